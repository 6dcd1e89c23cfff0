"""Probe: pass 2 (write-combined scatter) with its OUTPUT in the peer's memory (P2P stores over NVLink)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
import torch.distributed._symmetric_memory as symm_mem
from fugue_b200 import kernels as K
rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr); dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev)
n = int(float(os.environ.get("FB_ROWS", "1e8"))); ncols = 8; num = 256
arena = symm_mem.empty(n * ncols, dtype=torch.int64, device=dev)
hdl = symm_mem.rendezvous(arena, group=dist.group.WORLD)
g = torch.Generator(device=dev).manual_seed(rank)
key = torch.randint(0, 1 << 16, (n,), dtype=torch.int64, device=dev, generator=g)
cols = [key] + [torch.randint(-(2**62), 2**62, (n,), dtype=torch.int64, device=dev, generator=g) for _ in range(ncols - 1)]
scratch = torch.empty(K.partition_scratch_bytes(dev, n, num) + 256, dtype=torch.uint8, device=dev)
plan = K.partition_plan([key], num, scratch=scratch)
peer = (rank + 1) % world
def views(r):
    buf = hdl.get_buffer(r, (n * ncols,), torch.int64)
    return [buf[c * n:(c + 1) * n] for c in range(ncols)]
local_out, peer_out = views(rank), views(peer)
ref = [torch.empty_like(c) for c in cols]
K.partition_apply(plan, cols, ref)
def timeit(outs, label):
    ts = []
    for it in range(5):
        torch.cuda.synchronize(); dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); K.partition_apply(plan, cols, outs); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    dist.barrier()
    if rank == 0:
        print(f"{label}: {min(ts):.3f} ms  ({n*ncols*8/min(ts)/1e6:.0f} GB/s written)", flush=True)
timeit(local_out, "scatter -> own arena ")
timeit(peer_out, "scatter -> peer arena")
# correctness of the remote write: my arena now holds what the peer pushed = the peer's partitioned table
torch.cuda.synchronize(); dist.barrier()
mine = views(rank)
g2 = torch.Generator(device=dev).manual_seed((rank - 1) % world)
pk = torch.randint(0, 1 << 16, (n,), dtype=torch.int64, device=dev, generator=g2)
pp = K.partition_plan([pk], num, scratch=scratch)
chk = torch.empty_like(pk); K.partition_apply(pp, [pk], [chk])
ok = bool((mine[0] == chk).all())
print(f"rank {rank}: pushed key column matches: {ok}", flush=True)
# half local, half remote (what a 2-GPU exchange does): columns 0-3 local, 4-7 remote as a crude stand-in
mixed = local_out[:4] + peer_out[4:]
timeit(mixed, "scatter -> 4 cols own + 4 cols peer")
dist.destroy_process_group()
