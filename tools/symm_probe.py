import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
import torch.distributed._symmetric_memory as symm_mem
from fugue_b200 import kernels as K
rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr); dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev)
n = 100_000_000
t = symm_mem.empty(n, dtype=torch.int64, device=dev)
hdl = symm_mem.rendezvous(t, group=dist.group.WORLD)
print(rank, "ptrs", [hex(p) for p in hdl.buffer_ptrs], "world", hdl.world_size, flush=True)
t.fill_(rank + 1)
torch.cuda.synchronize(); dist.barrier()
peer = (rank + 1) % world
pt = hdl.get_buffer(peer, (n,), torch.int64)
print(rank, "peer first", int(pt[0]), flush=True)
out = torch.empty(n, dtype=torch.int64, device=dev)
# pull with my segment-copy kernel: 1024 segments
nseg = 1024
ln = torch.full((nseg,), n // nseg, dtype=torch.int64, device=dev)
off = torch.arange(nseg, dtype=torch.int64, device=dev) * (n // nseg)
for it in range(3):
    torch.cuda.synchronize(); dist.barrier(); t0 = time.perf_counter()
    K.copy_segments([pt], [out], off, off, ln)
    torch.cuda.synchronize(); t1 = time.perf_counter()
    out2 = out.clone(); torch.cuda.synchronize()
    t2 = time.perf_counter(); out.copy_(pt); torch.cuda.synchronize(); t3 = time.perf_counter()
    print(rank, f"pull via fb_copy_segments: {0.8/(t1-t0):.0f} GB/s; torch copy_ from peer: {0.8/(t3-t2):.0f} GB/s", flush=True)
assert int(out[123]) == peer + 1
dist.destroy_process_group()
