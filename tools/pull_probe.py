import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
import torch.distributed._symmetric_memory as symm_mem
from fugue_b200 import kernels as K
rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr); dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev)
n = 100_000_000; ncols = 8
t = symm_mem.empty(n * ncols, dtype=torch.int64, device=dev)
hdl = symm_mem.rendezvous(t, group=dist.group.WORLD)
t.fill_(rank + 1)
torch.cuda.synchronize(); dist.barrier()
peer = (rank + 1) % world
# emulate the exchange at G=8: 256/8=32 owned partitions x 8 sources -> segments of ~48K rows, all from the peer
nseg = 256; seglen = n // 8 // nseg * 1  # 48828 rows
ln = torch.full((nseg,), seglen, dtype=torch.int64, device=dev)
off = torch.arange(nseg, dtype=torch.int64, device=dev) * seglen + 3   # odd offsets: 8-byte alignment only
doff = torch.arange(nseg, dtype=torch.int64, device=dev) * seglen
outs = [torch.empty(nseg * seglen, dtype=torch.int64, device=dev) for _ in range(ncols)]
src_ptrs = [int(hdl.buffer_ptrs[peer]) + c * n * 8 for c in range(ncols)]
tab = torch.zeros(nseg, dtype=torch.int32, device=dev)
tot = nseg * seglen * 8 * ncols
for variant in (0,):
    for pieces in ("", "4", "16"):
        os.environ["FB_COPY_VARIANT"] = str(variant)
        if pieces: os.environ["FB_COPY_PIECES"] = pieces
        else: os.environ.pop("FB_COPY_PIECES", None)
        ts = []
        for it in range(4):
            torch.cuda.synchronize(); dist.barrier(); t0 = time.perf_counter()
            K.copy_segments(None, outs, off, doff, ln, max_len=seglen, src_table=tab, src_ptrs=src_ptrs)
            torch.cuda.synchronize(); ts.append(time.perf_counter() - t0)
        if rank == 0: print(f"variant {variant} pieces {pieces or 'auto'}: {tot/min(ts)/1e9:.0f} GB/s ({min(ts)*1e3:.2f} ms for {tot/1e9:.2f} GB)", flush=True)
dist.destroy_process_group()
