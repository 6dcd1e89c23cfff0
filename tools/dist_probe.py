import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
from fugue_b200 import kernels as K
from fugue_b200.dist import ExchangePlan, gather_counts
rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr); dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev)
n, num = 100_000_000, 256
g = torch.Generator(device=dev).manual_seed(rank)
cols = [torch.randint(0, 1 << 16, (n,), dtype=torch.int64, device=dev, generator=g)] + [torch.randint(-(2**62), 2**62, (n,), dtype=torch.int64, device=dev, generator=g) for _ in range(7)]
def T():
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize(); return time.perf_counter()
plan = K.partition_plan([cols[0]], num)
parts = K.partition_apply(plan, cols)
counts = gather_counts(plan.offsets[1:] - plan.offsets[:-1])
ep = ExchangePlan(counts, rank)
recv = [torch.empty(ep.total_recv, dtype=torch.int64, device=dev) for _ in cols]
outs = [torch.empty(ep.total_recv, dtype=torch.int64, device=dev) for _ in cols]
ss, sd, sl = ep.seg_src_off.to(dev), ep.seg_dst_off.to(dev), ep.seg_len.to(dev)
for it in range(3):
    t0 = T(); plan = K.partition_plan([cols[0]], num, scratch=plan.scratch, offsets=plan.offsets)
    t1 = T(); K.partition_apply(plan, cols, parts)
    t2 = T()
    for c in cols: K.partition_apply(plan, [c], [parts[0]])
    t3 = T(); counts = gather_counts(plan.offsets[1:] - plan.offsets[:-1]); ep = ExchangePlan(counts, rank)
    t4 = T()
    for p, r in zip(parts, recv): dist.all_to_all_single(r, p, output_split_sizes=ep.recv_rows, input_split_sizes=ep.send_rows)
    t5 = T(); K.copy_segments(recv, outs, ss, sd, sl)
    t6 = T()
    for r, o in zip(recv, outs): K.copy_segments([r], [o], ss, sd, sl)
    t7 = T()
    if rank == 0:
        print(f"plan {1e3*(t1-t0):.2f} | apply8 {1e3*(t2-t1):.2f} | 8x apply1 {1e3*(t3-t2):.2f} | counts+plan {1e3*(t4-t3):.2f} | a2a x8 {1e3*(t5-t4):.2f} ({8*ep.send_rows[1-rank if world==2 else 0]*8/ (t5-t4)/1e9:.0f} GB/s out to one peer) | segcopy8 {1e3*(t6-t5):.2f} | 8x segcopy1 {1e3*(t7-t6):.2f}", flush=True)
dist.destroy_process_group()
