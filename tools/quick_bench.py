"""Developer timing loop for the partition kernels (not the contract bench)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from fugue_b200 import kernels as K

def main():
    n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 100_000_000
    num = int(sys.argv[2]) if len(sys.argv) > 2 else 256
    dev = torch.device("cuda", 0)
    g = torch.Generator(device=dev).manual_seed(0)
    key = torch.randint(0, 1 << 16, (n,), dtype=torch.int64, device=dev, generator=g)
    pattern = os.environ.get("FB_PATTERN", "")
    if pattern:
        # aligned-run experiment: every tile holds exactly `rep` consecutive... rows per partition
        reps = [int(x) for x in pattern.split(",")]
        rep = reps[0]
        cand = torch.arange(0, 1 << 16, dtype=torch.int64, device=dev)
        pid = K.partition_ids([cand], num).long()
        first = torch.full((num,), -1, dtype=torch.int64, device=dev)
        # pick one key per partition id
        order = torch.argsort(pid, stable=True)
        sp = pid[order]
        starts = torch.searchsorted(sp, torch.arange(num, device=dev))
        keys_per_pid = cand[order][starts]
        if len(reps) == 1:
            pat = keys_per_pid.repeat_interleave(rep)          # rep rows of pid 0, rep rows of pid 1, ...
        else:  # first half of the partitions gets reps[0] rows per tile, second half reps[1]
            cnt = torch.tensor([reps[0]] * (num // 2) + [reps[1]] * (num - num // 2), device=dev)
            pat = keys_per_pid.repeat_interleave(cnt)
        perm = torch.randperm(pat.numel(), device=dev, generator=g) if os.environ.get("FB_SHUFFLE") else None
        if perm is not None: pat = pat[perm]
        key = pat.repeat((n + pat.numel() - 1) // pat.numel())[:n].contiguous()
    cols = [key] + [torch.randint(-(2**62), 2**62, (n,), dtype=torch.int64, device=dev, generator=g) for _ in range(3)] \
        + [torch.randn(n, dtype=torch.float64, device=dev, generator=g) for _ in range(4)]
    ncols = int(os.environ.get("FB_NCOLS", "8"))
    cols = cols[:ncols]
    out = [torch.empty_like(c) for c in cols]
    scratch = torch.empty(K.partition_scratch_bytes(dev, n, num) + 256, dtype=torch.uint8, device=dev)
    off = torch.empty(num + 1, dtype=torch.int64, device=dev)
    for _ in range(3):
        K.partition_columns(cols, [0], num, out=out, scratch=scratch, offsets=off)
    torch.cuda.synchronize()
    # copy peak on this box
    a = torch.empty(n * 4, dtype=torch.int64, device=dev); b = torch.empty_like(a)
    for _ in range(2): b.copy_(a)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best = 1e9
    for _ in range(5):
        e0.record(); b.copy_(a); e1.record(); torch.cuda.synchronize(); best = min(best, e0.elapsed_time(e1))
    print(f"copy: {2*a.numel()*8/best/1e6:.1f} GB/s")
    del a, b
    ts = []
    for _ in range(10):
        e0.record(); K.partition_columns(cols, [0], num, out=out, scratch=scratch, offsets=off); e1.record()
        torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    t = sorted(ts)[len(ts)//2]
    print(f"partition n={n} num={num}: median {t:.3f} ms  min {min(ts):.3f} ms  {n/t/1e6:.2f} Grows/s  alg {128*n/t/1e6:.1f} GB/s")
    # plan only
    ts = []
    for _ in range(5):
        e0.record(); plan = K.partition_plan([key], num, scratch=scratch, offsets=off); e1.record()
        torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    print(f"plan only: {min(ts):.3f} ms ({8*n/min(ts)/1e6:.1f} GB/s)")
    ts = []
    for _ in range(5):
        e0.record(); K.partition_apply(plan, cols, out); e1.record()
        torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    print(f"apply only: {min(ts):.3f} ms ({16*len(cols)*n/min(ts)/1e6:.1f} GB/s alg, {len(cols)} cols)")

main()
