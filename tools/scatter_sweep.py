"""Scatter-kernel tuning sweep (one GPU): write group x columns per launch, 100 M rows x 8 columns."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from fugue_b200 import kernels as K

def main():
    n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 100_000_000
    dev = torch.device("cuda", 0)
    g = torch.Generator(device=dev).manual_seed(0)
    cols = [torch.randint(0, 1 << 16, (n,), dtype=torch.int64, device=dev, generator=g)]
    cols += [torch.randint(-(2**62), 2**62, (n,), dtype=torch.int64, device=dev, generator=g) for _ in range(3)]
    cols += [torch.randn(n, dtype=torch.float64, device=dev, generator=g) for _ in range(4)]
    outs = [torch.empty_like(c) for c in cols]
    plan = K.partition_plan([cols[0]], 256)
    ref = None
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for wg in (4,):
        for cpl in (4, 3, 2, 8):
            for _ in range(2):
                K.partition_apply(plan, cols, outs, cols_per_launch=cpl)
            torch.cuda.synchronize()
            e0.record()
            for _ in range(5):
                K.partition_apply(plan, cols, outs, cols_per_launch=cpl)
            e1.record(); torch.cuda.synchronize()
            chk = [int(o.view(torch.int64).sum().item()) for o in outs] + [int(outs[1][12345].item()), int(outs[5].view(torch.int64)[n - 7].item())]
            if ref is None:
                ref = chk
            print(json.dumps({"write_group": wg, "cols_per_launch": cpl, "ms": round(e0.elapsed_time(e1) / 5, 3),
                              "same_output": chk == ref}), flush=True)
    e0.record()
    for _ in range(5):
        K.partition_plan([cols[0]], 256, scratch=plan.scratch, offsets=plan.offsets)
    e1.record(); torch.cuda.synchronize()
    print(json.dumps({"pass1_ms": round(e0.elapsed_time(e1) / 5, 3)}))

if __name__ == "__main__":
    main()
