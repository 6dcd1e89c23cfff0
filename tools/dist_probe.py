"""Multi-GPU exchange probe (torchrun, one rank per GPU): parity check, then the headline transform
step for a sweep of exchange settings on one engine object.  Prints one line per setting on rank 0.
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port 29544 tools/dist_probe.py [--rows 125000000] [--sweep "dma:1-1-2-4:1,dma:2:1,kernel:4:16"]
"""
import argparse
import json
import os

os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")  # see fugue_b200/dist.py
import sys
import time

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=125_000_000)
    ap.add_argument("--sweep", default="dma:2:1,dma:4:1,dma:1:1,dma:8:1,kernel:4:16,kernel:2:16")
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--no-parity", action="store_true")
    ap.add_argument("--trace", action="store_true")
    args = ap.parse_args()
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local_rank = int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist.init_process_group("nccl", device_id=dev)
    from fugue_b200 import api as fa
    from fugue_b200.dataframe import B200DataFrame
    from fugue_b200.dist import DistributedB200Engine
    from fugue_b200.partition import PartitionSpec
    from fugue_b200.table import B200Table

    eng = DistributedB200Engine({"fugue.b200.device": local_rank})
    if not args.no_parity:
        import dist_gpu_check

        t0 = time.time()
        ok, msg = dist_gpu_check.run_checks(eng, rank, world, dev)
        if rank == 0:
            print(f"PARITY {ok} {msg} ({time.time() - t0:.1f}s)", flush=True)
    n = args.rows
    g = torch.Generator(device=dev).manual_seed(rank)
    cols = [torch.randint(0, 1 << 16, (n,), dtype=torch.int64, device=dev, generator=g)]
    cols += [torch.randint(-(2**62), 2**62, (n,), dtype=torch.int64, device=dev, generator=g) for _ in range(3)]
    cols += [torch.randn(n, dtype=torch.float64, device=dev, generator=g) for _ in range(4)]
    df = B200DataFrame(B200Table("key:long,i1:long,i2:long,i3:long,v0:double,v1:double,v2:double,v3:double", cols))
    spec = PartitionSpec(by="key", algo="hash", num=256)

    def identity(t: B200Table) -> B200Table:
        return t

    def sync():
        torch.cuda.synchronize(dev)
        dist.barrier()
        torch.cuda.synchronize(dev)

    for item in args.sweep.split(","):
        mode, gc, res = item.split(":")[:3]
        if len(item.split(":")) > 3:
            eng._dma_pieces = int(item.split(":")[3])
        if len(item.split(":")) > 4:
            eng._dma_overlap_flag = bool(int(item.split(":")[4]))
        eng._exchange, eng._group_cols, eng._sm_reserve = mode, [int(x) for x in gc.split('-')], int(res)
        try:
            for _ in range(2):
                out = fa.transform(df, identity, schema="*", partition=spec, engine=eng)
            sync()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            t0 = time.perf_counter()
            for _ in range(args.steps):
                out = fa.transform(df, identity, schema="*", partition=spec, engine=eng)
            host_ms = (time.perf_counter() - t0) * 1e3 / args.steps
            e1.record()
            sync()
            ms = e0.elapsed_time(e1) / args.steps
            t = torch.tensor([ms], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            rows_out = torch.tensor([out.count()], dtype=torch.int64, device=dev)
            dist.all_reduce(rows_out)
            if rank == 0:
                ms = float(t.item())
                print(json.dumps({"mode": mode, "group_cols": gc, "sm_reserve": int(res), "dma_pieces": eng._dma_pieces, "overlap_flag": eng._dma_overlap_flag, "world": world,
                                  "rows_per_gpu": n, "ms_per_step": round(ms, 3), "host_enqueue_ms": round(host_ms, 3),
                                  "G_rows_per_s": round(n * world / ms / 1e6, 2), "rows_out": int(rows_out.item()),
                                  "nvlink_GBps_out_per_gpu": round(64.0 * n * (world - 1) / world / ms / 1e6, 1)}),
                      flush=True)
            if args.trace:
                eng._trace = []
                out = fa.transform(df, identity, schema="*", partition=spec, engine=eng)
                sync()
                tr, eng._trace = eng._trace, None
                if rank in (0, world - 1):
                    t0e = tr[0][1]
                    print(f"[rank {rank}] trace {item}: " + " ".join(f"{lb}={t0e.elapsed_time(ev):.2f}" for lb, ev in tr[1:]),
                          flush=True)
            del out
        except Exception as e:  # noqa: BLE001
            print(f"[rank {rank}] {item}: {e!r}", flush=True)
            break
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
