"""All-to-all bandwidth probe over symmetric memory (torchrun, one rank per GPU).  Every rank receives
`--mb` MiB from each peer; variants: copy-engine pull / push with k streams, the SM pull kernel
(fb_copy_segments), NCCL all_to_all_single.  Prints GB/s received per GPU (max time over ranks)."""
import argparse
import json
import os
import sys

os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mb", type=int, default=1000)
    ap.add_argument("--reps", type=int, default=3)
    args = ap.parse_args()
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local_rank = int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist.init_process_group("nccl", device_id=dev)
    import torch.distributed._symmetric_memory as symm_mem

    from fugue_b200 import kernels as K

    nb = args.mb << 20
    src = symm_mem.empty(world * nb, dtype=torch.uint8, device=dev)     # slice d = what rank d pulls from me
    hdl = symm_mem.rendezvous(src, group=dist.group.WORLD)
    dst_sym = symm_mem.empty(world * nb, dtype=torch.uint8, device=dev)  # push target: slice s = from rank s
    hdl2 = symm_mem.rendezvous(dst_sym, group=dist.group.WORLD)
    src.fill_(rank + 1)
    dst = torch.empty(world * nb, dtype=torch.uint8, device=dev)
    base = [int(x) for x in hdl.buffer_ptrs]
    base2 = [int(x) for x in hdl2.buffer_ptrs]
    streams = [torch.cuda.Stream(dev) for _ in range(world)]
    hi = torch.cuda.Stream(dev, priority=-1)
    peers = [(rank + j) % world for j in range(1, world)]

    def sync():
        torch.cuda.synchronize(dev)
        dist.barrier()
        torch.cuda.synchronize(dev)

    def timed(name, fn, extra=None):
        fn()
        sync()
        best = 1e9
        for _ in range(args.reps):
            sync()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            torch.cuda.synchronize(dev)
            t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            best = min(best, float(t.item()))
        if rank == 0:
            out = {"variant": name, "ms": round(best, 3), "GBps_in_per_gpu": round((world - 1) * nb / best / 1e6, 1)}
            if extra:
                out.update(extra)
            print(json.dumps(out), flush=True)

    main_s = torch.cuda.current_stream(dev)

    def ce(nstreams, push, pieces=1):
        def fn():
            ev = torch.cuda.Event()
            ev.record(main_s)
            used = []
            for j, p in enumerate(peers):
                st = streams[j % nstreams]
                used.append(st)
                with torch.cuda.stream(st):
                    st.wait_event(ev)
                    for q in range(pieces):
                        o, ln = q * (nb // pieces), nb // pieces
                        if push:  # write my slice for p into p's buffer
                            K.copy_runs_dma(dev, [base[rank] + p * nb + o], [base2[p] + rank * nb + o], [ln])
                        else:
                            K.copy_runs_dma(dev, [base[p] + rank * nb + o], [dst.data_ptr() + p * nb + o], [ln])
            for st in set(used):
                main_s.wait_stream(st)
        return fn

    for k in (1, 2, 4, world - 1):
        timed(f"ce_pull_{k}streams", ce(k, False))
    timed(f"ce_pull_{world - 1}streams_8pieces", ce(world - 1, False, 8))
    for k in (1, world - 1):
        timed(f"ce_push_{k}streams", ce(k, True))
    # SM pull kernel: world-1 segments of nb/8 rows
    rows = nb // 8
    seg_src = torch.full((world - 1,), rank * rows, dtype=torch.int64, device=dev)
    seg_dst = torch.tensor([p * rows for p in peers], dtype=torch.int64, device=dev)
    seg_len = torch.full((world - 1,), rows, dtype=torch.int64, device=dev)
    seg_tab = torch.tensor(peers, dtype=torch.int32, device=dev)
    d64 = dst.view(torch.int64)
    timed("sm_pull_kernel_ld128", lambda: K.copy_segments(None, [d64], seg_src, seg_dst, seg_len, max_len=rows,
                                                          src_table=seg_tab, src_ptrs=base))
    if hasattr(K, "pull_runs_tma"):
        for ctas in (8, 16, 32, 64):
            timed(f"tma_pull_{ctas}ctas", lambda: K.pull_runs_tma(dev, [base[p] + rank * nb for p in peers],
                                                                   [dst.data_ptr() + p * nb for p in peers],
                                                                   [nb] * (world - 1), ctas))
    inp = src[:world * nb]
    timed("nccl_all_to_all_single", lambda: dist.all_to_all_single(dst, inp))
    # local reference: device-to-device copy of the same bytes on the copy engine and by a kernel
    timed("local_ce_copy", lambda: K.copy_runs_dma(dev, [src.data_ptr()], [dst.data_ptr()], [(world - 1) * nb]))
    timed("local_torch_copy", lambda: dst[:(world - 1) * nb].copy_(src[:(world - 1) * nb]))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
