"""Does a copy-engine P2P copy make progress next to a compute kernel?  (torchrun, 2 ranks)
Measures the duration of a 2 GB peer pull alone and while different kernels run on the main stream."""
import json
import os
import sys

os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local_rank = int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist.init_process_group("nccl", device_id=dev)
    import torch.distributed._symmetric_memory as symm_mem

    from fugue_b200 import kernels as K

    nb = 2 << 30
    src = symm_mem.empty(nb, dtype=torch.uint8, device=dev)
    hdl = symm_mem.rendezvous(src, group=dist.group.WORLD)
    src.fill_(rank + 1)
    dst = torch.empty(nb, dtype=torch.uint8, device=dev)
    base = [int(x) for x in hdl.buffer_ptrs]
    peer = (rank + 1) % world
    n = 100_000_000
    g = torch.Generator(device=dev).manual_seed(rank)
    cols = [torch.randint(0, 1 << 16, (n,), dtype=torch.int64, device=dev, generator=g)]
    cols += [torch.randn(n, dtype=torch.float64, device=dev, generator=g).view(torch.int64) for _ in range(3)]
    outs = [torch.empty_like(c) for c in cols]
    plan = K.partition_plan([cols[0]], 256)
    big = torch.empty(400_000_000, dtype=torch.float64, device=dev)
    big2 = torch.empty_like(big)
    s_copy = torch.cuda.Stream(dev)
    main_s = torch.cuda.current_stream(dev)

    def sync():
        torch.cuda.synchronize(dev)
        dist.barrier()
        torch.cuda.synchronize(dev)

    def run(name, kernel_fn, reps):
        sync()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        k0.record(main_s)
        if kernel_fn is not None:
            for _ in range(reps):
                kernel_fn()
        k1.record(main_s)
        with torch.cuda.stream(s_copy):
            e0.record(s_copy)
            K.copy_runs_dma(dev, [base[peer]], [dst.data_ptr()], [nb])
            e1.record(s_copy)
        sync()
        if rank == 0:
            print(json.dumps({"next_to": name, "copy_ms": round(e0.elapsed_time(e1), 3),
                              "copy_GBps": round(nb / e0.elapsed_time(e1) / 1e6, 1),
                              "kernel_ms": round(k0.elapsed_time(k1), 3)}), flush=True)

    for _ in range(2):
        run("nothing", None, 0)
    run("scatter (persistent, 1 CTA/SM, 4 columns x 4)", lambda: K.partition_apply(plan, cols, outs), 4)
    run("scatter with 32 SMs reserved", lambda: K.partition_apply(plan, cols, outs, sm_reserve=32), 4)
    run("rank kernel (pass 1, ALU-bound) x 12", lambda: K.partition_plan([cols[0]], 256, scratch=plan.scratch, offsets=plan.offsets), 12)
    run("torch copy kernel 3.2 GB x 6", lambda: big2.copy_(big), 6)
    run("torch fill kernel 3.2 GB x 8", lambda: big.fill_(1.0), 8)
    run("torch sum (read-only) 3.2 GB x 8", lambda: big.sum(), 8)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
