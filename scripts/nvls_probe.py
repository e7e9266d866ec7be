"""Where the gradient all-reduce's time goes on this fabric.  CUDA-event timings, max over ranks:
the two signal-pad barriers alone, the two-shot kernel alone for several CTA counts, the full barrier+kernel+barrier
sequence, NCCL's all-reduce of the same 23.6 MB buffer, and torch's own symmetric-memory all-reduces.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 scripts/nvls_probe.py"""
import ctypes as C
import json
import os
import sys

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
import torch
import torch.distributed as dist
import torch.distributed._symmetric_memory as symm_mem

from gaussianavatars_b200 import _native as N


def main():
    rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(lr)
    dev = torch.device("cuda", lr)
    dist.init_process_group("nccl", device_id=dev)
    n = 100_000 * 59
    try:
        symm_mem.set_backend("CUDA")
    except Exception:
        pass
    buf = symm_mem.empty(n, dtype=torch.float32, device=dev)
    h = symm_mem.rendezvous(buf, dist.group.WORLD)
    mc = int(h.multicast_ptr)
    plain = torch.zeros(n, device=dev)
    buf.zero_()
    lib = N.lib()
    iters = 50

    def kernel():
        stream = torch.cuda.current_stream(dev).cuda_stream
        N.check(lib.gab200_nvls_allreduce(C.c_void_p(mc), n, rank, world, C.c_void_p(stream)), "nvls")

    def timed(fn, graph=False):
        for _ in range(5):
            fn()
        torch.cuda.synchronize()
        run = fn
        if graph:
            g = torch.cuda.CUDAGraph()
            s = torch.cuda.Stream()
            s.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(s):
                with torch.cuda.graph(g, stream=s):
                    fn()
            run = g.replay
            run()
        dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            run()
        e1.record()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / iters * 1e3], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return round(float(t), 2)

    out = {"world": world, "floats": n, "unit": "us"}
    out["barrier_x2"] = timed(lambda: (h.barrier(channel=0), h.barrier(channel=1)))
    out["barrier_x1"] = timed(lambda: h.barrier(channel=0))
    for ctas in (32, 64, 148, 296, 592, 1184):
        N.tune(N.TUNE_NVLS_CTAS, ctas)
        out[f"kernel_{ctas}"] = timed(lambda: (h.barrier(channel=0), kernel())) - out["barrier_x1"]
    N.tune(N.TUNE_NVLS_CTAS, 0)
    out["two_shot_full"] = timed(lambda: (h.barrier(channel=0), kernel(), h.barrier(channel=1)))
    out["nccl"] = timed(lambda: dist.all_reduce(plain))
    for name in ("multimem_all_reduce_", "two_shot_all_reduce_", "one_shot_all_reduce"):
        try:
            op = getattr(torch.ops.symm_mem, name)
            gname = dist.group.WORLD.group_name
            out["torch_" + name] = timed(lambda: op(buf, "sum", gname))
        except Exception as e:
            out["torch_" + name] = f"{type(e).__name__}: {str(e)[:80]}"
    # the same inside CUDA graphs (what the step does)
    try:
        out["two_shot_full_graph"] = timed(lambda: (h.barrier(channel=0), kernel(), h.barrier(channel=1)), graph=True)
    except Exception as e:
        out["two_shot_full_graph"] = f"{type(e).__name__}: {str(e)[:80]}"
    try:
        out["nccl_graph"] = timed(lambda: dist.all_reduce(plain), graph=True)
    except Exception as e:
        out["nccl_graph"] = f"{type(e).__name__}: {str(e)[:80]}"
    if rank == 0:
        print(json.dumps(out), flush=True)
    torch.cuda.synchronize()
    dist.barrier()
    os._exit(0)


if __name__ == "__main__":
    main()
