#!/usr/bin/env python
"""MV_Aggregate (K6) latency / bandwidth sweep: one-shot P2P, two-shot P2P, NVLS (multimem) and the
NCCL all_reduce comparator, fp32 SUM in place, 4 B ... 256 MB, device-timed, max over ranks
(BASELINE.md section 5 row "Allreduce (one-shot P2P) 2/4/8, 4 B ... 1 MB").

    torchrun --nproc-per-node 8 bench/allreduce.py
Writes gpurun_out/allreduce_n<N>.json."""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import multiverso_b200 as mv
from multiverso_b200.parallel import aggregate


def timed(fn, iters, world):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    mv.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    us = torch.tensor([e0.elapsed_time(e1) / iters * 1e3], dtype=torch.float64, device="cuda")
    if world > 1:
        import torch.distributed as dist
        dist.all_reduce(us, op=dist.ReduceOp.MAX)
    mv.barrier()
    return float(us)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--max-mb", type=int, default=256)
    a = ap.parse_args()
    mv.init()
    world, rank = mv.size(), mv.rank()
    if world == 1:
        print(json.dumps({"skipped": "world size 1"}))
        mv.shutdown()
        return
    import torch.distributed as dist
    sizes = [4, 64, 1024, 16 * 1024, 128 * 1024, 1 << 20, 8 << 20, 64 << 20, a.max_mb << 20]
    out = []
    for nbytes in sizes:
        n = nbytes // 4
        x = torch.ones(n, device="cuda")
        iters = 200 if nbytes <= (1 << 20) else 20
        row = {"bytes": nbytes}
        for algo in ("fused", "oneshot", "twoshot", "nvls"):
            if algo == "oneshot" and nbytes > (8 << 20):
                continue                                   # every rank reads everything: pointless for big buffers
            if algo == "fused" and nbytes > (1 << 20):
                continue                                   # the single-launch latency path covers <= 1 MB
            try:
                row[algo + "_us"] = timed(lambda: aggregate(x, algo=algo), iters, world)
            except Exception as e:                         # e.g. no multicast support
                row[algo + "_error"] = str(e)[:80]
            x.fill_(1.0)
        if nbytes > (1 << 20):
            # the same reduction on a tensor that already lives in symmetric memory: no staging copy
            xs = mv.symm_tensor(n, "float32")
            xs.fill_(1.0)
            row["inplace_us"] = timed(lambda: aggregate(xs, algo="auto"), iters, world)
            row["inplace_twoshot_us"] = timed(lambda: aggregate(xs, algo="twoshot"), iters, world)
        row["nccl_us"] = timed(lambda: dist.all_reduce(x), iters, world)    # default group: cuda -> NCCL
        best = min(v for k, v in row.items() if k.endswith("_us") and not k.startswith("nccl"))
        # bus bandwidth convention: 2 (W-1)/W * bytes / time
        row["best_ours_us"] = best
        row["best_busbw_gbs"] = 2.0 * (world - 1) / world * nbytes / best / 1e3
        out.append(row)
    if rank == 0:
        os.makedirs("gpurun_out", exist_ok=True)
        with open(f"gpurun_out/allreduce_n{world}.json", "w") as f:
            json.dump(out, f, indent=1)
        print(json.dumps(out), flush=True)
    mv.shutdown()


if __name__ == "__main__":
    main()
