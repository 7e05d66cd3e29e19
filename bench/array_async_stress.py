#!/usr/bin/env python
"""BASELINE.json config 5: ArrayTable 4 GiB dense fp32, ASYNC Get/Add with the Momentum server updater,
staleness stress -- workers run at deliberately different speeds, every Add records how many Adds of OTHER
workers were applied to a shard between this worker's last Get and its own Add.

    python -m torch.distributed.run --nproc-per-node 8 ... bench/array_async_stress.py [--gb 4] [--iters 12]

Reports, per rank and aggregated: device-timed Get / Add GB/s (table bytes / time, max over ranks) and the
staleness distribution (Dashboard.staleness()).  Reference semantics: async Server (src/server.cpp:36-58) with
`-updater_type=momentum_sgd` (include/multiverso/updater/momentum_updater.h:17-25); the reference has no
staleness measurement at all."""
import argparse
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import multiverso_b200 as mv  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gb", type=float, default=4.0)
    ap.add_argument("--iters", type=int, default=12)
    ap.add_argument("--skew-ms", type=float, default=3.0, help="extra compute per iteration = rank * skew")
    args = ap.parse_args()
    mv.init(["-sync=false", "-updater_type=momentum_sgd", "-staleness=true"])
    r, W = mv.rank(), mv.size()
    n = int(args.gb * (1 << 30)) // 4
    t = mv.ArrayTable(n, "float32", updater="momentum_sgd")
    delta = torch.full((n,), 1e-3, device="cuda")
    buf = torch.empty(n, device="cuda")
    opt = mv.AddOption(momentum=0.9)
    t.get(buf); t.add(delta, opt); torch.cuda.synchronize(); mv.barrier()
    mv.Dashboard.reset()
    get_ms, add_ms = [], []
    t0 = time.time()
    # worker r does (W - r) ... no: every worker runs `iters` rounds, slower ranks take longer per round
    for it in range(args.iters):
        e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        e0.record()
        t.get(buf)
        e1.record()
        torch.cuda._sleep(int(1.9e6 * args.skew_ms * r))          # uneven "compute" (cycles at ~1.9 GHz)
        e2.record()
        t.add(delta, opt)
        e3 = torch.cuda.Event(enable_timing=True); e3.record()
        torch.cuda.synchronize()
        get_ms.append(e0.elapsed_time(e1)); add_ms.append(e2.elapsed_time(e3))
    wall = time.time() - t0
    mv.barrier()
    st = mv.Dashboard.staleness()
    mine = next(iter(st.values())) if st else {"adds": 0, "hist": [0] * 64}
    hist = torch.tensor(mine["hist"], dtype=torch.float64, device="cuda")
    per = torch.tensor([sum(get_ms) / len(get_ms), sum(add_ms) / len(add_ms)], dtype=torch.float64, device="cuda")
    hist_all = hist.clone(); mv.aggregate(hist_all)
    worst = per.clone()
    if W > 1:
        import torch.distributed as dist
        dist.all_reduce(worst, op=dist.ReduceOp.MAX)
    nbytes = n * 4
    row = {"rank": r, "get_ms": round(float(per[0]), 3), "add_ms": round(float(per[1]), 3),
           "staleness_mean": mine.get("mean"), "staleness_p50": mine.get("p50"), "staleness_p99": mine.get("p99"),
           "adds": mine["adds"], "wall_s": round(wall, 2)}
    print(json.dumps(row), flush=True)
    mv.barrier()
    if r == 0:
        h = [int(v) for v in hist_all.tolist()]
        tot = max(sum(h), 1)
        cum, p50, p99 = 0, None, None
        for k, v in enumerate(h):
            cum += v
            if p50 is None and cum >= 0.5 * tot: p50 = k
            if p99 is None and cum >= 0.99 * tot: p99 = k
        print(json.dumps({"bench": "array_async_stress", "n_gpus": W, "table_gib": args.gb, "updater": "momentum_sgd",
                          "mode": "async", "iters": args.iters, "skew_ms_per_rank": args.skew_ms,
                          "get_ms_max": round(float(worst[0]), 3), "add_ms_max": round(float(worst[1]), 3),
                          "get_gbs": round(nbytes / float(worst[0]) / 1e6, 1), "add_gbs": round(nbytes / float(worst[1]) / 1e6, 1),
                          "staleness": {"adds": tot, "mean": round(sum(k * v for k, v in enumerate(h)) / tot, 3),
                                        "p50": p50, "p99": p99, "hist": h[:32]}}), flush=True)
    mv.shutdown()


if __name__ == "__main__":
    main()
