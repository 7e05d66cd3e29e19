#!/usr/bin/env python
"""Device counterpart of the reference's manual perf test (Test/test_matrix_perf.cpp:32-171,
SURVEY P11): MatrixTable 1,000,000 x 50 fp32 (200 MB); for p = 10 % .. 100 %: Get all rows ->
every worker Adds its share of the first p of the rows -> Get all rows again, verified and timed
(device-timed, max over ranks).  `--sparse` uses the stale-row delta pull of the sparse Matrix
(the second Get only moves the rows that changed, matrix.cpp:460-514).

    python bench/matrix_perf.py [--rows N] [--sparse]
    torchrun --nproc-per-node 8 bench/matrix_perf.py
Writes gpurun_out/matrix_perf_n<N>[_sparse].json."""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import multiverso_b200 as mv
from multiverso_b200.tables.options import GetOption


def dev_ms(fn, world):
    torch.cuda.synchronize()
    mv.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    out = fn()
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device="cuda")
    if world > 1:
        import torch.distributed as dist
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    return float(ms), out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=1_000_000)
    ap.add_argument("--cols", type=int, default=50)
    ap.add_argument("--sparse", action="store_true")
    a = ap.parse_args()
    mv.init()
    world, rank = mv.size(), mv.rank()
    R, C = a.rows, a.cols
    nbytes = R * C * 4
    dev = torch.device("cuda", torch.cuda.current_device())
    rows_all = torch.arange(R, device=dev)
    delta_full = (torch.arange(R * C, device=dev, dtype=torch.float32) % 100003).view(R, C)
    out, ok_all = [], True
    for percent in range(10):
        t = mv.MatrixTable(R, C, "float32", init_value=0.0, is_sparse=a.sparse)
        opt = GetOption(worker_id=mv.worker_id())
        buf = torch.empty(R * C, device=dev)
        first_ms, _ = dev_ms(lambda: t.get(buf), world)
        if a.sparse:
            t.get_stale(opt)                                   # consume the initial "everything is stale" state
        mine = rows_all[(rows_all % 10 <= percent) & (rows_all % world == rank)]
        add_ms, _ = dev_ms(lambda: t.add_rows(mine, delta_full[mine]) if mine.numel() else None, world)
        mv.barrier()
        if a.sparse:
            get_ms, (ids, vals) = dev_ms(lambda: t.get_stale(opt), world)
            full = torch.zeros(R, C, device=dev)
            full[ids] = vals
            moved = int(ids.numel())
        else:
            get_ms, _ = dev_ms(lambda: t.get(buf), world)
            full, moved = buf.view(R, C), R
        expect = torch.where((rows_all % 10 <= percent)[:, None], delta_full, torch.zeros_like(delta_full))
        ok = bool(torch.equal(full, expect))
        ok_all = ok_all and ok
        out.append({"percent": (percent + 1) * 10, "first_get_ms": first_ms, "add_ms": add_ms,
                    "rows_added_per_worker": int(mine.numel()), "get_ms": get_ms, "rows_moved_by_get": moved,
                    "get_gbs": moved * C * 4 / get_ms / 1e6, "verified": ok})
        mv.barrier()
        t.free()
    if rank == 0:
        os.makedirs("gpurun_out", exist_ok=True)
        name = f"gpurun_out/matrix_perf_n{world}{'_sparse' if a.sparse else ''}.json"
        res = {"table": f"{R}x{C} fp32 ({nbytes / 1e6:.0f} MB)", "sparse": a.sparse, "n_gpus": world,
               "all_verified": ok_all, "turns": out}
        with open(name, "w") as f:
            json.dump(res, f, indent=1)
        print(json.dumps(res), flush=True)
    mv.shutdown()
    sys.exit(0 if ok_all else 1)


if __name__ == "__main__":
    main()
