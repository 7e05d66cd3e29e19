#!/usr/bin/env python
"""BASELINE config 4: LogisticRegression, 10M-dim sparse input, server-side AdaGrad updater,
synthetic CSR minibatches (no network for datasets).  Measures samples/s through the model's
public step (K8 forward/backward + fused AdaGrad Add + pipelined Get), device-timed, max over ranks."""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import multiverso_b200 as mv
from multiverso_b200.models.logreg import LogRegConfig, LogRegModel


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--dim", type=int, default=10_000_000)
    ap.add_argument("--nnz", type=int, default=40)
    ap.add_argument("--batch", type=int, default=65536)
    ap.add_argument("--steps", type=int, default=20)
    a = ap.parse_args()
    mv.init(sync=True)
    world, rank = mv.size(), mv.rank()
    cfg = LogRegConfig(input_size=a.dim, output_size=1, sparse=True, objective_type="sigmoid", updater_type="sgd",
                       use_ps=True, pipeline=False, sync_frequency=1, learning_rate=0.05, server_updater="adagrad",
                       minibatch_size=a.batch, regular_type="default")
    model = LogRegModel(cfg)
    g = torch.Generator(device="cuda").manual_seed(1 + rank)
    batches = []
    for _ in range(4):
        keys = torch.randint(0, a.dim, (a.batch * a.nnz,), device="cuda", generator=g)
        row_ptr = torch.arange(0, a.batch * a.nnz + 1, a.nnz, device="cuda")
        vals = torch.ones(a.batch * a.nnz, device="cuda")
        labels = (torch.rand(a.batch, device="cuda", generator=g) > 0.5).float()
        batches.append((row_ptr, keys, vals, labels))

    def step(i):
        rp, k, v, y = batches[i % 4]
        model.forward_backward_sparse(rp, k, v, y, None, train=True)
        model.apply_gradient(a.batch)

    for i in range(3):
        step(i)
    torch.cuda.synchronize()
    mv.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(a.steps):
        step(i)
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device="cuda")
    if world > 1:
        import torch.distributed as dist
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    res = {"config": f"LogReg sparse dim={a.dim} nnz={a.nnz} batch={a.batch}/GPU adagrad server updater", "n_gpus": world,
           "ms_per_step": float(ms) / a.steps, "samples_per_sec": a.batch * world * a.steps / (float(ms) / 1e3),
           "loss": float(model.loss.item()) / max(1, a.batch * (a.steps + 3))}
    if rank == 0:
        os.makedirs("gpurun_out", exist_ok=True)
        with open(f"gpurun_out/logreg_sparse_n{world}.json", "w") as f:
            json.dump(res, f, indent=1)
        print(json.dumps(res), flush=True)
    mv.barrier()
    mv.shutdown()


if __name__ == "__main__":
    main()
