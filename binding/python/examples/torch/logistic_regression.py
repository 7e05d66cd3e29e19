"""Data-parallel logistic regression with multiverso shared variables (PyTorch counterpart of
the reference's binding/python/examples/theano/logistic_regression.py:86-131,339-477).

    python logistic_regression.py                      (1 process)
    torchrun --nproc-per-node 4 logistic_regression.py (4 GPUs / processes)

Pattern of every reference example: mv.init() -> pick the device by worker id -> shard the
minibatches (idx % workers_num == worker_id) -> mv_sync after each batch -> mv.barrier() per
epoch -> the master validates -> mv.shutdown().  Synthetic MNIST-shaped data (no network)."""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
import torch

import multiverso as mv
from multiverso.torch_ext import mv_shared, sync_all_mv_shared_vars


def main(epochs=5, batch=500, lr=0.13):
    mv.init()
    wid, W = mv.worker_id(), mv.workers_num()
    dev = torch.device("cuda", torch.cuda.current_device()) if torch.cuda.is_available() else torch.device("cpu")
    g = torch.Generator().manual_seed(0)
    centers = torch.randn(10, 784, generator=g)
    y = torch.randint(0, 10, (12000,), generator=g)
    x = centers[y] * 0.5 + torch.randn(12000, 784, generator=g)
    xtr, ytr, xva, yva = x[:10000].to(dev), y[:10000].to(dev), x[10000:].to(dev), y[10000:].to(dev)
    # shared variables: every worker sees the master's initial value
    Wt = mv_shared(torch.zeros(784, 10, device=dev))
    b = mv_shared(torch.zeros(10, device=dev))
    n_batches = xtr.shape[0] // batch
    for epoch in range(epochs):
        for idx in range(n_batches):
            if idx % W != wid:
                continue                                   # this minibatch belongs to another worker
            xb, yb = xtr[idx * batch:(idx + 1) * batch], ytr[idx * batch:(idx + 1) * batch]
            w, bb = Wt.get_value().requires_grad_(True), b.get_value().requires_grad_(True)
            loss = torch.nn.functional.cross_entropy(xb @ w + bb, yb)
            gw, gb = torch.autograd.grad(loss, [w, bb])
            with torch.no_grad():
                w -= lr * gw
                bb -= lr * gb
            w.requires_grad_(False); bb.requires_grad_(False)
            sync_all_mv_shared_vars()                      # push delta, pull merged parameters
        mv.barrier()
        acc = None
        if mv.is_master_worker():
            with torch.no_grad():
                acc = ((xva @ Wt.get_value() + b.get_value()).argmax(1) == yva).float().mean().item()
            print(f"epoch {epoch}: validation accuracy {acc:.4f}")
    master = mv.is_master_worker()
    mv.shutdown()
    return acc if master else None


if __name__ == "__main__":
    main()
