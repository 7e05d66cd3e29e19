"""ResNet-32 on CIFAR-10-shaped data with ASGD through multiverso -- the PyTorch counterpart of the
reference's headline binding benchmark (binding/python/examples/theano/lasagne/
Deep_Residual_Learning_CIFAR-10.py:59-86 model, :271-397 training loop; numbers in
binding/python/docs/BENCHMARK.md:58-62: batch 64 per worker, lr 0.05 for 8 workers, sync after every
batch, barrier per epoch, lr /10 at epochs 41 and 61).

There is no network in the build image, so `--synthetic` (default) draws CIFAR-shaped random
images with a learnable label rule; pass --data DIR with `data_batch_*` pickles for the real set.

    python resnet_cifar10.py --epochs 1                     (1 worker)
    torchrun --nproc-per-node 8 resnet_cifar10.py           (8 workers, one GPU each)
"""
import argparse
import os
import pickle
import sys
import time

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

import multiverso as mv
from multiverso.torch_ext import TorchParamManager


class Block(nn.Module):
    """Basic pre-projection residual block (He et al. 2015, the Lasagne example's residual_block)."""

    def __init__(self, cin, cout, stride):
        super().__init__()
        self.c1 = nn.Conv2d(cin, cout, 3, stride, 1, bias=False)
        self.b1 = nn.BatchNorm2d(cout)
        self.c2 = nn.Conv2d(cout, cout, 3, 1, 1, bias=False)
        self.b2 = nn.BatchNorm2d(cout)
        self.proj = None
        if stride != 1 or cin != cout:
            self.proj = nn.Sequential(nn.Conv2d(cin, cout, 1, stride, bias=False), nn.BatchNorm2d(cout))

    def forward(self, x):
        y = F.relu(self.b1(self.c1(x)))
        y = self.b2(self.c2(y))
        return F.relu(y + (x if self.proj is None else self.proj(x)))


def resnet(n=5, classes=10):
    """6n+2 layers: n=5 -> ResNet-32 (the benchmark's depth)."""
    layers = [nn.Conv2d(3, 16, 3, 1, 1, bias=False), nn.BatchNorm2d(16), nn.ReLU()]
    cin = 16
    for cout, stride in ((16, 1), (32, 2), (64, 2)):
        for i in range(n):
            layers.append(Block(cin, cout, stride if i == 0 else 1))
            cin = cout
    layers += [nn.AdaptiveAvgPool2d(1), nn.Flatten(), nn.Linear(64, classes)]
    return nn.Sequential(*layers)


def load_data(path, n_synth, seed):
    if path:
        xs, ys = [], []
        for i in range(1, 6):
            with open(os.path.join(path, f"data_batch_{i}"), "rb") as f:
                d = pickle.load(f, encoding="latin1")
            xs.append(d["data"])
            ys.append(d["labels"])
        x = np.concatenate(xs).reshape(-1, 3, 32, 32).astype(np.float32) / 255.0
        y = np.concatenate(ys).astype(np.int64)
        x -= x.mean(0, keepdims=True)
        return torch.from_numpy(x), torch.from_numpy(y)
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(n_synth, 3, 32, 32, generator=g)
    # 10 classes from coarse image statistics: learnable, not trivially separable
    feats = torch.stack([x[:, c, 8 * i:8 * i + 16, 8 * j:8 * j + 16].mean((1, 2))
                         for c in range(3) for i in range(2) for j in range(2)], 1)
    proj = torch.randn(feats.shape[1], 10, generator=g)
    return x, (feats @ proj).argmax(1)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--epochs", type=int, default=2)
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--lr", type=float, default=None, help="default: 0.1 for 1 worker, 0.05 otherwise (BENCHMARK.md)")
    ap.add_argument("--n", type=int, default=5, help="blocks per stage: depth 6n+2")
    ap.add_argument("--data", default="")
    ap.add_argument("--samples", type=int, default=8192, help="synthetic training-set size")
    ap.add_argument("--sync-freq", type=int, default=1)
    a = ap.parse_args()
    mv.init()
    wid, W = mv.worker_id(), mv.workers_num()
    dev = torch.device("cuda", torch.cuda.current_device()) if torch.cuda.is_available() else torch.device("cpu")
    torch.manual_seed(1234)
    model = resnet(a.n).to(dev)
    pm = TorchParamManager(model)                # ONE ArrayTable for all parameters, master-initialised
    lr = a.lr if a.lr is not None else (0.1 if W == 1 else 0.05)
    opt = torch.optim.SGD(model.parameters(), lr=lr, momentum=0.9, weight_decay=1e-4)
    x, y = load_data(a.data, a.samples, seed=7)
    nb = x.shape[0] // a.batch
    for epoch in range(a.epochs):
        if epoch in (41, 61):
            for gparam in opt.param_groups:
                gparam["lr"] *= 0.1
        perm = torch.randperm(x.shape[0], generator=torch.Generator().manual_seed(epoch))
        t0, seen, tot_loss, correct = time.time(), 0, 0.0, 0
        model.train()
        for b in range(nb):
            if b % W != wid:                     # every worker trains its own share of the batches
                continue
            idx = perm[b * a.batch:(b + 1) * a.batch]
            xb, yb = x[idx].to(dev, non_blocking=True), y[idx].to(dev, non_blocking=True)
            if torch.rand(1).item() < 0.5:
                xb = xb.flip(3)                  # the example's horizontal-flip augmentation
            opt.zero_grad(set_to_none=True)
            out = model(xb)
            loss = F.cross_entropy(out, yb)
            loss.backward()
            opt.step()
            if (b // W) % a.sync_freq == 0:
                pm.sync_all_param()              # push delta, pull the merged model
            seen += xb.shape[0]
            tot_loss += loss.item() * xb.shape[0]
            correct += int((out.argmax(1) == yb).sum())
        mv.barrier()
        if mv.is_master_worker():
            print(f"epoch {epoch}: {time.time() - t0:.2f} s  loss {tot_loss / max(seen, 1):.4f}  "
                  f"acc {100.0 * correct / max(seen, 1):.2f}%  ({W} workers)", flush=True)
    mv.shutdown()


if __name__ == "__main__":
    main()
