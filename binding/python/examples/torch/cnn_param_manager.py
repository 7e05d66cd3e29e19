"""A small CNN trained with ASGD through the model parameter manager (PyTorch counterpart of
the reference's Lasagne ResNet / Keras examples: binding/python/examples/theano/lasagne/
Deep_Residual_Learning_CIFAR-10.py:59-86,271-397 and keras/addition_rnn_mv.py:169-194).
All parameters live in ONE ArrayTable; MVCallback syncs every `freq` batches."""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
import torch
import torch.nn as nn

import multiverso as mv
from multiverso.torch_ext import MVCallback


def main(epochs=2, batch=64, freq=1):
    mv.init()
    wid, W = mv.worker_id(), mv.workers_num()
    dev = torch.device("cuda", torch.cuda.current_device()) if torch.cuda.is_available() else torch.device("cpu")
    torch.manual_seed(0)
    model = nn.Sequential(nn.Conv2d(3, 16, 3, padding=1), nn.ReLU(), nn.MaxPool2d(2),
                          nn.Conv2d(16, 32, 3, padding=1), nn.ReLU(), nn.AdaptiveAvgPool2d(1), nn.Flatten(),
                          nn.Linear(32, 10)).to(dev)
    cb = MVCallback(model, freq=freq)            # master-initialises the table, pulls into every worker
    opt = torch.optim.SGD(model.parameters(), lr=0.05)
    g = torch.Generator().manual_seed(1)
    x = torch.randn(2048, 3, 32, 32, generator=g)
    y = (x.mean((1, 2, 3)) > 0).long() + 2 * (x[:, 0].mean((1, 2)) > 0).long()   # 4 synthetic classes
    x, y = x.to(dev), y.to(dev)
    for epoch in range(epochs):
        for idx in range(x.shape[0] // batch):
            if idx % W != wid:
                continue
            xb, yb = x[idx * batch:(idx + 1) * batch], y[idx * batch:(idx + 1) * batch]
            opt.zero_grad()
            loss = nn.functional.cross_entropy(model(xb), yb)
            loss.backward()
            opt.step()
            cb.on_batch_end(idx)
        mv.barrier()
        if mv.is_master_worker():
            print(f"epoch {epoch}: loss {loss.item():.4f}")
    mv.shutdown()


if __name__ == "__main__":
    main()
