"""Sequence-to-sequence addition ("535+61" -> "596") with an LSTM encoder/decoder, trained by several
workers through multiverso -- the PyTorch counterpart of the reference's Keras example
(binding/python/examples/theano/keras/addition_rnn_mv.py:169-194: MVCallback syncs the whole model
through one ArrayTable every `freq` batches; README.md:9-14 reports val_acc 0.99+ reached earlier
with 2 workers).

    python addition_rnn.py --iters 20
    torchrun --nproc-per-node 2 addition_rnn.py
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
import numpy as np
import torch
import torch.nn as nn

import multiverso as mv
from multiverso.torch_ext import MVCallback

CHARS = "0123456789+ "
C2I = {c: i for i, c in enumerate(CHARS)}


def make_data(n, digits, rng):
    maxlen, outlen = 2 * digits + 1, digits + 1
    n = min(n, int(0.8 * (10 ** digits) * (10 ** digits + 1) / 2))     # unordered pairs that exist
    seen, q, a = set(), [], []
    while len(q) < n:
        x, y = (int("".join(rng.choice(list("0123456789")) for _ in range(rng.integers(1, digits + 1)))) for _ in range(2))
        key = tuple(sorted((x, y)))
        if key in seen:
            continue
        seen.add(key)
        q.append(f"{x}+{y}".ljust(maxlen)[::-1])       # the Keras example reverses the query
        a.append(str(x + y).ljust(outlen))
    enc = lambda s: [C2I[c] for c in s]
    return torch.tensor([enc(s) for s in q]), torch.tensor([enc(s) for s in a])


class Seq2Seq(nn.Module):
    def __init__(self, hidden=128, outlen=4):
        super().__init__()
        self.emb = nn.Embedding(len(CHARS), len(CHARS))
        self.emb.weight.data.copy_(torch.eye(len(CHARS)))          # one-hot inputs like the example
        self.emb.weight.requires_grad_(False)
        self.enc = nn.LSTM(len(CHARS), hidden, batch_first=True)
        self.dec = nn.LSTM(hidden, hidden, batch_first=True)
        self.out = nn.Linear(hidden, len(CHARS))
        self.outlen = outlen

    def forward(self, q):
        _, (h, _) = self.enc(self.emb(q))
        rep = h[-1].unsqueeze(1).repeat(1, self.outlen, 1)          # RepeatVector(DIGITS + 1)
        y, _ = self.dec(rep)
        return self.out(y)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--digits", type=int, default=3)
    ap.add_argument("--samples", type=int, default=20000)
    ap.add_argument("--iters", type=int, default=30)
    ap.add_argument("--batch", type=int, default=128)
    ap.add_argument("--freq", type=int, default=1, help="sync every `freq` batches")
    a = ap.parse_args()
    mv.init()
    wid, W = mv.worker_id(), mv.workers_num()
    dev = torch.device("cuda", torch.cuda.current_device()) if torch.cuda.is_available() else torch.device("cpu")
    rng = np.random.default_rng(0)                                  # same data on every worker, sharded below
    q, ans = make_data(a.samples, a.digits, rng)
    n_val = len(q) // 10
    qv, av, qt, at = q[:n_val].to(dev), ans[:n_val].to(dev), q[n_val:], ans[n_val:]
    torch.manual_seed(0)
    model = Seq2Seq(outlen=a.digits + 1).to(dev)
    cb = MVCallback(model, freq=a.freq)
    opt = torch.optim.Adam([p for p in model.parameters() if p.requires_grad], lr=2e-3)
    nb = len(qt) // a.batch
    for it in range(a.iters):
        model.train()
        perm = torch.randperm(len(qt), generator=torch.Generator().manual_seed(it))
        for b in range(nb):
            if b % W != wid:
                continue
            idx = perm[b * a.batch:(b + 1) * a.batch]
            opt.zero_grad(set_to_none=True)
            logits = model(qt[idx].to(dev))
            loss = nn.functional.cross_entropy(logits.flatten(0, 1), at[idx].to(dev).flatten())
            loss.backward()
            opt.step()
            cb.on_batch_end(b // W)
        mv.barrier()
        model.eval()
        with torch.no_grad():
            acc = (model(qv).argmax(-1) == av).all(1).float().mean().item()
        if mv.is_master_worker():
            print(f"iteration {it}: loss {loss.item():.4f}  val_acc {acc:.4f}  ({W} workers)", flush=True)
    mv.shutdown()


if __name__ == "__main__":
    main()
