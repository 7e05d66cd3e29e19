"""The ``LogisticRegression`` application (reference: Applications/LogisticRegression,
``LogisticRegression <config>``; logreg.cpp:14-173).

    python -m multiverso_b200.apps.logreg mnist.config
    torchrun --nproc-per-node 8 -m multiverso_b200.apps.logreg ctr.config

Per epoch: reset the native async reader, stream minibatches (CSR for sparse input, dense rows
otherwise), K8 forward/backward, update (local or through the parameter server with pulls
every ``sync_frequency`` minibatches, optionally pipelined); log loss and timing every
``show_time_per_sample`` samples; ``Test()`` after every epoch writes predictions to
``output_file`` and logs the test error; finally ``SaveModel``.  With several ranks every
rank reads ``train_file`` and keeps the minibatches ``i % size == rank``; output files get the
``-<worker_id>`` suffix like the reference (ps_model.cpp:69-80).
"""
from __future__ import annotations

import ctypes as C
import sys
import time
from typing import Optional

import numpy as np

from ..utils import Log
from ._applib import lib


class SampleReader:
    """Native async reader (reader.cpp): libsvm / dense / weighted / bsparse -> CSR minibatches."""

    def __init__(self, files: str, reader_type: str, sparse: bool, input_size: int, buffer_samples: int):
        self.h = lib().MVA_LRReaderOpen(files.encode(), reader_type.encode(), int(sparse), int(input_size),
                                        int(buffer_samples))
        self.input_size = int(input_size)

    def reset(self):
        lib().MVA_LRReaderReset(self.h)

    def close(self):
        if self.h:
            lib().MVA_LRReaderClose(self.h)
            self.h = None

    def next(self, max_samples: int, max_nnz: Optional[int] = None):
        max_nnz = max_nnz or max_samples * min(self.input_size + 1, 4096)
        row_ptr = np.zeros(max_samples + 1, np.int64)
        labels = np.empty(max_samples, np.float32)
        weights = np.empty(max_samples, np.float32)
        while True:
            keys = np.empty(max_nnz, np.int64)
            vals = np.empty(max_nnz, np.float32)
            n = lib().MVA_LRReaderNext(self.h, max_samples, max_nnz, row_ptr.ctypes.data_as(C.c_void_p),
                                       keys.ctypes.data_as(C.c_void_p), vals.ctypes.data_as(C.c_void_p),
                                       labels.ctypes.data_as(C.c_void_p), weights.ctypes.data_as(C.c_void_p))
            if n >= 0:
                break
            max_nnz = max(2 * max_nnz, -n)      # one sample is wider than the buffer: grow and retry
        if n == 0:
            return None
        nnz = int(row_ptr[n])
        return row_ptr[:n + 1], keys[:nnz], vals[:nnz], labels[:n], weights[:n]


class LogReg:
    def __init__(self, cfg):
        import torch
        import multiverso_b200 as mv
        from ..models.logreg import LogRegModel
        self.cfg, self.mv, self.torch = cfg, mv, torch
        if cfg.use_ps:
            mv.init()
        elif not mv.runtime.Runtime.get().started:
            mv.init()
        self.rank, self.size = mv.rank(), mv.size()
        self.model = LogRegModel(cfg)
        if cfg.init_model_file:
            self.model.load(cfg.init_model_file)
        self.dev = self.model.dev

    def _to_dev(self, batch):
        torch = self.torch
        row_ptr, keys, vals, labels, weights = batch
        f = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(self.dev, non_blocking=True)
        return f(row_ptr), f(keys), f(vals), f(labels), f(weights)

    def _dense(self, batch):
        """CSR minibatch -> dense [n x dim] rows (dense-input path feeds the dense kernel)."""
        torch = self.torch
        row_ptr, keys, vals, labels, _ = batch
        n = len(labels)
        x = np.zeros((n, self.model.dim), np.float32)
        rows = np.repeat(np.arange(n), np.diff(row_ptr))
        x[rows, keys] = vals
        return torch.from_numpy(x).to(self.dev, non_blocking=True), torch.from_numpy(labels.copy()).to(self.dev)

    def _step(self, batch, train=True, pred=None):
        if self.cfg.sparse:
            rp, k, v, y, w = self._to_dev(batch)
            self.model.forward_backward_sparse(rp, k, v, y, w if self.cfg.reader_type == "weight" else None,
                                               train=train, pred=pred)
        else:
            x, y = self._dense(batch)
            self.model.forward_backward_dense(x, y, train=train, pred=pred)

    def train(self) -> dict:
        cfg, torch = self.cfg, self.torch
        reader = SampleReader(cfg.train_file, cfg.reader_type, cfg.sparse, cfg.input_size, cfg.read_buffer_size * 3)
        stats = {}
        t_start = time.time()
        for epoch in range(cfg.train_epoch):
            if epoch:
                reader.reset()
            self.model.loss.zero_()
            self.model.correct.zero_()
            seen, mb, shown, t0 = 0, 0, 0, time.time()
            while True:
                batch = reader.next(cfg.minibatch_size)
                if batch is None:
                    break
                mine = (mb % self.size) == self.rank
                mb += 1
                if not mine:
                    continue
                n = len(batch[3])
                self._step(batch, train=True)
                self.model.apply_gradient(n)
                seen += n
                if seen - shown >= cfg.show_time_per_sample:
                    shown = seen
                    loss = float(self.model.loss.item()) / max(seen, 1)
                    Log.info("Sample seen %d  train loss %.6f  (%.0f samples/s)", seen, loss,
                             seen / max(time.time() - t0, 1e-9))
            torch.cuda.synchronize()
            self.mv.barrier()
            el = time.time() - t0
            stats = {"epoch": epoch, "samples": seen, "seconds": el, "samples_per_sec": seen / max(el, 1e-9),
                     "train_loss": float(self.model.loss.item()) / max(seen, 1),
                     "train_acc": int(self.model.correct.item()) / max(seen, 1)}
            Log.info("epoch %d: %d samples in %.2fs (%.0f samples/s), loss %.5f acc %.4f", epoch, seen, el,
                     stats["samples_per_sec"], stats["train_loss"], stats["train_acc"])
            if cfg.test_file:
                stats["test_error"] = self.test()
        reader.close()
        stats["total_seconds"] = time.time() - t_start
        return stats

    def test(self) -> float:
        cfg, torch = self.cfg, self.torch
        if self.model.table is not None:
            self.mv.barrier()
            self.model.pull(blocking=True)
        reader = SampleReader(cfg.test_file, cfg.reader_type, cfg.sparse, cfg.input_size, cfg.read_buffer_size * 3)
        self.model.correct.zero_()
        saved_loss = self.model.loss.clone()
        total = 0
        suffix = f"-{self.mv.worker_id()}" if self.size > 1 else ""
        out = open(cfg.output_file + suffix, "w") if cfg.output_file else None
        while True:
            batch = reader.next(max(cfg.minibatch_size, 256))
            if batch is None:
                break
            n = len(batch[3])
            pred = torch.empty(n * self.model.out, device=self.dev)
            self._step(batch, train=False, pred=pred)
            total += n
            if out:
                p = pred.view(n, self.model.out).cpu().numpy()
                for row in p:
                    out.write(" ".join(f"{v:.6f}" for v in row) + "\n")
        reader.close()
        if out:
            out.close()
        self.model.loss.copy_(saved_loss)
        err = 1.0 - int(self.model.correct.item()) / max(total, 1)
        Log.info("test error: %.6f (%d samples)", err, total)
        return err

    def save_model(self):
        if self.cfg.output_model_file:
            self.model.save(self.cfg.output_model_file)


def run(config_file: str) -> dict:
    import torch
    if not torch.cuda.is_available():
        # no GPU: the native CPU implementation of the same application (csrc/host/apps/logreg)
        from ._applib import run_native
        return run_native("logreg", [config_file])
    from ..models.logreg import LogRegConfig
    import multiverso_b200 as mv
    cfg = LogRegConfig.from_file(config_file)
    app = LogReg(cfg)
    stats = app.train()
    app.save_model()
    mv.shutdown()
    return stats


if __name__ == "__main__":
    if len(sys.argv) < 2:
        print("usage: python -m multiverso_b200.apps.logreg <config file>")
        sys.exit(2)
    run(sys.argv[1])
