"""MVModelParamManager for torch modules (reference: theano_ext/param_manager.py:9-82 and its
Lasagne / Keras subclasses; keras_ext/callbacks.py:8-39).  All parameters of a model are
flattened into ONE ArrayTable; ``sync_all_param`` = add(delta) + get + scatter back."""
import numpy as np
import torch

import multiverso_b200 as _mv

from .. import api


class MVModelParamManager(object):
    """Base class: subclasses say how to read / write all parameters of ``model``."""

    def __init__(self, model):
        self.model = model
        self._device = _mv.runtime.Runtime.get().backend == "device"
        flat = self._flatten(self.get_all_param_values())
        self.shape_sizes = [(tuple(p.shape), p.numel()) for p in self.get_all_param_values()]
        self.tbh = _mv.ArrayTable(flat.numel(), "float32")
        init = flat if api.is_master_worker() else torch.zeros_like(flat)
        self.tbh.add(init.cuda() if self._device else init.cpu().numpy())
        api.barrier()
        self.all_param_list = self._get_table()
        self._set_all(self.all_param_list)

    def get_all_param_values(self):
        raise NotImplementedError()

    def set_all_param_values(self, params):
        raise NotImplementedError()

    @staticmethod
    def _flatten(params):
        return torch.cat([p.detach().reshape(-1).to(torch.float32) for p in params])

    def _get_table(self):
        v = self.tbh.get()
        return v.clone() if torch.is_tensor(v) else torch.from_numpy(np.array(v, copy=True))

    def _set_all(self, flat):
        out, n = [], 0
        for shape, size in self.shape_sizes:
            out.append(flat[n:n + size].view(shape))
            n += size
        self.set_all_param_values(out)

    def sync_all_param(self):
        """Push the local delta, pull the merged parameters, write them into the model."""
        cur = self._flatten(self.get_all_param_values())
        last = self.all_param_list.to(cur.device)
        delta = cur - last
        self.tbh.add(delta.cuda() if self._device else delta.cpu().numpy())
        self.all_param_list = self._get_table()
        self._set_all(self.all_param_list)


class TorchParamManager(MVModelParamManager):
    """``model`` is a torch.nn.Module (the Lasagne/Keras managers' role in the reference)."""

    def get_all_param_values(self):
        return [p.data for p in self.model.parameters()]

    def set_all_param_values(self, params):
        for p, v in zip(self.model.parameters(), params):
            p.data.copy_(v.to(p.device).view_as(p.data))


class MVCallback(object):
    """Training-loop callback: sync every ``freq`` batches (keras_ext/callbacks.py:8-39).
    Call ``on_batch_end(batch_idx)`` from your loop."""

    def __init__(self, model, freq=1):
        self.kpm = TorchParamManager(model)
        self.freq = int(freq)

    def on_batch_end(self, batch, logs=None):
        if (batch + 1) % self.freq == 0:
            self.kpm.sync_all_param()
