"""mv_shared / MVSharedVariable / sync_all_mv_shared_vars for torch tensors.

Reference: theano_ext/sharedvar.py:12-99 -- a shared variable owns an ArrayTable sized to
its flattened value, master-initialised; ``mv_sync()`` pushes (current - last synced) and pulls
the merged value back.  On the device backend the tensor, the delta and the pulled value never
leave HBM (the reference round-trips GPU -> numpy -> 4 host copies -> MPI)."""
import numpy as np
import torch

import multiverso_b200 as _mv

from .. import api


class MVSharedVariable(object):
    """Wraps a torch tensor (or nn.Parameter) kept in sync through a multiverso ArrayTable."""

    shared_vars = []

    def __init__(self, tensor):
        self._tensor = tensor
        data = tensor.data if isinstance(tensor, torch.nn.Parameter) else tensor
        self._device = _mv.runtime.Runtime.get().backend == "device"
        self._table = _mv.ArrayTable(data.numel(), "float32")
        init = data.detach().reshape(-1).to(torch.float32)
        zeros = torch.zeros_like(init)
        self._push(init if api.is_master_worker() else zeros, sync=True)
        api.barrier()
        self._last = self._pull()
        self._assign(self._last)
        MVSharedVariable.shared_vars.append(self)

    # -- backend plumbing
    def _push(self, flat, sync=False):
        x = flat.cuda() if self._device else flat.cpu().numpy()
        (self._table.add if sync else self._table.add_async)(x)

    def _pull(self):
        v = self._table.get()
        return v.clone() if torch.is_tensor(v) else torch.from_numpy(np.array(v, copy=True))

    def _assign(self, flat):
        data = self._tensor.data if isinstance(self._tensor, torch.nn.Parameter) else self._tensor
        data.copy_(flat.to(data.device).view_as(data))

    # -- public API
    def mv_sync(self):
        """Push the local change since the last sync, pull the merged value."""
        data = self._tensor.data if isinstance(self._tensor, torch.nn.Parameter) else self._tensor
        cur = data.detach().reshape(-1).to(torch.float32)
        last = self._last.to(cur.device)
        self._push(cur - last)
        self._last = self._pull()
        self._assign(self._last)

    def get_value(self):
        return self._tensor

    def __getattr__(self, name):
        return getattr(self.__dict__["_tensor"], name)

    def __getstate__(self):
        return {"tensor": self._tensor.detach().cpu()}


def mv_shared(value, requires_grad=False):
    """Drop-in for ``theano.shared``: returns an MVSharedVariable around a tensor / Parameter."""
    if not torch.is_tensor(value):
        value = torch.as_tensor(np.asarray(value, dtype=np.float32))
    if requires_grad and not isinstance(value, torch.nn.Parameter):
        value = torch.nn.Parameter(value)
    return MVSharedVariable(value)


def sync_all_mv_shared_vars():
    """Sync every shared variable created through mv_shared."""
    for sv in MVSharedVariable.shared_vars:
        sv.mv_sync()
