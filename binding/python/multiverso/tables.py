"""ArrayTableHandler / MatrixTableHandler (reference: binding/python/multiverso/tables.py:
14-165): float32 tables with the master-init protocol -- every worker issues a *sync* add
at construction, the master adds ``init_value`` and the others add zeros."""
import numpy as np

import multiverso_b200 as _mv

from . import api
from .utils import convert_data

__all__ = ["TableHandler", "ArrayTableHandler", "MatrixTableHandler"]


def _to_backend(arr):
    """numpy -> what the active backend wants (CUDA tensor on the device backend)."""
    if _mv.runtime.Runtime.get().backend == "device":
        import torch
        return torch.from_numpy(np.ascontiguousarray(arr)).cuda()
    return arr


def _to_numpy(x):
    try:
        import torch
        if torch.is_tensor(x):
            return x.detach().cpu().numpy()
    except ImportError:
        pass
    return np.asarray(x)


class TableHandler(object):
    """Interface of a table handler (init_value must be a float32-convertible array)."""

    def __init__(self, size, init_value=None):
        raise NotImplementedError("You must implement the __init__ method.")

    def get(self, size):
        raise NotImplementedError("You must implement the get method.")

    def add(self, data, sync=False):
        raise NotImplementedError("You must implement the add method.")


class ArrayTableHandler(TableHandler):
    """A dense 1-D float32 table shared by all workers."""

    def __init__(self, size, init_value=None):
        self._size = int(size)
        self._table = _mv.ArrayTable(self._size, "float32")
        if init_value is not None:
            init_value = convert_data(init_value).reshape(-1)
            # sync add is used because we want every worker to see the initial value
            self.add(init_value if api.is_master_worker() else np.zeros(self._size, np.float32), sync=True)
            api.barrier()

    def get(self):
        """Returns the whole table as a float32 numpy array."""
        return _to_numpy(self._table.get()).reshape(-1).astype(np.float32, copy=False)

    def add(self, data, sync=False):
        """Adds ``data`` to the table. With sync=False the call may return before the
        servers have applied the update."""
        data = convert_data(data).reshape(-1)
        assert data.size == self._size
        if sync:
            self._table.add(_to_backend(data))
        else:
            self._table.add_async(_to_backend(data))


class MatrixTableHandler(TableHandler):
    """A dense 2-D float32 table; rows can be fetched / updated individually."""

    def __init__(self, num_row, num_col, init_value=None):
        self._num_row, self._num_col = int(num_row), int(num_col)
        self._size = self._num_row * self._num_col
        self._table = _mv.MatrixTable(self._num_row, self._num_col, "float32")
        if init_value is not None:
            init_value = convert_data(init_value).reshape(-1)
            self.add(init_value if api.is_master_worker() else np.zeros(self._size, np.float32), sync=True)
            api.barrier()

    def get(self, row_ids=None):
        """Whole table (num_row x num_col) or the given rows (len(row_ids) x num_col)."""
        if row_ids is None:
            return _to_numpy(self._table.get()).reshape(self._num_row, self._num_col)
        ids = np.asarray(row_ids, dtype=np.int64)
        return _to_numpy(self._table.get_rows(_to_backend(ids) if _mv.runtime.Runtime.get().backend == "device" else ids)
                         ).reshape(len(ids), self._num_col)

    def add(self, data=None, row_ids=None, sync=False):
        """Adds ``data`` to the whole table or to the rows listed in ``row_ids``."""
        assert data is not None
        data = convert_data(data)
        if row_ids is None:
            assert data.size == self._size
            flat = _to_backend(data.reshape(-1))
            if sync:
                self._table.add(flat)
            else:
                self._table.add_async(flat)
        else:
            ids = np.asarray(row_ids, dtype=np.int64)
            assert data.size == len(ids) * self._num_col
            vals = _to_backend(data.reshape(len(ids), self._num_col))
            if _mv.runtime.Runtime.get().backend == "device":
                ids = _to_backend(ids)
            self._table.add_rows(ids, vals)
