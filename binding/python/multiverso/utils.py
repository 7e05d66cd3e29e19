"""Helpers of the reference binding (binding/python/multiverso/utils.py:15-79): the library
loader is replaced by multiverso_b200's in-tree build; ``convert_data`` keeps its contract
(anything -> contiguous float32 numpy)."""
import numpy as np


def convert_data(data):
    """Convert the data to a contiguous float32 numpy array."""
    try:
        import torch
        if torch.is_tensor(data):
            data = data.detach().cpu().numpy()
    except ImportError:
        pass
    return np.ascontiguousarray(np.asarray(data, dtype=np.float32))


class Loader(object):
    """Kept for API parity: returns the ctypes handle of the C++ host runtime."""
    LIB = None

    @classmethod
    def get_lib(cls):
        if cls.LIB is None:
            from multiverso_b200 import _native
            cls.LIB = _native.host_lib()
        return cls.LIB
