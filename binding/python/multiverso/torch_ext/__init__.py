"""PyTorch extension: the counterpart of the reference's theano_ext / lasagne_ext / keras_ext
(binding/python/multiverso/theano_ext/sharedvar.py, theano_ext/param_manager.py,
keras_ext/callbacks.py).  Theano shared variables become torch tensors / nn.Parameters."""
from .sharedvar import MVSharedVariable, mv_shared, sync_all_mv_shared_vars
from .param_manager import MVModelParamManager, TorchParamManager, MVCallback

__all__ = ["MVSharedVariable", "mv_shared", "sync_all_mv_shared_vars", "MVModelParamManager",
           "TorchParamManager", "MVCallback"]
