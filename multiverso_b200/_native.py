"""ctypes bindings of the two native libraries.

* :func:`cuda_lib` -> ``libmvb200.so`` (sm_100a kernels, ``csrc/cuda/mvb200.h``)
* :func:`host_lib` -> ``libmultiverso.so`` (C++ host runtime + C API, ``include/multiverso/c_api.h``)

The CUDA library is mandatory whenever a GPU is present: ops raise instead of silently
falling back to eager PyTorch, so a run that passes has really executed our kernels.
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

MAX_RANKS = 8
PAD_CHANNELS = 64
PAD_WORDS = PAD_CHANNELS * MAX_RANKS
EPOCH_FIN = 1 << 62

F32, F64, I32, I64, I8 = 0, 1, 2, 3, 4
UPD_DEFAULT, UPD_SGD, UPD_MOMENTUM, UPD_ADAGRAD, UPD_DCASGD, UPD_DCASGDA = range(6)
UPDATER_NAMES = {
    "default": UPD_DEFAULT, "sgd": UPD_SGD, "momentum_sgd": UPD_MOMENTUM,
    "adagrad": UPD_ADAGRAD, "dcasgd": UPD_DCASGD, "dcasgda": UPD_DCASGDA,
}

_LIBDIR = Path(__file__).resolve().parent / "_lib"
_cuda = None
_host = None

vp = C.c_void_p
i64 = C.c_int64
VP8 = vp * MAX_RANKS
I64x8 = i64 * MAX_RANKS
I32x8 = C.c_int * MAX_RANKS


class AddOpt(C.Structure):
    """Reference AddOption layout: {worker_id, momentum, lr, rho, lambda} (20 bytes)."""
    _fields_ = [("worker_id", C.c_int), ("momentum", C.c_float), ("lr", C.c_float),
                ("rho", C.c_float), ("lam", C.c_float)]


AddOptx8 = AddOpt * MAX_RANKS


class DenseAdd(C.Structure):
    _fields_ = [
        ("dtype", C.c_int), ("updater", C.c_int), ("shard", vp), ("state0", vp), ("state1", vp),
        ("shard_len", i64), ("shard_off", i64), ("state_stride", i64), ("nworkers", C.c_int),
        ("worker_mask", C.c_uint32), ("delta_ptrs", VP8), ("delta_multicast", vp), ("opts", AddOptx8), ("replica_ptrs", VP8), ("replica_multicast", vp), ("scale", C.c_float),
        ("clip", C.c_float), ("pads", C.POINTER(vp)), ("me", C.c_int), ("world", C.c_int),
        ("ch_ready", C.c_int), ("ch_done", C.c_int), ("epoch", C.c_uint64), ("worker_rank", I32x8),
        ("is_worker", C.c_int), ("err_flag", vp), ("fin_flag", vp), ("done_counter", vp),
        ("timeout_s", C.c_double), ("opt_box", VP8), ("my_worker", C.c_int),
    ]


class DenseGet(C.Structure):
    _fields_ = [
        ("dtype", C.c_int), ("out", vp), ("nservers", C.c_int), ("shard_ptrs", VP8),
        ("shard_offs", I64x8), ("shard_lens", I64x8), ("pads", C.POINTER(vp)), ("me", C.c_int),
        ("world", C.c_int), ("ch_done", C.c_int), ("epoch", C.c_uint64), ("server_rank", I32x8),
        ("err_flag", vp), ("timeout_s", C.c_double),
    ]


class RowMap(C.Structure):
    _fields_ = [("num_row", i64), ("num_col", i64), ("nservers", C.c_int),
                ("rows_per_server", i64), ("shard_ptrs", VP8)]


class KV(C.Structure):
    _fields_ = [("vtype", C.c_int), ("nservers", C.c_int), ("capacity", i64), ("keys", VP8),
                ("vals", VP8)]


class Allreduce(C.Structure):
    _fields_ = [("dtype", C.c_int), ("n", i64), ("bufs", VP8), ("out", vp),
                ("pads", C.POINTER(vp)), ("me", C.c_int), ("world", C.c_int), ("ch", C.c_int),
                ("epoch", C.c_uint64), ("err_flag", vp), ("done_counter", vp),
                ("timeout_s", C.c_double)]


class Sgns(C.Structure):
    _fields_ = [
        ("tokens", vp), ("n_tokens", i64), ("w_in", vp), ("w_out", vp), ("g2_in", vp),
        ("g2_out", vp), ("dim", C.c_int), ("ld", i64), ("window", C.c_int), ("negative", C.c_int),
        ("cbow", C.c_int), ("hs", C.c_int), ("use_adagrad", C.c_int), ("lr", C.c_float),
        ("alias_prob", vp), ("alias_idx", vp), ("vocab", C.c_int), ("neg_pool", vp),
        ("neg_pool_size", C.c_int), ("hs_points", vp), ("hs_codes", vp), ("hs_len", vp),
        ("hs_max_code", C.c_int), ("map_in", vp), ("map_out", vp), ("seed", C.c_uint64),
        ("loss_sum", vp), ("pair_count", vp), ("variant", C.c_int), ("max_ctas", C.c_int),
        ("scale_in", vp), ("scale_out", vp), ("neg_pool_size_ptr", vp),
        ("nservers", C.c_int), ("rows_per_server", i64), ("w_in_peers", VP8), ("w_out_peers", VP8),
    ]


class KeyMap(C.Structure):
    _fields_ = [("size", i64), ("per_server", i64), ("nservers", C.c_int), ("width", C.c_int),
                ("shard_ptrs", VP8), ("touched_ptrs", VP8)]


class RowBox(C.Structure):
    _fields_ = [
        ("map", RowMap), ("me", C.c_int), ("cap", i64), ("slot_bytes", i64), ("box", VP8), ("ack", VP8),
        ("seg", vp), ("done", vp), ("applied", vp), ("go", vp), ("err_flag", vp), ("timeout_s", C.c_double),
    ]


class WePrep(C.Structure):
    _fields_ = [
        ("tokens", vp), ("n_tokens", i64), ("vocab", C.c_int), ("negative", C.c_int),
        ("alias_prob", vp), ("alias_idx", vp), ("seed", C.c_uint64), ("bm_in", vp), ("bm_out", vp),
        ("chunk_sums", vp), ("map_in", vp), ("map_out", vp), ("ids_in", vp), ("ids_out", vp),
        ("neg_pool", vp), ("pool_cap", i64), ("counts", vp), ("cap_in", i64), ("cap_out", i64),
    ]


class LrSparse(C.Structure):
    _fields_ = [
        ("row_ptr", vp), ("keys", vp), ("vals", vp), ("labels", vp), ("sample_w", vp), ("n", i64),
        ("objective", C.c_int), ("w", vp), ("dim", i64), ("out", C.c_int), ("grad", vp),
        ("loss_sum", vp), ("correct", vp), ("pred", vp), ("err", vp), ("compute_grad", C.c_int),
    ]


class LrDense(C.Structure):
    _fields_ = [
        ("x", vp), ("labels", vp), ("n", i64), ("dim", i64), ("out", C.c_int),
        ("objective", C.c_int), ("w", vp), ("grad", vp), ("loss_sum", vp), ("correct", vp),
        ("pred", vp), ("err", vp), ("compute_grad", C.c_int),
    ]


class GetGemm(C.Structure):
    _fields_ = [("x", vp), ("y", vp), ("w_cache", vp), ("M", i64), ("N", i64), ("K", i64),
                ("wmap", RowMap), ("local_server", C.c_int)]


class NativeError(RuntimeError):
    pass


def _ensure_built(name: str) -> Path:
    path = _LIBDIR / name
    if os.environ.get("MVB200_NO_BUILD") == "1" and path.exists():
        return path
    from . import _build
    try:
        if name == "libmvb200.so":
            _build.build_cuda()
        else:
            _build.build_host()
    except Exception:
        if not path.exists():
            raise
    return path


def cuda_lib():
    """Load (building if needed) the sm_100a kernel library."""
    global _cuda
    if _cuda is None:
        lib = C.CDLL(str(_ensure_built("libmvb200.so")), mode=C.RTLD_GLOBAL)
        lib.mvb_last_error.restype = C.c_char_p
        _cuda = lib
    return _cuda


def host_lib():
    """Load (building if needed) the C++ host runtime."""
    global _host
    if _host is None:
        _host = C.CDLL(str(_ensure_built("libmultiverso.so")), mode=C.RTLD_GLOBAL)
    return _host


def check(rc: int, what: str = ""):
    if rc != 0:
        msg = cuda_lib().mvb_last_error()
        raise NativeError(f"{what or 'mvb200 call'} failed rc={rc}: {msg.decode() if msg else ''}")


def ptr(t) -> int:
    """Device/host pointer of a torch tensor (None -> NULL)."""
    return 0 if t is None else t.data_ptr()


def stream_ptr() -> int:
    import torch
    return torch.cuda.current_stream().cuda_stream


def dtype_code(torch_dtype) -> int:
    import torch
    return {torch.float32: F32, torch.float64: F64, torch.int32: I32, torch.int64: I64,
            torch.int8: I8}[torch_dtype]
