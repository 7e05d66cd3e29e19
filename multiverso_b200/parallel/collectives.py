"""MV_Aggregate and friends on the device backend (K6).

Reference: MV_Aggregate<T> -> net::Allreduce -> MPI_Allreduce(MPI_IN_PLACE, SUM)
(src/multiverso.cpp:53-56, include/multiverso/net/mpi_net.h:147-151), available in
model-averaging mode (``-ma=true``) where no parameter server is started.  Here it is a
hand-written P2P kernel over symmetric staging: one-shot for latency-bound sizes,
two-shot (reduce own slice, write back to all) above ``TWO_SHOT_BYTES``.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, Tuple

import torch

from .. import _native as N
from ..runtime import Runtime
from ..utils import FLAGS, monitor

TWO_SHOT_BYTES = 512 * 1024


class _AggregateState:
    def __init__(self):
        self.staging = None
        self.cap = 0
        self.epoch = 0
        self.counter = None


_state = _AggregateState()


def reset() -> None:
    global _state
    _state = _AggregateState()


def aggregate(data: torch.Tensor, algo: str = "auto") -> torch.Tensor:
    """In-place SUM all-reduce of a CUDA tensor across all ranks; returns ``data``."""
    rt = Runtime.get()
    assert data.is_cuda and data.is_contiguous()
    if rt.size == 1:
        return data
    lib = N.cuda_lib()
    nbytes = data.numel() * data.element_size()
    st = _state
    if st.cap < nbytes:
        if st.staging is not None:
            rt.barrier()
            rt.release_symm(st.staging)
        st.cap = max(nbytes, 1 << 20)
        st.staging = rt.alloc_multicast(st.cap) or rt.alloc_symm(st.cap)
        st.counter = rt.done_counter_ptr()
    ch = 1  # channels 1 (ready) and 2 (done) are reserved for aggregate
    stream = C.c_void_p(N.stream_ptr())
    if st.epoch > 0:
        # nobody may still be reading our staging buffer from the previous call
        N.check(lib.mvb_wait(rt.pads_array(), rt.rank, rt.size, ch + 1, C.c_uint64(st.epoch),
                             C.c_uint32((1 << rt.size) - 1), C.c_void_p(rt.err_flag.data_ptr()),
                             C.c_double(float(FLAGS.get("barrier_timeout_s"))), stream), "mvb_wait")
    stage = st.staging.tensor(data.dtype, data.numel())
    stage.copy_(data.view(-1))
    st.epoch += 1
    a = N.Allreduce()
    a.dtype, a.n = N.dtype_code(data.dtype), data.numel()
    for r in range(rt.size):
        a.bufs[r] = st.staging.ptrs[r]
    a.out = data.data_ptr()
    pads = rt.pads_array()
    a.pads = C.cast(pads, C.POINTER(C.c_void_p))
    a.me, a.world, a.ch, a.epoch = rt.rank, rt.size, ch, st.epoch
    a.err_flag = rt.err_flag.data_ptr()
    a.done_counter = st.counter
    a.timeout_s = float(FLAGS.get("barrier_timeout_s"))
    two = (algo == "twoshot") or (algo in ("auto", "nvls") and nbytes > TWO_SHOT_BYTES)
    mc = getattr(st.staging, "multicast_ptr", 0)
    use_nvls = bool(mc) and data.dtype == torch.float32 and (algo == "nvls" or (algo == "auto" and two))
    with monitor("MV_AGGREGATE", cuda=True, nbytes=nbytes):
        if use_nvls:
            N.check(lib.mvb_allreduce_nvls(C.byref(a), C.c_void_p(mc), stream), "mvb_allreduce_nvls")
        else:
            fn = lib.mvb_allreduce_twoshot if two else lib.mvb_allreduce_oneshot
            N.check(fn(C.byref(a), stream), "mvb_allreduce")
    return data
