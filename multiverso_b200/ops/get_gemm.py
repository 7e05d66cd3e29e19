"""Fused Get + GEMM (K2-fused): ``y = x @ W^T`` where W is a row-sharded MatrixTable.

The pulled row block is never materialised: a persistent, warp-specialised kernel streams each
W tile from its owner (local or peer HBM over NVLink) with TMA into shared memory and multiplies
it with tcgen05.mma (CTA pairs, cta_group::2, TF32 operands, fp32 accumulation in TMEM) against
the local activations; remote tiles are staged once in an L2-resident scratch, so W crosses NVLink
once whatever M is.  Knobs (env): MVB_GEMM_CTAS=1 single-CTA MMAs, MVB_GEMM_XC=1|2 accumulators per
work item, MVB_GEMM_WCACHE=0 no scratch staging, MVB_GEMM_PROF=1 in-kernel wait-cycle counters.
Reference analogue: MatrixWorkerTable::Get followed by the application's first GEMM
(e.g. LogReg Objective::Predict, Applications/LogisticRegression/src/objective/objective.cpp:113-120).
"""
from __future__ import annotations

import ctypes as C

import torch

from .. import _native as N


def get_gemm_supported() -> bool:
    """True when the fused tcgen05 kernel can run here (sm_100a device present)."""
    if not torch.cuda.is_available():
        return False
    return bool(N.cuda_lib().mvb_get_gemm_supported())


def get_gemm(table, x: torch.Tensor, out: torch.Tensor = None) -> torch.Tensor:
    """``x``: [M, K] fp32 CUDA, ``table``: MatrixDeviceTable with num_col == K. Returns [M, num_row].
    On the host backend (no GPU) the same call is a plain Get followed by a matmul."""
    from ..runtime import Runtime
    if Runtime.get().backend == "host":
        w = torch.as_tensor(table.get()).view(table.num_row, table.num_col).to(torch.float32)
        y = torch.as_tensor(x, dtype=torch.float32) @ w.t()
        if out is not None:
            out.copy_(y)
            return out
        return y
    assert x.is_cuda and x.dtype == torch.float32 and x.dim() == 2 and x.is_contiguous()
    M, K = x.shape
    assert K == table.num_col and table.dtype == torch.float32
    if K % 4 != 0:
        raise ValueError("get_gemm needs K % 4 == 0 (16-byte row pitch for TMA)")
    Nrows = table.num_row
    if out is None:
        out = torch.empty(M, Nrows, dtype=torch.float32, device=x.device)
    g = N.GetGemm()
    g.x, g.y, g.w_cache = x.data_ptr(), out.data_ptr(), None
    g.M, g.N, g.K = M, Nrows, K
    g.wmap.num_row, g.wmap.num_col = Nrows, K
    g.wmap.nservers = table.S
    g.wmap.rows_per_server = table.rps
    for s in range(table.S):
        g.wmap.shard_ptrs[s] = table.shard_ptrs[s]
    g.local_server = table.sid if table.S > 1 else 0
    N.check(N.cuda_lib().mvb_get_gemm_fused(C.byref(g), C.c_void_p(N.stream_ptr())), "mvb_get_gemm_fused")
    return out
