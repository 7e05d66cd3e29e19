"""A linear layer whose weight lives in the parameter server.

``y = x @ W^T`` with W = a row-sharded MatrixTable (``[out_features, in_features]``): the forward is
the fused Get+GEMM (the pulled weight is never materialised on the GPU path), the backward
returns ``dL/dx = dL/dy @ W`` to autograd and pushes the weight gradient ``dL/dW = dL/dy^T @ x``
straight into the table -- where the table's server-side updater (sgd, momentum, AdaGrad, DC-ASGD ...)
applies it.  This is the reference's usage pattern (pull parameters, compute, push deltas;
e.g. binding/python/multiverso/theano_ext/param_manager.py:9-82) expressed as a differentiable
module instead of a manual sync call, so very large output / embedding layers never need a local copy.
Works on both backends (on the host backend ``get_gemm`` is a Get followed by a matmul).
"""
from __future__ import annotations

from typing import Optional

import torch

from ..tables.options import AddOption
from .get_gemm import get_gemm


def _as_like(value, like: torch.Tensor) -> torch.Tensor:
    t = torch.as_tensor(value)
    return t.to(device=like.device, dtype=like.dtype)


class _PSLinearFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x: torch.Tensor, table, push_scale, option):
        ctx.table, ctx.push_scale, ctx.option = table, push_scale, option
        ctx.save_for_backward(x)
        return get_gemm(table, x.contiguous())

    @staticmethod
    def backward(ctx, dy: torch.Tensor):
        (x,) = ctx.saved_tensors
        table = ctx.table
        dy = dy.contiguous()
        dx = None
        if ctx.needs_input_grad[0]:
            w = _as_like(table.get(), dy).view(table.num_row, table.num_col)
            dx = dy @ w
        if ctx.push_scale is not None:
            dw = (dy.t() @ x) * ctx.push_scale                  # [out_features, in_features]
            table.add_async(dw.reshape(-1), ctx.option)
        return dx, None, None, None


def ps_linear(table, x: torch.Tensor, push_scale: Optional[float] = 1.0, option: Optional[AddOption] = None):
    """Differentiable ``x @ W^T`` against a MatrixTable.  ``push_scale``: the weight gradient times this
    factor is added to the table during backward (e.g. the learning rate with the ``sgd`` updater, which
    subtracts; ``-lr`` with the ``default`` updater, which adds; ``None``: do not push)."""
    return _PSLinearFn.apply(x, table, push_scale, option)


class PSLinear(torch.nn.Module):
    """``torch.nn.Module`` wrapper: ``PSLinear(table, lr)`` is a bias-free Linear(in, out) whose weight is
    the table; its optimiser step is the table's updater, executed on the owners."""

    def __init__(self, table, push_scale: Optional[float] = 1.0, option: Optional[AddOption] = None):
        super().__init__()
        self.table, self.push_scale, self.option = table, push_scale, option
        self.in_features, self.out_features = table.num_col, table.num_row

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        lead = x.shape[:-1]
        y = ps_linear(self.table, x.reshape(-1, self.in_features), self.push_scale, self.option)
        return y.view(*lead, self.out_features)

    def extra_repr(self) -> str:
        return f"in_features={self.in_features}, out_features={self.out_features}, weight=MatrixTable#{getattr(self.table, 'table_id', '?')}"
