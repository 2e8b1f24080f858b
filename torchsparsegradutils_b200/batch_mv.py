"""Batched sparse x dense mat-vec glue in front of ``sparse_mm``.

Mirror of ``_batch_sparse_mv`` (reference ``distributions/sparse_multivariate_normal.py:16-102``), the
immediate caller of ``sparse_mm`` inside ``SparseMultivariateNormal.rsample`` (``:362``, ``:365``).  The
reference hands the operator *views* -- ``bvec.t()`` and ``bvec.permute(1, 2, 0)`` (``:96``, ``:100``) -- and
ATen materialises a contiguous copy of them on every call.  Here the views go to the kernels as they are:
``_ops`` reads any element strides (packing once through ``tsgu_pack_dense`` when the 128-bit kernels want
it) and grad_B comes back in the operand's own layout, so the transposes below stay free views.
"""
from __future__ import annotations

from typing import Callable

import torch

from .sparse_matmul import sparse_mm


def batch_sparse_mv(bmat: torch.Tensor, bvec: torch.Tensor, op: Callable[..., torch.Tensor] = sparse_mm, **kwargs) -> torch.Tensor:
    """``bmat @ bvec`` for the four rank combinations the reference supports (no batch broadcasting):

    ============  ============  ==========
    ``bmat``      ``bvec``      result
    ============  ============  ==========
    ``(n, n)``    ``(n,)``      ``(n,)``
    ``(n, n)``    ``(k, n)``    ``(k, n)``   (k sample vectors as rows, as ``rsample`` draws them)
    ``(B, n, n)`` ``(B, n)``    ``(B, n)``
    ``(B, n, n)`` ``(k, B, n)`` ``(k, B, n)``
    ============  ============  ==========

    Any other pair raises ``ValueError("Invalid dimensions for bmat and bvec")`` (reference ``:102``).
    """
    if bmat.dim() == 2 and bvec.dim() == 1:
        return op(bmat, bvec.unsqueeze(-1), **kwargs).squeeze(-1)
    if bmat.dim() == 2 and bvec.dim() == 2:
        return op(bmat, bvec.t(), **kwargs).t()
    if bmat.dim() == 3 and bvec.dim() == 2:
        return op(bmat, bvec.unsqueeze(-1), **kwargs).squeeze(-1)
    if bmat.dim() == 3 and bvec.dim() == 3:
        return op(bmat, bvec.permute(1, 2, 0), **kwargs).permute(2, 0, 1)
    raise ValueError("Invalid dimensions for bmat and bvec")
