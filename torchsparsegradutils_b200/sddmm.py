"""Sampled dense-dense product on a sparse pattern -- the building block of every sparse gradient.

``sddmm(A, X, Y)[e] = <X[i_e, :], Y[j_e, :]>`` for each stored entry ``(i_e, j_e)`` of ``A``, in A's
storage order.  This is the fused form of the ``index_select x2 -> mul -> sum`` idiom the reference
repeats in ``sparse_matmul.py:201-205`` (grad_A of sparse_mm), ``sparse_solve.py:216-235`` / ``:487-504``
(triangular and generic solves, with a sign flip) and ``sparse_lstsq.py:239-256``; SURVEY.md section 8(f)
rank 1.  No ``nnz x K`` temporary is materialised.  Not differentiable itself (it *is* the backward).
"""
from __future__ import annotations

import torch

from . import _ops
from ._pattern import csr_pattern


def sddmm(A: torch.Tensor, X: torch.Tensor, Y: torch.Tensor) -> torch.Tensor:
    """Values of ``X @ Y^T`` sampled at A's stored entries.

    A : sparse COO (2-D) or CSR (2-D or batched) CUDA tensor, shape ``(n, m)`` / ``(b, n, m)``; only its
        pattern is read.
    X : dense ``(n, K)`` / ``(b, n, K)``;  Y : dense ``(m, K)`` / ``(b, m, K)`` (any strides).
    Returns a dense tensor shaped like ``A.values()`` (CSR) or ``A._values()`` (COO, storage order,
    duplicates each get the full dot product).
    """
    if A.layout not in (torch.sparse_coo, torch.sparse_csr):
        raise ValueError("A should be in either COO or CSR sparse format")
    if not (A.is_cuda and X.is_cuda and Y.is_cuda):
        raise RuntimeError("torchsparsegradutils_b200.sddmm runs on CUDA tensors only; there is no CPU fallback")
    if X.dim() != A.dim() or Y.dim() != A.dim():
        raise ValueError("A, X and Y must all be 2D or all be 3D tensors")
    if X.size(-2) != A.size(-2) or Y.size(-2) != A.size(-1) or X.size(-1) != Y.size(-1):
        raise ValueError(f"Incompatible shapes: A {tuple(A.shape)}, X {tuple(X.shape)}, Y {tuple(Y.shape)}")
    if X.dtype != Y.dtype:
        raise RuntimeError(f"sddmm: X and Y must have the same dtype, got {X.dtype} and {Y.dtype}")
    X, Y = X.detach(), Y.detach()
    if A.layout == torch.sparse_coo:
        if A.dim() != 2 or A.sparse_dim() != 2:
            raise ValueError("COO input to sddmm must be 2-D (use CSR for batched patterns)")
        row, col = A._indices()
        return _ops.sddmm_coo(row, col, X, Y)
    pat = csr_pattern(A)
    out = _ops.sddmm(pat, X, Y, None, pat.nnz_total)
    return out.view(A.values().shape)


def solve_grad_A(A: torch.Tensor, gradB: torch.Tensor, x: torch.Tensor, transpose: bool = False) -> torch.Tensor:
    """Sparse gradient of a linear solve ``A x = B`` (``A^T x = B`` if ``transpose``) w.r.t. A's stored entries.

    ``gradA[i, j] = -<gradB[i, :], x[j, :]>``; with ``transpose``: ``-<gradB[j, :], x[i, :]>`` -- the step the
    reference evaluates with ``repeat_interleave`` + two ``index_select`` + ``mul`` + ``sum`` in
    ``SparseTriangularSolve.backward`` (``sparse_solve.py:216-235``), ``SparseGenericSolve.backward``
    (``:487-504``) and the first term of ``SparseGenericLstsq.backward`` (``sparse_lstsq.py:239-246``), here one
    fused SDDMM launch with no ``nnz x K`` temporaries.  ``gradB`` is the dense gradient the caller obtained from
    its own (transposed) solve and ``x`` the forward solution, both ``(n, K)`` or ``(n,)``.  (The least-squares
    backward adds a second sampled product, ``sddmm(A, B - A x, lstsq(A, gradB))``, ``sparse_lstsq.py:248-258``.)

    Returns a sparse tensor with A's layout, shape and index tensors (CSR: A's own crow/col; COO: A's indices in
    storage order), like the reference (``sparse_solve.py:237-240``, ``:510-513``).
    """
    if A.layout not in (torch.sparse_coo, torch.sparse_csr):
        raise ValueError("A should be in either COO or CSR sparse format")
    if A.dim() != 2:
        raise ValueError("solve_grad_A expects a 2-D sparse matrix (split batched operands per item)")
    if gradB.dim() == 1:
        gradB = gradB.unsqueeze(-1)
    if x.dim() == 1:
        x = x.unsqueeze(-1)
    vals = sddmm(A, x, gradB) if transpose else sddmm(A, gradB, x)
    vals = vals.neg_()
    if vals.dtype != A.dtype:
        vals = vals.to(A.dtype)
    if A.layout == torch.sparse_coo:
        return torch.sparse_coo_tensor(A._indices(), vals, A.shape)
    return torch.sparse_csr_tensor(A.crow_indices(), A.col_indices(), vals, A.shape)
