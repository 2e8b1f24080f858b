"""Multi-GPU drivers for ``sparse_mm`` on one 8xB200 box (one process per GPU, ``torch.distributed``).

The reference has no distributed code; these drivers sit *outside* the
``sparse_mm(A, B)`` signature and only partition work:

* **Batched inputs shard by batch item, with no collective.**  The reference's batched semantics are
  block-diagonal (``sparse_matmul.py:151-153``, ``utils/utils.py:505-513``): item k of C, grad_A and
  grad_B depends only on item k of A, B and the upstream gradient.  Each rank simply calls
  ``sparse_mm`` on its slice (:func:`batch_shard_bounds`, :func:`shard_batched`).
* **One large matrix shards by nnz-balanced row blocks with the dense B replicated.**  Forward and the
  SDDMM are local; grad_B = sum_p A_p^T G_p needs the one exchange step of the path: an all-reduce of
  the (m, K) partials (NCCL over NVLink; gloo in the CPU tests).  :func:`sparse_mm_row_sharded`.

* **Alternative for one large matrix: shard the dense K columns** (:func:`sparse_mm_k_sharded`,
  :func:`k_shard_bounds`).  Row sharding cannot scale: every rank still reads all of the replicated B twice and
  writes a full m x K grad_B partial, then all-reduces it (m x K elements).  Sharding K instead -- rank p owns
  B[:, K_p] and G[:, K_p], the (small) sparse A is replicated -- makes forward and grad_B purely local
  (C[:, K_p] = A B[:, K_p], grad_B[:, K_p] = A^T G[:, K_p]) and leaves one exchange step: grad_A's sampled dot
  products are partial sums over K, so the nnz values are all-reduced (nnz elements instead of m x K: 8 MB vs
  512 MiB on BASELINE config 5).  Every rank does 1/P of the dense traffic.

``local_mm`` lets the tests drive the partition / collective logic on CPU ranks (gloo) with the CPU
oracle as the per-rank operator; the default is the CUDA ``sparse_mm``.
"""
from __future__ import annotations

from typing import Callable, List, Optional, Tuple

import torch
import torch.distributed as dist

from .sparse_matmul import SparseMatMul, _grad_A, _grad_B, sparse_mm


# ------------------------------------------------------------------------------ batch sharding
def batch_shard_bounds(batch: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous chunk [lo, hi) of the batch dimension owned by `rank` (sizes differ by at most 1)."""
    if not (0 <= rank < world):
        raise ValueError(f"rank {rank} outside world of size {world}")
    base, extra = divmod(batch, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_batched(A: torch.Tensor, world: int, rank: int, *dense: torch.Tensor):
    """Slice a batched CSR / COO tensor and its dense companions down to this rank's batch items."""
    if A.dim() != 3:
        raise ValueError("shard_batched expects a batched (3-D) sparse tensor")
    lo, hi = batch_shard_bounds(A.shape[0], world, rank)
    shape = (hi - lo,) + tuple(A.shape[1:])
    if A.layout == torch.sparse_csr:
        A_loc = torch.sparse_csr_tensor(A.crow_indices()[lo:hi].contiguous(), A.col_indices()[lo:hi].contiguous(),
                                        A.values()[lo:hi].contiguous(), shape)
    elif A.layout == torch.sparse_coo:
        idx, val = A._indices(), A._values()
        keep = (idx[0] >= lo) & (idx[0] < hi)
        sub = idx[:, keep].clone()
        sub[0] -= lo
        A_loc = torch.sparse_coo_tensor(sub, val[keep], shape)
    else:
        raise ValueError("A should be in either COO or CSR sparse format")
    return (A_loc,) + tuple(d[lo:hi].contiguous() for d in dense)


# -------------------------------------------------------------------------------- row sharding
def nnz_balanced_row_blocks(crow: torch.Tensor, world: int) -> List[int]:
    """Row boundaries r_0 = 0 <= r_1 <= ... <= r_world = n such that every block holds ~nnz/world
    stored entries (prefix-sum search on crow)."""
    n = crow.numel() - 1
    nnz = int(crow[-1])
    targets = torch.tensor([nnz * p // world for p in range(1, world)], dtype=crow.dtype, device=crow.device)
    cuts = torch.searchsorted(crow.contiguous(), targets, right=False).clamp_(0, n).tolist() if world > 1 else []
    bounds = [0] + [int(c) for c in cuts] + [n]
    for i in range(1, len(bounds)):  # monotone even with long empty stretches
        bounds[i] = max(bounds[i], bounds[i - 1])
    return bounds


def shard_rows_csr(A: torch.Tensor, lo: int, hi: int) -> torch.Tensor:
    """Rows [lo, hi) of a 2-D CSR tensor as a CSR tensor of shape (hi - lo, m) (crow rebased to 0)."""
    if A.layout != torch.sparse_csr or A.dim() != 2:
        raise ValueError("shard_rows_csr expects a 2-D CSR tensor")
    crow, col, val = A.crow_indices(), A.col_indices(), A.values()
    s, e = int(crow[lo]), int(crow[hi])
    return torch.sparse_csr_tensor((crow[lo:hi + 1] - crow[lo]).contiguous(), col[s:e].contiguous(), val[s:e].contiguous(),
                                   (hi - lo, A.shape[1]))


class _AllReduceGrad(torch.autograd.Function):
    """Identity in forward; sums the gradient over the process group in backward (replicated operand)."""

    @staticmethod
    def forward(ctx, x, group):
        ctx.group = group
        return x.view_as(x)

    @staticmethod
    def backward(ctx, grad):  # type: ignore[override]
        grad = grad.contiguous()
        dist.all_reduce(grad, op=dist.ReduceOp.SUM, group=ctx.group)
        return grad, None


class _RowShardedMatMul(torch.autograd.Function):
    """sparse_mm on this rank's row block with the grad_B all-reduce issued as soon as the local partial
    exists, so it travels over NVLink while the SDDMM (grad_A, purely local) is still running."""

    @staticmethod
    def forward(ctx, A_local, B, group):
        ctx.group = group
        return SparseMatMul.forward(ctx, A_local, B)

    @staticmethod
    def backward(ctx, grad):  # type: ignore[override]
        gradB = work = None
        if ctx.needs_input_grad[1]:
            gradB = _grad_B(ctx, grad).contiguous()
            work = dist.all_reduce(gradB, op=dist.ReduceOp.SUM, group=ctx.group, async_op=True)
        gradA = _grad_A(ctx, grad) if ctx.needs_input_grad[0] else None
        if work is not None:
            work.wait()
        return gradA, gradB, None


def sparse_mm_row_sharded(A_local: torch.Tensor, B: torch.Tensor, group: Optional[dist.ProcessGroup] = None,
                          local_mm: Callable[[torch.Tensor, torch.Tensor], torch.Tensor] = sparse_mm) -> torch.Tensor:
    """C_local = A_local @ B for this rank's row block; B is replicated on every rank.

    grad_A_local and C_local stay local; grad_B is all-reduced so every replica of B sees the full
    A^T G (the only collective on the path).
    """
    distributed = dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1
    if not distributed:
        return local_mm(A_local, B)
    if local_mm is sparse_mm:  # product path: collective overlapped with the SDDMM
        return _RowShardedMatMul.apply(A_local, B, group)
    return local_mm(A_local, _AllReduceGrad.apply(B, group))


# ----------------------------------------------------------------------------- K (dense-column) sharding
def k_shard_bounds(K: int, world: int, rank: int, align: int = 1) -> Tuple[int, int]:
    """Contiguous block [lo, hi) of the dense columns owned by `rank`; block edges are multiples of `align`
    (pass the 128-bit vector width, 4 for fp32 / 8 for bf16, so every shard stays on the vectorised kernels)."""
    if not (0 <= rank < world):
        raise ValueError(f"rank {rank} outside world of size {world}")
    units = -(-K // align)
    base, extra = divmod(units, world)
    lo = (rank * base + min(rank, extra)) * align
    hi = lo + (base + (1 if rank < extra else 0)) * align
    return min(lo, K), min(hi, K)


def _sparse_values(t: torch.Tensor) -> torch.Tensor:
    return t.values() if t.layout == torch.sparse_csr else t._values()


class _AllReduceSparseGrad(torch.autograd.Function):
    """Identity on a sparse operand in forward; sums the gradient's stored values over the group in backward
    (gloo / injected-operator path of :func:`sparse_mm_k_sharded`)."""

    @staticmethod
    def forward(ctx, A, group):
        ctx.group = group
        return A.detach().clone() if A.layout == torch.sparse_coo else torch.sparse_csr_tensor(
            A.crow_indices(), A.col_indices(), A.values().detach().clone(), A.shape)

    @staticmethod
    def backward(ctx, grad):  # type: ignore[override]
        if grad.layout == torch.strided:  # dense gradient from a generic operator: reduce it whole
            g = grad.contiguous()
            dist.all_reduce(g, op=dist.ReduceOp.SUM, group=ctx.group)
            return g, None
        if grad.layout == torch.sparse_coo and not grad.is_coalesced():
            grad = grad.coalesce()
        v = _sparse_values(grad).contiguous()
        dist.all_reduce(v, op=dist.ReduceOp.SUM, group=ctx.group)
        if grad.layout == torch.sparse_csr:
            return torch.sparse_csr_tensor(grad.crow_indices(), grad.col_indices(), v, grad.shape), None
        return torch.sparse_coo_tensor(grad._indices(), v, grad.shape, is_coalesced=True), None


class _KShardedMatMul(torch.autograd.Function):
    """sparse_mm on this rank's block of dense columns.  Backward: the SDDMM over the local K block gives partial
    grad_A values; their all-reduce is issued asynchronously and travels over NVLink while the (purely local)
    grad_B SpMM runs."""

    @staticmethod
    def forward(ctx, A, B_local, group):
        ctx.group = group
        return SparseMatMul.forward(ctx, A, B_local)

    @staticmethod
    def backward(ctx, grad):  # type: ignore[override]
        gradA = work = None
        if ctx.needs_input_grad[0]:
            gradA = _grad_A(ctx, grad)
            work = dist.all_reduce(_sparse_values(gradA), op=dist.ReduceOp.SUM, group=ctx.group, async_op=True)
        gradB = _grad_B(ctx, grad) if ctx.needs_input_grad[1] else None
        if work is not None:
            work.wait()
        return gradA, gradB, None


def sparse_mm_k_sharded(A: torch.Tensor, B_local: torch.Tensor, group: Optional[dist.ProcessGroup] = None,
                        local_mm: Callable[[torch.Tensor, torch.Tensor], torch.Tensor] = sparse_mm) -> torch.Tensor:
    """C[:, K_p] = A @ B[:, K_p] for this rank's block of dense columns; A (sparse) is replicated on every rank.

    C[:, K_p] and grad_B[:, K_p] stay local; grad_A's stored values are all-reduced so every replica of A sees the
    full sampled product (the only collective: nnz elements).  Results equal the single-GPU ones up to the order
    of the K-sum in grad_A (P partial sums are added by the collective)."""
    distributed = dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1
    if not distributed:
        return local_mm(A, B_local)
    if local_mm is sparse_mm:
        return _KShardedMatMul.apply(A, B_local, group)
    return local_mm(_AllReduceSparseGrad.apply(A, group), B_local)
