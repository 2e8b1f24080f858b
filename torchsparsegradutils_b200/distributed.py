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
