"""Index/layout helpers of the sparse_mm path, backed by the sm_100a index-builder kernels.

Mirrors the public names, argument meaning and error behaviour of the reference's
``torchsparsegradutils/utils/utils.py`` for the functions on the hot path (SURVEY.md section 8 rows
a6-a12); results are bit-identical to the reference's on the same inputs
(``tests/test_utils_parity.py``).  Validation happens on the host before any device work, so the
error paths behave the same on any device; the compute itself is CUDA-only (no CPU fallback).
"""
from __future__ import annotations

from typing import List, Tuple

import torch

from .. import _native as nat
from .. import _ops
from .._pattern import _sort_coo


def _require_cuda(t: torch.Tensor, fn: str) -> None:
    if not t.is_cuda:
        raise RuntimeError(f"{fn}: torchsparsegradutils_b200 runs on CUDA tensors only (got {t.device}); no CPU fallback")


# ------------------------------------------------------------------------------------ stack_csr
def stack_csr(tensors: List[torch.Tensor], dim: int = 0) -> torch.Tensor:
    """Stack 2-D CSR tensors of equal shape and nnz into a batched CSR tensor.

    Reference: ``utils/utils.py:6-88`` (checks at ``:65-78`` in the same order, stacks at ``:80-82``).
    """
    if not isinstance(tensors, (list, tuple)):
        raise TypeError("Expected a list of tensors, but got {}.".format(type(tensors)))
    if len(tensors) == 0:
        raise ValueError("Cannot stack empty list of tensors.")
    first_shape = tensors[0].shape
    if any(t.shape != first_shape for t in tensors):
        raise ValueError("All tensors must have the same shape.")
    if any(t.layout != torch.sparse_csr for t in tensors):
        raise ValueError("All tensors must be in CSR layout.")
    if any(t.ndim != 2 for t in tensors):
        raise ValueError("All tensors must be 2D.")
    parts = [(t.crow_indices(), t.col_indices(), t.values()) for t in tensors]
    crow, col, val = (torch.stack([p[i] for p in parts], dim=dim) for i in range(3))
    shape = list(first_shape)
    shape.insert(dim, len(tensors))
    return torch.sparse_csr_tensor(crow, col, val, tuple(shape))


# ------------------------------------------------------------------------------ COO sort / CSR
def _sort_coo_indices(indices: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """Lexicographically sort COO coordinates; returns ``(sorted_indices, permutation)``.

    Reference: ``utils/utils.py:91-149`` (``torch.unique(sorted, return_inverse)`` + ``argsort``).
    Here: one stable radix sort (``tsgu_coo_sort``).  Like the reference it expects duplicate-free
    coordinates; with duplicates ties keep their original order and nothing is dropped.
    """
    _require_cuda(indices, "_sort_coo_indices")
    ind = indices.to(torch.int64).contiguous()
    ndim, nnz = ind.shape
    if nnz == 0:
        return ind.clone(), torch.zeros(0, dtype=torch.int64, device=ind.device)
    # coordinate extents: one fused device reduction; exact per-dimension maxima keep the key narrow
    dims = (ind.max(dim=1).values + 1).tolist()
    perm, srt = _sort_coo(ind, dims, ndim, nat.I64, True)
    return srt, perm


def _compress_row_indices(row_indices: torch.Tensor, num_rows: int) -> torch.Tensor:
    """Row indices -> CSR ``crow_indices`` (histogram + inclusive scan).

    Reference: ``utils/utils.py:152-233``; same checks and messages (``:214-226``), output dtype equals
    the input dtype (``:228-231``).
    """
    if not isinstance(row_indices, torch.Tensor):
        raise TypeError("row_indices must be a torch.Tensor.")
    if row_indices.ndim != 1:
        raise ValueError(f"row_indices must be 1D, got shape {tuple(row_indices.shape)}.")
    if row_indices.dtype not in (torch.int32, torch.int64):
        raise TypeError("row_indices must have integer dtype (torch.int32 or torch.int64).")
    if not isinstance(num_rows, int) or num_rows <= 0:
        raise ValueError("num_rows must be a positive integer.")
    if row_indices.numel() > 0:
        lo, hi = torch.aminmax(row_indices)
        if int(lo) < 0:
            raise ValueError("row_indices contains negative entries.")
        if int(hi) >= num_rows:
            raise ValueError("row_indices contains entries >= num_rows.")
    _require_cuda(row_indices, "_compress_row_indices")
    rows = row_indices.contiguous()
    dev = rows.device
    crow = torch.empty(num_rows + 1, dtype=rows.dtype, device=dev)
    idt = nat.idx_enum(rows.dtype)
    L = nat.lib()
    with torch.cuda.device(dev):
        ws = nat.workspace(L.tsgu_compress_rows_workspace_bytes(num_rows, idt), dev)
        nat.check(L.tsgu_compress_rows(rows.data_ptr(), rows.numel(), num_rows, crow.data_ptr(), idt, ws.data_ptr(),
                                       ws.numel(), nat.stream_ptr(dev)), "tsgu_compress_rows")
    return crow


def convert_coo_to_csr_indices_values(coo_indices, num_rows, values=None):
    """COO coordinates (+values) -> CSR ``(crow, col, values-or-permutation)``.

    Reference: ``utils/utils.py:236-346``.  Unbatched ``(2, nnz)`` gives ``crow (n+1,)``; batched
    ``(3, nnz)`` (equal nnz per item, ``:339-344``) gives ``crow (b, n+1)``, ``col (b, nnz/b)``.
    With ``values=None`` the third output is the permutation (``:327``, ``:342``).
    """
    if coo_indices.shape[0] < 2:
        raise ValueError(
            f"Indices tensor must have at least 2 rows (row and column indices). Got {coo_indices.shape[0]} rows.")
    elif coo_indices.shape[0] > 3:
        raise ValueError(
            "Current implementation only supports single batch diomension, therefore indices tensor must have at "
            f"most 3 rows (batch, row and column indices). Got {coo_indices.shape[0]} rows.")
    if coo_indices[-2].max() >= num_rows:
        raise ValueError(
            f"Row indices must be less than num_rows ({num_rows}). Got max row index {coo_indices[-2].max()}")
    if values is not None and values.shape[0] != coo_indices.shape[1]:
        raise ValueError(
            f"Number of values ({values.shape[0]}) does not match number of indices ({coo_indices.shape[1]})")
    _require_cuda(coo_indices, "convert_coo_to_csr_indices_values")

    srt, perm = _sort_coo_indices(coo_indices)
    batched = srt.shape[0] == 3
    nnz = srt.shape[1]
    dev = srt.device
    if batched:
        present = torch.unique_consecutive(srt[0])  # items that own entries (reference :330-337)
        nb = present.shape[0]
        # items are renumbered densely, exactly like the reference's loop over torch.unique(batch)
        dense_batch = torch.searchsorted(present, srt[0])
        idx3 = torch.stack([dense_batch, srt[1], srt[2]])
    else:
        nb, idx3 = 1, srt
    rowptr = torch.empty(nb * num_rows + 1, dtype=torch.int64, device=dev)
    col = torch.empty(nnz, dtype=torch.int64, device=dev)
    with torch.cuda.device(dev):
        nat.check(nat.lib().tsgu_coo_to_csr(idx3.data_ptr(), idx3.shape[0], nnz, idx3.stride(0), nb, num_rows, None,
                                            rowptr.data_ptr(), col.data_ptr(), nat.I64, nat.stream_ptr(dev)),
                  "tsgu_coo_to_csr")
    out_vals = perm if values is None else _ops.gather_values(values.contiguous(), perm)
    if not batched:
        return rowptr, col, out_vals
    if nnz % nb:
        raise RuntimeError(f"shape '[{nb}, -1]' is invalid for input of size {nnz}")  # reference: reshape fails
    per = nnz // nb
    starts = torch.arange(nb, device=dev, dtype=torch.int64) * num_rows
    crow = rowptr[starts.unsqueeze(1) + torch.arange(num_rows + 1, device=dev)] - rowptr[starts].unsqueeze(1)
    return crow, col.reshape(nb, per), out_vals.reshape(nb, per)


def convert_coo_to_csr(sparse_coo_tensor: torch.Tensor) -> torch.Tensor:
    """COO tensor -> CSR tensor (coalescing first). Reference: ``utils/utils.py:349-410``."""
    if sparse_coo_tensor.layout == torch.sparse_coo:
        if sparse_coo_tensor.is_coalesced() is False:
            sparse_coo_tensor = sparse_coo_tensor.coalesce()
        crow, col, vals = convert_coo_to_csr_indices_values(
            sparse_coo_tensor.indices(), sparse_coo_tensor.size()[-2], sparse_coo_tensor.values())
        return torch.sparse_csr_tensor(crow, col, vals, sparse_coo_tensor.size())
    raise ValueError(f"Unsupported layout: {sparse_coo_tensor.layout}")


def _demcompress_crow_indices(crow_indices: torch.Tensor, num_rows: int) -> torch.Tensor:
    """``crow_indices`` -> per-entry row index, dtype preserved. Reference: ``utils/utils.py:413-470``."""
    _require_cuda(crow_indices, "_demcompress_crow_indices")
    crow = crow_indices.contiguous()
    nnz = int(crow[num_rows])  # output size is data dependent (the reference syncs in repeat_interleave)
    rows = torch.empty(nnz, dtype=crow.dtype, device=crow.device)
    with torch.cuda.device(crow.device):
        nat.check(nat.lib().tsgu_decompress_crow(crow.data_ptr(), num_rows, nnz, rows.data_ptr(),
                                                 nat.idx_enum(crow.dtype), nat.stream_ptr(crow.device)),
                  "tsgu_decompress_crow")
    return rows


# --------------------------------------------------------------------------- block-diagonal
def sparse_block_diag(*sparse_tensors: torch.Tensor) -> torch.Tensor:
    """Block-diagonal concatenation of 2-D sparse tensors (all COO or all CSR).

    Reference: ``utils/utils.py:474-645``.  ``sparse_mm`` itself no longer needs this (its kernels take
    batch strides); it is kept because it is part of the reference's public ``utils`` surface.
    """
    if len(sparse_tensors) == 0:
        raise ValueError("At least one sparse tensor must be provided.")
    if len(sparse_tensors) == 1 and isinstance(sparse_tensors[0], (list, tuple)):
        raise TypeError("Sparse tensors must be provided as separate arguments, not as a list or tuple.")
    if not all(isinstance(t, torch.Tensor) for t in sparse_tensors):
        raise TypeError("All inputs must be torch.Tensor objects.")
    if all(t.layout == torch.sparse_coo for t in sparse_tensors):
        layout = torch.sparse_coo
    elif all(t.layout == torch.sparse_csr for t in sparse_tensors):
        layout = torch.sparse_csr
    else:
        raise ValueError("Sparse tensors must either be all sparse_coo or all sparse_csr.")
    if not all(t.sparse_dim() == 2 for t in sparse_tensors):
        raise ValueError("All sparse tensors must have exactly two sparse dimensions.")
    if not all(t.dense_dim() == 0 for t in sparse_tensors):
        raise ValueError("All sparse tensors must have zero dense dimensions.")
    if len(sparse_tensors) == 1:
        return sparse_tensors[0]

    rows_total = sum(t.size(-2) for t in sparse_tensors)
    cols_total = sum(t.size(-1) for t in sparse_tensors)
    if layout == torch.sparse_coo:
        idx_parts, val_parts = [], []
        r_off = c_off = 0
        for t in sparse_tensors:
            t = t if t.is_coalesced() else t.coalesce()
            off = torch.tensor([[r_off], [c_off]], dtype=torch.int64, device=t.device)
            idx_parts.append(t.indices() + off)
            val_parts.append(t.values())
            r_off += t.size(-2)
            c_off += t.size(-1)
        return torch.sparse_coo_tensor(torch.cat(idx_parts, dim=1), torch.cat(val_parts), size=(rows_total, cols_total))

    crow_parts, col_parts, val_parts = [], [], []
    c_off = 0
    nnz_off = None  # device scalar: running nnz, never synchronised to the host
    for k, t in enumerate(sparse_tensors):
        crow = t.crow_indices()
        crow_parts.append(crow if k == 0 else crow[1:] + nnz_off)
        col_parts.append(t.col_indices() + c_off)
        val_parts.append(t.values())
        nnz_off = crow_parts[-1][-1]
        c_off += t.size(-1)
    return torch.sparse_csr_tensor(torch.cat(crow_parts), torch.cat(col_parts), torch.cat(val_parts),
                                   size=(rows_total, cols_total))


def sparse_block_diag_split(sparse_block_diag_tensor: torch.Tensor, *shapes: Tuple[int, int]) -> tuple:
    """Inverse of :func:`sparse_block_diag`. Reference: ``utils/utils.py:648-790``."""
    layout = sparse_block_diag_tensor.layout
    if layout not in (torch.sparse_coo, torch.sparse_csr):
        raise ValueError("Input tensor layout not supported. Only sparse_coo and sparse_csr are supported.")
    if not all(len(s) == 2 for s in shapes):
        raise ValueError("All shapes must be two-dimensional (rows, cols).")
    total_rows = sum(s[0] for s in shapes)
    total_cols = sum(s[1] for s in shapes)
    in_rows, in_cols = sparse_block_diag_tensor.size(-2), sparse_block_diag_tensor.size(-1)
    if (total_rows, total_cols) != (in_rows, in_cols):
        raise ValueError(
            f"Sum of provided block shapes ({total_rows}, {total_cols}) does not match "
            f"input tensor size ({in_rows}, {in_cols}).")
    row_starts = [0]
    col_starts = [0]
    for r, c in shapes:
        row_starts.append(row_starts[-1] + r)
        col_starts.append(col_starts[-1] + c)

    if layout == torch.sparse_coo:
        t = sparse_block_diag_tensor if sparse_block_diag_tensor.is_coalesced() else sparse_block_diag_tensor.coalesce()
        idx, vals = t.indices(), t.values()
        dev = idx.device
        # coalesced => rows are sorted, so each block is one contiguous slice; one host sync for all cuts
        cuts = torch.searchsorted(idx[0].contiguous(), torch.tensor(row_starts, device=dev)).tolist()
        blocks = []
        for k, (r, c) in enumerate(shapes):
            sl = slice(cuts[k], cuts[k + 1])
            sub = idx[:, sl]
            keep = (sub[1] >= col_starts[k]) & (sub[1] < col_starts[k + 1])  # off-block entries are dropped
            off = torch.tensor([[row_starts[k]], [col_starts[k]]], dtype=torch.int64, device=dev)
            blocks.append(torch.sparse_coo_tensor(sub[:, keep] - off, vals[sl][keep], size=(r, c), device=dev,
                                                  dtype=vals.dtype))
        return tuple(blocks)

    t = sparse_block_diag_tensor
    crow, ccol, vals = t.crow_indices(), t.col_indices(), t.values()
    ptrs = crow[torch.tensor(row_starts, device=crow.device)].tolist()  # one sync instead of 2 per block
    blocks = []
    for k, (r, c) in enumerate(shapes):
        sl = slice(ptrs[k], ptrs[k + 1])
        blocks.append(torch.sparse_csr_tensor(crow[row_starts[k]: row_starts[k + 1] + 1] - ptrs[k],
                                              ccol[sl] - col_starts[k], vals[sl], size=(r, c), device=t.device,
                                              dtype=vals.dtype))
    return tuple(blocks)
