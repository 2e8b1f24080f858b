"""ctypes/numpy front-end of the C oracle -- TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs may
import this module; nothing under ``torchsparsegradutils_b200/`` does.

Parity status: PINNED against fixtures generated from the real reference
(``tests/golden/make_golden.py`` -> ``tests/golden/*.npz``; checked by
``tests/test_oracle_golden.py``).

The high-level entry point :func:`sparse_mm_fwd_bwd` restates the data flow of the
reference's ``SparseMatMul`` (``torchsparsegradutils/sparse_matmul.py:141-234``):
batched inputs behave as a block-diagonal product (``:151-153`` /
``utils/utils.py:474-645``), unbatched COO gradients are produced per *stored* entry in
storage order (``:184-185``, ``:209``), batched COO items are coalesced first
(``utils/utils.py:580``) so their gradient lives on the sorted unique pattern
(``utils/utils.py:716-752``), CSR gradients reuse crow/col (``:211``).
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SRC = os.path.join(_HERE, "tsgu_oracle.c")
_LIB = os.path.join(_HERE, "liboracle.so")
_lib = None


def build(force: bool = False) -> str:
    """Compile the C restatement with gcc (seconds)."""
    if force or not os.path.exists(_LIB) or os.path.getmtime(_LIB) < os.path.getmtime(_SRC):
        subprocess.check_call(
            ["gcc", "-O2", "-std=c99", "-shared", "-fPIC", "-fvisibility=hidden", "-o", _LIB, _SRC]
        )
    return _LIB


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build())
    return _lib


def _p(a: np.ndarray):
    return a.ctypes.data_as(ctypes.c_void_p)


def _i64(a) -> np.ndarray:
    return np.ascontiguousarray(np.asarray(a), dtype=np.int64)


def _vals(a, dt) -> np.ndarray:
    return np.ascontiguousarray(np.asarray(a), dtype=dt)


def _suffix(dt) -> str:
    return {np.dtype(np.float32): "f32", np.dtype(np.float64): "f64"}[np.dtype(dt)]


I64 = ctypes.c_int64


def spmm_csr(rowptr, col, vals, B, acc64: bool = True) -> np.ndarray:
    """C = A @ B for one CSR matrix; B may be any 2-D strided numpy array."""
    B = np.asarray(B)
    dt = B.dtype
    rowptr, col, vals = _i64(rowptr), _i64(col), _vals(vals, dt)
    n, K = rowptr.shape[0] - 1, B.shape[1]
    C = np.zeros((n, K), dtype=dt)
    isz = B.itemsize
    getattr(lib(), "orc_spmm_csr_" + _suffix(dt))(
        I64(n), I64(K), _p(rowptr), _p(col), _p(vals), _p(B),
        I64(B.strides[0] // isz), I64(B.strides[1] // isz), _p(C), ctypes.c_int(int(acc64)))
    return C


def sddmm(row, col, G, B, acc64: bool = True) -> np.ndarray:
    """out[e] = <G[row[e]], B[col[e]]> in storage order."""
    B = np.asarray(B)
    dt = B.dtype
    G = _vals(G, dt)
    row, col = _i64(row), _i64(col)
    out = np.zeros(row.shape[0], dtype=dt)
    isz = B.itemsize
    getattr(lib(), "orc_sddmm_" + _suffix(dt))(
        I64(row.shape[0]), I64(B.shape[1]), _p(row), _p(col), _p(G), I64(G.shape[1]), _p(B),
        I64(B.strides[0] // isz), I64(B.strides[1] // isz), _p(out), ctypes.c_int(int(acc64)))
    return out


def spmm_t(row, col, vals, G, m: int, acc64: bool = True) -> np.ndarray:
    """gradB = A^T @ G given A as (row, col, vals) triplets."""
    G = np.ascontiguousarray(G)
    dt = G.dtype
    row, col, vals = _i64(row), _i64(col), _vals(vals, dt)
    out = np.zeros((m, G.shape[1]), dtype=dt)
    getattr(lib(), "orc_spmm_t_" + _suffix(dt))(
        I64(row.shape[0]), I64(m), I64(G.shape[1]), _p(row), _p(col), _p(vals), _p(G),
        I64(G.shape[1]), _p(out), ctypes.c_int(int(acc64)))
    return out


def coo_sort(idx):
    """Stable lexicographic sort of (ndim, nnz) coordinates -> (sorted, perm)."""
    idx = _i64(idx)
    ndim, nnz = idx.shape
    out = np.zeros_like(idx)
    perm = np.zeros(nnz, dtype=np.int64)
    lib().orc_coo_sort(I64(ndim), I64(nnz), _p(idx), _p(out), _p(perm))
    return out, perm


def compress_rows(rows, n: int) -> np.ndarray:
    rows = _i64(rows)
    crow = np.zeros(n + 1, dtype=np.int64)
    lib().orc_compress_rows(I64(rows.shape[0]), _p(rows), I64(n), _p(crow))
    return crow


def decompress_crow(crow) -> np.ndarray:
    crow = _i64(crow)
    n = crow.shape[0] - 1
    rows = np.zeros(int(crow[-1]), dtype=np.int64)
    lib().orc_decompress_crow(I64(n), _p(crow), _p(rows))
    return rows


def csr_transpose(rowptr, col, m: int):
    """(rowptrT, colT, permT) of A^T, stable in A's storage order."""
    rowptr, col = _i64(rowptr), _i64(col)
    n = rowptr.shape[0] - 1
    nnz = int(rowptr[-1])
    rowptrT = np.zeros(m + 1, dtype=np.int64)
    colT = np.zeros(nnz, dtype=np.int64)
    permT = np.zeros(nnz, dtype=np.int64)
    lib().orc_csr_transpose(I64(n), I64(m), _p(rowptr), _p(col), _p(rowptrT), _p(colT), _p(permT))
    return rowptrT, colT, permT


def coo_to_csr(idx, n: int):
    """Restates convert_coo_to_csr_indices_values(values=None), utils/utils.py:236-346.

    Unbatched (2, nnz): -> crow (n+1,), col (nnz,), perm (nnz,).
    Batched (3, nnz) with equal nnz per item: -> crow (b, n+1), col (b, nnz/b), perm (b, nnz/b).
    """
    idx = _i64(idx)
    srt, perm = coo_sort(idx)
    if idx.shape[0] == 2:
        return compress_rows(srt[0], n), srt[1].copy(), perm
    batches = np.unique(srt[0])
    crow = np.stack([compress_rows(srt[1][srt[0] == b], n) for b in batches])
    nb = batches.shape[0]
    return crow, srt[2].reshape(nb, -1).copy(), perm.reshape(nb, -1).copy()


# --------------------------------------------------------------------------------------
# High-level: forward + backward of sparse_mm on numpy inputs
# --------------------------------------------------------------------------------------
def _coalesce(idx2, vals):
    """Sorted unique coordinates with duplicate values summed (torch coalesce())."""
    srt, perm = coo_sort(idx2)
    nnz = srt.shape[1]
    if nnz == 0:
        return srt, vals[perm]
    first = np.ones(nnz, dtype=bool)
    first[1:] = np.any(srt[:, 1:] != srt[:, :-1], axis=0)
    seg = np.cumsum(first) - 1
    out = np.zeros(int(seg[-1]) + 1, dtype=np.float64)
    np.add.at(out, seg, vals[perm].astype(np.float64))
    return srt[:, first], out.astype(vals.dtype)


def sparse_mm_fwd_bwd(layout: str, shape, B, G, *, indices=None, crow=None, col=None, values=None,
                      acc64: bool = True):
    """Forward + backward of sparse_mm.

    Returns dict(C=..., gradA_values=..., gradA_indices=... (COO only), gradB=...), with gradA in
    exactly the order/pattern the reference returns it.
    """
    B = np.asarray(B)
    G = np.asarray(G)
    dt = B.dtype
    values = np.asarray(values, dtype=dt)
    batched = len(shape) == 3
    n, m = shape[-2], shape[-1]
    if not batched:
        if layout == "coo":
            idx = _i64(indices)
            rows, cols = idx[0], idx[1]
            srt, perm = coo_sort(idx)
            rp = compress_rows(srt[0], n)
            C = spmm_csr(rp, srt[1], values[perm], B, acc64)
            gA = sddmm(rows, cols, G, B, acc64)  # storage order, duplicates each get the dot
            gB = spmm_t(rows, cols, values, G, m, acc64)
            return dict(C=C, gradA_values=gA, gradA_indices=idx, gradB=gB)
        rp, cl = _i64(crow), _i64(col)
        rows = decompress_crow(rp)
        return dict(C=spmm_csr(rp, cl, values, B, acc64), gradA_values=sddmm(rows, cl, G, B, acc64),
                    gradB=spmm_t(rows, cl, values, G, m, acc64))
    b = shape[0]
    Cs, gAs, gBs, gidx = [], [], [], []
    for t in range(b):
        if layout == "coo":
            idx = _i64(indices)
            sel = idx[0] == t
            idx2, v = _coalesce(idx[1:, sel], values[sel])
            rp, cl, rows = compress_rows(idx2[0], n), idx2[1], idx2[0]
            gidx.append(np.concatenate([np.full((1, idx2.shape[1]), t, dtype=np.int64), idx2]))
        else:
            rp, cl, v = _i64(crow)[t], _i64(col)[t], values[t]
            rows = decompress_crow(rp)
        Cs.append(spmm_csr(rp, cl, v, B[t], acc64))
        gAs.append(sddmm(rows, cl, G[t], B[t], acc64))
        gBs.append(spmm_t(rows, cl, v, G[t], m, acc64))
    out = dict(C=np.stack(Cs), gradB=np.stack(gBs))
    if layout == "coo":
        out["gradA_values"] = np.concatenate(gAs) if gAs else np.zeros(0, dt)
        out["gradA_indices"] = np.concatenate(gidx, axis=1) if gidx else np.zeros((3, 0), np.int64)
    else:
        out["gradA_values"] = np.stack(gAs)
    return out
