"""GPU: BASELINE.json's full-size workloads against the ORACLE on a sampled sub-problem (SURVEY.md section 8(c)(3):
"oracle on a row-sampled subset -- random 1 % of rows for C and grad_A, random 1 % of the columns of A for grad_B").

The CUDA path runs the whole problem once.  For the sample, the entries of the sampled rows (columns) are cut out
on the GPU, their column (row) ids are renumbered densely, and only the dense rows they touch travel to the host,
where ``oracle.spmm_csr`` / ``oracle.sddmm`` / ``oracle.spmm_t`` (C restatement, fp64 accumulation) compute the
truth on exactly those entries.

Tolerances are north_star's: fp32 rtol 1e-5 / atol 1e-6, bf16 1e-2 against the oracle on fp32-upcast inputs.
They are applied element-wise on U(0,1) dense operands (the reference's own test distribution,
tests/test_sparse_matmul.py:81,110 -- sums without cancellation).  On the benchmark distribution (N(0,1) operands,
benchmarks/sparse_mm_rand.py:75) individual outputs are cancellation-heavy sums for which no fp32 summation order
meets a per-element rtol; there the same check is made relative to the reference's own arithmetic: the reference's
data flow (``oracle/reference_port.py``, the same ATen CPU calls in fp32) is run on the same sub-problem and the
CUDA result's worst excess over the north_star tolerance must not be larger than 3x the reference's own worst excess.
"""
import os
import sys

import numpy as np
import pytest
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import workloads as W  # noqa: E402

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
TOL32 = dict(rtol=1e-5, atol=1e-6)
TOLBF = dict(rtol=1e-2, atol=1e-2)


def _fwd_bwd(A, B, G):
    from torchsparsegradutils_b200 import sparse_mm

    A = A.detach().requires_grad_(True)
    B = B.detach().requires_grad_(True)
    C = sparse_mm(A, B)
    C.backward(G)
    return C.detach(), A.grad, B.grad


def _np32(t):
    return t.detach().float().cpu().numpy()


def _row_subproblem(crow, col, val, rows):
    """Entries of the sampled rows: (sub_rowptr, entry ids in A, unique cols, compact col per entry), on device."""
    crow = crow.long()
    lens = crow[rows + 1] - crow[rows]
    sub_rowptr = torch.zeros(rows.numel() + 1, dtype=torch.int64, device=crow.device)
    sub_rowptr[1:] = lens.cumsum(0)
    total = int(sub_rowptr[-1])
    owner = torch.repeat_interleave(torch.arange(rows.numel(), device=crow.device), lens)
    eid = crow[rows][owner] + (torch.arange(total, device=crow.device) - sub_rowptr[:-1][owner])
    ucols, ccol = torch.unique(col[eid].long(), return_inverse=True)
    return sub_rowptr, eid, owner, ucols, ccol


def _col_subproblem(crow, col, cols_sample, m):
    """Entries of A that sit in the sampled columns: entry ids, compact column, unique rows, compact row."""
    lut = torch.full((m,), -1, dtype=torch.int64, device=col.device)
    lut[cols_sample] = torch.arange(cols_sample.numel(), device=col.device)
    ccol_all = lut[col.long()]
    eid = torch.nonzero(ccol_all >= 0).flatten()
    ccol = ccol_all[eid]
    rows = torch.searchsorted(crow.long().contiguous(), eid, right=True) - 1
    urows, crow_compact = torch.unique(rows, return_inverse=True)
    return eid, ccol, urows, crow_compact


def _excess(x, truth, tol):
    return float(np.maximum(np.abs(x.astype(np.float64) - truth.astype(np.float64))
                            - (tol["atol"] + tol["rtol"] * np.abs(truth.astype(np.float64))), 0.0).max(initial=0.0))


def _judge(name, ours, truth, ref32, tol, strict):
    ex = _excess(ours, truth, tol)
    if strict or ref32 is None:
        assert ex == 0.0, f"{name}: worst excess over rtol {tol['rtol']} / atol {tol['atol']} is {ex:.3e}"
        return
    ex_ref = _excess(ref32, truth, tol)
    print(f"[{name}] worst excess over the north_star tolerance: CUDA {ex:.3e}, reference fp32 data flow {ex_ref:.3e}")
    assert ex <= 3.0 * ex_ref + 1e-6, (f"{name}: CUDA excess {ex:.3e} vs the reference's own fp32 excess {ex_ref:.3e} "
                                      "on the same sub-problem")


def _check_sampled(A2d, B, G, C, gvals, gB, tol, strict, seed, frac=0.01, with_ref=True):
    """A2d: 2-D CSR on the GPU; B (m, K), G (n, K), C (n, K), gvals (nnz,), gB (m, K)."""
    from oracle import oracle as orc
    from oracle import reference_port as rp

    n, m = A2d.shape
    crow, col, val = A2d.crow_indices(), A2d.col_indices(), A2d.values()
    g = torch.Generator().manual_seed(seed)
    # ---- 1 % of the rows: C and grad_A
    rows = torch.randperm(n, generator=g)[: max(int(n * frac), 8)].sort().values.to(DEV)
    sub_rowptr, eid, owner, ucols, ccol = _row_subproblem(crow, col, val, rows)
    rp_np, cc_np = sub_rowptr.cpu().numpy(), ccol.cpu().numpy()
    v_np, B_np, G_np = _np32(val[eid]), _np32(B[ucols]), _np32(G[rows])
    truth_C = orc.spmm_csr(rp_np, cc_np, v_np, B_np, acc64=True)
    truth_gA = orc.sddmm(owner.cpu().numpy(), cc_np, G_np, B_np, acc64=True)
    ref_C = ref_gA = None
    if with_ref and not strict:
        A_sub = torch.sparse_csr_tensor(torch.from_numpy(rp_np), torch.from_numpy(cc_np), torch.from_numpy(v_np),
                                        (rows.numel(), ucols.numel()))
        rC, rgA, _ = rp.forward_backward(A_sub, torch.from_numpy(B_np), torch.from_numpy(G_np), need_gradB=False)
        ref_C, ref_gA = rC.numpy(), rgA.numpy()
    _judge("C", _np32(C[rows]), truth_C, ref_C, tol, strict)
    _judge("grad_A", _np32(gvals[eid]), truth_gA, ref_gA, tol, strict)
    # ---- 1 % of the columns of A: grad_B
    cols = torch.randperm(m, generator=g)[: max(int(m * frac), 8)].sort().values.to(DEV)
    eid, ccol, urows, crow_c = _col_subproblem(crow, col, cols, m)
    v_np, G_np = _np32(val[eid]), _np32(G[urows])
    r_np, c_np = crow_c.cpu().numpy(), ccol.cpu().numpy()
    truth_gB = orc.spmm_t(r_np, c_np, v_np, G_np, cols.numel(), acc64=True)
    ref_gB = None
    if with_ref and not strict:
        A_sub = torch.sparse_coo_tensor(torch.from_numpy(np.stack([r_np, c_np])), torch.from_numpy(v_np),
                                        (urows.numel(), cols.numel())).coalesce().to_sparse_csr()
        _, _, rgB = rp.forward_backward(A_sub, torch.zeros(cols.numel(), G_np.shape[1]), torch.from_numpy(G_np),
                                        need_gradA=False)
        ref_gB = rgB.numpy()
    _judge("grad_B", _np32(gB[cols]), truth_gB, ref_gB, tol, strict)


def _dense(shape_nm, K, dt, dist, seed, lead=()):
    n, m = shape_nm
    g = torch.Generator(device=DEV).manual_seed(seed)
    draw = torch.rand if dist == "uniform" else torch.randn
    B = draw(lead + (m, K), generator=g, device=DEV, dtype=torch.float32).to(dt)
    G = draw(lead + (n, K), generator=g, device=DEV, dtype=torch.float32).to(dt)
    return B, G


@pytest.mark.parametrize("dist", ["uniform", "normal"])
def test_config2_sampled_oracle(dist):
    """batched CSR b=8, 65536^2, 16 nnz/row, K=128 fp32 (BASELINE configs[1]), two of the eight items."""
    A = W.uniform_rows_csr(8, 65536, 65536, 16, torch.float32, torch.int32, DEV, seed=2)
    B, G = _dense((65536, 65536), 128, torch.float32, dist, 100, lead=(8,))
    C, gA, gB = _fwd_bwd(A, B, G)
    for t in (0, 5):
        At = torch.sparse_csr_tensor(A.crow_indices()[t], A.col_indices()[t], A.values()[t], (65536, 65536))
        _check_sampled(At, B[t], G[t], C[t], gA.values()[t], gB[t], TOL32, dist == "uniform", seed=10 + t)


@pytest.mark.parametrize("dist", ["uniform", "normal"])
def test_config3_stencil_128_sampled_oracle(dist):
    """27-point stencil on 128^3 (55.7 M nnz), K=32 fp32, B column-major as rsample passes it (BASELINE configs[2])."""
    D = 128
    A = W.stencil27_csr(D, torch.float32, torch.int32, DEV, seed=3)
    assert A._nnz() == (3 * D - 2) ** 3 == 55742968
    n = D ** 3
    B, G = _dense((n, n), 32, torch.float32, dist, 101)
    Bc = B.t().contiguous().t()  # column-major view, same values
    C, gA, gB = _fwd_bwd(A, Bc, G)
    assert gB.stride() == Bc.stride() and C.is_contiguous()
    assert torch.equal(gA.crow_indices(), A.crow_indices()) and torch.equal(gA.col_indices(), A.col_indices())
    _check_sampled(A, B, G, C, gA.values(), gB, TOL32, dist == "uniform", seed=20)


@pytest.mark.parametrize("dist", ["uniform", "normal"])
def test_config4_rmat_scale22_bf16_sampled_oracle(dist):
    """R-MAT scale 22 (~65 M nnz, hub rows of > 10^5 entries), K=128 bf16 (BASELINE configs[3]): the size at which
    32-bit offsets and the merge-path tile carries are exercised.  Oracle on the fp32-upcast inputs, 1e-2."""
    from torchsparsegradutils_b200 import _native as nat
    from torchsparsegradutils_b200._pattern import csr_pattern

    A = W.rmat_csr(22, 16, torch.bfloat16, torch.int32, DEV, seed=4)
    n = A.shape[0]
    assert n == 1 << 22 and A._nnz() > 60_000_000
    assert csr_pattern(A).algo == nat.ALGO_MERGE
    B, G = _dense((n, n), 128, torch.bfloat16, dist, 102)
    if dist == "uniform":
        # hub rows / columns sum 10^5 terms: scale so bf16 outputs stay in a sane range
        B, G = (B / 64).to(torch.bfloat16), (G / 64).to(torch.bfloat16)
    C, gA, gB = _fwd_bwd(A, B, G)
    assert torch.equal(gA.crow_indices(), A.crow_indices()) and gA._nnz() == A._nnz()
    # bf16: the tolerance is relative to the fp32-upcast oracle; no CPU reference data flow exists for bf16 CSR
    _check_sampled(A, B, G, C, gA.values(), gB, TOLBF, True, seed=30, with_ref=False)


@pytest.mark.parametrize("dist", ["uniform", "normal"])
def test_config5_long_k_sampled_oracle(dist):
    """262144^2, 8 nnz/row, K=512 fp32 (BASELINE configs[4])."""
    A = W.uniform_rows_csr(None, 262144, 262144, 8, torch.float32, torch.int32, DEV, seed=5)
    B, G = _dense((262144, 262144), 512, torch.float32, dist, 103)
    C, gA, gB = _fwd_bwd(A, B, G)
    _check_sampled(A, B, G, C, gA.values(), gB, TOL32, dist == "uniform", seed=40)


def test_config5_bf16_sampled_oracle():
    A = W.uniform_rows_csr(None, 262144, 262144, 8, torch.bfloat16, torch.int32, DEV, seed=5)
    B, G = _dense((262144, 262144), 512, torch.bfloat16, "normal", 104)
    C, gA, gB = _fwd_bwd(A, B, G)
    _check_sampled(A, B, G, C, gA.values(), gB, TOLBF, True, seed=41, with_ref=False)
