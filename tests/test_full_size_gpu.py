"""GPU: BASELINE.json's full-size workloads, checked through size-independent properties (the oracle
cannot run them whole): the adjoint identity <A B, G> = <grad_B, B> = <grad_A, A> that ties the three
kernels together, linearity, exact fp64 checks on sampled rows / entries / columns, and structural
invariants of the transpose (involution, sortedness)."""
import os
import sys

import pytest
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import workloads as W  # noqa: E402

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _fwd_bwd(A, B, G):
    from torchsparsegradutils_b200 import sparse_mm

    A = A.detach().requires_grad_(True)
    B = B.detach().requires_grad_(True)
    C = sparse_mm(A, B)
    C.backward(G)
    return C.detach(), A.grad, B.grad


def _dot(x, y):
    return float((x.double() * y.double()).sum())


def _check_adjoint(A, B, G, C, gA, gB, rel):
    vals = A.values() if A.layout == torch.sparse_csr else A._values()
    gvals = gA.values() if gA.layout == torch.sparse_csr else gA._values()
    ref = _dot(C, G)
    scale = float((C.double().abs() * G.double().abs()).sum())  # conditioning of the sum
    assert abs(_dot(gB, B) - ref) <= rel * scale
    assert abs(_dot(gvals, vals) - ref) <= rel * scale


def _sampled_exact(A2d, B, G, C, gvals, gB, rows, cols, tol):
    """fp64 reference on a sample: rows of C and grad_A, columns of grad_B."""
    crow, col, val = A2d.crow_indices().long(), A2d.col_indices().long(), A2d.values()
    for r in rows.tolist():
        s, e = int(crow[r]), int(crow[r + 1])
        c, v = col[s:e], val[s:e].double()
        torch.testing.assert_close(C[r].double(), (v[:, None] * B[c].double()).sum(0), **tol)
        torch.testing.assert_close(gvals[s:e].double(), (B[c].double() * G[r].double()).sum(1), **tol)
    rowid = torch.repeat_interleave(torch.arange(A2d.shape[0], device=DEV), crow[1:] - crow[:-1])
    for j in cols.tolist():
        m = col == j
        ref = (val[m].double()[:, None] * G[rowid[m]].double()).sum(0)
        torch.testing.assert_close(gB[j].double(), ref, **tol)


def test_config1_full_size_vs_oracle():
    """COO 4096x4096 at 1 % x dense 4096x64 fp32 (BASELINE configs[0]): small enough for the whole oracle."""
    import numpy as np

    from oracle import oracle as orc

    A = W.uniform_coo(4096, 4096, 167772, torch.float32, DEV, seed=1)
    B = torch.rand(4096, 64, device=DEV)
    G = torch.rand(4096, 64, device=DEV)
    C, gA, gB = _fwd_bwd(A, B, G)
    ref = orc.sparse_mm_fwd_bwd("coo", (4096, 4096), B.cpu().numpy(), G.cpu().numpy(),
                                indices=A._indices().cpu().numpy(), values=A._values().cpu().numpy())
    tol = dict(rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(C.cpu().numpy(), ref["C"], **tol)
    np.testing.assert_allclose(gB.cpu().numpy(), ref["gradB"], **tol)
    np.testing.assert_allclose(gA._values().cpu().numpy(), ref["gradA_values"], **tol)
    assert torch.equal(gA._indices(), A._indices()) and gA._nnz() == 167772


def test_config2_full_size_properties():
    """batched CSR b=8, 65536^2, 16 nnz/row, K=128 fp32 (BASELINE configs[1])."""
    A = W.uniform_rows_csr(8, 65536, 65536, 16, torch.float32, torch.int32, DEV, seed=2)
    B, G = W.dense_operands((8, 65536, 65536), 128, torch.float32, DEV, seed=100)
    C, gA, gB = _fwd_bwd(A, B, G)
    assert C.shape == (8, 65536, 128) and gB.shape == B.shape and gA._nnz() == 65536 * 16
    assert torch.equal(gA.crow_indices(), A.crow_indices()) and torch.equal(gA.col_indices(), A.col_indices())
    _check_adjoint(A, B, G, C, gA, gB, rel=2e-6)
    # linearity in B (same kernel, different data): A(2 B1 - 3 B2) = 2 A B1 - 3 A B2
    from torchsparsegradutils_b200 import sparse_mm

    B2 = torch.randn_like(B)
    lhs = sparse_mm(A, 2 * B - 3 * B2)
    rhs = 2 * C - 3 * sparse_mm(A, B2)
    torch.testing.assert_close(lhs, rhs, rtol=1e-4, atol=2e-4)
    # exact fp64 on a sample of rows / columns of two batch items
    g = torch.Generator().manual_seed(0)
    tol = dict(rtol=1e-5, atol=1e-5)  # N(0,1) operands: atol scaled to |x| ~ 10 (see tests/test_sparse_mm_gpu._check)
    for t in (0, 7):
        At = torch.sparse_csr_tensor(A.crow_indices()[t], A.col_indices()[t], A.values()[t], (65536, 65536))
        _sampled_exact(At, B[t], G[t], C[t], gA.values()[t], gB[t], torch.randint(0, 65536, (64,), generator=g),
                       torch.randint(0, 65536, (16,), generator=g), tol)


def test_config2_transpose_structure_full_size(monkeypatch):
    import torchsparsegradutils_b200 as tsgu
    from torchsparsegradutils_b200 import _pattern
    from torchsparsegradutils_b200._pattern import csr_pattern

    # the exact structure A.t().to_sparse_csr() would give: row padding (a layout optimisation of the
    # cached transpose, covered by test_padded_transpose_is_equivalent) is switched off here
    monkeypatch.setattr(_pattern, "_PAD_MIN_NNZ", 1 << 62)
    tsgu.clear_pattern_cache()
    A = W.uniform_rows_csr(8, 65536, 65536, 16, torch.float32, torch.int32, DEV, seed=2)
    P = csr_pattern(A)
    T = P.transpose()
    nnz = 8 * 65536 * 16
    assert T.rowptr.numel() == 8 * 65536 + 1 and int(T.rowptr[0]) == 0 and int(T.rowptr[-1]) == nnz
    assert bool((T.rowptr[1:] >= T.rowptr[:-1]).all())
    # permT is a permutation; entries of a transposed row are in increasing A-row order (stable sort)
    assert torch.equal(torch.sort(T.perm.long()).values, torch.arange(nnz, device=DEV))
    seg = torch.repeat_interleave(torch.arange(8 * 65536, device=DEV), (T.rowptr[1:] - T.rowptr[:-1]).long())
    same = seg[1:] == seg[:-1]
    assert bool((T.colind[1:][same] > T.colind[:-1][same]).all())
    # the entry permT[k] of A really sits in column (transposed row) seg[k] of item seg[k] // m
    cols_of_A = A.col_indices().reshape(-1).long()
    assert torch.equal(cols_of_A[T.perm.long()], seg % 65536)
    # involution: transposing again gives A's own structure back
    TT = T.transpose()
    flat_crow = torch.cat([(A.crow_indices()[:, :-1].long() + torch.arange(8, device=DEV)[:, None] * 65536 * 16).reshape(-1),
                           torch.tensor([nnz], device=DEV)])
    assert torch.equal(TT.rowptr.long(), flat_crow)
    assert torch.equal(TT.colind.long(), cols_of_A)
    assert torch.equal(TT.perm.long(), torch.arange(nnz, device=DEV))


@pytest.mark.default_layout_policy
def test_transposed_layout_is_optimised_on_the_second_use():
    """Default policy: the first backward runs on the plain transposed CSR (a pattern used once does not pay the
    block sort + padding), the second one upgrades it; the gradients do not change by a bit, and the mailbox host read
    the builds rely on returns what .tolist() would."""
    import torchsparsegradutils_b200 as tsgu
    from torchsparsegradutils_b200 import _native as nat
    from torchsparsegradutils_b200._pattern import csr_pattern

    probe = torch.tensor([3, -1, 2**40, 0], device=DEV)
    assert nat.host_read(probe) == probe.tolist() and nat.host_read(probe.int()[:2]) == [3, -1]
    tsgu.clear_pattern_cache()
    A = W.uniform_rows_csr(2, 65536, 65536, 16, torch.float32, torch.int32, DEV, seed=17)
    B, G = W.dense_operands((2, 65536, 65536), 64, torch.float32, DEV, seed=18)
    nnz = 2 * 65536 * 16
    C1, gA1, gB1 = _fwd_bwd(A, B, G)
    T1 = csr_pattern(A)._transpose
    assert T1 is not None and T1.row_map is None and not T1.padded and T1.nnz_total == nnz
    C2, gA2, gB2 = _fwd_bwd(A, B, G)
    T2 = csr_pattern(A)._transpose
    assert T2 is not T1 and T2.row_map is not None and T2.padded and T2.nnz_total > nnz
    C3, gA3, gB3 = _fwd_bwd(A, B, G)
    assert csr_pattern(A)._transpose is T2
    for x, y, z in ((C1, C2, C3), (gA1.values(), gA2.values(), gA3.values()), (gB1, gB2, gB3)):
        assert torch.equal(x, y) and torch.equal(y, z)
    # a one-step pattern announced as such stays plain; an announced long-lived one is optimised up front
    tsgu.clear_pattern_cache()
    tsgu.prepare_pattern(A, reuse=False)
    assert not csr_pattern(A)._transpose.padded
    tsgu.clear_pattern_cache()
    tsgu.prepare_pattern(A)
    assert csr_pattern(A)._transpose.padded
    tsgu.clear_pattern_cache()


def test_padded_transpose_is_equivalent(monkeypatch):
    """Rows of the cached transpose are padded to multiples of 4 with explicit zeros: same grad_B, bit for bit."""
    import torchsparsegradutils_b200 as tsgu
    from torchsparsegradutils_b200 import _pattern
    from torchsparsegradutils_b200._pattern import csr_pattern

    tsgu.clear_pattern_cache()
    A = W.uniform_rows_csr(2, 65536, 65536, 16, torch.float32, torch.int32, DEV, seed=7)
    B, G = W.dense_operands((2, 65536, 65536), 128, torch.float32, DEV, seed=8)
    T = csr_pattern(A).transpose(optimise=True)  # (by default the layout is optimised on the second request)
    lens = T.rowptr[1:] - T.rowptr[:-1]
    assert bool((lens % 4 == 0).all()) and T.nnz_total > 2 * 65536 * 16 and int((T.perm < 0).sum()) == T.nnz_total - 2 * 65536 * 16
    _, _, gB_pad = _fwd_bwd(A, B, G)
    monkeypatch.setattr(_pattern, "_PAD_MIN_NNZ", 1 << 62)
    tsgu.clear_pattern_cache()
    assert csr_pattern(A).transpose().nnz_total == 2 * 65536 * 16
    _, _, gB_ref = _fwd_bwd(A, B, G)
    tsgu.clear_pattern_cache()
    assert torch.equal(gB_pad, gB_ref)


def test_config5_long_k_properties():
    """262144^2, 8 nnz/row, K=512 (BASELINE configs[4]): SDDMM-dominated backward, fp32 and bf16."""
    for dt, rel in ((torch.float32, 2e-6), (torch.bfloat16, 4e-3)):
        A = W.uniform_rows_csr(None, 262144, 262144, 8, dt, torch.int32, DEV, seed=5)
        B, G = W.dense_operands((262144, 262144), 512, dt, DEV, seed=100)
        C, gA, gB = _fwd_bwd(A, B, G)
        _check_adjoint(A, B, G, C, gA, gB, rel=rel)
        if dt == torch.float32:
            g = torch.Generator().manual_seed(1)
            _sampled_exact(A, B, G, C, gA.values(), gB, torch.randint(0, 262144, (32,), generator=g),
                           torch.randint(0, 262144, (8,), generator=g), dict(rtol=1e-5, atol=3e-5))
        del A, B, G, C, gA, gB
        torch.cuda.empty_cache()


def test_config4_rmat_merge_path_properties():
    """R-MAT scale 20 (the config-4 generator at 1/4 size to bound test time), bf16 and fp32: the skew
    heuristic must pick merge-path, which must agree with row-split and satisfy the adjoint identity."""
    import torchsparsegradutils_b200 as tsgu
    from torchsparsegradutils_b200 import _native as nat
    from torchsparsegradutils_b200._pattern import csr_pattern

    tsgu.clear_pattern_cache()
    A = W.rmat_csr(20, 16, torch.float32, torch.int32, DEV, seed=4)
    n = A.shape[0]
    assert csr_pattern(A).algo == nat.ALGO_MERGE and csr_pattern(A).transpose().algo == nat.ALGO_MERGE
    B, G = W.dense_operands((n, n), 128, torch.float32, DEV, seed=100)
    B, G = B.abs(), G.abs()  # hub rows sum 1e5 terms: keep the sums well conditioned for the comparison
    C, gA, gB = _fwd_bwd(A, B, G)
    _check_adjoint(A, B, G, C, gA, gB, rel=2e-6)
    os.environ["TSGU_B200_ALGO"] = "rowsplit"
    try:
        tsgu.clear_pattern_cache()
        C2, gA2, gB2 = _fwd_bwd(A, B, G)
    finally:
        os.environ.pop("TSGU_B200_ALGO")
        tsgu.clear_pattern_cache()
    torch.testing.assert_close(C, C2, rtol=2e-5, atol=1e-6)
    torch.testing.assert_close(gA.values(), gA2.values(), rtol=2e-5, atol=1e-6)
    torch.testing.assert_close(gB, gB2, rtol=2e-5, atol=1e-6)
    Ab = torch.sparse_csr_tensor(A.crow_indices(), A.col_indices(), A.values().bfloat16(), A.shape)
    Cb, gAb, gBb = _fwd_bwd(Ab, B.bfloat16(), G.bfloat16())
    torch.testing.assert_close(Cb.float(), C, rtol=2e-2, atol=1e-2)
    torch.testing.assert_close(gBb.float(), gB, rtol=2e-2, atol=1e-2)


def test_config3_stencil_pattern_and_column_major_B():
    """27-point stencil (config 3 at 64^3 to bound test time): nnz = (3D-2)^3, B column-major as rsample
    passes it; result equals the contiguous-B result bit for bit and grad_B comes back in B's layout."""
    from torchsparsegradutils_b200 import sparse_mm

    D = 64
    A = W.stencil27_csr(D, torch.float32, torch.int32, DEV, seed=3)
    assert A._nnz() == (3 * D - 2) ** 3
    n = D ** 3
    eps = torch.randn(32, n, device=DEV)
    Bc = eps.t()  # column-major (n, 32) view
    G = torch.randn(n, 32, device=DEV)
    B1 = Bc.detach().requires_grad_(True)
    C1 = sparse_mm(A, B1)
    C1.backward(G)
    B2 = Bc.contiguous().requires_grad_(True)
    C2 = sparse_mm(A, B2)
    C2.backward(G)
    assert torch.equal(C1, C2) and torch.equal(B1.grad, B2.grad)
    assert B1.grad.stride() == B1.stride() and C1.is_contiguous()


def test_k_sliced_forward_matches_unsliced():
    """Uniform rows + a dense operand larger than L2 (128 MiB): the forward SpMM runs K in L2-resident slices
    (TSGU_ALGO_FLAG_KSLICE).  Slices are independent columns of C, so the result must equal the row-split
    kernel's (same CSR accumulation order) -- and the ragged transposed pattern must not ask for slicing."""
    from torchsparsegradutils_b200 import _native as nat
    from torchsparsegradutils_b200 import _ops
    from torchsparsegradutils_b200._pattern import csr_pattern

    n = m = 65536
    A = W.uniform_rows_csr(None, n, m, 8, torch.float32, torch.int32, DEV, seed=11)
    B, _ = W.dense_operands((n, m), 512, torch.float32, DEV, seed=12)
    pat = csr_pattern(A)
    assert pat.uniform_rows and not pat.transpose().uniform_rows
    C_sliced = _ops.spmm(pat, A.values(), B)  # default path: flag set by the launcher
    C_plain = _ops.spmm(pat, A.values(), B, algo=nat.ALGO_ROWSPLIT)
    torch.testing.assert_close(C_sliced, C_plain, rtol=1e-6, atol=1e-6)
