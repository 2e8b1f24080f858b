"""CPU: pins the oracle (oracle/tsgu_oracle.c + oracle/oracle.py) to fixtures produced by the
real reference (tests/golden/make_golden.py).  fp tolerances follow BASELINE.json's north_star:
rtol 1e-5 / atol 1e-6 for fp32 (tighter than the reference's own 1e-4, tests/test_config.py:22-37);
index structures are bit-exact."""
import os

import numpy as np
import pytest

from oracle import oracle as orc

GOLDEN = np.load(os.path.join(os.path.dirname(__file__), "golden", "sparse_mm_cases.npz"))
GIDX = np.load(os.path.join(os.path.dirname(__file__), "golden", "index_cases.npz"))
MM_CASES = [str(c) for c in GOLDEN["__cases__"]]
IDX_CASES = [str(c) for c in GIDX["__cases__"]]


def tol(dt):
    return dict(rtol=1e-5, atol=1e-6) if np.dtype(dt) == np.float32 else dict(rtol=1e-10, atol=1e-12)


def strided(arr, strides_elems):
    """Rebuild the exact strided view the reference was given."""
    order = np.argsort(-np.asarray(strides_elems), kind="stable")
    base = np.ascontiguousarray(arr.transpose(order))
    inv = np.argsort(order)
    return base.transpose(inv)


@pytest.mark.parametrize("name", MM_CASES)
@pytest.mark.parametrize("acc64", [True, False])
def test_oracle_matches_reference_sparse_mm(name, acc64):
    g = lambda k: GOLDEN[f"{name}/{k}"]  # noqa: E731
    layout = str(g("layout"))
    shape = tuple(int(x) for x in g("shape"))
    B = strided(g("B"), g("B_strides"))
    kw = dict(indices=g("indices")) if layout == "coo" else dict(crow=g("crow"), col=g("col"))
    out = orc.sparse_mm_fwd_bwd(layout, shape, B, g("G"), values=g("values"), acc64=acc64, **kw)
    t = tol(B.dtype)
    np.testing.assert_allclose(out["C"], g("C"), **t)
    np.testing.assert_allclose(out["gradB"], g("gradB"), **t)
    if layout == "coo":
        # reference returns batched-COO grads flagged uncoalesced but already in sorted order
        assert np.array_equal(out["gradA_indices"], g("gradA_indices"))
    np.testing.assert_allclose(out["gradA_values"], g("gradA_values"), **t)


def test_known_answer():
    # Dockerfile.pip-install:47-52 : [[2,0],[3,4]] @ [[5],[7]] = [[10],[43]]
    C = orc.spmm_csr([0, 1, 3], [0, 0, 1], [2.0, 3.0, 4.0], np.array([[5.0], [7.0]], dtype=np.float32))
    assert np.array_equal(C, np.array([[10.0], [43.0]], dtype=np.float32))


def test_probed_coo_to_csr_golden():
    crow, col, perm = orc.coo_to_csr(GIDX["a11/idx"], int(GIDX["a11/n"]))
    assert crow.tolist() == [0, 2, 3, 4] and col.tolist() == [0, 3, 0, 1] and perm.tolist() == [3, 1, 2, 0]
    assert np.array_equal(crow, GIDX["a11/crow"]) and np.array_equal(perm, GIDX["a11/perm"])


@pytest.mark.parametrize("name", IDX_CASES)
def test_index_builders_bit_exact(name):
    g = lambda k: GIDX[f"{name}/{k}"]  # noqa: E731
    idx, n, m = g("idx"), int(g("n")), int(g("m"))
    s, p = orc.coo_sort(idx)
    assert np.array_equal(s, g("sorted")) and np.array_equal(p, g("perm"))
    crow, col, perm = orc.coo_to_csr(idx, n)
    assert np.array_equal(crow, g("crow")) and np.array_equal(col, g("col"))
    assert np.array_equal(perm, g("csr_perm"))
    if idx.shape[0] == 2:
        cr = orc.compress_rows(s[0], n)
        assert np.array_equal(cr, g("compress_i64")) and np.array_equal(cr, g("compress_i32"))
        assert np.array_equal(orc.decompress_crow(cr), g("decompress_i64"))
        rT, cT, pT = orc.csr_transpose(crow, col, m)
        for it in ("i32", "i64"):
            assert np.array_equal(rT, g(f"T_crow_{it}"))
            assert np.array_equal(cT, g(f"T_col_{it}"))
            assert np.array_equal(pT, g(f"T_perm_{it}"))


@pytest.mark.parametrize("name", MM_CASES)
def test_reference_port_matches_golden(name):
    """The torch-CPU port used for the CPU baseline issues the reference's ATen calls: same numbers."""
    import torch

    from oracle import reference_port as rp

    g = lambda k: GOLDEN[f"{name}/{k}"]  # noqa: E731
    shape = tuple(int(x) for x in g("shape"))
    vals = torch.from_numpy(g("values"))
    if str(g("layout")) == "coo":
        A = torch.sparse_coo_tensor(torch.from_numpy(g("indices")), vals, shape)
    else:
        A = torch.sparse_csr_tensor(torch.from_numpy(g("crow")), torch.from_numpy(g("col")), vals, shape)
    B = torch.from_numpy(strided(g("B"), g("B_strides")))
    C, gA, gB = rp.forward_backward(A, B, torch.from_numpy(g("G")))
    t = tol(g("B").dtype)
    np.testing.assert_allclose(C.numpy(), g("C"), **t)
    np.testing.assert_allclose(gB.numpy(), g("gradB"), **t)
    np.testing.assert_allclose(gA.numpy().reshape(g("gradA_values").shape), g("gradA_values"), **t)


def test_oracle_sddmm_reproduces_reference_solve_gradients():
    """The oracle's SDDMM, negated, equals A.grad of the reference's triangular / generic solves
    (tests/golden/solve_grad_cases.npz; sparse_solve.py:216-235, :487-504) -- pins SURVEY 8(f) rank 1."""
    import os

    import numpy as np

    from oracle import oracle as orc

    G = np.load(os.path.join(os.path.dirname(__file__), "golden", "solve_grad_cases.npz"))
    for name in [str(c) for c in G["__cases__"]]:
        g = lambda s: G[f"{name}/{s}"]  # noqa: E731
        n = int(g("shape")[0])
        if str(g("layout")) == "coo":
            row, col = g("indices")
        else:
            crow, col = g("crow"), g("col")
            row = np.repeat(np.arange(n), np.diff(crow))
        X, Y = (g("x"), g("gradB")) if bool(g("transpose")) else (g("gradB"), g("x"))
        got = -np.einsum("ek,ek->e", X[row], Y[col])
        np.testing.assert_allclose(got, g("gradA_values"), rtol=1e-12, atol=1e-12)
        np.testing.assert_allclose(-orc.sddmm(row.astype(np.int64), col.astype(np.int64), X, Y), g("gradA_values"),
                                   rtol=1e-12, atol=1e-12)
