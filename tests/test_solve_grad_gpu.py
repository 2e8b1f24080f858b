"""GPU parity of ``solve_grad_A`` (the SDDMM inside the solves' backward) against A.grad of
the reference's own ``sparse_triangular_solve`` / ``sparse_generic_solve`` (tests/golden/make_golden_solve_grad.py,
fp64): same pattern bit for bit, values at fp64 tolerance; fp32 inputs at the north_star tolerance."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

GOLDEN = np.load(os.path.join(os.path.dirname(__file__), "golden", "solve_grad_cases.npz"))
CASES = [str(c) for c in GOLDEN["__cases__"]]


def _A(g, dev, dtype):
    shape = tuple(int(x) for x in g("shape"))
    vals = torch.from_numpy(g("values")).to(dev, dtype)
    if str(g("layout")) == "coo":
        return torch.sparse_coo_tensor(torch.from_numpy(g("indices")).to(dev), vals, shape)
    return torch.sparse_csr_tensor(torch.from_numpy(g("crow")).to(dev), torch.from_numpy(g("col")).to(dev), vals, shape)


@pytest.mark.parametrize("dtype", [torch.float64, torch.float32])
@pytest.mark.parametrize("name", CASES)
def test_solve_grad_A_matches_reference(name, dtype):
    from torchsparsegradutils_b200 import solve_grad_A

    g = lambda s: GOLDEN[f"{name}/{s}"]  # noqa: E731
    dev = torch.device("cuda:0")
    A = _A(g, dev, dtype)
    x = torch.from_numpy(g("x")).to(dev, dtype)
    gradB = torch.from_numpy(g("gradB")).to(dev, dtype)
    gA = solve_grad_A(A, gradB, x, transpose=bool(g("transpose")))
    assert gA.layout == A.layout and gA.shape == A.shape and gA.dtype == dtype
    want = torch.from_numpy(g("gradA_values")).to(dtype)
    tol = dict(rtol=1e-12, atol=1e-12) if dtype == torch.float64 else dict(rtol=1e-5, atol=1e-5)
    if A.layout == torch.sparse_coo:
        assert torch.equal(gA._indices().cpu(), torch.from_numpy(g("indices")))
        torch.testing.assert_close(gA._values().cpu(), want, **tol)
    else:
        assert torch.equal(gA.crow_indices(), A.crow_indices()) and torch.equal(gA.col_indices(), A.col_indices())
        torch.testing.assert_close(gA.values().cpu(), want, **tol)


def test_solve_grad_A_vector_rhs():
    from torchsparsegradutils_b200 import solve_grad_A

    dev = "cuda:0"
    A = (torch.rand(50, 50, device=dev, dtype=torch.float64) * (torch.rand(50, 50, device=dev) < 0.2)).to_sparse_csr()
    gb, x = torch.rand(50, device=dev, dtype=torch.float64), torch.rand(50, device=dev, dtype=torch.float64)
    gA = solve_grad_A(A, gb, x)
    torch.testing.assert_close(gA.to_dense(), -(torch.outer(gb, x)) * (A.to_dense() != 0))
