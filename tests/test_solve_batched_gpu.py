"""GPU: batched plumbing of the solves (block-diagonal operand, gradient straight in the batched layout) and the
least-squares gradient, against A.grad / operands produced by the reference's own sparse_triangular_solve and
sparse_generic_lstsq (fixtures: tests/golden/make_golden_solve_batched.py)."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "solve_batched_cases.npz"))
t = lambda a: torch.from_numpy(a).to(DEV)  # noqa: E731


def _A(name):
    k = lambda s: G[f"{name}/{s}"]  # noqa: E731
    shape = tuple(int(v) for v in k("shape"))
    if str(k("layout")) == "csr":
        return torch.sparse_csr_tensor(t(k("crow")), t(k("col")), t(k("values")), shape)
    return torch.sparse_coo_tensor(t(k("indices")), t(k("values")), shape)


@pytest.mark.parametrize("name", [str(c) for c in G["__cases__"]])
def test_block_diag_operand_matches_reference_assembly(name):
    from torchsparsegradutils_b200 import block_diag_operand

    bd = block_diag_operand(_A(name))
    k = lambda s: G[f"{name}/{s}"]  # noqa: E731
    assert bd.layout == torch.sparse_csr and tuple(bd.shape) == (int(k("shape")[0]) * int(k("shape")[1]),) * 2
    assert torch.equal(bd.crow_indices().cpu().long(), torch.from_numpy(k("bd_crow")))
    assert torch.equal(bd.col_indices().cpu().long(), torch.from_numpy(k("bd_col")))
    assert torch.equal(bd.values().cpu(), torch.from_numpy(k("bd_values")))


@pytest.mark.parametrize("name", [str(c) for c in G["__cases__"]])
def test_batched_solve_gradient_matches_reference(name):
    from torchsparsegradutils_b200 import solve_grad_A

    k = lambda s: G[f"{name}/{s}"]  # noqa: E731
    A = _A(name)
    gA = solve_grad_A(A, t(k("gradB")), t(k("x")), transpose=bool(k("transpose")))
    assert gA.layout == A.layout and gA.shape == A.shape
    if A.layout == torch.sparse_csr:
        assert torch.equal(gA.crow_indices().cpu(), torch.from_numpy(k("gradA_crow")))
        assert torch.equal(gA.col_indices().cpu(), torch.from_numpy(k("gradA_col")))
        got = gA.values()
    else:
        gA = gA.coalesce()
        assert torch.equal(gA.indices().cpu(), torch.from_numpy(k("gradA_indices")))
        got = gA.values()
    torch.testing.assert_close(got.cpu(), torch.from_numpy(k("gradA_values")), rtol=1e-12, atol=1e-13)


@pytest.mark.parametrize("name", [str(c) for c in G["__lstsq_cases__"]])
@pytest.mark.parametrize("dtype", [torch.float64, torch.float32])
def test_lstsq_gradient_matches_reference(name, dtype):
    from torchsparsegradutils_b200 import lstsq_grad_A

    k = lambda s: G[f"{name}/{s}"]  # noqa: E731
    A = _A(name).to(dtype)
    gA = lstsq_grad_A(A, t(k("gradB")).to(dtype), t(k("x")).to(dtype), t(k("B")).to(dtype), t(k("Apgb")).to(dtype))
    assert gA.layout == A.layout and gA.shape == A.shape
    got = gA.values() if A.layout == torch.sparse_csr else gA.coalesce().values()
    tol = dict(rtol=1e-11, atol=1e-12) if dtype == torch.float64 else dict(rtol=1e-4, atol=1e-5)
    torch.testing.assert_close(got.cpu().double(), torch.from_numpy(k("gradA_values")), **tol)
