"""GPU parity of the ``batch_sparse_mv`` glue against fixtures produced by the reference's
``_batch_sparse_mv(sparse_mm, ...)`` (tests/golden/make_golden_batch_mv.py): forward, grad wrt the sparse values
(bit-exact pattern), grad wrt the dense vectors, through transposed / permuted views of the operand."""
import os

import numpy as np
import pytest
import torch

from helpers import TOL

pytestmark = pytest.mark.gpu

GOLDEN = np.load(os.path.join(os.path.dirname(__file__), "golden", "batch_mv_cases.npz"))
CASES = [str(c) for c in GOLDEN["__cases__"]]


@pytest.mark.parametrize("name", CASES)
def test_batch_sparse_mv_matches_reference(name):
    from torchsparsegradutils_b200 import batch_sparse_mv

    g = lambda s: GOLDEN[f"{name}/{s}"]  # noqa: E731
    dev = torch.device("cuda:0")
    shape = tuple(int(x) for x in g("shape"))
    vals = torch.from_numpy(g("values")).to(dev)
    if str(g("layout")) == "coo":
        A = torch.sparse_coo_tensor(torch.from_numpy(g("indices")).to(dev), vals, shape)
    else:
        A = torch.sparse_csr_tensor(torch.from_numpy(g("crow")).to(dev), torch.from_numpy(g("col")).to(dev), vals, shape)
    A.requires_grad_(True)
    v = torch.from_numpy(g("bvec")).to(dev).requires_grad_(True)
    out = batch_sparse_mv(A, v)
    assert out.shape == v.shape
    out.backward(torch.from_numpy(g("G")).to(dev))
    tol = TOL[v.dtype]
    torch.testing.assert_close(out.detach().cpu(), torch.from_numpy(g("out")), **tol)
    torch.testing.assert_close(v.grad.cpu(), torch.from_numpy(g("grad_bvec")), **tol)
    if str(g("layout")) == "coo":
        gA = A.grad.coalesce()
        assert torch.equal(gA.indices().cpu(), torch.from_numpy(g("gradA_indices")))
        torch.testing.assert_close(gA.values().cpu(), torch.from_numpy(g("gradA_values")), **tol)
    else:
        assert torch.equal(A.grad.crow_indices(), A.crow_indices()) and torch.equal(A.grad.col_indices(), A.col_indices())
        torch.testing.assert_close(A.grad.values().cpu(), torch.from_numpy(g("gradA_values")), **tol)


def test_rsample_shaped_call_reads_views_in_place():
    """k sample vectors against one (n, n) CSR factor, the shape `SparseMultivariateNormal.rsample` issues
    (distributions/sparse_multivariate_normal.py:365): result equals the dense product."""
    from torchsparsegradutils_b200 import batch_sparse_mv

    n, k = 4096, 32
    dense = torch.tril(torch.rand(n, n, device="cuda:0") * (torch.rand(n, n, device="cuda:0") < 0.01))
    A = dense.to_sparse_csr()
    eps = torch.rand(k, n, device="cuda:0")
    x = batch_sparse_mv(A, eps)
    torch.testing.assert_close(x, (dense.double() @ eps.double().t()).t().float(), rtol=1e-5, atol=1e-5)
