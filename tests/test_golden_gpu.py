"""GPU parity: the CUDA path (through the C ABI) against fixtures produced by the real reference.

Every case in tests/golden/sparse_mm_cases.npz is the reference's own `sparse_mm` forward+backward
(tests/golden/make_golden.py).  Index structures of the gradient are compared bit-exactly, values at
the north_star tolerance (fp32 rtol 1e-5 / atol 1e-6)."""
import os

import numpy as np
import pytest
import torch

from helpers import TOL, sparse_from_golden, strided_from

pytestmark = pytest.mark.gpu

GOLDEN = np.load(os.path.join(os.path.dirname(__file__), "golden", "sparse_mm_cases.npz"))
CASES = [str(c) for c in GOLDEN["__cases__"]]


@pytest.mark.parametrize("name", CASES)
def test_sparse_mm_matches_reference(name):
    from torchsparsegradutils_b200 import sparse_mm

    g = lambda s: GOLDEN[f"{name}/{s}"]  # noqa: E731
    dev = torch.device("cuda:0")
    A = sparse_from_golden(GOLDEN, name, dev)
    B = strided_from(g("B"), g("B_strides"), dev).requires_grad_(True)
    assert tuple(B.stride()) == tuple(int(s) for s in g("B_strides")) or B.numel() <= 1 or min(B.shape) == 1
    G = torch.from_numpy(g("G")).to(dev)

    C = sparse_mm(A, B)
    assert C.is_contiguous() and C.layout == torch.strided and C.dtype == B.dtype
    C.backward(G)
    tol = TOL[B.dtype]
    torch.testing.assert_close(C.detach().cpu(), torch.from_numpy(g("C")), **tol)
    torch.testing.assert_close(B.grad.cpu(), torch.from_numpy(g("gradB")), **tol)
    assert B.grad.shape == B.shape

    gA = A.grad
    assert gA.layout == A.layout and gA.shape == A.shape and gA.dtype == A.dtype
    if str(g("layout")) == "coo":
        assert torch.equal(gA._indices().cpu(), torch.from_numpy(g("gradA_indices")))
        assert gA.is_coalesced() == bool(g("gradA_coalesced"))
        torch.testing.assert_close(gA._values().cpu(), torch.from_numpy(g("gradA_values")), **tol)
    else:
        assert gA.crow_indices().dtype == A.crow_indices().dtype
        assert torch.equal(gA.crow_indices().cpu(), torch.from_numpy(g("gradA_crow")))
        assert torch.equal(gA.col_indices().cpu(), torch.from_numpy(g("gradA_col")))
        torch.testing.assert_close(gA.values().cpu(), torch.from_numpy(g("gradA_values")), **tol)
