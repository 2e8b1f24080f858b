"""GPU: rsample_transform (sparse_mm + one fused layout/add pass) against SparseMultivariateNormal.rsample of the
reference with the same eps -- samples and gradients w.r.t. scale_tril values, loc and diagonal
(fixtures: tests/golden/make_golden_rsample.py)."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "rsample_cases.npz"))
t = lambda a: torch.from_numpy(a).to(DEV)  # noqa: E731


@pytest.mark.parametrize("name", [str(c) for c in G["__cases__"]])
def test_rsample_transform_matches_reference(name):
    from torchsparsegradutils_b200 import rsample_transform

    k = lambda s: G[f"{name}/{s}"]  # noqa: E731
    shape = tuple(int(v) for v in k("shape"))
    vals = t(k("values")).requires_grad_(True)
    if str(k("layout")) == "csr":
        A = torch.sparse_csr_tensor(t(k("crow")), t(k("col")), vals, shape)
    else:
        A = torch.sparse_coo_tensor(t(k("indices")), vals, shape)
    loc = t(k("loc")).requires_grad_(True)
    ldl = f"{name}/diag" in G.files
    diag = t(k("diag")).requires_grad_(True) if ldl else None
    x = rsample_transform(A, t(k("eps")), loc, diag)
    f64 = vals.dtype == torch.float64
    tol = dict(rtol=1e-12, atol=1e-13) if f64 else dict(rtol=1e-5, atol=1e-6)
    assert x.shape == k("x").shape and x.is_contiguous()
    torch.testing.assert_close(x.detach().cpu(), torch.from_numpy(k("x")), **tol)
    (x * t(k("w"))).sum().backward()
    gtol = dict(rtol=1e-11, atol=1e-12) if f64 else dict(rtol=1e-4, atol=1e-5)
    torch.testing.assert_close(vals.grad.cpu(), torch.from_numpy(k("grad_values")), **gtol)
    torch.testing.assert_close(loc.grad.cpu(), torch.from_numpy(k("grad_loc")), **gtol)
    if ldl:
        torch.testing.assert_close(diag.grad.cpu(), torch.from_numpy(k("grad_diag")), **gtol)


def test_rank_errors_match_batch_sparse_mv():
    from torchsparsegradutils_b200 import rsample_transform

    A = torch.eye(4, device=DEV).to_sparse_csr()
    with pytest.raises(ValueError, match="Invalid dimensions for bmat and bvec"):
        rsample_transform(A, torch.randn(2, 3, 4, device=DEV), torch.zeros(4, device=DEV))
