"""Golden fixtures for the rsample affine map (the `+ eta` / `loc +` epilogues), from the REAL
reference distribution (build container only):

    PYTHONPATH=/root/reference python tests/golden/make_golden_rsample.py

``SparseMultivariateNormal(loc, diagonal, scale_tril).rsample(sample_shape)``
(distributions/sparse_multivariate_normal.py:354-389) with ``_standard_normal`` intercepted so that the eps it drew is
stored; a weighted sum of the sample is back-propagated to scale_tril's values, loc and diagonal.
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, "/root/reference")
import torchsparsegradutils.distributions.sparse_multivariate_normal as smvn  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
torch.manual_seed(17)
_drawn = {}
_orig = smvn._standard_normal


def _spy(shape, dtype, device):
    _drawn["eps"] = _orig(shape, dtype=dtype, device=device)
    return _drawn["eps"]


smvn._standard_normal = _spy


def tril(n, strict, dtype, batch=None):
    lead = () if batch is None else (batch,)
    M = torch.rand(*lead, n, n, dtype=dtype) * (torch.rand(n, n) < 0.3)
    M = torch.tril(M, -1)
    if not strict:
        M = M + torch.diag_embed(torch.rand(*lead, n, dtype=dtype) + 0.5)
    return M


def main():
    store, names = {}, []
    n = 11
    for dt, dname in ((torch.float32, "f32"), (torch.float64, "f64")):
        for layout in ("coo", "csr"):
            for batch in (None, 3):
                for ldl in (False, True):
                    for sshape in ((), (5,)):
                        name = f"{dname}_{layout}_{'b' if batch else 'u'}_{'ldl' if ldl else 'll'}_{'k5' if sshape else 'k0'}"
                        dense = tril(n, ldl, dt, batch)
                        if layout == "csr" and batch:  # batched CSR needs equal nnz per item: share one mask
                            mask = (dense != 0).any(0)
                            crow = torch.cat([torch.zeros(1, dtype=torch.int64), mask.sum(1).cumsum(0)])
                            col = mask.nonzero()[:, 1]
                            vals = torch.stack([d[mask] for d in dense]).requires_grad_(True)
                            A = torch.sparse_csr_tensor(crow.expand(batch, -1).contiguous(), col.expand(batch, -1).contiguous(), vals, (batch, n, n))
                        elif layout == "csr":
                            ref = dense.to_sparse_csr()
                            vals = ref.values().clone().requires_grad_(True)
                            A = torch.sparse_csr_tensor(ref.crow_indices(), ref.col_indices(), vals, ref.shape)
                        else:
                            ref = dense.to_sparse_coo().coalesce()
                            vals = ref.values().clone().requires_grad_(True)
                            A = torch.sparse_coo_tensor(ref.indices(), vals, ref.shape).coalesce()
                        lead = () if batch is None else (batch,)
                        loc = torch.randn(*lead, n, dtype=dt, requires_grad=True)
                        diag = (torch.rand(*lead, n, dtype=dt) + 0.5).requires_grad_(True) if ldl else None
                        d = smvn.SparseMultivariateNormal(loc, diagonal=diag, scale_tril=A)
                        x = d.rsample(sshape)
                        w = torch.randn_like(x)
                        (x * w).sum().backward()
                        p = name + "/"
                        store[p + "layout"], store[p + "shape"] = np.array(layout), np.array(A.shape, dtype=np.int64)
                        if layout == "csr":
                            store[p + "crow"], store[p + "col"] = A.crow_indices().detach().numpy(), A.col_indices().detach().numpy()
                        else:
                            store[p + "indices"] = A.indices().detach().numpy()
                        store[p + "values"], store[p + "grad_values"] = vals.detach().numpy(), vals.grad.numpy()
                        store[p + "loc"], store[p + "grad_loc"] = loc.detach().numpy(), loc.grad.numpy()
                        if ldl:
                            store[p + "diag"], store[p + "grad_diag"] = diag.detach().numpy(), diag.grad.numpy()
                        store[p + "eps"], store[p + "w"], store[p + "x"] = _drawn["eps"].numpy(), w.numpy(), x.detach().numpy()
                        names.append(name)
    store["__cases__"] = np.array(names)
    np.savez_compressed(os.path.join(HERE, "rsample_cases.npz"), **store)
    print(len(names), "cases")


if __name__ == "__main__":
    main()
