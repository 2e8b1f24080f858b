"""Golden fixtures for the ``_batch_sparse_mv`` glue, from the REAL reference.

    PYTHONPATH=/root/reference python tests/golden/make_golden_batch_mv.py

Runs ``_batch_sparse_mv(sparse_mm, bmat, bvec)`` of cai4cai/torchsparsegradutils
(``distributions/sparse_multivariate_normal.py:16-102``) forward + backward on CPU for the four supported
rank combinations, COO and CSR, fp32 and fp64, and stores inputs/outputs in ``batch_mv_cases.npz``.
The build container only; the tests read the committed .npz.
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, "/root/reference")
from torchsparsegradutils import sparse_mm  # noqa: E402
from torchsparsegradutils.distributions.sparse_multivariate_normal import _batch_sparse_mv  # noqa: E402
from torchsparsegradutils.utils import stack_csr  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
torch.manual_seed(7)


def lower_tri(n, density, dtype):
    """Random lower-triangular (with diagonal) matrix, as SparseMultivariateNormal's scale_tril."""
    M = torch.rand(n, n, dtype=dtype) * (torch.rand(n, n) < density)
    M = torch.tril(M, -1) + torch.diag(torch.rand(n, dtype=dtype) + 0.5)
    return M


def main():
    store, cases = {}, []
    n, B, k = 23, 3, 5
    for dtype in (torch.float32, torch.float64):
        for layout in ("coo", "csr"):
            for rank in ("2x1", "2x2", "3x2", "3x3"):
                name = f"{rank}_{layout}_{str(dtype).split('.')[-1]}"
                batched = rank[0] == "3"
                if batched:
                    # equal nnz per item (batched CSR requirement, utils/utils.py:339-344): shared pattern, own values
                    pat = lower_tri(n, 0.3, dtype) != 0
                    dense = torch.stack([pat * (torch.rand(n, n, dtype=dtype) + 0.1) for _ in range(B)])
                    if layout == "coo":
                        A = dense.to_sparse_coo()
                    else:
                        A = stack_csr([d.to_sparse_csr() for d in dense])
                else:
                    dense = lower_tri(n, 0.3, dtype)
                    A = dense.to_sparse_coo() if layout == "coo" else dense.to_sparse_csr()
                shape = {"2x1": (n,), "2x2": (k, n), "3x2": (B, n), "3x3": (k, B, n)}[rank]
                bvec = torch.randn(shape, dtype=dtype)
                A = A.detach().requires_grad_(True)
                v = bvec.clone().requires_grad_(True)
                out = _batch_sparse_mv(sparse_mm, A, v)
                G = torch.rand(out.shape, dtype=dtype)
                out.backward(G)
                p = name + "/"
                store[p + "layout"] = np.array(layout)
                store[p + "shape"] = np.array(A.shape, dtype=np.int64)
                if layout == "coo":
                    Ac = A.detach()
                    store[p + "indices"] = Ac._indices().numpy()
                    store[p + "values"] = Ac._values().numpy()
                    gA = A.grad.coalesce()
                    store[p + "gradA_indices"] = gA.indices().numpy()
                    store[p + "gradA_values"] = gA.values().numpy()
                else:
                    Ac = A.detach()
                    store[p + "crow"] = Ac.crow_indices().numpy()
                    store[p + "col"] = Ac.col_indices().numpy()
                    store[p + "values"] = Ac.values().numpy()
                    store[p + "gradA_values"] = A.grad.values().numpy()
                store[p + "bvec"] = bvec.numpy()
                store[p + "G"] = G.numpy()
                store[p + "out"] = out.detach().numpy()
                store[p + "grad_bvec"] = v.grad.numpy()
                cases.append(name)
    store["__cases__"] = np.array(cases)
    np.savez_compressed(os.path.join(HERE, "batch_mv_cases.npz"), **store)
    print(len(cases), "cases")


if __name__ == "__main__":
    main()
