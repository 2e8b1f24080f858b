"""Golden fixtures for the sparse-gradient (SDDMM) half of the solves' backward.

    PYTHONPATH=/root/reference python tests/golden/make_golden_solve_grad.py

Runs the reference's ``sparse_triangular_solve`` (``sparse_solve.py:151-250``) and ``sparse_generic_solve``
(``:427-515``) forward + backward on CPU in fp64 and stores, per case: A (pattern + values), the forward solution
x, the dense gradient gradB the reference computed with its own solver, and A.grad's values.  The product under
test only replaces the ``index_select x2 -> mul -> sum`` step (``:216-235``, ``:499-504``): given the same
(gradB, x) it must return the same values on the same pattern.  The build container only.
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, "/root/reference")
from torchsparsegradutils import sparse_generic_solve, sparse_triangular_solve  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
torch.manual_seed(11)


def tri(n, upper, dtype=torch.float64):
    M = torch.rand(n, n, dtype=dtype) * (torch.rand(n, n) < 0.2)
    M = torch.triu(M, 1) if upper else torch.tril(M, -1)
    return M + torch.diag(torch.rand(n, dtype=dtype) + 1.0)


def record(store, name, A, x, gradB, transpose):
    p = name + "/"
    layout = "coo" if A.layout == torch.sparse_coo else "csr"
    store[p + "layout"] = np.array(layout)
    store[p + "shape"] = np.array(A.shape, dtype=np.int64)
    store[p + "transpose"] = np.array(bool(transpose))
    Ad, gA = A.detach(), A.grad
    if layout == "coo":
        Ad, gA = Ad.coalesce(), gA.coalesce()
        store[p + "indices"] = Ad.indices().numpy()
        store[p + "values"] = Ad.values().numpy()
        assert torch.equal(gA.indices(), Ad.indices())
        store[p + "gradA_values"] = gA.values().numpy()
    else:
        store[p + "crow"] = Ad.crow_indices().numpy()
        store[p + "col"] = Ad.col_indices().numpy()
        store[p + "values"] = Ad.values().numpy()
        store[p + "gradA_values"] = gA.values().numpy()
    store[p + "x"] = x.detach().numpy()
    store[p + "gradB"] = gradB.numpy()


def main():
    store, cases = {}, []
    n, k = 37, 6
    for layout in ("coo", "csr"):
        for upper in (False, True):
            for transpose in (False, True):
                name = f"tri_{layout}_{'upper' if upper else 'lower'}_{'T' if transpose else 'N'}"
                dense = tri(n, upper)
                A = (dense.to_sparse_coo() if layout == "coo" else dense.to_sparse_csr()).requires_grad_(True)
                B = torch.randn(n, k, dtype=torch.float64, requires_grad=True)
                x = sparse_triangular_solve(A, B, upper=upper, transpose=transpose)
                x.backward(torch.rand_like(x))
                record(store, name, A, x, B.grad, transpose)
                cases.append(name)
        name = f"generic_{layout}"
        dense = tri(n, False) + tri(n, True)  # well conditioned (dominant diagonal), non-symmetric
        A = (dense.to_sparse_coo() if layout == "coo" else dense.to_sparse_csr()).requires_grad_(True)
        B = torch.randn(n, k, dtype=torch.float64, requires_grad=True)
        dense_solve = lambda A_, B_: torch.linalg.solve(A_.to_dense(), B_)  # noqa: E731
        dense_tsolve = lambda A_, B_: torch.linalg.solve(A_.to_dense().t(), B_)  # noqa: E731
        x = sparse_generic_solve(A, B, solve=dense_solve, transpose_solve=dense_tsolve)
        x.backward(torch.rand_like(x))
        record(store, name, A, x, B.grad, False)
        cases.append(name)
    store["__cases__"] = np.array(cases)
    np.savez_compressed(os.path.join(HERE, "solve_grad_cases.npz"), **store)
    print(len(cases), "cases")


if __name__ == "__main__":
    main()
