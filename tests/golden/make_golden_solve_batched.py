"""Golden fixtures for the BATCHED plumbing of the solves and the least-squares gradient
(rank 1, second sampled product), from the REAL reference (build container only):

    PYTHONPATH=/root/reference python tests/golden/make_golden_solve_batched.py

* ``sparse_triangular_solve`` on batched CSR / batched COO operands (``sparse_solve.py:162-250``): the block-diagonal
  CSR operand the reference assembles (``:172-178``), the solution, gradB, and A.grad in the batched layout it returns
  after ``sparse_block_diag_split`` + ``stack_csr`` / ``torch.stack`` (``:242-250``).
* ``sparse_generic_lstsq`` on a tall sparse matrix (``sparse_lstsq.py:166-265``) with dense solvers injected: x, gradB,
  Apgb = A^+ gradB and A.grad.
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, "/root/reference")
from torchsparsegradutils import sparse_generic_lstsq, sparse_triangular_solve  # noqa: E402
from torchsparsegradutils.utils import convert_coo_to_csr, sparse_block_diag  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
torch.manual_seed(13)


def tri(n, upper):
    M = torch.rand(n, n, dtype=torch.float64) * (torch.rand(n, n) < 0.25)
    M = torch.triu(M, 1) if upper else torch.tril(M, -1)
    return M + torch.diag(torch.rand(n, dtype=torch.float64) + 1.0)


def main():
    store, cases = {}, []
    b, n, k = 3, 19, 4
    for layout in ("csr", "coo"):
        for upper, transpose in ((False, False), (True, True)):
            name = f"btri_{layout}_{'upper' if upper else 'lower'}_{'T' if transpose else 'N'}"
            dense = torch.stack([tri(n, upper) for _ in range(b)])
            if layout == "csr":
                # batched CSR needs equal nnz per item: pad the pattern with explicit entries of a common mask
                mask = (dense != 0).any(0)
                items = [torch.where(mask, d, torch.zeros_like(d)) for d in dense]
                crow = torch.stack([torch.cat([torch.zeros(1, dtype=torch.int64), mask.sum(1).cumsum(0)])] * b)
                col = torch.stack([mask.nonzero()[:, 1]] * b)
                vals = torch.stack([d[mask] for d in items])
                A = torch.sparse_csr_tensor(crow, col, vals, (b, n, n)).requires_grad_(True)
            else:
                A = dense.to_sparse_coo().requires_grad_(True)  # sparse_dim 3
            B = torch.randn(b, n, k, dtype=torch.float64, requires_grad=True)
            x = sparse_triangular_solve(A, B, upper=upper, transpose=transpose)
            x.backward(torch.rand_like(x))
            p = name + "/"
            store[p + "layout"], store[p + "transpose"] = np.array(layout), np.array(bool(transpose))
            store[p + "shape"] = np.array(A.shape, dtype=np.int64)
            store[p + "x"], store[p + "gradB"] = x.detach().numpy(), B.grad.numpy()
            Ad = A.detach()
            bd = sparse_block_diag(*Ad)
            if layout == "coo":
                bd = convert_coo_to_csr(bd)
                Ac = Ad.coalesce()
                store[p + "indices"], store[p + "values"] = Ac.indices().numpy(), Ac.values().numpy()
                gA = A.grad.coalesce()
                store[p + "gradA_indices"], store[p + "gradA_values"] = gA.indices().numpy(), gA.values().numpy()
            else:
                store[p + "crow"], store[p + "col"], store[p + "values"] = crow.numpy(), col.numpy(), vals.numpy()
                store[p + "gradA_crow"] = A.grad.crow_indices().numpy()
                store[p + "gradA_col"] = A.grad.col_indices().numpy()
                store[p + "gradA_values"] = A.grad.values().numpy()
            store[p + "bd_crow"], store[p + "bd_col"], store[p + "bd_values"] = (
                bd.crow_indices().numpy(), bd.col_indices().numpy(), bd.values().numpy())
            cases.append(name)

    # least squares: tall full-rank A (45 x 17), dense pseudo-inverse solvers injected
    lcases = []
    for layout in ("csr", "coo"):
        name = f"lstsq_{layout}"
        nr, nc = 45, 17
        dense = torch.rand(nr, nc, dtype=torch.float64) * (torch.rand(nr, nc) < 0.3)
        dense[:nc] += torch.eye(nc, dtype=torch.float64) * 2.0
        A = (dense.to_sparse_csr() if layout == "csr" else dense.to_sparse_coo()).requires_grad_(True)
        B = torch.randn(nr, k, dtype=torch.float64, requires_grad=True)
        lstsq = lambda A_, B_: torch.linalg.pinv(A_.to_dense()) @ B_  # noqa: E731
        tlstsq = lambda A_, B_: torch.linalg.pinv(A_.to_dense().t()) @ B_  # noqa: E731
        x = sparse_generic_lstsq(A, B, lstsq=lstsq, transpose_lstsq=tlstsq)
        g = torch.rand_like(x)
        x.backward(g)
        gradB = B.grad
        p = name + "/"
        store[p + "layout"], store[p + "shape"] = np.array(layout), np.array(A.shape, dtype=np.int64)
        Ad, gA = A.detach(), A.grad
        if layout == "coo":
            Ad, gA = Ad.coalesce(), gA.coalesce()
            store[p + "indices"], store[p + "values"] = Ad.indices().numpy(), Ad.values().numpy()
        else:
            store[p + "crow"], store[p + "col"], store[p + "values"] = (
                Ad.crow_indices().numpy(), Ad.col_indices().numpy(), Ad.values().numpy())
        store[p + "gradA_values"] = gA.values().numpy()
        store[p + "x"], store[p + "B"], store[p + "gradB"] = x.detach().numpy(), B.detach().numpy(), gradB.numpy()
        store[p + "Apgb"] = lstsq(A.detach(), gradB).numpy()
        lcases.append(name)
    store["__cases__"], store["__lstsq_cases__"] = np.array(cases), np.array(lcases)
    np.savez_compressed(os.path.join(HERE, "solve_batched_cases.npz"), **store)
    print(len(cases), "batched solve cases,", len(lcases), "lstsq cases")


if __name__ == "__main__":
    main()
