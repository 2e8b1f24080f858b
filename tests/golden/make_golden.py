"""Generate golden fixtures from the REAL reference (run in the build container only).

    PYTHONPATH=/root/reference python tests/golden/make_golden.py

Imports cai4cai/torchsparsegradutils v0.2.5 from /root/reference (read-only), runs its own
``sparse_mm`` forward+backward and its index helpers on seeded inputs, and stores inputs and
outputs in ``tests/golden/*.npz``.  /root/reference does not exist on the GPU box, so the
tests only ever read the committed .npz files.

Case list mirrors the reference's parity contract: tests/test_sparse_matmul.py:16-24 (TEST_DATA),
tests/test_utils.py:53-133, tests/test_distributions.py:323-347 (strided B), the known answer in
Dockerfile.pip-install:47-52, plus the [probed] edge cases of SURVEY.md section 8(b).
"""
import os
import random
import sys

import numpy as np
import torch

sys.path.insert(0, "/root/reference")
from torchsparsegradutils import sparse_mm  # noqa: E402
from torchsparsegradutils.utils import (  # noqa: E402
    convert_coo_to_csr_indices_values,
    rand_sparse,
    sparse_block_diag,
    sparse_block_diag_split,
    stack_csr,
)
from torchsparsegradutils.utils.utils import (  # noqa: E402
    _compress_row_indices,
    _demcompress_crow_indices,
    _sort_coo_indices,
)

HERE = os.path.dirname(os.path.abspath(__file__))
torch.manual_seed(42)
np.random.seed(42)
random.seed(42)  # the reference's rand_sparse draws coordinates with Python's `random` (utils/random_sparse.py)


def npy(t):
    return t.detach().cpu().numpy()


def run_case(store, name, A, B, G=None):
    """Run reference fwd+bwd, record everything under prefix `name/`."""
    A = A.detach().clone().requires_grad_(True)
    Bl = B.detach().clone().requires_grad_(True)
    if B.stride() != B.contiguous().stride():  # keep the exact strided view
        base = B.detach().clone()
        Bl = base.requires_grad_(True)
    C = sparse_mm(A, Bl)
    if G is None:
        G = torch.rand_like(C)
    C.backward(G)
    p = name + "/"
    store[p + "shape"] = np.array(A.shape, dtype=np.int64)
    store[p + "B"] = npy(B)
    store[p + "B_strides"] = np.array(B.stride(), dtype=np.int64)
    store[p + "G"] = npy(G)
    store[p + "C"] = npy(C)
    store[p + "gradB"] = npy(Bl.grad)
    gA = A.grad
    if A.layout == torch.sparse_coo:
        store[p + "layout"] = np.array("coo")
        store[p + "indices"] = npy(A._indices())
        store[p + "values"] = npy(A._values())
        store[p + "gradA_indices"] = npy(gA._indices())
        store[p + "gradA_values"] = npy(gA._values())
        store[p + "gradA_coalesced"] = np.array(gA.is_coalesced())
    else:
        store[p + "layout"] = np.array("csr")
        store[p + "crow"] = npy(A.crow_indices())
        store[p + "col"] = npy(A.col_indices())
        store[p + "values"] = npy(A.values())
        store[p + "gradA_crow"] = npy(gA.crow_indices())
        store[p + "gradA_col"] = npy(gA.col_indices())
        store[p + "gradA_values"] = npy(gA.values())
    return name


def main():
    store = {}
    names = []
    TEST_DATA = [
        ("unbat0", (4, 6), (6, 2), 8),
        ("unbat1", (8, 16), (16, 10), 32),
        ("unbat2", (7, 4), (4, 9), 14),
        ("bat0", (1, 4, 6), (1, 6, 2), 8),
        ("bat1", (4, 8, 16), (4, 16, 10), 32),
        ("bat2", (11, 7, 4), (11, 4, 9), 14),
        ("vec128", (33, 40), (40, 128), 200),   # K that takes the vectorised kernels
        ("bvec32", (3, 20, 24), (3, 24, 32), 90),
    ]
    for tag, As, Bs, nnz in TEST_DATA:
        for layout, lname in ((torch.sparse_coo, "coo"), (torch.sparse_csr, "csr")):
            for vd, vname in ((torch.float32, "f32"), (torch.float64, "f64")):
                for idt, iname in ((torch.int32, "i32"), (torch.int64, "i64")):
                    if lname == "coo" and iname == "i32":
                        continue  # COO indices are always int64 in torch
                    A = rand_sparse(As, nnz, layout, indices_dtype=idt, values_dtype=vd)
                    B = torch.rand(*Bs, dtype=vd)
                    names.append(run_case(store, f"{tag}_{lname}_{vname}_{iname}", A, B))

    # known answer, Dockerfile.pip-install:47-52
    A = torch.tensor([[2.0, 0.0], [3.0, 4.0]]).to_sparse_coo()
    names.append(run_case(store, "known_answer_coo", A, torch.tensor([[5.0], [7.0]])))
    names.append(run_case(store, "known_answer_csr", A.to_sparse_csr(), torch.tensor([[5.0], [7.0]])))

    # uncoalesced COO with duplicates and unsorted entries (SURVEY 8b [probed])
    idx = torch.tensor([[2, 0, 2, 1, 0, 2], [1, 3, 1, 0, 3, 0]])
    val = torch.tensor([1.0, 2.0, 3.0, 4.0, 5.0, 6.0])
    A = torch.sparse_coo_tensor(idx, val, (3, 4))
    names.append(run_case(store, "dup_coo_f32", A, torch.rand(4, 5)))
    A = torch.sparse_coo_tensor(idx, val.double(), (3, 4))
    names.append(run_case(store, "dup_coo_f64", A, torch.rand(4, 5, dtype=torch.float64)))

    # ragged batched COO (items with different nnz, one empty)
    idx = torch.tensor([[0, 0, 2, 2, 2, 0], [1, 0, 2, 0, 1, 2], [2, 1, 0, 3, 3, 0]])
    val = torch.tensor([1.0, -2.0, 3.0, 0.5, 7.0, 4.0])
    A = torch.sparse_coo_tensor(idx, val, (3, 3, 4))
    names.append(run_case(store, "ragged_bcoo_f32", A, torch.rand(3, 4, 6)))
    # batched COO with duplicates inside an item (items are coalesced by the reference)
    idx = torch.tensor([[0, 0, 1, 1, 1, 0], [1, 1, 2, 0, 2, 2], [2, 2, 0, 3, 0, 0]])
    A = torch.sparse_coo_tensor(idx, val, (2, 3, 4))
    names.append(run_case(store, "dup_bcoo_f32", A, torch.rand(2, 4, 3)))

    # nnz = 0
    A = torch.sparse_coo_tensor(torch.zeros((2, 0), dtype=torch.int64), torch.zeros(0), (3, 4))
    names.append(run_case(store, "empty_coo", A, torch.rand(4, 2)))
    A = torch.sparse_csr_tensor(torch.zeros(4, dtype=torch.int64), torch.zeros(0, dtype=torch.int64),
                                torch.zeros(0), (3, 4))
    names.append(run_case(store, "empty_csr", A, torch.rand(4, 2)))

    # K = 1 and strided B (tests/test_distributions.py:323-347; _batch_sparse_mv :93-100)
    A = rand_sparse((9, 9), 30, torch.sparse_csr, indices_dtype=torch.int32)
    names.append(run_case(store, "k1_csr", A, torch.rand(9, 1)))
    V = torch.rand(5, 9)
    names.append(run_case(store, "strided_t_csr", A, V.t()))
    A = rand_sparse((9, 9), 30, torch.sparse_coo)
    names.append(run_case(store, "strided_t_coo", A, V.t()))
    Ab = rand_sparse((3, 9, 9), 20, torch.sparse_csr, indices_dtype=torch.int64)
    V = torch.rand(6, 3, 9)
    names.append(run_case(store, "strided_perm_bcsr", Ab, V.permute(1, 2, 0)))
    Ab = rand_sparse((3, 9, 9), 20, torch.sparse_coo)
    names.append(run_case(store, "strided_perm_bcoo", Ab, V.permute(1, 2, 0)))
    # batched CSR built with stack_csr (README / docs idiom)
    A1 = rand_sparse((6, 5), 11, torch.sparse_csr)
    A2 = rand_sparse((6, 5), 11, torch.sparse_csr)
    names.append(run_case(store, "stacked_bcsr", stack_csr([A1, A2]), torch.rand(2, 5, 4)))

    store["__cases__"] = np.array(names)
    np.savez_compressed(os.path.join(HERE, "sparse_mm_cases.npz"), **store)
    print("sparse_mm cases:", len(names))

    # ---------------------------------------------------------------- index helpers
    u = {}
    # probed golden of SURVEY 8(a11)
    idx = torch.tensor([[2, 0, 1, 0], [1, 3, 0, 0]])
    crow, col, perm = convert_coo_to_csr_indices_values(idx, 3)
    u["a11/idx"], u["a11/n"] = npy(idx), np.array(3)
    u["a11/crow"], u["a11/col"], u["a11/perm"] = npy(crow), npy(col), npy(perm)
    unames = []
    for k, (shape, nnz) in enumerate([((4, 4), 12), ((8, 16), 32), ((7, 4), 14), ((64, 50), 500),
                                      ((2, 4, 4), 12), ((4, 8, 16), 32), ((5, 7, 4), 14)]):
        A = rand_sparse(shape, nnz, torch.sparse_coo)
        ind = A._indices()
        # shuffle the storage order so the sort has work to do
        sh = torch.randperm(ind.shape[1])
        ind = ind[:, sh].contiguous()
        s, p = _sort_coo_indices(ind)
        crow, col, perm = convert_coo_to_csr_indices_values(ind, shape[-2])
        nm = f"sort{k}"
        unames.append(nm)
        u[nm + "/idx"], u[nm + "/n"], u[nm + "/m"] = npy(ind), np.array(shape[-2]), np.array(shape[-1])
        u[nm + "/sorted"], u[nm + "/perm"] = npy(s), npy(p)
        u[nm + "/crow"], u[nm + "/col"], u[nm + "/csr_perm"] = npy(crow), npy(col), npy(perm)
        if len(shape) == 2:
            for idt, iname in ((torch.int32, "i32"), (torch.int64, "i64")):
                rows_sorted = s[0].to(idt)
                cr = _compress_row_indices(rows_sorted, shape[0])
                u[nm + f"/compress_{iname}"] = npy(cr)
                u[nm + f"/decompress_{iname}"] = npy(_demcompress_crow_indices(cr, shape[0]))
            # CSR transpose oracle: A.t().to_sparse_csr() keeps index dtype
            for idt, iname in ((torch.int32, "i32"), (torch.int64, "i64")):
                csr = torch.sparse_csr_tensor(crow.to(idt), col.to(idt),
                                              torch.arange(1, nnz + 1, dtype=torch.float64), shape)
                T = csr.t().to_sparse_csr()
                u[nm + f"/T_crow_{iname}"] = npy(T.crow_indices())
                u[nm + f"/T_col_{iname}"] = npy(T.col_indices())
                # values are 1..nnz, so values-1 IS the transpose permutation
                u[nm + f"/T_perm_{iname}"] = npy(T.values()).astype(np.int64) - 1
    u["__cases__"] = np.array(unames)

    # block-diag / split / stack (tests/test_utils.py:33-47, :137-156, :209-230)
    for lname, layout in (("coo", torch.sparse_coo), ("csr", torch.sparse_csr)):
        mats = [rand_sparse(s, z, layout) for s, z in (((4, 6), 8), ((3, 2), 4), ((5, 5), 9))]
        bd = sparse_block_diag(*mats)
        pre = f"bd_{lname}/"
        for i, t in enumerate(mats):
            u[pre + f"in{i}_dense"] = npy(t.to_dense())
            if lname == "coo":
                u[pre + f"in{i}_indices"], u[pre + f"in{i}_values"] = npy(t._indices()), npy(t._values())
            else:
                u[pre + f"in{i}_crow"], u[pre + f"in{i}_col"], u[pre + f"in{i}_values"] = (
                    npy(t.crow_indices()), npy(t.col_indices()), npy(t.values()))
        u[pre + "dense"] = npy(bd.to_dense())
        if lname == "coo":
            u[pre + "indices"], u[pre + "values"] = npy(bd._indices()), npy(bd._values())
        else:
            u[pre + "crow"], u[pre + "col"], u[pre + "values"] = (
                npy(bd.crow_indices()), npy(bd.col_indices()), npy(bd.values()))
        parts = sparse_block_diag_split(bd, (4, 6), (3, 2), (5, 5))
        for i, t in enumerate(parts):
            u[pre + f"split{i}_dense"] = npy(t.to_dense())
    same = [rand_sparse((4, 5), 7, torch.sparse_csr, indices_dtype=torch.int32) for _ in range(3)]
    st = stack_csr(same)
    for i, t in enumerate(same):
        u[f"stack/in{i}_crow"], u[f"stack/in{i}_col"], u[f"stack/in{i}_values"] = (
            npy(t.crow_indices()), npy(t.col_indices()), npy(t.values()))
    u["stack/crow"], u["stack/col"], u["stack/values"] = (
        npy(st.crow_indices()), npy(st.col_indices()), npy(st.values()))
    u["stack/dense"] = npy(st.to_dense())
    np.savez_compressed(os.path.join(HERE, "index_cases.npz"), **u)
    print("index cases:", len(unames))


if __name__ == "__main__":
    main()
