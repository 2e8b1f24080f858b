"""Shared helpers for the parity tests (test infrastructure)."""
import numpy as np
import torch

# north_star tolerances (BASELINE.json): fp32 rtol 1e-5 / atol 1e-6, bf16 1e-2; fp64 near machine precision
TOL = {
    torch.float32: dict(rtol=1e-5, atol=1e-6),
    torch.float64: dict(rtol=1e-12, atol=1e-13),
    torch.bfloat16: dict(rtol=1e-2, atol=1e-2),
}


def strided_from(arr: np.ndarray, strides_elems, device) -> torch.Tensor:
    """Rebuild on `device` a tensor with the logical values of `arr` and the given element strides."""
    t = torch.from_numpy(np.ascontiguousarray(arr))
    order = sorted(range(t.dim()), key=lambda d: -int(strides_elems[d]))
    base = t.permute(order).contiguous().to(device)
    inv = [order.index(d) for d in range(t.dim())]
    out = base.permute(inv)
    assert out.shape == t.shape
    return out


def sparse_from_golden(g, name, device, requires_grad=True):
    """Re-create the reference's input A of golden case `name` on `device`."""
    k = lambda s: g[f"{name}/{s}"]  # noqa: E731
    shape = tuple(int(x) for x in k("shape"))
    vals = torch.from_numpy(k("values")).to(device)
    if str(k("layout")) == "coo":
        A = torch.sparse_coo_tensor(torch.from_numpy(k("indices")).to(device), vals, shape)
    else:
        A = torch.sparse_csr_tensor(torch.from_numpy(k("crow")).to(device), torch.from_numpy(k("col")).to(device),
                                    vals, shape)
    return A.requires_grad_(requires_grad)


def rand_csr(n, m, nnz_per_row, batch=None, dtype=torch.float32, index_dtype=torch.int32, device="cuda", seed=0,
             ragged=False):
    """Seeded random CSR with distinct sorted columns per row (output contract of the reference's
    utils/random_sparse.py generators: unique coordinates, sorted CSR)."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    b = batch or 1
    crows, cols = [], []
    for _ in range(b):
        if ragged:
            cnt = torch.randint(0, 2 * nnz_per_row + 1, (n,), generator=g).clamp_(max=m)
            if batch is not None:  # batched CSR needs equal nnz per item: fix the total
                cnt = torch.full((n,), nnz_per_row)
        else:
            cnt = torch.full((n,), min(nnz_per_row, m))
        crow = torch.zeros(n + 1, dtype=torch.int64)
        crow[1:] = torch.cumsum(cnt, 0)
        keys = torch.rand(n, m, generator=g)
        order = keys.argsort(dim=1)
        mask = torch.arange(m).unsqueeze(0) < cnt.unsqueeze(1)
        sel = torch.where(mask, order, torch.full_like(order, m)).sort(dim=1).values
        col = sel[sel < m]
        crows.append(crow)
        cols.append(col)
    crow = torch.stack(crows) if batch is not None else crows[0]
    col = torch.stack(cols) if batch is not None else cols[0]
    vals = torch.rand(col.shape, generator=g, dtype=torch.float64).to(dtype)
    shape = (batch, n, m) if batch is not None else (n, m)
    return torch.sparse_csr_tensor(crow.to(index_dtype).to(device), col.to(index_dtype).to(device), vals.to(device),
                                   shape)
