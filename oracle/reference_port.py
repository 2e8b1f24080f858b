"""CPU port of the reference's sparse_mm data flow on top of torch CPU ops -- TEST INFRASTRUCTURE /
CPU-BASELINE ONLY (bench.py's ``cpu_baseline`` and ``--impl reference`` legs; never the product path).

The reference's arithmetic is not its own: it calls PyTorch ATen (``pyproject.toml:23`` torch>=2.5;
2.11.0+cu128 here), i.e. MKL / OpenMP kernels on the host.  This port issues the *same ATen calls in
the same order* as ``torchsparsegradutils/sparse_matmul.py`` so that timing it on the GPU box's host
cores measures what the reference's CPU path costs there:

  forward   block-diagonal assembly of batched A (utils/utils.py:604-645 for CSR, :570-602 for COO),
            B.reshape(-1, K) (:153), torch.sparse.mm (:155)
  grad_A    row expansion by repeat_interleave (:190-192), two index_select (:201-202),
            elementwise product and sum over K (:205)
  grad_B    torch.sparse.mm(A.t(), G) (:229)

Parity status: PINNED -- tests/test_oracle_golden.py::test_reference_port_matches_golden checks it
against the fixtures generated from the real reference.  It is written as plain functions (no
autograd.Function) because only the timing and the numbers matter here.
"""
from __future__ import annotations

import torch


def _block_diag_csr(crow: torch.Tensor, col: torch.Tensor, val: torch.Tensor, n: int, m: int):
    """Batched CSR (b, .) -> one (b*n, b*m) CSR, the way the reference stitches it item by item."""
    b = crow.shape[0]
    crow_parts, col_parts, val_parts = [], [], []
    running = None
    for t in range(b):
        c = crow[t]
        crow_parts.append(c if t == 0 else c[1:] + running)
        col_parts.append(col[t] + t * m)
        val_parts.append(val[t])
        running = crow_parts[-1][-1].clone()
    return torch.sparse_csr_tensor(torch.cat(crow_parts), torch.cat(col_parts), torch.cat(val_parts), size=(b * n, b * m))


def _block_diag_coo(A: torch.Tensor):
    """Batched COO -> block-diagonal COO; every item is coalesced first (utils/utils.py:580)."""
    b, n, m = A.shape
    rows, cols, vals = [], [], []
    for t in range(b):
        item = A[t]
        item = item if item.is_coalesced() else item.coalesce()
        r, c = item.indices()
        rows.append(r + t * n)
        cols.append(c + t * m)
        vals.append(item.values())
    return torch.sparse_coo_tensor(torch.stack([torch.cat(rows), torch.cat(cols)]), torch.cat(vals), size=(b * n, b * m))


def forward_backward(A: torch.Tensor, B: torch.Tensor, G: torch.Tensor, need_gradA: bool = True, need_gradB: bool = True):
    """Returns (C, gradA_values_in_reference_order, gradB) for CPU tensors."""
    batched = B.dim() == 3
    if batched:
        b, n, m = A.shape
        K = B.shape[-1]
        if A.layout == torch.sparse_csr:
            Af = _block_diag_csr(A.crow_indices(), A.col_indices(), A.values(), n, m)
        else:
            Af = _block_diag_coo(A)
        Bf = B.reshape(-1, K)
        Gf = G.reshape(-1, K)
    else:
        Af, Bf, Gf = A, B, G
    C = torch.sparse.mm(Af, Bf)
    if batched:
        C = C.view(b, n, K)

    gA = gB = None
    if need_gradA:
        if Af.layout == torch.sparse_coo:
            ridx, cidx = Af._indices()
        else:
            cidx = Af.col_indices()
            crow = Af.crow_indices()
            ridx = torch.repeat_interleave(torch.arange(Af.size(0)), crow[1:] - crow[:-1])
        gsel = Gf.index_select(0, ridx)
        bsel = Bf.index_select(0, cidx)
        gA = (gsel * bsel).sum(dim=1)
        if batched and Af.layout == torch.sparse_csr:
            gA = gA.view(b, -1)
    if need_gradB:
        gB = torch.sparse.mm(Af.t(), Gf)
        if batched:
            gB = gB.view(B.shape)
    return C, gA, gB
