"""CPU, world_size 2, gloo: partition and collective logic of torchsparsegradutils_b200.distributed.
The per-rank operator is injected (torch's own CPU sparse mm) because the product kernels are CUDA-only;
what is under test is the N>1 host logic: shard bounds, nnz-balanced row blocks, grad_B all-reduce."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from torchsparsegradutils_b200 import distributed as D


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _problem(seed=0, n=37, m=23, K=5, density=0.3):
    g = torch.Generator().manual_seed(seed)
    dense = torch.rand(n, m, generator=g, dtype=torch.float64) * (torch.rand(n, m, generator=g) < density)
    dense[5:12] = 0  # a stretch of empty rows
    return dense.to_sparse_csr(), torch.rand(m, K, generator=g, dtype=torch.float64), torch.rand(n, K, generator=g, dtype=torch.float64)


def test_batch_shard_bounds_cover_and_balance():
    for batch in (1, 7, 8, 11):
        for world in (1, 2, 3, 8):
            spans = [D.batch_shard_bounds(batch, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == batch
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        D.batch_shard_bounds(8, 2, 2)


def test_nnz_balanced_row_blocks():
    A, _, _ = _problem()
    crow = A.crow_indices()
    for world in (1, 2, 3, 5):
        b = D.nnz_balanced_row_blocks(crow, world)
        assert b[0] == 0 and b[-1] == A.shape[0] and len(b) == world + 1 and b == sorted(b)
        per = [int(crow[b[i + 1]] - crow[b[i]]) for i in range(world)]
        assert sum(per) == int(crow[-1])
        # balanced to within one row's worth of entries
        longest = int((crow[1:] - crow[:-1]).max())
        assert max(per) - min(per) <= 2 * longest + 1
    parts = [D.shard_rows_csr(A, lo, hi) for lo, hi in zip(b, b[1:])]
    assert torch.equal(torch.cat([p.to_dense() for p in parts]), A.to_dense())


def test_shard_batched_csr_and_coo():
    g = torch.Generator().manual_seed(1)
    dense = torch.rand(5, 4, 6, generator=g) * (torch.rand(5, 4, 6, generator=g) < 0.5)
    B = torch.rand(5, 6, 3, generator=g)
    coo = dense.to_sparse_coo()
    per = [dense[t].to_sparse_csr() for t in range(5)]
    got = []
    for r in range(2):
        Al, Bl = D.shard_batched(coo, 2, r, B)
        lo, hi = D.batch_shard_bounds(5, 2, r)
        assert torch.equal(Al.to_dense(), dense[lo:hi]) and torch.equal(Bl, B[lo:hi])
        got.append(Al.shape[0])
    assert got == [3, 2]
    assert all(p.layout == torch.sparse_csr for p in per)


def _row_sharded_worker(rank, world, port, out):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        A, B, G = _problem()
        bounds = D.nnz_balanced_row_blocks(A.crow_indices(), world)
        lo, hi = bounds[rank], bounds[rank + 1]
        A_loc = D.shard_rows_csr(A, lo, hi)
        Bl = B.clone().requires_grad_(True)
        C_loc = D.sparse_mm_row_sharded(A_loc, Bl, local_mm=torch.sparse.mm)
        C_loc.backward(G[lo:hi])
        # every replica of B must now hold the FULL A^T G
        ref_C = A.to_dense() @ B
        ref_gB = A.to_dense().t() @ G
        ok = torch.allclose(C_loc, ref_C[lo:hi]) and torch.allclose(Bl.grad, ref_gB)
        gathered = [None] * world
        dist.all_gather_object(gathered, bool(ok))
        if rank == 0:
            out.put(all(gathered))
    finally:
        dist.destroy_process_group()


def _batch_sharded_worker(rank, world, port, out):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = torch.Generator().manual_seed(3)
        b, n, m, K = 5, 6, 7, 3
        dense = torch.rand(b, n, m, generator=g, dtype=torch.float64) * (torch.rand(b, n, m, generator=g) < 0.4)
        B = torch.rand(b, m, K, generator=g, dtype=torch.float64)
        crow = torch.stack([dense[t].to_sparse_csr().crow_indices() for t in range(b)]) if False else None
        A = dense.to_sparse_coo()
        A_loc, B_loc = D.shard_batched(A, world, rank, B)
        lo, hi = D.batch_shard_bounds(b, world, rank)
        # no collective on the data path: the local product is the slice of the global product
        C_loc = torch.bmm(A_loc.to_dense(), B_loc)
        ok = torch.allclose(C_loc, torch.bmm(dense, B)[lo:hi])
        t = torch.tensor([float(C_loc.sum())], dtype=torch.float64)
        dist.all_reduce(t)  # only the harness reduces (checksum of checksums)
        ok = ok and torch.allclose(t, torch.bmm(dense, B).sum().reshape(1))
        gathered = [None] * world
        dist.all_gather_object(gathered, bool(ok))
        if rank == 0:
            out.put(all(gathered))
    finally:
        dist.destroy_process_group()


class _CpuSparseMM(torch.autograd.Function):
    """Per-rank operator for the gloo tests: the reference's data flow on CPU (oracle/reference_port.py) with a
    sparse gradient on A's pattern -- stands in for the CUDA sparse_mm, which cannot run here."""

    @staticmethod
    def forward(ctx, A, B):
        ctx.save_for_backward(A, B)
        return torch.sparse.mm(A.detach(), B.detach())

    @staticmethod
    def backward(ctx, G):
        from oracle import reference_port as ref

        A, B = ctx.saved_tensors
        _, gA, gB = ref.forward_backward(A.detach(), B.detach(), G)
        return torch.sparse_csr_tensor(A.crow_indices(), A.col_indices(), gA, A.shape), gB


def test_k_shard_bounds_cover_and_align():
    for K, align in ((512, 4), (10, 4), (128, 8), (7, 1), (3, 4)):
        for world in (1, 2, 3, 8):
            spans = [D.k_shard_bounds(K, world, r, align) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == K
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert all(lo % align == 0 for lo, _ in spans if lo < K)
    with pytest.raises(ValueError):
        D.k_shard_bounds(8, 2, 2)


def _k_sharded_worker(rank, world, port, out):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        A, B, G = _problem(K=6)
        lo, hi = D.k_shard_bounds(B.shape[1], world, rank)
        Al = A.detach().requires_grad_(True)  # replicated sparse operand
        Bl = B[:, lo:hi].clone().requires_grad_(True)
        C_loc = D.sparse_mm_k_sharded(Al, Bl, local_mm=_CpuSparseMM.apply)
        C_loc.backward(G[:, lo:hi].contiguous())
        dense = A.to_dense()
        ref_C, ref_gB = dense @ B, dense.t() @ G
        ref_gA = (G @ B.t()) * (dense != 0)
        # C and grad_B column blocks are local; every replica of A must hold the FULL sampled product
        ok = (torch.allclose(C_loc, ref_C[:, lo:hi]) and torch.allclose(Bl.grad, ref_gB[:, lo:hi])
              and torch.allclose(Al.grad.to_dense(), ref_gA) and torch.equal(Al.grad.col_indices(), A.col_indices()))
        gathered = [None] * world
        dist.all_gather_object(gathered, bool(ok))
        if rank == 0:
            out.put(all(gathered))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("worker", [_row_sharded_worker, _batch_sharded_worker, _k_sharded_worker],
                         ids=["row_sharded", "batch_sharded", "k_sharded"])
def test_world_size_2_gloo(worker):
    ctx = mp.get_context("spawn")
    out = ctx.SimpleQueue()
    port = _free_port()
    procs = [ctx.Process(target=worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert out.get() is True


def test_row_block_bounds_partition_rows():
    from torchsparsegradutils_b200.distributed import row_block_bounds

    for m, world in ((10, 4), (4194304, 8), (7, 8), (16, 2)):
        blocks = [row_block_bounds(m, world, r) for r in range(world)]
        assert blocks[0][0] == 0 and blocks[-1][1] == m
        assert all(b[0] <= b[1] for b in blocks) and all(blocks[i][1] == blocks[i + 1][0] for i in range(world - 1))
        per = -(-m // world)
        assert all(b[1] - b[0] <= per for b in blocks)
