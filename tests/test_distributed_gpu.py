"""GPU, >= 2 devices (skipped on a single-GPU box): row-sharded sparse_mm with the overlapped grad_B
all-reduce over NCCL equals the single-GPU result."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        import sys

        sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
        import workloads as W
        from torchsparsegradutils_b200 import distributed as D
        from torchsparsegradutils_b200 import sparse_mm

        A = W.uniform_rows_csr(None, 20000, 3000, 12, torch.float32, torch.int32, dev, seed=9)
        B, G = W.dense_operands((20000, 3000), 64, torch.float32, dev, seed=10)
        bounds = D.nnz_balanced_row_blocks(A.crow_indices(), world)
        lo, hi = bounds[rank], bounds[rank + 1]
        A_loc = D.shard_rows_csr(A, lo, hi).requires_grad_(True)
        B_rep = B.clone().requires_grad_(True)
        C_loc = D.sparse_mm_row_sharded(A_loc, B_rep)
        C_loc.backward(G[lo:hi])
        A_full = A.detach().requires_grad_(True)
        B_full = B.clone().requires_grad_(True)
        C = sparse_mm(A_full, B_full)
        C.backward(G)
        s, e = int(A.crow_indices()[lo]), int(A.crow_indices()[hi])
        # a row block may take a different kernel variant than the whole matrix (fewer rows): same sums,
        # different lane tiling of K, so compare at fp32 rounding rather than bit for bit
        ok = (torch.allclose(C_loc, C[lo:hi], rtol=1e-5, atol=1e-5)
              and torch.allclose(A_loc.grad.values(), A_full.grad.values()[s:e], rtol=1e-5, atol=1e-5)
              and torch.allclose(B_rep.grad, B_full.grad, rtol=1e-5, atol=1e-5))
        # reduce-scatter mode: this rank's block of rows of grad_B is the full sum, the rest of B.grad is zero
        A_rs = D.shard_rows_csr(A, lo, hi).requires_grad_(True)
        B_rs = B.clone().requires_grad_(True)
        D.sparse_mm_row_sharded(A_rs, B_rs, grad_b="reduce_scatter").backward(G[lo:hi])
        blo, bhi = D.row_block_bounds(B.shape[0], world, rank)
        mask = torch.zeros(B.shape[0], dtype=torch.bool, device=dev)
        mask[blo:bhi] = True
        ok = (ok and torch.allclose(B_rs.grad[blo:bhi], B_full.grad[blo:bhi], rtol=1e-5, atol=1e-5)
              and bool((B_rs.grad[~mask] == 0).all()) and torch.allclose(A_rs.grad.values(), A_loc.grad.values()))
        # not overlapped variant gives the same numbers
        A_no = D.shard_rows_csr(A, lo, hi).requires_grad_(True)
        B_no = B.clone().requires_grad_(True)
        D.sparse_mm_row_sharded(A_no, B_no, overlap=False).backward(G[lo:hi])
        ok = ok and torch.equal(B_no.grad, B_rep.grad)
        if not ok:
            print("rank", rank, "max diffs", float((C_loc - C[lo:hi]).abs().max()),
                  float((A_loc.grad.values() - A_full.grad.values()[s:e]).abs().max()),
                  float((B_rep.grad - B_full.grad).abs().max()), flush=True)
        flags = [None] * world
        dist.all_gather_object(flags, bool(ok))
        if rank == 0:
            out.put(all(flags))
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_row_sharded_nccl_matches_single_gpu():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    out = ctx.SimpleQueue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(300)
        assert p.exitcode == 0
    assert out.get() is True


def _k_worker(rank, world, port, out):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        import sys

        sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
        import workloads as W
        from torchsparsegradutils_b200 import distributed as D
        from torchsparsegradutils_b200 import sparse_mm

        ok = True
        for layout in ("csr", "coo"):
            A = W.uniform_rows_csr(None, 20000, 3000, 12, torch.float32, torch.int32, dev, seed=9)
            if layout == "coo":
                A = A.to_sparse_coo()
            B, G = W.dense_operands((20000, 3000), 128, torch.float32, dev, seed=10)
            lo, hi = D.k_shard_bounds(128, world, rank, align=4)
            A_rep = A.detach().requires_grad_(True)
            B_loc = B[:, lo:hi].contiguous().requires_grad_(True)
            C_loc = D.sparse_mm_k_sharded(A_rep, B_loc)
            C_loc.backward(G[:, lo:hi].contiguous())
            A_full = A.detach().requires_grad_(True)
            B_full = B.clone().requires_grad_(True)
            C = sparse_mm(A_full, B_full)
            C.backward(G)
            vals = lambda t: t.values() if t.layout == torch.sparse_csr else t._values()  # noqa: E731
            # C and grad_B blocks: same per-element sums; grad_A: the K-sum is split into `world` partial sums
            ok = ok and (torch.allclose(C_loc, C[:, lo:hi], rtol=1e-5, atol=1e-5)
                         and torch.allclose(B_loc.grad, B_full.grad[:, lo:hi], rtol=1e-5, atol=1e-5)
                         and torch.allclose(vals(A_rep.grad), vals(A_full.grad), rtol=1e-5, atol=1e-4)
                         and A_rep.grad.layout == A.layout)
        flags = [None] * world
        dist.all_gather_object(flags, bool(ok))
        if rank == 0:
            out.put(all(flags))
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_k_sharded_nccl_matches_single_gpu():
    """Dense-column sharding: local forward / grad_B, grad_A values all-reduced over NCCL (overlapped with the
    grad_B SpMM) equal the single-GPU results."""
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    out = ctx.SimpleQueue()
    procs = [ctx.Process(target=_k_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(300)
        assert p.exitcode == 0
    assert out.get() is True
