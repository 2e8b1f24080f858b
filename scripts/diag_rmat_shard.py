"""Diagnosis: the second row block of BASELINE config 4 (R-MAT scale 22, nnz-balanced 2-way split) on ONE GPU:
which kernel family the pattern heuristic picks and what each family costs (forward SpMM / SDDMM)."""
import os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import workloads as W
from torchsparsegradutils_b200 import _native as nat, _ops, distributed as D
from torchsparsegradutils_b200._pattern import csr_pattern, clear_pattern_cache

dev = torch.device("cuda:0")
A = W.rmat_csr(22, 16, torch.bfloat16, torch.int32, dev, seed=4)
bounds = D.nnz_balanced_row_blocks(A.crow_indices(), 2)
B, G = W.dense_operands(tuple(A.shape), 128, torch.bfloat16, dev, seed=100)
for r in (0, 1):
    lo, hi = bounds[r], bounds[r + 1]
    Al = D.shard_rows_csr(A, lo, hi)
    crow = Al.crow_indices()
    lens = (crow[1:] - crow[:-1])
    pat = csr_pattern(Al)
    print(f"shard {r}: rows {hi - lo} nnz {Al._nnz()} max_row {int(lens.max())} mean {Al._nnz() / (hi - lo):.2f} empty {int((lens == 0).sum())} algo {pat.algo}")
    Gl = G[lo:hi].contiguous()
    for name, algo in (("auto", nat.ALGO_AUTO), ("merge", nat.ALGO_MERGE)):
        for what in ("spmm", "sddmm"):
            fn = (lambda: _ops.spmm(pat, Al.values(), B, algo=algo)) if what == "spmm" else (lambda: _ops.sddmm(pat, Gl, B, None, pat.nnz_total, algo=algo))
            for _ in range(2):
                fn()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(5):
                fn()
            e1.record()
            torch.cuda.synchronize()
            print(f"   {what:6s} {name:6s} {e0.elapsed_time(e1) / 5:8.3f} ms")
