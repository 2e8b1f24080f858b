"""CUDA-graph replay of a ``sparse_mm`` forward + backward for a fixed sparsity pattern.

Enqueueing one eager step costs ~0.3-0.4 ms of host time (Python, ctypes, autograd); a small problem
(BASELINE config 1: three ~10 us kernels) or a strongly scaled batch shard is bound by that, not by the
GPU.  Training loops that keep the pattern and the shapes fixed (``tests/test_sparse_matmul.py:295-338``)
can capture the step once and replay it: config 1 goes from 0.32 ms to 0.031 ms per step.
"""
from __future__ import annotations

from typing import Tuple

import torch

from ._pattern import coo_pattern, csr_pattern, pin_pattern
from .sparse_matmul import sparse_mm


class GraphedSparseMM:
    """``C, grad_A_values, grad_B = step(values, B, G)`` replayed from one captured CUDA graph.

    Parameters are *templates*: ``A`` fixes the pattern (COO or CSR, unbatched or batched), ``B`` and
    ``G`` fix shapes / dtypes of the dense operand and of the upstream gradient.  Every call copies the
    new data into the graph's static buffers, replays, and returns the graph's static output tensors
    (valid until the next call).
    """

    def __init__(self, A: torch.Tensor, B: torch.Tensor, G: torch.Tensor, warmup: int = 3):
        if not (A.is_cuda and B.is_cuda and G.is_cuda):
            raise RuntimeError("GraphedSparseMM needs CUDA tensors")
        dev = B.device
        self._csr = A.layout == torch.sparse_csr
        with torch.no_grad():
            if self._csr:
                A_static = torch.sparse_csr_tensor(A.crow_indices(), A.col_indices(), A.values().clone(), A.shape)
            else:
                A_static = torch.sparse_coo_tensor(A._indices(), A._values().clone(), A.shape,
                                                   is_coalesced=A.is_coalesced())
        self._A = A_static.requires_grad_(True)
        self._B = B.detach().clone().contiguous().requires_grad_(True)
        self._G = G.detach().clone().contiguous()

        def step():
            self._A.grad = None
            self._B.grad = None
            C = sparse_mm(self._A, self._B)
            C.backward(self._G)
            return C

        side = torch.cuda.Stream(dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            # The pattern builds (sort / transpose, with their host syncs) are not capturable: build the pattern
            # here and PIN it to this object, so the capture below -- and nothing later -- depends on what the
            # LRU pattern cache happens to hold (it may have capacity 0, or evict the entry between steps).
            self._pattern = csr_pattern(self._A) if self._csr else coo_pattern(self._A)
            pin_pattern(self._pattern.cache_key, self._pattern)
            (self._pattern if self._csr else self._pattern.csr).transpose(optimise=True)  # final layout before capture
            for _ in range(max(warmup, 1)):
                step()
        torch.cuda.current_stream(dev).wait_stream(side)
        self._graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self._graph):
            self._C = step()
        gA = self._A.grad
        self._gA_values = gA.values() if self._csr else gA._values()
        self._gB = self._B.grad

    def _values(self) -> torch.Tensor:
        return self._A.values() if self._csr else self._A._values()

    @torch.no_grad()
    def __call__(self, values: torch.Tensor, B: torch.Tensor, G: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
        self._values().copy_(values.reshape(self._values().shape))
        self._B.copy_(B)
        self._G.copy_(G)
        self._graph.replay()
        return self._C.detach(), self._gA_values, self._gB
