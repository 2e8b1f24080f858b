"""Value assembly of ``PairwiseEncoder`` -- the step in front of ``sparse_mm`` for BASELINE config 3.

The reference's ``PairwiseEncoder.__call__`` (``encoders/pairwise_encoder.py:751-849``) turns per-offset value volumes
``(N, C, *spatial)`` into the values of the sparse precision / scale matrix in four materialising steps: a Python loop
that trims every offset's volume and flattens it (``_calc_values``, ``:731-749``), a ``torch.cat`` of the N pieces, an
``index_select`` by ``csr_permutation`` (``:837``; for COO a ``coalesce()`` sort, ``:830-832``), and -- for batched
input -- a ``repeat`` of the index tensors per call (``:840-841``, ``:819-826``).

Here the composition "trim -> flatten -> cat -> permute" is evaluated ONCE, on the index side, when the assembler is
built: ``source[e]`` = flat position inside the untrimmed ``(N, C, *spatial)`` input of the value that ends up at
stored entry ``e``.  Every call is then a single gather kernel (``tsgu_gather_values``) straight into CSR / coalesced
COO value order, the backward is the adjoint scatter (``tsgu_scatter_values``), and the batched index tensors are
zero-copy expanded views (CSR) or cached per batch size (COO), so nothing index-sized is rebuilt per call and
``sparse_mm``'s pattern cache keeps hitting.

The offset enumeration and the index pattern itself (``calc_pairwise_coo_indices_nd``, ``:383-505``) are one-off
construction-time work and stay with the caller: build the assembler from an existing encoder object
(:meth:`PairwiseValueAssembler.from_encoder`, duck-typed on the reference's attribute names) or from its arrays.
"""
from __future__ import annotations

from functools import reduce
from operator import mul
from typing import Optional, Sequence, Tuple

import torch

from . import _native as nat
from . import _ops
from ._pattern import _sort_coo


def _trimmed_positions(volume_shape: Sequence[int], offset: Sequence[int], device) -> torch.Tensor:
    """Flat positions (row-major over `volume_shape`) that ``_trim_nd(vol, offset).flatten()`` keeps, in its order
    (reference ``:15-84``: offset k > 0 keeps ``[k:]``, k < 0 keeps ``[:k]``)."""
    idx = torch.arange(reduce(mul, volume_shape), device=device, dtype=torch.int64).reshape(tuple(volume_shape))
    sl = tuple(slice(None if o < 0 else o, None if o > -1 else o) for o in offset)
    return idx[sl].reshape(-1)


class _AssembleValues(torch.autograd.Function):
    """values_out[b, e] = values_in[b].flatten()[source[e]]; backward scatters (source is injective)."""

    @staticmethod
    def forward(ctx, values, source, numel_in):
        ctx.source, ctx.numel_in, ctx.in_shape = source, numel_in, values.shape
        flat = values.reshape(-1, numel_in) if values.numel() else values.reshape(0, numel_in)
        flat = flat.contiguous()
        out = torch.empty((flat.shape[0], source.numel()), dtype=values.dtype, device=values.device)
        for b in range(flat.shape[0]):  # one gather per batch item, no index tensor scaled by the batch
            out[b] = _ops.gather_values(flat[b], source)
        return out

    @staticmethod
    def backward(ctx, grad):  # type: ignore[override]
        grad = grad.contiguous()
        gin = torch.stack([_ops.scatter_values(grad[b], ctx.source, ctx.numel_in) for b in range(grad.shape[0])])
        return gin.reshape(ctx.in_shape), None, None


class PairwiseValueAssembler:
    """``assembler(values)`` == ``PairwiseEncoder(...)(values)`` of the reference, one gather kernel per call.

    Parameters mirror the attributes of the reference encoder (``:686-712``): ``offsets`` (list of
    ``(c, *spatial)`` tuples in value order), ``volume_shape`` ``(C, *spatial)``, ``layout``, and the index pattern --
    ``indices`` (2, M) for COO, or ``crow_indices`` / ``col_indices`` / ``csr_permutation`` for CSR.
    """

    def __init__(self, offsets, volume_shape: Tuple[int, ...], layout, indices: Optional[torch.Tensor] = None,
                 crow_indices: Optional[torch.Tensor] = None, col_indices: Optional[torch.Tensor] = None,
                 csr_permutation: Optional[torch.Tensor] = None):
        self.offsets = [tuple(int(v) for v in o) for o in offsets]
        self.volume_shape = tuple(int(v) for v in volume_shape)
        self.volume_numel = reduce(mul, self.volume_shape)
        self.layout = layout
        ref = indices if layout == torch.sparse_coo else crow_indices
        if ref is None:
            raise ValueError("indices (COO) or crow_indices/col_indices/csr_permutation (CSR) are required")
        if not ref.is_cuda:
            raise RuntimeError("PairwiseValueAssembler runs on CUDA tensors only (move the encoder with .to(device) first)")
        dev = ref.device
        # value order of the reference's _calc_values: offset blocks concatenated, each trimmed and flattened
        cat_src = torch.cat([k * self.volume_numel + _trimmed_positions(self.volume_shape, o, dev)
                             for k, o in enumerate(self.offsets)])
        self.numel_in = len(self.offsets) * self.volume_numel
        small = self.numel_in < 2**31 - 1
        if layout == torch.sparse_csr:
            if crow_indices is None or col_indices is None or csr_permutation is None:
                raise ValueError("CSR layout needs crow_indices, col_indices and csr_permutation")
            self.crow_indices, self.col_indices = crow_indices, col_indices
            src = cat_src.index_select(0, csr_permutation.to(torch.int64))
        elif layout == torch.sparse_coo:
            if cat_src.numel() != indices.shape[1]:
                raise ValueError("indices do not match the offsets / volume_shape")
            # the reference returns .coalesce(): entries in sorted coordinate order (pairs are unique: nothing is summed)
            ind64 = indices.to(torch.int64).contiguous()
            n = self.volume_numel
            perm, srt = _sort_coo(ind64, (n, n), 2, nat.I64, True)
            self.indices = srt.to(indices.dtype)
            src = cat_src.index_select(0, perm)
            self._batched_indices = {}
        else:
            raise ValueError("layout must be either torch.sparse_coo or torch.sparse_csr")
        self.source = src.to(torch.int32 if small else torch.int64).contiguous()

    @classmethod
    def from_encoder(cls, enc) -> "PairwiseValueAssembler":
        """Build from an object with the reference encoder's attributes (``offsets``, ``volume_shape``, ``layout``,
        ``indices`` | ``crow_indices``/``col_indices``/``csr_permutation``), already on the CUDA device."""
        return cls(enc.offsets, enc.volume_shape, enc.layout, indices=getattr(enc, "indices", None),
                   crow_indices=getattr(enc, "crow_indices", None), col_indices=getattr(enc, "col_indices", None),
                   csr_permutation=getattr(enc, "csr_permutation", None))

    def _check(self, values: torch.Tensor) -> bool:
        """Same checks and messages as the reference (``:769-799``); returns whether the input is batched."""
        spatial = len(self.volume_shape) - 1
        full = spatial + 2
        if len(values.shape) < full or len(values.shape) > full + 1:
            raise ValueError(f"values must have {full} dimensions (N, C, *spatial_dims) "
                             f"or {full + 1} dimensions (B, N, C, *spatial_dims)")
        got, want = values.shape[-spatial:], self.volume_shape[-spatial:]
        if tuple(got) != tuple(want):
            raise ValueError(f"Spatial dimensions do not match: expected {tuple(want)}, got {tuple(got)}")
        if values.shape[-full] != len(self.offsets):
            raise ValueError(f"Shape of values at index {-full} ({values.shape[-full]}) "
                             f"must match number of offsets ({len(self.offsets)})")
        if values.dtype not in [torch.float32, torch.float64]:
            raise ValueError("values must be either torch.float32 or torch.float64 for sparse tensors")
        return len(values.shape) == full + 1

    def __call__(self, values: torch.Tensor) -> torch.Tensor:
        batched = self._check(values)
        if not values.is_cuda:
            raise RuntimeError("PairwiseValueAssembler runs on CUDA tensors only; there is no CPU fallback")
        n = self.volume_numel
        out = _AssembleValues.apply(values if batched else values.unsqueeze(0), self.source, self.numel_in)
        b = out.shape[0]
        if self.layout == torch.sparse_csr:
            if not batched:
                return torch.sparse_csr_tensor(self.crow_indices, self.col_indices, out[0], size=(n, n))
            # zero-copy batch views of the index tensors (the reference repeats them on every call, :840-841)
            return torch.sparse_csr_tensor(self.crow_indices.expand(b, -1), self.col_indices.expand(b, -1), out,
                                           size=(b, n, n))
        if not batched:
            return torch.sparse_coo_tensor(self.indices, out[0], size=(n, n), is_coalesced=True)
        idx = self._batched_indices.get(b)
        if idx is None:  # batch-major sorted coordinates, built once per batch size
            m = self.indices.shape[1]
            bdim = torch.arange(b, dtype=self.indices.dtype, device=self.indices.device).repeat_interleave(m).unsqueeze(0)
            idx = self._batched_indices[b] = torch.cat([bdim, self.indices.repeat(1, b)])
        return torch.sparse_coo_tensor(idx, out.reshape(-1), size=(b, n, n), is_coalesced=True)
