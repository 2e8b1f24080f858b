"""GPU: PairwiseValueAssembler (one gather kernel) reproduces the reference PairwiseEncoder.__call__ bit for bit --
values, index tensors and the gradient with respect to the input volumes (fixtures: make_golden_encoder.py)."""
import os
from types import SimpleNamespace

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "encoder_cases.npz"))


def _encoder_like(name):
    """An object with the reference encoder's attribute names, on the GPU (what from_encoder duck-types on)."""
    k = lambda s: G[f"{name}/{s}"]  # noqa: E731
    csr = str(k("layout")) == "csr"
    t = lambda a: torch.from_numpy(a).to(DEV)  # noqa: E731
    ns = SimpleNamespace(offsets=[tuple(int(v) for v in o) for o in k("offsets")],
                         volume_shape=tuple(int(v) for v in k("volume_shape")),
                         layout=torch.sparse_csr if csr else torch.sparse_coo)
    if csr:
        ns.crow_indices, ns.col_indices, ns.csr_permutation = t(k("crow")), t(k("col")), t(k("perm"))
    else:
        ns.indices = t(k("indices"))
    return ns


@pytest.mark.parametrize("name", [str(c) for c in G["__cases__"]])
@pytest.mark.parametrize("tag", ["u", "b"])
@pytest.mark.parametrize("dname", ["f32", "f64"])
def test_value_assembly_matches_reference(name, tag, dname):
    from torchsparsegradutils_b200.encoders import PairwiseValueAssembler

    asm = PairwiseValueAssembler.from_encoder(_encoder_like(name))
    q = f"{name}/{tag}_{dname}/"
    vals = torch.from_numpy(G[q + "values_in"]).to(DEV).requires_grad_(True)
    A = asm(vals)
    out = A.values()
    assert torch.equal(out.detach().cpu(), torch.from_numpy(G[q + "values_out"]))  # a gather: exact
    if asm.layout == torch.sparse_csr:
        assert A.layout == torch.sparse_csr
        assert torch.equal(A.crow_indices().cpu(), torch.from_numpy(G[q + "crow_out"]))
        assert torch.equal(A.col_indices().cpu(), torch.from_numpy(G[q + "col_out"]))
        assert A.crow_indices().dtype == torch.from_numpy(G[q + "crow_out"]).dtype
    else:
        assert A.layout == torch.sparse_coo and A.is_coalesced()
        assert torch.equal(A.indices().cpu(), torch.from_numpy(G[q + "indices_out"]))
    (out * torch.from_numpy(G[q + "w"]).to(DEV)).sum().backward()
    assert torch.equal(vals.grad.cpu(), torch.from_numpy(G[q + "grad_in"]))  # a scatter: exact


def test_assembled_matrix_feeds_sparse_mm_and_gradients_reach_the_volumes():
    """encoder -> sparse_mm -> loss: gradients flow back to the per-offset value volumes, and repeated calls reuse one
    cached sparsity pattern (same index tensors every call)."""
    from torchsparsegradutils_b200 import _pattern, sparse_mm
    from torchsparsegradutils_b200.encoders import PairwiseValueAssembler

    name = "csr3d_tril_i32"
    asm = PairwiseValueAssembler.from_encoder(_encoder_like(name))
    n = asm.volume_numel
    vals = torch.randn(len(asm.offsets), *asm.volume_shape, device=DEV, requires_grad=True)
    B = torch.randn(n, 8, device=DEV)
    _pattern.clear_pattern_cache()
    for _ in range(3):
        vals.grad = None
        C = sparse_mm(asm(vals), B)
        C.sum().backward()
    assert len(_pattern._cache) == 1
    dense = torch.zeros(n, n, device=DEV)
    A = asm(vals.detach())
    crow, col = A.crow_indices().long(), A.col_indices().long()
    rows = torch.repeat_interleave(torch.arange(n, device=DEV), crow[1:] - crow[:-1])
    v2 = vals.detach().clone().requires_grad_(True)
    A2 = asm(v2)
    dense = torch.zeros(n, n, device=DEV).index_put((rows, col), A2.values())
    (dense @ B).sum().backward()
    torch.testing.assert_close(vals.grad, v2.grad, rtol=1e-5, atol=1e-5)


def test_validation_messages_match_reference():
    from torchsparsegradutils_b200.encoders import PairwiseValueAssembler

    asm = PairwiseValueAssembler.from_encoder(_encoder_like("coo2d_lower"))
    with pytest.raises(ValueError, match="values must have 4 dimensions"):
        asm(torch.randn(4, 5, device=DEV))
    with pytest.raises(ValueError, match="Spatial dimensions do not match"):
        asm(torch.randn(len(asm.offsets), 1, 4, 6, device=DEV))
    with pytest.raises(ValueError, match="must match number of offsets"):
        asm(torch.randn(len(asm.offsets) + 1, 1, 4, 5, device=DEV))
    with pytest.raises(ValueError, match="float32 or torch.float64"):
        asm(torch.randn(len(asm.offsets), 1, 4, 5, device=DEV).half())
