"""GPU: index builders bit-exact against the reference's golden outputs, the oracle and torch."""
import os

import numpy as np
import pytest
import torch

from oracle import oracle as orc

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
GIDX = np.load(os.path.join(os.path.dirname(__file__), "golden", "index_cases.npz"))
CASES = [str(c) for c in GIDX["__cases__"]]


def _t(a):
    return torch.from_numpy(np.asarray(a)).to(DEV)


@pytest.mark.parametrize("name", CASES)
def test_builders_match_reference_golden(name):
    from torchsparsegradutils_b200.utils import utils as U

    g = lambda k: GIDX[f"{name}/{k}"]  # noqa: E731
    idx, n, m = _t(g("idx")), int(g("n")), int(g("m"))
    s, p = U._sort_coo_indices(idx)
    assert s.dtype == torch.int64 and p.dtype == torch.int64
    assert torch.equal(s.cpu(), torch.from_numpy(g("sorted"))) and torch.equal(p.cpu(), torch.from_numpy(g("perm")))
    crow, col, perm = U.convert_coo_to_csr_indices_values(idx, n)
    assert torch.equal(crow.cpu(), torch.from_numpy(g("crow")))
    assert torch.equal(col.cpu(), torch.from_numpy(g("col")))
    assert torch.equal(perm.cpu(), torch.from_numpy(g("csr_perm")))
    vals = torch.rand(idx.shape[1], device=DEV)
    _, _, v = U.convert_coo_to_csr_indices_values(idx, n, vals)
    assert torch.equal(v.flatten(), vals[perm.flatten()])
    if idx.shape[0] == 2:
        for it, tdt in (("i32", torch.int32), ("i64", torch.int64)):
            cr = U._compress_row_indices(s[0].to(tdt), n)
            assert cr.dtype == tdt and torch.equal(cr.cpu(), torch.from_numpy(g(f"compress_{it}")))
            rows = U._demcompress_crow_indices(cr, n)
            assert rows.dtype == tdt and torch.equal(rows.cpu(), torch.from_numpy(g(f"decompress_{it}")))
            # transpose: structure A.t().to_sparse_csr() gives, through the pattern machinery
            from torchsparsegradutils_b200._pattern import csr_pattern

            A = torch.sparse_csr_tensor(crow.to(tdt), col.to(tdt), torch.ones(col.numel(), device=DEV), (n, m))
            T = csr_pattern(A).transpose()
            assert torch.equal(T.rowptr.cpu().long(), torch.from_numpy(g(f"T_crow_{it}")).long())
            assert torch.equal(T.colind.cpu().long(), torch.from_numpy(g(f"T_col_{it}")).long())
            assert torch.equal(T.perm.cpu().long(), torch.from_numpy(g(f"T_perm_{it}")).long())


def test_probed_golden_a11():
    from torchsparsegradutils_b200.utils import utils as U

    crow, col, perm = U.convert_coo_to_csr_indices_values(_t(GIDX["a11/idx"]), 3)
    assert crow.tolist() == [0, 2, 3, 4] and col.tolist() == [0, 3, 0, 1] and perm.tolist() == [3, 1, 2, 0]


@pytest.mark.parametrize("shape,nnz", [((1000, 777), 50000), ((6, 300, 200), 9000), ((70000, 3), 100000)])
def test_sort_and_csr_vs_torch_and_oracle(shape, nnz):
    """tests/test_utils.py:53-117: sort == coalesce() indices, CSR == to_sparse_csr(), bit-exact."""
    from torchsparsegradutils_b200.utils import utils as U

    g = torch.Generator().manual_seed(7)
    b = shape[0] if len(shape) == 3 else 1
    n, m = shape[-2:]
    parts = []
    for t in range(b):
        flat = torch.randperm(n * m, generator=g)[:nnz]
        rc = torch.stack([flat // m, flat % m])
        parts.append(torch.cat([torch.full((1, nnz), t), rc]) if len(shape) == 3 else rc)
    idx = torch.cat(parts, dim=1)
    idx = idx[:, torch.randperm(idx.shape[1], generator=g)]
    vals = torch.rand(idx.shape[1], generator=g, dtype=torch.float64)
    s, p = U._sort_coo_indices(idx.to(DEV))
    co = torch.sparse_coo_tensor(idx, vals, shape).coalesce()
    assert torch.equal(s.cpu(), co.indices())
    assert torch.equal(vals[p.cpu()], co.values())
    so, po = orc.coo_sort(idx.numpy())
    assert np.array_equal(s.cpu().numpy(), so) and np.array_equal(p.cpu().numpy(), po)
    csr = U.convert_coo_to_csr(torch.sparse_coo_tensor(idx.to(DEV), vals.to(DEV), shape))
    assert csr.layout == torch.sparse_csr and csr.shape == torch.Size(shape)
    if len(shape) == 2:
        ref = co.to_sparse_csr()
        assert torch.equal(csr.crow_indices().cpu(), ref.crow_indices())
        assert torch.equal(csr.col_indices().cpu(), ref.col_indices())
        assert torch.equal(csr.values().cpu(), ref.values())
    else:
        for t in range(b):
            ref = co[t].coalesce().to_sparse_csr()
            assert torch.equal(csr.crow_indices()[t].cpu(), ref.crow_indices())
            assert torch.equal(csr.col_indices()[t].cpu(), ref.col_indices())
            assert torch.equal(csr.values()[t].cpu(), ref.values())


@pytest.mark.parametrize("idt", [torch.int32, torch.int64])
@pytest.mark.parametrize("batch", [None, 3])
def test_transpose_vs_torch_and_oracle(idt, batch):
    from helpers import rand_csr
    from torchsparsegradutils_b200._pattern import csr_pattern

    n, m = 513, 300
    A = rand_csr(n, m, 11, batch=batch, index_dtype=idt, seed=2, ragged=batch is None)
    T = csr_pattern(A).transpose()
    b = batch or 1
    crow = A.crow_indices().cpu().long().reshape(b, n + 1)
    col = A.col_indices().cpu().long().reshape(b, -1)
    off = 0
    for t in range(b):
        rT, cT, pT = orc.csr_transpose(crow[t].numpy(), col[t].numpy(), m)
        seg = slice(t * m, (t + 1) * m + 1)
        got_r = T.rowptr.cpu().long()[seg] - off
        assert np.array_equal(got_r.numpy(), rT)
        e0, e1 = off, off + int(rT[-1])
        assert np.array_equal(T.colind.cpu().long()[e0:e1].numpy(), cT)
        assert np.array_equal(T.perm.cpu().long()[e0:e1].numpy(), pT + t * col.shape[1])
        off = e1
    if batch is None:  # torch's own answer
        vals = torch.arange(1, col.numel() + 1, dtype=torch.float64)
        ref = torch.sparse_csr_tensor(crow[0], col[0], vals, (n, m)).t().to_sparse_csr()
        assert torch.equal(T.rowptr.cpu().long(), ref.crow_indices())
        assert torch.equal(T.colind.cpu().long(), ref.col_indices())
        assert torch.equal(T.perm.cpu().long(), ref.values().long() - 1)


def test_compress_unsorted_rows_and_empty():
    from torchsparsegradutils_b200.utils import utils as U

    rows = torch.tensor([4, 0, 4, 2, 0, 4], device=DEV)
    assert U._compress_row_indices(rows, 6).tolist() == [0, 2, 2, 3, 3, 6, 6]  # bincount semantics: any order
    assert U._compress_row_indices(torch.zeros(0, dtype=torch.int32, device=DEV), 3).tolist() == [0, 0, 0, 0]
    assert U._demcompress_crow_indices(torch.tensor([0, 0, 3, 3, 4], device=DEV), 4).tolist() == [1, 1, 1, 3]


@pytest.mark.parametrize("layout", ["coo", "csr"])
def test_block_diag_and_split_match_reference(layout):
    """tests/test_utils.py:137-156 and :209-230 against the reference's recorded outputs."""
    from torchsparsegradutils_b200.utils import utils as U

    pre = f"bd_{layout}/"
    shapes = [(4, 6), (3, 2), (5, 5)]
    mats = []
    for i, sh in enumerate(shapes):
        if layout == "coo":
            mats.append(torch.sparse_coo_tensor(_t(GIDX[pre + f"in{i}_indices"]), _t(GIDX[pre + f"in{i}_values"]), sh))
        else:
            mats.append(torch.sparse_csr_tensor(_t(GIDX[pre + f"in{i}_crow"]), _t(GIDX[pre + f"in{i}_col"]),
                                                _t(GIDX[pre + f"in{i}_values"]), sh))
    bd = U.sparse_block_diag(*mats)
    assert torch.equal(bd.to_dense().cpu(), torch.from_numpy(GIDX[pre + "dense"]))
    if layout == "coo":
        assert torch.equal(bd._indices().cpu(), torch.from_numpy(GIDX[pre + "indices"]))
    else:
        assert torch.equal(bd.crow_indices().cpu(), torch.from_numpy(GIDX[pre + "crow"]))
        assert torch.equal(bd.col_indices().cpu(), torch.from_numpy(GIDX[pre + "col"]))
    for i, part in enumerate(U.sparse_block_diag_split(bd, *shapes)):
        assert torch.equal(part.to_dense().cpu(), torch.from_numpy(GIDX[pre + f"split{i}_dense"]))


def test_stack_csr_matches_reference():
    from torchsparsegradutils_b200.utils import utils as U

    ts = [torch.sparse_csr_tensor(_t(GIDX[f"stack/in{i}_crow"]), _t(GIDX[f"stack/in{i}_col"]),
                                  _t(GIDX[f"stack/in{i}_values"]), (4, 5)) for i in range(3)]
    st = U.stack_csr(ts)
    assert st.crow_indices().dtype == torch.int32
    assert torch.equal(st.crow_indices().cpu(), torch.from_numpy(GIDX["stack/crow"]))
    assert torch.equal(st.col_indices().cpu(), torch.from_numpy(GIDX["stack/col"]))
    assert torch.equal(st.to_dense().cpu(), torch.from_numpy(GIDX["stack/dense"]))


def test_convert_preserves_int32_index_dtype_and_rejects_negatives():
    """Reference utils/utils.py:228-231, :324: crow / col keep the dtype of the incoming coordinates; the permutation
    is int64 (argsort).  Negative rows fail in the reference's _compress_row_indices check (:217)."""
    import pytest
    from torchsparsegradutils_b200.utils.utils import _sort_coo_indices, convert_coo_to_csr_indices_values

    idx = torch.tensor([[2, 0, 1, 0], [1, 3, 0, 0]], dtype=torch.int32, device="cuda:0")
    crow, col, perm = convert_coo_to_csr_indices_values(idx, 3)
    assert crow.dtype == torch.int32 and col.dtype == torch.int32 and perm.dtype == torch.int64
    assert crow.tolist() == [0, 2, 3, 4] and col.tolist() == [0, 3, 0, 1] and perm.tolist() == [3, 1, 2, 0]
    srt, p = _sort_coo_indices(idx)
    assert srt.dtype == torch.int32 and p.dtype == torch.int64
    # values of a dtype / shape the native gather does not cover go through index_select like the reference's values[perm]
    hv = torch.arange(8, device="cuda:0", dtype=torch.float16).reshape(4, 2)
    _, _, v = convert_coo_to_csr_indices_values(idx, 3, hv)
    assert torch.equal(v, hv[perm])
    bad = torch.tensor([[0, -1], [0, 1]], device="cuda:0")
    with pytest.raises(ValueError, match="negative"):
        convert_coo_to_csr_indices_values(bad, 3)
