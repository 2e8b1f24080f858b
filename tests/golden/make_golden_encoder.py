"""Golden fixtures for the PairwiseEncoder value assembly, from the REAL reference encoder
(build container only):

    PYTHONPATH=/root/reference python tests/golden/make_golden_encoder.py

For every case the reference ``PairwiseEncoder`` (encoders/pairwise_encoder.py:562-849) is built, called on seeded
value volumes (unbatched and batched), and a weighted sum of the produced sparse values is back-propagated to the
input volumes.  Stored: the encoder's own pattern arrays (what PairwiseValueAssembler.from_encoder consumes), the
inputs, the produced index / value arrays, and the input gradients.
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, "/root/reference")
from torchsparsegradutils.encoders import PairwiseEncoder  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
torch.manual_seed(5)

CASES = [
    # name, radius, volume_shape, diag, upper, relation, layout, index dtype
    ("csr3d_inter", 1.5, (2, 3, 4, 3), True, None, "inter", torch.sparse_csr, torch.int64),
    ("coo2d_lower", 1.0, (1, 4, 5), False, False, "indep", torch.sparse_coo, torch.int64),
    ("csr3d_tril_i32", 1.8, (1, 5, 4, 3), True, False, "indep", torch.sparse_csr, torch.int32),
    ("coo3d_intra", 1.0, (2, 3, 3, 2), True, True, "intra", torch.sparse_coo, torch.int64),
]


def npy(t):
    return t.detach().cpu().numpy()


def main():
    store = {}
    names = []
    for name, radius, shape, diag, upper, rel, layout, idt in CASES:
        enc = PairwiseEncoder(radius=radius, volume_shape=shape, diag=diag, upper=upper, channel_voxel_relation=rel,
                              layout=layout, indices_dtype=idt)
        p = name + "/"
        store[p + "offsets"] = np.array(enc.offsets, dtype=np.int64)
        store[p + "volume_shape"] = np.array(shape, dtype=np.int64)
        store[p + "layout"] = np.array("csr" if layout == torch.sparse_csr else "coo")
        if layout == torch.sparse_csr:
            store[p + "crow"], store[p + "col"], store[p + "perm"] = npy(enc.crow_indices), npy(enc.col_indices), npy(enc.csr_permutation)
        else:
            store[p + "indices"] = npy(enc.indices)
        for tag, lead in (("u", ()), ("b", (3,))):
            for dt, dname in ((torch.float32, "f32"), (torch.float64, "f64")):
                vals = torch.randn(*lead, len(enc.offsets), *shape, dtype=dt, requires_grad=True)
                A = enc(vals)
                out_vals = A.values()
                w = torch.randn(out_vals.shape, dtype=dt)
                (out_vals * w).sum().backward()
                q = f"{p}{tag}_{dname}/"
                store[q + "values_in"], store[q + "w"], store[q + "grad_in"] = npy(vals), npy(w), npy(vals.grad)
                store[q + "values_out"] = npy(out_vals)
                if layout == torch.sparse_csr:
                    store[q + "crow_out"], store[q + "col_out"] = npy(A.crow_indices()), npy(A.col_indices())
                else:
                    store[q + "indices_out"] = npy(A.indices())
        names.append(name)
        print(name, "offsets", len(enc.offsets), "nnz", int(out_vals.numel()))
    store["__cases__"] = np.array(names)
    np.savez_compressed(os.path.join(HERE, "encoder_cases.npz"), **store)


if __name__ == "__main__":
    main()
