"""Golden sparsity patterns of BASELINE config 3 from the REAL reference encoder (build container only).

    PYTHONPATH=/root/reference python tests/golden/make_golden_stencil.py

``PairwiseEncoder(radius=1.8, volume_shape=(1, D, D, D), diag=True, upper=None | False, layout=csr)``
(reference ``encoders/pairwise_encoder.py:383-505`` pattern, ``:665-712`` CSR conversion) at D = 8 and 16.
``workloads.stencil27_csr`` -- the arithmetic generator bench.py and the full-size tests use at D = 128, where
the reference's generator is too slow -- must reproduce these crow/col arrays bit for bit
(``tests/test_stencil_pattern.py``).  Index arrays are stored as int32 to keep the fixture small.
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, "/root/reference")
from torchsparsegradutils.encoders import PairwiseEncoder  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    store = {}
    for D in (8, 16):
        for tag, upper in (("full", None), ("lower", False)):
            enc = PairwiseEncoder(radius=1.8, volume_shape=(1, D, D, D), diag=True, upper=upper,
                                  layout=torch.sparse_csr, indices_dtype=torch.int64)
            crow, col = enc.crow_indices, enc.col_indices
            store[f"D{D}_{tag}/crow"] = crow.numpy().astype(np.int32)
            store[f"D{D}_{tag}/col"] = col.numpy().astype(np.int32)
            store[f"D{D}_{tag}/num_offsets"] = np.array(len(enc.offsets))
            print(D, tag, "offsets", len(enc.offsets), "nnz", col.numel())
    np.savez_compressed(os.path.join(HERE, "stencil_patterns.npz"), **store)


if __name__ == "__main__":
    main()
