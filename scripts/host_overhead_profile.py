"""cProfile of the eager host path (Python + ctypes + autograd) of one small sparse_mm fwd+bwd step on the GPU box.
    python scripts/host_overhead_profile.py [config]      -> top cumulative entries; kernels are tiny, the host dominates."""
import cProfile
import os
import pstats
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from torchsparsegradutils_b200 import sparse_mm  # noqa: E402

cfg = bench.CONFIGS[sys.argv[1] if len(sys.argv) > 1 else "batched128"]
dev = torch.device("cuda:0")
A, B, G = bench.build_inputs(cfg, dev)
A = A.requires_grad_(True)
B = B.requires_grad_(True)


def step():
    A.grad = None
    B.grad = None
    C = sparse_mm(A, B)
    C.backward(G)


for _ in range(20):
    step()
torch.cuda.synchronize()
N = 300
t0 = time.perf_counter()
for _ in range(N):
    step()
t1 = time.perf_counter()
torch.cuda.synchronize()
print("host enqueue per step: %.3f ms (wall incl. drain %.3f ms)" % ((t1 - t0) * 1e3 / N, (time.perf_counter() - t0) * 1e3 / N))
pr = cProfile.Profile()
pr.enable()
for _ in range(N):
    step()
pr.disable()
torch.cuda.synchronize()
st = pstats.Stats(pr)
st.sort_stats("cumulative").print_stats(45)
