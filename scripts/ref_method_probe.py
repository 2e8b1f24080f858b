"""Reference-methodology wall clock (clone A, clone B, op, sum().backward(), sync; empty_cache before every repeat) for one
comparability config, with a cProfile of one fwd+bwd repeat."""
import cProfile, pstats, sys, os, time, io
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from torchsparsegradutils_b200 import sparse_mm
cfg = bench.CONFIGS[sys.argv[1]]
A, B, G = bench.build_inputs(cfg, torch.device("cuda:0"))
print(sys.argv[1], bench.reference_methodology_ms(A, B, sparse_mm, repeats=5))
def one():
    torch.cuda.empty_cache(); torch.cuda.synchronize()
    A1 = A.detach().clone().requires_grad_(True); B1 = B.detach().clone().requires_grad_(True)
    out = sparse_mm(A1, B1); out.sum().backward(); torch.cuda.synchronize()
one()
pr = cProfile.Profile(); pr.enable(); one(); pr.disable()
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(22); print(s.getvalue()[:3500])
