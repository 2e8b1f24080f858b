"""cProfile of prepare_pattern on one item of config 2 (what the e2e leg pays per item)."""
import cProfile, pstats, io, os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from torchsparsegradutils_b200 import clear_pattern_cache, prepare_pattern
A, B, G = bench.build_inputs(bench.CONFIGS["2shard8"], torch.device("cuda:0"))
A = torch.sparse_csr_tensor(A.crow_indices()[0], A.col_indices()[0], A.values()[0], tuple(A.shape[1:]))
for _ in range(3):
    clear_pattern_cache(); prepare_pattern(A)
torch.cuda.synchronize()
ts = []
for _ in range(10):
    clear_pattern_cache(); torch.cuda.synchronize(); t = time.perf_counter(); prepare_pattern(A); torch.cuda.synchronize(); ts.append((time.perf_counter() - t) * 1e3)
print("prepare_pattern ms:", [round(x, 2) for x in ts])
clear_pattern_cache()
pr = cProfile.Profile(); pr.enable(); prepare_pattern(A); torch.cuda.synchronize(); pr.disable()
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(14); print(s.getvalue()[:2600])
