"""Where the cold-pattern time goes: csr_pattern / transpose / window plan / sort / pad, each synchronised, twice
(the second round shows what is left once the caching allocator holds blocks of the right sizes)."""
import os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from torchsparsegradutils_b200 import _pattern as P, sparse_mm

cfg_id = sys.argv[1] if len(sys.argv) > 1 else "2"
dev = torch.device("cuda:0")
A, B, G = bench.build_inputs(bench.CONFIGS[cfg_id], dev)
A.requires_grad_(True); B.requires_grad_(True)
def step():
    A.grad = None; B.grad = None
    C = sparse_mm(A, B); C.backward(G)
for _ in range(3): step()
torch.cuda.synchronize()
def T(f):
    torch.cuda.synchronize(); t = time.perf_counter(); r = f(); torch.cuda.synchronize(); return r, (time.perf_counter() - t) * 1e3
for rnd in range(3):
    P.clear_pattern_cache()
    _, t_all = T(step)
    P.clear_pattern_cache()
    pat, t1 = T(lambda: P.csr_pattern(A.detach()))
    _, t2 = T(lambda: P.window_plan(pat))
    # transpose pieces
    orig_sort, orig_pad, orig_wp = P._sort_rows_by_length, P._pad_rows, P.window_plan
    times = {}
    def wrap(name, f):
        def g(*a, **k):
            r, t = T(lambda: f(*a, **k)); times[name] = times.get(name, 0) + t; return r
        return g
    P._sort_rows_by_length, P._pad_rows, P.window_plan = wrap("sort_rows", orig_sort), wrap("pad_rows", orig_pad), wrap("window_plan_T", orig_wp)
    _, t3 = T(lambda: pat.transpose())
    P._sort_rows_by_length, P._pad_rows, P.window_plan = orig_sort, orig_pad, orig_wp
    print(f"round {rnd}: cold step {t_all:.2f} ms | csr_pattern {t1:.2f} | window_plan(fwd) {t2:.2f} | transpose total {t3:.2f} of which {{{', '.join(f'{k} {v:.2f}' for k, v in times.items())}}}")
