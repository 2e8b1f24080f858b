"""Rebuild profiles/r<N>_all_configs.md from the raw JSON lines in profiles/r<N>_bench_full/ (scripts/bench_full_record.sh).

    python profiles/make_all_configs_md.py [r1|r2]
"""
import json, os, sys
RND = sys.argv[1] if len(sys.argv) > 1 else "r2"
HERE = os.path.dirname(os.path.abspath(__file__))
rows = []
for c in ["1", "2", "3", "4", "5", "5bf16"]:
    d = json.load(open(os.path.join(HERE, f"{RND}_bench_full", f"bench_full_cfg{c}.json")))
    k, e, cb = d["kernels"], d.get("e2e") or {}, d.get("cpu_baseline") or {}
    ks = " / ".join("%.3f" % k[n]["ms"] for n in ("spmm_fwd", "sddmm", "spmm_gradB"))
    rows.append("| %s | %s | %s | %.3f | %.2f | %.0f | %.3f | %s | %.3f (%.1f ms) | %.2f (%d thr; %s) | %.0fx / %.0fx |" % (
        c, d["config"]["workload"], f"{d['nnz_per_step']:,}", d["ms_per_step"], d["value"] / 1e9, d["gflops"],
        d["step_frac_of_hbm_peak"], ks, e["value"] / 1e9, e["ms_per_step"], cb["value"] / 1e6, cb["cores"], cb["sample"],
        d["value"] / cb["value"], e["value"] / cb["value"]))
hdr = f"""# Round {RND[1:]}: every BASELINE config on one B200 (`scripts/bench_full_record.sh`, raw JSON lines in `{RND}_bench_full/`)

Device-resident step = forward + backward of `sparse_mm` through the public API with cached patterns (30 steps,
CUDA events; round 2: replayed from a CUDA graph, per-kernel times from the eager region beside it); e2e = the same from pinned host buffers (H2D of A, B, G and D2H of C, grad_A, grad_B inside the
timed region, cold pattern every step); CPU = the reference's data flow (`oracle/reference_port.py`, torch CPU ops,
all host threads) on the stated sample.  `step frac` = algorithmic bytes of the whole step / time / 6548.5 GB/s.
(Regenerate with `python profiles/make_all_configs_md.py`.)

| cfg | workload | nnz/step | step ms | Gnnz/s | GFLOP/s | step frac | fwd / SDDMM / grad_B ms | e2e Gnnz/s | CPU Mnnz/s | GPU/CPU (resident / e2e) |
|---|---|---|---|---|---|---|---|---|---|---|
"""
open(os.path.join(HERE, f"{RND}_all_configs.md"), "w").write(hdr + "\n".join(rows) + "\n")
print("\n".join(rows))
