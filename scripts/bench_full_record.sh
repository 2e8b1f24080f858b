#!/bin/bash
# full bench line (device-resident value, e2e from host, CPU baseline) for every BASELINE config
mkdir -p gpurun_out
for c in ${CONFIGS:-1 2 3 4 5 5bf16}; do
  timeout 900 python bench.py --config $c --steps ${STEPS:-30} > gpurun_out/bench_full_cfg$c.json 2> gpurun_out/bench_full_cfg$c.err
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/bench_full_cfg$c.json"))
    e = d.get("e2e") or {}
    cb = d.get("cpu_baseline") or {}
    print("cfg $c: step %.3f ms  %.3f Gnnz/s  %.0f GFLOP/s  step_frac %.3f | e2e %.3f Gnnz/s (%.1f ms) | cpu %.2f Mnnz/s x%d cores (%s) | cold %.0f ms" % (
        d["ms_per_step"], d["value"] / 1e9, d["gflops"], d["step_frac_of_hbm_peak"], (e.get("value") or 0) / 1e9, e.get("ms_per_step") or 0,
        (cb.get("value") or 0) / 1e6, cb.get("cores") or 0, cb.get("sample"), d.get("cold_first_step_ms") or 0))
except Exception as ex:
    print("cfg $c failed:", ex)
PY
done
