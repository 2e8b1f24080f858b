#!/bin/bash
# short device-resident bench of every BASELINE config (no e2e / CPU legs); prints one summary line each
for c in ${CONFIGS:-1 2 3 5 5bf16 4}; do
  echo "== config $c"
  python bench.py --config $c --no-e2e --no-cpu-baseline --steps ${STEPS:-10} 2>gpurun_out/bench_all_last.err | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(round(d['ms_per_step'],4), 'ms  Gnnz/s', round(d['value']/1e9,3), 'nnz', d['nnz_per_step'], 'step_frac', round(d['step_frac_of_hbm_peak'],3), {k:(round(v['ms'],4), round(v.get('frac',0),3), round(v.get('gather_gbs',0))) for k,v in d['kernels'].items()}, d['clocks'])"
done
