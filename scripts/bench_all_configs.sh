for c in 1 3 5 5bf16 4; do
  echo "== config $c"
  python bench.py --config $c --no-e2e --no-cpu-baseline --steps 10 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(round(d['ms_per_step'],4), 'Gnnz/s', round(d['value']/1e9,3), 'nnz', d['nnz_per_step'], 'step_frac', round(d['step_frac_of_hbm_peak'],3), {k:(round(v['ms'],4), round(v['frac'],3), round(v['gather_gbs'])) for k,v in d['kernels'].items()})"
done
