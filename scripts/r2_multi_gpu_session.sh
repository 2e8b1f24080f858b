#!/bin/bash
# One multi-GPU gpurun session (N = $N GPUs): NCCL tests, collective probe, row / K sharding of configs 4 and 5 with the
# collective overlapped or not and all-reduce vs reduce-scatter, and the strong-scaling run of config 2.
N=${N:-2}
mkdir -p gpurun_out/r2_multi_gpu
O=gpurun_out/r2_multi_gpu
[ -n "$SKIP_TESTS" ] || timeout 200 python -m pytest tests/test_distributed_gpu.py -x -q 2>&1 | tail -2 | tee $O/dist_tests_n$N.txt
TR="timeout ${RUN_TIMEOUT:-150} python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533"
[ -n "$SKIP_PROBE" ] || $TR scripts/nccl_collective_probe.py 2>/dev/null | tee $O/nccl_probe_n$N.txt
run() {  # name, bench args...
  name=$1; shift
  $TR bench.py --gpus $N "$@" --no-e2e --no-cpu-baseline --steps 20 > $O/${name}_n$N.json 2> $O/${name}_n$N.err
  python - <<PY
import json
try:
    d = [json.loads(l) for l in open("$O/${name}_n$N.json") if l.startswith("{")][-1]
    print("$name N=$N |", d["config"]["sharding"][:90], "| step ms", round(d["ms_per_step"], 3), "eager", round(d["eager_ms_per_step"], 3), "Gnnz/s", round(d["value"] / 1e9, 3), {k: round(v["ms"], 3) for k, v in d["kernels"].items()}, "host", round(d["host_enqueue_ms_per_step"], 3))
except Exception as ex:
    print("$name failed:", ex); print(open("$O/${name}_n$N.err").read()[-1200:])
PY
}
for c in ${CONFIGS:-5 4}; do
  run rows_ar_cfg$c --config $c --sharding rows
  [ -n "$QUICK" ] || run rows_ar_noov_cfg$c --config $c --sharding rows --no-overlap
  run rows_rs_cfg$c --config $c --sharding rows --grad-b reduce_scatter
  run k_cfg$c --config $c --sharding k
done 2>&1 | tee $O/sharding_n$N.txt
run strong_cfg2 --config 2 2>&1 | tee $O/strong_cfg2_n$N.txt
