#!/bin/bash
# 8-GPU (or $N) session: the driver's own SCALE command for the default config (strong scaling of BASELINE configs[1], e2e
# included), then the large single matrices row- and K-sharded.  Every run under its own timeout.
N=${N:-8}
O=gpurun_out/r2_multi_gpu
mkdir -p $O
TR="timeout ${RUN_TIMEOUT:-170} python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541"
$TR bench.py --gpus $N --steps 20 --warmup 5 > $O/scale_default_n$N.json 2> $O/scale_default_n$N.err
python - <<PY
import json
try:
    d = [json.loads(l) for l in open("$O/scale_default_n$N.json") if l.startswith("{")][-1]
    print("default N=$N |", d["scaling"], "| step ms", round(d["ms_per_step"], 4), "eager", round(d["eager_ms_per_step"], 4), "Gnnz/s", round(d["value"] / 1e9, 3), "e2e Gnnz/s", round(d["e2e"]["value"] / 1e9, 3), "host", d["host_enqueue_ms_per_step"], d["timed_region"][:20], [r["kernels_ms"] for r in d["per_rank"]][:2])
except Exception as ex:
    print("default failed:", ex); print(open("$O/scale_default_n$N.err").read()[-1500:])
PY
run() {
  name=$1; shift
  $TR bench.py --gpus $N "$@" --no-e2e --no-cpu-baseline --steps 20 > $O/${name}_n$N.json 2> $O/${name}_n$N.err
  python - <<PY
import json
try:
    d = [json.loads(l) for l in open("$O/${name}_n$N.json") if l.startswith("{")][-1]
    print("$name N=$N |", d["config"]["sharding"][:80], "| step ms", round(d["ms_per_step"], 3), "Gnnz/s", round(d["value"] / 1e9, 3), [sum(r["kernels_ms"].values()) for r in d["per_rank"]])
except Exception as ex:
    print("$name failed:", ex); print(open("$O/${name}_n$N.err").read()[-1200:])
PY
}
for spec in ${RUNS:-"rows_ar_cfg4:4:rows" "rows_ar_cfg5:5:rows" "k_cfg5:5:k"}; do
  IFS=: read name cfg sh <<< "$spec"
  run $name --config $cfg --sharding $sh
done 2>&1 | tee $O/sharding_n$N.txt
