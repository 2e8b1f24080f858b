#!/bin/bash
# N-GPU comparison of the two shardings of one large matrix (rows: north_star's scheme; k: dense columns), configs 5 and 4
N=${N:-2}
mkdir -p gpurun_out
for sh in ${SHARDINGS:-rows k}; do for c in ${CONFIGS:-5 4}; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $N --config $c --sharding $sh --no-e2e --no-cpu-baseline --steps 20 > gpurun_out/shard_${sh}_cfg${c}_n$N.json 2> gpurun_out/shard_${sh}_cfg${c}_n$N.err
  python - <<PY
import json
try:
    d = [json.loads(l) for l in open("gpurun_out/shard_${sh}_cfg${c}_n$N.json") if l.startswith("{")][-1]  # NCCL prints a banner first
    print("cfg", d["config"]["config_id"], "N=$N", d["config"]["sharding"][:70], "| step ms", round(d["ms_per_step"], 3), "Gnnz/s", round(d["value"] / 1e9, 3), {k: round(v["ms"], 3) for k, v in d["kernels"].items()})
except Exception as ex:
    print("cfg $c sharding $sh failed:", ex)
    print(open("gpurun_out/shard_${sh}_cfg${c}_n$N.err").read()[-1500:])
PY
done; done
