#!/bin/bash
# config 5 (B does not fit L2) with different K-slice budgets; 0 = slicing off
for f in 0 0.3 0.55 0.8; do
  echo "== TSGU_L2_SLICE_FRAC=$f"
  TSGU_L2_SLICE_FRAC=$f CONFIGS="${CONFIGS:-5 5bf16}" bash scripts/bench_all_configs.sh
done
