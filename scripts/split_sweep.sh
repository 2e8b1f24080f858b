#!/bin/bash
# config 4 (R-MAT): merge-path vs split-row mode at several piece bounds
echo "== merge"; TSGU_B200_ALGO=merge CONFIGS=4 STEPS=5 bash scripts/bench_all_configs.sh | tail -1
for b in 64 128 256 512; do
  echo "== split bound $b"; TSGU_B200_ALGO=split TSGU_B200_SPLIT_BOUND=$b CONFIGS=4 STEPS=5 bash scripts/bench_all_configs.sh | tail -1
done
