#!/bin/bash
# config 4 (R-MAT, merge-path kernels): device-resident bench for every kernel-variant library under variants_tmp/
# (built with scripts/build_variant.sh; TSGU_B200_LIB selects the library).
for lib in ${LIBS:-$(ls variants_tmp/lib_*.so)}; do
  echo "== $lib"
  TSGU_B200_LIB=$PWD/$lib CONFIGS=${CONFIGS:-4} STEPS=${STEPS:-10} bash scripts/bench_all_configs.sh | grep -v "^=="
done
