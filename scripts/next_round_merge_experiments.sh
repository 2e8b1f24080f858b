#!/bin/bash
# Round-2 starting point for config 4: the two gated merge-path variants written but not
# measured in round 1.  Step 1 (CPU box, ~4 min): build the variant libraries.  Step 2 (ONE gpurun call): parity of
# every variant on the merge-path tests, then the config-4 sweep.
#
#   bash scripts/next_round_merge_experiments.sh build
#   gpurun --timeout 600 -- 'bash scripts/next_round_merge_experiments.sh run'
set -e
cd "$(dirname "$0")/.."
case "$1" in
  build)
    python -m torchsparsegradutils_b200.csrc.build
    mkdir -p variants_tmp && cp torchsparsegradutils_b200/libtsgu_b200.so variants_tmp/lib_cur.so
    scripts/build_variant.sh PF    merge -DTSGU_MERGE_SDDMM_PREFETCH=1 &
    scripts/build_variant.sh PF2   merge -DTSGU_MERGE_SDDMM_PREFETCH=1 -DTSGU_MERGE_SDDMM_MINB=2 &
    scripts/build_variant.sh HR3K  merge -DTSGU_MERGE_HEAD_IN_REGS=1 &
    scripts/build_variant.sh HR4K  merge -DTSGU_MERGE_HEAD_IN_REGS=1 -DTSGU_MERGE_P=4096 &
    wait ;;
  run)
    mkdir -p gpurun_out
    for lib in variants_tmp/lib_*.so; do
      echo "== parity $lib"
      TSGU_B200_LIB=$PWD/$lib python -m pytest tests/test_sparse_mm_gpu.py -q -x -k merge 2>&1 | tail -1
    done | tee gpurun_out/merge_variants_parity.txt
    bash scripts/merge_variant_sweep.sh 2>&1 | tee gpurun_out/merge_variants_sweep.txt
    # host-side experiment: value gather of the transposed pass on a side stream, overlapped with the SDDMM
    echo "== overlap gather: parity"; TSGU_B200_OVERLAP_GATHER=1 python -m pytest tests/test_sparse_mm_gpu.py tests/test_golden_gpu.py -q -x 2>&1 | tail -1
    for v in 0 1; do echo "== TSGU_B200_OVERLAP_GATHER=$v"; TSGU_B200_OVERLAP_GATHER=$v CONFIGS="2 3" bash scripts/bench_all_configs.sh; done 2>&1 | tee gpurun_out/overlap_gather.txt ;;
  *) echo "usage: $0 build|run"; exit 2 ;;
esac
