#!/bin/bash
# Round-2 evidence on ONE GPU: ncu launch list of the default bench command, ncu --set full of one step of configs 2 and 3,
# compute-sanitizer over the window / encoder / plumbing tests.
mkdir -p gpurun_out/r2_prof
O=gpurun_out/r2_prof
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file $O/r2_cfg2_launches.csv \
    python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline > $O/launch_bench.log 2>&1
for c in 2 3; do
  ncu --set full --clock-control none -k regex:"tile_kernel|window_kernel|gather_values" -s 12 -c 4 \
      -o $O/r2_cfg${c}_full python bench.py --config $c --no-e2e --no-cpu-baseline --steps 2 --warmup 3 --no-graph > $O/ncu_cfg$c.log 2>&1
  ncu -i $O/r2_cfg${c}_full.ncu-rep --page raw --csv > $O/r2_cfg${c}_full_raw.csv 2>/dev/null
  rm -f $O/r2_cfg${c}_full.ncu-rep   # the report with imported sources is > 64 MiB: keep the CSV pages only
done
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_window_gpu.py tests/test_encoder_gpu.py tests/test_solve_batched_gpu.py -x -q -k "not plan_reconstructs" > $O/sanitizer_memcheck.txt 2>&1; echo "memcheck rc=$?" >> $O/sanitizer_memcheck.txt
timeout 600 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_window_gpu.py -x -q -k "stencil_window_kernels_vs_oracle or banded" > $O/sanitizer_racecheck.txt 2>&1; echo "racecheck rc=$?" >> $O/sanitizer_racecheck.txt
tail -n 3 $O/sanitizer_memcheck.txt; tail -n 3 $O/sanitizer_racecheck.txt
rm -f $O/*.ncu-rep.tmp
ls -la $O
