#!/bin/bash
# Round 2, GPU call 2 (1 GPU): the persistent streaming Schur kernels.  gpurun --timeout 1800 -- 'bash tools/gpu_r2_call2.sh'
set -u
OUT=gpurun_out/r2c2
mkdir -p "$OUT"
run() {  # run <seconds> <logname> <command...>
  local t=$1 name=$2; shift 2
  echo "=== $name: $*" | tee -a "$OUT/summary.txt"
  local t0=$(date +%s)
  timeout "$t" "$@" > "$OUT/$name.log" 2>&1
  echo "exit $? after $(( $(date +%s) - t0 )) s" | tee -a "$OUT/summary.txt"
  tail -4 "$OUT/$name.log" >> "$OUT/summary.txt"
}
run 600 parity python -m pytest tests/test_gpu_parity.py -m gpu -q -x
run 900 bench_c3 python bench.py --steps 10 --warmup 3 --no-cpu-baseline
run 1200 pytest_rest python -m pytest tests -m gpu -q --durations=10 --deselect tests/test_gpu_parity.py
run 400 bench_c2 python bench.py --workload c2_1kcam --steps 10 --warmup 3 --no-cpu-baseline --no-experiments
run 400 bench_c4 python bench.py --workload c4_radtan --steps 10 --warmup 3 --no-cpu-baseline --no-experiments
run 600 ncu_full_schur ncu --set full --clock-control none --import-source on -k regex:k_schur -c 4 -o "$OUT/r2_stream" -f \
    python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-e2e --no-experiments
SUB='test_residuals_match_golden or (test_stage_parity and pinhole_shared and not True-) or (test_full_solve_parity and radtan_per_camera and 1)'
run 500 sanitizer_memcheck compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "$SUB"
run 500 sanitizer_racecheck compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "$SUB"
cat "$OUT/summary.txt" | cut -c1-600
