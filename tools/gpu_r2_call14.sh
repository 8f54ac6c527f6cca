#!/bin/bash
# Round 2, GPU call 14 (1 GPU): final single-GPU validation -- whole GPU suite, racecheck / memcheck of the matcher's new kernels, the
# bench lines of every single-GPU config (c3 with microbenchmarks, CPU baseline and experiments; reference arm; c2, c4, c5).
set -u
OUT=gpurun_out/r2c14
mkdir -p "$OUT"
run() {  # run <seconds> <logname> <command...>
  local t=$1 name=$2; shift 2
  echo "=== $name: $*" | tee -a "$OUT/summary.txt"
  local t0=$(date +%s)
  timeout "$t" "$@" > "$OUT/$name.log" 2>&1
  echo "exit $? after $(( $(date +%s) - t0 )) s" | tee -a "$OUT/summary.txt"
  tail -4 "$OUT/$name.log" | cut -c1-600 >> "$OUT/summary.txt"
}
run 300 matcher_tests python -m pytest tests/test_xx_matcher_gpu.py -q -m gpu -x
run 200 bench_c5 python bench.py --workload c5_matcher --steps 3 --warmup 1
run 300 sanitizer_racecheck_matcher compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_xx_matcher_gpu.py -q -m gpu -x -k "exact_pass or reference_cases or random_descriptors"
run 300 sanitizer_memcheck_matcher compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_xx_matcher_gpu.py -q -m gpu -x -k "exact_pass or reference_cases or random_descriptors"
run 600 pytest_gpu python -m pytest tests -m gpu -q -x --durations=5
run 900 bench_c3 python bench.py
run 400 bench_ref python bench.py --impl reference
run 200 bench_c2 python bench.py --workload c2_1kcam --steps 10 --warmup 3 --no-cpu-baseline --no-experiments
run 200 bench_c4 python bench.py --workload c4_radtan --steps 10 --warmup 3 --no-cpu-baseline --no-experiments
SUB='test_residuals_match_golden or (test_stage_parity and pinhole_shared and not True-) or (test_full_solve_parity and radtan_per_camera and 1) or not_positive'
run 400 sanitizer_racecheck_ba compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "$SUB"
find gpurun_out -size +8M -delete
cat "$OUT/summary.txt" | cut -c1-300
