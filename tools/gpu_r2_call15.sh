#!/bin/bash
# Round 2, GPU call 15 (1 GPU): the two variants of the matcher's exact pass (dense staged work list vs per-lane loads, both with the
# register-staged double-buffered exhaustive scan): tests, racecheck, bench.
set -u
OUT=gpurun_out/r2c15
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
TBM_EXACT=lanes run 300 matcher_tests_lanes python -m pytest tests/test_xx_matcher_gpu.py -q -m gpu -x
run 200 bench_c5_staged python bench.py --workload c5_matcher --steps 3 --warmup 1 --no-cpu-baseline
TBM_EXACT=lanes run 200 bench_c5_lanes python bench.py --workload c5_matcher --steps 3 --warmup 1 --no-cpu-baseline
run 300 sanitizer_racecheck_matcher compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_xx_matcher_gpu.py -q -m gpu -x -k "exact_pass or random_descriptors"
TBM_EXACT=lanes run 300 sanitizer_racecheck_matcher_lanes compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_xx_matcher_gpu.py -q -m gpu -x -k "exact_pass"
run 200 ncu_exact ncu --set full --clock-control none -k regex:k_exact_top2 -c 1 -o "$OUT/r2_exact" -f python bench.py --workload c5_matcher --steps 1 --warmup 0 --no-cpu-baseline
python profiles/summarize.py full "$OUT/r2_exact.ncu-rep" > "$OUT/ncu_exact_summary.txt" 2>&1
rm -f "$OUT/r2_exact.ncu-rep"
find gpurun_out -size +8M -delete
cat "$OUT/summary.txt" | cut -c1-300
