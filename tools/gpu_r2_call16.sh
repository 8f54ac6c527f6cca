#!/bin/bash
# Round 2, GPU call 16 (1 GPU): the four variants of the matcher's exact pass (listed candidates staged / per-lane  x  exhaustive scan
# pipelined / plain), tests of the default.
set -u
OUT=gpurun_out/r2c16
mkdir -p "$OUT"
run() {  # run <seconds> <logname> <command...>
  local t=$1 name=$2; shift 2
  echo "=== $name: $*" | tee -a "$OUT/summary.txt"
  local t0=$(date +%s)
  timeout "$t" "$@" > "$OUT/$name.log" 2>&1
  echo "exit $? after $(( $(date +%s) - t0 )) s" | tee -a "$OUT/summary.txt"
  tail -3 "$OUT/$name.log" | cut -c1-400 >> "$OUT/summary.txt"
}
run 200 matcher_tests python -m pytest tests/test_xx_matcher_gpu.py -q -m gpu -x
TBM_EXACT=lanes TBM_EXH=simple run 200 matcher_tests_lanes_simple python -m pytest tests/test_xx_matcher_gpu.py -q -m gpu -x
run 100 bench_staged_pipe python bench.py --workload c5_matcher --steps 3 --warmup 1 --no-cpu-baseline
TBM_EXH=simple run 100 bench_staged_simple python bench.py --workload c5_matcher --steps 3 --warmup 1 --no-cpu-baseline
TBM_EXACT=lanes run 100 bench_lanes_pipe python bench.py --workload c5_matcher --steps 3 --warmup 1 --no-cpu-baseline
TBM_EXACT=lanes TBM_EXH=simple run 100 bench_lanes_simple python bench.py --workload c5_matcher --steps 3 --warmup 1 --no-cpu-baseline
TBM_EXACT=lanes TBM_EXH=simple run 200 racecheck_lanes_simple compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_xx_matcher_gpu.py -q -m gpu -x -k "exact_pass"
cat "$OUT/summary.txt" | cut -c1-300
