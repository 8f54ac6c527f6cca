#!/bin/bash
# Round 2, GPU call 8 (1 GPU): matcher with the 8-warp streaming epilogue.
set -u
OUT=gpurun_out/r2c9
mkdir -p "$OUT"
run() {  # run <seconds> <logname> <command...>
  local t=$1 name=$2; shift 2
  echo "=== $name: $*" | tee -a "$OUT/summary.txt"
  local t0=$(date +%s)
  timeout "$t" "$@" > "$OUT/$name.log" 2>&1
  echo "exit $? after $(( $(date +%s) - t0 )) s" | tee -a "$OUT/summary.txt"
  tail -6 "$OUT/$name.log" | cut -c1-600 >> "$OUT/summary.txt"
}
run 150 matcher_tests python -m pytest tests/test_xx_matcher_gpu.py -m gpu -q -x
run 200 bench_c5 python bench.py --workload c5_matcher --steps 3 --warmup 1
run 200 ncu_matcher ncu --set full --clock-control none --import-source on -k regex:k_nn_candidates -c 1 -o "$OUT/r2_matcher" -f \
    python bench.py --workload c5_matcher --steps 1 --warmup 0 --no-cpu-baseline
cat "$OUT/summary.txt" | cut -c1-300
