#!/bin/bash
# Round 2, GPU call 6 (1 GPU): matvec variants (J-row caching, mbarrier wait form), matcher epilogue v2.
set -u
OUT=gpurun_out/r2c6
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
run 150 bench_c3_A python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-experiments --no-e2e
THEIA_BA_B200_LIB=$PWD/theiasfm_b200/libtheia_ba_b200_varB.so run 150 bench_c3_B python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-experiments --no-e2e
THEIA_BA_B200_LIB=$PWD/theiasfm_b200/libtheia_ba_b200_varC.so run 150 bench_c3_C python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-experiments --no-e2e
run 200 ncu_matcher ncu --set full --clock-control none --import-source on -k regex:k_nn_candidates -c 1 -o "$OUT/r2_matcher" -f \
    python bench.py --workload c5_matcher --steps 1 --warmup 0 --no-cpu-baseline
grep -o '"ms_per_step": [0-9.]*\|"avg_launch_ms": [0-9.]*\|"frac": [0-9.]*' "$OUT"/bench_c3_*.log | head -20
cat "$OUT/summary.txt" | cut -c1-300
