#!/bin/bash
# Round 2, GPU call 4 (1 GPU): remaining test files with the de-parallelised small-problem oracle, bench with the fused prepare kernel.
set -u
OUT=gpurun_out/r2c4
mkdir -p "$OUT"
run() {  # run <seconds> <logname> <command...>
  local t=$1 name=$2; shift 2
  echo "=== $name: $*" | tee -a "$OUT/summary.txt"
  local t0=$(date +%s)
  timeout "$t" "$@" > "$OUT/$name.log" 2>&1
  echo "exit $? after $(( $(date +%s) - t0 )) s" | tee -a "$OUT/summary.txt"
  tail -4 "$OUT/$name.log" | cut -c1-400 >> "$OUT/summary.txt"
}
run 200 parity python -m pytest tests/test_gpu_parity.py -m gpu -q -x
run 300 bench_c3 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-experiments
for f in test_xx_inner_iterations_gpu test_xx_track_estimator_gpu test_xx_two_view_gpu test_z_adapter_gpu test_zz_experiments_gpu test_x_exact_schur_gpu; do
  run 150 "$f" python -m pytest "tests/$f.py" -q -m gpu
done
run 200 bench_c2 python bench.py --workload c2_1kcam --steps 10 --warmup 3 --no-cpu-baseline --no-experiments
run 200 bench_c4 python bench.py --workload c4_radtan --steps 10 --warmup 3 --no-cpu-baseline --no-experiments
run 300 ncu_lin ncu --set full --clock-control none --import-source on -k "regex:k_linearize|k_cost|k_prepare_stream|k_schur_stream" -c 7 \
    -o "$OUT/r2_lin" -f python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-e2e --no-experiments
cat "$OUT/summary.txt" | cut -c1-300
