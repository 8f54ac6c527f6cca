#!/bin/bash
# Round 2, GPU call 3 (1 GPU): diagnose the inner-iterations hang of call 2, the rest of the suite file by file, bench with the fused
# PCG kernels + compact camera record.   gpurun --timeout 1500 -- 'bash tools/gpu_r2_call3.sh'
set -u
OUT=gpurun_out/r2c3
mkdir -p "$OUT"
run() {  # run <seconds> <logname> <command...>
  local t=$1 name=$2; shift 2
  echo "=== $name: $*" | tee -a "$OUT/summary.txt"
  local t0=$(date +%s)
  timeout "$t" "$@" > "$OUT/$name.log" 2>&1
  echo "exit $? after $(( $(date +%s) - t0 )) s" | tee -a "$OUT/summary.txt"
  tail -4 "$OUT/$name.log" | cut -c1-400 >> "$OUT/summary.txt"
}
run 160 inner_stream python -X faulthandler -m pytest tests/test_xx_inner_iterations_gpu.py -m gpu -x -q -o faulthandler_timeout=90
TBA_MATVEC=tile run 160 inner_tile python -X faulthandler -m pytest tests/test_xx_inner_iterations_gpu.py -m gpu -x -q -o faulthandler_timeout=90
run 300 parity python -m pytest tests/test_gpu_parity.py -m gpu -q -x
for f in test_x_bench_sequence_gpu test_x_exact_schur_gpu test_x_fountain_gpu test_x_fullsize_gpu test_x_track_filter_gpu test_xx_camera_models_gpu \
         test_xx_matcher_gpu test_xx_track_estimator_gpu test_xx_two_view_gpu test_z_adapter_gpu test_zz_experiments_gpu; do
  run 300 "$f" python -m pytest "tests/$f.py" -q -m gpu
done
run 400 bench_c3 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-experiments
run 300 bench_c2 python bench.py --workload c2_1kcam --steps 10 --warmup 3 --no-cpu-baseline --no-experiments
run 300 bench_c4 python bench.py --workload c4_radtan --steps 10 --warmup 3 --no-cpu-baseline --no-experiments
run 400 ncu_lin ncu --set full --clock-control none --import-source on -k "regex:k_linearize|k_cost|k_schur_stream" -c 6 \
    -o "$OUT/r2_lin" -f python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-e2e --no-experiments
cat "$OUT/summary.txt" | cut -c1-300
