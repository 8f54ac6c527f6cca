#!/bin/bash
# Round 2, GPU call 1 (1 GPU):  gpurun --timeout 2400 -- 'bash tools/gpu_r2_call1.sh'
# Every step runs under its own inner timeout; results in gpurun_out/r2c1/.
set -u
OUT=gpurun_out/r2c1
mkdir -p "$OUT"
run() {  # run <seconds> <logname> <command...>
  local t=$1 name=$2; shift 2
  echo "=== $name: $*" | tee -a "$OUT/summary.txt"
  local t0=$(date +%s)
  timeout "$t" "$@" > "$OUT/$name.log" 2>&1
  echo "exit $? after $(( $(date +%s) - t0 )) s" | tee -a "$OUT/summary.txt"
  tail -4 "$OUT/$name.log" >> "$OUT/summary.txt"
}
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > "$OUT/smi.txt" 2>&1
nproc > "$OUT/nproc.txt"; lscpu | head -25 >> "$OUT/nproc.txt"
# 1. the whole single-GPU suite (new: six losses, config 2 / config 4 full-size trajectories)
run 1200 pytest_gpu python -m pytest tests -m gpu -q -x --durations=15
# 2. the bench line (c3, new defaults) with the diagnostic pass (round-1 kernels, matvec ablations)
run 900 bench_c3 python bench.py --steps 20 --warmup 5
# 3. the other single-GPU configs of BASELINE.json
run 400 bench_c2 python bench.py --workload c2_1kcam --steps 10 --warmup 3 --no-cpu-baseline --no-experiments
run 400 bench_c4 python bench.py --workload c4_radtan --steps 10 --warmup 3 --no-cpu-baseline --no-experiments
# 4. reference arm twice (stability of the pinned CPU baseline)
run 300 ref_a python bench.py --impl reference --steps 5 --warmup 1
run 300 ref_b python bench.py --impl reference --steps 5 --warmup 1
# 5. ncu: launch list of c3 and full captures of the shipped kernels
run 600 ncu_launches ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file "$OUT/launches_c3.csv" \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-experiments
run 600 ncu_full_schur ncu --set full --clock-control none --import-source on -k regex:k_schur -c 4 -o "$OUT/r2_schur" -f \
    python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-e2e --no-experiments
run 600 ncu_full_other ncu --set full --clock-control none --import-source on -k "regex:k_linearize|k_precond_ext|k_precond_intr|k_cost" -c 5 \
    -o "$OUT/r2_other" -f python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-e2e --no-experiments
# 6. compute-sanitizer on a small parity subset
SUB='test_residuals_match_golden or (test_stage_parity and pinhole_shared and not True-) or (test_full_solve_parity and radtan_per_camera and 1)'
run 500 sanitizer_memcheck compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "$SUB"
run 500 sanitizer_racecheck compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "$SUB"
cat "$OUT/summary.txt"
