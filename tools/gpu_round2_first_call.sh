#!/bin/bash
# First GPU call of round 2 (NOTES.md section 1), ready to paste:
#   /usr/local/graft/bin/gpurun --timeout 1500 -- 'bash tools/gpu_round2_first_call.sh'
# Every step has its own inner timeout; nothing here can hang the box.  Results land in gpurun_out/r2_first/.
# Budget: about 12-15 GPU-minutes on one B200.
set -u
OUT=gpurun_out/r2_first
mkdir -p "$OUT"
run() {  # run <seconds> <logname> <command...>
  local t=$1 name=$2; shift 2
  echo "=== $name: $*" | tee -a "$OUT/summary.txt"
  timeout "$t" "$@" > "$OUT/$name.log" 2>&1
  echo "exit $?" | tee -a "$OUT/summary.txt"
  tail -3 "$OUT/$name.log" >> "$OUT/summary.txt"
}
# 1. the hardware-verified parity suite first, then everything that has never run on hardware (one file at a time so that
#    one failure does not hide the others)
run 600 parity python -m pytest tests/test_gpu_parity.py -q -m gpu -x
for f in test_x_fountain_gpu test_x_fullsize_gpu test_xx_matcher_gpu test_x_track_filter_gpu test_xx_camera_models_gpu \
         test_x_exact_schur_gpu test_xx_inner_iterations_gpu test_xx_track_estimator_gpu test_xx_two_view_gpu test_z_adapter_gpu; do
  run 420 "$f" python -m pytest "tests/$f.py" -q -m gpu
done
# 2. smoke + the bench line (default workload c3, N = 1)
run 300 smoke python -c "import __graft_entry__ as g; g.smoke()"
run 600 bench_c3 python bench.py
# 3. the compiled-in experiments: parity subset, then the bench line (its "stage_ms_per_step" gives the per-kernel effect).
#    TBA_TRED first: it is the one expected to matter most (a third of the L2 sector operations of every RED-bound kernel).
TBA_TRED=1 run 300 parity_tred python -m pytest tests/test_gpu_parity.py -q -m gpu -k "stage_parity or full_solve"
TBA_TRED=1 run 400 bench_c3_tred python bench.py --no-cpu-baseline
TBA_TRED=1 TBA_PACK_SORT=1 run 400 bench_c3_tred_packsort python bench.py --no-cpu-baseline
TBA_MATVEC_BULKRED=1 run 300 parity_bulkred python -m pytest tests/test_gpu_parity.py -q -m gpu -k "stage_parity or full_solve"
TBA_MATVEC_BULKRED=1 run 400 bench_c3_bulkred python bench.py --no-cpu-baseline
TBA_FAST_SEG=1 run 300 parity_fastseg python -m pytest tests/test_gpu_parity.py -q -m gpu -k "stage_parity or full_solve"
TBA_FAST_SEG=1 run 400 bench_c3_fastseg python bench.py --no-cpu-baseline
TBA_PACK_SORT=1 run 300 parity_packsort python -m pytest tests/test_gpu_parity.py -q -m gpu -k "stage_parity or full_solve"
TBA_PACK_SORT=1 run 400 bench_c3_packsort python bench.py --no-cpu-baseline
# 4. ncu: launch list of the bench command and one full capture of the matvec / linearise kernels
run 600 ncu_launches ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file "$OUT/launches.csv" \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-experiments
run 600 ncu_full_schur ncu --set full --clock-control none --import-source on -k regex:k_schur -s 4 -c 2 -o "$OUT/r2_schur" -f \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-experiments
run 600 ncu_full_linearize ncu --set full --clock-control none --import-source on -k regex:k_linearize -s 1 -c 1 -o "$OUT/r2_linearize" -f \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-experiments
cat "$OUT/summary.txt"
