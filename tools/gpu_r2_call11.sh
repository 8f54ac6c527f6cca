#!/bin/bash
# Round 2, GPU call 11 (1 GPU): call 10's measurements again with SMALL outputs (call 10's two .ncu-rep files exceeded the 64 MiB that
# gpurun merges back: nothing came home) + the matcher with the decisions on the device and the trimmed epilogue.
set -u
OUT=gpurun_out/r2c11
mkdir -p "$OUT"
run() {  # run <seconds> <logname> <command...>
  local t=$1 name=$2; shift 2
  echo "=== $name: $*" | tee -a "$OUT/summary.txt"
  local t0=$(date +%s)
  timeout "$t" "$@" > "$OUT/$name.log" 2>&1
  echo "exit $? after $(( $(date +%s) - t0 )) s" | tee -a "$OUT/summary.txt"
  tail -4 "$OUT/$name.log" | cut -c1-600 >> "$OUT/summary.txt"
}
nproc > "$OUT/nproc.txt"
run 300 matcher_tests python -m pytest tests/test_xx_matcher_gpu.py -q -m gpu -x
run 200 bench_c5 python bench.py --workload c5_matcher --steps 3 --warmup 1 --no-cpu-baseline
TBA_UPLOAD_TRACE=1 TBA_TRACE_LM=1 run 300 bench_c3_traced python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-experiments
run 300 bench_c3 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-experiments
run 600 n3_bench python tools/bench_n3.py
cp gpurun_out/n3_bench.json "$OUT/n3_bench.json" 2>/dev/null
run 400 ncu_launches ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file "$OUT/launches_c3.csv" \
    python bench.py --steps 3 --warmup 1 --no-cpu-baseline --no-e2e --no-experiments
run 400 ncu_full_hot ncu --set full --clock-control none --import-source on -k "regex:k_schur_stream|k_prepare_stream|k_linearize|k_cost" -c 7 \
    -o "$OUT/r2_final_hot" -f python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-e2e --no-experiments
python profiles/summarize.py full "$OUT/r2_final_hot.ncu-rep" > "$OUT/ncu_full_hot_summary.txt" 2>&1
rm -f "$OUT/r2_final_hot.ncu-rep"
run 400 ncu_full_n3 ncu --set full --clock-control none -k "regex:k_estimate_tracks|k_adjust_tracks|k_two_view_ba" -c 8 \
    -o "$OUT/r2_n3" -f python tools/bench_n3.py --only-big --big 400000 --cpu-tracks 2000 --pairs 4000 --cpu-pairs 8 --repeat 1
python profiles/summarize.py full "$OUT/r2_n3.ncu-rep" > "$OUT/ncu_full_n3_summary.txt" 2>&1
rm -f "$OUT/r2_n3.ncu-rep"
find gpurun_out -size +8M -delete
du -sh gpurun_out | tee -a "$OUT/summary.txt"
cat "$OUT/summary.txt" | cut -c1-300
