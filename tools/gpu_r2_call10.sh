#!/bin/bash
# Round 2, GPU call 10 (1 GPU): whole GPU suite after the upload / LM-loop latency changes, stage tracer + upload trace on c3, N3 kernel
# throughput (tools/bench_n3.py), launch list of the shipped c3 step, full ncu captures of the shipped hot kernels and of the N3 kernels.
set -u
OUT=gpurun_out/r2c10
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
run 600 pytest_gpu python -m pytest tests -m gpu -q -x --durations=8
run 300 bench_c3 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-experiments
TBA_UPLOAD_TRACE=1 TBA_TRACE_LM=1 run 300 bench_c3_traced python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-experiments
run 600 n3_bench python tools/bench_n3.py
run 400 ncu_launches ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file "$OUT/launches_c3.csv" \
    python bench.py --steps 3 --warmup 1 --no-cpu-baseline --no-e2e --no-experiments
run 400 ncu_full_hot ncu --set full --clock-control none --import-source on -k "regex:k_schur_stream|k_prepare_stream|k_linearize|k_cost" -c 7 \
    -o "$OUT/r2_final_hot" -f python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-e2e --no-experiments
run 400 ncu_full_n3 ncu --set full --clock-control none --import-source on -k "regex:k_estimate_tracks|k_adjust_tracks|k_two_view_ba" -c 8 \
    -o "$OUT/r2_n3" -f python tools/bench_n3.py --only-big --big 400000 --cpu-tracks 2000 --pairs 4000 --cpu-pairs 8 --repeat 1
run 200 bench_c2 python bench.py --workload c2_1kcam --steps 10 --warmup 3 --no-cpu-baseline --no-experiments
run 200 bench_c4 python bench.py --workload c4_radtan --steps 10 --warmup 3 --no-cpu-baseline --no-experiments
cat "$OUT/summary.txt" | cut -c1-400
