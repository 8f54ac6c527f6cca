#!/bin/bash
# Round 2, GPU call 13 (1 GPU): exhaustive scans of the matcher's exact pass by the whole CTA through shared memory (tests + bench),
# device-side span of the PCG launches against its wall clock, cost of the stage events inside the timed region.
set -u
OUT=gpurun_out/r2c13
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
run 200 bench_c5 python bench.py --workload c5_matcher --steps 3 --warmup 1 --no-cpu-baseline
TBA_UPLOAD_TRACE=1 TBA_TRACE_LM=1 run 300 bench_c3_traced python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-experiments
TBA_BENCH_NOPROF=1 run 300 bench_c3_noprof python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-experiments
run 300 bench_c3 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-experiments
run 200 ncu_exact ncu --set full --clock-control none -k regex:k_exact_top2 -c 1 -o "$OUT/r2_exact" -f python bench.py --workload c5_matcher --steps 1 --warmup 0 --no-cpu-baseline
python profiles/summarize.py full "$OUT/r2_exact.ncu-rep" > "$OUT/ncu_exact_summary.txt" 2>&1
rm -f "$OUT/r2_exact.ncu-rep"
find gpurun_out -size +8M -delete
cat "$OUT/summary.txt" | cut -c1-300
