#!/bin/bash
# Round 2, GPU call 12 (1 GPU): fused PCG vector kernel (k_pcg_fused: one launch per CG iteration next to the matvec) against the split
# kernels, pack fill on all host threads; whole GPU suite (new variant split_pcg in test_zz_experiments_gpu.py).
set -u
OUT=gpurun_out/r2c12
mkdir -p "$OUT"
run() {  # run <seconds> <logname> <command...>
  local t=$1 name=$2; shift 2
  echo "=== $name: $*" | tee -a "$OUT/summary.txt"
  local t0=$(date +%s)
  timeout "$t" "$@" > "$OUT/$name.log" 2>&1
  echo "exit $? after $(( $(date +%s) - t0 )) s" | tee -a "$OUT/summary.txt"
  tail -4 "$OUT/$name.log" | cut -c1-600 >> "$OUT/summary.txt"
}
run 300 parity python -m pytest tests/test_gpu_parity.py -m gpu -q -x
run 300 bench_c3_fused python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-experiments
TBA_PCG=split run 300 bench_c3_split python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-experiments
TBA_UPLOAD_TRACE=1 TBA_TRACE_LM=1 run 300 bench_c3_traced python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-experiments
run 600 pytest_gpu python -m pytest tests -m gpu -q -x --durations=5
run 200 bench_c2 python bench.py --workload c2_1kcam --steps 10 --warmup 3 --no-cpu-baseline --no-experiments
run 200 bench_c4 python bench.py --workload c4_radtan --steps 10 --warmup 3 --no-cpu-baseline --no-experiments
find gpurun_out -size +8M -delete
cat "$OUT/summary.txt" | cut -c1-300
