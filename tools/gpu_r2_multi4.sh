#!/bin/bash
# Round 2, 4-GPU call (charged 4x: keep it short): the world-4 cases of the multi-GPU tests, bench c3 and c5 at N=4 with the final code.
set -u
OUT=gpurun_out/r2_multi4
mkdir -p "$OUT"
run() {  # run <seconds> <logname> <command...>
  local t=$1 name=$2; shift 2
  echo "=== $name" | tee -a "$OUT/summary.txt"
  local t0=$(date +%s)
  timeout "$t" "$@" > "$OUT/$name.log" 2>&1
  echo "exit $? after $(( $(date +%s) - t0 )) s" | tee -a "$OUT/summary.txt"
  tail -3 "$OUT/$name.log" | cut -c1-400 >> "$OUT/summary.txt"
}
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
run 200 multi_tests python -m pytest tests/test_y_multi_gpu.py -q -m gpu -x
run 120 bench_c3_n4 $TR --nproc-per-node 4 --master-port 29504 bench.py --gpus 4 --steps 10 --warmup 3 --no-cpu-baseline --no-experiments
TBA_TRACE_LM=1 run 120 bench_c3_n4_traced $TR --nproc-per-node 4 --master-port 29505 bench.py --gpus 4 --steps 10 --warmup 3 --no-cpu-baseline --no-experiments
run 100 bench_c5_n4 $TR --nproc-per-node 4 --master-port 29510 bench.py --workload c5_matcher --gpus 4 --steps 3 --warmup 1
cat "$OUT/summary.txt" | cut -c1-300
