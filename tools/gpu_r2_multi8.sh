#!/bin/bash
# Round 2, 8-GPU call (charged 8x: keep it short): bench c3 at N=8 with the fused peer-memory all-reduce and with NCCL, N=4, matcher N=8.
set -u
OUT=gpurun_out/r2_multi8
mkdir -p "$OUT"
run() {  # run <seconds> <logname> <command...>
  local t=$1 name=$2; shift 2
  echo "=== $name" | tee -a "$OUT/summary.txt"
  local t0=$(date +%s)
  timeout "$t" "$@" > "$OUT/$name.log" 2>&1
  echo "exit $? after $(( $(date +%s) - t0 )) s" | tee -a "$OUT/summary.txt"
}
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
run 120 bench_c3_n8_p2p $TR --nproc-per-node 8 --master-port 29508 bench.py --gpus 8 --steps 10 --warmup 3 --no-cpu-baseline --no-experiments
TBA_P2P=0 run 120 bench_c3_n8_nccl $TR --nproc-per-node 8 --master-port 29509 bench.py --gpus 8 --steps 10 --warmup 3 --no-cpu-baseline --no-experiments
run 120 bench_c3_n4_p2p $TR --nproc-per-node 4 --master-port 29504 bench.py --gpus 4 --steps 10 --warmup 3 --no-cpu-baseline --no-experiments
run 100 bench_c5_n8 $TR --nproc-per-node 8 --master-port 29510 bench.py --workload c5_matcher --gpus 8 --steps 3 --warmup 1
cat "$OUT/summary.txt"
