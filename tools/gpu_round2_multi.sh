#!/bin/bash
# Multi-GPU calls of round 2 (every step under its own inner timeout; torchrun rendezvous on 127.0.0.1).
#   /usr/local/graft/bin/gpurun --gpus 2 --timeout 900 -- 'bash tools/gpu_round2_multi.sh 2'     (charged 2x)
#   /usr/local/graft/bin/gpurun --gpus 8 --timeout 600 -- 'bash tools/gpu_round2_multi.sh 8'     (charged 8x: bench only)
set -u
N=${1:-2}
OUT=gpurun_out/r2_multi_$N
mkdir -p "$OUT"
run() { local t=$1 name=$2; shift 2; echo "=== $name: $*" | tee -a "$OUT/summary.txt"; timeout "$t" "$@" > "$OUT/$name.log" 2>&1; echo "exit $?" | tee -a "$OUT/summary.txt"; tail -3 "$OUT/$name.log" >> "$OUT/summary.txt"; }
bench() {  # bench <n>
  run 400 "bench_n$1" python -m torch.distributed.run --nnodes=1 --nproc-per-node "$1" --master-addr 127.0.0.1 --master-port $((29500 + $1)) \
      bench.py --gpus "$1" --steps 5 --warmup 3
}
if [ "$N" -le 2 ]; then
  run 600 multi_tests python -m pytest tests/test_y_multi_gpu.py -q -m gpu     # sharded world-2 parity, tba_solve_multi (subprocess), two-view sharding
  bench 2
else
  for n in 2 4 "$N"; do [ "$n" -le "$N" ] && bench "$n"; done
fi
cat "$OUT/summary.txt"
