#!/bin/bash
# Round 2, 2-GPU call: multi-GPU tests (fused matvec + peer-memory all-reduce), bench N=2 with and without it, matcher N=2.
#   /usr/local/graft/bin/gpurun --gpus 2 --timeout 900 -- 'bash tools/gpu_r2_multi2.sh'   (charged 2x)
set -u
OUT=gpurun_out/r2_multi2
mkdir -p "$OUT"
run() {  # run <seconds> <logname> <command...>
  local t=$1 name=$2; shift 2
  echo "=== $name: $*" | tee -a "$OUT/summary.txt"
  local t0=$(date +%s)
  timeout "$t" "$@" > "$OUT/$name.log" 2>&1
  echo "exit $? after $(( $(date +%s) - t0 )) s" | tee -a "$OUT/summary.txt"
  tail -5 "$OUT/$name.log" | cut -c1-500 >> "$OUT/summary.txt"
}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
nvidia-smi topo -m > "$OUT/topo.txt" 2>&1
run 300 multi_tests python -m pytest tests/test_y_multi_gpu.py -q -m gpu -x
run 200 bench_c3_n2_p2p $TR --master-port 29502 bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu-baseline --no-experiments
TBA_P2P=0 run 200 bench_c3_n2_nccl $TR --master-port 29503 bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu-baseline --no-experiments
run 150 bench_c3_n1 python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu-baseline --no-experiments
run 200 bench_c5_n2 $TR --master-port 29504 bench.py --workload c5_matcher --gpus 2 --steps 3 --warmup 1
run 200 bench_c5_n1 python bench.py --workload c5_matcher --steps 3 --warmup 1 --no-cpu-baseline
run 100 matcher_tests python -m pytest tests/test_xx_matcher_gpu.py -m gpu -q -x
cat "$OUT/summary.txt" | cut -c1-300
