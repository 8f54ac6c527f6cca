#!/bin/bash
# Round 2, second 2-GPU call: fused PCG kernel + merged collectives on two ranks (tests, bench N=2 fused / split PCG, traced), the
# launch-gap microbenchmark and the carve-out hint at N=1, one ncu capture of the matcher's exact pass.
#   /usr/local/graft/bin/gpurun --gpus 2 --timeout 900 -- 'bash tools/gpu_r2_multi2b.sh'   (charged 2x)
set -u
OUT=gpurun_out/r2_multi2b
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
run 300 multi_tests python -m pytest tests/test_y_multi_gpu.py -q -m gpu -x
run 200 bench_c3_n2 $TR --master-port 29502 bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu-baseline --no-experiments
TBA_PCG=split run 200 bench_c3_n2_split $TR --master-port 29503 bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu-baseline --no-experiments
TBA_TRACE_LM=1 run 200 bench_c3_n2_traced $TR --master-port 29505 bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu-baseline --no-experiments
run 150 bench_c3_n1 python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu-baseline --no-experiments
TBA_CARVEOUT=0 run 150 bench_c3_n1_nohint python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu-baseline --no-experiments
run 100 gaps python -c "
import ctypes
L = ctypes.CDLL('theiasfm_b200/libtheia_microbench_b200.so'); out = (ctypes.c_double * 5)(); print(L.tba_microbench_gaps(0, out), list(out))"
run 200 ncu_exact ncu --set full --clock-control none -k regex:k_exact_top2 -c 1 -o "$OUT/r2_exact" -f python bench.py --workload c5_matcher --steps 1 --warmup 0 --no-cpu-baseline
python profiles/summarize.py full "$OUT/r2_exact.ncu-rep" > "$OUT/ncu_exact_summary.txt" 2>&1
rm -f "$OUT/r2_exact.ncu-rep"
run 200 bench_c5_n2 $TR --master-port 29504 bench.py --workload c5_matcher --gpus 2 --steps 3 --warmup 1
find gpurun_out -size +8M -delete
cat "$OUT/summary.txt" | cut -c1-300
