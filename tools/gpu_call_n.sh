#!/bin/bash
# N GPUs (gpurun --gpus N -- tools/gpu_call_n.sh N): slab parity at N ranks, weak-scaling bench with the slab
# self-check, the plume (configs[3]) on N slabs
N=${1:-8}
mkdir -p gpurun_out
make -s -j5 -C nyles_b200/csrc && make -s -C oracle all
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
( time timeout 600 $TR tests/dist_check.py ) > gpurun_out/r2f_dist_check_n$N.log 2>&1
grep -E "dist_check|FAIL" gpurun_out/r2f_dist_check_n$N.log | head -5
timeout 900 $TR bench.py --gpus $N --steps 10 --warmup 3 --e2e-steps 3 > gpurun_out/r2f_bench_n$N.json 2> gpurun_out/r2f_bench_n$N.err
python tools/show_bench.py gpurun_out/r2f_bench_n$N.json | head -15
grep -o '"parity_check": "[a-zA-Z]*"' gpurun_out/r2f_bench_n$N.json; grep -o '"nvlink": {[^}]*}' gpurun_out/r2f_bench_n$N.json
if [ "$N" = "8" ] && [ -n "$PLUME" ]; then
  timeout 900 $TR bench.py --gpus $N --workload plume --steps 8 --warmup 3 --e2e-steps 2 --no-selfcheck > gpurun_out/r2f_bench_plume_n$N.json 2> gpurun_out/r2f_bench_plume_n$N.err
  python tools/show_bench.py gpurun_out/r2f_bench_plume_n$N.json | head -15; tail -2 gpurun_out/r2f_bench_plume_n$N.err
  NY_MG_GATHER_CELLS=2200000 timeout 900 $TR bench.py --gpus $N --steps 10 --warmup 3 --e2e-steps 1 --no-selfcheck > gpurun_out/r2f_bench_n${N}_gather128.json 2> gpurun_out/r2f_bench_n${N}_gather128.err
  python tools/show_bench.py gpurun_out/r2f_bench_n${N}_gather128.json | head -4
fi
