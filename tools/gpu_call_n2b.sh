#!/bin/bash
mkdir -p gpurun_out
make -s -j5 -C nyles_b200/csrc && make -s -C oracle all
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
( time timeout 600 $TR tests/dist_check.py ) > gpurun_out/r2i_dist_check_n2.log 2>&1
grep -E "dist_check|FAIL" gpurun_out/r2i_dist_check_n2.log | head -5
timeout 600 $TR bench.py --gpus 2 --steps 10 --warmup 3 --e2e-steps 3 > gpurun_out/r2i_bench_n2.json 2> gpurun_out/r2i_bench_n2.err
python tools/show_bench.py gpurun_out/r2i_bench_n2.json | head -14
grep -o '"parity_check": "[a-zA-Z]*"' gpurun_out/r2i_bench_n2.json; grep -o '"nvlink": {[^}]*}' gpurun_out/r2i_bench_n2.json; tail -3 gpurun_out/r2i_bench_n2.err
