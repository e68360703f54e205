#!/bin/bash
mkdir -p gpurun_out
make -s -j5 -C nyles_b200/csrc
timeout 300 ncu --set full --import-source on --clock-control none -k regex:k_vcycle_tail --launch-skip 60 --launch-count 1 -o gpurun_out/tail_lock -f python bench.py --workload lock --steps 20 --warmup 3 --no-cpu --e2e-steps 1 > gpurun_out/tail_lock_ncu.log 2>&1
tail -3 gpurun_out/tail_lock_ncu.log
python bench.py --workload lock --steps 200 --warmup 5 --no-cpu --e2e-steps 10 > gpurun_out/tail_bench_lock.json 2> gpurun_out/tail_bench_lock.err
python tools/show_bench.py gpurun_out/tail_bench_lock.json | head -14
