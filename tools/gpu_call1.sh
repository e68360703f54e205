#!/bin/bash
# round 2, first visit: full GPU suite (with the 256^3 oracle comparisons), plume256 sanity, sanitizers
mkdir -p gpurun_out
make -s -j5 -C nyles_b200/csrc && make -s -C oracle all
( time timeout 1200 python -m pytest tests -m gpu -x -q ) > gpurun_out/r2a_pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/r2a_pytest_gpu.log
tail -5 gpurun_out/r2a_pytest_gpu.log
timeout 300 python bench.py --workload plume256 --steps 5 --warmup 3 --no-cpu --e2e-steps 2 > gpurun_out/r2a_bench_plume256.json 2> gpurun_out/r2a_bench_plume256.err
timeout 300 python bench.py --workload plume256 --steps 5 --warmup 3 --no-cpu --e2e-steps 2 --unfused-forcing > gpurun_out/r2a_bench_plume256_unfused.json 2> gpurun_out/r2a_bench_plume256_unfused.err
timeout 600 python bench.py --steps 10 --warmup 3 --e2e-steps 5 > gpurun_out/r2a_bench_n1.json 2> gpurun_out/r2a_bench_n1.err
tools/sanitize.sh memcheck synccheck
nproc; free -g | head -2
