#!/bin/bash
# two GPUs: slab parity test, bench with the slab self-check (weak 512^3 per GPU), configs[2] strong (one 512^3 box on 2 slabs)
mkdir -p gpurun_out
make -s -j5 -C nyles_b200/csrc && make -s -C oracle all
( time timeout 600 python -m pytest tests/test_gpu_slabs.py -m gpu -x -q ) > gpurun_out/r2c_pytest_slabs_n2.log 2>&1
tail -3 gpurun_out/r2c_pytest_slabs_n2.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 600 $TR bench.py --gpus 2 --steps 10 --warmup 3 --e2e-steps 3 > gpurun_out/r2c_bench_n2.json 2> gpurun_out/r2c_bench_n2.err
python tools/show_bench.py gpurun_out/r2c_bench_n2.json | head -14
timeout 600 $TR bench.py --gpus 2 --workload rt512strong --steps 10 --warmup 3 --e2e-steps 3 --no-selfcheck > gpurun_out/r2c_bench_rt512strong_n2.json 2> gpurun_out/r2c_bench_rt512strong_n2.err
python tools/show_bench.py gpurun_out/r2c_bench_rt512strong_n2.json | head -14
grep -o '"parity_check[^}]*}' gpurun_out/r2c_bench_n2.json | cut -c1-400
