#!/bin/bash
mkdir -p gpurun_out
make -s -j5 -C nyles_b200/csrc && make -s -C oracle all
for kc in 127 64 32 20 12; do echo "kchunk $kc"; NY_MOM3_KCHUNK=$kc timeout 300 python tools/kbench.py --what rhs --mom 2 2>&1 | head -1; done | tee gpurun_out/r2b_kbench_kchunk.log
( time timeout 900 python -m pytest tests/test_gpu_operators.py tests/test_gpu_multigrid.py -m gpu -x -q ) > gpurun_out/r2b_pytest_ops.log 2>&1
tail -4 gpurun_out/r2b_pytest_ops.log
timeout 300 python tools/kbench.py --what mg 2>&1 | tee gpurun_out/r2b_kbench_mg.log
