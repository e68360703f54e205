#!/bin/bash
mkdir -p gpurun_out
make -s -j5 -C nyles_b200/csrc && make -s -C oracle all
( time timeout 900 python -m pytest tests/test_gpu_operators.py tests/test_gpu_model.py -m gpu -x -q ) > gpurun_out/r2d_pytest.log 2>&1
grep -E "passed|failed|error" gpurun_out/r2d_pytest.log | tail -3; grep -E "^FAILED|^ERROR|Error" gpurun_out/r2d_pytest.log | head -5
timeout 300 python tools/kbench.py --what rhs --mom 2 --sustained 1 2>&1 | head -5 | tee gpurun_out/r2d_kbench_rhs.log
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name regex:'k_up3' --launch-skip 4 --launch-count 1 \
   -o gpurun_out/r2d_up3 -f python tools/kbench.py --what rhs --mom 2 --n 512 2>&1 | tail -2 | cut -c1-200
