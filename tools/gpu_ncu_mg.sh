#!/bin/bash
# Full ncu capture of the fine-level multigrid legs (17 consecutive k_vleg launches span one whole V-cycle)
# and of the projection / diagnostics glue kernels.
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on \
    --kernel-name regex:'k_vleg' --launch-skip 192 --launch-count 17 \
    -o gpurun_out/full_mg -f python bench.py --steps 1 --warmup 3 --no-cpu --e2e-steps 1 > gpurun_out/b_ncu_full_mg.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on \
    --kernel-name regex:'k_diag_post|k_extract_gradp|k_div_embed|k_resid|k_ts' --launch-skip 12 --launch-count 8 \
    -o gpurun_out/full_glue -f python bench.py --steps 1 --warmup 3 --no-cpu --e2e-steps 1 > gpurun_out/b_ncu_full_glue.log 2>&1
python tools/kbench.py --what mg > gpurun_out/kbench_mg.log 2>&1
ls -la gpurun_out
