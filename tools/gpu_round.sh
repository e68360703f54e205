#!/bin/bash
# One GPU-box visit: parity tests, full ncu capture of the dominant kernels (-> DRAM traffic per launch for the
# bench line), the judged bench line, the reference arm and the ncu launch list.  Everything lands in
# gpurun_out/ (scratch); summaries are copied to profiles/ by hand.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.max.mem --format=csv > gpurun_out/gpu.txt
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
timeout 700 ncu --set full --clock-control none --import-source on \
    --kernel-name regex:'k_momentum|k_upwind2' --launch-skip 8 --launch-count 4 \
    -o gpurun_out/full_rhs -f python bench.py --steps 1 --warmup 3 --no-cpu --e2e-steps 1 > gpurun_out/b_ncu_full_rhs.log 2>&1
timeout 700 ncu --set full --clock-control none --import-source on \
    --kernel-name regex:'k_vleg' --launch-skip 100 --launch-count 12 \
    -o gpurun_out/full_mg -f python bench.py --steps 1 --warmup 3 --no-cpu --e2e-steps 1 > gpurun_out/b_ncu_full_mg.log 2>&1
python tools/ncu_traffic.py gpurun_out/ncu_traffic.json gpurun_out/full_rhs.ncu-rep gpurun_out/full_mg.ncu-rep > gpurun_out/ncu_traffic.log 2>&1
python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu --e2e-steps 1 > gpurun_out/b_ncu.log 2>&1
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
ls -la gpurun_out
