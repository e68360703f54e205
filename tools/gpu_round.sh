#!/bin/bash
# One GPU-box visit for the record: parity tests, full ncu capture of the dominant kernels (-> DRAM traffic and fp64
# instruction counts per launch for the bench line), the judged bench line, the reference arm, the ncu launch list,
# smoke().  Everything lands in gpurun_out/ (scratch); summaries are copied to profiles/ by hand.
TAG=${1:-r2}
mkdir -p gpurun_out
make -s -j5 -C nyles_b200/csrc && make -s -C oracle all
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.max.mem --format=csv > gpurun_out/gpu.txt
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/${TAG}_pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/${TAG}_pytest_gpu.log
grep -E "passed|failed" gpurun_out/${TAG}_pytest_gpu.log | tail -2
timeout 700 ncu --set full --clock-control none --import-source on \
    --kernel-name regex:'k_mom3|k_momentum|k_upwind2|k_up3' --launch-skip 12 --launch-count 6 \
    -o gpurun_out/${TAG}_full_rhs -f python bench.py --steps 1 --warmup 3 --no-cpu --e2e-steps 1 > gpurun_out/${TAG}_ncu_full_rhs.log 2>&1
timeout 700 ncu --set full --clock-control none --import-source on \
    --kernel-name regex:'k_vleg' --launch-skip 100 --launch-count 12 \
    -o gpurun_out/${TAG}_full_mg -f python bench.py --steps 1 --warmup 3 --no-cpu --e2e-steps 1 > gpurun_out/${TAG}_ncu_full_mg.log 2>&1
python tools/ncu_traffic.py gpurun_out/ncu_traffic.json 134217728 gpurun_out/${TAG}_full_rhs.ncu-rep gpurun_out/${TAG}_full_mg.ncu-rep > gpurun_out/${TAG}_ncu_traffic.log 2>&1
python bench.py > gpurun_out/${TAG}_bench_n1.json 2> gpurun_out/${TAG}_bench_n1.err
python tools/show_bench.py gpurun_out/${TAG}_bench_n1.json | head -16
[ -n "$SKIP_REF" ] || python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${TAG}_bench_ref.json 2> gpurun_out/${TAG}_bench_ref.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu --e2e-steps 1 > gpurun_out/${TAG}_ncu_launches.log 2>&1
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1; tail -1 gpurun_out/${TAG}_smoke.log
