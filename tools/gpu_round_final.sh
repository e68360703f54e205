#!/bin/bash
# Round record at the final sources: tools/gpu_round.sh, the small configs (configs[0] lock exchange, configs[1] TGV 256^3),
# and memcheck / synccheck over the multigrid solves that run the cooperative V-cycle tail (grid barrier).
TAG=${1:-r2m}
SKIP_REF=1 tools/gpu_round.sh $TAG
python bench.py --workload lock --steps 200 --warmup 5 --no-cpu --e2e-steps 10 > gpurun_out/${TAG}_bench_lock.json 2> gpurun_out/${TAG}_bench_lock.err
python tools/show_bench.py gpurun_out/${TAG}_bench_lock.json | head -3
python bench.py --workload tgv256 --steps 10 --warmup 3 --no-cpu --e2e-steps 3 > gpurun_out/${TAG}_bench_tgv256.json 2> gpurun_out/${TAG}_bench_tgv256.err
python tools/show_bench.py gpurun_out/${TAG}_bench_tgv256.json | head -3
for tool in memcheck synccheck; do
    timeout 170 compute-sanitizer --tool $tool --error-exitcode 86 --print-limit 20 \
        python -m pytest tests/test_gpu_multigrid.py -q -x -m gpu -k "test_solve_point_sources" -p no:cacheprovider \
        > gpurun_out/${TAG}_san_${tool}_tail.log 2>&1
    echo "$tool (V-cycle tail, solves against the oracle): exit $? | $(grep -E 'passed|failed' gpurun_out/${TAG}_san_${tool}_tail.log | tail -1) | $(grep -E 'ERROR SUMMARY' gpurun_out/${TAG}_san_${tool}_tail.log | tail -1)" | tee -a gpurun_out/${TAG}_san_tail_summary.txt
done
