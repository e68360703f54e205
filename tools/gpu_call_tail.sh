#!/bin/bash
# GPU visit for the wide V-cycle tail: bit-identity tests, then the threshold sweep on three workloads.
mkdir -p gpurun_out
make -s -j5 -C nyles_b200/csrc && make -s -C oracle all
( time timeout 600 python -m pytest tests/test_gpu_multigrid.py tests/test_gpu_model.py -x -q ) > gpurun_out/tail_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/tail_pytest.log
tail -5 gpurun_out/tail_pytest.log
timeout 500 python tools/sweep_tail.py ${SWEEP_ARGS} > gpurun_out/tail_sweep.jsonl 2> gpurun_out/tail_sweep.err
echo "sweep exit $?"
cat gpurun_out/tail_sweep.jsonl | cut -c1-330
tail -3 gpurun_out/tail_sweep.err
