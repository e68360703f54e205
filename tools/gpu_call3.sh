#!/bin/bash
tools/gpu_round.sh r2j
# one reference-arm line on the workload itself (512^3, not the 256^3 sample): same_config record
python bench.py --impl reference --cpu-sample same --steps 2 --warmup 0 > gpurun_out/r2j_bench_ref_same_config.json 2> gpurun_out/r2j_bench_ref_same_config.err
python -c "
import json; d=json.load(open('gpurun_out/r2j_bench_ref_same_config.json')); print('reference arm on the 512^3 workload: %.3e cell-updates/s, %.0f ms/step' % (d['value'], d['ms_per_step']))"
