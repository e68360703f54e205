#!/bin/bash
# The round record (tools/gpu_round.sh) plus one reference-arm line on the 512^3 workload itself (~70 s of host time).
TAG=${1:-r2j}
tools/gpu_round.sh $TAG
# one reference-arm line on the workload itself (512^3, not the 256^3 sample): same_config record
python bench.py --impl reference --cpu-sample same --steps 2 --warmup 0 > gpurun_out/${TAG}_bench_ref_same_config.json 2> gpurun_out/${TAG}_bench_ref_same_config.err
python -c "
import json; d=json.load(open('gpurun_out/${TAG}_bench_ref_same_config.json')); print('reference arm on the 512^3 workload: %.3e cell-updates/s, %.0f ms/step' % (d['value'], d['ms_per_step']))"
