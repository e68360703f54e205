#!/bin/bash
mkdir -p gpurun_out
make -s -j5 -C nyles_b200/csrc
for v in 0 200000 3000000 20000000; do
  for wl in lock tgv256; do
    NY_MG_LEG_MIN_CELLS=$v timeout 200 python bench.py --workload $wl --steps 20 --warmup 3 --no-cpu --e2e-steps 1 > gpurun_out/tmp_leg.json 2>/dev/null
    echo "leg_min_cells $v $wl: $(python -c "
import json; d=json.load(open('gpurun_out/tmp_leg.json')); s=d['roofline_step']['kernel_time_share']; print('%.3f ms/step' % d['ms_per_step'], {k: round(v*d['ms_per_step'],2) for k,v in list(s.items())[:6]})")"
  done
done
