#!/bin/bash
mkdir -p gpurun_out
make -s -j5 -C nyles_b200/csrc
for b in "4,2" "8,2" "8,4" "4,4" "16,2" "2,8"; do
  NY_GP_BLOCK=$b timeout 200 python bench.py --steps 6 --warmup 3 --no-cpu --e2e-steps 1 > gpurun_out/tmp_gp.json 2>/dev/null
  echo "block 32,$b: $(python -c "
import json; d=json.load(open('gpurun_out/tmp_gp.json')); s=d['roofline_step']['kernel_time_share']; print('%.2f ms/step, gradp_vorticity_ke %.2f ms/step' % (d['ms_per_step'], s.get('gradp_vorticity_ke',0)*d['ms_per_step']))")"
done
