#!/bin/bash
# the 512^3 oracle comparison (benchmarked grid): Rayleigh-Taylor closed LES, diag0 / one RHS / Euler + 3 LFAM3 steps
mkdir -p gpurun_out
make -s -j5 -C nyles_b200/csrc && make -s -C oracle all
rm -f gpurun_out/r2h_parity512.json
( time NY_LARGE_N=512 NY_LARGE_RECORD=gpurun_out/r2h_parity512.json timeout 1500 python -m pytest tests/test_gpu_large.py -m gpu -x -q -k "rt" ) > gpurun_out/r2h_parity512.log 2>&1
tail -5 gpurun_out/r2h_parity512.log; python -c "
import json; r=json.load(open('gpurun_out/r2h_parity512.json'))
for c in r: print({k: c[k] for k in ('case','n','diag0','rhs0','final','vcycles_per_step','oracle_seconds_per_step')})"
