#!/bin/bash
# compute-sanitizer passes over the small-grid GPU parity tests ("sanitizers"):
#   memcheck  : out-of-bounds / misaligned global, shared and local accesses (every kernel family)
#   racecheck : shared-memory hazards (k_vleg's TMA ring + y/z tiles, k_vcycle_tail, k_upwind2's flux exchange)
#   synccheck : divergent / invalid barrier and mbarrier use
# Each pass is bounded by `timeout`; logs go to gpurun_out/san_<tool>.log, a one-line verdict per tool to
# gpurun_out/san_summary.txt.  Usage: tools/sanitize.sh [memcheck racecheck synccheck]
mkdir -p gpurun_out
TOOLS=${@:-memcheck racecheck synccheck}
# small and quick, but through every kernel: operators (all closures), fused RHS + time scheme, multigrid level
# operations, complete solves, fused legs (TMA) against the per-operator kernels, the one-launch tail, one model run
SEL_FAST='test_upwind or test_vortex_force_and_bernoulli or test_vorticity or test_kin_div or test_rhs_step or test_halo_fill_self or test_level_operations or test_solve_point_sources or test_fused_legs_equal or test_weno3 or test_timescheme_kernels or test_max_speed2'
SEL_SLOW='test_upwind or test_rhs_step or test_level_operations or test_fused_legs_equal'
: > gpurun_out/san_summary.txt
for tool in $TOOLS; do
    sel="$SEL_FAST"; lim=1500
    if [ "$tool" = racecheck ]; then sel="$SEL_SLOW"; lim=1700; fi
    log=gpurun_out/san_$tool.log
    start=$(date +%s)
    timeout $lim compute-sanitizer --tool $tool --error-exitcode 86 --print-limit 20 \
        python -m pytest tests/test_gpu_operators.py tests/test_gpu_multigrid.py -q -x -m gpu -k "$sel" \
        -p no:cacheprovider > $log 2>&1
    rc=$?
    took=$(( $(date +%s) - start ))
    errs=$(grep -c "^=========.*\(Invalid\|Race\|hazard\|Barrier error\|Error:\)" $log)
    tail_line=$(grep -E "passed|failed|error" $log | tail -1)
    verdict=$(grep -E "ERROR SUMMARY|RACECHECK SUMMARY" $log | tail -1)
    echo "$tool: exit $rc, ${took}s, sanitizer error lines $errs | pytest: $tail_line | $verdict" | tee -a gpurun_out/san_summary.txt
done
# the model path end to end under memcheck (three LES steps against the oracle)
if echo "$TOOLS" | grep -q memcheck; then
    timeout 900 compute-sanitizer --tool memcheck --error-exitcode 86 --print-limit 20 \
        python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/san_memcheck_smoke.log 2>&1
    echo "memcheck smoke(): exit $? | $(grep -E 'ERROR SUMMARY' gpurun_out/san_memcheck_smoke.log | tail -1) | $(grep 'smoke ok' gpurun_out/san_memcheck_smoke.log)" | tee -a gpurun_out/san_summary.txt
fi
