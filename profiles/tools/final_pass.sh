#!/bin/bash
# Round-end measurement pass on one B200 (run under gpurun): bench lines, ncu launch list, one ncu --set full capture
# per kernel of the step.  Each ncu command only after its own command has exited 0 without ncu.
set -x
O=gpurun_out
python bench.py > $O/bench_default_1gpu.json 2> $O/bench_default.err || exit 1
python bench.py --impl reference > $O/bench_reference_arm.json 2> $O/bench_reference.err
python bench.py --workload fit --steps 20 --warmup 3 2>/dev/null | tail -1 > $O/bench_fit_config5.json
python bench.py --workload fk --steps 50 --warmup 5 2>/dev/null | tail -1 > $O/bench_fk_config3.json
python bench.py --hands 4096 --rotate 16 --steps 200 --warmup 10 --no-e2e --no-cpu-baseline 2>/dev/null | tail -1 > $O/bench_config2_4096.json
python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > /dev/null 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_final.csv \
    python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > $O/ncu_launches.log 2>&1
ncu --set full --clock-control none --import-source on \
    -k 'regex:pose_forward_lh_kernel|blend_tc_forward_mres|skin_forward_kernel|skin_backward_kernel|blend_tc_backward_kernel|pose_backward_lh_kernel' \
    --launch-skip 18 --launch-count 6 -o $O/prof_r1_final2 -f python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > $O/ncu_full.log 2>&1
tail -3 $O/ncu_full.log
