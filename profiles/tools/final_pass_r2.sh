#!/bin/bash
# Round-2 measurement pass on one B200 (run under gpurun): bench lines of both arms, ncu launch list, one `ncu --set full`
# capture per kernel of the training step and of the forward-only (fused) launch.  Every ncu command runs only after the
# same command has exited 0 without ncu.
set -x
O=gpurun_out
python bench.py > $O/bench_default_1gpu.json 2> $O/bench_default.err || exit 1
python bench.py --impl reference > $O/bench_reference_arm.json 2> $O/bench_reference.err
B="python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-extras"
$B > /dev/null 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_r2.csv $B > $O/ncu_launches_r2.log 2>&1
ncu --set full --clock-control none --import-source on \
    -k 'regex:pose_forward_lh_kernel|blend_tc_forward_mres|skin_forward_kernel|skin_backward_kernel|blend_tc_backward_kernel|pose_backward_lh_kernel' \
    --launch-skip 18 --launch-count 6 -o $O/prof_r2_step -f $B > $O/ncu_full_step.log 2>&1
tail -2 $O/ncu_full_step.log
# forward-only launches of the same command: 3 training warm-ups + 2 timed + 2 profiled steps come first (no fused kernel in
# them), then 3 forward warm-ups; capture the first timed forward (bone-operand pre-pass + fused kernel)
ncu --set full --clock-control none --import-source on -k 'regex:vskin_forward_kernel|vs_bones_operand_kernel' \
    --launch-skip 6 --launch-count 2 -o $O/prof_r2_fused_fwd -f $B > $O/ncu_full_fwd.log 2>&1
tail -2 $O/ncu_full_fwd.log
