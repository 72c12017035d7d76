#!/bin/bash
# Round-2 measurement pass on one B200 (run under gpurun): bench lines of both arms, ncu launch list, one `ncu --set full`
# capture per kernel of the training step (fused forward writing the rest-pose scratch), of the forward-only launch and of the
# FK kernels.  Every ncu command runs only after the same command has exited 0 without ncu.
set -x
O=gpurun_out
python bench.py > $O/bench_default_1gpu.json 2> $O/bench_default.err || exit 1
python bench.py --impl reference > $O/bench_reference_arm.json 2> $O/bench_reference.err
B="python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-extras"
$B > /dev/null 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_r2.csv $B > $O/ncu_launches_r2.log 2>&1
# training step = 6 launches: pose_forward_lh, vs_bones_operand, vskin_forward, skin_backward, blend_tc_backward, pose_backward_lh;
# 3 warm-up steps are skipped, the first timed step is captured
ncu --set full --clock-control none --import-source on \
    -k 'regex:pose_forward_lh_kernel|vs_bones_operand_kernel|vskin_forward_kernel|skin_backward_kernel|blend_tc_backward_kernel|pose_backward_lh_kernel' \
    --launch-skip 18 --launch-count 6 -o $O/prof_r2_step -f $B > $O/ncu_full_step.log 2>&1
tail -2 $O/ncu_full_step.log
# forward-only launches of the same command: 7 training steps (3 warm-up + 2 timed + 2 stage-timed) x 2 launches of this set come
# first, then 3 forward warm-ups; capture the first timed forward (bone-operand pre-pass + fused kernel, no scratch written)
ncu --set full --clock-control none --import-source on -k 'regex:vskin_forward_kernel|vs_bones_operand_kernel' \
    --launch-skip 20 --launch-count 2 -o $O/prof_r2_fused_fwd -f $B > $O/ncu_full_fwd.log 2>&1
tail -2 $O/ncu_full_fwd.log
# FK kernels at 2^20 - 1 samples (a partial last tile): the separate pair and the pair with the L2Loss terms inside
F="python bench.py --workload fk --hands 1048575 --steps 4 --warmup 3 --no-cpu-baseline"
$F > $O/bench_fk_1m.json 2> /dev/null || exit 1
python bench.py --workload fk --no-cpu-baseline > $O/bench_fk_65536.json 2> /dev/null
ncu --set full --clock-control none --import-source on -k 'regex:fk_forward_kernel|fk_backward_kernel' \
    --launch-skip 8 --launch-count 2 -o $O/prof_r2_fk -f $F > $O/ncu_full_fk.log 2>&1
tail -2 $O/ncu_full_fk.log
