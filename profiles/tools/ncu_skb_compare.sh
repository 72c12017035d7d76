B="python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-extras"
for v in default RING; do
  if [ "$v" = "default" ]; then unset MANO_B200_LIB; else export MANO_B200_LIB=$PWD/3dhandposeestimation_b200/libmano_b200_$v.so; fi
  ncu --set full --clock-control none --import-source on -k regex:skin_backward_kernel --launch-skip 3 --launch-count 1 -o gpurun_out/prof_skb_$v -f $B > gpurun_out/ncu_skb_$v.log 2>&1
  tail -1 gpurun_out/ncu_skb_$v.log
done
