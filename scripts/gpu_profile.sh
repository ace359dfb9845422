# ncu --set full captures of the top kernels of the bf16x3 step (one launch each, warm: -s skips the first step's launches)
mkdir -p gpurun_out
B="python bench.py --steps 1 --warmup 1 --math bf16x3 --skip-e2e --no-cpu-baseline"
ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel -s 120 -c 2 -o gpurun_out/r1_conv_tc -f $B > gpurun_out/ncu_conv_tc.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:stem_forward_kernel -s 2 -c 1 -o gpurun_out/r1_stem_fwd -f $B > gpurun_out/ncu_stem_fwd.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:stem_wgrad_kernel -s 2 -c 1 -o gpurun_out/r1_stem_wgrad -f $B > gpurun_out/ncu_stem_wgrad.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:nce_gather_kernel -s 2 -c 1 -o gpurun_out/r1_nce -f $B > gpurun_out/ncu_nce.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:bn_relu_backward_apply -s 90 -c 1 -o gpurun_out/r1_bn_bwd -f $B > gpurun_out/ncu_bn.log 2>&1
ls -la gpurun_out/*.ncu-rep
