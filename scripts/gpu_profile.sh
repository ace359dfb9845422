# profile set (one B200): ncu --set full captures of the top kernels of a warm bf16x3 step, NCE at K = 16384 and the CMA scan
mkdir -p gpurun_out
B="python bench.py --steps 1 --warmup 1 --math bf16x3 --skip-e2e --no-cpu-baseline"
ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel -s 110 -c 4 -o gpurun_out/r1_conv_tc -f $B > gpurun_out/ncu_conv_tc.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:wgrad_tc_kernel -s 48 -c 4 -o gpurun_out/r1_wgrad_tc -f $B > gpurun_out/ncu_wgrad.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:stem_forward_kernel -s 2 -c 1 -o gpurun_out/r1_stem_fwd -f $B > gpurun_out/ncu_stem_fwd.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:stem_wgrad_kernel -s 3 -c 1 -o gpurun_out/r1_stem_wgrad -f $B > gpurun_out/ncu_stem_wgrad.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:bn_relu_backward_apply_kernel -s 50 -c 1 -o gpurun_out/r1_bn_bwd -f $B > gpurun_out/ncu_bn_bwd.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:nce_gather_kernel -s 3 -c 1 -o gpurun_out/r1_nce_k16384 -f python scripts/bench_nce.py --banks 2000000 --negatives 16384 --iters 3 --warmup 2 > gpurun_out/ncu_nce.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:cma_scan_tc_kernel -c 1 -o gpurun_out/r1_cma_scan_tc -f python scripts/probe_cma.py 60000 > gpurun_out/ncu_cma.log 2>&1
ls -la gpurun_out/*.ncu-rep
