# round-1 profile set (one B200): per-launch list of one warm bf16x3 step + ncu --set full captures of the top kernels
mkdir -p gpurun_out
B="python bench.py --steps 1 --warmup 1 --math bf16x3 --skip-e2e --no-cpu-baseline"
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active --clock-control none --csv --log-file gpurun_out/launches_bf16x3.csv $B > gpurun_out/bench_under_ncu.log 2>&1
wc -l gpurun_out/launches_bf16x3.csv
ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel -s 110 -c 3 -o gpurun_out/r1_conv_tc -f $B > gpurun_out/ncu_conv_tc.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:wgrad_tc_kernel -s 50 -c 2 -o gpurun_out/r1_wgrad_tc -f $B > gpurun_out/ncu_wgrad.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:stem_forward_kernel -s 2 -c 1 -o gpurun_out/r1_stem_fwd -f $B > gpurun_out/ncu_stem_fwd.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:stem_wgrad_kernel -s 2 -c 1 -o gpurun_out/r1_stem_wgrad -f $B > gpurun_out/ncu_stem_wgrad.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:nce_gather_kernel -s 12 -c 1 -o gpurun_out/r1_nce_k16384 -f python scripts/bench_nce.py --banks 2000000 --negatives 16384 --iters 3 --warmup 2 > gpurun_out/ncu_nce.log 2>&1
ls -la gpurun_out/*.ncu-rep
