# launch list of ONE warm bf16x3 training step with time, DRAM bytes and tensor-pipe activity per launch
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active --clock-control none --csv --log-file gpurun_out/launches_bf16x3.csv python bench.py --steps 1 --warmup 1 --math bf16x3 --skip-e2e --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
wc -l gpurun_out/launches_bf16x3.csv
