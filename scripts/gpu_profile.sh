mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q --no-header -p no:cacheprovider -k "config1" > gpurun_out/pytest_cfg1.log 2>&1; tail -3 gpurun_out/pytest_cfg1.log
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r1_fp32.csv python bench.py --steps 1 --warmup 1 --skip-e2e --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
wc -l gpurun_out/launches_r1_fp32.csv
ncu --set full --clock-control none --import-source on -k regex:conv_igemm -s 40 -c 3 -o gpurun_out/prof_conv_fp32 -f python bench.py --steps 1 --warmup 1 --skip-e2e --no-cpu-baseline > gpurun_out/ncu_conv.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:nce_gather -s 1 -c 1 -o gpurun_out/prof_nce -f python bench.py --steps 1 --warmup 1 --skip-e2e --no-cpu-baseline > gpurun_out/ncu_nce.log 2>&1
ls -la gpurun_out
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r1_fp32.json 2> gpurun_out/bench_r1_fp32.err; cat gpurun_out/bench_r1_fp32.json
