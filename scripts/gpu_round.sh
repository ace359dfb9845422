# full GPU check: smoke, gpu tests, bench at each math mode, launch list of the TC path
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke.log
timeout 1500 python -m pytest tests -m gpu -q --no-header -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
grep -E "^(FAILED|ERROR|[0-9]+ (passed|failed))" gpurun_out/pytest_gpu.log | head -40
for m in bf16x3 bf16; do
timeout 600 python bench.py --steps 10 --warmup 3 --math $m --no-cpu-baseline > gpurun_out/bench_$m.json 2> gpurun_out/bench_$m.err; echo "bench $m rc=$?"; tail -2 gpurun_out/bench_$m.err; cat gpurun_out/bench_$m.json | cut -c1-1500
done
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_bf16x3.csv python bench.py --steps 1 --warmup 1 --math bf16x3 --skip-e2e --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
wc -l gpurun_out/launches_bf16x3.csv
