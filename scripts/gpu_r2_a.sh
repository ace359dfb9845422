# round 2, call A: descriptor probe, GPU tests, bench (default / two-stream towers / stock-torch baseline), NCE sweep
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv,noheader
./scripts/probes/halo_desc > gpurun_out/halo_desc.log 2>&1; echo "probe rc=$?"; tail -3 gpurun_out/halo_desc.log
timeout 1500 python -m pytest tests -m gpu -q --no-header -p no:cacheprovider -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
grep -E "^(FAILED|ERROR|[0-9]+ (passed|failed))" gpurun_out/pytest_gpu.log | head -20; tail -5 gpurun_out/pytest_gpu.log
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --dump-launches gpurun_out/launch_table_r2a.txt > gpurun_out/bench_r2a.json 2> gpurun_out/bench_r2a.err; echo "bench rc=$?"; tail -2 gpurun_out/bench_r2a.err; cut -c1-600 gpurun_out/bench_r2a.json
AVID_TOWER_STREAMS=1 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-baseline > gpurun_out/bench_r2a_streams.json 2> gpurun_out/bench_r2a_streams.err; echo "bench streams rc=$?"; tail -2 gpurun_out/bench_r2a_streams.err; cut -c1-300 gpurun_out/bench_r2a_streams.json
timeout 600 python scripts/bench_nce.py --out gpurun_out/nce_sweep_r2a.json > gpurun_out/nce_sweep_r2a.log 2>&1; echo "nce rc=$?"; tail -12 gpurun_out/nce_sweep_r2a.log
