# 2 GPUs: NCCL world-2 tests through the real kernels (sharded bank, sharded CMA, sharded Adam) and a short sharded bench with the in-run parity check
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_sharded_cma.py tests/test_sharded_bank.py tests/test_sharded_adam.py -m gpu -q --no-header -p no:cacheprovider > gpurun_out/pytest_2gpu.log 2>&1; echo "pytest rc=$?"
grep -E "^(FAILED|ERROR)|passed|failed|Error|assert" gpurun_out/pytest_2gpu.log | head -20
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29621 bench.py --gpus 2 --steps 6 --warmup 3 > gpurun_out/bench_2gpu.json 2> gpurun_out/bench_2gpu.err; echo "bench rc=$?"; tail -2 gpurun_out/bench_2gpu.err | cut -c1-300
python - <<PY
import json
d=json.loads(open("gpurun_out/bench_2gpu.json").read().strip().splitlines()[-1])
print({k:d.get(k) for k in ("value","ms_per_step","n_gpus","gpu_launches","parity_check")}, d["e2e"], d.get("config3"), d.get("config4"))
PY
