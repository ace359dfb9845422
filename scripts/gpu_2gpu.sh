mkdir -p gpurun_out
nvidia-smi -L | head -3
timeout 600 python -m pytest tests/test_sharded_bank.py -m gpu -q --no-header -p no:cacheprovider -x > gpurun_out/pytest_2gpu.log 2>&1; echo "pytest rc=$?"
grep -E "^(FAILED|ERROR)|passed|failed|assert|Error|error" gpurun_out/pytest_2gpu.log | head -20
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_2gpu.json 2> gpurun_out/bench_2gpu.err; echo "bench rc=$?"; tail -3 gpurun_out/bench_2gpu.err
python - <<PY
import json
d=json.loads(open("gpurun_out/bench_2gpu.json").read().strip().splitlines()[-1])
print({k:d[k] for k in ("value","ms_per_step","n_gpus","gpu_launches","last_loss","clocks")}, d["e2e"]["value"], d["config"]["bank_layout"])
PY
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29612 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 2>/dev/null | cut -c1-400
