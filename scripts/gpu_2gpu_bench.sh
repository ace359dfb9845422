mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29621 bench.py --gpus 2 --steps 6 --warmup 3 --skip-e2e > gpurun_out/bench_2gpu.json 2> gpurun_out/bench_2gpu.err; echo "bench rc=$?"; tail -2 gpurun_out/bench_2gpu.err | cut -c1-300
python - <<PY
import json
d=json.loads(open("gpurun_out/bench_2gpu.json").read().strip().splitlines()[-1])
print({k:d[k] for k in ("value","ms_per_step","n_gpus","gpu_launches","last_loss")}, d["config"]["bank_layout"])
PY
