# round 2, call F (2 GPUs): ShardedAdam test, N=2 bench fused vs ddp
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_sharded_adam.py tests/test_sharded_bank.py -m gpu -q --no-header -p no:cacheprovider > gpurun_out/pytest_f.log 2>&1; echo "pytest rc=$?"; tail -25 gpurun_out/pytest_f.log
run() { name=$1; shift; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 "$@" > gpurun_out/bench2_$name.json 2> gpurun_out/bench2_$name.err; echo "bench2 $name rc=$?"; grep -v "OMP_NUM\|\*\*\*\*\|UserWarning\|last_loss\|^$" gpurun_out/bench2_$name.err | tail -12; python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/bench2_$name.json').read().strip().splitlines()[-1])
    print('$name', round(d['value'],1), 'ms', round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value'],1), d['config']['grad_sync'], 'loss', d['last_loss'])
except Exception as e: print('parse', e)
PY
}
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-baseline > gpurun_out/bench1_r2f.json 2> gpurun_out/bench1_r2f.err; python -c "
import json; d=json.loads(open('gpurun_out/bench1_r2f.json').read()); print('N=1', round(d['value'],1), round(d['ms_per_step'],2), d['last_loss'])"
run fused --no-subrecords
run ddp --no-subrecords --grad-sync ddp
AVID_BENCH_NO_DDP=1 run nosync --no-subrecords
