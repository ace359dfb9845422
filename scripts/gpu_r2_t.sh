mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 4 --steps 10 --warmup 3 > gpurun_out/scale_n4.json 2> gpurun_out/scale_n4.err; echo "N=4 rc=$?"; grep -v "OMP_NUM\|\*\*\*\*\|Warning\|warn\|last_loss\|^$\|run_backward\|Consider using" gpurun_out/scale_n4.err | tail -8
python - <<'PY'
import json
b=json.loads(open('gpurun_out/scale_n4.json').read().strip().splitlines()[-1])
print('N=4', round(b['value'],1), round(b['ms_per_step'],2), 'e2e', round(b['e2e']['value'],1))
for k in ('config3','config4'): print(k, {x:(round(y,3) if isinstance(y,float) else y) for x,y in b.get(k,{}).items() if x not in ('workload','bank_init')})
PY
