# last session: N = 8 line of the final code alone (the budget left did not allow the N = 1 run on the same 8-GPU box)
mkdir -p gpurun_out
timeout 95 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/scale_n8_s3.json 2> gpurun_out/scale_n8_s3.err; echo "N=8 rc=$?"
python - <<'PY'
import json
b=json.loads(open('gpurun_out/scale_n8_s3.json').read().strip().splitlines()[-1])
print('N=8', round(b['value'],1), round(b['ms_per_step'],2), 'e2e', round(b['e2e']['value'],1), b['config'].get('grad_sync'))
print('parity', b.get('parity_check'))
for k in ('config3','config4'): print(k, {x:(round(y,3) if isinstance(y,float) else y) for x,y in b.get(k,{}).items() if x!='workload'})
PY
