# round 2: ONE 8-GPU box: N=1 reference, then N=8 with the fused gradient exchange (parity_check + config-3 / config-4 sub-records)
mkdir -p gpurun_out
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-gpu-baseline > gpurun_out/scale_n1.json 2> gpurun_out/scale_n1.err; echo "N=1 rc=$?"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 8 --steps 20 --warmup 3 > gpurun_out/scale_n8.json 2> gpurun_out/scale_n8.err; echo "N=8 rc=$?"; grep -v "OMP_NUM\|\*\*\*\*\|Warning\|warn\|last_loss\|^$\|run_backward" gpurun_out/scale_n8.err | tail -8
python - <<'PY'
import json
a=json.loads(open('gpurun_out/scale_n1.json').read()); b=json.loads(open('gpurun_out/scale_n8.json').read().strip().splitlines()[-1])
print('N=1', round(a['value'],1), round(a['ms_per_step'],2), 'e2e', round(a['e2e']['value'],1))
print('N=8', round(b['value'],1), round(b['ms_per_step'],2), 'e2e', round(b['e2e']['value'],1), 'eff', round(b['value']/8/a['value'],3), 'e2e eff', round(b['e2e']['value']/8/a['e2e']['value'],3), b['config']['grad_sync'])
print('parity', b.get('parity_check'))
for k in ('config3','config4'): print(k, {x:(round(y,3) if isinstance(y,float) else y) for x,y in b.get(k,{}).items() if x!='workload'})
PY
