# round 2, call B (2 GPUs): criterion tests incl. the NCCL world-2 ones, NCE sweep (cp.async ring), N=1 vs N=2 scaling diagnostics
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --no-header -p no:cacheprovider > gpurun_out/pytest_gpu2.log 2>&1; echo "pytest rc=$?"
grep -E "^(FAILED|ERROR|[0-9]+ (passed|failed))" gpurun_out/pytest_gpu2.log | head -20; tail -3 gpurun_out/pytest_gpu2.log
timeout 600 python scripts/bench_nce.py --banks 2000000 --out gpurun_out/nce_sweep_r2b.json > gpurun_out/nce_sweep_r2b.log 2>&1; echo "nce rc=$?"; grep -o '"K": [0-9]*\|"ms_median": [0-9.]*\|"frac_of_measured_hbm": [0-9.]*' gpurun_out/nce_sweep_r2b.log | paste - - -
run() { name=$1; shift; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 "$@" > gpurun_out/bench2_$name.json 2> gpurun_out/bench2_$name.err; echo "bench2 $name rc=$?"; tail -2 gpurun_out/bench2_$name.err; python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/bench2_$name.json').read().strip().splitlines()[-1])
    print('$name', round(d['value'],1), 'ms', round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value'],1), d.get('parity_check'), {k:d[k] for k in ('config3','config4') if k in d})
except Exception as e: print('parse', e)
PY
}
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-baseline > gpurun_out/bench1_r2b.json 2> gpurun_out/bench1_r2b.err; python -c "
import json; d=json.loads(open('gpurun_out/bench1_r2b.json').read()); print('N=1', round(d['value'],1), round(d['ms_per_step'],2))"
run default

AVID_BENCH_NO_DDP=1 run noddp_real --no-subrecords
run replicated --bank-mode replicated --no-subrecords
