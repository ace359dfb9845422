# round 2, call C: NCE (register loads + L2 prefetch) correctness + split sweep; CUDA-graph launch-gap experiment
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_criterion_gpu.py tests/test_criterion_edges_gpu.py tests/test_warm_start_gpu.py -m gpu -q --no-header -p no:cacheprovider > gpurun_out/pytest_c.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_c.log
for kc in 0 128 192 256 384; do
  echo "== AVID_NCE_KC=$kc"
  AVID_NCE_KC=$kc timeout 300 python scripts/bench_nce.py --banks 2000000 --negatives 256 1024 4096 16384 --iters 20 2>&1 | grep -o '"K": [0-9]*\|"ms_median": [0-9.]*\|"frac_of_measured_hbm": [0-9.]*' | paste - - -
done
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-baseline --graph-experiment > gpurun_out/bench_graph.json 2> gpurun_out/bench_graph.err; echo "bench rc=$?"; tail -3 gpurun_out/bench_graph.err
python -c "
import json; d=json.loads(open('gpurun_out/bench_graph.json').read()); print('eager', round(d['value'],1), round(d['ms_per_step'],2), 'graph', d.get('graph_experiment'))"
AVID_TOWER_STREAMS=1 timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-baseline --graph-experiment > gpurun_out/bench_graph_s.json 2> gpurun_out/bench_graph_s.err; echo "bench rc=$?"; tail -3 gpurun_out/bench_graph_s.err
python -c "
import json; d=json.loads(open('gpurun_out/bench_graph_s.json').read()); print('streams eager', round(d['value'],1), round(d['ms_per_step'],2), 'graph', d.get('graph_experiment'))"
