# round 2, call D (2 GPUs): symmetric-memory probe, drop-in test, NCE baseline recheck
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 scripts/probes/symm_probe.py > gpurun_out/symm_probe.log 2>&1; echo "symm rc=$?"; grep -v "OMP_NUM\|\*\*\*\*" gpurun_out/symm_probe.log | tail -30
timeout 900 python -m pytest tests/test_dropin.py tests/test_criterion_gpu.py -m gpu -q --no-header -p no:cacheprovider > gpurun_out/pytest_d.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_d.log
timeout 300 python scripts/bench_nce.py --banks 2000000 --iters 20 2>&1 | grep -o '"K": [0-9]*\|"ms_median": [0-9.]*\|"frac_of_measured_hbm": [0-9.]*' | paste - - -
