mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_sharded_adam.py -m gpu -q --no-header -p no:cacheprovider > gpurun_out/pytest_g.log 2>&1; echo "pytest rc=$?"; grep -n "mismatch\|passed\|failed" gpurun_out/pytest_g.log | cut -c1-3000 | head -8
timeout 600 python -m pytest tests/test_criterion_gpu.py tests/test_criterion_edges_gpu.py tests/test_warm_start_gpu.py tests/test_sharded_bank.py -m gpu -q --no-header -p no:cacheprovider > gpurun_out/pytest_g2.log 2>&1; echo "pytest nce-tma rc=$?"; tail -5 gpurun_out/pytest_g2.log
for tma in 1 0; do echo "== AVID_NCE_TMA=$tma"; AVID_NCE_TMA=$tma timeout 300 python scripts/bench_nce.py --banks 2000000 --iters 20 2>&1 | grep -o '"K": [0-9]*\|"ms_median": [0-9.]*\|"frac_of_measured_hbm": [0-9.]*\|Error.*' | paste - - -; done
