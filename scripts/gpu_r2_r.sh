mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_sharded_bank.py tests/test_sharded_adam.py tests/test_sharded_cma.py tests/test_config2_gpu.py tests/test_graphs_gpu.py -m gpu -q --no-header -p no:cacheprovider > gpurun_out/pytest_r.log 2>&1; echo "pytest rc=$?"; grep -E "^(FAILED|ERROR)|mismatch" gpurun_out/pytest_r.log | cut -c1-600 | head; tail -3 gpurun_out/pytest_r.log
