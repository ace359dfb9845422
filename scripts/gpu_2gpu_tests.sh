mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_sharded_cma.py tests/test_sharded_bank.py -m gpu -q --no-header -p no:cacheprovider > gpurun_out/pytest_2gpu.log 2>&1; echo "pytest rc=$?"
grep -E "^(FAILED|ERROR)|passed|failed|Error|assert" gpurun_out/pytest_2gpu.log | head -20
