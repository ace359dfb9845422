mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_train_gpu.py tests/test_towers_gpu.py -m gpu -q --no-header -p no:cacheprovider -x > gpurun_out/pytest_glue.log 2>&1; echo "pytest rc=$?"
grep -E "^(FAILED|ERROR)|passed|failed|assert|Error|error" gpurun_out/pytest_glue.log | head -30
