mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_conv_tc_gpu.py -m gpu -q --no-header -p no:cacheprovider -x > gpurun_out/pytest_tc.log 2>&1; echo "pytest rc=$?"
grep -E "^(FAILED|ERROR|[0-9]+ (passed|failed))|assert|Error|error" gpurun_out/pytest_tc.log | head -40
timeout 300 python -m pytest tests/test_conv_tc_gpu.py -m gpu -q --no-header -p no:cacheprovider > gpurun_out/pytest_tc_all.log 2>&1; echo "pytest rc=$?"
grep -E "^(FAILED|ERROR|[0-9]+ (passed|failed))" gpurun_out/pytest_tc_all.log | head -60
