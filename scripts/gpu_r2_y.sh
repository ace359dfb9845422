mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_conv_tc_gpu.py -m gpu -q --no-header -p no:cacheprovider -x > gpurun_out/pytest_y.log 2>&1; echo "pytest conv rc=$?"; grep -E "^(FAILED|ERROR)|Error|assert " gpurun_out/pytest_y.log | head -10; tail -3 gpurun_out/pytest_y.log
