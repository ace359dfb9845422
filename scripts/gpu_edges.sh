mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_spectrogram_gpu.py tests/test_abi.py -m gpu -q --no-header -p no:cacheprovider > gpurun_out/pytest_edges.log 2>&1; echo "pytest rc=$?"
grep -E "^(FAILED|ERROR)|passed|failed|Error|assert" gpurun_out/pytest_edges.log | head -30
