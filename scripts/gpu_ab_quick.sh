# tower / conv / graph tests of the working tree, then the same-box A/B of scripts/gpu_ab_old_new.sh (the old tree = _ab_old/)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_conv_tc_gpu.py tests/test_towers_gpu.py tests/test_graphs_gpu.py tests/test_config2_gpu.py -m gpu -q --no-header -p no:cacheprovider -x > gpurun_out/pytest_quick.log 2>&1; echo "pytest rc=$?"
grep -E "^(FAILED|ERROR)|passed|failed|Error" gpurun_out/pytest_quick.log | head -10
bash scripts/gpu_ab_old_new.sh
