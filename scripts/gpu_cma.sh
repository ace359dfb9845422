mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_criterion_gpu.py -m gpu -q --no-header -p no:cacheprovider -x -k "cma" > gpurun_out/pytest_cma.log 2>&1; echo "pytest rc=$?"
grep -E "^(FAILED|ERROR)|passed|failed|assert|Error|error" gpurun_out/pytest_cma.log | head -20
timeout 300 python scripts/bench_cma.py --out gpurun_out/cma_240k.json 2>&1 | tail -3
