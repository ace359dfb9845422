mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:nce_gather_kernel -s 3 -c 1 -o gpurun_out/r1_nce_k1024 -f python scripts/bench_nce.py --banks 2000000 --negatives 1024 --iters 3 --warmup 2 > gpurun_out/ncu_nce1024.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:nce_gather_kernel -s 3 -c 1 -o gpurun_out/r1_nce_k16384 -f python scripts/bench_nce.py --banks 2000000 --negatives 16384 --iters 3 --warmup 2 > gpurun_out/ncu_nce16384.log 2>&1
ls -la gpurun_out/r1_nce*.ncu-rep
