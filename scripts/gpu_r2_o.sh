mkdir -p gpurun_out
python scripts/probes/stem_time.py 10
ncu --set full --clock-control none --import-source on -k regex:stem_forward_kernel -s 2 -c 1 -o gpurun_out/r2_stem_fwd python scripts/probes/stem_time.py 1 > /dev/null 2>&1; echo "ncu fwd rc=$?"
ncu --set full --clock-control none --import-source on -k regex:stem_wgrad_kernel -s 2 -c 1 -o gpurun_out/r2_stem_wgrad python scripts/probes/stem_time.py 1 > /dev/null 2>&1; echo "ncu wgrad rc=$?"
ls -la gpurun_out/r2_stem_*.ncu-rep
