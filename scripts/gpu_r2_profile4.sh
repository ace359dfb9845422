# round 2, last session: tests of the conv kernels, quick bench, steady-state launch list of the final kernels, --set full capture of the
# strided conv3x-entry input gradient (the 72nd conv_tc_kernel launch of a step)
mkdir -p gpurun_out
bash scripts/gpu_quick.sh 2>&1 | grep -v Warning | tail -14
export AVID_CUDA_GRAPH=0 AVID_TOWER_STREAMS=0
B="python bench.py --steps 2 --warmup 3 --skip-e2e --no-cpu-baseline --no-gpu-baseline"
ncu --nvtx --nvtx-include "avid_timed" --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,l1tex__m_xbar2l1tex_read_bytes.sum --clock-control none --csv --log-file gpurun_out/r2_launches_s3.csv $B > gpurun_out/r2_bench_under_ncu_s3.log 2>&1
echo "launch list rc=$? lines $(wc -l < gpurun_out/r2_launches_s3.csv)"
ncu --nvtx --nvtx-include "avid_timed" --set full --clock-control none --import-source on -k regex:conv_tc_kernel -s 71 -c 1 -o gpurun_out/r2_conv_tc_strided -f $B > gpurun_out/r2_ncu_conv_tc_strided.log 2>&1; echo "strided rc=$?"
ls -la gpurun_out/r2_conv_tc_strided.ncu-rep
