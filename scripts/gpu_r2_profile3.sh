# round 2, final kernels: steady-state launch list (NVTX range around the timed steps, eager launches on one stream so that every kernel
# is a separate ncu record) + --set full captures of the dominant kernels.  Numbers printed under ncu are never bench values.
mkdir -p gpurun_out
export AVID_CUDA_GRAPH=0 AVID_TOWER_STREAMS=0
B="python bench.py --steps 2 --warmup 3 --skip-e2e --no-cpu-baseline --no-gpu-baseline"
ncu --nvtx --nvtx-include "avid_timed" --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,l1tex__m_xbar2l1tex_read_bytes.sum --clock-control none --csv --log-file gpurun_out/r2_launches.csv $B > gpurun_out/r2_bench_under_ncu.log 2>&1
echo "launch list rc=$? lines $(wc -l < gpurun_out/r2_launches.csv)"
for k in conv_pair_kernel:2 conv_tc_kernel:3 wgrad_tc_kernel:2 stem_forward_kernel:1 stem_wgrad_kernel:1; do
  name=${k%%:*}; cnt=${k##*:}
  ncu --nvtx --nvtx-include "avid_timed" --set full --clock-control none --import-source on -k regex:$name -c $cnt -o gpurun_out/r2_$name -f $B > gpurun_out/r2_ncu_$name.log 2>&1; echo "$name rc=$?"
done
ncu --nvtx --nvtx-include "avid_timed" --set full --clock-control none --import-source on -k regex:bn_relu_backward_apply_kernel -c 1 -s 30 -o gpurun_out/r2_bn_bwd -f $B > gpurun_out/r2_ncu_bn.log 2>&1; echo "bn rc=$?"
ls -la gpurun_out/r2_*.ncu-rep
