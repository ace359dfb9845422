mkdir -p gpurun_out
./scripts/probes/gather4_tma > gpurun_out/gather4.log 2>&1; echo "gather4 rc=$?"; cat gpurun_out/gather4.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 scripts/probes/symm_probe.py > gpurun_out/symm_probe.log 2>&1; echo "symm rc=$?"; grep -v "OMP_NUM\|\*\*\*\*" gpurun_out/symm_probe.log | tail -30
