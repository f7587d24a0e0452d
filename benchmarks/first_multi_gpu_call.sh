#!/bin/bash
# The multi-GPU counterpart of first_gpu_call.sh: what changed on the distributed path since it last ran on hardware
# (TSC on slabs; the __grid_constant__ peer table of the slab transposes; vector reductions on the z-binned scatter).
#
#   /usr/local/graft/bin/gpurun --gpus 2 --timeout 600 -- 'bash benchmarks/first_multi_gpu_call.sh 2'
#
# Charged N x the box time: keep it to the parity check and one bench line per setting.
N=${1:-2}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
( time MGC_TSC=1 timeout 500 $TR tests/multi_gpu_check.py ) > gpurun_out/r2_mgc_${N}gpu.log 2>&1
echo "multi_gpu_check rc=$?"; grep -E "MULTI_GPU_CHECK|FAIL|TSC" gpurun_out/r2_mgc_${N}gpu.log | tail -12
( time timeout 400 $TR bench.py --gpus $N --steps 5 --warmup 3 ) > gpurun_out/r2_bench_${N}gpu.log 2>&1
echo "bench rc=$?"; tail -1 gpurun_out/r2_bench_${N}gpu.log | cut -c1-400
