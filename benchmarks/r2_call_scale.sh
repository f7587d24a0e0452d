#!/bin/bash
# Round 2 scaling call (gpurun --gpus N, N = 4 or 8): the bench line at N GPUs, IterativeRecon at 2048^3 / 8e8 particles
# (the north star's scaling target: efficiency 4 -> 8), and at 8 GPUs MultigridRecon at 2048^3 / 1e9 (BASELINE configs[4])
# and the lognormal catalog.
N=${1:-8}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
( time timeout 400 $TR --master-port 29701 bench.py --gpus $N --steps 5 --warmup 3 ) > gpurun_out/r2_bench_${N}gpu.log 2>&1; echo "bench rc=$?"
( time timeout 500 $TR --master-port 29702 benchmarks/c5_multigrid_dist.py --algorithm iterative --mesh 2048 --particles 8e8 --steps 3 --warmup 2 ) > gpurun_out/r2_iter2048_${N}gpu.log 2>&1; echo "iter2048 rc=$?"
if [ "$N" = "8" ]; then
  ( time timeout 500 $TR --master-port 29703 benchmarks/c5_multigrid_dist.py --mesh 2048 --particles 1e9 --steps 2 --warmup 1 ) > gpurun_out/r2_c5_8gpu.log 2>&1; echo "c5 rc=$?"
  ( time timeout 400 $TR --master-port 29704 bench.py --gpus $N --steps 5 --warmup 3 --catalog lognormal --no-e2e ) > gpurun_out/r2_bench_lognormal_${N}gpu.log 2>&1; echo "lognormal rc=$?"
fi
python benchmarks/summarize_bench.py gpurun_out/r2_bench_${N}gpu.log gpurun_out/r2_bench_lognormal_${N}gpu.log 2>/dev/null | cut -c1-1500
for f in gpurun_out/r2_iter2048_${N}gpu.log gpurun_out/r2_c5_8gpu.log; do [ -f $f ] && grep -E '^\{|rror|real' $f | cut -c1-1800; done
