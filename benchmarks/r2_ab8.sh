#!/bin/bash
# A/B of the exchange knobs on N GPUs (gpurun --gpus N): one bench line per setting, device-resident value only.
N=${1:-8}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
i=0
for cfg in "${@:2}"; do
  i=$((i+1))
  env $cfg $TR --master-port $((29600+i)) bench.py --gpus $N --steps 5 --warmup 3 --no-e2e > gpurun_out/r2_ab${N}_$i.log 2>&1
  echo "== $cfg"; python benchmarks/summarize_bench.py gpurun_out/r2_ab${N}_$i.log
done
