#!/bin/bash
# Multi-GPU call (run with gpurun --gpus N): parity of the slab path with N ranks, then the bench line.
#   /usr/local/graft/bin/gpurun --gpus 2 --timeout 900 -- 'bash benchmarks/r2_multi.sh 2'
N=${1:-2}
mkdir -p gpurun_out
LOG=gpurun_out/r2_multi_${N}.log; : > $LOG
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
run() { local name=$1; shift; echo "== $name" | tee -a $LOG; ( time timeout 600 "$@" ) > "gpurun_out/r2_${name}_${N}gpu.log" 2>&1; echo "   rc=$?" | tee -a $LOG; grep -E "OK|FAIL|PASS|^\{|rror" "gpurun_out/r2_${name}_${N}gpu.log" | tail -40 >> $LOG; }
nvidia-smi topo -m >> $LOG 2>&1
if [ -z "$SKIP_CHECK" ]; then MGC_TSC=1 run check $TR --master-port 29511 tests/multi_gpu_check.py; fi
run bench $TR --master-port 29512 bench.py --gpus $N --steps 5 --warmup 3
if [ -n "$AB_NCCL" ]; then BAOREC_EXCHANGE=nccl run bench_nccl $TR --master-port 29513 bench.py --gpus $N --steps 5 --warmup 3 --no-e2e; fi
if [ -n "$REF_ARM" ]; then run bench_ref $TR --master-port 29514 bench.py --impl reference --gpus $N --steps 1 --warmup 1; fi
cat $LOG
