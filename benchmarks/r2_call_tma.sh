#!/bin/bash
# Round 2, TMA call (1 GPU): parity of the tensor-map staged tile gather against the cp.async staging, then the A/B at
# the bench workload (uniform and lognormal) and one full ncu capture of both gather kernels.
mkdir -p gpurun_out
LOG=gpurun_out/r2_tma.log; : > $LOG
run() { local name=$1; shift; echo "== $name" | tee -a $LOG; ( time timeout 600 "$@" ) > "gpurun_out/r2_$name.log" 2>&1; echo "   rc=$?" | tee -a $LOG; tail -4 "gpurun_out/r2_$name.log" >> $LOG; }
run tma_tests python -m pytest tests/test_gpu_zzz6_gather_stage.py -q -m gpu -p no:cacheprovider -k "tma or staged"
run tma_ab python benchmarks/ab_options.py --only-set --set gather_stage=2 --steps 8
run tma_ab_lognormal python benchmarks/ab_options.py --only-set --set gather_stage=2 --steps 8 --catalog lognormal
run tma_ncu ncu --set full --clock-control none --import-source on -k regex:"gather_tile" -c 4 -o gpurun_out/r2_gather_tma python benchmarks/ab_options.py --only-set --set gather_stage=2 --steps 1 --warmup 1
cat $LOG
