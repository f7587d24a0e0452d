#!/bin/bash
# Round 2, sort call (1 GPU): four particles per thread in usort_reorder / unsort -- parity tests that cover the sort, then the bench workload.
mkdir -p gpurun_out
LOG=gpurun_out/r2_sortcall.log; : > $LOG
run() { local name=$1; shift; echo "== $name" | tee -a $LOG; ( time timeout 900 "$@" ) > "gpurun_out/r2_$name.log" 2>&1; echo "   rc=$?" | tee -a $LOG; tail -4 "gpurun_out/r2_$name.log" >> $LOG; }
run sort_tests python -m pytest tests/test_gpu_options.py tests/test_gpu_zzz6_gather_stage.py tests/test_gpu_mas.py tests/test_gpu_iterative.py -q -m gpu -p no:cacheprovider
run sort_ab python benchmarks/ab_options.py --only-set --steps 8
run sort_ab_lognormal python benchmarks/ab_options.py --only-set --steps 8 --catalog lognormal
cat $LOG
