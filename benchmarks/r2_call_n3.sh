#!/bin/bash
# Round 2, N3 call (1 GPU): PCS + interlacing + compute_auto_box parity, then the estimator's timing at the reference
# helpers' configuration and the full GPU suite.
mkdir -p gpurun_out
LOG=gpurun_out/r2_n3.log; : > $LOG
run() { local name=$1; shift; echo "== $name" | tee -a $LOG; ( time timeout 1200 "$@" ) > "gpurun_out/r2_$name.log" 2>&1; echo "   rc=$?" | tee -a $LOG; tail -4 "gpurun_out/r2_$name.log" >> $LOG; }
run n3_tests python -m pytest tests/test_gpu_mas.py tests/test_gpu_zzz4_pk.py tests/test_gpu_zzz3_tsc_slabs_deterministic.py -q -m gpu -p no:cacheprovider
run n3_pk_bench python benchmarks/pk_bench.py
run n3_tests_all python -m pytest tests -q -m gpu -p no:cacheprovider --durations=8
cat $LOG
