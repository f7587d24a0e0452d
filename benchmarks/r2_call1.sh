#!/bin/bash
# Round 2, first gpurun call: the whole GPU suite (no -x, per-file logs), then the measurements that had a flag but no
# number.  Every step writes its own log under gpurun_out/; nothing here is a bench value under ncu.
#   /usr/local/graft/bin/gpurun --timeout 1500 -- 'bash benchmarks/r2_call1.sh'
mkdir -p gpurun_out
LOG=gpurun_out/r2_call1.log; : > $LOG
run() { local name=$1; shift; echo "== $name" | tee -a $LOG; ( time timeout 900 "$@" ) > "gpurun_out/r2_$name.log" 2>&1; echo "   rc=$?" | tee -a $LOG; tail -4 "gpurun_out/r2_$name.log" >> $LOG; }
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv >> $LOG
nproc >> $LOG
run smoke python -c "import __graft_entry__ as G; G.smoke()"
run tests_all python -m pytest tests -q -m gpu -p no:cacheprovider --durations=15
run bench python bench.py --steps 5 --warmup 3 --e2e-batch 4
run ab_uniform python benchmarks/ab_options.py --steps 3 --set scatter_pairs=2
run secondary python benchmarks/secondary.py c1 c2 c2tsc c3
run bench_lognormal python bench.py --steps 5 --warmup 3 --catalog lognormal --no-cpu-baseline
cat $LOG
