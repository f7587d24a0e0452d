#!/bin/bash
# Round 2, final call (1 GPU): whole GPU suite, smoke, the bench line and the ncu launch list of the bench command.
mkdir -p gpurun_out
LOG=gpurun_out/r2_final.log; : > $LOG
run() { local name=$1; shift; echo "== $name" | tee -a $LOG; ( time timeout 900 "$@" ) > "gpurun_out/r2_$name.log" 2>&1; echo "   rc=$?" | tee -a $LOG; tail -3 "gpurun_out/r2_$name.log" >> $LOG; }
run tests_all python -m pytest tests -q -m gpu -p no:cacheprovider --durations=5
run smoke python -c "import __graft_entry__ as G; G.smoke()"
run bench python bench.py --steps 5 --warmup 3
run ncu_launches ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_ncu.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline
run secondary python benchmarks/secondary.py c1 c2 c2tsc c3
cat $LOG
