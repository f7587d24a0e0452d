#!/bin/bash
# Round 2, last call (1 GPU): whole GPU suite, smoke, the bench line, the ncu launch list of the bench command, the
# file-driven batch pipeline at two sizes, a finer L2-prefetch A/B for the y passes.
mkdir -p gpurun_out
LOG=gpurun_out/r2_final2.log; : > $LOG
run() { local name=$1; shift; echo "== $name" | tee -a $LOG; ( time timeout 400 "$@" ) > "gpurun_out/r2f_$name.log" 2>&1; echo "   rc=$?" | tee -a $LOG; tail -3 "gpurun_out/r2f_$name.log" >> $LOG; }
run tests_all python -m pytest tests -q -m gpu -p no:cacheprovider --durations=5
run smoke python -c "import __graft_entry__ as G; G.smoke()"
run bench python bench.py --steps 5 --warmup 3
run ncu_launches ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2f_launches_ncu.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline
run io_bench_512 python benchmarks/batch_files_bench.py --mesh 512 --particles 1e7 --catalogs 6
run io_bench_1024 python benchmarks/batch_files_bench.py --mesh 1024 --particles 1e8 --catalogs 4
run ab_prefetch python benchmarks/ab_options.py --only-set --steps 5 --warmup 3 --set fft_prefetch=37 --set fft_prefetch=74
cat $LOG
