#!/bin/bash
# Round 2 profile call (1 GPU): whole GPU suite, the bench line, the ncu launch list of the bench command and full
# captures of the dominant hand-written kernels.  Numbers printed under ncu are never bench values.
mkdir -p gpurun_out
LOG=gpurun_out/r2_profile.log; : > $LOG
run() { local name=$1; shift; echo "== $name" | tee -a $LOG; ( time timeout 900 "$@" ) > "gpurun_out/r2_$name.log" 2>&1; echo "   rc=$?" | tee -a $LOG; tail -3 "gpurun_out/r2_$name.log" >> $LOG; }
run tests_all python -m pytest tests -q -m gpu -p no:cacheprovider --durations=8
run smoke python -c "import __graft_entry__ as G; G.smoke()"
run bench python bench.py --steps 5 --warmup 3
run ncu_launches ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_ncu.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline
run ncu_full ncu --set full --clock-control none --import-source on -k regex:"fft_z_disp|fft_z_solve|gather_tile|scatter_records|usort_reorder|usort_count|unsort|fft_cols" -s 60 -c 12 -o gpurun_out/r2_own_kernels python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline
cat $LOG
