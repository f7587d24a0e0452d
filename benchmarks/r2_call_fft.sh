#!/bin/bash
# Round 2, FFT call (1 GPU): parity of the own-FFT paths and the gather staging variants, then the bench line.
mkdir -p gpurun_out
LOG=gpurun_out/r2_fftcall.log; : > $LOG
run() { local name=$1; shift; echo "== $name" | tee -a $LOG; ( time timeout 900 "$@" ) > "gpurun_out/r2_$name.log" 2>&1; echo "   rc=$?" | tee -a $LOG; tail -4 "gpurun_out/r2_$name.log" >> $LOG; }
run fft_tests python -m pytest tests/test_gpu_options.py tests/test_gpu_zzz6_gather_stage.py tests/test_gpu_a_configs.py -q -m gpu -p no:cacheprovider -k "own_fft or tma or staged or all_paths or c1"
run fft_dist_tests python -m pytest tests/test_gpu_dist.py -q -m gpu -p no:cacheprovider -k "slab_z or roundtrip"
run fft_zplain python benchmarks/fft_zplain_probe.py
run fft_bench python bench.py --steps 8 --warmup 3 --no-cpu-baseline
cat $LOG
