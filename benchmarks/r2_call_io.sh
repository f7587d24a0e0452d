#!/bin/bash
# Round 2, catalog-file pipeline on hardware (1 GPU): the batch tests (array and file sources), the file-driven batch
# at two sizes, and the L2-prefetch A/B of the own FFT kernels.
mkdir -p gpurun_out
LOG=gpurun_out/r2_io.log; : > $LOG
run() { local name=$1; shift; echo "== $name" | tee -a $LOG; ( time timeout 150 "$@" ) > "gpurun_out/r2_$name.log" 2>&1; echo "   rc=$?" | tee -a $LOG; tail -4 "gpurun_out/r2_$name.log" >> $LOG; }
run io_tests python -m pytest tests/test_gpu_zzz5_batch.py -q -p no:cacheprovider
run io_bench_512 python benchmarks/batch_files_bench.py --mesh 512 --particles 1e7 --catalogs 6
run io_bench_1024 python benchmarks/batch_files_bench.py --mesh 1024 --particles 1e8 --catalogs 4
run ab_prefetch python benchmarks/ab_options.py --only-set --steps 5 --warmup 3 --set fft_prefetch=148 --set fft_prefetch=296 --set fft_prefetch=592
cat $LOG
