#!/bin/bash
# Everything that was written after round 1's GPU budget was spent, in ONE gpurun call (about 6-8 minutes of box
# time), cheapest and most informative first; every step writes its own log under gpurun_out/ so that a failure
# in one does not hide the others (no `set -e`, no `pytest -x` across files).
#
#   /usr/local/graft/bin/gpurun --timeout 900 -- 'bash benchmarks/first_gpu_call.sh'
#
# Not part of the product; nothing here is a bench value (the ncu passes serialise and flush caches).
mkdir -p gpurun_out
run() { local name=$1; shift; echo "== $name" | tee -a gpurun_out/r2_first.log; ( time timeout 600 "$@" ) > "gpurun_out/r2_$name.log" 2>&1; echo "   rc=$?" | tee -a gpurun_out/r2_first.log; tail -3 "gpurun_out/r2_$name.log" >> gpurun_out/r2_first.log; }

# 1. the GPU tests that have never run, one file at a time
for f in zzz1_golden_catalog zzz2_fullsize zzz3_tsc_slabs_deterministic zzz4_pk zzz5_batch zzz6_gather_stage; do
  run "test_$f" python -m pytest "tests/test_gpu_$f.py" -q -m gpu
done
# 2. everything that had passed before (regression: mas.o's TSC kernels changed for the slab layout)
run test_validated python -m pytest tests -q -m gpu -x --ignore-glob='tests/test_gpu_zzz[1-9]_*'
# 3. the bench with the batched host pipeline next to the one-at-a-time e2e
run bench_batch python bench.py --steps 5 --warmup 3 --e2e-batch 4
# 3b. the secondary configs, TSC included (BASELINE configs[1] names TSC; only CIC was measured in round 1)
run secondary python benchmarks/secondary.py c2 c2tsc c3
# 4. option A/B on the bench workload, uniform and lognormal
run ab_uniform python benchmarks/ab_options.py --steps 3 --set scatter_pairs=2
run ab_lognormal python benchmarks/ab_options.py --steps 3 --catalog lognormal
# 5. ncu: launch list of one multipole estimate + catalog kernels, full capture of pk_kernel and the two conversions
cat > /tmp/pk_probe.py <<'PY'
import sys, numpy as np, torch
sys.path.insert(0, '.')
import __graft_entry__ as G
B = G.load_package()
n, L = 1024, 2500.0
rho = torch.rand((n, n, n), device='cuda') + 0.5
for _ in range(3):
    r = B.power_multipoles(rho, np.full(3, L, np.float32), dk=0.005, nbins=256, mas='cic')
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
ctx = B.Context.get(0); ctx.profile(True); e0.record()
for _ in range(5):
    r = B.power_multipoles(rho, np.full(3, L, np.float32), dk=0.005, nbins=256, mas='cic')
e1.record(); torch.cuda.synchronize()
print('power_multipoles 1024^3: %.3f ms per call' % (e0.elapsed_time(e1) / 5), {k: round(v[0] / v[1], 3) for k, v in ctx.profile_read().items()})
PY
run pk_time python /tmp/pk_probe.py
run ncu_pk ncu --set full --clock-control none --import-source on -k regex:pk_kernel -c 2 -o gpurun_out/r2_pk_kernel python /tmp/pk_probe.py
run ncu_catalog ncu --set full --clock-control none --import-source on -k regex:'sky_to_cartesian|cartesian_to_sky' -c 4 -o gpurun_out/r2_catalog python benchmarks/catalog_bench.py
cat gpurun_out/r2_first.log
