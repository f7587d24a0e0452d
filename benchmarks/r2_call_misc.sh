#!/bin/bash
# Round 2, miscellaneous measurements on one GPU: config tests after the radial-update fusion, secondary configs,
# the multipole estimator and the catalog kernels under ncu.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_a_configs.py tests/test_gpu_iterative.py -q -m gpu -p no:cacheprovider 2>&1 | tail -4
python benchmarks/secondary.py c1 c2 c2tsc c3 > gpurun_out/r2_secondary_b.log 2>&1; cut -c1-400 gpurun_out/r2_secondary_b.log
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
python /tmp/pk_probe.py > gpurun_out/r2_pk_time.log 2>&1; tail -1 gpurun_out/r2_pk_time.log
ncu --set full --clock-control none -k regex:pk_kernel -c 1 -o gpurun_out/r2_pk_kernel python /tmp/pk_probe.py > /dev/null 2>&1
ncu --set full --clock-control none -k regex:'sky_to_cartesian|cartesian_to_sky' -c 2 -o gpurun_out/r2_catalog python benchmarks/catalog_bench.py > gpurun_out/r2_catalog_bench.log 2>&1; tail -3 gpurun_out/r2_catalog_bench.log
ls -la gpurun_out/*.ncu-rep
