#!/bin/bash
# Round 2: the two file-driven example flows on hardware (small synthetic files written on the box).
mkdir -p gpurun_out /tmp/ex_in /tmp/ex_lc
python - <<'PY'
import numpy as np
rng = np.random.default_rng(1)
for i in range(3):
    a = rng.uniform(0, 1000, (150_000 + 20_000 * i, 3)).astype(np.float32)
    if i == 1: np.save(f"/tmp/ex_in/mock_{i}.npy", a)
    else: np.savetxt(f"/tmp/ex_in/mock_{i}.dat", a, fmt="%.9g")
def sector(n):
    return np.stack([120 + 60 * rng.random(n), -10 + 60 * rng.random(n), 2000 * rng.random(n), 0.7 + 0.4 * rng.random(n), 2e-4 * (0.5 + rng.random(n))], 1).astype(np.float32)
np.savetxt("/tmp/ex_lc/data.dat", sector(100_000), fmt="%.9g")
np.savetxt("/tmp/ex_lc/randoms.dat", sector(500_000), fmt="%.9g")
PY
( time timeout 25 python examples/many_mocks.py --files /tmp/ex_in --out /tmp/ex_out --grid 128 ) > gpurun_out/r2_example_many_mocks.log 2>&1; echo "many_mocks rc=$?"
( time timeout 30 python examples/lightcone_gpu.py --grid 128 --data-file /tmp/ex_lc/data.dat --randoms-file /tmp/ex_lc/randoms.dat --out /tmp/ex_lc_out ) > gpurun_out/r2_example_lightcone.log 2>&1; echo "lightcone rc=$?"
ls -la /tmp/ex_out /tmp/ex_lc_out >> gpurun_out/r2_example_many_mocks.log 2>&1
tail -4 gpurun_out/r2_example_many_mocks.log; tail -12 gpurun_out/r2_example_lightcone.log
