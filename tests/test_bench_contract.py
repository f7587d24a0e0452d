"""CPU: the parts of bench.py's contract that need no GPU -- the reference arm (`--impl reference`) prints one
JSON line with the agreed keys and times the CPU oracle on a bounded sample; ranks other than 0 print nothing;
the algorithmic-bytes table knows every kernel name the engine profiles."""
import json
import os
import re
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent


def run_ref(extra_env=None):
    env = dict(os.environ, **(extra_env or {}))
    r = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                        "--ref-mesh", "32", "--particles", "3.2e6"], capture_output=True, text=True, env=env, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    return [l for l in r.stdout.splitlines() if l.strip()]


def test_reference_arm_prints_one_valid_line():
    lines = run_ref()
    assert len(lines) == 1
    d = json.loads(lines[0])
    base = json.loads((ROOT / "BASELINE.json").read_text())
    assert d["impl"] == "reference" and d["unit"] == "ms" and d["higher_is_better"] is False
    assert d["metric"].split("(")[0].strip() == base["metric"].split("(")[0].strip() == "ms per reconstruction"
    for k in ("value", "n_gpus", "steps", "warmup", "ms_per_step", "scaling", "vs_baseline", "dtype", "data", "config",
              "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["vs_baseline"] is None and d["data"] == "synthetic" and "workload" in d["config"]
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and "sample" in cb
    assert d["e2e"] == {"value": d["value"], "unit": "ms", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    # a 32^3 sample scaled x(1024/32)^3 to the 1024^3 workload
    assert abs(d["value"] / d["ms_per_step"] - 32768) < 1e-6 and d["ms_per_step"] > 0
    assert "C/OpenMP" in cb["sample"] or "numpy port" in cb["sample"]


def test_reference_arm_is_silent_on_other_ranks():
    assert run_ref({"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"}) == []


def test_algorithmic_bytes_cover_the_profiled_kernels():
    sys.path.insert(0, str(ROOT))
    import bench
    M, Mc, N = 1024 ** 3, 513 * 1024 * 1024, 10 ** 8
    # kernel names as the engine's per-launch profiling reports them in the bench workload (profiles/r1_launches_ncu_final_summary.csv)
    for name in ("usort_count_kernel", "usort_reorder_kernel", "scatter_records_kernel", "hash_positions_kernel",
                 "gather_tile_kernel<3>", "unsort_kernel", "kspace_kernel<FusedLosOp>", "kspace_kernel<DispOp>", "cufft_r2c", "cufft_c2r"):
        b = bench.algorithmic_bytes(name, M, Mc, N)
        assert b is not None and b > 0, name
    assert bench.algorithmic_bytes("gather_tile_kernel<3>", M, Mc, N) == 32 * N + 12 * M
    assert bench.algorithmic_bytes("scatter_records_kernel", M, Mc, N) == 16 * N + 4 * M
    src = (ROOT / "bench.py").read_text()
    for key in ('"roofline"', '"cpu_baseline"', '"e2e"', '"clocks"', '"gpu_launches"', '"h2d_bytes_per_step"', '"d2h_bytes_per_step"'):
        assert key in src, key
    assert re.search(r"--warmup.*default=3", src) and "torch.cuda.Event" in src


def test_the_catalog_is_the_same_at_every_rank_count():
    """bench.py --gpus N: rank r holds the r-th contiguous part of the catalog --gpus 1 draws (strong scaling of one workload)."""
    sys.path.insert(0, str(ROOT))
    import numpy as np
    import bench
    n = (1 << 24) + 12345            # spans two generation chunks
    (fx, fy, fz), fw = bench.make_catalog(n, 2500.0, seed=42)
    for world in (2, 8):
        for rank in (0, world - 1, world // 2):
            lo, hi = rank * n // world, (rank + 1) * n // world
            (x, y, z), w = bench.make_catalog(n, 2500.0, seed=42, share=(lo, hi))
            assert len(w) == hi - lo
            for a, b in ((x, fx), (y, fy), (z, fz)):
                assert np.array_equal(a.numpy(), b.numpy()[lo:hi])


def test_reference_arm_restores_the_host_threads_under_torchrun():
    """torchrun exports OMP_NUM_THREADS=1; scipy's FFT sizes its pool from it.  The CPU arm sets the thread count itself."""
    src = (ROOT / "bench.py").read_text()
    assert 'os.environ["OMP_NUM_THREADS"] = ncpu' in src and 'os.environ["DUCC0_NUM_THREADS"] = ncpu' in src
    assert len(run_ref({"OMP_NUM_THREADS": "1"})) == 1
