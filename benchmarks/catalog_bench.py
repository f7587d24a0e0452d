#!/usr/bin/env python
"""Catalog pre/post-processing kernels (csrc/catalog.cu) at the bench workload's catalog size:
ms per call (CUDA events, inputs resident in HBM, catalogs larger than L2) and algorithmic GB/s
(12 B in + 12 B out per particle for the conversions, 4 + 4 for the weights, 12 + 12 for the re-wrap)
against the measured copy peak.  Prints one JSON line.   python benchmarks/catalog_bench.py [N]"""
import json
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import __graft_entry__ as G  # noqa: E402

B = G.load_package()


def timed(fn, reps=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def main():
    N = int(float(sys.argv[1])) if len(sys.argv) > 1 else 100_000_000
    peak = 6532.5
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        peak = float(json.loads(p.read_text())["hbm_gbs"])
    g = torch.Generator(device="cuda").manual_seed(42)
    ra = 360 * torch.rand(N, device="cuda", generator=g)
    dec = 180 * torch.rand(N, device="cuda", generator=g) - 90
    red = 0.05 + 2.9 * torch.rand(N, device="cuda", generator=g)
    nz = 1e-3 * torch.rand(N, device="cuda", generator=g)
    cosmo = B.Cosmology(z_tab_max=3)
    x, y, z = B.sky_to_cartesian(ra, dec, red, cosmo)
    ctx = B.Context.get(0)
    ctx.profile(True)
    ms = {
        "sky_to_cartesian": timed(lambda: B.sky_to_cartesian(ra, dec, red, cosmo)),
        "cartesian_to_sky": timed(lambda: B.cartesian_to_sky(x, y, z, cosmo)),
        "fkp_weights": timed(lambda: B.fkp_weights(nz, 5e3)),
        "wrap_positions": timed(lambda: B.wrap_positions(x, y, z, (4000.0,) * 3, (-2000.0,) * 3)),
    }
    prof = ctx.profile_read()
    ctx.profile(False)
    nbytes = {"sky_to_cartesian": 24 * N, "cartesian_to_sky": 24 * N, "fkp_weights": 8 * N, "wrap_positions": 24 * N}
    out = {"workload": f"catalog kernels, N = {N:.0e} particles, Float32 SoA", "peak_GBs": peak, "api_ms_per_call": ms, "kernels": {}}
    for name, (tot, cnt) in prof.items():
        key = name.replace("_kernel", "")
        per = tot / max(cnt, 1)
        nb = nbytes.get(key)
        out["kernels"][name] = {"ms_per_launch": round(per, 4), "launches": cnt,
                                "alg_GBs": round(nb / per / 1e6, 1) if nb and per > 0 else None,
                                "frac_of_peak": round(nb / per / 1e6 / peak, 3) if nb and per > 0 else None}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
