#!/usr/bin/env python
"""Timing of the multipole estimator as the reference's helpers configure it (test_helpers/powspec_auto.conf: 512^3 grid,
TSC, interlaced; test_helpers/simulation.py:36-70: compute_auto_box on the catalog before / after reconstruction):
one baorec_compute_auto_box_f32 call on a 1e8-particle periodic catalog, per scheme, with the kernels behind it.

    python benchmarks/pk_bench.py [--mesh 512] [--particles 1e8] [--steps 3]"""
import argparse
import json
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "benchmarks"))
import __graft_entry__ as G  # noqa: E402
import catalogs  # noqa: E402

B = G.load_package()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mesh", type=int, default=512)
    ap.add_argument("--particles", type=float, default=1e8)
    ap.add_argument("--steps", type=int, default=3)
    args = ap.parse_args()
    n, N, L = args.mesh, int(args.particles), 2500.0
    pos, w = catalogs.uniform_box(N, L, seed=42, device="cuda")
    bs = np.full(3, L, np.float32)
    ctx = B.Context.get(0)
    for mas, interlace in (("cic", False), ("tsc", False), ("tsc", True), ("pcs", True)):
        kw = dict(mas=mas, interlace=interlace, dk=0.005, nbins=128, shot=L ** 3 / N)
        for _ in range(2):
            r = B.compute_auto_box(*pos, w, bs, (n, n, n), **kw)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ctx.profile(True)
        e0.record()
        for _ in range(args.steps):
            r = B.compute_auto_box(*pos, w, bs, (n, n, n), **kw)
        e1.record()
        torch.cuda.synchronize()
        prof = {k: round(v[0] / args.steps, 3) for k, v in sorted(ctx.profile_read().items(), key=lambda kv: -kv[1][0])[:8]}
        ctx.profile(False)
        hi = np.nanmax(np.abs(r["p0"][-8:]))                                   # residual of the shot-noise-subtracted monopole near k_max
        print(json.dumps({"workload": f"compute_auto_box {n}^3, {N:.0e} particles, {mas}{' interlaced' if interlace else ''}",
                          "ms": e0.elapsed_time(e1) / args.steps, "kernels_ms": prof, "max_abs_p0_minus_shot_last_bins": float(hi),
                          "shot": L ** 3 / N}), flush=True)


if __name__ == "__main__":
    main()
