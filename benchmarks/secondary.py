#!/usr/bin/env python
"""Secondary workloads (BASELINE.json configs 0-2): timings + per-kernel breakdown on one GPU.
  C1: IterativeRecon periodic box 256^3, 1e6 particles (the CPU config)
  C2: IterativeRecon lightcone, radial LOS + 10x randoms, 512^3 (CIC and TSC)
  C3: MultigridRecon lightcone 512^3
Prints one JSON line per workload."""
import json
import sys
import time
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
import __graft_entry__ as G  # noqa: E402
from util import lightcone  # noqa: E402

B = G.load_package()
dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()


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


def breakdown(ctx, fn):
    ctx.profile(True)
    fn()
    pr = ctx.profile_read()
    ctx.profile(False)
    agg = {}
    for k, (ms, c) in pr.items():
        agg[k] = {"ms": round(ms, 3), "launches": c}
    return dict(sorted(agg.items(), key=lambda kv: -kv[1]["ms"]))


def main():
    which = sys.argv[1:] or ["c1", "c2", "c2tsc", "c3"]
    ctx = B.Context.get(0)
    rng = np.random.default_rng(42)
    if "c1" in which:
        n, N, L = 256, 1_000_000, 2500.0
        pos = [dev((rng.random(N) * L).astype(np.float32)) for _ in range(3)]
        w = dev(np.ones(N, np.float32))
        kw = dict(bias=2.2, f=0.757, smoothing_radius=15.0, box_size=np.full(3, L, np.float32),
                  box_min=np.zeros(3, np.float32), los=(0.0, 0.0, 1.0), n_iter=3)
        rec = B.IterativeRecon(**kw)

        def step():
            m = B.run(rec, (n, n, n), *pos, w)
            return B.read_shifts(rec, *pos, m, field="sum")
        print(json.dumps({"workload": "C1 iterative box 256^3 1e6", "ms": timed(step), "kernels": breakdown(ctx, step)}))
    if any(k in which for k in ("c2", "c2tsc", "c3")):
        nd, nr = 5_000_000, 50_000_000
        d, wd, r, wr = lightcone(nd, nr, seed=42, rmin=1900.0, rmax=2300.0, half_angle_deg=30.0)
        gd, gr, gwd, gwr = [dev(p) for p in d], [dev(p) for p in r], dev(wd), dev(wr)
        n = 512
        for name, Rec, mas in (("c2", B.IterativeRecon, "cic"), ("c2tsc", B.IterativeRecon, "tsc"),
                               ("c3", B.MultigridRecon, "cic")):
            if name not in which:
                continue
            rec = Rec(bias=2.2, f=0.757, smoothing_radius=15.0, los=None, mas=mas)

            def step():
                m = B.run(rec, (n, n, n), *gd, gwd, *gr, gwr)
                return B.read_shifts(rec, *gd, m, field="sum")
            ms = timed(step, reps=3, warm=2)
            print(json.dumps({"workload": f"{name}: {Rec.__name__} lightcone radial + 10x randoms 512^3 {mas} "
                                          f"({nd:.0e} data, {nr:.0e} randoms)", "ms": ms,
                              "box": [float(v) for v in rec.box_size], "kernels": breakdown(ctx, step)}))


if __name__ == "__main__":
    main()
