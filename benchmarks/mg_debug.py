#!/usr/bin/env python
"""Debug aid: multigrid primitives on the device against the compiled oracle at production mesh sizes, radial line
of sight, lightcone-like box (observer at the origin outside the box).  python benchmarks/mg_debug.py [n ...]"""
import sys
import time
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "oracle")); sys.path.insert(0, str(ROOT / "tests"))
import __graft_entry__ as G  # noqa: E402
import baorec_oracle_fast as fast  # noqa: E402
from util import rel_rms  # noqa: E402

B = G.load_package()
F = fast.load()
f32 = np.float32
dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
for n in [int(a) for a in sys.argv[1:]] or [128, 256, 512]:
    bs = np.full(3, 2798.33, f32)
    bm = np.array([1400.0, -1400.0, -1400.0], f32)
    rng = np.random.default_rng(n)
    # smooth-ish right-hand side with a sharp mask, like a survey
    f = rng.standard_normal((n, n, n)).astype(f32)
    zz, yy, xx = np.meshgrid(*[np.linspace(-1, 1, n, dtype=f32)] * 3, indexing="ij")
    f *= ((xx ** 2 + yy ** 2 + zz ** 2) < 0.7).astype(f32)
    f -= f.mean(dtype=np.float64).astype(f32)
    v = (0.1 * rng.standard_normal((n, n, n))).astype(f32)
    beta, w = f32(0.344), f32(0.4)
    xv = F.x_vec((n, n, n), bs, bm, f32)
    out = {}
    t0 = time.time()
    for sweeps in (1, 5):
        o = F.jacobi(v.copy(), f, xv, bs, bm, beta, w, sweeps, None)
        g = B.jacobi(dev(v), dev(f), None, bs, bm, float(beta), float(w), sweeps, los=None)
        out[f"jacobi x{sweeps}"] = rel_rms(g.cpu().numpy(), o)
    o = F.residual(v, f, xv, bs, bm, beta, None)
    g = torch.empty((n, n, n), dtype=torch.float32, device="cuda")
    B.residual(g, dev(v), dev(f), None, bs, bm, float(beta), los=None)
    out["residual"] = rel_rms(g.cpu().numpy(), o)
    c = F.restrict(f)
    gc = torch.empty((n // 2,) * 3, dtype=torch.float32, device="cuda")
    B.reduce(gc, dev(f), bs, bm)
    out["restrict"] = rel_rms(gc.cpu().numpy(), c)
    o = F.prolong(np.zeros((n, n, n), f32), c)
    g = torch.zeros((n, n, n), dtype=torch.float32, device="cuda")
    B.prolong(g, dev(c), bs, bm)
    out["prolong"] = rel_rms(g.cpu().numpy(), o)
    o = F.vcycle(v.copy(), f, bs, bm, beta, w, 5, None)
    g = B.vcycle(dev(v), dev(f), bs, bm, float(beta), float(w), 5, los=None)
    gg, oo = g.cpu().numpy(), o
    out["vcycle"] = rel_rms(gg - gg.mean(), oo - oo.mean())
    for lname, los in (("radial", None), ("fixed", (0.0, 0.0, 1.0))):
        o = F.fmg(f.copy(), np.zeros((n, n, n), f32), bs, bm, beta, w, 5, 6, los)
        g = B.fmg(dev(f), None, bs, bm, float(beta), float(w), 5, 6, los=los)
        gg, oo = g.cpu().numpy(), o
        out[f"fmg {lname}"] = rel_rms(gg - gg.mean(), oo - oo.mean())
        out[f"fmg {lname} |phi| rms"] = float(np.sqrt(np.mean((oo - oo.mean()) ** 2)))
    print(n, {k: float(f"{v:.3e}") for k, v in out.items()}, f"{time.time() - t0:.0f}s", flush=True)
