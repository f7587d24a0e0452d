#!/usr/bin/env python
"""Per-pass timings of the column FFT kernels (csrc/fft.cu, option own_fft) against cuFFT's 3-D plans at n^3:
smooth! (R2C + Gaussian + C2R) with own_fft = 0 / 1, then the fused run! + read_shifts of the bench workload.
    python benchmarks/fft_probe.py [n] [particles]"""
import json
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import __graft_entry__ as G  # noqa: E402

B = G.load_package()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
N = int(float(sys.argv[2])) if len(sys.argv) > 2 else 20_000_000
L = 2500.0 * n / 1024
bs, bm = np.full(3, L, np.float32), np.zeros(3, np.float32)
ctx = B.Context.get(0)
fld = torch.rand((n, n, n), device="cuda")
g = torch.Generator(device="cuda").manual_seed(1)
pos = [torch.rand(N, device="cuda", generator=g) * (L * 0.999) for _ in range(3)]
w = torch.ones(N, device="cuda")
kw = dict(bias=2.2, f=0.757, smoothing_radius=15.0, n_iter=3, los=(0.0, 0.0, 1.0), box_size=bs, box_min=bm)
for own in (0, 1):
    ctx.set_option("own_fft", own)
    rec = B.IterativeRecon(**kw)
    for it in range(3):
        if it == 2:
            ctx.profile(True)
        B.smooth(fld, 15.0, bs)
        mesh = B.run(rec, (n, n, n), *pos, w)
        s = B.read_shifts(rec, *pos, mesh, field="sum")
    torch.cuda.synchronize()
    prof = ctx.profile_read()
    ctx.profile(False)
    keep = {k: (round(v[0], 3), v[1]) for k, v in prof.items() if "fft" in k or "kspace" in k}
    print(json.dumps({"own_fft": own, "n": n, "ms (total, launches)": keep, "sum_ms": round(sum(v[0] for v in keep.values()), 3),
                      "checksum": float(s[2].double().abs().mean())}), flush=True)
ctx.set_option("own_fft", -1)
