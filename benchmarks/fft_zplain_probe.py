#!/usr/bin/env python
"""A/B of the tile width of the plain z passes (8 or 16 columns per tile): B.smooth (x R2C, y and z forward passes,
Gaussian, z and y inverse passes, x C2R) at 512^3 and 1024^3, ms per call.  (The run kept under profiles/ used a
separate experimental option for the z passes, 512^3 included; what it established is now the automatic choice of
option "fft_tile_cols": 16 columns for the z passes of a 1024-point axis.  This script compares 8 everywhere with auto.)"""
import json
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import __graft_entry__ as G  # noqa: E402

B = G.load_package()
ctx = B.Context.get(0)
for n in (512, 1024):
    fld = torch.rand((n, n, n), device="cuda")
    bs = np.full(3, 2500.0 * n / 1024, np.float32)
    for cols in (8, -1, 8, -1):
        ctx.set_option("own_fft", 1)
        ctx.set_option("fft_tile_cols", cols)
        for _ in range(3):
            B.smooth(fld, 15.0, bs)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ctx.profile(True)
        e0.record()
        for _ in range(10):
            B.smooth(fld, 15.0, bs)
        e1.record()
        torch.cuda.synchronize()
        prof = {k: round(v[0] / 10, 3) for k, v in ctx.profile_read().items()}
        ctx.profile(False)
        print(json.dumps({"mesh": n, "fft_tile_cols": cols, "ms_per_smooth": e0.elapsed_time(e1) / 10, "kernels_ms": prof}), flush=True)
ctx.set_option("fft_tile_cols", -1)
ctx.set_option("own_fft", -1)
