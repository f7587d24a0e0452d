#!/usr/bin/env python
"""A/B harness: the bench workload (or a smaller one) under a list of option settings, one JSON line each, so that
one gpurun call answers "does option X pay off?" (ms per reconstruction + the five most expensive kernels).

    python benchmarks/ab_options.py [--mesh 1024] [--particles 1e8] [--catalog uniform|lognormal] [--steps 5]
                                    [--set name=value ...]      # extra configurations, e.g. --set deterministic_scatter=1

Default configurations: baseline; unified_sort=0; gather_tiles=0; deterministic_scatter=1; fuse_kspace=0; gather_stage=0 / 1; scatter_pairs=0; own_fft=0."""
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
DEFAULTS = {"unified_sort": 1, "gather_tiles": 1, "deterministic_scatter": 0, "fuse_kspace": 1, "gather_stage": 2, "scatter_pairs": 2, "own_fft": -1}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mesh", type=int, default=1024)
    ap.add_argument("--particles", type=float, default=1e8)
    ap.add_argument("--catalog", default="uniform", choices=["uniform", "lognormal"])
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--set", action="append", default=[], metavar="NAME=VALUE")
    ap.add_argument("--only-set", action="store_true", help="baseline + the --set configurations only")
    args = ap.parse_args()
    n, N = args.mesh, int(args.particles)
    L = 2500.0 * n / 1024.0
    if args.catalog == "lognormal":
        pos, w = catalogs.lognormal_box(N, L, seed=42, device="cuda", n_gen=min(n, 512), f_rsd=0.757)
    else:
        pos, w = catalogs.uniform_box(N, L, seed=42, device="cuda")
    kw = dict(bias=2.2, f=0.757, smoothing_radius=15.0, n_iter=3, los=(0.0, 0.0, 1.0), box_size=np.full(3, L, np.float32),
              box_min=np.zeros(3, np.float32))
    ctx = B.Context.get(0)
    configs = [{}] + ([] if args.only_set else [{k: 0 if v else 1} for k, v in DEFAULTS.items()] + [{"gather_stage": 1}])   # (own_fft: auto -> 0 = cuFFT's 3-D plans)
    for item in args.set:
        k, v = item.split("=")
        configs.append({k: int(v)})
    mesh_buf = torch.empty((n, n, n), dtype=torch.float32, device="cuda")
    out_buf = tuple(torch.empty_like(pos[0]) for _ in range(3))
    for cfg in configs:
        for k, v in {**DEFAULTS, **cfg}.items():
            ctx.set_option(k, v)
        rec = B.IterativeRecon(**kw)

        def step():
            mesh = B.run(rec, (n, n, n), *pos, w, mesh_out=mesh_buf)
            return B.read_shifts(rec, *pos, mesh, field="sum", out=out_buf)

        for _ in range(args.warmup):
            step()
        torch.cuda.synchronize()
        ctx.profile(True)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            s = step()
        e1.record()
        torch.cuda.synchronize()
        prof = ctx.profile_read()
        ctx.profile(False)
        top = sorted(prof.items(), key=lambda kv: -kv[1][0])[:14]
        print(json.dumps({"options": cfg or "defaults", "catalog": args.catalog, "mesh": n, "particles": N,
                          "ms_per_reconstruction": round(e0.elapsed_time(e1) / args.steps, 3),
                          "top_kernels_ms_per_step": {k: round(v[0] / args.steps, 3) for k, v in top},
                          "checksum": float(s[2].double().abs().mean())}), flush=True)
    for k, v in DEFAULTS.items():
        ctx.set_option(k, v)


if __name__ == "__main__":
    main()
