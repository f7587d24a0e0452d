#!/usr/bin/env python
"""The reference's examples/simulation_gpu.jl, line for line, on this engine (Python mirror of the same C ABI the
Julia shim binds).  The DESI mock the reference reads is not public, so a synthetic lognormal box of the same shape
stands in for it (--npy x.npy y.npy z.npy loads a real catalog instead).

    python examples/simulation_gpu.py [--grid 512] [--particles 5e6] [--out DIR]
"""
import argparse
import sys
import time
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "benchmarks"))
import __graft_entry__ as G  # noqa: E402
import catalogs  # noqa: E402

BAOrec = G.load_package()


def timed(label, fn):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    out = fn()
    torch.cuda.synchronize()
    print(f"{label}: {1e3 * (time.perf_counter() - t0):.1f} ms")
    return out


def multipoles(label, pos, w, grid_size, box_size, box_min, randoms=None):
    """P_0 / P_2 of a catalog (minus shifted randoms) -- what test_helpers/simulation.py:36-70 plots with pypowspec
    (compute_auto_box / compute_auto_box_rand), here on the device."""
    def mesh(cat, ww):
        rho = torch.zeros(grid_size[::-1], dtype=torch.float32, device="cuda")
        BAOrec.cic(rho, *[p.clone() for p in cat], ww, box_size, box_min, wrap=True)
        return rho
    shot = float(np.prod(box_size)) / len(w) * (1 + (len(w) / len(randoms[0]) if randoms is not None else 0))
    pk = BAOrec.power_multipoles(mesh(pos, w), box_size, los=(0.0, 0.0, 1.0), kmin=0.0, dk=0.01, nbins=10, mas="cic", shot=shot,
                                 randoms=None if randoms is None else mesh(randoms, torch.ones_like(randoms[0])))
    print(f"{label}: k = {np.round(pk['k'][1:6], 3)}  P0 = {np.round(pk['p0'][1:6], 0)}  P2/P0 = {np.round(pk['p2'][1:6] / pk['p0'][1:6], 2)}")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--grid", type=int, default=512)
    ap.add_argument("--particles", type=float, default=5e6)
    ap.add_argument("--npy", nargs=3, metavar=("X", "Y", "Z"))
    ap.add_argument("--out", default=None)
    args = ap.parse_args()

    box_size = np.float32([1000.0, 1000.0, 1000.0])
    box_min = np.float32([0.0, 0.0, 0.0])
    grid_size = (args.grid,) * 3
    los = (0.0, 0.0, 1.0)
    if args.npy:
        data_cat_pos = [torch.from_numpy(np.load(f).astype(np.float32)).cuda() for f in args.npy]
    else:
        data_cat_pos, _ = catalogs.lognormal_box(int(args.particles), 1000.0, seed=42, device="cuda", n_gen=256, f_rsd=0.757)
    data_cat_w = torch.zeros_like(data_cat_pos[0]) + 1

    multipoles("Pre", data_cat_pos, data_cat_w, grid_size, box_size, box_min)
    for name, cls, extra in (("iterative", BAOrec.IterativeRecon, dict(n_iter=3)), ("multigrid", BAOrec.MultigridRecon, {})):
        recon = cls(bias=2.2, f=0.757, smoothing_radius=15.0, box_size=box_size, box_min=box_min, los=los, **extra)
        print(f"Run {name}")
        timed("run!", lambda: BAOrec.run(recon, grid_size, *data_cat_pos, data_cat_w))
        new_pos = timed("reconstructed_positions(data, :sum)", lambda: BAOrec.reconstructed_positions(recon, *data_cat_pos, field="sum"))
        g = torch.Generator(device="cuda").manual_seed(1)
        recon_cat_pos = [float(box_size[i]) * torch.rand(10 * len(data_cat_pos[i]), device="cuda", generator=g) for i in range(3)]
        new_rand_sym = timed("reconstructed_positions(randoms, :sum)", lambda: BAOrec.reconstructed_positions(recon, *recon_cat_pos, field="sum"))
        new_rand_iso = timed("reconstructed_positions(randoms, :disp)", lambda: BAOrec.reconstructed_positions(recon, *recon_cat_pos, field="disp"))
        # the reference's helper scripts re-wrap before measuring P(k) (test_helpers/simulation.py:51-52); here on the device
        for cat in (new_pos, new_rand_sym, new_rand_iso):
            BAOrec.wrap_positions(*cat, box_size, box_min)
        multipoles(f"{name}: Iso", new_pos, data_cat_w, grid_size, box_size, box_min, randoms=new_rand_iso)
        multipoles(f"{name}: Sym", new_pos, data_cat_w, grid_size, box_size, box_min, randoms=new_rand_sym)
        if args.out:
            out = Path(args.out)
            out.mkdir(parents=True, exist_ok=True)
            for tag, cat in (("dat.rec", new_pos), ("ran.rec.sym", new_rand_sym), ("ran.rec.iso", new_rand_iso)):
                np.save(out / f"{name}.{tag}.npy", np.stack([t.cpu().numpy() for t in cat], axis=1))


if __name__ == "__main__":
    main()
