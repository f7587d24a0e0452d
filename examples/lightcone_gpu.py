#!/usr/bin/env python
"""The reference's examples/lightcone_gpu.jl on this engine: sky -> Cartesian with the comoving-distance
interpolator, FKP weights, run! with randoms (box from setup_box(randoms, 500)), reconstructed positions of the
data and of the randoms (sym / iso), Cartesian -> sky.  Every step runs on the device; the reference does the
conversions on CPU threads.  A synthetic survey sector (0.8 < z < 1.0) stands in for the DESI mock, or -- like the
reference's script (examples/lightcone.jl:19-26) -- the catalogs come from space-delimited text files with the columns
ra dec d z nz, cut to 0.8 < z < 1, and the reconstructed sky coordinates go to NPY files (lines 141-150 there).

    python examples/lightcone_gpu.py [--grid 512] [--data 5e6] [--randoms-per-data 10]
    python examples/lightcone_gpu.py --data-file DATA.dat --randoms-file RANDOMS.dat [--out DIR]
"""
import argparse
import sys
import time
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import __graft_entry__ as G  # noqa: E402

BAOrec = G.load_package()
P0 = 5e3


def timed(label, fn):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    out = fn()
    torch.cuda.synchronize()
    print(f"{label}: {1e3 * (time.perf_counter() - t0):.1f} ms")
    return out


def survey(n, gen):
    """ra, dec [deg], redshift, n(z) of a 60 x 60 degree sector."""
    u = lambda: torch.rand(n, device="cuda", generator=gen)
    ra, dec = 120.0 + 60.0 * u(), -10.0 + 60.0 * u()
    z = 0.8 + 0.2 * u()
    nz = 2e-4 * (0.5 + u())
    return ra, dec, z, nz


def from_file(path):
    """CSV.File(fn, delim = " ", header = [:ra, :dec, :d, :z, :nz], types = Float32) + the redshift cut, then to the device."""
    cols = BAOrec.read_text_catalog(path, [0, 1, 3, 4])                         # ra dec z nz into pinned memory
    cols = BAOrec.select_rows(cols, 2, 0.8, 1.0)                                # data_cat[map(z -> ((z > 0.8) & (z < 1)), data_cat.z), :]
    return tuple(c.cuda(non_blocking=True) for c in cols)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--grid", type=int, default=512)
    ap.add_argument("--data", type=float, default=5e6)
    ap.add_argument("--randoms-per-data", type=int, default=10)
    ap.add_argument("--data-file", default=None)
    ap.add_argument("--randoms-file", default=None)
    ap.add_argument("--out", default=None, help="directory for the reconstructed (ra, dec, z) NPY files")
    args = ap.parse_args()
    gen = torch.Generator(device="cuda").manual_seed(42)
    cosmo = BAOrec.Cosmology(z_tab_max=10)                                     # const cosmo = BAOrec.Cosmology(z_tab_max = 10)
    if args.data_file and args.randoms_file:
        d_ra, d_dec, d_z, d_nz = timed("read + cut data", lambda: from_file(args.data_file))
        r_ra, r_dec, r_z, r_nz = timed("read + cut randoms", lambda: from_file(args.randoms_file))
    else:
        d_ra, d_dec, d_z, d_nz = survey(int(args.data), gen)
        r_ra, r_dec, r_z, r_nz = survey(int(args.data) * args.randoms_per_data, gen)

    print("Coordinate conversion")
    data_cat_pos = timed("sky_to_cartesian(data)", lambda: BAOrec.sky_to_cartesian(d_ra, d_dec, d_z, cosmo))
    rand_cat_pos = timed("sky_to_cartesian(randoms)", lambda: BAOrec.sky_to_cartesian(r_ra, r_dec, r_z, cosmo))
    print("Weights")
    data_cat_w, rand_cat_w = BAOrec.fkp_weights(d_nz, P0), BAOrec.fkp_weights(r_nz, P0)
    grid_size = (args.grid,) * 3
    for name, cls, extra in (("iterative", BAOrec.IterativeRecon, dict(n_iter=3)), ("multigrid", BAOrec.MultigridRecon, {})):
        recon = cls(bias=2.2, f=0.757, smoothing_radius=15.0, los=None, **extra)
        print(f"Run {name}")
        timed("run!", lambda: BAOrec.run(recon, grid_size, *data_cat_pos, data_cat_w, *rand_cat_pos, rand_cat_w))
        print("Reading new positions")
        new_pos = timed("data :sum", lambda: BAOrec.reconstructed_positions(recon, *data_cat_pos, field="sum"))
        new_rand_sym = timed("randoms :sum", lambda: BAOrec.reconstructed_positions(recon, *rand_cat_pos, field="sum"))
        new_rand_iso = timed("randoms :disp", lambda: BAOrec.reconstructed_positions(recon, *rand_cat_pos, field="disp"))
        print("Coordinate conversion")
        sky = timed("cartesian_to_sky x3", lambda: [BAOrec.cartesian_to_sky(*c, cosmo) for c in (new_pos, new_rand_sym, new_rand_iso)])
        print("tenth galaxy (ra, dec, z):", [float(c[9]) for c in sky[0]])
        if args.out:                                                            # npzwrite("...dat.rec.npy", hcat(new_pos...)) etc.
            out = Path(args.out)
            out.mkdir(parents=True, exist_ok=True)
            for tag, cat, nz in zip(("dat.rec", "ran.rec.sym", "ran.rec.iso"), sky, (d_nz, r_nz, r_nz)):
                BAOrec.write_npy(out / f"{name}.{tag}.npy", *[c.cpu() for c in cat], nz.cpu())   # hcat(new_pos', nz): (N, 4) ra dec z nz


if __name__ == "__main__":
    main()
