#!/usr/bin/env python
"""Generates the committed golden fixtures under tests/golden/ from the CPU oracle.

    python tests/golden/make_golden.py            # rewrites tests/golden/*.npz + MANIFEST.json

The reference (BAOrec.jl) cannot run in this image (no Julia) and ships no usable vectors
(SURVEY.md section 8c), so these fixtures are outputs of oracle/baorec_oracle.py -- itself pinned
by the analytic known-answer tests in tests/test_oracle_kat.py -- on small seeded catalogs.  They
do two jobs: (1) a regression lock on the oracle (tests/test_golden.py, CPU), (2) fixed vectors the
CUDA path is compared against through the C ABI without running the oracle
(tests/test_gpu_golden.py).  Inputs are stored next to the outputs, so a fixture never depends on
a random-number generator's stream staying stable.

Every fixture carries the fp32-faithful oracle result and the fp64 one (the tolerance yardstick of
BASELINE.json: rel. rms <= 1e-4, max |ds| <= 1e-3 Mpc/h)."""
from __future__ import annotations

import hashlib
import json
import sys
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
ROOT = HERE.parent.parent
sys.path.insert(0, str(ROOT / "oracle"))
sys.path.insert(0, str(ROOT / "tests"))

import baorec_oracle as O  # noqa: E402
import catalog_oracle as CO  # noqa: E402
import pk_oracle as PK  # noqa: E402
from util import clustered_box, lightcone  # noqa: E402

PARAMS = dict(bias=2.2, f=0.757, smoothing_radius=15.0)
FIELDS = ("disp", "rsd", "sum")


def _shifts(rec, pos, mesh, prefix, out):
    for f in FIELDS:
        s = O.read_shifts(rec, *pos, mesh, f)
        for a, ax in enumerate("xyz"):
            out[f"{prefix}_{f}_{ax}"] = np.asarray(s[a], np.float32)


def box_case(algorithm, n, L, N, los, seed):
    """Periodic box (examples/simulation.jl style): run! + read_shifts for the three fields."""
    pos, w = clustered_box(N, L, seed=seed)
    out = {"x": pos[0], "y": pos[1], "z": pos[2], "w": w,
           "box_size": np.full(3, L, np.float32), "box_min": np.zeros(3, np.float32),
           "los": np.asarray(los, np.float32), "n": np.int64(n)}
    cls = O.IterativeRecon if algorithm == "iterative" else O.MultigridRecon
    for T, tag in ((np.float32, "f32"), (np.float64, "f64")):
        kw = dict(PARAMS, box_size=np.full(3, L, T), box_min=np.zeros(3, T), los=tuple(los))
        if algorithm == "iterative":
            kw["n_iter"] = 3
        rec = cls(**kw)
        p = [q.astype(T) for q in pos]
        mesh = O.run(rec, (n, n, n), *[q.copy() for q in p], w.astype(T))
        out[f"mesh_{tag}"] = mesh.astype(np.float32)      # fp64 truth stored rounded to fp32 (tolerances are 1e-4)
        _shifts(rec, p, mesh, f"shift_{tag}", out)
    # bit-exact material: scatter cells / weights (src/mas.jl:13-38), gather cells / weights for the
    # CPU and the GPU coordinate formula of the reference (src/mas.jl:224 vs :274)
    bs, bm = out["box_size"], out["box_min"]
    wrapped, i0, i1, w0, w1 = O.cic_cells(*[q.copy() for q in pos], (n, n, n), bs, bm, wrap=True)
    out["cic_wrapped"] = np.stack(wrapped)
    out["cic_i0"], out["cic_i1"] = np.stack(i0).astype(np.int32), np.stack(i1).astype(np.int32)
    out["cic_w0"], out["cic_w1"] = np.stack(w0), np.stack(w1)
    for formula in ("cpu", "gpu"):
        idn, iup, wd, wu = O.gather_cells(*pos, (n, n, n), bs, bm, wrap=True, formula=formula)
        out[f"gather_{formula}_id"] = np.stack(idn).astype(np.int32)
        out[f"gather_{formula}_iu"] = np.stack(iup).astype(np.int32)
        out[f"gather_{formula}_wd"], out[f"gather_{formula}_wu"] = np.stack(wd), np.stack(wu)
    return out


def lightcone_case(algorithm, n, nd, nr, seed):
    """Lightcone: radial line of sight + randoms, box from setup_box(randoms, 500)
    (examples/lightcone.jl / lightcone_mg.jl style).  The threshold mask of the fp32 run is stored
    and forced on the fp64 run (the cut is a discontinuity)."""
    d, wd, r, wr = lightcone(nd, nr, seed=seed, rmin=500.0, rmax=800.0, half_angle_deg=25.0)
    out = {"x": d[0], "y": d[1], "z": d[2], "w": wd, "rx": r[0], "ry": r[1], "rz": r[2], "rw": wr,
           "n": np.int64(n)}
    cls = O.IterativeRecon if algorithm == "iterative" else O.MultigridRecon
    kw = dict(PARAMS, los=None)
    if algorithm == "iterative":
        kw["n_iter"] = 3
    rec = cls(**kw)
    rec.box_size, rec.box_min = O.setup_box(*r, np.float32(500))
    info = {}
    O.setup_overdensity(np.zeros((n, n, n), np.float32), rec, *d, wd, *r, wr, info=info)
    mask = info["ran"] > info["threshold"]
    out["box_size"], out["box_min"] = rec.box_size, rec.box_min
    out["mask"] = mask
    out["ran_over_threshold"] = (info["ran"].astype(np.float64) / info["threshold"]).astype(np.float32)
    for T, tag in ((np.float32, "f32"), (np.float64, "f64")):
        rec = cls(**kw)
        dd, rr = [q.astype(T) for q in d], [q.astype(T) for q in r]
        mesh = O.run(rec, (n, n, n), *dd, wd.astype(T), *rr, wr.astype(T), force_mask=mask)
        out[f"mesh_{tag}"] = mesh.astype(np.float32)
        _shifts(rec, dd, mesh, f"shift_{tag}", out)
    return out


def mas_case(n, L, N, seed):
    """Mass assignment alone: CIC (reference) and TSC (extension) density meshes and gathers."""
    pos, w = clustered_box(N, L, seed=seed)
    bs, bm = np.full(3, L, np.float32), np.zeros(3, np.float32)
    out = {"x": pos[0], "y": pos[1], "z": pos[2], "w": w, "box_size": bs, "box_min": bm, "n": np.int64(n)}
    rho = np.zeros((n, n, n), np.float32)
    O.cic_scatter(rho, *[q.copy() for q in pos], w, bs, bm, wrap=True)
    out["rho_cic"] = rho
    rho = np.zeros((n, n, n), np.float32)
    O.tsc_scatter(rho, *[q.copy() for q in pos], w, bs, bm, wrap=True)
    out["rho_tsc"] = rho
    rng = np.random.default_rng(seed + 1)
    fld = rng.standard_normal((n, n, n)).astype(np.float32)
    out["field"] = fld
    out["read_cic_cpu"] = O.read_cic(fld, *pos, bs, bm, wrap=True, formula="cpu")
    out["read_cic_gpu"] = O.read_cic(fld, *pos, bs, bm, wrap=True, formula="gpu")
    out["read_tsc"] = O.read_tsc(fld, *pos, bs, bm, wrap=True)
    return out


def multigrid_ops_case(n, L, seed):
    """One call of each multigrid operator (src/multigrid.jl) on a random mesh, both LOS modes."""
    rng = np.random.default_rng(seed)
    v = rng.standard_normal((n, n, n)).astype(np.float32)
    f = rng.standard_normal((n, n, n)).astype(np.float32)
    f -= f.mean()
    bs = np.full(3, L, np.float32)
    out = {"v": v, "f": f, "box_size": bs, "n": np.int64(n), "beta": np.float32(0.344),
           "omega": np.float32(0.4)}
    for tag, los, lo in (("los", (0.0, 0.0, 1.0), 0.0), ("radial", None, 900.0)):
        bm = np.full(3, lo, np.float32)
        xv = O.x_vec((n, n, n), bs, bm, np.float32)
        out[f"box_min_{tag}"] = bm
        out[f"jacobi_{tag}"] = O.jacobi(v.copy(), f, xv, bs, bm, out["beta"], out["omega"], 3, los)
        out[f"residual_{tag}"] = O.residual(v, f, xv, bs, bm, out["beta"], los)
        out[f"fmg_{tag}"] = O.fmg(f, np.zeros_like(f), bs, bm, out["beta"], out["omega"], 5, 6, los)
    out["restrict"] = O.restrict(v)
    out["prolong"] = O.prolong(np.zeros_like(v), out["restrict"])
    return out


def catalog_case(n, seed):
    """Catalog pre/post-processing (src/cosmo.jl, examples/lightcone.jl:30-82): cosmology tables (every 1000th knot),
    sky -> Cartesian -> sky, FKP weights, periodic re-wrap."""
    rng = np.random.default_rng(seed)
    c = CO.Cosmology(z_tab_max=3)
    ra, dec = (360 * rng.random(n)).astype(np.float32), (180 * rng.random(n) - 90).astype(np.float32)
    red = (0.01 + 2.9 * rng.random(n)).astype(np.float32)
    nz = (1e-3 * rng.random(n)).astype(np.float32)
    zt, rt = CO.tables(c)
    x, y, z = CO.sky_to_cartesian(ra, dec, red, c)
    ra2, dec2, red2 = CO.cartesian_to_sky(x, y, z, c)
    box = (np.float32([1500.0, 2000.0, 2500.0]), np.float32([-700.0, 0.0, 100.0]))
    wx, wy, wz = CO.wrap_positions(x, y, z, *box)
    dens = {k: np.float32(getattr(c, k)) for k in ("h", "H0", "Omega_b0", "Omega_c0", "Omega_g0", "Omega_nu0", "Omega_L0")}
    return dict(ra=ra, dec=dec, red=red, nz=nz, z_tab=zt[::1000], r_tab=rt[::1000], x=x, y=y, z=z, ra2=ra2, dec2=dec2, red2=red2,
                fkp=CO.fkp_weights(nz, np.float32(5e3)), wrap_box_size=box[0], wrap_box_min=box[1], wx=wx, wy=wy, wz=wz, **dens)


def pk_case(shape_xyz, L, N, seed):
    """Power-spectrum multipoles (oracle/pk_oracle.py) of a CIC density mesh: the mesh itself is the input."""
    nx, ny, nz = shape_xyz
    bs = np.asarray(L, np.float32)
    rng = np.random.default_rng(seed)
    centres = rng.random((40, 3))
    p = (centres[rng.integers(0, 40, N)] + 0.05 * rng.standard_normal((N, 3))) % 1.0
    pos = [np.minimum((bs[a] * p[:, a]).astype(np.float32), np.nextafter(bs[a], np.float32(0))) for a in range(3)]
    w = (0.5 + rng.random(N)).astype(np.float32)
    rho = O.cic_scatter(np.zeros((nz, ny, nx), np.float32), *pos, w, bs, np.zeros(3, np.float32), True)
    los = np.float32([0.2, -0.4, 0.9])
    shot = float(np.prod(bs.astype(np.float64))) * float((w.astype(np.float64) ** 2).sum()) / float(w.sum(dtype=np.float64)) ** 2
    out = dict(rho=rho, box_size=bs, los=los, kmin=np.float64(0.004), dk=np.float64(0.011), nbins=np.int64(30), shot=np.float64(shot))
    for power, tag in ((2, "cic"), (0, "raw")):
        r = PK.power_multipoles(rho, bs, los=los, kmin=0.004, dk=0.011, nbins=30, mas_power=power, shot=shot)
        for k, v in r.items():
            out[f"{tag}_{k}"] = np.asarray(v, np.float64)
    return out


CASES = {
    "iterative_box_32": lambda: box_case("iterative", 32, 431.7, 6000, (0.0, 0.0, 1.0), 101),
    "multigrid_box_32": lambda: box_case("multigrid", 32, 431.7, 6000, (0.0, 0.0, 1.0), 102),
    "iterative_lightcone_48": lambda: lightcone_case("iterative", 48, 4000, 30000, 103),
    "multigrid_lightcone_48": lambda: lightcone_case("multigrid", 48, 4000, 30000, 104),
    "mas_24": lambda: mas_case(24, 300.0, 5000, 105),
    "multigrid_ops_32": lambda: multigrid_ops_case(32, 500.0, 106),
    "catalog_4000": lambda: catalog_case(4000, 107),
    "pk_40": lambda: pk_case((40, 36, 44), (400.0, 360.0, 440.0), 50000, 108),
}


def main():
    only = sys.argv[1:]                      # `make_golden.py pk_40` regenerates one fixture and keeps the others' entries
    mf = HERE / "MANIFEST.json"
    manifest = json.loads(mf.read_text()) if (only and mf.exists()) else {}
    for name, fn in CASES.items():
        if only and name not in only:
            continue
        data = fn()
        path = HERE / f"{name}.npz"
        np.savez_compressed(path, **data)
        h = hashlib.sha256()
        for k in sorted(data):
            h.update(k.encode())
            h.update(np.ascontiguousarray(data[k]).tobytes())
        manifest[name] = {"arrays": len(data), "sha256_of_arrays": h.hexdigest(),
                          "bytes": path.stat().st_size}
        print(f"{name}: {len(data)} arrays, {path.stat().st_size / 1024:.0f} KiB")
    (HERE / "MANIFEST.json").write_text(json.dumps(manifest, indent=1) + "\n")


if __name__ == "__main__":
    main()
