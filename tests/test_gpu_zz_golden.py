"""GPU parity against the committed golden fixtures (tests/golden/*.npz; made by
tests/golden/make_golden.py from the oracle, whose own agreement with them is checked on the CPU by
tests/test_golden.py).  Nothing here runs the oracle: the CUDA path, called through the C ABI, is
compared with fixed vectors.  Bit-exact: cells, weights, wrapped positions, gathers of a given
field.  Tolerance (BASELINE.json): mesh / potential rel. rms <= 1e-4, shifts rel. rms <= 1e-4 and
max |ds| <= 1e-3 Mpc/h against the fp32 and the fp64 vectors."""
from pathlib import Path

import numpy as np
import pytest

from util import rel_rms, maxabs

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

GOLD = Path(__file__).resolve().parent / "golden"
PARAMS = dict(bias=2.2, f=0.757, smoothing_radius=15.0)
TOL_RMS, TOL_MAX_SHIFT = 1e-4, 1e-3


def load(name):
    with np.load(GOLD / f"{name}.npz") as z:
        return {k: z[k] for k in z.files}


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def u32(a):
    return np.ascontiguousarray(a).view(np.uint32)


@pytest.mark.parametrize("name", ["iterative_box_32", "multigrid_box_32"])
def test_cells_and_weights_bit_exact(B, name):
    g = load(name)
    n = int(g["n"])
    bs, bm = g["box_size"], g["box_min"]
    d = [dev(g[a]) for a in "xyz"]
    i0, i1, w0, w1 = (t.cpu().numpy() for t in B.cic_cells((n, n, n), *d, bs, bm, True))
    assert np.array_equal(i0, g["cic_i0"]) and np.array_equal(i1, g["cic_i1"])
    assert np.array_equal(u32(w0), u32(g["cic_w0"])) and np.array_equal(u32(w1), u32(g["cic_w1"]))
    for formula in ("cpu", "gpu"):
        idn, iup, wd, wu = (t.cpu().numpy() for t in B.gather_cells((n, n, n), *d, bs, bm, gpu_formula=(formula == "gpu")))
        assert np.array_equal(idn, g[f"gather_{formula}_id"]) and np.array_equal(iup, g[f"gather_{formula}_iu"])
        assert np.array_equal(u32(wd), u32(g[f"gather_{formula}_wd"]))
        assert np.array_equal(u32(wu), u32(g[f"gather_{formula}_wu"]))


@pytest.mark.parametrize("name,cls", [("iterative_box_32", "IterativeRecon"), ("multigrid_box_32", "MultigridRecon")])
def test_box_reconstruction(B, name, cls):
    g = load(name)
    n = int(g["n"])
    kw = dict(PARAMS, box_size=g["box_size"], box_min=g["box_min"], los=tuple(float(v) for v in g["los"]))
    if cls == "IterativeRecon":
        kw["n_iter"] = 3
    rec = getattr(B, cls)(**kw)
    d = [dev(g[a]) for a in "xyz"]
    mesh = B.run(rec, (n, n, n), *d, dev(g["w"]))
    # cic!(wrap=true) writes the wrapped positions back into the caller's arrays (src/mas.jl:8-10)
    for a in range(3):
        assert np.array_equal(u32(d[a].cpu().numpy()), u32(g["cic_wrapped"][a]))
    for tag in ("f32", "f64"):
        assert rel_rms(mesh.cpu().numpy(), g[f"mesh_{tag}"]) < TOL_RMS
    for f in ("disp", "rsd", "sum"):
        s = B.read_shifts(rec, *d, mesh, field=f)
        for a, ax in enumerate("xyz"):
            got = s[a].cpu().numpy()
            for tag in ("f32", "f64"):
                ref = g[f"shift_{tag}_{f}_{ax}"]
                if np.abs(ref).max() == 0:
                    assert np.abs(got).max() == 0
                    continue
                assert rel_rms(got, ref) < TOL_RMS
                assert maxabs(got, ref) < TOL_MAX_SHIFT


@pytest.mark.parametrize("name,cls", [("iterative_lightcone_48", "IterativeRecon"),
                                      ("multigrid_lightcone_48", "MultigridRecon")])
def test_lightcone_reconstruction(B, name, cls):
    g = load(name)
    n = int(g["n"])
    kw = dict(PARAMS, los=None)
    if cls == "IterativeRecon":
        kw["n_iter"] = 3
    d, r = [dev(g[a]) for a in "xyz"], [dev(g[a]) for a in ("rx", "ry", "rz")]
    w, rw = dev(g["w"]), dev(g["rw"])
    # set-up alone: box from the randoms, then the `ran > threshold` decisions (src/recon.jl:85)
    rec = getattr(B, cls)(**kw)
    rec.box_size, rec.box_min = B.setup_box(*r, 500.0)
    assert np.array_equal(rec.box_size, g["box_size"]) and np.array_equal(rec.box_min, g["box_min"])
    delta = torch.zeros((n, n, n), dtype=torch.float32, device="cuda")
    B.setup_fft(rec, delta)
    B.setup_overdensity(delta, rec, *d, w, *r, rw)
    flips = (delta.cpu().numpy() != 0) != g["mask"]
    # the cut is a discontinuity: a cell may only flip if its smoothed randoms density sits within
    # fp32 transform noise of the threshold
    assert flips.sum() <= 5
    assert (np.abs(g["ran_over_threshold"][flips] - 1.0) < 2e-3).all()
    # the whole reconstruction
    rec = getattr(B, cls)(**kw)
    mesh = B.run(rec, (n, n, n), *d, w, *r, rw)
    got = mesh.cpu().numpy()
    m32, m64 = g["mesh_f32"], g["mesh_f64"]
    if cls == "MultigridRecon":      # periodic Poisson problem: the potential is defined up to a constant
        got, m32, m64 = got - got.mean(), m32 - m32.mean(), m64 - m64.mean()
    sg = B.read_shifts(rec, *d, mesh, field="sum")
    if flips.sum() == 0:
        # yardstick = the fp32 oracle's own distance to fp64 (1/(alpha ran) amplifies rounding at the survey edge)
        assert rel_rms(got, m64) < max(TOL_RMS, 2 * rel_rms(m32, m64))
        for a, ax in enumerate("xyz"):
            s, s32, s64 = sg[a].cpu().numpy(), g[f"shift_f32_sum_{ax}"], g[f"shift_f64_sum_{ax}"]
            assert rel_rms(s, s64) < max(TOL_RMS, 2 * rel_rms(s32, s64))
            assert maxabs(s, s64) < max(TOL_MAX_SHIFT, 3 * maxabs(s32, s64))
    else:   # a legitimately flipped cell changes the answer globally: sanity bound only
        for a, ax in enumerate("xyz"):
            err = np.abs(sg[a].cpu().numpy() - g[f"shift_f64_sum_{ax}"])
            assert np.median(err) < 5e-3 and np.quantile(err, 0.9) < 2e-2


def test_mass_assignment(B):
    g = load("mas_24")
    n = int(g["n"])
    bs, bm = g["box_size"], g["box_min"]
    for mas, key in (("cic", "rho_cic"), ("tsc", "rho_tsc")):
        rho = torch.zeros((n, n, n), dtype=torch.float32, device="cuda")
        B.cic(rho, *(dev(g[a]) for a in "xyz"), dev(g["w"]), bs, bm, wrap=True, mas=mas)
        # serial (fixture) vs atomic (GPU) summation order: a few ulp of the cell value
        assert maxabs(rho.cpu().numpy(), g[key]) <= 2e-5 * max(1.0, float(g[key].max()))
    N = len(g["x"])
    fld = dev(g["field"])
    out = torch.empty(N, dtype=torch.float32, device="cuda")
    B.read_cic(out, fld, *(dev(g[a]) for a in "xyz"), bs, bm)
    assert np.array_equal(u32(out.cpu().numpy()), u32(g["read_cic_cpu"]))     # the parity target: CPU formula
    B.read_cic(out, fld, *(dev(g[a]) for a in "xyz"), bs, bm, mas="tsc")
    assert np.array_equal(u32(out.cpu().numpy()), u32(g["read_tsc"]))


def test_multigrid_operators(B):
    g = load("multigrid_ops_32")
    bs, beta = g["box_size"], float(g["beta"])
    for tag, los in (("los", (0.0, 0.0, 1.0)), ("radial", None)):
        bm = g[f"box_min_{tag}"]
        v = dev(g["v"])
        B.jacobi(v, dev(g["f"]), None, bs, bm, beta, 0.4, 3, los=los)
        assert rel_rms(v.cpu().numpy(), g[f"jacobi_{tag}"]) < 1e-5
        r = torch.empty_like(v)
        B.residual(r, dev(g["v"]), dev(g["f"]), None, bs, bm, beta, los=los)
        assert rel_rms(r.cpu().numpy(), g[f"residual_{tag}"]) < 1e-5
        phi = torch.zeros_like(v)
        B.fmg(dev(g["f"]), phi, bs, bm, beta, 0.4, 5, 6, los=los)
        assert rel_rms(phi.cpu().numpy(), g[f"fmg_{tag}"]) < 1e-4
    n = int(g["n"])
    c = torch.empty((n // 2,) * 3, dtype=torch.float32, device="cuda")
    B.reduce(c, dev(g["v"]))
    assert rel_rms(c.cpu().numpy(), g["restrict"]) < 1e-6
    fine = torch.full((n, n, n), float("nan"), dtype=torch.float32, device="cuda")
    B.prolong(fine, dev(g["restrict"]))
    assert rel_rms(fine.cpu().numpy(), g["prolong"]) < 1e-6
