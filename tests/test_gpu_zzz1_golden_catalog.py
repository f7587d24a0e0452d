"""GPU: the catalog kernels against the committed golden fixture tests/golden/catalog_4000.npz (no oracle run).
Same calls and tolerances as tests/test_gpu_zz_catalog.py (which has run on a B200); this file was added after the
GPU budget was spent and sorts after everything validated on hardware."""
from pathlib import Path

import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu
f32 = np.float32


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def ulps(a, b, scale=None):
    a, b = np.asarray(a, f32), np.asarray(b, f32)
    s = np.maximum(np.abs(a), np.abs(b)).astype(f32) if scale is None else np.asarray(scale, f32)
    return np.abs(a.astype(np.float64) - b.astype(np.float64)) / np.spacing(np.maximum(s, f32(1e-30)))


def test_catalog_kernels_match_the_golden_fixture(B):
    with np.load(Path(__file__).resolve().parent / "golden" / "catalog_4000.npz") as zf:
        g = {k: zf[k] for k in zf.files}
    cosmo = B.Cosmology(z_tab_max=3)
    for k in ("h", "H0", "Omega_b0", "Omega_c0", "Omega_g0", "Omega_nu0", "Omega_L0"):
        assert f32(getattr(cosmo, k)) == g[k], k
    z, r = cosmo.tables()
    assert np.abs(z[::1000] - g["z_tab"]).max() < 1e-13 and np.abs(r[::1000][1:] / g["r_tab"][1:] - 1).max() < 1e-12
    x, y, zz = (t.cpu().numpy() for t in B.sky_to_cartesian(dev(g["ra"]), dev(g["dec"]), dev(g["red"]), cosmo))
    scale = np.sqrt(g["x"].astype(float) ** 2 + g["y"].astype(float) ** 2 + g["z"].astype(float) ** 2).astype(f32)
    for a, b in ((x, g["x"]), (y, g["y"]), (zz, g["z"])):
        assert ulps(a, b, scale).max() <= 2
    a, d, q = (t.cpu().numpy() for t in B.cartesian_to_sky(dev(g["x"]), dev(g["y"]), dev(g["z"]), cosmo))
    assert ulps(a, g["ra2"]).max() <= 2 and ulps(d, g["dec2"], np.maximum(np.abs(g["dec2"]), 1e-3)).max() <= 2 and ulps(q, g["red2"]).max() <= 2
    assert np.array_equal(B.fkp_weights(dev(g["nz"]), 5e3).cpu().numpy().view(np.uint32), g["fkp"].view(np.uint32))
    p = [dev(g[k]) for k in "xyz"]
    B.wrap_positions(*p, g["wrap_box_size"], g["wrap_box_min"])
    for t, k in zip(p, ("wx", "wy", "wz")):
        assert np.array_equal(t.cpu().numpy().view(np.uint32), g[k].view(np.uint32))
