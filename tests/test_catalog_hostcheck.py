"""CPU checks of the catalog pre/post-processing that need no GPU:
  * the product's host-side comoving-distance table (baorec_cosmo_build_table) against the oracle;
  * the `Cosmology` mirror's derived densities against the oracle's restatement of src/cosmo.jl;
  * the per-particle arithmetic the CUDA kernels execute (baorec.jl_b200/csrc/catalog_math.cuh),
    compiled as plain C++ by tests/hostcheck/ and run on the CPU, against the oracle -- bit-exact for
    the Float32-only formulas (FKP weights, re-wrap), within 1 ulp otherwise.
The same comparisons run on the device in tests/test_gpu_zz_catalog.py."""
import ctypes as C
import subprocess
from pathlib import Path

import numpy as np
import pytest

import catalog_oracle as CO

ROOT = Path(__file__).resolve().parent.parent
f32 = np.float32
_F, _D = C.POINTER(C.c_float), C.POINTER(C.c_double)


def fp(a):
    return a.ctypes.data_as(_F)


@pytest.fixture(scope="module")
def HC():
    out = ROOT / "tests" / "_build" / "libcatalog_hostcheck.so"
    src = ROOT / "tests" / "hostcheck" / "catalog_hostcheck.cpp"
    hdr = ROOT / "baorec.jl_b200" / "csrc" / "catalog_math.cuh"
    if not out.exists() or out.stat().st_mtime < max(src.stat().st_mtime, hdr.stat().st_mtime):
        out.parent.mkdir(exist_ok=True)
        gxx = "/usr/bin/g++" if Path("/usr/bin/g++").exists() else "g++"
        subprocess.run([gxx, "-O2", "-ffp-contract=off", "-shared", "-fPIC", "-o", str(out), str(src)], check=True)
    lib = C.CDLL(str(out))
    i64, d, f = C.c_int64, C.c_double, C.c_float
    lib.hc_sky_to_cartesian.restype = i64
    lib.hc_sky_to_cartesian.argtypes = [_F, _F, _F, i64, f, _D, d, d, d, i64, _F, _F, _F]
    lib.hc_cartesian_to_sky.restype = i64
    lib.hc_cartesian_to_sky.argtypes = [_F, _F, _F, i64, f, _D, d, d, i64, _F, _F, _F, C.c_int, C.c_int, C.POINTER(C.c_int)]
    lib.hc_sincos.restype = None
    lib.hc_sincos.argtypes = [_F, i64, _D, _D]
    lib.hc_fkp_weights.restype = None
    lib.hc_fkp_weights.argtypes = [_F, i64, f, _F]
    lib.hc_wrap_positions.restype = None
    lib.hc_wrap_positions.argtypes = [_F, _F, _F, i64, _F, _F]
    lib.hc_atan2.restype = None
    lib.hc_atan2.argtypes = [_D, _D, i64, _D]
    return lib


def product_table(B, cosmo):
    p = cosmo._params()
    n = int(cosmo.z_tab_num)
    z, r = np.empty(n), np.empty(n)
    B.lib_loader.check(B.lib_loader.load().baorec_cosmo_build_table(C.byref(p), z.ctypes.data_as(_D), r.ctypes.data_as(_D)))
    return z, r


def ulps(a, b):
    a, b = np.asarray(a, f32), np.asarray(b, f32)
    return np.abs(a.astype(np.float64) - b.astype(np.float64)) / np.spacing(np.maximum(np.abs(a), np.abs(b)).astype(f32))


@pytest.mark.parametrize("kw", [dict(), dict(z_tab_max=10), dict(z_tab_min=0.4, z_tab_max=1.6, z_tab_num=4097),
                                dict(h=0.6736, omega_b=0.02237, Neff=3.0328, z_tab_num=1000), dict(w0=-0.9, wa=0.1, z_tab_max=2)])
def test_cosmology_mirror_and_distance_table(B, kw):
    mine, ref = B.Cosmology(**kw), CO.Cosmology(**kw)
    for k in ("h", "h2", "H0", "Omega_b0", "Omega_c0", "Omega_g0", "Omega_nu0", "Omega_k0", "Omega_L0", "Omega_m0", "w0", "wa"):
        assert f32(getattr(mine, k)) == f32(getattr(ref, k)), k
    z, r = product_table(B, mine)
    zo, ro = CO.tables(ref)
    assert np.abs(z - zo).max() < 1e-14 * max(1.0, zo[-1])
    assert np.abs(r[1:] / ro[1:] - 1).max() < 1e-12 and abs(r[0] - ro[0]) < 1e-9 * max(1.0, ro[0])
    i = len(z) // 3
    assert abs(r[i] / CO.comoving_distance(ref, z[i]) - 1) < 1e-9            # vs adaptive quadrature (quadgk's rtol is 1e-8)


def test_desi_cosmology_mirror(B):
    a, b = B.DESICosmology(z_tab_max=4), CO.DESICosmology(z_tab_max=4)
    for k in ("h", "Neff", "Omega_b0", "Omega_c0", "Omega_g0", "Omega_nu0", "Omega_L0"):
        assert f32(getattr(a, k)) == f32(getattr(b, k)), k


def test_bad_cosmology_is_an_error(B):
    lib = B.lib_loader.load()
    r = np.empty(8)
    p = B.Cosmology(z_tab_num=8)._params()
    assert lib.baorec_cosmo_build_table(C.byref(p), None, None) == B.lib_loader.ERR_INVALID
    p.z_tab_num = 1
    assert lib.baorec_cosmo_build_table(C.byref(p), None, r.ctypes.data_as(_D)) == B.lib_loader.ERR_INVALID
    p.z_tab_num, p.z_tab_max = 8, -1.0
    assert lib.baorec_cosmo_build_table(C.byref(p), None, r.ctypes.data_as(_D)) == B.lib_loader.ERR_INVALID
    assert b"z_tab" in lib.baorec_last_error()


def sky_catalog(n, seed, zmax):
    rng = np.random.default_rng(seed)
    ra, dec = (360 * rng.random(n)).astype(f32), (180 * rng.random(n) - 90).astype(f32)
    red = (zmax * rng.random(n)).astype(f32)
    ra[:4], dec[:4] = f32([0, 90, 180, 359.99]), f32([0, -90, 90, 45])
    red[:3] = f32([0, zmax, zmax / 2])
    return ra, dec, red


@pytest.mark.parametrize("kw", [dict(z_tab_max=3), dict(z_tab_max=10), dict(z_tab_max=2, z_tab_num=1500)])
def test_device_arithmetic_of_sky_to_cartesian(HC, B, kw):
    cosmo, ref = B.Cosmology(**kw), CO.Cosmology(**kw)
    z, r = product_table(B, cosmo)
    ra, dec, red = sky_catalog(50000, 3, float(kw["z_tab_max"]))
    x, y, zz = (np.empty_like(ra) for _ in range(3))
    h = f32(cosmo.H0 / f32(100))
    bad = HC.hc_sky_to_cartesian(fp(ra), fp(dec), fp(red), len(ra), h, r.ctypes.data_as(_D), z[0], z[-1], (z[-1] - z[0]) / (len(z) - 1), len(z), fp(x), fp(y), fp(zz))
    assert bad == 0
    ox, oy, oz = CO.sky_to_cartesian(ra, dec, red, ref)
    scale = np.sqrt(ox.astype(float) ** 2 + oy.astype(float) ** 2 + oz.astype(float) ** 2).astype(f32)
    for a, b in ((x, ox), (y, oy), (zz, oz)):
        assert (np.abs(a.astype(float) - b.astype(float)) <= 1.01 * np.spacing(scale)).all()     # <= 1 ulp of |p|
        assert (a == b).mean() > 0.99


def test_reduced_sincos_is_accurate_enough_to_round_correctly(HC):
    rng = np.random.default_rng(11)
    x = np.concatenate([(4 * np.pi * rng.random(200000) - 2 * np.pi), 1e4 * rng.standard_normal(50000),
                        [0.0, np.pi / 2, np.pi, -np.pi / 4, 7e4, -9.9e4, 1e5, 3e7, -1e30]]).astype(f32)
    s, c = np.empty(len(x)), np.empty(len(x))
    HC.hc_sincos(fp(x), len(x), s.ctypes.data_as(_D), c.ctypes.data_as(_D))
    xs = x.astype(np.float64)
    small = np.abs(xs) < 1e5
    # Cody-Waite + Taylor path: absolute error ~1e-16 (the library path beyond 1e5 is the platform's own sincos)
    assert np.abs(s - np.sin(xs))[small].max() < 4e-16 and np.abs(c - np.cos(xs))[small].max() < 4e-16
    assert np.abs(s - np.sin(xs))[~small].max() < 1e-15
    # after the single rounding to Float32 the result is the correctly rounded one (here: always)
    assert np.array_equal(s.astype(f32), np.sin(xs).astype(f32)) and np.array_equal(c.astype(f32), np.cos(xs).astype(f32))
    bad = f32([np.nan, np.inf])
    HC.hc_sincos(fp(bad), 2, s.ctypes.data_as(_D), c.ctypes.data_as(_D))
    assert np.isnan(s[:2]).all() and np.isnan(c[:2]).all()


def test_polynomial_atan2(HC):
    rng = np.random.default_rng(12)
    n = 300000
    y, x = rng.standard_normal(n) * 10.0 ** rng.integers(-3, 4, n), rng.standard_normal(n) * 10.0 ** rng.integers(-3, 4, n)
    sp_y = np.array([0.0, -0.0, 0.0, -0.0, 1.0, -1.0, 1.0, -1.0, 0.0, 0.0, 1.0, -1.0, np.inf, -np.inf, np.inf, np.inf, 1.0, 1e-300, 1.0])
    sp_x = np.array([0.0, 0.0, -0.0, -0.0, 0.0, 0.0, -0.0, -0.0, 1.0, -1.0, 1.0, -1.0, np.inf, np.inf, -np.inf, 1.0, np.inf, 1e-300, -np.inf])
    y, x = np.concatenate([y, sp_y]), np.concatenate([x, sp_x])
    out = np.empty(len(y))
    HC.hc_atan2(y.ctypes.data_as(_D), x.ctypes.data_as(_D), len(y), out.ctypes.data_as(_D))
    ref = np.arctan2(y, x)
    assert np.abs(out - ref).max() < 2e-15
    assert np.array_equal(np.signbit(out), np.signbit(ref))                       # signed zeros: atan2(-0, +1) = -0 ...
    assert np.array_equal(out.astype(f32), ref.astype(f32))                       # the rounded Float32 is the correctly rounded one
    small = np.abs(ref) < 1e-3
    sel = small & (ref != 0)
    assert (np.abs(out[sel] / ref[sel] - 1) < 3e-15).all()              # relative accuracy near zero
    nan = np.array([np.nan, 1.0, np.nan])
    HC.hc_atan2(nan.ctypes.data_as(_D), np.array([1.0, np.nan, np.nan]).ctypes.data_as(_D), 3, out.ctypes.data_as(_D))
    assert np.isnan(out[:3]).all()


@pytest.mark.parametrize("vec4,corr", [(0, -1), (1, -1), (1, 0), (0, 3)])
@pytest.mark.parametrize("kw", [dict(z_tab_max=3), dict(z_tab_max=3, z_tab_num=37), dict(z_tab_max=3, z_tab_num=2049),
                                dict(z_tab_max=10), dict(z_tab_max=3, z_tab_num=1 << 20), dict(z_tab_min=0.4, z_tab_max=1.6, z_tab_num=3)])
def test_device_arithmetic_of_cartesian_to_sky_and_round_trip(HC, B, kw, vec4, corr):
    """corr = -1: the correction steps measured from the table (the product's setting); 0: every guess that misses its
    interval falls through to the bisection; 3: more steps than needed.  All give the same answer."""
    cosmo, ref = B.Cosmology(**kw), CO.Cosmology(**kw)
    z, r = product_table(B, cosmo)
    zmin, zmax = float(kw.get("z_tab_min", 0.0)), float(kw["z_tab_max"])
    ra, dec, red = sky_catalog(50001, 4, 1.0)
    red = (zmin + (zmax - zmin) * np.clip(red, 0.003, 0.9999)).astype(f32)
    if zmin == 0:
        red[8] = f32(0.0)                                 # the first knot itself
    x, y, zz = CO.sky_to_cartesian(ra, dec, red, ref)
    a, d, q = (np.empty_like(ra) for _ in range(3))
    measured = C.c_int(-1)
    bad = HC.hc_cartesian_to_sky(fp(x), fp(y), fp(zz), len(x), f32(cosmo.h), r.ctypes.data_as(_D), z[0], (z[-1] - z[0]) / (len(z) - 1), len(z), fp(a), fp(d), fp(q), vec4, corr, C.byref(measured))
    assert bad == 0 and (measured.value == 1 or len(z) > 200000 or len(z) < 100)
    oa, od, oq = CO.cartesian_to_sky(x, y, zz, ref)
    assert ulps(a, oa).max() <= 1 and ulps(d, od).max() <= 1 and ulps(q, oq).max() <= 1
    assert (a <= 0).all() and (a > -360).all()
    sel = (np.abs(dec) < 89) & (red > 0)
    coarse_tab = len(z) < 1000          # linear interpolation error of a 37-knot table dominates there
    assert np.abs(((a - ra + 180) % 360) - 180)[sel].max() < 2e-4
    assert np.abs(q[red > 0] / red[red > 0] - 1).max() < (2e-3 if coarse_tab else 3e-6)


@pytest.mark.parametrize("vec4", [0, 1])
def test_out_of_table_is_counted_and_nan(HC, B, vec4):
    cosmo = B.Cosmology(z_tab_max=1, z_tab_num=101)
    z, r = product_table(B, cosmo)
    ra, dec, red = f32([10, 20, 30, 40]), f32([1, 2, 3, 4]), f32([0.5, 1.0000001, -0.1, np.nan])
    x, y, zz = (np.empty_like(ra) for _ in range(3))
    bad = HC.hc_sky_to_cartesian(fp(ra), fp(dec), fp(red), 4, f32(0.67), r.ctypes.data_as(_D), z[0], z[-1], (z[-1] - z[0]) / (len(z) - 1), len(z), fp(x), fp(y), fp(zz))
    assert bad == 3 and np.isfinite(x[0]) and np.isnan(x[1:]).all() and np.isnan(zz[1:]).all()
    far = f32([r[-1] * 0.67 * 1.01, 100.0])
    a, d, q = (np.empty_like(far) for _ in range(3))
    bad = HC.hc_cartesian_to_sky(fp(far), fp(f32([0, 0])), fp(f32([0, 0])), 2, f32(0.67), r.ctypes.data_as(_D), z[0], (z[-1] - z[0]) / (len(z) - 1), len(z), fp(a), fp(d), fp(q), vec4, -1, None)
    assert bad == 1 and np.isnan(q[0]) and np.isfinite(q[1]) and np.isfinite(a).all()


def test_fkp_and_wrap_bit_exact(HC):
    rng = np.random.default_rng(5)
    nz = (1e-3 * rng.random(10000)).astype(f32)
    w = np.empty_like(nz)
    HC.hc_fkp_weights(fp(nz), len(nz), f32(5e3), fp(w))
    assert np.array_equal(w.view(np.uint32), CO.fkp_weights(nz, f32(5e3)).view(np.uint32))
    for L, mn in (((1000.0,) * 3, (0.0,) * 3), ((500.0, 750.0, 1250.0), (-250.0, 10.0, 1e3))):
        pos = [(m - 0.3 * l + 1.6 * l * rng.random(20000)).astype(f32) for l, m in zip(L, mn)]
        pos[0][:5] = f32([mn[0], mn[0] + L[0], mn[0] - L[0], mn[0] + 1e-6, mn[0] - 1e-6])
        ref = CO.wrap_positions(*pos, L, mn)
        got = [p.copy() for p in pos]
        HC.hc_wrap_positions(fp(got[0]), fp(got[1]), fp(got[2]), len(got[0]), fp(f32(L)), fp(f32(mn)))
        for g, o, l, m in zip(got, ref, L, mn):
            assert np.array_equal(g.view(np.uint32), o.view(np.uint32))
            assert (g >= f32(m)).all() and (g <= f32(m) + f32(l)).all()


def test_device_arithmetic_against_the_golden_fixture(HC, B):
    """The committed catalog fixture (tests/golden/catalog_4000.npz) against the kernels' arithmetic, no oracle run."""
    with np.load(ROOT / "tests" / "golden" / "catalog_4000.npz") as zf:
        g = {k: zf[k] for k in zf.files}
    cosmo = B.Cosmology(z_tab_max=3)
    z, r = product_table(B, cosmo)
    assert np.abs(r[::1000][1:] / g["r_tab"][1:] - 1).max() < 1e-12
    n = len(g["ra"])
    x, y, zz = (np.empty(n, f32) for _ in range(3))
    assert HC.hc_sky_to_cartesian(fp(g["ra"]), fp(g["dec"]), fp(g["red"]), n, f32(cosmo.H0 / f32(100)), r.ctypes.data_as(_D), z[0], z[-1],
                                  (z[-1] - z[0]) / (len(z) - 1), len(z), fp(x), fp(y), fp(zz)) == 0
    scale = np.sqrt(g["x"].astype(float) ** 2 + g["y"].astype(float) ** 2 + g["z"].astype(float) ** 2).astype(f32)
    for a, b in ((x, g["x"]), (y, g["y"]), (zz, g["z"])):
        assert (np.abs(a.astype(float) - b.astype(float)) <= 1.01 * np.spacing(scale)).all() and (a == b).mean() > 0.99
    a, d, q = (np.empty(n, f32) for _ in range(3))
    assert HC.hc_cartesian_to_sky(fp(g["x"]), fp(g["y"]), fp(g["z"]), n, f32(cosmo.h), r.ctypes.data_as(_D), z[0],
                                  (z[-1] - z[0]) / (len(z) - 1), len(z), fp(a), fp(d), fp(q), 1, -1, None) == 0
    assert ulps(a, g["ra2"]).max() <= 1 and ulps(d, g["dec2"]).max() <= 1 and ulps(q, g["red2"]).max() <= 1
    w = np.empty(n, f32)
    HC.hc_fkp_weights(fp(g["nz"]), n, f32(5e3), fp(w))
    assert np.array_equal(w.view(np.uint32), g["fkp"].view(np.uint32))
    p = [g[k].copy() for k in "xyz"]
    HC.hc_wrap_positions(fp(p[0]), fp(p[1]), fp(p[2]), n, fp(g["wrap_box_size"]), fp(g["wrap_box_min"]))
    for got, k in zip(p, ("wx", "wy", "wz")):
        assert np.array_equal(got.view(np.uint32), g[k].view(np.uint32))


@pytest.mark.parametrize("kw", [dict(), dict(z_tab_max=10), dict(w0=-0.9, wa=0.1, z_tab_max=2), dict(z_tab_min=0.4, z_tab_max=1.6, z_tab_num=4097)])
def test_scalar_cosmology_functions_mirror_the_reference(B, kw):
    """E, H, comoving_distance and the two interpolators of src/cosmo.jl:70-98 as host callables of the product mirror
    (table from the library's own quadrature) against the oracle's restatement (scipy quad, rtol 1e-10)."""
    mine, ref = B.Cosmology(**kw), CO.Cosmology(**kw)
    z = np.array([0.0, 1e-3, 0.5, 0.8754, 1.0, 1.5999])
    assert np.allclose(B.E(mine, z), CO.E(ref, z), rtol=1e-14, atol=0)
    assert np.allclose(B.H(mine, z), CO.H(ref, z), rtol=1e-14, atol=0)
    for zi in (0.0, 0.3, 1.0, 2.5):
        assert B.comoving_distance(mine, zi) == pytest.approx(CO.comoving_distance(ref, zi), rel=1e-9, abs=0)
    r_fun, z_fun = B.comoving_distance_interp(mine), B.redshift_interp(mine)
    r_ref, z_ref = CO.comoving_distance_interp(ref), CO.redshift_interp(ref)
    zq = np.linspace(max(float(mine.z_tab_min), 0.41), 1.59, 1001)
    assert np.allclose(r_fun(zq), r_ref(zq), rtol=1e-12, atol=0)
    rq = np.asarray(r_ref(zq))
    assert np.allclose(z_fun(rq), z_ref(rq), rtol=1e-11, atol=1e-13)
    assert np.allclose(z_fun(r_fun(zq)), zq, rtol=0, atol=1e-9)               # the two tables invert each other
    assert float(r_fun(0.9)) == pytest.approx(float(r_ref(0.9)), rel=1e-12)   # scalars work too
    with pytest.raises(B.OutOfRangeError):                                    # Interpolations.jl: BoundsError
        r_fun(float(mine.z_tab_max) + 1.0)
    with pytest.raises(B.OutOfRangeError):
        z_fun(-1.0)
