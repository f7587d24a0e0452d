"""GPU parity: catalog pre/post-processing kernels (baorec.jl_b200/csrc/catalog.cu) through the C ABI
against oracle/catalog_oracle.py (src/cosmo.jl:70-103, examples/lightcone.jl:30-82).
Bit-exact: FKP weights and the periodic re-wrap (Float32-only formulas).  Conversions: <= 2 ulp of
|position| / of the angle / of the redshift (the Float64 sincos / atan2 of CUDA's and the host's libm
may differ in the last place before the single rounding to Float32).
(File name sorts last on purpose: these kernels were written after this round's GPU budget was spent,
so under `pytest -x` everything validated earlier runs first.)"""
import numpy as np
import pytest

import catalog_oracle as CO

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu
f32 = np.float32


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def host(t):
    return t.cpu().numpy()


def ulps(a, b, scale=None):
    a, b = np.asarray(a, f32), np.asarray(b, f32)
    s = np.maximum(np.abs(a), np.abs(b)).astype(f32) if scale is None else np.asarray(scale, f32)
    return np.abs(a.astype(np.float64) - b.astype(np.float64)) / np.spacing(s)


def sky_catalog(n, seed, zmax):
    rng = np.random.default_rng(seed)
    ra, dec = (360 * rng.random(n)).astype(f32), (180 * rng.random(n) - 90).astype(f32)
    red = (zmax * rng.random(n)).astype(f32)
    ra[:4], dec[:4] = f32([0, 90, 180, 359.99]), f32([0, -90, 90, 45])
    red[:3] = f32([0, zmax, zmax / 2])
    return ra, dec, red


@pytest.mark.parametrize("kw", [dict(z_tab_max=3), dict(z_tab_max=10), dict(z_tab_min=0.4, z_tab_max=1.6, z_tab_num=4097)])
def test_tables_on_the_device_context(B, kw):
    cosmo, ref = B.Cosmology(**kw), CO.Cosmology(**kw)
    z, r = cosmo.tables()
    zo, ro = CO.tables(ref)
    assert np.abs(z - zo).max() < 1e-13 and np.abs(r[1:] / ro[1:] - 1).max() < 1e-12


@pytest.mark.parametrize("kw,n", [(dict(z_tab_max=3), 300_000), (dict(z_tab_max=10), 70_001), (dict(z_tab_max=2, z_tab_num=1500), 1000)])
def test_sky_to_cartesian(B, kw, n):
    cosmo, ref = B.Cosmology(**kw), CO.Cosmology(**kw)
    ra, dec, red = sky_catalog(n, 3, float(kw["z_tab_max"]))
    x, y, z = (host(t) for t in B.sky_to_cartesian(dev(ra), dev(dec), dev(red), cosmo))
    ox, oy, oz = CO.sky_to_cartesian(ra, dec, red, ref)
    scale = np.maximum(np.sqrt(ox.astype(float) ** 2 + oy.astype(float) ** 2 + oz.astype(float) ** 2), 1e-30).astype(f32)
    for a, b in ((x, ox), (y, oy), (z, oz)):
        assert ulps(a, b, scale).max() <= 2
        assert (a == b).mean() > 0.98


def test_cartesian_to_sky_and_round_trip(B):
    kw = dict(z_tab_max=3)
    cosmo, ref = B.Cosmology(**kw), CO.Cosmology(**kw)
    ra, dec, red = sky_catalog(200_000, 4, 2.95)
    red = np.maximum(red, f32(0.01))
    gx, gy, gz = B.sky_to_cartesian(dev(ra), dev(dec), dev(red), cosmo)
    a, d, q = (host(t) for t in B.cartesian_to_sky(gx, gy, gz, cosmo))
    oa, od, oq = CO.cartesian_to_sky(host(gx), host(gy), host(gz), ref)
    assert ulps(a, oa).max() <= 2 and ulps(d, od, np.maximum(np.abs(od), 1e-3)).max() <= 2 and ulps(q, oq).max() <= 2
    assert (a <= 0).all() and (a > -360).all()                     # the reference's `(lon - 360) % 360`
    ok = np.abs(dec) < 89
    assert np.abs(((a - ra + 180) % 360) - 180)[ok].max() < 2e-4 and np.abs(d - dec).max() < 2e-4
    assert np.abs(q / red - 1).max() < 3e-6


def test_inverse_interpolation_does_not_depend_on_the_guess(B):
    """cartesian_to_sky finds its table interval from a guess table + measured number of +-1 steps; with the steps
    forced to 0 every miss falls through to the bisection of the whole table.  Same redshifts bit for bit."""
    cosmo = B.Cosmology(z_tab_max=10)
    ra, dec, red = sky_catalog(100_003, 7, 9.9)
    g = B.sky_to_cartesian(dev(ra), dev(dec), dev(np.maximum(red, f32(1e-3))), cosmo)
    ctx = B.Context.get(0)
    base = [host(t) for t in B.cartesian_to_sky(*g, cosmo)]
    try:
        for corr in (0, 5):
            ctx.set_option("catalog_corr", corr)
            for a, b in zip(base, [host(t) for t in B.cartesian_to_sky(*g, cosmo)]):
                assert np.array_equal(a.view(np.uint32), b.view(np.uint32))
    finally:
        ctx.set_option("catalog_corr", -1)


def test_out_of_table_raises_and_marks_nan(B):
    cosmo = B.Cosmology(z_tab_max=1, z_tab_num=101)
    ra, dec, red = f32([10, 20, 30, 40]), f32([1, 2, 3, 4]), f32([0.5, 1.0000001, -0.1, np.nan])
    with pytest.raises(B.OutOfRangeError):
        B.sky_to_cartesian(dev(ra), dev(dec), dev(red), cosmo)
    x, _, _ = B.sky_to_cartesian(dev(ra[:1]), dev(dec[:1]), dev(red[:1]), cosmo)      # the context stays usable
    assert np.isfinite(host(x)).all()
    _, r = cosmo.tables()
    far = f32([r[-1] * 0.67 * 1.01])
    with pytest.raises(B.OutOfRangeError):
        B.cartesian_to_sky(dev(far), dev(f32([0])), dev(f32([0])), cosmo)


def test_switching_cosmologies_rebuilds_the_table(B):
    a, b = B.Cosmology(z_tab_max=3), B.Cosmology(h=0.7, z_tab_max=3, z_tab_num=5000)
    ra, dec, red = sky_catalog(5000, 6, 2.9)
    for cosmo in (a, b, a):
        ref = CO.Cosmology(h=float(cosmo.h), z_tab_max=3, z_tab_num=int(cosmo.z_tab_num))
        x, _, _ = B.sky_to_cartesian(dev(ra), dev(dec), dev(red), cosmo)
        ox, oy, oz = CO.sky_to_cartesian(ra, dec, red, ref)
        scale = np.sqrt(ox.astype(float) ** 2 + oy.astype(float) ** 2 + oz.astype(float) ** 2).astype(f32)
        assert ulps(host(x), ox, np.maximum(scale, 1e-30)).max() <= 2


def test_fkp_weights_and_wrap_bit_exact(B):
    rng = np.random.default_rng(5)
    nz = (1e-3 * rng.random(100_000)).astype(f32)
    w = host(B.fkp_weights(dev(nz), 5e3))
    assert np.array_equal(w.view(np.uint32), CO.fkp_weights(nz, f32(5e3)).view(np.uint32))
    for L, mn in (((1000.0,) * 3, (0.0,) * 3), ((500.0, 750.0, 1250.0), (-250.0, 10.0, 1e3))):
        pos = [(m - 0.3 * l + 1.6 * l * rng.random(200_000)).astype(f32) for l, m in zip(L, mn)]
        pos[0][:5] = f32([mn[0], mn[0] + L[0], mn[0] - L[0], mn[0] + 1e-6, mn[0] - 1e-6])
        ref = CO.wrap_positions(*pos, L, mn)
        g = [dev(p) for p in pos]
        out = B.wrap_positions(*g, L, mn)
        assert out[0] is g[0]                                        # in place
        for t, o in zip(g, ref):
            assert np.array_equal(host(t).view(np.uint32), o.view(np.uint32))


def test_unaligned_arrays_and_tails_take_the_scalar_path(B):
    """16-byte aligned arrays go through the float4 kernels (+ a scalar launch for the <= 3 tail particles);
    anything else through the scalar instantiation: same results bit for bit."""
    cosmo = B.Cosmology(z_tab_max=3)
    ra, dec, red = sky_catalog(10_007, 9, 2.9)
    nz = (1e-3 * np.random.default_rng(2).random(10_007)).astype(f32)
    base = [host(t) for t in B.sky_to_cartesian(dev(ra), dev(dec), dev(red), cosmo)]
    pad = lambda a: dev(np.concatenate([f32([0.5]), a]))[1:]          # contiguous, 4 bytes off a 16-byte boundary
    off = [host(t) for t in B.sky_to_cartesian(pad(ra), pad(dec), pad(red), cosmo)]
    for a, b in zip(base, off):
        assert np.array_equal(a.view(np.uint32), b.view(np.uint32))
    sky = [host(t) for t in B.cartesian_to_sky(*(dev(a) for a in base), cosmo)]
    sky_off = [host(t) for t in B.cartesian_to_sky(*(pad(a) for a in base), cosmo)]
    for a, b in zip(sky, sky_off):
        assert np.array_equal(a.view(np.uint32), b.view(np.uint32))
    assert np.array_equal(host(B.fkp_weights(pad(nz), 5e3)), host(B.fkp_weights(dev(nz), 5e3)))
    w0, w1 = [dev(a) for a in base], [pad(a) for a in base]
    B.wrap_positions(*w0, (1000.0,) * 3)
    B.wrap_positions(*w1, (1000.0,) * 3)
    for a, b in zip(w0, w1):
        assert np.array_equal(host(a), host(b)) and float(a.min()) >= 0 and float(a.max()) <= 1000


def test_empty_catalogs(B):
    e = torch.empty(0, dtype=torch.float32, device="cuda")
    cosmo = B.Cosmology()
    assert all(t.numel() == 0 for t in B.sky_to_cartesian(e, e, e, cosmo))
    assert all(t.numel() == 0 for t in B.cartesian_to_sky(e, e, e, cosmo))
    assert B.fkp_weights(e, 5e3).numel() == 0
    B.wrap_positions(e, e, e, (1.0, 1.0, 1.0))


def test_lightcone_example_flow(B, O):
    """examples/lightcone.jl end to end on the device: sky -> Cartesian, FKP weights, run! with randoms,
    reconstructed_positions, Cartesian -> sky; against the oracle doing the same on the CPU."""
    rng = np.random.default_rng(8)
    cosmo, ref = B.Cosmology(z_tab_max=3), CO.Cosmology(z_tab_max=3)

    def cat(n):
        ra, dec = (20 + 25 * rng.random(n)).astype(f32), (-10 + 25 * rng.random(n)).astype(f32)
        return ra, dec, (0.8 + 0.2 * rng.random(n)).astype(f32), (2e-4 * (0.5 + rng.random(n))).astype(f32)

    (ra, dec, red, nz), (rra, rdec, rred, rnz) = cat(40_000), cat(400_000)
    n = 64
    kw = dict(bias=2.2, f=0.757, smoothing_radius=15.0, los=None, n_iter=3)
    # oracle
    od, orr = CO.sky_to_cartesian(ra, dec, red, ref), CO.sky_to_cartesian(rra, rdec, rred, ref)
    ow, orw = CO.fkp_weights(nz, f32(5e3)), CO.fkp_weights(rnz, f32(5e3))
    # device
    gd = B.sky_to_cartesian(dev(ra), dev(dec), dev(red), cosmo)
    gr = B.sky_to_cartesian(dev(rra), dev(rdec), dev(rred), cosmo)
    gw, grw = B.fkp_weights(dev(nz), 5e3), B.fkp_weights(dev(rnz), 5e3)
    assert np.array_equal(host(gw), ow)
    # the `ran > threshold` decisions of the device set-up (a discontinuity: DESIGN.md section 5) are handed to the oracle
    rec = B.IterativeRecon(**kw)
    rec.box_size, rec.box_min = B.setup_box(*gr, 500.0)
    delta = torch.zeros((n, n, n), dtype=torch.float32, device="cuda")
    B.setup_fft(rec, delta)
    B.setup_overdensity(delta, rec, *gd, gw, *gr, grw)
    mask = host(delta) != 0
    rec = B.IterativeRecon(**kw)
    mesh = B.run(rec, (n, n, n), *gd, gw, *gr, grw)
    orec = O.IterativeRecon(**kw)
    # the oracle runs on the device's own Cartesian catalog (differences there are <= 2 ulp, tested above) and mask
    hd, hr = [host(t) for t in gd], [host(t) for t in gr]
    omesh = O.run(orec, (n, n, n), *hd, ow, *hr, orw, force_mask=mask)
    assert np.array_equal(rec.box_size, orec.box_size)
    new = B.reconstructed_positions(rec, *gd, field="sum")
    onew = O.reconstructed_positions(orec, *hd, omesh, "sum")
    for a in range(3):
        err = np.abs(host(new[a]) - onew[a])
        assert np.median(err) < 1e-3 and err.max() < 5e-2          # positions ~2e3 Mpc/h: Float32 spacing is 1.2e-4
    sky, osky = B.cartesian_to_sky(*new, cosmo), CO.cartesian_to_sky(*onew, ref)
    assert np.abs(host(sky[0]) - osky[0]).max() < 1e-3 and np.abs(host(sky[1]) - osky[1]).max() < 1e-3
    assert np.abs(host(sky[2]) - osky[2]).max() < 1e-4
    for t in od + orr:
        assert t.dtype == np.float32
