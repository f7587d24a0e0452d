"""CPU: known-answer tests pinning oracle/catalog_oracle.py (cosmology tables, sky <-> Cartesian, FKP
weights, periodic re-wrap; src/cosmo.jl, examples/lightcone.jl:30-82).  The reference has no tests or
vectors for these helpers, so they are pinned analytically."""
import numpy as np
import pytest

import catalog_oracle as CO

f32 = np.float32


def eds():
    """Einstein-de Sitter: Omega_m = 1, nothing else -> r(z) = (2c/H0) (1 - 1/sqrt(1+z))."""
    c = CO.Cosmology(z_tab_max=5, z_tab_num=5001)
    c.Omega_g0 = c.Omega_nu0 = c.Omega_k0 = c.Omega_L0 = f32(0)
    c.Omega_b0, c.Omega_c0 = f32(0.25), f32(0.75)
    return c


def test_derived_densities_follow_the_struct():
    c = CO.Cosmology()
    assert c.H0 == f32(f32(0.67) * f32(100)) and c.h2 == f32(f32(0.67) * f32(0.67))
    assert c.Omega_b0 == f32(f32(0.0225) / c.h2) and c.Omega_c0 == f32(f32(0.12) / c.h2)
    # photons: T_cmb = 2.725 K, h = 0.67 -> Omega_gamma h^2 = 2.47e-5; 3.044 neutrinos x 0.2271
    assert abs(float(c.Omega_g0) * float(c.h2) / 2.47e-5 - 1) < 5e-3
    assert abs(float(c.Omega_nu0) / float(c.Omega_g0) - 3.044 * 0.22711) < 1e-4
    tot = sum(float(v) for v in (c.Omega_b0, c.Omega_c0, c.Omega_g0, c.Omega_nu0, c.Omega_k0, c.Omega_L0))
    assert abs(tot - 1.0) < 3e-7                      # flat by construction, in Float32
    d = CO.DESICosmology()
    assert d.h == f32(0.6736) and d.Neff == f32(f32(2.0328) + f32(1))


def test_einstein_de_sitter_closed_form():
    c = eds()
    H0 = float(f32(c.h * f32(100)))
    for z in (0.1, 0.5, 1.0, 3.0):
        exact = 2 * CO.speed_of_light_km_s / H0 * (1 - 1 / np.sqrt(1 + z))
        assert abs(CO.comoving_distance(c, z) / exact - 1) < 1e-9
    zt, rt = CO.tables(c)
    exact = 2 * CO.speed_of_light_km_s / H0 * (1 - 1 / np.sqrt(1 + zt))
    assert np.abs(rt[1:] / exact[1:] - 1).max() < 1e-12 and rt[0] == 0.0


def test_table_matches_adaptive_quadrature_and_its_derivative_is_c_over_H():
    c = CO.Cosmology(z_tab_max=10)
    zt, rt = CO.tables(c)
    assert len(zt) == 100000 and zt[0] == 0 and zt[-1] == 10
    for i in (1, 777, 31415, 99999):
        assert abs(rt[i] / CO.comoving_distance(c, zt[i]) - 1) < 1e-9      # the reference: quadgk, rtol 1e-8
    mid = 0.5 * (zt[1:] + zt[:-1])
    slope = np.diff(rt) / np.diff(zt)
    assert np.abs(slope * CO.H(c, mid) / CO.speed_of_light_km_s - 1).max() < 1e-8


def test_nonzero_table_start():
    a, b = CO.Cosmology(z_tab_min=0.5, z_tab_max=1.5, z_tab_num=1001), CO.Cosmology(z_tab_max=1.5, z_tab_num=1501)
    (za, ra), (zb, rb) = CO.tables(a), CO.tables(b)
    assert np.allclose(za, zb[500:], atol=1e-13) and np.abs(ra / rb[500:] - 1).max() < 1e-10


def test_gridded_linear_interpolation():
    c = CO.Cosmology(z_tab_max=2, z_tab_num=21)
    zt, rt = CO.tables(c)
    r_fun, z_fun = CO.comoving_distance_interp(c), CO.redshift_interp(c)
    assert np.array_equal(r_fun(zt), rt) and np.allclose(z_fun(rt), zt, rtol=0, atol=1e-15)
    assert r_fun(0.25 * zt[3] + 0.75 * zt[4]) == pytest.approx(0.25 * rt[3] + 0.75 * rt[4], rel=1e-14)
    assert z_fun(0.5 * (rt[7] + rt[8])) == pytest.approx(0.5 * (zt[7] + zt[8]), rel=1e-14)
    for bad in (-1e-9, 2.0000001, np.nan):
        with pytest.raises(CO.OutOfTableError):
            r_fun(np.array([0.5, bad]))
    with pytest.raises(CO.OutOfTableError):
        z_fun(np.array([rt[-1] * 1.0000001]))


def test_sky_to_cartesian_known_directions():
    c = CO.Cosmology()
    r1 = CO.comoving_distance_interp(c)(np.float64(f32(1.0))) * float(f32(c.H0 / f32(100)))
    ra, dec, red = f32([0, 90, 180, 0, 45]), f32([0, 0, 0, 90, -30]), np.full(5, 1.0, f32)
    x, y, z = CO.sky_to_cartesian(ra, dec, red, c)
    assert x.dtype == np.float32
    d = np.sqrt(x.astype(float) ** 2 + y.astype(float) ** 2 + z.astype(float) ** 2)
    assert np.abs(d / r1 - 1).max() < 2e-7
    assert abs(x[0] / r1 - 1) < 1e-7 and abs(y[1] / r1 - 1) < 1e-7 and abs(x[2] / r1 + 1) < 1e-7 and abs(z[3] / r1 - 1) < 1e-7
    assert abs(z[4] / r1 + 0.5) < 1e-7 and abs(x[4] / y[4] - 1) < 1e-6
    # Float32(pi) is not pi: cos(Float32(pi)/2) = -4.37e-8, kept (not "fixed") like the reference
    assert abs(x[1]) < 1e-3 and x[1] != 0


def test_round_trip_and_the_ra_quirk():
    rng = np.random.default_rng(1)
    c = CO.Cosmology(z_tab_max=3)
    ra, dec = (360 * rng.random(2000)).astype(f32), (180 * rng.random(2000) - 90).astype(f32)
    red = (0.05 + 2.9 * rng.random(2000)).astype(f32)
    x, y, z = CO.sky_to_cartesian(ra, dec, red, c)
    ra2, dec2, red2 = CO.cartesian_to_sky(x, y, z, c)
    assert (ra2 <= 0).all() and (ra2 > -360).all()               # `(lon - 360) % 360`, truncated remainder
    assert np.abs(((ra2 - ra + 180) % 360) - 180).max() < 1e-4   # same direction modulo 360
    assert np.abs(dec2 - dec).max() < 1e-4 and np.abs(red2 / red - 1).max() < 3e-6


def test_fkp_weights_and_wrap():
    nz = f32([0, 1e-4, 2e-4, 5e-3])
    assert np.array_equal(CO.fkp_weights(nz, 5e3), (f32(1) / (f32(1) + nz * f32(5e3))).astype(f32))
    assert CO.fkp_weights(nz, 5e3)[0] == 1 and abs(CO.fkp_weights(nz, 5e3)[2] - 0.5) < 1e-7
    x, y, z = f32([-1, 0, 999.5, 1000, 1001.25]), f32([5, 5, 5, 5, 5]), f32([-0.25, 1e-8, 500, 1999, -999])
    wx, wy, wz = CO.wrap_positions(x, y, z, (1000, 1000, 1000))
    assert np.array_equal(wx, f32([999, 0, 999.5, 0, 1.25])) and np.array_equal(wy, y)
    assert np.array_equal(wz, f32([999.75, 0, 500, 999, 1]))      # 1e-8 + 1000 rounds to 1000 in Float32: (pos + L) % L as written
    sx, _, _ = CO.wrap_positions(f32([-260, 240, 10]), y[:3], y[:3], (500, 500, 500), (-250, -250, -250))
    assert np.array_equal(sx, f32([240, -250 + 490, 10]))
