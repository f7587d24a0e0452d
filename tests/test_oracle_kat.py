"""Known-answer tests that pin the CPU oracle (oracle/baorec_oracle.py).

The reference ships no usable golden vectors (its CSVs come from DESI mocks that are not
shipped; Julia is not installed), so the oracle is pinned analytically: plane waves with closed
forms, conservation laws, beta = 0 identities, discrete-operator eigenfunctions."""
import numpy as np
import pytest

import baorec_oracle as O


def box(n, L, lo=0.0, T=np.float64):
    return np.full(3, L, T), np.full(3, lo, T)


def plane_wave(n, L, m, T=np.float64, phase=0.3):
    """delta(x) = cos(k.x + phase) sampled on mesh points x_i = i L/n; m = integer mode numbers (mx,my,mz)."""
    i = np.arange(n) * (L / n)
    kx, ky, kz = (2 * np.pi * mm / L for mm in m)
    arg = kz * i[:, None, None] + ky * i[None, :, None] + kx * i[None, None, :] + phase
    return np.cos(arg).astype(T), np.sin(arg).astype(T), np.array([kx, ky, kz])


# ---- utils.jl ---------------------------------------------------------------------------------
@pytest.mark.parametrize("n", [8, 12, 64])
def test_k_vec_matches_fftfreq(n):
    L = 1000.0
    kx, ky, kz = O.k_vec((n, n, n), np.full(3, L, np.float64), np.float64)
    assert np.allclose(kx, 2 * np.pi * np.fft.rfftfreq(n, d=L / n))
    assert np.allclose(ky, 2 * np.pi * np.fft.fftfreq(n, d=L / n))
    assert kx[-1] > 0 and ky[n // 2] < 0 and kz[n // 2] < 0       # Nyquist: + for x, - for y,z


def test_k_vec_float32_rounding_path():
    # fs = Float32(2 pi n / L) from a Float64 product, multiplier fs/n in Float32, value = i * multiplier
    n, L = 256, np.float32(2500.0)
    kx, _, _ = O.k_vec((n, n, n), np.full(3, L, np.float32), np.float32)
    fs = np.float32(2.0 * np.pi * n / np.float64(L))
    mult = np.float32(fs / np.float32(n))
    assert kx.dtype == np.float32 and kx[17] == np.float32(17) * mult


def test_x_vec_cell_centres():
    xv = O.x_vec((4, 8, 16), np.array([8.0, 8.0, 8.0]), np.array([-4.0, 0.0, 10.0]), np.float64)
    assert np.allclose(xv[0], [-3, -1, 1, 3]) and np.allclose(xv[1], 0.5 + np.arange(8)) and len(xv[2]) == 16
    assert np.isclose(xv[2][0], 10.25)


def test_setup_box():
    x, y, z = np.array([1.0, 9.0]), np.array([-5.0, 5.0]), np.array([0.0, 2.0])
    size, mn = O.setup_box(x, y, z, 500.0)
    assert np.allclose(mn, [-249, -255, -250]) and np.allclose(size, 510.0)     # cubic: largest extent + pad


def test_smooth_plane_wave():
    n, L, R = 32, 100.0, 7.0
    c, _, k = plane_wave(n, L, (2, 1, 3))
    out = O.smooth(c.copy(), R, np.full(3, L))
    assert np.allclose(out, np.exp(-0.5 * R * R * (k @ k)) * c, atol=1e-12)


# ---- mas.jl -----------------------------------------------------------------------------------
def test_cic_particle_on_grid_point_and_mass_conservation():
    n, L = 8, 16.0
    bs, bm = box(n, L)
    rho = np.zeros((n, n, n))
    O.cic_scatter(rho, np.array([4.0]), np.array([6.0]), np.array([14.0]), np.array([2.5]), bs, bm)
    assert rho[7, 3, 2] == 2.5 and rho.sum() == 2.5              # [iz, iy, ix] = (14, 6, 4)/2
    rng = np.random.default_rng(0)
    p = [rng.uniform(0, L, 1000) for _ in range(3)]
    w = rng.uniform(0.5, 2, 1000)
    rho = O.cic_scatter(np.zeros((n, n, n)), *p, w, bs, bm)
    assert np.isclose(rho.sum(), w.sum())


def test_cic_linear_weights_and_periodic_wrap():
    n, L = 8, 8.0
    bs, bm = box(n, L)
    rho = O.cic_scatter(np.zeros((n, n, n)), np.array([7.25]), np.array([0.0]), np.array([0.0]), np.array([1.0]), bs, bm)
    assert np.isclose(rho[0, 0, 7], 0.75) and np.isclose(rho[0, 0, 0], 0.25)     # wraps to cell 0
    # wrap quirk (src/mas.jl:8-10): p - min > L  ->  p - L, written back into the caller's array
    x = np.array([8.5])
    O.cic_scatter(np.zeros((n, n, n)), x, np.array([1.0]), np.array([1.0]), np.array([1.0]), bs, bm)
    assert x[0] == 0.5
    with pytest.raises(O.OutOfBoxError):
        O.cic_scatter(np.zeros((n, n, n)), np.array([-0.5]), np.array([1.0]), np.array([1.0]), np.array([1.0]), bs, bm)
    with pytest.raises(O.OutOfBoxError):      # wrap=false and base cell == n
        O.cic_scatter(np.zeros((n, n, n)), np.array([7.5]), np.array([1.0]), np.array([1.0]), np.array([1.0]), bs, bm,
                      wrap=False)


def test_gather_reproduces_trilinear_field_and_adjointness():
    n, L = 16, 16.0
    bs, bm = box(n, L)
    i = np.arange(n, dtype=np.float64)
    fld = (2.0 * i[None, None, :] + 3.0 * i[None, :, None] - 1.0 * i[:, None, None]) + 5.0   # linear in x,y,z
    rng = np.random.default_rng(1)
    p = [rng.uniform(0, L - 1.0, 500) for _ in range(3)]         # stay off the periodic seam
    got = O.read_cic(fld, *p, bs, bm)
    assert np.allclose(got, 2 * p[0] + 3 * p[1] - p[2] + 5.0)
    # scatter and gather are adjoint for in-box particles: <scatter(w), f> = <w, gather(f)>
    w = rng.uniform(0, 1, 500)
    f = rng.standard_normal((n, n, n))
    lhs = (O.cic_scatter(np.zeros((n, n, n)), *[q.copy() for q in p], w, bs, bm) * f).sum()
    assert np.isclose(lhs, (w * O.read_cic(f, *p, bs, bm)).sum())


def test_gather_cpu_and_gpu_formulas_agree_in_float64():
    rng = np.random.default_rng(2)
    p = [rng.uniform(0, 100, 1000) for _ in range(3)]
    a = O.gather_cells(*p, (32, 32, 32), np.full(3, 100.0), np.zeros(3), True, "cpu")
    b = O.gather_cells(*p, (32, 32, 32), np.full(3, 100.0), np.zeros(3), True, "gpu")
    assert all(np.array_equal(x, y) for x, y in zip(a[0], b[0]))
    assert all(np.allclose(x, y, atol=1e-12) for x, y in zip(a[3], b[3]))


def test_tsc_partition_of_unity_and_quadratic_exactness():
    n, L = 16, 16.0
    bs, bm = box(n, L)
    rng = np.random.default_rng(3)
    p = [rng.uniform(0, L, 2000) for _ in range(3)]
    w = rng.uniform(0.5, 1.5, 2000)
    rho = O.tsc_scatter(np.zeros((n, n, n)), *p, w, bs, bm)
    assert np.isclose(rho.sum(), w.sum())
    i = np.arange(n, dtype=np.float64)
    fld = 1.5 * i[None, None, :] + 0 * i[None, :, None] + 0 * i[:, None, None]
    q = [rng.uniform(2, L - 3, 300) for _ in range(3)]
    assert np.allclose(O.read_tsc(fld, *q, bs, bm), 1.5 * q[0])          # TSC reproduces linear fields


# ---- recon.jl set-up -----------------------------------------------------------------------------
def test_setup_overdensity_uniform_lattice_is_zero():
    n, L = 16, 64.0
    g = (np.arange(n) + 0.5) * (L / n)
    zz, yy, xx = np.meshgrid(g, g, g, indexing="ij")
    rec = O.IterativeRecon(bias=2.0, f=0.8, smoothing_radius=6.0, box_size=np.full(3, L), box_min=np.zeros(3))
    d = O.setup_overdensity(np.zeros((n, n, n)), rec, xx.ravel().copy(), yy.ravel().copy(), zz.ravel().copy(),
                            np.ones(n ** 3))
    assert np.abs(d).max() < 1e-12


def test_setup_overdensity_randoms_equal_to_data_is_zero_and_threshold():
    n, L = 16, 64.0
    rng = np.random.default_rng(4)
    p = [rng.uniform(8, 40, 20000) for _ in range(3)]       # occupies part of the box only
    w = np.ones(20000)
    rec = O.IterativeRecon(bias=2.0, f=0.8, smoothing_radius=4.0, box_size=np.full(3, L), box_min=np.zeros(3))
    info = {}
    d = O.setup_overdensity(np.zeros((n, n, n)), rec, *p, w, *[q.copy() for q in p], w, info=info)
    assert np.abs(d).max() < 1e-10                           # data == randoms -> alpha = 1, delta = 0
    assert np.isclose(info["alpha"], 1.0)
    assert np.isclose(info["threshold"], 0.01 * info["ran"].sum() / 20000)
    assert (info["ran"] <= info["threshold"]).sum() > 0      # empty region exists and is masked to 0


# ---- iterative.jl -----------------------------------------------------------------------------------
def test_displacement_of_plane_wave():
    """delta = cos(k.x)  ->  Psi = i k delta_k / k^2  ->  Psi(x) = -(k/k^2) sin(k.x);  -div Psi = delta."""
    n, L = 32, 200.0
    c, s, k = plane_wave(n, L, (1, 2, 2))
    rec = O.IterativeRecon(bias=1.0, f=0.5, smoothing_radius=5.0, box_size=np.full(3, L), box_min=np.zeros(3))
    psi = O.displacement_meshes(c, rec)
    for a in range(3):
        assert np.allclose(psi[a], -(k[a] / (k @ k)) * s, atol=1e-12)


def test_iterate_identities():
    n, L = 16, 100.0
    rng = np.random.default_rng(5)
    ds = rng.standard_normal((n, n, n))
    kv = O.k_vec((n, n, n), np.full(3, L), np.float64)
    xv = O.x_vec((n, n, n), np.full(3, L), np.full(3, 500.0), np.float64)
    for los in ((0.0, 0.0, 1.0), None):
        out = O.iterate(rng.standard_normal((n, n, n)), ds, kv, 2, 0.0, los, xv)
        assert np.allclose(out, ds)                                     # beta = 0: delta_r = delta_s
    # fixed LOS along z, plane wave along z: mu = 1  ->  delta^(1) = delta_s (1 - beta/(1+beta))
    c, _, _ = plane_wave(n, L, (0, 0, 3))
    beta = 0.35
    out = O.iterate(c.copy(), c, kv, 1, beta, (0.0, 0.0, 1.0), None)
    assert np.allclose(out, c / (1 + beta), atol=1e-12)
    out2 = O.iterate(out.copy(), c, kv, 2, beta, (0.0, 0.0, 1.0), None)
    assert np.allclose(out2, c * (1 - beta / (1 + beta)), atol=1e-12)      # delta_s - beta * delta^(1)
    # plane wave perpendicular to the LOS: untouched
    cx, _, _ = plane_wave(n, L, (2, 0, 0))
    assert np.allclose(O.iterate(cx.copy(), cx, kv, 1, beta, (0.0, 0.0, 1.0), None), cx, atol=1e-12)


def test_radial_iteration_far_observer_tends_to_fixed_los():
    """With the box far away along +z the radial operator tends to the plane-parallel one."""
    n, L = 16, 50.0
    c, _, _ = plane_wave(n, L, (1, 0, 2))
    kv = O.k_vec((n, n, n), np.full(3, L), np.float64)
    far = np.array([-L / 2, -L / 2, 1e7])
    xv = O.x_vec((n, n, n), np.full(3, L), far, np.float64)
    a = O.iterate(c.copy(), c, kv, 1, 0.4, None, xv)
    b = O.iterate(c.copy(), c, kv, 1, 0.4, (0.0, 0.0, 1.0), None)
    assert np.abs(a - b).max() < 1e-4


def test_read_shifts_rsd_projection():
    n, L = 16, 100.0
    c, _, _ = plane_wave(n, L, (1, 1, 0))
    rec = O.IterativeRecon(bias=1.0, f=0.6, smoothing_radius=5.0, box_size=np.full(3, L), box_min=np.zeros(3),
                           los=(0.0, 1.0, 0.0))
    rng = np.random.default_rng(6)
    p = [rng.uniform(0, L, 200) for _ in range(3)]
    disp = O.read_shifts(rec, *p, c, "disp")
    rsd = O.read_shifts(rec, *p, c, "rsd")
    tot = O.read_shifts(rec, *p, c, "sum")
    assert np.allclose(rsd[0], 0) and np.allclose(rsd[2], 0) and np.allclose(rsd[1], 0.6 * disp[1])
    assert all(np.allclose(t, d + r) for t, d, r in zip(tot, disp, rsd))
    new = O.reconstructed_positions(rec, *p, c, "sum")
    assert all(np.allclose(nw, q - t) for nw, q, t in zip(new, p, tot))
    rec.los = None                                                          # radial: r_hat = p/|p|
    rsd = O.read_shifts(rec, *p, c, "rsd")
    r = np.sqrt(sum(q * q for q in p))
    dot = sum(d * q / r for d, q in zip(disp, p))
    assert all(np.allclose(rs, 0.6 * dot * q / r) for rs, q in zip(rsd, p))


# ---- multigrid.jl ---------------------------------------------------------------------------------------
def fd_eigen(n, L, m, beta=0.0, los=None):
    h = L / n
    lam = [(2 - 2 * np.cos(2 * np.pi * mm / n)) / h ** 2 for mm in m]
    out = sum(lam)
    if los is not None:
        l = np.asarray(los, float)
        l2 = l @ l
        out += beta * sum(lam[a] * l[a] ** 2 for a in range(3)) / l2
    return out


@pytest.mark.parametrize("los,beta", [((0.0, 0.0, 1.0), 0.0), ((0.0, 0.0, 1.0), 0.35), ((1.0, 0.0, 0.0), 0.5)])
def test_stencil_eigenfunction_jacobi_fixed_point_and_residual(los, beta):
    """Axis-aligned plane waves are eigenfunctions of the 19-point operator (cross terms vanish for an
    axis-aligned LOS): L cos = lambda cos, so v = f/lambda is a fixed point of jacobi! with zero residual."""
    n, L, m = 16, 32.0, (1, 2, 3)
    c, _, _ = plane_wave(n, L, m)
    lam = fd_eigen(n, L, m, beta, los)
    bs, bm = box(n, L)
    xv = O.x_vec((n, n, n), bs, bm, np.float64)
    v = c / lam
    assert np.abs(O.residual(v, c, xv, bs, bm, beta, los)).max() < 1e-12
    assert np.allclose(O.jacobi(v.copy(), c, xv, bs, bm, beta, 0.4, 3, los), v, atol=1e-13)


def test_restrict_prolong_constants_and_linear():
    rng = np.random.default_rng(7)
    assert np.allclose(O.restrict(np.full((8, 8, 8), 3.25)), 3.25)
    assert np.allclose(O.prolong(np.empty((8, 8, 8)), np.full((4, 4, 4), -1.5)), -1.5)
    # full weighting preserves the mean; prolongation is exact for fields linear away from the seam
    f = rng.standard_normal((8, 8, 8))
    assert np.isclose(O.restrict(f).mean(), f.mean())
    i = np.arange(8, dtype=float)
    coarse = 2.0 * i[None, None, :] + 0 * i[None, :, None] + 0 * i[:, None, None]
    fine = O.prolong(np.empty((16, 16, 16)), coarse)
    j = np.arange(16, dtype=float)
    expect = 2.0 * (j - 1) / 2                          # fine 2c+1 sits on coarse c
    assert np.allclose(fine[3, 5, 1:15], expect[1:15])


def test_fmg_converges_to_discrete_solution():
    n, L = 32, 64.0
    bs, bm = box(n, L)
    rng = np.random.default_rng(8)
    f = sum(rng.standard_normal() * plane_wave(n, L, m, phase=rng.uniform(0, 6))[0]
            for m in [(1, 0, 0), (0, 2, 1), (3, 1, 2), (2, 2, 0)])
    beta, los = 0.3, (0.0, 0.0, 1.0)
    v = O.fmg(f, None, bs, bm, beta, 0.4, 5, 6, los)
    xv = O.x_vec((n, n, n), bs, bm, np.float64)
    r = O.residual(v, f, xv, bs, bm, beta, los)
    assert np.sqrt((r ** 2).mean()) < 2e-4 * np.sqrt((f ** 2).mean())       # ~0.2 contraction per V-cycle
    # one more V-cycle contracts the residual further
    v2 = O.vcycle(v.copy(), f, bs, bm, beta, 0.4, 5, los)
    r2 = O.residual(v2, f, xv, bs, bm, beta, los)
    assert (r2 ** 2).sum() < 0.2 * (r ** 2).sum()


def test_multigrid_matches_iterative_for_smooth_field():
    """Both solvers reconstruct the same displacement for a well-resolved field (O(h^2) apart)."""
    n, L, N = 32, 400.0, 60000
    rng = np.random.default_rng(9)
    base = [rng.uniform(0, L, N) for _ in range(3)]
    pos = [base[0] + 6.0 * np.sin(2 * np.pi * base[0] / L), base[1], base[2] + 5.0 * np.sin(2 * np.pi * 2 * base[2] / L)]
    pos = [q % L for q in pos]
    w = np.ones(N)
    kw = dict(bias=1.0, f=0.0, smoothing_radius=30.0, box_size=np.full(3, L), box_min=np.zeros(3), los=(0.0, 0.0, 1.0))
    it = O.IterativeRecon(**kw)
    mg = O.MultigridRecon(**kw)
    m1 = O.run(it, (n, n, n), *[q.copy() for q in pos], w)
    m2 = O.run(mg, (n, n, n), *[q.copy() for q in pos], w)
    s1 = O.read_shifts(it, *pos, m1, "disp")
    s2 = O.read_shifts(mg, *pos, m2, "disp")
    scale = max(np.abs(s).max() for s in s1)
    assert scale > 1.0
    assert max(np.abs(a - b).max() for a, b in zip(s1, s2)) < 0.05 * scale


def test_float32_oracle_tracks_float64():
    n, L, N = 32, 500.0, 40000
    rng = np.random.default_rng(10)
    pos = [rng.uniform(0, L, N) for _ in range(3)]
    out = {}
    for T in (np.float32, np.float64):
        rec = O.IterativeRecon(bias=2.2, f=0.757, smoothing_radius=15.0, box_size=np.full(3, L, T),
                               box_min=np.zeros(3, T), los=(0.0, 0.0, 1.0))
        p = [q.astype(T) for q in pos]
        mesh = O.run(rec, (n, n, n), *[q.copy() for q in p], np.ones(N, T))
        out[T] = O.read_shifts(rec, *p, mesh, "sum")
    for a in range(3):
        assert np.abs(out[np.float32][a] - out[np.float64][a]).max() < 1e-3


def test_fmg_does_not_feel_the_mean_of_the_right_hand_side_except_through_float32_rounding():
    """On a periodic cubic mesh the diagonal of the multigrid operator is the same in every cell (2 (3 + beta) / cell^2,
    src/multigrid.jl:82), so a non-zero mean of the right-hand side only makes the damped-Jacobi iterate drift by a
    constant (omega mean(f) / diag per sweep): the fluctuating part of the potential -- all that the shifts see -- is the
    same with and without it.  True in Float64; in Float32 the drifting constant (hundreds of times the rms of the
    fluctuations) eats the low bits and the reference arithmetic moves away from its own Float64 result, while the same
    arithmetic on the mean-free right-hand side does not.  tests/test_gpu_a_configs.py relies on this for config 3
    (a survey's delta has a non-zero mean): the device is held to the Float64-equivalent answer."""
    import baorec_oracle as O
    n = 32
    f32 = np.float32
    bs, bm = np.full(3, 2798.33, f32), np.array([1400.0, -1400.0, -1400.0], f32)       # observer outside the box: radial LOS
    rng = np.random.default_rng(4)
    f = rng.standard_normal((n, n, n)).astype(f32)
    f = (f - f32(f.mean(dtype=np.float64)) + f32(0.05)).astype(f32)
    f0 = (f - f32(f.mean(dtype=np.float64))).astype(f32)
    dm = lambda a: a - a.mean()
    rel = lambda a, b: float(np.sqrt(np.mean((dm(a) - dm(b)) ** 2)) / dm(b).std())
    a64 = O.fmg(f.astype(np.float64), np.zeros((n, n, n)), bs.astype(np.float64), bm.astype(np.float64), 0.344, 0.4, 5, 6, None)
    b64 = O.fmg(f0.astype(np.float64), np.zeros((n, n, n)), bs.astype(np.float64), bm.astype(np.float64), 0.344, 0.4, 5, 6, None)
    a32 = O.fmg(f.copy(), np.zeros((n, n, n), f32), bs, bm, f32(0.344), f32(0.4), 5, 6, None)
    b32 = O.fmg(f0.copy(), np.zeros((n, n, n), f32), bs, bm, f32(0.344), f32(0.4), 5, 6, None)
    assert abs(a64.mean()) > 100 * dm(a64).std()          # the drift
    assert rel(a64, b64) < 1e-5                           # ... is a pure constant
    assert rel(b32, a64) < 2e-5                           # Float32 on the mean-free right-hand side: fine
    assert rel(a32, a64) > 5 * rel(b32, a64)              # Float32 with the drift: visibly worse


def test_read_grad_cic_known_answers():
    """The finite-difference read-back (oracle restatement of the reference's commented-out read_grad_cic!,
    src/mas.jl:388-466, with the intended geometry).  A particle ON a mesh point gets exactly the central difference
    there; for a plane wave the result is the spectral gradient times sin(kh)/(kh) (the central difference) up to the
    CIC interpolation; the Float64 and Float32 evaluations agree to rounding."""
    import baorec_oracle as O
    n, L = 32, 640.0
    h = L / n
    bs, bm = np.full(3, L), np.array([100.0, -50.0, 0.0])
    rng = np.random.default_rng(1)
    phi = rng.standard_normal((n, n, n))
    i, j, k = 5, 31, 0                                       # wraps in y (upper neighbour) and z (lower neighbour)
    g = O.read_grad_cic(phi, np.array([bm[0] + i * h]), np.array([bm[1] + j * h]), np.array([bm[2] + k * h]), bs, bm)
    assert np.isclose(g[0][0], (phi[k, j, i + 1] - phi[k, j, i - 1]) / (2 * h), rtol=1e-12)
    assert np.isclose(g[1][0], (phi[k, 0, i] - phi[k, j - 1, i]) / (2 * h), rtol=1e-12)
    assert np.isclose(g[2][0], (phi[1, j, i] - phi[n - 1, j, i]) / (2 * h), rtol=1e-12)
    m = 2
    kx = 2 * np.pi * m / L
    xs = bm[0] + h * np.arange(n)
    wave = np.broadcast_to(np.sin(kx * (xs - bm[0]))[None, None, :], (n, n, n)).copy()
    N = 2000
    x, y, z = (bm[a] + L * rng.random(N) for a in range(3))
    g = O.read_grad_cic(wave, x, y, z, bs, bm)
    want = kx * np.cos(kx * (x - bm[0])) * np.sin(kx * h) / (kx * h)
    assert np.abs(g[0] - want).max() < 0.02 * kx and np.abs(g[1]).max() < 1e-12 and np.abs(g[2]).max() < 1e-12
    g32 = O.read_grad_cic(phi.astype(np.float32), x.astype(np.float32), y.astype(np.float32), z.astype(np.float32),
                          bs.astype(np.float32), bm.astype(np.float32))
    g64 = O.read_grad_cic(phi.astype(np.float32).astype(np.float64), x.astype(np.float32).astype(np.float64),
                          y.astype(np.float32).astype(np.float64), z.astype(np.float32).astype(np.float64), bs, bm)
    # (the Float32 cell index can differ from Float64's for a particle within rounding of a mesh point: compare the bulk)
    err = np.abs(g32[0] - g64[0])
    assert np.median(err) < 1e-6 and np.quantile(err, 0.99) < 1e-4
    # and the multigrid read-back through it
    rec = O.MultigridRecon(bias=2.0, f=0.8, smoothing_radius=10.0, box_size=bs.astype(np.float32), box_min=bm.astype(np.float32), los=(0.0, 0.0, 1.0))
    rec.fd_gradient = True
    s = O.read_shifts(rec, x.astype(np.float32), y.astype(np.float32), z.astype(np.float32), wave.astype(np.float32), "disp")
    assert np.allclose(s[0], want, atol=0.02 * kx)
