"""CPU: known-answer tests of the power-spectrum multipole oracle (oracle/pk_oracle.py), and -- through it -- a
physics known-answer test of the reconstruction oracle itself.

The reference ships no assertion about its results: it validates by eye, plotting P_0/P_2 before and after
(test_helpers/simulation.py:56-75).  The last test below turns that into numbers: a lognormal box with a linear
redshift-space shift shows the Kaiser quadrupole; IterativeRecon + reconstructed positions with field = :rsd
(src/recon.jl:366-388) must remove it and bring the monopole back to its real-space value.  A sign error anywhere
in the chain (overdensity normalisation, the RSD recurrence, -ik/k^2, the shift epilogue) makes the quadrupole
grow instead."""
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest

import baorec_oracle as O
import baorec_oracle_fast as fast
import pk_oracle as PK
from util import lognormal_radial

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "benchmarks"))
f32 = np.float32


def test_plane_wave_lands_in_its_bin_with_the_right_amplitude():
    n, L, A, m = 32, np.array([100.0, 100.0, 100.0]), 0.1, 3
    x = np.arange(n) * L[0] / n
    wave = 1 + A * np.cos(2 * np.pi * m * x / L[0])
    kf = 2 * np.pi / 100.0
    for axis, los, l2, l4 in ((2, (0, 0, 1), -0.5, 0.375), (2, (1, 0, 0), 1.0, 1.0), (0, (0, 0, 1), 1.0, 1.0)):
        shape = [1, 1, 1]
        shape[axis] = n
        rho = np.ones((n, n, n)) * wave.reshape(shape)
        r = PK.power_multipoles(rho, L, los=los, kmin=0.5 * kf, dk=kf, nbins=8, mas_power=0)
        b = m - 1                                            # bins are centred on the multiples of kf
        # two modes (+-k0) of power V A^2 / 4 each among the modes of the bin
        assert np.isclose(r["p0"][b], L.prod() * A * A / 4 * 2 / r["nmodes"][b], rtol=1e-10)
        assert np.isclose(r["p2"][b] / r["p0"][b], 5 * l2, rtol=1e-6) and np.isclose(r["p4"][b] / r["p0"][b], 9 * l4, rtol=1e-6)
        others = np.delete(np.arange(8), b)
        assert np.abs(r["p0"][others]).max() < 1e-20 * r["p0"][b] + 1e-12


def test_every_mode_is_counted_once():
    for shape, L in (((16, 16, 16), (50.0, 50.0, 50.0)), ((12, 10, 14), (30.0, 25.0, 35.0)), ((8, 8, 9), (20.0, 20.0, 20.0))):
        rho = np.random.default_rng(0).random(shape)
        r = PK.power_multipoles(rho, L, kmin=0.0, dk=0.05, nbins=400, mas_power=2)
        assert r["nmodes"].sum() == np.prod(shape) - 1       # all of the Hermitian mesh but k = 0
        k, mu, W, wt = PK.mode_table(shape, L, (0, 0, 1))
        assert np.abs(mu).max() <= 1 + 1e-12 and W.min() > 0 and W.max() == 1.0


def test_poisson_catalog_has_only_shot_noise():
    n, L, N = 64, 500.0, 200_000
    rng = np.random.default_rng(3)
    pos = [(L * rng.random(N)).astype(f32) for _ in range(3)]
    bs, bm = np.full(3, L, f32), np.zeros(3, f32)
    rho = O.cic_scatter(np.zeros((n, n, n), f32), *pos, np.ones(N, f32), bs, bm, True)
    shot = L ** 3 / N
    r = PK.power_multipoles(rho, bs, kmin=0.0, dk=0.02, nbins=8, mas_power=2, shot=shot)   # up to 0.4 of Nyquist
    err = shot * np.sqrt(2.0 / r["nmodes"])
    assert np.all(np.abs(r["p0"][1:]) < 5 * err[1:] + 0.03 * shot)      # compensated CIC: aliasing stays below 3 % there
    assert np.all(np.abs(r["p2"][1:]) < 5 * np.sqrt(5.0) * err[1:] + 0.03 * shot)


@pytest.fixture(scope="module")
def F():
    if not fast.available():
        subprocess.run(["make", "-C", str(ROOT / "oracle"), "-s"], check=True)
    return fast.load()


@pytest.mark.parametrize("algorithm", ["iterative", "multigrid"])
def test_reconstruction_removes_the_kaiser_quadrupole(F, algorithm):
    import catalogs as Cat
    L, n, N, f, R = 1000.0, 64, 2_000_000, 0.757, 10.0
    red, w = Cat.lognormal_box(N, L, seed=5, n_gen=64, sigma=0.8, f_rsd=f)
    real, _ = Cat.lognormal_box(N, L, seed=5, n_gen=64, sigma=0.8, f_rsd=0.0)     # same particles, no RSD shift
    red, real, w = [p.numpy() for p in red], [p.numpy() for p in real], w.numpy()
    bs, bm = np.full(3, L, f32), np.zeros(3, f32)

    def multipoles(p):
        rho = F.cic_scatter(np.zeros((n, n, n), f32), *[q.copy() for q in p], w, bs, bm, True)
        return PK.power_multipoles(rho, bs, los=(0, 0, 1), kmin=0.0, dk=0.02, nbins=4, mas_power=2, shot=L ** 3 / N)

    r_real, r_red = multipoles(real), multipoles(red)
    b = 1                                                    # 0.02 <= k < 0.04 h/Mpc: ~1000 modes, exp(-k^2 R^2 / 2) = 0.95
    beta = f                                                 # the tracer is the field itself: bias 1
    kaiser_q = (4 * beta / 3 + 4 * beta ** 2 / 7) / (1 + 2 * beta / 3 + beta ** 2 / 5)      # 0.826
    q_real, q_red = r_real["p2"][b] / r_real["p0"][b], r_red["p2"][b] / r_red["p0"][b]
    assert abs(q_real) < 0.15                                # isotropic up to the sample variance of ~1000 modes
    assert abs(q_red - q_real - kaiser_q) < 0.2              # measured: +0.68
    assert 1.3 < r_red["p0"][b] / r_real["p0"][b] < 1.8      # Kaiser monopole boost 1.62 (linear theory)
    kw = dict(bias=1.0, f=f, smoothing_radius=R, box_size=bs, box_min=bm, los=(0.0, 0.0, 1.0))
    rec = F.IterativeRecon(n_iter=3, **kw) if algorithm == "iterative" else F.MultigridRecon(**kw)   # both solvers: -0.018 / -0.020, 0.967 / 0.965
    mesh = F.run(rec, (n, n, n), *[q.copy() for q in red], w)
    new = F.reconstructed_positions(rec, *red, mesh, field="rsd")
    top = np.nextafter(f32(L), f32(0))
    new = [np.clip(np.mod(q, f32(L)), 0, top).astype(f32) for q in new]
    r_new = multipoles(new)
    assert abs(r_new["p2"][b] / r_new["p0"][b] - q_real) < 0.12   # the quadrupole is back at its real-space value (measured: -0.08) ...
    assert abs(r_new["p0"][b] / r_real["p0"][b] - 1) < 0.15  # ... and so is the monopole (measured: 0.967)
    # the displacement itself points the right way: removing it as well (field = :sum) lowers the large-scale power
    both = F.reconstructed_positions(rec, *red, mesh, field="sum")
    both = [np.clip(np.mod(q, f32(L)), 0, top).astype(f32) for q in both]
    assert multipoles(both)["p0"][b] < 0.2 * r_real["p0"][b]     # measured: ~0 (k R << 1: the smoothed field is the field)


@pytest.mark.parametrize("algorithm", ["iterative", "multigrid"])
def test_rec_sym_keeps_and_rec_iso_removes_the_kaiser_boost(F, algorithm):
    """The reference's own post-reconstruction check (test_helpers/simulation.py:48-70): data displaced with
    field = :sum, randoms displaced with :sum ("sym") or :disp ("iso"), power spectrum of data minus shifted randoms.
    Linear theory: RecSym keeps the redshift-space clustering on large scales (Kaiser boost and quadrupole unchanged),
    RecIso returns the real-space one (no quadrupole).  Exercises read_shifts on a second catalog and the sign of Psi:
    with the wrong sign the shifted data and the shifted randoms do not cancel and the large-scale power is lost."""
    import catalogs as Cat
    L, n, N, f, R = 1000.0, 64, 2_000_000, 0.757, 10.0
    red, w = Cat.lognormal_box(N, L, seed=5, n_gen=64, sigma=0.8, f_rsd=f)
    real, _ = Cat.lognormal_box(N, L, seed=5, n_gen=64, sigma=0.8, f_rsd=0.0)
    red, real, w = [p.numpy() for p in red], [p.numpy() for p in real], w.numpy()
    bs, bm = np.full(3, L, f32), np.zeros(3, f32)
    rng = np.random.default_rng(9)
    NR = 4 * N
    ran, wr = [(L * rng.random(NR)).astype(f32) for _ in range(3)], np.ones(NR, f32)
    top = np.nextafter(f32(L), f32(0))

    def wrap(ps):
        return [np.clip(np.mod(q, f32(L)), 0, top).astype(f32) for q in ps]

    def mesh(p, ww):
        return F.cic_scatter(np.zeros((n, n, n), f32), *[q.copy() for q in p], ww, bs, bm, True)

    def multipoles(p, r=None):
        shot = L ** 3 / N * (1 + (N / NR if r is not None else 0))
        return PK.power_multipoles(mesh(p, w), bs, los=(0, 0, 1), kmin=0.0, dk=0.02, nbins=4, mas_power=2, shot=shot,
                                   randoms=None if r is None else mesh(r, wr))

    b = 1
    r_real, r_red = multipoles(real), multipoles(red)
    assert abs(multipoles(red, ran)["p0"][b] / r_red["p0"][b] - 1) < 0.03      # unshifted randoms carry no signal
    kw = dict(bias=1.0, f=f, smoothing_radius=R, box_size=bs, box_min=bm, los=(0.0, 0.0, 1.0))
    rec = F.IterativeRecon(n_iter=3, **kw) if algorithm == "iterative" else F.MultigridRecon(**kw)
    m = F.run(rec, (n, n, n), *[q.copy() for q in red], w)
    d_sum = wrap(F.reconstructed_positions(rec, *red, m, field="sum"))
    sym = multipoles(d_sum, wrap(F.reconstructed_positions(rec, *ran, m, field="sum")))
    iso = multipoles(d_sum, wrap(F.reconstructed_positions(rec, *ran, m, field="disp")))
    q = lambda r: r["p2"][b] / r["p0"][b]
    assert abs(sym["p0"][b] / r_red["p0"][b] - 1) < 0.05 and abs(q(sym) - q(r_red)) < 0.05     # measured: 0.995, +0.010
    assert abs(iso["p0"][b] / r_real["p0"][b] - 1) < 0.1 and abs(q(iso) - q(r_real)) < 0.12    # measured: 0.958, -0.083


@pytest.mark.parametrize("algorithm", ["iterative", "multigrid"])
def test_lightcone_mode_removes_a_radial_kaiser_quadrupole(F, algorithm):
    """The other half of the hot path -- radial line of sight (los = nothing) with a random catalog: setup_box with
    its padding, (data - alpha randoms) / randoms with the `ran > threshold` mask (src/recon.jl:60-91), the radial
    iteration x_i x_j / |x|^2 (src/iterative.jl:14-40) or the radial multigrid stencil, and the per-particle line of
    sight of the shift epilogue (src/recon.jl:277-304).  A periodic lognormal box is seen from 3000 Mpc/h below its
    centre (line of sight within 10 degrees of z), shifted radially, reconstructed as a survey would be -- observer at
    the origin, uniform randoms in the box, padded mesh -- and brought back into the periodic box to measure the
    multipoles.  Redshift space: P2/P0 = 0.77, boost 1.56; after field = :rsd: 0.017 / 0.020 (real space: 0.010) and
    P0 = 0.971 / 0.972 of the real-space value."""
    L, ng, f, R = 1000.0, 64, 0.757, 10.0
    obs = np.array([500.0, 500.0, -3000.0])
    real, red = lognormal_radial(2_000_000, L, ng, 0.8, f, obs, 5)
    N = len(real)
    bs, bm, w = np.full(3, L, f32), np.zeros(3, f32), np.ones(N, f32)
    top = np.nextafter(f32(L), f32(0))

    def multipoles(a):
        p = [np.clip(np.mod(a[:, i].astype(f32), f32(L)), 0, top).astype(f32) for i in range(3)]
        rho = F.cic_scatter(np.zeros((ng, ng, ng), f32), *p, w, bs, bm, True)
        return PK.power_multipoles(rho, bs, los=(0, 0, 1), kmin=0.0, dk=0.02, nbins=4, mas_power=2, shot=L ** 3 / N)

    b = 1
    q = lambda r: r["p2"][b] / r["p0"][b]
    r_real, r_red = multipoles(real), multipoles(red)
    assert abs(q(r_real)) < 0.1 and abs(q(r_red) - q(r_real) - 0.826) < 0.2 and 1.3 < r_red["p0"][b] / r_real["p0"][b] < 1.8
    NR = 4 * N
    ran = np.random.default_rng(11).random((NR, 3)) * L
    cat = lambda a: [(a[:, i] - obs[i]).astype(f32) for i in range(3)]                # survey coordinates: observer at the origin
    d, r_ = cat(red), cat(ran)
    kw = dict(bias=1.0, f=f, smoothing_radius=R, los=None)
    rec = F.IterativeRecon(n_iter=3, **kw) if algorithm == "iterative" else F.MultigridRecon(**kw)
    mesh = F.run(rec, (128, 128, 128), *[p.copy() for p in d], w, *[p.copy() for p in r_], np.ones(NR, f32))
    assert np.allclose(rec.box_size, 1500.0, rtol=1e-3)                              # setup_box: extent of the randoms + 500
    new = F.reconstructed_positions(rec, *d, mesh, field="rsd")
    r_new = multipoles(np.stack([new[i].astype(np.float64) + obs[i] for i in range(3)], 1))
    assert abs(q(r_new) - q(r_real)) < 0.1
    assert abs(r_new["p0"][b] / r_real["p0"][b] - 1) < 0.1


# ---- interlacing and the PCS window: the analytic image sum of ONE particle ---------------------------------
def _image_sum(k3, x0, h, p, mmax, even_only):
    """sum over images m of W(k + 2 k_N m) exp(-i (k + 2 k_N m) . x0), W = prod sinc(q h / 2)^p: the transform of a unit
    particle at x0 painted with a B-spline of order p on a mesh of spacing h (per axis), exactly; with interlacing only
    the images with m_x + m_y + m_z even survive."""
    ms = np.arange(-mmax, mmax + 1)
    out = np.zeros(len(k3), np.complex128)
    per_axis = []
    for a in range(3):
        q = k3[:, a][:, None] + 2 * np.pi / h[a] * ms[None, :]                      # [modes][images]
        xa = q * h[a] / 2
        w = np.where(xa == 0, 1.0, np.sin(xa) / np.where(xa == 0, 1.0, xa)) ** p
        per_axis.append(w * np.exp(-1j * q * x0[a]))
    par = (np.abs(ms) % 2)                                                           # parity of m per axis
    for ex in (0, 1):
        for ey in (0, 1):
            for ez in (0, 1):
                if even_only and (ex + ey + ez) % 2:
                    continue
                out += per_axis[0][:, par == ex].sum(1) * per_axis[1][:, par == ey].sum(1) * per_axis[2][:, par == ez].sum(1)
    return out


@pytest.mark.parametrize("mas,p", [("tsc", 3), ("pcs", 4)])
def test_interlacing_cancels_the_odd_images(mas, p):
    """One particle, mesh 8 x 10 x 12: the oracle's painted transform IS the image sum, and the interlaced combination
    is the sum over the even images only -- to the truncation of the sum (|m| <= 40: the tail falls as m^-p)."""
    grid, L = (8, 10, 12), np.array([80.0, 90.0, 120.0])
    h = L / np.array(grid, np.float64)
    x0 = np.array([13.7, 41.2, 66.6])
    f64 = np.float64
    pos = [np.array([v], f64) for v in x0]
    w = np.ones(1, f64)
    scatter = O.tsc_scatter if mas == "tsc" else O.pcs_scatter
    nx, ny, nz = grid
    m1 = scatter(np.zeros((nz, ny, nx), f64), *pos, w, L, np.zeros(3), True)
    m2 = scatter(np.zeros((nz, ny, nx), f64), *PK.interlace_positions(*pos, grid, L, np.zeros(3)), w, L, np.zeros(3), True)
    assert abs(m1.sum() - 1) < 1e-12 and abs(m2.sum() - 1) < 1e-12
    kx, ky, kz = O.k_vec(grid, L.astype(np.float32), np.float32)
    KZ, KY, KX = np.meshgrid(kz.astype(f64), ky.astype(f64), kx.astype(f64), indexing="ij")
    k3 = np.stack([KX.ravel(), KY.ravel(), KZ.ravel()], 1)
    plain = PK.field_k(m1, L).ravel()
    inter = PK.field_k(m1, L, rho_shifted=m2).ravel()
    ref_all = _image_sum(k3, x0, h, p, 40, False)
    ref_even = _image_sum(k3, x0, h, p, 40, True)
    # the Nyquist planes of a real-to-complex transform fold +k_N and -k_N together; compare away from them
    inner = (np.abs(k3[:, 0]) < np.pi / h[0] * 0.99) & (np.abs(k3[:, 1]) < np.pi / h[1] * 0.99) & (np.abs(k3[:, 2]) < np.pi / h[2] * 0.99)
    tol = 3e-4 if p == 3 else 2e-6
    assert np.abs(plain - ref_all)[inner].max() < tol
    assert np.abs(inter - ref_even)[inner].max() < tol
    # and it matters: near the Nyquist frequency the odd images are the bulk of the aliasing
    hi = inner & (np.sqrt((k3 ** 2).sum(1)) > 0.6 * np.pi / h.max())
    exact = np.exp(-1j * (k3 * x0).sum(1))
    W = np.prod([np.where(k3[:, a] == 0, 1.0, np.sin(k3[:, a] * h[a] / 2) / np.where(k3[:, a] == 0, 1.0, k3[:, a] * h[a] / 2)) ** p for a in range(3)], 0)
    err_plain = np.abs(plain / W - exact)[hi].mean()
    err_inter = np.abs(inter / W - exact)[hi].mean()
    assert err_inter < 0.5 * err_plain


def test_compute_auto_box_interlaced_pcs_recovers_shot_noise_to_nyquist():
    """A Poisson catalog through compute_auto_box (PCS, interlaced): P_0 = V / N up to the Nyquist frequency -- without
    interlacing the aliased shot noise lifts the last bins (CIC: by 30 %)."""
    rng = np.random.default_rng(5)
    N, L, n = 200_000, 1000.0, 32
    pos = [rng.uniform(0, L, N).astype(np.float32) for _ in range(3)]
    w = np.ones(N, np.float32)
    bs = np.full(3, L, np.float32)
    kny = np.pi * n / L
    kw = dict(kmin=0.0, dk=kny / 8, nbins=8)
    good = PK.compute_auto_box(*pos, w, bs, (n, n, n), mas="pcs", interlace=True, **kw)
    bad = PK.compute_auto_box(*pos, w, bs, (n, n, n), mas="cic", interlace=False, **kw)
    shot = L ** 3 / N
    assert np.abs(good["p0"][5:] / shot - 1).max() < 0.02                 # the three bins below Nyquist (thousands of modes each)
    assert bad["p0"][-1] / shot - 1 > 0.2 and bad["p0"][-2] / shot - 1 > 0.05   # measured: +30 % and +9 % of aliased shot noise
    plain = PK.compute_auto_box(*pos, w, bs, (n, n, n), mas="pcs", interlace=False, **kw)
    assert plain["p0"][-1] / shot - 1 > 0.08                              # PCS alone: +11 % in the last bin; interlaced: +1.5 %
