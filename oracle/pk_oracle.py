"""TEST INFRASTRUCTURE (CPU oracle) -- power-spectrum multipoles of a density mesh.

The reference validates a reconstruction by eye, through the power-spectrum multipoles of the catalogs
before and after (test_helpers/simulation.py:56-75, test_helpers/pyrecon_simulation.py:70-86: pypowspec
`compute_auto_box` on a 512^3 mesh, CIC, line of sight along z, multipoles 0/2/4, linear k bins).  pypowspec is a
third-party C library that is not under /root/reference; what it computes for a periodic box is the textbook
FFT estimator restated here (parity unpinned: no reference vectors exist for it; pinned by the analytic
known-answer tests in tests/test_pk_oracle.py -- plane waves, Poisson shot noise, the Kaiser quadrupole):

    delta(x)  = rho(x) / mean(rho) - 1                       on the mesh, rho from cic! (or TSC)
    delta_k   = sum_x delta(x) exp(-i k x)                   unnormalised forward transform (src/recon.jl:36)
    P(k)      = V |delta_k|^2 / M^2 / W(k)^2                 V = Lx Ly Lz, M = nx ny nz,
                W(k) = prod_a sinc(k_a h_a / 2)^p            mass-assignment window (p = 2 CIC, 3 TSC, 0 = none)
    P_l(k_i)  = (2l+1) <P(k) L_l(mu)>_{k in bin i} - [l == 0] shot     mu = k . los / |k|

Every mode of the full (Hermitian) mesh counts once: on the half mesh the planes kx = 0 and kx = Nyquist have
weight 1, the others weight 2.  k = 0 is excluded.  Bins are [kmin + i dk, kmin + (i+1) dk), i < nbins
(index = floor((|k| - kmin) * (1/dk)) in Float64).

Interlacing (the GRID_INTERLACE = T of the reference's helper configuration, test_helpers/powspec_auto.conf:125;
Sefusatti et al. 2016, the scheme of pypowspec / nbodykit / pypower): a second mesh is painted from the same catalog
with every position moved by +h/2 along each axis (periodic), and the two transforms are combined as
    f_k = [ f1_k + f2_k exp(+i (k_x h_x + k_y h_y + k_z h_z) / 2) ] / 2 ,
which cancels the aliasing images k + 2 k_N m with m_x + m_y + m_z odd (tests/test_pk_oracle.py checks exactly that
against the analytic image sum of one particle).
Only tests/ (and the product's parity tests) may import this module."""
from __future__ import annotations

import numpy as np
import scipy.fft

import baorec_oracle as O


def mode_table(shape_zyx, box_size, los, mas_power=2):
    """Per-mode |k| (from the reference's Float32 k_vec tables, src/utils.jl:3-10, combined in Float64), mu, window
    and Hermitian weight on the half mesh [nz][ny][nx/2+1]."""
    nz, ny, nx = shape_zyx
    L = np.asarray(box_size, np.float32)
    f32 = np.float32
    kx, ky, kz = O.k_vec((nx, ny, nz), L, f32)          # the tables the device context holds (src/utils.jl:3-10)
    KX, KY, KZ = kx[None, None, :], ky[None, :, None], kz[:, None, None]
    KX, KY, KZ = KX.astype(np.float64), KY.astype(np.float64), KZ.astype(np.float64)
    k = np.sqrt(KX * KX + KY * KY + KZ * KZ)             # Float64 arithmetic on the Float32 table values
    lv = np.asarray(los, np.float64)
    lv = lv / np.sqrt((lv * lv).sum())
    with np.errstate(invalid="ignore", divide="ignore"):
        mu = (KX * lv[0] + KY * lv[1] + KZ * lv[2]) / k
    mu[0, 0, 0] = 0.0
    h = L.astype(np.float64) / np.array([nx, ny, nz], np.float64)

    def sinc(k1, h1):   # sin(x)/x with x = k h / 2
        x = k1.astype(np.float64) * h1 / 2
        return np.where(x == 0, 1.0, np.sin(x) / np.where(x == 0, 1.0, x))

    W = (sinc(KX, h[0]) * sinc(KY, h[1]) * sinc(KZ, h[2])) ** mas_power if mas_power else np.ones_like(k)
    wt = np.full(k.shape, 2.0)
    wt[:, :, 0] = 1.0
    if nx % 2 == 0:
        wt[:, :, -1] = 1.0
    wt[0, 0, 0] = 0.0
    return k, mu, W, wt


def interlace_positions(x, y, z, grid_xyz, box_size, box_min):
    """The catalog of the second (interlaced) mesh: every position moved by half a cell, q = p + (L/n)/2 in the
    positions' own precision, and back by L where that leaves the box (q - min >= L)."""
    out = []
    for a, p in enumerate((x, y, z)):
        p = np.asarray(p)
        T = p.dtype.type
        L, mn = T(np.asarray(box_size, T)[a]), T(np.asarray(box_min, T)[a])
        half = T(T(L / T(grid_xyz[a])) / T(2))
        q = (p + half).astype(T)
        q = np.where((q - mn).astype(T) >= L, (q - L).astype(T), q)
        out.append(np.ascontiguousarray(q))
    return out


def field_k(rho, box_size, randoms=None, rho_shifted=None, randoms_shifted=None):
    """The normalised field of the estimator on the half mesh, complex128: rho_k / rho_0 (minus ran_k / ran_0), the
    interlaced combination when the shifted meshes are given."""
    def one(r, q):
        rk = scipy.fft.rfftn(np.asarray(r, np.float64), workers=-1)
        f = rk * (1.0 / float(rk[0, 0, 0].real))
        if q is not None:
            sk = scipy.fft.rfftn(np.asarray(q, np.float64), workers=-1)
            f = f - sk * (1.0 / float(sk[0, 0, 0].real))
        return f
    # Float64 transform: a Float32 transform of the raw density carries the rounding of its DC term (~1e-7 sum rho)
    # into every mode, which is 1e-4 of the signal in a sparse low-k bin -- the oracle is the truth, not a second
    # Float32 estimate
    f = one(rho, randoms)
    if rho_shifted is not None:
        nz, ny, nx = np.asarray(rho).shape
        L = np.asarray(box_size, np.float32)
        kx, ky, kz = O.k_vec((nx, ny, nz), L, np.float32)
        h = L.astype(np.float64) / np.array([nx, ny, nz], np.float64)
        ph = (np.exp(0.5j * kx.astype(np.float64) * h[0])[None, None, :] * np.exp(0.5j * ky.astype(np.float64) * h[1])[None, :, None]) \
            * np.exp(0.5j * kz.astype(np.float64) * h[2])[:, None, None]
        f = 0.5 * (f + one(rho_shifted, randoms_shifted) * ph)
    return f


def power_multipoles(rho, box_size, los=(0.0, 0.0, 1.0), kmin=0.0, dk=None, nbins=None, mas_power=2, shot=0.0, randoms=None,
                     rho_shifted=None, randoms_shifted=None):
    """Multipoles l = 0, 2, 4 of the density mesh `rho` ([nz][ny][nx], any normalisation; Float32 or Float64).
    `randoms`: optional density mesh of a (shifted) random catalog: the field is then rho / sum(rho) - ran / sum(ran),
    the "data minus shifted randoms" estimate of a reconstructed catalog (test_helpers/simulation.py:52-70,
    compute_auto_box_rand).  `rho_shifted` (and `randoms_shifted`): the meshes of the same catalogs painted half a cell
    further along every axis (interlace_positions) -> the interlaced estimate.
    Returns dict(k=<mean k per bin>, nmodes, p0, p2, p4); empty bins hold NaN."""
    rho = np.asarray(rho)
    nz, ny, nx = rho.shape
    L = np.asarray(box_size, np.float64)
    if dk is None:
        dk = 2 * np.pi / float(L.max())
    if nbins is None:
        nbins = int((np.pi * min(nx / L[0], ny / L[1], nz / L[2]) - kmin) / dk)
    f = field_k(rho, box_size, randoms, rho_shifted, randoms_shifted)
    re, im = f.real, f.imag
    k, mu, W, wt = mode_table(rho.shape, box_size, los, mas_power)
    V = float(L.prod())
    p = (re * re + im * im) * V / (W * W)
    b = np.floor((k - kmin) * (1.0 / dk)).astype(np.int64)     # 1/dk rounded once, like the device kernel
    ok = (b >= 0) & (b < nbins) & (wt > 0)
    b, w = b[ok], wt[ok]
    p, mu, kk = p[ok], mu[ok], k[ok]
    mu2 = mu * mu
    l2 = 1.5 * mu2 - 0.5
    l4 = (35.0 * mu2 * mu2 - 30.0 * mu2 + 3.0) / 8.0
    cnt = np.bincount(b, w, nbins)
    with np.errstate(invalid="ignore", divide="ignore"):
        out = dict(nmodes=cnt,
                   k=np.bincount(b, w * kk, nbins) / cnt,
                   p0=np.bincount(b, w * p, nbins) / cnt - shot,
                   p2=5.0 * np.bincount(b, w * p * l2, nbins) / cnt,
                   p4=9.0 * np.bincount(b, w * p * l4, nbins) / cnt)
    return out


MAS_POWER = {"cic": 2, "tsc": 3, "pcs": 4}


def compute_auto_box(x, y, z, w, box_size, grid, mas="tsc", interlace=True, los=(0.0, 0.0, 1.0), kmin=0.0, dk=None, nbins=None,
                     shot=0.0, rx=None, ry=None, rz=None, rw=None, box_min=(0.0, 0.0, 0.0)):
    """compute_auto_box / compute_auto_box_rand of the reference's helpers (test_helpers/simulation.py:36-70, the
    settings of test_helpers/powspec_auto.conf: PARTICLE_ASSIGN unset = TSC, GRID_INTERLACE = T) on a periodic box:
    paint (cic! / tsc / pcs of the reconstruction oracle), interlace, transform, multipoles.  grid = (nx, ny, nz)."""
    nx, ny, nz = grid
    scatter = {"cic": O.cic_scatter, "tsc": O.tsc_scatter, "pcs": O.pcs_scatter}[mas]
    bs, bm = np.asarray(box_size, np.float32), np.asarray(box_min, np.float32)

    def paint(px, py, pz, pw):
        return scatter(np.zeros((nz, ny, nx), np.float32), px.copy(), py.copy(), pz.copy(), pw, bs, bm, True)

    def both(px, py, pz, pw):
        if px is None:
            return None, None
        m1 = paint(px, py, pz, pw)
        m2 = paint(*interlace_positions(px, py, pz, grid, bs, bm), pw) if interlace else None
        return m1, m2
    d1, d2 = both(x, y, z, w)
    r1, r2 = both(rx, ry, rz, rw)
    return power_multipoles(d1, bs, los, kmin, dk, nbins, MAS_POWER[mas], shot, r1, d2, r2)
