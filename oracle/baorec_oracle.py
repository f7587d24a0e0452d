"""CPU oracle for the BAOrec.jl hot path -- TEST INFRASTRUCTURE ONLY.

This module is a numpy/scipy restatement of the reference's `Array` (CPU)
methods.  It is the checker for the CUDA library, never the product: only
`tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / reference
arm may import it.  The product path (`baorec.jl_b200/`) never does.

PARITY UNPINNED by reference vectors: the reference ships no runnable test,
no golden vector whose inputs are available (test/bchmk_*.csv were produced
from DESI mocks that are not shipped) and Julia is not installed here, so the
reference itself cannot be run.  The oracle is instead pinned by analytic
known-answer tests in tests/test_oracle_kat.py (plane waves, mass
conservation, beta=0 identities, spectral-vs-finite-difference agreement)
and by physics known-answer tests of whole reconstructions in
tests/test_pk_oracle.py (the reference's own by-eye validation,
test_helpers/simulation.py:36-70, as numbers: the Kaiser quadrupole of a
lognormal redshift-space box is removed, RecSym keeps and RecIso removes the
Kaiser boost, and the radial-line-of-sight + randoms mode does the same for a
radially shifted box -- for IterativeRecon and MultigridRecon).

Every function cites the reference file:line it follows (paths relative to
/root/reference).  Array convention: a Julia `Array{T,3}` A[ix,iy,iz]
(column-major, x fastest) is the numpy C-order array a[iz,iy,ix]; all
cell indices returned here are 0-based.

Precision: every routine is generic in the dtype of its inputs.  float32
reproduces the reference's Float32 arithmetic step by step where bit-exactness
is claimed (cell indices and interpolation weights); float64 is the "truth" run
used for tolerances.
"""
from __future__ import annotations

import os
from dataclasses import dataclass, field as _dcfield
from typing import Optional, Sequence, Tuple

import numpy as np
import scipy.fft as sfft

_WORKERS = os.cpu_count() or 1


class OutOfBoxError(ValueError):
    """A particle falls outside the mesh (the reference raises BoundsError /
    performs an out-of-bounds atomic: src/mas.jl:33-35, 82-84)."""


def _T(a):
    return np.asarray(a).dtype.type


def _vec3(v, T):
    a = np.asarray(v, dtype=T)
    if a.shape == ():
        a = np.repeat(a, 3)
    assert a.shape == (3,)
    return a


# --------------------------------------------------------------------------
# utils.jl
# --------------------------------------------------------------------------
def k_vec(n_xyz: Sequence[int], box_size, T=np.float32):
    """src/utils.jl:3-10.  sample_rate = T(2pi*n/L) (computed in Float64, then
    rounded); kx = rfftfreq(nx, fs) = (0..nx/2) * (fs/nx); ky,kz = fftfreq =
    (0..n/2-1, -n/2..-1) * (fs/n) (AbstractFFTs.Frequencies: integer * T
    multiplier)."""
    L = _vec3(box_size, T)
    out = []
    for a in range(3):
        n = int(n_xyz[a])
        fs = T(2.0 * np.pi * n / np.float64(L[a]))
        mult = T(fs / T(n))
        if a == 0:
            idx = np.arange(n // 2 + 1)
        else:
            nn = (n + 1) >> 1
            idx = np.concatenate([np.arange(nn), np.arange(nn - n, 0)])
        out.append((idx.astype(T) * mult).astype(T))
    return tuple(out)


def x_vec(n_xyz: Sequence[int], box_size, box_min, T=np.float32):
    """src/utils.jl:21-25.  Cell-centre coordinates.  cell = L/n in T, the range
    `min + 0.5*cell : cell : min + L` is Float64 (the 0.5 literal promotes) and
    each element is then rounded to T."""
    L = _vec3(box_size, T)
    mn = _vec3(box_min, T)
    out = []
    for a in range(3):
        n = int(n_xyz[a])
        cell = T(L[a] / T(n))
        start = np.float64(mn[a]) + 0.5 * np.float64(cell)
        out.append((start + np.arange(n, dtype=np.float64) * np.float64(cell)).astype(T))
    return tuple(out)


def setup_box(pos_x, pos_y, pos_z, box_pad):
    """src/utils.jl:100-109.  Cubic box: min - pad/2, size = max extent + pad."""
    T = _T(pos_x)
    pad = T(box_pad)
    mn = np.array([p.min() for p in (pos_x, pos_y, pos_z)], dtype=T) - pad / T(2)
    mx = np.array([p.max() for p in (pos_x, pos_y, pos_z)], dtype=T) + pad / T(2)
    size = np.full(3, (mx - mn).max(), dtype=T)
    return size, mn


def rfft(a):
    """plan_rfft * a (src/recon.jl:35-37): unnormalised forward R2C, x halved."""
    return sfft.rfftn(a, axes=(0, 1, 2), workers=_WORKERS)


def irfft(ak, shape):
    """ldiv!(y, plan, x): inverse C2R scaled by 1/M."""
    return sfft.irfftn(ak, s=shape, axes=(0, 1, 2), workers=_WORKERS)


def _k2(kv, T):
    kx, ky, kz = kv
    return ((kx * kx)[None, None, :] + (ky * ky)[None, :, None]) + (kz * kz)[:, None, None]


def smooth(fld, smoothing_radius, box_size):
    """src/utils.jl:45-56.  field_k *= exp(-0.5 R^2 k^2) with k^2 in T and the
    exponent / exp / product in Float64, then stored back to Complex{T}."""
    T = _T(fld)
    nz, ny, nx = fld.shape
    fk = rfft(fld)
    kv = k_vec((nx, ny, nz), box_size, T)
    k2 = _k2(kv, T)
    R2 = T(T(smoothing_radius) * T(smoothing_radius))
    g = np.exp(-0.5 * np.float64(R2) * k2.astype(np.float64))
    fk = (fk.astype(np.complex128) * g).astype(fk.dtype)
    fld[...] = irfft(fk, fld.shape).astype(T)
    return fld


# --------------------------------------------------------------------------
# mas.jl  (mass assignment)
# --------------------------------------------------------------------------
def cic_cells(x, y, z, n_xyz, box_size, box_min, wrap=True):
    """Index / weight part of cic! (src/mas.jl:7-35).  Returns
    (wrapped positions (3 arrays), i0[3][N], i1[3][N], w0[3][N], w1[3][N]) with
    0-based indices and the *unweighted* interpolation weights.
    Quirk kept: the wrap test uses box_min[1]/box_size[1] for all axes."""
    T = _T(x)
    L = _vec3(box_size, T)
    mn = _vec3(box_min, T)
    pos = []
    for p in (x, y, z):
        p = np.asarray(p, dtype=T)
        if wrap:
            p = np.where((p - mn[0]) > L[0], (p - L[0]).astype(T), p)
        pos.append(p.astype(T))
    i0, i1, w0, w1 = [], [], [], []
    for a in range(3):
        n = int(n_xyz[a])
        g = ((pos[a] - mn[a]) * T(n)).astype(T)
        g = (g / L[a]).astype(T)
        g = (g + T(1)).astype(T)
        f0 = np.floor(g)
        c0 = f0.astype(np.int64)              # 1-based like the reference
        wa1 = (g - f0).astype(T)
        wa0 = (T(1) - wa1).astype(T)
        c0 = np.where(c0 == n + 1, 1, c0)
        if wrap:
            c1 = np.where(c0 == n, 1, c0 + 1)
        else:
            c1 = c0 + 1
        i0.append(c0 - 1)
        i1.append(c1 - 1)
        w0.append(wa0)
        w1.append(wa1)
    return pos, i0, i1, w0, w1


def cic_scatter(rho, x, y, z, w, box_size, box_min, wrap=True):
    """cic! CPU method (src/mas.jl:1-52): serial loop over particles, deposits
    in the order 000,100,010,001,110,101,011,111 with weights (wx*wy)*wz where
    wx already carries the particle weight.  Mutates x,y,z when wrap (:8-10).
    rho is accumulated into (caller zero-fills)."""
    T = _T(rho)
    nz, ny, nx = rho.shape
    pos, i0, i1, w0, w1 = cic_cells(x, y, z, (nx, ny, nz), box_size, box_min, wrap)
    for a, n in enumerate((nx, ny, nz)):
        if (i0[a] < 0).any() or (i0[a] >= n).any() or (i1[a] >= n).any():
            raise OutOfBoxError("particle outside the mesh (reference: BoundsError, src/mas.jl:33-49)")
    if wrap:
        x[...] = pos[0]
        y[...] = pos[1]
        z[...] = pos[2]
    w = np.asarray(w, dtype=T)
    wx0 = (w0[0] * w).astype(T)
    wx1 = (w1[0] * w).astype(T)
    wy0, wy1, wz0, wz1 = w0[1], w1[1], w0[2], w1[2]
    X0, X1, Y0, Y1, Z0, Z1 = i0[0], i1[0], i0[1], i1[1], i0[2], i1[2]
    lin = lambda zz, yy, xx: (zz * ny + yy) * nx + xx
    idx = np.stack([lin(Z0, Y0, X0), lin(Z0, Y0, X1), lin(Z0, Y1, X0), lin(Z1, Y0, X0),
                    lin(Z0, Y1, X1), lin(Z1, Y0, X1), lin(Z1, Y1, X0), lin(Z1, Y1, X1)], axis=1)
    val = np.stack([(wx0 * wy0).astype(T) * wz0, (wx1 * wy0).astype(T) * wz0,
                    (wx0 * wy1).astype(T) * wz0, (wx0 * wy0).astype(T) * wz1,
                    (wx1 * wy1).astype(T) * wz0, (wx1 * wy0).astype(T) * wz1,
                    (wx0 * wy1).astype(T) * wz1, (wx1 * wy1).astype(T) * wz1], axis=1).astype(T)
    np.add.at(rho.reshape(-1), idx.reshape(-1), val.reshape(-1))   # particle-major serial order
    return rho


def gather_cells(x, y, z, n_xyz, box_size, box_min, wrap=True, formula="cpu"):
    """Index / weight part of read_cic! (src/mas.jl:220-255 CPU; :274-306 GPU).
    formula="cpu": d = (p-min)/cell, cell = T(L/n)   (the parity target)
    formula="gpu": d = (p-min)*n/L
    Returns (id[3][N], iu[3][N], wd[3][N], wu[3][N]), 0-based."""
    T = _T(x)
    L = _vec3(box_size, T)
    mn = _vec3(box_min, T)
    idn, iup, wd, wu = [], [], [], []
    for a, p in enumerate((x, y, z)):
        n = int(n_xyz[a])
        p = np.asarray(p, dtype=T)
        if formula == "cpu":
            cell = T(L[a] / T(n))
            d = ((p - mn[a]).astype(T) / cell).astype(T)
        else:
            d = (((p - mn[a]).astype(T) * T(n)).astype(T) / L[a]).astype(T)
        f = np.floor(d)
        u = (d - f).astype(T)
        dl = (T(1) - u).astype(T)
        i = f.astype(np.int64) + 1
        if wrap:
            i = np.where(i > n, i - n, i)
        j = i + 1
        if wrap:
            j = np.where(j > n, j - n, j)
        idn.append(i - 1)
        iup.append(j - 1)
        wd.append(dl)
        wu.append(u)
    return idn, iup, wd, wu


def read_cic(fld, x, y, z, box_size, box_min, wrap=True, formula="cpu"):
    """read_cic! (src/mas.jl:218-269): trilinear gather, sum of 8 products
    field*wx*wy*wz left-associated in the order ddd,ddu,dud,duu,udd,udu,uud,uuu
    (letters = x,y,z)."""
    T = _T(fld)
    nz, ny, nx = fld.shape
    idn, iup, wd, wu = gather_cells(x, y, z, (nx, ny, nz), box_size, box_min, wrap, formula)
    for a, n in enumerate((nx, ny, nz)):
        if (idn[a] < 0).any() or (idn[a] >= n).any() or (iup[a] >= n).any():
            raise OutOfBoxError("particle outside the mesh in read_cic (reference: @inbounds UB)")
    (xd, yd, zd), (xu, yu, zu) = idn, iup
    (dx, dy, dz), (ux, uy, uz) = wd, wu

    def term(ix, iy, iz, a, b, c):
        return (((fld[iz, iy, ix] * a).astype(T) * b).astype(T) * c).astype(T)

    out = term(xd, yd, zd, dx, dy, dz)
    out = (out + term(xd, yd, zu, dx, dy, uz)).astype(T)
    out = (out + term(xd, yu, zd, dx, uy, dz)).astype(T)
    out = (out + term(xd, yu, zu, dx, uy, uz)).astype(T)
    out = (out + term(xu, yd, zd, ux, dy, dz)).astype(T)
    out = (out + term(xu, yd, zu, ux, dy, uz)).astype(T)
    out = (out + term(xu, yu, zd, ux, uy, dz)).astype(T)
    out = (out + term(xu, yu, zu, ux, uy, uz)).astype(T)
    return out


# ---- TSC (extension: NOT in the reference, which only has CIC; BASELINE config 2
# asks for it.  Same grid convention as cic!: mesh points are cell corners at
# min + i*cell.  Parity is unpinned by definition.) -------------------------
def tsc_cells(x, y, z, n_xyz, box_size, box_min, wrap=True):
    """Nearest grid point ic = floor(g + 0.5), d = g - ic with g = (p-min)*n/L;
    weights w(-1) = 0.5(0.5-d)^2, w(0) = 0.75-d^2, w(+1) = 0.5(0.5+d)^2.
    Returns (ic[3][N] 0-based *unwrapped* centre index, wm, w0, wp)."""
    T = _T(x)
    L = _vec3(box_size, T)
    mn = _vec3(box_min, T)
    ic, wm, wc, wp = [], [], [], []
    for a, p in enumerate((x, y, z)):
        n = int(n_xyz[a])
        p = np.asarray(p, dtype=T)
        g = (((p - mn[a]).astype(T) * T(n)).astype(T) / L[a]).astype(T)
        c = np.floor((g + T(0.5)).astype(T))
        d = (g - c).astype(T)
        hm = (T(0.5) - d).astype(T)
        hp = (T(0.5) + d).astype(T)
        wm.append((T(0.5) * (hm * hm).astype(T)).astype(T))
        wc.append((T(0.75) - (d * d).astype(T)).astype(T))
        wp.append((T(0.5) * (hp * hp).astype(T)).astype(T))
        ic.append(c.astype(np.int64))
    return ic, wm, wc, wp


def _tsc_axis_index(c, off, n, wrap):
    i = c + off
    if wrap:
        return np.mod(i, n)
    if (i < 0).any() or (i >= n).any():
        raise OutOfBoxError("TSC stencil leaves the mesh with wrap=False")
    return i


def tsc_scatter(rho, x, y, z, w, box_size, box_min, wrap=True):
    T = _T(rho)
    nz, ny, nx = rho.shape
    ic, wm, wc, wp = tsc_cells(x, y, z, (nx, ny, nz), box_size, box_min, wrap)
    w = np.asarray(w, dtype=T)
    W = [(wm[a], wc[a], wp[a]) for a in range(3)]
    flat = rho.reshape(-1)
    for oz in range(3):
        iz = _tsc_axis_index(ic[2], oz - 1, nz, wrap)
        for oy in range(3):
            iy = _tsc_axis_index(ic[1], oy - 1, ny, wrap)
            for ox in range(3):
                ix = _tsc_axis_index(ic[0], ox - 1, nx, wrap)
                val = (((W[0][ox] * w).astype(T) * W[1][oy]).astype(T) * W[2][oz]).astype(T)
                np.add.at(flat, (iz * ny + iy) * nx + ix, val)
    return rho


def read_tsc(fld, x, y, z, box_size, box_min, wrap=True):
    T = _T(fld)
    nz, ny, nx = fld.shape
    ic, wm, wc, wp = tsc_cells(x, y, z, (nx, ny, nz), box_size, box_min, wrap)
    W = [(wm[a], wc[a], wp[a]) for a in range(3)]
    out = np.zeros(len(x), dtype=T)
    for oz in range(3):
        iz = _tsc_axis_index(ic[2], oz - 1, nz, wrap)
        for oy in range(3):
            iy = _tsc_axis_index(ic[1], oy - 1, ny, wrap)
            for ox in range(3):
                ix = _tsc_axis_index(ic[0], ox - 1, nx, wrap)
                out = (out + (((fld[iz, iy, ix] * W[0][ox]).astype(T) * W[1][oy]).astype(T) * W[2][oz]).astype(T)).astype(T)
    return out


# PCS (piecewise cubic spline; SURVEY.md section 8f N3 -- the fourth assignment scheme of the power-spectrum tool the
# reference's helpers configure, test_helpers/powspec_auto.conf:117-124; not in src/mas.jl).  Same grid convention as
# cic! / tsc: mesh points at min + i * cell.  The coordinate is cic!'s own (src/mas.jl:13-19): g = (p - min) n / L + 1
# (1-based), c = floor(g) - 1 the 0-based base point, t = g - floor(g) -- cic!'s upper weight, so the base point of a
# particle is the same cell under both schemes, last bit included (the slab decomposition assigns particles by that
# cell).  u = 1 - t; the four points c-1 .. c+2 carry the cubic B-spline  w(-1) = u^3/6,  w(0) = ((3t - 6) t^2 + 4)/6,
# w(+1) = ((3u - 6) u^2 + 4)/6,  w(+2) = t^3/6, every operation rounded to T in the order written (the device kernels
# spell out the same sequence).
def pcs_cells(x, y, z, n_xyz, box_size, box_min, wrap=True):
    """Returns (c[3][N] 0-based *unwrapped* floor index, [w(-1), w(0), w(+1), w(+2)] per axis)."""
    T = _T(x)
    L = _vec3(box_size, T)
    mn = _vec3(box_min, T)
    ic, W = [], []
    for a, p in enumerate((x, y, z)):
        n = int(n_xyz[a])
        p = np.asarray(p, dtype=T)
        g = ((((p - mn[a]).astype(T) * T(n)).astype(T) / L[a]).astype(T) + T(1)).astype(T)
        f0 = np.floor(g)
        c = f0 - 1
        t = (g - f0).astype(T)
        u = (T(1) - t).astype(T)
        t2 = (t * t).astype(T)
        u2 = (u * u).astype(T)
        w_m = ((u2 * u).astype(T) / T(6)).astype(T)
        w_0 = (((((T(3) * t).astype(T) - T(6)).astype(T) * t2).astype(T) + T(4)).astype(T) / T(6)).astype(T)
        w_1 = (((((T(3) * u).astype(T) - T(6)).astype(T) * u2).astype(T) + T(4)).astype(T) / T(6)).astype(T)
        w_2 = ((t2 * t).astype(T) / T(6)).astype(T)
        W.append((w_m, w_0, w_1, w_2))
        ic.append(c.astype(np.int64))
    return ic, W


def pcs_scatter(rho, x, y, z, w, box_size, box_min, wrap=True):
    T = _T(rho)
    nz, ny, nx = rho.shape
    ic, W = pcs_cells(x, y, z, (nx, ny, nz), box_size, box_min, wrap)
    w = np.asarray(w, dtype=T)
    flat = rho.reshape(-1)
    for oz in range(4):
        iz = _tsc_axis_index(ic[2], oz - 1, nz, wrap)
        for oy in range(4):
            iy = _tsc_axis_index(ic[1], oy - 1, ny, wrap)
            for ox in range(4):
                ix = _tsc_axis_index(ic[0], ox - 1, nx, wrap)
                val = (((W[0][ox] * w).astype(T) * W[1][oy]).astype(T) * W[2][oz]).astype(T)
                np.add.at(flat, (iz * ny + iy) * nx + ix, val)
    return rho


def read_pcs(fld, x, y, z, box_size, box_min, wrap=True):
    T = _T(fld)
    nz, ny, nx = fld.shape
    ic, W = pcs_cells(x, y, z, (nx, ny, nz), box_size, box_min, wrap)
    out = np.zeros(len(x), dtype=T)
    for oz in range(4):
        iz = _tsc_axis_index(ic[2], oz - 1, nz, wrap)
        for oy in range(4):
            iy = _tsc_axis_index(ic[1], oy - 1, ny, wrap)
            for ox in range(4):
                ix = _tsc_axis_index(ic[0], ox - 1, nx, wrap)
                out = (out + (((fld[iz, iy, ix] * W[0][ox]).astype(T) * W[1][oy]).astype(T) * W[2][oz]).astype(T)).astype(T)
    return out


# --------------------------------------------------------------------------
# recon.jl : parameter bags and overdensity set-up
# --------------------------------------------------------------------------
@dataclass
class IterativeRecon:
    """src/recon.jl:2-16."""
    bias: float
    f: float
    smoothing_radius: float
    box_size: Optional[np.ndarray] = None
    box_min: Optional[np.ndarray] = None
    los: Optional[Sequence[float]] = None
    n_iter: int = 3
    beta: Optional[float] = None
    mas: str = "cic"            # extension ("tsc", "pcs"); the reference hard-codes cic!
    result_cache: Optional[np.ndarray] = _dcfield(default=None, repr=False)

    def __post_init__(self):
        if self.beta is None:
            self.beta = np.float32(np.float32(self.f) / np.float32(self.bias))


@dataclass
class MultigridRecon:
    """src/recon.jl:18-33."""
    bias: float
    f: float
    smoothing_radius: float
    box_size: Optional[np.ndarray] = None
    box_min: Optional[np.ndarray] = None
    los: Optional[Sequence[float]] = None
    jacobi_damping_factor: float = 0.4
    jacobi_niterations: int = 5
    vcycle_niterations: int = 6
    beta: Optional[float] = None
    mas: str = "cic"
    result_cache: Optional[np.ndarray] = _dcfield(default=None, repr=False)

    def __post_init__(self):
        if self.beta is None:
            self.beta = np.float32(np.float32(self.f) / np.float32(self.bias))


def _scatter(recon, rho, x, y, z, w, wrap):
    if getattr(recon, "mas", "cic") == "tsc":
        return tsc_scatter(rho, x, y, z, w, recon.box_size, recon.box_min, wrap)
    if getattr(recon, "mas", "cic") == "pcs":
        return pcs_scatter(rho, x, y, z, w, recon.box_size, recon.box_min, wrap)
    return cic_scatter(rho, x, y, z, w, recon.box_size, recon.box_min, wrap)


def setup_overdensity(delta, recon, x, y, z, w, rx=None, ry=None, rz=None, rw=None,
                      wrap=True, ran_min=0.01, force_mask=None, info=None):
    """src/recon.jl:42-57 (no randoms) and :60-91 (randoms).

    Test hooks (not in the reference): `force_mask` replaces the `ran > threshold` decision by a
    given boolean mesh (the cut is a discontinuity: cells whose smoothed randoms density sits
    within fp32 FFT noise of the threshold legitimately flip between implementations); `info`, if
    a dict, receives the smoothed randoms mesh and the threshold."""
    T = _T(delta)
    if rx is None:
        _scatter(recon, delta, x, y, z, w, wrap)
        smooth(delta, T(recon.smoothing_radius), recon.box_size)
        mean = T(delta.sum(dtype=T) / T(delta.size))
        delta[...] = (((delta / mean).astype(T) - T(1)).astype(T) / T(recon.bias)).astype(T)
        return delta
    ran = np.zeros_like(delta)
    _scatter(recon, delta, x, y, z, w, False)
    _scatter(recon, ran, rx, ry, rz, rw, False)
    smooth(delta, T(recon.smoothing_radius), recon.box_size)
    smooth(ran, T(recon.smoothing_radius), recon.box_size)
    sd = T(delta.sum(dtype=T))
    sr = T(ran.sum(dtype=T))
    alpha = T(sd / sr)
    thr = np.float64(ran_min) * np.float64(sr) / len(rx)      # Float64: ran_min is a Float64 literal (:70)
    delta -= (alpha * ran).astype(T)
    with np.errstate(divide="ignore", invalid="ignore"):
        q = (delta / ((T(recon.bias) * alpha).astype(T) * ran).astype(T)).astype(T)
    mask = (ran.astype(np.float64) > thr) if force_mask is None else force_mask
    if info is not None:
        info.update(ran=ran, threshold=thr, alpha=alpha)
    delta[...] = np.where(mask, q, T(0))
    return delta


# --------------------------------------------------------------------------
# iterative.jl
# --------------------------------------------------------------------------
def iterate(delta_r, delta_s, kv, it, beta, los=None, xv=None):
    """iterate! CPU method (src/iterative.jl:5-67).  `it` is 1-based."""
    T = _T(delta_r)
    beta = T(beta)
    shape = delta_r.shape
    dk = rfft(delta_r)
    k2 = _k2(kv, T)
    k2s = np.where(k2 == 0, T(1), k2)
    dk = (dk / k2s).astype(dk.dtype)
    dk[0, 0, 0] = 0
    delta_r[...] = delta_s
    kb = (kv[0][None, None, :], kv[1][None, :, None], kv[2][:, None, None])
    if los is None:
        assert xv is not None
        xb = (xv[0][None, None, :], xv[1][None, :, None], xv[2][:, None, None])
        x2 = ((xb[0] * xb[0]) + (xb[1] * xb[1])).astype(T) + (xb[2] * xb[2])
        x2 = x2.astype(T)
        for i in range(3):
            for j in range(i, 3):
                fac = T((1.0 + float(i != j)) * np.float64(beta))
                if it == 1:
                    fac = T(fac / (T(1) + beta))
                dpk = ((kb[i] * kb[j]).astype(T) * dk).astype(dk.dtype)
                dpx = irfft(dpk, shape).astype(T)
                with np.errstate(divide="ignore", invalid="ignore"):
                    upd = (delta_r - (((fac * dpx).astype(T) * xb[i]).astype(T) * xb[j]).astype(T) / x2).astype(T)
                delta_r[...] = np.where(x2 > 0, upd, T(0))
    else:
        los = np.asarray(los, dtype=T)
        for i in range(3):
            if los[i] == 0:
                continue
            fac = beta
            if it == 1:
                fac = T(fac / (T(1) + beta))
            dpk = (((kb[i] * kb[i]).astype(T) * los[i]).astype(T) * dk).astype(dk.dtype)
            dpx = irfft(dpk, shape).astype(T)
            delta_r[...] = (delta_r - (fac * dpx).astype(T)).astype(T)
    return delta_r


def reconstructed_overdensity(delta, recon: IterativeRecon, x, y, z, w, rx=None, ry=None, rz=None, rw=None, force_mask=None):
    """src/recon.jl:93-132."""
    T = _T(delta)
    nz, ny, nx = delta.shape
    setup_overdensity(delta, recon, x, y, z, w, rx, ry, rz, rw, force_mask=force_mask)
    delta_s = delta.copy()
    kv = k_vec((nx, ny, nz), recon.box_size, T)
    xv = x_vec((nx, ny, nz), recon.box_size, recon.box_min, T) if recon.los is None else None
    for it in range(1, recon.n_iter + 1):
        iterate(delta, delta_s, kv, it, recon.beta, recon.los, xv)
    return delta


def compute_displacements_iterative(delta, x, y, z, recon, formula="cpu"):
    """src/iterative.jl:253-275: Psi_a = irfft(i k_a delta_k / k^2), gathered."""
    T = _T(delta)
    nz, ny, nx = delta.shape
    kv = k_vec((nx, ny, nz), recon.box_size, T)
    dk = rfft(delta)
    k2 = _k2(kv, T)
    kb = (kv[0][None, None, :], kv[1][None, :, None], kv[2][:, None, None])
    out = []
    for a in range(3):
        with np.errstate(divide="ignore", invalid="ignore"):
            pk = ((1j * kb[a]).astype(dk.dtype) * dk / k2).astype(dk.dtype)
        pk = np.where(k2 > 0, pk, 0).astype(dk.dtype)
        pr = irfft(pk, delta.shape).astype(T)
        out.append(_read(recon, pr, x, y, z, formula))
    return out


def _read(recon, fld, x, y, z, formula="cpu"):
    if getattr(recon, "mas", "cic") == "tsc":
        return read_tsc(fld, x, y, z, recon.box_size, recon.box_min, True)
    if getattr(recon, "mas", "cic") == "pcs":
        return read_pcs(fld, x, y, z, recon.box_size, recon.box_min, True)
    return read_cic(fld, x, y, z, recon.box_size, recon.box_min, True, formula)


def displacement_meshes(mesh, recon):
    """The three real-space displacement meshes (before the gather)."""
    T = _T(mesh)
    nz, ny, nx = mesh.shape
    kv = k_vec((nx, ny, nz), recon.box_size, T)
    dk = rfft(mesh)
    k2 = _k2(kv, T)
    kb = (kv[0][None, None, :], kv[1][None, :, None], kv[2][:, None, None])
    out = []
    for a in range(3):
        if isinstance(recon, MultigridRecon):
            pk = ((1j * kb[a]).astype(dk.dtype) * dk).astype(dk.dtype)
        else:
            with np.errstate(divide="ignore", invalid="ignore"):
                pk = ((1j * kb[a]).astype(dk.dtype) * dk / k2).astype(dk.dtype)
            pk = np.where(k2 > 0, pk, 0).astype(dk.dtype)
        out.append(irfft(pk, mesh.shape).astype(T))
    return out


def compute_displacements_multigrid(phi, x, y, z, recon, formula="cpu"):
    """src/multigrid.jl:755-773: Psi_a = irfft(i k_a phi_k), gathered."""
    T = _T(phi)
    nz, ny, nx = phi.shape
    kv = k_vec((nx, ny, nz), recon.box_size, T)
    pk0 = rfft(phi)
    kb = (kv[0][None, None, :], kv[1][None, :, None], kv[2][:, None, None])
    out = []
    for a in range(3):
        pk = ((1j * kb[a]).astype(pk0.dtype) * pk0).astype(pk0.dtype)
        pr = irfft(pk, phi.shape).astype(T)
        out.append(_read(recon, pr, x, y, z, formula))
    return out


def read_grad_cic(fld, x, y, z, box_size, box_min):
    """read_grad_cic! -- the finite-difference read-back the reference sketches and leaves commented out
    (src/mas.jl:388-466; call sites src/multigrid.jl:759, 785): CIC interpolation of the central differences
    fld[i+1] - fld[i-1] at the eight corners, divided by 2 cell.  Restated with the INTENDED geometry: the sketch
    locates the particle with the doubled cell size it needs for the difference quotient (dist = (p - min) / (2 L / n)),
    which would halve every coordinate; here the particle's cell is read_cic!'s (src/mas.jl:221-224) and only the final
    division uses 2 L / n.  Accumulation order of the sketch: corners 000, 100, 010, 001, 110, 101, 011, 111
    (x, y, z upper), weight (wx wy) wz, one division at the end.  Always periodic (wrap = true, like every caller)."""
    T = _T(fld)
    nz, ny, nx = fld.shape
    L, mn = _vec3(box_size, T), _vec3(box_min, T)
    idx, lo, up, cells = [], [], [], []
    for a, (p, n) in enumerate(zip((x, y, z), (nx, ny, nz))):
        cell = T(L[a] / T(n))
        d = ((np.asarray(p, dtype=T) - mn[a]).astype(T) / cell).astype(T)
        if not ((d >= 0) & (d < 2 * n)).all():
            raise OutOfBoxError("particle outside the mesh in read_grad_cic")
        f = np.floor(d)
        u = (d - f).astype(T)
        i = f.astype(np.int64) + 1
        i = np.where(i > n, i - n, i)
        iu = np.where(i + 1 > n, i + 1 - n, i + 1)
        ipp = np.where(iu + 1 > n, iu + 1 - n, iu + 1)
        im = np.where(i - 1 < 1, i - 1 + n, i - 1)
        idx.append([im - 1, i - 1, iu - 1, ipp - 1])
        lo.append((T(1) - u).astype(T))
        up.append(u)
        cells.append(cell)
    X, Y, Z = idx
    phi = lambda ix, iy, iz: fld[Z[iz], Y[iy], X[ix]]
    g = [None, None, None]
    for k, (cx, cy, cz) in enumerate(((0, 0, 0), (1, 0, 0), (0, 1, 0), (0, 0, 1), (1, 1, 0), (1, 0, 1), (0, 1, 1), (1, 1, 1))):
        wt = (((up[0] if cx else lo[0]) * (up[1] if cy else lo[1])).astype(T) * (up[2] if cz else lo[2])).astype(T)
        terms = (((phi(2 + cx, 1 + cy, 1 + cz) - phi(cx, 1 + cy, 1 + cz)).astype(T) * wt).astype(T),
                 ((phi(1 + cx, 2 + cy, 1 + cz) - phi(1 + cx, cy, 1 + cz)).astype(T) * wt).astype(T),
                 ((phi(1 + cx, 1 + cy, 2 + cz) - phi(1 + cx, 1 + cy, cz)).astype(T) * wt).astype(T))
        g = [t if k == 0 else (acc + t).astype(T) for acc, t in zip(g, terms)]
    return [(g[a] / (T(2) * cells[a]).astype(T)).astype(T) for a in range(3)]


def compute_displacements(mesh, x, y, z, recon, formula="cpu"):
    if isinstance(recon, MultigridRecon) and getattr(recon, "fd_gradient", False):
        return read_grad_cic(mesh, x, y, z, recon.box_size, recon.box_min)
    if isinstance(recon, MultigridRecon):
        return compute_displacements_multigrid(mesh, x, y, z, recon, formula)
    return compute_displacements_iterative(mesh, x, y, z, recon, formula)


def read_shifts(recon, x, y, z, mesh, field="disp", formula="cpu"):
    """src/recon.jl:264-306.  field in {"disp","rsd","sum"}."""
    T = _T(mesh)
    disp = compute_displacements(mesh, x, y, z, recon, formula)
    if field == "disp":
        return disp
    f = T(recon.f)
    if recon.los is None:
        dist = np.sqrt(((x * x) + (y * y)).astype(T) + (z * z)).astype(T)
        lx, ly, lz = (x / dist).astype(T), (y / dist).astype(T), (z / dist).astype(T)
    else:
        lx, ly, lz = (T(v) for v in recon.los)
    dot = (((disp[0] * lx) + (disp[1] * ly)).astype(T) + (disp[2] * lz)).astype(T)
    rsd = [((f * dot).astype(T) * l).astype(T) for l in (lx, ly, lz)]
    if field == "rsd":
        return rsd
    if field == "sum":
        return [(d + r).astype(T) for d, r in zip(disp, rsd)]
    raise ValueError(field)


def reconstructed_positions(recon, x, y, z, mesh=None, field="disp", formula="cpu"):
    """src/recon.jl:366-388: pos - shift, no periodic re-wrap."""
    if mesh is None:
        mesh = recon.result_cache
    s = read_shifts(recon, x, y, z, mesh, field, formula)
    return [(p - d).astype(p.dtype) for p, d in zip((x, y, z), s)]


# --------------------------------------------------------------------------
# multigrid.jl
# --------------------------------------------------------------------------
def _nb(v, dz, dy, dx):
    """v[ix+dx, iy+dy, iz+dz] with periodic wrap."""
    return np.roll(v, (-dz, -dy, -dx), axis=(0, 1, 2))


def _mg_coeffs(shape, xv, box_size, beta, los, T):
    """Coefficient block shared by jacobi!/residual! (src/multigrid.jl:17-20, 55-62)."""
    nz, ny, nx = shape
    L = _vec3(box_size, T)
    cell = np.array([T(L[0] / T(nx)), T(L[1] / T(ny)), T(L[2] / T(nz))], dtype=T)
    cell2 = (cell * cell).astype(T)
    icell2 = (T(1) / cell2).astype(T)
    if los is None:
        px = (xv[0] / cell[0]).astype(T)[None, None, :]
        py = (xv[1] / cell[1]).astype(T)[None, :, None]
        pz = (xv[2] / cell[2]).astype(T)[:, None, None]
    else:
        ln = (np.asarray(los, dtype=T) / cell).astype(T)
        px, py, pz = (np.full((1, 1, 1), ln[a], dtype=T) for a in range(3))
    den = ((cell2[0] * (px * px)).astype(T) + (cell2[1] * (py * py)).astype(T)).astype(T) + (cell2[2] * (pz * pz)).astype(T)
    g = (T(beta) / den.astype(T)).astype(T)
    gpx2 = (icell2[0] + (g * (px * px)).astype(T)).astype(T)
    gpy2 = (icell2[1] + (g * (py * py)).astype(T)).astype(T)
    gpz2 = (icell2[2] + (g * (pz * pz)).astype(T)).astype(T)
    return px, py, pz, g, gpx2, gpy2, gpz2


def _offdiag(v, px, py, pz, g, gpx2, gpy2, gpz2, radial, T):
    """Sum of the off-diagonal stencil terms (src/multigrid.jl:65-79 / :263-276)."""
    xp, xm = _nb(v, 0, 0, 1), _nb(v, 0, 0, -1)
    yp, ym = _nb(v, 0, 1, 0), _nb(v, 0, -1, 0)
    zp, zm = _nb(v, 1, 0, 0), _nb(v, -1, 0, 0)
    s = gpx2 * (xp + xm) + gpy2 * (yp + ym) + gpz2 * (zp + zm)
    cxy = _nb(v, 0, 1, 1) + _nb(v, 0, -1, -1) - _nb(v, 0, 1, -1) - _nb(v, 0, -1, 1)
    cxz = _nb(v, 1, 0, 1) + _nb(v, -1, 0, -1) - _nb(v, 1, 0, -1) - _nb(v, -1, 0, 1)
    cyz = _nb(v, 1, 1, 0) + _nb(v, -1, -1, 0) - _nb(v, 1, -1, 0) - _nb(v, -1, 1, 0)
    s = s + g / T(2) * (px * py * cxy + px * pz * cxz + py * pz * cyz)
    if radial:
        s = s + g * (px * (xp - xm) + py * (yp - ym) + pz * (zp - zm))
    return s.astype(T)


def jacobi(v, f, xv, box_size, box_min, beta, damping_factor, niterations, los=None):
    """jacobi! (src/multigrid.jl:14-135): niterations damped-Jacobi sweeps."""
    T = _T(v)
    px, py, pz, g, gpx2, gpy2, gpz2 = _mg_coeffs(v.shape, xv, box_size, beta, los, T)
    diag = (T(2) * (gpx2 + gpy2 + gpz2)).astype(T)
    w = T(damping_factor)
    for _ in range(niterations):
        jac = ((f + _offdiag(v, px, py, pz, g, gpx2, gpy2, gpz2, los is None, T)) / diag).astype(T)
        v[...] = ((T(1) - w) * v + w * jac).astype(T)
    return v


def residual(v, f, xv, box_size, box_min, beta, los=None):
    """residual! (src/multigrid.jl:213-326): r = f - L v."""
    T = _T(v)
    px, py, pz, g, gpx2, gpy2, gpz2 = _mg_coeffs(v.shape, xv, box_size, beta, los, T)
    diag = (T(2) * (gpx2 + gpy2 + gpz2)).astype(T)
    r = (diag * v - _offdiag(v, px, py, pz, g, gpx2, gpy2, gpz2, los is None, T)).astype(T)
    return (f - r).astype(T)


def restrict(v1h):
    """reduce! (src/multigrid.jl:520-584): 27-point full weighting; coarse c
    (0-based) is centred on fine 2c+1; weights 8/4/2/1, /64."""
    T = _T(v1h)
    acc = np.zeros(tuple(s // 2 for s in v1h.shape), dtype=T)
    wts = {0: T(8), 1: T(4), 2: T(2), 3: T(1)}
    for dz in (-1, 0, 1):
        for dy in (-1, 0, 1):
            for dx in (-1, 0, 1):
                wgt = wts[abs(dz) + abs(dy) + abs(dx)]
                acc += wgt * _nb(v1h, dz, dy, dx)[1::2, 1::2, 1::2]
    return (acc / T(64)).astype(T)


def prolong(v1h, v2h):
    """prolong! (src/multigrid.jl:398-453): trilinear; fine 2c+1 <- coarse c,
    fine (2c+2) mod n <- mean of coarse c and c+1 (per axis).  Overwrites all of v1h."""
    T = _T(v2h)
    ncz, ncy, ncx = v2h.shape
    nfz, nfy, nfx = v1h.shape
    C = v2h
    sh = lambda dz, dy, dx: _nb(C, dz, dy, dx)
    zo, yo, xo = 2 * np.arange(ncz) + 1, 2 * np.arange(ncy) + 1, 2 * np.arange(ncx) + 1
    ze, ye, xe = (zo + 1) % nfz, (yo + 1) % nfy, (xo + 1) % nfx
    v1h[np.ix_(zo, yo, xo)] = C
    v1h[np.ix_(zo, yo, xe)] = ((C + sh(0, 0, 1)) / T(2)).astype(T)
    v1h[np.ix_(zo, ye, xo)] = ((C + sh(0, 1, 0)) / T(2)).astype(T)
    v1h[np.ix_(ze, yo, xo)] = ((C + sh(1, 0, 0)) / T(2)).astype(T)
    v1h[np.ix_(zo, ye, xe)] = ((C + sh(0, 0, 1) + sh(0, 1, 0) + sh(0, 1, 1)) / T(4)).astype(T)
    v1h[np.ix_(ze, ye, xo)] = ((C + sh(0, 1, 0) + sh(1, 0, 0) + sh(1, 1, 0)) / T(4)).astype(T)
    v1h[np.ix_(ze, yo, xe)] = ((C + sh(0, 0, 1) + sh(1, 0, 0) + sh(1, 0, 1)) / T(4)).astype(T)
    v1h[np.ix_(ze, ye, xe)] = ((C + sh(0, 0, 1) + sh(0, 1, 0) + sh(1, 0, 0)
                                + sh(0, 1, 1) + sh(1, 0, 1) + sh(1, 1, 0) + sh(1, 1, 1)) / T(8)).astype(T)
    return v1h


def _recurse_ok(shape):
    return all((n > 4) and (n % 2 == 0) for n in shape)


def vcycle(v, f, box_size, box_min, beta, damping_factor, niterations, los=None):
    """vcycle! (src/multigrid.jl:654-687)."""
    T = _T(v)
    nz, ny, nx = v.shape
    xv = x_vec((nx, ny, nz), box_size, box_min, T)
    jacobi(v, f, xv, box_size, box_min, beta, damping_factor, niterations, los)
    if _recurse_ok(v.shape):
        r = residual(v, f, xv, box_size, box_min, beta, los)
        f2h = restrict(r)
        v2h = np.zeros_like(f2h)
        vcycle(v2h, f2h, box_size, box_min, beta, damping_factor, niterations, los)
        v1h = np.zeros_like(v)
        prolong(v1h, v2h)
        v += v1h
    jacobi(v, f, xv, box_size, box_min, beta, damping_factor, niterations, los)
    return v


def fmg(f1h, v1h, box_size, box_min, beta, damping_factor, jacobi_niterations, vcycle_niterations, los=None):
    """fmg (src/multigrid.jl:722-752)."""
    if _recurse_ok(f1h.shape):
        f2h = restrict(f1h)
        v2h = fmg(f2h, None, box_size, box_min, beta, damping_factor, jacobi_niterations, vcycle_niterations, los)
        if v1h is None:
            v1h = np.zeros_like(f1h)
        prolong(v1h, v2h)
    elif v1h is None:
        v1h = np.zeros_like(f1h)
    for _ in range(vcycle_niterations):
        vcycle(v1h, f1h, box_size, box_min, beta, damping_factor, jacobi_niterations, los)
    return v1h


def reconstructed_potential(phi, recon: MultigridRecon, x, y, z, w, rx=None, ry=None, rz=None, rw=None, force_mask=None):
    """src/recon.jl:184-212."""
    T = _T(phi)
    delta = np.zeros_like(phi)
    setup_overdensity(delta, recon, x, y, z, w, rx, ry, rz, rw, force_mask=force_mask)
    fmg(delta, phi, recon.box_size, recon.box_min, T(recon.beta), T(recon.jacobi_damping_factor),
        recon.jacobi_niterations, recon.vcycle_niterations, recon.los)
    return phi


# --------------------------------------------------------------------------
# run! (src/recon.jl:134-180, 215-261)
# --------------------------------------------------------------------------
def run(recon, grid_size_xyz, x, y, z, w, rx=None, ry=None, rz=None, rw=None, force_mask=None):
    T = _T(x)
    nx, ny, nz = grid_size_xyz
    mesh = np.zeros((nz, ny, nx), dtype=T)
    if rx is not None:
        recon.box_size, recon.box_min = setup_box(rx, ry, rz, T(500))
    if isinstance(recon, MultigridRecon):
        reconstructed_potential(mesh, recon, x, y, z, w, rx, ry, rz, rw, force_mask)
    else:
        reconstructed_overdensity(mesh, recon, x, y, z, w, rx, ry, rz, rw, force_mask)
    recon.result_cache = mesh
    return mesh
