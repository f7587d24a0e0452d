"""Catalog pre/post-processing on the device: the host-side mirror of src/cosmo.jl and of the helper
functions every lightcone example of the reference defines around `run!`
(examples/lightcone.jl:30-82, identical in lightcone_gpu.jl / lightcone_mg*.jl) -- same names, same
argument order -- on top of the C ABI (baorec_cosmo_set, baorec_sky_to_cartesian_f32,
baorec_cartesian_to_sky_f32, baorec_fkp_weights_f32, baorec_wrap_positions_f32).

`Cosmology` is a parameter bag: its derived densities are scalar set-up arithmetic done here exactly
as the reference's `@with_kw` struct does them (Float32 fields, Float64 intermediates); the
comoving-distance table, the interpolations and all per-particle work happen in the CUDA library.
Catalog columns are contiguous 1-D float32 CUDA tensors.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Optional

import numpy as np
import torch

from . import lib_loader as L
from .host import Context, _chk_vec, _ptr, _stream

_f32 = np.float32

# src/cosmo.jl:11-21
_MPC, _C_KM_S, _G, _HBAR, _EVC2, _KB = 3.085677581491367e22, 299792.458, 6.67408e-11, 6.582119514e-16, 1.782661907e-36, 8.6173303e-5


@dataclass
class Cosmology:
    """`BAOrec.Cosmology(; kwargs...)` src/cosmo.jl:22-64 (keyword names transliterated:
    ω_b -> omega_b, Ω_k₀ -> Omega_k0, ...)."""
    omega_b: float = 0.0225
    omega_c: float = 0.12
    h: float = 0.67
    Neff: float = 3.044
    Omega_k0: float = 0.0
    T_cmb: float = 2.725
    w0: float = -1.0
    wa: float = 0.0
    ns: float = 0.96
    z_tab_min: float = 0
    z_tab_max: float = 3
    z_tab_num: int = 100000

    def __post_init__(self):
        T = _f32
        for k in ("omega_b", "omega_c", "h", "Neff", "Omega_k0", "T_cmb", "w0", "wa", "ns"):
            setattr(self, k, T(getattr(self, k)))
        self.h2 = T(self.h * self.h)
        self.H0 = T(self.h * T(100))
        rho_g0 = T(3 * 100 ** 2 / (8 * np.pi * _G) * _HBAR ** 3 * (1.0 / (1e-3 * _MPC)) ** 2 * (1.0 / _EVC2) * (_C_KM_S * 1e3) ** 3)
        self.Omega_g0 = T((np.pi ** 2 / 15) * (2.725 * _KB) ** 4 / float(rho_g0) * (float(self.T_cmb) / 2.725) ** 4 / float(self.h2))
        self.Omega_nu0 = T(float(self.Neff) * (7 / 8 * (4 / 11) ** (4 / 3)) * float(self.Omega_g0))
        self.Omega_b0 = T(self.omega_b / self.h2)
        self.Omega_c0 = T(self.omega_c / self.h2)
        self.Omega_L0 = T(T(T(T(T(T(1) - self.Omega_k0) - self.Omega_b0) - self.Omega_c0) - self.Omega_nu0) - self.Omega_g0)
        self.Omega_m0 = T(self.Omega_b0 + self.Omega_c0)
        self._bound = {}        # device -> True once the tables of THIS cosmology are the context's

    def _params(self) -> L.CosmologyParams:
        return L.CosmologyParams(float(self.h), float(self.Omega_b0), float(self.Omega_c0), float(self.Omega_nu0),
                                 float(self.Omega_g0), float(self.Omega_k0), float(self.Omega_L0), float(self.w0),
                                 float(self.wa), float(self.z_tab_min), float(self.z_tab_max), int(self.z_tab_num))

    def bind(self, ctx: Optional[Context] = None) -> Context:
        """Builds the comoving-distance table in the library (the `cache` of src/cosmo.jl:86-91) if the
        context does not hold this cosmology's yet."""
        ctx = ctx or Context.get()
        if getattr(ctx, "cosmology", None) is not self:
            p = self._params()
            L.check(ctx.lib.baorec_cosmo_set(ctx.handle, C.byref(p)))
            ctx.cosmology = self
        return ctx

    def tables(self, ctx: Optional[Context] = None):
        """(z, r) of the cache as float64 numpy arrays (parity probe)."""
        ctx = self.bind(ctx)
        n = int(self.z_tab_num)
        z, r = np.empty(n, np.float64), np.empty(n, np.float64)
        L.check(ctx.lib.baorec_cosmo_tables(ctx.handle, z.ctypes.data_as(C.POINTER(C.c_double)),
                                            r.ctypes.data_as(C.POINTER(C.c_double)), n))
        return z, r


def DESICosmology(**kwargs) -> Cosmology:
    """src/cosmo.jl:66-68."""
    base = dict(omega_b=0.02237, omega_c=0.1200, h=0.6736, ns=0.9649, Neff=float(_f32(_f32(2.0328) + _f32(1))), w0=-1.0, wa=0.0)
    base.update(kwargs)
    return Cosmology(**base)


# ---- the scalar functions of src/cosmo.jl:70-98 (host set-up arithmetic, Float64 like the reference's quadgk nodes) -------
def E(c: Cosmology, z):
    """`E(c, z)` src/cosmo.jl:70-78: sqrt of the density parameters scaled to redshift z (CPL dark energy)."""
    z = np.asarray(z, np.float64)
    a1 = 1.0 + z
    w0, wa = float(c.w0), float(c.wa)
    fde = -3.0 * (1.0 + w0)
    if wa != 0.0:
        a = 1.0 / a1
        with np.errstate(divide="ignore", invalid="ignore"):
            fde = fde + 3.0 * wa * ((a - 1.0) / np.log(a) - 1.0)
    return np.sqrt((float(c.Omega_nu0) + float(c.Omega_g0)) * a1 ** 4 + (float(c.Omega_b0) + float(c.Omega_c0)) * a1 ** 3
                   + float(c.Omega_k0) * a1 ** 2 + float(c.Omega_L0) * a1 ** (-fde))


def H(c: Cosmology, z):
    """`H(c, z) = c.h * 100 * E(c, z)` src/cosmo.jl:79 [km/s/Mpc]; `c.h * 100` is the Float32 product."""
    return float(_f32(c.h * _f32(100))) * E(c, z)


def _host_table(c: Cosmology, z_min=None, z_max=None, num=None):
    """The library's host-side comoving-distance table (baorec_cosmo_build_table: 7-point Gauss-Legendre per knot
    interval); the cosmology's own table range unless another is given."""
    p = c._params()
    if z_max is not None:
        p.z_tab_min, p.z_tab_max, p.z_tab_num = float(z_min), float(z_max), int(num)
    n = int(p.z_tab_num)
    z, r = np.empty(n, np.float64), np.empty(n, np.float64)
    L.check(L.load().baorec_cosmo_build_table(C.byref(p), z.ctypes.data_as(C.POINTER(C.c_double)),
                                              r.ctypes.data_as(C.POINTER(C.c_double))))
    return z, r


def comoving_distance(c: Cosmology, z):
    """`comoving_distance(c, z)` src/cosmo.jl:80-82: c * int_0^z dz'/H(z') [Mpc], one redshift (the reference: quadgk,
    rtol 1e-8; here the library's quadrature over 4096 sub-intervals of [0, z])."""
    z = float(z)
    if z == 0.0:
        return 0.0
    return float(_host_table(c, 0.0, z, 4097)[1][-1])


def _gridded_linear(knots, values, x, what):
    x = np.asarray(x, np.float64)
    if x.size and (np.any(~(x >= knots[0])) or np.any(~(x <= knots[-1]))):
        raise L.OutOfRangeError(L.ERR_OUT_OF_RANGE, f"{what} outside the tabulated range [{knots[0]}, {knots[-1]}]")
    i = np.clip(np.searchsorted(knots, x, side="right") - 1, 0, len(knots) - 2)
    t = (x - knots[i]) / (knots[i + 1] - knots[i])
    return (1.0 - t) * values[i] + t * values[i + 1]


def comoving_distance_interp(c: Cosmology):
    """`comoving_distance_interp(c)` src/cosmo.jl:83-90: the Gridded(Linear()) interpolator r(z) [Mpc] over the cached
    table, as a host callable (scalars or numpy arrays; the per-particle device path is `sky_to_cartesian`).  Outside the
    table it raises like Interpolations.jl's BoundsError."""
    if getattr(c, "_host_cache", None) is None:
        c._host_cache = _host_table(c)
    z, r = c._host_cache
    return lambda x: _gridded_linear(z, r, x, "redshift")


def redshift_interp(c: Cosmology):
    """`redshift_interp(c)` src/cosmo.jl:91-98: the inverse table z(r), r in Mpc (host callable; device path:
    `cartesian_to_sky`)."""
    if getattr(c, "_host_cache", None) is None:
        c._host_cache = _host_table(c)
    z, r = c._host_cache
    return lambda x: _gridded_linear(r, z, x, "distance")


def sky_to_cartesian(data_cat_ra, data_cat_dec, data_cat_red, cosmo: Cosmology):
    """examples/lightcone.jl:30-49: (ra, dec [deg], redshift) -> (x, y, z) [Mpc/h], three new tensors."""
    n = _chk_vec(data_cat_ra, data_cat_dec, data_cat_red)
    ctx = cosmo.bind(Context.get(data_cat_ra.device.index))
    out = tuple(torch.empty_like(data_cat_ra) for _ in range(3))
    h = float(_f32(cosmo.H0 / _f32(100)))
    L.check(ctx.lib.baorec_sky_to_cartesian_f32(ctx.handle, _ptr(data_cat_ra), _ptr(data_cat_dec), _ptr(data_cat_red), n, h,
                                                _ptr(out[0]), _ptr(out[1]), _ptr(out[2]), _stream()))
    return out


def cartesian_to_sky(data_x, data_y, data_z, cosmo: Cosmology):
    """examples/lightcone.jl:51-80: (x, y, z) [Mpc/h] -> (ra, dec [deg], redshift); ra in (-360, 0] like the reference."""
    n = _chk_vec(data_x, data_y, data_z)
    ctx = cosmo.bind(Context.get(data_x.device.index))
    out = tuple(torch.empty_like(data_x) for _ in range(3))
    L.check(ctx.lib.baorec_cartesian_to_sky_f32(ctx.handle, _ptr(data_x), _ptr(data_y), _ptr(data_z), n, float(cosmo.h),
                                                _ptr(out[0]), _ptr(out[1]), _ptr(out[2]), _stream()))
    return out


def fkp_weights(nz, P0):
    """fkp_weights.(nz, Ref(P0)) examples/lightcone.jl:82,90-91."""
    n = _chk_vec(nz)
    ctx = Context.get(nz.device.index)
    w = torch.empty_like(nz)
    L.check(ctx.lib.baorec_fkp_weights_f32(ctx.handle, _ptr(nz), n, float(_f32(P0)), _ptr(w), _stream()))
    return w


def wrap_positions(pos_x, pos_y, pos_z, box_size, box_min=(0.0, 0.0, 0.0)):
    """In-place periodic re-wrap of (reconstructed) positions into [box_min, box_min + box_size):
    `(pos + box_size) % box_size` of test_helpers/simulation.py:38,51-52, for any box origin."""
    n = _chk_vec(pos_x, pos_y, pos_z)
    ctx = Context.get(pos_x.device.index)
    L.check(ctx.lib.baorec_wrap_positions_f32(ctx.handle, _ptr(pos_x), _ptr(pos_y), _ptr(pos_z), n, L.f3(np.broadcast_to(box_size, 3)),
                                              L.f3(np.broadcast_to(box_min, 3)), _stream()))
    return pos_x, pos_y, pos_z


MAS_POWER = {None: 0, "none": 0, "ngp": 1, "cic": 2, "tsc": 3, "pcs": 4}


def _bins(nx, ny, nz, box_size, kmin, dk, nbins):
    Lb = np.broadcast_to(np.asarray(box_size, np.float64), 3)
    if dk is None:
        dk = 2 * np.pi / float(Lb.max())
    if nbins is None:
        nbins = int((np.pi * min(nx / Lb[0], ny / Lb[1], nz / Lb[2]) - kmin) / dk)
    return float(dk), max(1, min(int(nbins), 1024))


def power_multipoles(rho, box_size, los=(0.0, 0.0, 1.0), kmin=0.0, dk=None, nbins=None, mas="cic", shot=0.0,
                     box_min=(0.0, 0.0, 0.0), randoms=None, rho_shifted=None, randoms_shifted=None):
    """P_0, P_2, P_4 of the density mesh `rho` (device tensor [nz][ny][nx] from `cic`; not modified) for a periodic
    box: the before / after check of test_helpers/simulation.py:56-75 (pypowspec compute_auto_box), on the device.
    mas: "cic" / "tsc" / "pcs" / None selects the window the estimate is compensated for; `shot` (e.g. V / N) is subtracted
    from the monopole.  `randoms`: optional density mesh of a shifted random catalog -> the spectrum of "data minus
    shifted randoms" (compute_auto_box_rand in the reference's helpers: RecIso / RecSym).  `rho_shifted` (and
    `randoms_shifted`): the meshes of the same catalogs painted from `interlace_positions` -> the interlaced estimate
    (GRID_INTERLACE = T, test_helpers/powspec_auto.conf:125).  Defaults: dk = the fundamental 2 pi / max(L), bins up to
    the Nyquist frequency.  Returns dict(k, nmodes, p0, p2, p4) of numpy Float64 arrays (NaN in empty bins)."""
    from .host import _chk_mesh, _plan_for
    nx, ny, nz = _chk_mesh(rho)
    dk, nbins = _bins(nx, ny, nz, box_size, kmin, dk, nbins)
    for m in (randoms, rho_shifted, randoms_shifted):
        if m is not None and _chk_mesh(m) != (nx, ny, nz):
            raise ValueError("every mesh must have the shape of the data mesh")
    if (randoms is None) != (randoms_shifted is None) and rho_shifted is not None:
        raise ValueError("interlaced estimate with randoms: pass both randoms meshes")
    ctx = _plan_for(rho, box_size, box_min)
    out = [np.empty(nbins, np.float64) for _ in range(5)]
    dp = [o.ctypes.data_as(C.POINTER(C.c_double)) for o in out]
    power = MAS_POWER[mas]
    opt = lambda t: _ptr(t) if t is not None else None                       # noqa: E731
    if rho_shifted is None:
        L.check(ctx.lib.baorec_power_multipoles_f32(ctx.handle, _ptr(rho), opt(randoms), L.f3(np.asarray(los, np.float32)), float(kmin), dk,
                                                    nbins, power, float(shot), *dp, _stream()))
    else:
        L.check(ctx.lib.baorec_power_multipoles_interlaced_f32(ctx.handle, _ptr(rho), opt(randoms), _ptr(rho_shifted), opt(randoms_shifted),
                                                               L.f3(np.asarray(los, np.float32)), float(kmin), dk, nbins, power, float(shot),
                                                               *dp, _stream()))
    return dict(k=out[0], nmodes=out[1], p0=out[2], p2=out[3], p4=out[4])


def interlace_positions(pos_x, pos_y, pos_z, grid_size, box_size, box_min=(0.0, 0.0, 0.0)):
    """The catalog of the interlaced mesh: every position half a cell further along each axis, periodic.  Returns new tensors."""
    import torch
    from .host import _plan_grid
    n = _chk_vec(pos_x, pos_y, pos_z)
    ctx = _plan_grid(pos_x.device.index, grid_size, box_size, box_min)
    out = tuple(torch.empty_like(pos_x) for _ in range(3))
    L.check(ctx.lib.baorec_interlace_positions_f32(ctx.handle, _ptr(pos_x), _ptr(pos_y), _ptr(pos_z), n, *(_ptr(o) for o in out), _stream()))
    return out


def compute_auto_box(data_x, data_y, data_z, data_w, box_size, grid_size=(512, 512, 512), mas="tsc", interlace=True, los=(0.0, 0.0, 1.0),
                     kmin=0.0, dk=None, nbins=None, shot=0.0, box_min=(0.0, 0.0, 0.0), rand_x=None, rand_y=None, rand_z=None, rand_w=None):
    """compute_auto_box(x, y, z, w, ...) / compute_auto_box_rand(x, y, z, w, rx, ry, rz, rw, ...) of the reference's
    helpers (test_helpers/simulation.py:36-70 with test_helpers/powspec_auto.conf: 512^3 grid, TSC, interlaced, line of
    sight along z, l = 0, 2, 4) in ONE library call on device catalogs: paint, interlace, transform, bin.  The catalogs
    are not modified.  Returns dict(k, nmodes, p0, p2, p4)."""
    from .host import _plan_grid
    n = _chk_vec(data_x, data_y, data_z, data_w)
    nr = _chk_vec(rand_x, rand_y, rand_z, rand_w) if rand_x is not None else 0
    nx, ny, nz = (int(v) for v in grid_size)
    dk, nbins = _bins(nx, ny, nz, box_size, kmin, dk, nbins)
    ctx = _plan_grid(data_x.device.index, grid_size, box_size, box_min)
    out = [np.empty(nbins, np.float64) for _ in range(5)]
    dp = [o.ctypes.data_as(C.POINTER(C.c_double)) for o in out]
    opt = lambda t: _ptr(t) if t is not None else None                       # noqa: E731
    L.check(ctx.lib.baorec_compute_auto_box_f32(ctx.handle, _ptr(data_x), _ptr(data_y), _ptr(data_z), _ptr(data_w), n, opt(rand_x), opt(rand_y),
                                                opt(rand_z), opt(rand_w), nr, L.MAS[mas], int(bool(interlace)), L.f3(np.asarray(los, np.float32)),
                                                float(kmin), dk, nbins, float(shot), *dp, _stream()))
    return dict(k=out[0], nmodes=out[1], p0=out[2], p2=out[3], p4=out[4])
