"""CPU oracle for the catalog pre/post-processing that brackets every lightcone run
 -- TEST INFRASTRUCTURE ONLY, same rules as baorec_oracle.py.

Restates, in numpy/scipy:
  * the `Cosmology` parameter bag and its derived densities     src/cosmo.jl:22-64
  * E(z), H(z), comoving_distance (adaptive quadrature, rtol 1e-8) src/cosmo.jl:70-85
  * comoving_distance_interp / redshift_interp (gridded linear)   src/cosmo.jl:86-103
  * sky_to_cartesian, cartesian_to_sky, fkp_weights               examples/lightcone.jl:30-82
  * the periodic re-wrap of reconstructed positions               test_helpers/simulation.py:38,51-52

PARITY UNPINNED by reference vectors (Julia is not installed; the reference has no tests for these
helpers).  Pinned analytically in tests/test_catalog_oracle.py: Einstein-de Sitter closed form,
d r/d z = c / H(z), round trips, hand-computed angles.

Precision follows Julia's promotion rules for the Float32 catalogs of the examples: the tables
are Float64; a Float32 redshift is widened before the interpolation; `ra * pi / 180` is Float32;
cos/sin of a Float32 are (in Julia) evaluated through a Float64 kernel and rounded once; the
product `dist * cos(dec) * cos(ra) * h` is Float64 and rounds once when stored.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Optional

import numpy as np
from scipy import integrate

f32 = np.float32

# src/cosmo.jl:11-21
eV = 1.782661907e-36
mpc = 3.085677581491367e22
speed_of_light_km_s = 299792.458
G = 6.67408e-11
hbar = 6.582119514e-16
evc2 = 1.782661907e-36
kb = 8.6173303e-5


@dataclass
class Cosmology:
    """src/cosmo.jl:22-64 (`@with_kw mutable struct Cosmology`): Float32 defaults, T = Float32."""
    omega_b: float = f32(0.0225)
    omega_c: float = f32(0.12)
    h: float = f32(0.67)
    Neff: float = f32(3.044)
    Omega_k0: float = f32(0.0)
    T_cmb: float = f32(2.725)
    w0: float = f32(-1.0)
    wa: float = f32(0.0)
    ns: float = f32(0.96)
    z_tab_min: float = 0
    z_tab_max: float = 3
    z_tab_num: int = 100000
    cache: Optional[dict] = field(default=None, repr=False)

    def __post_init__(self):
        T = f32
        for k in ("omega_b", "omega_c", "h", "Neff", "Omega_k0", "T_cmb", "w0", "wa", "ns"):
            setattr(self, k, T(getattr(self, k)))
        self.h2 = T(self.h * self.h)                                            # :35  h^2 in T
        self.H0 = T(self.h * T(100))                                            # :36
        rho_g0 = T(3 * 100 ** 2 / (8 * np.pi * G) * hbar ** 3 * (1.0 / (1e-3 * mpc)) ** 2 * (1.0 / evc2)
                   * (speed_of_light_km_s * 1e3) ** 3)                          # :40
        # :41  (pi^2/15) (2.725 kb)^4 / rho (T_cmb/2.725)^4 / h^2 ; Float64 literals x Float32 fields -> Float64, then T()
        self.Omega_g0 = T((np.pi ** 2 / 15) * (2.725 * kb) ** 4 / np.float64(rho_g0)
                          * (np.float64(self.T_cmb) / 2.725) ** 4 / np.float64(self.h2))
        self.Omega_nu0 = T(np.float64(self.Neff) * (7 / 8 * (4 / 11) ** (4 / 3)) * np.float64(self.Omega_g0))   # :44
        self.Omega_b0 = T(self.omega_b / self.h2)                               # :48
        self.Omega_c0 = T(self.omega_c / self.h2)                               # :50
        self.Omega_L0 = T(T(T(T(T(T(1) - self.Omega_k0) - self.Omega_b0) - self.Omega_c0) - self.Omega_nu0) - self.Omega_g0)  # :54
        self.Omega_m0 = T(self.Omega_b0 + self.Omega_c0)


def DESICosmology(**kw):
    """src/cosmo.jl:66-68."""
    base = dict(omega_b=f32(0.02237), omega_c=f32(0.1200), h=f32(0.6736), ns=f32(0.9649), Neff=f32(f32(2.0328) + f32(1)),
                w0=f32(-1), wa=f32(0))
    base.update(kw)
    return Cosmology(**base)


def E(c: Cosmology, z):
    """src/cosmo.jl:70-79; z is Float64 in every caller (quadgk nodes)."""
    z = np.asarray(z, np.float64)
    a1 = 1.0 + z
    w0, wa = np.float64(c.w0), np.float64(c.wa)
    if wa == 0.0:
        fde = -3.0 * (1.0 + w0)                      # the 0 * ((a-1)/log a - 1) term (NaN only at a == 1, never a quadrature node)
    else:
        a = 1.0 / a1
        with np.errstate(divide="ignore", invalid="ignore"):
            fde = -3.0 * (1.0 + w0) + 3.0 * wa * ((a - 1.0) / np.log(a) - 1.0)
    om = (np.float64(c.Omega_nu0) * a1 ** 4 + np.float64(c.Omega_g0) * a1 ** 4 + np.float64(c.Omega_b0) * a1 ** 3
          + np.float64(c.Omega_c0) * a1 ** 3 + np.float64(c.Omega_k0) * a1 ** 2 + np.float64(c.Omega_L0) * a1 ** (-fde))
    return np.sqrt(om)


def H(c: Cosmology, z):
    return np.float64(f32(c.h * f32(100))) * E(c, z)          # c.h * 100 is a Float32 product (:80)


def comoving_distance(c: Cosmology, z):
    """src/cosmo.jl:81-83: speed_of_light * quadgk(1/H, 0, z, rtol=1e-8)  [Mpc]."""
    if z == 0:
        return 0.0
    val, _ = integrate.quad(lambda t: 1.0 / H(c, t), 0.0, float(z), epsrel=1e-10, epsabs=0.0, limit=200)
    return speed_of_light_km_s * val


def tables(c: Cosmology):
    """The cache of src/cosmo.jl:86-91: z = range(z_tab_min, z_tab_max, length = z_tab_num), r = comoving_distance.(z).
    Built by cumulative quadrature over the knots (same integrals, O(n) instead of O(n) full integrals)."""
    if c.cache is None:
        z = np.linspace(np.float64(c.z_tab_min), np.float64(c.z_tab_max), int(c.z_tab_num))
        # 7-point Gauss-Legendre per knot interval (relative error ~1e-15 for intervals of 3e-5 .. 1e-3)
        xg, wg = np.polynomial.legendre.leggauss(7)
        a, b = z[:-1], z[1:]
        mid, half = 0.5 * (a + b), 0.5 * (b - a)
        nodes = mid[:, None] + half[:, None] * xg[None, :]
        seg = (wg[None, :] / H(c, nodes)).sum(axis=1) * half
        r = np.concatenate([[0.0], np.cumsum(seg)]) * speed_of_light_km_s
        if z[0] != 0:
            r += comoving_distance(c, z[0])
        c.cache = {"z": z, "r": r}
    return c.cache["z"], c.cache["r"]


class OutOfTableError(ValueError):
    """Interpolations.jl raises BoundsError outside the knots."""


def _gridded_linear(knots, values, x):
    """interpolate((knots,), values, Gridded(Linear()))(x): i = searchsortedlast(knots, x) clamped to [1, n-1],
    t = (x - k_i) / (k_{i+1} - k_i), (1 - t) v_i + t v_{i+1}; Float64 throughout."""
    x = np.asarray(x, np.float64)
    if x.size and (np.any(~(x >= knots[0])) or np.any(~(x <= knots[-1]))):
        raise OutOfTableError("argument outside the tabulated range")
    i = np.clip(np.searchsorted(knots, x, side="right") - 1, 0, len(knots) - 2)
    t = (x - knots[i]) / (knots[i + 1] - knots[i])
    return (1.0 - t) * values[i] + t * values[i + 1]


def comoving_distance_interp(c: Cosmology):
    z, r = tables(c)
    return lambda x: _gridded_linear(z, r, x)


def redshift_interp(c: Cosmology):
    z, r = tables(c)
    return lambda x: _gridded_linear(r, z, x)


def _trig32(fn, a):
    """Julia's cos/sin(::Float32): Float64 kernel, one rounding."""
    return fn(np.asarray(a, np.float32).astype(np.float64)).astype(np.float32)


def sky_to_cartesian(ra, dec, red, c: Cosmology):
    """examples/lightcone.jl:30-49.  ra, dec in degrees, red = redshift (Float32 arrays) -> (x, y, z) in Mpc/h."""
    T = np.asarray(ra).dtype.type
    r_fun = comoving_distance_interp(c)
    h = T(c.H0 / f32(100))                                  # h::eltype = cosmo.H0 / 100
    dist = r_fun(np.asarray(red, T).astype(np.float64))     # Float64
    ra_r = ((np.asarray(ra, T) * T(np.pi)).astype(T) / T(180)).astype(T)
    dec_r = ((np.asarray(dec, T) * T(np.pi)).astype(T) / T(180)).astype(T)
    if T is np.float32:
        cd, sd, cr, sr = _trig32(np.cos, dec_r), _trig32(np.sin, dec_r), _trig32(np.cos, ra_r), _trig32(np.sin, ra_r)
    else:
        cd, sd, cr, sr = np.cos(dec_r), np.sin(dec_r), np.cos(ra_r), np.sin(ra_r)
    h64 = np.float64(h)
    x = (dist * cd.astype(np.float64) * cr.astype(np.float64) * h64).astype(T)
    y = (dist * cd.astype(np.float64) * sr.astype(np.float64) * h64).astype(T)
    z = (dist * sd.astype(np.float64) * h64).astype(T)
    return x, y, z


def cartesian_to_sky(x, y, z, c: Cosmology):
    """examples/lightcone.jl:51-80 -> (ra, dec, redshift).  Quirk kept: `lon = (lon - 360) % 360` with Julia's
    truncated remainder leaves the right ascension in (-360, 0]."""
    T = np.asarray(x).dtype.type
    z_fun = redshift_interp(c)
    x, y, z = (np.asarray(a, T) for a in (x, y, z))
    r = (np.sqrt(((x * x + y * y).astype(T) + z * z).astype(T)).astype(T) / T(c.h)).astype(T)
    red = z_fun(r.astype(np.float64))
    s = np.sqrt((x * x + y * y).astype(T)).astype(T)
    lon = np.arctan2(y.astype(np.float64), x.astype(np.float64)).astype(T)
    lat = np.arctan2(z.astype(np.float64), s.astype(np.float64)).astype(T)
    lon = ((lon * T(180)).astype(T) / T(np.pi)).astype(T)
    lat = ((lat * T(180)).astype(T) / T(np.pi)).astype(T)
    lon = np.fmod((lon - T(360)).astype(T), T(360)).astype(T)
    return lon, lat, red.astype(T)


def fkp_weights(nz, P0):
    """examples/lightcone.jl:82: 1 / (1 + nz * P0) in the catalog's precision."""
    T = np.asarray(nz).dtype.type
    return (T(1) / (T(1) + (np.asarray(nz, T) * T(P0)).astype(T)).astype(T)).astype(T)


def wrap_positions(x, y, z, box_size, box_min=(0.0, 0.0, 0.0)):
    """test_helpers/simulation.py:38,51-52: pos = (pos + L) % L (floored modulo), generalised to a box that
    starts at box_min; evaluated in the catalog's precision."""
    out = []
    for p, L, mn in zip((x, y, z), box_size, box_min):
        T = np.asarray(p).dtype.type
        q = ((np.asarray(p, T) - T(mn)).astype(T) + T(L)).astype(T)
        q = np.mod(q, T(L)).astype(T)
        out.append((q + T(mn)).astype(T))
    return out
