// Per-particle arithmetic of the catalog pre/post-processing kernels (catalog.cu), written once as
// __host__ __device__ functions.  The product library only ever runs them on the device; the header
// also compiles as plain C++ so that tests/hostcheck/ can execute the very same statements on the CPU
// of the build container (no GPU there) against the CPU checker.  Host compilation must use
// -ffp-contract=off so that every Float32 operation rounds exactly like the *_rn intrinsics.
#pragma once
#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define BR_HD __host__ __device__ __forceinline__
#else
#define BR_HD inline
#endif

namespace baorec {
namespace catalog {

constexpr int kCoarse = 2048;          // entries of the shared-memory coarse table (16 KB)
constexpr float kPiF = 3.14159274f;    // Float32(pi)

#if defined(__CUDA_ARCH__)
BR_HD float fmul(float a, float b) { return __fmul_rn(a, b); }
BR_HD float fadd(float a, float b) { return __fadd_rn(a, b); }
BR_HD float fsub(float a, float b) { return __fsub_rn(a, b); }
BR_HD float fdiv(float a, float b) { return __fdiv_rn(a, b); }
BR_HD float fsqrt(float a) { return __fsqrt_rn(a); }
BR_HD double ld(const double* p) { return __ldg(p); }
BR_HD float quiet_nan() { return __int_as_float(0x7fc00000); }
#else
BR_HD float fmul(float a, float b) { return a * b; }
BR_HD float fadd(float a, float b) { return a + b; }
BR_HD float fsub(float a, float b) { return a - b; }
BR_HD float fdiv(float a, float b) { return a / b; }
BR_HD float fsqrt(float a) { return sqrtf(a); }
BR_HD double ld(const double* p) { return *p; }
BR_HD float quiet_nan() { return nanf(""); }
#endif

// Layout of the coarse table over an n-knot table: every `stride`-th knot, the last entry clamped to knot n-1.
BR_HD void coarse_layout(int64_t ntab, int64_t* stride, int* ncoarse) {
  const int64_t s = (ntab - 1 + (kCoarse - 2)) / (kCoarse - 1);   // ceil((n-1)/(C-1)) >= 1
  *stride = s;
  *ncoarse = (int)((ntab - 1 + s - 1) / s) + 1;                   // covers knot n-1; <= kCoarse
}

// interpolate((range(k0, step = dk, length = n),), v, Gridded(Linear()))(x)  (src/cosmo.jl:92):
// i = searchsortedlast clamped to [0, n-2], t = (x - k_i)/(k_{i+1} - k_i), (1 - t) v_i + t v_{i+1}
BR_HD bool interp_uniform(const double* v, double k0, double k1, double dk, int64_t n, double x, double* out) {
  if (!(x >= k0 && x <= k1)) return false;   // [k0, k1] = the table's own end points; also catches NaN
  int64_t i = (int64_t)floor((x - k0) / dk);
  i = i < 0 ? 0 : (i > n - 2 ? n - 2 : i);
  const double ka = fma((double)i, dk, k0), kb = fma((double)(i + 1), dk, k0);
  const double t = (x - ka) / (kb - ka);
  *out = (1.0 - t) * ld(v + i) + t * ld(v + i + 1);
  return true;
}

// interpolate((r,), range(z0, step = dz, length = n), Gridded(Linear()))(x)  (src/cosmo.jl:102): knots r[]
// increasing.  searchsortedlast is narrowed by the coarse table, then finished in the full table.
BR_HD bool interp_inverse(const double* r, const double* coarse, int ncoarse, int64_t stride, double z0, double dz, int64_t n,
                          double x, double* out) {
  if (!(x >= ld(r) && x <= ld(r + n - 1))) return false;
  int lo = 0, hi = ncoarse - 1;              // coarse[lo] <= x; x < coarse[hi] unless hi is the last entry
  while (hi - lo > 1) {
    const int m = (lo + hi) >> 1;
    if (coarse[m] <= x) lo = m; else hi = m;
  }
  int64_t a = (int64_t)lo * stride, b = a + stride;
  if (b > n - 1) b = n - 1;
  while (b - a > 1) {                        // r[a] <= x <= r[b]
    const int64_t m = (a + b) >> 1;
    if (ld(r + m) <= x) a = m; else b = m;
  }
  if (a > n - 2) a = n - 2;
  const double ra = ld(r + a), rb = ld(r + a + 1);
  const double t = (x - ra) / (rb - ra);
  *out = (1.0 - t) * fma((double)a, dz, z0) + t * fma((double)(a + 1), dz, z0);
  return true;
}

// examples/lightcone.jl:36-46.  Julia promotion: ra * pi / 180 in Float32; cos/sin(::Float32) through a
// Float64 kernel rounded once; dist (Float64) * cos * cos * h in Float64, rounded when stored.
BR_HD bool sky_to_cartesian_one(float ra, float dec, float red, float h, const double* rtab, double z0, double z1, double dz,
                                int64_t ntab, float* ox, float* oy, float* oz) {
  double dist;
  const bool ok = interp_uniform(rtab, z0, z1, dz, ntab, (double)red, &dist);
  const float ra_r = fdiv(fmul(ra, kPiF), 180.0f);
  const float dec_r = fdiv(fmul(dec, kPiF), 180.0f);
  double sd, cd, sr, cr;
  sincos((double)dec_r, &sd, &cd);
  sincos((double)ra_r, &sr, &cr);
  const double cdf = (double)(float)cd, sdf = (double)(float)sd, crf = (double)(float)cr, srf = (double)(float)sr;
  if (!ok) {
    *ox = *oy = *oz = quiet_nan();
    return false;
  }
  *ox = (float)(dist * cdf * crf * (double)h);
  *oy = (float)(dist * cdf * srf * (double)h);
  *oz = (float)(dist * sdf * (double)h);
  return true;
}

// examples/lightcone.jl:55-77; `(lon - 360) % 360` (truncated remainder) leaves ra in (-360, 0].
BR_HD bool cartesian_to_sky_one(float px, float py, float pz, float h, const double* rtab, const double* coarse, int ncoarse,
                                int64_t stride, double z0, double dz, int64_t ntab, float* ora, float* odec, float* ored) {
  const float xy2 = fadd(fmul(px, px), fmul(py, py));
  const float r = fdiv(fsqrt(fadd(xy2, fmul(pz, pz))), h);
  double red;
  const bool ok = interp_inverse(rtab, coarse, ncoarse, stride, z0, dz, ntab, (double)r, &red);
  const float s = fsqrt(xy2);
  float lon = (float)atan2((double)py, (double)px);
  float lat = (float)atan2((double)pz, (double)s);
  lon = fdiv(fmul(lon, 180.0f), kPiF);
  lat = fdiv(fmul(lat, 180.0f), kPiF);
  *ora = fmodf(fsub(lon, 360.0f), 360.0f);
  *odec = lat;
  *ored = ok ? (float)red : quiet_nan();
  return ok;
}

// examples/lightcone.jl:82
BR_HD float fkp_one(float nz, float P0) { return fdiv(1.0f, fadd(1.0f, fmul(nz, P0))); }

// test_helpers/simulation.py:38: pos = (pos + L) % L (floored modulo), for a box starting at mn
BR_HD float wrap_one(float p, float L, float mn) {
  const float q = fadd(fsub(p, mn), L);
  float m = fmodf(q, L);
  if (m != 0.0f) {
    if (m < 0.0f) m = fadd(m, L);
  } else {
    m = 0.0f;
  }
  return fadd(m, mn);
}

// ---- host only: the comoving-distance table (src/cosmo.jl:70-91) ---------------------------------
struct CosmoPars {
  double h, Omega_b0, Omega_c0, Omega_nu0, Omega_g0, Omega_k0, Omega_L0, w0, wa;
};

inline double E_of_z(const CosmoPars& c, double z) {   // src/cosmo.jl:70-79, Float64 like the quadrature nodes
  const double a1 = 1.0 + z;
  double fde = -3.0 * (1.0 + c.w0);
  if (c.wa != 0.0) {
    const double a = 1.0 / a1;
    const double la = log(a);
    fde += 3.0 * c.wa * ((la != 0.0 ? (a - 1.0) / la : 1.0) - 1.0);
  }
  const double a2 = a1 * a1, a3 = a2 * a1, a4 = a2 * a2;
  return sqrt(c.Omega_nu0 * a4 + c.Omega_g0 * a4 + c.Omega_b0 * a3 + c.Omega_c0 * a3 + c.Omega_k0 * a2 +
              c.Omega_L0 * pow(a1, -fde));
}

inline double integrate_inv_H(const CosmoPars& c, double H0, double a, double b) {   // 7-point Gauss-Legendre
  static const double gx[7] = {-0.9491079123427585, -0.7415311855993945, -0.4058451513773972, 0.0,
                               0.4058451513773972,  0.7415311855993945,  0.9491079123427585};
  static const double gw[7] = {0.1294849661688697, 0.2797053914892766, 0.3818300505051189, 0.4179591836734694,
                               0.3818300505051189, 0.2797053914892766, 0.1294849661688697};
  const double mid = 0.5 * (a + b), half = 0.5 * (b - a);
  double s = 0.0;
  for (int i = 0; i < 7; i++) s += gw[i] / (H0 * E_of_z(c, mid + half * gx[i]));
  return s * half;
}

// r[i] = c * int_0^{z_i} dz / H(z), z_i = z0 + i dz; returns the first index where r stops increasing, or -1.
inline int64_t build_distance_table(const CosmoPars& c, double z0, double dz, int64_t n, double* r) {
  const double H0 = (double)((float)c.h * 100.0f);   // c.h * 100 is a Float32 product (src/cosmo.jl:80)
  const double clight = 299792.458;                  // src/cosmo.jl:13
  double acc = 0.0;
  if (z0 > 0.0) {                                    // [0, z_tab_min] in 4096 panels
    const int panels = 4096;
    for (int i = 0; i < panels; i++) acc += integrate_inv_H(c, H0, z0 * i / panels, z0 * (i + 1) / panels);
  }
  r[0] = clight * acc;
  for (int64_t i = 1; i < n; i++) {
    acc += integrate_inv_H(c, H0, z0 + dz * (double)(i - 1), z0 + dz * (double)i);
    r[i] = clight * acc;
    if (!(r[i] > r[i - 1])) return i;
  }
  return -1;
}

}  // namespace catalog
}  // namespace baorec
