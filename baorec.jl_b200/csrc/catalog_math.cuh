// Per-particle arithmetic of the catalog pre/post-processing kernels (catalog.cu), written once as
// __host__ __device__ functions.  The product library only ever runs them on the device; the header
// also compiles as plain C++ so that tests/hostcheck/ can execute the very same statements on the CPU
// of the build container (no GPU there) against the CPU checker.  Host compilation must use
// -ffp-contract=off so that every Float32 operation rounds exactly like the *_rn intrinsics.
#pragma once
#include <math.h>
#include <stdint.h>
#include <string.h>

#if defined(__CUDACC__)
#define BR_HD __host__ __device__ __forceinline__   // layout helpers the API functions also call on the host
#define BR_D __device__ __forceinline__              // per-particle arithmetic: device only in the product
#define BR_TABLE static __constant__ double           // polynomial coefficients: constant-bank operands of DFMA
#else
#define BR_HD inline
#define BR_D inline
#define BR_TABLE static const double
#endif

namespace baorec {
namespace catalog {

constexpr float kPiF = 3.14159274f;    // Float32(pi)

#if defined(__CUDA_ARCH__)
BR_D float fmul(float a, float b) { return __fmul_rn(a, b); }
BR_D float fadd(float a, float b) { return __fadd_rn(a, b); }
BR_D float fsub(float a, float b) { return __fsub_rn(a, b); }
BR_D float fdiv(float a, float b) { return __fdiv_rn(a, b); }
BR_D float fsqrt(float a) { return __fsqrt_rn(a); }
BR_D double ld(const double* p) { return __ldg(p); }
BR_D float quiet_nan() { return __int_as_float(0x7fc00000); }
#else
BR_D float fmul(float a, float b) { return a * b; }
BR_D float fadd(float a, float b) { return a + b; }
BR_D float fsub(float a, float b) { return a - b; }
BR_D float fdiv(float a, float b) { return a / b; }
BR_D float fsqrt(float a) { return sqrtf(a); }
BR_D double ld(const double* p) { return *p; }
BR_D float quiet_nan() { return nanf(""); }
#endif

// interpolate((range(k0, step = dk, length = n),), v, Gridded(Linear()))(x)  (src/cosmo.jl:92):
// i = searchsortedlast clamped to [0, n-2], t = (x - k_i)/(k_{i+1} - k_i), (1 - t) v_i + t v_{i+1}
// (inv_dk = 1/dk: the knots are uniform, so both divisions become multiplications; t moves by <= 2^-52)
BR_D bool interp_uniform(const double* v, double k0, double k1, double dk, double inv_dk, int64_t n, double x, double* out) {
  if (!(x >= k0 && x <= k1)) return false;   // [k0, k1] = the table's own end points; also catches NaN
  int64_t i = (int64_t)floor((x - k0) * inv_dk);
  i = i < 0 ? 0 : (i > n - 2 ? n - 2 : i);
  const double t = (x - fma((double)i, dk, k0)) * inv_dk;
  *out = (1.0 - t) * ld(v + i) + t * ld(v + i + 1);
  return true;
}

// Degree-17 polynomial in z = t^2 for atan(t)/t on [0, 1] (interpolated at Chebyshev nodes: |error| < 8e-16),
// highest power first.  A constant-bank table: the Horner steps read their coefficients as uniform operands.
BR_TABLE kAtanC[18] = {-0x1.7f90f9cfdb736p-15, 0x1.e38c3cf752829p-12,  -0x1.2025beb28c550p-9, 0x1.b354a4c8666c3p-8,
                       -0x1.d9c94754ffed9p-7,  0x1.92de6e09c3905p-6,   -0x1.1db1227853a4fp-5, 0x1.669453ee33ed7p-5,
                       -0x1.a4125e7343365p-5,  0x1.dededc10302e9p-5,   -0x1.10c1eefc7e7fdp-4, 0x1.3b07c37190c98p-4,
                       -0x1.745bd222cc3dcp-4,  0x1.c71c5aa9341dfp-4,   -0x1.249248a39fa5ap-3, 0x1.999999969e074p-3,
                       -0x1.5555555551d01p-2,  0x1.ffffffffffff5p-1};

// sin and cos of a Float32 angle in Float64, accurate to ~1e-16 so that the single rounding to Float32 that
// follows is (all but never) the correctly rounded value, like Julia's sin/cos(::Float32).  Cody-Waite
// reduction by pi/2 (two FMAs, exact for |x| < 1e5) and Taylor polynomials on [-pi/4, pi/4] (truncation
// < 1e-16): ~25 Float64 operations instead of the ~100 of the library sincos.  Larger |x|, NaN and Inf take
// the library path.  (Measured on B200, 1e8 particles: library sincos 1.18 ms, this 0.96 ms; a variant with
// constant-bank coefficient tables, shift-based rounding and an out-of-line library call measured 1.34 ms --
// the pointer outputs of the non-inlined call put s and c on the stack -- and was dropped.)
BR_D void sincos_reduced(double x, double* s, double* c) {
  if (!(fabs(x) < 1.0e5)) {
    sincos(x, s, c);
    return;
  }
  const double k = rint(x * 0.63661977236758134308);     // 2/pi
  double r = fma(-k, 1.57079632679489655800e+00, x);     // pi/2, high part
  r = fma(-k, 6.12323399573676603587e-17, r);            // pi/2, low part
  const double z = r * r;
  double ps = -1.0 / 1307674368000.0;                    // -1/15!
  ps = fma(ps, z, 1.0 / 6227020800.0);                   // 1/13!
  ps = fma(ps, z, -1.0 / 39916800.0);                    // -1/11!
  ps = fma(ps, z, 1.0 / 362880.0);                       // 1/9!
  ps = fma(ps, z, -1.0 / 5040.0);                        // -1/7!
  ps = fma(ps, z, 1.0 / 120.0);                          // 1/5!
  ps = fma(ps, z, -1.0 / 6.0);                           // -1/3!
  const double sr = fma(ps * z, r, r);
  double pc = 1.0 / 20922789888000.0;                    // 1/16!
  pc = fma(pc, z, -1.0 / 87178291200.0);                 // -1/14!
  pc = fma(pc, z, 1.0 / 479001600.0);                    // 1/12!
  pc = fma(pc, z, -1.0 / 3628800.0);                     // -1/10!
  pc = fma(pc, z, 1.0 / 40320.0);                        // 1/8!
  pc = fma(pc, z, -1.0 / 720.0);                         // -1/6!
  pc = fma(pc, z, 1.0 / 24.0);                           // 1/4!
  pc = fma(pc, z, -0.5);
  const double cr = fma(pc, z, 1.0);
  const int q = (int)((long long)k & 3);                 // quadrant (two's complement: also right for k < 0)
  *s = (q == 0) ? sr : (q == 1) ? cr : (q == 2) ? -sr : -cr;
  *c = (q == 0) ? cr : (q == 1) ? -sr : (q == 2) ? -cr : sr;
}

// atan2(y, x) in Float64 with |error| < 1e-15 (one division, 18 Horner steps) instead of the library's
// ~90 instructions; same conventions (signed zeros, infinities, NaN propagates).
BR_D double atan2_poly(double y, double x) {
  const double ax = fabs(x), ay = fabs(y);
  const double mx = fmax(ax, ay), mn = fmin(ax, ay);
  double t = mn / mx;                                    // in [0, 1]
  if (mn == mx) t = (mx == 0.0) ? 0.0 : 1.0;             // 0/0 and inf/inf
  if (x != x || y != y) t = x + y;                       // fmax / fmin drop NaNs: put it back
  const double z = t * t;
  double p = kAtanC[0];
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
  for (int i = 1; i < 18; i++) p = fma(p, z, kAtanC[i]);
  double r = p * t;
  if (ay > ax) r = (1.57079632679489655800e+00 - r) + 6.12323399573676603587e-17;
  if (signbit(x)) r = (3.14159265358979311600e+00 - r) + 1.22464679914735320717e-16;
  return copysign(r, y);
}

// interpolate((r,), range(z0, step = dz, length = n), Gridded(Linear()))(x)  (src/cosmo.jl:102): knots r[]
// increasing, for VEC arguments at once.  Instead of bisecting the 100 000-knot table (17 dependent
// steps; the kernel was issue-bound on them) the interval is GUESSED from a small auxiliary table:
// g[j] = the continuous knot coordinate u(r) = i + (r - r_i)/(r_{i+1} - r_i) at kGuessBins + 1 uniform
// distances.  u(r) is smooth (u' = H/(c dz)), so interpolating g linearly lands within `corr` knots of the
// truth -- corr is measured exactly on the host when the table is built (the error is piecewise linear with
// its extrema at the knots) and is 1 for every realistic table.  `corr` branch-free +-1 steps then settle on
// the bracketing interval; a lane that still does not bracket (never, by construction) bisects the whole
// table, so the result does not depend on the guess being good.
constexpr int kGuessBins = 4096;

template <int VEC>
BR_D void interp_inverse_vec(const double* r, const double* g, double r0, double r1, double inv_h, int corr, double z0, double dz,
                             int n, const double (&x)[VEC], double (&out)[VEC], bool (&ok)[VEC]) {
  double xc[VEC], ra[VEC], rb[VEC];
  int k[VEC];
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
  for (int i = 0; i < VEC; i++) {
    ok[i] = (x[i] >= r0 && x[i] <= r1);      // also catches NaN
    xc[i] = ok[i] ? x[i] : r0;
    const double s = (xc[i] - r0) * inv_h;
    int j = (int)s;
    j = j < 0 ? 0 : (j > kGuessBins - 1 ? kGuessBins - 1 : j);
    const double g0 = ld(g + j), g1 = ld(g + j + 1);
    const double u = fma(s - (double)j, g1 - g0, g0);
    int kk = (int)u;
    kk = kk < 0 ? 0 : (kk > n - 2 ? n - 2 : kk);
    k[i] = kk;
    ra[i] = ld(r + kk);
    rb[i] = ld(r + kk + 1);
  }
  for (int c = 0; c < corr; c++) {
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int i = 0; i < VEC; i++) {
      int kk = k[i] + (xc[i] > rb[i] ? 1 : 0) - (xc[i] < ra[i] ? 1 : 0);
      kk = kk < 0 ? 0 : (kk > n - 2 ? n - 2 : kk);
      k[i] = kk;
      ra[i] = ld(r + kk);
      rb[i] = ld(r + kk + 1);
    }
  }
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
  for (int i = 0; i < VEC; i++) {
    if (!(xc[i] >= ra[i] && xc[i] <= rb[i])) {          // cold: the guess was off by more than corr knots
      int a = 0, b = n - 1;
      while (b - a > 1) {
        const int m = (a + b) >> 1;
        if (ld(r + m) <= xc[i]) a = m; else b = m;
      }
      k[i] = a;
      ra[i] = ld(r + a);
      rb[i] = ld(r + a + 1);
    }
    const double t = (xc[i] - ra[i]) / (rb[i] - ra[i]);
    out[i] = (1.0 - t) * fma((double)k[i], dz, z0) + t * fma((double)(k[i] + 1), dz, z0);
  }
}

// examples/lightcone.jl:36-46.  Julia promotion: ra * pi / 180 in Float32; cos/sin(::Float32) through a
// Float64 kernel rounded once; dist (Float64) * cos * cos * h in Float64, rounded when stored.
BR_D bool sky_to_cartesian_one(float ra, float dec, float red, float h, const double* rtab, double z0, double z1, double dz,
                                double inv_dz, int64_t ntab, float* ox, float* oy, float* oz) {
  double dist;
  const bool ok = interp_uniform(rtab, z0, z1, dz, inv_dz, ntab, (double)red, &dist);
  const float ra_r = fdiv(fmul(ra, kPiF), 180.0f);
  const float dec_r = fdiv(fmul(dec, kPiF), 180.0f);
  double sd, cd, sr, cr;
  sincos_reduced((double)dec_r, &sd, &cd);
  sincos_reduced((double)ra_r, &sr, &cr);
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

// examples/lightcone.jl:55-77, VEC particles at once; `(lon - 360) % 360` (truncated remainder) leaves ra in (-360, 0].
// Returns the number of particles whose distance lies outside the table (their redshift is NaN).
template <int VEC>
BR_D int cartesian_to_sky_vec(const float (&px)[VEC], const float (&py)[VEC], const float (&pz)[VEC], float h, const double* rtab,
                              const double* gtab, double r0, double r1, double inv_h, int corr, double z0, double dz, int ntab,
                              float (&ora)[VEC], float (&odec)[VEC], float (&ored)[VEC]) {
  double rr[VEC], red[VEC];
  bool ok[VEC];
  float xy2[VEC];
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
  for (int i = 0; i < VEC; i++) {
    xy2[i] = fadd(fmul(px[i], px[i]), fmul(py[i], py[i]));
    rr[i] = (double)fdiv(fsqrt(fadd(xy2[i], fmul(pz[i], pz[i]))), h);
  }
  interp_inverse_vec<VEC>(rtab, gtab, r0, r1, inv_h, corr, z0, dz, ntab, rr, red, ok);
  int bad = 0;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
  for (int i = 0; i < VEC; i++) {
    const float s = fsqrt(xy2[i]);
    float lon = (float)atan2_poly((double)py[i], (double)px[i]);
    float lat = (float)atan2_poly((double)pz[i], (double)s);
    lon = fdiv(fmul(lon, 180.0f), kPiF);
    lat = fdiv(fmul(lat, 180.0f), kPiF);
    ora[i] = fmodf(fsub(lon, 360.0f), 360.0f);
    odec[i] = lat;
    ored[i] = ok[i] ? (float)red[i] : quiet_nan();
    bad += ok[i] ? 0 : 1;
  }
  return bad;
}

// examples/lightcone.jl:82
BR_D float fkp_one(float nz, float P0) { return fdiv(1.0f, fadd(1.0f, fmul(nz, P0))); }

// test_helpers/simulation.py:38: pos = (pos + L) % L (floored modulo), for a box starting at mn
BR_D float wrap_one(float p, float L, float mn) {
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

// The guess table of interp_inverse_vec for knots r[0..n): g[j] = u(r0 + j h), j = 0..kGuessBins, h = (r1 - r0)/kGuessBins.
// Returns the number of +-1 correction steps the device must take: the largest distance, over all knots,
// between the interpolated guess and the knot's own index, rounded up (the exact maximum over all r: guess
// and truth are both piecewise linear, equal at the bin edges, so the error peaks at a knot).
inline int build_inverse_guess(const double* r, int64_t n, double* g, double* inv_h_out) {
  const double r0 = r[0], r1 = r[n - 1];
  const double h = (r1 - r0) / kGuessBins, inv_h = 1.0 / h;
  int64_t i = 0;
  for (int j = 0; j <= kGuessBins; j++) {
    const double x = (j == kGuessBins) ? r1 : r0 + h * j;
    while (i < n - 2 && r[i + 1] <= x) i++;
    g[j] = (double)i + (x - r[i]) / (r[i + 1] - r[i]);
  }
  double worst = 0.0;
  for (int64_t k = 0; k < n; k++) {
    const double s = (r[k] - r0) * inv_h;
    int j = (int)s;
    j = j < 0 ? 0 : (j > kGuessBins - 1 ? kGuessBins - 1 : j);
    const double u = fma(s - (double)j, g[j + 1] - g[j], g[j]);
    const double e = fabs(u - (double)k);
    if (e > worst) worst = e;
  }
  *inv_h_out = inv_h;
  const double steps = ceil(worst);
  return steps < 1.0 ? 1 : (steps > 64.0 ? 64 : (int)steps);   // beyond that the (always correct) bisection path takes over
}

}  // namespace catalog
}  // namespace baorec
