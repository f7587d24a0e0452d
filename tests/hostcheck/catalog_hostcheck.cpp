// TEST INFRASTRUCTURE: compiles the product's per-particle catalog arithmetic
// (baorec.jl_b200/csrc/catalog_math.cuh, the __host__ __device__ functions catalog.cu's kernels call)
// as plain C++ and runs it in loops on the CPU, so that the build container -- which has no GPU --
// can check the very statements the device executes against oracle/catalog_oracle.py
// (tests/test_catalog_hostcheck.py).  Built by the test fixture with
//   g++ -O2 -ffp-contract=off -shared -fPIC   (no FMA contraction == the *_rn intrinsics)
// into tests/_build/ (git-ignored).  Never linked into, or loaded by, the product library.
#include <vector>

#include "../../baorec.jl_b200/csrc/catalog_math.cuh"

using namespace baorec::catalog;

extern "C" {

int64_t hc_sky_to_cartesian(const float* ra, const float* dec, const float* red, int64_t n, float h, const double* rtab,
                            double z0, double z1, double dz, int64_t ntab, float* x, float* y, float* z) {
  int64_t bad = 0;
  for (int64_t p = 0; p < n; p++)
    if (!sky_to_cartesian_one(ra[p], dec[p], red[p], h, rtab, z0, z1, dz, 1.0 / dz, ntab, x + p, y + p, z + p)) bad++;
  return bad;
}

// corr < 0: the number of correction steps the product would measure; otherwise forced (0 = every miss bisects)
int64_t hc_cartesian_to_sky(const float* x, const float* y, const float* z, int64_t n, float h, const double* rtab, double z0,
                            double dz, int64_t ntab, float* ra, float* dec, float* red, int vec4, int corr, int* corr_measured) {
  std::vector<double> g((size_t)kGuessBins + 1);          // what baorec_cosmo_set uploads next to the distance table
  double inv_h;
  const int measured = build_inverse_guess(rtab, ntab, g.data(), &inv_h);
  if (corr_measured) *corr_measured = measured;
  if (corr < 0) corr = measured;
  const double r0 = rtab[0], r1 = rtab[ntab - 1];
  int64_t bad = 0, p = 0;
  for (; p + 4 <= n && vec4; p += 4) {                  // the VEC = 4 instantiation the aligned kernel runs
    float px[4], py[4], pz[4], a[4], d[4], r[4];
    for (int i = 0; i < 4; i++) px[i] = x[p + i], py[i] = y[p + i], pz[i] = z[p + i];
    bad += cartesian_to_sky_vec<4>(px, py, pz, h, rtab, g.data(), r0, r1, inv_h, corr, z0, dz, (int)ntab, a, d, r);
    for (int i = 0; i < 4; i++) ra[p + i] = a[i], dec[p + i] = d[i], red[p + i] = r[i];
  }
  for (; p < n; p++) {                                  // VEC = 1: unaligned arrays and the tail
    float px[1] = {x[p]}, py[1] = {y[p]}, pz[1] = {z[p]}, a[1], d[1], r[1];
    bad += cartesian_to_sky_vec<1>(px, py, pz, h, rtab, g.data(), r0, r1, inv_h, corr, z0, dz, (int)ntab, a, d, r);
    ra[p] = a[0], dec[p] = d[0], red[p] = r[0];
  }
  return bad;
}

void hc_fkp_weights(const float* nz, int64_t n, float P0, float* w) {
  for (int64_t p = 0; p < n; p++) w[p] = fkp_one(nz[p], P0);
}

void hc_wrap_positions(float* x, float* y, float* z, int64_t n, const float* L, const float* mn) {
  for (int64_t p = 0; p < n; p++) {
    x[p] = wrap_one(x[p], L[0], mn[0]);
    y[p] = wrap_one(y[p], L[1], mn[1]);
    z[p] = wrap_one(z[p], L[2], mn[2]);
  }
}

void hc_sincos(const float* x, int64_t n, double* s, double* c) {
  for (int64_t p = 0; p < n; p++) sincos_reduced((double)x[p], s + p, c + p);
}

void hc_atan2(const double* y, const double* x, int64_t n, double* out) {
  for (int64_t p = 0; p < n; p++) out[p] = atan2_poly(y[p], x[p]);
}


}  // extern "C"
