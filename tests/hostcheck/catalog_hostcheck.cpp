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
    if (!sky_to_cartesian_one(ra[p], dec[p], red[p], h, rtab, z0, z1, dz, ntab, x + p, y + p, z + p)) bad++;
  return bad;
}

int64_t hc_cartesian_to_sky(const float* x, const float* y, const float* z, int64_t n, float h, const double* rtab, double z0,
                            double dz, int64_t ntab, float* ra, float* dec, float* red) {
  int64_t stride;
  int ncoarse;
  coarse_layout(ntab, &stride, &ncoarse);
  if (ncoarse > kCoarse) return -1;
  std::vector<double> coarse((size_t)ncoarse);          // what the kernel stages in shared memory
  for (int j = 0; j < ncoarse; j++) {
    int64_t k = (int64_t)j * stride;
    coarse[(size_t)j] = rtab[k > ntab - 1 ? ntab - 1 : k];
  }
  int64_t bad = 0;
  for (int64_t p = 0; p < n; p++)
    if (!cartesian_to_sky_one(x[p], y[p], z[p], h, rtab, coarse.data(), ncoarse, stride, z0, dz, ntab, ra + p, dec + p, red + p))
      bad++;
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

void hc_coarse_layout(int64_t ntab, int64_t* stride, int* ncoarse) { coarse_layout(ntab, stride, ncoarse); }

}  // extern "C"
