// TEST INFRASTRUCTURE: compiles the per-mode arithmetic of the multipole estimator (baorec.jl_b200/csrc/pk_ops.cuh
// -- what pk_kernel calls on the device) as plain C++ and walks a whole k-space half mesh on the CPU in the
// kernel's own index convention, so that tests/test_pk_hostcheck.py can hold it to oracle/pk_oracle.py without a
// GPU.  Built with g++ -O2 -ffp-contract=off into tests/_build/.  Never linked into, or loaded by, the product.
#include <stdint.h>

#include "../../baorec.jl_b200/csrc/pk_ops.cuh"

using namespace baorec;

extern "C" {

// acc: [5][nbins] Float64 sums (w, w k, w P, w P L2, w P L4), zero-filled by the caller.  Returns the number of
// modes that fell into a bin.
int64_t hc_pk(const float* rk, const float* rk2, double sa, double sb, const float* kx, const float* ky, const float* kz, int nx, int ny, int nz, const double* wx,
              const double* wy, const double* wz, const double* los, double kmin, double dk, int nbins, double* acc) {
  PkGeom g;
  g.wx = wx;
  g.wy = wy;
  g.wz = wz;
  for (int a = 0; a < 3; a++) g.los[a] = los[a];
  g.kmin = kmin;
  g.inv_dk = 1.0 / dk;
  g.nbins = nbins;
  g.xh = nx / 2 + 1;
  g.nyq_x = nx % 2 == 0 ? nx / 2 : -1;
  const float2* in = (const float2*)rk;
  const float2* in2 = (const float2*)rk2;  // randoms mesh, or NULL
  int64_t used = 0;
  for (int iz = 0; iz < nz; iz++)
    for (int iy = 0; iy < ny; iy++)
      for (int ix = 0; ix < g.xh; ix++) {
        double c[5];
        const size_t idx = ((size_t)iz * ny + iy) * g.xh + ix;
        const float2 v = in[idx], v2 = in2 ? in2[idx] : make_float2(0.f, 0.f);
        const double re = (double)v.x * sa - (double)v2.x * sb, im = (double)v.y * sa - (double)v2.y * sb;  // as in pk_kernel
        const int b = pk_mode(g, re, im, kx[ix], ky[iy], kz[iz], ix, iy, iz, c);
        if (b < 0) continue;
        used++;
        for (int q = 0; q < 5; q++) acc[(size_t)q * nbins + b] += c[q];
      }
  return used;
}
}
