// Per-mode arithmetic of the power-spectrum multipole estimator (pk.cu).  The reference validates a
// reconstruction through P_0 / P_2 / P_4 of the catalogs before and after (test_helpers/simulation.py:56-75, with
// the third-party pypowspec); this is the textbook FFT estimator for a periodic box, on the device, so that the
// check needs no second pass over the mesh on the host.
// Device code in the product; also compiles as plain C++ (host_shim.cuh) for tests/hostcheck/.
#pragma once
#include "host_shim.cuh"

namespace baorec {

struct PkGeom {
  const double* wx;  // per-axis inverse squared mass-assignment window 1 / sinc(k_a h_a / 2)^(2p), Float64, tabulated on the host
  const double* wy;
  const double* wz;
  double los[3];     // unit line of sight
  double kmin, inv_dk;  // bin = floor((|k| - kmin) * inv_dk), inv_dk = 1 / dk rounded once (the oracle does the same)
  int nbins;
  int xh;            // nx/2 + 1
  int nyq_x;         // index of the x-Nyquist plane (nx even), or -1
};

// Contributions of one mode of the half mesh.  Returns the bin (or -1: k = 0 / outside the bins) and
// c[0..4] = w, w k, w P, w P L_2(mu), w P L_4(mu)  with  P = |f_k|^2 / W(k)^2  for the field value (re, im) of the mode
// -- rho_k / rho_0, minus ran_k / ran_0 when a randoms mesh is given; the caller multiplies by V once per bin --
// and w the Hermitian weight (1 on the planes kx = 0 and kx = Nyquist, else 2).
// Everything in Float64 from the Float32 k tables of the context (src/utils.jl:3-10): the bin index is then
// identical to the oracle's, value for value.
__device__ __forceinline__ int pk_mode(const PkGeom& g, double re, double im, float kx, float ky, float kz, int ix, int iy,
                                       int iz, double c[5]) {
  const double x = (double)kx, y = (double)ky, z = (double)kz;
  const double k2 = x * x + y * y + z * z;
  if (!(k2 > 0.0)) return -1;
  const double k = sqrt(k2);
  const double fb = floor((k - g.kmin) * g.inv_dk);
  if (!(fb >= 0.0 && fb < (double)g.nbins)) return -1;
  const double mu = (x * g.los[0] + y * g.los[1] + z * g.los[2]) / k;
  const double mu2 = mu * mu;
  const double iw2 = __ldg(g.wx + ix) * __ldg(g.wy + iy) * __ldg(g.wz + iz);
  const double p = (re * re + im * im) * iw2;
  const double w = (ix == 0 || ix == g.nyq_x) ? 1.0 : 2.0;
  const double wp = w * p;
  c[0] = w;
  c[1] = w * k;
  c[2] = wp;
  c[3] = wp * (1.5 * mu2 - 0.5);
  c[4] = wp * ((35.0 * mu2 * mu2 - 30.0 * mu2 + 3.0) / 8.0);
  return (int)fb;
}

}  // namespace baorec
