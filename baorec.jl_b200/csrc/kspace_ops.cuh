// The per-mode k-space operators of kspace.cu (the Op functors kspace_kernel<Op> applies): Gaussian smoothing,
// overdensity normalisation, the RSD iteration terms, ALL fixed-LOS iterations folded into one recurrence
// (FusedLosOp) and the displacement components.  Device code in the product; the header
// also compiles as plain C++ (host_shim.cuh) so that tests/hostcheck/ can apply the very same functors to a
// k-space mesh on the CPU and hold them to the reference's sequence of iterate! calls.
#pragma once
#include "host_shim.cuh"

namespace baorec {

__device__ __forceinline__ float ksq(float kx, float ky, float kz) {
  // k² = kx^2 + ky^2 + kz^2, left to right in Float32 (src/utils.jl:51)
  return __fadd_rn(__fadd_rn(__fmul_rn(kx, kx), __fmul_rn(ky, ky)), __fmul_rn(kz, kz));
}

// exp(-0.5 R² k²) in Float64 like the reference (src/utils.jl:52, 81), as the product of three
// per-axis factors exp(-0.5 R² k_a²) tabulated in Float64 at plan time (gauss_tables): the Float64
// `exp` per mode made the pass compute-bound.  k_a² is the Float32 square the reference forms; only
// the two Float32 roundings of its k² sum are not reproduced (<= 1.2e-7 relative in the exponent).
struct GaussTab {
  const double* gx;
  const double* gy;
  const double* gz;
  __device__ __forceinline__ double at(int ix, int iy, int iz) const {
    return __ldg(gx + ix) * __ldg(gy + iy) * __ldg(gz + iz);
  }
};

struct GaussOp {
  static const char* name() { return "kspace_kernel<GaussOp>"; }  // smooth!: field_k *= exp(-0.5 R² k²), then /M
  float2* out;
  GaussTab gt;
  double invM;
  __device__ __forceinline__ void apply(size_t idx, float2 v, float kx, float ky, float kz, bool, int ix, int iy, int iz) const {
    double s = gt.at(ix, iy, iz) * invM;
    out[idx] = make_float2((float)((double)v.x * s), (float)((double)v.y * s));
  }
};

struct SetupBoxOp {
  static const char* name() { return "kspace_kernel<SetupBoxOp>"; }  // smooth + (rho/mean - 1)/bias in k-space: delta_k = rho_k g /(A0 bias), DC -> 0
  float2* out;
  GaussTab gt;
  float bias;
  const double* dc;  // Re A_k[0] = sum(rho)
  __device__ __forceinline__ void apply(size_t idx, float2 v, float kx, float ky, float kz, bool is_dc, int ix, int iy, int iz) const {
    double s = gt.at(ix, iy, iz) * __ldg(dc + 8);  // dc[8] = 1 / (A0 bias)
    if (is_dc) s = 0.0;
    out[idx] = make_float2((float)((double)v.x * s), (float)((double)v.y * s));
  }
};

struct IterLosOp {
  static const char* name() { return "kspace_kernel<IterLosOp>"; }  // fixed LOS: sum_a k_a² los_a delta_k / k² (src/iterative.jl:10-14, 53), then /M
  float2* out;
  float los[3];
  float invM;
  __device__ __forceinline__ void apply(size_t idx, float2 v, float kx, float ky, float kz, bool, int, int, int) const {
    float k2 = ksq(kx, ky, kz);
    float c = __fmul_rn(__fmul_rn(kx, kx), los[0]);
    c = __fadd_rn(c, __fmul_rn(__fmul_rn(ky, ky), los[1]));
    c = __fadd_rn(c, __fmul_rn(__fmul_rn(kz, kz), los[2]));
    float2 o = make_float2(0.f, 0.f);
    if (k2 > 0.f) {
      o.x = __fmul_rn(__fmul_rn(__fdiv_rn(v.x, k2), c), invM);
      o.y = __fmul_rn(__fmul_rn(__fdiv_rn(v.y, k2), c), invM);
    }
    out[idx] = o;
  }
};

struct IterPairOp {
  static const char* name() { return "kspace_kernel<IterPairOp>"; }  // radial: k_i k_j delta_k / k² (src/iterative.jl:27), then /M
  float2* out;
  int i, j;
  float invM;
  __device__ __forceinline__ void apply(size_t idx, float2 v, float kx, float ky, float kz, bool, int, int, int) const {
    float k2 = ksq(kx, ky, kz);
    float ki = i == 0 ? kx : (i == 1 ? ky : kz);
    float kj = j == 0 ? kx : (j == 1 ? ky : kz);
    float c = __fmul_rn(ki, kj);
    float2 o = make_float2(0.f, 0.f);
    if (k2 > 0.f) {
      o.x = __fmul_rn(__fmul_rn(c, __fdiv_rn(v.x, k2)), invM);
      o.y = __fmul_rn(__fmul_rn(c, __fdiv_rn(v.y, k2)), invM);
    }
    out[idx] = o;
  }
};

// All n_iter fixed-LOS RSD iterations in one pass.  For a constant line of sight iterate! is
// linear and diagonal in k-space:  d^(n+1)_k = ds_k - fac_n mu_k d^(n)_k  with
// mu_k = sum_a k_a^2 los_a / k^2, fac_1 = beta/(1+beta), fac_n = beta (src/iterative.jl:43-62),
// so the per-iteration C2R/R2C round trips of the reference collapse into a per-mode recurrence.
// MODE 0: input is rho_k (smoothing and (rho/mean-1)/bias applied here, src/recon.jl:53-55);
// MODE 1: input is the R2C of delta_s.
template <int MODE>
struct FusedLosOp {
  static const char* name() { return MODE == 0 ? "kspace_kernel<FusedLosOp<rho>>" : "kspace_kernel<FusedLosOp<delta>>"; }
  float2* out_c2r;   // delta_final_k / M  (input of the C2R)
  float2* out_keep;  // delta_final_k (unnormalised), or nullptr
  GaussTab gt;
  float bias;
  const double* dc;
  float los[3];
  float beta;
  int n_iter;
  float invM;
  __device__ __forceinline__ void apply(size_t idx, float2 v, float kx, float ky, float kz, bool is_dc, int ix, int iy, int iz) const {
    float k2 = ksq(kx, ky, kz);
    float dsx, dsy;
    if (MODE == 0) {
      // delta_k (unnormalised R2C convention) = rho_k g M / (A0 bias): mean(rho) = A0 / M
      double s = gt.at(ix, iy, iz) * __ldg(dc + 8);  // dc[8] = M / (A0 bias)
      if (is_dc) s = 0.0;
      dsx = (float)((double)v.x * s);
      dsy = (float)((double)v.y * s);
    } else {
      dsx = v.x;
      dsy = v.y;
    }
    float c = __fmul_rn(__fmul_rn(kx, kx), los[0]);
    c = __fadd_rn(c, __fmul_rn(__fmul_rn(ky, ky), los[1]));
    c = __fadd_rn(c, __fmul_rn(__fmul_rn(kz, kz), los[2]));
    float mu = k2 > 0.f ? __fdiv_rn(c, k2) : 0.f;
    float drx = dsx, dry = dsy;
    for (int it = 1; it <= n_iter; it++) {
      float fac = it == 1 ? __fdiv_rn(beta, __fadd_rn(1.0f, beta)) : beta;
      float fm = __fmul_rn(fac, mu);
      drx = __fsub_rn(dsx, __fmul_rn(fm, drx));
      dry = __fsub_rn(dsy, __fmul_rn(fm, dry));
    }
    if (MODE == 1 && is_dc) {  // the reference zeroes the k=0 mode of the Hessian term only
      drx = dsx;
      dry = dsy;
    }
    out_c2r[idx] = make_float2(__fmul_rn(drx, invM), __fmul_rn(dry, invM));
    if (out_keep) out_keep[idx] = make_float2(drx, dry);
  }
};

template <bool POTENTIAL>
struct DispOp {
  static const char* name() { return POTENTIAL ? "kspace_kernel<DispOp<potential>>" : "kspace_kernel<DispOp<density>>"; }  // Psi_a = i k_a delta_k / k² (src/iterative.jl:268) or i k_a phi_k (src/multigrid.jl:767), /M
  float2* o0;
  float2* o1;
  float2* o2;
  float invM;
  __device__ __forceinline__ void apply(size_t idx, float2 v, float kx, float ky, float kz, bool, int, int, int) const {
    float s;
    if (POTENTIAL) {
      s = invM;
    } else {
      float k2 = ksq(kx, ky, kz);
      s = k2 > 0.f ? __fdiv_rn(invM, k2) : 0.f;
    }
    float re = __fmul_rn(-v.y, s), im = __fmul_rn(v.x, s);  // i * v * s
    o0[idx] = make_float2(__fmul_rn(re, kx), __fmul_rn(im, kx));
    o1[idx] = make_float2(__fmul_rn(re, ky), __fmul_rn(im, ky));
    o2[idx] = make_float2(__fmul_rn(re, kz), __fmul_rn(im, kz));
  }
};

}  // namespace baorec
