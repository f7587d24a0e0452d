// TEST INFRASTRUCTURE: compiles the product's per-mode k-space operators (baorec.jl_b200/csrc/kspace_ops.cuh --
// the functors kspace_kernel<Op> applies on the device) as plain C++ and applies them to a whole k-space mesh on
// the CPU, with exactly the call convention of kspace_kernel in kspace.cu, so that tests/test_kspace_hostcheck.py
// can hold them to the oracle (and FusedLosOp to the reference's SEQUENCE of iterate! calls) without a GPU.
// Built with g++ -O2 -ffp-contract=off into tests/_build/.  Never linked into, or loaded by, the product library.
#include <stdint.h>

#include "../../baorec.jl_b200/csrc/kspace_ops.cuh"

using namespace baorec;

namespace {

struct Mesh {
  const float* kx;
  const float* ky;
  const float* kz;
  int xh, ny, nz;
};

template <class Op>
void apply_all(const Mesh& g, const float2* in, const Op& op) {
  for (int iz = 0; iz < g.nz; iz++)
    for (int iy = 0; iy < g.ny; iy++)
      for (int ix = 0; ix < g.xh; ix++) {
        const size_t idx = ((size_t)iz * g.ny + iy) * g.xh + ix;
        op.apply(idx, in[idx], g.kx[ix], g.ky[iy], g.kz[iz], (ix | iy | iz) == 0, ix, iy, iz);
      }
}

}  // namespace

extern "C" {

// tables: gx[xh], gy[ny], gz[nz] = exp(-0.5 R^2 k_a^2) in Float64 (ctx.cu: gauss_tables); dc[8] as the drivers set it
void hc_gauss(const float* in, float* out, const float* kx, const float* ky, const float* kz, int xh, int ny, int nz,
              const double* gx, const double* gy, const double* gz, double invM) {
  GaussOp op{(float2*)out, GaussTab{gx, gy, gz}, invM};
  apply_all(Mesh{kx, ky, kz, xh, ny, nz}, (const float2*)in, op);
}

void hc_setup_box(const float* in, float* out, const float* kx, const float* ky, const float* kz, int xh, int ny, int nz,
                  const double* gx, const double* gy, const double* gz, float bias, const double* dc) {
  SetupBoxOp op{(float2*)out, GaussTab{gx, gy, gz}, bias, dc};
  apply_all(Mesh{kx, ky, kz, xh, ny, nz}, (const float2*)in, op);
}

void hc_iter_los(const float* in, float* out, const float* kx, const float* ky, const float* kz, int xh, int ny, int nz,
                 const float* los, float invM) {
  IterLosOp op{(float2*)out, {los[0], los[1], los[2]}, invM};
  apply_all(Mesh{kx, ky, kz, xh, ny, nz}, (const float2*)in, op);
}

void hc_iter_pair(const float* in, float* out, const float* kx, const float* ky, const float* kz, int xh, int ny, int nz, int i,
                  int j, float invM) {
  IterPairOp op{(float2*)out, i, j, invM};
  apply_all(Mesh{kx, ky, kz, xh, ny, nz}, (const float2*)in, op);
}

void hc_fused_los(int mode, const float* in, float* out_c2r, float* out_keep, const float* kx, const float* ky, const float* kz,
                  int xh, int ny, int nz, const double* gx, const double* gy, const double* gz, float bias, const double* dc,
                  const float* los, float beta, int n_iter, float invM) {
  const Mesh g{kx, ky, kz, xh, ny, nz};
  if (mode == 0) {
    FusedLosOp<0> op{(float2*)out_c2r, (float2*)out_keep, GaussTab{gx, gy, gz}, bias, dc, {los[0], los[1], los[2]}, beta, n_iter, invM};
    apply_all(g, (const float2*)in, op);
  } else {
    FusedLosOp<1> op{(float2*)out_c2r, (float2*)out_keep, GaussTab{gx, gy, gz}, bias, dc, {los[0], los[1], los[2]}, beta, n_iter, invM};
    apply_all(g, (const float2*)in, op);
  }
}

void hc_disp(int potential, const float* in, float* o0, float* o1, float* o2, const float* kx, const float* ky, const float* kz,
             int xh, int ny, int nz, float invM) {
  const Mesh g{kx, ky, kz, xh, ny, nz};
  if (potential) {
    DispOp<true> op{(float2*)o0, (float2*)o1, (float2*)o2, invM};
    apply_all(g, (const float2*)in, op);
  } else {
    DispOp<false> op{(float2*)o0, (float2*)o1, (float2*)o2, invM};
    apply_all(g, (const float2*)in, op);
  }
}

}  // extern "C"
