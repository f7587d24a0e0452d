// TEST INFRASTRUCTURE: compiles the product's cell / weight arithmetic of the mass-assignment kernels
// (baorec.jl_b200/csrc/mas_math.cuh: wrap_pos, cic_axis, gather_axis, tsc_axis -- the device functions every
// scatter / gather kernel of mas.cu calls) as plain C++ and runs it in loops on the CPU, so that the build
// container -- which has no GPU -- can hold the very statements the device executes to the oracle's
// bit-exact cell indices and weights (tests/test_mas_hostcheck.py).  The loops mirror cic_cells_kernel /
// gather_cells_kernel of mas.cu.  Built with g++ -O2 -ffp-contract=off (no FMA contraction == the *_rn
// intrinsics) into tests/_build/.  Never linked into, or loaded by, the product library.
#include <stdint.h>

#include "../../baorec.jl_b200/csrc/mas_math.cuh"

using namespace baorec;

extern "C" {

// i0/i1/w0/w1 are [3][n] like baorec_cic_cells_f32; out-of-box particles get index -1
void hc_cic_cells(const float* x, const float* y, const float* z, int64_t n, const int* ng, const float* L, const float* mn,
                  int wrap, int32_t* i0, int32_t* i1, float* w0, float* w1, float* wrapped) {
  const float* p[3] = {x, y, z};
  for (int64_t i = 0; i < n; i++)
    for (int a = 0; a < 3; a++) {
      float v = p[a][i];
      if (wrap) v = wrap_pos(v, mn[0], L[0]);
      if (wrapped) wrapped[a * n + i] = v;
      int c0 = -1, c1 = -1;
      float a0 = 0.f, a1 = 0.f;
      if (!cic_axis(v, mn[a], L[a], ng[a], wrap != 0, c0, c1, a0, a1)) c0 = c1 = -1;
      i0[a * n + i] = c0;
      i1[a * n + i] = c1;
      w0[a * n + i] = a0;
      w1[a * n + i] = a1;
    }
}

void hc_gather_cells(const float* x, const float* y, const float* z, int64_t n, const int* ng, const float* L, const float* mn,
                     const float* cell, int gpu_formula, int32_t* id, int32_t* iu, float* wd, float* wu) {
  const float* p[3] = {x, y, z};
  for (int64_t i = 0; i < n; i++)
    for (int a = 0; a < 3; a++) {
      int c0 = -1, c1 = -1;
      float a0 = 0.f, a1 = 0.f;
      if (!gather_axis(p[a][i], mn[a], L[a], cell[a], ng[a], gpu_formula != 0, c0, c1, a0, a1)) c0 = c1 = -1;
      id[a * n + i] = c0;
      iu[a * n + i] = c1;
      wd[a * n + i] = a0;
      wu[a * n + i] = a1;
    }
}

// idx / w are [3 axes][3 stencil points][n]
void hc_tsc_cells(const float* x, const float* y, const float* z, int64_t n, const int* ng, const float* L, const float* mn, int wrap,
                  int32_t* idx, float* w, int32_t* ok) {
  const float* p[3] = {x, y, z};
  for (int64_t i = 0; i < n; i++) {
    int good = 1;
    for (int a = 0; a < 3; a++) {
      int id3[3] = {-1, -1, -1};
      float w3[3] = {0.f, 0.f, 0.f};
      if (!tsc_axis(p[a][i], mn[a], L[a], ng[a], wrap != 0, id3, w3)) good = 0;
      for (int o = 0; o < 3; o++) {
        idx[(a * 3 + o) * n + i] = id3[o];
        w[(a * 3 + o) * n + i] = w3[o];
      }
    }
    ok[i] = good;
  }
}

}  // extern "C"
