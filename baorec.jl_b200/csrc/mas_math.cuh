// Cell / weight arithmetic of the mass-assignment kernels (mas.cu): cic! (src/mas.jl:7-35), read_cic!
// (src/mas.jl:224-255 CPU formula, :274-306 GPU formula) and the TSC extension, as the device functions every
// scatter / gather kernel calls.  Coordinate arithmetic uses the explicit round-to-nearest intrinsics so that
// nvcc never contracts it into FMAs: indices and weights are bit-identical to the reference's Float32 CPU
// arithmetic.  The product only runs these on the device; the header also compiles as plain C++
// (-ffp-contract=off, the intrinsics mapped to the plain operators by host_shim.cuh) so that tests/hostcheck/ can run the
// very same statements on the CPU of the build container against the CPU checker.
#pragma once
#include <stddef.h>
#include <stdint.h>

#include "baorec_b200.h"
#include "host_shim.cuh"

namespace baorec {

struct BoxGeom {
  float mn[3];
  float L[3];
  float cell[3];
  int n[3];
  // slab decomposition (multi-GPU): the mesh buffer holds nzp planes; global plane z maps to local
  // plane (z - z_lo + zoff) (periodic); slab == 0 -> the whole mesh, planes wrap at n[2]
  int slab, z_lo, zoff, nzp;
};

// ---- CIC scatter cell (src/mas.jl:13-35) -------------------------------------------------
__device__ __forceinline__ float wrap_pos(float p, float mn0, float L0) {
  // src/mas.jl:8-10 -- note: axis-1 box_min/box_size for every axis (reference quirk, kept).
  return (__fsub_rn(p, mn0) > L0) ? __fsub_rn(p, L0) : p;
}

__device__ __forceinline__ bool cic_axis(float p, float mn, float L, int n, bool wrap, int& i0, int& i1,
                                         float& w0, float& w1) {
  float g = __fadd_rn(__fdiv_rn(__fmul_rn(__fsub_rn(p, mn), (float)n), L), 1.0f);
  if (!(g >= 1.0f && g < (float)(n + 2))) return false;
  float f0 = floorf(g);
  int c0 = (int)f0;  // 1-based, as in the reference
  w1 = __fsub_rn(g, f0);
  w0 = __fsub_rn(1.0f, w1);
  if (c0 == n + 1) c0 = 1;
  int c1;
  if (c0 == n) {
    if (!wrap) return false;  // reference: index n+1 -> BoundsError / OOB atomic
    c1 = 1;
  } else {
    c1 = c0 + 1;
  }
  i0 = c0 - 1;
  i1 = c1 - 1;
  return true;
}

// ---- CIC gather cell (src/mas.jl:224-255 CPU formula; :274-306 GPU formula) ----------------
__device__ __forceinline__ bool gather_axis(float p, float mn, float L, float cell, int n, bool gpu_formula,
                                            int& id, int& iu, float& wd, float& wu) {
  float d = gpu_formula ? __fdiv_rn(__fmul_rn(__fsub_rn(p, mn), (float)n), L) : __fdiv_rn(__fsub_rn(p, mn), cell);
  if (!(d >= 0.0f && d < (float)(2 * n))) return false;
  float f = floorf(d);
  wu = __fsub_rn(d, f);
  wd = __fsub_rn(1.0f, wu);
  int i = (int)f + 1;  // 1-based
  if (i > n) i -= n;
  int j = i + 1;
  if (j > n) j -= n;
  id = i - 1;
  iu = j - 1;
  return true;
}

// ---- TSC (extension; same grid convention as cic!: mesh points at min + i*cell) -------------
__device__ __forceinline__ bool tsc_axis(float p, float mn, float L, int n, bool wrap, int idx[3], float w[3]) {
  float g = __fdiv_rn(__fmul_rn(__fsub_rn(p, mn), (float)n), L);
  if (!(g >= -1.0f && g <= (float)(n + 1))) return false;
  float c = floorf(__fadd_rn(g, 0.5f));
  float d = __fsub_rn(g, c);
  float hm = __fsub_rn(0.5f, d), hp = __fadd_rn(0.5f, d);
  w[0] = __fmul_rn(0.5f, __fmul_rn(hm, hm));
  w[1] = __fsub_rn(0.75f, __fmul_rn(d, d));
  w[2] = __fmul_rn(0.5f, __fmul_rn(hp, hp));
  int ic = (int)c;
#pragma unroll
  for (int o = 0; o < 3; o++) {
    int i = ic + o - 1;
    if (wrap) {
      i = i < 0 ? i + n : (i >= n ? i - n : i);
      if (i < 0 || i >= n) return false;
    } else if (i < 0 || i >= n) {
      return false;
    }
    idx[o] = i;
  }
  return true;
}

// ---- PCS (extension, SURVEY.md 8f N3: the piecewise cubic spline of the power-spectrum tool the reference's helpers
// configure, test_helpers/powspec_auto.conf:117-124).  The coordinate is cic!'s: g = (p - min) n / L + 1, base point
// c = floor(g) - 1, t = g - floor(g) (cic!'s upper weight), u = 1 - t; the four points c-1 .. c+2 carry the cubic
// B-spline u^3/6, ((3t-6)t^2+4)/6, ((3u-6)u^2+4)/6, t^3/6 -- every operation rounded to Float32 in the order written,
// as in the CPU restatement the parity tests compare with.  A particle owned by a slab (owner = slab of cic!'s base
// plane, the same cell) reaches the same planes as TSC does: one below the slab to two above it.
__device__ __forceinline__ bool pcs_axis(float p, float mn, float L, int n, bool wrap, int idx[4], float w[4]) {
  float g = __fadd_rn(__fdiv_rn(__fmul_rn(__fsub_rn(p, mn), (float)n), L), 1.0f);  // cic!'s 1-based coordinate
  if (!(g >= 0.0f && g <= (float)(n + 2))) return false;
  float c = floorf(g);
  float t = __fsub_rn(g, c), u = __fsub_rn(1.0f, t);
  c = c - 1.0f;  // 0-based base point: the cell cic! deposits into, last bit included
  float t2 = __fmul_rn(t, t), u2 = __fmul_rn(u, u);
  w[0] = __fdiv_rn(__fmul_rn(u2, u), 6.0f);
  w[1] = __fdiv_rn(__fadd_rn(__fmul_rn(__fsub_rn(__fmul_rn(3.0f, t), 6.0f), t2), 4.0f), 6.0f);
  w[2] = __fdiv_rn(__fadd_rn(__fmul_rn(__fsub_rn(__fmul_rn(3.0f, u), 6.0f), u2), 4.0f), 6.0f);
  w[3] = __fdiv_rn(__fmul_rn(t2, t), 6.0f);
  int ic = (int)c;
#pragma unroll
  for (int o = 0; o < 4; o++) {
    int i = ic + o - 1;
    if (wrap) {
      i = i < 0 ? i + n : (i >= n ? i - n : i);
      if (i < 0 || i >= n) return false;
    } else if (i < 0 || i >= n) {
      return false;
    }
    idx[o] = i;
  }
  return true;
}

// The stencil schemes behind one interface: SW points per axis (arrays are sized for the widest), point 1 is the one
// the sorts key on (TSC: the nearest grid point; PCS: the base point floor(g)).
template <int MAS>
struct MasStencil {
  static constexpr int SW = MAS == BAOREC_MAS_PCS ? 4 : 3;
};
template <int MAS>
__device__ __forceinline__ bool stencil_axis(float p, float mn, float L, int n, bool wrap, int idx[4], float w[4]) {
  if (MAS == BAOREC_MAS_PCS) return pcs_axis(p, mn, L, n, wrap, idx, w);
  return tsc_axis(p, mn, L, n, wrap, idx, w);
}

// Global plane indices (lower, upper-wrapped) -> plane indices in the local buffer.
__device__ __forceinline__ bool local_planes(const BoxGeom& g, int z0, int z1, int& l0, int& l1) {
  if (!g.slab) {
    l0 = z0;
    l1 = z1;
    return true;
  }
  const int nz = g.n[2];
  int dz = z0 - g.z_lo;          // in (-nz, nz)
  if (dz < 0) dz += nz;          // periodic distance above the slab base, in [0, nz)
  if (dz + g.zoff > g.nzp - 2) dz -= nz;  // not reachable from below: it is a plane under the slab
  l0 = dz + g.zoff;
  l1 = l0 + 1;
  return l0 >= 0 && l1 < g.nzp;
}

// One global plane index (already wrapped into [0, nz)) -> its plane in the local buffer.  The TSC stencil of a
// particle owned by this slab (owner = slab of its cic! base plane floor(g), the same sharding as CIC) reaches
// from one plane below the slab to two above it: both TSC slab layouts carry (nz_loc + 3) planes with the real
// planes at 1 .. nz_loc.
__device__ __forceinline__ bool local_plane1(const BoxGeom& g, int z, int& l) {
  if (!g.slab) {
    l = z;
    return true;
  }
  const int nz = g.n[2];
  int dz = z - g.z_lo;
  if (dz < 0) dz += nz;                   // periodic distance above the slab base, in [0, nz)
  if (dz + g.zoff > g.nzp - 1) dz -= nz;  // not reachable from below: it is a plane under the slab
  l = dz + g.zoff;
  return l >= 0 && l < g.nzp;
}

// ---- per-particle bodies ---------------------------------------------------------------------
// Deposit one (already wrapped) particle.  Returns false if it is outside the mesh.
template <int MAS>
__device__ __forceinline__ bool deposit(float* __restrict__ rho, float px, float py, float pz, float ww,
                                        const BoxGeom& g, bool wrap) {
  const size_t nx = g.n[0], ny = g.n[1];
  if (MAS == BAOREC_MAS_CIC) {
    int x0, x1, y0, y1, z0, z1;
    float wx0, wx1, wy0, wy1, wz0, wz1;
    bool ok = cic_axis(px, g.mn[0], g.L[0], g.n[0], wrap, x0, x1, wx0, wx1);
    ok = cic_axis(py, g.mn[1], g.L[1], g.n[1], wrap, y0, y1, wy0, wy1) && ok;
    ok = cic_axis(pz, g.mn[2], g.L[2], g.n[2], wrap, z0, z1, wz0, wz1) && ok;
    if (!ok || !local_planes(g, z0, z1, z0, z1)) return false;
    wx0 = __fmul_rn(wx0, ww);
    wx1 = __fmul_rn(wx1, ww);
    size_t r00 = ((size_t)z0 * ny + y0) * nx, r10 = ((size_t)z0 * ny + y1) * nx;
    size_t r01 = ((size_t)z1 * ny + y0) * nx, r11 = ((size_t)z1 * ny + y1) * nx;
    float a00 = __fmul_rn(wx0, wy0), a10 = __fmul_rn(wx1, wy0), a01 = __fmul_rn(wx0, wy1),
          a11 = __fmul_rn(wx1, wy1);
    atomicAdd(rho + r00 + x0, __fmul_rn(a00, wz0));
    atomicAdd(rho + r00 + x1, __fmul_rn(a10, wz0));
    atomicAdd(rho + r10 + x0, __fmul_rn(a01, wz0));
    atomicAdd(rho + r01 + x0, __fmul_rn(a00, wz1));
    atomicAdd(rho + r10 + x1, __fmul_rn(a11, wz0));
    atomicAdd(rho + r01 + x1, __fmul_rn(a10, wz1));
    atomicAdd(rho + r11 + x0, __fmul_rn(a01, wz1));
    atomicAdd(rho + r11 + x1, __fmul_rn(a11, wz1));
    return true;
  } else {
    constexpr int SW = MasStencil<MAS>::SW;
    int ix[4], iy[4], iz[4];
    float wx[4], wy[4], wz[4];
    bool ok = stencil_axis<MAS>(px, g.mn[0], g.L[0], g.n[0], wrap, ix, wx);
    ok = stencil_axis<MAS>(py, g.mn[1], g.L[1], g.n[1], wrap, iy, wy) && ok;
    ok = stencil_axis<MAS>(pz, g.mn[2], g.L[2], g.n[2], wrap, iz, wz) && ok;
    if (!ok) return false;
    if (g.slab) {  // slab layout (multi-GPU): the stencil planes in the local (nz_loc + 3)-plane buffer
#pragma unroll
      for (int c = 0; c < SW; c++) ok = local_plane1(g, iz[c], iz[c]) && ok;
      if (!ok) return false;
    }
#pragma unroll
    for (int c = 0; c < SW; c++) {
#pragma unroll
      for (int b = 0; b < SW; b++) {
        size_t row = ((size_t)iz[c] * ny + iy[b]) * nx;
#pragma unroll
        for (int a = 0; a < SW; a++) {
          float v = __fmul_rn(__fmul_rn(__fmul_rn(wx[a], ww), wy[b]), wz[c]);
          atomicAdd(rho + row + ix[a], v);
        }
      }
    }
    return true;
  }
}

// ---- paired deposit (option "scatter_pairs") ---------------------------------------------------------------------
// The eight cells of cic! are four rows of two x-neighbours.  When x0 is even (and the mesh row length is, and the
// pair does not straddle the periodic wrap) the two cells of a row are one 8-byte aligned pair, and sm_90+ has a
// vector reduction for exactly that: red.global.add.v2.f32 -- one L2 reduction instead of two.  Half of the particles
// qualify, so a catalog issues 6 reductions per particle on average instead of 8, in a kernel that is bound by L2
// reduction throughput.  Same cells, same Float32 values; only the grouping differs.
__device__ __forceinline__ void red_add_v2(float* p, float a, float b) {
#if defined(__CUDA_ARCH__)
  asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" ::"l"(__cvta_generic_to_global(p)), "f"(a), "f"(b) : "memory");
#else
  p[0] += a;
  p[1] += b;
#endif
}

__device__ __forceinline__ void red_add_v4(float* p, float a, float b, float c, float d) {
#if defined(__CUDA_ARCH__)
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(__cvta_generic_to_global(p)), "f"(a), "f"(b), "f"(c), "f"(d)
               : "memory");
#else
  p[0] += a;
  p[1] += b;
  p[2] += c;
  p[3] += d;
#endif
}

// cic! for one (already wrapped) particle, pairs where possible.  `rho` must be 8-byte aligned (MODE 1) / 16-byte
// aligned (MODE 2).  MODE 2 is branch-free: every row is ONE red.global.add.v4.f32 on the aligned 16-byte block that
// holds x0, the two values at the pair's offset in the block and +0 in the other slots (adding +0 leaves a cell as it
// is), plus a scalar reduction for the quarter of the particles whose pair straddles two blocks -- 5 reductions per
// particle on average instead of 6 (MODE 1, where the warp also diverges) or 8.
template <int MODE>
__device__ __forceinline__ bool deposit_pairs(float* __restrict__ rho, float px, float py, float pz, float ww,
                                              const BoxGeom& g, bool wrap) {
  const size_t nx = g.n[0], ny = g.n[1];
  int x0, x1, y0, y1, z0, z1;
  float wx0, wx1, wy0, wy1, wz0, wz1;
  bool ok = cic_axis(px, g.mn[0], g.L[0], g.n[0], wrap, x0, x1, wx0, wx1);
  ok = cic_axis(py, g.mn[1], g.L[1], g.n[1], wrap, y0, y1, wy0, wy1) && ok;
  ok = cic_axis(pz, g.mn[2], g.L[2], g.n[2], wrap, z0, z1, wz0, wz1) && ok;
  if (!ok || !local_planes(g, z0, z1, z0, z1)) return false;
  wx0 = __fmul_rn(wx0, ww);
  wx1 = __fmul_rn(wx1, ww);
  const size_t r00 = ((size_t)z0 * ny + y0) * nx, r10 = ((size_t)z0 * ny + y1) * nx;
  const size_t r01 = ((size_t)z1 * ny + y0) * nx, r11 = ((size_t)z1 * ny + y1) * nx;
  const float a00 = __fmul_rn(wx0, wy0), a10 = __fmul_rn(wx1, wy0), a01 = __fmul_rn(wx0, wy1), a11 = __fmul_rn(wx1, wy1);
  const float v000 = __fmul_rn(a00, wz0), v100 = __fmul_rn(a10, wz0), v010 = __fmul_rn(a01, wz0), v110 = __fmul_rn(a11, wz0);
  const float v001 = __fmul_rn(a00, wz1), v101 = __fmul_rn(a10, wz1), v011 = __fmul_rn(a01, wz1), v111 = __fmul_rn(a11, wz1);
  if (MODE == 2 && (g.n[0] & 3) == 0) {
    // every lane: ONE quad on the aligned block that holds x0 (the pair sits at offset o of it; +0 elsewhere), and for
    // o = 3 -- the pair straddles two blocks, or wraps -- a scalar for the second cell: 1.25 reductions per row
    const int o = x0 & 3;
    const size_t xb = (size_t)(x0 - o);
    const bool s0 = o == 0, s1 = o == 1, s2 = o == 2, s3 = o == 3;
#define BAOREC_QUAD(row, va, vb)                                                                            \
  red_add_v4(rho + (row) + xb, s0 ? (va) : 0.f, s0 ? (vb) : (s1 ? (va) : 0.f), s1 ? (vb) : (s2 ? (va) : 0.f), \
             s2 ? (vb) : (s3 ? (va) : 0.f));                                                                \
  if (s3) atomicAdd(rho + (row) + x1, (vb));
    BAOREC_QUAD(r00, v000, v100)
    BAOREC_QUAD(r10, v010, v110)
    BAOREC_QUAD(r01, v001, v101)
    BAOREC_QUAD(r11, v011, v111)
#undef BAOREC_QUAD
  } else if (((x0 | g.n[0]) & 1) == 0 && x1 == x0 + 1) {  // rows start at even offsets (nx even), x0 even: aligned pairs
    red_add_v2(rho + r00 + x0, v000, v100);
    red_add_v2(rho + r10 + x0, v010, v110);
    red_add_v2(rho + r01 + x0, v001, v101);
    red_add_v2(rho + r11 + x0, v011, v111);
  } else {
    atomicAdd(rho + r00 + x0, v000);
    atomicAdd(rho + r00 + x1, v100);
    atomicAdd(rho + r10 + x0, v010);
    atomicAdd(rho + r01 + x0, v001);
    atomicAdd(rho + r10 + x1, v110);
    atomicAdd(rho + r01 + x1, v101);
    atomicAdd(rho + r11 + x0, v011);
    atomicAdd(rho + r11 + x1, v111);
  }
  return true;
}

// TSC with vector reductions (option "scatter_pairs"): the three x-neighbours of a stencil row lie in ONE aligned
// 16-byte block (first cell at offset 0 or 1 of the block) or in that block and the first half of the next (offset 2
// or 3): one quad, plus one pair for half of the particles, +0 in the unused slots -- 13.5 reductions per particle
// on average instead of 27, and two issued instructions per row instead of three.
// Rows that touch the periodic wrap, and meshes whose row length is not a multiple of 4, use the scalar reductions.
__device__ __forceinline__ void red_row3(float* row, int first, float v0, float v1, float v2) {
  // branch-free: one quad on the aligned block that holds `first` (values at offset o, +0 elsewhere) and, for o >= 2,
  // one pair on the first half of the next block
  const int o = first & 3;
  float* b = row + (first - o);
  const bool s0 = o == 0, s1 = o == 1, s2 = o == 2, s3 = o == 3;
  red_add_v4(b, s0 ? v0 : 0.f, s0 ? v1 : (s1 ? v0 : 0.f), s0 ? v2 : (s1 ? v1 : (s2 ? v0 : 0.f)),
             s1 ? v2 : (s2 ? v1 : (s3 ? v0 : 0.f)));
  if (o >= 2) red_add_v2(b + 4, s2 ? v2 : v1, s2 ? 0.f : v2);
}

// deposit<TSC> with vector reductions; `rho` must be 16-byte aligned.  Same cells, same Float32 values.
__device__ __forceinline__ bool deposit_tsc_vec(float* __restrict__ rho, float px, float py, float pz, float ww,
                                                const BoxGeom& g, bool wrap) {
  const size_t nx = g.n[0], ny = g.n[1];
  int ix[3], iy[3], iz[3];
  float wx[3], wy[3], wz[3];
  bool ok = tsc_axis(px, g.mn[0], g.L[0], g.n[0], wrap, ix, wx);
  ok = tsc_axis(py, g.mn[1], g.L[1], g.n[1], wrap, iy, wy) && ok;
  ok = tsc_axis(pz, g.mn[2], g.L[2], g.n[2], wrap, iz, wz) && ok;
  if (!ok) return false;
  if (g.slab) {
#pragma unroll
    for (int c = 0; c < 3; c++) ok = local_plane1(g, iz[c], iz[c]) && ok;
    if (!ok) return false;
  }
  // the row is contiguous unless the stencil wraps in x; first + 2 < nx and nx = 4 k keep both pairs inside the row
  const bool vec = (g.n[0] & 3) == 0 && ix[1] == ix[0] + 1 && ix[2] == ix[0] + 2;
#pragma unroll
  for (int c = 0; c < 3; c++) {
#pragma unroll
    for (int b = 0; b < 3; b++) {
      float* row = rho + ((size_t)iz[c] * ny + iy[b]) * nx;
      const float v0 = __fmul_rn(__fmul_rn(__fmul_rn(wx[0], ww), wy[b]), wz[c]);
      const float v1 = __fmul_rn(__fmul_rn(__fmul_rn(wx[1], ww), wy[b]), wz[c]);
      const float v2 = __fmul_rn(__fmul_rn(__fmul_rn(wx[2], ww), wy[b]), wz[c]);
      if (vec) {
        red_row3(row, ix[0], v0, v1, v2);
      } else {
        atomicAdd(row + ix[0], v0);
        atomicAdd(row + ix[1], v1);
        atomicAdd(row + ix[2], v2);
      }
    }
  }
  return true;
}

// Four contiguous cells of a row (PCS): the aligned 16-byte block that holds `first` takes the values at offset o
// (+0 in the slots before it), the next block the rest -- one quad for a quarter of the particles, two for the others.
__device__ __forceinline__ void red_row4(float* row, int first, float v0, float v1, float v2, float v3) {
  const int o = first & 3;
  float* b = row + (first - o);
  const bool s0 = o == 0, s1 = o == 1, s2 = o == 2;
  red_add_v4(b, s0 ? v0 : 0.f, s0 ? v1 : (s1 ? v0 : 0.f), s0 ? v2 : (s1 ? v1 : (s2 ? v0 : 0.f)), s0 ? v3 : (s1 ? v2 : (s2 ? v1 : v0)));
  if (o != 0) red_add_v4(b + 4, s1 ? v3 : (s2 ? v2 : v1), s1 ? 0.f : (s2 ? v3 : v2), (s1 || s2) ? 0.f : v3, 0.f);
}

// deposit<PCS> with vector reductions (16 rows of four cells: 16 - 32 quads instead of 64 scalar reductions); `rho` must
// be 16-byte aligned.  Same cells, same Float32 values.
__device__ __forceinline__ bool deposit_pcs_vec(float* __restrict__ rho, float px, float py, float pz, float ww,
                                                const BoxGeom& g, bool wrap) {
  const size_t nx = g.n[0], ny = g.n[1];
  int ix[4], iy[4], iz[4];
  float wx[4], wy[4], wz[4];
  bool ok = pcs_axis(px, g.mn[0], g.L[0], g.n[0], wrap, ix, wx);
  ok = pcs_axis(py, g.mn[1], g.L[1], g.n[1], wrap, iy, wy) && ok;
  ok = pcs_axis(pz, g.mn[2], g.L[2], g.n[2], wrap, iz, wz) && ok;
  if (!ok) return false;
  if (g.slab) {
#pragma unroll
    for (int c = 0; c < 4; c++) ok = local_plane1(g, iz[c], iz[c]) && ok;
    if (!ok) return false;
  }
  // the row is contiguous unless the stencil wraps in x; first + 3 < nx and nx = 4 k keep both blocks inside the row
  const bool vec = (g.n[0] & 3) == 0 && ix[3] == ix[0] + 3;
  const float x0 = __fmul_rn(wx[0], ww), x1 = __fmul_rn(wx[1], ww), x2 = __fmul_rn(wx[2], ww), x3 = __fmul_rn(wx[3], ww);
#pragma unroll
  for (int c = 0; c < 4; c++) {
#pragma unroll
    for (int b = 0; b < 4; b++) {
      float* row = rho + ((size_t)iz[c] * ny + iy[b]) * nx;
      const float v0 = __fmul_rn(__fmul_rn(x0, wy[b]), wz[c]), v1 = __fmul_rn(__fmul_rn(x1, wy[b]), wz[c]);
      const float v2 = __fmul_rn(__fmul_rn(x2, wy[b]), wz[c]), v3 = __fmul_rn(__fmul_rn(x3, wy[b]), wz[c]);
      if (vec) {
        red_row4(row, ix[0], v0, v1, v2, v3);
      } else {
        atomicAdd(row + ix[0], v0);
        atomicAdd(row + ix[1], v1);
        atomicAdd(row + ix[2], v2);
        atomicAdd(row + ix[3], v3);
      }
    }
  }
  return true;
}

// ---- deterministic deposit (option "deterministic_scatter") -------------------------------------------------------
// Float additions do not commute, so a mesh accumulated with float reductions differs from run to run in the last
// bits (and with it the cells that sit on the `ran > threshold` cut).  Integer additions do
// commute: every Float32 deposit value is converted to 2^-40 fixed point -- exactly, unless it is below 2^-17, where
// the bits under 2^-40 are rounded away -- and accumulated with 64-bit integer reductions; the sum is then rounded to
// Float32 ONCE.  Bit-reproducible whatever the order, and closer to the exact sum than any Float32 summation order.
constexpr float kFixedScale = 1099511627776.0f;          // 2^40
constexpr double kFixedInv = 1.0 / 1099511627776.0;

__device__ __forceinline__ unsigned long long to_fixed(float v) {
#if defined(__CUDA_ARCH__)
  return (unsigned long long)__float2ll_rn(__fmul_rn(v, kFixedScale));
#else
  return (unsigned long long)llrintf(v * kFixedScale);
#endif
}

__device__ __forceinline__ float from_fixed(unsigned long long a) { return (float)((double)(long long)a * kFixedInv); }

// cic! for one (already wrapped) particle into the fixed-point accumulator: the cells, weights and the eight
// products are deposit<CIC>'s, only the accumulation differs.  Whole mesh only (no slab layout).
__device__ __forceinline__ bool deposit_fixed(unsigned long long* __restrict__ acc, float px, float py, float pz, float ww,
                                              const BoxGeom& g, bool wrap) {
  const size_t nx = g.n[0], ny = g.n[1];
  int x0, x1, y0, y1, z0, z1;
  float wx0, wx1, wy0, wy1, wz0, wz1;
  bool ok = cic_axis(px, g.mn[0], g.L[0], g.n[0], wrap, x0, x1, wx0, wx1);
  ok = cic_axis(py, g.mn[1], g.L[1], g.n[1], wrap, y0, y1, wy0, wy1) && ok;
  ok = cic_axis(pz, g.mn[2], g.L[2], g.n[2], wrap, z0, z1, wz0, wz1) && ok;
  if (!ok) return false;
  wx0 = __fmul_rn(wx0, ww);
  wx1 = __fmul_rn(wx1, ww);
  const size_t r00 = ((size_t)z0 * ny + y0) * nx, r10 = ((size_t)z0 * ny + y1) * nx;
  const size_t r01 = ((size_t)z1 * ny + y0) * nx, r11 = ((size_t)z1 * ny + y1) * nx;
  const float a00 = __fmul_rn(wx0, wy0), a10 = __fmul_rn(wx1, wy0), a01 = __fmul_rn(wx0, wy1), a11 = __fmul_rn(wx1, wy1);
  atomicAdd(acc + r00 + x0, to_fixed(__fmul_rn(a00, wz0)));
  atomicAdd(acc + r00 + x1, to_fixed(__fmul_rn(a10, wz0)));
  atomicAdd(acc + r10 + x0, to_fixed(__fmul_rn(a01, wz0)));
  atomicAdd(acc + r01 + x0, to_fixed(__fmul_rn(a00, wz1)));
  atomicAdd(acc + r10 + x1, to_fixed(__fmul_rn(a11, wz0)));
  atomicAdd(acc + r01 + x1, to_fixed(__fmul_rn(a10, wz1)));
  atomicAdd(acc + r11 + x0, to_fixed(__fmul_rn(a01, wz1)));
  atomicAdd(acc + r11 + x1, to_fixed(__fmul_rn(a11, wz1)));
  return true;
}

// ---- read_shifts / reconstructed_positions epilogue -----------------------------------------------------------
struct GatherArgs {
  const float* f[3];
  const float* x;
  const float* y;
  const float* z;
  float* o[3];
  int64_t n;
  int field;      // BAOREC_FIELD_*
  int positions;  // write pos - shift
  int has_los;
  float los[3];
  float fgrowth;
  float4* sorted_out;  // if non-null: write (s0,s1,s2,0) at the record's sorted position instead of o[c][idx]
};

// read_grad_cic! -- the finite-difference read-back the reference sketches and leaves commented out
// (src/mas.jl:388-466; the call site src/multigrid.jl:759, 785): grad phi at the particle = CIC interpolation of the
// central differences (phi[i+1] - phi[i-1]) at the eight corners, divided by 2 cell.  One gather over phi replaces the
// 1 R2C + 3 C2R + 3 gathers of the spectral read-back.  Restated with the INTENDED geometry: the sketch computes the
// particle's cell with the doubled cell size it needs for the difference quotient (dist = (p - min) / (2 L / n)), which
// would halve every coordinate; here the cell is read_cic!'s (src/mas.jl:221-224: d = (p - min) / T(L / n)) and only the
// final division uses 2 L / n.  Same accumulation order as the sketch: corners 000, 100, 010, 001, 110, 101, 011, 111
// (letters = x, y, z upper), weight = (wx wy) wz left to right, sum left to right, one division at the end.
__device__ __forceinline__ bool fd_axis(float p, float mn, float cell, int n, int idx[4], float& wd, float& wu) {
  float d = __fdiv_rn(__fsub_rn(p, mn), cell);
  if (!(d >= 0.0f && d < (float)(2 * n))) return false;
  float f = floorf(d);
  wu = __fsub_rn(d, f);
  wd = __fsub_rn(1.0f, wu);
  int i = (int)f + 1;  // 1-based like the reference
  if (i > n) i -= n;
  int u = i + 1;
  if (u > n) u -= n;
  int pp = u + 1;
  if (pp > n) pp -= n;
  int m = i - 1;
  if (m < 1) m += n;
  idx[0] = m - 1;   // i - 1
  idx[1] = i - 1;   // i
  idx[2] = u - 1;   // i + 1
  idx[3] = pp - 1;  // i + 2
  return true;
}

__device__ __forceinline__ bool fd_gradient(const float* __restrict__ phi, const BoxGeom& g, float px, float py, float pz,
                                            float (&out)[3]) {
  int X[4], Y[4], Z[4];
  float wx, dx, wy, dy, wz, dz;  // w = lower weight (1 - u), d = upper weight (u), the sketch's names
  bool ok = fd_axis(px, g.mn[0], g.cell[0], g.n[0], X, wx, dx);
  ok = fd_axis(py, g.mn[1], g.cell[1], g.n[1], Y, wy, dy) && ok;
  ok = fd_axis(pz, g.mn[2], g.cell[2], g.n[2], Z, wz, dz) && ok;
  if (!ok) return false;
  const size_t nx = g.n[0], ny = g.n[1];
#define PHI(ix, iy, iz) __ldg(phi + ((size_t)Z[iz] * ny + Y[iy]) * nx + X[ix])
  float gx = 0.f, gy = 0.f, gz = 0.f;
#pragma unroll
  for (int k = 0; k < 8; k++) {
    // corner order of the sketch: 000, 100, 010, 001, 110, 101, 011, 111
    const int cx = (k == 1 || k == 4 || k == 5 || k == 7), cy = (k == 2 || k == 4 || k == 6 || k == 7),
              cz = (k == 3 || k == 5 || k == 6 || k == 7);
    const float wt = __fmul_rn(__fmul_rn(cx ? dx : wx, cy ? dy : wy), cz ? dz : wz);
    const float ax = __fmul_rn(__fsub_rn(PHI(2 + cx, 1 + cy, 1 + cz), PHI(cx, 1 + cy, 1 + cz)), wt);
    const float ay = __fmul_rn(__fsub_rn(PHI(1 + cx, 2 + cy, 1 + cz), PHI(1 + cx, cy, 1 + cz)), wt);
    const float az = __fmul_rn(__fsub_rn(PHI(1 + cx, 1 + cy, 2 + cz), PHI(1 + cx, 1 + cy, cz)), wt);
    gx = k ? __fadd_rn(gx, ax) : ax;
    gy = k ? __fadd_rn(gy, ay) : ay;
    gz = k ? __fadd_rn(gz, az) : az;
  }
#undef PHI
  out[0] = __fdiv_rn(gx, __fmul_rn(2.0f, g.cell[0]));
  out[1] = __fdiv_rn(gy, __fmul_rn(2.0f, g.cell[1]));
  out[2] = __fdiv_rn(gz, __fmul_rn(2.0f, g.cell[2]));
  return true;
}

// read_shifts epilogue (src/recon.jl:277-304 / kernels :308-330) and optionally pos - shift (:376-378)
template <int NF>
__device__ __forceinline__ void shifts_epilogue(const GatherArgs& a, const float (&val)[NF], float px, float py,
                                                float pz, int64_t out_idx, int64_t sorted_pos = -1) {
  if (NF == 1) {
    a.o[0][out_idx] = val[0];
    return;
  }
  float s0 = val[0], s1 = val[NF > 1 ? 1 : 0], s2 = val[NF > 2 ? 2 : 0];
  if (a.field != BAOREC_FIELD_DISP) {
    float lx, ly, lz;
    if (a.has_los) {
      lx = a.los[0];
      ly = a.los[1];
      lz = a.los[2];
    } else {
      float dist = __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(px, px), __fmul_rn(py, py)), __fmul_rn(pz, pz)));
      lx = __fdiv_rn(px, dist);
      ly = __fdiv_rn(py, dist);
      lz = __fdiv_rn(pz, dist);
    }
    float dot = __fadd_rn(__fadd_rn(__fmul_rn(s0, lx), __fmul_rn(s1, ly)), __fmul_rn(s2, lz));
    float fd = __fmul_rn(a.fgrowth, dot);
    float r0 = __fmul_rn(fd, lx), r1 = __fmul_rn(fd, ly), r2 = __fmul_rn(fd, lz);
    if (a.field == BAOREC_FIELD_RSD) {
      s0 = r0;
      s1 = r1;
      s2 = r2;
    } else {
      s0 = __fadd_rn(s0, r0);
      s1 = __fadd_rn(s1, r1);
      s2 = __fadd_rn(s2, r2);
    }
  }
  if (a.positions) {
    s0 = __fsub_rn(px, s0);
    s1 = __fsub_rn(py, s1);
    s2 = __fsub_rn(pz, s2);
  }
  if (a.sorted_out && sorted_pos >= 0) {
    a.sorted_out[sorted_pos] = make_float4(s0, s1, s2, 0.f);  // coalesced; un-permuted by unsort_kernel
    return;
  }
  a.o[0][out_idx] = s0;
  if (NF > 1) a.o[1][out_idx] = s1;
  if (NF > 2) a.o[2][out_idx] = s2;
}

}  // namespace baorec
