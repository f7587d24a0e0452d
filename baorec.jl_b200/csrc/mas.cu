// Mass assignment: CIC/TSC scatter (particles -> mesh) and gather (mesh -> particles),
// the index/weight probes used by the parity tests, and setup_box reductions.
//
// Coordinate arithmetic uses explicit round-to-nearest intrinsics (__fsub_rn, __fmul_rn,
// __fdiv_rn, __fadd_rn) so that nvcc never contracts it into FMAs: cell indices and weights
// are bit-identical to the reference's Float32 CPU arithmetic (src/mas.jl:7-35, 224-255).
#include "internal.cuh"

namespace baorec {

struct BoxGeom {
  float mn[3];
  float L[3];
  float cell[3];
  int n[3];
};

static BoxGeom geom_of(const baorec_ctx* ctx) {
  BoxGeom g;
  for (int a = 0; a < 3; a++) {
    g.mn[a] = ctx->mn[a];
    g.L[a] = ctx->L[a];
    g.cell[a] = ctx->cell[a];
  }
  g.n[0] = ctx->nx;
  g.n[1] = ctx->ny;
  g.n[2] = ctx->nz;
  return g;
}

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

__global__ void __launch_bounds__(256)
cic_scatter_kernel(float* __restrict__ rho, float* __restrict__ x, float* __restrict__ y, float* __restrict__ z,
                   const float* __restrict__ w, int64_t n, BoxGeom g, int wrap, unsigned long long* oob) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float px = x[i], py = y[i], pz = z[i];
  if (wrap) {
    float qx = wrap_pos(px, g.mn[0], g.L[0]);
    float qy = wrap_pos(py, g.mn[0], g.L[0]);
    float qz = wrap_pos(pz, g.mn[0], g.L[0]);
    if (qx != px) x[i] = qx;  // write-back like the reference (src/mas.jl:57-59)
    if (qy != py) y[i] = qy;
    if (qz != pz) z[i] = qz;
    px = qx;
    py = qy;
    pz = qz;
  }
  int x0, x1, y0, y1, z0, z1;
  float wx0, wx1, wy0, wy1, wz0, wz1;
  bool ok = cic_axis(px, g.mn[0], g.L[0], g.n[0], wrap, x0, x1, wx0, wx1);
  ok = cic_axis(py, g.mn[1], g.L[1], g.n[1], wrap, y0, y1, wy0, wy1) && ok;
  ok = cic_axis(pz, g.mn[2], g.L[2], g.n[2], wrap, z0, z1, wz0, wz1) && ok;
  if (!ok) {
    atomicAdd(oob, 1ULL);
    return;
  }
  float ww = w[i];
  wx0 = __fmul_rn(wx0, ww);
  wx1 = __fmul_rn(wx1, ww);
  size_t nx = g.n[0], ny = g.n[1];
  size_t r00 = ((size_t)z0 * ny + y0) * nx, r10 = ((size_t)z0 * ny + y1) * nx;
  size_t r01 = ((size_t)z1 * ny + y0) * nx, r11 = ((size_t)z1 * ny + y1) * nx;
  float a00 = __fmul_rn(wx0, wy0), a10 = __fmul_rn(wx1, wy0), a01 = __fmul_rn(wx0, wy1), a11 = __fmul_rn(wx1, wy1);
  atomicAdd(rho + r00 + x0, __fmul_rn(a00, wz0));
  atomicAdd(rho + r00 + x1, __fmul_rn(a10, wz0));
  atomicAdd(rho + r10 + x0, __fmul_rn(a01, wz0));
  atomicAdd(rho + r01 + x0, __fmul_rn(a00, wz1));
  atomicAdd(rho + r10 + x1, __fmul_rn(a11, wz0));
  atomicAdd(rho + r01 + x1, __fmul_rn(a10, wz1));
  atomicAdd(rho + r11 + x0, __fmul_rn(a01, wz1));
  atomicAdd(rho + r11 + x1, __fmul_rn(a11, wz1));
}

__global__ void __launch_bounds__(256)
tsc_scatter_kernel(float* __restrict__ rho, const float* __restrict__ x, const float* __restrict__ y,
                   const float* __restrict__ z, const float* __restrict__ w, int64_t n, BoxGeom g, int wrap,
                   unsigned long long* oob) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int ix[3], iy[3], iz[3];
  float wx[3], wy[3], wz[3];
  bool ok = tsc_axis(x[i], g.mn[0], g.L[0], g.n[0], wrap, ix, wx);
  ok = tsc_axis(y[i], g.mn[1], g.L[1], g.n[1], wrap, iy, wy) && ok;
  ok = tsc_axis(z[i], g.mn[2], g.L[2], g.n[2], wrap, iz, wz) && ok;
  if (!ok) {
    atomicAdd(oob, 1ULL);
    return;
  }
  float ww = w[i];
  size_t nx = g.n[0], ny = g.n[1];
#pragma unroll
  for (int c = 0; c < 3; c++) {
#pragma unroll
    for (int b = 0; b < 3; b++) {
      size_t row = ((size_t)iz[c] * ny + iy[b]) * nx;
#pragma unroll
      for (int a = 0; a < 3; a++) {
        float v = __fmul_rn(__fmul_rn(__fmul_rn(wx[a], ww), wy[b]), wz[c]);
        atomicAdd(rho + row + ix[a], v);
      }
    }
  }
}

__global__ void cic_cells_kernel(const float* __restrict__ x, const float* __restrict__ y, const float* __restrict__ z,
                                 int64_t n, BoxGeom g, int wrap, int32_t* i0, int32_t* i1, float* w0, float* w1) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float* p[3] = {x, y, z};
#pragma unroll
  for (int a = 0; a < 3; a++) {
    float v = p[a][i];
    if (wrap) v = wrap_pos(v, g.mn[0], g.L[0]);
    int c0 = -1, c1 = -1;
    float a0 = 0.f, a1 = 0.f;
    if (!cic_axis(v, g.mn[a], g.L[a], g.n[a], wrap, c0, c1, a0, a1)) {
      c0 = c1 = -1;
    }
    i0[a * n + i] = c0;
    i1[a * n + i] = c1;
    w0[a * n + i] = a0;
    w1[a * n + i] = a1;
  }
}

__global__ void gather_cells_kernel(const float* __restrict__ x, const float* __restrict__ y,
                                    const float* __restrict__ z, int64_t n, BoxGeom g, int gpu_formula, int32_t* id,
                                    int32_t* iu, float* wd, float* wu) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float* p[3] = {x, y, z};
#pragma unroll
  for (int a = 0; a < 3; a++) {
    int c0 = -1, c1 = -1;
    float a0 = 0.f, a1 = 0.f;
    if (!gather_axis(p[a][i], g.mn[a], g.L[a], g.cell[a], g.n[a], gpu_formula != 0, c0, c1, a0, a1)) c0 = c1 = -1;
    id[a * n + i] = c0;
    iu[a * n + i] = c1;
    wd[a * n + i] = a0;
    wu[a * n + i] = a1;
  }
}

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
};

// read_cic! (src/mas.jl:258-265): sum of field*wx*wy*wz, left-associated, in the order
// ddd,ddu,dud,duu,udd,udu,uud,uuu (letters = x,y,z).
template <int NF, int MAS>
__global__ void __launch_bounds__(256) gather_kernel(GatherArgs a, BoxGeom g, unsigned long long* oob) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= a.n) return;
  float px = a.x[i], py = a.y[i], pz = a.z[i];
  float val[NF];
  size_t nx = g.n[0], ny = g.n[1];
  bool ok;
  if (MAS == BAOREC_MAS_CIC) {
    int xd, xu, yd, yu, zd, zu;
    float dx, ux, dy, uy, dz, uz;
    ok = gather_axis(px, g.mn[0], g.L[0], g.cell[0], g.n[0], false, xd, xu, dx, ux);
    ok = gather_axis(py, g.mn[1], g.L[1], g.cell[1], g.n[1], false, yd, yu, dy, uy) && ok;
    ok = gather_axis(pz, g.mn[2], g.L[2], g.cell[2], g.n[2], false, zd, zu, dz, uz) && ok;
    if (ok) {
      size_t rdd = ((size_t)zd * ny + yd) * nx, rdu = ((size_t)zu * ny + yd) * nx;
      size_t rud = ((size_t)zd * ny + yu) * nx, ruu = ((size_t)zu * ny + yu) * nx;
#pragma unroll
      for (int c = 0; c < NF; c++) {
        const float* f = a.f[c];
        float v;
        v = __fmul_rn(__fmul_rn(__fmul_rn(__ldg(f + rdd + xd), dx), dy), dz);
        v = __fadd_rn(v, __fmul_rn(__fmul_rn(__fmul_rn(__ldg(f + rdu + xd), dx), dy), uz));
        v = __fadd_rn(v, __fmul_rn(__fmul_rn(__fmul_rn(__ldg(f + rud + xd), dx), uy), dz));
        v = __fadd_rn(v, __fmul_rn(__fmul_rn(__fmul_rn(__ldg(f + ruu + xd), dx), uy), uz));
        v = __fadd_rn(v, __fmul_rn(__fmul_rn(__fmul_rn(__ldg(f + rdd + xu), ux), dy), dz));
        v = __fadd_rn(v, __fmul_rn(__fmul_rn(__fmul_rn(__ldg(f + rdu + xu), ux), dy), uz));
        v = __fadd_rn(v, __fmul_rn(__fmul_rn(__fmul_rn(__ldg(f + rud + xu), ux), uy), dz));
        v = __fadd_rn(v, __fmul_rn(__fmul_rn(__fmul_rn(__ldg(f + ruu + xu), ux), uy), uz));
        val[c] = v;
      }
    }
  } else {
    int ix[3], iy[3], iz[3];
    float wx[3], wy[3], wz[3];
    ok = tsc_axis(px, g.mn[0], g.L[0], g.n[0], true, ix, wx);
    ok = tsc_axis(py, g.mn[1], g.L[1], g.n[1], true, iy, wy) && ok;
    ok = tsc_axis(pz, g.mn[2], g.L[2], g.n[2], true, iz, wz) && ok;
    if (ok) {
#pragma unroll
      for (int c = 0; c < NF; c++) val[c] = 0.f;
#pragma unroll
      for (int oz = 0; oz < 3; oz++)
#pragma unroll
        for (int oy = 0; oy < 3; oy++) {
          size_t row = ((size_t)iz[oz] * ny + iy[oy]) * nx;
#pragma unroll
          for (int ox = 0; ox < 3; ox++)
#pragma unroll
            for (int c = 0; c < NF; c++)
              val[c] = __fadd_rn(
                  val[c], __fmul_rn(__fmul_rn(__fmul_rn(__ldg(a.f[c] + row + ix[ox]), wx[ox]), wy[oy]), wz[oz]));
        }
    }
  }
  if (!ok) {
    atomicAdd(oob, 1ULL);
#pragma unroll
    for (int c = 0; c < NF; c++) val[c] = 0.f;
  }
  if (NF == 1) {
    a.o[0][i] = val[0];
    return;
  }
  // read_shifts epilogue (src/recon.jl:277-304 / kernels :308-330)
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
  a.o[0][i] = s0;
  if (NF > 1) a.o[1][i] = s1;
  if (NF > 2) a.o[2][i] = s2;
}

// ---- setup_box reductions (src/utils.jl:100-109) --------------------------------------------
__device__ __forceinline__ unsigned f2ord(float f) {
  unsigned u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__host__ __device__ __forceinline__ float ord2f(unsigned u) {
  unsigned v = (u & 0x80000000u) ? (u & 0x7fffffffu) : ~u;
#ifdef __CUDA_ARCH__
  return __uint_as_float(v);
#else
  float f;
  memcpy(&f, &v, 4);
  return f;
#endif
}

__global__ void minmax_init_kernel(unsigned* mm) {
  int t = threadIdx.x;
  if (t < 3) mm[t] = 0xffffffffu;
  else if (t < 6) mm[t] = 0u;
}

__global__ void __launch_bounds__(256)
minmax_kernel(const float* __restrict__ x, const float* __restrict__ y, const float* __restrict__ z, int64_t n,
              unsigned* mm) {
  unsigned lo[3] = {0xffffffffu, 0xffffffffu, 0xffffffffu}, hi[3] = {0u, 0u, 0u};
  const float* p[3] = {x, y, z};
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
#pragma unroll
    for (int a = 0; a < 3; a++) {
      unsigned o = f2ord(p[a][i]);
      lo[a] = min(lo[a], o);
      hi[a] = max(hi[a], o);
    }
  }
#pragma unroll
  for (int a = 0; a < 3; a++) {
    lo[a] = __reduce_min_sync(0xffffffffu, lo[a]);
    hi[a] = __reduce_max_sync(0xffffffffu, hi[a]);
  }
  if ((threadIdx.x & 31) == 0) {
#pragma unroll
    for (int a = 0; a < 3; a++) {
      atomicMin(mm + a, lo[a]);
      atomicMax(mm + 3 + a, hi[a]);
    }
  }
}

int scatter(baorec_ctx* ctx, float* rho, float* x, float* y, float* z, const float* w, int64_t n, int wrap, int mas,
            cudaStream_t st) {
  if (n == 0) return BAOREC_OK;
  BoxGeom g = geom_of(ctx);
  unsigned grid = cdiv((size_t)n, 256);
  if (mas == BAOREC_MAS_TSC) {
    BR_LAUNCH(ctx, tsc_scatter_kernel, grid, 256, 0, st, rho, x, y, z, w, n, g, wrap, ctx->d_oob);
  } else {
    BR_LAUNCH(ctx, cic_scatter_kernel, grid, 256, 0, st, rho, x, y, z, w, n, g, wrap, ctx->d_oob);
  }
  return BAOREC_OK;
}

int gather3(baorec_ctx* ctx, const float* fx, const float* fy, const float* fz, const float* x, const float* y,
            const float* z, int64_t n, float* ox, float* oy, float* oz, int mas, int field, float f, int has_los,
            const float* los, int positions, cudaStream_t st) {
  if (n == 0) return BAOREC_OK;
  BoxGeom g = geom_of(ctx);
  GatherArgs a;
  a.f[0] = fx;
  a.f[1] = fy;
  a.f[2] = fz;
  a.x = x;
  a.y = y;
  a.z = z;
  a.o[0] = ox;
  a.o[1] = oy;
  a.o[2] = oz;
  a.n = n;
  a.field = field;
  a.positions = positions;
  a.has_los = has_los;
  for (int c = 0; c < 3; c++) a.los[c] = (has_los && los) ? los[c] : 0.f;
  a.fgrowth = f;
  unsigned grid = cdiv((size_t)n, 256);
  bool one = (fy == nullptr);
  if (mas == BAOREC_MAS_TSC) {
    if (one) BR_LAUNCH(ctx, (gather_kernel<1, BAOREC_MAS_TSC>), grid, 256, 0, st, a, g, ctx->d_oob);
    else BR_LAUNCH(ctx, (gather_kernel<3, BAOREC_MAS_TSC>), grid, 256, 0, st, a, g, ctx->d_oob);
  } else {
    if (one) BR_LAUNCH(ctx, (gather_kernel<1, BAOREC_MAS_CIC>), grid, 256, 0, st, a, g, ctx->d_oob);
    else BR_LAUNCH(ctx, (gather_kernel<3, BAOREC_MAS_CIC>), grid, 256, 0, st, a, g, ctx->d_oob);
  }
  return BAOREC_OK;
}

int setup_box_dev(baorec_ctx* ctx, const float* x, const float* y, const float* z, int64_t n, float pad,
                  float L_out[3], float mn_out[3], cudaStream_t st) {
  BR_REQUIRE(n > 0, "setup_box needs at least one particle");
  unsigned* mm = (unsigned*)ctx->d_minmax;
  BR_LAUNCH(ctx, minmax_init_kernel, 1, 32, 0, st, mm);
  unsigned grid = cdiv((size_t)n, 256);
  if (grid > 148 * 16) grid = 148 * 16;
  BR_LAUNCH(ctx, minmax_kernel, grid, 256, 0, st, x, y, z, n, mm);
  unsigned h[6];
  BR_CUDA(cudaMemcpyAsync(h, mm, sizeof(h), cudaMemcpyDeviceToHost, st));
  BR_CUDA(cudaStreamSynchronize(st));
  // box_min = min - pad/2 ; box_max = max + pad/2 ; size = max over axes of (max - min), cubic.
  float half = pad / 2.0f, ext = 0.f;
  for (int a = 0; a < 3; a++) {
    float lo = ord2f(h[a]) - half, hi = ord2f(h[3 + a]) + half;
    mn_out[a] = lo;
    float e = hi - lo;
    if (a == 0 || e > ext) ext = e;
  }
  for (int a = 0; a < 3; a++) L_out[a] = ext;
  return BAOREC_OK;
}

}  // namespace baorec

using namespace baorec;

extern "C" {

int baorec_cic_scatter_f32(baorec_ctx* ctx, float* d_rho, float* d_x, float* d_y, float* d_z, const float* d_w,
                           int64_t n, int wrap, int mas, baorec_stream stream) {
  BR_NEED_PLAN(ctx);
  BR_REQUIRE(n >= 0, "n < 0");
  BR_REQUIRE(n == 0 || (d_rho && d_x && d_y && d_z && d_w), "NULL device pointer");
  BR_REQUIRE(mas == BAOREC_MAS_CIC || mas == BAOREC_MAS_TSC, "unknown mas");
  cudaStream_t st = (cudaStream_t)stream;
  BR_TRY(reset_oob(ctx, st));
  BR_TRY(scatter(ctx, d_rho, d_x, d_y, d_z, d_w, n, wrap, mas, st));
  return check_oob(ctx, st, "cic scatter");
}

int baorec_cic_cells_f32(baorec_ctx* ctx, const float* d_x, const float* d_y, const float* d_z, int64_t n, int wrap,
                         int32_t* d_i0, int32_t* d_i1, float* d_w0, float* d_w1, baorec_stream stream) {
  BR_NEED_PLAN(ctx);
  BR_REQUIRE(n >= 0, "n < 0");
  if (n == 0) return BAOREC_OK;
  BR_REQUIRE(d_x && d_y && d_z && d_i0 && d_i1 && d_w0 && d_w1, "NULL device pointer");
  BR_LAUNCH(ctx, cic_cells_kernel, cdiv((size_t)n, 256), 256, 0, (cudaStream_t)stream, d_x, d_y, d_z, n, geom_of(ctx),
            wrap, d_i0, d_i1, d_w0, d_w1);
  return BAOREC_OK;
}

int baorec_gather_cells_f32(baorec_ctx* ctx, const float* d_x, const float* d_y, const float* d_z, int64_t n,
                            int gpu_formula, int32_t* d_id, int32_t* d_iu, float* d_wd, float* d_wu,
                            baorec_stream stream) {
  BR_NEED_PLAN(ctx);
  BR_REQUIRE(n >= 0, "n < 0");
  if (n == 0) return BAOREC_OK;
  BR_REQUIRE(d_x && d_y && d_z && d_id && d_iu && d_wd && d_wu, "NULL device pointer");
  BR_LAUNCH(ctx, gather_cells_kernel, cdiv((size_t)n, 256), 256, 0, (cudaStream_t)stream, d_x, d_y, d_z, n,
            geom_of(ctx), gpu_formula, d_id, d_iu, d_wd, d_wu);
  return BAOREC_OK;
}

int baorec_gather_f32(baorec_ctx* ctx, const float* d_field, const float* d_x, const float* d_y, const float* d_z,
                      int64_t n, float* d_out, int mas, baorec_stream stream) {
  BR_NEED_PLAN(ctx);
  BR_REQUIRE(n >= 0, "n < 0");
  if (n == 0) return BAOREC_OK;
  BR_REQUIRE(d_field && d_x && d_y && d_z && d_out, "NULL device pointer");
  cudaStream_t st = (cudaStream_t)stream;
  BR_TRY(reset_oob(ctx, st));
  BR_TRY(gather3(ctx, d_field, nullptr, nullptr, d_x, d_y, d_z, n, d_out, nullptr, nullptr, mas, BAOREC_FIELD_DISP,
                 0.f, 0, nullptr, 0, st));
  return check_oob(ctx, st, "gather");
}

int baorec_setup_box_f32(baorec_ctx* ctx, const float* d_x, const float* d_y, const float* d_z, int64_t n, float pad,
                         float box_size_out[3], float box_min_out[3], baorec_stream stream) {
  BR_REQUIRE(ctx != nullptr, "ctx is NULL");
  BR_REQUIRE(d_x && d_y && d_z && box_size_out && box_min_out, "NULL pointer");
  BR_CUDA(cudaSetDevice(ctx->device));
  return setup_box_dev(ctx, d_x, d_y, d_z, n, pad, box_size_out, box_min_out, (cudaStream_t)stream);
}

}  // extern "C"
