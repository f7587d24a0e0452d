// Mass assignment: CIC/TSC scatter (particles -> mesh) and gather (mesh -> particles),
// the index/weight probes used by the parity tests, and setup_box reductions.
//
// Coordinate arithmetic uses explicit round-to-nearest intrinsics (__fsub_rn, __fmul_rn,
// __fdiv_rn, __fadd_rn) so that nvcc never contracts it into FMAs: cell indices and weights
// are bit-identical to the reference's Float32 CPU arithmetic (src/mas.jl:7-35, 224-255).
//
// Large catalogs go through a z-slab binning pass first (one counting sort on the base z
// plane, ZG planes per bin): the scatter's float atomics and the gather's 8x3 loads then
// walk the mesh one (ZG+1)-plane window at a time, which stays resident in the 126 MB L2,
// so HBM sees each mesh sector once instead of once per particle.
#include <cuda.h>
#include <string.h>

#include "internal.cuh"
#include "mas_math.cuh"

namespace baorec {

static BoxGeom geom_for(const baorec_ctx* ctx, int slab_mode) {
  BoxGeom g;
  for (int a = 0; a < 3; a++) {
    g.mn[a] = ctx->mn[a];
    g.L[a] = ctx->L[a];
    g.cell[a] = ctx->cell[a];
  }
  g.n[0] = ctx->nx;
  g.n[1] = ctx->ny;
  g.n[2] = ctx->nz;
  g.slab = slab_mode != 0;
  g.z_lo = g.slab ? ctx->z0 : 0;
  // slab_mode 1: CIC scatter (one ghost plane above); 2: gather (one halo plane below, two above);
  // 3: TSC scatter (one ghost plane below, two above -- the same layout as the gather's)
  g.zoff = slab_mode >= 2 ? 1 : 0;
  g.nzp = slab_mode == 1 ? ctx->nz_loc + 1 : (slab_mode >= 2 ? ctx->nz_loc + 3 : ctx->nz);
  return g;
}
static BoxGeom geom_of(const baorec_ctx* ctx) { return geom_for(ctx, ctx->slab_mode); }

// read_cic! (src/mas.jl:258-265): sum of field*wx*wy*wz, left-associated, in the order
// ddd,ddu,dud,duu,udd,udu,uud,uuu (letters = x,y,z); then the read_shifts epilogue
// (src/recon.jl:277-304 / kernels :308-330) and optionally pos - shift (:376-378).
template <int NF, int MAS>
__device__ __forceinline__ bool gather_one(const GatherArgs& a, const BoxGeom& g, float px, float py, float pz,
                                           int64_t out_idx) {
  float val[NF];
  const size_t nx = g.n[0], ny = g.n[1];
  bool ok;
  if (MAS == BAOREC_MAS_CIC) {
    int xd, xu, yd, yu, zd, zu;
    float dx, ux, dy, uy, dz, uz;
    ok = gather_axis(px, g.mn[0], g.L[0], g.cell[0], g.n[0], false, xd, xu, dx, ux);
    ok = gather_axis(py, g.mn[1], g.L[1], g.cell[1], g.n[1], false, yd, yu, dy, uy) && ok;
    ok = gather_axis(pz, g.mn[2], g.L[2], g.cell[2], g.n[2], false, zd, zu, dz, uz) && ok;
    ok = ok && local_planes(g, zd, zu, zd, zu);
    if (ok) {
      size_t rdd = ((size_t)zd * ny + yd) * nx, rdu = ((size_t)zu * ny + yd) * nx;
      size_t rud = ((size_t)zd * ny + yu) * nx, ruu = ((size_t)zu * ny + yu) * nx;
      float l[NF][8];
#pragma unroll
      for (int c = 0; c < NF; c++) {  // issue all loads first (memory-level parallelism)
        const float* f = a.f[c];
        l[c][0] = __ldg(f + rdd + xd);
        l[c][1] = __ldg(f + rdu + xd);
        l[c][2] = __ldg(f + rud + xd);
        l[c][3] = __ldg(f + ruu + xd);
        l[c][4] = __ldg(f + rdd + xu);
        l[c][5] = __ldg(f + rdu + xu);
        l[c][6] = __ldg(f + rud + xu);
        l[c][7] = __ldg(f + ruu + xu);
      }
#pragma unroll
      for (int c = 0; c < NF; c++) {
        float v;
        v = __fmul_rn(__fmul_rn(__fmul_rn(l[c][0], dx), dy), dz);
        v = __fadd_rn(v, __fmul_rn(__fmul_rn(__fmul_rn(l[c][1], dx), dy), uz));
        v = __fadd_rn(v, __fmul_rn(__fmul_rn(__fmul_rn(l[c][2], dx), uy), dz));
        v = __fadd_rn(v, __fmul_rn(__fmul_rn(__fmul_rn(l[c][3], dx), uy), uz));
        v = __fadd_rn(v, __fmul_rn(__fmul_rn(__fmul_rn(l[c][4], ux), dy), dz));
        v = __fadd_rn(v, __fmul_rn(__fmul_rn(__fmul_rn(l[c][5], ux), dy), uz));
        v = __fadd_rn(v, __fmul_rn(__fmul_rn(__fmul_rn(l[c][6], ux), uy), dz));
        v = __fadd_rn(v, __fmul_rn(__fmul_rn(__fmul_rn(l[c][7], ux), uy), uz));
        val[c] = v;
      }
    }
  } else {
    constexpr int SW = MasStencil<MAS>::SW;
    int ix[4], iy[4], iz[4];
    float wx[4], wy[4], wz[4];
    ok = stencil_axis<MAS>(px, g.mn[0], g.L[0], g.n[0], true, ix, wx);
    ok = stencil_axis<MAS>(py, g.mn[1], g.L[1], g.n[1], true, iy, wy) && ok;
    ok = stencil_axis<MAS>(pz, g.mn[2], g.L[2], g.n[2], true, iz, wz) && ok;
    if (ok && g.slab) {  // slab layout (multi-GPU): halo planes below / above the slab
#pragma unroll
      for (int c = 0; c < SW; c++) ok = local_plane1(g, iz[c], iz[c]) && ok;
    }
    if (ok) {
#pragma unroll
      for (int c = 0; c < NF; c++) val[c] = 0.f;
#pragma unroll
      for (int oz = 0; oz < SW; oz++)
#pragma unroll
        for (int oy = 0; oy < SW; oy++) {
          size_t row = ((size_t)iz[oz] * ny + iy[oy]) * nx;
#pragma unroll
          for (int ox = 0; ox < SW; ox++)
#pragma unroll
            for (int c = 0; c < NF; c++)
              val[c] = __fadd_rn(
                  val[c], __fmul_rn(__fmul_rn(__fmul_rn(__ldg(a.f[c] + row + ix[ox]), wx[ox]), wy[oy]), wz[oz]));
        }
    }
  }
  if (!ok) {
#pragma unroll
    for (int c = 0; c < NF; c++) val[c] = 0.f;
  }
  shifts_epilogue<NF>(a, val, px, py, pz, out_idx);
  return ok;
}

// ---- direct (unbinned) kernels: one thread per particle in catalog order -------------------------
template <int MAS>
__global__ void __launch_bounds__(256)
scatter_direct_kernel(float* __restrict__ rho, float* __restrict__ x, float* __restrict__ y, float* __restrict__ z,
                      const float* __restrict__ w, int64_t n, BoxGeom g, int wrap, unsigned long long* oob) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float px = x[i], py = y[i], pz = z[i];
  if (wrap && MAS == BAOREC_MAS_CIC) {
    float qx = wrap_pos(px, g.mn[0], g.L[0]);
    float qy = wrap_pos(py, g.mn[0], g.L[0]);
    float qz = wrap_pos(pz, g.mn[0], g.L[0]);
    if (qx != px) x[i] = qx;  // write-back like the reference (src/mas.jl:57-59)
    if (qy != py) y[i] = qy;
    if (qz != pz) z[i] = qz;
    if (qx != px || qy != py || qz != pz) atomicAdd(oob + 1, 1ULL);
    px = qx;
    py = qy;
    pz = qz;
  }
  if (!deposit<MAS>(rho, px, py, pz, w[i], g, wrap != 0)) atomicAdd(oob, 1ULL);
}

template <int NF, int MAS>
__global__ void __launch_bounds__(256) gather_direct_kernel(GatherArgs a, BoxGeom g, unsigned long long* oob) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= a.n) return;
  if (!gather_one<NF, MAS>(a, g, a.x[i], a.y[i], a.z[i], i)) atomicAdd(oob, 1ULL);
}

// ---- z-slab binning ---------------------------------------------------------------------------------
// Key = base z plane / zg (scatter: cic z0 after wrap, or TSC centre; gather: lower z index).
// Out-of-box particles get key = nbins (a trash bin that the sorted kernels never visit).
constexpr int BIN_THREADS = 256;
constexpr int BIN_PPT = 8;  // particles per thread in the reorder kernel
constexpr int BIN_MAX = 1024;

enum { BIN_SCATTER = 0, BIN_GATHER = 1 };

template <int MODE, int MAS>
__device__ __forceinline__ int bin_key(float& px, float& py, float& pz, const BoxGeom& g, int wrap, int zg, int nbins,
                                       bool& wrapped) {
  wrapped = false;
  bool ok;
  int zb;
  if (MODE == BIN_SCATTER && MAS == BAOREC_MAS_CIC) {
    if (wrap) {
      float qx = wrap_pos(px, g.mn[0], g.L[0]), qy = wrap_pos(py, g.mn[0], g.L[0]),
            qz = wrap_pos(pz, g.mn[0], g.L[0]);
      wrapped = (qx != px) || (qy != py) || (qz != pz);
      px = qx;
      py = qy;
      pz = qz;
    }
    int i0, i1;
    float w0, w1;
    ok = cic_axis(px, g.mn[0], g.L[0], g.n[0], wrap, i0, i1, w0, w1);
    ok = cic_axis(py, g.mn[1], g.L[1], g.n[1], wrap, i0, i1, w0, w1) && ok;
    ok = cic_axis(pz, g.mn[2], g.L[2], g.n[2], wrap, i0, i1, w0, w1) && ok;
    ok = ok && local_planes(g, i0, i1, i0, i1);
    zb = i0;
  } else if (MAS != BAOREC_MAS_CIC) {
    constexpr int SW = MasStencil<MAS>::SW;
    int idx[4];
    float w[4];
    bool wr = MODE == BIN_GATHER ? true : (wrap != 0);
    ok = stencil_axis<MAS>(px, g.mn[0], g.L[0], g.n[0], wr, idx, w);
    ok = stencil_axis<MAS>(py, g.mn[1], g.L[1], g.n[1], wr, idx, w) && ok;
    ok = stencil_axis<MAS>(pz, g.mn[2], g.L[2], g.n[2], wr, idx, w) && ok;
    zb = idx[1];
    if (ok && g.slab) {  // bins are planes of the local buffer: plane 1 of the stencil, provided the whole stencil is local
      int lo, hi;
      ok = local_plane1(g, idx[0], lo) && local_plane1(g, idx[SW - 1], hi) && local_plane1(g, idx[1], zb);
    }
  } else {
    int id, iu;
    float wd, wu;
    ok = gather_axis(px, g.mn[0], g.L[0], g.cell[0], g.n[0], false, id, iu, wd, wu);
    ok = gather_axis(py, g.mn[1], g.L[1], g.cell[1], g.n[1], false, id, iu, wd, wu) && ok;
    ok = gather_axis(pz, g.mn[2], g.L[2], g.cell[2], g.n[2], false, id, iu, wd, wu) && ok;
    ok = ok && local_planes(g, id, iu, id, iu);
    zb = id;
  }
  return ok ? zb / zg : nbins;
}

template <int MODE, int MAS>
__global__ void __launch_bounds__(BIN_THREADS)
bin_count_kernel(const float* __restrict__ x, const float* __restrict__ y, const float* __restrict__ z, int64_t n,
                 BoxGeom g, int wrap, int zg, int nbins, unsigned* __restrict__ counts) {
  __shared__ unsigned s_cnt[BIN_MAX + 1];
  for (int t = threadIdx.x; t <= nbins; t += blockDim.x) s_cnt[t] = 0;
  __syncthreads();
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    float px = x[i], py = y[i], pz = z[i];
    bool wr;
    int k = bin_key<MODE, MAS>(px, py, pz, g, wrap, zg, nbins, wr);
    atomicAdd(&s_cnt[k], 1u);
  }
  __syncthreads();
  for (int t = threadIdx.x; t <= nbins; t += blockDim.x)
    if (s_cnt[t]) atomicAdd(counts + t, s_cnt[t]);
}

// exclusive scan of counts[0..nbins] into cursor[0..nbins] (single block)
__global__ void bin_scan_kernel(const unsigned* __restrict__ counts, unsigned* __restrict__ cursor,
                                unsigned* __restrict__ starts, int nbins, unsigned long long* oob) {
  __shared__ unsigned s[BIN_MAX + 2];
  int t = threadIdx.x;
  for (int i = t; i <= nbins; i += blockDim.x) s[i] = counts[i];
  __syncthreads();
  if (t == 0) {
    unsigned acc = 0;
    for (int i = 0; i <= nbins; i++) {
      unsigned c = s[i];
      s[i] = acc;
      acc += c;
    }
    s[nbins + 1] = acc;
    if (counts[nbins]) atomicAdd(oob, (unsigned long long)counts[nbins]);
  }
  __syncthreads();
  for (int i = t; i <= nbins + 1; i += blockDim.x) {
    if (i <= nbins) cursor[i] = s[i];
    starts[i] = s[i];
  }
}

// Reorder into float4 records: scatter (x,y,z,w) ; gather (x,y,z, original index bits).
template <int MODE, int MAS>
__global__ void __launch_bounds__(BIN_THREADS)
bin_reorder_kernel(float* __restrict__ x, float* __restrict__ y, float* __restrict__ z, const float* __restrict__ w,
                   int64_t n, BoxGeom g, int wrap, int zg, int nbins, unsigned* __restrict__ cursor,
                   float4* __restrict__ rec, unsigned long long* oob) {
  __shared__ unsigned s_cnt[BIN_MAX + 1];
  __shared__ unsigned s_base[BIN_MAX + 1];
  for (int t = threadIdx.x; t <= nbins; t += blockDim.x) s_cnt[t] = 0;
  __syncthreads();
  const int64_t base = (int64_t)blockIdx.x * (BIN_THREADS * BIN_PPT);
  float4 r[BIN_PPT];
  int key[BIN_PPT];
  unsigned rank[BIN_PPT];
#pragma unroll
  for (int u = 0; u < BIN_PPT; u++) {
    int64_t i = base + u * BIN_THREADS + threadIdx.x;
    key[u] = -1;
    if (i < n) {
      float px = x[i], py = y[i], pz = z[i];
      float ox = px, oy = py, oz = pz;
      bool wr;
      key[u] = bin_key<MODE, MAS>(px, py, pz, g, wrap, zg, nbins, wr);
      if (MODE == BIN_SCATTER && wr) {  // cic! writes the wrapped position back (src/mas.jl:57-59)
        if (px != ox) x[i] = px;
        if (py != oy) y[i] = py;
        if (pz != oz) z[i] = pz;
        atomicAdd(oob + 1, 1ULL);
      }
      r[u] = make_float4(px, py, pz, MODE == BIN_SCATTER ? w[i] : __uint_as_float((unsigned)i));
      rank[u] = atomicAdd(&s_cnt[key[u]], 1u);
    }
  }
  __syncthreads();
  for (int t = threadIdx.x; t <= nbins; t += blockDim.x)
    if (s_cnt[t]) s_base[t] = atomicAdd(cursor + t, s_cnt[t]);
  __syncthreads();
#pragma unroll
  for (int u = 0; u < BIN_PPT; u++)
    if (key[u] >= 0) rec[s_base[key[u]] + rank[u]] = r[u];
}

template <int MAS>
__global__ void __launch_bounds__(256)
scatter_sorted_kernel(float* __restrict__ rho, const float4* __restrict__ rec, const unsigned* __restrict__ n_valid,
                      BoxGeom g, int wrap) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (int64_t)__ldg(n_valid)) return;  // records beyond this are the out-of-box trash bin
  float4 p = rec[i];
  deposit<MAS>(rho, p.x, p.y, p.z, p.w, g, wrap != 0);
}

// option "scatter_pairs" on the z-binned path (slab-decomposed scatters, unified_sort = 0): CIC with one quad per row
__global__ void __launch_bounds__(256)
scatter_sorted_cic_vec_kernel(float* __restrict__ rho, const float4* __restrict__ rec, const unsigned* __restrict__ n_valid,
                              BoxGeom g, int wrap) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (int64_t)__ldg(n_valid)) return;
  float4 p = rec[i];
  deposit_pairs<2>(rho, p.x, p.y, p.z, p.w, g, wrap != 0);
}

// option "scatter_pairs": the binned PCS scatter with vector reductions (deposit_pcs_vec in mas_math.cuh)
__global__ void __launch_bounds__(256)
scatter_sorted_pcs_vec_kernel(float* __restrict__ rho, const float4* __restrict__ rec, const unsigned* __restrict__ n_valid,
                              BoxGeom g, int wrap) {
  unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= *n_valid) return;
  float4 p = rec[i];
  deposit_pcs_vec(rho, p.x, p.y, p.z, p.w, g, wrap != 0);
}

// option "scatter_pairs": the binned TSC scatter with vector reductions (deposit_tsc_vec in mas_math.cuh)
__global__ void __launch_bounds__(256)
scatter_sorted_tsc_vec_kernel(float* __restrict__ rho, const float4* __restrict__ rec, const unsigned* __restrict__ n_valid,
                              BoxGeom g, int wrap) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (int64_t)__ldg(n_valid)) return;
  float4 p = rec[i];
  deposit_tsc_vec(rho, p.x, p.y, p.z, p.w, g, wrap != 0);
}

template <int NF, int MAS>
__global__ void __launch_bounds__(256)
gather_sorted_kernel(GatherArgs a, const float4* __restrict__ rec, const unsigned* __restrict__ n_valid, int64_t n,
                     BoxGeom g) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float4 p = rec[i];
  const int64_t idx = (int64_t)__float_as_uint(p.w);
  if (i >= (int64_t)__ldg(n_valid)) {  // out-of-box particle (trash bin): zero outputs
    for (int c = 0; c < NF; c++) a.o[c][idx] = 0.f;
    return;
  }
  gather_one<NF, MAS>(a, g, p.x, p.y, p.z, idx);
}

// zero-fill the outputs of out-of-box particles in the binned gather (they sit in the trash bin)
__global__ void gather_trash_kernel(GatherArgs a, const float4* __restrict__ rec, const unsigned* __restrict__ first,
                                    int64_t n_total, int nf) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_total || i < (int64_t)__ldg(first)) return;
  int64_t idx = (int64_t)__float_as_uint(rec[i].w);
  for (int c = 0; c < nf; c++) a.o[c][idx] = 0.f;
}


// ---- fine (tile) binning for the gather ---------------------------------------------------------
// Key = (z plane, y / TILE_Y, x / TILE_X) in that order.  Consecutive records then share the
// 2 x (TILE_Y+1) x (TILE_X+1) x 3-field window of their tile (~28 KB), which lives in L1 while
// a thread block works through it; successive tiles along x, y and z reuse the neighbouring rows
// and plane from L2.  Counting sort with one global counter per tile.
constexpr int TILE_X = 128;
constexpr int TILE_Y = 8;

struct TileGeom {
  int nxc, nyc;       // tiles per row / per plane column
  unsigned ntiles;    // nz * nyc * nxc  (the trash bin has index ntiles)
};

// Scatter flavour (option "scatter_tiles"): the same tile order for the deposit -- the 8 atomics of
// a warp's particles then fall into one 2 x 9 x 129-cell window instead of anywhere in a 2-plane
// slab.  Key from cic!'s own cell (src/mas.jl:13-35) of the wrapped position, which is returned.
__device__ __forceinline__ unsigned tile_key_scatter(float& px, float& py, float& pz, const BoxGeom& g,
                                                     const TileGeom& t, int wrap, bool& wrapped) {
  wrapped = false;
  if (wrap) {
    float qx = wrap_pos(px, g.mn[0], g.L[0]), qy = wrap_pos(py, g.mn[0], g.L[0]), qz = wrap_pos(pz, g.mn[0], g.L[0]);
    wrapped = (qx != px) || (qy != py) || (qz != pz);
    px = qx;
    py = qy;
    pz = qz;
  }
  int ix, iy, iz, i1;
  float w0, w1;
  bool ok = cic_axis(px, g.mn[0], g.L[0], g.n[0], wrap, ix, i1, w0, w1);
  ok = cic_axis(py, g.mn[1], g.L[1], g.n[1], wrap, iy, i1, w0, w1) && ok;
  ok = cic_axis(pz, g.mn[2], g.L[2], g.n[2], wrap, iz, i1, w0, w1) && ok;
  ok = ok && local_planes(g, iz, i1, iz, i1);
  if (!ok) return t.ntiles;
  return ((unsigned)iz * t.nyc + (unsigned)(iy / TILE_Y)) * t.nxc + (unsigned)(ix / TILE_X);
}

__global__ void __launch_bounds__(256)
tile_count_scatter_kernel(const float* __restrict__ x, const float* __restrict__ y, const float* __restrict__ z,
                          int64_t n, BoxGeom g, TileGeom t, int wrap, unsigned* __restrict__ counts) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float px = x[i], py = y[i], pz = z[i];
  bool wr;
  atomicAdd(counts + tile_key_scatter(px, py, pz, g, t, wrap, wr), 1u);
}

// records (x, y, z, w) in tile order; wrapped positions are written back like cic! does (src/mas.jl:57-59)
__global__ void __launch_bounds__(256)
tile_reorder_scatter_kernel(float* __restrict__ x, float* __restrict__ y, float* __restrict__ z,
                            const float* __restrict__ w, int64_t n, BoxGeom g, TileGeom t, int wrap,
                            unsigned* __restrict__ cursor, float4* __restrict__ rec, unsigned long long* oob) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float px = x[i], py = y[i], pz = z[i];
  const float ox = px, oy = py, oz = pz;
  bool wr;
  unsigned k = tile_key_scatter(px, py, pz, g, t, wrap, wr);
  if (wr) {
    if (px != ox) x[i] = px;
    if (py != oy) y[i] = py;
    if (pz != oz) z[i] = pz;
    atomicAdd(oob + 1, 1ULL);
  }
  unsigned dst = atomicAdd(cursor + k, 1u);
  rec[dst] = make_float4(px, py, pz, w[i]);
}

template <int MAS>
__device__ __forceinline__ unsigned tile_key(float px, float py, float pz, const BoxGeom& g, const TileGeom& t) {
  int ix, iy, iz;
  bool ok;
  if (MAS != BAOREC_MAS_CIC) {
    constexpr int SW = MasStencil<MAS>::SW;
    int idx[4];
    float w[4];
    ok = stencil_axis<MAS>(px, g.mn[0], g.L[0], g.n[0], true, idx, w);
    ix = idx[1];
    ok = stencil_axis<MAS>(py, g.mn[1], g.L[1], g.n[1], true, idx, w) && ok;
    iy = idx[1];
    ok = stencil_axis<MAS>(pz, g.mn[2], g.L[2], g.n[2], true, idx, w) && ok;
    iz = idx[1];
    if (ok && g.slab) {  // tiles are planes of the local buffer: plane 1 of the stencil, provided the whole stencil is local
      int lo, hi;
      ok = local_plane1(g, idx[0], lo) && local_plane1(g, idx[SW - 1], hi) && local_plane1(g, idx[1], iz);
    }
  } else {
    int iu;
    float wd, wu;
    ok = gather_axis(px, g.mn[0], g.L[0], g.cell[0], g.n[0], false, ix, iu, wd, wu);
    ok = gather_axis(py, g.mn[1], g.L[1], g.cell[1], g.n[1], false, iy, iu, wd, wu) && ok;
    ok = gather_axis(pz, g.mn[2], g.L[2], g.cell[2], g.n[2], false, iz, iu, wd, wu) && ok;
    ok = ok && local_planes(g, iz, iu, iz, iu);
  }
  if (!ok) return t.ntiles;
  return ((unsigned)iz * t.nyc + (unsigned)(iy / TILE_Y)) * t.nxc + (unsigned)(ix / TILE_X);
}

// ---- unified sort (run! sorts once, the read-back of the same catalog reuses it) -----------------
// Order-independent 64-bit content hash of a catalog: sum over particles of a mixed (index, x, y, z).
__device__ __forceinline__ unsigned long long mix64(unsigned long long v) {  // splitmix64 finaliser
  v ^= v >> 30;
  v *= 0xbf58476d1ce4e5b9ULL;
  v ^= v >> 27;
  v *= 0x94d049bb133111ebULL;
  v ^= v >> 31;
  return v;
}
__device__ __forceinline__ unsigned long long particle_hash(int64_t i, float px, float py, float pz) {
  unsigned long long a = (unsigned long long)__float_as_uint(px) | ((unsigned long long)__float_as_uint(py) << 32);
  unsigned long long b = (unsigned long long)__float_as_uint(pz) ^ ((unsigned long long)(i + 1) * 0x9e3779b97f4a7c15ULL);
  return mix64(a ^ mix64(b));
}
__device__ __forceinline__ void hash_accumulate(unsigned long long h, unsigned long long* __restrict__ dst) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) h += __shfl_down_sync(0xffffffffu, h, o);
  if ((threadIdx.x & 31) == 0) atomicAdd(dst, h);
}

__global__ void __launch_bounds__(256)
usort_count_kernel(const float* __restrict__ x, const float* __restrict__ y, const float* __restrict__ z, int64_t n,
                   BoxGeom g, TileGeom t, int wrap, unsigned* __restrict__ counts) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float px = x[i], py = y[i], pz = z[i];
  if (wrap) {  // cic!'s upper-face wrap (src/mas.jl:8-10); the key is the gather's tile of the wrapped position
    px = wrap_pos(px, g.mn[0], g.L[0]);
    py = wrap_pos(py, g.mn[0], g.L[0]);
    pz = wrap_pos(pz, g.mn[0], g.L[0]);
  }
  atomicAdd(counts + tile_key<BAOREC_MAS_CIC>(px, py, pz, g, t), 1u);
}

// records (x, y, z, w) in the gather's tile order + inverse permutation + content hash of the
// (wrapped, written-back) positions
// Four particles per thread, phase by phase (loads, keys, the four returning atomics back to back, stores): every
// thread keeps four of the atomics that return a record's slot in flight instead of one (ncu on the one-per-thread
// version: 55 cycles of long-scoreboard stall per issue at 85 % occupancy).  Measured at 1e8 particles: 3.89 -> 3.55 ms.
// What remains is traffic: the 16-byte records land in 10^6 open tiles, so most 32-byte sectors leave L2 half written
// and come back for their other half (ncu: 6.8 GB of DRAM traffic for 3.6 GB of algorithmic bytes).
constexpr int USORT_PPT = 4;
__global__ void __launch_bounds__(256)
usort_reorder_kernel(float* __restrict__ x, float* __restrict__ y, float* __restrict__ z, const float* __restrict__ w,
                     int64_t n, BoxGeom g, TileGeom t, int wrap, unsigned* __restrict__ cursor,
                     float4* __restrict__ rec, unsigned* __restrict__ inv, unsigned long long* __restrict__ oob,
                     unsigned long long* __restrict__ hash) {
  const int64_t base = (int64_t)blockIdx.x * (256 * USORT_PPT) + threadIdx.x;
  unsigned long long h = 0;
  float px[USORT_PPT], py[USORT_PPT], pz[USORT_PPT], pw[USORT_PPT];
  unsigned key[USORT_PPT], dst[USORT_PPT];
#pragma unroll
  for (int k = 0; k < USORT_PPT; k++) {
    const int64_t i = base + k * 256;
    const bool on = i < n;
    px[k] = on ? x[i] : 0.f;
    py[k] = on ? y[i] : 0.f;
    pz[k] = on ? z[i] : 0.f;
    pw[k] = on ? w[i] : 0.f;
  }
#pragma unroll
  for (int k = 0; k < USORT_PPT; k++) {
    const int64_t i = base + k * 256;
    if (i < n && wrap) {
      const float ox = px[k], oy = py[k], oz = pz[k];
      px[k] = wrap_pos(ox, g.mn[0], g.L[0]);
      py[k] = wrap_pos(oy, g.mn[0], g.L[0]);
      pz[k] = wrap_pos(oz, g.mn[0], g.L[0]);
      if (px[k] != ox) x[i] = px[k];  // write-back like the reference (src/mas.jl:57-59)
      if (py[k] != oy) y[i] = py[k];
      if (pz[k] != oz) z[i] = pz[k];
      if (px[k] != ox || py[k] != oy || pz[k] != oz) atomicAdd(oob + 1, 1ULL);
    }
    key[k] = tile_key<BAOREC_MAS_CIC>(px[k], py[k], pz[k], g, t);
  }
#pragma unroll
  for (int k = 0; k < USORT_PPT; k++)
    if (base + k * 256 < n) dst[k] = atomicAdd(cursor + key[k], 1u);
#pragma unroll
  for (int k = 0; k < USORT_PPT; k++) {
    const int64_t i = base + k * 256;
    if (i < n) {
      rec[dst[k]] = make_float4(px[k], py[k], pz[k], pw[k]);
      inv[i] = dst[k];
      h += particle_hash(i, px[k], py[k], pz[k]);
    }
  }
  hash_accumulate(h, hash);
}

__global__ void __launch_bounds__(256)
hash_positions_kernel(const float* __restrict__ x, const float* __restrict__ y, const float* __restrict__ z, int64_t n,
                      unsigned long long* __restrict__ hash) {
  unsigned long long h = 0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    h += particle_hash(i, x[i], y[i], z[i]);
  hash_accumulate(h, hash);
}
__global__ void hash_compare_kernel(unsigned long long* h) { h[2] = h[0] == h[1] ? 1ULL : 0ULL; }

// deposit every record (the order is the gather's; validity is cic!'s own: src/mas.jl:33-35)
__global__ void __launch_bounds__(256)
scatter_records_kernel(float* __restrict__ rho, const float4* __restrict__ rec, int64_t n, BoxGeom g, int wrap,
                       unsigned long long* __restrict__ oob) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float4 p = rec[i];
  if (!deposit<BAOREC_MAS_CIC>(rho, p.x, p.y, p.z, p.w, g, wrap != 0)) atomicAdd(oob, 1ULL);
}

// option "scatter_pairs": the tile-ordered scatter with vector reductions (deposit_pairs in mas_math.cuh)
template <int MODE>
__global__ void __launch_bounds__(256)
scatter_records_pairs_kernel(float* __restrict__ rho, const float4* __restrict__ rec, int64_t n, BoxGeom g, int wrap,
                             unsigned long long* __restrict__ oob) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float4 p = rec[i];
  if (!deposit_pairs<MODE>(rho, p.x, p.y, p.z, p.w, g, wrap != 0)) atomicAdd(oob, 1ULL);
}

// ---- deterministic scatter (option "deterministic_scatter"; see deposit_fixed in mas_math.cuh) -------------------
__global__ void __launch_bounds__(256)
scatter_records_fixed_kernel(unsigned long long* __restrict__ acc, const float4* __restrict__ rec, int64_t n, BoxGeom g,
                             int wrap, unsigned long long* __restrict__ oob) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float4 p = rec[i];
  if (!deposit_fixed(acc, p.x, p.y, p.z, p.w, g, wrap != 0)) atomicAdd(oob, 1ULL);
}

// catalog order (small catalogs, or unified_sort off); wraps and writes back like scatter_direct_kernel
__global__ void __launch_bounds__(256)
scatter_direct_fixed_kernel(unsigned long long* __restrict__ acc, float* __restrict__ x, float* __restrict__ y,
                            float* __restrict__ z, const float* __restrict__ w, int64_t n, BoxGeom g, int wrap,
                            unsigned long long* oob) {
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
    if (qx != px || qy != py || qz != pz) atomicAdd(oob + 1, 1ULL);
    px = qx;
    py = qy;
    pz = qz;
  }
  if (!deposit_fixed(acc, px, py, pz, w[i], g, wrap != 0)) atomicAdd(oob, 1ULL);
}

// rho += Float32(sum): the one rounding of the deterministic scatter (cic! accumulates into rho)
__global__ void __launch_bounds__(256)
fixed_to_float_kernel(float* __restrict__ rho, const unsigned long long* __restrict__ acc, size_t n) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) {
    const unsigned long long a = acc[i];
    if (a) rho[i] = __fadd_rn(rho[i], from_fixed(a));
  }
}

template <int MAS>
__global__ void __launch_bounds__(256)
tile_count_kernel(const float* __restrict__ x, const float* __restrict__ y, const float* __restrict__ z, int64_t n,
                  BoxGeom g, TileGeom t, unsigned* __restrict__ counts) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  atomicAdd(counts + tile_key<MAS>(x[i], y[i], z[i], g, t), 1u);
}

// exclusive scan of m counters, 3 phases (SCAN_CHUNK elements per block)
constexpr int SCAN_CHUNK = 2048;
__global__ void __launch_bounds__(256) scan_partial_kernel(const unsigned* __restrict__ in, unsigned* __restrict__ sums,
                                                           unsigned m) {
  __shared__ unsigned s[8];
  unsigned base = blockIdx.x * SCAN_CHUNK, acc = 0;
  for (unsigned i = threadIdx.x; i < SCAN_CHUNK; i += 256)
    if (base + i < m) acc += in[base + i];
  acc = __reduce_add_sync(0xffffffffu, acc);
  if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned t = 0;
    for (int w = 0; w < 8; w++) t += s[w];
    sums[blockIdx.x] = t;
  }
}
__global__ void scan_sums_kernel(unsigned* sums, unsigned nblocks, unsigned* total) {
  // single thread: nblocks is at most a few thousand
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    unsigned acc = 0;
    for (unsigned b = 0; b < nblocks; b++) {
      unsigned c = sums[b];
      sums[b] = acc;
      acc += c;
    }
    *total = acc;
  }
}
__global__ void __launch_bounds__(256) scan_final_kernel(const unsigned* __restrict__ in,
                                                         const unsigned* __restrict__ sums,
                                                         unsigned* __restrict__ out, unsigned* __restrict__ out2,
                                                         unsigned m) {
  // each thread scans 8 consecutive elements; block-level exclusive scan of the thread totals
  __shared__ unsigned s_warp[8];
  const unsigned base = blockIdx.x * SCAN_CHUNK + threadIdx.x * 8;
  unsigned v[8], tot = 0;
#pragma unroll
  for (int k = 0; k < 8; k++) {
    v[k] = (base + k < m) ? in[base + k] : 0u;
    tot += v[k];
  }
  unsigned incl = tot;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    unsigned up = __shfl_up_sync(0xffffffffu, incl, o);
    if ((threadIdx.x & 31) >= o) incl += up;
  }
  if ((threadIdx.x & 31) == 31) s_warp[threadIdx.x >> 5] = incl;
  __syncthreads();
  unsigned woff = 0;
  for (int w = 0; w < (int)(threadIdx.x >> 5); w++) woff += s_warp[w];
  unsigned run = sums[blockIdx.x] + woff + incl - tot;
#pragma unroll
  for (int k = 0; k < 8; k++) {
    if (base + k < m) {
      out[base + k] = run;
      out2[base + k] = run;
    }
    run += v[k];
  }
}

template <int MAS>
__global__ void __launch_bounds__(256)
tile_reorder_kernel(const float* __restrict__ x, const float* __restrict__ y, const float* __restrict__ z, int64_t n,
                    BoxGeom g, TileGeom t, unsigned* __restrict__ cursor, float4* __restrict__ rec,
                    unsigned* __restrict__ inv) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float px = x[i], py = y[i], pz = z[i];
  unsigned k = tile_key<MAS>(px, py, pz, g, t);
  unsigned dst = atomicAdd(cursor + k, 1u);
  rec[dst] = make_float4(px, py, pz, __uint_as_float((unsigned)i));
  inv[i] = dst;  // coalesced: where particle i went
}


// Tile gather: one thread block per (z, y/TILE_Y, x/TILE_X) tile.  The 2 x (TILE_Y+1) x (TILE_X+1)
// window of each field is staged in shared memory with coalesced row loads (one 128 B line per
// warp instruction instead of one line per lane), then every particle of the tile reads its 8 x NF
// values from shared memory.  Same arithmetic and order as gather_one -> identical results.
constexpr int TILE_WXP = TILE_X + 4;  // padded row length in shared memory

// The particle loop of the TMA-staged tile gather (gather_tile_tma_kernel below): every particle of the tile reads its
// 8 x NF values from the staged window (PLANE floats per (field, plane), rows of TILE_WXP).  The same arithmetic in the
// same order as the loop of gather_tile_kernel and as gather_one -> identical results.
template <int NF, int PLANE, bool FIRST>
__device__ __forceinline__ void tile_particles(const float* __restrict__ sm, const GatherArgs& a,
                                               const float4* __restrict__ rec, unsigned beg, unsigned end,
                                               float4 p_first, const BoxGeom& g, int x0, int y0) {
  const int nx = g.n[0], ny = g.n[1], nz = g.n[2];
  for (unsigned i = beg + threadIdx.x; i < end; i += 128) {
    float4 p = (FIRST && i < beg + 128) ? p_first : rec[i];
    const float px = p.x, py = p.y, pz = p.z;
    const int64_t out_idx = (int64_t)__float_as_uint(p.w);
    int xd, xu, yd, yu, zd, zu;
    float dx, ux, dy, uy, dz, uz;
    gather_axis(px, g.mn[0], g.L[0], g.cell[0], nx, false, xd, xu, dx, ux);
    gather_axis(py, g.mn[1], g.L[1], g.cell[1], ny, false, yd, yu, dy, uy);
    gather_axis(pz, g.mn[2], g.L[2], g.cell[2], nz, false, zd, zu, dz, uz);
    const float* w00 = sm + (yd - y0) * TILE_WXP + (xd - x0);  // plane 0, row ly, column lx
    float val[NF];
#pragma unroll
    for (int c = 0; c < NF; c++) {
      const float* q0 = w00 + (2 * c) * PLANE;
      const float* q1 = q0 + PLANE;
      float v;
      v = __fmul_rn(__fmul_rn(__fmul_rn(q0[0], dx), dy), dz);
      v = __fadd_rn(v, __fmul_rn(__fmul_rn(__fmul_rn(q1[0], dx), dy), uz));
      v = __fadd_rn(v, __fmul_rn(__fmul_rn(__fmul_rn(q0[TILE_WXP], dx), uy), dz));
      v = __fadd_rn(v, __fmul_rn(__fmul_rn(__fmul_rn(q1[TILE_WXP], dx), uy), uz));
      v = __fadd_rn(v, __fmul_rn(__fmul_rn(__fmul_rn(q0[1], ux), dy), dz));
      v = __fadd_rn(v, __fmul_rn(__fmul_rn(__fmul_rn(q1[1], ux), dy), uz));
      v = __fadd_rn(v, __fmul_rn(__fmul_rn(__fmul_rn(q0[TILE_WXP + 1], ux), uy), dz));
      v = __fadd_rn(v, __fmul_rn(__fmul_rn(__fmul_rn(q1[TILE_WXP + 1], ux), uy), uz));
      val[c] = v;
    }
    shifts_epilogue<NF>(a, val, px, py, pz, out_idx, (int64_t)i);
  }
}

// STAGE selects how the fast path issues its cp.async rows: 0 = rows dealt round-robin over a fully unrolled loop
// (the version measured in round 1: 1719 of its 2700 SASS instructions are integer address arithmetic, and every warp
// steps over all 54 row predicates); 1 = each warp owns rows py = warp, warp + 4 (, 8) of every (field, plane):
// one plane base pointer per (field, plane), one multiply-add per row, compile-time shared-memory offsets; the tile
// index is decoded with shifts when the tile counts are powers of two; a thread's first record is requested before
// the window is staged
// (option "gather_stage", off until it has been measured).  Identical shared-memory contents either way.
template <int NF, int STAGE>
__global__ void __launch_bounds__(128)
gather_tile_kernel(GatherArgs a, const float4* __restrict__ rec, const unsigned* __restrict__ starts, BoxGeom g,
                   TileGeom t) {
  __shared__ __align__(16) float sm[NF][2][TILE_Y + 1][TILE_WXP];
  const unsigned tile = blockIdx.x;
  const unsigned beg = starts[tile], end = starts[tile + 1];
  if (beg == end) return;
  const int nx = g.n[0], ny = g.n[1], nz = g.n[2];
  int tx, ty, iz;
  if (STAGE == 1 && ((t.nxc & (t.nxc - 1)) | (t.nyc & (t.nyc - 1))) == 0) {  // power-of-two tile counts: shifts, no division
    const int sx = __ffs(t.nxc) - 1, sy = __ffs(t.nyc) - 1;
    tx = (int)(tile & (unsigned)(t.nxc - 1));
    ty = (int)((tile >> sx) & (unsigned)(t.nyc - 1));
    iz = (int)(tile >> (sx + sy));
  } else {
    tx = tile % t.nxc;
    const unsigned r = tile / t.nxc;
    ty = r % t.nyc;
    iz = r / t.nyc;
  }
  const int x0 = tx * TILE_X, y0 = ty * TILE_Y;
  const int wx = min(TILE_X, nx - x0), wy = min(TILE_Y, ny - y0);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const bool fast = (wx == TILE_X) && (wy == TILE_Y) && (nx % 4 == 0) &&
                    ((((uintptr_t)a.f[0] | (uintptr_t)a.f[NF > 1 ? 1 : 0] | (uintptr_t)a.f[NF > 2 ? 2 : 0]) & 15) == 0);
  float4 p_first = make_float4(0.f, 0.f, 0.f, 0.f);
  if (STAGE == 1 && beg + threadIdx.x < end) p_first = rec[beg + threadIdx.x];  // in flight while the window is staged
  if (fast && STAGE == 1) {
    int xe = x0 + TILE_X;
    if (xe >= nx) xe -= nx;
    const int z1 = g.slab ? iz + 1 : (iz + 1 >= nz ? 0 : iz + 1);
    int y8 = y0 + TILE_Y;  // only the last row of the window can wrap: the tile is full, so y0 + TILE_Y - 1 < ny
    if (y8 >= ny) y8 -= ny;
    const size_t plane_elems = (size_t)ny * nx;
    const unsigned s_lane = (unsigned)__cvta_generic_to_shared(&sm[0][0][0][0]) + 16u * lane;
    const unsigned s_edge = (unsigned)__cvta_generic_to_shared(&sm[0][0][0][0]) + 4u * TILE_X;
#pragma unroll
    for (int p = 0; p < NF * 2; p++) {  // p = field * 2 + plane
      const int f = p >> 1;
      const float* fld = f == 0 ? a.f[0] : (f == 1 ? a.f[NF > 1 ? 1 : 0] : a.f[NF > 2 ? 2 : 0]);
      const float* base = fld + (size_t)((p & 1) ? z1 : iz) * plane_elems + x0;
#pragma unroll
      for (int k = 0; k < 3; k++) {
        const int py = warp + 4 * k;  // rows warp, warp + 4 and -- warp 0 only -- TILE_Y
        if (py <= TILE_Y) {
          const int yy = py == TILE_Y ? y8 : y0 + py;
          const float* src = base + (size_t)yy * nx;
          const unsigned row = (unsigned)((p * (TILE_Y + 1) + py) * TILE_WXP) * 4u;
          asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s_lane + row), "l"(src + 4 * lane));
          if (lane == 0)
            asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(s_edge + row), "l"(src - x0 + xe));
        }
      }
    }
    asm volatile("cp.async.wait_all;" ::: "memory");
  } else if (fast) {
    // One warp per row, rows dealt round-robin to the 4 warps with compile-time indices (no
    // integer division); each lane issues one 16-byte cp.async (LDGSTS) per row straight into
    // shared memory -- no register staging, all of a warp's rows in flight at once.
    int xe = x0 + TILE_X;
    if (xe >= nx) xe -= nx;
    int zrow[2];
    zrow[0] = iz;
    zrow[1] = g.slab ? iz + 1 : (iz + 1 >= nz ? 0 : iz + 1);
#pragma unroll
    for (int r = 0; r < NF * 2 * (TILE_Y + 1); r++) {
      if ((r & 3) != warp) continue;
      constexpr int RY = TILE_Y + 1;
      const int py = r % RY, pz = (r / RY) & 1, f = r / (2 * RY);
      int yy = y0 + py;
      if (yy >= ny) yy -= ny;
      const float* fld = f == 0 ? a.f[0] : (f == 1 ? a.f[NF > 1 ? 1 : 0] : a.f[NF > 2 ? 2 : 0]);
      const float* src = fld + ((size_t)zrow[pz] * ny + yy) * nx;
      float* dst = &sm[f][pz][py][0];
      unsigned d16 = (unsigned)__cvta_generic_to_shared(dst + 4 * lane);
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d16), "l"(src + x0 + 4 * lane));
      if (lane == 0) {
        unsigned d4 = (unsigned)__cvta_generic_to_shared(dst + TILE_X);
        asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(d4), "l"(src + xe));
      }
    }
    asm volatile("cp.async.wait_all;" ::: "memory");
  } else {
    const int nrows = NF * 2 * (wy + 1);
    for (int row = warp; row < nrows; row += 4) {
      int py = row % (wy + 1);
      int q = row / (wy + 1);
      int pz = q & 1, f = q >> 1;
      int zz = iz + pz;
      if (!g.slab && zz >= nz) zz -= nz;
      int yy = y0 + py;
      if (yy >= ny) yy -= ny;
      const float* fld = f == 0 ? a.f[0] : (f == 1 ? a.f[NF > 1 ? 1 : 0] : a.f[NF > 2 ? 2 : 0]);
      const float* src = fld + ((size_t)zz * ny + yy) * nx;
      float* dst = &sm[f][pz][py][0];
      for (int c = lane; c <= wx; c += 32) {
        int xx = x0 + c;
        if (xx >= nx) xx -= nx;
        dst[c] = __ldg(src + xx);
      }
    }
  }
  __syncthreads();
  for (unsigned i = beg + threadIdx.x; i < end; i += 128) {
    float4 p = (STAGE == 1 && i < beg + 128) ? p_first : rec[i];
    const float px = p.x, py = p.y, pz = p.z;
    const int64_t out_idx = (int64_t)__float_as_uint(p.w);
    int xd, xu, yd, yu, zd, zu;
    float dx, ux, dy, uy, dz, uz;
    gather_axis(px, g.mn[0], g.L[0], g.cell[0], nx, false, xd, xu, dx, ux);
    gather_axis(py, g.mn[1], g.L[1], g.cell[1], ny, false, yd, yu, dy, uy);
    gather_axis(pz, g.mn[2], g.L[2], g.cell[2], nz, false, zd, zu, dz, uz);
    const int lx = xd - x0, ly = yd - y0;
    float val[NF];
#pragma unroll
    for (int c = 0; c < NF; c++) {
      float v;
      v = __fmul_rn(__fmul_rn(__fmul_rn(sm[c][0][ly][lx], dx), dy), dz);
      v = __fadd_rn(v, __fmul_rn(__fmul_rn(__fmul_rn(sm[c][1][ly][lx], dx), dy), uz));
      v = __fadd_rn(v, __fmul_rn(__fmul_rn(__fmul_rn(sm[c][0][ly + 1][lx], dx), uy), dz));
      v = __fadd_rn(v, __fmul_rn(__fmul_rn(__fmul_rn(sm[c][1][ly + 1][lx], dx), uy), uz));
      v = __fadd_rn(v, __fmul_rn(__fmul_rn(__fmul_rn(sm[c][0][ly][lx + 1], ux), dy), dz));
      v = __fadd_rn(v, __fmul_rn(__fmul_rn(__fmul_rn(sm[c][1][ly][lx + 1], ux), dy), uz));
      v = __fadd_rn(v, __fmul_rn(__fmul_rn(__fmul_rn(sm[c][0][ly + 1][lx + 1], ux), uy), dz));
      v = __fadd_rn(v, __fmul_rn(__fmul_rn(__fmul_rn(sm[c][1][ly + 1][lx + 1], ux), uy), uz));
      val[c] = v;
    }
    shifts_epilogue<NF>(a, val, px, py, pz, out_idx, (int64_t)i);
  }
}

// ---- tile gather staged by the TMA engine (option "gather_stage" = 2) ------------------------------------------------
// One 3-D tensor map per field (x fastest, then y, then the nzp planes of the buffer), box = TILE_WXP x (TILE_Y+1) x 1:
// ONE thread issues 2 x NF cp.async.bulk.tensor.3d copies (UTMALDG) for the whole window -- the 54 row copies of the
// cp.async versions with their address arithmetic leave the issue slots -- and the block waits on one mbarrier.
// The engine fills what lies beyond the mesh with zeros, it does not wrap: the last tile along x takes its column
// TILE_X from x = 0 and the last tile along y its row TILE_Y from y = 0 with ordinary loads, requested before the wait
// and stored after it.  The two planes are separate copies, so the periodic z + 1 is just a coordinate.
// Shared-memory destinations of tensor copies must be 128-byte aligned: each (field, plane) block is padded from
// 9 x 132 = 1188 floats to 1216.  Tiles that are not full (mesh sizes off the tile grid) stage with plain loads.
struct alignas(64) GatherMaps {
  CUtensorMap m[3];
};
constexpr int TMA_PLANE = 1216;
static_assert(TMA_PLANE >= (TILE_Y + 1) * TILE_WXP && (TMA_PLANE * 4) % 128 == 0, "128-byte aligned (field, plane) blocks");
static_assert((TILE_WXP * 4) % 16 == 0, "the box's inner extent is a multiple of 16 bytes");

template <int NF>
__global__ void __launch_bounds__(128)
gather_tile_tma_kernel(const __grid_constant__ GatherMaps maps, GatherArgs a, const float4* __restrict__ rec,
                       const unsigned* __restrict__ starts, BoxGeom g, TileGeom t) {
  __shared__ __align__(128) float sm[NF * 2 * TMA_PLANE];
  __shared__ __align__(8) unsigned long long mbar_store;
  const unsigned tile = blockIdx.x;
  const unsigned beg = starts[tile], end = starts[tile + 1];
  if (beg == end) return;
  const int nx = g.n[0], ny = g.n[1], nz = g.n[2];
  int tx, ty, iz;
  if (((t.nxc & (t.nxc - 1)) | (t.nyc & (t.nyc - 1))) == 0) {
    const int sx = __ffs(t.nxc) - 1, sy = __ffs(t.nyc) - 1;
    tx = (int)(tile & (unsigned)(t.nxc - 1));
    ty = (int)((tile >> sx) & (unsigned)(t.nyc - 1));
    iz = (int)(tile >> (sx + sy));
  } else {
    tx = tile % t.nxc;
    const unsigned r = tile / t.nxc;
    ty = r % t.nyc;
    iz = r / t.nyc;
  }
  const int x0 = tx * TILE_X, y0 = ty * TILE_Y;
  const int wx = min(TILE_X, nx - x0), wy = min(TILE_Y, ny - y0);
  const int z1 = g.slab ? iz + 1 : (iz + 1 >= nz ? 0 : iz + 1);
  const int tid = threadIdx.x;
  float4 p_first = make_float4(0.f, 0.f, 0.f, 0.f);
  if (beg + tid < end) p_first = rec[beg + tid];  // in flight while the window is staged
  if (wx == TILE_X && wy == TILE_Y) {
    const unsigned mbar = (unsigned)__cvta_generic_to_shared(&mbar_store);
    const unsigned s0 = (unsigned)__cvta_generic_to_shared(sm);
    if (tid == 0) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(mbar) : "memory");
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
      constexpr unsigned BYTES = NF * 2 * (TILE_Y + 1) * TILE_WXP * 4;
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mbar), "r"(BYTES) : "memory");
#pragma unroll
      for (int p = 0; p < NF * 2; p++) {  // p = field * 2 + plane
        const unsigned long long map = (unsigned long long)(uintptr_t)&maps.m[p >> 1];
        asm volatile(
            "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
            ::"r"(s0 + (unsigned)(p * TMA_PLANE * 4)), "l"(map), "r"(x0), "r"(y0), "r"((p & 1) ? z1 : iz), "r"(mbar)
            : "memory");
      }
    }
    // what the engine cannot fetch: the periodic images of column TILE_X (last tile along x) and row TILE_Y (last along y)
    const bool wrapx = x0 + TILE_X >= nx, wrapy = y0 + TILE_Y >= ny;
    const size_t plane_elems = (size_t)ny * nx;
    float ex = 0.f;
    int ex_at = -1;
    if (wrapx && tid < NF * 2 * (TILE_Y + 1)) {
      const int p = tid / (TILE_Y + 1), py = tid - p * (TILE_Y + 1);
      int yy = y0 + py;
      if (yy >= ny) yy -= ny;
      const int f = p >> 1;
      const float* fld = f == 0 ? a.f[0] : (f == 1 ? a.f[NF > 1 ? 1 : 0] : a.f[NF > 2 ? 2 : 0]);
      ex = __ldg(fld + (size_t)((p & 1) ? z1 : iz) * plane_elems + (size_t)yy * nx);
      ex_at = p * TMA_PLANE + py * TILE_WXP + TILE_X;
    }
    __syncthreads();  // the barrier is initialised before anyone polls it
    unsigned done = 0;
    while (!done) {
      asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n selp.u32 %0, 1, 0, p;\n}"
                   : "=r"(done)
                   : "r"(mbar)
                   : "memory");
    }
    if (wrapx | wrapy) {
      if (ex_at >= 0) sm[ex_at] = ex;
      if (wrapy) {
        for (int idx = tid; idx < NF * 2 * (TILE_X + 1); idx += 128) {
          const int p = idx / (TILE_X + 1), c = idx - p * (TILE_X + 1);
          int xx = x0 + c;
          if (xx >= nx) xx -= nx;
          const int f = p >> 1;
          const float* fld = f == 0 ? a.f[0] : (f == 1 ? a.f[NF > 1 ? 1 : 0] : a.f[NF > 2 ? 2 : 0]);
          sm[p * TMA_PLANE + TILE_Y * TILE_WXP + c] = __ldg(fld + (size_t)((p & 1) ? z1 : iz) * plane_elems + xx);
        }
      }
      __syncthreads();
    }
  } else {
    const int lane = tid & 31, warp = tid >> 5;
    const int nrows = NF * 2 * (wy + 1);
    for (int row = warp; row < nrows; row += 4) {
      const int py = row % (wy + 1), p = row / (wy + 1);
      int yy = y0 + py;
      if (yy >= ny) yy -= ny;
      const int f = p >> 1;
      const float* fld = f == 0 ? a.f[0] : (f == 1 ? a.f[NF > 1 ? 1 : 0] : a.f[NF > 2 ? 2 : 0]);
      const float* src = fld + ((size_t)((p & 1) ? z1 : iz) * ny + yy) * nx;
      float* dst = sm + p * TMA_PLANE + py * TILE_WXP;
      for (int c = lane; c <= wx; c += 32) {
        int xx = x0 + c;
        if (xx >= nx) xx -= nx;
        dst[c] = __ldg(src + xx);
      }
    }
    __syncthreads();
  }
  tile_particles<NF, TMA_PLANE, true>(sm, a, rec, beg, end, p_first, g, x0, y0);
}

// cuTensorMapEncodeTiled through the runtime's driver entry point query: no link-time dependency on libcuda
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn encode_tiled_fn() {
  static EncodeTiledFn fn = [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q = cudaDriverEntryPointSymbolNotFound;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
      p = nullptr;
    return (EncodeTiledFn)p;
  }();
  return fn;
}

// The maps of the gather's fields; false when the layout cannot be described (row pitch or base off 16 bytes) or the
// driver has no tensor maps: the caller stages with cp.async then.
static bool gather_maps(const GatherArgs& a, int nf, const BoxGeom& g, GatherMaps* out) {
  EncodeTiledFn enc = encode_tiled_fn();
  if (!enc || g.n[0] % 4 != 0) return false;
  for (int f = 0; f < 3; f++) {
    const float* base = a.f[f < nf ? f : 0];
    if (((uintptr_t)base & 15) != 0) return false;
    const cuuint64_t dims[3] = {(cuuint64_t)g.n[0], (cuuint64_t)g.n[1], (cuuint64_t)g.nzp};
    const cuuint64_t strides[2] = {(cuuint64_t)g.n[0] * 4, (cuuint64_t)g.n[0] * g.n[1] * 4};
    const cuuint32_t box[3] = {TILE_WXP, TILE_Y + 1, 1};
    const cuuint32_t estr[3] = {1, 1, 1};
    if (enc(&out->m[f], CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, (void*)base, dims, strides, box, estr,
            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,
            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
      return false;
  }
  return true;
}

__global__ void add_oob_kernel(const unsigned* __restrict__ count, unsigned long long* oob) {
  if (*count) atomicAdd(oob, (unsigned long long)*count);
}

// out[c][i] = sorted_out[inv[i]].c : coalesced index read and output writes, one random 16 B read.
// One particle per thread.  (Four per thread, the four random reads requested before the first use, measured SLOWER on
// B200: 2.05 instead of 1.96 ms at 1e8 particles -- the kernel sits at the DRAM's random 64-byte access rate, not on
// the latency of one request.)
__global__ void __launch_bounds__(256)
unsort_kernel(GatherArgs a, const float4* __restrict__ sorted_out, const unsigned* __restrict__ inv, int64_t n,
              const unsigned* __restrict__ n_valid) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  unsigned j = inv[i];
  float4 v = j < __ldg(n_valid) ? __ldg(sorted_out + j) : make_float4(0.f, 0.f, 0.f, 0.f);
  a.o[0][i] = v.x;
  a.o[1][i] = v.y;
  a.o[2][i] = v.z;
}

struct BinResult {
  float4* rec;
  const unsigned* n_valid;  // device: number of in-box records (start of the trash bin)
  const unsigned* starts = nullptr;  // tile binning: first record of every tile (+ trash, + end)
  unsigned ntiles = 0;
  unsigned* inv = nullptr;           // tile binning: sorted position of every particle
};

// Builds the sorted records (asynchronous: the number of in-box records stays on the device).
template <int MODE>
static int bin_particles(baorec_ctx* ctx, float* x, float* y, float* z, const float* w, int64_t n, int wrap, int mas,
                         cudaStream_t st, BinResult* out) {
  BoxGeom g = geom_of(ctx);
  int zg = MODE == BIN_SCATTER ? ctx->opt_zg_scatter : ctx->opt_zg_gather;
  const int nzp = g.nzp;
  if (zg <= 0) zg = MODE == BIN_SCATTER ? (ctx->nz + 511) / 512 : 1;  // auto: window << L2
  if (zg < 1) zg = 1;
  int nbins = (nzp + zg - 1) / zg;
  if (nbins > BIN_MAX) {
    zg = (nzp + BIN_MAX - 1) / BIN_MAX;
    nbins = (nzp + zg - 1) / zg;
  }
  unsigned* cnt;
  float4* rec;
  BR_TRY(need_t(ctx, BUF_BINKEY, (size_t)3 * (BIN_MAX + 2), &cnt));
  BR_TRY(need_t(ctx, BUF_BINIDX, (size_t)n, &rec));
  unsigned* cursor = cnt + (BIN_MAX + 2);
  unsigned* starts = cnt + 2 * (BIN_MAX + 2);
  BR_CUDA(cudaMemsetAsync(cnt, 0, sizeof(unsigned) * (BIN_MAX + 2), st));
  unsigned grid_c = cdiv((size_t)n, BIN_THREADS * 4);
  if (grid_c > 148 * 8) grid_c = 148 * 8;
  unsigned grid_r = cdiv((size_t)n, BIN_THREADS * BIN_PPT);
  const bool tsc = mas != BAOREC_MAS_CIC, pcs = mas == BAOREC_MAS_PCS;  // tsc: a stencil scheme (TSC or PCS)
  if (pcs) {
    BR_LAUNCH(ctx, (bin_count_kernel<MODE, BAOREC_MAS_PCS>), grid_c, BIN_THREADS, 0, st, x, y, z, n, g, wrap, zg, nbins,
              cnt);
  } else if (tsc) {
    BR_LAUNCH(ctx, (bin_count_kernel<MODE, BAOREC_MAS_TSC>), grid_c, BIN_THREADS, 0, st, x, y, z, n, g, wrap, zg, nbins,
              cnt);
  } else {
    BR_LAUNCH(ctx, (bin_count_kernel<MODE, BAOREC_MAS_CIC>), grid_c, BIN_THREADS, 0, st, x, y, z, n, g, wrap, zg, nbins,
              cnt);
  }
  BR_LAUNCH(ctx, bin_scan_kernel, 1, 256, 0, st, cnt, cursor, starts, nbins, ctx->d_oob);
  if (pcs) {
    BR_LAUNCH(ctx, (bin_reorder_kernel<MODE, BAOREC_MAS_PCS>), grid_r, BIN_THREADS, 0, st, x, y, z, w, n, g, wrap, zg,
              nbins, cursor, rec, ctx->d_oob);
  } else if (tsc) {
    BR_LAUNCH(ctx, (bin_reorder_kernel<MODE, BAOREC_MAS_TSC>), grid_r, BIN_THREADS, 0, st, x, y, z, w, n, g, wrap, zg,
              nbins, cursor, rec, ctx->d_oob);
  } else {
    BR_LAUNCH(ctx, (bin_reorder_kernel<MODE, BAOREC_MAS_CIC>), grid_r, BIN_THREADS, 0, st, x, y, z, w, n, g, wrap, zg,
              nbins, cursor, rec, ctx->d_oob);
  }
  out->rec = rec;
  out->n_valid = starts + nbins;  // stays on the device: no host round trip
  return BAOREC_OK;
}

// Fine binning for the gather (see TILE_X / TILE_Y).  Asynchronous.
static int bin_tiles(baorec_ctx* ctx, const float* x, const float* y, const float* z, int64_t n, int mas,
                     cudaStream_t st, BinResult* out) {
  BoxGeom g = geom_of(ctx);
  TileGeom t;
  t.nxc = (ctx->nx + TILE_X - 1) / TILE_X;
  t.nyc = (ctx->ny + TILE_Y - 1) / TILE_Y;
  t.ntiles = (unsigned)g.nzp * t.nyc * t.nxc;
  const unsigned m = t.ntiles + 1;  // + trash bin
  const unsigned nsb = cdiv(m, SCAN_CHUNK);
  unsigned* cnt;
  float4* rec;
  BR_TRY(need_t(ctx, BUF_BINTMP, (size_t)3 * m + nsb + 16, &cnt));
  BR_TRY(need_t(ctx, BUF_BINIDX, (size_t)n, &rec));
  unsigned* inv;
  BR_TRY(need_t(ctx, BUF_BININV, (size_t)n, &inv));
  unsigned* cursor = cnt + m;
  unsigned* starts = cursor + m;  // m + 1 entries: starts[m] = n
  unsigned* sums = starts + m + 1;
  unsigned* total = sums + nsb;
  BR_CUDA(cudaMemsetAsync(cnt, 0, sizeof(unsigned) * m, st));
  const unsigned grid = cdiv((size_t)n, 256);
  const bool tsc = mas != BAOREC_MAS_CIC, pcs = mas == BAOREC_MAS_PCS;  // tsc: a stencil scheme (TSC or PCS)
  if (pcs) BR_LAUNCH(ctx, tile_count_kernel<BAOREC_MAS_PCS>, grid, 256, 0, st, x, y, z, n, g, t, cnt);
  else if (tsc) BR_LAUNCH(ctx, tile_count_kernel<BAOREC_MAS_TSC>, grid, 256, 0, st, x, y, z, n, g, t, cnt);
  else BR_LAUNCH(ctx, tile_count_kernel<BAOREC_MAS_CIC>, grid, 256, 0, st, x, y, z, n, g, t, cnt);
  BR_LAUNCH(ctx, scan_partial_kernel, nsb, 256, 0, st, cnt, sums, m);
  BR_LAUNCH(ctx, scan_sums_kernel, 1, 32, 0, st, sums, nsb, total);
  BR_LAUNCH(ctx, scan_final_kernel, nsb, 256, 0, st, cnt, sums, cursor, starts, m);
  BR_LAUNCH(ctx, add_oob_kernel, 1, 1, 0, st, cnt + t.ntiles, ctx->d_oob);  // trash-bin size -> out-of-box counter
  if (pcs) BR_LAUNCH(ctx, tile_reorder_kernel<BAOREC_MAS_PCS>, grid, 256, 0, st, x, y, z, n, g, t, cursor, rec, inv);
  else if (tsc) BR_LAUNCH(ctx, tile_reorder_kernel<BAOREC_MAS_TSC>, grid, 256, 0, st, x, y, z, n, g, t, cursor, rec, inv);
  else BR_LAUNCH(ctx, tile_reorder_kernel<BAOREC_MAS_CIC>, grid, 256, 0, st, x, y, z, n, g, t, cursor, rec, inv);
  out->rec = rec;
  out->n_valid = starts + t.ntiles;
  out->starts = starts;
  out->ntiles = t.ntiles;
  out->inv = inv;
  return BAOREC_OK;
}

// Tile order for the scatter (CIC).  Asynchronous.
static int bin_tiles_scatter(baorec_ctx* ctx, float* x, float* y, float* z, const float* w, int64_t n, int wrap,
                             cudaStream_t st, BinResult* out) {
  BoxGeom g = geom_of(ctx);
  TileGeom t;
  t.nxc = (ctx->nx + TILE_X - 1) / TILE_X;
  t.nyc = (ctx->ny + TILE_Y - 1) / TILE_Y;
  t.ntiles = (unsigned)g.nzp * t.nyc * t.nxc;
  const unsigned m = t.ntiles + 1;  // + trash bin
  const unsigned nsb = cdiv(m, SCAN_CHUNK);
  unsigned* cnt;
  float4* rec;
  BR_TRY(need_t(ctx, BUF_BINTMP, (size_t)3 * m + nsb + 16, &cnt));
  BR_TRY(need_t(ctx, BUF_BINIDX, (size_t)n, &rec));
  unsigned* cursor = cnt + m;
  unsigned* starts = cursor + m;
  unsigned* sums = starts + m + 1;
  unsigned* total = sums + nsb;
  BR_CUDA(cudaMemsetAsync(cnt, 0, sizeof(unsigned) * m, st));
  const unsigned grid = cdiv((size_t)n, 256);
  BR_LAUNCH(ctx, tile_count_scatter_kernel, grid, 256, 0, st, x, y, z, n, g, t, wrap, cnt);
  BR_LAUNCH(ctx, scan_partial_kernel, nsb, 256, 0, st, cnt, sums, m);
  BR_LAUNCH(ctx, scan_sums_kernel, 1, 32, 0, st, sums, nsb, total);
  BR_LAUNCH(ctx, scan_final_kernel, nsb, 256, 0, st, cnt, sums, cursor, starts, m);
  BR_LAUNCH(ctx, add_oob_kernel, 1, 1, 0, st, cnt + t.ntiles, ctx->d_oob);
  BR_LAUNCH(ctx, tile_reorder_scatter_kernel, grid, 256, 0, st, x, y, z, w, n, g, t, wrap, cursor, rec, ctx->d_oob);
  out->rec = rec;
  out->n_valid = starts + t.ntiles;
  return BAOREC_OK;
}

struct TileScratch {
  unsigned *cnt, *cursor, *starts, *sums, *total;
  unsigned m, nsb;
};
static int tile_scratch(baorec_ctx* ctx, const TileGeom& t, TileScratch* s) {
  s->m = t.ntiles + 1;  // + trash bin
  s->nsb = cdiv(s->m, SCAN_CHUNK);
  BR_TRY(need_t(ctx, BUF_BINTMP, (size_t)3 * s->m + s->nsb + 16, &s->cnt));
  s->cursor = s->cnt + s->m;
  s->starts = s->cursor + s->m;  // m + 1 entries: starts[m] = n
  s->sums = s->starts + s->m + 1;
  s->total = s->sums + s->nsb;
  return BAOREC_OK;
}

// run!'s sort: the catalog into the gather's tile order, once (see opt_unified_sort).  Asynchronous.
static int unified_sort(baorec_ctx* ctx, float* x, float* y, float* z, const float* w, int64_t n, int wrap,
                        cudaStream_t st, BinResult* out) {
  ctx->sortc_valid = false;
  // the key is the GATHER's tile: on a slab that is the gather's halo layout (mode 2: one plane below, two above the
  // slab), whatever layout the scatter that called us deposits into (mode 1: one ghost plane above)
  const int key_mode = ctx->slab_mode ? 2 : 0;
  BoxGeom g = geom_for(ctx, key_mode);
  TileGeom t;
  t.nxc = (ctx->nx + TILE_X - 1) / TILE_X;
  t.nyc = (ctx->ny + TILE_Y - 1) / TILE_Y;
  t.ntiles = (unsigned)g.nzp * t.nyc * t.nxc;
  TileScratch sc;
  BR_TRY(tile_scratch(ctx, t, &sc));
  float4* rec;
  unsigned* inv;
  BR_TRY(need_t(ctx, BUF_BINIDX, (size_t)n, &rec));
  BR_TRY(need_t(ctx, BUF_BININV, (size_t)n, &inv));
  BR_CUDA(cudaMemsetAsync(sc.cnt, 0, sizeof(unsigned) * sc.m, st));
  BR_CUDA(cudaMemsetAsync(ctx->d_hash, 0, sizeof(unsigned long long), st));
  const unsigned grid = cdiv((size_t)n, 256);
  BR_LAUNCH(ctx, usort_count_kernel, grid, 256, 0, st, x, y, z, n, g, t, wrap, sc.cnt);
  BR_LAUNCH(ctx, scan_partial_kernel, sc.nsb, 256, 0, st, sc.cnt, sc.sums, sc.m);
  BR_LAUNCH(ctx, scan_sums_kernel, 1, 32, 0, st, sc.sums, sc.nsb, sc.total);
  BR_LAUNCH(ctx, scan_final_kernel, sc.nsb, 256, 0, st, sc.cnt, sc.sums, sc.cursor, sc.starts, sc.m);
  BR_LAUNCH(ctx, usort_reorder_kernel, cdiv((size_t)n, 256 * USORT_PPT), 256, 0, st, x, y, z, w, n, g, t, wrap, sc.cursor, rec, inv, ctx->d_oob,
            ctx->d_hash);
  out->rec = rec;
  out->n_valid = sc.starts + t.ntiles;
  out->starts = sc.starts;
  out->ntiles = t.ntiles;
  out->inv = inv;
  ctx->sortc_x = x;
  ctx->sortc_y = y;
  ctx->sortc_z = z;
  ctx->sortc_n = n;
  ctx->sortc_ntiles = t.ntiles;
  ctx->sortc_slab_mode = key_mode;
  ctx->sortc_valid = true;
  return BAOREC_OK;
}

// Does the sort kept by the last scatter describe exactly these arrays?  Pointers, count and a content
// hash recomputed from the arrays (one 12 B/particle streaming pass + a 4-byte read-back).
static int sort_cache_hit(baorec_ctx* ctx, const float* x, const float* y, const float* z, int64_t n, cudaStream_t st,
                          bool* hit) {
  *hit = false;
  if (!ctx->sortc_valid || ctx->slab_mode != ctx->sortc_slab_mode || ctx->sortc_x != x || ctx->sortc_y != y ||
      ctx->sortc_z != z || ctx->sortc_n != n)
    return BAOREC_OK;
  BR_CUDA(cudaMemsetAsync(ctx->d_hash + 1, 0, sizeof(unsigned long long), st));
  unsigned grid = cdiv((size_t)n, 256 * 8);
  BR_LAUNCH(ctx, hash_positions_kernel, grid, 256, 0, st, x, y, z, n, ctx->d_hash + 1);
  BR_LAUNCH(ctx, hash_compare_kernel, 1, 1, 0, st, ctx->d_hash);
  unsigned long long flag = 0;
  BR_CUDA(cudaMemcpyAsync(&flag, ctx->d_hash + 2, sizeof(flag), cudaMemcpyDeviceToHost, st));
  BR_CUDA(cudaStreamSynchronize(st));
  *hit = flag != 0;
  return BAOREC_OK;
}

static bool use_binning(const baorec_ctx* ctx, int64_t n) {
  return n >= ctx->opt_bin_min_particles && n < ((int64_t)1 << 32) && ctx->nz >= 8;
}

int scatter(baorec_ctx* ctx, float* rho, float* x, float* y, float* z, const float* w, int64_t n, int wrap, int mas,
            cudaStream_t st) {
  if (ctx->prebin_valid) {  // a forked read-back sort that was never joined (error path) owns the binning buffers
    BR_CUDA(cudaStreamWaitEvent(st, ctx->ev_join, 0));
    ctx->prebin_valid = false;
  }
  if (n == 0) return BAOREC_OK;
  BoxGeom g = geom_of(ctx);
  const bool tsc = mas != BAOREC_MAS_CIC, pcs = mas == BAOREC_MAS_PCS;  // tsc: a stencil scheme (TSC or PCS)
  if (ctx->opt_det_scatter && (tsc || ctx->slab_mode != 0)) {
    set_error("option deterministic_scatter covers the single-GPU CIC scatter only (requested: %s%s); unset it for this call",
              tsc ? "TSC / PCS" : "CIC", ctx->slab_mode != 0 ? " on slabs" : "");
    return BAOREC_ERR_INVALID;  // not silently ignored: the caller asked for bit-reproducible meshes
  }
  if (ctx->opt_det_scatter) {
    // bit-reproducible CIC: 64-bit fixed-point integer reductions, one rounding to Float32 at the end
    unsigned long long* acc;
    BR_TRY(need_t(ctx, BUF_DET, ctx->M, &acc));
    BR_CUDA(cudaMemsetAsync(acc, 0, ctx->M * sizeof(unsigned long long), st));
    if (use_binning(ctx, n) && ctx->opt_unified_sort && ctx->opt_gather_tiles) {
      BinResult b;
      BR_TRY(unified_sort(ctx, x, y, z, w, n, wrap, st, &b));  // same sort (and the read-back reuses it)
      BR_LAUNCH(ctx, scatter_records_fixed_kernel, cdiv((size_t)n, 256), 256, 0, st, acc, b.rec, n, g, wrap, ctx->d_oob);
    } else {
      ctx->sortc_valid = false;
      BR_LAUNCH(ctx, scatter_direct_fixed_kernel, cdiv((size_t)n, 256), 256, 0, st, acc, x, y, z, w, n, g, wrap,
                ctx->d_oob);
    }
    size_t blocks = cdiv(ctx->M, 256);
    if (blocks > (size_t)148 * 32) blocks = (size_t)148 * 32;
    BR_LAUNCH(ctx, fixed_to_float_kernel, (unsigned)blocks, 256, 0, st, rho, acc, ctx->M);
    return BAOREC_OK;
  }
  if (use_binning(ctx, n) && ctx->opt_unified_sort && !tsc && ctx->slab_mode <= 1 && ctx->opt_gather_tiles) {
    BinResult b;
    BR_TRY(unified_sort(ctx, x, y, z, w, n, wrap, st, &b));
    if (ctx->opt_scatter_pairs >= 2 && ((uintptr_t)rho & 15) == 0)
      BR_LAUNCH(ctx, scatter_records_pairs_kernel<2>, cdiv((size_t)n, 256), 256, 0, st, rho, b.rec, n, g, wrap, ctx->d_oob);
    else if (ctx->opt_scatter_pairs && ((uintptr_t)rho & 7) == 0)
      BR_LAUNCH(ctx, scatter_records_pairs_kernel<1>, cdiv((size_t)n, 256), 256, 0, st, rho, b.rec, n, g, wrap, ctx->d_oob);
    else
      BR_LAUNCH(ctx, scatter_records_kernel, cdiv((size_t)n, 256), 256, 0, st, rho, b.rec, n, g, wrap, ctx->d_oob);
    return BAOREC_OK;
  }
  ctx->sortc_valid = false;  // the paths below reuse the binning buffers
  if (use_binning(ctx, n)) {
    BinResult b;
    if (ctx->opt_scatter_tiles && !tsc) BR_TRY(bin_tiles_scatter(ctx, x, y, z, w, n, wrap, st, &b));
    else BR_TRY(bin_particles<BIN_SCATTER>(ctx, x, y, z, w, n, wrap, mas, st, &b));
    unsigned grid = cdiv((size_t)n, 256);
    if (pcs && ctx->opt_scatter_pairs && ((uintptr_t)rho & 15) == 0)
      BR_LAUNCH(ctx, scatter_sorted_pcs_vec_kernel, grid, 256, 0, st, rho, b.rec, b.n_valid, g, wrap);
    else if (pcs) BR_LAUNCH(ctx, scatter_sorted_kernel<BAOREC_MAS_PCS>, grid, 256, 0, st, rho, b.rec, b.n_valid, g, wrap);
    else if (tsc && ctx->opt_scatter_pairs && ((uintptr_t)rho & 15) == 0)
      BR_LAUNCH(ctx, scatter_sorted_tsc_vec_kernel, grid, 256, 0, st, rho, b.rec, b.n_valid, g, wrap);
    else if (tsc) BR_LAUNCH(ctx, scatter_sorted_kernel<BAOREC_MAS_TSC>, grid, 256, 0, st, rho, b.rec, b.n_valid, g, wrap);
    else if (ctx->opt_scatter_pairs && ((uintptr_t)rho & 15) == 0)
      BR_LAUNCH(ctx, scatter_sorted_cic_vec_kernel, grid, 256, 0, st, rho, b.rec, b.n_valid, g, wrap);
    else BR_LAUNCH(ctx, scatter_sorted_kernel<BAOREC_MAS_CIC>, grid, 256, 0, st, rho, b.rec, b.n_valid, g, wrap);
    return BAOREC_OK;
  }
  unsigned grid = cdiv((size_t)n, 256);
  if (pcs) BR_LAUNCH(ctx, scatter_direct_kernel<BAOREC_MAS_PCS>, grid, 256, 0, st, rho, x, y, z, w, n, g, wrap,
                     ctx->d_oob);
  else if (tsc) BR_LAUNCH(ctx, scatter_direct_kernel<BAOREC_MAS_TSC>, grid, 256, 0, st, rho, x, y, z, w, n, g, wrap,
                     ctx->d_oob);
  else BR_LAUNCH(ctx, scatter_direct_kernel<BAOREC_MAS_CIC>, grid, 256, 0, st, rho, x, y, z, w, n, g, wrap,
                 ctx->d_oob);
  return BAOREC_OK;
}

int gather_prebin(baorec_ctx* ctx, const float* x, const float* y, const float* z, int64_t n, int mas,
                  cudaStream_t st) {
  ctx->prebin_valid = false;
  if (!ctx->opt_overlap_sort || n == 0 || !use_binning(ctx, n) || !ctx->opt_gather_tiles || mas != BAOREC_MAS_CIC)
    return BAOREC_OK;
  if (ctx->sortc_valid && ctx->slab_mode == ctx->sortc_slab_mode && ctx->sortc_x == x && ctx->sortc_y == y &&
      ctx->sortc_z == z && ctx->sortc_n == n)
    return BAOREC_OK;  // candidate for the sort run! kept: gather3 validates and reuses it
  ctx->sortc_valid = false;
  BR_CUDA(cudaEventRecord(ctx->ev_fork, st));
  BR_CUDA(cudaStreamWaitEvent(ctx->side_stream, ctx->ev_fork, 0));
  BinResult b;
  BR_TRY(bin_tiles(ctx, x, y, z, n, mas, ctx->side_stream, &b));
  BR_CUDA(cudaEventRecord(ctx->ev_join, ctx->side_stream));
  ctx->prebin_x = x;
  ctx->prebin_y = y;
  ctx->prebin_z = z;
  ctx->prebin_n = n;
  ctx->prebin_mas = mas;
  ctx->prebin_slab_mode = ctx->slab_mode;
  ctx->prebin_rec = b.rec;
  ctx->prebin_nvalid = b.n_valid;
  ctx->prebin_starts = b.starts;
  ctx->prebin_ntiles = b.ntiles;
  ctx->prebin_inv = b.inv;
  ctx->prebin_valid = true;
  return BAOREC_OK;
}

int gather3(baorec_ctx* ctx, const float* fx, const float* fy, const float* fz, const float* x, const float* y,
            const float* z, int64_t n, float* ox, float* oy, float* oz, int mas, int field, float f, int has_los,
            const float* los, int positions, cudaStream_t st) {
  const bool joined = ctx->prebin_valid && ctx->prebin_x == x && ctx->prebin_y == y && ctx->prebin_z == z &&
                      ctx->prebin_n == n && ctx->prebin_mas == mas && ctx->prebin_slab_mode == ctx->slab_mode;
  if (ctx->prebin_valid) {
    // whatever happens next must be ordered after the forked sort (it owns the binning buffers)
    BR_CUDA(cudaStreamWaitEvent(st, ctx->ev_join, 0));
    ctx->prebin_valid = false;
  }
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
  a.sorted_out = nullptr;
  const bool one = (fy == nullptr);
  const bool tsc = mas != BAOREC_MAS_CIC, pcs = mas == BAOREC_MAS_PCS;  // tsc: a stencil scheme (TSC or PCS)
  if (use_binning(ctx, n)) {
    BinResult b;
    bool reuse = false;
    if (!one && !tsc && !joined && ctx->opt_gather_tiles) BR_TRY(sort_cache_hit(ctx, x, y, z, n, st, &reuse));
    if (reuse) {
      // the sort run! made for the scatter is this catalog's sort: records, tile starts, inverse permutation
      TileGeom t;
      t.nxc = (ctx->nx + TILE_X - 1) / TILE_X;
      t.nyc = (ctx->ny + TILE_Y - 1) / TILE_Y;
      t.ntiles = ctx->sortc_ntiles;
      TileScratch sc;
      BR_TRY(tile_scratch(ctx, t, &sc));
      b.rec = (float4*)ctx->bufs[BUF_BINIDX].p;
      b.n_valid = sc.starts + t.ntiles;
      b.starts = sc.starts;
      b.ntiles = t.ntiles;
      b.inv = (unsigned*)ctx->bufs[BUF_BININV].p;
      BR_LAUNCH(ctx, add_oob_kernel, 1, 1, 0, st, sc.cnt + t.ntiles, ctx->d_oob);  // trash-bin size -> out-of-box counter
      ctx->n_sort_reuse++;
    } else if (joined) {
      b.rec = (float4*)ctx->prebin_rec;
      b.n_valid = ctx->prebin_nvalid;
      b.starts = ctx->prebin_starts;
      b.ntiles = ctx->prebin_ntiles;
      b.inv = ctx->prebin_inv;
    } else if (ctx->opt_gather_tiles) {
      ctx->sortc_valid = false;
      BR_TRY(bin_tiles(ctx, x, y, z, n, mas, st, &b));
    } else {
      ctx->sortc_valid = false;
      BR_TRY(bin_particles<BIN_GATHER>(ctx, (float*)x, (float*)y, (float*)z, nullptr, n, 0, mas, st, &b));
    }
    if (b.starts && !tsc) {
      TileGeom t;
      t.nxc = (ctx->nx + TILE_X - 1) / TILE_X;
      t.nyc = (ctx->ny + TILE_Y - 1) / TILE_Y;
      t.ntiles = b.ntiles;
      static bool carve = false;
      if (!carve) {
        cudaFuncSetAttribute(gather_tile_kernel<3, 0>, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
        cudaFuncSetAttribute(gather_tile_kernel<3, 1>, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
        carve = true;
      }
      GatherMaps maps;
      const bool tma = ctx->opt_gather_stage == 2 && gather_maps(a, one ? 1 : 3, g, &maps);
      if (one) {
        if (tma) BR_LAUNCH_NAMED(ctx, "gather_tile_tma_kernel<1>", gather_tile_tma_kernel<1>, b.ntiles, 128, 0, st, maps, a, b.rec, b.starts, g, t);
        else BR_LAUNCH_NAMED(ctx, "gather_tile_kernel<1>", (gather_tile_kernel<1, 0>), b.ntiles, 128, 0, st, a, b.rec, b.starts, g, t);
      } else {
        // results land in sorted order (coalesced), then one un-permute pass: the 3 x 4 B scattered
        // stores of the direct version cost three random DRAM read-modify-writes per particle
        float4* so;
        BR_TRY(need_t(ctx, BUF_BINOUT, (size_t)n, &so));
        a.sorted_out = so;
        if (tma) {
          static bool carve_tma = false;
          if (!carve_tma) {
            cudaFuncSetAttribute(gather_tile_tma_kernel<3>, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
            carve_tma = true;
          }
          BR_LAUNCH_NAMED(ctx, "gather_tile_tma_kernel<3>", gather_tile_tma_kernel<3>, b.ntiles, 128, 0, st, maps, a, b.rec, b.starts, g, t);
        } else if (ctx->opt_gather_stage) BR_LAUNCH_NAMED(ctx, "gather_tile_kernel<3> [stage=1]", (gather_tile_kernel<3, 1>), b.ntiles, 128, 0, st, a, b.rec, b.starts, g, t);
        else BR_LAUNCH_NAMED(ctx, "gather_tile_kernel<3>", (gather_tile_kernel<3, 0>), b.ntiles, 128, 0, st, a, b.rec, b.starts, g, t);
        BR_LAUNCH(ctx, unsort_kernel, cdiv((size_t)n, 256), 256, 0, st, a, so, b.inv, n, b.n_valid);
        return BAOREC_OK;
      }
      // single-field tile gather wrote in place; out-of-box outputs are zeroed below
      BR_LAUNCH(ctx, gather_trash_kernel, cdiv((size_t)n, 256), 256, 0, st, a, b.rec, b.n_valid, n, 1);
    } else {
      unsigned grid = cdiv((size_t)n, 256);
      if (pcs) {
        if (one) BR_LAUNCH(ctx, (gather_sorted_kernel<1, BAOREC_MAS_PCS>), grid, 256, 0, st, a, b.rec, b.n_valid, n, g);
        else BR_LAUNCH(ctx, (gather_sorted_kernel<3, BAOREC_MAS_PCS>), grid, 256, 0, st, a, b.rec, b.n_valid, n, g);
      } else if (tsc) {
        if (one) BR_LAUNCH(ctx, (gather_sorted_kernel<1, BAOREC_MAS_TSC>), grid, 256, 0, st, a, b.rec, b.n_valid, n, g);
        else BR_LAUNCH(ctx, (gather_sorted_kernel<3, BAOREC_MAS_TSC>), grid, 256, 0, st, a, b.rec, b.n_valid, n, g);
      } else {
        if (one) BR_LAUNCH(ctx, (gather_sorted_kernel<1, BAOREC_MAS_CIC>), grid, 256, 0, st, a, b.rec, b.n_valid, n, g);
        else BR_LAUNCH(ctx, (gather_sorted_kernel<3, BAOREC_MAS_CIC>), grid, 256, 0, st, a, b.rec, b.n_valid, n, g);
      }
    }
    return BAOREC_OK;
  }
  unsigned grid = cdiv((size_t)n, 256);
  if (pcs) {
    if (one) BR_LAUNCH(ctx, (gather_direct_kernel<1, BAOREC_MAS_PCS>), grid, 256, 0, st, a, g, ctx->d_oob);
    else BR_LAUNCH(ctx, (gather_direct_kernel<3, BAOREC_MAS_PCS>), grid, 256, 0, st, a, g, ctx->d_oob);
  } else if (tsc) {
    if (one) BR_LAUNCH(ctx, (gather_direct_kernel<1, BAOREC_MAS_TSC>), grid, 256, 0, st, a, g, ctx->d_oob);
    else BR_LAUNCH(ctx, (gather_direct_kernel<3, BAOREC_MAS_TSC>), grid, 256, 0, st, a, g, ctx->d_oob);
  } else {
    if (one) BR_LAUNCH(ctx, (gather_direct_kernel<1, BAOREC_MAS_CIC>), grid, 256, 0, st, a, g, ctx->d_oob);
    else BR_LAUNCH(ctx, (gather_direct_kernel<3, BAOREC_MAS_CIC>), grid, 256, 0, st, a, g, ctx->d_oob);
  }
  return BAOREC_OK;
}

// ---- finite-difference read-back (fd_gradient in mas_math.cuh) -------------------------------------------------------
// One thread per particle; in tile order (the unified sort's records) the 32 cells a particle reads are its
// neighbours' cells too and come from L1 / L2.
__global__ void __launch_bounds__(256)
gather_fd_sorted_kernel(GatherArgs a, const float4* __restrict__ rec, const unsigned* __restrict__ n_valid, int64_t n, BoxGeom g) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n || i >= (int64_t)__ldg(n_valid)) return;  // the trash bin is zero-filled by unsort_kernel
  const float4 p = rec[i];
  float val[3];
  if (!fd_gradient(a.f[0], g, p.x, p.y, p.z, val)) val[0] = val[1] = val[2] = 0.f;
  shifts_epilogue<3>(a, val, p.x, p.y, p.z, 0, i);
}
__global__ void __launch_bounds__(256)
gather_fd_direct_kernel(GatherArgs a, BoxGeom g, unsigned long long* __restrict__ oob) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= a.n) return;
  const float px = a.x[i], py = a.y[i], pz = a.z[i];
  float val[3];
  if (!fd_gradient(a.f[0], g, px, py, pz, val)) {
    val[0] = val[1] = val[2] = 0.f;
    atomicAdd(oob, 1ULL);
  }
  shifts_epilogue<3>(a, val, px, py, pz, i);
}

// grad phi interpolated to the particles by finite differences + the read_shifts epilogue (single GPU)
int gather_fd(baorec_ctx* ctx, const float* phi, const float* x, const float* y, const float* z, int64_t n, float* ox,
              float* oy, float* oz, int field, float f, int has_los, const float* los, int positions, cudaStream_t st) {
  if (ctx->prebin_valid) {
    BR_CUDA(cudaStreamWaitEvent(st, ctx->ev_join, 0));
    ctx->prebin_valid = false;
  }
  if (n == 0) return BAOREC_OK;
  BoxGeom g = geom_of(ctx);
  GatherArgs a;
  a.f[0] = phi;
  a.f[1] = a.f[2] = nullptr;
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
  a.sorted_out = nullptr;
  if (use_binning(ctx, n) && ctx->opt_gather_tiles) {
    BinResult b;
    bool reuse = false;
    BR_TRY(sort_cache_hit(ctx, x, y, z, n, st, &reuse));
    if (reuse) {
      TileGeom t;
      t.nxc = (ctx->nx + TILE_X - 1) / TILE_X;
      t.nyc = (ctx->ny + TILE_Y - 1) / TILE_Y;
      t.ntiles = ctx->sortc_ntiles;
      TileScratch sc;
      BR_TRY(tile_scratch(ctx, t, &sc));
      b.rec = (float4*)ctx->bufs[BUF_BINIDX].p;
      b.n_valid = sc.starts + t.ntiles;
      b.inv = (unsigned*)ctx->bufs[BUF_BININV].p;
      BR_LAUNCH(ctx, add_oob_kernel, 1, 1, 0, st, sc.cnt + t.ntiles, ctx->d_oob);
      ctx->n_sort_reuse++;
    } else {
      ctx->sortc_valid = false;
      BR_TRY(bin_tiles(ctx, x, y, z, n, BAOREC_MAS_CIC, st, &b));
    }
    float4* so;
    BR_TRY(need_t(ctx, BUF_BINOUT, (size_t)n, &so));
    a.sorted_out = so;
    BR_LAUNCH(ctx, gather_fd_sorted_kernel, cdiv((size_t)n, 256), 256, 0, st, a, b.rec, b.n_valid, n, g);
    BR_LAUNCH(ctx, unsort_kernel, cdiv((size_t)n, 256), 256, 0, st, a, so, b.inv, n, b.n_valid);
    return BAOREC_OK;
  }
  BR_LAUNCH(ctx, gather_fd_direct_kernel, cdiv((size_t)n, 256), 256, 0, st, a, g, ctx->d_oob);
  return BAOREC_OK;
}

// ---- parity probes -------------------------------------------------------------------------------------
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

// ---- setup_box reductions (src/utils.jl:100-109) --------------------------------------------
__device__ __forceinline__ unsigned f2ord(float f) {
  unsigned u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
static float ord2f(unsigned u) {
  unsigned v = (u & 0x80000000u) ? (u & 0x7fffffffu) : ~u;
  float f;
  memcpy(&f, &v, 4);
  return f;
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
  BR_REQUIRE(mas == BAOREC_MAS_CIC || mas == BAOREC_MAS_TSC || mas == BAOREC_MAS_PCS, "unknown mas");
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
