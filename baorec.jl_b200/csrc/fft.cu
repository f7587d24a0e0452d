// Strided batched 1-D FFT passes (the y and z axes of the 3-D transform) with the k-space
// operators folded into the z pass.
//
// cuFFT's 3-D R2C/C2R at 1024^3 spends 2 x 2.5 ms in its two strided passes (3.4 TB/s) and the
// k-space multiplies then cost another read + write of the half-mesh.  Here the x axis stays a
// batched contiguous cuFFT 1-D transform and the y / z axes are our own column kernel:
//   * a CTA owns a tile of 8 adjacent columns (64-byte row segments: the columns are contiguous in memory) and
//     32 threads per column; thread t of a column loads rows t, t + 32, ... straight into registers;
//   * the length-N transform is a four-step FFT N = 32 x M (M = N / 32 = 8 .. 64): an M-point transform in each
//     thread's registers (fft_radix.cuh: generated, fully unrolled radix-2 butterflies with literal twiddles), the
//     W_N^(t k) twiddles, ONE exchange through shared memory (padded: conflict-free both ways), a 32-point
//     transform in registers -- ~40 instructions per point instead of the ~110 of the first version, whose 32-point
//     stage ran across the lanes of a warp with shuffles and made the pass issue-bound (3.2 ms against cuFFT's 2.5);
//   * the output of a forward transform sits in the registers exactly where the inverse transform wants its input
//     (thread t holds indices t + 32 j), so the z pass of run! applies the k-space operator in registers and runs the
//     inverse z transform in the same kernel (forward + operator + inverse = ONE read and ONE write of the
//     half-mesh instead of four passes), and the read-back emits the three displacement fields from one read.
#include <math.h>

#include "internal.cuh"
#include "fft_column.cuh"
#include "kspace_ops.cuh"

namespace baorec {

// resident CTAs per SM.  The kernels fit 3 (80 registers, 20 bytes of spills at N = 1024), but measured on B200 at 1024^3
// that is SLOWER (plain pass 3.0 instead of 2.2 ms, fused z pass 9.7 instead of 6.9 ms): a z pass touches every plane
// of the mesh with ~19 KB per plane in flight whatever the tile shape, so more tiles in flight means more open DRAM
// rows and less shared memory left for L1 -- two CTAs per SM it is.
#ifndef FFT_CTAS_1024
#define FFT_CTAS_1024 2
#endif

struct ColGeom {
  int ncols;            // length of the contiguous axis (nx/2+1)
  int nouter;           // z passes: number of rows (y) -- their columns are tiled over the flat index iy * ncols + ix
  size_t stride;        // elements between consecutive points of a transform
  size_t outer_stride;  // elements between consecutive tiles rows (blockIdx.y)
  const float* kx;      // k tables: contiguous axis, outer axis, transform axis
  const float* kouter;
  const float* ktrans;
  int outer_is_y;       // 1: outer = y, transform = z (z pass); 0: outer = z, transform = y (y pass)
  int prefetch;         // > 0: every CTA asks L2 for the tile of the CTA `prefetch` places further on in launch order (option "fft_prefetch")
};

// The fused z passes keep the twiddle table and the k_z / Gaussian tables in SHARED memory, after the exchange buffer.
// Read through L1 (__ldg) the streaming column loads and stores keep evicting them (ncu: L1 hit rate 14 - 21 %) and
// every twiddle multiply and every k_z^2 waits for L2: a third of the warp-stall samples of fft_z_disp sat on the first
// use of a k_z value.  Measured at 1024^3: fft_z_solve 6.93 -> 6.17 ms, fft_z_disp 8.41 -> 7.96 ms.  The plain pass
// (fft_cols_kernel) is better off with __ldg: there the copy and its barrier cost more than they save (2.02 -> 2.24 ms).
template <int N, int TX>
__device__ __forceinline__ float2* stage_twiddles(float2* S, const float2* __restrict__ tw) {
  float2* T = S + TX * Xch<N, TX>::COL;
  for (int i = threadIdx.x; i < N; i += 32 * TX) T[i] = __ldg(tw + i);
  return T;
}

// STREAM: evict-first loads (ld.global.cs).  Measured at 1024^3: together with evict-first stores they help the read-back's
// z pass (fft_z_disp 5.73 -> 5.53 ms: one read, three write streams), do nothing for fft_z_solve and slow the plain
// y pass down (2.02 -> 2.21 ms) -- so only fft_z_disp uses them.
template <int N, bool STREAM = false>
__device__ __forceinline__ void load_column(float2 (&v)[N / 32], const float2* __restrict__ src, size_t stride, int t) {
  const float2* p = src + (size_t)t * stride;
  const size_t step = 32 * stride;
#pragma unroll
  for (int j = 0; j < N / 32; j++) v[j] = STREAM ? __ldcs(p + (size_t)j * step) : p[(size_t)j * step];
}

// L2 prefetch of the tile another CTA will load `dist` CTAs later (roughly: when this CTA's slot is free again): the
// column loads of that CTA then wait for L2 instead of DRAM.  One prefetch at the first and one at the last byte of
// every row segment (the segments of a y pass are not aligned), rows dealt over the threads of the CTA.
template <int N, int TX>
__device__ __forceinline__ void prefetch_tile(const float2* __restrict__ in, const ColGeom& cg, int dist, bool flat) {
  const unsigned nbx = gridDim.x;
  const unsigned long long lin = (unsigned long long)blockIdx.y * nbx + blockIdx.x + (unsigned)dist;
  const unsigned by = (unsigned)(lin / nbx), bx = (unsigned)(lin - (unsigned long long)by * nbx);
  if (by >= gridDim.y) return;
  const size_t first = (size_t)by * cg.outer_stride + (size_t)bx * TX;
  const size_t limit = flat ? (size_t)cg.ncols * cg.nouter : (size_t)by * cg.outer_stride + cg.ncols;  // one past the last column of the row
  size_t last = first + TX - 1;
  if (last >= limit) last = limit - 1;
  if (first >= limit) return;
  for (int r = threadIdx.x; r < N; r += 32 * TX) {
    const float2* a = in + first + (size_t)r * cg.stride;
    const float2* b = in + last + (size_t)r * cg.stride;
    asm volatile("prefetch.global.L2 [%0];" ::"l"(a));
    asm volatile("prefetch.global.L2 [%0];" ::"l"(b));
  }
}

// ---- operators applied in the z pass (frequency index f along the transform axis) ----------------
// smoothing + (rho/mean - 1)/bias + all n_iter fixed-LOS iterations: FusedLosOp<0> of kspace_ops.cuh, value-returning
// (same Float32 / Float64 products, Gaussian from the per-axis Float64 tables).  Returns the unnormalised delta_final_k.
struct OpLosSolve {
  GaussTab gt;
  const double* scal;  // scal[8] = M / (A0 bias)
  float los[3];
  float beta;
  int n_iter;
  float invM;
  // per-column invariants (one thread works on ONE (ix, iy) column): the x-y part of the Gaussian times the
  // normalisation, the x-y part of the line-of-sight numerator, the first iteration's factor.  (gx gy) gz dc is the
  // product FusedLosOp forms as ((gx gy) gz) dc -- Float64, the association differs in the last bit only.
  struct Col {
    double gxy;
    float kx, ky, cxy, fac1;
  };
  __device__ __forceinline__ Col column(float kx, float ky, int ix, int iy) const {
    Col c;
    c.gxy = __ldg(gt.gx + ix) * __ldg(gt.gy + iy);
    c.kx = kx;
    c.ky = ky;
    c.cxy = __fadd_rn(__fmul_rn(__fmul_rn(kx, kx), los[0]), __fmul_rn(__fmul_rn(ky, ky), los[1]));
    c.fac1 = __fdiv_rn(beta, __fadd_rn(1.0f, beta));
    return c;
  }
  __device__ __forceinline__ float2 solve(float2 v, const Col& c, float kz, double gz_dc, bool is_dc) const {
    float k2 = ksq(c.kx, c.ky, kz);
    // the Gaussian of the three axes and the normalisation multiply in Float64 and are rounded ONCE; the mode is then
    // scaled in Float32 (FusedLosOp rounds (double) v * s instead: one ulp apart, and four conversions per mode more)
    float s = (float)(c.gxy * gz_dc);
    if (is_dc) s = 0.f;
    float dsx = __fmul_rn(v.x, s), dsy = __fmul_rn(v.y, s);
    float num = __fadd_rn(c.cxy, __fmul_rn(__fmul_rn(kz, kz), los[2]));
    float mu = k2 > 0.f ? __fdividef(num, k2) : 0.f;  // 2 ulp; the IEEE division is 12 instructions and a slow-path branch per mode
    float drx = dsx, dry = dsy;
    if (n_iter >= 1) {
      float fm = __fmul_rn(c.fac1, mu);
      drx = __fsub_rn(dsx, __fmul_rn(fm, drx));
      dry = __fsub_rn(dsy, __fmul_rn(fm, dry));
      fm = __fmul_rn(beta, mu);
      for (int it = 2; it <= n_iter; it++) {
        drx = __fsub_rn(dsx, __fmul_rn(fm, drx));
        dry = __fsub_rn(dsy, __fmul_rn(fm, dry));
      }
    }
    return make_float2(drx, dry);
  }
};

// Psi_c = i k_c delta_k / k^2 / M (density) or i k_c phi_k / M (potential)
struct OpDisp {
  int potential;
  float invM;
  __device__ __forceinline__ float2 comp(float2 v, float kx, float ky, float kz, int c) const {
    float s;
    if (potential) {
      s = invM;
    } else {
      float k2 = ksq(kx, ky, kz);
      s = k2 > 0.f ? __fdiv_rn(invM, k2) : 0.f;
    }
    float kc = c == 0 ? kx : (c == 1 ? ky : kz);
    return make_float2(__fmul_rn(__fmul_rn(-v.y, s), kc), __fmul_rn(__fmul_rn(v.x, s), kc));
  }
};

// ---- kernels -----------------------------------------------------------------------------------------
// plain pass: out = FFT_DIR(in) along the strided axis (in place allowed: a CTA reads its whole tile before it writes)
template <int N, int DIR, int TX = FFT_TX>
__global__ void __launch_bounds__(32 * TX, N >= 2048 ? 1 : (FFT_CTAS_1024 * 8) / TX)
fft_cols_kernel(const float2* __restrict__ in, float2* __restrict__ out, ColGeom cg, const float2* __restrict__ tw,
                int hermitian_edges) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float2* S = reinterpret_cast<float2*>(smem_raw);
  constexpr int M = N / 32, MT = M >= 32 ? M / 32 : 1;
  const int c = threadIdx.x & (TX - 1), t = threadIdx.x / TX;
  const int ix = blockIdx.x * TX + c;
  const bool valid = ix < cg.ncols;
  const size_t base = (size_t)blockIdx.y * cg.outer_stride + ix;
  float2 v[M], X[MT][32];
  if (valid) load_column<N>(v, in + base, cg.stride, t);
  else {
#pragma unroll
    for (int j = 0; j < M; j++) v[j] = make_float2(0.f, 0.f);
  }
  if (cg.prefetch > 0) prefetch_tile<N, TX>(in, cg, cg.prefetch, false);
  fft_column<N, DIR, false, TX>(v, X, S, tw, c, t);
  // Last pass before the C2R along x: the kx = 0 and kx = Nyquist columns must be real there.  FFTW
  // and pocketfft ignore their imaginary part; cuFFT's 1-D C2R does not, so it is dropped here
  // (it is non-zero only for inputs with power at the Nyquist modes, e.g. i k multiplications).
  const bool drop = hermitian_edges && (ix == 0 || ix == cg.ncols - 1);
  if (valid) {
#pragma unroll
    for (int m = 0; m < MT; m++) {
      const int k2 = t + 32 * m;
      if (M >= 32 || k2 < M) {
        float2* o = out + base + (size_t)k2 * cg.stride;
#pragma unroll
        for (int k1 = 0; k1 < 32; k1++) {
          float2 r = X[m][fft_bitrev<32>(k1)];
          if (drop) r.y = 0.f;
          o[(size_t)k1 * M * cg.stride] = r;
        }
      }
    }
  }
}

// z pass of the fused fixed-LOS solve (N >= 1024: the forward output is the inverse input, register for register):
// forward z FFT, operator, [delta_k kept], inverse z FFT; in place.
template <int N, int TX = FFT_TX>
__global__ void __launch_bounds__(32 * TX, N >= 2048 ? 1 : (FFT_CTAS_1024 * 8) / TX)
fft_z_solve_kernel(float2* __restrict__ data, float2* __restrict__ keep, ColGeom cg, const float2* __restrict__ tw,
                   OpLosSolve op) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float2* S = reinterpret_cast<float2*>(smem_raw);
  constexpr int M = N / 32, MT = M / 32;
  static_assert(M >= 32, "the fused z pass needs N >= 1024");
  // One tile row per y (grid.y = ny; the last tile of a row holds one column: xh = nx/2 + 1 is odd).  The flat tiling of
  // the other z passes (fft_z_disp_kernel) measured SLOWER here: 5.79 instead of 5.02 ms at 1024^3 -- this kernel
  // writes two planes-apart streams (delta_k kept, the column) next to the one it reads.
  const int c = threadIdx.x & (TX - 1), t = threadIdx.x / TX;
  const int ix0 = blockIdx.x * TX + c, iy = blockIdx.y;
  const bool valid = ix0 < cg.ncols;
  const int ix = valid ? ix0 : 0;
  const size_t base = (size_t)iy * cg.outer_stride + ix0;
  float2 v[M], X[MT][32];
  if (valid) load_column<N>(v, data + base, cg.stride, t);
  else {
#pragma unroll
    for (int j = 0; j < M; j++) v[j] = make_float2(0.f, 0.f);
  }
  if (cg.prefetch > 0) prefetch_tile<N, TX>(data, cg, cg.prefetch, false);
  float2* T = stage_twiddles<N, TX>(S, tw);
  double* GZ = reinterpret_cast<double*>(T + N);   // Gaussian of the z axis (Float64), then k_z
  float* KZ = reinterpret_cast<float*>(GZ + N);
  for (int i = threadIdx.x; i < N; i += 32 * TX) {
    GZ[i] = __ldg(op.gt.gz + i);
    KZ[i] = __ldg(cg.ktrans + i);
  }
  __syncthreads();
  fft_column<N, 1, true, TX>(v, X, S, T, c, t);
  const OpLosSolve::Col col = op.column(__ldg(cg.kx + ix), __ldg(cg.kouter + iy), ix, iy);
  const double dc8 = __ldg(op.scal + 8);
#pragma unroll
  for (int m = 0; m < MT; m++)
#pragma unroll
    for (int k1 = 0; k1 < 32; k1++) {
      const int f = t + 32 * m + M * k1;      // = t + 32 (m + MT k1): the inverse transform's register m + MT k1
      float2 d = make_float2(0.f, 0.f);
      if (valid) {
        d = op.solve(X[m][fft_bitrev<32>(k1)], col, KZ[f], GZ[f] * dc8, (ix | iy | f) == 0);
        if (keep != nullptr) keep[base + (size_t)f * cg.stride] = d;
      }
      v[m + MT * k1] = make_float2(d.x * op.invM, d.y * op.invM);
    }
  __syncthreads();  // everybody has read the exchange buffer of the forward transform
  fft_column<N, -1, true, TX>(v, X, S, T, c, t);
  if (valid) {
#pragma unroll
    for (int m = 0; m < MT; m++)
#pragma unroll
      for (int k1 = 0; k1 < 32; k1++)
        data[base + (size_t)(t + 32 * m + M * k1) * cg.stride] = X[m][fft_bitrev<32>(k1)];
  }
}

// z pass of the displacement read-back from a kept delta_k (phi_k).  i k_x and i k_y are constants of a column, so the x
// and y fields share ONE inverse z transform of G = delta_k / (k^2 M) (phi_k / M); the z field is the inverse transform
// of H = i k_z G.  Two transforms and three stores per column; delta_k is read twice (the second read of the 64 KB
// tile comes from L2) so that no copy of it has to stay in registers while a transform runs -- the first version
// kept it, needed 255 registers (one CTA per SM) and ran three transforms: 11.3 ms at 1024^3.
template <int N, int TX = FFT_TX>
__global__ void __launch_bounds__(32 * TX, N >= 2048 ? 1 : (FFT_CTAS_1024 * 8) / TX)
fft_z_disp_kernel(const float2* __restrict__ in, float2* __restrict__ o0, float2* __restrict__ o1,
                  float2* __restrict__ o2, ColGeom cg, const float2* __restrict__ tw, OpDisp op) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float2* S = reinterpret_cast<float2*>(smem_raw);
  constexpr int M = N / 32, MT = M >= 32 ? M / 32 : 1;
  const int c = threadIdx.x & (TX - 1), t = threadIdx.x / TX;
  // tiles run over the FLAT column index q = iy * xh + ix (the columns of a z pass are contiguous in memory across
  // rows): every tile is one aligned 64- / 128-byte segment per plane and none is partly empty (xh = nx/2 + 1 is odd).
  // Measured at 1024^3: 6.48 -> 5.70 ms here, the plain z pass 2.35 -> 2.07 ms.
  const unsigned q = blockIdx.x * TX + c;
  const bool valid = q < (unsigned)cg.ncols * (unsigned)cg.nouter;
  const int iy = valid ? (int)(q / (unsigned)cg.ncols) : 0, ix = valid ? (int)(q - (unsigned)iy * (unsigned)cg.ncols) : 0;
  const size_t base = q;
  const float kx = __ldg(cg.kx + ix), ky = __ldg(cg.kouter + iy);
  if (cg.prefetch > 0) prefetch_tile<N, TX>(in, cg, cg.prefetch, true);
  float2* T = stage_twiddles<N, TX>(S, tw);
  float* KZ = reinterpret_cast<float*>(T + N);
  for (int i = threadIdx.x; i < N; i += 32 * TX) KZ[i] = __ldg(cg.ktrans + i);
  __syncthreads();
#pragma unroll 1
  for (int pass = 0; pass < 2; pass++) {   // 0: H -> z field;  1: G -> x and y fields
    float2 v[M], X[MT][32];
    if (valid) {
      if (pass == 0) load_column<N, false>(v, in + base, cg.stride, t);  // read again by pass 1: stays in L2
      else load_column<N, true>(v, in + base, cg.stride, t);
    }
#pragma unroll
    for (int j = 0; j < M; j++) {
      const float kz = KZ[t + 32 * j];
      float s;
      if (op.potential) {
        s = op.invM;
      } else {
        float k2 = ksq(kx, ky, kz);
        s = k2 > 0.f ? __fdividef(op.invM, k2) : 0.f;  // MUFU.RCP + FMUL (2 ulp) instead of the ~12-instruction IEEE division
      }
      const float2 d = valid ? v[j] : make_float2(0.f, 0.f);
      // the products of DispOp: re = (-v.y s) k, im = (v.x s) k;  G itself for the shared transform
      v[j] = pass == 0 ? make_float2(__fmul_rn(__fmul_rn(-d.y, s), kz), __fmul_rn(__fmul_rn(d.x, s), kz))
                       : make_float2(__fmul_rn(d.x, s), __fmul_rn(d.y, s));
    }
    if (pass) __syncthreads();  // the first transform's exchange buffer has been read
    fft_column<N, -1, true, TX>(v, X, S, T, c, t);
    if (valid) {
#pragma unroll
      for (int m = 0; m < MT; m++) {
        const int k2 = t + 32 * m;
        if (M >= 32 || k2 < M) {
#pragma unroll
          for (int k1 = 0; k1 < 32; k1++) {
            const size_t o = base + (size_t)(k2 + M * k1) * cg.stride;
            const float2 r = X[m][fft_bitrev<32>(k1)];
            if (pass == 0) {
              __stcs(o2 + o, r);
            } else {  // i k g = (-g.y k, g.x k)
              __stcs(o0 + o, make_float2(__fmul_rn(-r.y, kx), __fmul_rn(r.x, kx)));
              __stcs(o1 + o, make_float2(__fmul_rn(-r.y, ky), __fmul_rn(r.x, ky)));
            }
          }
        }
      }
    }
  }
}

// sum over z of A(kx=0, ky=0, z) after the x and y passes = the DC mode of the 3-D transform
__global__ void dc_from_xy_kernel(const float2* __restrict__ a, size_t zstride, int nz, double* scal, int slot,
                                  double mul) {
  __shared__ double sh[32];
  double acc = 0.0;
  for (int z = threadIdx.x; z < nz; z += blockDim.x) acc += (double)a[(size_t)z * zstride].x;
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); w++) t += sh[w];
    scal[slot] = t;
    scal[slot + 8] = mul / t;
  }
}

// ---- host side ---------------------------------------------------------------------------------------------
static bool pow2_ok(int n) { return n >= 256 && n <= 2048 && (n & (n - 1)) == 0; }

// the fused z pass (forward + operator + inverse in one kernel) needs N >= 1024: below that the last step of the
// four-step transform leaves its output in 16 (8) of a column's 32 threads, not where the inverse reads its input
bool own_fft_fused_available(const baorec_ctx* ctx) { return own_fft_available(ctx) && ctx->nz >= 1024; }

bool own_fft_available(const baorec_ctx* ctx) {
  // -1 = auto: on where it was measured faster than cuFFT's 3-D plans (both strided axes a power of two >= 512:
  // 49.3 vs 52.8 ms per reconstruction at 1024^3, 6.5 vs 6.8 ms of transforms at 512^3; slower at 256^3)
  const bool want = ctx->opt_own_fft > 0 || (ctx->opt_own_fft < 0 && ctx->ny >= 512 && ctx->nz >= 512);
  return want && ctx->have_x_plans && pow2_ok(ctx->ny) && pow2_ok(ctx->nz) && ctx->d_tw[0] && ctx->d_tw[1];
}

// twiddle tables exp(-2 pi i k / n) for the y and z axes, computed in double
int own_fft_setup(baorec_ctx* ctx) {
  for (int a = 0; a < 2; a++) {
    if (ctx->d_tw[a]) {
      cudaFree(ctx->d_tw[a]);
      ctx->d_tw[a] = nullptr;
    }
    int n = a == 0 ? ctx->ny : ctx->nz;
    if (!pow2_ok(n)) continue;
    std::vector<float2> h(n);
    for (int k = 0; k < n; k++) {
      double ang = -2.0 * M_PI * (double)k / (double)n;
      h[k] = make_float2((float)cos(ang), (float)sin(ang));
    }
    BR_CUDA(cudaMalloc(&ctx->d_tw[a], sizeof(float2) * n));
    BR_CUDA(cudaMemcpy(ctx->d_tw[a], h.data(), sizeof(float2) * n, cudaMemcpyHostToDevice));
  }
  return BAOREC_OK;
}

// exchange buffer; the fused z passes add their tables (twiddles 8 N, + extra)
template <int N, int TX = FFT_TX>
static size_t tile_bytes() { return Xch<N, TX>::BYTES; }
template <int N, int TX = FFT_TX>
static size_t fused_bytes(size_t extra) { return Xch<N, TX>::BYTES + (size_t)N * sizeof(float2) + extra; }

// Columns per tile (option "fft_tile_cols"; the widths 4 and 16 are instantiated for N = 1024 only).  -1 = auto: the
// z passes of a 1024-point axis -- fused or plain -- take 16 columns (128-byte row segments, one CTA of 512 threads per
// SM), every other pass 8 (64-byte segments, two CTAs of 256 threads).  Measured at 1024^3
// (profiles/r2_ab_fft_tile_cols.jsonl, r2_fft_zplain_probe.jsonl): with 16 columns fft_z_solve 6.14 -> 5.00 ms,
// fft_z_disp 7.91 -> 6.42 ms, the plain z pass 3.0 -> 2.35 ms -- along z consecutive rows of a column are a whole plane
// (4 MB) apart, every row segment opens its own DRAM page, and a 128-byte segment moves twice the data per activation
// -- while the y pass loses (2.03 -> 2.21 ms: its rows are 4 KB apart, and one CTA per SM overlaps load, exchange and
// store worse than two), the 512-point z pass loses slightly (0.55 -> 0.58 ms at 512^3: 1 MB planes) and 4 columns
// (32-byte segments, four CTAs) lose everywhere (fft_z_disp 13.8 ms).
static int tile_cols(const baorec_ctx* ctx, int n, bool z_axis) {
  if (n != 1024) return FFT_TX;
  const int o = ctx->opt_fft_tile_cols;
  if (o == 4 || o == 16 || o == 8) return o;
  return z_axis ? 16 : FFT_TX;
}

#define FFT_DISPATCH_N(n, CALL)     \
  switch (n) {                      \
    case 256: { CALL(256); } break; \
    case 512: { CALL(512); } break; \
    case 1024: { CALL(1024); } break; \
    case 2048: { CALL(2048); } break; \
    default: set_error("own FFT: unsupported length %d", n); return BAOREC_ERR_INVALID; \
  }

template <class K>
static int set_smem(K kernel, size_t bytes) {
  if (bytes > 48 * 1024) BR_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
  return BAOREC_OK;
}

static ColGeom geom_y(const baorec_ctx* ctx) {  // transform along y, tiles over (x, z)
  ColGeom g;
  g.ncols = ctx->xh;
  g.nouter = ctx->nz;
  g.stride = ctx->xh;
  g.outer_stride = (size_t)ctx->xh * ctx->ny;
  g.kx = ctx->d_k[0];
  g.kouter = ctx->d_k[2];
  g.ktrans = ctx->d_k[1];
  g.outer_is_y = 0;
  // y passes: every CTA asks L2 for the tile of the CTA one wave (148 SMs) further on -- measured at 1024^3: 2.04 -> 1.93 ms
  // per pass with a distance of 148, 1.94 / 1.98 with 296 / 592 (profiles/r2_ab_fft_prefetch.jsonl)
  g.prefetch = ctx->opt_fft_prefetch < 0 ? 148 : ctx->opt_fft_prefetch;
  return g;
}
static ColGeom geom_z(const baorec_ctx* ctx) {  // transform along z, tiles over (x, y)
  ColGeom g;
  g.ncols = ctx->xh;
  g.nouter = ctx->ny;
  g.stride = (size_t)ctx->xh * ctx->ny;
  g.outer_stride = ctx->xh;
  g.kx = ctx->d_k[0];
  g.kouter = ctx->d_k[1];
  g.ktrans = ctx->d_k[2];
  g.outer_is_y = 1;
  // z passes: the same prefetch costs time (fft_z_solve 5.03 -> 5.74 ms, fft_z_disp 5.52 -> 6.14 ms: their loads already
  // keep DRAM busy, the extra requests only thrash the row buffers) -- only when asked for explicitly
  g.prefetch = ctx->opt_fft_prefetch > 0 ? ctx->opt_fft_prefetch : 0;
  return g;
}

template <int N, int DIR, int TX>
static int launch_cols_tx(baorec_ctx* ctx, const float2* in, float2* out, const ColGeom& g, int nouter, const float2* tw,
                          int herm, cudaStream_t st) {
  BR_TRY(set_smem((fft_cols_kernel<N, DIR, TX>), (tile_bytes<N, TX>())));
  dim3 grid(cdiv(g.ncols, TX), nouter);
  BR_LAUNCH_NAMED(ctx, DIR > 0 ? "fft_cols_kernel<fwd>" : "fft_cols_kernel<inv>", (fft_cols_kernel<N, DIR, TX>), grid,
                  32 * TX, (tile_bytes<N, TX>()), st, in, out, g, tw, herm);
  return BAOREC_OK;
}
template <int N, int DIR>
static int launch_cols(baorec_ctx* ctx, const float2* in, float2* out, const ColGeom& g, int nouter, const float2* tw,
                       int herm, bool z_axis, cudaStream_t st) {
  if (N == 1024) {
    const int tx = tile_cols(ctx, N, z_axis);
    if (tx == 4) return launch_cols_tx<1024, DIR, 4>(ctx, in, out, g, nouter, tw, herm, st);
    if (tx == 16) return launch_cols_tx<1024, DIR, 16>(ctx, in, out, g, nouter, tw, herm, st);
  }
  return launch_cols_tx<N, DIR, FFT_TX>(ctx, in, out, g, nouter, tw, herm, st);
}

static int cols_pass(baorec_ctx* ctx, float2* data, int axis /*1 = y, 2 = z*/, int dir, cudaStream_t st,
                     int herm = 0) {
  ColGeom g = axis == 1 ? geom_y(ctx) : geom_z(ctx);
  const int n = axis == 1 ? ctx->ny : ctx->nz;
  int nouter = axis == 1 ? ctx->nz : ctx->ny;
  if (axis == 2) {  // the columns of a z pass are contiguous across rows: one flat row of xh * ny columns (aligned tiles, none partly empty)
    g.ncols = ctx->xh * ctx->ny;
    nouter = 1;
  }
  const float2* tw = ctx->d_tw[axis - 1];
#define CALL(NN)                                                                    \
  if (dir > 0) BR_TRY((launch_cols<NN, 1>(ctx, data, data, g, nouter, tw, herm, axis == 2, st)));    \
  else BR_TRY((launch_cols<NN, -1>(ctx, data, data, g, nouter, tw, herm, axis == 2, st)));
  FFT_DISPATCH_N(n, CALL)
#undef CALL
  return BAOREC_OK;
}

// z pass of the slab-decomposed layout K[z][yl][x] (multi-GPU, peer-copy exchange): nyl * xh interleaved columns, one
// k-plane of the rank (nyl * xh values) between consecutive z.  Replaces cuFFT's strided 1-D plan where the column
// kernel is the faster one (the 1024-point axis with 16 columns per tile: 2.35 against 3.3 ms per full 1024^3 mesh).
bool own_slab_z_available(const baorec_ctx* ctx) {
  return ctx->opt_own_fft != 0 && ctx->nz == 1024 && ctx->d_tw[1] != nullptr && (ctx->xh * ctx->ny_loc) > 0;
}
int own_slab_z(baorec_ctx* ctx, const float2* in, float2* out, int dir, cudaStream_t st) {
  ColGeom g;
  g.ncols = ctx->xh * ctx->ny_loc;  // flat: the nyl * xh columns of the rank are contiguous
  g.nouter = 1;
  g.stride = (size_t)ctx->xh * ctx->ny_loc;
  g.outer_stride = 0;
  g.kx = ctx->d_k[0];
  g.kouter = ctx->d_k[1];
  g.ktrans = ctx->d_k[2];
  g.outer_is_y = 1;
  g.prefetch = 0;
  if (dir > 0) return launch_cols<1024, 1>(ctx, in, out, g, 1, ctx->d_tw[1], 0, true, st);
  return launch_cols<1024, -1>(ctx, in, out, g, 1, ctx->d_tw[1], 0, true, st);
}

static int x_r2c(baorec_ctx* ctx, const float* in, float2* out, cudaStream_t st) {
  BR_CUFFT(cufftSetStream(ctx->px_r2c, st));
  int pi = prof_begin(ctx, "cufft_1d_x_r2c", st);
  BR_CUFFT(cufftExecR2C(ctx->px_r2c, (cufftReal*)in, (cufftComplex*)out));
  prof_end(ctx, pi, st);
  ctx->n_fft++;
  return BAOREC_OK;
}
static int x_c2r(baorec_ctx* ctx, float2* in, float* out, cudaStream_t st) {
  BR_CUFFT(cufftSetStream(ctx->px_c2r, st));
  int pi = prof_begin(ctx, "cufft_1d_x_c2r", st);
  BR_CUFFT(cufftExecC2R(ctx->px_c2r, (cufftComplex*)in, (cufftReal*)out));
  prof_end(ctx, pi, st);
  ctx->n_fft++;
  return BAOREC_OK;
}

// full transforms (drop-in for the cuFFT 3-D plans; same layout and normalisation)
int own_r2c(baorec_ctx* ctx, const float* in, float2* out, cudaStream_t st) {
  BR_TRY(x_r2c(ctx, in, out, st));
  BR_TRY(cols_pass(ctx, out, 1, +1, st));
  return cols_pass(ctx, out, 2, +1, st);
}
int own_c2r(baorec_ctx* ctx, float2* in, float* out, cudaStream_t st) {
  BR_TRY(cols_pass(ctx, in, 2, -1, st));
  BR_TRY(cols_pass(ctx, in, 1, -1, st, 1));
  return x_c2r(ctx, in, out, st);
}

// run! for a periodic box with a fixed line of sight: rho (real, already scattered into `mesh`)
// -> delta_final in `mesh`; `keep` (may be NULL) receives delta_final_k.
int own_fused_los_solve(baorec_ctx* ctx, const baorec_params* p, float* mesh, float2* work, float2* keep,
                        cudaStream_t st) {
  BR_TRY(x_r2c(ctx, mesh, work, st));
  BR_TRY(cols_pass(ctx, work, 1, +1, st));
  BR_LAUNCH(ctx, dc_from_xy_kernel, 1, 256, 0, st, work, (size_t)ctx->xh * ctx->ny, ctx->nz, ctx->d_scal, 0,
            (double)ctx->M / (double)p->bias);
  OpLosSolve op;
  BR_TRY(gauss_tables(ctx, p->smoothing_radius, &op.gt.gx, &op.gt.gy, &op.gt.gz, st));
  op.scal = ctx->d_scal;
  for (int a = 0; a < 3; a++) op.los[a] = p->los[a];
  op.beta = p->beta;
  op.n_iter = p->n_iter;
  op.invM = (float)(1.0 / (double)ctx->M);
  const ColGeom g = geom_z(ctx);
#define CALL(NN, TX)                                                                                               \
  {                                                                                                                \
    dim3 grid(cdiv(g.ncols, TX), ctx->ny);                                                                         \
    BR_TRY(set_smem((fft_z_solve_kernel<NN, TX>), (fused_bytes<NN, TX>(NN * 12))));                                \
    BR_LAUNCH_NAMED(ctx, "fft_z_solve_kernel", (fft_z_solve_kernel<NN, TX>), grid, 32 * TX, (fused_bytes<NN, TX>(NN * 12)), st, \
                    work, keep, g, ctx->d_tw[1], op);                                                              \
  }
  switch (ctx->nz) {
    case 1024: {
      const int tx = tile_cols(ctx, 1024, true);
      if (tx == 4) CALL(1024, 4) else if (tx == 16) CALL(1024, 16) else CALL(1024, FFT_TX)
    } break;
    case 2048: CALL(2048, FFT_TX) break;
    default: set_error("own FFT: the fused z pass needs nz = 1024 or 2048"); return BAOREC_ERR_INVALID;
  }
#undef CALL
  BR_TRY(cols_pass(ctx, work, 1, -1, st, 1));
  return x_c2r(ctx, work, mesh, st);
}

// displacement meshes: from the real mesh (from_k == nullptr) or from a cached delta_k
int own_displacements(baorec_ctx* ctx, const float* mesh, const float2* from_k, int algorithm, float2* w0, float2* w1,
                      float2* w2, float* px, float* py, float* pz, cudaStream_t st) {
  OpDisp op;
  op.potential = algorithm == BAOREC_MULTIGRID;
  op.invM = (float)(1.0 / (double)ctx->M);
  const ColGeom g = geom_z(ctx);
  const float2* src = from_k;
  if (!from_k) {  // delta_k (phi_k) first: the z kernel starts in k space
    BR_TRY(x_r2c(ctx, mesh, w0, st));
    BR_TRY(cols_pass(ctx, w0, 1, +1, st));
    BR_TRY(cols_pass(ctx, w0, 2, +1, st));
    src = w0;
  }
#define CALL_TX(NN, TX)                                                                                            \
  {                                                                                                                \
    dim3 grid_tx(cdiv((size_t)g.ncols * g.nouter, TX), 1);                                                         \
    BR_TRY(set_smem((fft_z_disp_kernel<NN, TX>), (fused_bytes<NN, TX>(NN * 4))));                                  \
    BR_LAUNCH_NAMED(ctx, "fft_z_disp_kernel", (fft_z_disp_kernel<NN, TX>), grid_tx, 32 * TX, (fused_bytes<NN, TX>(NN * 4)), st, \
                    src, w0, w1, w2, g, ctx->d_tw[1], op);                                                         \
  }
#define CALL(NN) CALL_TX(NN, FFT_TX)
  if (ctx->nz == 1024 && tile_cols(ctx, 1024, true) == 4) CALL_TX(1024, 4)
  else if (ctx->nz == 1024 && tile_cols(ctx, 1024, true) == 16) CALL_TX(1024, 16)
  else {
    FFT_DISPATCH_N(ctx->nz, CALL)
  }
#undef CALL
#undef CALL_TX
  float2* w[3] = {w0, w1, w2};
  float* o[3] = {px, py, pz};
  for (int c = 0; c < 3; c++) {
    BR_TRY(cols_pass(ctx, w[c], 1, -1, st, 1));
    BR_TRY(x_c2r(ctx, w[c], o[c], st));
  }
  return BAOREC_OK;
}

}  // namespace baorec
