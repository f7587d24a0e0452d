// Strided batched 1-D FFT passes (the y and z axes of the 3-D transform) with the k-space
// operators folded into the z pass.
//
// cuFFT's 3-D R2C/C2R at 1024^3 spends 2 x 2.5 ms in its two strided passes (3.4 TB/s) and the
// k-space multiplies then cost another read + write of the half-mesh.  Here the x axis stays a
// batched contiguous cuFFT 1-D transform (already at ~5.3 TB/s) and the y / z axes are our own
// column kernel:
//   * a CTA stages a tile of 8 adjacent columns x N rows in shared memory with coalesced
//     64-byte row segments (the columns are contiguous in memory), one warp per column;
//   * the warp runs the length-N transform as a 32 x M four-step FFT: an M-point DFT in each
//     lane's registers, the W_N^(lane*q) twiddle, and a 32-point DFT across the lanes with
//     __shfl_xor butterflies -- no shared-memory traffic inside the transform;
//   * the z pass applies the k-space operator while the column sits in registers and runs the
//     inverse z transform in the same kernel (forward + operator + inverse = ONE read and ONE
//     write of the half-mesh instead of four passes), optionally emitting three outputs
//     (the displacement components) from one read.
#include <math.h>

#include "internal.cuh"

namespace baorec {

constexpr int FFT_TX = 8;          // columns per tile (8 x 8 B = 64 B row segments)
constexpr int FFT_THREADS = 256;   // 8 warps = 8 columns

__host__ __device__ constexpr int ilog2(int n) { return n <= 1 ? 0 : 1 + ilog2(n >> 1); }
__host__ __device__ constexpr int bitrev(int v, int bits) {
  int r = 0;
  for (int i = 0; i < bits; i++) r |= ((v >> i) & 1) << (bits - 1 - i);
  return r;
}

__device__ __forceinline__ float2 cmul(float2 a, float2 b) {
  return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
template <int DIR>
__device__ __forceinline__ float2 twd(const float2* __restrict__ tw, int idx) {
  float2 w = __ldg(tw + idx);  // exp(-2 pi i idx / N)
  if (DIR < 0) w.y = -w.y;
  return w;
}

// Length-N transform of one column held by a warp.
//   in : lane l holds x[l + 32 j] in v[j], j = 0..M-1      (M = N / 32)
//   out: lane l holds X[M * bitrev5(l) + q] in v[q], q = 0..M-1
// tw = exp(-2 pi i k / N), k = 0..N-1.   DIR = +1 forward (e^-), -1 inverse (e^+), unnormalised.
template <int N, int DIR>
__device__ __forceinline__ void warp_fft(float2 (&v)[N / 32], const float2* __restrict__ tw, int lane) {
  constexpr int M = N / 32;
  constexpr int LM = ilog2(M);
  // step 1: M-point DIF DFT over j in registers (result in bit-reversed register order)
#pragma unroll
  for (int half = M / 2; half >= 1; half >>= 1) {
#pragma unroll
    for (int t = 0; t < half; t++) {
      const float2 w = twd<DIR>(tw, t * (N / (2 * half)));  // W_{2 half}^t
#pragma unroll
      for (int g0 = 0; g0 < M; g0 += 2 * half) {
        float2 a = v[g0 + t], b = v[g0 + t + half];
        v[g0 + t] = cadd(a, b);
        float2 d = csub(a, b);
        v[g0 + t + half] = t == 0 ? d : cmul(d, w);
      }
    }
  }
  float2 y[M];
#pragma unroll
  for (int q = 0; q < M; q++) y[q] = v[bitrev(q, LM)];
  // step 2: twiddle W_N^(lane q).  Only the log2(M) anchors W_N^(lane 2^b) are loaded (lane-strided
  // table reads cost one L1 wavefront per distinct line); the other powers are composed from them
  // along the binary digits of q (at most log2(M)-1 products, <= ~3 ulp).
  {
    float2 anchor[LM > 0 ? LM : 1];
#pragma unroll
    for (int b = 0; b < LM; b++) anchor[b] = twd<DIR>(tw, (lane << b) & (N - 1));
#pragma unroll
    for (int q = 1; q < M; q++) {
      float2 w = make_float2(1.f, 0.f);
      bool first = true;
#pragma unroll
      for (int b = 0; b < LM; b++) {
        if (q & (1 << b)) {
          w = first ? anchor[b] : cmul(w, anchor[b]);
          first = false;
        }
      }
      y[q] = cmul(y[q], w);
    }
  }
  // step 3: 32-point DIF DFT across lanes (result for output r lands in lane bitrev5(r))
#pragma unroll
  for (int half = 16; half >= 1; half >>= 1) {
    const bool upper = (lane & half) != 0;
    const float2 w = twd<DIR>(tw, (lane & (half - 1)) * (N / (2 * half)));  // W_{2 half}^(lane mod half)
#pragma unroll
    for (int q = 0; q < M; q++) {
      float2 o;
      o.x = __shfl_xor_sync(0xffffffffu, y[q].x, half);
      o.y = __shfl_xor_sync(0xffffffffu, y[q].y, half);
      y[q] = upper ? cmul(csub(o, y[q]), w) : cadd(y[q], o);
    }
  }
#pragma unroll
  for (int q = 0; q < M; q++) v[q] = y[q];
}

struct ColGeom {
  int ncols;            // length of the contiguous axis (nx/2+1)
  size_t stride;        // elements between consecutive points of a transform
  size_t outer_stride;  // elements between consecutive tiles rows (blockIdx.y)
  const float* kx;      // k tables: contiguous axis, outer axis, transform axis
  const float* kouter;
  const float* ktrans;
  int outer_is_y;       // 1: outer = y, transform = z (z pass); 0: outer = z, transform = y (y pass)
};

// shared-memory tile: physical row = f + f/32 (one pad row per 32) so that both the strided
// per-lane reads (rows l + 32 j) and the per-lane chunk writes (rows M r + q) are conflict-free
template <int N>
struct Tile {
  static constexpr int ROWS = N + N / 32;
  float2 s[ROWS][FFT_TX + 1];
  __device__ __forceinline__ float2& at(int f, int c) { return s[f + (f >> 5)][c]; }
};

// Fills the tile with 8-byte cp.async (LDGSTS): every thread has all of its N/32 row segments in
// flight at once and no registers are spent on staging.  Call tile_load_wait() before the barrier.
template <int N>
__device__ __forceinline__ void tile_load(Tile<N>& T, const float2* __restrict__ src, size_t stride, int ncol) {
  const int c = threadIdx.x & (FFT_TX - 1), r0 = threadIdx.x / FFT_TX;  // 32 rows x 8 columns per sweep
  if (c < ncol) {
#pragma unroll
    for (int r = r0; r < N; r += FFT_THREADS / FFT_TX) {
      unsigned d = (unsigned)__cvta_generic_to_shared(&T.at(r, c));
      asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(d), "l"(src + (size_t)r * stride + c));
    }
  }
  asm volatile("cp.async.wait_all;" ::: "memory");
}
template <int N>
__device__ __forceinline__ void tile_store(Tile<N>& T, float2* __restrict__ dst, size_t stride, int ncol,
                                           bool drop_imag = false) {
  const int c = threadIdx.x & (FFT_TX - 1), r0 = threadIdx.x / FFT_TX;
  if (c < ncol) {
#pragma unroll 8
    for (int r = r0; r < N; r += FFT_THREADS / FFT_TX) {
      float2 v = T.at(r, c);
      if (drop_imag) v.y = 0.f;
      dst[(size_t)r * stride + c] = v;
    }
  }
}
template <int N>
__device__ __forceinline__ void col_read(Tile<N>& T, int col, int lane, float2 (&v)[N / 32]) {
#pragma unroll
  for (int j = 0; j < N / 32; j++) v[j] = T.at(lane + 32 * j, col);
}
// after warp_fft lane l holds indices (N/32) * bitrev5(l) + q
template <int N>
__device__ __forceinline__ void col_write(Tile<N>& T, int col, int lane, const float2 (&v)[N / 32]) {
  const int base = (N / 32) * (int)(__brev((unsigned)lane) >> 27);
#pragma unroll
  for (int q = 0; q < N / 32; q++) T.at(base + q, col) = v[q];
}

// ---- operators applied in the z pass (frequency index f along the transform axis) ----------------
struct OpNone {
  static constexpr int NOUT = 1;
  static constexpr bool HAS_OP = false;
  __device__ __forceinline__ float2 apply(int, float2 v, float, float, float, bool) const { return v; }
};

// smoothing + (rho/mean - 1)/bias + all n_iter fixed-LOS iterations (see FusedLosOp in kspace.cu);
// returns delta_final_k / M; `keep` receives the unnormalised delta_final_k
struct OpLosSolve {
  static constexpr int NOUT = 1;
  static constexpr bool HAS_OP = true;
  float R2;
  const double* scal;  // scal[8] = M / (A0 bias)
  float los[3];
  float beta;
  int n_iter;
  float invM;
  __device__ __forceinline__ float2 solve(float2 v, float kx, float ky, float kz, bool is_dc) const {
    float k2 = __fadd_rn(__fadd_rn(__fmul_rn(kx, kx), __fmul_rn(ky, ky)), __fmul_rn(kz, kz));
    double s = exp(-0.5 * (double)R2 * (double)k2) * __ldg(scal + 8);
    if (is_dc) s = 0.0;
    float dsx = (float)((double)v.x * s), dsy = (float)((double)v.y * s);
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
      float k2 = __fadd_rn(__fadd_rn(__fmul_rn(kx, kx), __fmul_rn(ky, ky)), __fmul_rn(kz, kz));
      s = k2 > 0.f ? __fdiv_rn(invM, k2) : 0.f;
    }
    float kc = c == 0 ? kx : (c == 1 ? ky : kz);
    return make_float2(__fmul_rn(__fmul_rn(-v.y, s), kc), __fmul_rn(__fmul_rn(v.x, s), kc));
  }
};

// ---- kernels -----------------------------------------------------------------------------------------
// plain pass: out = FFT_DIR(in) along the strided axis (in place allowed)
template <int N, int DIR>
__global__ void __launch_bounds__(FFT_THREADS)
fft_cols_kernel(const float2* __restrict__ in, float2* __restrict__ out, ColGeom cg, const float2* __restrict__ tw,
                int hermitian_edges) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  Tile<N>& T = *reinterpret_cast<Tile<N>*>(smem_raw);
  const int x0 = blockIdx.x * FFT_TX;
  const int ncol = min(FFT_TX, cg.ncols - x0);
  const size_t base = (size_t)blockIdx.y * cg.outer_stride + x0;
  tile_load<N>(T, in + base, cg.stride, ncol);
  __syncthreads();
  const int lane = threadIdx.x & 31, col = threadIdx.x >> 5;
  if (col < ncol) {
    float2 v[N / 32];
    col_read<N>(T, col, lane, v);
    warp_fft<N, DIR>(v, tw, lane);
    __syncwarp();
    col_write<N>(T, col, lane, v);
  }
  __syncthreads();
  // Last pass before the C2R along x: the kx = 0 and kx = Nyquist columns must be real there.  FFTW
  // and pocketfft ignore their imaginary part; cuFFT's 1-D C2R does not, so it is dropped here
  // (it is non-zero only for inputs with power at the Nyquist modes, e.g. i k multiplications).
  const int cme = threadIdx.x & (FFT_TX - 1);
  const bool drop = hermitian_edges && (x0 + cme == 0 || x0 + cme == cg.ncols - 1);
  tile_store<N>(T, out + base, cg.stride, ncol, drop);
}

// z pass of the fused fixed-LOS solve: forward z FFT, operator, inverse z FFT; in place.
template <int N>
__global__ void __launch_bounds__(FFT_THREADS)
fft_z_solve_kernel(float2* __restrict__ data, float2* __restrict__ keep, ColGeom cg, const float2* __restrict__ tw,
                   OpLosSolve op) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  Tile<N>& T = *reinterpret_cast<Tile<N>*>(smem_raw);
  constexpr int M = N / 32;
  const int x0 = blockIdx.x * FFT_TX;
  const int ncol = min(FFT_TX, cg.ncols - x0);
  const int iy = blockIdx.y;
  const size_t base = (size_t)iy * cg.outer_stride + x0;
  tile_load<N>(T, data + base, cg.stride, ncol);
  __syncthreads();
  const int lane = threadIdx.x & 31, col = threadIdx.x >> 5;
  float2 v[M];
  if (col < ncol) {
    col_read<N>(T, col, lane, v);
    warp_fft<N, 1>(v, tw, lane);
    const int ix = x0 + col;
    const float kx = __ldg(cg.kx + ix), ky = __ldg(cg.kouter + iy);
    const int f0 = M * (int)(__brev((unsigned)lane) >> 27);
#pragma unroll
    for (int q = 0; q < M; q++) {
      const int f = f0 + q;
      v[q] = op.solve(v[q], kx, ky, __ldg(cg.ktrans + f), (ix | iy | f) == 0);
    }
    __syncwarp();
    col_write<N>(T, col, lane, v);
  }
  if (keep != nullptr) {  // uniform branch: delta_final_k for the read-back cache
    __syncthreads();
    tile_store<N>(T, keep + base, cg.stride, ncol);
    __syncthreads();
  } else {
    __syncwarp();
  }
  if (col < ncol) {
    col_read<N>(T, col, lane, v);
#pragma unroll
    for (int q = 0; q < M; q++) v[q] = make_float2(v[q].x * op.invM, v[q].y * op.invM);
    warp_fft<N, -1>(v, tw, lane);
    __syncwarp();
    col_write<N>(T, col, lane, v);
  }
  __syncthreads();
  tile_store<N>(T, data + base, cg.stride, ncol);
}

// z pass of the displacement read-back: (forward z FFT of `in` unless FROM_K) then for each of the
// three components operator + inverse z FFT into out[c].  One read, three writes.
template <int N, bool FROM_K>
__global__ void __launch_bounds__(FFT_THREADS)
fft_z_disp_kernel(const float2* __restrict__ in, float2* __restrict__ o0, float2* __restrict__ o1,
                  float2* __restrict__ o2, ColGeom cg, const float2* __restrict__ tw, OpDisp op) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  Tile<N>& T = *reinterpret_cast<Tile<N>*>(smem_raw);
  constexpr int M = N / 32;
  const int x0 = blockIdx.x * FFT_TX;
  const int ncol = min(FFT_TX, cg.ncols - x0);
  const int iy = blockIdx.y;
  const size_t base = (size_t)iy * cg.outer_stride + x0;
  tile_load<N>(T, in + base, cg.stride, ncol);
  __syncthreads();
  const int lane = threadIdx.x & 31, col = threadIdx.x >> 5;
  const int f0 = M * (int)(__brev((unsigned)lane) >> 27);
  float2 dk[M];  // delta_k (or phi_k) of this column, frequency f0 + q
  float kx = 0.f, ky = 0.f;
  if (col < ncol) {
    if (FROM_K) {
#pragma unroll
      for (int q = 0; q < M; q++) dk[q] = T.at(f0 + q, col);
    } else {
      col_read<N>(T, col, lane, dk);
      warp_fft<N, 1>(dk, tw, lane);
    }
    kx = __ldg(cg.kx + x0 + col);
    ky = __ldg(cg.kouter + iy);
  }
#pragma unroll 1
  for (int c = 0; c < 3; c++) {
    float2* outc = c == 0 ? o0 : (c == 1 ? o1 : o2);
    __syncthreads();  // previous tile contents fully consumed / stored
    if (col < ncol) {
      float2 v[M];
#pragma unroll
      for (int q = 0; q < M; q++) T.at(f0 + q, col) = op.comp(dk[q], kx, ky, __ldg(cg.ktrans + f0 + q), c);
      __syncwarp();
      col_read<N>(T, col, lane, v);
      warp_fft<N, -1>(v, tw, lane);
      __syncwarp();
      col_write<N>(T, col, lane, v);
    }
    __syncthreads();
    tile_store<N>(T, outc + base, cg.stride, ncol);
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
static bool pow2_ok(int n) { return n >= 64 && n <= 2048 && (n & (n - 1)) == 0; }

bool own_fft_available(const baorec_ctx* ctx) {
  return ctx->opt_own_fft && ctx->have_x_plans && pow2_ok(ctx->ny) && pow2_ok(ctx->nz) && ctx->d_tw[0] && ctx->d_tw[1];
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

template <int N>
static size_t tile_bytes() { return sizeof(Tile<N>); }

#define FFT_DISPATCH_N(n, CALL)     \
  switch (n) {                      \
    case 64: { CALL(64); } break;   \
    case 128: { CALL(128); } break; \
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
  g.stride = ctx->xh;
  g.outer_stride = (size_t)ctx->xh * ctx->ny;
  g.kx = ctx->d_k[0];
  g.kouter = ctx->d_k[2];
  g.ktrans = ctx->d_k[1];
  g.outer_is_y = 0;
  return g;
}
static ColGeom geom_z(const baorec_ctx* ctx) {  // transform along z, tiles over (x, y)
  ColGeom g;
  g.ncols = ctx->xh;
  g.stride = (size_t)ctx->xh * ctx->ny;
  g.outer_stride = ctx->xh;
  g.kx = ctx->d_k[0];
  g.kouter = ctx->d_k[1];
  g.ktrans = ctx->d_k[2];
  g.outer_is_y = 1;
  return g;
}

template <int N, int DIR>
static int launch_cols(baorec_ctx* ctx, const float2* in, float2* out, const ColGeom& g, int nouter, const float2* tw,
                       int herm, cudaStream_t st) {
  BR_TRY(set_smem(fft_cols_kernel<N, DIR>, tile_bytes<N>()));
  dim3 grid(cdiv(g.ncols, FFT_TX), nouter);
  BR_LAUNCH_NAMED(ctx, DIR > 0 ? "fft_cols_kernel<fwd>" : "fft_cols_kernel<inv>", (fft_cols_kernel<N, DIR>), grid,
                  FFT_THREADS, tile_bytes<N>(), st, in, out, g, tw, herm);
  return BAOREC_OK;
}

static int cols_pass(baorec_ctx* ctx, float2* data, int axis /*1 = y, 2 = z*/, int dir, cudaStream_t st,
                     int herm = 0) {
  const ColGeom g = axis == 1 ? geom_y(ctx) : geom_z(ctx);
  const int n = axis == 1 ? ctx->ny : ctx->nz;
  const int nouter = axis == 1 ? ctx->nz : ctx->ny;
  const float2* tw = ctx->d_tw[axis - 1];
#define CALL(NN)                                                                    \
  if (dir > 0) BR_TRY((launch_cols<NN, 1>(ctx, data, data, g, nouter, tw, herm, st)));    \
  else BR_TRY((launch_cols<NN, -1>(ctx, data, data, g, nouter, tw, herm, st)));
  FFT_DISPATCH_N(n, CALL)
#undef CALL
  return BAOREC_OK;
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
  op.R2 = p->smoothing_radius * p->smoothing_radius;
  op.scal = ctx->d_scal;
  for (int a = 0; a < 3; a++) op.los[a] = p->los[a];
  op.beta = p->beta;
  op.n_iter = p->n_iter;
  op.invM = (float)(1.0 / (double)ctx->M);
  const ColGeom g = geom_z(ctx);
  dim3 grid(cdiv(g.ncols, FFT_TX), ctx->ny);
#define CALL(NN)                                                                                          \
  BR_TRY(set_smem(fft_z_solve_kernel<NN>, tile_bytes<NN>()));                                             \
  BR_LAUNCH_NAMED(ctx, "fft_z_solve_kernel", fft_z_solve_kernel<NN>, grid, FFT_THREADS, tile_bytes<NN>(), st, work, \
                  keep, g, ctx->d_tw[1], op);
  FFT_DISPATCH_N(ctx->nz, CALL)
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
  dim3 grid(cdiv(g.ncols, FFT_TX), ctx->ny);
  const float2* src = from_k;
  if (!from_k) {
    BR_TRY(x_r2c(ctx, mesh, w0, st));
    BR_TRY(cols_pass(ctx, w0, 1, +1, st));
    src = w0;
  }
#define CALL(NN)                                                                                               \
  if (from_k) {                                                                                                \
    BR_TRY(set_smem(fft_z_disp_kernel<NN, true>, tile_bytes<NN>()));                                           \
    BR_LAUNCH_NAMED(ctx, "fft_z_disp_kernel<from_k>", (fft_z_disp_kernel<NN, true>), grid, FFT_THREADS,         \
                    tile_bytes<NN>(), st, src, w0, w1, w2, g, ctx->d_tw[1], op);                                \
  } else {                                                                                                     \
    BR_TRY(set_smem(fft_z_disp_kernel<NN, false>, tile_bytes<NN>()));                                          \
    BR_LAUNCH_NAMED(ctx, "fft_z_disp_kernel<from_mesh>", (fft_z_disp_kernel<NN, false>), grid, FFT_THREADS,     \
                    tile_bytes<NN>(), st, src, w0, w1, w2, g, ctx->d_tw[1], op);                                \
  }
  FFT_DISPATCH_N(ctx->nz, CALL)
#undef CALL
  float2* w[3] = {w0, w1, w2};
  float* o[3] = {px, py, pz};
  for (int c = 0; c < 3; c++) {
    BR_TRY(cols_pass(ctx, w[c], 1, -1, st, 1));
    BR_TRY(x_c2r(ctx, w[c], o[c], st));
  }
  return BAOREC_OK;
}

}  // namespace baorec
