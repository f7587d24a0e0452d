// k-space and real-space mesh passes around the cuFFT R2C/C2R transforms:
// Gaussian smoothing, overdensity set-up (box and randoms), the RSD iteration
// (iterate!), and the displacement meshes.  Every pass is a single streaming
// kernel (one read + one write of each operand); the 1/M normalisation of the
// inverse transform (`ldiv!`) is folded into the k-space multiply.
#include "internal.cuh"
#include "kspace_ops.cuh"

namespace baorec {

struct KGeom {
  const float* kx;
  const float* ky;
  const float* kz;
  int xh, ny, nz;
  int y0;  // global index of the first row held (0 on one GPU; the y slab's offset in the distributed K[z][yl][x] layout)
};

static KGeom kgeom_of(const baorec_ctx* ctx) {
  KGeom g;
  g.kx = ctx->d_k[0];
  g.ky = ctx->d_k[1];
  g.kz = ctx->d_k[2];
  g.xh = ctx->xh;
  g.ny = ctx->ny;
  g.nz = ctx->nz;
  g.y0 = 0;
  return g;
}

constexpr int KS_THREADS = 256;
constexpr int KS_UNROLL = 4;

// One thread handles KS_UNROLL complex values of one z-plane (loads issued back to back,
// then the stores), blockIdx.y = iz.  Op::apply(idx, v, kx, ky, kz, dc) writes its outputs.
template <class Op>
__global__ void __launch_bounds__(KS_THREADS) kspace_kernel(KGeom g, const float2* __restrict__ in, Op op) {
  const int iz = blockIdx.y;
  const unsigned plane = (unsigned)g.xh * (unsigned)g.ny;
  const float kz = __ldg(g.kz + iz);
  const unsigned base = blockIdx.x * (KS_THREADS * KS_UNROLL) + threadIdx.x;
  const size_t off = (size_t)iz * plane;
  float2 v[KS_UNROLL];
#pragma unroll
  for (int u = 0; u < KS_UNROLL; u++) {
    unsigned p = base + u * KS_THREADS;
    if (p < plane) v[u] = in[off + p];
  }
#pragma unroll
  for (int u = 0; u < KS_UNROLL; u++) {
    unsigned p = base + u * KS_THREADS;
    if (p < plane) {
      unsigned iy = p / (unsigned)g.xh;
      unsigned ix = p - iy * (unsigned)g.xh;
      iy += (unsigned)g.y0;  // g.ny rows are held, rows y0 .. y0 + ny - 1 of the mesh
      op.apply(off + p, v[u], __ldg(g.kx + ix), __ldg(g.ky + iy), kz, (ix | iy | (unsigned)iz) == 0u, (int)ix, (int)iy, iz);
    }
  }
}

// scal[slot] = Re A_k[0] (= sum of the real mesh); scal[slot + 8] = mul / Re A_k[0]
__global__ void stash_dc_kernel(const float2* __restrict__ ck, double* scal, int slot, double mul) {
  double a0 = (double)ck[0].x;
  scal[slot] = a0;
  scal[slot + 8] = mul / a0;
}

static int gauss_tab(baorec_ctx* ctx, float R, GaussTab* t, cudaStream_t st) {
  return gauss_tables(ctx, R, &t->gx, &t->gy, &t->gz, st);
}

template <class Op>
static int run_kspace(baorec_ctx* ctx, const float2* in, Op op, cudaStream_t st) {
  size_t plane = (size_t)ctx->xh * ctx->ny;
  dim3 grid(cdiv(plane, KS_THREADS * KS_UNROLL), ctx->nz);
  BR_LAUNCH_NAMED(ctx, Op::name(), kspace_kernel<Op>, grid, KS_THREADS, 0, st, kgeom_of(ctx), in, op);
  return BAOREC_OK;
}

// ---- real-space passes -------------------------------------------------------------------------
// delta_r = src - fac * X    (src/iterative.jl:59), 4 cells per thread
__global__ void __launch_bounds__(256)
axpy4_kernel(float4* __restrict__ out, const float4* __restrict__ src, const float4* __restrict__ X, float fac,
             size_t n4) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  size_t stride = (size_t)gridDim.x * blockDim.x;
  for (; i < n4; i += stride) {
    float4 s = src[i], x = X[i];
    out[i] = make_float4(__fsub_rn(s.x, __fmul_rn(fac, x.x)), __fsub_rn(s.y, __fmul_rn(fac, x.y)),
                         __fsub_rn(s.z, __fmul_rn(fac, x.z)), __fsub_rn(s.w, __fmul_rn(fac, x.w)));
  }
}
__global__ void axpy1_kernel(float* __restrict__ out, const float* __restrict__ src, const float* __restrict__ X,
                             float fac, size_t n) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  size_t stride = (size_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) out[i] = __fsub_rn(src[i], __fmul_rn(fac, X[i]));
}

// radial update (src/iterative.jl:32-35): delta_r = x²>0 ? src - fac X x_i x_j / x² : 0
__global__ void __launch_bounds__(256)
radial_update_kernel(float* out, const float* src, const float* __restrict__ X,
                     const float* __restrict__ xv0, const float* __restrict__ xv1, const float* __restrict__ xv2,
                     int nx, int ny, int ci, int cj, float fac) {
  const int iz = blockIdx.y;
  const unsigned plane = (unsigned)nx * (unsigned)ny;
  const float zc = __ldg(xv2 + iz);
  const size_t off = (size_t)iz * plane;
  for (unsigned p = blockIdx.x * blockDim.x + threadIdx.x; p < plane; p += gridDim.x * blockDim.x) {
    unsigned iy = p / (unsigned)nx, ix = p - iy * (unsigned)nx;
    float xc = __ldg(xv0 + ix), yc = __ldg(xv1 + iy);
    float x2 = __fadd_rn(__fadd_rn(__fmul_rn(xc, xc), __fmul_rn(yc, yc)), __fmul_rn(zc, zc));
    float a = ci == 0 ? xc : (ci == 1 ? yc : zc);
    float b = cj == 0 ? xc : (cj == 1 ? yc : zc);
    float s = src[off + p], x = X[off + p];
    float t = __fdiv_rn(__fmul_rn(__fmul_rn(__fmul_rn(fac, x), a), b), x2);
    out[off + p] = x2 > 0.f ? __fsub_rn(s, t) : 0.f;
  }
}

// Three pair terms of the radial update in ONE pass (the six of an iteration in two): reads src and three X meshes,
// writes out -- 20 B/cell instead of 3 x 12.  Every cell runs the reference's three sequential updates
// (src/iterative.jl:32-35) in registers, so the result is bit-identical to three radial_update_kernel launches.
struct Pair3 {
  int i[3], j[3];
  float fac[3];
};
__global__ void __launch_bounds__(256)
radial_update3_kernel(float* out, const float* src, const float* __restrict__ X0, const float* __restrict__ X1,
                      const float* __restrict__ X2, const float* __restrict__ xv0, const float* __restrict__ xv1,
                      const float* __restrict__ xv2, int nx, int ny, Pair3 pr) {
  const int iz = blockIdx.y;
  const unsigned nx4 = (unsigned)nx >> 2, plane4 = nx4 * (unsigned)ny;
  const float zc = __ldg(xv2 + iz);
  const size_t off = (size_t)iz * nx * ny;
  for (unsigned p = blockIdx.x * blockDim.x + threadIdx.x; p < plane4; p += gridDim.x * blockDim.x) {
    const unsigned iy = p / nx4, ix = (p - iy * nx4) << 2;
    const float yc = __ldg(xv1 + iy);
    const size_t c = off + (size_t)iy * nx + ix;
    const float4 s4 = *reinterpret_cast<const float4*>(src + c);
    const float4 a4 = *reinterpret_cast<const float4*>(X0 + c), b4 = *reinterpret_cast<const float4*>(X1 + c),
                 c4 = *reinterpret_cast<const float4*>(X2 + c);
    const float4 x4 = *reinterpret_cast<const float4*>(xv0 + ix);
    const float sv[4] = {s4.x, s4.y, s4.z, s4.w}, xs[4] = {x4.x, x4.y, x4.z, x4.w};
    const float Xv[3][4] = {{a4.x, a4.y, a4.z, a4.w}, {b4.x, b4.y, b4.z, b4.w}, {c4.x, c4.y, c4.z, c4.w}};
    float o[4];
#pragma unroll
    for (int q = 0; q < 4; q++) {
      const float xc = xs[q];
      const float x2 = __fadd_rn(__fadd_rn(__fmul_rn(xc, xc), __fmul_rn(yc, yc)), __fmul_rn(zc, zc));
      float d = sv[q];
#pragma unroll
      for (int k = 0; k < 3; k++) {
        const float a = pr.i[k] == 0 ? xc : (pr.i[k] == 1 ? yc : zc);
        const float b = pr.j[k] == 0 ? xc : (pr.j[k] == 1 ? yc : zc);
        const float t = __fdiv_rn(__fmul_rn(__fmul_rn(__fmul_rn(pr.fac[k], Xv[k][q]), a), b), x2);
        d = x2 > 0.f ? __fsub_rn(d, t) : 0.f;
      }
      o[q] = d;
    }
    *reinterpret_cast<float4*>(out + c) = make_float4(o[0], o[1], o[2], o[3]);
  }
}

// randoms combine (src/recon.jl:78-85)
__global__ void __launch_bounds__(256)
randoms_combine_kernel(float* __restrict__ out, const float* __restrict__ dat, const float* __restrict__ ran,
                       const double* __restrict__ scal, float bias, double ran_min, double n_ran, size_t n) {
  const float sd = (float)scal[0], sr = (float)scal[1];
  const float alpha = __fdiv_rn(sd, sr);
  if (n_ran < 0.0) n_ran = scal[2];  // distributed: global randoms count, all-reduced into scal[2]
  const double thr = ran_min * (double)sr / n_ran;  // Float64 like the reference (0.01 literal, src/recon.jl:70,81)
  const float ba = __fmul_rn(bias, alpha);
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  size_t stride = (size_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) {
    float d = dat[i], r = ran[i];
    float num = __fsub_rn(d, __fmul_rn(alpha, r));
    out[i] = (double)r > thr ? __fdiv_rn(num, __fmul_rn(ba, r)) : 0.f;
  }
}

static unsigned stream_grid(size_t work_items, int threads) {
  size_t g = (work_items + threads - 1) / threads;
  size_t cap = (size_t)148 * 32;
  return (unsigned)(g < cap ? (g ? g : 1) : cap);
}

static int axpy(baorec_ctx* ctx, float* out, const float* src, const float* X, float fac, cudaStream_t st) {
  size_t n = ctx->M;
  bool al = (((uintptr_t)out | (uintptr_t)src | (uintptr_t)X) & 15) == 0;
  if (n % 4 == 0 && al) {
    BR_LAUNCH(ctx, axpy4_kernel, stream_grid(n / 4, 256), 256, 0, st, (float4*)out, (const float4*)src,
              (const float4*)X, fac, n / 4);
  } else {
    BR_LAUNCH(ctx, axpy1_kernel, stream_grid(n, 256), 256, 0, st, out, src, X, fac, n);
  }
  return BAOREC_OK;
}

// ---- drivers -------------------------------------------------------------------------------------
int smooth(baorec_ctx* ctx, float* mesh, float R, cudaStream_t st) {
  float2* ck0;
  BR_TRY(need_t(ctx, BUF_CK0, ctx->Mc, &ck0));
  BR_TRY(fft_r2c(ctx, mesh, ck0, st));
  GaussTab gt;
  BR_TRY(gauss_tab(ctx, R, &gt, st));
  GaussOp op{ck0, gt, 1.0 / (double)ctx->M};
  BR_TRY(run_kspace(ctx, ck0, op, st));
  BR_TRY(fft_c2r(ctx, ck0, mesh, st));
  return BAOREC_OK;
}

// Scatters into `mesh` (zero-filled by the caller) and leaves delta in `delta_out`
// (which may be `mesh` itself).
int setup_overdensity_into(baorec_ctx* ctx, const baorec_params* p, float* mesh, float* delta_out, float* x, float* y,
                           float* z, const float* w, int64_t n, float* rx, float* ry, float* rz, const float* rw,
                           int64_t nr, int wrap, cudaStream_t st) {
  float2* ck0;
  BR_TRY(need_t(ctx, BUF_CK0, ctx->Mc, &ck0));
  GaussTab gt;
  BR_TRY(gauss_tab(ctx, p->smoothing_radius, &gt, st));
  BR_TRY(reset_oob(ctx, st));
  if (nr == 0) {
    BR_TRY(scatter(ctx, mesh, x, y, z, w, n, wrap, p->mas, st));
    BR_TRY(fft_r2c(ctx, mesh, ck0, st));
    BR_LAUNCH(ctx, stash_dc_kernel, 1, 1, 0, st, ck0, ctx->d_scal, 0, 1.0 / (double)p->bias);
    SetupBoxOp op{ck0, gt, p->bias, ctx->d_scal};
    BR_TRY(run_kspace(ctx, ck0, op, st));
    BR_TRY(fft_c2r(ctx, ck0, delta_out, st));
  } else {
    float* ran;
    BR_TRY(need_t(ctx, BUF_RAN, ctx->M, &ran));
    BR_CUDA(cudaMemsetAsync(ran, 0, ctx->M * sizeof(float), st));
    // randoms first: the sort kept for the read-back is then the data catalog's (unified sort)
    BR_TRY(scatter(ctx, ran, rx, ry, rz, rw, nr, 0, p->mas, st));
    BR_TRY(scatter(ctx, mesh, x, y, z, w, n, 0, p->mas, st));
    GaussOp op{ck0, gt, 1.0 / (double)ctx->M};
    BR_TRY(fft_r2c(ctx, mesh, ck0, st));
    BR_LAUNCH(ctx, stash_dc_kernel, 1, 1, 0, st, ck0, ctx->d_scal, 0, 1.0);
    BR_TRY(run_kspace(ctx, ck0, op, st));
    BR_TRY(fft_c2r(ctx, ck0, mesh, st));
    BR_TRY(fft_r2c(ctx, ran, ck0, st));
    BR_LAUNCH(ctx, stash_dc_kernel, 1, 1, 0, st, ck0, ctx->d_scal, 1, 1.0);
    BR_TRY(run_kspace(ctx, ck0, op, st));
    BR_TRY(fft_c2r(ctx, ck0, ran, st));
    BR_LAUNCH(ctx, randoms_combine_kernel, stream_grid(ctx->M, 256), 256, 0, st, delta_out, mesh, ran, ctx->d_scal,
              p->bias, (double)p->ran_min, (double)nr, ctx->M);
  }
  return check_oob(ctx, st, "setup_overdensity");
}

int setup_overdensity(baorec_ctx* ctx, const baorec_params* p, float* mesh, float* x, float* y, float* z,
                      const float* w, int64_t n, float* rx, float* ry, float* rz, const float* rw, int64_t nr, int wrap,
                      cudaStream_t st) {
  return setup_overdensity_into(ctx, p, mesh, mesh, x, y, z, w, n, rx, ry, rz, rw, nr, wrap, st);
}

// One RSD iteration: delta_k from `fft_src` (the current delta_r), result into `delta_r`.
// On iteration 1 the reference has delta_r == delta_s, so the driver passes fft_src = delta_s
// and never materialises `copy(δ_r)` (src/recon.jl:101).
static int iterate_impl(baorec_ctx* ctx, const float* fft_src, float* delta_r, const float* delta_s, int iter,
                        float beta, const float* los, cudaStream_t st) {
  float2* ck0;
  float* X;
  BR_TRY(need_t(ctx, BUF_CK0, ctx->Mc, &ck0));
  BR_TRY(need_t(ctx, BUF_RX, ctx->M, &X));
  ctx->disp_valid = false;
  const float invM = (float)(1.0 / (double)ctx->M);
  BR_TRY(fft_r2c(ctx, fft_src, ck0, st));
  if (los) {
    float fac = beta;
    if (iter == 1) fac = fac / (1.0f + beta);
    IterLosOp op{ck0, {los[0], los[1], los[2]}, invM};
    BR_TRY(run_kspace(ctx, ck0, op, st));
    BR_TRY(fft_c2r(ctx, ck0, X, st));
    BR_TRY(axpy(ctx, delta_r, delta_s, X, fac, st));
  } else {
    float2* ck1;
    BR_TRY(need_t(ctx, BUF_CK1, ctx->Mc, &ck1));
    bool first = true;
    float *X1, *X2;
    BR_TRY(need_t(ctx, BUF_RY, ctx->M, &X1));
    BR_TRY(need_t(ctx, BUF_RZ, ctx->M, &X2));
    const bool vec = ctx->nx % 4 == 0 && ((((uintptr_t)delta_r | (uintptr_t)delta_s | (uintptr_t)X | (uintptr_t)X1 | (uintptr_t)X2) & 15) == 0);
    if (vec) {
      // three pair terms per real-space pass: (xx, xy, xz) then (yy, yz, zz), the reference's order
      float* Xs[3] = {X, X1, X2};
      Pair3 pr;
      int k = 0;
      for (int i = 0; i < 3; i++)
        for (int j = i; j < 3; j++) {
          float fac = (float)((1.0 + (i != j ? 1.0 : 0.0)) * (double)beta);
          if (iter == 1) fac = fac / (1.0f + beta);
          IterPairOp op{ck1, i, j, invM};
          BR_TRY(run_kspace(ctx, ck0, op, st));
          BR_TRY(fft_c2r(ctx, ck1, Xs[k], st));
          pr.i[k] = i;
          pr.j[k] = j;
          pr.fac[k] = fac;
          if (++k == 3) {
            unsigned gx = stream_grid((size_t)(ctx->nx / 4) * ctx->ny, 256);
            dim3 grid(gx > 64 ? 64 : gx, ctx->nz);
            BR_LAUNCH(ctx, radial_update3_kernel, grid, 256, 0, st, delta_r, first ? delta_s : delta_r, Xs[0], Xs[1], Xs[2],
                      ctx->d_xv[0], ctx->d_xv[1], ctx->d_xv[2], ctx->nx, ctx->ny, pr);
            first = false;
            k = 0;
          }
        }
      return BAOREC_OK;
    }
    for (int i = 0; i < 3; i++)
      for (int j = i; j < 3; j++) {
        float fac = (float)((1.0 + (i != j ? 1.0 : 0.0)) * (double)beta);
        if (iter == 1) fac = fac / (1.0f + beta);
        IterPairOp op{ck1, i, j, invM};
        BR_TRY(run_kspace(ctx, ck0, op, st));
        BR_TRY(fft_c2r(ctx, ck1, X, st));
        dim3 grid(stream_grid((size_t)ctx->nx * ctx->ny, 256) > 64 ? 64 : stream_grid((size_t)ctx->nx * ctx->ny, 256),
                  ctx->nz);
        BR_LAUNCH(ctx, radial_update_kernel, grid, 256, 0, st, delta_r, first ? delta_s : delta_r, X, ctx->d_xv[0],
                  ctx->d_xv[1], ctx->d_xv[2], ctx->nx, ctx->ny, i, j, fac);
        first = false;
      }
  }
  return BAOREC_OK;
}

int iterate(baorec_ctx* ctx, float* delta_r, const float* delta_s, int iter, float beta, const float* los,
            cudaStream_t st) {
  return iterate_impl(ctx, delta_r, delta_r, delta_s, iter, beta, los, st);
}

int displacement_meshes(baorec_ctx* ctx, const float* mesh, int algorithm, float* px, float* py, float* pz,
                        cudaStream_t st, bool use_kcache) {
  float2 *ck0, *ck1, *ck2;
  BR_TRY(need_t(ctx, BUF_CK0, ctx->Mc, &ck0));
  BR_TRY(need_t(ctx, BUF_CK1, ctx->Mc, &ck1));
  BR_TRY(need_t(ctx, BUF_CK2, ctx->Mc, &ck2));
  const float invM = (float)(1.0 / (double)ctx->M);
  if (own_fft_available(ctx)) {
    const float2* from_k = (use_kcache && ctx->kcache_valid && algorithm == BAOREC_ITERATIVE)
                               ? (const float2*)ctx->bufs[BUF_CKCACHE].p : nullptr;
    return own_displacements(ctx, mesh, from_k, algorithm, ck0, ck1, ck2, px, py, pz, st);
  }
  const float2* src = ck0;
  if (use_kcache && ctx->kcache_valid && algorithm == BAOREC_ITERATIVE) {
    src = (const float2*)ctx->bufs[BUF_CKCACHE].p;  // delta_k kept by the fused solve: no R2C needed
  } else {
    BR_TRY(fft_r2c(ctx, mesh, ck0, st));
  }
  if (algorithm == BAOREC_MULTIGRID) {
    DispOp<true> op{ck0, ck1, ck2, invM};
    BR_TRY(run_kspace(ctx, src, op, st));
  } else {
    DispOp<false> op{ck0, ck1, ck2, invM};
    BR_TRY(run_kspace(ctx, src, op, st));
  }
  BR_TRY(fft_c2r(ctx, ck0, px, st));
  BR_TRY(fft_c2r(ctx, ck1, py, st));
  BR_TRY(fft_c2r(ctx, ck2, pz, st));
  return BAOREC_OK;
}

int reconstructed_overdensity(baorec_ctx* ctx, const baorec_params* p, float* mesh, float* x, float* y, float* z,
                              const float* w, int64_t n, float* rx, float* ry, float* rz, const float* rw, int64_t nr,
                              cudaStream_t st) {
  ctx->kcache_valid = false;
  ctx->disp_valid = false;
  ctx->mg_result_mesh = nullptr;
  const float* los = p->has_los ? p->los : nullptr;
  const float invM = (float)(1.0 / (double)ctx->M);
  if (los && ctx->opt_fuse_kspace) {
    // Fixed line of sight: scatter -> R2C -> one fused k-space pass (smooth, normalise, all
    // n_iter iterations) -> C2R.  2 transforms instead of 2 + 2 n_iter.
    float2* ck0;
    BR_TRY(need_t(ctx, BUF_CK0, ctx->Mc, &ck0));
    float2* keep = nullptr;
    if (ctx->want_kcache || ctx->opt_keep_delta_k) BR_TRY(need_t(ctx, BUF_CKCACHE, ctx->Mc, &keep));
    if (nr == 0 && own_fft_fused_available(ctx)) {
      // x: cuFFT 1-D; y: column kernel; z: forward + solve + inverse in one kernel; y; x
      BR_TRY(reset_oob(ctx, st));
      BR_TRY(scatter(ctx, mesh, x, y, z, w, n, 1, p->mas, st));
      BR_TRY(own_fused_los_solve(ctx, p, mesh, ck0, keep, st));
      BR_TRY(check_oob(ctx, st, "reconstructed_overdensity"));
    } else if (nr == 0) {
      BR_TRY(reset_oob(ctx, st));
      BR_TRY(scatter(ctx, mesh, x, y, z, w, n, 1, p->mas, st));
      BR_TRY(fft_r2c(ctx, mesh, ck0, st));
      BR_LAUNCH(ctx, stash_dc_kernel, 1, 1, 0, st, ck0, ctx->d_scal, 0, (double)ctx->M / (double)p->bias);
      GaussTab gt;
      BR_TRY(gauss_tab(ctx, p->smoothing_radius, &gt, st));
      FusedLosOp<0> op{ck0, keep, gt, p->bias, ctx->d_scal,
                       {los[0], los[1], los[2]}, p->beta, p->n_iter, invM};
      BR_TRY(run_kspace(ctx, ck0, op, st));
      BR_TRY(fft_c2r(ctx, ck0, mesh, st));
      BR_TRY(check_oob(ctx, st, "reconstructed_overdensity"));
    } else {
      float* ds;
      BR_TRY(need_t(ctx, BUF_RS, ctx->M, &ds));
      BR_TRY(setup_overdensity_into(ctx, p, mesh, ds, x, y, z, w, n, rx, ry, rz, rw, nr, 0, st));
      BR_TRY(fft_r2c(ctx, ds, ck0, st));
      FusedLosOp<1> op{ck0, keep, GaussTab{}, p->bias, ctx->d_scal, {los[0], los[1], los[2]}, p->beta, p->n_iter, invM};
      BR_TRY(run_kspace(ctx, ck0, op, st));
      BR_TRY(fft_c2r(ctx, ck0, mesh, st));
    }
    ctx->kcache_valid = keep != nullptr;
    ctx->kcache_mesh = mesh;
    return BAOREC_OK;
  }
  float* ds;
  BR_TRY(need_t(ctx, BUF_RS, ctx->M, &ds));
  // delta_s lands in RS (no `copy(δ_r)`, src/recon.jl:101); the first iteration reads it directly.
  BR_TRY(setup_overdensity_into(ctx, p, mesh, ds, x, y, z, w, n, rx, ry, rz, rw, nr, 1, st));
  if (p->n_iter <= 0) {
    BR_CUDA(cudaMemcpyAsync(mesh, ds, ctx->M * sizeof(float), cudaMemcpyDeviceToDevice, st));
    return BAOREC_OK;
  }
  for (int it = 1; it <= p->n_iter; it++)
    BR_TRY(iterate_impl(ctx, it == 1 ? ds : mesh, mesh, ds, it, p->beta, los, st));
  return BAOREC_OK;
}

// ---- transposed (distributed) layout -----------------------------------------------------------
// After the slab transpose rank r holds T[yl][ix][iz] (z contiguous) for global rows
// y = y0 + yl.  Same Ops, different index decode.
template <class Op>
__global__ void __launch_bounds__(KS_THREADS)
kspace_kernel_t(KGeom g, int y0, const float2* __restrict__ in, Op op) {
  const int yl = blockIdx.y;
  const unsigned plane = (unsigned)g.xh * (unsigned)g.nz;
  const float ky = __ldg(g.ky + y0 + yl);
  const unsigned base = blockIdx.x * (KS_THREADS * KS_UNROLL) + threadIdx.x;
  const size_t off = (size_t)yl * plane;
  float2 v[KS_UNROLL];
#pragma unroll
  for (int u = 0; u < KS_UNROLL; u++) {
    unsigned p = base + u * KS_THREADS;
    if (p < plane) v[u] = in[off + p];
  }
#pragma unroll
  for (int u = 0; u < KS_UNROLL; u++) {
    unsigned p = base + u * KS_THREADS;
    if (p < plane) {
      unsigned ix = p / (unsigned)g.nz;
      unsigned iz = p - ix * (unsigned)g.nz;
      op.apply(off + p, v[u], __ldg(g.kx + ix), ky, __ldg(g.kz + iz), (ix | (unsigned)(y0 + yl) | iz) == 0u, (int)ix, y0 + yl, (int)iz);
    }
  }
}

template <class Op>
static int run_kspace_t(baorec_ctx* ctx, const float2* in, Op op, cudaStream_t st) {
  if (ctx->p2p) {
    // peer-copy exchange: K[z][yl][x] -- the single-GPU layout restricted to this rank's rows
    KGeom g = kgeom_of(ctx);
    g.ny = ctx->ny_loc;
    g.y0 = ctx->y0;
    size_t kplane = (size_t)ctx->xh * ctx->ny_loc;
    dim3 grid(cdiv(kplane, KS_THREADS * KS_UNROLL), ctx->nz);
    BR_LAUNCH_NAMED(ctx, Op::name(), kspace_kernel<Op>, grid, KS_THREADS, 0, st, g, in, op);
    return BAOREC_OK;
  }
  size_t plane = (size_t)ctx->xh * ctx->nz;
  dim3 grid(cdiv(plane, KS_THREADS * KS_UNROLL), ctx->ny_loc);
  BR_LAUNCH_NAMED(ctx, Op::name(), kspace_kernel_t<Op>, grid, KS_THREADS, 0, st, kgeom_of(ctx), ctx->y0, in, op);
  return BAOREC_OK;
}

template <bool POTENTIAL>
struct DispCompOp {  // one component of Psi = i k delta_k / k^2 (src/iterative.jl:268) or i k phi_k (src/multigrid.jl:767), /M
  static const char* name() { return POTENTIAL ? "kspace_kernel_t<DispCompOp<potential>>" : "kspace_kernel_t<DispCompOp<density>>"; }
  float2* out;
  int comp;
  float invM;
  __device__ __forceinline__ void apply(size_t idx, float2 v, float kx, float ky, float kz, bool, int, int, int) const {
    float s;
    if (POTENTIAL) {
      s = invM;
    } else {
      float k2 = ksq(kx, ky, kz);
      s = k2 > 0.f ? __fdiv_rn(invM, k2) : 0.f;
    }
    float kc = comp == 0 ? kx : (comp == 1 ? ky : kz);
    out[idx] = make_float2(__fmul_rn(__fmul_rn(-v.y, s), kc), __fmul_rn(__fmul_rn(v.x, s), kc));
  }
};

// G = delta_k / (k^2 M) (or phi_k / M) and H = i k_z G: the x and y displacement fields are i k_x G and i k_y G, and those
// factors commute with the inverse z transform, so they are applied AFTER it (disp_xy_planes_kernel, dist.cu): two
// distributed inverse transforms per read-back instead of three.  Same Float32 products as DispOp / DispCompOp.
template <bool POTENTIAL>
struct DispGHOp {
  static const char* name() { return POTENTIAL ? "kspace_kernel<DispGHOp<potential>>" : "kspace_kernel<DispGHOp<density>>"; }
  float2* g;
  float2* h;
  float invM;
  __device__ __forceinline__ void apply(size_t idx, float2 v, float kx, float ky, float kz, bool, int, int, int) const {
    float s;
    if (POTENTIAL) {
      s = invM;
    } else {
      float k2 = ksq(kx, ky, kz);
      s = k2 > 0.f ? __fdiv_rn(invM, k2) : 0.f;
    }
    const float gre = __fmul_rn(v.x, s), gim = __fmul_rn(v.y, s);
    g[idx] = make_float2(gre, gim);
    h[idx] = make_float2(__fmul_rn(-gim, kz), __fmul_rn(gre, kz));
  }
};

int kpass_disp_gh(baorec_ctx* ctx, const float2* in, float2* g, float2* h, bool potential, cudaStream_t st) {
  const float invM = (float)(1.0 / (double)ctx->M);
  if (potential) {
    DispGHOp<true> op{g, h, invM};
    return run_kspace_t(ctx, in, op, st);
  }
  DispGHOp<false> op{g, h, invM};
  return run_kspace_t(ctx, in, op, st);
}

int stash_dc(baorec_ctx* ctx, const float2* ck, int slot, double mul, cudaStream_t st) {
  BR_LAUNCH(ctx, stash_dc_kernel, 1, 1, 0, st, ck, ctx->d_scal, slot, mul);
  return BAOREC_OK;
}

int kpass_fused_T(baorec_ctx* ctx, const float2* in, float2* out_c2r, float2* keep, const baorec_params* p,
                  cudaStream_t st) {
  const float invM = (float)(1.0 / (double)ctx->M);
  GaussTab gt;
  BR_TRY(gauss_tab(ctx, p->smoothing_radius, &gt, st));
  FusedLosOp<0> op{out_c2r, keep, gt, p->bias, ctx->d_scal,
                   {p->los[0], p->los[1], p->los[2]}, p->beta, p->n_iter, invM};
  return run_kspace_t(ctx, in, op, st);
}

// same recurrence on the R2C of delta_s (randoms set-up done in real space)
int kpass_fused_delta_T(baorec_ctx* ctx, const float2* in, float2* out_c2r, float2* keep, const baorec_params* p,
                        cudaStream_t st) {
  const float invM = (float)(1.0 / (double)ctx->M);
  FusedLosOp<1> op{out_c2r, keep, GaussTab{}, p->bias, ctx->d_scal, {p->los[0], p->los[1], p->los[2]}, p->beta, p->n_iter,
                   invM};
  return run_kspace_t(ctx, in, op, st);
}

int kpass_disp_T(baorec_ctx* ctx, const float2* in, float2* out, int comp, bool potential, cudaStream_t st) {
  const float invM = (float)(1.0 / (double)ctx->M);
  if (potential) {
    DispCompOp<true> op{out, comp, invM};
    return run_kspace_t(ctx, in, op, st);
  }
  DispCompOp<false> op{out, comp, invM};
  return run_kspace_t(ctx, in, op, st);
}

// dc[8] must hold 1 / (A0 bias) (stash_dc with mul = 1/bias)
int kpass_setup_box_T(baorec_ctx* ctx, const float2* in, float2* out, const baorec_params* p, cudaStream_t st) {
  GaussTab gt;
  BR_TRY(gauss_tab(ctx, p->smoothing_radius, &gt, st));
  SetupBoxOp op{out, gt, p->bias, ctx->d_scal};
  return run_kspace_t(ctx, in, op, st);
}

int kpass_gauss_T(baorec_ctx* ctx, const float2* in, float2* out, float R, cudaStream_t st) {
  GaussTab gt;
  BR_TRY(gauss_tab(ctx, R, &gt, st));
  GaussOp op{out, gt, 1.0 / (double)ctx->M};
  return run_kspace_t(ctx, in, op, st);
}

int kpass_iter_pair_T(baorec_ctx* ctx, const float2* in, float2* out, int i, int j, cudaStream_t st) {
  IterPairOp op{out, i, j, (float)(1.0 / (double)ctx->M)};
  return run_kspace_t(ctx, in, op, st);
}

int radial_update_slab(baorec_ctx* ctx, float* out, const float* src, const float* X, int nzl, int z_lo, int ci,
                       int cj, float fac, cudaStream_t st) {
  unsigned gx = stream_grid((size_t)ctx->nx * ctx->ny, 256);
  dim3 grid(gx > 64 ? 64 : gx, nzl);
  BR_LAUNCH(ctx, radial_update_kernel, grid, 256, 0, st, out, src, X, ctx->d_xv[0], ctx->d_xv[1], ctx->d_xv[2] + z_lo,
            ctx->nx, ctx->ny, ci, cj, fac);
  return BAOREC_OK;
}

int randoms_combine(baorec_ctx* ctx, float* out, const float* dat, const float* ran, float bias, float ran_min,
                    double n_ran, size_t cells, cudaStream_t st) {
  BR_LAUNCH(ctx, randoms_combine_kernel, stream_grid(cells, 256), 256, 0, st, out, dat, ran, ctx->d_scal, bias,
            (double)ran_min, n_ran, cells);
  return BAOREC_OK;
}

}  // namespace baorec
