// Context lifecycle, planning (cuFFT plans, k/x tables), scratch arena, FFT wrappers.
#include <math.h>
#include <stdarg.h>
#include <string.h>

#include "internal.cuh"

namespace baorec {

static thread_local char g_err[1024] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

static cudaEvent_t prof_event(baorec_ctx* ctx) {
  if (!ctx->prof_pool.empty()) {
    cudaEvent_t e = ctx->prof_pool.back();
    ctx->prof_pool.pop_back();
    return e;
  }
  cudaEvent_t e = nullptr;
  cudaEventCreate(&e);
  return e;
}

int prof_begin(baorec_ctx* ctx, const char* name, cudaStream_t st) {
  if (!ctx->prof_on) return -1;
  ProfRec r{name, prof_event(ctx), prof_event(ctx)};
  cudaEventRecord(r.a, st);
  ctx->prof.push_back(r);
  return (int)ctx->prof.size() - 1;
}

void prof_end(baorec_ctx* ctx, int idx, cudaStream_t st) {
  if (idx < 0) return;
  cudaEventRecord(ctx->prof[idx].b, st);
}

int need(baorec_ctx* ctx, BufId id, size_t bytes, void** out) {
  Buf& b = ctx->bufs[id];
  if (b.bytes < bytes) {
    if (id == BUF_RX || id == BUF_RY || id == BUF_RZ) ctx->disp_valid = false;
    if (b.p) {
      BR_CUDA(cudaFree(b.p));
      b.p = nullptr;
      b.bytes = 0;
    }
    cudaError_t e = cudaMalloc(&b.p, bytes);
    if (e != cudaSuccess) {
      set_error("cudaMalloc of %zu bytes for scratch buffer %d failed: %s", bytes, (int)id, cudaGetErrorString(e));
      b.p = nullptr;
      return BAOREC_ERR_NOMEM;
    }
    b.bytes = bytes;
  }
  *out = b.p;
  return BAOREC_OK;
}

void release(baorec_ctx* ctx, BufId id) {
  Buf& b = ctx->bufs[id];
  if (b.p) cudaFree(b.p);
  b.p = nullptr;
  b.bytes = 0;
}

int reset_oob(baorec_ctx* ctx, cudaStream_t st) {
  BR_CUDA(cudaMemsetAsync(ctx->d_oob, 0, 2 * sizeof(unsigned long long), st));
  return BAOREC_OK;
}

int check_oob(baorec_ctx* ctx, cudaStream_t st, const char* what) {
  unsigned long long hh[2] = {0, 0};
  BR_CUDA(cudaMemcpyAsync(hh, ctx->d_oob, sizeof(hh), cudaMemcpyDeviceToHost, st));
  BR_CUDA(cudaStreamSynchronize(st));
  ctx->last_wrapped = (int64_t)hh[1];
  unsigned long long h = hh[0];
  if (h != 0) {
    set_error("%s: %llu particle(s) outside the mesh (the reference would raise BoundsError / write out of bounds)",
              what, h);
    return BAOREC_ERR_OUT_OF_BOX;
  }
  return BAOREC_OK;
}

static void destroy_split_plans(baorec_ctx* ctx) {
  if (ctx->split_planes_planned) {
    cufftDestroy(ctx->ps_r2c);
    cufftDestroy(ctx->ps_c2r);
    cufftDestroy(ctx->ps_z);
    ctx->split_planes_planned = 0;
  }
}

// Plans of the split path: 2-D R2C/C2R over C planes, strided 1-D C2C along z (in place).
static int split_setup(baorec_ctx* ctx) {
  int C = ctx->opt_fft_split;
  if (C > ctx->nz) C = ctx->nz;
  while (ctx->nz % C) C--;
  if (ctx->split_planes_planned == C) return BAOREC_OK;
  destroy_split_plans(ctx);
  size_t w[3] = {0, 0, 0};
  int n2[2] = {ctx->ny, ctx->nx};
  int n1[1] = {ctx->nz};
  const int stride = ctx->ny * ctx->xh;
  BR_CUFFT(cufftCreate(&ctx->ps_r2c));
  BR_CUFFT(cufftCreate(&ctx->ps_c2r));
  BR_CUFFT(cufftCreate(&ctx->ps_z));
  ctx->split_planes_planned = C;
  BR_CUFFT(cufftSetAutoAllocation(ctx->ps_r2c, 0));
  BR_CUFFT(cufftSetAutoAllocation(ctx->ps_c2r, 0));
  BR_CUFFT(cufftSetAutoAllocation(ctx->ps_z, 0));
  BR_CUFFT(cufftMakePlanMany(ctx->ps_r2c, 2, n2, nullptr, 1, 0, nullptr, 1, 0, CUFFT_R2C, C, &w[0]));
  BR_CUFFT(cufftMakePlanMany(ctx->ps_c2r, 2, n2, nullptr, 1, 0, nullptr, 1, 0, CUFFT_C2R, C, &w[1]));
  BR_CUFFT(cufftMakePlanMany(ctx->ps_z, 1, n1, n1, stride, 1, n1, stride, 1, CUFFT_C2C, stride, &w[2]));
  size_t wmax = w[0] > w[1] ? w[0] : w[1];
  if (w[2] > wmax) wmax = w[2];
  if (wmax > ctx->bufs[BUF_WORK].bytes) {
    // the shared work area is about to move: rebind every plan that uses it
    BR_CUDA(cudaDeviceSynchronize());
    void* work = nullptr;
    BR_TRY(need(ctx, BUF_WORK, wmax, &work));
    if (ctx->have_plans) {
      BR_CUFFT(cufftSetWorkArea(ctx->r2c, work));
      BR_CUFFT(cufftSetWorkArea(ctx->c2r, work));
    }
    if (ctx->have_x_plans) {
      BR_CUFFT(cufftSetWorkArea(ctx->px_r2c, work));
      BR_CUFFT(cufftSetWorkArea(ctx->px_c2r, work));
    }
    if (ctx->have_dist_plans) {
      BR_CUFFT(cufftSetWorkArea(ctx->p2d_r2c, work));
      BR_CUFFT(cufftSetWorkArea(ctx->p2d_c2r, work));
      BR_CUFFT(cufftSetWorkArea(ctx->p1d, work));
    }
  }
  void* work = ctx->bufs[BUF_WORK].p;
  BR_CUFFT(cufftSetWorkArea(ctx->ps_r2c, work));
  BR_CUFFT(cufftSetWorkArea(ctx->ps_c2r, work));
  BR_CUFFT(cufftSetWorkArea(ctx->ps_z, work));
  return BAOREC_OK;
}

static int split_r2c(baorec_ctx* ctx, const float* in, float2* out, cudaStream_t st) {
  BR_TRY(split_setup(ctx));
  const int C = ctx->split_planes_planned;
  const size_t rplane = (size_t)ctx->nx * ctx->ny, cplane = (size_t)ctx->xh * ctx->ny;
  BR_CUFFT(cufftSetStream(ctx->ps_r2c, st));
  BR_CUFFT(cufftSetStream(ctx->ps_z, st));
  int pi = prof_begin(ctx, "cufft_split_2d_r2c", st);
  for (int z = 0; z < ctx->nz; z += C)
    BR_CUFFT(cufftExecR2C(ctx->ps_r2c, (cufftReal*)(in + (size_t)z * rplane), (cufftComplex*)(out + (size_t)z * cplane)));
  prof_end(ctx, pi, st);
  pi = prof_begin(ctx, "cufft_split_z", st);
  BR_CUFFT(cufftExecC2C(ctx->ps_z, (cufftComplex*)out, (cufftComplex*)out, CUFFT_FORWARD));
  prof_end(ctx, pi, st);
  ctx->n_fft++;
  return BAOREC_OK;
}

static int split_c2r(baorec_ctx* ctx, float2* in, float* out, cudaStream_t st) {
  BR_TRY(split_setup(ctx));
  const int C = ctx->split_planes_planned;
  const size_t rplane = (size_t)ctx->nx * ctx->ny, cplane = (size_t)ctx->xh * ctx->ny;
  BR_CUFFT(cufftSetStream(ctx->ps_c2r, st));
  BR_CUFFT(cufftSetStream(ctx->ps_z, st));
  int pi = prof_begin(ctx, "cufft_split_z", st);
  BR_CUFFT(cufftExecC2C(ctx->ps_z, (cufftComplex*)in, (cufftComplex*)in, CUFFT_INVERSE));
  prof_end(ctx, pi, st);
  pi = prof_begin(ctx, "cufft_split_2d_c2r", st);
  for (int z = 0; z < ctx->nz; z += C)
    BR_CUFFT(cufftExecC2R(ctx->ps_c2r, (cufftComplex*)(in + (size_t)z * cplane), (cufftReal*)(out + (size_t)z * rplane)));
  prof_end(ctx, pi, st);
  ctx->n_fft++;
  return BAOREC_OK;
}

int fft_r2c(baorec_ctx* ctx, const float* in, float2* out, cudaStream_t st) {
  if (own_fft_available(ctx)) return own_r2c(ctx, in, out, st);
  if (ctx->opt_fft_split > 0) return split_r2c(ctx, in, out, st);
  BR_CUFFT(cufftSetStream(ctx->r2c, st));
  int pi = prof_begin(ctx, "cufft_r2c", st);
  BR_CUFFT(cufftExecR2C(ctx->r2c, (cufftReal*)in, (cufftComplex*)out));
  prof_end(ctx, pi, st);
  ctx->n_fft++;
  return BAOREC_OK;
}

int fft_c2r(baorec_ctx* ctx, float2* in, float* out, cudaStream_t st) {
  if (own_fft_available(ctx)) return own_c2r(ctx, in, out, st);
  if (ctx->opt_fft_split > 0) return split_c2r(ctx, in, out, st);
  BR_CUFFT(cufftSetStream(ctx->c2r, st));
  int pi = prof_begin(ctx, "cufft_c2r", st);
  BR_CUFFT(cufftExecC2R(ctx->c2r, (cufftComplex*)in, (cufftReal*)out));
  prof_end(ctx, pi, st);
  ctx->n_fft++;
  return BAOREC_OK;
}

// k_vec (src/utils.jl:3-10): fs = T(2pi n / L) from a Float64 product, multiplier fs/n in T,
// value = T(integer) * multiplier.
static void host_kvec(int n, float L, bool half, std::vector<float>& out) {
  float fs = (float)(2.0 * M_PI * (double)n / (double)L);
  float mult = fs / (float)n;
  if (half) {
    out.resize(n / 2 + 1);
    for (int i = 0; i <= n / 2; i++) out[i] = (float)i * mult;
  } else {
    out.resize(n);
    int nn = (n + 1) >> 1;
    for (int i = 0; i < n; i++) {
      int f = i < nn ? i : i - n;
      out[i] = (float)f * mult;
    }
  }
}

// x_vec (src/utils.jl:21-25): cell = L/n in T; Float64 range min + 0.5 cell + i cell, rounded to T.
void host_xvec(int n, float L, float mn, std::vector<float>& out) {
  float cell = L / (float)n;
  double start = (double)mn + 0.5 * (double)cell;
  out.resize(n);
  for (int i = 0; i < n; i++) out[i] = (float)(start + (double)i * (double)cell);
}

int gauss_tables(baorec_ctx* ctx, float R, const double** gx, const double** gy, const double** gz, cudaStream_t st) {
  const int len[3] = {ctx->xh, ctx->ny, ctx->nz};
  const int n[3] = {ctx->nx, ctx->ny, ctx->nz};
  if (!ctx->d_gauss) BR_CUDA(cudaMalloc(&ctx->d_gauss, sizeof(double) * (size_t)(len[0] + len[1] + len[2])));
  if (!ctx->gauss_valid || ctx->gauss_R != R) {
    const float R2 = R * R;  // T(R)^2 in Float32, promoted (src/utils.jl:52)
    std::vector<double> h;
    h.reserve((size_t)len[0] + len[1] + len[2]);
    for (int a = 0; a < 3; a++) {
      std::vector<float> k;
      host_kvec(n[a], ctx->L[a], a == 0, k);
      for (int i = 0; i < len[a]; i++) {
        const float k2 = k[i] * k[i];
        h.push_back(exp(-0.5 * (double)R2 * (double)k2));
      }
    }
    // earlier passes on `st` may still read the old tables: the copy is ordered after them
    BR_CUDA(cudaMemcpyAsync(ctx->d_gauss, h.data(), h.size() * sizeof(double), cudaMemcpyHostToDevice, st));
    BR_CUDA(cudaStreamSynchronize(st));  // h is pageable and local
    ctx->gauss_R = R;
    ctx->gauss_valid = true;
  }
  *gx = ctx->d_gauss;
  *gy = ctx->d_gauss + len[0];
  *gz = ctx->d_gauss + len[0] + len[1];
  return BAOREC_OK;
}

static int upload_tables(baorec_ctx* ctx) {
  ctx->gauss_valid = false;
  ctx->sortc_valid = false;  // tile keys depend on the box
  int n[3] = {ctx->nx, ctx->ny, ctx->nz};
  for (int a = 0; a < 3; a++) {
    std::vector<float> k, x;
    host_kvec(n[a], ctx->L[a], a == 0, k);
    host_xvec(n[a], ctx->L[a], ctx->mn[a], x);
    BR_CUDA(cudaMemcpy(ctx->d_k[a], k.data(), k.size() * sizeof(float), cudaMemcpyHostToDevice));
    BR_CUDA(cudaMemcpy(ctx->d_xv[a], x.data(), x.size() * sizeof(float), cudaMemcpyHostToDevice));
    ctx->cell[a] = ctx->L[a] / (float)n[a];
  }
  ctx->mg_radial_tables = false;
  return BAOREC_OK;
}

static void destroy_plans(baorec_ctx* ctx) {
  destroy_split_plans(ctx);
  if (ctx->have_x_plans) {
    cufftDestroy(ctx->px_r2c);
    cufftDestroy(ctx->px_c2r);
    ctx->have_x_plans = false;
  }
  if (ctx->have_plans) {
    cufftDestroy(ctx->r2c);
    cufftDestroy(ctx->c2r);
    ctx->have_plans = false;
  }
  if (ctx->have_dist_plans) {
    cufftDestroy(ctx->p2d_r2c);
    cufftDestroy(ctx->p2d_c2r);
    cufftDestroy(ctx->p1d);
    ctx->have_dist_plans = false;
  }
  if (ctx->chunk_planes) {
    cufftDestroy(ctx->pc_r2c);
    cufftDestroy(ctx->pc_c2r);
    ctx->chunk_planes = 0;
  }
}

int plan_common(baorec_ctx* ctx, int nx, int ny, int nz, const float L[3], const float mn[3]) {
  BR_REQUIRE(ctx != nullptr, "ctx is NULL");
  BR_REQUIRE(nx >= 2 && ny >= 2 && nz >= 2, "mesh must be at least 2 cells per axis");
  BR_REQUIRE(nx % 2 == 0, "nx must be even (R2C halves the x axis)");
  BR_REQUIRE(nz <= 65535 && ny <= 65535, "ny, nz must be <= 65535");
  BR_REQUIRE(L && mn, "box_size / box_min are NULL");
  BR_REQUIRE(L[0] > 0 && L[1] > 0 && L[2] > 0, "box_size must be positive");
  BR_CUDA(cudaSetDevice(ctx->device));
  bool same = ctx->planned && ctx->nx == nx && ctx->ny == ny && ctx->nz == nz;
  if (!same) {
    destroy_plans(ctx);
    for (int a = 0; a < 3; a++) {
      if (ctx->d_k[a]) cudaFree(ctx->d_k[a]);
      if (ctx->d_xv[a]) cudaFree(ctx->d_xv[a]);
      ctx->d_k[a] = ctx->d_xv[a] = nullptr;
    }
    ctx->levels.clear();
    ctx->dlevels.clear();
    if (ctx->d_gauss) cudaFree(ctx->d_gauss);
    ctx->d_gauss = nullptr;
    ctx->gauss_valid = false;
    ctx->nx = nx;
    ctx->ny = ny;
    ctx->nz = nz;
    ctx->xh = nx / 2 + 1;
    ctx->M = (size_t)nx * ny * nz;
    ctx->Mc = (size_t)ctx->xh * ny * nz;
    int n[3] = {nx, ny, nz};
    for (int a = 0; a < 3; a++) {
      BR_CUDA(cudaMalloc(&ctx->d_k[a], sizeof(float) * (a == 0 ? ctx->xh : n[a])));
      BR_CUDA(cudaMalloc(&ctx->d_xv[a], sizeof(float) * n[a]));
    }
    ctx->cache_valid = false;
  }
  ctx->kcache_valid = false;
  ctx->disp_valid = false;
  ctx->mg_result_mesh = nullptr;
  for (int a = 0; a < 3; a++) {
    ctx->L[a] = L[a];
    ctx->mn[a] = mn[a];
  }
  BR_TRY(upload_tables(ctx));
  return same ? 1 : 0;
}

}  // namespace baorec

using namespace baorec;

extern "C" {

int baorec_version(void) { return BAOREC_VERSION; }
const char* baorec_last_error(void) { return g_err; }

int baorec_create(int device, baorec_ctx** out) {
  BR_REQUIRE(out != nullptr, "out is NULL");
  int ndev = 0;
  BR_CUDA(cudaGetDeviceCount(&ndev));
  BR_REQUIRE(device >= 0 && device < ndev, "device index out of range");
  BR_CUDA(cudaSetDevice(device));
  baorec_ctx* ctx = new baorec_ctx();
  ctx->device = device;
  BR_CUDA(cudaMalloc(&ctx->d_oob, 2 * sizeof(unsigned long long)));  // [0] out-of-box, [1] positions wrapped
  BR_CUDA(cudaMemset(ctx->d_oob, 0, 2 * sizeof(unsigned long long)));
  BR_CUDA(cudaMalloc(&ctx->d_scal, 32 * sizeof(double)));
  BR_CUDA(cudaMalloc(&ctx->d_hash, 4 * sizeof(unsigned long long)));
  BR_CUDA(cudaMalloc(&ctx->d_minmax, 8 * sizeof(float)));
  BR_CUDA(cudaStreamCreateWithFlags(&ctx->own_stream, cudaStreamNonBlocking));
  BR_CUDA(cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
  BR_CUDA(cudaStreamCreateWithFlags(&ctx->side_stream, cudaStreamNonBlocking));
  {
    // the communication streams carry copies and one-warp flag kernels: highest priority, so that a flag kernel is
    // dispatched at once instead of queueing behind the blocks of whatever large grid the main stream is running
    // (measured: 190 us per flag kernel at default priority next to cuFFT)
    int lo = 0, hi = 0;
    BR_CUDA(cudaDeviceGetStreamPriorityRange(&lo, &hi));
    BR_CUDA(cudaStreamCreateWithPriority(&ctx->comm_stream, cudaStreamNonBlocking, hi));
    BR_CUDA(cudaStreamCreateWithPriority(&ctx->comm_stream2, cudaStreamNonBlocking, hi));
    BR_CUDA(cudaStreamCreateWithPriority(&ctx->comm_stream3, cudaStreamNonBlocking, hi));
    BR_CUDA(cudaStreamCreateWithPriority(&ctx->comm_stream4, cudaStreamNonBlocking, hi));
  }
  for (int i = 0; i < 8; i++) {
    BR_CUDA(cudaEventCreateWithFlags(&ctx->ev_local[i], cudaEventDisableTiming));
    BR_CUDA(cudaEventCreateWithFlags(&ctx->ev_split[i], cudaEventDisableTiming));
    BR_CUDA(cudaEventCreateWithFlags(&ctx->ev_split2[i], cudaEventDisableTiming));
    BR_CUDA(cudaEventCreateWithFlags(&ctx->ev_chunk[i], cudaEventDisableTiming));
    BR_CUDA(cudaEventCreateWithFlags(&ctx->ev_a2a[i], cudaEventDisableTiming));
  }
  BR_CUDA(cudaEventCreateWithFlags(&ctx->ev_fork, cudaEventDisableTiming));
  BR_CUDA(cudaEventCreateWithFlags(&ctx->ev_join, cudaEventDisableTiming));
  BR_CUDA(cudaEventCreateWithFlags(&ctx->ev_copy, cudaEventDisableTiming));
  for (int i = 0; i < 8; i++) BR_CUDA(cudaEventCreate(&ctx->ev[i]));
  *out = ctx;
  return BAOREC_OK;
}

int baorec_comm_destroy_internal(baorec_ctx* ctx);

int baorec_destroy(baorec_ctx* ctx) {
  if (!ctx) return BAOREC_OK;
  cudaSetDevice(ctx->device);
  cudaDeviceSynchronize();
  baorec_comm_destroy_internal(ctx);
  destroy_plans(ctx);
  for (int i = 0; i < BUF_COUNT; i++) release(ctx, (BufId)i);
  for (int a = 0; a < 3; a++) {
    if (ctx->d_k[a]) cudaFree(ctx->d_k[a]);
    if (ctx->d_xv[a]) cudaFree(ctx->d_xv[a]);
  }
  for (int a = 0; a < 2; a++)
    if (ctx->d_tw[a]) cudaFree(ctx->d_tw[a]);
  if (ctx->d_gauss) cudaFree(ctx->d_gauss);
  for (auto& hs : ctx->batch_host_slots)
    if (hs.base) cudaFreeHost(hs.base);
  ctx->batch_host_slots.clear();
  if (ctx->d_oob) cudaFree(ctx->d_oob);
  if (ctx->d_scal) cudaFree(ctx->d_scal);
  if (ctx->d_hash) cudaFree(ctx->d_hash);
  if (ctx->d_minmax) cudaFree(ctx->d_minmax);
  if (ctx->d_cosmo_r) cudaFree(ctx->d_cosmo_r);
  if (ctx->d_cosmo_g) cudaFree(ctx->d_cosmo_g);
  for (int i = 0; i < 8; i++)
    if (ctx->ev[i]) cudaEventDestroy(ctx->ev[i]);
  for (auto& r : ctx->prof) {
    cudaEventDestroy(r.a);
    cudaEventDestroy(r.b);
  }
  for (auto e : ctx->prof_pool) cudaEventDestroy(e);
  if (ctx->own_stream) cudaStreamDestroy(ctx->own_stream);
  if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
  if (ctx->side_stream) cudaStreamDestroy(ctx->side_stream);
  if (ctx->comm_stream) cudaStreamDestroy(ctx->comm_stream);
  if (ctx->comm_stream2) cudaStreamDestroy(ctx->comm_stream2);
  if (ctx->comm_stream3) cudaStreamDestroy(ctx->comm_stream3);
  if (ctx->comm_stream4) cudaStreamDestroy(ctx->comm_stream4);
  for (int i = 0; i < 8; i++) {
    if (ctx->ev_local[i]) cudaEventDestroy(ctx->ev_local[i]);
    if (ctx->ev_split[i]) cudaEventDestroy(ctx->ev_split[i]);
    if (ctx->ev_split2[i]) cudaEventDestroy(ctx->ev_split2[i]);
    if (ctx->ev_chunk[i]) cudaEventDestroy(ctx->ev_chunk[i]);
    if (ctx->ev_a2a[i]) cudaEventDestroy(ctx->ev_a2a[i]);
  }
  if (ctx->ev_fork) cudaEventDestroy(ctx->ev_fork);
  if (ctx->ev_join) cudaEventDestroy(ctx->ev_join);
  if (ctx->ev_copy) cudaEventDestroy(ctx->ev_copy);
  delete ctx;
  return BAOREC_OK;
}

int baorec_plan(baorec_ctx* ctx, int nx, int ny, int nz, const float box_size[3], const float box_min[3]) {
  int s = plan_common(ctx, nx, ny, nz, box_size, box_min);
  if (s < 0) return s;
  ctx->dist = false;
  if (!ctx->have_plans) {
    size_t w1 = 0, w2 = 0;
    BR_CUFFT(cufftCreate(&ctx->r2c));
    BR_CUFFT(cufftCreate(&ctx->c2r));
    ctx->have_plans = true;
    BR_CUFFT(cufftSetAutoAllocation(ctx->r2c, 0));
    BR_CUFFT(cufftSetAutoAllocation(ctx->c2r, 0));
    BR_CUFFT(cufftMakePlan3d(ctx->r2c, nz, ny, nx, CUFFT_R2C, &w1));
    BR_CUFFT(cufftMakePlan3d(ctx->c2r, nz, ny, nx, CUFFT_C2R, &w2));
    ctx->work_bytes = w1 > w2 ? w1 : w2;
    // batched contiguous 1-D transforms along x for the own-FFT path
    size_t w3 = 0, w4 = 0;
    int n1[1] = {nx};
    BR_CUFFT(cufftCreate(&ctx->px_r2c));
    BR_CUFFT(cufftCreate(&ctx->px_c2r));
    ctx->have_x_plans = true;
    BR_CUFFT(cufftSetAutoAllocation(ctx->px_r2c, 0));
    BR_CUFFT(cufftSetAutoAllocation(ctx->px_c2r, 0));
    BR_CUFFT(cufftMakePlanMany(ctx->px_r2c, 1, n1, nullptr, 1, 0, nullptr, 1, 0, CUFFT_R2C, ny * nz, &w3));
    BR_CUFFT(cufftMakePlanMany(ctx->px_c2r, 1, n1, nullptr, 1, 0, nullptr, 1, 0, CUFFT_C2R, ny * nz, &w4));
    if (w3 > ctx->work_bytes) ctx->work_bytes = w3;
    if (w4 > ctx->work_bytes) ctx->work_bytes = w4;
    void* work = nullptr;
    BR_TRY(need(ctx, BUF_WORK, ctx->work_bytes > 0 ? ctx->work_bytes : 16, &work));
    BR_CUFFT(cufftSetWorkArea(ctx->r2c, work));
    BR_CUFFT(cufftSetWorkArea(ctx->c2r, work));
    BR_CUFFT(cufftSetWorkArea(ctx->px_r2c, work));
    BR_CUFFT(cufftSetWorkArea(ctx->px_c2r, work));
    BR_TRY(own_fft_setup(ctx));
  }
  ctx->planned = true;
  return BAOREC_OK;
}

int baorec_set_box(baorec_ctx* ctx, const float box_size[3], const float box_min[3]) {
  BR_REQUIRE(ctx != nullptr, "ctx is NULL");
  if (!ctx->planned) {
    set_error("baorec_set_box: call baorec_plan first");
    return BAOREC_ERR_NOT_PLANNED;
  }
  BR_REQUIRE(box_size && box_min, "box_size / box_min are NULL");
  BR_REQUIRE(box_size[0] > 0 && box_size[1] > 0 && box_size[2] > 0, "box_size must be positive");
  BR_CUDA(cudaSetDevice(ctx->device));
  BR_CUDA(cudaDeviceSynchronize());
  ctx->disp_valid = false;
  for (int a = 0; a < 3; a++) {
    ctx->L[a] = box_size[a];
    ctx->mn[a] = box_min[a];
  }
  return upload_tables(ctx);
}

int64_t baorec_scratch_bytes(const baorec_ctx* ctx) {
  if (!ctx) return 0;
  int64_t t = 0;
  for (int i = 0; i < BUF_COUNT; i++) t += (int64_t)ctx->bufs[i].bytes;
  return t;
}

int64_t baorec_sort_reuse_count(const baorec_ctx* ctx) { return ctx ? ctx->n_sort_reuse : 0; }

int baorec_launch_counts(const baorec_ctx* ctx, int64_t* kernels, int64_t* fft_execs) {
  BR_REQUIRE(ctx != nullptr, "ctx is NULL");
  if (kernels) *kernels = ctx->n_kernels;
  if (fft_execs) *fft_execs = ctx->n_fft;
  return BAOREC_OK;
}

int baorec_last_stage_ms(const baorec_ctx* ctx, float* out, int cap) {
  if (!ctx || !out) return 0;
  int n = ctx->n_stage < cap ? ctx->n_stage : cap;
  for (int i = 0; i < n; i++) out[i] = ctx->stage_ms[i];
  return n;
}

float* baorec_result_cache(baorec_ctx* ctx) {
  if (!ctx || !ctx->cache_valid) return nullptr;
  return (float*)ctx->bufs[BUF_CACHE].p;
}

int baorec_profile_enable(baorec_ctx* ctx, int on) {
  BR_REQUIRE(ctx != nullptr, "ctx is NULL");
  BR_CUDA(cudaSetDevice(ctx->device));
  BR_CUDA(cudaDeviceSynchronize());
  for (auto& r : ctx->prof) {
    ctx->prof_pool.push_back(r.a);
    ctx->prof_pool.push_back(r.b);
  }
  ctx->prof.clear();
  ctx->prof_on = on != 0;
  return BAOREC_OK;
}

int baorec_profile_read(baorec_ctx* ctx, char* names, int name_stride, float* total_ms, int32_t* counts, int cap) {
  BR_REQUIRE(ctx != nullptr && names && total_ms && counts && name_stride > 1 && cap > 0, "profile_read arguments");
  BR_CUDA(cudaSetDevice(ctx->device));
  BR_CUDA(cudaDeviceSynchronize());
  int n = 0;
  for (auto& r : ctx->prof) {
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, r.a, r.b) != cudaSuccess) continue;
    int k = -1;
    for (int i = 0; i < n; i++)
      if (strncmp(names + (size_t)i * name_stride, r.name, name_stride - 1) == 0) {
        k = i;
        break;
      }
    if (k < 0) {
      if (n >= cap) continue;
      k = n++;
      strncpy(names + (size_t)k * name_stride, r.name, name_stride - 1);
      names[(size_t)k * name_stride + name_stride - 1] = 0;
      total_ms[k] = 0.f;
      counts[k] = 0;
    }
    total_ms[k] += ms;
    counts[k] += 1;
  }
  return n;
}

int baorec_host_alloc(void** out, int64_t bytes) {
  BR_REQUIRE(out != nullptr && bytes > 0, "baorec_host_alloc arguments");
  BR_CUDA(cudaHostAlloc(out, (size_t)bytes, cudaHostAllocDefault));
  return BAOREC_OK;
}

int baorec_host_free(void* p) {
  if (p) BR_CUDA(cudaFreeHost(p));
  return BAOREC_OK;
}

}  // extern "C"
