// Power-spectrum multipoles P_0, P_2, P_4 of a density mesh on the device: the check the
// reference makes by eye on every reconstruction (test_helpers/simulation.py:56-75, pypowspec compute_auto_box on
// the catalogs before / after), as one R2C and ONE pass over the half mesh.
//
// pk_kernel: HBM-bound reduction, algorithmic bytes 8 Mc (16 Mc with a randoms mesh): every mode read once; the
// result is 5 * nbins doubles.
// Persistent grid (a multiple of the SM count), each block walks chunks of PK_THREADS * PK_UNROLL modes of one z
// plane with coalesced float2 loads issued back to back, bins them into a shared-memory histogram of Float64
// sums -- after a segmented shuffle reduction over the warp's runs of equal bins (|k| is monotonic along a row, so
// a warp holds a handful of runs): shared Float64 atomics are compare-and-swap loops on sm_100a
// (ATOMS.CAST.SPIN.64), one per run instead of one per lane keeps them conflict-free -- and flushes its histogram
// to global memory once at the end (REDG.ADD.F64).
#include <math.h>

#include <algorithm>
#include <vector>

#include "internal.cuh"
#include "pk_ops.cuh"

namespace baorec {

constexpr int PK_THREADS = 256;
constexpr int PK_UNROLL = 4;
constexpr int PK_MAX_BINS = 1024;  // 5 * 1024 doubles = 40 KB of shared memory

template <bool INTERLACED>
__global__ void __launch_bounds__(PK_THREADS)
pk_kernel(PkGeom g, const float* __restrict__ tkx, const float* __restrict__ tky, const float* __restrict__ tkz, int ny,
          unsigned chunks_per_plane, unsigned nchunks, const float2* __restrict__ in, const float2* __restrict__ in2,
          const double2* __restrict__ phx, const double2* __restrict__ phy, const double2* __restrict__ phz, double sa,
          double* __restrict__ acc) {
  extern __shared__ double s_acc[];  // [5][nbins]
  const int nb = g.nbins;
  for (int t = threadIdx.x; t < 5 * nb; t += blockDim.x) s_acc[t] = 0.0;
  __syncthreads();
  const unsigned plane = (unsigned)g.xh * (unsigned)ny;
  const int lane = threadIdx.x & 31;
  for (unsigned chunk = blockIdx.x; chunk < nchunks; chunk += gridDim.x) {  // block-uniform trip count
    const unsigned iz = chunk / chunks_per_plane;
    const unsigned base = (chunk - iz * chunks_per_plane) * (PK_THREADS * PK_UNROLL) + threadIdx.x;
    const size_t off = (size_t)iz * plane;
    const float kz = __ldg(tkz + iz);
    const double2 pz = INTERLACED ? __ldg(phz + iz) : make_double2(1.0, 0.0);
    float2 v[PK_UNROLL], v2[PK_UNROLL];
#pragma unroll
    for (int u = 0; u < PK_UNROLL; u++) {
      const unsigned p = base + u * PK_THREADS;
      v[u] = p < plane ? in[off + p] : make_float2(0.f, 0.f);
      v2[u] = (INTERLACED && p < plane) ? in2[off + p] : make_float2(0.f, 0.f);
    }
#pragma unroll
    for (int u = 0; u < PK_UNROLL; u++) {
      const unsigned p = base + u * PK_THREADS;
      int bin = -1;
      double c[5] = {0.0, 0.0, 0.0, 0.0, 0.0};
      if (p < plane) {
        const unsigned iy = p / (unsigned)g.xh;
        const unsigned ix = p - iy * (unsigned)g.xh;
        // field of the mode: delta_k / M; interlaced: [f1 + f2 exp(i (k_x h_x + k_y h_y + k_z h_z) / 2)] / 2, the phase
        // as the product of three tabulated per-axis factors (Float64)
        double re = (double)v[u].x, im = (double)v[u].y;
        if (INTERLACED) {
          const double2 a = __ldg(phx + ix), b = __ldg(phy + iy);
          const double abx = a.x * b.x - a.y * b.y, aby = a.x * b.y + a.y * b.x;
          const double cx = abx * pz.x - aby * pz.y, cy = abx * pz.y + aby * pz.x;
          const double r2 = (double)v2[u].x, i2 = (double)v2[u].y;
          re = 0.5 * (re + (r2 * cx - i2 * cy));
          im = 0.5 * (im + (r2 * cy + i2 * cx));
        }
        re *= sa;
        im *= sa;
        bin = pk_mode(g, re, im, __ldg(tkx + ix), __ldg(tky + iy), kz, (int)ix, (int)iy, (int)iz, c);
      }
      // Segmented reduction over the warp's runs of equal bins (|k| is monotonic along a row, so equal bins sit in
      // consecutive lanes): run id = number of run heads up to this lane; five shuffle steps leave the sum of each
      // run in its head lane, which issues the only atomics (two lanes of one instruction meet on an address only where a
      // row boundary inside the warp repeats a bin).
      const int prev = __shfl_up_sync(0xffffffffu, bin, 1);
      const bool head = lane == 0 || prev != bin;
      const unsigned heads = __ballot_sync(0xffffffffu, head);
      const int run = __popc(heads & (0xffffffffu >> (31 - lane)));
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const int r = __shfl_down_sync(0xffffffffu, run, d);
        const bool take = lane + d < 32 && r == run;
#pragma unroll
        for (int q = 0; q < 5; q++) {
          const double t = __shfl_down_sync(0xffffffffu, c[q], d);
          if (take) c[q] += t;
        }
      }
      if (head && bin >= 0) {
#pragma unroll
        for (int q = 0; q < 5; q++) atomicAdd(&s_acc[q * nb + bin], c[q]);
      }
    }
  }
  __syncthreads();
  for (int t = threadIdx.x; t < 5 * nb; t += blockDim.x) {
    const double s = s_acc[t];
    if (s != 0.0) atomicAdd(acc + t, s);
  }
}

// Float64 sums of the data (and randoms) mesh: sums[0] = sum rho, sums[1] = sum ran.  4 B/cell per mesh, HBM-bound.
__global__ void __launch_bounds__(256)
pk_sum_kernel(const float* __restrict__ rho, const float* __restrict__ ran, size_t n, double* __restrict__ sums) {
  double a = 0.0, b = 0.0;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    a += (double)rho[i];
    if (ran != nullptr) b += (double)ran[i];
  }
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) {
    a += __shfl_down_sync(0xffffffffu, a, d);
    b += __shfl_down_sync(0xffffffffu, b, d);
  }
  __shared__ double sa[8], sb[8];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) {
    sa[warp] = a;
    sb[warp] = b;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < 8; w++) {
      a += sa[w];
      b += sb[w];
    }
    atomicAdd(sums, a);
    if (ran != nullptr) atomicAdd(sums + 1, b);
  }
}

// The field whose spectrum is estimated, formed in REAL space so that the Float32 transform never sees the mean:
// out = rho * (M / sum rho) - 1, or rho * (M / sum rho) - ran * (M / sum ran) with a randoms mesh (compute_auto_box /
// compute_auto_box_rand of the reference's helpers).  Transforming the raw density instead leaves every mode with
// the transform's rounding of the DC term (~1e-7 * sum rho), which on a sparse low-k bin is 1e-4 of the signal: the
// failure of the first hardware run (32 x 32 x 33 mesh).  Float64 arithmetic per cell, rounded once.  12 B/cell.
__global__ void __launch_bounds__(256)
pk_contrast_kernel(const float* __restrict__ rho, const float* __restrict__ ran, size_t n, double ca, double cb,
                   float* __restrict__ out) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
    out[i] = (float)((double)rho[i] * ca - (ran != nullptr ? (double)ran[i] * cb : 1.0));
}

// 1 / sinc(x)^(2p): the inverse squared window of one axis
static double inv_window2(double x, int p) {
  const double s = x == 0.0 ? 1.0 : sin(x) / x;
  double r = 1.0;
  for (int i = 0; i < p; i++) r *= s;
  return 1.0 / (r * r);
}

// rho2 / ran2: NULL, or the meshes of the same catalogs painted half a cell further along every axis (interlacing)
int power_multipoles(baorec_ctx* ctx, const float* rho, const float* ran, const float* rho2, const float* ran2,
                     const float los[3], double kmin, double dk, int nbins, int mas_power, double shot, double* h_k,
                     double* h_nmodes, double* h_p0, double* h_p2, double* h_p4, cudaStream_t st) {
  const int n[3] = {ctx->nx, ctx->ny, ctx->nz};
  const int len[3] = {ctx->xh, ctx->ny, ctx->nz};
  const size_t ntab = (size_t)len[0] + len[1] + len[2];
  // window tables from the context's own k tables: 1 / W_a^2, W_a = sinc(k_a h_a / 2)^p, in Float64
  std::vector<float> kt(ntab);
  const size_t ntab2 = (ntab + 1) & ~(size_t)1;     // the phase table (double2) starts 16-byte aligned
  std::vector<double> wt(rho2 ? ntab2 + 2 * ntab : ntab);   // windows; interlaced: + (cos, sin)(k_a h_a / 2) per axis
  size_t o = 0;
  for (int a = 0; a < 3; a++) {
    BR_CUDA(cudaMemcpyAsync(kt.data() + o, ctx->d_k[a], sizeof(float) * len[a], cudaMemcpyDeviceToHost, st));
    o += len[a];
  }
  BR_CUDA(cudaStreamSynchronize(st));
  o = 0;
  for (int a = 0; a < 3; a++) {
    const double h = (double)ctx->L[a] / (double)n[a];
    for (int i = 0; i < len[a]; i++) {
      wt[o + i] = inv_window2((double)kt[o + i] * h / 2.0, mas_power);
      if (rho2) {
        wt[ntab2 + 2 * (o + i)] = cos((double)kt[o + i] * h / 2.0);
        wt[ntab2 + 2 * (o + i) + 1] = sin((double)kt[o + i] * h / 2.0);
      }
    }
    o += len[a];
  }
  double* d;
  BR_TRY(need_t(ctx, BUF_PK, wt.size() + (size_t)5 * nbins, &d));
  double* acc = d + wt.size();
  const double2* phase = reinterpret_cast<const double2*>(d + ntab2);
  BR_CUDA(cudaMemcpyAsync(d, wt.data(), sizeof(double) * wt.size(), cudaMemcpyHostToDevice, st));
  BR_CUDA(cudaMemsetAsync(acc, 0, sizeof(double) * 5 * nbins, st));
  // sums in Float64 -> contrast in real space -> ONE transform (the randoms are subtracted before it, not after)
  double* sums = ctx->d_scal + 16;  // four sums: rho, ran, rho shifted, ran shifted
  BR_CUDA(cudaMemsetAsync(sums, 0, 4 * sizeof(double), st));
  int sms = 148;
  BR_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, ctx->device));
  const unsigned rgrid = (unsigned)std::min<size_t>((size_t)sms * 8, cdiv(ctx->M, 256));
  BR_LAUNCH(ctx, pk_sum_kernel, rgrid, 256, 0, st, rho, ran, ctx->M, sums);
  if (rho2) BR_LAUNCH(ctx, pk_sum_kernel, rgrid, 256, 0, st, rho2, ran2, ctx->M, sums + 2);
  double hs[4] = {0.0, 1.0, 1.0, 1.0};
  BR_CUDA(cudaMemcpyAsync(hs, sums, 4 * sizeof(double), cudaMemcpyDeviceToHost, st));
  BR_CUDA(cudaStreamSynchronize(st));
  if (!ran) hs[1] = 1.0;
  if (!rho2) hs[2] = 1.0;
  if (!ran2) hs[3] = 1.0;
  if (!(hs[0] != 0.0) || !(hs[1] != 0.0) || !(hs[2] != 0.0) || !(hs[3] != 0.0)) {
    set_error("baorec_power_multipoles_f32: a mesh sums to zero (pass densities, not overdensities)");
    return BAOREC_ERR_INVALID;
  }
  float* contrast;
  float2 *ck0, *ck1 = nullptr;
  BR_TRY(need_t(ctx, BUF_RS, ctx->M, &contrast));
  BR_TRY(need_t(ctx, BUF_CK0, ctx->Mc, &ck0));
  BR_LAUNCH(ctx, pk_contrast_kernel, rgrid, 256, 0, st, rho, ran, ctx->M, (double)ctx->M / hs[0],
            ran ? (double)ctx->M / hs[1] : 0.0, contrast);
  BR_TRY(fft_r2c(ctx, contrast, ck0, st));
  if (rho2) {  // the interlaced mesh: same contrast, second transform (stream order frees `contrast` for it)
    BR_TRY(need_t(ctx, BUF_CK1, ctx->Mc, &ck1));
    BR_LAUNCH(ctx, pk_contrast_kernel, rgrid, 256, 0, st, rho2, ran2, ctx->M, (double)ctx->M / hs[2],
              ran2 ? (double)ctx->M / hs[3] : 0.0, contrast);
    BR_TRY(fft_r2c(ctx, contrast, ck1, st));
  }
  const double sa = 1.0 / (double)ctx->M;  // delta_k / M  ==  rho_k / rho_0 for k != 0
  PkGeom g;
  g.wx = d;
  g.wy = d + len[0];
  g.wz = d + len[0] + len[1];
  const double ln = sqrt((double)los[0] * los[0] + (double)los[1] * los[1] + (double)los[2] * los[2]);
  for (int a = 0; a < 3; a++) g.los[a] = (double)los[a] / ln;
  g.kmin = kmin;
  g.inv_dk = 1.0 / dk;
  g.nbins = nbins;
  g.xh = ctx->xh;
  g.nyq_x = (ctx->nx % 2 == 0) ? ctx->nx / 2 : -1;
  const size_t plane = (size_t)ctx->xh * ctx->ny;
  const unsigned cpp = cdiv(plane, PK_THREADS * PK_UNROLL);
  const size_t nchunks = (size_t)cpp * ctx->nz;
  BR_REQUIRE(nchunks < ((size_t)1 << 32), "mesh too large for the multipole kernel's chunk index");
  // persistent grid: SM count x resident blocks per SM (76 registers -> 3 blocks of 256 threads)
  int occ = 1;
  if (ck1) BR_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, pk_kernel<true>, PK_THREADS, sizeof(double) * 5 * nbins));
  else BR_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, pk_kernel<false>, PK_THREADS, sizeof(double) * 5 * nbins));
  unsigned grid = (unsigned)sms * (unsigned)(occ > 0 ? occ : 1);
  if (grid > nchunks) grid = (unsigned)nchunks;
  if (ck1)
    BR_LAUNCH_NAMED(ctx, "pk_kernel<interlaced>", pk_kernel<true>, grid, PK_THREADS, sizeof(double) * 5 * nbins, st, g, ctx->d_k[0],
                    ctx->d_k[1], ctx->d_k[2], ctx->ny, cpp, (unsigned)nchunks, ck0, (const float2*)ck1, phase, phase + len[0],
                    phase + len[0] + len[1], sa, acc);
  else
    BR_LAUNCH_NAMED(ctx, "pk_kernel", pk_kernel<false>, grid, PK_THREADS, sizeof(double) * 5 * nbins, st, g, ctx->d_k[0], ctx->d_k[1],
                    ctx->d_k[2], ctx->ny, cpp, (unsigned)nchunks, ck0, (const float2*)nullptr, phase, phase, phase, sa, acc);
  std::vector<double> h((size_t)5 * nbins);
  BR_CUDA(cudaMemcpyAsync(h.data(), acc, sizeof(double) * 5 * nbins, cudaMemcpyDeviceToHost, st));
  BR_CUDA(cudaStreamSynchronize(st));
  const double norm = (double)ctx->L[0] * (double)ctx->L[1] * (double)ctx->L[2];  // P = V |rho_k / rho_0|^2 = V |delta_k|^2 / M^2
  const double nan = ::nan("");
  for (int b = 0; b < nbins; b++) {
    const double cnt = h[b];
    h_nmodes[b] = cnt;
    if (cnt > 0.0) {
      h_k[b] = h[nbins + b] / cnt;
      h_p0[b] = h[2 * nbins + b] / cnt * norm - shot;
      h_p2[b] = 5.0 * h[3 * nbins + b] / cnt * norm;
      h_p4[b] = 9.0 * h[4 * nbins + b] / cnt * norm;
    } else {
      h_k[b] = h_p0[b] = h_p2[b] = h_p4[b] = nan;
    }
  }
  return BAOREC_OK;
}

}  // namespace baorec

using namespace baorec;

static int check_pk_args(const void* d_rho, const float* los, int nbins, double kmin, double dk, int mas_power, const void* a,
                         const void* b, const void* c, const void* d, const void* e) {
  BR_REQUIRE(d_rho != nullptr && los != nullptr, "NULL mesh / line of sight");
  BR_REQUIRE(a && b && c && d && e, "NULL output array");
  BR_REQUIRE(nbins >= 1 && nbins <= PK_MAX_BINS, "nbins must be in 1..1024");
  BR_REQUIRE(dk > 0.0 && kmin >= 0.0, "kmin >= 0 and dk > 0");
  BR_REQUIRE(mas_power >= 0 && mas_power <= 4, "mas_power (window exponent) must be in 0..4");
  BR_REQUIRE(los[0] != 0.f || los[1] != 0.f || los[2] != 0.f, "line of sight is the zero vector");
  return BAOREC_OK;
}

extern "C" int baorec_power_multipoles_f32(baorec_ctx* ctx, const float* d_rho, const float* d_ran, const float los[3], double kmin,
                                           double dk, int nbins, int mas_power, double shot, double* h_k,
                                           double* h_nmodes, double* h_p0, double* h_p2, double* h_p4,
                                           baorec_stream stream) {
  BR_NEED_PLAN(ctx);
  BR_TRY(check_pk_args(d_rho, los, nbins, kmin, dk, mas_power, h_k, h_nmodes, h_p0, h_p2, h_p4));
  return power_multipoles(ctx, d_rho, d_ran, nullptr, nullptr, los, kmin, dk, nbins, mas_power, shot, h_k, h_nmodes, h_p0,
                          h_p2, h_p4, (cudaStream_t)stream);
}

extern "C" int baorec_power_multipoles_interlaced_f32(baorec_ctx* ctx, const float* d_rho, const float* d_ran,
                                                      const float* d_rho_shifted, const float* d_ran_shifted,
                                                      const float los[3], double kmin, double dk, int nbins, int mas_power,
                                                      double shot, double* h_k, double* h_nmodes, double* h_p0,
                                                      double* h_p2, double* h_p4, baorec_stream stream) {
  BR_NEED_PLAN(ctx);
  BR_TRY(check_pk_args(d_rho, los, nbins, kmin, dk, mas_power, h_k, h_nmodes, h_p0, h_p2, h_p4));
  BR_REQUIRE(d_rho_shifted != nullptr, "NULL interlaced mesh");
  BR_REQUIRE((d_ran == nullptr) == (d_ran_shifted == nullptr), "randoms: both meshes or neither");
  return power_multipoles(ctx, d_rho, d_ran, d_rho_shifted, d_ran_shifted, los, kmin, dk, nbins, mas_power, shot, h_k,
                          h_nmodes, h_p0, h_p2, h_p4, (cudaStream_t)stream);
}

// q = p + (L/n)/2, and back by L where that leaves the box: the catalog of the interlaced mesh
__global__ void __launch_bounds__(256)
interlace_positions_kernel(const float* __restrict__ x, const float* __restrict__ y, const float* __restrict__ z, int64_t n,
                           float hx, float hy, float hz, float Lx, float Ly, float Lz, float mx, float my, float mz,
                           float* __restrict__ ox, float* __restrict__ oy, float* __restrict__ oz) {
  const int64_t step = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += step) {
    float a = __fadd_rn(x[i], hx), b = __fadd_rn(y[i], hy), c = __fadd_rn(z[i], hz);
    if (__fsub_rn(a, mx) >= Lx) a = __fsub_rn(a, Lx);
    if (__fsub_rn(b, my) >= Ly) b = __fsub_rn(b, Ly);
    if (__fsub_rn(c, mz) >= Lz) c = __fsub_rn(c, Lz);
    ox[i] = a;
    oy[i] = b;
    oz[i] = c;
  }
}

namespace baorec {
static int interlace_positions(baorec_ctx* ctx, const float* x, const float* y, const float* z, int64_t n, float* ox, float* oy,
                               float* oz, cudaStream_t st) {
  if (n == 0) return BAOREC_OK;
  const float hx = ctx->L[0] / (float)ctx->nx, hy = ctx->L[1] / (float)ctx->ny, hz = ctx->L[2] / (float)ctx->nz;  // Float32 divisions
  unsigned grid = (unsigned)std::min<size_t>((size_t)148 * 16, cdiv((size_t)n, 256));
  BR_LAUNCH(ctx, interlace_positions_kernel, grid, 256, 0, st, x, y, z, n, hx / 2.0f, hy / 2.0f, hz / 2.0f, ctx->L[0], ctx->L[1],
            ctx->L[2], ctx->mn[0], ctx->mn[1], ctx->mn[2], ox, oy, oz);
  return BAOREC_OK;
}

// paint one catalog into `mesh` (zeroed here) from a private copy of the positions (cic! writes wrapped positions back)
static int paint(baorec_ctx* ctx, float* mesh, const float* x, const float* y, const float* z, const float* w, int64_t n, int mas,
                 bool shifted, float* px, float* py, float* pz, cudaStream_t st) {
  BR_CUDA(cudaMemsetAsync(mesh, 0, sizeof(float) * ctx->M, st));
  if (shifted) {
    BR_TRY(interlace_positions(ctx, x, y, z, n, px, py, pz, st));
  } else {
    BR_CUDA(cudaMemcpyAsync(px, x, sizeof(float) * n, cudaMemcpyDeviceToDevice, st));
    BR_CUDA(cudaMemcpyAsync(py, y, sizeof(float) * n, cudaMemcpyDeviceToDevice, st));
    BR_CUDA(cudaMemcpyAsync(pz, z, sizeof(float) * n, cudaMemcpyDeviceToDevice, st));
  }
  return scatter(ctx, mesh, px, py, pz, w, n, 1, mas, st);
}
}  // namespace baorec

extern "C" int baorec_interlace_positions_f32(baorec_ctx* ctx, const float* d_x, const float* d_y, const float* d_z, int64_t n,
                                              float* d_ox, float* d_oy, float* d_oz, baorec_stream stream) {
  BR_NEED_PLAN(ctx);
  BR_REQUIRE(n >= 0 && (n == 0 || (d_x && d_y && d_z && d_ox && d_oy && d_oz)), "NULL position array");
  return interlace_positions(ctx, d_x, d_y, d_z, n, d_ox, d_oy, d_oz, (cudaStream_t)stream);
}

extern "C" int baorec_compute_auto_box_f32(baorec_ctx* ctx, const float* d_x, const float* d_y, const float* d_z, const float* d_w,
                                           int64_t n, const float* d_rx, const float* d_ry, const float* d_rz, const float* d_rw,
                                           int64_t nr, int mas, int interlace, const float los[3], double kmin, double dk,
                                           int nbins, double shot, double* h_k, double* h_nmodes, double* h_p0, double* h_p2,
                                           double* h_p4, baorec_stream stream) {
  BR_NEED_PLAN(ctx);
  BR_REQUIRE(mas == BAOREC_MAS_CIC || mas == BAOREC_MAS_TSC || mas == BAOREC_MAS_PCS, "unknown mas");
  const int mas_power = mas == BAOREC_MAS_CIC ? 2 : (mas == BAOREC_MAS_TSC ? 3 : 4);
  BR_TRY(check_pk_args(d_x, los, nbins, kmin, dk, mas_power, h_k, h_nmodes, h_p0, h_p2, h_p4));
  BR_REQUIRE(n > 0 && d_y && d_z && d_w, "empty catalog / NULL column");
  BR_REQUIRE(nr >= 0 && (nr == 0 || (d_rx && d_ry && d_rz && d_rw)), "NULL randoms column");
  cudaStream_t st = (cudaStream_t)stream;
  float *rho, *ran = nullptr, *rho2 = nullptr, *ran2 = nullptr, *pos;
  BR_TRY(need_t(ctx, BUF_RX, ctx->M, &rho));
  if (nr) BR_TRY(need_t(ctx, BUF_RAN, ctx->M, &ran));
  if (interlace) BR_TRY(need_t(ctx, BUF_RY, ctx->M, &rho2));
  if (interlace && nr) BR_TRY(need_t(ctx, BUF_RZ, ctx->M, &ran2));
  const size_t np = (size_t)std::max(n, nr);
  BR_TRY(need_t(ctx, BUF_PART, 3 * np, &pos));
  BR_TRY(paint(ctx, rho, d_x, d_y, d_z, d_w, n, mas, false, pos, pos + np, pos + 2 * np, st));
  if (interlace) BR_TRY(paint(ctx, rho2, d_x, d_y, d_z, d_w, n, mas, true, pos, pos + np, pos + 2 * np, st));
  if (nr) BR_TRY(paint(ctx, ran, d_rx, d_ry, d_rz, d_rw, nr, mas, false, pos, pos + np, pos + 2 * np, st));
  if (nr && interlace) BR_TRY(paint(ctx, ran2, d_rx, d_ry, d_rz, d_rw, nr, mas, true, pos, pos + np, pos + 2 * np, st));
  return power_multipoles(ctx, rho, ran, rho2, ran2, los, kmin, dk, nbins, mas_power, shot, h_k, h_nmodes, h_p0, h_p2, h_p4, st);
}
