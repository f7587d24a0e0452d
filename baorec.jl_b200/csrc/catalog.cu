// Catalog pre/post-processing on the device: the per-particle streaming
// work that brackets every lightcone reconstruction in the reference's examples and runs there on
// CPU threads -- sky <-> Cartesian with the comoving-distance / redshift interpolators
// (src/cosmo.jl:70-103, examples/lightcone.jl:30-80), FKP weights (examples/lightcone.jl:82) and the
// periodic re-wrap of reconstructed positions (test_helpers/simulation.py:38,51-52).
//
// Precision follows Julia's promotion rules for the Float32 catalogs of the examples (DESIGN.md
// section 10): Float64 tables, `ra * pi / 180` in Float32, cos/sin through a Float64
// kernel rounded once to Float32, the product dist * cos(dec) * cos(ra) * h in Float64 rounded once.
// All kernels are grid-stride over the particles with coalesced SoA accesses (12 B in, 12 B out per
// particle); the 0.8 MB tables stay L2-resident, and the inverse interpolation narrows its search
// with a 2048-entry coarse table in shared memory before touching them.
#include <math.h>

#include <vector>

#include "catalog_math.cuh"
#include "internal.cuh"

namespace baorec {

namespace {

using namespace baorec::catalog;

constexpr int kBlock = 256;

inline unsigned catalog_grid(int64_t n) {
  // persistent-style grid: at most 8 resident blocks of 256 threads on each of the 148 SMs
  int64_t want = (n + kBlock - 1) / kBlock;
  int64_t cap = 148 * 8;
  return (unsigned)(want < 1 ? 1 : (want < cap ? want : cap));
}

// examples/lightcone.jl:30-49
__global__ void __launch_bounds__(kBlock) sky_to_cartesian_kernel(const float* ra, const float* dec, const float* red, int64_t n,
                                                                  float h, const double* __restrict__ rtab, double z0, double z1,
                                                                  double dz, int64_t ntab, float* ox, float* oy, float* oz,
                                                                  unsigned long long* __restrict__ oob) {
  const int64_t step = (int64_t)gridDim.x * blockDim.x;
  unsigned bad = 0;
  for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < n; p += step) {
    float x, y, z;
    if (!sky_to_cartesian_one(ra[p], dec[p], red[p], h, rtab, z0, z1, dz, ntab, &x, &y, &z)) bad++;
    ox[p] = x;
    oy[p] = y;
    oz[p] = z;
  }
  if (bad) atomicAdd(oob, (unsigned long long)bad);
}

// examples/lightcone.jl:51-80
__global__ void __launch_bounds__(kBlock) cartesian_to_sky_kernel(const float* x, const float* y, const float* z, int64_t n, float h,
                                                                  const double* __restrict__ rtab, double z0, double dz,
                                                                  int64_t ntab, int64_t stride, int ncoarse, float* ora,
                                                                  float* odec, float* ored, unsigned long long* __restrict__ oob) {
  __shared__ double coarse[kCoarse];
  for (int j = threadIdx.x; j < ncoarse; j += blockDim.x) {
    int64_t k = (int64_t)j * stride;
    coarse[j] = __ldg(rtab + (k > ntab - 1 ? ntab - 1 : k));
  }
  __syncthreads();
  const int64_t step = (int64_t)gridDim.x * blockDim.x;
  unsigned bad = 0;
  for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < n; p += step) {
    float a, d, r;
    if (!cartesian_to_sky_one(x[p], y[p], z[p], h, rtab, coarse, ncoarse, stride, z0, dz, ntab, &a, &d, &r)) bad++;
    ora[p] = a;
    odec[p] = d;
    ored[p] = r;
  }
  if (bad) atomicAdd(oob, (unsigned long long)bad);
}

// examples/lightcone.jl:82
__global__ void __launch_bounds__(kBlock) fkp_weights_kernel(const float* nz, int64_t n, float P0, float* w) {
  const int64_t step = (int64_t)gridDim.x * blockDim.x;
  for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < n; p += step) w[p] = fkp_one(nz[p], P0);
}

__global__ void __launch_bounds__(kBlock) wrap_positions_kernel(float* x, float* y, float* z, int64_t n, float Lx, float Ly, float Lz,
                                                                float mx, float my, float mz) {
  const int64_t step = (int64_t)gridDim.x * blockDim.x;
  for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < n; p += step) {
    x[p] = wrap_one(x[p], Lx, mx);
    y[p] = wrap_one(y[p], Ly, my);
    z[p] = wrap_one(z[p], Lz, mz);
  }
}

int check_table_range(baorec_ctx* ctx, cudaStream_t st, const char* what) {
  unsigned long long h = 0;
  BR_CUDA(cudaMemcpyAsync(&h, ctx->d_oob, sizeof(h), cudaMemcpyDeviceToHost, st));
  BR_CUDA(cudaStreamSynchronize(st));
  if (h != 0) {
    set_error("%s: %llu particle(s) outside the tabulated range (the reference's interpolator raises BoundsError); "
              "their outputs are NaN",
              what, h);
    return BAOREC_ERR_OUT_OF_RANGE;
  }
  return BAOREC_OK;
}

}  // namespace
}  // namespace baorec

using namespace baorec;

extern "C" {

static int check_cosmology(const baorec_cosmology* c) {
  BR_REQUIRE(c != nullptr, "NULL pointer");
  BR_REQUIRE(c->z_tab_num >= 2 && c->z_tab_num <= (int64_t)1 << 26, "z_tab_num must be in [2, 2^26]");
  BR_REQUIRE(c->z_tab_max > c->z_tab_min && c->z_tab_min >= 0.0, "need 0 <= z_tab_min < z_tab_max");
  BR_REQUIRE(c->h > 0.0, "h must be positive");
  return BAOREC_OK;
}

int baorec_cosmo_build_table(const baorec_cosmology* c, double* h_z, double* h_r) {
  BR_TRY(check_cosmology(c));
  BR_REQUIRE(h_r != nullptr, "h_r is NULL");
  const int64_t n = c->z_tab_num;
  const double z0 = c->z_tab_min, dz = (c->z_tab_max - c->z_tab_min) / (double)(n - 1);
  const CosmoPars pars = {c->h, c->Omega_b0, c->Omega_c0, c->Omega_nu0, c->Omega_g0, c->Omega_k0, c->Omega_L0, c->w0, c->wa};
  const int64_t bad = build_distance_table(pars, z0, dz, n, h_r);
  if (bad >= 0) {
    set_error("comoving distance table is not increasing at z = %g (check the density parameters)", z0 + dz * (double)bad);
    return BAOREC_ERR_INVALID;
  }
  if (h_z) {
    for (int64_t i = 0; i < n; i++) h_z[i] = fma((double)i, dz, z0);
    h_z[n - 1] = c->z_tab_max;   // a Julia range hits its end point exactly
  }
  return BAOREC_OK;
}

int baorec_cosmo_set(baorec_ctx* ctx, const baorec_cosmology* c) {
  BR_REQUIRE(ctx != nullptr, "ctx is NULL");
  BR_TRY(check_cosmology(c));
  const int64_t n = c->z_tab_num;
  std::vector<double> r((size_t)n);
  BR_TRY(baorec_cosmo_build_table(c, nullptr, r.data()));
  BR_CUDA(cudaSetDevice(ctx->device));
  BR_CUDA(cudaDeviceSynchronize());   // no kernel may still be reading the previous table
  if (ctx->d_cosmo_r && ctx->cosmo_n != n) {
    cudaFree(ctx->d_cosmo_r);
    ctx->d_cosmo_r = nullptr;
  }
  if (!ctx->d_cosmo_r) BR_CUDA(cudaMalloc(&ctx->d_cosmo_r, (size_t)n * sizeof(double)));
  BR_CUDA(cudaMemcpy(ctx->d_cosmo_r, r.data(), (size_t)n * sizeof(double), cudaMemcpyHostToDevice));
  ctx->h_cosmo_r.swap(r);
  ctx->cosmo_n = n;
  ctx->cosmo_z0 = c->z_tab_min;
  ctx->cosmo_z1 = c->z_tab_max;
  ctx->cosmo_dz = (c->z_tab_max - c->z_tab_min) / (double)(n - 1);
  return BAOREC_OK;
}

int baorec_cosmo_tables(const baorec_ctx* ctx, double* h_z, double* h_r, int64_t cap) {
  BR_REQUIRE(ctx != nullptr, "ctx is NULL");
  BR_REQUIRE(ctx->cosmo_n > 0, "call baorec_cosmo_set first");
  BR_REQUIRE(cap >= ctx->cosmo_n, "output arrays are shorter than z_tab_num");
  for (int64_t i = 0; i < ctx->cosmo_n; i++) {
    if (h_z) h_z[i] = (i == ctx->cosmo_n - 1) ? ctx->cosmo_z1 : fma((double)i, ctx->cosmo_dz, ctx->cosmo_z0);
    if (h_r) h_r[i] = ctx->h_cosmo_r[(size_t)i];
  }
  return BAOREC_OK;
}

int baorec_sky_to_cartesian_f32(baorec_ctx* ctx, const float* d_ra, const float* d_dec, const float* d_red, int64_t n, float h,
                                float* d_x, float* d_y, float* d_z, baorec_stream stream) {
  BR_REQUIRE(ctx != nullptr, "ctx is NULL");
  BR_REQUIRE(ctx->cosmo_n > 0, "call baorec_cosmo_set first");
  BR_REQUIRE(n >= 0, "negative particle count");
  if (n == 0) return BAOREC_OK;
  BR_REQUIRE(d_ra && d_dec && d_red && d_x && d_y && d_z, "NULL pointer");
  BR_CUDA(cudaSetDevice(ctx->device));
  cudaStream_t st = (cudaStream_t)stream;
  BR_TRY(reset_oob(ctx, st));
  BR_LAUNCH(ctx, sky_to_cartesian_kernel, catalog_grid(n), kBlock, 0, st, d_ra, d_dec, d_red, n, h, ctx->d_cosmo_r,
            ctx->cosmo_z0, ctx->cosmo_z1, ctx->cosmo_dz, ctx->cosmo_n, d_x, d_y, d_z, ctx->d_oob);
  return check_table_range(ctx, st, "sky_to_cartesian");
}

int baorec_cartesian_to_sky_f32(baorec_ctx* ctx, const float* d_x, const float* d_y, const float* d_z, int64_t n, float h,
                                float* d_ra, float* d_dec, float* d_red, baorec_stream stream) {
  BR_REQUIRE(ctx != nullptr, "ctx is NULL");
  BR_REQUIRE(ctx->cosmo_n > 0, "call baorec_cosmo_set first");
  BR_REQUIRE(n >= 0, "negative particle count");
  BR_REQUIRE(h > 0.0f, "h must be positive");
  if (n == 0) return BAOREC_OK;
  BR_REQUIRE(d_ra && d_dec && d_red && d_x && d_y && d_z, "NULL pointer");
  BR_CUDA(cudaSetDevice(ctx->device));
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t ntab = ctx->cosmo_n;
  int64_t stride;
  int ncoarse;
  coarse_layout(ntab, &stride, &ncoarse);
  BR_TRY(reset_oob(ctx, st));
  BR_LAUNCH(ctx, cartesian_to_sky_kernel, catalog_grid(n), kBlock, 0, st, d_x, d_y, d_z, n, h, ctx->d_cosmo_r, ctx->cosmo_z0,
            ctx->cosmo_dz, ntab, stride, ncoarse, d_ra, d_dec, d_red, ctx->d_oob);
  return check_table_range(ctx, st, "cartesian_to_sky");
}

int baorec_fkp_weights_f32(baorec_ctx* ctx, const float* d_nz, int64_t n, float P0, float* d_w, baorec_stream stream) {
  BR_REQUIRE(ctx != nullptr, "ctx is NULL");
  BR_REQUIRE(n >= 0, "negative particle count");
  if (n == 0) return BAOREC_OK;
  BR_REQUIRE(d_nz && d_w, "NULL pointer");
  BR_CUDA(cudaSetDevice(ctx->device));
  cudaStream_t st = (cudaStream_t)stream;
  BR_LAUNCH(ctx, fkp_weights_kernel, catalog_grid(n), kBlock, 0, st, d_nz, n, P0, d_w);
  return BAOREC_OK;
}

int baorec_wrap_positions_f32(baorec_ctx* ctx, float* d_x, float* d_y, float* d_z, int64_t n, const float box_size[3],
                              const float box_min[3], baorec_stream stream) {
  BR_REQUIRE(ctx != nullptr, "ctx is NULL");
  BR_REQUIRE(n >= 0, "negative particle count");
  BR_REQUIRE(box_size && box_min, "NULL pointer");
  BR_REQUIRE(box_size[0] > 0 && box_size[1] > 0 && box_size[2] > 0, "box_size must be positive");
  if (n == 0) return BAOREC_OK;
  BR_REQUIRE(d_x && d_y && d_z, "NULL pointer");
  BR_CUDA(cudaSetDevice(ctx->device));
  cudaStream_t st = (cudaStream_t)stream;
  BR_LAUNCH(ctx, wrap_positions_kernel, catalog_grid(n), kBlock, 0, st, d_x, d_y, d_z, n, box_size[0], box_size[1],
            box_size[2], box_min[0], box_min[1], box_min[2]);
  return BAOREC_OK;
}

}  // extern "C"
