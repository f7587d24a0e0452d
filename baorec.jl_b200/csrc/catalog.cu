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
// particle, float4 accesses, four particles per thread); the 0.8 MB distance table stays L2-resident, and
// the inverse interpolation finds its interval from a 32 KB guess table plus one +-1 correction step
// instead of bisecting (catalog_math.cuh).
#include <math.h>

#include <vector>

#include "catalog_math.cuh"
#include "internal.cuh"

namespace baorec {

namespace {

using namespace baorec::catalog;

constexpr int kBlock = 256;

// Every kernel handles VEC consecutive particles per thread and iteration: VEC = 4 moves them with one
// 16-byte access per array (and gives the Float64 pipeline / the table searches four independent
// chains to overlap); VEC = 1 is the same code for unaligned arrays and for the <= 3 tail particles.
template <int VEC>
__device__ __forceinline__ void load_vec(const float* p, int64_t g, float (&v)[VEC]) {
  if (VEC == 4) {
    const float4 t = reinterpret_cast<const float4*>(p)[g];
    v[0] = t.x, v[1 % VEC] = t.y, v[2 % VEC] = t.z, v[3 % VEC] = t.w;
  } else {
    v[0] = p[g];
  }
}
template <int VEC>
__device__ __forceinline__ void store_vec(float* p, int64_t g, const float (&v)[VEC]) {
  if (VEC == 4) reinterpret_cast<float4*>(p)[g] = make_float4(v[0], v[1 % VEC], v[2 % VEC], v[3 % VEC]);
  else p[g] = v[0];
}

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

// grid-stride launches: `per_sm` resident blocks of 256 threads on each of the 148 SMs at most
inline unsigned catalog_grid(int64_t groups, int per_sm) {
  int64_t want = (groups + kBlock - 1) / kBlock;
  int64_t cap = (int64_t)148 * per_sm;
  return (unsigned)(want < 1 ? 1 : (want < cap ? want : cap));
}

// examples/lightcone.jl:30-49
template <int VEC>
__global__ void __launch_bounds__(kBlock) sky_to_cartesian_kernel(const float* ra, const float* dec, const float* red, int64_t groups,
                                                                  float h, const double* __restrict__ rtab, double z0, double z1,
                                                                  double dz, double inv_dz, int64_t ntab, float* ox, float* oy,
                                                                  float* oz, unsigned long long* __restrict__ oob) {
  const int64_t step = (int64_t)gridDim.x * blockDim.x;
  unsigned bad = 0;
  for (int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; g < groups; g += step) {
    float a[VEC], d[VEC], r[VEC], x[VEC], y[VEC], z[VEC];
    load_vec<VEC>(ra, g, a);
    load_vec<VEC>(dec, g, d);
    load_vec<VEC>(red, g, r);
#pragma unroll
    for (int i = 0; i < VEC; i++)
      if (!sky_to_cartesian_one(a[i], d[i], r[i], h, rtab, z0, z1, dz, inv_dz, ntab, &x[i], &y[i], &z[i])) bad++;
    store_vec<VEC>(ox, g, x);
    store_vec<VEC>(oy, g, y);
    store_vec<VEC>(oz, g, z);
  }
  if (bad) atomicAdd(oob, (unsigned long long)bad);
}

// examples/lightcone.jl:51-80
template <int VEC>
__global__ void __launch_bounds__(kBlock) cartesian_to_sky_kernel(const float* x, const float* y, const float* z, int64_t groups,
                                                                  float h, const double* __restrict__ rtab,
                                                                  const double* __restrict__ gtab, double r0, double r1, double inv_h,
                                                                  int corr, double z0, double dz, int ntab, float* ora, float* odec,
                                                                  float* ored, unsigned long long* __restrict__ oob) {
  const int64_t step = (int64_t)gridDim.x * blockDim.x;
  unsigned bad = 0;
  for (int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; g < groups; g += step) {
    float px[VEC], py[VEC], pz[VEC], a[VEC], d[VEC], r[VEC];
    load_vec<VEC>(x, g, px);
    load_vec<VEC>(y, g, py);
    load_vec<VEC>(z, g, pz);
    bad += cartesian_to_sky_vec<VEC>(px, py, pz, h, rtab, gtab, r0, r1, inv_h, corr, z0, dz, ntab, a, d, r);
    store_vec<VEC>(ora, g, a);
    store_vec<VEC>(odec, g, d);
    store_vec<VEC>(ored, g, r);
  }
  if (bad) atomicAdd(oob, (unsigned long long)bad);
}

// examples/lightcone.jl:82
template <int VEC>
__global__ void __launch_bounds__(kBlock) fkp_weights_kernel(const float* nz, int64_t groups, float P0, float* w) {
  const int64_t step = (int64_t)gridDim.x * blockDim.x;
  for (int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; g < groups; g += step) {
    float v[VEC];
    load_vec<VEC>(nz, g, v);
#pragma unroll
    for (int i = 0; i < VEC; i++) v[i] = fkp_one(v[i], P0);
    store_vec<VEC>(w, g, v);
  }
}

template <int VEC>
__global__ void __launch_bounds__(kBlock) wrap_positions_kernel(float* x, float* y, float* z, int64_t groups, float Lx, float Ly,
                                                                float Lz, float mx, float my, float mz) {
  const int64_t step = (int64_t)gridDim.x * blockDim.x;
  for (int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; g < groups; g += step) {
    float a[VEC], b[VEC], c[VEC];
    load_vec<VEC>(x, g, a);
    load_vec<VEC>(y, g, b);
    load_vec<VEC>(z, g, c);
#pragma unroll
    for (int i = 0; i < VEC; i++) {
      a[i] = wrap_one(a[i], Lx, mx);
      b[i] = wrap_one(b[i], Ly, my);
      c[i] = wrap_one(c[i], Lz, mz);
    }
    store_vec<VEC>(x, g, a);
    store_vec<VEC>(y, g, b);
    store_vec<VEC>(z, g, c);
  }
}

// Splits [0, n) into a float4 body (when every array is 16-byte aligned) and a scalar remainder:
// body(groups) launches the VEC = 4 kernel on the first 4 * groups particles, tail(offset, count) the
// VEC = 1 kernel on the rest.
template <class Body, class Tail>
int split_launch(int64_t n, bool aligned, Body body, Tail tail) {
  const int64_t n4 = aligned ? n / 4 : 0;
  if (n4 > 0) BR_TRY(body(n4));
  if (n - 4 * n4 > 0) BR_TRY(tail(4 * n4, n - 4 * n4));
  return BAOREC_OK;
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
  std::vector<double> g((size_t)kGuessBins + 1);
  ctx->cosmo_corr = build_inverse_guess(r.data(), n, g.data(), &ctx->cosmo_inv_h);
  if (!ctx->d_cosmo_g) BR_CUDA(cudaMalloc(&ctx->d_cosmo_g, g.size() * sizeof(double)));
  BR_CUDA(cudaMemcpy(ctx->d_cosmo_g, g.data(), g.size() * sizeof(double), cudaMemcpyHostToDevice));
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
  const double z0 = ctx->cosmo_z0, z1 = ctx->cosmo_z1, dz = ctx->cosmo_dz, inv_dz = 1.0 / ctx->cosmo_dz;
  const bool al = aligned16(d_ra) && aligned16(d_dec) && aligned16(d_red) && aligned16(d_x) && aligned16(d_y) && aligned16(d_z);
  BR_TRY(split_launch(
      n, al,
      [&](int64_t groups) -> int {
        BR_LAUNCH_NAMED(ctx, "sky_to_cartesian_kernel", sky_to_cartesian_kernel<4>, catalog_grid(groups, 8), kBlock, 0, st, d_ra,
                        d_dec, d_red, groups, h, ctx->d_cosmo_r, z0, z1, dz, inv_dz, ctx->cosmo_n, d_x, d_y, d_z, ctx->d_oob);
        return (int)BAOREC_OK;
      },
      [&](int64_t off, int64_t cnt) -> int {
        BR_LAUNCH_NAMED(ctx, "sky_to_cartesian_kernel", sky_to_cartesian_kernel<1>, catalog_grid(cnt, 8), kBlock, 0, st, d_ra + off,
                        d_dec + off, d_red + off, cnt, h, ctx->d_cosmo_r, z0, z1, dz, inv_dz, ctx->cosmo_n, d_x + off, d_y + off,
                        d_z + off, ctx->d_oob);
        return (int)BAOREC_OK;
      }));
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
  const int ntab = (int)ctx->cosmo_n;
  const double r0 = ctx->h_cosmo_r.front(), r1 = ctx->h_cosmo_r.back(), inv_h = ctx->cosmo_inv_h;
  const int corr = ctx->opt_catalog_corr >= 0 ? ctx->opt_catalog_corr : ctx->cosmo_corr;
  BR_TRY(reset_oob(ctx, st));
  const double z0 = ctx->cosmo_z0, dz = ctx->cosmo_dz;
  const bool al = aligned16(d_ra) && aligned16(d_dec) && aligned16(d_red) && aligned16(d_x) && aligned16(d_y) && aligned16(d_z);
  BR_TRY(split_launch(
      n, al,
      [&](int64_t groups) -> int {
        BR_LAUNCH_NAMED(ctx, "cartesian_to_sky_kernel", cartesian_to_sky_kernel<4>, catalog_grid(groups, 8), kBlock, 0, st, d_x, d_y,
                        d_z, groups, h, ctx->d_cosmo_r, ctx->d_cosmo_g, r0, r1, inv_h, corr, z0, dz, ntab, d_ra, d_dec, d_red,
                        ctx->d_oob);
        return (int)BAOREC_OK;
      },
      [&](int64_t off, int64_t cnt) -> int {
        BR_LAUNCH_NAMED(ctx, "cartesian_to_sky_kernel", cartesian_to_sky_kernel<1>, catalog_grid(cnt, 8), kBlock, 0, st, d_x + off,
                        d_y + off, d_z + off, cnt, h, ctx->d_cosmo_r, ctx->d_cosmo_g, r0, r1, inv_h, corr, z0, dz, ntab, d_ra + off,
                        d_dec + off, d_red + off, ctx->d_oob);
        return (int)BAOREC_OK;
      }));
  return check_table_range(ctx, st, "cartesian_to_sky");
}

int baorec_fkp_weights_f32(baorec_ctx* ctx, const float* d_nz, int64_t n, float P0, float* d_w, baorec_stream stream) {
  BR_REQUIRE(ctx != nullptr, "ctx is NULL");
  BR_REQUIRE(n >= 0, "negative particle count");
  if (n == 0) return BAOREC_OK;
  BR_REQUIRE(d_nz && d_w, "NULL pointer");
  BR_CUDA(cudaSetDevice(ctx->device));
  cudaStream_t st = (cudaStream_t)stream;
  return split_launch(
      n, aligned16(d_nz) && aligned16(d_w),
      [&](int64_t groups) -> int {
        BR_LAUNCH_NAMED(ctx, "fkp_weights_kernel", fkp_weights_kernel<4>, catalog_grid(groups, 16), kBlock, 0, st, d_nz, groups, P0, d_w);
        return (int)BAOREC_OK;
      },
      [&](int64_t off, int64_t cnt) -> int {
        BR_LAUNCH_NAMED(ctx, "fkp_weights_kernel", fkp_weights_kernel<1>, catalog_grid(cnt, 16), kBlock, 0, st, d_nz + off, cnt, P0,
                        d_w + off);
        return (int)BAOREC_OK;
      });
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
  const float Lx = box_size[0], Ly = box_size[1], Lz = box_size[2], mx = box_min[0], my = box_min[1], mz = box_min[2];
  return split_launch(
      n, aligned16(d_x) && aligned16(d_y) && aligned16(d_z),
      [&](int64_t groups) -> int {
        BR_LAUNCH_NAMED(ctx, "wrap_positions_kernel", wrap_positions_kernel<4>, catalog_grid(groups, 16), kBlock, 0, st, d_x, d_y, d_z,
                        groups, Lx, Ly, Lz, mx, my, mz);
        return (int)BAOREC_OK;
      },
      [&](int64_t off, int64_t cnt) -> int {
        BR_LAUNCH_NAMED(ctx, "wrap_positions_kernel", wrap_positions_kernel<1>, catalog_grid(cnt, 16), kBlock, 0, st, d_x + off,
                        d_y + off, d_z + off, cnt, Lx, Ly, Lz, mx, my, mz);
        return (int)BAOREC_OK;
      });
}

}  // extern "C"
