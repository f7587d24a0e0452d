// extern "C" entry points for the mesh set-up, iterative solver, read-back and the
// host-buffer pipelines (run! / reconstructed_positions equivalents).
#include "internal.cuh"

#include <chrono>

#include "batch.cuh"

using namespace baorec;

namespace baorec {

static int check_params(const baorec_params* p) {
  BR_REQUIRE(p != nullptr, "params is NULL");
  BR_REQUIRE(p->bias != 0.f, "bias must be non-zero");
  BR_REQUIRE(p->mas == BAOREC_MAS_CIC || p->mas == BAOREC_MAS_TSC || p->mas == BAOREC_MAS_PCS, "unknown mas");
  BR_REQUIRE(p->n_iter >= 0 && p->jacobi_niterations >= 0 && p->vcycle_niterations >= 0, "negative iteration count");
  return BAOREC_OK;
}

static int check_catalog(int64_t n, const void* x, const void* y, const void* z, const void* w, const char* what) {
  if (n < 0) {
    set_error("invalid argument: %s count < 0", what);
    return BAOREC_ERR_INVALID;
  }
  if (n > 0 && !(x && y && z && w)) {
    set_error("invalid argument: NULL %s array", what);
    return BAOREC_ERR_INVALID;
  }
  return BAOREC_OK;
}

// Displacement meshes of `mesh` into RX/RY/RZ, then the fused gather + read_shifts epilogue.
static int read_common(baorec_ctx* ctx, const baorec_params* p, int algorithm, const float* mesh, const float* x,
                       const float* y, const float* z, int64_t n, int field, int positions, float* ox, float* oy,
                       float* oz, cudaStream_t st, bool use_kcache = false, bool trusted = false) {
  if (algorithm == BAOREC_MULTIGRID && ctx->opt_mg_fd_gradient && p->mas == BAOREC_MAS_CIC) {
    // Psi = grad(phi) by finite differences: one gather over phi, no transforms (src/multigrid.jl:759, commented there)
    BR_TRY(reset_oob(ctx, st));
    BR_CUDA(cudaEventRecord(ctx->ev[5], st));
    BR_TRY(gather_fd(ctx, mesh, x, y, z, n, ox, oy, oz, field, p->f, p->has_los, p->los, positions, st));
    return check_oob(ctx, st, "read_shifts");
  }
  float *px, *py, *pz;
  BR_TRY(need_t(ctx, BUF_RX, ctx->M, &px));
  BR_TRY(need_t(ctx, BUF_RY, ctx->M, &py));
  BR_TRY(need_t(ctx, BUF_RZ, ctx->M, &pz));
  // `trusted`: the caller asserts that `mesh` is the unmodified cached result, so displacement
  // meshes computed from it by an earlier call are still its displacement meshes
  const bool reuse = trusted && ctx->disp_valid && ctx->disp_algo == algorithm && ctx->disp_mesh == mesh &&
                     !own_fft_available(ctx);
  BR_TRY(reset_oob(ctx, st));
  if (!reuse) BR_TRY(gather_prebin(ctx, x, y, z, n, p->mas, st));  // the sort overlaps the transforms
  if (!reuse) {
    ctx->disp_valid = false;
    BR_TRY(displacement_meshes(ctx, mesh, algorithm, px, py, pz, st, use_kcache));
    if (trusted) {
      ctx->disp_valid = true;
      ctx->disp_algo = algorithm;
      ctx->disp_mesh = mesh;
    }
  }
  BR_CUDA(cudaEventRecord(ctx->ev[5], st));
  BR_TRY(gather3(ctx, px, py, pz, x, y, z, n, ox, oy, oz, p->mas, field, p->f, p->has_los, p->los, positions, st));
  return check_oob(ctx, st, "read_shifts");
}

// ---- many catalogs per process: the host pipeline, software-pipelined over a batch --------------------------------
// README.md:11 of the reference: "one process, many reconstructions" (the examples loop over mocks).  For each
// catalog this is baorec_run_host_f32 followed by baorec_read_host_f32 on the same catalog -- periodic box, no
// randoms -- but the PCIe transfers of neighbouring catalogs overlap the solve: the upload of catalog i+1 (copy
// stream) and the download of the results of catalog i-1 (download stream) run while catalog i is reconstructed
// (main stream); catalogs and results are double-buffered on the device, the read-back uses the positions
// already resident (no second upload) and the tile sort run! made (same device arrays -> hash hit).
// The catalogs come from `src` (batch.cuh): arrays the caller holds, or files read by an I/O thread.  Sizes need not
// be known in advance: a staging buffer grows when a larger catalog arrives, at a point where nothing in flight uses it.
int batch_pipeline(baorec_ctx* ctx, const baorec_params* p, int algorithm, int n_catalogs, int field, int shifts_only,
                   BatchSource& src) {
  BR_TRY(check_params(p));
  BR_REQUIRE(algorithm == BAOREC_ITERATIVE || algorithm == BAOREC_MULTIGRID, "unknown algorithm");
  BR_REQUIRE(field >= BAOREC_FIELD_DISP && field <= BAOREC_FIELD_SUM, "unknown field");
  BR_REQUIRE(n_catalogs >= 1, "n_catalogs");
  cudaStream_t st = ctx->own_stream, up = ctx->copy_stream;
  const BufId part_id[2] = {BUF_PART, BUF_PART2};
  float *part[2] = {nullptr, nullptr}, *out = nullptr, *mesh, *px, *py, *pz;
  int64_t cap_part[2] = {0, 0}, cap_out = 0;
  BR_TRY(need_t(ctx, BUF_CACHE, ctx->M, &mesh));
  BR_TRY(need_t(ctx, BUF_RX, ctx->M, &px));
  BR_TRY(need_t(ctx, BUF_RY, ctx->M, &py));
  BR_TRY(need_t(ctx, BUF_RZ, ctx->M, &pz));
  cudaStream_t dn = nullptr;
  cudaEvent_t ev_up[2] = {nullptr, nullptr}, ev_gather[2] = {nullptr, nullptr}, ev_dn[2] = {nullptr, nullptr};
  BR_CUDA(cudaStreamCreateWithFlags(&dn, cudaStreamNonBlocking));
  for (int b = 0; b < 2; b++) {
    BR_CUDA(cudaEventCreateWithFlags(&ev_up[b], cudaEventDisableTiming));
    BR_CUDA(cudaEventCreateWithFlags(&ev_gather[b], cudaEventDisableTiming));
    BR_CUDA(cudaEventCreateWithFlags(&ev_dn[b], cudaEventDisableTiming));
  }
  ctx->cache_valid = false;
  ctx->kcache_valid = false;
  ctx->disp_valid = false;
  BatchItem item[2];
  // Device staging for a catalog of n particles in half b.  part[b] was last read by the gather of the catalog two
  // before, which has completed (check_oob synchronises the main stream); the result halves are re-cut only after
  // the downloads in flight have drained.  With max_rows() known both happen once, before the first upload.
  auto reserve = [&](int b, int64_t n) -> int {
    if (n > cap_part[b]) {
      const int64_t cap = cap_part[b] == 0 ? n : n + n / 16;
      BR_TRY(need_t(ctx, part_id[b], (size_t)cap * 4, &part[b]));
      cap_part[b] = cap;
    }
    if (n > cap_out) {
      BR_CUDA(cudaStreamSynchronize(dn));
      const int64_t cap = cap_out == 0 ? n : n + n / 16;
      BR_TRY(need_t(ctx, BUF_OUT, (size_t)cap * 6, &out));
      cap_out = cap;
    }
    return BAOREC_OK;
  };
  auto upload = [&](int i) -> int {  // catalog i -> part[i & 1] (x, y, z, w at stride n) on the copy stream
    const int b = i & 1;
    BatchItem& it = item[b];
    BR_TRY(src.acquire(i, &it));
    BR_REQUIRE(it.n > 0 && it.x && it.y && it.z && it.w && it.ox && it.oy && it.oz, "catalog arrays");
    BR_TRY(reserve(b, it.n));
    const float* h[4] = {it.x, it.y, it.z, it.w};
    for (int c = 0; c < 4; c++)
      BR_CUDA(cudaMemcpyAsync(part[b] + (size_t)c * it.n, h[c], (size_t)it.n * sizeof(float), cudaMemcpyHostToDevice, up));
    BR_CUDA(cudaEventRecord(ev_up[b], up));
    return BAOREC_OK;
  };
  auto body = [&]() -> int {
    if (src.max_rows() > 0) {
      BR_TRY(reserve(0, src.max_rows()));
      BR_TRY(reserve(1, src.max_rows()));
    }
    // everything queued on the main stream so far (an earlier call's work on these buffers) precedes the first upload
    BR_CUDA(cudaEventRecord(ev_gather[0], st));
    BR_CUDA(cudaStreamWaitEvent(up, ev_gather[0], 0));
    BR_TRY(upload(0));
    for (int i = 0; i < n_catalogs; i++) {
      const int b = i & 1;
      if (i + 1 < n_catalogs) {
        // part[b ^ 1] was last read by the gather of catalog i-1, which has completed (check_oob synchronises the
        // main stream); the event keeps the order explicit on the device as well
        if (i >= 1) BR_CUDA(cudaStreamWaitEvent(up, ev_gather[b ^ 1], 0));
        BR_TRY(upload(i + 1));
      }
      const BatchItem& it = item[b];
      const int64_t ni = it.n;
      float *dx = part[b], *dy = part[b] + ni, *dz = part[b] + 2 * ni, *dw = part[b] + 3 * ni;
      float* ob = out + (size_t)b * 3 * (size_t)cap_out;
      BR_CUDA(cudaStreamWaitEvent(st, ev_up[b], 0));
      BR_CUDA(cudaMemsetAsync(mesh, 0, ctx->M * sizeof(float), st));
      ctx->want_kcache = true;
      int s = algorithm == BAOREC_MULTIGRID
                  ? reconstructed_potential(ctx, p, mesh, dx, dy, dz, dw, ni, nullptr, nullptr, nullptr, nullptr, 0, st)
                  : reconstructed_overdensity(ctx, p, mesh, dx, dy, dz, dw, ni, nullptr, nullptr, nullptr, nullptr, 0, st);
      ctx->want_kcache = false;
      if (s != BAOREC_OK) return s;
      if (ctx->last_wrapped > 0) {  // cic!(wrap = true) mutates the caller's positions (src/mas.jl:8-10)
        float* hdst[3] = {it.x, it.y, it.z};
        for (int c = 0; c < 3; c++)
          BR_CUDA(cudaMemcpyAsync(hdst[c], part[b] + (size_t)c * ni, (size_t)ni * sizeof(float), cudaMemcpyDeviceToHost, st));
      }
      ctx->disp_valid = false;
      BR_TRY(displacement_meshes(ctx, mesh, algorithm, px, py, pz, st, /*use_kcache=*/true));
      if (i >= 2) BR_CUDA(cudaStreamWaitEvent(st, ev_dn[b], 0));  // this half still holds the results of catalog i-2
      BR_TRY(reset_oob(ctx, st));
      BR_TRY(gather3(ctx, px, py, pz, dx, dy, dz, ni, ob, ob + ni, ob + 2 * ni, p->mas, field, p->f, p->has_los, p->los,
                     shifts_only ? 0 : 1, st));
      BR_CUDA(cudaEventRecord(ev_gather[b], st));
      BR_TRY(check_oob(ctx, st, "read_shifts"));
      BR_CUDA(cudaStreamWaitEvent(dn, ev_gather[b], 0));
      float* hdst[3] = {it.ox, it.oy, it.oz};
      for (int c = 0; c < 3; c++)
        BR_CUDA(cudaMemcpyAsync(hdst[c], ob + (size_t)c * ni, (size_t)ni * sizeof(float), cudaMemcpyDeviceToHost, dn));
      BR_CUDA(cudaEventRecord(ev_dn[b], dn));
      if (i >= 1) {  // the download of catalog i-1 ran under this catalog's reconstruction
        BR_CUDA(cudaEventSynchronize(ev_dn[b ^ 1]));
        BR_TRY(src.release(i - 1));
      }
    }
    BR_CUDA(cudaStreamSynchronize(dn));
    BR_CUDA(cudaStreamSynchronize(st));
    return src.release(n_catalogs - 1);
  };
  const int status = body();
  if (status != BAOREC_OK) {  // drain whatever is in flight before the buffers are reused
    cudaStreamSynchronize(up);
    cudaStreamSynchronize(dn);
    cudaStreamSynchronize(st);
  }
  for (int b = 0; b < 2; b++) {
    cudaEventDestroy(ev_up[b]);
    cudaEventDestroy(ev_gather[b]);
    cudaEventDestroy(ev_dn[b]);
  }
  cudaStreamDestroy(dn);
  ctx->sortc_valid = false;  // the sort belongs to a staging buffer the next call overwrites
  ctx->cache_valid = status == BAOREC_OK;
  ctx->disp_valid = false;
  return status;
}

}  // namespace baorec

extern "C" {

int baorec_smooth_f32(baorec_ctx* ctx, float* d_mesh, float smoothing_radius, baorec_stream stream) {
  BR_NEED_PLAN(ctx);
  BR_REQUIRE(d_mesh != nullptr, "mesh is NULL");
  return smooth(ctx, d_mesh, smoothing_radius, (cudaStream_t)stream);
}

int baorec_setup_overdensity_f32(baorec_ctx* ctx, const baorec_params* p, float* d_mesh, float* d_x, float* d_y,
                                 float* d_z, const float* d_w, int64_t n, float* d_rx, float* d_ry, float* d_rz,
                                 const float* d_rw, int64_t n_ran, int wrap, baorec_stream stream) {
  BR_NEED_PLAN(ctx);
  BR_TRY(check_params(p));
  BR_REQUIRE(d_mesh != nullptr, "mesh is NULL");
  BR_TRY(check_catalog(n, d_x, d_y, d_z, d_w, "data"));
  BR_TRY(check_catalog(n_ran, d_rx, d_ry, d_rz, d_rw, "randoms"));
  return setup_overdensity(ctx, p, d_mesh, d_x, d_y, d_z, d_w, n, d_rx, d_ry, d_rz, d_rw, n_ran, wrap,
                           (cudaStream_t)stream);
}

int baorec_iterate_f32(baorec_ctx* ctx, float* d_delta_r, const float* d_delta_s, int iter, float beta,
                       const float* h_los, baorec_stream stream) {
  BR_NEED_PLAN(ctx);
  BR_REQUIRE(d_delta_r && d_delta_s && iter >= 1, "iterate arguments");
  return iterate(ctx, d_delta_r, d_delta_s, iter, beta, h_los, (cudaStream_t)stream);
}

int baorec_reconstructed_overdensity_f32(baorec_ctx* ctx, const baorec_params* p, float* d_mesh, float* d_x,
                                         float* d_y, float* d_z, const float* d_w, int64_t n, float* d_rx,
                                         float* d_ry, float* d_rz, const float* d_rw, int64_t n_ran,
                                         baorec_stream stream) {
  BR_NEED_PLAN(ctx);
  BR_TRY(check_params(p));
  BR_REQUIRE(d_mesh != nullptr, "mesh is NULL");
  BR_TRY(check_catalog(n, d_x, d_y, d_z, d_w, "data"));
  BR_TRY(check_catalog(n_ran, d_rx, d_ry, d_rz, d_rw, "randoms"));
  return reconstructed_overdensity(ctx, p, d_mesh, d_x, d_y, d_z, d_w, n, d_rx, d_ry, d_rz, d_rw, n_ran,
                                   (cudaStream_t)stream);
}

int baorec_reconstructed_potential_f32(baorec_ctx* ctx, const baorec_params* p, float* d_phi, float* d_x, float* d_y,
                                       float* d_z, const float* d_w, int64_t n, float* d_rx, float* d_ry, float* d_rz,
                                       const float* d_rw, int64_t n_ran, baorec_stream stream) {
  BR_NEED_PLAN(ctx);
  BR_TRY(check_params(p));
  BR_REQUIRE(d_phi != nullptr, "phi is NULL");
  BR_TRY(check_catalog(n, d_x, d_y, d_z, d_w, "data"));
  BR_TRY(check_catalog(n_ran, d_rx, d_ry, d_rz, d_rw, "randoms"));
  return reconstructed_potential(ctx, p, d_phi, d_x, d_y, d_z, d_w, n, d_rx, d_ry, d_rz, d_rw, n_ran,
                                 (cudaStream_t)stream);
}

int baorec_displacement_meshes_f32(baorec_ctx* ctx, const float* d_mesh, int algorithm, float* d_psix, float* d_psiy,
                                   float* d_psiz, baorec_stream stream) {
  BR_NEED_PLAN(ctx);
  BR_REQUIRE(d_mesh && d_psix && d_psiy && d_psiz, "NULL mesh pointer");
  BR_REQUIRE(algorithm == BAOREC_ITERATIVE || algorithm == BAOREC_MULTIGRID, "unknown algorithm");
  return displacement_meshes(ctx, d_mesh, algorithm, d_psix, d_psiy, d_psiz, (cudaStream_t)stream);
}

int baorec_compute_displacements_f32(baorec_ctx* ctx, const float* d_mesh, int algorithm, const float* d_x,
                                     const float* d_y, const float* d_z, int64_t n, float* d_px, float* d_py,
                                     float* d_pz, int mas, baorec_stream stream) {
  BR_NEED_PLAN(ctx);
  BR_REQUIRE(d_mesh != nullptr, "mesh is NULL");
  BR_REQUIRE(algorithm == BAOREC_ITERATIVE || algorithm == BAOREC_MULTIGRID, "unknown algorithm");
  BR_REQUIRE(n >= 0 && (n == 0 || (d_x && d_y && d_z && d_px && d_py && d_pz)), "particle arrays");
  baorec_params p = {};
  p.bias = 1.f;
  p.mas = mas;
  p.has_los = 1;
  return read_common(ctx, &p, algorithm, d_mesh, d_x, d_y, d_z, n, BAOREC_FIELD_DISP, 0, d_px, d_py, d_pz,
                     (cudaStream_t)stream);
}

int baorec_read_shifts_f32(baorec_ctx* ctx, const baorec_params* p, int algorithm, const float* d_mesh,
                           const float* d_x, const float* d_y, const float* d_z, int64_t n, int field, float* d_sx,
                           float* d_sy, float* d_sz, baorec_stream stream) {
  BR_NEED_PLAN(ctx);
  BR_TRY(check_params(p));
  BR_REQUIRE(d_mesh != nullptr, "mesh is NULL");
  BR_REQUIRE(field >= BAOREC_FIELD_DISP && field <= BAOREC_FIELD_SUM, "unknown field");
  BR_REQUIRE(n >= 0 && (n == 0 || (d_x && d_y && d_z && d_sx && d_sy && d_sz)), "particle arrays");
  return read_common(ctx, p, algorithm, d_mesh, d_x, d_y, d_z, n, field, 0, d_sx, d_sy, d_sz, (cudaStream_t)stream);
}

// read_shifts / reconstructed_positions against recon.result_cache: the caller asserts that d_mesh
// is the (unmodified) mesh returned by the last reconstructed_overdensity! on this context, so the
// delta_k kept by that solve is used and the forward transform of the mesh is skipped.
int baorec_read_result_cache_f32(baorec_ctx* ctx, const baorec_params* p, int algorithm, const float* d_mesh,
                                 const float* d_x, const float* d_y, const float* d_z, int64_t n, int field,
                                 int positions, float* d_ox, float* d_oy, float* d_oz, baorec_stream stream) {
  BR_NEED_PLAN(ctx);
  BR_TRY(check_params(p));
  BR_REQUIRE(d_mesh != nullptr, "mesh is NULL");
  BR_REQUIRE(field >= BAOREC_FIELD_DISP && field <= BAOREC_FIELD_SUM, "unknown field");
  BR_REQUIRE(n >= 0 && (n == 0 || (d_x && d_y && d_z && d_ox && d_oy && d_oz)), "particle arrays");
  const bool hit = ctx->kcache_valid && ctx->kcache_mesh == d_mesh && algorithm == BAOREC_ITERATIVE;
  // MultigridRecon keeps no phi_k, but the displacement meshes of the asserted-unmodified mesh are reusable
  const bool trusted = hit || (algorithm == BAOREC_MULTIGRID && ctx->mg_result_mesh == d_mesh);
  return read_common(ctx, p, algorithm, d_mesh, d_x, d_y, d_z, n, field, positions, d_ox, d_oy, d_oz,
                     (cudaStream_t)stream, hit, trusted);
}

int baorec_reconstructed_positions_f32(baorec_ctx* ctx, const baorec_params* p, int algorithm, const float* d_mesh,
                                       const float* d_x, const float* d_y, const float* d_z, int64_t n, int field,
                                       float* d_ox, float* d_oy, float* d_oz, baorec_stream stream) {
  BR_NEED_PLAN(ctx);
  BR_TRY(check_params(p));
  BR_REQUIRE(d_mesh != nullptr, "mesh is NULL");
  BR_REQUIRE(field >= BAOREC_FIELD_DISP && field <= BAOREC_FIELD_SUM, "unknown field");
  BR_REQUIRE(n >= 0 && (n == 0 || (d_x && d_y && d_z && d_ox && d_oy && d_oz)), "particle arrays");
  return read_common(ctx, p, algorithm, d_mesh, d_x, d_y, d_z, n, field, 1, d_ox, d_oy, d_oz, (cudaStream_t)stream);
}

int baorec_set_option(baorec_ctx* ctx, const char* name, int64_t value) {
  BR_REQUIRE(ctx != nullptr && name != nullptr, "NULL argument");
  std::string s(name);
  if (s == "bin_min_particles") ctx->opt_bin_min_particles = value;
  else if (s == "fuse_kspace") ctx->opt_fuse_kspace = (int)value;
  else if (s == "keep_delta_k") ctx->opt_keep_delta_k = (int)value;
  else if (s == "own_fft") ctx->opt_own_fft = (int)value;
  else if (s == "fft_split_planes") ctx->opt_fft_split = (int)value;
  else if (s == "gather_tiles") ctx->opt_gather_tiles = (int)value;
  else if (s == "gather_stage") ctx->opt_gather_stage = (int)value;
  else if (s == "fft_tile_cols") ctx->opt_fft_tile_cols = (int)value;
  else if (s == "fft_prefetch") ctx->opt_fft_prefetch = (int)value;
  else if (s == "scatter_pairs") ctx->opt_scatter_pairs = (int)value;
  else if (s == "scatter_tiles") ctx->opt_scatter_tiles = (int)value;
  else if (s == "batch_slots") ctx->opt_batch_slots = (int)value;
  else if (s == "deterministic_scatter") ctx->opt_det_scatter = (int)value;
  else if (s == "unified_sort") {
    ctx->opt_unified_sort = (int)value;
    ctx->sortc_valid = false;
  }
  else if (s == "bin_zg_scatter") ctx->opt_zg_scatter = (int)value;
  else if (s == "bin_zg_gather") ctx->opt_zg_gather = (int)value;
  else if (s == "mg_kernel") ctx->opt_mg_kernel = (int)value;
  else if (s == "mg_ring") ctx->opt_mg_ring = (int)value;
  else if (s == "overlap_sort") ctx->opt_overlap_sort = (int)value;
  else if (s == "a2a_chunks") ctx->opt_a2a_chunks = (int)value;
  else if (s == "comm_split") ctx->opt_comm_split = (int)value;
  else if (s == "push_sm") ctx->opt_push_sm = (int)value;
  else if (s == "peer_halo") ctx->opt_peer_halo = (int)value;
  else if (s == "dist_exchange") {
    ctx->opt_dist_exchange = value != 0;
    baorec::dist_refresh_mode(ctx);
  }
  else if (s == "mg_coarse") ctx->opt_mg_coarse = (int)value;
  else if (s == "mg_remove_mean") ctx->opt_mg_remove_mean = (int)value;
  else if (s == "mg_fd_gradient") ctx->opt_mg_fd_gradient = (int)value;
  else if (s == "mg_bulk") ctx->opt_mg_bulk = (int)value;
  else if (s == "catalog_corr") ctx->opt_catalog_corr = (int)value;   // test hook (catalog.cu): -1 = measured value
  else if (s == "mg_slab_min_cells") {
    ctx->opt_mg_slab_min_cells = value;
    ctx->dlevels.clear();
  }
  else {
    set_error("baorec_set_option: unknown option '%s'", name);
    return BAOREC_ERR_INVALID;
  }
  return BAOREC_OK;
}

// ---- host pipelines ---------------------------------------------------------------------------------
int baorec_run_host_f32(baorec_ctx* ctx, const baorec_params* p, int algorithm, float* h_x, float* h_y, float* h_z,
                        const float* h_w, int64_t n, const float* h_rx, const float* h_ry, const float* h_rz,
                        const float* h_rw, int64_t n_ran, float* h_mesh_out, float box_size_out[3],
                        float box_min_out[3]) {
  BR_NEED_PLAN(ctx);
  BR_TRY(check_params(p));
  BR_REQUIRE(algorithm == BAOREC_ITERATIVE || algorithm == BAOREC_MULTIGRID, "unknown algorithm");
  BR_TRY(check_catalog(n, h_x, h_y, h_z, h_w, "data"));
  BR_TRY(check_catalog(n_ran, h_rx, h_ry, h_rz, h_rw, "randoms"));
  cudaStream_t st = ctx->own_stream;
  ctx->cache_valid = false;
  ctx->kcache_valid = false;
  ctx->disp_valid = false;
  float *dp, *dr = nullptr, *mesh;
  BR_TRY(need_t(ctx, BUF_PART, (size_t)(n > 0 ? n : 1) * 4, &dp));
  if (n_ran > 0) BR_TRY(need_t(ctx, BUF_PART2, (size_t)n_ran * 4, &dr));
  BR_TRY(need_t(ctx, BUF_CACHE, ctx->M, &mesh));
  BR_CUDA(cudaEventRecord(ctx->ev[0], st));
  const float* hsrc[4] = {h_x, h_y, h_z, h_w};
  for (int c = 0; c < 4 && n > 0; c++)
    BR_CUDA(cudaMemcpyAsync(dp + (size_t)c * n, hsrc[c], (size_t)n * sizeof(float), cudaMemcpyHostToDevice, st));
  const float* hrs[4] = {h_rx, h_ry, h_rz, h_rw};
  for (int c = 0; c < 4 && n_ran > 0; c++)
    BR_CUDA(cudaMemcpyAsync(dr + (size_t)c * n_ran, hrs[c], (size_t)n_ran * sizeof(float), cudaMemcpyHostToDevice, st));
  BR_CUDA(cudaMemsetAsync(mesh, 0, ctx->M * sizeof(float), st));
  if (n_ran > 0) {
    // run! with randoms overrides the box: setup_box(rand..., 500)  (src/recon.jl:172, 253)
    float L[3], mn[3];
    BR_TRY(setup_box_dev(ctx, dr, dr + n_ran, dr + 2 * n_ran, n_ran, p->box_pad, L, mn, st));
    BR_TRY(baorec_set_box(ctx, L, mn));
  }
  if (box_size_out && box_min_out)
    for (int a = 0; a < 3; a++) {
      box_size_out[a] = ctx->L[a];
      box_min_out[a] = ctx->mn[a];
    }
  BR_CUDA(cudaEventRecord(ctx->ev[1], st));
  float* rx = dr;
  float* ry = dr ? dr + n_ran : nullptr;
  float* rz = dr ? dr + 2 * n_ran : nullptr;
  float* rw = dr ? dr + 3 * n_ran : nullptr;
  int s;
  ctx->want_kcache = true;
  if (algorithm == BAOREC_MULTIGRID)
    s = reconstructed_potential(ctx, p, mesh, dp, dp + n, dp + 2 * n, dp + 3 * n, n, rx, ry, rz, rw, n_ran, st);
  else
    s = reconstructed_overdensity(ctx, p, mesh, dp, dp + n, dp + 2 * n, dp + 3 * n, n, rx, ry, rz, rw, n_ran, st);
  ctx->want_kcache = false;
  if (s != BAOREC_OK) return s;
  BR_CUDA(cudaEventRecord(ctx->ev[2], st));
  if (n_ran == 0 && n > 0 && ctx->last_wrapped > 0) {
    // cic!(wrap = true) mutates the caller's positions (src/mas.jl:8-10): copy them back
    // (only when at least one particle was actually wrapped).
    float* hdst[3] = {h_x, h_y, h_z};
    for (int c = 0; c < 3; c++)
      BR_CUDA(cudaMemcpyAsync(hdst[c], dp + (size_t)c * n, (size_t)n * sizeof(float), cudaMemcpyDeviceToHost, st));
  }
  if (h_mesh_out) BR_CUDA(cudaMemcpyAsync(h_mesh_out, mesh, ctx->M * sizeof(float), cudaMemcpyDeviceToHost, st));
  BR_CUDA(cudaEventRecord(ctx->ev[3], st));
  BR_CUDA(cudaStreamSynchronize(st));
  ctx->cache_valid = true;
  ctx->n_stage = 3;
  for (int i = 0; i < 3; i++) BR_CUDA(cudaEventElapsedTime(&ctx->stage_ms[i], ctx->ev[i], ctx->ev[i + 1]));
  return BAOREC_OK;
}

int baorec_read_host_f32(baorec_ctx* ctx, const baorec_params* p, int algorithm, const float* h_mesh_or_null,
                         const float* h_x, const float* h_y, const float* h_z, int64_t n, int field, int shifts_only,
                         float* h_ox, float* h_oy, float* h_oz) {
  BR_NEED_PLAN(ctx);
  BR_TRY(check_params(p));
  BR_REQUIRE(algorithm == BAOREC_ITERATIVE || algorithm == BAOREC_MULTIGRID, "unknown algorithm");
  BR_REQUIRE(field >= BAOREC_FIELD_DISP && field <= BAOREC_FIELD_SUM, "unknown field");
  BR_REQUIRE(n >= 0 && (n == 0 || (h_x && h_y && h_z && h_ox && h_oy && h_oz)), "particle arrays");
  cudaStream_t st = ctx->own_stream;
  float *mesh, *dp, *dout;
  BR_TRY(need_t(ctx, BUF_CACHE, ctx->M, &mesh));
  if (h_mesh_or_null) {
    BR_CUDA(cudaMemcpyAsync(mesh, h_mesh_or_null, ctx->M * sizeof(float), cudaMemcpyHostToDevice, st));
    ctx->cache_valid = true;
    ctx->kcache_valid = false;
    ctx->disp_valid = false;
  }
  if (!ctx->cache_valid) {
    set_error("baorec_read_host_f32: no cached result mesh (call baorec_run_host_f32 first or pass a mesh)");
    return BAOREC_ERR_INVALID;
  }
  size_t nn = (size_t)(n > 0 ? n : 1);
  BR_TRY(need_t(ctx, BUF_OUT, nn * 6, &dout));
  dp = dout + 3 * nn;
  // A catalog of the size run! scattered is uploaded to the very device arrays run! used (x, y, z at
  // stride n in BUF_PART): if it is the same catalog -- decided by the content hash in gather3 --
  // the tile sort made for the scatter is reused instead of sorting again.
  if (ctx->sortc_valid && ctx->sortc_n == n && n > 0 && ctx->sortc_x == (const float*)ctx->bufs[BUF_PART].p &&
      ctx->bufs[BUF_PART].bytes >= (size_t)n * 4 * sizeof(float)) {
    dp = (float*)ctx->bufs[BUF_PART].p;
    nn = (size_t)n;
  }
  BR_CUDA(cudaEventRecord(ctx->ev[4], st));
  const float* hsrc[3] = {h_x, h_y, h_z};
  // the upload of the positions (copy stream) overlaps the displacement-mesh transforms (main stream)
  for (int c = 0; c < 3 && n > 0; c++)
    BR_CUDA(cudaMemcpyAsync(dp + c * nn, hsrc[c], (size_t)n * sizeof(float), cudaMemcpyHostToDevice,
                            ctx->copy_stream));
  BR_TRY(reset_oob(ctx, ctx->copy_stream));
  BR_CUDA(cudaEventRecord(ctx->ev_copy, ctx->copy_stream));
  // ... and so does the tile sort of the uploaded positions (side stream, ordered after the copies)
  BR_TRY(gather_prebin(ctx, dp, dp + nn, dp + 2 * nn, n, p->mas, ctx->copy_stream));
  {
    float *px, *py, *pz;
    BR_TRY(need_t(ctx, BUF_RX, ctx->M, &px));
    BR_TRY(need_t(ctx, BUF_RY, ctx->M, &py));
    BR_TRY(need_t(ctx, BUF_RZ, ctx->M, &pz));
    // the cached mesh is library-owned: displacement meshes of an earlier read are still valid
    if (!(ctx->disp_valid && ctx->disp_algo == algorithm && ctx->disp_mesh == mesh && !own_fft_available(ctx))) {
      ctx->disp_valid = false;
      BR_TRY(displacement_meshes(ctx, mesh, algorithm, px, py, pz, st, /*use_kcache=*/true));
      ctx->disp_valid = true;
      ctx->disp_algo = algorithm;
      ctx->disp_mesh = mesh;
    }
    BR_CUDA(cudaEventRecord(ctx->ev[5], st));
    BR_CUDA(cudaStreamWaitEvent(st, ctx->ev_copy, 0));
    BR_TRY(gather3(ctx, px, py, pz, dp, dp + nn, dp + 2 * nn, n, dout, dout + nn, dout + 2 * nn, p->mas, field, p->f,
                   p->has_los, p->los, shifts_only ? 0 : 1, st));
    BR_TRY(check_oob(ctx, st, "read_shifts"));
  }
  BR_CUDA(cudaEventRecord(ctx->ev[6], st));
  float* hdst[3] = {h_ox, h_oy, h_oz};
  for (int c = 0; c < 3 && n > 0; c++)
    BR_CUDA(cudaMemcpyAsync(hdst[c], dout + c * nn, (size_t)n * sizeof(float), cudaMemcpyDeviceToHost, st));
  BR_CUDA(cudaEventRecord(ctx->ev[7], st));
  BR_CUDA(cudaStreamSynchronize(st));
  float a = 0, b = 0, c2 = 0;
  BR_CUDA(cudaEventElapsedTime(&a, ctx->ev[4], ctx->ev[5]));
  BR_CUDA(cudaEventElapsedTime(&b, ctx->ev[5], ctx->ev[6]));
  BR_CUDA(cudaEventElapsedTime(&c2, ctx->ev[6], ctx->ev[7]));
  ctx->stage_ms[3] = a;
  ctx->stage_ms[4] = b;
  ctx->stage_ms[5] = c2;
  ctx->n_stage = 6;
  return BAOREC_OK;
}

// ---- many catalogs per process (batch_pipeline above) --------------------------------------------------------------
int baorec_batch_host_f32(baorec_ctx* ctx, const baorec_params* p, int algorithm, int n_catalogs, float* const* h_x,
                          float* const* h_y, float* const* h_z, const float* const* h_w, const int64_t* n, int field,
                          int shifts_only, float* const* h_ox, float* const* h_oy, float* const* h_oz) {
  BR_NEED_PLAN(ctx);
  BR_REQUIRE(n_catalogs >= 1 && h_x && h_y && h_z && h_w && n && h_ox && h_oy && h_oz, "batch arrays");
  struct ArraySource : BatchSource {
    int count;
    float* const *x, *const *y, *const *z, *const *ox, *const *oy, *const *oz;
    const float* const* w;
    const int64_t* n;
    int64_t nmax = 0;
    int acquire(int i, BatchItem* it) override {
      it->x = x[i], it->y = y[i], it->z = z[i], it->w = w[i], it->n = n[i];
      it->ox = ox[i], it->oy = oy[i], it->oz = oz[i];
      return BAOREC_OK;
    }
    int release(int) override { return BAOREC_OK; }
    int64_t max_rows() const override { return nmax; }
  } src;
  src.count = n_catalogs;
  src.x = h_x, src.y = h_y, src.z = h_z, src.w = h_w, src.n = n, src.ox = h_ox, src.oy = h_oy, src.oz = h_oz;
  for (int i = 0; i < n_catalogs; i++) {
    BR_REQUIRE(n[i] > 0 && h_x[i] && h_y[i] && h_z[i] && h_w[i] && h_ox[i] && h_oy[i] && h_oz[i], "catalog arrays");
    if (n[i] > src.nmax) src.nmax = n[i];
  }
  return batch_pipeline(ctx, p, algorithm, n_catalogs, field, shifts_only, src);
}

// The same pipeline fed from catalog FILES (text or NPY by extension) and writing NPY files: a reader thread parses
// catalog i+1 (and further ahead, as host buffer sets come free) into pinned memory and a writer thread stores the
// results of catalog i-1 while the device reconstructs catalog i -- examples/simulation.jl:12-40 looped over mocks.
int baorec_batch_files_f32(baorec_ctx* ctx, const baorec_params* p, int algorithm, int n_catalogs,
                           const char* const* in_paths, char delim, const int cols[4], int field, int shifts_only,
                           const char* const* out_paths, int n_threads, int64_t* n_rows, double* seconds) {
  BR_NEED_PLAN(ctx);
  BR_TRY(check_params(p));
  BR_REQUIRE(n_catalogs >= 1 && in_paths && cols, "n_catalogs / in_paths / cols");
  BR_REQUIRE(cols[0] >= 0 && cols[1] >= 0 && cols[2] >= 0, "columns of x, y, z");
  for (int i = 0; i < n_catalogs; i++) BR_REQUIRE(in_paths[i] != nullptr, "in_paths[i]");
  FileBatchConfig cfg;
  cfg.n_catalogs = n_catalogs;
  cfg.in_paths = in_paths;
  cfg.out_paths = out_paths;
  cfg.delim = delim;
  for (int c = 0; c < 4; c++) cfg.cols[c] = cols[c];
  cfg.n_threads = n_threads;
  cfg.n_slots = ctx->opt_batch_slots;
  HostAllocator al;
  al.alloc = [](void** out, size_t bytes) -> int { return cudaHostAlloc(out, bytes, cudaHostAllocDefault) == cudaSuccess ? 0 : 1; };
  al.free = [](void* q) { cudaFreeHost(q); };
  const auto t0 = std::chrono::steady_clock::now();
  int status;
  {
    FileBatchSource src(cfg, al, &ctx->batch_host_slots);  // the pinned buffer sets outlive the call
    status = batch_pipeline(ctx, p, algorithm, n_catalogs, field, shifts_only, src);
    if (status == BAOREC_OK) status = src.finish();
    else {
      const std::string msg = baorec_last_error();  // the abort below may record an I/O message of its own
      src.finish(true);
      set_error("%s", msg.c_str());
    }
    if (n_rows)
      for (int i = 0; i < n_catalogs; i++) n_rows[i] = src.rows()[i];
    if (seconds) {
      seconds[0] = src.read_seconds();
      seconds[1] = src.write_seconds();
      seconds[2] = src.wait_seconds();
    }
  }
  if (seconds) seconds[3] = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  return status;
}

}  // extern "C"
