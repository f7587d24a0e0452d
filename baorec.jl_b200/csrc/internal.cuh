// Internal definitions shared by the translation units of libbaorec_b200.so.
#pragma once
#include <cuda_runtime.h>
#include <cufft.h>
#include <stdint.h>
#include <stdio.h>
#include <string>
#include <vector>

#include "baorec_b200.h"
#include "host_slot.cuh"

namespace baorec {

void set_error(const char* fmt, ...);

#define BR_CUDA(expr)                                                                      \
  do {                                                                                     \
    cudaError_t _e = (expr);                                                               \
    if (_e != cudaSuccess) {                                                               \
      baorec::set_error("%s:%d CUDA error %s in %s", __FILE__, __LINE__,                   \
                        cudaGetErrorString(_e), #expr);                                    \
      return BAOREC_ERR_CUDA;                                                              \
    }                                                                                      \
  } while (0)

#define BR_CUFFT(expr)                                                                     \
  do {                                                                                     \
    cufftResult _e = (expr);                                                               \
    if (_e != CUFFT_SUCCESS) {                                                             \
      baorec::set_error("%s:%d cuFFT error %d in %s", __FILE__, __LINE__, (int)_e, #expr); \
      return BAOREC_ERR_CUFFT;                                                             \
    }                                                                                      \
  } while (0)

#define BR_TRY(expr)             \
  do {                           \
    int _s = (expr);             \
    if (_s != BAOREC_OK) return _s; \
  } while (0)

#define BR_REQUIRE(cond, msg)                         \
  do {                                                \
    if (!(cond)) {                                    \
      baorec::set_error("invalid argument: %s", msg); \
      return BAOREC_ERR_INVALID;                      \
    }                                                 \
  } while (0)

// Count + check a kernel launch; when profiling is on, bracket it with CUDA events.
#define BR_LAUNCH(ctx, kernel, grid, block, smem, stream, ...) \
  BR_LAUNCH_NAMED(ctx, #kernel, kernel, grid, block, smem, stream, __VA_ARGS__)
#define BR_LAUNCH_NAMED(ctx, name, kernel, grid, block, smem, stream, ...) \
  do {                                                                    \
    int _pi = baorec::prof_begin((ctx), (name), (stream));                \
    kernel<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__);           \
    baorec::prof_end((ctx), _pi, (stream));                               \
    (ctx)->n_kernels++;                                                   \
    BR_CUDA(cudaGetLastError());                                          \
  } while (0)

enum BufId {
  BUF_CK0 = 0,  // complex half-mesh (R2C output / k-space work)
  BUF_CK1,
  BUF_CK2,
  BUF_CKCACHE,  // delta_k of the cached result mesh (host pipeline)
  BUF_RS,       // real mesh: delta_s / multigrid right-hand side
  BUF_RX,       // real mesh: C2R output / displacement x
  BUF_RY,
  BUF_RZ,
  BUF_RAN,      // real mesh: randoms density
  BUF_WORK,     // cuFFT work area (shared by all plans)
  BUF_CACHE,    // result mesh of the host pipeline (recon.result_cache)
  BUF_PART,     // particle staging for host pipelines
  BUF_PART2,    // randoms staging
  BUF_OUT,      // per-particle outputs for host pipelines
  BUF_BINKEY,   // particle binning
  BUF_BINIDX,
  BUF_BINTMP,
  BUF_BININV,   // inverse permutation of the tile sort
  BUF_BINOUT,   // gather results in sorted order
  BUF_MG,       // multigrid level hierarchy (one slab allocation)
  BUF_A2A_SEND, // distributed FFT staging
  BUF_A2A_RECV,
  BUF_A2A_RECV2,
  BUF_A2A_RECV3,
  BUF_HALO_INBOX,  // halo / ghost planes the ring neighbours store into (mapped by all peers)
  BUF_HALO,
  BUF_MGD,      // multigrid level hierarchy of the slab-decomposed solver
  BUF_DET,      // 64-bit fixed-point accumulator of the deterministic scatter
  BUF_PK,       // multipole estimator: window tables + per-bin Float64 sums
  BUF_SHARD_SEND,   // particle sharding, data catalog: columns grouped by destination rank (then: results coming back)
  BUF_SHARD_RECV,   //   this rank's slab particles (x, y, z, w columns)
  BUF_SHARD_MAP,    //   position of every original particle in the send order
  BUF_SHARD_SEND2,  // the same three for the randoms catalog
  BUF_SHARD_RECV2,
  BUF_SHARD_MAP2,
  BUF_SHARD_OUT,    // per-particle results in slab order before they travel back
  BUF_COUNT
};

struct Buf {
  void* p = nullptr;
  size_t bytes = 0;
};

struct ProfRec {
  const char* name;
  cudaEvent_t a, b;
};

struct MgLevel {
  int nx = 0, ny = 0, nz = 0;
  size_t cells = 0;
  float* f = nullptr;   // right-hand side
  float* va = nullptr;  // solution ping
  float* vb = nullptr;  // solution pong / residual scratch
  float* px = nullptr;  // radial tables x_centre/cell per axis (length nx, ny, nz)
  float* py = nullptr;
  float* pz = nullptr;
  float cell[3];
  // slab-decomposed hierarchy (multi-GPU): slab != 0 -> this rank stores nzl real planes at plane
  // index 1..nzl of an (nzl+2)-plane buffer (planes 0 and nzl+1 are halos); slab == 0 -> the level
  // is replicated on every rank.  tmp: slab-layout window of the first replicated level.
  int nzl = 0, slab = 0;
  float* tmp = nullptr;
};

}  // namespace baorec

struct baorec_ctx {
  int device = 0;
  bool planned = false;
  int nx = 0, ny = 0, nz = 0, xh = 0;
  size_t M = 0, Mc = 0;
  float L[3] = {0, 0, 0}, mn[3] = {0, 0, 0};
  float cell[3] = {0, 0, 0};  // T(L/n), the gather's cell_size (src/mas.jl:221)
  cufftHandle r2c = 0, c2r = 0;
  bool have_plans = false;
  // own FFT path: batched 1-D cuFFT along x + our column kernels along y and z (fft.cu)
  cufftHandle px_r2c = 0, px_c2r = 0;
  bool have_x_plans = false;
  float2* d_tw[2] = {nullptr, nullptr};  // twiddle tables for the y and z axes
  // split FFT path (option "fft_split_planes" = C > 0): batched 2-D transforms over chunks of C planes
  // (the intermediate of cuFFT's x and y passes stays L2-resident) + one strided 1-D pass along z
  cufftHandle ps_r2c = 0, ps_c2r = 0, ps_z = 0;
  int split_planes_planned = 0;
  int opt_fft_split = 0;
  int opt_fft_tile_cols = -1;  // columns per tile of the own FFT kernels at N = 1024: -1 = auto (z passes 16, y passes 8), 4 / 8 / 16 = everywhere
  int opt_fft_prefetch = -1;   // L2 prefetch distance of the own FFT kernels in CTAs: -1 = auto (y passes 148, z passes none), 0 = none, > 0 = every pass
  int opt_own_fft = -1;  // column FFT kernels of fft.cu instead of cuFFT's 3-D plans: -1 = auto (ny, nz powers of two >= 512), 0 = never, 1 = wherever supported (>= 256)
  size_t work_bytes = 0;
  float* d_k[3] = {nullptr, nullptr, nullptr};  // k tables (xh, ny, nz)
  float* d_xv[3] = {nullptr, nullptr, nullptr}; // cell-centre tables (nx, ny, nz)
  // per-axis Gaussian factors exp(-0.5 R^2 k_a^2) in Float64 (xh + ny + nz doubles) for radius gauss_R
  double* d_gauss = nullptr;
  float gauss_R = 0.f;
  bool gauss_valid = false;
  baorec::Buf bufs[baorec::BUF_COUNT];
  unsigned long long* d_oob = nullptr;  // out-of-box particle counter
  double* d_scal = nullptr;             // small device scalars (DC modes, sums)
  float* d_minmax = nullptr;            // 6 floats for setup_box
  cudaStream_t own_stream = nullptr;    // used by host pipelines
  cudaStream_t copy_stream = nullptr;   // H2D uploads that overlap compute
  // tile sort of the read-back catalog, forked onto side_stream so that it overlaps the
  // displacement transforms (gather_prebin / gather3 in mas.cu)
  cudaStream_t side_stream = nullptr;
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
  bool prebin_valid = false;
  const float *prebin_x = nullptr, *prebin_y = nullptr, *prebin_z = nullptr;
  int64_t prebin_n = 0;
  int prebin_mas = 0, prebin_slab_mode = 0;
  void* prebin_rec = nullptr;
  const unsigned* prebin_nvalid = nullptr;
  const unsigned* prebin_starts = nullptr;
  unsigned prebin_ntiles = 0;
  unsigned* prebin_inv = nullptr;
  int opt_overlap_sort = 0;  // measured: no gain on one GPU (the transforms already saturate HBM), off by default
  cudaEvent_t ev_copy = nullptr;
  cudaEvent_t ev[8] = {};
  float stage_ms[8] = {};
  int n_stage = 0;
  int64_t n_kernels = 0, n_fft = 0;
  bool cache_valid = false;
  bool kcache_valid = false;   // BUF_CKCACHE holds the unnormalised R2C of the cached result mesh
  bool want_kcache = false;    // set by the host pipeline around the solve
  const float* kcache_mesh = nullptr;  // the result mesh delta_k belongs to
  int opt_keep_delta_k = 1;    // device API: keep delta_k of the last reconstructed_overdensity! result
  int64_t opt_bin_min_particles = 1 << 18;  // catalogs at least this large are z-binned first
  int opt_fuse_kspace = 1;     // fixed-LOS iterations folded into one k-space pass
  std::vector<baorec::HostSlot> batch_host_slots;  // pinned buffer sets of baorec_batch_files_f32, kept between calls
  int opt_batch_slots = 4;     // baorec_batch_files_f32: host buffer sets (>= 3: reading ahead | uploading + in flight | being written)
  int opt_scatter_tiles = 0;   // scatter: (z, y/8, x/128) tile order instead of z slabs
  int opt_det_scatter = 0;     // CIC scatter through 64-bit fixed-point integer reductions: bit-reproducible meshes
  // Unified sort: run! sorts the catalog ONCE into the gather's tile order (records x,y,z,w + inverse
  // permutation + a 64-bit content hash of the wrapped positions); the scatter deposits in that order
  // and a read-back of the same, unmodified position arrays (same pointers, same count, same hash --
  // recomputed from the arrays on every read) reuses it instead of sorting again.
  int opt_unified_sort = 1;
  bool sortc_valid = false;
  const float *sortc_x = nullptr, *sortc_y = nullptr, *sortc_z = nullptr;
  int64_t sortc_n = 0;
  unsigned sortc_ntiles = 0;
  int sortc_slab_mode = 0;  // 0: whole-mesh tiles; 2: tiles of the slab's gather layout (the sort was made by a slab scatter)
  unsigned long long* d_hash = nullptr;  // [0] hash at sort time, [1] hash at read time, [2] match flag
  int64_t n_sort_reuse = 0;              // read-backs that reused the run!'s sort (diagnostics)
  int opt_gather_tiles = 1;    // gather: fine (z, y/8, x/128) tile binning instead of z slabs
  int opt_scatter_pairs = 2;   // vector reductions: 1 = red.global.add.v2.f32 for aligned x pairs (6 per particle), 2 = one red.global.add.v4.f32 per CIC row (5 per particle); measured at C4: scatter 3.89 -> 3.51 -> 2.99 ms, default 2
  int opt_gather_stage = 2;    // tile gather staging: 2 = tensor-map TMA (gather_tile_tma_kernel: six cp.async.bulk.tensor.3d copies per tile; C4: 3.48 -> 2.90 ms = 0.85 of HBM), 1 = strength-reduced cp.async rows (5.28 -> 3.48 ms), 0 = the round-1 loop; bit-identical shifts
  int opt_zg_scatter = 0, opt_zg_gather = 0;  // z planes per bin (0 = auto)
  int64_t last_wrapped = 0;  // particles whose position cic! wrapped in the last scatter
  // per-launch profiling (baorec_profile_*)
  bool prof_on = false;
  std::vector<baorec::ProfRec> prof;      // records of the current window
  std::vector<cudaEvent_t> prof_pool;     // recycled events
  // multigrid
  std::vector<baorec::MgLevel> levels;
  std::vector<baorec::MgLevel> dlevels;  // slab-decomposed hierarchy (baorec_plan_dist)
  int opt_mg_fd_gradient = 0;  // MultigridRecon read-back: 1 = finite-difference gradient of phi in ONE gather (no transforms) instead of the reference's spectral gradient
  int opt_mg_remove_mean = 1;  // multigrid solves on f - mean(f) (cubic cells): see mg_fmg in multigrid.cu
  int opt_mg_coarse = 1;  // levels of <= 4096 cells: the bottom of the V-cycle in one thread block
  int opt_mg_bulk = 0;    // staged kernel: row bodies by TMA bulk copies; measured SLOWER than cp.async staging (554 vs 355 us), off
  int opt_mg_ring = 6;    // planes in the staged kernel's shared-memory ring (3 or 6)
  int opt_mg_kernel = 0;  // 0: staged shared-memory kernel where it applies; 1: register march; 2: generic
  int64_t opt_mg_slab_min_cells = 1 << 22;  // levels with fewer cells (128^3 and below) are replicated, not slab-decomposed
  bool mg_radial_tables = false;
  // distributed
  void* comm = nullptr;  // ncclComm_t
  int rank = 0, nranks = 1;
  bool dist = false;
  int nz_loc = 0, z0 = 0, ny_loc = 0, y0 = 0;
  // peer-to-peer transposes: recv buffers of every rank mapped through CUDA IPC (double-buffered)
  // peer-copy exchange (dist.cu): [0] = K-layout receive buffer (forward), [1] = plane-layout receive buffer
  // (inverse) of every rank, and every rank's flag block, mapped through CUDA IPC
  bool p2p = false;
  int opt_dist_exchange = 1;  // 1 = peer copies + flags (default), 0 = pack / transpose kernels + NCCL all-to-all
  float2* peer_recv[3][16] = {};   // [0] k layout (forward), [1], [2] plane layout (inverse, two slots)
  float2* own_recv[3] = {nullptr, nullptr, nullptr};
  void* peer_flags[16] = {};
  void* d_flags = nullptr;
  unsigned seq_k = 0, seq_a[2] = {0, 0};
  // halo / ghost planes through peer memory (halo_send_kernel / halo_recv_kernel in dist.cu)
  void* peer_inbox[16] = {};
  float* own_inbox = nullptr;
  size_t inbox_slot_floats = 0;
  unsigned seq_halo = 0;
  unsigned* d_halo_done = nullptr;
  int opt_peer_halo = 1;  // 0 = grouped ncclSend / ncclRecv for the ring exchanges  // forward / inverse transforms issued since the flags were reset (same on every rank)
  // particle sharding by slab (baorec_shard_catalog_f32): routing tables of the last call per slot (0 data, 1 randoms)
  struct ShardState {
    bool valid = false;
    int64_t n = 0, n_local = 0;            // particles passed in / particles of this rank's slab
    int64_t send_cnt[16] = {}, send_off[16] = {};   // per destination rank, in the send order
    int64_t recv_cnt[16] = {}, recv_off[16] = {};   // per source rank, in the slab order
    int64_t off_at_dst[16] = {}, off_at_src[16] = {};  // my segment's offset inside rank r's receive columns / send order
    int64_t send_stride = 0, recv_stride = 0;       // column strides of the two buffers
    bool peer = false;                              // routed with peer copies (else NCCL)
  } shard[2];
  // peer-copy sharding: every rank's receive columns [0] and send staging [1] per slot, mapped through CUDA IPC, with the
  // capacities (particles per column) all ranks agree on
  void* shard_peer[2][2][16] = {};
  int64_t shard_cap[2][2][16] = {};
  unsigned seq_shard[2][2] = {};   // [slot][0 shard / 1 unshard]
  void* d_ipc_stage = nullptr;
  unsigned long long* d_shard_cnt = nullptr;  // [0..15] per-owner counts, [16] out-of-box, [17..33] cursors; [40..] gathered
  // displacement meshes (BUF_RX/RY/RZ) of the cached result, kept across read_shifts /
  // reconstructed_positions calls (the examples read data, randoms-sym and randoms-iso back from
  // one reconstruction: examples/simulation.jl:32-35) so that only the first call pays for the transforms
  bool disp_valid = false;
  int disp_algo = 0;
  const float* disp_mesh = nullptr;
  const float* mg_result_mesh = nullptr;  // phi produced by the last reconstructed_potential! on this context
  bool kcache_potential = false;  // BUF_CKCACHE holds phi_k (MultigridRecon) instead of delta_k
  int slab_mode = 0;  // 0: whole mesh; 1: scatter into a slab (+1 ghost plane); 2: gather from a slab (+3 halo planes)
  cufftHandle p2d_r2c = 0, p2d_c2r = 0, p1d = 0, p1d_s = 0;  // p1d_s: strided z transform of the K[z][yl][x] layout
  bool have_dist_plans = false;
  // pipelined slab transposes: the local planes are processed in opt_a2a_chunks chunks so that the
  // all-to-all of chunk c (comm_stream) overlaps the 2-D transforms / pack / unpack of its neighbours
  cufftHandle pc_r2c = 0, pc_c2r = 0;
  int chunk_planes = 0;  // planes per chunk the chunk plans were made for (0 = none)
  int opt_a2a_chunks = 4;
  cudaStream_t comm_stream = nullptr;
  cudaStream_t comm_stream2 = nullptr;  // the local 1/P block of a peer-copy exchange (an HBM copy) runs beside the NVLink copies
  cudaStream_t comm_stream4 = nullptr;
  cudaStream_t comm_stream3 = nullptr;  // every other remote copy (option "comm_split"): two copy engines, two peers at a time
  int opt_comm_split = 1;
  int opt_push_sm = -1;  // remote blocks of the peer exchange pushed by an SM kernel instead of the copy engines: -1 = with 5+ ranks, 0 never, 1 always
  cudaEvent_t ev_chunk[8] = {}, ev_a2a[8] = {}, ev_local[8] = {}, ev_split[8] = {}, ev_split2[8] = {};
  // catalog pre/post-processing (catalog.cu): comoving-distance table r(z) on uniform z knots
  double* d_cosmo_r = nullptr;
  std::vector<double> h_cosmo_r;
  int64_t cosmo_n = 0;
  double cosmo_z0 = 0.0, cosmo_z1 = 0.0, cosmo_dz = 0.0;
  double* d_cosmo_g = nullptr;   // guess table of the inverse interpolation (catalog_math.cuh)
  double cosmo_inv_h = 0.0;
  int cosmo_corr = 1;            // +-1 correction steps after the guess, measured when the table is built
  int opt_catalog_corr = -1;     // test hook: >= 0 overrides it (0 forces the bisection path for every miss)
};

namespace baorec {

int prof_begin(baorec_ctx* ctx, const char* name, cudaStream_t st);
void prof_end(baorec_ctx* ctx, int idx, cudaStream_t st);
int need(baorec_ctx* ctx, BufId id, size_t bytes, void** out);
template <class T>
inline int need_t(baorec_ctx* ctx, BufId id, size_t count, T** out) {
  void* p = nullptr;
  int s = need(ctx, id, count * sizeof(T), &p);
  *out = (T*)p;
  return s;
}
void release(baorec_ctx* ctx, BufId id);
int check_oob(baorec_ctx* ctx, cudaStream_t st, const char* what);
int reset_oob(baorec_ctx* ctx, cudaStream_t st);

// FFT wrappers (count executions, bind stream).
int fft_r2c(baorec_ctx* ctx, const float* in, float2* out, cudaStream_t st);
int fft_c2r(baorec_ctx* ctx, float2* in, float* out, cudaStream_t st);

// mas.cu
int scatter(baorec_ctx* ctx, float* rho, float* x, float* y, float* z, const float* w, int64_t n, int wrap,
            int mas, cudaStream_t st);
// Starts the tile sort of a read-back catalog on the side stream (ordered after everything already
// queued on `st`); the next gather3 on the same arrays joins it instead of sorting.  A no-op for
// catalogs / modes that do not use the tile gather.
int gather_prebin(baorec_ctx* ctx, const float* x, const float* y, const float* z, int64_t n, int mas,
                  cudaStream_t st);
// mode: 0 disp, 1 rsd, 2 sum ; positions: write pos - shift
int gather3(baorec_ctx* ctx, const float* fx, const float* fy, const float* fz, const float* x, const float* y,
            const float* z, int64_t n, float* ox, float* oy, float* oz, int mas, int field, float f, int has_los,
            const float* los, int positions, cudaStream_t st);

// grad(phi) at the particles by finite differences (the reference's commented-out read_grad_cic!, src/mas.jl:388-466)
// + the read_shifts epilogue; single GPU
int gather_fd(baorec_ctx* ctx, const float* phi, const float* x, const float* y, const float* z, int64_t n, float* ox,
              float* oy, float* oz, int field, float f, int has_los, const float* los, int positions, cudaStream_t st);

// kspace.cu
int smooth(baorec_ctx* ctx, float* mesh, float R, cudaStream_t st);
int setup_overdensity(baorec_ctx* ctx, const baorec_params* p, float* mesh, float* x, float* y, float* z,
                      const float* w, int64_t n, float* rx, float* ry, float* rz, const float* rw, int64_t nr,
                      int wrap, cudaStream_t st);
int iterate(baorec_ctx* ctx, float* delta_r, const float* delta_s, int iter, float beta, const float* los,
            cudaStream_t st);
int displacement_meshes(baorec_ctx* ctx, const float* mesh, int algorithm, float* px, float* py, float* pz,
                        cudaStream_t st, bool use_kcache = false);

int setup_overdensity_into(baorec_ctx* ctx, const baorec_params* p, float* mesh, float* delta_out, float* x, float* y,
                           float* z, const float* w, int64_t n, float* rx, float* ry, float* rz, const float* rw,
                           int64_t nr, int wrap, cudaStream_t st);
int reconstructed_overdensity(baorec_ctx* ctx, const baorec_params* p, float* mesh, float* x, float* y, float* z,
                              const float* w, int64_t n, float* rx, float* ry, float* rz, const float* rw, int64_t nr,
                              cudaStream_t st);
int setup_box_dev(baorec_ctx* ctx, const float* x, const float* y, const float* z, int64_t n, float pad,
                  float L_out[3], float mn_out[3], cudaStream_t st);
// ctx.cu
// device tables of the separable Gaussian for smoothing radius R (rebuilt when R or the box changes)
int gauss_tables(baorec_ctx* ctx, float R, const double** gx, const double** gy, const double** gz, cudaStream_t st);
int plan_common(baorec_ctx* ctx, int nx, int ny, int nz, const float L[3], const float mn[3]);
void host_xvec(int n, float L, float mn, std::vector<float>& out);

int stash_dc(baorec_ctx* ctx, const float2* ck, int slot, double mul, cudaStream_t st);
int kpass_fused_T(baorec_ctx* ctx, const float2* in, float2* out_c2r, float2* keep, const baorec_params* p,
                  cudaStream_t st);
int kpass_fused_delta_T(baorec_ctx* ctx, const float2* in, float2* out_c2r, float2* keep, const baorec_params* p,
                        cudaStream_t st);
int kpass_disp_T(baorec_ctx* ctx, const float2* in, float2* out, int comp, bool potential, cudaStream_t st);
int kpass_disp_gh(baorec_ctx* ctx, const float2* in, float2* g, float2* h, bool potential, cudaStream_t st);
int kpass_setup_box_T(baorec_ctx* ctx, const float2* in, float2* out, const baorec_params* p, cudaStream_t st);
int kpass_gauss_T(baorec_ctx* ctx, const float2* in, float2* out, float R, cudaStream_t st);
int kpass_iter_pair_T(baorec_ctx* ctx, const float2* in, float2* out, int i, int j, cudaStream_t st);
// real-space passes on a slab of `nzl` planes starting at global plane z_lo
int radial_update_slab(baorec_ctx* ctx, float* out, const float* src, const float* X, int nzl, int z_lo, int ci,
                       int cj, float fac, cudaStream_t st);
// n_ran < 0: the randoms count is read from d_scal[2]
int randoms_combine(baorec_ctx* ctx, float* out, const float* dat, const float* ran, float bias, float ran_min,
                    double n_ran, size_t cells, cudaStream_t st);

// fft.cu
bool own_fft_available(const baorec_ctx* ctx);
bool own_fft_fused_available(const baorec_ctx* ctx);
bool own_slab_z_available(const baorec_ctx* ctx);
int own_slab_z(baorec_ctx* ctx, const float2* in, float2* out, int dir, cudaStream_t st);
int own_fft_setup(baorec_ctx* ctx);
int own_r2c(baorec_ctx* ctx, const float* in, float2* out, cudaStream_t st);
int own_c2r(baorec_ctx* ctx, float2* in, float* out, cudaStream_t st);
int own_fused_los_solve(baorec_ctx* ctx, const baorec_params* p, float* mesh, float2* work, float2* keep,
                        cudaStream_t st);
int own_displacements(baorec_ctx* ctx, const float* mesh, const float2* from_k, int algorithm, float2* w0, float2* w1,
                      float2* w2, float* px, float* py, float* pz, cudaStream_t st);

// multigrid.cu
int mg_setup_levels(baorec_ctx* ctx);
int reconstructed_potential(baorec_ctx* ctx, const baorec_params* p, float* phi, float* x, float* y, float* z,
                            const float* w, int64_t n, float* rx, float* ry, float* rz, const float* rw, int64_t nr,
                            cudaStream_t st);
int mg_fmg(baorec_ctx* ctx, const float* f, float* v, float beta, float damping, int n_jacobi, int n_vcycle,
           const float* los, cudaStream_t st);
// Slab-decomposed fmg: f_slab is an (nz_loc+2)-plane buffer with this rank's right-hand side at
// planes 1..nz_loc; *result is the library buffer (same layout) holding the solution.
int mg_fmg_dist(baorec_ctx* ctx, float* f_slab, float** result, float beta, float damping, int n_jacobi,
                int n_vcycle, const float* los, cudaStream_t st);
// dist.cu: both halo planes of a slab-layout buffer (ring neighbours, periodic); all-gather of slabs
int mg_halo_exchange(baorec_ctx* ctx, float* buf, size_t plane, int nzl, cudaStream_t st);
int mg_allgather(baorec_ctx* ctx, const float* send, float* recv, size_t count, cudaStream_t st);
int mg_allreduce_sum(baorec_ctx* ctx, double* d_value, cudaStream_t st);  // one double, summed over the ranks in place
void dist_refresh_mode(baorec_ctx* ctx);  // re-evaluates ctx->p2p (peer-copy exchange usable?) after an option / mapping change

#define BR_NEED_PLAN(ctx)                                        \
  do {                                                           \
    BR_REQUIRE((ctx) != nullptr, "ctx is NULL");                 \
    if (!(ctx)->planned) {                                       \
      baorec::set_error("call baorec_plan before using the context"); \
      return BAOREC_ERR_NOT_PLANNED;                             \
    }                                                            \
    BR_CUDA(cudaSetDevice((ctx)->device));                       \
  } while (0)

inline unsigned cdiv(size_t a, size_t b) { return (unsigned)((a + b - 1) / b); }

}  // namespace baorec
