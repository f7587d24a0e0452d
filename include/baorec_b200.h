/* baorec_b200.h -- C ABI of libbaorec_b200.so, the B200 (sm_100a) engine behind
 * BAOrec.jl's device hot path.
 *
 * The reference has no FFI today: its device boundary is Julia multiple
 * dispatch on CuArray (cited per entry point below, paths relative to the
 * reference checkout).  Each function here is what a `ccall` from the Julia
 * shim (baorec.jl_b200/julia/BAOrecB200.jl) or the Python ctypes mirror
 * (baorec.jl_b200/host.py) binds instead of that CuArray method.
 *
 * Conventions
 *  - Every function returns 0 on success, a negative BAOREC_ERR_* otherwise;
 *    the message is available from baorec_last_error() (thread-local).
 *    No exception or longjmp ever crosses the boundary.
 *  - Meshes are float32, C-order [iz][iy][ix] with x fastest == Julia
 *    Array{Float32,3}(nx,ny,nz) memory.  Particle arrays are SoA float32.
 *  - Pointers named d_* are device pointers owned by the caller; the library
 *    never frees or retains them past the call.  Pointers named h_* are host
 *    pointers (pinned memory makes the copies asynchronous, pageable works).
 *  - `stream` is a cudaStream_t passed as void* (CUDA.stream().handle in
 *    Julia, torch.cuda.current_stream().cuda_stream in Python); 0 = legacy
 *    default stream.  Primitives enqueue work on it and return; those that
 *    can detect out-of-box particles synchronise the stream before returning
 *    so the error code is reliable (documented per function).
 *  - A context is bound to one device, is not re-entrant, and owns all
 *    scratch (FFT plans and work area, k/x tables, multigrid level buffers,
 *    binning buffers, NCCL staging): no per-call cudaMalloc after warm-up.
 */
#ifndef BAOREC_B200_H
#define BAOREC_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BAOREC_VERSION 100

enum {
  BAOREC_OK = 0,
  BAOREC_ERR_INVALID = -1,     /* bad argument */
  BAOREC_ERR_CUDA = -2,        /* CUDA runtime failure */
  BAOREC_ERR_CUFFT = -3,       /* cuFFT failure */
  BAOREC_ERR_NCCL = -4,        /* NCCL failure */
  BAOREC_ERR_OUT_OF_BOX = -5,  /* particle outside the mesh (reference: BoundsError /
                                  out-of-bounds atomic, src/mas.jl:33-49, 82-97) */
  BAOREC_ERR_NOT_PLANNED = -6, /* baorec_plan has not been called */
  BAOREC_ERR_NOMEM = -7,
  BAOREC_ERR_OUT_OF_RANGE = -8, /* redshift / distance outside the cosmology tables (reference:
                                  Interpolations.jl BoundsError) */
  BAOREC_ERR_IO = -9            /* a catalog file cannot be opened, read, parsed or written */
};

enum { BAOREC_MAS_CIC = 0, BAOREC_MAS_TSC = 1, BAOREC_MAS_PCS = 2 };  /* TSC and PCS are extensions */
enum { BAOREC_FIELD_DISP = 0, BAOREC_FIELD_RSD = 1, BAOREC_FIELD_SUM = 2 }; /* :disp :rsd :sum */
enum { BAOREC_ITERATIVE = 0, BAOREC_MULTIGRID = 1 };

typedef struct baorec_ctx baorec_ctx;
typedef void* baorec_stream;

/* Mirrors the keyword fields of IterativeRecon / MultigridRecon
 * (src/recon.jl:2-16, 18-33) plus the constants the reference hard-codes. */
typedef struct baorec_params {
  float bias;
  float f;
  float smoothing_radius;
  float beta;                  /* f / bias unless overridden */
  int32_t n_iter;              /* IterativeRecon.n_iter (default 3) */
  int32_t has_los;             /* 0: los = nothing (radial, observer at origin) */
  float los[3];
  float jacobi_damping_factor; /* default 0.4 */
  int32_t jacobi_niterations;  /* default 5 */
  int32_t vcycle_niterations;  /* default 6 */
  int32_t mas;                 /* BAOREC_MAS_* ; reference = CIC */
  float ran_min;               /* 0.01, src/recon.jl:70 */
  float box_pad;               /* 500,  src/recon.jl:172,253 */
} baorec_params;

/* ---- lifecycle ------------------------------------------------------------ */
int baorec_version(void);
const char* baorec_last_error(void);

/* Create a context on CUDA device `device` (runtime API, primary context:
 * shares the context CUDA.jl / torch already use). */
int baorec_create(int device, baorec_ctx** out);
int baorec_destroy(baorec_ctx* ctx);

/* Replaces setup_fft! (src/recon.jl:35-37) and the per-call k_vec/x_vec uploads
 * (src/iterative.jl:153,168; src/utils.jl:89): builds cuFFT R2C/C2R plans, the
 * k and x tables and (lazily) all scratch for an (nx,ny,nz) mesh in the given box. */
int baorec_plan(baorec_ctx* ctx, int nx, int ny, int nz, const float box_size[3], const float box_min[3]);
/* Change the box without re-planning the FFTs (run! with randoms overrides it,
 * src/recon.jl:172). */
int baorec_set_box(baorec_ctx* ctx, const float box_size[3], const float box_min[3]);
/* Tuning knobs (all default to the fast path):
 *   "bin_min_particles" (default 262144): catalogs at least this large are counting-sorted by
 *       z slab before the scatter / gather so the mesh window stays L2-resident; 0 = always.
 *   "fuse_kspace" (default 1): with a fixed line of sight reconstructed_overdensity! folds the
 *       smoothing, normalisation and all n_iter iterations into ONE k-space pass between one
 *       R2C and one C2R (iterate! is linear and diagonal in k for a constant LOS); 0 = run the
 *       reference's sequence of iterate! calls (2 + 2 n_iter transforms).
 *   "scatter_pairs" (default 2): 0 = eight scalar red.global.add.f32 per particle; 1 = the tile-ordered CIC scatter deposits the two x-neighbours of a row with one
 *       vector reduction (red.global.add.v2.f32) when they form an aligned pair -- same cells, same values, 6 instead
 *       of 8 L2 reductions per particle on average; 2 = one red.global.add.v4.f32 per row on the aligned
 *       block that holds the pair (+0 in the unused slots) plus a scalar where the pair straddles two blocks (5 on
 *       average, branch-free).  Either value also switches the binned TSC scatter to one quad (+ one pair for half of
 *       the particles) per stencil row: 13.5 instead of 27.  Measured on B200 at 1024^3 / 1e8 particles: scatter 3.89 ms
 *       (0), 3.51 ms (1), 2.99 ms (2).
 *   "gather_stage" (default 2): how the tile gather stages the 2 x 9 x 129 window of each displacement mesh in shared
 *       memory.  2 = the TMA engine: one 3-D tensor map per field, six cp.async.bulk.tensor.3d copies per tile issued by
 *       one thread, one mbarrier (the periodic column / row the engine cannot fetch come from ordinary loads); needs a
 *       row pitch and base that are multiples of 16 bytes and a driver with cuTensorMapEncodeTiled, else 1 is used.
 *       1 = cp.async rows with one base pointer per (field, plane) and one multiply-add per row; 0 = the first kernel
 *       (twice the instructions of 1).  Bit-identical results.  Measured on B200 at 1024^3 / 1e8 particles: gather
 *       5.28 ms (0), 3.48 ms (1), 2.90 ms (2) = 0.85 of the measured HBM copy rate.
 *   "deterministic_scatter" (default 0): 1 = the CIC scatter (single GPU) accumulates 2^-40 fixed-point values with
 *       64-bit integer reductions and rounds to Float32 once: meshes, `ran > threshold` masks and everything after
 *       them are bit-reproducible from run to run and independent of the particle order (float reductions are not);
 *       costs an 8-byte-per-cell scratch mesh and one conversion pass.  With the option set, a TSC / PCS scatter or a
 *       scatter on slabs (multi-GPU) returns BAOREC_ERR_INVALID instead of falling back to float reductions.  A cell may
 *       accumulate up to 2^23 (8.4e6) in weight before the 2^-40 fixed-point sum wraps.
 *   "mg_remove_mean" (default 1): MultigridRecon solves on delta - mean(delta) when the cells are cubic.  A survey's
 *       delta has a non-zero mean; damped Jacobi on a periodic mesh then drifts by a constant that exact arithmetic
 *       does not feel (the operator's diagonal is uniform, src/multigrid.jl:82) but Float32 does: the reference's
 *       Float32 arithmetic ends 1e-2 from the reference's own Float64 run at 512^3.  1 = Float64-equivalent shifts
 *       (the potential comes without the drift constant); 0 = the reference's Float32 arithmetic, drift included.
 *   "mg_fd_gradient" (default 0): 1 = read_shifts / reconstructed_positions of a MultigridRecon (device API, CIC) take
 *       Psi = grad(phi) by finite differences in ONE gather over phi -- the read_grad_cic! the reference sketches and
 *       leaves commented out (src/mas.jl:388-466, src/multigrid.jl:759) -- instead of 1 R2C + 3 C2R + 3 gathers.
 *       Differs from the reference's spectral gradient by the finite-difference error, hence an option.
 *   "own_fft" (default -1): -1 = the column FFT kernels of csrc/fft.cu replace cuFFT's 3-D plans where they were
 *       measured faster (ny, nz powers of two >= 512); 0 = never; 1 = wherever supported (>= 256).
 *   "dist_exchange" (default 1), "push_sm" (default -1), "comm_split" (default 1): exchange of the slab transforms, see
 *       baorec_dist_ipc_export below.
 *   "mg_slab_min_cells" (default 4194304): slab-decomposed multigrid levels with fewer cells are
 *       replicated on every rank instead of exchanging halos. */
int baorec_set_option(baorec_ctx* ctx, const char* name, int64_t value);
/* Bytes of device scratch currently owned by the context. */
int64_t baorec_scratch_bytes(const baorec_ctx* ctx);
/* Read-backs that reused the tile sort of the preceding run! (option "unified_sort", default 1: the
 * catalog is sorted once per reconstruction when read_shifts / reconstructed_positions is given the
 * same, unmodified position arrays; validated by a 64-bit content hash on every read). */
int64_t baorec_sort_reuse_count(const baorec_ctx* ctx);
/* Kernel launches issued by this library since creation (cuFFT launches are
 * counted separately in *fft_execs). */
int baorec_launch_counts(const baorec_ctx* ctx, int64_t* kernels, int64_t* fft_execs);
/* Per-stage timing of the last host-pipeline calls (CUDA events on the context's stream), ms:
 * [0] H2D of the catalogs (+ setup_box), [1] set-up + solve, [2] D2H (wrapped positions, mesh);
 * after baorec_read_host_f32 also [3] H2D of positions + displacement meshes, [4] gather,
 * [5] D2H of the outputs.  Returns the count written (<= cap). */
int baorec_last_stage_ms(const baorec_ctx* ctx, float* out, int cap);
/* Per-launch profiling: while enabled every kernel launch and cuFFT execution issued by the
 * context is bracketed by CUDA events on its stream.  baorec_profile_enable(ctx, on) clears
 * the current window; baorec_profile_read synchronises the device and aggregates the window
 * by kernel name: names[i*name_stride ...] (NUL-terminated), total_ms[i], counts[i]; returns
 * the number of distinct names (<= cap). */
int baorec_profile_enable(baorec_ctx* ctx, int on);
int baorec_profile_read(baorec_ctx* ctx, char* names, int name_stride, float* total_ms, int32_t* counts, int cap);

/* ---- multi-GPU (one process per GPU; slabs along z) ------------------------ */
/* 128-byte NCCL unique id, created on rank 0 and broadcast by the host side
 * (torch.distributed / MPI.jl). */
int baorec_comm_unique_id(void* out128);
int baorec_comm_init(baorec_ctx* ctx, int rank, int nranks, const void* unique_id128);
/* Distributed plan: this rank owns z-planes [rank*nz/P, (rank+1)*nz/P) in real space and the
 * rows y in [rank*ny/P, (rank+1)*ny/P) in k space (nz, ny divisible by P).  With nranks == 1 no
 * communicator is needed (baorec_comm_init(ctx, 0, 1, NULL)). */
int baorec_plan_dist(baorec_ctx* ctx, int nx, int ny, int nz, const float box_size[3], const float box_min[3]);
int baorec_slab_range(const baorec_ctx* ctx, int* z_lo, int* nz_loc);
/* Peer-copy exchange of the slab transforms (option "dist_exchange" = 1, the default; after baorec_plan_dist): every
 * rank exports the CUDA IPC handles of its three receive buffers (k-layout for the forward, two plane-layout ones for
 * the inverse transforms, so that two fields can be in flight), of its halo inbox (ghost / halo planes of the ring
 * neighbours land there: two small kernels per exchange instead of a grouped ncclSend / ncclRecv; option "peer_halo")
 * and of its flag block (5 x 64 bytes; the call also resets the rank's flags, so it must precede the
 * gathering of the handles on every rank), the host side all-gathers them, and each rank opens the others'.  From then
 * on an exchange is one strided copy-engine copy per peer and plane chunk over NVLink, straight from the 2-D
 * transform's output into the peer's buffer, announced by a sequence flag stored into the peer's flag block: no
 * pack / transpose kernels, no staging buffer, no collective, no barrier.  With one rank the library maps itself.
 * "dist_exchange" = 0 selects round 1's scheme: pack + tile-transpose kernels and grouped ncclSend/ncclRecv. */
int baorec_dist_ipc_close(baorec_ctx* ctx);  /* drop the mappings (before a re-plan frees the buffers) */
int baorec_dist_ipc_export(baorec_ctx* ctx, void* out320);
int baorec_dist_ipc_open(baorec_ctx* ctx, const void* all_handles /* nranks x 320 bytes */, int nranks);
int baorec_dist_exchange_mode(const baorec_ctx* ctx); /* 1: peer copies (k space K[z][yl][x]); 0: NCCL (T[yl][x][z]) */
/* Owner rank of every particle = slab of its cic! base plane (src/mas.jl:15-30), -1 if out of box. */
int baorec_slab_owner_f32(baorec_ctx* ctx, const float* d_z, int64_t n, int32_t* d_owner, baorec_stream stream);
/* Particle sharding by slab ("one bucketed all-to-all(v)"; the reference's own attempt,
 * send_to_relevant_process src/mas.jl:112-140, is unfinished).  Every rank passes ANY part of the catalog (device
 * arrays, untouched); on return *d_sx .. *d_sw point to library-owned columns holding the *n_local particles whose
 * base plane lies in this rank's slab (valid until the next call with the same slot; run_dist's cic! write-back of
 * wrapped positions goes into them).  slot 0 = data, 1 = randoms: one routing table each is kept.  Device counting
 * sort by owner + all-gathered count matrix + one grouped ncclSend/ncclRecv per column and peer.  Out-of-box particles
 * on ANY rank make EVERY rank return BAOREC_ERR_OUT_OF_BOX.  Synchronises the stream (the counts are needed on the host). */
int baorec_shard_catalog_f32(baorec_ctx* ctx, int slot, const float* d_x, const float* d_y, const float* d_z,
                             const float* d_w, int64_t n, float** d_sx, float** d_sy, float** d_sz, float** d_sw,
                             int64_t* n_local, baorec_stream stream);
/* The way back: up to three columns in slab order (n_local values each; NULL = skip) -> the order of the arrays that
 * were passed to baorec_shard_catalog_f32 (n values each). */
int baorec_unshard_f32(baorec_ctx* ctx, int slot, const float* d_a, const float* d_b, const float* d_c, float* d_oa,
                       float* d_ob, float* d_oc, baorec_stream stream);
/* Slab-decomposed transforms (unnormalised): real slab [nz_loc][ny][nx] <-> this rank's y slab of k space (complex):
 * K[nz][ny_loc][nx/2+1] with the peer-copy exchange (2-D cuFFT over the local planes, peer copies, strided 1-D cuFFT
 * along z), T[ny_loc][nx/2+1][nz] with the NCCL exchange (2-D cuFFT, pack / tile-transpose kernels, grouped
 * ncclSend/ncclRecv, contiguous 1-D cuFFT along z). */
int baorec_dist_r2c_f32(baorec_ctx* ctx, const float* d_slab, float* d_kslab_t, baorec_stream stream);
int baorec_dist_c2r_f32(baorec_ctx* ctx, float* d_kslab_t /* destroyed */, float* d_slab, baorec_stream stream);
/* run! (src/recon.jl:134-180 box, :215-261 randoms) across the ranks: every rank passes the
 * particles (data and, with has_randoms != 0, randoms) whose base plane lies in its slab
 * (baorec_slab_owner_f32) and receives its slab of the result mesh (delta_r for BAOREC_ITERATIVE,
 * phi for BAOREC_MULTIGRID).  The box is the plan's: with randoms the host side all-reduces
 * setup_box (src/utils.jl:100-109) over the ranks before planning.
 *   scatter into slab + ghost plane (sent to the next rank and added), distributed R2C, k-space
 *   passes in the transposed layout (DC modes broadcast from rank 0), distributed C2R;
 *   IterativeRecon, fixed LOS: all iterations in one k-space pass; radial LOS: iterate!
 *   (src/iterative.jl:151-211) on slabs with 1 R2C + 6 C2R per iteration;
 *   MultigridRecon: fmg (src/multigrid.jl:722-752) on slabs, one halo-plane exchange with both
 *   ring neighbours after every Jacobi sweep / residual / prolongation; levels smaller than
 *   option "mg_slab_min_cells" (default 2^22 cells: 128^3 and below) are all-gathered and solved on every rank.
 * delta_k (phi_k) is kept, transposed, for baorec_read_shifts_dist_f32.
 * p->mas = BAOREC_MAS_TSC: the 27-point stencil reaches one plane below the slab and two above it; the deposit goes
 * into a (nz_loc + 3)-plane scratch and the boundary planes are exchanged with both ring neighbours and added
 * (needs nz_loc >= 2).  Particles are sharded exactly as for CIC (baorec_slab_owner_f32). */
int baorec_run_dist_f32(baorec_ctx* ctx, const baorec_params* p, int algorithm, float* d_x, float* d_y, float* d_z,
                        const float* d_w, int64_t n_local, float* d_rx, float* d_ry, float* d_rz, const float* d_rw,
                        int64_t n_ran_local, int has_randoms, float* d_mesh_slab, baorec_stream stream);
/* read_shifts / reconstructed_positions (positions != 0) for this rank's particles against the
 * result of the last baorec_run_dist_f32: three distributed C2R of i k delta_k / k^2 (i k phi_k
 * for MultigridRecon), halo planes exchanged with both neighbours, gather. */
int baorec_read_shifts_dist_f32(baorec_ctx* ctx, const baorec_params* p, const float* d_x, const float* d_y,
                                const float* d_z, int64_t n_local, int field, int positions, float* d_sx,
                                float* d_sy, float* d_sz, baorec_stream stream);
/* run! + read_shifts / reconstructed_positions from an UNSHARDED catalog: every rank passes any part of the data (and,
 * has_randoms != 0, randoms) catalog as device arrays; the library shards by slab (baorec_shard_catalog_f32), runs
 * baorec_run_dist_f32 + baorec_read_shifts_dist_f32 on the slabs and routes the results back into the caller's order.
 * The box is the planned one (with randoms: setup_box over all ranks first, src/recon.jl:172,253). */
int baorec_reconstruct_dist_f32(baorec_ctx* ctx, const baorec_params* p, int algorithm, const float* d_x, const float* d_y,
                                const float* d_z, const float* d_w, int64_t n, const float* d_rx, const float* d_ry,
                                const float* d_rz, const float* d_rw, int64_t n_ran, int has_randoms, int field,
                                int positions, float* d_ox, float* d_oy, float* d_oz, baorec_stream stream);
/* The same for HOST arrays (periodic box, no randoms): upload of this rank's share, the call above, download. */
int baorec_reconstruct_dist_host_f32(baorec_ctx* ctx, const baorec_params* p, int algorithm, const float* h_x,
                                     const float* h_y, const float* h_z, const float* h_w, int64_t n, int field,
                                     int positions, float* h_ox, float* h_oy, float* h_oz);

/* ---- mass assignment (replaces cic!/read_cic! CuArray methods) ------------- */
/* cic!(rho::CuArray, ...; wrap) src/mas.jl:53-107.  Accumulates into d_rho (caller
 * zero-fills).  With wrap != 0 the wrapped positions are written back into
 * d_x/d_y/d_z exactly as the reference does (src/mas.jl:56-60).  Synchronises
 * `stream`; returns BAOREC_ERR_OUT_OF_BOX if any particle was outside the mesh
 * (those particles are skipped). */
int baorec_cic_scatter_f32(baorec_ctx* ctx, float* d_rho, float* d_x, float* d_y, float* d_z,
                           const float* d_w, int64_t n, int wrap, int mas, baorec_stream stream);
/* Parity probe: the cell indices (0-based) and interpolation weights the
 * scatter uses, per particle, axis-major: i0/i1/w0/w1 are [3][n]. Does not
 * modify positions. (src/mas.jl:7-35) */
int baorec_cic_cells_f32(baorec_ctx* ctx, const float* d_x, const float* d_y, const float* d_z, int64_t n,
                         int wrap, int32_t* d_i0, int32_t* d_i1, float* d_w0, float* d_w1, baorec_stream stream);
/* Same for the gather (read_cic!, src/mas.jl:224-255 CPU formula when
 * gpu_formula == 0, :274-306 GPU formula otherwise). */
int baorec_gather_cells_f32(baorec_ctx* ctx, const float* d_x, const float* d_y, const float* d_z, int64_t n,
                            int gpu_formula, int32_t* d_id, int32_t* d_iu, float* d_wd, float* d_wu,
                            baorec_stream stream);
/* read_cic!(out::CuArray, field::CuArray, ...) src/mas.jl:271-327 (always periodic). */
int baorec_gather_f32(baorec_ctx* ctx, const float* d_field, const float* d_x, const float* d_y,
                      const float* d_z, int64_t n, float* d_out, int mas, baorec_stream stream);

/* setup_box (src/utils.jl:100-109) on device arrays. */
int baorec_setup_box_f32(baorec_ctx* ctx, const float* d_x, const float* d_y, const float* d_z, int64_t n,
                         float pad, float box_size_out[3], float box_min_out[3], baorec_stream stream);

/* ---- mesh set-up ------------------------------------------------------------ */
/* smooth!(field::CuArray, R, L, plan) src/utils.jl:85-96. In place. */
int baorec_smooth_f32(baorec_ctx* ctx, float* d_mesh, float smoothing_radius, baorec_stream stream);
/* setup_overdensity! src/recon.jl:42-57 (n_ran == 0: periodic box, wrap as given)
 * and :60-91 (randoms; wrap ignored = false).  d_mesh is zero-filled by the
 * caller (run! does it) and receives delta. Synchronises `stream`. */
int baorec_setup_overdensity_f32(baorec_ctx* ctx, const baorec_params* p, float* d_mesh,
                                 float* d_x, float* d_y, float* d_z, const float* d_w, int64_t n,
                                 float* d_rx, float* d_ry, float* d_rz, const float* d_rw, int64_t n_ran,
                                 int wrap, baorec_stream stream);

/* ---- iterative solver --------------------------------------------------------- */
/* iterate!(δ_r::CuArray, δ_s, k⃗, iter, β, plan; r̂, x⃗) src/iterative.jl:151-211;
 * `iter` is 1-based; d_los == NULL means radial (uses the x table of the plan). */
int baorec_iterate_f32(baorec_ctx* ctx, float* d_delta_r, const float* d_delta_s, int iter, float beta,
                       const float* h_los_or_null, baorec_stream stream);
/* reconstructed_overdensity! src/recon.jl:93-132 (set-up + n_iter iterations). */
int baorec_reconstructed_overdensity_f32(baorec_ctx* ctx, const baorec_params* p, float* d_mesh,
                                         float* d_x, float* d_y, float* d_z, const float* d_w, int64_t n,
                                         float* d_rx, float* d_ry, float* d_rz, const float* d_rw, int64_t n_ran,
                                         baorec_stream stream);

/* ---- multigrid solver (src/multigrid.jl) ---------------------------------------- */
/* All meshes are (nx,ny,nz) device arrays of the *given* size, which may be any
 * level of the hierarchy; the box is the plan's.  h_los == NULL -> radial. */
int baorec_mg_jacobi_f32(baorec_ctx* ctx, float* d_v, const float* d_f, int nx, int ny, int nz, float beta,
                         float damping, int niterations, const float* h_los_or_null, baorec_stream stream); /* :193-210 */
int baorec_mg_residual_f32(baorec_ctx* ctx, float* d_r, const float* d_v, const float* d_f, int nx, int ny, int nz,
                           float beta, const float* h_los_or_null, baorec_stream stream);                /* :378-395 */
int baorec_mg_restrict_f32(baorec_ctx* ctx, float* d_v2h, const float* d_v1h, int nx, int ny, int nz,
                           baorec_stream stream); /* reduce! :641-651 ; (nx,ny,nz) = fine size */
int baorec_mg_prolong_f32(baorec_ctx* ctx, float* d_v1h, const float* d_v2h, int nx, int ny, int nz,
                          baorec_stream stream);  /* prolong! :511-518 ; (nx,ny,nz) = fine size */
int baorec_mg_vcycle_f32(baorec_ctx* ctx, float* d_v, const float* d_f, float beta, float damping,
                         int niterations, const float* h_los_or_null, baorec_stream stream);       /* :689-719, plan size */
int baorec_mg_fmg_f32(baorec_ctx* ctx, const float* d_f, float* d_v, float beta, float damping, int n_jacobi,
                      int n_vcycle, const float* h_los_or_null, baorec_stream stream);              /* :722-752, plan size */
/* reconstructed_potential! src/recon.jl:184-212. */
int baorec_reconstructed_potential_f32(baorec_ctx* ctx, const baorec_params* p, float* d_phi,
                                       float* d_x, float* d_y, float* d_z, const float* d_w, int64_t n,
                                       float* d_rx, float* d_ry, float* d_rz, const float* d_rw, int64_t n_ran,
                                       baorec_stream stream);

/* ---- read-back ------------------------------------------------------------------ */
/* compute_displacements(mesh::CuArray, x,y,z, recon) src/iterative.jl:229-250
 * (algorithm BAOREC_ITERATIVE: Psi = irfft(i k delta_k/k^2)) and
 * src/multigrid.jl:781-799 (BAOREC_MULTIGRID: Psi = irfft(i k phi_k)). */
int baorec_compute_displacements_f32(baorec_ctx* ctx, const float* d_mesh, int algorithm,
                                     const float* d_x, const float* d_y, const float* d_z, int64_t n,
                                     float* d_px, float* d_py, float* d_pz, int mas, baorec_stream stream);
/* read_shifts(recon, x,y,z, mesh; field) src/recon.jl:333-364. */
int baorec_read_shifts_f32(baorec_ctx* ctx, const baorec_params* p, int algorithm, const float* d_mesh,
                           const float* d_x, const float* d_y, const float* d_z, int64_t n, int field,
                           float* d_sx, float* d_sy, float* d_sz, baorec_stream stream);
/* reconstructed_positions src/recon.jl:366-388: pos - shift, no re-wrap. */
int baorec_reconstructed_positions_f32(baorec_ctx* ctx, const baorec_params* p, int algorithm, const float* d_mesh,
                                       const float* d_x, const float* d_y, const float* d_z, int64_t n, int field,
                                       float* d_ox, float* d_oy, float* d_oz, baorec_stream stream);
/* read_shifts / reconstructed_positions (positions != 0) against recon.result_cache
 * (src/recon.jl:380-388 passes recon.result_cache as the mesh): the caller asserts that d_mesh is
 * the unmodified mesh produced by the last reconstructed_overdensity! on this context.  The library
 * keeps delta_k of that result (option "keep_delta_k", default 1), so the forward transform the
 * reference repeats on every call (src/iterative.jl:236) is skipped; if the pointer does not match
 * or no delta_k is held the call behaves exactly like baorec_read_shifts_f32. */
int baorec_read_result_cache_f32(baorec_ctx* ctx, const baorec_params* p, int algorithm, const float* d_mesh,
                                 const float* d_x, const float* d_y, const float* d_z, int64_t n, int field,
                                 int positions, float* d_ox, float* d_oy, float* d_oz, baorec_stream stream);
/* The three real-space displacement meshes themselves (parity probe). */
int baorec_displacement_meshes_f32(baorec_ctx* ctx, const float* d_mesh, int algorithm,
                                   float* d_psix, float* d_psiy, float* d_psiz, baorec_stream stream);

/* ---- whole pipeline, host buffers (what run! + reconstructed_positions do) ------- */
/* run!(recon, grid_size, data..., [rand...]) src/recon.jl:134-180 / 215-261 with HOST
 * catalogs: uploads them, (randoms: setup_box with p->box_pad and baorec_set_box,
 * returned in box_*_out), zero-fills an internal result mesh (recon.result_cache),
 * solves.  h_mesh_out may be NULL (mesh stays on the device as the cache).  When
 * wrap applies, the wrapped positions are written back to h_x/h_y/h_z like cic!. */
int baorec_run_host_f32(baorec_ctx* ctx, const baorec_params* p, int algorithm,
                        float* h_x, float* h_y, float* h_z, const float* h_w, int64_t n,
                        const float* h_rx, const float* h_ry, const float* h_rz, const float* h_rw, int64_t n_ran,
                        float* h_mesh_out, float box_size_out[3], float box_min_out[3]);
/* reconstructed_positions / read_shifts on HOST catalogs against the cached mesh
 * (or h_mesh_or_null uploaded first).  shifts_only != 0 -> read_shifts. */
int baorec_read_host_f32(baorec_ctx* ctx, const baorec_params* p, int algorithm, const float* h_mesh_or_null,
                         const float* h_x, const float* h_y, const float* h_z, int64_t n, int field,
                         int shifts_only, float* h_ox, float* h_oy, float* h_oz);
/* Many catalogs per process (README.md:11 "one process, many reconstructions"; the examples loop over mocks):
 * for every catalog i < n_catalogs, run! (periodic box, no randoms; src/recon.jl:134-180 / 215-261) followed by
 * reconstructed_positions / read_shifts of THAT catalog (src/recon.jl:333-388) -- baorec_run_host_f32 +
 * baorec_read_host_f32 per catalog, software-pipelined: the upload of catalog i+1 and the download of the results of
 * catalog i-1 overlap the reconstruction of catalog i (catalogs and results double-buffered on the device; the
 * read-back uses the resident positions and the tile sort run! made).  h_x[i] ... are HOST arrays of n[i] floats
 * (pinned memory for the overlap to happen); wrapped positions are written back to h_x/h_y/h_z like cic!.
 * shifts_only != 0 -> read_shifts.  Afterwards the context's result cache is the mesh of the last catalog. */
int baorec_batch_host_f32(baorec_ctx* ctx, const baorec_params* p, int algorithm, int n_catalogs, float* const* h_x,
                          float* const* h_y, float* const* h_z, const float* const* h_w, const int64_t* n, int field,
                          int shifts_only, float* const* h_ox, float* const* h_oy, float* const* h_oz);
/* The same pipeline fed from catalog FILES and writing files -- examples/simulation.jl:12-40 looped over mocks:
 *   in_paths[i]  text catalog (fields separated as `delim` says, see baorec_text_catalog_scan) or, by its extension,
 *                an .npy matrix (N, K); cols[0..2] = the 0-based columns of x, y, z, cols[3] = the column of the
 *                weights or -1 for weights of one (examples/simulation.jl:16);
 *   out_paths[i] NPY file of the reconstructed positions (shifts_only != 0: of the shifts) as the (N, 3) Float32
 *                matrix npzwrite(fn, hcat(new_pos...)) stores; out_paths or any entry may be NULL (nothing written).
 * A reader thread parses catalog i+1 (and further ahead, as buffer sets come free) into pinned host memory and a
 * writer thread stores the results of catalog i-1 while the device reconstructs catalog i; the catalogs may differ in
 * size (staging grows when a larger one arrives).  n_threads: parser threads (<= 0: one per core).  n_rows (NULL or
 * n_catalogs entries): rows of each catalog.  seconds (NULL or 4 doubles): time the reader spent reading + parsing,
 * the writer writing, the pipeline waiting for the reader, and the whole call.  Option "batch_slots" (default 4, >= 3)
 * sets the number of host buffer sets.  Errors: BAOREC_ERR_IO with the file (and line) named. */
int baorec_batch_files_f32(baorec_ctx* ctx, const baorec_params* p, int algorithm, int n_catalogs,
                           const char* const* in_paths, char delim, const int cols[4], int field, int shifts_only,
                           const char* const* out_paths, int n_threads, int64_t* n_rows, double* seconds);
/* Device pointer of the cached result mesh (recon.result_cache), or NULL. */
float* baorec_result_cache(baorec_ctx* ctx);

/* ---- catalog pre/post-processing (the callers either side of the path) ------ */
/* The density parameters of `Cosmology` (src/cosmo.jl:22-64) after its Float32 derivations, widened
 * to double, and the table range (`z_tab_min`, `z_tab_max`, `z_tab_num`). */
typedef struct baorec_cosmology {
  double h;                               /* Float32(h) */
  double Omega_b0, Omega_c0, Omega_nu0, Omega_g0, Omega_k0, Omega_L0;
  double w0, wa;
  double z_tab_min, z_tab_max;
  int64_t z_tab_num;
} baorec_cosmology;
/* Replaces the cache built by comoving_distance_interp / redshift_interp (src/cosmo.jl:86-103):
 * r(z) = c * int_0^z dz'/H(z') [Mpc] on z = range(z_tab_min, z_tab_max, length = z_tab_num), Float64,
 * by 7-point Gauss-Legendre quadrature per knot interval (the reference: quadgk, rtol 1e-8), kept on
 * the host and uploaded to the device.  Needs no baorec_plan. */
int baorec_cosmo_set(baorec_ctx* ctx, const baorec_cosmology* c);
/* The same table on the host only (no context, no device): h_r[z_tab_num] (and the knots in h_z unless
 * NULL).  Set-up arithmetic like k_vec / x_vec; lets a caller inspect or cache the table. */
int baorec_cosmo_build_table(const baorec_cosmology* c, double* h_z, double* h_r);
/* Host copy of the tables (parity probe); either output may be NULL; cap >= z_tab_num. */
int baorec_cosmo_tables(const baorec_ctx* ctx, double* h_z, double* h_r, int64_t cap);
/* sky_to_cartesian (examples/lightcone.jl:30-49): ra, dec [deg], redshift -> x, y, z [Mpc/h];
 * dist = comoving_distance_interp(redshift), h = Float32(cosmo.H0 / 100).  Outputs may alias inputs.
 * Synchronises `stream`; BAOREC_ERR_OUT_OF_RANGE if a redshift lies outside the table (outputs NaN). */
int baorec_sky_to_cartesian_f32(baorec_ctx* ctx, const float* d_ra, const float* d_dec, const float* d_red, int64_t n,
                                float h, float* d_x, float* d_y, float* d_z, baorec_stream stream);
/* cartesian_to_sky (examples/lightcone.jl:51-80): x, y, z [Mpc/h] -> ra, dec [deg], redshift =
 * redshift_interp(|p| / h) with h = cosmo.h.  The right ascension is left in (-360, 0] exactly like
 * the reference's `(lon - 360) % 360`.  Synchronises `stream`; BAOREC_ERR_OUT_OF_RANGE as above. */
int baorec_cartesian_to_sky_f32(baorec_ctx* ctx, const float* d_x, const float* d_y, const float* d_z, int64_t n,
                                float h, float* d_ra, float* d_dec, float* d_red, baorec_stream stream);
/* fkp_weights.(nz, P0) (examples/lightcone.jl:82): w = 1 / (1 + nz * P0), Float32, bit-exact. */
int baorec_fkp_weights_f32(baorec_ctx* ctx, const float* d_nz, int64_t n, float P0, float* d_w, baorec_stream stream);
/* Periodic re-wrap of reconstructed positions, in place: p = min + mod(p - min + L, L) (floored
 * modulo; test_helpers/simulation.py:38,51-52 with min = 0).  reconstructed_positions itself does not
 * re-wrap (src/recon.jl:366-388). */
int baorec_wrap_positions_f32(baorec_ctx* ctx, float* d_x, float* d_y, float* d_z, int64_t n, const float box_size[3],
                              const float box_min[3], baorec_stream stream);

/* Power-spectrum multipoles of a density mesh (periodic box): what the reference checks by eye on every
 * reconstruction, through the third-party pypowspec (test_helpers/simulation.py:56-75, pyrecon_simulation.py:70-86:
 * compute_auto_box on the catalogs before / after, line of sight along z, l = 0, 2, 4, linear bins).
 *   d_rho: density mesh from baorec_cic_scatter_f32 on this context's grid (any normalisation; not modified);
 *   d_ran: NULL, or the density mesh of a (shifted) random catalog: the field is then rho/sum(rho) - ran/sum(ran),
 *          "data minus shifted randoms" (compute_auto_box_rand, test_helpers/simulation.py:52-70: RecIso / RecSym);
 *   P(k) = V |rho_k|^2 / rho_0^2 / W(k)^2, W = prod_a sinc(k_a h_a / 2)^mas_power (2 = CIC, 3 = TSC, 4 = PCS, 0 = none);
 *   P_l(bin i) = (2l+1) <P L_l(mu)> over the modes with kmin + i dk <= |k| < kmin + (i+1) dk, mu = k.los/|k|,
 *   every mode of the Hermitian mesh counted once, k = 0 excluded; `shot` is subtracted from the monopole.
 * One R2C per mesh + one pass over the half mesh (Float64 sums).  Outputs are HOST arrays of nbins doubles (mean k of the
 * bin, number of modes, P_0, P_2, P_4; NaN where a bin is empty).  nbins <= 1024.  Synchronises `stream`. */
int baorec_power_multipoles_f32(baorec_ctx* ctx, const float* d_rho, const float* d_ran, const float los[3], double kmin,
                                double dk, int nbins, int mas_power, double shot, double* h_k, double* h_nmodes, double* h_p0,
                                double* h_p2, double* h_p4, baorec_stream stream);

/* The same estimator with interlacing (Sefusatti et al. 2016; GRID_INTERLACE = T in the reference's helper configuration,
 * test_helpers/powspec_auto.conf:125): d_rho_shifted (d_ran_shifted) is the mesh of the same catalog painted half a cell
 * further along every axis (baorec_interlace_positions_f32); the two transforms are combined per mode as
 * [f1 + f2 exp(i (k_x h_x + k_y h_y + k_z h_z) / 2)] / 2, which cancels the aliasing images k + 2 k_N m with
 * m_x + m_y + m_z odd.  mas_power: 2 = CIC, 3 = TSC, 4 = PCS.  Two R2C transforms + one pass over both half meshes. */
int baorec_power_multipoles_interlaced_f32(baorec_ctx* ctx, const float* d_rho, const float* d_ran, const float* d_rho_shifted,
                                           const float* d_ran_shifted, const float los[3], double kmin, double dk, int nbins,
                                           int mas_power, double shot, double* h_k, double* h_nmodes, double* h_p0,
                                           double* h_p2, double* h_p4, baorec_stream stream);

/* Positions of the interlaced mesh on this context's grid: q = p + (L/n)/2 per axis in Float32, minus L where
 * q - min >= L.  Output arrays may not alias the inputs. */
int baorec_interlace_positions_f32(baorec_ctx* ctx, const float* d_x, const float* d_y, const float* d_z, int64_t n, float* d_ox,
                                   float* d_oy, float* d_oz, baorec_stream stream);

/* compute_auto_box / compute_auto_box_rand of the reference's helpers (test_helpers/simulation.py:36-70, settings of
 * test_helpers/powspec_auto.conf: TSC + interlacing on a 512^3 grid) in one call on this context's grid and box: paint
 * the catalog (mas = BAOREC_MAS_CIC / TSC / PCS, periodic; the caller's arrays are not modified), optionally the
 * randoms (nr > 0: the field is data/sum - randoms/sum, "data minus shifted randoms") and the interlaced meshes
 * (interlace != 0), then the multipoles as above with the window exponent of the scheme. */
int baorec_compute_auto_box_f32(baorec_ctx* ctx, const float* d_x, const float* d_y, const float* d_z, const float* d_w, int64_t n,
                                const float* d_rx, const float* d_ry, const float* d_rz, const float* d_rw, int64_t nr, int mas,
                                int interlace, const float los[3], double kmin, double dk, int nbins, double shot, double* h_k,
                                double* h_nmodes, double* h_p0, double* h_p2, double* h_p4, baorec_stream stream);

/* ---- catalog files (the data formats either side of the path) ------------------------------
 * Host code: none of these needs a context or a GPU.  They fill / read caller-owned SoA Float32 arrays -- the layout
 * baorec_run_host_f32 and baorec_batch_host_f32 take; give them pinned memory (baorec_host_alloc) and the uploads
 * that follow are asynchronous.  n_threads <= 0: one thread per host core (never more than one per 4 MiB of work). */
enum { BAOREC_DTYPE_F32 = 4, BAOREC_DTYPE_F64 = 8 };
/* Rows and columns of a delimited text catalog -- what CSV.File(fn, delim = ' ', ignorerepeated = true,
 * header = [...], types = [Float32 ...]) reads (examples/simulation.jl:12-13, examples/lightcone.jl:22-23).
 * delim == ' ': fields are separated by runs of blanks / tabs (ignorerepeated); any other character separates
 * exactly, blanks around a field are dropped.  Blank lines and lines starting with '#' are not rows.  n_cols = the
 * fields of the first row (may be NULL). */
int baorec_text_catalog_scan(const char* path, char delim, int64_t* n_rows, int* n_cols, int n_threads);
/* Reads the file columns cols[0..n_out) (0-based) into h_out[j][0..*n_read): correctly rounded Float32 (like
 * CSV.jl's Float32 parser), `inf` / `nan` accepted.  BAOREC_ERR_INVALID when the file holds more than `capacity`
 * rows; BAOREC_ERR_IO naming the line when a requested field is missing or not a number. */
int baorec_text_catalog_read_f32(const char* path, char delim, int n_out, const int* cols, float* const* h_out,
                                 int64_t capacity, int64_t* n_read, int n_threads);
/* Header of an NPY file (format 1.0 - 3.0; little-endian f4 / f8; a vector (N,) or a matrix (N, K) in either
 * memory order; npzread in the reference's scripts).  Any output may be NULL. */
int baorec_npy_info(const char* path, int* dtype, int* fortran_order, int64_t* n_rows, int64_t* n_cols);
/* Columns cols[0..n_out) of the (N, K) matrix into h_out[j][0..N) (Float64 data are rounded to Float32). */
int baorec_npy_read_columns_f32(const char* path, int n_out, const int* cols, float* const* h_out, int64_t capacity,
                                int64_t* n_read, int n_threads);
/* npzwrite(fn, hcat(cols...)) (examples/simulation.jl:38-40): the SoA columns written as an (n, n_cols) Float32
 * matrix with 'fortran_order': True -- byte for byte what NPZ.jl stores for the column-major matrix, and what
 * numpy.load returns as an (n, n_cols) array.  n_cols == 1 writes a vector (n,). */
int baorec_npy_write_columns_f32(const char* path, int n_cols, const float* const* h_cols, int64_t n);
/* data_cat[map(z -> ((z > lo) & (z < hi)), data_cat.z), :] (examples/lightcone.jl:25-26): keeps the rows with
 * lo < h_cols[key][r] < hi, in order, in place in every column; *n_kept rows remain. */
int baorec_catalog_select_f32(int n_cols, float* const* h_cols, int64_t n, int key, float lo, float hi, int64_t* n_kept,
                              int n_threads);

/* Pinned host memory helpers for callers without their own (Julia: CUDA.Mem.alloc(HostBuffer)). */
int baorec_host_alloc(void** out, int64_t bytes);
int baorec_host_free(void* p);

#ifdef __cplusplus
}
#endif
#endif /* BAOREC_B200_H */
