// Multi-GPU path: one process per GPU, the mesh decomposed into z slabs.
//
//  real space : rank r owns planes [r nz/P, (r+1) nz/P)            slab[zl][y][x]
//  k space    : rank r owns rows   [r ny/P, (r+1) ny/P), z contiguous   T[yl][ix][iz]
//
//  Two exchange schemes (option "dist_exchange"):
//  1 = peer copies (default; needs the receive buffers of all ranks mapped through CUDA IPC, dist.plan does that):
//      k space is K[z][yl][x] -- the layout the per-peer blocks of the 2-D transform output already have, so there is
//      NO pack and NO transpose kernel.  forward = batched 2-D R2C over a chunk of planes -> one strided copy per peer
//      (copy engines, NVLink, no SMs) straight into the peer's K buffer + a flag store -> strided 1-D C2C along z;
//      inverse mirrors it (the copies land in the peer's plane-layout buffer, the 2-D C2R reads it directly).
//      Flow control = per-peer sequence flags in peer memory (arrived / buffer-free), no barrier, no NCCL.
//  0 = NCCL (round 1): k space is T[yl][ix][iz]; 2-D R2C -> pack (row copy) -> grouped ncclSend/ncclRecv
//      all-to-all -> unpack with an (x,z) tile transpose -> contiguous 1-D C2C along z; inverse mirrors it.
//  Particles are owned by the slab of their base cell: the scatter writes one ghost plane that
//  is sent to the next rank and added; the gather reads three halo planes.
#include <nccl.h>
#include <string.h>

#include "internal.cuh"

using namespace baorec;

#define BR_NCCL(expr)                                                                              \
  do {                                                                                             \
    ncclResult_t _e = (expr);                                                                      \
    if (_e != ncclSuccess) {                                                                       \
      baorec::set_error("%s:%d NCCL error %s in %s", __FILE__, __LINE__, ncclGetErrorString(_e), #expr); \
      return BAOREC_ERR_NCCL;                                                                      \
    }                                                                                              \
  } while (0)

namespace baorec {

// ---- pack / unpack kernels ---------------------------------------------------------------------------
// Row copy between A[zl][y][x] and the per-peer blocks B[peer][zl][yl][x]  (y = peer*nyl + yl).
template <bool TO_BLOCKS>
__global__ void __launch_bounds__(128)
rows_kernel(float2* __restrict__ dst, const float2* __restrict__ src, int nzl, int ny, int nyl, int xh,
            unsigned row0) {
  const unsigned row = row0 + blockIdx.x;  // zl * ny + y
  const int zl = row / ny, y = row - zl * ny;
  const int peer = y / nyl, yl = y - peer * nyl;
  const size_t a = (size_t)row * xh;
  const size_t b = (((size_t)peer * nzl + zl) * nyl + yl) * xh;
  const float2* s = src + (TO_BLOCKS ? a : b);
  float2* d = dst + (TO_BLOCKS ? b : a);
  for (int x = threadIdx.x; x < xh; x += blockDim.x) d[x] = s[x];
}

struct PeerTab {
  float2* p[16];
};

// Batched strided 2-D transpose through shared memory: dst[c][r] = src[r][c].
struct TrGeom {
  int rows, cols;
  size_t src_row_stride, dst_row_stride;
  int nb1;  // batch index b -> (b0 = b / nb1, b1 = b % nb1)
  size_t src_b0, src_b1, dst_b0, dst_b1;
};

// (`tab` is __grid_constant__: indexing it with the run-time b0 reads the constant bank; passed by plain value the
// compiler copied all 16 pointers to every thread's stack first -- 16 local stores per thread, found by
// benchmarks/sass_census.py.)
// If `tab` is given, batch b0 (= destination peer) is written into tab.p[b0] + tab_off instead of
// dst + b0 * dst_b0 (fused transpose + exchange over peer memory).
__global__ void __launch_bounds__(256)
transpose_kernel(float2* __restrict__ dst, const float2* __restrict__ src, TrGeom t, const __grid_constant__ PeerTab tab,
                 int use_tab,
                 size_t tab_off) {
  __shared__ float2 tile[32][33];
  const int b = blockIdx.z, b0 = b / t.nb1, b1 = b - b0 * t.nb1;
  const float2* s = src + b0 * t.src_b0 + b1 * t.src_b1;
  float2* d = (use_tab ? tab.p[b0] + tab_off : dst + b0 * t.dst_b0) + b1 * t.dst_b1;
  const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
#pragma unroll
  for (int k = 0; k < 4; k++) {
    int r = r0 + ty + 8 * k, c = c0 + tx;
    if (r < t.rows && c < t.cols) tile[ty + 8 * k][tx] = s[(size_t)r * t.src_row_stride + c];
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < 4; k++) {
    int c = c0 + ty + 8 * k, r = r0 + tx;
    if (r < t.rows && c < t.cols) d[(size_t)c * t.dst_row_stride + r] = tile[tx][ty + 8 * k];
  }
}

__global__ void add_plane_kernel(float* __restrict__ dst, const float* __restrict__ src, size_t n) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] += src[i];
}

__global__ void slab_owner_kernel(const float* __restrict__ z, int64_t n, float mn, float L, int nz, int nzl,
                                  int32_t* __restrict__ owner) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  // base plane of cic! (src/mas.jl:15,19,30) after the upper-face wrap (:10)
  float p = z[i];
  if (__fsub_rn(p, mn) > L) p = __fsub_rn(p, L);
  float g = __fadd_rn(__fdiv_rn(__fmul_rn(__fsub_rn(p, mn), (float)nz), L), 1.0f);
  int c0 = (int)floorf(g);
  if (c0 == nz + 1) c0 = 1;
  owner[i] = (g >= 1.0f && c0 >= 1 && c0 <= nz) ? (c0 - 1) / nzl : -1;
}

static ncclComm_t comm_of(baorec_ctx* ctx) { return (ncclComm_t)ctx->comm; }

// defined with the peer-copy machinery below
static bool peer_halos(const baorec_ctx* ctx);
static int halo_exchange_peer(baorec_ctx* ctx, int nx, const float* const send[2], const int to[2], float* const recv[2],
                              const int from[2], const size_t count[2], cudaStream_t st);

// blocks of `blk` complex values per peer: S[peer] -> R[peer]
static int all_to_all(baorec_ctx* ctx, const float2* S, float2* R, size_t blk, cudaStream_t st) {
  const int P = ctx->nranks;
  if (P == 1) {
    BR_CUDA(cudaMemcpyAsync(R, S, blk * sizeof(float2), cudaMemcpyDeviceToDevice, st));
    return BAOREC_OK;
  }
  int pi = prof_begin(ctx, "nccl_all_to_all", st);
  BR_NCCL(ncclGroupStart());
  for (int peer = 0; peer < P; peer++) {
    BR_NCCL(ncclSend(S + (size_t)peer * blk, blk * 2, ncclFloat, peer, comm_of(ctx), st));
    BR_NCCL(ncclRecv(R + (size_t)peer * blk, blk * 2, ncclFloat, peer, comm_of(ctx), st));
  }
  BR_NCCL(ncclGroupEnd());
  prof_end(ctx, pi, st);
  return BAOREC_OK;
}

// send `count` floats to rank `to`, receive as many from rank `from`
static int ring_exchange(baorec_ctx* ctx, const float* send, int to, float* recv, int from, size_t count,
                         cudaStream_t st) {
  if (ctx->nranks == 1) {
    BR_CUDA(cudaMemcpyAsync(recv, send, count * sizeof(float), cudaMemcpyDeviceToDevice, st));
    return BAOREC_OK;
  }
  if (peer_halos(ctx) && count % 4 == 0 && count <= ctx->inbox_slot_floats) {
    const float* sp[2] = {send, nullptr};
    float* rp[2] = {recv, nullptr};
    const int t2[2] = {to, to}, f2[2] = {from, from};
    const size_t c2[2] = {count, 0};
    return halo_exchange_peer(ctx, 1, sp, t2, rp, f2, c2, st);
  }
  int pi = prof_begin(ctx, "nccl_halo", st);
  BR_NCCL(ncclGroupStart());
  BR_NCCL(ncclSend(send, count, ncclFloat, to, comm_of(ctx), st));
  BR_NCCL(ncclRecv(recv, count, ncclFloat, from, comm_of(ctx), st));
  BR_NCCL(ncclGroupEnd());
  prof_end(ctx, pi, st);
  return BAOREC_OK;
}

// Both halo planes of a slab-layout buffer ((nzl+2) planes, real planes at 1..nzl): plane 0 <- last
// real plane of the previous rank, plane nzl+1 <- first real plane of the next rank (periodic ring).
// One NCCL group of two sends and two receives; with two ranks both neighbours are the same peer
// and the in-order matching of NCCL pairs (last plane -> plane 0) and (first plane -> plane nzl+1).
int mg_halo_exchange(baorec_ctx* ctx, float* buf, size_t plane, int nzl, cudaStream_t st) {
  float* lo = buf;
  float* first = buf + plane;
  float* last = buf + (size_t)nzl * plane;
  float* hi = buf + (size_t)(nzl + 1) * plane;
  const int P = ctx->nranks;
  if (P == 1) {
    BR_CUDA(cudaMemcpyAsync(lo, last, plane * sizeof(float), cudaMemcpyDeviceToDevice, st));
    BR_CUDA(cudaMemcpyAsync(hi, first, plane * sizeof(float), cudaMemcpyDeviceToDevice, st));
    return BAOREC_OK;
  }
  const int next = (ctx->rank + 1) % P, prev = (ctx->rank + P - 1) % P;
  if (peer_halos(ctx) && plane % 4 == 0 && plane <= ctx->inbox_slot_floats) {
    const float* sp[2] = {last, first};      // last plane -> next rank's plane 0; first plane -> previous rank's plane nzl+1
    float* rp[2] = {lo, hi};
    const int t2[2] = {next, prev}, f2[2] = {prev, next};
    const size_t c2[2] = {plane, plane};
    return halo_exchange_peer(ctx, 2, sp, t2, rp, f2, c2, st);
  }
  int pi = prof_begin(ctx, "nccl_mg_halo", st);
  BR_NCCL(ncclGroupStart());
  BR_NCCL(ncclSend(last, plane, ncclFloat, next, comm_of(ctx), st));
  BR_NCCL(ncclSend(first, plane, ncclFloat, prev, comm_of(ctx), st));
  BR_NCCL(ncclRecv(lo, plane, ncclFloat, prev, comm_of(ctx), st));
  BR_NCCL(ncclRecv(hi, plane, ncclFloat, next, comm_of(ctx), st));
  BR_NCCL(ncclGroupEnd());
  prof_end(ctx, pi, st);
  return BAOREC_OK;
}

// recv[r * count ...] = rank r's send[0 .. count)
int mg_allgather(baorec_ctx* ctx, const float* send, float* recv, size_t count, cudaStream_t st) {
  if (ctx->nranks == 1) {
    BR_CUDA(cudaMemcpyAsync(recv, send, count * sizeof(float), cudaMemcpyDeviceToDevice, st));
    return BAOREC_OK;
  }
  int pi = prof_begin(ctx, "nccl_mg_allgather", st);
  BR_NCCL(ncclAllGather(send, recv, count, ncclFloat, comm_of(ctx), st));
  prof_end(ctx, pi, st);
  return BAOREC_OK;
}

int mg_allreduce_sum(baorec_ctx* ctx, double* d_value, cudaStream_t st) {
  if (ctx->nranks > 1) BR_NCCL(ncclAllReduce(d_value, d_value, 1, ncclDouble, ncclSum, comm_of(ctx), st));
  return BAOREC_OK;
}

__global__ void set_scalar_kernel(double* p, double v) { *p = v; }

// ---- peer-copy exchange: flags --------------------------------------------------------------------------------
// Every rank owns one PeerFlags block (device memory, mapped by all peers through CUDA IPC).  Peers WRITE into it,
// the owner spins on it: sequence numbers only ever grow, so a late reader never misses a signal.
//   ARR_*[s]   chunks (or calls) of an exchange that rank s has finished copying into my buffer
//   FREE_*[d]  transforms (calls) after which rank d has finished reading ITS receive buffer (I may overwrite it)
// K = k-layout buffer of the forward transform; A0 / A1 = the two plane-layout buffers of the inverse transforms
// (two, so that the exchange of one displacement field overlaps the transforms of the other).
enum PeerFlagId { F_ARR_K = 0, F_FREE_K, F_ARR_A0, F_FREE_A0, F_ARR_A1, F_FREE_A1,
                  F_ARR_S0, F_FREE_S0, F_ARR_B0, F_FREE_B0, F_ARR_S1, F_FREE_S1, F_ARR_B1, F_FREE_B1,
                  F_ARR_H0, F_ARR_H1, F_FREE_H, F_COUNT = 20 };  // H: halo / ghost planes through the peers' inboxes  // S: sharded catalog, B: results coming back; 0 data, 1 randoms
struct PeerFlags {
  unsigned w[F_COUNT][16];
  unsigned err, pad[15];
};
struct FlagTab {
  unsigned* p[16];
};

// System-scope RELAXED accesses, no fences: the ordering comes from the stream.  The copies a flag announces are
// earlier operations of the signalling stream (complete, i.e. performed in the peer's memory, before the flag kernel
// starts), and whatever consumes the data is a later kernel of the waiting stream.
__device__ __forceinline__ unsigned ld_relaxed_sys(const unsigned* p) {
  unsigned v;
  asm volatile("ld.relaxed.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_relaxed_sys(unsigned* p, unsigned v) {
  asm volatile("st.relaxed.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned long long global_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

// thread t stores `value` into rank t's flag word for this rank (stream order puts it after the copies it announces)
__global__ void peer_signal_kernel(const __grid_constant__ FlagTab tab, unsigned value, int P) {
  const int t = threadIdx.x;
  if (t < P) st_relaxed_sys(tab.p[t], value);
}

// thread t waits until flags[t] >= value (wrap-safe); gives up after ~20 s and raises the error flag instead of
// hanging the stream forever when a peer died
__global__ void peer_wait_kernel(const unsigned* flags, unsigned value, int P, unsigned* err) {
  const int t = threadIdx.x;
  if (t < P) {
    const unsigned long long t0 = global_ns();
    while ((int)(ld_relaxed_sys(flags + t) - value) < 0) {
      __nanosleep(200);
      if (global_ns() - t0 > 20000000000ull) {
        atomicExch(err, 1u);
        break;
      }
    }
  }
  __syncthreads();
}

// The remote blocks of one chunk pushed by the SMs, to ALL peers at once: blockIdx.y = peer step k (destination
// (rank + k) mod P), PUSH_CTAS blocks per destination, 16-byte loads from the local rows and 16-byte stores into the
// peer's rows over NVLink.  Measured on 8 x B200 (benchmarks/probe/p2p_probe.cu, 16.8 MB per pair): 640 GB/s per
// direction against 430 GB/s for copy-engine copies in any stream arrangement -- the copy engines cannot keep seven
// peers busy, 112 thread blocks can.  With 2 - 4 ranks the copy engines are the faster (700 - 755 GB/s) and SM-free.
constexpr int PUSH_CTAS = 18;  // per destination: 126 blocks with 8 ranks
struct PushGeom {
  size_t src_peer_stride;        // source offset per destination rank (float2 elements)
  size_t src_pitch, dst_pitch;   // row pitches (float2 elements)
  size_t dst_off;                // offset of this rank's rows in every destination buffer
  unsigned w4;                   // row width in float4
  int rows, P, me;
};
__global__ void __launch_bounds__(256)
peer_push_kernel(const __grid_constant__ PeerTab dst, const float2* __restrict__ src, const PushGeom g) {
  const int d = (g.me + (int)blockIdx.y + 1) % g.P;
  const float2* s0 = src + (size_t)d * g.src_peer_stride;
  float2* o0 = dst.p[d] + g.dst_off;
  for (int row = blockIdx.x; row < g.rows; row += gridDim.x) {
    const float4* s = (const float4*)(s0 + (size_t)row * g.src_pitch);
    float4* o = (float4*)(o0 + (size_t)row * g.dst_pitch);
    unsigned i = threadIdx.x;
    for (; i + 3 * 256 < g.w4; i += 4 * 256) {
      const float4 a = s[i], b = s[i + 256], c = s[i + 512], e = s[i + 768];
      o[i] = a;
      o[i + 256] = b;
      o[i + 512] = c;
      o[i + 768] = e;
    }
    for (; i < g.w4; i += 256) o[i] = s[i];
  }
}

// ---- halo / ghost planes through peer memory ------------------------------------------------------------------------
// A ring exchange (every rank sends `count` floats to one neighbour and receives as many from another) as TWO small
// kernels instead of a grouped ncclSend / ncclRecv (~100 us each at 8 ranks for a 4 MB plane; a multigrid solve makes
// 700 of them): the send kernel stores the plane(s) straight into slot (call number mod HALO_SLOTS) of the neighbour's
// inbox and its last block raises the arrival flag; the receive kernel spins on the flag, copies inbox -> destination
// and its last block tells both neighbours that the slot is free again.  Up to two transfers per call (the multigrid
// halo exchange sends to both neighbours).
constexpr int HALO_SLOTS = 8;
struct HaloXfer {
  const float* src;   // send: local source            receive: my inbox slot
  float* dst;         // send: the neighbour's inbox   receive: local destination
  unsigned* flag;     // send: the neighbour's ARR_H word for me
  const unsigned* wait;  // send: my FREE_H word of that neighbour   receive: my ARR_H word of the sender
  size_t count;       // floats (multiple of 4)
};
struct HaloArgs {
  HaloXfer x[2];
  int nx;             // transfers in this call (1 or 2)
  unsigned seq;       // call number
  unsigned* done;     // block counter (zeroed by the last block)
  unsigned* free_flag[2];  // receive: the FREE_H words of my two neighbours that belong to me
  unsigned* err;
};

__device__ __forceinline__ bool spin_until(const unsigned* p, unsigned value, unsigned* err) {
  const unsigned long long t0 = global_ns();
  while ((int)(ld_relaxed_sys(p) - value) < 0) {
    __nanosleep(100);
    if (global_ns() - t0 > 20000000000ull) {
      atomicExch(err, 1u);
      return false;
    }
  }
  return true;
}

__global__ void __launch_bounds__(256) halo_send_kernel(const __grid_constant__ HaloArgs a) {
  if (threadIdx.x == 0) {  // the slot must have been consumed HALO_SLOTS calls ago
    for (int k = 0; k < a.nx; k++)
      if (a.seq > HALO_SLOTS) spin_until(a.x[k].wait, a.seq - HALO_SLOTS, a.err);
  }
  __syncthreads();
  for (int k = 0; k < a.nx; k++) {
    const float4* s = reinterpret_cast<const float4*>(a.x[k].src);
    float4* d = reinterpret_cast<float4*>(a.x[k].dst);
    const size_t n4 = a.x[k].count >> 2;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) d[i] = s[i];
  }
  __threadfence_system();   // my stores are performed in the neighbour's memory before the flag can be
  __syncthreads();
  if (threadIdx.x == 0) {
    if (atomicAdd(a.done, 1u) == gridDim.x - 1) {
      *a.done = 0u;
      __threadfence_system();
      for (int k = 0; k < a.nx; k++) st_relaxed_sys(a.x[k].flag, a.seq);
    }
  }
}

__global__ void __launch_bounds__(256) halo_recv_kernel(const __grid_constant__ HaloArgs a) {
  if (threadIdx.x == 0) {
    for (int k = 0; k < a.nx; k++) spin_until(a.x[k].wait, a.seq, a.err);
    __threadfence_system();
  }
  __syncthreads();
  for (int k = 0; k < a.nx; k++) {
    const float4* s = reinterpret_cast<const float4*>(a.x[k].src);
    float4* d = reinterpret_cast<float4*>(a.x[k].dst);
    const size_t n4 = a.x[k].count >> 2;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x)
      d[i] = __ldcg(s + i);   // the inbox was written by a peer: read it from L2, not from a stale L1 line
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    if (atomicAdd(a.done, 1u) == gridDim.x - 1) {
      *a.done = 0u;
      st_relaxed_sys(a.free_flag[0], a.seq);
      st_relaxed_sys(a.free_flag[1], a.seq);
    }
  }
}

static PeerFlags* own_flags(const baorec_ctx* ctx) { return (PeerFlags*)ctx->d_flags; }

static bool peer_halos(const baorec_ctx* ctx) { return ctx->p2p && ctx->nranks > 1 && ctx->opt_peer_halo && ctx->peer_inbox[ctx->rank]; }

// up to two transfers: send[k] -> rank to[k], recv[k] <- rank from[k]; counts in floats
static int halo_exchange_peer(baorec_ctx* ctx, int nx, const float* const send[2], const int to[2], float* const recv[2],
                              const int from[2], const size_t count[2], cudaStream_t st) {
  const int P = ctx->nranks, me = ctx->rank, next = (me + 1) % P, prev = (me + P - 1) % P;
  const unsigned seq = ++ctx->seq_halo;
  const size_t slot_floats = ctx->inbox_slot_floats;
  const size_t slot = (size_t)(seq % HALO_SLOTS) * 2;   // two half-slots per call
  PeerFlags* mine = own_flags(ctx);
  HaloArgs sa, ra;
  sa.nx = ra.nx = nx;
  sa.seq = ra.seq = seq;
  sa.done = ctx->d_halo_done;
  ra.done = ctx->d_halo_done + 1;
  sa.err = ra.err = &mine->err;
  sa.free_flag[0] = sa.free_flag[1] = nullptr;
  ra.free_flag[0] = ((PeerFlags*)ctx->peer_flags[prev])->w[F_FREE_H] + me;
  ra.free_flag[1] = ((PeerFlags*)ctx->peer_flags[next])->w[F_FREE_H] + me;
  for (int k = 0; k < nx; k++) {
    BR_REQUIRE(count[k] % 4 == 0 && count[k] <= slot_floats, "halo plane does not fit an inbox slot");
    // the two transfers of a call go to different half-slots so that prev == next (two ranks) does not collide
    sa.x[k].src = send[k];
    sa.x[k].dst = (float*)ctx->peer_inbox[to[k]] + (slot + k) * slot_floats;
    sa.x[k].flag = ((PeerFlags*)ctx->peer_flags[to[k]])->w[k ? F_ARR_H1 : F_ARR_H0] + me;
    sa.x[k].wait = mine->w[F_FREE_H] + to[k];
    sa.x[k].count = count[k];
    ra.x[k].src = (const float*)ctx->peer_inbox[me] + (slot + k) * slot_floats;
    ra.x[k].dst = recv[k];
    ra.x[k].flag = nullptr;
    ra.x[k].wait = mine->w[k ? F_ARR_H1 : F_ARR_H0] + from[k];
    ra.x[k].count = count[k];
  }
  const size_t work = (count[0] + (nx > 1 ? count[1] : 0)) / 4;
  unsigned grid = (unsigned)((work + 2047) / 2048);
  if (grid > 64) grid = 64;
  if (grid < 1) grid = 1;
  BR_LAUNCH(ctx, halo_send_kernel, grid, 256, 0, st, sa);
  BR_LAUNCH(ctx, halo_recv_kernel, grid, 256, 0, st, ra);
  return BAOREC_OK;
}

static PeerTab recv_tab(const baorec_ctx* ctx, int which) {
  PeerTab t;
  for (int i = 0; i < 16; i++) t.p[i] = ctx->peer_recv[which][i];
  return t;
}

static bool push_with_sms(const baorec_ctx* ctx, size_t kplane, size_t cplane) {
  const int o = ctx->opt_push_sm;
  const bool aligned = (kplane % 2 == 0) && (cplane % 2 == 0);   // 16-byte rows
  return aligned && ctx->nranks > 1 && (o > 0 || (o < 0 && ctx->nranks >= 5));
}

// the word of every peer's block that belongs to this rank
static FlagTab flag_tab(const baorec_ctx* ctx, int which) {
  FlagTab t;
  for (int r = 0; r < 16; r++) {
    PeerFlags* f = (PeerFlags*)ctx->peer_flags[r];
    t.p[r] = f ? f->w[which] + ctx->rank : nullptr;
  }
  return t;
}

static int peer_signal(baorec_ctx* ctx, int which, unsigned value, cudaStream_t st) {
  BR_LAUNCH(ctx, peer_signal_kernel, 1, 32, 0, st, flag_tab(ctx, which), value, ctx->nranks);
  return BAOREC_OK;
}

static int peer_wait(baorec_ctx* ctx, int which, unsigned value, cudaStream_t st) {
  if (value == 0) return BAOREC_OK;
  PeerFlags* f = own_flags(ctx);
  BR_LAUNCH(ctx, peer_wait_kernel, 1, 32, 0, st, f->w[which], value, ctx->nranks, &f->err);
  return BAOREC_OK;
}

struct DistBufs {
  float2 *A, *S, *R, *T;
};

static int dist_bufs(baorec_ctx* ctx, DistBufs* b) {
  const size_t slab_c = (size_t)ctx->nz_loc * ctx->ny * ctx->xh;  // == ny_loc * xh * nz
  BR_TRY(need_t(ctx, BUF_CK0, slab_c, &b->A));
  BR_TRY(need_t(ctx, BUF_CK1, slab_c, &b->T));
  b->S = nullptr;
  if (!ctx->p2p) BR_TRY(need_t(ctx, BUF_A2A_SEND, slab_c, &b->S));  // the peer copies need no send staging
  BR_TRY(need_t(ctx, BUF_A2A_RECV, slab_c, &b->R));
  return BAOREC_OK;
}

// ---- pipelined variants ------------------------------------------------------------------------------
// chunk c of the all-to-all: planes [c nzc, (c+1) nzc) of every per-peer block (contiguous inside the block)
static int all_to_all_chunk(baorec_ctx* ctx, const float2* S, float2* R, size_t blk, size_t off, size_t cnt,
                            cudaStream_t st) {
  const int P = ctx->nranks;
  if (P == 1) {
    BR_CUDA(cudaMemcpyAsync(R + off, S + off, cnt * sizeof(float2), cudaMemcpyDeviceToDevice, st));
    return BAOREC_OK;
  }
  int pi = prof_begin(ctx, "nccl_all_to_all", st);
  BR_NCCL(ncclGroupStart());
  for (int peer = 0; peer < P; peer++) {
    BR_NCCL(ncclSend(S + (size_t)peer * blk + off, cnt * 2, ncclFloat, peer, comm_of(ctx), st));
    BR_NCCL(ncclRecv(R + (size_t)peer * blk + off, cnt * 2, ncclFloat, peer, comm_of(ctx), st));
  }
  BR_NCCL(ncclGroupEnd());
  prof_end(ctx, pi, st);
  return BAOREC_OK;
}

// number of chunks (0 = not pipelined) and the 2-D plans for nzl / chunks planes
static int chunk_setup(baorec_ctx* ctx, int* chunks_out) {
  *chunks_out = 0;
  int C = ctx->opt_a2a_chunks;
  if (C < 2) return BAOREC_OK;
  if (C > 8) C = 8;
  while (C > 1 && (ctx->nz_loc % C != 0)) C--;
  if (C < 2) return BAOREC_OK;
  const int nzc = ctx->nz_loc / C;
  if (ctx->chunk_planes != nzc) {
    if (ctx->chunk_planes) {
      cufftDestroy(ctx->pc_r2c);
      cufftDestroy(ctx->pc_c2r);
      ctx->chunk_planes = 0;
    }
    size_t w[2] = {0, 0};
    int n2[2] = {ctx->ny, ctx->nx};
    BR_CUFFT(cufftCreate(&ctx->pc_r2c));
    BR_CUFFT(cufftCreate(&ctx->pc_c2r));
    ctx->chunk_planes = nzc;
    BR_CUFFT(cufftSetAutoAllocation(ctx->pc_r2c, 0));
    BR_CUFFT(cufftSetAutoAllocation(ctx->pc_c2r, 0));
    BR_CUFFT(cufftMakePlanMany(ctx->pc_r2c, 2, n2, nullptr, 1, 0, nullptr, 1, 0, CUFFT_R2C, nzc, &w[0]));
    BR_CUFFT(cufftMakePlanMany(ctx->pc_c2r, 2, n2, nullptr, 1, 0, nullptr, 1, 0, CUFFT_C2R, nzc, &w[1]));
    const size_t wmax = w[0] > w[1] ? w[0] : w[1];
    if (wmax > ctx->bufs[BUF_WORK].bytes) {
      set_error("chunk plans need a larger cuFFT work area than the slab plans (%zu > %zu bytes)", wmax,
                ctx->bufs[BUF_WORK].bytes);
      return BAOREC_ERR_CUFFT;
    }
    BR_CUFFT(cufftSetWorkArea(ctx->pc_r2c, ctx->bufs[BUF_WORK].p));
    BR_CUFFT(cufftSetWorkArea(ctx->pc_c2r, ctx->bufs[BUF_WORK].p));
  }
  *chunks_out = C;
  return BAOREC_OK;
}

static int dist_r2c_pipelined(baorec_ctx* ctx, const float* slab, float2* T, int C, cudaStream_t st) {
  DistBufs b;
  BR_TRY(dist_bufs(ctx, &b));
  const int P = ctx->nranks, nzl = ctx->nz_loc, nyl = ctx->ny_loc, ny = ctx->ny, xh = ctx->xh, nz = ctx->nz;
  const int nzc = nzl / C;
  const size_t rplane = (size_t)ny * ctx->nx, cplane = (size_t)ny * xh;
  const size_t blk = (size_t)nzl * nyl * xh, cblk = (size_t)nzc * nyl * xh;
  cudaStream_t cs = ctx->comm_stream;
  BR_CUFFT(cufftSetStream(ctx->pc_r2c, st));
  for (int c = 0; c < C; c++) {
    int pi = prof_begin(ctx, "cufft_2d_r2c", st);
    BR_CUFFT(cufftExecR2C(ctx->pc_r2c, (cufftReal*)(slab + (size_t)c * nzc * rplane),
                          (cufftComplex*)(b.A + (size_t)c * nzc * cplane)));
    prof_end(ctx, pi, st);
    BR_LAUNCH(ctx, rows_kernel<true>, (unsigned)(nzc * ny), 128, 0, st, b.S, b.A, nzl, ny, nyl, xh,
              (unsigned)(c * nzc * ny));
    BR_CUDA(cudaEventRecord(ctx->ev_chunk[c], st));
    BR_CUDA(cudaStreamWaitEvent(cs, ctx->ev_chunk[c], 0));
    BR_TRY(all_to_all_chunk(ctx, b.S, b.R, blk, (size_t)c * cblk, cblk, cs));
    BR_CUDA(cudaEventRecord(ctx->ev_a2a[c], cs));
  }
  ctx->n_fft++;
  for (int c = 0; c < C; c++) {
    BR_CUDA(cudaStreamWaitEvent(st, ctx->ev_a2a[c], 0));
    // T[yl][x][s*nzl + zl] = R[s][zl][yl][x] for the planes zl of chunk c
    TrGeom t;
    t.rows = nzc;
    t.cols = xh;
    t.src_row_stride = (size_t)nyl * xh;
    t.dst_row_stride = nz;
    t.nb1 = nyl;
    t.src_b0 = blk;
    t.src_b1 = xh;
    t.dst_b0 = nzl;
    t.dst_b1 = (size_t)xh * nz;
    dim3 grid(cdiv(xh, 32), cdiv(nzc, 32), P * nyl);
    BR_LAUNCH(ctx, transpose_kernel, grid, 256, 0, st, T + (size_t)c * nzc, b.R + (size_t)c * cblk, t, PeerTab{}, 0,
              (size_t)0);
  }
  BR_CUFFT(cufftSetStream(ctx->p1d, st));
  int pi = prof_begin(ctx, "cufft_1d_z", st);
  BR_CUFFT(cufftExecC2C(ctx->p1d, (cufftComplex*)T, (cufftComplex*)T, CUFFT_FORWARD));
  prof_end(ctx, pi, st);
  ctx->n_fft++;
  return BAOREC_OK;
}

static int dist_c2r_pipelined(baorec_ctx* ctx, float2* T, float* slab, int C, cudaStream_t st) {
  DistBufs b;
  BR_TRY(dist_bufs(ctx, &b));
  const int P = ctx->nranks, nzl = ctx->nz_loc, nyl = ctx->ny_loc, ny = ctx->ny, xh = ctx->xh, nz = ctx->nz;
  const int nzc = nzl / C;
  const size_t rplane = (size_t)ny * ctx->nx, cplane = (size_t)ny * xh;
  const size_t blk = (size_t)nzl * nyl * xh, cblk = (size_t)nzc * nyl * xh;
  cudaStream_t cs = ctx->comm_stream;
  BR_CUFFT(cufftSetStream(ctx->p1d, st));
  int pi = prof_begin(ctx, "cufft_1d_z", st);
  BR_CUFFT(cufftExecC2C(ctx->p1d, (cufftComplex*)T, (cufftComplex*)T, CUFFT_INVERSE));
  prof_end(ctx, pi, st);
  ctx->n_fft++;
  for (int c = 0; c < C; c++) {
    // S[d][zl][yl][x] = T[yl][x][d*nzl + zl] for the planes zl of chunk c
    TrGeom t;
    t.rows = xh;
    t.cols = nzc;
    t.src_row_stride = nz;
    t.dst_row_stride = (size_t)nyl * xh;
    t.nb1 = nyl;
    t.src_b0 = nzl;
    t.src_b1 = (size_t)xh * nz;
    t.dst_b0 = blk;
    t.dst_b1 = xh;
    dim3 grid(cdiv(nzc, 32), cdiv(xh, 32), P * nyl);
    BR_LAUNCH(ctx, transpose_kernel, grid, 256, 0, st, b.S + (size_t)c * cblk, T + (size_t)c * nzc, t, PeerTab{}, 0,
              (size_t)0);
    BR_CUDA(cudaEventRecord(ctx->ev_chunk[c], st));
    BR_CUDA(cudaStreamWaitEvent(cs, ctx->ev_chunk[c], 0));
    BR_TRY(all_to_all_chunk(ctx, b.S, b.R, blk, (size_t)c * cblk, cblk, cs));
    BR_CUDA(cudaEventRecord(ctx->ev_a2a[c], cs));
  }
  BR_CUFFT(cufftSetStream(ctx->pc_c2r, st));
  for (int c = 0; c < C; c++) {
    BR_CUDA(cudaStreamWaitEvent(st, ctx->ev_a2a[c], 0));
    BR_LAUNCH(ctx, rows_kernel<false>, (unsigned)(nzc * ny), 128, 0, st, b.A, b.R, nzl, ny, nyl, xh,
              (unsigned)(c * nzc * ny));
    pi = prof_begin(ctx, "cufft_2d_c2r", st);
    BR_CUFFT(cufftExecC2R(ctx->pc_c2r, (cufftComplex*)(b.A + (size_t)c * nzc * cplane),
                          (cufftReal*)(slab + (size_t)c * nzc * rplane)));
    prof_end(ctx, pi, st);
  }
  ctx->n_fft++;
  return BAOREC_OK;
}

// ---- peer-copy transforms (dist_exchange = 1) ---------------------------------------------------------------------
// Forward: slab[zl][y][x] -> K[z][yl][x].  The block of plane zl that belongs to rank d, A[zl][d nyl .. (d+1) nyl)[:],
// is contiguous (nyl * xh values) and its home in rank d's K buffer, RK_d[(rank nzl + zl)][:][:], is contiguous as
// well, so the whole exchange with one peer is ONE strided copy (height = planes of the chunk) issued to the copy
// engines on the communication stream; the 2-D transforms of the next chunk run meanwhile.  Then a flag per chunk.
static int dist_r2c_peer(baorec_ctx* ctx, const float* slab, float2* K, int C, cudaStream_t st) {
  DistBufs b;
  BR_TRY(dist_bufs(ctx, &b));
  const int P = ctx->nranks, nzl = ctx->nz_loc, nyl = ctx->ny_loc, ny = ctx->ny, xh = ctx->xh;
  if (C < 1) C = 1;
  const int nzc = nzl / C;
  const size_t rplane = (size_t)ny * ctx->nx, cplane = (size_t)ny * xh, kplane = (size_t)nyl * xh;
  cudaStream_t cs = ctx->comm_stream, cs2 = ctx->comm_stream2, cs3 = ctx->comm_stream3, cs4 = ctx->comm_stream4;
  const bool split = ctx->opt_comm_split && P > 2;  // two streams of remote copies: two copy engines, two peers at a time
  const bool sm_push = push_with_sms(ctx, kplane, cplane);
  const unsigned seq = ++ctx->seq_k, base = (seq - 1) * (unsigned)C;
  cufftHandle plan = C > 1 ? ctx->pc_r2c : ctx->p2d_r2c;
  BR_CUFFT(cufftSetStream(plan, st));
  // every destination must have finished reading its K buffer of the previous forward transform
  BR_CUDA(cudaEventRecord(ctx->ev_a2a[0], st));   // (orders cs after whatever the caller queued: A is about to be overwritten
  BR_CUDA(cudaStreamWaitEvent(cs, ctx->ev_a2a[0], 0));  //  only by st itself, but RK_d is written from cs)
  BR_TRY(peer_wait(ctx, F_FREE_K, seq - 1, cs));
  BR_CUDA(cudaEventRecord(ctx->ev_a2a[2], cs));
  BR_CUDA(cudaStreamWaitEvent(cs2, ctx->ev_a2a[2], 0));  // (my own K buffer included: the local copies wait too)
  if (split) {
    BR_CUDA(cudaStreamWaitEvent(cs3, ctx->ev_a2a[2], 0));
    BR_CUDA(cudaStreamWaitEvent(cs4, ctx->ev_a2a[2], 0));
  }
  for (int c = 0; c < C; c++) {
    int pi = prof_begin(ctx, "cufft_2d_r2c", st);
    BR_CUFFT(cufftExecR2C(plan, (cufftReal*)(slab + (size_t)c * nzc * rplane), (cufftComplex*)(b.A + (size_t)c * nzc * cplane)));
    prof_end(ctx, pi, st);
    BR_CUDA(cudaEventRecord(ctx->ev_chunk[c], st));
    BR_CUDA(cudaStreamWaitEvent(cs, ctx->ev_chunk[c], 0));
    // this rank's own block is an HBM copy: second stream, beside the NVLink copies
    BR_CUDA(cudaStreamWaitEvent(cs2, ctx->ev_chunk[c], 0));
    if (split) {
      BR_CUDA(cudaStreamWaitEvent(cs3, ctx->ev_chunk[c], 0));
      BR_CUDA(cudaStreamWaitEvent(cs4, ctx->ev_chunk[c], 0));
    }
    if (sm_push) {
      // own block: copy engine; all remote blocks: one kernel
      BR_CUDA(cudaMemcpy2DAsync(ctx->peer_recv[0][ctx->rank] + ((size_t)ctx->rank * nzl + (size_t)c * nzc) * kplane, kplane * sizeof(float2),
                                b.A + (size_t)c * nzc * cplane + (size_t)ctx->rank * kplane, cplane * sizeof(float2),
                                kplane * sizeof(float2), (size_t)nzc, cudaMemcpyDefault, cs2));
      PushGeom pg{kplane, cplane, kplane, ((size_t)ctx->rank * nzl + (size_t)c * nzc) * kplane, (unsigned)(kplane / 2), nzc, P, ctx->rank};
      BR_LAUNCH_NAMED(ctx, "peer_copies", peer_push_kernel, dim3(PUSH_CTAS, P - 1), 256, 0, cs, recv_tab(ctx, 0),
                      b.A + (size_t)c * nzc * cplane, pg);
    } else {
      for (int k = 0; k < P; k++) {
        const int d = (ctx->rank + k) % P;  // staggered: at step k every rank writes to a different peer
        const float2* src = b.A + (size_t)c * nzc * cplane + (size_t)d * kplane;
        float2* dst = ctx->peer_recv[0][d] + ((size_t)ctx->rank * nzl + (size_t)c * nzc) * kplane;
        if (k == 1) pi = prof_begin(ctx, "peer_copies", cs);
        BR_CUDA(cudaMemcpy2DAsync(dst, kplane * sizeof(float2), src, cplane * sizeof(float2), kplane * sizeof(float2),
                                  (size_t)nzc, cudaMemcpyDefault, k == 0 ? cs2 : (!split ? cs : (k % 3 == 1 ? cs : (k % 3 == 2 ? cs3 : cs4)))));
      }
      if (split) {  // the copies to two thirds of the peers ran on two more streams: join them before the flag
        BR_CUDA(cudaEventRecord(ctx->ev_split[c], cs3));
        BR_CUDA(cudaStreamWaitEvent(cs, ctx->ev_split[c], 0));
        BR_CUDA(cudaEventRecord(ctx->ev_split2[c], cs4));
        BR_CUDA(cudaStreamWaitEvent(cs, ctx->ev_split2[c], 0));
      }
      if (P > 1) prof_end(ctx, pi, cs);
    }
    BR_CUDA(cudaEventRecord(ctx->ev_local[c], cs2));
    BR_CUDA(cudaStreamWaitEvent(cs, ctx->ev_local[c], 0));
    BR_TRY(peer_signal(ctx, F_ARR_K, base + (unsigned)c + 1u, cs));
  }
  BR_CUDA(cudaEventRecord(ctx->ev_a2a[1], cs));
  ctx->n_fft++;
  BR_TRY(peer_wait(ctx, F_ARR_K, base + (unsigned)C, st));  // all chunks of all ranks have landed in my K buffer
  if (own_slab_z_available(ctx)) {
    BR_TRY(own_slab_z(ctx, ctx->own_recv[0], K, +1, st));
  } else {
    BR_CUFFT(cufftSetStream(ctx->p1d_s, st));
    int pi = prof_begin(ctx, "cufft_1d_z", st);
    BR_CUFFT(cufftExecC2C(ctx->p1d_s, (cufftComplex*)ctx->own_recv[0], (cufftComplex*)K, CUFFT_FORWARD));
    prof_end(ctx, pi, st);
  }
  ctx->n_fft++;
  BR_TRY(peer_signal(ctx, F_FREE_K, seq, st));                  // my K receive buffer may be overwritten
  BR_CUDA(cudaStreamWaitEvent(st, ctx->ev_a2a[1], 0));     // A may be reused once my own copies have left it
  return BAOREC_OK;
}

// Inverse: K[z][yl][x] (destroyed) -> slab.  Rank d's planes are the contiguous block K[d nzl .. (d+1) nzl); they go
// into rank d's plane-layout buffer RA_d[zl][rank nyl + yl][x] with one strided copy per (peer, chunk); the 2-D C2R of
// chunk c starts as soon as chunk c has arrived from every rank.  Three phases, so that a caller can start the
// exchange of a second field (buffer slot 1) while the first is in flight:
//   c2r_peer_send  : z transform, then the copies + one flag per chunk (communication streams)
//   c2r_peer_wait  : chunk c of `slot` has arrived from every rank (main stream)
//   c2r_peer_done  : my plane buffer may be overwritten; K may be reused once my own copies have left it
struct C2RPeer {
  int slot, C, nzc;
  unsigned seq, base;
  size_t cplane, rplane;
  float2* RA;  // my plane-layout receive buffer of this slot
};

static int c2r_peer_send(baorec_ctx* ctx, float2* K, int slot, int C, C2RPeer* h, cudaStream_t st) {
  const int P = ctx->nranks, nzl = ctx->nz_loc, nyl = ctx->ny_loc, ny = ctx->ny, xh = ctx->xh;
  if (C < 1) C = 1;
  h->slot = slot;
  h->C = C;
  h->nzc = nzl / C;
  h->rplane = (size_t)ny * ctx->nx;
  h->cplane = (size_t)ny * xh;
  h->RA = ctx->own_recv[1 + slot];
  const size_t cplane = h->cplane, kplane = (size_t)nyl * xh;
  const int nzc = h->nzc;
  cudaStream_t cs = ctx->comm_stream, cs2 = ctx->comm_stream2, cs3 = ctx->comm_stream3, cs4 = ctx->comm_stream4;
  const bool split = ctx->opt_comm_split && P > 2;
  const bool sm_push = push_with_sms(ctx, kplane, cplane);
  h->seq = ++ctx->seq_a[slot];
  h->base = (h->seq - 1) * (unsigned)C;
  const int f_arr = slot ? F_ARR_A1 : F_ARR_A0, f_free = slot ? F_FREE_A1 : F_FREE_A0;
  cudaEvent_t ev_z = ctx->ev_a2a[slot ? 4 : 0], ev_w = ctx->ev_a2a[slot ? 5 : 2];
  int pi = 0;
  if (own_slab_z_available(ctx)) {
    BR_TRY(own_slab_z(ctx, K, K, -1, st));
  } else {
    BR_CUFFT(cufftSetStream(ctx->p1d_s, st));
    pi = prof_begin(ctx, "cufft_1d_z", st);
    BR_CUFFT(cufftExecC2C(ctx->p1d_s, (cufftComplex*)K, (cufftComplex*)K, CUFFT_INVERSE));
    prof_end(ctx, pi, st);
  }
  ctx->n_fft++;
  BR_CUDA(cudaEventRecord(ev_z, st));
  BR_CUDA(cudaStreamWaitEvent(cs, ev_z, 0));
  BR_TRY(peer_wait(ctx, f_free, h->seq - 1, cs));  // every destination has consumed this slot's buffer of the previous inverse
  BR_CUDA(cudaEventRecord(ev_w, cs));
  BR_CUDA(cudaStreamWaitEvent(cs2, ev_w, 0));  // (my own plane buffer included: the local copies wait too)
  if (split) {
    BR_CUDA(cudaStreamWaitEvent(cs3, ev_w, 0));
    BR_CUDA(cudaStreamWaitEvent(cs4, ev_w, 0));
  }
  for (int c = 0; c < C; c++) {
    if (sm_push) {
      BR_CUDA(cudaMemcpy2DAsync(ctx->peer_recv[1 + slot][ctx->rank] + (size_t)c * nzc * cplane + (size_t)ctx->rank * kplane,
                                cplane * sizeof(float2), K + ((size_t)ctx->rank * nzl + (size_t)c * nzc) * kplane,
                                kplane * sizeof(float2), kplane * sizeof(float2), (size_t)nzc, cudaMemcpyDefault, cs2));
      PushGeom pg{(size_t)nzl * kplane, kplane, cplane, (size_t)c * nzc * cplane + (size_t)ctx->rank * kplane, (unsigned)(kplane / 2), nzc, P,
                  ctx->rank};
      BR_LAUNCH_NAMED(ctx, "peer_copies", peer_push_kernel, dim3(PUSH_CTAS, P - 1), 256, 0, cs, recv_tab(ctx, 1 + slot),
                      K + (size_t)c * nzc * kplane, pg);
    } else {
      for (int k = 0; k < P; k++) {
        const int d = (ctx->rank + k) % P;
        const float2* src = K + ((size_t)d * nzl + (size_t)c * nzc) * kplane;
        float2* dst = ctx->peer_recv[1 + slot][d] + (size_t)c * nzc * cplane + (size_t)ctx->rank * kplane;
        if (k == 1) pi = prof_begin(ctx, "peer_copies", cs);
        BR_CUDA(cudaMemcpy2DAsync(dst, cplane * sizeof(float2), src, kplane * sizeof(float2), kplane * sizeof(float2),
                                  (size_t)nzc, cudaMemcpyDefault, k == 0 ? cs2 : (!split ? cs : (k % 3 == 1 ? cs : (k % 3 == 2 ? cs3 : cs4)))));
      }
      if (split) {
        BR_CUDA(cudaEventRecord(ctx->ev_split[c], cs3));
        BR_CUDA(cudaStreamWaitEvent(cs, ctx->ev_split[c], 0));
        BR_CUDA(cudaEventRecord(ctx->ev_split2[c], cs4));
        BR_CUDA(cudaStreamWaitEvent(cs, ctx->ev_split2[c], 0));
      }
      if (P > 1) prof_end(ctx, pi, cs);
    }
    BR_CUDA(cudaEventRecord(ctx->ev_local[c], cs2));
    BR_CUDA(cudaStreamWaitEvent(cs, ctx->ev_local[c], 0));
    BR_TRY(peer_signal(ctx, f_arr, h->base + (unsigned)c + 1u, cs));
  }
  BR_CUDA(cudaEventRecord(ctx->ev_a2a[slot ? 3 : 1], cs));
  return BAOREC_OK;
}

static int c2r_peer_wait(baorec_ctx* ctx, const C2RPeer& h, int c, cudaStream_t st) {
  return peer_wait(ctx, h.slot ? F_ARR_A1 : F_ARR_A0, h.base + (unsigned)c + 1u, st);
}

static int c2r_peer_done(baorec_ctx* ctx, const C2RPeer& h, cudaStream_t st) {
  BR_TRY(peer_signal(ctx, h.slot ? F_FREE_A1 : F_FREE_A0, h.seq, st));
  BR_CUDA(cudaStreamWaitEvent(st, ctx->ev_a2a[h.slot ? 3 : 1], 0));
  return BAOREC_OK;
}

static int c2r_2d_chunk(baorec_ctx* ctx, int C, float2* in_planes, float* slab, int c, int nzc, size_t cplane, size_t rplane,
                        cudaStream_t st) {
  cufftHandle plan = C > 1 ? ctx->pc_c2r : ctx->p2d_c2r;
  BR_CUFFT(cufftSetStream(plan, st));
  int pi = prof_begin(ctx, "cufft_2d_c2r", st);
  BR_CUFFT(cufftExecC2R(plan, (cufftComplex*)(in_planes + (size_t)c * nzc * cplane), (cufftReal*)(slab + (size_t)c * nzc * rplane)));
  prof_end(ctx, pi, st);
  return BAOREC_OK;
}

static int dist_c2r_peer(baorec_ctx* ctx, float2* K, float* slab, int C, cudaStream_t st) {
  DistBufs b;
  BR_TRY(dist_bufs(ctx, &b));
  C2RPeer h;
  BR_TRY(c2r_peer_send(ctx, K, 0, C, &h, st));
  for (int c = 0; c < h.C; c++) {
    BR_TRY(c2r_peer_wait(ctx, h, c, st));
    BR_TRY(c2r_2d_chunk(ctx, h.C, h.RA, slab, c, h.nzc, h.cplane, h.rplane, st));
  }
  ctx->n_fft++;
  return c2r_peer_done(ctx, h, st);
}

// i k_x G and i k_y G of a chunk of planes G[zl][y][x] (plane layout, after the z transform and the exchange):
// X overwrites G, Y goes to `y_out`.  The multiplications commute with the z transform, so the x and y displacement
// fields share ONE inverse z transform and ONE exchange (DispGHOp below makes G and the z field H).
__global__ void __launch_bounds__(256)
disp_xy_planes_kernel(float2* __restrict__ g_io, float2* __restrict__ y_out, const float* __restrict__ kx,
                      const float* __restrict__ ky, int xh, int ny, size_t n) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const unsigned row = (unsigned)(i / (unsigned)xh);
    const unsigned ix = (unsigned)(i - (size_t)row * xh), iy = row % (unsigned)ny;
    const float2 v = g_io[i];
    const float kxv = __ldg(kx + ix), kyv = __ldg(ky + iy);
    g_io[i] = make_float2(__fmul_rn(-v.y, kxv), __fmul_rn(v.x, kxv));
    y_out[i] = make_float2(__fmul_rn(-v.y, kyv), __fmul_rn(v.x, kyv));
  }
}

// The three displacement slabs from the kept delta_k (phi_k): TWO inverse z transforms and TWO exchanges (G for x and
// y, H for z) instead of three, software-pipelined over the two plane buffers.  psi[c] point at plane 1 of the
// (nzl + 3)-plane halo buffers.
static int dist_displacements_peer(baorec_ctx* ctx, const float2* keep, bool potential, float* psi[3], cudaStream_t st) {
  DistBufs b;
  BR_TRY(dist_bufs(ctx, &b));
  const size_t slab_c = (size_t)ctx->nz_loc * ctx->ny * ctx->xh;
  float2* KH;
  BR_TRY(need_t(ctx, BUF_CK2, slab_c, &KH));
  float2* KG = b.T;
  int C = 0;
  BR_TRY(chunk_setup(ctx, &C));
  BR_TRY(kpass_disp_gh(ctx, keep, KG, KH, potential, st));
  C2RPeer hh, hg;
  BR_TRY(c2r_peer_send(ctx, KH, 0, C, &hh, st));
  BR_TRY(c2r_peer_send(ctx, KG, 1, C, &hg, st));   // its z transform runs while H is on the wire
  for (int c = 0; c < hh.C; c++) {
    BR_TRY(c2r_peer_wait(ctx, hh, c, st));
    BR_TRY(c2r_2d_chunk(ctx, hh.C, hh.RA, psi[2], c, hh.nzc, hh.cplane, hh.rplane, st));
  }
  ctx->n_fft++;
  BR_TRY(c2r_peer_done(ctx, hh, st));
  for (int c = 0; c < hg.C; c++) {
    BR_TRY(c2r_peer_wait(ctx, hg, c, st));
    const size_t cnt = (size_t)hg.nzc * hg.cplane;
    float2* gx = hg.RA + (size_t)c * cnt;
    float2* gy = b.A + (size_t)c * cnt;
    BR_LAUNCH(ctx, disp_xy_planes_kernel, 148 * 8, 256, 0, st, gx, gy, ctx->d_k[0], ctx->d_k[1], ctx->xh, ctx->ny, cnt);
    BR_TRY(c2r_2d_chunk(ctx, hg.C, hg.RA, psi[0], c, hg.nzc, hg.cplane, hg.rplane, st));
    BR_TRY(c2r_2d_chunk(ctx, hg.C, b.A, psi[1], c, hg.nzc, hg.cplane, hg.rplane, st));
  }
  ctx->n_fft += 2;
  return c2r_peer_done(ctx, hg, st);
}

// slab[nz_loc][ny][nx] (real) -> T[ny_loc][xh][nz] (unnormalised forward transform)
static int dist_r2c(baorec_ctx* ctx, const float* slab, float2* T, cudaStream_t st) {
  int chunks = 0;
  BR_TRY(chunk_setup(ctx, &chunks));
  if (ctx->p2p) return dist_r2c_peer(ctx, slab, T, chunks, st);
  if (chunks >= 2) return dist_r2c_pipelined(ctx, slab, T, chunks, st);
  DistBufs b;
  BR_TRY(dist_bufs(ctx, &b));
  const int P = ctx->nranks, nzl = ctx->nz_loc, nyl = ctx->ny_loc, ny = ctx->ny, xh = ctx->xh, nz = ctx->nz;
  BR_CUFFT(cufftSetStream(ctx->p2d_r2c, st));
  int pi = prof_begin(ctx, "cufft_2d_r2c", st);
  BR_CUFFT(cufftExecR2C(ctx->p2d_r2c, (cufftReal*)slab, (cufftComplex*)b.A));
  prof_end(ctx, pi, st);
  ctx->n_fft++;
  const size_t blk = (size_t)nzl * nyl * xh;
  const float2* Rsrc = b.R;
  BR_LAUNCH(ctx, rows_kernel<true>, (unsigned)(nzl * ny), 128, 0, st, b.S, b.A, nzl, ny, nyl, xh, 0u);
  BR_TRY(all_to_all(ctx, b.S, b.R, blk, st));
  // T[yl][x][s*nzl + zl] = R[s][zl][yl][x]
  TrGeom t;
  t.rows = nzl;
  t.cols = xh;
  t.src_row_stride = (size_t)nyl * xh;
  t.dst_row_stride = nz;
  t.nb1 = nyl;
  t.src_b0 = blk;
  t.src_b1 = xh;
  t.dst_b0 = nzl;
  t.dst_b1 = (size_t)xh * nz;
  dim3 grid(cdiv(xh, 32), cdiv(nzl, 32), P * nyl);
  BR_LAUNCH(ctx, transpose_kernel, grid, 256, 0, st, T, Rsrc, t, PeerTab{}, 0, (size_t)0);
  BR_CUFFT(cufftSetStream(ctx->p1d, st));
  pi = prof_begin(ctx, "cufft_1d_z", st);
  BR_CUFFT(cufftExecC2C(ctx->p1d, (cufftComplex*)T, (cufftComplex*)T, CUFFT_FORWARD));
  prof_end(ctx, pi, st);
  ctx->n_fft++;
  return BAOREC_OK;
}

// T[ny_loc][xh][nz] (destroyed) -> slab[nz_loc][ny][nx]; unnormalised inverse
static int dist_c2r(baorec_ctx* ctx, float2* T, float* slab, cudaStream_t st) {
  int chunks = 0;
  BR_TRY(chunk_setup(ctx, &chunks));
  if (ctx->p2p) return dist_c2r_peer(ctx, T, slab, chunks, st);
  if (chunks >= 2) return dist_c2r_pipelined(ctx, T, slab, chunks, st);
  DistBufs b;
  BR_TRY(dist_bufs(ctx, &b));
  const int P = ctx->nranks, nzl = ctx->nz_loc, nyl = ctx->ny_loc, ny = ctx->ny, xh = ctx->xh, nz = ctx->nz;
  BR_CUFFT(cufftSetStream(ctx->p1d, st));
  int pi = prof_begin(ctx, "cufft_1d_z", st);
  BR_CUFFT(cufftExecC2C(ctx->p1d, (cufftComplex*)T, (cufftComplex*)T, CUFFT_INVERSE));
  prof_end(ctx, pi, st);
  ctx->n_fft++;
  const size_t blk = (size_t)nzl * nyl * xh;
  // S[d][zl][yl][x] = T[yl][x][d*nzl + zl]
  TrGeom t;
  t.rows = xh;
  t.cols = nzl;
  t.src_row_stride = nz;
  t.dst_row_stride = (size_t)nyl * xh;
  t.nb1 = nyl;
  t.src_b0 = nzl;
  t.src_b1 = (size_t)xh * nz;
  t.dst_b0 = blk;
  t.dst_b1 = xh;
  dim3 grid(cdiv(nzl, 32), cdiv(xh, 32), P * nyl);
  const float2* Rsrc = b.R;
  BR_LAUNCH(ctx, transpose_kernel, grid, 256, 0, st, b.S, T, t, PeerTab{}, 0, (size_t)0);
  BR_TRY(all_to_all(ctx, b.S, b.R, blk, st));
  BR_LAUNCH(ctx, rows_kernel<false>, (unsigned)(nzl * ny), 128, 0, st, b.A, Rsrc, nzl, ny, nyl, xh, 0u);
  BR_CUFFT(cufftSetStream(ctx->p2d_c2r, st));
  pi = prof_begin(ctx, "cufft_2d_c2r", st);
  BR_CUFFT(cufftExecC2R(ctx->p2d_c2r, (cufftComplex*)b.A, (cufftReal*)slab));
  prof_end(ctx, pi, st);
  ctx->n_fft++;
  return BAOREC_OK;
}

// peer copies are used when the option asks for them AND every rank's buffers and flags are mapped
void dist_refresh_mode(baorec_ctx* ctx) {
  bool ok = ctx->opt_dist_exchange != 0 && ctx->d_flags != nullptr;
  for (int r = 0; r < ctx->nranks && ok; r++)
    ok = ctx->peer_recv[0][r] && ctx->peer_recv[1][r] && ctx->peer_recv[2][r] && ctx->peer_flags[r];
  if (ok != ctx->p2p) {  // the k-space layout changes with the scheme: cached k-space meshes are void
    ctx->kcache_valid = false;
    ctx->disp_valid = false;
  }
  ctx->p2p = ok;
}

}  // namespace baorec

#define BR_NEED_DIST(ctx)                                                     \
  do {                                                                        \
    BR_REQUIRE((ctx) != nullptr, "ctx is NULL");                              \
    if (!(ctx)->planned || !(ctx)->dist) {                                    \
      baorec::set_error("call baorec_plan_dist before the *_dist entry points"); \
      return BAOREC_ERR_NOT_PLANNED;                                          \
    }                                                                         \
    BR_CUDA(cudaSetDevice((ctx)->device));                                    \
  } while (0)

extern "C" {

static void shard_close_peers(baorec_ctx* ctx, int slot);

int baorec_comm_unique_id(void* out128) {
  BR_REQUIRE(out128 != nullptr, "out128 is NULL");
  static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is expected to be 128 bytes");
  ncclUniqueId id;
  BR_NCCL(ncclGetUniqueId(&id));
  memcpy(out128, &id, sizeof(id));
  return BAOREC_OK;
}

int baorec_comm_init(baorec_ctx* ctx, int rank, int nranks, const void* unique_id128) {
  BR_REQUIRE(ctx != nullptr, "ctx is NULL");
  BR_REQUIRE(nranks >= 1 && rank >= 0 && rank < nranks, "rank / nranks");
  BR_CUDA(cudaSetDevice(ctx->device));
  if (ctx->comm) {
    ncclCommDestroy((ncclComm_t)ctx->comm);
    ctx->comm = nullptr;
  }
  ctx->rank = rank;
  ctx->nranks = nranks;
  if (nranks == 1) return BAOREC_OK;  // single rank: no communicator needed
  BR_REQUIRE(unique_id128 != nullptr, "unique id is NULL");
  ncclUniqueId id;
  memcpy(&id, unique_id128, sizeof(id));
  ncclComm_t comm;
  BR_NCCL(ncclCommInitRank(&comm, nranks, id, rank));
  ctx->comm = comm;
  return BAOREC_OK;
}

static void close_ipc(baorec_ctx* ctx) {
  for (int r = 0; r < 16; r++) {
    for (int k = 0; k < 3; k++) {
      if (ctx->peer_recv[k][r] && ctx->peer_recv[k][r] != ctx->own_recv[k]) cudaIpcCloseMemHandle(ctx->peer_recv[k][r]);
      ctx->peer_recv[k][r] = nullptr;
    }
    if (ctx->peer_flags[r] && ctx->peer_flags[r] != ctx->d_flags) cudaIpcCloseMemHandle(ctx->peer_flags[r]);
    ctx->peer_flags[r] = nullptr;
    if (ctx->peer_inbox[r] && r != ctx->rank) cudaIpcCloseMemHandle(ctx->peer_inbox[r]);
    ctx->peer_inbox[r] = nullptr;
  }
  ctx->p2p = false;
  cudaGetLastError();  // a peer that already exited makes the close fail; nothing to do about it
}

int baorec_comm_destroy_internal(baorec_ctx* ctx) {
  if (ctx) {
    close_ipc(ctx);
    shard_close_peers(ctx, 0);
    shard_close_peers(ctx, 1);
    if (ctx->d_flags) cudaFree(ctx->d_flags);
    ctx->d_flags = nullptr;
    if (ctx->d_halo_done) cudaFree(ctx->d_halo_done);
    ctx->d_halo_done = nullptr;
  }
  if (ctx && ctx->comm) {
    ncclCommDestroy((ncclComm_t)ctx->comm);
    ctx->comm = nullptr;
  }
  return BAOREC_OK;
}

int baorec_plan_dist(baorec_ctx* ctx, int nx, int ny, int nz, const float box_size[3], const float box_min[3]) {
  BR_REQUIRE(ctx != nullptr, "ctx is NULL");
  const int P = ctx->nranks;
  BR_REQUIRE(P == 1 || ctx->comm != nullptr, "call baorec_comm_init first");
  BR_REQUIRE(nz % P == 0 && ny % P == 0, "nz and ny must be divisible by the number of ranks");
  BR_REQUIRE(nz / P >= 2, "each slab needs at least two planes");
  int s = plan_common(ctx, nx, ny, nz, box_size, box_min);
  if (s < 0) return s;
  const int nzl = nz / P, nyl = ny / P;
  if (!ctx->have_dist_plans || s == 0 || ctx->nz_loc != nzl) {
    ctx->dlevels.clear();
    if (ctx->chunk_planes) {
      cufftDestroy(ctx->pc_r2c);
      cufftDestroy(ctx->pc_c2r);
      ctx->chunk_planes = 0;
    }
    if (ctx->have_dist_plans) {
      cufftDestroy(ctx->p2d_r2c);
      cufftDestroy(ctx->p2d_c2r);
      cufftDestroy(ctx->p1d);
      cufftDestroy(ctx->p1d_s);
      ctx->have_dist_plans = false;
    }
    size_t w[4] = {0, 0, 0, 0};
    int n2[2] = {ny, nx};
    int n1[1] = {nz};
    BR_CUFFT(cufftCreate(&ctx->p2d_r2c));
    BR_CUFFT(cufftCreate(&ctx->p2d_c2r));
    BR_CUFFT(cufftCreate(&ctx->p1d));
    BR_CUFFT(cufftCreate(&ctx->p1d_s));
    ctx->have_dist_plans = true;
    BR_CUFFT(cufftSetAutoAllocation(ctx->p2d_r2c, 0));
    BR_CUFFT(cufftSetAutoAllocation(ctx->p2d_c2r, 0));
    BR_CUFFT(cufftSetAutoAllocation(ctx->p1d, 0));
    BR_CUFFT(cufftSetAutoAllocation(ctx->p1d_s, 0));
    BR_CUFFT(cufftMakePlanMany(ctx->p2d_r2c, 2, n2, nullptr, 1, 0, nullptr, 1, 0, CUFFT_R2C, nzl, &w[0]));
    BR_CUFFT(cufftMakePlanMany(ctx->p2d_c2r, 2, n2, nullptr, 1, 0, nullptr, 1, 0, CUFFT_C2R, nzl, &w[1]));
    BR_CUFFT(cufftMakePlanMany(ctx->p1d, 1, n1, nullptr, 1, 0, nullptr, 1, 0, CUFFT_C2C, nyl * (nx / 2 + 1), &w[2]));
    {
      // z transform of the K[z][yl][x] layout (peer-copy exchange): nyl * xh interleaved columns, stride = one z plane
      const int kplane = nyl * (nx / 2 + 1);
      int emb[1] = {nz};
      BR_CUFFT(cufftMakePlanMany(ctx->p1d_s, 1, n1, emb, kplane, 1, emb, kplane, 1, CUFFT_C2C, kplane, &w[3]));
    }
    size_t wmax = w[0] > w[1] ? w[0] : w[1];
    if (w[2] > wmax) wmax = w[2];
    if (w[3] > wmax) wmax = w[3];
    if (wmax < 16) wmax = 16;
    if (ctx->bufs[BUF_WORK].bytes > wmax) wmax = ctx->bufs[BUF_WORK].bytes;
    void* work = nullptr;
    BR_TRY(need(ctx, BUF_WORK, wmax, &work));
    BR_CUFFT(cufftSetWorkArea(ctx->p2d_r2c, work));
    BR_CUFFT(cufftSetWorkArea(ctx->p2d_c2r, work));
    BR_CUFFT(cufftSetWorkArea(ctx->p1d, work));
    BR_CUFFT(cufftSetWorkArea(ctx->p1d_s, work));
    if (ctx->have_plans) {  // the single-GPU plans share the (possibly re-allocated) work area
      BR_CUFFT(cufftSetWorkArea(ctx->r2c, work));
      BR_CUFFT(cufftSetWorkArea(ctx->c2r, work));
    }
  }
  BR_TRY(own_fft_setup(ctx));  // twiddle tables of the column kernels (the slab z pass uses them where they beat cuFFT's strided plan)
  {
    // receive buffers are allocated here at their final size: their addresses are exported via IPC
    const size_t slab_c = (size_t)nzl * ny * (nx / 2 + 1);
    void* r0 = ctx->bufs[BUF_A2A_RECV].p;
    void* r1 = ctx->bufs[BUF_A2A_RECV2].p;
    void* r2 = ctx->bufs[BUF_A2A_RECV3].p;
    BR_TRY(need_t(ctx, BUF_A2A_RECV, slab_c, &ctx->own_recv[0]));
    BR_TRY(need_t(ctx, BUF_A2A_RECV2, slab_c, &ctx->own_recv[1]));
    BR_TRY(need_t(ctx, BUF_A2A_RECV3, slab_c, &ctx->own_recv[2]));
    if (r0 != ctx->own_recv[0] || r1 != ctx->own_recv[1] || r2 != ctx->own_recv[2]) {  // re-export needed
      close_ipc(ctx);
    }
    if (!ctx->d_flags) {
      BR_CUDA(cudaMalloc(&ctx->d_flags, sizeof(PeerFlags)));
      BR_CUDA(cudaMemset(ctx->d_flags, 0, sizeof(PeerFlags)));
    }
    if (!ctx->d_halo_done) {
      BR_CUDA(cudaMalloc(&ctx->d_halo_done, 2 * sizeof(unsigned)));
      BR_CUDA(cudaMemset(ctx->d_halo_done, 0, 2 * sizeof(unsigned)));
    }
    {
      // halo / ghost planes travel through an inbox the ring neighbours write into: HALO_SLOTS calls in flight, two
      // transfers of up to two planes each per call
      const size_t slot = 2 * (size_t)ny * nx;
      void* old = ctx->bufs[BUF_HALO_INBOX].p;
      float* inbox;
      BR_TRY(need_t(ctx, BUF_HALO_INBOX, (size_t)HALO_SLOTS * 2 * slot, &inbox));
      if (old != (void*)inbox) close_ipc(ctx);
      ctx->inbox_slot_floats = slot;
      ctx->own_inbox = inbox;
    }
    if (P == 1) {
      // a single rank is its own (only) peer: the peer-copy path runs with local copies and local flags
      for (int k = 0; k < 3; k++) ctx->peer_recv[k][0] = ctx->own_recv[k];
      ctx->peer_flags[0] = ctx->d_flags;
      ctx->peer_inbox[0] = ctx->own_inbox;
    }
    dist_refresh_mode(ctx);
  }
  ctx->nz_loc = nzl;
  ctx->z0 = ctx->rank * nzl;
  ctx->ny_loc = nyl;
  ctx->y0 = ctx->rank * nyl;
  ctx->dist = true;
  ctx->planned = true;
  ctx->kcache_valid = false;
  return BAOREC_OK;
}

int baorec_dist_ipc_close(baorec_ctx* ctx) {
  BR_REQUIRE(ctx != nullptr, "ctx is NULL");
  BR_CUDA(cudaSetDevice(ctx->device));
  BR_CUDA(cudaDeviceSynchronize());
  close_ipc(ctx);
  return BAOREC_OK;
}

int baorec_dist_ipc_export(baorec_ctx* ctx, void* out320) {
  BR_NEED_DIST(ctx);
  BR_REQUIRE(out320 != nullptr, "out320 is NULL");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is expected to be 64 bytes");
  // A new generation of the exchange starts here: every rank calls this before the handles are gathered (a
  // synchronisation point of the caller), so no peer can have written a flag of the new generation yet.
  BR_CUDA(cudaDeviceSynchronize());
  BR_CUDA(cudaMemset(ctx->d_flags, 0, sizeof(PeerFlags)));
  BR_CUDA(cudaDeviceSynchronize());
  ctx->seq_k = ctx->seq_a[0] = ctx->seq_a[1] = 0;
  for (int k = 0; k < 2; k++) ctx->seq_shard[k][0] = ctx->seq_shard[k][1] = 0;
  ctx->seq_halo = 0;
  cudaIpcMemHandle_t h[5];
  for (int k = 0; k < 3; k++) BR_CUDA(cudaIpcGetMemHandle(&h[k], ctx->own_recv[k]));
  BR_CUDA(cudaIpcGetMemHandle(&h[3], ctx->d_flags));
  BR_CUDA(cudaIpcGetMemHandle(&h[4], ctx->own_inbox));
  memcpy(out320, h, sizeof(h));
  return BAOREC_OK;
}

int baorec_dist_ipc_open(baorec_ctx* ctx, const void* all_handles, int nranks) {
  BR_NEED_DIST(ctx);
  BR_REQUIRE(all_handles != nullptr && nranks == ctx->nranks && nranks <= 16, "ipc_open arguments");
  const cudaIpcMemHandle_t* h = (const cudaIpcMemHandle_t*)all_handles;
  close_ipc(ctx);
  for (int r = 0; r < nranks; r++) {
    if (r == ctx->rank) {
      for (int k = 0; k < 3; k++) ctx->peer_recv[k][r] = ctx->own_recv[k];
      ctx->peer_flags[r] = ctx->d_flags;
      ctx->peer_inbox[r] = ctx->own_inbox;
      continue;
    }
    void* p[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
    for (int k = 0; k < 5; k++) BR_CUDA(cudaIpcOpenMemHandle(&p[k], h[5 * r + k], cudaIpcMemLazyEnablePeerAccess));
    for (int k = 0; k < 3; k++) ctx->peer_recv[k][r] = (float2*)p[k];
    ctx->peer_flags[r] = p[3];
    ctx->peer_inbox[r] = p[4];
  }
  dist_refresh_mode(ctx);
  return BAOREC_OK;
}

// the wait kernels raise PeerFlags::err instead of spinning forever when a peer never signals
static int peer_check(baorec_ctx* ctx, cudaStream_t st) {
  if (!ctx->p2p || !ctx->d_flags) return BAOREC_OK;
  unsigned e = 0;
  BR_CUDA(cudaMemcpyAsync(&e, &own_flags(ctx)->err, sizeof(e), cudaMemcpyDeviceToHost, st));
  BR_CUDA(cudaStreamSynchronize(st));
  if (e) {
    set_error("peer-copy exchange: a rank never signalled (20 s timeout); results are invalid");
    return BAOREC_ERR_NCCL;
  }
  return BAOREC_OK;
}

int baorec_dist_exchange_mode(const baorec_ctx* ctx) { return ctx != nullptr && ctx->p2p ? 1 : 0; }

int baorec_slab_range(const baorec_ctx* ctx, int* z_lo, int* nz_loc) {
  BR_REQUIRE(ctx != nullptr && ctx->dist, "not a distributed plan");
  if (z_lo) *z_lo = ctx->z0;
  if (nz_loc) *nz_loc = ctx->nz_loc;
  return BAOREC_OK;
}

int baorec_slab_owner_f32(baorec_ctx* ctx, const float* d_z, int64_t n, int32_t* d_owner, baorec_stream stream) {
  BR_NEED_DIST(ctx);
  BR_REQUIRE(n >= 0 && (n == 0 || (d_z && d_owner)), "slab_owner arguments");
  if (n == 0) return BAOREC_OK;
  // note: the wrap of cic! uses the axis-1 box for every axis (src/mas.jl:8-10)
  BR_LAUNCH(ctx, slab_owner_kernel, cdiv((size_t)n, 256), 256, 0, (cudaStream_t)stream, d_z, n, ctx->mn[2], ctx->L[2],
            ctx->nz, ctx->nz_loc, d_owner);
  return BAOREC_OK;
}

// ---- particle sharding by slab ------------------------------------------------------------------------------------
// "particle redistribution by slab is one bucketed all-to-all(v)".  Reference precedent:
// send_to_relevant_process, src/mas.jl:112-140 (unfinished there).  Every rank passes ANY part of the catalog; the
// owner of a particle is the slab of its cic! base plane (slab_owner_kernel's arithmetic).  Device counting sort by
// owner (block-local ranking in shared memory, one global reservation per block and owner), the P x (P+1) count
// matrix all-gathered (the extra column carries the out-of-box counts so that every rank fails together), then one
// grouped ncclSend / ncclRecv per column and peer.  The map of original index -> send position stays on the sender,
// so results computed in slab order travel back the same way and land in the caller's order (baorec_unshard_f32).
constexpr int SHARD_THREADS = 256, SHARD_ITEMS = 8;

__device__ __forceinline__ int slab_owner_of(float p, float mn, float L, float mn0, float L0, int nz, int nzl) {
  if (__fsub_rn(p, mn0) > L0) p = __fsub_rn(p, L0);  // upper-face wrap of cic! (axis-1 box for every axis, src/mas.jl:8-10)
  float g = __fadd_rn(__fdiv_rn(__fmul_rn(__fsub_rn(p, mn), (float)nz), L), 1.0f);
  int c0 = (int)floorf(g);
  if (c0 == nz + 1) c0 = 1;
  return (g >= 1.0f && c0 >= 1 && c0 <= nz) ? (c0 - 1) / nzl : 16;  // 16 = out of box
}

// cnt[0..15] += particles per owner, cnt[16] += out-of-box
__global__ void __launch_bounds__(SHARD_THREADS)
shard_count_kernel(const float* __restrict__ z, int64_t n, float mn, float L, float mn0, float L0, int nz, int nzl,
                   unsigned long long* __restrict__ cnt) {
  __shared__ unsigned h[17];
  if (threadIdx.x < 17) h[threadIdx.x] = 0;
  __syncthreads();
  const int64_t base = (int64_t)blockIdx.x * (SHARD_THREADS * SHARD_ITEMS);
#pragma unroll
  for (int k = 0; k < SHARD_ITEMS; k++) {
    int64_t i = base + k * SHARD_THREADS + threadIdx.x;
    if (i < n) atomicAdd(&h[slab_owner_of(z[i], mn, L, mn0, L0, nz, nzl)], 1u);
  }
  __syncthreads();
  if (threadIdx.x < 17 && h[threadIdx.x]) atomicAdd(cnt + threadIdx.x, (unsigned long long)h[threadIdx.x]);
}

// cursor[o] starts at the offset of owner o in the send order; out-of-box particles are dropped (map = ~0)
__global__ void __launch_bounds__(SHARD_THREADS)
shard_reorder_kernel(const float* __restrict__ x, const float* __restrict__ y, const float* __restrict__ z,
                     const float* __restrict__ w, int64_t n, float mn, float L, float mn0, float L0, int nz, int nzl,
                     unsigned long long* __restrict__ cursor, float* __restrict__ sx, float* __restrict__ sy,
                     float* __restrict__ sz, float* __restrict__ sw, unsigned* __restrict__ map) {
  __shared__ unsigned h[17];
  __shared__ unsigned long long start[17];
  if (threadIdx.x < 17) h[threadIdx.x] = 0;
  __syncthreads();
  const int64_t base = (int64_t)blockIdx.x * (SHARD_THREADS * SHARD_ITEMS);
  int own[SHARD_ITEMS];
  unsigned rank_in[SHARD_ITEMS];
  float pz[SHARD_ITEMS];
#pragma unroll
  for (int k = 0; k < SHARD_ITEMS; k++) {
    int64_t i = base + k * SHARD_THREADS + threadIdx.x;
    own[k] = -1;
    if (i < n) {
      pz[k] = z[i];
      own[k] = slab_owner_of(pz[k], mn, L, mn0, L0, nz, nzl);
      rank_in[k] = atomicAdd(&h[own[k]], 1u);
    }
  }
  __syncthreads();
  if (threadIdx.x < 16 && h[threadIdx.x]) start[threadIdx.x] = atomicAdd(cursor + threadIdx.x, (unsigned long long)h[threadIdx.x]);
  __syncthreads();
#pragma unroll
  for (int k = 0; k < SHARD_ITEMS; k++) {
    int64_t i = base + k * SHARD_THREADS + threadIdx.x;
    if (own[k] < 0) continue;
    if (own[k] == 16) {
      map[i] = 0xffffffffu;
      continue;
    }
    const size_t slot = (size_t)(start[own[k]] + rank_in[k]);
    sx[slot] = x[i];
    sy[slot] = y[i];
    sz[slot] = pz[k];
    sw[slot] = w[i];
    map[i] = (unsigned)slot;
  }
}

__global__ void shard_cursor_kernel(unsigned long long* cnt) {
  // cursors [17 .. 33) = exclusive prefix sum of the per-owner counts
  if (threadIdx.x == 0) {
    unsigned long long s = 0;
    for (int o = 0; o < 16; o++) {
      cnt[17 + o] = s;
      s += cnt[o];
    }
  }
}

// out[i] = back[map[i]] for up to three columns
__global__ void __launch_bounds__(256)
unshard_gather_kernel(const unsigned* __restrict__ map, int64_t n, const float* __restrict__ ba, const float* __restrict__ bb,
                      const float* __restrict__ bc, float* __restrict__ oa, float* __restrict__ ob, float* __restrict__ oc) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const unsigned s = map[i];
  if (s == 0xffffffffu) return;
  if (oa) oa[i] = ba[s];
  if (ob) ob[i] = bb[s];
  if (oc) oc[i] = bc[s];
}

// ---- sharding: buffers that peers write into --------------------------------------------------------------------
// With the peer mappings of the slab transforms in place (ctx->p2p) the catalog columns and the results travel as
// plain copy-engine copies into the peers' buffers too (NCCL's grouped send/recv moved them at ~200 GB/s).  The
// receiving buffers grow with the catalog, so their IPC handles are exchanged INSIDE the library whenever any rank has
// to grow one -- every rank sees the whole count matrix and therefore takes the same decision at the same call.
static void shard_close_peers(baorec_ctx* ctx, int slot) {
  for (int r = 0; r < 16; r++)
    for (int k = 0; k < 2; k++) {
      void*& q = ctx->shard_peer[slot][k][r];
      if (q && r != ctx->rank) cudaIpcCloseMemHandle(q);
      q = nullptr;
    }
  cudaGetLastError();
}

// capacities (particles) every rank's two buffers must have; grows + re-maps when any of them is too small
static int shard_ensure_peer_buffers(baorec_ctx* ctx, int slot, const int64_t need_recv[16], const int64_t need_send[16],
                                     cudaStream_t st) {
  const int P = ctx->nranks;
  bool grow = false;
  for (int r = 0; r < P; r++)
    grow = grow || need_recv[r] > ctx->shard_cap[slot][0][r] || need_send[r] > ctx->shard_cap[slot][1][r] ||
           !ctx->shard_peer[slot][0][r] || !ctx->shard_peer[slot][1][r];
  if (!grow) return BAOREC_OK;
  // (1) nobody frees a buffer a peer still maps: drop the mappings, then a barrier
  BR_CUDA(cudaStreamSynchronize(st));
  shard_close_peers(ctx, slot);
  if (!ctx->d_ipc_stage) BR_CUDA(cudaMalloc(&ctx->d_ipc_stage, 17 * 128));
  if (P > 1) {
    BR_NCCL(ncclAllReduce(ctx->d_ipc_stage, ctx->d_ipc_stage, 1, ncclInt, ncclMax, comm_of(ctx), st));
    BR_CUDA(cudaStreamSynchronize(st));
  }
  // (2) new capacities: 25 % headroom so that a slightly different catalog does not grow again
  for (int r = 0; r < P; r++) {
    const int64_t nr = need_recv[r] + need_recv[r] / 4 + 1024, ns = need_send[r] + need_send[r] / 4 + 1024;
    if (need_recv[r] > ctx->shard_cap[slot][0][r]) ctx->shard_cap[slot][0][r] = nr;
    if (need_send[r] > ctx->shard_cap[slot][1][r]) ctx->shard_cap[slot][1][r] = ns;
  }
  float *recv, *send;
  BR_TRY(need_t(ctx, slot ? BUF_SHARD_RECV2 : BUF_SHARD_RECV, 4 * (size_t)ctx->shard_cap[slot][0][ctx->rank], &recv));
  BR_TRY(need_t(ctx, slot ? BUF_SHARD_SEND2 : BUF_SHARD_SEND, 4 * (size_t)ctx->shard_cap[slot][1][ctx->rank], &send));
  // (3) handles: all-gather through device memory, then open
  cudaIpcMemHandle_t mine[2], all[2 * 16];
  BR_CUDA(cudaIpcGetMemHandle(&mine[0], recv));
  BR_CUDA(cudaIpcGetMemHandle(&mine[1], send));
  char* stage = (char*)ctx->d_ipc_stage;
  BR_CUDA(cudaMemcpyAsync(stage + 128 * 16, mine, 128, cudaMemcpyHostToDevice, st));
  if (P > 1) {
    BR_NCCL(ncclAllGather(stage + 128 * 16, stage, 128, ncclChar, comm_of(ctx), st));
  } else {
    BR_CUDA(cudaMemcpyAsync(stage, stage + 128 * 16, 128, cudaMemcpyDeviceToDevice, st));
  }
  BR_CUDA(cudaMemcpyAsync(all, stage, 128 * (size_t)P, cudaMemcpyDeviceToHost, st));
  BR_CUDA(cudaStreamSynchronize(st));
  for (int r = 0; r < P; r++)
    for (int k = 0; k < 2; k++) {
      if (r == ctx->rank) {
        ctx->shard_peer[slot][k][r] = k == 0 ? (void*)recv : (void*)send;
        continue;
      }
      void* q = nullptr;
      BR_CUDA(cudaIpcOpenMemHandle(&q, all[2 * r + k], cudaIpcMemLazyEnablePeerAccess));
      ctx->shard_peer[slot][k][r] = q;
    }
  return BAOREC_OK;
}

static int shard_catalog(baorec_ctx* ctx, int slot, const float* x, const float* y, const float* z, const float* w,
                         int64_t n, float** ox, float** oy, float** oz, float** ow, int64_t* n_local, cudaStream_t st) {
  const int P = ctx->nranks;
  BR_REQUIRE(P <= 16, "at most 16 ranks");
  BR_REQUIRE(n < ((int64_t)1 << 32) - 1, "at most 2^32 - 2 particles per rank and call");
  baorec_ctx::ShardState& S = ctx->shard[slot];
  S.valid = false;
  const bool peer = ctx->p2p;
  const unsigned seq = peer ? ++ctx->seq_shard[slot][0] : 0;
  // my receive columns of the previous call are dead once everything queued so far has run: peers may overwrite them
  if (peer) BR_TRY(peer_signal(ctx, slot ? F_FREE_S1 : F_FREE_S0, seq, st));
  if (!ctx->d_shard_cnt) BR_CUDA(cudaMalloc(&ctx->d_shard_cnt, sizeof(unsigned long long) * (40 + 17 * 16)));
  unsigned long long* cnt = ctx->d_shard_cnt;
  BR_CUDA(cudaMemsetAsync(cnt, 0, sizeof(unsigned long long) * 40, st));
  const unsigned grid = cdiv((size_t)(n > 0 ? n : 1), SHARD_THREADS * SHARD_ITEMS);
  if (n > 0)
    BR_LAUNCH(ctx, shard_count_kernel, grid, SHARD_THREADS, 0, st, z, n, ctx->mn[2], ctx->L[2], ctx->mn[0], ctx->L[0], ctx->nz,
              ctx->nz_loc, cnt);
  BR_LAUNCH(ctx, shard_cursor_kernel, 1, 32, 0, st, cnt);
  // count matrix: row s = rank s's 17 counts
  unsigned long long* all = cnt + 40;
  if (P > 1) {
    BR_NCCL(ncclAllGather(cnt, all, 17, ncclUint64, comm_of(ctx), st));
  } else {
    BR_CUDA(cudaMemcpyAsync(all, cnt, 17 * sizeof(unsigned long long), cudaMemcpyDeviceToDevice, st));
  }
  unsigned long long h[17 * 16];
  BR_CUDA(cudaMemcpyAsync(h, all, sizeof(unsigned long long) * 17 * P, cudaMemcpyDeviceToHost, st));
  BR_CUDA(cudaStreamSynchronize(st));
  unsigned long long oob = 0;
  for (int s = 0; s < P; s++) oob += h[17 * s + 16];
  if (oob) {  // every rank sees the same matrix: all of them return the error together
    set_error("shard_catalog: %llu particle(s) outside the box over all ranks (the reference would raise BoundsError)", oob);
    return BAOREC_ERR_OUT_OF_BOX;
  }
  int64_t so = 0, ro = 0, need_recv[16] = {}, need_send[16] = {};
  for (int r = 0; r < P; r++) {
    S.send_cnt[r] = (int64_t)h[17 * ctx->rank + r];
    S.send_off[r] = so;
    so += S.send_cnt[r];
    S.recv_cnt[r] = (int64_t)h[17 * r + ctx->rank];
    S.recv_off[r] = ro;
    ro += S.recv_cnt[r];
    S.off_at_dst[r] = S.off_at_src[r] = 0;
    for (int q = 0; q < P; q++) {
      need_recv[r] += (int64_t)h[17 * q + r];
      need_send[r] += (int64_t)h[17 * r + q];
      if (q < ctx->rank) {
        S.off_at_dst[r] += (int64_t)h[17 * q + r];   // where my segment starts in rank r's receive columns
        S.off_at_src[r] += (int64_t)h[17 * r + q];   // where my results start in rank r's send order
      }
    }
  }
  S.n = n;
  S.n_local = ro;
  float *send, *recv;
  unsigned* map;
  size_t ns = (size_t)(n > 0 ? n : 1), nl = (size_t)(ro > 0 ? ro : 1);
  if (peer) {
    BR_TRY(shard_ensure_peer_buffers(ctx, slot, need_recv, need_send, st));
    ns = (size_t)ctx->shard_cap[slot][1][ctx->rank];   // column strides = the capacities every rank knows
    nl = (size_t)ctx->shard_cap[slot][0][ctx->rank];
    send = (float*)ctx->shard_peer[slot][1][ctx->rank];
    recv = (float*)ctx->shard_peer[slot][0][ctx->rank];
  } else {
    BR_TRY(need_t(ctx, slot ? BUF_SHARD_SEND2 : BUF_SHARD_SEND, 4 * ns, &send));
    BR_TRY(need_t(ctx, slot ? BUF_SHARD_RECV2 : BUF_SHARD_RECV, 4 * nl, &recv));
  }
  S.send_stride = (int64_t)ns;
  S.recv_stride = (int64_t)nl;
  BR_TRY(need_t(ctx, slot ? BUF_SHARD_MAP2 : BUF_SHARD_MAP, (size_t)(n > 0 ? n : 1), &map));
  if (n > 0)
    BR_LAUNCH(ctx, shard_reorder_kernel, grid, SHARD_THREADS, 0, st, x, y, z, w, n, ctx->mn[2], ctx->L[2], ctx->mn[0], ctx->L[0],
              ctx->nz, ctx->nz_loc, cnt + 17, send, send + ns, send + 2 * ns, send + 3 * ns, map);
  if (peer) {
    // one copy per (peer, column) straight into the peer's receive columns, after the peer has entered this call
    BR_TRY(peer_wait(ctx, slot ? F_FREE_S1 : F_FREE_S0, seq, st));
    int pi = prof_begin(ctx, "peer_shard_copies", st);
    for (int k = 0; k < P; k++) {
      const int d = (ctx->rank + k) % P;
      if (!S.send_cnt[d]) continue;
      float* dcol = (float*)ctx->shard_peer[slot][0][d];
      const size_t dstride = (size_t)ctx->shard_cap[slot][0][d];
      for (int c = 0; c < 4; c++)
        BR_CUDA(cudaMemcpyAsync(dcol + c * dstride + S.off_at_dst[d], send + c * ns + S.send_off[d],
                                (size_t)S.send_cnt[d] * sizeof(float), cudaMemcpyDefault, st));
    }
    prof_end(ctx, pi, st);
    BR_TRY(peer_signal(ctx, slot ? F_ARR_S1 : F_ARR_S0, seq, st));
    BR_TRY(peer_wait(ctx, slot ? F_ARR_S1 : F_ARR_S0, seq, st));
  } else if (P > 1) {
    int pi = prof_begin(ctx, "nccl_shard_all_to_all_v", st);
    BR_NCCL(ncclGroupStart());
    for (int r = 0; r < P; r++)
      for (int c = 0; c < 4; c++) {
        if (S.send_cnt[r]) BR_NCCL(ncclSend(send + c * ns + S.send_off[r], (size_t)S.send_cnt[r], ncclFloat, r, comm_of(ctx), st));
        if (S.recv_cnt[r]) BR_NCCL(ncclRecv(recv + c * nl + S.recv_off[r], (size_t)S.recv_cnt[r], ncclFloat, r, comm_of(ctx), st));
      }
    BR_NCCL(ncclGroupEnd());
    prof_end(ctx, pi, st);
  } else {
    for (int c = 0; c < 4 && n > 0; c++)
      BR_CUDA(cudaMemcpyAsync(recv + c * nl, send + c * ns, (size_t)n * sizeof(float), cudaMemcpyDeviceToDevice, st));
  }
  *ox = recv;
  *oy = recv + nl;
  *oz = recv + 2 * nl;
  *ow = recv + 3 * nl;
  *n_local = ro;
  S.peer = peer;
  S.valid = true;
  return BAOREC_OK;
}

// columns a, b, c (n_local values each, slab order) -> oa, ob, oc (n values each, the order of the arrays that were sharded)
static int unshard(baorec_ctx* ctx, int slot, const float* a, const float* b, const float* c, float* oa, float* ob, float* oc,
                   cudaStream_t st) {
  const int P = ctx->nranks;
  baorec_ctx::ShardState& S = ctx->shard[slot];
  if (!S.valid) {
    set_error("baorec_unshard_f32: no routing table for slot %d (call baorec_shard_catalog_f32 first)", slot);
    return BAOREC_ERR_INVALID;
  }
  const size_t ns = (size_t)S.send_stride;
  float* back = (float*)ctx->bufs[slot ? BUF_SHARD_SEND2 : BUF_SHARD_SEND].p;  // the send staging is free again: 4 columns
  const unsigned* map = (const unsigned*)ctx->bufs[slot ? BUF_SHARD_MAP2 : BUF_SHARD_MAP].p;
  const float* src[3] = {a, b, c};
  if (S.peer && ctx->p2p) {
    const unsigned seq = ++ctx->seq_shard[slot][1];
    const int f_free = slot ? F_FREE_B1 : F_FREE_B0, f_arr = slot ? F_ARR_B1 : F_ARR_B0;
    BR_TRY(peer_signal(ctx, f_free, seq, st));  // my shard copies have left the staging (stream order): results may land in it
    BR_TRY(peer_wait(ctx, f_free, seq, st));
    int pi = prof_begin(ctx, "peer_unshard_copies", st);
    for (int k = 0; k < P; k++) {
      const int s = (ctx->rank + k) % P;
      if (!S.recv_cnt[s]) continue;
      float* dcol = (float*)ctx->shard_peer[slot][1][s];
      const size_t dstride = (size_t)ctx->shard_cap[slot][1][s];
      for (int q = 0; q < 3; q++)
        if (src[q])
          BR_CUDA(cudaMemcpyAsync(dcol + q * dstride + S.off_at_src[s], src[q] + S.recv_off[s],
                                  (size_t)S.recv_cnt[s] * sizeof(float), cudaMemcpyDefault, st));
    }
    prof_end(ctx, pi, st);
    BR_TRY(peer_signal(ctx, f_arr, seq, st));
    BR_TRY(peer_wait(ctx, f_arr, seq, st));
  } else if (P > 1) {
    int pi = prof_begin(ctx, "nccl_unshard_all_to_all_v", st);
    BR_NCCL(ncclGroupStart());
    for (int r = 0; r < P; r++)
      for (int k = 0; k < 3; k++) {
        if (!src[k]) continue;
        if (S.recv_cnt[r]) BR_NCCL(ncclSend(src[k] + S.recv_off[r], (size_t)S.recv_cnt[r], ncclFloat, r, comm_of(ctx), st));
        if (S.send_cnt[r]) BR_NCCL(ncclRecv(back + k * ns + S.send_off[r], (size_t)S.send_cnt[r], ncclFloat, r, comm_of(ctx), st));
      }
    BR_NCCL(ncclGroupEnd());
    prof_end(ctx, pi, st);
  } else {
    for (int k = 0; k < 3; k++)
      if (src[k] && S.n > 0)
        BR_CUDA(cudaMemcpyAsync(back + k * ns, src[k], (size_t)S.n * sizeof(float), cudaMemcpyDeviceToDevice, st));
  }
  if (S.n > 0)
    BR_LAUNCH(ctx, unshard_gather_kernel, cdiv((size_t)S.n, 256), 256, 0, st, map, S.n, back, back + ns, back + 2 * ns,
              a ? oa : nullptr, b ? ob : nullptr, c ? oc : nullptr);
  return BAOREC_OK;
}

int baorec_shard_catalog_f32(baorec_ctx* ctx, int slot, const float* d_x, const float* d_y, const float* d_z,
                             const float* d_w, int64_t n, float** d_sx, float** d_sy, float** d_sz, float** d_sw,
                             int64_t* n_local, baorec_stream stream) {
  BR_NEED_DIST(ctx);
  BR_REQUIRE(slot == 0 || slot == 1, "slot is 0 (data) or 1 (randoms)");
  BR_REQUIRE(n >= 0 && (n == 0 || (d_x && d_y && d_z && d_w)), "particle arrays");
  BR_REQUIRE(d_sx && d_sy && d_sz && d_sw && n_local, "NULL output argument");
  return shard_catalog(ctx, slot, d_x, d_y, d_z, d_w, n, d_sx, d_sy, d_sz, d_sw, n_local, (cudaStream_t)stream);
}

int baorec_unshard_f32(baorec_ctx* ctx, int slot, const float* d_a, const float* d_b, const float* d_c, float* d_oa,
                       float* d_ob, float* d_oc, baorec_stream stream) {
  BR_NEED_DIST(ctx);
  BR_REQUIRE(slot == 0 || slot == 1, "slot is 0 (data) or 1 (randoms)");
  BR_REQUIRE((!d_a || d_oa) && (!d_b || d_ob) && (!d_c || d_oc), "an input column without its output column");
  return unshard(ctx, slot, d_a, d_b, d_c, d_oa, d_ob, d_oc, (cudaStream_t)stream);
}

int baorec_dist_r2c_f32(baorec_ctx* ctx, const float* d_slab, float* d_kslab_t, baorec_stream stream) {
  BR_NEED_DIST(ctx);
  BR_REQUIRE(d_slab && d_kslab_t, "NULL pointer");
  return dist_r2c(ctx, d_slab, (float2*)d_kslab_t, (cudaStream_t)stream);
}

int baorec_dist_c2r_f32(baorec_ctx* ctx, float* d_kslab_t, float* d_slab, baorec_stream stream) {
  BR_NEED_DIST(ctx);
  BR_REQUIRE(d_slab && d_kslab_t, "NULL pointer");
  return dist_c2r(ctx, (float2*)d_kslab_t, d_slab, (cudaStream_t)stream);
}

// TSC scatter on slabs.  The 27-point stencil of a particle owned by this slab (owner = slab of its cic! base
// plane, the same sharding as CIC) reaches one plane below the slab and two above it, so the deposit goes into a
// scratch buffer of (nz_loc + 3) planes with the real planes at 1 .. nz_loc (slab_mode 3 = the gather's halo
// layout).  Then the boundary-cell exchange: the plane below goes to the previous rank and is added to its last
// real plane, the two planes above go to the next rank and are added to its first two; finally the real planes
// are copied to `loc` (planes 0 .. nz_loc-1), where the transform expects them.
static int dist_scatter_tsc(baorec_ctx* ctx, float* loc, float* x, float* y, float* z, const float* w, int64_t n,
                            int wrap, int mas, cudaStream_t st) {
  const int P = ctx->nranks, nzl = ctx->nz_loc;
  const size_t plane = (size_t)ctx->ny * ctx->nx;
  const int next = (ctx->rank + 1) % P, prev = (ctx->rank + P - 1) % P;
  BR_REQUIRE(nzl >= 2, "TSC on slabs needs at least two planes per rank");
  float* buf;
  BR_TRY(need_t(ctx, BUF_HALO, (size_t)(nzl + 5) * plane, &buf));
  float* halo = buf + (size_t)(nzl + 3) * plane;  // two planes of receive space
  BR_CUDA(cudaMemsetAsync(buf, 0, (size_t)(nzl + 3) * plane * sizeof(float), st));
  ctx->slab_mode = 3;
  int s = scatter(ctx, buf, x, y, z, w, n, wrap, mas, st);
  ctx->slab_mode = 0;
  if (s != BAOREC_OK) return s;
  // plane 0 (below the slab) -> previous rank; what arrives from the next rank belongs to our last real plane
  BR_TRY(ring_exchange(ctx, buf, prev, halo, next, plane, st));
  BR_LAUNCH(ctx, add_plane_kernel, cdiv(plane, 256), 256, 0, st, buf + (size_t)nzl * plane, halo, plane);
  // planes nz_loc+1, nz_loc+2 (above the slab) -> next rank; what arrives from the previous rank belongs to our first two
  BR_TRY(ring_exchange(ctx, buf + (size_t)(nzl + 1) * plane, next, halo, prev, 2 * plane, st));
  BR_LAUNCH(ctx, add_plane_kernel, cdiv(2 * plane, 256), 256, 0, st, buf + plane, halo, 2 * plane);
  BR_CUDA(cudaMemcpyAsync(loc, buf + plane, (size_t)nzl * plane * sizeof(float), cudaMemcpyDeviceToDevice, st));
  return BAOREC_OK;
}

// Scatter this rank's particles into `loc` ((nz_loc+1) planes, zero-filled here); the ghost plane
// (global plane z_lo + nz_loc) belongs to the next rank: send it, add what arrives from below.
static int dist_scatter(baorec_ctx* ctx, float* loc, float* x, float* y, float* z, const float* w, int64_t n,
                        int wrap, int mas, cudaStream_t st) {
  if (mas != BAOREC_MAS_CIC) return dist_scatter_tsc(ctx, loc, x, y, z, w, n, wrap, mas, st);  // TSC and PCS reach the same planes
  const int P = ctx->nranks, nzl = ctx->nz_loc;
  const size_t plane = (size_t)ctx->ny * ctx->nx;
  float* halo;
  BR_TRY(need_t(ctx, BUF_HALO, plane, &halo));
  BR_CUDA(cudaMemsetAsync(loc, 0, (size_t)(nzl + 1) * plane * sizeof(float), st));
  ctx->slab_mode = 1;
  int s = scatter(ctx, loc, x, y, z, w, n, wrap, mas, st);
  ctx->slab_mode = 0;
  if (s != BAOREC_OK) return s;
  BR_TRY(ring_exchange(ctx, loc + (size_t)nzl * plane, (ctx->rank + 1) % P, halo, (ctx->rank + P - 1) % P, plane, st));
  BR_LAUNCH(ctx, add_plane_kernel, cdiv(plane, 256), 256, 0, st, loc, halo, plane);
  return BAOREC_OK;
}

// rank 0 holds the DC mode of a transposed k slab: scal[slot] = sum(mesh), scal[slot+8] = mul / sum
static int dist_stash_dc(baorec_ctx* ctx, const float2* T, int slot, double mul, cudaStream_t st) {
  if (ctx->rank == 0) BR_TRY(stash_dc(ctx, T, slot, mul, st));
  return BAOREC_OK;
}

static int dist_bcast_scal(baorec_ctx* ctx, cudaStream_t st) {
  if (ctx->nranks > 1) BR_NCCL(ncclBroadcast(ctx->d_scal, ctx->d_scal, 16, ncclDouble, 0, comm_of(ctx), st));
  return BAOREC_OK;
}

// setup_overdensity! (src/recon.jl:42-57 box, :60-91 randoms) on slabs: delta lands in `delta_out`
// (nz_loc planes; must not alias BUF_RS plane 0.. when randoms are used -> see callers).
static int dist_setup_overdensity(baorec_ctx* ctx, const baorec_params* p, float* delta_out, float* x, float* y,
                                  float* z, const float* w, int64_t n, float* rx, float* ry, float* rz,
                                  const float* rw, int64_t nr, bool has_randoms, cudaStream_t st) {
  const int nzl = ctx->nz_loc;
  const size_t plane = (size_t)ctx->ny * ctx->nx;
  DistBufs b;
  BR_TRY(dist_bufs(ctx, &b));
  float* loc;
  BR_TRY(need_t(ctx, BUF_RS, (size_t)(nzl + 2) * plane, &loc));
  if (!has_randoms) {
    BR_TRY(dist_scatter(ctx, loc, x, y, z, w, n, 1, p->mas, st));
    BR_TRY(dist_r2c(ctx, loc, b.T, st));
    BR_TRY(dist_stash_dc(ctx, b.T, 0, 1.0 / (double)p->bias, st));
    BR_TRY(dist_bcast_scal(ctx, st));
    BR_TRY(kpass_setup_box_T(ctx, b.T, b.T, p, st));
    BR_TRY(dist_c2r(ctx, b.T, delta_out, st));
    return BAOREC_OK;
  }
  float *ran, *dat;
  BR_TRY(need_t(ctx, BUF_RAN, (size_t)(nzl + 1) * plane, &ran));
  BR_TRY(need_t(ctx, BUF_RX, (size_t)(nzl + 3) * plane, &dat));
  BR_TRY(dist_scatter(ctx, loc, x, y, z, w, n, 0, p->mas, st));
  BR_TRY(dist_scatter(ctx, ran, rx, ry, rz, rw, nr, 0, p->mas, st));
  BR_TRY(dist_r2c(ctx, loc, b.T, st));
  BR_TRY(dist_stash_dc(ctx, b.T, 0, 1.0, st));
  BR_TRY(kpass_gauss_T(ctx, b.T, b.T, p->smoothing_radius, st));
  BR_TRY(dist_c2r(ctx, b.T, dat, st));
  BR_TRY(dist_r2c(ctx, ran, b.T, st));
  BR_TRY(dist_stash_dc(ctx, b.T, 1, 1.0, st));
  BR_TRY(kpass_gauss_T(ctx, b.T, b.T, p->smoothing_radius, st));
  BR_TRY(dist_c2r(ctx, b.T, ran, st));
  BR_TRY(dist_bcast_scal(ctx, st));
  // global randoms count for the threshold (src/recon.jl:81)
  BR_LAUNCH(ctx, set_scalar_kernel, 1, 1, 0, st, ctx->d_scal + 2, (double)nr);
  if (ctx->nranks > 1)
    BR_NCCL(ncclAllReduce(ctx->d_scal + 2, ctx->d_scal + 2, 1, ncclDouble, ncclSum, comm_of(ctx), st));
  BR_TRY(randoms_combine(ctx, delta_out, dat, ran, p->bias, p->ran_min, -1.0, (size_t)nzl * plane, st));
  return BAOREC_OK;
}

int baorec_run_dist_f32(baorec_ctx* ctx, const baorec_params* p, int algorithm, float* d_x, float* d_y, float* d_z,
                        const float* d_w, int64_t n_local, float* d_rx, float* d_ry, float* d_rz, const float* d_rw,
                        int64_t n_ran_local, int has_randoms, float* d_mesh_slab, baorec_stream stream) {
  BR_NEED_DIST(ctx);
  BR_REQUIRE(p != nullptr && d_mesh_slab != nullptr, "NULL argument");
  BR_REQUIRE(algorithm == BAOREC_ITERATIVE || algorithm == BAOREC_MULTIGRID, "unknown algorithm");
  BR_REQUIRE(n_local >= 0 && (n_local == 0 || (d_x && d_y && d_z && d_w)), "particle arrays");
  BR_REQUIRE(n_ran_local >= 0 && (n_ran_local == 0 || (d_rx && d_ry && d_rz && d_rw)), "randoms arrays");
  BR_REQUIRE(has_randoms || n_ran_local == 0, "randoms passed with has_randoms == 0");
  BR_REQUIRE(p->mas == BAOREC_MAS_CIC || p->mas == BAOREC_MAS_TSC || p->mas == BAOREC_MAS_PCS, "unknown mass-assignment scheme");
  cudaStream_t st = (cudaStream_t)stream;
  const int nzl = ctx->nz_loc;
  const size_t plane = (size_t)ctx->ny * ctx->nx;
  const size_t slab_r = (size_t)nzl * plane;
  const size_t slab_c = (size_t)ctx->ny_loc * ctx->xh * ctx->nz;
  const float* los = p->has_los ? p->los : nullptr;
  ctx->kcache_valid = false;
  ctx->kcache_potential = false;
  ctx->disp_valid = false;
  BR_TRY(reset_oob(ctx, st));
  DistBufs b;
  BR_TRY(dist_bufs(ctx, &b));
  float2* keep;
  BR_TRY(need_t(ctx, BUF_CKCACHE, slab_c, &keep));
  float* rs;
  BR_TRY(need_t(ctx, BUF_RS, (size_t)(nzl + 2) * plane, &rs));

  if (algorithm == BAOREC_MULTIGRID) {
    // reconstructed_potential! (src/recon.jl:184-212): delta -> planes 1..nzl of the slab-layout rhs
    float* f_slab;
    if (has_randoms) {
      f_slab = rs;
      BR_TRY(dist_setup_overdensity(ctx, p, d_mesh_slab, d_x, d_y, d_z, d_w, n_local, d_rx, d_ry, d_rz, d_rw,
                                    n_ran_local, true, st));
      BR_CUDA(cudaMemcpyAsync(f_slab + plane, d_mesh_slab, slab_r * sizeof(float), cudaMemcpyDeviceToDevice, st));
    } else {
      // the C2R output goes straight to plane 1 (the scatter target at plane 0.. is dead by then)
      f_slab = rs;
      BR_TRY(dist_setup_overdensity(ctx, p, rs + plane, d_x, d_y, d_z, d_w, n_local, nullptr, nullptr, nullptr,
                                    nullptr, 0, false, st));
    }
    float* phi;
    BR_TRY(mg_fmg_dist(ctx, f_slab, &phi, p->beta, p->jacobi_damping_factor, p->jacobi_niterations,
                       p->vcycle_niterations, los, st));
    BR_CUDA(cudaMemcpyAsync(d_mesh_slab, phi + plane, slab_r * sizeof(float), cudaMemcpyDeviceToDevice, st));
    // phi_k for read_shifts (the reference transforms phi in every compute_displacements call,
    // src/multigrid.jl:787)
    BR_TRY(dist_r2c(ctx, d_mesh_slab, keep, st));
    ctx->kcache_potential = true;
  } else if (los && !has_randoms) {
    // fixed line of sight, periodic box: smoothing, normalisation and all iterations in one pass
    BR_TRY(dist_scatter(ctx, rs, d_x, d_y, d_z, d_w, n_local, 1, p->mas, st));
    BR_TRY(dist_r2c(ctx, rs, b.T, st));
    // sum(rho) = DC mode, held by rank 0: broadcast it with the derived scale
    BR_TRY(dist_stash_dc(ctx, b.T, 0, (double)ctx->M / (double)p->bias, st));
    BR_TRY(dist_bcast_scal(ctx, st));
    BR_TRY(kpass_fused_T(ctx, b.T, b.T, keep, p, st));
    BR_TRY(dist_c2r(ctx, b.T, d_mesh_slab, st));
  } else {
    // delta_s in planes 1..nzl of RS (kept for the whole iteration)
    float* ds = rs + plane;
    if (has_randoms) {
      BR_TRY(dist_setup_overdensity(ctx, p, d_mesh_slab, d_x, d_y, d_z, d_w, n_local, d_rx, d_ry, d_rz, d_rw,
                                    n_ran_local, true, st));
      BR_CUDA(cudaMemcpyAsync(ds, d_mesh_slab, slab_r * sizeof(float), cudaMemcpyDeviceToDevice, st));
    } else {
      BR_TRY(dist_setup_overdensity(ctx, p, ds, d_x, d_y, d_z, d_w, n_local, nullptr, nullptr, nullptr, nullptr, 0,
                                    false, st));
    }
    if (los) {
      BR_TRY(dist_r2c(ctx, ds, b.T, st));
      BR_TRY(kpass_fused_delta_T(ctx, b.T, b.T, keep, p, st));
      BR_TRY(dist_c2r(ctx, b.T, d_mesh_slab, st));
    } else {
      // radial line of sight: iterate! (src/iterative.jl:151-211) on slabs, 1 R2C + 6 C2R per iteration
      float2* pair;
      float* X;
      BR_TRY(need_t(ctx, BUF_CK2, slab_c, &pair));
      BR_TRY(need_t(ctx, BUF_RX, (size_t)(nzl + 3) * plane, &X));
      if (p->n_iter <= 0) BR_CUDA(cudaMemcpyAsync(d_mesh_slab, ds, slab_r * sizeof(float), cudaMemcpyDeviceToDevice, st));
      for (int it = 1; it <= p->n_iter; it++) {
        BR_TRY(dist_r2c(ctx, it == 1 ? ds : d_mesh_slab, b.T, st));
        bool first = true;
        for (int i = 0; i < 3; i++)
          for (int j = i; j < 3; j++) {
            float fac = (float)((1.0 + (i != j ? 1.0 : 0.0)) * (double)p->beta);
            if (it == 1) fac = fac / (1.0f + p->beta);
            BR_TRY(kpass_iter_pair_T(ctx, b.T, pair, i, j, st));
            BR_TRY(dist_c2r(ctx, pair, X, st));
            BR_TRY(radial_update_slab(ctx, d_mesh_slab, first ? ds : d_mesh_slab, X, nzl, ctx->z0, i, j, fac, st));
            first = false;
          }
      }
      BR_TRY(dist_r2c(ctx, d_mesh_slab, keep, st));  // delta_k of the result for read_shifts
    }
  }
  BR_TRY(check_oob(ctx, st, "run_dist (particles must lie in this rank's slab)"));
  BR_TRY(peer_check(ctx, st));
  ctx->kcache_valid = true;
  return BAOREC_OK;
}

int baorec_read_shifts_dist_f32(baorec_ctx* ctx, const baorec_params* p, const float* d_x, const float* d_y,
                                const float* d_z, int64_t n_local, int field, int positions, float* d_sx,
                                float* d_sy, float* d_sz, baorec_stream stream) {
  BR_NEED_DIST(ctx);
  BR_REQUIRE(p != nullptr, "params is NULL");
  BR_REQUIRE(field >= BAOREC_FIELD_DISP && field <= BAOREC_FIELD_SUM, "unknown field");
  BR_REQUIRE(n_local >= 0 && (n_local == 0 || (d_x && d_y && d_z && d_sx && d_sy && d_sz)), "particle arrays");
  if (!ctx->kcache_valid) {
    set_error("baorec_read_shifts_dist_f32: call baorec_run_dist_f32 first (no cached delta_k / phi_k)");
    return BAOREC_ERR_INVALID;
  }
  cudaStream_t st = (cudaStream_t)stream;
  const int P = ctx->nranks, nzl = ctx->nz_loc;
  const size_t plane = (size_t)ctx->ny * ctx->nx;
  const int next = (ctx->rank + 1) % P, prev = (ctx->rank + P - 1) % P;
  DistBufs b;
  BR_TRY(dist_bufs(ctx, &b));
  const float2* keep = (const float2*)ctx->bufs[BUF_CKCACHE].p;
  float* psi[3];
  BR_TRY(need_t(ctx, BUF_RX, (size_t)(nzl + 3) * plane, &psi[0]));
  BR_TRY(need_t(ctx, BUF_RY, (size_t)(nzl + 3) * plane, &psi[1]));
  BR_TRY(need_t(ctx, BUF_RZ, (size_t)(nzl + 3) * plane, &psi[2]));
  // the displacement slabs (+ halos) of the last run are kept across calls (data, then randoms, ...)
  const bool reuse = ctx->disp_valid && ctx->disp_algo == -1;
  ctx->disp_valid = false;
  BR_TRY(reset_oob(ctx, st));
  if (!reuse) {  // the tile sort of the catalog overlaps the three distributed transforms
    ctx->slab_mode = 2;
    int sp = gather_prebin(ctx, d_x, d_y, d_z, n_local, p->mas, st);
    ctx->slab_mode = 0;
    if (sp != BAOREC_OK) return sp;
  }
  if (!reuse && ctx->p2p) {
    float* own[3] = {psi[0] + plane, psi[1] + plane, psi[2] + plane};  // own planes at local index 1 .. nzl
    BR_TRY(dist_displacements_peer(ctx, keep, ctx->kcache_potential, own, st));
  }
  for (int c = 0; c < 3 && !reuse; c++) {
    if (!ctx->p2p) {
      BR_TRY(kpass_disp_T(ctx, keep, b.T, c, ctx->kcache_potential, st));
      BR_TRY(dist_c2r(ctx, b.T, psi[c] + plane, st));  // own planes at local index 1 .. nzl
    }
    // halo: plane 0 <- previous rank's last plane; planes nzl+1, nzl+2 <- next rank's first two
    BR_TRY(ring_exchange(ctx, psi[c] + (size_t)nzl * plane, next, psi[c], prev, plane, st));
    BR_TRY(ring_exchange(ctx, psi[c] + plane, prev, psi[c] + (size_t)(nzl + 1) * plane, next, 2 * plane, st));
  }
  ctx->slab_mode = 2;
  int s = gather3(ctx, psi[0], psi[1], psi[2], d_x, d_y, d_z, n_local, d_sx, d_sy, d_sz, p->mas, field, p->f,
                  p->has_los, p->los, positions, st);
  ctx->slab_mode = 0;
  if (s != BAOREC_OK) return s;
  ctx->disp_valid = true;
  ctx->disp_algo = -1;  // marks slab-layout displacement meshes
  BR_TRY(check_oob(ctx, st, "read_shifts_dist (particles must lie in this rank's slab)"));
  return peer_check(ctx, st);
}

// One reconstruction from an UNSHARDED catalog: every rank passes any part of the data (and randoms) catalog as
// device arrays; the library shards by slab, runs run! + read_shifts / reconstructed_positions on the slabs and routes
// the per-particle results back into the caller's order.  The box must have been planned (with randoms: from
// setup_box over all ranks, src/recon.jl:172,253).  The result mesh slab stays in the library (BUF_CACHE).
int baorec_reconstruct_dist_f32(baorec_ctx* ctx, const baorec_params* p, int algorithm, const float* d_x, const float* d_y,
                                const float* d_z, const float* d_w, int64_t n, const float* d_rx, const float* d_ry,
                                const float* d_rz, const float* d_rw, int64_t n_ran, int has_randoms, int field,
                                int positions, float* d_ox, float* d_oy, float* d_oz, baorec_stream stream) {
  BR_NEED_DIST(ctx);
  BR_REQUIRE(p != nullptr, "params is NULL");
  BR_REQUIRE(n >= 0 && (n == 0 || (d_x && d_y && d_z && d_w && d_ox && d_oy && d_oz)), "particle arrays");
  BR_REQUIRE(n_ran >= 0 && (n_ran == 0 || (d_rx && d_ry && d_rz && d_rw)), "randoms arrays");
  cudaStream_t st = (cudaStream_t)stream;
  float *sx, *sy, *sz, *sw, *rx = nullptr, *ry = nullptr, *rz = nullptr, *rw = nullptr;
  int64_t nl = 0, nrl = 0;
  BR_TRY(shard_catalog(ctx, 0, d_x, d_y, d_z, d_w, n, &sx, &sy, &sz, &sw, &nl, st));
  if (has_randoms) BR_TRY(shard_catalog(ctx, 1, d_rx, d_ry, d_rz, d_rw, n_ran, &rx, &ry, &rz, &rw, &nrl, st));
  float* mesh;
  BR_TRY(need_t(ctx, BUF_CACHE, (size_t)ctx->nz_loc * ctx->ny * ctx->nx, &mesh));
  BR_TRY(baorec_run_dist_f32(ctx, p, algorithm, sx, sy, sz, sw, nl, rx, ry, rz, rw, nrl, has_randoms, mesh, stream));
  float* out;
  const size_t no = (size_t)(nl > 0 ? nl : 1);
  BR_TRY(need_t(ctx, BUF_SHARD_OUT, 3 * no, &out));
  BR_TRY(baorec_read_shifts_dist_f32(ctx, p, sx, sy, sz, nl, field, positions, out, out + no, out + 2 * no, stream));
  return unshard(ctx, 0, out, out + no, out + 2 * no, d_ox, d_oy, d_oz, st);
}

// The same with HOST arrays (this rank's share of the catalog; pinned memory copies at PCIe speed): upload, the call
// above, download.  This is what a host program that holds an unsharded catalog calls on every rank.
int baorec_reconstruct_dist_host_f32(baorec_ctx* ctx, const baorec_params* p, int algorithm, const float* h_x,
                                     const float* h_y, const float* h_z, const float* h_w, int64_t n, int field,
                                     int positions, float* h_ox, float* h_oy, float* h_oz) {
  BR_NEED_DIST(ctx);
  BR_REQUIRE(n >= 0 && (n == 0 || (h_x && h_y && h_z && h_w && h_ox && h_oy && h_oz)), "particle arrays");
  cudaStream_t st = ctx->own_stream;
  const size_t nn = (size_t)(n > 0 ? n : 1);
  float *dp, *dout;
  BR_TRY(need_t(ctx, BUF_PART, 4 * nn, &dp));
  BR_TRY(need_t(ctx, BUF_OUT, 3 * nn, &dout));
  const float* hs[4] = {h_x, h_y, h_z, h_w};
  for (int c = 0; c < 4 && n > 0; c++)
    BR_CUDA(cudaMemcpyAsync(dp + c * nn, hs[c], (size_t)n * sizeof(float), cudaMemcpyHostToDevice, st));
  BR_TRY(baorec_reconstruct_dist_f32(ctx, p, algorithm, dp, dp + nn, dp + 2 * nn, dp + 3 * nn, n, nullptr, nullptr, nullptr,
                                     nullptr, 0, 0, field, positions, dout, dout + nn, dout + 2 * nn, (baorec_stream)st));
  float* hd[3] = {h_ox, h_oy, h_oz};
  for (int c = 0; c < 3 && n > 0; c++)
    BR_CUDA(cudaMemcpyAsync(hd[c], dout + c * nn, (size_t)n * sizeof(float), cudaMemcpyDeviceToHost, st));
  BR_CUDA(cudaStreamSynchronize(st));
  return BAOREC_OK;
}

}  // extern "C"
