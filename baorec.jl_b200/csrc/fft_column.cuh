// The four-step column transform of csrc/fft.cu (N = 32 x M; M-point FFT in registers, W_N^(t k) twiddles, one exchange
// through shared memory, 32-point FFT in registers) as a header: the kernels of fft.cu include it, and
// tests/hostcheck/fft_hostcheck.cpp compiles the same functions as plain C++ (csrc/host_shim.cuh) to hold them to
// numpy's FFT on the CPU -- generated butterflies, twiddle indices, exchange layout and output order.
#pragma once
#include "host_shim.cuh"
#include "fft_radix.cuh"

namespace baorec {

constexpr int FFT_TX = 8;          // columns per tile (8 x 8 B = 64 B row segments): the default; 4 and 16 are instantiated for N = 1024
                                   // (option "fft_tile_cols") -- 32 threads per column, so 128 / 256 / 512 threads per CTA

__device__ __forceinline__ float2 cmul(float2 a, float2 b) {
  return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
// TWS: the table is the CTA's shared-memory copy (fused z passes) / the global table through L1 (plain passes)
template <int DIR, bool TWS>
__device__ __forceinline__ float2 twd(const float2* tw, int idx) {
  float2 w = TWS ? tw[idx] : __ldg(tw + idx);  // exp(-2 pi i idx / N)
  if (DIR < 0) w.y = -w.y;
  return w;
}

// Exchange buffer of the four-step FFT: per column M rows (k2) of 33 slots (n1; 33 = one pad slot, so that the reads
// of a row by consecutive threads of a warp fall into different banks) and a column stride that is 16 bytes off a
// multiple of 128: a half-warp (8 columns x 2 threads, one 8-byte access each) then covers the 32 banks exactly once
// for the writes (slot t) and for the reads (row t).  (First version: 32 bytes off, columns c and c + 4 collided:
// ncu counted 2.2 extra wavefronts per shared-memory instruction.)
template <int N, int TX = FFT_TX>
struct Xch {
  static constexpr int M = N / 32;
  static constexpr int pad() {  // column stride = 128 / TX bytes off a multiple of 128 (TX columns x 16 / TX threads per half-warp)
    int p = 0;
    while (((M * 33 + p) * 8) % 128 != 128 / TX) p++;
    return p;
  }
  static constexpr int COL = M * 33 + pad();
  static constexpr size_t BYTES = (size_t)TX * COL * sizeof(float2);
};

// One length-N transform per (column c, 32 threads t): v[j] = x[t + 32 j] in, out[m][r] with
// out[m][fft_bitrev<32>(k1)] = X[(t + 32 m) + M k1] for the transforms this thread owns in the last step
// (m < MT = max(1, M / 32); for M < 32 only threads t < M own one).  S = this CTA's exchange buffer.
// Split at the barrier into fft_column_spread / fft_column_collect so that tests/hostcheck/fft_hostcheck.cpp can walk the
// very same code over the 32 threads of every column on the CPU (same SASS as the one-piece version).
template <int N, int DIR, bool TWS, int TX = FFT_TX>
__device__ __forceinline__ void fft_column_spread(float2 (&v)[N / 32], float2* __restrict__ S, const float2* tw, int c, int t) {
  constexpr int M = N / 32, COL = Xch<N, TX>::COL;
  fft_reg<M, DIR>(v);
  float2* col = S + c * COL;
#pragma unroll
  for (int k2 = 0; k2 < M; k2++) {
    float2 y = v[fft_bitrev<M>(k2)];
    if (k2) y = cmul(y, twd<DIR, TWS>(tw, (t * k2) & (N - 1)));
    col[k2 * 33 + t] = y;
  }
}
template <int N, int DIR, int TX = FFT_TX>
__device__ __forceinline__ void fft_column_collect(float2 (&out)[(N / 32 >= 32 ? N / 1024 : 1)][32], const float2* __restrict__ S, int c, int t) {
  constexpr int M = N / 32, MT = M >= 32 ? M / 32 : 1, COL = Xch<N, TX>::COL;
  const float2* col = S + c * COL;
#pragma unroll
  for (int m = 0; m < MT; m++) {
    const int k2 = t + 32 * m;
    if (M >= 32 || k2 < M) {
#pragma unroll
      for (int n1 = 0; n1 < 32; n1++) out[m][n1] = col[k2 * 33 + n1];
      fft_reg<32, DIR>(out[m]);
    }
  }
}

#if defined(__CUDACC__)
template <int N, int DIR, bool TWS, int TX = FFT_TX>
__device__ __forceinline__ void fft_column(float2 (&v)[N / 32], float2 (&out)[(N / 32 >= 32 ? N / 1024 : 1)][32],
                                           float2* __restrict__ S, const float2* tw, int c, int t) {
  fft_column_spread<N, DIR, TWS, TX>(v, S, tw, c, t);
  __syncthreads();
  fft_column_collect<N, DIR, TX>(out, S, c, t);
}
#endif

}  // namespace baorec
