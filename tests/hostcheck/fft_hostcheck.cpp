// Host check of the own column FFT (baorec.jl_b200/csrc/fft_column.cuh + the generated butterflies of fft_radix.cuh):
// the product's device functions compiled as plain C++ (csrc/host_shim.cuh) and walked over the 32 "threads" of every
// column of a tile -- first fft_column_spread for all of them (what precedes the kernel's __syncthreads), then
// fft_column_collect -- with an ordinary array for the CTA's shared-memory exchange buffer.  Outputs are stored where the
// kernels store them (frequency k2 + M k1 from X[m][fft_bitrev<32>(k1)], k2 = t + 32 m), so that tests/test_fft_hostcheck.py
// can hold them to numpy's FFT.
#define __host__
#include <vector>

#include "fft_column.cuh"

using namespace baorec;

template <int R, int DIR>
static void reg_fft(const float* in, float* out) {
  float2 v[R];
  for (int j = 0; j < R; j++) v[j] = make_float2(in[2 * j], in[2 * j + 1]);
  fft_reg<R, DIR>(v);
  for (int k = 0; k < R; k++) {
    out[2 * k] = v[fft_bitrev<R>(k)].x;
    out[2 * k + 1] = v[fft_bitrev<R>(k)].y;
  }
}

extern "C" int hc_fft_reg(int R, int dir, const float* in, float* out) {
#define CASE(RR)                                 \
  if (R == RR) {                                 \
    if (dir > 0) reg_fft<RR, 1>(in, out);        \
    else reg_fft<RR, -1>(in, out);               \
    return 0;                                    \
  }
  CASE(8) CASE(16) CASE(32) CASE(64)
#undef CASE
  return -1;
}

// in / out: TX columns of N complex values, column-major (column c at offset c * N); tw[k] = exp(-2 pi i k / N)
template <int N, int DIR, int TX>
static void tile_fft(const float2* in, const float2* tw, float2* out) {
  constexpr int M = N / 32, MT = M >= 32 ? M / 32 : 1;
  std::vector<float2> S((size_t)TX * Xch<N, TX>::COL, make_float2(0.f, 0.f));
  for (int c = 0; c < TX; c++)
    for (int t = 0; t < 32; t++) {
      float2 v[M];
      for (int j = 0; j < M; j++) v[j] = in[(size_t)c * N + t + 32 * j];   // load_column: rows t, t + 32, ...
      fft_column_spread<N, DIR, false, TX>(v, S.data(), tw, c, t);
    }
  for (int c = 0; c < TX; c++)
    for (int t = 0; t < 32; t++) {
      float2 X[MT][32];
      fft_column_collect<N, DIR, TX>(X, S.data(), c, t);
      for (int m = 0; m < MT; m++) {
        const int k2 = t + 32 * m;
        if (M >= 32 || k2 < M)
          for (int k1 = 0; k1 < 32; k1++) out[(size_t)c * N + k2 + M * k1] = X[m][fft_bitrev<32>(k1)];
      }
    }
}

extern "C" int hc_fft_tile(int N, int dir, int TX, const float* in, const float* tw, float* out) {
  const float2* i2 = reinterpret_cast<const float2*>(in);
  const float2* t2 = reinterpret_cast<const float2*>(tw);
  float2* o2 = reinterpret_cast<float2*>(out);
#define CASE(NN, TT)                                   \
  if (N == NN && TX == TT) {                           \
    if (dir > 0) tile_fft<NN, 1, TT>(i2, t2, o2);      \
    else tile_fft<NN, -1, TT>(i2, t2, o2);             \
    return 0;                                          \
  }
  CASE(256, 8) CASE(512, 8) CASE(1024, 8) CASE(2048, 8) CASE(1024, 4) CASE(1024, 16)
#undef CASE
  return -1;
}

extern "C" int hc_fft_exchange_floats(int N, int TX) {   // size of the exchange buffer, for the bank-layout assertions
  if (N == 1024 && TX == 16) return (int)(Xch<1024, 16>::BYTES / sizeof(float));
  if (N == 1024 && TX == 8) return (int)(Xch<1024, 8>::BYTES / sizeof(float));
  if (N == 1024 && TX == 4) return (int)(Xch<1024, 4>::BYTES / sizeof(float));
  if (N == 2048 && TX == 8) return (int)(Xch<2048, 8>::BYTES / sizeof(float));
  if (N == 512 && TX == 8) return (int)(Xch<512, 8>::BYTES / sizeof(float));
  if (N == 256 && TX == 8) return (int)(Xch<256, 8>::BYTES / sizeof(float));
  return -1;
}
