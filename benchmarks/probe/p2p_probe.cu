// Probe (not product): how fast can ONE process drive an all-to-all of P GPUs with (a) copy-engine copies, one stream
// per GPU, peers in staggered order, contiguous; (b) the same with 2-D strided copies (pitch != width), (c) one stream
// per (GPU, peer), (d) an SM kernel that stores to all peers at once.  Message = `mb` MB per (src, dst) pair.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a p2p_probe.cu -o p2p_probe && ./p2p_probe 16.8
#include <cuda_runtime.h>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <vector>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e)); exit(1); } } while (0)

struct Tab { float4* p[8]; };

__global__ void __launch_bounds__(256) push_all(const __grid_constant__ Tab dst, const float4* __restrict__ src, size_t n4, int P, int me) {
  // blockIdx.y = peer step k (1 .. P-1); grid-stride over the message
  const int k = blockIdx.y + 1, d = (me + k) % P;
  const float4* s = src + (size_t)d * n4;
  float4* o = dst.p[d] + (size_t)me * n4;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  for (; i + 3 * stride < n4; i += 4 * stride) {
    float4 a = s[i], b = s[i + stride], c = s[i + 2 * stride], e = s[i + 3 * stride];
    o[i] = a; o[i + stride] = b; o[i + 2 * stride] = c; o[i + 3 * stride] = e;
  }
  for (; i < n4; i += stride) o[i] = s[i];
}

int main(int argc, char** argv) {
  double mb = argc > 1 ? atof(argv[1]) : 16.8;
  int P = 0;
  CK(cudaGetDeviceCount(&P));
  if (argc > 2) P = atoi(argv[2]);
  const size_t n4 = (size_t)(mb * 1e6 / 16), bytes = n4 * 16;
  std::vector<float4*> S(P), R(P);
  std::vector<std::vector<cudaStream_t>> st(P, std::vector<cudaStream_t>(P));
  for (int g = 0; g < P; g++) {
    CK(cudaSetDevice(g));
    for (int h = 0; h < P; h++) if (h != g) { cudaError_t e = cudaDeviceEnablePeerAccess(h, 0); if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) { printf("no P2P %d->%d\n", g, h); return 1; } cudaGetLastError(); }
    CK(cudaMalloc(&S[g], bytes * P * 2));   // x2: room for a strided source (pitch = 2 * width)
    CK(cudaMalloc(&R[g], bytes * P));
    CK(cudaMemset(S[g], 1, bytes * P * 2));
    for (int h = 0; h < P; h++) CK(cudaStreamCreateWithFlags(&st[g][h], cudaStreamNonBlocking));
  }
  auto sync_all = [&]() { for (int g = 0; g < P; g++) { CK(cudaSetDevice(g)); CK(cudaDeviceSynchronize()); } };
  const int reps = 20;
  const size_t rows = 32, width = bytes / rows;
  for (int mode = 0; mode < 5; mode++) {
    double best = 1e30;
    for (int trial = 0; trial < 3; trial++) {
      sync_all();
      auto t0 = std::chrono::high_resolution_clock::now();
      for (int r = 0; r < reps; r++)
        for (int g = 0; g < P; g++) {
          CK(cudaSetDevice(g));
          if (mode == 4) {
            Tab t; for (int h = 0; h < P; h++) t.p[h] = R[h];
            dim3 grid(mode == 4 ? 16 : 8, P - 1);
            push_all<<<grid, 256, 0, st[g][0]>>>(t, S[g], n4, P, g);
            continue;
          }
          for (int k = 1; k < P; k++) {
            const int d = (g + k) % P;
            cudaStream_t s = (mode == 2 || mode == 3) ? st[g][k] : st[g][0];
            if (mode == 1 || mode == 3)
              CK(cudaMemcpy2DAsync((char*)R[d] + (size_t)g * bytes, width, (char*)S[g] + (size_t)d * bytes * 2, 2 * width, width, rows, cudaMemcpyDefault, s));
            else
              CK(cudaMemcpyAsync((char*)R[d] + (size_t)g * bytes, (char*)S[g] + (size_t)d * bytes, bytes, cudaMemcpyDefault, s));
          }
        }
      sync_all();
      double dt = std::chrono::duration<double>(std::chrono::high_resolution_clock::now() - t0).count();
      if (dt < best) best = dt;
    }
    const char* names[5] = {"CE contiguous, 1 stream/GPU, staggered peers", "CE 2-D strided (32 rows), 1 stream/GPU", "CE contiguous, 1 stream per peer",
                            "CE 2-D strided, 1 stream per peer", "SM kernel, all peers at once (16 x (P-1) CTAs)"};
    printf("P=%d msg=%.1f MB  %-48s %7.1f GB/s per GPU per direction  (%.1f us per all-to-all)\n", P, mb, names[mode],
           (double)bytes * (P - 1) * reps / best / 1e9, best / reps * 1e6);
  }
  return 0;
}
