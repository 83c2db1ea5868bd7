// Microbenchmark: the v3 loader structure of bjx_lik_tc.cu in isolation -- W loader warps per SM, each owning a
// quarter-block unit (16 rows x 256 B, U x LDG.128 per lane), depth 1 per warp, with a busy "convert+store" delay of
// `work` cycles per unit.  Reports achieved HBM GB/s.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

template <int W, int ROWSTEPS>   // ROWSTEPS row-steps of 4 rows: unit = 4*ROWSTEPS rows x 256 B
__global__ void __launch_bounds__(W * 32, 1) k(const float* __restrict__ X, int64_t N, int work, float* out) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int rl = lane >> 3, c = lane & 7;
  constexpr int UNITS_PER_BLOCK = 64 / (4 * ROWSTEPS);
  const int qd = warp % UNITS_PER_BLOCK, bl = warp / UNITS_PER_BLOCK;
  constexpr int BL = W / UNITS_PER_BLOCK;
  const int64_t n_tiles = N / 64;
  const int64_t my_tiles = (n_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x;
  const int64_t total = my_tiles * 8;
  float acc = 0.f;
  for (int64_t g = bl; g < total; g += BL) {
    const int64_t it = g / 8; const int b = (int)(g % 8);
    const int64_t row0 = (blockIdx.x + it * gridDim.x) * 64 + 4 * ROWSTEPS * qd + rl;
    float4 v[2 * ROWSTEPS];
#pragma unroll
    for (int i = 0; i < ROWSTEPS; ++i) {
      const float4* p = reinterpret_cast<const float4*>(X + (row0 + 4 * i) * 512 + b * 64 + c * 8);
      v[2 * i] = __ldg(p); v[2 * i + 1] = __ldg(p + 1);
    }
#pragma unroll
    for (int i = 0; i < 2 * ROWSTEPS; ++i) acc += v[i].x + v[i].y + v[i].z + v[i].w;
    const long long t0 = clock64();
    while (clock64() - t0 < work) {}
  }
  if (acc == 123.456f) out[0] = acc;
}

template <int W, int RS>
void run(const float* X, int64_t N, float* out, int work) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<W, RS><<<148, W * 32>>>(X, N, work, out);
  cudaEventRecord(e0);
  for (int i = 0; i < 3; ++i) k<W, RS><<<148, W * 32>>>(X, N, work, out);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= 3;
  printf("  warps=%2d unit=%2d rows (%4.1f KB/warp, %3d KB/SM in flight) work=%4d: %.3f ms  %.0f GB/s\n", W, 4 * RS, RS * 1.0,
         W * RS, work, ms, N * 2048.0 / ms / 1e6);
}

int main() {
  const int64_t N = 4000000;
  float *X, *out; cudaMalloc(&X, N * 2048); cudaMalloc(&out, 4); cudaMemset(X, 0, N * 2048);
  for (int work : {0, 600}) {
    run<16, 4>(X, N, out, work); run<24, 4>(X, N, out, work); run<32, 4>(X, N, out, work);
    run<16, 8>(X, N, out, work); run<8, 8>(X, N, out, work); run<16, 16>(X, N, out, work);
  }
  return 0;
}
