// Microbenchmark: achievable HBM read bandwidth of the loader access pattern of bjx_lik_tc.cu (no MMA, no smem).
// 148 CTAs x 256 threads; tile = 64 rows x 512 fp32 (row pitch 2 KB); "col" = 8 blocks of [64 rows x 64 cols] visited
// block by block (what the kernel does); "row" = 8 slabs of [8 rows x 512 cols] (16 KB contiguous each).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

template <int PF, int ROWMAJOR>
__global__ void __launch_bounds__(256, 1) k(const float* __restrict__ X, int64_t N, float* out) {
  const int tid = threadIdx.x;
  const int64_t n_tiles = N / 64;
  const int64_t my_tiles = (n_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x;
  const int64_t total = my_tiles * 8;
  float4 buf[PF][4];
  float acc = 0.f;
  auto issue = [&](int64_t g, float4 (&b)[4]) {
    if (g >= total) return;
    const int64_t it = g / 8; const int blk = (int)(g % 8);
    const int64_t row0 = (blockIdx.x + it * gridDim.x) * 64;
    if (ROWMAJOR) {   // slab = 8 rows x 512 cols: thread -> (row = tid/32, 16 consecutive floats)
      const float4* p = reinterpret_cast<const float4*>(X + (row0 + blk * 8 + (tid >> 5)) * 512 + (tid & 31) * 16);
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = __ldg(p + j);
    } else {          // block = 64 rows x 64 cols: thread -> rows tid/8, tid/8+32; 8 floats at col blk*64 + (tid%8)*8
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const float4* p = reinterpret_cast<const float4*>(X + (row0 + (tid >> 3) + 32 * h) * 512 + blk * 64 + (tid & 7) * 8);
        b[2 * h] = __ldg(p); b[2 * h + 1] = __ldg(p + 1);
      }
    }
  };
  auto use = [&](float4 (&b)[4]) {
#pragma unroll
    for (int j = 0; j < 4; ++j) acc += b[j].x + b[j].y + b[j].z + b[j].w;
  };
#pragma unroll
  for (int p = 0; p < PF; ++p) issue(p, buf[p]);
  for (int64_t g = 0; g < total; g += PF) {
#pragma unroll
    for (int p = 0; p < PF; ++p) {
      if (g + p < total) use(buf[p]);
      issue(g + p + PF, buf[p]);
    }
  }
  if (acc == 123.456f) out[0] = acc;
}

template <int PF, int RM>
void run(const float* X, int64_t N, float* out, const char* name) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<PF, RM><<<148, 256>>>(X, N, out);
  cudaEventRecord(e0);
  for (int i = 0; i < 3; ++i) k<PF, RM><<<148, 256>>>(X, N, out);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= 3;
  printf("  %-28s PF=%d: %.3f ms  %.0f GB/s\n", name, PF, ms, N * 2048.0 / ms / 1e6);
}

int main() {
  const int64_t N = 4000000;
  float *X, *out; cudaMalloc(&X, N * 2048); cudaMalloc(&out, 4); cudaMemset(X, 0, N * 2048);
  run<1, 0>(X, N, out, "column blocks (kernel)"); run<2, 0>(X, N, out, "column blocks (kernel)");
  run<4, 0>(X, N, out, "column blocks (kernel)"); run<8, 0>(X, N, out, "column blocks (kernel)");
  run<1, 1>(X, N, out, "row slabs (contiguous)"); run<2, 1>(X, N, out, "row slabs (contiguous)");
  run<4, 1>(X, N, out, "row slabs (contiguous)"); run<8, 1>(X, N, out, "row slabs (contiguous)");
  return 0;
}
