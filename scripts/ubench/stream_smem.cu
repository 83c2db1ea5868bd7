// Microbenchmark: does the shared-memory carve-out (which shrinks the L1 data cache to ~28 KB at 200+ KB of smem) limit
// the bytes a loader-warp structure can keep in flight?  W loader warps per SM, unit = 16 rows x 256 B per warp
// (8 x LDG.128 per lane), depth 1 per warp; dynamic smem 0 / 100 / 200 KB; load flavours: ld.global.nc, +L1::no_allocate,
// ld.global.cv-free variants.  Reports achieved HBM GB/s.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

template <int MODE>
__device__ __forceinline__ float4 ld16(const float4* p) {
  float4 v;
  if (MODE == 0) {
    v = __ldg(p);
  } else if (MODE == 1) {
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
  } else if (MODE == 2) {
    asm volatile("ld.global.cg.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
  } else {
    asm volatile("ld.global.nc.L1::no_allocate.L2::256B.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
  }
  return v;
}

template <int W, int MODE>
__global__ void __launch_bounds__(W * 32, 1) k(const float* __restrict__ X, int64_t N, int pf, float* out) {
  extern __shared__ float dummy[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int rl = lane >> 3, c = lane & 7;
  const int qd = warp % 4, bl = warp / 4;
  constexpr int BL = W / 4;
  const int64_t n_tiles = N / 64;
  const int64_t my_tiles = (n_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x;
  const int64_t total = my_tiles * 8;
  float acc = 0.f;
  for (int64_t g = bl; g < total; g += BL) {
    const int64_t it = g / 8; const int b = (int)(g % 8);
    const int64_t row0 = (blockIdx.x + it * gridDim.x) * 64 + 16 * qd + rl;
    float4 v[8];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float4* p = reinterpret_cast<const float4*>(X + (row0 + 4 * i) * 512 + b * 64 + c * 8);
      v[2 * i] = ld16<MODE>(p); v[2 * i + 1] = ld16<MODE>(p + 1);
    }
    if (pf) {
      const int64_t gp = g + 8;
      const int64_t prow = (blockIdx.x + (gp / 8) * gridDim.x) * 64 + 16 * qd + (lane >> 1);
      if (gp < total) asm volatile("prefetch.global.L2 [%0];" ::"l"(X + prow * 512 + (gp % 8) * 64 + (lane & 1) * 32));
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) acc += v[i].x + v[i].y + v[i].z + v[i].w;
  }
  if (acc == 123.456f) { out[0] = acc; dummy[threadIdx.x] = acc; }
}

template <int W, int MODE>
void run(const float* X, int64_t N, float* out, int smem_kb, int pf) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaFuncSetAttribute(k<W, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
  const size_t sm = (size_t)smem_kb * 1024;
  k<W, MODE><<<148, W * 32, sm>>>(X, N, pf, out);
  cudaEventRecord(e0);
  for (int i = 0; i < 3; ++i) k<W, MODE><<<148, W * 32, sm>>>(X, N, pf, out);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= 3;
  printf("  warps=%2d mode=%d smem=%3d KB pf=%d: %.3f ms  %.0f GB/s  (%s)\n", W, MODE, smem_kb, pf, ms, N * 2048.0 / ms / 1e6,
         cudaGetErrorString(cudaGetLastError()));
}

int main() {
  const int64_t N = 4000000;
  float *X, *out; cudaMalloc(&X, N * 2048); cudaMalloc(&out, 4); cudaMemset(X, 0, N * 2048);
  for (int pf : {0, 1})
    for (int smem : {0, 100, 200}) {
      run<16, 0>(X, N, out, smem, pf); run<16, 1>(X, N, out, smem, pf); run<16, 2>(X, N, out, smem, pf); run<16, 3>(X, N, out, smem, pf);
      run<24, 0>(X, N, out, smem, pf); run<24, 1>(X, N, out, smem, pf);
    }
  return 0;
}
