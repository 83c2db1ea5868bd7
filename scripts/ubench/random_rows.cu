// What HBM delivers for RANDOM ROW reads: N rows of ROW_BYTES each (a [N, 512] fp32 matrix = 2 KB rows, or two bf16 planes of
// 1 KB rows), B row indices drawn uniformly with replacement, every row read exactly as the minibatch step reads it (whole row,
// 16 bytes per lane) and reduced into a register -- no writes.  Compared with the same kernel on idx[r] = r (streaming).
// This is the roofline of the copy-free minibatch path: nothing that gathers rows through an index can go faster.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o random_rows random_rows.cu
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

template <int ROW_BYTES>
__global__ void __launch_bounds__(512) k_rows(const uint4* __restrict__ X, const int32_t* __restrict__ idx, int64_t B, uint32_t* __restrict__ sink) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  uint32_t acc = 0;
  constexpr int PER_LANE = ROW_BYTES / 512;  // uint4 loads per lane per row
  for (int64_t r = warp; r < B; r += nwarps * 4) {  // four rows in flight per warp
    uint4 v[4][PER_LANE];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int64_t rr = r + u * nwarps;
      if (rr < B) {
        const uint4* p = X + (int64_t)idx[rr] * (ROW_BYTES / 16) + lane;
#pragma unroll
        for (int j = 0; j < PER_LANE; ++j) v[u][j] = __ldcs(p + 32 * j);
      } else {
#pragma unroll
        for (int j = 0; j < PER_LANE; ++j) v[u][j] = make_uint4(0, 0, 0, 0);
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u)
#pragma unroll
      for (int j = 0; j < PER_LANE; ++j) acc ^= v[u][j].x ^ v[u][j].y ^ v[u][j].z ^ v[u][j].w;
  }
  if (acc == 0x12345678u) sink[0] = acc;
}

template <int ROW_BYTES>
static double run(const uint4* X, const int32_t* idx, int64_t B, uint32_t* sink, int iters) {
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  for (int i = 0; i < 3; ++i) k_rows<ROW_BYTES><<<sms * 4, 512>>>(X, idx, B, sink);
  cudaEventRecord(e0);
  for (int i = 0; i < iters; ++i) k_rows<ROW_BYTES><<<sms * 4, 512>>>(X, idx, B, sink);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  return (double)B * ROW_BYTES * iters / (ms * 1e-3) / 1e9;  // GB/s
}

int main(int argc, char** argv) {
  const int64_t N = argc > 1 ? atoll(argv[1]) : 8000000;  // rows of 2 KB: 16 GB
  const int64_t B = argc > 2 ? atoll(argv[2]) : 4000000;  // 8 GB read per launch (>> L2)
  uint4* X;
  cudaMalloc(&X, (size_t)N * 2048);
  cudaMemset(X, 1, (size_t)N * 2048);
  std::vector<int32_t> h(B), s(B);
  uint64_t st = 88172645463325252ull;
  for (int64_t i = 0; i < B; ++i) {
    st ^= st << 13; st ^= st >> 7; st ^= st << 17;
    h[i] = (int32_t)(st % (uint64_t)N);
    s[i] = (int32_t)(i % N);
  }
  int32_t *dr, *ds;
  uint32_t* sink;
  cudaMalloc(&dr, B * 4);
  cudaMalloc(&ds, B * 4);
  cudaMalloc(&sink, 4);
  cudaMemcpy(dr, h.data(), B * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(ds, s.data(), B * 4, cudaMemcpyHostToDevice);
  printf("{\"rows\": %lld, \"batch\": %lld,\n", (long long)N, (long long)B);
  printf(" \"streaming_2KB_rows_GBps\": %.1f,\n", run<2048>(X, ds, B, sink, 10));
  printf(" \"random_2KB_rows_GBps\": %.1f,\n", run<2048>(X, dr, B, sink, 10));
  // the planes layout: two 1 KB rows per datum -> 2N rows of 1 KB, index pairs (i, i + N)
  std::vector<int32_t> hp(2 * B);
  for (int64_t i = 0; i < B; ++i) { hp[2 * i] = h[i]; hp[2 * i + 1] = (int32_t)(h[i] + N); }
  int32_t* dp;
  cudaMalloc(&dp, 2 * B * 4);
  cudaMemcpy(dp, hp.data(), 2 * B * 4, cudaMemcpyHostToDevice);
  printf(" \"random_1KB_plane_rows_GBps\": %.1f,\n", run<1024>(X, dp, 2 * B, sink, 10));
  printf(" \"error\": \"%s\"}\n", cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}
