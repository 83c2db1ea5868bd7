// Microbenchmark: cycles per tcgen05.mma (kind::f16, bf16) for the operand shapes/layouts of bjx_lik_tc.cu, SS mode.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o umma_rate umma_rate.cu ; run: ./umma_rate
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)(sbo >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
__host__ __device__ constexpr uint32_t make_idesc(int M, int N, int am, int bm) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)am << 15) | ((uint32_t)bm << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void umma(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}

// mode 0: A K-major (rows x 128B blocks), B K-major; mode 1: both MN-major
__device__ __forceinline__ void umma_ts(uint32_t d, uint32_t a_tmem, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d), "r"(a_tmem), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ uint32_t elect_one() {
  uint32_t pred = 0;
  asm volatile("{\n\t.reg .b32 rx;\n\t.reg .pred px;\n\telect.sync rx|px, 0xffffffff;\n\t@px mov.s32 %0, 1;\n\t}" : "+r"(pred));
  return pred;
}
template <int NACC, int ELECT, int mode>
__global__ void __launch_bounds__(128, 1) k(int M, int N, int iters, long long* out) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
  __shared__ unsigned long long bar;
  __shared__ uint32_t tmem_base;
  for (int i = threadIdx.x; i < 160 * 1024 / 4; i += blockDim.x) ((uint32_t*)smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;");
  }
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("fence.proxy.async.shared::cta;");
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t tmem = tmem_base;
  if (ELECT ? (threadIdx.x < 32) : (threadIdx.x == 0)) {
    const uint32_t idesc = mode == 2 ? make_idesc(M, N, 0, 0) : mode == 3 ? make_idesc(M, N, 1, 0) : make_idesc(M, N, mode, mode);
    const uint32_t a0 = smem_u32(smem), b0 = smem_u32(smem) + 128 * 1024;
    // 16 distinct operand positions, descriptors hoisted out of the timed loop (constant offsets -> immediate adds)
    uint64_t da0, db0;
    uint32_t astep_blk, astep_k;
    if (mode == 0 || mode == 2) { da0 = make_desc(a0, 16, 1024); db0 = make_desc(b0, 16, 1024); astep_blk = 16384 >> 4; astep_k = 32 >> 4; }
    else if (mode == 3) { da0 = make_desc(a0, 16384, 1024); db0 = make_desc(b0, 16, 1024); astep_blk = 16384 >> 4; astep_k = 2048 >> 4; }
    else           { da0 = make_desc(a0, 8192, 1024); db0 = make_desc(b0, 8192, 1024); astep_blk = 16384 >> 4; astep_k = 2048 >> 4; }
    long long t0 = clock64();
    for (int i = 0; i < iters; i += 16) {
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const uint64_t da = da0 + (uint64_t)((j / 4) * astep_blk + (j % 4) * astep_k);
        const uint64_t db = db0 + (uint64_t)((j % 4) * astep_k);
        if (mode == 2) { if (elect_one()) umma_ts(tmem + 256 + (j % 2) * 64, tmem + (j / 4) * 32 + (j % 4) * 8, da, idesc, 1); __syncwarp(); continue; }
        if (mode == 3) { if (elect_one()) umma(tmem + 384 + (j / 4) * 16, da, db0 + (uint64_t)((j % 4) * 2), idesc, 1); __syncwarp(); continue; }
        if (ELECT) { if (elect_one()) umma(tmem + (j % NACC) * 256, da, db, idesc, 1); __syncwarp(); }
        else umma(tmem + (j % NACC) * 256, da, db, idesc, 1);
      }
    }
    if (!ELECT || elect_one()) asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    long long t1 = clock64();
    asm volatile("{\n\t.reg .pred p;\n\tW: mbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n\t@p bra D;\n\tbra W;\n\tD:\n\t}" ::"r"(smem_u32(&bar)) : "memory");
    long long t2 = clock64();
    if (blockIdx.x == 0 && threadIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512));
}

int main() {
  long long* out; cudaMallocManaged(&out, 16);
  cudaFuncSetAttribute(k<1,1,0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  cudaFuncSetAttribute(k<1,1,1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  cudaFuncSetAttribute(k<1,1,2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  cudaFuncSetAttribute(k<1,1,3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  struct { int M, N, mode; const char* name; } cases[] = {
    {64, 64, 0, "G1 x1*[t1;t2]  M=64  N=64  K-major"}, {64, 32, 0, "G1 x2*t1       M=64  N=32  K-major"},
    {128, 64, 0, "             M=128 N=64  K-major"}, {128, 32, 0, "             M=128 N=32  K-major"},
    {128, 128, 0, "             M=128 N=128 K-major"}, {128, 256, 0, "             M=128 N=256 K-major"},
    {128, 64, 1, "G2 x1^T*[e1;e2] M=128 N=64 MN-major"}, {128, 32, 1, "G2 x2^T*e1    M=128 N=32 MN-major"},
    {128, 64, 2, "T-G1 theta[tmem]*x  M=128 N=64 TS"}, {128, 128, 2, "T-G1 theta[tmem]*[x1;x2] M=128 N=128 TS"},
    {128, 32, 3, "T-G2 x^T(MN)*e(K)   M=128 N=32"}, {128, 16, 3, "T-G2 x^T(MN)*e(K)   M=128 N=16"}, {64, 64, 2, "     theta[tmem]*x  M=64 N=64 TS"},
  };
  for (int grid : {148}) {
    printf("grid=%d CTAs (1 per SM)\n", grid);
    for (auto& c : cases) {
      const int iters = 2048;
      for (int nacc : {1}) {
      switch (c.mode) {
        case 0: k<1,1,0><<<grid, 128, 200 * 1024>>>(c.M, c.N, iters, out); break;
        case 1: k<1,1,1><<<grid, 128, 200 * 1024>>>(c.M, c.N, iters, out); break;
        case 2: k<1,1,2><<<grid, 128, 200 * 1024>>>(c.M, c.N, iters, out); break;
        default: k<1,1,3><<<grid, 128, 200 * 1024>>>(c.M, c.N, iters, out); break;
      }
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("%s: %s\n", c.name, cudaGetErrorString(e)); return 1; }
      printf("  %-38s elect=%d issue %.1f cyc/mma, complete %.1f cyc/mma  (%.0f MAC/cyc/SM)\n", c.name, 1, (double)out[0] / iters, (double)out[1] / iters,
             (double)c.M * c.N * 16 / ((double)out[1] / iters));
      }
    }
  }
  return 0;
}
