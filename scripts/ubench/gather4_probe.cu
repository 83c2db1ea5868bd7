// Probe of cp.async.bulk.tensor.2d ... tile::gather4 on sm_100a: which box shape the tensor map needs, how many bytes the
// mbarrier sees, and where the four rows land under the 128-byte swizzle when the destination is 512 bytes into a 1024-byte
// swizzle atom.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o gather4_probe gather4_probe.cu -lcuda
#include <cuda.h>
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <vector>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void probe(const __grid_constant__ CUtensorMap map, int col0, int r0, int r1, int r2, int r3, int dst_off, uint32_t expect,
                      uint16_t* out, int* status) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
  __shared__ unsigned long long bar;
  for (int i = threadIdx.x; i < 4096; i += blockDim.x) reinterpret_cast<uint16_t*>(smem)[i] = 0xFFFF;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;");
  }
  __syncthreads();
  asm volatile("fence.proxy.async.shared::cta;");
  if (threadIdx.x == 0) {
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n\t}" ::"r"(smem_u32(&bar)), "r"(expect) : "memory");
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile::gather4.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];" ::"r"(
            smem_u32(smem + dst_off)),
        "l"(&map), "r"(smem_u32(&bar)), "r"(col0), "r"(r0), "r"(r1), "r"(r2), "r"(r3)
        : "memory");
    const long long t0 = clock64();
    int done = 0;
    while (!done && clock64() - t0 < 200000000ll) {
      uint32_t ok;
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(&bar)) : "memory");
      done = (int)ok;
    }
    status[0] = done;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 4096; i += blockDim.x) out[i] = reinterpret_cast<uint16_t*>(smem)[i];
}

int main() {
  const int R = 4096, C = 128;
  std::vector<uint16_t> h((size_t)R * C);
  for (int r = 0; r < R; ++r)
    for (int c = 0; c < C; ++c) h[(size_t)r * C + c] = (uint16_t)(((r & 0x3FF) << 6) | (c & 63));  // row id (10 bits) | column in block
  uint16_t* d;
  cudaMalloc(&d, h.size() * 2);
  cudaMemcpy(d, h.data(), h.size() * 2, cudaMemcpyHostToDevice);
  uint16_t* out;
  int* status;
  cudaMalloc(&out, 8192);
  cudaMalloc(&status, 4);
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 16384);
  const int rows[4] = {5, 100, 7, 999};
  for (int box_rows : {1, 4}) {
    CUtensorMap m;
    cuuint64_t gdim[2] = {(cuuint64_t)C, (cuuint64_t)R};
    cuuint64_t gstr[1] = {(cuuint64_t)C * 2};
    cuuint32_t box[2] = {64, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult rc = cuTensorMapEncodeTiled(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, d, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                         CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("box_rows=%d encode rc=%d\n", box_rows, (int)rc);
    if (rc != CUDA_SUCCESS) continue;
    for (int dst_off : {0, 512}) {
      for (uint32_t expect : {512u, 2048u}) {
        if (box_rows == 1 && expect == 2048u) continue;
        cudaMemset(status, 0xFF, 4);
        probe<<<1, 128, 16384>>>(m, 64, rows[0], rows[1], rows[2], rows[3], dst_off, expect, out, status);
        cudaError_t e = cudaDeviceSynchronize();
        int st = -1;
        std::vector<uint16_t> o(4096);
        if (e == cudaSuccess) {
          cudaMemcpy(&st, status, 4, cudaMemcpyDeviceToHost);
          cudaMemcpy(o.data(), out, 8192, cudaMemcpyDeviceToHost);
        }
        printf(" box_rows=%d dst_off=%d expect_tx=%u -> %s, barrier completed=%d\n", box_rows, dst_off, expect, cudaGetErrorString(e), st);
        if (e != cudaSuccess) return 1;
        // which source (row, 16-byte chunk) sits in each 16-byte chunk of the first 2 KB of the slot
        for (int line = 0; line < 16; ++line) {
          bool any = false;
          for (int ch = 0; ch < 8; ++ch) any |= o[line * 64 + ch * 8] != 0xFFFF;
          if (!any) continue;
          printf("   smem row %2d:", line);
          for (int ch = 0; ch < 8; ++ch) {
            const uint16_t v = o[line * 64 + ch * 8];
            if (v == 0xFFFF) printf("  ----  ");
            else printf(" r%03d:c%d", v >> 6, (v & 63) / 8);
          }
          printf("\n");
        }
      }
    }
  }
  return 0;
}
