// The exchange step of the path as ONE kernel over NVLink peer memory instead of an NCCL call:
// a one-shot all-reduce (sum) of the packed [loss | grads] buffer.  The message is tiny (4.1 KB at D = 512), so the step
// is latency-bound: every rank pushes its vector straight into a slot of every peer's buffer with plain stores over
// NVLink (the buffers are cudaMalloc'ed here and mapped into the other processes with CUDA IPC), raises a flag, waits
// for the peers' flags and sums the slots in rank order (deterministic, identical on every rank).
//
// Buffer of one rank:  slots[2][world][n_pad] floats  |  flags[2][world] uint64 (the epoch that filled the slot)
// Two slot sets alternate by epoch parity.  A rank can start epoch e+1 (writing set (e+1)&1) while a peer still reads set
// e&1, and it cannot reach epoch e+2 before every peer has signalled e+1, i.e. has finished reading set e&1.
#include <cstdio>
#include <cstring>

#include "bjx_common.cuh"

namespace bjx {

struct P2PPeers {
  float* buf[8];
};

__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}

// A dead or hung peer must not hang this rank forever: the flag wait gives up after g_timeout_ns (default 30 s) and traps,
// which surfaces on the host as a launch failure of this stream at the next synchronisation (BJX_ERR_CUDA from then on).
__device__ unsigned long long g_p2p_timeout_ns = 30ull * 1000ull * 1000ull * 1000ull;

__device__ __forceinline__ unsigned long long globaltimer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}

// push -> flag -> wait -> rank-ordered sum, by ONE block of 1024 threads; result in out[0..n) on every rank
__device__ __forceinline__ void allreduce_p2p_block(float* __restrict__ out, int64_t n, int64_t n_pad, const P2PPeers& peers, int rank,
                                                    int world, unsigned long long epoch) {
  const int tid = threadIdx.x;
  const int set = (int)(epoch & 1ull);
  const size_t slot_floats = (size_t)n_pad;
  // 1. push my vector into slot [set][rank] of every rank's buffer (peer stores over NVLink; my own copy too)
  for (int p = 0; p < world; ++p) {
    float* dst = peers.buf[p] + ((size_t)set * world + rank) * slot_floats;
    for (int64_t i = tid; i < n; i += 1024) dst[i] = out[i];
  }
  __threadfence_system();
  __syncthreads();
  // 2. tell every rank that my slot of this epoch is complete
  if (tid < world) {
    unsigned long long* flags = reinterpret_cast<unsigned long long*>(peers.buf[tid] + (size_t)2 * world * slot_floats);
    st_release_sys(flags + (size_t)set * world + rank, epoch);
  }
  // 3. wait for every rank's slot in MY buffer
  if (tid < world) {
    const unsigned long long* flags = reinterpret_cast<const unsigned long long*>(peers.buf[rank] + (size_t)2 * world * slot_floats);
    const unsigned long long t0 = globaltimer_ns();
    unsigned spins = 0;
    while (ld_acquire_sys(flags + (size_t)set * world + tid) != epoch) {
      if ((++spins & 0x3FFu) == 0 && globaltimer_ns() - t0 > g_p2p_timeout_ns) {
        printf("bjx: peer-memory all-reduce timed out on rank %d waiting for rank %d at epoch %llu\n", rank, tid, epoch);
        __trap();
      }
    }
  }
  __syncthreads();
  // 4. sum in rank order
  const float* mine = peers.buf[rank] + (size_t)set * world * slot_floats;
  for (int64_t i = tid; i < n; i += 1024) {
    float acc = 0.0f;
    for (int p = 0; p < world; ++p) acc += __ldcv(mine + (size_t)p * slot_floats + i);
    out[i] = acc;
  }
}

__global__ void __launch_bounds__(1024) k_allreduce_p2p(float* __restrict__ out, int64_t n, int64_t n_pad, P2PPeers peers, int rank,
                                                        int world, unsigned long long epoch) {
  allreduce_p2p_block(out, n, n_pad, peers, rank, world, epoch);
}

// The exchange step of a device-resident training step (bjx_train_steps_sharded): the epoch comes from the DEVICE step
// counter (epoch_base + *steps_done + 1), so a captured CUDA graph of whole training steps replays without updates.
__global__ void __launch_bounds__(1024) k_allreduce_p2p_dev(float* __restrict__ out, int64_t n, int64_t n_pad, P2PPeers peers,
                                                            int rank, int world, unsigned long long epoch_base,
                                                            const int64_t* __restrict__ steps_done) {
  allreduce_p2p_block(out, n, n_pad, peers, rank, world, epoch_base + (unsigned long long)(*steps_done) + 1ull);
}

static int64_t pad_floats(int64_t n) { return (n + 63) / 64 * 64; }

}  // namespace bjx

extern "C" {

size_t bjx_p2p_buffer_bytes(int64_t n, int32_t world) {
  if (n < 1 || world < 1 || world > 8) return 0;
  return (size_t)2 * world * bjx::pad_floats(n) * sizeof(float) + (size_t)2 * world * sizeof(unsigned long long);
}

// cudaMalloc'ed (IPC handles refer to whole allocations), zero-filled; handle_out receives the 64-byte cudaIpcMemHandle_t
int bjx_p2p_alloc(size_t bytes, void** ptr_out, void* handle_out_64_bytes) {
  using namespace bjx;
  BJX_REQUIRE(bytes > 0 && ptr_out && handle_out_64_bytes, "bad arguments");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
  void* p = nullptr;
  BJX_CUDA_TRY(cudaMalloc(&p, bytes));
  BJX_CUDA_TRY(cudaMemset(p, 0, bytes));
  BJX_CUDA_TRY(cudaDeviceSynchronize());
  cudaIpcMemHandle_t h;
  BJX_CUDA_TRY(cudaIpcGetMemHandle(&h, p));
  memcpy(handle_out_64_bytes, &h, sizeof(h));
  *ptr_out = p;
  return BJX_OK;
}

int bjx_p2p_open(const void* handle_64_bytes, void** ptr_out) {
  using namespace bjx;
  BJX_REQUIRE(handle_64_bytes && ptr_out, "null pointer");
  cudaIpcMemHandle_t h;
  memcpy(&h, handle_64_bytes, sizeof(h));
  BJX_CUDA_TRY(cudaIpcOpenMemHandle(ptr_out, h, cudaIpcMemLazyEnablePeerAccess));
  return BJX_OK;
}

int bjx_p2p_close(void* peer_ptr) {
  using namespace bjx;
  BJX_CUDA_TRY(cudaIpcCloseMemHandle(peer_ptr));
  return BJX_OK;
}

int bjx_p2p_free(void* own_ptr) {
  using namespace bjx;
  BJX_CUDA_TRY(cudaFree(own_ptr));
  return BJX_OK;
}

// out[0..n) <- sum over ranks, in place, enqueued on `stream`.  bufs_host[world]: bufs_host[rank] is this rank's own buffer,
// the others are the IPC-opened peers; epoch must be 1, 2, 3, ... in lockstep on every rank (one call per step).
int bjx_allreduce_packed_p2p(float* out, int64_t n, void* const* bufs_host, int32_t rank, int32_t world, uint64_t epoch, void* stream) {
  using namespace bjx;
  BJX_REQUIRE(out && bufs_host && n >= 1, "bad arguments");
  BJX_REQUIRE(world >= 1 && world <= 8 && rank >= 0 && rank < world, "world must be 1..8 and rank inside it");
  BJX_REQUIRE(epoch >= 1, "epochs start at 1");
  P2PPeers peers;
  for (int p = 0; p < 8; ++p) peers.buf[p] = p < world ? (float*)bufs_host[p] : nullptr;
  k_allreduce_p2p<<<1, 1024, 0, (cudaStream_t)stream>>>(out, n, pad_floats(n), peers, rank, world, (unsigned long long)epoch);
  BJX_LAUNCH_CHECK("k_allreduce_p2p");
  return BJX_OK;
}

// The same exchange with the epoch taken from a device-resident step counter: epoch = epoch_base + *steps_done + 1.
int bjx_allreduce_packed_p2p_dev(float* out, int64_t n, void* const* bufs_host, int32_t rank, int32_t world, uint64_t epoch_base,
                                 const int64_t* steps_done, void* stream) {
  using namespace bjx;
  BJX_REQUIRE(out && bufs_host && steps_done && n >= 1, "bad arguments");
  BJX_REQUIRE(world >= 1 && world <= 8 && rank >= 0 && rank < world, "world must be 1..8 and rank inside it");
  P2PPeers peers;
  for (int p = 0; p < 8; ++p) peers.buf[p] = p < world ? (float*)bufs_host[p] : nullptr;
  k_allreduce_p2p_dev<<<1, 1024, 0, (cudaStream_t)stream>>>(out, n, pad_floats(n), peers, rank, world, (unsigned long long)epoch_base,
                                                            steps_done);
  BJX_LAUNCH_CHECK("k_allreduce_p2p_dev");
  return BJX_OK;
}

// how long a rank waits for a peer's flag before it traps (0 restores the default of 30 s)
int bjx_p2p_set_timeout_ms(int64_t ms) {
  using namespace bjx;
  BJX_REQUIRE(ms >= 0, "timeout must be >= 0");
  const unsigned long long ns = ms == 0 ? 30ull * 1000ull * 1000ull * 1000ull : (unsigned long long)ms * 1000000ull;
  BJX_CUDA_TRY(cudaMemcpyToSymbol(g_p2p_timeout_ns, &ns, sizeof(ns)));
  return BJX_OK;
}

}  // extern "C"
