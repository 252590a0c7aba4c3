// Peer-memory plumbing for the row-partitioned propagation: buffers that every rank of one node can store into
// (CUDA IPC over NVLink / NVSwitch), and the two tiny kernels of the per-layer hand-shake.
//
// The exchange itself is NOT a separate collective: the SpMM epilogue (spmm.cu) stores every finished row straight
// into all peers' copies of the table while it computes the next rows, so the all-gather of SURVEY 5.8 / 8(e) rides
// under the gathers of the same kernel.  What is left per layer is  signal (my rows are out)  ->  wait (everyone's are
// in), two single-block kernels that are captured in the step graph like everything else.
#include <string.h>
#include "common.cuh"

namespace b200rec {

// flags[src] on every peer <- epoch  (after a system-scope fence: the SpMM's peer stores are ordered before it)
__global__ void peer_signal_kernel(unsigned long long* const* peer_flags, int n_peers, int my_rank, const int64_t* epoch_ptr,
                                   int64_t epoch_add) {
  if (threadIdx.x < n_peers) {
    __threadfence_system();
    const unsigned long long e = (unsigned long long)(*epoch_ptr + epoch_add);
    volatile unsigned long long* f = peer_flags[threadIdx.x] + my_rank;
    *f = e;
    __threadfence_system();
  }
}
// spin until flags[p] >= epoch for every p (bounded: a lost peer must abort the launch, not hang the GPU)
__global__ void peer_wait_kernel(const unsigned long long* my_flags, int n_peers, int64_t* epoch_ptr, int64_t epoch_add) {
  const unsigned long long e = (unsigned long long)(*epoch_ptr + epoch_add);
  __syncthreads();
  if (threadIdx.x < n_peers) {
    const volatile unsigned long long* f = my_flags + threadIdx.x;
    unsigned long long spins = 0;
    while (*f < e) {
      if (++spins > (1ull << 31)) __trap();
    }
    __threadfence_system();
  }
  __syncthreads();
  if (threadIdx.x == 0) *epoch_ptr = (int64_t)e;  // the hand-shake advances the epoch: the next one waits for e + epoch_add
}

}  // namespace b200rec
using namespace b200rec;

extern "C" int b200rec_peer_alloc(int64_t bytes, void** ptr_out, uint8_t* handle_out /*HOST [64]*/) {
  B2_REQUIRE(bytes > 0 && ptr_out && handle_out, "bad argument");
  void* p = nullptr;
  B2_CUDA(cudaMalloc(&p, (size_t)bytes));
  B2_CUDA(cudaMemset(p, 0, (size_t)bytes));
  cudaIpcMemHandle_t h;
  B2_CUDA(cudaIpcGetMemHandle(&h, p));
  static_assert(sizeof(h) == 64, "cudaIpcMemHandle_t is 64 bytes");
  memcpy(handle_out, &h, 64);
  *ptr_out = p;
  return 0;
}
extern "C" int b200rec_peer_open(const uint8_t* handle /*HOST [64]*/, void** ptr_out) {
  B2_REQUIRE(handle && ptr_out, "bad argument");
  cudaIpcMemHandle_t h;
  memcpy(&h, handle, 64);
  B2_CUDA(cudaIpcOpenMemHandle(ptr_out, h, cudaIpcMemLazyEnablePeerAccess));
  return 0;
}
extern "C" int b200rec_peer_close(void* ptr) {
  B2_CUDA(cudaIpcCloseMemHandle(ptr));
  return 0;
}
extern "C" int b200rec_peer_free(void* ptr) {
  B2_CUDA(cudaFree(ptr));
  return 0;
}
extern "C" int b200rec_peer_signal(void* const* peer_flags /*device array [n_peers] of flag arrays*/, int32_t n_peers,
                                   int32_t my_rank, const int64_t* epoch /*device*/, int64_t epoch_add, void* stream) {
  B2_REQUIRE(peer_flags && epoch && n_peers > 0 && n_peers <= 32, "bad argument");
  peer_signal_kernel<<<1, 32, 0, (cudaStream_t)stream>>>((unsigned long long* const*)peer_flags, n_peers, my_rank, epoch, epoch_add);
  B2_LAUNCHED();
  return 0;
}
extern "C" int b200rec_peer_wait(const void* my_flags, int32_t n_peers, int64_t* epoch, int64_t epoch_add, void* stream) {
  B2_REQUIRE(my_flags && epoch && n_peers > 0 && n_peers <= 32, "bad argument");
  peer_wait_kernel<<<1, 32, 0, (cudaStream_t)stream>>>((const unsigned long long*)my_flags, n_peers, epoch, epoch_add);
  B2_LAUNCHED();
  return 0;
}
