// Stage-boundary transport over NVLink without NCCL: CUDA-IPC mailboxes + peer copies + device-side sequence flags.
//
// One process per GPU (as the reference launches, README.md:118).  The receiving stage owns a mailbox (slots + one
// "ready" flag per slot) allocated with cudaMalloc and exported with cudaIpcGetMemHandle; the sending stage maps it
// and pushes a micro-batch's boundary tuple with cudaMemcpyPeerAsync, then publishes the slot's sequence number with a
// one-thread kernel (st.release.sys).  The receiver's compute stream waits for that number with a one-thread polling
// kernel (ld.acquire.sys), so no host thread ever blocks on the data path and there is no rendezvous: the copy engine
// moves the bytes while both SM arrays keep computing.  Flow control runs the other way through a "free" flag array
// owned by the sender.
//
// Replaces DeepSpeed's _exec_send/recv_activations/_grads over torch.distributed p2p (emitted by the schedule at
// reference utils/patches.py:134-143).
#include <string.h>

#include "host_util.h"

namespace dpipe {

__global__ void flag_write_kernel(unsigned long long* flag, unsigned long long value) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(flag), "l"(value) : "memory");
}

__global__ void flag_wait_kernel(const unsigned long long* flag, unsigned long long value, long long timeout_cycles) {
  const long long t0 = clock64();
  unsigned long long v;
  while (true) {
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(flag) : "memory");
    if (v >= value) break;
    if (timeout_cycles > 0 && clock64() - t0 > timeout_cycles) {
      printf("dpipe: stage-link flag wait timed out (have %llu, want %llu)\n", v, value);
      __trap();
    }
    __nanosleep(200);
  }
}

}  // namespace dpipe

using namespace dpipe;

extern "C" int dpipe_ipc_alloc(int64_t bytes, void** dev_ptr, unsigned char* handle64) {
  if (bytes <= 0 || !dev_ptr || !handle64) return fail(DPIPE_EINVAL, "dpipe_ipc_alloc: bad arguments");
  void* p = nullptr;
  DPIPE_CUDA_CHECK(cudaMalloc(&p, (size_t)bytes));
  DPIPE_CUDA_CHECK(cudaMemset(p, 0, (size_t)bytes));
  cudaIpcMemHandle_t h;
  cudaError_t e = cudaIpcGetMemHandle(&h, p);
  if (e != cudaSuccess) {
    cudaFree(p);
    return fail(DPIPE_ECUDA, "cudaIpcGetMemHandle: %s", cudaGetErrorString(e));
  }
  static_assert(sizeof(h) == 64, "cudaIpcMemHandle_t is 64 bytes");
  memcpy(handle64, &h, 64);
  DPIPE_CUDA_CHECK(cudaDeviceSynchronize());
  *dev_ptr = p;
  return 0;
}

extern "C" int dpipe_ipc_open(const unsigned char* handle64, void** dev_ptr) {
  if (!handle64 || !dev_ptr) return fail(DPIPE_EINVAL, "dpipe_ipc_open: bad arguments");
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, 64);
  void* p = nullptr;
  DPIPE_CUDA_CHECK(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
  *dev_ptr = p;
  return 0;
}

extern "C" int dpipe_ipc_close(void* dev_ptr) {
  if (!dev_ptr) return 0;
  DPIPE_CUDA_CHECK(cudaIpcCloseMemHandle(dev_ptr));
  return 0;
}

extern "C" int dpipe_ipc_free(void* dev_ptr) {
  if (!dev_ptr) return 0;
  DPIPE_CUDA_CHECK(cudaFree(dev_ptr));
  return 0;
}

extern "C" int dpipe_peer_copy(void* dst, int dst_device, const void* src, int src_device, int64_t bytes, void* stream) {
  if (!dst || !src || bytes < 0) return fail(DPIPE_EINVAL, "dpipe_peer_copy: bad arguments");
  if (bytes == 0) return 0;
  DPIPE_CUDA_CHECK(cudaMemcpyPeerAsync(dst, dst_device, src, src_device, (size_t)bytes, (cudaStream_t)stream));
  return 0;
}

extern "C" int dpipe_flag_write(void* flag, uint64_t value, void* stream) {
  if (!flag) return fail(DPIPE_EINVAL, "dpipe_flag_write: null flag");
  flag_write_kernel<<<1, 1, 0, (cudaStream_t)stream>>>((unsigned long long*)flag, (unsigned long long)value);
  DPIPE_CUDA_CHECK(cudaGetLastError());
  return 0;
}

extern "C" int dpipe_flag_wait_geq(const void* flag, uint64_t value, double timeout_s, void* stream) {
  if (!flag) return fail(DPIPE_EINVAL, "dpipe_flag_wait_geq: null flag");
  const long long cycles = timeout_s > 0 ? (long long)(timeout_s * 1.9e9) : 0;
  flag_wait_kernel<<<1, 1, 0, (cudaStream_t)stream>>>((const unsigned long long*)flag, (unsigned long long)value, cycles);
  DPIPE_CUDA_CHECK(cudaGetLastError());
  return 0;
}
