// Backend shim.  The product is the CUDA build (nvcc, sm_100a).  Defining
// PGBP_HOST_EMUL compiles the very same kernel bodies as plain C++ loops; that
// build exists ONLY so the CPU-side test-suite (`pytest -m "not gpu"`) can
// exercise the plan compiler, index tables and ABI without a GPU.  The Python
// package never loads it (see _lib.py: it refuses to run without the CUDA .so).
#pragma once
#include <atomic>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>

#ifdef PGBP_HOST_EMUL
#define PGBP_HD inline
#define PGBP_D inline
typedef void* pgbp_stream_t;
#else
#include <cuda_runtime.h>
#define PGBP_HD __host__ __device__ __forceinline__
#define PGBP_D __device__ __forceinline__
typedef cudaStream_t pgbp_stream_t;
#endif

namespace pgbp {

void set_error(const std::string& msg);
#define PGBP_FAIL(code, ...)                            \
  do {                                                  \
    char _b[512];                                       \
    snprintf(_b, sizeof _b, __VA_ARGS__);               \
    ::pgbp::set_error(_b);                              \
    return (code);                                      \
  } while (0)

enum { PGBP_OK = 0, PGBP_EINVAL = -1, PGBP_ECUDA = -2, PGBP_ENOMEM = -3, PGBP_ESTATE = -4 };

#ifdef PGBP_HOST_EMUL
#define PGBP_CUDA(x) (x)
inline int dev_malloc(void** p, size_t n) {
  *p = n ? malloc(n) : nullptr;
  return (n && !*p) ? PGBP_ENOMEM : 0;
}
inline void dev_free(void* p) { free(p); }
inline int dev_memset(void* p, int v, size_t n, pgbp_stream_t) {
  if (n) memset(p, v, n);
  return 0;
}
inline int h2d(void* d, const void* h, size_t n, pgbp_stream_t) {
  if (n) memcpy(d, h, n);
  return 0;
}
inline int d2h(void* h, const void* d, size_t n, pgbp_stream_t) {
  if (n) memcpy(h, d, n);
  return 0;
}
inline int d2d(void* d, const void* s, size_t n, pgbp_stream_t) {
  if (n) memmove(d, s, n);
  return 0;
}
inline int stream_sync(pgbp_stream_t) { return 0; }
inline int set_device(int) { return 0; }
inline int check_launch(const char*) { return 0; }
#else
#define PGBP_CUDA(call)                                                                    \
  do {                                                                                     \
    cudaError_t _e = (call);                                                               \
    if (_e != cudaSuccess)                                                                 \
      PGBP_FAIL(::pgbp::PGBP_ECUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(_e), \
                __FILE__, __LINE__);                                                       \
  } while (0)
inline int dev_malloc(void** p, size_t n) {
  *p = nullptr;
  if (!n) return 0;
  PGBP_CUDA(cudaMalloc(p, n));
  return 0;
}
inline void dev_free(void* p) {
  if (p) cudaFree(p);
}
inline int dev_memset(void* p, int v, size_t n, pgbp_stream_t s) {
  if (n) PGBP_CUDA(cudaMemsetAsync(p, v, n, s));
  return 0;
}
inline int h2d(void* d, const void* h, size_t n, pgbp_stream_t s) {
  if (n) PGBP_CUDA(cudaMemcpyAsync(d, h, n, cudaMemcpyHostToDevice, s));
  return 0;
}
inline int d2h(void* h, const void* d, size_t n, pgbp_stream_t s) {
  if (n) PGBP_CUDA(cudaMemcpyAsync(h, d, n, cudaMemcpyDeviceToHost, s));
  return 0;
}
inline int d2d(void* d, const void* s_, size_t n, pgbp_stream_t s) {
  if (n) PGBP_CUDA(cudaMemcpyAsync(d, s_, n, cudaMemcpyDeviceToDevice, s));
  return 0;
}
inline int stream_sync(pgbp_stream_t s) {
  PGBP_CUDA(cudaStreamSynchronize(s));
  return 0;
}
inline int set_device(int d) {
  PGBP_CUDA(cudaSetDevice(d));
  return 0;
}
// cudaFuncSetAttribute applies to the CURRENT device only: one flag per (kernel instantiation, device), so that a
// process that drives several GPUs raises the dynamic shared-memory limit on each of them.  Returns true the first
// time it is called on the current device.
struct AttrOnce {
  std::atomic<uint64_t> mask{0};
  bool first() {
    int dev = 0;
    cudaGetDevice(&dev);
    const uint64_t bit = 1ull << (dev & 63);
    return (mask.fetch_or(bit, std::memory_order_relaxed) & bit) == 0;
  }
};
inline int check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) PGBP_FAIL(PGBP_ECUDA, "launch of %s failed: %s", what, cudaGetErrorString(e));
  return 0;
}
#endif

#define PGBP_TRY(x)          \
  do {                       \
    int _r = (x);            \
    if (_r != 0) return _r;  \
  } while (0)

// first-failure-wins status update (smaller word = earlier message in the
// reference's sequential order)
PGBP_HD void status_fail(int32_t* status, int64_t e, int32_t word) {
#if defined(__CUDA_ARCH__)
  int old = atomicCAS(&status[e], 0, word);
  if (old != 0) atomicMin(&status[e], word);
#else
  if (status[e] == 0 || word < status[e]) status[e] = word;
#endif
}

}  // namespace pgbp
