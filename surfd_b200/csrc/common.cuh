// common.cuh -- shared host-side helpers for the surfd_b200 CUDA library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/surfd_b200.h"

namespace surfd {

extern thread_local char g_last_error[512];
extern int64_t g_launch_count;

inline int set_error(int code, const char* what, const char* file, int line) {
  snprintf(g_last_error, sizeof(g_last_error), "%s (%s:%d)", what, file, line);
  return code;
}

#define SURFD_CUDA(expr)                                                              \
  do {                                                                                \
    cudaError_t _e = (expr);                                                          \
    if (_e != cudaSuccess) {                                                          \
      return ::surfd::set_error(-(int)_e, cudaGetErrorString(_e), __FILE__, __LINE__); \
    }                                                                                 \
  } while (0)

#define SURFD_CHECK_LAUNCH()                                                          \
  do {                                                                                \
    ++::surfd::g_launch_count;                                                        \
    cudaError_t _e = cudaGetLastError();                                              \
    if (_e != cudaSuccess) {                                                          \
      return ::surfd::set_error(-(int)_e, cudaGetErrorString(_e), __FILE__, __LINE__); \
    }                                                                                 \
  } while (0)

#define SURFD_REQUIRE(cond, msg)                                                      \
  do {                                                                                \
    if (!(cond)) return ::surfd::set_error(SURFD_BAD_ARGUMENT, msg, __FILE__, __LINE__); \
  } while (0)

#define SURFD_TRY(expr)          \
  do {                           \
    int _s = (expr);             \
    if (_s != 0) return _s;      \
  } while (0)

inline int64_t cdiv(int64_t a, int64_t b) { return (a + b - 1) / b; }

// Grow-only device buffer owned by a handle.
struct DevBuf {
  void* p = nullptr;
  size_t bytes = 0;
  // Grow-only, with 25% head-room rounded to 2 MiB: sizes that creep up from shape to shape (query lists) must not
  // re-allocate every time -- cudaFree() is device-synchronising and would wait for the marching-cubes replays that
  // run on other streams.
  int reserve(size_t need) {
    if (need <= bytes) return 0;
    size_t want = need + need / 4;
    const size_t gran = (size_t)2 << 20;
    want = (want + gran - 1) / gran * gran;
    if (p) cudaFree(p);
    p = nullptr; bytes = 0;
    cudaError_t e = cudaMalloc(&p, want);
    if (e != cudaSuccess) {
      e = cudaMalloc(&p, need);   // fall back to the exact size when memory is tight
      want = need;
    }
    if (e != cudaSuccess) return set_error(-(int)e, cudaGetErrorString(e), __FILE__, __LINE__);
    bytes = want;
    return 0;
  }
  void release() { if (p) cudaFree(p); p = nullptr; bytes = 0; }
  template <typename T> T* as() const { return reinterpret_cast<T*>(p); }
};

// ---- ordered bit compaction (compact.cu): positions of set bits, ascending -------------------
struct Compactor {
  DevBuf block_counts;   // int32 per block (+1)
  int64_t* d_total = nullptr;   // device scalar
  int64_t* h_total = nullptr;   // mapped pinned host scalar (written by the scan kernel)
  int64_t* h_total_dev = nullptr;   // its device-side alias
  int init();
  void destroy();
  // count(): per-block popcounts + exclusive offsets + total (device scalar d_total).
  // scatter(): must follow count() on the same bits; writes the first `cap` positions (total > cap = truncated).
  int count(const uint32_t* bits, int64_t n_words, cudaStream_t st);
  int scatter(const uint32_t* bits, int64_t n_words, int32_t* out_list, int64_t cap, cudaStream_t st);
  // word_prefix(): must follow count() on the same bits; number of set bits before each word
  int word_prefix(const uint32_t* bits, int64_t n_words, int32_t* prefix, cudaStream_t st);
  // copies d_total to host and synchronises the stream
  int read_total(int64_t* total, cudaStream_t st);
};

}  // namespace surfd
