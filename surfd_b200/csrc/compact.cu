// compact.cu -- ordered stream compaction of a bit mask (positions of set bits, ascending).
// Used for the marching-cubes candidate list (raster order == the reference's scan order,
// _marching_cubes_lewiner_cy.pyx:1194-1206) and for the GridFiller / gradient query lists
// (boolean-mask indexing order of meshudf.py:176,199-203,295-296).
#include "common.cuh"

namespace surfd {

thread_local char g_last_error[512] = {0};
int64_t g_launch_count = 0;

namespace {

constexpr int kThreads = 256;

__device__ __forceinline__ int warp_incl_scan(int v) {
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    int n = __shfl_up_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) >= o) v += n;
  }
  return v;
}

// inclusive block scan for kThreads threads; returns inclusive value, total in *total
__device__ __forceinline__ int block_incl_scan(int v, int* total) {
  __shared__ int warp_sums[kThreads / 32];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  int inc = warp_incl_scan(v);
  if (lane == 31) warp_sums[wid] = inc;
  __syncthreads();
  if (wid == 0) {
    int s = lane < kThreads / 32 ? warp_sums[lane] : 0;
    s = warp_incl_scan(s);
    if (lane < kThreads / 32) warp_sums[lane] = s;
  }
  __syncthreads();
  const int base = wid > 0 ? warp_sums[wid - 1] : 0;
  *total = warp_sums[kThreads / 32 - 1];
  return inc + base;
}

__global__ void __launch_bounds__(kThreads) popc_blocks_kernel(const uint32_t* __restrict__ bits, int64_t n_words,
                                                               int32_t* __restrict__ block_counts) {
  const int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x;
  int c = i < n_words ? __popc(bits[i]) : 0;
  int total;
  block_incl_scan(c, &total);
  if (threadIdx.x == 0) block_counts[blockIdx.x] = total;
}

// single block: exclusive scan of block_counts[0..nb) in place; total -> *total
__global__ void __launch_bounds__(1024) scan_blocks_kernel(int32_t* __restrict__ counts, int nb, int64_t* __restrict__ total,
                                                           int64_t* __restrict__ total_host) {
  __shared__ int64_t part[1024];
  const int t = threadIdx.x;
  const int per = (nb + 1023) / 1024;
  const int lo = t * per, hi = min(nb, lo + per);
  int64_t s = 0;
  for (int i = lo; i < hi; ++i) s += counts[i];
  part[t] = s;
  __syncthreads();
  // Hillis-Steele over 1024 partials
  for (int o = 1; o < 1024; o <<= 1) {
    int64_t v = t >= o ? part[t - o] : 0;
    __syncthreads();
    part[t] += v;
    __syncthreads();
  }
  int64_t run = t > 0 ? part[t - 1] : 0;
  for (int i = lo; i < hi; ++i) {
    int c = counts[i];
    counts[i] = (int32_t)run;
    run += c;
  }
  if (t == 1023) { *total = part[1023]; *total_host = part[1023]; }
}

__global__ void __launch_bounds__(kThreads) scatter_bits_kernel(const uint32_t* __restrict__ bits, int64_t n_words,
                                                                const int32_t* __restrict__ block_offsets,
                                                                int32_t* __restrict__ out, int64_t cap) {
  const int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x;
  uint32_t w = i < n_words ? bits[i] : 0u;
  int c = __popc(w);
  int total;
  int inc = block_incl_scan(c, &total);
  int64_t pos = (int64_t)block_offsets[blockIdx.x] + (inc - c);
  const int32_t base = (int32_t)(i * 32);
  while (w) {
    int b = __ffs(w) - 1;
    if (pos < cap) out[pos] = base + b;   // the caller compares the total with cap to detect truncation
    ++pos;
    w &= w - 1;
  }
}

// set bits before each 32-bit word (rank structure: rank(i) = prefix[i >> 5] + popc(bits[i >> 5] & low_mask(i)))
__global__ void __launch_bounds__(kThreads) word_prefix_kernel(const uint32_t* __restrict__ bits, int64_t n_words,
                                                               const int32_t* __restrict__ block_offsets,
                                                               int32_t* __restrict__ prefix) {
  const int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x;
  const int c = i < n_words ? __popc(bits[i]) : 0;
  int total;
  const int inc = block_incl_scan(c, &total);
  if (i < n_words) prefix[i] = block_offsets[blockIdx.x] + (inc - c);
}

}  // namespace

int Compactor::init() {
  SURFD_CUDA(cudaMalloc(&d_total, sizeof(int64_t)));
  // The host reads the total from mapped pinned memory the scan kernel writes directly: no device->host copy is queued, so
  // the read-back cannot sit behind another stream's pending copy in a copy-engine queue.
  SURFD_CUDA(cudaHostAlloc(&h_total, sizeof(int64_t), cudaHostAllocMapped));
  SURFD_CUDA(cudaHostGetDevicePointer(&h_total_dev, h_total, 0));
  *h_total = 0;
  return 0;
}

void Compactor::destroy() {
  block_counts.release();
  if (d_total) cudaFree(d_total);
  if (h_total) cudaFreeHost(h_total);
  d_total = nullptr; h_total = nullptr;
}

int Compactor::count(const uint32_t* bits, int64_t n_words, cudaStream_t st) {
  const int64_t nb = cdiv(n_words, kThreads);
  SURFD_REQUIRE(nb < (1ll << 30), "compaction input too large");
  SURFD_TRY(block_counts.reserve((size_t)(nb + 1) * sizeof(int32_t)));
  if (nb == 0) {
    SURFD_CUDA(cudaMemsetAsync(d_total, 0, sizeof(int64_t), st));
    SURFD_CUDA(cudaStreamSynchronize(st));
    *h_total = 0;
    return 0;
  }
  popc_blocks_kernel<<<(unsigned)nb, kThreads, 0, st>>>(bits, n_words, block_counts.as<int32_t>());
  SURFD_CHECK_LAUNCH();
  scan_blocks_kernel<<<1, 1024, 0, st>>>(block_counts.as<int32_t>(), (int)nb, d_total, h_total_dev);
  SURFD_CHECK_LAUNCH();
  return 0;
}

int Compactor::scatter(const uint32_t* bits, int64_t n_words, int32_t* out_list, int64_t cap, cudaStream_t st) {
  const int64_t nb = cdiv(n_words, kThreads);
  if (nb == 0) return 0;
  scatter_bits_kernel<<<(unsigned)nb, kThreads, 0, st>>>(bits, n_words, block_counts.as<int32_t>(), out_list, cap);
  SURFD_CHECK_LAUNCH();
  return 0;
}

int Compactor::word_prefix(const uint32_t* bits, int64_t n_words, int32_t* prefix, cudaStream_t st) {
  const int64_t nb = cdiv(n_words, kThreads);
  if (nb == 0) return 0;
  word_prefix_kernel<<<(unsigned)nb, kThreads, 0, st>>>(bits, n_words, block_counts.as<int32_t>(), prefix);
  SURFD_CHECK_LAUNCH();
  return 0;
}

int Compactor::read_total(int64_t* total, cudaStream_t st) {
  SURFD_CUDA(cudaStreamSynchronize(st));
  *total = *reinterpret_cast<volatile int64_t*>(h_total);
  return 0;
}

}  // namespace surfd

extern "C" int surfd_version(void) { return 100; }
extern "C" const char* surfd_last_error(void) { return surfd::g_last_error; }
extern "C" int64_t surfd_launch_count(int reset) {
  int64_t v = surfd::g_launch_count;
  if (reset) surfd::g_launch_count = 0;
  return v;
}
