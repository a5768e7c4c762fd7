// mc.cu -- MeshUDF marching cubes on device: HBM-bound candidate classification + ordered replay.
// Compiled with -fmad=false: the replay must reproduce the reference's unfused IEEE arithmetic
// (meshudf/_marching_cubes_lewiner_cy.pyx built with gcc for x86-64 has no FMA contraction).
//
// Stage 1 (classify_kernel): the reference's raster scan evaluates avg_cube < 1.05*voxel and
// max_cube <= 1.74*voxel for all (N-1)^3 cubes on one CPU thread (pyx:1157-1158, 1194-1218, 1825-1841).
// Here it is one pass over the udf lattice: each warp owns 32 consecutive x positions of a (y, y+1) row
// pair and marches in z keeping the previous plane in registers, so every lattice value is loaded ~2x
// from L1/L2 and ~1x from HBM (algorithmic traffic 4 B/voxel); the x+1 neighbour comes by warp shuffle.
// Output: 1 bit per lattice index (word = one ballot), then ordered compaction (compact.cu) gives the
// raster-sorted candidate list == the order in which the reference's scan meets the candidates.
// Stage 2 (replay_kernel): mc_core.h, O(surface).
#include "common.cuh"
#include <algorithm>
#include "mc_chain.h"

namespace surfd {

constexpr int kZChunk = 16;

// grid: x = ceil(N/32) * (N-1) rows... one warp per (x-segment, y), blockIdx.y = z chunk
__global__ void __launch_bounds__(256)
classify_kernel(const float* __restrict__ im, int N, float avg_t, float max_t, uint32_t* __restrict__ bits, int aligned) {
  const int warps_per_block = blockDim.x >> 5;
  const int lane = threadIdx.x & 31;
  const int wid = blockIdx.x * warps_per_block + (threadIdx.x >> 5);
  const int xsegs = (N + 31) >> 5;
  const int y = wid / xsegs;
  const int x0 = (wid % xsegs) << 5;
  if (y >= N - 1) return;
  const int z_begin = blockIdx.y * kZChunk;
  const int z_end = min(N - 1, z_begin + kZChunk);  // cubes z in [z_begin, z_end)
  if (z_begin >= N - 1) return;
  const int x = x0 + lane;
  const bool in_x = x < N;
  const bool has_next = (x0 + 32) < N;  // lane 31 needs lattice point x0+32
  const size_t plane = (size_t)N * N;

  // load one plane's row pair for this lane (+ lane 31's extra neighbour)
  float a, b, an, bn;  // a = im[z][y][x], b = im[z][y+1][x]; an/bn = values at x+1
  auto load_plane = [&](int z, float& pa, float& pb, float& pan, float& pbn) {
    const float* r0 = im + (size_t)z * plane + (size_t)y * N;
    const float* r1 = r0 + N;
    pa = in_x ? r0[x] : 0.f;
    pb = in_x ? r1[x] : 0.f;
    float ea = 0.f, eb = 0.f;
    if (lane == 31 && has_next) { ea = r0[x0 + 32]; eb = r1[x0 + 32]; }
    pan = __shfl_down_sync(0xffffffffu, pa, 1);
    pbn = __shfl_down_sync(0xffffffffu, pb, 1);
    if (lane == 31) { pan = ea; pbn = eb; }
  };
  load_plane(z_begin, a, b, an, bn);
  for (int z = z_begin; z < z_end; ++z) {
    float c, d, cn, dn;
    load_plane(z + 1, c, d, cn, dn);
    // reference corner order v1..v8 = (z,y,x) (z,y,x+1) (z,y+1,x+1) (z,y+1,x) (z+1,y,x) (z+1,y,x+1) (z+1,y+1,x+1) (z+1,y+1,x)
    float s = a + an;
    s = s + bn; s = s + b; s = s + c; s = s + cn; s = s + dn; s = s + d;
    const float avg = 0.125f * s;
    const float m = fmaxf(fmaxf(fmaxf(a, an), fmaxf(bn, b)), fmaxf(fmaxf(c, cn), fmaxf(dn, d)));
    const bool cand = (x < N - 1) && (avg < avg_t) && (m <= max_t);
    const uint32_t w = __ballot_sync(0xffffffffu, cand);
    if (lane == 0) {
      const size_t i0 = (size_t)z * plane + (size_t)y * N + x0;
      if (aligned) {
        bits[i0 >> 5] = w;
      } else if (w) {
        const int sh = (int)(i0 & 31);
        atomicOr(&bits[i0 >> 5], w << sh);
        if (sh) atomicOr(&bits[(i0 >> 5) + 1], w >> (32 - sh));
      }
    }
    a = c; b = d; an = cn; bn = dn;
  }
}

// ---- O(surface) replay (mc_chain.h): records -> one-warp chain -> parallel emission ----

// lattice vertices that are a corner of a candidate cube
__global__ void __launch_bounds__(256) mark_vertices_kernel(const int32_t* __restrict__ list, const int64_t* __restrict__ n_cand_dev,
                                                            int64_t cap, int N, uint32_t* __restrict__ vbits) {
  const int64_t n = min(*n_cand_dev, cap);
  for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += (int64_t)gridDim.x * blockDim.x) {
    const int64_t i = list[k];
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      const int64_t v = i + (int64_t)MC_CZ(c) * N * N + (int64_t)MC_CY(c) * N + MC_CX(c);
      atomicOr(&vbits[v >> 5], 1u << (v & 31));
    }
  }
}

__global__ void __launch_bounds__(128) build_records_kernel(const surfd_mccore::Chain* __restrict__ gp, const int64_t* __restrict__ n_cand_dev,
                                                            const int64_t* __restrict__ n_vtx_dev, int64_t cap_cand) {
  const int64_t n = *n_cand_dev;
  if (n > cap_cand || *n_vtx_dev > gp->cap_vtx) return;   // the chain reports MC_CAPACITY
  surfd_mccore::Chain g = *gp;
  for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += (int64_t)gridDim.x * blockDim.x)
    surfd_mccore::build_record(g, k);
}

// One warp per shape: all 32 lanes run replay_r() on a private copy of the counters (identical control flow, same-value
// stores) and split the two gathers of each visit.  The candidate / vertex counts come from the device-side compaction
// totals, so the host never has to read them before the launch.
__global__ void __launch_bounds__(32) chain_kernel(surfd_mccore::Chain* gp, surfd_mccore::Chain* result_host,
                                                   const int64_t* __restrict__ n_cand_dev, const int64_t* __restrict__ n_vtx_dev,
                                                   int64_t cap_cand) {
  __shared__ surfd_mccore::ChainCache cc;
  surfd_mccore::mc_lut_load();
  surfd_mccore::Chain g = *gp;
  const int64_t n = *n_cand_dev;
  g.n_cand = n < cap_cand ? n : cap_cand;
  g.n_cand_total = n;
  g.n_vtx = *n_vtx_dev;
  g.n_v = 0; g.n_f3 = 0; g.n_accept = 0; g.n_seed = 0; g.n_unsure_push = 0; g.n_nontrivial_push = 0;
  if (n == 0) {
    g.status = surfd_mccore::MC_EMPTY;
  } else if (n > cap_cand || g.n_vtx > g.cap_vtx) {
    g.status = surfd_mccore::MC_CAPACITY;   // candidate list truncated / vertex state too small: caller retries with more room
  } else {
    surfd_mccore::replay_r(g, cc, gp);
  }
  __syncwarp();
  // results go straight to mapped pinned host memory: no device->host copy has to be queued behind this long kernel
  if (threadIdx.x == 0) { *gp = g; *result_host = g; }
}

__global__ void __launch_bounds__(128) emit_kernel(const surfd_mccore::Chain* __restrict__ gp) {
  surfd_mccore::mc_lut_load();
  if (gp->status != surfd_mccore::MC_OK) return;
  surfd_mccore::Chain g = *gp;
  for (int64_t a = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; a < g.n_accept; a += (int64_t)gridDim.x * blockDim.x)
    surfd_mccore::emit_cube(g, a);
}

}  // namespace surfd

using namespace surfd;

struct surfd_mc {
  DevBuf bits, list, cand_prefix, vbits, vtx_prefix, recs, vs, done, slot, acc, verts, faces, queues, grid_dev;
  surfd_mccore::Chain* grid_host = nullptr;   // mapped pinned: the kernel writes the results here
  surfd_mccore::Chain* grid_host_dev = nullptr;
  surfd_mccore::Chain* grid_stage = nullptr;  // pinned: launch parameters
  Compactor comp, comp_v;
  int64_t n_v = 0, n_f3 = 0;
  int64_t cap_cand = 0;
  int q_shift = 0;     // extra doublings of the queue capacity after a queue overflow
  int N = 0;
  bool pending = false;
  cudaStream_t pending_stream = nullptr;
};

extern "C" int surfd_mc_create(surfd_mc** out) {
  SURFD_REQUIRE(out != nullptr, "null argument");
  surfd_mc* m = new surfd_mc();
  int st = m->comp.init();
  if (!st) st = m->comp_v.init();
  if (st) { delete m; return st; }
  cudaError_t e = cudaHostAlloc(&m->grid_host, sizeof(surfd_mccore::Chain), cudaHostAllocMapped);
  if (e == cudaSuccess) e = cudaHostGetDevicePointer(&m->grid_host_dev, m->grid_host, 0);
  if (e == cudaSuccess) e = cudaMallocHost(&m->grid_stage, sizeof(surfd_mccore::Chain));
  if (e != cudaSuccess) { m->comp.destroy(); m->comp_v.destroy(); delete m; return set_error(-(int)e, cudaGetErrorString(e), __FILE__, __LINE__); }
  st = m->grid_dev.reserve(sizeof(surfd_mccore::Chain));
  if (st) { surfd_mc_destroy(m); return st; }
  // The chain is a single long-running warp that shares the GPU with the decoder's persistent tcgen05 kernel (one 215 KB
  // CTA per SM).  An SM can only host both if they agree on the L1/shared split, so ask for the same max-shared carve-out;
  // otherwise every SM holding a chain is lost to the GEMM and its 148-CTA grid needs a second wave (measured ~2x).
  cudaFuncSetAttribute(chain_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
  *out = m;
  return 0;
}

extern "C" void surfd_mc_destroy(surfd_mc* m) {
  if (!m) return;
  DevBuf* all[] = {&m->bits, &m->list, &m->cand_prefix, &m->vbits, &m->vtx_prefix, &m->recs, &m->vs, &m->done, &m->slot, &m->acc,
                   &m->verts, &m->faces, &m->queues, &m->grid_dev};
  for (DevBuf* b : all) b->release();
  if (m->grid_host) cudaFreeHost(m->grid_host);
  if (m->grid_stage) cudaFreeHost(m->grid_stage);
  m->comp.destroy();
  m->comp_v.destroy();
  delete m;
}

static int run_classify(surfd_mc* m, const float* udf, int N, cudaStream_t st) {
  const int64_t n3 = (int64_t)N * N * N;
  const int64_t words = cdiv(n3, 32);
  SURFD_TRY(m->bits.reserve((size_t)(words + 1) * sizeof(uint32_t)));
  SURFD_CUDA(cudaMemsetAsync(m->bits.p, 0, (size_t)(words + 1) * sizeof(uint32_t), st));
  const double voxel = 2.0 / (N - 1);              // pyx:1131 (hard-coded [-1,1] range)
  const float avg_t = (float)(1.05 * voxel);       // pyx:1157
  const float max_t = (float)(1.74 * voxel);       // pyx:1158
  const int xsegs = (N + 31) / 32;
  const int64_t warps = (int64_t)xsegs * (N - 1);
  dim3 grid((unsigned)cdiv(warps, 8), (unsigned)cdiv(N - 1, kZChunk));
  classify_kernel<<<grid, 256, 0, st>>>(udf, N, avg_t, max_t, m->bits.as<uint32_t>(), (N % 32) == 0 ? 1 : 0);
  SURFD_CHECK_LAUNCH();
  SURFD_TRY(m->comp.count(m->bits.as<uint32_t>(), words, st));
  return 0;
}

extern "C" int surfd_mc_classify(surfd_mc* m, const float* udf_dev, int N, uint32_t* bits_out, int64_t* n_cand_host,
                                 void* stream) {
  SURFD_REQUIRE(m && udf_dev, "null argument");
  SURFD_REQUIRE(N >= 2 && N <= 1024, "Input array must be at least 2x2x2.");
  cudaStream_t st = (cudaStream_t)stream;
  SURFD_TRY(run_classify(m, udf_dev, N, st));
  if (bits_out) {
    const int64_t words = cdiv((int64_t)N * N * N, 32);
    SURFD_CUDA(cudaMemcpyAsync(bits_out, m->bits.p, (size_t)words * sizeof(uint32_t), cudaMemcpyDeviceToDevice, st));
  }
  if (n_cand_host) SURFD_TRY(m->comp.read_total(n_cand_host, st));
  return 0;
}

// Measurement hook (bench.py): `iters` back-to-back launches of classify_kernel alone over a resident lattice, CUDA events on
// `stream`; the kernel's algorithmic traffic is 4 bytes per lattice point read + 1 bit written.
extern "C" int surfd_mc_time_classify(surfd_mc* m, const float* udf_dev, int N, int iters, float* ms_per_launch, void* stream) {
  SURFD_REQUIRE(m && udf_dev && ms_per_launch && iters >= 1, "null argument");
  SURFD_REQUIRE(N >= 2 && N <= 1024, "Input array must be at least 2x2x2.");
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t words = cdiv((int64_t)N * N * N, 32);
  SURFD_TRY(m->bits.reserve((size_t)(words + 1) * sizeof(uint32_t)));
  SURFD_CUDA(cudaMemsetAsync(m->bits.p, 0, (size_t)(words + 1) * sizeof(uint32_t), st));
  const double voxel = 2.0 / (N - 1);
  const float avg_t = (float)(1.05 * voxel), max_t = (float)(1.74 * voxel);
  const int xsegs = (N + 31) / 32;
  dim3 grid((unsigned)cdiv((int64_t)xsegs * (N - 1), 8), (unsigned)cdiv(N - 1, kZChunk));
  cudaEvent_t e0, e1;
  SURFD_CUDA(cudaEventCreate(&e0));
  SURFD_CUDA(cudaEventCreate(&e1));
  for (int i = 0; i < 2; ++i) { classify_kernel<<<grid, 256, 0, st>>>(udf_dev, N, avg_t, max_t, m->bits.as<uint32_t>(), (N % 32) == 0 ? 1 : 0); SURFD_CHECK_LAUNCH(); }
  SURFD_CUDA(cudaEventRecord(e0, st));
  for (int i = 0; i < iters; ++i) { classify_kernel<<<grid, 256, 0, st>>>(udf_dev, N, avg_t, max_t, m->bits.as<uint32_t>(), (N % 32) == 0 ? 1 : 0); SURFD_CHECK_LAUNCH(); }
  SURFD_CUDA(cudaEventRecord(e1, st));
  SURFD_CUDA(cudaEventSynchronize(e1));
  float ms = 0.f;
  SURFD_CUDA(cudaEventElapsedTime(&ms, e0, e1));
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  *ms_per_launch = ms / iters;
  return 0;
}

// Enqueue classification + compaction + records + ordered chain + emission on `stream` without any host synchronisation:
// the candidate / vertex counts stay on the device and every buffer is sized from a per-handle capacity (grown on retry).
extern "C" int surfd_mc_launch(surfd_mc* m, const float* udf_dev, const float* grad_dev, int N, void* stream) {
  SURFD_REQUIRE(m && udf_dev && grad_dev, "null argument");
  SURFD_REQUIRE(N >= 2 && N <= 1024, "Input array must be at least 2x2x2.");
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t n3 = (int64_t)N * N * N;
  const int64_t words = cdiv(n3, 32);
  m->N = N;
  SURFD_TRY(run_classify(m, udf_dev, N, st));
  // capacity: candidates live in a thin shell around the surface, O(N^2); start at 3*N^2 (a sphere of radius 0.5
  // gives ~1.55*N^2) and let surfd_mc_finish() report SURFD_CAPACITY so the wrapper can grow it.
  int64_t cap = m->cap_cand > 0 ? m->cap_cand : std::max<int64_t>(3ll * N * N, 4096);
  if (cap > n3) cap = n3;
  m->cap_cand = cap;
  // a vertex is shared by up to 8 candidate cubes; a shell 2-3 cubes thick has ~1.3 vertices per cube
  int64_t cap_vtx = 4 * cap;
  if (cap_vtx > n3) cap_vtx = n3;
  SURFD_TRY(m->list.reserve((size_t)cap * sizeof(int32_t)));
  SURFD_TRY(m->comp.scatter(m->bits.as<uint32_t>(), words, m->list.as<int32_t>(), cap, st));
  SURFD_TRY(m->cand_prefix.reserve((size_t)(words + 1) * sizeof(int32_t)));
  SURFD_TRY(m->comp.word_prefix(m->bits.as<uint32_t>(), words, m->cand_prefix.as<int32_t>(), st));
  // vertex set + its rank structure
  SURFD_TRY(m->vbits.reserve((size_t)(words + 1) * sizeof(uint32_t)));
  SURFD_TRY(m->vtx_prefix.reserve((size_t)(words + 1) * sizeof(int32_t)));
  SURFD_CUDA(cudaMemsetAsync(m->vbits.p, 0, (size_t)(words + 1) * sizeof(uint32_t), st));
  mark_vertices_kernel<<<592, 256, 0, st>>>(m->list.as<int32_t>(), m->comp.d_total, cap, N, m->vbits.as<uint32_t>());
  SURFD_CHECK_LAUNCH();
  SURFD_TRY(m->comp_v.count(m->vbits.as<uint32_t>(), words, st));
  SURFD_TRY(m->comp_v.word_prefix(m->vbits.as<uint32_t>(), words, m->vtx_prefix.as<int32_t>(), st));
  // O(surface) state
  SURFD_TRY(m->recs.reserve((size_t)cap * sizeof(surfd_mccore::Rec)));
  SURFD_TRY(m->vs.reserve((size_t)cap_vtx));
  SURFD_TRY(m->done.reserve((size_t)cap));
  SURFD_TRY(m->slot.reserve((size_t)cap_vtx * 4 * sizeof(int32_t)));
  SURFD_TRY(m->acc.reserve((size_t)cap * sizeof(surfd_mccore::Accept)));
  SURFD_CUDA(cudaMemsetAsync(m->vs.p, 0, (size_t)cap_vtx, st));
  SURFD_CUDA(cudaMemsetAsync(m->done.p, 0, (size_t)cap, st));
  SURFD_CUDA(cudaMemsetAsync(m->slot.p, 0xFF, (size_t)cap_vtx * 4 * sizeof(int32_t), st));
  // observed V ~ 0.8 n_cand, 3F ~ 4.7 n_cand on closed surfaces; 3x / 12x leaves room for noisy fields
  const int64_t cap_v = 3 * cap + 64;
  const int64_t cap_f3 = 12 * cap + 64;
  SURFD_TRY(m->verts.reserve((size_t)cap_v * 3 * sizeof(float)));
  SURFD_TRY(m->faces.reserve((size_t)cap_f3 * sizeof(int32_t)));
  // every accepted cube pushes <= 6 entries and is accepted once: 6*cap bounds the BFS queue's lifetime traffic, hence its
  // occupancy; the two priority queues hold cubes waiting for a second look (grown on MC_QUEUE_OVERFLOW)
  uint32_t qcap = 1024, qcap2 = 1024;
  while ((int64_t)qcap < 6 * cap + 1024) qcap <<= 1;
  while ((int64_t)qcap2 < cap + 1024) qcap2 <<= 1;
  qcap <<= m->q_shift; qcap2 <<= m->q_shift;
  SURFD_TRY(m->queues.reserve(((size_t)qcap + 2 * (size_t)qcap2) * sizeof(int32_t)));

  surfd_mccore::Chain& g = *m->grid_stage;
  memset(&g, 0, sizeof(g));
  g.N = N; g.im = udf_dev; g.grads = grad_dev;
  g.cand_bits = m->bits.as<uint32_t>(); g.cand_prefix = m->cand_prefix.as<int32_t>();
  g.vtx_bits = m->vbits.as<uint32_t>(); g.vtx_prefix = m->vtx_prefix.as<int32_t>();
  g.cand_list = m->list.as<int32_t>(); g.n_cand = 0; g.cap_vtx = cap_vtx;
  g.recs = m->recs.as<surfd_mccore::Rec>(); g.vs = m->vs.as<uint8_t>(); g.done = m->done.as<uint8_t>();
  g.slot = m->slot.as<int32_t>(); g.acc = m->acc.as<surfd_mccore::Accept>();
  g.verts = m->verts.as<float>(); g.cap_v = cap_v; g.faces = m->faces.as<int32_t>(); g.cap_f3 = cap_f3;
  g.q.buf = m->queues.as<int32_t>(); g.q_unsure.buf = g.q.buf + qcap; g.q_nontrivial.buf = g.q_unsure.buf + qcap2;
  g.q.mask = qcap - 1; g.q_unsure.mask = g.q_nontrivial.mask = qcap2 - 1;
  SURFD_CUDA(cudaMemcpyAsync(m->grid_dev.p, m->grid_stage, sizeof(g), cudaMemcpyHostToDevice, st));
  surfd_mccore::Chain* gd = m->grid_dev.as<surfd_mccore::Chain>();
  build_records_kernel<<<592, 128, 0, st>>>(gd, m->comp.d_total, m->comp_v.d_total, cap);
  SURFD_CHECK_LAUNCH();
  chain_kernel<<<1, 32, 0, st>>>(gd, m->grid_host_dev, m->comp.d_total, m->comp_v.d_total, cap);
  SURFD_CHECK_LAUNCH();
  emit_kernel<<<592, 128, 0, st>>>(gd);
  SURFD_CHECK_LAUNCH();
  m->pending_stream = st;
  m->pending = true;
  return 0;
}

// Wait for the launch on its stream and report counts / status.
extern "C" int surfd_mc_finish(surfd_mc* m, int64_t* n_v, int64_t* n_f, int64_t* stats) {
  SURFD_REQUIRE(m && n_v && n_f, "null argument");
  SURFD_REQUIRE(m->pending, "surfd_mc_finish without surfd_mc_launch");
  SURFD_CUDA(cudaStreamSynchronize(m->pending_stream));
  m->pending = false;
  surfd_mccore::Chain g;
  memcpy(&g, m->grid_host, sizeof(g));   // written by the kernel into mapped pinned memory; the stream sync above orders it
  m->n_v = g.n_v; m->n_f3 = g.n_f3;
  *n_v = g.n_v; *n_f = g.n_f3 / 3;
  if (stats) {
    for (int i = 0; i < 8; ++i) stats[i] = 0;
    stats[0] = g.n_cand_total; stats[1] = g.n_seed; stats[2] = g.n_accept; stats[3] = g.n_unsure_push; stats[4] = g.n_nontrivial_push;
  }
  if (g.status == surfd_mccore::MC_EMPTY) { m->n_v = m->n_f3 = 0; return SURFD_EMPTY_SURFACE; }
  if (g.status == surfd_mccore::MC_CAPACITY) {
    // grow for the retry: enough for the whole candidate set, or double when vertices/faces overflowed
    int64_t want = g.n_cand_total > m->cap_cand ? g.n_cand_total + g.n_cand_total / 8 : 2 * m->cap_cand;
    if (g.n_vtx > g.cap_vtx) want = std::max<int64_t>(g.n_cand_total + g.n_cand_total / 8, g.n_vtx / 4 + g.n_vtx / 32);
    m->cap_cand = want;
    m->n_v = m->n_f3 = 0;
    return set_error(SURFD_CAPACITY, "marching cubes capacity exceeded; call again (buffers were grown)", __FILE__, __LINE__);
  }
  if (g.status == surfd_mccore::MC_QUEUE_OVERFLOW) {
    if (m->q_shift < 4) {   // retry with larger queues (same path as a capacity miss)
      ++m->q_shift;
      m->n_v = m->n_f3 = 0;
      return set_error(SURFD_CAPACITY, "marching cubes queue capacity exceeded; call again (queues were grown)", __FILE__, __LINE__);
    }
    return set_error(SURFD_QUEUE_OVERFLOW, "marching cubes BFS queue overflow", __FILE__, __LINE__);
  }
  return 0;
}

extern "C" int surfd_mc_profile(surfd_mc* m, int64_t* prof) {
  SURFD_REQUIRE(m && prof, "null argument");
  SURFD_REQUIRE(!m->pending, "surfd_mc_profile while a launch is pending");
  for (int i = 0; i < 8; ++i) prof[i] = m->grid_host->prof[i];
  return 0;
}

extern "C" int surfd_mc_udf(surfd_mc* m, const float* udf_dev, const float* grad_dev, int N, int64_t* n_v, int64_t* n_f,
                            int64_t* stats, void* stream) {
  for (int attempt = 0; attempt < 6; ++attempt) {
    SURFD_TRY(surfd_mc_launch(m, udf_dev, grad_dev, N, stream));
    const int rc = surfd_mc_finish(m, n_v, n_f, stats);
    if (rc != SURFD_CAPACITY) return rc;
  }
  return set_error(SURFD_CAPACITY, "marching cubes capacity exceeded after retries", __FILE__, __LINE__);
}

extern "C" int surfd_mc_fetch(surfd_mc* m, float* verts_dev, int32_t* faces_dev, void* stream) {
  SURFD_REQUIRE(m != nullptr, "null argument");
  cudaStream_t st = (cudaStream_t)stream;
  if (m->n_v > 0) {
    SURFD_REQUIRE(verts_dev != nullptr, "null vertex buffer");
    SURFD_CUDA(cudaMemcpyAsync(verts_dev, m->verts.p, (size_t)m->n_v * 3 * sizeof(float), cudaMemcpyDeviceToDevice, st));
  }
  if (m->n_f3 > 0) {
    SURFD_REQUIRE(faces_dev != nullptr, "null face buffer");
    SURFD_CUDA(cudaMemcpyAsync(faces_dev, m->faces.p, (size_t)m->n_f3 * sizeof(int32_t), cudaMemcpyDeviceToDevice, st));
  }
  return 0;
}
