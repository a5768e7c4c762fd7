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
#include "mc_core.h"

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

__global__ void replay_kernel(surfd_mccore::Grid* g) {
  if (threadIdx.x == 0 && blockIdx.x == 0) surfd_mccore::replay(*g);
}

}  // namespace surfd

using namespace surfd;

struct surfd_mc {
  DevBuf bits, list, sgn, flg, face_layer, verts, faces, queues, grid_dev;
  surfd_mccore::Grid* grid_host = nullptr;  // pinned
  Compactor comp;
  int64_t n_v = 0, n_f3 = 0;
};

extern "C" int surfd_mc_create(surfd_mc** out) {
  SURFD_REQUIRE(out != nullptr, "null argument");
  surfd_mc* m = new surfd_mc();
  int st = m->comp.init();
  if (st) { delete m; return st; }
  cudaError_t e = cudaMallocHost(&m->grid_host, sizeof(surfd_mccore::Grid));
  if (e != cudaSuccess) { m->comp.destroy(); delete m; return set_error(-(int)e, cudaGetErrorString(e), __FILE__, __LINE__); }
  st = m->grid_dev.reserve(sizeof(surfd_mccore::Grid));
  if (st) { surfd_mc_destroy(m); return st; }
  *out = m;
  return 0;
}

extern "C" void surfd_mc_destroy(surfd_mc* m) {
  if (!m) return;
  m->bits.release(); m->list.release(); m->sgn.release(); m->flg.release(); m->face_layer.release();
  m->verts.release(); m->faces.release(); m->queues.release(); m->grid_dev.release();
  if (m->grid_host) cudaFreeHost(m->grid_host);
  m->comp.destroy();
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

extern "C" int surfd_mc_udf(surfd_mc* m, const float* udf_dev, const float* grad_dev, int N, int64_t* n_v, int64_t* n_f,
                            int64_t* stats, void* stream) {
  SURFD_REQUIRE(m && udf_dev && grad_dev && n_v && n_f, "null argument");
  SURFD_REQUIRE(N >= 2 && N <= 1024, "Input array must be at least 2x2x2.");
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t n3 = (int64_t)N * N * N;
  const int64_t words = cdiv(n3, 32);
  SURFD_TRY(run_classify(m, udf_dev, N, st));
  int64_t n_cand = 0;
  SURFD_TRY(m->comp.read_total(&n_cand, st));
  *n_v = 0; *n_f = 0; m->n_v = 0; m->n_f3 = 0;
  if (stats) { for (int i = 0; i < 8; ++i) stats[i] = 0; stats[0] = n_cand; }
  if (n_cand == 0) return SURFD_EMPTY_SURFACE;
  SURFD_TRY(m->list.reserve((size_t)n_cand * sizeof(int32_t)));
  SURFD_TRY(m->comp.scatter(m->bits.as<uint32_t>(), words, m->list.as<int32_t>(), st));

  SURFD_TRY(m->sgn.reserve((size_t)n3));
  SURFD_TRY(m->flg.reserve((size_t)n3));
  SURFD_TRY(m->face_layer.reserve((size_t)n3 * 4 * sizeof(int32_t)));
  SURFD_CUDA(cudaMemsetAsync(m->sgn.p, 0, (size_t)n3, st));
  SURFD_CUDA(cudaMemsetAsync(m->flg.p, 0, (size_t)n3, st));
  SURFD_CUDA(cudaMemsetAsync(m->face_layer.p, 0xFF, (size_t)n3 * 4 * sizeof(int32_t), st));
  // every emitted vertex owns one of the 4 slots of some cell adjacent to a candidate cube; 13 corners per
  // accepted cube is the hard bound for faces (12 triangles x 3 in the largest Lewiner tiling).
  const int64_t cap_v = 13 * n_cand + 64;
  const int64_t cap_f3 = 36 * n_cand + 64;
  SURFD_TRY(m->verts.reserve((size_t)cap_v * 3 * sizeof(float)));
  SURFD_TRY(m->faces.reserve((size_t)cap_f3 * sizeof(int32_t)));
  uint32_t qcap = 1024;
  while ((int64_t)qcap < 16 * n_cand + 1024) qcap <<= 1;
  SURFD_TRY(m->queues.reserve((size_t)qcap * 3 * sizeof(int32_t)));

  surfd_mccore::Grid& g = *m->grid_host;
  memset(&g, 0, sizeof(g));
  g.N = N; g.im = udf_dev; g.grads = grad_dev;
  g.cand_bits = m->bits.as<uint32_t>(); g.cand_list = m->list.as<int32_t>(); g.n_cand = n_cand;
  g.sgn = m->sgn.as<int8_t>(); g.flg = m->flg.as<uint8_t>(); g.face_layer = m->face_layer.as<int32_t>();
  g.verts = m->verts.as<float>(); g.cap_v = cap_v; g.faces = m->faces.as<int32_t>(); g.cap_f3 = cap_f3;
  g.q.buf = m->queues.as<int32_t>(); g.q_unsure.buf = g.q.buf + qcap; g.q_nontrivial.buf = g.q.buf + 2 * (size_t)qcap;
  g.q.mask = g.q_unsure.mask = g.q_nontrivial.mask = qcap - 1;
  SURFD_CUDA(cudaMemcpyAsync(m->grid_dev.p, &g, sizeof(g), cudaMemcpyHostToDevice, st));
  replay_kernel<<<1, 32, 0, st>>>(m->grid_dev.as<surfd_mccore::Grid>());
  SURFD_CHECK_LAUNCH();
  SURFD_CUDA(cudaMemcpyAsync(&g, m->grid_dev.p, sizeof(g), cudaMemcpyDeviceToHost, st));
  SURFD_CUDA(cudaStreamSynchronize(st));
  m->n_v = g.n_v; m->n_f3 = g.n_f3;
  *n_v = g.n_v; *n_f = g.n_f3 / 3;
  if (stats) { stats[1] = g.n_seed; stats[2] = g.n_accept; stats[3] = g.n_unsure_push; stats[4] = g.n_nontrivial_push; }
  if (g.status == surfd_mccore::MC_EMPTY) return SURFD_EMPTY_SURFACE;
  if (g.status == surfd_mccore::MC_CAPACITY) return set_error(SURFD_CAPACITY, "marching cubes output bound exceeded", __FILE__, __LINE__);
  if (g.status == surfd_mccore::MC_QUEUE_OVERFLOW) return set_error(SURFD_QUEUE_OVERFLOW, "marching cubes BFS queue overflow", __FILE__, __LINE__);
  return 0;
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
