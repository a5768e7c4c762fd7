// decoder_tc.cu -- the decoder's 512x512 layer GEMM on the 5th-generation tensor cores (sm_100a):
// C[M][512] = A[M][512] * W[512][512]^T with the same fused epilogue as the FFMA kernel (bias, CBN/ReLU mask,
// residual, next layer's CBN+ReLU activation), operands in TF32 (kind::tf32, fp32 accumulate in TMEM).
//
// Structure (one persistent CTA per SM, 10 warps, no cluster):
//   warp 0  TMA producer : cp.async.bulk.tensor.2d of the A tile (128 x 32 fp32, 128B-swizzled) and the W tile
//                          (256 x 32) into a 4-stage shared-memory ring, mbarrier complete_tx
//   warp 1  MMA issuer   : one elected lane issues tcgen05.mma.cta_group::1.kind::tf32 (M=128, N=256, K=8) x4 per
//                          stage, tcgen05.commit frees the stage; the 128x256 fp32 accumulator lives in TMEM and
//                          is double-buffered (2 x 256 columns) so the epilogue of tile i overlaps the MMAs of tile i+1
//   warps 2-9 epilogue   : tcgen05.ld 32x32b (one accumulator row per thread), fused epilogue, vectorised row stores;
//                          two warps per TMEM lane quadrant, one half of the tile's columns each
// Both operands are K-major (activations [points][K], weights [out][K], K contiguous), so A and B tiles use the same
// canonical SWIZZLE_128B K-major layout that TMA writes and the UMMA shared-memory descriptor reads.
#include <cuda.h>

#include "common.cuh"
#include "decoder_common.cuh"

namespace surfd {
namespace tc {

constexpr int BM = 128, BN = 256, BK = 32, STAGES = 4;
constexpr int A_BYTES = BM * BK * 4;               // 16 KB
constexpr int B_BYTES = BN * BK * 4;               // 32 KB
constexpr int STAGE_BYTES = A_BYTES + B_BYTES;     // 48 KB
constexpr int EPI_WARPS = 8;                       // two per TMEM lane quadrant: each takes one half of the tile's 256 columns
constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align slack*/ + 256 /*barriers*/ + EPI_WARPS * 32 * 32 * 4 /*epilogue stages*/;
constexpr int K_TOTAL = 512, KBLOCKS = K_TOTAL / BK;
constexpr int THREADS = 64 + 32 * EPI_WARPS;
constexpr uint32_t SPIN_LIMIT = 1u << 27;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// bounded spin: a protocol bug traps (error reported by the runtime) instead of hanging the GPU
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity, int* err, int code) {
  uint32_t done = 0, spins = 0;
  while (!done) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    if (!done && ++spins > SPIN_LIMIT) {
      if (err) *err = code;
      __threadfence_system();
      asm volatile("trap;");
    }
  }
}

__device__ __forceinline__ void tma_load_2d(const CUtensorMap* map, uint64_t* bar, void* dst, int x, int y) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(x), "r"(y)
      : "memory");
}

// UMMA shared-memory descriptor, K-major, SWIZZLE_128B: rows of 128 B, 8-row atoms of 1024 B (SBO), version 1.
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);        // start address  [0,14)
  d |= (uint64_t)1 << 16;                          // leading byte offset (ignored for swizzled K-major) [16,30)
  d |= (uint64_t)(1024 >> 4) << 32;                // stride byte offset = 1024 B [32,46)
  d |= (uint64_t)1 << 46;                          // descriptor version (Blackwell) [46,48)
  d |= (uint64_t)2 << 61;                          // layout type SWIZZLE_128B [61,64)
  return d;
}

// kind::tf32 instruction descriptor: D=F32, A=B=TF32, both K-major, N=256, M=128
__device__ __forceinline__ uint32_t umma_idesc() {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
}

__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
        "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
        "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// One epilogue warp's share of a 128 x 256 accumulator tile: TMEM lane quadrant `quad` (32 rows), columns [half * 128, + 128).
// ncu (r2, source view): the epilogue, not HBM or the tensor pipe, bounds a tile -- ~5,000 dependent instructions per warp and
// tile with ONE warp per scheduler (0.2 IPC: 13 us against 4.2 us of MMAs).  Eight epilogue warps (two per scheduler) halve
// each warp's share and let the schedulers overlap the two.
// Each thread owns one accumulator row in TMEM, but row-per-lane global accesses are fully divergent, so every 32 x 32 block is
// transposed through a shared-memory stage (16-byte pieces XOR-swizzled by the row: conflict-free for the row-wise writes
// and the 8-lanes-per-row reads) and the global side runs with 8 lanes per row: 128 contiguous bytes per row, 4 rows per access.
__device__ __forceinline__ void epilogue_tile(const Epilogue& e, int M, int m0, int n0, int quad, uint32_t tmem_tile, float* stage, int lane) {
  const int srow = lane >> 3, scol4 = lane & 7, scol = scol4 * 4;
#pragma unroll 1
  for (int c = 0; c < BN / 32 / 2; ++c) {
    // Residual / mask operands of this 32x32 block first: all 8 (+8) row segments are requested before anything is stored
    // (R and C are the same buffer for the in-place residual update: loads placed after stores would be serialised).
    float4 rq[8], mk[8];
    // per-column constants of this block: requested first (ncu r2: the bias add right behind its __ldg was the top stall of the
    // kernel, 14 % of all warp samples -- one exposed L1/L2 round trip per block)
    const int n = n0 + c * 32 + scol;
    float4 bias = make_float4(0.f, 0.f, 0.f, 0.f), ms = bias, s2 = bias, t2 = bias;
    if (e.bias) bias = __ldg(reinterpret_cast<const float4*>(e.bias + n));
    if (e.mask) ms = __ldg(reinterpret_cast<const float4*>(e.mscale + n));
    if (e.act) { s2 = __ldg(reinterpret_cast<const float4*>(e.s2 + n)); t2 = __ldg(reinterpret_cast<const float4*>(e.t2 + n)); }
    {
      const int nn = n0 + c * 32 + scol;
#pragma unroll
      for (int it = 0; it < 8; ++it) {
        const int m = m0 + quad * 32 + it * 4 + srow;
        const size_t off = (size_t)(m < M ? m : M - 1) * e.ld + nn;
        if (e.R) rq[it] = *reinterpret_cast<const float4*>(e.R + off);
        if (e.mask) mk[it] = *reinterpret_cast<const float4*>(e.mask + off);
      }
    }
    uint32_t r[32];
    tmem_ld32(tmem_tile + ((uint32_t)(quad * 32) << 16) + (uint32_t)(c * 32), r);
    __syncwarp();
#pragma unroll
    for (int q = 0; q < 8; ++q)
      *reinterpret_cast<float4*>(stage + lane * 32 + ((q ^ (lane & 7)) << 2)) =
          make_float4(__uint_as_float(r[4 * q]), __uint_as_float(r[4 * q + 1]), __uint_as_float(r[4 * q + 2]), __uint_as_float(r[4 * q + 3]));
    __syncwarp();
#pragma unroll
    for (int it = 0; it < 8; ++it) {
      const int rr = it * 4 + srow;
      const int m = m0 + quad * 32 + rr;
      if (m >= M) continue;
      float4 v = *reinterpret_cast<const float4*>(stage + rr * 32 + ((scol4 ^ (rr & 7)) << 2));
      const size_t off = (size_t)m * e.ld + n;
      v.x += bias.x; v.y += bias.y; v.z += bias.z; v.w += bias.w;
      if (e.mask) {
        const float4 k4 = mk[it];
        v.x = k4.x > 0.f ? v.x * ms.x : 0.f; v.y = k4.y > 0.f ? v.y * ms.y : 0.f;
        v.z = k4.z > 0.f ? v.z * ms.z : 0.f; v.w = k4.w > 0.f ? v.w * ms.w : 0.f;
      }
      if (e.R) { const float4 q4 = rq[it]; v.x += q4.x; v.y += q4.y; v.z += q4.z; v.w += q4.w; }
      if (e.C) {
        float4 o = v;
        if (e.round_c) { o.x = round_to_tf32(o.x); o.y = round_to_tf32(o.y); o.z = round_to_tf32(o.z); o.w = round_to_tf32(o.w); }
        *reinterpret_cast<float4*>(e.C + off) = o;
      }
      if (e.act) {
        float4 a;
        a.x = fmaxf(fmaf(s2.x, v.x, t2.x), 0.f); a.y = fmaxf(fmaf(s2.y, v.y, t2.y), 0.f);
        a.z = fmaxf(fmaf(s2.z, v.z, t2.z), 0.f); a.w = fmaxf(fmaf(s2.w, v.w, t2.w), 0.f);
        if (e.round_act) { a.x = round_to_tf32(a.x); a.y = round_to_tf32(a.y); a.z = round_to_tf32(a.z); a.w = round_to_tf32(a.w); }
        *reinterpret_cast<float4*>(e.act + off) = a;
      }
    }
  }
}

__global__ void __launch_bounds__(THREADS, 1)
tc_gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, int M, Epilogue e, int* err) {
  extern __shared__ uint8_t smem_raw[];
  // 1024-byte alignment by pointer arithmetic on the __shared__ array (not through an integer cast): the compiler keeps the
  // shared address space and the epilogue's staging accesses compile to LDS / STS instead of generic LD / ST (ncu r2: the
  // generic loads were the top stall of the epilogue warps)
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);
  uint64_t* full = bars;                 // [STAGES]
  uint64_t* empty = bars + STAGES;       // [STAGES]
  uint64_t* tfull = bars + 2 * STAGES;   // [2]
  uint64_t* tempty = tfull + 2;          // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m_tiles = (M + BM - 1) / BM;
  const int n_tiles = 2 * m_tiles;

  if (threadIdx.x == 0) {
    for (int i = 0; i < STAGES; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&tfull[i], 1); mbar_init(&tempty[i], EPI_WARPS); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmA)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmB)) : "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(tmem_slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===== TMA producer =====
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int m0 = (tile >> 1) * BM, n0 = (tile & 1) * BN;
        for (int kb = 0; kb < KBLOCKS; ++kb) {
          mbar_wait(&empty[stage], phase ^ 1, err, 1);
          uint8_t* sa = smem + stage * STAGE_BYTES;
          mbar_expect_tx(&full[stage], STAGE_BYTES);
          tma_load_2d(&tmA, &full[stage], sa, kb * BK, m0);
          tma_load_2d(&tmB, &full[stage], sa + A_BYTES, kb * BK, n0);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    if (lane == 0) {
      const uint32_t idesc = umma_idesc();
      int stage = 0; uint32_t phase = 0;
      int acc = 0; uint32_t acc_phase = 0;
      for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        mbar_wait(&tempty[acc], acc_phase ^ 1, err, 2);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * BN);
        for (int kb = 0; kb < KBLOCKS; ++kb) {
          mbar_wait(&full[stage], phase, err, 3);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + stage * STAGE_BYTES);
          const uint64_t adesc = umma_desc(sa), bdesc = umma_desc(sa + A_BYTES);
#pragma unroll
          for (int k = 0; k < BK / 8; ++k) {
            // advance 8 tf32 = 32 bytes along K inside the 128-byte swizzle atom: +2 in the (addr >> 4) field
            umma_tf32(d_tmem, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, (kb | k) != 0 ? 1u : 0u);
          }
          umma_commit(&empty[stage]);                      // stage reusable once these MMAs have read it
          if (kb == KBLOCKS - 1) umma_commit(&tfull[acc]);  // accumulator complete
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else {
    // ===== epilogue warps 2..9: TMEM lane quadrant = warp % 4, column half = (warp - 2) / 4 =====
    const int quad = warp & 3, half = (warp - 2) >> 2;
    float* stage = reinterpret_cast<float*>(smem + STAGES * STAGE_BYTES + 256) + (warp - 2) * (32 * 32);
    int acc = 0; uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
      const int m0 = (tile >> 1) * BM, n0 = (tile & 1) * BN + half * (BN / 2);
      mbar_wait(&tfull[acc], acc_phase, err, 4);
      tc_fence_after();
      epilogue_tile(e, M, m0, n0, quad, tmem_base + (uint32_t)(acc * BN + half * (BN / 2)), stage, lane);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty[acc]);
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base) : "memory");
  }
}

// ---- the layer chain in one launch ---------------------------------------------------------------------------------
// Every layer of the decoder is ROW-LOCAL: row m of a layer's output depends on row m of its input only.  tc_chain_kernel
// therefore gives each CTA whole 128-row panels and runs ALL the layers of a pass on a panel before moving on: what a layer's
// epilogue writes is read back by the same CTA's TMA loads a few tiles later.  No CTA depends on another one -- no grid
// barrier, no cooperative launch, no launch gaps or pipeline refills between layers.  Two panels P, Q are interleaved per
// trip -- P:l:h0, P:l:h1, Q:l:h0, Q:l:h1 for layer l = 0 .. n-1 (h = column half) -- so that the MMAs of one panel cover the
// epilogue drain of the other; the producer waits on `ready[slot]` (all 16 epilogue-warp arrivals of the panel's two
// layer-(l-1) tiles, each after a gpu-scope fence and a generic->async proxy fence) before it requests the panel's layer-l
// rows.  Same roles, stage ring and TMEM double buffer as tc_gemm_kernel.
// Measured (profiles/r2_probe_panels.log): 4-5 % faster than the same layers behind grid barriers.  The HBM traffic is
// unchanged (ncu: 9.2 GB per 161,280-point forward pass = 5.7 KB per point and layer, L2 hit rate 61 %): with 140 CTAs x 2
// panels x 768 KB in flight the panels' activations are evicted before they are read back (L2: 126 MB), so the chain stays
// bound by that round trip (68 % of the measured copy bandwidth).
struct alignas(64) ChainLayer {
  CUtensorMap tmA;   // this layer's input activations [M][512]
  CUtensorMap tmB;   // its weights [512][512]
  Epilogue e;
};
constexpr int CHAIN_MAX_LAYERS = 10;
struct ChainProg {
  ChainLayer layer[CHAIN_MAX_LAYERS];
};
static_assert(sizeof(ChainProg) <= 3968, "the program travels in the kernel parameters");

__global__ void __launch_bounds__(THREADS, 1)
tc_chain_kernel(const __grid_constant__ ChainProg prog, int n_layers, int M, int* err) {
  extern __shared__ uint8_t smem_raw[];
  // 1024-byte alignment by pointer arithmetic on the __shared__ array (not through an integer cast): the compiler keeps the
  // shared address space and the epilogue's staging accesses compile to LDS / STS instead of generic LD / ST
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);
  uint64_t* full = bars;                 // [STAGES]
  uint64_t* empty = bars + STAGES;       // [STAGES]
  uint64_t* tfull = bars + 2 * STAGES;   // [2]
  uint64_t* tempty = tfull + 2;          // [2]
  uint64_t* ready = tempty + 2;          // [2]  panel slot P / Q: the previous layer's rows are complete
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(ready + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m_tiles = (M + BM - 1) / BM;                                               // panels
  const int npan = (int)blockIdx.x < m_tiles ? (m_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;

  if (threadIdx.x == 0) {
    for (int i = 0; i < STAGES; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&tfull[i], 1); mbar_init(&tempty[i], EPI_WARPS); mbar_init(&ready[i], 2 * EPI_WARPS); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(tmem_slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // pipeline positions: each role advances its own copies, all roles walk the same sequence of tiles and k-blocks
  int stage = 0; uint32_t phase = 0;
  int acc = 0; uint32_t acc_phase = 0;
  if (warp == 0) {
    // ===== TMA producer =====
    if (lane == 0) {
      uint32_t rph[2] = {0u, 0u};
      for (int g = 0; g < npan; g += 2) {
        const int np = npan - g < 2 ? npan - g : 2;
        for (int l = 0; l < n_layers; ++l) {
          const ChainLayer& L = prog.layer[l];
          for (int sl = 0; sl < np; ++sl) {
            const int m0 = ((int)blockIdx.x + (g + sl) * (int)gridDim.x) * BM;
            if (l > 0) {
              mbar_wait(&ready[sl], rph[sl] & 1u, err, 6);
              ++rph[sl];
              asm volatile("fence.proxy.async;" ::: "memory");   // those rows were written with generic stores
            }
            for (int h = 0; h < 2; ++h) {
              for (int kb = 0; kb < KBLOCKS; ++kb) {
                mbar_wait(&empty[stage], phase ^ 1, err, 1);
                uint8_t* sa = smem + stage * STAGE_BYTES;
                mbar_expect_tx(&full[stage], STAGE_BYTES);
                tma_load_2d(&L.tmA, &full[stage], sa, kb * BK, m0);
                tma_load_2d(&L.tmB, &full[stage], sa + A_BYTES, kb * BK, h * BN);
                if (++stage == STAGES) { stage = 0; phase ^= 1; }
              }
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    if (lane == 0) {
      const uint32_t idesc = umma_idesc();
      for (int g = 0; g < npan; g += 2) {
        const int np = npan - g < 2 ? npan - g : 2;
        for (int t = 0; t < n_layers * np * 2; ++t) {
          mbar_wait(&tempty[acc], acc_phase ^ 1, err, 2);
          tc_fence_after();
          const uint32_t d_tmem = tmem_base + (uint32_t)(acc * BN);
          for (int kb = 0; kb < KBLOCKS; ++kb) {
            mbar_wait(&full[stage], phase, err, 3);
            tc_fence_after();
            const uint32_t sa = smem_u32(smem + stage * STAGE_BYTES);
            const uint64_t adesc = umma_desc(sa), bdesc = umma_desc(sa + A_BYTES);
#pragma unroll
            for (int k = 0; k < BK / 8; ++k)
              umma_tf32(d_tmem, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, (kb | k) != 0 ? 1u : 0u);
            umma_commit(&empty[stage]);
            if (kb == KBLOCKS - 1) umma_commit(&tfull[acc]);
            if (++stage == STAGES) { stage = 0; phase ^= 1; }
          }
          if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
      }
    }
  } else {
    // ===== epilogue warps 2..9 (see tc_gemm_kernel) =====
    const int quad = warp & 3, half = (warp - 2) >> 2;
    float* stg = reinterpret_cast<float*>(smem + STAGES * STAGE_BYTES + 256) + (warp - 2) * (32 * 32);
    for (int g = 0; g < npan; g += 2) {
      const int np = npan - g < 2 ? npan - g : 2;
      for (int l = 0; l < n_layers; ++l) {
        const ChainLayer& L = prog.layer[l];
        for (int sl = 0; sl < np; ++sl) {
          const int m0 = ((int)blockIdx.x + (g + sl) * (int)gridDim.x) * BM;
          for (int h = 0; h < 2; ++h) {
            mbar_wait(&tfull[acc], acc_phase, err, 4);
            tc_fence_after();
            epilogue_tile(L.e, M, m0, h * BN + half * (BN / 2), quad, tmem_base + (uint32_t)(acc * BN + half * (BN / 2)), stg, lane);
            tc_fence_before();
            if (l + 1 < n_layers) {
              // this warp's part of the panel's layer-l rows goes to the same CTA's TMA loads of layer l + 1 (async proxy, via L2)
              __threadfence();
              asm volatile("fence.proxy.async;" ::: "memory");
            }
            __syncwarp();
            if (lane == 0) {
              mbar_arrive(&tempty[acc]);
              if (l + 1 < n_layers) mbar_arrive(&ready[sl]);
            }
            if (++acc == 2) { acc = 0; acc_phase ^= 1; }
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base) : "memory");
  }
}

// ---- host side -------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

// 2-D fp32 tensor [rows][512] row-major, box = [box_rows][32 floats], 128-byte swizzle, zero fill out of bounds
static int make_map(CUtensorMap* map, const float* base, int64_t rows, int box_rows) {
  EncodeTiledFn enc = get_encode();
  if (!enc) return set_error(SURFD_BAD_ARGUMENT, "cuTensorMapEncodeTiled not available from the driver", __FILE__, __LINE__);
  cuuint64_t dims[2] = {(cuuint64_t)K_TOTAL, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)K_TOTAL * sizeof(float)};
  cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return set_error(SURFD_BAD_ARGUMENT, "cuTensorMapEncodeTiled failed", __FILE__, __LINE__);
  return 0;
}

}  // namespace tc

// One launch for `n` consecutive 512x512 layers over the same M rows (A[i] -> epilogue e[i]).
int launch_chain_tc(const float* const* A, const float* const* W, const Epilogue* e, int n, int M, unsigned* gsync, int* err_flag,
                    int num_sms, cudaStream_t st) {
  SURFD_REQUIRE(n >= 1 && n <= tc::CHAIN_MAX_LAYERS, "bad layer count");
  static bool attr_set = false;
  if (!attr_set) {
    SURFD_CUDA(cudaFuncSetAttribute(tc::tc_chain_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::SMEM_BYTES));
    attr_set = true;
  }
  static tc::ChainProg prog;   // (host staging for the by-value kernel parameter; the launch copies it)
  for (int i = 0; i < n; ++i) {
    SURFD_TRY(tc::make_map(&prog.layer[i].tmA, A[i], M, tc::BM));
    SURFD_TRY(tc::make_map(&prog.layer[i].tmB, W[i], 512, tc::BN));
    prog.layer[i].e = e[i];
  }
  (void)gsync;   // (no grid barrier any more: every dependency of the chain is CTA-local)
  const int panels = (int)cdiv(M, tc::BM);
  const int grid = panels < num_sms ? panels : num_sms;
  tc::tc_chain_kernel<<<grid, tc::THREADS, tc::SMEM_BYTES, st>>>(prog, n, M, err_flag);
  SURFD_CHECK_LAUNCH();
  return 0;
}

int launch_gemm_tc(const float* A, const float* W, int M, const Epilogue& e, int* err_flag, int num_sms, cudaStream_t st) {
  static bool attr_set = false;
  if (!attr_set) {
    SURFD_CUDA(cudaFuncSetAttribute(tc::tc_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::SMEM_BYTES));
    attr_set = true;
  }
  CUtensorMap ma, mb;
  SURFD_TRY(tc::make_map(&ma, A, M, tc::BM));
  SURFD_TRY(tc::make_map(&mb, W, 512, tc::BN));
  const int tiles = 2 * (int)cdiv(M, tc::BM);
  const int grid = tiles < num_sms ? tiles : num_sms;
  tc::tc_gemm_kernel<<<grid, tc::THREADS, tc::SMEM_BYTES, st>>>(ma, mb, M, e, err_flag);
  SURFD_CHECK_LAUNCH();
  return 0;
}

}  // namespace surfd
