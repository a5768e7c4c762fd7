// unet.cu -- the denoiser (MDM / UNetModel forward) as an op program interpreted on the device, and the
// reverse-diffusion loop around it, captured once as a CUDA graph and replayed per step.
//
// Reference being replaced (read-only, /root/reference):
//   models/openaimodel.py:710-749   UNetModel.forward          :255-275 ResBlock._forward
//   models/openaimodel.py:318-324   AttentionBlock._forward    :356-372 QKVAttentionLegacy.forward
//   models/openaimodel.py:91-119,134-160  Upsample (nearest x2 + conv3) / Downsample (conv3 stride 2)
//   utils/ldm_utils.py:165-185      timestep_embedding         :244-249 GroupNorm32(32, C), eps 1e-5
//   diffusion/gaussian_diffusion.py:471-520 p_sample, :234-256 q_posterior_mean_variance, :635-708 loop
//   diffusion/respace.py:116-132    _WrappedModel timestep remap
//
// Layout: activations channels-last [B][T][C] fp32 (the reference is [B][C][T]); conv weights repacked to
// [tap][Cout][Cin] so a k=3 convolution is three token-shifted GEMMs over contiguous channel vectors; the
// UNet's skip concatenation and the 1x1 skip convolution are extra K segments of the same GEMM; the 22
// emb_layers linears are one batched GEMM per step.  The reference issues ~450 library kernels and 6 host->device
// table uploads per step; here a step is one graph launch of ~170 small kernels reading the step index from
// device memory.
#include <math.h>
#include <stdlib.h>
#include <vector>

#include <cooperative_groups.h>

#include "common.cuh"

namespace cg = cooperative_groups;

namespace surfd {

// Programmatic dependent launch: every kernel of the step first lets its successor start launching (its CTAs become
// resident and park at their own wait) and then waits for its predecessor to finish and flush.  This hides the
// kernel-to-kernel launch latency of the ~170-node step graph.  Both instructions are no-ops for ordinary launches.
#define PDL_PROLOGUE()                                            \
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); \
  asm volatile("griddepcontrol.wait;" ::: "memory")

enum { OP_GN = 1, OP_CONV = 2, OP_ATTN = 3, OP_INCONV = 4, OP_OUTCONV = 5 };
constexpr int REC = 32;
constexpr int EMB = 896;
constexpr int TCH = 224;
constexpr int CTX = 512;

// ------------------------------------------------------------------------------------------------
// GroupNorm(32, C) (+ SiLU) over a virtual channel concat [in1 | in2]; one CTA per (batch, group)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
gn_kernel(const float* __restrict__ in1, int C1, const float* __restrict__ in2, int C2, int T, const float* __restrict__ gamma,
          const float* __restrict__ beta, int silu, float* __restrict__ out, float* __restrict__ raw) {
  PDL_PROLOGUE();
  const int C = C1 + C2;
  const int cg = C / 32;
  const int b = blockIdx.x >> 5, g = blockIdx.x & 31;
  const int n = T * cg;
  __shared__ float red[4];
  auto load = [&](int idx) -> float {
    const int t = idx / cg, c = g * cg + idx % cg;
    return c < C1 ? in1[((size_t)b * T + t) * C1 + c] : in2[((size_t)b * T + t) * C2 + (c - C1)];
  };
  auto block_sum = [&](float v) -> float {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    return red[0] + red[1] + red[2] + red[3];
  };
  // the group's values are read once and kept in registers for the three passes (n <= 32 * 128; longer groups re-read)
  constexpr int GN_CACHE = 32;
  float xc[GN_CACHE];
#pragma unroll
  for (int k = 0; k < GN_CACHE; ++k) { const int i = threadIdx.x + 128 * k; xc[k] = i < n ? load(i) : 0.f; }
  float s = 0.f;
#pragma unroll
  for (int k = 0; k < GN_CACHE; ++k) if ((int)threadIdx.x + 128 * k < n) s += xc[k];
  for (int i = threadIdx.x + 128 * GN_CACHE; i < n; i += 128) s += load(i);
  const float mean = block_sum(s) / (float)n;
  float q = 0.f;
#pragma unroll
  for (int k = 0; k < GN_CACHE; ++k) if ((int)threadIdx.x + 128 * k < n) { const float d = xc[k] - mean; q += d * d; }
  for (int i = threadIdx.x + 128 * GN_CACHE; i < n; i += 128) { const float d = load(i) - mean; q += d * d; }
  const float var = block_sum(q) / (float)n;
  const float rstd = 1.0f / sqrtf(var + 1e-5f);
  auto emit = [&](int i, float x) {
    const int t = i / cg, c = g * cg + i % cg;
    float y = (x - mean) * rstd * gamma[c] + beta[c];
    if (silu) y = y / (1.0f + expf(-y));
    const size_t o = ((size_t)b * T + t) * C + c;
    out[o] = y;
    if (raw) raw[o] = x;
  };
#pragma unroll
  for (int k = 0; k < GN_CACHE; ++k) if ((int)threadIdx.x + 128 * k < n) emit((int)threadIdx.x + 128 * k, xc[k]);
  for (int i = threadIdx.x + 128 * GN_CACHE; i < n; i += 128) emit(i, load(i));
}

// ------------------------------------------------------------------------------------------------
// Token GEMM: out[m][n] = sum over K segments/taps of A[token(m,tap)][ci] * W[tap][n][ci]  (+bias +emb +residual)
// CTA tile 32 tokens x 32 outputs; 8 warps split the K chunks (32 channels each) and reduce through smem.
// ------------------------------------------------------------------------------------------------
struct Seg {
  const float* A;  // [B][T_in][Cin]  (wide units: channels [0, C1) of a virtual concat [A | A2])
  const float* W;  // [taps][N][Cin]
  int Cin, taps, stride, up, T_in;
  const float* A2; // wide units only: channels [C1, Cin) come from A2 [B][T_in][Cin - C1]; C1 == Cin: single source
  int C1;
  const float* Wh; // tcgen05 units only: W with every 32-float block replaced by its fp16 split [32 x hi | 32 x lo] (same offsets)
};
struct ConvArgs {
  Seg seg[2];
  int nseg;
  int B, T_out, N;
  const float* bias;      // [N]
  const float* emb;       // [B][emb_ld] (already offset to this block's columns) or null
  int emb_ld;
  const float* residual;  // [B][T_out][N] or null
  float* out;             // [B][T_out][N]
  // wide units only: GroupNorm(32, Cin) (+ SiLU) of segment 0's input fused into this GEMM (pn_on): the unit computes the
  // statistics of the (sample, group) pairs its K slice touches and normalises the token rows in shared memory
  const float* pn_gamma;
  const float* pn_beta;
  int pn_on, pn_silu, pn_cg, pn_logT;
  float pn_inv_cg;
};

constexpr int CT = 32;        // tile edge
constexpr int CTP = CT + 4;   // padded row (floats)
constexpr int KSPLIT = 8;     // max CTAs per cluster: the K reduction is split over a thread-block cluster (1, 2, 4 or 8)
constexpr int CONV_STAGES = 3; // per-warp cp.async stages of the chunk stream (two chunks in flight behind the one being multiplied)

// One output tile (32 tokens x 32 channels) is owned by a cluster of KSPLIT CTAs.  K chunks (32 input channels of one
// tap of one segment) are dealt round-robin to the 8 x 8 = 64 warps of the cluster, so even the deepest layers
// (M = 32 tokens, K = 5376) put 200+ CTAs on the machine and every weight byte is fetched exactly once.  Partial
// tiles are reduced first across the warps of a CTA (shared memory) and then across the cluster through distributed
// shared memory in a fixed order -- deterministic, no atomics, no second kernel.
// MODE 0: fp32 FFMA.  MODE 1: tensor cores, mma.sync m16n8k8 with the 3xTF32 split (a = hi + lo, products hi*hi + hi*lo +
// lo*hi accumulate in fp32: ~2^-21 relative, fp32-class results at a tenth of the issue slots).  MODE 2: single TF32 pass.
__device__ __forceinline__ uint32_t to_tf32(float x) {
  uint32_t u;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(x));
  return u;
}
__device__ __forceinline__ void mma_tf32(float* c, const uint32_t* a, const uint32_t* b) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

// K slice `crank` of `ks` of output tile (n0, m0): every warp accumulates its chunks into register fragments.
template <int MODE>
__device__ __forceinline__ void conv_accumulate(const ConvArgs& a, int n0, int m0, int ks, int crank, float* smem, float (&acc)[8][4],
                                                long long* prof = nullptr, bool pdl_wait = false) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* stage0 = smem + warp * (CONV_STAGES * 2 * CT * CTP);   // per warp: CONV_STAGES x (A tile | W tile)
  const int M = a.B * a.T_out;

  // Copy pattern: one cp.async instruction moves 4 rows x 128 B (lanes 0-7 cover row 0, 8-15 row 1, ...), i.e. 4 full
  // lines per instruction.  (One row per lane would touch 32 different lines per instruction and saturate the L1
  // wavefront pipeline: measured 2.5 us per 8 KB chunk.)  This lane serves rows lr + 4j, j = 0..7, 16-byte piece lp.
  const int lr = lane >> 3, lp = lane & 7;
  int rbv[8], rlv[8];
  unsigned rokv = 0;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int m_row = m0 + lr + 4 * j;
    rbv[j] = m_row / a.T_out;
    rlv[j] = m_row - rbv[j] * a.T_out;
    rokv |= (m_row < M ? 1u : 0u) << j;
  }

  const int ly = lane >> 3, lx = lane & 7;   // FFMA mapping: rows ly + 4i, cols lx + 8j
  const int fg = lane >> 2, ft = lane & 3;   // MMA fragment mapping: group id / thread in group
  // acc: MODE 0: [i][j];  MODE 1/2: [mt*4 + nt][c0..c3]
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  const int my_slot = warp * ks + crank;       // chunk c belongs to slot c % (8*ks); consecutive chunks go to different CTAs
  // This warp owns the flattened chunks f = my_slot, my_slot + 8*ks, ...  (f enumerates segment, tap, 32-channel block).
  // Chunks stream through a per-warp cp.async double buffer: while chunk f is multiplied, the 32 token rows and 32
  // weight rows (128 B each) of chunk f + 8*ks are already in flight (weights from HBM: 553 MB per step >> L2; the
  // activation rows are L2-resident).  Copies bypass L1 (.cg), so the persistent kernel never sees stale activations.
  const int n_chunks0 = a.seg[0].taps * (a.seg[0].Cin / CT);
  const int n_chunks = n_chunks0 + (a.nseg > 1 ? a.seg[1].taps * (a.seg[1].Cin / CT) : 0);
  const int stride_f = 8 * ks;
  // what = 1: weight rows only, 2: token rows only, 3: both
  auto issue = [&](int f, float* As_, float* Ws_, int what) {
    const int s = f < n_chunks0 ? 0 : 1;
    const Seg& sg = a.seg[s];
    const int g = s ? f - n_chunks0 : f;
    const int cpt = sg.Cin / CT;
    const int tap = g / cpt, c = g - tap * cpt;
    const int T_eff = sg.up ? 2 * sg.T_in : sg.T_in;
    const float* abase = sg.A + c * CT + 4 * lp;
    const float* wbase = sg.W + ((size_t)tap * a.N + n0 + lr) * sg.Cin + c * CT + 4 * lp;
    const uint32_t da = (uint32_t)__cvta_generic_to_shared(As_ + lr * CTP + 4 * lp);
    const uint32_t dw = (uint32_t)__cvta_generic_to_shared(Ws_ + lr * CTP + 4 * lp);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      if (what & 2) {
        const int src = rlv[j] * sg.stride + tap - (sg.taps >> 1);
        const bool ok = ((rokv >> j) & 1u) && src >= 0 && src < T_eff;
        const int st = sg.up ? (src >> 1) : src;
        const float* arow = abase + (ok ? (size_t)rbv[j] * sg.T_in + st : (size_t)0) * sg.Cin;
        const int asz = ok ? 16 : 0;   // src-size 0: the 16 destination bytes are zero-filled (padding rows / taps outside the sequence)
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(da + j * (4 * CTP * 4)), "l"(arow), "r"(asz) : "memory");
      }
      if (what & 1)
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dw + j * (4 * CTP * 4)), "l"(wbase + (size_t)(4 * j) * sg.Cin) : "memory");
    }
  };
  int f = my_slot;
  int stg = 0;
  const bool pf = prof != nullptr && threadIdx.x == 0;
  // Programmatic dependent launch: the weight rows do not depend on the previous kernel, so they are requested before
  // griddepcontrol.wait (their HBM latency overlaps the predecessor's tail); the token rows follow the wait.
#pragma unroll
  for (int p = 0; p < CONV_STAGES - 1; ++p) {
    const int fp = f + p * stride_f;
    if (fp < n_chunks) issue(fp, stage0 + p * (2 * CT * CTP), stage0 + p * (2 * CT * CTP) + CT * CTP, 1);
  }
  if (pdl_wait) asm volatile("griddepcontrol.wait;" ::: "memory");
#pragma unroll
  for (int p = 0; p < CONV_STAGES - 1; ++p) {
    const int fp = f + p * stride_f;
    if (fp < n_chunks) issue(fp, stage0 + p * (2 * CT * CTP), stage0 + p * (2 * CT * CTP) + CT * CTP, 2);
    asm volatile("cp.async.commit_group;" ::: "memory");
  }
  for (; f < n_chunks; f += stride_f) {
    {
      {
        float* As = stage0 + stg * (2 * CT * CTP);
        float* Ws = As + CT * CTP;
        const int fn = f + (CONV_STAGES - 1) * stride_f;
        const int sn = stg + CONV_STAGES - 1 >= CONV_STAGES ? stg - 1 : stg + CONV_STAGES - 1;
        if (fn < n_chunks) issue(fn, stage0 + sn * (2 * CT * CTP), stage0 + sn * (2 * CT * CTP) + CT * CTP, 3);
        asm volatile("cp.async.commit_group;" ::: "memory");
        long long tw0 = 0;
        if (pf) tw0 = clock64();
        asm volatile("cp.async.wait_group %0;" ::"n"(CONV_STAGES - 1) : "memory");
        __syncwarp();
        if (pf) { prof[27] += clock64() - tw0; prof[26] += 1; }
        if (MODE == 0) {
#pragma unroll
          for (int kk = 0; kk < CT; kk += 4) {
            float4 wj[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) wj[j] = *reinterpret_cast<const float4*>(Ws + (lx + 8 * j) * CTP + kk);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const float4 ai = *reinterpret_cast<const float4*>(As + (ly + 4 * i) * CTP + kk);
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                acc[i][j] = fmaf(ai.x, wj[j].x, acc[i][j]);
                acc[i][j] = fmaf(ai.y, wj[j].y, acc[i][j]);
                acc[i][j] = fmaf(ai.z, wj[j].z, acc[i][j]);
                acc[i][j] = fmaf(ai.w, wj[j].w, acc[i][j]);
              }
            }
          }
        } else {
#pragma unroll
          for (int k0 = 0; k0 < CT; k0 += 8) {
            // A fragments (16x8, row-major): a0 (g, t) a1 (g+8, t) a2 (g, t+4) a3 (g+8, t+4)
            uint32_t ahi[2][4], alo[2][4];
#pragma unroll
            for (int mt = 0; mt < 2; ++mt) {
              const float* ap = As + (mt * 16 + fg) * CTP + k0 + ft;
              const float x[4] = {ap[0], ap[8 * CTP], ap[4], ap[8 * CTP + 4]};
#pragma unroll
              for (int r = 0; r < 4; ++r) {
                ahi[mt][r] = to_tf32(x[r]);
                if (MODE == 1) alo[mt][r] = to_tf32(x[r] - __uint_as_float(ahi[mt][r]));
              }
            }
#pragma unroll
            for (int nt = 0; nt < 4; ++nt) {
              // B fragment (8x8, k x n): b0 (k = t, n = g) b1 (k = t+4, n = g); W tile is [n][k]
              const float* bp = Ws + (nt * 8 + fg) * CTP + k0 + ft;
              const float y[2] = {bp[0], bp[4]};
              uint32_t bhi[2], blo[2];
#pragma unroll
              for (int r = 0; r < 2; ++r) {
                bhi[r] = to_tf32(y[r]);
                if (MODE == 1) blo[r] = to_tf32(y[r] - __uint_as_float(bhi[r]));
              }
#pragma unroll
              for (int mt = 0; mt < 2; ++mt) {
                float* c = acc[mt * 4 + nt];
                if (MODE == 1) { mma_tf32(c, alo[mt], bhi); mma_tf32(c, ahi[mt], blo); }
                mma_tf32(c, ahi[mt], bhi);
              }
            }
          }
        }
        __syncwarp();   // every lane is done reading this stage before the next iteration's copy overwrites it
        stg = stg + 1 == CONV_STAGES ? 0 : stg + 1;
      }
    }
  }
  asm volatile("cp.async.wait_group 0;" ::: "memory");
}

// (1) cross-warp reduction inside the CTA: red[warp][row][col] (rows padded to 33) -> part[row][col]; returns `part`
// (not yet synchronised: the caller's next barrier publishes it)
template <int MODE>
__device__ __forceinline__ float* conv_cta_partial(float (&acc)[8][4], float* smem) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ly = lane >> 3, lx = lane & 7;
  const int fg = lane >> 2, ft = lane & 3;
  __syncthreads();
  float* red = smem;
  float* part = smem + 8 * CT * 33;   // [32][32] CTA partial, read by the other CTAs of the cluster
  if (MODE == 0) {
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) red[(warp * CT + (ly + 4 * i)) * 33 + (lx + 8 * j)] = acc[i][j];
  } else {
    // accumulator fragment: c0 (g, 2t) c1 (g, 2t+1) c2 (g+8, 2t) c3 (g+8, 2t+1)
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
      for (int nt = 0; nt < 4; ++nt)
#pragma unroll
        for (int r = 0; r < 4; ++r)
          red[(warp * CT + (mt * 16 + fg + (r >> 1) * 8)) * 33 + (nt * 8 + 2 * ft + (r & 1))] = acc[mt * 4 + nt][r];
  }
  __syncthreads();
  {
    const int r = threadIdx.x >> 3, c0 = (threadIdx.x & 7) * 4;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float sum = 0.f;
#pragma unroll
      for (int w = 0; w < 8; ++w) sum += red[(w * CT + r) * 33 + c0 + j];
      part[r * CT + c0 + j] = sum;
    }
  }
  return part;
}

// epilogue of one output element: + bias (+ per-sample embedding column) (+ residual), in this order
__device__ __forceinline__ float conv_epilogue_operands(const ConvArgs& a, int m0, int n0, int idx, float v) {
  const int r = idx >> 5, col = idx & 31;
  const int m = m0 + r;
  if (m < a.B * a.T_out) {
    const int n = n0 + col;
    const int b = m / a.T_out;
    v += a.bias[n];
    if (a.emb) v += a.emb[(size_t)b * a.emb_ld + n];
    if (a.residual) v += a.residual[(size_t)m * a.N + n];
  }
  return v;
}
__device__ __forceinline__ void conv_store_value(const ConvArgs& a, int m0, int n0, int idx, float v) {
  const int r = idx >> 5, col = idx & 31;
  const int m = m0 + r;
  if (m < a.B * a.T_out) a.out[(size_t)m * a.N + n0 + col] = v;
}
__device__ __forceinline__ void conv_store(const ConvArgs& a, int m0, int n0, int idx, float v) {
  conv_store_value(a, m0, n0, idx, conv_epilogue_operands(a, m0, n0, idx, v));
}

template <int MODE>
__global__ void __launch_bounds__(256)
conv_gemm_kernel(ConvArgs a) {
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");   // the wait sits inside conv_accumulate, after the weight requests
  extern __shared__ __align__(16) float smem[];
  cg::cluster_group cluster = cg::this_cluster();
  const int crank = (int)cluster.block_rank();
  const int ks = (int)cluster.num_blocks();    // K split of this launch: 1, 2, 4 or 8 CTAs per output tile
  const int n0 = blockIdx.x * CT;
  const int m0 = blockIdx.y * CT;
  float acc[8][4];
  conv_accumulate<MODE>(a, n0, m0, ks, crank, smem, acc, nullptr, true);
  float* part = conv_cta_partial<MODE>(acc, smem);
  // (2) cross-CTA reduction over distributed shared memory: CTA `crank` finishes its 1/ks share of the 32x32 tile
  cluster.sync();
  {
    const int per = CT * CT / ks;
    // <= 4 elements per thread: all remote partials and all epilogue operands (bias, embedding column, residual) are
    // requested before the first store -- a load placed after a store to `out` could not be moved ahead of it by the compiler
    float v[4], add[4];
    int idxs[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int idx = crank * per + (int)threadIdx.x + 256 * i;
      idxs[i] = idx < (crank + 1) * per ? idx : -1;
      float pv[KSPLIT];
#pragma unroll
      for (int j = 0; j < KSPLIT; ++j) pv[j] = (idxs[i] >= 0 && j < ks) ? cluster.map_shared_rank(part, j)[idx] : 0.f;
      float acc1 = 0.f;
#pragma unroll
      for (int j = 0; j < KSPLIT; ++j)
        if (j < ks) acc1 += pv[j];
      v[i] = acc1;
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) add[i] = idxs[i] >= 0 ? conv_epilogue_operands(a, m0, n0, idxs[i], v[i]) : 0.f;
#pragma unroll
    for (int i = 0; i < 4; ++i)
      if (idxs[i] >= 0) conv_store_value(a, m0, n0, idxs[i], add[i]);
  }
  cluster.sync();   // keep this CTA's shared memory alive until every peer has read it
}

// ------------------------------------------------------------------------------------------------
// QKVAttentionLegacy: one (batch, head) unit per thread group.  qkv [B][T][3C] with per-head [q|k|v] channel interleave.
// `nthr` threads (a multiple of 32) with ids `tid` cooperate; `sync()` is their barrier.  Every output element is produced
// by one thread (or one warp for a softmax row) in a fixed order, so the result does not depend on nthr.
// Shared memory: q, k [T][ch+1] (padded: lanes walk k rows), v [T][ch], w [T][T+1].
// ------------------------------------------------------------------------------------------------
__host__ __device__ inline int attn_smem_floats(int T, int ch) { return 2 * T * (ch + 1) + T * ch + T * (T + 1); }

// (i / d, i % d) for small non-negative i without an integer division: a lone warp pays ~35 dependent instructions per
// division by a run-time value, and the unit below did one per element
__device__ __forceinline__ int attn_div(int x, int d, float inv) {
  int q = (int)((float)x * inv);
  const int r = x - q * d;
  if (r < 0) --q;
  else if (r >= d) ++q;
  return q;
}

template <typename Sync, typename Load>
__device__ __forceinline__ void attn_unit(const float* qkv, int C, int T, int heads, float scale, float* out, int b, int h, float* sm,
                                          int tid, int nthr, Sync sync, Load load) {
  const int ch = C / heads;
  const int chp = ch + 1;
  const float inv_ch = 1.0f / (float)ch, inv_T = 1.0f / (float)T;
  float* q = sm;               // [T][chp] (scaled)
  float* k = q + T * chp;      // [T][chp] (scaled)
  float* v = k + T * chp;      // [T][ch]
  float* w = v + T * ch;       // [T][T+1]
  const float* base = qkv + (size_t)b * T * 3 * C + (size_t)h * 3 * ch;
  for (int i0 = tid; i0 < T * ch; i0 += 4 * nthr) {   // four (q, k, v) triples in flight per thread
    float qv[4], kv[4], vv[4];
    int tt[4], cc[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int i = i0 + u * nthr;
      qv[u] = kv[u] = vv[u] = 0.f;
      tt[u] = cc[u] = 0;
      if (i < T * ch) {
        const int t = attn_div(i, ch, inv_ch), c = i - t * ch;
        tt[u] = t; cc[u] = c;
        const float* p = base + (size_t)t * 3 * C + c;
        qv[u] = load(p, b * T + t, h * 3 * ch + c); kv[u] = load(p + ch, b * T + t, h * 3 * ch + ch + c);
        vv[u] = load(p + 2 * ch, b * T + t, h * 3 * ch + 2 * ch + c);
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int i = i0 + u * nthr;
      if (i < T * ch) {
        q[tt[u] * chp + cc[u]] = qv[u] * scale;
        k[tt[u] * chp + cc[u]] = kv[u] * scale;
        v[i] = vv[u];
      }
    }
  }
  sync();
  // scores: one (query, key) pair per thread; four independent partial sums over the channels (a single accumulator is a
  // chain of ch dependent FMAs: 112 x 4 cycles at the 4-token levels, where only 16 threads have work)
  for (int i = tid; i < T * T; i += nthr) {
    const int t = attn_div(i, T, inv_T), s_ = i - t * T;
    const float* qr = q + t * chp;
    const float* kr = k + s_ * chp;
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
    int c = 0;
    for (; c + 4 <= ch; c += 4) {
      a0 = fmaf(qr[c], kr[c], a0); a1 = fmaf(qr[c + 1], kr[c + 1], a1);
      a2 = fmaf(qr[c + 2], kr[c + 2], a2); a3 = fmaf(qr[c + 3], kr[c + 3], a3);
    }
    for (; c < ch; ++c) a0 = fmaf(qr[c], kr[c], a0);
    w[t * (T + 1) + s_] = (a0 + a1) + (a2 + a3);
  }
  sync();
  // softmax: one warp per row, lanes over the keys
  const int lane = tid & 31;
  for (int t = tid >> 5; t < T; t += nthr >> 5) {
    float* wr = w + t * (T + 1);
    float mx = -INFINITY;
    for (int s = lane; s < T; s += 32) mx = fmaxf(mx, wr[s]);
#pragma unroll
    for (int of = 16; of > 0; of >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, of));
    float sum = 0.f;
    for (int s = lane; s < T; s += 32) { const float e = expf(wr[s] - mx); wr[s] = e; sum += e; }
#pragma unroll
    for (int of = 16; of > 0; of >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, of);
    const float inv = 1.0f / sum;
    for (int s = lane; s < T; s += 32) wr[s] *= inv;
  }
  sync();
  for (int i = tid; i < T * ch; i += nthr) {
    const int t = attn_div(i, ch, inv_ch), c = i - t * ch;
    const float* wr = w + t * (T + 1);
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
    int s = 0;
    for (; s + 4 <= T; s += 4) {
      a0 = fmaf(wr[s], v[s * ch + c], a0); a1 = fmaf(wr[s + 1], v[(s + 1) * ch + c], a1);
      a2 = fmaf(wr[s + 2], v[(s + 2) * ch + c], a2); a3 = fmaf(wr[s + 3], v[(s + 3) * ch + c], a3);
    }
    for (; s < T; ++s) a0 = fmaf(wr[s], v[s * ch + c], a0);
    out[((size_t)b * T + t) * C + h * ch + c] = (a0 + a1) + (a2 + a3);
  }
  sync();   // the buffers may be reused for the next unit
}

__global__ void __launch_bounds__(128)
attn_kernel(const float* __restrict__ qkv, int C, int T, int heads, float scale, float* __restrict__ out) {
  PDL_PROLOGUE();
  extern __shared__ __align__(16) float sm[];
  attn_unit(qkv, C, T, heads, scale, out, (int)blockIdx.x / heads, (int)blockIdx.x % heads, sm, (int)threadIdx.x, 128,
            [] { __syncthreads(); }, [](const float* p, int, int) { return *p; });
}

// first conv (1 -> 224, k=3, pad 1) and last conv (224 -> 1)
__global__ void inconv_kernel(const float* __restrict__ x, int B, int L, int N, const float* __restrict__ W /*[3][N][1]*/,
                              const float* __restrict__ bias, float* __restrict__ out) {
  PDL_PROLOGUE();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * L * N) return;
  const int n = i % N, l = (i / N) % L, b = i / (N * L);
  const float* xr = x + (size_t)b * L;
  float acc = 0.f;
  if (l > 0) acc = fmaf(W[n], xr[l - 1], acc);
  acc = fmaf(W[N + n], xr[l], acc);
  if (l < L - 1) acc = fmaf(W[2 * N + n], xr[l + 1], acc);
  out[i] = acc + bias[n];
}

__global__ void outconv_kernel(const float* __restrict__ a, int B, int L, int C, const float* __restrict__ W /*[3][1][C]*/,
                               const float* __restrict__ bias, float* __restrict__ out) {
  PDL_PROLOGUE();
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= B * L) return;
  const int l = warp % L, b = warp / L;
  float acc = 0.f;
  for (int tap = 0; tap < 3; ++tap) {
    const int src = l + tap - 1;
    if (src < 0 || src >= L) continue;
    const float* ar = a + ((size_t)b * L + src) * C;
    const float* wr = W + (size_t)tap * C;
    for (int c = lane; c < C; c += 32) acc = fmaf(ar[c], wr[c], acc);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if (lane == 0) out[warp] = acc + bias[0];
}

// ---- embedding path ----------------------------------------------------------------------------
// timestep_embedding: [cos(t*f_k) | sin(t*f_k)], f_k = exp(-ln(1e4) * k / 112) in fp32 (utils/ldm_utils.py:165-185)
__global__ void temb_kernel(const int64_t* __restrict__ t, int B, float* __restrict__ out) {
  PDL_PROLOGUE();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * (TCH / 2)) return;
  const int k = i % (TCH / 2), b = i / (TCH / 2);
  const float c = -9.210340371976184f;  // (float)(-math.log(10000))
  const float f = expf(__fdiv_rn(__fmul_rn(c, (float)k), 112.0f));
  const float arg = __fmul_rn((float)t[b], f);
  out[(size_t)b * TCH + k] = cosf(arg);
  out[(size_t)b * TCH + TCH / 2 + k] = sinf(arg);
}

// out[m][n] (+)= sum_k act(in[m][k]) * W[n][k] + bias[n];  M small (<= 8 rows per CTA pass); one warp per n.
__global__ void __launch_bounds__(256)
linear_rows_kernel(const float* __restrict__ in, int M, int K, const float* __restrict__ W, const float* __restrict__ bias, int N,
                   int in_silu, int out_silu, int accumulate, float* __restrict__ out) {
  PDL_PROLOGUE();
  const int n = blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  const int mb = blockIdx.y * 8;
  if (n >= N) return;
  float acc[8];
#pragma unroll
  for (int r = 0; r < 8; ++r) acc[r] = 0.f;
  const float* wr = W + (size_t)n * K;
  for (int k = lane * 4; k < K; k += 128) {
    const float4 w = *reinterpret_cast<const float4*>(wr + k);
#pragma unroll
    for (int r = 0; r < 8; ++r) {
      if (mb + r < M) {
        float4 x = *reinterpret_cast<const float4*>(in + (size_t)(mb + r) * K + k);
        if (in_silu) {
          x.x = x.x / (1.0f + expf(-x.x)); x.y = x.y / (1.0f + expf(-x.y));
          x.z = x.z / (1.0f + expf(-x.z)); x.w = x.w / (1.0f + expf(-x.w));
        }
        acc[r] = fmaf(x.x, w.x, acc[r]); acc[r] = fmaf(x.y, w.y, acc[r]);
        acc[r] = fmaf(x.z, w.z, acc[r]); acc[r] = fmaf(x.w, w.w, acc[r]);
      }
    }
  }
#pragma unroll
  for (int r = 0; r < 8; ++r)
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc[r] += __shfl_xor_sync(0xffffffffu, acc[r], o);
  if (lane == 0) {
    const float bv = bias[n];
#pragma unroll
    for (int r = 0; r < 8; ++r) {
      if (mb + r < M) {
        float v = acc[r] + bv;
        if (out_silu) v = v / (1.0f + expf(-v));
        float* o = out + (size_t)(mb + r) * N + n;
        *o = accumulate ? (*o + v) : v;
      }
    }
  }
}

__global__ void label_add_kernel(const int64_t* __restrict__ y, int B, const float* __restrict__ table, float* __restrict__ emb) {
  PDL_PROLOGUE();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * EMB) return;
  const int b = i / EMB, c = i % EMB;
  emb[i] += table[(size_t)y[b] * EMB + c];
}

// ---- reverse-diffusion step bookkeeping -----------------------------------------------------------
struct StepState {
  int iter;        // loop iteration 0..n_steps-1 (step index = n_steps-1-iter)
  int n_steps;
};

__global__ void step_begin_kernel(const StepState* __restrict__ st, const int64_t* __restrict__ tmap, int B, int64_t* __restrict__ t_cur) {
  PDL_PROLOGUE();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B) return;
  const int idx = st->n_steps - 1 - st->iter;
  t_cur[i] = tmap[idx];   // _WrappedModel: new_ts = map_tensor[ts]
}

// sample = (c1*x0 + c2*x_t) + ((t != 0) * std) * noise   -- unfused multiplies/adds like the reference's torch ops.
// x0b != null: classifier-free guidance replay  x0 = x0b + scale * (x0a - x0b)   (models/cfg_sampler.py:19-26)
__global__ void ddpm_update_kernel(StepState* __restrict__ st, const float* __restrict__ coef, const float* __restrict__ x0a,
                                   const float* __restrict__ x0b, float scale, const float* __restrict__ noise, int64_t noise_row_stride, int n,
                                   float* __restrict__ x) {
  PDL_PROLOGUE();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int iter = st->iter, ns = st->n_steps;
  const int idx = ns - 1 - iter;
  if (i < n) {
    float x0 = x0a[i];
    if (x0b) x0 = __fadd_rn(x0b[i], __fmul_rn(scale, __fsub_rn(x0a[i], x0b[i])));
    const float c1 = coef[idx], c2 = coef[ns + idx], sd = coef[2 * ns + idx];
    const float mean = __fadd_rn(__fmul_rn(c1, x0), __fmul_rn(c2, x[i]));
    const float mask = idx != 0 ? 1.0f : 0.0f;
    const float nz = noise[(size_t)(1 + iter) * noise_row_stride + i];   // noise is [n_steps+1][B_total][L]; this lane owns a column block
    x[i] = __fadd_rn(mean, __fmul_rn(__fmul_rn(mask, sd), nz));
  }
}

__global__ void step_advance_kernel(StepState* st) {
  PDL_PROLOGUE();
  st->iter += 1;
}


// =================================================================================================================
// Persistent sampler: the whole reverse-diffusion loop (n_steps x ~170 ops) in ONE cooperative kernel.
//
// The step is latency-bound: ~170 dependent small ops on <= 256 tokens, 553 MB of weights streamed per step.  As a
// graph of kernels every op pays a launch + ramp + drain (~13 us per node at batch 8).  Here one CTA per SM stays
// resident, interprets the same op list, and separates ops by a grid barrier (one atomic + a spin, ~1.5 us); while a
// CTA waits at the barrier the weight lines of its NEXT op's units are already on their way to L2.
//   * token GEMM: same chunk -> (warp, K slice) mapping and the same summation order as conv_gemm_kernel; the K
//     slices of a tile meet through an L2 scratch tile + a per-tile semaphore (last arriver reduces in slice order and
//     runs the epilogue), so results are bit-identical to the cluster/DSMEM kernel and independent of the grid size.
//   * GroupNorm / attention: two 128-thread halves per CTA work on different (batch, group|head) units (named barriers),
//     arithmetic identical to gn_kernel / attn_kernel.
//   * embedding linears: input rows staged in shared memory (SiLU applied once), two weight rows in flight per warp.
//   * the DDPM update is the epilogue of the last op (each output element needs only its own x0).
// Activations are read with plain loads after the barrier's fence (weights are immutable and may use the read-only path).
// =================================================================================================================
constexpr int P_SYNC_WORDS = 64;   // [0] barrier counter, [1] abort flag, [2] an activation left the fp16 range of the split products
constexpr int P_MAX_KS = 16;   // most K slices per output tile in the persistent kernel
enum { P_EMB1 = 1, P_LIN = 2, P_INCONV = 3, P_GN = 4, P_CONV = 5, P_ATTN = 6, P_OUTCONV = 7 };

// A token-GEMM output that its producer left as K-slice partial tiles ("deferred"): the consumer (GroupNorm or attention)
// sums the slices in slice order and applies the GEMM's epilogue while it loads.  ks == 0: a plain [B][T][C] tensor.
struct PartSrc {
  const float* part;          // [tile][slice][32][128]
  const float* bias;
  const float* emb;           // [B][emb_ld] or null
  const float* res;           // [B*T][N] or null
  int ks, tiles_n, emb_ld, N;
  const unsigned* p2p;        // non-null: the producing GEMM is NOT followed by a grid barrier -- its units count their arrivals per
                              // output tile here (monotonic over the run) and the consumer waits for exactly the tiles it reads
};

struct POp {
  int type;
  int ks, tiles_n, tiles_m;   // P_CONV: K split, tile grid
  ConvArgs conv;
  const float* in0;           // activations
  const float* in1;
  const float* w0;            // weights / bias / tables (immutable)
  const float* w1;
  const float* w2;
  const int64_t* lab;
  float* out0;
  float* out1;
  int i0, i1, i2, i3, i4, i5;
  float f0;
  float inv_ks, inv_tn, inv_T; // P_CONV: 1 / ks, 1 / tiles_n, 1 / T_out (unit decoding without integer divisions, see div_small)
  int wide;                   // P_CONV: 1 = wide unit (32 x 128 tile, distributed slice reduction), 0 = narrow unit
  int defer;                  // P_CONV (wide, ks > 1): leave the K-slice partial tiles to the next op (no exchange inside this op)
  unsigned* p2p;              // P_CONV (deferred, weight-stream flow): per-tile arrival counters; no grid barrier behind this op
  PartSrc ps;                 // P_GN / P_ATTN: the first input is a deferred token-GEMM output
};

struct PersistArgs {
  const POp* ops;
  int n_emb, n_prog;          // per step: ops[0, n_emb) once, then ops[n_emb, n_emb + n_prog) once per pass
  int n_steps, B, L, n_pass;
  const int64_t* tmap;
  const float* coef;
  const float* noise;
  long long noise_stride;
  float guidance;
  float* x;                   // [B][L] current sample, updated in place
  float* x0a;                 // [B][L] first-pass prediction when n_pass == 2
  float* partials;            // [tile][slice][32*32] K-slice partial tiles
  unsigned* sems;             // per-tile arrival counters: [2 banks for the wide units | narrow units (zero between ops)] x sem_bank
  int sem_bank;
  int use_tc;                 // precision mode 1: wide units run on tcgen05 (ops with wide == 2); the kernel holds 64 TMEM columns
  unsigned* sync;             // [0] barrier counter, [1] abort flag
  long long* prof;            // diagnostics (may be null): [cta 0 | cta G-1][op type][body cycles, barrier cycles, count]
  const uint8_t* const* wlist;   // weight stream (use_tma): image pointers of one pass, per CTA in consumption order
  const int* wl_start;           // [grid] first entry of the CTA's list
  const int* wl_count;           // [grid] its length
  int use_tma;                // tcgen05 units take their weights from the packed stream through cp.async.bulk (see WJob)
  int pair_B;                 // classifier-free guidance as ONE pass over 2 * pair_B samples (rows b and pair_B + b are the two
                              // forwards of sample b; the out-conv combines them); 0: n_pass sequential passes over B samples
};

// Weight stream of the tcgen05 units.  The weights of a unit are immutable and the op list is static, so they do not have
// to wait for the op chain: at load time every (op, output tile, K slice) gets its chunk pairs packed as ready-made 32 KB
// shared-memory images (fp16 hi plane | lo plane, 128 rows x 128 B each, SWIZZLE_128B applied), and one thread of each CTA
// walks the CTA's future units in program order and copies the next images into a ring with cp.async.bulk (TMA, completion
// on an mbarrier) whenever a stage is free -- across grid barriers, GroupNorm and attention ops.  HBM keeps streaming while
// the dependent chain synchronises, and a token GEMM finds its first pairs already resident.
struct WJob {
  const uint8_t* base;        // [tiles_n * ks][pairs_max][32 KB]
  int tiles_n, tiles_m, ks, n_chunks, pairs_max, pad;
};

__device__ __forceinline__ void named_bar(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }

__device__ __forceinline__ void red_release_add(unsigned* p, unsigned v) {
  asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned ld_acquire(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

__device__ __forceinline__ void st_release(unsigned* p, unsigned v) {
  asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// Grid barrier over all CTAs of the (co-resident) grid, split in two so that work which does not depend on the other
// CTAs (the next op's weight copies) can be issued between arriving and waiting.  One release-increment and an
// acquire-poll of one counter by thread 0; the CTA barriers around them order the other threads.  (Measured on B200:
// 1.85 us per barrier at 148 CTAs; per-CTA flag words polled by a warp, and a two-level counter tree, were both slower.)
__device__ __forceinline__ void grid_arrive(unsigned* sync) {   // call after a __syncthreads()
  if (threadIdx.x == 0) red_release_add(sync, 1u);
}
// returns false when the run was aborted (a CTA waited > ~1 s: never expected)
template <typename Idle>
__device__ __forceinline__ bool grid_wait(unsigned* sync, unsigned& target, unsigned G, int* s_ok, Idle idle) {
  if (threadIdx.x == 0) {
    target += G;
    int ok = 1;
    const long long t0 = clock64();
    unsigned spins = 0;
    for (;;) {
      const unsigned v = ld_acquire(sync);
      if ((int)(v - target) >= 0) break;
      idle();   // the weight stream keeps issuing while this CTA waits
      if ((++spins & 1023u) == 0u) {
        if (*(volatile unsigned*)(sync + 1) != 0u || clock64() - t0 > 2000000000LL) {
          atomicExch(sync + 1, 1u);
          ok = 0;
          break;
        }
      }
    }
    *s_ok = ok;
  }
  __syncthreads();
  return *s_ok != 0;
}

// ---- token GEMM units ------------------------------------------------------------------------------------------
template <int MODE>
__device__ __forceinline__ void p_conv(const POp& o, float* smem, float* partials, unsigned* sems, int* s_last, int cta, int G, long long* prof) {
  const bool pf = prof != nullptr && cta == 0 && threadIdx.x == 0;   // diagnostics: phase split of the first CTA's units
  const ConvArgs& a = o.conv;
  const int ks = o.ks;
  const int n_units = o.tiles_n * o.tiles_m * ks;
  for (int u = cta; u < n_units; u += G) {
    const int crank = u % ks, tile = u / ks;
    const int tn = tile % o.tiles_n, tm = tile / o.tiles_n;
    const int n0 = tn * CT, m0 = tm * CT;
    float acc[8][4];
    long long t0 = 0, t1 = 0, t2 = 0, t3 = 0;
    if (pf) t0 = clock64();
    conv_accumulate<MODE>(a, n0, m0, ks, crank, smem, acc, pf ? prof : nullptr);
    if (pf) t1 = clock64();
    float* part = conv_cta_partial<MODE>(acc, smem);
    __syncthreads();
    if (pf) t2 = clock64();
    if (ks == 1) {
      float vv[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        float v = 0.f;
        v += part[(int)threadIdx.x + 256 * i];
        vv[i] = conv_epilogue_operands(a, m0, n0, (int)threadIdx.x + 256 * i, v);
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) conv_store_value(a, m0, n0, (int)threadIdx.x + 256 * i, vv[i]);
    } else {
      float* mine = partials + ((size_t)tile * ks + crank) * (CT * CT);
      for (int idx = threadIdx.x; idx < CT * CT; idx += 256) __stcg(mine + idx, part[idx]);
      __syncthreads();
      if (threadIdx.x == 0) {
        __threadfence();   // release: cumulative over the CTA's stores ordered before it by the barrier above
        const unsigned old = atomicAdd(sems + tile, 1u);
        const int last = old == (unsigned)(ks - 1);
        if (last) sems[tile] = 0u;   // every slice has arrived; the next use is at least one grid barrier away
        *s_last = last;
      }
      __syncthreads();
      if (pf) t3 = clock64();
      if (*s_last) {
        __threadfence();
        // all K-slice partials of this thread's 4 elements are fetched first (independent L2 loads in flight together),
        // then summed in slice order -- the order that makes the result independent of which CTA arrives last
        const float* base = partials + (size_t)tile * ks * (CT * CT) + threadIdx.x;
        float pv[4][P_MAX_KS];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < P_MAX_KS; ++j) pv[i][j] = j < ks ? __ldcg(base + (size_t)j * (CT * CT) + 256 * i) : 0.f;
        float vv[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          float v = 0.f;
#pragma unroll
          for (int j = 0; j < P_MAX_KS; ++j)
            if (j < ks) v += pv[i][j];
          vv[i] = conv_epilogue_operands(a, m0, n0, (int)threadIdx.x + 256 * i, v);
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) conv_store_value(a, m0, n0, (int)threadIdx.x + 256 * i, vv[i]);
      }
    }
    __syncthreads();   // the staging buffers are reused by the next unit
    if (pf) {
      const long long t4 = clock64();
      prof[0] += t1 - t0;                       // chunk stream (loads + MMA)
      prof[1] += t2 - t1;                       // warp partials -> CTA partial
      prof[2] += ks == 1 ? 0 : t3 - t2;         // partial tile to L2 + semaphore
      prof[24] += ks == 1 ? t4 - t2 : t4 - t3;  // slice reduction + epilogue (only when this CTA arrived last)
      prof[25] += 1;
    }
  }
}

// weight lines of this CTA's units of a token-GEMM op -> L2 (hint; issued before the barrier in front of that op)
__device__ __forceinline__ void p_conv_prefetch(const POp& o, int cta, int G) {
  const ConvArgs& a = o.conv;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ks = o.ks;
  const int n_units = o.tiles_n * o.tiles_m * ks;
  const int n_chunks0 = a.seg[0].taps * (a.seg[0].Cin / CT);
  const int n_chunks = n_chunks0 + (a.nseg > 1 ? a.seg[1].taps * (a.seg[1].Cin / CT) : 0);
  for (int u = cta; u < n_units; u += G) {
    const int crank = u % ks, tile = u / ks;
    const int n0 = (tile % o.tiles_n) * CT;
    for (int f = warp * ks + crank; f < n_chunks; f += 8 * ks) {
      const int s = f < n_chunks0 ? 0 : 1;
      const Seg& sg = a.seg[s];
      const int g = s ? f - n_chunks0 : f;
      const int cpt = sg.Cin / CT;
      const int tap = g / cpt, c = g - tap * cpt;
      const float* wrow = sg.W + ((size_t)tap * a.N + n0 + lane) * sg.Cin + c * CT;
      asm volatile("prefetch.global.L2 [%0];" ::"l"(wrow));
      asm volatile("prefetch.global.L2 [%0];" ::"l"(wrow + CT - 1));
    }
  }
}

// ---- wide token-GEMM units (default engine) --------------------------------------------------------------------
// The narrow unit above moves one 4 KB activation chunk per 4 KB weight chunk (32 x 32 tile): half of the L2 -> SM
// traffic of a weight-streaming op is activations, and every warp runs its own short, latency-exposed copy chain.
// The wide unit owns 32 tokens x 128 outputs of one contiguous K slice:
//   * CTA-wide cp.async ring of W_STAGES stages; a stage holds TWO K chunks (32 channels each): 2 x (32 x 32) token rows
//     + 2 x (128 x 32) weight rows = 40 KB of payload, three stages (120 KB) in flight per SM -- activations are 1/5
//     of the traffic and a slice of <= 8 chunks is requested in full before the first MMA;
//   * warp w multiplies n-subtile (w & 3) of chunk (w >> 2) of the stage; the 3xTF32 products of the 8 independent
//     accumulator fragments are issued term by term (lo*hi for all, hi*lo for all, hi*hi for all), so consecutive
//     mma.sync instructions never depend on each other;
//   * the two chunk groups meet in shared memory, the K slices of a tile meet through L2 partial tiles and a per-tile
//     arrival counter; once all slices have arrived EVERY slice's CTA reduces 1/ks of the tile in slice order (fixed
//     summation order: the result depends on the slice count but not on timing) and applies the epilogue.
// A unit's slices must be co-resident (they wait for each other): the host only splits K when tiles * ks <= grid.
constexpr int WN = 128;                                // outputs per wide tile
constexpr int W_STAGES = 4;
constexpr int WTP = 40;                                // padded row (floats): 8-byte fragment loads of rows g, g+1, .. hit disjoint banks
constexpr int W_STAGE = (2 * CT + 2 * WN) * WTP;       // floats per stage
constexpr int WIDE_SMEM = (W_STAGES * W_STAGE + 1536) * (int)sizeof(float);   // staging ring + fused-GroupNorm tables (PN_TAB)
constexpr int P_MAX_KS_WIDE = 32;

__device__ __forceinline__ void mma_tf32_nv(float* c, const uint32_t* a, const uint32_t* b) {
  asm("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

__device__ __forceinline__ void mma_f16(float* c, const uint32_t* a, const uint32_t* b) {
  asm("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
// x = hi + lo / 4096 with hi, lo in fp16 (11 significant bits each, the same split as 3xTF32's hi + lo; the residual is
// scaled by 2^12 so that it stays a normal fp16 number).  Packed pairs: (x.x, x.y) -> low / high half.
constexpr float F16_LO_SCALE = 4096.0f;
__device__ __forceinline__ void split_f16x2(float2 x, uint32_t& hi, uint32_t& lo) {
  asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(hi) : "f"(x.y), "f"(x.x));
  float hx, hy;
  asm("{ .reg .f16 l, h; mov.b32 {l, h}, %2; cvt.f32.f16 %0, l; cvt.f32.f16 %1, h; }" : "=f"(hx), "=f"(hy) : "r"(hi));
  const float rx = (x.x - hx) * F16_LO_SCALE, ry = (x.y - hy) * F16_LO_SCALE;
  asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(lo) : "f"(ry), "f"(rx));
}

// One warp: acc[mt*4+nt] += As[32 tok][32 k] * Ws[32 n][32 k]^T (rows padded to WTP floats).
// MODE 1: fp32-class product on the fp16 tensor path, mma.sync m16n8k16: acc += a_hi b_hi, acc2 += a_hi b_lo + a_lo b_hi
// (acc2 carries the 2^12 scale; the caller adds acc2 / 4096).  Half the tensor-pipe work of the 3xTF32 m16n8k8 split, which
// bounded the wide units (measured 1.55 us per chunk pair and SM).  Values must lie inside the fp16 range (|x| < 65504).
// MODE 2: single-pass TF32, m16n8k8.
template <int MODE>
__device__ __forceinline__ void mma_chunk(const float* __restrict__ As, const float* __restrict__ Ws, float (&acc)[8][4], float (&acc2)[8][4],
                                          int fg, int ft) {
  if (MODE == 1) {
#pragma unroll
    for (int k0 = 0; k0 < CT; k0 += 16) {
      // A fragment (16x16, row-major): a0 (g, 2t..) a1 (g+8, 2t..) a2 (g, 2t+8..) a3 (g+8, 2t+8..); B (16x8): b0 (k 2t.., n g) b1 (k 2t+8.., n g)
      uint32_t ahi[2][4], alo[2][4], bhi[4][2], blo[4][2];
#pragma unroll
      for (int mt = 0; mt < 2; ++mt) {
        const float* ap = As + (mt * 16 + fg) * WTP + k0 + 2 * ft;
        split_f16x2(*reinterpret_cast<const float2*>(ap), ahi[mt][0], alo[mt][0]);
        split_f16x2(*reinterpret_cast<const float2*>(ap + 8 * WTP), ahi[mt][1], alo[mt][1]);
        split_f16x2(*reinterpret_cast<const float2*>(ap + 8), ahi[mt][2], alo[mt][2]);
        split_f16x2(*reinterpret_cast<const float2*>(ap + 8 * WTP + 8), ahi[mt][3], alo[mt][3]);
      }
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) {
        const float* bp = Ws + (nt * 8 + fg) * WTP + k0 + 2 * ft;
        split_f16x2(*reinterpret_cast<const float2*>(bp), bhi[nt][0], blo[nt][0]);
        split_f16x2(*reinterpret_cast<const float2*>(bp + 8), bhi[nt][1], blo[nt][1]);
      }
#pragma unroll
      for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) mma_f16(acc2[mt * 4 + nt], alo[mt], bhi[nt]);
#pragma unroll
      for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) mma_f16(acc[mt * 4 + nt], ahi[mt], bhi[nt]);
#pragma unroll
      for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) mma_f16(acc2[mt * 4 + nt], ahi[mt], blo[nt]);
    }
  } else {
#pragma unroll
    for (int k0 = 0; k0 < CT; k0 += 8) {
      uint32_t ahi[2][4], bhi[4][2];
#pragma unroll
      for (int mt = 0; mt < 2; ++mt) {
        const float* ap = As + (mt * 16 + fg) * WTP + k0 + ft;
        ahi[mt][0] = to_tf32(ap[0]); ahi[mt][1] = to_tf32(ap[8 * WTP]); ahi[mt][2] = to_tf32(ap[4]); ahi[mt][3] = to_tf32(ap[8 * WTP + 4]);
      }
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) {
        const float* bp = Ws + (nt * 8 + fg) * WTP + k0 + ft;
        bhi[nt][0] = to_tf32(bp[0]); bhi[nt][1] = to_tf32(bp[4]);
      }
#pragma unroll
      for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) mma_tf32_nv(acc[mt * 4 + nt], ahi[mt], bhi[nt]);
    }
  }
}

// cp.async copies of chunk pair `pair` of a wide unit into stage `stg`: what & 1 = the 2 x 128 weight rows, what & 2 = the
// 2 x 32 token rows.  Copy mapping: thread -> 16-byte piece tid & 7 of row tid >> 3 (8 consecutive lanes = one 128 B line).
template <int PITCH>
__device__ __forceinline__ void wide_issue(const ConvArgs& a, int n_chunks0, int f_begin, int f_end, int n0, int m0, int pair, float* stg, int what) {
  const int tid = (int)threadIdx.x;
  const int lp = tid & 7, cr = tid >> 3;
  const int m_row = m0 + cr;
  const bool rok = m_row < a.B * a.T_out;
  const int rb = m_row / a.T_out, rl = m_row - rb * a.T_out;
#pragma unroll
  for (int g = 0; g < 2; ++g) {
    const int f = f_begin + 2 * pair + g;
    if (f < f_end) {
      const int s = f < n_chunks0 ? 0 : 1;
      const Seg& sg = a.seg[s];
      const int q = s ? f - n_chunks0 : f;
      const int cb = q / sg.taps, tap = q - cb * sg.taps;   // block-major: the taps of one 32-channel block are adjacent chunks
      if (what & 2) {
        const int T_eff = sg.up ? 2 * sg.T_in : sg.T_in;
        const int src = rl * sg.stride + tap - (sg.taps >> 1);
        const bool ok = rok && src >= 0 && src < T_eff;
        const int st = sg.up ? (src >> 1) : src;
        const bool second = cb * CT >= sg.C1;
        const float* base = second ? sg.A2 : sg.A;
        const int ld = second ? sg.Cin - sg.C1 : sg.C1;
        const int coff = second ? cb * CT - sg.C1 : cb * CT;
        const float* arow = base + (ok ? ((size_t)rb * sg.T_in + st) * ld : (size_t)0) + coff + 4 * lp;
        const uint32_t da = (uint32_t)__cvta_generic_to_shared(stg + (g * CT + cr) * PITCH + 4 * lp);
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(da), "l"(arow), "r"(ok ? 16 : 0) : "memory");
      }
      if (what & 1) {
        const float* wbase = sg.W + (size_t)tap * a.N * sg.Cin + cb * CT + 4 * lp;
        const uint32_t dw = (uint32_t)__cvta_generic_to_shared(stg + (2 * CT + g * WN + cr) * PITCH + 4 * lp);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int n = n0 + cr + 32 * j;
          const bool okw = n < a.N;
          asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dw + j * (32 * PITCH * 4)), "l"(wbase + (size_t)(okw ? n : 0) * sg.Cin),
                       "r"(okw ? 16 : 0) : "memory");
        }
      }
    }
  }
}

// x / d for 0 <= x < 2^20 with inv = 1.0f / d (reciprocal multiply + one correction step)
__device__ __forceinline__ int div_small(int x, int d, float inv) {
  int q = (int)((float)x * inv);
  const int r = x - q * d;
  if (r < 0) --q;
  else if (r >= d) ++q;
  return q;
}

// Fused GroupNorm, part 1: statistics of the (sample, group) pairs touched by this unit's segment-0 chunks, and the
// scale / shift of its channel range, into the tables behind the staging ring.  One warp per pair; a pair's T * cg values
// are padded to whole 32-lane slots and the slots of up to 32 / slots-per-pair pairs are requested together (one L2 round
// trip per round; two-pass mean / variance in registers like gn_kernel).  tab: gam[512] bet[512] mu[256] rs[256].
constexpr int PN_TAB = 1536;
struct PnRange { int c_lo, g_lo, ng; };
__device__ __forceinline__ PnRange wide_prenorm_stats(const ConvArgs& a, int n_chunks0, int f_begin, int f_end, int m0, float* tab) {
  PnRange R{0, 0, 0};
  const int fe0 = f_end < n_chunks0 ? f_end : n_chunks0;
  if (f_begin >= fe0) return R;
  const Seg& sg = a.seg[0];
  const int tid = (int)threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int cb_lo = f_begin / sg.taps, cb_hi = (fe0 - 1) / sg.taps;
  const int c_lo = cb_lo * CT, c_hi = (cb_hi + 1) * CT;
  const int cg = a.pn_cg, T = 1 << a.pn_logT;
  const int g_lo = c_lo / cg, g_hi = (c_hi - 1) / cg;
  const int ng = g_hi - g_lo + 1;
  R.c_lo = c_lo; R.g_lo = g_lo; R.ng = ng;
  float* gam = tab; float* bet = tab + 512; float* mu = tab + 1024; float* rs = tab + 1280;
  for (int c = c_lo + tid; c < c_hi; c += 256) { gam[c - c_lo] = a.pn_gamma[c]; bet[c - c_lo] = a.pn_beta[c]; }
  const int M = a.B * a.T_out;
  const int rows = M - m0 < CT ? M - m0 : CT;
  const int nb = rows >> a.pn_logT, b0 = m0 >> a.pn_logT;
  const int P = nb * ng;
  const int n = T * cg;
  const int sp = (n + 31) >> 5;          // slots per pair
  const int kpr = 32 / sp;               // pairs per round (sp <= 21 in this network; the host checks sp <= 32)
  const int C1 = sg.C1, C2 = sg.Cin - sg.C1;
  const int pw = P > warp ? (P - warp + 7) >> 3 : 0;   // pairs of this warp: warp, warp + 8, ...
  for (int k0 = 0; k0 < pw; k0 += kpr) {
    float xv[32];
#pragma unroll
    for (int q = 0; q < 32; ++q) {
      xv[q] = 0.f;
      const int kk = q / sp, k = k0 + kk;
      const int e = (q - kk * sp) * 32 + lane;
      if (kk < kpr && k < pw && e < n) {
        const int pr = warp + 8 * k;
        const int bl = pr / ng, gi = pr - bl * ng;
        const int t = div_small(e, cg, a.pn_inv_cg), cc = e - t * cg;
        const int c = (g_lo + gi) * cg + cc;
        const size_t m = (size_t)(b0 + bl) * T + t;
        xv[q] = c < C1 ? __ldcg(sg.A + m * C1 + c) : __ldcg(sg.A2 + m * C2 + (c - C1));
      }
    }
    for (int kk = 0; kk < kpr && k0 + kk < pw; ++kk) {
      float sum = 0.f;
#pragma unroll
      for (int q = 0; q < 32; ++q)
        if (q / sp == kk) sum += xv[q];
#pragma unroll
      for (int of = 16; of > 0; of >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, of);
      const float mean = sum / (float)n;
      float sq = 0.f;
#pragma unroll
      for (int q = 0; q < 32; ++q)
        if (q / sp == kk && (q - kk * sp) * 32 + lane < n) { const float d = xv[q] - mean; sq += d * d; }
#pragma unroll
      for (int of = 16; of > 0; of >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, of);
      if (lane == 0) {
        const int pr = warp + 8 * (k0 + kk);
        mu[pr] = mean;
        rs[pr] = 1.0f / sqrtf(sq / (float)n + 1e-5f);
      }
    }
  }
  return R;
}

// Fused GroupNorm, part 2: the landed token rows of a stage's segment-0 chunks are normalised (+ SiLU) in place.  Rows that
// the copy zero-filled (conv padding, rows past the batch) stay zero.  Thread -> column tid & 31, rows (tid >> 5) + 8 i.
__device__ __forceinline__ void wide_prenorm_apply(const ConvArgs& a, int n_chunks0, int f_begin, int f_end, int m0, int pair, float* stg,
                                                   const float* tab, const PnRange& R) {
  const Seg& sg = a.seg[0];
  const int tid = (int)threadIdx.x, j = tid & 31, r0 = tid >> 5;
  const int T = 1 << a.pn_logT;
  const int M = a.B * a.T_out;
  const float* gam = tab; const float* bet = tab + 512; const float* mu = tab + 1024; const float* rs = tab + 1280;
#pragma unroll
  for (int g = 0; g < 2; ++g) {
    const int f = f_begin + 2 * pair + g;
    if (f < f_end && f < n_chunks0) {
      const int cb = f / sg.taps, tap = f - cb * sg.taps;
      const int c = cb * CT + j;
      const int gi = div_small(c, a.pn_cg, a.pn_inv_cg) - R.g_lo;
      const float gm = gam[c - R.c_lo], bt = bet[c - R.c_lo];
      float* base = stg + g * CT * WTP + j;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int r = r0 + 8 * i;
        const int src = (r & (T - 1)) + tap - (sg.taps >> 1);
        if (m0 + r < M && src >= 0 && src < T) {
          const int pr = (r >> a.pn_logT) * R.ng + gi;
          float y = (base[r * WTP] - mu[pr]) * rs[pr] * gm + bt;
          if (a.pn_silu) y = y / (1.0f + expf(-y));
          base[r * WTP] = y;
        }
      }
    }
  }
}

// weight rows of the first W_STAGES - 1 chunk pairs of this CTA's first unit of a wide token GEMM, requested while the
// CTA waits at the grid barrier in front of that op (weights do not depend on the other CTAs).  Not committed: the
// copies join the first commit group of the unit's own prologue.
template <int PITCH, int STAGE_FLOATS, int N_STAGES>
__device__ __forceinline__ void wide_preissue(const POp& o, float* smem, int cta) {
  const ConvArgs& a = o.conv;
  const int ks = o.ks;
  if (cta >= o.tiles_n * o.tiles_m * ks) return;
  const int crank = cta % ks, tile = cta / ks;
  const int n0 = (tile % o.tiles_n) * WN, m0 = (tile / o.tiles_n) * CT;
  const int n_chunks0 = a.seg[0].taps * (a.seg[0].Cin / CT);
  const int n_chunks = n_chunks0 + (a.nseg > 1 ? a.seg[1].taps * (a.seg[1].Cin / CT) : 0);
  const int f_begin = (crank * n_chunks) / ks, f_end = ((crank + 1) * n_chunks) / ks;
  const int n_pairs = (f_end - f_begin + 1) >> 1;
#pragma unroll
  for (int p = 0; p < N_STAGES - 1; ++p)
    if (p < n_pairs) wide_issue<PITCH>(a, n_chunks0, f_begin, f_end, n0, m0, p, smem + p * STAGE_FLOATS, 1);
}

// spin until *p >= want (another CTA's release); false on abort / timeout (never expected)
__device__ __forceinline__ bool spin_until(const unsigned* p, unsigned want, unsigned* abort_flag) {
  const long long t0 = clock64();
  unsigned spins = 0;
  for (;;) {
    if (ld_acquire(p) >= want) return true;
    if ((++spins & 1023u) == 0u) {
      if (*(volatile unsigned*)abort_flag != 0u || clock64() - t0 > 2000000000LL) {
        atomicExch(abort_flag, 1u);
        return false;
      }
    }
  }
}

template <int MODE>
__device__ __forceinline__ void p_conv_wide(const POp& o, float* smem, float* partials, unsigned* sems, unsigned* abort_flag, bool pre_issued,
                                            int cta, int G, long long* prof) {
  const bool pf = prof != nullptr && cta == 0 && threadIdx.x == 0;
  const ConvArgs& a = o.conv;
  const int ks = o.ks;
  const int n_units = o.tiles_n * o.tiles_m * ks;
  const int tid = (int)threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int nsub = warp & 3, kg = warp >> 2;
  const int fg = lane >> 2, ft = lane & 3;
  const int M = a.B * a.T_out;
  const int n_chunks0 = a.seg[0].taps * (a.seg[0].Cin / CT);
  const int n_chunks = n_chunks0 + (a.nseg > 1 ? a.seg[1].taps * (a.seg[1].Cin / CT) : 0);
  for (int u = cta; u < n_units; u += G) {
    long long t0 = 0, t1 = 0, t2 = 0, t3 = 0;
    if (pf) t0 = clock64();
    // unit -> (tile, slice) -> chunk range: integer divisions by run-time values cost a lone warp ~35 dependent instructions
    // each (seven of them were 0.7 us at the head of every unit); the divisors' reciprocals come with the descriptor
    const int tile = div_small(u, ks, o.inv_ks), crank = u - tile * ks;
    const int tm = div_small(tile, o.tiles_n, o.inv_tn), tn = tile - tm * o.tiles_n;
    const int n0 = tn * WN, m0 = tm * CT;
    const int f_begin = div_small(crank * n_chunks, ks, o.inv_ks), f_end = div_small((crank + 1) * n_chunks, ks, o.inv_ks);
    const int n_pairs = (f_end - f_begin + 1) >> 1;

    auto issue_pair = [&](int pair, float* stg, int what) {
      wide_issue<WTP>(a, n_chunks0, f_begin, f_end, n0, m0, pair, stg, what);
      asm volatile("cp.async.commit_group;" ::: "memory");
    };

    float acc[8][4], acc2[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) { acc[i][j] = 0.f; acc2[i][j] = 0.f; }
    const int first_what = (pre_issued && u == cta) ? 2 : 3;   // this CTA's first unit: the weight rows are already in flight
#pragma unroll
    for (int p = 0; p < W_STAGES - 1; ++p) {
      if (p < n_pairs) issue_pair(p, smem + p * W_STAGE, first_what);
      else asm volatile("cp.async.commit_group;" ::: "memory");
    }
    float* tab = smem + W_STAGES * W_STAGE;
    PnRange pn{0, 0, 0};
    if (a.pn_on) pn = wide_prenorm_stats(a, n_chunks0, f_begin, f_end, m0, tab);   // overlaps the copies in flight
    int stg = 0;
    for (int p = 0; p < n_pairs; ++p) {
      asm volatile("cp.async.wait_group %0;" ::"n"(W_STAGES - 2) : "memory");
      __syncthreads();   // pair p has landed for every thread; everybody is done with pair p-1 (its stage is refilled below)
      if (pf && p == 0) { prof[27] += clock64() - t0; prof[26] += n_pairs; }   // time to the first pair, pairs of this unit
      const int nxt = p + W_STAGES - 1;
      if (nxt < n_pairs) issue_pair(nxt, smem + (stg == 0 ? W_STAGES - 1 : stg - 1) * W_STAGE, 3);
      else asm volatile("cp.async.commit_group;" ::: "memory");
      if (a.pn_on && f_begin + 2 * p < n_chunks0) {
        wide_prenorm_apply(a, n_chunks0, f_begin, f_end, m0, p, smem + stg * W_STAGE, tab, pn);
        __syncthreads();
      }
      if (f_begin + 2 * p + kg < f_end) {
        const float* As = smem + stg * W_STAGE + kg * CT * WTP;
        const float* Ws = smem + stg * W_STAGE + (2 * CT + kg * WN + nsub * 32) * WTP;
        mma_chunk<MODE>(As, Ws, acc, acc2, fg, ft);
      }
      stg = stg + 1 == W_STAGES ? 0 : stg + 1;
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();
    if (MODE == 1) {
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] += acc2[i][j] * (1.0f / F16_LO_SCALE);
    }
    if (pf) t1 = clock64();
    // the two chunk groups meet: group 1 parks its fragments (lane-major, conflict-free), group 0 adds them
    float* red = smem;
    if (kg == 1) {
#pragma unroll
      for (int i = 0; i < 32; ++i) red[(nsub * 32 + i) * 32 + lane] = acc[i >> 2][i & 3];
    }
    __syncthreads();
    if (kg == 0) {
#pragma unroll
      for (int i = 0; i < 32; ++i) acc[i >> 2][i & 3] += red[(nsub * 32 + i) * 32 + lane];
    }
    if (pf) t2 = clock64();
    // accumulator fragment: c0 (g, 2t) c1 (g, 2t+1) c2 (g+8, 2t) c3 (g+8, 2t+1)
    if (ks == 1) {
      if (kg == 0) {
#pragma unroll
        for (int mt = 0; mt < 2; ++mt)
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const int m = m0 + mt * 16 + fg + h * 8;
            if (m < M) {
              const int b = m / a.T_out;
#pragma unroll
              for (int nt = 0; nt < 4; ++nt) {
                const int n = n0 + nsub * 32 + nt * 8 + 2 * ft;
                if (n < a.N) {
                  float2 v = make_float2(acc[mt * 4 + nt][2 * h], acc[mt * 4 + nt][2 * h + 1]);
                  const float2 bv = *reinterpret_cast<const float2*>(a.bias + n);
                  v.x += bv.x; v.y += bv.y;
                  if (a.emb) { const float2 e = *reinterpret_cast<const float2*>(a.emb + (size_t)b * a.emb_ld + n); v.x += e.x; v.y += e.y; }
                  if (a.residual) { const float2 r = *reinterpret_cast<const float2*>(a.residual + (size_t)m * a.N + n); v.x += r.x; v.y += r.y; }
                  *reinterpret_cast<float2*>(a.out + (size_t)m * a.N + n) = v;
                }
              }
            }
          }
      }
      if (pf) t3 = clock64();
    } else {
      float* mine = partials + ((size_t)tile * ks + crank) * (CT * WN);
      if (o.defer) {   // the consumer sums the slices: publish and go straight to the grid barrier
        if (kg == 0) {
#pragma unroll
          for (int mt = 0; mt < 2; ++mt)
#pragma unroll
            for (int h = 0; h < 2; ++h)
#pragma unroll
              for (int nt = 0; nt < 4; ++nt)
                __stcg(reinterpret_cast<float2*>(mine + (mt * 16 + fg + h * 8) * WN + nsub * 32 + nt * 8 + 2 * ft),
                       make_float2(acc[mt * 4 + nt][2 * h], acc[mt * 4 + nt][2 * h + 1]));
        }
        __syncthreads();
        if (pf) { const long long t4 = clock64(); prof[0] += t1 - t0; prof[1] += t2 - t1; prof[2] += t4 - t2; prof[25] += 1; }
        continue;
      }
      if (kg == 0) {
#pragma unroll
        for (int mt = 0; mt < 2; ++mt)
#pragma unroll
          for (int h = 0; h < 2; ++h)
#pragma unroll
            for (int nt = 0; nt < 4; ++nt)
              __stcg(reinterpret_cast<float2*>(mine + (mt * 16 + fg + h * 8) * WN + nsub * 32 + nt * 8 + 2 * ft),
                     make_float2(acc[mt * 4 + nt][2 * h], acc[mt * 4 + nt][2 * h + 1]));
      }
      __syncthreads();
      if (tid == 0) {
        red_release_add(sems + tile, 1u);   // releases the CTA's partial tile (ordered before it by the barrier above)
        spin_until(sems + tile, (unsigned)ks, abort_flag);   // acquires the other slices' tiles
      }
      __syncthreads();
      if (pf) t3 = clock64();
      // this slice's share of the tile: elements [e0, e1), summed over the slices in slice order
      const int e0 = (crank * (CT * WN)) / ks, e1 = ((crank + 1) * (CT * WN)) / ks;
      const float* base = partials + (size_t)tile * ks * (CT * WN);
      for (int idx = e0 + tid; idx < e1; idx += 256) {
        // every slice's value and the epilogue operands are requested together (one L2 round trip), then summed in slice order
        const int m = m0 + (idx >> 7), n = n0 + (idx & (WN - 1));
        const bool live = m < M && n < a.N;
        float pv[P_MAX_KS_WIDE];
#pragma unroll
        for (int j = 0; j < P_MAX_KS_WIDE; ++j) pv[j] = j < ks ? __ldcg(base + (size_t)j * (CT * WN) + idx) : 0.f;
        float bv = 0.f, ev = 0.f, rv = 0.f;
        if (live) {
          bv = a.bias[n];
          if (a.emb) ev = a.emb[(size_t)(m / a.T_out) * a.emb_ld + n];
          if (a.residual) rv = a.residual[(size_t)m * a.N + n];
        }
        float v = 0.f;
#pragma unroll
        for (int j = 0; j < P_MAX_KS_WIDE; ++j)
          if (j < ks) v += pv[j];
        if (live) {
          v += bv;
          if (a.emb) v += ev;
          if (a.residual) v += rv;
          a.out[(size_t)m * a.N + n] = v;
        }
      }
    }
    __syncthreads();   // the staging buffers are reused by the next unit
    if (pf) {
      const long long t4 = clock64();
      prof[0] += t1 - t0;      // chunk stream (copies + MMA)
      prof[1] += t2 - t1;      // chunk groups -> CTA partial
      prof[2] += t3 - t2;      // publish + wait for the other slices (or the direct epilogue when ks == 1)
      prof[24] += t4 - t3;     // distributed slice reduction + epilogue
      prof[25] += 1;
    }
  }
}

// value of element (m, c) of a deferred token-GEMM output: slices summed in slice order, then + bias (+ emb) (+ residual) --
// the same order of additions as the in-op epilogue of p_conv_wide.  All loads are requested together.
// b = m / T when the caller knows it (the division by a run-time T costs ~35 dependent instructions per element), else -1
__device__ __forceinline__ float part_value(const PartSrc& ps, int T, int m, int c, int b = -1) {
  const float* base = ps.part + ((size_t)((m >> 5) * ps.tiles_n + (c >> 7)) * ps.ks) * (CT * WN) + (m & 31) * WN + (c & (WN - 1));
  float pv[P_MAX_KS_WIDE];
#pragma unroll
  for (int j0 = 0; j0 < P_MAX_KS_WIDE; j0 += 8) {   // whole groups of 8 slices are skipped (uniform branch)
    if (j0 < ps.ks) {
#pragma unroll
      for (int j = j0; j < j0 + 8; ++j) pv[j] = j < ps.ks ? __ldcg(base + (size_t)j * (CT * WN)) : 0.f;
    } else {
#pragma unroll
      for (int j = j0; j < j0 + 8; ++j) pv[j] = 0.f;
    }
  }
  const float bv = ps.bias[c];
  const float ev = ps.emb ? ps.emb[(size_t)(b >= 0 ? b : m / T) * ps.emb_ld + c] : 0.f;
  const float rv = ps.res ? ps.res[(size_t)m * ps.N + c] : 0.f;
  float v = 0.f;
#pragma unroll
  for (int j0 = 0; j0 < P_MAX_KS_WIDE; j0 += 8) {
    if (j0 < ps.ks) {
#pragma unroll
      for (int j = j0; j < j0 + 8; ++j)
        if (j < ps.ks) v += pv[j];
    }
  }
  v += bv;
  if (ps.emb) v += ev;
  if (ps.res) v += rv;
  return v;
}

// ---- wide token-GEMM units on the 5th-generation tensor cores (tcgen05 + TMEM), precision mode 1 --------------------------
// Same unit, slicing, exchange and summation orders as p_conv_wide; the product moves from mma.sync to tcgen05.mma.
// The fp16 two-term split (x = hi + lo / 4096) of the WEIGHTS is done once at load time: every 32-float block of a weight row
// is stored as [32 x hi | 32 x lo] halves in the same 128 bytes (`Seg::Wh`, same offsets as the fp32 blob), so the per-step
// weight stream has the same size and 16-byte cp.async copies drop the planes straight into K-major SWIZZLE_128B operand
// tiles -- no conversion pass over the weights (the F2F conversions bounded both the mma.sync and a first tcgen05 variant:
// 1.06 / 1.97 us per chunk pair).  Only the token rows (32 x 64 values per stage) are split on the fly.
//   stage (40 KB, 4-deep ring) = one K = 64 step: W_hi, W_lo [128 rows x 128 B], A_hi, A_lo [32 rows x 128 B];
//   one thread issues 4 k16-steps x 3 tcgen05.mma.kind::f16 (M = 128 outputs, N = 32 tokens): D1 += W_hi A_hi,
//   D2 += W_lo A_hi + W_hi A_lo (32 TMEM columns each); tcgen05.commit -> mbarrier frees the stage for the copies of the
//   pair three ahead; warps 0-3 read the accumulators (tcgen05.ld 32x32b: thread = output channel, 32 token values).
constexpr int TC_STAGES = 4;
constexpr int TC_OP_BYTES = (WN + CT) * 128 * 2;               // hi + lo planes of W (128 rows) and A (32 rows), 128-byte rows
constexpr int TC_W_LO = WN * 128, TC_A_HI = 2 * WN * 128, TC_A_LO = 2 * WN * 128 + CT * 128;
constexpr int TC_ASTG = 2 * CT * CT;                            // floats per fp32 token-row stage (2 chunks x 32 rows x 32)
constexpr int TC_SMEM = 1024 + TC_STAGES * TC_OP_BYTES + TC_STAGES * TC_ASTG * 4;
// weight-stream layout (use_tma): [W ring: TC_STAGES x 32 KB, written by cp.async.bulk only][A operand tiles: TC_STAGES x 8 KB]
// [fp32 token-row staging: TC_STAGES x 8 KB -- also the scratch of the attention / embedding ops, which run between GEMMs]
constexpr int TCX_W_BYTES = 2 * WN * 128;                       // one chunk pair of weights: hi + lo planes
constexpr int TCX_A_BYTES = 2 * CT * 128;
constexpr int TCX_A_OFF = TC_STAGES * TCX_W_BYTES;
constexpr int TCX_STG_OFF = TCX_A_OFF + TC_STAGES * TCX_A_BYTES;
static_assert(TCX_STG_OFF + TC_STAGES * TC_ASTG * 4 + 1024 == TC_SMEM, "both layouts use the same dynamic shared memory");
__device__ __forceinline__ uint8_t* tc_ops(float* smem) {   // operand ring: first 1024-byte boundary of the dynamic shared memory
  // (pointer arithmetic on the __shared__ array, not an integer cast: keeps the shared address space -> LDS / STS, not generic LD / ST)
  return reinterpret_cast<uint8_t*>(smem) + ((1024u - ((uint32_t)__cvta_generic_to_shared(smem) & 1023u)) & 1023u);
}
__device__ __forceinline__ float* tc_astage(float* smem) { return reinterpret_cast<float*>(tc_ops(smem) + TC_STAGES * TC_OP_BYTES); }

__device__ __forceinline__ uint32_t tcw_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void tcw_mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(tcw_smem_u32(bar)), "r"(count));
}
// bounded wait: false after ~1 s (never expected; the caller aborts the run)
__device__ __forceinline__ bool tcw_mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t done = 0, spins = 0;
  while (!done) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(tcw_smem_u32(bar)), "r"(parity)
        : "memory");
    if (!done && ++spins > (1u << 24)) return false;
  }
  return true;
}
// K-major SWIZZLE_128B operand tile: 128-byte rows, 8-row atoms of 1024 B (same encoding as decoder_tc.cu)
__device__ __forceinline__ uint64_t tcw_desc(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
// kind::f16 instruction descriptor: D = F32 (bit 4), A = B = F16 (format 0), both K-major, N = 32 tokens, M = 128 outputs
__device__ __forceinline__ uint32_t tcw_idesc() { return (1u << 4) | ((uint32_t)(CT >> 3) << 17) | ((uint32_t)(WN >> 4) << 24); }
__device__ __forceinline__ void tcw_mma(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tcw_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(tcw_smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tcw_ld32(uint32_t taddr, uint32_t* r) {
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

// What the token-row copies of one K chunk need, computed ONCE per chunk (by one thread) instead of by all 256: the chunk ->
// (segment, channel block, tap) decoding is the same for every row of the tile.
struct ChunkDesc {
  const float* base;   // source tensor (first or second concat source), already offset to the chunk's channels
  int ld;              // its row stride in floats
  int shift;           // tap - taps / 2: token offset of this tap
  int stride, up, T_in;
  int valid;           // 0: the pair's second chunk does not exist (odd chunk count)
};
// per-CTA state of the tcgen05 path (static shared memory of the persistent kernel)
struct TcState {
  uint64_t stage_free[TC_STAGES];   // the MMAs that read operand stage s have completed
  uint64_t full[TC_STAGES];         // weight stream: the image of ring stage s has landed (complete_tx)
  uint64_t done;                    // all MMAs of the unit have completed
  uint32_t tmem_base;
  uint32_t pad;
  ChunkDesc cd[2 * 4];              // token-row descriptors of the current batch of chunk pairs (2 per pair)
};
// thread-local, uniform: commits issued so far per barrier (a wait targets the phase of the latest commit; waiting twice for
// the same phase is harmless, so no wait is ever "owed")
struct TcPhase { uint32_t stage[TC_STAGES]; uint32_t done; uint32_t pairs; };   // pairs: chunk pairs consumed so far (weight stream: ring position)

// ---- weight-stream producer (thread 0 of every CTA) ----
__device__ __forceinline__ bool tcw_mbar_test(uint64_t* bar, uint32_t parity) {
  uint32_t done;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(done)
      : "r"(tcw_smem_u32(bar)), "r"(parity)
      : "memory");
  return done != 0;
}
// The CTA's images of one pass over the op list, in consumption order, are listed at load time (persist_build): the producer
// only walks a pointer list -- a single thread runs this code, and every dependent scalar instruction costs it 5-8 cycles
// (a first version that derived unit / slice / pair from the op table with integer divisions spent 1.8 us per unit here).
struct WCursor {
  const uint8_t* const* list;   // this CTA's image pointers of one pass
  uint32_t n, idx, reps;        // images per pass, position, passes still to stream (this one included)
  uint32_t issued;              // images issued so far
};
// issue as many images as the ring takes (never blocks).  The pointers of the next TC_STAGES images are requested together
// up front: one L2 round trip per call instead of one per image.
__device__ __forceinline__ void wstream_pump(WCursor& c, TcState* ts, uint8_t* ring) {
  if (c.reps == 0) return;
  {   // anything to do at all?  (the common case in the grid-barrier spin loop: the ring is full)
    const uint32_t s = c.issued & (TC_STAGES - 1), k = c.issued / TC_STAGES;
    if (k > 0 && !tcw_mbar_test(&ts->stage_free[s], (k - 1u) & 1u)) return;
  }
  uint32_t idx = c.idx, reps = c.reps, issued = c.issued;
  const uint8_t* ptr[TC_STAGES];
  {
    uint32_t j = idx;
#pragma unroll
    for (int i = 0; i < TC_STAGES; ++i) {
      ptr[i] = c.list[j];
      if (++j == c.n) j = 0;
    }
  }
#pragma unroll
  for (int i = 0; i < TC_STAGES; ++i) {
    if (reps == 0) break;
    const uint32_t s = issued & (TC_STAGES - 1), k = issued / TC_STAGES;
    // use k of stage s overwrites use k-1: its MMAs must have completed (commit -> stage_free phase k-1)
    if (k > 0 && !tcw_mbar_test(&ts->stage_free[s], (k - 1u) & 1u)) break;
    const uint32_t bar = tcw_smem_u32(&ts->full[s]);
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"((uint32_t)TCX_W_BYTES) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(tcw_smem_u32(ring + s * TCX_W_BYTES)), "l"(ptr[i]), "r"((uint32_t)TCX_W_BYTES), "r"(bar) : "memory");
    ++issued;
    if (++idx == c.n) { idx = 0; --reps; }
  }
  c.idx = idx; c.reps = reps; c.issued = issued;
}

// load-time split of a weight tensor: block of 32 floats -> [32 x hi | 32 x lo] halves in the same 128 bytes
__global__ void split_weights_kernel(const float* __restrict__ W, float* __restrict__ Wh, size_t n_blocks) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;   // one thread per pair of floats
  if (i >= n_blocks * 16) return;
  const size_t blk = i >> 4;
  const int j = (int)(i & 15);
  const float2 x = *reinterpret_cast<const float2*>(W + blk * 32 + 2 * j);
  uint32_t hi, lo;
  split_f16x2(x, hi, lo);
  uint32_t* dst = reinterpret_cast<uint32_t*>(Wh + blk * 32);
  dst[j] = hi;
  dst[16 + j] = lo;
}

// load-time packing of the weight stream: block (weight unit = tile_n * ks + slice, pair) writes one 32 KB image, same
// thread -> 16-byte piece mapping and the same swizzle as tc_issue()'s shared-memory destination
__global__ void __launch_bounds__(256) pack_wstream_kernel(ConvArgs a, int ks, int n_chunks0, int n_chunks, int pairs_max, uint8_t* dst) {
  const int wu = (int)blockIdx.x, pair = (int)blockIdx.y;
  const int crank = wu % ks, tn = wu / ks;
  const int n0 = tn * WN;
  const int f_begin = (crank * n_chunks) / ks, f_end = ((crank + 1) * n_chunks) / ks;
  uint8_t* img = dst + ((size_t)wu * pairs_max + pair) * TCX_W_BYTES;
  const int tid = (int)threadIdx.x, lp = tid & 7, cr = tid >> 3;
#pragma unroll
  for (int g = 0; g < 2; ++g) {
    const int f = f_begin + 2 * pair + g;
    const bool have = f < f_end;
    const int s = (have && f >= n_chunks0) ? 1 : 0;
    const Seg& sg = a.seg[s];
    const int q = have ? (s ? f - n_chunks0 : f) : 0;
    const int cb = q / sg.taps, tap = q - cb * sg.taps;
    const float* wbase = sg.Wh + (size_t)tap * a.N * sg.Cin + cb * CT + 4 * lp;
    const int plane = lp >> 2, p16 = 4 * g + (lp & 3);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int rr = cr + 32 * j;
      const int n = n0 + rr;
      uint4 v = make_uint4(0u, 0u, 0u, 0u);
      if (have && n < a.N) v = *reinterpret_cast<const uint4*>(wbase + (size_t)n * sg.Cin);
      *reinterpret_cast<uint4*>(img + (plane ? TC_W_LO : 0) + rr * 128 + ((p16 ^ (rr & 7)) << 4)) = v;
    }
  }
}

// cp.async copies of chunk pair `pair` of a tcgen05 unit: what & 1 = split weight rows straight into the swizzled operand
// tiles of `ob`; what & 2 = fp32 token rows into `astg`.  Weight mapping: thread -> 16-byte piece tid & 7 of the 128-byte
// block (pieces 0-3 hi, 4-7 lo) of row (tid >> 3) + 32 j: 8 consecutive lanes read one full line.
__device__ __forceinline__ void tc_issue(const ConvArgs& a, int n_chunks0, int f_begin, int f_end, int n0, int m0, int pair, uint8_t* ob,
                                         float* astg, int what) {
  const int tid = (int)threadIdx.x;
  const int lp = tid & 7, cr = tid >> 3;
  const int m_row = m0 + cr;
  const bool rok = m_row < a.B * a.T_out;
  const int rb = m_row / a.T_out, rl = m_row - rb * a.T_out;
#pragma unroll
  for (int g = 0; g < 2; ++g) {
    const int f = f_begin + 2 * pair + g;
    if (f < f_end) {
      const int s = f < n_chunks0 ? 0 : 1;
      const Seg& sg = a.seg[s];
      const int q = s ? f - n_chunks0 : f;
      const int cb = q / sg.taps, tap = q - cb * sg.taps;
      if (what & 2) {
        const int T_eff = sg.up ? 2 * sg.T_in : sg.T_in;
        const int src = rl * sg.stride + tap - (sg.taps >> 1);
        const bool ok = rok && src >= 0 && src < T_eff;
        const int st = sg.up ? (src >> 1) : src;
        const bool second = cb * CT >= sg.C1;
        const float* base = second ? sg.A2 : sg.A;
        const int ld = second ? sg.Cin - sg.C1 : sg.C1;
        const int coff = second ? cb * CT - sg.C1 : cb * CT;
        const float* arow = base + (ok ? ((size_t)rb * sg.T_in + st) * ld : (size_t)0) + coff + 4 * lp;
        const uint32_t da = tcw_smem_u32(astg + (g * CT + cr) * CT + 4 * lp);
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(da), "l"(arow), "r"(ok ? 16 : 0) : "memory");
      }
      if (what & 1) {
        const float* wbase = sg.Wh + (size_t)tap * a.N * sg.Cin + cb * CT + 4 * lp;
        const int plane = lp >> 2, p = 4 * g + (lp & 3);   // operand plane (hi / lo), 16-byte piece of the 128-byte K row
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int rr = cr + 32 * j;
          const int n = n0 + rr;
          const bool okw = n < a.N;
          const uint32_t dw = tcw_smem_u32(ob + (plane ? TC_W_LO : 0) + rr * 128 + ((p ^ (rr & 7)) << 4));
          asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dw), "l"(wbase + (size_t)(okw ? n : 0) * sg.Cin), "r"(okw ? 16 : 0)
                       : "memory");
        }
      }
    }
  }
}

// Token rows of chunk pair `pair` of a weight-stream unit (the `what & 2` part of tc_issue).  One thread runs ~30 of these
// per op on the critical path, so the segment descriptors are read into registers before the first copy (every cp.async
// carries a memory clobber: interleaved with the copies, each field was re-read from shared memory behind a dependent
// address computation -- measured 1.3 us per unit for issuing 0.17 us worth of loads), the row -> (sample, token) division
// is done once per unit by the caller, and the chunk -> (channel block, tap) division is by a constant.
__device__ __forceinline__ ChunkDesc make_chunk_desc(const ConvArgs& a, int n_chunks0, int f_begin, int f_end, int pair, int g) {
  ChunkDesc d;
  const int f = f_begin + 2 * pair + g;
  d.valid = f < f_end ? 1 : 0;
  const Seg& sg = a.seg[(d.valid && f >= n_chunks0) ? 1 : 0];
  const int q = d.valid ? (f >= n_chunks0 ? f - n_chunks0 : f) : 0;
  const int taps = sg.taps;
  const int cb = taps == 3 ? q / 3 : (taps == 1 ? q : q / taps);
  const int tap = q - cb * taps;
  const bool second = cb * CT >= sg.C1;
  d.base = (second ? sg.A2 : sg.A) + (second ? cb * CT - sg.C1 : cb * CT);
  d.ld = second ? sg.Cin - sg.C1 : sg.C1;
  d.shift = tap - (taps >> 1);
  d.stride = sg.stride; d.up = sg.up; d.T_in = sg.T_in;
  return d;
}
// token rows of one chunk pair from its two descriptors (shared memory)
__device__ __forceinline__ void tc_issue_rows_d(const ChunkDesc* cd, float* astg, int rb, int rl, bool rok) {
  const int tid = (int)threadIdx.x;
  const int lp = tid & 7, cr = tid >> 3;
  const ChunkDesc c0 = cd[0], c1 = cd[1];
  const float* src[2];
  int bytes[2];
#pragma unroll
  for (int g = 0; g < 2; ++g) {
    const ChunkDesc& c = g ? c1 : c0;
    const int T_eff = c.up ? 2 * c.T_in : c.T_in;
    const int sr = rl * c.stride + c.shift;
    const bool ok = c.valid && rok && sr >= 0 && sr < T_eff;
    const int st = c.up ? (sr >> 1) : sr;
    src[g] = c.base + (ok ? ((size_t)rb * c.T_in + st) * c.ld : (size_t)0) + 4 * lp;
    bytes[g] = ok ? 16 : 0;
  }
#pragma unroll
  for (int g = 0; g < 2; ++g) {
    if (g == 0 ? c0.valid : c1.valid) {
      const uint32_t da = tcw_smem_u32(astg + (g * CT + cr) * CT + 4 * lp);
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(da), "l"(src[g]), "r"(bytes[g]) : "memory");
    }
  }
}

struct SegRows { const float* A; const float* A2; int Cin, taps, stride, up, T_in, C1; };   // what the token-row copies need of a Seg
__device__ __forceinline__ SegRows seg_rows(const Seg& g) { return SegRows{g.A, g.A2, g.Cin, g.taps, g.stride, g.up, g.T_in, g.C1}; }
__device__ __forceinline__ void tc_issue_rows(const SegRows& s0, const SegRows& s1, int n_chunks0, int f_begin, int f_end, int pair, float* astg,
                                              int rb, int rl, bool rok) {
  const int tid = (int)threadIdx.x;
  const int lp = tid & 7, cr = tid >> 3;
  const float* src[2];
  int bytes[2];
#pragma unroll
  for (int g = 0; g < 2; ++g) {
    const int f = f_begin + 2 * pair + g;
    const bool have = f < f_end;
    const bool one = f >= n_chunks0;
    const int taps = one ? s1.taps : s0.taps, stride = one ? s1.stride : s0.stride, up = one ? s1.up : s0.up;
    const int T_in = one ? s1.T_in : s0.T_in, C1 = one ? s1.C1 : s0.C1, Cin = one ? s1.Cin : s0.Cin;
    const float* A = one ? s1.A : s0.A;
    const float* A2 = one ? s1.A2 : s0.A2;
    const int q = one ? f - n_chunks0 : f;
    const int cb = taps == 3 ? q / 3 : (taps == 1 ? q : q / taps);
    const int tap = q - cb * taps;
    const int T_eff = up ? 2 * T_in : T_in;
    const int sr = rl * stride + tap - (taps >> 1);
    const bool ok = have && rok && sr >= 0 && sr < T_eff;
    const int st = up ? (sr >> 1) : sr;
    const bool second = cb * CT >= C1;
    const float* base = second ? A2 : A;
    const int ld = second ? Cin - C1 : C1;
    const int coff = second ? cb * CT - C1 : cb * CT;
    src[g] = base + (ok ? ((size_t)rb * T_in + st) * ld : (size_t)0) + (have ? coff : 0) + 4 * lp;
    bytes[g] = ok ? 16 : 0;
  }
#pragma unroll
  for (int g = 0; g < 2; ++g) {
    if (f_begin + 2 * pair + g < f_end) {
      const uint32_t da = tcw_smem_u32(astg + (g * CT + cr) * CT + 4 * lp);
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(da), "l"(src[g]), "r"(bytes[g]) : "memory");
    }
  }
}

// weight copies of the first two chunk pairs of this CTA's first unit, issued before the barrier in front of the op
__device__ __forceinline__ void tc_preissue(const POp& o, float* smem, int cta) {
  const ConvArgs& a = o.conv;
  const int ks = o.ks;
  if (cta >= o.tiles_n * o.tiles_m * ks) return;
  const int crank = cta % ks, tile = cta / ks;
  const int n0 = (tile % o.tiles_n) * WN, m0 = (tile / o.tiles_n) * CT;
  const int n_chunks0 = a.seg[0].taps * (a.seg[0].Cin / CT);
  const int n_chunks = n_chunks0 + (a.nseg > 1 ? a.seg[1].taps * (a.seg[1].Cin / CT) : 0);
  const int f_begin = (crank * n_chunks) / ks, f_end = ((crank + 1) * n_chunks) / ks;
  const int n_pairs = (f_end - f_begin + 1) >> 1;
  uint8_t* ops = tc_ops(smem);
#pragma unroll
  for (int p = 0; p < 2; ++p)
    if (p < n_pairs) tc_issue(a, n_chunks0, f_begin, f_end, n0, m0, p, ops + p * TC_OP_BYTES, nullptr, 1);
}

// returns false when a wait timed out (the caller aborts the run)
// tma: the weights arrive through the weight stream (wstream_pump) instead of this CTA's cp.async copies
__device__ __forceinline__ bool p_conv_tc(const POp& o, float* smem, float* partials, unsigned* sems, unsigned* abort_flag, bool pre_issued,
                                          int cta, int G, long long* prof, TcState* ts, TcPhase& ph, bool tma, WCursor& wc) {
  const bool pf = prof != nullptr && cta == 0 && threadIdx.x == 0;
  const ConvArgs& a = o.conv;
  const int ks = o.ks;
  const int n_units = o.tiles_n * o.tiles_m * ks;
  const int tid = (int)threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int M = a.B * a.T_out;
  const int n_chunks0 = a.seg[0].taps * (a.seg[0].Cin / CT);
  const int n_chunks = n_chunks0 + (a.nseg > 1 ? a.seg[1].taps * (a.seg[1].Cin / CT) : 0);
  uint8_t* ops = tc_ops(smem);
  float* astg = tma ? reinterpret_cast<float*>(ops + TCX_STG_OFF) : tc_astage(smem);
  // operand tiles of ring stage s: weights (hi plane, lo plane at + TC_W_LO) and token rows (hi, lo at + CT * 128)
  auto w_tile = [&](int st) { return tma ? ops + st * TCX_W_BYTES : ops + st * TC_OP_BYTES; };
  auto a_tile = [&](int st) { return tma ? ops + TCX_A_OFF + st * TCX_A_BYTES : ops + st * TC_OP_BYTES + TC_A_HI; };
  const uint32_t idesc = tcw_idesc();
  const uint32_t tmem = ts->tmem_base;
  bool ok = true;
  // the operand stage of pair `pr` may be rewritten once the MMAs of its previous user have completed
  auto wait_stage = [&](int s) {
    if (ph.stage[s] > 0u && !tcw_mbar_wait(&ts->stage_free[s], (ph.stage[s] - 1u) & 1u)) ok = false;
  };
  for (int u = cta; u < n_units; u += G) {
    long long t0 = 0, t1 = 0, t3 = 0;
    if (pf) t0 = clock64();
    // unit -> (tile, slice) -> chunk range without integer divisions (see p_conv_wide)
    const int tile = div_small(u, ks, o.inv_ks), crank = u - tile * ks;
    const int tm = div_small(tile, o.tiles_n, o.inv_tn), tn = tile - tm * o.tiles_n;
    const int n0 = tn * WN, m0 = tm * CT;
    const int f_begin = div_small(crank * n_chunks, ks, o.inv_ks), f_end = div_small((crank + 1) * n_chunks, ks, o.inv_ks);
    const int n_pairs = (f_end - f_begin + 1) >> 1;
    // ring stage of the unit's pair: the weight stream numbers the pairs of the whole run, the local ring restarts per unit
    const uint32_t gp0 = tma ? ph.pairs : 0u;
    const int m_row = m0 + (tid >> 3);   // this thread's token row of the tile -> (sample, token)
    const bool row_ok = m_row < M;
    const int row_b = div_small(m_row, a.T_out, o.inv_T), row_l = m_row - row_b * a.T_out;
    auto stage_of = [&](int pair) { return (int)((gp0 + (uint32_t)pair) & (TC_STAGES - 1)); };
    auto issue_pair = [&](int pair, int what) {
      const int s = stage_of(pair);
      tc_issue(a, n_chunks0, f_begin, f_end, n0, m0, pair, w_tile(s), astg + s * TC_ASTG, tma ? (what & 2) : what);
      asm volatile("cp.async.commit_group;" ::: "memory");
    };
    // split one pair's token rows (fp32 staging -> fp16 hi / lo operand tiles): thread -> 8 consecutive channels (one 16-byte
    // piece of halves) of one row
    auto split_pair = [&](int p, int s) {
      const int rr = tid >> 3, pc = tid & 7;
      const int g = pc >> 2, kk0 = (pc & 3) * 8;
      const float* src = astg + s * TC_ASTG + (g * CT + rr) * CT + kk0;
      float4 x0 = *reinterpret_cast<const float4*>(src), x1 = *reinterpret_cast<const float4*>(src + 4);
      if (g == 1 && !(f_begin + 2 * p + 1 < f_end)) { x0 = make_float4(0.f, 0.f, 0.f, 0.f); x1 = x0; }   // odd chunk count
      {   // the fp16 split needs |x| < 65504 (and finite): an outlier activation is reported, not silently turned into inf
        const float m = fmaxf(fmaxf(fmaxf(fabsf(x0.x), fabsf(x0.y)), fmaxf(fabsf(x0.z), fabsf(x0.w))),
                              fmaxf(fmaxf(fabsf(x1.x), fabsf(x1.y)), fmaxf(fabsf(x1.z), fabsf(x1.w))));
        if (!(m < 6.0e4f)) atomicOr(abort_flag + 1, 1u);
      }
      uint4 hi, lo;
      split_f16x2(make_float2(x0.x, x0.y), hi.x, lo.x);
      split_f16x2(make_float2(x0.z, x0.w), hi.y, lo.y);
      split_f16x2(make_float2(x1.x, x1.y), hi.z, lo.z);
      split_f16x2(make_float2(x1.z, x1.w), hi.w, lo.w);
      uint8_t* at = a_tile(s);
      const int off = rr * 128 + ((pc ^ (rr & 7)) << 4);
      *reinterpret_cast<uint4*>(at + off) = hi;
      *reinterpret_cast<uint4*>(at + CT * 128 + off) = lo;
    };
    // thread 0: the pair's 12 (6) MMAs + the commits that free the stage / complete the unit
    auto mma_pair = [&](int p, int s) {
      if (tma) {   // this pair's weight image: issued by the stream long ago in the steady state
        const uint32_t par = ((gp0 + (uint32_t)p) / TC_STAGES) & 1u;
        uint32_t spins = 0;
        const long long tw0 = pf ? clock64() : 0;
        while (!tcw_mbar_test(&ts->full[s], par)) {
          wstream_pump(wc, ts, ops);
          if (++spins > (1u << 24)) { ok = false; break; }
        }
        if (pf) { prof[48] += clock64() - tw0; prof[51] += spins ? 1 : 0; }
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      }
      const uint32_t sw = tcw_smem_u32(w_tile(s)), sa = tcw_smem_u32(a_tile(s));
      const uint64_t whi = tcw_desc(sw), wlo = tcw_desc(sw + TC_W_LO), ahi = tcw_desc(sa), alo = tcw_desc(sa + CT * 128);
      const bool has2 = f_begin + 2 * p + 1 < f_end;
#pragma unroll
      for (int k = 0; k < 4; ++k) {   // 16 halves = 32 bytes along K inside the swizzle row: +2 in the (addr >> 4) field
        if (k >= 2 && !has2) break;    // odd chunk count: the second half of the K row is absent
        // Six independent accumulators (set = k & 1; W_hi A_hi | W_lo A_hi | W_hi A_lo): an MMA with 32 token columns is far
        // shorter than the tensor pipe's accumulate latency, so back-to-back MMAs into ONE accumulator ran at ~226 cycles each
        // (measured: 12 MMAs = 1.3 us per pair, the bound of the pair loop); now an accumulator is revisited every 6th MMA.
        const uint32_t acc = (p != 0 || k >= 2) ? 1u : 0u;
        const uint32_t set = tmem + (uint32_t)((k & 1) * 3 * CT);
        tcw_mma(set + CT, wlo + (uint64_t)(2 * k), ahi + (uint64_t)(2 * k), idesc, acc);
        tcw_mma(set, whi + (uint64_t)(2 * k), ahi + (uint64_t)(2 * k), idesc, acc);
        tcw_mma(set + 2 * CT, whi + (uint64_t)(2 * k), alo + (uint64_t)(2 * k), idesc, acc);
      }
      tcw_commit(&ts->stage_free[s]);
      if (p == n_pairs - 1) tcw_commit(&ts->done);
    };
    // Weight-stream flow: the MMAs of a pair are issued by TWO threads (a single thread pays ~110 cycles per tcgen05.mma: 12 per
    // pair were 1.9 us per unit).  Thread 0: W_hi x [A_hi ; A_lo] as ONE MMA with 64 token columns (the two token planes are
    // adjacent 32-row tiles) -> columns [0, 64) of the set; thread 32: W_lo x A_hi -> columns [64, 96).  4 + 4 instead of 12
    // instructions per pair; the accumulators are disjoint, both threads commit to the stage / unit barriers (count 2).
    auto full_wait = [&](int p, int s, int which) {   // the pair's weight image has landed (issued long ago in the steady state)
      const uint32_t par = ((gp0 + (uint32_t)p) / TC_STAGES) & 1u;
      uint32_t spins = 0;
      while (!tcw_mbar_test(&ts->full[s], par)) {
        if (which == 0) wstream_pump(wc, ts, ops);
        if (++spins > (1u << 24)) { ok = false; break; }
      }
      if (pf) prof[51] += spins ? 1 : 0;
    };
    auto mma_pair2 = [&](int p, int s, int which) {
      const uint32_t sw = tcw_smem_u32(w_tile(s)), sa = tcw_smem_u32(a_tile(s));
      const uint64_t wd = tcw_desc(sw + (which ? TC_W_LO : 0)), ad = tcw_desc(sa);
      const uint32_t id = which ? idesc : ((1u << 4) | ((uint32_t)((2 * CT) >> 3) << 17) | ((uint32_t)(WN >> 4) << 24));
      const bool has2 = f_begin + 2 * p + 1 < f_end;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        if (k >= 2 && !has2) break;
        const uint32_t acc = (p != 0 || k >= 2) ? 1u : 0u;
        const uint32_t set = tmem + (uint32_t)((k & 1) * 3 * CT) + (which ? 2u * CT : 0u);
        tcw_mma(set, wd + (uint64_t)(2 * k), ad + (uint64_t)(2 * k), id, acc);
      }
      tcw_commit(&ts->stage_free[s]);
      if (p == n_pairs - 1) tcw_commit(&ts->done);
    };
    if (tma) {
      // Weight stream: the only generic-proxy writes to the operand tiles are the split token rows, so a unit needs ONE
      // proxy fence per batch of up to TC_STAGES pairs instead of one per pair.  (Measured: with a fence behind every pair
      // the pair loop cost 1.3-2 us per pair although no weight image was ever late -- the fence waits for the copies that
      // are still in flight for the following pairs, which serialises the ring.)  All token rows of the batch are requested
      // at once: one L2 round trip per batch.
      for (int b0 = 0; b0 < n_pairs; b0 += TC_STAGES) {
        const int nb = n_pairs - b0 < TC_STAGES ? n_pairs - b0 : TC_STAGES;
        if (b0 > 0) {   // the stages' previous users are this unit's previous batch
          const long long ts0 = pf ? clock64() : 0;
          for (int i = 0; i < nb; ++i) wait_stage(stage_of(b0 + i));
          if (pf) prof[49] += clock64() - ts0;
          if (tid == 128) wstream_pump(wc, ts, ops);   // the previous batch's stages are free: request this batch's images now
        }
        if (tid < 2 * nb) ts->cd[tid] = make_chunk_desc(a, n_chunks0, f_begin, f_end, b0 + (tid >> 1), tid & 1);
        __syncthreads();
        for (int i = 0; i < nb; ++i) tc_issue_rows_d(ts->cd + 2 * i, astg + stage_of(b0 + i) * TC_ASTG, row_b, row_l, row_ok);
        asm volatile("cp.async.commit_group;" ::: "memory");
        const long long ti1 = pf ? clock64() : 0;
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        const long long ti2 = pf ? clock64() : 0;
        __syncthreads();   // the batch's token rows have landed for every thread
        if (pf && b0 == 0) { prof[56] += ti1 - t0; prof[57] += ti2 - ti1; }
        if (pf && b0 == 0) { prof[27] += clock64() - t0; prof[26] += n_pairs; }
        const long long tq0 = pf ? clock64() : 0;
        for (int i = 0; i < nb; ++i) split_pair(b0 + i, stage_of(b0 + i));
        const long long tq1 = pf ? clock64() : 0;
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes (split) -> tensor core reads
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        const long long tq2 = pf ? clock64() : 0;
        __syncthreads();
        if (tid == 0 || tid == 32) {
          const int which = tid == 32 ? 1 : 0;
          const long long tq3 = pf ? clock64() : 0;
          {   // the batch's weight images: the tests are independent (their latencies overlap); waiting is the rare path
            uint32_t landed = 1u;
            for (int i = 0; i < nb; ++i)
              landed &= tcw_mbar_test(&ts->full[stage_of(b0 + i)], ((gp0 + (uint32_t)(b0 + i)) / TC_STAGES) & 1u) ? 1u : 0u;
            if (!landed)
              for (int i = 0; i < nb; ++i) full_wait(b0 + i, stage_of(b0 + i), which);
          }
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          if (pf) prof[48] += clock64() - tq3;
          for (int i = 0; i < nb; ++i) mma_pair2(b0 + i, stage_of(b0 + i), which);
          const long long tq4 = pf ? clock64() : 0;
          if (pf) { prof[52] += tq1 - tq0; prof[53] += tq2 - tq1; prof[54] += tq3 - tq2; prof[55] += tq4 - tq3; }
        }
        for (int i = 0; i < nb; ++i) ph.stage[stage_of(b0 + i)] += 1u;
      }
    } else {
    // (every MMA of the previous unit / op has completed: its `done` wait; stages 0 and 1 are free)
    // Token rows are requested three pairs ahead; pair p + 3 reuses the stage of pair p - 1, whose MMAs were committed one
    // iteration ago.
    const bool w_pre = pre_issued && u == cta;   // local ring: the weights of pairs 0 and 1 were copied before the barrier
#pragma unroll
    for (int p = 0; p < 3; ++p) {
      if (p < n_pairs) issue_pair(p, (w_pre && p < 2) ? 2 : 3);
      else asm volatile("cp.async.commit_group;" ::: "memory");
    }
    for (int p = 0; p < n_pairs; ++p) {
      const int s = stage_of(p);
      if (p + 3 < n_pairs) {
        const long long ts0 = pf ? clock64() : 0;
        wait_stage(stage_of(p + 3));   // read by the MMAs of pair p - 1 (issued one iteration ago)
        if (pf) prof[49] += clock64() - ts0;
        issue_pair(p + 3, 3);
      } else {
        asm volatile("cp.async.commit_group;" ::: "memory");
      }
      asm volatile("cp.async.wait_group 3;" ::: "memory");
      __syncthreads();   // pair p has landed for every thread
      if (pf && p == 0) { prof[27] += clock64() - t0; prof[26] += n_pairs; }
      split_pair(p, s);
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes (copies, split) -> tensor core reads
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncthreads();
      if (tid == 0) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        mma_pair(p, s);
      }
      ph.stage[s] += 1u;
    }
    }
    ph.pairs += (uint32_t)n_pairs;
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    ph.done += 1u;
    const long long td0 = pf ? clock64() : 0;
    // one thread polls the mbarrier, the others sleep at the CTA barrier (255 threads spinning on try_wait slowed the MMA
    // issuer's own shared-memory operations down)
    if (tid == 0 && !tcw_mbar_wait(&ts->done, (ph.done - 1u) & 1u)) ok = false;
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (pf) t1 = clock64();
    if (pf) prof[50] += t1 - td0;
    if (tma && tid == 128) wstream_pump(wc, ts, ops);   // the unit's stages are free: refill them while warps 0-3 read the accumulators
    float* mine = partials + ((size_t)tile * ks + crank) * (CT * WN);
    if (warp < 4) {
      uint32_t r1[32], r2[32];
      {   // sum the two accumulator sets: r1 = W_hi A_hi, r2 = W_lo A_hi + W_hi A_lo (scaled by 4096)
        const uint32_t row = tmem + ((uint32_t)(warp * 32) << 16);
        uint32_t t[32];
        tcw_ld32(row, r1);
        tcw_ld32(row + 3 * CT, t);
#pragma unroll
        for (int j = 0; j < 32; ++j) r1[j] = __float_as_uint(__uint_as_float(r1[j]) + __uint_as_float(t[j]));
        tcw_ld32(row + CT, r2);
        tcw_ld32(row + 2 * CT, t);
#pragma unroll
        for (int j = 0; j < 32; ++j) r2[j] = __float_as_uint(__uint_as_float(r2[j]) + __uint_as_float(t[j]));
        tcw_ld32(row + 4 * CT, t);
#pragma unroll
        for (int j = 0; j < 32; ++j) r2[j] = __float_as_uint(__uint_as_float(r2[j]) + __uint_as_float(t[j]));
        tcw_ld32(row + 5 * CT, t);
#pragma unroll
        for (int j = 0; j < 32; ++j) r2[j] = __float_as_uint(__uint_as_float(r2[j]) + __uint_as_float(t[j]));
      }
      const int n = n0 + warp * 32 + lane;
      if (ks == 1) {
        if (n < a.N) {
          const float bv = a.bias[n];
#pragma unroll
          for (int j = 0; j < CT; ++j) {
            const int m = m0 + j;
            if (m < M) {
              float v = __uint_as_float(r1[j]) + __uint_as_float(r2[j]) * (1.0f / F16_LO_SCALE);
              v += bv;
              if (a.emb) v += a.emb[(size_t)(m / a.T_out) * a.emb_ld + n];
              if (a.residual) v += a.residual[(size_t)m * a.N + n];
              a.out[(size_t)m * a.N + n] = v;
            }
          }
        }
      } else {
#pragma unroll
        for (int j = 0; j < CT; ++j)
          __stcg(mine + j * WN + warp * 32 + lane, __uint_as_float(r1[j]) + __uint_as_float(r2[j]) * (1.0f / F16_LO_SCALE));
      }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();   // the accumulators have been read: the next unit may overwrite them
    if (o.p2p != nullptr && tid == 0) red_release_add(o.p2p + tile, 1u);   // this slice's partial tile is published
    if (ks > 1 && !o.defer) {
      if (tid == 0) {
        red_release_add(sems + tile, 1u);
        spin_until(sems + tile, (unsigned)ks, abort_flag);
      }
      __syncthreads();
      if (pf) t3 = clock64();
      const int e0 = (crank * (CT * WN)) / ks, e1 = ((crank + 1) * (CT * WN)) / ks;
      const float* base = partials + (size_t)tile * ks * (CT * WN);
      for (int idx = e0 + tid; idx < e1; idx += 256) {
        const int m = m0 + (idx >> 7), n = n0 + (idx & (WN - 1));
        const bool live = m < M && n < a.N;
        float pv[P_MAX_KS_WIDE];
#pragma unroll
        for (int j = 0; j < P_MAX_KS_WIDE; ++j) pv[j] = j < ks ? __ldcg(base + (size_t)j * (CT * WN) + idx) : 0.f;
        float bv = 0.f, ev = 0.f, rv = 0.f;
        if (live) {
          bv = a.bias[n];
          if (a.emb) ev = a.emb[(size_t)(m / a.T_out) * a.emb_ld + n];
          if (a.residual) rv = a.residual[(size_t)m * a.N + n];
        }
        float v = 0.f;
#pragma unroll
        for (int j = 0; j < P_MAX_KS_WIDE; ++j)
          if (j < ks) v += pv[j];
        if (live) {
          v += bv;
          if (a.emb) v += ev;
          if (a.residual) v += rv;
          a.out[(size_t)m * a.N + n] = v;
        }
      }
      __syncthreads();
    }
    if (pf) {
      const long long t4 = clock64();
      prof[0] += t1 - t0;                      // chunk stream (copies + split + MMA)
      prof[2] += (t3 ? t3 : t4) - t1;          // accumulator read-out, publish (+ wait for the other slices)
      prof[24] += t3 ? t4 - t3 : 0;            // distributed slice reduction + epilogue
      prof[25] += 1;
    }
  }
  return ok;
}

// ---- GroupNorm unit (128 threads = one half of the CTA), arithmetic of gn_kernel -------------------------------
// at least `want` arrivals (wrap-safe); false: aborted
__device__ __forceinline__ bool spin_at_least(const unsigned* p, unsigned want, unsigned* abort_flag) {
  const long long t0 = clock64();
  unsigned spins = 0;
  while ((int)(ld_acquire(p) - want) < 0) {
    if ((++spins & 1023u) == 0u) {
      if (*(volatile unsigned*)abort_flag != 0u || clock64() - t0 > 2000000000LL) {
        atomicExch(abort_flag, 1u);
        return false;
      }
    }
  }
  return true;
}

// exec: how many times this op has run in this launch, this time included (p2p hand-off: arrivals per tile = ks * exec)
__device__ __forceinline__ void p_gn_unit(const POp& o, int b, int g, int half, float* red, long long* prof = nullptr, unsigned exec = 0u,
                                          unsigned* abort_flag = nullptr) {
  const int tid = threadIdx.x & 127;
  const bool pf = prof != nullptr && threadIdx.x == 0;
  const long long tg0 = pf ? clock64() : 0;
  const float* in1 = o.in0;
  const float* in2 = o.in1;
  const int C1 = o.i0, C2 = o.i1, T = o.i2, silu = o.i3;
  const float* gamma = o.w0;
  const float* beta = o.w1;
  float* out = o.out0;
  float* raw = o.out1;
  const int C = C1 + C2;
  const int cg = C / 32;
  const int n = T * cg;
  const PartSrc ps = o.ps;   // into registers: the descriptor lives in shared memory and the barriers below carry memory clobbers
  const bool deferred = ps.ks > 0;
  if (deferred && ps.p2p != nullptr) {
    // No grid barrier separates this op from the GEMM that produced its first input: wait for the K slices of exactly the
    // output tiles this (sample, group) unit reads -- the sample's token rows x the group's channels, at most 2 x 2 tiles.
    if (tid == 0) {
      const int c_lo = g * cg, c_hi = min((g + 1) * cg, C1) - 1;
      if (c_lo <= c_hi) {
        const unsigned want = (unsigned)ps.ks * exec;
        for (int tm = (b * T) >> 5; tm <= (b * T + T - 1) >> 5; ++tm)
          for (int tn = c_lo >> 7; tn <= c_hi >> 7; ++tn) spin_at_least(ps.p2p + tm * ps.tiles_n + tn, want, abort_flag);
      }
    }
    named_bar(1 + half, 128);
  }
  float* fin = const_cast<float*>(o.in0);   // a deferred first input is finalised in place for its later consumers
  auto load = [&](int idx) -> float {
    const int t = idx / cg, c = g * cg + idx % cg;
    if (deferred && c < C1) return part_value(ps, T, b * T + t, c, b);
    return c < C1 ? in1[((size_t)b * T + t) * C1 + c] : in2[((size_t)b * T + t) * C2 + (c - C1)];
  };
  auto block_sum = [&](float v) -> float {
#pragma unroll
    for (int of = 16; of > 0; of >>= 1) v += __shfl_xor_sync(0xffffffffu, v, of);
    named_bar(1 + half, 128);
    if ((tid & 31) == 0) red[tid >> 5] = v;
    named_bar(1 + half, 128);
    return red[0] + red[1] + red[2] + red[3];
  };
  // the unit's values are read once and kept in registers for the three passes (n <= 6 * 128 in this network; longer
  // groups re-read); same per-thread element order and the same summation trees as gn_kernel
  constexpr int PGN_CACHE = 6;
  float xc[PGN_CACHE], gc[PGN_CACHE], bc[PGN_CACHE];
  int tt[PGN_CACHE], cc[PGN_CACHE];   // (token, channel) of the cached elements, computed once (the unit is issue-bound on index math)
  const float inv_cg = 1.0f / (float)cg;
  auto load_tc = [&](int t, int c) -> float {
    if (deferred && c < C1) return part_value(ps, T, b * T + t, c, b);
    return c < C1 ? in1[((size_t)b * T + t) * C1 + c] : in2[((size_t)b * T + t) * C2 + (c - C1)];
  };
#pragma unroll
  for (int k = 0; k < PGN_CACHE; ++k) {   // scale / shift are requested with the values (not after the statistics)
    const int i = tid + 128 * k;
    const int t = i < n ? div_small(i, cg, inv_cg) : 0;
    tt[k] = t; cc[k] = g * cg + (i < n ? i - t * cg : 0);
    gc[k] = gamma[cc[k]]; bc[k] = beta[cc[k]];
  }
#pragma unroll
  for (int k = 0; k < PGN_CACHE; ++k) xc[k] = tid + 128 * k < n ? load_tc(tt[k], cc[k]) : 0.f;
  float s = 0.f;
#pragma unroll
  for (int k = 0; k < PGN_CACHE; ++k) if (tid + 128 * k < n) s += xc[k];
  for (int i = tid + 128 * PGN_CACHE; i < n; i += 128) s += load(i);
  const long long tg1 = pf ? clock64() : 0;
  const float mean = block_sum(s) / (float)n;
  const long long tg2 = pf ? clock64() : 0;
  float q = 0.f;
#pragma unroll
  for (int k = 0; k < PGN_CACHE; ++k) if (tid + 128 * k < n) { const float d = xc[k] - mean; q += d * d; }
  for (int i = tid + 128 * PGN_CACHE; i < n; i += 128) { const float d = load(i) - mean; q += d * d; }
  const float var = block_sum(q) / (float)n;
  const float rstd = 1.0f / sqrtf(var + 1e-5f);
  const long long tg3 = pf ? clock64() : 0;
  auto emit = [&](int t, int c, float x, float gm, float bt) {
    float y = (x - mean) * rstd * gm + bt;
    if (silu) y = y / (1.0f + expf(-y));
    const size_t oo = ((size_t)b * T + t) * C + c;
    out[oo] = y;
    if (raw) raw[oo] = x;
    if (deferred && c < C1) fin[((size_t)b * T + t) * C1 + c] = x;
  };
#pragma unroll
  for (int k = 0; k < PGN_CACHE; ++k) if (tid + 128 * k < n) emit(tt[k], cc[k], xc[k], gc[k], bc[k]);
  for (int i = tid + 128 * PGN_CACHE; i < n; i += 128) { const int c = g * cg + i % cg; emit(i / cg, c, load(i), gamma[c], beta[c]); }
  if (pf) { prof[58] += tg1 - tg0; prof[59] += tg2 - tg1; prof[60] += tg3 - tg2; prof[61] += clock64() - tg3; prof[62] += 1; }
}

// ---- small-M linears: rows [mb, mb+8) staged in shared memory (xs, row stride K), two output columns per warp pass ----
// Same per-lane k order and the same shuffle tree as linear_rows_kernel.  K <= 896 (7 float4 per lane).
__device__ __forceinline__ void p_lin_cols(const POp& o, const float* xs, int mb, int rows, int cta, int G) {
  const int K = o.i1, N = o.i2, out_silu = o.i4;
  const float* W = o.w0;
  const float* bias = o.w1;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_stride = G * 8;
  // the weight rows of pass i+1 are requested before pass i is multiplied (two column pairs in flight per warp)
  float4 wa[7], wb[7], na[7], nb[7];
  float ba = 0.f, bb = 0.f, nba = 0.f, nbb = 0.f;
  auto fetch = [&](int nA, float4 (&ra)[7], float4 (&rb)[7], float& bA, float& bB) {
    const int nB = nA + n_stride;
    bA = nA < N ? __ldg(bias + nA) : 0.f;
    bB = nB < N ? __ldg(bias + nB) : 0.f;
#pragma unroll
    for (int i = 0; i < 7; ++i) {
      const int k = lane * 4 + 128 * i;
      ra[i] = make_float4(0.f, 0.f, 0.f, 0.f);
      rb[i] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (k < K && nA < N) {
        ra[i] = __ldg(reinterpret_cast<const float4*>(W + (size_t)nA * K + k));
        if (nB < N) rb[i] = __ldg(reinterpret_cast<const float4*>(W + (size_t)nB * K + k));
      }
    }
  };
  fetch(cta * 8 + warp, wa, wb, ba, bb);
  for (int nA = cta * 8 + warp; nA < N; nA += 2 * n_stride) {
    const int nB = nA + n_stride;
    const bool hasB = nB < N;
    fetch(nA + 2 * n_stride, na, nb, nba, nbb);
    float accA[8], accB[8];
#pragma unroll
    for (int r = 0; r < 8; ++r) { accA[r] = 0.f; accB[r] = 0.f; }
#pragma unroll
    for (int i = 0; i < 7; ++i) {
      const int k = lane * 4 + 128 * i;
      if (k < K) {
#pragma unroll
        for (int r = 0; r < 8; ++r) {
          if (r < rows) {
            const float4 x = *reinterpret_cast<const float4*>(xs + r * K + k);
            accA[r] = fmaf(x.x, wa[i].x, accA[r]); accA[r] = fmaf(x.y, wa[i].y, accA[r]);
            accA[r] = fmaf(x.z, wa[i].z, accA[r]); accA[r] = fmaf(x.w, wa[i].w, accA[r]);
            accB[r] = fmaf(x.x, wb[i].x, accB[r]); accB[r] = fmaf(x.y, wb[i].y, accB[r]);
            accB[r] = fmaf(x.z, wb[i].z, accB[r]); accB[r] = fmaf(x.w, wb[i].w, accB[r]);
          }
        }
      }
    }
#pragma unroll
    for (int r = 0; r < 8; ++r)
#pragma unroll
      for (int of = 16; of > 0; of >>= 1) {
        accA[r] += __shfl_xor_sync(0xffffffffu, accA[r], of);
        accB[r] += __shfl_xor_sync(0xffffffffu, accB[r], of);
      }
    // every lane holds all 16 sums: lane r finishes row r of column A, lane 8 + r row r of column B
    if (lane < 16) {
      const int r = lane & 7, c = lane >> 3;
      float sum = 0.f;
#pragma unroll
      for (int rr = 0; rr < 8; ++rr)
        if (r == rr) sum = c ? accB[rr] : accA[rr];
      const int n = c ? nB : nA;
      if (r < rows && (c == 0 || hasB)) {
        float v = sum + (c ? bb : ba);
        if (out_silu) v = v / (1.0f + expf(-v));
        const int row = mb + r;
        if (o.lab) v = v + o.w2[(size_t)o.lab[row] * EMB + n];    // label_emb row (label_add_kernel)
        if (o.in1) v = v + o.in1[(size_t)row * N + n];            // projected context (accumulating linear)
        o.out0[(size_t)row * N + n] = v;
        if (o.out1) o.out1[(size_t)row * N + n] = v / (1.0f + expf(-v));   // SiLU once, for the consumer
      }
    }
#pragma unroll
    for (int i = 0; i < 7; ++i) { wa[i] = na[i]; wb[i] = nb[i]; }
    ba = nba; bb = nbb;
  }
}

template <int MODE>
__global__ void __launch_bounds__(256, 1) unet_persistent_kernel(PersistArgs pa) {
  extern __shared__ __align__(16) float smem[];
  __shared__ POp sop[2];
  __shared__ float gn_red[2][4];
  __shared__ int s_ok, s_last;
  const int G = (int)gridDim.x, cta = (int)blockIdx.x, tid = (int)threadIdx.x;
  const int half = tid >> 7;
  const int n_ops = pa.n_emb + pa.n_prog;
  unsigned target = 0;
  unsigned seq = 0;   // ops executed so far (selects the arrival-counter bank)
  bool pre_issued = false;   // the current op's first weight copies were issued before the barrier in front of it
  auto fetch_op = [&](int slot, int oi) {
    const int* src = reinterpret_cast<const int*>(pa.ops + oi);
    int* dst = reinterpret_cast<int*>(&sop[slot]);
    for (int i = tid; i < (int)(sizeof(POp) / sizeof(int)); i += 256) dst[i] = src[i];
  };
  static_assert(sizeof(POp) <= 256 * sizeof(int), "descriptor must fit one word per thread");
  fetch_op(0, 0);
  __shared__ TcState tcs;
  __shared__ WCursor s_wcur;   // weight-stream position: advanced by thread 0 (op start, grid barrier) and, inside a token
                               // GEMM, by thread 128 (a warp that has no part in the accumulator read-out); CTA barriers separate them
  TcPhase tph{};
  const bool use_tc = MODE == 1 && pa.use_tc != 0;
  const bool use_tma = use_tc && pa.use_tma != 0;
  if (use_tma && tid == 0) {
    WCursor c{};
    c.n = pa.wl_count[cta];
    c.list = pa.wlist + pa.wl_start[cta];
    c.reps = c.n ? (uint32_t)(pa.n_steps * pa.n_pass) : 0u;
    s_wcur = c;
  }
  // scratch of the attention / embedding ops: behind the weight ring when the stream owns the front of the shared memory
  float* const scr = use_tma ? reinterpret_cast<float*>(tc_ops(smem) + TCX_STG_OFF) : smem;
  auto pump = [&]() { if (MODE == 1 && use_tma) wstream_pump(s_wcur, &tcs, tc_ops(smem)); };   // thread 0 only
  if (use_tc) {   // tcgen05 path: mbarriers + 64 TMEM columns (two 128 x 32 fp32 accumulators), held for the whole run
    if (tid == 0) {
      // weight-stream flow: two MMA-issuing threads commit to the stage / unit barriers
      for (int i = 0; i < TC_STAGES; ++i) tcw_mbar_init(&tcs.stage_free[i], use_tma ? 2 : 1);
      for (int i = 0; i < TC_STAGES; ++i) tcw_mbar_init(&tcs.full[i], 1);
      tcw_mbar_init(&tcs.done, use_tma ? 2 : 1);
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (tid < 32) {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 256;" ::"r"(tcw_smem_u32(&tcs.tmem_base)) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  }
  __syncthreads();
  if (use_tc) asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  if (use_tma && tid == 0) pump();
  int slot = 0;
  [&]() {   // the op loop; `return` leaves it early when the run is aborted (TMEM is released below in every case)
  for (int iter = 0; iter < pa.n_steps; ++iter) {
    const int sidx = pa.n_steps - 1 - iter;
    for (int pass = 0; pass < pa.n_pass; ++pass) {
      for (int oi = pass == 0 ? 0 : pa.n_emb; oi < n_ops; ++oi) {
        // descriptor of the op after this one (same order as this loop nest)
        int next = oi + 1;
        bool has_next = true;
        if (next == n_ops) {
          if (pass + 1 < pa.n_pass) next = pa.n_emb;
          else if (iter + 1 < pa.n_steps) next = 0;
          else has_next = false;
        }
        // the next descriptor travels through a register while this op runs (its load latency stays off the critical path)
        int next_word = 0;
        if (has_next && tid < (int)(sizeof(POp) / sizeof(int))) next_word = reinterpret_cast<const int*>(pa.ops + next)[tid];
        const POp& o = sop[slot];
        const bool profiled = pa.prof != nullptr && tid == 0 && (cta == 0 || cta == G - 1);
        long long t_op0 = 0;
        if (profiled) t_op0 = clock64();
        if (tid == 0) pump();
        switch (o.type) {
          case P_EMB1: {   // timestep embedding (temb_kernel) + first time_embed linear with SiLU
            const int B = pa.B;
            const float tf = (float)pa.tmap[sidx];
            for (int mb = 0; mb < B; mb += 8) {
              const int rows = B - mb < 8 ? B - mb : 8;
              __syncthreads();
              for (int i = tid; i < rows * (TCH / 2); i += 256) {
                const int k = i % (TCH / 2), r = i / (TCH / 2);
                const float c = -9.210340371976184f;
                const float f = expf(__fdiv_rn(__fmul_rn(c, (float)k), 112.0f));
                const float arg = __fmul_rn(tf, f);
                scr[r * TCH + k] = cosf(arg);
                scr[r * TCH + TCH / 2 + k] = sinf(arg);
              }
              __syncthreads();
              p_lin_cols(o, scr, mb, rows, cta, G);
            }
            break;
          }
          case P_LIN: {
            const int M = o.i0, K = o.i1, in_silu = o.i3;
            for (int mb = 0; mb < M; mb += 8) {
              const int rows = M - mb < 8 ? M - mb : 8;
              __syncthreads();
              // independent loads are requested eight at a time (a plain loop pays one L2 round trip per iteration)
              for (int i0 = tid; i0 < rows * K; i0 += 8 * 256) {
                float xv[8];
#pragma unroll
                for (int q = 0; q < 8; ++q) { const int i = i0 + 256 * q; xv[q] = i < rows * K ? o.in0[(size_t)mb * K + i] : 0.f; }
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                  const int i = i0 + 256 * q;
                  if (i < rows * K) {
                    float x = xv[q];
                    if (in_silu) x = x / (1.0f + expf(-x));
                    scr[i] = x;
                  }
                }
              }
              __syncthreads();
              p_lin_cols(o, scr, mb, rows, cta, G);
            }
            break;
          }
          case P_INCONV: {
            const int L = o.i0, N = o.i1;
            const int total = pa.B * L * N;
            for (int i = cta * 256 + tid; i < total; i += G * 256) {
              const int n = i % N, l = (i / N) % L, b = i / (N * L);
              const float* xr = o.in0 + (size_t)b * L;
              float acc = 0.f;
              if (l > 0) acc = fmaf(o.w0[n], xr[l - 1], acc);
              acc = fmaf(o.w0[N + n], xr[l], acc);
              if (l < L - 1) acc = fmaf(o.w0[2 * N + n], xr[l + 1], acc);
              o.out0[i] = acc + o.w1[n];
            }
            break;
          }
          case P_GN: {
            const int n_units = pa.B * 32;
            for (int u = cta * 2 + half; u < n_units; u += 2 * G) p_gn_unit(o, u >> 5, u & 31, half, gn_red[half], cta == 0 ? pa.prof : nullptr,
                                                                              (unsigned)(iter * pa.n_pass + pass + 1), pa.sync + 1);
            break;
          }
          case P_CONV:
            if (MODE == 1 && o.wide == 2) {
              unsigned* bank = pa.sems + (size_t)(seq & 1u) * pa.sem_bank;
              if (!p_conv_tc(o, smem, pa.partials, bank, pa.sync + 1, pre_issued, cta, G, pa.prof, &tcs, tph, use_tma, s_wcur)) {
                if (tid == 0) atomicExch(pa.sync + 1, 1u);   // a tensor-core wait timed out: abort the run (reported by surfd_unet_status)
              }
            } else if (MODE != 0 && o.wide) {
              // arrival counters: ops alternate between two banks; the bank of the previous op is cleared here (its
              // waiters all passed the grid barrier in front of this op, its next users arrive behind the one after it)
              unsigned* bank = pa.sems + (size_t)(seq & 1u) * pa.sem_bank;
              p_conv_wide<(MODE == 0 ? 1 : MODE)>(o, smem, pa.partials, bank, pa.sync + 1, pre_issued, cta, G, pa.prof);
            } else {
              p_conv<MODE>(o, smem, pa.partials, pa.sems + 2 * (size_t)pa.sem_bank, &s_last, cta, G, pa.prof);
            }
            break;
          case P_ATTN: {   // one (batch, head) unit per CTA pass, same arithmetic as attn_kernel
            const int heads = o.i2;
            const int n_units = pa.B * heads;
            for (int u = cta; u < n_units; u += G)
              if (o.ps.ks > 0)
                attn_unit(o.in0, o.i0, o.i1, heads, o.f0, o.out0, u / heads, u % heads, scr, tid, 256, [] { __syncthreads(); },
                          [&](const float*, int m, int c) { return part_value(o.ps, o.i1, m, c); });
              else
                attn_unit(o.in0, o.i0, o.i1, heads, o.f0, o.out0, u / heads, u % heads, scr, tid, 256, [] { __syncthreads(); },
                          [](const float* p, int, int) { return *p; });
            break;
          }
          case P_OUTCONV: {   // last conv (224 -> 1) + the DDPM update of the element it produces
            const int C = o.i0, L = o.i1;
            const int warp = tid >> 5, lane = tid & 31;
            if (pa.pair_B > 0) {
              // guidance pair in one batch: x0 = x0b + scale * (x0a - x0b) with x0a / x0b the predictions of rows b / pair_B + b
              // (models/cfg_sampler.py:19-26); both rows receive the updated sample
              const int half = pa.pair_B * L;
              for (int wi = cta * 8 + warp; wi < half; wi += G * 8) {
                const int l = wi % L, b = wi / L;
                float acc_a = 0.f, acc_b = 0.f;
                for (int tap = 0; tap < 3; ++tap) {
                  const int src = l + tap - 1;
                  if (src < 0 || src >= L) continue;
                  const float* ar = o.in0 + ((size_t)b * L + src) * C;
                  const float* br = ar + (size_t)half * C;
                  const float* wr = o.w0 + (size_t)tap * C;
                  for (int c = lane; c < C; c += 32) { acc_a = fmaf(ar[c], wr[c], acc_a); acc_b = fmaf(br[c], wr[c], acc_b); }
                }
#pragma unroll
                for (int of = 16; of > 0; of >>= 1) {
                  acc_a += __shfl_xor_sync(0xffffffffu, acc_a, of);
                  acc_b += __shfl_xor_sync(0xffffffffu, acc_b, of);
                }
                if (lane == 0) {
                  const float x0a = acc_a + o.w1[0], x0b = acc_b + o.w1[0];
                  const float x0 = __fadd_rn(x0b, __fmul_rn(pa.guidance, __fsub_rn(x0a, x0b)));
                  const int ns = pa.n_steps;
                  const float c1 = pa.coef[sidx], c2 = pa.coef[ns + sidx], sd = pa.coef[2 * ns + sidx];
                  const float mean = __fadd_rn(__fmul_rn(c1, x0), __fmul_rn(c2, pa.x[wi]));
                  const float mask = sidx != 0 ? 1.0f : 0.0f;
                  const float nz = pa.noise[(size_t)(1 + iter) * pa.noise_stride + wi];
                  const float xn = __fadd_rn(mean, __fmul_rn(__fmul_rn(mask, sd), nz));
                  pa.x[wi] = xn;
                  pa.x[half + wi] = xn;
                }
              }
              break;
            }
            for (int wi = cta * 8 + warp; wi < pa.B * L; wi += G * 8) {
              const int l = wi % L, b = wi / L;
              float acc = 0.f;
              for (int tap = 0; tap < 3; ++tap) {
                const int src = l + tap - 1;
                if (src < 0 || src >= L) continue;
                const float* ar = o.in0 + ((size_t)b * L + src) * C;
                const float* wr = o.w0 + (size_t)tap * C;
                for (int c = lane; c < C; c += 32) acc = fmaf(ar[c], wr[c], acc);
              }
#pragma unroll
              for (int of = 16; of > 0; of >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, of);
              if (lane == 0) {
                float x0 = acc + o.w1[0];
                if (pass + 1 < pa.n_pass) {
                  pa.x0a[wi] = x0;
                } else {
                  if (pa.n_pass == 2) {   // classifier-free guidance replay: x0 = x0b + scale * (x0a - x0b)
                    const float x0a = pa.x0a[wi];
                    x0 = __fadd_rn(x0, __fmul_rn(pa.guidance, __fsub_rn(x0a, x0)));
                  }
                  const int ns = pa.n_steps;
                  const float c1 = pa.coef[sidx], c2 = pa.coef[ns + sidx], sd = pa.coef[2 * ns + sidx];
                  const float mean = __fadd_rn(__fmul_rn(c1, x0), __fmul_rn(c2, pa.x[wi]));
                  const float mask = sidx != 0 ? 1.0f : 0.0f;
                  const float nz = pa.noise[(size_t)(1 + iter) * pa.noise_stride + wi];
                  pa.x[wi] = __fadd_rn(mean, __fmul_rn(__fmul_rn(mask, sd), nz));
                }
              }
            }
            break;
          }
          default:
            break;
        }
        if (cta == 0) {
          unsigned* other = pa.sems + (size_t)((seq + 1u) & 1u) * pa.sem_bank;
          for (int i = tid; i < pa.sem_bank; i += 256) other[i] = 0u;
        }
        ++seq;
        if (has_next && tid < (int)(sizeof(POp) / sizeof(int))) reinterpret_cast<int*>(&sop[slot ^ 1])[tid] = next_word;
        __syncthreads();   // the op is complete in this CTA, the next descriptor is in place
        long long t_op1 = 0;
        if (profiled) t_op1 = clock64();
        // A deferred token GEMM with per-tile arrival counters hands its partial tiles to the GroupNorm behind it point to
        // point: no grid barrier here, the GroupNorm units wait for the tiles they read (p_gn_unit).
        const bool no_barrier = MODE == 1 && sop[slot].type == P_CONV && sop[slot].p2p != nullptr;
        pre_issued = false;
        if (!no_barrier) {
          grid_arrive(pa.sync);
          if (tid == 0) pump();
          if (MODE != 0 && !use_tma && has_next && sop[slot ^ 1].type == P_CONV && sop[slot ^ 1].wide) {
            if (sop[slot ^ 1].wide == 2) tc_preissue(sop[slot ^ 1], smem, cta);
            else wide_preissue<WTP, W_STAGE, W_STAGES>(sop[slot ^ 1], smem, cta);
            pre_issued = true;
          }
          if (!grid_wait(pa.sync, target, (unsigned)G, &s_ok, pump)) return;
        } else if (tid == 0) {
          pump();
        }
        if (profiled) {
          long long* pr = pa.prof + ((cta == 0 ? 0 : 8) + sop[slot].type) * 3;
          pr[0] += t_op1 - t_op0;
          pr[1] += clock64() - t_op1;
          pr[2] += 1;
        }
        slot ^= 1;
      }
    }
  }
  }();
  if (use_tc) {
    __syncthreads();
    if (tid < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 256;" ::"r"(tcs.tmem_base) : "memory");
  }
}

}  // namespace surfd

using namespace surfd;

// cudaLaunchKernelEx with an optional cluster dimension (z) and the programmatic-dependent-launch attribute
#define UNET_LAUNCH(pdl, kernel, grid, block, smem, st, cluster_z, ...)                                   \
  do {                                                                                                    \
    cudaLaunchConfig_t _cfg{};                                                                            \
    _cfg.gridDim = (grid); _cfg.blockDim = (block); _cfg.dynamicSmemBytes = (smem); _cfg.stream = (st);   \
    cudaLaunchAttribute _attr[2];                                                                         \
    unsigned _na = 0;                                                                                     \
    if ((cluster_z) > 0) {                                                                                \
      _attr[_na].id = cudaLaunchAttributeClusterDimension;                                                \
      _attr[_na].val.clusterDim.x = 1; _attr[_na].val.clusterDim.y = 1;                                   \
      _attr[_na].val.clusterDim.z = (unsigned)(cluster_z); ++_na;                                         \
    }                                                                                                     \
    if (pdl) {                                                                                            \
      _attr[_na].id = cudaLaunchAttributeProgrammaticStreamSerialization;                                 \
      _attr[_na].val.programmaticStreamSerializationAllowed = 1; ++_na;                                   \
    }                                                                                                     \
    _cfg.attrs = _attr; _cfg.numAttrs = _na;                                                              \
    SURFD_CUDA(cudaLaunchKernelEx(&_cfg, kernel, __VA_ARGS__));                                           \
    SURFD_CHECK_LAUNCH();                                                                                 \
  } while (0)

// per-warp A/W staging (8 warps x CONV_STAGES x 2 x 32 x 36 floats); its first part is reused for the 8 x 32 x 33 warp
// partials, followed (at 8*32*33 floats) by the 32x32 CTA partial
static constexpr int CONV_SMEM = (8 * CONV_STAGES * 2 * CT * CTP) * (int)sizeof(float);
static_assert(8 * CONV_STAGES * 2 * CT * CTP >= 8 * CT * 33 + CT * CT, "reduction scratch must fit in the staging area");

// One lane = private activation pool + step state + captured step graph + stream.  The denoiser is latency-bound (about 170
// dependent small kernels per step on <= 256 tokens), so a batch is split over lanes that run concurrently on their own
// streams; the weights are shared.
struct Lane {
  DevBuf pool, emb_all, temb, e1, emb, t_cur, x0a, x0b, xcur, state;
  // persistent sampler: op descriptors, K-slice scratch, semaphores, barrier word + abort flag
  DevBuf emb_silu, ctxv, p_ops, p_partials, p_sems, p_sync, p_prof, p_wlist, p_wlidx, p_wstream, ctx2, lab2, p_p2p;
  size_t p_p2p_words = 0;
  int p_use_tma = 0;
  int p_B = -1, p_n_emb = 0, p_n_prog = 0, p_smem = 0, p_grid = 0, p_split = -1, p_wide = -1, p_sem_bank = 0, p_use_tc = 0;
  const float* p_ctx = nullptr;
  const int64_t* p_lab = nullptr;
  cudaStream_t stream = nullptr;
  cudaEvent_t done = nullptr;
  int cap = 0;   // batch capacity of the buffers
  cudaGraphExec_t graph_exec = nullptr;
  int graph_B = -1;
  const float* graph_ctx = nullptr;
  const int64_t* graph_lab = nullptr;
  const int64_t* graph_tmap = nullptr;
  const float* graph_coef = nullptr;
  const float* graph_noise = nullptr;
  int64_t graph_stride = 0;
  float graph_guidance = 1.f;
  int64_t graph_launches = 0;
  std::vector<size_t> buf_off;
  float* buf(int64_t id) const { return pool.as<float>() + buf_off[(size_t)id]; }
  void release() {
    if (graph_exec) cudaGraphExecDestroy(graph_exec);
    graph_exec = nullptr;
    pool.release(); emb_all.release(); temb.release(); e1.release(); emb.release(); t_cur.release(); x0a.release(); x0b.release();
    xcur.release(); state.release();
    emb_silu.release(); ctxv.release(); p_ops.release(); p_partials.release(); p_sems.release(); p_sync.release(); p_prof.release();
    p_wlist.release(); p_wlidx.release(); p_wstream.release(); ctx2.release(); lab2.release(); p_p2p.release();
    p_B = -1;
    if (stream) cudaStreamDestroy(stream);
    if (done) cudaEventDestroy(done);
    stream = nullptr; done = nullptr;
  }
};

struct surfd_unet {
  int L = 0, max_batch = 0;
  bool pdl = true;     // programmatic dependent launch between the step's kernels (falls back to false if capture rejects it)
  int sampler = 1;     // surfd_sample: 1 = persistent cooperative kernel (default), 0 = CUDA-graph replay of the step kernels
                       // (also the fallback when cooperative launch is unavailable)
  int sampler_sms = 0; // CTAs of the persistent kernel (0 = one per SM)
  bool profile = false; // persistent kernel: per-op-type cycle counters (diagnostics)
  int persist_tc = 1;      // precision mode 1: the wide units' products run on tcgen05 (0 / SURFD_UNET_DEBUG bit 5: mma.sync)
  int persist_fuse_gn = 0; // wide units: 1 = GroupNorm ops are folded into the token GEMM behind them.  Off: measured slower (1.95 vs
                           // 1.33 ms/step) -- the per-unit statistics prologue is instruction-bound and every GEMM then pays the in-op
                           // K-slice exchange instead of leaving it to the GroupNorm op.  SURFD_UNET_DEBUG bit 4 switches it on.
  int persist_defer = 1; // wide units: consumers sum the K slices of the GEMM in front of them (0 = exchange inside the GEMM op)
  int persist_split = 1; // token-GEMM K split of the persistent kernel: 0 = the graph path's rule (bit-identical samples),
                         // 1 = as many slices as fit in ONE round of the resident CTAs (faster; same fp32-class accuracy)
  int num_sms = 0;
  bool coop = false;
  unsigned* h_abort = nullptr;   // pinned: abort flag of the last persistent run
  int precision = 1;   // token GEMMs: 0 fp32 FFMA, 1 3xTF32 mma.sync (fp32-class accuracy, default), 2 single-pass TF32
  DevBuf weights;
  DevBuf weights_split;   // tcgen05 units: token-GEMM weights with every 32-float block stored as its fp16 hi | lo split
  bool split_done = false;
  std::vector<int64_t> hdr, buf_sizes;
  std::vector<std::vector<int64_t>> prog;
  int emb_cols = 0;
  size_t n_floats = 0;
  std::vector<Lane> lanes;
  cudaEvent_t fork = nullptr;
  const float* w(int64_t off) const { return weights.as<float>() + off; }
};

extern "C" size_t surfd_unet_packed_floats(void) { return 0; }  // size depends on cond_mode; python validates via the arch walk

static int lane_init(surfd_unet* u, Lane& ln, int cap) {
  ln.cap = cap;
  size_t tot = 0;
  ln.buf_off.clear();
  for (size_t i = 0; i < u->buf_sizes.size(); ++i) { ln.buf_off.push_back(tot); tot += (size_t)u->buf_sizes[i] * cap; }
  const size_t B = (size_t)cap;
  SURFD_TRY(ln.pool.reserve(tot * sizeof(float)));
  SURFD_TRY(ln.emb_all.reserve(B * u->emb_cols * sizeof(float)));
  SURFD_TRY(ln.temb.reserve(B * TCH * sizeof(float)));
  SURFD_TRY(ln.e1.reserve(B * EMB * sizeof(float)));
  SURFD_TRY(ln.emb.reserve(B * EMB * sizeof(float)));
  SURFD_TRY(ln.t_cur.reserve(B * sizeof(int64_t)));
  SURFD_TRY(ln.x0a.reserve(B * u->L * sizeof(float)));
  SURFD_TRY(ln.x0b.reserve(B * u->L * sizeof(float)));
  SURFD_TRY(ln.xcur.reserve(B * u->L * sizeof(float)));
  SURFD_TRY(ln.state.reserve(sizeof(StepState)));
  SURFD_TRY(ln.emb_silu.reserve(B * EMB * sizeof(float)));
  SURFD_TRY(ln.ctxv.reserve(B * EMB * sizeof(float)));
  ln.p_B = -1;
  if (!ln.stream) SURFD_CUDA(cudaStreamCreateWithFlags(&ln.stream, cudaStreamNonBlocking));
  if (!ln.done) SURFD_CUDA(cudaEventCreateWithFlags(&ln.done, cudaEventDisableTiming));
  return 0;
}

extern "C" int surfd_unet_set_precision(surfd_unet* u, int mode) {
  SURFD_REQUIRE(u != nullptr && mode >= 0 && mode <= 2, "precision mode must be 0 (fp32), 1 (3xTF32) or 2 (TF32)");
  if (mode != u->precision) {
    SURFD_CUDA(cudaDeviceSynchronize());
    for (auto& ln : u->lanes) {   // captured step graphs embed the kernel variant
      if (ln.graph_exec) { cudaGraphExecDestroy(ln.graph_exec); ln.graph_exec = nullptr; }
    }
    u->precision = mode;
  }
  return 0;
}

extern "C" int surfd_unet_set_lanes(surfd_unet* u, int n_lanes) {
  SURFD_REQUIRE(u != nullptr, "null argument");
  SURFD_REQUIRE(n_lanes >= 1 && n_lanes <= 64, "n_lanes out of range");
  SURFD_CUDA(cudaDeviceSynchronize());
  for (auto& ln : u->lanes) ln.release();
  u->lanes.assign((size_t)n_lanes, Lane());
  const int cap0 = 2 * u->max_batch;                               // lane 0 also serves surfd_unet_forward at full batch and the
                                                                   // persistent engine's guidance pairs (2 rows per sample)
  const int cap = (u->max_batch + n_lanes - 1) / n_lanes;
  for (int i = 0; i < n_lanes; ++i) SURFD_TRY(lane_init(u, u->lanes[(size_t)i], i == 0 ? cap0 : cap));
  return 0;
}

extern "C" int surfd_unet_create(const float* packed, size_t n_floats, const int64_t* program, size_t n_prog, int L, int max_batch,
                                 surfd_unet** out) {
  SURFD_REQUIRE(packed && program && out, "null argument");
  SURFD_REQUIRE(n_prog >= 16, "program too short");
  SURFD_REQUIRE(max_batch >= 1 && max_batch <= 4096, "max_batch out of range");
  surfd_unet* u = new surfd_unet();
  u->L = L; u->max_batch = max_batch; u->n_floats = n_floats;
  u->hdr.assign(program, program + 16);
  const int64_t nb = u->hdr[0], np = u->hdr[1];
  if ((size_t)(16 + nb + np * REC) != n_prog || u->hdr[13] != L) {
    delete u;
    return set_error(SURFD_BAD_ARGUMENT, "program/packing mismatch", __FILE__, __LINE__);
  }
  u->emb_cols = (int)u->hdr[2];
  u->buf_sizes.assign(program + 16, program + 16 + nb);
  for (int64_t i = 0; i < np; ++i) u->prog.emplace_back(program + 16 + nb + i * REC, program + 16 + nb + (i + 1) * REC);
  auto fail = [&](int code) { surfd_unet_destroy(u); return code; };
  int st;
  if ((st = u->weights.reserve(n_floats * sizeof(float)))) return fail(st);
  cudaError_t ce = cudaMemcpy(u->weights.p, packed, n_floats * sizeof(float), cudaMemcpyDefault);
  if (ce != cudaSuccess) return fail(set_error(-(int)ce, cudaGetErrorString(ce), __FILE__, __LINE__));
  ce = cudaFuncSetAttribute(conv_gemm_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, CONV_SMEM);
  if (ce == cudaSuccess) ce = cudaFuncSetAttribute(conv_gemm_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, CONV_SMEM);
  if (ce == cudaSuccess) ce = cudaFuncSetAttribute(conv_gemm_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, CONV_SMEM);
  if (ce != cudaSuccess) return fail(set_error(-(int)ce, cudaGetErrorString(ce), __FILE__, __LINE__));
  ce = cudaEventCreateWithFlags(&u->fork, cudaEventDisableTiming);
  if (ce != cudaSuccess) return fail(set_error(-(int)ce, cudaGetErrorString(ce), __FILE__, __LINE__));
  {
    int dev = 0, v = 0;
    cudaGetDevice(&dev);
    if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess) u->num_sms = v;
    if (cudaDeviceGetAttribute(&v, cudaDevAttrCooperativeLaunch, dev) == cudaSuccess) u->coop = v != 0;
    ce = cudaHostAlloc(&u->h_abort, 2 * sizeof(unsigned), cudaHostAllocDefault);
    if (ce != cudaSuccess) return fail(set_error(-(int)ce, cudaGetErrorString(ce), __FILE__, __LINE__));
    u->h_abort[0] = u->h_abort[1] = 0u;
  }
  // measured on B200 (B=8, L=32): 1 lane 2.6 ms/step, 8 concurrent lanes 5.6 ms/step -- many tiny cluster launches from
  // several streams contend in the front end, so the default is a single lane; set_lanes() stays for experiments.
  if ((st = surfd_unet_set_lanes(u, 1))) return fail(st);
  *out = u;
  return 0;
}

extern "C" void surfd_unet_destroy(surfd_unet* u) {
  if (!u) return;
  cudaDeviceSynchronize();
  for (auto& ln : u->lanes) ln.release();
  if (u->fork) cudaEventDestroy(u->fork);
  if (u->h_abort) cudaFreeHost(u->h_abort);
  u->weights.release();
  u->weights_split.release();
  delete u;
}

// one model evaluation on a lane's buffers: x [B][L], t [B] -> x0 [B][L]
static int unet_run(surfd_unet* u, Lane& ln, int B, const float* x, const int64_t* t, const float* ctx, const int64_t* lab, float* x0,
                    cudaStream_t st) {
  const auto& h = u->hdr;
  // ---- embedding ----
  UNET_LAUNCH(u->pdl, temb_kernel, dim3((unsigned)cdiv((int64_t)B * (TCH / 2), 128)), dim3(128), 0, st, 0, t, B, ln.temb.as<float>());
  const dim3 g1((unsigned)cdiv(EMB, 8), (unsigned)cdiv(B, 8));
  UNET_LAUNCH(u->pdl, linear_rows_kernel, dim3(g1), dim3(256), 0, st, 0, ln.temb.as<float>(), B, TCH, u->w(h[5]), u->w(h[6]), EMB, 0, 1, 0, ln.e1.as<float>());
  UNET_LAUNCH(u->pdl, linear_rows_kernel, dim3(g1), dim3(256), 0, st, 0, ln.e1.as<float>(), B, EMB, u->w(h[7]), u->w(h[8]), EMB, 0, 0, 0, ln.emb.as<float>());
  if (lab) {
    SURFD_REQUIRE(h[11] >= 0, "labels given but the checkpoint has no label_emb");
    UNET_LAUNCH(u->pdl, label_add_kernel, dim3((unsigned)cdiv((int64_t)B * EMB, 256)), dim3(256), 0, st, 0, lab, B, u->w(h[11]), ln.emb.as<float>());
  }
  if (ctx) {
    UNET_LAUNCH(u->pdl, linear_rows_kernel, dim3(g1), dim3(256), 0, st, 0, ctx, B, CTX, u->w(h[9]), u->w(h[10]), EMB, 0, 0, 1, ln.emb.as<float>());
  }
  const dim3 g2((unsigned)cdiv(u->emb_cols, 8), (unsigned)cdiv(B, 8));
  UNET_LAUNCH(u->pdl, linear_rows_kernel, dim3(g2), dim3(256), 0, st, 0, ln.emb.as<float>(), B, EMB, u->w(h[3]), u->w(h[4]), u->emb_cols, 1, 0, 0, ln.emb_all.as<float>());
  // ---- program ----
  for (const auto& r : u->prog) {
    switch (r[0]) {
      case OP_INCONV: {
        const int N = (int)r[2], L = (int)r[3];
        UNET_LAUNCH(u->pdl, inconv_kernel, dim3((unsigned)cdiv((int64_t)B * L * N, 256)), dim3(256), 0, st, 0, x, B, L, N, u->w(r[4]), u->w(r[5]), ln.buf(r[1]));
        break;
      }
      case OP_GN: {
        const int C1 = (int)r[2], C2 = (int)r[4], T = (int)r[5];
        UNET_LAUNCH(u->pdl, gn_kernel, dim3((unsigned)(B * 32)), dim3(128), 0, st, 0, ln.buf(r[1]), C1, r[3] >= 0 ? ln.buf(r[3]) : nullptr, C2, T, u->w(r[9]), u->w(r[10]),
                                                      (int)r[8], ln.buf(r[6]), r[7] >= 0 ? ln.buf(r[7]) : nullptr);
        break;
      }
      case OP_CONV: {
        ConvArgs a{};
        a.out = ln.buf(r[1]); a.N = (int)r[2]; a.T_out = (int)r[3]; a.nseg = (int)r[4]; a.B = B;
        for (int s = 0; s < a.nseg; ++s) {
          const int64_t* q = &r[5 + 7 * s];
          a.seg[s].A = ln.buf(q[0]); a.seg[s].Cin = (int)q[1]; a.seg[s].taps = (int)q[2]; a.seg[s].stride = (int)q[3];
          a.seg[s].up = (int)q[4]; a.seg[s].T_in = (int)q[5]; a.seg[s].W = u->w(q[6]);
        }
        a.bias = u->w(r[19]);
        a.emb = r[20] >= 0 ? ln.emb_all.as<float>() + r[20] : nullptr;
        a.emb_ld = u->emb_cols;
        a.residual = r[21] >= 0 ? ln.buf(r[21]) : nullptr;
        // K split: enough CTAs to cover the SMs, but at least ~2 chunks of work per warp (8 warps per CTA)
        int chunks = 0;
        for (int s = 0; s < a.nseg; ++s) chunks += a.seg[s].taps * (a.seg[s].Cin / CT);
        // (decided from the per-sample tile count so the fp32 summation order -- and the result -- does not depend on B)
        const int tiles = (a.N / CT) * (int)cdiv((int64_t)a.T_out, CT);
        int ksplit = 1;
        while (ksplit < KSPLIT && tiles * ksplit < 148 && chunks >= 16 * (ksplit * 2)) ksplit *= 2;
        const dim3 cgrid((unsigned)(a.N / CT), (unsigned)cdiv((int64_t)B * a.T_out, CT), (unsigned)ksplit);
        if (u->precision == 0) UNET_LAUNCH(u->pdl, conv_gemm_kernel<0>, cgrid, dim3(256), CONV_SMEM, st, ksplit, a);
        else if (u->precision == 1) UNET_LAUNCH(u->pdl, conv_gemm_kernel<1>, cgrid, dim3(256), CONV_SMEM, st, ksplit, a);
        else UNET_LAUNCH(u->pdl, conv_gemm_kernel<2>, cgrid, dim3(256), CONV_SMEM, st, ksplit, a);
        break;
      }
      case OP_ATTN: {
        const int C = (int)r[2], T = (int)r[3], heads = (int)r[5];
        const int ch = C / heads;
        const float scale = (float)(1.0 / sqrt(sqrt((double)ch)));
        const size_t smem = (size_t)attn_smem_floats(T, ch) * sizeof(float);
        UNET_LAUNCH(u->pdl, attn_kernel, dim3((unsigned)(B * heads)), dim3(128), smem, st, 0, ln.buf(r[1]), C, T, heads, scale, ln.buf(r[4]));
        break;
      }
      case OP_OUTCONV: {
        const int C = (int)r[2], T = (int)r[3];
        UNET_LAUNCH(u->pdl, outconv_kernel, dim3((unsigned)cdiv((int64_t)B * T * 32, 256)), dim3(256), 0, st, 0, ln.buf(r[1]), B, T, C, u->w(r[4]), u->w(r[5]), x0);
        break;
      }
      default:
        return set_error(SURFD_BAD_ARGUMENT, "unknown op in program", __FILE__, __LINE__);
    }
  }
  return 0;
}

extern "C" int surfd_unet_forward(surfd_unet* u, int B, const float* x_dev, const int64_t* t_dev, const float* context_dev,
                                  const int64_t* labels_dev, float* out_dev, void* stream) {
  SURFD_REQUIRE(u && x_dev && t_dev && out_dev, "null argument");
  SURFD_REQUIRE(B >= 1 && B <= u->max_batch, "batch exceeds max_batch");
  return unet_run(u, u->lanes[0], B, x_dev, t_dev, context_dev, labels_dev, out_dev, (cudaStream_t)stream);
}

static int record_step(surfd_unet* u, Lane& ln, int B, const int64_t* tmap, const float* coef, const float* noise, int64_t noise_stride,
                       const float* ctx, const int64_t* lab, float guidance, cudaStream_t st) {
  StepState* ss = ln.state.as<StepState>();
  UNET_LAUNCH(u->pdl, step_begin_kernel, dim3((unsigned)cdiv(B, 128)), dim3(128), 0, st, 0, ss, tmap, B, ln.t_cur.as<int64_t>());
  SURFD_TRY(unet_run(u, ln, B, ln.xcur.as<float>(), ln.t_cur.as<int64_t>(), ctx, lab, ln.x0a.as<float>(), st));
  const bool cfg = guidance != 1.0f;
  if (cfg) SURFD_TRY(unet_run(u, ln, B, ln.xcur.as<float>(), ln.t_cur.as<int64_t>(), ctx, lab, ln.x0b.as<float>(), st));
  const int n = B * u->L;
  UNET_LAUNCH(u->pdl, ddpm_update_kernel, dim3((unsigned)cdiv(n, 128)), dim3(128), 0, st, 0, ss, coef, ln.x0a.as<float>(), cfg ? ln.x0b.as<float>() : nullptr, guidance, noise,
                                                             noise_stride, n, ln.xcur.as<float>());
  UNET_LAUNCH(u->pdl, step_advance_kernel, dim3(1), dim3(1), 0, st, 0, ss);
  return 0;
}

// ---- persistent sampler (host side) ---------------------------------------------------------------------------------
extern "C" int surfd_unet_set_sampler(surfd_unet* u, int mode, int n_sms) {
  SURFD_REQUIRE(u != nullptr && mode >= 0 && mode <= 2, "sampler mode must be 0 (graph replay), 1 (persistent kernel) or 2 (persistent, graph-identical K split)");
  SURFD_REQUIRE(n_sms >= 0, "n_sms must be >= 0 (0 = one CTA per SM)");
  u->sampler = mode ? 1 : 0;
  u->persist_split = mode == 2 ? 0 : 1;
  u->sampler_sms = n_sms;
  return 0;
}

// Diagnostics: enable (out == NULL, on != 0) / disable per-op-type cycle counters of the persistent sampler, or read the
// counters of the last run (out != NULL): out[(half * 8 + op type) * 3 + {0 body cycles, 1 barrier cycles, 2 count}],
// half 0 = first CTA, half 1 = last CTA.  Synchronises the device when reading.
extern "C" int surfd_unet_profile(surfd_unet* u, int on, int64_t* out) {
  SURFD_REQUIRE(u != nullptr, "null argument");
  if (!out) { u->profile = on != 0; return 0; }
  SURFD_CUDA(cudaDeviceSynchronize());
  for (int i = 0; i < 64; ++i) out[i] = 0;
  if (u->lanes[0].p_prof.p) SURFD_CUDA(cudaMemcpy(out, u->lanes[0].p_prof.p, 64 * sizeof(long long), cudaMemcpyDeviceToHost));
  return 0;
}

extern "C" int surfd_unet_status(surfd_unet* u) {
  SURFD_REQUIRE(u != nullptr, "null argument");
  if (u->h_abort && u->h_abort[0] != 0u)
    return set_error(SURFD_ABORTED, "persistent sampler aborted: a grid barrier timed out", __FILE__, __LINE__);
  if (u->h_abort && u->h_abort[1] != 0u)
    return set_error(SURFD_RANGE, "persistent sampler: an activation left the fp16 range of the split-product token GEMMs (|x| >= 6e4 or "
                                  "not finite); the samples of this call are invalid -- use set_precision(0) or the graph engine (set_sampler(0))",
                     __FILE__, __LINE__);
  return 0;
}

template <int MODE>
static int persist_launch(const PersistArgs& pa, int grid, int smem, cudaStream_t st) {
  SURFD_CUDA(cudaFuncSetAttribute(unet_persistent_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  PersistArgs args = pa;
  void* kargs[] = {&args};
  SURFD_CUDA(cudaLaunchCooperativeKernel((void*)unet_persistent_kernel<MODE>, dim3((unsigned)grid), dim3(256), kargs, (size_t)smem, st));
  g_launch_count += 1;
  return 0;
}

template <int MODE>
static int persist_max_grid(int smem, int num_sms, int* out) {
  SURFD_CUDA(cudaFuncSetAttribute(unet_persistent_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  int per_sm = 0;
  SURFD_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, unet_persistent_kernel<MODE>, 256, (size_t)smem));
  *out = per_sm * num_sms;
  return 0;
}

// op descriptors for batch B on lane `ln` (cached per (B, ctx, lab))
static int persist_build(surfd_unet* u, Lane& ln, int B, const float* ctx, const int64_t* lab, int grid) {
  const bool wide = u->persist_split == 1 && u->precision != 0;
  const char* dbg_env0 = getenv("SURFD_UNET_DEBUG");
  const bool use_tc = wide && u->precision == 1 && u->persist_tc && !((dbg_env0 ? atoi(dbg_env0) : 0) & 32);
  // diagnostics (tools/sampler_profile.py): SURFD_UNET_DEBUG bit 1 = no deferred exchange, bit 3 = no fused GroupNorm, bit 4 = fused
  // GroupNorm on, bit 5 = mma.sync instead of tcgen05 units, bit 6 = attention sums the qkv GEMM's K slices,
  // bit 2 = every op replaced by an empty one (barrier cost only; results are garbage)
  const char* dbg_env = getenv("SURFD_UNET_DEBUG");
  const int dbg = dbg_env ? atoi(dbg_env) : 0;
  if (ln.p_B == B && ln.p_ctx == ctx && ln.p_lab == lab && ln.p_grid == grid && ln.p_split == u->persist_split && ln.p_wide == (int)wide + (int)use_tc) return 0;
  const auto& h = u->hdr;
  if (use_tc && !u->split_done) {
    // load-time fp16 split of every token-GEMM weight tensor (same offsets as the fp32 blob), see p_conv_tc
    SURFD_TRY(u->weights_split.reserve(u->n_floats * sizeof(float)));
    for (const auto& r : u->prog) {
      if (r[0] != OP_CONV) continue;
      for (int sgi = 0; sgi < (int)r[4]; ++sgi) {
        const int64_t* q = &r[5 + 7 * sgi];
        const size_t n_blocks = (size_t)q[2] * (size_t)r[2] * (size_t)q[1] / 32;   // taps * N * Cin / 32
        split_weights_kernel<<<(unsigned)cdiv((int64_t)(n_blocks * 16), 256), 256>>>(u->w(q[6]), u->weights_split.as<float>() + q[6], n_blocks);
        SURFD_CHECK_LAUNCH();
      }
    }
    SURFD_CUDA(cudaDeviceSynchronize());
    u->split_done = true;
  }
  std::vector<POp> ops;
  auto blank = [](int type) { POp o; memset(&o, 0, sizeof(o)); o.type = type; return o; };
  {
    POp o = blank(P_EMB1);   // e1 = SiLU(W1 temb(t) + b1)
    o.w0 = u->w(h[5]); o.w1 = u->w(h[6]); o.i0 = B; o.i1 = TCH; o.i2 = EMB; o.i4 = 1; o.out0 = ln.e1.as<float>();
    ops.push_back(o);
  }
  {
    POp o = blank(P_LIN);    // emb = W2 e1 + b2 (+ label row) (+ projected context); also SiLU(emb) for the next op
    o.in0 = ln.e1.as<float>(); o.i0 = B; o.i1 = EMB; o.i2 = EMB; o.w0 = u->w(h[7]); o.w1 = u->w(h[8]);
    if (lab) {
      SURFD_REQUIRE(h[11] >= 0, "labels given but the checkpoint has no label_emb");
      o.lab = lab; o.w2 = u->w(h[11]);
    }
    if (ctx) o.in1 = ln.ctxv.as<float>();
    o.out0 = ln.emb.as<float>(); o.out1 = ln.emb_silu.as<float>();
    ops.push_back(o);
  }
  {
    POp o = blank(P_LIN);    // all 22 emb_layers at once
    o.in0 = ln.emb_silu.as<float>(); o.i0 = B; o.i1 = EMB; o.i2 = u->emb_cols; o.w0 = u->w(h[3]); o.w1 = u->w(h[4]);
    o.out0 = ln.emb_all.as<float>();
    ops.push_back(o);
  }
  const int n_emb = (int)ops.size();
  size_t max_partial_tiles = 0, max_tiles = 1;
  int smem_floats = (wide ? (use_tc ? (TC_SMEM > WIDE_SMEM ? TC_SMEM : WIDE_SMEM) : WIDE_SMEM) : CONV_SMEM) / (int)sizeof(float);
  if (smem_floats < 8 * EMB) smem_floats = 8 * EMB;
  for (const auto& r : u->prog) {
    switch (r[0]) {
      case OP_INCONV: {
        POp o = blank(P_INCONV);
        o.in0 = ln.xcur.as<float>(); o.i0 = (int)r[3]; o.i1 = (int)r[2]; o.w0 = u->w(r[4]); o.w1 = u->w(r[5]); o.out0 = ln.buf(r[1]);
        ops.push_back(o);
        break;
      }
      case OP_GN: {
        POp o = blank(P_GN);
        o.in0 = ln.buf(r[1]); o.i0 = (int)r[2]; o.in1 = r[3] >= 0 ? ln.buf(r[3]) : nullptr; o.i1 = (int)r[4]; o.i2 = (int)r[5];
        o.i3 = (int)r[8]; o.w0 = u->w(r[9]); o.w1 = u->w(r[10]); o.out0 = ln.buf(r[6]); o.out1 = r[7] >= 0 ? ln.buf(r[7]) : nullptr;
        ops.push_back(o);
        break;
      }
      case OP_CONV: {
        POp o = blank(P_CONV);
        ConvArgs& a = o.conv;
        a.out = ln.buf(r[1]); a.N = (int)r[2]; a.T_out = (int)r[3]; a.nseg = (int)r[4]; a.B = B;
        int chunks = 0;
        for (int s = 0; s < a.nseg; ++s) {
          const int64_t* q = &r[5 + 7 * s];
          a.seg[s].A = ln.buf(q[0]); a.seg[s].Cin = (int)q[1]; a.seg[s].taps = (int)q[2]; a.seg[s].stride = (int)q[3];
          a.seg[s].up = (int)q[4]; a.seg[s].T_in = (int)q[5]; a.seg[s].W = u->w(q[6]);
          a.seg[s].A2 = nullptr; a.seg[s].C1 = a.seg[s].Cin;
          a.seg[s].Wh = use_tc ? u->weights_split.as<float>() + q[6] : nullptr;
          chunks += a.seg[s].taps * (a.seg[s].Cin / CT);
        }
        a.bias = u->w(r[19]);
        a.emb = r[20] >= 0 ? ln.emb_all.as<float>() + r[20] : nullptr;
        a.emb_ld = u->emb_cols;
        a.residual = r[21] >= 0 ? ln.buf(r[21]) : nullptr;
        const int tiles = (a.N / CT) * (int)cdiv((int64_t)a.T_out, CT);
        int ksplit = 1;
        if (wide) {
          // wide units: 32 x 128 tiles; one round -- every (tile, slice) unit gets its own CTA, and a slice keeps at
          // least one chunk pair; K is only split when all slices of all tiles are co-resident
          o.wide = use_tc ? 2 : 1;
          o.tiles_n = (int)cdiv((int64_t)a.N, WN); o.tiles_m = (int)cdiv((int64_t)B * a.T_out, CT);
          const size_t nt = (size_t)o.tiles_n * o.tiles_m;
          ksplit = (int)((int64_t)grid / (int64_t)nt);
          if (ksplit > chunks / 2) ksplit = chunks / 2;
          if (ksplit > P_MAX_KS_WIDE) ksplit = P_MAX_KS_WIDE;
          if (ksplit < 1) ksplit = 1;
          o.ks = ksplit;
          o.inv_ks = 1.0f / (float)o.ks; o.inv_tn = 1.0f / (float)o.tiles_n; o.inv_T = 1.0f / (float)a.T_out;
          if (nt > max_tiles) max_tiles = nt;
          if (ksplit > 1 && nt * ksplit * (WN / CT) > max_partial_tiles) max_partial_tiles = nt * ksplit * (WN / CT);
          ops.push_back(o);
          break;
        }
        o.tiles_n = a.N / CT; o.tiles_m = (int)cdiv((int64_t)B * a.T_out, CT);
        const size_t nt = (size_t)o.tiles_n * o.tiles_m;
        if (u->persist_split == 0) {
          // same K-split rule as the graph path (decided per sample, so results do not depend on B)
          while (ksplit < KSPLIT && tiles * ksplit < 148 && chunks >= 16 * (ksplit * 2)) ksplit *= 2;
        } else {
          // one round: every (tile, slice) unit gets its own CTA (any slice count, not only powers of two); a slice
          // keeps at least one chunk per warp
          ksplit = (int)((int64_t)grid / (int64_t)nt);
          if (ksplit > chunks / 8) ksplit = chunks / 8;
          if (ksplit > P_MAX_KS) ksplit = P_MAX_KS;
          if (ksplit < 1) ksplit = 1;
        }
        o.ks = ksplit;
        if (nt > max_tiles) max_tiles = nt;
        if (ksplit > 1 && nt * ksplit > max_partial_tiles) max_partial_tiles = nt * ksplit;
        ops.push_back(o);
        break;
      }
      case OP_ATTN: {
        POp o = blank(P_ATTN);
        const int C = (int)r[2], T = (int)r[3], heads = (int)r[5];
        const int ch = C / heads;
        o.in0 = ln.buf(r[1]); o.i0 = C; o.i1 = T; o.i2 = heads; o.i3 = attn_smem_floats(T, ch);
        o.f0 = (float)(1.0 / sqrt(sqrt((double)ch)));
        o.out0 = ln.buf(r[4]);
        if (o.i3 > smem_floats) smem_floats = o.i3;
        ops.push_back(o);
        break;
      }
      case OP_OUTCONV: {
        POp o = blank(P_OUTCONV);
        o.in0 = ln.buf(r[1]); o.i0 = (int)r[2]; o.i1 = (int)r[3]; o.w0 = u->w(r[4]); o.w1 = u->w(r[5]);
        ops.push_back(o);
        break;
      }
      default:
        return set_error(SURFD_BAD_ARGUMENT, "unknown op in program", __FILE__, __LINE__);
    }
  }
  // fused GroupNorm: a GroupNorm op whose output feeds segment 0 of the wide token GEMM right behind it disappears -- the
  // GEMM's units normalise their own token rows (statistics recomputed per unit from the raw input, which the producer
  // finalised); a later 1x1 skip segment that read the GroupNorm's raw concat copy reads the two concat sources instead
  if (wide && (u->persist_fuse_gn || (dbg & 16)) && !(dbg & 8)) {
    struct RawSrc { const float* raw; const float* in1; const float* in2; int C1; };
    std::vector<RawSrc> raws;
    std::vector<POp> fused;
    for (size_t i = 0; i < ops.size(); ++i) {
      POp g = ops[i];
      if (g.type == P_CONV && g.conv.nseg > 1) {
        for (const auto& rsrc : raws)
          if (g.conv.seg[1].A == rsrc.raw) { g.conv.seg[1].A = rsrc.in1; g.conv.seg[1].A2 = rsrc.in2; g.conv.seg[1].C1 = rsrc.C1; }
      }
      if (g.type == P_GN && i + 1 < ops.size() && ops[i + 1].type == P_CONV && ops[i + 1].wide) {
        POp c = ops[i + 1];
        Seg& s0 = c.conv.seg[0];
        const int C1 = g.i0, C2 = g.i1, T = g.i2, C = C1 + C2;
        bool ok = s0.A == g.out0 && s0.stride == 1 && s0.up == 0 && s0.T_in == T && c.conv.T_out == T && s0.Cin == C &&
                  (T & (T - 1)) == 0 && T <= CT && C1 % CT == 0 && C % 32 == 0;
        if (ok) {
          int n_chunks = 0;
          for (int sgi = 0; sgi < c.conv.nseg; ++sgi) n_chunks += c.conv.seg[sgi].taps * (c.conv.seg[sgi].Cin / CT);
          const int per_unit = (n_chunks + c.ks - 1) / c.ks + 1;
          const int channels = (per_unit / s0.taps + 2) * CT;
          const int cg = C / 32;
          const int ng_max = channels / cg + 2, nb = CT / T;
          ok = channels <= 512 && nb * ng_max <= 256 && (T * cg + 31) / 32 <= 32;
          if (ok) {
            int logT = 0;
            while ((1 << logT) < T) ++logT;
            s0.A = g.in0; s0.A2 = g.in1; s0.C1 = C1;
            c.wide = 1;   // the fused normalisation lives in the mma.sync unit
            c.conv.pn_on = 1; c.conv.pn_silu = g.i3; c.conv.pn_cg = cg; c.conv.pn_logT = logT; c.conv.pn_inv_cg = 1.0f / (float)cg;
            c.conv.pn_gamma = g.w0; c.conv.pn_beta = g.w1;
            if (g.out1) raws.push_back(RawSrc{g.out1, g.in0, g.in1, C1});
            if (c.conv.nseg > 1) {
              for (const auto& rsrc : raws)
                if (c.conv.seg[1].A == rsrc.raw) { c.conv.seg[1].A = rsrc.in1; c.conv.seg[1].A2 = rsrc.in2; c.conv.seg[1].C1 = rsrc.C1; }
            }
            fused.push_back(c);
            ++i;
            continue;
          }
        }
      }
      fused.push_back(g);
    }
    ops.swap(fused);
  }
  // deferred K-slice exchange: a wide token GEMM whose output is consumed first by the very next op (GroupNorm of it, or the
  // attention that follows a qkv projection) leaves its partial tiles to that op
  if (wide && u->persist_defer && !(dbg & 2)) {
    for (size_t i = 0; i + 1 < ops.size(); ++i) {
      POp& c = ops[i];
      POp& nx = ops[i + 1];
      if (c.type != P_CONV || !c.wide || c.ks <= 1) continue;
      // (attention consumers were measured slower with the slice sum in their load phase than the qkv GEMM's own exchange;
      // SURFD_UNET_DEBUG bit 6 enables it for experiments)
      const bool attn_ok = (dbg & 64) && nx.type == P_ATTN;
      if (!((nx.type == P_GN || attn_ok) && nx.in0 == c.conv.out)) continue;
      if (nx.type == P_GN && nx.i0 != c.conv.N) continue;
      if (nx.type == P_ATTN && (c.conv.emb || c.conv.residual || 3 * nx.i0 != c.conv.N)) continue;
      c.defer = 1;
      nx.ps.part = ln.p_partials.as<float>();   // patched below once the scratch buffer has its final address
      nx.ps.bias = c.conv.bias; nx.ps.emb = c.conv.emb; nx.ps.res = c.conv.residual;
      nx.ps.ks = c.ks; nx.ps.tiles_n = c.tiles_n; nx.ps.emb_ld = c.conv.emb_ld; nx.ps.N = c.conv.N;
    }
  }
  // weight stream (tcgen05 units): packed 32 KB images per (op, output tile, K slice, chunk pair) + the job table
  ln.p_use_tma = 0;
  if (use_tc && !(dbg & (4 | 128))) {
    std::vector<WJob> jobs;
    std::vector<size_t> offs;
    size_t total = 0;
    bool all_tc = true;
    for (const auto& o : ops) {
      if (o.type != P_CONV) continue;
      if (o.wide != 2) { all_tc = false; break; }
      const ConvArgs& a = o.conv;
      const int n_chunks0 = a.seg[0].taps * (a.seg[0].Cin / CT);
      const int n_chunks = n_chunks0 + (a.nseg > 1 ? a.seg[1].taps * (a.seg[1].Cin / CT) : 0);
      int pairs_max = 0;
      for (int c = 0; c < o.ks; ++c) {
        const int fb = (c * n_chunks) / o.ks, fe = ((c + 1) * n_chunks) / o.ks;
        pairs_max = std::max(pairs_max, (fe - fb + 1) >> 1);
      }
      WJob j{};
      j.tiles_n = o.tiles_n; j.tiles_m = o.tiles_m; j.ks = o.ks; j.n_chunks = n_chunks; j.pairs_max = pairs_max;
      offs.push_back(total);
      total += (size_t)o.tiles_n * o.ks * pairs_max * TCX_W_BYTES;
      jobs.push_back(j);
    }
    if (all_tc && !jobs.empty()) {
      SURFD_TRY(ln.p_wstream.reserve(total));
      size_t ji = 0;
      for (const auto& o : ops) {
        if (o.type != P_CONV) continue;
        WJob& j = jobs[ji];
        j.base = ln.p_wstream.as<uint8_t>() + offs[ji];
        const ConvArgs& a = o.conv;
        const int n_chunks0 = a.seg[0].taps * (a.seg[0].Cin / CT);
        pack_wstream_kernel<<<dim3((unsigned)(o.tiles_n * o.ks), (unsigned)j.pairs_max), 256>>>(a, o.ks, n_chunks0, j.n_chunks, j.pairs_max,
                                                                                                  ln.p_wstream.as<uint8_t>() + offs[ji]);
        SURFD_CHECK_LAUNCH();
        ++ji;
      }
      SURFD_CUDA(cudaDeviceSynchronize());
      // per-CTA image lists of one pass, in the order p_conv_tc consumes them: ops in program order, units u = cta, cta + grid,
      // ..., pairs in order
      std::vector<const uint8_t*> list;
      std::vector<int> idx(2 * (size_t)grid);
      for (int c = 0; c < grid; ++c) {
        idx[(size_t)c] = (int)list.size();
        for (const WJob& j : jobs) {
          const int n_units = j.tiles_n * j.tiles_m * j.ks;
          for (int uu = c; uu < n_units; uu += grid) {
            const int crank = uu % j.ks, tn = (uu / j.ks) % j.tiles_n;
            const int fb = (crank * j.n_chunks) / j.ks, fe = ((crank + 1) * j.n_chunks) / j.ks;
            const int np = (fe - fb + 1) >> 1;
            for (int pp = 0; pp < np; ++pp) list.push_back(j.base + ((size_t)(tn * j.ks + crank) * j.pairs_max + pp) * TCX_W_BYTES);
          }
        }
        idx[(size_t)grid + c] = (int)list.size() - idx[(size_t)c];
      }
      if (list.empty()) list.push_back(nullptr);
      SURFD_TRY(ln.p_wlist.reserve(list.size() * sizeof(void*)));
      SURFD_TRY(ln.p_wlidx.reserve(idx.size() * sizeof(int)));
      SURFD_CUDA(cudaMemcpy(ln.p_wlist.p, list.data(), list.size() * sizeof(void*), cudaMemcpyHostToDevice));
      SURFD_CUDA(cudaMemcpy(ln.p_wlidx.p, idx.data(), idx.size() * sizeof(int), cudaMemcpyHostToDevice));
      ln.p_use_tma = 1;
    }
  }
  // point-to-point hand-off GEMM -> GroupNorm (weight-stream flow): per-tile arrival counters, one array per deferred GEMM
  // (monotonic over a launch: a tile has received ks * executions arrivals), zeroed before every launch
  ln.p_p2p_words = 0;
  if (ln.p_use_tma && !(dbg & 256)) {
    size_t words = 0;
    for (size_t i = 0; i + 1 < ops.size(); ++i)
      if (ops[i].type == P_CONV && ops[i].defer && ops[i].wide == 2 && ops[i + 1].type == P_GN) words += (size_t)ops[i].tiles_n * ops[i].tiles_m;
    SURFD_TRY(ln.p_p2p.reserve((words ? words : 1) * sizeof(unsigned)));
    size_t off = 0;
    for (size_t i = 0; i + 1 < ops.size(); ++i) {
      if (ops[i].type == P_CONV && ops[i].defer && ops[i].wide == 2 && ops[i + 1].type == P_GN) {
        ops[i].p2p = ln.p_p2p.as<unsigned>() + off;
        ops[i + 1].ps.p2p = ops[i].p2p;
        off += (size_t)ops[i].tiles_n * ops[i].tiles_m;
      }
    }
    ln.p_p2p_words = words;
  }
  if (dbg & 4) for (auto& o : ops) o.type = 0;
  SURFD_TRY(ln.p_partials.reserve((max_partial_tiles ? max_partial_tiles : 1) * CT * CT * sizeof(float)));
  for (auto& o : ops) if (o.ps.ks > 0) o.ps.part = ln.p_partials.as<float>();
  SURFD_TRY(ln.p_ops.reserve(ops.size() * sizeof(POp)));
  SURFD_CUDA(cudaMemcpy(ln.p_ops.p, ops.data(), ops.size() * sizeof(POp), cudaMemcpyHostToDevice));
  max_tiles = (max_tiles + 63) / 64 * 64;
  SURFD_TRY(ln.p_sems.reserve(3 * max_tiles * sizeof(unsigned)));
  SURFD_TRY(ln.p_sync.reserve(P_SYNC_WORDS * sizeof(unsigned)));
  SURFD_TRY(ln.p_prof.reserve(64 * sizeof(long long)));
  ln.p_sem_bank = (int)max_tiles;
  ln.p_n_emb = n_emb; ln.p_n_prog = (int)ops.size() - n_emb; ln.p_smem = smem_floats * (int)sizeof(float);
  ln.p_wide = (int)wide + (int)use_tc;
  ln.p_use_tc = (int)use_tc;
  ln.p_B = B; ln.p_ctx = ctx; ln.p_lab = lab; ln.p_grid = grid; ln.p_split = u->persist_split;
  return 0;
}

static int sample_persistent(surfd_unet* u, int B, int n_steps, const int64_t* tmap_dev, const float* coef_dev, const float* noise_dev,
                             const float* ctx, const int64_t* lab, float guidance, float* out_dev, cudaStream_t st) {
  Lane& ln = u->lanes[0];
  SURFD_REQUIRE(B <= ln.cap, "lane capacity exceeded");
  SURFD_TRY(surfd_unet_status(u));
  const int L = u->L;
  int grid = u->sampler_sms > 0 ? u->sampler_sms : u->num_sms;
  if (grid > u->num_sms) grid = u->num_sms;     // one CTA per SM (234 registers x 256 threads)
  int pair_max = 64;   // rows per launch in the verified range of the engine (see surfd_sample)
  if (const char* e = getenv("SURFD_PERSIST_MAX_BATCH")) pair_max = atoi(e);
  // Classifier-free guidance (two forwards per step, models/cfg_sampler.py:19-26): the one-round engine runs the pair as ONE
  // pass over 2B rows -- the same FLOPs, but the 553 MB of weights stream once per step instead of twice and the op chain is
  // walked once (C5: 2.99 -> 1.5 s per 1000 steps at B = 4).  The graph-identical engine (persist_split 0) keeps two passes.
  const bool pair = guidance != 1.0f && u->persist_split == 1 && 2 * B <= ln.cap && 2 * B <= pair_max;
  const int BB = pair ? 2 * B : B;
  if (pair) {
    const size_t xb = (size_t)B * L * sizeof(float);
    SURFD_CUDA(cudaMemcpyAsync(ln.xcur.p, noise_dev, xb, cudaMemcpyDeviceToDevice, st));   // x_T = noise row 0, for both rows of a pair
    SURFD_CUDA(cudaMemcpyAsync(ln.xcur.as<char>() + xb, noise_dev, xb, cudaMemcpyDeviceToDevice, st));
    if (ctx) {
      const size_t cb = (size_t)B * CTX * sizeof(float);
      SURFD_TRY(ln.ctx2.reserve(2 * cb));
      SURFD_CUDA(cudaMemcpyAsync(ln.ctx2.p, ctx, cb, cudaMemcpyDeviceToDevice, st));
      SURFD_CUDA(cudaMemcpyAsync(ln.ctx2.as<char>() + cb, ctx, cb, cudaMemcpyDeviceToDevice, st));
      ctx = ln.ctx2.as<float>();
    }
    if (lab) {
      const size_t lb = (size_t)B * sizeof(int64_t);
      SURFD_TRY(ln.lab2.reserve(2 * lb));
      SURFD_CUDA(cudaMemcpyAsync(ln.lab2.p, lab, lb, cudaMemcpyDeviceToDevice, st));
      SURFD_CUDA(cudaMemcpyAsync(ln.lab2.as<char>() + lb, lab, lb, cudaMemcpyDeviceToDevice, st));
      lab = ln.lab2.as<int64_t>();
    }
  } else {
    SURFD_CUDA(cudaMemcpyAsync(ln.xcur.p, noise_dev, (size_t)B * L * sizeof(float), cudaMemcpyDeviceToDevice, st));   // x_T = noise row 0
  }
  SURFD_TRY(persist_build(u, ln, BB, ctx, lab, grid));
  int max_grid = 0;
  if (u->precision == 0) SURFD_TRY(persist_max_grid<0>(ln.p_smem, u->num_sms, &max_grid));
  else if (u->precision == 1) SURFD_TRY(persist_max_grid<1>(ln.p_smem, u->num_sms, &max_grid));
  else SURFD_TRY(persist_max_grid<2>(ln.p_smem, u->num_sms, &max_grid));
  SURFD_REQUIRE(max_grid >= grid, "persistent sampler kernel: the requested CTAs cannot all be resident");
  SURFD_CUDA(cudaMemsetAsync(ln.p_sync.p, 0, P_SYNC_WORDS * sizeof(unsigned), st));
  SURFD_CUDA(cudaMemsetAsync(ln.p_sems.p, 0, 3 * (size_t)ln.p_sem_bank * sizeof(unsigned), st));
  if (ln.p_p2p_words) SURFD_CUDA(cudaMemsetAsync(ln.p_p2p.p, 0, ln.p_p2p_words * sizeof(unsigned), st));
  if (ctx) {   // projected context: constant over the loop, computed once (the graph path accumulates it every step)
    const dim3 g1((unsigned)cdiv(EMB, 8), (unsigned)cdiv(BB, 8));
    const auto& h = u->hdr;
    UNET_LAUNCH(false, linear_rows_kernel, dim3(g1), dim3(256), 0, st, 0, ctx, BB, CTX, u->w(h[9]), u->w(h[10]), EMB, 0, 0, 0, ln.ctxv.as<float>());
  }
  PersistArgs pa{};
  pa.ops = ln.p_ops.as<POp>(); pa.n_emb = ln.p_n_emb; pa.n_prog = ln.p_n_prog;
  pa.n_steps = n_steps; pa.B = BB; pa.L = L; pa.n_pass = (guidance != 1.0f && !pair) ? 2 : 1;
  pa.pair_B = pair ? B : 0;
  pa.tmap = tmap_dev; pa.coef = coef_dev; pa.noise = noise_dev; pa.noise_stride = (long long)B * L; pa.guidance = guidance;
  pa.x = ln.xcur.as<float>(); pa.x0a = ln.x0a.as<float>();
  pa.partials = ln.p_partials.as<float>(); pa.sems = ln.p_sems.as<unsigned>(); pa.sem_bank = ln.p_sem_bank; pa.use_tc = ln.p_use_tc; pa.sync = ln.p_sync.as<unsigned>();
  pa.wlist = ln.p_wlist.as<const uint8_t*>(); pa.wl_start = ln.p_wlidx.as<int>(); pa.wl_count = ln.p_wlidx.as<int>() + grid;
  pa.use_tma = ln.p_use_tma;
  pa.prof = nullptr;
  if (u->profile) {
    SURFD_CUDA(cudaMemsetAsync(ln.p_prof.p, 0, 64 * sizeof(long long), st));
    pa.prof = ln.p_prof.as<long long>();
  }
  if (u->precision == 0) SURFD_TRY(persist_launch<0>(pa, grid, ln.p_smem, st));
  else if (u->precision == 1) SURFD_TRY(persist_launch<1>(pa, grid, ln.p_smem, st));
  else SURFD_TRY(persist_launch<2>(pa, grid, ln.p_smem, st));
  SURFD_CUDA(cudaMemcpyAsync(u->h_abort, ln.p_sync.as<unsigned>() + 1, 2 * sizeof(unsigned), cudaMemcpyDeviceToHost, st));   // abort, range flags
  SURFD_CUDA(cudaMemcpyAsync(out_dev, ln.xcur.p, (size_t)B * L * sizeof(float), cudaMemcpyDeviceToDevice, st));
  return 0;
}

extern "C" int surfd_sample(surfd_unet* u, int B, int n_steps, const int64_t* tmap_dev, const float* coef_dev, const float* noise_dev,
                            const float* context_dev, const int64_t* labels_dev, float guidance, float* out_dev, void* stream) {
  SURFD_REQUIRE(u && tmap_dev && coef_dev && noise_dev && out_dev, "null argument");
  SURFD_REQUIRE(B >= 1 && B <= u->max_batch, "batch exceeds max_batch");
  SURFD_REQUIRE(n_steps >= 1, "n_steps must be positive");
  cudaStream_t st = (cudaStream_t)stream;
  // The persistent engine is used for batches of at most PERSIST_MAX_BATCH samples per call and (wide units) at least
  // PERSIST_MIN_GRID CTAs.  Above 8 samples a token GEMM has more units than resident CTAs (several rounds per op): verified on
  // hardware up to 64 samples per call, both latent sizes, against the graph engine (tests/test_gpu_unet.py large_batches,
  // tools/probe_large_batch.py; the round-1 report of an unfinished 12 / 40-sample run did not reproduce in three separate
  // sessions, nor under compute-sanitizer -- profiles/r2_sanitizer.md).  Larger batches take the CUDA-graph engine.
  constexpr int PERSIST_MIN_GRID = 100;
  int PERSIST_MAX_BATCH = 64;
  if (const char* e = getenv("SURFD_PERSIST_MAX_BATCH")) PERSIST_MAX_BATCH = atoi(e);   // diagnostics
  if (u->sampler == 1 && u->coop && B <= PERSIST_MAX_BATCH) {
    int grid = u->sampler_sms > 0 ? u->sampler_sms : u->num_sms;
    if (grid > u->num_sms) grid = u->num_sms;
    if (u->persist_split == 0 || grid >= PERSIST_MIN_GRID)
      return sample_persistent(u, B, n_steps, tmap_dev, coef_dev, noise_dev, context_dev, labels_dev, guidance, out_dev, st);
  }
  const int L = u->L;
  const int n_lanes = (int)u->lanes.size() < B ? (int)u->lanes.size() : B;
  const int64_t stride = (int64_t)B * L;
  SURFD_CUDA(cudaEventRecord(u->fork, st));
  int off = 0;
  for (int li = 0; li < n_lanes; ++li) {
    Lane& ln = u->lanes[(size_t)li];
    const int b = B / n_lanes + (li < B % n_lanes ? 1 : 0);
    SURFD_REQUIRE(b <= ln.cap, "lane capacity exceeded");
    const float* noise = noise_dev + (size_t)off * L;
    const float* ctx = context_dev ? context_dev + (size_t)off * CTX : nullptr;
    const int64_t* lab = labels_dev ? labels_dev + off : nullptr;
    SURFD_CUDA(cudaStreamWaitEvent(ln.stream, u->fork, 0));
    SURFD_CUDA(cudaMemcpyAsync(ln.xcur.p, noise, (size_t)b * L * sizeof(float), cudaMemcpyDeviceToDevice, ln.stream));   // x_T = noise row 0
    StepState init{0, n_steps};
    SURFD_CUDA(cudaMemcpyAsync(ln.state.p, &init, sizeof(init), cudaMemcpyHostToDevice, ln.stream));
    const bool reuse = ln.graph_exec && ln.graph_B == b && ln.graph_ctx == ctx && ln.graph_lab == lab && ln.graph_tmap == tmap_dev &&
                       ln.graph_coef == coef_dev && ln.graph_noise == noise && ln.graph_stride == stride && ln.graph_guidance == guidance;
    if (!reuse) {
      if (ln.graph_exec) { cudaGraphExecDestroy(ln.graph_exec); ln.graph_exec = nullptr; }
      for (int attempt = 0; attempt < 2; ++attempt) {
        cudaStream_t cs;
        SURFD_CUDA(cudaStreamCreateWithFlags(&cs, cudaStreamNonBlocking));
        cudaGraph_t graph = nullptr;
        cudaError_t ce = cudaStreamBeginCapture(cs, cudaStreamCaptureModeThreadLocal);
        int rc = 0;
        if (ce == cudaSuccess) {
          const int64_t launches_before = g_launch_count;
          rc = record_step(u, ln, b, tmap_dev, coef_dev, noise, stride, ctx, lab, guidance, cs);
          ln.graph_launches = g_launch_count - launches_before;
          g_launch_count = launches_before;   // capture is not execution; launches are counted per replay below
          cudaError_t ce2 = cudaStreamEndCapture(cs, &graph);
          if (ce == cudaSuccess) ce = ce2;
        }
        if (ce == cudaSuccess && rc == 0) ce = cudaGraphInstantiate(&ln.graph_exec, graph, 0);
        if (graph) cudaGraphDestroy(graph);
        cudaStreamDestroy(cs);
        if (ce == cudaSuccess && rc == 0) break;
        ln.graph_exec = nullptr;
        cudaGetLastError();   // clear
        if (u->pdl && attempt == 0) { u->pdl = false; continue; }   // programmatic edges rejected: capture again without them
        if (rc) return rc;
        return set_error(-(int)ce, cudaGetErrorString(ce), __FILE__, __LINE__);
      }
      ln.graph_B = b; ln.graph_ctx = ctx; ln.graph_lab = lab; ln.graph_tmap = tmap_dev; ln.graph_coef = coef_dev;
      ln.graph_noise = noise; ln.graph_stride = stride; ln.graph_guidance = guidance;
    }
    off += b;
  }
  // interleave the lanes' graph launches step by step so they progress together (and share weight lines in L2)
  for (int i = 0; i < n_steps; ++i) {
    for (int li = 0; li < n_lanes; ++li) {
      Lane& ln = u->lanes[(size_t)li];
      SURFD_CUDA(cudaGraphLaunch(ln.graph_exec, ln.stream));
      g_launch_count += ln.graph_launches;
    }
  }
  off = 0;
  for (int li = 0; li < n_lanes; ++li) {
    Lane& ln = u->lanes[(size_t)li];
    const int b = B / n_lanes + (li < B % n_lanes ? 1 : 0);
    SURFD_CUDA(cudaMemcpyAsync(out_dev + (size_t)off * L, ln.xcur.p, (size_t)b * L * sizeof(float), cudaMemcpyDeviceToDevice, ln.stream));
    SURFD_CUDA(cudaEventRecord(ln.done, ln.stream));
    SURFD_CUDA(cudaStreamWaitEvent(st, ln.done, 0));
    off += b;
  }
  return 0;
}
