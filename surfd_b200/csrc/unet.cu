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
  float s = 0.f;
  for (int i = threadIdx.x; i < n; i += 128) s += load(i);
  const float mean = block_sum(s) / (float)n;
  float q = 0.f;
  for (int i = threadIdx.x; i < n; i += 128) { const float d = load(i) - mean; q += d * d; }
  const float var = block_sum(q) / (float)n;
  const float rstd = 1.0f / sqrtf(var + 1e-5f);
  for (int i = threadIdx.x; i < n; i += 128) {
    const int t = i / cg, c = g * cg + i % cg;
    const float x = load(i);
    float y = (x - mean) * rstd * gamma[c] + beta[c];
    if (silu) y = y / (1.0f + expf(-y));
    const size_t o = ((size_t)b * T + t) * C + c;
    out[o] = y;
    if (raw) raw[o] = x;
  }
}

// ------------------------------------------------------------------------------------------------
// Token GEMM: out[m][n] = sum over K segments/taps of A[token(m,tap)][ci] * W[tap][n][ci]  (+bias +emb +residual)
// CTA tile 32 tokens x 32 outputs; 8 warps split the K chunks (32 channels each) and reduce through smem.
// ------------------------------------------------------------------------------------------------
struct Seg {
  const float* A;  // [B][T_in][Cin]
  const float* W;  // [taps][N][Cin]
  int Cin, taps, stride, up, T_in;
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
};

constexpr int CT = 32;        // tile edge
constexpr int CTP = CT + 4;   // padded row (floats)
constexpr int KSPLIT = 8;     // max CTAs per cluster: the K reduction is split over a thread-block cluster (1, 2, 4 or 8)

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
__device__ __forceinline__ void conv_accumulate(const ConvArgs& a, int n0, int m0, int ks, int crank, float* smem, float (&acc)[8][4]) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* As = smem + warp * (2 * CT * CTP);
  float* Ws = As + CT * CTP;
  const int M = a.B * a.T_out;

  // this lane's token row for loading
  const int m_row = m0 + lane;
  const int rb = m_row / a.T_out, rl = m_row % a.T_out;
  const bool row_ok = m_row < M;

  const int ly = lane >> 3, lx = lane & 7;   // FFMA mapping: rows ly + 4i, cols lx + 8j
  const int fg = lane >> 2, ft = lane & 3;   // MMA fragment mapping: group id / thread in group
  // acc: MODE 0: [i][j];  MODE 1/2: [mt*4 + nt][c0..c3]
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  const int my_slot = warp * ks + crank;       // chunk c belongs to slot c % (8*ks); consecutive chunks go to different CTAs
  // This warp owns the flattened chunks f = my_slot, my_slot + 8*ks, ...  (f enumerates segment, tap, 32-channel block).
  // The weight rows of the NEXT chunk are prefetched into registers while the current one is multiplied: weights come
  // from HBM (553 MB per step >> L2) and are the long pole of the per-warp chain; the activation rows are L2-resident.
  const int n_chunks0 = a.seg[0].taps * (a.seg[0].Cin / CT);
  const int n_chunks = n_chunks0 + (a.nseg > 1 ? a.seg[1].taps * (a.seg[1].Cin / CT) : 0);
  const int stride_f = 8 * ks;
  auto locate = [&](int f, const float*& arow, const float*& wrow, bool& ok) {
    const int s = f < n_chunks0 ? 0 : 1;
    const Seg& sg = a.seg[s];
    const int g = s ? f - n_chunks0 : f;
    const int cpt = sg.Cin / CT;
    const int tap = g / cpt, c = g - tap * cpt;
    const int T_eff = sg.up ? 2 * sg.T_in : sg.T_in;
    const int src = rl * sg.stride + tap - (sg.taps >> 1);
    ok = row_ok && src >= 0 && src < T_eff;
    const int st = sg.up ? (src >> 1) : src;
    arow = sg.A + ((size_t)rb * sg.T_in + (ok ? st : 0)) * sg.Cin + c * CT;
    wrow = sg.W + ((size_t)tap * a.N + n0 + lane) * sg.Cin + c * CT;
  };
  float4 wv[8], wn[8];
  const float* arow = nullptr; const float* wrow = nullptr; bool ok = false;
  int f = my_slot;
  if (f < n_chunks) {
    locate(f, arow, wrow, ok);
#pragma unroll
    for (int j = 0; j < 8; ++j) wv[j] = *reinterpret_cast<const float4*>(wrow + 4 * j);
  }
  for (; f < n_chunks; f += stride_f) {
    {
      {
        float4 av[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) av[j] = ok ? *reinterpret_cast<const float4*>(arow + 4 * j) : make_float4(0.f, 0.f, 0.f, 0.f);
        const int fn = f + stride_f;
        if (fn < n_chunks) {
          locate(fn, arow, wrow, ok);
#pragma unroll
          for (int j = 0; j < 8; ++j) wn[j] = *reinterpret_cast<const float4*>(wrow + 4 * j);
        }
        __syncwarp();
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          *reinterpret_cast<float4*>(As + lane * CTP + 4 * j) = av[j];
          *reinterpret_cast<float4*>(Ws + lane * CTP + 4 * j) = wv[j];
        }
        __syncwarp();
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
#pragma unroll
        for (int j = 0; j < 8; ++j) wv[j] = wn[j];
      }
    }
  }
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

// epilogue of one output element: + bias (+ per-sample embedding column) (+ residual)
__device__ __forceinline__ void conv_store(const ConvArgs& a, int m0, int n0, int idx, float v) {
  const int r = idx >> 5, col = idx & 31;
  const int m = m0 + r;
  if (m < a.B * a.T_out) {
    const int n = n0 + col;
    const int b = m / a.T_out;
    v += a.bias[n];
    if (a.emb) v += a.emb[(size_t)b * a.emb_ld + n];
    const size_t o = (size_t)m * a.N + n;
    if (a.residual) v += a.residual[o];
    a.out[o] = v;
  }
}

template <int MODE>
__global__ void __launch_bounds__(256)
conv_gemm_kernel(ConvArgs a) {
  PDL_PROLOGUE();
  extern __shared__ __align__(16) float smem[];
  cg::cluster_group cluster = cg::this_cluster();
  const int crank = (int)cluster.block_rank();
  const int ks = (int)cluster.num_blocks();    // K split of this launch: 1, 2, 4 or 8 CTAs per output tile
  const int n0 = blockIdx.x * CT;
  const int m0 = blockIdx.y * CT;
  float acc[8][4];
  conv_accumulate<MODE>(a, n0, m0, ks, crank, smem, acc);
  float* part = conv_cta_partial<MODE>(acc, smem);
  // (2) cross-CTA reduction over distributed shared memory: CTA `crank` finishes its 1/ks share of the 32x32 tile
  cluster.sync();
  {
    const int per = CT * CT / ks;
    for (int idx = crank * per + threadIdx.x; idx < (crank + 1) * per; idx += 256) {
      float v = 0.f;
      for (int j = 0; j < ks; ++j) v += cluster.map_shared_rank(part, j)[idx];
      conv_store(a, m0, n0, idx, v);
    }
  }
  cluster.sync();   // keep this CTA's shared memory alive until every peer has read it
}

// ------------------------------------------------------------------------------------------------
// QKVAttentionLegacy: one CTA per (batch, head).  qkv [B][T][3C] with per-head [q|k|v] channel interleave.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
attn_kernel(const float* __restrict__ qkv, int C, int T, int heads, float scale, float* __restrict__ out) {
  PDL_PROLOGUE();
  extern __shared__ __align__(16) float sm[];
  const int ch = C / heads;
  const int b = blockIdx.x / heads, h = blockIdx.x % heads;
  float* q = sm;               // [T][ch] (scaled)
  float* k = q + T * ch;       // [T][ch] (scaled)
  float* v = k + T * ch;       // [T][ch]
  float* w = v + T * ch;       // [T][T+1]
  const float* base = qkv + (size_t)b * T * 3 * C + (size_t)h * 3 * ch;
  for (int i = threadIdx.x; i < T * ch; i += 128) {
    const int t = i / ch, c = i % ch;
    const float* p = base + (size_t)t * 3 * C + c;
    q[i] = p[0] * scale;
    k[i] = p[ch] * scale;
    v[i] = p[2 * ch];
  }
  __syncthreads();
  for (int i = threadIdx.x; i < T * T; i += 128) {
    const int t = i / T, s = i % T;
    float acc = 0.f;
    for (int c = 0; c < ch; ++c) acc = fmaf(q[t * ch + c], k[s * ch + c], acc);
    w[t * (T + 1) + s] = acc;
  }
  __syncthreads();
  for (int t = threadIdx.x; t < T; t += 128) {
    float mx = -INFINITY;
    for (int s = 0; s < T; ++s) mx = fmaxf(mx, w[t * (T + 1) + s]);
    float sum = 0.f;
    for (int s = 0; s < T; ++s) { const float e = expf(w[t * (T + 1) + s] - mx); w[t * (T + 1) + s] = e; sum += e; }
    const float inv = 1.0f / sum;
    for (int s = 0; s < T; ++s) w[t * (T + 1) + s] *= inv;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < T * ch; i += 128) {
    const int t = i / ch, c = i % ch;
    float acc = 0.f;
    for (int s = 0; s < T; ++s) acc = fmaf(w[t * (T + 1) + s], v[s * ch + c], acc);
    out[((size_t)b * T + t) * C + h * ch + c] = acc;
  }
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

// per-warp A/W staging (8 warps x 2 x 32 x 36 floats) is reused for the 8 x 32 x 33 warp partials; + the 32x32 CTA partial
static constexpr int CONV_SMEM = (8 * 2 * CT * CTP + CT * CT) * (int)sizeof(float);

// One lane = private activation pool + step state + captured step graph + stream.  The denoiser is latency-bound (about 170
// dependent small kernels per step on <= 256 tokens), so a batch is split over lanes that run concurrently on their own
// streams; the weights are shared.
struct Lane {
  DevBuf pool, emb_all, temb, e1, emb, t_cur, x0a, x0b, xcur, state;
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
    if (stream) cudaStreamDestroy(stream);
    if (done) cudaEventDestroy(done);
    stream = nullptr; done = nullptr;
  }
};

struct surfd_unet {
  int L = 0, max_batch = 0;
  bool pdl = true;     // programmatic dependent launch between the step's kernels (falls back to false if capture rejects it)
  int precision = 1;   // token GEMMs: 0 fp32 FFMA, 1 3xTF32 mma.sync (fp32-class accuracy, default), 2 single-pass TF32
  DevBuf weights;
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
  const int cap0 = u->max_batch;                                   // lane 0 also serves surfd_unet_forward at full batch
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
  u->weights.release();
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
        const size_t smem = ((size_t)3 * T * ch + (size_t)T * (T + 1)) * sizeof(float);
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

extern "C" int surfd_sample(surfd_unet* u, int B, int n_steps, const int64_t* tmap_dev, const float* coef_dev, const float* noise_dev,
                            const float* context_dev, const int64_t* labels_dev, float guidance, float* out_dev, void* stream) {
  SURFD_REQUIRE(u && tmap_dev && coef_dev && noise_dev && out_dev, "null argument");
  SURFD_REQUIRE(B >= 1 && B <= u->max_batch, "batch exceeds max_batch");
  SURFD_REQUIRE(n_steps >= 1, "n_steps must be positive");
  cudaStream_t st = (cudaStream_t)stream;
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
