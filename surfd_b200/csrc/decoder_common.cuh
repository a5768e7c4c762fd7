// decoder_common.cuh -- epilogue contract shared by the FFMA (decoder.cu) and tcgen05 (decoder_tc.cu) layer GEMMs.
#pragma once
#include "common.cuh"

namespace surfd {

// C[m][n] = sum_k A[m][k] W[n][k];  then, in this order:
//   v = acc + bias[n]
//   v = mask[m][n] > 0 ? v * mscale[n] : 0        (CBN/ReLU backward)
//   v += R[m][n]                                   (residual; may alias C)
//   C[m][n] = v                                    (optionally rounded to TF32 when it feeds a tensor-core GEMM)
//   act[m][n] = relu(s2[n] * v + t2[n])            (next layer's conditional batch-norm + ReLU)
struct Epilogue {
  const float* bias;
  const float* mask;
  const float* mscale;
  const float* R;
  float* C;
  float* act;
  const float* s2;
  const float* t2;
  int ld;
  int round_c;
  int round_act;
};

__device__ __forceinline__ float round_to_tf32(float x) {
  uint32_t u;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(x));
  return __uint_as_float(u);
}

// decoder_tc.cu: A [M][512], W [512][512] (both K contiguous), persistent tcgen05 kernel
int launch_gemm_tc(const float* A, const float* W, int M, const Epilogue& e, int* err_flag, int num_sms, cudaStream_t st);

// decoder_tc.cu: n consecutive ROW-LOCAL layers over the same rows in ONE launch (each CTA runs all layers on its own 128-row panels, so no CTA
// depends on another one).  gsync: unused (kept for the signature).
int launch_chain_tc(const float* const* A, const float* const* W, const Epilogue* e, int n, int M, unsigned* gsync, int* err_flag,
                    int num_sms, cudaStream_t st);

}  // namespace surfd
