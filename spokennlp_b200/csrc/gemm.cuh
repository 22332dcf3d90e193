// Shared definitions of the tcgen05 GEMM family: epilogue ids and the argument block.  The kernel
// lives in gemm2.cuh (2-CTA cta_group::2 tiles).  The single-CTA first generation was removed in round 2 once the A/B data was
// in profiles/ (r01a-r01d launch lists: 1289 -> 2127 seq/s when the 2-CTA kernel replaced it); git history keeps it (01a2ca4).
//
//   C[M,N] = epilogue( A[M,K] * B[N,K]^T ),  fp16 operands, fp32 accumulation in TMEM.
//
// Operands may be K-major (contraction dim contiguous: activations x, weights W[out,in] in the forward) or
// MN-major (contraction dim strided: W in dgrad, dY^T and X^T in wgrad).  MN-major tiles are loaded as 64-wide
// column blocks ([k rows][128 B], 128B swizzle) and described to the tensor core with a_major/b_major = 1, so the
// backward needs no transposed copies of weights or activations.
#pragma once
#include "ptx.cuh"

namespace b200 {

constexpr int GEMM_BK = 64;            // K block per pipeline stage (one 128-byte swizzle row of fp16)

enum : int {
  EPI_STORE = 0,      // out = acc
  EPI_BIAS = 1,       // out = acc + bias[n]
  EPI_BIAS_GELU = 2,  // z = acc + bias[n]; out = gelu_erf(z); out2 = gelu'(z) (if given; consumed by EPI_DGELU)
  EPI_BIAS_RES = 3,   // out = acc + bias[n] + aux[m,n]
  EPI_DGELU = 4,      // out = acc * aux[m,n]   (aux = gelu'(z) saved by the forward)
  EPI_ADD = 5,        // out = acc + aux[m,n]
  EPI_ATOMIC = 6,     // out(fp32) += alpha * acc   (split-K reduction with red.global.add)
  EPI_BIAS_RES32 = 7, // out = acc + bias[n] + aux32[m,n]   (fp32 residual stream)
  // gemm2 only, opt-in:
  EPI_RESADD = 8,     // out(fp32) += drop(acc + bias[n])  — out already holds the residual; TMA reduce-add, so the epilogue reads no
                      // aux at all and a tile may be finished by several CTA pairs (stream-K schedule, GemmArgs::k_splits = -1)
  EPI_STORE_DELTA = 9,// out = acc (fp16) and delta[b,h,q] = sum_d out[row, 64h+d] * aux[row, 64h+d]  (attention-backward row statistic
                      // fused into the output-projection dgrad; aux = the forward's context, row = b*delta_S + q)
};

struct GemmArgs {
  int M, N, K;
  int k_splits;
  const float* bias;
  const __half* aux;    // fp16 [M, ld_aux]; reinterpreted as const float* by EPI_BIAS_RES32
  int ld_aux;
  void* out;
  int ld_out;
  __half* out2;
  int ld_out2;
  const float* alpha;   // optional device scalar
  DropCfg drop;         // dropout on (acc + bias) before the residual is added (EPI_BIAS_RES / EPI_BIAS_RES32, gemm2 only)
  float* colsum;        // EPI_DGELU (gemm2 only), optional: colsum[n] += *col_alpha * sum_m out[m, n] — the bias gradient of the layer
  const float* col_alpha;   // that produced the saved activation, taken from the fp16-rounded tile while it sits in the staging slab
  // EPI_STORE_DELTA keeps its row statistic in the out2 slot: out2 (as float*) = delta [B, N/64, ld_out2], ld_out2 = tokens per sequence.
};

}  // namespace b200
