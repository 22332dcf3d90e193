// Shared definitions of the fused attention kernels for head_dim 64 on tcgen05, and the side kernel that
// materialises probabilities for `output_attentions=True`.
//
//   O = softmax(Q K^T / sqrt(64) + key_bias) V      per (batch, head), never materialising P in HBM
//   (bert_model.py:309 QK^T, :328 /sqrt(d), :329-331 +mask, :334 softmax, :343 P V, :345-347 merge heads)
//
// The forward kernel is attn_fwd3.cuh (persistent, third generation).  The first generation that used to live here (one CTA per
// (batch, head, 256-query pair), 93 us at the bench shape against 68 us) was removed in round 2; its A/B data is in
// profiles/r01d-r01f and git history keeps the code (01a2ca4).
// Q, K, V are read in place from the packed [tokens, 3H] projection output through TMA coordinates (no head
// split / permute kernels) and the context is written straight into [tokens, H].
#pragma once
#include "ptx.cuh"

namespace b200 {

constexpr int ATT_D = 64;
constexpr int ATT_BQ = 128;
constexpr int ATT_BK = 128;
constexpr float kAttScaleLog2 = 1.4426950408889634f / 8.0f;      // log2(e) / sqrt(ATT_D): scores -> exp2 exponents

struct AttnFwdArgs {
  int B, heads, Sq, Sk;
  int q_col0, k_col0, v_col0;        // column offsets of head 0 inside the Q map / KV map
  const float* key_bias;             // [B, Sk] additive bias in natural-log units (0 / -inf or finite), may be null
  const int* kv_len;                 // [B] keys beyond kv_len[b] are fully masked (block skipping), may be null
  __half* out;                       // [B*Sq, ld_out], head h at columns h*64
  int ld_out;
  float* lse2;                       // [B, heads, Sq] log2-domain log-sum-exp (for backward), may be null
  float scale_log2;                  // log2(e) / sqrt(d)
  DropCfg drop;                      // attention-probability dropout (bert_model.py:338), off when seed_base is null
  const int* cu_seqlens;             // [B+1] or null.  Non-null = PACKED rows: sequence b occupies rows
                                     // [cu[b], cu[b+1]) of the Q / KV / output buffers, its length is both its query and its key
                                     // count, and Sq / Sk are the MAXIMUM length (they still shape lse2 / the dropout indices, so a
                                     // packed run draws the same masks as the padded one).  Self-attention only.
};

// Materialised attention probabilities for `output_attentions=True` (ditto/evaluation_ditto.py:121-127 reads the
// diagonal of one head).  P[b,h,i,j] = exp2(q_i.k_j * scale_log2 + bias_j*log2e - lse2[b,h,i]).  Plain SIMT: this side
// output is only requested by the 128-token inference path and is HBM-write bound ([B,h,S,S] fp32).
__global__ void attn_probs_kernel(const __half* __restrict__ q, int ldq, const __half* __restrict__ k, int ldk,
                                  const float* __restrict__ key_bias, const float* __restrict__ lse2, float* __restrict__ probs,
                                  int B, int heads, int Sq, int Sk, float scale_log2) {
  __shared__ __half qs[32][ATT_D + 8];
  __shared__ __half ks[32][ATT_D + 8];
  const int b = blockIdx.z / heads, h = blockIdx.z % heads;
  const int i0 = blockIdx.y * 32, j0 = blockIdx.x * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;          // 32 x 8 threads
  for (int rr = ty; rr < 32; rr += 8) {
    for (int c = tx; c < ATT_D; c += 32) {
      qs[rr][c] = (i0 + rr < Sq) ? q[(static_cast<size_t>(b) * Sq + i0 + rr) * ldq + h * ATT_D + c] : __float2half(0.f);
      ks[rr][c] = (j0 + rr < Sk) ? k[(static_cast<size_t>(b) * Sk + j0 + rr) * ldk + h * ATT_D + c] : __float2half(0.f);
    }
  }
  __syncthreads();
  for (int rr = ty; rr < 32; rr += 8) {
    const int i = i0 + rr, j = j0 + tx;
    if (i >= Sq || j >= Sk) continue;
    float s = 0.f;
#pragma unroll 16
    for (int c = 0; c < ATT_D; ++c) s = fmaf(__half2float(qs[rr][c]), __half2float(ks[tx][c]), s);
    const float bias = key_bias ? key_bias[static_cast<size_t>(b) * Sk + j] * 1.4426950408889634f : 0.f;
    const float L = lse2[(static_cast<size_t>(b) * heads + h) * Sq + i];
    probs[((static_cast<size_t>(b) * heads + h) * Sq + i) * Sk + j] = exp2f(fmaf(s, scale_log2, bias) - L);
  }
}

}  // namespace b200
