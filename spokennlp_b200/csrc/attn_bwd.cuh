// Shared definitions of the fused attention backward for head_dim 64 on tcgen05 — the autograd of
// bert_model.py:309-350 — and its two streaming side kernels (row statistic, dQ cast).  The backward kernel is attn_bwd3.cuh
// (persistent, natural orientation).  The first generation that used to live here (transposed orientation, one CTA per key
// block: 222 us at the bench shape against 172 us) was removed in round 2; profiles/r01d-r01f hold its A/B data, git history
// the code (01a2ca4).
#pragma once
#include "attn_fwd.cuh"

namespace b200 {

constexpr int ATTB_THREADS = 320;   // warp 0 TMA, warp 1 MMA, warps 2-9 softmax / gradient epilogue

struct AttnBwdArgs {
  int B, heads, Sq, Sk;
  int q_col0, k_col0, v_col0;       // head-0 columns inside the Q map / KV map
  const float* key_bias;            // [B,Sk] or null
  const int* kv_len;                // [B] or null
  const float* lse2;                // [B,heads,Sq]
  const float* delta;               // [B,heads,Sq]  rowsum(dO o O)
  float* dq_acc;                    // [B*Sq, ld_dq] fp32, zero-initialised by the caller; head h at columns 64h
  int ld_dq;
  __half* dk;                       // [B*Sk, ld_dkv]; head h at columns dk_col0 + 64h
  __half* dv;
  int ld_dkv, dk_col0, dv_col0;
  float scale_log2;                 // log2(e)/sqrt(d)
  float inv_sqrt_d;
  DropCfg drop;                     // the forward's attention-probability dropout (mask regenerated here)
  const int* cu_seqlens;            // [B+1] or null: packed rows, as in AttnFwdArgs (Sq / Sk = the maximum length)
  int dq_half;                      // != 0: dQ leaves as an fp16 TMA reduce-add straight into the (zeroed) dQ columns of the gradient
  int dq_col0;                      // buffer (tmDQ is then an fp16 map with 64-byte boxes): no fp32 accumulator, memset or cast pass
};

// delta[b,h,q] = sum_d dO[q, 64h+d] * O[q, 64h+d]   (one warp per token row; 8 lanes share a head)
// Packed rows (row_ex / row_rank non-null): row r belongs to sequence row_ex[r] at position row_rank[r]; delta keeps its padded
// [B, heads, Sq] layout.
__global__ void __launch_bounds__(128) attn_delta_kernel(const __half* __restrict__ dout, int ld_do, const __half* __restrict__ out, int ld_o,
                                                         float* __restrict__ delta, int rows, int heads, int Sq,
                                                         const int* __restrict__ row_ex, const int* __restrict__ row_rank) {
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * 4 + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int b = row_ex ? row_ex[row] : row / Sq, q = row_rank ? row_rank[row] : row % Sq;
  const int nvec = heads * 8;
  for (int v = lane; v < nvec; v += 32) {
    const uint4 ra = *reinterpret_cast<const uint4*>(dout + static_cast<size_t>(row) * ld_do + v * 8);
    const uint4 rb = *reinterpret_cast<const uint4*>(out + static_cast<size_t>(row) * ld_o + v * 8);
    const __half2* ha = reinterpret_cast<const __half2*>(&ra);
    const __half2* hb = reinterpret_cast<const __half2*>(&rb);
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float2 x = __half22float2(ha[i]), y = __half22float2(hb[i]);
      s = fmaf(x.x, y.x, s);
      s = fmaf(x.y, y.y, s);
    }
    s += __shfl_xor_sync(0xffffffffu, s, 1);
    s += __shfl_xor_sync(0xffffffffu, s, 2);
    s += __shfl_xor_sync(0xffffffffu, s, 4);
    if ((lane & 7) == 0) delta[(static_cast<size_t>(b) * heads + (v >> 3)) * Sq + q] = s;
  }
}

// dq(fp16, strided) = dq_acc(fp32)
__global__ void dq_cast_kernel(const float* __restrict__ src, int ld_src, __half* __restrict__ dst, int ld_dst, int rows, int cols8) {
  const size_t total = static_cast<size_t>(rows) * cols8;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total; i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const size_t r = i / cols8, c = (i % cols8) * 8;
    const float4 x = *reinterpret_cast<const float4*>(src + r * ld_src + c);
    const float4 y = *reinterpret_cast<const float4*>(src + r * ld_src + c + 4);
    __half2 h[4] = {__floats2half2_rn(x.x, x.y), __floats2half2_rn(x.z, x.w), __floats2half2_rn(y.x, y.y), __floats2half2_rn(y.z, y.w)};
    *reinterpret_cast<uint4*>(dst + r * ld_dst + c) = *reinterpret_cast<uint4*>(h);
  }
}

}  // namespace b200
