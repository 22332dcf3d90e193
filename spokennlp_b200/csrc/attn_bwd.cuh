// Fused attention backward for head_dim 64 on tcgen05 — the autograd of bert_model.py:309-350.
//
// One CTA owns one 128-key block of one (batch, head) and walks the query blocks.  Everything is computed in the
// TRANSPOSED orientation (TMEM lane == key row), so that the two tiles the softmax threads write, P^T and dS^T
// ([key, query], fp16, 128B-swizzled), feed all three gradient MMAs without any transposition:
//     S^T  = K   Q^T          (A = K  K-major,  B = Q  K-major)     recompute scores
//     dP^T = V   dO^T         (A = V  K-major,  B = dO K-major)
//     P^T  = exp2(S^T c + bias_k - lse2_q),   dS^T = P^T o (dP^T - delta_q) / sqrt(d)        [CUDA cores, row per thread]
//     dV  += P^T  dO          (A = P^T  K-major, B = dO MN-major)   accumulates in TMEM over query blocks
//     dK  += dS^T Q           (A = dS^T K-major, B = Q  MN-major)   accumulates in TMEM over query blocks
//     dQ_i = dS   K           (A = dS^T viewed MN-major, B = K MN-major) -> fp32 red.global.add into dq_acc
// Q / K / V / dO tiles are read in place from the packed projection buffers with TMA; MN-major views of the same
// swizzled smem tiles are selected purely through the UMMA descriptors.
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
  int dbg;                          // measurement knobs: 0x10000 skip dQ reductions, 0x20000 skip gradient MMAs, 0x40000 skip exp math
  DropCfg drop;                     // the forward's attention-probability dropout (mask regenerated here)
};

struct AttnBwdSmem {
  static constexpr int T = ATT_BK * ATT_D * 2;                 // 16 KB tile
  static constexpr int OFF_K = 0;
  static constexpr int OFF_V = OFF_K + T;
  static constexpr int OFF_QDO = OFF_V + T;                    // 2 stages x (Q_i, dO_i)
  static constexpr int OFF_P = OFF_QDO + 2 * 2 * T;            // P^T  [128 keys][128 q] fp16 = 32 KB
  static constexpr int OFF_DS = OFF_P + 2 * T;                 // dS^T 32 KB
  static constexpr int STAT_Q = 2048;                          // queries whose (lse2, delta) are staged at once
  static constexpr int OFF_STAT = OFF_DS + 2 * T;              // [2 (lse2, delta)][STAT_Q] floats
  static constexpr int OFF_BAR = OFF_STAT + 2 * STAT_Q * 4;
  static constexpr int TOTAL = OFF_BAR + 256 + 1024;
};

// Schedule per query block i (tensor core and CUDA cores overlap):
//   MMA warp : wait ds_full(i) -> issue S^T(i+1), dP^T(i+1) (their TMEM is free: block i was read out) -> s_full(i+1)
//              -> issue dV += , dK += , dQ(i) = -> grad_done(i)
//   softmax  : wait s_full(i+1) -> exp / dS math for block i+1 into registers WHILE the gradient MMAs of block i run
//              -> wait grad_done(i) -> drain dQ(i) (fp32 reduction) -> write P^T / dS^T tiles -> ds_full(i+1)
template <bool DROP>
__global__ void __launch_bounds__(ATTB_THREADS, 1)
attn_bwd_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmKV,
                const __grid_constant__ CUtensorMap tmDO, const AttnBwdArgs a) {
  using S = AttnBwdSmem;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + S::OFF_BAR);
  uint64_t* kv_full = bars;            // 1
  uint64_t* qdo_full = bars + 1;       // 2
  uint64_t* qdo_empty = bars + 3;      // 2
  uint64_t* s_full = bars + 5;         // 1
  uint64_t* ds_full = bars + 6;        // 1 (256 arrivals)
  uint64_t* grad_done = bars + 7;      // 1
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.z, h = blockIdx.y;
  const int k0 = blockIdx.x * ATT_BK;
  int kv_len = a.kv_len ? a.kv_len[b] : a.Sk;
  const bool general_bias = a.key_bias != nullptr && (a.kv_len == nullptr || kv_len < 0);   // see attn_fwd.cuh
  kv_len = max(1, min(kv_len < 0 ? -kv_len : kv_len, a.Sk));
  const int nq = (a.Sq + ATT_BQ - 1) / ATT_BQ;
  const bool dead_block = k0 >= kv_len;      // every key of this block is masked: dK = dV = 0, no dQ contribution

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmKV);
    tma_prefetch_desc(&tmDO);
    mbar_init(kv_full, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&qdo_full[i], 1);
      mbar_init(&qdo_empty[i], 1);
    }
    mbar_init(s_full, 1);
    mbar_init(ds_full, 256);
    mbar_init(grad_done, 1);
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  // TMEM columns: S^T [0,128)  dP^T [128,256)  dV [256,320)  dK [320,384)  dQ ping-pong [384,448) / [448,512)

  if (dead_block) {
    if (warp >= 2 && warp < 6) {
      const int r = (warp & 3) * 32 + lane;
      if (k0 + r < a.Sk) {
        const size_t row = static_cast<size_t>(b) * a.Sk + k0 + r;
        const uint4 z = make_uint4(0, 0, 0, 0);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          *reinterpret_cast<uint4*>(a.dk + row * a.ld_dkv + a.dk_col0 + h * ATT_D + i * 8) = z;
          *reinterpret_cast<uint4*>(a.dv + row * a.ld_dkv + a.dv_col0 + h * ATT_D + i * 8) = z;
        }
      }
    }
  } else if (warp == 0) {
    if (lane == 0) {
      mbar_expect_tx(kv_full, 2 * S::T);
      tma_load_2d(smem + S::OFF_K, &tmKV, kv_full, a.k_col0 + h * ATT_D, b * a.Sk + k0);
      tma_load_2d(smem + S::OFF_V, &tmKV, kv_full, a.v_col0 + h * ATT_D, b * a.Sk + k0);
      for (int i = 0; i < nq; ++i) {
        const int st = i & 1;
        mbar_wait(&qdo_empty[st], ((i >> 1) & 1) ^ 1);
        mbar_expect_tx(&qdo_full[st], 2 * S::T);
        uint8_t* dst = smem + S::OFF_QDO + st * 2 * S::T;
        tma_load_2d(dst, &tmQ, &qdo_full[st], a.q_col0 + h * ATT_D, b * a.Sq + i * ATT_BQ);
        tma_load_2d(dst + S::T, &tmDO, &qdo_full[st], h * ATT_D, b * a.Sq + i * ATT_BQ);
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    constexpr uint32_t idesc_s = make_idesc_f16(128, 128, 0, 0);
    constexpr uint32_t idesc_g = make_idesc_f16(128, 64, 0, 1);     // dV, dK : A K-major, B MN-major
    constexpr uint32_t idesc_q = make_idesc_f16(128, 64, 1, 1);     // dQ     : both MN-major
    const uint32_t ka = smem_u32(smem + S::OFF_K), va = smem_u32(smem + S::OFF_V);
    const uint32_t pa = smem_u32(smem + S::OFF_P), dsa = smem_u32(smem + S::OFF_DS);
    auto issue_scores = [&](int st) {
      const uint32_t qa = smem_u32(smem + S::OFF_QDO + st * 2 * S::T), doa = qa + S::T;
#pragma unroll
      for (int kk = 0; kk < 4; ++kk)
        umma_ss(tmem + 0, make_smem_desc(ka + kk * 32, 0, 1024), make_smem_desc(qa + kk * 32, 0, 1024), idesc_s, kk > 0);
#pragma unroll
      for (int kk = 0; kk < 4; ++kk)
        umma_ss(tmem + 128, make_smem_desc(va + kk * 32, 0, 1024), make_smem_desc(doa + kk * 32, 0, 1024), idesc_s, kk > 0);
      umma_commit(s_full);
    };
    mbar_wait(kv_full, 0);
    mbar_wait(&qdo_full[0], 0);
    tc_fence_after();
    if (lane == 0) issue_scores(0);
    __syncwarp();
    for (int i = 0; i < nq; ++i) {
      const int st = i & 1;
      mbar_wait(ds_full, i & 1);
      tc_fence_after();
      if (i + 1 < nq) {                       // scores of the next block first: the softmax warps start on them at once
        mbar_wait(&qdo_full[st ^ 1], ((i + 1) >> 1) & 1);
        tc_fence_after();
        if (lane == 0) issue_scores(st ^ 1);
        __syncwarp();
      }
      if (lane == 0) {
        const uint32_t qa = smem_u32(smem + S::OFF_QDO + st * 2 * S::T), doa = qa + S::T;
        if (!(a.dbg & 0x20000)) {
#pragma unroll
        for (int kk = 0; kk < 8; ++kk)        // dV += P^T dO
          umma_ss(tmem + 256, make_smem_desc(pa + (kk >> 2) * 16384 + (kk & 3) * 32, 0, 1024), make_smem_desc(doa + kk * 2048, 8192, 1024),
                  idesc_g, (i > 0 || kk > 0) ? 1u : 0u);
#pragma unroll
        for (int kk = 0; kk < 8; ++kk)        // dK += dS^T Q
          umma_ss(tmem + 320, make_smem_desc(dsa + (kk >> 2) * 16384 + (kk & 3) * 32, 0, 1024), make_smem_desc(qa + kk * 2048, 8192, 1024),
                  idesc_g, (i > 0 || kk > 0) ? 1u : 0u);
#pragma unroll
        for (int kk = 0; kk < 8; ++kk)        // dQ_i = dS K   (A = dS^T tile viewed MN-major: M = query, K = key rows)
          umma_ss(tmem + 384 + (i & 1) * 64, make_smem_desc(dsa + kk * 2048, 16384, 1024), make_smem_desc(ka + kk * 2048, 8192, 1024), idesc_q, kk > 0);
        }
        umma_commit(&qdo_empty[st]);
        umma_commit(grad_done);
      }
      __syncwarp();
    }
  } else {
    const int qd = warp & 3;                      // TMEM lane quadrant
    const int half = (warp - 2) >> 2;             // which 64 of the 128 query columns (and which 32 of the 64 d columns)
    const int r = qd * 32 + lane;                 // key row inside the block == TMEM lane (query row for the dQ tile)
    const int t = threadIdx.x - 64;               // 0..255
    const uint32_t lane_addr = static_cast<uint32_t>(qd * 32) << 16;
    const uint32_t stat = smem_u32(smem + S::OFF_STAT);
    const uint32_t p_row = smem_u32(smem + S::OFF_P) + half * 16384 + r * 128;
    const uint32_t ds_row = smem_u32(smem + S::OFF_DS) + half * 16384 + r * 128;
    const int key = k0 + r;
    float bias = -INFINITY;
    if (key < kv_len) bias = general_bias ? a.key_bias[static_cast<size_t>(b) * a.Sk + key] * 1.4426950408889634f : 0.f;
    const size_t stat_base = (static_cast<size_t>(b) * a.heads + h) * a.Sq;
    const uint32_t dseed = DROP ? drop_seed(a.drop) : 0u;

    auto drain_dq = [&](int i) {                  // dQ_i tile: TMEM lane == query row; this warp owns 32 of the 64 d columns
      const int q = i * ATT_BQ + r;
      uint32_t o[32];
      tmem_ld_x32(tmem + lane_addr + 384 + (i & 1) * 64 + half * 32, o);
      tmem_wait_ld();
      if (q < a.Sq && !(a.dbg & 0x10000)) {
        float* dst = a.dq_acc + (static_cast<size_t>(b) * a.Sq + q) * a.ld_dq + h * ATT_D + half * 32;
#pragma unroll
        for (int k = 0; k < 8; ++k)
          red_add_v4(dst + k * 4, __uint_as_float(o[4 * k]), __uint_as_float(o[4 * k + 1]), __uint_as_float(o[4 * k + 2]),
                     __uint_as_float(o[4 * k + 3]));
      }
    };

    // (lse2, delta) of the queries are staged STAT_Q at a time (all of them at once for Sq <= 2048), so the hot loop
    // carries no block-wide barrier; queries past Sq get lse = +inf (P = 0) and delta = 0
    const int blocks_per_stage = S::STAT_Q / ATT_BQ;
    for (int i = 0; i < nq; ++i) {
      if (i % blocks_per_stage == 0) {
        if (i > 0) asm volatile("bar.sync 1, 256;" ::: "memory");       // everyone is done with the previous stage
        for (int qq = t; qq < S::STAT_Q; qq += 256) {
          const int q = i * ATT_BQ + qq;
          sts_f32(stat + qq * 4, q < a.Sq ? a.lse2[stat_base + q] : INFINITY);
          sts_f32(stat + S::STAT_Q * 4 + qq * 4, q < a.Sq ? a.delta[stat_base + q] : 0.f);
        }
        asm volatile("bar.sync 1, 256;" ::: "memory");
      }
      const uint32_t lse_s = stat + ((i % blocks_per_stage) * ATT_BQ + half * 64) * 4;
      const uint32_t del_s = lse_s + S::STAT_Q * 4;
      mbar_wait(s_full, i & 1);
      tc_fence_after();
      uint32_t pk[32], dk[32];                    // this thread's 64 P^T / dS^T values, packed fp16
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        uint32_t sv[32], dp[32];
        tmem_ld_x32(tmem + lane_addr + half * 64 + c * 32, sv);
        tmem_ld_x32(tmem + lane_addr + 128 + half * 64 + c * 32, dp);
        tmem_wait_ld();
        if (a.dbg & 0x40000) {
#pragma unroll
          for (int e = 0; e < 16; ++e) pk[c * 16 + e] = dk[c * 16 + e] = sv[e] ^ dp[e];
          continue;
        }
#pragma unroll
        for (int e = 0; e < 16; ++e) {
          const int qi = c * 32 + 2 * e;
          float p0 = fast_exp2(fmaf(__uint_as_float(sv[2 * e]), a.scale_log2, bias) - lds_f32(lse_s + qi * 4));
          float p1 = fast_exp2(fmaf(__uint_as_float(sv[2 * e + 1]), a.scale_log2, bias) - lds_f32(lse_s + qi * 4 + 4));
          float g0 = __uint_as_float(dp[2 * e]), g1 = __uint_as_float(dp[2 * e + 1]);
          float pd0 = p0, pd1 = p1;
          if (DROP) {       // P_drop = P o mask/(1-p) feeds dV; dP flows back through the same mask; delta is unchanged
            const int q = min(i * ATT_BQ + half * 64 + qi, a.Sq - 1);
            const uint32_t skp = static_cast<uint32_t>((a.Sk + 1) >> 1), kc = static_cast<uint32_t>(min(key, a.Sk - 1));
            const uint32_t e0 = (static_cast<uint32_t>(stat_base + q) * skp + (kc >> 1)) * 2u + (kc & 1u);   // (pair, lane) as the forward drew it
            const float m0 = drop_one(e0, dseed, a.drop.thr15, a.drop.scale);
            const float m1 = drop_one(e0 + 2u * skp, dseed, a.drop.thr15, a.drop.scale);
            pd0 *= m0; pd1 *= m1;
            g0 *= m0; g1 *= m1;
          }
          const float d0 = p0 * (g0 - lds_f32(del_s + qi * 4)) * a.inv_sqrt_d;
          const float d1 = p1 * (g1 - lds_f32(del_s + qi * 4 + 4)) * a.inv_sqrt_d;
          const __half2 hp = __floats2half2_rn(pd0, pd1), hd = __floats2half2_rn(d0, d1);
          pk[c * 16 + e] = *reinterpret_cast<const uint32_t*>(&hp);
          dk[c * 16 + e] = *reinterpret_cast<const uint32_t*>(&hd);
        }
      }
      if (i > 0) {                                // gradient MMAs of block i-1 are done: dQ(i-1) is complete, P^T/dS^T are free
        mbar_wait(grad_done, (i - 1) & 1);
        tc_fence_after();
      }
#pragma unroll
      for (int ch = 0; ch < 8; ++ch) {
        const int off = ((ch ^ (r & 7)) << 4);
        sts128(p_row + off, pk[4 * ch], pk[4 * ch + 1], pk[4 * ch + 2], pk[4 * ch + 3]);
        sts128(ds_row + off, dk[4 * ch], dk[4 * ch + 1], dk[4 * ch + 2], dk[4 * ch + 3]);
      }
      fence_proxy_async_smem();
      tc_fence_before();
      mbar_arrive(ds_full);
      // dQ(i-1) sits in the other TMEM buffer: reduce it into HBM off the critical path (after releasing the MMA warp)
      if (i > 0) drain_dq(i - 1);
    }
    mbar_wait(grad_done, (nq - 1) & 1);
    tc_fence_after();
    drain_dq(nq - 1);
    // dV, dK: TMEM lane == key row; this warp owns 32 of the 64 d columns
#pragma unroll 1
    for (int which = 0; which < 2; ++which) {
      uint32_t o[32];
      tmem_ld_x32(tmem + lane_addr + 256 + which * 64 + half * 32, o);
      tmem_wait_ld();
      if (key < a.Sk) {
        __half* dst = (which == 0 ? a.dv + a.dv_col0 : a.dk + a.dk_col0) + (static_cast<size_t>(b) * a.Sk + key) * a.ld_dkv + h * ATT_D + half * 32;
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          uint32_t w[4];
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const __half2 hv = __floats2half2_rn(__uint_as_float(o[8 * e + 2 * k]), __uint_as_float(o[8 * e + 2 * k + 1]));
            w[k] = *reinterpret_cast<const uint32_t*>(&hv);
          }
          *reinterpret_cast<uint4*>(dst + e * 8) = make_uint4(w[0], w[1], w[2], w[3]);
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem, 512);
  }
}

// delta[b,h,q] = sum_d dO[q, 64h+d] * O[q, 64h+d]   (one warp per token row; 8 lanes share a head)
__global__ void __launch_bounds__(128) attn_delta_kernel(const __half* __restrict__ dout, int ld_do, const __half* __restrict__ out, int ld_o,
                                                         float* __restrict__ delta, int B, int heads, int Sq) {
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * 4 + (threadIdx.x >> 5);
  if (row >= B * Sq) return;
  const int b = row / Sq, q = row % Sq;
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
