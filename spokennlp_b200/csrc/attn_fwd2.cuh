// Fused attention forward, second generation (same contract as attn_fwd.cuh: head_dim 64, fp16 operands, fp32 softmax).
//
// What changed against attn_fwd_kernel, and why (profiles/r01d: 15 % tensor-pipe, 40 % issue utilisation, a quarter of
// the softmax warps' time spent waiting for their next score tile):
//   * ONE pass over the scores: a thread pulls its whole 128-key row out of TMEM into registers (4 back-to-back
//     tcgen05.ld, one wait), so every score is read from TMEM once instead of twice;
//   * the score buffer is handed back to the tensor core as soon as it sits in registers (`s_free`), so S(j+1) is
//     computed WHILE the softmax of block j runs — the softmax warps never wait for the MMA round trip again;
//   * each query tile has its own MMA-issuing warp (the two tiles no longer serialise behind one static issue order);
//   * masking is a pre-pass that only runs in blocks that contain masked keys (prefix masks: an index compare; general
//     additive bias: staged per block), so the common block carries ~4 instructions per score
//     (FMNMX3/2 + FFMA + MUFU.EX2 + FADD + F2FP/2).
//   warp 0       TMA producer (Q_A, Q_B once; K_j / V_j ring)
//   warp 1, 2    MMA issuers for query tile A / B
//   warps 3-10   softmax A (3-6) / softmax B (7-10): one thread per query row
#pragma once
#include "attn_fwd.cuh"

namespace b200 {

constexpr int ATT2_THREADS = 352;

template <bool DROP>
__global__ void __launch_bounds__(ATT2_THREADS, 1)
attn_fwd2_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmKV,
                 const __grid_constant__ CUtensorMap tmO, const AttnFwdArgs a) {
  using S = AttnFwdSmem;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + S::OFF_BAR);
  uint64_t* q_full = bars;                       // 1
  uint64_t* kv_full = bars + 1;                  // KV_STAGES
  uint64_t* kv_empty = kv_full + ATT_KV_STAGES;  // KV_STAGES
  uint64_t* s_full = kv_empty + ATT_KV_STAGES;   // 2
  uint64_t* p_full = s_full + 2;                 // 2
  uint64_t* o_full = p_full + 2;                 // 2
  uint64_t* s_free = o_full + 2;                 // 2
  uint64_t* b_go = s_free + 2;                   // 1: tile B starts half a period after tile A (see the MMA warps)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(b_go + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.z, h = blockIdx.y;
  const int q0 = blockIdx.x * 2 * ATT_BQ;
  const bool tileB = (q0 + ATT_BQ) < a.Sq;
  int kv_len = a.kv_len ? a.kv_len[b] : a.Sk;
  const bool general_bias = a.key_bias != nullptr && (a.kv_len == nullptr || kv_len < 0);   // see attn_fwd.cuh
  kv_len = max(1, min(kv_len < 0 ? -kv_len : kv_len, a.Sk));
  const int n_blocks = (kv_len + ATT_BK - 1) / ATT_BK;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmKV);
    tma_prefetch_desc(&tmO);
    mbar_init(q_full, 1);
    for (int i = 0; i < ATT_KV_STAGES; ++i) {
      mbar_init(&kv_full[i], 1);
      mbar_init(&kv_empty[i], tileB ? 2 : 1);     // one commit per MMA warp that consumed the stage
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&s_full[i], 1);
      mbar_init(&p_full[i], 128);
      mbar_init(&o_full[i], 1);
      mbar_init(&s_free[i], 128);
    }
    mbar_init(b_go, 1);
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
  // TMEM columns: S_A [0,128)  S_B [128,256)  O_A [256,320)  O_B [320,384)

  if (warp == 0) {
    if (lane == 0) {
      mbar_expect_tx(q_full, (tileB ? 2 : 1) * S::Q_BYTES);
      tma_load_2d(smem + S::OFF_Q, &tmQ, q_full, a.q_col0 + h * ATT_D, b * a.Sq + q0);
      if (tileB) tma_load_2d(smem + S::OFF_Q + S::Q_BYTES, &tmQ, q_full, a.q_col0 + h * ATT_D, b * a.Sq + q0 + ATT_BQ);
      uint32_t stage = 0, phase = 0;
      for (int j = 0; j < n_blocks; ++j) {
        mbar_wait(&kv_empty[stage], phase ^ 1);
        mbar_expect_tx(&kv_full[stage], 2 * S::KV_BYTES);
        uint8_t* dst = smem + S::OFF_KV + stage * 2 * S::KV_BYTES;
        tma_load_2d(dst, &tmKV, &kv_full[stage], a.k_col0 + h * ATT_D, b * a.Sk + j * ATT_BK);
        tma_load_2d(dst + S::KV_BYTES, &tmKV, &kv_full[stage], a.v_col0 + h * ATT_D, b * a.Sk + j * ATT_BK);
        if (++stage == ATT_KV_STAGES) { stage = 0; phase ^= 1; }
      }
    }
    __syncwarp();
  } else if (warp <= 2) {
    // ------------------------------------------------------------------ MMA issuer of query tile x
    const int x = warp - 1;
    if (x == 0 || tileB) {
      constexpr uint32_t idesc_s = make_idesc_f16(128, ATT_BK, 0, 0);
      constexpr uint32_t idesc_o = make_idesc_f16(128, ATT_D, 0, 1);
      const uint32_t qa = smem_u32(smem + S::OFF_Q + x * S::Q_BYTES);
      const uint32_t pa = smem_u32(smem + S::OFF_P + x * S::P_BYTES);
      auto issue_s = [&](uint32_t stage) {          // S_x = Q_x K^T
        const uint32_t ka = smem_u32(smem + S::OFF_KV + stage * 2 * S::KV_BYTES);
#pragma unroll
        for (int kk = 0; kk < ATT_D / 16; ++kk)
          umma_ss(tmem + x * 128, make_smem_desc(qa + kk * 32, 0, 1024), make_smem_desc(ka + kk * 32, 0, 1024), idesc_s, kk > 0);
        umma_commit(&s_full[x]);
      };
      mbar_wait(q_full, 0);
      mbar_wait(&kv_full[0], 0);
      // The TMEM read port (score rows -> registers) and the MUFU (exp2) are the two busy resources of a block and a
      // softmax group uses them one after the other.  Tile B therefore starts when tile A has pulled its first score
      // rows: from then on one group reads TMEM while the other one exponentiates.
      if (x == 1) mbar_wait(b_go, 0);
      tc_fence_after();
      if (lane == 0) issue_s(0);
      __syncwarp();
      uint32_t stage = 0, phase = 0;
      for (int j = 0; j < n_blocks; ++j) {
        uint32_t nstage = stage + 1, nphase = phase;
        if (nstage == ATT_KV_STAGES) { nstage = 0; nphase ^= 1; }
        if (j + 1 < n_blocks) {                     // next scores as soon as the softmax threads hold block j in registers
          mbar_wait(&kv_full[nstage], nphase);
          mbar_wait(&s_free[x], j & 1);
          tc_fence_after();
          if (lane == 0) issue_s(nstage);
          __syncwarp();
        }
        mbar_wait(&p_full[x], j & 1);
        tc_fence_after();
        if (lane == 0) {
          const uint32_t va = smem_u32(smem + S::OFF_KV + stage * 2 * S::KV_BYTES + S::KV_BYTES);
#pragma unroll
          for (int kk = 0; kk < ATT_BK / 16; ++kk)
            umma_ss(tmem + 256 + x * 64, make_smem_desc(pa + (kk >> 2) * 16384 + (kk & 3) * 32, 0, 1024),
                    make_smem_desc(va + kk * 2048, 8192, 1024), idesc_o, (j > 0 || kk > 0) ? 1u : 0u);
          umma_commit(&o_full[x]);
          umma_commit(&kv_empty[stage]);
        }
        __syncwarp();
        stage = nstage;
        phase = nphase;
      }
    }
  } else {
    // ------------------------------------------------------------------ softmax groups
    const int x = (warp - 3) >> 2;                // 0 = tile A, 1 = tile B
    if (x == 0 || tileB) {
      const int qd = warp & 3;                    // TMEM lane quadrant of this warp
      const int r = qd * 32 + lane;               // row inside the query tile
      const int t = ((warp - 3) & 3) * 32 + lane; // 0..127 inside the group (bias staging)
      const uint32_t lane_addr = static_cast<uint32_t>(qd * 32) << 16;
      const uint32_t bias_s = smem_u32(smem + S::OFF_BIAS) + x * 2 * ATT_BK * 4;
      const uint32_t p_row = smem_u32(smem + S::OFF_P + x * S::P_BYTES) + r * 128;
      const float NEG_INF = -INFINITY;
      const float sc = a.scale_log2;
      const float inv_sc = 1.0f / sc;
      const uint32_t dseed = DROP ? drop_seed(a.drop) : 0u;
      const uint32_t dbase = DROP ? static_cast<uint32_t>(((static_cast<size_t>(b) * a.heads + h) * a.Sq + min(q0 + x * ATT_BQ + r, a.Sq - 1)) * ((a.Sk + 1) >> 1)) : 0u;
      float m = NEG_INF, l = 0.f;
      for (int j = 0; j < n_blocks; ++j) {
        const bool partial = (j + 1) * ATT_BK > kv_len;
        const uint32_t bj = bias_s + (j & 1) * ATT_BK * 4;
        if (general_bias) {                       // bias in units of raw scores: (s + bias/scale) * scale = s*scale + bias
          const int key = j * ATT_BK + t;
          float bv = NEG_INF;
          if (key < kv_len) bv = a.key_bias[static_cast<size_t>(b) * a.Sk + key] * 1.4426950408889634f * inv_sc;
          sts_f32(bj + t * 4, bv);
          asm volatile("bar.sync %0, 128;" ::"r"(1 + x) : "memory");
        }
        mbar_wait(&s_full[x], j & 1);
        tc_fence_after();
        uint32_t v[128];
        tmem_ld_x32(tmem + lane_addr + x * 128, *reinterpret_cast<uint32_t(*)[32]>(&v[0]));
        tmem_ld_x32(tmem + lane_addr + x * 128 + 32, *reinterpret_cast<uint32_t(*)[32]>(&v[32]));
        tmem_ld_x32(tmem + lane_addr + x * 128 + 64, *reinterpret_cast<uint32_t(*)[32]>(&v[64]));
        tmem_ld_x32(tmem + lane_addr + x * 128 + 96, *reinterpret_cast<uint32_t(*)[32]>(&v[96]));
        tmem_wait_ld();
        tc_fence_before();
        mbar_arrive(&s_free[x]);                  // the tensor core may overwrite S_x with block j+1 now
        if (x == 0 && j == 0 && t == 0) mbar_arrive(b_go);
        // ---- masked keys (only blocks that have any)
        if (general_bias) {
#pragma unroll
          for (int i = 0; i < 128; i += 4) {
            const uint4 bb = lds128(bj + i * 4);
            v[i] = __float_as_uint(__uint_as_float(v[i]) + __uint_as_float(bb.x));
            v[i + 1] = __float_as_uint(__uint_as_float(v[i + 1]) + __uint_as_float(bb.y));
            v[i + 2] = __float_as_uint(__uint_as_float(v[i + 2]) + __uint_as_float(bb.z));
            v[i + 3] = __float_as_uint(__uint_as_float(v[i + 3]) + __uint_as_float(bb.w));
          }
        } else if (partial) {
          const int lim = kv_len - j * ATT_BK;    // keys [0, lim) of this block are kept
#pragma unroll
          for (int i = 0; i < 128; ++i) v[i] = (i < lim) ? v[i] : 0xff800000u;
        }
        // ---- row maximum
        float m0 = fmaxf(__uint_as_float(v[0]), __uint_as_float(v[1])), m1 = fmaxf(__uint_as_float(v[2]), __uint_as_float(v[3]));
#pragma unroll
        for (int i = 4; i < 128; i += 4) {
          m0 = fmaxf(m0, fmaxf(__uint_as_float(v[i]), __uint_as_float(v[i + 1])));
          m1 = fmaxf(m1, fmaxf(__uint_as_float(v[i + 2]), __uint_as_float(v[i + 3])));
        }
        const float mx = fmaxf(m0, m1) * sc;      // sc > 0
        // ---- lazy rescale: keep the running reference max unless some row of this warp grew by more than 2^8
        const float m_new = fmaxf(m, mx);
        const bool grow = (m_new - m) > 8.0f || m == NEG_INF;
        const bool rescale = __any_sync(0xffffffffu, grow);
        float m_use = rescale ? m_new : m;
        if (m_use == NEG_INF) m_use = 0.f;
        const float alpha = rescale ? fast_exp2(m - m_use) : 1.0f;     // m == -inf -> 0
        if (j > 0) {                              // P.V of block j-1 finished: P smem is free, O may be rescaled
          mbar_wait(&o_full[x], (j - 1) & 1);
          tc_fence_after();
          if (rescale) {
#pragma unroll
            for (int c = 0; c < 4; ++c) {
              uint32_t o[16];
              tmem_ld_x16(tmem + lane_addr + 256 + x * 64 + c * 16, o);
              tmem_wait_ld();
#pragma unroll
              for (int i = 0; i < 16; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
              tmem_st_x16(tmem + lane_addr + 256 + x * 64 + c * 16, o);
            }
            tmem_wait_st();
          }
        }
        // ---- p = exp2(s * scale - m_use), row sum, fp16 P into the swizzled smem tile (16 chunks of 8 keys)
        const float neg_m = -m_use;
        float rs0 = 0.f, rs1 = 0.f;
#pragma unroll
        for (int ch = 0; ch < 16; ++ch) {
          uint32_t pk[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int i = ch * 8 + 2 * e;
            float p0 = fast_exp2(fmaf(__uint_as_float(v[i]), sc, neg_m));
            float p1 = fast_exp2(fmaf(__uint_as_float(v[i + 1]), sc, neg_m));
            rs0 += p0;
            rs1 += p1;
            if (DROP) {                             // the row sum (softmax denominator) is taken before dropout
              float k0, k1;
              drop_pair(dbase + ((j * ATT_BK + i) >> 1), dseed, a.drop.thr15, a.drop.scale, k0, k1);
              p0 *= k0;
              p1 *= k1;
            }
            const __half2 hp = __floats2half2_rn(p0, p1);
            pk[e] = *reinterpret_cast<const uint32_t*>(&hp);
          }
          sts128(p_row + (ch >> 3) * 16384 + (((ch & 7) ^ (r & 7)) << 4), pk[0], pk[1], pk[2], pk[3]);
        }
        l = fmaf(l, alpha, rs0 + rs1);
        m = (m_use == 0.f && m_new == NEG_INF) ? NEG_INF : m_use;
        fence_proxy_async_smem();
        tc_fence_before();
        mbar_arrive(&p_full[x]);
      }
      // ---------------------------------------------------------------- finalise: O / l -> ctx, LSE
      mbar_wait(&o_full[x], (n_blocks - 1) & 1);
      tc_fence_after();
      const int qrow = q0 + x * ATT_BQ + r;
      const float inv_l = l > 0.f ? 1.0f / l : 0.f;
      uint32_t o[2][32];
      tmem_ld_x32(tmem + lane_addr + 256 + x * 64, o[0]);
      tmem_ld_x32(tmem + lane_addr + 256 + x * 64 + 32, o[1]);
      tmem_wait_ld();
      // The 32 rows of this warp leave as ONE TMA store from a 128B-swizzled staging patch (the P tile is free by now):
      // a thread-per-row st.global touches 32 different 128-byte lines per instruction and kept the LSU busy for ~1.4 us
      // per CTA (profiles/r01e).  A warp whose rows straddle the end of the batch element stores them directly.
      uint32_t w[32];
#pragma unroll
      for (int k = 0; k < 32; ++k) {
        const __half2 hv = __floats2half2_rn(__uint_as_float(o[k >> 4][2 * (k & 15)]) * inv_l, __uint_as_float(o[k >> 4][2 * (k & 15) + 1]) * inv_l);
        w[k] = *reinterpret_cast<const uint32_t*>(&hv);
      }
      const int wrow0 = q0 + x * ATT_BQ + qd * 32;           // first query row of this warp
      if (wrow0 + 32 <= a.Sq) {
#pragma unroll
        for (int ch = 0; ch < 8; ++ch) sts128(p_row + ((ch ^ (r & 7)) << 4), w[4 * ch], w[4 * ch + 1], w[4 * ch + 2], w[4 * ch + 3]);
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) {
          tma_store_2d(&tmO, smem + S::OFF_P + x * S::P_BYTES + qd * 32 * 128, h * ATT_D, b * a.Sq + wrow0);
          tma_commit_group();
          tma_wait_group_read<0>();
        }
      } else if (qrow < a.Sq) {
        __half* dst = a.out + (static_cast<size_t>(b) * a.Sq + qrow) * a.ld_out + h * ATT_D;
#pragma unroll
        for (int ch = 0; ch < 8; ++ch) *reinterpret_cast<uint4*>(dst + ch * 8) = make_uint4(w[4 * ch], w[4 * ch + 1], w[4 * ch + 2], w[4 * ch + 3]);
      }
      if (qrow < a.Sq) {
        if (a.lse2) a.lse2[(static_cast<size_t>(b) * a.heads + h) * a.Sq + qrow] = (l > 0.f) ? (m + log2f(l)) : NEG_INF;
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

}  // namespace b200
