// Fused scaled-dot-product attention forward for head_dim 64 on tcgen05.
//
//   O = softmax(Q K^T / sqrt(64) + key_bias) V      per (batch, head), never materialising P in HBM
//   (bert_model.py:309 QK^T, :328 /sqrt(d), :329-331 +mask, :334 softmax, :343 P V, :345-347 merge heads)
//
// One CTA owns TWO 128-row query tiles (A, B) of one (batch, head) and walks the keys in blocks of 128:
//   warp 0       TMA producer: Q_A, Q_B once; K_j / V_j into a 3-stage ring (shared by both query tiles)
//   warp 1       MMA issuer  : S_X = Q_X K_j^T (M128 N128 K64) into TMEM; O_X += P_X V_j (M128 N64 K128), V read
//                              MN-major straight from its row-major [key, d] tile (no V^T copy)
//   warps 2-5    softmax A   : one thread per query row (TMEM lane == row, so row max / sum need no shuffles);
//   warps 6-9    softmax B     exp2 on pre-scaled scores, online rescale of O in TMEM, P written fp16 into a
//                              128B-swizzled smem tile that feeds the P.V MMA
// While softmax A works on block j the tensor core runs S_B / P_B V: the two query tiles ping-pong.
// Q, K, V are read in place from the packed [tokens, 3H] projection output through TMA coordinates (no head
// split / permute kernels) and the context is written straight into [tokens, H].
#pragma once
#include "ptx.cuh"

namespace b200 {

constexpr int ATT_D = 64;
constexpr int ATT_BQ = 128;
constexpr int ATT_BK = 128;
constexpr int ATT_KV_STAGES = 3;
constexpr int ATT_THREADS = 320;

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
};

struct AttnFwdSmem {
  static constexpr int Q_BYTES = ATT_BQ * ATT_D * 2;                 // 16 KB per query tile
  static constexpr int KV_BYTES = ATT_BK * ATT_D * 2;                // 16 KB each for K_j and V_j
  static constexpr int P_BYTES = ATT_BQ * ATT_BK * 2;                // 32 KB per query tile
  static constexpr int OFF_Q = 0;
  static constexpr int OFF_KV = OFF_Q + 2 * Q_BYTES;
  static constexpr int OFF_P = OFF_KV + ATT_KV_STAGES * 2 * KV_BYTES;
  static constexpr int OFF_BIAS = OFF_P + 2 * P_BYTES;               // [2 tiles][2 buffers][128] floats
  static constexpr int OFF_BAR = OFF_BIAS + 2 * 2 * ATT_BK * 4;
  static constexpr int TOTAL = OFF_BAR + 256 + 1024;
};

template <bool DROP>
__global__ void __launch_bounds__(ATT_THREADS, 1)
attn_fwd_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmKV, const AttnFwdArgs a) {
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
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_full + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.z, h = blockIdx.y;
  const int q0 = blockIdx.x * 2 * ATT_BQ;
  const bool tileB = (q0 + ATT_BQ) < a.Sq;
  int kv_len = a.kv_len ? a.kv_len[b] : a.Sk;
  // a negative kv_len (or a bias without kv_len) means "arbitrary additive bias": every key reads its bias value;
  // otherwise the kept keys are the prefix [0, kv_len) and only the last partial block needs any masking
  const bool general_bias = a.key_bias != nullptr && (a.kv_len == nullptr || kv_len < 0);
  kv_len = max(1, min(kv_len < 0 ? -kv_len : kv_len, a.Sk));
  const int n_blocks = (kv_len + ATT_BK - 1) / ATT_BK;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmKV);
    mbar_init(q_full, 1);
    for (int i = 0; i < ATT_KV_STAGES; ++i) {
      mbar_init(&kv_full[i], 1);
      mbar_init(&kv_empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&s_full[i], 1);
      mbar_init(&p_full[i], 128);
      mbar_init(&o_full[i], 1);
    }
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
  } else if (warp == 1) {
    constexpr uint32_t idesc_s = make_idesc_f16(128, ATT_BK, 0, 0);
    constexpr uint32_t idesc_o = make_idesc_f16(128, ATT_D, 0, 1);
    const int n_tiles = tileB ? 2 : 1;
    auto issue_s = [&](int x, uint32_t stage) {          // S_X = Q_X K^T
      const uint32_t qa = smem_u32(smem + S::OFF_Q + x * S::Q_BYTES);
      const uint32_t ka = smem_u32(smem + S::OFF_KV + stage * 2 * S::KV_BYTES);
#pragma unroll
      for (int kk = 0; kk < ATT_D / 16; ++kk)
        umma_ss(tmem + x * 128, make_smem_desc(qa + kk * 32, 0, 1024), make_smem_desc(ka + kk * 32, 0, 1024), idesc_s, kk > 0);
      umma_commit(&s_full[x]);
    };
    mbar_wait(q_full, 0);
    mbar_wait(&kv_full[0], 0);
    tc_fence_after();
    if (lane == 0)
      for (int x = 0; x < n_tiles; ++x) issue_s(x, 0);
    __syncwarp();
    uint32_t stage = 0, phase = 0;
    for (int j = 0; j < n_blocks; ++j) {
      uint32_t nstage = stage + 1, nphase = phase;
      if (nstage == ATT_KV_STAGES) { nstage = 0; nphase ^= 1; }
      if (j + 1 < n_blocks) mbar_wait(&kv_full[nstage], nphase);
      for (int x = 0; x < n_tiles; ++x) {
        mbar_wait(&p_full[x], j & 1);
        tc_fence_after();
        if (lane == 0) {
          const uint32_t pa = smem_u32(smem + S::OFF_P + x * S::P_BYTES);
          const uint32_t va = smem_u32(smem + S::OFF_KV + stage * 2 * S::KV_BYTES + S::KV_BYTES);
#pragma unroll
          for (int kk = 0; kk < ATT_BK / 16; ++kk)
            umma_ss(tmem + 256 + x * 64, make_smem_desc(pa + (kk >> 2) * 16384 + (kk & 3) * 32, 0, 1024),
                    make_smem_desc(va + kk * 2048, 8192, 1024), idesc_o, (j > 0 || kk > 0) ? 1u : 0u);
          umma_commit(&o_full[x]);
          if (x == n_tiles - 1) umma_commit(&kv_empty[stage]);
          if (j + 1 < n_blocks) issue_s(x, nstage);
        }
        __syncwarp();
      }
      stage = nstage;
      phase = nphase;
    }
  } else {
    // ------------------------------------------------------------------ softmax groups
    const int x = (warp - 2) >> 2;                // 0 = tile A, 1 = tile B
    if (x == 0 || tileB) {
      const int qd = warp & 3;                    // TMEM lane quadrant
      const int r = qd * 32 + lane;               // row inside the query tile
      const int t = threadIdx.x - 64 - x * 128;   // 0..127 inside the group (bias staging)
      const uint32_t lane_addr = static_cast<uint32_t>(qd * 32) << 16;
      const uint32_t bias_s = smem_u32(smem + S::OFF_BIAS) + x * 2 * ATT_BK * 4;
      const uint32_t p_row = smem_u32(smem + S::OFF_P + x * S::P_BYTES) + r * 128;
      const float NEG_INF = -INFINITY;
      const float sc = a.scale_log2;
      // one hash covers two consecutive keys: pair index of (row, key) = ((b*heads + h)*Sq + q) * ceil(Sk/2) + key/2
      const uint32_t dseed = DROP ? drop_seed(a.drop) : 0u;
      const uint32_t dbase = DROP ? static_cast<uint32_t>(((static_cast<size_t>(b) * a.heads + h) * a.Sq + min(q0 + x * ATT_BQ + r, a.Sq - 1)) * ((a.Sk + 1) >> 1)) : 0u;
      float m = NEG_INF, l = 0.f;
      for (int j = 0; j < n_blocks; ++j) {
        // Only blocks that contain a biased / removed key pay for the per-key bias (CTA-uniform decision).
        const bool biased = general_bias || ((j + 1) * ATT_BK > kv_len);
        const uint32_t bj = bias_s + (j & 1) * ATT_BK * 4;
        if (biased) {
          const int key = j * ATT_BK + t;
          float bv = NEG_INF;
          if (key < kv_len) bv = general_bias ? a.key_bias[static_cast<size_t>(b) * a.Sk + key] * 1.4426950408889634f : 0.f;
          sts_f32(bj + t * 4, bv);
          asm volatile("bar.sync %0, 128;" ::"r"(1 + x) : "memory");
        }
        mbar_wait(&s_full[x], j & 1);
        tc_fence_after();
        // ---- pass 1: row maximum (scores are re-read from TMEM in pass 2: keeps the register footprint small)
        float mx = NEG_INF;
#pragma unroll 1
        for (int c = 0; c < 4; ++c) {
          uint32_t v[32];
          tmem_ld_x32(tmem + lane_addr + x * 128 + c * 32, v);
          tmem_wait_ld();
          if (biased) {
#pragma unroll
            for (int i = 0; i < 32; ++i) mx = fmaxf(mx, fmaf(__uint_as_float(v[i]), sc, lds_f32(bj + (c * 32 + i) * 4)));
          } else {
            float m0 = __uint_as_float(v[0]), m1 = __uint_as_float(v[1]);
#pragma unroll
            for (int i = 2; i < 32; i += 2) {
              m0 = fmaxf(m0, __uint_as_float(v[i]));
              m1 = fmaxf(m1, __uint_as_float(v[i + 1]));
            }
            mx = fmaxf(mx, fmaxf(m0, m1) * sc);
          }
        }
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
        // ---- pass 2: p = exp2(s * scale - m_use) (+ bias), row sum, fp16 P into the swizzled smem tile
        const float neg_m = -m_use;
        float rs0 = 0.f, rs1 = 0.f;
#pragma unroll 1
        for (int c = 0; c < 4; ++c) {
          uint32_t v[32];
          tmem_ld_x32(tmem + lane_addr + x * 128 + c * 32, v);
          tmem_wait_ld();
          uint32_t pk[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            float t0, t1;
            if (biased) {
              t0 = fmaf(__uint_as_float(v[2 * i]), sc, lds_f32(bj + (c * 32 + 2 * i) * 4)) + neg_m;
              t1 = fmaf(__uint_as_float(v[2 * i + 1]), sc, lds_f32(bj + (c * 32 + 2 * i + 1) * 4)) + neg_m;
            } else {
              t0 = fmaf(__uint_as_float(v[2 * i]), sc, neg_m);
              t1 = fmaf(__uint_as_float(v[2 * i + 1]), sc, neg_m);
            }
            float p0 = fast_exp2(t0), p1 = fast_exp2(t1);
            rs0 += p0;
            rs1 += p1;
            if (DROP) {                               // the row sum (softmax denominator) is taken before dropout
              float m0, m1;
              drop_pair(dbase + ((j * ATT_BK + c * 32) >> 1) + i, dseed, a.drop.thr15, a.drop.scale, m0, m1);
              p0 *= m0;
              p1 *= m1;
            }
            const __half2 hp = __floats2half2_rn(p0, p1);
            pk[i] = *reinterpret_cast<const uint32_t*>(&hp);
          }
          // 32 keys = four 16-byte chunks of this row; chunk index (4c+i) -> half (4c+i)/8, swizzled slot ((4c+i)%8) ^ (r%8)
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int ch = 4 * c + i;
            sts128(p_row + (ch >> 3) * 16384 + (((ch & 7) ^ (r & 7)) << 4), pk[4 * i], pk[4 * i + 1], pk[4 * i + 2], pk[4 * i + 3]);
          }
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
      if (qrow < a.Sq) {
        __half* dst = a.out + (static_cast<size_t>(b) * a.Sq + qrow) * a.ld_out + h * ATT_D;
#pragma unroll
        for (int c = 0; c < 2; ++c)
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            uint32_t w[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const __half2 hv = __floats2half2_rn(__uint_as_float(o[c][8 * i + 2 * k]) * inv_l, __uint_as_float(o[c][8 * i + 2 * k + 1]) * inv_l);
              w[k] = *reinterpret_cast<const uint32_t*>(&hv);
            }
            *reinterpret_cast<uint4*>(dst + c * 32 + i * 8) = make_uint4(w[0], w[1], w[2], w[3]);
          }
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
