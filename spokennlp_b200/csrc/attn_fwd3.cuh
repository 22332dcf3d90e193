// Fused attention forward, production version (same contract as attn_fwd.cuh, which is kept as the first generation for
// A/B runs): head_dim 64, fp16 operands, fp32 softmax statistics, context never materialises P in HBM.
//
// Inside one 128 x 128 score block (what changed against attn_fwd_kernel, profiles/r01d -> r01e):
//   * ONE pass over the scores: a thread pulls its whole 128-key row out of TMEM into registers (4 back-to-back
//     tcgen05.ld, one wait) instead of reading every score twice;
//   * the score buffer is handed back to the tensor core as soon as it sits in registers (`s_free`), so S(j+1) is computed
//     WHILE the softmax of block j runs — the softmax warps never wait for the MMA round trip;
//   * each query tile has its own MMA-issuing warp, and tile B starts half a period after tile A: the MUFU (exp2, the
//     busiest unit at d = 64: 8 clk per warp instruction) and the rest of a block alternate between the two tiles;
//   * masking is a pre-pass that only runs in blocks that contain masked keys (prefix masks: an index compare; general
//     additive bias: staged per block), so the common block carries ~4 instructions per score
//     (FMNMX3/2 + FFMA + MUFU.EX2 + FADD + F2FP/2); dropout tests both lanes of a hash at once (ptx.cuh: drop_z);
//   * a warp's 32 context rows leave as ONE TMA store from a 128B-swizzled staging patch (a thread-per-row st.global
//     touches 32 different 128-byte lines per instruction and kept the LSU busy for ~1.4 us per CTA).
// Across blocks (Sk sweep of the non-persistent form of this kernel, profiles/r01e: 74 us at the bench shape = 41 us of
// per-block work + 33 us of per-CTA fixed cost — launch, TMEM allocation, first TMA round trip, pipeline ramp, output
// drain, all exposed because one CTA owns the SM): ONE CTA per SM stays resident and walks (batch, head, 256-query pair)
// items;
//   * the TMA warp runs ahead across item boundaries (Q double-buffered by item parity, K and V in separate 2-stage rings
//     — K_{j+1} is wanted a whole block earlier than V_{j+1}, so it must not queue behind it),
//   * the MMA warps issue the first S of the next item while the softmax warps are still writing the previous context,
//   * barrier phases come from running counters, so nothing is re-initialised between items.
//   warp 0       TMA producer
//   warp 1, 2    MMA issuers for query tile A / B
//   warps 3-10   softmax A (3-6) / softmax B (7-10): one thread per query row
#pragma once
#include "attn_fwd.cuh"

namespace b200 {

constexpr int ATTP_THREADS = 352;

struct AttnFwd3Smem {
  static constexpr int TILE = ATT_BQ * ATT_D * 2;                    // 16 KB: one Q / K / V tile
  static constexpr int P_BYTES = ATT_BQ * ATT_BK * 2;                // 32 KB per query tile
  static constexpr int OFF_Q = 0;                                    // [2 item parities][2 tiles]
  static constexpr int OFF_K = OFF_Q + 4 * TILE;                     // [2 stages]
  static constexpr int OFF_V = OFF_K + 2 * TILE;                     // [2 stages]
  static constexpr int OFF_P = OFF_V + 2 * TILE;                     // [2 tiles]
  static constexpr int OFF_BIAS = OFF_P + 2 * P_BYTES;               // [2 tiles][2 buffers][128] floats
  static constexpr int OFF_BAR = OFF_BIAS + 2 * 2 * ATT_BK * 4;
  static constexpr int TOTAL = OFF_BAR + 256 + 1024;
};

template <bool DROP>
__global__ void __launch_bounds__(ATTP_THREADS, 1)
attn_fwd3_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmKV,
                 const __grid_constant__ CUtensorMap tmO, const AttnFwdArgs a) {
  using S = AttnFwd3Smem;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + S::OFF_BAR);
  uint64_t* q_full = bars;             // 2
  uint64_t* q_empty = bars + 2;        // 2 (one arrival per MMA warp)
  uint64_t* k_full = bars + 4;         // 2
  uint64_t* k_empty = bars + 6;        // 2 (one arrival per MMA warp)
  uint64_t* v_full = bars + 8;         // 2
  uint64_t* v_empty = bars + 10;       // 2 (one arrival per MMA warp)
  uint64_t* s_full = bars + 12;        // 2 (per tile)
  uint64_t* s_free = bars + 14;        // 2 (one arrival per softmax warp)
  uint64_t* p_full = bars + 16;        // 2 (one arrival per softmax warp)
  uint64_t* o_full = bars + 18;        // 2
  uint64_t* b_go = bars + 20;          // 1
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 21);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  ATT_TRACE_INIT;
  const int nqp = (a.Sq + 2 * ATT_BQ - 1) / (2 * ATT_BQ);
  const int n_items = a.B * a.heads * nqp;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmKV);
    tma_prefetch_desc(&tmO);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&q_full[i], 1);
      mbar_init(&q_empty[i], 2);
      mbar_init(&k_full[i], 1);
      mbar_init(&k_empty[i], 2);
      mbar_init(&v_full[i], 1);
      mbar_init(&v_empty[i], 2);
      mbar_init(&s_full[i], 1);
      mbar_init(&s_free[i], 4);
      mbar_init(&p_full[i], 4);
      mbar_init(&o_full[i], 1);
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

  // per-item geometry, identical in every role
  struct Item {
    int b, h, q0, kv_len, n_blocks;
    int qrow0, krow0, sq;              // first row of this sequence in the Q / KV buffers, its query count
    bool tileB, general_bias, skip;
  };
  auto decode = [&](int item) {
    Item it;
    const int qp = item % nqp, bh = item / nqp;
    it.h = bh % a.heads;
    it.b = bh / a.heads;
    it.q0 = qp * 2 * ATT_BQ;
    if (a.cu_seqlens) {                // packed rows: the sequence's own length bounds queries and keys
      const int r0 = a.cu_seqlens[it.b], len = a.cu_seqlens[it.b + 1] - r0;
      it.qrow0 = it.krow0 = r0;
      it.sq = len;
      it.skip = it.q0 >= len;          // a shorter sequence has fewer query pairs than the longest one
      it.general_bias = false;
      it.kv_len = max(1, len);
    } else {
      it.qrow0 = it.b * a.Sq;
      it.krow0 = it.b * a.Sk;
      it.sq = a.Sq;
      it.skip = false;
      int kv = a.kv_len ? a.kv_len[it.b] : a.Sk;
      it.general_bias = a.key_bias != nullptr && (a.kv_len == nullptr || kv < 0);   // see attn_fwd.cuh
      it.kv_len = max(1, min(kv < 0 ? -kv : kv, a.Sk));
    }
    it.tileB = (it.q0 + ATT_BQ) < it.sq;
    it.n_blocks = (it.kv_len + ATT_BK - 1) / ATT_BK;
    return it;
  };

  if (warp == 0) {
    if (lane == 0) {
      uint32_t n = 0, kb = 0;
      for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++n) {
        const Item it = decode(item);
        if (it.skip) { --n; continue; }          // (the loop header still counts it)
        const uint32_t buf = n & 1;
        mbar_wait(&q_empty[buf], ((n >> 1) & 1) ^ 1);
        mbar_expect_tx(&q_full[buf], (it.tileB ? 2 : 1) * S::TILE);
        uint8_t* qd = smem + S::OFF_Q + buf * 2 * S::TILE;
        tma_load_2d(qd, &tmQ, &q_full[buf], a.q_col0 + it.h * ATT_D, it.qrow0 + it.q0);
        if (it.tileB) tma_load_2d(qd + S::TILE, &tmQ, &q_full[buf], a.q_col0 + it.h * ATT_D, it.qrow0 + it.q0 + ATT_BQ);
        for (int j = 0; j < it.n_blocks; ++j, ++kb) {
          const uint32_t s = kb & 1, ph = (kb >> 1) & 1;
          mbar_wait(&k_empty[s], ph ^ 1);
          mbar_expect_tx(&k_full[s], S::TILE);
          tma_load_2d(smem + S::OFF_K + s * S::TILE, &tmKV, &k_full[s], a.k_col0 + it.h * ATT_D, it.krow0 + j * ATT_BK);
          mbar_wait(&v_empty[s], ph ^ 1);
          mbar_expect_tx(&v_full[s], S::TILE);
          tma_load_2d(smem + S::OFF_V + s * S::TILE, &tmKV, &v_full[s], a.v_col0 + it.h * ATT_D, it.krow0 + j * ATT_BK);
        }
      }
    }
    __syncwarp();
  } else if (warp <= 2) {
    // ------------------------------------------------------------------ MMA issuer of query tile x
    const int x = warp - 1;
    constexpr uint32_t idesc_s = make_idesc_f16(128, ATT_BK, 0, 0);
    constexpr uint32_t idesc_o = make_idesc_f16(128, ATT_D, 0, 1);
    const uint32_t pa = smem_u32(smem + S::OFF_P + x * S::P_BYTES);
    uint32_t n = 0, kb = 0, t = 0;     // items seen, K/V ring position, blocks this tile has processed
    for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++n) {
      const Item it = decode(item);
      if (it.skip) { --n; continue; }
      const uint32_t buf = n & 1;
      const int nb = it.n_blocks;
      if (x == 1 && !it.tileB) {       // no second tile in this item: keep the shared rings' arrival counts balanced
        mbar_wait(&q_full[buf], (n >> 1) & 1);
        if (lane == 0) mbar_arrive(&q_empty[buf]);
        for (int j = 0; j < nb; ++j, ++kb) {
          const uint32_t s = kb & 1, ph = (kb >> 1) & 1;
          mbar_wait(&k_full[s], ph);
          if (lane == 0) mbar_arrive(&k_empty[s]);
          mbar_wait(&v_full[s], ph);
          if (lane == 0) mbar_arrive(&v_empty[s]);
        }
        __syncwarp();
        continue;
      }
      const uint32_t qa = smem_u32(smem + S::OFF_Q + (buf * 2 + x) * S::TILE);
      auto issue_s = [&](uint32_t s, bool last) {          // S_x = Q_x K^T from K ring stage s
        const uint32_t ka = smem_u32(smem + S::OFF_K + s * S::TILE);
#pragma unroll
        for (int kk = 0; kk < ATT_D / 16; ++kk)
          umma_ss(tmem + x * 128, make_smem_desc(qa + kk * 32, 0, 1024), make_smem_desc(ka + kk * 32, 0, 1024), idesc_s, kk > 0);
        umma_commit(&s_full[x]);
        umma_commit(&k_empty[s]);
        ATT_TRACE(10);                            // S issued
        if (last) umma_commit(&q_empty[buf]);                // Q of this item is not needed after its last score block
      };
      mbar_wait(&q_full[buf], (n >> 1) & 1);
      mbar_wait(&k_full[kb & 1], (kb >> 1) & 1);
      if (t > 0) mbar_wait(&s_free[x], (t - 1) & 1);         // the previous block's score rows sit in registers
      else if (x == 1) mbar_wait(b_go, 0);                   // tile B starts half a period after tile A (header)
      tc_fence_after();
      if (elect_one_sync()) issue_s(kb & 1, nb == 1);
      __syncwarp();
      for (int j = 0; j < nb; ++j, ++t, ++kb) {
        const uint32_t s = kb & 1, ph = (kb >> 1) & 1;
        if (j + 1 < nb) {                                    // next scores as soon as the softmax threads hold block j
          const uint32_t s1 = (kb + 1) & 1, ph1 = ((kb + 1) >> 1) & 1;
          mbar_wait(&k_full[s1], ph1);
          mbar_wait(&s_free[x], t & 1);
          tc_fence_after();
          if (elect_one_sync()) issue_s(s1, j + 2 == nb);
          __syncwarp();
        }
        mbar_wait(&p_full[x], t & 1);
        mbar_wait(&v_full[s], ph);
        tc_fence_after();
        if (elect_one_sync()) {
          const uint32_t va = smem_u32(smem + S::OFF_V + s * S::TILE);
#pragma unroll
          for (int kk = 0; kk < ATT_BK / 16; ++kk)
            umma_ss(tmem + 256 + x * 64, make_smem_desc(pa + (kk >> 2) * 16384 + (kk & 3) * 32, 0, 1024),
                    make_smem_desc(va + kk * 2048, 8192, 1024), idesc_o, (j > 0 || kk > 0) ? 1u : 0u);
          umma_commit(&o_full[x]);
          umma_commit(&v_empty[s]);
          ATT_TRACE(11);                          // P.V issued
        }
        __syncwarp();
      }
    }
  } else {
    // ------------------------------------------------------------------ softmax groups
    const int x = (warp - 3) >> 2;                // 0 = tile A, 1 = tile B
    const int qd = warp & 3;                      // TMEM lane quadrant of this warp
    const int r = qd * 32 + lane;                 // row inside the query tile
    const int tg = ((warp - 3) & 3) * 32 + lane;  // 0..127 inside the group (bias staging)
    const uint32_t lane_addr = static_cast<uint32_t>(qd * 32) << 16;
    const uint32_t bias_s = smem_u32(smem + S::OFF_BIAS) + x * 2 * ATT_BK * 4;
    const uint32_t p_row = smem_u32(smem + S::OFF_P + x * S::P_BYTES) + r * 128;
    const float NEG_INF = -INFINITY;
    constexpr float sc = kAttScaleLog2;           // log2(e) / sqrt(64): an immediate operand of the packed FMA below
    constexpr float inv_sc = 1.0f / kAttScaleLog2;
    const uint32_t dseed = DROP ? drop_seed(a.drop) : 0u;
    uint32_t t = 0;                               // blocks this tile has processed (barrier phases)
    if (x == 1) asm volatile("bar.arrive 3, 256;" ::: "memory");      // tile A opens
    for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
      const Item it = decode(item);
      if (it.skip || (x == 1 && !it.tileB)) continue;
      const int b = it.b, h = it.h, kv_len = it.kv_len, n_blocks = it.n_blocks;
      const int qrow = it.q0 + x * ATT_BQ + r;
      // dropout: quad index of (row, key) = rowbase + key / 4; (quad + seed) * C1 is walked by adding multiples of C1 (ptx.cuh: drop4_z)
      const uint32_t dpre = DROP ? drop_premix(static_cast<uint32_t>(((static_cast<size_t>(b) * a.heads + h) * a.Sq + min(qrow, a.Sq - 1)) * ((a.Sk + 3) >> 2)), dseed) : 0u;
      const uint32_t dtt = a.drop.thr15;          // the four packed lane thresholds
      float m = NEG_INF, l = 0.f;
      if (lane == 0) tma_wait_group_read<0>();    // the previous item's context store has read this warp's staging rows
      __syncwarp();
      for (int j = 0; j < n_blocks; ++j, ++t) {
        const bool partial = (j + 1) * ATT_BK > kv_len;
        const uint32_t bj = bias_s + (t & 1) * ATT_BK * 4;
        if (it.general_bias) {                    // bias in units of raw scores: (s + bias/scale) * scale = s*scale + bias
          const int key = j * ATT_BK + tg;
          float bv = NEG_INF;
          if (key < kv_len) bv = a.key_bias[static_cast<size_t>(b) * a.Sk + key] * 1.4426950408889634f * inv_sc;
          sts_f32(bj + tg * 4, bv);
          asm volatile("bar.sync %0, 128;" ::"r"(1 + x) : "memory");
        }
        ATT_TRACE(1);                             // waiting for S
        mbar_wait(&s_full[x], t & 1);
        tc_fence_after();
        ATT_TRACE(2);                             // S arrived
        uint32_t v[128];
        tmem_ld_x32(tmem + lane_addr + x * 128, *reinterpret_cast<uint32_t(*)[32]>(&v[0]));
        tmem_ld_x32(tmem + lane_addr + x * 128 + 32, *reinterpret_cast<uint32_t(*)[32]>(&v[32]));
        tmem_ld_x32(tmem + lane_addr + x * 128 + 64, *reinterpret_cast<uint32_t(*)[32]>(&v[64]));
        tmem_ld_x32(tmem + lane_addr + x * 128 + 96, *reinterpret_cast<uint32_t(*)[32]>(&v[96]));
        tmem_wait_ld();
        tc_fence_before();
        __syncwarp();                             // the tensor core may overwrite S_x with the next block now:
        if (lane == 0) mbar_arrive(&s_free[x]);   // one arrival per warp (128 same-address arrivals cost 9 us per backward, r02a)
        if (x == 0 && t == 0 && tg == 0) mbar_arrive(b_go);
        ATT_TRACE(3);                             // scores in registers
        // ---- masked keys (only blocks that have any)
        if (it.general_bias) {
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
        ATT_TRACE(4);                             // row max done, waiting for O(j-1)
        if (j > 0) {                              // P.V of block j-1 finished: P smem is free, O may be rescaled
          mbar_wait(&o_full[x], (t - 1) & 1);
          tc_fence_after();
          ATT_TRACE(rescale ? 6 : 5);             // O(j-1) arrived (6: this block rescales)
          if (rescale) {
#pragma unroll
            for (int c = 0; c < 4; ++c) {
              uint32_t o[16];
              tmem_ld_x16(tmem + lane_addr + 256 + x * 64 + c * 16, o);
              tmem_wait_ld();
              const uint64_t al2 = pack2(alpha, alpha);
#pragma unroll
              for (int i = 0; i < 16; i += 2) {
                float r0, r1;
                unpack2(mul2(pack2u(o[i], o[i + 1]), al2), r0, r1);
                o[i] = __float_as_uint(r0);
                o[i + 1] = __float_as_uint(r1);
              }
              tmem_st_x16(tmem + lane_addr + 256 + x * 64 + c * 16, o);
            }
            tmem_wait_st();
          }
        }
        // ---- the two tiles take turns in the exp2 section (MUFU-bound: two warps of a sub-partition inside it at once each take
        //      twice as long, and then both sit in the latency-bound rest of a block — barrier waits, TMEM loads — together;
        //      profiles/r02m_trace_fwd.txt: 2850 clk for the section against ~1300 alone, 4400 per block).  A token passes
        //      A -> B -> A -> ... through two named barriers; both tiles of an item have the same number of key blocks.
        if (it.tileB) asm volatile("bar.sync %0, 256;" ::"r"(3 + x) : "memory");
        // ---- p = exp2(s * scale - m_use), row sum, fp16 P into the swizzled smem tile (16 chunks of 8 keys)
        const uint64_t nm2 = pack2(-m_use, -m_use), sc2 = pack2(sc, sc);
        const uint32_t dpre_j = dpre + static_cast<uint32_t>(j * (ATT_BK / 4)) * kDropC1;
        uint64_t rs2 = pack2(0.f, 0.f);
#pragma unroll
        for (int ch = 0; ch < 16; ++ch) {
          uint32_t pk[4], z[2];
          if (DROP) {                               // this chunk's 8 keys = quads 2 ch, 2 ch + 1 of the block
            z[0] = drop4_z(dpre_j + static_cast<uint32_t>(2 * ch) * kDropC1, dtt);
            z[1] = drop4_z(dpre_j + static_cast<uint32_t>(2 * ch + 1) * kDropC1, dtt);
          }
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int i = ch * 8 + 2 * e;
            float x0, x1;
            unpack2(fma2(pack2u(v[i], v[i + 1]), sc2, nm2), x0, x1);        // two exponents per instruction
            const float p0 = fast_exp2(x0), p1 = fast_exp2(x1);
            rs2 = add2(rs2, pack2(p0, p1));         // the row sum (softmax denominator) is taken before dropout
            const __half2 hp = __floats2half2_rn(p0, p1);
            pk[e] = *reinterpret_cast<const uint32_t*>(&hp);
            if (DROP)                               // zero the dropped lanes of the packed pair; 1/(1-p) is applied to O at the end
              pk[e] &= (e & 1) ? drop4_keep_h2_hi(z[e >> 1]) : drop4_keep_h2_lo(z[e >> 1]);
          }
          sts128(p_row + (ch >> 3) * 16384 + (((ch & 7) ^ (r & 7)) << 4), pk[0], pk[1], pk[2], pk[3]);
        }
        float rs0, rs1;
        unpack2(rs2, rs0, rs1);
        l = fmaf(l, alpha, rs0 + rs1);
        m = (m_use == 0.f && m_new == NEG_INF) ? NEG_INF : m_use;
        fence_proxy_async_smem();
        tc_fence_before();
        __syncwarp();                             // every lane has fenced its own P rows; one lane publishes the warp's 32 rows
        if (lane == 0) mbar_arrive(&p_full[x]);
        if (it.tileB) asm volatile("bar.arrive %0, 256;" ::"r"(3 + (x ^ 1)) : "memory");      // the other tile's turn
        ATT_TRACE(7);                             // P published
      }
      // ---------------------------------------------------------------- finalise: O / l -> ctx, LSE
      mbar_wait(&o_full[x], (t - 1) & 1);
      tc_fence_after();
      ATT_TRACE(8);                               // last O arrived
      const float inv_l = (l > 0.f ? 1.0f / l : 0.f) * (DROP ? a.drop.scale : 1.0f);
      uint32_t o[2][32];
      tmem_ld_x32(tmem + lane_addr + 256 + x * 64, o[0]);
      tmem_ld_x32(tmem + lane_addr + 256 + x * 64 + 32, o[1]);
      tmem_wait_ld();
      tc_fence_before();
      uint32_t w[32];
#pragma unroll
      for (int k = 0; k < 32; ++k) {
        const __half2 hv = __floats2half2_rn(__uint_as_float(o[k >> 4][2 * (k & 15)]) * inv_l, __uint_as_float(o[k >> 4][2 * (k & 15) + 1]) * inv_l);
        w[k] = *reinterpret_cast<const uint32_t*>(&hv);
      }
      const int wrow0 = it.q0 + x * ATT_BQ + qd * 32;        // first query row of this warp
      if (wrow0 + 32 <= it.sq) {                             // one TMA store per warp from a swizzled staging patch (never across a sequence end)
#pragma unroll
        for (int ch = 0; ch < 8; ++ch) sts128(p_row + ((ch ^ (r & 7)) << 4), w[4 * ch], w[4 * ch + 1], w[4 * ch + 2], w[4 * ch + 3]);
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) {
          tma_store_2d(&tmO, smem + S::OFF_P + x * S::P_BYTES + qd * 32 * 128, h * ATT_D, it.qrow0 + wrow0);
          tma_commit_group();
        }
      } else if (qrow < it.sq) {
        __half* dst = a.out + (static_cast<size_t>(it.qrow0) + qrow) * a.ld_out + h * ATT_D;
#pragma unroll
        for (int ch = 0; ch < 8; ++ch) *reinterpret_cast<uint4*>(dst + ch * 8) = make_uint4(w[4 * ch], w[4 * ch + 1], w[4 * ch + 2], w[4 * ch + 3]);
      }
      if (qrow < it.sq && a.lse2) a.lse2[(static_cast<size_t>(b) * a.heads + h) * a.Sq + qrow] = (l > 0.f) ? (m + log2f(l)) : NEG_INF;
      ATT_TRACE(9);                               // item finalised
    }
    if (lane == 0) tma_wait_group_read<0>();      // staging memory stays valid until the bulk stores have read it
  }

  ATT_TRACE_FINI;
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem, 512);
  }
}

}  // namespace b200
