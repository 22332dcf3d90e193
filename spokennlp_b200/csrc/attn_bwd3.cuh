// Fused attention backward, production version (same contract and workspace as attn_bwd.cuh, which is kept as the first
// generation for A/B runs) — the autograd of bert_model.py:309-350 for head_dim 64.
//
// One work item = one 128-key block of one (batch, head), walked over the 128-query blocks i:
//     S  = Q_i K^T , dP = dO_i V^T                     (K-major operands, M = queries, N = 64 keys per half)
//     P  = exp2(S c + bias_k - lse2_q),  dS = P o (dP - delta_q) / sqrt(d)          [CUDA cores, row per thread]
//     dV += P^T  dO_i   (A = P  tile viewed MN-major, B = dO_i MN-major)            accumulates in TMEM over i
//     dK += dS^T Q_i    (A = dS tile viewed MN-major, B = Q_i  MN-major)            accumulates in TMEM over i
//     dQ_i = dS K       (A = dS K-major, B = K MN-major) -> fp32 TMA reduce-add into dq_acc
// Inside an item (what changed against attn_bwd_kernel, profiles/r01d -> r01f):
//   * natural orientation, TMEM lane == QUERY row: lse2 / delta are two registers per thread instead of two shared-memory
//     loads per score, and the dropout mask is regenerated pairwise exactly as the forward drew it;
//   * the score tile is split by KEY half between the two softmax groups, each with its own score buffers and barriers:
//     S_g(i+1) and dP_g(i+1) are issued as soon as group g holds block i in registers, so the tensor core recomputes
//     scores while the CUDA cores do the exp / dS math; group 1 starts half a period after group 0;
//   * a 3-deep Q / dO ring keeps the operands of block i+1 resident that early;
//   * dQ leaves as ONE TMA reduce-add per warp from a 128B-swizzled staging patch (8 scattered red.global.add.v4 per
//     thread kept the LSU busy for ~2k cycles per block and stalled the next block on the register hazard); rows past Sq
//     carry exact zeros (P = 0 there), so the box needs no row clipping inside a batch element.
// Across items (Sk sweep, profiles/r01e: ~60 of 150 us were per-CTA fixed cost paid 10.4 times per SM): ONE CTA per SM
// stays resident and walks (batch, head, key-block) items;
//   * the TMA warp fetches Q_0 / dO_0 of the next item while the current one is still running and its K / V as soon as the
//     last gradient MMAs have retired — that latency hides behind the dK / dV drain of the softmax warps,
//   * barrier phases, the Q / dO ring position and the dQ ping-pong buffer come from running counters,
//   * dK / dV leave through the dQ staging patches (free at that point) as two TMA stores, so the P / dS tiles never have
//     to be protected across an item boundary.
//   warp 0 TMA producer, warp 1 MMA issuer, warps 2-5 softmax group 0 (keys 0-63), warps 6-9 group 1 (keys 64-127)
#pragma once
#include "attn_bwd.cuh"

namespace b200 {

constexpr int ATTB3_QDO_STAGES = 3;

struct AttnBwd3Smem {
  static constexpr int T = ATT_BK * ATT_D * 2;                 // 16 KB tile
  static constexpr int OFF_K = 0;
  static constexpr int OFF_V = OFF_K + T;
  static constexpr int OFF_QDO = OFF_V + T;                    // 3 stages x (Q_i, dO_i)
  static constexpr int OFF_P = OFF_QDO + ATTB3_QDO_STAGES * 2 * T;   // P  [128 q][128 keys] fp16 = 32 KB (two 64-key halves)
  static constexpr int OFF_DS = OFF_P + 2 * T;                 // dS 32 KB
  static constexpr int OFF_DQS = OFF_DS + 2 * T;               // dQ staging for the TMA reduction: 8 warps x [32 q][32 d] fp32 (4 KB each)
  static constexpr int OFF_BIAS = OFF_DQS + 8 * 4096;          // [128] floats: additive bias of this item's keys (log2 units)
  static constexpr int OFF_BAR = OFF_BIAS + ATT_BK * 4;
  static constexpr int TOTAL = OFF_BAR + 256 + 1024;
};

template <bool DROP>
__global__ void __launch_bounds__(ATTB_THREADS, 1)
attn_bwd3_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmKV,
                 const __grid_constant__ CUtensorMap tmDO, const __grid_constant__ CUtensorMap tmDQ,
                 const __grid_constant__ CUtensorMap tmDKV, const AttnBwdArgs a) {
  using S = AttnBwd3Smem;
  constexpr int NST = ATTB3_QDO_STAGES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + S::OFF_BAR);
  uint64_t* kv_full = bars;                   // 1
  uint64_t* kv_empty = bars + 1;              // 1
  uint64_t* qdo_full = bars + 2;              // 3
  uint64_t* qdo_empty = qdo_full + NST;       // 3
  uint64_t* s_full = qdo_empty + NST;         // 2 (one per key half)
  uint64_t* s_free = s_full + 2;              // 2 (one arrival per softmax warp of the half)
  uint64_t* ds_full = s_free + 2;             // 1 (one arrival per softmax warp)
  uint64_t* grad_done = ds_full + 1;          // 1
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(grad_done + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  ATT_TRACE_INIT;
  const int n_kb = (a.Sk + ATT_BK - 1) / ATT_BK;
  const int n_items = a.B * a.heads * n_kb;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmKV);
    tma_prefetch_desc(&tmDO);
    tma_prefetch_desc(&tmDQ);
    tma_prefetch_desc(&tmDKV);
    mbar_init(kv_full, 1);
    mbar_init(kv_empty, 1);
    for (int i = 0; i < NST; ++i) {
      mbar_init(&qdo_full[i], 1);
      mbar_init(&qdo_empty[i], 1);
    }
    for (int g = 0; g < 2; ++g) {
      mbar_init(&s_full[g], 1);
      mbar_init(&s_free[g], 4);
    }
    mbar_init(ds_full, 8);
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
  // TMEM columns: S [0,128) (half g at 64g)  dP [128,256)  dV [256,320)  dK [320,384)  dQ ping-pong [384,448) / [448,512)

  struct Item {
    int b, h, k0, kv_len;
    int qrow0, krow0, sq, sk, nq;     // first row of this sequence in the Q / KV buffers, its query / key counts, query blocks
    bool general_bias, dead;
  };
  auto decode = [&](int item) {
    Item it;
    const int kbk = item % n_kb, bh = item / n_kb;
    it.h = bh % a.heads;
    it.b = bh / a.heads;
    it.k0 = kbk * ATT_BK;
    if (a.cu_seqlens) {               // packed rows: the sequence's own length bounds queries and keys
      const int r0 = a.cu_seqlens[it.b], len = a.cu_seqlens[it.b + 1] - r0;
      it.qrow0 = it.krow0 = r0;
      it.sq = it.sk = len;
      it.general_bias = false;
      it.kv_len = max(1, len);
    } else {
      it.qrow0 = it.b * a.Sq;
      it.krow0 = it.b * a.Sk;
      it.sq = a.Sq;
      it.sk = a.Sk;
      int kv = a.kv_len ? a.kv_len[it.b] : a.Sk;
      it.general_bias = a.key_bias != nullptr && (a.kv_len == nullptr || kv < 0);   // see attn_fwd.cuh
      it.kv_len = max(1, min(kv < 0 ? -kv : kv, a.Sk));
    }
    it.nq = (it.sq + ATT_BQ - 1) / ATT_BQ;
    it.dead = it.k0 >= it.kv_len || it.nq == 0;      // every key of this block is masked: dK = dV = 0, no dQ contribution
    return it;
  };

  if (warp == 0) {
    if (lane == 0) {
      uint32_t n = 0, qs = 0;           // live items seen, Q / dO ring position
      for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
        const Item it = decode(item);
        if (it.dead) continue;
        auto load_qdo = [&](int i) {
          const uint32_t st = qs % NST, ph = (qs / NST) & 1;
          mbar_wait(&qdo_empty[st], ph ^ 1);
          mbar_expect_tx(&qdo_full[st], 2 * S::T);
          uint8_t* dst = smem + S::OFF_QDO + st * 2 * S::T;
          tma_load_2d(dst, &tmQ, &qdo_full[st], a.q_col0 + it.h * ATT_D, it.qrow0 + i * ATT_BQ);
          tma_load_2d(dst + S::T, &tmDO, &qdo_full[st], it.h * ATT_D, it.qrow0 + i * ATT_BQ);
          ++qs;
        };
        load_qdo(0);                    // the first Q / dO block does not wait for the item switch ...
        mbar_wait(kv_empty, (n & 1) ^ 1);   // ... K / V do: the previous item's last gradient MMAs read them
        mbar_expect_tx(kv_full, 2 * S::T);
        tma_load_2d(smem + S::OFF_K, &tmKV, kv_full, a.k_col0 + it.h * ATT_D, it.krow0 + it.k0);
        tma_load_2d(smem + S::OFF_V, &tmKV, kv_full, a.v_col0 + it.h * ATT_D, it.krow0 + it.k0);
        for (int i = 1; i < it.nq; ++i) load_qdo(i);
        ++n;
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    constexpr uint32_t idesc_s = make_idesc_f16(128, 64, 0, 0);     // S_g, dP_g : both K-major, N = 64 keys
    constexpr uint32_t idesc_t = make_idesc_f16(128, 64, 1, 1);     // dV, dK   : A and B MN-major
    constexpr uint32_t idesc_q = make_idesc_f16(128, 64, 0, 1);     // dQ       : A K-major, B MN-major
    const uint32_t ka = smem_u32(smem + S::OFF_K), va = smem_u32(smem + S::OFF_V);
    const uint32_t pa = smem_u32(smem + S::OFF_P), dsa = smem_u32(smem + S::OFF_DS);
    auto issue_scores = [&](uint32_t st, int g) {
      const uint32_t qa = smem_u32(smem + S::OFF_QDO + st * 2 * S::T), doa = qa + S::T;
      const uint32_t kg = ka + g * 64 * 128, vg = va + g * 64 * 128;            // rows [64g, 64g+64) of the K / V tiles
#pragma unroll
      for (int kk = 0; kk < 4; ++kk)
        umma_ss(tmem + g * 64, make_smem_desc(qa + kk * 32, 0, 1024), make_smem_desc(kg + kk * 32, 0, 1024), idesc_s, kk > 0);
#pragma unroll
      for (int kk = 0; kk < 4; ++kk)
        umma_ss(tmem + 128 + g * 64, make_smem_desc(doa + kk * 32, 0, 1024), make_smem_desc(vg + kk * 32, 0, 1024), idesc_s, kk > 0);
      umma_commit(&s_full[g]);
      ATT_TRACE(10 + g);                          // S_g / dP_g issued
    };
    uint32_t n = 0, qs = 0, ir = 0;     // live items seen, Q / dO ring position, query blocks processed so far
    for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
      const Item it = decode(item);
      if (it.dead) continue;
      const int nq = it.nq;
      mbar_wait(kv_full, n & 1);
      mbar_wait(&qdo_full[qs % NST], (qs / NST) & 1);
      if (ir > 0) mbar_wait(&s_free[0], (ir - 1) & 1);      // group 0 holds its last block in registers
      tc_fence_after();
      if (elect_one_sync()) issue_scores(qs % NST, 0);
      __syncwarp();
      for (int i = 0; i < nq; ++i, ++ir, ++qs) {
        const uint32_t st = qs % NST;
        if (i == 0) {
          // group 1's first scores of the item.  At the very start of the kernel they are held back until group 0 has pulled
          // its own block out of TMEM, so that the two groups run half a period apart (header).
          if (ir > 0) mbar_wait(&s_free[1], (ir - 1) & 1);
          else mbar_wait(&s_free[0], 0);
          tc_fence_after();
          if (elect_one_sync()) issue_scores(st, 1);
          __syncwarp();
        }
        if (i + 1 < nq) {                     // scores of the next block as soon as a group holds block i in registers
          const uint32_t nst = (qs + 1) % NST;
          mbar_wait(&qdo_full[nst], ((qs + 1) / NST) & 1);
          for (int g = 0; g < 2; ++g) {
            mbar_wait(&s_free[g], ir & 1);
            tc_fence_after();
            if (elect_one_sync()) issue_scores(nst, g);
            __syncwarp();
          }
        }
        mbar_wait(ds_full, ir & 1);
        tc_fence_after();
        ATT_TRACE(13);                            // P / dS arrived
        if (elect_one_sync()) {
          const uint32_t qa = smem_u32(smem + S::OFF_QDO + st * 2 * S::T), doa = qa + S::T;
          {
#pragma unroll
            for (int kk = 0; kk < 8; ++kk)        // dV += P^T dO_i      (K = 16 query rows per step)
              umma_ss(tmem + 256, make_smem_desc(pa + kk * 2048, 16384, 1024), make_smem_desc(doa + kk * 2048, 8192, 1024), idesc_t,
                      (i > 0 || kk > 0) ? 1u : 0u);
#pragma unroll
            for (int kk = 0; kk < 8; ++kk)        // dK += dS^T Q_i
              umma_ss(tmem + 320, make_smem_desc(dsa + kk * 2048, 16384, 1024), make_smem_desc(qa + kk * 2048, 8192, 1024), idesc_t,
                      (i > 0 || kk > 0) ? 1u : 0u);
#pragma unroll
            for (int kk = 0; kk < 8; ++kk)        // dQ_i = dS K         (K = 16 keys per step)
              umma_ss(tmem + 384 + (ir & 1) * 64, make_smem_desc(dsa + (kk >> 2) * 16384 + (kk & 3) * 32, 0, 1024),
                      make_smem_desc(ka + kk * 2048, 8192, 1024), idesc_q, kk > 0);
          }
          umma_commit(&qdo_empty[st]);
          umma_commit(grad_done);
          ATT_TRACE(12);                          // gradient MMAs issued
          if (i + 1 == nq) umma_commit(kv_empty);             // K / V of this item are dead once these MMAs retire
        }
        __syncwarp();
      }
      ++n;
    }
  } else {
    const int qd = warp & 3;                      // TMEM lane quadrant
    const int g = (warp - 2) >> 2;                // key half (score columns [64g, 64g+64)); also which 32 of the 64 d columns
    const int r = qd * 32 + lane;                 // query row inside the block == TMEM lane (key row for dV / dK)
    const int t = threadIdx.x - 64;               // 0..255
    const uint32_t lane_addr = static_cast<uint32_t>(qd * 32) << 16;
    const uint32_t bias_s = smem_u32(smem + S::OFF_BIAS);
    const uint32_t p_row = smem_u32(smem + S::OFF_P) + g * 16384 + r * 128;
    const uint32_t ds_row = smem_u32(smem + S::OFF_DS) + g * 16384 + r * 128;
    const uint32_t dseed = DROP ? drop_seed(a.drop) : 0u;
    const uint32_t skq = static_cast<uint32_t>((a.Sk + 3) >> 2);       // dropout quads per score row (ptx.cuh: drop4_z)
    constexpr float sc = kAttScaleLog2;
    const uint32_t dtt = a.drop.thr15;            // the four packed lane thresholds
    const uint32_t cs_bits = __float_as_uint(a.inv_sqrt_d * a.drop.scale);   // dP coefficient of a kept element
    uint8_t* dqs = smem + S::OFF_DQS + (warp - 2) * 4096;     // this warp's [32 q][32 d] fp32 staging patch
    const uint32_t dqs_row = smem_u32(dqs) + lane * 128;
    uint32_t ir = 0;                              // query blocks processed so far (barrier phases, dQ buffer)
    for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
      const Item it = decode(item);
      const int b = it.b, h = it.h, k0 = it.k0, kv_len = it.kv_len, nq = it.nq;
      if (it.dead) {
        if (warp < 6 && k0 + r < it.sk) {
          const size_t row = static_cast<size_t>(it.krow0) + k0 + r;
          const uint4 z = make_uint4(0, 0, 0, 0);
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            *reinterpret_cast<uint4*>(a.dk + row * a.ld_dkv + a.dk_col0 + h * ATT_D + i * 8) = z;
            *reinterpret_cast<uint4*>(a.dv + row * a.ld_dkv + a.dv_col0 + h * ATT_D + i * 8) = z;
          }
        }
        continue;
      }
      const size_t stat_base = (static_cast<size_t>(b) * a.heads + h) * a.Sq;
      const int kg0 = k0 + g * 64;                // first key of this thread's columns
      const int lim = kv_len - kg0;               // keys [0, lim) of this half are kept (prefix masks)
      const bool partial = lim < 64;
      if (it.general_bias) {                      // this item's 128 bias values (the barrier also fences the previous item's reads)
        asm volatile("bar.sync 1, 256;" ::: "memory");
        if (t < ATT_BK) {
          const int key = k0 + t;
          float bv = -INFINITY;
          if (key < kv_len) bv = a.key_bias[static_cast<size_t>(b) * a.Sk + key] * 1.4426950408889634f;
          sts_f32(bias_s + t * 4, bv);
        }
        asm volatile("bar.sync 1, 256;" ::: "memory");
      }

      // dQ tile of block (ir_blk): TMEM lane == query row; one TMA reduce-add per warp from its swizzled staging patch
      auto drain_dq = [&](uint32_t ir_blk, int i) {
        uint32_t o[32];
        tmem_ld_x32(tmem + lane_addr + 384 + (ir_blk & 1) * 64 + g * 32, o);
        tmem_wait_ld();
        if (lane == 0) tma_wait_group_read<0>();   // the previous reduction has finished reading the staging patch
        __syncwarp();
        if (a.dq_half) {
          // fp16 reduce-add into the dQ columns of the gradient buffer: [32 q][32 d] halves = 64-byte rows, SWIZZLE_64B (16-byte
          // chunk c of row r sits at chunk c ^ ((r >> 1) & 3)).  At most Sk / 128 partial sums meet per element (4 at S = 512).
          uint32_t w[16];
#pragma unroll
          for (int k = 0; k < 16; ++k) {
            const __half2 hv = __floats2half2_rn(__uint_as_float(o[2 * k]), __uint_as_float(o[2 * k + 1]));
            w[k] = *reinterpret_cast<const uint32_t*>(&hv);
          }
          const uint32_t row_s = smem_u32(dqs) + lane * 64;
#pragma unroll
          for (int ch = 0; ch < 4; ++ch) sts128(row_s + ((ch ^ ((lane >> 1) & 3)) << 4), w[4 * ch], w[4 * ch + 1], w[4 * ch + 2], w[4 * ch + 3]);
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) {
            tma_reduce_add_2d(&tmDQ, dqs, a.dq_col0 + h * ATT_D + g * 32, it.qrow0 + i * ATT_BQ + qd * 32);
            tma_commit_group();
          }
          return;
        }
#pragma unroll
        for (int ch = 0; ch < 8; ++ch) sts128(dqs_row + ((ch ^ (lane & 7)) << 4), o[4 * ch], o[4 * ch + 1], o[4 * ch + 2], o[4 * ch + 3]);
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) {
          tma_reduce_add_2d(&tmDQ, dqs, h * ATT_D + g * 32, it.qrow0 + i * ATT_BQ + qd * 32);
          tma_commit_group();
        }
      };

      // per-query statistics of block 0 (queries past Sq: lse = +inf -> P = 0, delta = 0)
      float lse_n = (r < it.sq) ? a.lse2[stat_base + r] : INFINITY;
      float del_n = (r < it.sq) ? a.delta[stat_base + r] : 0.f;
      for (int i = 0; i < nq; ++i, ++ir) {
        const float neg_lse = -lse_n, ndc = -del_n * a.inv_sqrt_d;
        const int q = i * ATT_BQ + r;
        if (i + 1 < nq) {                         // prefetch the next block's statistics
          const int qn = q + ATT_BQ;
          lse_n = (qn < it.sq) ? a.lse2[stat_base + qn] : INFINITY;
          del_n = (qn < it.sq) ? a.delta[stat_base + qn] : 0.f;
        }
        ATT_TRACE(1);                             // waiting for S / dP
        mbar_wait(&s_full[g], ir & 1);
        tc_fence_after();
        ATT_TRACE(2);                             // S / dP arrived
        uint32_t sv[64], dp[64];
        tmem_ld_x32(tmem + lane_addr + g * 64, *reinterpret_cast<uint32_t(*)[32]>(&sv[0]));
        tmem_ld_x32(tmem + lane_addr + g * 64 + 32, *reinterpret_cast<uint32_t(*)[32]>(&sv[32]));
        tmem_ld_x32(tmem + lane_addr + 128 + g * 64, *reinterpret_cast<uint32_t(*)[32]>(&dp[0]));
        tmem_ld_x32(tmem + lane_addr + 128 + g * 64 + 32, *reinterpret_cast<uint32_t(*)[32]>(&dp[32]));
        tmem_wait_ld();
        tc_fence_before();
        __syncwarp();                             // the tensor core may overwrite S_g / dP_g with the next block now
        if (lane == 0) mbar_arrive(&s_free[g]);
        ATT_TRACE(3);                             // scores in registers
        // masked keys: turn their scores into -inf (P = 0, dS = 0)
        if (it.general_bias) {
#pragma unroll
          for (int c = 0; c < 64; c += 4) {
            const uint4 bb = lds128(bias_s + (g * 64 + c) * 4);
            sv[c] = __float_as_uint(fmaf(__uint_as_float(bb.x), 1.0f / sc, __uint_as_float(sv[c])));
            sv[c + 1] = __float_as_uint(fmaf(__uint_as_float(bb.y), 1.0f / sc, __uint_as_float(sv[c + 1])));
            sv[c + 2] = __float_as_uint(fmaf(__uint_as_float(bb.z), 1.0f / sc, __uint_as_float(sv[c + 2])));
            sv[c + 3] = __float_as_uint(fmaf(__uint_as_float(bb.w), 1.0f / sc, __uint_as_float(sv[c + 3])));
          }
        } else if (partial) {
#pragma unroll
          for (int c = 0; c < 64; ++c) sv[c] = (c < lim) ? sv[c] : 0xff800000u;
        }
        uint32_t pk[32], dk[32];                  // this thread's 64 P / dS values, packed fp16
        // dropout: (quad + seed) * C1 of this row's first quad; consecutive quads add C1 (ptx.cuh: drop4_z)
        const uint32_t dpre = DROP ? drop_premix(static_cast<uint32_t>(stat_base + min(q, a.Sq - 1)) * skq + (static_cast<uint32_t>(kg0) >> 2), dseed) : 0u;
        const uint64_t nl2 = pack2(neg_lse, neg_lse), sc2 = pack2(sc, sc), ndc2 = pack2(ndc, ndc);
        const uint64_t cc2 = pack2(a.inv_sqrt_d, a.inv_sqrt_d);
#pragma unroll
        for (int qd4 = 0; qd4 < 16; ++qd4) {      // four keys (one dropout quad) per step, two packed fp32 pairs
          const uint32_t z = DROP ? drop4_z(dpre + static_cast<uint32_t>(qd4) * kDropC1, dtt) : 0u;
#pragma unroll
          for (int hf = 0; hf < 2; ++hf) {
            const int e = qd4 * 2 + hf;
            float x0, x1;
            unpack2(fma2(pack2u(sv[2 * e], sv[2 * e + 1]), sc2, nl2), x0, x1);
            const float p0 = fast_exp2(x0), p1 = fast_exp2(x1);
            const __half2 hp = __floats2half2_rn(p0, p1);
            pk[e] = *reinterpret_cast<const uint32_t*>(&hp);
            uint64_t c2 = cc2;
            if (DROP) {       // P o mask feeds dV (its 1/(1-p) is applied to dV at the end); dP flows back through mask/(1-p)
              pk[e] &= hf ? drop4_keep_h2_hi(z) : drop4_keep_h2_lo(z);
              c2 = pack2u((hf ? drop4_keep<2>(z) : drop4_keep<0>(z)) & cs_bits, (hf ? drop4_keep<3>(z) : drop4_keep<1>(z)) & cs_bits);
            }
            float d0, d1;                                                 // P o (dP_eff - delta) / sqrt(d), two per instruction
            unpack2(mul2(pack2(p0, p1), fma2(pack2u(dp[2 * e], dp[2 * e + 1]), c2, ndc2)), d0, d1);
            const __half2 hd = __floats2half2_rn(d0, d1);
            dk[e] = *reinterpret_cast<const uint32_t*>(&hd);
          }
        }
        ATT_TRACE(4);                             // P / dS computed, waiting for the previous block's gradient MMAs
        if (i > 0) {                              // gradient MMAs of the previous block are done: its dQ is complete, P / dS are free
          mbar_wait(grad_done, (ir - 1) & 1);
          tc_fence_after();
          ATT_TRACE(5);                           // P / dS tiles free
        } else if (t == 0) {
          tma_wait_group_read<0>();               // the previous item's dK / dV stores have read the staging patches (see the epilogue)
        }
#pragma unroll
        for (int ch = 0; ch < 8; ++ch) {
          const int off = ((ch ^ (r & 7)) << 4);
          sts128(p_row + off, pk[4 * ch], pk[4 * ch + 1], pk[4 * ch + 2], pk[4 * ch + 3]);
          sts128(ds_row + off, dk[4 * ch], dk[4 * ch + 1], dk[4 * ch + 2], dk[4 * ch + 3]);
        }
        fence_proxy_async_smem();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(ds_full);
        ATT_TRACE(6);                             // P / dS published
        // dQ of the previous block sits in the other TMEM buffer: reduce it into HBM off the critical path
        if (i > 0) drain_dq(ir - 1, i - 1);
        ATT_TRACE(7);                             // previous dQ drained
      }
      // ------------------------------------------------------------------ item epilogue
      mbar_wait(grad_done, (ir - 1) & 1);
      tc_fence_after();
      ATT_TRACE(8);                               // last gradient MMAs done
      drain_dq(ir - 1, nq - 1);
      // dV, dK: TMEM lane == key row; this thread owns 32 of the 64 d columns.  A full key block leaves as two TMA stores from
      // the dQ staging patches ([128 keys][64 d] fp16 each); a block that straddles the end of the batch element stores directly.
      const int key = k0 + r;
      const bool full_block = k0 + ATT_BK <= it.sk;      // (packed rows: a tile must not run into the next sequence)
      if (full_block) {
        if (lane == 0) tma_wait_group_read<0>();  // every warp's last dQ reduction has read its patch ...
        asm volatile("bar.sync 1, 256;" ::: "memory");   // ... before anybody overwrites it
      }
#pragma unroll 1
      for (int which = 0; which < 2; ++which) {
        uint32_t o[32];
        tmem_ld_x32(tmem + lane_addr + 256 + which * 64 + g * 32, o);
        tmem_wait_ld();
        uint32_t w[16];
        const float osc = (DROP && which == 0) ? a.drop.scale : 1.0f;      // dV = (P o mask / (1-p))^T dO
#pragma unroll
        for (int k = 0; k < 16; ++k) {
          const __half2 hv = __floats2half2_rn(__uint_as_float(o[2 * k]) * osc, __uint_as_float(o[2 * k + 1]) * osc);
          w[k] = *reinterpret_cast<const uint32_t*>(&hv);
        }
        if (full_block) {
          const uint32_t row_s = smem_u32(smem + S::OFF_DQS + which * S::T) + r * 128;
#pragma unroll
          for (int e = 0; e < 4; ++e) sts128(row_s + (((g * 4 + e) ^ (r & 7)) << 4), w[4 * e], w[4 * e + 1], w[4 * e + 2], w[4 * e + 3]);
        } else if (key < it.sk) {
          __half* dst = (which == 0 ? a.dv + a.dv_col0 : a.dk + a.dk_col0) + (static_cast<size_t>(it.krow0) + key) * a.ld_dkv + h * ATT_D + g * 32;
#pragma unroll
          for (int e = 0; e < 4; ++e) *reinterpret_cast<uint4*>(dst + e * 8) = make_uint4(w[4 * e], w[4 * e + 1], w[4 * e + 2], w[4 * e + 3]);
        }
      }
      tc_fence_before();
      if (full_block) {
        fence_proxy_async_smem();
        asm volatile("bar.sync 1, 256;" ::: "memory");
        if (t == 0) {
          tma_store_2d(&tmDKV, smem + S::OFF_DQS, a.dv_col0 + h * ATT_D, it.krow0 + k0);
          tma_store_2d(&tmDKV, smem + S::OFF_DQS + S::T, a.dk_col0 + h * ATT_D, it.krow0 + k0);
          tma_commit_group();
        }
      }
      ATT_TRACE(9);                               // dK / dV on their way
    }
    if (lane == 0) tma_wait_group_read<0>();      // staging memory stays valid until the bulk copies have read it
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
