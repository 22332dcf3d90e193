// 2-CTA (cta_group::2) tcgen05 GEMM with a TMA-driven epilogue — the production path for the encoder's dense
// contractions.  Same math / operand layouts as gemm.cuh, re-tiled for the B200 SM pair:
//
//   * a cluster of two CTAs (one TPC) owns a 256 x 256 output tile; each CTA stages its own 128 rows of A and HALF of
//     the B tile (128 of 256 rows), so per-SM shared-memory fill and L2->SM traffic drop from 48 KB to 32 KB per
//     64-wide K block while every tcgen05.mma does 256x256x16;
//   * the leader CTA's MMA warp issues tcgen05.mma.cta_group::2 for both SMs; completion is multicast to the
//     mbarriers of both CTAs (smem-slot release, accumulator-ready);
//   * accumulators: 2 x 256 TMEM columns per CTA (epilogue of tile i overlaps the mainloop of tile i+1);
//   * epilogue: eight warps per CTA (two per TMEM lane quadrant), one thread per accumulator row: TMEM -> registers -> bias / GELU / residual /
//     dGELU math -> fp16/fp32 128B-swizzled staging rows -> TMA store (or TMA reduce-add for split-K wgrad).
//     Auxiliary row-major inputs (residual, saved activation derivative) are read 128 contiguous bytes per thread
//     straight into registers one chunk ahead of their use.
#pragma once
#include <type_traits>

#include "gemm.cuh"

namespace b200 {

constexpr int G2_THREADS = 320;           // warp 0 TMA, warp 1 MMA, warps 2..9 epilogue
constexpr int G2_BM = 256;                // rows per CTA pair
constexpr int G2_SMEM_MAX = 232448;       // 227 KB opt-in limit per CTA

// Shared-memory plan: 8 epilogue warps x 2 output slabs (64 KB); the operand ring takes the rest (5 stages of 32 KB).
template <int BN, int EPI>
struct Gemm2Smem {
  static constexpr int A_BYTES = 128 * GEMM_BK * 2;          // 16 KB : this CTA's 128 rows of A
  static constexpr int B_BYTES = (BN / 2) * GEMM_BK * 2;     // 16 KB : this CTA's half of the B tile
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int SLAB = 32 * 128;                      // 4 KB : 32 rows x 128 B staging slab
#ifndef G2_SLABS
#define G2_SLABS 2
#endif
#ifndef G2_STAGE_CAP
#define G2_STAGE_CAP 16
#endif
  static constexpr int SLABS = EPI == EPI_BIAS_GELU ? 2 : G2_SLABS;     // output slabs per epilogue warp (GELU: one per output)
  static constexpr int EPI_PER_WARP = SLABS * SLAB;
  static constexpr int EPI_WARPS = 8;
  static constexpr int STAGES_FIT = (G2_SMEM_MAX - 1024 - 512 - EPI_WARPS * EPI_PER_WARP) / STAGE_BYTES;
  static constexpr int STAGES = STAGES_FIT < G2_STAGE_CAP ? STAGES_FIT : G2_STAGE_CAP;
  static constexpr int OFF_EPI = STAGES * STAGE_BYTES;
  static constexpr int OFF_BAR = OFF_EPI + EPI_WARPS * EPI_PER_WARP;
  static constexpr int TOTAL = OFF_BAR + 512 + 1024;
  static_assert(STAGES >= 2 && TOTAL <= G2_SMEM_MAX, "shared-memory plan");
};

// ------------------------------------------------------------------------------------------------ cluster helpers
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t mapa_u32(uint32_t local_smem_addr, uint32_t cta_rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_smem_addr), "r"(cta_rank));
  return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// TMA load whose completion bytes are signalled on an mbarrier that may live in the peer CTA of the pair
__device__ __forceinline__ void tma_load_2d_cg2(void* smem_dst, const CUtensorMap* m, uint32_t mbar_cluster_addr, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      :: "r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(mbar_cluster_addr), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tmem_alloc_cg2(uint32_t* smem_result, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_result)), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tmem_relinquish_cg2() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_cg2(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma2_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      :: "r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// completion of all prior MMAs -> arrive on the barrier at this smem offset in BOTH CTAs of the pair
__device__ __forceinline__ void umma2_commit_mc(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(smem_u32(bar)), "h"(static_cast<uint16_t>(3)) : "memory");
}
// (TMA store / reduce helpers live in ptx.cuh)

struct Gemm2Maps {
  CUtensorMap a, b, out, aux, out2;
};

// Staging slab: 32 rows x 128 B, 128B-swizzled (16-byte chunk c of row r lives at chunk slot c ^ (r & 7)).
__device__ __forceinline__ uint32_t slab_chunk(uint32_t slab_saddr, int row, int chunk) {
  return slab_saddr + row * 128 + ((chunk ^ (row & 7)) << 4);
}

// OutT = __half / float, or __nv_bfloat16: bf16 OPERANDS (and a bf16 or — with OutT float via BF16IN — fp32 result) for the plain
// epilogues (SURVEY.md config 4's bf16 arm: same tensor-core rate, 8 mantissa bits instead of 11).
template <int BN, int A_MN, int B_MN, int EPI, typename OutT, bool BF16IN = false>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(G2_THREADS, 1)
gemm2_f16_kernel(const __grid_constant__ Gemm2Maps maps, const GemmArgs g) {
  using S = Gemm2Smem<BN, EPI>;
  constexpr int G2_STAGES = S::STAGES;
  constexpr bool OUT32 = sizeof(OutT) == 4;
  constexpr bool BF16 = BF16IN || std::is_same<OutT, __nv_bfloat16>::value;
  static_assert(!BF16 || EPI == EPI_STORE || EPI == EPI_BIAS || EPI == EPI_ATOMIC, "bf16 operands: plain epilogues only");
  constexpr int CW = OUT32 ? 32 : 64;                 // accumulator columns per staging slab (128 B of output per row)
  constexpr bool RESADD = EPI == EPI_RESADD;         // out32 += drop(acc + bias): residual already in `out`, TMA reduce-add
  constexpr bool DELTA = EPI == EPI_STORE_DELTA;      // out16 = acc, delta[b,h,q] = <out row, aux row> per 64-column head
  constexpr bool HAS_AUX = EPI == EPI_BIAS_RES || EPI == EPI_BIAS_RES32 || EPI == EPI_DGELU || EPI == EPI_ADD || DELTA;
  constexpr bool HAS_BIAS = EPI == EPI_BIAS || EPI == EPI_BIAS_GELU || EPI == EPI_BIAS_RES || EPI == EPI_BIAS_RES32 || RESADD;
  static_assert(!(EPI == EPI_BIAS_RES32) || OUT32, "fp32 residual stream implies fp32 output");
  static_assert(!(EPI == EPI_BIAS_RES || EPI == EPI_DGELU || EPI == EPI_ADD || EPI == EPI_BIAS_GELU) || !OUT32, "fp16-aux epilogues write fp16");
  static_assert(!(EPI == EPI_ATOMIC) || OUT32, "split-K reduction is fp32");
  static_assert(!RESADD || OUT32, "the residual stream is fp32");
  static_assert(!DELTA || !OUT32, "the fused row statistic rides on the fp16 store (one 64-column chunk = one head)");

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;
  uint8_t* sB = smem + G2_STAGES * S::A_BYTES;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + S::OFF_BAR);
  uint64_t* empty_bar = full_bar + G2_STAGES;
  uint64_t* tfull_bar = empty_bar + G2_STAGES;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int pair = blockIdx.x >> 1, n_pairs = gridDim.x >> 1;

  const int m_tiles = (g.M + G2_BM - 1) / G2_BM;
  const int n_tiles = (g.N + BN - 1) / BN;
  const int k_blocks = (g.K + GEMM_BK - 1) / GEMM_BK;
  const int splits = g.k_splits > 0 ? g.k_splits : 1;
  const int kb_per_split = (k_blocks + splits - 1) / splits;
  const int units = m_tiles * n_tiles * splits;
  const int u_begin = pair, u_end = units;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&maps.a);
    tma_prefetch_desc(&maps.b);
    tma_prefetch_desc(&maps.out);
    for (int i = 0; i < G2_STAGES; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull_bar[i], 1);
      mbar_init(&tempty_bar[i], 16);              // 8 epilogue warps x 2 CTAs (only the leader's copy is used)
    }
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc_cg2(tmem_slot, 2 * BN);
    tmem_relinquish_cg2();
  }
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  auto decode = [&](int u, int& mt, int& nt, int& kb0, int& kb1) {
    const int sp = u % splits;
    const int t = u / splits;
    nt = t % n_tiles;
    mt = t / n_tiles;
    kb0 = sp * kb_per_split;
    kb1 = min(k_blocks, kb0 + kb_per_split);
  };
  auto next_unit = [&](int u, int, int) { return u + n_pairs; };

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer (both CTAs)
    // The whole warp walks the loop; one elected lane issues (see ptx.cuh: elect_one_sync).
    uint32_t stage = 0, phase = 0;
    int mt, nt, kb0, kb1;
    for (int u = u_begin; u < u_end; u = next_unit(u, kb0, kb1)) {
      decode(u, mt, nt, kb0, kb1);
      const int m0 = mt * G2_BM + static_cast<int>(rank) * 128;          // this CTA's A rows
      const int n0 = nt * BN + static_cast<int>(rank) * (BN / 2);        // this CTA's half of the B tile
      for (int kb = kb0; kb < kb1; ++kb) {
        mbar_wait(&empty_bar[stage], phase ^ 1);
        if (elect_one_sync()) {
#ifdef G2_DIAG_NO_TMA                                     // diagnostic build (tools/gemm_stages.py): the mainloop without its loads
          if (leader) mbar_arrive(&full_bar[stage]);
#else
          const uint32_t full0 = mapa_u32(smem_u32(&full_bar[stage]), 0);   // the leader's barrier collects both CTAs' bytes
#ifdef G2_DIAG_HALF_B                                     // diagnostic build: every other k-block skips its B loads (-25 % L2 reads)
          const bool skip_b = (kb & 1) != 0;
          if (leader) mbar_expect_tx(&full_bar[stage], 2 * (S::A_BYTES + (skip_b ? 0 : S::B_BYTES)));
#else
          const bool skip_b = false;
          if (leader) mbar_expect_tx(&full_bar[stage], 2 * (S::A_BYTES + S::B_BYTES));
#endif
          uint8_t* a_dst = sA + stage * S::A_BYTES;
          uint8_t* b_dst = sB + stage * S::B_BYTES;
          const int k0 = kb * GEMM_BK;
          if (A_MN) {
#pragma unroll
            for (int j = 0; j < 2; ++j) tma_load_2d_cg2(a_dst + j * 8192, &maps.a, full0, m0 + 64 * j, k0);
          } else {
            tma_load_2d_cg2(a_dst, &maps.a, full0, k0, m0);
          }
          if (skip_b) {
          } else if (B_MN) {
#pragma unroll
            for (int j = 0; j < BN / 128; ++j) tma_load_2d_cg2(b_dst + j * 8192, &maps.b, full0, n0 + 64 * j, k0);
          } else {
            tma_load_2d_cg2(b_dst, &maps.b, full0, k0, n0);
          }
#endif
        }
        __syncwarp();
        if (++stage == G2_STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer (leader CTA only)
    if (leader) {
      constexpr uint32_t idesc = make_idesc_f16(G2_BM, BN, A_MN, B_MN, BF16 ? 1 : 0);
      uint32_t stage = 0, phase = 0, acc = 0, acc_phase = 0;
      int mt, nt, kb0, kb1;
      for (int u = u_begin; u < u_end; u = next_unit(u, kb0, kb1)) {
        decode(u, mt, nt, kb0, kb1);
        mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN;
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          if (elect_one_sync()) {                  // (not `lane == 0`: see ptx.cuh — 4-5 instead of 10+ instructions per MMA)
            const uint32_t a_addr = smem_u32(sA + stage * S::A_BYTES);
            const uint32_t b_addr = smem_u32(sB + stage * S::B_BYTES);
#pragma unroll
            for (int kk = 0; kk < GEMM_BK / 16; ++kk) {
              const uint64_t da = A_MN ? make_smem_desc(a_addr + kk * 2048, 8192, 1024) : make_smem_desc(a_addr + kk * 32, 0, 1024);
              const uint64_t db = B_MN ? make_smem_desc(b_addr + kk * 2048, 8192, 1024) : make_smem_desc(b_addr + kk * 32, 0, 1024);
#ifndef G2_DIAG_NO_MMA                                    // diagnostic build: the load pipeline without the MMAs
              umma2_ss(d_tmem, da, db, idesc, (kb > kb0 || kk > 0) ? 1u : 0u);
#else
              (void)da; (void)db;
#endif
            }
            umma2_commit_mc(&empty_bar[stage]);
            if (kb == kb1 - 1) umma2_commit_mc(&tfull_bar[acc]);
          }
          __syncwarp();
          if (++stage == G2_STAGES) { stage = 0; phase ^= 1; }
        }
        if (kb1 <= kb0 && elect_one_sync()) umma2_commit_mc(&tfull_bar[acc]);
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1;
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue (warps 2..9, both CTAs)
    // Two warps per TMEM lane quadrant: warp (q, half) owns rows 32q..32q+31 and one half of the tile's columns; one
    // thread per accumulator row.  Outputs leave through double-buffered swizzled slabs + TMA store (asynchronous,
    // fully coalesced).  Auxiliary inputs (residual / saved activation derivative) are fetched straight into registers,
    // 128 contiguous bytes per thread, one chunk AHEAD of their use, so their latency hides behind the math of the
    // current chunk without costing shared memory.
    const int q = warp & 3;                               // TMEM lane quadrant
    const int ew = warp - 2;                              // 0..7
    const int half = ew >> 2;                             // which half of the BN columns
    uint8_t* out_s = smem + S::OFF_EPI + ew * S::EPI_PER_WARP;      // 2 slabs
    const uint32_t tempty0[2] = {mapa_u32(smem_u32(&tempty_bar[0]), 0), mapa_u32(smem_u32(&tempty_bar[1]), 0)};
    const float alpha = g.alpha ? __ldg(g.alpha) : 1.0f;
    const uint32_t lane_addr = static_cast<uint32_t>(q * 32) << 16;
    uint32_t acc = 0, acc_phase = 0, out_uses = 0;
    constexpr int NCH = BN / CW / 2;                      // chunks per warp per tile
    constexpr int AUX_ESZ = EPI == EPI_BIAS_RES32 ? 4 : 2;
    int seg_kb0 = 0, seg_kb1 = 0;                         // K-block range of the unit last decoded by tile_origin
    auto tile_origin = [&](int u, int& row0, int& col0, bool& has_data) {
      int mt, nt;
      decode(u, mt, nt, seg_kb0, seg_kb1);
      row0 = mt * G2_BM + static_cast<int>(rank) * 128 + q * 32;      // first output row of this warp
      col0 = nt * BN + half * (BN / 2);                               // first output column of this warp
      has_data = seg_kb1 > seg_kb0;
    };
    // 128 bytes of this thread's aux row for the chunk starting at column gc0 (zeros outside the matrix)
    auto load_aux = [&](uint4 (&dst)[8], int row, int gc0) {
      const char* base = reinterpret_cast<const char*>(g.aux) + (static_cast<size_t>(row) * g.ld_aux + gc0) * AUX_ESZ;
      const bool row_ok = row < g.M;
#pragma unroll
      for (int ch = 0; ch < 8; ++ch) {
        const bool ok = row_ok && (gc0 + ch * (16 / AUX_ESZ)) < g.N;
        dst[ch] = ok ? __ldg(reinterpret_cast<const uint4*>(base) + ch) : make_uint4(0, 0, 0, 0);
      }
    };
    uint4 aux_next[8];
    if (HAS_AUX && u_begin < u_end) {
      int r0, c0;
      bool hd;
      tile_origin(u_begin, r0, c0, hd);
      load_aux(aux_next, r0 + lane, c0);
    }
    int kb0 = 0, kb1 = 0;
    for (int u = u_begin; u < u_end; u = next_unit(u, kb0, kb1)) {
      int row0, col0;
      bool has_data;
      tile_origin(u, row0, col0, has_data);
      kb0 = seg_kb0;
      kb1 = seg_kb1;
      mbar_wait(&tfull_bar[acc], acc_phase);
      tc_fence_after();
#pragma unroll 1
      for (int c = 0; c < NCH; ++c) {
        const int gc0 = col0 + c * CW;
        uint32_t v[CW];
#pragma unroll
        for (int i = 0; i < CW / 32; ++i)
          tmem_ld_x32(tmem_base + lane_addr + acc * BN + half * (BN / 2) + c * CW + i * 32, *reinterpret_cast<uint32_t(*)[32]>(&v[i * 32]));
        uint4 araw[8];
        if (HAS_AUX) {
#pragma unroll
          for (int ch = 0; ch < 8; ++ch) araw[ch] = aux_next[ch];
          if (c + 1 < NCH) {
            load_aux(aux_next, row0 + lane, gc0 + CW);
          } else if (u + n_pairs < units) {               // first chunk of this warp's next tile
            int r1, c1;
            bool hd;
            tile_origin(u + n_pairs, r1, c1, hd);
            load_aux(aux_next, r1 + lane, c1);
          }
        }
        tmem_wait_ld();
        if (c == NCH - 1) {                               // accumulator drained -> hand it back to the leader's MMA warp
          tc_fence_before();
          if (lane == 0) mbar_arrive_cluster(tempty0[acc]);
        }
        float f[CW];
#pragma unroll
        for (int i = 0; i < CW; ++i) f[i] = __uint_as_float(v[i]);
        if (!RESADD && g.alpha != nullptr) {            // in practice only the wgrad reduction scales (one uniform branch elsewhere)
          const uint64_t al2 = pack2(alpha, alpha);
#pragma unroll
          for (int i = 0; i < CW; i += 2) unpack2(mul2(pack2(f[i], f[i + 1]), al2), f[i], f[i + 1]);
        }
        if (HAS_BIAS) {
#pragma unroll
          for (int i = 0; i < CW / 4; ++i) {
            float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f);
            if (gc0 + 4 * i < g.N) b4 = __ldg(reinterpret_cast<const float4*>(g.bias + gc0 + 4 * i));
            unpack2(add2(pack2(f[4 * i], f[4 * i + 1]), pack2(b4.x, b4.y)), f[4 * i], f[4 * i + 1]);
            unpack2(add2(pack2(f[4 * i + 2], f[4 * i + 3]), pack2(b4.z, b4.w)), f[4 * i + 2], f[4 * i + 3]);
          }
        }
        if ((EPI == EPI_BIAS_RES32 || EPI == EPI_BIAS_RES || RESADD) && g.drop.seed_base != nullptr) {
          // hidden dropout of BertSelfOutput / BertOutput (bert_model.py:373,451): applied to dense(x)+bias, before the residual
          const uint32_t seed = drop_seed(g.drop);
          const uint32_t e0 = static_cast<uint32_t>(row0 + lane) * static_cast<uint32_t>(g.N) + static_cast<uint32_t>(gc0);
#pragma unroll
          const uint32_t pre0 = drop_premix(e0 >> 1, seed), tt = g.drop.thr15 * 0x00010001u;
          const uint64_t sc2 = pack2(g.drop.scale, g.drop.scale);
          for (int i = 0; i < CW / 2; ++i) {     // both decisions of a pair from one SWAR compare, the 1/(1-p) as one packed multiply
            const uint32_t z = drop_z(pre0 + static_cast<uint32_t>(i) * kDropC1, tt);
            const uint64_t kept = pack2u(__float_as_uint(f[2 * i]) & drop_keep_lo(z), __float_as_uint(f[2 * i + 1]) & drop_keep_hi(z));
            unpack2(mul2(kept, sc2), f[2 * i], f[2 * i + 1]);
          }
        }
        if (HAS_AUX && !DELTA) {
#pragma unroll
          for (int ch = 0; ch < 8; ++ch) {
            const uint4 raw = araw[ch];
            if (EPI == EPI_BIAS_RES32) {                  // 4 fp32 per 16-byte chunk
              f[4 * ch] += __uint_as_float(raw.x); f[4 * ch + 1] += __uint_as_float(raw.y);
              f[4 * ch + 2] += __uint_as_float(raw.z); f[4 * ch + 3] += __uint_as_float(raw.w);
            } else {                                      // 8 fp16 per chunk
              const __half2* hp = reinterpret_cast<const __half2*>(&raw);
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                const float2 x = __half22float2(hp[k]);
                if (EPI == EPI_DGELU) {
                  f[8 * ch + 2 * k] *= x.x;
                  f[8 * ch + 2 * k + 1] *= x.y;
                } else {
                  f[8 * ch + 2 * k] += x.x;
                  f[8 * ch + 2 * k + 1] += x.y;
                }
              }
            }
          }
        }
        uint32_t ow[32];                                  // packed output words of this thread's 128-byte row segment
        if (EPI == EPI_BIAS_GELU) {
          uint32_t zw[32];
#pragma unroll
          for (int k = 0; k < 32; ++k) {
            float y0, y1, d0, d1;
            uint64_t y2, d2;
#ifdef G2_DIAG_NO_ERF                                     // diagnostic build: both outputs stored, no erf math
            y2 = pack2(f[2 * k], f[2 * k + 1]);
            d2 = y2;
#else
            gelu_erf_both2(pack2(f[2 * k], f[2 * k + 1]), y2, d2);
#endif
            unpack2(y2, y0, y1);
            unpack2(d2, d0, d1);
            const __half2 hz = __floats2half2_rn(d0, d1), hh = __floats2half2_rn(y0, y1);
            zw[k] = *reinterpret_cast<const uint32_t*>(&hz);
            ow[k] = *reinterpret_cast<const uint32_t*>(&hh);
          }
          // GELU variant: slab 0 carries h, slab 1 the saved derivative, both single-buffered — the erf math of the next
          // chunk (microseconds) separates a TMA store from the next write to its slab, so the wait below is free.
          if (lane == 0) tma_wait_group_read<0>();
          __syncwarp();
          const uint32_t hs = smem_u32(out_s), ds = smem_u32(out_s + S::SLAB);
#pragma unroll
          for (int ch = 0; ch < 8; ++ch) {
            sts128(slab_chunk(hs, lane, ch), ow[4 * ch], ow[4 * ch + 1], ow[4 * ch + 2], ow[4 * ch + 3]);
            sts128(slab_chunk(ds, lane, ch), zw[4 * ch], zw[4 * ch + 1], zw[4 * ch + 2], zw[4 * ch + 3]);
          }
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) {
            if (has_data) {
              tma_store_2d(&maps.out, out_s, gc0, row0);
#ifndef G2_DIAG_NO_OUT2                                   // diagnostic build: erf math, derivative not stored
              if (g.out2) tma_store_2d(&maps.out2, out_s + S::SLAB, gc0, row0);
#endif
            }
            tma_commit_group();
          }
          continue;
        } else if (OUT32) {
#pragma unroll
          for (int k = 0; k < 32; ++k) ow[k] = __float_as_uint(f[k]);
        } else {
          float dsum = 0.f;
#pragma unroll
          for (int k = 0; k < 32; ++k) {
            const __half2 hv = __floats2half2_rn(f[2 * k], f[2 * k + 1]);
            ow[k] = *reinterpret_cast<const uint32_t*>(&hv);
            if (std::is_same<OutT, __nv_bfloat16>::value) {
              const __nv_bfloat162 bv = __floats2bfloat162_rn(f[2 * k], f[2 * k + 1]);
              ow[k] = *reinterpret_cast<const uint32_t*>(&bv);
            }
            if (DELTA) {                                  // the statistic is taken on the ROUNDED output, the values attention backward reads
              const float2 o2 = __half22float2(hv);
              const float2 c2 = __half22float2(reinterpret_cast<const __half2*>(araw)[k]);
              dsum = fmaf(o2.x, c2.x, dsum);
              dsum = fmaf(o2.y, c2.y, dsum);
            }
          }
          if (DELTA) {                                    // one 64-column chunk is one head: row (b, q), head gc0 / 64
            const int row = row0 + lane;
            if (has_data && row < g.M && gc0 < g.N) {
              const int Sq = g.ld_out2, b = row / Sq;     // delta [B, N/64, Sq] rides in the out2 slot (see GemmArgs)
              reinterpret_cast<float*>(g.out2)[(static_cast<size_t>(b) * (g.N >> 6) + (gc0 >> 6)) * Sq + (row - b * Sq)] = dsum;
            }
          }
        }
        // double-buffered slab: the TMA store issued two chunks ago must have finished reading this buffer
        uint8_t* os_ptr = out_s + (out_uses % S::SLABS) * S::SLAB;
        if (lane == 0) tma_wait_group_read<S::SLABS - 1>();
        __syncwarp();
        const uint32_t os = smem_u32(os_ptr);
#pragma unroll
        for (int ch = 0; ch < 8; ++ch) sts128(slab_chunk(os, lane, ch), ow[4 * ch], ow[4 * ch + 1], ow[4 * ch + 2], ow[4 * ch + 3]);
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) {
          if (has_data) {
            if (EPI == EPI_ATOMIC || RESADD) tma_reduce_add_2d(&maps.out, os_ptr, gc0, row0);
            else tma_store_2d(&maps.out, os_ptr, gc0, row0);
          }
          tma_commit_group();
        }
        if (EPI == EPI_DGELU && g.colsum != nullptr && has_data) {
          // Fused bias gradient (replaces a separate column-sum pass that re-read the whole [M, N] tensor from HBM): the warp's
          // 32 rows x 64 columns sit in the slab as fp16; lane l owns columns gc0 + 2l, 2l + 1 = 32-bit word l of every row
          // (word w of row r lives in chunk (w >> 2) ^ (r & 7): all lanes read one row per step, distinct banks).  Rows past M
          // hold exact zeros (their A rows and aux were zero-filled).
          float s0 = 0.f, s1 = 0.f;
#pragma unroll
          for (int rr = 0; rr < 32; ++rr) {
            uint32_t wv;
            asm volatile("ld.shared.b32 %0, [%1];" : "=r"(wv) : "r"(os + rr * 128 + ((((lane >> 2) ^ (rr & 7)) << 4) | ((lane & 3) << 2))) : "memory");
            const float2 x = __half22float2(*reinterpret_cast<const __half2*>(&wv));
            s0 += x.x;
            s1 += x.y;
          }
          const float ca = g.col_alpha ? __ldg(g.col_alpha) : 1.0f;
          const int cc = gc0 + 2 * lane;
          if (cc < g.N) atomicAdd(g.colsum + cc, s0 * ca);
          if (cc + 1 < g.N) atomicAdd(g.colsum + cc + 1, s1 * ca);
        }
        ++out_uses;
      }
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
    if (lane == 0) tma_wait_group_all();
  }

  tc_fence_before();
  cluster_sync_all();                                     // the peer may still signal our barriers / read our smem
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_cg2(tmem_base, 2 * BN);
  }
}

}  // namespace b200
