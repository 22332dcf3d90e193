// tcgen05 GEMM family for the encoder's dense contractions.
//
//   C[M,N] = epilogue( A[M,K] * B[N,K]^T ),  fp16 operands, fp32 accumulation in TMEM.
//
// One persistent CTA per SM, 192 threads, warp-specialised:
//   warp 0      TMA producer   : global -> 128B-swizzled smem ring (4 stages of 64-wide K blocks)
//   warp 1      MMA issuer     : one lane issues tcgen05.mma (M=128, N=BN, K=16), accumulator in TMEM,
//                                two accumulator stages so the epilogue of tile i overlaps the MMAs of tile i+1
//   warps 2..5  epilogue       : tcgen05.ld (row-per-thread) -> warp-private smem transpose -> fully coalesced
//                                bias / GELU / residual / dGELU / fp32-atomic (split-K) stores
//
// Operands may be K-major (contraction dim contiguous: activations x, weights W[out,in] in the forward) or
// MN-major (contraction dim strided: W in dgrad, dY^T and X^T in wgrad).  MN-major tiles are loaded as 64-wide
// column blocks ([k rows][128 B], 128B swizzle) and described to the tensor core with a_major/b_major = 1, so the
// backward needs no transposed copies of weights or activations.
#pragma once
#include "ptx.cuh"

namespace b200 {

constexpr int GEMM_BM = 128;
constexpr int GEMM_BK = 64;
constexpr int GEMM_STAGES = 4;
constexpr int GEMM_THREADS = 192;
constexpr int GEMM_STAGE_PITCH = 36;   // floats per staged row (32 + 4 pad -> conflict-free v4 access)

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
  int dbg;              // measurement knobs (gemm2 only): 1 skip A loads, 2 skip B loads, 4 skip MMA issue, 8 skip epilogue stores,
                        // bits 8..15: L2 prefetch distance in K blocks
  // The opt-in epilogues reuse fields they do not otherwise need, so that this struct — and with it the machine code of
  // every round-1 kernel — stays exactly as validated (tools/sass_diff.py):
  //   EPI_RESADD      : k_splits == -1 selects the stream-K schedule (equal share of K blocks per CTA pair)
  //   EPI_STORE_DELTA : out2 (as float*) = delta [B, N/64, ld_out2], ld_out2 = tokens per sequence (Sq)
};

template <int BN>
struct GemmSmem {
  static constexpr int A_BYTES = GEMM_BM * GEMM_BK * 2;
  static constexpr int B_BYTES = BN * GEMM_BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int EPI_BYTES = 4 * 32 * GEMM_STAGE_PITCH * 4;
  static constexpr int BAR_OFFSET = GEMM_STAGES * STAGE_BYTES + EPI_BYTES;
  static constexpr int TOTAL = BAR_OFFSET + 256 + 1024;   // + barriers + alignment slack
};

// Epilogue of one 32-row x BN-column slab owned by one warp: TMEM -> registers (row per thread) -> warp-private smem
// transpose -> coalesced bias / activation / residual math and global stores.  `release_acc` is invoked (lane 0) as soon
// as the accumulator has been read out of TMEM.
template <int BN, int EPI, typename OutT, typename ReleaseFn>
__device__ __forceinline__ void gemm_epilogue_tile(const GemmArgs& g, float* st, uint32_t tmem_acc, int m0, int n0, int lane, float alpha,
                                                   bool has_data, ReleaseFn release_acc) {
  const int cl = (lane & 7) * 4;                // column group handled in the coalesced phase
  const int rl = lane >> 3;                     // row within a group of 4
#pragma unroll 1
  for (int c = 0; c < BN / 32; ++c) {
    uint32_t v[32];
    tmem_ld_x32(tmem_acc + c * 32, v);
    tmem_wait_ld();
    if (c == BN / 32 - 1) {                   // accumulator drained: hand the TMEM stage back to the MMA warp
      tc_fence_before();
      if (lane == 0) release_acc();
    }
    float* my = st + lane * GEMM_STAGE_PITCH;
#pragma unroll
    for (int j = 0; j < 8; ++j)
      *reinterpret_cast<float4*>(my + 4 * j) = make_float4(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1]),
                                                           __uint_as_float(v[4 * j + 2]), __uint_as_float(v[4 * j + 3]));
    __syncwarp();
    const int gc = n0 + c * 32 + cl;
    if (gc < g.N && has_data) {
      float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f);
      if (EPI == EPI_BIAS || EPI == EPI_BIAS_GELU || EPI == EPI_BIAS_RES || EPI == EPI_BIAS_RES32) b4 = __ldg(reinterpret_cast<const float4*>(g.bias + gc));
#pragma unroll
      for (int it = 0; it < 8; ++it) {
        const int r = it * 4 + rl;
        const int gr = m0 + r;
        if (gr >= g.M) continue;
        float4 a = *reinterpret_cast<const float4*>(st + r * GEMM_STAGE_PITCH + cl);
        a.x = fmaf(a.x, alpha, b4.x); a.y = fmaf(a.y, alpha, b4.y); a.z = fmaf(a.z, alpha, b4.z); a.w = fmaf(a.w, alpha, b4.w);
        if (EPI == EPI_BIAS_RES32) {
          const float4 r4 = __ldg(reinterpret_cast<const float4*>(reinterpret_cast<const float*>(g.aux) + static_cast<size_t>(gr) * g.ld_aux + gc));
          a.x += r4.x; a.y += r4.y; a.z += r4.z; a.w += r4.w;
        }
        if (EPI == EPI_BIAS_RES || EPI == EPI_DGELU || EPI == EPI_ADD) {
          const uint2 raw = __ldg(reinterpret_cast<const uint2*>(g.aux + static_cast<size_t>(gr) * g.ld_aux + gc));
          const float2 x01 = __half22float2(*reinterpret_cast<const __half2*>(&raw.x));
          const float2 x23 = __half22float2(*reinterpret_cast<const __half2*>(&raw.y));
          if (EPI == EPI_DGELU) {
            a.x *= x01.x; a.y *= x01.y; a.z *= x23.x; a.w *= x23.y;
          } else {
            a.x += x01.x; a.y += x01.y; a.z += x23.x; a.w += x23.y;
          }
        }
        if (EPI == EPI_BIAS_GELU) {
          float4 d, y;
          gelu_erf_both(a.x, y.x, d.x); gelu_erf_both(a.y, y.y, d.y); gelu_erf_both(a.z, y.z, d.z); gelu_erf_both(a.w, y.w, d.w);
          a = y;
          if (g.out2) {
            __half2 z01 = __floats2half2_rn(d.x, d.y), z23 = __floats2half2_rn(d.z, d.w);
            uint2 zr;
            zr.x = *reinterpret_cast<uint32_t*>(&z01);
            zr.y = *reinterpret_cast<uint32_t*>(&z23);
            *reinterpret_cast<uint2*>(g.out2 + static_cast<size_t>(gr) * g.ld_out2 + gc) = zr;
          }
        }
        if (EPI == EPI_ATOMIC) {
          red_add_v4(reinterpret_cast<float*>(g.out) + static_cast<size_t>(gr) * g.ld_out + gc, a.x, a.y, a.z, a.w);
        } else if (sizeof(OutT) == 4) {
          *reinterpret_cast<float4*>(reinterpret_cast<float*>(g.out) + static_cast<size_t>(gr) * g.ld_out + gc) = a;
        } else {
          __half2 o01 = __floats2half2_rn(a.x, a.y), o23 = __floats2half2_rn(a.z, a.w);
          uint2 o;
          o.x = *reinterpret_cast<uint32_t*>(&o01);
          o.y = *reinterpret_cast<uint32_t*>(&o23);
          *reinterpret_cast<uint2*>(reinterpret_cast<__half*>(g.out) + static_cast<size_t>(gr) * g.ld_out + gc) = o;
        }
      }
    }
    __syncwarp();
  }
}

template <int BN, int A_MN, int B_MN, int EPI, typename OutT>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_f16_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const GemmArgs g) {
  using S = GemmSmem<BN>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;
  uint8_t* sB = smem + GEMM_STAGES * S::A_BYTES;
  float* sEpi = reinterpret_cast<float*>(smem + GEMM_STAGES * S::STAGE_BYTES);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + S::BAR_OFFSET);
  uint64_t* empty_bar = full_bar + GEMM_STAGES;
  uint64_t* tfull_bar = empty_bar + GEMM_STAGES;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  const int m_tiles = (g.M + GEMM_BM - 1) / GEMM_BM;
  const int n_tiles = (g.N + BN - 1) / BN;
  const int k_blocks = (g.K + GEMM_BK - 1) / GEMM_BK;
  const int splits = g.k_splits > 0 ? g.k_splits : 1;
  const int kb_per_split = (k_blocks + splits - 1) / splits;
  const int units = m_tiles * n_tiles * splits;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int i = 0; i < GEMM_STAGES; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull_bar[i], 1);
      mbar_init(&tempty_bar[i], 4);
    }
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, 2 * BN);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // unit -> (m_tile, n_tile, split); n fastest so CTAs running together share the A row panel in L2.
  auto decode = [&](int u, int& mt, int& nt, int& kb0, int& kb1) {
    const int sp = u % splits;
    const int t = u / splits;
    nt = t % n_tiles;
    mt = t / n_tiles;
    kb0 = sp * kb_per_split;
    kb1 = min(k_blocks, kb0 + kb_per_split);
  };

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      uint32_t stage = 0, phase = 0;
      for (int u = blockIdx.x; u < units; u += gridDim.x) {
        int mt, nt, kb0, kb1;
        decode(u, mt, nt, kb0, kb1);
        const int m0 = mt * GEMM_BM, n0 = nt * BN;
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          mbar_expect_tx(&full_bar[stage], S::STAGE_BYTES);
          uint8_t* a_dst = sA + stage * S::A_BYTES;
          uint8_t* b_dst = sB + stage * S::B_BYTES;
          const int k0 = kb * GEMM_BK;
          if (A_MN) {
#pragma unroll
            for (int j = 0; j < GEMM_BM / 64; ++j) tma_load_2d(a_dst + j * 8192, &tmA, &full_bar[stage], m0 + 64 * j, k0);
          } else {
            tma_load_2d(a_dst, &tmA, &full_bar[stage], k0, m0);
          }
          if (B_MN) {
#pragma unroll
            for (int j = 0; j < BN / 64; ++j) tma_load_2d(b_dst + j * 8192, &tmB, &full_bar[stage], n0 + 64 * j, k0);
          } else {
            tma_load_2d(b_dst, &tmB, &full_bar[stage], k0, n0);
          }
          if (++stage == GEMM_STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    constexpr uint32_t idesc = make_idesc_f16(GEMM_BM, BN, A_MN, B_MN);
    uint32_t stage = 0, phase = 0, acc = 0, acc_phase = 0;
    for (int u = blockIdx.x; u < units; u += gridDim.x) {
      int mt, nt, kb0, kb1;
      decode(u, mt, nt, kb0, kb1);
      mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + acc * BN;
      for (int kb = kb0; kb < kb1; ++kb) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        if (lane == 0) {
          const uint32_t a_addr = smem_u32(sA + stage * S::A_BYTES);
          const uint32_t b_addr = smem_u32(sB + stage * S::B_BYTES);
#pragma unroll
          for (int kk = 0; kk < GEMM_BK / 16; ++kk) {
            const uint64_t da = A_MN ? make_smem_desc(a_addr + kk * 2048, 8192, 1024) : make_smem_desc(a_addr + kk * 32, 0, 1024);
            const uint64_t db = B_MN ? make_smem_desc(b_addr + kk * 2048, 8192, 1024) : make_smem_desc(b_addr + kk * 32, 0, 1024);
            umma_ss(d_tmem, da, db, idesc, (kb > kb0 || kk > 0) ? 1u : 0u);
          }
          umma_commit(&empty_bar[stage]);
          if (kb == kb1 - 1) umma_commit(&tfull_bar[acc]);
        }
        __syncwarp();
        if (++stage == GEMM_STAGES) { stage = 0; phase ^= 1; }
      }
      if (kb1 <= kb0 && lane == 0) umma_commit(&tfull_bar[acc]);   // empty split: nothing accumulated
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
  } else {
    // ------------------------------------------------------------------ epilogue (warps 2..5)
    const int q = warp & 3;                       // TMEM lane quadrant this warp may read
    float* st = sEpi + (warp - 2) * 32 * GEMM_STAGE_PITCH;
    const float alpha = g.alpha ? __ldg(g.alpha) : 1.0f;
    uint32_t acc = 0, acc_phase = 0;
    for (int u = blockIdx.x; u < units; u += gridDim.x) {
      int mt, nt, kb0, kb1;
      decode(u, mt, nt, kb0, kb1);
      const int m0 = mt * GEMM_BM + q * 32, n0 = nt * BN;
      mbar_wait(&tfull_bar[acc], acc_phase);
      tc_fence_after();
      const bool has_data = kb1 > kb0;
      gemm_epilogue_tile<BN, EPI, OutT>(g, st, tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * BN, m0, n0, lane, alpha, has_data,
                                        [&] { mbar_arrive(&tempty_bar[acc]); });
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 2 * BN);
  }
}

}  // namespace b200
