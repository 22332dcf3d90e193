// Thin inline-PTX layer for sm_100a: mbarrier, TMA, tcgen05 (MMA / TMEM), descriptors.
// Everything here is hand-written; CUTLASS/CuTe is not used.  Bit layouts follow the
// PTX ISA "tcgen05 matrix descriptors" / "instruction descriptor" tables.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace b200 {

#ifndef B200_SPIN_LIMIT_CYCLES
#define B200_SPIN_LIMIT_CYCLES (6000000000ll)   // ~3-4 s at B200 clocks: turn a protocol deadlock into a trap
#endif

// Timeline instrumentation for tools/attn_trace.py (a separate -DB200_ATT_TRACE build, never the product library): lane 0 of
// every warp of CTA 0 appends (event, clock) records to its own buffer (a plain store: an atomic slot counter cost ~600 clk per
// record and bent the timeline).  Compiled out otherwise.
#ifdef B200_ATT_TRACE
constexpr int kAttTraceCap = 4096;                       // records per warp
__device__ unsigned long long g_att_trace[16 * kAttTraceCap];
__device__ unsigned int g_att_trace_n[16];
#define ATT_TRACE_INIT unsigned int att_trace_i_ = 0
#define ATT_TRACE(ev)                                                                                                    \
  do {                                                                                                                   \
    if (blockIdx.x == 0 && (threadIdx.x & 31) == 0 && att_trace_i_ < kAttTraceCap)                                       \
      g_att_trace[(threadIdx.x >> 5) * kAttTraceCap + att_trace_i_++] =                                                  \
          (static_cast<unsigned long long>(ev) << 48) | (static_cast<unsigned long long>(clock64()) & 0xFFFFFFFFFFFFull); \
  } while (0)
#define ATT_TRACE_FINI                                                                                                   \
  do {                                                                                                                   \
    if (blockIdx.x == 0 && (threadIdx.x & 31) == 0) g_att_trace_n[threadIdx.x >> 5] = att_trace_i_;                      \
  } while (0)
#else
#define ATT_TRACE_INIT
#define ATT_TRACE(ev)
#define ATT_TRACE_FINI
#endif

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ uint32_t lane_id() { return threadIdx.x & 31; }

// ----------------------------------------------------------------------------- mbarrier
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ uint32_t mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok;
}
// Bounded wait: a pipeline-protocol bug must surface as a launch failure, never as a hung GPU.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  uint32_t n = 0;
  while (!mbar_try_wait(bar, parity)) {
    if (((++n) & 0xFFF) == 0 && (clock64() - t0) > B200_SPIN_LIMIT_CYCLES) {
      printf("b200: mbarrier wait timed out (block %d,%d thread %d bar@%u parity %u)\n", blockIdx.x, blockIdx.y,
             threadIdx.x, smem_u32(bar), parity);
      __trap();
    }
  }
}

// One lane of a converged warp.  Guarding a tcgen05.mma / tcgen05.commit sequence with this instead of `lane == 0` matters: ptxas
// knows a single thread is active and moves the operands to uniform registers directly (4-5 instructions per MMA); after a lane
// compare it wraps every MMA in an ELECT / broadcast loop with a branch (10+ instructions, ~90 clk per MMA in the attention
// backward, whose single issuing warp then set the pace of the whole kernel: 40 MMAs x 90 clk per 3.9 k clk block, r02r).
__device__ __forceinline__ bool elect_one_sync() {
  uint32_t p;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(p));
  return p != 0;
}

// explicit shared-window accesses (the dynamic-smem base is re-aligned by hand, so the compiler cannot prove the
// address space and would otherwise emit slower generic LD/ST)
__device__ __forceinline__ void sts128(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ uint4 lds128(uint32_t addr) {
  uint4 r;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "r"(addr) : "memory");
  return r;
}
__device__ __forceinline__ float lds_f32(uint32_t addr) {
  float r;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(r) : "r"(addr) : "memory");
  return r;
}
__device__ __forceinline__ void sts_f32(uint32_t addr, float v) {
  asm volatile("st.shared.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory");
}

// generic-proxy smem writes -> visible to the async proxy (TMA / tcgen05.mma reads)
__device__ __forceinline__ void fence_proxy_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

// ----------------------------------------------------------------------------- TMA
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* m) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
// 2D tiled load: c0 = coordinate along the contiguous dim, c1 = row coordinate.
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      :: "r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d_hint(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1,
                                                 uint64_t policy) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint"
      " [%0], [%1, {%3, %4}], [%2], %5;"
      :: "r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "l"(policy)
      : "memory");
}
constexpr uint64_t kEvictFirst = 0x12F0000000000000ull;
constexpr uint64_t kEvictLast = 0x14F0000000000000ull;

// TMA store / reduce (smem -> global), bulk-group completion
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* m, const void* smem_src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
               ::"l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(smem_src)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_reduce_add_2d(const CUtensorMap* m, const void* smem_src, int c0, int c1) {
  asm volatile("cp.reduce.async.bulk.tensor.2d.global.shared::cta.add.bulk_group [%0, {%2, %3}], [%1];"
               ::"l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(smem_src)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_commit_group() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void tma_wait_group_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void tma_wait_group_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

// ----------------------------------------------------------------------------- tcgen05 / TMEM
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_result, uint32_t ncols) {   // whole warp
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_result)), "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {        // whole warp
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// D[tmem] (+)= A[smem] * B[smem];  issued by ONE thread.
__device__ __forceinline__ void umma_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      :: "r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// D[tmem] (+)= A[tmem] * B[smem]
__device__ __forceinline__ void umma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      :: "r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// All previously issued tcgen05.mma of this thread arrive on `bar` when complete
// (implies tcgen05.fence::before_thread_sync).
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// TMEM -> registers, shape 32x32b: lane i of the warp reads TMEM lane (base_lane + i), N consecutive 32-bit columns.
__device__ __forceinline__ void tmem_ld_x32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_x16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_st_x16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
      :: "r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
         "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}

// ----------------------------------------------------------------------------- descriptors
// Shared-memory matrix descriptor, SWIZZLE_128B, sm_100 version field = 1.
//   bits [0,14)  start address >> 4        bits [16,30) leading-dim byte offset >> 4
//   bits [32,46) stride-dim byte offset>>4 bits [46,48) version (1)   bits [61,64) layout (2 = SWIZZLE_128B)
// K-major tile  (rows of 64 halves = 128 B, 8-row 1024 B atoms):  SBO = 1024, LBO unused.
// MN-major tile (64-half-wide column blocks, each [k rows][128 B]): SBO = 1024 (next 8 k rows),
//                LBO = bytes between 64-wide MN blocks.
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((saddr & 0x3FFFFu) >> 4);
  d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFFu) << 32;
  d |= 1ull << 46;
  d |= 2ull << 61;
  return d;
}

// Instruction descriptor for kind::f16, fp16 (or bf16) A/B, fp32 accumulate.
//   [4,6) c_format (1 = f32)  [7,10) a_format (0 = f16, 1 = bf16)  [10,13) b_format  [15] a_major  [16] b_major (1 = MN-major)
//   [17,23) N >> 3   [24,29) M >> 4
__host__ __device__ constexpr uint32_t make_idesc_f16(int M, int N, int a_mn_major, int b_mn_major, int bf16_operands = 0) {
  return (1u << 4) | (static_cast<uint32_t>(bf16_operands) << 7) | (static_cast<uint32_t>(bf16_operands) << 10) | (static_cast<uint32_t>(a_mn_major) << 15) |
         (static_cast<uint32_t>(b_mn_major) << 16) | (static_cast<uint32_t>(N >> 3) << 17) |
         (static_cast<uint32_t>(M >> 4) << 24);
}

// fire-and-forget vectorised fp32 reduction into global memory (split-K wgrad, dQ accumulation)
__device__ __forceinline__ void red_add_v4(float* p, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

// ----------------------------------------------------------------------------- dropout (counter-based, stateless)
// The keep/drop decision of element `idx` of a tensor is a pure function of (seed, idx): the backward regenerates the
// forward's mask instead of storing it.  One 32-bit hash decides TWO consecutive elements (15 bits each: bits [0,15) for
// the even element, bits [16,31) for the odd one; keep iff the lane >= thr15 = round(p * 32768), so p = 0.1 is met to
// 6e-6).  `seed` = per-step base seed (device memory, so a replayed CUDA graph still draws fresh masks) mixed with a
// per-site constant.  The index enters by ADDITION, so a kernel that walks consecutive pairs of one row pays one integer
// add per pair for (idx + seed) * C1 (see drop_premix / drop_z) instead of an xor and a multiply.
struct DropCfg {
  const uint32_t* seed_base;   // device pointer, null = dropout off
  uint32_t site;               // which dropout site of the model (layer * 8 + kind)
  uint32_t thr15;              // round(p * 32768)
  float scale;                 // 1 / (1 - p)
};
constexpr uint32_t kDropC1 = 0x9E3779B1u, kDropC2 = 0x85EBCA6Bu;
__device__ __forceinline__ uint32_t drop_premix(uint32_t pair_idx, uint32_t seed) { return (pair_idx + seed) * kDropC1; }
__device__ __forceinline__ uint32_t drop_finish(uint32_t h) {      // h = drop_premix(...)
  h ^= h >> 15;
  h *= kDropC2;
  h ^= h >> 13;
  return h;
}
__device__ __forceinline__ uint32_t drop_hash(uint32_t pair_idx, uint32_t seed) { return drop_finish(drop_premix(pair_idx, seed)); }
__device__ __forceinline__ uint32_t drop_seed(const DropCfg& d) {
  return drop_hash(d.site * 0x632BE5ABu + 0x7F4A7C15u, __ldg(d.seed_base));
}
// multipliers (0 or scale) for elements 2*pair_idx and 2*pair_idx+1
__device__ __forceinline__ void drop_pair(uint32_t pair_idx, uint32_t seed, uint32_t thr15, float scale, float& m0, float& m1) {
  const uint32_t h = drop_hash(pair_idx, seed);
  m0 = (h & 0x7FFFu) >= thr15 ? scale : 0.f;
  m1 = ((h >> 16) & 0x7FFFu) >= thr15 ? scale : 0.f;
}
__device__ __forceinline__ float drop_one(uint32_t idx, uint32_t seed, uint32_t thr15, float scale) {
  const uint32_t h = drop_hash(idx >> 1, seed);
  return (((idx & 1) ? (h >> 16) : h) & 0x7FFFu) >= thr15 ? scale : 0.f;
}
// Both decisions of a pair at once (SWAR), for the attention kernels where the mask costs more issue slots than the exp:
// z = ((lane | 0x8000) - thr15) per 16-bit lane, so bit 15 / bit 31 of z say "keep" (no borrow crosses the lanes because
// every lane is >= 0x8000 before the subtraction).  tt = thr15 * 0x00010001.
__device__ __forceinline__ uint32_t drop_z(uint32_t premixed, uint32_t tt) {
  uint32_t h = premixed;
  h ^= h >> 15;
  h *= kDropC2;
  return ((h ^ (h >> 13)) | 0x80008000u) - tt;
}
// 0xFFFF in every 16-bit lane that is kept (for packed half2 data): byte-permute with sign replication of bytes 1 and 3
__device__ __forceinline__ uint32_t drop_keep_h2(uint32_t z) {
  uint32_t m;
  asm("prmt.b32 %0, %1, %1, 0xBB99;" : "=r"(m) : "r"(z));
  return m;
}
// 0xFFFFFFFF if the even (lo) / odd (hi) element of the pair is kept
__device__ __forceinline__ uint32_t drop_keep_lo(uint32_t z) {
  uint32_t m;
  asm("prmt.b32 %0, %1, %1, 0x9999;" : "=r"(m) : "r"(z));
  return m;
}
__device__ __forceinline__ uint32_t drop_keep_hi(uint32_t z) {
  uint32_t m;
  asm("prmt.b32 %0, %1, %1, 0xBBBB;" : "=r"(m) : "r"(z));
  return m;
}

// Attention-probability sites draw ONE hash per FOUR consecutive keys (7-bit lanes, one per byte): the persistent attention
// kernels are instruction-issue bound (profiles/r02a: 0.46 warp instructions per cycle per scheduler, ~0.5 is what 3-register
// operand instructions sustain), and the pair-wise mask above cost 8.5 instructions per pair — more than the softmax itself.
//   quad index = ((b * heads + h) * Sq + q) * ceil(Sk / 4) + key / 4;   lane l = key & 3 -> bits [8l, 8l + 7) of the hash;
//   element kept iff lane >= thr[l], the four thresholds dithered so that their sum is round(512 p) (p = 0.1: 13,13,13,12 ->
//   an effective drop rate of 51/512 = 0.0996); kept elements are scaled by 512 / (512 - sum) — unbiased for that rate.
// DropCfg carries the packed thresholds in `thr15` and that scale in `scale` for these sites (api.cu: make_drop_attn).
__device__ __forceinline__ uint32_t drop4_z(uint32_t premixed, uint32_t tt) {   // bit 7 of byte l of the result = "keep lane l"
  uint32_t h = premixed;
  h ^= h >> 15;
  h *= kDropC2;
  return ((h ^ (h >> 13)) | 0x80808080u) - tt;       // every byte is >= 0x80 before the subtraction: no borrow crosses the lanes
}
template <uint32_t SEL>
__device__ __forceinline__ uint32_t prmt_sel(uint32_t z) {                      // byte permute with sign replication (selector nibble | 8)
  uint32_t m;
  asm("prmt.b32 %0, %1, %1, %2;" : "=r"(m) : "r"(z), "n"(SEL));
  return m;
}
// 0xFFFF per kept half of the packed pair (lanes 0,1 / lanes 2,3); 0xFFFFFFFF if lane l is kept
__device__ __forceinline__ uint32_t drop4_keep_h2_lo(uint32_t z) { return prmt_sel<0x9988>(z); }
__device__ __forceinline__ uint32_t drop4_keep_h2_hi(uint32_t z) { return prmt_sel<0xBBAA>(z); }
template <int L>
__device__ __forceinline__ uint32_t drop4_keep(uint32_t z) { return prmt_sel<0x8888u + 0x1111u * L>(z); }

// ----------------------------------------------------------------------------- packed fp32 pairs (FFMA2 / FADD2 / FMUL2)
// sm_100 executes two fp32 operations per instruction on an aligned register pair: half the issue slots for the softmax /
// epilogue arithmetic of kernels that are issue-bound.  pack2 / unpack2 are register renames when the pair is adjacent.
__device__ __forceinline__ uint64_t pack2(float a, float b) {
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
  return r;
}
__device__ __forceinline__ uint64_t pack2u(uint32_t a, uint32_t b) {
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "r"(a), "r"(b));
  return r;
}
__device__ __forceinline__ void unpack2(uint64_t v, float& a, float& b) { asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ uint64_t fma2(uint64_t a, uint64_t b, uint64_t c) {
  uint64_t d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ uint64_t add2(uint64_t a, uint64_t b) {
  uint64_t d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ uint64_t mul2(uint64_t a, uint64_t b) {
  uint64_t d;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}

// ----------------------------------------------------------------------------- small math helpers
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

__device__ __forceinline__ float fast_exp2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// erf via the single-precision rational approximation x*P(x^2)/Q(x^2) on |x| <= 4 (abs err < 2e-7 after
// clamping): ~13 FMA + 1 reciprocal, no branches — cheap enough for a GEMM epilogue.
__device__ __forceinline__ float fast_erf(float x) {
  x = fminf(fmaxf(x, -4.0f), 4.0f);
  const float x2 = x * x;
  float p = -2.72614225801306e-10f;
  p = fmaf(p, x2, 2.77068142495902e-08f);
  p = fmaf(p, x2, -2.10102402082508e-06f);
  p = fmaf(p, x2, -5.69250639462346e-05f);
  p = fmaf(p, x2, -7.34990630326855e-04f);
  p = fmaf(p, x2, -2.95459980854025e-03f);
  p = fmaf(p, x2, -1.60960333262415e-02f);
  p *= x;
  float q = -1.45660718464996e-05f;
  q = fmaf(q, x2, -2.13374055278905e-04f);
  q = fmaf(q, x2, -1.68282697438203e-03f);
  q = fmaf(q, x2, -7.37332916720468e-03f);
  q = fmaf(q, x2, -1.42647390514189e-02f);
  return __fdividef(p, q);
}
__device__ __forceinline__ float fast_rcp(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// Exact (erf) GELU and its derivative for GEMM epilogues, ~17 instructions per element for BOTH values:
//   Phi(x) = 1 - 0.5 * poly(t) * exp(-x^2/2),  t = 1/(1 + p|x|/sqrt2)   (Abramowitz-Stegun 7.1.26 applied to erf(x/sqrt2))
// so the single exponential exp(-x^2/2) serves the erf tail AND the Gaussian pdf of the derivative.  Max abs error vs
// the exact functions: 4.2e-7 (gelu), 2.7e-7 (gelu') — three orders of magnitude below the fp16 rounding of the outputs.
__device__ __forceinline__ void gelu_erf_both(float x, float& y, float& dy) {
  const float t = fast_rcp(fmaf(fabsf(x), 0.2316418882f, 1.0f));
  const float he = 0.5f * fast_exp2(x * x * -0.7213475204444817f);            // 0.5 * exp(-x^2/2)
  float poly = fmaf(t, 1.061405429f, -1.453152027f);
  poly = fmaf(poly, t, 1.421413741f);
  poly = fmaf(poly, t, -0.284496736f);
  poly = fmaf(poly, t, 0.254829592f);
  const float tail = poly * t * he;                                           // 1 - Phi(|x|)
  const float cdf = x >= 0.f ? 1.0f - tail : tail;
  y = x * cdf;
  dy = fmaf(x, 0.7978845608028654f * he, cdf);                                // Phi(x) + x * phi(x)
}
// Two elements at a time on packed fp32 pairs: the GEMM epilogue that applies this is issue-bound (8 epilogue warps, ~0.5
// instructions per cycle per scheduler), and 21.5 scalar instructions per element left half of it exposed behind a K = 768
// mainloop.  Same arithmetic as gelu_erf_both; the 0.5 is folded into the exponent, the sign select into a
// sign-bit transfer: cdf = 0.5 + copysign(0.5 - tail, x).
__device__ __forceinline__ void gelu_erf_both2(uint64_t x2, uint64_t& y2, uint64_t& dy2) {
  float x0, x1;
  unpack2(x2, x0, x1);
  const float t0 = fast_rcp(fmaf(fabsf(x0), 0.2316418882f, 1.0f)), t1 = fast_rcp(fmaf(fabsf(x1), 0.2316418882f, 1.0f));
  const uint64_t t2 = pack2(t0, t1);
  float a0, a1;
  unpack2(fma2(mul2(x2, x2), pack2(-0.7213475204444817f, -0.7213475204444817f), pack2(-1.0f, -1.0f)), a0, a1);
  const uint64_t he2 = pack2(fast_exp2(a0), fast_exp2(a1));                   // 0.5 * exp(-x^2/2)
  uint64_t poly = fma2(t2, pack2(1.061405429f, 1.061405429f), pack2(-1.453152027f, -1.453152027f));
  poly = fma2(poly, t2, pack2(1.421413741f, 1.421413741f));
  poly = fma2(poly, t2, pack2(-0.284496736f, -0.284496736f));
  poly = fma2(poly, t2, pack2(0.254829592f, 0.254829592f));
  const uint64_t tail = mul2(mul2(poly, t2), he2);                            // 1 - Phi(|x|)
  float q0, q1;
  unpack2(fma2(tail, pack2(-1.0f, -1.0f), pack2(0.5f, 0.5f)), q0, q1);         // 0.5 - tail  (>= 0)
  q0 = __uint_as_float(__float_as_uint(q0) ^ (__float_as_uint(x0) & 0x80000000u));
  q1 = __uint_as_float(__float_as_uint(q1) ^ (__float_as_uint(x1) & 0x80000000u));
  const uint64_t cdf = add2(pack2(q0, q1), pack2(0.5f, 0.5f));
  y2 = mul2(x2, cdf);
  dy2 = fma2(x2, mul2(he2, pack2(0.7978845608028654f, 0.7978845608028654f)), cdf);      // Phi(x) + x * phi(x)
}
__device__ __forceinline__ float gelu_erf(float x) {
  float y, dy;
  gelu_erf_both(x, y, dy);
  return y;
}
__device__ __forceinline__ float gelu_erf_grad(float x) {
  float y, dy;
  gelu_erf_both(x, y, dy);
  return dy;
}

}  // namespace b200
