// The forward softmax block body (attn_fwd3.cuh) in isolation: 128 scores per thread from TMEM (or registers), row max,
// exp2 on packed pairs, row sum, fp16 pack (+ quad dropout), swizzled STS — no MMA, no barriers.  Cycles per block per
// warp for 4 / 8 warps per SM (one / two per sub-partition), with parts switched off one at a time: says whether the instruction mix alone reaches the
// MUFU bound (128 EX2 x 8 clk per warp-block, two warps per sub-partition -> 2048 clk per iteration).
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I spokennlp_b200/csrc -o tools/micro/softmax_body tools/micro/softmax_body.cu
#include <cstdio>
#include <cstdint>
#include "ptx.cuh"

using namespace b200;

enum { F_TMEM = 1, F_MAX = 2, F_EXP = 4, F_SUM = 8, F_PACK = 16, F_STS = 32, F_DROP = 64, F_FENCE = 128 };

template <int F>
__global__ void __launch_bounds__(256, 1) body(int iters, long long* out, float* sink, float sc) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0) { tmem_alloc(&tmem_slot, 512); tmem_relinquish(); }
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  const int x = (warp >> 2) & 1, qd = warp & 3, r = qd * 32 + lane;
  const uint32_t lane_addr = static_cast<uint32_t>(qd * 32) << 16;
  const uint32_t p_row = smem_u32(smem) + x * 32768 + r * 128;
  uint32_t v[128];
#pragma unroll
  for (int i = 0; i < 128; ++i) v[i] = __float_as_uint(-0.01f * ((threadIdx.x * 7 + i * 13) & 255));
  float m = -1e30f, l = 0.f;
  uint32_t dpre = threadIdx.x * 2654435761u;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    if (F & F_TMEM) {
      tmem_ld_x32(tmem + lane_addr + x * 128, *reinterpret_cast<uint32_t(*)[32]>(&v[0]));
      tmem_ld_x32(tmem + lane_addr + x * 128 + 32, *reinterpret_cast<uint32_t(*)[32]>(&v[32]));
      tmem_ld_x32(tmem + lane_addr + x * 128 + 64, *reinterpret_cast<uint32_t(*)[32]>(&v[64]));
      tmem_ld_x32(tmem + lane_addr + x * 128 + 96, *reinterpret_cast<uint32_t(*)[32]>(&v[96]));
      tmem_wait_ld();
    } else {
#pragma unroll
      for (int i = 0; i < 128; ++i) asm volatile("" : "+r"(v[i]));
    }
    float m_use = m;
    if (F & F_MAX) {
      float m0 = fmaxf(__uint_as_float(v[0]), __uint_as_float(v[1])), m1 = fmaxf(__uint_as_float(v[2]), __uint_as_float(v[3]));
#pragma unroll
      for (int i = 4; i < 128; i += 4) {
        m0 = fmaxf(m0, fmaxf(__uint_as_float(v[i]), __uint_as_float(v[i + 1])));
        m1 = fmaxf(m1, fmaxf(__uint_as_float(v[i + 2]), __uint_as_float(v[i + 3])));
      }
      const float mx = fmaxf(m0, m1) * sc;
      const float m_new = fmaxf(m, mx);
      const bool grow = (m_new - m) > 8.0f;
      const bool rescale = __any_sync(0xffffffffu, grow);
      m_use = rescale ? m_new : m;
      m = m_use;
    }
    const uint64_t nm2 = pack2(-m_use, -m_use), sc2 = pack2(sc, sc);
    uint64_t rs2 = pack2(0.f, 0.f);
    const uint32_t dpre_j = dpre + static_cast<uint32_t>(it * 32) * kDropC1;
#pragma unroll
    for (int ch = 0; ch < 16; ++ch) {
      uint32_t pk[4], z[2];
      if (F & F_DROP) {
        z[0] = drop4_z(dpre_j + static_cast<uint32_t>(2 * ch) * kDropC1, 0x0d0d0d0cu);
        z[1] = drop4_z(dpre_j + static_cast<uint32_t>(2 * ch + 1) * kDropC1, 0x0d0d0d0cu);
      }
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int i = ch * 8 + 2 * e;
        float x0, x1;
        unpack2(fma2(pack2u(v[i], v[i + 1]), sc2, nm2), x0, x1);
        float p0 = x0, p1 = x1;
        if (F & F_EXP) { p0 = fast_exp2(x0); p1 = fast_exp2(x1); }
        if (F & F_SUM) rs2 = add2(rs2, pack2(p0, p1));
        if (F & F_PACK) {
          const __half2 hp = __floats2half2_rn(p0, p1);
          pk[e] = *reinterpret_cast<const uint32_t*>(&hp);
        } else {
          pk[e] = __float_as_uint(p0) ^ __float_as_uint(p1);
        }
        if (F & F_DROP) pk[e] &= (e & 1) ? drop4_keep_h2_hi(z[e >> 1]) : drop4_keep_h2_lo(z[e >> 1]);
      }
      if (F & F_STS) sts128(p_row + (ch >> 3) * 16384 + (((ch & 7) ^ (r & 7)) << 4), pk[0], pk[1], pk[2], pk[3]);
      else asm volatile("" ::"r"(pk[0]), "r"(pk[1]), "r"(pk[2]), "r"(pk[3]));
    }
    float rs0, rs1;
    unpack2(rs2, rs0, rs1);
    l += rs0 + rs1;
    if (F & F_FENCE) { fence_proxy_async_smem(); tc_fence_before(); __syncwarp(); }
  }
  __syncthreads();
  const long long t1 = clock64();
  if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = t1 - t0;
  if (l == 123.456f) sink[0] = l + m;
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

template <int F>
void run(const char* name, long long* out, float* sink) {
  const int iters = 2000;
  cudaFuncSetAttribute(body<F>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
  for (int threads : {128, 256}) {
    body<F><<<148, threads, 160 * 1024>>>(iters, out, sink, 0.18f);
    body<F><<<148, threads, 160 * 1024>>>(iters, out, sink, 0.18f);
    cudaDeviceSynchronize();
    long long h;
    cudaMemcpy(&h, out, 8, cudaMemcpyDeviceToHost);
    printf("%-44s %2d warps/SM  %8.1f clk per block-iteration  %s\n", name, threads / 32, double(h) / iters, cudaGetErrorString(cudaGetLastError()));
  }
}

int main() {
  long long* out; float* sink;
  cudaMalloc(&out, 8); cudaMalloc(&sink, 4);
  constexpr int ALL = F_TMEM | F_MAX | F_EXP | F_SUM | F_PACK | F_STS | F_FENCE;
  run<ALL>("full body", out, sink);
  run<ALL | F_DROP>("full body + dropout", out, sink);
  run<ALL & ~F_TMEM>("no tmem load", out, sink);
  run<ALL & ~F_MAX>("no max", out, sink);
  run<ALL & ~F_EXP>("no exp2", out, sink);
  run<ALL & ~F_SUM>("no row sum", out, sink);
  run<ALL & ~F_PACK>("no fp16 pack", out, sink);
  run<ALL & ~F_STS>("no sts", out, sink);
  run<ALL & ~F_FENCE>("no fence", out, sink);
  run<F_EXP>("exp2 + ffma2 only", out, sink);
  run<F_EXP | F_PACK | F_STS>("exp2 + pack + sts", out, sink);
  run<F_EXP | F_SUM | F_MAX>("exp2 + sum + max", out, sink);
  return 0;
}
