// Micro-benchmarks that decide the attention kernels' design (B200): TMEM read bandwidth per SM, MUFU.EX2 rate,
// and the fixed cost of a 1-CTA/SM launch wave with ~200 KB of shared memory and a full TMEM allocation.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../spokennlp_b200/csrc/ptx.cuh"
using namespace b200;

__global__ void __launch_bounds__(512, 1) tmem_read_kernel(int iters, int nwarps, long long* out, float* sink) {
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) { tmem_alloc(&slot, 512); tmem_relinquish(); }
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tmem = slot;
  const uint32_t lane_addr = static_cast<uint32_t>((warp & 3) * 32) << 16;
  float acc = 0.f;
  __syncthreads();
  const long long t0 = clock64();
  if (warp < nwarps) {
    for (int it = 0; it < iters; ++it) {
      uint32_t v[128];
      tmem_ld_x32(tmem + lane_addr + ((warp >> 2) & 1) * 128, *reinterpret_cast<uint32_t(*)[32]>(&v[0]));
      tmem_ld_x32(tmem + lane_addr + ((warp >> 2) & 1) * 128 + 32, *reinterpret_cast<uint32_t(*)[32]>(&v[32]));
      tmem_ld_x32(tmem + lane_addr + ((warp >> 2) & 1) * 128 + 64, *reinterpret_cast<uint32_t(*)[32]>(&v[64]));
      tmem_ld_x32(tmem + lane_addr + ((warp >> 2) & 1) * 128 + 96, *reinterpret_cast<uint32_t(*)[32]>(&v[96]));
      tmem_wait_ld();
#pragma unroll
      for (int i = 0; i < 128; i += 16) acc += __uint_as_float(v[i]);
    }
  }
  __syncthreads();
  const long long t1 = clock64();
  if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = t1 - t0;
  if (acc == 123.456f) sink[0] = acc;
  tc_fence_before(); __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tmem, 512); }
}

__global__ void __launch_bounds__(512, 1) mufu_kernel(int iters, int nwarps, long long* out, float* sink) {
  const int warp = threadIdx.x >> 5;
  float x[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) x[i] = -0.001f * (threadIdx.x + i);
  __syncthreads();
  const long long t0 = clock64();
  if (warp < nwarps) {
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int i = 0; i < 16; ++i) x[i] = fast_exp2(x[i]) - 1.0f;
    }
  }
  __syncthreads();
  const long long t1 = clock64();
  if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = t1 - t0;
  float acc = 0.f;
#pragma unroll
  for (int i = 0; i < 16; ++i) acc += x[i];
  if (acc == 123.456f) sink[0] = acc;
}

// pack-convert rate
__global__ void __launch_bounds__(512, 1) f2fp_kernel(int iters, int nwarps, long long* out, float* sink) {
  const int warp = threadIdx.x >> 5;
  float x[16];
  uint32_t acc = 0;
#pragma unroll
  for (int i = 0; i < 16; ++i) x[i] = 0.001f * (threadIdx.x + i);
  __syncthreads();
  const long long t0 = clock64();
  if (warp < nwarps) {
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int i = 0; i < 16; i += 2) {
        const __half2 h = __floats2half2_rn(x[i], x[i + 1]);
        acc ^= *reinterpret_cast<const uint32_t*>(&h);
        x[i] += 1.0f;
      }
    }
  }
  __syncthreads();
  const long long t1 = clock64();
  if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = t1 - t0;
  if (acc == 123456u) sink[0] = 1.f;
}

__global__ void __launch_bounds__(352, 1) empty_cta_kernel(int alloc_tmem) {
  extern __shared__ uint8_t smem[];
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  if (alloc_tmem) {
    if (warp == 1) { tmem_alloc(&slot, 512); tmem_relinquish(); }
    tc_fence_before(); __syncthreads(); tc_fence_after();
    const uint32_t tmem = slot;
    smem[threadIdx.x] = 1;
    tc_fence_before(); __syncthreads();
    if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem, 512); }
  } else {
    smem[threadIdx.x] = 1;
  }
}

int main() {
  long long* out; float* sink;
  cudaMalloc(&out, 8); cudaMalloc(&sink, 4);
  long long h;
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  const double ghz = 1.0;  // report cycles
  for (int nw : {4, 8, 16}) {
    const int iters = 2000;
    tmem_read_kernel<<<148, 512>>>(iters, nw, out, sink);
    tmem_read_kernel<<<148, 512>>>(iters, nw, out, sink);
    cudaDeviceSynchronize();
    cudaMemcpy(&h, out, 8, cudaMemcpyDeviceToHost);
    const double bytes = double(nw) * iters * 128 * 32 * 4;
    printf("tmem_read  warps=%2d: %.1f B/clk/SM (%lld cycles; one 128x128 fp32 tile = %.0f cycles) %s\n", nw, bytes / h, h, 65536.0 / (bytes / h), cudaGetErrorString(cudaGetLastError()));
  }
  for (int nw : {4, 8, 16}) {
    const int iters = 2000;
    mufu_kernel<<<148, 512>>>(iters, nw, out, sink);
    mufu_kernel<<<148, 512>>>(iters, nw, out, sink);
    cudaDeviceSynchronize();
    cudaMemcpy(&h, out, 8, cudaMemcpyDeviceToHost);
    printf("mufu.ex2   warps=%2d: %.2f ex2/clk/SM (%lld cycles) %s\n", nw, double(nw) * 32 * iters * 16 / h, h, cudaGetErrorString(cudaGetLastError()));
  }
  for (int nw : {4, 8, 16}) {
    const int iters = 2000;
    f2fp_kernel<<<148, 512>>>(iters, nw, out, sink);
    f2fp_kernel<<<148, 512>>>(iters, nw, out, sink);
    cudaDeviceSynchronize();
    cudaMemcpy(&h, out, 8, cudaMemcpyDeviceToHost);
    printf("f2fp.pack  warps=%2d: %.2f cvt-pairs/clk/SM (%lld cycles) %s\n", nw, double(nw) * 32 * iters * 8 / h, h, cudaGetErrorString(cudaGetLastError()));
  }
  cudaFuncSetAttribute(empty_cta_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int alloc : {0, 1}) for (int smem : {0, 200 * 1024}) for (int ctas : {148, 768, 1536}) {
    empty_cta_kernel<<<ctas, 352, smem>>>(alloc);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    for (int i = 0; i < 20; ++i) empty_cta_kernel<<<ctas, 352, smem>>>(alloc);
    cudaEventRecord(e1);
    cudaDeviceSynchronize();
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    printf("empty CTAs tmem_alloc=%d smem=%6d ctas=%4d: %.2f us per launch (%.2f us per wave) %s\n", alloc, smem, ctas, ms * 1000 / 20, ms * 1000 / 20 / ((ctas + 147) / 148.0), cudaGetErrorString(cudaGetLastError()));
  }
  (void)ghz;
  return 0;
}
