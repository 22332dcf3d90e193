// Cost of one tcgen05.mma (cta_group::1, M = 128, K = 16, fp16 -> fp32) by N and operand majorness, issued back to back by one
// thread per SM exactly as the attention kernels issue them (same descriptors).  The attention backward issues 40 such
// instructions with N = 64 per 128 x 128 block; if an instruction costs the same for N = 64 as for N = 128, half of the tensor
// pipe's time is lost to the shape.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I spokennlp_b200/csrc -o tools/micro/mma_shapes tools/micro/mma_shapes.cu
#include <cstdio>
#include <cstdint>
#include "ptx.cuh"

using namespace b200;

// MODE 0: A K-major, B K-major  (S = Q K^T)        MODE 1: A MN-major, B MN-major (dV = P^T dO)
// MODE 2: A K-major, B MN-major (O = P V, dQ = dS K)
template <int N, int MODE>
__global__ void __launch_bounds__(128, 1) k(int reps, long long* out, int fill) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // operands: zeros, or small random fp16 values (fill != 0) — data toggling costs power, and a throttled tensor pipe would show here
  for (int i = threadIdx.x; i < 96 * 1024 / 4; i += blockDim.x) {
    uint32_t h = (i + 1) * 2654435761u;
    h ^= h >> 15; h *= 0x85EBCA6Bu; h ^= h >> 13;
    reinterpret_cast<uint32_t*>(smem)[i] = fill ? ((h & 0x83FF83FFu) | 0x38003800u) : 0u;      // +-[0.5, 1)
  }
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
  if (warp == 0) { tmem_alloc(&slot, 512); tmem_relinquish(); }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = slot;
  if (warp == 0 && lane == 0) {
    constexpr uint32_t idesc = make_idesc_f16(128, N, MODE == 1 ? 1 : 0, MODE == 0 ? 0 : 1);
    const uint32_t a = smem_u32(smem), b = smem_u32(smem + 32768);
    const long long t0 = clock64();
    for (int r = 0; r < reps; ++r) {
#pragma unroll
      for (int kk = 0; kk < 8; ++kk) {
        uint64_t da, db;
        if (MODE == 0) { da = make_smem_desc(a + (kk & 3) * 32 + (kk >> 2) * 16384, 0, 1024); db = make_smem_desc(b + (kk & 3) * 32 + (kk >> 2) * 16384, 0, 1024); }
        if (MODE == 1) { da = make_smem_desc(a + kk * 2048, 16384, 1024); db = make_smem_desc(b + kk * 2048, 8192, 1024); }
        if (MODE == 2) { da = make_smem_desc(a + (kk >> 2) * 16384 + (kk & 3) * 32, 0, 1024); db = make_smem_desc(b + kk * 2048, 8192, 1024); }
        umma_ss(tmem + (r & 1) * 256, da, db, idesc, kk > 0);
      }
    }
    umma_commit(&bar);
    mbar_wait(&bar, 0);
    const long long t1 = clock64();
    if (blockIdx.x == 0) out[0] = t1 - t0;
  }
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tmem, 512); }
}

template <int N, int MODE>
void run(const char* name, long long* out) {
  const int reps = 4096;
  cudaFuncSetAttribute(k<N, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  for (int fill : {0, 1}) {
    const int ctas = 148;
    k<N, MODE><<<ctas, 128, 100 * 1024>>>(reps, out, fill);
    k<N, MODE><<<ctas, 128, 100 * 1024>>>(reps, out, fill);
    cudaDeviceSynchronize();
    long long h;
    cudaMemcpy(&h, out, 8, cudaMemcpyDeviceToHost);
    const double per = double(h) / (reps * 8);
    printf("%-26s N=%3d %s  %7.1f clk per MMA  (%5.1f %% of 8192 flop/clk/SM)  %s\n", name, N, fill ? "random" : "zeros ", per,
           100.0 * (2.0 * 128 * N * 16 / per) / 8192.0, cudaGetErrorString(cudaGetLastError()));
  }
}

int main() {
  long long* out;
  cudaMalloc(&out, 8);
  run<64, 0>("A K-major,  B K-major", out);
  run<128, 0>("A K-major,  B K-major", out);
  run<256, 0>("A K-major,  B K-major", out);
  run<64, 1>("A MN-major, B MN-major", out);
  run<128, 1>("A MN-major, B MN-major", out);
  run<64, 2>("A K-major,  B MN-major", out);
  run<128, 2>("A K-major,  B MN-major", out);
  run<256, 2>("A K-major,  B MN-major", out);
  return 0;
}
