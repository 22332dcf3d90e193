// Which SM pipe do the softmax-loop instructions occupy?  Rates per SM for dependent chains (16 independent chains per
// thread, 16 warps per SM) of MUFU.EX2, F2FP.PACK, FMNMX3, FFMA, FADD and 1:1 mixes; a mix that costs the SUM of its parts
// shares a pipe, a mix that costs the MAX does not.
#include <cstdio>
#include <cstdint>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

__device__ __forceinline__ float ex2(float x) { float y; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ uint32_t pack(float a, float b) { uint32_t r; asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ float fmax3(float a, float b, float c) { return fmaxf(a, fmaxf(b, c)); }

template <int MODE>
__global__ void __launch_bounds__(512, 1) k(int iters, long long* out, float* sink) {
  float x[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) x[i] = -0.001f * (threadIdx.x + i + 1);
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      if (MODE == 0) x[i] = ex2(x[i]);
      if (MODE == 1) x[i] = __uint_as_float(pack(x[i], x[i]));
      if (MODE == 2) { x[i] = ex2(x[i]); x[(i + 8) & 15] = __uint_as_float(pack(x[(i + 8) & 15], x[(i + 8) & 15])); }
      if (MODE == 3) x[i] = fmax3(x[i], x[(i + 1) & 15], x[(i + 2) & 15]);
      if (MODE == 4) x[i] = fmaf(x[i], 1.0001f, 0.5f);
      if (MODE == 5) x[i] = x[i] + 1.5f;
      if (MODE == 6) { x[i] = ex2(x[i]); x[(i + 8) & 15] = fmaf(x[(i + 8) & 15], 1.0001f, 0.5f); }
      if (MODE == 7) { x[i] = __uint_as_float(pack(x[i], x[i])); x[(i + 8) & 15] = fmaf(x[(i + 8) & 15], 1.0001f, 0.5f); }
      if (MODE == 8) { x[i] = __uint_as_float(pack(x[i], x[i])); x[(i + 8) & 15] = fmax3(x[(i + 8) & 15], x[(i + 9) & 15], x[(i + 10) & 15]); }
      if (MODE == 9) { x[i] = fmaf(x[i], 1.0001f, 0.5f); x[(i + 8) & 15] = x[(i + 8) & 15] + 1.5f; }
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

template <int MODE>
void run(const char* name, int per_iter, long long* out, float* sink) {
  const int iters = 1000;
  k<MODE><<<148, 512>>>(iters, out, sink);
  k<MODE><<<148, 512>>>(iters, out, sink);
  cudaDeviceSynchronize();
  long long h;
  cudaMemcpy(&h, out, 8, cudaMemcpyDeviceToHost);
  printf("%-28s %8.2f warp-instr/clk/SM  (%6.2f clk per warp-instr per SMSP)  %s\n", name, 16.0 * iters * per_iter / h, double(h) * 4 / (16.0 * iters * per_iter),
         cudaGetErrorString(cudaGetLastError()));
}

int main() {
  long long* out; float* sink;
  cudaMalloc(&out, 8); cudaMalloc(&sink, 4);
  run<0>("ex2", 16, out, sink);
  run<1>("f2fp.pack", 16, out, sink);
  run<2>("ex2 + f2fp (1:1)", 32, out, sink);
  run<3>("fmnmx3", 16, out, sink);
  run<4>("ffma", 16, out, sink);
  run<5>("fadd", 16, out, sink);
  run<6>("ex2 + ffma (1:1)", 32, out, sink);
  run<7>("f2fp + ffma (1:1)", 32, out, sink);
  run<8>("f2fp + fmnmx3 (1:1)", 32, out, sink);
  run<9>("ffma + fadd (1:1)", 32, out, sink);
  return 0;
}
