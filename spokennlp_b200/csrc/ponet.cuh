// PoNet pooling mixer (reference call site alimeeting4mug/src/models/modeling_ponet.py:68-79, algorithm
// restated in oracle/ponet_oracle.py — the modelscope source is absent, parity is against that restatement).
//
// Input: the five projections of one layer packed token-major in ONE fp16 buffer proj[B*S, 5H] = [Q | K | O | Sg | Lc]
// (written by a single [5H,H] tcgen05 GEMM), the key-padding mask and the monotone segment ids.  HBM-bound streaming
// kernels, 128-bit accesses, every projection element read once (K twice, second time from L2):
//   1. ponet_qsum_kernel        qsum[b,:]  = sum over valid s of Q[b,s,:], cnt[b]            (global branch, step 1)
//   2. ponet_global_part_kernel per (b, 128-key chunk, head): online-softmax partials of qbar.K_s/8 over the chunk
//   3. ponet_global_comb_kernel g[b,:] = combine partials                                     (one vector per head)
//   4. ponet_segmax_kernel      segmax[b,seg,:] = max over the run of Sg (masked rows = -1e4); runs are contiguous, so
//                               a thread keeps a running max and touches memory once per run (atomic max)
//   5. ponet_mix_kernel         out = (g + segmax[seg_s]) * O_s + max3(Lc_{s-1}, Lc_s, Lc_{s+1}), zero on padding
#pragma once
#include "rowwise.cuh"

namespace b200 {

constexpr float PONET_NEG = -10000.0f;

__device__ __forceinline__ void atomic_max_f32(float* addr, float v) {
  if (v >= 0.f) atomicMax(reinterpret_cast<int*>(addr), __float_as_int(v));
  else atomicMin(reinterpret_cast<unsigned int*>(addr), __float_as_uint(v));
}

// grid (ceil(S/128), B), block H/8 threads (each owns 8 columns)
__global__ void ponet_qsum_kernel(const __half* __restrict__ proj, int ld, const float* __restrict__ key_bias, float* __restrict__ qsum,
                                  float* __restrict__ cnt, int S, int H) {
  const int b = blockIdx.y, s0 = blockIdx.x * 128, c = threadIdx.x * 8;
  if (c >= H) return;
  float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  int n = 0;
  for (int s = s0; s < min(S, s0 + 128); ++s) {
    if (key_bias && key_bias[static_cast<size_t>(b) * S + s] != 0.f) continue;
    const Vec8 v = load8(proj + (static_cast<size_t>(b) * S + s) * ld + c);
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] += v.v[j];
    ++n;
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) atomicAdd(qsum + static_cast<size_t>(b) * H + c + j, acc[j]);
  if (threadIdx.x == 0) atomicAdd(cnt + b, static_cast<float>(n));
}

// grid (ceil(S/128), heads, B), block 128: thread t <-> key s0+t.  part[b,h,chunk] = {m, l, acc[64]} (log2 domain).
__global__ void __launch_bounds__(128) ponet_global_part_kernel(const __half* __restrict__ proj, int ld, const float* __restrict__ key_bias,
                                                                const float* __restrict__ qsum, const float* __restrict__ cnt,
                                                                float* __restrict__ part, int S, int H, int heads) {
  __shared__ float qb[64];
  __shared__ float p[128];
  __shared__ float red[4];
  const int b = blockIdx.z, h = blockIdx.y, chunk = blockIdx.x, nchunks = gridDim.x;
  const int t = threadIdx.x, s = chunk * 128 + t;
  if (t < 64) qb[t] = qsum[static_cast<size_t>(b) * H + h * 64 + t] / fmaxf(cnt[b], 1.0f) * (0.125f * 1.4426950408889634f);
  __syncthreads();
  const bool valid = s < S && !(key_bias && key_bias[static_cast<size_t>(b) * S + s] != 0.f);
  float sc = -INFINITY;
  const __half* krow = proj + (static_cast<size_t>(b) * S + min(s, S - 1)) * ld + H + h * 64;
  if (valid) {
    sc = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const Vec8 kv = load8(krow + i * 8);
#pragma unroll
      for (int j = 0; j < 8; ++j) sc = fmaf(kv.v[j], qb[i * 8 + j], sc);
    }
  }
  float m = warp_max(sc);
  if ((t & 31) == 0) red[t >> 5] = m;
  __syncthreads();
  m = fmaxf(fmaxf(red[0], red[1]), fmaxf(red[2], red[3]));
  const float pe = (valid && m > -INFINITY) ? exp2f(sc - m) : 0.f;
  p[t] = pe;
  float l = warp_sum(pe);
  __syncthreads();
  if ((t & 31) == 0) red[t >> 5] = l;
  __syncthreads();
  l = red[0] + red[1] + red[2] + red[3];
  float* dst = part + ((static_cast<size_t>(b) * heads + h) * nchunks + chunk) * 66;
  if (t < 64) {
    float acc = 0.f;
    const int rows = min(128, S - chunk * 128);
    const __half* kc = proj + (static_cast<size_t>(b) * S + chunk * 128) * ld + H + h * 64 + t;
    for (int r = 0; r < rows; ++r) acc = fmaf(p[r], __half2float(kc[static_cast<size_t>(r) * ld]), acc);
    dst[2 + t] = acc;
  }
  if (t == 0) {
    dst[0] = m;
    dst[1] = l;
  }
}

// grid (heads, B), block 64
__global__ void ponet_global_comb_kernel(const float* __restrict__ part, float* __restrict__ g, int nchunks, int H, int heads) {
  const int b = blockIdx.y, h = blockIdx.x, d = threadIdx.x;
  const float* src = part + (static_cast<size_t>(b) * heads + h) * nchunks * 66;
  float m = -INFINITY;
  for (int c = 0; c < nchunks; ++c) m = fmaxf(m, src[c * 66]);
  float l = 0.f, acc = 0.f;
  for (int c = 0; c < nchunks; ++c) {
    const float w = src[c * 66] > -INFINITY ? exp2f(src[c * 66] - m) : 0.f;
    l = fmaf(w, src[c * 66 + 1], l);
    acc = fmaf(w, src[c * 66 + 2 + d], acc);
  }
  g[static_cast<size_t>(b) * H + h * 64 + d] = l > 0.f ? acc / l : 0.f;
}

// grid (ceil(S/64), B), block H/8.  segmax must be pre-filled with -inf.
__global__ void ponet_segmax_kernel(const __half* __restrict__ proj, int ld, const float* __restrict__ key_bias,
                                    const int64_t* __restrict__ seg, float* __restrict__ segmax, int S, int H, int nseg) {
  const int b = blockIdx.y, s0 = blockIdx.x * 64, c = threadIdx.x * 8;
  if (c >= H) return;
  float run[8];
  long cur = -1;
  for (int s = s0; s < min(S, s0 + 64); ++s) {
    const long id = seg[static_cast<size_t>(b) * S + s];
    const bool pad = key_bias && key_bias[static_cast<size_t>(b) * S + s] != 0.f;
    Vec8 v;
    if (pad) {
#pragma unroll
      for (int j = 0; j < 8; ++j) v.v[j] = PONET_NEG;
    } else {
      v = load8(proj + (static_cast<size_t>(b) * S + s) * ld + 3 * H + c);
    }
    if (id != cur) {
      if (cur >= 0 && cur < nseg)
#pragma unroll
        for (int j = 0; j < 8; ++j) atomic_max_f32(segmax + (static_cast<size_t>(b) * nseg + cur) * H + c + j, run[j]);
      cur = id;
#pragma unroll
      for (int j = 0; j < 8; ++j) run[j] = v.v[j];
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) run[j] = fmaxf(run[j], v.v[j]);
    }
  }
  if (cur >= 0 && cur < nseg)
#pragma unroll
    for (int j = 0; j < 8; ++j) atomic_max_f32(segmax + (static_cast<size_t>(b) * nseg + cur) * H + c + j, run[j]);
}

// grid ceil(B*S / 4), block 4 warps: warp per token row
__global__ void __launch_bounds__(ROW_WARPS * 32) ponet_mix_kernel(const __half* __restrict__ proj, int ld, const float* __restrict__ key_bias,
                                                                    const int64_t* __restrict__ seg, const float* __restrict__ g,
                                                                    const float* __restrict__ segmax, __half* __restrict__ out, int B, int S,
                                                                    int H, int nseg) {
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * ROW_WARPS + (threadIdx.x >> 5);
  if (row >= B * S) return;
  const int b = row / S, s = row % S;
  const bool pad = key_bias && key_bias[row] != 0.f;
  const int nv = lane_vecs(H, lane);
  const bool has_l = s > 0, has_r = s + 1 < S;
  const bool pad_l = has_l && key_bias && key_bias[row - 1] != 0.f, pad_r = has_r && key_bias && key_bias[row + 1] != 0.f;
  const long id = seg[row];
#pragma unroll
  for (int i = 0; i < ROW_MAXV; ++i)
    if (i < nv) {
      const int c = (i * 32 + lane) * 8;
      Vec8 o;
      if (pad) {
#pragma unroll
        for (int j = 0; j < 8; ++j) o.v[j] = 0.f;
      } else {
        const __half* base = proj + static_cast<size_t>(row) * ld;
        const Vec8 ov = load8(base + 2 * H + c);
        Vec8 loc = load8(base + 4 * H + c);
        if (has_l) {
          if (pad_l) {
#pragma unroll
            for (int j = 0; j < 8; ++j) loc.v[j] = fmaxf(loc.v[j], PONET_NEG);
          } else {
            const Vec8 t = load8(base - ld + 4 * H + c);
#pragma unroll
            for (int j = 0; j < 8; ++j) loc.v[j] = fmaxf(loc.v[j], t.v[j]);
          }
        }
        if (has_r) {
          if (pad_r) {
#pragma unroll
            for (int j = 0; j < 8; ++j) loc.v[j] = fmaxf(loc.v[j], PONET_NEG);
          } else {
            const Vec8 t = load8(base + ld + 4 * H + c);
#pragma unroll
            for (int j = 0; j < 8; ++j) loc.v[j] = fmaxf(loc.v[j], t.v[j]);
          }
        }
        const Vec8 gv = load8(g + static_cast<size_t>(b) * H + c);
        const Vec8 sm = load8(segmax + (static_cast<size_t>(b) * nseg + min(static_cast<long>(nseg - 1), max(0l, id))) * H + c);
#pragma unroll
        for (int j = 0; j < 8; ++j) o.v[j] = fmaf(gv.v[j] + sm.v[j], ov.v[j], loc.v[j]);
      }
      store8(out + static_cast<size_t>(row) * H + c, o);
    }
}

}  // namespace b200
