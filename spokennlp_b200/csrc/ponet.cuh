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
// Grids of the streaming kernels below.  At the BASELINE shape ([2, 4096]) the batch offers no parallelism, so the grids come
// from the sequence, and two things bound these kernels long before HBM does (profiles/r02_hbm_rooflines.md):
//   * DRAM round trips: a thread issues ALL loads of its positions before the first use (PONET_*_POS positions per batch);
//   * same-address atomics: a per-sequence sum that every block adds into serialises at L2 (~50 ns each: 512 blocks -> 25 us).
//     Blocks therefore carry PONET_GROUPS position groups, reduce them in shared memory, and add once per block.
constexpr int PONET_QSUM_POS = 16;
constexpr int PONET_RUN_POS = 16;
constexpr int PONET_GROUPS = 4;        // position groups per block of ponet_qsum_kernel / ponet_bwd_sums_kernel
constexpr int PONET_QSUM_ROUNDS = 2;   // batches of PONET_QSUM_POS positions per group: 128 positions per block
constexpr int PONET_MAX_H = 1024;      // static shared memory of the in-block reductions

__device__ __forceinline__ void atomic_max_f32(float* addr, float v) {
  if (v >= 0.f) atomicMax(reinterpret_cast<int*>(addr), __float_as_int(v));
  else atomicMin(reinterpret_cast<unsigned int*>(addr), __float_as_uint(v));
}

// Raw 8-half vectors: the streaming kernels below issue ALL loads of their positions before the first use (one DRAM round trip
// per block instead of one per position — at [2, 4096] these kernels are latency-bound, not bandwidth-bound) and convert late.
__device__ __forceinline__ uint4 ldraw8(const __half* p) { return *reinterpret_cast<const uint4*>(p); }
__device__ __forceinline__ Vec8 cvt8(const uint4& raw) {
  const __half2* h = reinterpret_cast<const __half2*>(&raw);
  Vec8 r;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 f = __half22float2(h[i]);
    r.v[2 * i] = f.x;
    r.v[2 * i + 1] = f.y;
  }
  return r;
}
constexpr uint32_t PONET_NEG_H2 = 0xF0E2F0E2u;    // two fp16 -10000 (exact)
constexpr uint32_t PONET_NINF_H2 = 0xFC00FC00u;   // two fp16 -inf

// grid (ceil(S / (PONET_GROUPS * PONET_QSUM_ROUNDS * PONET_QSUM_POS)), B), block (H/8, PONET_GROUPS): thread (x, y) owns 8
// columns of position group y
__global__ void ponet_qsum_kernel(const __half* __restrict__ proj, int ld, const float* __restrict__ key_bias, float* __restrict__ qsum,
                                  float* __restrict__ cnt, int S, int H) {
  __shared__ float red[PONET_GROUPS][PONET_MAX_H];
  __shared__ int nred[PONET_GROUPS];
  const int b = blockIdx.y, c = threadIdx.x * 8, y = threadIdx.y;
  float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  int n = 0;
#pragma unroll 1
  for (int rd = 0; rd < PONET_QSUM_ROUNDS; ++rd) {
    const int s0 = ((blockIdx.x * PONET_GROUPS + y) * PONET_QSUM_ROUNDS + rd) * PONET_QSUM_POS;
    if (s0 >= S) break;
    bool keep[PONET_QSUM_POS];
    uint4 raw[PONET_QSUM_POS];
#pragma unroll
    for (int k = 0; k < PONET_QSUM_POS; ++k) {
      const int s = s0 + k;
      keep[k] = s < S && !(key_bias && key_bias[static_cast<size_t>(b) * S + min(s, S - 1)] != 0.f);
    }
#pragma unroll
    for (int k = 0; k < PONET_QSUM_POS; ++k) raw[k] = ldraw8(proj + (static_cast<size_t>(b) * S + min(s0 + k, S - 1)) * ld + c);
#pragma unroll
    for (int k = 0; k < PONET_QSUM_POS; ++k)
      if (keep[k]) {
        const Vec8 v = cvt8(raw[k]);
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] += v.v[j];
        ++n;
      }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) red[y][c + j] = acc[j];
  if (threadIdx.x == 0) nred[y] = n;
  __syncthreads();
  if (y == 0) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float t = 0.f;
#pragma unroll
      for (int g = 0; g < PONET_GROUPS; ++g) t += red[g][c + j];
      atomicAdd(qsum + static_cast<size_t>(b) * H + c + j, t);
    }
    if (threadIdx.x == 0) {
      int t = 0;
#pragma unroll
      for (int g = 0; g < PONET_GROUPS; ++g) t += nred[g];
      atomicAdd(cnt + b, static_cast<float>(t));
    }
  }
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
#pragma unroll 8
    for (int r = 0; r < rows; ++r) acc = fmaf(p[r], __half2float(kc[static_cast<size_t>(r) * ld]), acc);
    dst[2 + t] = acc;
  }
  if (t == 0) {
    dst[0] = m;
    dst[1] = l;
  }
}

// grid (heads, B), block 64.  All partial statistics in one round trip (thread c fetches chunk c's), then the accumulators
// sixteen chunks at a time.
__global__ void ponet_global_comb_kernel(const float* __restrict__ part, float* __restrict__ g, int nchunks, int H, int heads) {
  __shared__ float wgt[64];
  __shared__ float red[2][2];
  const int b = blockIdx.y, h = blockIdx.x, d = threadIdx.x;
  const float* src = part + (static_cast<size_t>(b) * heads + h) * nchunks * 66;
  float l = 0.f, acc = 0.f, m = -INFINITY;
  for (int c0 = 0; c0 < nchunks; c0 += 64) {      // (one pass up to S = 8192)
    const int c = c0 + d;
    const float mc = c < nchunks ? src[c * 66] : -INFINITY, lc = c < nchunks ? src[c * 66 + 1] : 0.f;
    float mb = warp_max(mc);
    if ((d & 31) == 0) red[0][d >> 5] = mb;
    __syncthreads();
    mb = fmaxf(red[0][0], red[0][1]);
    const float m_new = fmaxf(m, mb);
    const float rescale = m > -INFINITY ? exp2f(m - m_new) : 0.f;
    const float w = mc > -INFINITY ? exp2f(mc - m_new) : 0.f;
    wgt[d] = w;
    float lb = warp_sum(w * lc);
    if ((d & 31) == 0) red[1][d >> 5] = lb;
    __syncthreads();
    l = l * rescale + red[1][0] + red[1][1];
    acc *= rescale;
    const int nc = min(64, nchunks - c0);
#pragma unroll 16
    for (int k = 0; k < nc; ++k) acc = fmaf(wgt[k], src[(c0 + k) * 66 + 2 + d], acc);
    m = m_new;
    __syncthreads();
  }
  g[static_cast<size_t>(b) * H + h * 64 + d] = l > 0.f ? acc / l : 0.f;
}

// grid (ceil(S/PONET_RUN_POS), B), block H/8.  segmax must be pre-filled with -inf.
__global__ void ponet_segmax_kernel(const __half* __restrict__ proj, int ld, const float* __restrict__ key_bias,
                                    const int64_t* __restrict__ seg, float* __restrict__ segmax, int S, int H, int nseg) {
  const int b = blockIdx.y, s0 = blockIdx.x * PONET_RUN_POS, c = threadIdx.x * 8;
  if (c >= H) return;
  const int np = min(PONET_RUN_POS, S - s0);
  long ids[PONET_RUN_POS];
  bool pad[PONET_RUN_POS];
  uint4 raw[PONET_RUN_POS];
#pragma unroll
  for (int k = 0; k < PONET_RUN_POS; ++k) {
    const size_t row = static_cast<size_t>(b) * S + min(s0 + k, S - 1);
    ids[k] = seg[row];
    pad[k] = key_bias && key_bias[row] != 0.f;
  }
#pragma unroll
  for (int k = 0; k < PONET_RUN_POS; ++k) raw[k] = ldraw8(proj + (static_cast<size_t>(b) * S + min(s0 + k, S - 1)) * ld + 3 * H + c);
  float run[8];
  long cur = -1;
#pragma unroll
  for (int k = 0; k < PONET_RUN_POS; ++k) {
    if (k >= np) break;
    if (pad[k]) raw[k] = make_uint4(PONET_NEG_H2, PONET_NEG_H2, PONET_NEG_H2, PONET_NEG_H2);
    const Vec8 v = cvt8(raw[k]);
    if (ids[k] != cur) {
      if (cur >= 0 && cur < nseg)
#pragma unroll
        for (int j = 0; j < 8; ++j) atomic_max_f32(segmax + (static_cast<size_t>(b) * nseg + cur) * H + c + j, run[j]);
      cur = ids[k];
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

// ------------------------------------------------------------------------------------------------ backward
// out_s = (g + seg_s) o O_s + loc_s on kept rows.  With t_s = dout_s o O_s:
//   dO_s = dout_s o (g + seg_s);   dg = sum_s t_s;   dSg_u = [Sg_u == segmax(seg_u)] * sum_{s in seg_u} t_s;
//   dLc_u = sum over the (<=3) windows whose arg-max is u of dout_s;
//   global branch (g = sum_s a_s K_s, a = softmax(qbar.K/8)):  dz_s = a_s (dg.K_s - dg.g),
//   dK_s = a_s dg + dz_s qbar / 8,  dqbar = sum_s dz_s K_s / 8,  dQ_s = dqbar / cnt on kept rows.
// Ties of the segment max are shared evenly (see pass 1); ties inside a local window go to the first row (max_pool1d).

// pass 1: dg[b,:], segsum[b,seg,:] and segties[b,seg,:] (running sums over the contiguous runs).
// grid (ceil(S / (PONET_GROUPS * PONET_RUN_POS)), B), block (H/8, PONET_GROUPS): thread (x, y) owns 8 columns of position group y
// segties counts the rows of a segment that attain its maximum: the fp16 grid makes exact ties real, and amax's
// gradient is shared evenly between tied rows (torch scatter_reduce "amax" semantics, which the restatement follows).
__global__ void __launch_bounds__(PONET_MAX_H / 8 * PONET_GROUPS)
ponet_bwd_sums_kernel(const __half* __restrict__ proj, int ld, const __half* __restrict__ dout, const float* __restrict__ key_bias,
                      const int64_t* __restrict__ seg, const float* __restrict__ segmax, float* __restrict__ dg,
                      float* __restrict__ segsum, float* __restrict__ segties, int S, int H, int nseg) {
  __shared__ float red[PONET_GROUPS][PONET_MAX_H];
  const int b = blockIdx.y, s0 = (blockIdx.x * PONET_GROUPS + threadIdx.y) * PONET_RUN_POS, c = threadIdx.x * 8;
  float run[8], tot[8], ties[8];
  Vec8 mx;
#pragma unroll
  for (int j = 0; j < 8; ++j) run[j] = tot[j] = ties[j] = mx.v[j] = 0.f;
  long cur = -1;
  auto flush = [&]() {
    if (cur >= 0 && cur < nseg) {
      const size_t o = (static_cast<size_t>(b) * nseg + cur) * H + c;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        if (run[j] != 0.f) atomicAdd(segsum + o + j, run[j]);
        if (ties[j] != 0.f) atomicAdd(segties + o + j, ties[j]);
      }
    }
  };
  constexpr int NB = 4;                          // positions per batch: 3 x 4 raw vectors in flight per thread
  static_assert(PONET_RUN_POS % NB == 0, "batches tile the block");
#pragma unroll 1
  for (int sb = s0; sb < min(S, s0 + PONET_RUN_POS); sb += NB) {
    long ids[NB];
    bool pad[NB];
    uint4 rd[NB], ro[NB], rs[NB];
#pragma unroll
    for (int k = 0; k < NB; ++k) {
      const size_t row = static_cast<size_t>(b) * S + min(sb + k, S - 1);
      ids[k] = seg[row];
      pad[k] = key_bias && key_bias[row] != 0.f;
    }
#pragma unroll
    for (int k = 0; k < NB; ++k) {
      const size_t row = static_cast<size_t>(b) * S + min(sb + k, S - 1);
      rd[k] = ldraw8(dout + row * H + c);
      ro[k] = ldraw8(proj + row * ld + 2 * H + c);
      rs[k] = ldraw8(proj + row * ld + 3 * H + c);
    }
#pragma unroll
    for (int k = 0; k < NB; ++k) {
      if (sb + k >= S) break;
      if (ids[k] != cur) {
        flush();
        cur = ids[k];
#pragma unroll
        for (int j = 0; j < 8; ++j) run[j] = ties[j] = 0.f;
        if (cur >= 0 && cur < nseg) mx = load8(segmax + (static_cast<size_t>(b) * nseg + cur) * H + c);
      }
      if (pad[k]) continue;
      const Vec8 d = cvt8(rd[k]), o = cvt8(ro[k]), sg = cvt8(rs[k]);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float t = d.v[j] * o.v[j];
        run[j] += t;
        tot[j] += t;
        ties[j] += sg.v[j] == mx.v[j] ? 1.f : 0.f;
      }
    }
  }
  flush();
#pragma unroll
  for (int j = 0; j < 8; ++j) red[threadIdx.y][c + j] = tot[j];
  __syncthreads();
  if (threadIdx.y == 0) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float t = 0.f;
#pragma unroll
      for (int g = 0; g < PONET_GROUPS; ++g) t += red[g][c + j];
      atomicAdd(dg + static_cast<size_t>(b) * H + c + j, t);
    }
  }
}

// forward helper for the backward: lse[b,h] (log2 domain) of the global softmax.  grid (heads, B), block 32
__global__ void ponet_global_lse_kernel(const float* __restrict__ part, float* __restrict__ lse, int nchunks, int heads) {
  const int b = blockIdx.y, h = blockIdx.x;
  const float* src = part + (static_cast<size_t>(b) * heads + h) * nchunks * 66;
  float m = -INFINITY;
  for (int c = threadIdx.x; c < nchunks; c += 32) m = fmaxf(m, src[c * 66]);
  m = warp_max(m);
  float l = 0.f;
  for (int c = threadIdx.x; c < nchunks; c += 32) l += src[c * 66] > -INFINITY ? exp2f(src[c * 66] - m) * src[c * 66 + 1] : 0.f;
  l = warp_sum(l);
  if (threadIdx.x == 0) lse[b * heads + h] = l > 0.f ? m + log2f(l) : INFINITY;
}

// pass 2: global branch.  grid (ceil(S/128), heads, B), block 128: thread t <-> key.  Writes dK rows, accumulates dqbar.
__global__ void __launch_bounds__(128) ponet_bwd_global_kernel(const __half* __restrict__ proj, int ld, const float* __restrict__ key_bias,
                                                               const float* __restrict__ qsum, const float* __restrict__ cnt,
                                                               const float* __restrict__ g, const float* __restrict__ lse,
                                                               const float* __restrict__ dg, float* __restrict__ dqbar,
                                                               __half* __restrict__ dproj, int ld_d, int S, int H, int heads) {
  __shared__ float qb[64], dgs[64], acc[4][64];
  __shared__ float c0s;
  const int b = blockIdx.z, h = blockIdx.y, t = threadIdx.x, s = blockIdx.x * 128 + t;
  if (t < 64) {
    qb[t] = qsum[static_cast<size_t>(b) * H + h * 64 + t] / fmaxf(cnt[b], 1.0f);
    dgs[t] = dg[static_cast<size_t>(b) * H + h * 64 + t];
  }
  __syncthreads();
  if (t < 32) {
    float v = dgs[t] * g[static_cast<size_t>(b) * H + h * 64 + t] + dgs[t + 32] * g[static_cast<size_t>(b) * H + h * 64 + t + 32];
    v = warp_sum(v);
    if (t == 0) c0s = v;
  }
  __syncthreads();
  const bool valid = s < S && !(key_bias && key_bias[static_cast<size_t>(b) * S + s] != 0.f);
  float dz = 0.f, a = 0.f;
  Vec8 kv[8];
  if (s < S) {
    const __half* krow = proj + (static_cast<size_t>(b) * S + s) * ld + H + h * 64;
#pragma unroll
    for (int i = 0; i < 8; ++i) kv[i] = load8(krow + i * 8);
  }
  if (valid) {
    float z = 0.f, dd = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        z = fmaf(kv[i].v[j], qb[i * 8 + j], z);
        dd = fmaf(kv[i].v[j], dgs[i * 8 + j], dd);
      }
    a = exp2f(z * (0.125f * 1.4426950408889634f) - lse[b * heads + h]);
    dz = a * (dd - c0s);
  }
  if (s < S) {
    __half* drow = dproj + (static_cast<size_t>(b) * S + s) * ld_d + H + h * 64;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      Vec8 o;
#pragma unroll
      for (int j = 0; j < 8; ++j) o.v[j] = valid ? fmaf(dz * 0.125f, qb[i * 8 + j], a * dgs[i * 8 + j]) : 0.f;
      store8(drow + i * 8, o);
    }
  }
  // dqbar[d] += sum_s dz_s K_s[d] / 8 : transposing butterfly over the 32 keys of a warp — at offset o a lane keeps one half
  // of its values and trades the other (62 shuffles instead of 64 x 5); lane l ends with columns 2l, 2l+1 — then across warps
  {
    float v[64];
    const float wz = valid ? dz * 0.125f : 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int j = 0; j < 8; ++j) v[i * 8 + j] = wz * kv[i].v[j];
#pragma unroll
    for (int o = 16, n = 32; o >= 1; o >>= 1, n >>= 1) {
      const bool upper = (t & o) != 0;
#pragma unroll
      for (int k = 0; k < n; ++k) {
        const float send = upper ? v[k] : v[k + n];
        const float keep = upper ? v[k + n] : v[k];
        v[k] = keep + __shfl_xor_sync(0xffffffffu, send, o);
      }
    }
    acc[t >> 5][2 * (t & 31)] = v[0];
    acc[t >> 5][2 * (t & 31) + 1] = v[1];
  }
  __syncthreads();
  if (t < 64) atomicAdd(dqbar + static_cast<size_t>(b) * H + h * 64 + t, acc[0][t] + acc[1][t] + acc[2][t] + acc[3][t]);
}

// pass 3: everything row-local; writes dQ, dO, dSg, dLc (dK was written by pass 2).  grid (ceil(S/PONET_ROWS_POS), B), block H/8:
// a thread owns 8 columns and walks PONET_ROWS_POS consecutive positions with the local window (Lc of s-2..s+2, dout of s-1..s+1)
// held in registers — every projection element is fetched once (+ a halo of 4 / 2 rows per block) instead of up to five times by
// five different warps, and all loads of the block are issued before the first use.  Was warp-per-row: 57 us at [2, 4096].
constexpr int PONET_ROWS_POS = 8;
__global__ void __launch_bounds__(128) ponet_bwd_rows_kernel(const __half* __restrict__ proj, int ld, const __half* __restrict__ dout,
                                                             const float* __restrict__ key_bias, const int64_t* __restrict__ seg,
                                                             const float* __restrict__ g, const float* __restrict__ segmax,
                                                             const float* __restrict__ segsum, const float* __restrict__ segties,
                                                             const float* __restrict__ dqbar, const float* __restrict__ cnt,
                                                             __half* __restrict__ dproj, int ld_d, int B, int S, int H, int nseg) {
  constexpr int P = PONET_ROWS_POS;
  const int b = blockIdx.y, s0 = blockIdx.x * P, c = threadIdx.x * 8;
  if (c >= H) return;
  auto padded = [&](int ss) { return key_bias && key_bias[static_cast<size_t>(b) * S + ss] != 0.f; };
  // flags of positions s0-2 .. s0+P+1: 0 = in range and kept, 1 = padded, 2 = out of range
  int flag[P + 4];
  long ids[P];
#pragma unroll
  for (int k = 0; k < P + 4; ++k) {
    const int ss = s0 - 2 + k;
    flag[k] = (ss < 0 || ss >= S) ? 2 : (padded(ss) ? 1 : 0);
  }
#pragma unroll
  for (int k = 0; k < P; ++k) ids[k] = min(static_cast<long>(nseg - 1), max(0l, static_cast<long>(seg[static_cast<size_t>(b) * S + min(s0 + k, S - 1)])));
  uint4 lc[P + 4], dd[P + 2], sg[P];
#pragma unroll
  for (int k = 0; k < P + 4; ++k) {
    const int ss = min(max(s0 - 2 + k, 0), S - 1);
    lc[k] = ldraw8(proj + (static_cast<size_t>(b) * S + ss) * ld + 4 * H + c);
  }
#pragma unroll
  for (int k = 0; k < P + 2; ++k) {
    const int ss = min(max(s0 - 1 + k, 0), S - 1);
    dd[k] = ldraw8(dout + (static_cast<size_t>(b) * S + ss) * H + c);
  }
#pragma unroll
  for (int k = 0; k < P; ++k) sg[k] = ldraw8(proj + (static_cast<size_t>(b) * S + min(s0 + k, S - 1)) * ld + 3 * H + c);
  const Vec8 gv = load8(g + static_cast<size_t>(b) * H + c), dqb = load8(dqbar + static_cast<size_t>(b) * H + c);
  const float inv_cnt = 1.0f / fmaxf(cnt[b], 1.0f);
  // masked local-branch values: padded -> -1e4, out of range -> -inf (never wins a window); dropped rows carry no dout
#pragma unroll
  for (int k = 0; k < P + 4; ++k) {
    if (flag[k] == 1) lc[k] = make_uint4(PONET_NEG_H2, PONET_NEG_H2, PONET_NEG_H2, PONET_NEG_H2);
    if (flag[k] == 2) lc[k] = make_uint4(PONET_NINF_H2, PONET_NINF_H2, PONET_NINF_H2, PONET_NINF_H2);
  }
#pragma unroll
  for (int k = 0; k < P + 2; ++k)
    if (flag[k + 1] != 0) dd[k] = make_uint4(0, 0, 0, 0);
  Vec8 sm, ss, st;
  long cur = -1;
#pragma unroll
  for (int k = 0; k < P; ++k) {
    const int s = s0 + k;
    if (s >= S) break;
    __half* drow = dproj + (static_cast<size_t>(b) * S + s) * ld_d;
    Vec8 dq, dO, dsg, dlc;
    if (flag[k + 2] == 1) {
#pragma unroll
      for (int j = 0; j < 8; ++j) dq.v[j] = dO.v[j] = dsg.v[j] = dlc.v[j] = 0.f;
    } else {
      if (ids[k] != cur) {                        // the segment tables change once per sentence
        cur = ids[k];
        const size_t o = (static_cast<size_t>(b) * nseg + cur) * H + c;
        sm = load8(segmax + o);
        ss = load8(segsum + o);
        st = load8(segties + o);
      }
      const Vec8 d = cvt8(dd[k + 1]), sgv = cvt8(sg[k]);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        dq.v[j] = dqb.v[j] * inv_cnt;
        dO.v[j] = d.v[j] * (gv.v[j] + sm.v[j]);
        dsg.v[j] = (sgv.v[j] == sm.v[j]) ? ss.v[j] / fmaxf(st.v[j], 1.f) : 0.f;
      }
      // dLc_u: u = s is the arg-max of window w (centred at w in {s-1, s, s+1}) iff Lc_s beats the other two taps
      const Vec8 l0 = cvt8(lc[k]), l1 = cvt8(lc[k + 1]), l2 = cvt8(lc[k + 2]), l3 = cvt8(lc[k + 3]), l4 = cvt8(lc[k + 4]);
      const Vec8 dm = cvt8(dd[k]), dp = cvt8(dd[k + 2]);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float me = l2.v[j];
        float acc = 0.f;
        // first maximum wins inside a window (taps ordered left to right), like an arg-max scan
        if (me > l0.v[j] && me > l1.v[j]) acc += dm.v[j];                    // window centred at s-1: taps (s-2, s-1, s)
        if (me >= l3.v[j] && me > l1.v[j]) acc += d.v[j];                    // window centred at s  : taps (s-1, s, s+1)
        if (me >= l3.v[j] && me >= l4.v[j]) acc += dp.v[j];                  // window centred at s+1: taps (s, s+1, s+2)
        dlc.v[j] = acc;
      }
    }
    store8(drow + c, dq);
    store8(drow + 2 * H + c, dO);
    store8(drow + 3 * H + c, dsg);
    store8(drow + 4 * H + c, dlc);
  }
}

}  // namespace b200
