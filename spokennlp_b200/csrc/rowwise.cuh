// HBM-bound row-wise kernels: warp-per-row, 128-bit vectorised, warp-shuffle
// reductions, fp32 statistics.  Rows hold H <= 1024 elements (H % 8 == 0) and stay in registers between the
// statistics pass and the normalisation pass, so every activation byte is read exactly once.
#pragma once
#include "ptx.cuh"

namespace b200 {

constexpr int ROW_MAXV = 4;          // 4 x (32 lanes x 8 elements) = 1024 columns max
constexpr int ROW_WARPS = 4;

struct Vec8 {
  float v[8];
};

__device__ __forceinline__ Vec8 load8(const __half* p) {
  const uint4 raw = *reinterpret_cast<const uint4*>(p);
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
__device__ __forceinline__ Vec8 load8(const float* p) {
  const float4 a = *reinterpret_cast<const float4*>(p);
  const float4 b = *reinterpret_cast<const float4*>(p + 4);
  Vec8 r;
  r.v[0] = a.x; r.v[1] = a.y; r.v[2] = a.z; r.v[3] = a.w;
  r.v[4] = b.x; r.v[5] = b.y; r.v[6] = b.z; r.v[7] = b.w;
  return r;
}
__device__ __forceinline__ void store8(__half* p, const Vec8& r) {
  uint4 raw;
  __half2* h = reinterpret_cast<__half2*>(&raw);
#pragma unroll
  for (int i = 0; i < 4; ++i) h[i] = __floats2half2_rn(r.v[2 * i], r.v[2 * i + 1]);
  *reinterpret_cast<uint4*>(p) = raw;
}
__device__ __forceinline__ void store8(float* p, const Vec8& r) {
  *reinterpret_cast<float4*>(p) = make_float4(r.v[0], r.v[1], r.v[2], r.v[3]);
  *reinterpret_cast<float4*>(p + 4) = make_float4(r.v[4], r.v[5], r.v[6], r.v[7]);
}

// multiply the 8 consecutive elements starting at (even) linear index e0 by their dropout multipliers
__device__ __forceinline__ void drop8(Vec8& v, uint32_t e0, uint32_t seed, const DropCfg& d) {
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float m0, m1;
    drop_pair((e0 >> 1) + i, seed, d.thr15, d.scale, m0, m1);
    v.v[2 * i] *= m0;
    v.v[2 * i + 1] *= m1;
  }
}

// Row statistics over values already in registers.  Two-pass (mean, then centred variance) like the reference.
__device__ __forceinline__ void row_stats(const Vec8 (&x)[ROW_MAXV], int nv_lane, int H, float eps, float& mean, float& rstd) {
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < ROW_MAXV; ++i)
    if (i < nv_lane)
#pragma unroll
      for (int j = 0; j < 8; ++j) s += x[i].v[j];
  mean = warp_sum(s) / H;
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < ROW_MAXV; ++i)
    if (i < nv_lane)
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float d = x[i].v[j] - mean;
        q = fmaf(d, d, q);
      }
  rstd = rsqrtf(warp_sum(q) / H + eps);
}

// number of 8-wide vectors this lane owns: vector index = i*32 + lane, valid while (i*32+lane)*8 < H
__device__ __forceinline__ int lane_vecs(int H, int lane) {
  const int nvec = H >> 3;
  return (nvec - lane + 31) / 32 > 0 ? (nvec - lane + 31) / 32 : 0;
}

// ------------------------------------------------------------------------------------------------ K7: LayerNorm fwd
// y = (x - mean) * rstd * gamma + beta.  x: fp32 or fp16 pre-LN sum; y: fp16 (+ optional fp32 copy).
template <typename InT>
__global__ void __launch_bounds__(ROW_WARPS * 32) ln_fwd_kernel(const InT* __restrict__ x, const float* __restrict__ gamma,
                                                                 const float* __restrict__ beta, __half* __restrict__ y,
                                                                 float* __restrict__ y32, float* __restrict__ mean_out,
                                                                 float* __restrict__ rstd_out, int rows, int H, float eps) {
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * ROW_WARPS + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int nv = lane_vecs(H, lane);
  Vec8 v[ROW_MAXV];
  const InT* xr = x + static_cast<size_t>(row) * H;
#pragma unroll
  for (int i = 0; i < ROW_MAXV; ++i)
    if (i < nv) v[i] = load8(xr + (i * 32 + lane) * 8);
  float mean, rstd;
  row_stats(v, nv, H, eps, mean, rstd);
  if (lane == 0) {
    if (mean_out) mean_out[row] = mean;
    if (rstd_out) rstd_out[row] = rstd;
  }
#pragma unroll
  for (int i = 0; i < ROW_MAXV; ++i)
    if (i < nv) {
      const int c = (i * 32 + lane) * 8;
      const Vec8 g = load8(gamma + c), b = load8(beta + c);
      Vec8 o;
#pragma unroll
      for (int j = 0; j < 8; ++j) o.v[j] = fmaf((v[i].v[j] - mean) * rstd, g.v[j], b.v[j]);
      store8(y + static_cast<size_t>(row) * H + c, o);
      if (y32) store8(y32 + static_cast<size_t>(row) * H + c, o);
    }
}

// ------------------------------------------------------------------------------------------------ K10: LayerNorm bwd
// dx = rstd * (g - mean(g) - xhat * mean(g * xhat)),  g = dy * gamma
// dgamma += alpha * sum_rows dy * xhat;  dbeta += alpha * sum_rows dy;  dbias += alpha * sum_rows dx (optional: the
// bias of the dense layer that produced the pre-LN sum).  `dy2` (optional) is a second upstream gradient added to dy
// (the residual branch that by-passes the next block).

// The first generation kept a whole row per warp, which cost 72 accumulator registers per lane for the three column sums
// (dgamma, dbeta, dbias): 168 registers, 12 resident warps per SM, one row in flight per warp — 2.4 TB/s (profiles/r01d; removed in
// round 2, git 01a2ca4 keeps it).  Here a row is spread over `tpr` threads (one 8-column chunk each, so the column sums
// are 24 registers), a CTA runs lnb_threads / tpr independent row slots, and every slot pulls LNB_R rows per trip, so an SM keeps
// 24 warps x LNB_R rows of loads in flight.  The per-row sums cross the slot's warps through shared memory behind a
// slot-local named barrier (parity-double-buffered: one barrier per trip).
// CTA size: 768 threads per SM either way, but four CTAs of 192 threads (two row slots each at H = 768) interleave their load
// and compute phases better than two of 384 (30.7 -> 28.3 us with dropout at the bench shape; eight of 96 are worse: 33.5 us).
constexpr int lnb_threads(int wps) { return wps == 4 ? 256 : 192; }      // a multiple of tpr = 32 wps
constexpr int lnb_ctas(int wps) { return 768 / lnb_threads(wps); }
constexpr int LNB_R = 2;
static_assert(LNB_R == 2, "the row-sum exchange below moves one float4 (2 rows x 2 sums) per warp");

__device__ __forceinline__ Vec8 unpack8(const uint4& raw) {
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

// WPS = warps per row slot (tpr = 32 * WPS threads share a row)
template <typename InT, int WPS>
__global__ void __launch_bounds__(lnb_threads(WPS), lnb_ctas(WPS)) ln_bwd2_kernel(const __half* __restrict__ dy, const __half* __restrict__ dy2,
                                                                  const InT* __restrict__ x, const float* __restrict__ mean_in,
                                                                  const float* __restrict__ rstd_in, const float* __restrict__ gamma,
                                                                  __half* __restrict__ dx, float* __restrict__ dgamma,
                                                                  float* __restrict__ dbeta, float* __restrict__ dbias,
                                                                  const float* __restrict__ alpha_ptr, int rows, int H,
                                                                  __half* __restrict__ dx_drop, DropCfg drop) {
  extern __shared__ float red[];   // [3][slots][H] column sums, then [2 parities][slots][WPS] float4 row-sum partials
  constexpr int tpr = 32 * WPS, slots = lnb_threads(WPS) / tpr;
  const int slot = threadIdx.x / tpr, tl = threadIdx.x - slot * tpr;
  const int lane = threadIdx.x & 31, wis = tl >> 5;            // warp inside the slot
  const int c = tl * 8;
  const bool active = c < H;
  float4* part = reinterpret_cast<float4*>(red + 3 * slots * H);
  const bool dropping = dx_drop != nullptr && drop.seed_base != nullptr;
  const uint32_t dseed = dropping ? drop_seed(drop) : 0u;
  const float inv_h = 1.0f / H;
  Vec8 ag, ab, ad;
#pragma unroll
  for (int j = 0; j < 8; ++j) ag.v[j] = ab.v[j] = ad.v[j] = 0.f;
  int par = 0;
  for (int base = (blockIdx.x * slots + slot) * LNB_R; base < rows; base += gridDim.x * slots * LNB_R, par ^= 1) {
    // ---- every load of the trip is issued before anything is consumed (LNB_R rows x 48..64 bytes per thread in flight)
    uint4 xr0[LNB_R], xr1[LNB_R], dr[LNB_R], d2r[LNB_R];
    float mean[LNB_R], rs[LNB_R];
    bool ok[LNB_R];
#pragma unroll
    for (int k = 0; k < LNB_R; ++k) {
      const int row = base + k;
      ok[k] = row < rows && active;
      xr0[k] = xr1[k] = dr[k] = d2r[k] = make_uint4(0, 0, 0, 0);
      mean[k] = rs[k] = 0.f;
      if (ok[k]) {
        const size_t off = static_cast<size_t>(row) * H + c;
        xr0[k] = *reinterpret_cast<const uint4*>(x + off);
        if (sizeof(InT) == 4) xr1[k] = *reinterpret_cast<const uint4*>(reinterpret_cast<const char*>(x + off) + 16);
        dr[k] = *reinterpret_cast<const uint4*>(dy + off);
        if (dy2) d2r[k] = *reinterpret_cast<const uint4*>(dy2 + off);
        mean[k] = mean_in[row];
        rs[k] = rstd_in[row];
      }
    }
    Vec8 gm;
#pragma unroll
    for (int j = 0; j < 8; ++j) gm.v[j] = 0.f;
    if (active) gm = load8(gamma + c);
    Vec8 xh[LNB_R], d[LNB_R];
    float s[2 * LNB_R];
#pragma unroll
    for (int k = 0; k < LNB_R; ++k) {
      Vec8 xv;
      if (sizeof(InT) == 4) {
        const float4 a4 = *reinterpret_cast<const float4*>(&xr0[k]), b4 = *reinterpret_cast<const float4*>(&xr1[k]);
        xv.v[0] = a4.x; xv.v[1] = a4.y; xv.v[2] = a4.z; xv.v[3] = a4.w;
        xv.v[4] = b4.x; xv.v[5] = b4.y; xv.v[6] = b4.z; xv.v[7] = b4.w;
      } else {
        xv = unpack8(xr0[k]);
      }
      d[k] = unpack8(dr[k]);
      if (dy2) {
        const Vec8 d2 = unpack8(d2r[k]);
#pragma unroll
        for (int j = 0; j < 8; ++j) d[k].v[j] += d2.v[j];
      }
      s[2 * k] = s[2 * k + 1] = 0.f;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        xh[k].v[j] = ok[k] ? (xv.v[j] - mean[k]) * rs[k] : 0.f;
        const float g = d[k].v[j] * gm.v[j];
        s[2 * k] += g;
        s[2 * k + 1] = fmaf(g, xh[k].v[j], s[2 * k + 1]);
        ag.v[j] = fmaf(d[k].v[j], xh[k].v[j], ag.v[j]);
        ab.v[j] += d[k].v[j];
      }
    }
#pragma unroll
    for (int k = 0; k < 2 * LNB_R; ++k) s[k] = warp_sum(s[k]);
    if (WPS > 1) {
      float4* mine = part + (par * slots + slot) * WPS;
      if (lane == 0) mine[wis] = make_float4(s[0], s[1], s[2], s[3]);
      asm volatile("bar.sync %0, %1;" ::"r"(1 + slot), "n"(tpr) : "memory");
      float4 t = mine[0];
#pragma unroll
      for (int w = 1; w < WPS; ++w) {
        const float4 u = mine[w];
        t.x += u.x; t.y += u.y; t.z += u.z; t.w += u.w;
      }
      s[0] = t.x; s[1] = t.y; s[2] = t.z; s[3] = t.w;
    }
#pragma unroll
    for (int k = 0; k < LNB_R; ++k) {
      if (ok[k]) {
        const int row = base + k;
        const float m1 = s[2 * k] * inv_h, m2 = s[2 * k + 1] * inv_h;
        Vec8 o;
#pragma unroll
        for (int j = 0; j < 8; ++j) o.v[j] = rs[k] * (d[k].v[j] * gm.v[j] - m1 - xh[k].v[j] * m2);
        store8(dx + static_cast<size_t>(row) * H + c, o);
        if (dropping) {
          drop8(o, static_cast<uint32_t>(row) * static_cast<uint32_t>(H) + c, dseed, drop);
          store8(dx_drop + static_cast<size_t>(row) * H + c, o);
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) ad.v[j] += o.v[j];
      }
    }
  }
  // column sums: slots -> one atomic per column per CTA
  if (active) {
    store8(red + (0 * slots + slot) * H + c, ag);
    store8(red + (1 * slots + slot) * H + c, ab);
    store8(red + (2 * slots + slot) * H + c, ad);
  }
  __syncthreads();
  const float alpha = alpha_ptr ? *alpha_ptr : 1.0f;
  for (int col = threadIdx.x; col < H; col += blockDim.x) {
    float a = 0.f, b = 0.f, dd = 0.f;
#pragma unroll
    for (int w = 0; w < slots; ++w) {
      a += red[(0 * slots + w) * H + col];
      b += red[(1 * slots + w) * H + col];
      dd += red[(2 * slots + w) * H + col];
    }
    atomicAdd(dgamma + col, a * alpha);
    atomicAdd(dbeta + col, b * alpha);
    if (dbias) atomicAdd(dbias + col, dd * alpha);
  }
}

// ------------------------------------------------------------------------------------------------ K1: embeddings
// y = LN(word[ids] + pos[pos_ids] + type[tt]); tables are the fp32 master parameters (gathered rows, 128-bit loads).
// bert_model.py:184-210.  `inputs_embeds` (fp32 [rows,H]) replaces the word gather when given.
__global__ void __launch_bounds__(ROW_WARPS * 32) embed_ln_fwd_kernel(const int64_t* __restrict__ ids, const int64_t* __restrict__ tt,
                                                                       const int64_t* __restrict__ pos, const float* __restrict__ inputs_embeds,
                                                                       const float* __restrict__ word, const float* __restrict__ pos_tab,
                                                                       const float* __restrict__ type_tab, const float* __restrict__ gamma,
                                                                       const float* __restrict__ beta, __half* __restrict__ y,
                                                                       float* __restrict__ y32, int rows, int S, int H, float eps,
                                                                       DropCfg drop) {
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * ROW_WARPS + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int nv = lane_vecs(H, lane);
  const uint32_t dseed = drop.seed_base ? drop_seed(drop) : 0u;
  const float* w = inputs_embeds ? inputs_embeds + static_cast<size_t>(row) * H : word + static_cast<size_t>(ids[row]) * H;
  const float* p = pos_tab + static_cast<size_t>(pos ? pos[row] : (row % S)) * H;
  const float* t = type_tab + static_cast<size_t>(tt ? tt[row] : 0) * H;
  Vec8 v[ROW_MAXV];
#pragma unroll
  for (int i = 0; i < ROW_MAXV; ++i)
    if (i < nv) {
      const int c = (i * 32 + lane) * 8;
      const Vec8 a = load8(w + c), b = load8(p + c), d = load8(t + c);
#pragma unroll
      for (int j = 0; j < 8; ++j) v[i].v[j] = a.v[j] + d.v[j] + b.v[j];
    }
  float mean, rstd;
  row_stats(v, nv, H, eps, mean, rstd);
#pragma unroll
  for (int i = 0; i < ROW_MAXV; ++i)
    if (i < nv) {
      const int c = (i * 32 + lane) * 8;
      const Vec8 g = load8(gamma + c), b = load8(beta + c);
      Vec8 o;
#pragma unroll
      for (int j = 0; j < 8; ++j) o.v[j] = fmaf((v[i].v[j] - mean) * rstd, g.v[j], b.v[j]);
      if (drop.seed_base) drop8(o, static_cast<uint32_t>(row) * static_cast<uint32_t>(H) + c, dseed, drop);     // bert_model.py:209
      store8(y + static_cast<size_t>(row) * H + c, o);
      if (y32) store8(y32 + static_cast<size_t>(row) * H + c, o);
    }
}

// Backward of the embedding block: recompute the pre-LN sum and its statistics (cheaper than saving them), LN bwd,
// then scatter-add dx into the three fp32 gradient tables.  Column-sum-like targets (dgamma, dbeta, the type-0 row
// that almost every token hits) are accumulated in registers over a grid-stride loop and reduced once per block;
// word / position rows receive vectorised red.global.add.v4.f32.
__global__ void __launch_bounds__(ROW_WARPS * 32) embed_ln_bwd_kernel(const __half* __restrict__ dy, const __half* __restrict__ dy2,
                                                                       const int64_t* __restrict__ ids, const int64_t* __restrict__ tt,
                                                                       const int64_t* __restrict__ pos, const float* __restrict__ word,
                                                                       const float* __restrict__ pos_tab, const float* __restrict__ type_tab,
                                                                       const float* __restrict__ gamma, float* __restrict__ dword,
                                                                       float* __restrict__ dpos, float* __restrict__ dtype_tab,
                                                                       float* __restrict__ dgamma, float* __restrict__ dbeta,
                                                                       const float* __restrict__ alpha_ptr, int rows, int S, int H, float eps,
                                                                       DropCfg drop, long long pad_id, const float* __restrict__ inputs_embeds,
                                                                       float* __restrict__ d_inputs_embeds) {
  extern __shared__ float red[];   // [3][ROW_WARPS][H]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const float alpha = alpha_ptr ? *alpha_ptr : 1.0f;
  const uint32_t dseed = drop.seed_base ? drop_seed(drop) : 0u;
  const int nv = lane_vecs(H, lane);
  Vec8 ag[ROW_MAXV], ab[ROW_MAXV], at0[ROW_MAXV], gm[ROW_MAXV];
#pragma unroll
  for (int i = 0; i < ROW_MAXV; ++i) {
#pragma unroll
    for (int j = 0; j < 8; ++j) ag[i].v[j] = ab[i].v[j] = at0[i].v[j] = 0.f;
    if (i < nv) gm[i] = load8(gamma + (i * 32 + lane) * 8);
  }
  for (int row = blockIdx.x * ROW_WARPS + warp; row < rows; row += gridDim.x * ROW_WARPS) {
    // word row: gathered by id, or the caller's inputs_embeds row (then the gradient goes to d_inputs_embeds instead of the table).
    // nn.Embedding(padding_idx=pad_token_id) (bert_model.py:171): the pad row takes part in the forward but receives NO gradient.
    const long long wid = ids ? static_cast<long long>(ids[row]) : -1;
    const size_t pi = static_cast<size_t>(pos ? pos[row] : (row % S)), ti = static_cast<size_t>(tt ? tt[row] : 0);
    const float* wrow = ids ? word + static_cast<size_t>(wid) * H : inputs_embeds + static_cast<size_t>(row) * H;
    float* dwrow = ids ? ((dword && wid != pad_id) ? dword + static_cast<size_t>(wid) * H : nullptr)
                       : (d_inputs_embeds ? d_inputs_embeds + static_cast<size_t>(row) * H : nullptr);
    Vec8 v[ROW_MAXV];
#pragma unroll
    for (int i = 0; i < ROW_MAXV; ++i)
      if (i < nv) {
        const int c = (i * 32 + lane) * 8;
        const Vec8 a = load8(wrow + c), b = load8(pos_tab + pi * H + c), d = load8(type_tab + ti * H + c);
#pragma unroll
        for (int j = 0; j < 8; ++j) v[i].v[j] = a.v[j] + d.v[j] + b.v[j];
      }
    float mean, rstd;
    row_stats(v, nv, H, eps, mean, rstd);
    Vec8 g[ROW_MAXV];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < ROW_MAXV; ++i)
      if (i < nv) {
        const size_t off = static_cast<size_t>(row) * H + (i * 32 + lane) * 8;
        Vec8 d = load8(dy + off);
        if (dy2) {
          const Vec8 d2 = load8(dy2 + off);
#pragma unroll
          for (int j = 0; j < 8; ++j) d.v[j] += d2.v[j];
        }
        if (drop.seed_base) drop8(d, static_cast<uint32_t>(off), dseed, drop);      // gradient of the embedding dropout
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          v[i].v[j] = (v[i].v[j] - mean) * rstd;      // xhat
          g[i].v[j] = d.v[j] * gm[i].v[j];
          s1 += g[i].v[j];
          s2 = fmaf(g[i].v[j], v[i].v[j], s2);
          ag[i].v[j] = fmaf(d.v[j], v[i].v[j], ag[i].v[j]);
          ab[i].v[j] += d.v[j];
        }
      }
    s1 = warp_sum(s1) / H;
    s2 = warp_sum(s2) / H;
#pragma unroll
    for (int i = 0; i < ROW_MAXV; ++i)
      if (i < nv) {
        const int c = (i * 32 + lane) * 8;
        float dx[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) dx[j] = rstd * (g[i].v[j] - s1 - v[i].v[j] * s2) * alpha;
        if (dwrow) {
          if (ids) {
            red_add_v4(dwrow + c, dx[0], dx[1], dx[2], dx[3]);
            red_add_v4(dwrow + c + 4, dx[4], dx[5], dx[6], dx[7]);
          } else {                                      // one row per token: a plain store
            *reinterpret_cast<float4*>(dwrow + c) = make_float4(dx[0], dx[1], dx[2], dx[3]);
            *reinterpret_cast<float4*>(dwrow + c + 4) = make_float4(dx[4], dx[5], dx[6], dx[7]);
          }
        }
        red_add_v4(dpos + pi * H + c, dx[0], dx[1], dx[2], dx[3]);
        red_add_v4(dpos + pi * H + c + 4, dx[4], dx[5], dx[6], dx[7]);
        if (ti == 0) {
#pragma unroll
          for (int j = 0; j < 8; ++j) at0[i].v[j] += dx[j];
        } else {
          red_add_v4(dtype_tab + ti * H + c, dx[0], dx[1], dx[2], dx[3]);
          red_add_v4(dtype_tab + ti * H + c + 4, dx[4], dx[5], dx[6], dx[7]);
        }
      }
  }
  float* rg = red;
  float* rb = red + ROW_WARPS * H;
  float* rt = red + 2 * ROW_WARPS * H;
#pragma unroll
  for (int i = 0; i < ROW_MAXV; ++i)
    if (i < nv) {
      const int c = (i * 32 + lane) * 8;
      store8(rg + warp * H + c, ag[i]);
      store8(rb + warp * H + c, ab[i]);
      store8(rt + warp * H + c, at0[i]);
    }
  __syncthreads();
  for (int c = threadIdx.x; c < H; c += blockDim.x) {
    float a = 0.f, b = 0.f, t = 0.f;
#pragma unroll
    for (int w = 0; w < ROW_WARPS; ++w) {
      a += rg[w * H + c];
      b += rb[w * H + c];
      t += rt[w * H + c];
    }
    atomicAdd(dgamma + c, a * alpha);
    atomicAdd(dbeta + c, b * alpha);
    atomicAdd(dtype_tab + c, t);
  }
}

// ------------------------------------------------------------------------------------------------ K11: token-cls head
// logits[row, c] = h[row,:] . W[c,:] + b[c]   (C <= 4 classes; Linear(H -> 2/3): loss_calculator.py:17,42,
// modeling_ponet.py:43,83-84).  HBM-bound GEMV-like: each warp streams one row with 128-bit loads.
// Each warp owns CLS_RPW consecutive rows per trip and issues ALL of their 128-bit loads before the first use (a row of H = 768
// halves is only 3 loads per lane: one row per warp left ~1.5 KB in flight per warp and the kernel at 0.14-0.28 of the HBM
// roofline, profiles/r01g); the C weight rows are loop-invariant and live in registers.  Grid-stride over row groups.
constexpr int CLS_RPW = 7;            // rows in flight per warp for 2-byte activations (fp32: 4): 16384 rows / (148 SMs x 16 warps) = 6.9,
constexpr int CLS_WARPS = 16;         // so the bench shape is ONE trip per warp with every load issued up front
// eight elements as they sit in memory (fp16: one 128-bit word, fp32: two): what a load keeps in flight; converted at the use
template <typename T> struct Raw8;
template <> struct Raw8<__half> {
  uint4 a;
  __device__ __forceinline__ void load(const __half* p) { a = *reinterpret_cast<const uint4*>(p); }
  __device__ __forceinline__ Vec8 get() const {
    const __half2* hp = reinterpret_cast<const __half2*>(&a);
    Vec8 r;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float2 f = __half22float2(hp[i]);
      r.v[2 * i] = f.x;
      r.v[2 * i + 1] = f.y;
    }
    return r;
  }
};
template <> struct Raw8<float> {
  float4 a, b;
  __device__ __forceinline__ void load(const float* p) {
    a = *reinterpret_cast<const float4*>(p);
    b = *reinterpret_cast<const float4*>(p + 4);
  }
  __device__ __forceinline__ Vec8 get() const {
    Vec8 r;
    r.v[0] = a.x; r.v[1] = a.y; r.v[2] = a.z; r.v[3] = a.w;
    r.v[4] = b.x; r.v[5] = b.y; r.v[6] = b.z; r.v[7] = b.w;
    return r;
  }
};
// NV = 8-wide vectors per lane = ceil(H / 256): a template parameter so that the in-flight rows cost NV (not ROW_MAXV) registers each
template <int C, typename T, int NV>
__global__ void __launch_bounds__(CLS_WARPS * 32, 1) cls_head_fwd_kernel(const T* __restrict__ h, const float* __restrict__ W,
                                                                       const float* __restrict__ b, float* __restrict__ logits,
                                                                       int32_t* __restrict__ argmax_out, int rows, int H, DropCfg drop) {
  extern __shared__ float w_s[];            // [C][H]: the weight rows, read back 128 bits per lane (conflict-free), so that the
  const int lane = threadIdx.x & 31;        // registers hold activations in flight instead of loop-invariant weights
  for (int k = threadIdx.x; k < C * H; k += blockDim.x) w_s[k] = W[k];
  __syncthreads();
  const int nv = lane_vecs(H, lane);
  const uint32_t dseed = drop.seed_base ? drop_seed(drop) : 0u;
  float bias[C];
#pragma unroll
  for (int c = 0; c < C; ++c) bias[c] = b[c];
  // every warp owns one contiguous range of rows and walks it CLS_RPW rows at a time, all loads of a trip issued up front
  const int n_warps = gridDim.x * CLS_WARPS, wid = blockIdx.x * CLS_WARPS + (threadIdx.x >> 5);
  const int per_warp = (rows + n_warps - 1) / n_warps;
  const int row_end = min(rows, (wid + 1) * per_warp);
  constexpr int RPW = sizeof(T) == 2 ? CLS_RPW : 4;
  for (int row0 = wid * per_warp; row0 < row_end; row0 += RPW) {
    Raw8<T> x[RPW][NV];
#pragma unroll
    for (int rr = 0; rr < RPW; ++rr)
#pragma unroll
      for (int i = 0; i < NV; ++i)
        if (i < nv && row0 + rr < row_end) x[rr][i].load(h + static_cast<size_t>(row0 + rr) * H + (i * 32 + lane) * 8);
    float acc[RPW][C];
#pragma unroll
    for (int rr = 0; rr < RPW; ++rr)
#pragma unroll
      for (int c = 0; c < C; ++c) acc[rr][c] = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i)
      if (i < nv) {
        const int col = (i * 32 + lane) * 8;
        Vec8 w[C];
#pragma unroll
        for (int c = 0; c < C; ++c) w[c] = load8(w_s + c * H + col);
#pragma unroll
        for (int rr = 0; rr < RPW; ++rr)
          if (row0 + rr < row_end) {
            Vec8 xv = x[rr][i].get();
            if (drop.seed_base) drop8(xv, static_cast<uint32_t>(row0 + rr) * static_cast<uint32_t>(H) + col, dseed, drop);   // bert_for_ts.py:66-67
#pragma unroll
            for (int c = 0; c < C; ++c)
#pragma unroll
              for (int j = 0; j < 8; ++j) acc[rr][c] = fmaf(xv.v[j], w[c].v[j], acc[rr][c]);
          }
      }
#pragma unroll
    for (int rr = 0; rr < RPW; ++rr) {
      const int row = row0 + rr;
      if (row >= row_end) break;
      float a[C];
#pragma unroll
      for (int c = 0; c < C; ++c) a[c] = warp_sum(acc[rr][c]) + bias[c];
      if (lane == 0) {
        int best = 0;
#pragma unroll
        for (int c = 0; c < C; ++c) {
          logits[static_cast<size_t>(row) * C + c] = a[c];
          if (a[c] > a[best]) best = c;     // first maximum wins, like np.argmax
        }
        if (argmax_out) argmax_out[row] = best;
      }
    }
  }
}

// Cross-entropy statistics over logits (ignore_index = -100, optional class weights): stats[0] += sum w*nll,
// stats[1] += sum w.   utils.py:173-182 (gamma == 0 path).
template <int C>
__global__ void ce_stats_kernel(const float* __restrict__ logits, const int64_t* __restrict__ labels, const float* __restrict__ cw,
                                float* __restrict__ stats, int rows) {
  float l = 0.f, w = 0.f;
  for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < rows; r += gridDim.x * blockDim.x) {
    const int64_t y = labels[r];
    if (y < 0 || y >= C) continue;
    float m = -INFINITY;
#pragma unroll
    for (int c = 0; c < C; ++c) m = fmaxf(m, logits[static_cast<size_t>(r) * C + c]);
    float s = 0.f;
#pragma unroll
    for (int c = 0; c < C; ++c) s += expf(logits[static_cast<size_t>(r) * C + c] - m);
    const float wy = cw ? cw[y] : 1.0f;
    l += wy * (m + logf(s) - logits[static_cast<size_t>(r) * C + y]);
    w += wy;
  }
  l = warp_sum(l);
  w = warp_sum(w);
  if ((threadIdx.x & 31) == 0) {
    atomicAdd(stats, l);
    atomicAdd(stats + 1, w);
  }
}

// Backward of head + mean CE in one pass over h: dlogits = w_y (softmax - onehot) / sum_w;
// dh = scale * dlogits . W (fp16);  dW += dlogits^T h;  db += sum dlogits.
template <int C>
__global__ void __launch_bounds__(ROW_WARPS * 32) cls_head_bwd_kernel(const __half* __restrict__ h, const float* __restrict__ logits,
                                                                       const int64_t* __restrict__ labels, const float* __restrict__ cw,
                                                                       const float* __restrict__ stats, const float* __restrict__ W,
                                                                       const float* __restrict__ scale_ptr, __half* __restrict__ dh,
                                                                       float* __restrict__ dW, float* __restrict__ db, int rows, int H,
                                                                       DropCfg drop) {
  extern __shared__ float red[];   // [ROW_WARPS][C][H]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int nv = lane_vecs(H, lane);
  const uint32_t dseed = drop.seed_base ? drop_seed(drop) : 0u;
  const float inv_w = 1.0f / stats[1];
  const float scale = scale_ptr ? *scale_ptr : 1.0f;
  Vec8 aw[C][ROW_MAXV];
  float abias[C];
#pragma unroll
  for (int c = 0; c < C; ++c) {
    abias[c] = 0.f;
#pragma unroll
    for (int i = 0; i < ROW_MAXV; ++i)
#pragma unroll
      for (int j = 0; j < 8; ++j) aw[c][i].v[j] = 0.f;
  }
  for (int row = blockIdx.x * ROW_WARPS + warp; row < rows; row += gridDim.x * ROW_WARPS) {
    const int64_t y = labels[row];
    float dl[C];
    const bool valid = (y >= 0 && y < C);
    if (valid) {
      float m = -INFINITY;
#pragma unroll
      for (int c = 0; c < C; ++c) m = fmaxf(m, logits[static_cast<size_t>(row) * C + c]);
      float s = 0.f;
#pragma unroll
      for (int c = 0; c < C; ++c) {
        dl[c] = expf(logits[static_cast<size_t>(row) * C + c] - m);
        s += dl[c];
      }
      const float wy = (cw ? cw[y] : 1.0f) * inv_w;
#pragma unroll
      for (int c = 0; c < C; ++c) dl[c] = wy * (dl[c] / s - (c == y ? 1.0f : 0.0f));
    } else {
#pragma unroll
      for (int c = 0; c < C; ++c) dl[c] = 0.f;
    }
#pragma unroll
    for (int i = 0; i < ROW_MAXV; ++i)
      if (i < nv) {
        const int col = (i * 32 + lane) * 8;
        const size_t off = static_cast<size_t>(row) * H + col;
        Vec8 o;
#pragma unroll
        for (int j = 0; j < 8; ++j) o.v[j] = 0.f;
        if (valid) {
          Vec8 x = load8(h + off);
          if (drop.seed_base) drop8(x, static_cast<uint32_t>(off), dseed, drop);
#pragma unroll
          for (int c = 0; c < C; ++c) {
            const Vec8 w = load8(W + static_cast<size_t>(c) * H + col);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              o.v[j] = fmaf(dl[c] * scale, w.v[j], o.v[j]);
              aw[c][i].v[j] = fmaf(dl[c], x.v[j], aw[c][i].v[j]);
            }
          }
          if (drop.seed_base) drop8(o, static_cast<uint32_t>(off), dseed, drop);
        }
        store8(dh + off, o);
      }
    if (valid && lane == 0)
#pragma unroll
      for (int c = 0; c < C; ++c) abias[c] += dl[c];
  }
#pragma unroll
  for (int c = 0; c < C; ++c)
#pragma unroll
    for (int i = 0; i < ROW_MAXV; ++i)
      if (i < nv) store8(red + (warp * C + c) * H + (i * 32 + lane) * 8, aw[c][i]);
  __syncthreads();
  for (int idx = threadIdx.x; idx < C * H; idx += blockDim.x) {
    float a = 0.f;
#pragma unroll
    for (int w = 0; w < ROW_WARPS; ++w) a += red[w * C * H + idx];
    atomicAdd(dW + idx, a);
  }
  if (lane == 0)
#pragma unroll
    for (int c = 0; c < C; ++c)
      if (abias[c] != 0.f) atomicAdd(db + c, abias[c]);
}

// ------------------------------------------------------------------------------------------------ misc streaming kernels
// db[n] += alpha * sum_m dY[m, n]  (bias gradients of QKV / FFN-up).  Each block: 64 columns x a slab of rows.
__global__ void __launch_bounds__(256) colsum_kernel(const __half* __restrict__ dy, int ld, float* __restrict__ db,
                                                     const float* __restrict__ alpha_ptr, int rows, int cols, int rows_per_block) {
  __shared__ float red[32][65];
  const int cg = threadIdx.x & 7, rl = threadIdx.x >> 3;
  const int c0 = blockIdx.x * 64 + cg * 8;
  const int r0 = blockIdx.y * rows_per_block;
  const int r1 = min(rows, r0 + rows_per_block);
  float acc[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[j] = 0.f;
  if (c0 < cols)
    for (int r = r0 + rl; r < r1; r += 32) {
      const Vec8 v = load8(dy + static_cast<size_t>(r) * ld + c0);
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] += v.v[j];
    }
#pragma unroll
  for (int j = 0; j < 8; ++j) red[rl][cg * 8 + j] = acc[j];
  __syncthreads();
  if (threadIdx.x < 64) {
    float s = 0.f;
#pragma unroll
    for (int r = 0; r < 32; ++r) s += red[r][threadIdx.x];
    const int c = blockIdx.x * 64 + threadIdx.x;
    if (c < cols) atomicAdd(db + c, s * (alpha_ptr ? *alpha_ptr : 1.0f));
  }
}

// fp32 -> fp16 (parameter compute copies)
__global__ void cast_f32_f16_kernel(const float* __restrict__ src, __half* __restrict__ dst, size_t n8) {
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n8; i += static_cast<size_t>(gridDim.x) * blockDim.x)
    store8(dst + i * 8, load8(src + i * 8));
}
// fp32 -> bf16 (operands of the bf16 GEMM arm)
__global__ void cast_f32_bf16_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst, size_t n8) {
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n8; i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const Vec8 v = load8(src + i * 8);
    uint4 raw;
    __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&raw);
#pragma unroll
    for (int k = 0; k < 4; ++k) h[k] = __floats2bfloat162_rn(v.v[2 * k], v.v[2 * k + 1]);
    *reinterpret_cast<uint4*>(dst + i * 8) = raw;
  }
}
// fp16 -> fp32
__global__ void cast_f16_f32_kernel(const __half* __restrict__ src, float* __restrict__ dst, size_t n8) {
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n8; i += static_cast<size_t>(gridDim.x) * blockDim.x)
    store8(dst + i * 8, load8(src + i * 8));
}

// dst(fp32) = src(fp16) * scale[1]   (carry a scaled fp16 gradient back out to an fp32 autograd graph)
__global__ void unscale_cast_f16_f32_kernel(const __half* __restrict__ src, const float* __restrict__ scale, float* __restrict__ dst, size_t n8) {
  const float s = scale ? scale[1] : 1.0f;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n8; i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    Vec8 v = load8(src + i * 8);
#pragma unroll
    for (int j = 0; j < 8; ++j) v.v[j] *= s;
    store8(dst + i * 8, v);
  }
}

// amax(|x|) over an fp32 tensor -> slot[0] (as float bits via atomicMax on uint: valid for non-negative floats)
__global__ void amax_f32_kernel(const float* __restrict__ x, size_t n, unsigned int* __restrict__ slot) {
  float m = 0.f;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n; i += static_cast<size_t>(gridDim.x) * blockDim.x)
    m = fmaxf(m, fabsf(x[i]));
  m = warp_max(m);
  if ((threadIdx.x & 31) == 0) atomicMax(slot, __float_as_uint(m));
}
// scale[0] = 2^floor(log2(target / amax)), scale[1] = 1/scale[0]   (power of two: exact to apply and undo)
__global__ void pick_scale_kernel(const unsigned int* __restrict__ amax_slot, float target, float* __restrict__ scale) {
  const float a = __uint_as_float(*amax_slot);
  float s = 1.0f;
  if (a > 0.f && isfinite(a)) s = exp2f(floorf(log2f(target / a)));
  s = fminf(fmaxf(s, 1.0f), 16777216.0f);
  scale[0] = s;
  scale[1] = 1.0f / s;
}
// dst(fp16) = src(fp32) * scale[0]
__global__ void scale_cast_f32_f16_kernel(const float* __restrict__ src, const float* __restrict__ scale, __half* __restrict__ dst, size_t n8) {
  const float s = scale ? scale[0] : 1.0f;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n8; i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    Vec8 v = load8(src + i * 8);
#pragma unroll
    for (int j = 0; j < 8; ++j) v.v[j] *= s;
    store8(dst + i * 8, v);
  }
}

}  // namespace b200
