// Loss heads of the topic-segmentation wrapper on the labelled [BOS] rows, as device kernels.
//
// What they replace (emnlp2023-topic_segmentation/src/models/modules/):
//   loss_calculator.py:38-71   ts loss ("lt": Linear + CE / focal — the full-position head lives in rowwise.cuh; "cos": BCE on the
//                              pair similarities), + cl_loss_weight * CSSL, + tssp_loss_weight * TSSP
//   utils.py:116-138           EopPairCosineSimilarity: a Python loop over examples, boolean-mask gathers, roll, cosine
//   utils.py:141-182           FocalLoss / get_loss_fct
//   tssp.py:26-34              boolean-mask gather -> Linear(H, 3) -> CE
//   cssl.py:20-72, 86-126, 236-263   segment amax + index_select gather, the n x n similarity matrix ("eop_matrix") or the
//                              positive / negative lists ("eop_list"), topic ids by a Python loop over examples
// The reference spends its time here in per-example Python loops and host<->device synchronisations (boolean indexing);
// arithmetic is O(labelled rows x H).  Everything below works on the FLATTENED list of labelled rows: a compaction pass turns
// the [B, S] label tensor into (flat position, example, rank inside the example) triples, a gather makes the [n, H] fp32 row
// matrix R, every head reads R and adds into dR, and one scatter writes dR back into the [B*S, H] gradient of the encoder output.
// The row count n is known on the host (one 16-byte read after the compaction, which depends on the labels only and can run
// before the encoder): grids are exact and nothing is allocated for the worst case.
#pragma once
#include "ptx.cuh"

namespace b200 {

constexpr int HD_WARPS = 8;

// ------------------------------------------------------------------------------------------------ compaction
// Stage 1, one block per example: positions s with key[b, s] != ignore, in increasing s, into tmp[b*S + r]; cnt[b].
__global__ void __launch_bounds__(256) heads_compact_rows_kernel(const int64_t* __restrict__ key, long long ignore, int S, int32_t* __restrict__ tmp,
                                                                 int32_t* __restrict__ cnt) {
  const int b = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  __shared__ int wsum[8];
  __shared__ int base_s;
  if (threadIdx.x == 0) base_s = 0;
  __syncthreads();
  for (int s0 = 0; s0 < S; s0 += 256) {
    const int s = s0 + threadIdx.x;
    const bool keep = s < S && key[static_cast<size_t>(b) * S + s] != ignore;
    const unsigned bal = __ballot_sync(0xffffffffu, keep);
    if (lane == 0) wsum[warp] = __popc(bal);
    __syncthreads();
    int off = base_s;
    for (int w = 0; w < warp; ++w) off += wsum[w];
    if (keep) tmp[static_cast<size_t>(b) * S + off + __popc(bal & ((1u << lane) - 1))] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
      int t = 0;
      for (int w = 0; w < 8; ++w) t += wsum[w];
      base_s += t;
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) cnt[b] = base_s;
}

// Stage 2, one block: start[b] = exclusive scan of cnt, totals[0] = n, totals[1] = max cnt; then the flat lists
// idx[start[b] + r] = b*S + s, ex[..] = b, rank[..] = r  (row-major order of the labelled positions, as `labels != -100` gives).
__global__ void __launch_bounds__(1024) heads_compact_finish_kernel(const int32_t* __restrict__ tmp, const int32_t* __restrict__ cnt, int B, int S,
                                                                    int32_t* __restrict__ start, int32_t* __restrict__ totals, int32_t* __restrict__ idx,
                                                                    int32_t* __restrict__ ex, int32_t* __restrict__ rank) {
  __shared__ int total_s, max_s;
  if (threadIdx.x == 0) {
    int acc = 0, mx = 0;
    for (int b = 0; b < B; ++b) {
      start[b] = acc;
      acc += cnt[b];
      mx = max(mx, cnt[b]);
    }
    start[B] = acc;                     // start has B + 1 entries: it doubles as cu_seqlens of the packed layout
    total_s = acc;
    max_s = mx;
    totals[0] = acc;
    totals[1] = mx;
  }
  __syncthreads();
  for (int e = threadIdx.x; e < B * S; e += blockDim.x) {
    const int b = e / S, r = e - b * S;
    if (r < cnt[b]) {
      const int o = start[b] + r;
      idx[o] = b * S + tmp[e];
      ex[o] = b;
      rank[o] = r;
    }
  }
}

// vals[i] = key[idx[i]] (int64 -> int32): the labels / classes of the compacted rows
__global__ void heads_gather_keys_kernel(const int64_t* __restrict__ key, const int32_t* __restrict__ idx, int n, int32_t* __restrict__ vals) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) vals[i] = static_cast<int32_t>(key[idx[i]]);
}

// out[i] = key[idx[i]] (int64 -> int64): token ids / token types / labels of the packed rows
__global__ void heads_gather_i64_kernel(const int64_t* __restrict__ key, const int32_t* __restrict__ idx, int n, int64_t* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = key[idx[i]];
}
// dst[idx[i], :] = src[i, :] (fp32 rows): packed activations back into the padded [B*S, H] layout the HF interface returns
__global__ void __launch_bounds__(HD_WARPS * 32) heads_unpack_rows_kernel(const float* __restrict__ src, const int32_t* __restrict__ idx, int n, int H,
                                                                         float* __restrict__ dst) {
  const int row = blockIdx.x * HD_WARPS + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= n) return;
  const float4* a = reinterpret_cast<const float4*>(src + static_cast<size_t>(row) * H);
  float4* d = reinterpret_cast<float4*>(dst + static_cast<size_t>(idx[row]) * H);
  for (int c = lane; c < H / 4; c += 32) d[c] = a[c];
}

// topic ids of the labelled rows (cssl.py:252-263 / utils.py:29-40): consecutive rows share an id until a row labelled 0
// (a boundary AFTER it) or the end of an example; seg[i] = number of boundaries among rows < i.  One block, chunked scan.
__global__ void __launch_bounds__(1024) heads_topic_ids_kernel(const int32_t* __restrict__ lab, const int32_t* __restrict__ ex, int n,
                                                               int32_t* __restrict__ seg) {
  __shared__ int wsum[32];
  __shared__ int carry;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (int i0 = 0; i0 < n; i0 += 1024) {
    const int i = i0 + threadIdx.x;
    int bnd = 0;
    if (i < n) bnd = (lab[i] == 0 || i + 1 == n || ex[i + 1] != ex[i]) ? 1 : 0;
    int incl = bnd;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += t;
    }
    if (lane == 31) wsum[warp] = incl;
    __syncthreads();
    int off = carry;
    for (int w = 0; w < warp; ++w) off += wsum[w];
    if (i < n) seg[i] = off + incl - bnd;
    __syncthreads();
    if (threadIdx.x == 1023) carry = off + incl;
    __syncthreads();
  }
}

// ------------------------------------------------------------------------------------------------ row gather / scatter
__global__ void __launch_bounds__(HD_WARPS * 32) heads_gather_rows_kernel(const float* __restrict__ h, const int32_t* __restrict__ idx, int n, int H,
                                                                         float* __restrict__ R) {
  const int row = blockIdx.x * HD_WARPS + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= n) return;
  const float4* src = reinterpret_cast<const float4*>(h + static_cast<size_t>(idx[row]) * H);
  float4* dst = reinterpret_cast<float4*>(R + static_cast<size_t>(row) * H);
  for (int c = lane; c < H / 4; c += 32) dst[c] = src[c];
}
// dh[idx[i], :] += scale * dR[i, :]   (the positions of one list are distinct: plain read-modify-write)
__global__ void __launch_bounds__(HD_WARPS * 32) heads_scatter_rows_kernel(const float* __restrict__ dR, const int32_t* __restrict__ idx, int n, int H,
                                                                          float scale, float* __restrict__ dh) {
  const int row = blockIdx.x * HD_WARPS + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= n) return;
  const float4* src = reinterpret_cast<const float4*>(dR + static_cast<size_t>(row) * H);
  float4* dst = reinterpret_cast<float4*>(dh + static_cast<size_t>(idx[row]) * H);
  for (int c = lane; c < H / 4; c += 32) {
    const float4 a = src[c];
    float4 d = dst[c];
    d.x = fmaf(scale, a.x, d.x); d.y = fmaf(scale, a.y, d.y); d.z = fmaf(scale, a.z, d.z); d.w = fmaf(scale, a.w, d.w);
    dst[c] = d;
  }
}

// ------------------------------------------------------------------------------------------------ CSSL features (cssl.py:236-247)
// F[f, :] = max over positions s of example b with seg_ids[b, s] == slot of h[b, s, :]  (zeros when no position matches:
// scatter_reduce(amax, include_self=False) on a zero tensor), for the slots (b, slot = eop_index[b, j] != 0) in row-major order.
__global__ void __launch_bounds__(HD_WARPS * 32) heads_segmax_fwd_kernel(const float* __restrict__ h, const int64_t* __restrict__ seg_ids,
                                                                        const int32_t* __restrict__ slot_ex, const int32_t* __restrict__ slot_id,
                                                                        int nf, int S, int H, float* __restrict__ F) {
  const int f = blockIdx.x * HD_WARPS + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (f >= nf) return;
  const int b = slot_ex[f];
  const long long k = slot_id[f];
  float* out = F + static_cast<size_t>(f) * H;
  bool any = false;
  for (int c = lane; c < H; c += 32) out[c] = 0.f;
  for (int s = 0; s < S; ++s) {
    if (seg_ids[static_cast<size_t>(b) * S + s] != k) continue;      // warp-uniform
    const float* src = h + (static_cast<size_t>(b) * S + s) * H;
    for (int c = lane; c < H; c += 32) out[c] = any ? fmaxf(out[c], src[c]) : src[c];
    any = true;
  }
}
// backward of the above: the gradient of F[f, c] goes, in equal parts, to the positions that attain the maximum
__global__ void __launch_bounds__(HD_WARPS * 32) heads_segmax_bwd_kernel(const float* __restrict__ h, const int64_t* __restrict__ seg_ids,
                                                                        const int32_t* __restrict__ slot_ex, const int32_t* __restrict__ slot_id,
                                                                        const float* __restrict__ F, const float* __restrict__ dF, int nf, int S, int H,
                                                                        float scale, float* __restrict__ dh) {
  const int f = blockIdx.x * HD_WARPS + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (f >= nf) return;
  const int b = slot_ex[f];
  const long long k = slot_id[f];
  for (int c = lane; c < H; c += 32) {
    const float m = F[static_cast<size_t>(f) * H + c];
    int ties = 0;
    for (int s = 0; s < S; ++s)
      if (seg_ids[static_cast<size_t>(b) * S + s] == k && h[(static_cast<size_t>(b) * S + s) * H + c] == m) ++ties;
    if (!ties) continue;
    const float g = scale * dF[static_cast<size_t>(f) * H + c] / ties;
    for (int s = 0; s < S; ++s)
      if (seg_ids[static_cast<size_t>(b) * S + s] == k && h[(static_cast<size_t>(b) * S + s) * H + c] == m)
        atomicAdd(dh + (static_cast<size_t>(b) * S + s) * H + c, g);       // two slots of one example may share a position id
  }
}

// ------------------------------------------------------------------------------------------------ row norms
// Xn[i] = X[i] / max(|X[i]|, eps), inv[i] = 1 / max(|X[i]|, eps)   (F.cosine_similarity normalises each operand first)
__global__ void __launch_bounds__(HD_WARPS * 32) heads_normalize_kernel(const float* __restrict__ X, int n, int H, float eps, float* __restrict__ Xn,
                                                                       float* __restrict__ inv) {
  const int row = blockIdx.x * HD_WARPS + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= n) return;
  const float* x = X + static_cast<size_t>(row) * H;
  float ss = 0.f;
  for (int c = lane; c < H; c += 32) ss = fmaf(x[c], x[c], ss);
  ss = warp_sum(ss);
  const float r = 1.0f / fmaxf(sqrtf(ss), eps);
  for (int c = lane; c < H; c += 32) Xn[static_cast<size_t>(row) * H + c] = x[c] * r;
  if (lane == 0) inv[row] = r;
}
// gradient through the normalisation: dX[i] += scale * (dXn[i] - Xn[i] <Xn[i], dXn[i]>) * inv[i]
__global__ void __launch_bounds__(HD_WARPS * 32) heads_normalize_bwd_kernel(const float* __restrict__ Xn, const float* __restrict__ inv,
                                                                           const float* __restrict__ dXn, int n, int H, float scale,
                                                                           const float* __restrict__ gscale, float* __restrict__ dX) {
  const int row = blockIdx.x * HD_WARPS + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= n) return;
  const float* xn = Xn + static_cast<size_t>(row) * H;
  const float* g = dXn + static_cast<size_t>(row) * H;
  float dot = 0.f;
  for (int c = lane; c < H; c += 32) dot = fmaf(xn[c], g[c], dot);
  dot = warp_sum(dot);
  const float r = inv[row] * scale * (gscale ? *gscale : 1.0f);      // gscale: the upstream gradient of the scalar loss (device)
  for (int c = lane; c < H; c += 32) dX[static_cast<size_t>(row) * H + c] += (g[c] - xn[c] * dot) * r;
}

__device__ __forceinline__ float warp_dot(const float* __restrict__ a, const float* __restrict__ b, int H, int lane) {
  float s = 0.f;
  for (int c = lane * 4; c < H; c += 128) {
    const float4 x = *reinterpret_cast<const float4*>(a + c), y = *reinterpret_cast<const float4*>(b + c);
    s = fmaf(x.x, y.x, s); s = fmaf(x.y, y.y, s); s = fmaf(x.z, y.z, s); s = fmaf(x.w, y.w, s);
  }
  return warp_sum(s);
}

// ------------------------------------------------------------------------------------------------ adjacent-pair cosine (utils.py:116-138)
// cos[i] = <Rn[i], Rn[next(i)]> / temp, next = cyclic successor inside the example; also scattered into out[b, rank] (a [B, ld]
// matrix pre-filled with -100 by the caller).
__global__ void __launch_bounds__(HD_WARPS * 32) heads_pair_cos_fwd_kernel(const float* __restrict__ Rn, const int32_t* __restrict__ ex,
                                                                          const int32_t* __restrict__ rank, const int32_t* __restrict__ start,
                                                                          const int32_t* __restrict__ cnt, int n, int H, float inv_temp,
                                                                          float* __restrict__ cos_rows, float* __restrict__ out, int ld) {
  const int i = blockIdx.x * HD_WARPS + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (i >= n) return;
  const int b = ex[i], r = rank[i];
  const int j = start[b] + (r + 1) % cnt[b];
  const float c = warp_dot(Rn + static_cast<size_t>(i) * H, Rn + static_cast<size_t>(j) * H, H, lane) * inv_temp;
  if (lane == 0) {
    cos_rows[i] = c;
    out[static_cast<size_t>(b) * ld + r] = c;
  }
}
// dRn[i] += (g[i] Rn[next(i)] + g[prev(i)] Rn[prev(i)]) / temp   (row i is the left operand of its own pair and the right one of prev's)
__global__ void __launch_bounds__(HD_WARPS * 32) heads_pair_cos_bwd_kernel(const float* __restrict__ Rn, const int32_t* __restrict__ ex,
                                                                          const int32_t* __restrict__ rank, const int32_t* __restrict__ start,
                                                                          const int32_t* __restrict__ cnt, const float* __restrict__ g, int n, int H,
                                                                          float inv_temp, float* __restrict__ dRn) {
  const int i = blockIdx.x * HD_WARPS + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (i >= n) return;
  const int b = ex[i], r = rank[i], c = cnt[b];
  const int nx = start[b] + (r + 1) % c, pv = start[b] + (r + c - 1) % c;
  const float gn = g[i] * inv_temp, gp = g[pv] * inv_temp;
  const float* a = Rn + static_cast<size_t>(nx) * H;
  const float* p = Rn + static_cast<size_t>(pv) * H;
  for (int k = lane; k < H; k += 32) dRn[static_cast<size_t>(i) * H + k] += gn * a[k] + gp * p[k];
}

// "cos" score predictor (loss_calculator.py:45-49): BCE-with-logits over the PADDED [B, max_n] matrix, padding (-100 logits with
// -100.0 targets) included as upstream; probs = sigmoid(cos).  stats[0] += sum of the n real terms; the caller adds the padding.
__global__ void heads_bce_fwd_kernel(const float* __restrict__ cos_rows, const int32_t* __restrict__ lab, int n, float* __restrict__ stats) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  float l = 0.f;
  if (i < n) {
    const float x = cos_rows[i], y = static_cast<float>(lab[i]);
    l = fmaxf(x, 0.f) - x * y + log1pf(expf(-fabsf(x)));
  }
  l = warp_sum(l);
  if ((threadIdx.x & 31) == 0 && l != 0.f) atomicAdd(stats, l);
}
__global__ void heads_bce_bwd_kernel(const float* __restrict__ cos_rows, const int32_t* __restrict__ lab, int n, float scale,
                                     const float* __restrict__ gscale, float* __restrict__ g) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) g[i] += scale * (gscale ? *gscale : 1.0f) * (1.0f / (1.0f + expf(-cos_rows[i])) - static_cast<float>(lab[i]));
}

// ------------------------------------------------------------------------------------------------ CSSL "eop_matrix" (cssl.py:20-72)
// One block per column j: e_ij = exp(<Fn_i, Fn_j> / temp); num_j = sum over OTHER rows of j's topic, den_j = num_j + sum over rows
// of other topics.  E [n, n] keeps e_ij for the backward.
__global__ void __launch_bounds__(HD_WARPS * 32) heads_cssl_matrix_fwd_kernel(const float* __restrict__ Fn, const int32_t* __restrict__ seg, int n, int H,
                                                                             float inv_temp, float* __restrict__ E, float* __restrict__ num,
                                                                             float* __restrict__ den) {
  const int j = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  __shared__ float s_num[HD_WARPS], s_oth[HD_WARPS];
  const float* fj = Fn + static_cast<size_t>(j) * H;
  const int tj = seg[j];
  float a = 0.f, o = 0.f;
  for (int i = warp; i < n; i += HD_WARPS) {
    const float e = expf(warp_dot(Fn + static_cast<size_t>(i) * H, fj, H, lane) * inv_temp);
    if (lane == 0) {
      E[static_cast<size_t>(i) * n + j] = e;
      if (seg[i] == tj) a += (i != j) ? e : 0.f;
      else o += e;
    }
  }
  if (lane == 0) {
    s_num[warp] = a;
    s_oth[warp] = o;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    float sa = 0.f, so = 0.f;
    for (int w = 0; w < HD_WARPS; ++w) {
      sa += s_num[w];
      so += s_oth[w];
    }
    num[j] = sa;
    den[j] = sa + so;
  }
}
// loss = mean over the columns with num/den != 0 of -log(num/den) (0 when there are no more than two rows or a single topic:
// cssl.py:264-266); out[0] += weight * loss, coef[j] = -weight / V for the valid columns (0 otherwise).  One block.
__global__ void __launch_bounds__(1024) heads_cssl_matrix_loss_kernel(const float* __restrict__ num, const float* __restrict__ den,
                                                                      const int32_t* __restrict__ seg, int n, float weight, float* __restrict__ out,
                                                                      float* __restrict__ coef) {
  __shared__ float s_l[32];
  __shared__ int s_v[32];
  const bool gate = n > 2 && seg[n - 1] != 0;
  float l = 0.f;
  int v = 0;
  for (int j = threadIdx.x; j < n; j += blockDim.x) {
    const float p = num[j] / den[j];
    if (gate && p != 0.f) {
      l -= logf(p);
      ++v;
    }
  }
  l = warp_sum(l);
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  if ((threadIdx.x & 31) == 0) {
    s_l[threadIdx.x >> 5] = l;
    s_v[threadIdx.x >> 5] = v;
  }
  __syncthreads();
  __shared__ float inv_v;
  if (threadIdx.x == 0) {
    float tl = 0.f;
    int tv = 0;
    for (int w = 0; w < 32; ++w) {
      tl += s_l[w];
      tv += s_v[w];
    }
    inv_v = tv > 0 ? 1.0f / tv : 0.f;
    if (tv > 0) out[0] += weight * tl * inv_v;
  }
  __syncthreads();
  for (int j = threadIdx.x; j < n; j += blockDim.x) {
    const float p = num[j] / den[j];
    coef[j] = (gate && p != 0.f) ? -weight * inv_v : 0.f;
  }
}
// dL/d sim_ij = coef_j e_ij ([same, i != j] / num_j - [i != j or other topic] / den_j);  dFn_r = sum_c (G_rc + G_cr) Fn_c / temp.
// One block per row r; the H-wide accumulator is spread over the block's threads.
__global__ void __launch_bounds__(256) heads_cssl_matrix_bwd_kernel(const float* __restrict__ Fn, const int32_t* __restrict__ seg,
                                                                    const float* __restrict__ E, const float* __restrict__ num,
                                                                    const float* __restrict__ den, const float* __restrict__ coef, int n, int H,
                                                                    float inv_temp, float* __restrict__ dFn) {
  const int r = blockIdx.x;
  extern __shared__ float wgt[];            // [n] : G_rc + G_cr
  const int tr = seg[r];
  const float cr = coef[r], nr = num[r], dr = den[r];
  for (int c = threadIdx.x; c < n; c += blockDim.x) {
    float w = 0.f;
    if (c != r) {
      const float e = E[static_cast<size_t>(r) * n + c];
      const bool same = seg[c] == tr;
      const float cc = coef[c];
      if (cc != 0.f) w += cc * e * ((same ? 1.0f / num[c] : 0.f) - 1.0f / den[c]);      // G_rc: column c
      if (cr != 0.f) w += cr * e * ((same ? 1.0f / nr : 0.f) - 1.0f / dr);              // G_cr: column r (e is symmetric)
    }
    wgt[c] = w * inv_temp;
  }
  __syncthreads();
  for (int k = threadIdx.x; k < H; k += blockDim.x) {
    float acc = 0.f;
    for (int c = 0; c < n; ++c) acc = fmaf(wgt[c], Fn[static_cast<size_t>(c) * H + k], acc);
    dFn[static_cast<size_t>(r) * H + k] += acc;
  }
}

// ------------------------------------------------------------------------------------------------ CSSL "eop_list" (cssl.py:86-166)
// pos [kp, n] / neg [kn, n] index lists (drawn on the host with Python's `random`, as the reference does).
// loss_i = -log(sum_k e_pos / (sum_k e_pos + sum_k e_neg)); out[0] += weight * mean_i loss_i; the per-pair gradient weights are
// kept in gw [kp + kn, n] for the backward.
__global__ void __launch_bounds__(HD_WARPS * 32) heads_cssl_list_fwd_kernel(const float* __restrict__ Fn, const int32_t* __restrict__ pos,
                                                                           const int32_t* __restrict__ neg, int kp, int kn, int n, int H,
                                                                           float inv_temp, float weight, float* __restrict__ out,
                                                                           float* __restrict__ gw) {
  const int i = blockIdx.x * HD_WARPS + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (i >= n) return;
  const float* fi = Fn + static_cast<size_t>(i) * H;
  float sp = 0.f, sn = 0.f;
  for (int k = 0; k < kp + kn; ++k) {
    const int j = k < kp ? pos[static_cast<size_t>(k) * n + i] : neg[static_cast<size_t>(k - kp) * n + i];
    const float e = expf(warp_dot(fi, Fn + static_cast<size_t>(j) * H, H, lane) * inv_temp);
    if (k < kp) sp += e;
    else sn += e;
    if (lane == 0) gw[static_cast<size_t>(k) * n + i] = e;
  }
  __syncwarp();
  if (lane == 0) {
    const float all = sp + sn, w = weight / n;
    atomicAdd(out, -w * logf(sp / all));
    for (int k = 0; k < kp + kn; ++k) {
      const float e = gw[static_cast<size_t>(k) * n + i];
      gw[static_cast<size_t>(k) * n + i] = w * e * (1.0f / all - (k < kp ? 1.0f / sp : 0.f)) * inv_temp;     // dL / d<Fn_i, Fn_j>
    }
  }
}
__global__ void __launch_bounds__(HD_WARPS * 32) heads_cssl_list_bwd_kernel(const float* __restrict__ Fn, const int32_t* __restrict__ pos,
                                                                           const int32_t* __restrict__ neg, const float* __restrict__ gw, int kp, int kn,
                                                                           int n, int H, float* __restrict__ dFn) {
  const int i = blockIdx.x * HD_WARPS + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (i >= n) return;
  for (int k = 0; k < kp + kn; ++k) {
    const int j = k < kp ? pos[static_cast<size_t>(k) * n + i] : neg[static_cast<size_t>(k - kp) * n + i];
    const float w = gw[static_cast<size_t>(k) * n + i];
    for (int c = lane; c < H; c += 32) {
      atomicAdd(dFn + static_cast<size_t>(i) * H + c, w * Fn[static_cast<size_t>(j) * H + c]);
      atomicAdd(dFn + static_cast<size_t>(j) * H + c, w * Fn[static_cast<size_t>(i) * H + c]);
    }
  }
}

// ------------------------------------------------------------------------------------------------ Linear(H, C) + CE on compact rows (TSSP, tssp.py:26-34)
// logits[i] = R[i] . W^T + b; stats[0] += nll_i (target tgt[i]); probs kept for the backward.
template <int C>
__global__ void __launch_bounds__(HD_WARPS * 32) heads_rows_ce_fwd_kernel(const float* __restrict__ R, const float* __restrict__ W, const float* __restrict__ bias,
                                                                         const int32_t* __restrict__ tgt, int n, int H, float* __restrict__ probs,
                                                                         float* __restrict__ stats) {
  const int i = blockIdx.x * HD_WARPS + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (i >= n) return;
  float z[C];
#pragma unroll
  for (int c = 0; c < C; ++c) z[c] = warp_dot(R + static_cast<size_t>(i) * H, W + static_cast<size_t>(c) * H, H, lane) + bias[c];
  if (lane == 0) {
    float m = z[0];
#pragma unroll
    for (int c = 1; c < C; ++c) m = fmaxf(m, z[c]);
    float s = 0.f;
#pragma unroll
    for (int c = 0; c < C; ++c) {
      z[c] = expf(z[c] - m);
      s += z[c];
    }
    const int y = tgt[i];
#pragma unroll
    for (int c = 0; c < C; ++c) probs[static_cast<size_t>(i) * C + c] = z[c] / s;
    atomicAdd(stats, -logf(z[y] / s));
  }
}
// dlogits = scale (probs - onehot); dR[i] += dlogits . W; dW += dlogits^T R; db += sum dlogits
template <int C>
__global__ void __launch_bounds__(HD_WARPS * 32) heads_rows_ce_bwd_kernel(const float* __restrict__ R, const float* __restrict__ W, const int32_t* __restrict__ tgt,
                                                                         const float* __restrict__ probs, int n, int H, float scale,
                                                                         const float* __restrict__ gscale, float* __restrict__ dR,
                                                                         float* __restrict__ dW, float* __restrict__ db) {
  const int i = blockIdx.x * HD_WARPS + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (i >= n) return;
  float dl[C];
  const int y = tgt[i];
  scale *= gscale ? *gscale : 1.0f;
#pragma unroll
  for (int c = 0; c < C; ++c) dl[c] = scale * (probs[static_cast<size_t>(i) * C + c] - (c == y ? 1.0f : 0.0f));
  for (int k = lane; k < H; k += 32) {
    const float x = R[static_cast<size_t>(i) * H + k];
    float acc = 0.f;
#pragma unroll
    for (int c = 0; c < C; ++c) {
      acc = fmaf(dl[c], W[static_cast<size_t>(c) * H + k], acc);
      atomicAdd(dW + static_cast<size_t>(c) * H + k, dl[c] * x);
    }
    dR[static_cast<size_t>(i) * H + k] += acc;
  }
  if (lane == 0)
#pragma unroll
    for (int c = 0; c < C; ++c) atomicAdd(db + c, dl[c]);
}

// ------------------------------------------------------------------------------------------------ focal / CE over ALL positions (utils.py:141-182)
// stats[0] += sum w_y nll (labelled rows), stats[1] += sum w_y, stats[2] += sum over ALL rows of (1 - p_target)^gamma with the
// target of an ignored row taken as class 0 (utils.py:162).  loss = gamma == 0 ? CE : (stats[2] / rows) * CE, CE = stats[0]/stats[1].
template <int C>
__global__ void heads_focal_stats_kernel(const float* __restrict__ logits, const int64_t* __restrict__ labels, const float* __restrict__ cw, float gamma,
                                         int rows, float* __restrict__ stats) {
  float l = 0.f, w = 0.f, f = 0.f;
  for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < rows; r += gridDim.x * blockDim.x) {
    const int64_t yl = labels[r];
    const bool lab = yl >= 0 && yl < C;
    const int y = lab ? static_cast<int>(yl) : 0;
    float m = -INFINITY;
#pragma unroll
    for (int c = 0; c < C; ++c) m = fmaxf(m, logits[static_cast<size_t>(r) * C + c]);
    float s = 0.f;
#pragma unroll
    for (int c = 0; c < C; ++c) s += expf(logits[static_cast<size_t>(r) * C + c] - m);
    const float lp = logits[static_cast<size_t>(r) * C + y] - m - logf(s);
    if (lab) {
      const float wy = cw ? cw[y] : 1.0f;
      l -= wy * lp;
      w += wy;
    }
    if (gamma != 0.f) f += powf(fmaxf(1.0f - expf(lp), 0.f), gamma);
  }
  l = warp_sum(l);
  w = warp_sum(w);
  f = warp_sum(f);
  if ((threadIdx.x & 31) == 0) {
    atomicAdd(stats, l);
    atomicAdd(stats + 1, w);
    if (gamma != 0.f) atomicAdd(stats + 2, f);
  }
}
// Backward of the full-position head under that loss, fp32 in / fp32 out: per row dlogits = a (softmax - onehot_y) with
//   a = scale * ( F w_y / sum_w [labelled]  +  CE (gamma / rows) (1 - p)^(gamma-1) p [gamma != 0, every row] ),  F = gamma ? stats[2]/rows : 1.
// dh[row] += dlogits . W;  dW += dlogits^T h;  db += sum dlogits.
template <int C>
__global__ void __launch_bounds__(HD_WARPS * 32) heads_cls_bwd_kernel(const float* __restrict__ h, const float* __restrict__ logits, const int64_t* __restrict__ labels,
                                                                     const float* __restrict__ cw, const float* __restrict__ stats, const float* __restrict__ W,
                                                                     float gamma, float scale, const float* __restrict__ gscale, int rows, int H,
                                                                     float* __restrict__ dh, float* __restrict__ dW, float* __restrict__ db) {
  extern __shared__ float red[];            // [HD_WARPS][C][H]
  scale *= gscale ? *gscale : 1.0f;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const float ce = stats[0] / stats[1];
  const float Fm = gamma != 0.f ? stats[2] / rows : 1.0f;
  float* my = red + static_cast<size_t>(warp) * C * H;
  for (int k = lane; k < C * H; k += 32) my[k] = 0.f;
  float abias[C];
#pragma unroll
  for (int c = 0; c < C; ++c) abias[c] = 0.f;
  for (int row = blockIdx.x * HD_WARPS + warp; row < rows; row += gridDim.x * HD_WARPS) {
    const int64_t yl = labels[row];
    const bool lab = yl >= 0 && yl < C;
    if (!lab && gamma == 0.f) continue;       // no gradient reaches this row
    const int y = lab ? static_cast<int>(yl) : 0;
    float p[C], m = -INFINITY;
#pragma unroll
    for (int c = 0; c < C; ++c) m = fmaxf(m, logits[static_cast<size_t>(row) * C + c]);
    float s = 0.f;
#pragma unroll
    for (int c = 0; c < C; ++c) {
      p[c] = expf(logits[static_cast<size_t>(row) * C + c] - m);
      s += p[c];
    }
#pragma unroll
    for (int c = 0; c < C; ++c) p[c] /= s;
    float a = lab ? Fm * (cw ? cw[y] : 1.0f) / stats[1] : 0.f;
    if (gamma != 0.f) a += ce * (gamma / rows) * powf(fmaxf(1.0f - p[y], 0.f), gamma - 1.0f) * p[y];
    a *= scale;
    float dl[C];
#pragma unroll
    for (int c = 0; c < C; ++c) dl[c] = a * (p[c] - (c == y ? 1.0f : 0.0f));
    for (int k = lane; k < H; k += 32) {
      const float x = h[static_cast<size_t>(row) * H + k];
      float acc = 0.f;
#pragma unroll
      for (int c = 0; c < C; ++c) {
        acc = fmaf(dl[c], W[static_cast<size_t>(c) * H + k], acc);
        my[c * H + k] = fmaf(dl[c], x, my[c * H + k]);
      }
      dh[static_cast<size_t>(row) * H + k] += acc;
    }
    if (lane == 0)
#pragma unroll
      for (int c = 0; c < C; ++c) abias[c] += dl[c];
  }
  __syncthreads();
  for (int k = threadIdx.x; k < C * H; k += blockDim.x) {
    float a = 0.f;
#pragma unroll
    for (int w = 0; w < HD_WARPS; ++w) a += red[static_cast<size_t>(w) * C * H + k];
    if (a != 0.f) atomicAdd(dW + k, a);
  }
  if (lane == 0)
#pragma unroll
    for (int c = 0; c < C; ++c)
      if (abias[c] != 0.f) atomicAdd(db + c, abias[c]);
}

}  // namespace b200
