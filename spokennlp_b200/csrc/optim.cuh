// Optimizer-side streaming kernels: global grad norm, clip coefficient, fused AdamW that also
// refreshes the fp16 compute copy of the parameters.  Pure HBM streaming over the flat parameter buffer.
#pragma once
#include "ptx.cuh"

namespace b200 {

// sumsq[0] += sum g^2   (non-finite gradients propagate to inf/nan and are caught by clip_coef_kernel)
__global__ void __launch_bounds__(256) grad_sumsq_kernel(const float* __restrict__ g, size_t n4, float* __restrict__ sumsq) {
  float s = 0.f;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n4; i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const float4 v = reinterpret_cast<const float4*>(g)[i];
    s = fmaf(v.x, v.x, s);
    s = fmaf(v.y, v.y, s);
    s = fmaf(v.z, v.z, s);
    s = fmaf(v.w, v.w, s);
  }
  s = warp_sum(s);
  __shared__ float red[8];
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int i = 0; i < 8; ++i) t += red[i];
    atomicAdd(sumsq, t);
  }
}

// coef[0] = grad_mult * min(1, max_norm / (grad_mult * sqrt(sumsq) + 1e-6))   (torch.nn.utils.clip_grad_norm_)
// coef[1] = 1 if the norm is finite else 0 (step is skipped);  coef[2] = the (unclipped) global norm
__global__ void clip_coef_kernel(const float* __restrict__ sumsq, float max_norm, float grad_mult, float* __restrict__ coef) {
  const float norm = grad_mult * sqrtf(sumsq[0]);
  const bool ok = isfinite(norm);
  float c = 1.0f;
  if (max_norm > 0.f && ok) c = fminf(1.0f, max_norm / (norm + 1e-6f));
  coef[0] = ok ? c * grad_mult : 0.f;
  coef[1] = ok ? 1.f : 0.f;
  coef[2] = norm;
}

// torch.optim.AdamW semantics (decoupled weight decay), fp32 state; p16 (optional) receives the fp16 compute copy.
__global__ void __launch_bounds__(256) adamw_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                                    float* __restrict__ v, __half* __restrict__ p16, size_t n4, float lr, float beta1,
                                                    float beta2, float eps, float wd, float bc1, float bc2, const float* __restrict__ coef) {
  const float gm = coef ? coef[0] : 1.0f;
  const bool ok = coef ? coef[1] != 0.f : true;
  const float step = lr / bc1, inv_sqrt_bc2 = rsqrtf(bc2);
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n4; i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    float4 pv = reinterpret_cast<float4*>(p)[i];
    if (ok) {
      const float4 gv = reinterpret_cast<const float4*>(g)[i];
      float4 mv = reinterpret_cast<float4*>(m)[i], vv = reinterpret_cast<float4*>(v)[i];
      float* pp = &pv.x;
      const float* gp = &gv.x;
      float* mp = &mv.x;
      float* vp = &vv.x;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float gk = gp[k] * gm;
        mp[k] = fmaf(beta1, mp[k], (1.f - beta1) * gk);
        vp[k] = fmaf(beta2, vp[k], (1.f - beta2) * gk * gk);
        const float denom = sqrtf(vp[k]) * inv_sqrt_bc2 + eps;
        pp[k] = pp[k] * (1.f - lr * wd) - step * mp[k] / denom;
      }
      reinterpret_cast<float4*>(p)[i] = pv;
      reinterpret_cast<float4*>(m)[i] = mv;
      reinterpret_cast<float4*>(v)[i] = vv;
    }
    if (p16) {
      const __half2 a = __floats2half2_rn(pv.x, pv.y), b2 = __floats2half2_rn(pv.z, pv.w);
      uint2 o;
      o.x = *reinterpret_cast<const uint32_t*>(&a);
      o.y = *reinterpret_cast<const uint32_t*>(&b2);
      reinterpret_cast<uint2*>(p16)[i] = o;
    }
  }
}

// clip_coef_kernel + dynamic loss scaling (torch.cuda.amp.GradScaler semantics, on the device so that it also runs inside a
// replayed CUDA graph): a step whose global gradient norm is not finite is skipped (coef[1] = 0), the loss scale is multiplied
// by `backoff`, and state[1] counts it; after `growth_interval` consecutive good steps the scale is multiplied by `growth`.
//   loss_scale = {scale, 1/scale} (read by the head's backward / the unscaling reductions of the NEXT step),
//   state      = {consecutive good steps, skipped steps in total}.
__global__ void clip_coef_scaled_kernel(const float* __restrict__ sumsq, float max_norm, float grad_mult, float* __restrict__ coef,
                                        float* __restrict__ loss_scale, float* __restrict__ state, float growth_interval, float backoff,
                                        float growth, float min_scale, float max_scale) {
  const float norm = grad_mult * sqrtf(sumsq[0]);
  const bool ok = isfinite(norm);
  float c = 1.0f;
  if (max_norm > 0.f && ok) c = fminf(1.0f, max_norm / (norm + 1e-6f));
  coef[0] = ok ? c * grad_mult : 0.f;
  coef[1] = ok ? 1.f : 0.f;
  coef[2] = norm;
  float sc = loss_scale[0], good = state[0];
  if (!ok) {
    sc = fmaxf(min_scale, sc * backoff);
    good = 0.f;
    state[1] += 1.f;
  } else if (++good >= growth_interval) {
    sc = fminf(max_scale, sc * growth);
    good = 0.f;
  }
  state[0] = good;
  loss_scale[0] = sc;
  loss_scale[1] = 1.0f / sc;
}

// hyper = {lr, beta1, beta2, eps, weight_decay, bias_corr1, bias_corr2, -}: device-resident so that a captured CUDA graph
// of the training step can be replayed with a new learning rate / step count every iteration.
__global__ void set_hyper_kernel(float* __restrict__ hyper, float a, float b, float c, float d, float e, float f, float g, float h) {
  hyper[0] = a; hyper[1] = b; hyper[2] = c; hyper[3] = d; hyper[4] = e; hyper[5] = f; hyper[6] = g; hyper[7] = h;
}
// ZERO: the gradient buffer is cleared on the way out (it is read exactly once per step, so the next step's accumulating
// wgrad / bias reductions start from zero without a separate 0.44 GB fill pass).
template <bool ZERO>
__global__ void __launch_bounds__(256) adamw_dev_kernel(float* __restrict__ p, float* __restrict__ g, float* __restrict__ m,
                                                        float* __restrict__ v, __half* __restrict__ p16, size_t n4,
                                                        const float* __restrict__ hyper, const float* __restrict__ coef) {
  const float lr = hyper[0], beta1 = hyper[1], beta2 = hyper[2], eps = hyper[3], wd = hyper[4], bc1 = hyper[5], bc2 = hyper[6];
  const float gm = coef ? coef[0] : 1.0f;
  const bool ok = coef ? coef[1] != 0.f : true;
  const float step = lr / bc1, inv_sqrt_bc2 = rsqrtf(bc2);
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n4; i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    float4 pv = reinterpret_cast<float4*>(p)[i];
    if (ok) {
      const float4 gv = reinterpret_cast<const float4*>(g)[i];
      float4 mv = reinterpret_cast<float4*>(m)[i], vv = reinterpret_cast<float4*>(v)[i];
      float* pp = &pv.x;
      const float* gp = &gv.x;
      float* mp = &mv.x;
      float* vp = &vv.x;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float gk = gp[k] * gm;
        mp[k] = fmaf(beta1, mp[k], (1.f - beta1) * gk);
        vp[k] = fmaf(beta2, vp[k], (1.f - beta2) * gk * gk);
        const float denom = sqrtf(vp[k]) * inv_sqrt_bc2 + eps;
        pp[k] = pp[k] * (1.f - lr * wd) - step * mp[k] / denom;
      }
      reinterpret_cast<float4*>(p)[i] = pv;
      reinterpret_cast<float4*>(m)[i] = mv;
      reinterpret_cast<float4*>(v)[i] = vv;
    }
    if (ZERO) reinterpret_cast<float4*>(g)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (p16) {
      const __half2 a = __floats2half2_rn(pv.x, pv.y), b2 = __floats2half2_rn(pv.z, pv.w);
      uint2 o;
      o.x = *reinterpret_cast<const uint32_t*>(&a);
      o.y = *reinterpret_cast<const uint32_t*>(&b2);
      reinterpret_cast<uint2*>(p16)[i] = o;
    }
  }
}

}  // namespace b200
