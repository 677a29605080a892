// Multi-tensor AdamW, gradient norm and clip coefficient: the training step around the hot path.
//
// Reference: utils/optim_factory.py:138-176 builds torch.optim.AdamW(betas (0.9, 0.95), weight decay on every
// parameter) and utils/native_scaler.py:20-82 wraps backward / grad-norm (one torch.norm per parameter, stacked) /
// optional clip / step.  Here one launch updates every parameter tensor of the model (HBM bound: 7 fp32 streams per
// element), optionally scaling the gradients by a device-resident clip coefficient (no host sync), and writes the bf16
// weight image the next forward's GEMMs read, so the per-step fp32 -> bf16 weight casts disappear.
#include "common.cuh"
#include "mmf_b200.h"

#include <atomic>

namespace mmf {
extern std::atomic<int64_t> g_launch_count;

constexpr int OPT_THREADS = 256;

__global__ void __launch_bounds__(OPT_THREADS) adamw_kernel(const MmfAdamWTensor* __restrict__ tensors,
                                                            const int32_t* __restrict__ chunk_tensor,
                                                            const int32_t* __restrict__ chunk_index, int chunk_elems, float lr,
                                                            float beta1, float beta2, float eps, float weight_decay,
                                                            const float* __restrict__ grad_scale) {
  const MmfAdamWTensor t = tensors[chunk_tensor[blockIdx.x]];
  const int64_t begin = (int64_t)chunk_index[blockIdx.x] * chunk_elems;
  if (begin >= t.n) return;   // n == 0: the tensor has no gradient this step
  const int64_t end = min(begin + (int64_t)chunk_elems, t.n);
  const float gs = grad_scale ? __ldg(grad_scale) : 1.0f;
  const float decay = 1.0f - lr * weight_decay;
  const float step_size = lr / t.bias_correction1;
  const float inv_sqrt_bc2 = rsqrtf(t.bias_correction2);
  __nv_bfloat16* w16[2] = {reinterpret_cast<__nv_bfloat16*>(t.w16a), reinterpret_cast<__nv_bfloat16*>(t.w16b)};
  const int64_t pitch[2] = {t.pitch16a, t.pitch16b};
  const bool vec = ((reinterpret_cast<uintptr_t>(t.p) | reinterpret_cast<uintptr_t>(t.g) | reinterpret_cast<uintptr_t>(t.m) |
                     reinterpret_cast<uintptr_t>(t.v)) & 15) == 0 && (begin & 3) == 0 &&
                   (t.cols == 0 || ((t.cols & 3) == 0 && (pitch[0] & 3) == 0 && (pitch[1] & 3) == 0 &&
                                    ((reinterpret_cast<uintptr_t>(w16[0]) | reinterpret_cast<uintptr_t>(w16[1])) & 7) == 0));
  auto update = [&](float& p, float g, float& m, float& v) {
    g *= gs;
    p *= decay;                                  // decoupled weight decay (torch.optim.AdamW order)
    m = beta1 * m + (1.0f - beta1) * g;
    v = beta2 * v + (1.0f - beta2) * g * g;
    p -= step_size * m / (sqrtf(v) * inv_sqrt_bc2 + eps);
  };
  auto image_offset = [&](int64_t i, int which) -> int64_t {   // element i of p -> element of the bf16 image
    if (t.cols == 0) return i;
    return (i / t.cols) * pitch[which] + (i % t.cols);
  };
  if (vec) {
    const int64_t n4 = (end - begin) >> 2;
    for (int64_t j = threadIdx.x; j < n4; j += OPT_THREADS) {
      const int64_t i = begin + 4 * j;
      float4 p = *reinterpret_cast<const float4*>(t.p + i), m = *reinterpret_cast<const float4*>(t.m + i),
             v = *reinterpret_cast<const float4*>(t.v + i);
      const float4 g = *reinterpret_cast<const float4*>(t.g + i);
      update(p.x, g.x, m.x, v.x); update(p.y, g.y, m.y, v.y); update(p.z, g.z, m.z, v.z); update(p.w, g.w, m.w, v.w);
      *reinterpret_cast<float4*>(t.p + i) = p;
      *reinterpret_cast<float4*>(t.m + i) = m;
      *reinterpret_cast<float4*>(t.v + i) = v;
      const uint2 b = make_uint2(pack_bf16(p.x, p.y), pack_bf16(p.z, p.w));
#pragma unroll
      for (int w = 0; w < 2; ++w)
        if (w16[w]) *reinterpret_cast<uint2*>(w16[w] + image_offset(i, w)) = b;
    }
    for (int64_t i = begin + 4 * n4 + threadIdx.x; i < end; i += OPT_THREADS) {
      float p = t.p[i], m = t.m[i], v = t.v[i];
      update(p, t.g[i], m, v);
      t.p[i] = p; t.m[i] = m; t.v[i] = v;
      for (int w = 0; w < 2; ++w)
        if (w16[w]) w16[w][image_offset(i, w)] = __float2bfloat16(p);
    }
  } else {
    for (int64_t i = begin + threadIdx.x; i < end; i += OPT_THREADS) {
      float p = t.p[i], m = t.m[i], v = t.v[i];
      update(p, t.g[i], m, v);
      t.p[i] = p; t.m[i] = m; t.v[i] = v;
      for (int w = 0; w < 2; ++w)
        if (w16[w]) w16[w][image_offset(i, w)] = __float2bfloat16(p);
    }
  }
}

// sum of squared gradient elements over every tensor (one atomic per CTA)
__global__ void __launch_bounds__(OPT_THREADS) grad_sqnorm_kernel(const MmfAdamWTensor* __restrict__ tensors,
                                                                  const int32_t* __restrict__ chunk_tensor,
                                                                  const int32_t* __restrict__ chunk_index, int chunk_elems,
                                                                  float* __restrict__ out) {
  const MmfAdamWTensor t = tensors[chunk_tensor[blockIdx.x]];
  const int64_t begin = (int64_t)chunk_index[blockIdx.x] * chunk_elems;
  if (begin >= t.n) return;
  const int64_t end = min(begin + (int64_t)chunk_elems, t.n);
  float s = 0.f;
  if ((reinterpret_cast<uintptr_t>(t.g) & 15) == 0 && (begin & 3) == 0) {
    const int64_t n4 = (end - begin) >> 2;
    for (int64_t j = threadIdx.x; j < n4; j += OPT_THREADS) {
      const float4 g = *reinterpret_cast<const float4*>(t.g + begin + 4 * j);
      s += g.x * g.x + g.y * g.y + g.z * g.z + g.w * g.w;
    }
    for (int64_t i = begin + 4 * n4 + threadIdx.x; i < end; i += OPT_THREADS) s += t.g[i] * t.g[i];
  } else {
    for (int64_t i = begin + threadIdx.x; i < end; i += OPT_THREADS) s += t.g[i] * t.g[i];
  }
  __shared__ float red[OPT_THREADS / 32];
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x < 32) {
    s = threadIdx.x < OPT_THREADS / 32 ? red[threadIdx.x] : 0.f;
    s = warp_sum(s);
    if (threadIdx.x == 0) atomicAdd(out, s);
  }
}

__global__ void clip_coef_kernel(const float* __restrict__ sqnorm, float max_norm, float* __restrict__ coef, float* __restrict__ norm) {
  const float n = sqrtf(*sqnorm);
  if (norm) *norm = n;
  if (coef) *coef = max_norm > 0.f ? fminf(1.0f, max_norm / (n + 1e-6f)) : 1.0f;   // torch.nn.utils.clip_grad_norm_
}

}  // namespace mmf

extern "C" int mmf_adamw_step(const MmfAdamWTensor* tensors, const int32_t* chunk_tensor, const int32_t* chunk_index,
                              int32_t nchunks, int32_t chunk_elems, float lr, float beta1, float beta2, float eps,
                              float weight_decay, const float* grad_scale, mmf_stream_t stream) {
  using namespace mmf;
  if (!tensors || !chunk_tensor || !chunk_index) MMF_BAD_ARG(1);
  if (nchunks <= 0) return 0;
  if (chunk_elems <= 0 || (chunk_elems & 3)) MMF_BAD_ARG(2);
  adamw_kernel<<<nchunks, OPT_THREADS, 0, reinterpret_cast<cudaStream_t>(stream)>>>(tensors, chunk_tensor, chunk_index, chunk_elems, lr,
                                                                                   beta1, beta2, eps, weight_decay, grad_scale);
  g_launch_count.fetch_add(1, std::memory_order_relaxed);
  MMF_LAUNCH_CHECK();
  return 0;
}

extern "C" int mmf_grad_norm(const MmfAdamWTensor* tensors, const int32_t* chunk_tensor, const int32_t* chunk_index, int32_t nchunks,
                             int32_t chunk_elems, float max_norm, float* sqnorm, float* norm, float* clip_coef, mmf_stream_t stream) {
  using namespace mmf;
  if (!tensors || !chunk_tensor || !chunk_index || !sqnorm) MMF_BAD_ARG(1);
  if (chunk_elems <= 0 || (chunk_elems & 3)) MMF_BAD_ARG(2);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  cudaError_t e = cudaMemsetAsync(sqnorm, 0, sizeof(float), st);
  if (e != cudaSuccess) return (int)e;
  if (nchunks > 0) {
    grad_sqnorm_kernel<<<nchunks, OPT_THREADS, 0, st>>>(tensors, chunk_tensor, chunk_index, chunk_elems, sqnorm);
    g_launch_count.fetch_add(1, std::memory_order_relaxed);
  }
  clip_coef_kernel<<<1, 1, 0, st>>>(sqnorm, max_norm, clip_coef, norm);
  g_launch_count.fetch_add(1, std::memory_order_relaxed);
  MMF_LAUNCH_CHECK();
  return 0;
}
