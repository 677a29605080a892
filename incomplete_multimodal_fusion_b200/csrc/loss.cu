// Masked reconstruction losses (reference criterion.py:85-115 MaskedMSELoss, :142-172 MaskedL1Loss,
// norm_pix=False).  The reference makes ~6 elementwise passes over [B, C, H, W] (mse, channel mean,
// mask upsample, multiply, two flatten-sums) plus a host sync on `mask.sum() == 0`; here one pass
// reads only the pixels of MASKED patches, and a one-warp finalize does the per-sample divide and
// the batch nanmean on the device.
#include "common.cuh"
#include "mmf_b200.h"

#include <atomic>

namespace mmf {
extern std::atomic<int64_t> g_launch_count;

template <typename TP>
__device__ __forceinline__ float ld_pred(const TP* p, int64_t i);
template <>
__device__ __forceinline__ float ld_pred<float>(const float* p, int64_t i) { return p[i]; }
template <>
__device__ __forceinline__ float ld_pred<__nv_bfloat16>(const __nv_bfloat16* p, int64_t i) { return __bfloat162float(p[i]); }

// grid: (patches-chunks, B).  Each warp takes one (patch, channel, row-of-patch) line at a time.
// work[0..B) = per-sample error sum (already divided by C), work[B..2B) = per-sample mask pixel count
template <typename TP>
__global__ void masked_loss_fwd_kernel(const TP* __restrict__ pred, const float* __restrict__ target,
                                       const int64_t* __restrict__ mask, int64_t mask_bstride, int C, int H, int W, int P,
                                       int kind, float* __restrict__ work, int64_t B) {
  const int b = blockIdx.y;
  const int nw = W / P, nh = H / P, F = nw * nh;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
  float acc = 0.f;
  int cnt = 0;
  const int64_t lines = (int64_t)F * C * P;
  for (int64_t l = (int64_t)blockIdx.x * nwarp + warp; l < lines; l += (int64_t)gridDim.x * nwarp) {
    const int ph = (int)(l % P);
    const int c = (int)((l / P) % C);
    const int patch = (int)(l / ((int64_t)P * C));
    if (mask != nullptr && mask[b * mask_bstride + patch] == 0) continue;
    const int py = patch / nw, px = patch % nw;
    const int64_t base = (((int64_t)b * C + c) * H + (py * P + ph)) * W + px * P;
    for (int x = lane; x < P; x += 32) {
      const float d = ld_pred<TP>(pred, base + x) - target[base + x];
      acc += kind == 0 ? d * d : fabsf(d);
    }
    if (c == 0 && ph == 0 && lane == 0) cnt += P * P;
  }
  acc = warp_sum(acc);
  if (lane == 0) {
    if (acc != 0.f) atomicAdd(work + b, acc / (float)C);
    if (cnt) atomicAdd(work + B + b, (float)cnt);
  }
}

// one warp: loss = nanmean_b(sum_b / cnt_b) (mask given) or sum / (B*H*W) (no mask); work[2B] = #valid samples
__global__ void masked_loss_finalize_kernel(float* work, int64_t B, int has_mask, float hw, float* loss) {
  const int lane = threadIdx.x;
  float s = 0.f, n = 0.f;
  for (int64_t b = lane; b < B; b += 32) {
    if (has_mask) {
      if (work[B + b] > 0.f) { s += work[b] / work[B + b]; n += 1.f; }
    } else {
      s += work[b];
    }
  }
  s = warp_sum(s);
  n = warp_sum(n);
  if (lane == 0) {
    if (has_mask) { work[2 * B] = n; loss[0] = n > 0.f ? s / n : 0.f; }  // all-zero mask: reference returns 0
    else { work[2 * B] = (float)B; loss[0] = s / ((float)B * hw); }
  }
}

template <typename TP>
__global__ void masked_loss_bwd_kernel(const TP* __restrict__ pred, const float* __restrict__ target,
                                       const int64_t* __restrict__ mask, int64_t mask_bstride, int C, int H, int W, int P,
                                       int kind, const float* __restrict__ work, int64_t B, const float* __restrict__ dloss,
                                       TP* __restrict__ dpred) {
  const int nw = W / P;
  const int F = nw * (H / P);
  const int64_t total = B * C * H * W;
  const float g = dloss[0];
  const float nvalid = work[2 * B];
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int x = (int)(i % W);
    const int y = (int)((i / W) % H);
    const int64_t b = i / ((int64_t)W * H * C);
    float coef;
    if (mask != nullptr) {
      const int patch = (y / P) * nw + x / P;
      const float cnt = work[B + b];
      coef = (mask[b * mask_bstride + patch] != 0 && cnt > 0.f && nvalid > 0.f) ? g / (nvalid * cnt * (float)C) : 0.f;
    } else {
      coef = g / ((float)B * (float)C * (float)H * (float)W);
    }
    float o = 0.f;
    if (coef != 0.f) {
      const float d = ld_pred<TP>(pred, i) - target[i];
      o = kind == 0 ? 2.f * d * coef : (d > 0.f ? coef : (d < 0.f ? -coef : 0.f));
    }
    dpred[i] = (TP)o;
  }
  (void)F;
}

}  // namespace mmf

using namespace mmf;

extern "C" int mmf_masked_loss_fwd(const void* pred, int32_t pred_f32, const float* target, const int64_t* mask,
                                   int64_t mask_bstride, int64_t B, int32_t C, int32_t H, int32_t W, int32_t P, int32_t kind,
                                   float* work, float* loss, mmf_stream_t stream) {
  if (!pred || !target || !work || !loss) MMF_BAD_ARG(1);
  if (B <= 0 || B > 65535 || C <= 0 || P <= 0 || H % P || W % P || (kind != 0 && kind != 1)) MMF_BAD_ARG(2);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  cudaError_t e = cudaMemsetAsync(work, 0, (2 * B + 2) * sizeof(float), st);
  if (e != cudaSuccess) return (int)e;
  const int64_t lines = (int64_t)(H / P) * (W / P) * C * P;
  int gx = (int)ceil_div64(lines, 8 * 4);
  const int cap = (int)ceil_div64(148 * 8, B);
  if (gx > cap) gx = cap;
  if (gx < 1) gx = 1;
  dim3 grid(gx, (unsigned)B);
  if (pred_f32) masked_loss_fwd_kernel<float><<<grid, 256, 0, st>>>((const float*)pred, target, mask, mask_bstride, C, H, W, P, kind, work, B);
  else masked_loss_fwd_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>((const __nv_bfloat16*)pred, target, mask, mask_bstride, C, H, W, P, kind, work, B);
  // without a mask the mean is over B*C*H*W: per-sample sums were divided by C already
  masked_loss_finalize_kernel<<<1, 32, 0, st>>>(work, B, mask != nullptr, (float)H * (float)W, loss);
  g_launch_count.fetch_add(2, std::memory_order_relaxed);
  MMF_LAUNCH_CHECK();
  return 0;
}

extern "C" int mmf_masked_loss_bwd(const void* pred, int32_t pred_f32, const float* target, const int64_t* mask,
                                   int64_t mask_bstride, int64_t B, int32_t C, int32_t H, int32_t W, int32_t P, int32_t kind,
                                   const float* work, const float* dloss, void* dpred, mmf_stream_t stream) {
  if (!pred || !target || !work || !dloss || !dpred) MMF_BAD_ARG(1);
  if (B <= 0 || C <= 0 || P <= 0 || H % P || W % P) MMF_BAD_ARG(2);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int64_t total = B * C * H * W;
  int64_t g = ceil_div64(total, 256);
  if (g > 148 * 16) g = 148 * 16;
  if (pred_f32) masked_loss_bwd_kernel<float><<<(int)g, 256, 0, st>>>((const float*)pred, target, mask, mask_bstride, C, H, W, P, kind, work, B, dloss, (float*)dpred);
  else masked_loss_bwd_kernel<__nv_bfloat16><<<(int)g, 256, 0, st>>>((const __nv_bfloat16*)pred, target, mask, mask_bstride, C, H, W, P, kind, work, B, dloss, (__nv_bfloat16*)dpred);
  g_launch_count.fetch_add(1, std::memory_order_relaxed);
  MMF_LAUNCH_CHECK();
  return 0;
}
