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

// Vectorised form for patch widths that are multiples of 8 (every configuration of the reference: P = 16): a thread
// owns 8 consecutive pixels of one patch row (16-byte bf16 / 2 x 16-byte fp32 loads), a warp owns 32 such pieces of the
// same sample; only pieces of MASKED patches are read.  work[] as in the scalar kernel.
template <typename TP>
__global__ void __launch_bounds__(256) masked_loss_fwd_vec_kernel(const TP* __restrict__ pred, const float* __restrict__ target,
                                                                  const int64_t* __restrict__ mask, int64_t mask_bstride, int C, int H,
                                                                  int W, int P, int kind, float* __restrict__ work, int64_t B) {
  const int b = blockIdx.y;
  const int nw = W / P, W8 = W >> 3, P8 = P >> 3;
  const int pieces = C * H * W8;                      // 8-pixel pieces of this sample
  float acc = 0.f;
  int cnt = 0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < pieces; i += gridDim.x * blockDim.x) {
    const int x8 = i % W8, y = (i / W8) % H, c = i / (W8 * H);
    const int patch = (y / P) * nw + x8 / P8;
    if (mask != nullptr && mask[b * mask_bstride + patch] == 0) continue;
    const int64_t off = (((int64_t)b * C + c) * H + y) * W + x8 * 8;
    float pv[8];
    if (sizeof(TP) == 2) {
      const uint4 u = *reinterpret_cast<const uint4*>(pred + off);
      const uint32_t* pu = &u.x;
#pragma unroll
      for (int k = 0; k < 4; ++k) { const float2 f = unpack_bf16(pu[k]); pv[2 * k] = f.x; pv[2 * k + 1] = f.y; }
    } else {
      const float4 a = *reinterpret_cast<const float4*>(pred + off), bq = *reinterpret_cast<const float4*>(pred + off + 4);
      pv[0] = a.x; pv[1] = a.y; pv[2] = a.z; pv[3] = a.w; pv[4] = bq.x; pv[5] = bq.y; pv[6] = bq.z; pv[7] = bq.w;
    }
    const float4 t0 = *reinterpret_cast<const float4*>(target + off), t1 = *reinterpret_cast<const float4*>(target + off + 4);
    const float tv[8] = {t0.x, t0.y, t0.z, t0.w, t1.x, t1.y, t1.z, t1.w};
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const float d = pv[k] - tv[k];
      acc += kind == 0 ? d * d : fabsf(d);
    }
    if (c == 0) cnt += 8;
  }
  __shared__ float red[2][8];
  acc = warp_sum(acc);
  float fc = warp_sum((float)cnt);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) { red[0][warp] = acc; red[1][warp] = fc; }
  __syncthreads();
  if (warp == 0) {
    acc = lane < 8 ? red[0][lane] : 0.f;
    fc = lane < 8 ? red[1][lane] : 0.f;
    acc = warp_sum(acc);
    fc = warp_sum(fc);
    if (lane == 0) {
      if (acc != 0.f) atomicAdd(work + b, acc / (float)C);
      if (fc != 0.f) atomicAdd(work + B + b, fc);
    }
  }
}

template <typename TP>
__global__ void __launch_bounds__(256) masked_loss_bwd_vec_kernel(const TP* __restrict__ pred, const float* __restrict__ target,
                                                                  const int64_t* __restrict__ mask, int64_t mask_bstride, int C, int H,
                                                                  int W, int P, int kind, const float* __restrict__ work, int64_t B,
                                                                  const float* __restrict__ dloss, TP* __restrict__ dpred) {
  const int b = blockIdx.y;
  const int nw = W / P, W8 = W >> 3, P8 = P >> 3;
  const int pieces = C * H * W8;
  const float g = dloss[0], nvalid = work[2 * B], cntb = work[B + b];
  const float coef_m = (cntb > 0.f && nvalid > 0.f) ? g / (nvalid * cntb * (float)C) : 0.f;
  const float coef_u = g / ((float)B * (float)C * (float)H * (float)W);
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < pieces; i += gridDim.x * blockDim.x) {
    const int x8 = i % W8, y = (i / W8) % H, c = i / (W8 * H);
    float coef = coef_u;
    if (mask != nullptr) coef = mask[b * mask_bstride + (y / P) * nw + x8 / P8] != 0 ? coef_m : 0.f;
    const int64_t off = (((int64_t)b * C + c) * H + y) * W + x8 * 8;
    float o[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) o[k] = 0.f;
    if (coef != 0.f) {
      float pv[8];
      if (sizeof(TP) == 2) {
        const uint4 u = *reinterpret_cast<const uint4*>(pred + off);
        const uint32_t* pu = &u.x;
#pragma unroll
        for (int k = 0; k < 4; ++k) { const float2 f = unpack_bf16(pu[k]); pv[2 * k] = f.x; pv[2 * k + 1] = f.y; }
      } else {
        const float4 a = *reinterpret_cast<const float4*>(pred + off), bq = *reinterpret_cast<const float4*>(pred + off + 4);
        pv[0] = a.x; pv[1] = a.y; pv[2] = a.z; pv[3] = a.w; pv[4] = bq.x; pv[5] = bq.y; pv[6] = bq.z; pv[7] = bq.w;
      }
      const float4 t0 = *reinterpret_cast<const float4*>(target + off), t1 = *reinterpret_cast<const float4*>(target + off + 4);
      const float tv[8] = {t0.x, t0.y, t0.z, t0.w, t1.x, t1.y, t1.z, t1.w};
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const float d = pv[k] - tv[k];
        o[k] = kind == 0 ? 2.f * d * coef : (d > 0.f ? coef : (d < 0.f ? -coef : 0.f));
      }
    }
    if (sizeof(TP) == 2) {
      *reinterpret_cast<uint4*>(dpred + off) = make_uint4(pack_bf16(o[0], o[1]), pack_bf16(o[2], o[3]), pack_bf16(o[4], o[5]), pack_bf16(o[6], o[7]));
    } else {
      *reinterpret_cast<float4*>(dpred + off) = make_float4(o[0], o[1], o[2], o[3]);
      *reinterpret_cast<float4*>(dpred + off + 4) = make_float4(o[4], o[5], o[6], o[7]);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Masked cross-entropy (criterion.py:24-58 MaskedCrossEntropyLoss, label_smoothing = 0): per pixel
// logsumexp_c(logits) - logits[target]; patch-mask weighted, per-sample normalised, batch nanmean (same work[] /
// finalize as the reconstruction losses).  A thread owns 8 consecutive pixels and walks the C class planes.
// ------------------------------------------------------------------------------------------------
template <typename TP>
__device__ __forceinline__ void ld8(const TP* p, float (&v)[8]) {
  if (sizeof(TP) == 2) {
    const uint4 u = *reinterpret_cast<const uint4*>(p);
    const uint32_t* pu = &u.x;
#pragma unroll
    for (int k = 0; k < 4; ++k) { const float2 f = unpack_bf16(pu[k]); v[2 * k] = f.x; v[2 * k + 1] = f.y; }
  } else {
    const float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
  }
}

template <typename TP, bool BWD>
__global__ void __launch_bounds__(256) masked_ce_kernel(const TP* __restrict__ logits, const int64_t* __restrict__ target,
                                                        const int64_t* __restrict__ mask, int64_t mask_bstride, int C, int H, int W,
                                                        int P, float* __restrict__ work, int64_t B, const float* __restrict__ dloss,
                                                        TP* __restrict__ dlogits) {
  const int b = blockIdx.y;
  const int nw = W / P, W8 = W >> 3, P8 = P >> 3;
  const int pieces = H * W8;
  const int64_t plane = (int64_t)H * W;
  float acc = 0.f, cnt = 0.f;
  float coef_m = 0.f, coef_u = 0.f;
  if (BWD) {
    const float g = dloss[0], nvalid = work[2 * B], cntb = work[B + b];
    coef_m = (cntb > 0.f && nvalid > 0.f) ? g / (nvalid * cntb) : 0.f;
    coef_u = g / ((float)B * (float)H * (float)W);
  }
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < pieces; i += gridDim.x * blockDim.x) {
    const int x8 = i % W8, y = i / W8;
    const bool on = mask == nullptr || mask[b * mask_bstride + (y / P) * nw + x8 / P8] != 0;
    const int64_t pix = (int64_t)y * W + x8 * 8;
    const TP* base = logits + (int64_t)b * C * plane + pix;
    if (!on) {
      if (BWD) {
        for (int c = 0; c < C; ++c) {
          if (sizeof(TP) == 2) *reinterpret_cast<uint4*>(dlogits + (int64_t)b * C * plane + c * plane + pix) = make_uint4(0, 0, 0, 0);
          else {
            *reinterpret_cast<float4*>(dlogits + (int64_t)b * C * plane + c * plane + pix) = make_float4(0.f, 0.f, 0.f, 0.f);
            *reinterpret_cast<float4*>(dlogits + (int64_t)b * C * plane + c * plane + pix + 4) = make_float4(0.f, 0.f, 0.f, 0.f);
          }
        }
      }
      continue;
    }
    int64_t t[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) t[k] = target[(int64_t)b * plane + pix + k];
    float m[8], s[8], tl[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) { m[k] = -INFINITY; s[k] = 0.f; tl[k] = 0.f; }
    for (int c = 0; c < C; ++c) {
      float v[8];
      ld8<TP>(base + c * plane, v);
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const float mn = fmaxf(m[k], v[k]);
        s[k] = s[k] * __expf(m[k] - mn) + __expf(v[k] - mn);
        m[k] = mn;
        if (t[k] == c) tl[k] = v[k];
      }
    }
    if (!BWD) {
#pragma unroll
      for (int k = 0; k < 8; ++k) acc += m[k] + __logf(s[k]) - tl[k];
      cnt += 8.f;
    } else {
      const float coef = mask != nullptr ? coef_m : coef_u;
      for (int c = 0; c < C; ++c) {
        float v[8], o[8];
        ld8<TP>(base + c * plane, v);
#pragma unroll
        for (int k = 0; k < 8; ++k) o[k] = (__expf(v[k] - m[k]) / s[k] - (t[k] == c ? 1.f : 0.f)) * coef;
        TP* dst = dlogits + (int64_t)b * C * plane + c * plane + pix;
        if (sizeof(TP) == 2) {
          *reinterpret_cast<uint4*>(dst) = make_uint4(pack_bf16(o[0], o[1]), pack_bf16(o[2], o[3]), pack_bf16(o[4], o[5]), pack_bf16(o[6], o[7]));
        } else {
          *reinterpret_cast<float4*>(dst) = make_float4(o[0], o[1], o[2], o[3]);
          *reinterpret_cast<float4*>(dst + 4) = make_float4(o[4], o[5], o[6], o[7]);
        }
      }
    }
  }
  if (!BWD) {
    __shared__ float red[2][8];
    acc = warp_sum(acc);
    cnt = warp_sum(cnt);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) { red[0][warp] = acc; red[1][warp] = cnt; }
    __syncthreads();
    if (warp == 0) {
      acc = lane < 8 ? red[0][lane] : 0.f;
      cnt = lane < 8 ? red[1][lane] : 0.f;
      acc = warp_sum(acc);
      cnt = warp_sum(cnt);
      if (lane == 0 && cnt != 0.f) { atomicAdd(work + b, acc); atomicAdd(work + B + b, cnt); }
    }
  }
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
  const bool vec = (P % 8 == 0) && (W % 8 == 0) && (reinterpret_cast<uintptr_t>(pred) % 16 == 0) && (reinterpret_cast<uintptr_t>(target) % 16 == 0);
  if (vec) {
    const int64_t pieces = (int64_t)C * H * (W / 8);
    if (pieces > INT32_MAX) MMF_BAD_ARG(3);
    int vx = (int)ceil_div64(pieces, 256 * 4);
    const int vcap = (int)ceil_div64(148 * 16, B);
    if (vx > vcap) vx = vcap;
    if (vx < 1) vx = 1;
    dim3 vgrid(vx, (unsigned)B);
    if (pred_f32) masked_loss_fwd_vec_kernel<float><<<vgrid, 256, 0, st>>>((const float*)pred, target, mask, mask_bstride, C, H, W, P, kind, work, B);
    else masked_loss_fwd_vec_kernel<__nv_bfloat16><<<vgrid, 256, 0, st>>>((const __nv_bfloat16*)pred, target, mask, mask_bstride, C, H, W, P, kind, work, B);
  } else if (pred_f32) masked_loss_fwd_kernel<float><<<grid, 256, 0, st>>>((const float*)pred, target, mask, mask_bstride, C, H, W, P, kind, work, B);
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
  const bool vec = (P % 8 == 0) && (W % 8 == 0) && B <= 65535 && (reinterpret_cast<uintptr_t>(pred) % 16 == 0) &&
                   (reinterpret_cast<uintptr_t>(target) % 16 == 0) && (reinterpret_cast<uintptr_t>(dpred) % 16 == 0) &&
                   (int64_t)C * H * (W / 8) <= INT32_MAX;
  if (vec) {
    const int64_t pieces = (int64_t)C * H * (W / 8);
    int vx = (int)ceil_div64(pieces, 256 * 4);
    const int vcap = (int)ceil_div64(148 * 16, B);
    if (vx > vcap) vx = vcap;
    if (vx < 1) vx = 1;
    dim3 vgrid(vx, (unsigned)B);
    if (pred_f32) masked_loss_bwd_vec_kernel<float><<<vgrid, 256, 0, st>>>((const float*)pred, target, mask, mask_bstride, C, H, W, P, kind, work, B, dloss, (float*)dpred);
    else masked_loss_bwd_vec_kernel<__nv_bfloat16><<<vgrid, 256, 0, st>>>((const __nv_bfloat16*)pred, target, mask, mask_bstride, C, H, W, P, kind, work, B, dloss, (__nv_bfloat16*)dpred);
  } else if (pred_f32) masked_loss_bwd_kernel<float><<<(int)g, 256, 0, st>>>((const float*)pred, target, mask, mask_bstride, C, H, W, P, kind, work, B, dloss, (float*)dpred);
  else masked_loss_bwd_kernel<__nv_bfloat16><<<(int)g, 256, 0, st>>>((const __nv_bfloat16*)pred, target, mask, mask_bstride, C, H, W, P, kind, work, B, dloss, (__nv_bfloat16*)dpred);
  g_launch_count.fetch_add(1, std::memory_order_relaxed);
  MMF_LAUNCH_CHECK();
  return 0;
}

static int ce_grid(int64_t pieces, int64_t B) {
  int vx = (int)mmf::ceil_div64(pieces, 256 * 2);
  const int vcap = (int)mmf::ceil_div64(148 * 16, B);
  if (vx > vcap) vx = vcap;
  return vx < 1 ? 1 : vx;
}

extern "C" int mmf_masked_ce_fwd(const void* logits, int32_t logits_f32, const int64_t* target, const int64_t* mask,
                                 int64_t mask_bstride, int64_t B, int32_t C, int32_t H, int32_t W, int32_t P, float* work,
                                 float* loss, mmf_stream_t stream) {
  if (!logits || !target || !work || !loss) MMF_BAD_ARG(1);
  if (B <= 0 || B > 65535 || C <= 0 || P <= 0 || H % P || W % P || (P & 7) || (W & 7)) MMF_BAD_ARG(2);
  if (reinterpret_cast<uintptr_t>(logits) & 15) MMF_BAD_ARG(3);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  cudaError_t e = cudaMemsetAsync(work, 0, (2 * B + 2) * sizeof(float), st);
  if (e != cudaSuccess) return (int)e;
  dim3 grid(ce_grid((int64_t)H * (W / 8), B), (unsigned)B);
  if (logits_f32) masked_ce_kernel<float, false><<<grid, 256, 0, st>>>((const float*)logits, target, mask, mask_bstride, C, H, W, P, work, B, nullptr, nullptr);
  else masked_ce_kernel<__nv_bfloat16, false><<<grid, 256, 0, st>>>((const __nv_bfloat16*)logits, target, mask, mask_bstride, C, H, W, P, work, B, nullptr, nullptr);
  masked_loss_finalize_kernel<<<1, 32, 0, st>>>(work, B, mask != nullptr, (float)H * (float)W, loss);
  g_launch_count.fetch_add(2, std::memory_order_relaxed);
  MMF_LAUNCH_CHECK();
  return 0;
}

extern "C" int mmf_masked_ce_bwd(const void* logits, int32_t logits_f32, const int64_t* target, const int64_t* mask,
                                 int64_t mask_bstride, int64_t B, int32_t C, int32_t H, int32_t W, int32_t P, const float* work,
                                 const float* dloss, void* dlogits, mmf_stream_t stream) {
  if (!logits || !target || !work || !dloss || !dlogits) MMF_BAD_ARG(1);
  if (B <= 0 || B > 65535 || C <= 0 || P <= 0 || H % P || W % P || (P & 7) || (W & 7)) MMF_BAD_ARG(2);
  if ((reinterpret_cast<uintptr_t>(logits) | reinterpret_cast<uintptr_t>(dlogits)) & 15) MMF_BAD_ARG(3);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  dim3 grid(ce_grid((int64_t)H * (W / 8), B), (unsigned)B);
  if (logits_f32) masked_ce_kernel<float, true><<<grid, 256, 0, st>>>((const float*)logits, target, mask, mask_bstride, C, H, W, P, const_cast<float*>(work), B, dloss, (float*)dlogits);
  else masked_ce_kernel<__nv_bfloat16, true><<<grid, 256, 0, st>>>((const __nv_bfloat16*)logits, target, mask, mask_bstride, C, H, W, P, const_cast<float*>(work), B, dloss, (__nv_bfloat16*)dlogits);
  g_launch_count.fetch_add(1, std::memory_order_relaxed);
  MMF_LAUNCH_CHECK();
  return 0;
}
