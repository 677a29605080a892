// DINO-style distillation loss between a student and a (detached) teacher feature (reference criterion.py:328-335,
// called three times per step at pretrain_mmae.py:489-493):
//     s = normalize(student), t = normalize(teacher)        (F.normalize: x / max(||x||, 1e-12))
//     loss = mean_b( - sum_d softmax(t / Tt)_d * log_softmax(s / Ts)_d )
// The reference runs ~12 ATen launches forward and ~20 backward per call on a [B, D] problem.  Here one CTA per sample
// does the whole row in one pass and also emits d loss / d student for an upstream gradient of 1 (the teacher is
// detached); the autograd node scales it by the incoming scalar.  fp32 throughout (all of these ops are fp32 under
// the reference's autocast, Appendix A #19).
#include "common.cuh"
#include "mmf_b200.h"
#include <atomic>

namespace mmf {
extern std::atomic<int64_t> g_launch_count;

constexpr int DINO_THREADS = 256;
constexpr int DINO_MAX_PER_THREAD = 8;   // D <= 2048

__device__ __forceinline__ float block_sum(float v, float* red) {
  v = warp_sum(v);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) red[warp] = v;
  __syncthreads();
  float t = (threadIdx.x < (blockDim.x >> 5)) ? red[threadIdx.x] : 0.f;
  if (warp == 0) t = warp_sum(t);
  if (threadIdx.x == 0) red[0] = t;
  __syncthreads();
  return red[0];
}
__device__ __forceinline__ float block_max(float v, float* red) {
  v = warp_max(v);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) red[warp] = v;
  __syncthreads();
  float t = (threadIdx.x < (blockDim.x >> 5)) ? red[threadIdx.x] : -INFINITY;
  if (warp == 0) t = warp_max(t);
  if (threadIdx.x == 0) red[0] = t;
  __syncthreads();
  return red[0];
}

template <typename T>
__device__ __forceinline__ float ld_feat(const T* p, int64_t i);
template <>
__device__ __forceinline__ float ld_feat<float>(const float* p, int64_t i) { return p[i]; }
template <>
__device__ __forceinline__ float ld_feat<__nv_bfloat16>(const __nv_bfloat16* p, int64_t i) { return __bfloat162float(p[i]); }

// grid = B.  row_loss[b] = per-sample loss; dstudent[b, :] = d(mean loss)/d student[b, :] for upstream gradient 1.
template <typename T>
__global__ void __launch_bounds__(DINO_THREADS) dino_loss_kernel(const T* __restrict__ student, int64_t lds, const T* __restrict__ teacher,
                                                                int64_t ldt, int B, int D, float inv_ts, float inv_tt,
                                                                float* __restrict__ row_loss, float* __restrict__ dstudent) {
  __shared__ float red[32];
  const int b = blockIdx.x;
  float s[DINO_MAX_PER_THREAD], t[DINO_MAX_PER_THREAD];
  float ss = 0.f, tt = 0.f;
#pragma unroll
  for (int i = 0; i < DINO_MAX_PER_THREAD; ++i) {
    const int d = threadIdx.x + i * DINO_THREADS;
    s[i] = d < D ? ld_feat<T>(student, (int64_t)b * lds + d) : 0.f;
    t[i] = d < D ? ld_feat<T>(teacher, (int64_t)b * ldt + d) : 0.f;
    ss += s[i] * s[i];
    tt += t[i] * t[i];
  }
  const float ns = fmaxf(sqrtf(block_sum(ss, red)), 1e-12f);
  const float nt = fmaxf(sqrtf(block_sum(tt, red)), 1e-12f);
  // logits
  float ms = -INFINITY, mt = -INFINITY;
#pragma unroll
  for (int i = 0; i < DINO_MAX_PER_THREAD; ++i) {
    const int d = threadIdx.x + i * DINO_THREADS;
    s[i] = s[i] / ns * inv_ts;
    t[i] = t[i] / nt * inv_tt;
    if (d < D) { ms = fmaxf(ms, s[i]); mt = fmaxf(mt, t[i]); }
  }
  ms = block_max(ms, red);
  mt = block_max(mt, red);
  float es = 0.f, et = 0.f;
#pragma unroll
  for (int i = 0; i < DINO_MAX_PER_THREAD; ++i) {
    const int d = threadIdx.x + i * DINO_THREADS;
    if (d < D) { es += __expf(s[i] - ms); et += __expf(t[i] - mt); }
  }
  const float lse_s = ms + logf(block_sum(es, red));
  const float inv_et = 1.0f / block_sum(et, red);
  // loss = - sum_d p_t * (s_d - lse_s);   dL/d logit_s = softmax(s) - p_t   (sum p_t = 1)
  float l = 0.f, dot = 0.f;
  float g[DINO_MAX_PER_THREAD];
#pragma unroll
  for (int i = 0; i < DINO_MAX_PER_THREAD; ++i) {
    const int d = threadIdx.x + i * DINO_THREADS;
    g[i] = 0.f;
    if (d < D) {
      const float pt = __expf(t[i] - mt) * inv_et;
      l -= pt * (s[i] - lse_s);
      g[i] = (__expf(s[i] - lse_s) - pt) * inv_ts / (float)B;   // w.r.t. the normalised student feature
      dot += g[i] * (s[i] / inv_ts);                            // <g, s_n>
    }
  }
  l = block_sum(l, red);
  dot = block_sum(dot, red);
  if (threadIdx.x == 0) row_loss[b] = l;
  // back through x / ||x||: dx = (g - s_n <g, s_n>) / ||x||
#pragma unroll
  for (int i = 0; i < DINO_MAX_PER_THREAD; ++i) {
    const int d = threadIdx.x + i * DINO_THREADS;
    if (d < D) dstudent[(int64_t)b * D + d] = (g[i] - (s[i] / inv_ts) * dot) / ns;
  }
}

}  // namespace mmf

extern "C" int mmf_dino_loss(const void* student, int64_t lds, const void* teacher, int64_t ldt, int32_t is_f32, int32_t B,
                             int32_t D, float student_temp, float teacher_temp, float* row_loss, float* dstudent,
                             mmf_stream_t stream) {
  using namespace mmf;
  if (!student || !teacher || !row_loss || !dstudent) MMF_BAD_ARG(1);
  if (B <= 0 || D <= 0 || D > DINO_THREADS * DINO_MAX_PER_THREAD || lds < D || ldt < D) MMF_BAD_ARG(2);
  if (!(student_temp > 0.f) || !(teacher_temp > 0.f)) MMF_BAD_ARG(3);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (is_f32)
    dino_loss_kernel<float><<<B, DINO_THREADS, 0, st>>>(reinterpret_cast<const float*>(student), lds, reinterpret_cast<const float*>(teacher),
                                                        ldt, B, D, 1.0f / student_temp, 1.0f / teacher_temp, row_loss, dstudent);
  else
    dino_loss_kernel<__nv_bfloat16><<<B, DINO_THREADS, 0, st>>>(reinterpret_cast<const __nv_bfloat16*>(student), lds,
                                                                reinterpret_cast<const __nv_bfloat16*>(teacher), ldt, B, D,
                                                                1.0f / student_temp, 1.0f / teacher_temp, row_loss, dstudent);
  g_launch_count.fetch_add(1, std::memory_order_relaxed);
  MMF_LAUNCH_CHECK();
  return 0;
}
