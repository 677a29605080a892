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


// ------------------------------------------------------------------------------------------------
// Debiased hard-negative contrastive loss (reference criterion.py:214-268 HardNegtive_loss, called three times per step at
// pretrain_mmae_s2dsm.py:482-492), forward AND the gradients w.r.t. both inputs for an upstream gradient of 1, in three
// launches on a [2B, D] problem (the reference: a B-iteration Python loop for the mask + ~15 ATen launches forward, ~30
// backward).
//   o = cat(normalize(out_1), normalize(out_2))                                                   [2B, D]
//   sim[r, c] = <bf16(o_r), bf16(o_c)> (fp32 accumulate: torch.mm under the reference's autocast), neg = exp(sim / T) over
//   the columns c != r mod B, c != r mod B + B;  pos_r = exp(<o1_b, o2_b> / T) in fp32, b = r mod B
//   hard:  imp = neg^beta;  Ng = max((-tau+ N pos + N sum(imp neg) / sum(imp)) / (1 - tau+), N e^(-1/T)),  N = 2B - 2
//   easy:  Ng = sum(neg);          loss = mean_r(-log(pos / (pos + Ng)))
// hn_prep_kernel (grid B): normalises rows b and b + B, keeps o (fp32), its bf16 rounding and <o1_b, o2_b>.
// hn_row_kernel  (grid 2B): row r's similarities (one warp per column, lanes over D), the row statistics, loss_r and
//                 G[r, :] = d(mean loss) / d sim[r, :] (0 on the two excluded columns), gp[r] = d(mean loss)/d<o1_b,o2_b>.
// hn_grad_kernel (grid 2B): do_r = sum_c (G[r,c] + G[c,r]) bf16(o_c) + (gp[b] + gp[b+B]) o_partner, back through
//                 x / max(|x|, eps); block 0 also sums the row losses.
// ------------------------------------------------------------------------------------------------
constexpr int HN_THREADS = 256;
constexpr int HN_MAX_D = 2048;

__global__ void __launch_bounds__(HN_THREADS) hn_prep_kernel(const float* __restrict__ x1, int64_t ld1, const float* __restrict__ x2, int64_t ld2,
                                                             int B, int D, float* __restrict__ o, float* __restrict__ ob,
                                                             float* __restrict__ nrm, float* __restrict__ pdot) {
  __shared__ float red[32];
  const int b = blockIdx.x;
  float a[HN_MAX_D / HN_THREADS], c[HN_MAX_D / HN_THREADS];
  float sa = 0.f, sc = 0.f;
#pragma unroll
  for (int i = 0; i < HN_MAX_D / HN_THREADS; ++i) {
    const int d = threadIdx.x + i * HN_THREADS;
    a[i] = d < D ? x1[(int64_t)b * ld1 + d] : 0.f;
    c[i] = d < D ? x2[(int64_t)b * ld2 + d] : 0.f;
    sa += a[i] * a[i];
    sc += c[i] * c[i];
  }
  const float na = fmaxf(sqrtf(block_sum(sa, red)), 1e-12f);
  const float nc = fmaxf(sqrtf(block_sum(sc, red)), 1e-12f);
  float dot = 0.f;
#pragma unroll
  for (int i = 0; i < HN_MAX_D / HN_THREADS; ++i) {
    const int d = threadIdx.x + i * HN_THREADS;
    if (d < D) {
      const float u = a[i] / na, v = c[i] / nc;
      o[(int64_t)b * D + d] = u;
      o[(int64_t)(b + B) * D + d] = v;
      ob[(int64_t)b * D + d] = __bfloat162float(__float2bfloat16_rn(u));
      ob[(int64_t)(b + B) * D + d] = __bfloat162float(__float2bfloat16_rn(v));
      dot += u * v;
    }
  }
  dot = block_sum(dot, red);
  if (threadIdx.x == 0) { nrm[b] = na; nrm[b + B] = nc; pdot[b] = dot; }
}

__global__ void __launch_bounds__(HN_THREADS) hn_row_kernel(const float* __restrict__ ob, const float* __restrict__ pdot, int B, int D,
                                                            float tau_plus, float beta, float inv_t, int easy,
                                                            float* __restrict__ G, float* __restrict__ gp, float* __restrict__ row_loss) {
  extern __shared__ float hn_smem[];
  float* srow = hn_smem;             // [D] this row, bf16-rounded
  float* sneg = hn_smem + D;         // [2B] neg of every column (0 on the excluded ones)
  __shared__ float red[32];
  const int r = blockIdx.x, n2 = 2 * B, b = r % B;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = HN_THREADS / 32;
  for (int d = threadIdx.x; d < D; d += HN_THREADS) srow[d] = ob[(int64_t)r * D + d];
  __syncthreads();
  float A = 0.f, S = 0.f;            // sum neg^(beta+1), sum neg^beta  (easy: A = sum neg)
  for (int c = warp; c < n2; c += nwarp) {
    const float* oc = ob + (int64_t)c * D;
    float acc = 0.f;
    for (int d = lane; d < D; d += 32) acc = fmaf(srow[d], oc[d], acc);
    acc = warp_sum(acc);
    if (lane == 0) {
      float neg = 0.f;
      if (c != b && c != b + B) {
        neg = __expf(acc * inv_t);
        if (easy) { A += neg; } else { const float imp = __expf(beta * acc * inv_t); A += imp * neg; S += imp; }
      }
      sneg[c] = neg;
    }
  }
  A = block_sum(lane == 0 ? A : 0.f, red);
  S = block_sum(lane == 0 ? S : 0.f, red);
  const float N = (float)(n2 - 2);
  const float pos = __expf(pdot[b] * inv_t);
  float Ng, dNg_dpos = 0.f;
  bool live = true;                   // does the gradient reach the negatives (not clamped)?
  if (easy) {
    Ng = A;
  } else {
    const float raw = (-tau_plus * N * pos + N * A / S) / (1.f - tau_plus);
    const float floor_ = N * __expf(-inv_t);
    live = raw >= floor_;
    Ng = live ? raw : floor_;
    dNg_dpos = live ? -tau_plus * N / (1.f - tau_plus) : 0.f;
  }
  const float w = 1.0f / (float)n2;
  const float dL_dNg = w / (pos + Ng);
  const float dL_dpos = w * (1.f / (pos + Ng) - 1.f / pos) + dL_dNg * dNg_dpos;
  if (threadIdx.x == 0) {
    row_loss[r] = -logf(pos / (pos + Ng));
    gp[r] = dL_dpos * pos * inv_t;
  }
  const float ratio = easy ? 0.f : A / S;
  const float k = easy ? dL_dNg * inv_t : dL_dNg * N / (1.f - tau_plus) * inv_t / S;
  for (int c = threadIdx.x; c < n2; c += HN_THREADS) {
    const float neg = sneg[c];
    float g = 0.f;
    if (neg > 0.f && live) {
      if (easy) g = k * neg;
      else g = k * __powf(neg, beta) * ((beta + 1.f) * neg - beta * ratio);
    }
    G[(int64_t)r * n2 + c] = g;
  }
}

__global__ void __launch_bounds__(HN_THREADS) hn_grad_kernel(const float* __restrict__ o, const float* __restrict__ ob, const float* __restrict__ nrm,
                                                             const float* __restrict__ G, const float* __restrict__ gp,
                                                             const float* __restrict__ row_loss, int B, int D, float* __restrict__ loss,
                                                             float* __restrict__ d1, float* __restrict__ d2) {
  extern __shared__ float hn_smem[];
  float* sw = hn_smem;                // [2B] G[r, c] + G[c, r]
  __shared__ float red[32];
  const int r = blockIdx.x, n2 = 2 * B, b = r % B, partner = r < B ? r + B : r - B;
  for (int c = threadIdx.x; c < n2; c += HN_THREADS) sw[c] = G[(int64_t)r * n2 + c] + G[(int64_t)c * n2 + r];
  __syncthreads();
  const float gpp = gp[b] + gp[b + B];
  float g[HN_MAX_D / HN_THREADS], on[HN_MAX_D / HN_THREADS];
  float dot = 0.f;
#pragma unroll
  for (int i = 0; i < HN_MAX_D / HN_THREADS; ++i) {
    const int d = threadIdx.x + i * HN_THREADS;
    g[i] = on[i] = 0.f;
    if (d < D) {
      float acc = 0.f;
      for (int c = 0; c < n2; ++c) acc = fmaf(sw[c], ob[(int64_t)c * D + d], acc);
      acc = fmaf(gpp, o[(int64_t)partner * D + d], acc);
      g[i] = acc;
      on[i] = o[(int64_t)r * D + d];
      dot += acc * on[i];
    }
  }
  dot = block_sum(dot, red);
  const float inv = 1.0f / nrm[r];
  float* dst = r < B ? d1 + (int64_t)r * D : d2 + (int64_t)(r - B) * D;
#pragma unroll
  for (int i = 0; i < HN_MAX_D / HN_THREADS; ++i) {
    const int d = threadIdx.x + i * HN_THREADS;
    if (d < D) dst[d] = (g[i] - on[i] * dot) * inv;
  }
  if (r == 0) {
    float l = 0.f;
    for (int c = threadIdx.x; c < n2; c += HN_THREADS) l += row_loss[c];
    l = block_sum(l, red);
    if (threadIdx.x == 0) *loss = l / (float)n2;
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

extern "C" int64_t mmf_hardneg_workspace_floats(int32_t B, int32_t D) {
  const int64_t n2 = 2 * (int64_t)B;
  return 2 * n2 * D + n2 * n2 + 4 * n2 + B;
}

extern "C" int mmf_hardneg_loss(const float* out1, int64_t ld1, const float* out2, int64_t ld2, int32_t B, int32_t D, float tau_plus,
                                float beta, float temperature, int32_t easy, float* work, float* loss, float* dout1, float* dout2,
                                mmf_stream_t stream) {
  using namespace mmf;
  if (!out1 || !out2 || !work || !loss || !dout1 || !dout2) MMF_BAD_ARG(1);
  if (B < 2 || D <= 0 || D > HN_MAX_D || ld1 < D || ld2 < D) MMF_BAD_ARG(2);
  if (!(temperature > 0.f) || !(tau_plus < 1.f)) MMF_BAD_ARG(3);
  const int64_t n2 = 2 * (int64_t)B;
  const size_t smem_row = (size_t)(D + n2) * sizeof(float), smem_grad = (size_t)n2 * sizeof(float);
  if (smem_row > 200 * 1024) MMF_BAD_ARG(4);
  float* o = work;
  float* ob = o + n2 * D;
  float* G = ob + n2 * D;
  float* nrm = G + n2 * n2;
  float* gp = nrm + n2;
  float* row_loss = gp + n2;
  float* pdot = row_loss + n2;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (smem_row > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(hn_row_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_row);
    if (e != cudaSuccess) return (int)e;
  }
  hn_prep_kernel<<<B, HN_THREADS, 0, st>>>(out1, ld1, out2, ld2, B, D, o, ob, nrm, pdot);
  hn_row_kernel<<<(int)n2, HN_THREADS, smem_row, st>>>(ob, pdot, B, D, tau_plus, beta, 1.0f / temperature, easy, G, gp, row_loss);
  hn_grad_kernel<<<(int)n2, HN_THREADS, smem_grad, st>>>(o, ob, nrm, G, gp, row_loss, B, D, loss, dout1, dout2);
  g_launch_count.fetch_add(3, std::memory_order_relaxed);
  MMF_LAUNCH_CHECK();
  return 0;
}
