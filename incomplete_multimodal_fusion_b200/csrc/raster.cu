// Device half of the input pipeline: everything the reference's dataset does to a raster AFTER the GeoTIFF
// decode (utils/multimodal_dfc2023.py:99-141 load_dsm / load_rgb / load_sar, the RandomCrop of :53-94), one launch per
// modality on the raw, still-integer (or fp32) batch:
//     transform (SAR: 10 log10(x + 1e-7), clip to [-25, 0]; nan_to_num)  ->  cv2.resize(.., INTER_AREA) by an integer
//     factor  ->  float32  ->  z-score with per-band constants (rgb, sar) or the image's own mean / variance (dsm)
//     ->  crop window (top, left) per sample  ->  fp32 [B, C, Ho, Wo], the tensor the patch embedding reads.
// The host then ships uint8 / uint16 rasters at native size (RGB: 1 byte per value instead of the 4 of the normalised
// fp32 tensor) and never touches the pixels.
//
// cv2's INTER_AREA for integer factors is restated exactly (probed against cv2 4.13, pinned by tests/golden/raster.pt):
//   integers, factor 2 : (a + b + c + d + 2) >> 2
//   integers, factor f : rint(float(sum) * (1.f / f^2)), round-half-even, saturated
//   float32,  factor 2 : ((a + b) + (c + d)) * 0.25f
//   float32,  factor f : the f^2 values in row-major order, summed as sum += ((s0 + s1) + s2) + s3 per group of four,
//                        then singly; result sum * (1.f / f^2)
// and numpy's float64 z-score ((float32 - float64) / float64, rounded once to float32) is done in fp64.
#include "common.cuh"
#include "mmf_b200.h"
#include <atomic>
#include <cooperative_groups.h>
#include <float.h>

namespace mmf {
extern std::atomic<int64_t> g_launch_count;

constexpr int RASTER_MAX_C = 8;
constexpr int RASTER_MAX_F = 8;

struct RasterParams {
  const void* src;       // [B, C, Hs, Ws] raw raster
  int64_t B;
  int C, Hs, Ws, f, Hr, Wr, mode, Ho, Wo;
  double mean[RASTER_MAX_C], stdv[RASTER_MAX_C];
  const int32_t* top;    // [B] crop origin in the RESIZED image, or null (0)
  const int32_t* left;
  float* out;            // [B, C, Ho, Wo]
};

template <typename T> struct RasterT;
template <> struct RasterT<uint8_t> { static constexpr bool is_int = true; static constexpr int maxv = 255; };
template <> struct RasterT<uint16_t> { static constexpr bool is_int = true; static constexpr int maxv = 65535; };
template <> struct RasterT<float> { static constexpr bool is_int = false; static constexpr int maxv = 0; };

// np.nan_to_num on float32
__device__ __forceinline__ float nan_to_num(float v) {
  if (v != v) return 0.f;
  if (v > FLT_MAX) return FLT_MAX;
  if (v < -FLT_MAX) return -FLT_MAX;
  return v;
}

// per-pixel transform ahead of the resize (load_sar: multimodal_dfc2023.py:131-133; load_rgb / load_dsm: nan_to_num)
template <int MODE>
__device__ __forceinline__ float pre_transform(float x) {
  if (MODE == 1) {
    float v = __fmul_rn(10.f, log10f(__fadd_rn(x, 1e-7f)));
    if (v == v) v = fminf(fmaxf(v, -25.f), 0.f);   // np.clip keeps NaN; nan_to_num then zeroes it
    return nan_to_num(v);
  }
  return nan_to_num(x);
}

// value of the resized image at (y, x) of plane `img` -- cv2.resize(INTER_AREA) by the integer factor f, as float32.
// F = compile-time factor (1, 2) or 0 = runtime `f`; factor 2 reads each source row's pixel pair with one load.
template <typename T, int MODE, int F>
__device__ __forceinline__ float resized_at(const T* __restrict__ img, int y, int x, int Ws, int f_rt) {
  const int f = F ? F : f_rt;
  const T* s = img + (int64_t)y * f * Ws + (int64_t)x * f;
  if constexpr (RasterT<T>::is_int) {
    if (f == 1) return (float)s[0];
    if (f == 2) {
      if constexpr (sizeof(T) == 1) {
        const uint32_t a = *reinterpret_cast<const uint16_t*>(s), b = *reinterpret_cast<const uint16_t*>(s + Ws);
        return (float)(((a & 0xff) + (a >> 8) + (b & 0xff) + (b >> 8) + 2) >> 2);
      } else {
        const uint32_t a = *reinterpret_cast<const uint32_t*>(s), b = *reinterpret_cast<const uint32_t*>(s + Ws);
        return (float)(((a & 0xffff) + (a >> 16) + (b & 0xffff) + (b >> 16) + 2) >> 2);
      }
    }
    int sum = 0;
    for (int dy = 0; dy < f; ++dy)
      for (int dx = 0; dx < f; ++dx) sum += (int)s[(int64_t)dy * Ws + dx];
    const float r = rintf(__fmul_rn((float)sum, 1.f / (float)(f * f)));
    return fminf(fmaxf(r, 0.f), (float)RasterT<T>::maxv);
  } else {
    if (f == 1) return pre_transform<MODE>(s[0]);
    if (f == 2) {
      const float2 r0 = *reinterpret_cast<const float2*>(s), r1 = *reinterpret_cast<const float2*>(s + Ws);
      return __fmul_rn(__fadd_rn(__fadd_rn(pre_transform<MODE>(r0.x), pre_transform<MODE>(r0.y)),
                                 __fadd_rn(pre_transform<MODE>(r1.x), pre_transform<MODE>(r1.y))), 0.25f);
    }
    const int area = f * f;
    float sum = 0.f;
    int k = 0;
    auto at = [&](int kk) { return pre_transform<MODE>(s[(int64_t)(kk / f) * Ws + (kk % f)]); };
    for (; k + 4 <= area; k += 4)
      sum = __fadd_rn(sum, __fadd_rn(__fadd_rn(__fadd_rn(at(k), at(k + 1)), at(k + 2)), at(k + 3)));
    for (; k < area; ++k) sum = __fadd_rn(sum, at(k));
    return __fmul_rn(sum, 1.f / (float)area);
  }
}

// modes 0 (constant z-score) and 1 (SAR: dB, clip, constant z-score).  A CTA owns 8 output rows of one (sample, band)
// plane, a warp one row, lanes consecutive pixels: no index division anywhere, source reads and output writes are
// contiguous per warp.  uint8 rasters: the resized value is one of 256 integers, so the fp64 z-score is taken from a
// per-CTA table (one fp64 division per thread instead of one per pixel) -- same bits.
constexpr int RASTER_ROWS = 8;
constexpr int RASTER_U = 4;   // pixels per lane per round: their source loads are all issued before the first store
template <typename T, int MODE, int F>
__global__ void __launch_bounds__(RASTER_ROWS * 32) raster_const_kernel(const RasterParams p) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int c = blockIdx.y;
  const int64_t b = blockIdx.z;
  constexpr bool LUT = sizeof(T) == 1;
  __shared__ float lut[LUT ? 256 : 1];
  const double mean = p.mean[c], stdv = p.stdv[c];
  const double rstd = 1.0 / stdv;   // dB path only (log10f already sits ~1 ulp from numpy's): no fp64 division per pixel
  if (LUT) {
    lut[threadIdx.x] = (float)(((double)threadIdx.x - mean) / stdv);   // RASTER_ROWS * 32 == 256 threads
    __syncthreads();
  }
  const int oy = blockIdx.x * RASTER_ROWS + warp;
  if (oy >= p.Ho) return;
  const int y = oy + (p.top ? p.top[b] : 0), x0 = p.left ? p.left[b] : 0;
  const T* img = reinterpret_cast<const T*>(p.src) + (b * p.C + c) * (int64_t)p.Hs * p.Ws;
  float* orow = p.out + ((b * p.C + c) * p.Ho + oy) * (int64_t)p.Wo;
  for (int base = 0; base < p.Wo; base += 32 * RASTER_U) {
    float v[RASTER_U];
#pragma unroll
    for (int u = 0; u < RASTER_U; ++u) {
      const int ox = min(base + 32 * u + lane, p.Wo - 1);   // clamped: no predicated loads, the store is guarded
      v[u] = resized_at<T, MODE, F>(img, y, x0 + ox, p.Ws, p.f);
    }
#pragma unroll
    for (int u = 0; u < RASTER_U; ++u) {
      const int ox = base + 32 * u + lane;
      if (ox < p.Wo)
        orow[ox] = LUT ? lut[(int)v[u]] : (MODE == 1 ? (float)(((double)v[u] - mean) * rstd) : (float)(((double)v[u] - mean) / stdv));
    }
  }
}

// mode 2 (load_dsm, multimodal_dfc2023.py:99-112): the image's own mean and variance over the WHOLE resized raster
// (all bands), then the crop.  A cluster of RASTER_CL CTAs per sample (one CTA per sample left 148 SMs with 256 uneven
// jobs), a warp per row.  Sweep 1 accumulates sum(x) and sum(x^2) in fp64 (the variance about the fp32-rounded mean m
// follows as (sum(x^2) - 2 m sum(x) + n m^2) / n, exact to fp64 rounding); the CTAs' partial sums meet through
// distributed shared memory, summed in rank order by every CTA (same bits everywhere); sweep 2 re-reads the crop window
// (L2) and writes it.
constexpr int RASTER_CL = 4;
template <typename T, int F>
__global__ void __cluster_dims__(RASTER_CL, 1, 1) __launch_bounds__(256) raster_standardize_kernel(const RasterParams p) {
  namespace cg = cooperative_groups;
  cg::cluster_group cluster = cg::this_cluster();
  __shared__ double sh[2][8];
  __shared__ double part[2];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
  const int rank = (int)cluster.block_rank();
  const int64_t b = blockIdx.x / RASTER_CL;
  const T* img = reinterpret_cast<const T*>(p.src) + b * p.C * (int64_t)p.Hs * p.Ws;
  double s1 = 0.0, s2 = 0.0;
  for (int r = rank * nwarp + warp; r < p.C * p.Hr; r += RASTER_CL * nwarp) {
    const int c = r / p.Hr, y = r - c * p.Hr;
    const T* plane = img + (int64_t)c * p.Hs * p.Ws;
    for (int base = 0; base < p.Wr; base += 32 * RASTER_U) {
      float v[RASTER_U];
#pragma unroll
      for (int u = 0; u < RASTER_U; ++u) v[u] = resized_at<T, 0, F>(plane, y, min(base + 32 * u + lane, p.Wr - 1), p.Ws, p.f);
#pragma unroll
      for (int u = 0; u < RASTER_U; ++u)
        if (base + 32 * u + lane < p.Wr) {
          s1 += (double)v[u];
          s2 += (double)v[u] * (double)v[u];
        }
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    s1 += __shfl_xor_sync(0xffffffffu, s1, o);
    s2 += __shfl_xor_sync(0xffffffffu, s2, o);
  }
  if (lane == 0) { sh[0][warp] = s1; sh[1][warp] = s2; }
  __syncthreads();
  if (threadIdx.x == 0) {
    s1 = s2 = 0.0;
    for (int w = 0; w < nwarp; ++w) { s1 += sh[0][w]; s2 += sh[1][w]; }
    part[0] = s1; part[1] = s2;
  }
  cluster.sync();
  s1 = s2 = 0.0;
  for (int r = 0; r < RASTER_CL; ++r) {
    const double* q = cluster.map_shared_rank(part, r);
    s1 += q[0]; s2 += q[1];
  }
  cluster.sync();   // no CTA leaves (or reuses `part`) while a peer still reads it
  const double n = (double)p.C * p.Hr * p.Wr;
  const float mean = (float)(s1 / n);
  const double md = (double)mean;
  const float var = (float)(fmax(s2 - 2.0 * md * s1 + n * md * md, 0.0) / n);
  const float den = __fsqrt_rn(__fadd_rn(var, 1e-6f));
  const int top = p.top ? p.top[b] : 0, left = p.left ? p.left[b] : 0;
  for (int r = rank * nwarp + warp; r < p.C * p.Ho; r += RASTER_CL * nwarp) {
    const int c = r / p.Ho, oy = r - c * p.Ho;
    const T* plane = img + (int64_t)c * p.Hs * p.Ws;
    float* orow = p.out + ((b * p.C + c) * p.Ho + oy) * (int64_t)p.Wo;
    for (int base = 0; base < p.Wo; base += 32 * RASTER_U) {
      float v[RASTER_U];
#pragma unroll
      for (int u = 0; u < RASTER_U; ++u)
        v[u] = resized_at<T, 0, F>(plane, oy + top, min(base + 32 * u + lane, p.Wo - 1) + left, p.Ws, p.f);
#pragma unroll
      for (int u = 0; u < RASTER_U; ++u)
        if (base + 32 * u + lane < p.Wo) orow[base + 32 * u + lane] = __fdiv_rn(__fsub_rn(v[u], mean), den);
    }
  }
}

// Truncated depth standardisation of the training loop (pretrain_mmae.py:452-459, `--standardize_depth`): per sample,
// sort the n = C*H*W values, keep sorted[k_lo : k_hi] (the reference drops the bottom and top 10 %), and standardise the
// WHOLE image with that slice's mean and unbiased variance.  No sort here: the sample sits in shared memory (n <= 56 K
// floats), two 4-pass radix selects on order-preserving keys find v_lo = sorted[k_lo] and v_hi = sorted[k_hi - 1]; the
// slice is then {v_lo < x < v_hi} plus the right number of copies of the two boundary values (ties), summed in fp64.
constexpr int TRUNC_THREADS = 1024;
__device__ __forceinline__ uint32_t order_key(float f) {
  const uint32_t u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
// value of rank k (0-based, ascending) among xs[0..n): 8 bits per pass, shared-memory histogram of the keys that match
// the prefix found so far
__device__ float radix_select(const float* xs, int n, int k, uint32_t* hist, uint32_t* bcast) {
  uint32_t prefix = 0, mask = 0;
  for (int shift = 24; shift >= 0; shift -= 8) {
    for (int i = threadIdx.x; i < 256; i += blockDim.x) hist[i] = 0;
    __syncthreads();
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
      const uint32_t key = order_key(xs[i]);
      if ((key & mask) == prefix) atomicAdd(&hist[(key >> shift) & 255], 1u);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      uint32_t acc = 0;
      int bkt = 0;
      for (; bkt < 256; ++bkt) {
        if (acc + hist[bkt] > (uint32_t)k) break;
        acc += hist[bkt];
      }
      bcast[0] = (uint32_t)bkt;
      bcast[1] = acc;
    }
    __syncthreads();
    prefix |= bcast[0] << shift;
    mask |= 255u << shift;
    k -= (int)bcast[1];
    __syncthreads();
  }
  const uint32_t u = (prefix & 0x80000000u) ? (prefix & 0x7fffffffu) : ~prefix;
  return __uint_as_float(u);
}

__device__ __forceinline__ double block_sum_d(double v, double* sh) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
  __syncthreads();
  double t = 0.0;
  for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += sh[w];
  return t;
}

__global__ void __launch_bounds__(TRUNC_THREADS) trunc_standardize_kernel(const float* __restrict__ x, float* __restrict__ out,
                                                                          int n, int k_lo, int k_hi) {
  extern __shared__ float xs[];
  __shared__ uint32_t hist[256];
  __shared__ uint32_t bcast[2];
  __shared__ double shd[32];
  const float* xi = x + (int64_t)blockIdx.x * n;
  float* oi = out + (int64_t)blockIdx.x * n;
  for (int i = threadIdx.x; i < n; i += blockDim.x) xs[i] = xi[i];
  __syncthreads();
  const float v_lo = radix_select(xs, n, k_lo, hist, bcast);
  const float v_hi = radix_select(xs, n, k_hi - 1, hist, bcast);
  const int m = k_hi - k_lo;
  // counts of the boundary values inside the slice
  double c_le_lo = 0.0, c_lt_hi = 0.0, s_mid = 0.0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const float v = xs[i];
    c_le_lo += v <= v_lo ? 1.0 : 0.0;
    c_lt_hi += v < v_hi ? 1.0 : 0.0;
    s_mid += (v > v_lo && v < v_hi) ? (double)v : 0.0;
  }
  c_le_lo = block_sum_d(c_le_lo, shd);
  c_lt_hi = block_sum_d(c_lt_hi, shd);
  s_mid = block_sum_d(s_mid, shd);
  double n_lo, n_hi;
  if (v_lo == v_hi) { n_lo = (double)m; n_hi = 0.0; }
  else { n_lo = c_le_lo - (double)k_lo; n_hi = (double)k_hi - c_lt_hi; }
  const double mean = (s_mid + n_lo * (double)v_lo + n_hi * (double)v_hi) / (double)m;
  double q = 0.0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const float v = xs[i];
    const double d = (double)v - mean;
    q += (v > v_lo && v < v_hi) ? d * d : 0.0;
  }
  q = block_sum_d(q, shd);
  q += n_lo * ((double)v_lo - mean) * ((double)v_lo - mean) + n_hi * ((double)v_hi - mean) * ((double)v_hi - mean);
  const float var = (float)(q / (double)(m - 1));   // torch.var: unbiased
  const float mean_f = (float)mean;
  const float den = __fsqrt_rn(__fadd_rn(var, 1e-6f));
  for (int i = threadIdx.x; i < n; i += blockDim.x) oi[i] = __fdiv_rn(__fsub_rn(xs[i], mean_f), den);
}

template <typename T>
static int raster_launch(const RasterParams& p, cudaStream_t st) {
  if (p.mode == 2) {
    const unsigned grid = (unsigned)(p.B * RASTER_CL);
    if (p.f == 2) raster_standardize_kernel<T, 2><<<grid, 256, 0, st>>>(p);
    else if (p.f == 1) raster_standardize_kernel<T, 1><<<grid, 256, 0, st>>>(p);
    else raster_standardize_kernel<T, 0><<<grid, 256, 0, st>>>(p);
  } else {
    if (p.B > 65535) return -12;
    const dim3 grid((unsigned)ceil_div(p.Ho, RASTER_ROWS), (unsigned)p.C, (unsigned)p.B);
#define MMF_RASTER_LAUNCH(MODE)                                                                         \
  do {                                                                                                  \
    if (p.f == 2) raster_const_kernel<T, MODE, 2><<<grid, RASTER_ROWS * 32, 0, st>>>(p);                \
    else if (p.f == 1) raster_const_kernel<T, MODE, 1><<<grid, RASTER_ROWS * 32, 0, st>>>(p);           \
    else raster_const_kernel<T, MODE, 0><<<grid, RASTER_ROWS * 32, 0, st>>>(p);                         \
  } while (0)
    if (p.mode == 1) MMF_RASTER_LAUNCH(1);
    else MMF_RASTER_LAUNCH(0);
#undef MMF_RASTER_LAUNCH
  }
  g_launch_count.fetch_add(1, std::memory_order_relaxed);
  MMF_LAUNCH_CHECK();
  return 0;
}

}  // namespace mmf

extern "C" int mmf_raster_prep(const void* src, int32_t src_dtype, int64_t batch, int32_t C, int32_t Hs, int32_t Ws,
                               int32_t factor, int32_t mode, const double* mean_host, const double* std_host,
                               const int32_t* crop_top, const int32_t* crop_left, int32_t Ho, int32_t Wo, float* out,
                               mmf_stream_t stream) {
  using namespace mmf;
  if (!src || !out) MMF_BAD_ARG(1);
  if (batch <= 0) return 0;
  if (C <= 0 || C > RASTER_MAX_C || Hs <= 0 || Ws <= 0) MMF_BAD_ARG(2);
  if (factor < 1 || factor > RASTER_MAX_F || Hs % factor || Ws % factor) MMF_BAD_ARG(3);   // integer INTER_AREA factors only
  if (mode < 0 || mode > 2) MMF_BAD_ARG(4);
  if (mode == 1 && src_dtype != 2) MMF_BAD_ARG(5);      // the dB transform is restated for float32 rasters
  if (mode != 2 && (!mean_host || !std_host)) MMF_BAD_ARG(6);
  const int Hr = Hs / factor, Wr = Ws / factor;
  if (Ho <= 0 || Wo <= 0 || Ho > Hr || Wo > Wr) MMF_BAD_ARG(7);
  if ((crop_top == nullptr) != (crop_left == nullptr)) MMF_BAD_ARG(8);
  if (!crop_top && (Ho != Hr || Wo != Wr)) MMF_BAD_ARG(9);   // a smaller window needs its origin
  const int elem = src_dtype == 0 ? 1 : (src_dtype == 1 ? 2 : 4);
  if (reinterpret_cast<uintptr_t>(src) & (uintptr_t)(2 * elem - 1)) MMF_BAD_ARG(10);   // factor 2 reads pixel pairs with one load
  RasterParams p{};
  p.src = src; p.B = batch; p.C = C; p.Hs = Hs; p.Ws = Ws; p.f = factor; p.Hr = Hr; p.Wr = Wr; p.mode = mode;
  p.Ho = Ho; p.Wo = Wo; p.top = crop_top; p.left = crop_left; p.out = out;
  for (int c = 0; c < C; ++c) {
    p.mean[c] = mode == 2 ? 0.0 : mean_host[c];
    p.stdv[c] = mode == 2 ? 1.0 : std_host[c];
  }
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  switch (src_dtype) {
    case 0: return raster_launch<uint8_t>(p, st);
    case 1: return raster_launch<uint16_t>(p, st);
    case 2: return raster_launch<float>(p, st);
    default: MMF_BAD_ARG(11);
  }
}

extern "C" int mmf_trunc_standardize(const float* x, float* out, int64_t batch, int32_t n, int32_t k_lo, int32_t k_hi,
                                     mmf_stream_t stream) {
  using namespace mmf;
  if (!x || !out) MMF_BAD_ARG(1);
  if (batch <= 0) return 0;
  const size_t smem = (size_t)n * sizeof(float);
  if (n <= 1 || smem > 220 * 1024) MMF_BAD_ARG(2);      // the sample lives in shared memory
  if (k_lo < 0 || k_hi > n || k_hi - k_lo < 2) MMF_BAD_ARG(3);
  static DeviceOnce attr_done;
  const int attr_dev = current_device();
  if (!attr_done.done(attr_dev)) {
    cudaError_t e = cudaFuncSetAttribute(trunc_standardize_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
    if (e != cudaSuccess) return (int)e;
    attr_done.set(attr_dev);
  }
  trunc_standardize_kernel<<<(unsigned)batch, TRUNC_THREADS, smem, reinterpret_cast<cudaStream_t>(stream)>>>(x, out, n, k_lo, k_hi);
  g_launch_count.fetch_add(1, std::memory_order_relaxed);
  MMF_LAUNCH_CHECK();
  return 0;
}
