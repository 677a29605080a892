// HBM-bound data-movement / elementwise kernels of the MultiMAE path: dtype casts with padding
// (bf16 weight images for the GEMMs), GEGLU / GELU backward, column sums (bias gradients), the
// patch im2col gather (only visible patches), un-patchify, row gathers, batch broadcast / reduce.
// All use 128-bit accesses where alignment allows and grid-stride loops sized to the SM count.
#include "common.cuh"
#include "mmf_b200.h"

#include <atomic>

namespace mmf {
extern std::atomic<int64_t> g_launch_count;

static inline int ew_grid(int64_t work_items, int threads) {
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  int64_t need = ceil_div64(work_items, threads);
  int64_t cap = (int64_t)sms * 16;
  if (need < 1) need = 1;
  return (int)(need < cap ? need : cap);
}
#define MMF_COUNT_LAUNCH() g_launch_count.fetch_add(1, std::memory_order_relaxed)

// ------------------------------------------------------------------------------------------------
// f32 [rows, cols] (ld_src) -> bf16 [rows_pad, cols_pad] (ld_dst), zero padding, optional scale
// ------------------------------------------------------------------------------------------------
__global__ void cast_pad_kernel(const float* __restrict__ src, int64_t rows, int64_t cols, int64_t ld_src,
                                __nv_bfloat16* __restrict__ dst, int64_t rows_pad, int64_t cols_pad, int64_t ld_dst,
                                float scale) {
  pdl_wait();   // launched through launch_pdl (common.cuh)
  const int64_t total = rows_pad * cols_pad;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / cols_pad, c = i % cols_pad;
    float v = 0.f;
    if (r < rows && c < cols) v = src[r * ld_src + c] * scale;
    dst[r * ld_dst + c] = __float2bfloat16(v);
  }
}
// contiguous fast path: n % 4 == 0, 16B aligned
__global__ void cast_vec_kernel(const float4* __restrict__ src, uint2* __restrict__ dst, int64_t n4) {
  pdl_wait();   // launched through launch_pdl (common.cuh)
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    const float4 v = src[i];
    dst[i] = make_uint2(pack_bf16(v.x, v.y), pack_bf16(v.z, v.w));
  }
}

// ------------------------------------------------------------------------------------------------
// GEGLU backward (zorro_utils.py:115-118): u = [value | gate] (bf16 [rows, 2*ipad]), dg bf16 [rows, ipad]
//   dvalue = dg * gelu(gate);  dgate = dg * value * gelu'(gate)
// ------------------------------------------------------------------------------------------------
__global__ void geglu_bwd_kernel(const uint4* __restrict__ u, const uint4* __restrict__ dg, uint4* __restrict__ du,
                                 int64_t rows, int ipad8) {
  const int64_t total = rows * ipad8;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / ipad8;
    const int c = (int)(i % ipad8);
    const uint4 val = u[r * 2 * ipad8 + c];
    const uint4 gat = u[r * 2 * ipad8 + ipad8 + c];
    const uint4 d = dg[i];
    uint4 ov, og;
    const uint32_t* pv = &val.x; const uint32_t* pg = &gat.x; const uint32_t* pd = &d.x;
    uint32_t* qv = &ov.x; uint32_t* qg = &og.x;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float2 fv = unpack_bf16(pv[k]), fg = unpack_bf16(pg[k]), fd = unpack_bf16(pd[k]);
      qv[k] = pack_bf16(fd.x * gelu_erf(fg.x), fd.y * gelu_erf(fg.y));
      qg[k] = pack_bf16(fd.x * fv.x * gelu_erf_grad(fg.x), fd.y * fv.y * gelu_erf_grad(fg.y));
    }
    du[r * 2 * ipad8 + c] = ov;
    du[r * 2 * ipad8 + ipad8 + c] = og;
  }
}

// dpre = dy * gelu'(pre)   (bf16, n % 8 == 0)
__global__ void gelu_bwd_kernel(const uint4* __restrict__ pre, const uint4* __restrict__ dy, uint4* __restrict__ dpre, int64_t n8) {
  pdl_wait();   // launched through launch_pdl (common.cuh)
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += (int64_t)gridDim.x * blockDim.x) {
    const uint4 a = pre[i], d = dy[i];
    uint4 o;
    const uint32_t* pa = &a.x; const uint32_t* pd = &d.x; uint32_t* po = &o.x;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float2 fa = unpack_bf16(pa[k]), fd = unpack_bf16(pd[k]);
      po[k] = pack_bf16(fd.x * gelu_erf_grad(fa.x), fd.y * gelu_erf_grad(fa.y));
    }
    dpre[i] = o;
  }
}

// ------------------------------------------------------------------------------------------------
// column sums of a [rows, cols] matrix (bf16 or f32) into f32 out[cols] (atomic accumulate; caller zeroes)
// each CTA handles a slab of rows; threads own columns (coalesced along the row)
// ------------------------------------------------------------------------------------------------
template <typename T>
__global__ void colsum_kernel(const T* __restrict__ x, int64_t rows, int cols, int64_t ld, float* __restrict__ out,
                              int rows_per_cta) {
  const int64_t r0 = (int64_t)blockIdx.y * rows_per_cta;
  const int64_t r1 = min(r0 + rows_per_cta, rows);
  for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < cols; c += gridDim.x * blockDim.x) {
    float acc = 0.f;
    for (int64_t r = r0; r < r1; ++r) acc += (float)x[r * ld + c];
    atomicAdd(out + c, acc);
  }
}

// bf16 rows, cols % 256 == 0: a thread owns 8 adjacent columns (16-byte loads), 32 column-threads x 8 row-lanes per CTA,
// four rows in flight per thread; the row-lanes are combined in shared memory before one atomic per column
__global__ void __launch_bounds__(256) colsum_vec_kernel(const __nv_bfloat16* __restrict__ x, int64_t rows, int64_t ld,
                                                         float* __restrict__ out, int rows_per_cta) {
  pdl_wait();   // launched through launch_pdl (common.cuh)
  const int tcol = threadIdx.x & 31, trow = threadIdx.x >> 5;
  const int64_t col = ((int64_t)blockIdx.x * 32 + tcol) * 8;
  const int64_t r0 = (int64_t)blockIdx.y * rows_per_cta, r1 = min(r0 + rows_per_cta, rows);
  float acc[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) acc[k] = 0.f;
  auto add8 = [&](const uint4& u) {
    const uint32_t* pu = &u.x;
#pragma unroll
    for (int k = 0; k < 4; ++k) { const float2 f = unpack_bf16(pu[k]); acc[2 * k] += f.x; acc[2 * k + 1] += f.y; }
  };
  int64_t r = r0 + trow;
  for (; r + 24 < r1; r += 32) {
    const uint4 a = *reinterpret_cast<const uint4*>(x + r * ld + col), b = *reinterpret_cast<const uint4*>(x + (r + 8) * ld + col),
                c = *reinterpret_cast<const uint4*>(x + (r + 16) * ld + col), d = *reinterpret_cast<const uint4*>(x + (r + 24) * ld + col);
    add8(a); add8(b); add8(c); add8(d);
  }
  for (; r < r1; r += 8) add8(*reinterpret_cast<const uint4*>(x + r * ld + col));
  __shared__ float red[8][32][9];
#pragma unroll
  for (int k = 0; k < 8; ++k) red[trow][tcol][k] = acc[k];
  __syncthreads();
  const int c = threadIdx.x;   // 256 columns of this CTA, one per thread
  float s = 0.f;
#pragma unroll
  for (int t = 0; t < 8; ++t) s += red[t][c >> 3][c & 7];
  atomicAdd(out + (int64_t)blockIdx.x * 256 + c, s);
}

// ------------------------------------------------------------------------------------------------
// out[b, r, :] (f32, batch stride) = src[r, :]   -- batch-invariant rows (fusion tokens + pos-emb)
// and its backward: dsrc[r, :] = sum_b dout[b, r, :]
// ------------------------------------------------------------------------------------------------
__global__ void bcast_rows_kernel(const float4* __restrict__ src, float* __restrict__ dst, int64_t batch, int64_t rows,
                                  int d4, int64_t dst_batch_stride) {
  const int64_t total = batch * rows * d4;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t b = i / (rows * d4), rem = i % (rows * d4);
    reinterpret_cast<float4*>(dst + b * dst_batch_stride)[rem] = src[rem];
  }
}
__global__ void reduce_batch_kernel(const float* __restrict__ src, float4* __restrict__ dst, int64_t batch, int64_t rows,
                                    int d4, int64_t src_batch_stride) {
  const int64_t total = rows * d4;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int64_t b = 0; b < batch; ++b) {
      const float4 v = reinterpret_cast<const float4*>(src + b * src_batch_stride)[i];
      a.x += v.x; a.y += v.y; a.z += v.z; a.w += v.w;
    }
    dst[i] = a;
  }
}

// ------------------------------------------------------------------------------------------------
// Patch im2col gather (input_adapters.py:110 as a GEMM operand): for the n_keep visible patches only,
// A[b*n_keep + i, (c, ph, pw)] = bf16(img[b, c, py*P + ph, px*P + pw]),  patch = idx[i] = py*nw + px.
// One warp per (row, c, ph) line of P pixels when P*4B is 16B aligned.
// ------------------------------------------------------------------------------------------------
__global__ void im2col_gather_kernel(const float* __restrict__ img, const int32_t* __restrict__ idx,
                                     __nv_bfloat16* __restrict__ out, int64_t batch, int C, int H, int W, int P, int n_keep,
                                     int64_t ld_out) {
  const int nw = W / P;
  const int P4 = P >> 2;  // float4 per patch line
  const int64_t lines = batch * n_keep * C * P;
  const int64_t total = lines * P4;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
    const int q = (int)(t % P4);
    int64_t l = t / P4;
    const int ph = (int)(l % P); l /= P;
    const int c = (int)(l % C); l /= C;
    const int i = (int)(l % n_keep);
    const int64_t b = l / n_keep;
    const int patch = idx[i];
    const int py = patch / nw, px = patch % nw;
    const float4 v = *reinterpret_cast<const float4*>(img + (((b * C + c) * H + (py * P + ph)) * (int64_t)W + px * P + q * 4));
    __nv_bfloat16* o = out + (b * n_keep + i) * ld_out + (c * P + ph) * P + q * 4;
    *reinterpret_cast<uint2*>(o) = make_uint2(pack_bf16(v.x, v.y), pack_bf16(v.z, v.w));
  }
}

// Token-table im2col over all modalities (see mmf_im2col_tokens in the header): a row is written in 4-column (8-byte)
// chunks; a chunk inside the token's own modality block converts one float4 of the patch line, the indicator chunk holds
// the one-hot modality flag, every other chunk is zero.
struct TokIm2colParams {
  const float* img[4];
  int C[4], col_off[4], tok_off[5];
  int M, nenc, H, W, P, ind_col;
  const int32_t* tok;
  __nv_bfloat16* out;
  int64_t ld_out, batch;
};
__global__ void im2col_tokens_kernel(const TokIm2colParams p) {
  const int chunks = (int)(p.ld_out >> 2);
  const int nw = p.W / p.P, PP = p.P * p.P;
  const int64_t total = p.batch * p.nenc * chunks;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
    const int ch = (int)(t % chunks);
    const int64_t row = t / chunks;
    const int i = (int)(row % p.nenc);
    const int64_t b = row / p.nenc;
    const int id = p.tok[i];
    int m = 0;
    while (m + 1 < p.M && id >= p.tok_off[m + 1]) ++m;
    const int col = ch * 4;
    uint2 v = make_uint2(0u, 0u);
    const int rel = col - p.col_off[m];
    if (rel >= 0 && rel < p.C[m] * PP) {
      const int patch = id - p.tok_off[m];
      const int py = patch / nw, px = patch % nw;
      const int c = rel / PP, r2 = rel % PP;
      const int ph = r2 / p.P, pw = r2 % p.P;
      const float4 f = *reinterpret_cast<const float4*>(p.img[m] + (((b * p.C[m] + c) * p.H + (py * p.P + ph)) * (int64_t)p.W + px * p.P + pw));
      v = make_uint2(pack_bf16(f.x, f.y), pack_bf16(f.z, f.w));
    } else if (col == p.ind_col) {   // (ind_col is a multiple of 4 and M <= 4: the flags live in this one chunk)
      float o[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) o[q] = (col + q == p.ind_col + m) ? 1.0f : 0.0f;
      v = make_uint2(pack_bf16(o[0], o[1]), pack_bf16(o[2], o[3]));
    }
    *reinterpret_cast<uint2*>(p.out + row * p.ld_out + col) = v;
  }
}

// One-hot im2col of a class map (SemSegInputAdapter, input_adapters.py:209-328): row b*n_keep + i of `out`
// ([.., num_classes*P*P] bf16, ZEROED by the caller) gets a 1 at column cls*P*P + ph*P + pw for every pixel of visible
// patch idx[i].  With it the adapter's embedding lookup + Conv2d(k = s = P) becomes one GEMM against the
// [num_classes*P*P, D] table  T[c, ph, pw, :] = W[:, :, ph, pw] . class_emb[c]  (exact in bf16: the operand is 0 / 1).
__global__ void onehot_im2col_kernel(const int64_t* __restrict__ cls, const int32_t* __restrict__ idx, __nv_bfloat16* __restrict__ out,
                                     int64_t batch, int H, int W, int P, int n_keep, int num_classes, int64_t ld_out) {
  const int nw = W / P;
  const int64_t total = batch * n_keep * P * P;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
    const int pw = (int)(t % P);
    int64_t l = t / P;
    const int ph = (int)(l % P); l /= P;
    const int i = (int)(l % n_keep);
    const int64_t b = l / n_keep;
    const int patch = idx[i];
    const int py = patch / nw, px = patch % nw;
    const int64_t c = cls[(b * H + (py * P + ph)) * (int64_t)W + px * P + pw];
    if (c >= 0 && c < num_classes) out[(b * n_keep + i) * ld_out + (c * P + ph) * P + pw] = __float2bfloat16(1.0f);
  }
}

// ------------------------------------------------------------------------------------------------
// Un-patchify 'b (nh nw) (c ph pw) -> b c (nh ph) (nw pw)' (output_adapters_simple.py:183-186), bf16,
// and its inverse (for the gradient).  8-element (16 B) granules; requires P % 8 == 0.
// ------------------------------------------------------------------------------------------------
__global__ void unpatchify_kernel(const uint4* __restrict__ tok, uint4* __restrict__ img, int64_t batch, int C, int H,
                                  int W, int P, int inverse) {
  pdl_wait();   // launched through launch_pdl (common.cuh)
  const int nw = W / P, nh = H / P, P8 = P >> 3;
  const int64_t total = batch * C * H * (W >> 3);
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
    // t indexes the image in 8-pixel granules
    const int xg = (int)(t % (W >> 3));
    int64_t l = t / (W >> 3);
    const int y = (int)(l % H); l /= H;
    const int c = (int)(l % C);
    const int64_t b = l / C;
    const int px = xg / P8, q = xg % P8, py = y / P, ph = y % P;
    const int64_t tok_idx = ((b * nh * nw + py * nw + px) * (int64_t)(C * P * P) + (c * P + ph) * P) / 8 + q;
    if (!inverse) img[t] = tok[tok_idx];
    else          const_cast<uint4*>(tok)[tok_idx] = img[t];
  }
}

// ------------------------------------------------------------------------------------------------
// Row gather with cast: dst[b*n + i, :] = cast(src[b*src_batch_rows + (idx ? idx[i] : i) + row_off, :])
// src f32 or bf16; dst bf16 or f32.  d % 4 == 0.
// ------------------------------------------------------------------------------------------------
template <typename TS, typename TD>
__global__ void gather_rows_kernel(const TS* __restrict__ src, int64_t ld_src, int64_t src_batch_rows, int64_t row_off,
                                   const int32_t* __restrict__ idx, TD* __restrict__ dst, int64_t ld_dst, int64_t batch,
                                   int n, int d) {
  const int d4 = d >> 2;
  const int64_t total = batch * n * d4;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(t % d4);
    const int64_t r = t / d4;
    const int i = (int)(r % n);
    const int64_t b = r / n;
    const int64_t sr = b * src_batch_rows + row_off + (idx ? idx[i] : i);
    float v[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) v[k] = (float)src[sr * ld_src + c * 4 + k];
#pragma unroll
    for (int k = 0; k < 4; ++k) dst[r * ld_dst + c * 4 + k] = (TD)v[k];
  }
}

// y (f32) += x (f32), n % 4 == 0
__global__ void add_inplace_kernel(float4* __restrict__ y, const float4* __restrict__ x, int64_t n4) {
  pdl_wait();   // launched through launch_pdl (common.cuh)
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    float4 a = y[i];
    const float4 b = x[i];
    a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
    y[i] = a;
  }
}

// out (f32) = x (f32) + d (bf16), n % 4 == 0: materialises the residual stream after the last sub-layer
__global__ void add_bf16_kernel(float4* __restrict__ out, const float4* __restrict__ x, const uint2* __restrict__ d, int64_t n4) {
  pdl_wait();   // launched through launch_pdl (common.cuh)
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    float4 a = x[i];
    const uint2 u = d[i];
    const float2 lo = unpack_bf16(u.x), hi = unpack_bf16(u.y);
    a.x += lo.x; a.y += lo.y; a.z += hi.x; a.w += hi.y;
    out[i] = a;
  }
}

}  // namespace mmf

using namespace mmf;

extern "C" int mmf_cast_f32_bf16(const float* src, int64_t rows, int64_t cols, int64_t ld_src, void* dst, int64_t rows_pad,
                                 int64_t cols_pad, int64_t ld_dst, float scale, mmf_stream_t stream) {
  if (!src || !dst) MMF_BAD_ARG(1);
  if (rows_pad < rows || cols_pad < cols || ld_dst < cols_pad || ld_src < cols) MMF_BAD_ARG(2);
  if (rows_pad * cols_pad == 0) return 0;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const bool contiguous = rows == rows_pad && cols == cols_pad && ld_src == cols && ld_dst == cols && scale == 1.0f;
  const int64_t n = rows * cols;
  if (contiguous && (n & 3) == 0 && !(reinterpret_cast<uintptr_t>(src) & 15) && !(reinterpret_cast<uintptr_t>(dst) & 7)) {
    launch_pdl(cast_vec_kernel, dim3(ew_grid(n / 4, 256)), dim3(256), 0, st, reinterpret_cast<const float4*>(src), reinterpret_cast<uint2*>(dst), n / 4);
  } else {
    launch_pdl(cast_pad_kernel, dim3(ew_grid(rows_pad * cols_pad, 256)), dim3(256), 0, st, src, rows, cols, ld_src,
                                                                     reinterpret_cast<__nv_bfloat16*>(dst), rows_pad, cols_pad, ld_dst, scale);
  }
  MMF_COUNT_LAUNCH();
  MMF_LAUNCH_CHECK();
  return 0;
}

extern "C" int mmf_geglu_bwd(const void* u, const void* dg, void* du, int64_t rows, int64_t ipad, mmf_stream_t stream) {
  if (!u || !dg || !du) MMF_BAD_ARG(1);
  if (ipad & 7) MMF_BAD_ARG(2);
  if (rows * ipad == 0) return 0;
  geglu_bwd_kernel<<<ew_grid(rows * ipad / 8, 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const uint4*>(u), reinterpret_cast<const uint4*>(dg), reinterpret_cast<uint4*>(du), rows, (int)(ipad / 8));
  MMF_COUNT_LAUNCH();
  MMF_LAUNCH_CHECK();
  return 0;
}

extern "C" int mmf_gelu_bwd(const void* pre, const void* dy, void* dpre, int64_t n, mmf_stream_t stream) {
  if (!pre || !dy || !dpre) MMF_BAD_ARG(1);
  if (n & 7) MMF_BAD_ARG(2);
  if (n == 0) return 0;
  launch_pdl(gelu_bwd_kernel, dim3(ew_grid(n / 8, 256)), dim3(256), 0, reinterpret_cast<cudaStream_t>(stream), 
      reinterpret_cast<const uint4*>(pre), reinterpret_cast<const uint4*>(dy), reinterpret_cast<uint4*>(dpre), n / 8);
  MMF_COUNT_LAUNCH();
  MMF_LAUNCH_CHECK();
  return 0;
}

extern "C" int mmf_colsum(const void* x, int32_t x_f32, int64_t rows, int64_t cols, int64_t ld, float* out, mmf_stream_t stream) {
  if (!x || !out) MMF_BAD_ARG(1);
  if (rows * cols == 0) return 0;
  if (!x_f32 && cols % 256 == 0 && ld % 8 == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0) {
    const int gxv = (int)(cols / 256);
    int gyv = (int)ceil_div64(148 * 4, gxv);
    if (gyv > rows / 32) gyv = (int)(rows / 32 > 0 ? rows / 32 : 1);
    const int rpc = (int)ceil_div64(rows, gyv);
    gyv = (int)ceil_div64(rows, rpc);
    launch_pdl(colsum_vec_kernel, dim3(dim3(gxv, gyv)), dim3(256), 0, reinterpret_cast<cudaStream_t>(stream), reinterpret_cast<const __nv_bfloat16*>(x), rows, ld, out, rpc);
    MMF_COUNT_LAUNCH();
    MMF_LAUNCH_CHECK();
    return 0;
  }
  const int threads = 128;
  const int gx = (int)ceil_div64(cols, threads);
  int gy = (int)ceil_div64(148 * 8, gx);
  if (gy > rows) gy = (int)rows;
  if (gy < 1) gy = 1;
  const int rows_per_cta = (int)ceil_div64(rows, gy);
  gy = (int)ceil_div64(rows, rows_per_cta);
  dim3 grid(gx, gy);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (x_f32) colsum_kernel<float><<<grid, threads, 0, st>>>(reinterpret_cast<const float*>(x), rows, (int)cols, ld, out, rows_per_cta);
  else colsum_kernel<__nv_bfloat16><<<grid, threads, 0, st>>>(reinterpret_cast<const __nv_bfloat16*>(x), rows, (int)cols, ld, out, rows_per_cta);
  MMF_COUNT_LAUNCH();
  MMF_LAUNCH_CHECK();
  return 0;
}

extern "C" int mmf_bcast_rows(const float* src, float* dst, int64_t batch, int64_t rows, int64_t d, int64_t dst_batch_stride,
                              mmf_stream_t stream) {
  if (!src || !dst) MMF_BAD_ARG(1);
  if ((d & 3) || (dst_batch_stride & 3)) MMF_BAD_ARG(2);
  if (batch * rows * d == 0) return 0;
  bcast_rows_kernel<<<ew_grid(batch * rows * d / 4, 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const float4*>(src), dst, batch, rows, (int)(d / 4), dst_batch_stride);
  MMF_COUNT_LAUNCH();
  MMF_LAUNCH_CHECK();
  return 0;
}

extern "C" int mmf_reduce_batch(const float* src, float* dst, int64_t batch, int64_t rows, int64_t d, int64_t src_batch_stride,
                                mmf_stream_t stream) {
  if (!src || !dst) MMF_BAD_ARG(1);
  if ((d & 3) || (src_batch_stride & 3)) MMF_BAD_ARG(2);
  if (rows * d == 0) return 0;
  reduce_batch_kernel<<<ew_grid(rows * d / 4, 128), 128, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      src, reinterpret_cast<float4*>(dst), batch, rows, (int)(d / 4), src_batch_stride);
  MMF_COUNT_LAUNCH();
  MMF_LAUNCH_CHECK();
  return 0;
}

extern "C" int mmf_im2col_gather(const float* img, const int32_t* idx, void* out, int64_t batch, int32_t C, int32_t H, int32_t W,
                                 int32_t P, int32_t n_keep, int64_t ld_out, mmf_stream_t stream) {
  if (!img || !idx || !out) MMF_BAD_ARG(1);
  if (P <= 0 || (P & 3) || H % P || W % P || (W & 3) || (ld_out & 3) || ld_out < (int64_t)C * P * P) MMF_BAD_ARG(2);
  if (batch * n_keep == 0) return 0;
  const int64_t total = batch * n_keep * C * P * (P / 4);
  im2col_gather_kernel<<<ew_grid(total, 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      img, idx, reinterpret_cast<__nv_bfloat16*>(out), batch, C, H, W, P, n_keep, ld_out);
  MMF_COUNT_LAUNCH();
  MMF_LAUNCH_CHECK();
  return 0;
}

extern "C" int mmf_im2col_tokens(const float* const* imgs, const int32_t* chans, const int32_t* col_off, const int32_t* tok_off,
                                 int32_t M, const int32_t* tok, int32_t nenc, void* out, int64_t ld_out, int32_t ind_col, int64_t batch,
                                 int32_t H, int32_t W, int32_t P, mmf_stream_t stream) {
  if (!imgs || !chans || !col_off || !tok_off || !tok || !out) MMF_BAD_ARG(1);
  if (M <= 0 || M > 4 || P <= 0 || (P & 3) || H % P || W % P || (W & 3) || (ld_out & 3) || ind_col < 0 || (ind_col & 3) || ind_col + 4 > ld_out) MMF_BAD_ARG(2);
  if (batch * nenc == 0) return 0;
  TokIm2colParams p;
  for (int m = 0; m < 4; ++m) { p.img[m] = nullptr; p.C[m] = 0; p.col_off[m] = 0; }
  for (int m = 0; m < M; ++m) {
    if (!imgs[m] || chans[m] <= 0 || (col_off[m] & 3) || col_off[m] + chans[m] * P * P > ind_col) MMF_BAD_ARG(3);
    p.img[m] = imgs[m]; p.C[m] = chans[m]; p.col_off[m] = col_off[m];
  }
  for (int m = 0; m <= M; ++m) p.tok_off[m] = tok_off[m];
  for (int m = M + 1; m < 5; ++m) p.tok_off[m] = tok_off[M];
  p.M = M; p.nenc = nenc; p.H = H; p.W = W; p.P = P; p.ind_col = ind_col; p.tok = tok;
  p.out = reinterpret_cast<__nv_bfloat16*>(out); p.ld_out = ld_out; p.batch = batch;
  const int64_t total = batch * nenc * (ld_out >> 2);
  im2col_tokens_kernel<<<ew_grid(total, 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(p);
  MMF_COUNT_LAUNCH();
  MMF_LAUNCH_CHECK();
  return 0;
}

extern "C" int mmf_onehot_im2col(const int64_t* cls, const int32_t* idx, void* out, int64_t batch, int32_t H, int32_t W, int32_t P,
                                 int32_t n_keep, int32_t num_classes, int64_t ld_out, mmf_stream_t stream) {
  if (!cls || !idx || !out) MMF_BAD_ARG(1);
  if (P <= 0 || H % P || W % P || num_classes <= 0 || ld_out < (int64_t)num_classes * P * P) MMF_BAD_ARG(2);
  if (batch * n_keep == 0) return 0;
  onehot_im2col_kernel<<<ew_grid(batch * n_keep * P * P, 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      cls, idx, reinterpret_cast<__nv_bfloat16*>(out), batch, H, W, P, n_keep, num_classes, ld_out);
  MMF_COUNT_LAUNCH();
  MMF_LAUNCH_CHECK();
  return 0;
}

extern "C" int mmf_unpatchify_bf16(void* tokens, void* image, int64_t batch, int32_t C, int32_t H, int32_t W, int32_t P,
                                   int32_t inverse, mmf_stream_t stream) {
  if (!tokens || !image) MMF_BAD_ARG(1);
  if (P <= 0 || (P & 7) || H % P || W % P) MMF_BAD_ARG(2);
  const int64_t total = batch * C * H * (W / 8);
  if (total == 0) return 0;
  launch_pdl(unpatchify_kernel, dim3(ew_grid(total, 256)), dim3(256), 0, reinterpret_cast<cudaStream_t>(stream), 
      reinterpret_cast<const uint4*>(tokens), reinterpret_cast<uint4*>(image), batch, C, H, W, P, inverse);
  MMF_COUNT_LAUNCH();
  MMF_LAUNCH_CHECK();
  return 0;
}

extern "C" int mmf_gather_rows(const void* src, int32_t src_f32, int64_t ld_src, int64_t src_batch_rows, int64_t row_off,
                               const int32_t* idx, void* dst, int32_t dst_f32, int64_t ld_dst, int64_t batch, int32_t n,
                               int32_t d, mmf_stream_t stream) {
  if (!src || !dst) MMF_BAD_ARG(1);
  if (d & 3) MMF_BAD_ARG(2);
  const int64_t total = batch * n * (d / 4);
  if (total == 0) return 0;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int g = ew_grid(total, 256);
#define MMF_GR(TS, TD)                                                                                                  \
  gather_rows_kernel<TS, TD><<<g, 256, 0, st>>>(reinterpret_cast<const TS*>(src), ld_src, src_batch_rows, row_off, idx, \
                                                reinterpret_cast<TD*>(dst), ld_dst, batch, n, d)
  if (src_f32 && dst_f32) MMF_GR(float, float);
  else if (src_f32 && !dst_f32) MMF_GR(float, __nv_bfloat16);
  else if (!src_f32 && dst_f32) MMF_GR(__nv_bfloat16, float);
  else MMF_GR(__nv_bfloat16, __nv_bfloat16);
#undef MMF_GR
  MMF_COUNT_LAUNCH();
  MMF_LAUNCH_CHECK();
  return 0;
}

extern "C" int mmf_add_inplace_f32(float* y, const float* x, int64_t n, mmf_stream_t stream) {
  if (!y || !x) MMF_BAD_ARG(1);
  if (n & 3) MMF_BAD_ARG(2);
  if (n == 0) return 0;
  launch_pdl(add_inplace_kernel, dim3(ew_grid(n / 4, 256)), dim3(256), 0, reinterpret_cast<cudaStream_t>(stream), 
      reinterpret_cast<float4*>(y), reinterpret_cast<const float4*>(x), n / 4);
  MMF_COUNT_LAUNCH();
  MMF_LAUNCH_CHECK();
  return 0;
}

extern "C" int mmf_add_bf16_f32(float* out, const float* x, const void* d, int64_t n, mmf_stream_t stream) {
  if (!out || !x || !d) MMF_BAD_ARG(1);
  if (n & 3) MMF_BAD_ARG(2);
  if (n == 0) return 0;
  launch_pdl(add_bf16_kernel, dim3(ew_grid(n / 4, 256)), dim3(256), 0, reinterpret_cast<cudaStream_t>(stream), 
      reinterpret_cast<float4*>(out), reinterpret_cast<const float4*>(x), reinterpret_cast<const uint2*>(d), n / 4);
  MMF_COUNT_LAUNCH();
  MMF_LAUNCH_CHECK();
  return 0;
}
