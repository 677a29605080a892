// LayerNorm kernels (HBM-bound, warp-per-row, 128-bit accesses, fp32 math).
//
// The reference's encoder applies LayerNorm twice back to back before every attention and every
// FFN (Block.norm1 -> Attention.norm, Block.norm2 -> FeedForward[0]: zorro_utils.py:238->176,
// 239->124), each a separate ATen launch reading and writing the fp32 [rows, D] stream.  Here the
// pair is one pass: read fp32 x once, write the bf16 GEMM operand once.  The same kernel does the
// single nn.LayerNorm(eps=1e-6, bias) of the decoders (multimae_utils.py:217-232) and the final
// encoder norm (multimae.py:431).  The backward fuses both LN backwards, the residual-gradient add
// and the bf16 copy the next dgrad GEMM needs.
#include "common.cuh"
#include "mmf_b200.h"
#include <atomic>
#include <cstdlib>

namespace mmf {
extern std::atomic<int64_t> g_launch_count;

constexpr int LN_MAX_D = 1024;  // up to 8 float4 chunks per lane
constexpr int LN_WARPS = 8;

struct LnParams {
  const float* x;
  const float* x2;   // optional second source for rows >= x_split
  int64_t x_split;
  int64_t rows, ldx;
  int D;
  const float* g1;
  const float* b1;
  const float* g2;
  float eps1, eps2;
  void* y;
  int64_t ldy;
  int y_f32;
  float* stats;  // [rows, 4] mean1, rstd1, mean2, rstd2 (nullable)
  // fused residual add: rows >= delta_row0 are normalised as x + delta[row - delta_row0] (delta bf16: the output of the
  // sub-layer's last GEMM) and that sum, the new fp32 residual stream, is written to xout[row - delta_row0]
  const __nv_bfloat16* delta;
  int64_t delta_row0, lddelta;
  float* xout;
  int64_t ldxout;
  int l2_prefetch;  // L2-prefetch x (and delta) this many row steps ahead, 0 = off (MMF_LN_FWD_L2PF)
};

// gamma1 / bias1 / gamma2 staged once per CTA in shared memory (index = float4 chunk)
__device__ __forceinline__ void ln_stage_params(float4* sg1, float4* sb1, float4* sg2, const float* g1, const float* b1,
                                                const float* g2, int nchunk) {
  for (int c = threadIdx.x; c < nchunk; c += blockDim.x) {
    sg1[c] = __ldg(reinterpret_cast<const float4*>(g1) + c);
    sb1[c] = b1 ? __ldg(reinterpret_cast<const float4*>(b1) + c) : make_float4(0.f, 0.f, 0.f, 0.f);
    sg2[c] = g2 ? __ldg(reinterpret_cast<const float4*>(g2) + c) : make_float4(1.f, 1.f, 1.f, 1.f);
  }
  __syncthreads();
}

// FULL: D == NC * 128, every chunk of every lane is in range.  The `chunk < nchunk` tests then fold away at compile time;
// with them each chunk is its own predicated basic block and the row's arithmetic cannot be interleaved.
template <int NC, bool FULL>
__global__ void __launch_bounds__(LN_WARPS * 32, (NC <= 6 ? 3 : 2)) ln_fwd_kernel(const LnParams p) {
  constexpr int LN_MAX_CHUNKS = NC;
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int nchunk = p.D >> 2;
  const float invD = 1.0f / (float)p.D;
  __shared__ float4 g1[NC * 32], b1[NC * 32], g2[NC * 32];
  pdl_wait();   // launched with the programmatic-serialization attribute (launch_pdl): nothing is read before this
  ln_stage_params(g1, b1, g2, p.g1, p.b1, p.g2, nchunk);
  for (int64_t row = (int64_t)blockIdx.x * LN_WARPS + warp; row < p.rows; row += (int64_t)gridDim.x * LN_WARPS) {
    const float4* xr = reinterpret_cast<const float4*>(
        (p.x2 && row >= p.x_split) ? p.x2 + (row - p.x_split) * p.ldx : p.x + row * p.ldx);
    float4 v[LN_MAX_CHUNKS];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < LN_MAX_CHUNKS; ++i) {
      const int c = lane + 32 * i;
      v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (FULL || c < nchunk) v[i] = xr[c];
    }
    if (p.l2_prefetch) {  // same reasoning as the backward kernel: one row of loads per warp is too few bytes in flight
      const int64_t r2 = row + (int64_t)p.l2_prefetch * gridDim.x * LN_WARPS;
      if (r2 < p.rows) {
        const float* xr2 = (p.x2 && r2 >= p.x_split) ? p.x2 + (r2 - p.x_split) * p.ldx : p.x + r2 * p.ldx;
        const bool pd = p.delta && r2 >= p.delta_row0;
#pragma unroll
        for (int i = 0; i < LN_MAX_CHUNKS; ++i) {
          const int c = lane + 32 * i;
          if (FULL || c < nchunk) {
            if ((lane & 7) == 0) asm volatile("prefetch.global.L2 [%0];" ::"l"(xr2 + 4 * c));
            if (pd && (lane & 15) == 0) asm volatile("prefetch.global.L2 [%0];" ::"l"(p.delta + (r2 - p.delta_row0) * p.lddelta + 4 * c));
          }
        }
      }
    }
    if (p.delta && row >= p.delta_row0) {
      const uint2* dr = reinterpret_cast<const uint2*>(p.delta + (row - p.delta_row0) * p.lddelta);
      float4* xo = reinterpret_cast<float4*>(p.xout + (row - p.delta_row0) * p.ldxout);
      uint2 u[LN_MAX_CHUNKS];
#pragma unroll
      for (int i = 0; i < LN_MAX_CHUNKS; ++i)
        if (FULL || lane + 32 * i < nchunk) u[i] = dr[lane + 32 * i];
#pragma unroll
      for (int i = 0; i < LN_MAX_CHUNKS; ++i) {
        if (FULL || lane + 32 * i < nchunk) {
          const float2 lo = unpack_bf16(u[i].x), hi = unpack_bf16(u[i].y);
          v[i].x += lo.x; v[i].y += lo.y; v[i].z += hi.x; v[i].w += hi.y;
          xo[lane + 32 * i] = v[i];
        }
      }
    }
#pragma unroll
    for (int i = 0; i < LN_MAX_CHUNKS; ++i) s += v[i].x + v[i].y + v[i].z + v[i].w;
    const float mean1 = warp_sum(s) * invD;
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < LN_MAX_CHUNKS; ++i) {
      const int c = lane + 32 * i;
      if (FULL || c < nchunk) {
        const float a = v[i].x - mean1, b = v[i].y - mean1, cc = v[i].z - mean1, d = v[i].w - mean1;
        q += a * a + b * b + cc * cc + d * d;
      }
    }
    const float rstd1 = rsqrtf(warp_sum(q) * invD + p.eps1);
    float mean2 = 0.f, rstd2 = 1.f;
#pragma unroll
    for (int i = 0; i < LN_MAX_CHUNKS; ++i) {
      v[i].x = (v[i].x - mean1) * rstd1 * g1[lane + 32 * i].x + b1[lane + 32 * i].x;
      v[i].y = (v[i].y - mean1) * rstd1 * g1[lane + 32 * i].y + b1[lane + 32 * i].y;
      v[i].z = (v[i].z - mean1) * rstd1 * g1[lane + 32 * i].z + b1[lane + 32 * i].z;
      v[i].w = (v[i].w - mean1) * rstd1 * g1[lane + 32 * i].w + b1[lane + 32 * i].w;
    }
    if (p.g2) {
      s = 0.f;
#pragma unroll
      for (int i = 0; i < LN_MAX_CHUNKS; ++i)
        if (FULL || lane + 32 * i < nchunk) s += v[i].x + v[i].y + v[i].z + v[i].w;
      mean2 = warp_sum(s) * invD;
      q = 0.f;
#pragma unroll
      for (int i = 0; i < LN_MAX_CHUNKS; ++i) {
        if (FULL || lane + 32 * i < nchunk) {
          const float a = v[i].x - mean2, b = v[i].y - mean2, cc = v[i].z - mean2, d = v[i].w - mean2;
          q += a * a + b * b + cc * cc + d * d;
        }
      }
      rstd2 = rsqrtf(warp_sum(q) * invD + p.eps2);
#pragma unroll
      for (int i = 0; i < LN_MAX_CHUNKS; ++i) {
        v[i].x = (v[i].x - mean2) * rstd2 * g2[lane + 32 * i].x;
        v[i].y = (v[i].y - mean2) * rstd2 * g2[lane + 32 * i].y;
        v[i].z = (v[i].z - mean2) * rstd2 * g2[lane + 32 * i].z;
        v[i].w = (v[i].w - mean2) * rstd2 * g2[lane + 32 * i].w;
      }
    }
    if (p.y_f32) {
      float4* yr = reinterpret_cast<float4*>(reinterpret_cast<float*>(p.y) + row * p.ldy);
#pragma unroll
      for (int i = 0; i < LN_MAX_CHUNKS; ++i)
        if (FULL || lane + 32 * i < nchunk) yr[lane + 32 * i] = v[i];
    } else {
      uint2* yr = reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(p.y) + row * p.ldy);
#pragma unroll
      for (int i = 0; i < LN_MAX_CHUNKS; ++i)
        if (FULL || lane + 32 * i < nchunk) yr[lane + 32 * i] = make_uint2(pack_bf16(v[i].x, v[i].y), pack_bf16(v[i].z, v[i].w));
    }
    if (p.stats && lane == 0) reinterpret_cast<float4*>(p.stats)[row] = make_float4(mean1, rstd1, mean2, rstd2);
  }
}

struct LnBwdParams {
  const void* dy;  // bf16 or f32 [rows, D]
  int64_t lddy;
  int dy_f32;
  const float* x;
  const float* x2;
  int64_t x_split;
  int64_t rows, ldx;
  int D;
  const float* g1;
  const float* b1;
  const float* g2;
  const float* stats;
  const float* dres;  // optional f32 [rows, D] added to dx (the residual branch's gradient)
  int64_t lddres;
  float* dx;          // f32 [rows, D]
  int64_t lddx;
  void* dx_bf16;      // optional bf16 copy of dx
  int64_t lddxb;
  float* dg1;         // f32 [D] accumulated with atomics (must be zeroed by the caller)
  float* db1;         // optional
  float* dg2;         // optional (double LN)
  int l2_prefetch;    // issue L2 prefetches this many row steps ahead, 0 = off (MMF_LN_BWD_L2PF, default 2)
};

// PF (software pipelining across rows): ncu shows the plain row loop stalled on its own loads (long-scoreboard 9.7 per
// issue at 24 warps per SM; no pipe above 50 %): a warp has nothing in flight while it does a row's ~1300 instructions.
// With PF the NEXT row's x and dy are requested before the current row's arithmetic starts (NC float4 + NC uint2 more
// live registers: 2 CTAs per SM instead of 3), and the residual-branch gradient is fetched in one batch ahead of the
// first LayerNorm's reductions instead of chunk by chunk behind the dx stores.  Tried and dropped: two rows per warp
// (spills, 0.42 ms) and packed fp32x2 row arithmetic (the pair packing moves cost more than the FFMA2s save: 0.41 ms
// against 0.33 ms for this form at the cfg-2 shape).
// HB: the first LayerNorm has a bias (decoders' nn.LayerNorm); the zorro LayerNorms have none (beta is a zero buffer,
// zorro_utils.py:103-110): without it the bias reads and adds of the second norm's recomputation drop out.
template <int NC, bool PF, bool FULL, bool HB>
__global__ void __launch_bounds__(LN_WARPS * 32, (PF ? 2 : (NC <= 6 ? 3 : 2))) ln_bwd_kernel(const LnBwdParams p) {
  constexpr int LN_MAX_CHUNKS = NC;
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int nchunk = p.D >> 2;
  const float invD = 1.0f / (float)p.D;
  const bool dbl = p.g2 != nullptr;
  __shared__ float4 g1[NC * 32], b1[NC * 32], g2[NC * 32];
  ln_stage_params(g1, b1, g2, p.g1, p.b1, p.g2, nchunk);
  // Parameter-gradient accumulators live in shared memory, one private slice per warp (lane-contiguous float4s:
  // conflict free, no synchronisation).  The kernel is latency bound (ncu: long-scoreboard stalls, 48 % of DRAM
  // peak at 16 resident warps), so registers are what is budgeted: only the normalised input and the running
  // gradient stay live per row (2 x NC float4); the second LayerNorm's xhat is recomputed from them and the
  // residual-branch gradient is read where it is added.  That fits 3 CTAs (24 warps) per SM for D <= 768.
  extern __shared__ float4 ln_acc[];
  float4* adg1 = ln_acc + (size_t)warp * NC * 32;
  float4* adg2 = adg1 + (size_t)LN_WARPS * NC * 32;
  float4* adb1 = adg1 + (size_t)(dbl ? 2 : 1) * LN_WARPS * NC * 32;
#pragma unroll
  for (int i = 0; i < LN_MAX_CHUNKS; ++i) {
    adg1[lane + 32 * i] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (dbl) adg2[lane + 32 * i] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (p.db1) adb1[lane + 32 * i] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  const int64_t row_first = (int64_t)blockIdx.x * LN_WARPS + warp, row_step = (int64_t)gridDim.x * LN_WARPS;
  const bool pf_dy = PF && !p.dy_f32;   // bf16 upstream gradients ride the prefetch; the rare fp32 ones are read in place
  float4 nx[PF ? LN_MAX_CHUNKS : 1];
  uint2 ndy[PF ? LN_MAX_CHUNKS : 1];
  float4 nst = make_float4(0.f, 1.f, 0.f, 1.f);
  auto fetch_row = [&](int64_t r) {
    const float4* xr = reinterpret_cast<const float4*>((p.x2 && r >= p.x_split) ? p.x2 + (r - p.x_split) * p.ldx : p.x + r * p.ldx);
    const uint2* dr = reinterpret_cast<const uint2*>(reinterpret_cast<const __nv_bfloat16*>(p.dy) + r * p.lddy);
    nst = __ldg(reinterpret_cast<const float4*>(p.stats) + r);
#pragma unroll
    for (int i = 0; i < LN_MAX_CHUNKS; ++i) {
      const int c = lane + 32 * i;
      if (FULL || c < nchunk) {
        nx[PF ? i : 0] = xr[c];
        if (pf_dy) ndy[PF ? i : 0] = dr[c];
      }
    }
  };
  if (PF && row_first < p.rows) fetch_row(row_first);
  for (int64_t row = row_first; row < p.rows; row += row_step) {
    const float4 st = PF ? nst : __ldg(reinterpret_cast<const float4*>(p.stats) + row);
    const float mean1 = st.x, rstd1 = st.y, mean2 = st.z, rstd2 = st.w;
    const float nm1 = -mean1 * rstd1, nm2 = -mean2 * rstd2;   // xhat = fma(x, rstd, -mean * rstd): one instruction per value
    const float4* xr = reinterpret_cast<const float4*>(
        (p.x2 && row >= p.x_split) ? p.x2 + (row - p.x_split) * p.ldx : p.x + row * p.ldx);
    float4 xh1[LN_MAX_CHUNKS], d[LN_MAX_CHUNKS], rs[PF ? LN_MAX_CHUNKS : 1];
#pragma unroll
    for (int i = 0; i < LN_MAX_CHUNKS; ++i) {  // all global loads of the row are issued before any arithmetic
      const int c = lane + 32 * i;
      xh1[i] = d[i] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (FULL || c < nchunk) {
        xh1[i] = PF ? nx[PF ? i : 0] : xr[c];
        if (p.dy_f32) {
          d[i] = reinterpret_cast<const float4*>(reinterpret_cast<const float*>(p.dy) + row * p.lddy)[c];
        } else {
          const uint2 u = PF ? ndy[PF ? i : 0] : reinterpret_cast<const uint2*>(reinterpret_cast<const __nv_bfloat16*>(p.dy) + row * p.lddy)[c];
          const float2 lo = unpack_bf16(u.x), hi = unpack_bf16(u.y);
          d[i] = make_float4(lo.x, lo.y, hi.x, hi.y);
        }
      }
    }
    if (PF && row + row_step < p.rows) fetch_row(row + row_step);   // in flight during this row's arithmetic
    if (PF && p.l2_prefetch) {
      // L2 prefetch two rows ahead (x, dy) and of the next row's residual-branch gradient: a warp has only ONE row of loads
      // in flight for ~1/4 of the time a row takes it (ncu: DRAM 58 %, issue 30 %: latency-bound, too few bytes in flight);
      // with the lines already in L2 the register prefetch above completes in a third of the time
      const int64_t r2 = row + p.l2_prefetch * row_step, r1 = row + (p.l2_prefetch - 1) * row_step;
      if (r2 < p.rows) {
        const float* xr2 = (p.x2 && r2 >= p.x_split) ? p.x2 + (r2 - p.x_split) * p.ldx : p.x + r2 * p.ldx;
        const __nv_bfloat16* dr2 = reinterpret_cast<const __nv_bfloat16*>(p.dy) + r2 * p.lddy;
#pragma unroll
        for (int i = 0; i < LN_MAX_CHUNKS; ++i) {
          const int c = lane + 32 * i;
          if (FULL || c < nchunk) {
            if ((lane & 7) == 0) asm volatile("prefetch.global.L2 [%0];" ::"l"(xr2 + 4 * c));
            if (pf_dy && (lane & 15) == 0) asm volatile("prefetch.global.L2 [%0];" ::"l"(dr2 + 4 * c));
          }
        }
      }
      if (p.dres && r1 < p.rows) {
#pragma unroll
        for (int i = 0; i < LN_MAX_CHUNKS; ++i) {
          const int c = lane + 32 * i;
          if ((FULL || c < nchunk) && (lane & 7) == 0) asm volatile("prefetch.global.L2 [%0];" ::"l"(p.dres + r1 * p.lddres + 4 * c));
        }
      }
    }
#pragma unroll
    for (int i = 0; i < LN_MAX_CHUNKS; ++i) {
      const bool in = FULL || lane + 32 * i < nchunk;
      xh1[i].x = in ? fmaf(xh1[i].x, rstd1, nm1) : 0.f; xh1[i].y = in ? fmaf(xh1[i].y, rstd1, nm1) : 0.f;
      xh1[i].z = in ? fmaf(xh1[i].z, rstd1, nm1) : 0.f; xh1[i].w = in ? fmaf(xh1[i].w, rstd1, nm1) : 0.f;
    }
    if (dbl) {
      // second LN: y = xh2 * g2, xh2 = (y1 - mean2) * rstd2, y1 = xh1 * g1 + b1
      float s1 = 0.f, s2 = 0.f;
#pragma unroll
      for (int i = 0; i < LN_MAX_CHUNKS; ++i) {
        if (FULL || lane + 32 * i < nchunk) {
          const float4 G1 = g1[lane + 32 * i], G2 = g2[lane + 32 * i];
          const float4 B1 = HB ? b1[lane + 32 * i] : make_float4(0.f, 0.f, 0.f, 0.f);
          float4 xh2;
          if (HB) {
            xh2.x = (xh1[i].x * G1.x + B1.x - mean2) * rstd2; xh2.y = (xh1[i].y * G1.y + B1.y - mean2) * rstd2;
            xh2.z = (xh1[i].z * G1.z + B1.z - mean2) * rstd2; xh2.w = (xh1[i].w * G1.w + B1.w - mean2) * rstd2;
          } else {
            xh2.x = fmaf(xh1[i].x * G1.x, rstd2, nm2); xh2.y = fmaf(xh1[i].y * G1.y, rstd2, nm2);
            xh2.z = fmaf(xh1[i].z * G1.z, rstd2, nm2); xh2.w = fmaf(xh1[i].w * G1.w, rstd2, nm2);
          }
          float4 t = adg2[lane + 32 * i];
          t.x += d[i].x * xh2.x; t.y += d[i].y * xh2.y; t.z += d[i].z * xh2.z; t.w += d[i].w * xh2.w;
          adg2[lane + 32 * i] = t;
          d[i].x *= G2.x; d[i].y *= G2.y; d[i].z *= G2.z; d[i].w *= G2.w;
          s1 += d[i].x + d[i].y + d[i].z + d[i].w;
          s2 += d[i].x * xh2.x + d[i].y * xh2.y + d[i].z * xh2.z + d[i].w * xh2.w;
        }
      }
      s1 = warp_sum(s1) * invD;
      s2 = warp_sum(s2) * invD;
      const float c2a = -rstd2 * s1, c2b = -rstd2 * s2;
#pragma unroll
      for (int i = 0; i < LN_MAX_CHUNKS; ++i) {
        if (FULL || lane + 32 * i < nchunk) {
          const float4 G1 = g1[lane + 32 * i];
          if (HB) {
            const float4 B1 = b1[lane + 32 * i];
            d[i].x = rstd2 * (d[i].x - s1 - (xh1[i].x * G1.x + B1.x - mean2) * rstd2 * s2);
            d[i].y = rstd2 * (d[i].y - s1 - (xh1[i].y * G1.y + B1.y - mean2) * rstd2 * s2);
            d[i].z = rstd2 * (d[i].z - s1 - (xh1[i].z * G1.z + B1.z - mean2) * rstd2 * s2);
            d[i].w = rstd2 * (d[i].w - s1 - (xh1[i].w * G1.w + B1.w - mean2) * rstd2 * s2);
          } else {
            // rstd2 * (d - s1 - xh2 * s2) as two dependent FMAs on top of xh2's two
            d[i].x = fmaf(d[i].x, rstd2, fmaf(fmaf(xh1[i].x * G1.x, rstd2, nm2), c2b, c2a));
            d[i].y = fmaf(d[i].y, rstd2, fmaf(fmaf(xh1[i].y * G1.y, rstd2, nm2), c2b, c2a));
            d[i].z = fmaf(d[i].z, rstd2, fmaf(fmaf(xh1[i].z * G1.z, rstd2, nm2), c2b, c2a));
            d[i].w = fmaf(d[i].w, rstd2, fmaf(fmaf(xh1[i].w * G1.w, rstd2, nm2), c2b, c2a));
          }
        }
      }
    }
    // first LN
    if (PF) {
#pragma unroll
      for (int i = 0; i < LN_MAX_CHUNKS; ++i) {
        rs[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (p.dres && (FULL || lane + 32 * i < nchunk)) rs[i] = reinterpret_cast<const float4*>(p.dres + row * p.lddres)[lane + 32 * i];
      }
    }
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < LN_MAX_CHUNKS; ++i) {
      if (FULL || lane + 32 * i < nchunk) {
        float4 t = adg1[lane + 32 * i];
        t.x += d[i].x * xh1[i].x; t.y += d[i].y * xh1[i].y; t.z += d[i].z * xh1[i].z; t.w += d[i].w * xh1[i].w;
        adg1[lane + 32 * i] = t;
        if (p.db1) {
          float4 u = adb1[lane + 32 * i];
          u.x += d[i].x; u.y += d[i].y; u.z += d[i].z; u.w += d[i].w;
          adb1[lane + 32 * i] = u;
        }
        const float4 G1 = g1[lane + 32 * i];
        d[i].x *= G1.x; d[i].y *= G1.y; d[i].z *= G1.z; d[i].w *= G1.w;
        s1 += d[i].x + d[i].y + d[i].z + d[i].w;
        s2 += d[i].x * xh1[i].x + d[i].y * xh1[i].y + d[i].z * xh1[i].z + d[i].w * xh1[i].w;
      }
    }
    s1 = warp_sum(s1) * invD;
    s2 = warp_sum(s2) * invD;
    const float c1a = -rstd1 * s1, c1b = -rstd1 * s2;   // rstd1 * (d - s1 - xh1 * s2) = fma(d, rstd1, fma(xh1, c1b, c1a))
#pragma unroll
    for (int i = 0; i < LN_MAX_CHUNKS; ++i) {
      const int c = lane + 32 * i;
      if (FULL || c < nchunk) {
        float4 o;
        o.x = fmaf(d[i].x, rstd1, fmaf(xh1[i].x, c1b, c1a));
        o.y = fmaf(d[i].y, rstd1, fmaf(xh1[i].y, c1b, c1a));
        o.z = fmaf(d[i].z, rstd1, fmaf(xh1[i].z, c1b, c1a));
        o.w = fmaf(d[i].w, rstd1, fmaf(xh1[i].w, c1b, c1a));
        if (PF) {
          o.x += rs[PF ? i : 0].x; o.y += rs[PF ? i : 0].y; o.z += rs[PF ? i : 0].z; o.w += rs[PF ? i : 0].w;
        } else if (p.dres) {
          const float4 r = reinterpret_cast<const float4*>(p.dres + row * p.lddres)[c];
          o.x += r.x; o.y += r.y; o.z += r.z; o.w += r.w;
        }
        reinterpret_cast<float4*>(p.dx + row * p.lddx)[c] = o;
        if (p.dx_bf16)
          reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(p.dx_bf16) + row * p.lddxb)[c] =
              make_uint2(pack_bf16(o.x, o.y), pack_bf16(o.z, o.w));
      }
    }
  }
  // parameter gradients: sum the warps' slices, then one atomic per column
  __syncthreads();
  const int narr = 1 + (dbl ? 1 : 0) + (p.db1 ? 1 : 0);
  for (int a = 0; a < narr; ++a) {
    float* dst = a == 0 ? p.dg1 : ((a == 1 && dbl) ? p.dg2 : p.db1);
    const float4* base = ln_acc + (size_t)a * LN_WARPS * NC * 32;
    for (int c = threadIdx.x; c < nchunk; c += blockDim.x) {
      float4 t = base[c];
#pragma unroll
      for (int w = 1; w < LN_WARPS; ++w) {
        const float4 u = base[(size_t)w * NC * 32 + c];
        t.x += u.x; t.y += u.y; t.z += u.z; t.w += u.w;
      }
      atomicAdd(dst + 4 * c, t.x); atomicAdd(dst + 4 * c + 1, t.y);
      atomicAdd(dst + 4 * c + 2, t.z); atomicAdd(dst + 4 * c + 3, t.w);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Ring variant of the backward kernel (D = NC * 128, bf16 upstream gradient, 16-byte aligned rows): the rows travel
// global -> shared memory as 1-D bulk copies (cp.async.bulk + mbarrier complete_tx), issued by each warp's lane 0 into the
// warp's PRIVATE two-deep ring two rows ahead of the row being worked on, so nothing in the row loop waits on HBM and no
// register holds prefetched data.  Why: the register-prefetch kernel above is latency bound (ncu: 16 warps / SM, one
// instruction per ~11 clocks per warp, 5.6 of them long-scoreboard); inside a training step the SM clock sits at the power
// cap (~1.46 GHz against 1.9 GHz for the kernel run alone) and its rate falls with the clock (0.87 of the HBM peak alone,
// 0.69 in the step).  Here ~160 KB per SM are in flight whatever the clock.  One CTA per SM, 12 warps (10 at D = 1024),
// up to 168 registers: the parameter-gradient accumulators move from shared memory into registers, which also removes
// four LDS/STS per chunk.  (x, dy, stats) and the residual-branch gradient have separate barriers: the first group is
// copied to registers and re-armed at once, the second is read where it is added at the end of the row.
// ------------------------------------------------------------------------------------------------
template <int NC>
struct LnRing {
  static constexpr int WARPS = NC <= 6 ? 12 : 10;
  static constexpr int XB = NC * 512, DYB = NC * 256;
  static constexpr int STAGE_A = XB + DYB + 16;   // x row, dy row, the row's (mean1, rstd1, mean2, rstd2)
  static constexpr int STAGE_R = XB;              // dres row
  static constexpr int WARP_BYTES = 2 * STAGE_A + 2 * STAGE_R;
  static constexpr int PARAM_BYTES = 3 * NC * 512;
  static constexpr int BAR_BYTES = ((WARPS * 4 * 8 + 127) / 128) * 128;
  static constexpr int SMEM = PARAM_BYTES + BAR_BYTES + WARPS * WARP_BYTES;
  static_assert(WARP_BYTES >= 3 * NC * 512, "the ring doubles as the staging area of the final parameter-gradient reduction");
};

__device__ __forceinline__ void bulk_load_1d(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(smem_dst)), "l"(gsrc), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

template <int NC, bool HB>
__global__ void __launch_bounds__(LnRing<NC>::WARPS * 32, 1) ln_bwd_ring_kernel(const LnBwdParams p) {
  using R = LnRing<NC>;
  constexpr int W = R::WARPS;
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const float invD = 1.0f / (float)p.D;
  const bool dbl = p.g2 != nullptr;
  extern __shared__ __align__(128) uint8_t ring_smem[];
  float4* g1 = reinterpret_cast<float4*>(ring_smem);
  float4* b1 = g1 + NC * 32;
  float4* g2 = b1 + NC * 32;
  uint64_t* bar_a = reinterpret_cast<uint64_t*>(ring_smem + R::PARAM_BYTES) + warp * 4;
  uint64_t* bar_r = bar_a + 2;
  uint8_t* wbase = ring_smem + R::PARAM_BYTES + R::BAR_BYTES + (size_t)warp * R::WARP_BYTES;
  if (lane == 0) {
    for (int s = 0; s < 4; ++s) mbar_init(&bar_a[s], 1);
    mbar_fence_init();
  }
  pdl_wait();   // launched through launch_pdl: the barrier set-up above overlaps the previous kernel's tail
  ln_stage_params(g1, b1, g2, p.g1, p.b1, p.g2, NC * 32);   // ends with __syncthreads()
  const int64_t row_first = (int64_t)blockIdx.x * W + warp, row_step = (int64_t)gridDim.x * W;
  const __nv_bfloat16* dyp = reinterpret_cast<const __nv_bfloat16*>(p.dy);
  auto issue_a = [&](int64_t r, int s) {
    uint8_t* sa = wbase + s * R::STAGE_A;
    mbar_expect_tx(&bar_a[s], R::STAGE_A);
    bulk_load_1d(sa, (p.x2 && r >= p.x_split) ? p.x2 + (r - p.x_split) * p.ldx : p.x + r * p.ldx, R::XB, &bar_a[s]);
    bulk_load_1d(sa + R::XB, dyp + r * p.lddy, R::DYB, &bar_a[s]);
    bulk_load_1d(sa + R::XB + R::DYB, p.stats + 4 * r, 16, &bar_a[s]);
  };
  auto issue_r = [&](int64_t r, int s) {
    mbar_expect_tx(&bar_r[s], R::STAGE_R);
    bulk_load_1d(wbase + 2 * R::STAGE_A + s * R::STAGE_R, p.dres + r * p.lddres, R::STAGE_R, &bar_r[s]);
  };
  if (lane == 0) {
#pragma unroll
    for (int s = 0; s < 2; ++s) {
      const int64_t r = row_first + s * row_step;
      if (r < p.rows) {
        issue_a(r, s);
        if (p.dres) issue_r(r, s);
      }
    }
  }
  float4 adg1[NC], adg2[NC], adb1[HB ? NC : 1];
#pragma unroll
  for (int i = 0; i < NC; ++i) {
    adg1[i] = adg2[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (HB) adb1[HB ? i : 0] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  int s = 0;
  uint32_t ph = 0;
  for (int64_t row = row_first; row < p.rows; row += row_step) {
    const uint8_t* sa = wbase + s * R::STAGE_A;
    mbar_wait(&bar_a[s], ph);
    float4 xh1[NC], d[NC];
#pragma unroll
    for (int i = 0; i < NC; ++i) {
      xh1[i] = reinterpret_cast<const float4*>(sa)[lane + 32 * i];
      const uint2 u = reinterpret_cast<const uint2*>(sa + R::XB)[lane + 32 * i];
      const float2 lo = unpack_bf16(u.x), hi = unpack_bf16(u.y);
      d[i] = make_float4(lo.x, lo.y, hi.x, hi.y);
    }
    const float4 st = *reinterpret_cast<const float4*>(sa + R::XB + R::DYB);
    __syncwarp();   // every lane's reads of the stage are done: lane 0 re-arms it for the row after next
    if (lane == 0 && row + 2 * row_step < p.rows) {
      fence_proxy_async_smem();
      issue_a(row + 2 * row_step, s);
    }
    const float rstd1 = st.y, mean2 = st.z, rstd2 = st.w;
    const float nm1 = -st.x * rstd1, nm2 = -mean2 * rstd2;
#pragma unroll
    for (int i = 0; i < NC; ++i) {
      xh1[i].x = fmaf(xh1[i].x, rstd1, nm1); xh1[i].y = fmaf(xh1[i].y, rstd1, nm1);
      xh1[i].z = fmaf(xh1[i].z, rstd1, nm1); xh1[i].w = fmaf(xh1[i].w, rstd1, nm1);
    }
    if (dbl) {
      float s1 = 0.f, s2 = 0.f;
#pragma unroll
      for (int i = 0; i < NC; ++i) {
        const float4 G1 = g1[lane + 32 * i], G2 = g2[lane + 32 * i];
        float4 xh2;
        if (HB) {
          const float4 B1 = b1[lane + 32 * i];
          xh2.x = (xh1[i].x * G1.x + B1.x - mean2) * rstd2; xh2.y = (xh1[i].y * G1.y + B1.y - mean2) * rstd2;
          xh2.z = (xh1[i].z * G1.z + B1.z - mean2) * rstd2; xh2.w = (xh1[i].w * G1.w + B1.w - mean2) * rstd2;
        } else {
          xh2.x = fmaf(xh1[i].x * G1.x, rstd2, nm2); xh2.y = fmaf(xh1[i].y * G1.y, rstd2, nm2);
          xh2.z = fmaf(xh1[i].z * G1.z, rstd2, nm2); xh2.w = fmaf(xh1[i].w * G1.w, rstd2, nm2);
        }
        adg2[i].x += d[i].x * xh2.x; adg2[i].y += d[i].y * xh2.y; adg2[i].z += d[i].z * xh2.z; adg2[i].w += d[i].w * xh2.w;
        d[i].x *= G2.x; d[i].y *= G2.y; d[i].z *= G2.z; d[i].w *= G2.w;
        s1 += d[i].x + d[i].y + d[i].z + d[i].w;
        s2 += d[i].x * xh2.x + d[i].y * xh2.y + d[i].z * xh2.z + d[i].w * xh2.w;
      }
      s1 = warp_sum(s1) * invD;
      s2 = warp_sum(s2) * invD;
      const float c2a = -rstd2 * s1, c2b = -rstd2 * s2;
#pragma unroll
      for (int i = 0; i < NC; ++i) {
        const float4 G1 = g1[lane + 32 * i];
        if (HB) {
          const float4 B1 = b1[lane + 32 * i];
          d[i].x = rstd2 * (d[i].x - s1 - (xh1[i].x * G1.x + B1.x - mean2) * rstd2 * s2);
          d[i].y = rstd2 * (d[i].y - s1 - (xh1[i].y * G1.y + B1.y - mean2) * rstd2 * s2);
          d[i].z = rstd2 * (d[i].z - s1 - (xh1[i].z * G1.z + B1.z - mean2) * rstd2 * s2);
          d[i].w = rstd2 * (d[i].w - s1 - (xh1[i].w * G1.w + B1.w - mean2) * rstd2 * s2);
        } else {
          d[i].x = fmaf(d[i].x, rstd2, fmaf(fmaf(xh1[i].x * G1.x, rstd2, nm2), c2b, c2a));
          d[i].y = fmaf(d[i].y, rstd2, fmaf(fmaf(xh1[i].y * G1.y, rstd2, nm2), c2b, c2a));
          d[i].z = fmaf(d[i].z, rstd2, fmaf(fmaf(xh1[i].z * G1.z, rstd2, nm2), c2b, c2a));
          d[i].w = fmaf(d[i].w, rstd2, fmaf(fmaf(xh1[i].w * G1.w, rstd2, nm2), c2b, c2a));
        }
      }
    }
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < NC; ++i) {
      adg1[i].x += d[i].x * xh1[i].x; adg1[i].y += d[i].y * xh1[i].y; adg1[i].z += d[i].z * xh1[i].z; adg1[i].w += d[i].w * xh1[i].w;
      if (HB) {
        adb1[HB ? i : 0].x += d[i].x; adb1[HB ? i : 0].y += d[i].y; adb1[HB ? i : 0].z += d[i].z; adb1[HB ? i : 0].w += d[i].w;
      }
      const float4 G1 = g1[lane + 32 * i];
      d[i].x *= G1.x; d[i].y *= G1.y; d[i].z *= G1.z; d[i].w *= G1.w;
      s1 += d[i].x + d[i].y + d[i].z + d[i].w;
      s2 += d[i].x * xh1[i].x + d[i].y * xh1[i].y + d[i].z * xh1[i].z + d[i].w * xh1[i].w;
    }
    s1 = warp_sum(s1) * invD;
    s2 = warp_sum(s2) * invD;
    const float c1a = -rstd1 * s1, c1b = -rstd1 * s2;
    const float4* sr = reinterpret_cast<const float4*>(wbase + 2 * R::STAGE_A + s * R::STAGE_R);
    if (p.dres) mbar_wait(&bar_r[s], ph);
    float* dxr = p.dx + row * p.lddx;
    __nv_bfloat16* dxb = p.dx_bf16 ? reinterpret_cast<__nv_bfloat16*>(p.dx_bf16) + row * p.lddxb : nullptr;
#pragma unroll
    for (int i = 0; i < NC; ++i) {
      const int c = lane + 32 * i;
      float4 o;
      o.x = fmaf(d[i].x, rstd1, fmaf(xh1[i].x, c1b, c1a));
      o.y = fmaf(d[i].y, rstd1, fmaf(xh1[i].y, c1b, c1a));
      o.z = fmaf(d[i].z, rstd1, fmaf(xh1[i].z, c1b, c1a));
      o.w = fmaf(d[i].w, rstd1, fmaf(xh1[i].w, c1b, c1a));
      if (p.dres) {
        const float4 r = sr[c];
        o.x += r.x; o.y += r.y; o.z += r.z; o.w += r.w;
      }
      reinterpret_cast<float4*>(dxr)[c] = o;
      if (dxb) reinterpret_cast<uint2*>(dxb)[c] = make_uint2(pack_bf16(o.x, o.y), pack_bf16(o.z, o.w));
    }
    if (p.dres) {
      __syncwarp();
      if (lane == 0 && row + 2 * row_step < p.rows) {
        fence_proxy_async_smem();
        issue_r(row + 2 * row_step, s);
      }
    }
    s ^= 1;
    if (s == 0) ph ^= 1;
  }
  // parameter gradients: every copy this warp issued has been waited for, so its ring is free: park the accumulators there,
  // sum the warps' slices, then one atomic per column
  __syncwarp();
  {
    float4* acc = reinterpret_cast<float4*>(wbase);
#pragma unroll
    for (int i = 0; i < NC; ++i) {
      acc[lane + 32 * i] = adg1[i];
      acc[NC * 32 + lane + 32 * i] = adg2[i];
      if (HB) acc[2 * NC * 32 + lane + 32 * i] = adb1[HB ? i : 0];
    }
  }
  __syncthreads();
  for (int a = 0; a < 3; ++a) {
    float* dst = a == 0 ? p.dg1 : (a == 1 ? (dbl ? p.dg2 : nullptr) : (HB ? p.db1 : nullptr));
    if (!dst) continue;
    const uint8_t* base = ring_smem + R::PARAM_BYTES + R::BAR_BYTES + (size_t)a * NC * 512;
    for (int c = threadIdx.x; c < NC * 32; c += blockDim.x) {
      float4 t = reinterpret_cast<const float4*>(base)[c];
#pragma unroll
      for (int w = 1; w < W; ++w) {
        const float4 u = reinterpret_cast<const float4*>(base + (size_t)w * R::WARP_BYTES)[c];
        t.x += u.x; t.y += u.y; t.z += u.z; t.w += u.w;
      }
      atomicAdd(dst + 4 * c, t.x); atomicAdd(dst + 4 * c + 1, t.y);
      atomicAdd(dst + 4 * c + 2, t.z); atomicAdd(dst + 4 * c + 3, t.w);
    }
  }
}

// Persistent grid = exactly the CTAs that are co-resident (occupancy x SMs): the row loop is grid-strided, so a grid
// that is not a whole number of resident waves leaves SMs idle during the last wave.
template <typename Kern>
static int ln_grid(Kern kern, size_t dyn_smem, int64_t rows, int waves = 1) {
  int dev = 0, sms = 148, per_sm = 1;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, LN_WARPS * 32, dyn_smem) != cudaSuccess || per_sm < 1) per_sm = 1;
  const int64_t need = ceil_div64(rows, LN_WARPS);
  const int64_t cap = (int64_t)sms * per_sm * waves;
  return (int)(need < cap ? need : cap);
}

}  // namespace mmf

extern "C" int mmf_layernorm_fwd(const float* x, const float* x2, int64_t x_split, int64_t rows, int32_t D, int64_t ldx, const float* g1, const float* b1,
                                 float eps1, const float* g2, float eps2, void* y, int64_t ldy, int32_t y_f32,
                                 float* stats, const void* delta, int64_t delta_row0, int64_t lddelta, float* xout,
                                 int64_t ldxout, mmf_stream_t stream) {
  using namespace mmf;
  if (!x || !g1 || !y) MMF_BAD_ARG(1);
  if (rows <= 0) return 0;
  if (D <= 0 || (D & 3) || D > LN_MAX_D) MMF_BAD_ARG(2);
  if ((ldx & 3) || (ldy & 3) || ldx < D || ldy < D) MMF_BAD_ARG(3);
  if ((reinterpret_cast<uintptr_t>(x) & 15) || (reinterpret_cast<uintptr_t>(y) & 15)) MMF_BAD_ARG(4);
  if (delta && (!xout || delta_row0 < 0 || (lddelta & 3) || (ldxout & 3) || lddelta < D || ldxout < D ||
                (reinterpret_cast<uintptr_t>(delta) & 7) || (reinterpret_cast<uintptr_t>(xout) & 15)))
    MMF_BAD_ARG(5);
  LnParams p{x, x2, x_split, rows, ldx, D, g1, b1, g2, eps1, eps2, y, ldy, y_f32, stats,
             reinterpret_cast<const __nv_bfloat16*>(delta), delta_row0, lddelta, xout, ldxout, 0};
  const char* l2e = getenv("MMF_LN_FWD_L2PF");      // read per call (A/B runs switch it between launches)
  // isolated, cfg-2 shape: the plain kernel gains 7 % one row step ahead (5.73 -> 6.15 TB/s), the residual-add variant is at
  // 6.1 TB/s without it and loses with it
  p.l2_prefetch = l2e ? atoi(l2e) : (delta ? 0 : 1);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int nc = ceil_div(D, 128);
#define MMF_LN_FWD_LAUNCH(NCV, F) launch_pdl(ln_fwd_kernel<NCV, F>, dim3(ln_grid(ln_fwd_kernel<NCV, F>, 0, rows, 4)), dim3(LN_WARPS * 32), 0, st, p)
  if (nc <= 2) { if (D == 256) MMF_LN_FWD_LAUNCH(2, true); else MMF_LN_FWD_LAUNCH(2, false); }
  else if (nc <= 4) { if (D == 512) MMF_LN_FWD_LAUNCH(4, true); else MMF_LN_FWD_LAUNCH(4, false); }
  else if (nc <= 6) { if (D == 768) MMF_LN_FWD_LAUNCH(6, true); else MMF_LN_FWD_LAUNCH(6, false); }
  else { if (D == 1024) MMF_LN_FWD_LAUNCH(8, true); else MMF_LN_FWD_LAUNCH(8, false); }
#undef MMF_LN_FWD_LAUNCH
  g_launch_count.fetch_add(1, std::memory_order_relaxed);
  MMF_LAUNCH_CHECK();
  return 0;
}

extern "C" int mmf_layernorm_bwd(const void* dy, int64_t lddy, int32_t dy_f32, const float* x, const float* x2,
                                 int64_t x_split, int64_t rows, int32_t D, int64_t ldx, const float* g1, const float* b1, const float* g2, const float* stats,
                                 const float* dres, int64_t lddres, float* dx, int64_t lddx, void* dx_bf16,
                                 int64_t lddxb, float* dg1, float* db1, float* dg2, mmf_stream_t stream) {
  using namespace mmf;
  if (!dy || !x || !g1 || !stats || !dx || !dg1) MMF_BAD_ARG(1);
  if (rows <= 0) return 0;
  if (D <= 0 || (D & 3) || D > LN_MAX_D) MMF_BAD_ARG(2);
  if ((ldx & 3) || (lddy & 3) || (lddx & 3) || (dres && (lddres & 3)) || (dx_bf16 && (lddxb & 3))) MMF_BAD_ARG(3);
  if (g2 && !dg2) MMF_BAD_ARG(4);
  const char* l2e = getenv("MMF_LN_BWD_L2PF");      // read per call (A/B runs switch it between launches)
  LnBwdParams p{dy, lddy, dy_f32, x, x2, x_split, rows, ldx, D, g1, b1, g2, stats, dres, lddres, dx, lddx, dx_bf16, lddxb, dg1, db1,
                g2 ? dg2 : nullptr, l2e ? atoi(l2e) : 2};
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int nc = ceil_div(D, 128);
  const int ncp = nc <= 2 ? 2 : (nc <= 4 ? 4 : (nc <= 6 ? 6 : 8));
  const int narr = 1 + (g2 ? 1 : 0) + (db1 ? 1 : 0);
  const size_t smem = (size_t)narr * LN_WARPS * ncp * 32 * sizeof(float4);
  // ~58 KB per CTA at D = 768: ask for the full shared-memory carve-out so that three CTAs fit (the default carve-out
  // stops at two and shared memory, not registers, would set the occupancy)
  // row prefetch: 0.382 -> 0.329 ms at the cfg-2 shape (D = 768); at D = 1024 its registers spill, so it stays off there
  // ring variant (bulk copies into per-warp shared-memory rings): MMF_LN_BWD_RING=0 falls back to the register-prefetch kernel
  const char* ring_e = getenv("MMF_LN_BWD_RING");   // read per call (A/B runs switch it between launches)
  const bool ring_ok = (ring_e ? atoi(ring_e) : 1) != 0 && D == ncp * 128 && !dy_f32 && (!b1 || db1) && !(lddy & 7) &&
                       !(reinterpret_cast<uintptr_t>(dy) & 15) && !(reinterpret_cast<uintptr_t>(x) & 15) &&
                       !(x2 && (reinterpret_cast<uintptr_t>(x2) & 15)) && !(reinterpret_cast<uintptr_t>(stats) & 15) &&
                       !(dres && (reinterpret_cast<uintptr_t>(dres) & 15)) && !(reinterpret_cast<uintptr_t>(dx) & 15) &&
                       !(dx_bf16 && (reinterpret_cast<uintptr_t>(dx_bf16) & 7));
  if (ring_ok) {
#define MMF_LN_RING_LAUNCH(NCV, HBV)                                                                                          \
  do {                                                                                                                      \
    static DeviceOnce attr_done;                                                                                            \
    const int attr_dev = current_device();                                                                                  \
    if (!attr_done.done(attr_dev)) {                                                                                        \
      cudaFuncSetAttribute(ln_bwd_ring_kernel<NCV, HBV>, cudaFuncAttributeMaxDynamicSharedMemorySize, LnRing<NCV>::SMEM);   \
      attr_done.set(attr_dev);                                                                                              \
    }                                                                                                                       \
    int sms = 148;                                                                                                          \
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, attr_dev);                                                 \
    const int64_t need = ceil_div64(rows, LnRing<NCV>::WARPS);                                                              \
    launch_pdl(ln_bwd_ring_kernel<NCV, HBV>, dim3((int)(need < sms ? need : sms)), dim3(LnRing<NCV>::WARPS * 32), LnRing<NCV>::SMEM, st, p); \
  } while (0)
    if (b1) {
      if (ncp == 2) MMF_LN_RING_LAUNCH(2, true); else if (ncp == 4) MMF_LN_RING_LAUNCH(4, true);
      else if (ncp == 6) MMF_LN_RING_LAUNCH(6, true); else MMF_LN_RING_LAUNCH(8, true);
    } else {
      if (ncp == 2) MMF_LN_RING_LAUNCH(2, false); else if (ncp == 4) MMF_LN_RING_LAUNCH(4, false);
      else if (ncp == 6) MMF_LN_RING_LAUNCH(6, false); else MMF_LN_RING_LAUNCH(8, false);
    }
#undef MMF_LN_RING_LAUNCH
    g_launch_count.fetch_add(1, std::memory_order_relaxed);
    MMF_LAUNCH_CHECK();
    return 0;
  }
  static const int variant = getenv("MMF_LN_BWD_PF") ? atoi(getenv("MMF_LN_BWD_PF")) : 1;
  const bool early = variant != 0 && nc <= 6;
#define MMF_LN_BWD_LAUNCH(NCV, E, F, HBV)                                                                                       \
  do {                                                                                                                    \
    static DeviceOnce attr_done;                                                                                          \
    const int attr_dev = current_device();                                                                                \
    if (!attr_done.done(attr_dev)) {                                                                                      \
      cudaFuncSetAttribute(ln_bwd_kernel<NCV, E, F, HBV>, cudaFuncAttributeMaxDynamicSharedMemorySize, 3 * LN_WARPS * NCV * 32 * 16); \
      cudaFuncSetAttribute(ln_bwd_kernel<NCV, E, F, HBV>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared); \
      attr_done.set(attr_dev);                                                                                            \
    }                                                                                                                     \
    ln_bwd_kernel<NCV, E, F, HBV><<<ln_grid(ln_bwd_kernel<NCV, E, F, HBV>, smem, rows), LN_WARPS * 32, smem, st>>>(p);               \
  } while (0)
#define MMF_LN_BWD_DISPATCH_B(E, F, HBV)                 \
  do {                                                  \
    if (nc <= 2) MMF_LN_BWD_LAUNCH(2, E, F, HBV);       \
    else if (nc <= 4) MMF_LN_BWD_LAUNCH(4, E, F, HBV);  \
    else if (nc <= 6) MMF_LN_BWD_LAUNCH(6, E, F, HBV);  \
    else MMF_LN_BWD_LAUNCH(8, E, F, HBV);               \
  } while (0)
#define MMF_LN_BWD_DISPATCH(E, F)                                                           \
  do {                                                                                      \
    if (b1) MMF_LN_BWD_DISPATCH_B(E, F, true); else MMF_LN_BWD_DISPATCH_B(E, F, false);    \
  } while (0)
  const bool full = D == 256 || D == 512 || D == 768 || D == 1024;   // = NC * 128 of the kernel chosen below
  if (early) { if (full) MMF_LN_BWD_DISPATCH(true, true); else MMF_LN_BWD_DISPATCH(true, false); }
  else { if (full) MMF_LN_BWD_DISPATCH(false, true); else MMF_LN_BWD_DISPATCH(false, false); }
#undef MMF_LN_BWD_DISPATCH
#undef MMF_LN_BWD_DISPATCH_B
#undef MMF_LN_BWD_LAUNCH
  g_launch_count.fetch_add(1, std::memory_order_relaxed);
  MMF_LAUNCH_CHECK();
  return 0;
}
