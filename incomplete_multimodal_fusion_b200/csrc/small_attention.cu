// Two small-attention kernels of the MultiMAE path (HBM/latency-bound, CUDA-core math):
//
//  * slot attention  -- the "modality attention" of Block_Fusion (reference: downstream/
//    instance_segmentation/modeling/multimae/zorro_utils.py:243-258, call site
//    pretraining/multimae/multimae_crossattn.py:450-470).  For every spatial position p the
//    reference scatters the visible tokens of each modality into a clone of `mask_embedding`,
//    stacks [s1(p), s2(p), dem(p), fusion(p)] and runs full self-attention over the S slots, then
//    keeps only the fusion slot.  Here nothing is scattered: the fusion token's query attends to S
//    key/value rows looked up through a slot map (visible token row, or the batch-invariant
//    mask-embedding row), and only the fusion slot's output is computed.
//
//  * pool attention  -- the learned return-token queries (multimae.py:434-455,
//    multimae_crossattn.py:529-543): a handful of queries over all N tokens with a dense boolean
//    mask, reproducing masked_fill(-finfo.max) exactly: a row with no allowed key becomes the
//    UNIFORM distribution over all keys (mode 0), or zero output for an empty context (mode 1).
#include "common.cuh"
#include "mmf_b200.h"

#include <atomic>

namespace mmf {
extern std::atomic<int64_t> g_launch_count;

constexpr int SLOT_MAX = 8;

struct SlotParams {
  const __nv_bfloat16* q;       // [B*F, H*64]
  const __nv_bfloat16* kv_tok;  // planar token rows, [.., 2*H*64] = [k | v]
  const __nv_bfloat16* kv_me;   // [F, 2*H*64] mask-embedding rows
  const int32_t* slotmap;       // [S-1, F] rank in the modality's visible list or -1
  const int32_t* seg;           // [S] start of each modality segment inside the head plane
  __nv_bfloat16* out;           // [B*F, H*64]
  float* probs;                 // [B*F, H, S]
  int64_t ldq, ldkv, ldme, ldo;
  int B, F, H, S, n_head;
  float scale;
  // backward
  const __nv_bfloat16* dout;    // [B*F, H*64]
  __nv_bfloat16* dq;            // [B*F, H*64]
  __nv_bfloat16* dkv_tok;       // planar token rows (every row written exactly once)
  float* dkv_me;                // [F, 2*H*64] f32, atomically accumulated (caller zeroes)
  int64_t lddout, lddq, lddkv, lddme;
  float* me_scratch;            // [B*F, H, S-1, 2] f32: (ds, w) of every modality slot (wide backward path)
};

__device__ __forceinline__ int64_t slot_row(const SlotParams& p, int b, int pos, int s, bool& is_me) {
  is_me = false;
  if (s == p.S - 1) return (int64_t)p.B * p.n_head + (int64_t)b * p.F + pos;  // the fusion token itself
  const int r = p.slotmap[s * p.F + pos];
  if (r < 0) { is_me = true; return pos; }
  return (int64_t)b * p.n_head + p.seg[s] + r;
}

// one warp per (b, pos, head); dh = 64 -> 2 elements per lane
template <bool BWD>
__global__ void slot_attn_kernel(const SlotParams p) {
  const int lane = threadIdx.x & 31;
  const int h = threadIdx.x >> 5;
  const int64_t bp = blockIdx.x;  // b * F + pos
  const int b = (int)(bp / p.F), pos = (int)(bp % p.F);
  const int HD = p.H * 64;
  const float2 q = unpack_bf16(*reinterpret_cast<const uint32_t*>(p.q + bp * p.ldq + h * 64 + 2 * lane));
  float2 k[SLOT_MAX], v[SLOT_MAX];
  float s[SLOT_MAX];
  int64_t rows[SLOT_MAX];
  bool me[SLOT_MAX];
  float mx = -INFINITY;
#pragma unroll
  for (int t = 0; t < SLOT_MAX; ++t) {
    if (t < p.S) {
      rows[t] = slot_row(p, b, pos, t, me[t]);
      const __nv_bfloat16* base = me[t] ? p.kv_me + rows[t] * p.ldme : p.kv_tok + rows[t] * p.ldkv;
      k[t] = unpack_bf16(*reinterpret_cast<const uint32_t*>(base + h * 64 + 2 * lane));
      v[t] = unpack_bf16(*reinterpret_cast<const uint32_t*>(base + HD + h * 64 + 2 * lane));
      s[t] = warp_sum(q.x * k[t].x + q.y * k[t].y) * p.scale;
      mx = fmaxf(mx, s[t]);
    }
  }
  float sum = 0.f;
#pragma unroll
  for (int t = 0; t < SLOT_MAX; ++t)
    if (t < p.S) { s[t] = __expf(s[t] - mx); sum += s[t]; }
  const float inv = 1.0f / sum;
  if (!BWD) {
    float2 o = make_float2(0.f, 0.f);
#pragma unroll
    for (int t = 0; t < SLOT_MAX; ++t)
      if (t < p.S) { const float w = s[t] * inv; o.x += w * v[t].x; o.y += w * v[t].y; }
    *reinterpret_cast<uint32_t*>(p.out + bp * p.ldo + h * 64 + 2 * lane) = pack_bf16(o.x, o.y);
    if (p.probs && lane < p.S) {
      float w = 0.f;
#pragma unroll
      for (int t = 0; t < SLOT_MAX; ++t) if (t == lane) w = s[t] * inv;
      p.probs[(bp * p.H + h) * p.S + lane] = w;
    }
  } else {
    const float2 d = unpack_bf16(*reinterpret_cast<const uint32_t*>(p.dout + bp * p.lddout + h * 64 + 2 * lane));
    float dp[SLOT_MAX], dot = 0.f;
#pragma unroll
    for (int t = 0; t < SLOT_MAX; ++t)
      if (t < p.S) { s[t] *= inv; dp[t] = warp_sum(d.x * v[t].x + d.y * v[t].y); dot += s[t] * dp[t]; }
    float2 dq = make_float2(0.f, 0.f);
#pragma unroll
    for (int t = 0; t < SLOT_MAX; ++t) {
      if (t < p.S) {
        const float ds = s[t] * (dp[t] - dot) * p.scale;
        dq.x += ds * k[t].x; dq.y += ds * k[t].y;
        const float2 dk = make_float2(ds * q.x, ds * q.y);
        const float2 dv = make_float2(s[t] * d.x, s[t] * d.y);
        if (me[t]) {
          float* base = p.dkv_me + rows[t] * p.lddme;
          atomicAdd(base + h * 64 + 2 * lane, dk.x); atomicAdd(base + h * 64 + 2 * lane + 1, dk.y);
          atomicAdd(base + HD + h * 64 + 2 * lane, dv.x); atomicAdd(base + HD + h * 64 + 2 * lane + 1, dv.y);
        } else {
          __nv_bfloat16* base = p.dkv_tok + rows[t] * p.lddkv;
          *reinterpret_cast<uint32_t*>(base + h * 64 + 2 * lane) = pack_bf16(dk.x, dk.y);
          *reinterpret_cast<uint32_t*>(base + HD + h * 64 + 2 * lane) = pack_bf16(dv.x, dv.y);
        }
      }
    }
    *reinterpret_cast<uint32_t*>(p.dq + bp * p.lddq + h * 64 + 2 * lane) = pack_bf16(dq.x, dq.y);
  }
}

// ---- wide variant (H <= 8, 16-byte aligned rows): ONE warp per (b, pos) covers all heads: lane = (head, quarter of the
// head's 64 dims), 32-byte loads per lane and 4-lane dot-product reductions, instead of one warp per head with 4-byte
// loads.  Backward: whether slot s of position p is a mask-embedding row does not depend on the sample (masks are
// shared by the batch), so the mask-embedding gradients are a reduction over the batch: this kernel stores the two
// scalars (ds, w) per (b, p, head, slot) and slot_me_reduce_kernel sums ds*q / w*dout over b -- no atomics.
__device__ __forceinline__ void ld16(const __nv_bfloat16* p, float (&f)[16]) {
  const uint4 a = *reinterpret_cast<const uint4*>(p), b = *reinterpret_cast<const uint4*>(p + 8);
  const uint32_t u[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
  for (int i = 0; i < 8; ++i) { const float2 t = unpack_bf16(u[i]); f[2 * i] = t.x; f[2 * i + 1] = t.y; }
}
__device__ __forceinline__ void st16(__nv_bfloat16* p, const float (&f)[16]) {
  *reinterpret_cast<uint4*>(p) = make_uint4(pack_bf16(f[0], f[1]), pack_bf16(f[2], f[3]), pack_bf16(f[4], f[5]), pack_bf16(f[6], f[7]));
  *reinterpret_cast<uint4*>(p + 8) = make_uint4(pack_bf16(f[8], f[9]), pack_bf16(f[10], f[11]), pack_bf16(f[12], f[13]), pack_bf16(f[14], f[15]));
}
__device__ __forceinline__ float quad_sum(float v) {
  v += __shfl_xor_sync(0xffffffffu, v, 1);
  v += __shfl_xor_sync(0xffffffffu, v, 2);
  return v;
}

constexpr int SLOTW_WARPS = 4;

struct Pk16 { uint4 a, b; };   // 16 bf16 values, kept packed in registers until used
__device__ __forceinline__ Pk16 ldpk(const __nv_bfloat16* p) {
  Pk16 r;
  r.a = *reinterpret_cast<const uint4*>(p);
  r.b = *reinterpret_cast<const uint4*>(p + 8);
  return r;
}
__device__ __forceinline__ void unpk(const Pk16& r, float (&f)[16]) {
  const uint32_t u[8] = {r.a.x, r.a.y, r.a.z, r.a.w, r.b.x, r.b.y, r.b.z, r.b.w};
#pragma unroll
  for (int i = 0; i < 8; ++i) { const float2 t = unpack_bf16(u[i]); f[2 * i] = t.x; f[2 * i + 1] = t.y; }
}
__device__ __forceinline__ float dot16(const Pk16& r, const float (&x)[16]) {
  float f[16];
  unpk(r, f);
  float d = 0.f;
#pragma unroll
  for (int i = 0; i < 16; ++i) d += f[i] * x[i];
  return d;
}
__device__ __forceinline__ void axpy16(float a, const Pk16& r, float (&y)[16]) {
  float f[16];
  unpk(r, f);
#pragma unroll
  for (int i = 0; i < 16; ++i) y[i] += a * f[i];
}

template <bool BWD, int S>
__global__ void __launch_bounds__(SLOTW_WARPS * 32) slot_attn_wide_kernel(const SlotParams p) {
  pdl_wait();   // launched through launch_pdl (common.cuh): nothing another kernel owns is touched before this
  const int lane = threadIdx.x & 31;
  const int64_t bp = (int64_t)blockIdx.x * SLOTW_WARPS + (threadIdx.x >> 5);   // b * F + pos
  if (bp >= (int64_t)p.B * p.F) return;
  const int b = (int)(bp / p.F), pos = (int)(bp % p.F);
  const int h = lane >> 2, sub = lane & 3;
  const bool act = h < p.H;
  const int HD = p.H * 64;
  const int col = (act ? h : 0) * 64 + sub * 16;
  float q[16];
  ld16(p.q + bp * p.ldq + col, q);
  Pk16 k[S], v[S];
  float sc[S];
  int64_t rows[S];
  bool me[S];
#pragma unroll
  for (int t = 0; t < S; ++t) {   // all row loads are issued before any arithmetic
    rows[t] = slot_row(p, b, pos, t, me[t]);
    const __nv_bfloat16* base = me[t] ? p.kv_me + rows[t] * p.ldme : p.kv_tok + rows[t] * p.ldkv;
    k[t] = ldpk(base + col);
    v[t] = ldpk(base + HD + col);
  }
  float d[16];
  if (BWD) ld16(p.dout + bp * p.lddout + col, d);
  float mx = -INFINITY;
#pragma unroll
  for (int t = 0; t < S; ++t) {
    sc[t] = quad_sum(dot16(k[t], q)) * p.scale;
    mx = fmaxf(mx, sc[t]);
  }
  float sum = 0.f;
#pragma unroll
  for (int t = 0; t < S; ++t) { sc[t] = __expf(sc[t] - mx); sum += sc[t]; }
  const float inv = 1.0f / sum;
  if (!BWD) {
    float o[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) o[i] = 0.f;
#pragma unroll
    for (int t = 0; t < S; ++t) axpy16(sc[t] * inv, v[t], o);
    if (act) st16(p.out + bp * p.ldo + col, o);
  } else {
    float dp[S], dot = 0.f;
#pragma unroll
    for (int t = 0; t < S; ++t) {
      sc[t] *= inv;
      dp[t] = quad_sum(dot16(v[t], d));
      dot += sc[t] * dp[t];
    }
    float dq[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) dq[i] = 0.f;
#pragma unroll
    for (int t = 0; t < S; ++t) {
      const float ds = sc[t] * (dp[t] - dot) * p.scale;
      axpy16(ds, k[t], dq);
      if (t < S - 1 && act && sub == 0)
        *reinterpret_cast<float2*>(p.me_scratch + ((bp * p.H + h) * (S - 1) + t) * 2) = make_float2(ds, sc[t]);
      if (!me[t] && act) {
        float dk[16], dv[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) { dk[i] = ds * q[i]; dv[i] = sc[t] * d[i]; }
        __nv_bfloat16* base = p.dkv_tok + rows[t] * p.lddkv;
        st16(base + col, dk);
        st16(base + HD + col, dv);
      }
    }
    if (act) st16(p.dq + bp * p.lddq + col, dq);
  }
}

template <bool BWD>
static void launch_slot_wide(const SlotParams& p, cudaStream_t st) {
  const int64_t n = (int64_t)p.B * p.F;
  const unsigned grid = (unsigned)((n + SLOTW_WARPS - 1) / SLOTW_WARPS);
  switch (p.S) {
    case 2: launch_pdl(slot_attn_wide_kernel<BWD, 2>, dim3(grid), dim3(SLOTW_WARPS * 32), 0, st, p); break;
    case 3: launch_pdl(slot_attn_wide_kernel<BWD, 3>, dim3(grid), dim3(SLOTW_WARPS * 32), 0, st, p); break;
    case 4: launch_pdl(slot_attn_wide_kernel<BWD, 4>, dim3(grid), dim3(SLOTW_WARPS * 32), 0, st, p); break;
    default: launch_pdl(slot_attn_wide_kernel<BWD, 5>, dim3(grid), dim3(SLOTW_WARPS * 32), 0, st, p); break;
  }
}

// dkv_me[pos, :] = sum over the batch of the mask-embedding slots' (ds * q | w * dout); one CTA of 8 warps per
// (pos, head): warp w takes the samples b = w, w + 8, ..; lane = 2 of the head's 64 dims; the warps' partial sums meet in
// shared memory.  Positions / slots that are real tokens contribute nothing (written as 0).
constexpr int MERED_WARPS = 8;
__global__ void __launch_bounds__(MERED_WARPS * 32) slot_me_reduce_kernel(const SlotParams p) {
  pdl_wait();   // launched through launch_pdl (common.cuh): nothing another kernel owns is touched before this
  __shared__ float4 part[MERED_WARPS][32];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int pos = blockIdx.x / p.H, h = blockIdx.x % p.H;
  const int HD = p.H * 64;
  const int nm = p.S - 1;
  bool is_me[SLOT_MAX];
  bool any = false;
#pragma unroll
  for (int t = 0; t < SLOT_MAX; ++t) { is_me[t] = t < nm && p.slotmap[t * p.F + pos] < 0; any |= is_me[t]; }
  float2 dk = make_float2(0.f, 0.f), dv = make_float2(0.f, 0.f);
  if (any) {
#pragma unroll 4
    for (int b = w; b < p.B; b += MERED_WARPS) {
      const int64_t bp = (int64_t)b * p.F + pos;
      const float2 q = unpack_bf16(*reinterpret_cast<const uint32_t*>(p.q + bp * p.ldq + h * 64 + 2 * lane));
      const float2 d = unpack_bf16(*reinterpret_cast<const uint32_t*>(p.dout + bp * p.lddout + h * 64 + 2 * lane));
      const float* sw = p.me_scratch + (bp * p.H + h) * nm * 2;
      float ds = 0.f, wt = 0.f;
#pragma unroll
      for (int t = 0; t < SLOT_MAX; ++t)
        if (is_me[t]) { ds += __ldg(sw + 2 * t); wt += __ldg(sw + 2 * t + 1); }
      dk.x += ds * q.x; dk.y += ds * q.y;
      dv.x += wt * d.x; dv.y += wt * d.y;
    }
  }
  part[w][lane] = make_float4(dk.x, dk.y, dv.x, dv.y);
  __syncthreads();
  if (w == 0) {
    float4 t = part[0][lane];
#pragma unroll
    for (int i = 1; i < MERED_WARPS; ++i) { const float4 u = part[i][lane]; t.x += u.x; t.y += u.y; t.z += u.z; t.w += u.w; }
    float* base = p.dkv_me + (int64_t)pos * p.lddme;
    *reinterpret_cast<float2*>(base + h * 64 + 2 * lane) = make_float2(t.x, t.y);
    *reinterpret_cast<float2*>(base + HD + h * 64 + 2 * lane) = make_float2(t.z, t.w);
  }
}

// ------------------------------------------------------------------------------------------------
// pool attention
// ------------------------------------------------------------------------------------------------
constexpr int POOL_MAX_N = 2048;

struct PoolParams {
  const __nv_bfloat16* q;    // [Bq, R, H*64] (q_bstride = 0 when batch-invariant)
  const __nv_bfloat16* kv;   // planar token rows [.., 2*H*64]
  const uint8_t* mask;       // [R, N] 1 = allowed
  const int32_t* mode;       // [R] all-masked row: 0 -> uniform over N, 1 -> zero output
  __nv_bfloat16* out;        // [B, R, H*64]
  float* stat;               // [B, R, H, 2] = (max, sum) ; sum < 0 flags "all masked"
  int64_t q_bstride, ldkv;
  int B, R, H, N, n_head, n_tail;
  float scale;
  // backward
  const __nv_bfloat16* dout; // [B, R, H*64]
  float* dq;                 // [Bq, R, H*64] f32, atomically accumulated (caller zeroes)
  int64_t dq_bstride;
  __nv_bfloat16* dkv;        // planar token rows [.., 2*H*64]; every row written
  int64_t lddkv;
};

__device__ __forceinline__ float dot64(const __nv_bfloat16* a_row, const float* q) {
  float acc = 0.f;
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    const uint4 u = *reinterpret_cast<const uint4*>(a_row + c * 8);
    const uint32_t* pu = &u.x;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float2 f = unpack_bf16(pu[k]);
      acc += f.x * q[c * 8 + 2 * k] + f.y * q[c * 8 + 2 * k + 1];
    }
  }
  return acc;
}

// one warp per (b, h, r)
__global__ void pool_attn_fwd_kernel(const PoolParams p) {
  extern __shared__ float sc_all[];
  const int lane = threadIdx.x & 31, r = threadIdx.x >> 5;
  const int h = blockIdx.x, b = blockIdx.y;
  float* sc = sc_all + (size_t)r * p.N;
  __shared__ float qs[16][64];
  const int HD = p.H * 64;
  {
    const float2 f = unpack_bf16(*reinterpret_cast<const uint32_t*>(p.q + b * p.q_bstride + (int64_t)r * HD + h * 64 + 2 * lane));
    qs[r][2 * lane] = f.x; qs[r][2 * lane + 1] = f.y;
  }
  __syncwarp();
  const int64_t head_rows = (int64_t)p.B * p.n_head;
  float mx = -INFINITY;
  bool any = false;
  for (int j = lane; j < p.N; j += 32) {
    const bool ok = p.mask[r * p.N + j] != 0;
    float s = -INFINITY;
    if (ok) {
      const int64_t row = j < p.n_head ? (int64_t)b * p.n_head + j : head_rows + (int64_t)b * p.n_tail + (j - p.n_head);
      s = dot64(p.kv + row * p.ldkv + h * 64, qs[r]) * p.scale;
      any = true;
    }
    sc[j] = s;
    mx = fmaxf(mx, s);
  }
  mx = warp_max(mx);
  any = __any_sync(0xffffffffu, any);
  const bool uniform = !any && p.mode[r] == 0;
  float sum = 0.f;
  for (int j = lane; j < p.N; j += 32) {
    const float e = any ? __expf(sc[j] - mx) : (uniform ? 1.f : 0.f);
    sc[j] = e;
    sum += e;
  }
  sum = warp_sum(sum);
  __syncwarp();
  // weighted sum of the values: lane-per-key like the score pass (each lane accumulates all 64 dims over its keys: N/32
  // independent 128-byte row reads instead of N dependent 4-byte ones), then a transposing reduction: after the
  // butterfly lane l holds dims 2l, 2l+1
  float2 o = make_float2(0.f, 0.f);
  if (sum > 0.f) {
    float acc[64];
#pragma unroll
    for (int d = 0; d < 64; ++d) acc[d] = 0.f;
    for (int j = lane; j < p.N; j += 32) {
      const float w = sc[j];
      if (w == 0.f) continue;
      const int64_t row = j < p.n_head ? (int64_t)b * p.n_head + j : head_rows + (int64_t)b * p.n_tail + (j - p.n_head);
      const __nv_bfloat16* vr = p.kv + row * p.ldkv + HD + h * 64;
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        const uint4 u = *reinterpret_cast<const uint4*>(vr + c * 8);
        const uint32_t* pu = &u.x;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const float2 f = unpack_bf16(pu[k]);
          acc[c * 8 + 2 * k] += w * f.x; acc[c * 8 + 2 * k + 1] += w * f.y;
        }
      }
    }
    // reduce-scatter over the warp: at step s (16, 8, 4, 2, 1) a lane keeps the half of its remaining dims selected
    // by bit s of its lane id and receives the partner's partial sums for that half
    float buf[32];
#pragma unroll
    for (int d = 0; d < 32; ++d) {
      const bool hi = (lane & 16) != 0;
      const float send = hi ? acc[d] : acc[d + 32], keep = hi ? acc[d + 32] : acc[d];
      buf[d] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
    }
#pragma unroll
    for (int s = 8, n = 16; s >= 1; s >>= 1, n >>= 1) {
#pragma unroll
      for (int d = 0; d < 16; ++d) {
        if (d < n) {
          const bool hi = (lane & s) != 0;
          const float send = hi ? buf[d] : buf[d + n], keep = hi ? buf[d + n] : buf[d];
          buf[d] = keep + __shfl_xor_sync(0xffffffffu, send, s);
        }
      }
    }
    // lane l now holds the two dims [32*b4 + 16*b3 + 8*b2 + 4*b1 + 2*b0 + {0, 1}] with b_i the bits of l: 2l, 2l + 1
    const float inv = 1.0f / sum;
    o.x = buf[0] * inv; o.y = buf[1] * inv;
  }
  *reinterpret_cast<uint32_t*>(p.out + ((int64_t)b * p.R + r) * HD + h * 64 + 2 * lane) = pack_bf16(o.x, o.y);
  if (lane == 0) {
    float* st = p.stat + (((int64_t)b * p.R + r) * p.H + h) * 2;
    st[0] = any ? mx : 0.f;
    st[1] = any ? sum : (uniform ? -1.f : 0.f);  // -1: uniform row, 0: empty row
  }
}

// backward: 4 lanes per key (16 of the 64 dims each), 64 keys per 256-thread block, looping over the
// R queries.  delta[b,r,h] = dout.out sits right after the (max,sum) pairs in `stat`.
__global__ void __launch_bounds__(256) pool_attn_bwd_kernel(const PoolParams p) {
  // dq_r = sum_j ds_rj k_j is taken in a second phase from shared copies of the block's 64 keys and its [R, 64] ds
  // values: thread (r, dim) runs over the keys (the first form reduced every (r, dim) product over the warp's keys with
  // three shuffles and a shared atomic: 340 shuffles per thread)
  __shared__ float qs[16][64], ds_[16][64], dss[16][64], ks[64][65], mxs[16], sums[16], dls[16];
  const int h = blockIdx.y, b = blockIdx.z;
  const int HD = p.H * 64;
  const int tid = threadIdx.x, sub = tid & 3;
  for (int t = tid; t < p.R * 64; t += blockDim.x) {
    const int r = t / 64, d = t % 64;
    qs[r][d] = __bfloat162float(p.q[b * p.q_bstride + (int64_t)r * HD + h * 64 + d]);
    ds_[r][d] = __bfloat162float(p.dout[((int64_t)b * p.R + r) * HD + h * 64 + d]);
  }
  if (tid < p.R) {
    const int64_t brh = ((int64_t)b * p.R + tid) * p.H + h;
    mxs[tid] = p.stat[brh * 2];
    sums[tid] = p.stat[brh * 2 + 1];
    dls[tid] = p.stat[(int64_t)p.B * p.R * p.H * 2 + brh];
  }
  __syncthreads();
  const int j = blockIdx.x * 64 + (tid >> 2);
  const bool valid = j < p.N;
  const int64_t head_rows = (int64_t)p.B * p.n_head;
  const int64_t row = !valid ? 0 : (j < p.n_head ? (int64_t)b * p.n_head + j : head_rows + (int64_t)b * p.n_tail + (j - p.n_head));
  float kf[16], vf[16], dk[16], dv[16];
#pragma unroll
  for (int d = 0; d < 16; ++d) kf[d] = vf[d] = dk[d] = dv[d] = 0.f;
  if (valid) {
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      const uint4 uk = *reinterpret_cast<const uint4*>(p.kv + row * p.ldkv + h * 64 + sub * 16 + c * 8);
      const uint4 uv = *reinterpret_cast<const uint4*>(p.kv + row * p.ldkv + HD + h * 64 + sub * 16 + c * 8);
      const uint32_t* pk = &uk.x; const uint32_t* pv = &uv.x;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float2 a = unpack_bf16(pk[k]), c2 = unpack_bf16(pv[k]);
        kf[c * 8 + 2 * k] = a.x; kf[c * 8 + 2 * k + 1] = a.y;
        vf[c * 8 + 2 * k] = c2.x; vf[c * 8 + 2 * k + 1] = c2.y;
      }
    }
  }
#pragma unroll
  for (int d = 0; d < 16; ++d) ks[tid >> 2][sub * 16 + d] = kf[d];
  for (int r = 0; r < p.R; ++r) {
    const float sum = sums[r];
    float dp = 0.f, sdot = 0.f;
#pragma unroll
    for (int d = 0; d < 16; ++d) { dp += ds_[r][sub * 16 + d] * vf[d]; sdot += qs[r][sub * 16 + d] * kf[d]; }
    dp += __shfl_xor_sync(0xffffffffu, dp, 1);   dp += __shfl_xor_sync(0xffffffffu, dp, 2);
    sdot += __shfl_xor_sync(0xffffffffu, sdot, 1); sdot += __shfl_xor_sync(0xffffffffu, sdot, 2);
    float pj = 0.f, dsj = 0.f;
    if (sum > 0.f) {
      if (valid && p.mask[r * p.N + j] != 0) pj = __expf(sdot * p.scale - mxs[r]) / sum;
      dsj = pj * (dp - dls[r]) * p.scale;
    } else if (sum < 0.f) {
      pj = valid ? 1.0f / (float)p.N : 0.f;  // uniform row: constant w.r.t. the scores, ds = 0
    }
#pragma unroll
    for (int d = 0; d < 16; ++d) { dk[d] += dsj * qs[r][sub * 16 + d]; dv[d] += pj * ds_[r][sub * 16 + d]; }
    if (sub == 0) dss[r][tid >> 2] = dsj;   // 0 for masked / invalid keys and for uniform or empty rows
  }
  __syncthreads();
  for (int t = tid; t < p.R * 64; t += blockDim.x) {
    const int r = t >> 6, d = t & 63;
    float c = 0.f;
#pragma unroll 8
    for (int k = 0; k < 64; ++k) c += dss[r][k] * ks[k][d];
    if (c != 0.f) atomicAdd(p.dq + b * p.dq_bstride + (int64_t)r * HD + h * 64 + d, c);
  }
  if (valid) {
    __nv_bfloat16* ok = p.dkv + row * p.lddkv + h * 64 + sub * 16;
    __nv_bfloat16* ov = p.dkv + row * p.lddkv + HD + h * 64 + sub * 16;
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      *reinterpret_cast<uint4*>(ok + c * 8) = make_uint4(pack_bf16(dk[c * 8], dk[c * 8 + 1]), pack_bf16(dk[c * 8 + 2], dk[c * 8 + 3]),
                                                        pack_bf16(dk[c * 8 + 4], dk[c * 8 + 5]), pack_bf16(dk[c * 8 + 6], dk[c * 8 + 7]));
      *reinterpret_cast<uint4*>(ov + c * 8) = make_uint4(pack_bf16(dv[c * 8], dv[c * 8 + 1]), pack_bf16(dv[c * 8 + 2], dv[c * 8 + 3]),
                                                        pack_bf16(dv[c * 8 + 4], dv[c * 8 + 5]), pack_bf16(dv[c * 8 + 6], dv[c * 8 + 7]));
    }
  }
}

// delta[b, r, h] = dout[b, r, h, :] . out[b, r, h, :]   (one warp per (b, r, h))
__global__ void pool_delta_kernel(const __nv_bfloat16* dout, const __nv_bfloat16* out, float* delta, int64_t n_brh) {
  const int lane = threadIdx.x & 31;
  const int64_t w = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (w >= n_brh) return;
  const float2 a = unpack_bf16(*reinterpret_cast<const uint32_t*>(dout + w * 64 + 2 * lane));
  const float2 c = unpack_bf16(*reinterpret_cast<const uint32_t*>(out + w * 64 + 2 * lane));
  const float s = warp_sum(a.x * c.x + a.y * c.y);
  if (lane == 0) delta[w] = s;
}

}  // namespace mmf

using namespace mmf;

static int slot_check(const MmfSlotAttnArgs* a) {
  if (!a || !a->q || !a->kv_tok || !a->kv_me || !a->slotmap || !a->seg) return 1;
  if (a->S < 2 || a->S > SLOT_MAX || a->H < 1 || a->H > 32 || a->dh != 64) return 2;
  if ((a->ldq & 1) || (a->ldkv & 1) || (a->ldme & 1)) return 3;
  return 0;
}
static SlotParams slot_params(const MmfSlotAttnArgs& a) {
  SlotParams p{};
  p.q = (const __nv_bfloat16*)a.q; p.kv_tok = (const __nv_bfloat16*)a.kv_tok; p.kv_me = (const __nv_bfloat16*)a.kv_me;
  p.slotmap = a.slotmap; p.seg = a.seg; p.out = (__nv_bfloat16*)a.out; p.probs = a.probs;
  p.ldq = a.ldq; p.ldkv = a.ldkv; p.ldme = a.ldme; p.ldo = a.ldo;
  p.B = a.B; p.F = a.F; p.H = a.H; p.S = a.S; p.n_head = a.n_head; p.scale = a.scale;
  p.dout = (const __nv_bfloat16*)a.dout; p.dq = (__nv_bfloat16*)a.dq; p.dkv_tok = (__nv_bfloat16*)a.dkv_tok; p.dkv_me = a.dkv_me;
  p.lddout = a.lddout; p.lddq = a.lddq; p.lddkv = a.lddkv; p.lddme = a.lddme;
  p.me_scratch = a.me_scratch;
  return p;
}
// the wide kernels need every row slice 16-byte aligned and all heads inside one warp
static bool slot_wide_ok(const MmfSlotAttnArgs& a, bool bwd) {
  auto al = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0; };
  bool ok = a.H <= 8 && a.S >= 2 && a.S <= 5 && !(a.ldq & 7) && !(a.ldkv & 7) && !(a.ldme & 7) && al(a.q) && al(a.kv_tok) && al(a.kv_me) && !a.probs;
  if (!bwd) return ok && !(a.ldo & 7) && al(a.out);
  return ok && a.me_scratch && !(a.lddout & 7) && !(a.lddq & 7) && !(a.lddkv & 7) && !(a.lddme & 1) && al(a.dout) && al(a.dq) &&
         al(a.dkv_tok) && (reinterpret_cast<uintptr_t>(a.dkv_me) & 7) == 0;
}

extern "C" int mmf_slot_attn_fwd(const MmfSlotAttnArgs* a, mmf_stream_t stream) {
  int rc = slot_check(a);
  if (rc) MMF_BAD_ARG(rc);
  if (!a->out || (a->ldo & 1)) MMF_BAD_ARG(10);
  if ((int64_t)a->B * a->F == 0) return 0;
  if (slot_wide_ok(*a, false)) {
    launch_slot_wide<false>(slot_params(*a), reinterpret_cast<cudaStream_t>(stream));
  } else
  slot_attn_kernel<false><<<(unsigned)((int64_t)a->B * a->F), 32 * a->H, 0, reinterpret_cast<cudaStream_t>(stream)>>>(slot_params(*a));
  g_launch_count.fetch_add(1, std::memory_order_relaxed);
  MMF_LAUNCH_CHECK();
  return 0;
}

extern "C" int mmf_slot_attn_bwd(const MmfSlotAttnArgs* a, mmf_stream_t stream) {
  int rc = slot_check(a);
  if (rc) MMF_BAD_ARG(rc);
  if (!a->dout || !a->dq || !a->dkv_tok || !a->dkv_me) MMF_BAD_ARG(11);
  if ((a->lddout & 1) || (a->lddq & 1) || (a->lddkv & 1)) MMF_BAD_ARG(12);
  if ((int64_t)a->B * a->F == 0) return 0;
  if (slot_wide_ok(*a, true)) {
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    launch_slot_wide<true>(slot_params(*a), st);
    launch_pdl(slot_me_reduce_kernel, dim3((unsigned)(a->F * a->H)), dim3(MERED_WARPS * 32), 0, st, slot_params(*a));
    g_launch_count.fetch_add(1, std::memory_order_relaxed);
  } else
  slot_attn_kernel<true><<<(unsigned)((int64_t)a->B * a->F), 32 * a->H, 0, reinterpret_cast<cudaStream_t>(stream)>>>(slot_params(*a));
  g_launch_count.fetch_add(1, std::memory_order_relaxed);
  MMF_LAUNCH_CHECK();
  return 0;
}

static int pool_check(const MmfPoolAttnArgs* a) {
  if (!a || !a->q || !a->kv || !a->mask || !a->mode || !a->stat) return 1;
  if (a->R < 1 || a->R > 16 || a->H < 1 || a->dh != 64 || a->N < 1 || a->N > POOL_MAX_N) return 2;
  if (a->n_head + a->n_tail != a->N || (a->ldkv & 7)) return 3;
  return 0;
}
static PoolParams pool_params(const MmfPoolAttnArgs& a) {
  PoolParams p{};
  p.q = (const __nv_bfloat16*)a.q; p.kv = (const __nv_bfloat16*)a.kv; p.mask = a.mask; p.mode = a.mode;
  p.out = (__nv_bfloat16*)a.out; p.stat = a.stat; p.q_bstride = a.q_bstride; p.ldkv = a.ldkv;
  p.B = a.B; p.R = a.R; p.H = a.H; p.N = a.N; p.n_head = a.n_head; p.n_tail = a.n_tail; p.scale = a.scale;
  p.dout = (const __nv_bfloat16*)a.dout; p.dq = a.dq; p.dq_bstride = a.dq_bstride; p.dkv = (__nv_bfloat16*)a.dkv; p.lddkv = a.lddkv;
  return p;
}

extern "C" int mmf_pool_attn_fwd(const MmfPoolAttnArgs* a, mmf_stream_t stream) {
  int rc = pool_check(a);
  if (rc) MMF_BAD_ARG(rc);
  if (!a->out) MMF_BAD_ARG(10);
  const size_t smem = (size_t)a->R * a->N * sizeof(float);
  if (smem > 200 * 1024) MMF_BAD_ARG(11);
  static size_t configured = 0;
  if (smem > 40 * 1024 && smem > configured) {
    cudaError_t e = cudaFuncSetAttribute(pool_attn_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    configured = smem;
  }
  pool_attn_fwd_kernel<<<dim3(a->H, a->B), 32 * a->R, smem, reinterpret_cast<cudaStream_t>(stream)>>>(pool_params(*a));
  g_launch_count.fetch_add(1, std::memory_order_relaxed);
  MMF_LAUNCH_CHECK();
  return 0;
}

extern "C" int mmf_pool_attn_bwd(const MmfPoolAttnArgs* a, mmf_stream_t stream) {
  int rc = pool_check(a);
  if (rc) MMF_BAD_ARG(rc);
  if (!a->out || !a->dout || !a->dq || !a->dkv || (a->lddkv & 7) || a->dq_bstride < 0) MMF_BAD_ARG(12);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  // stat layout: [B,R,H,2] (max,sum) followed by [B,R,H] delta scratch
  const int64_t n_brh = (int64_t)a->B * a->R * a->H;
  float* delta = a->stat + n_brh * 2;
  pool_delta_kernel<<<(unsigned)((n_brh + 7) / 8), 256, 0, st>>>((const __nv_bfloat16*)a->dout, (const __nv_bfloat16*)a->out, delta, n_brh);
  pool_attn_bwd_kernel<<<dim3((a->N + 63) / 64, a->H, a->B), 256, 0, st>>>(pool_params(*a));
  g_launch_count.fetch_add(2, std::memory_order_relaxed);
  MMF_LAUNCH_CHECK();
  return 0;
}
