// Zorro-masked flash attention, forward + backward, for the MultiMAE encoder and decoders.
//
// Replaces Attention.forward's einsum / masked_fill(-finfo.max) / softmax / einsum chain
// (reference zorro_utils.py:181-193) and the decoder attention (multimae_utils.py:170-180).  The
// reference materialises S and P as [B, h, N, N] tensors 3-4x per layer; here they never leave
// the SM.  The zorro mask (multimae.py:410-426) is NOT applied as -inf: it is a segment table
// seg[0..nseg] (token ranges per modality, last segment = fusion tokens).  Query tiles are cut per
// segment, and a modality tile only ever visits its own segment's key blocks; fusion tiles visit
// all of them.  Only the ragged end of a key range is masked.
//
// v1 data path: cp.async -> XOR-swizzled smem -> ldmatrix -> mma.sync.m16n8k16 (bf16, fp32 acc).
// Token (b, i) lives at row  i < n_head ? b*n_head + i : head_rows + b*n_tail + (i - n_head)
// ("planar" layout: all modality tokens of the batch, then all fusion tokens).
#include "common.cuh"

#include <type_traits>
#include "mmf_b200.h"

#include <atomic>
#include <cstdlib>

namespace mmf {
extern std::atomic<int64_t> g_launch_count;

constexpr int ATT_BM = 64;       // query rows per CTA
constexpr int ATT_BN = 64;       // keys per inner iteration
constexpr int ATT_THREADS = 128; // 4 warps x 16 rows
constexpr int ATT_MAX_SEG = 8;

struct AttnParams {
  const __nv_bfloat16* q; const __nv_bfloat16* k; const __nv_bfloat16* v;
  __nv_bfloat16* o;
  float* lse;                      // [B, H, Nq] natural-log logsumexp of the scaled scores
  int64_t ldq, ldk, ldv, ldo;      // row strides in elements
  int B, H, Nq, Nk;
  int n_head_q, n_tail_q;          // planar split of the query tokens (n_head_q + n_tail_q == Nq)
  int n_head_k, n_tail_k;
  int64_t head_rows_q, head_rows_k;
  float scale;                     // softmax(scale * q.k)
  const int32_t* seg;              // device int32[nseg+1] or NULL (= one segment, everything attends everything)
  int nseg;
  // backward
  const __nv_bfloat16* d_o; int64_t lddo;
  const float* delta;              // [B, H, Nq]
  __nv_bfloat16* dq; __nv_bfloat16* dk; __nv_bfloat16* dv;
  int64_t lddq, lddk, lddv;
};

__device__ __forceinline__ int64_t tok_row(int b, int i, int n_head, int n_tail, int64_t head_rows) {
  return i < n_head ? (int64_t)b * n_head + i : head_rows + (int64_t)b * n_tail + (i - n_head);
}

// ---- smem tiles: [rows][DH] bf16, 16-byte chunks XOR-swizzled so ldmatrix is conflict-free ----
template <int DH>
__device__ __forceinline__ uint32_t tile_off(int row, int chunk) {
  if (DH == 64) return row * 128 + ((chunk ^ (row & 7)) << 4);
  return row * 64 + ((chunk ^ ((row >> 1) & 3)) << 4);  // DH == 32
}
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, bool valid) {
  const int sz = valid ? 16 : 0;  // src-size 0 => zero fill
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void mma16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// load `nrows` token rows [i0, i0+64) (valid while < i_end) of one head into a swizzled tile
template <int DH>
__device__ __forceinline__ void load_tile(uint32_t smem_base, const __nv_bfloat16* base, int64_t ld, int b, int h, int i0,
                                          int i_end, int n_head, int n_tail, int64_t head_rows) {
  constexpr int CH = DH / 8;  // 16B chunks per row
  for (int t = threadIdx.x; t < 64 * CH; t += ATT_THREADS) {
    const int r = t / CH, c = t % CH;
    const int i = i0 + r;
    const bool ok = i < i_end;
    const __nv_bfloat16* src = base + (ok ? tok_row(b, i, n_head, n_tail, head_rows) * ld + h * DH + c * 8 : 0);
    cp_async16(smem_base + tile_off<DH>(r, c), src, ok);
  }
}

// A fragments (16 rows x DH) of this warp's rows from a tile: frag[ks][4]
template <int DH>
__device__ __forceinline__ void load_a_frags(uint32_t tile, int row0, uint32_t (&frag)[DH / 16][4]) {
  const int lane = threadIdx.x & 31;
#pragma unroll
  for (int ks = 0; ks < DH / 16; ++ks)
    ldsm_x4(tile + tile_off<DH>(row0 + (lane & 15), ks * 2 + (lane >> 4)), frag[ks][0], frag[ks][1], frag[ks][2], frag[ks][3]);
}

// acc[16 x 64] += A[16 x DH] . T[64 x DH]^T   (T rows are the "n" index, DH contiguous: non-transposed ldmatrix)
// RAGGED (the last block of a ragged range, e.g. 196 tokens = 3 x 64 + 4): only the first 16 * npairs rows of T are live
// and the other accumulator columns stay zero.  A template flag, not a runtime test on the full blocks: with the test inside
// the unrolled loops the full-block path lost its ldmatrix / mma interleaving (0.121 -> 0.149 ms forward at N = 256).  Used by
// the two backward kernels (0.404 -> 0.368 ms at the decoders' shape); the forward kernel measured no faster with it.
template <int DH, bool RAGGED = false>
__device__ __forceinline__ void mma_a_tT(float (&acc)[8][4], const uint32_t (&a)[DH / 16][4], uint32_t tile, int npairs = 4) {
  const int lane = threadIdx.x & 31;
  const int mi = lane >> 3, r = lane & 7;
#pragma unroll
  for (int ks = 0; ks < DH / 16; ++ks) {
#pragma unroll
    for (int np = 0; np < 4; ++np) {  // pairs of n-tiles
      if (RAGGED && np >= npairs) continue;
      uint32_t b0, b1, b2, b3;
      ldsm_x4(tile + tile_off<DH>(np * 16 + (mi >> 1) * 8 + r, ks * 2 + (mi & 1)), b0, b1, b2, b3);
      mma16816(acc[2 * np], a[ks], b0, b1);
      mma16816(acc[2 * np + 1], a[ks], b2, b3);
    }
  }
}

// acc[16 x DH] += P[16 x 64] . T[64 x DH]   (T rows are the "k" index: transposed ldmatrix); P given as accumulators
template <int DH, bool RAGGED = false>
__device__ __forceinline__ void mma_p_t(float (&acc)[DH / 8][4], const float (&p)[8][4], uint32_t tile, int nks = 4) {
  const int lane = threadIdx.x & 31;
  const int mi = lane >> 3, r = lane & 7;
#pragma unroll
  for (int ks = 0; ks < 4; ++ks) {  // 16 keys per step
    if (RAGGED && ks >= nks) continue;
    uint32_t a[4];
    a[0] = pack_bf16(p[2 * ks][0], p[2 * ks][1]);
    a[1] = pack_bf16(p[2 * ks][2], p[2 * ks][3]);
    a[2] = pack_bf16(p[2 * ks + 1][0], p[2 * ks + 1][1]);
    a[3] = pack_bf16(p[2 * ks + 1][2], p[2 * ks + 1][3]);
#pragma unroll
    for (int np = 0; np < DH / 16; ++np) {
      uint32_t b0, b1, b2, b3;
      ldsm_x4_t(tile + tile_off<DH>(ks * 16 + (mi & 1) * 8 + r, np * 2 + (mi >> 1)), b0, b1, b2, b3);
      mma16816(acc[2 * np], a, b0, b1);
      mma16816(acc[2 * np + 1], a, b2, b3);
    }
  }
}

// map a tile index to (segment rows [r0, r1), key range [k0, k1)); returns false if the tile does not exist
__device__ __forceinline__ bool map_tile(const AttnParams& p, int tile, int& r0, int& r1, int& k0, int& k1) {
  if (p.seg == nullptr) {
    r0 = tile * ATT_BM; r1 = min(r0 + ATT_BM, p.Nq); k0 = 0; k1 = p.Nk;
    return r0 < p.Nq;
  }
  for (int s = 0; s < p.nseg; ++s) {
    const int a = p.seg[s], e = p.seg[s + 1];
    const int nt = (e - a + ATT_BM - 1) / ATT_BM;
    if (tile < nt) {
      r0 = a + tile * ATT_BM; r1 = min(r0 + ATT_BM, e);
      if (s == p.nseg - 1) { k0 = 0; k1 = p.Nk; } else { k0 = a; k1 = e; }
      return true;
    }
    tile -= nt;
  }
  return false;
}

// ------------------------------------------------------------------------------------------------
// forward
// ------------------------------------------------------------------------------------------------
template <int DH>
__global__ void __launch_bounds__(ATT_THREADS) attn_fwd_kernel(const AttnParams p) {
  pdl_wait();   // launched through launch_pdl (common.cuh): nothing another kernel owns is touched before this
  __shared__ __align__(128) uint8_t smem[(64 + 4 * 64) * DH * 2];
  int r0, r1, k0, k1;
  if (!map_tile(p, blockIdx.x, r0, r1, k0, k1)) return;
  const int h = blockIdx.y, b = blockIdx.z;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t sQ = smem_u32(smem);
  const uint32_t sK = sQ + 64 * DH * 2;          // 2 stages
  const uint32_t sV = sK + 2 * 64 * DH * 2;      // 2 stages
  constexpr uint32_t TILE = 64 * DH * 2;

  load_tile<DH>(sQ, p.q, p.ldq, b, h, r0, r1, p.n_head_q, p.n_tail_q, p.head_rows_q);
  load_tile<DH>(sK, p.k, p.ldk, b, h, k0, k1, p.n_head_k, p.n_tail_k, p.head_rows_k);
  load_tile<DH>(sV, p.v, p.ldv, b, h, k0, k1, p.n_head_k, p.n_tail_k, p.head_rows_k);
  cp_async_commit();

  float o[DH / 8][4];
#pragma unroll
  for (int i = 0; i < DH / 8; ++i) o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f;
  float m[2] = {-INFINITY, -INFINITY}, l[2] = {0.f, 0.f};
  const float sl2 = p.scale * 1.4426950408889634f;
  uint32_t qf[DH / 16][4];

  const int nblk = (k1 - k0 + ATT_BN - 1) / ATT_BN;
  for (int blk = 0; blk < nblk; ++blk) {
    const int st = blk & 1;
    if (blk + 1 < nblk) {
      const int kn = k0 + (blk + 1) * ATT_BN;
      load_tile<DH>(sK + (st ^ 1) * TILE, p.k, p.ldk, b, h, kn, k1, p.n_head_k, p.n_tail_k, p.head_rows_k);
      load_tile<DH>(sV + (st ^ 1) * TILE, p.v, p.ldv, b, h, kn, k1, p.n_head_k, p.n_tail_k, p.head_rows_k);
      cp_async_commit();
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();
    if (blk == 0) load_a_frags<DH>(sQ, warp * 16, qf);

    float s[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i) s[i][0] = s[i][1] = s[i][2] = s[i][3] = 0.f;
    mma_a_tT<DH>(s, qf, sK + st * TILE);

    const int kb = k0 + blk * ATT_BN;
    if (kb + ATT_BN > k1) {  // ragged end of the key range
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int c = kb + i * 8 + 2 * (lane & 3);
        if (c >= k1) s[i][0] = s[i][2] = -INFINITY;
        if (c + 1 >= k1) s[i][1] = s[i][3] = -INFINITY;
      }
    }
    float mx[2] = {m[0], m[1]};
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      mx[0] = fmaxf(mx[0], fmaxf(s[i][0], s[i][1]));
      mx[1] = fmaxf(mx[1], fmaxf(s[i][2], s[i][3]));
    }
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      mx[j] = fmaxf(mx[j], __shfl_xor_sync(0xffffffffu, mx[j], 1));
      mx[j] = fmaxf(mx[j], __shfl_xor_sync(0xffffffffu, mx[j], 2));
    }
    float corr[2], rs[2] = {0.f, 0.f};
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      corr[j] = exp2f((m[j] - mx[j]) * sl2);  // m = -inf on the first block -> 0
      m[j] = mx[j];
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      s[i][0] = exp2f((s[i][0] - mx[0]) * sl2);
      s[i][1] = exp2f((s[i][1] - mx[0]) * sl2);
      s[i][2] = exp2f((s[i][2] - mx[1]) * sl2);
      s[i][3] = exp2f((s[i][3] - mx[1]) * sl2);
      rs[0] += s[i][0] + s[i][1];
      rs[1] += s[i][2] + s[i][3];
    }
#pragma unroll
    for (int j = 0; j < 2; ++j) l[j] = l[j] * corr[j] + rs[j];
#pragma unroll
    for (int i = 0; i < DH / 8; ++i) {
      o[i][0] *= corr[0]; o[i][1] *= corr[0]; o[i][2] *= corr[1]; o[i][3] *= corr[1];
    }
    mma_p_t<DH>(o, s, sV + st * TILE);
    __syncthreads();  // everyone done with stage st before it is refilled
  }

#pragma unroll
  for (int j = 0; j < 2; ++j) {
    l[j] += __shfl_xor_sync(0xffffffffu, l[j], 1);
    l[j] += __shfl_xor_sync(0xffffffffu, l[j], 2);
  }
#pragma unroll
  for (int j = 0; j < 2; ++j) {
    const int i = r0 + warp * 16 + (lane >> 2) + j * 8;
    if (i < r1) {
      const float inv = 1.0f / l[j];
      __nv_bfloat16* orow = p.o + tok_row(b, i, p.n_head_q, p.n_tail_q, p.head_rows_q) * p.ldo + h * DH;
#pragma unroll
      for (int t = 0; t < DH / 8; ++t)
        *reinterpret_cast<uint32_t*>(orow + t * 8 + 2 * (lane & 3)) = pack_bf16(o[t][2 * j] * inv, o[t][2 * j + 1] * inv);
      if (p.lse && (lane & 3) == 0) p.lse[((int64_t)b * p.H + h) * p.Nq + i] = m[j] * p.scale + logf(l[j]);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// backward, part 0: delta[b,h,i] = sum_d dO[i,d] * O[i,d]
// ------------------------------------------------------------------------------------------------
template <int DH>
__global__ void attn_delta_kernel(const AttnParams p) {
  pdl_wait();   // launched through launch_pdl (common.cuh): nothing another kernel owns is touched before this
  // one warp per (b, i): lanes sweep the H*DH row in 16-byte chunks; heads are DH/8 chunks wide
  constexpr int CPH = DH / 8;
  const int lane = threadIdx.x & 31;
  const int64_t w = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (w >= (int64_t)p.B * p.Nq) return;
  const int b = (int)(w / p.Nq), i = (int)(w % p.Nq);
  const int64_t row = tok_row(b, i, p.n_head_q, p.n_tail_q, p.head_rows_q);
  const int nchunk = p.H * CPH;
  for (int c0 = 0; c0 < nchunk; c0 += 32) {
    const int c = c0 + lane;
    float acc = 0.f;
    if (c < nchunk) {
      const uint4 a = *reinterpret_cast<const uint4*>(p.o + row * p.ldo + c * 8);
      const uint4 d = *reinterpret_cast<const uint4*>(p.d_o + row * p.lddo + c * 8);
      const uint32_t* pa = &a.x; const uint32_t* pd = &d.x;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float2 fa = unpack_bf16(pa[k]), fd = unpack_bf16(pd[k]);
        acc += fa.x * fd.x + fa.y * fd.y;
      }
    }
    // reduce within groups of CPH lanes (CPH = 8 or 4, power of two, groups aligned)
#pragma unroll
    for (int off = CPH / 2; off > 0; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
    if (c < nchunk && (lane % CPH) == 0) {
      const int h = c / CPH;
      const_cast<float*>(p.delta)[((int64_t)b * p.H + h) * p.Nq + i] = acc;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// backward, part 1: dQ.  Same tiling as the forward (query tiles per segment).
//   P = exp(scale*S - lse);  dP = dO.V^T;  dS = P*(dP - delta);  dQ = scale * dS.K
// ------------------------------------------------------------------------------------------------
template <int DH>
__global__ void __launch_bounds__(ATT_THREADS, (DH == 32 ? 4 : 3)) attn_bwd_dq_kernel(const AttnParams p) {
  pdl_wait();   // launched through launch_pdl (common.cuh): nothing another kernel owns is touched before this   // register caps of 4 / 3 CTAs per SM (128 / 168)
  extern __shared__ __align__(128) uint8_t smem[];
  int r0, r1, k0, k1;
  if (!map_tile(p, blockIdx.x, r0, r1, k0, k1)) return;
  const int h = blockIdx.y, b = blockIdx.z;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr uint32_t TILE = 64 * DH * 2;
  const uint32_t sQ = smem_u32(smem), sdO = sQ + TILE, sK = sdO + TILE, sV = sK + 2 * TILE;

  load_tile<DH>(sQ, p.q, p.ldq, b, h, r0, r1, p.n_head_q, p.n_tail_q, p.head_rows_q);
  load_tile<DH>(sdO, p.d_o, p.lddo, b, h, r0, r1, p.n_head_q, p.n_tail_q, p.head_rows_q);
  load_tile<DH>(sK, p.k, p.ldk, b, h, k0, k1, p.n_head_k, p.n_tail_k, p.head_rows_k);
  load_tile<DH>(sV, p.v, p.ldv, b, h, k0, k1, p.n_head_k, p.n_tail_k, p.head_rows_k);
  cp_async_commit();

  float lse[2], dl[2];
#pragma unroll
  for (int j = 0; j < 2; ++j) {
    const int i = r0 + warp * 16 + (lane >> 2) + j * 8;
    const bool ok = i < r1;
    lse[j] = ok ? p.lse[((int64_t)b * p.H + h) * p.Nq + i] : INFINITY;
    dl[j] = ok ? p.delta[((int64_t)b * p.H + h) * p.Nq + i] : 0.f;
  }
  float dq[DH / 8][4];
#pragma unroll
  for (int i = 0; i < DH / 8; ++i) dq[i][0] = dq[i][1] = dq[i][2] = dq[i][3] = 0.f;
  uint32_t qf[DH / 16][4], dof[DH / 16][4];
  const float L2E = 1.4426950408889634f;

  const int nblk = (k1 - k0 + ATT_BN - 1) / ATT_BN;
  auto stage_in = [&](int blk) {   // every thread: prefetch the next key block, wait for this one
    const int st = blk & 1;
    if (blk + 1 < nblk) {
      const int kn = k0 + (blk + 1) * ATT_BN;
      load_tile<DH>(sK + (st ^ 1) * TILE, p.k, p.ldk, b, h, kn, k1, p.n_head_k, p.n_tail_k, p.head_rows_k);
      load_tile<DH>(sV + (st ^ 1) * TILE, p.v, p.ldv, b, h, kn, k1, p.n_head_k, p.n_tail_k, p.head_rows_k);
      cp_async_commit();
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();
  };
  auto body = [&](auto ragged, int blk) {   // RG: the ragged last key block: only its 16-key groups that hold keys are worked on
    constexpr bool RG = decltype(ragged)::value;
    const int st = blk & 1;
    const int kb = k0 + blk * ATT_BN;
    const int npairs = RG ? min(4, (k1 - kb + 15) >> 4) : 4;
    float s[8][4], dp[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      s[i][0] = s[i][1] = s[i][2] = s[i][3] = 0.f;
      dp[i][0] = dp[i][1] = dp[i][2] = dp[i][3] = 0.f;
    }
    mma_a_tT<DH, RG>(s, qf, sK + st * TILE, npairs);
    mma_a_tT<DH, RG>(dp, dof, sV + st * TILE, npairs);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (RG && i >= 2 * npairs) { s[i][0] = s[i][1] = s[i][2] = s[i][3] = 0.f; continue; }
      const int c = kb + i * 8 + 2 * (lane & 3);
      const bool v0 = !RG || c < k1, v1 = !RG || c + 1 < k1;
      const float p0 = v0 ? exp2f((s[i][0] * p.scale - lse[0]) * L2E) : 0.f;
      const float p1 = v1 ? exp2f((s[i][1] * p.scale - lse[0]) * L2E) : 0.f;
      const float p2 = v0 ? exp2f((s[i][2] * p.scale - lse[1]) * L2E) : 0.f;
      const float p3 = v1 ? exp2f((s[i][3] * p.scale - lse[1]) * L2E) : 0.f;
      s[i][0] = p0 * (dp[i][0] - dl[0]);
      s[i][1] = p1 * (dp[i][1] - dl[0]);
      s[i][2] = p2 * (dp[i][2] - dl[1]);
      s[i][3] = p3 * (dp[i][3] - dl[1]);
    }
    mma_p_t<DH, RG>(dq, s, sK + st * TILE, npairs);  // dQ += dS . K
  };
  // A warp whose 16 query rows all lie past the tile's end (the last 64-row tile of 196 tokens holds 4 rows) only takes part in
  // the cooperative loads and the barriers.
  const bool warp_live = r0 + warp * 16 < r1;
  const bool ragged_last = ((k1 - k0) % ATT_BN) != 0;
  const int nfull = ragged_last ? nblk - 1 : nblk;
  if (!warp_live) {
    for (int blk = 0; blk < nblk; ++blk) { stage_in(blk); __syncthreads(); }
  } else {
    for (int blk = 0; blk < nfull; ++blk) {
      stage_in(blk);
      if (blk == 0) {
        load_a_frags<DH>(sQ, warp * 16, qf);
        load_a_frags<DH>(sdO, warp * 16, dof);
      }
      body(std::false_type{}, blk);
      __syncthreads();
    }
    if (ragged_last) {
      stage_in(nfull);
      if (nfull == 0) {
        load_a_frags<DH>(sQ, warp * 16, qf);
        load_a_frags<DH>(sdO, warp * 16, dof);
      }
      body(std::true_type{}, nfull);
      __syncthreads();
    }
  }
#pragma unroll
  for (int j = 0; j < 2; ++j) {
    const int i = r0 + warp * 16 + (lane >> 2) + j * 8;
    if (i < r1) {
      __nv_bfloat16* row = p.dq + tok_row(b, i, p.n_head_q, p.n_tail_q, p.head_rows_q) * p.lddq + h * DH;
#pragma unroll
      for (int t = 0; t < DH / 8; ++t)
        *reinterpret_cast<uint32_t*>(row + t * 8 + 2 * (lane & 3)) = pack_bf16(dq[t][2 * j] * p.scale, dq[t][2 * j + 1] * p.scale);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// backward, part 2: dK, dV.  One CTA per 64-key tile (cut per segment); it visits the query blocks
// that attend to it: its own segment's rows plus the fusion (last) segment's rows.
//   work on transposed tiles so that each warp owns 16 keys:
//   S^T = K.Q^T;  P^T = exp(scale*S^T - lse[q]);  dV += P^T.dO;  dP^T = V.dO^T;
//   dS^T = P^T*(dP^T - delta[q]);  dK += scale * dS^T.Q
// ------------------------------------------------------------------------------------------------
template <int DH>
__global__ void __launch_bounds__(ATT_THREADS) attn_bwd_dkv_kernel(const AttnParams p) {
  pdl_wait();   // launched through launch_pdl (common.cuh): nothing another kernel owns is touched before this
  extern __shared__ __align__(128) uint8_t smem[];
  // key tile: reuse map_tile on the key axis (self-attention: Nq == Nk, same segments)
  int c0, c1;
  int seg_idx = -1;
  {
    int tile = blockIdx.x;
    if (p.seg == nullptr) {
      c0 = tile * 64; c1 = min(c0 + 64, p.Nk);
      if (c0 >= p.Nk) return;
    } else {
      bool found = false;
      for (int s = 0; s < p.nseg; ++s) {
        const int a = p.seg[s], e = p.seg[s + 1];
        const int nt = (e - a + 63) / 64;
        if (tile < nt) { c0 = a + tile * 64; c1 = min(c0 + 64, e); seg_idx = s; found = true; break; }
        tile -= nt;
      }
      if (!found) return;
    }
  }
  const int h = blockIdx.y, b = blockIdx.z;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr uint32_t TILE = 64 * DH * 2;
  const uint32_t sK = smem_u32(smem), sV = sK + TILE, sQ = sV + TILE, sdO = sQ + 2 * TILE;
  float* s_lse = reinterpret_cast<float*>(smem + 6 * TILE);  // [2][64]
  float* s_dl = s_lse + 128;                                 // [2][64]

  // query ranges attending to this key tile
  int qa[2], qe[2], nrange;
  if (p.seg == nullptr) { qa[0] = 0; qe[0] = p.Nq; nrange = 1; }
  else if (seg_idx == p.nseg - 1) { qa[0] = p.seg[seg_idx]; qe[0] = p.seg[seg_idx + 1]; nrange = 1; }
  else { qa[0] = p.seg[seg_idx]; qe[0] = p.seg[seg_idx + 1]; qa[1] = p.seg[p.nseg - 1]; qe[1] = p.seg[p.nseg]; nrange = 2; }
  const int nb0 = (qe[0] - qa[0] + 63) / 64;
  const int nb1 = nrange > 1 ? (qe[1] - qa[1] + 63) / 64 : 0;
  const int nblk = nb0 + nb1;

  auto issue_q = [&](int blk, int st) {
    int a, e;
    if (blk < nb0) { a = qa[0] + blk * 64; e = qe[0]; } else { a = qa[1] + (blk - nb0) * 64; e = qe[1]; }
    load_tile<DH>(sQ + st * TILE, p.q, p.ldq, b, h, a, e, p.n_head_q, p.n_tail_q, p.head_rows_q);
    load_tile<DH>(sdO + st * TILE, p.d_o, p.lddo, b, h, a, e, p.n_head_q, p.n_tail_q, p.head_rows_q);
    if (threadIdx.x < 64) {
      const int i = a + threadIdx.x;
      const bool ok = i < e;
      s_lse[st * 64 + threadIdx.x] = ok ? p.lse[((int64_t)b * p.H + h) * p.Nq + i] : INFINITY;
      s_dl[st * 64 + threadIdx.x] = ok ? p.delta[((int64_t)b * p.H + h) * p.Nq + i] : 0.f;
    }
  };

  load_tile<DH>(sK, p.k, p.ldk, b, h, c0, c1, p.n_head_k, p.n_tail_k, p.head_rows_k);
  load_tile<DH>(sV, p.v, p.ldv, b, h, c0, c1, p.n_head_k, p.n_tail_k, p.head_rows_k);
  if (nblk > 0) issue_q(0, 0);
  cp_async_commit();

  float dk[DH / 8][4], dv[DH / 8][4];
#pragma unroll
  for (int i = 0; i < DH / 8; ++i) {
    dk[i][0] = dk[i][1] = dk[i][2] = dk[i][3] = 0.f;
    dv[i][0] = dv[i][1] = dv[i][2] = dv[i][3] = 0.f;
  }
  uint32_t kf[DH / 16][4], vf[DH / 16][4];
  const float L2E = 1.4426950408889634f;
  const int key_a = c0 + warp * 16 + (lane >> 2);  // this thread's two key rows: key_a, key_a + 8

  auto stage_in = [&](int blk) {   // every thread: prefetch the next query block, wait for this one
    const int st = blk & 1;
    if (blk + 1 < nblk) {
      issue_q(blk + 1, st ^ 1);
      cp_async_commit();
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();
  };
  auto body = [&](auto ragged, int blk, int qrows) {   // RG: a query block with fewer than 64 live rows (lse = +inf beyond them)
    constexpr bool RG = decltype(ragged)::value;
    const int st = blk & 1;
    const int npairs = RG ? min(4, (qrows + 15) >> 4) : 4;
    float s[8][4], dp[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      s[i][0] = s[i][1] = s[i][2] = s[i][3] = 0.f;
      dp[i][0] = dp[i][1] = dp[i][2] = dp[i][3] = 0.f;
    }
    mma_a_tT<DH, RG>(s, kf, sQ + st * TILE, npairs);     // S^T[key, q]
    mma_a_tT<DH, RG>(dp, vf, sdO + st * TILE, npairs);   // dP^T[key, q]
    const bool kv0 = key_a < c1, kv1 = key_a + 8 < c1;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (RG && i >= 2 * npairs) {
        s[i][0] = s[i][1] = s[i][2] = s[i][3] = 0.f;
        dp[i][0] = dp[i][1] = dp[i][2] = dp[i][3] = 0.f;
        continue;
      }
      const int qc = i * 8 + 2 * (lane & 3);
      const float l0 = s_lse[st * 64 + qc], l1 = s_lse[st * 64 + qc + 1];
      const float d0 = s_dl[st * 64 + qc], d1 = s_dl[st * 64 + qc + 1];
      const float p0 = kv0 ? exp2f((s[i][0] * p.scale - l0) * L2E) : 0.f;
      const float p1 = kv0 ? exp2f((s[i][1] * p.scale - l1) * L2E) : 0.f;
      const float p2 = kv1 ? exp2f((s[i][2] * p.scale - l0) * L2E) : 0.f;
      const float p3 = kv1 ? exp2f((s[i][3] * p.scale - l1) * L2E) : 0.f;
      s[i][0] = p0; s[i][1] = p1; s[i][2] = p2; s[i][3] = p3;
      dp[i][0] = p0 * (dp[i][0] - d0);
      dp[i][1] = p1 * (dp[i][1] - d1);
      dp[i][2] = p2 * (dp[i][2] - d0);
      dp[i][3] = p3 * (dp[i][3] - d1);
    }
    mma_p_t<DH, RG>(dv, s, sdO + st * TILE, npairs);  // dV += P^T . dO
    mma_p_t<DH, RG>(dk, dp, sQ + st * TILE, npairs);  // dK += dS^T . Q
  };
  const bool warp_live = c0 + warp * 16 < c1;   // as in attn_bwd_dq_kernel (here the warp's rows are keys, the blocks queries)
  if (!warp_live) {
    for (int blk = 0; blk < nblk; ++blk) { stage_in(blk); __syncthreads(); }
  } else {
    for (int blk = 0; blk < nblk; ++blk) {
      stage_in(blk);
      if (blk == 0) {
        load_a_frags<DH>(sK, warp * 16, kf);
        load_a_frags<DH>(sV, warp * 16, vf);
      }
      const int qrows = blk < nb0 ? qe[0] - (qa[0] + blk * 64) : qe[1] - (qa[1] + (blk - nb0) * 64);   // warp-uniform
      if (qrows >= 64) body(std::false_type{}, blk, 64); else body(std::true_type{}, blk, qrows);
      __syncthreads();
    }
  }
#pragma unroll
  for (int j = 0; j < 2; ++j) {
    const int i = key_a + j * 8;
    if (i < c1) {
      const int64_t row = tok_row(b, i, p.n_head_k, p.n_tail_k, p.head_rows_k);
      __nv_bfloat16* rk = p.dk + row * p.lddk + h * DH;
      __nv_bfloat16* rv = p.dv + row * p.lddv + h * DH;
#pragma unroll
      for (int t = 0; t < DH / 8; ++t) {
        *reinterpret_cast<uint32_t*>(rk + t * 8 + 2 * (lane & 3)) = pack_bf16(dk[t][2 * j] * p.scale, dk[t][2 * j + 1] * p.scale);
        *reinterpret_cast<uint32_t*>(rv + t * 8 + 2 * (lane & 3)) = pack_bf16(dv[t][2 * j], dv[t][2 * j + 1]);
      }
    }
  }
}

int attn_fwd_tc_launch(const MmfAttnArgs* a, cudaStream_t stream);  // attention_tc.cu (tcgen05 path, dh = 64 self-attention)
int attn_bwd_tc_launch(const MmfAttnArgs* a, cudaStream_t stream);

static int check_common(const MmfAttnArgs* a) {
  if (!a || !a->q || !a->k || !a->v) return 1;
  if (a->dh != 64 && a->dh != 32) return 2;
  if (a->B <= 0 || a->H <= 0 || a->Nq <= 0 || a->Nk <= 0) return 3;
  if ((a->ldq & 7) || (a->ldk & 7) || (a->ldv & 7)) return 4;
  if (a->n_head_q + a->n_tail_q != a->Nq || a->n_head_k + a->n_tail_k != a->Nk) return 5;
  if (a->seg && (a->nseg < 1 || a->nseg > ATT_MAX_SEG || a->Nq != a->Nk)) return 6;
  if ((reinterpret_cast<uintptr_t>(a->q) & 15) || (reinterpret_cast<uintptr_t>(a->k) & 15) || (reinterpret_cast<uintptr_t>(a->v) & 15)) return 7;
  return 0;
}

static AttnParams to_params(const MmfAttnArgs& a) {
  AttnParams p{};
  p.q = reinterpret_cast<const __nv_bfloat16*>(a.q); p.k = reinterpret_cast<const __nv_bfloat16*>(a.k);
  p.v = reinterpret_cast<const __nv_bfloat16*>(a.v); p.o = reinterpret_cast<__nv_bfloat16*>(a.o);
  p.lse = a.lse; p.ldq = a.ldq; p.ldk = a.ldk; p.ldv = a.ldv; p.ldo = a.ldo;
  p.B = a.B; p.H = a.H; p.Nq = a.Nq; p.Nk = a.Nk;
  p.n_head_q = a.n_head_q; p.n_tail_q = a.n_tail_q; p.n_head_k = a.n_head_k; p.n_tail_k = a.n_tail_k;
  p.head_rows_q = (int64_t)a.B * a.n_head_q; p.head_rows_k = (int64_t)a.B * a.n_head_k;
  p.scale = a.scale; p.seg = a.seg; p.nseg = a.nseg;
  p.d_o = reinterpret_cast<const __nv_bfloat16*>(a.d_o); p.lddo = a.lddo; p.delta = a.delta;
  p.dq = reinterpret_cast<__nv_bfloat16*>(a.dq); p.dk = reinterpret_cast<__nv_bfloat16*>(a.dk);
  p.dv = reinterpret_cast<__nv_bfloat16*>(a.dv); p.lddq = a.lddq; p.lddk = a.lddk; p.lddv = a.lddv;
  return p;
}

}  // namespace mmf

using namespace mmf;

extern "C" int mmf_attn_fwd(const MmfAttnArgs* a, mmf_stream_t stream) {
  int rc = check_common(a);
  if (rc) MMF_BAD_ARG(rc);
  if (!a->o || (a->ldo & 1)) MMF_BAD_ARG(20);
  {
    // dh = 64 self-attention runs on the tcgen05 kernel; dh = 32 (decoders) and cross-attention use the kernel below
    static const bool tc_on = !(getenv("MMF_ATTN_TC") && atoi(getenv("MMF_ATTN_TC")) == 0);
    if (tc_on) {
      const int rc_tc = attn_fwd_tc_launch(a, reinterpret_cast<cudaStream_t>(stream));
      if (rc_tc != -1000) return rc_tc;
    }
  }
  AttnParams p = to_params(*a);
  const int tiles = (a->Nq + ATT_BM - 1) / ATT_BM + (a->seg ? a->nseg : 0);
  dim3 grid(tiles, a->H, a->B);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (a->dh == 64) launch_pdl(attn_fwd_kernel<64>, dim3(grid), dim3(ATT_THREADS), 0, st, p);
  else launch_pdl(attn_fwd_kernel<32>, dim3(grid), dim3(ATT_THREADS), 0, st, p);
  g_launch_count.fetch_add(1, std::memory_order_relaxed);
  MMF_LAUNCH_CHECK();
  return 0;
}

extern "C" int mmf_attn_bwd(const MmfAttnArgs* a, mmf_stream_t stream) {
  int rc = check_common(a);
  if (rc) MMF_BAD_ARG(rc);
  if (!a->o || !a->lse || !a->d_o || !a->delta || !a->dq || !a->dk || !a->dv) MMF_BAD_ARG(21);
  if ((a->ldo & 7) || (a->lddo & 7) || (a->lddq & 1) || (a->lddk & 1) || (a->lddv & 1)) MMF_BAD_ARG(22);
  AttnParams p = to_params(*a);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int64_t rows = (int64_t)a->B * a->Nq;
  const int dgrid = (int)((rows + 7) / 8);
  const int qtiles = (a->Nq + 63) / 64 + (a->seg ? a->nseg : 0);
  const int ktiles = (a->Nk + 63) / 64 + (a->seg ? a->nseg : 0);
  {
    static const bool tc_on = !(getenv("MMF_ATTN_TC") && atoi(getenv("MMF_ATTN_TC")) == 0);
    if (tc_on && a->dh == 64 && a->Nq == a->Nk && a->n_head_q == a->n_head_k && !(a->ldo & 7) && !(a->lddo & 7)) {
      launch_pdl(attn_delta_kernel<64>, dim3(dgrid), dim3(256), 0, st, p);
      g_launch_count.fetch_add(1, std::memory_order_relaxed);
      const int rc_tc = attn_bwd_tc_launch(a, st);
      if (rc_tc != -1000) return rc_tc;
    }
  }
  const int smem_dq = 6 * 64 * a->dh * 2, smem_dkv = 6 * 64 * a->dh * 2 + 4 * 64 * 4;
  if (a->dh == 64) {
    static DeviceOnce attr;
    const int attr_dev = current_device();
    if (!attr.done(attr_dev)) {
      cudaFuncSetAttribute(attn_bwd_dq_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_dq);
      cudaFuncSetAttribute(attn_bwd_dkv_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_dkv);
      attr.set(attr_dev);
    }
    launch_pdl(attn_delta_kernel<64>, dim3(dgrid), dim3(256), 0, st, p);
    launch_pdl(attn_bwd_dq_kernel<64>, dim3(dim3(qtiles, a->H, a->B)), dim3(ATT_THREADS), smem_dq, st, p);
    launch_pdl(attn_bwd_dkv_kernel<64>, dim3(dim3(ktiles, a->H, a->B)), dim3(ATT_THREADS), smem_dkv, st, p);
  } else {
    launch_pdl(attn_delta_kernel<32>, dim3(dgrid), dim3(256), 0, st, p);
    launch_pdl(attn_bwd_dq_kernel<32>, dim3(dim3(qtiles, a->H, a->B)), dim3(ATT_THREADS), smem_dq, st, p);
    launch_pdl(attn_bwd_dkv_kernel<32>, dim3(dim3(ktiles, a->H, a->B)), dim3(ATT_THREADS), smem_dkv, st, p);
  }
  g_launch_count.fetch_add(3, std::memory_order_relaxed);
  MMF_LAUNCH_CHECK();
  return 0;
}
