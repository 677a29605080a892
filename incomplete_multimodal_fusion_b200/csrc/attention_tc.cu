// Zorro-masked flash attention FORWARD on tcgen05 tensor cores (dh = 64).
//
//   per CTA: one 128-row query tile of one (batch, head).  Q, K, V tiles arrive by TMA (128B swizzle) straight
//   from the fused [rows, 3*H*64] qkv matrix; S = Q.K^T accumulates in TMEM (128 lanes x 128 fp32 columns);
//   four softmax warps (one thread per query row = one TMEM lane) read S with tcgen05.ld, keep the running
//   max / sum in registers, and write P as packed bf16 back to TMEM with tcgen05.st; O += P.V is then a
//   tcgen05.mma with the A operand in TMEM and V read MN-major from shared memory, accumulating in TMEM.
//   The O rescale of online softmax is lazy (only when the row max grows by more than 2^8).
//
//   The zorro mask (multimae.py:410-426) is the segment table: query tiles are cut per segment; a modality
//   tile visits only its own segment's key blocks, a fusion tile visits everything.  Masked key blocks are
//   never loaded; only the ragged end of a key range is masked in registers, and the MMA N / K extents
//   shrink to the valid keys (multiples of 16).
//
//   Key blocks are 64 wide and P is written IN PLACE over the S columns it was computed from (the P.V MMA and the
//   next block's S MMA are issued in that order and the tensor pipe executes in order), so a CTA needs only 128 TMEM
//   columns (S/P 64 | O 64) and ~65 KB of shared memory: THREE CTAs share an SM.  The kernel is paced by the softmax
//   (16 ex2/clk/SM; clock instrumentation: a CTA spends ~60 % of its time in the exponent passes and ~20 % waiting for
//   the next S), so what matters is how many independent CTAs keep the MUFU pipe busy while others wait on the tensor
//   pipe or on TMA.
#include "common.cuh"
#include "mmf_b200.h"

#include <atomic>
#include <mutex>

namespace mmf {
extern std::atomic<int64_t> g_launch_count;

constexpr int TC_BM = 128;   // query rows per CTA (UMMA M)
constexpr int TC_BN = 64;    // keys per block (UMMA N of S, K extent of P.V)
constexpr int TC_THREADS = 192;
constexpr int TC_CTAS_PER_SM = 3;   // four would fit TMEM (4 x 128 columns) but not the 64-register tcgen05.ld of a whole S row
constexpr int TC_QBUF = 2;          // Q tile buffers: the next head's Q arrives while the current head is in softmax
constexpr int TC_TILE_BYTES = 128 * 64 * 2;      // one 128 x 64 bf16 tile (Q; also the backward kernels' 128-row tiles)
constexpr int TC_KV_BYTES = TC_BN * 64 * 2;      // one key / value block
constexpr int TC_SMEM = TC_QBUF * TC_TILE_BYTES + 4 * TC_KV_BYTES + 1024 + 128;   // Q + 2x(K,V) + align + barriers
constexpr uint32_t TMEM_S = 0, TMEM_P = 0, TMEM_O = 64, TMEM_COLS = 128;

struct AttnTcParams {
  __nv_bfloat16* o;
  float* lse;
  int64_t ldo;
  int B, H, N;                // self-attention: Nq == Nk == N
  int n_head, n_tail;
  int64_t head_rows;
  float scale_log2;           // scale * log2(e)
  const int32_t* seg;
  int nseg;
  int tiles;                  // query tiles per sample the grid was sized for (upper bound)
  int pf;                     // L2 prefetch distance in key blocks (0 = off)
};

__device__ __forceinline__ void umma_bf16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc, bool accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(tmem_d),
      "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"((uint32_t)accumulate)
      : "memory");
}
__device__ __forceinline__ void tmem_st_32x16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
      "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ void tmem_ld_32x16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_st_32x32(uint32_t taddr, const uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
      "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]),
      "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]),
      "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
      : "memory");
}

// Row-wise softmax pieces on one 64-key block of S held in registers (one query row per thread); `nv` = valid keys.
__device__ __forceinline__ float row_max64(const uint32_t (&v)[64], int nv) {
  float m = -INFINITY;
  if (nv >= 64) {
#pragma unroll
    for (int t = 0; t < 64; ++t) m = fmaxf(m, __uint_as_float(v[t]));
  } else {
#pragma unroll
    for (int t = 0; t < 64; ++t)
      if (t < nv) m = fmaxf(m, __uint_as_float(v[t]));
  }
  return m;
}
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// P = exp2(s * scale_log2 - m_ref) packed to bf16 pairs (keys >= nv -> 0); returns the row sum
__device__ __forceinline__ float exp_pack64(const uint32_t (&v)[64], uint32_t (&pk)[32], int nv, float scale_log2, float m_ref) {
  float rs0 = 0.f, rs1 = 0.f;
  if (nv >= 64) {
#pragma unroll
    for (int t = 0; t < 32; ++t) {
      const float p0 = ex2_approx(fmaf(__uint_as_float(v[2 * t]), scale_log2, -m_ref));
      const float p1 = ex2_approx(fmaf(__uint_as_float(v[2 * t + 1]), scale_log2, -m_ref));
      rs0 += p0; rs1 += p1;
      pk[t] = pack_bf16(p0, p1);
    }
  } else {
#pragma unroll
    for (int t = 0; t < 32; ++t) {
      const float p0 = 2 * t < nv ? ex2_approx(fmaf(__uint_as_float(v[2 * t]), scale_log2, -m_ref)) : 0.f;
      const float p1 = 2 * t + 1 < nv ? ex2_approx(fmaf(__uint_as_float(v[2 * t + 1]), scale_log2, -m_ref)) : 0.f;
      rs0 += p0; rs1 += p1;
      pk[t] = pack_bf16(p0, p1);
    }
  }
  return rs0 + rs1;
}

// key blocks of a query tile: the key range [k0, k1) is cut at the head/tail plane boundary, each part into
// blocks of TC_BN keys.  Block j -> (global row of its first key, number of valid keys).
struct KeyBlocks {
  int a0, e0, a1, e1, nb0, nb;
  int n_head, n_tail, b;
  int64_t head_rows;
  __device__ void init(int k0, int k1, int n_head_, int n_tail_, int64_t head_rows_, int b_) {
    n_head = n_head_; n_tail = n_tail_; head_rows = head_rows_; b = b_;
    a0 = k0; e0 = min(k1, n_head_);            // part in the head plane
    a1 = max(k0, n_head_); e1 = k1;            // part in the tail plane
    if (e0 < a0) e0 = a0;
    if (e1 < a1) e1 = a1;
    nb0 = (e0 - a0 + TC_BN - 1) / TC_BN;
    nb = nb0 + (e1 - a1 + TC_BN - 1) / TC_BN;
  }
  __device__ void get(int j, int64_t& row, int& nvalid) const {
    if (j < nb0) {
      const int k = a0 + j * TC_BN;
      row = (int64_t)b * n_head + k;
      nvalid = min(TC_BN, e0 - k);
    } else {
      const int k = a1 + (j - nb0) * TC_BN;
      row = head_rows + (int64_t)b * n_tail + (k - n_head);
      nvalid = min(TC_BN, e1 - k);
    }
  }
};


// CTA -> (query tile, sample, head range) for the kernels whose CTAs own a query tile.  Tiles of the LAST segment (the
// fusion tokens: they visit every key block, ~4.5x the work of a modality tile at cfg 2) are numbered first, for all
// samples, and the light tiles afterwards: with the natural order the last heavy CTAs start late and the grid ends on a
// long tail (2.9 waves of 3 CTAs per SM); forward 0.29 -> 0.24 ms at cfg 2.  A heavy tile can also be split into
// TC_HSPLIT CTAs over the heads (measured: no further gain, so 1).  `tiles` = the host's upper bound of tiles per sample;
// the grid is (tiles + (TC_HSPLIT - 1) * heavy_max) * B CTAs, surplus ones exit.
constexpr int TC_HSPLIT = 1;
constexpr int TC_FWD_DEFAULT = 2;
constexpr int TC_PF_DEFAULT = 0;    // MMF_ATTN_PF default: L2 prefetch distance in blocks
constexpr int TC_DQ_DEFAULT = 2;    // MMF_ATTN_DQ default (see attn_bwd_tc_launch)   // MMF_ATTN_FWD default (see attn_fwd_tc_launch)
__device__ __forceinline__ void lpt_tile(const int32_t* seg, int nseg, int tiles, int B, int H, int& tile, int& b, int& h0, int& h1) {
  const int L = blockIdx.x;
  h0 = 0; h1 = H;
  if (seg == nullptr) { tile = L % tiles; b = L / tiles; if (b >= B) tile = 1 << 20; return; }
  int first_heavy = 0;
  for (int s = 0; s + 1 < nseg; ++s) first_heavy += (seg[s + 1] - seg[s] + TC_BM - 1) / TC_BM;
  const int th = (seg[nseg] - seg[nseg - 1] + TC_BM - 1) / TC_BM;
  const int hs = (H % TC_HSPLIT == 0) ? TC_HSPLIT : 1;
  if (L < th * B * hs) {
    const int part = L % hs, t = L / hs;
    b = t / th; tile = first_heavy + t % th;
    h0 = part * (H / hs); h1 = h0 + H / hs;
    return;
  }
  const int tl = tiles - th, L2 = L - th * B * hs;
  b = L2 / tl;
  const int t = L2 % tl;
  tile = (t < first_heavy && b < B) ? t : (1 << 20);   // surplus indices land past the last tile
}

// The same idea for the dK/dV kernel, whose CTAs own a KEY tile: there the modality tiles are the heavier ones (their
// keys are read by their own segment's queries and by every fusion query), so all samples' modality tiles come first.
__device__ __forceinline__ void lpt_key_tile(const int32_t* seg, int nseg, int tiles, int B, int& tile, int& b) {
  const int L = blockIdx.x;
  if (seg == nullptr) { tile = L % tiles; b = L / tiles; if (b >= B) tile = 1 << 20; return; }
  int nmod = 0;
  for (int s = 0; s + 1 < nseg; ++s) nmod += (seg[s + 1] - seg[s] + TC_BM - 1) / TC_BM;
  const int th = (seg[nseg] - seg[nseg - 1] + TC_BM - 1) / TC_BM;
  if (nmod > 0 && L < nmod * B) { b = L / nmod; tile = L % nmod; return; }
  const int tl = tiles - nmod, L2 = L - nmod * B;
  b = L2 / tl;
  const int t = L2 % tl;
  tile = (t < th && b < B) ? nmod + t : (1 << 20);
}

#ifdef MMF_ATTN_CLOCKS
__device__ unsigned long long g_attn_clk[16];
#define CLK(i, expr) do { if (dbg_on) { const long long t__ = clock64(); g_attn_clk[i] += (unsigned long long)(t__ - t_last); t_last = t__; } } while (0)
#define CLK2(i, expr) do { if (dbg_on) { const unsigned t__ = clock(); acc_clk[i] += t__ - t_last; t_last = t__; } } while (0)
#else
#define CLK(i, expr) do { } while (0)
#define CLK2(i, expr) do { } while (0)
#endif

__global__ void __launch_bounds__(TC_THREADS, TC_CTAS_PER_SM)
attn_fwd_tc_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_k,
                   const __grid_constant__ CUtensorMap tmap_v, const AttnTcParams p) {
  extern __shared__ uint8_t smem_raw[];
  // ---- which query tile? (tiles are cut per segment) ----
  int r0 = 0, r1 = 0, k0 = 0, k1 = 0, b = 0, h0 = 0, h1 = 0;
  {
    int tile;
    lpt_tile(p.seg, p.nseg, p.tiles, p.B, p.H, tile, b, h0, h1);
    bool found = false;
    if (p.seg == nullptr) {
      r0 = tile * TC_BM; r1 = min(r0 + TC_BM, p.N); k0 = 0; k1 = p.N; found = r0 < p.N;
    } else {
      for (int s = 0; s < p.nseg; ++s) {
        const int a = p.seg[s], e = p.seg[s + 1];
        const int nt = (e - a + TC_BM - 1) / TC_BM;
        if (tile < nt) {
          r0 = a + tile * TC_BM; r1 = min(r0 + TC_BM, e);
          if (s == p.nseg - 1) { k0 = 0; k1 = p.N; } else { k0 = a; k1 = e; }
          found = true;
          break;
        }
        tile -= nt;
      }
    }
    if (!found) return;
  }
  const int nh = h1 - h0;   // this CTA's heads are h0 .. h1-1; h below counts from h0
  // One CTA serves this query tile for ALL heads of one sample: barriers / TMEM are set up once and the TMA
  // producer runs ahead into the next head's Q/K/V while the current head is in softmax.
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;                            // 2 buffers
  uint8_t* sK = smem + TC_QBUF * TC_TILE_BYTES;  // 2 stages of 64 keys
  uint8_t* sV = sK + 2 * TC_KV_BYTES;            // 2 stages
  uint64_t* bars = reinterpret_cast<uint64_t*>(sV + 2 * TC_KV_BYTES);
  uint64_t* q_full = bars;         // [2]
  uint64_t* q_empty = bars + 2;    // [2]
  uint64_t* kv_full = bars + 4;    // [2]
  uint64_t* kv_empty = bars + 6;   // [2]
  uint64_t* s_full = bars + 8;
  uint64_t* p_full = bars + 9;
  uint64_t* o_full = bars + 10;
  uint64_t* o_empty = bars + 11;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 12);

  // all of a query tile's rows live in one plane (segments never straddle the head/tail boundary)
  const int64_t q_row0 = r0 < p.n_head ? (int64_t)b * p.n_head + r0 : p.head_rows + (int64_t)b * p.n_tail + (r0 - p.n_head);
  KeyBlocks kb;
  kb.init(k0, k1, p.n_head, p.n_tail, p.head_rows, b);

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_q);
    tma_prefetch_desc(&tmap_k);
    tma_prefetch_desc(&tmap_v);
  }
  if (warp == 1) {
    if (lane == 0) {
      for (int s = 0; s < 2; ++s) {
        mbar_init(&q_full[s], 1); mbar_init(&q_empty[s], 1);
        mbar_init(&kv_full[s], 1); mbar_init(&kv_empty[s], 1);
      }
      mbar_init(s_full, 1);
      mbar_init(p_full, 4);
      mbar_init(o_full, 1);
      mbar_init(o_empty, 4);
      mbar_fence_init();
    }
    __syncwarp();
    tmem_alloc(tmem_slot, TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp == 0) {
    // ------------------------------ TMA producer ------------------------------
    if (lane == 0) {
      int g = 0;  // running key-block counter over all heads
      for (int h = 0; h < nh; ++h) {
        const int qs = h % TC_QBUF;
        mbar_wait(&q_empty[qs], ((h / TC_QBUF) & 1) ^ 1);
        mbar_expect_tx(&q_full[qs], TC_TILE_BYTES);
        tma_load_2d(sQ + qs * TC_TILE_BYTES, &tmap_q, &q_full[qs], (h0 + h) * 64, (int)q_row0);
        for (int j = 0; j < kb.nb; ++j, ++g) {
          const int st = g & 1;
          mbar_wait(&kv_empty[st], ((g >> 1) & 1) ^ 1);
          int64_t row; int nvalid;
          kb.get(j, row, nvalid);
          mbar_expect_tx(&kv_full[st], 2 * TC_KV_BYTES);
          tma_load_2d(sK + st * TC_KV_BYTES, &tmap_k, &kv_full[st], (h0 + h) * 64, (int)row);
          tma_load_2d(sV + st * TC_KV_BYTES, &tmap_v, &kv_full[st], (h0 + h) * 64, (int)row);
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------ MMA issuer ------------------------------
    if (lane == 0) {
      const uint32_t idesc_pv = umma_idesc_bf16(TC_BM, 64, false, true);   // A = P (TMEM, K-major), B = V (MN-major)
      int g = 0;
      auto issue_s = [&](int h, int j, int gg) {
        const int st = gg & 1;
        if (j == 0) {
          mbar_wait(&q_full[h % TC_QBUF], (h / TC_QBUF) & 1);
        }
        mbar_wait(&kv_full[st], (gg >> 1) & 1);
        tc_fence_after();
        int64_t row; int nvalid;
        kb.get(j, row, nvalid);
        const int n16 = (nvalid + 15) & ~15;
        const uint32_t idesc = umma_idesc_bf16(TC_BM, n16, false, false);
        const uint32_t q_addr = smem_u32(sQ + (h % TC_QBUF) * TC_TILE_BYTES);
        const uint32_t k_addr = smem_u32(sK + st * TC_KV_BYTES);
#pragma unroll
        for (int k = 0; k < 4; ++k)   // dh = 64 = 4 x 16
          umma_bf16(tmem + TMEM_S, umma_smem_desc(q_addr + k * 32, 16, 1024), umma_smem_desc(k_addr + k * 32, 16, 1024), idesc, k > 0);
        umma_commit(s_full);
      };
      issue_s(0, 0, 0);
      for (int h = 0; h < nh; ++h) {
        for (int j = 0; j < kb.nb; ++j, ++g) {
          const int st = g & 1;
          int64_t row; int nvalid;
          kb.get(j, row, nvalid);
          mbar_wait(p_full, g & 1);     // softmax wrote P and is done reading S
          if (j == 0 && h > 0) mbar_wait(o_empty, (h - 1) & 1);   // previous head's O has been read out
          tc_fence_after();
          const uint32_t v_addr = smem_u32(sV + st * TC_KV_BYTES);
          const int ksteps = (nvalid + 15) >> 4;
          for (int k = 0; k < ksteps; ++k)   // 16 keys per step: P advances 8 packed columns, V 16 rows of 128 B
            umma_bf16_ts(tmem + TMEM_O, tmem + TMEM_P + k * 8, umma_smem_desc(v_addr + k * 2048, 8192, 1024), idesc_pv, (j > 0) || (k > 0));
          umma_commit(&kv_empty[st]);          // K/V stage consumed
          if (j + 1 < kb.nb) {
            issue_s(h, j + 1, g + 1);          // in-order MMA pipe: S(next) completes after this P.V
          } else {
            umma_commit(o_full);
            umma_commit(&q_empty[h % TC_QBUF]);
            if (h + 1 < nh) issue_s(h + 1, 0, g + 1);
          }
        }
      }
    }
  } else {
    // ------------------------------ softmax / output warps (2..5) ------------------------------
    const int quarter = warp & 3;
    const int row_in_tile = quarter * 32 + lane;
    const uint32_t lane_addr = tmem + ((uint32_t)(quarter * 32) << 16);
    const int i = r0 + row_in_tile;
    uint32_t raw[32];
    uint32_t sreg[64];
    int g = 0;
#ifdef MMF_ATTN_CLOCKS
    const bool dbg_on = (b == 3) && (r0 == p.n_head) && warp == 2 && lane == 0;   // first fusion tile of sample 3
    if (dbg_on) g_attn_clk[15] = (unsigned long long)kb.nb * nh;
    long long t_last = clock64();
#endif
    for (int h = 0; h < nh; ++h) {
      float m_ref = -INFINITY, l = 0.f;
      for (int j = 0; j < kb.nb; ++j, ++g) {
        int64_t row; int nvalid;
        kb.get(j, row, nvalid);
        CLK(4, 0);
        mbar_wait(s_full, g & 1);
        tc_fence_after();
        CLK(0, 0);
        // ONE read of the S row (64 keys -> 64 registers): TMEM reads run at only ~64 B/clk/SM, so a separate
        // max pass + exp pass over S would cost as much as the exponentials themselves
        tmem_ld_32x64(lane_addr + TMEM_S, sreg);
        tmem_wait_ld();
        const float mx = row_max64(sreg, nvalid);
        CLK(1, 0);
        const float m_new = mx * p.scale_log2;
        // lazy rescale: only when the max grew by more than 2^8 (warp-uniform decision: tcgen05.ld/st are warp-wide)
        const bool grow = m_new > m_ref + 8.0f;
        if (j == 0) {
          m_ref = m_new;
        } else if (__any_sync(0xffffffffu, grow)) {
          const float new_ref = grow ? m_new : m_ref;
          const float alpha = exp2f(m_ref - new_ref);
          m_ref = new_ref;
          l *= alpha;
#pragma unroll
          for (int c = 0; c < 2; ++c) {   // O row: 64 fp32 columns (the previous P.V has completed: S finished after it)
            tmem_ld_32x32(lane_addr + TMEM_O + c * 32, raw);
            tmem_wait_ld();
#pragma unroll
            for (int t = 0; t < 32; ++t) raw[t] = __float_as_uint(__uint_as_float(raw[t]) * alpha);
            tmem_st_32x32(lane_addr + TMEM_O + c * 32, raw);
          }
          tmem_wait_st();
        }
        // P = exp2(s*scale*log2e - m_ref) as packed bf16, written in place over the S columns (P.V reads ceil16(nvalid) keys)
        float rs;
        {
          uint32_t pk[32];
          rs = exp_pack64(sreg, pk, nvalid, p.scale_log2, m_ref);
          tmem_st_32x32(lane_addr + TMEM_P, pk);
        }
        l += rs;
        CLK(2, 0);
        tmem_wait_st();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(p_full);
        CLK(3, 0);
      }
      // ---- head epilogue: O / l -> bf16 row, log-sum-exp ----
      mbar_wait(o_full, h & 1);
      tc_fence_after();
      CLK(5, 0);
      const float inv = 1.0f / l;
      __nv_bfloat16* orow = p.o + (q_row0 + row_in_tile) * p.ldo + (h0 + h) * 64;
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        tmem_ld_32x32(lane_addr + TMEM_O + c * 32, raw);
        tmem_wait_ld();
        if (i < r1) {
#pragma unroll
          for (int q = 0; q < 4; ++q)
            *reinterpret_cast<uint4*>(orow + c * 32 + q * 8) = make_uint4(
                pack_bf16(__uint_as_float(raw[8 * q]) * inv, __uint_as_float(raw[8 * q + 1]) * inv),
                pack_bf16(__uint_as_float(raw[8 * q + 2]) * inv, __uint_as_float(raw[8 * q + 3]) * inv),
                pack_bf16(__uint_as_float(raw[8 * q + 4]) * inv, __uint_as_float(raw[8 * q + 5]) * inv),
                pack_bf16(__uint_as_float(raw[8 * q + 6]) * inv, __uint_as_float(raw[8 * q + 7]) * inv));
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(o_empty);   // the MMA warp may overwrite O for the next head
      if (p.lse && i < r1) p.lse[((int64_t)b * p.H + h0 + h) * p.N + i] = (m_ref + log2f(l)) * 0.6931471805599453f;
      CLK(6, 0);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem, TMEM_COLS);
  }
}


// ================================================================================================
// FORWARD, second generation (round 2): the same tile / TMEM / barrier skeleton, software-pipelined on HALF blocks.
//
// Round-1 measurements (ncu + clock instrumentation): tensor pipe 18 % active, MUFU 46 % busy, a
// fusion-tile CTA spent 1030 of its 3370 clk per 64-key block WAITING for the next S: per CTA the chain
//   S MMA -> commit -> tcgen05.ld -> max -> exp -> tcgen05.st -> arrive -> P.V MMA -> S MMA ...
// is strictly serial, so the only overlap came from the three CTAs sharing an SM.  Here a 64-key block is two 32-key
// halves A / B with their own S columns (A: 0..31, B: 32..63; P written in place) and their own s_full / p_full
// barriers.  The MMA thread issues   P.V_A(g), S_A(g+1), P.V_B(g), S_B(g+1), ...   so while the softmax warps work on
// half B of block g the tensor pipe accumulates half A and already computes half A of block g+1, and vice versa: in steady
// state neither side waits for the other, and the per-thread register footprint halves (32 S values instead of 64), which
// admits FOUR CTAs per SM (template CTAS; with a single Q buffer the shared memory fits too).
// A fraction of the exponentials (POLY of every 4) runs on the FMA pipe as Cody-Waite range reduction + a degree-3 minimax
// polynomial (max rel. error 7.5e-5, far below the bf16 rounding of P) to unload the 16-lane MUFU pipe that bounds the
// kernel (tools/micro/pipes.cu); warps whose 32 query rows are all padding (rows >= r1) skip the arithmetic.
// Every s_full / p_full barrier completes exactly once per block (a missing half B still commits / arrives), so all phase
// parities are simply (g & 1).
// ================================================================================================
constexpr uint32_t F2_SA = 0, F2_SB = 32, F2_O = 64;

// 2^x on the FMA / ALU pipes: x = n + f with n = round(x) (magic-number add), 2^f by a degree-3 minimax polynomial on
// [-0.5, 0.5], scaled by 2^n through the exponent field.  x <= ~+120; x < -125 saturates at 2^-125 (~0).
__device__ __forceinline__ float exp2_poly(float x) {
  x = fmaxf(x, -125.f);
  const float t = x + 12582912.f;          // 1.5 * 2^23: round(x) lands in the low mantissa bits
  const float f = x - (t - 12582912.f);
  float p = fmaf(f, 0.0551716685f, 0.2426111251f);
  p = fmaf(p, f, 0.6932609677f);
  p = fmaf(p, f, 0.9999280572f);
  return __int_as_float(__float_as_int(p) + (__float_as_int(t) << 23));
}

__device__ __forceinline__ float row_max32(const uint32_t (&v)[32], int nv) {
  float m0 = -INFINITY, m1 = -INFINITY;
  if (nv >= 32) {
#pragma unroll
    for (int t = 0; t < 16; ++t) { m0 = fmaxf(m0, __uint_as_float(v[2 * t])); m1 = fmaxf(m1, __uint_as_float(v[2 * t + 1])); }
  } else {
#pragma unroll
    for (int t = 0; t < 32; ++t)
      if (t < nv) m0 = fmaxf(m0, __uint_as_float(v[t]));
  }
  return fmaxf(m0, m1);
}

template <int POLY>
__device__ __forceinline__ float exp_pack32(const uint32_t (&v)[32], uint32_t (&pk)[16], int nv, float scale_log2, float m_ref) {
  float rs0 = 0.f, rs1 = 0.f;
#pragma unroll
  for (int t = 0; t < 16; ++t) {
    const float x0 = fmaf(__uint_as_float(v[2 * t]), scale_log2, -m_ref);
    const float x1 = fmaf(__uint_as_float(v[2 * t + 1]), scale_log2, -m_ref);
    float p0, p1;
    if (POLY == 8 || POLY == 10) {   // TIMING ABLATION ONLY (wrong results): no exponentials
      p0 = x0; p1 = x1;
    } else {
      p0 = ((2 * t) & 3) < POLY ? exp2_poly(x0) : ex2_approx(x0);
      p1 = ((2 * t + 1) & 3) < POLY ? exp2_poly(x1) : ex2_approx(x1);
    }
    if (nv < 32) {
      if (2 * t >= nv) p0 = 0.f;
      if (2 * t + 1 >= nv) p1 = 0.f;
    }
    rs0 += p0; rs1 += p1;
    pk[t] = pack_bf16(p0, p1);
  }
  return rs0 + rs1;
}

template <int CTAS, int QBUF, int KVST, int POLY>
__global__ void __launch_bounds__(TC_THREADS, CTAS)
attn_fwd_tc2_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_k,
                    const __grid_constant__ CUtensorMap tmap_v, const AttnTcParams p) {
  pdl_wait();   // launched through launch_pdl (common.cuh): nothing another kernel owns is touched before this
  extern __shared__ uint8_t smem_raw[];
  int r0 = 0, r1 = 0, k0 = 0, k1 = 0, b = 0, h0 = 0, h1 = 0;
  {
    int tile;
    lpt_tile(p.seg, p.nseg, p.tiles, p.B, p.H, tile, b, h0, h1);
    bool found = false;
    if (p.seg == nullptr) {
      r0 = tile * TC_BM; r1 = min(r0 + TC_BM, p.N); k0 = 0; k1 = p.N; found = r0 < p.N;
    } else {
      for (int s = 0; s < p.nseg; ++s) {
        const int a = p.seg[s], e = p.seg[s + 1];
        const int nt = (e - a + TC_BM - 1) / TC_BM;
        if (tile < nt) {
          r0 = a + tile * TC_BM; r1 = min(r0 + TC_BM, e);
          if (s == p.nseg - 1) { k0 = 0; k1 = p.N; } else { k0 = a; k1 = e; }
          found = true;
          break;
        }
        tile -= nt;
      }
    }
    if (!found) return;
  }
  const int nh = h1 - h0;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;                            // QBUF buffers
  uint8_t* sK = smem + QBUF * TC_TILE_BYTES;     // KVST stages of 64 keys
  uint8_t* sV = sK + KVST * TC_KV_BYTES;         // KVST stages
  uint64_t* bars = reinterpret_cast<uint64_t*>(sV + KVST * TC_KV_BYTES);
  uint64_t* q_full = bars;         // [2]
  uint64_t* q_empty = bars + 2;    // [2]
  uint64_t* s_full = bars + 4;     // [2] per half
  uint64_t* p_full = bars + 6;     // [2] per half
  uint64_t* o_full = bars + 8;
  uint64_t* o_empty = bars + 9;
  uint64_t* kv_full = bars + 10;   // [KVST]
  uint64_t* kv_empty = kv_full + KVST;   // [KVST]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(kv_empty + KVST);

  const int64_t q_row0 = r0 < p.n_head ? (int64_t)b * p.n_head + r0 : p.head_rows + (int64_t)b * p.n_tail + (r0 - p.n_head);
  KeyBlocks kb;
  kb.init(k0, k1, p.n_head, p.n_tail, p.head_rows, b);

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_q);
    tma_prefetch_desc(&tmap_k);
    tma_prefetch_desc(&tmap_v);
  }
  if (warp == 1) {
    if (lane == 0) {
      for (int s = 0; s < 2; ++s) {
        mbar_init(&q_full[s], 1); mbar_init(&q_empty[s], 1);
        mbar_init(&s_full[s], 1); mbar_init(&p_full[s], 4);
      }
      for (int s = 0; s < KVST; ++s) { mbar_init(&kv_full[s], 1); mbar_init(&kv_empty[s], 1); }
      mbar_init(o_full, 1);
      mbar_init(o_empty, 4);
      mbar_fence_init();
    }
    __syncwarp();
    tmem_alloc(tmem_slot, TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const int total = kb.nb * nh;      // blocks this CTA processes, g = h * kb.nb + j

  if (warp == 0) {
    // ------------------------------ TMA producer ------------------------------
    if (lane == 0) {
      // L2 prefetch of the operands p.pf blocks / heads ahead of the shared-memory loads (the ring holds only KVST key
      // blocks, i.e. KVST - 1 loads in flight, against a global -> shared latency of a few thousand cycles)
      auto prefetch_kv = [&](int gg) {
        if (gg >= total) return;
        const int hh = gg / kb.nb, jj = gg - hh * kb.nb;
        int64_t row; int nvalid;
        kb.get(jj, row, nvalid);
        tma_prefetch_2d(&tmap_k, (h0 + hh) * 64, (int)row);
        tma_prefetch_2d(&tmap_v, (h0 + hh) * 64, (int)row);
      };
      if (p.pf > 0) {
        for (int gg = 0; gg < p.pf; ++gg) prefetch_kv(gg);
        for (int hh = 1; hh < nh && hh <= 2; ++hh) tma_prefetch_2d(&tmap_q, (h0 + hh) * 64, (int)q_row0);
      }
      int g = 0;
      for (int h = 0; h < nh; ++h) {
        const int qs = h % QBUF;
        if (p.pf > 0 && h + 3 < nh) tma_prefetch_2d(&tmap_q, (h0 + h + 3) * 64, (int)q_row0);
        mbar_wait(&q_empty[qs], ((h / QBUF) & 1) ^ 1);
        mbar_expect_tx(&q_full[qs], TC_TILE_BYTES);
        tma_load_2d(sQ + qs * TC_TILE_BYTES, &tmap_q, &q_full[qs], (h0 + h) * 64, (int)q_row0);
        for (int j = 0; j < kb.nb; ++j, ++g) {
          const int st = g % KVST;
          if (p.pf > 0) prefetch_kv(g + p.pf);
          mbar_wait(&kv_empty[st], ((g / KVST) & 1) ^ 1);
          int64_t row; int nvalid;
          kb.get(j, row, nvalid);
          mbar_expect_tx(&kv_full[st], 2 * TC_KV_BYTES);
          tma_load_2d(sK + st * TC_KV_BYTES, &tmap_k, &kv_full[st], (h0 + h) * 64, (int)row);
          tma_load_2d(sV + st * TC_KV_BYTES, &tmap_v, &kv_full[st], (h0 + h) * 64, (int)row);
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------ MMA issuer ------------------------------
    // The WHOLE warp runs this loop and one elected lane issues the tcgen05 instructions.  Round 1 wrapped the loop in
    // `if (lane == 0)`: inside a divergent region nothing is provably warp-uniform, so ptxas moved every descriptor to
    // the uniform datapath through an R2UR + vote loop (~17 SASS instructions per UTCHMMA, ~400 dependent
    // single-thread instructions per key block) -- and THAT scalar stream, not MUFU / TMEM / TMA, paced the kernel: a
    // lone CTA took ~3200 clk per 64-key block whatever the softmax did (tools/attn_ab.py ablations).
    // Here every operand is computed in converged code from uniform sources (shuffle-broadcast where the compiler cannot
    // see it), the descriptors are running sums, and (h, j) advance incrementally instead of through divisions.
    if (total > 0) {
      const bool leader = elect_one();
      const uint32_t tm = __shfl_sync(0xffffffffu, tmem, 0);
      const uint32_t idesc_pv = umma_idesc_bf16(TC_BM, 64, false, true);   // A = P (TMEM, K-major), B = V (MN-major)
      const uint32_t idesc_s0 = umma_idesc_bf16(TC_BM, 0, false, false);   // N is added per half
      const uint64_t dq0 = umma_smem_desc(__shfl_sync(0xffffffffu, smem_u32(sQ), 0), 16, 1024);
      const uint64_t dk0 = umma_smem_desc(__shfl_sync(0xffffffffu, smem_u32(sK), 0), 16, 1024);
      const uint64_t dv0 = umma_smem_desc(__shfl_sync(0xffffffffu, smem_u32(sV), 0), 8192, 1024);
      const int nb = __shfl_sync(0xffffffffu, kb.nb, 0);
      // (h, j, stage, stage parity) of the NEXT S to issue
      int sh = 0, sj = 0, sst = 0, sph = 0;
      auto nvalid_of = [&](int j) { int64_t row; int nv; kb.get(j, row, nv); return nv; };
      auto issue_s = [&](int hf) {      // half hf of block (sh, sj); hf == 1 also advances the cursor
        if (hf == 0) {
          if (sj == 0) mbar_wait(&q_full[sh % QBUF], (sh / QBUF) & 1);
          mbar_wait(&kv_full[sst], sph);
          tc_fence_after();
        }
        const int nvh = min(32, nvalid_of(sj) - 32 * hf);
        if (nvh > 0 && leader) {
          const uint32_t idesc = idesc_s0 | ((uint32_t)((nvh + 15) >> 4) << 18);     // n_dim = n16 >> 3 at bit 17
          const uint64_t dq = dq0 + (uint64_t)((sh % QBUF) * (TC_TILE_BYTES >> 4));
          const uint64_t dk = dk0 + (uint64_t)(sst * (TC_KV_BYTES >> 4) + hf * (4096 >> 4));
#pragma unroll
          for (int k = 0; k < 4; ++k)   // dh = 64 = 4 x 16: 32 bytes per step
            umma_bf16(tm + (hf ? F2_SB : F2_SA), dq + 2 * k, dk + 2 * k, idesc, k > 0);
        }
        if (leader) {
          umma_commit(&s_full[hf]);     // completes once per block and half, whether or not the half has keys
          // the head's Q tile has no reader after the last block's half-B S
          if (hf == 1 && sj + 1 == nb) umma_commit(&q_empty[sh % QBUF]);
        }
        __syncwarp();
        if (hf == 1) {
          if (++sj == nb) { sj = 0; ++sh; }
          if (++sst == KVST) { sst = 0; sph ^= 1; }
        }
      };
#ifdef MMF_ATTN_CLOCKS
      const bool dbg_on = (b == 3) && (r0 == p.n_head) && lane == 0;   // first fusion tile of sample 3
      unsigned t_last = clock();
      unsigned acc_clk[16] = {0};
#endif
      issue_s(0);
      issue_s(1);
      int h = 0, j = 0, st = 0;
      for (int g = 0; g < total; ++g) {
        const int nvalid = nvalid_of(j);
        const uint64_t dv = dv0 + (uint64_t)(st * (TC_KV_BYTES >> 4));
#pragma unroll
        for (int hf = 0; hf < 2; ++hf) {
          CLK2(8, 0);
          mbar_wait(&p_full[hf], g & 1);      // the softmax warps wrote P of this half (and are done reading its S)
          CLK2(9 + hf, 0);
          if (j == 0 && hf == 0 && h > 0) mbar_wait(o_empty, (h - 1) & 1);   // previous head's O has been read out
          tc_fence_after();
          CLK2(11, 0);
          const int nvh = min(32, nvalid - 32 * hf);
          if (leader) {
            const int ksteps = nvh > 0 ? (nvh + 15) >> 4 : 0;
            for (int k = 0; k < ksteps; ++k)   // 16 keys per step: P advances 8 packed columns, V 16 rows of 128 B
              umma_bf16_ts(tm + F2_O, tm + (hf ? F2_SB : F2_SA) + k * 8, dv + (uint64_t)((2 * hf + k) * (2048 >> 4)), idesc_pv,
                           (j > 0) || (hf > 0) || (k > 0));
            if (hf == 1) {
              umma_commit(&kv_empty[st]);        // every MMA reading this K/V stage has been issued
              if (j + 1 == nb) umma_commit(o_full);
            }
          }
          __syncwarp();
          CLK2(12, 0);
          // the half's S columns are free again (in-order tensor pipe: S(next) executes after this P.V)
          if (g + 1 < total) issue_s(hf);
          else if (leader) umma_commit(&s_full[hf]);       // final phase flip: lets a rescale at the last block wait for this P.V
        }
        if (++j == nb) { j = 0; ++h; }
        if (++st == KVST) st = 0;
      }
#ifdef MMF_ATTN_CLOCKS
      if (dbg_on) for (int i = 8; i < 13; ++i) g_attn_clk[i] = acc_clk[i];
#endif
    }
  } else {
    // ------------------------------ softmax / output warps (2..5) ------------------------------
    const int quarter = warp & 3;
    const int row_in_tile = quarter * 32 + lane;
    const uint32_t lane_addr = tmem + ((uint32_t)(quarter * 32) << 16);
    const int i = r0 + row_in_tile;
    const bool warp_live = r0 + quarter * 32 < r1;     // any real query row in this warp?
    uint32_t sreg[32];
    int g = 0;
#ifdef MMF_ATTN_CLOCKS
    const bool dbg_on = (b == 3) && (r0 == p.n_head) && warp == 2 && lane == 0;   // first fusion tile of sample 3
    if (dbg_on) g_attn_clk[15] = (unsigned long long)kb.nb * nh;
    unsigned t_last = clock();
    unsigned acc_clk[16] = {0};
#endif
    for (int h = 0; h < nh; ++h) {
      float m_ref = -INFINITY, l = 0.f;
      for (int j = 0; j < kb.nb; ++j, ++g) {
        int64_t row; int nvalid;
        kb.get(j, row, nvalid);
#pragma unroll
        for (int hf = 0; hf < 2; ++hf) {
          const int nvh = min(32, nvalid - 32 * hf);
          CLK2(4, 0);
          mbar_wait(&s_full[hf], g & 1);
          CLK2(0, 0);
          if (warp_live && nvh > 0) {
            tc_fence_after();
            if (POLY == 9 || POLY == 10) {   // TIMING ABLATION ONLY (wrong results): no TMEM read of S
#pragma unroll
              for (int t = 0; t < 32; ++t) sreg[t] = __float_as_uint((float)(t + lane) * 0.01f);
            } else {
              tmem_ld_32x32(lane_addr + (hf ? F2_SB : F2_SA), sreg);
              tmem_wait_ld();
            }
            const float m_new = row_max32(sreg, nvh) * p.scale_log2;
            CLK2(1, 0);
            const bool grow = m_new > m_ref + 8.0f;
            if (j == 0 && hf == 0) {
              m_ref = m_new;
            } else if (__any_sync(0xffffffffu, grow)) {
              // lazy rescale of O.  The P.V of the PREVIOUS half may still be accumulating: it is complete once the
              // other half's s_full has flipped for (hf == 0: this block; hf == 1: the next block / the final flip)
              mbar_wait(&s_full[hf ^ 1], hf == 0 ? (g & 1) : ((g + 1) & 1));
              tc_fence_after();
              const float new_ref = grow ? m_new : m_ref;
              const float alpha = exp2f(m_ref - new_ref);
              m_ref = new_ref;
              l *= alpha;
              // (rare path; 16 columns at a time so that it does not push the S row out of the register file)
#pragma unroll 1
              for (int c = 0; c < 4; ++c) {
                uint32_t raw[16];
                tmem_ld_32x16(lane_addr + F2_O + c * 16, raw);
                tmem_wait_ld();
#pragma unroll
                for (int t = 0; t < 16; ++t) raw[t] = __float_as_uint(__uint_as_float(raw[t]) * alpha);
                tmem_st_32x16(lane_addr + F2_O + c * 16, raw);
              }
              tmem_wait_st();
            }
            uint32_t pk[16];
            l += exp_pack32<POLY>(sreg, pk, nvh, p.scale_log2, m_ref);
            tmem_st_32x16(lane_addr + (hf ? F2_SB : F2_SA), pk);
            CLK2(2, 0);
            tmem_wait_st();
          }
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&p_full[hf]);
          CLK2(3, 0);
        }
      }
      // ---- head epilogue: O / l -> bf16 row, log-sum-exp ----
      mbar_wait(o_full, h & 1);
      tc_fence_after();
      CLK2(5, 0);
      if (warp_live) {
        const float inv = 1.0f / l;
        __nv_bfloat16* orow = p.o + (q_row0 + row_in_tile) * p.ldo + (h0 + h) * 64;
        uint32_t raw[32];
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          tmem_ld_32x32(lane_addr + F2_O + c * 32, raw);
          tmem_wait_ld();
          if (i < r1) {
#pragma unroll
            for (int q = 0; q < 4; ++q)
              *reinterpret_cast<uint4*>(orow + c * 32 + q * 8) = make_uint4(
                  pack_bf16(__uint_as_float(raw[8 * q]) * inv, __uint_as_float(raw[8 * q + 1]) * inv),
                  pack_bf16(__uint_as_float(raw[8 * q + 2]) * inv, __uint_as_float(raw[8 * q + 3]) * inv),
                  pack_bf16(__uint_as_float(raw[8 * q + 4]) * inv, __uint_as_float(raw[8 * q + 5]) * inv),
                  pack_bf16(__uint_as_float(raw[8 * q + 6]) * inv, __uint_as_float(raw[8 * q + 7]) * inv));
          }
        }
        if (p.lse && i < r1) p.lse[((int64_t)b * p.H + h0 + h) * p.N + i] = (m_ref + log2f(l)) * 0.6931471805599453f;
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(o_empty);   // the MMA warp may overwrite O for the next head
      CLK2(6, 0);
    }
#ifdef MMF_ATTN_CLOCKS
    if (dbg_on) for (int i = 0; i < 7; ++i) g_attn_clk[i] = acc_clk[i];
#endif
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem, TMEM_COLS);
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn tc_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(ptr);
  });
  return fn;
}
static int tc_make_tmap_box(CUtensorMap* map, const void* base, int64_t rows, int64_t cols, int64_t ld, int box_rows);

// returns -1000 if this problem is not eligible for the tcgen05 kernel (caller falls through to the generic kernel)
int attn_fwd_tc_launch(const MmfAttnArgs* a, cudaStream_t stream) {
  if (a->dh != 64 || a->Nq != a->Nk || a->n_head_q != a->n_head_k) return -1000;
  if ((a->ldq & 7) || (a->ldk & 7) || (a->ldv & 7) || (a->ldo & 7)) return -1000;
  const int64_t rows = (int64_t)a->B * a->Nq;
  CUtensorMap tq, tk, tv;
  int rc;
  if ((rc = tc_make_tmap_box(&tq, a->q, rows, (int64_t)a->H * 64, a->ldq, TC_BM))) return rc;
  if ((rc = tc_make_tmap_box(&tk, a->k, rows, (int64_t)a->H * 64, a->ldk, TC_BN))) return rc;
  if ((rc = tc_make_tmap_box(&tv, a->v, rows, (int64_t)a->H * 64, a->ldv, TC_BN))) return rc;
  AttnTcParams p;
  p.o = reinterpret_cast<__nv_bfloat16*>(a->o); p.lse = a->lse; p.ldo = a->ldo;
  p.B = a->B; p.H = a->H; p.N = a->Nq; p.n_head = a->n_head_q; p.n_tail = a->n_tail_q;
  p.head_rows = (int64_t)a->B * a->n_head_q;
  p.scale_log2 = a->scale * 1.4426950408889634f;
  p.seg = a->seg; p.nseg = a->nseg;
  {
    const char* pe = getenv("MMF_ATTN_PF");
    p.pf = pe ? atoi(pe) : TC_PF_DEFAULT;
    if (p.pf < 0 || p.pf > 64) p.pf = TC_PF_DEFAULT;
  }
  const int tiles = (a->Nq + TC_BM - 1) / TC_BM + (a->seg ? a->nseg : 0);
  p.tiles = tiles;
  const int heavy_max = a->seg ? (a->Nq + TC_BM - 1) / TC_BM : 0;
  const int grid = (tiles + (TC_HSPLIT - 1) * heavy_max) * a->B;
  // MMF_ATTN_FWD selects the kernel generation / configuration for same-box A/B runs (default: the best measured):
  //   0 = round-1 kernel (64-key blocks, 3 CTAs/SM);  v2 (half-block pipeline): 1 = 3 CTAs/SM, 2 = 4 CTAs/SM,
  //   3 / 4 = the same with 1 of 4 exponentials on the FMA pipe, 5 / 6 = 2 of 4
  const char* env = getenv("MMF_ATTN_FWD");      // read per call: a bench process may switch variants between launches
  int variant = env ? atoi(env) : TC_FWD_DEFAULT;
  if (variant < 0 || variant > 9) variant = TC_FWD_DEFAULT;
  static std::atomic<unsigned> attr_done_v[10];   // per variant: bit d set = function attributes applied on device d
  std::atomic<unsigned>& attr_done = attr_done_v[variant];
  int dev = 0;
  cudaGetDevice(&dev);
  const unsigned bit = 1u << (dev & 31);
  auto prep = [&](const void* fn, int smem) -> int {
    if (attr_done.load(std::memory_order_acquire) & bit) return 0;
    cudaError_t e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return (int)e;
    // the default carve-out is sized for fewer CTAs: ask for the whole shared memory so that all resident CTAs fit
    cudaFuncSetAttribute(fn, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    return 0;
  };
#define MMF_FWD2(CT, QB, KS, PL)                                                                                       \
  do {                                                                                                                 \
    constexpr int smem = QB * TC_TILE_BYTES + 2 * KS * TC_KV_BYTES + 1024 + 512;                                       \
    if ((rc = prep(reinterpret_cast<const void*>(attn_fwd_tc2_kernel<CT, QB, KS, PL>), smem))) return rc;              \
    launch_pdl(attn_fwd_tc2_kernel<CT, QB, KS, PL>, dim3(grid), dim3(TC_THREADS), smem, stream, tq, tk, tv, p);                            \
  } while (0)
  switch (variant) {
    case 0:
      if ((rc = prep(reinterpret_cast<const void*>(attn_fwd_tc_kernel), TC_SMEM))) return rc;
      attn_fwd_tc_kernel<<<grid, TC_THREADS, TC_SMEM, stream>>>(tq, tk, tv, p);
      break;
    case 2: MMF_FWD2(4, 1, 2, 0); break;
    case 3: MMF_FWD2(3, 1, 3, 0); break;    // deeper K/V rings: 3 CTAs x 3 stages, 2 CTAs x 5 stages, 1 CTA x 12 stages
    case 4: MMF_FWD2(2, 2, 5, 0); break;
    case 5: MMF_FWD2(1, 2, 12, 0); break;
    case 6: MMF_FWD2(2, 2, 5, 1); break;    // + 1 of 4 exponentials on the FMA pipe
    case 7: MMF_FWD2(3, 2, 2, 8); break;    // 7 .. 9: timing ablations (WRONG RESULTS): no exp / no TMEM read of S / neither
    case 8: MMF_FWD2(3, 2, 2, 9); break;
    case 9: MMF_FWD2(3, 2, 2, 10); break;
    default: MMF_FWD2(3, 2, 2, 0); break;
  }
#undef MMF_FWD2
  attr_done.fetch_or(bit, std::memory_order_release);
  g_launch_count.fetch_add(1, std::memory_order_relaxed);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? 0 : (int)e;
}


// ================================================================================================
// BACKWARD on tcgen05 (dh = 64).  Two kernels on the same skeleton as the forward (TMA producer warp, one MMA
// issuing thread, four warps of one-thread-per-TMEM-lane elementwise work):
//   dQ  kernel: CTA = 128 query rows, loops key blocks of 64:   S = Q.K^T, dP = dO.V^T  ->  dS = P*(dP - delta)
//               (bf16, back into TMEM over S)  ->  dQ += dS.K   (A from TMEM, K read MN-major)
//   dKV kernel: CTA = 128 keys, loops query blocks of 64 (own segment + fusion rows):  S^T = K.Q^T, dP^T = V.dO^T
//               ->  P^T, dS^T (bf16 into TMEM)  ->  dV += P^T.dO,  dK += dS^T.Q   (dO / Q read MN-major)
// P is recomputed from the saved log-sum-exp: P = exp2(s*scale*log2e - lse*log2e).
// TMEM: 256 columns per CTA in both kernels, two CTAs per SM.
// ================================================================================================
constexpr int BW_BLK = 64;                        // rows of the inner (streamed) operand per block
constexpr int BW_BLK_BYTES = BW_BLK * 64 * 2;     // 64 x 64 bf16 tile = 8 KB

// blocks of BW_BLK token indices over up to two index ranges (each range lies inside one plane)
struct RowBlocks {
  int a0, e0, a1, e1, nb0, nb;   // scalars, not arrays: a runtime index would put them in local memory
  int n_head, n_tail, b;
  int64_t head_rows;
  __device__ void init(int a0_, int e0_, int a1_, int e1_, int n_head_, int n_tail_, int64_t head_rows_, int b_) {
    n_head = n_head_; n_tail = n_tail_; head_rows = head_rows_; b = b_;
    a0 = a0_; e0 = max(e0_, a0_); a1 = a1_; e1 = max(e1_, a1_);
    nb0 = (e0 - a0 + BW_BLK - 1) / BW_BLK;
    nb = nb0 + (e1 - a1 + BW_BLK - 1) / BW_BLK;
  }
  // block j -> first token index, global row of that token, number of valid rows
  __device__ void get(int j, int& tok, int64_t& row, int& nvalid) const {
    const bool second = j >= nb0;
    tok = second ? a1 + (j - nb0) * BW_BLK : a0 + j * BW_BLK;
    nvalid = min(BW_BLK, (second ? e1 : e0) - tok);
    row = tok < n_head ? (int64_t)b * n_head + tok : head_rows + (int64_t)b * n_tail + (tok - n_head);
  }
};

struct AttnBwdTcParams {
  const float* lse;      // [B, H, N] natural log
  const float* delta;    // [B, H, N]
  __nv_bfloat16* dq; __nv_bfloat16* dk; __nv_bfloat16* dv;
  int64_t lddq, lddk, lddv;
  int B, H, N, n_head, n_tail;
  int64_t head_rows;
  float scale, scale_log2;
  const int32_t* seg;
  int nseg;
  int tiles;             // tiles per sample the grids were sized for (upper bound)
  int pf;                // L2 prefetch distance in streamed blocks (0 = off)
};

__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// ---------------------------------------- dQ ----------------------------------------
constexpr uint32_t DQ_S = 0, DQ_DP = 64, DQ_ACC = 128, DQ_DS = 192;   // S | dP | dQ accumulator | dS (packed bf16)
constexpr int DQ_SMEM = 4 * TC_TILE_BYTES + 4 * BW_BLK_BYTES + 1024 + 256;   // 2x(Q,dO) + 2x(K,V)

__global__ void __launch_bounds__(TC_THREADS, 2)
attn_bwd_dq_tc_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_do,
                      const __grid_constant__ CUtensorMap tmap_k, const __grid_constant__ CUtensorMap tmap_v,
                      const AttnBwdTcParams p) {
  extern __shared__ uint8_t smem_raw[];
  int r0 = 0, r1 = 0, k0 = 0, k1 = 0, b = 0, h0 = 0, h1 = 0;
  {
    int tile;
    lpt_tile(p.seg, p.nseg, p.tiles, p.B, p.H, tile, b, h0, h1);
    bool found = false;
    if (p.seg == nullptr) {
      r0 = tile * TC_BM; r1 = min(r0 + TC_BM, p.N); k0 = 0; k1 = p.N; found = r0 < p.N;
    } else {
      for (int s = 0; s < p.nseg; ++s) {
        const int a = p.seg[s], e = p.seg[s + 1];
        const int nt = (e - a + TC_BM - 1) / TC_BM;
        if (tile < nt) {
          r0 = a + tile * TC_BM; r1 = min(r0 + TC_BM, e);
          if (s == p.nseg - 1) { k0 = 0; k1 = p.N; } else { k0 = a; k1 = e; }
          found = true;
          break;
        }
        tile -= nt;
      }
    }
    if (!found) return;
  }
  const int nh = h1 - h0;   // this CTA's heads are h0 .. h1-1; h below counts from h0
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;                                   // 2 buffers of 128x64
  uint8_t* sdO = smem + 2 * TC_TILE_BYTES;              // 2 buffers
  uint8_t* sK = smem + 4 * TC_TILE_BYTES;               // 2 stages of 64x64
  uint8_t* sV = sK + 2 * BW_BLK_BYTES;                  // 2 stages
  uint64_t* bars = reinterpret_cast<uint64_t*>(sV + 2 * BW_BLK_BYTES);
  uint64_t* q_full = bars;         // [2] (Q and dO of a head)
  uint64_t* q_empty = bars + 2;    // [2]
  uint64_t* kv_full = bars + 4;    // [2]
  uint64_t* kv_empty = bars + 6;   // [2]
  uint64_t* s_full = bars + 8;     // S and dP ready
  uint64_t* ds_full = bars + 9;    // dS written (and S / dP consumed)
  uint64_t* acc_full = bars + 10;  // dQ of the head complete
  uint64_t* acc_empty = bars + 11; // dQ read out
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 12);

  const int64_t q_row0 = r0 < p.n_head ? (int64_t)b * p.n_head + r0 : p.head_rows + (int64_t)b * p.n_tail + (r0 - p.n_head);
  RowBlocks kb;   // key range split at the plane boundary
  kb.init(k0, min(k1, p.n_head), max(k0, p.n_head), k1, p.n_head, p.n_tail, p.head_rows, b);

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_q); tma_prefetch_desc(&tmap_do); tma_prefetch_desc(&tmap_k); tma_prefetch_desc(&tmap_v);
  }
  if (warp == 1) {
    if (lane == 0) {
      for (int s = 0; s < 2; ++s) {
        mbar_init(&q_full[s], 1); mbar_init(&q_empty[s], 1);
        mbar_init(&kv_full[s], 1); mbar_init(&kv_empty[s], 1);
      }
      mbar_init(s_full, 1); mbar_init(ds_full, 4); mbar_init(acc_full, 1); mbar_init(acc_empty, 4);
      mbar_fence_init();
    }
    __syncwarp();
    tmem_alloc(tmem_slot, 256);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      int g = 0;
      for (int h = 0; h < nh; ++h) {
        const int qs = h & 1;
        mbar_wait(&q_empty[qs], ((h >> 1) & 1) ^ 1);
        mbar_expect_tx(&q_full[qs], 2 * TC_TILE_BYTES);
        tma_load_2d(sQ + qs * TC_TILE_BYTES, &tmap_q, &q_full[qs], (h0 + h) * 64, (int)q_row0);
        tma_load_2d(sdO + qs * TC_TILE_BYTES, &tmap_do, &q_full[qs], (h0 + h) * 64, (int)q_row0);
        for (int j = 0; j < kb.nb; ++j, ++g) {
          const int st = g & 1;
          mbar_wait(&kv_empty[st], ((g >> 1) & 1) ^ 1);
          int tok, nvalid; int64_t row;
          kb.get(j, tok, row, nvalid);
          mbar_expect_tx(&kv_full[st], 2 * BW_BLK_BYTES);
          tma_load_2d(sK + st * BW_BLK_BYTES, &tmap_k, &kv_full[st], (h0 + h) * 64, (int)row);
          tma_load_2d(sV + st * BW_BLK_BYTES, &tmap_v, &kv_full[st], (h0 + h) * 64, (int)row);
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc_s = umma_idesc_bf16(TC_BM, BW_BLK, false, false);   // [128 x 64] = A[128 x 64dh] . B[64 keys x 64dh]^T
      const uint32_t idesc_acc = umma_idesc_bf16(TC_BM, 64, false, true);      // dQ[128 x 64dh] += dS[128 x keys] . K (MN-major)
      int g = 0;
      auto issue_s = [&](int h, int j, int gg) {
        const int st = gg & 1;
        if (j == 0) mbar_wait(&q_full[h & 1], (h >> 1) & 1);
        mbar_wait(&kv_full[st], (gg >> 1) & 1);
        tc_fence_after();
        const uint32_t q_addr = smem_u32(sQ + (h & 1) * TC_TILE_BYTES), do_addr = smem_u32(sdO + (h & 1) * TC_TILE_BYTES);
        const uint32_t k_addr = smem_u32(sK + st * BW_BLK_BYTES), v_addr = smem_u32(sV + st * BW_BLK_BYTES);
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_bf16(tmem + DQ_S, umma_smem_desc(q_addr + k * 32, 16, 1024), umma_smem_desc(k_addr + k * 32, 16, 1024), idesc_s, k > 0);
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_bf16(tmem + DQ_DP, umma_smem_desc(do_addr + k * 32, 16, 1024), umma_smem_desc(v_addr + k * 32, 16, 1024), idesc_s, k > 0);
        umma_commit(s_full);
      };
      issue_s(0, 0, 0);
      for (int h = 0; h < nh; ++h) {
        for (int j = 0; j < kb.nb; ++j, ++g) {
          const int st = g & 1;
          int tok, nvalid; int64_t row;
          kb.get(j, tok, row, nvalid);
          mbar_wait(ds_full, g & 1);
          if (j == 0 && h > 0) mbar_wait(acc_empty, (h - 1) & 1);
          tc_fence_after();
          const uint32_t k_addr = smem_u32(sK + st * BW_BLK_BYTES);
          const int ksteps = (nvalid + 15) >> 4;
          for (int k = 0; k < ksteps; ++k)
            umma_bf16_ts(tmem + DQ_ACC, tmem + DQ_DS + k * 8, umma_smem_desc(k_addr + k * 2048, 8192, 1024), idesc_acc, (j > 0) || (k > 0));
          umma_commit(&kv_empty[st]);
          if (j + 1 < kb.nb) {
            issue_s(h, j + 1, g + 1);
          } else {
            umma_commit(acc_full);
            umma_commit(&q_empty[h & 1]);
            if (h + 1 < nh) issue_s(h + 1, 0, g + 1);
          }
        }
      }
    }
  } else {
    const int quarter = warp & 3;
    const int row_in_tile = quarter * 32 + lane;
    const uint32_t lane_addr = tmem + ((uint32_t)(quarter * 32) << 16);
    const int i = r0 + row_in_tile;
    const bool row_ok = i < r1;
    uint32_t rs[32], rd[32];
    int g = 0;
    // this row's lse / delta of the NEXT head are requested while the current head is processed (fetched at the head's
    // start they cost a full memory round trip per head, i.e. per 2 key blocks on a modality tile)
    const int64_t stat0 = ((int64_t)b * p.H + h0) * p.N + (row_ok ? i : r0);
    float lse_raw = p.lse[stat0], dl_raw = p.delta[stat0];
    for (int h = 0; h < nh; ++h) {
      const float lse2 = row_ok ? lse_raw * 1.4426950408889634f : INFINITY;   // invalid rows -> P = 0
      const float dl = row_ok ? dl_raw : 0.f;
      if (h + 1 < nh) {
        lse_raw = p.lse[stat0 + (int64_t)(h + 1) * p.N];
        dl_raw = p.delta[stat0 + (int64_t)(h + 1) * p.N];
      }
      for (int j = 0; j < kb.nb; ++j, ++g) {
        int tok, nvalid; int64_t row;
        kb.get(j, tok, row, nvalid);
        mbar_wait(s_full, g & 1);
        tc_fence_after();
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          tmem_ld_32x32(lane_addr + DQ_S + c * 32, rs);
          tmem_ld_32x32(lane_addr + DQ_DP + c * 32, rd);
          tmem_wait_ld();
          uint32_t pk[16];
#pragma unroll
          for (int t = 0; t < 16; ++t) {
            const int c0 = c * 32 + 2 * t;
            const float p0 = c0 < nvalid ? ex2(fmaf(__uint_as_float(rs[2 * t]), p.scale_log2, -lse2)) : 0.f;
            const float p1 = c0 + 1 < nvalid ? ex2(fmaf(__uint_as_float(rs[2 * t + 1]), p.scale_log2, -lse2)) : 0.f;
            pk[t] = pack_bf16(p0 * (__uint_as_float(rd[2 * t]) - dl) * p.scale, p1 * (__uint_as_float(rd[2 * t + 1]) - dl) * p.scale);
          }
          tmem_st_32x16(lane_addr + DQ_DS + c * 16, pk);
        }
        tmem_wait_st();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(ds_full);
      }
      mbar_wait(acc_full, h & 1);
      tc_fence_after();
      __nv_bfloat16* orow = p.dq + (q_row0 + row_in_tile) * p.lddq + (h0 + h) * 64;
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        tmem_ld_32x32(lane_addr + DQ_ACC + c * 32, rs);
        tmem_wait_ld();
        if (row_ok) {
#pragma unroll
          for (int q = 0; q < 4; ++q)
            *reinterpret_cast<uint4*>(orow + c * 32 + q * 8) = make_uint4(
                pack_bf16(__uint_as_float(rs[8 * q]), __uint_as_float(rs[8 * q + 1])),
                pack_bf16(__uint_as_float(rs[8 * q + 2]), __uint_as_float(rs[8 * q + 3])),
                pack_bf16(__uint_as_float(rs[8 * q + 4]), __uint_as_float(rs[8 * q + 5])),
                pack_bf16(__uint_as_float(rs[8 * q + 6]), __uint_as_float(rs[8 * q + 7])));
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(acc_empty);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem, 256);
  }
}

// ---------------------------------------- dQ, second generation ----------------------------------------
// The round-1 dQ kernel above is one serial chain per CTA (S, dP MMA -> elementwise -> dQ MMA -> next S ...; two CTAs per
// SM).  Same restructuring as the forward (attn_fwd_tc2_kernel): a 64-key block is two 32-key halves with their own
// S / dP columns and barriers, dS is written in place over the half's S columns, and the MMA thread issues
//   dQ += dS_A(g).K_A,  S_A / dP_A (g+1),  dQ += dS_B(g).K_B,  S_B / dP_B (g+1), ...
// so the tensor pipe works on one half while the elementwise warps work on the other.  Every barrier completes exactly
// once per block (phase parity = g & 1).  TMEM: S_A | S_B | dP_A | dP_B | dQ = 192 of 256 allocated columns.
constexpr uint32_t DQ2_HALF = 64, DQ2_DP_OFF = 32, DQ2_ACC = 128;   // half hf: S at 64 hf, dP at 64 hf + 32 (one 64-column TMEM read)

template <int KVST, int POLY>
__global__ void __launch_bounds__(TC_THREADS, 2)
attn_bwd_dq_tc2_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_do,
                       const __grid_constant__ CUtensorMap tmap_k, const __grid_constant__ CUtensorMap tmap_v,
                       const AttnBwdTcParams p) {
  pdl_wait();   // launched through launch_pdl (common.cuh): nothing another kernel owns is touched before this
  // 1024-byte aligned by declaration (the 128B-swizzle atoms need it): no alignment slack in the allocation, which is what
  // lets 2 CTAs x (64 KB of Q / dO + three 16 KB K / V stages) share an SM
  extern __shared__ __align__(1024) uint8_t smem_dq2[];
  int r0 = 0, r1 = 0, k0 = 0, k1 = 0, b = 0, h0 = 0, h1 = 0;
  {
    int tile;
    lpt_tile(p.seg, p.nseg, p.tiles, p.B, p.H, tile, b, h0, h1);
    bool found = false;
    if (p.seg == nullptr) {
      r0 = tile * TC_BM; r1 = min(r0 + TC_BM, p.N); k0 = 0; k1 = p.N; found = r0 < p.N;
    } else {
      for (int s = 0; s < p.nseg; ++s) {
        const int a = p.seg[s], e = p.seg[s + 1];
        const int nt = (e - a + TC_BM - 1) / TC_BM;
        if (tile < nt) {
          r0 = a + tile * TC_BM; r1 = min(r0 + TC_BM, e);
          if (s == p.nseg - 1) { k0 = 0; k1 = p.N; } else { k0 = a; k1 = e; }
          found = true;
          break;
        }
        tile -= nt;
      }
    }
    if (!found) return;
  }
  const int nh = h1 - h0;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint8_t* smem = smem_dq2;
  uint8_t* sQ = smem;                                   // 2 buffers of 128x64
  uint8_t* sdO = smem + 2 * TC_TILE_BYTES;              // 2 buffers
  uint8_t* sK = smem + 4 * TC_TILE_BYTES;               // KVST stages of 64x64
  uint8_t* sV = sK + KVST * BW_BLK_BYTES;               // KVST stages
  uint64_t* bars = reinterpret_cast<uint64_t*>(sV + KVST * BW_BLK_BYTES);
  uint64_t* q_full = bars;         // [2] (Q and dO of a head)
  uint64_t* q_empty = bars + 2;    // [2]
  uint64_t* s_full = bars + 4;     // [2] S and dP of a half ready
  uint64_t* ds_full = bars + 6;    // [2] dS of a half written (its S / dP consumed)
  uint64_t* acc_full = bars + 8;   // dQ of the head complete
  uint64_t* acc_empty = bars + 9;  // dQ read out
  uint64_t* kv_full = bars + 10;   // [KVST]
  uint64_t* kv_empty = kv_full + KVST;   // [KVST]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(kv_empty + KVST);

  const int64_t q_row0 = r0 < p.n_head ? (int64_t)b * p.n_head + r0 : p.head_rows + (int64_t)b * p.n_tail + (r0 - p.n_head);
  RowBlocks kb;   // key range split at the plane boundary
  kb.init(k0, min(k1, p.n_head), max(k0, p.n_head), k1, p.n_head, p.n_tail, p.head_rows, b);

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_q); tma_prefetch_desc(&tmap_do); tma_prefetch_desc(&tmap_k); tma_prefetch_desc(&tmap_v);
  }
  if (warp == 1) {
    if (lane == 0) {
      for (int s = 0; s < 2; ++s) {
        mbar_init(&q_full[s], 1); mbar_init(&q_empty[s], 1);
        mbar_init(&s_full[s], 1); mbar_init(&ds_full[s], 4);
      }
      for (int s = 0; s < KVST; ++s) { mbar_init(&kv_full[s], 1); mbar_init(&kv_empty[s], 1); }
      mbar_init(acc_full, 1); mbar_init(acc_empty, 4);
      mbar_fence_init();
    }
    __syncwarp();
    tmem_alloc(tmem_slot, 256);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const int total = kb.nb * nh;

  if (warp == 0) {
    if (lane == 0) {
      auto prefetch_kv = [&](int gg) {     // L2 prefetch ahead of the shared-memory ring (see attn_fwd_tc2_kernel)
        if (gg >= total) return;
        const int hh = gg / kb.nb, jj = gg - hh * kb.nb;
        int tok, nvalid; int64_t row;
        kb.get(jj, tok, row, nvalid);
        tma_prefetch_2d(&tmap_k, (h0 + hh) * 64, (int)row);
        tma_prefetch_2d(&tmap_v, (h0 + hh) * 64, (int)row);
      };
      if (p.pf > 0) {
        for (int gg = 0; gg < p.pf; ++gg) prefetch_kv(gg);
        for (int hh = 1; hh < nh && hh <= 2; ++hh) {
          tma_prefetch_2d(&tmap_q, (h0 + hh) * 64, (int)q_row0);
          tma_prefetch_2d(&tmap_do, (h0 + hh) * 64, (int)q_row0);
        }
      }
      int g = 0;
      for (int h = 0; h < nh; ++h) {
        const int qs = h & 1;
        if (p.pf > 0 && h + 3 < nh) {
          tma_prefetch_2d(&tmap_q, (h0 + h + 3) * 64, (int)q_row0);
          tma_prefetch_2d(&tmap_do, (h0 + h + 3) * 64, (int)q_row0);
        }
        mbar_wait(&q_empty[qs], ((h >> 1) & 1) ^ 1);
        mbar_expect_tx(&q_full[qs], 2 * TC_TILE_BYTES);
        tma_load_2d(sQ + qs * TC_TILE_BYTES, &tmap_q, &q_full[qs], (h0 + h) * 64, (int)q_row0);
        tma_load_2d(sdO + qs * TC_TILE_BYTES, &tmap_do, &q_full[qs], (h0 + h) * 64, (int)q_row0);
        for (int j = 0; j < kb.nb; ++j, ++g) {
          const int st = g % KVST;
          if (p.pf > 0) prefetch_kv(g + p.pf);
          mbar_wait(&kv_empty[st], ((g / KVST) & 1) ^ 1);
          int tok, nvalid; int64_t row;
          kb.get(j, tok, row, nvalid);
          mbar_expect_tx(&kv_full[st], 2 * BW_BLK_BYTES);
          tma_load_2d(sK + st * BW_BLK_BYTES, &tmap_k, &kv_full[st], (h0 + h) * 64, (int)row);
          tma_load_2d(sV + st * BW_BLK_BYTES, &tmap_v, &kv_full[st], (h0 + h) * 64, (int)row);
        }
      }
    }
  } else if (warp == 1) {
    // whole warp, one elected issuing lane, running-sum descriptors (see attn_fwd_tc2_kernel)
    if (total > 0) {
      const bool leader = elect_one();
      const uint32_t tm = __shfl_sync(0xffffffffu, tmem, 0);
      const uint32_t idesc_acc = umma_idesc_bf16(TC_BM, 64, false, true);      // dQ[128 x 64dh] += dS[128 x keys] . K (MN-major)
      const uint32_t idesc_s0 = umma_idesc_bf16(TC_BM, 0, false, false);
      const uint64_t dq0 = umma_smem_desc(__shfl_sync(0xffffffffu, smem_u32(sQ), 0), 16, 1024);
      const uint64_t ddo0 = umma_smem_desc(__shfl_sync(0xffffffffu, smem_u32(sdO), 0), 16, 1024);
      const uint64_t dk0 = umma_smem_desc(__shfl_sync(0xffffffffu, smem_u32(sK), 0), 16, 1024);
      const uint64_t dv0 = umma_smem_desc(__shfl_sync(0xffffffffu, smem_u32(sV), 0), 16, 1024);
      const uint64_t dkm0 = umma_smem_desc(__shfl_sync(0xffffffffu, smem_u32(sK), 0), 8192, 1024);   // K read MN-major for dQ
      const int nb = __shfl_sync(0xffffffffu, kb.nb, 0);
      int sh = 0, sj = 0, sst = 0, sph = 0;   // (h, j, stage, stage parity) of the NEXT S / dP to issue
      auto nvalid_of = [&](int j) { int tok, nv; int64_t row; kb.get(j, tok, row, nv); return nv; };
      auto issue_s = [&](int hf) {
        if (hf == 0) {
          if (sj == 0) mbar_wait(&q_full[sh & 1], (sh >> 1) & 1);
          mbar_wait(&kv_full[sst], sph);
          tc_fence_after();
        }
        const int nvh = min(32, nvalid_of(sj) - 32 * hf);
        if (nvh > 0 && leader) {
          const uint32_t idesc = idesc_s0 | ((uint32_t)((nvh + 15) >> 4) << 18);
          const uint64_t qd = dq0 + (uint64_t)((sh & 1) * (TC_TILE_BYTES >> 4)), od = ddo0 + (uint64_t)((sh & 1) * (TC_TILE_BYTES >> 4));
          const uint64_t kd = dk0 + (uint64_t)(sst * (BW_BLK_BYTES >> 4) + hf * (4096 >> 4));
          const uint64_t vd = dv0 + (uint64_t)(sst * (BW_BLK_BYTES >> 4) + hf * (4096 >> 4));
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_bf16(tm + hf * DQ2_HALF, qd + 2 * k, kd + 2 * k, idesc, k > 0);
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_bf16(tm + hf * DQ2_HALF + DQ2_DP_OFF, od + 2 * k, vd + 2 * k, idesc, k > 0);
        }
        if (leader) {
          umma_commit(&s_full[hf]);
          if (hf == 1 && sj + 1 == nb) umma_commit(&q_empty[sh & 1]);     // Q / dO of the head have no later reader
        }
        __syncwarp();
        if (hf == 1) {
          if (++sj == nb) { sj = 0; ++sh; }
          if (++sst == KVST) { sst = 0; sph ^= 1; }
        }
      };
      issue_s(0);
      issue_s(1);
      int h = 0, j = 0, st = 0;
      for (int g = 0; g < total; ++g) {
        const int nvalid = nvalid_of(j);
        const uint64_t kd = dkm0 + (uint64_t)(st * (BW_BLK_BYTES >> 4));
#pragma unroll
        for (int hf = 0; hf < 2; ++hf) {
          mbar_wait(&ds_full[hf], g & 1);
          if (j == 0 && hf == 0 && h > 0) mbar_wait(acc_empty, (h - 1) & 1);
          tc_fence_after();
          const int nvh = min(32, nvalid - 32 * hf);
          if (leader) {
            const int ksteps = nvh > 0 ? (nvh + 15) >> 4 : 0;
            for (int k = 0; k < ksteps; ++k)
              umma_bf16_ts(tm + DQ2_ACC, tm + hf * DQ2_HALF + k * 8, kd + (uint64_t)((2 * hf + k) * (2048 >> 4)), idesc_acc,
                           (j > 0) || (hf > 0) || (k > 0));
            if (hf == 1) {
              umma_commit(&kv_empty[st]);
              if (j + 1 == nb) umma_commit(acc_full);
            }
          }
          __syncwarp();
          if (g + 1 < total) issue_s(hf);
        }
        if (++j == nb) { j = 0; ++h; }
        if (++st == KVST) st = 0;
      }
    }
  } else {
    const int quarter = warp & 3;
    const int row_in_tile = quarter * 32 + lane;
    const uint32_t lane_addr = tmem + ((uint32_t)(quarter * 32) << 16);
    const int i = r0 + row_in_tile;
    const bool row_ok = i < r1;
    const bool warp_live = r0 + quarter * 32 < r1;
    uint32_t rs[32], rd[32];
    int g = 0;
    const int64_t stat0 = ((int64_t)b * p.H + h0) * p.N + (row_ok ? i : r0);
    float lse_raw = p.lse[stat0], dl_raw = p.delta[stat0];
    for (int h = 0; h < nh; ++h) {
      const float lse2 = row_ok ? lse_raw * 1.4426950408889634f : INFINITY;   // invalid rows -> P = 0
      const float dl = row_ok ? dl_raw : 0.f;
      if (h + 1 < nh) {
        lse_raw = p.lse[stat0 + (int64_t)(h + 1) * p.N];
        dl_raw = p.delta[stat0 + (int64_t)(h + 1) * p.N];
      }
      for (int j = 0; j < kb.nb; ++j, ++g) {
        int tok, nvalid; int64_t row;
        kb.get(j, tok, row, nvalid);
#pragma unroll
        for (int hf = 0; hf < 2; ++hf) {
          const int nvh = min(32, nvalid - 32 * hf);
          mbar_wait(&s_full[hf], g & 1);
          if (warp_live && nvh > 0) {
            tc_fence_after();
            uint32_t sd[64];                             // S (0..31) | dP (32..63) of the half, one TMEM read
            tmem_ld_32x64(lane_addr + hf * DQ2_HALF, sd);
            tmem_wait_ld();
            uint32_t pk[16];
#pragma unroll
            for (int t = 0; t < 16; ++t) {
              const float x0 = fmaf(__uint_as_float(sd[2 * t]), p.scale_log2, -lse2);
              const float x1 = fmaf(__uint_as_float(sd[2 * t + 1]), p.scale_log2, -lse2);
              float p0 = ((2 * t) & 3) < POLY ? exp2_poly(x0) : ex2(x0);
              float p1 = ((2 * t + 1) & 3) < POLY ? exp2_poly(x1) : ex2(x1);
              if (nvh < 32) {
                if (2 * t >= nvh) p0 = 0.f;
                if (2 * t + 1 >= nvh) p1 = 0.f;
              }
              pk[t] = pack_bf16(p0 * (__uint_as_float(sd[32 + 2 * t]) - dl) * p.scale, p1 * (__uint_as_float(sd[32 + 2 * t + 1]) - dl) * p.scale);
            }
            tmem_st_32x16(lane_addr + hf * DQ2_HALF, pk);
            tmem_wait_st();
          }
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&ds_full[hf]);
        }
      }
      mbar_wait(acc_full, h & 1);
      tc_fence_after();
      if (warp_live) {
        __nv_bfloat16* orow = p.dq + (q_row0 + row_in_tile) * p.lddq + (h0 + h) * 64;
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          tmem_ld_32x32(lane_addr + DQ2_ACC + c * 32, rs);
          tmem_wait_ld();
          if (row_ok) {
#pragma unroll
            for (int q = 0; q < 4; ++q)
              *reinterpret_cast<uint4*>(orow + c * 32 + q * 8) = make_uint4(
                  pack_bf16(__uint_as_float(rs[8 * q]), __uint_as_float(rs[8 * q + 1])),
                  pack_bf16(__uint_as_float(rs[8 * q + 2]), __uint_as_float(rs[8 * q + 3])),
                  pack_bf16(__uint_as_float(rs[8 * q + 4]), __uint_as_float(rs[8 * q + 5])),
                  pack_bf16(__uint_as_float(rs[8 * q + 6]), __uint_as_float(rs[8 * q + 7])));
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(acc_empty);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem, 256);
  }
}

// ---------------------------------------- dK, dV ----------------------------------------
// Round 2: a half's S^T and dP^T are ADJACENT (half c: S^T at 64c, dP^T at 64c + 32), so the elementwise warps fetch both with
// one 64-column tcgen05.ld and return P^T (64c .. +15) and dS^T (64c + 16 .. +31) with one 32-column tcgen05.st: the TMEM
// instructions carry a large fixed cost (splitting them further was measured 1.75x SLOWER, profiles/r02_attn_*chunked16*).
constexpr uint32_t KV_HALF = 64, KV_DPT_OFF = 32, KV_DS_OFF = 16, KV_DV = 128, KV_DK = 192;
constexpr int DKV_SMEM = 4 * TC_TILE_BYTES + 4 * BW_BLK_BYTES + 4 * 2 * 2 * BW_BLK * 4 + 1024 + 256;   // + per-warp lse / delta stages

__global__ void __launch_bounds__(TC_THREADS, 2)
attn_bwd_dkv_tc_kernel(const __grid_constant__ CUtensorMap tmap_k, const __grid_constant__ CUtensorMap tmap_v,
                       const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_do,
                       const AttnBwdTcParams p) {
  pdl_wait();   // launched through launch_pdl (common.cuh): nothing another kernel owns is touched before this
  extern __shared__ uint8_t smem_raw[];
  // key tile (cut per segment) and the query ranges that attend to it
  int c0 = 0, c1 = 0, qa0 = 0, qe0 = 0, qa1 = 0, qe1 = 0, b = 0;
  {
    int tile;
    lpt_key_tile(p.seg, p.nseg, p.tiles, p.B, tile, b);
    bool found = false;
    if (p.seg == nullptr) {
      c0 = tile * TC_BM; c1 = min(c0 + TC_BM, p.N); qa0 = 0; qe0 = p.N; found = c0 < p.N;
    } else {
      for (int s = 0; s < p.nseg; ++s) {
        const int a = p.seg[s], e = p.seg[s + 1];
        const int nt = (e - a + TC_BM - 1) / TC_BM;
        if (tile < nt) {
          c0 = a + tile * TC_BM; c1 = min(c0 + TC_BM, e);
          qa0 = a; qe0 = e;                                       // its own segment ...
          if (s != p.nseg - 1) { qa1 = p.seg[p.nseg - 1]; qe1 = p.seg[p.nseg]; }   // ... and the fusion rows
          found = true;
          break;
        }
        tile -= nt;
      }
    }
    if (!found) return;
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sK = smem;                                   // 2 buffers of 128x64 (per head)
  uint8_t* sV = smem + 2 * TC_TILE_BYTES;
  uint8_t* sQ = smem + 4 * TC_TILE_BYTES;               // 2 stages of 64x64
  uint8_t* sdO = sQ + 2 * BW_BLK_BYTES;
  float* s_lse = reinterpret_cast<float*>(sdO + 2 * BW_BLK_BYTES);   // [4 warps][2 stages][64]
  float* s_dl = s_lse + 4 * 2 * BW_BLK;                              // [4 warps][2 stages][64]
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_dl + 4 * 2 * BW_BLK);
  uint64_t* kvt_full = bars;        // [2] K and V tile of a head
  uint64_t* kvt_empty = bars + 2;   // [2]
  uint64_t* qb_full = bars + 4;     // [2] Q / dO block
  uint64_t* qb_empty = bars + 6;    // [2]
  uint64_t* s_full = bars + 8;      // [2] S^T / dP^T of one 32-query half of the block
  uint64_t* ds_full = bars + 10;    // [2] P^T / dS^T of that half written (S^T / dP^T consumed)
  uint64_t* acc_full = bars + 12;
  uint64_t* acc_empty = bars + 13;
  uint64_t* st_full = bars + 14;    // [2] lse / delta of a query block staged in shared memory
  uint64_t* st_empty = bars + 16;   // [2] all four elementwise warps are done with that stage
  uint64_t* acc_done = bars + 18;   // [2] the dV / dK MMAs that read a half's P^T / dS^T have completed: its columns are free
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 20);

  const int64_t k_row0 = c0 < p.n_head ? (int64_t)b * p.n_head + c0 : p.head_rows + (int64_t)b * p.n_tail + (c0 - p.n_head);
  RowBlocks qb;
  qb.init(qa0, qe0, qa1, qe1, p.n_head, p.n_tail, p.head_rows, b);

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_q); tma_prefetch_desc(&tmap_do); tma_prefetch_desc(&tmap_k); tma_prefetch_desc(&tmap_v);
  }
  if (warp == 1) {
    if (lane == 0) {
      for (int s = 0; s < 2; ++s) {
        mbar_init(&kvt_full[s], 1); mbar_init(&kvt_empty[s], 1);
        mbar_init(&qb_full[s], 1); mbar_init(&qb_empty[s], 1);
      }
      for (int s = 0; s < 2; ++s) { mbar_init(&s_full[s], 1); mbar_init(&ds_full[s], 4); }
      for (int s = 0; s < 2; ++s) { mbar_init(&st_full[s], 1); mbar_init(&st_empty[s], 4); mbar_init(&acc_done[s], 1); }
      mbar_init(acc_full, 1); mbar_init(acc_empty, 4);
      mbar_fence_init();
    }
    __syncwarp();
    tmem_alloc(tmem_slot, 256);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  // Two issuing warps (round 2).  One warp issuing everything spent ~2400 of the ~3600 clk a block takes on the 24 MMAs, the
  // commits and the waits of a block (~90 clk each, profiles/r02_attn_dkv_phase_clocks.txt), in series with the elementwise
  // warps it feeds.  Now warp 0 loads (TMA) and issues S^T / dP^T, warp 1 issues the dV / dK accumulations and stages the
  // lse / delta statistics; each half's chain  S -> elementwise -> accumulate -> S(next)  crosses the two warps through
  // completion barriers (s_full, ds_full, acc_done: tcgen05.commit only orders the issuing thread's own MMAs, so "its columns
  // are free" is the COMPLETION of the accumulate MMAs, not their issue).
  const int total = qb.nb * p.H;
  if (warp == 0) {
    // ------------------------------ TMA producer + S^T / dP^T issuer ------------------------------
    if (total > 0) {
      const bool leader = elect_one();
      const uint32_t tm = __shfl_sync(0xffffffffu, tmem, 0);
      const uint32_t idesc = umma_idesc_bf16(TC_BM, 32, false, false);    // S^T[128 keys x 32 q] = K . Q^T (one half); N = 32
      // whatever the number of valid queries: columns past them hold finite values and are zeroed through lse = +inf
      const uint64_t dk0 = umma_smem_desc(__shfl_sync(0xffffffffu, smem_u32(sK), 0), 16, 1024);
      const uint64_t dv0 = umma_smem_desc(__shfl_sync(0xffffffffu, smem_u32(sV), 0), 16, 1024);
      const uint64_t dq0 = umma_smem_desc(__shfl_sync(0xffffffffu, smem_u32(sQ), 0), 16, 1024);
      const uint64_t ddo0 = umma_smem_desc(__shfl_sync(0xffffffffu, smem_u32(sdO), 0), 16, 1024);
      const int nb = __shfl_sync(0xffffffffu, qb.nb, 0);
      int cnt0 = 0, cnt1 = 0;       // S^T halves issued so far: the n-th issue of a half waits for acc_done completion n - 1
      int h = 0, j = 0;             // block being loaded / issued
      for (int g = 0; g < total; ++g) {
        const int st = g & 1, ks = h & 1;
        int tok, nvalid; int64_t row;
        qb.get(j, tok, row, nvalid);
        if (j == 0) {               // K / V tile of the head
          mbar_wait(&kvt_empty[ks], ((h >> 1) & 1) ^ 1);
          if (leader) {
            mbar_expect_tx(&kvt_full[ks], 2 * TC_TILE_BYTES);
            tma_load_2d(sK + ks * TC_TILE_BYTES, &tmap_k, &kvt_full[ks], h * 64, (int)k_row0);
            tma_load_2d(sV + ks * TC_TILE_BYTES, &tmap_v, &kvt_full[ks], h * 64, (int)k_row0);
          }
        }
        mbar_wait(&qb_empty[st], ((g >> 1) & 1) ^ 1);
        if (leader) {
          mbar_expect_tx(&qb_full[st], 2 * BW_BLK_BYTES);
          tma_load_2d(sQ + st * BW_BLK_BYTES, &tmap_q, &qb_full[st], h * 64, (int)row);
          tma_load_2d(sdO + st * BW_BLK_BYTES, &tmap_do, &qb_full[st], h * 64, (int)row);
        }
        __syncwarp();
#pragma unroll
        for (int hf = 0; hf < 2; ++hf) {
          if (nvalid <= 32 * hf) continue;
          const int n_prev = hf == 0 ? cnt0 : cnt1;
          if (n_prev > 0) mbar_wait(&acc_done[hf], (n_prev - 1) & 1);     // the half's previous P^T / dS^T have been consumed
          if (hf == 0) {
            if (j == 0) mbar_wait(&kvt_full[ks], (h >> 1) & 1);
            mbar_wait(&qb_full[st], (g >> 1) & 1);
          }
          tc_fence_after();
          if (leader) {
            const uint64_t kd = dk0 + (uint64_t)(ks * (TC_TILE_BYTES >> 4)), vd = dv0 + (uint64_t)(ks * (TC_TILE_BYTES >> 4));
            const uint64_t qd = dq0 + (uint64_t)(st * (BW_BLK_BYTES >> 4) + hf * (4096 >> 4));
            const uint64_t od = ddo0 + (uint64_t)(st * (BW_BLK_BYTES >> 4) + hf * (4096 >> 4));
#pragma unroll
            for (int k2 = 0; k2 < 4; ++k2) umma_bf16(tm + hf * KV_HALF, kd + 2 * k2, qd + 2 * k2, idesc, k2 > 0);
#pragma unroll
            for (int k2 = 0; k2 < 4; ++k2) umma_bf16(tm + hf * KV_HALF + KV_DPT_OFF, vd + 2 * k2, od + 2 * k2, idesc, k2 > 0);
            umma_commit(&s_full[hf]);
          }
          __syncwarp();
          if (hf == 0) ++cnt0; else ++cnt1;
        }
        if (++j == nb) { j = 0; ++h; }
      }
    }
  } else if (warp == 1) {
    // ------------------------------ dV / dK issuer + statistics stager ------------------------------
    if (total > 0) {
      const bool leader = elect_one();
      const uint32_t tm = __shfl_sync(0xffffffffu, tmem, 0);
      const uint32_t idesc_acc = umma_idesc_bf16(TC_BM, 64, false, true);     // dV/dK[128 x 64dh] += A(TMEM)[128 x q] . B (MN-major)
      const uint64_t dqm0 = umma_smem_desc(__shfl_sync(0xffffffffu, smem_u32(sQ), 0), 8192, 1024);    // Q / dO read MN-major
      const uint64_t ddom0 = umma_smem_desc(__shfl_sync(0xffffffffu, smem_u32(sdO), 0), 8192, 1024);
      const int nb = __shfl_sync(0xffffffffu, qb.nb, 0);
      // lse * log2e and delta of a block's 64 query rows (lane l: rows l and l + 32), staged ONE block ahead for the four
      // elementwise warps: the global loads of block g + 1 are issued during block g and consumed at its end
      float l0 = 0.f, l1 = 0.f, d0 = 0.f, d1 = 0.f;
      bool ok0 = false, ok1 = false;
      auto load_stats = [&](int hh, int jj) {
        int tok, nvalid; int64_t row;
        qb.get(jj, tok, row, nvalid);
        const int64_t sb_ = ((int64_t)b * p.H + hh) * p.N + tok;
        ok0 = lane < nvalid; ok1 = lane + 32 < nvalid;
        l0 = p.lse[sb_ + (ok0 ? lane : 0)]; d0 = p.delta[sb_ + (ok0 ? lane : 0)];
        l1 = p.lse[sb_ + (ok1 ? lane + 32 : 0)]; d1 = p.delta[sb_ + (ok1 ? lane + 32 : 0)];
      };
      auto store_stats = [&](int gg) {
        const int stg = gg & 1;
        mbar_wait(&st_empty[stg], ((gg >> 1) & 1) ^ 1);
        sts_f1(smem_u32(s_lse + stg * BW_BLK + lane), ok0 ? l0 * 1.4426950408889634f : INFINITY);        // +inf beyond nvalid -> P = 0
        sts_f1(smem_u32(s_lse + stg * BW_BLK + lane + 32), ok1 ? l1 * 1.4426950408889634f : INFINITY);
        sts_f1(smem_u32(s_dl + stg * BW_BLK + lane), ok0 ? d0 : 0.f);
        sts_f1(smem_u32(s_dl + stg * BW_BLK + lane + 32), ok1 ? d1 : 0.f);
        __syncwarp();
        if (lane == 0) mbar_arrive(&st_full[stg]);
      };
      load_stats(0, 0);
      store_stats(0);
      if (total > 1) load_stats(nb > 1 ? 0 : 1, nb > 1 ? 1 : 0);     // block 1, stored at the top of iteration 0
      int used0 = 0, used1 = 0;   // completed waits on ds_full[0] / ds_full[1]
#ifdef MMF_ATTN_CLOCKS
      const bool dbg_on = (b == 3) && (c0 == 0) && lane == 0;
      unsigned t_last = clock();
      unsigned acc_clk[16] = {0};
#endif
      int h = 0, j = 0;
      for (int g = 0; g < total; ++g) {
        const int st = g & 1;
        int tok, nvalid; int64_t row;
        qb.get(j, tok, row, nvalid);
        // statistics run TWO blocks ahead of the accumulation this warp issues: block g + 1 is published now (the
        // elementwise warps reach it while this warp still waits for block g's ds_full), block g + 2 is requested
        if (g + 1 < total) store_stats(g + 1);
        if (g + 2 < total) {
          int h2 = h, j2 = j + 2;
          while (j2 >= nb) { j2 -= nb; ++h2; }
          load_stats(h2, j2);
        }
        const uint64_t qd = dqm0 + (uint64_t)(st * (BW_BLK_BYTES >> 4)), od = ddom0 + (uint64_t)(st * (BW_BLK_BYTES >> 4));
#pragma unroll
        for (int hf = 0; hf < 2; ++hf) {
          const int nvh = min(32, nvalid - 32 * hf);
          if (nvh <= 0) continue;
          CLK2(8, 0);
          if (hf == 0) { mbar_wait(&ds_full[0], used0 & 1); ++used0; } else { mbar_wait(&ds_full[1], used1 & 1); ++used1; }
          CLK2(9, 0);
          if (j == 0 && hf == 0 && h > 0) mbar_wait(acc_empty, (h - 1) & 1);   // previous head's dV / dK have been read out
          tc_fence_after();
          CLK2(10, 0);
          if (leader) {
            const int ksteps = (nvh + 15) >> 4;
            for (int k2 = 0; k2 < ksteps; ++k2)   // dV += P^T . dO
              umma_bf16_ts(tm + KV_DV, tm + hf * KV_HALF + k2 * 8, od + (uint64_t)((2 * hf + k2) * (2048 >> 4)), idesc_acc,
                           (j > 0) || (hf > 0) || (k2 > 0));
            for (int k2 = 0; k2 < ksteps; ++k2)   // dK += dS^T . Q
              umma_bf16_ts(tm + KV_DK, tm + hf * KV_HALF + KV_DS_OFF + k2 * 8, qd + (uint64_t)((2 * hf + k2) * (2048 >> 4)), idesc_acc,
                           (j > 0) || (hf > 0) || (k2 > 0));
            umma_commit(&acc_done[hf]);           // the half's columns are free once these have completed
          }
          __syncwarp();
          CLK2(11, 0);
        }
        if (leader) {
          umma_commit(&qb_empty[st]);             // every MMA reading this Q / dO stage has completed by then
          if (j + 1 == nb) {
            umma_commit(acc_full);
            umma_commit(&kvt_empty[h & 1]);
          }
        }
        __syncwarp();
        CLK2(12, 0);
        if (++j == nb) { j = 0; ++h; }
      }
#ifdef MMF_ATTN_CLOCKS
      if (dbg_on) for (int i = 8; i < 13; ++i) g_attn_clk[i] = acc_clk[i];
#endif
    }
  } else {
    const int quarter = warp & 3;
    const int row_in_tile = quarter * 32 + lane;      // key row of this thread
    const uint32_t lane_addr = tmem + ((uint32_t)(quarter * 32) << 16);
    const bool key_ok = c0 + row_in_tile < c1;
    uint32_t rs[32], rd[32];
    int g = 0;
    int used0 = 0, used1 = 0;   // completed waits on s_full[0] / s_full[1]
#ifdef MMF_ATTN_CLOCKS
    const bool dbg_on = (b == 3) && (c0 == 0) && warp == 2 && lane == 0;   // first modality key tile of sample 3
    if (dbg_on) g_attn_clk[15] = (unsigned long long)qb.nb * p.H;
    unsigned t_last = clock();
    unsigned acc_clk[16] = {0};
#endif
    // lse * log2e and delta of the block's 64 query rows come staged in shared memory from the producer warp (st_full / st_empty)
    for (int h = 0; h < p.H; ++h) {
      for (int j = 0; j < qb.nb; ++j, ++g) {
        const int st = g & 1;
        int tok, nvalid; int64_t row;
        qb.get(j, tok, row, nvalid);
        CLK2(4, 0);
        mbar_wait(&st_full[st], (g >> 1) & 1);
        CLK2(5, 0);
        const uint32_t ls = smem_u32(s_lse + st * BW_BLK), dl = smem_u32(s_dl + st * BW_BLK);   // explicit ld.shared below: as float* these were generic LD.E.128
#pragma unroll
        for (int c = 0; c < 2; ++c) {   // the block's two 32-query halves (see the MMA warp)
          if (nvalid <= 32 * c) continue;
          if (c == 0) { mbar_wait(&s_full[0], used0 & 1); ++used0; } else { mbar_wait(&s_full[1], used1 & 1); ++used1; }
          tc_fence_after();
          CLK2(0, 0);
          uint32_t sd[64];                               // S^T (0..31) | dP^T (32..63) of the half, one TMEM read
          tmem_ld_32x64(lane_addr + c * KV_HALF, sd);
          tmem_wait_ld();
          uint32_t pd[32];                               // P^T (0..15) | dS^T (16..31), one TMEM write
#pragma unroll
          for (int t = 0; t < 8; ++t) {   // four query columns per step: one 16-byte broadcast read of lse and of delta
            const float4 l4 = lds_f4(ls + (c * 8 + t) * 16), d4 = lds_f4(dl + (c * 8 + t) * 16);
            const float p0 = ex2(fmaf(__uint_as_float(sd[4 * t]), p.scale_log2, -l4.x));        // lse = +inf beyond nvalid -> 0
            const float p1 = ex2(fmaf(__uint_as_float(sd[4 * t + 1]), p.scale_log2, -l4.y));
            const float p2 = ex2(fmaf(__uint_as_float(sd[4 * t + 2]), p.scale_log2, -l4.z));
            const float p3 = ex2(fmaf(__uint_as_float(sd[4 * t + 3]), p.scale_log2, -l4.w));
            pd[2 * t] = pack_bf16(p0, p1);
            pd[2 * t + 1] = pack_bf16(p2, p3);
            pd[16 + 2 * t] = pack_bf16(p0 * (__uint_as_float(sd[32 + 4 * t]) - d4.x), p1 * (__uint_as_float(sd[32 + 4 * t + 1]) - d4.y));
            pd[16 + 2 * t + 1] = pack_bf16(p2 * (__uint_as_float(sd[32 + 4 * t + 2]) - d4.z), p3 * (__uint_as_float(sd[32 + 4 * t + 3]) - d4.w));
          }
          tmem_st_32x32(lane_addr + c * KV_HALF, pd);    // in place over the half's consumed S^T columns
          tmem_wait_st();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&ds_full[c]);
          CLK2(1, 0);
        }
        CLK2(2, 0);
        __syncwarp();
        if (lane == 0) mbar_arrive(&st_empty[st]);      // this warp has read the stage's statistics
        CLK2(3, 0);
      }
      if (qb.nb == 0) continue;
      mbar_wait(acc_full, h & 1);
      tc_fence_after();
      const int64_t orow = k_row0 + row_in_tile;
      __nv_bfloat16* vrow = p.dv + orow * p.lddv + h * 64;
      __nv_bfloat16* krow = p.dk + orow * p.lddk + h * 64;
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        tmem_ld_32x32(lane_addr + KV_DV + c * 32, rs);
        tmem_ld_32x32(lane_addr + KV_DK + c * 32, rd);
        tmem_wait_ld();
        if (key_ok) {
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            *reinterpret_cast<uint4*>(vrow + c * 32 + q * 8) = make_uint4(
                pack_bf16(__uint_as_float(rs[8 * q]), __uint_as_float(rs[8 * q + 1])),
                pack_bf16(__uint_as_float(rs[8 * q + 2]), __uint_as_float(rs[8 * q + 3])),
                pack_bf16(__uint_as_float(rs[8 * q + 4]), __uint_as_float(rs[8 * q + 5])),
                pack_bf16(__uint_as_float(rs[8 * q + 6]), __uint_as_float(rs[8 * q + 7])));
            *reinterpret_cast<uint4*>(krow + c * 32 + q * 8) = make_uint4(
                pack_bf16(__uint_as_float(rd[8 * q]) * p.scale, __uint_as_float(rd[8 * q + 1]) * p.scale),
                pack_bf16(__uint_as_float(rd[8 * q + 2]) * p.scale, __uint_as_float(rd[8 * q + 3]) * p.scale),
                pack_bf16(__uint_as_float(rd[8 * q + 4]) * p.scale, __uint_as_float(rd[8 * q + 5]) * p.scale),
                pack_bf16(__uint_as_float(rd[8 * q + 6]) * p.scale, __uint_as_float(rd[8 * q + 7]) * p.scale));
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(acc_empty);
      CLK2(6, 0);
    }
#ifdef MMF_ATTN_CLOCKS
    if (dbg_on) for (int i = 0; i < 7; ++i) g_attn_clk[i] = acc_clk[i];
#endif
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem, 256);
  }
}

static int tc_make_tmap_box(CUtensorMap* map, const void* base, int64_t rows, int64_t cols, int64_t ld, int box_rows) {
  EncodeTiledFn enc = tc_encode_fn();
  if (!enc) return 1000;
  cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t gstride[1] = {(cuuint64_t)ld * 2};
  cuuint32_t box[2] = {64, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), gdim, gstride, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : 2000 + (int)r;
}

// dq / dk / dv of the tcgen05 backward (delta must already be computed).  -1000: not eligible.
int attn_bwd_tc_launch(const MmfAttnArgs* a, cudaStream_t stream) {
  if (a->dh != 64 || a->Nq != a->Nk || a->n_head_q != a->n_head_k) return -1000;
  if ((a->ldq & 7) || (a->ldk & 7) || (a->ldv & 7) || (a->lddo & 7) || (a->lddq & 7) || (a->lddk & 7) || (a->lddv & 7)) return -1000;
  const int64_t rows = (int64_t)a->B * a->Nq;
  const int64_t cols = (int64_t)a->H * 64;
  CUtensorMap q128, do128, k64, v64, k128, v128, q64, do64;
  int rc;
  if ((rc = tc_make_tmap_box(&q128, a->q, rows, cols, a->ldq, 128))) return rc;
  if ((rc = tc_make_tmap_box(&do128, a->d_o, rows, cols, a->lddo, 128))) return rc;
  if ((rc = tc_make_tmap_box(&k64, a->k, rows, cols, a->ldk, 64))) return rc;
  if ((rc = tc_make_tmap_box(&v64, a->v, rows, cols, a->ldv, 64))) return rc;
  if ((rc = tc_make_tmap_box(&k128, a->k, rows, cols, a->ldk, 128))) return rc;
  if ((rc = tc_make_tmap_box(&v128, a->v, rows, cols, a->ldv, 128))) return rc;
  if ((rc = tc_make_tmap_box(&q64, a->q, rows, cols, a->ldq, 64))) return rc;
  if ((rc = tc_make_tmap_box(&do64, a->d_o, rows, cols, a->lddo, 64))) return rc;
  AttnBwdTcParams p;
  p.lse = a->lse; p.delta = a->delta;
  p.dq = reinterpret_cast<__nv_bfloat16*>(a->dq); p.dk = reinterpret_cast<__nv_bfloat16*>(a->dk); p.dv = reinterpret_cast<__nv_bfloat16*>(a->dv);
  p.lddq = a->lddq; p.lddk = a->lddk; p.lddv = a->lddv;
  p.B = a->B; p.H = a->H; p.N = a->Nq; p.n_head = a->n_head_q; p.n_tail = a->n_tail_q;
  p.head_rows = (int64_t)a->B * a->n_head_q;
  p.scale = a->scale; p.scale_log2 = a->scale * 1.4426950408889634f;
  p.seg = a->seg; p.nseg = a->nseg;
  {
    const char* pe = getenv("MMF_ATTN_PF");
    p.pf = pe ? atoi(pe) : TC_PF_DEFAULT;
    if (p.pf < 0 || p.pf > 64) p.pf = TC_PF_DEFAULT;
  }
  const int tiles = (a->Nq + TC_BM - 1) / TC_BM + (a->seg ? a->nseg : 0);
  p.tiles = tiles;
  const int heavy_max = a->seg ? (a->Nq + TC_BM - 1) / TC_BM : 0;
  const int grid_q = (tiles + (TC_HSPLIT - 1) * heavy_max) * a->B;
  // MMF_ATTN_DQ: 0 = round-1 dQ kernel, 1 = half-block pipeline (attn_bwd_dq_tc2_kernel), 2 / 3 = the same with 1 / 2 of
  // every 4 exponentials on the FMA pipe.  Read per call (A/B runs switch it between launches).
  const char* env = getenv("MMF_ATTN_DQ");
  int vq = env ? atoi(env) : TC_DQ_DEFAULT;
  if (vq < 0 || vq > 3) vq = TC_DQ_DEFAULT;
  static std::atomic<unsigned> attr_done_v[5];   // [0..3] dQ variants, [4] dK/dV: bit d set = attributes applied on device d
  int dev = 0;
  cudaGetDevice(&dev);
  const unsigned bit = 1u << (dev & 31);
  auto prep = [&](std::atomic<unsigned>& done, const void* fn, int smem) -> int {
    if (done.load(std::memory_order_acquire) & bit) return 0;
    cudaError_t e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return (int)e;
    cudaFuncSetAttribute(fn, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    done.fetch_or(bit, std::memory_order_release);
    return 0;
  };
#define MMF_DQ2(IDX, KS, PL)                                                                                          \
  do {                                                                                                                 \
    constexpr int smem = 4 * TC_TILE_BYTES + 2 * KS * BW_BLK_BYTES + 256;                                              \
    if ((rc = prep(attr_done_v[IDX], reinterpret_cast<const void*>(attn_bwd_dq_tc2_kernel<KS, PL>), smem))) return rc;  \
    launch_pdl(attn_bwd_dq_tc2_kernel<KS, PL>, dim3(grid_q), dim3(TC_THREADS), smem, stream, q128, do128, k64, v64, p);                    \
  } while (0)
  switch (vq) {
    case 0:
      if ((rc = prep(attr_done_v[0], reinterpret_cast<const void*>(attn_bwd_dq_tc_kernel), DQ_SMEM))) return rc;
      attn_bwd_dq_tc_kernel<<<grid_q, TC_THREADS, DQ_SMEM, stream>>>(q128, do128, k64, v64, p);
      break;
    case 2: MMF_DQ2(2, 3, 0); break;      // three K/V stages (2 CTAs per SM still fit)
    case 3: MMF_DQ2(3, 3, 1); break;      // + 1 of 4 exponentials on the FMA pipe
    default: MMF_DQ2(1, 2, 0); break;
  }
#undef MMF_DQ2
  if ((rc = prep(attr_done_v[4], reinterpret_cast<const void*>(attn_bwd_dkv_tc_kernel), DKV_SMEM))) return rc;
  launch_pdl(attn_bwd_dkv_tc_kernel, dim3(tiles * a->B), dim3(TC_THREADS), DKV_SMEM, stream, k128, v128, q64, do64, p);
  g_launch_count.fetch_add(2, std::memory_order_relaxed);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? 0 : (int)e;
}

}  // namespace mmf

#ifdef MMF_ATTN_CLOCKS
extern "C" int mmf_debug_attn_clocks(unsigned long long* out16, int reset) {
  if (reset) {
    unsigned long long z[16] = {0};
    return (int)cudaMemcpyToSymbol(mmf::g_attn_clk, z, sizeof(z));
  }
  return (int)cudaMemcpyFromSymbol(out16, mmf::g_attn_clk, 16 * sizeof(unsigned long long));
}
#endif
