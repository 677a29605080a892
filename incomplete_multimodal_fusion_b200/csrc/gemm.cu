// tcgen05 GEMM for sm_100a:  out[M,N] = epilogue(alpha * A[M,K] . B[N,K]^T), bf16 in, fp32 accumulate.
//
// Persistent, warp-specialised, one CTA per SM:
//   warp 0      : TMA producer   (cp.async.bulk.tensor -> 128B-swizzled smem ring, mbarrier tx-count)
//   warp 1      : MMA issuer     (one elected lane issues tcgen05.mma 128 x BLOCK_N x 16, accumulators in TMEM)
//   warps 2..5  : epilogue       (tcgen05.ld TMEM -> registers -> fused epilogue -> global)
// Two TMEM accumulator stages let the epilogue of tile i overlap the main loop of tile i+1.
// Operands may be K-major ([rows, K]) or MN-major (stored [K, rows]); the latter serves dgrad / wgrad
// without materialising transposes.  split_k > 1 accumulates with fp32 red.global.add.
#include "common.cuh"
#include "mmf_b200.h"

#include <atomic>
#include <cstdlib>
#include <mutex>

namespace mmf {

std::atomic<int64_t> g_launch_count{0};
std::atomic<int> g_reserved_sms{0};   // SMs the persistent GEMM grids leave free (for a concurrent NCCL all-reduce)

constexpr int BLOCK_M = 128;
constexpr int BLOCK_K = 64;  // 64 bf16 = 128 B = one swizzle atom
constexpr int UMMA_K = 16;
constexpr int GEMM_THREADS = 192;
constexpr int SMEM_BUDGET = 200 * 1024;
constexpr int STG_FLOATS = 32 * 32;  // per-warp 32x32 fp32 epilogue staging block

template <int BLOCK_N>
struct GemmCfg {
  static constexpr int A_BYTES = BLOCK_M * BLOCK_K * 2;
  static constexpr int B_BYTES = BLOCK_N * BLOCK_K * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int STAGES = SMEM_BUDGET / STAGE_BYTES;  // 4 @256, 6 @128
  static constexpr int TMEM_COLS = 2 * BLOCK_N;             // two accumulator stages
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align*/ + 256 /*barriers*/ + 4 * STG_FLOATS * 4 /*epilogue staging*/;
};

struct GemmParams {
  void* out;
  void* out2;
  const float* bias;
  const float* residual;
  const float* residual2;
  int64_t res_split;
  const int32_t* res_row_map;
  int64_t M, N, K;
  int64_t ldo, ldo2, ldr;
  int32_t out_f32, act, split_k, res_period, out_period, out_batch_rows, accumulate, a_mn, b_mn;
  int32_t m_tiles, n_tiles, num_kb, kb_per_split;
  int64_t geglu_ipad;  // act==2: row offset of the gate half inside B
  float alpha;
  int32_t fast_ok;  // all pitches / pointers allow the vectorised epilogue
  int32_t one;      // always 1, but opaque to the compiler: pins basic-block boundaries in the epilogues (see GEGLU)
  int32_t epi_flags;  // debug builds (-DMMF_GEMM_CLOCKS) only: epilogue ablation bits, see GABL
};

__device__ __forceinline__ int64_t shfl_i64(int64_t v, int src) {
  const uint32_t lo = __shfl_sync(0xffffffffu, (uint32_t)(v & 0xffffffffu), src);
  const uint32_t hi = __shfl_sync(0xffffffffu, (uint32_t)((uint64_t)v >> 32), src);
  return (int64_t)(((uint64_t)hi << 32) | lo);
}

// Epilogue kinds (compile-time: the four epilogue warps are instruction-issue bound, so each kind gets a lean
// instruction stream instead of one generic, branchy one that overflows the instruction cache).
enum : int {
  EPI_BF16 = 0,     // out bf16 = alpha*acc (+bias)
  EPI_GELU = 1,     // out bf16 = gelu(alpha*acc + bias), out2 (optional) = pre-activation bf16
  EPI_F32 = 2,      // out f32 = alpha*acc (+bias) (+residual), out2 (optional) = bf16 copy
  EPI_ATOMIC = 3,   // out f32 += alpha*acc   (split-K / accumulate)
  EPI_GEGLU = 4,    // out bf16 = gelu(gate)*value, out2 (optional) = [value | gate] bf16
  EPI_BF16_ACC = 5, // out bf16 += alpha*acc
  EPI_GEGLU_BWD = 6 // acc = dg (gradient of the GEGLU output); out2 = saved [value | gate] (INPUT), out = [dvalue | dgate] bf16
};

#ifdef MMF_GEMM_CLOCKS
// timing experiments only (tools/gemm_clocks.py, debug build): per-phase clock totals of one cluster's warps
__device__ unsigned long long g_gemm_clk[32];
#define GCLK(i) do { if (dbg_on) { const unsigned t__ = clock(); acc_clk[i] += t__ - t_last; t_last = t__; } } while (0)
#define GCLK_DECL(cond) const bool dbg_on = (cond); unsigned acc_clk[10] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0}; unsigned t_last = clock(); (void)t_last
#define GCLK_FLUSH(base, n) do { if (dbg_on && (threadIdx.x & 31) == 0) for (int i__ = 0; i__ < (n); ++i__) g_gemm_clk[(base) + i__] = acc_clk[i__]; } while (0)
#define GABL(bit) ((p.epi_flags & (bit)) != 0)   // ablations of the GEGLU-backward epilogue (MMF_GEGLU_BWD_ABL): 2 math, 4 stores, 8 loads
#else
#define GABL(bit) false
#define GCLK(i) do { } while (0)
#define GCLK_DECL(cond) do { } while (0)
#define GCLK_FLUSH(base, n) do { } while (0)
#endif

// Per-warp 32x32 fp32 staging block, 128-byte rows, 16-byte groups XOR-swizzled with (row & 7): the row-per-lane
// 128-bit writes and the row-contiguous 128-bit reads both run at one 128-byte wavefront per 8 lanes.

__device__ __forceinline__ float4 gelu4(float4 x) { return make_float4(gelu_erf(x.x), gelu_erf(x.y), gelu_erf(x.z), gelu_erf(x.w)); }
__device__ __forceinline__ uint2 pack4(float4 x) { return make_uint2(pack_bf16(x.x, x.y), pack_bf16(x.z, x.w)); }

// The accumulator arrives "one row per lane" (tcgen05.ld 32x32b).  Global accesses in that shape touch 32 different
// rows per instruction, so each warp transposes its 32x32 block through shared memory; afterwards lane l owns the 4
// columns 4*(l%8).. of rows 4*i + l/8 (i = 0..7): 8 lanes x 16 B cover one 128-byte row segment.
__device__ __forceinline__ void stage_raw(float* stg, const uint32_t (&raw)[32], int lane) {
#pragma unroll
  for (int j = 0; j < 8; ++j)
    *reinterpret_cast<uint4*>(stg + lane * 32 + ((j ^ (lane & 7)) << 2)) = make_uint4(raw[4 * j], raw[4 * j + 1], raw[4 * j + 2], raw[4 * j + 3]);
  __syncwarp();
}
__device__ __forceinline__ float4 stage_read(const float* stg, int i, int lane) {
  const int row = i * 4 + (lane >> 3);
  return *reinterpret_cast<const float4*>(stg + row * 32 + (((lane & 7) ^ (row & 7)) << 2));
}

// residual values of one 32-column chunk in the post-transpose ownership (8 rows x 4 columns per lane)
__device__ __forceinline__ void load_res8(float4 (&resv)[8], const int64_t (&orow_i)[8], const float* const (&res_i)[8],
                                          int64_t col) {
#pragma unroll
  for (int i = 0; i < 8; ++i)
    if (orow_i[i] >= 0) resv[i] = __ldg(reinterpret_cast<const float4*>(res_i[i] + col));
}

// one full, aligned 32-column chunk; `resv` = the chunk's residual values (EPI_F32 with a residual only)
template <int EPI>
__device__ __forceinline__ void epi_chunk(const GemmParams& p, float* stg, const uint32_t (&raw)[32], int lane,
                                          const int64_t (&orow_i)[8], const float4 (&resv)[8], int64_t col0) {
  const int c = (lane & 7) * 4;
  const bool has_res = (EPI == EPI_F32) && p.residual != nullptr;
  float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f);
  if (EPI != EPI_ATOMIC && EPI != EPI_BF16_ACC && p.bias != nullptr) b4 = __ldg(reinterpret_cast<const float4*>(p.bias + col0 + c));
  stage_raw(stg, raw, lane);
  const float alpha = p.alpha;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int64_t orow = orow_i[i];
    if (orow < 0) continue;
    float4 x = stage_read(stg, i, lane);
    x.x = fmaf(x.x, alpha, b4.x); x.y = fmaf(x.y, alpha, b4.y); x.z = fmaf(x.z, alpha, b4.z); x.w = fmaf(x.w, alpha, b4.w);
    if (EPI == EPI_GELU) {
      if (p.out2) *reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(p.out2) + orow * p.ldo2 + col0 + c) = pack4(x);
      x = gelu4(x);
    }
    if (EPI == EPI_F32) {
      if (has_res) { x.x += resv[i].x; x.y += resv[i].y; x.z += resv[i].z; x.w += resv[i].w; }
      *reinterpret_cast<float4*>(reinterpret_cast<float*>(p.out) + orow * p.ldo + col0 + c) = x;
      if (p.out2) *reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(p.out2) + orow * p.ldo2 + col0 + c) = pack4(x);
    } else if (EPI == EPI_ATOMIC) {
      float* o = reinterpret_cast<float*>(p.out) + orow * p.ldo + col0 + c;
      // one 16-byte vector reduction instead of four scalar ones: split-K wgrads are paced by L2 atomic throughput
      asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(o), "f"(x.x), "f"(x.y), "f"(x.z), "f"(x.w) : "memory");
    } else {
      __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(p.out) + orow * p.ldo + col0 + c;
      if (EPI == EPI_BF16_ACC) {
        const uint2 old = *reinterpret_cast<const uint2*>(o);
        const float2 a = unpack_bf16(old.x), b = unpack_bf16(old.y);
        x.x += a.x; x.y += a.y; x.z += b.x; x.w += b.y;
      }
      *reinterpret_cast<uint2*>(o) = pack4(x);
    }
  }
  __syncwarp();
}

__device__ __forceinline__ void epi_chunk_geglu(const GemmParams& p, float* stg, const uint32_t (&raw)[32],
                                                const uint32_t (&rawg)[32], int lane, const int64_t (&orow_i)[8], int64_t col0) {
  const int c = (lane & 7) * 4;
  float4 val[8];
  stage_raw(stg, raw, lane);
#pragma unroll
  for (int i = 0; i < 8; ++i) val[i] = stage_read(stg, i, lane);
  __syncwarp();
  stage_raw(stg, rawg, lane);
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int64_t orow = orow_i[i];
    if (orow < 0) continue;
    const float4 g = stage_read(stg, i, lane);
    if (p.out2) {
      __nv_bfloat16* u = reinterpret_cast<__nv_bfloat16*>(p.out2) + orow * p.ldo2 + col0 + c;
      *reinterpret_cast<uint2*>(u) = pack4(val[i]);
      *reinterpret_cast<uint2*>(u + p.geglu_ipad) = pack4(g);
    }
    const float4 ge = gelu4(g);
    *reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(p.out) + orow * p.ldo + col0 + c) =
        pack4(make_float4(ge.x * val[i].x, ge.y * val[i].y, ge.z * val[i].z, ge.w * val[i].w));
  }
  __syncwarp();
}

// ragged / unaligned chunk: row-per-lane scalar code, deliberately compact (runs only on the last column tile of
// odd-sized problems); `g` is the gate accumulator in GEGLU mode
template <int EPI>
__device__ __noinline__ void epi_chunk_slow(const GemmParams& p, const uint32_t (&raw)[32], const uint32_t (&rawg)[32],
                                            int64_t orow, const float* res_row, int64_t col0, int nvalid) {
  if (orow < 0) return;
#pragma unroll 1
  for (int j = 0; j < nvalid; ++j) {
    float x = __uint_as_float(raw[j]) * p.alpha;
    const int64_t col = col0 + j;
    if (EPI == EPI_GEGLU) {
      const float g = __uint_as_float(rawg[j]);
      if (p.out2) {
        __nv_bfloat16* u = reinterpret_cast<__nv_bfloat16*>(p.out2) + orow * p.ldo2 + col;
        u[0] = __float2bfloat16(x);
        u[p.geglu_ipad] = __float2bfloat16(g);
      }
      reinterpret_cast<__nv_bfloat16*>(p.out)[orow * p.ldo + col] = __float2bfloat16(gelu_erf(g) * x);
      continue;
    }
    if (EPI != EPI_ATOMIC && EPI != EPI_BF16_ACC && p.bias != nullptr) x += __ldg(p.bias + col);
    if (EPI == EPI_GELU) {
      if (p.out2) reinterpret_cast<__nv_bfloat16*>(p.out2)[orow * p.ldo2 + col] = __float2bfloat16(x);
      x = gelu_erf(x);
    }
    if (EPI == EPI_F32) {
      if (res_row != nullptr) x += __ldg(res_row + col);
      reinterpret_cast<float*>(p.out)[orow * p.ldo + col] = x;
      if (p.out2) reinterpret_cast<__nv_bfloat16*>(p.out2)[orow * p.ldo2 + col] = __float2bfloat16(x);
    } else if (EPI == EPI_ATOMIC) {
      atomicAdd(reinterpret_cast<float*>(p.out) + orow * p.ldo + col, x);
    } else {
      __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(p.out) + orow * p.ldo + col;
      if (EPI == EPI_BF16_ACC) x += __bfloat162float(*o);
      *o = __float2bfloat16(x);
    }
  }
}

template <int BLOCK_N, int EPI>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_tcgen05_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                    const GemmParams p) {
  using Cfg = GemmCfg<BLOCK_N>;
  constexpr int STAGES = Cfg::STAGES;
  constexpr bool geglu = (EPI == EPI_GEGLU);
  extern __shared__ uint8_t smem_raw[];
  // 128B swizzle atoms need 1024-byte aligned tiles
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + STAGES * Cfg::A_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * Cfg::STAGE_BYTES);
  uint64_t* full_bar = bars;                     // [STAGES]
  uint64_t* empty_bar = bars + STAGES;           // [STAGES]
  uint64_t* tmem_full = bars + 2 * STAGES;       // [2]
  uint64_t* tmem_empty = bars + 2 * STAGES + 2;  // [2]
  uint32_t* tmem_base_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4);
  float* stage_all = reinterpret_cast<float*>(smem + STAGES * Cfg::STAGE_BYTES + 256);  // 16-byte aligned

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  // in GEGLU mode one 256-wide MMA tile yields 128 output columns (value | gate halves)
  constexpr int out_cols_per_tile = geglu ? BLOCK_N / 2 : BLOCK_N;
  const int tiles_mn = p.m_tiles * p.n_tiles;
  const int total_work = tiles_mn * p.split_k;
  const bool a_mn = p.a_mn != 0, b_mn = p.b_mn != 0;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_b);
  }
  if (warp == 1) {
    if (lane == 0) {
      for (int s = 0; s < STAGES; ++s) {
        mbar_init(&full_bar[s], 1);
        mbar_init(&empty_bar[s], 1);
      }
      for (int s = 0; s < 2; ++s) {
        mbar_init(&tmem_full[s], 1);
        mbar_init(&tmem_empty[s], 4);  // one arrive per epilogue warp
      }
      mbar_fence_init();
    }
    __syncwarp();
    tmem_alloc(tmem_base_slot, Cfg::TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_base_slot;

  if (warp == 0) {
    // ------------------------------ TMA producer ------------------------------
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int w = blockIdx.x; w < total_work; w += gridDim.x) {
        const int tile = w % tiles_mn;
        const int split = w / tiles_mn;
        const int m0 = (tile / p.n_tiles) * BLOCK_M;
        const int n0 = (tile % p.n_tiles) * out_cols_per_tile;
        const int kb0 = split * p.kb_per_split;
        const int kb1 = min(kb0 + p.kb_per_split, p.num_kb);
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          mbar_expect_tx(&full_bar[stage], Cfg::STAGE_BYTES);
          uint8_t* sa = smem_a + stage * Cfg::A_BYTES;
          uint8_t* sb = smem_b + stage * Cfg::B_BYTES;
          const int k0 = kb * BLOCK_K;
          if (!a_mn) {
            tma_load_2d(sa, &tmap_a, &full_bar[stage], k0, m0);  // box {64 k, 128 rows}
          } else {
#pragma unroll
            for (int j = 0; j < BLOCK_M / 64; ++j)  // box {64 m, 64 k}
              tma_load_2d(sa + j * 8192, &tmap_a, &full_bar[stage], m0 + j * 64, k0);
          }
          if (geglu) {  // box {64 k, BLOCK_N/2 rows} twice: value rows then gate rows
            tma_load_2d(sb, &tmap_b, &full_bar[stage], k0, n0);
            tma_load_2d(sb + Cfg::B_BYTES / 2, &tmap_b, &full_bar[stage], k0, (int)p.geglu_ipad + n0);
          } else if (!b_mn) {
            tma_load_2d(sb, &tmap_b, &full_bar[stage], k0, n0);  // box {64 k, BLOCK_N rows}
          } else {
#pragma unroll
            for (int j = 0; j < BLOCK_N / 64; ++j)  // box {64 n, 64 k}
              tma_load_2d(sb + j * 8192, &tmap_b, &full_bar[stage], n0 + j * 64, k0);
          }
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------ MMA issuer ------------------------------
    if (lane == 0) {
      const uint32_t idesc = umma_idesc_bf16(BLOCK_M, BLOCK_N, a_mn, b_mn);
      // K-major: advance 16 elements = 32 B inside the swizzle atom; SBO = 1024 B (8 rows x 128 B).
      // MN-major: advance 16 k-rows = 2048 B; LBO = 8192 B (next 64-wide MN atom), SBO = 1024 B.
      const uint32_t a_step = a_mn ? 2048 : 32, a_lbo = a_mn ? 8192 : 16;
      const uint32_t b_step = b_mn ? 2048 : 32, b_lbo = b_mn ? 8192 : 16;
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int w = blockIdx.x; w < total_work; w += gridDim.x) {
        const int split = w / tiles_mn;
        const int kb0 = split * p.kb_per_split;
        const int kb1 = min(kb0 + p.kb_per_split, p.num_kb);
        mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + acc * BLOCK_N;
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem_a + stage * Cfg::A_BYTES);
          const uint32_t sb = smem_u32(smem_b + stage * Cfg::B_BYTES);
#pragma unroll
          for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
            const uint64_t da = umma_smem_desc(sa + k * a_step, a_lbo, 1024);
            const uint64_t db = umma_smem_desc(sb + k * b_step, b_lbo, 1024);
            umma_bf16(tmem_d, da, db, idesc, (kb > kb0) || (k > 0));
          }
          umma_commit(&empty_bar[stage]);  // frees the smem slot once these MMAs have read it
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        umma_commit(&tmem_full[acc]);  // accumulator ready for the epilogue
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else {
    // ------------------------------ epilogue (warps 2..5) ------------------------------
    const int quarter = warp & 3;  // TMEM lane quarter this warp may access
    float* stg = stage_all + (warp - 2) * STG_FLOATS;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int w = blockIdx.x; w < total_work; w += gridDim.x) {
      const int tile = w % tiles_mn;
      const int m0 = (tile / p.n_tiles) * BLOCK_M;
      const int n0 = (tile % p.n_tiles) * out_cols_per_tile;
      // row bookkeeping (independent of the accumulator: done before waiting for the MMA)
      const int64_t row = (int64_t)m0 + quarter * 32 + lane;
      int64_t orow = -1;  // -1 marks a row beyond M
      if (row < p.M) orow = p.out_period > 0 ? (row / p.out_period) * p.out_batch_rows + (row % p.out_period) : row;
      const float* res_row = nullptr;
      if (EPI == EPI_F32 && p.residual != nullptr && row < p.M) {
        int64_t rr = row;
        if (p.res_period > 0) {
          rr = row % p.res_period;
          if (p.res_row_map) rr = p.res_row_map[rr];
        }
        res_row = (p.residual2 && rr >= p.res_split) ? p.residual2 + (rr - p.res_split) * p.ldr : p.residual + rr * p.ldr;
      }
      int64_t orow_i[8];
      const float* res_i[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int r = i * 4 + (lane >> 3);
        orow_i[i] = shfl_i64(orow, r);
        res_i[i] = (EPI == EPI_F32) ? reinterpret_cast<const float*>(shfl_i64(reinterpret_cast<int64_t>(res_row), r)) : nullptr;
      }
      mbar_wait(&tmem_full[acc], acc_phase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + acc * BLOCK_N;
      uint32_t raw[32];
      if (!geglu) {
#pragma unroll 1
        for (int c = 0; c < BLOCK_N / 32; ++c) {
          const int64_t col0 = (int64_t)n0 + c * 32;
          if (col0 >= p.N) break;  // warp-uniform
          tmem_ld_32x32(taddr + c * 32, raw);
          tmem_wait_ld();
          if (col0 + 32 <= p.N && p.fast_ok) {
            float4 resv[8];
            if (EPI == EPI_F32 && p.residual != nullptr) load_res8(resv, orow_i, res_i, col0 + (lane & 7) * 4);
            epi_chunk<EPI>(p, stg, raw, lane, orow_i, resv, col0);
          } else {
            epi_chunk_slow<EPI>(p, raw, raw, orow, res_row, col0, (int)min((int64_t)32, p.N - col0));
          }
        }
      } else {
        // GEGLU: accumulator columns [0, BLOCK_N/2) = value, [BLOCK_N/2, BLOCK_N) = gate of the same features
        uint32_t rawg[32];
#pragma unroll 1
        for (int c = 0; c < BLOCK_N / 64; ++c) {
          const int64_t col0 = (int64_t)n0 + c * 32;
          if (col0 >= p.N) break;
          tmem_ld_32x32(taddr + c * 32, raw);
          tmem_ld_32x32(taddr + BLOCK_N / 2 + c * 32, rawg);
          tmem_wait_ld();
          if (col0 + 32 <= p.N && p.fast_ok) epi_chunk_geglu(p, stg, raw, rawg, lane, orow_i, col0);
          else epi_chunk_slow<EPI_GEGLU>(p, raw, rawg, orow, nullptr, col0, (int)min((int64_t)32, p.N - col0));
        }
      }
      // release the accumulator stage back to the MMA warp
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty[acc]);
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  }
}

// ================================================================================================
// CTA-pair variant (cta_group::2): a 2-CTA cluster owns one 256 x 256 output tile.  Each CTA loads its own 128 rows
// of A and 128 of the 256 B rows per k-block (32 KB per stage per CTA instead of 48 KB for the same 128 x 256 MMA
// work), which is what lifts the kernel off the L2 -> SM bandwidth ceiling the single-CTA tile sits on.  The leader
// (cluster rank 0) issues tcgen05.mma.cta_group::2; both CTAs run 8 epilogue warps over their own 128 accumulator rows.
//   barriers:  full[s]   leader only   1 arrival (leader's expect_tx) + the bytes of BOTH CTAs' TMA loads
//              empty[s]  each CTA      multicast tcgen05.commit from the leader
//              tmem_full[a]  each CTA  multicast tcgen05.commit from the leader
//              tmem_empty[a] leader    16 arrivals: 8 epilogue warps x 2 CTAs (remote arrive from the peer)
// ================================================================================================
constexpr int GEMM2_THREADS = 320;     // warp 0 TMA, warp 1 MMA / TMEM owner, warps 2..9 epilogue
constexpr int G2_BN = 256;
constexpr int G2_A_BYTES = 128 * BLOCK_K * 2;
constexpr int G2_B_BYTES = 128 * BLOCK_K * 2;
constexpr int G2_STAGE_BYTES = G2_A_BYTES + G2_B_BYTES;
// TS = "TMA-store epilogue" (bf16 / GEGLU outputs with a plain row mapping): every epilogue warp pulls its whole share
// of the accumulator out of TMEM in one go (a tcgen05.ld issued while MMAs are queued completes late, so serial
// load -> use -> load round trips made the epilogue the pacing stage), releases the accumulator, packs bf16 into a
// 128B-swizzled 32-row box in shared memory and hands it to the TMA store engine.
// x = alpha * acc + bias on 64 accumulator columns held in registers (bias nullable; uniform loads)
__device__ __forceinline__ void bias_scale64(uint32_t (&v)[64], const float* bias, float alpha) {
#pragma unroll
  for (int t = 0; t < 64; ++t) {
    float x = __uint_as_float(v[t]) * alpha;
    if (bias != nullptr) x += __ldg(bias + t);
    v[t] = __float_as_uint(x);
  }
}

// GEGLU backward fused into the dgrad GEMM that produces dg = dY . W2 (zorro_utils.py:115-128 under autograd): the
// epilogue warps TMA-load the saved pre-activation boxes [value | gate] of their accumulator share into shared memory
// while the tile's MMAs run, turn them IN PLACE into [dvalue | dgate] = [dg * gelu(gate) | dg * value * gelu'(gate)]
// and hand the same boxes to the TMA store engine.  dg never goes to HBM (it was a bf16 [M, I] write + read) and the
// elementwise pass over [M, 2I] rides in the shadow of the MMAs.
template <int EPI, bool TS>
struct G2Cfg {
  static constexpr bool GB = (EPI == EPI_GEGLU_BWD);
  // epilogue warps per CTA: 2 per TMEM lane quarter (128 columns each); the GEGLU backward does ~30 FMA-pipe
  // instructions + 2 MUFU per accumulator element, so it runs 4 per quarter (64 columns each) to have the issue slots
  // GELU TMA-store epilogue (two outputs through the same boxes): 4 warps per quarter as well, 64 columns and ONE box each:
  // 83 -> 68 us on the decoders' fc1 (M = 50176, N = 1024, K = 256).  The plain bf16 kind measured no faster that way
  // (44.5 vs 46.2 us on the same shape, 37.9 vs 40.6 us at N = 768, K = 512: the chain is not per-warp pack latency) and
  // keeps 2 warps per quarter; -DMMF_TS_EPI16=1 builds it with 4 for A/B runs.
#ifndef MMF_TS_EPI16
#define MMF_TS_EPI16 0
#endif
  static constexpr bool WIDE16 = TS && (EPI == EPI_GELU || (EPI == EPI_BF16 && (MMF_TS_EPI16 != 0)));
  static constexpr int EPI_WARPS = (GB || WIDE16) ? 16 : 8;
  static constexpr int THREADS = 64 + 32 * EPI_WARPS;
  static constexpr int STAGES = GB ? 3 : (TS ? 5 : 6);
  // per epilogue warp: two 4 KB boxes / one 32x32 fp32 block
  static constexpr int STG_WARP = WIDE16 ? 4096 : ((TS || GB) ? 8192 : STG_FLOATS * 4);
  static constexpr int STG_BYTES = EPI_WARPS * STG_WARP;
  static constexpr int SMEM_BYTES = STAGES * G2_STAGE_BYTES + STG_BYTES + 1024 /*align*/ + 512 /*barriers*/;
};

// (gelu(g0) * v0, gelu(g1) * v1): gelu_fast (common.cuh) on a pair, polynomial and products as packed fp32x2
__device__ __forceinline__ void geglu_fwd_pair(float& v0, float& v1, float g0, float g1) {
  float t0, t1;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t0) : "f"(fmaf(0.3275911f * 0.70710678118654752f, fabsf(g0), 1.0f)));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t1) : "f"(fmaf(0.3275911f * 0.70710678118654752f, fabsf(g1), 1.0f)));
  const uint64_t t = f2_pack(t0, t1), x = f2_pack(g0, g1);
  uint64_t q = f2_fma(t, F2C(0.5f * 1.061405429f), F2C(0.5f * -1.453152027f));
  q = f2_fma(q, t, F2C(0.5f * 1.421413741f));
  q = f2_fma(q, t, F2C(0.5f * -0.284496736f));
  q = f2_fma(q, t, F2C(0.5f * 0.254829592f));
  float a0, a1, e0, e1;
  f2_unpack(f2_mul(f2_mul(x, x), F2C(-0.5f * 1.4426950408889634f)), a0, a1);
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e0) : "f"(a0));
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e1) : "f"(a1));
  float h0, h1;
  f2_unpack(f2_mul(f2_mul(q, t), f2_pack(e0, e1)), h0, h1);   // Phi(-|x|)
  const uint64_t cdf = f2_pack(g0 < 0.f ? h0 : 1.0f - h0, g1 < 0.f ? h1 : 1.0f - h1);
  f2_unpack(f2_mul(f2_mul(x, cdf), f2_pack(v0, v1)), v0, v1);
}

// GEGLU backward on one bf16 pair: (dvalue, dgate) = (dg * gelu(gate), dg * value * gelu'(gate)), exact-erf GELU through
// the same Abramowitz-Stegun erfc form as gelu_fast: Phi(-|x|) = q(t) * t * exp(-x^2/2) / 2, t = 1 / (1 + 0.3275911 |x| / sqrt2);
// gelu'(x) = Phi(x) + x * exp(-x^2/2) / sqrt(2 pi).  4 MUFU + ~22 issue slots per pair.
__device__ __forceinline__ void geglu_bwd_pair(uint32_t vraw, uint32_t graw, float d0, float d1, uint64_t alpha2, uint32_t& ov, uint32_t& og) {
  const float2 fv = unpack_bf16(vraw), fg = unpack_bf16(graw);
  float t0, t1;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t0) : "f"(fmaf(0.3275911f * 0.70710678118654752f, fabsf(fg.x), 1.0f)));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t1) : "f"(fmaf(0.3275911f * 0.70710678118654752f, fabsf(fg.y), 1.0f)));
  const uint64_t t = f2_pack(t0, t1), x = f2_pack(fg.x, fg.y);
  uint64_t q = f2_fma(t, F2C(0.5f * 1.061405429f), F2C(0.5f * -1.453152027f));
  q = f2_fma(q, t, F2C(0.5f * 1.421413741f));
  q = f2_fma(q, t, F2C(0.5f * -0.284496736f));
  q = f2_fma(q, t, F2C(0.5f * 0.254829592f));
  float a0, a1, e0, e1;
  f2_unpack(f2_mul(f2_mul(x, x), F2C(-0.5f * 1.4426950408889634f)), a0, a1);
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e0) : "f"(a0));
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e1) : "f"(a1));
  const uint64_t e = f2_pack(e0, e1);
  float h0, h1;
  f2_unpack(f2_mul(f2_mul(q, t), e), h0, h1);                 // Phi(-|x|)
  const uint64_t cdf = f2_pack(fg.x < 0.f ? h0 : 1.0f - h0, fg.y < 0.f ? h1 : 1.0f - h1);
  const uint64_t d = f2_mul(f2_pack(d0, d1), alpha2);
  float o0, o1;
  f2_unpack(f2_mul(d, f2_mul(x, cdf)), o0, o1);
  ov = pack_bf16(o0, o1);
  const uint64_t gp = f2_fma(f2_mul(x, e), F2C(0.3989422804014327f), cdf);
  f2_unpack(f2_mul(f2_mul(d, f2_pack(fv.x, fv.y)), gp), o0, o1);
  og = pack_bf16(o0, o1);
}

// explicit shared-window accesses: through uint8_t* the box pointers are generic to ptxas (ST.E.128 / LD.E.128 in the SASS)
__device__ __forceinline__ void sts128(uint32_t addr, uint4 v) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ uint4 lds128(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr) : "memory");
  return v;
}

// 32 columns of the fused GEGLU backward, in place on this lane's row of a (value, gate) box pair (32 rows x 64 bytes,
// CU_TENSOR_MAP_SWIZZLE_64B: 16-byte chunk index XOR address bits 7-8)
__device__ __forceinline__ void geglu_bwd_box32(uint8_t* vbox, uint8_t* gbox, const uint32_t* acc, int lane, float alpha) {
  const uint64_t alpha2 = f2_pack(alpha, alpha);
  const uint32_t vb = smem_u32(vbox), gb = smem_u32(gbox);
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int off = lane * 64 + ((j ^ ((lane >> 1) & 3)) << 4);
    const uint4 uv = lds128(vb + off);
    const uint4 ug = lds128(gb + off);
    const uint32_t* pv = &uv.x; const uint32_t* pg = &ug.x;
    uint4 ov, og;
    uint32_t* qv = &ov.x; uint32_t* qg = &og.x;
#pragma unroll
    for (int k = 0; k < 4; ++k)
      geglu_bwd_pair(pv[k], pg[k], __uint_as_float(acc[8 * j + 2 * k]), __uint_as_float(acc[8 * j + 2 * k + 1]), alpha2, qv[k], qg[k]);
    sts128(vb + off, ov);
    sts128(gb + off, og);
  }
}

// 64 fp32 accumulator columns of this lane's row -> bf16 -> one 32-row x 128-byte box, 16-byte groups XOR-swizzled
// with (row & 7) (= CU_TENSOR_MAP_SWIZZLE_128B for a 1024-byte aligned box)
__device__ __forceinline__ void pack_box(uint8_t* box, const uint32_t (&v)[64], int lane, const float* bias, float alpha) {
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    float x[8];
#pragma unroll
    for (int t = 0; t < 8; ++t) {
      x[t] = __uint_as_float(v[8 * j + t]) * alpha;
      if (bias != nullptr) x[t] += __ldg(bias + 8 * j + t);
    }
    sts128(smem_u32(box) + lane * 128 + ((j ^ (lane & 7)) << 4),
           make_uint4(pack_bf16(x[0], x[1]), pack_bf16(x[2], x[3]), pack_bf16(x[4], x[5]), pack_bf16(x[6], x[7])));
  }
}

template <int EPI, bool TS>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__((G2Cfg<EPI, TS>::THREADS), 1)
gemm2_tcgen05_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                     const __grid_constant__ CUtensorMap tmap_o, const __grid_constant__ CUtensorMap tmap_o2,
                     const GemmParams p) {
  constexpr bool geglu = (EPI == EPI_GEGLU);
  constexpr bool geglu_bwd = (EPI == EPI_GEGLU_BWD);
  constexpr int STAGES = G2Cfg<EPI, TS>::STAGES;
  static_assert(!TS || EPI == EPI_BF16 || EPI == EPI_GEGLU || EPI == EPI_GELU || EPI == EPI_GEGLU_BWD, "TMA-store epilogue: bf16 outputs only");
  static_assert(TS || !geglu_bwd, "the fused GEGLU backward exists for the TMA-store epilogue only");
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + STAGES * G2_A_BYTES;
  uint8_t* stage_bytes = smem + STAGES * G2_STAGE_BYTES;   // 1024-byte aligned (TMA-store boxes need it)
  uint64_t* bars = reinterpret_cast<uint64_t*>(stage_bytes + G2Cfg<EPI, TS>::STG_BYTES);
  uint64_t* full_bar = bars;                     // [STAGES]
  uint64_t* empty_bar = bars + STAGES;           // [STAGES]
  uint64_t* tmem_full = bars + 2 * STAGES;       // [2]
  uint64_t* tmem_empty = bars + 2 * STAGES + 2;  // [2]
  uint32_t* tmem_base_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4);
  uint64_t* box_bar = bars + 2 * STAGES + 6;     // [2 * EPI_WARPS] GEGLU backward: per epilogue warp, its two pre-activation half-box pairs landed
  float* stage_all = reinterpret_cast<float*>(stage_bytes);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int cluster_id = blockIdx.x >> 1;
  const int num_clusters = gridDim.x >> 1;
  constexpr int out_cols_per_tile = geglu ? G2_BN / 2 : G2_BN;
  const int tiles_mn = p.m_tiles * p.n_tiles;     // m_tiles counts 256-row pair tiles here
  const int total_work = tiles_mn * p.split_k;
  const bool a_mn = p.a_mn != 0, b_mn = p.b_mn != 0;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_b);
    if (TS) {
      tma_prefetch_desc(&tmap_o);
      if (geglu || geglu_bwd || EPI == EPI_GELU) tma_prefetch_desc(&tmap_o2);
    }
  }
  if (warp == 1) {
    if (lane == 0) {
      for (int s = 0; s < STAGES; ++s) {
        mbar_init(&full_bar[s], 1);
        mbar_init(&empty_bar[s], 1);
      }
      for (int s = 0; s < 2; ++s) {
        mbar_init(&tmem_full[s], 1);
        mbar_init(&tmem_empty[s], 2 * G2Cfg<EPI, TS>::EPI_WARPS);
      }
      if (geglu_bwd)
        for (int s = 0; s < 2 * G2Cfg<EPI, TS>::EPI_WARPS; ++s) mbar_init(&box_bar[s], 1);
      mbar_fence_init();
    }
    __syncwarp();
    tmem_alloc_cg2(tmem_base_slot, 512);
    tmem_relinquish_cg2();
  }
  tc_fence_before();
  cluster_sync_all();   // the peer's barriers are initialised before any remote arrive / multicast commit lands
  tc_fence_after();
  const uint32_t tmem_base = *tmem_base_slot;
  // Programmatic dependent launch: the grid may have been scheduled before the previous kernel of the stream has
  // finished (the prologue above -- barrier init, TMEM allocation, descriptor prefetch -- touches no global data);
  // everything below reads or writes memory other kernels own, so every thread waits for them here.
  asm volatile("griddepcontrol.wait;" ::: "memory");

  if (warp == 0) {
    // ------------------------------ TMA producer (both CTAs, each its own halves) ------------------------------
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      GCLK_DECL(cluster_id == 5 && leader);
      for (int w = cluster_id; w < total_work; w += num_clusters) {
        const int tile = w % tiles_mn;
        const int split = w / tiles_mn;
        const int m0 = (tile / p.n_tiles) * 256 + (int)rank * 128;
        const int n0 = (tile % p.n_tiles) * out_cols_per_tile;
        const int kb0 = split * p.kb_per_split;
        const int kb1 = min(kb0 + p.kb_per_split, p.num_kb);
        for (int kb = kb0; kb < kb1; ++kb) {
          GCLK(0);
          mbar_wait(&empty_bar[stage], phase ^ 1);
          GCLK(1);
          if (leader) mbar_expect_tx(&full_bar[stage], 2 * G2_STAGE_BYTES);
          uint8_t* sa = smem_a + stage * G2_A_BYTES;
          uint8_t* sb = smem_b + stage * G2_B_BYTES;
          const int k0 = kb * BLOCK_K;
          if (!a_mn) {
            tma_load_2d_cg2(sa, &tmap_a, &full_bar[stage], k0, m0);               // box {64 k, 128 rows}
          } else {
            tma_load_2d_cg2(sa, &tmap_a, &full_bar[stage], m0, k0);               // box {64 m, 64 k} x 2
            tma_load_2d_cg2(sa + 8192, &tmap_a, &full_bar[stage], m0 + 64, k0);
          }
          if (geglu) {   // leader: 128 value rows, peer: the 128 gate rows of the same features
            tma_load_2d_cg2(sb, &tmap_b, &full_bar[stage], k0, (leader ? 0 : (int)p.geglu_ipad) + n0);
          } else if (!b_mn) {
            tma_load_2d_cg2(sb, &tmap_b, &full_bar[stage], k0, n0 + (int)rank * 128);   // box {64 k, 128 rows}
          } else {
            tma_load_2d_cg2(sb, &tmap_b, &full_bar[stage], n0 + (int)rank * 128, k0);   // box {64 n, 64 k} x 2
            tma_load_2d_cg2(sb + 8192, &tmap_b, &full_bar[stage], n0 + (int)rank * 128 + 64, k0);
          }
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
      GCLK_FLUSH(20, 2);
    }
  } else if (warp == 1) {
    // ------------------------------ MMA issuer (leader CTA only) ------------------------------
    // The whole warp runs the loop in converged code and ONE elected lane issues: inside an `if (lane == 0)` region
    // nothing is provably warp-uniform and ptxas feeds every UTCHMMA's descriptors to the uniform datapath through an
    // R2UR + vote loop (~19 SASS instructions per MMA; found on the attention kernels).  Descriptors are
    // a per-stage base plus a constant step.
    if (leader) {
      const bool issuer = elect_one();
      const uint32_t idesc = umma_idesc_bf16(256, G2_BN, a_mn, b_mn);
      const uint32_t a_step = (a_mn ? 2048 : 32) >> 4, a_lbo = a_mn ? 8192 : 16;
      const uint32_t b_step = (b_mn ? 2048 : 32) >> 4, b_lbo = b_mn ? 8192 : 16;
      const uint64_t da0 = umma_smem_desc(__shfl_sync(0xffffffffu, smem_u32(smem_a), 0), a_lbo, 1024);
      const uint64_t db0 = umma_smem_desc(__shfl_sync(0xffffffffu, smem_u32(smem_b), 0), b_lbo, 1024);
      const uint32_t tmem0 = __shfl_sync(0xffffffffu, tmem_base, 0);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      GCLK_DECL(cluster_id == 5 && lane == 0);
      for (int w = cluster_id; w < total_work; w += num_clusters) {
        const int split = w / tiles_mn;
        const int kb0 = split * p.kb_per_split;
        const int kb1 = min(kb0 + p.kb_per_split, p.num_kb);
        GCLK(0);
        mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
        tc_fence_after();
        GCLK(1);
        const uint32_t tmem_d = tmem0 + acc * G2_BN;
        for (int kb = kb0; kb < kb1; ++kb) {
          GCLK(0);
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          GCLK(2);
          if (issuer) {
            const uint64_t da = da0 + (uint64_t)(stage * (G2_A_BYTES >> 4));
            const uint64_t db = db0 + (uint64_t)(stage * (G2_B_BYTES >> 4));
#pragma unroll
            for (int k = 0; k < BLOCK_K / UMMA_K; ++k)
              umma_bf16_cg2(tmem_d, da + (uint64_t)(k * a_step), db + (uint64_t)(k * b_step), idesc, (kb > kb0) || (k > 0));
            umma_commit_cg2(&empty_bar[stage]);   // frees this smem slot in both CTAs
          }
          __syncwarp();
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        if (issuer) umma_commit_cg2(&tmem_full[acc]);       // accumulator ready, both CTAs' epilogues
        __syncwarp();
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
#ifdef MMF_GEMM_CLOCKS
        if (dbg_on) g_gemm_clk[31] += 1;
#endif
      }
      GCLK_FLUSH(16, 3);
    }
  } else {
    // ------------------------------ epilogue (warps 2..9, both CTAs) ------------------------------
    const int ew = warp - 2;
    const int quarter = warp & 3;   // TMEM lane quarter this warp may access
    const int half = ew >> 2;       // which half of the tile's columns
    float* stg = stage_all + ew * STG_FLOATS;
    int acc = 0;
    uint32_t acc_phase = 0;
    if constexpr (TS) {
      uint8_t* sbox = stage_bytes + ew * G2Cfg<EPI, TS>::STG_WARP;   // two 4 KB boxes
      uint32_t box_phase = 0;
      GCLK_DECL(cluster_id == 5 && rank == 0 && ew == 5);   // the whole warp takes the clocks (a lane-0 condition would diverge the measured code)
      for (int w = cluster_id; w < total_work; w += num_clusters) {
        const int tile = w % tiles_mn;
        const int m0 = (tile / p.n_tiles) * 256 + (int)rank * 128;
        const int n0 = (tile % p.n_tiles) * out_cols_per_tile;
        const int row0 = m0 + quarter * 32;
        const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + acc * G2_BN;
        uint32_t v0[64], v1[64];
        if constexpr (geglu_bwd) {
          // This warp: rows [row0, +32), dg columns [cb, +64) as two 32-column halves h; half h works in place on the
          // box pair (value, gate) = u[:, cb + 32h ..], u[:, ipad + cb + 32h ..] held in its own 4 KB buffer.  The
          // boxes of the NEXT tile are requested as soon as a half's stores have been read out of shared memory, so
          // their HBM latency is covered by the other half's arithmetic and the next tile's accumulator wait.
          const int ipad = (int)p.geglu_ipad;
          const int part = ew >> 2;
          auto coords = [&](int ww, int& cb, int& r0, bool& lv) {
            const int t = ww % tiles_mn;
            cb = (t % p.n_tiles) * out_cols_per_tile + part * 64;
            r0 = (t / p.n_tiles) * 256 + (int)rank * 128 + quarter * 32;
            lv = ww < total_work && cb < p.N && r0 < p.M;   // warp-uniform
          };
          auto load_half = [&](int h, int cb, int r0) {
            if (GABL(8)) return;
            mbar_expect_tx(&box_bar[2 * ew + h], 4096);
            tma_load_2d(sbox + h * 4096, &tmap_o2, &box_bar[2 * ew + h], cb + 32 * h, r0);
            tma_load_2d(sbox + h * 4096 + 2048, &tmap_o2, &box_bar[2 * ew + h], ipad + cb + 32 * h, r0);
          };
          int cb, r0, cbn, r0n;
          bool live, live_n;
#ifdef MMF_GEMM_CLOCKS
          t_last = clock();
#endif
          coords(w, cb, r0, live);
          if (w == cluster_id && live && lane == 0) { load_half(0, cb, r0); load_half(1, cb, r0); }
          coords(w + num_clusters, cbn, r0n, live_n);
          GCLK(0);
          mbar_wait(&tmem_full[acc], acc_phase);
          tc_fence_after();
          GCLK(1);
          if (!live) {
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(&tmem_empty[acc], 0);
          }
          if (live) {
            // the whole 64-column share in ONE tcgen05.ld (a TMEM read costs ~1.5-3 k clk here whatever its width while the
            // other accumulator's MMAs run: two 32-column reads were 5.8 k of the 15.3 k clk per tile), accumulator released at once
            uint32_t v64[64];
            tmem_ld_32x64(taddr + part * 64, v64);
            tmem_wait_ld();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(&tmem_empty[acc], 0);
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              const uint32_t* vh = v64 + 32 * h;
              GCLK(2);
              if (!GABL(8)) mbar_wait(&box_bar[2 * ew + h], box_phase);
              GCLK(3 + h);
              uint8_t* hb = sbox + h * 4096;
              if (!GABL(2)) geglu_bwd_box32(hb, hb + 2048, vh, lane, p.alpha);
              fence_proxy_async_smem();
              __syncwarp();
              GCLK(5);
              if (lane == 0) {
                if (!GABL(4)) {
                  tma_store_2d(&tmap_o, hb, cb + 32 * h, r0);
                  tma_store_2d(&tmap_o, hb + 2048, ipad + cb + 32 * h, r0);
                }
                tma_store_commit();
                GCLK(6);
                if (live_n) {
                  tma_store_wait_read<0>();   // this half's boxes have left shared memory
                  GCLK(7);
                  load_half(h, cbn, r0n);
                }
              }
            }
            box_phase ^= 1;
          } else if (live_n && lane == 0) {   // (not reachable with row-major tile order; kept for safety)
            load_half(0, cbn, r0n); load_half(1, cbn, r0n);
          }
          if (++acc == 2) { acc = 0; acc_phase ^= 1; }
          GCLK(8);
          GCLK_FLUSH(0, 9);
          continue;
        }
#ifdef MMF_GEMM_CLOCKS
        t_last = clock();
#endif
        mbar_wait(&tmem_full[acc], acc_phase);
        tc_fence_after();
        GCLK(1);
        if constexpr (!geglu) {
          constexpr bool W16 = G2Cfg<EPI, TS>::WIDE16;   // 16 epilogue warps: `half` is 0..3 and selects 64 columns, one box per warp
          constexpr int EW_COLS = W16 ? 64 : 128;
          const int col_base = n0 + half * EW_COLS;
          const bool live0 = col_base < p.N && row0 < p.M, live1 = !W16 && col_base + 64 < p.N && row0 < p.M;   // warp-uniform
          if (live0) tmem_ld_32x64(taddr + half * EW_COLS, v0);
          if (live1) tmem_ld_32x64(taddr + half * EW_COLS + 64, v1);
          tmem_wait_ld();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive_cluster(&tmem_empty[acc], 0);   // accumulator released before any store work
          GCLK(2);
          if (live0) {
            if (lane == 0) tma_store_wait_read<0>();   // the previous tile's stores have drained this warp's boxes
            __syncwarp();
            GCLK(3);
            if constexpr (EPI == EPI_GELU) {
              // x = alpha * acc + bias; out2 (optional) = x, out = gelu(x): both leave through the same two boxes
              bias_scale64(v0, p.bias ? p.bias + col_base : nullptr, p.alpha);
              if (live1) bias_scale64(v1, p.bias ? p.bias + col_base + 64 : nullptr, p.alpha);
              if (p.out2) {
                pack_box(sbox, v0, lane, nullptr, 1.0f);
                if (live1) pack_box(sbox + 4096, v1, lane, nullptr, 1.0f);
                fence_proxy_async_smem();
                __syncwarp();
                if (lane == 0) {
                  tma_store_2d(&tmap_o2, sbox, col_base, row0);
                  if (live1) tma_store_2d(&tmap_o2, sbox + 4096, col_base + 64, row0);
                  tma_store_commit();
                }
              }
#pragma unroll
              for (int t = 0; t < 64; ++t) v0[t] = __float_as_uint(gelu_fast(__uint_as_float(v0[t])));
              if (live1) {
#pragma unroll
                for (int t = 0; t < 64; ++t) v1[t] = __float_as_uint(gelu_fast(__uint_as_float(v1[t])));
              }
              if (p.out2) {
                if (lane == 0) tma_store_wait_read<0>();
                __syncwarp();
              }
              pack_box(sbox, v0, lane, nullptr, 1.0f);
              if (live1) pack_box(sbox + 4096, v1, lane, nullptr, 1.0f);
            } else {
              pack_box(sbox, v0, lane, p.bias ? p.bias + col_base : nullptr, p.alpha);
              GCLK(6);
              if (live1) pack_box(sbox + 4096, v1, lane, p.bias ? p.bias + col_base + 64 : nullptr, p.alpha);
              GCLK(7);
            }
            fence_proxy_async_smem();
            GCLK(8);
            __syncwarp();
            GCLK(4);
            if (lane == 0) {
              tma_store_2d(&tmap_o, sbox, col_base, row0);
              if (live1) tma_store_2d(&tmap_o, sbox + 4096, col_base + 64, row0);
              tma_store_commit();
            }
            GCLK(5);
          }
          GCLK_FLUSH(0, 9);
        } else {
          // value columns [half*64, +64) and the gate columns of the same 64 features
          const int col_base = n0 + half * 64;
          const bool live = col_base < p.N && row0 < p.M;
          if (live) {
            tmem_ld_32x64(taddr + half * 64, v0);
            tmem_ld_32x64(taddr + G2_BN / 2 + half * 64, v1);
          }
          tmem_wait_ld();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive_cluster(&tmem_empty[acc], 0);
          if (live) {
            if (lane == 0) tma_store_wait_read<0>();
            __syncwarp();
            if (p.out2) {   // pre-activations [value | gate] for the backward pass
              pack_box(sbox, v0, lane, nullptr, 1.0f);
              pack_box(sbox + 4096, v1, lane, nullptr, 1.0f);
              fence_proxy_async_smem();
              __syncwarp();
              if (lane == 0) {
                tma_store_2d(&tmap_o2, sbox, col_base, row0);
                tma_store_2d(&tmap_o2, sbox + 4096, (int)p.geglu_ipad + col_base, row0);
                tma_store_commit();
              }
            }
            // Own basic block: left free, ptxas interleaves this math with the packing of the pre-activation boxes
            // above, keeps raw and activated values live together and spills (measured: 0.746 vs 0.679 ms at cfg 2).
            if (p.one) {
#pragma unroll
              for (int t = 0; t < 64; t += 2) {
                float a = __uint_as_float(v0[t]), b = __uint_as_float(v0[t + 1]);
                geglu_fwd_pair(a, b, __uint_as_float(v1[t]), __uint_as_float(v1[t + 1]));
                v0[t] = __float_as_uint(a); v0[t + 1] = __float_as_uint(b);
              }
            }
            if (lane == 0) tma_store_wait_read<0>();   // box 0 is read out (the GELU math above covered the wait)
            __syncwarp();
            pack_box(sbox, v0, lane, nullptr, 1.0f);
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) {
              tma_store_2d(&tmap_o, sbox, col_base, row0);
              tma_store_commit();
            }
          }
        }
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
      if (lane == 0) tma_store_wait_all<0>();   // stores complete before this CTA's shared memory goes away
    } else
    for (int w = cluster_id; w < total_work; w += num_clusters) {
      const int tile = w % tiles_mn;
      const int m0 = (tile / p.n_tiles) * 256 + (int)rank * 128;
      const int n0 = (tile % p.n_tiles) * out_cols_per_tile;
      const int64_t row = (int64_t)m0 + quarter * 32 + lane;
      int64_t orow = -1;
      if (row < p.M) orow = p.out_period > 0 ? (row / p.out_period) * p.out_batch_rows + (row % p.out_period) : row;
      const float* res_row = nullptr;
      if (EPI == EPI_F32 && p.residual != nullptr && row < p.M) {
        int64_t rr = row;
        if (p.res_period > 0) {
          rr = row % p.res_period;
          if (p.res_row_map) rr = p.res_row_map[rr];
        }
        res_row = (p.residual2 && rr >= p.res_split) ? p.residual2 + (rr - p.res_split) * p.ldr : p.residual + rr * p.ldr;
      }
      int64_t orow_i[8];
      const float* res_i[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int r = i * 4 + (lane >> 3);
        orow_i[i] = shfl_i64(orow, r);
        res_i[i] = (EPI == EPI_F32) ? reinterpret_cast<const float*>(shfl_i64(reinterpret_cast<int64_t>(res_row), r)) : nullptr;
      }
      const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + acc * G2_BN;
      uint32_t raw[32];
      if (!geglu) {
        const bool has_res = (EPI == EPI_F32) && p.residual != nullptr;
        const int cl = (lane & 7) * 4;
        float4 res_next[8];
        // the first chunk's residual is fetched while the MMA of this tile is still running
        if (has_res && p.fast_ok && (int64_t)n0 + half * 128 + 32 <= p.N) load_res8(res_next, orow_i, res_i, (int64_t)n0 + half * 128 + cl);
        mbar_wait(&tmem_full[acc], acc_phase);
        tc_fence_after();
#pragma unroll 1
        for (int cc = 0; cc < 4; ++cc) {
          const int c = half * 4 + cc;
          const int64_t col0 = (int64_t)n0 + c * 32;
          if (col0 < p.N) {   // warp-uniform
            tmem_ld_32x32(taddr + c * 32, raw);
            float4 resv[8];
            if (has_res) {
#pragma unroll
              for (int i = 0; i < 8; ++i) resv[i] = res_next[i];
              if (cc + 1 < 4 && p.fast_ok && col0 + 64 <= p.N) load_res8(res_next, orow_i, res_i, col0 + 32 + cl);
            }
            tmem_wait_ld();
            if (col0 + 32 <= p.N && p.fast_ok) epi_chunk<EPI>(p, stg, raw, lane, orow_i, resv, col0);
            else epi_chunk_slow<EPI>(p, raw, raw, orow, res_row, col0, (int)min((int64_t)32, p.N - col0));
          }
        }
      } else {
        uint32_t rawg[32];
        mbar_wait(&tmem_full[acc], acc_phase);
        tc_fence_after();
#pragma unroll 1
        for (int cc = 0; cc < 2; ++cc) {
          const int c = half * 2 + cc;
          const int64_t col0 = (int64_t)n0 + c * 32;
          if (col0 >= p.N) break;
          tmem_ld_32x32(taddr + c * 32, raw);
          tmem_ld_32x32(taddr + G2_BN / 2 + c * 32, rawg);
          tmem_wait_ld();
          if (col0 + 32 <= p.N && p.fast_ok) epi_chunk_geglu(p, stg, raw, rawg, lane, orow_i, col0);
          else epi_chunk_slow<EPI_GEGLU>(p, raw, rawg, orow, nullptr, col0, (int)min((int64_t)32, p.N - col0));
        }
      }
      // release the accumulator stage: the MMA warp of the LEADER waits for both CTAs' epilogue warps
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(&tmem_empty[acc], 0);
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  }

  tc_fence_before();
  cluster_sync_all();   // no CTA of the pair may free TMEM / exit while the peer still signals or reads
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_cg2(tmem_base, 512);
  }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  // resolved through the runtime so the library has no link-time dependency on libcuda.so
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

// 2-D bf16 tensor map over a row-major [rows, cols] matrix (cols contiguous), 128B swizzle.
static int make_tmap(CUtensorMap* map, const void* base, int64_t rows, int64_t cols, int64_t ld, int box_cols,
                     int box_rows, CUtensorMapSwizzle swizzle = CU_TENSOR_MAP_SWIZZLE_128B) {
  EncodeTiledFn enc = get_encode_fn();
  if (!enc) return 1000;
  cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t gstride[1] = {(cuuint64_t)ld * 2};
  cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), gdim, gstride, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : 2000 + (int)r;
}

static int num_sms() {     // per device (a process may drive more than one GPU)
  static int n[64] = {0};
  const int dev = current_device() & 63;
  int v = __atomic_load_n(&n[dev], __ATOMIC_RELAXED);
  if (v == 0) {
    cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev);
    __atomic_store_n(&n[dev], v, __ATOMIC_RELAXED);
  }
  return v;
}

template <int BLOCK_N, int EPI>
static int launch_gemm(const MmfGemmArgs& a, cudaStream_t stream) {
  using Cfg = GemmCfg<BLOCK_N>;
  const bool geglu = a.act == 2;
  const bool A_MN = a.a_mn != 0, B_MN = a.b_mn != 0;
  CUtensorMap ta, tb;
  int rc;
  if (!A_MN) rc = make_tmap(&ta, a.a, a.M, a.K, a.lda, BLOCK_K, BLOCK_M);
  else       rc = make_tmap(&ta, a.a, a.K, a.M, a.lda, 64, BLOCK_K);
  if (rc) return rc;
  const int64_t b_rows = geglu ? 2 * a.N : a.N;  // geglu: caller passes N = I_pad, B has 2*I_pad rows
  if (!B_MN) rc = make_tmap(&tb, a.b, b_rows, a.K, a.ldb, BLOCK_K, geglu ? BLOCK_N / 2 : BLOCK_N);
  else       rc = make_tmap(&tb, a.b, a.K, b_rows, a.ldb, 64, BLOCK_K);
  if (rc) return rc;

  GemmParams p;
  p.out = a.out; p.out2 = a.out2; p.bias = a.bias; p.residual = a.residual; p.residual2 = a.residual2; p.res_split = a.res_split; p.res_row_map = a.res_row_map;
  p.M = a.M; p.N = a.N; p.K = a.K; p.ldo = a.ldo; p.ldo2 = a.ldo2; p.ldr = a.ldr;
  p.out_f32 = a.out_f32; p.act = a.act; p.split_k = a.split_k < 1 ? 1 : a.split_k;
  p.res_period = a.res_period; p.out_period = a.out_period; p.out_batch_rows = a.out_batch_rows;
  p.accumulate = a.accumulate;
  p.a_mn = a.a_mn; p.b_mn = a.b_mn;
  p.m_tiles = (int)ceil_div64(a.M, BLOCK_M);
  p.n_tiles = (int)ceil_div64(a.N, geglu ? BLOCK_N / 2 : BLOCK_N);
  p.num_kb = (int)ceil_div64(a.K, BLOCK_K);
  if (p.split_k > p.num_kb) p.split_k = p.num_kb;
  p.kb_per_split = ceil_div(p.num_kb, p.split_k);
  p.split_k = ceil_div(p.num_kb, p.kb_per_split);  // no empty splits
  p.geglu_ipad = a.N;
  p.alpha = a.alpha;
  p.one = 1;
  auto al16 = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0; };
  p.fast_ok = (a.ldo % 4 == 0) && al16(a.out) && (!a.residual || (a.ldr % 4 == 0 && al16(a.residual))) &&
              (!a.residual2 || al16(a.residual2)) && (!a.bias || al16(a.bias)) && (!a.out2 || (a.ldo2 % 4 == 0 && al16(a.out2))) &&
              (a.act != 2 || a.N % 4 == 0);

  static DeviceOnce attr_once;
  const int attr_dev = current_device();
  auto kern = gemm_tcgen05_kernel<BLOCK_N, EPI>;
  if (!attr_once.done(attr_dev)) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES);
    if (e != cudaSuccess) return (int)e;
    attr_once.set(attr_dev);
  }
  const int64_t total = (int64_t)p.m_tiles * p.n_tiles * p.split_k;
  const int grid = (int)(total < num_sms() ? total : num_sms());
  kern<<<grid, GEMM_THREADS, Cfg::SMEM_BYTES, stream>>>(ta, tb, p);
  g_launch_count.fetch_add(1, std::memory_order_relaxed);
  MMF_LAUNCH_CHECK();
  return 0;
}

template <int EPI, bool TS>
static int launch_gemm2(const MmfGemmArgs& a, cudaStream_t stream) {
  const bool geglu = a.act == 2;
  const bool A_MN = a.a_mn != 0, B_MN = a.b_mn != 0;
  CUtensorMap ta, tb;
  int rc;
  if (!A_MN) rc = make_tmap(&ta, a.a, a.M, a.K, a.lda, BLOCK_K, 128);
  else       rc = make_tmap(&ta, a.a, a.K, a.M, a.lda, 64, BLOCK_K);
  if (rc) return rc;
  const int64_t b_rows = geglu ? 2 * a.N : a.N;
  if (!B_MN) rc = make_tmap(&tb, a.b, b_rows, a.K, a.ldb, BLOCK_K, 128);
  else       rc = make_tmap(&tb, a.b, a.K, b_rows, a.ldb, 64, BLOCK_K);
  if (rc) return rc;

  GemmParams p;
  p.out = a.out; p.out2 = a.out2; p.bias = a.bias; p.residual = a.residual; p.residual2 = a.residual2; p.res_split = a.res_split; p.res_row_map = a.res_row_map;
  p.M = a.M; p.N = a.N; p.K = a.K; p.ldo = a.ldo; p.ldo2 = a.ldo2; p.ldr = a.ldr;
  p.out_f32 = a.out_f32; p.act = a.act; p.split_k = a.split_k < 1 ? 1 : a.split_k;
  p.res_period = a.res_period; p.out_period = a.out_period; p.out_batch_rows = a.out_batch_rows;
  p.accumulate = a.accumulate;
  p.a_mn = a.a_mn; p.b_mn = a.b_mn;
  p.m_tiles = (int)ceil_div64(a.M, 256);
  p.n_tiles = (int)ceil_div64(a.N, geglu ? G2_BN / 2 : G2_BN);
  p.num_kb = (int)ceil_div64(a.K, BLOCK_K);
  if (p.split_k > p.num_kb) p.split_k = p.num_kb;
  p.kb_per_split = ceil_div(p.num_kb, p.split_k);
  p.split_k = ceil_div(p.num_kb, p.kb_per_split);
  p.geglu_ipad = a.N;
  p.alpha = a.alpha;
#ifdef MMF_GEMM_CLOCKS
  { const char* e = getenv("MMF_GEGLU_BWD_ABL"); p.epi_flags = e ? atoi(e) : 0; }
#else
  p.epi_flags = 0;
#endif
  p.one = 1;
  auto al16 = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0; };
  p.fast_ok = (a.ldo % 4 == 0) && al16(a.out) && (!a.residual || (a.ldr % 4 == 0 && al16(a.residual))) &&
              (!a.residual2 || al16(a.residual2)) && (!a.bias || al16(a.bias)) && (!a.out2 || (a.ldo2 % 4 == 0 && al16(a.out2))) &&
              (a.act != 2 || a.N % 4 == 0);
  CUtensorMap to = ta, to2 = ta;   // placeholders unless TS
  if (TS) {   // bf16 outputs as 32-row x 64-column boxes
    if (EPI == EPI_GEGLU_BWD) {   // out = [dvalue | dgate], out2 = [value | gate]: both [M, 2N], 32 x 32 boxes
      if ((rc = make_tmap(&to, a.out, a.M, 2 * a.N, a.ldo, 32, 32, CU_TENSOR_MAP_SWIZZLE_64B))) return rc;
      if ((rc = make_tmap(&to2, a.out2, a.M, 2 * a.N, a.ldo2, 32, 32, CU_TENSOR_MAP_SWIZZLE_64B))) return rc;
    } else {
      if ((rc = make_tmap(&to, a.out, a.M, a.N, a.ldo, 64, 32))) return rc;
      if (a.out2 && (rc = make_tmap(&to2, a.out2, a.M, geglu ? 2 * a.N : a.N, a.ldo2, 64, 32))) return rc;
    }
  }

  static DeviceOnce attr_once;
  const int attr_dev = current_device();
  auto kern = gemm2_tcgen05_kernel<EPI, TS>;
  constexpr int SMEM = G2Cfg<EPI, TS>::SMEM_BYTES;
  if (!attr_once.done(attr_dev)) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM);
    if (e != cudaSuccess) return (int)e;
    attr_once.set(attr_dev);
  }
  const int64_t total = (int64_t)p.m_tiles * p.n_tiles * p.split_k;
  int max_clusters = (num_sms() - g_reserved_sms.load(std::memory_order_relaxed)) / 2;
  if (max_clusters < 1) max_clusters = 1;
  const int clusters = (int)(total < max_clusters ? total : max_clusters);
  static const bool pdl = !(getenv("MMF_PDL") && atoi(getenv("MMF_PDL")) == 0);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(2 * clusters);
  cfg.blockDim = dim3(G2Cfg<EPI, TS>::THREADS);
  cfg.dynamicSmemBytes = SMEM;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl ? 1 : 0;
  cudaError_t le = cudaLaunchKernelEx(&cfg, kern, ta, tb, to, to2, p);
  if (le != cudaSuccess) return (int)le;
  g_launch_count.fetch_add(1, std::memory_order_relaxed);
  MMF_LAUNCH_CHECK();
  return 0;
}

}  // namespace mmf

extern "C" int mmf_gemm_bf16(const MmfGemmArgs* args, mmf_stream_t stream_) {
  using namespace mmf;
  if (!args) MMF_BAD_ARG(1);
  const MmfGemmArgs& a = *args;
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  if (!a.a || !a.b || !a.out) MMF_BAD_ARG(2);
  if (a.M <= 0 || a.N <= 0 || a.K <= 0) MMF_BAD_ARG(3);
  if (a.M > INT32_MAX || a.N > INT32_MAX || a.K > INT32_MAX) MMF_BAD_ARG(4);
  // TMA: 16-byte aligned base and row pitch
  if ((reinterpret_cast<uintptr_t>(a.a) & 15) || (reinterpret_cast<uintptr_t>(a.b) & 15)) MMF_BAD_ARG(5);
  if ((a.lda & 7) || (a.ldb & 7)) MMF_BAD_ARG(6);
  if (a.lda < (a.a_mn ? a.M : a.K) || a.ldb < (a.b_mn ? a.N : a.K) || a.ldo < a.N) MMF_BAD_ARG(7);
  if (a.split_k > 1 && (!a.out_f32 || a.act != 0 || a.bias || a.residual || a.out2)) MMF_BAD_ARG(8);
  if (a.act < 0 || a.act > 3) MMF_BAD_ARG(9);
  if (a.act == 3) {   // fused GEGLU backward: CTA-pair TMA-store kernel only, N = I_pad (a multiple of 64)
    auto al16 = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0; };
    if (a.out_f32 || a.bias || a.residual || a.split_k > 1 || a.accumulate || !a.out2 || (a.N & 63) || a.out_period ||
        (a.ldo & 7) || (a.ldo2 & 7) || a.ldo < 2 * a.N || a.ldo2 < 2 * a.N || !al16(a.out) || !al16(a.out2) ||
        (a.block_n != 0 && a.block_n != 256))
      MMF_BAD_ARG(21);
    return launch_gemm2<EPI_GEGLU_BWD, true>(a, stream);
  }
  if (a.act == 2 && (a.b_mn || a.out_f32 || a.bias || a.residual || (a.N & 7) || (a.out2 && (a.ldo2 & 7)) ||
                     (a.ldo & 7) || a.split_k > 1))
    MMF_BAD_ARG(10);
  if (a.residual && a.ldr < a.N) MMF_BAD_ARG(12);
  if (a.out_period < 0 || a.res_period < 0) MMF_BAD_ARG(13);
  if (a.accumulate && (a.act == 2 || a.out2)) MMF_BAD_ARG(16);

  int bn = a.block_n;
  if (bn == 0) {
    if (a.act == 2) bn = 256;
    else {
      const int64_t pad256 = ceil_div64(a.N, 256) * 256, pad128 = ceil_div64(a.N, 128) * 128;
      bn = (pad256 == pad128) ? 256 : 128;
    }
  }
  if (bn != 128 && bn != 256) MMF_BAD_ARG(14);
  if (a.act == 2 && bn != 256) MMF_BAD_ARG(15);
  int epi;
  if (a.act == 2) epi = EPI_GEGLU;
  else if (a.split_k > 1 || (a.accumulate && a.out_f32)) epi = EPI_ATOMIC;
  else if (a.accumulate) epi = EPI_BF16_ACC;
  else if (a.out_f32) epi = EPI_F32;
  else if (a.act == 1) epi = EPI_GELU;
  else epi = EPI_BF16;
  if (epi == EPI_F32 && a.act == 1) MMF_BAD_ARG(17);            // GELU is provided for bf16 outputs only
  if (epi != EPI_F32 && a.residual) MMF_BAD_ARG(18);            // the residual is an fp32 stream
  if ((epi == EPI_ATOMIC || epi == EPI_BF16_ACC) && a.bias) MMF_BAD_ARG(19);
  if (epi == EPI_BF16 && a.out2) MMF_BAD_ARG(11);
#define MMF_DISPATCH(BN)                                                              \
  switch (epi) {                                                                      \
    case EPI_BF16: return launch_gemm<BN, EPI_BF16>(a, stream);                       \
    case EPI_GELU: return launch_gemm<BN, EPI_GELU>(a, stream);                       \
    case EPI_F32: return launch_gemm<BN, EPI_F32>(a, stream);                         \
    case EPI_ATOMIC: return launch_gemm<BN, EPI_ATOMIC>(a, stream);                   \
    case EPI_BF16_ACC: return launch_gemm<BN, EPI_BF16_ACC>(a, stream);               \
    default: break;                                                                   \
  }
  static const bool pair_on = !(getenv("MMF_GEMM_2CTA") && atoi(getenv("MMF_GEMM_2CTA")) == 0);
  if (bn == 256 && pair_on) {   // CTA-pair kernel (cta_group::2, 256 x 256 tile per cluster)
    static const bool ts_on = !(getenv("MMF_GEMM_TMA_STORE") && atoi(getenv("MMF_GEMM_TMA_STORE")) == 0);
    auto al16 = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0; };
    const bool ts = ts_on && a.out_period == 0 && (a.ldo & 7) == 0 && al16(a.out) && (!a.out2 || ((a.ldo2 & 7) == 0 && al16(a.out2)));
    switch (epi) {
      case EPI_BF16: return ts ? launch_gemm2<EPI_BF16, true>(a, stream) : launch_gemm2<EPI_BF16, false>(a, stream);
      case EPI_GELU: return ts ? launch_gemm2<EPI_GELU, true>(a, stream) : launch_gemm2<EPI_GELU, false>(a, stream);
      case EPI_F32: return launch_gemm2<EPI_F32, false>(a, stream);
      case EPI_ATOMIC: return launch_gemm2<EPI_ATOMIC, false>(a, stream);
      case EPI_BF16_ACC: return launch_gemm2<EPI_BF16_ACC, false>(a, stream);
      case EPI_GEGLU: return ts ? launch_gemm2<EPI_GEGLU, true>(a, stream) : launch_gemm2<EPI_GEGLU, false>(a, stream);
      default: break;
    }
  }
  if (bn == 256) {
    if (epi == EPI_GEGLU) return launch_gemm<256, EPI_GEGLU>(a, stream);
    MMF_DISPATCH(256)
  }
  MMF_DISPATCH(128)
  MMF_BAD_ARG(20);
#undef MMF_DISPATCH
}

extern "C" int mmf_abi_version(void) { return 1; }
extern "C" void mmf_set_gemm_reserved_sms(int32_t n) { mmf::g_reserved_sms.store(n < 0 ? 0 : n); }
extern "C" int64_t mmf_launch_count(void) { return mmf::g_launch_count.load(); }
extern "C" void mmf_reset_launch_count(void) { mmf::g_launch_count.store(0); }

#ifdef MMF_GEMM_CLOCKS
extern "C" int mmf_debug_gemm_clocks(unsigned long long* out32, int reset) {
  if (reset) {
    unsigned long long z[32] = {0};
    return (int)cudaMemcpyToSymbol(mmf::g_gemm_clk, z, sizeof(z));
  }
  return (int)cudaMemcpyFromSymbol(out32, mmf::g_gemm_clk, 32 * sizeof(unsigned long long));
}
#endif
