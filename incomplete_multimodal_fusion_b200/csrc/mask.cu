// Mask sampling downstream of the random draws (reference: MultiMAE.generate_random_masks, multimae.py:182-255, and the
// per-modality token selection / zorro-mask bookkeeping of forward, multimae.py:378-426) in ONE single-CTA launch.
//
// The reference does this with ~25 tiny ATen launches on ONE row of <= 1024 keys (masks are shared by the whole batch)
// plus three .nonzero() host syncs.  The random draws themselves stay torch calls on the host side (same generator, same
// shapes, same order -> the same noise as the reference); everything after the noise is integer / ordering work:
//   per task t:   order = argsort(noise1_t);  pre[p] = order[p] < want_t ? 0 : 1        want_t = round(share_t * nenc)
//   all tasks:    ids_shuffle = argsort(pre + noise2);  ids_restore = argsort(ids_shuffle);  ids_keep = ids_shuffle[:nenc]
//                 mask[p] = ids_restore[p] < nenc ? 0 : 1
//   per task t:   idx_t = ascending positions with mask == 0 (the nonzero() of multimae.py:378-382), counts, the zorro
//                 segment table [0, c0, c0+c1, .., nenc, nenc+F] and slotmap[t][p] = rank of p in idx_t or -1
// argsort = rank counting in shared memory (n^2 compares, n <= 4096): rank_i = #{j : key_j < key_i or (key_j == key_i and
// j < i)}, i.e. a STABLE sort; torch's CUDA argsort leaves the order of exactly-equal keys unspecified.
#include "common.cuh"
#include "mmf_b200.h"
#include <atomic>

namespace mmf {
extern std::atomic<int64_t> g_launch_count;

constexpr int MASK_MAX_TASKS = 8;
constexpr int MASK_MAX_N = 4096;
constexpr int MASK_THREADS = 1024;

struct MaskParams {
  const float* noise1;   // [n_total] per-task noise, tasks concatenated
  const float* noise2;   // [n_total]
  const float* share;    // [T] Dirichlet sample
  int T, n_total, nenc, n_fusion;
  int off[MASK_MAX_TASKS + 1];
  int64_t* mask;         // [n_total] 0 = visible
  int64_t* ids_restore;  // [n_total]
  int64_t* ids_keep;     // [nenc]
  int32_t* idx;          // [n_total]: task t's ascending visible positions at off[t] .. off[t] + counts[t]
  int32_t* counts;       // [T]
  int32_t* seg;          // [T + 2]
  int32_t* slotmap;      // [T, n_fusion] or null
  int32_t* tok;          // [nenc] or null: global ids (off[t] + position) of the visible tokens in encoder order
  const int64_t* given;  // explicit mode: [n_total] caller's mask row (0 = visible); null = sampled mode
  int32_t* err;          // explicit mode: set to 1 when the mask keeps a number of tokens different from nenc
};

__global__ void __launch_bounds__(MASK_THREADS) mask_build_kernel(const MaskParams p) {
  __shared__ float key[MASK_MAX_N];
  __shared__ int32_t order[MASK_MAX_N];    // per-task argsort, later ids_restore
  __shared__ uint8_t keep[MASK_MAX_N];
  __shared__ int32_t s_counts[MASK_MAX_TASKS];
  const int n = p.n_total;
  if (threadIdx.x < MASK_MAX_TASKS) s_counts[threadIdx.x] = 0;
  if (p.given != nullptr) {
    // ---- explicit masks (multimae.py:372-376): ids_shuffle = STABLE argsort of the 0 / 1 row, i.e. a stable partition;
    // the reference's CUDA argsort leaves the order among equal keys unspecified (documented difference) ----
    __shared__ int s_nzero;
    for (int i = threadIdx.x; i < n; i += blockDim.x) keep[i] = p.given[i] == 0;
    if (threadIdx.x == 0) s_nzero = 0;
    __syncthreads();
    for (int i = threadIdx.x; i < n; i += blockDim.x) atomicAdd(&s_nzero, (int)keep[i]);
    __syncthreads();
    const int nzero = s_nzero;
    if (threadIdx.x == 0 && p.err) *p.err = nzero != p.nenc;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
      int before = 0;                       // equal keys before i
      for (int j = 0; j < i; ++j) before += keep[j] == keep[i];
      const int r = keep[i] ? before : nzero + before;
      p.ids_restore[i] = r;
      if (r < p.nenc) p.ids_keep[r] = i;
      if (p.mask) p.mask[i] = keep[i] ? 0 : 1;
      order[i] = r;
    }
    __syncthreads();
    // the tables below always describe exactly nenc tokens (the first nenc of the partition order = ids_keep), so that a
    // wrong num_encoded_tokens -- reported through *err -- can never make a consumer run past its buffers
    for (int i = threadIdx.x; i < n; i += blockDim.x) keep[i] = order[i] < p.nenc;
    __syncthreads();
  } else {
  // ---- per-task argsort of noise1 ----
  for (int i = threadIdx.x; i < n; i += blockDim.x) key[i] = p.noise1[i];
  __syncthreads();
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    int t = 0;
    while (i >= p.off[t + 1]) ++t;
    const int a = p.off[t], e = p.off[t + 1];
    const float k = key[i];
    int r = 0;
    for (int j = a; j < e; ++j) r += (key[j] < k) || (key[j] == k && j < i);
    order[a + r] = i - a;
  }
  __syncthreads();
  // ---- pre-mask and the global shuffle key ----
  float k2 = 0.f;   // (n <= 4 * blockDim: keep up to 4 keys per thread in registers)
  float k2v[MASK_MAX_N / MASK_THREADS];
  int c = 0;
  for (int i = threadIdx.x; i < n; i += blockDim.x, ++c) {
    int t = 0;
    while (i >= p.off[t + 1]) ++t;
    const long long want = (long long)rintf(p.share[t] * (float)p.nenc);   // torch.round: half to even
    const float pre = (long long)order[i] < want ? 0.f : 1.f;
    k2v[c] = pre + p.noise2[i];
  }
  __syncthreads();
  c = 0;
  for (int i = threadIdx.x; i < n; i += blockDim.x, ++c) key[i] = k2v[c];
  __syncthreads();
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    k2 = key[i];
    int r = 0;
    for (int j = 0; j < n; ++j) r += (key[j] < k2) || (key[j] == k2 && j < i);
    order[i] = r;   // rank of i = ids_restore[i]
    p.ids_restore[i] = r;
    if (r < p.nenc) p.ids_keep[r] = i;
    keep[i] = r < p.nenc;
    p.mask[i] = r < p.nenc ? 0 : 1;
  }
  __syncthreads();
  }
  // ---- per-task ascending index lists, counts, slot map ----
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    int t = 0;
    while (i >= p.off[t + 1]) ++t;
    const int a = p.off[t];
    int before = 0;
    for (int j = a; j < i; ++j) before += keep[j];
    if (keep[i]) {
      p.idx[a + before] = i - a;
      atomicAdd(&s_counts[t], 1);
    }
    if (p.slotmap && i - a < p.n_fusion) p.slotmap[t * p.n_fusion + (i - a)] = keep[i] ? before : -1;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    int acc = 0;
    p.seg[0] = 0;
    for (int t = 0; t < p.T; ++t) {
      p.counts[t] = s_counts[t];
      acc += s_counts[t];
      p.seg[t + 1] = acc;
    }
    p.seg[p.T + 1] = acc + p.n_fusion;
  }
  if (p.tok) {   // encoder order = modality-major, ascending position = ascending global id: rank among the kept ids
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
      if (!keep[i]) continue;
      int before = 0;
      for (int j = 0; j < i; ++j) before += keep[j];
      if (before < p.nenc) p.tok[before] = i;
    }
  }
}

}  // namespace mmf

static int mask_launch(mmf::MaskParams& p, int32_t T, const int32_t* sizes, int32_t nenc, int32_t n_fusion, bool slot, mmf_stream_t stream) {
  using namespace mmf;
  if (T <= 0 || T > MASK_MAX_TASKS) MMF_BAD_ARG(2);
  p.T = T; p.nenc = nenc; p.n_fusion = n_fusion;
  p.off[0] = 0;
  for (int t = 0; t < T; ++t) {
    if (sizes[t] < 0) MMF_BAD_ARG(3);
    if (slot && sizes[t] != n_fusion) MMF_BAD_ARG(6);   // the slot map is per fusion-token position
    p.off[t + 1] = p.off[t] + sizes[t];
  }
  for (int t = T; t < MASK_MAX_TASKS; ++t) p.off[t + 1] = p.off[T];
  p.n_total = p.off[T];
  if (p.n_total <= 0 || p.n_total > MASK_MAX_N) MMF_BAD_ARG(4);
  if (nenc < 0 || nenc > p.n_total) MMF_BAD_ARG(5);
  mask_build_kernel<<<1, MASK_THREADS, 0, reinterpret_cast<cudaStream_t>(stream)>>>(p);
  g_launch_count.fetch_add(1, std::memory_order_relaxed);
  MMF_LAUNCH_CHECK();
  return 0;
}

extern "C" int mmf_mask_build(const float* noise1, const float* noise2, const float* share, int32_t T, const int32_t* sizes,
                              int32_t nenc, int32_t n_fusion, int64_t* mask, int64_t* ids_restore, int64_t* ids_keep,
                              int32_t* idx, int32_t* counts, int32_t* seg, int32_t* slotmap, int32_t* tok, mmf_stream_t stream) {
  using namespace mmf;
  if (!noise1 || !noise2 || !share || !sizes || !mask || !ids_restore || !ids_keep || !idx || !counts || !seg) MMF_BAD_ARG(1);
  MaskParams p;
  p.noise1 = noise1; p.noise2 = noise2; p.share = share;
  p.mask = mask; p.ids_restore = ids_restore; p.ids_keep = ids_keep; p.idx = idx; p.counts = counts; p.seg = seg; p.slotmap = slotmap;
  p.tok = tok; p.given = nullptr; p.err = nullptr;
  return mask_launch(p, T, sizes, nenc, n_fusion, slotmap != nullptr, stream);
}

extern "C" int mmf_mask_explicit(const int64_t* given, int32_t T, const int32_t* sizes, int32_t nenc, int32_t n_fusion,
                                 int64_t* ids_restore, int64_t* ids_keep, int32_t* idx, int32_t* counts, int32_t* seg,
                                 int32_t* slotmap, int32_t* tok, int32_t* err, mmf_stream_t stream) {
  using namespace mmf;
  if (!given || !sizes || !ids_restore || !ids_keep || !idx || !counts || !seg || !err) MMF_BAD_ARG(1);
  MaskParams p;
  p.noise1 = nullptr; p.noise2 = nullptr; p.share = nullptr;
  p.mask = nullptr; p.ids_restore = ids_restore; p.ids_keep = ids_keep; p.idx = idx; p.counts = counts; p.seg = seg; p.slotmap = slotmap;
  p.tok = tok; p.given = given; p.err = err;
  return mask_launch(p, T, sizes, nenc, n_fusion, slotmap != nullptr, stream);
}
