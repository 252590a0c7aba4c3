// Full-rank evaluation, exact fp32 CUDA-core path (model.py:122-127 predict; trainer.py:146-170 eval):
//   scores = U . I^T accumulated with fmaf in d = 0..D-1 order (bit-identical to oracle/oracle_c.c),
//   train/val-item masking, banned item range, per-row top-K ordered (score desc, item id asc).
// The fused kernel never writes the b x n_items score matrix: every block streams item tiles against a resident tile
// of users and keeps, per user, a shared-memory candidate list guarded by a rising threshold (the K-th best so far).
// The tcgen05 candidate pass (score_tc.cu) reuses the merge / final-select kernels of this file.
#include <float.h>
#include "common.cuh"

namespace b200rec {

constexpr int TU = 64;  // users per block (dense kernel; the top-K kernel uses 16 * UPT)
constexpr int TI = 64;  // items per tile
constexpr int KMAX = 128;

struct Cand {
  float s;
  int id;
};
__device__ __forceinline__ bool cand_before(const Cand& a, const Cand& b) {  // (score desc, id asc)
  return a.s > b.s || (a.s == b.s && a.id < b.id);
}
// warp-cooperative bitonic sort of np (power of two) candidates in shared memory, best first
__device__ __forceinline__ void warp_sort(Cand* c, int np, int lane) {
  for (int k = 2; k <= np; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int t = lane; t < np; t += 32) {
        const int o = t ^ j;
        if (o > t) {
          const bool up = ((t & k) == 0);
          const Cand a = c[t], b = c[o];
          if (cand_before(b, a) == up) { c[t] = b; c[o] = a; }
        }
      }
      __syncwarp();
    }
  }
}

__device__ __forceinline__ bool csr_row_has(const int32_t* __restrict__ ptr, const int32_t* __restrict__ idx, int64_t row,
                                            int key) {
  if (!ptr) return false;
  int lo = __ldg(ptr + row), hi = __ldg(ptr + row + 1);
  const int end = hi;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (__ldg(idx + mid) < key) lo = mid + 1; else hi = mid;
  }
  return lo < end && __ldg(idx + lo) == key;
}

struct TopkParams {
  const float* rep_users;
  const int64_t* users;
  int n_users;
  const float* rep_items;
  int n_items;
  const int32_t *excl_ptr_a, *excl_idx_a, *excl_ptr_b, *excl_idx_b;
  int banned_lo, banned_hi;
  int k;
  int cap;        // candidate slots per user row (pow2 >= k + TI)
  int n_split;    // item range splits (grid.y)
  Cand* partial;  // [n_split, n_users, k]
};

// dynamic smem: us[D][TUK] | it[D][TI] | cand[TUK][cap] | thr[TUK] | cnt[TUK];  TUK = 16 * UPT users per block
template <int D, int UPT>
__global__ void __launch_bounds__(256) score_topk_f32_kernel(const TopkParams p) {
  constexpr int TUK = 16 * UPT;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float* us = reinterpret_cast<float*>(smem_raw);
  float* it = us + D * TUK;
  Cand* cand = reinterpret_cast<Cand*>(it + D * TI);
  float* thr = reinterpret_cast<float*>(cand + (size_t)TUK * p.cap);
  int* cnt = reinterpret_cast<int*>(thr + TUK);
  __shared__ int64_t s_user[TUK];

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int tx = tid & 15, ty = tid >> 4;  // 16 x 16 threads, UPT users x 4 items each
  const int u0 = blockIdx.x * TUK;
  const int cap = p.cap, K = p.k;
  // item range of this split, aligned to TI
  const int tiles_total = (p.n_items + TI - 1) / TI;
  const int tiles_per = (tiles_total + p.n_split - 1) / p.n_split;
  const int tile_lo = blockIdx.y * tiles_per, tile_hi = min(tiles_total, tile_lo + tiles_per);

  if (tid < TUK) {
    const int u = u0 + tid;
    s_user[tid] = (u < p.n_users) ? p.users[u] : -1;
    thr[tid] = -INFINITY;
    cnt[tid] = 0;
  }
  __syncthreads();
  // user tile, transposed to [d][user]
  for (int e = tid; e < TUK * (D / 4); e += 256) {
    const int r = e / (D / 4), c4 = e % (D / 4);
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (s_user[r] >= 0) v = ldg_f4(p.rep_users + (size_t)s_user[r] * D + c4 * 4);
    us[(c4 * 4 + 0) * TUK + r] = v.x; us[(c4 * 4 + 1) * TUK + r] = v.y;
    us[(c4 * 4 + 2) * TUK + r] = v.z; us[(c4 * 4 + 3) * TUK + r] = v.w;
  }

  for (int tile = tile_lo; tile < tile_hi; ++tile) {
    const int i0 = tile * TI;
    __syncthreads();  // previous tile fully consumed (and user tile stored)
    for (int e = tid; e < TI * (D / 4); e += 256) {
      const int r = e / (D / 4), c4 = e % (D / 4);
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (i0 + r < p.n_items) v = ldg_f4(p.rep_items + (size_t)(i0 + r) * D + c4 * 4);
      it[(c4 * 4 + 0) * TI + r] = v.x; it[(c4 * 4 + 1) * TI + r] = v.y;
      it[(c4 * 4 + 2) * TI + r] = v.z; it[(c4 * 4 + 3) * TI + r] = v.w;
    }
    __syncthreads();
    float acc[UPT][4];
#pragma unroll
    for (int a = 0; a < UPT; ++a)
#pragma unroll
      for (int b = 0; b < 4; ++b) acc[a][b] = 0.f;
#pragma unroll 8
    for (int d = 0; d < D; ++d) {
      float ua[UPT];
#pragma unroll
      for (int a = 0; a < UPT; ++a) ua[a] = us[d * TUK + ty * UPT + a];
      const float4 vv = *reinterpret_cast<const float4*>(it + d * TI + tx * 4);
      const float va[4] = {vv.x, vv.y, vv.z, vv.w};
#pragma unroll
      for (int a = 0; a < UPT; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) acc[a][b] = fmaf(ua[a], va[b], acc[a][b]);
    }
    // candidates
#pragma unroll
    for (int a = 0; a < UPT; ++a) {
      const int r = ty * UPT + a;
      const int64_t user = s_user[r];
      if (user < 0) continue;
      const float t = thr[r];
#pragma unroll
      for (int b = 0; b < 4; ++b) {
        const int item = i0 + tx * 4 + b;
        const float s = acc[a][b];
        if (item < p.n_items && s > t) {
          if (item >= p.banned_lo && item < p.banned_hi) continue;
          if (csr_row_has(p.excl_ptr_a, p.excl_idx_a, user, item)) continue;
          if (csr_row_has(p.excl_ptr_b, p.excl_idx_b, user, item)) continue;
          const int slot = atomicAdd(&cnt[r], 1);
          cand[(size_t)r * cap + slot] = Cand{s, item};  // cap >= k + TI and compaction below keep slot < cap
        }
      }
    }
    __syncthreads();
    // compaction: rows that could overflow on the next tile keep their K best and raise the threshold
    for (int r = warp; r < TUK; r += 8) {
      const int c = cnt[r];
      if (c > cap - TI) {
        Cand* row = cand + (size_t)r * cap;
        for (int t2 = c + lane; t2 < cap; t2 += 32) row[t2] = Cand{-INFINITY, INT32_MAX};
        __syncwarp();
        warp_sort(row, cap, lane);
        if (lane == 0) {
          cnt[r] = min(c, K);
          if (c >= K) thr[r] = row[K - 1].s;
        }
      }
    }
  }
  __syncthreads();
  // final: sort every row, emit K (padding: -inf / -1 when fewer than K unmasked items were seen)
  for (int r = warp; r < TUK; r += 8) {
    if (s_user[r] < 0) continue;
    const int c = cnt[r];
    Cand* row = cand + (size_t)r * cap;
    for (int t2 = c + lane; t2 < cap; t2 += 32) row[t2] = Cand{-INFINITY, INT32_MAX};
    __syncwarp();
    warp_sort(row, cap, lane);
    Cand* out = p.partial + ((size_t)blockIdx.y * p.n_users + (u0 + r)) * K;
    for (int j = lane; j < K; j += 32) out[j] = (j < c) ? row[j] : Cand{-INFINITY, -1};
  }
}

// merge the per-split partial lists of one user (n_split * k candidates, k <= 128) -> final K
__global__ void __launch_bounds__(128) topk_merge_kernel(const Cand* __restrict__ partial, int n_split, int n_users, int k,
                                                         int np, int32_t* __restrict__ out_ids,
                                                         float* __restrict__ out_scores) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  Cand* buf = reinterpret_cast<Cand*>(smem_raw);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int u = blockIdx.x * 4 + warp;
  if (u >= n_users) return;
  Cand* row = buf + (size_t)warp * np;
  const int total = n_split * k;
  for (int t = lane; t < np; t += 32) {
    Cand c{-INFINITY, INT32_MAX};
    if (t < total) {
      c = partial[((size_t)(t / k) * n_users + u) * k + (t % k)];
      if (c.id < 0) c.id = INT32_MAX;
    }
    row[t] = c;
  }
  __syncwarp();
  warp_sort(row, np, lane);
  for (int j = lane; j < k; j += 32) {
    const Cand c = row[j];
    out_ids[(size_t)u * k + j] = (c.id == INT32_MAX) ? -1 : c.id;
    out_scores[(size_t)u * k + j] = c.s;
  }
}

// dense scores for predict(): same fmaf chain, written out
template <int D>
__global__ void __launch_bounds__(256) score_dense_kernel(const float* __restrict__ rep_users, const int64_t* __restrict__ users,
                                                          int n_users, const float* __restrict__ rep_items, int n_items,
                                                          float* __restrict__ scores) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float* us = reinterpret_cast<float*>(smem_raw);
  float* it = us + D * TU;
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int u0 = blockIdx.y * TU, i0 = blockIdx.x * TI;
  for (int e = tid; e < TU * (D / 4); e += 256) {
    const int r = e / (D / 4), c4 = e % (D / 4);
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (u0 + r < n_users) v = ldg_f4(rep_users + (size_t)users[u0 + r] * D + c4 * 4);
    us[(c4 * 4 + 0) * TU + r] = v.x; us[(c4 * 4 + 1) * TU + r] = v.y;
    us[(c4 * 4 + 2) * TU + r] = v.z; us[(c4 * 4 + 3) * TU + r] = v.w;
  }
  for (int e = tid; e < TI * (D / 4); e += 256) {
    const int r = e / (D / 4), c4 = e % (D / 4);
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (i0 + r < n_items) v = ldg_f4(rep_items + (size_t)(i0 + r) * D + c4 * 4);
    it[(c4 * 4 + 0) * TI + r] = v.x; it[(c4 * 4 + 1) * TI + r] = v.y;
    it[(c4 * 4 + 2) * TI + r] = v.z; it[(c4 * 4 + 3) * TI + r] = v.w;
  }
  __syncthreads();
  float acc[4][4];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) acc[a][b] = 0.f;
#pragma unroll 8
  for (int d = 0; d < D; ++d) {
    const float4 uu = *reinterpret_cast<const float4*>(us + d * TU + ty * 4);
    const float4 vv = *reinterpret_cast<const float4*>(it + d * TI + tx * 4);
    const float ua[4] = {uu.x, uu.y, uu.z, uu.w}, va[4] = {vv.x, vv.y, vv.z, vv.w};
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int b = 0; b < 4; ++b) acc[a][b] = fmaf(ua[a], va[b], acc[a][b]);
  }
#pragma unroll
  for (int a = 0; a < 4; ++a) {
    const int u = u0 + ty * 4 + a;
    if (u >= n_users) continue;
#pragma unroll
    for (int b = 0; b < 4; ++b) {
      const int i = i0 + tx * 4 + b;
      if (i < n_items) scores[(size_t)u * n_items + i] = acc[a][b];
    }
  }
}

__global__ void hit_matrix_kernel(const int32_t* __restrict__ rec, int n_rows, int k, int64_t user0,
                                  const int32_t* __restrict__ eval_ptr, const int32_t* __restrict__ eval_idx,
                                  float* __restrict__ hit) {
  const int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (t >= (int64_t)n_rows * k) return;
  const int64_t r = t / k;
  const int id = rec[t];
  hit[t] = (id >= 0 && csr_row_has(eval_ptr, eval_idx, user0 + r, id)) ? 1.f : 0.f;
}

// Precision / Recall / NDCG @k for a list of cut-offs in one pass (trainer.py:123-143): one thread per user walks its K
// recommended ids once, keeping the hit count and the discounted gains; at every requested k it adds its three ratios to
// per-thread sums.  Sums are fp64 and reduced in a fixed order (warp butterfly, warps in order, one partial per block):
// the caller adds the block partials.  Users without eval items do not count (user_masks, trainer.py:139).
constexpr int METRIC_MAX_KS = 32;
__global__ void __launch_bounds__(128) rank_metrics_kernel(const int32_t* __restrict__ rec, int n_rows, int k, int64_t user0,
                                                           const int32_t* __restrict__ eval_ptr,
                                                           const int32_t* __restrict__ eval_idx,
                                                           const int32_t* __restrict__ topks, int nk, double* block_sums) {
  __shared__ double s_part[4][3 * METRIC_MAX_KS + 1];
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  double acc[3 * METRIC_MAX_KS + 1];
#pragma unroll
  for (int i = 0; i < 3 * METRIC_MAX_KS + 1; ++i) acc[i] = 0.0;
  if (r < n_rows) {
    const int b = eval_ptr[user0 + r], e = eval_ptr[user0 + r + 1];
    const int n_eval = e - b;
    if (n_eval > 0) {
      int n_hit = 0, ti = 0;
      double dcg = 0.0, idcg = 0.0;
      for (int j = 0; j < k && ti < nk; ++j) {
        const int id = rec[(size_t)r * k + j];
        const double gain = 1.0 / log2((double)(j + 2));
        if (id >= 0 && csr_row_has(eval_ptr, eval_idx, user0 + r, id)) {
          ++n_hit;
          dcg += gain;
        }
        if (j < n_eval) idcg += gain;
#pragma unroll 1
        while (ti < nk && topks[ti] == j + 1) {
          acc[3 * ti + 0] = (double)n_hit / (double)(j + 1);
          acc[3 * ti + 1] = (double)n_hit / (double)n_eval;
          acc[3 * ti + 2] = dcg / idcg;
          ++ti;
        }
      }
      acc[3 * nk] = 1.0;
    }
  }
  for (int i = 0; i <= 3 * nk; ++i) {
    double v = acc[i];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0) s_part[warp][i] = v;
  }
  __syncthreads();
  for (int i = threadIdx.x; i <= 3 * nk; i += blockDim.x)
    block_sums[(size_t)blockIdx.x * (3 * nk + 1) + i] = ((s_part[0][i] + s_part[1][i]) + s_part[2][i]) + s_part[3][i];
}

static int next_pow2(int v) {
  int p = 1;
  while (p < v) p <<= 1;
  return p;
}
static int topk_upt(int d) { return d >= 256 ? 2 : 4; }
static int choose_split(int n_users, int n_items, int k, int d) {
  const int user_tiles = ceil_div(n_users, 16 * topk_upt(d));
  const int tiles_total = ceil_div(n_items, TI);
  int split = ceil_div(2 * 148, user_tiles);
  if (split > tiles_total) split = tiles_total;
  if (split < 1) split = 1;
  if (split > 64) split = 64;
  if (split > 2048 / k) split = 2048 / k;  // merge list must fit shared memory
  if (split < 1) split = 1;
  return split;
}

}  // namespace b200rec
using namespace b200rec;

extern "C" int64_t b200rec_score_topk_workspace(int32_t n_batch_users, int32_t n_items, int32_t d, int32_t k,
                                                int32_t precision);
// score_tc.cu
int64_t b200rec_score_topk_tc_workspace(int nb, int ni, int d);
bool b200rec_score_topk_tc_supported(int d);
int b200rec_score_topk_tc(const float* rep_users, const int64_t* users, int nb, const float* rep_items, int ni, int d,
                          const int32_t* ea_ptr, const int32_t* ea_idx, const int32_t* eb_ptr, const int32_t* eb_idx, int blo,
                          int bhi, int k, int32_t* out_ids, float* out_scores, int32_t* out_overflow, void* workspace,
                          cudaStream_t st);

extern "C" int b200rec_score_dense_f32(const float* rep_users, const int64_t* users, int32_t n_batch_users,
                                       const float* rep_items, int32_t n_items, int32_t d, float* scores, void* stream) {
  B2_REQUIRE(rep_users && users && rep_items && scores && n_batch_users > 0 && n_items > 0, "bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  dim3 grid(ceil_div(n_items, TI), ceil_div(n_batch_users, TU));
#define B2_DENSE(DD)                                                                                               \
  case DD: {                                                                                                       \
    const size_t smem = (size_t)DD * (TU + TI) * sizeof(float);                                                    \
    B2_CUDA(cudaFuncSetAttribute(score_dense_kernel<DD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
    score_dense_kernel<DD><<<grid, 256, smem, st>>>(rep_users, users, n_batch_users, rep_items, n_items, scores);  \
  } break;
  switch (d) {
    B2_DENSE(16) B2_DENSE(32) B2_DENSE(64) B2_DENSE(128) B2_DENSE(256)
    default: return fail(B200REC_ERR_UNSUPPORTED, "%s: %s", __func__, "embedding size must be 16/32/64/128/256");
  }
#undef B2_DENSE
  B2_LAUNCHED();
  return 0;
}

extern "C" int64_t b200rec_score_topk_workspace(int32_t n_batch_users, int32_t n_items, int32_t d, int32_t k,
                                                int32_t precision) {
  (void)precision;
  if (k < 1) k = 1;
  if (precision == 1 && b200rec_score_topk_tc_supported(d)) return b200rec_score_topk_tc_workspace(n_batch_users, n_items, d);
  const int split = choose_split(n_batch_users, n_items, k, d);
  return (int64_t)split * n_batch_users * k * (int64_t)sizeof(Cand) + 256;
}

extern "C" int b200rec_score_topk(const float* rep_users, const int64_t* users, int32_t n_batch_users,
                                  const float* rep_items, int32_t n_items, int32_t d, const int32_t* excl_ptr_a,
                                  const int32_t* excl_idx_a, const int32_t* excl_ptr_b, const int32_t* excl_idx_b,
                                  int32_t banned_lo, int32_t banned_hi, int32_t k, int32_t precision, int32_t* out_ids,
                                  float* out_scores, int32_t* out_overflow, void* workspace, void* stream) {
  B2_REQUIRE(rep_users && users && rep_items && out_ids && out_scores && workspace, "null argument");
  B2_REQUIRE(n_batch_users > 0 && n_items > 0, "empty input");
  B2_REQUIRE(k >= 1 && k <= KMAX, "k must be in [1,128]");
  B2_REQUIRE(!excl_ptr_a || excl_idx_a, "exclusion CSR a incomplete");
  B2_REQUIRE(!excl_ptr_b || excl_idx_b, "exclusion CSR b incomplete");
  B2_REQUIRE(precision == 0 || precision == 1, "precision must be 0 (exact fp32) or 1 (tcgen05 bf16 candidates + exact re-score)");
  cudaStream_t st = (cudaStream_t)stream;
  if (precision == 1) {
    B2_REQUIRE(b200rec_score_topk_tc_supported(d), "precision 1 supports embedding size 64 / 128");
    B2_REQUIRE(out_overflow, "precision 1 needs out_overflow");
    return b200rec_score_topk_tc(rep_users, users, n_batch_users, rep_items, n_items, d, excl_ptr_a, excl_idx_a, excl_ptr_b,
                                 excl_idx_b, banned_lo, banned_hi, k, out_ids, out_scores, out_overflow, workspace, st);
  }
  if (out_overflow) B2_CUDA(cudaMemsetAsync(out_overflow, 0, (size_t)n_batch_users * sizeof(int32_t), st));
  TopkParams p;
  p.rep_users = rep_users; p.users = users; p.n_users = n_batch_users; p.rep_items = rep_items; p.n_items = n_items;
  p.excl_ptr_a = excl_ptr_a; p.excl_idx_a = excl_idx_a; p.excl_ptr_b = excl_ptr_b; p.excl_idx_b = excl_idx_b;
  p.banned_lo = banned_lo; p.banned_hi = banned_hi; p.k = k;
  p.cap = next_pow2(k + TI);
  p.n_split = choose_split(n_batch_users, n_items, k, d);
  p.partial = reinterpret_cast<Cand*>(workspace);
  const int upt = topk_upt(d), tuk = 16 * upt;
  dim3 grid(ceil_div(n_batch_users, tuk), p.n_split);
#define B2_TOPK(DD, UPT)                                                                                                   \
  case DD: {                                                                                                          \
    const size_t smem = (size_t)DD * (tuk + TI) * sizeof(float) + (size_t)tuk * p.cap * sizeof(Cand) + tuk * 8;       \
    B2_CUDA(cudaFuncSetAttribute(score_topk_f32_kernel<DD, UPT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
    score_topk_f32_kernel<DD, UPT><<<grid, 256, smem, st>>>(p);                                                            \
  } break;
  switch (d) {
    B2_TOPK(16, 4) B2_TOPK(32, 4) B2_TOPK(64, 4) B2_TOPK(128, 4) B2_TOPK(256, 2)
    default: return fail(B200REC_ERR_UNSUPPORTED, "%s: %s", __func__, "embedding size must be 16/32/64/128/256");
  }
#undef B2_TOPK
  B2_LAUNCHED();
  const int np = next_pow2(p.n_split * k);
  const size_t msmem = (size_t)4 * np * sizeof(Cand);
  B2_REQUIRE(msmem <= 200 * 1024, "merge list too large");
  B2_CUDA(cudaFuncSetAttribute(topk_merge_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)msmem));
  topk_merge_kernel<<<ceil_div(n_batch_users, 4), 128, msmem, st>>>(p.partial, p.n_split, n_batch_users, k, np, out_ids,
                                                                    out_scores);
  B2_LAUNCHED();
  return 0;
}

extern "C" int b200rec_rank_metrics(const int32_t* rec_ids, int32_t n_rows, int32_t k, int64_t user0, const int32_t* eval_ptr,
                                    const int32_t* eval_idx, const int32_t* topks, int32_t n_topks, double* block_sums,
                                    void* stream) {
  B2_REQUIRE(rec_ids && eval_ptr && eval_idx && topks && block_sums && n_rows > 0 && k > 0, "bad argument");
  B2_REQUIRE(n_topks > 0 && n_topks <= METRIC_MAX_KS, "1..32 cut-offs");
  rank_metrics_kernel<<<ceil_div(n_rows, 128), 128, 0, (cudaStream_t)stream>>>(rec_ids, n_rows, k, user0, eval_ptr, eval_idx,
                                                                              topks, n_topks, block_sums);
  B2_LAUNCHED();
  return 0;
}

extern "C" int b200rec_hit_matrix(const int32_t* rec_ids, int32_t n_rows, int32_t k, int64_t user0, const int32_t* eval_ptr,
                                  const int32_t* eval_idx, float* hit, void* stream) {
  B2_REQUIRE(rec_ids && eval_ptr && eval_idx && hit && n_rows > 0 && k > 0, "bad argument");
  hit_matrix_kernel<<<ceil_div((long long)n_rows * k, 256), 256, 0, (cudaStream_t)stream>>>(rec_ids, n_rows, k, user0,
                                                                                           eval_ptr, eval_idx, hit);
  B2_LAUNCHED();
  return 0;
}
