// BPR step kernels: device sampler, row gather / scatter-add, fused BPR forward+backward, layer-0 L2, Adam.
// Reference: dataset.py:119-131 (sampler), model.py:112-120 / :4046-4052 / :66-71 (bpr_forward),
// trainer.py:414-428 / :542-553 (loss, backward, optimiser step).
#include "common.cuh"

namespace b200rec {

// ---------------------------------------------------------------------------------------------- sampler
__device__ __forceinline__ void philox_pair(uint64_t seed, uint32_t slot, uint64_t step, uint32_t draw, uint64_t& a,
                                            uint64_t& b) {
  uint32_t c[4] = {slot, (uint32_t)step, draw, (uint32_t)(step >> 32)};
  Philox::gen(c, (uint32_t)seed, (uint32_t)(seed >> 32));
  a = (uint64_t)c[0] | ((uint64_t)c[1] << 32);
  b = (uint64_t)c[2] | ((uint64_t)c[3] << 32);
}
__device__ __forceinline__ bool row_contains(const int32_t* row, int n, int key) {
  int lo = 0, hi = n;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    const int v = __ldg(row + mid);
    if (v < key) lo = mid + 1; else hi = mid;
  }
  return lo < n && __ldg(row + lo) == key;
}
__global__ void bpr_sample_kernel(const int32_t* __restrict__ user_ptr, const int32_t* __restrict__ user_items,
                                  int n_users, int n_items, uint64_t seed, const int64_t* __restrict__ step_ptr,
                                  int batch, int64_t* __restrict__ out) {
  const int slot = blockIdx.x * blockDim.x + threadIdx.x;
  if (slot >= batch) return;
  const uint64_t step = (uint64_t)(*step_ptr);
  uint32_t draw = 0;
  uint64_t a, b;
  int user = 0, s = 0, deg = 0;
  for (;; ++draw) {  // user uniform, redrawn while its train row is empty (dataset.py:120-122)
    philox_pair(seed, (uint32_t)slot, step, draw, a, b);
    user = (int)__umul64hi(a, (uint64_t)n_users);
    s = __ldg(user_ptr + user);
    deg = __ldg(user_ptr + user + 1) - s;
    if (deg > 0 || draw > (1u << 20)) break;
  }
  const int32_t* row = user_items + s;
  const int pos = deg > 0 ? __ldg(row + (int)__umul64hi(b, (uint64_t)deg)) : 0;  // np.random.choice(train_data[user])
  int neg = 0;
  for (++draw;; ++draw) {  // negative uniform over items, rejected while in the row (dataset.py:126-128)
    philox_pair(seed, (uint32_t)slot, step, draw, a, b);
    neg = (int)__umul64hi(a, (uint64_t)n_items);
    if (!row_contains(row, deg, neg)) break;
    neg = (int)__umul64hi(b, (uint64_t)n_items);
    if (!row_contains(row, deg, neg) || draw > (1u << 20)) break;
  }
  out[3 * (size_t)slot + 0] = user;
  out[3 * (size_t)slot + 1] = pos;
  out[3 * (size_t)slot + 2] = neg;
}

// Edge-dropout keep mask of NGCF.dropout_sp_mat (model.py:4016-4021): keep[e] = floor((1-p) + r[e]) with r uniform in
// [0,1) at 24-bit resolution (torch.rand fp32).  The reference draws r on the CPU generator; the device stream is
// Philox4x32-10, key = seed, counter = (e >> 2, step_lo, step_hi, 0xD0), lane e & 3.  One thread fills one 32-bit word.
__global__ void dropout_mask_kernel(int nnz, float keep_base, uint64_t seed, const int64_t* __restrict__ step_ptr,
                                    uint32_t* __restrict__ bits) {
  const int w = blockIdx.x * blockDim.x + threadIdx.x;
  const int n_words = (nnz + 31) >> 5;
  if (w >= n_words) return;
  const uint64_t step = (uint64_t)(*step_ptr);
  uint32_t word = 0;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    uint32_t c[4] = {(uint32_t)(w * 8 + j), (uint32_t)step, (uint32_t)(step >> 32), 0xD0u};
    Philox::gen(c, (uint32_t)seed, (uint32_t)(seed >> 32));
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const float r = (float)(c[q] >> 8) * (1.0f / 16777216.0f);
      if (floorf(keep_base + r) >= 1.f) word |= 1u << (j * 4 + q);
    }
  }
  const int rem = nnz - w * 32;
  if (rem < 32) word &= (1u << rem) - 1u;
  bits[w] = word;
}

// rows a BPR batch touches: flags[u] = flags[off+pos] = flags[off+neg] = 1 (flags pre-zeroed)
__global__ void mark_rows_kernel(const int64_t* __restrict__ batch, int n_batch, int64_t item_offset, uint8_t* __restrict__ flags) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= 3 * n_batch) return;
  const int64_t r = batch[t] + ((t % 3) ? item_offset : 0);
  flags[r] = 1;
}

// Rows within one hop of a batch: the sampled rows themselves and every item a sampled user interacted with (one warp per
// sample walks the user's sorted train row).  The user half of `flags` is not written here: the caller keeps it all-ones
// (the users within a hop of the sampled items are nearly all users of a power-law graph, and a superset is always valid).
__global__ void mark_reach_kernel(const int64_t* __restrict__ batch, int n_batch, int64_t item_offset,
                                  const int32_t* __restrict__ user_ptr, const int32_t* __restrict__ user_items,
                                  uint8_t* __restrict__ flags) {
  const int s = (int)((blockIdx.x * (unsigned)blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
  if (s >= n_batch) return;
  const int64_t u = batch[3 * s];
  if (lane == 0) flags[u] = 1;
  if (lane == 1) flags[item_offset + batch[3 * s + 1]] = 1;
  if (lane == 2) flags[item_offset + batch[3 * s + 2]] = 1;
  const int lo = __ldg(user_ptr + u), hi = __ldg(user_ptr + u + 1);
  for (int k = lo + lane; k < hi; k += 32) flags[item_offset + __ldg(user_items + k)] = 1;
}

// ---------------------------------------------------------------------------------------------- gather / scatter
template <int G, int VPL>
__global__ void __launch_bounds__(256) gather_rows_kernel(const float* __restrict__ table, const int64_t* __restrict__ idx,
                                                          int64_t offset, int n, float* __restrict__ out,
                                                          float* __restrict__ sqnorm) {
  constexpr int D = G * VPL * 4;
  const int g = (int)((blockIdx.x * (unsigned)blockDim.x + threadIdx.x) / G), gl = threadIdx.x & (G - 1);
  const bool active = g < n;
  const int64_t row = active ? idx[g] + offset : 0;
  float ss = 0.f;
#pragma unroll
  for (int t = 0; t < VPL; ++t) {
    const int c = (gl + t * G) * 4;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (active) {
      v = ldg_f4(table + (size_t)row * D + c);
      st_f4(out + (size_t)g * D + c, v);
    }
    ss += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
  }
  ss = group_sum<G>(ss);
  if (active && sqnorm && gl == 0) sqnorm[g] = ss;
}
template <int G, int VPL>
__global__ void __launch_bounds__(256) scatter_add_rows_kernel(float* __restrict__ table, const int64_t* __restrict__ idx,
                                                               int64_t offset, int n, const float* __restrict__ src) {
  constexpr int D = G * VPL * 4;
  const int g = (int)((blockIdx.x * (unsigned)blockDim.x + threadIdx.x) / G), gl = threadIdx.x & (G - 1);
  if (g >= n) return;
  const int64_t row = idx[g] + offset;
#pragma unroll
  for (int t = 0; t < VPL; ++t) {
    const int c = (gl + t * G) * 4;
    red_add_f4(table + (size_t)row * D + c, ldg_f4(src + (size_t)g * D + c));
  }
}

// ---------------------------------------------------------------------------------------------- fused BPR
// Deterministic cross-block reduction: every block stores its partial, the last block to finish (ticket counter in
// scratch[0], self-resetting) adds the partials in block order.
__device__ __forceinline__ bool last_block_ticket(float* scratch) {
  __shared__ bool is_last;
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned* counter = reinterpret_cast<unsigned*>(scratch);
    const unsigned t = atomicAdd(counter, 1u);
    is_last = (t == gridDim.x - 1);
    if (is_last) *counter = 0u;
  }
  __syncthreads();
  if (is_last) __threadfence();
  return is_last;
}

__device__ __forceinline__ float softplus_t(float x) { return x > 20.f ? x : log1pf(expf(x)); }  // F.softplus defaults
// derivative torch uses for softplus: z/(z+1) with z = exp(x) (x <= threshold), else 1
__device__ __forceinline__ float sigmoid_t(float x) {
  if (x > 20.f) return 1.f;
  const float z = expf(x);
  return z / (z + 1.f);
}

template <int G, int VPL, bool HAS_W>
__global__ void __launch_bounds__(256) bpr_fused_kernel(const float* rep, const int64_t* batch,
                                                        int n_batch, int64_t item_offset, float l2_reg, int reg_mode,
                                                        const float* __restrict__ w, float loss_scale,
                                                        float* __restrict__ g_rep, float* __restrict__ g_w,
                                                        float* __restrict__ loss_out, float* scratch,
                                                        float* __restrict__ dots, int dots_mode, float loss_weight,
                                                        float* __restrict__ coef_out, const float* __restrict__ emb0,
                                                        float l2_emb0) {
  // emb0 != NULL: LightGCN's layer-0 regulariser (model.py:114-117) is folded into this launch's loss -- l2_emb0 * mean of
  // ||e_u||^2 + ||e_p||^2 + ||e_n||^2 over the layer-0 rows; its gradient is formed per distinct row by
  // bpr_grad_rows_kernel<L2ROWS> after the backward chain.
  // coef_out != NULL (ordered mode): the per-sample score derivative is stored and NO gradient row is scattered here;
  // bpr_grad_rows_kernel then forms every touched row's gradient once, in sample order (deterministic, no atomics).
  // dots_mode 0: everything in one launch.  Embedding-dimension sharding (each rank holds D/P columns) splits it:
  // 1 = write this rank's partial (pos, neg, l2) per sample to dots[B,3] and stop; after the caller's all-reduce,
  // 2 = take the full (pos, neg, l2) from dots and produce this rank's columns of the gradient.
  constexpr int D = G * VPL * 4;
  constexpr int SPB = 256 / G;  // samples per block
  __shared__ float s_loss[SPB];
  __shared__ float s_w[HAS_W ? SPB : 1][HAS_W ? D : 1];
  const int gib = threadIdx.x / G, gl = threadIdx.x & (G - 1);
  const int smp = blockIdx.x * SPB + gib;
  const bool active = smp < n_batch;
  pdl_trigger();
  pdl_wait();  // rep / batch / dots come from the previous kernels of the step
  int64_t ru = 0, rp = 0, rn = 0;
  if (active) {
    ru = batch[3 * (size_t)smp];
    rp = batch[3 * (size_t)smp + 1] + item_offset;
    rn = batch[3 * (size_t)smp + 2] + item_offset;
  }
  float4 u[VPL], p[VPL], n[VPL], wv[VPL];
  float pos = 0.f, neg = 0.f, l2 = 0.f;
#pragma unroll
  for (int t = 0; t < VPL; ++t) {
    const int c = (gl + t * G) * 4;
    u[t] = p[t] = n[t] = make_float4(0.f, 0.f, 0.f, 0.f);
    wv[t] = make_float4(1.f, 1.f, 1.f, 1.f);
    if (active) {
      u[t] = ldc_f4(rep + (size_t)ru * D + c);  // coherent loads: rep is the output of the kernel this one overlaps (PDL)
      p[t] = ldc_f4(rep + (size_t)rp * D + c);
      n[t] = ldc_f4(rep + (size_t)rn * D + c);
      if constexpr (HAS_W) wv[t] = ldg_f4(w + c);
    }
    pos += u[t].x * p[t].x * wv[t].x + u[t].y * p[t].y * wv[t].y + u[t].z * p[t].z * wv[t].z + u[t].w * p[t].w * wv[t].w;
    neg += u[t].x * n[t].x * wv[t].x + u[t].y * n[t].y * wv[t].y + u[t].z * n[t].z * wv[t].z + u[t].w * n[t].w * wv[t].w;
    l2 += u[t].x * u[t].x + u[t].y * u[t].y + u[t].z * u[t].z + u[t].w * u[t].w;
    l2 += p[t].x * p[t].x + p[t].y * p[t].y + p[t].z * p[t].z + p[t].w * p[t].w;
    l2 += n[t].x * n[t].x + n[t].y * n[t].y + n[t].z * n[t].z + n[t].w * n[t].w;
  }
  pos = group_sum<G>(pos);
  neg = group_sum<G>(neg);
  l2 = group_sum<G>(l2);
  if (dots_mode == 1) {
    if (active && gl == 0) {
      dots[3 * (size_t)smp] = pos; dots[3 * (size_t)smp + 1] = neg; dots[3 * (size_t)smp + 2] = l2;
    }
    return;
  }
  if (dots_mode == 2 && active) {
    pos = dots[3 * (size_t)smp]; neg = dots[3 * (size_t)smp + 1]; l2 = dots[3 * (size_t)smp + 2];
  }
  const float x = neg - pos;
  const float inv_b = 1.f / (float)n_batch;
  float loss = softplus_t(x);
  if (reg_mode == 1) loss += l2_reg * l2;
  loss *= loss_weight;
  if (emb0) {  // every rank adds its own columns' share of the layer-0 term
    float l0 = 0.f;
    if (active) {
#pragma unroll
      for (int t = 0; t < VPL; ++t) {
        const int c = (gl + t * G) * 4;
        const float4 a = ldc_f4(emb0 + (size_t)ru * D + c), b = ldc_f4(emb0 + (size_t)rp * D + c), e = ldc_f4(emb0 + (size_t)rn * D + c);
        l0 += a.x * a.x + a.y * a.y + a.z * a.z + a.w * a.w + b.x * b.x + b.y * b.y + b.z * b.z + b.w * b.w +
              e.x * e.x + e.y * e.y + e.z * e.z + e.w * e.w;
      }
    }
    l0 = group_sum<G>(l0);
    loss += l2_emb0 * l0;
  }
  const float coef = loss_scale * inv_b * sigmoid_t(x);
  const float rc = (reg_mode == 1) ? 2.f * l2_reg * loss_scale * inv_b : 0.f;
  if (coef_out && active && gl == 0) coef_out[smp] = coef;
  if (active) {
#pragma unroll
    for (int t = 0; t < VPL; ++t) {
      const int c = (gl + t * G) * 4;
      const float4 cw = make_float4(coef * wv[t].x, coef * wv[t].y, coef * wv[t].z, coef * wv[t].w);
      if (!coef_out) {
      float4 gu = make_float4(cw.x * (n[t].x - p[t].x) + rc * u[t].x, cw.y * (n[t].y - p[t].y) + rc * u[t].y,
                              cw.z * (n[t].z - p[t].z) + rc * u[t].z, cw.w * (n[t].w - p[t].w) + rc * u[t].w);
      float4 gp = make_float4(-cw.x * u[t].x + rc * p[t].x, -cw.y * u[t].y + rc * p[t].y, -cw.z * u[t].z + rc * p[t].z,
                              -cw.w * u[t].w + rc * p[t].w);
      float4 gn = make_float4(cw.x * u[t].x + rc * n[t].x, cw.y * u[t].y + rc * n[t].y, cw.z * u[t].z + rc * n[t].z,
                              cw.w * u[t].w + rc * n[t].w);
      red_add_f4(g_rep + (size_t)ru * D + c, gu);
      red_add_f4(g_rep + (size_t)rp * D + c, gp);
      red_add_f4(g_rep + (size_t)rn * D + c, gn);
      }
      if constexpr (HAS_W) {  // d(score)/dw = u * (n - p)
        s_w[gib][c + 0] = coef * u[t].x * (n[t].x - p[t].x);
        s_w[gib][c + 1] = coef * u[t].y * (n[t].y - p[t].y);
        s_w[gib][c + 2] = coef * u[t].z * (n[t].z - p[t].z);
        s_w[gib][c + 3] = coef * u[t].w * (n[t].w - p[t].w);
      }
    }
  } else if constexpr (HAS_W) {
#pragma unroll
    for (int t = 0; t < VPL; ++t) {
      const int c = (gl + t * G) * 4;
      s_w[gib][c] = s_w[gib][c + 1] = s_w[gib][c + 2] = s_w[gib][c + 3] = 0.f;
    }
  }
  if (gl == 0) s_loss[gib] = active ? loss : 0.f;
  __syncthreads();
  // block partials -> scratch[1 + block], scratch[1 + gridDim + block*D + d]
  float* part_loss = scratch + 4;
  float* part_w = scratch + 4 + gridDim.x;
  if (threadIdx.x == 0) {
    float acc = 0.f;
    for (int i = 0; i < SPB; ++i) acc += s_loss[i];
    part_loss[blockIdx.x] = acc;
  }
  if constexpr (HAS_W) {
    for (int d = threadIdx.x; d < D; d += blockDim.x) {
      float acc = 0.f;
      for (int i = 0; i < SPB; ++i) acc += s_w[i][d];
      part_w[(size_t)blockIdx.x * D + d] = acc;
    }
  }
  if (!last_block_ticket(scratch)) return;
  if (threadIdx.x == 0) {
    float acc = 0.f;
    for (unsigned i = 0; i < gridDim.x; ++i) acc += __ldcg(part_loss + i);
    *loss_out += loss_scale * acc * inv_b;
  }
  if (HAS_W) {
    for (int d = threadIdx.x; d < D; d += blockDim.x) {
      float acc = 0.f;
      for (unsigned i = 0; i < gridDim.x; ++i) acc += __ldcg(part_w + (size_t)i * D + d);
      g_w[d] = acc;
    }
  }
}

// ---------------------------------------------------------------------------------------------- ordered scatter
// Duplicate ids inside a batch are the rule (popular items; a user drawn twice), and a red.global.add per sample sums
// them in arrival order -- different bits from run to run.  Instead the 3B (row, slot) pairs of the batch are ranked once
// (bpr_group_rows_kernel, beside the forward layers on the side stream) and every DISTINCT row gets its
// gradient from one lane group that adds the contributions of its slots in slot order and issues one plain store: the
// aggregation north_star asks of the scatter (one write per distinct row instead of one atomic per sample), with a
// fixed summation order on top.
constexpr int GROUP_CAP = 12288;    // slots (3 * batch) the grouping holds in 48 KB of shared memory: batches up to 4096
constexpr int GROUP_EPW = 4;        // slots ranked per warp
constexpr unsigned GROUP_CONT = 0x80000000u;  // order[k] bit 31: position k continues the row of position k - 1
// Rank sort over the whole grid instead of a sort in one block: the position of slot e in (row, slot) order is the number
// of slots with a smaller (row, slot) -- every warp counts that for GROUP_EPW slots against all 3B rows held in shared
// memory (37.7 M compares at B = 2048, a few microseconds of the whole GPU, no serial stage) and stores
// order[position] = e, with bit 31 set when an earlier slot names the same row.  Rows are compared as int32: the tables
// of this library have < 2^31 rows (int32 column indices in b200rec_csr).
__global__ void __launch_bounds__(256) bpr_group_rows_kernel(const int64_t* __restrict__ batch, int n_slots, int64_t item_offset,
                                                             int32_t* __restrict__ order) {
  extern __shared__ int32_t s_row[];  // [n_slots rounded up to 32]
  const int n_pad = (n_slots + 31) & ~31;
  for (int i = threadIdx.x; i < n_pad; i += 256)
    s_row[i] = i < n_slots ? (int32_t)(batch[i] + ((i % 3) ? item_offset : 0)) : 0x7fffffff;
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int e0 = (blockIdx.x * 8 + (threadIdx.x >> 5)) * GROUP_EPW;
  if (e0 >= n_slots) return;
  int re[GROUP_EPW], lt[GROUP_EPW], eqb[GROUP_EPW];
#pragma unroll
  for (int j = 0; j < GROUP_EPW; ++j) {
    re[j] = e0 + j < n_slots ? s_row[e0 + j] : 0x7fffffff;
    lt[j] = eqb[j] = 0;
  }
#pragma unroll 4
  for (int f = lane; f < n_pad; f += 32) {
    const int rf = s_row[f];
#pragma unroll
    for (int j = 0; j < GROUP_EPW; ++j) {
      lt[j] += rf < re[j];
      eqb[j] += (rf == re[j]) & (f < e0 + j);
    }
  }
#pragma unroll
  for (int j = 0; j < GROUP_EPW; ++j) {
    const int l = __reduce_add_sync(0xffffffffu, lt[j]), q = __reduce_add_sync(0xffffffffu, eqb[j]);
    if (lane == j && e0 + j < n_slots) order[l + q] = (int32_t)((unsigned)(e0 + j) | (q ? GROUP_CONT : 0u));
  }
}

// gradient of the BPR terms w.r.t. one touched row = sum over the row's slots, in slot order, of
//   user slot:  coef * w * (n - p) + rc * u      pos slot:  -coef * w * u + rc * p      neg slot:  coef * w * u + rc * n
// (the per-sample values bpr_fused_kernel scatters in its atomic mode).  L2ROWS: the contribution is rc * e[row] per slot
// instead (LightGCN's layer-0 regulariser).  ACCUM: add the sum to what the row holds (a term on top of a propagated
// gradient), else overwrite (g_rep is zero everywhere else).
template <int G, int VPL, bool HAS_W, bool L2ROWS, bool ACCUM>
__global__ void __launch_bounds__(256) bpr_grad_rows_kernel(const float* rep, const int64_t* __restrict__ batch, int64_t item_offset,
                                                            const float* __restrict__ coef, const float* __restrict__ w,
                                                            float rc, const int32_t* __restrict__ order, int n_slots,
                                                            float* __restrict__ g_out, float* __restrict__ clear_table) {
  constexpr int D = G * VPL * 4;
  int k = (int)((blockIdx.x * (unsigned)blockDim.x + threadIdx.x) / G);
  const int gl = threadIdx.x & (G - 1);
  pdl_trigger();
  pdl_wait();
  if (k >= n_slots) return;
  unsigned o = (unsigned)order[k];
  if (o & GROUP_CONT) return;  // one lane group per distinct row: the one at the row's first position
  float4 acc[VPL], wv[VPL];
#pragma unroll
  for (int t = 0; t < VPL; ++t) {
    acc[t] = make_float4(0.f, 0.f, 0.f, 0.f);
    wv[t] = make_float4(1.f, 1.f, 1.f, 1.f);
    if constexpr (HAS_W) wv[t] = ldg_f4(w + (gl + t * G) * 4);
  }
  int64_t row = 0;
  for (bool more = true; more; more = ++k < n_slots && ((o = (unsigned)order[k]) & GROUP_CONT)) {
    const int slot = (int)(o & ~GROUP_CONT);
    const int smp = slot / 3, role = slot - 3 * smp;
    const int64_t ru = batch[3 * (size_t)smp];
    const int64_t rp = batch[3 * (size_t)smp + 1] + item_offset;
    const int64_t rn = batch[3 * (size_t)smp + 2] + item_offset;
    row = role == 0 ? ru : (role == 1 ? rp : rn);
    if (L2ROWS) {  // layer-0 L2: rc * e[row] per slot
#pragma unroll
      for (int t = 0; t < VPL; ++t) {
        const float4 ev = ldc_f4(rep + (size_t)row * D + (gl + t * G) * 4);
        acc[t].x += rc * ev.x; acc[t].y += rc * ev.y; acc[t].z += rc * ev.z; acc[t].w += rc * ev.w;
      }
      continue;
    }
    const float cf = coef[smp];
#pragma unroll
    for (int t = 0; t < VPL; ++t) {
      const int c = (gl + t * G) * 4;
      const float4 cw = make_float4(cf * wv[t].x, cf * wv[t].y, cf * wv[t].z, cf * wv[t].w);
      const float4 u = ldc_f4(rep + (size_t)ru * D + c);
      float4 g;
      if (role == 0) {
        const float4 p = ldc_f4(rep + (size_t)rp * D + c), n = ldc_f4(rep + (size_t)rn * D + c);
        g = make_float4(cw.x * (n.x - p.x) + rc * u.x, cw.y * (n.y - p.y) + rc * u.y, cw.z * (n.z - p.z) + rc * u.z,
                        cw.w * (n.w - p.w) + rc * u.w);
      } else if (role == 1) {
        const float4 p = (rc != 0.f) ? ldc_f4(rep + (size_t)rp * D + c) : make_float4(0.f, 0.f, 0.f, 0.f);
        g = make_float4(-cw.x * u.x + rc * p.x, -cw.y * u.y + rc * p.y, -cw.z * u.z + rc * p.z, -cw.w * u.w + rc * p.w);
      } else {
        const float4 n = (rc != 0.f) ? ldc_f4(rep + (size_t)rn * D + c) : make_float4(0.f, 0.f, 0.f, 0.f);
        g = make_float4(cw.x * u.x + rc * n.x, cw.y * u.y + rc * n.y, cw.z * u.z + rc * n.z, cw.w * u.w + rc * n.w);
      }
      acc[t].x += g.x; acc[t].y += g.y; acc[t].z += g.z; acc[t].w += g.w;
    }
  }
#pragma unroll
  for (int t = 0; t < VPL; ++t) {
    if (clear_table) st_f4(clear_table + (size_t)row * D + (gl + t * G) * 4, make_float4(0.f, 0.f, 0.f, 0.f));  // G back to all-zero
    float* dst = g_out + (size_t)row * D + (gl + t * G) * 4;
    if (ACCUM) {
      const float4 old = ld_f4(dst);
      st_f4(dst, make_float4(old.x + acc[t].x, old.y + acc[t].y, old.z + acc[t].z, old.w + acc[t].w));
    } else {
      st_f4(dst, acc[t]);
    }
  }
}

// zero the rows a batch touched (g_rep is kept all-zero between steps instead of being memset whole every step)
template <int G, int VPL>
__global__ void __launch_bounds__(256) clear_rows_kernel(const int64_t* __restrict__ batch, int n_slots, int64_t item_offset,
                                                         float* __restrict__ table) {
  constexpr int D = G * VPL * 4;
  const int slot = (int)((blockIdx.x * (unsigned)blockDim.x + threadIdx.x) / G), gl = threadIdx.x & (G - 1);
  pdl_trigger();
  pdl_wait();
  if (slot >= n_slots) return;
  const int64_t row = batch[slot] + ((slot % 3) ? item_offset : 0);
#pragma unroll
  for (int t = 0; t < VPL; ++t) st_f4(table + (size_t)row * D + (gl + t * G) * 4, make_float4(0.f, 0.f, 0.f, 0.f));
}

// LightGCN regulariser on layer-0 rows (model.py:114-117)
template <int G, int VPL>
__global__ void __launch_bounds__(256) bpr_l2_emb0_kernel(const float* __restrict__ emb0, const int64_t* batch,
                                                          int n_batch, int64_t item_offset, float l2_reg,
                                                          float* __restrict__ g_emb0, float* __restrict__ loss_out,
                                                          float* scratch) {  // g_emb0 == NULL: loss term only
  constexpr int D = G * VPL * 4;
  constexpr int SPB = 256 / G;
  __shared__ float s_loss[SPB];
  const int gib = threadIdx.x / G, gl = threadIdx.x & (G - 1);
  const int smp = blockIdx.x * SPB + gib;
  const bool active = smp < n_batch;
  const float inv_b = 1.f / (float)n_batch;
  const float rc = 2.f * l2_reg * inv_b;
  float l2 = 0.f;
  pdl_trigger();
  pdl_wait();  // g_emb0 is the previous kernel's output
  if (active) {
    int64_t rows[3] = {batch[3 * (size_t)smp], batch[3 * (size_t)smp + 1] + item_offset,
                       batch[3 * (size_t)smp + 2] + item_offset};
#pragma unroll
    for (int q = 0; q < 3; ++q) {
#pragma unroll
      for (int t = 0; t < VPL; ++t) {
        const int c = (gl + t * G) * 4;
        const float4 e = ldc_f4(emb0 + (size_t)rows[q] * D + c);
        l2 += e.x * e.x + e.y * e.y + e.z * e.z + e.w * e.w;
        if (g_emb0) red_add_f4(g_emb0 + (size_t)rows[q] * D + c, make_float4(rc * e.x, rc * e.y, rc * e.z, rc * e.w));
      }
    }
  }
  l2 = group_sum<G>(l2);
  if (gl == 0) s_loss[gib] = active ? l2 : 0.f;
  __syncthreads();
  float* part_loss = scratch + 4;
  if (threadIdx.x == 0) {
    float acc = 0.f;
    for (int i = 0; i < SPB; ++i) acc += s_loss[i];
    part_loss[blockIdx.x] = acc;
  }
  if (!last_block_ticket(scratch)) return;
  if (threadIdx.x == 0) {
    float acc = 0.f;
    for (unsigned i = 0; i < gridDim.x; ++i) acc += __ldcg(part_loss + i);
    *loss_out += l2_reg * acc * inv_b;
  }
}

// ---------------------------------------------------------------------------------------------- Adam
struct AdamPeers {
  int n;
  float* p[8];  // the same element range of the parameter table on the other ranks (row-partitioned multi-GPU)
};
__global__ void __launch_bounds__(256) adam_kernel(float* __restrict__ p, const float* g, float* __restrict__ m,
                                                   float* __restrict__ v, int64_t n, double lr, double beta1d, double beta2d,
                                                   float eps, const int64_t* __restrict__ step_ptr, const AdamPeers peers) {
  __shared__ float s_step_size, s_bc2_sqrt;
  if (threadIdx.x == 0) {
    const double t = (double)(*step_ptr + 1);
    const double bc1 = 1.0 - pow(beta1d, t), bc2 = 1.0 - pow(beta2d, t);
    s_step_size = (float)(lr / bc1);
    s_bc2_sqrt = (float)sqrt(bc2);
  }
  pdl_trigger();
  pdl_wait();  // the bias corrections above (two fp64 pow) overlap the tail of the kernel that produces g
  __syncthreads();
  const float step_size = s_step_size, bc2_sqrt = s_bc2_sqrt;
  const float w1 = (float)(1.0 - beta1d), w2 = (float)(1.0 - beta2d), beta2 = (float)beta2d;
  const int64_t n4 = n >> 2;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    const float4 gg = ld_f4(g + 4 * i);
    float4 mm = ld_f4(m + 4 * i), vv = ld_f4(v + 4 * i), pp = ld_f4(p + 4 * i);
#define B2_ADAM1(c)                                                            \
    mm.c = mm.c + w1 * (gg.c - mm.c);            /* exp_avg.lerp_(grad, 1-b1) */ \
    vv.c = vv.c * beta2 + (w2 * gg.c) * gg.c;    /* mul_(b2).addcmul_(g,g,1-b2) */ \
    pp.c = pp.c - step_size * (mm.c / (sqrtf(vv.c) / bc2_sqrt + eps));
    B2_ADAM1(x) B2_ADAM1(y) B2_ADAM1(z) B2_ADAM1(w)
    st_f4(m + 4 * i, mm); st_f4(v + 4 * i, vv); st_f4(p + 4 * i, pp);
    for (int q = 0; q < peers.n; ++q) st_f4(peers.p[q] + 4 * i, pp);  // updated rows go straight to the peers' tables
  }
  // tail (n not a multiple of 4)
  if (blockIdx.x == 0 && threadIdx.x < (n & 3)) {
    const int64_t i = (n4 << 2) + threadIdx.x;
    const float gg = g[i];
    float mm = m[i], vv = v[i];
    mm = mm + w1 * (gg - mm);
    vv = vv * beta2 + (w2 * gg) * gg;
    const float pn = p[i] - step_size * (mm / (sqrtf(vv) / bc2_sqrt + eps));
    p[i] = pn;
    for (int q = 0; q < peers.n; ++q) peers.p[q][i] = pn;
    m[i] = mm; v[i] = vv;
  }
#undef B2_ADAM1
}

__global__ void step_advance_kernel(int64_t* step, int64_t* step_b, const float* loss, double* loss_accum, int n_batch) {
  pdl_wait();
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    if (step) *step += 1;
    if (step_b) *step_b += 1;
    if (loss && loss_accum) {  // AverageMeter.update(loss.item(), B)  (utils.py:286-289)
      loss_accum[0] += (double)(*loss) * (double)n_batch;
      loss_accum[1] += (double)n_batch;
    }
  }
}

template <int G, int VPL>
static int launch_bpr(const float* rep, const int64_t* batch, int nb, int64_t off, float l2_reg, int reg_mode,
                      const float* w, float loss_scale, float* g_rep, float* g_w, float* loss_out, float* scratch,
                      float* dots, int dots_mode, float loss_weight, cudaStream_t st, float* coef = nullptr,
                      const int32_t* order = nullptr,
                      bool accumulate = false, const float* emb0 = nullptr, float l2_emb0 = 0.f) {
  const int grid = ceil_div(nb, 256 / G);
  if (w) B2_LAUNCH_PDL(bpr_fused_kernel<G, VPL, true>, grid, 256, 0, st, rep, batch, nb, off, l2_reg, reg_mode, w, loss_scale, g_rep, g_w, loss_out, scratch, dots, dots_mode, loss_weight, coef, emb0, l2_emb0);
  else B2_LAUNCH_PDL(bpr_fused_kernel<G, VPL, false>, grid, 256, 0, st, rep, batch, nb, off, l2_reg, reg_mode, w, loss_scale, g_rep, g_w, loss_out, scratch, dots, dots_mode, loss_weight, coef, emb0, l2_emb0);
  if (coef && dots_mode != 1) {  // ordered mode: one lane group per distinct row
    const float rc = (reg_mode == 1) ? 2.f * l2_reg * loss_scale / (float)nb : 0.f;
    const int grid2 = ceil_div(3 * nb, 256 / G);
#define B2_GRAD_ROWS(W, A) B2_LAUNCH_PDL(bpr_grad_rows_kernel<G, VPL, W, false, A>, grid2, 256, 0, st, rep, batch, off, (const float*)coef, w, rc, order, 3 * nb, g_rep, (float*)nullptr)
    if (w) { if (accumulate) B2_GRAD_ROWS(true, true); else B2_GRAD_ROWS(true, false); }
    else { if (accumulate) B2_GRAD_ROWS(false, true); else B2_GRAD_ROWS(false, false); }
#undef B2_GRAD_ROWS
  }
  return 0;
}

}  // namespace b200rec
using namespace b200rec;

#define B2_DISPATCH_D(d, ...)                                                                                     \
  switch (d) {                                                                                                    \
    case 8: { constexpr int G = 2, VPL = 1; __VA_ARGS__; } break;                                                        \
    case 16: { constexpr int G = 4, VPL = 1; __VA_ARGS__; } break;                                                       \
    case 32: { constexpr int G = 8, VPL = 1; __VA_ARGS__; } break;                                                       \
    case 64: { constexpr int G = 16, VPL = 1; __VA_ARGS__; } break;                                                      \
    case 128: { constexpr int G = 32, VPL = 1; __VA_ARGS__; } break;                                                     \
    case 256: { constexpr int G = 32, VPL = 2; __VA_ARGS__; } break;                                                     \
    default: return fail(B200REC_ERR_UNSUPPORTED, "%s: %s", __func__, "embedding size must be 8/16/32/64/128/256"); \
  }

extern "C" int b200rec_bpr_sample(const int32_t* user_ptr, const int32_t* user_items, int32_t n_users, int32_t n_items,
                                  uint64_t seed, const int64_t* step, int32_t batch, int64_t* out_batch, void* stream) {
  B2_REQUIRE(user_ptr && user_items && step && out_batch && n_users > 0 && n_items > 0 && batch > 0, "bad argument");
  bpr_sample_kernel<<<ceil_div(batch, 128), 128, 0, (cudaStream_t)stream>>>(user_ptr, user_items, n_users, n_items, seed,
                                                                           step, batch, out_batch);
  B2_LAUNCHED();
  return 0;
}

extern "C" int b200rec_mark_rows(const int64_t* batch, int32_t n_batch, int64_t item_offset, uint8_t* flags, void* stream) {
  B2_REQUIRE(batch && flags && n_batch > 0, "bad argument");
  mark_rows_kernel<<<ceil_div(3 * n_batch, 256), 256, 0, (cudaStream_t)stream>>>(batch, n_batch, item_offset, flags);
  B2_LAUNCHED();
  return 0;
}

extern "C" int b200rec_mark_reach(const int64_t* batch, int32_t n_batch, int64_t item_offset, const int32_t* user_ptr,
                                  const int32_t* user_items, uint8_t* flags, void* stream) {
  B2_REQUIRE(batch && flags && user_ptr && user_items && n_batch > 0, "bad argument");
  mark_reach_kernel<<<ceil_div(32 * n_batch, 256), 256, 0, (cudaStream_t)stream>>>(batch, n_batch, item_offset, user_ptr,
                                                                                    user_items, flags);
  B2_LAUNCHED();
  return 0;
}

extern "C" int b200rec_gather_rows(const float* table, int32_t d, const int64_t* idx, int64_t offset, int32_t n,
                                   float* out, float* sqnorm, void* stream) {
  B2_REQUIRE(table && idx && out && n >= 0, "bad argument");
  if (n == 0) return 0;
  B2_DISPATCH_D(d, (gather_rows_kernel<G, VPL><<<ceil_div(n, 256 / G), 256, 0, (cudaStream_t)stream>>>(table, idx, offset, n, out, sqnorm)));
  B2_LAUNCHED();
  return 0;
}

extern "C" int b200rec_scatter_add_rows(float* table, int32_t d, const int64_t* idx, int64_t offset, int32_t n,
                                        const float* src, void* stream) {
  B2_REQUIRE(table && idx && src && n >= 0, "bad argument");
  if (n == 0) return 0;
  B2_DISPATCH_D(d, (scatter_add_rows_kernel<G, VPL><<<ceil_div(n, 256 / G), 256, 0, (cudaStream_t)stream>>>(table, idx, offset, n, src)));
  B2_LAUNCHED();
  return 0;
}

extern "C" int64_t b200rec_bpr_scratch_floats(int32_t n_batch, int32_t d) {
  int g = d / 4;  // lanes per sample
  if (g < 2) g = 2;
  if (g > 32) g = 32;
  const int64_t samples_per_block = 256 / g;
  const int64_t blocks = (n_batch + samples_per_block - 1) / samples_per_block;
  return 4 + blocks * (int64_t)(d + 1);
}

extern "C" int b200rec_bpr_fwd_bwd(const float* rep, int32_t d, const int64_t* batch, int32_t n_batch, int64_t item_offset,
                                   float l2_reg, int32_t reg_mode, const float* w, float loss_scale, float* g_rep,
                                   float* g_w, float* loss_out, float* block_scratch, void* stream) {
  B2_REQUIRE(rep && batch && g_rep && loss_out && block_scratch && n_batch > 0, "bad argument");
  B2_REQUIRE(reg_mode == 0 || reg_mode == 1, "reg_mode must be 0 or 1 (layer-0 L2 is b200rec_bpr_l2_emb0)");
  B2_REQUIRE(!w || g_w, "g_w required with w");
  B2_DISPATCH_D(d, { int rc = launch_bpr<G, VPL>(rep, batch, n_batch, item_offset, l2_reg, reg_mode, w, loss_scale, g_rep, g_w, loss_out, block_scratch, nullptr, 0, 1.f, (cudaStream_t)stream); if (rc) return rc; });
  return 0;
}

extern "C" int b200rec_bpr_fwd_bwd_sharded(const float* rep, int32_t d, const int64_t* batch, int32_t n_batch,
                                           int64_t item_offset, float l2_reg, int32_t reg_mode, const float* w,
                                           float loss_scale, float* g_rep, float* g_w, float* loss_out,
                                           float* block_scratch, float* dots, int32_t phase, float loss_weight,
                                           void* stream) {
  B2_REQUIRE(rep && batch && dots && block_scratch && n_batch > 0, "bad argument");
  B2_REQUIRE(phase == 1 || phase == 2, "phase must be 1 (partial dots) or 2 (gradients from reduced dots)");
  B2_REQUIRE(phase == 1 || (g_rep && loss_out), "phase 2 needs g_rep and loss_out");
  B2_REQUIRE(reg_mode == 0 || reg_mode == 1, "reg_mode must be 0 or 1");
  B2_REQUIRE(!w || g_w || phase == 1, "g_w required with w");
  B2_DISPATCH_D(d, { int rc = launch_bpr<G, VPL>(rep, batch, n_batch, item_offset, l2_reg, reg_mode, w, loss_scale, g_rep, g_w, loss_out, block_scratch, dots, phase, loss_weight, (cudaStream_t)stream); if (rc) return rc; });
  return 0;
}

extern "C" int b200rec_bpr_group_rows(const int64_t* batch, int32_t n_batch, int64_t item_offset, int32_t* order,
                                      void* stream) {
  B2_REQUIRE(batch && order && n_batch > 0, "bad argument");
  if (3 * (int64_t)n_batch > GROUP_CAP)
    return fail(B200REC_ERR_UNSUPPORTED, "%s: %s", __func__, "batch too large for the grouping (3 * n_batch <= 12288)");
  const int n_slots = 3 * n_batch;
  const size_t smem = (size_t)((n_slots + 31) & ~31) * sizeof(int32_t);  // <= 48 KB: no opt-in needed
  bpr_group_rows_kernel<<<ceil_div(n_slots, 8 * GROUP_EPW), 256, smem, (cudaStream_t)stream>>>(batch, n_slots, item_offset, order);
  B2_LAUNCHED();
  return 0;
}

extern "C" int b200rec_bpr_fwd_bwd_ordered(const float* rep, int32_t d, const int64_t* batch, int32_t n_batch,
                                           int64_t item_offset, float l2_reg, int32_t reg_mode, const float* w,
                                           float loss_scale, float* g_rep, float* g_w, float* loss_out,
                                           float* block_scratch, float* dots, int32_t phase, float loss_weight,
                                           float* coef, const int32_t* order, int32_t accumulate, const float* emb0, float l2_emb0,
                                           void* stream) {
  B2_REQUIRE(rep && batch && block_scratch && n_batch > 0, "bad argument");
  B2_REQUIRE(phase >= 0 && phase <= 2, "phase must be 0 (one GPU), 1 (partial dots) or 2 (gradients from reduced dots)");
  B2_REQUIRE(phase == 0 || dots, "dots required when the dimension is sharded");
  B2_REQUIRE(phase == 1 || (g_rep && loss_out && coef && order), "null output / grouping");
  B2_REQUIRE(reg_mode == 0 || reg_mode == 1, "reg_mode must be 0 or 1");
  B2_REQUIRE(!w || g_w || phase == 1, "g_w required with w");
  B2_DISPATCH_D(d, { int rc = launch_bpr<G, VPL>(rep, batch, n_batch, item_offset, l2_reg, reg_mode, w, loss_scale, g_rep, g_w, loss_out, block_scratch, dots, phase, loss_weight, (cudaStream_t)stream, phase == 1 ? nullptr : coef, order, accumulate != 0, phase == 1 ? nullptr : emb0, l2_emb0); if (rc) return rc; });
  return 0;
}

extern "C" int b200rec_bpr_l2_emb0_ordered(const float* emb0, int32_t d, const int64_t* batch, int32_t n_batch,
                                           int64_t item_offset, float l2_reg, float* g_emb0, float* loss_out,
                                           float* block_scratch, const int32_t* order, float* clear_table,
                                           void* stream) {
  B2_REQUIRE(emb0 && batch && g_emb0 && order && n_batch > 0, "bad argument");
  B2_REQUIRE(!loss_out || block_scratch, "block_scratch required with loss_out");
  const float rc = 2.f * l2_reg / (float)n_batch;
  B2_DISPATCH_D(d, {
    if (loss_out)  // NULL: the loss term was folded into b200rec_bpr_fwd_bwd_ordered (emb0 / l2_emb0)
      B2_LAUNCH_PDL(bpr_l2_emb0_kernel<G, VPL>, ceil_div(n_batch, 256 / G), 256, 0, (cudaStream_t)stream, emb0, batch, n_batch, item_offset, l2_reg, (float*)nullptr, loss_out, block_scratch);
    B2_LAUNCH_PDL(bpr_grad_rows_kernel<G, VPL, false, true, true>, ceil_div(3 * n_batch, 256 / G), 256, 0, (cudaStream_t)stream, emb0, batch, item_offset, (const float*)nullptr, (const float*)nullptr, rc, order, 3 * n_batch, g_emb0, clear_table);
  });
  return 0;
}

extern "C" int b200rec_clear_rows(const int64_t* batch, int32_t n_batch, int64_t item_offset, float* table, int32_t d,
                                  void* stream) {
  B2_REQUIRE(batch && table && n_batch > 0, "bad argument");
  B2_DISPATCH_D(d, { B2_LAUNCH_PDL(clear_rows_kernel<G, VPL>, ceil_div(3 * n_batch, 256 / G), 256, 0, (cudaStream_t)stream, batch, 3 * n_batch, item_offset, table); });
  return 0;
}

extern "C" int b200rec_bpr_l2_emb0(const float* emb0, int32_t d, const int64_t* batch, int32_t n_batch, int64_t item_offset,
                                   float l2_reg, float* g_emb0, float* loss_out, float* block_scratch, void* stream) {
  B2_REQUIRE(emb0 && batch && g_emb0 && loss_out && block_scratch && n_batch > 0, "bad argument");
  B2_DISPATCH_D(d, { B2_LAUNCH_PDL(bpr_l2_emb0_kernel<G, VPL>, ceil_div(n_batch, 256 / G), 256, 0, (cudaStream_t)stream, emb0, batch, n_batch, item_offset, l2_reg, g_emb0, loss_out, block_scratch); });
  return 0;
}

extern "C" int b200rec_adam_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n, double lr,
                                 double beta1, double beta2, double eps, const int64_t* step, void* stream) {
  B2_REQUIRE(param && grad && exp_avg && exp_avg_sq && step && n > 0, "bad argument");
  B2_REQUIRE((((uintptr_t)param | (uintptr_t)grad | (uintptr_t)exp_avg | (uintptr_t)exp_avg_sq) & 15) == 0, "16-byte alignment");
  int grid = ceil_div(n / 4 + 1, 256);
  if (grid > 148 * 16) grid = 148 * 16;
  AdamPeers none;
  none.n = 0;
  B2_LAUNCH_PDL(adam_kernel, grid, 256, 0, (cudaStream_t)stream, param, grad, exp_avg, exp_avg_sq, n, lr, beta1, beta2, (float)eps, step, none);
  return 0;
}

extern "C" int b200rec_adam_step_peer(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n, double lr,
                                      double beta1, double beta2, double eps, const int64_t* step, int32_t n_peers,
                                      float* const* peer_param /*HOST [n_peers]*/, void* stream) {
  B2_REQUIRE(param && grad && exp_avg && exp_avg_sq && step && n > 0, "bad argument");
  B2_REQUIRE(n_peers >= 0 && n_peers <= 8 && (n_peers == 0 || peer_param), "at most 8 peers");
  B2_REQUIRE((((uintptr_t)param | (uintptr_t)grad | (uintptr_t)exp_avg | (uintptr_t)exp_avg_sq) & 15) == 0, "16-byte alignment");
  AdamPeers pr;
  pr.n = n_peers;
  for (int q = 0; q < 8; ++q) pr.p[q] = q < n_peers ? peer_param[q] : nullptr;
  int grid = ceil_div(n / 4 + 1, 256);
  if (grid > 148 * 16) grid = 148 * 16;
  B2_LAUNCH_PDL(adam_kernel, grid, 256, 0, (cudaStream_t)stream, param, grad, exp_avg, exp_avg_sq, n, lr, beta1, beta2, (float)eps, step, pr);
  return 0;
}

extern "C" int b200rec_dropout_mask(int32_t nnz, float p, uint64_t seed, const int64_t* step, uint32_t* keep_bits,
                                    void* stream) {
  B2_REQUIRE(nnz > 0 && step && keep_bits && p >= 0.f && p < 1.f, "bad argument");
  const int n_words = (nnz + 31) >> 5;
  dropout_mask_kernel<<<ceil_div(n_words, 256), 256, 0, (cudaStream_t)stream>>>(nnz, (float)(1.0 - (double)p), seed, step,
                                                                               keep_bits);
  B2_LAUNCHED();
  return 0;
}

extern "C" int b200rec_step_advance(int64_t* step, int64_t* step_b, const float* loss, double* loss_accum, int32_t n_batch,
                                    void* stream) {
  B2_LAUNCH_PDL(step_advance_kernel, 1, 32, 0, (cudaStream_t)stream, step, step_b, loss, loss_accum, n_batch);
  return 0;
}
