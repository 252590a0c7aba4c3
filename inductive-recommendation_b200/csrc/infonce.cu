// InfoNCE between two views' user rows, forward + backward in one call (SGL.cal_loss model.py:206-214, HALF :322-332;
// the `info_nce` package's 'unpaired' form with negative_keys == positive_key, temperature 0.1):
//   qn = q / max(|q|, 1e-12), kn likewise;  logits_i = [qn_i.kn_i, qn_i.kn_0, ..., qn_i.kn_{n-1}] / T;  loss = mean_i CE(logits_i, 0)
// The n x n logit matrix (2048 x 2048 per step) is never stored: a statistics pass keeps a running (max, sum) per row,
// a gradient pass (both views in one launch) recomputes the tiles and contracts them with the other view on the fly,
// then apply the normalisation's Jacobian row-locally.  All sums run in a fixed order (reproducible run to run; only the
// scatter-add of rows that occur twice in `rows` depends on arrival order, like every other gradient scatter here).
#include "common.cuh"

namespace b200rec {

constexpr int NCE_TM = 32;  // rows of the block's own view per CTA
constexpr float NCE_EPS = 1e-12f;

// one warp per sample: gather (optional), normalise both views, positive logit
template <int D>
__global__ void __launch_bounds__(256) infonce_prep_kernel(const float* q, const float* k, const int64_t* rows, int row_stride,
                                                           int n, float inv_t, float* qn, float* kn, float* qden, float* kden,
                                                           float* pos) {
  pdl_trigger();
  pdl_wait();
  const int i = (int)((blockIdx.x * (unsigned)blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
  if (i >= n) return;
  const int64_t r = rows ? rows[(size_t)i * row_stride] : i;
  const float* qr = q + (size_t)r * D;
  const float* kr = k + (size_t)r * D;
  constexpr int PER = (D + 31) / 32;
  float a[PER], b[PER];
  float sq = 0.f, sk = 0.f;
#pragma unroll
  for (int t = 0; t < PER; ++t) {
    const int c = lane + 32 * t;
    a[t] = c < D ? qr[c] : 0.f;
    b[t] = c < D ? kr[c] : 0.f;
    sq = fmaf(a[t], a[t], sq);
    sk = fmaf(b[t], b[t], sk);
  }
  sq = warp_sum(sq);
  sk = warp_sum(sk);
  const float dq = fmaxf(sqrtf(sq), NCE_EPS), dk = fmaxf(sqrtf(sk), NCE_EPS);
  float dot = 0.f;
#pragma unroll
  for (int t = 0; t < PER; ++t) {
    const int c = lane + 32 * t;
    a[t] = a[t] / dq;
    b[t] = b[t] / dk;
    dot = fmaf(a[t], b[t], dot);
    if (c < D) {
      qn[(size_t)i * D + c] = a[t];
      kn[(size_t)i * D + c] = b[t];
    }
  }
  dot = warp_sum(dot);
  if (lane == 0) {
    qden[i] = dq;
    kden[i] = dk;
    pos[i] = dot * inv_t;
  }
}

// One CTA (128 threads) = 32 rows of its OWN view against tiles of 64 rows of the OTHER view.  Both contractions are
// register-tiled 4 x 4 per thread with 128-bit shared-memory loads (shared memory feeds 32 floats per clock per SM
// against 128 FMAs, so every loaded value has to be used at least four times):
//   logits  S[i][j]  = sum_c own[i][c] * other[j][c]        thread (ti, tj): rows ti+8r, columns tj+16q
//   output  O[i][d] += sum_j P[i][j] * other[j][d]          thread (ti, td): rows ti+8r, columns td*CW .. td*CW+CW-1
// The other view's rows are split into column parts over more CTAs (grid.y / grid.z): 64 row blocks alone would leave
// most of the 148 SMs with one 4-warp CTA or none.
//  MODE 0 (grid.y = column parts): running (max, sum) of exp(logit) per row over this part's columns.
//  MODE 1 (grid.y = 0: gradient of the query view, own = qn, softmax offset = lse[own row];
//          grid.y = 1: gradient of the key view,   own = kn, softmax offset = lse[other row]; grid.z = column parts):
//          P tile -> shared memory -> this part's share of O -> workspace; infonce_finish_kernel completes the row.
constexpr int NCE_TN2 = 64;
constexpr int NCE_THREADS = 128;

template <int D>
struct NceSmem {
  static constexpr int XS = D + 4;              // row stride of both operand tiles: 16-byte aligned, rows 4 banks apart
  static constexpr int PS = NCE_TN2 + 4;
  static constexpr int A_FLOATS = NCE_TM * XS;
  static constexpr int B_FLOATS = NCE_TN2 * XS;
  static constexpr int P_FLOATS = NCE_TM * PS;
  static constexpr size_t BYTES = (size_t)(A_FLOATS + B_FLOATS + P_FLOATS + NCE_TN2) * sizeof(float);
};

template <int D, int MODE>
__global__ void __launch_bounds__(NCE_THREADS) infonce_pass_kernel(const float* qn, const float* kn, const float* qden,
                                                                   const float* kden, const float* pos, const float* lse,
                                                                   float* part_m, float* part_l, int n, float inv_t,
                                                                   float* opart) {
  using L = NceSmem<D>;
  extern __shared__ __align__(16) float smem[];
  float* As = smem;                    // [TM][XS]
  float* Bs = As + L::A_FLOATS;        // [TN][XS]
  float* Ps = Bs + L::B_FLOATS;        // [TM][PS]
  float* s_off = Ps + L::P_FLOATS;     // [TN]
  const int tid = threadIdx.x, ti = tid >> 4, tj = tid & 15;
  const int row0 = blockIdx.x * NCE_TM;
  const bool key_side = (MODE == 1) && blockIdx.y == 1;
  const int part = (MODE == 0) ? blockIdx.y : blockIdx.z, n_part = (MODE == 0) ? gridDim.y : gridDim.z;
  const float* own = key_side ? kn : qn;
  const float* other = key_side ? qn : kn;
  pdl_trigger();
  pdl_wait();
  for (int e = tid * 4; e < NCE_TM * D; e += NCE_THREADS * 4) {
    const int i = e / D, c = e - i * D;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (row0 + i < n) v = *reinterpret_cast<const float4*>(own + (size_t)row0 * D + e);
    *reinterpret_cast<float4*>(As + i * L::XS + c) = v;
  }
  // the other view's rows are split over n_part CTAs on tile boundaries (a 2048-sample batch has only 64 row blocks)
  const int tiles = (n + NCE_TN2 - 1) / NCE_TN2, per = (tiles + n_part - 1) / n_part;
  const int j_begin = min(n, part * per * NCE_TN2), j_end = min(n, (part + 1) * per * NCE_TN2);
  float m_run[4], l_run[4], own_off[4];
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    m_run[r] = -INFINITY;
    l_run[r] = 0.f;
    const int i = row0 + ti + 8 * r;
    own_off[r] = (MODE == 1 && !key_side && i < n) ? lse[i] : 0.f;
  }
  constexpr int CW = (D >= 16) ? D / 16 : 1;  // output columns per thread (16 threads per row; D = 8: half of them idle)
  const int oc0 = tj * CW;
  const bool oc_live = oc0 < D;
  float acc[4][CW];
#pragma unroll
  for (int r = 0; r < 4; ++r)
#pragma unroll
    for (int c = 0; c < CW; ++c) acc[r][c] = 0.f;
  __syncthreads();

  for (int j0 = j_begin; j0 < j_end; j0 += NCE_TN2) {
    for (int e = tid * 4; e < NCE_TN2 * D; e += NCE_THREADS * 4) {
      const int j = e / D, c = e - j * D;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (j0 + j < n) v = *reinterpret_cast<const float4*>(other + (size_t)j0 * D + e);
      *reinterpret_cast<float4*>(Bs + j * L::XS + c) = v;
    }
    if (key_side && tid < NCE_TN2) s_off[tid] = (j0 + tid < n) ? lse[j0 + tid] : 0.f;
    __syncthreads();
    float s[4][4];
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
      for (int q = 0; q < 4; ++q) s[r][q] = 0.f;
#pragma unroll 2
    for (int c = 0; c < D; c += 4) {
      float4 a[4], b[4];
#pragma unroll
      for (int r = 0; r < 4; ++r) a[r] = *reinterpret_cast<const float4*>(As + (ti + 8 * r) * L::XS + c);
#pragma unroll
      for (int q = 0; q < 4; ++q) b[q] = *reinterpret_cast<const float4*>(Bs + (tj + 16 * q) * L::XS + c);
#pragma unroll
      for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          s[r][q] = fmaf(a[r].x, b[q].x, s[r][q]);
          s[r][q] = fmaf(a[r].y, b[q].y, s[r][q]);
          s[r][q] = fmaf(a[r].z, b[q].z, s[r][q]);
          s[r][q] = fmaf(a[r].w, b[q].w, s[r][q]);
        }
    }
    if (MODE == 0) {
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        float tm = -INFINITY;
#pragma unroll
        for (int q = 0; q < 4; ++q)
          if (j0 + tj + 16 * q < j_end) tm = fmaxf(tm, s[r][q] * inv_t);
        if (tm > -INFINITY) {
          const float m_new = fmaxf(m_run[r], tm);
          float ex = 0.f;
#pragma unroll
          for (int q = 0; q < 4; ++q)
            if (j0 + tj + 16 * q < j_end) ex += expf(s[r][q] * inv_t - m_new);
          l_run[r] = l_run[r] * expf(m_run[r] - m_new) + ex;  // exp(-inf) = 0 on the first tile
          m_run[r] = m_new;
        }
      }
    } else {
#pragma unroll
      for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int j = tj + 16 * q;
          const float off = key_side ? s_off[j] : own_off[r];
          Ps[(ti + 8 * r) * L::PS + j] = (j0 + j < n) ? expf(s[r][q] * inv_t - off) : 0.f;
        }
      __syncthreads();
      if (oc_live) {
#pragma unroll 2
        for (int j = 0; j < NCE_TN2; j += 4) {
          float4 pv[4];
#pragma unroll
          for (int r = 0; r < 4; ++r) pv[r] = *reinterpret_cast<const float4*>(Ps + (ti + 8 * r) * L::PS + j);
#pragma unroll
          for (int jj = 0; jj < 4; ++jj) {
            float bv[CW];
            const float* bp = Bs + (j + jj) * L::XS + oc0;
            if (CW % 4 == 0) {
#pragma unroll
              for (int c = 0; c < CW; c += 4) {
                const float4 t = *reinterpret_cast<const float4*>(bp + c);
                bv[c] = t.x; bv[c + 1] = t.y; bv[c + 2] = t.z; bv[c + 3] = t.w;
              }
            } else if (CW == 2) {
              const float2 t = *reinterpret_cast<const float2*>(bp);
              bv[0] = t.x; bv[CW - 1] = t.y;
            } else {
              bv[0] = bp[0];
            }
#pragma unroll
            for (int r = 0; r < 4; ++r) {
              const float p = jj == 0 ? pv[r].x : jj == 1 ? pv[r].y : jj == 2 ? pv[r].z : pv[r].w;
#pragma unroll
              for (int c = 0; c < CW; ++c) acc[r][c] = fmaf(p, bv[c], acc[r][c]);
            }
          }
        }
      }
    }
    __syncthreads();
  }

  if (MODE == 0) {
    // merge the (max, sum) pairs of the 16 threads that share a row: xor butterfly, the same order every run
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      float m = m_run[r], l = l_run[r];
#pragma unroll
      for (int o = 1; o < 16; o <<= 1) {
        const float mo = __shfl_xor_sync(0xffffffffu, m, o), lo = __shfl_xor_sync(0xffffffffu, l, o);
        const float m_new = fmaxf(m, mo);
        if (m_new > -INFINITY) l = l * expf(m - m_new) + lo * expf(mo - m_new);
        m = m_new;
      }
      const int i = row0 + ti + 8 * r;
      if (tj == 0 && i < n) {
        part_m[(size_t)blockIdx.y * n + i] = m;
        part_l[(size_t)blockIdx.y * n + i] = l;
      }
    }
    return;
  }

  // this part's share of O; infonce_finish_kernel adds the parts in order
  float* obase = opart + ((size_t)(key_side ? n_part : 0) + part) * n * D;
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const int i = row0 + ti + 8 * r;
    if (i < n && oc_live) {
#pragma unroll
      for (int c = 0; c < CW; ++c) obase[(size_t)i * D + oc0 + c] = acc[r][c];
    }
  }
}

// One warp per (view, sample): O = sum of the column parts (fixed order) + the positive-pair term, scaled; then the
// Jacobian of x -> x / max(|x|, eps) and the store / scatter-add of the gradient row.
template <int D>
__global__ void __launch_bounds__(256) infonce_finish_kernel(const float* qn, const float* kn, const float* qden, const float* kden,
                                                             const float* pos, const float* lse, const float* opart, int n_part,
                                                             int n, float coef, const int64_t* rows, int row_stride, float* gq,
                                                             float* gk) {
  pdl_trigger();
  pdl_wait();
  const int wid = (int)((blockIdx.x * (unsigned)blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
  if (wid >= 2 * n) return;
  const bool key_side = wid >= n;
  const int i = key_side ? wid - n : wid;
  const float* own = (key_side ? kn : qn) + (size_t)i * D;
  const float* other = (key_side ? qn : kn) + (size_t)i * D;
  const float* ob = opart + (size_t)(key_side ? n_part : 0) * n * D + (size_t)i * D;
  const float ppos = expf(pos[i] - lse[i]) - 1.f;  // d loss_i / d pos_i (the positive sits at label 0)
  const float den = key_side ? kden[i] : qden[i];
  constexpr int PER = (D + 31) / 32;
  float o[PER], x[PER];
  float dotp = 0.f;
#pragma unroll
  for (int t = 0; t < PER; ++t) {
    const int c = lane + 32 * t;
    o[t] = 0.f;
    x[t] = 0.f;
    if (c < D) {
      for (int h = 0; h < n_part; ++h) o[t] += ob[(size_t)h * n * D + c];
      o[t] = (o[t] + ppos * other[c]) * coef;
      x[t] = own[c];
      dotp = fmaf(x[t], o[t], dotp);
    }
  }
  dotp = warp_sum(dotp);
  const bool clamped = den <= NCE_EPS;  // F.normalize's clamp: the quotient is then linear in x
  const int64_t r_out = rows ? rows[(size_t)i * row_stride] : i;
  float* g_out = (key_side ? gk : gq) + (size_t)r_out * D;
#pragma unroll
  for (int t = 0; t < PER; ++t) {
    const int c = lane + 32 * t;
    if (c < D) {
      const float gx = clamped ? o[t] / den : (o[t] - x[t] * dotp) / den;
      if (rows) atomicAdd(g_out + c, gx);
      else g_out[c] = gx;
    }
  }
}

// lse_i over [pos_i, row i's logits] from the column halves' partials; loss_out[0] += scale * sum_i (lse_i - pos_i),
// summed in a fixed order (thread-strided partials, then a shared-memory tree)
__global__ void __launch_bounds__(256) infonce_lse_kernel(const float* part_m, const float* part_l, int n_half, const float* pos,
                                                          int n, float scale, float* lse, float* loss_out) {
  __shared__ float s_sum[256];
  pdl_trigger();
  pdl_wait();
  float t = 0.f;
  for (int i = threadIdx.x; i < n; i += 256) {
    float m = pos[i], l = 1.f;
    for (int h = 0; h < n_half; ++h) {
      const float mh = part_m[(size_t)h * n + i], lh = part_l[(size_t)h * n + i];
      if (mh > -INFINITY) {
        const float m_new = fmaxf(m, mh);
        l = l * expf(m - m_new) + lh * expf(mh - m_new);
        m = m_new;
      }
    }
    const float v = m + logf(l);
    lse[i] = v;
    t += v - pos[i];
  }
  s_sum[threadIdx.x] = t;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) s_sum[threadIdx.x] += s_sum[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) loss_out[0] += s_sum[0] * scale;
}

constexpr int NCE_MAX_PARTS = 8;

template <int D>
static int infonce_launch(const float* q, const float* k, const int64_t* rows, int row_stride, int n, float temperature,
                          float loss_scale, float* loss_out, float* gq, float* gk, float* ws, cudaStream_t st) {
  float* qn = ws;
  float* kn = qn + (size_t)n * D;
  float* qden = kn + (size_t)n * D;
  float* kden = qden + n;
  float* pos = kden + n;
  float* lse = pos + n;
  float* part_m = lse + n;
  float* part_l = part_m + (size_t)NCE_MAX_PARTS * n;
  float* opart = part_l + (size_t)NCE_MAX_PARTS * n;  // [2][parts][n][D]
  const int nblk = ceil_div(n, NCE_TM), tiles = ceil_div(n, NCE_TN2);
  // enough CTAs for ~3 per SM: the row blocks alone are too few (64 for a 2048-sample batch)
  const int parts_stat = max(1, min(min(NCE_MAX_PARTS, tiles), ceil_div(444, nblk)));
  const int parts_grad = max(1, min(min(NCE_MAX_PARTS, tiles), ceil_div(444, 2 * nblk)));
  const float inv_t = 1.f / temperature;
  const size_t smem = NceSmem<D>::BYTES;
  // opt in to > 48 KB of dynamic shared memory: the attribute is per DEVICE, so it is set on every call (cheap) rather
  // than once per process -- a process that later runs on another GPU would otherwise fail to launch
  B2_CUDA(cudaFuncSetAttribute(infonce_pass_kernel<D, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  B2_CUDA(cudaFuncSetAttribute(infonce_pass_kernel<D, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  B2_LAUNCH_PDL(infonce_prep_kernel<D>, ceil_div((long long)n * 32, 256), 256, 0, st, q, k, rows, row_stride, n, inv_t, qn, kn,
                qden, kden, pos);
  B2_LAUNCH_PDL((infonce_pass_kernel<D, 0>), dim3(nblk, parts_stat), NCE_THREADS, smem, st, (const float*)qn, (const float*)kn,
                (const float*)qden, (const float*)kden, (const float*)pos, (const float*)lse, part_m, part_l, n, inv_t,
                (float*)nullptr);
  B2_LAUNCH_PDL(infonce_lse_kernel, 1, 256, 0, st, (const float*)part_m, (const float*)part_l, parts_stat, (const float*)pos, n,
                loss_scale / (float)n, lse, loss_out);
  B2_LAUNCH_PDL((infonce_pass_kernel<D, 1>), dim3(nblk, 2, parts_grad), NCE_THREADS, smem, st, (const float*)qn,
                (const float*)kn, (const float*)qden, (const float*)kden, (const float*)pos, (const float*)lse, part_m, part_l, n,
                inv_t, opart);
  const float coef = loss_scale * inv_t / (float)n;
  B2_LAUNCH_PDL(infonce_finish_kernel<D>, ceil_div((long long)2 * n * 32, 256), 256, 0, st, (const float*)qn, (const float*)kn,
                (const float*)qden, (const float*)kden, (const float*)pos, (const float*)lse, (const float*)opart, parts_grad, n,
                coef, rows, row_stride, gq, gk);
  return 0;
}

}  // namespace b200rec

using namespace b200rec;

extern "C" int64_t b200rec_infonce_workspace_floats(int32_t n, int32_t d) {
  if (n <= 0 || d <= 0) return 0;
  return 2ll * n * d + 4ll * n + 2ll * NCE_MAX_PARTS * n + 2ll * NCE_MAX_PARTS * n * d + 64;
}

extern "C" int b200rec_infonce_fwd_bwd(const float* q, const float* k, const int64_t* rows, int32_t row_stride, int32_t n,
                                       int32_t d, float temperature, float loss_scale, float* loss_out, float* gq, float* gk,
                                       float* workspace, void* stream) {
  B2_REQUIRE(q && k && loss_out && gq && gk && workspace && n > 0, "null argument");
  B2_REQUIRE(temperature > 0.f, "temperature must be positive");
  B2_REQUIRE(!rows || row_stride >= 1, "row_stride must be >= 1");
  cudaStream_t st = (cudaStream_t)stream;
  switch (d) {
    case 8: return infonce_launch<8>(q, k, rows, row_stride, n, temperature, loss_scale, loss_out, gq, gk, workspace, st);
    case 16: return infonce_launch<16>(q, k, rows, row_stride, n, temperature, loss_scale, loss_out, gq, gk, workspace, st);
    case 32: return infonce_launch<32>(q, k, rows, row_stride, n, temperature, loss_scale, loss_out, gq, gk, workspace, st);
    case 64: return infonce_launch<64>(q, k, rows, row_stride, n, temperature, loss_scale, loss_out, gq, gk, workspace, st);
    case 128: return infonce_launch<128>(q, k, rows, row_stride, n, temperature, loss_scale, loss_out, gq, gk, workspace, st);
    case 256: return infonce_launch<256>(q, k, rows, row_stride, n, temperature, loss_scale, loss_out, gq, gk, workspace, st);
    default: return fail(B200REC_ERR_UNSUPPORTED, "%s: %s", __func__, "embedding size must be 8/16/32/64/128/256");
  }
}
