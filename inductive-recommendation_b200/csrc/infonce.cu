// InfoNCE between two views' user rows, forward + backward in one call (SGL.cal_loss model.py:206-214, HALF :322-332;
// the `info_nce` package's 'unpaired' form with negative_keys == positive_key, temperature 0.1):
//   qn = q / max(|q|, 1e-12), kn likewise;  logits_i = [qn_i.kn_i, qn_i.kn_0, ..., qn_i.kn_{n-1}] / T;  loss = mean_i CE(logits_i, 0)
// The n x n logit matrix (2048 x 2048 per step) is never stored: a statistics pass keeps a running (max, sum) per row,
// two gradient passes recompute the tiles and contract them with the other view on the fly (flash-attention style),
// then apply the normalisation's Jacobian row-locally.  All sums run in a fixed order (reproducible run to run; only the
// scatter-add of rows that occur twice in `rows` depends on arrival order, like every other gradient scatter here).
#include "common.cuh"

namespace b200rec {

constexpr int NCE_TM = 32;  // rows of the block's own view per CTA
constexpr int NCE_TN = 32;  // rows of the other view per tile
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

// MODE 0: per-row log-sum-exp over [pos_i, S_i0 .. S_i,n-1] and the row's loss term.
// MODE 1: gradient w.r.t. the query view  (own = qn, other = kn, softmax rows indexed by the OWN row).
// MODE 2: gradient w.r.t. the key view    (own = kn, other = qn, softmax rows indexed by the OTHER row).
template <int D, int MODE>
__global__ void __launch_bounds__(256) infonce_pass_kernel(const float* own, const float* other, const float* own_den,
                                                           const float* pos, float* lse, float* loss_part, int n, float inv_t,
                                                           float coef, const int64_t* rows, int row_stride, float* g_out) {
  extern __shared__ __align__(16) float smem[];
  constexpr int AS = D;          // own tile: broadcast reads
  constexpr int BS = D + 1;      // other tile: lane <-> row, padded against bank conflicts
  float* As = smem;                       // [TM][AS]
  float* Bs = As + NCE_TM * AS;           // [TN][BS]
  float* Ps = Bs + NCE_TN * BS;           // [TM][TN+1]
  float* s_lse = Ps + NCE_TM * (NCE_TN + 1);  // [TN] softmax offsets of the other rows (MODE 2)
  const int tid = threadIdx.x, tx = tid & 31, ty = tid >> 5;
  const int row0 = blockIdx.x * NCE_TM;
  pdl_trigger();
  pdl_wait();
  for (int e = tid; e < NCE_TM * D; e += 256) {
    const int i = e / D, c = e - i * D;
    As[i * AS + c] = (row0 + i < n) ? own[(size_t)(row0 + i) * D + c] : 0.f;
  }
  // MODE 0: running (max, sum) of the 4 rows this thread shares with its warp, seeded with the positive logit
  float m_run[4], l_run[4], own_lse[4];
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const int i = row0 + ty * 4 + r;
    m_run[r] = (MODE == 0 && i < n) ? pos[i] : 0.f;
    l_run[r] = 1.f;
    own_lse[r] = (MODE == 1 && i < n) ? lse[i] : 0.f;
  }
  // MODE 1/2: output accumulators, thread = (row tid/8, columns (tid%8) + 8*c)
  constexpr int CPT = (D + 7) / 8;
  const int orow = tid >> 3, oc0 = tid & 7;
  float acc[CPT];
#pragma unroll
  for (int c = 0; c < CPT; ++c) acc[c] = 0.f;
  __syncthreads();

  for (int j0 = 0; j0 < n; j0 += NCE_TN) {
    for (int e = tid; e < NCE_TN * D; e += 256) {
      const int j = e / D, c = e - j * D;
      Bs[j * BS + c] = (j0 + j < n) ? other[(size_t)(j0 + j) * D + c] : 0.f;
    }
    if (MODE == 2 && tid < NCE_TN) s_lse[tid] = (j0 + tid < n) ? lse[j0 + tid] : 0.f;
    __syncthreads();
    float s[4] = {0.f, 0.f, 0.f, 0.f};
    const float* brow = Bs + tx * BS;
#pragma unroll 2
    for (int c = 0; c < D; c += 4) {
      const float b0 = brow[c], b1 = brow[c + 1], b2 = brow[c + 2], b3 = brow[c + 3];
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const float4 a = *reinterpret_cast<const float4*>(As + (ty * 4 + r) * AS + c);  // broadcast LDS.128
        s[r] = fmaf(a.x, b0, s[r]);
        s[r] = fmaf(a.y, b1, s[r]);
        s[r] = fmaf(a.z, b2, s[r]);
        s[r] = fmaf(a.w, b3, s[r]);
      }
    }
    const bool jv = j0 + tx < n;
    if (MODE == 0) {
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const float z = jv ? s[r] * inv_t : -INFINITY;
        float tm = z;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) tm = fmaxf(tm, __shfl_xor_sync(0xffffffffu, tm, o));
        const float m_new = fmaxf(m_run[r], tm);
        float ex = jv ? expf(z - m_new) : 0.f;
        ex = warp_sum(ex);
        l_run[r] = l_run[r] * expf(m_run[r] - m_new) + ex;
        m_run[r] = m_new;
      }
    } else {
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const float off = (MODE == 1) ? own_lse[r] : s_lse[tx];
        Ps[(ty * 4 + r) * (NCE_TN + 1) + tx] = jv ? expf(s[r] * inv_t - off) : 0.f;
      }
      __syncthreads();
#pragma unroll 4
      for (int j = 0; j < NCE_TN; ++j) {
        const float pv = Ps[orow * (NCE_TN + 1) + j];
#pragma unroll
        for (int c = 0; c < CPT; ++c) {
          const int col = oc0 + 8 * c;
          if (col < D) acc[c] = fmaf(pv, Bs[j * BS + col], acc[c]);
        }
      }
    }
    __syncthreads();
  }

  if (MODE == 0) {
    // thread tx == 0 of each warp holds the finished rows; block-ordered loss partial
    float* s_loss = Ps;
    if (tx == 0) {
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const int i = row0 + ty * 4 + r;
        float li = 0.f;
        if (i < n) {
          const float v = m_run[r] + logf(l_run[r]);
          lse[i] = v;
          li = v - pos[i];
        }
        s_loss[ty * 4 + r] = li;
      }
    }
    __syncthreads();
    if (tid == 0) {
      float t = 0.f;
      for (int i = 0; i < NCE_TM; ++i) t += s_loss[i];
      loss_part[blockIdx.x] = t;
    }
    return;
  }

  // positive-pair term, scale, and the Jacobian of x -> x / max(|x|, eps), all local to the row's 8 threads
  const int i = row0 + orow;
  const bool iv = i < n;
  const float ppos = iv ? expf(pos[i] - lse[i]) - 1.f : 0.f;  // d loss_i / d pos_i (the positive sits at label 0)
  const float den = iv ? own_den[i] : 1.f;
  float dotp = 0.f;
#pragma unroll
  for (int c = 0; c < CPT; ++c) {
    const int col = oc0 + 8 * c;
    if (col < D && iv) {
      acc[c] = (acc[c] + ppos * other[(size_t)i * D + col]) * coef;
      dotp = fmaf(As[orow * AS + col], acc[c], dotp);
    }
  }
  dotp += __shfl_xor_sync(0xffffffffu, dotp, 1);
  dotp += __shfl_xor_sync(0xffffffffu, dotp, 2);
  dotp += __shfl_xor_sync(0xffffffffu, dotp, 4);
  if (!iv) return;
  const bool clamped = den <= NCE_EPS;  // F.normalize's clamp: the quotient is then linear in x
  const int64_t r_out = rows ? rows[(size_t)i * row_stride] : i;
#pragma unroll
  for (int c = 0; c < CPT; ++c) {
    const int col = oc0 + 8 * c;
    if (col < D) {
      const float gx = clamped ? acc[c] / den : (acc[c] - As[orow * AS + col] * dotp) / den;
      if (rows) atomicAdd(g_out + (size_t)r_out * D + col, gx);
      else g_out[(size_t)r_out * D + col] = gx;
    }
  }
}

__global__ void infonce_loss_kernel(const float* loss_part, int n_part, float scale, float* loss_out) {
  pdl_wait();
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    float t = 0.f;
    for (int i = 0; i < n_part; ++i) t += loss_part[i];
    loss_out[0] += t * scale;
  }
}

template <int D>
static int infonce_launch(const float* q, const float* k, const int64_t* rows, int row_stride, int n, float temperature,
                          float loss_scale, float* loss_out, float* gq, float* gk, float* ws, cudaStream_t st) {
  float* qn = ws;
  float* kn = qn + (size_t)n * D;
  float* qden = kn + (size_t)n * D;
  float* kden = qden + n;
  float* pos = kden + n;
  float* lse = pos + n;
  const int nblk = ceil_div(n, NCE_TM);
  float* part = lse + n;
  const float inv_t = 1.f / temperature;
  const size_t smem = (size_t)(NCE_TM * D + NCE_TN * (D + 1) + NCE_TM * (NCE_TN + 1) + NCE_TN) * sizeof(float);
  static bool attr_done = false;
  if (!attr_done) {
    B2_CUDA(cudaFuncSetAttribute(infonce_pass_kernel<D, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    B2_CUDA(cudaFuncSetAttribute(infonce_pass_kernel<D, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    B2_CUDA(cudaFuncSetAttribute(infonce_pass_kernel<D, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_done = true;
  }
  B2_LAUNCH_PDL(infonce_prep_kernel<D>, ceil_div((long long)n * 32, 256), 256, 0, st, q, k, rows, row_stride, n, inv_t, qn, kn,
                qden, kden, pos);
  B2_LAUNCH_PDL(infonce_pass_kernel<D, 0>, nblk, 256, smem, st, (const float*)qn, (const float*)kn, (const float*)qden,
                (const float*)pos, lse, part, n, inv_t, 0.f, (const int64_t*)nullptr, 0, (float*)nullptr);
  B2_LAUNCH_PDL(infonce_loss_kernel, 1, 32, 0, st, (const float*)part, nblk, loss_scale / (float)n, loss_out);
  const float coef = loss_scale * inv_t / (float)n;
  B2_LAUNCH_PDL(infonce_pass_kernel<D, 1>, nblk, 256, smem, st, (const float*)qn, (const float*)kn, (const float*)qden,
                (const float*)pos, lse, part, n, inv_t, coef, rows, row_stride, gq);
  B2_LAUNCH_PDL(infonce_pass_kernel<D, 2>, nblk, 256, smem, st, (const float*)kn, (const float*)qn, (const float*)kden,
                (const float*)pos, lse, part, n, inv_t, coef, rows, row_stride, gk);
  return 0;
}

}  // namespace b200rec

using namespace b200rec;

extern "C" int64_t b200rec_infonce_workspace_floats(int32_t n, int32_t d) {
  if (n <= 0 || d <= 0) return 0;
  return 2ll * n * d + 4ll * n + (n + NCE_TM - 1) / NCE_TM + 64;
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
