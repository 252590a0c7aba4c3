/* CPU restatement (plain C) of the integer / order-sensitive parts of the hot path.
 * TEST INFRASTRUCTURE ONLY: loaded by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg through
 * oracle/oracle_c.py; the product never links it.
 *
 *  - Philox4x32-10 streams of the device BPR sampler and edge-dropout mask.  The reference's sampler
 *    (dataset.py:119-131: Python `random` MT19937 + NumPy legacy choice) and dropout draw (model.py:4019-4021:
 *    CPU torch.rand) cannot be reproduced on a GPU; the contract is "same distribution, stated stream", and this
 *    file is the bit-exact statement of that stream.  Parity unpinned against the reference for these two
 *    (nothing to pin: different RNG by design); distribution checks live in tests/test_oracle_sampler.py.
 *  - fp32 scores accumulated with fmaf in d = 0..D-1 order + masking + top-K ordered (score desc, id asc):
 *    model.py:122-127 predict, trainer.py:152-169.  Pinned against tests/golden (ids equal wherever the
 *    reference's own adjacent scores differ by more than 1e-6).
 *  - CSR SpMM in CSR order with fmaf, including the chunked order used for rows longer than `chunk`:
 *    model.py:106 gspmm 'mul','sum'.  Pinned against tests/golden to 1e-5 relative.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

static void philox_round(uint32_t c[4], uint32_t k0, uint32_t k1) {
  const uint64_t p0 = (uint64_t)0xD2511F53u * c[0], p1 = (uint64_t)0xCD9E8D57u * c[2];
  const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c[1] ^ k0, n1 = (uint32_t)p1;
  const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c[3] ^ k1, n3 = (uint32_t)p0;
  c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
}
void oracle_philox4x32_10(uint32_t c[4], uint32_t k0, uint32_t k1) {
  for (int i = 0; i < 10; ++i) {
    philox_round(c, k0, k1);
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
}
static uint64_t mulhi64(uint64_t a, uint64_t b) { return (uint64_t)(((unsigned __int128)a * b) >> 64); }
static void philox_pair(uint64_t seed, uint32_t slot, uint64_t step, uint32_t draw, uint64_t* a, uint64_t* b) {
  uint32_t c[4] = {slot, (uint32_t)step, draw, (uint32_t)(step >> 32)};
  oracle_philox4x32_10(c, (uint32_t)seed, (uint32_t)(seed >> 32));
  *a = (uint64_t)c[0] | ((uint64_t)c[1] << 32);
  *b = (uint64_t)c[2] | ((uint64_t)c[3] << 32);
}
static int row_contains(const int32_t* row, int n, int key) {
  int lo = 0, hi = n;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (row[mid] < key) lo = mid + 1; else hi = mid;
  }
  return lo < n && row[lo] == key;
}

/* BasicDataset.__getitem__ (dataset.py:119-131) with the stated Philox stream */
void oracle_bpr_sample(const int32_t* user_ptr, const int32_t* user_items, int32_t n_users, int32_t n_items,
                       uint64_t seed, uint64_t step, int32_t batch, int64_t* out) {
  for (int32_t slot = 0; slot < batch; ++slot) {
    uint32_t draw = 0;
    uint64_t a, b;
    int user = 0, s = 0, deg = 0;
    for (;; ++draw) {
      philox_pair(seed, (uint32_t)slot, step, draw, &a, &b);
      user = (int)mulhi64(a, (uint64_t)n_users);
      s = user_ptr[user];
      deg = user_ptr[user + 1] - s;
      if (deg > 0 || draw > (1u << 20)) break;
    }
    const int32_t* row = user_items + s;
    const int pos = deg > 0 ? row[(int)mulhi64(b, (uint64_t)deg)] : 0;
    int neg = 0;
    for (++draw;; ++draw) {
      philox_pair(seed, (uint32_t)slot, step, draw, &a, &b);
      neg = (int)mulhi64(a, (uint64_t)n_items);
      if (!row_contains(row, deg, neg)) break;
      neg = (int)mulhi64(b, (uint64_t)n_items);
      if (!row_contains(row, deg, neg) || draw > (1u << 20)) break;
    }
    out[3 * (size_t)slot] = user;
    out[3 * (size_t)slot + 1] = pos;
    out[3 * (size_t)slot + 2] = neg;
  }
}

/* NGCF.dropout_sp_mat keep mask (model.py:4016-4021): keep = floor((1-p) + r), r = 24-bit uniform */
void oracle_dropout_mask(int32_t nnz, float p, uint64_t seed, uint64_t step, uint32_t* bits) {
  const float keep_base = (float)(1.0 - (double)p);
  const int n_words = (nnz + 31) >> 5;
  for (int w = 0; w < n_words; ++w) {
    uint32_t word = 0;
    for (int j = 0; j < 8; ++j) {
      uint32_t c[4] = {(uint32_t)(w * 8 + j), (uint32_t)step, (uint32_t)(step >> 32), 0xD0u};
      oracle_philox4x32_10(c, (uint32_t)seed, (uint32_t)(seed >> 32));
      for (int q = 0; q < 4; ++q) {
        const float r = (float)(c[q] >> 8) * (1.0f / 16777216.0f);
        if (floorf(keep_base + r) >= 1.f) word |= 1u << (j * 4 + q);
      }
    }
    const int rem = nnz - w * 32;
    if (rem < 32) word &= (1u << rem) - 1u;
    bits[w] = word;
  }
}

/* scores[b, n_items] = U[users[b], :] . I[j, :], fmaf in d order (model.py:126 torch.mm restated with a fixed order) */
void oracle_scores_f32(const float* rep_users, const int64_t* users, int32_t nb, const float* rep_items, int32_t n_items,
                       int32_t d, float* scores) {
  for (int32_t b = 0; b < nb; ++b) {
    const float* u = rep_users + (size_t)users[b] * d;
    for (int32_t j = 0; j < n_items; ++j) {
      const float* v = rep_items + (size_t)j * d;
      float acc = 0.f;
      for (int32_t k = 0; k < d; ++k) acc = fmaf(u[k], v[k], acc);
      scores[(size_t)b * n_items + j] = acc;
    }
  }
}

typedef struct { float s; int32_t id; } cand_t;
static int cand_cmp(const void* pa, const void* pb) {
  const cand_t *a = (const cand_t*)pa, *b = (const cand_t*)pb;
  if (a->s > b->s) return -1;
  if (a->s < b->s) return 1;
  return (a->id > b->id) - (a->id < b->id);
}
/* trainer.py:152-169: mask train (+val) items and the banned range, top-K by (score desc, id asc); rows with fewer
 * than K unmasked items are padded with id -1 / -inf */
void oracle_mask_topk(const float* scores, const int64_t* users, int32_t nb, int32_t n_items,
                      const int32_t* ptr_a, const int32_t* idx_a, const int32_t* ptr_b, const int32_t* idx_b,
                      int32_t banned_lo, int32_t banned_hi, int32_t k, int32_t* out_ids, float* out_scores) {
  cand_t* buf = (cand_t*)malloc(sizeof(cand_t) * (size_t)n_items);
  for (int32_t b = 0; b < nb; ++b) {
    const int64_t u = users[b];
    int32_t n = 0;
    for (int32_t j = 0; j < n_items; ++j) {
      if (j >= banned_lo && j < banned_hi) continue;
      if (ptr_a && row_contains(idx_a + ptr_a[u], ptr_a[u + 1] - ptr_a[u], j)) continue;
      if (ptr_b && row_contains(idx_b + ptr_b[u], ptr_b[u + 1] - ptr_b[u], j)) continue;
      buf[n].s = scores[(size_t)b * n_items + j];
      buf[n].id = j;
      ++n;
    }
    qsort(buf, (size_t)n, sizeof(cand_t), cand_cmp);
    for (int32_t j = 0; j < k; ++j) {
      out_ids[(size_t)b * k + j] = j < n ? buf[j].id : -1;
      out_scores[(size_t)b * k + j] = j < n ? buf[j].s : -INFINITY;
    }
  }
  free(buf);
}

/* y = A x in CSR order with fmaf; rows longer than `chunk` are summed chunk by chunk and the chunk sums added in
 * order (the deterministic order of the device kernel) */
void oracle_spmm_csr(const int32_t* rowptr, const int32_t* colidx, const float* vals, int32_t n_rows, const float* x,
                     int32_t d, int32_t chunk, float* y) {
  float* acc = (float*)malloc(sizeof(float) * (size_t)d);
  float* tot = (float*)malloc(sizeof(float) * (size_t)d);
  for (int32_t r = 0; r < n_rows; ++r) {
    const int32_t s = rowptr[r], e = rowptr[r + 1];
    const int split = (e - s) > chunk;
    memset(tot, 0, sizeof(float) * (size_t)d);
    for (int32_t b = s; b < e || b == s; b += chunk) {
      const int32_t be = (b + chunk < e) ? b + chunk : e;
      memset(acc, 0, sizeof(float) * (size_t)d);
      for (int32_t k = b; k < be; ++k) {
        const float v = vals ? vals[k] : 1.f;
        const float* xr = x + (size_t)colidx[k] * d;
        for (int32_t c = 0; c < d; ++c) acc[c] = fmaf(v, xr[c], acc[c]);
      }
      if (!split) { memcpy(tot, acc, sizeof(float) * (size_t)d); break; }
      for (int32_t c = 0; c < d; ++c) tot[c] += acc[c];
      if (be >= e) break;
    }
    memcpy(y + (size_t)r * d, tot, sizeof(float) * (size_t)d);
  }
  free(acc);
  free(tot);
}

/* InfoNCE as SGL / HALF call it (model.py:206-214, :322-332; info_nce package, 'unpaired', negative_keys == positive_key,
 * reduction 'mean'): rows L2-normalised with the norm clamped at 1e-12, logits_i = [qn_i.kn_i, qn_i.kn_0 .. qn_i.kn_{n-1}]/T,
 * cross-entropy against index 0.  Plain double arithmetic; returns the loss and writes dLoss/dq, dLoss/dk (may be NULL). */
double oracle_infonce(const float* q, const float* k, int32_t n, int32_t d, double temperature, double* gq, double* gk) {
  double* qn = (double*)malloc(sizeof(double) * (size_t)n * d);
  double* kn = (double*)malloc(sizeof(double) * (size_t)n * d);
  double* qden = (double*)malloc(sizeof(double) * (size_t)n);
  double* kden = (double*)malloc(sizeof(double) * (size_t)n);
  double* dqn = (double*)calloc((size_t)n * d, sizeof(double));
  double* dkn = (double*)calloc((size_t)n * d, sizeof(double));
  double* row = (double*)malloc(sizeof(double) * (size_t)n);
  for (int32_t i = 0; i < n; ++i) {
    double sq = 0, sk = 0;
    for (int32_t c = 0; c < d; ++c) {
      sq += (double)q[(size_t)i * d + c] * q[(size_t)i * d + c];
      sk += (double)k[(size_t)i * d + c] * k[(size_t)i * d + c];
    }
    qden[i] = fmax(sqrt(sq), 1e-12);
    kden[i] = fmax(sqrt(sk), 1e-12);
    for (int32_t c = 0; c < d; ++c) {
      qn[(size_t)i * d + c] = q[(size_t)i * d + c] / qden[i];
      kn[(size_t)i * d + c] = k[(size_t)i * d + c] / kden[i];
    }
  }
  double loss = 0;
  for (int32_t i = 0; i < n; ++i) {
    double mx = -INFINITY;
    for (int32_t j = 0; j < n; ++j) {
      double s = 0;
      for (int32_t c = 0; c < d; ++c) s += qn[(size_t)i * d + c] * kn[(size_t)j * d + c];
      row[j] = s / temperature;
      if (row[j] > mx) mx = row[j];
    }
    const double pos = row[i];
    double z = exp(pos - mx);
    for (int32_t j = 0; j < n; ++j) z += exp(row[j] - mx);
    const double lse = mx + log(z);
    loss += lse - pos;
    /* d loss_i / d logits: softmax over [pos, row[0..n-1]] minus the one-hot at the positive slot */
    const double ppos = exp(pos - lse) - 1.0;
    for (int32_t j = 0; j < n; ++j) {
      double w = exp(row[j] - lse);
      if (j == i) w += ppos;
      w /= temperature * n;
      for (int32_t c = 0; c < d; ++c) {
        dqn[(size_t)i * d + c] += w * kn[(size_t)j * d + c];
        dkn[(size_t)j * d + c] += w * qn[(size_t)i * d + c];
      }
    }
  }
  /* Jacobian of x -> x / max(|x|, eps): (g - xn (xn.g)) / |x| above the clamp, g / eps below it */
  for (int side = 0; side < 2; ++side) {
    double* g = side ? gk : gq;
    if (!g) continue;
    const double* xn = side ? kn : qn;
    const double* den = side ? kden : qden;
    const double* dxn = side ? dkn : dqn;
    for (int32_t i = 0; i < n; ++i) {
      double dot = 0;
      for (int32_t c = 0; c < d; ++c) dot += xn[(size_t)i * d + c] * dxn[(size_t)i * d + c];
      for (int32_t c = 0; c < d; ++c)
        g[(size_t)i * d + c] = den[i] <= 1e-12 ? dxn[(size_t)i * d + c] / den[i]
                                               : (dxn[(size_t)i * d + c] - xn[(size_t)i * d + c] * dot) / den[i];
    }
  }
  free(qn); free(kn); free(qden); free(kden); free(dqn); free(dkn); free(row);
  return loss / n;
}

/* Precision / Recall / NDCG at each cut-off (calculate_metrics, trainer.py:115-144): hit = rec[u][j] in the user's eval
 * row (sorted CSR), gains 1/log2(j+2), ideal gains over min(len, k), mean over users with at least one eval item.
 * out: [3][n_topks] (precision, recall, ndcg rows); returns the number of users counted. */
int32_t oracle_rank_metrics(const int32_t* rec, int32_t n_users, int32_t k, const int32_t* eval_ptr, const int32_t* eval_idx,
                            const int32_t* topks, int32_t n_topks, double* out) {
  for (int32_t t = 0; t < 3 * n_topks; ++t) out[t] = 0;
  int32_t counted = 0;
  for (int32_t u = 0; u < n_users; ++u) {
    const int32_t b = eval_ptr[u], len = eval_ptr[u + 1] - eval_ptr[u];
    if (len <= 0) continue;
    ++counted;
    for (int32_t t = 0; t < n_topks; ++t) {
      const int32_t kk = topks[t] < k ? topks[t] : k;
      int32_t hits = 0;
      double dcg = 0, idcg = 0;
      for (int32_t j = 0; j < kk; ++j) {
        const double gain = 1.0 / log2((double)(j + 2));
        const int32_t id = rec[(size_t)u * k + j];
        if (id >= 0 && row_contains(eval_idx + b, len, id)) { ++hits; dcg += gain; }
        if (j < len) idcg += gain;
      }
      out[0 * n_topks + t] += (double)hits / topks[t];
      out[1 * n_topks + t] += (double)hits / len;
      out[2 * n_topks + t] += dcg / idcg;
    }
  }
  for (int32_t t = 0; t < 3 * n_topks && counted; ++t) out[t] /= counted;
  return counted;
}
