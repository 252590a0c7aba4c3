// Device-side construction of the sparse operands (one-time, per graph):
//   b200rec_csr_build  -- COO pairs -> coalesced int32 CSR (utils.py:42-50 generate_daj_mat's COO->CSR with duplicates
//                         summed; utils.py:33-39 get_sparse_tensor's coalesce: row-major, columns ascending);
//   b200rec_plan_build -- CSR -> the load-balanced work decomposition spmm_items_kernel consumes, optionally cut into
//                         passes over blocks of source rows that fit L2 (b200rec_spmm_f32_blocked).
// Setup-time entry points: they allocate their own scratch (cudaMalloc), use CUB's radix sort / scan / run-length
// encode for the index plumbing, and synchronise the stream before returning their sizes.
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_run_length_encode.cuh>
#include <cub/device/device_scan.cuh>
#include "common.cuh"

namespace b200rec {

struct Scratch {  // cudaMalloc'ed scratch released on every exit path
  void* p[16];
  int n = 0;
  ~Scratch() {
    for (int i = 0; i < n; ++i) cudaFree(p[i]);
  }
  template <typename T>
  cudaError_t get(T** out, size_t count) {
    void* q = nullptr;
    cudaError_t e = cudaMalloc(&q, (count ? count : 1) * sizeof(T));
    if (e == cudaSuccess) p[n++] = q;
    *out = (T*)q;
    return e;
  }
};

// ---------------------------------------------------------------------------------------------------- COO -> CSR
__global__ void coo_keys_kernel(const int64_t* rows, const int64_t* cols, int64_t n, int64_t row_off, int64_t col_off,
                                uint64_t* keys, int32_t* pos, int32_t n_rows, int32_t n_cols, int32_t* bad) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int64_t r = rows[i] + row_off, c = cols[i] + col_off;
  if (r < 0 || r >= n_rows || c < 0 || c >= n_cols) atomicExch(bad, 1);
  keys[i] = ((uint64_t)(uint32_t)r << 32) | (uint32_t)c;
  pos[i] = (int32_t)i;
}
// the symmetric bipartite adjacency of utils.py:42-50: entry i < E is (u, U + item), entry E + i its mirror
__global__ void bipartite_keys_kernel(const int64_t* users, const int64_t* items, int64_t e, int32_t n_users, int32_t n_items,
                                      uint64_t* keys, int32_t* pos, int32_t* bad) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= e) return;
  const int64_t u = users[i], it = items[i];
  if (u < 0 || u >= n_users || it < 0 || it >= n_items) atomicExch(bad, 1);
  const uint32_t r = (uint32_t)u, c = (uint32_t)(it + n_users);
  keys[i] = ((uint64_t)r << 32) | c;
  keys[e + i] = ((uint64_t)c << 32) | r;
  pos[i] = (int32_t)i;
  pos[e + i] = (int32_t)(e + i);
}
__global__ void csr_emit_kernel(const uint64_t* ukeys, const int32_t* counts, int32_t nnz, int32_t* colidx, float* mult) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nnz) return;
  colidx[i] = (int32_t)(uint32_t)ukeys[i];
  if (mult) mult[i] = (float)counts[i];
}
// rowptr[r] = first unique key whose row is >= r (keys ascend); one thread per row boundary
__global__ void csr_rowptr_kernel(const uint64_t* ukeys, int32_t nnz, int32_t n_rows, int32_t* rowptr) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r > n_rows) return;
  const uint64_t target = (uint64_t)(uint32_t)r << 32;
  int lo = 0, hi = nnz;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (ukeys[mid] < target) lo = mid + 1; else hi = mid;
  }
  rowptr[r] = lo;
}
// position (in the caller's entry list) of the first entry merged into each output nnz: counts' exclusive scan indexes
// the sorted entry list, whose payload is the original position
__global__ void csr_first_pos_kernel(const int32_t* sorted_pos, const int32_t* count_scan, int32_t nnz, int32_t* first_pos) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < nnz) first_pos[i] = sorted_pos[count_scan[i]];
}

static int csr_from_keys(Scratch& sc, uint64_t* keys, int32_t* pos, int64_t n, int32_t n_rows, int32_t n_cols,
                         int32_t* rowptr, int32_t* colidx, float* mult, int32_t* first_pos, int32_t* nnz_out,
                         int32_t* has_dup_out, cudaStream_t st) {
  uint64_t *keys2 = nullptr, *ukeys = nullptr;
  int32_t *pos2 = nullptr, *counts = nullptr, *nruns = nullptr, *scan = nullptr;
  B2_CUDA(sc.get(&keys2, n)); B2_CUDA(sc.get(&pos2, n)); B2_CUDA(sc.get(&ukeys, n)); B2_CUDA(sc.get(&counts, n));
  B2_CUDA(sc.get(&nruns, 1));
  int row_bits = 1, col_bits = 1;
  while ((1ll << row_bits) < n_rows) ++row_bits;
  while ((1ll << col_bits) < n_cols) ++col_bits;
  size_t tb = 0, tb2 = 0, tb3 = 0;
  // LSD radix sort is stable: entries with equal (row, col) keep their list order, so first_pos is the FIRST duplicate
  B2_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tb, keys, keys2, pos, pos2, n, 0, 32 + row_bits, st));
  B2_CUDA(cub::DeviceRunLengthEncode::Encode(nullptr, tb2, keys2, ukeys, counts, nruns, n, st));
  B2_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tb3, counts, counts, n, st));
  if (tb2 > tb) tb = tb2;
  if (tb3 > tb) tb = tb3;
  uint8_t* tmp = nullptr;
  B2_CUDA(sc.get(&tmp, tb));
  (void)col_bits;
  B2_CUDA(cub::DeviceRadixSort::SortPairs(tmp, tb, keys, keys2, pos, pos2, n, 0, 32 + row_bits, st));
  B2_CUDA(cub::DeviceRunLengthEncode::Encode(tmp, tb, keys2, ukeys, counts, nruns, n, st));
  int32_t nnz = 0;
  B2_CUDA(cudaMemcpyAsync(&nnz, nruns, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
  B2_CUDA(cudaStreamSynchronize(st));
  *nnz_out = nnz;
  *has_dup_out = (nnz != n) ? 1 : 0;
  if (nnz > 0) {
    csr_emit_kernel<<<ceil_div(nnz, 256), 256, 0, st>>>(ukeys, counts, nnz, colidx, mult);
    B2_LAUNCHED();
  }
  csr_rowptr_kernel<<<ceil_div(n_rows + 1, 256), 256, 0, st>>>(ukeys, nnz, n_rows, rowptr);
  B2_LAUNCHED();
  if (first_pos && nnz > 0) {
    B2_CUDA(sc.get(&scan, nnz));
    B2_CUDA(cub::DeviceScan::ExclusiveSum(tmp, tb, counts, scan, nnz, st));
    csr_first_pos_kernel<<<ceil_div(nnz, 256), 256, 0, st>>>(pos2, scan, nnz, first_pos);
    B2_LAUNCHED();
  }
  B2_CUDA(cudaStreamSynchronize(st));
  return 0;
}

// ---------------------------------------------------------------------------------------------------- work plan
struct PlanIn {
  const int32_t* rowptr;
  const int32_t* colidx;
  int n_rows, chunk, order_split, n_blocks, row_order;
  int col_sort;           // single-pass plans: full hub chunks ordered by first source column (see walk_row)
  const int32_t* bounds;  // device [n_blocks + 1]
};

__device__ __forceinline__ int lower_bound_i32(const int32_t* a, int lo, int hi, int key) {
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (__ldg(a + mid) < key) lo = mid + 1; else hi = mid;
  }
  return lo;
}
__device__ __forceinline__ int block_of(const int32_t* bounds, int n_blocks, int col) {  // bounds[b] <= col < bounds[b+1]
  int lo = 0, hi = n_blocks;
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (__ldg(bounds + mid) <= col) lo = mid; else hi = mid;
  }
  return lo;
}

// Walks the pieces of one row (a row of at most `chunk` entries is one piece, a longer one ceil(len/chunk) pieces with
// their own slots) and, inside a piece, its slices per block of source columns.  EMIT = false only counts.
template <bool EMIT>
__device__ __forceinline__ void walk_row(const PlanIn& in, int r, int& n_items, int& n_slots, int item_off, int slot_off,
                                         int long_idx, int32_t* item_start, int32_t* item_end, int32_t* item_dst,
                                         int32_t* item_row, uint64_t* item_key, int32_t* slot_long) {
  const int s = __ldg(in.rowptr + r), e = __ldg(in.rowptr + r + 1);
  const int len = e - s;
  const bool is_long = len > in.chunk;
  const int pieces = is_long ? (len + in.chunk - 1) / in.chunk : 1;
  int ni = 0;
  const uint64_t order_hi = (in.n_blocks <= 1 && in.order_split > 0 && r >= in.order_split) ? 1 : 0;
  for (int q = 0; q < pieces; ++q) {
    const int vs = is_long ? s + q * in.chunk : s;
    const int ve = is_long ? min(vs + in.chunk, e) : e;
    if (EMIT && is_long) slot_long[slot_off + q] = long_idx;
    const uint32_t idcode = is_long ? (uint32_t)(slot_off + q) : (uint32_t)r;
    if (ve == vs || in.n_blocks <= 1) {  // empty row, or no column blocking: one item
      if (EMIT) {
        const int k = item_off + ni;
        item_start[k] = vs; item_end[k] = ve;
        item_dst[k] = is_long ? (int32_t)~idcode : (int32_t)idcode;
        item_row[k] = r;
        // phase, then longest first; FULL chunks of hub rows (all the same length) are then ordered by their first source
        // column instead of row by row (B200REC_PLAN_COLSORT, default on): the chunks in flight at one time cover one
        // window of source rows across many hub rows -- the dense hub rows of a power-law graph share most of it -- so that
        // window is fetched from DRAM once instead of once per hub row.  Chunk boundaries and the slot-ordered sum are
        // untouched: the same bits.
        const uint64_t len_key = in.row_order ? 0u : (uint32_t)(0x7fffffff - (ve - vs));
        const uint64_t col_key = (in.col_sort && is_long && (ve - vs) == in.chunk) ? (uint32_t)__ldg(in.colidx + vs) : 0u;
        // (column-blocked plans keep the pass number in the high word: an empty row is an item of pass 0)
        item_key[k] = (in.n_blocks <= 1) ? ((order_hi << 63) | (len_key << 32) | col_key) : len_key;
      }
      ++ni;
      continue;
    }
    const int b_lo = block_of(in.bounds, in.n_blocks, __ldg(in.colidx + vs));
    const int b_hi = block_of(in.bounds, in.n_blocks, __ldg(in.colidx + ve - 1));
    int pos = vs;
    bool first = true;
    for (int b = b_lo; b <= b_hi; ++b) {
      const int nxt = (b == b_hi) ? ve : lower_bound_i32(in.colidx, pos, ve, __ldg(in.bounds + b + 1));
      if (nxt > pos) {
        if (EMIT) {
          const int k = item_off + ni;
          uint32_t code = idcode;
          if (!first) code |= 1u << 29;      // continues a running sum
          if (nxt < ve) code |= 1u << 30;    // parks it again
          item_start[k] = pos; item_end[k] = nxt;
          item_dst[k] = is_long ? (int32_t)~code : (int32_t)code;
          item_row[k] = r;
          // inside a pass: longest first (one work item per launch slot), or -- for the record stream -- the emission
          // order itself (rows ascending: the stable sort keeps it), so that carried sums and outputs stream through DRAM
          item_key[k] = ((uint64_t)b << 32) | (in.row_order ? 0u : (uint32_t)(0x7fffffff - (nxt - pos)));
        }
        ++ni;
        first = false;
      }
      pos = nxt;
    }
  }
  n_items = ni;
  n_slots = is_long ? pieces : 0;
}

__global__ void plan_count_kernel(const PlanIn in, int32_t* cnt_items, int32_t* cnt_slots, int32_t* cnt_long) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= in.n_rows) return;
  int ni, ns;
  walk_row<false>(in, r, ni, ns, 0, 0, 0, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr);
  cnt_items[r] = ni; cnt_slots[r] = ns; cnt_long[r] = ns > 0 ? 1 : 0;
}
__global__ void plan_emit_kernel(const PlanIn in, const int32_t* off_items, const int32_t* off_slots, const int32_t* off_long,
                                 int32_t* item_start, int32_t* item_end, int32_t* item_dst, int32_t* item_row,
                                 uint64_t* item_key, int32_t* long_row, int32_t* long_slot0, int32_t* long_nslot,
                                 int32_t* slot_long) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= in.n_rows) return;
  int ni, ns;
  const int li = off_long[r];
  walk_row<true>(in, r, ni, ns, off_items[r], off_slots[r], li, item_start, item_end, item_dst, item_row, item_key, slot_long);
  if (ns > 0) { long_row[li] = r; long_slot0[li] = off_slots[r]; long_nslot[li] = ns; }
}
__global__ void iota_kernel(int32_t* a, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) a[i] = i;
}
__global__ void plan_permute_kernel(const int32_t* perm, int n, const int32_t* s0, const int32_t* s1, const int32_t* s2,
                                    const int32_t* s3, int32_t* d0, int32_t* d1, int32_t* d2, int32_t* d3) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int j = perm[i];
  d0[i] = s0[j]; d1[i] = s1[j]; d2[i] = s2[j]; d3[i] = s3[j];
}
// pass_ptr[b] = first sorted item whose pass is >= b
__global__ void plan_pass_ptr_kernel(const uint64_t* sorted_keys, int n, int n_passes, int32_t* pass_ptr) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b > n_passes) return;
  const uint64_t target = (uint64_t)b << 32;
  int lo = 0, hi = n;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (sorted_keys[mid] < target) lo = mid + 1; else hi = mid;
  }
  pass_ptr[b] = lo;
}

}  // namespace b200rec

using namespace b200rec;

extern "C" int b200rec_csr_build(const int64_t* rows, const int64_t* cols, int64_t n, int64_t row_offset, int64_t col_offset,
                                 int32_t n_rows, int32_t n_cols, int32_t* rowptr, int32_t* colidx, float* mult,
                                 int32_t* first_pos, int32_t* nnz_out, int32_t* has_dup_out, void* stream) {
  B2_REQUIRE(rows && cols && rowptr && colidx && nnz_out && has_dup_out, "null argument");
  B2_REQUIRE(n >= 0 && n < (1ll << 31) && n_rows > 0 && n_cols > 0, "bad size (entries must fit int32)");
  cudaStream_t st = (cudaStream_t)stream;
  Scratch sc;
  uint64_t* keys = nullptr;
  int32_t *pos = nullptr, *bad = nullptr;
  B2_CUDA(sc.get(&keys, n)); B2_CUDA(sc.get(&pos, n)); B2_CUDA(sc.get(&bad, 1));
  B2_CUDA(cudaMemsetAsync(bad, 0, sizeof(int32_t), st));
  if (n > 0) {
    coo_keys_kernel<<<ceil_div(n, 256), 256, 0, st>>>(rows, cols, n, row_offset, col_offset, keys, pos, n_rows, n_cols, bad);
    B2_LAUNCHED();
  }
  int32_t hbad = 0;
  B2_CUDA(cudaMemcpyAsync(&hbad, bad, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
  B2_CUDA(cudaStreamSynchronize(st));
  B2_REQUIRE(!hbad, "an index is outside [0, n_rows) x [0, n_cols)");
  return csr_from_keys(sc, keys, pos, n, n_rows, n_cols, rowptr, colidx, mult, first_pos, nnz_out, has_dup_out, st);
}

extern "C" int b200rec_adj_build(const int64_t* users, const int64_t* items, int64_t n_pairs, int32_t n_users, int32_t n_items,
                                 int32_t* rowptr, int32_t* colidx, float* mult, int32_t* nnz_out, int32_t* has_dup_out,
                                 void* stream) {
  B2_REQUIRE(users && items && rowptr && colidx && nnz_out && has_dup_out, "null argument");
  B2_REQUIRE(n_pairs >= 0 && 2 * n_pairs < (1ll << 31) && n_users > 0 && n_items > 0, "bad size (2E must fit int32)");
  B2_REQUIRE((long long)n_users + n_items < (1ll << 31), "too many nodes");
  cudaStream_t st = (cudaStream_t)stream;
  Scratch sc;
  uint64_t* keys = nullptr;
  int32_t *pos = nullptr, *bad = nullptr;
  const int64_t n = 2 * n_pairs;
  B2_CUDA(sc.get(&keys, n)); B2_CUDA(sc.get(&pos, n)); B2_CUDA(sc.get(&bad, 1));
  B2_CUDA(cudaMemsetAsync(bad, 0, sizeof(int32_t), st));
  if (n_pairs > 0) {
    bipartite_keys_kernel<<<ceil_div(n_pairs, 256), 256, 0, st>>>(users, items, n_pairs, n_users, n_items, keys, pos, bad);
    B2_LAUNCHED();
  }
  int32_t hbad = 0;
  B2_CUDA(cudaMemcpyAsync(&hbad, bad, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
  B2_CUDA(cudaStreamSynchronize(st));
  B2_REQUIRE(!hbad, "a user or item id is out of range");
  return csr_from_keys(sc, keys, pos, n, n_users + n_items, n_users + n_items, rowptr, colidx, mult, nullptr, nnz_out,
                       has_dup_out, st);
}

extern "C" int b200rec_plan_build(const int32_t* rowptr, const int32_t* colidx, int32_t n_rows, int32_t chunk,
                                  int32_t order_split, const int32_t* col_bounds, int32_t n_blocks, int32_t row_order,
                                  int32_t* sizes,
                                  int32_t* item_start, int32_t* item_end, int32_t* item_dst, int32_t* item_row,
                                  int32_t* long_row, int32_t* long_slot0, int32_t* long_nslot, int32_t* slot_long,
                                  int32_t* pass_ptr, void* stream) {
  B2_REQUIRE(rowptr && colidx && sizes && n_rows > 0 && chunk >= 32, "bad argument");
  B2_REQUIRE(n_blocks >= 1 && (n_blocks == 1 || col_bounds), "col_bounds missing");
  B2_REQUIRE(n_rows < (1 << 29), "at most 2^29 rows");
  cudaStream_t st = (cudaStream_t)stream;
  Scratch sc;
  PlanIn in;
  in.rowptr = rowptr; in.colidx = colidx; in.n_rows = n_rows; in.chunk = chunk; in.order_split = order_split;
  in.n_blocks = n_blocks; in.bounds = nullptr; in.row_order = row_order;
  {
    // column-ordered hub chunks pay when the gathered table is much larger than L2 (C4: -2.6 % per step); on the L2-resident
    // shapes they only bunch the hub rows' finishers together (C2: 0.384 -> 0.390 ms/step), so: graphs of >= 16 M entries
    int32_t nnz = 0;
    B2_CUDA(cudaMemcpyAsync(&nnz, rowptr + n_rows, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    B2_CUDA(cudaStreamSynchronize(st));
    const char* e = getenv("B200REC_PLAN_COLSORT");
    in.col_sort = e ? (atoi(e) != 0) : (nnz >= (1 << 24));
  }
  if (n_blocks > 1) {
    for (int b = 0; b < n_blocks; ++b) B2_REQUIRE(col_bounds[b] <= col_bounds[b + 1], "col_bounds must ascend");
    int32_t* dbounds = nullptr;
    B2_CUDA(sc.get(&dbounds, (size_t)n_blocks + 1));
    B2_CUDA(cudaMemcpyAsync(dbounds, col_bounds, sizeof(int32_t) * ((size_t)n_blocks + 1), cudaMemcpyHostToDevice, st));
    in.bounds = dbounds;
  }
  int32_t *cnt = nullptr, *off = nullptr;  // [3][n_rows + 1]: items, slots, long rows
  const size_t stride = (size_t)n_rows + 1;
  B2_CUDA(sc.get(&cnt, 3 * stride)); B2_CUDA(sc.get(&off, 3 * stride));
  B2_CUDA(cudaMemsetAsync(cnt, 0, 3 * stride * sizeof(int32_t), st));
  const int grid_rows = ceil_div(n_rows, 128);
  plan_count_kernel<<<grid_rows, 128, 0, st>>>(in, cnt, cnt + stride, cnt + 2 * stride);
  B2_LAUNCHED();
  size_t tb = 0, tb2 = 0;
  B2_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tb, cnt, off, (int)stride, st));
  uint8_t* tmp = nullptr;
  B2_CUDA(sc.get(&tmp, tb));
  for (int k = 0; k < 3; ++k) B2_CUDA(cub::DeviceScan::ExclusiveSum(tmp, tb, cnt + k * stride, off + k * stride, (int)stride, st));
  int32_t tot[3];
  for (int k = 0; k < 3; ++k)
    B2_CUDA(cudaMemcpyAsync(&tot[k], off + k * stride + n_rows, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
  B2_CUDA(cudaStreamSynchronize(st));
  const int n_items = tot[0], n_slots = tot[1], n_long = tot[2];
  B2_REQUIRE(n_slots < (1 << 29), "too many hub pieces");
  sizes[0] = n_items; sizes[1] = n_long; sizes[2] = n_slots;
  if (!item_start) return 0;  // size query
  B2_REQUIRE(item_end && item_dst && item_row && pass_ptr, "null output");
  B2_REQUIRE(n_long == 0 || (long_row && long_slot0 && long_nslot && slot_long), "null long-row output");
  int32_t *u0 = nullptr, *u1 = nullptr, *u2 = nullptr, *u3 = nullptr, *perm = nullptr, *perm2 = nullptr, *dpass = nullptr;
  uint64_t *key = nullptr, *key2 = nullptr;
  B2_CUDA(sc.get(&u0, n_items)); B2_CUDA(sc.get(&u1, n_items)); B2_CUDA(sc.get(&u2, n_items)); B2_CUDA(sc.get(&u3, n_items));
  B2_CUDA(sc.get(&perm, n_items)); B2_CUDA(sc.get(&perm2, n_items)); B2_CUDA(sc.get(&key, n_items)); B2_CUDA(sc.get(&key2, n_items));
  B2_CUDA(sc.get(&dpass, (size_t)n_blocks + 1));
  plan_emit_kernel<<<grid_rows, 128, 0, st>>>(in, off, off + stride, off + 2 * stride, u0, u1, u2, u3, key, long_row, long_slot0,
                                               long_nslot, slot_long);
  B2_LAUNCHED();
  iota_kernel<<<ceil_div(n_items, 256), 256, 0, st>>>(perm, n_items);
  B2_LAUNCHED();
  // stable sort by (pass, longest first): equal keys keep their (row, piece, block) emission order -> deterministic plan
  B2_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tb2, key, key2, perm, perm2, n_items, 0, 64, st));
  uint8_t* tmp2 = nullptr;
  B2_CUDA(sc.get(&tmp2, tb2));
  B2_CUDA(cub::DeviceRadixSort::SortPairs(tmp2, tb2, key, key2, perm, perm2, n_items, 0, 64, st));
  plan_permute_kernel<<<ceil_div(n_items, 256), 256, 0, st>>>(perm2, n_items, u0, u1, u2, u3, item_start, item_end, item_dst,
                                                              item_row);
  B2_LAUNCHED();
  if (n_blocks > 1) {
    plan_pass_ptr_kernel<<<ceil_div(n_blocks + 1, 128), 128, 0, st>>>(key2, n_items, n_blocks, dpass);
    B2_LAUNCHED();
    B2_CUDA(cudaMemcpyAsync(pass_ptr, dpass, sizeof(int32_t) * ((size_t)n_blocks + 1), cudaMemcpyDeviceToHost, st));
  }
  B2_CUDA(cudaStreamSynchronize(st));
  if (n_blocks <= 1) { pass_ptr[0] = 0; pass_ptr[1] = n_items; }  // one launch; rows >= order_split are merely scheduled last
  return 0;
}

// ---------------------------------------------------------------------------------------------------- global top-m
// The global cut of DOSE's similarity mining (model.py:503-560: the flattened U x I cosine matrix goes through
// torch.topk(aug_num)): the m largest of n candidate values with their positions, largest first.  A descending radix sort
// on order-preserving keys (setup-time: once per epoch; n = U x 128 candidates from the per-row pass).
namespace b200rec {
__global__ void topk_keys_kernel(const float* __restrict__ vals, int64_t n, uint32_t* keys, int32_t* idx) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint32_t b = __float_as_uint(vals[i]);
  keys[i] = (b & 0x80000000u) ? ~b : (b | 0x80000000u);  // unsigned order == float order (NaN sorts above +inf)
  idx[i] = (int32_t)i;
}
__global__ void topk_take_kernel(const int32_t* sorted_idx, const float* vals, int m, int32_t* out_idx, float* out_vals) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= m) return;
  const int j = sorted_idx[i];
  out_idx[i] = j;
  if (out_vals) out_vals[i] = vals[j];
}
}  // namespace b200rec

extern "C" int b200rec_topk_global(const float* vals, int64_t n, int32_t m, int32_t* out_idx, float* out_vals, void* stream) {
  B2_REQUIRE(vals && out_idx && n > 0 && n < (1ll << 31) && m > 0 && m <= n, "bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  Scratch sc;
  uint32_t *k0 = nullptr, *k1 = nullptr;
  int32_t *i0 = nullptr, *i1 = nullptr;
  B2_CUDA(sc.get(&k0, n)); B2_CUDA(sc.get(&k1, n)); B2_CUDA(sc.get(&i0, n)); B2_CUDA(sc.get(&i1, n));
  topk_keys_kernel<<<ceil_div(n, 256), 256, 0, st>>>(vals, n, k0, i0);
  B2_LAUNCHED();
  size_t tb = 0;
  B2_CUDA(cub::DeviceRadixSort::SortPairsDescending(nullptr, tb, k0, k1, i0, i1, (int)n, 0, 32, st));
  uint8_t* tmp = nullptr;
  B2_CUDA(sc.get(&tmp, tb));
  B2_CUDA(cub::DeviceRadixSort::SortPairsDescending(tmp, tb, k0, k1, i0, i1, (int)n, 0, 32, st));  // stable: ties keep position order
  topk_take_kernel<<<ceil_div(m, 256), 256, 0, st>>>(i1, vals, m, out_idx, out_vals);
  B2_LAUNCHED();
  B2_CUDA(cudaStreamSynchronize(st));
  return 0;
}
