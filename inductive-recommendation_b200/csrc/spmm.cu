// CSR SpMM with fused epilogue: the propagation leaf (model.py:106, :4184, :4196 dgl.ops.gspmm 'mul','sum').
//
// Mapping: one group of G = D/4 lanes owns one work item (a row, or a chunk of a hub row).  Each lane keeps a
// float4 slice of the row sum, so a gathered neighbour row is one 128-bit load per lane, 256 B (D=64, two rows per
// warp) or 512 B (D=128) fully coalesced.  (Column id, value) pairs are loaded >= 16 at a time, coalesced, parked in
// shared memory and re-read with one broadcast LDS.64 per neighbour; U neighbour rows are in flight per lane before
// the FMAs retire them.  The sum runs in CSR order, so results do not depend on the grid, the item order or
// (multi-GPU) the row partition.  Launched with programmatic stream serialisation: the work-plan loads run while the
// previous kernel of the step drains (common.cuh).
#include <stdlib.h>
#include "common.cuh"

#ifndef B200REC_SPMM_SCAN
#define B200REC_SPMM_SCAN 4
#endif

namespace b200rec {

struct SpmmParams {
  const int32_t* colidx;
  const float* vals;
  const float* nbr_scale;
  const float* row_scale;
  const int32_t* eid;
  const uint32_t* keep_bits;
  const uint8_t* dst_flags;  // optional: only rows with a non-zero flag are computed (others left untouched)
  const uint8_t* src_flags;  // optional: gathered rows with a zero flag are known to be all-zero and are skipped
  const uint8_t* out_rows;   // optional: the rows where `out` / `addend` matter (out_mode)
  int out_mode;              // 1: out is written (addend read) only at flagged rows; 2: addend is zero outside them (not read)
  const int32_t* live_items; // optional: compacted list of the work items to run (b200rec_live_items), instead of all of them
  const int32_t* live_count; // device scalar: entries in live_items
  const int32_t* item_start;
  const int32_t* item_end;
  const int32_t* item_dst;
  const int32_t* item_row;
  int n_items;
  const float* x;
  float* y;
  const float* addend;
  float* out;
  float out_scale;
  float post_scale;
  float* partial;
  const int32_t* slot_long;   // [n_slots] index of the split row a slot belongs to
  const int32_t* long_row;
  const int32_t* long_slot0;
  const int32_t* long_nslot;
  int32_t* long_cnt;          // [n_long] arrival counters (zero between launches: the finisher resets its own)
  // row-partitioned multi-GPU: the same row is also stored into the peers' copies of y / out (all-gather by push)
  int n_peer;
  float* peer_y[8];
  float* peer_out[8];
  int cold_evict_first;  // HINT variant: rows not flagged hot are loaded L2::evict_first (else with the default policy)
};

// L2 residency control for tables larger than L2 (HINT variant).  The column ids then carry a "hot" flag in bit 31
// (b200rec.graph: the highest-degree source rows of each side, as many as fit the persisting set-aside): hot rows are
// gathered with an L2::evict_last policy, which the hardware honours inside the set-aside carved out by
// cudaLimitPersistingL2CacheSize (b200rec_l2_persist) -- without a set-aside the hint is a no-op, which is what round 1
// measured.  Coherent loads (no .nc): the gathered table is the previous kernel's output.
__device__ __forceinline__ uint64_t make_policy(int kind) {  // 0 evict_normal, 1 evict_first, 2 evict_last
  uint64_t p;
  if (kind == 2) asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
  else if (kind == 1) asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
  else asm volatile("createpolicy.fractional.L2::evict_normal.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ float4 ldc_f4_policy(const float* ptr, uint64_t pol) {
  float4 v;
  asm volatile("ld.global.L2::cache_hint.v4.f32 {%0, %1, %2, %3}, [%4], %5;"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(ptr), "l"(pol));
  return v;
}

#ifndef B200REC_SPMM_MINB
#define B200REC_SPMM_MINB 4      // resident 256-thread blocks per SM the D <= 128 kernels are compiled for (register cap)
#endif
#ifndef B200REC_SPMM_EBMUL
#define B200REC_SPMM_EBMUL 1     // edges staged per group and block = max(G, 16) * this (unfiltered variants)
#endif
#ifndef B200REC_SPMM_PREFETCH
#define B200REC_SPMM_PREFETCH 0  // 1: the next block's (column, value) pairs are loaded under the current block's gathers
#endif

template <int G, int VPL>
__device__ __forceinline__ void epilogue_row(const SpmmParams& p, int row, int gl, float4 (&acc)[VPL]) {
  constexpr int D = G * VPL * 4;
  float sc = p.post_scale;
  if (p.row_scale) sc *= p.row_scale[row];
  // a training step reads the layer sum only at the sampled rows, and its gradient G is zero outside them: the 2 x N x D
  // (forward) / N x D (backward) bytes of out / addend traffic per layer shrink to those rows
  const bool sel = !p.out_rows || ldc_u8(p.out_rows + row) != 0;
  const bool wr_out = sel || p.out_mode != 1, rd_add = sel || p.out_mode != 2;
#pragma unroll
  for (int t = 0; t < VPL; ++t) {
    const size_t off = (size_t)row * D + (size_t)(gl + t * G) * 4;
    float4 s = make_float4(acc[t].x * sc, acc[t].y * sc, acc[t].z * sc, acc[t].w * sc);
    if (p.y) {
      st_f4(p.y + off, s);
      for (int q = 0; q < p.n_peer; ++q)
        if (p.peer_y[q]) st_f4(p.peer_y[q] + off, s);  // NVLink store into the peer's table, overlapped with the gathers
    }
    if (p.out && wr_out) {
      float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
      if (p.addend && rd_add) a = ld_f4(p.addend + off);
      const float os = p.out_scale;
      const float4 o4 = make_float4((a.x + s.x) * os, (a.y + s.y) * os, (a.z + s.z) * os, (a.w + s.w) * os);
      st_f4(p.out + off, o4);
      for (int q = 0; q < p.n_peer; ++q)
        if (p.peer_out[q]) st_f4(p.peer_out[q] + off, o4);
    }
  }
}

// Inner-loop economics (ncu, C2 shape): the LSU data pipe is the busiest unit -- every gathered 256 B row costs two
// wavefronts whether it hits L1 or not -- so everything else is kept off that pipe and off the issue slots:
//  * (column, value) pairs are loaded G at a time, coalesced, and parked in shared memory as one 64-bit word; each
//    neighbour then costs ONE broadcast LDS.64 instead of two SHFLs;
//  * edges that contribute nothing (past the row end, dropped by the edge-dropout mask, or whose source row is
//    flagged all-zero) are squeezed out with a ballot/popc compaction before the gather loop; the slack of the last
//    sub-block is padded with (first column of the row, weight 0), so the loop body carries no predicates or selects.
template <int G, int VPL, bool HAS_VALS, bool HAS_NBR, bool HAS_MASK, bool HAS_EID, bool HAS_SRCF, bool HINT>
__global__ void __launch_bounds__(256, (VPL == 1) ? B200REC_SPMM_MINB : 2) spmm_items_kernel(const SpmmParams p) {
  constexpr int D = G * VPL * 4;
  constexpr int U = 8;                    // neighbour rows in flight per lane
  constexpr int GPW = 32 / G;             // groups per warp
  constexpr bool COMPACT = HAS_MASK || HAS_SRCF;
  // edges staged per group per block (narrow rows: several edges per lane).  The filtered variants scan SCAN x as many:
  // their index / flag loads are independent and pipeline, and the survivors of the wider window fill the U-deep gather
  // batches (a 16-edge window with ~30 % survivors issues 5 real loads and 3 pads per L2 round trip).
  constexpr int SCAN = !COMPACT ? 1 : (G >= 4 ? B200REC_SPMM_SCAN : 2);  // G = 2: 128 groups per block, shared-memory budget
  constexpr int EB = ((G >= 16) ? G : 16) * SCAN * (COMPACT ? 1 : B200REC_SPMM_EBMUL);
  constexpr int EPL = EB / G;             // edges loaded per lane per block
  constexpr int CVS = (G == 32) ? EB : EB + 1;  // slot stride per group: +1 keeps the groups' broadcast reads off one bank
  __shared__ int2 s_cv[(256 / G) * CVS];
  const int lane = threadIdx.x & 31;
  const int gl = lane & (G - 1);
  const int warp = (int)((blockIdx.x * (unsigned)blockDim.x + threadIdx.x) >> 5);
  int item = warp * GPW + lane / G;       // position in the plan (or in the live list)
  int it = item;                          // plan index
  int2* my_cv = s_cv + (threadIdx.x / G) * CVS;  // this group's EB slots
  int start = 0, end = 0, dst = 0;
  if (!p.live_items && item < p.n_items) {
    start = __ldg(p.item_start + it);
    end = __ldg(p.item_end + it);
    dst = __ldg(p.item_dst + it);
  }
  // the work plan above is static; everything below may read what the previous kernel of the step wrote
  pdl_trigger();
  pdl_wait();
  if (p.live_items) {  // row-restricted call: the grid covers an upper bound of the list, most of it a few blocks
    it = (item < *p.live_count) ? p.live_items[item] : -1;
    if (it >= 0) {
      start = __ldg(p.item_start + it);
      end = __ldg(p.item_end + it);
      dst = __ldg(p.item_dst + it);
    } else {
      item = p.n_items;
    }
  }
  if (item < p.n_items && p.dst_flags && !ldc_u8(p.dst_flags + __ldg(p.item_row + it))) end = start;  // row not needed
  const bool hub = dst < 0;
  const int id = hub ? ~dst : dst;  // row, or slot of a hub piece
  const bool live = end > start;
  int maxlen = end - start;
  if (GPW > 1) {
#pragma unroll
    for (int o = G; o < 32; o <<= 1) maxlen = max(maxlen, __shfl_xor_sync(0xffffffffu, maxlen, o));
  }
  const int cfirst = live ? __ldg(p.colidx + start) : 0;  // padding target: a row this sum reads anyway (HINT: flag bit kept)
  uint64_t pol_hot = 0, pol_cold = 0;
  if (HINT) { pol_hot = make_policy(2); pol_cold = make_policy(p.cold_evict_first ? 1 : 0); }
  float4 acc[VPL];
#pragma unroll
  for (int t = 0; t < VPL; ++t) acc[t] = make_float4(0.f, 0.f, 0.f, 0.f);
  const float* xg = p.x + (size_t)gl * 4;

  int c_nx[EPL];
  float v_nx[EPL];
  if (!COMPACT && B200REC_SPMM_PREFETCH) {
#pragma unroll
    for (int e = 0; e < EPL; ++e) {
      const int k = start + e * G + gl;
      c_nx[e] = cfirst;
      v_nx[e] = 0.f;
      if (k < end) {
        c_nx[e] = ld_stream_i32(p.colidx + k);
        v_nx[e] = HAS_VALS ? ld_stream_f32(p.vals + k) : 1.f;
        if (HAS_NBR) v_nx[e] *= p.nbr_scale[c_nx[e]];
      }
    }
  }
  for (int base = 0; base < maxlen; base += EB) {
    int cnt = 0;  // contributing edges of this group in this block of EB; cntmax: warp-uniform loop bound
    if (COMPACT) {
#pragma unroll
      for (int e = 0; e < EPL; ++e) my_cv[e * G + gl] = make_int2(cfirst, 0);  // pad
      __syncwarp();
    }
    if (!COMPACT && B200REC_SPMM_PREFETCH) {
      // software pipeline: the (column, value) pairs of this block were loaded while the previous block's rows were being
      // gathered; park them and start the loads of the next block before this block's gathers
#pragma unroll
      for (int e = 0; e < EPL; ++e) my_cv[e * G + gl] = make_int2(c_nx[e], __float_as_int(v_nx[e]));
#pragma unroll
      for (int e = 0; e < EPL; ++e) {
        const int k = start + base + EB + e * G + gl;
        c_nx[e] = cfirst;
        v_nx[e] = 0.f;
        if (k < end) {
          c_nx[e] = ld_stream_i32(p.colidx + k);
          v_nx[e] = HAS_VALS ? ld_stream_f32(p.vals + k) : 1.f;
          if (HAS_NBR) v_nx[e] *= p.nbr_scale[c_nx[e]];
        }
      }
    } else {
#pragma unroll
      for (int e = 0; e < EPL; ++e) {
        const int k = start + base + e * G + gl;
        int c = cfirst;
        float v = 0.f;
        bool valid = k < end;
        if (valid) {
          c = ld_stream_i32(p.colidx + k);
          v = HAS_VALS ? ld_stream_f32(p.vals + k) : 1.f;
          if (HAS_MASK) {
            const int eidx = HAS_EID ? ld_stream_i32(p.eid + k) : k;
            valid = ((*(p.keep_bits + (eidx >> 5)) >> (eidx & 31)) & 1u) != 0u;
          }
          if (HAS_SRCF) valid = valid && (ldc_u8(p.src_flags + c) != 0);
          if (HAS_NBR && valid) v *= p.nbr_scale[c];
        }
        if (COMPACT) {
          const unsigned bal = __ballot_sync(0xffffffffu, valid);
          const unsigned gm = (G == 32) ? bal : ((bal >> (lane & ~(G - 1))) & ((1u << (G & 31)) - 1u));
          const int rank = cnt + __popc(gm & ((1u << gl) - 1u));
          if (valid) my_cv[rank] = make_int2(c, __float_as_int(v));
          cnt += __popc(gm);
        } else {
          my_cv[e * G + gl] = valid ? make_int2(c, __float_as_int(v)) : make_int2(cfirst, 0);
        }
      }
    }
    if (!COMPACT) cnt = min(EB, max(end - start - base, 0));
    int cntmax = cnt;
    if (GPW > 1) {
#pragma unroll
      for (int o = G; o < 32; o <<= 1) cntmax = max(cntmax, __shfl_xor_sync(0xffffffffu, cntmax, o));
    }
    __syncwarp();
#pragma unroll
    for (int j0 = 0; j0 < EB; j0 += U) {
      if (j0 >= cntmax) break;  // warp-uniform
      float4 xv[U][VPL];
      float vv[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int2 cv = my_cv[j0 + u];  // broadcast LDS.64
        vv[u] = __int_as_float(cv.y);
        if (HINT) {
          const float* r = xg + (size_t)(((unsigned)cv.x & 0x7fffffffu) * (unsigned)D);
          const uint64_t pol = (cv.x < 0) ? pol_hot : pol_cold;
#pragma unroll
          for (int t = 0; t < VPL; ++t) xv[u][t] = ldc_f4_policy(r + t * G * 4, pol);
        } else {
          const float* r = xg + (size_t)((unsigned)cv.x * (unsigned)D);
#pragma unroll
          for (int t = 0; t < VPL; ++t) xv[u][t] = ldc_f4(r + t * G * 4);
        }
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
#pragma unroll
        for (int t = 0; t < VPL; ++t) {
          acc[t].x = fmaf(vv[u], xv[u][t].x, acc[t].x);
          acc[t].y = fmaf(vv[u], xv[u][t].y, acc[t].y);
          acc[t].z = fmaf(vv[u], xv[u][t].z, acc[t].z);
          acc[t].w = fmaf(vv[u], xv[u][t].w, acc[t].w);
        }
      }
    }
    __syncwarp();  // slots are rewritten by the next block
  }
  if (item >= p.n_items) return;
  if (p.dst_flags && !ldc_u8(p.dst_flags + __ldg(p.item_row + it))) return;
  if (!hub) {
    epilogue_row<G, VPL>(p, id, gl, acc);
    return;
  }
  // A chunk of a split (hub) row: park the raw partial sum; the group that parks the LAST chunk of the row adds all
  // chunk sums in slot order (so the result does not depend on which group finishes last) and runs the epilogue.
  const int slot = id;
#pragma unroll
  for (int t = 0; t < VPL; ++t) __stcg(reinterpret_cast<float4*>(p.partial + (size_t)slot * D + (size_t)(gl + t * G) * 4), acc[t]);
  const int li = __ldg(p.slot_long + slot);
  const int ns = __ldg(p.long_nslot + li);
  __threadfence();
  __syncwarp();
  int old = 0;
  if (gl == 0) old = atomicAdd(p.long_cnt + li, 1);
  old = __shfl_sync(0xffffffffu, old, 0, G);
  if (old != ns - 1) return;
  if (gl == 0) p.long_cnt[li] = 0;
  __threadfence();
  const float* pbase = p.partial + (size_t)__ldg(p.long_slot0 + li) * D + (size_t)gl * 4;
#pragma unroll
  for (int t = 0; t < VPL; ++t) acc[t] = make_float4(0.f, 0.f, 0.f, 0.f);
  int sidx = 0;
  for (; sidx + 8 <= ns; sidx += 8) {
    float4 v8[8][VPL];
#pragma unroll
    for (int q = 0; q < 8; ++q)
#pragma unroll
      for (int t = 0; t < VPL; ++t) v8[q][t] = __ldcg(reinterpret_cast<const float4*>(pbase + (size_t)(sidx + q) * D + t * G * 4));
#pragma unroll
    for (int q = 0; q < 8; ++q)
#pragma unroll
      for (int t = 0; t < VPL; ++t) {
        acc[t].x += v8[q][t].x; acc[t].y += v8[q][t].y; acc[t].z += v8[q][t].z; acc[t].w += v8[q][t].w;
      }
  }
  for (; sidx < ns; ++sidx) {
#pragma unroll
    for (int t = 0; t < VPL; ++t) {
      const float4 v1 = __ldcg(reinterpret_cast<const float4*>(pbase + (size_t)sidx * D + t * G * 4));
      acc[t].x += v1.x; acc[t].y += v1.y; acc[t].z += v1.z; acc[t].w += v1.w;
    }
  }
  epilogue_row<G, VPL>(p, __ldg(p.long_row + li), gl, acc);
}

template <int G, int VPL>
static int launch_spmm(const b200rec_csr* a, const SpmmParams& p, cudaStream_t st, int max_live) {
  constexpr int GPW = 32 / G;
  // threads per block: 256 measured best or equal on every shape (C2: 57-58 us at 32/64/256 for D=64; C4 D=16:
  // 1.86 ms at 256 vs 4.9 ms at 32); B200REC_SPMM_TPB overrides for experiments
  static const int tpb = [] {
    const char* e = getenv("B200REC_SPMM_TPB");
    const int v = e ? atoi(e) : 256;
    return (v == 32 || v == 64 || v == 128 || v == 256) ? v : 256;
  }();
  const int items_per_block = (tpb / 32) * GPW;
  if (a->n_items > 0) {
    const int grid = ceil_div(p.live_items ? min(max_live, a->n_items) : a->n_items, items_per_block);
    const bool hv = p.vals != nullptr, hn = p.nbr_scale != nullptr, hm = p.keep_bits != nullptr, he = p.eid != nullptr && hm;
    const bool hs = p.src_flags != nullptr;
    if (a->col_hint) {  // hot-row residency hints: plain valued operand only
      if (hv && !hn && !hm && !hs) {
        B2_LAUNCH_PDL(spmm_items_kernel<G, VPL, true, false, false, false, false, true>, grid, tpb, 0, st, p);
        return 0;
      }
      return fail(B200REC_ERR_UNSUPPORTED, "%s: %s", "b200rec_spmm_f32", "col_hint operands take no masks or scales");
    }
#define B2_SPMM_CASE(V, N, M, E, S)                                                      \
  if (hv == V && hn == N && hm == M && he == E && hs == S) {                             \
    B2_LAUNCH_PDL(spmm_items_kernel<G, VPL, V, N, M, E, S, false>, grid, tpb, 0, st, p); \
  } else
    B2_SPMM_CASE(true, false, false, false, false)   // normalised adjacency
    B2_SPMM_CASE(true, false, false, false, true)    // normalised adjacency, sparse source (first backward hop)
    B2_SPMM_CASE(false, false, false, false, false)  // template features, eval
    B2_SPMM_CASE(false, false, true, false, false)   // template features, training (edge dropout)
    B2_SPMM_CASE(false, true, false, false, false)   // transposed features (backward), eval-mode graph
    B2_SPMM_CASE(false, true, true, true, false)     // transposed features (backward), dropout mask by edge id
    B2_SPMM_CASE(true, false, true, false, false)    // valued operand with dropout
    B2_SPMM_CASE(true, true, false, false, false)
    { return fail(B200REC_ERR_UNSUPPORTED, "%s: %s", "b200rec_spmm_f32", "operand combination (vals/nbr_scale/keep_bits/eid/src_flags) not instantiated"); }
#undef B2_SPMM_CASE
  }
  return 0;
}

static int spmm_dispatch(const b200rec_csr* a, const float* x, int d, const uint32_t* keep_bits, float post_scale,
                         float* y, const float* addend, float* out, float out_scale, const uint8_t* dst_flags,
                         const uint8_t* src_flags, cudaStream_t st, int n_peer = 0, float* const* peer_y = nullptr,
                         float* const* peer_out = nullptr, const int32_t* live_items = nullptr,
                         const int32_t* live_count = nullptr, int max_live = 0, const uint8_t* out_rows = nullptr,
                         int out_mode = 0) {
  B2_REQUIRE(a && x, "null operand");
  B2_REQUIRE(!out_rows || out_mode == 1 || out_mode == 2, "out_rows needs out_mode 1 or 2");
  B2_REQUIRE(y || out, "no output");
  B2_REQUIRE(a->n_items == 0 || (a->item_start && a->item_end && a->item_dst && a->colidx), "csr plan missing");
  B2_REQUIRE(a->n_long == 0 || (a->partial && a->long_row && a->long_slot0 && a->long_nslot && a->slot_long && a->long_cnt),
             "long-row plan missing");
  B2_REQUIRE(!dst_flags || a->item_row, "dst_flags needs item_row in the plan");
  B2_REQUIRE((long long)a->n_cols * d < (1ll << 32), "gathered table must have < 2^32 elements");
  B2_REQUIRE(a->n_passes <= 1, "column-blocked operand: call b200rec_spmm_f32_blocked");
  SpmmParams p;
  p.colidx = a->colidx; p.vals = a->vals; p.nbr_scale = a->nbr_scale; p.row_scale = a->row_scale; p.eid = a->eid;
  p.keep_bits = keep_bits; p.dst_flags = dst_flags; p.src_flags = src_flags;
  p.live_items = live_items; p.live_count = live_count;
  p.out_rows = out_rows; p.out_mode = out_rows ? out_mode : 0;
  B2_REQUIRE(!live_items || (live_count && max_live > 0), "live_items needs live_count and max_live");
  p.item_start = a->item_start; p.item_end = a->item_end; p.item_dst = a->item_dst; p.item_row = a->item_row;
  p.n_items = a->n_items;
  p.n_peer = n_peer;
  for (int q = 0; q < 8; ++q) {
    p.peer_y[q] = (q < n_peer && peer_y) ? peer_y[q] : nullptr;
    p.peer_out[q] = (q < n_peer && peer_out) ? peer_out[q] : nullptr;
  }
  p.x = x; p.y = y; p.addend = addend; p.out = out; p.out_scale = out_scale; p.post_scale = post_scale;
  p.cold_evict_first = (a->col_hint == 2);
  p.partial = a->partial; p.slot_long = a->slot_long; p.long_row = a->long_row; p.long_slot0 = a->long_slot0;
  p.long_nslot = a->long_nslot; p.long_cnt = a->long_cnt;
  switch (d) {
    case 8: return launch_spmm<2, 1>(a, p, st, max_live);
    case 16: return launch_spmm<4, 1>(a, p, st, max_live);
    case 32: return launch_spmm<8, 1>(a, p, st, max_live);
    case 64: return launch_spmm<16, 1>(a, p, st, max_live);
    case 128: return launch_spmm<32, 1>(a, p, st, max_live);
    case 256: return launch_spmm<32, 2>(a, p, st, max_live);
    default: return fail(B200REC_ERR_UNSUPPORTED, "%s: %s", "b200rec_spmm_f32", "embedding size must be 8/16/32/64/128/256");
  }
}

// ---- compacted work list for a row-restricted call -----------------------------------------------------------------
// A BPR step needs its last forward layer only at the <= 3B sampled rows: of the 72 103 work items of the C2 plan ~8 000
// are live, and 4 507 blocks that mostly load three plan entries and a flag and exit cost ~15 us.  The list is built
// beside the first forward layers (side stream); the order of its entries only changes the schedule, never a sum.
__global__ void live_items_kernel(const int32_t* __restrict__ item_row, int n_items, const uint8_t* row_flags,
                                  int32_t* __restrict__ live_items, int32_t* live_count) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const bool live = i < n_items && ldc_u8(row_flags + __ldg(item_row + i)) != 0;
  const unsigned bal = __ballot_sync(0xffffffffu, live);
  if (bal == 0) return;
  const int lane = threadIdx.x & 31;
  int base = 0;
  if (lane == 0) base = atomicAdd(live_count, __popc(bal));
  base = __shfl_sync(0xffffffffu, base, 0);
  if (live) live_items[base + __popc(bal & ((1u << lane) - 1u))] = i;
}

// ---- adjacency normalisation (model.py:89-98) ----------------------------------------------------------
__global__ void adj_dinv_kernel(const int32_t* rowptr, const float* mult, int n_rows, float* dinv) {
  const int warp = (int)((blockIdx.x * (unsigned)blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
  if (warp >= n_rows) return;
  const int s = rowptr[warp], e = rowptr[warp + 1];
  float deg;
  if (mult) {
    // np.sum(adj, axis=1) accumulates the fp32 multiplicities; they are small integers, so any order is exact
    float acc = 0.f;
    for (int k = s + lane; k < e; k += 32) acc += mult[k];
    deg = warp_sum(acc);
  } else {
    deg = (float)(e - s);
  }
  deg = fmaxf(1.f, deg);
  if (lane == 0) dinv[warp] = (float)(1.0 / sqrt((double)deg));  // deg^-0.5 correctly rounded to fp32
}
__global__ void adj_vals_kernel(const int32_t* rowptr, const int32_t* colidx, const float* mult, int n_rows,
                                const float* dinv, float* vals) {
  const int warp = (int)((blockIdx.x * (unsigned)blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
  if (warp >= n_rows) return;
  const int s = rowptr[warp], e = rowptr[warp + 1];
  const float dr = dinv[warp];
  for (int k = s + lane; k < e; k += 32) {
    const float m = mult ? mult[k] : 1.f;
    vals[k] = __fmul_rn(__fmul_rn(dr, m), dinv[colidx[k]]);  // d_mat.dot(adj).dot(d_mat): (d_r * a) * d_c
  }
}

}  // namespace b200rec

using namespace b200rec;

extern "C" int b200rec_l2_persist(int64_t bytes, int64_t* granted_out) {
  int dev = 0, max_persist = 0;
  B2_CUDA(cudaGetDevice(&dev));
  B2_CUDA(cudaDeviceGetAttribute(&max_persist, cudaDevAttrMaxPersistingL2CacheSize, dev));
  size_t want = bytes < 0 ? 0 : (size_t)bytes;
  if (want > (size_t)max_persist) want = (size_t)max_persist;
  B2_CUDA(cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, want));
  size_t got = 0;
  B2_CUDA(cudaDeviceGetLimit(&got, cudaLimitPersistingL2CacheSize));
  if (granted_out) *granted_out = (int64_t)got;
  return 0;
}

extern "C" int b200rec_adj_normalize(const int32_t* rowptr, const int32_t* colidx, const float* mult, int32_t n_rows,
                                     float* dinv, float* vals, void* stream) {
  B2_REQUIRE(rowptr && colidx && dinv && vals && n_rows > 0, "null argument");
  cudaStream_t st = (cudaStream_t)stream;
  const int grid = ceil_div((long long)n_rows * 32, 256);
  adj_dinv_kernel<<<grid, 256, 0, st>>>(rowptr, mult, n_rows, dinv);
  B2_LAUNCHED();
  adj_vals_kernel<<<grid, 256, 0, st>>>(rowptr, colidx, mult, n_rows, dinv, vals);
  B2_LAUNCHED();
  return 0;
}

extern "C" int b200rec_spmm_f32(const b200rec_csr* a, const float* x, int32_t d, const uint32_t* keep_bits,
                                float post_scale, float* y, const float* addend, float* out, float out_scale,
                                void* stream) {
  return spmm_dispatch(a, x, d, keep_bits, post_scale, y, addend, out, out_scale, nullptr, nullptr, (cudaStream_t)stream);
}

extern "C" int b200rec_spmm_f32_ex(const b200rec_csr* a, const float* x, int32_t d, const uint32_t* keep_bits,
                                   float post_scale, float* y, const float* addend, float* out, float out_scale,
                                   const uint8_t* dst_flags, const uint8_t* src_flags, void* stream) {
  return spmm_dispatch(a, x, d, keep_bits, post_scale, y, addend, out, out_scale, dst_flags, src_flags,
                       (cudaStream_t)stream);
}

extern "C" int b200rec_spmm_f32_sel(const b200rec_csr* a, const float* x, int32_t d, float post_scale, float* y,
                                    const float* addend, float* out, float out_scale, const uint8_t* dst_flags,
                                    const uint8_t* src_flags, const uint8_t* out_rows, int32_t out_mode, void* stream) {
  return spmm_dispatch(a, x, d, nullptr, post_scale, y, addend, out, out_scale, dst_flags, src_flags, (cudaStream_t)stream,
                       0, nullptr, nullptr, nullptr, nullptr, 0, out_rows, out_mode);
}

extern "C" int b200rec_spmm_f32_peer(const b200rec_csr* a, const float* x, int32_t d, const uint32_t* keep_bits,
                                     float post_scale, float* y, const float* addend, float* out, float out_scale,
                                     const uint8_t* dst_flags, const uint8_t* src_flags, int32_t n_peers,
                                     float* const* peer_y /*HOST [n_peers] or NULL*/, float* const* peer_out /*HOST or NULL*/,
                                     const uint8_t* out_rows, int32_t out_mode, void* stream) {
  B2_REQUIRE(n_peers >= 0 && n_peers <= 8, "at most 8 peers");
  return spmm_dispatch(a, x, d, keep_bits, post_scale, y, addend, out, out_scale, dst_flags, src_flags, (cudaStream_t)stream,
                       n_peers, peer_y, peer_out, nullptr, nullptr, 0, out_rows, out_mode);
}

extern "C" int b200rec_live_items(const b200rec_csr* a, const uint8_t* row_flags, int32_t* live_items, int32_t* live_count,
                                  void* stream) {
  B2_REQUIRE(a && row_flags && live_items && live_count && a->item_row, "null argument");
  cudaStream_t st = (cudaStream_t)stream;
  B2_CUDA(cudaMemsetAsync(live_count, 0, sizeof(int32_t), st));
  if (a->n_items > 0) {
    live_items_kernel<<<ceil_div(a->n_items, 256), 256, 0, st>>>(a->item_row, a->n_items, row_flags, live_items, live_count);
    B2_LAUNCHED();
  }
  return 0;
}

extern "C" int b200rec_spmm_f32_live(const b200rec_csr* a, const float* x, int32_t d, float post_scale, float* y,
                                     const float* addend, float* out, float out_scale, const int32_t* live_items,
                                     const int32_t* live_count, int32_t max_live, void* stream) {
  B2_REQUIRE(live_items && live_count && max_live > 0, "live list missing");
  return spmm_dispatch(a, x, d, nullptr, post_scale, y, addend, out, out_scale, nullptr, nullptr, (cudaStream_t)stream, 0,
                       nullptr, nullptr, live_items, live_count, max_live);
}

extern "C" int b200rec_propagate_fwd(const b200rec_csr* a, const float* x0, int32_t d, int32_t n_layers, float* buf0,
                                     float* buf1, float* mean_out, const uint8_t* needed_rows, void* stream) {
  B2_REQUIRE(a && x0 && mean_out && n_layers >= 0, "null argument");
  cudaStream_t st = (cudaStream_t)stream;
  if (n_layers == 0) {
    B2_CUDA(cudaMemcpyAsync(mean_out, x0, (size_t)a->n_rows * d * sizeof(float), cudaMemcpyDeviceToDevice, st));
    return 0;
  }
  B2_REQUIRE(n_layers == 1 || buf0, "buf0 required for n_layers >= 2");
  B2_REQUIRE(n_layers <= 2 || buf1, "buf1 required for n_layers >= 3");
  float* bufs[2] = {buf0, buf1};
  const float inv = 1.f / (float)(n_layers + 1);
  const float* src = x0;
  for (int k = 0; k < n_layers; ++k) {
    const bool last = (k == n_layers - 1);
    float* y = last ? nullptr : bufs[k & 1];
    const float* addend = (k == 0) ? x0 : mean_out;  // running layer sum lives in mean_out
    // only the last layer can be restricted: every earlier layer feeds rows that the needed rows read
    int rc = spmm_dispatch(a, src, d, nullptr, 1.f, y, addend, mean_out, last ? inv : 1.f, last ? needed_rows : nullptr,
                           nullptr, st);
    if (rc) return rc;
    src = y;
  }
  return 0;
}

extern "C" int b200rec_propagate_bwd(const b200rec_csr* a, const float* g, int32_t d, int32_t n_layers, float* buf0,
                                     float* buf1, float* dx0_out, const uint8_t* nonzero_rows, void* stream) {
  B2_REQUIRE(a && g && dx0_out && n_layers >= 0, "null argument");
  cudaStream_t st = (cudaStream_t)stream;
  const float inv = 1.f / (float)(n_layers + 1);
  if (n_layers == 0) {
    B2_CUDA(cudaMemcpyAsync(dx0_out, g, (size_t)a->n_rows * d * sizeof(float), cudaMemcpyDeviceToDevice, st));
    return 0;
  }
  B2_REQUIRE(n_layers == 1 || buf0, "buf0 required for n_layers >= 2");
  B2_REQUIRE(n_layers <= 2 || buf1, "buf1 required for n_layers >= 3");
  float* bufs[2] = {buf0, buf1};
  const float* src = g;  // H_0 = G
  for (int k = 1; k <= n_layers; ++k) {
    const bool last = (k == n_layers);
    float* dstp = last ? dx0_out : bufs[(k - 1) & 1];
    // H_k = G + A H_{k-1}; on the first hop H_0 = G is zero outside nonzero_rows, so those gathers are skipped
    int rc = spmm_dispatch(a, src, d, nullptr, 1.f, nullptr, g, dstp, last ? inv : 1.f, nullptr,
                           (k == 1) ? nonzero_rows : nullptr, st);
    if (rc) return rc;
    src = dstp;
  }
  return 0;
}
