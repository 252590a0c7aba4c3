// Column-blocked SpMM for gathered tables that do not fit L2 (b200rec_spmm_f32_blocked; C4: 3M x 128 fp32 = 1.5 GB
// against 126 MB of L2, and its 16-column shards on 8 GPUs: 192 MB).
//
// Why: random 512 B row gathers come out of L2 at ~17.5 TB/s while the table fits (<= ~80 MB) and at ~7.5 TB/s when it is
// 1.5 GB (profiles/r02_l2_gather_ceiling.txt) -- the single-pass kernel already sits on that second number (11.9 ms per C4
// layer, HBM at 81 % of peak moving 16x the algorithmic bytes).  So the source rows are cut into L2-sized blocks and the
// product runs as one PASS per block: every gather of a pass hits L2.
//
// How, without changing a bit of the result: CSR columns ascend inside a row, so a pass owns one contiguous slice
// ("segment") of every row it touches, and a segment continues the row's fmaf chain from the running sum the previous
// pass parked in `carry`.  The first attempt launched one work item per segment (round-2 sweep, profiles/
// r02_spmm_block_sweep.txt): 12.5-34 ms per layer against 11.9 single-pass -- segments average ~5 edges, and a lane group
// paid three dependent memory round trips (plan entry -> indices -> gathers) for each of them.  This kernel removes the
// per-segment latency instead of the per-segment work: the operand is re-laid out PASS-MAJOR as one record stream
//     [carry-in row r, 1.0] [col, val] [col, val] ... [end of segment: park / finish row r]   (8 B per record)
// and a lane group walks a fixed-size WINDOW of that stream, whatever segments it holds: records are fetched 32+ at a
// time, coalesced, one block ahead; a carry-in is just one more gathered row with weight 1.0 (fmaf(1, c, 0) == c, so the
// chain restarts exactly where it was parked); an end record stores the running sum (park) or runs the epilogue (finish)
// and zeroes the accumulator.  Windows hold whole segments and equal record counts, so the load is balanced by
// construction; hub rows keep their chunk pieces (each piece is a virtual row parked in `partial`, combined in slot order
// by whichever group finishes the row's last piece) -- the summation order is exactly the single-pass kernel's.
#include <cub/device/device_scan.cuh>
#include "common.cuh"

#ifndef B200REC_STREAM_BLOCKS
#define B200REC_STREAM_BLOCKS 3  // resident 256-thread blocks per SM: 3 -> 80 registers, no spills (4 -> 64 registers spills ~60 B)
#endif

namespace b200rec {

// record.x: bits 31..30 type, bit 29 hub (id is a `partial` slot instead of a row), bits 28..0 column / row / slot
constexpr unsigned REC_GATHER = 0u, REC_CARRY = 1u, REC_PARK = 2u, REC_FINISH = 3u;
constexpr unsigned REC_ID_MASK = 0x1fffffffu;

struct StreamParams {
  const int2* rec;
  const int32_t* win_start;  // [n_windows_total + 1] record offsets; window w = [win_start[w], win_start[w+1])
  int win_base, n_windows;   // this launch = windows [win_base, win_base + n_windows) (one pass)
  int32_t* win_counter;      // this pass's "next window" counter (zero at launch)
  const float* x;
  float* y;
  const float* addend;
  float* out;
  float out_scale, post_scale;
  const float* row_scale;
  float* carry;
  int ld;                    // row stride (floats) of x, y, addend, out, carry
  float* partial;
  const int32_t *slot_long, *long_row, *long_slot0, *long_nslot;
  int32_t* long_cnt;
};

template <int G, int VPL>
__device__ __forceinline__ void stream_epilogue(const StreamParams& p, int row, int gl, const float4 (&acc)[VPL]) {
  float sc = p.post_scale;
  if (p.row_scale) sc *= p.row_scale[row];
#pragma unroll
  for (int t = 0; t < VPL; ++t) {
    const size_t off = (size_t)row * (size_t)p.ld + (size_t)(gl + t * G) * 4;
    const float4 s = make_float4(acc[t].x * sc, acc[t].y * sc, acc[t].z * sc, acc[t].w * sc);
    if (p.y) st_cs_f4(p.y + off, s);
    if (p.out) {
      float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
      if (p.addend) a = ld_cs_f4(p.addend + off);
      const float os = p.out_scale;
      st_cs_f4(p.out + off, make_float4((a.x + s.x) * os, (a.y + s.y) * os, (a.z + s.z) * os, (a.w + s.w) * os));
    }
  }
}

// the last piece of a hub row: add the row's pieces in slot order (same order as the single-pass kernel) and finish it
template <int G, int VPL>
__device__ __forceinline__ void stream_finish_hub(const StreamParams& p, int slot, int gl, float4 (&acc)[VPL]) {
  constexpr int D = G * VPL * 4;
#pragma unroll
  for (int t = 0; t < VPL; ++t) __stcg(reinterpret_cast<float4*>(p.partial + (size_t)slot * D + (size_t)(gl + t * G) * 4), acc[t]);
  const int li = __ldg(p.slot_long + slot);
  const int ns = __ldg(p.long_nslot + li);
  __threadfence();
  int old = 0;
  if (gl == 0) old = atomicAdd(p.long_cnt + li, 1);
  // the G lanes of this group are converged here (the caller's branch is uniform inside a group)
  const unsigned gmask = (G == 32) ? 0xffffffffu : (((1u << G) - 1u) << ((threadIdx.x & 31) & ~(G - 1)));
  old = __shfl_sync(gmask, old, 0, G);
  if (old != ns - 1) return;
  if (gl == 0) p.long_cnt[li] = 0;
  __threadfence();
  const float* pbase = p.partial + (size_t)__ldg(p.long_slot0 + li) * D + (size_t)gl * 4;
  float4 tot[VPL];
#pragma unroll
  for (int t = 0; t < VPL; ++t) tot[t] = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int sidx = 0; sidx < ns; ++sidx) {
#pragma unroll
    for (int t = 0; t < VPL; ++t) {
      const float4 v1 = __ldcg(reinterpret_cast<const float4*>(pbase + (size_t)sidx * D + t * G * 4));
      tot[t].x += v1.x; tot[t].y += v1.y; tot[t].z += v1.z; tot[t].w += v1.w;
    }
  }
  stream_epilogue<G, VPL>(p, __ldg(p.long_row + li), gl, tot);
}

// finishing a row or a hub piece (once per row per layer; parks are handled inline).  Out of line: the record loop is
// unrolled 8x and this body (epilogue loads and stores, the hub hand-over) would be replicated in every slot -- the first
// build of this kernel was > 64 KB of SASS and ran at a quarter of the single-pass kernel's speed out of the
// instruction cache.  `p` is a __grid_constant__ kernel parameter, so passing it by reference costs no local copy.
template <int G, int VPL>
__device__ __noinline__ void stream_finish(const StreamParams& p, unsigned code, int gl, float4 a0, float4 a1) {
  float4 acc[VPL];
  acc[0] = a0;
  if (VPL > 1) acc[VPL - 1] = a1;
  const int id = (int)(code & REC_ID_MASK);
  if (!(code & (1u << 29))) stream_epilogue<G, VPL>(p, id, gl, acc);
  else stream_finish_hub<G, VPL>(p, id, gl, acc);
}

__device__ __forceinline__ void prefetch_l2(const void* ptr) { asm volatile("prefetch.global.L2 [%0];" ::"l"(ptr)); }

// Persistent grid; lane groups take windows from a per-pass counter (a warp's groups finish their windows at different
// times -- a window holds whole segments, so one long segment makes it several times the nominal size -- and a static
// assignment left the other groups of the warp idle).  While a block of records is being folded in, the next block is
// already staged in registers, and the rows it will need from DRAM rather than from the L2-resident source block -- the
// carried sums of segments that continue a row, the addend rows of segments that finish one -- are prefetched into L2, so
// that a batch of eight gathers is never held up by one DRAM round trip.
template <int G, int VPL>
__global__ void __launch_bounds__(256, (VPL == 1) ? B200REC_STREAM_BLOCKS : 2) spmm_stream_kernel(const __grid_constant__ StreamParams p) {
  constexpr int D = G * VPL * 4;
  constexpr int U = 8;                      // rows in flight per lane
  constexpr int EB = (G >= 16) ? G : 16;    // records staged per group per block
  constexpr int EPL = EB / G;
  constexpr int CVS = (G == 32) ? EB : EB + 1;
  constexpr int LINES = (D * 4 + 127) / 128;  // 128-byte lines per row
  __shared__ int2 s_rec[(256 / G) * CVS];
  const int lane = threadIdx.x & 31;
  const int gl = lane & (G - 1);
  const unsigned gmask = (G == 32) ? 0xffffffffu : (((1u << G) - 1u) << (lane & ~(G - 1)));
  int2* my = s_rec + (threadIdx.x / G) * CVS;
  const size_t gl4 = (size_t)gl * 4;
  const unsigned ld = (unsigned)p.ld;
  const float* xb = p.x + gl4;
  float* cb = p.carry + gl4;
  float* pb = p.partial + gl4;
  int w_pos = 0, w_end = 0;   // records of the current window not staged yet
  bool more = true;           // the counter has not run out
  int2 nxt[EPL];
  int nxt_cnt = 0;

  auto take_window = [&]() {  // next window of this pass (possibly empty); group-uniform
    int w = 0;
    if (gl == 0) w = atomicAdd(p.win_counter, 1);
    w = __shfl_sync(gmask, w, 0, G);
    if (w < p.n_windows) {
      w_pos = __ldg(p.win_start + p.win_base + w);
      w_end = __ldg(p.win_start + p.win_base + w + 1);
    } else {
      more = false;
      w_pos = w_end = 0;
    }
  };
  auto stage_next = [&]() {   // records [w_pos, w_pos + nxt_cnt) -> registers; DRAM-side rows they name -> L2
    nxt_cnt = min(EB, w_end - w_pos);
#pragma unroll
    for (int e = 0; e < EPL; ++e) {
      const int k = e * G + gl;
      nxt[e] = (k < nxt_cnt) ? __ldcs(p.rec + w_pos + k) : make_int2(0, 0);  // padding: a gather of row 0, never folded in
    }
    w_pos += nxt_cnt;
  };
  auto prefetch_staged = [&]() {
#pragma unroll
    for (int e = 0; e < EPL; ++e) {
      const unsigned code = (unsigned)nxt[e].x;
      const unsigned type = code >> 30, id = code & REC_ID_MASK;
      if (type == REC_CARRY) {
        const float* r = (code & (1u << 29)) ? p.partial + (size_t)id * D : p.carry + (size_t)(id * ld);
#pragma unroll
        for (int q = 0; q < LINES; ++q) prefetch_l2(r + q * 32);
      } else if (type == REC_FINISH && p.addend && !(code & (1u << 29))) {
        const float* r = p.addend + (size_t)(id * ld);
#pragma unroll
        for (int q = 0; q < LINES; ++q) prefetch_l2(r + q * 32);
      }
    }
  };

  // the record stream, the windows and the counter are static / owned by this launch: fetched while the previous pass drains
  take_window();
  stage_next();
  pdl_trigger();
  pdl_wait();
  prefetch_staged();  // carried sums are the previous pass's output
  float4 acc[VPL];
#pragma unroll
  for (int t = 0; t < VPL; ++t) acc[t] = make_float4(0.f, 0.f, 0.f, 0.f);

  for (;;) {
    const int cnt = nxt_cnt;
#pragma unroll
    for (int e = 0; e < EPL; ++e) my[e * G + gl] = nxt[e];
    if (w_pos >= w_end && more) take_window();
    stage_next();
    prefetch_staged();
    int cntmax = cnt;
    bool any_more = more || nxt_cnt > 0;
    if (G < 32) {
#pragma unroll
      for (int o = G; o < 32; o <<= 1) cntmax = max(cntmax, __shfl_xor_sync(0xffffffffu, cntmax, o));
    }
    any_more = __any_sync(0xffffffffu, any_more);
    __syncwarp();
#pragma unroll 1
    for (int j0 = 0; j0 < EB; j0 += U) {  // not unrolled: code size (see stream_finish)
      if (j0 >= cntmax) break;  // warp-uniform
      float4 xv[U][VPL];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const unsigned code = (unsigned)my[j0 + u].x;  // broadcast LDS
        const unsigned type = code >> 30, id = code & REC_ID_MASK;
        if (type <= REC_CARRY) {
          const float* r = xb + (size_t)(id * ld);
          if (type == REC_CARRY) r = (code & (1u << 29)) ? pb + (size_t)id * D : cb + (size_t)(id * ld);
#pragma unroll
          for (int t = 0; t < VPL; ++t) xv[u][t] = ldc_f4(r + t * G * 4);
        }
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        if (j0 + u >= cnt) break;  // padding is fetched (row 0) but never folded in: 0 * inf would be NaN
        const int2 rc = my[j0 + u];  // re-read (broadcast LDS.64) rather than kept live across the gathers
        const unsigned code = (unsigned)rc.x;
        const unsigned type = code >> 30;
        if (type <= REC_CARRY) {
          const float w = __int_as_float(rc.y);
#pragma unroll
          for (int t = 0; t < VPL; ++t) {
            acc[t].x = fmaf(w, xv[u][t].x, acc[t].x);
            acc[t].y = fmaf(w, xv[u][t].y, acc[t].y);
            acc[t].z = fmaf(w, xv[u][t].z, acc[t].z);
            acc[t].w = fmaf(w, xv[u][t].w, acc[t].w);
          }
        } else {
          if (type == REC_PARK) {  // the row (or hub piece) continues in a later pass: park the running sum
            const unsigned id = code & REC_ID_MASK;
            float* dst = (code & (1u << 29)) ? pb + (size_t)id * D : cb + (size_t)(id * ld);
#pragma unroll
            for (int t = 0; t < VPL; ++t) st_cs_f4(dst + t * G * 4, acc[t]);
          } else {
            stream_finish<G, VPL>(p, code, gl, acc[0], acc[VPL - 1]);
          }
#pragma unroll
          for (int t = 0; t < VPL; ++t) acc[t] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
      }
    }
    __syncwarp();  // slots are rewritten by the next block
    if (!any_more) break;
  }
}

// ---------------------------------------------------------------------------------------------------- stream build
// items: the column-blocked plan of b200rec_plan_build (sorted by pass; codes with bits 29 / 30).  Record count of an
// item = its edges + 1 carry-in (if it continues a row) + 1 end record.
__global__ void stream_count_kernel(const int32_t* item_start, const int32_t* item_end, const int32_t* item_dst, int n_items,
                                    int32_t* cnt) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i > n_items) return;
  if (i == n_items) { cnt[i] = 0; return; }
  const int dst = item_dst[i];
  const unsigned code = dst < 0 ? ~(unsigned)dst : (unsigned)dst;
  cnt[i] = (item_end[i] - item_start[i]) + (int)((code >> 29) & 1u) + 1;
}
__global__ void stream_emit_kernel(const int32_t* __restrict__ colidx, const float* __restrict__ vals,
                                   const int32_t* __restrict__ item_start, const int32_t* __restrict__ item_end,
                                   const int32_t* __restrict__ item_dst, int n_items, const int32_t* __restrict__ rec_off,
                                   int2* __restrict__ rec) {
  // one warp per item: a hub piece has up to `chunk` edges, an ordinary segment a handful
  const int i = (int)((blockIdx.x * (unsigned)blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
  if (i >= n_items) return;
  const int s = item_start[i], e = item_end[i], dst = item_dst[i];
  const bool hub = dst < 0;
  const unsigned code = hub ? ~(unsigned)dst : (unsigned)dst;
  const unsigned id = code & REC_ID_MASK, hubbit = hub ? (1u << 29) : 0u;
  const bool not_first = (code >> 29) & 1u, not_last = (code >> 30) & 1u;
  int o = rec_off[i];
  if (not_first) {
    if (lane == 0) rec[o] = make_int2((int)((REC_CARRY << 30) | hubbit | id), __float_as_int(1.0f));
    ++o;
  }
  for (int k = s + lane; k < e; k += 32) rec[o + (k - s)] = make_int2(colidx[k], __float_as_int(vals[k]));
  if (lane == 0) rec[o + (e - s)] = make_int2((int)(((not_last ? REC_PARK : REC_FINISH) << 30) | hubbit | id), 0);
}
// window k of a pass starts at the first item whose record offset is >= pass_rec0 + k * window
__global__ void stream_windows_kernel(const int32_t* rec_off, int item_lo, int item_hi, int pass_rec0, int window, int n_win,
                                      int32_t* win_start) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n_win) return;
  const int target = pass_rec0 + k * window;
  int lo = item_lo, hi = item_hi;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (rec_off[mid] < target) lo = mid + 1; else hi = mid;
  }
  win_start[k] = rec_off[lo];  // rec_off[item_hi] = first record of the next pass
}

template <int G, int VPL>
static int launch_stream(const b200rec_csr* a, StreamParams p, cudaStream_t st) {
  const int groups_per_block = 256 / G;
  int dev = 0, sms = 148;
  B2_CUDA(cudaGetDevice(&dev));
  B2_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const int resident = sms * ((VPL == 1) ? B200REC_STREAM_BLOCKS : 2);  // persistent grid: one wave
  B2_CUDA(cudaMemsetAsync(a->win_counter, 0, sizeof(int32_t) * (size_t)a->n_passes, st));
  for (int pass = 0; pass < a->n_passes; ++pass) {
    p.win_base = a->pass_win_ptr[pass];
    p.n_windows = a->pass_win_ptr[pass + 1] - a->pass_win_ptr[pass];
    p.win_counter = a->win_counter + pass;
    if (p.n_windows <= 0) continue;
    B2_LAUNCH_PDL(spmm_stream_kernel<G, VPL>, min(resident, ceil_div(p.n_windows, groups_per_block)), 256, 0, st, p);
  }
  return 0;
}

}  // namespace b200rec
using namespace b200rec;

extern "C" int b200rec_stream_build(const int32_t* colidx, const float* vals, int32_t n_items, const int32_t* item_start,
                                    const int32_t* item_end, const int32_t* item_dst, const int32_t* pass_ptr,
                                    int32_t n_passes, int32_t window, int64_t* n_records_out, int32_t* n_windows_out,
                                    void* records, int32_t* win_start, int32_t* pass_win_ptr, void* stream) {
  B2_REQUIRE(colidx && vals && item_start && item_end && item_dst && pass_ptr && n_records_out && n_windows_out, "null argument");
  B2_REQUIRE(n_items > 0 && n_passes >= 1 && window >= 32, "bad size");
  cudaStream_t st = (cudaStream_t)stream;
  int32_t *cnt = nullptr, *off = nullptr;
  const size_t n1 = (size_t)n_items + 1;
  B2_CUDA(cudaMalloc(&cnt, n1 * sizeof(int32_t)));
  if (cudaMalloc(&off, n1 * sizeof(int32_t)) != cudaSuccess) { cudaFree(cnt); return fail(B200REC_ERR_CUDA, "%s: %s", __func__, "out of memory"); }
  struct Free { int32_t *a, *b; void* c = nullptr; ~Free() { cudaFree(a); cudaFree(b); if (c) cudaFree(c); } } guard{cnt, off};
  stream_count_kernel<<<ceil_div((long long)n1, 256), 256, 0, st>>>(item_start, item_end, item_dst, n_items, cnt);
  B2_LAUNCHED();
  size_t tb = 0;
  B2_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tb, cnt, off, (int)n1, st));
  B2_CUDA(cudaMalloc(&guard.c, tb ? tb : 1));
  B2_CUDA(cub::DeviceScan::ExclusiveSum(guard.c, tb, cnt, off, (int)n1, st));
  // record offsets at the pass boundaries -> windows per pass
  int32_t* pass_rec = new int32_t[(size_t)n_passes + 1];
  struct Del { int32_t* p; ~Del() { delete[] p; } } del{pass_rec};
  for (int b = 0; b <= n_passes; ++b)
    B2_CUDA(cudaMemcpyAsync(pass_rec + b, off + pass_ptr[b], sizeof(int32_t), cudaMemcpyDeviceToHost, st));
  B2_CUDA(cudaStreamSynchronize(st));
  const long long n_rec = pass_rec[n_passes];
  B2_REQUIRE(n_rec < (1ll << 31), "record stream must fit int32 offsets");
  long long n_win = 0;
  for (int b = 0; b < n_passes; ++b) n_win += ceil_div(pass_rec[b + 1] - pass_rec[b], window);
  *n_records_out = n_rec;
  *n_windows_out = (int32_t)n_win;
  if (!records) return 0;  // size query
  B2_REQUIRE(win_start && pass_win_ptr, "null output");
  stream_emit_kernel<<<ceil_div((long long)n_items * 32, 256), 256, 0, st>>>(colidx, vals, item_start, item_end, item_dst, n_items,
                                                                              off, reinterpret_cast<int2*>(records));
  B2_LAUNCHED();
  int w0 = 0;
  for (int b = 0; b < n_passes; ++b) {
    pass_win_ptr[b] = w0;
    const int nw = ceil_div(pass_rec[b + 1] - pass_rec[b], window);
    if (nw > 0) {
      stream_windows_kernel<<<ceil_div(nw, 256), 256, 0, st>>>(off, pass_ptr[b], pass_ptr[b + 1], pass_rec[b], window, nw,
                                                               win_start + w0);
      B2_LAUNCHED();
    }
    w0 += nw;
  }
  pass_win_ptr[n_passes] = w0;
  const int32_t last = (int32_t)n_rec;
  B2_CUDA(cudaMemcpyAsync(win_start + w0, &last, sizeof(int32_t), cudaMemcpyHostToDevice, st));
  B2_CUDA(cudaStreamSynchronize(st));
  return 0;
}

extern "C" int b200rec_spmm_f32_blocked(const b200rec_csr* a, const float* x, int32_t d, int32_t ld, float post_scale, float* y,
                                        const float* addend, float* out, float out_scale, float* carry, void* stream) {
  B2_REQUIRE(a && x && carry, "null operand");
  B2_REQUIRE(y || out, "no output");
  B2_REQUIRE(a->n_passes >= 1 && a->records && a->win_start && a->pass_win_ptr && a->win_counter,
             "operand has no record stream (b200rec_plan_build with col_bounds + b200rec_stream_build)");
  B2_REQUIRE(a->n_long == 0 || (a->partial && a->long_row && a->long_slot0 && a->long_nslot && a->slot_long && a->long_cnt),
             "long-row plan missing");
  if (ld == 0) ld = d;
  B2_REQUIRE(ld >= d && (ld % 4) == 0, "row stride must be >= d and a multiple of 4 floats");
  B2_REQUIRE((long long)a->n_cols * ld < (1ll << 32) && (long long)a->n_rows * ld < (1ll << 32), "tables must have < 2^32 elements");
  StreamParams p;
  p.rec = reinterpret_cast<const int2*>(a->records); p.win_start = a->win_start; p.win_base = 0; p.n_windows = 0;
  p.x = x; p.y = y; p.addend = addend; p.out = out; p.out_scale = out_scale; p.post_scale = post_scale;
  p.row_scale = a->row_scale; p.carry = carry; p.ld = ld; p.partial = a->partial; p.slot_long = a->slot_long;
  p.long_row = a->long_row; p.long_slot0 = a->long_slot0; p.long_nslot = a->long_nslot; p.long_cnt = a->long_cnt;
  cudaStream_t st = (cudaStream_t)stream;
  switch (d) {
    case 8: return launch_stream<2, 1>(a, p, st);
    case 16: return launch_stream<4, 1>(a, p, st);
    case 32: return launch_stream<8, 1>(a, p, st);
    case 64: return launch_stream<16, 1>(a, p, st);
    case 128: return launch_stream<32, 1>(a, p, st);
    case 256: return launch_stream<32, 2>(a, p, st);
    default: return fail(B200REC_ERR_UNSUPPORTED, "%s: %s", __func__, "embedding size must be 8/16/32/64/128/256");
  }
}
