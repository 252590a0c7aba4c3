// Full-rank evaluation, tensor-core path (north_star (4); model.py:122-127 predict + trainer.py:152-169 mask + topk):
// a TMA-fed tcgen05 bf16 user x item score GEMM whose epilogue never writes the score matrix.
//
//   prep kernel     : gather the user rows, convert users / items to bf16 (K-major), row norms, max item norm
//   score_tc2_kernel: one CTA per 128 users, 18 warps; warp 0 = TMA producer (item tiles, ring of STAGES), warp 1 =
//                     tcgen05.mma issuer (128 x BN x 16 UMMAs, fp32 accumulators in TMEM, two accumulator stages),
//                     warps 2-9 = drain (TMEM -> registers, group-of-8 maxima against the row's cut, hit groups pushed
//                     into shared-memory rings), warps 10-17 = consumers (per-value cut / banned-range / exclusion
//                     tests, append to the row's candidate list, re-derive the row's cut when the list grows long).
//                     The cut-off is a lower bound of the K-th best exact score (below).
//   rescore kernel  : exact fp32 scores (the same fmaf chain in d-order as the precision-0 path and the C oracle) for the
//                     candidates, final (score desc, id asc) top-K.
// Exactness: |approx_i - exact_i| <= eps_i = c_u * ||v_i||, c_u = 1.05 * 2^-7 * ||u|| (two bf16 roundings, Cauchy-Schwarz;
// fp32 accumulation error is three orders smaller); the kernels use eps_t = c_u * max ||v|| over the item's 128-item TILE
// (>= eps_i; one scalar per tile, so the test stays one compare per group of 8 scores).  L_i = approx_i - eps_t is a
// lower bound of the exact score and U_i = approx_i + eps_t an upper bound.  If T is the exact K-th best score, the K-th
// largest L over any subset of the items is <= T (those K items all have exact >= it), and every true top-K item has
// U >= T.  So with cutL = K-th largest L over the candidates kept so far, "U_i >= cutL" keeps every true top-K item:
// the ids and scores returned are identical to the exact path by construction.  The band is 2 * eps_t wide, i.e. it
// follows the norms of the items actually in the tile rather than the largest norm of the catalogue (round 1), which made
// the candidate volume -- hence the sweep time -- depend on the training state.  A row whose candidate list overflows is
// flagged and must be redone by the caller with precision 0.
#include <cuda.h>
#include <cuda_bf16.h>
#include <stdlib.h>
#include "common.cuh"

namespace b200rec {

constexpr int TC_M = 128;      // users per CTA (UMMA M, cta_group::1)
constexpr int TC_CAP = 1024;   // candidate slots per user row
struct Cand {
  float s;
  int id;
};

// ------------------------------------------------------------------------------------------------ PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  return ok != 0;
}
// bounded wait: a mis-programmed pipeline must abort the launch (trap -> CUDA error), never hang the GPU
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t spins = 0;
  while (!mbar_try(bar, parity)) {
    if (++spins > 40000000u) __trap();
  }
}
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
          smem_u32(smem_dst)),
      "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_mma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// A operand in tensor memory (row m of the 128 x K tile in lane m, two bf16 per 32-bit column, element 2c in the low half:
// the memory image of a K-major bf16 row): the user tile is the same for every item tile of the sweep, and an SS-form
// 128 x 128 x 16 UMMA reads 8 KB of shared memory per 64 cycles -- all 128 B/clk an SM has, which the record rings and
// candidate lists also need.  With A in TMEM the tensor core reads 4 KB (B only) per instruction.
__device__ __forceinline__ void tc_mma_bf16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(tmem_d),
      "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tc_st32(uint32_t taddr, const uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]),
      "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]),
      "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]),
      "r"(r[31])
      : "memory");
}
__device__ __forceinline__ void tc_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tc_ld32_nowait(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
}
// ring flags in shared memory: acquire loads are plain LDS (ordering only), a release store is one MEMBAR.ALL.CTA + STS --
// the rings used to bracket every poll with __threadfence_block() (MEMBAR.SC.CTA, ~100 cycles each, five per pass of a
// consumer warp), which cost more than the records' own processing
__device__ __forceinline__ int lds_acquire(const volatile int* p) {
  int v;
  asm volatile("ld.acquire.cta.shared.b32 %0, [%1];" : "=r"(v) : "r"(smem_u32(const_cast<const int*>(p))) : "memory");
  return v;
}
__device__ __forceinline__ void sts_release(volatile int* p, int v) {
  asm volatile("st.release.cta.shared.b32 [%0], %1;" ::"r"(smem_u32(const_cast<int*>(p))), "r"(v) : "memory");
}
// The compiler does not know that tcgen05.ld is asynchronous: tie the destination registers to a point after
// tcgen05.wait::ld so no use of them can be scheduled above the wait (no instruction is emitted)
__device__ __forceinline__ void tc_regs_fence(uint32_t (&r)[32]) {
  asm volatile("" : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]), "+r"(r[8]),
               "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15]) :: "memory");
  asm volatile("" : "+r"(r[16]), "+r"(r[17]), "+r"(r[18]), "+r"(r[19]), "+r"(r[20]), "+r"(r[21]), "+r"(r[22]), "+r"(r[23]),
               "+r"(r[24]), "+r"(r[25]), "+r"(r[26]), "+r"(r[27]), "+r"(r[28]), "+r"(r[29]), "+r"(r[30]), "+r"(r[31]) :: "memory");
}
__device__ __forceinline__ float fmax3(float a, float b, float c) {
  float r;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
  return r;
}
__device__ __forceinline__ void tc_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tc_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
// K-major operand tile with 128-byte swizzle (rows of 64 bf16 = 128 B, 8-row atoms of 1024 B): start>>4, LBO (unused) 1,
// SBO = 1024 B >> 4, descriptor version 1 (sm_100), layout type 2 = SWIZZLE_128B  (cute::UMMA::SmemDescriptor)
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t smem_addr) {
  return (uint64_t)((smem_addr >> 4) & 0x3FFF) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) |
         ((uint64_t)2 << 61);
}
// cute::UMMA::InstrDescriptor for kind::f16: D = F32, A = B = BF16, both K-major, N >> 3 at bit 17, M >> 4 at bit 24
__host__ __device__ constexpr uint32_t umma_idesc_bf16(int m, int n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

// ------------------------------------------------------------------------------------------------ prep
// rows -> bf16 (row-major, D contiguous = K-major), ||row||_2, optional gather, optional max of the norms per tile of
// `tile_rows` rows (items: the band of the candidate test is scaled by the tile's largest norm)
template <int D>
__global__ void __launch_bounds__(256) tc_prep_kernel(const float* __restrict__ table, const int64_t* __restrict__ gather,
                                                      int n_rows, __nv_bfloat16* __restrict__ out, float* __restrict__ norms,
                                                      unsigned* __restrict__ tile_max_bits, int tile_rows) {
  constexpr int G = D / 4;  // lanes per row (D = 64 -> 16, 128 -> 32)
  const int gi = (int)((blockIdx.x * (unsigned)blockDim.x + threadIdx.x) / G), gl = threadIdx.x & (G - 1);
  float ss = 0.f;
  if (gi < n_rows) {
    const int64_t src = gather ? gather[gi] : gi;
    const float4 v = ldg_f4(table + (size_t)src * D + gl * 4);
    ss = v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
    __nv_bfloat162 a = __floats2bfloat162_rn(v.x, v.y), b = __floats2bfloat162_rn(v.z, v.w);
    uint2 pk;
    pk.x = *reinterpret_cast<unsigned*>(&a);
    pk.y = *reinterpret_cast<unsigned*>(&b);
    *reinterpret_cast<uint2*>(out + (size_t)gi * D + gl * 4) = pk;
  }
  ss = group_sum<G>(ss);
  if (gi < n_rows && gl == 0) {
    const float nrm = sqrtf(ss);
    norms[gi] = nrm;
    if (tile_max_bits) atomicMax(tile_max_bits + gi / tile_rows, __float_as_uint(nrm));  // non-negative floats order like their bits
  }
}

// ------------------------------------------------------------------------------------------------ main kernel
struct TcParams {
  const int64_t* users;      // [b] absolute user ids (exclusion rows are indexed by them)
  int n_users, n_items, k;
  const int32_t *excl_ptr_a, *excl_idx_a, *excl_ptr_b, *excl_idx_b;
  int banned_lo, banned_hi;
  const float* unorm;        // [b]
  const __nv_bfloat16* ub;   // [b, D] the users' rows in bf16 (tc_prep_kernel)
  const unsigned* tile_vmax_bits; // [n_tiles] largest item norm of each BN-item tile (float bits)
  Cand* cand;                // [b, TC_CAP]
  int* cand_cnt;             // [b]
  int* overflow;             // [b]
#ifdef B200REC_TC_PROF
  unsigned long long* prof;  // [16] clock / event counters summed over the grid (debug build only)
#endif
};
#ifdef B200REC_TC_PROF
#define TCPROF_ADD(i, v) atomicAdd(p.prof + (i), (unsigned long long)(v))
#define TCPROF_CLK() clock64()
#else
#define TCPROF_ADD(i, v) ((void)0)
#define TCPROF_CLK() 0ll
#endif

// order-preserving float <-> int map (for redux.sync, which has no float form)
__device__ __forceinline__ int float_to_ord(float f) { const int b = __float_as_int(f); return b >= 0 ? b : b ^ 0x7fffffff; }
__device__ __forceinline__ float ord_to_float(int k) { return __int_as_float(k >= 0 ? k : k ^ 0x7fffffff); }
__device__ __forceinline__ bool cand_before(const Cand& a, const Cand& b) { return a.s > b.s || (a.s == b.s && a.id < b.id); }
__device__ __forceinline__ void warp_sort_desc(Cand* c, int np, int lane) {
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
__device__ __forceinline__ bool row_has(const int32_t* __restrict__ ptr, const int32_t* __restrict__ idx, int64_t row, int key) {
  if (!ptr) return false;
  int lo = __ldg(ptr + row), hi = __ldg(ptr + row + 1);
  const int end = hi;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (__ldg(idx + mid) < key) lo = mid + 1; else hi = mid;
  }
  return lo < end && __ldg(idx + lo) == key;
}

// ------------------------------------------------------------------------------------------------ main kernel
// History: ncu on the first generation (four warps doing drain + per-hit work, deleted in round 2) showed the tensor pipe 3 % busy with the MMA warp spinning on `tempty` --
// four latency-exposed warps both drained TMEM and ran the branchy per-hit work (range / exclusion tests, list appends,
// cut refinement).  Here eight DRAIN warps (2-9: two per TMEM lane quadrant, each half of a tile's columns) load their
// 64 columns with one wait, hand the accumulator stage back, reduce to eight group maxima, compare with the row's cut
// and push the hit groups (row, first item, 8 scores) of the tile as one batch into single-producer / single-consumer
// rings in shared memory (ring = quadrant x column half x row half); eight CONSUMER warps (10-17), each owning 16 rows
// and their two rings, pop the records, do the per-value tests, append to the row lists, and re-derive a row's cut when
// its list grows long.
// No locks: a row is touched by exactly one consumer; a stale cut only costs extra
// records.  The train / val exclusion test must not be a binary search per candidate: a heavy user's thousands of train
// items all score above its cut, never raise it (they are excluded) and would each cost a chain of dependent global
// loads in the one warp that owns the row.  Instead the test is a MERGE: the records of one (row, column half) arrive in
// ascending item order, so a cursor into the row's sorted exclusion list only moves forward; the 8 lanes that handle a
// record advance it 8 entries per coalesced load and compare the record's 8 consecutive items against the 8 entries
// at the cursor.  Total work per row is O(E) for the whole sweep.  (Measured alternatives: deferring the test to the
// re-score kernel with a K+E cut tripled the candidate volume, 8 -> 20 ms on the C2 sweep; a hash-bitmap prefilter left
// the heavy rows at 13 ms.)
#ifndef B200REC_TC_ACC_STAGES
#define B200REC_TC_ACC_STAGES 2
#endif
constexpr int TC_ACC = B200REC_TC_ACC_STAGES;  // accumulator stages in TMEM (x 128 columns; 4 = all 512): the MMA warp runs up
                                               // to TC_ACC - 1 tiles ahead of the drain warps, hiding the commit -> drain ->
                                               // release -> issue hand-over latency (2 stages in round 1: tensor pipe 40 %)
static_assert(TC_ACC == 2, "512 TMEM columns: TC_ACC x 128 accumulator columns + D / 2 columns of the user tile");
constexpr int TC_ACC_SHIFT = (TC_ACC == 4) ? 2 : 1;
constexpr int TC_QCAP = 64;       // records per ring (power of two); 16 rings: (quadrant, column half, row half)
constexpr int TC_RINGS = 16;
constexpr int TC_CONSUMERS = 8;    // consumer warp (quadrant, row half) owns 16 rows and their two rings
constexpr int TC2_THREADS = (2 + 8 + TC_CONSUMERS) * 32;
#ifndef B200REC_TC_REFINE_AT
#define B200REC_TC_REFINE_AT 128
#endif
constexpr int TC_REFINE_AT = B200REC_TC_REFINE_AT; // list length that triggers a cut refinement
struct __align__(16) HitRec {
  float v[8];
  int base;   // item id of v[0]
  int row;    // row inside the CTA tile
  float vmax; // largest item norm of the record's tile (the consumer's eps_t = c_u * vmax without a global load)
  int pad;
};

template <int D, int BN, int STAGES>
__global__ void __launch_bounds__(TC2_THREADS, 1) score_tc2_kernel(const __grid_constant__ CUtensorMap tm_items,
                                                                            const TcParams p) {
  constexpr int KB = D / 64;
  constexpr uint32_t B_STAGE_BYTES = BN * D * 2;
  constexpr uint32_t TMEM_A_COL = TC_ACC * BN;   // the user tile sits behind the accumulator stages
  constexpr uint32_t TMEM_COLS = 512;            // power of two >= TC_ACC * BN + D / 2
  static_assert(TC_ACC * BN + D / 2 <= 512, "TMEM budget");
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sB = smem;
  HitRec* queues = reinterpret_cast<HitRec*>(sB + (size_t)STAGES * B_STAGE_BYTES);      // [TC_RINGS][TC_QCAP]
  Cand* sort_area = reinterpret_cast<Cand*>(queues + TC_RINGS * TC_QCAP);                 // [TC_CONSUMERS][TC_CAP]
  float* s_cut = reinterpret_cast<float*>(sort_area + TC_CONSUMERS * TC_CAP);                        // [128]
  int* s_cnt = reinterpret_cast<int*>(s_cut + TC_M);                                      // [128]
  float* s_cu = reinterpret_cast<float*>(s_cnt + TC_M);                                   // [128] c_u = 1.05 * 2^-7 * ||u||
  int* s_kk = reinterpret_cast<int*>(s_cu + TC_M);                                        // [128] rank the cut is derived from
  int* s_xend = s_kk + TC_M;                                                              // [2 lists][128] end of the row's exclusion entries
  int* s_xcur = s_xend + 2 * TC_M;                                                        // [2 lists][2 halves][128] merge cursors
  int* s_xnext = s_xcur + 4 * TC_M;                                                       // [2 lists][2 halves][128] value at the cursor
  volatile int* s_tail = reinterpret_cast<volatile int*>(s_xnext + 4 * TC_M);              // [TC_RINGS] records published
  volatile int* s_head = s_tail + TC_RINGS;                                               // records consumed
  volatile int* s_done = s_head + TC_RINGS;                                               // producer finished
  uint64_t* bars = reinterpret_cast<uint64_t*>(const_cast<int*>(s_done) + TC_RINGS);
  uint64_t* full = bars;
  uint64_t* empty = bars + STAGES;
  uint64_t* tfull = bars + 2 * STAGES;
  uint64_t* tempty = tfull + TC_ACC;
  uint64_t* afull = tempty + TC_ACC;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(afull + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int u0 = blockIdx.x * TC_M;
  const int n_tiles = (p.n_items + BN - 1) / BN;

  if (threadIdx.x < TC_M) {
    const int u = u0 + (int)threadIdx.x;
    float cut0 = -INFINITY;
    int kk = p.k;
    for (int which = 0; which < 2; ++which) {
      const int32_t* ptr = which ? p.excl_ptr_b : p.excl_ptr_a;
      const int32_t* xidx = which ? p.excl_idx_b : p.excl_idx_a;
      int b = 0, e = 0;
      if (ptr && u < p.n_users) { const int64_t user = p.users[u]; b = ptr[user]; e = ptr[user + 1]; }
      const int first = (xidx && b < e) ? xidx[b] : INT32_MAX;
      s_xend[which * TC_M + threadIdx.x] = e;
      s_xcur[(which * 2 + 0) * TC_M + threadIdx.x] = b;
      s_xcur[(which * 2 + 1) * TC_M + threadIdx.x] = b;
      s_xnext[(which * 2 + 0) * TC_M + threadIdx.x] = first;
      s_xnext[(which * 2 + 1) * TC_M + threadIdx.x] = first;
    }
    s_cut[threadIdx.x] = cut0; s_cnt[threadIdx.x] = 0; s_kk[threadIdx.x] = kk;
    s_cu[threadIdx.x] = (u < p.n_users) ? 1.05f * 0.0078125f * __ldg(p.unorm + u) : 0.f;
  }
  if (threadIdx.x < TC_RINGS) { s_tail[threadIdx.x] = 0; s_head[threadIdx.x] = 0; s_done[threadIdx.x] = 0; }
  if (warp == 0 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(full + s, 1); mbar_init(empty + s, 1); }
    for (int s = 0; s < TC_ACC; ++s) { mbar_init(tfull + s, 1); mbar_init(tempty + s, 8); }
    mbar_init(afull, 4);  // the four drain warps that store the user tile into TMEM
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(TMEM_COLS));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {  // ===== TMA producer =====
      for (int t = 0; t < n_tiles; ++t) {
        const int s = t % STAGES;
        mbar_wait(empty + s, ((t / STAGES) & 1) ^ 1);
        mbar_expect_tx(full + s, B_STAGE_BYTES);
        uint8_t* dst = sB + (size_t)s * B_STAGE_BYTES;
        for (int kb = 0; kb < KB; ++kb) tma_load_2d(dst + (size_t)kb * BN * 128, &tm_items, kb * 64, t * BN, full + s);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {  // ===== MMA issuer =====
      constexpr uint32_t idesc = umma_idesc_bf16(TC_M, BN);
      mbar_wait(afull, 0);
      tc_fence_after();
      long long w_acc = 0, w_tile = 0;
      const long long mm0 = TCPROF_CLK();
      for (int t = 0; t < n_tiles; ++t) {
        const int s = t % STAGES, as = t & (TC_ACC - 1);
        const long long m0 = TCPROF_CLK();
        mbar_wait(tempty + as, ((t >> TC_ACC_SHIFT) & 1) ^ 1);
        const long long m1 = TCPROF_CLK();
        mbar_wait(full + s, (t / STAGES) & 1);
        w_acc += m1 - m0;                       // MMA warp waiting for a free accumulator stage
        w_tile += TCPROF_CLK() - m1;            // ... for an item tile
        tc_fence_after();
        const uint32_t b0 = smem_u32(sB + (size_t)s * B_STAGE_BYTES);
#pragma unroll
        for (int kb = 0; kb < KB; ++kb) {
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const uint64_t bd = umma_desc_sw128(b0 + kb * BN * 128 + k * 32);
            // K-step (kb, k) = elements 64 kb + 16 k .. + 15 of every row = 8 TMEM columns
            tc_mma_bf16_ts(tmem_base + as * BN, tmem_base + TMEM_A_COL + (uint32_t)(kb * 32 + k * 8), bd, idesc, (kb | k) ? 1u : 0u);
          }
        }
        tc_commit(empty + s);
        tc_commit(tfull + as);
      }
      TCPROF_ADD(0, w_acc);
      TCPROF_ADD(1, w_tile);
      TCPROF_ADD(14, TCPROF_CLK() - mm0);       // MMA warp: whole sweep
    }
  } else if (warp < 10) {
    // ===== drain warps: TMEM -> group maxima -> hit records.  Two warps per TMEM lane quadrant (a warp may only touch
    // lanes 32 * (warp % 4) ..), each taking half of a tile's columns in two 32-column chunks.  The chunks are software-
    // pipelined: while one chunk's maxima / records are formed, the tcgen05.ld of the next chunk (the other half of this
    // tile, or the first half of the next tile) is in flight -- TMEM reads are the unit this kernel is bound by (64 B/clk
    // per SM: 1 024 cycles per 128 x 128 fp32 tile), and a drain warp that loads, waits, then computes leaves it idle a
    // third of the time (measured 1.5 k cycles per tile at D = 128 before this).
    const int quad = warp & 3;
    const int half = (warp - 2) >> 2;
    const int lh = lane >> 4;                          // row half: lanes 0-15 feed one consumer, 16-31 the other
    const int ring = (quad * 2 + half) * 2 + lh;
    const int row = quad * 32 + lane;
    const bool active = (u0 + row) < p.n_users;
    HitRec* q = queues + ring * TC_QCAP;
    int tail = 0;
    const float cu = s_cu[row];
    static_assert(BN == 128, "drain layout: two 32-column chunks per warp");
    const uint32_t tlane = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(half * 64);

    // one 32-column chunk: group-of-8 maxima against the row's cut ("U_i >= cutL": approx_i >= cutL - c_u * largest norm in
    // the tile; cutL is re-read per chunk, the consumer may have raised it), hit groups pushed as ONE batch -- per-lane
    // record count -> prefix inside the 16-lane row half -> one wait for room, one publish.  A lane's records stay
    // consecutive and ascending in item id (what the consumer's forward-only exclusion cursor relies on); a batch holds at
    // most 16 lanes x 4 groups = TC_QCAP records, so it always fits an empty ring.
    long long d_wait_mma = 0, d_wait_ring = 0, d_batches = 0, d_records = 0;
    auto process = [&](const uint32_t (&v)[32], int t, int cc) {
      const float tvmax = __uint_as_float(__ldg(p.tile_vmax_bits + t));
      const float cut = active ? s_cut[row] - cu * tvmax : INFINITY;
      unsigned hit4 = 0;
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        // eight values in four instructions (FMNMX3, sm_100)
        const float a = fmax3(__uint_as_float(v[g * 8 + 0]), __uint_as_float(v[g * 8 + 1]), __uint_as_float(v[g * 8 + 2]));
        const float b = fmax3(__uint_as_float(v[g * 8 + 3]), __uint_as_float(v[g * 8 + 4]), __uint_as_float(v[g * 8 + 5]));
        const float c2 = fmax3(__uint_as_float(v[g * 8 + 6]), __uint_as_float(v[g * 8 + 7]), a);
        if (fmaxf(b, c2) >= cut) hit4 |= 1u << g;
      }
      if (!__any_sync(0xffffffffu, hit4 != 0)) return;
      const int mine_n = __popc(hit4);
      int pre = mine_n;  // prefix sum inside each 16-lane row half (each half has its own ring; `tail` is per half)
#pragma unroll
      for (int o = 1; o < 16; o <<= 1) {
        const int t2 = __shfl_up_sync(0xffffffffu, pre, o, 16);
        if ((lane & 15) >= o) pre += t2;
      }
      const int total = __shfl_sync(0xffffffffu, pre, 15, 16);
      pre -= mine_n;
      static_assert(16 * 4 <= TC_QCAP, "a chunk's batch must fit the ring");
      ++d_batches;
      d_records += mine_n;
      unsigned spins = 0;
      const long long r0 = TCPROF_CLK();
      while (__any_sync(0xffffffffu, tail + total - s_head[ring] > TC_QCAP)) {
        if (++spins > 200000000u) __trap();
      }
      d_wait_ring += TCPROF_CLK() - r0;  // drain warp waiting for ring room
      int k2 = tail + pre;
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        if ((hit4 >> g) & 1u) {
          HitRec* r = q + (k2 & (TC_QCAP - 1));
          *reinterpret_cast<uint4*>(&r->v[0]) = make_uint4(v[g * 8 + 0], v[g * 8 + 1], v[g * 8 + 2], v[g * 8 + 3]);
          *reinterpret_cast<uint4*>(&r->v[4]) = make_uint4(v[g * 8 + 4], v[g * 8 + 5], v[g * 8 + 6], v[g * 8 + 7]);
          r->base = t * BN + half * 64 + cc * 32 + g * 8;
          r->row = row;
          r->vmax = tvmax;
          ++k2;
        }
      }
      tail += total;
      __syncwarp();  // orders the lanes' record stores before the publishing lane's release
      if ((lane & 15) == 0) sts_release(s_tail + ring, tail);  // publish (one lane per half)
    };

    uint32_t va[32], vb[32];
    if (half == 0) {
      // the user tile -> TMEM, once per CTA: this lane's row (zeros past the last user), 32 columns (64 bf16) per store
      const uint4* src = reinterpret_cast<const uint4*>(p.ub + (size_t)(u0 + row) * D);
#pragma unroll
      for (int c = 0; c < D / 64; ++c) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const uint4 w = active ? __ldg(src + c * 8 + j) : make_uint4(0u, 0u, 0u, 0u);
          va[j * 4 + 0] = w.x; va[j * 4 + 1] = w.y; va[j * 4 + 2] = w.z; va[j * 4 + 3] = w.w;
        }
        tc_st32(tmem_base + ((uint32_t)(quad * 32) << 16) + TMEM_A_COL + (uint32_t)(c * 32), va);
      }
      tc_wait_st();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(afull);
    }
    {
      const long long d0 = TCPROF_CLK();
      mbar_wait(tfull + 0, 0);
      d_wait_mma += TCPROF_CLK() - d0;
      tc_fence_after();
      tc_ld32_nowait(tlane, va);
      tc_wait_ld();
      tc_regs_fence(va);
    }
    for (int t = 0; t < n_tiles; ++t) {
      const int as = t & (TC_ACC - 1);
      tc_ld32_nowait(tlane + (uint32_t)(as * BN + 32), vb);  // second chunk of this tile: in flight under the first's work
      process(va, t, 0);
      tc_wait_ld();
      tc_regs_fence(vb);
      // this warp has read all its columns of the tile: hand the accumulator stage back to the MMA warp
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty + as);
      if (t + 1 < n_tiles) {
        const int an = (t + 1) & (TC_ACC - 1);
        const long long d0 = TCPROF_CLK();
        mbar_wait(tfull + an, ((t + 1) >> TC_ACC_SHIFT) & 1);
        d_wait_mma += TCPROF_CLK() - d0;  // drain warp waiting for the MMA
        tc_fence_after();
        tc_ld32_nowait(tlane + (uint32_t)(an * BN), va);     // first chunk of the next tile: under the second chunk's work
      }
      process(vb, t, 1);
      if (t + 1 < n_tiles) {
        tc_wait_ld();
        tc_regs_fence(va);
      }
    }
    if (lane == 0) {
      TCPROF_ADD(2, d_wait_mma);
      TCPROF_ADD(3, d_wait_ring);
      TCPROF_ADD(5, d_batches);
    }
    TCPROF_ADD(6, d_records);
    __syncwarp();
    if ((lane & 15) == 0) sts_release(s_done + ring, 1);
  } else {
    // ===== consumer warps: warp 10 + 2q + h owns rows 32q + 16h .. +15 and their two rings (one per column half) =====
    const int cw = warp - 10;
    const int quad = cw >> 1, rh = cw & 1;
    const int row_base = quad * 32 + rh * 16;
    Cand* my_sort = sort_area + (size_t)cw * TC_CAP;
    int head2[2] = {0, 0};

    long long c_loop = 0, c_refine = 0, c_iters = 0, c_passes = 0, c_nref = 0, c_app = 0;
    const long long c_begin = TCPROF_CLK();
    auto refine_row = [&](int row) {
      // warp-cooperative: lower bound of the K'-th best approximate score by value bisection, then keep the band above
      // (bound - margin) with a ballot compaction (see the exactness argument at the top of the file)
      const int u = u0 + row;
      const int rc = min(s_cnt[row], TC_CAP);
      const int kk = s_kk[row];
      if (rc < kk) return;
      Cand* list = p.cand + (size_t)u * TC_CAP;
      const float cu2 = 2.f * s_cu[row];  // an entry holds L = approx - eps_t; its upper bound is U = L + 2 * eps_t
      const float rcut = s_cut[row];
      float mx = -INFINITY, mn = INFINITY;
      for (int t = lane; t < rc; t += 32) {
        const Cand c = list[t];
        my_sort[t] = c;
        mx = fmaxf(mx, c.s);
        mn = fminf(mn, c.s);
      }
      mx = ord_to_float(__reduce_max_sync(0xffffffffu, float_to_ord(mx)));  // one REDUX each instead of five shuffles
      mn = ord_to_float(__reduce_min_sync(0xffffffffu, float_to_ord(mn)));
      __syncwarp();
      float lo = (rcut > -INFINITY) ? rcut : mn;  // #(entries with L >= lo) >= K' always holds
      lo = fminf(lo, mx);
      float hi = mx;
      for (int itn = 0; itn < 14; ++itn) {
        const float pv = 0.5f * (lo + hi);
        if (!(pv > lo && pv < hi)) break;
        int c = 0;
        for (int t = lane; t < rc; t += 32) c += (my_sort[t].s >= pv) ? 1 : 0;
        c = __reduce_add_sync(0xffffffffu, c);
        if (c >= kk) { lo = pv; if (c <= kk + kk / 2 + 8) break; } else { hi = pv; }
      }
      const float ncut = lo;  // lower bound of the K'-th largest L: the new cutL
      int base = 0;
      for (int t0 = 0; t0 < rc; t0 += 32) {
        const int t = t0 + lane;
        bool kp = false;
        if (t < rc) {
          const Cand c = my_sort[t];
          kp = c.s + cu2 * __uint_as_float(__ldg(p.tile_vmax_bits + c.id / BN)) >= ncut;  // U_i >= cutL
        }
        const unsigned bal = __ballot_sync(0xffffffffu, kp);
        if (kp) list[base + __popc(bal & ((1u << lane) - 1u))] = my_sort[t];
        base += __popc(bal);
      }
      __syncwarp();
      if (lane == 0) { s_cnt[row] = base; s_cut[row] = ncut; }
      __syncwarp();
    };

    unsigned idle = 0;
    for (;;) {
      bool progressed = false, all_done = true;
#pragma unroll
      for (int hf = 0; hf < 2; ++hf) {
        const int ring = (quad * 2 + hf) * 2 + rh;
        HitRec* q = queues + ring * TC_QCAP;
        const int done = lds_acquire(s_done + ring);    // read BEFORE the tail: done => the tail read below is final
        const int tl_all = lds_acquire(s_tail + ring);  // acquire: the records below were written before this tail
        int head = head2[hf];
        // at most 16 records per ring per pass: a pass appends <= 2 * 16 * 8 = 256 values to one row (all records may
        // belong to the same row, e.g. a tile with a single live user), and the refinement check below runs before a list
        // that was under its threshold (<= 3/4 TC_CAP) can reach TC_CAP
        const int tl = min(tl_all, head + 16);
        const bool any = head < tl;
        const long long l0 = TCPROF_CLK();
        if (any) ++c_passes;
        while (head < tl) {
          ++c_iters;
          const int nrec = min(4, tl - head);
          const int ri = lane >> 3, j = lane & 7;   // 8 lanes per record, lane j <-> item base + j
          const unsigned gmask = 0xffu << (ri * 8);
          float sc = -INFINITY, tvmax = 0.f;
          int row = 0, base = 0;
          if (ri < nrec) {
            const HitRec* r = q + ((head + ri) & (TC_QCAP - 1));
            sc = r->v[j];
            row = r->row;
            base = r->base;
            tvmax = r->vmax;
          }
          const int item = base + j;
          // records carry approx; the list keeps L = approx - eps_t, the test is U = approx + eps_t >= cutL
          const float eps_t = s_cu[row] * tvmax;
          bool pass = ri < nrec && sc + eps_t >= s_cut[row] && item < p.n_items && !(item >= p.banned_lo && item < p.banned_hi);
          // merge test against the row's sorted exclusion lists (group-cooperative, warp-uniform control flow)
#pragma unroll
          for (int which = 0; which < 2; ++which) {
            const int32_t* idx = which ? p.excl_idx_b : p.excl_idx_a;
            if (!idx) continue;
            // the value at the row's cursor is kept in shared memory: a record whose 8 items all lie below it contains no
            // excluded item and needs no cursor move -- the common case, and it costs no global load
            int* nxp = s_xnext + (which * 2 + hf) * TC_M + row;
            const bool gneed = (__ballot_sync(0xffffffffu, pass) & gmask) != 0u  // some lane of my record still passes
                               && *nxp < base + 8;
            if (!__any_sync(0xffffffffu, gneed)) continue;
            int* curp = s_xcur + (which * 2 + hf) * TC_M + row;
            const int end = s_xend[which * TC_M + row];
            int cur = gneed ? *curp : end;
            // advance to the first entry >= base, 8 entries per step
            for (;;) {
              const int ev = (gneed && cur + j < end) ? __ldg(idx + cur + j) : INT32_MAX;
              const int nlt = __popc(__ballot_sync(0xffffffffu, ev < base) & gmask);
              cur += nlt;
              if (!__any_sync(0xffffffffu, nlt == 8)) break;
            }
            const int ev = (gneed && cur + j < end) ? __ldg(idx + cur + j) : INT32_MAX;
            unsigned bits = (ev >= base && ev < base + 8) ? (1u << (ev - base)) : 0u;
            bits |= __shfl_xor_sync(0xffffffffu, bits, 1);
            bits |= __shfl_xor_sync(0xffffffffu, bits, 2);
            bits |= __shfl_xor_sync(0xffffffffu, bits, 4);
            if ((bits >> j) & 1u) pass = false;
            if (gneed && j == 0) {  // two records of one row in the same pass: the later base wins both maxima
              atomicMax(curp, cur);
              atomicMax(nxp, ev);
            }
          }
          if (pass) {
            ++c_app;
            const int pos = atomicAdd(&s_cnt[row], 1);
            if (pos < TC_CAP) p.cand[(size_t)(u0 + row) * TC_CAP + pos] = Cand{sc - eps_t, item};
            else p.overflow[u0 + row] = 1;
          }
          head += nrec;
        }
        __syncwarp();
        c_loop += TCPROF_CLK() - l0;
        if (any) {
          head2[hf] = head;
          if (lane == 0) sts_release(s_head + ring, head);  // free the slots (the __syncwarp above ordered the lanes' reads)
          progressed = true;
        }
        if (!(done && head == tl_all)) all_done = false;
      }
      if (progressed) {
        // rows whose list grew long: tighten their cut (the list entries were written by other lanes of this warp: the
        // __syncwarp after the record loop ordered them)
        const int row = row_base + (lane & 15);
        const int kk = s_kk[row];
        const bool need = lane < 16 && (u0 + row) < p.n_users && s_cnt[row] >= max(TC_REFINE_AT, min(2 * kk, 3 * TC_CAP / 4));
        unsigned needm = __ballot_sync(0xffffffffu, need);
        const long long f0 = TCPROF_CLK();
        while (needm) {
          const int r = __ffs(needm) - 1;
          needm &= needm - 1;
          refine_row(row_base + r);
          ++c_nref;
        }
        c_refine += TCPROF_CLK() - f0;
        idle = 0;
      } else {
        if (all_done) break;
        if (++idle > 400000000u) __trap();
      }
    }
    if (lane == 0) {
      TCPROF_ADD(7, TCPROF_CLK() - c_begin);  // consumer warp: total, record loop, refinement
      TCPROF_ADD(8, c_loop);
      TCPROF_ADD(9, c_refine);
      TCPROF_ADD(10, c_iters);
      TCPROF_ADD(11, c_passes);
      TCPROF_ADD(12, c_nref);
    }
    TCPROF_ADD(13, c_app);
    // final: lower-bound pass for every row, publish counts
    __threadfence_block();
    for (int r = 0; r < 16; ++r) {
      const int row = row_base + r;
      if ((u0 + row) >= p.n_users) continue;
      refine_row(row);
      if (lane == 0) p.cand_cnt[u0 + row] = min(s_cnt[row], TC_CAP);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS));
  }
}

// ------------------------------------------------------------------------------------------------ exact re-score + top-K
template <int D>
__global__ void __launch_bounds__(128) tc_rescore_kernel(const float* __restrict__ rep_users, const int64_t* __restrict__ users,
                                                         int n_users, const float* __restrict__ rep_items,
                                                         const Cand* __restrict__ cand, const int* __restrict__ cand_cnt, int k,
                                                         int32_t* __restrict__ out_ids, float* __restrict__ out_scores,
                                                         const TcParams p) {
  extern __shared__ __align__(16) unsigned char raw[];
  Cand* buf = reinterpret_cast<Cand*>(raw);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int u = blockIdx.x * 4 + warp;
  if (u >= n_users) return;
  Cand* row = buf + (size_t)warp * TC_CAP;
  const int n = cand_cnt[u];
  const float* ur = rep_users + (size_t)users[u] * D;
  // the user's row is shared by every candidate of the warp: one copy in shared memory, read as broadcast 128-bit words
  __shared__ __align__(16) float s_user[4][D];
  for (int d = lane; d < D; d += 32) s_user[warp][d] = __ldg(ur + d);
  __syncwarp();
  int np = 32;
  while (np < n) np <<= 1;
  for (int t = lane; t < np; t += 32) {
    Cand c{-INFINITY, INT32_MAX};
    if (t < n) {
      const int id = cand[(size_t)u * TC_CAP + t].id;
      // masks (trainer.py:155-167): train / val items of this user, banned range -- dropped before the exact score
      const bool masked = (id >= p.banned_lo && id < p.banned_hi) || row_has(p.excl_ptr_a, p.excl_idx_a, users[u], id) ||
                          row_has(p.excl_ptr_b, p.excl_idx_b, users[u], id);
      if (!masked) {
        // one lane per candidate (the oracle's chain runs d = 0..D-1 in order, so a row cannot be split over lanes); the row
        // is read 16 B at a time with 8 loads in flight -- a lane consumes whole sectors while they are in flight instead of
        // counting on L1 to keep 32 lanes x 512 B rows alive between 4-byte loads (measured: 2.67 -> see DESIGN 4.6)
        const float4* vr = reinterpret_cast<const float4*>(rep_items + (size_t)id * D);
        const float4* uw = reinterpret_cast<const float4*>(s_user[warp]);
        float acc = 0.f;
#pragma unroll 8
        for (int d4 = 0; d4 < D / 4; ++d4) {
          const float4 v = __ldg(vr + d4);
          const float4 w4 = uw[d4];
          acc = fmaf(w4.x, v.x, acc);
          acc = fmaf(w4.y, v.y, acc);
          acc = fmaf(w4.z, v.z, acc);
          acc = fmaf(w4.w, v.w, acc);
        }
        c = Cand{acc, id};
      }
    }
    row[t] = c;
  }
  __syncwarp();
  warp_sort_desc(row, np, lane);
  for (int j = lane; j < k; j += 32) {
    const bool ok = j < n && row[j].id != INT32_MAX;  // masked candidates sorted to the end
    out_ids[(size_t)u * k + j] = ok ? row[j].id : -1;
    out_scores[(size_t)u * k + j] = ok ? row[j].s : -INFINITY;
  }
}

// ------------------------------------------------------------------------------------------------ host
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* sym = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(sym);
  }
  return fn;
}
// bf16 [rows, D] row-major; box = 64 columns (128 B, SWIZZLE_128B) x box_rows
static int make_map(CUtensorMap* m, const void* base, int rows, int d, int box_rows) {
  EncodeTiledFn enc = get_encode();
  if (!enc) return fail(B200REC_ERR_CUDA, "%s: %s", "score_topk_tc", "cuTensorMapEncodeTiled not available");
  cuuint64_t dims[2] = {(cuuint64_t)d, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)d * 2};
  cuuint32_t box[2] = {64, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(B200REC_ERR_CUDA, "%s: %s", "score_topk_tc", "cuTensorMapEncodeTiled failed");
  return 0;
}

static size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

}  // namespace b200rec
using namespace b200rec;

// workspace carve-up shared by the size query and the launch
struct TcWorkspace {
  size_t users_bf16, items_bf16, unorm, vnorm, vmax, cand, cnt, ovf, total;
};
static TcWorkspace tc_layout(int nb, int ni, int d) {
  TcWorkspace w;
  size_t o = 0;
  w.users_bf16 = o; o = align_up(o + (size_t)nb * d * 2, 1024);
  w.items_bf16 = o; o = align_up(o + (size_t)ni * d * 2, 1024);
  w.unorm = o; o = align_up(o + (size_t)nb * 4, 256);
  w.vnorm = o; o = align_up(o + (size_t)ni * 4, 256);
  w.vmax = o; o = align_up(o + ((size_t)ni / 128 + 2) * 4, 256);  // per-tile max item norm (BN = 128)
  w.cand = o; o = align_up(o + (size_t)nb * TC_CAP * sizeof(Cand), 256);
  w.cnt = o; o = align_up(o + (size_t)nb * 4, 256);
  w.ovf = o; o = align_up(o + (size_t)nb * 4, 256);
  w.total = o + 1024;
  return w;
}

int64_t b200rec_score_topk_tc_workspace(int nb, int ni, int d) { return (int64_t)tc_layout(nb, ni, d).total; }
bool b200rec_score_topk_tc_supported(int d) { return d == 64 || d == 128; }

template <int D, int BN, int STAGES>
static int tc_launch(const float* rep_users, const int64_t* users, int nb, const float* rep_items, int ni,
                     const int32_t* ea_ptr, const int32_t* ea_idx, const int32_t* eb_ptr, const int32_t* eb_idx, int blo, int bhi,
                     int k, int32_t* out_ids, float* out_scores, int32_t* out_overflow, void* workspace, cudaStream_t st) {
  uint8_t* ws = reinterpret_cast<uint8_t*>(align_up((size_t)workspace, 1024));
  const TcWorkspace w = tc_layout(nb, ni, D);
  __nv_bfloat16* ub = reinterpret_cast<__nv_bfloat16*>(ws + w.users_bf16);
  __nv_bfloat16* ib = reinterpret_cast<__nv_bfloat16*>(ws + w.items_bf16);
  float* unorm = reinterpret_cast<float*>(ws + w.unorm);
  float* vnorm = reinterpret_cast<float*>(ws + w.vnorm);
  unsigned* vmax = reinterpret_cast<unsigned*>(ws + w.vmax);
  Cand* cand = reinterpret_cast<Cand*>(ws + w.cand);
  int* cnt = reinterpret_cast<int*>(ws + w.cnt);
  int* ovf = out_overflow ? out_overflow : reinterpret_cast<int*>(ws + w.ovf);
  B2_CUDA(cudaMemsetAsync(vmax, 0, ((size_t)ni / BN + 2) * 4, st));
  static_assert(BN == 128, "workspace layout assumes 128-item tiles");
  constexpr int G = D / 4;
  tc_prep_kernel<D><<<ceil_div((long long)nb * G, 256), 256, 0, st>>>(rep_users, users, nb, ub, unorm, nullptr, 1);
  B2_LAUNCHED();
  tc_prep_kernel<D><<<ceil_div((long long)ni * G, 256), 256, 0, st>>>(rep_items, nullptr, ni, ib, vnorm, vmax, BN);
  B2_LAUNCHED();
  CUtensorMap mi;
  int rc = make_map(&mi, ib, ni, D, BN);
  if (rc) return rc;
  TcParams p;
  p.users = users; p.n_users = nb; p.n_items = ni; p.k = k;
  p.excl_ptr_a = ea_ptr; p.excl_idx_a = ea_idx; p.excl_ptr_b = eb_ptr; p.excl_idx_b = eb_idx;
  p.banned_lo = blo; p.banned_hi = bhi; p.unorm = unorm; p.ub = ub; p.tile_vmax_bits = vmax; p.cand = cand; p.cand_cnt = cnt; p.overflow = ovf;
  B2_CUDA(cudaMemsetAsync(ovf, 0, (size_t)nb * sizeof(int), st));
  const size_t smem = 1024 + (size_t)STAGES * BN * D * 2 + TC_RINGS * TC_QCAP * sizeof(HitRec) +
                      TC_CONSUMERS * TC_CAP * sizeof(Cand) + 14 * TC_M * 4 + 1024;
  B2_CUDA(cudaFuncSetAttribute(score_tc2_kernel<D, BN, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
#ifdef B200REC_TC_PROF
  static unsigned long long* prof_d = nullptr;
  if (!prof_d) B2_CUDA(cudaMalloc(&prof_d, 16 * sizeof(unsigned long long)));
  B2_CUDA(cudaMemsetAsync(prof_d, 0, 16 * sizeof(unsigned long long), st));
  p.prof = prof_d;
#endif
  score_tc2_kernel<D, BN, STAGES><<<ceil_div(nb, TC_M), TC2_THREADS, smem, st>>>(mi, p);
  B2_LAUNCHED();
#ifdef B200REC_TC_PROF
  {
    unsigned long long h[16];
    B2_CUDA(cudaStreamSynchronize(st));
    B2_CUDA(cudaMemcpy(h, prof_d, sizeof(h), cudaMemcpyDeviceToHost));
    const double ctas = (double)ceil_div(nb, TC_M), tiles = (double)ceil_div(ni, BN);
    fprintf(stderr, "[tc prof] ctas %.0f tiles %.0f | per CTA (cycles): mma sweep %.0f, wait acc %.0f, wait tile %.0f | per drain warp: wait mma %.0f, "
            "wait ring %.0f | per consumer warp: total %.0f, record loop %.0f, refine %.0f, iterations %.0f, passes %.0f, "
            "refines %.0f | per CTA: record batches %.0f, records %.0f, appended %.0f\n", ctas, tiles, h[14] / ctas, h[0] / ctas, h[1] / ctas,
            h[2] / ctas / 8, h[3] / ctas / 8, h[7] / ctas / 8, h[8] / ctas / 8, h[9] / ctas / 8, h[10] / ctas / 8,
            h[11] / ctas / 8, h[12] / ctas / 8, h[5] / ctas, h[6] / ctas, h[13] / ctas);
  }
#endif
  const size_t rsmem = 4 * TC_CAP * sizeof(Cand);
  tc_rescore_kernel<D><<<ceil_div(nb, 4), 128, rsmem, st>>>(rep_users, users, nb, rep_items, cand, cnt, k, out_ids, out_scores, p);
  B2_LAUNCHED();
  return 0;
}

int b200rec_score_topk_tc(const float* rep_users, const int64_t* users, int nb, const float* rep_items, int ni, int d,
                          const int32_t* ea_ptr, const int32_t* ea_idx, const int32_t* eb_ptr, const int32_t* eb_idx, int blo,
                          int bhi, int k, int32_t* out_ids, float* out_scores, int32_t* out_overflow, void* workspace,
                          cudaStream_t st) {
  if (d == 64)
    return tc_launch<64, 128, 3>(rep_users, users, nb, rep_items, ni, ea_ptr, ea_idx, eb_ptr, eb_idx, blo, bhi, k, out_ids,
                                 out_scores, out_overflow, workspace, st);
  if (d == 128)
    return tc_launch<128, 128, 3>(rep_users, users, nb, rep_items, ni, ea_ptr, ea_idx, eb_ptr, eb_idx, blo, bhi, k, out_ids,
                                  out_scores, out_overflow, workspace, st);
  return fail(B200REC_ERR_UNSUPPORTED, "%s: %s", "score_topk_tc", "tensor-core scoring supports embedding size 64 / 128");
}
