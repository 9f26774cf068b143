// K2: query-by-corpus scoring on the 5th-gen tensor cores with a streaming top-k' selection in
// the epilogue.  The [Q, N] score matrix never reaches HBM.
//
// Replaces, per corpus shard, the hot loop of sentence-transformers 2.2.2
// InformationRetrievalEvaluator.compute_metrices (constructed at
// /root/reference/ir_evauation_script.py:107-123, models/evaluators.py:572-588):
//     pair_scores = score_function(query_embeddings, sub_corpus_embeddings)   # torch.mm
//     torch.topk(pair_scores, min(max_k, chunk), dim=1, largest=True, sorted=False)
//
// Structure (one persistent CTA per SM, 256 threads, warp-specialised).  Default: CTA PAIRS
// (cta_group::2, cluster of 2): the pair computes a 256-query x 256-corpus-row tile, each CTA
// stages its own 128 query rows and HALF of the corpus tile, which cuts the L2->smem traffic per
// FLOP by 1.5x against the single-CTA tile and leaves room for a 6-stage ring.
//   warp 4      TMA producer: 128x64 query tile + 128x64 (pair) / 256x64 (single) corpus tile per
//               k-block into a 128B-swizzled smem ring (mbarrier full/empty pipeline); in pair
//               mode both CTAs signal the LEADER's full barrier
//   warp 5      MMA issuer (leader CTA only in pair mode): one thread issues tcgen05.mma
//               (M=256|128, N=256, K=16) x4 per k-block, accumulating in TMEM; two 256-column
//               accumulators are double-buffered so the epilogue of tile t overlaps the MMAs of
//               tile t+1; tcgen05.commit (multicast to both CTAs) frees smem slots and publishes
//               accumulators
//   warp 6      TMEM allocator
//   warps 0-3   epilogue (lower warp ids: the arbiter serves the TMA/MMA warps first): thread r of the CTA owns query row r of the tile (TMEM lane r);
//               tcgen05.ld 32 columns at a time, reject against the row's running threshold
//               (k'-th best score seen), append survivors to the row's candidate buffer,
//               warp-cooperative radix-select compaction when a buffer fills up.
//
// Work decomposition: the corpus is cut into `stripes` of `tiles_per_stripe` 256-row tiles; a
// work unit is (stripe, 128-query tile).  Units are ordered stripe-major so the CTAs running at
// the same time walk the same corpus stripe -> every corpus tile is fetched from HBM once and
// re-read from L2 by the other query tiles.  Thresholds are shared between units of the same
// query row through a global hint array (atomicMax), which removes the cold start of later units.
#include "qst_common.cuh"
#include "sm100_ptx.cuh"
#include "select_common.cuh"
#include <stdlib.h>

namespace qst {

constexpr int BM = 128;          // query rows per tile  (UMMA M)
constexpr int BN = 256;          // corpus rows per tile (UMMA N)
constexpr int BK = 64;           // bf16 elements per k-block = one 128-byte swizzle row
constexpr int UMMA_K = 16;
constexpr int kScoreThreads = 256;
constexpr uint32_t A_BYTES = BM * BK * 2;
constexpr uint32_t TMEM_COLS = 512;  // 2 accumulators x 256 fp32 columns
constexpr int MAX_STAGES = 6;

template <int CTAS>
struct Cfg {
  static constexpr int STAGES = CTAS == 1 ? 4 : 6;                // even: the two producer warps take one parity each
  static constexpr int B_ROWS = BN / CTAS;                 // corpus rows staged by one CTA per tile
  static constexpr uint32_t B_BYTES = B_ROWS * BK * 2;
  static constexpr uint32_t STAGE_BYTES = A_BYTES + B_BYTES;  // per CTA
  static constexpr int UNIT_ROWS = BM * CTAS;              // query rows of one work unit
  static constexpr size_t SMEM = (size_t)STAGES * STAGE_BYTES + 1024;  // + alignment slack
};

struct ScoreParams {
  int Q, N, num_kb;
  int m_tiles, n_tiles, stripes, tiles_per_stripe, units;
  int kunit, cap;      // entries a unit keeps per row after a compaction; buffer capacity
  uint32_t* thr_hint;  // [m_tiles*UNIT_ROWS] ordered keys, 0 = no threshold yet
  int* unit_cnt;       // [units*UNIT_ROWS]
  uint32_t* unit_thr;  // [units*UNIT_ROWS] final threshold key of the unit row (upper bound of its discards)
  uint2* unit_cand;    // [units*UNIT_ROWS*cap]  (key, corpus row)
  float* dense_out;    // dense mode only: [Q, N]
  // corpus-sharded runs: hint arrays of the OTHER ranks (peer-mapped device memory over
  // NVLink/NVSwitch).  A new row threshold is pushed to every peer with a remote atomicMax, so all
  // shards filter against the best threshold any shard has established for the query.
  uint32_t* peer_hint[QST_MAX_PEERS];
  int n_peers;
  int chunkmax;        // 1: rows also raise their threshold from the running top-16 chunk maxima
  int debug;           // QST_SCORE_DEBUG ablation bits (0 in production), see launch_score()
  // query-stationary kernel: the bf16 query operand as plain rows (it is copied into TMEM, not TMA-staged)
  const uint4* q_rows; // [Q, D_pad] bf16 viewed as 16-byte vectors
  int q_pitch16;       // D_pad / 8
};

// ------------------------------------------------------------------------------------------
// Register-resident compaction for buffers of at most 256 entries (8 per lane): ONE round trip to
// the buffer, a 256-bin histogram over the observed key range, keep every entry in or above the
// bin in which the running count (from the top) reaches k.  Keeps >= k entries (k plus whatever
// shares the boundary bin), drops only entries strictly below the returned threshold T.  Returns
// false (buffer untouched) when the histogram cannot separate the keys enough to free space, in
// which case the caller falls back to the exact radix select.
// ------------------------------------------------------------------------------------------
constexpr int kSmallCap = 256;

__device__ __forceinline__ bool warp_compact_small(uint2* __restrict__ buf, int n, int k, int max_keep, int* hist,
                                                   int lane, uint32_t& T_out, int& n_out) {
  uint2 e[kSmallCap / 32];
  uint32_t lo = 0xffffffffu, hi = 0u;
#pragma unroll
  for (int i = 0; i < kSmallCap / 32; ++i) {
    const int j = lane + 32 * i;
    e[i] = make_uint2(0u, 0u);
    if (j < n) {
      e[i] = buf[j];
      lo = min(lo, e[i].x);
      hi = max(hi, e[i].x);
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    lo = min(lo, __shfl_xor_sync(0xffffffffu, lo, o));
    hi = max(hi, __shfl_xor_sync(0xffffffffu, hi, o));
  }
  // bin = floor((key - lo) * 256 / (hi - lo + 1)) evaluated in fp32: rounding is monotone, so the
  // bin index is non-decreasing in the key, which is all the selection needs
  const float scale = 256.0f / ((float)(hi - lo) + 1.0f);
#pragma unroll
  for (int i = 0; i < 8; ++i) hist[lane * 8 + i] = 0;
  __syncwarp();
  int bin[kSmallCap / 32];
#pragma unroll
  for (int i = 0; i < kSmallCap / 32; ++i) {
    bin[i] = min(255, (int)((float)(e[i].x - lo) * scale));
    if (lane + 32 * i < n) atomicAdd(&hist[bin[i]], 1);
  }
  __syncwarp();
  int c[8], ls = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) { c[i] = hist[lane * 8 + i]; ls += c[i]; }
  int incl = ls;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int t = __shfl_down_sync(0xffffffffu, incl, o);
    if (lane + o < 32) incl += t;
  }
  const int above = incl - ls;
  const bool mine = (above < k) && (k <= above + ls);
  const unsigned who = __ballot_sync(0xffffffffu, mine);
  const int src = 31 - __clz(who);
  int bstar = 0, kept = 0;
  if (lane == src) {
    int run = above;
#pragma unroll
    for (int i = 7; i >= 0; --i) {
      if (run < k && k <= run + c[i]) { bstar = lane * 8 + i; kept = run + c[i]; }
      run += c[i];
    }
  }
  bstar = __shfl_sync(0xffffffffu, bstar, src);
  kept = __shfl_sync(0xffffffffu, kept, src);
  __syncwarp();
  if (kept > max_keep) return false;
  const unsigned lt = (1u << lane) - 1u;
  int w = 0;
  uint32_t tmin = 0xffffffffu;   // smallest kept key: everything dropped is strictly below it
#pragma unroll
  for (int i = 0; i < kSmallCap / 32; ++i) {
    const bool keep = (lane + 32 * i < n) && (bin[i] >= bstar);
    const unsigned km = __ballot_sync(0xffffffffu, keep);
    if (keep) {
      buf[w + __popc(km & lt)] = e[i];
      tmin = min(tmin, e[i].x);
    }
    w += __popc(km);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) tmin = min(tmin, __shfl_xor_sync(0xffffffffu, tmin, o));
  T_out = tmin;
  __syncwarp();
  n_out = w;
  return true;
}

// New row threshold -> hint array of this rank and (sharded runs) of every peer.
__device__ __forceinline__ void publish_threshold(const ScoreParams& P, int grow, uint32_t key) {
  atomicMax(&P.thr_hint[grow], key);
  for (int pr = 0; pr < P.n_peers; ++pr) atomicMax(&P.peer_hint[pr][grow], key);   // remote RED, fire and forget
}

// One row's compaction, OUT OF LINE on purpose: the selection code is large and rarely runs, and the
// epilogue has a single warp per scheduler, so every instruction-cache miss on the way through the
// filter is paid in full -- keeping this out of the filter's code path keeps that path short.
// Returns (new count << 32) | threshold key.
__device__ __noinline__ unsigned long long compact_one_row(uint2* bp, int n, int kunit, int max_keep, int cap,
                                                           int* hist, int lane) {
  uint32_t T = 0u;
  int n_new = kunit;
  // fast: one-pass histogram compaction in registers; exact 4-pass radix select otherwise
  // (large buffers, or keys the histogram cannot separate)
  if (!(cap <= kSmallCap && warp_compact_small(bp, n, kunit, max_keep, hist, lane, T, n_new))) {
    T = warp_select_compact(bp, n, kunit, hist, lane);
    n_new = kunit;
  }
  return ((unsigned long long)(uint32_t)n_new << 32) | T;
}

// Compacts the candidate buffers of the rows in `need` (one bit per lane) down to ~kunit entries
// and publishes each row's new threshold to the hint array.
__device__ __forceinline__ void compact_rows(unsigned need, int max_keep, const ScoreParams& P, int grow, float& thr,
                                             int& cnt, uint2* my_buf, int* hist, int lane) {
  while (need) {
    const int r = __ffs(need) - 1;
    need &= need - 1;
    uint2* bp = reinterpret_cast<uint2*>(__shfl_sync(0xffffffffu, (unsigned long long)my_buf, r));
    const int n = __shfl_sync(0xffffffffu, cnt, r);
    const unsigned long long res = compact_one_row(bp, n, P.kunit, max_keep, P.cap, hist, lane);
    if (lane == r) {
      const uint32_t T = (uint32_t)res;
      cnt = (int)(res >> 32);
      thr = fmaxf(thr, key_to_float(T));
      publish_threshold(P, grow, T);
    }
  }
}

// ------------------------------------------------------------------------------------------
// One 32-row x 32-column block of scores held by a warp (lane = query row, v[j] = column col0+j).
// Fast path: per-lane max against the row threshold, one ballot, nothing else.
// Slow path (some lane has a score above its threshold): the lanes with hits park their 32 values
// in a warp-private shared tile; the warp then serves one hit row at a time -- lane j looks at
// column j of that row, a ballot ranks the survivors and they are appended to the row's candidate
// buffer with consecutive (coalesced) stores.  No per-element branches, no dynamic register indexing.
// ------------------------------------------------------------------------------------------
constexpr int kStagePitch = 33;  // floats per staged row: conflict-free for both access patterns

constexpr int kChunkMax = 20;   // running top-20 of a row's per-chunk maxima (count-20 threshold guarantee)

// sorted (descending) insertion of x into cm[]
__device__ __forceinline__ void cm_insert(float (&cm)[kChunkMax], float x) {
#pragma unroll
  for (int i = kChunkMax - 1; i > 0; --i) cm[i] = fmaxf(cm[i], fminf(cm[i - 1], x));   // old cm[i-1]: descending i
  cm[0] = fmaxf(cm[0], x);
}

// End of a unit: drop every entry of every row that is not above the row's final threshold (most
// were appended while the threshold was still maturing).  One coalesced load and one store per
// row; the loads of FOUR rows are issued back to back before any of them is compacted, so a warp's
// 32 rows cost eight memory latencies instead of 32 dependent round trips.  Rows that still hold
// more than `max_final` entries afterwards (heavy ties) are returned in the mask and go through the
// selecting compaction.
__device__ __forceinline__ unsigned final_filter_rows(unsigned need, int max_final, float thr, int& cnt,
                                                      uint2* my_buf, int lane) {
  constexpr int E = kSmallCap / 32;
  constexpr int B = 4;
  const unsigned lt = (1u << lane) - 1u;
  unsigned still = 0u;
  while (need) {
    int r[B], n[B];
    uint2* bp[B];
    uint32_t tkey[B];
    uint2 e[B][E];
#pragma unroll
    for (int b = 0; b < B; ++b) {
      const bool ok = need != 0u;
      r[b] = ok ? __ffs(need) - 1 : 0;
      need &= need - 1;   // 0 stays 0
      bp[b] = reinterpret_cast<uint2*>(__shfl_sync(0xffffffffu, (unsigned long long)my_buf, r[b]));
      n[b] = ok ? __shfl_sync(0xffffffffu, cnt, r[b]) : 0;
      tkey[b] = float_to_key(__shfl_sync(0xffffffffu, thr, r[b]));
      if (!ok) r[b] = -1;
#pragma unroll
      for (int i = 0; i < E; ++i) {
        const int j = lane + 32 * i;
        e[b][i] = j < n[b] ? bp[b][j] : make_uint2(0u, 0u);
      }
    }
#pragma unroll
    for (int b = 0; b < B; ++b) {
      int w = 0;
#pragma unroll
      for (int i = 0; i < E; ++i) {
        const bool keep = (lane + 32 * i < n[b]) && (e[b][i].x > tkey[b]);
        const unsigned km = __ballot_sync(0xffffffffu, keep);
        if (keep) bp[b][w + __popc(km & lt)] = e[b][i];
        w += __popc(km);
      }
      if (lane == r[b]) cnt = w;
      if (r[b] >= 0 && w > max_final) still |= 1u << r[b];
    }
  }
  __syncwarp();
  return still;
}

template <bool DENSE>
__device__ __forceinline__ void epilogue_chunk(uint32_t (&v)[32], int col0, const ScoreParams& P, int grow, bool row_ok,
                                               float& thr, int& cnt, uint2* my_buf, float* stage, int* hist, int lane,
                                               float (&cm)[kChunkMax], bool update_cm) {
  if (P.debug & 2) {  // ablation: TMEM reads only
    if (v[0] == 0x7fc12345u) P.unit_cnt[0] = 1;
    return;
  }
  if (col0 + 32 > P.N) {  // ragged last tile: columns >= N were zero-filled by TMA
#pragma unroll
    for (int j = 0; j < 32; ++j)
      if (col0 + j >= P.N) v[j] = 0xff800000u;  // -inf
  }
  if (DENSE) {
    if (row_ok) {
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (col0 + j < P.N) P.dense_out[(size_t)grow * P.N + col0 + j] = __uint_as_float(v[j]);
    }
    return;
  }
  // row maximum as a shallow tree (3-input max): the epilogue runs one warp per scheduler, so the
  // dependency depth of this reduction is paid in full
  float m8[8];
#pragma unroll
  for (int g = 0; g < 8; ++g)
    m8[g] = fmaxf(fmaxf(__uint_as_float(v[4 * g]), __uint_as_float(v[4 * g + 1])),
                  fmaxf(__uint_as_float(v[4 * g + 2]), __uint_as_float(v[4 * g + 3])));
  const float mx = fmaxf(fmaxf(fmaxf(m8[0], m8[1]), fmaxf(m8[2], m8[3])), fmaxf(fmaxf(m8[4], m8[5]), fmaxf(m8[6], m8[7])));
  const bool hit = mx > thr;
  if (__ballot_sync(0xffffffffu, hit) == 0u) return;
  // Running top-20 of this row's per-chunk maxima (sorted registers): 20 distinct documents of this
  // unit score at least cm[19], so it is a threshold with the same kind of guarantee as a
  // compaction's, but it matures chunk by chunk -- a unit does not have to fill and compact its
  // buffer to get going (a cold unit seeds cm[] from its whole first tile, see the kernel).  Takes
  // effect after this chunk's appends.
  float thr_next = thr;
  if (P.chunkmax && hit && update_cm) {
    cm_insert(cm, mx);
    thr_next = fmaxf(thr, cm[kChunkMax - 1]);
  }
  __syncwarp();
  // Some lane has a score above its threshold.  Which ones: a 32-bit mask per lane (predicated, no
  // branches).  Every lane then appends its own hits to its own row's buffer: one hit (by far the
  // common case) is the row maximum and needs nothing else; several hits are read back one by one
  // from the lane's row of the staging tile (registers cannot be indexed by the bit position).
  // The stores are 8 bytes per lane to 32 different buffers -- uncoalesced but few, and far cheaper
  // in instructions than serving the rows cooperatively (the epilogue is issue-bound: one warp per
  // scheduler).
  // (A group-wise build -- only look at a group of four columns when some lane's group maximum beats its
  // threshold -- executes fewer instructions but was 5 % SLOWER: eight votes and branches per chunk cost
  // more than 32 predicated compare + select pairs.  profiles/r02_k2_ab_grouped_mask.txt)
  unsigned above = 0u;
#pragma unroll
  for (int j = 0; j < 32; ++j) above |= (__uint_as_float(v[j]) > thr) ? (1u << j) : 0u;
  const int n_above = __popc(above);
  if (n_above == 1) {
    my_buf[cnt] = make_uint2(float_to_key(mx), (uint32_t)(col0 + __ffs(above) - 1));
    ++cnt;
  } else if (n_above > 1) {
#pragma unroll
    for (int j = 0; j < 32; ++j) stage[lane * kStagePitch + j] = __uint_as_float(v[j]);
    while (above) {
      const int j = __ffs(above) - 1;
      above &= above - 1;
      my_buf[cnt] = make_uint2(float_to_key(stage[lane * kStagePitch + j]), (uint32_t)(col0 + j));
      ++cnt;
    }
  }
  __syncwarp();
  thr = thr_next;
  // keep room for the next chunk's worst case (32 appends)
  compact_rows(__ballot_sync(0xffffffffu, cnt > P.cap - 32), P.cap - 64, P, grow, thr, cnt, my_buf, hist, lane);
}

// debug trace (QST_SCORE_DEBUG & 32): per CTA, per tile of its FIRST unit: globaltimer ns, candidate count
// and threshold of the CTA's row 0
constexpr int kTraceTiles = 64;
__device__ long long g_trace_ns[160 * kTraceTiles];
__device__ int g_trace_cnt[160 * kTraceTiles];
__device__ float g_trace_thr[160 * kTraceTiles];
__device__ __forceinline__ long long globaltimer_ns() {
  long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}

template <int CTAS, bool DENSE>
__global__ void __launch_bounds__(kScoreThreads, 1)
score_select_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_c,
                    const ScoreParams P) {
  using C = Cfg<CTAS>;
  constexpr int STAGES = C::STAGES;
  constexpr uint32_t STAGE_BYTES = C::STAGE_BYTES;
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t s_full[MAX_STAGES];
  __shared__ __align__(8) uint64_t s_empty[MAX_STAGES];
  __shared__ __align__(8) uint64_t s_tmem_full[2];
  __shared__ __align__(8) uint64_t s_tmem_empty[2];
  __shared__ uint32_t s_tmem_base;
  __shared__ int s_hist[4][256];
  __shared__ float s_stage[4][32 * kStagePitch];

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  // Role -> warp id.  The SM's warp arbiter favours higher warp ids, so the two single-thread
  // roles that feed the tensor pipe (TMA producer, MMA issuer) sit ABOVE the four epilogue warps
  // and are never queued behind their filtering code.
  constexpr int kWarpTma = 4, kWarpMma = 5, kWarpAlloc = 6, kWarpTma2 = 7;   // warps 0-3: epilogue
  const uint32_t rank = CTAS == 2 ? ptx::cluster_ctarank() : 0u;   // 0 = pair leader
  const int group = blockIdx.x / CTAS, n_groups = gridDim.x / CTAS;
  // 128B swizzle needs 1024-byte aligned stage bases
  const uint32_t smem_base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;

  if (warp == kWarpTma && lane == 0) {
    ptx::prefetch_tmap(&tmap_q);
    ptx::prefetch_tmap(&tmap_c);
  }
  if (warp == kWarpMma && lane == 0) {
    for (int s = 0; s < STAGES; ++s) {
      ptx::mbar_init(ptx::smem_u32(&s_full[s]), 1);
      ptx::mbar_init(ptx::smem_u32(&s_empty[s]), 1);
    }
    for (int a = 0; a < 2; ++a) {
      ptx::mbar_init(ptx::smem_u32(&s_tmem_full[a]), 1);
      ptx::mbar_init(ptx::smem_u32(&s_tmem_empty[a]), 4 * CTAS);  // one arrive per epilogue warp (of both CTAs)
    }
    ptx::fence_mbar_init();
  }
  if (warp == kWarpAlloc) {
    if (CTAS == 2) ptx::tmem_alloc_pair(ptx::smem_u32(&s_tmem_base), TMEM_COLS);
    else ptx::tmem_alloc(ptx::smem_u32(&s_tmem_base), TMEM_COLS);
  }
  ptx::tc_fence_before();
  if (CTAS == 2) ptx::cluster_sync_all(); else __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = s_tmem_base;

  if (warp == kWarpTma || warp == kWarpTma2) {
    // ============================== TMA producers =============================
    // TWO producer threads, one per stage parity (STAGES is even).  A thread needs ~380 cycles for
    // wait + expect_tx + the first box and ~140 for the second (profiles/ubench_umma.cu: the cost is
    // the issuing thread's own latency chain, it scales with the number of issuing threads), i.e.
    // ~580 cycles per k-block against 512 cycles of MMA work: a single producer cannot keep the
    // tensor pipe fed (QST_SCORE_DEBUG=9: 7.7 ms of feed alone for 8 ms of MMA work).
    const uint32_t my_parity = warp == kWarpTma ? 0u : 1u;
    const bool one_producer = (P.debug & 64) != 0;   // ablation: the r01 single producer
    uint32_t stage = 0, phase = 0;
    for (int u = group; u < P.units; u += n_groups) {
      const int s = u / P.m_tiles, m = u - s * P.m_tiles;
      const int t0 = s * P.tiles_per_stripe;
      const int t1 = min(t0 + P.tiles_per_stripe, P.n_tiles);
      const int q_row = m * C::UNIT_ROWS + (int)rank * BM;
      for (int t = t0; t < t1; ++t) {
        const int c_row = (P.debug & 4) ? 0 : t * BN + (int)rank * C::B_ROWS;
        for (int kb = 0; kb < P.num_kb; ++kb) {
          if (one_producer ? (my_parity != 0u) : ((stage & 1u) != my_parity)) {
            if (++stage == STAGES) { stage = 0; phase ^= 1u; }
            continue;
          }
          ptx::mbar_wait(ptx::smem_u32(&s_empty[stage]), phase ^ 1u);
          if (ptx::elect_one_sync()) {   // whole warp runs the loop, one lane issues
            const uint32_t full = ptx::smem_u32(&s_full[stage]);
            const uint32_t sa = smem_base + stage * STAGE_BYTES;
            // queries are re-read by every corpus tile (keep in L2); a corpus tile is re-read by the
            // other query tiles walking the same stripe (normal priority)
            if (CTAS == 1) {
              ptx::mbar_arrive_expect_tx(full, STAGE_BYTES);
              ptx::tma_load_2d(sa, &tmap_q, full, kb * BK, q_row, ptx::kEvictLast);
              ptx::tma_load_2d(sa + A_BYTES, &tmap_c, full, kb * BK, c_row, ptx::kEvictNormal);
            } else {
              // both CTAs' bytes are accounted on the leader's barrier, which the MMA thread waits on
              if (rank == 0) ptx::mbar_arrive_expect_tx(full, STAGE_BYTES * 2);
              const uint32_t leader_full = ptx::mapa_shared(full, 0);
              ptx::tma_load_2d_pair(sa, &tmap_q, leader_full, kb * BK, q_row, ptx::kEvictLast);
              ptx::tma_load_2d_pair(sa + A_BYTES, &tmap_c, leader_full, kb * BK, c_row, ptx::kEvictNormal);
            }
          }
          __syncwarp();
          if (++stage == STAGES) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == kWarpMma && rank == 0) {
    // ============================== MMA issuer (whole warp, one lane issues) ================================
    constexpr uint32_t idesc = ptx::make_idesc_bf16_f32(BM * CTAS, BN);
    const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);   // warp-uniform for the compiler
    const bool quarter_mmas = (P.debug & 8) != 0;                      // ablation: one MMA per k-block
    uint32_t stage = 0, phase = 0, acc = 0, acc_phase = 0;
    for (int u = group; u < P.units; u += n_groups) {
      const int s = u / P.m_tiles;
      const int t0 = s * P.tiles_per_stripe;
      const int t1 = min(t0 + P.tiles_per_stripe, P.n_tiles);
      for (int t = t0; t < t1; ++t) {
        ptx::mbar_wait(ptx::smem_u32(&s_tmem_empty[acc]), acc_phase ^ 1u);
        ptx::tc_fence_after();
        const uint32_t d_tmem = tmem_u + acc * BN;
        for (int kb = 0; kb < P.num_kb; ++kb) {
          ptx::mbar_wait(ptx::smem_u32(&s_full[stage]), phase);
          ptx::tc_fence_after();
          if (ptx::elect_one_sync()) {
            const uint32_t sa = smem_base + stage * STAGE_BYTES;
            const uint64_t da = ptx::make_sw128_kmajor_desc(sa);
            const uint64_t db = ptx::make_sw128_kmajor_desc(sa + A_BYTES);
            const uint32_t first = kb != 0 ? 1u : 0u;
            // four K steps: +32 bytes (16 bf16) inside the 128-byte swizzle row = +2 in the addr>>4 field
            if (CTAS == 2) {
              ptx::umma_bf16_pair(d_tmem, da, db, idesc, first);
              if (!quarter_mmas) {
                ptx::umma_bf16_pair(d_tmem, da + 2, db + 2, idesc, 1u);
                ptx::umma_bf16_pair(d_tmem, da + 4, db + 4, idesc, 1u);
                ptx::umma_bf16_pair(d_tmem, da + 6, db + 6, idesc, 1u);
              }
              // frees the smem slot (in both CTAs) once the MMAs that read it have retired
              ptx::umma_commit_pair(ptx::smem_u32(&s_empty[stage]), 3);
              // last k-block: accumulator ready for the epilogue warps (of both CTAs)
              if (kb + 1 == P.num_kb) ptx::umma_commit_pair(ptx::smem_u32(&s_tmem_full[acc]), 3);
            } else {
              ptx::umma_bf16(d_tmem, da, db, idesc, first);
              if (!quarter_mmas) {
                ptx::umma_bf16(d_tmem, da + 2, db + 2, idesc, 1u);
                ptx::umma_bf16(d_tmem, da + 4, db + 4, idesc, 1u);
                ptx::umma_bf16(d_tmem, da + 6, db + 6, idesc, 1u);
              }
              ptx::umma_commit(ptx::smem_u32(&s_empty[stage]));
              if (kb + 1 == P.num_kb) ptx::umma_commit(ptx::smem_u32(&s_tmem_full[acc]));
            }
          }
          __syncwarp();
          if (++stage == STAGES) { stage = 0; phase ^= 1u; }
        }
        acc ^= 1u;
        if (acc == 0) acc_phase ^= 1u;
      }
    }
  } else if (warp < 4) {
    // ============================== epilogue ==================================
    const int quarter = warp;                     // TMEM lanes [32*quarter, 32*quarter+32)
    const int row_in_unit = (int)rank * BM + quarter * 32 + lane;
    int* hist = s_hist[quarter];
    float* stage = s_stage[quarter];
    uint32_t acc = 0, acc_phase = 0;
    for (int u = group; u < P.units; u += n_groups) {
      const int s = u / P.m_tiles, m = u - s * P.m_tiles;
      const int t0 = s * P.tiles_per_stripe;
      const int t1 = min(t0 + P.tiles_per_stripe, P.n_tiles);
      const int grow = m * C::UNIT_ROWS + row_in_unit;
      const bool row_ok = grow < P.Q && !(P.debug & 16);   // ablation 16: no row ever has a hit
      float thr = row_ok ? -INFINITY : INFINITY;
      float pub = thr;                       // last threshold this row published
      float cm[kChunkMax];
#pragma unroll
      for (int i = 0; i < kChunkMax; ++i) cm[i] = -INFINITY;
      int cnt = 0;
      uint2* my_buf = DENSE ? nullptr : P.unit_cand + ((size_t)u * C::UNIT_ROWS + row_in_unit) * (size_t)P.cap;
      uint32_t next_hint = (!DENSE && row_ok) ? __ldcg(&P.thr_hint[grow]) : 0u;
      const bool tracing = (P.debug & 32) && u == group && warp == 0 && lane == 0 && blockIdx.x < 160;
      if (tracing) g_trace_ns[blockIdx.x * kTraceTiles] = globaltimer_ns();
      bool unit_cold = false;
      for (int t = t0; t < t1; ++t) {
        if (!DENSE && row_ok) {
          // thresholds published by other units of this row; the load was issued one tile ago
          if (next_hint != 0u) thr = fmaxf(thr, key_to_float(next_hint));
          next_hint = __ldcg(&P.thr_hint[grow]);
        }
        if (t == t0) unit_cold = __ballot_sync(0xffffffffu, row_ok && thr == -INFINITY) != 0u;
        ptx::mbar_wait(ptx::smem_u32(&s_tmem_full[acc]), acc_phase);
        ptx::tc_fence_after();
        const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + acc * BN;
        if (P.debug & 1) {  // ablation: no TMEM reads at all
          ptx::tc_fence_before();
          __syncwarp();
          if (lane == 0) {
            if (CTAS == 2) ptx::mbar_arrive_cluster_relaxed(ptx::mapa_shared(ptx::smem_u32(&s_tmem_empty[acc]), 0));
            else ptx::mbar_arrive(ptx::smem_u32(&s_tmem_empty[acc]));
          }
          acc ^= 1u;
          if (acc == 0) acc_phase ^= 1u;
          continue;
        }
        // Cold start (first tile of a unit whose rows have no threshold yet, chunk-maximum mode):
        // pass 1 reads the whole tile once only to find (nearly) every row's 20 best scores (cm[]), so
        // that pass 2 -- the normal filter below -- appends those instead of most of the tile, and
        // the next tiles start from a "20th best of 256" threshold.  Without it a cold unit
        // spends its first three tiles appending and compacting (profiles/small_q_trace.py).
        // chunkmax == 2: the running top-20 of chunk maxima is only kept by units that started cold
        // (some row without a threshold); a warm unit filters against thresholds other units have
        // published, which the chunk maxima would not raise before the buffer's own compaction does,
        // and the 40-instruction sorted insertion on every chunk with a hit is the largest single item
        // of the epilogue's slow path
        bool update_cm = P.chunkmax != 2 || unit_cold;
        if (tracing && t == t0) g_trace_ns[blockIdx.x * kTraceTiles + 40] = globaltimer_ns();
        if (!DENSE && P.chunkmax && t == t0 && unit_cold) {
          uint32_t vc[32];
#pragma unroll 1
          for (int chunk = 0; chunk < BN / 32; ++chunk) {
            ptx::tmem_ld_32x32(taddr + chunk * 32, vc);
            ptx::tmem_ld_wait(vc);
            const int col0 = t * BN + chunk * 32;
            // the chunk's four best (sorted insertion, 2 min/max per step), then those four into
            // cm[]: 20 of the (up to) 32 documents collected this way score >= cm[19]
            float c4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              const float x = (col0 + j < P.N) ? __uint_as_float(vc[j]) : -INFINITY;
              c4[3] = fmaxf(c4[3], fminf(c4[2], x));
              c4[2] = fmaxf(c4[2], fminf(c4[1], x));
              c4[1] = fmaxf(c4[1], fminf(c4[0], x));
              c4[0] = fmaxf(c4[0], x);
            }
#pragma unroll
            for (int i = 0; i < 4; ++i) cm_insert(cm, c4[i]);
          }
          if (row_ok) thr = fmaxf(thr, cm[kChunkMax - 1]);
          update_cm = false;   // cm[] already holds this tile's documents
          if (tracing) g_trace_ns[blockIdx.x * kTraceTiles + 41] = globaltimer_ns();
        }
        // software pipeline over the 8 chunks of 32 columns: the TMEM load of chunk c+1 is in
        // flight while chunk c is filtered (two register buffers, loop unrolled by 2)
        uint32_t va[32], vb[32];
        ptx::tmem_ld_32x32(taddr, va);
        ptx::tmem_ld_wait(va);
#pragma unroll 1
        for (int chunk = 0; chunk < BN / 32; chunk += 2) {
          ptx::tmem_ld_32x32(taddr + (chunk + 1) * 32, vb);
          epilogue_chunk<DENSE>(va, t * BN + chunk * 32, P, grow, row_ok, thr, cnt, my_buf, stage, hist, lane, cm,
                                update_cm);
          ptx::tmem_ld_wait(vb);
          if (chunk + 2 < BN / 32) {
            ptx::tmem_ld_32x32(taddr + (chunk + 2) * 32, va);
          } else {
            // every column of this accumulator is now in registers: hand TMEM back to the MMA warp
            ptx::tc_fence_before();
            __syncwarp();
            if (lane == 0) {
              if (CTAS == 2) ptx::mbar_arrive_cluster_relaxed(ptx::mapa_shared(ptx::smem_u32(&s_tmem_empty[acc]), 0));
              else ptx::mbar_arrive(ptx::smem_u32(&s_tmem_empty[acc]));
            }
          }
          epilogue_chunk<DENSE>(vb, t * BN + (chunk + 1) * 32, P, grow, row_ok, thr, cnt, my_buf, stage, hist, lane, cm,
                                update_cm);
          if (chunk + 2 < BN / 32) ptx::tmem_ld_wait(va);
        }
        if (!DENSE && row_ok && thr > pub) {   // thresholds raised by the chunk maxima during this tile
          publish_threshold(P, grow, float_to_key(thr));
          pub = thr;
        }
        if (tracing && t - t0 + 1 < kTraceTiles) {
          g_trace_ns[blockIdx.x * kTraceTiles + t - t0 + 1] = globaltimer_ns();
          g_trace_cnt[blockIdx.x * kTraceTiles + t - t0 + 1] = cnt;
          g_trace_thr[blockIdx.x * kTraceTiles + t - t0 + 1] = thr;
        }
        acc ^= 1u;
        if (acc == 0) acc_phase ^= 1u;
      }
      if (!DENSE) {
        // leave ~kunit entries per row and publish the unit's final threshold: later units of the
        // same rows start from it, and K3 has less to gather
        unsigned need = __ballot_sync(0xffffffffu, cnt > P.kunit);
        if (P.cap <= kSmallCap && P.chunkmax) need = final_filter_rows(need, P.kunit + 16, thr, cnt, my_buf, lane);
        compact_rows(need, P.cap, P, grow, thr, cnt, my_buf, hist, lane);
        if (tracing) g_trace_ns[blockIdx.x * kTraceTiles + kTraceTiles - 1] = globaltimer_ns();
        P.unit_cnt[(size_t)u * C::UNIT_ROWS + row_in_unit] = cnt;
        P.unit_thr[(size_t)u * C::UNIT_ROWS + row_in_unit] = row_ok ? float_to_key(thr) : 0u;
      }
    }
  }

  // ------------------------------ teardown ------------------------------
  ptx::tc_fence_before();
  if (CTAS == 2) ptx::cluster_sync_all(); else __syncthreads();
  if (warp == kWarpAlloc) {
    ptx::tc_fence_after();
    if (CTAS == 2) ptx::tmem_dealloc_pair(tmem_base, TMEM_COLS);
    else ptx::tmem_dealloc(tmem_base, TMEM_COLS);
  }
}


// ==========================================================================================
// Query-stationary variant (CTA pairs, D_pad <= 768): the pair's 256-query block is loaded ONCE per
// work unit and stays on chip for the whole stripe; only corpus rows stream through the smem ring.
// Against the kernel above this halves the operand traffic L2 -> SM per FLOP (768 KB -> 384 KB per
// 256 x 256 block of scores at D = 768) and the shared-memory port load (no query tile is written
// and re-read per corpus tile), and the producer issues ONE 32 KB box per 1024 cycles of MMA work.
//
// Where the query block lives (per CTA: 128 rows x D_pad bf16 = up to 192 KB):
//   k-blocks [0, 8)      TENSOR MEMORY, columns [0, 256): lane r = query row r, column c = bf16
//                        elements (2c, 2c+1); the MMAs read it as their A operand from TMEM
//   k-blocks [8, 12)     shared memory, 128B-swizzled K-major like a TMA-staged tile (64 KB), read
//                        through a descriptor; loaded by one TMA box per unit
// TMEM columns [256, 512): two fp32 accumulators of ACC_N = 128 corpus rows, double-buffered (N = 128
// is the smallest N that runs at the full MMA rate: 64.0 cycles per instruction, N = 64 takes 44.6
// instead of 32 -- profiles/ubench_umma.cu).  A plan tile (256 corpus rows) is walked as two
// sub-tiles; a sub-tile's MMAs run over all k-blocks before the next one starts.
// smem ring: 4 stages of QS_KB_STAGE = 4 k-blocks of ONE CTA's half of a sub-tile (64 rows): one 3-D
// TMA box {64 k, 64 rows, 4 k-blocks} = 32 KB per stage.
// ==========================================================================================
constexpr int QS_KB_STAGE = 4;
constexpr int QS_MAX_DPAD = 768;
constexpr int QS_KB_TMEM = 8;                                   // k-blocks of the query block kept in TMEM
constexpr int QS_ACC_N = 128;
constexpr int QS_STAGES = 4;
constexpr int QS_B_ROWS = QS_ACC_N / 2;                          // corpus rows one CTA stages per sub-tile
constexpr uint32_t QS_KB_BYTES = QS_B_ROWS * BK * 2;             // 8 KB: one k-block of that half
constexpr uint32_t QS_STAGE_BYTES = QS_KB_BYTES * QS_KB_STAGE;   // 32 KB
constexpr uint32_t QS_ATAIL_BYTES = (QS_MAX_DPAD / BK - QS_KB_TMEM) * A_BYTES;   // 64 KB
constexpr size_t QS_SMEM = (size_t)QS_ATAIL_BYTES + (size_t)QS_STAGES * QS_STAGE_BYTES + 1024;
constexpr uint32_t QS_ACC_BASE = TMEM_COLS - 2 * QS_ACC_N;
constexpr int QS_SUB = BN / QS_ACC_N;                            // sub-tiles per plan tile

// This thread's query row (or zeros) -> its TMEM lane, 32 columns (= 64 bf16 = 8 x 16 bytes) per store.
__device__ __forceinline__ void qs_load_query_row(const uint4* __restrict__ qrow, bool ok, uint32_t taddr, int a_cols) {
#pragma unroll 2
  for (int c0 = 0; c0 < a_cols; c0 += 32) {
    uint32_t v[32];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const uint4 x = ok ? __ldg(qrow + (c0 >> 2) + j) : make_uint4(0u, 0u, 0u, 0u);
      v[4 * j] = x.x; v[4 * j + 1] = x.y; v[4 * j + 2] = x.z; v[4 * j + 3] = x.w;
    }
    ptx::tmem_st_32x32(taddr + (uint32_t)c0, v);
  }
  ptx::tmem_st_wait();
}

template <bool DENSE>
__global__ void __launch_bounds__(kScoreThreads, 1)
score_select_qs_kernel(const __grid_constant__ CUtensorMap tmap_c, const __grid_constant__ CUtensorMap tmap_q,
                       const ScoreParams P) {
  constexpr int ACC_N = QS_ACC_N;
  constexpr int NCHUNK = ACC_N / 32;
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t s_full[QS_STAGES];
  __shared__ __align__(8) uint64_t s_empty[QS_STAGES];
  __shared__ __align__(8) uint64_t s_tmem_full[2];
  __shared__ __align__(8) uint64_t s_tmem_empty[2];
  __shared__ __align__(8) uint64_t s_a_ready;
  __shared__ uint32_t s_tmem_base;
  __shared__ int s_hist[4][256];
  __shared__ float s_stage[4][32 * kStagePitch];

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  constexpr int kWarpTma = 4, kWarpMma = 5, kWarpAlloc = 6;
  const uint32_t rank = ptx::cluster_ctarank();
  const int group = blockIdx.x / 2, n_groups = gridDim.x / 2;
  const uint32_t smem_atail = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;   // query k-blocks >= QS_KB_TMEM
  const uint32_t smem_ring = smem_atail + QS_ATAIL_BYTES;
  const int stages_per_sub = (P.num_kb + QS_KB_STAGE - 1) / QS_KB_STAGE;
  const int n_sub = (P.N + ACC_N - 1) / ACC_N;
  const int kb_tmem = min(P.num_kb, QS_KB_TMEM);
  const int kb_tail = P.num_kb - kb_tmem;        // 0..4 k-blocks of the query block in shared memory

  if (warp == kWarpTma && lane == 0) { ptx::prefetch_tmap(&tmap_c); ptx::prefetch_tmap(&tmap_q); }
  if (warp == kWarpMma && lane == 0) {
    for (int s = 0; s < QS_STAGES; ++s) {
      ptx::mbar_init(ptx::smem_u32(&s_full[s]), 1);
      ptx::mbar_init(ptx::smem_u32(&s_empty[s]), 1);
    }
    for (int a = 0; a < 2; ++a) {
      ptx::mbar_init(ptx::smem_u32(&s_tmem_full[a]), 1);
      ptx::mbar_init(ptx::smem_u32(&s_tmem_empty[a]), 8);   // 4 epilogue warps of each CTA
    }
    // query block in place: one arrival per epilogue warp of both CTAs (the leader's warp 0 arrives
    // with the byte count of both CTAs' smem parts)
    ptx::mbar_init(ptx::smem_u32(&s_a_ready), 8);
    ptx::fence_mbar_init();
  }
  if (warp == kWarpAlloc) ptx::tmem_alloc_pair(ptx::smem_u32(&s_tmem_base), TMEM_COLS);
  ptx::tc_fence_before();
  ptx::cluster_sync_all();
  ptx::tc_fence_after();
  const uint32_t tmem_base = s_tmem_base;

  if (warp == kWarpTma) {
    // ============================== TMA producer (corpus rows only; whole warp, one lane issues) ===
    uint32_t stage = 0, phase = 0;
    for (int u = group; u < P.units; u += n_groups) {
      const int s = u / P.m_tiles;
      const int sub0 = s * P.tiles_per_stripe * QS_SUB;
      const int sub1 = min(min(s * P.tiles_per_stripe + P.tiles_per_stripe, P.n_tiles) * QS_SUB, n_sub);
      for (int sub = sub0; sub < sub1; ++sub) {
        const int c_row = (P.debug & 4) ? 0 : sub * ACC_N + (int)rank * QS_B_ROWS;
        for (int st = 0; st < stages_per_sub; ++st) {
          ptx::mbar_wait(ptx::smem_u32(&s_empty[stage]), phase ^ 1u);
          if (ptx::elect_one_sync()) {
            const uint32_t full = ptx::smem_u32(&s_full[stage]);
            // k-blocks past the end of the row (last stage of a sub-tile when num_kb % 4 != 0) are
            // zero-filled by TMA and still counted: the byte count of a stage is always the whole box
            if (rank == 0) ptx::mbar_arrive_expect_tx(full, QS_STAGE_BYTES * 2u);
            ptx::tma_load_3d_pair(smem_ring + stage * QS_STAGE_BYTES, &tmap_c, ptx::mapa_shared(full, 0), 0, c_row,
                                  st * QS_KB_STAGE, ptx::kEvictNormal);
          }
          __syncwarp();
          if (++stage == QS_STAGES) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == kWarpMma && rank == 0) {
    // ============================== MMA issuer (whole warp, one lane issues) ================================
    constexpr uint32_t idesc = ptx::make_idesc_bf16_f32(BM * 2, ACC_N);
    const uint64_t da_tail = ptx::make_sw128_kmajor_desc(smem_atail);
    const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);   // warp-uniform for the compiler
    uint32_t stage = 0, phase = 0, acc = 0, acc_phase = 0, a_phase = 0;
    for (int u = group; u < P.units; u += n_groups) {
      const int s = u / P.m_tiles;
      const int sub0 = s * P.tiles_per_stripe * QS_SUB;
      const int sub1 = min(min(s * P.tiles_per_stripe + P.tiles_per_stripe, P.n_tiles) * QS_SUB, n_sub);
      // the unit's query block is in TMEM / shared memory of both CTAs
      ptx::mbar_wait(ptx::smem_u32(&s_a_ready), a_phase);
      a_phase ^= 1u;
      ptx::tc_fence_after();
      for (int sub = sub0; sub < sub1; ++sub) {
        ptx::mbar_wait(ptx::smem_u32(&s_tmem_empty[acc]), acc_phase ^ 1u);
        ptx::tc_fence_after();
        const uint32_t d_tmem = tmem_u + QS_ACC_BASE + acc * ACC_N;
        for (int st = 0; st < stages_per_sub; ++st) {
          const int nkb = min(QS_KB_STAGE, P.num_kb - st * QS_KB_STAGE);
          ptx::mbar_wait(ptx::smem_u32(&s_full[stage]), phase);
          ptx::tc_fence_after();
          if (ptx::elect_one_sync()) {
            const uint64_t db = ptx::make_sw128_kmajor_desc(smem_ring + stage * QS_STAGE_BYTES);
            // a stage lies entirely on one side of the TMEM / smem split of the query block
            // (QS_KB_TMEM is a multiple of QS_KB_STAGE): ONE decision per 16 MMAs, and every operand
            // of the unrolled MMAs is a base plus a compile-time offset
            if (st * QS_KB_STAGE < QS_KB_TMEM) {
              const uint32_t a0 = tmem_u + (uint32_t)(st * QS_KB_STAGE * (BK / 2));
#pragma unroll
              for (int kbi = 0; kbi < QS_KB_STAGE; ++kbi) {
                if (kbi < nkb) {
#pragma unroll
                  for (int k = 0; k < BK / UMMA_K; ++k)   // A: 16 bf16 = 8 TMEM columns per K step
                    ptx::umma_bf16_pair_ts(d_tmem, a0 + (uint32_t)(kbi * (BK / 2) + k * (UMMA_K / 2)),
                                           db + (uint64_t)(kbi * (QS_KB_BYTES >> 4) + 2 * k), idesc,
                                           (kbi | k) != 0 ? 1u : (st != 0 ? 1u : 0u));
                }
              }
            } else {
              const uint64_t a0 = da_tail + (uint64_t)((st * QS_KB_STAGE - QS_KB_TMEM) * (A_BYTES >> 4));
#pragma unroll
              for (int kbi = 0; kbi < QS_KB_STAGE; ++kbi) {
                if (kbi < nkb) {
#pragma unroll
                  for (int k = 0; k < BK / UMMA_K; ++k)   // A: one 16 KB swizzled tile per k-block, +32 bytes per K step
                    ptx::umma_bf16_pair(d_tmem, a0 + (uint64_t)(kbi * (A_BYTES >> 4) + 2 * k),
                                        db + (uint64_t)(kbi * (QS_KB_BYTES >> 4) + 2 * k), idesc, 1u);
                }
              }
            }
            ptx::umma_commit_pair(ptx::smem_u32(&s_empty[stage]), 3);
            if (st + 1 == stages_per_sub) ptx::umma_commit_pair(ptx::smem_u32(&s_tmem_full[acc]), 3);
          }
          __syncwarp();
          if (++stage == QS_STAGES) { stage = 0; phase ^= 1u; }
        }
        acc ^= 1u;
        if (acc == 0) acc_phase ^= 1u;
      }
    }
  } else if (warp < 4) {
    // ============================== epilogue ==================================
    const int quarter = warp;
    const int row_in_unit = (int)rank * BM + quarter * 32 + lane;
    int* hist = s_hist[quarter];
    float* stage = s_stage[quarter];
    const uint32_t lane_base = tmem_base + ((uint32_t)(quarter * 32) << 16);
    const uint32_t a_ready_leader = ptx::mapa_shared(ptx::smem_u32(&s_a_ready), 0);
    const int a_cols = kb_tmem * (BK / 2);
    // Puts the query block of unit `un` in place (this thread: its own row -> TMEM; warp 0 lane 0:
    // the TMA box of the k-blocks that live in shared memory) and arrives on the leader's barrier.
    auto stage_queries = [&](int un) {
      const int gn = (un % P.m_tiles) * (2 * BM) + row_in_unit;
      if (P.debug & 128) {   // ablation: no staging work, only the handshake
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive_cluster_relaxed(a_ready_leader);
        return;
      }
      if (quarter == 0 && lane == 0 && kb_tail > 0) {
        // both CTAs' boxes are accounted on the leader's barrier; rows past Q are zero-filled
        ptx::tma_load_3d_pair(smem_atail, &tmap_q, a_ready_leader, 0, (un % P.m_tiles) * (2 * BM) + (int)rank * BM, QS_KB_TMEM,
                              ptx::kEvictLast);
      }
      qs_load_query_row(P.q_rows + (size_t)gn * P.q_pitch16, gn < P.Q, lane_base, a_cols);
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (quarter == 0 && rank == 0 && kb_tail > 0) ptx::mbar_arrive_expect_tx(ptx::smem_u32(&s_a_ready), 2u * QS_ATAIL_BYTES);
        else ptx::mbar_arrive_cluster_relaxed(a_ready_leader);
      }
    };
    uint32_t acc = 0, acc_phase = 0;
    bool a_stored = false;
    for (int u = group; u < P.units; u += n_groups) {
      const int s = u / P.m_tiles, m = u - s * P.m_tiles;
      const int sub0 = s * P.tiles_per_stripe * QS_SUB;
      const int sub1 = min(min(s * P.tiles_per_stripe + P.tiles_per_stripe, P.n_tiles) * QS_SUB, n_sub);
      const int grow = m * (2 * BM) + row_in_unit;
      if (!a_stored) {   // first unit of this CTA (later ones are staged at the end of the previous unit)
        stage_queries(u);
        a_stored = true;
      }
      const bool row_ok = grow < P.Q && !(P.debug & 16);
      float thr = row_ok ? -INFINITY : INFINITY;
      float pub = thr;
      float cm[kChunkMax];
#pragma unroll
      for (int i = 0; i < kChunkMax; ++i) cm[i] = -INFINITY;
      int cnt = 0;
      uint2* my_buf = DENSE ? nullptr : P.unit_cand + ((size_t)u * (2 * BM) + row_in_unit) * (size_t)P.cap;
      uint32_t next_hint = (!DENSE && row_ok) ? __ldcg(&P.thr_hint[grow]) : 0u;
      bool unit_cold = false;
      for (int sub = sub0; sub < sub1; ++sub) {
        const bool tile_start = ((sub - sub0) % QS_SUB) == 0;
        if (!DENSE && row_ok && tile_start) {
          if (next_hint != 0u) thr = fmaxf(thr, key_to_float(next_hint));
          next_hint = __ldcg(&P.thr_hint[grow]);
        }
        if (sub == sub0) unit_cold = __ballot_sync(0xffffffffu, row_ok && thr == -INFINITY) != 0u;
        ptx::mbar_wait(ptx::smem_u32(&s_tmem_full[acc]), acc_phase);
        ptx::tc_fence_after();
        const uint32_t taddr = lane_base + QS_ACC_BASE + acc * ACC_N;
        const bool last_of_unit = sub + 1 == sub1;
        if (P.debug & 1) {  // ablation: no TMEM reads at all
          ptx::tc_fence_before();
          __syncwarp();
          if (lane == 0) ptx::mbar_arrive_cluster_relaxed(ptx::mapa_shared(ptx::smem_u32(&s_tmem_empty[acc]), 0));
        } else {
          // Cold start (first sub-tile of a unit whose rows have no threshold yet): pass 1 collects
          // each chunk's four best scores into the running top-20, pass 2 is the normal filter against
          // the resulting threshold, so the unit appends ~20 entries per row instead of the sub-tile.
          bool update_cm = P.chunkmax != 2 || unit_cold;   // see the classic kernel
          if (!DENSE && P.chunkmax && sub == sub0 && unit_cold) {
            uint32_t vc[32];
#pragma unroll 1
            for (int chunk = 0; chunk < NCHUNK; ++chunk) {
              ptx::tmem_ld_32x32(taddr + chunk * 32, vc);
              ptx::tmem_ld_wait(vc);
              const int col0 = sub * ACC_N + chunk * 32;
              // eight best of the chunk (sorted insertion), then those into cm[]: 20 of the 32
              // documents collected over the sub-tile score >= cm[19]
              float c8[8];
#pragma unroll
              for (int i = 0; i < 8; ++i) c8[i] = -INFINITY;
#pragma unroll
              for (int j = 0; j < 32; ++j) {
                const float x = (col0 + j < P.N) ? __uint_as_float(vc[j]) : -INFINITY;
#pragma unroll
                for (int i = 7; i > 0; --i) c8[i] = fmaxf(c8[i], fminf(c8[i - 1], x));
                c8[0] = fmaxf(c8[0], x);
              }
#pragma unroll
              for (int i = 0; i < 8; ++i) cm_insert(cm, c8[i]);
            }
            if (row_ok) thr = fmaxf(thr, cm[kChunkMax - 1]);
            update_cm = false;
          }
          uint32_t va[32], vb[32];
          ptx::tmem_ld_32x32(taddr, va);
          ptx::tmem_ld_wait(va);
#pragma unroll 1
          for (int chunk = 0; chunk < NCHUNK; chunk += 2) {
            ptx::tmem_ld_32x32(taddr + (chunk + 1) * 32, vb);
            epilogue_chunk<DENSE>(va, sub * ACC_N + chunk * 32, P, grow, row_ok, thr, cnt, my_buf, stage, hist, lane, cm,
                                  update_cm);
            ptx::tmem_ld_wait(vb);
            if (chunk + 2 < NCHUNK) {
              ptx::tmem_ld_32x32(taddr + (chunk + 2) * 32, va);
            } else {
              // every column of this accumulator is in registers: hand it back to the MMA warp
              ptx::tc_fence_before();
              __syncwarp();
              if (lane == 0) ptx::mbar_arrive_cluster_relaxed(ptx::mapa_shared(ptx::smem_u32(&s_tmem_empty[acc]), 0));
            }
            epilogue_chunk<DENSE>(vb, sub * ACC_N + (chunk + 1) * 32, P, grow, row_ok, thr, cnt, my_buf, stage, hist, lane,
                                  cm, update_cm);
            if (chunk + 2 < NCHUNK) ptx::tmem_ld_wait(va);
          }
        }
        acc ^= 1u;
        if (acc == 0) acc_phase ^= 1u;
        if (!DENSE && row_ok && thr > pub && (((sub + 1 - sub0) % QS_SUB) == 0 || last_of_unit)) {
          publish_threshold(P, grow, float_to_key(thr));
          pub = thr;
        }
      }
      // Every accumulator of this unit has been read, so all of its MMAs have completed and the query
      // block (TMEM and smem part) is free: stage the NEXT unit's block first -- the tensor pipe
      // restarts while this unit's buffers are being compacted below.
      if (u + n_groups < P.units) stage_queries(u + n_groups);
      if (!DENSE) {
        unsigned need = __ballot_sync(0xffffffffu, cnt > P.kunit);
        if (P.cap <= kSmallCap && P.chunkmax) need = final_filter_rows(need, P.kunit + 16, thr, cnt, my_buf, lane);
        compact_rows(need, P.cap, P, grow, thr, cnt, my_buf, hist, lane);
        P.unit_cnt[(size_t)u * (2 * BM) + row_in_unit] = cnt;
        P.unit_thr[(size_t)u * (2 * BM) + row_in_unit] = row_ok ? float_to_key(thr) : 0u;
      }
    }
  }

  // ------------------------------ teardown ------------------------------
  ptx::tc_fence_before();
  ptx::cluster_sync_all();
  if (warp == kWarpAlloc) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc_pair(tmem_base, TMEM_COLS);
  }
}

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode_fn() {
  static PFN_encodeTiled fn = nullptr;
  if (fn) return fn;
  void* p = nullptr;
  cudaDriverEntryPointQueryResult qres;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) != cudaSuccess ||
      qres != cudaDriverEntryPointSuccess || p == nullptr) {
    cudaGetLastError();
    return nullptr;
  }
  fn = reinterpret_cast<PFN_encodeTiled>(p);
  return fn;
}

// bf16 [rows, d_pad] row-major -> tiles of box_rows x 64 columns, 128B swizzle, zero OOB fill.
static int make_tmap(CUtensorMap* map, const void* base, int64_t rows, int64_t d_pad, int box_rows) {
  PFN_encodeTiled enc = get_encode_fn();
  if (!enc) { set_error("cuTensorMapEncodeTiled entry point not available"); return QST_ERR_CUDA; }
  cuuint64_t gdim[2] = {(cuuint64_t)d_pad, (cuuint64_t)rows};
  cuuint64_t gstride[1] = {(cuuint64_t)d_pad * 2};
  cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), gdim, gstride, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed with CUresult %d", (int)r); return QST_ERR_CUDA; }
  return QST_OK;
}

template <int CTAS, bool DENSE>
static int launch_score_t(const void* q_bf16, const void* c_bf16, const ScoreParams& P, int64_t d_pad, int groups,
                          cudaStream_t st) {
  using C = Cfg<CTAS>;
  CUtensorMap tq, tc;
  int rc = make_tmap(&tq, q_bf16, P.Q, d_pad, BM);
  if (rc) return rc;
  rc = make_tmap(&tc, c_bf16, P.N, d_pad, C::B_ROWS);
  if (rc) return rc;
  auto kern = score_select_kernel<CTAS, DENSE>;
  QST_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM));
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)(groups * CTAS));
  cfg.blockDim = dim3(kScoreThreads);
  cfg.dynamicSmemBytes = C::SMEM;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CTAS;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  QST_CUDA(cudaLaunchKernelEx(&cfg, kern, tq, tc, P));
  count_launch();
  return QST_OK;
}

// bf16 [rows, d_pad] row-major seen as {64 k, rows, d_pad/64 k-blocks}: one box = box_rows rows x
// box_kb k-blocks, written to smem as box_kb consecutive 128B-swizzled [box_rows x 64] tiles.  Rows and
// k-blocks out of range are zero-filled.
static int make_tmap_3d(CUtensorMap* map, const void* base, int64_t rows, int64_t d_pad, int box_rows, int box_kb) {
  PFN_encodeTiled enc = get_encode_fn();
  if (!enc) { set_error("cuTensorMapEncodeTiled entry point not available"); return QST_ERR_CUDA; }
  cuuint64_t gdim[3] = {(cuuint64_t)BK, (cuuint64_t)rows, (cuuint64_t)(d_pad / BK)};
  cuuint64_t gstride[2] = {(cuuint64_t)d_pad * 2, (cuuint64_t)BK * 2};
  cuuint32_t box[3] = {(cuuint32_t)BK, (cuuint32_t)box_rows, (cuuint32_t)box_kb};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), gdim, gstride, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled (3-D) failed with CUresult %d", (int)r); return QST_ERR_CUDA; }
  return QST_OK;
}

template <bool DENSE>
static int launch_score_qs_t(const void* q_bf16, const void* c_bf16, ScoreParams P, int64_t d_pad, int groups,
                             cudaStream_t st) {
  CUtensorMap tc, tq;
  int rc = make_tmap_3d(&tc, c_bf16, P.N, d_pad, QS_B_ROWS, QS_KB_STAGE);
  if (rc) return rc;
  // query k-blocks [QS_KB_TMEM, 12) of one CTA's 128 rows in one box
  rc = make_tmap_3d(&tq, q_bf16, P.Q, d_pad, BM, QS_MAX_DPAD / BK - QS_KB_TMEM);
  if (rc) return rc;
  P.q_rows = reinterpret_cast<const uint4*>(q_bf16);
  P.q_pitch16 = (int)(d_pad / 8);
  auto kern = score_select_qs_kernel<DENSE>;
  QST_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)QS_SMEM));
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)(groups * 2));
  cfg.blockDim = dim3(kScoreThreads);
  cfg.dynamicSmemBytes = QS_SMEM;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  QST_CUDA(cudaLaunchKernelEx(&cfg, kern, tc, tq, P));
  count_launch();
  return QST_OK;
}

static int launch_score_qs(bool dense, const void* q_bf16, const void* c_bf16, const ScoreParams& P_in, int64_t d_pad,
                           int groups, cudaStream_t st) {
  ScoreParams P = P_in;
  const char* dbg = getenv("QST_SCORE_DEBUG");
  P.debug = dbg ? atoi(dbg) : 0;
  return dense ? launch_score_qs_t<true>(q_bf16, c_bf16, P, d_pad, groups, st)
               : launch_score_qs_t<false>(q_bf16, c_bf16, P, d_pad, groups, st);
}

// Query-stationary tiles for CTA pairs whenever the query block fits TMEM next to two accumulators;
// QST_SCORE_QS=0 forces the classic (query tiles re-staged per corpus tile) pair kernel.
static bool use_query_stationary(int ctas, int64_t d_pad) {
  const char* e = getenv("QST_SCORE_QS");
  if (e && e[0] == '0') return false;
  return ctas == 2 && d_pad <= QS_MAX_DPAD;
}

static int launch_score(int ctas, bool dense, const void* q_bf16, const void* c_bf16, const ScoreParams& P_in,
                        int64_t d_pad, int groups, cudaStream_t st) {
  // QST_SCORE_DEBUG: performance ablations only (results are wrong when set): 1 = epilogue skips
  // TMEM reads, 2 = epilogue reads TMEM but does not select, 4 = every tile loads corpus rows 0..255,
  // 8 = one MMA per k-block instead of four, 16 = thresholds at +inf (fast path only).
  ScoreParams P = P_in;
  const char* dbg = getenv("QST_SCORE_DEBUG");
  P.debug = dbg ? atoi(dbg) : 0;
  if (ctas == 2) {
    return dense ? launch_score_t<2, true>(q_bf16, c_bf16, P, d_pad, groups, st)
                 : launch_score_t<2, false>(q_bf16, c_bf16, P, d_pad, groups, st);
  }
  return dense ? launch_score_t<1, true>(q_bf16, c_bf16, P, d_pad, groups, st)
               : launch_score_t<1, false>(q_bf16, c_bf16, P, d_pad, groups, st);
}

// QST_SCORE_CTAS=1 forces the single-CTA tile (debugging / comparison); default is CTA pairs.
// CTA pairs (M = 256) unless the whole query batch fits one 128-row tile: a pair would then spend
// half of its tensor time on empty rows, and the pass is corpus-streaming bound (148 independent
// single-CTA walkers pull more HBM bandwidth than 74 pairs).  QST_SCORE_CTAS=1|2 overrides.
// (profiles/small_q_probe.py; for Q = 160..768 the two shapes are otherwise within 2 % of each other)
static int default_ctas(int64_t Q) {
  const char* e = getenv("QST_SCORE_CTAS");
  if (e && e[0] == '1') return 1;
  if (e && e[0] == '2') return 2;
  if (Q <= BM) return 1;
  // small batches whose last 256-row block would be at most half full (Q = 384: 0.50 ms single
  // against 0.62 ms in pairs): pairs pad the batch to a multiple of 256 rows, single tiles to 128
  if (Q <= 2048 && ((Q - 1) % (2 * BM)) < BM) return 1;
  return 2;
}

static int device_sm_count() {
  int dev = 0, sms = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 0;
  if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return 0;
  return sms;
}

}  // namespace qst

using namespace qst;

// workspace layout of a plan (depends on units, rows_per_unit, cap)
static void plan_layout(qst_topk_plan* plan) {
  size_t off = 0;
  const size_t ur = (size_t)plan->rows_per_unit;
  plan->off_thr = off;  off += round_up((size_t)plan->m_tiles * ur * sizeof(uint32_t), 256);
  plan->off_cnt = off;  off += round_up((size_t)plan->units * ur * sizeof(int), 256);
  plan->off_uthr = off; off += round_up((size_t)plan->units * ur * sizeof(uint32_t), 256);
  plan->off_cand = off; off += (size_t)plan->units * ur * (size_t)plan->cap * sizeof(uint2);
  plan->ws_bytes = off;
}

extern "C" int qst_topk_plan_make(int64_t Q, int64_t N, int64_t D, int k, int kprime, int score, int sm_count,
                                  qst_topk_plan* plan) {
  QST_CHECK_ARG(plan != nullptr, "plan_make: null plan");
  QST_CHECK_ARG(Q >= 1 && N >= 1 && D >= 1, "plan_make: bad shape Q=%lld N=%lld D=%lld", (long long)Q, (long long)N,
                (long long)D);
  QST_CHECK_ARG(Q < (1ll << 31) - BM && N < (1ll << 31) - BN, "plan_make: Q and N must fit in int32");
  QST_CHECK_ARG(k >= 1 && k <= 1024, "plan_make: k must be in [1, 1024], %d given", k);
  QST_CHECK_ARG(score >= QST_SCORE_COS && score <= QST_SCORE_EUCLID, "plan_make: unknown score function %d", score);
  if (sm_count <= 0) sm_count = device_sm_count();
  if (sm_count <= 0) sm_count = 148;
  if (kprime <= 0) {
    // head-room so that the bf16 ordering error cannot push a true top-k document out of the
    // rescored set (DESIGN.md "certificate"): k + max(1.2 k, 40), in steps of 16.  Measured on
    // config 3: k' = 192 leaves 1 query in 20 000 uncertified, k' = 224 none.
    int extra = (6 * k) / 5 > 40 ? (6 * k) / 5 : 40;
    kprime = k + extra;
  }
  kprime = (int)round_up(kprime, 16);
  if (kprime > 2048) kprime = 2048;
  QST_CHECK_ARG(kprime >= k && kprime <= 2048, "plan_make: kprime %d out of range [k, 2048]", kprime);
  memset(plan, 0, sizeof(*plan));
  plan->Q = Q; plan->N = N; plan->D = D; plan->D_pad = qst_padded_dim_for(D, score == QST_SCORE_EUCLID ? QST_PREP_EUCLID_CORPUS : QST_PREP_RAW);
  plan->k = k; plan->kprime = kprime;
  plan->score = score;
  plan->ctas = default_ctas(Q);
  plan->qs = use_query_stationary(plan->ctas, plan->D_pad) ? 1 : 0;
  plan->rows_per_unit = BM * plan->ctas;
  plan->m_tiles = (int)ceil_div(Q, plan->rows_per_unit);
  plan->n_tiles = (int)ceil_div(N, BN);
  const int groups_max = sm_count / plan->ctas > 0 ? sm_count / plan->ctas : 1;
  // stripes: minimise  waves * (tiles_per_stripe + per-unit cost)  over S, subject to a stripe
  // being small enough (<= 32 MB of bf16 rows) that the few stripes being walked at any time stay
  // resident in the 126 MB L2 even when their walkers drift apart
  const int r_min = 4;
  int s_max = plan->n_tiles / r_min;
  if (s_max < 1) s_max = 1;
  if (s_max > 160) s_max = 160;   // finalize.cu kMaxStripes
  const int64_t tile_bytes = (int64_t)BN * plan->D_pad * 2;
  int64_t r_l2 = (32ll << 20) / tile_bytes;
  if (r_l2 < r_min) r_l2 = r_min;
  int s_min = (int)ceil_div(plan->n_tiles, r_l2);
  if (s_min > s_max) s_min = s_max;
  double best = 1e300;
  int best_s = 1;
  for (int S = s_min; S <= s_max; ++S) {
    const int R = (int)ceil_div(plan->n_tiles, S);
    const int S_eff = (int)ceil_div(plan->n_tiles, R);  // stripes actually non-empty
    const int64_t units = (int64_t)plan->m_tiles * S_eff;
    const int64_t waves = ceil_div(units, groups_max);
    const double cost = (double)waves * ((double)R + 2.0);
    if (cost < best - 1e-9) { best = cost; best_s = S_eff; }
  }
  {
    const char* e = getenv("QST_STRIPES");   // tuning override
    if (e && atoi(e) >= 1) best_s = atoi(e) < plan->n_tiles ? atoi(e) : plan->n_tiles;
  }
  plan->tiles_per_stripe = (int)ceil_div(plan->n_tiles, best_s);
  plan->stripes = (int)ceil_div(plan->n_tiles, plan->tiles_per_stripe);
  plan->units = plan->m_tiles * plan->stripes;
  // Entries a unit keeps per row.  The certificate only needs every unit's final threshold to stay
  // below the k'-th best score overall, so with S stripes a unit needs ~k'/S entries plus slack
  // for uneven placement; QST_KUNIT overrides (kunit = kprime is the most conservative setting).
  {
    // Poisson tail: a stripe holds ~k'/S of the k' best documents; 3x that plus 16 keeps the chance
    // that a unit's threshold climbs above the k'-th best score negligible (at 28 stripes and
    // k' = 192, kunit = 16 left 56 of 20 000 queries uncertified, kunit = 32 none)
    int ku = (int)round_up(3 * (int)ceil_div(kprime, plan->stripes) + 8, 8);
    if (ku < 16) ku = 16;
    if (ku > kprime) ku = kprime;
    const char* e = getenv("QST_KUNIT");
    if (e && atoi(e) >= 8) { ku = (int)round_up(atoi(e), 8); if (ku > kprime) ku = kprime; }
    plan->kunit = ku;
    // slack between compactions: at least 128 entries; the whole register-resident compaction
    // window (256) when the unit keeps few, so that a unit rarely compacts before its end
    plan->cap = 2 * ku > ku + 128 ? 2 * ku : ku + 128;
    if (ku <= 64 && plan->cap < 256) plan->cap = 256;
  }
  plan->grid = plan->units < groups_max ? plan->units : groups_max;  // CTA groups (x ctas CTAs)
  {
    // QST_K2_GROUPS: leave SMs free for kernels of a neighbouring batch running on another stream
    // (spatial overlap experiments: K2 is persistent, nothing else becomes resident on an SM it holds)
    const char* e = getenv("QST_K2_GROUPS");
    if (e && atoi(e) >= 1 && atoi(e) < plan->grid) plan->grid = atoi(e);
  }
  plan_layout(plan);
  return QST_OK;
}

extern "C" int qst_topk_plan_set_kunit(qst_topk_plan* plan, int kunit) {
  QST_CHECK_ARG(plan != nullptr, "plan_set_kunit: null plan");
  QST_CHECK_ARG(kunit >= 8 && kunit <= 2048, "plan_set_kunit: kunit=%d out of range", kunit);
  plan->kunit = (int)round_up(kunit, 8);
  plan->cap = 2 * plan->kunit > plan->kunit + 128 ? 2 * plan->kunit : plan->kunit + 128;
  if (plan->kunit <= 64 && plan->cap < 256) plan->cap = 256;
  plan_layout(plan);
  return QST_OK;
}


static int score_select_impl(const qst_topk_plan* plan, const void* q_bf16, const void* c_bf16, void* workspace,
                             uint32_t* hint_local, uint32_t* const* peer_hints, int n_peers, qst_stream_t stream);

extern "C" int qst_score_select(const qst_topk_plan* plan, const void* q_bf16, const void* c_bf16, void* workspace,
                                qst_stream_t stream) {
  return score_select_impl(plan, q_bf16, c_bf16, workspace, nullptr, nullptr, 0, stream);
}

extern "C" int qst_score_select_peers(const qst_topk_plan* plan, const void* q_bf16, const void* c_bf16,
                                      void* workspace, void* hint_local, void* const* peer_hints, int n_peers,
                                      qst_stream_t stream) {
  QST_CHECK_ARG(hint_local != nullptr, "score_select_peers: null hint_local");
  QST_CHECK_ARG(n_peers >= 0 && n_peers <= QST_MAX_PEERS, "score_select_peers: n_peers=%d out of range", n_peers);
  QST_CHECK_ARG(n_peers == 0 || peer_hints != nullptr, "score_select_peers: null peer_hints");
  return score_select_impl(plan, q_bf16, c_bf16, workspace, reinterpret_cast<uint32_t*>(hint_local),
                           reinterpret_cast<uint32_t* const*>(peer_hints), n_peers, stream);
}

static int score_select_impl(const qst_topk_plan* plan, const void* q_bf16, const void* c_bf16, void* workspace,
                             uint32_t* hint_local, uint32_t* const* peer_hints, int n_peers, qst_stream_t stream) {
  QST_CHECK_ARG(plan && q_bf16 && c_bf16 && workspace, "score_select: null argument");
  QST_CHECK_ARG((reinterpret_cast<uintptr_t>(q_bf16) & 15u) == 0 && (reinterpret_cast<uintptr_t>(c_bf16) & 15u) == 0,
                "score_select: operands must be 16-byte aligned");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  uint8_t* ws = reinterpret_cast<uint8_t*>(workspace);
  ScoreParams P{};
  P.Q = (int)plan->Q; P.N = (int)plan->N; P.num_kb = (int)(plan->D_pad / BK);
  P.m_tiles = plan->m_tiles; P.n_tiles = plan->n_tiles; P.stripes = plan->stripes;
  P.tiles_per_stripe = plan->tiles_per_stripe; P.units = plan->units;
  P.kunit = plan->kunit; P.cap = plan->cap;
  // the count-20 chunk-maximum thresholds are as safe as the buffer's own when a unit is expected to
  // hold at most ~5 of the k' best documents (kunit = 3*lambda + 8 <= 24)
  P.chunkmax = plan->kunit <= 24 ? 2 : 0;   // 2 = cold units only, 1 = every unit (QST_CHUNKMAX overrides)
  { const char* e = getenv("QST_CHUNKMAX"); if (e) P.chunkmax = atoi(e); }
  // hint array: inside the workspace (zeroed here) or, for sharded runs, the caller's peer-visible
  // buffer, which the CALLER zeroes (it is written by other ranks, see qst_peer_buffer_*)
  P.thr_hint = hint_local ? hint_local : reinterpret_cast<uint32_t*>(ws + plan->off_thr);
  for (int i = 0; i < n_peers; ++i) P.peer_hint[i] = peer_hints[i];
  P.n_peers = n_peers;
  P.unit_cnt = reinterpret_cast<int*>(ws + plan->off_cnt);
  P.unit_thr = reinterpret_cast<uint32_t*>(ws + plan->off_uthr);
  P.unit_cand = reinterpret_cast<uint2*>(ws + plan->off_cand);
  QST_CHECK_ARG(plan->ctas == 1 || plan->ctas == 2, "score_select: plan->ctas must be 1 or 2");
  if (!hint_local)
    QST_CUDA(cudaMemsetAsync(P.thr_hint, 0, (size_t)plan->m_tiles * plan->rows_per_unit * sizeof(uint32_t), st));
  if (plan->qs) {
    QST_CHECK_ARG(plan->ctas == 2 && plan->D_pad <= QS_MAX_DPAD, "score_select: plan->qs needs CTA pairs and D_pad <= %d",
                  QS_MAX_DPAD);
    return launch_score_qs(false, q_bf16, c_bf16, P, plan->D_pad, plan->grid, st);
  }
  return launch_score(plan->ctas, false, q_bf16, c_bf16, P, plan->D_pad, plan->grid, st);
}

extern "C" int qst_debug_read_trace(long long* ns, int* cnt, float* thr, int n) {
  QST_CHECK_ARG(n >= 0 && n <= 160 * kTraceTiles, "debug_read_trace: bad n");
  QST_CUDA(cudaMemcpyFromSymbol(ns, g_trace_ns, sizeof(long long) * n));
  QST_CUDA(cudaMemcpyFromSymbol(cnt, g_trace_cnt, sizeof(int) * n));
  QST_CUDA(cudaMemcpyFromSymbol(thr, g_trace_thr, sizeof(float) * n));
  return QST_OK;
}

extern "C" int qst_score_dense(const void* q_bf16, int64_t Q, const void* c_bf16, int64_t N, int64_t D_pad, float* out,
                               qst_stream_t stream) {
  QST_CHECK_ARG(q_bf16 && c_bf16 && out, "score_dense: null argument");
  QST_CHECK_ARG(Q >= 1 && N >= 1 && D_pad >= BK && D_pad % BK == 0, "score_dense: bad shape");
  ScoreParams P{};
  P.Q = (int)Q; P.N = (int)N; P.num_kb = (int)(D_pad / BK);
  const int ctas = default_ctas(Q);
  P.m_tiles = (int)ceil_div(Q, BM * ctas); P.n_tiles = (int)ceil_div(N, BN);
  P.stripes = 1; P.tiles_per_stripe = P.n_tiles; P.units = P.m_tiles;
  P.dense_out = out;
  int sms = device_sm_count();
  if (sms <= 0) sms = 148;
  const int groups_max = sms / ctas;
  const int groups = P.units < groups_max ? P.units : groups_max;
  if (use_query_stationary(ctas, D_pad))
    return launch_score_qs(true, q_bf16, c_bf16, P, D_pad, groups, reinterpret_cast<cudaStream_t>(stream));
  return launch_score(ctas, true, q_bf16, c_bf16, P, D_pad, groups, reinterpret_cast<cudaStream_t>(stream));
}
