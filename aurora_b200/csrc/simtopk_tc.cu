// Fused similarity + top-k for sm_100a: S = Q . C^T on tcgen05 tensor cores with the
// query block resident in TMEM, the corpus streamed once from HBM by TMA, and a per-query
// candidate list kept in shared memory by the epilogue warps.
//
// Replaces the dense leg of collection.query.hybrid(...) / near_text(...) that the
// reference sends to Weaviate (server/routes/knowledge_base/weaviate_client.py:252-259,
// server/routes/incident_feedback/weaviate_client.py:286-291).
//
// One CTA, persistent, 1 CTA / SM (G = 1 or 2 epilogue groups; 32 * (4G + 3) threads):
//   warps 1..4G epilogue  : thread r of a group owns query r (TMEM lane r): tcgen05.ld its 64 scores of
//                           a tile, scale by the rows' inverse norms (FMUL2), keep one max
//                           per 16 scores and compare it with the query's threshold; a
//                           four-score group that reaches it is parked in a per-thread FIFO
//                           in shared memory and examined later, out of line and rarely
//                           (drain_fifo); large k without room for the FIFO pushes at once
//                           (push_group4).
//   warp 0     threshold   : serves the certified global threshold of the one or two queries
//                           assigned to this CTA (see "Threshold exchange").
//                           With two groups, group g takes tiles g, g+2, ... (TMEM buffer g):
//                           two MMA tile-times per tile, so the MMA rarely waits on them.
//   warp 4G+1  TMA producer: corpus tiles [64 rows x 256 k] -> smem ring (SWIZZLE_128B)
//   warp 4G+2  MMA issuer  : tcgen05.mma kind::f16, A = queries from TMEM (128 lanes = 128
//                           queries, dim/2 columns; dims past 768 from a swizzled tile in
//                           shared memory instead), B = corpus tile from smem, D = [128
//                           queries x 64 rows] fp32 in one of two TMEM buffers.
// (The scheduler favours the highest warp id of a sub-partition: the two latency-critical
// single-thread roles get the top ids, the background threshold warp the bottom one.)
// cta_group::2: a CTA pair shares every corpus tile -- each CTA TMA-loads 32 of the 64
// rows, the leader issues M=256 MMAs, each CTA's TMEM holds its own 128 queries.
// cta_group::1: M=128; when nq > 128 two CTAs take the same tiles for the two query halves.
//
// Threshold exchange.  A CTA sees only 1/74 of the corpus, so its own k-th best is a loose
// filter.  Every epilogue thread therefore publishes its best (or 2nd best) score so far
// into a [query][CTA] table.  Query q is served by CTA q mod 74: its threshold warp reads
// the row of q every couple of microseconds, takes the R-th largest entry (R * m >= k +
// slack) and publishes it: at least k + slack rows with a score >= that value exist
// somewhere, so nothing below it can reach the final top-k.  Every epilogue thread reads its
// query's current threshold once per tile.  A query then admits only a handful of rows per
// CTA over a 1M-row scan.
//
// Bootstrap.  On its first tile a thread only publishes the tile's best score and waits for
// the first certified threshold (all CTAs do this at the same time, ~5 us once), then
// examines the tile against it: no arbitrary rows ever enter a list.
//
// Candidate list.  Append-only, ksel slots per query.  If it fills up it is compacted:
// entries under the current threshold are dropped, and only if ksel live candidates remain
// does it fall back to replace-the-minimum.  At the end the survivors are appended to the
// query's compact row in global memory for the finalize kernel.
#include <cuda.h>
#include "internal.h"
#include "ptx.cuh"

// Bring-up timers compile to nothing unless the library is built with AUR_TC_PROFILE=1: every
// clock64() is a scheduling barrier inside the hot loops.
#ifdef AUR_TC_PROFILE
#define TCLK() clock64()
#else
#define TCLK() 0ll
#endif

namespace aur {
using namespace ptx;

namespace {

constexpr int kThrWarps = 1;
constexpr int kQStageBufs = 2;   // ring stages (the last ones) the query load borrows as its transpose buffer
constexpr uint32_t kSlot = kTcQRows * 8u;   // byte stride between list slots of one query

struct SmemLayout {
  uint32_t stage_bytes, box_bytes, lcap, fifo_recs, qs_kb;
  uint32_t off_qs, off_list, off_norm, off_mask, off_fifo, off_tau, off_bar, total;
};
__host__ __device__ inline SmemLayout make_layout(int cta_group, int epi_groups, int num_stages, int ksel, int dim) {
  SmemLayout L;
  L.box_bytes = (kTcTileN / cta_group) * 128u;
  L.stage_bytes = L.box_bytes * kTcKbPerStage;
  L.lcap = static_cast<uint32_t>(ksel);
  uint32_t o = L.stage_bytes * num_stages;
  // query dims beyond the 768 that fit TMEM: [128 queries x 64] bf16 K-major SWIZZLE_128B tiles, one per
  // 64-dim k-block (the SS-MMA A operand); 1024-byte aligned because the stages are
  L.qs_kb = dim > kTcTmemDim ? static_cast<uint32_t>((dim - kTcTmemDim) / kTcKBlock) : 0u;
  L.off_qs = o;     o += L.qs_kb * (kTcQRows * 128u);
  L.off_list = o;   o += static_cast<uint32_t>(epi_groups) * L.lcap * kSlot;
  L.off_norm = o;   o += static_cast<uint32_t>(epi_groups) * 4u * 2u * kTcTileN * 4u;
  L.off_mask = o;   o += static_cast<uint32_t>(epi_groups) * 4u * 2u * kTcTileN * 4u;   // per-row tenant-scope bit masks (kMask launches)
  // deferred-candidate FIFO: per thread kTcFifoRecs records of four adjacent scores (16 B) + a row tag
  L.fifo_recs = (epi_groups == 1 && ksel <= kTcFifoMaxKsel) ? kTcFifoRecs : 0u;
  L.off_fifo = o;   o += L.fifo_recs * kTcQRows * (16u + 4u);
  L.off_tau = o;    o += kTcQRows * 4u;   // certified thresholds of this CTA's 128 queries, refreshed by the threshold warp
  L.off_bar = o;    o += (2u * kTcMaxStages + 2u + 2u + 2u) * 8u + 16u;
  L.total = o;
  return L;
}

// Per-thread (= per-query) selection state of the epilogue.
struct TopkState {
  uint64_t min_key;   // smallest key, valid when the list holds exactly ksel entries (nfill == ksel)
  float tau_local;    // its score, else -inf
  float tau;          // admission filter = max(tau_local, certified global threshold)
  float top[4];       // best four chunk maxima seen so far (distinct rows), descending: top[xm - 1] is published
  int minpos, nfill;
};

__device__ __forceinline__ void scan_min(uint32_t list_a, int n, uint64_t& m, int& mp) {
  m = lds_u64(list_a); mp = 0;
#pragma unroll 8
  for (int t = 1; t < n; ++t) {
    const uint64_t v = lds_u64(list_a + t * kSlot);
    if (v < m) { m = v; mp = t; }
  }
}

// Drop what the threshold has made useless, then cut down to ksel entries.
__device__ __noinline__ TopkState compact_list(TopkState st, uint32_t list_a, int ksel) {
  const uint32_t thr = f32_to_ord(st.tau);
  int w = 0;
  for (int t = 0; t < st.nfill; ++t) {
    const uint64_t k = lds_u64(list_a + t * kSlot);
    if (static_cast<uint32_t>(k >> 32) >= thr) { sts_u64(list_a + w * kSlot, k); ++w; }
  }
  st.nfill = w;
  while (st.nfill > ksel) {  // rare: this CTA owns more than ksel of the current global best
    uint64_t m; int mp;
    scan_min(list_a, st.nfill, m, mp);
    --st.nfill;
    sts_u64(list_a + mp * kSlot, lds_u64(list_a + st.nfill * kSlot));
  }
  if (st.nfill == ksel) {
    scan_min(list_a, ksel, st.min_key, st.minpos);
    st.tau_local = key_score(st.min_key);
    st.tau = fmaxf(st.tau, st.tau_local);
  }
  return st;
}

// Admit one score into a query's list.
__device__ __forceinline__ TopkState push_one(TopkState st, float s, int row, uint32_t list_a, int ksel, int lcap) {
  const uint64_t key = make_key(s, row);
  if (st.nfill == lcap) st = compact_list(st, list_a, ksel);
  if (st.nfill < lcap) {
    sts_u64(list_a + st.nfill * kSlot, key);
    ++st.nfill;
  } else if (key > st.min_key) {  // lcap == ksel and the list is full of live candidates
    sts_u64(list_a + st.minpos * kSlot, key);
    scan_min(list_a, ksel, st.min_key, st.minpos);
    st.tau_local = key_score(st.min_key);
    st.tau = fmaxf(st.tau, st.tau_local);
  }
  return st;
}

// Examine four adjacent scores of one query (their max reached the threshold).  Out of
// line and small: at warp level some lane needs this about twice per tile, so it has to be
// cheap and its code has to stay resident next to the hot loop.
__device__ __noinline__ TopkState push_group4(TopkState st, float s0, float s1, float s2, float s3, int row,
                                              uint32_t list_a, int ksel) {
  if (s0 >= st.tau) st = push_one(st, s0, row + 0, list_a, ksel, ksel);   // `>=` also rejects NaN
  if (s1 >= st.tau) st = push_one(st, s1, row + 1, list_a, ksel, ksel);
  if (s2 >= st.tau) st = push_one(st, s2, row + 2, list_a, ksel, ksel);
  if (s3 >= st.tau) st = push_one(st, s3, row + 3, list_a, ksel, ksel);
  return st;
}

// Running best four of the per-16-row chunk maxima.  Every chunk maximum belongs to a different corpus row, so
// top[m - 1] >= v certifies "this CTA has m rows scoring at least v" -- what the threshold exchange needs when fewer
// than k + slack CTAs scan the corpus (m = ceil((k + slack) / CTAs), up to 4).  Branch-free insertion network.
__device__ __forceinline__ void top4_insert(float (&t)[4], float v) {   // v is never NaN (fmaxf drops NaN upstream)
  float a = v;
  const float n0 = fmaxf(t[0], a); a = fminf(t[0], a);
  const float n1 = fmaxf(t[1], a); a = fminf(t[1], a);
  const float n2 = fmaxf(t[2], a); a = fminf(t[2], a);
  t[3] = fmaxf(t[3], a); t[0] = n0; t[1] = n1; t[2] = n2;
}

__device__ __forceinline__ float top4_get(const float (&t)[4], int m) {   // t[m - 1] without dynamic register indexing
  return m == 1 ? t[0] : (m == 2 ? t[1] : (m == 3 ? t[2] : t[3]));
}

// Deferred slow path.  The hot loop only parks a group of four adjacent scores whose max reached the
// threshold (two predicated shared-memory stores); this routine runs when a thread's FIFO is nearly
// full and once at the end, re-tests the parked scores against the threshold as it stands NOW (usually
// much tighter) and admits the survivors.  The cold code is entered a handful of times per kernel
// instead of ~100x per warp, and fewer scores pass.
__device__ __noinline__ TopkState drain_fifo(TopkState st, uint32_t fifo_a, uint32_t ftag_a, int fcnt, uint32_t list_a,
                                             int ksel) {
#pragma unroll 1
  for (int rec = 0; rec < fcnt; ++rec) {
    int row;
    float s0, s1, s2, s3;
    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(row) : "r"(ftag_a + rec * (kTcQRows * 4u)));
    asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(s0), "=f"(s1), "=f"(s2), "=f"(s3)
                 : "r"(fifo_a + rec * (kTcQRows * 16u)));
    if (s0 >= st.tau) st = push_one(st, s0, row + 0, list_a, ksel, ksel);   // `>=` also rejects NaN
    if (s1 >= st.tau) st = push_one(st, s1, row + 1, list_a, ksel, ksel);
    if (s2 >= st.tau) st = push_one(st, s2, row + 2, list_a, ksel, ksel);
    if (s3 >= st.tau) st = push_one(st, s3, row + 3, list_a, ksel, ksel);
  }
  return st;
}

// Warp-cooperative: R-th largest of the values published for one query by up to 96 CTAs
// (lane i holds entries i, i+32, i+64; entries of other launches or not yet written are
// skipped).  Bisection on the value with ballot counts.  Returns -inf when fewer than R CTAs
// have published.  All lanes return the same value.  The three entries are loaded by the caller
// (exchange_load) so that several rows' loads can be in flight before any is consumed.
struct PubEntries { unsigned long long e[3]; };
__device__ __forceinline__ PubEntries exchange_load(const unsigned long long* pubrow, int nuse, int lane) {
  PubEntries r;
#pragma unroll
  for (int k = 0; k < 3; ++k) { const int i = lane + 32 * k; r.e[k] = (i < nuse) ? __ldcg(pubrow + i) : 0ull; }
  return r;
}
__device__ float exchange_select(const PubEntries& ent, int nuse, int R, uint32_t epoch, int lane) {
  float v[3];
  float lo = INFINITY, hi = -INFINITY;
  int nvalid = 0;
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    const int i = lane + 32 * k;
    const unsigned long long e = ent.e[k];
    const bool ok = (i < nuse) && static_cast<uint32_t>(e >> 32) == epoch;
    v[k] = ok ? ord_to_f32(static_cast<uint32_t>(e)) : __int_as_float(0x7FC00000);
    nvalid += __popc(__ballot_sync(0xffffffffu, ok));
    lo = fminf(lo, v[k]); hi = fmaxf(hi, v[k]);   // fminf / fmaxf skip NaN
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    lo = fminf(lo, __shfl_xor_sync(0xffffffffu, lo, o));
    hi = fmaxf(hi, __shfl_xor_sync(0xffffffffu, hi, o));
  }
  if (nvalid < R) return -INFINITY;
  // invariant: count(v >= lo) >= R
  for (int round = 0; round < 12; ++round) {
    const float mid = 0.5f * (lo + hi);
    int c = 0;
#pragma unroll
    for (int k = 0; k < 3; ++k) c += __popc(__ballot_sync(0xffffffffu, v[k] >= mid));
    if (c >= R) lo = mid; else hi = mid;
  }
  return lo;
}

__device__ __forceinline__ float read_threshold(const unsigned long long* tq, uint32_t epoch) {
  const unsigned long long e = __ldcg(tq);
  return (static_cast<uint32_t>(e >> 32) == epoch) ? __uint_as_float(static_cast<uint32_t>(e)) : -INFINITY;
}

constexpr uint32_t kDescHi = 0x40004040u;  // SBO = 1024 B, descriptor version 1, SWIZZLE_128B
constexpr uint32_t kNaNBits = 0x7FC00000u;

// kMask: the queries of the batch carry different tenant scopes (user_id == u OR org_id == o,
// weaviate_client.py:244-249).  A pre-pass has written one 32-bit word per corpus row -- bit s set when scope s of the
// batch may see the row -- and every query knows its scope's bit: scores of invisible rows become NaN before anything
// else looks at them, exactly like tombstones.  (One scope for the whole batch needs none of this: it folds into the
// inverse norms.)
template <int kCtaGroup, int kEpiGroups, bool kMask>
__global__ void __launch_bounds__(32 * (4 * kEpiGroups + 3), 1)
simtopk_tc_kernel(const __grid_constant__ CUtensorMap tmap, const TcParams p) {
  constexpr int kEpiWarps = 4 * kEpiGroups;
  constexpr int kProducerWarp = kEpiWarps + kThrWarps;
  constexpr int kMmaWarp = kProducerWarp + 1;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // SWIZZLE_128B tiles need 1024-byte alignment; the runtime only guarantees 16.  Offsetting
  // the declared array (rather than round-tripping through an integer) keeps the compiler's
  // shared-address-space inference, i.e. LDS/STS instead of generic loads.
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);

  const SmemLayout L = make_layout(kCtaGroup, kEpiGroups, p.num_stages, p.ksel, p.dim);
  float* normbuf = reinterpret_cast<float*>(smem + L.off_norm);      // [4 warps][2][64]
  uint32_t* maskbuf = reinterpret_cast<uint32_t*>(smem + L.off_mask);  // [4 warps][2][64]
  volatile float* tau_s = reinterpret_cast<volatile float*>(smem + L.off_tau);   // [128]
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L.off_bar);
  uint64_t* full_bar = bars;                              // [kTcMaxStages]
  uint64_t* empty_bar = bars + kTcMaxStages;              // [kTcMaxStages]
  uint64_t* tmem_full = bars + 2 * kTcMaxStages;          // [2]
  uint64_t* tmem_empty = bars + 2 * kTcMaxStages + 2;     // [2]
  uint64_t* q_ready = bars + 2 * kTcMaxStages + 4;        // [1]
  uint64_t* q_staged = bars + 2 * kTcMaxStages + 5;       // [1] local: the query load no longer uses ring stages as scratch
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 2 * kTcMaxStages + 6);
  volatile int* epi_done = reinterpret_cast<volatile int*>(tmem_ptr_smem + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = (kCtaGroup == 2) ? cluster_ctarank() : 0u;
  // launch counter tagging the exchange-table entries: read from device memory (the finalize kernel behind this launch
  // bumps it) so that a captured CUDA graph advances by itself when replayed
  const uint32_t epoch = p.epoch_ptr ? *p.epoch_ptr : p.epoch;

  // Work split.  tset = which set of corpus tiles this CTA (pair) walks; qblock = which 128 queries.
  // With more than 256 queries in a launch, S = n_qblocks / 2 CTA pairs ("query super-blocks") walk the SAME tile set
  // side by side, each with its own 256 queries in TMEM: the first pair to ask for a tile pulls it from HBM, its
  // siblings hit L2, so the corpus crosses the HBM interface once per 256 * S queries instead of once per 256.
  int qblock, tset, n_tsets;
  if constexpr (kCtaGroup == 2) {
    const int n_super = p.n_qblocks >> 1, pair = blockIdx.x >> 1;
    qblock = (pair % n_super) * 2 + static_cast<int>(rank);
    tset = pair / n_super;
    n_tsets = (gridDim.x >> 1) / n_super;
  } else {
    qblock = blockIdx.x % p.n_qblocks;
    tset = blockIdx.x / p.n_qblocks;
    n_tsets = gridDim.x / p.n_qblocks;
  }
  const int my_tiles = (p.n_tiles > tset) ? (p.n_tiles - tset + n_tsets - 1) / n_tsets : 0;
  const int kbs = p.dim / kTcKBlock;  // 128-byte k-blocks per row

  // exchange geometry: R-th largest of the m-th best of `nuse` CTAs is a valid threshold
  const int ksel = p.ksel;
  const int nuse = min(n_tsets, 96);
  const int xm = (ksel + nuse - 1) / nuse;            // rows every publishing CTA vouches for (1 .. 4)
  const int xR = (ksel + xm - 1) / xm;
  // (only CTAs that own at least one tile ever publish)
  const bool xchg = (p.pub != nullptr) && xm <= 4 && (xR <= min(nuse, p.n_tiles));

  if constexpr (kCtaGroup == 2) cluster_sync_all();  // both CTAs resident before the paired TMEM alloc

  if (warp == kMmaWarp && lane == 0) {
    prefetch_tmap(&tmap);
    for (int i = 0; i < p.num_stages; ++i) {
      mbar_init(&full_bar[i], kCtaGroup);  // leader's expect_tx arrive (+ the peer producer's arrive)
      mbar_init(&empty_bar[i], 1);         // one tcgen05.commit
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full[i], 1);                       // one tcgen05.commit
      mbar_init(&tmem_empty[i], 4 * kCtaGroup);          // the four warps (per CTA) that drain this buffer
    }
    mbar_init(q_ready, kEpiWarps * kCtaGroup);
    mbar_init(q_staged, kEpiWarps);
    *epi_done = 0;
    fence_mbar_init();
  } else if (warp == kProducerWarp) {
    tmem_alloc<kCtaGroup>(tmem_ptr_smem, 512);
    tmem_relinquish<kCtaGroup>();
  } else if (warp < kThrWarps) {
    for (int i = lane; i < kTcQRows; i += 32) tau_s[i] = -INFINITY;
  }
  tc_fence_before();
  if constexpr (kCtaGroup == 2) cluster_sync_all(); else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  grid_dep_launch();   // the exact re-rank kernel behind this one may start its prologue as CTAs here retire

  if (warp == kProducerWarp) {
    // ============================== TMA producer ==============================
    if (lane == 0) {
      // read once -> evict first; shared with sibling CTAs (two single CTAs or several query super-blocks) -> keep in L2
      const uint64_t hint = ((kCtaGroup == 2 && p.n_qblocks == 2) || p.n_qblocks == 1) ? kEvictFirst : kEvictNormal;
      int stage = 0; uint32_t phase = 0;
      bool q_done = p.num_stages < 2 * kQStageBufs;   // too few stages: the query load does not borrow any
      long long tp_wait = 0;
      const long long tp_begin = TCLK();
#ifdef AUR_TC_PROFILE
      unsigned long long gt0; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt0));
#endif
      for (int it = 0; it < my_tiles; ++it) {
        const int tile = tset + it * n_tsets;
        const int row0 = tile * kTcTileN + static_cast<int>(rank) * (kTcTileN / kCtaGroup);
        for (int kb0 = 0; kb0 < kbs; kb0 += kTcKbPerStage) {
          const int nkb = min(kTcKbPerStage, kbs - kb0);
          if (!q_done && stage >= p.num_stages - kQStageBufs) { mbar_wait(q_staged, 0); q_done = true; }   // (see the query load)
          {
            const long long t0 = TCLK();
            mbar_wait(&empty_bar[stage], phase ^ 1u);
            tp_wait += TCLK() - t0;
          }
          uint8_t* dst = smem + static_cast<uint32_t>(stage) * L.stage_bytes;
          if constexpr (kCtaGroup == 1) {
            mbar_arrive_expect_tx(&full_bar[stage], static_cast<uint32_t>(nkb) * L.box_bytes);
            for (int j = 0; j < nkb; ++j)
              tma_load_2d(dst + j * L.box_bytes, &tmap, &full_bar[stage], (kb0 + j) * kTcKBlock, row0, hint);
          } else {
            for (int j = 0; j < nkb; ++j)
              tma_load_2d_pair(dst + j * L.box_bytes, &tmap, &full_bar[stage], (kb0 + j) * kTcKBlock, row0, hint);
            if (rank == 0) mbar_arrive_expect_tx(&full_bar[stage], 2u * static_cast<uint32_t>(nkb) * L.box_bytes);
            else mbar_arrive_cluster(&full_bar[stage], 0);
          }
          if (++stage == p.num_stages) { stage = 0; phase ^= 1u; }
        }
      }
#ifdef AUR_TC_PROFILE
      if ((p.dbg_flags & 64) && p.dbg_scores != nullptr) {
        unsigned long long gt1; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt1));
        float* d = p.dbg_scores + static_cast<size_t>(blockIdx.x) * kTcQRows * kTcTileN + 36;
        d[0] = static_cast<float>(tp_wait); d[1] = static_cast<float>(TCLK() - tp_begin);
        d[2] = static_cast<float>(gt1 - gt0);     // ns: cycles / ns = the SM clock the kernel really ran at
      }
#endif
    }
  } else if (warp == kMmaWarp) {
    // ============================== MMA issuer ==============================
    // The whole warp walks the pipeline (so every operand stays warp-uniform and lives in
    // uniform registers); one elected lane issues the MMAs and their commits.
    if (rank == 0) {
      mbar_wait(q_ready, 0);  // queries of (both) CTAs are in TMEM
      tc_fence_after();
      constexpr uint32_t idesc = idesc_bf16_f32(128 * kCtaGroup, kTcTileN);
      constexpr uint32_t kBox16 = ((kTcTileN / kCtaGroup) * 128u) >> 4;  // box stride in descriptor units
      int stage = 0; uint32_t phase = 0;
      long long tm_empty = 0, tm_full = 0;
      const long long tm_begin = TCLK();
      for (int it = 0; it < my_tiles; ++it) {
        const int b = it & 1;
        {
          const long long t0 = TCLK();
          mbar_wait(&tmem_empty[b], ((static_cast<uint32_t>(it) >> 1) & 1u) ^ 1u);
          tm_empty += TCLK() - t0;
        }
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + kTcAccCol0 + b * kTcTileN;
        for (int kb0 = 0; kb0 < kbs; kb0 += kTcKbPerStage) {
          const int nkb = min(kTcKbPerStage, kbs - kb0);
          {
            const long long t0 = TCLK();
            mbar_wait(&full_bar[stage], phase);
            tm_full += TCLK() - t0;
          }
          tc_fence_after();
          const uint32_t base_lo = (smem_u32(smem + static_cast<uint32_t>(stage) * L.stage_bytes) & 0x3FFFFu) >> 4;
          const uint32_t a_col = tmem_base + static_cast<uint32_t>(kb0) * 32u;
          if (elect_one()) {
#pragma unroll
            for (int j = 0; j < kTcKbPerStage; ++j) {
              if (j < nkb && !(p.dbg_flags & 2)) {
                const int kb = kb0 + j;
                if (kb < kTcTmemDim / kTcKBlock) {
#pragma unroll
                  for (int k = 0; k < 4; ++k) {  // 4 x K=16 per 128-byte k-block, queries from TMEM
                    const uint32_t acc = (j | k) != 0 ? 1u : (kb0 != 0 ? 1u : 0u);
                    mma_ts_bf16<kCtaGroup>(d_tmem, a_col + j * 32 + k * 8, pack_u64(base_lo + j * kBox16 + k * 2, kDescHi),
                                           idesc, acc);
                  }
                } else {                         // dims past 768: queries from shared memory (SS)
                  const uint32_t qs_lo = ((smem_u32(smem + L.off_qs) & 0x3FFFFu) >> 4) +
                                         static_cast<uint32_t>(kb - kTcTmemDim / kTcKBlock) * ((kTcQRows * 128u) >> 4);
#pragma unroll
                  for (int k = 0; k < 4; ++k)
                    mma_ss_bf16<kCtaGroup>(d_tmem, pack_u64(qs_lo + k * 2, kDescHi), pack_u64(base_lo + j * kBox16 + k * 2, kDescHi),
                                           idesc, 1u);
                }
              }
            }
            mma_commit<kCtaGroup>(&empty_bar[stage]);  // smem slot free once these MMAs retire
            if (kb0 + kTcKbPerStage >= kbs) mma_commit<kCtaGroup>(&tmem_full[b]);  // accumulator complete
          }
          __syncwarp();
          if (++stage == p.num_stages) { stage = 0; phase ^= 1u; }
        }
      }
      if ((p.dbg_flags & 64) && p.dbg_scores != nullptr && lane == 0) {
        float* d = p.dbg_scores + static_cast<size_t>(blockIdx.x) * kTcQRows * kTcTileN + 32;
        d[0] = static_cast<float>(tm_empty); d[1] = static_cast<float>(tm_full);
        d[2] = static_cast<float>(TCLK() - tm_begin);
      }
    }
  } else {
    // ================= threshold warp (0) and epilogue warps (1-4) =================
    const int quarter = warp & 3;              // TMEM lane quarter this warp may touch
    const int r = quarter * 32 + lane;         // TMEM lane == query row inside the CTA
    const int qglob = qblock * kTcQRows + r;   // query index inside this launch
    const uint32_t lane_addr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16);
    // exchange table: [qblock][query][CTA] -> a query's row is contiguous (592 B for 74 CTAs);
    // behind it, one published threshold per query
    const int pub_stride = (n_tsets + 1) & ~1;
    unsigned long long* pub_base =
        reinterpret_cast<unsigned long long*>(p.pub) + static_cast<size_t>(qblock) * kTcQRows * pub_stride;
    unsigned long long* thr_base = reinterpret_cast<unsigned long long*>(p.pub) +
                                   static_cast<size_t>(p.n_qblocks) * kTcQRows * pub_stride + qblock * kTcQRows;
    unsigned long long* pubrow = pub_base + static_cast<size_t>(r) * pub_stride;
    const unsigned long long* thr_q = thr_base + r;

    if (warp < kThrWarps) {
      // ============================== threshold warp ==============================
      // Serve the queries r = tset, tset + n_tsets, ... of this CTA's query block.
      constexpr int kSlots = 8;                      // queries one CTA may have to serve: ceil(128 / tile sets), tile sets >= 16
      float cur[kSlots];
#pragma unroll
      for (int sl = 0; sl < kSlots; ++sl) cur[sl] = -INFINITY;
      int polls = 0;
      while (xchg && my_tiles > 0) {
        // every epilogue thread of the grid waits for the FIRST threshold (bootstrap): poll fast until it is out
        __nanosleep(polls < 48 ? 150 : 1500);
        ++polls;
        const bool done = *epi_done >= kEpiWarps;
        // every global load of this round is issued before the first one is consumed: one L2 round trip per round
        unsigned long long cached[kTcQRows / 32];
#pragma unroll
        for (int j = 0; j < kTcQRows / 32; ++j) cached[j] = __ldcg(thr_base + j * 32 + lane);
        PubEntries ent[kSlots];
        int nslot = 0;
#pragma unroll
        for (int sl = 0; sl < kSlots; ++sl) {
          const int rq = tset + sl * n_tsets;
          if (rq < kTcQRows) { ent[sl] = exchange_load(pub_base + static_cast<size_t>(rq) * pub_stride, nuse, lane); nslot = sl + 1; }
        }
#pragma unroll
        for (int sl = 0; sl < kSlots; ++sl) {
          if (sl < nslot) {
            const int rq = tset + sl * n_tsets;
            const float t = exchange_select(ent[sl], nuse, xR, epoch, lane);
            if (t > cur[sl]) {
              cur[sl] = t;
              if (lane == 0) __stcg(thr_base + rq, (static_cast<unsigned long long>(epoch) << 32) | __float_as_uint(t));
            }
          }
        }
        // Refresh this CTA's cache of its 128 queries' certified thresholds.  The epilogue reads them from shared
        // memory once per tile; a global (L2) read there sat on the per-tile critical path with its full latency.
#pragma unroll
        for (int j = 0; j < kTcQRows / 32; ++j)
          if (static_cast<uint32_t>(cached[j] >> 32) == epoch) tau_s[j * 32 + lane] = __uint_as_float(static_cast<uint32_t>(cached[j]));   // only ever rises
        if (done) break;
      }
    } else {
      // ============================== epilogue warps ==============================
      const int grp = (warp - kThrWarps) >> 2;   // epilogue group: takes tiles grp, grp + groups, ...
      const long long t_kernel0 = TCLK();
      // ---- park this thread's query row in TMEM (bf16 pairs, K ascending along columns);
      //      with two groups each loads every other k-block
      {
        // A thread needs ITS row (TMEM lane = query), but 32 lanes reading 32 different rows cost 32 L1 wavefronts per
        // load instruction -- 14k cycles for the block.  So a warp reads its 32 rows coalesced (lane l takes 16-byte
        // piece l of 4 consecutive rows per instruction), transposes through shared memory -- borrowed from the last
        // ring stages, which the TMA producer leaves alone until q_staged completes -- and every lane reads its own
        // row back.  Staging layout = the 128-byte-swizzled K-major tile TMA would produce, so dims past 768 (which
        // stay in shared memory as the SS-MMA operand) are written straight to their final place.
        const bool staged = p.num_stages >= 2 * kQStageBufs;
        const int ew = warp - kThrWarps;                      // epilogue warp 0 .. 4G-1
        constexpr int kBufs = (kEpiGroups == 1) ? 2 : 1;      // 4 KB transpose buffers per warp
        uint8_t* stg = smem + static_cast<uint32_t>(p.num_stages - kQStageBufs) * L.stage_bytes + ew * (kBufs * 4096);
        const uint8_t* qbase = reinterpret_cast<const uint8_t*>(p.q);
        const size_t row_bytes = static_cast<size_t>(p.dim) * 2;
        const int wrow0 = qblock * kTcQRows + quarter * 32;   // first query of this warp
        auto load_kb = [&](int kb, uint4 (&x)[8]) {
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int id = i * 32 + lane, row = id >> 3, c = id & 7;
            x[i] = make_uint4(0, 0, 0, 0);
            if (kb < kbs && wrow0 + row < p.nq) x[i] = ldg_nc_v4(qbase + static_cast<size_t>(wrow0 + row) * row_bytes + kb * 128 + c * 16);
          }
        };
        auto scatter_kb = [&](uint8_t* tile, const uint4 (&x)[8]) {   // tile: this warp's [32 rows x 128 B], swizzled
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int id = i * 32 + lane, row = id >> 3, c = id & 7;
            *reinterpret_cast<uint4*>(tile + row * 128 + ((c ^ (row & 7)) << 4)) = x[i];
          }
        };
        if (staged) {
          // every CTA pair reads the same 256 x dim block at the same moment: start each pair at a different
          // k-block (rotation by tile set) so they do not all queue on the same L2 lines
          const int n_j = (kbs - grp + kEpiGroups - 1) / kEpiGroups;      // k-blocks this group loads
          const int rot = n_j > 0 ? tset % n_j : 0;
          auto kb_of = [&](int j) { return (j < n_j) ? grp + ((j + rot) % n_j) * kEpiGroups : kbs; };
          // four k-blocks of loads (32 x 16 B per lane) are issued before the first is consumed: the block is read
          // in three L2 round trips instead of twelve
          constexpr int kDepth = 4;
          uint4 x[kDepth][8];
          for (int j0 = 0; j0 < n_j; j0 += kDepth) {
#pragma unroll
            for (int u = 0; u < kDepth; ++u) load_kb(kb_of(j0 + u), x[u]);
#pragma unroll
            for (int u = 0; u < kDepth; ++u) {
              const int j = j0 + u;
              if (j >= n_j) break;
              const int kb = kb_of(j);
              if (kb < kTcTmemDim / kTcKBlock) {
                uint8_t* buf = stg + (j % kBufs) * 4096;
                if (kBufs == 1) __syncwarp();
                scatter_kb(buf, x[u]);
                __syncwarp();
                uint32_t v[2][16];
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                  const uint4 t = *reinterpret_cast<const uint4*>(buf + lane * 128 + ((c ^ (lane & 7)) << 4));
                  v[c >> 2][(c & 3) * 4 + 0] = t.x; v[c >> 2][(c & 3) * 4 + 1] = t.y;
                  v[c >> 2][(c & 3) * 4 + 2] = t.z; v[c >> 2][(c & 3) * 4 + 3] = t.w;
                }
                tmem_st_x16(lane_addr + kb * 32, v[0]);
                tmem_st_x16(lane_addr + kb * 32 + 16, v[1]);
              } else {
                scatter_kb(smem + L.off_qs + static_cast<uint32_t>(kb - kTcTmemDim / kTcKBlock) * (kTcQRows * 128u) + quarter * 32 * 128, x[u]);
              }
            }
          }
        } else {   // hardly any ring (large k at dim > 768): every thread fetches its own row
          const uint4* src = reinterpret_cast<const uint4*>(p.q + static_cast<size_t>(qglob) * p.dim);
          for (int kb = grp; kb < kbs; kb += kEpiGroups) {
            uint32_t v[2][16];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              uint4 x = make_uint4(0, 0, 0, 0);
              if (qglob < p.nq) x = ldg_nc_v4(src + kb * 8 + i);
              v[i >> 2][(i & 3) * 4 + 0] = x.x; v[i >> 2][(i & 3) * 4 + 1] = x.y;
              v[i >> 2][(i & 3) * 4 + 2] = x.z; v[i >> 2][(i & 3) * 4 + 3] = x.w;
            }
            if (kb < kTcTmemDim / kTcKBlock) {
              tmem_st_x16(lane_addr + kb * 32, v[0]);
              tmem_st_x16(lane_addr + kb * 32 + 16, v[1]);
            } else {   // K-major tile with the 128-byte swizzle TMA would have produced: 16-byte chunk c of row r at c ^ (r & 7)
              uint8_t* qrow = smem + L.off_qs + static_cast<uint32_t>(kb - kTcTmemDim / kTcKBlock) * (kTcQRows * 128u) + r * 128u;
#pragma unroll
              for (int c = 0; c < 8; ++c)
                *reinterpret_cast<uint4*>(qrow + ((c ^ (r & 7)) << 4)) =
                    make_uint4(v[c >> 2][(c & 3) * 4 + 0], v[c >> 2][(c & 3) * 4 + 1], v[c >> 2][(c & 3) * 4 + 2], v[c >> 2][(c & 3) * 4 + 3]);
            }
          }
        }
        tmem_wait_st();
        fence_proxy_async_smem();   // (shared-memory part of the queries -> visible to the tensor core; the borrowed
                                    //  ring stages -> safe for TMA to overwrite)
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          mbar_arrive(q_staged);
          if constexpr (kCtaGroup == 2) mbar_arrive_cluster(q_ready, 0); else mbar_arrive(q_ready);
        }
      }
      float* mynorm = normbuf + (warp - kThrWarps) * 2 * kTcTileN;
      uint32_t* mymask = maskbuf + (warp - kThrWarps) * 2 * kTcTileN;
      uint32_t mybit = 0u;                                       // this query's tenant-scope bit
      if constexpr (kMask) { if (qglob < p.nq) mybit = 1u << (p.q_scope[qglob] & 31); }
      const uint32_t list_a = smem_u32(smem + L.off_list) + (static_cast<uint32_t>(grp) * L.lcap * kTcQRows + r) * 8u;
      // deferred-candidate FIFO (see drain_fifo) whenever shared memory has room for it
      const bool use_fifo = L.fifo_recs > 0;
      const uint32_t fifo_a = smem_u32(smem + L.off_fifo) + r * 16u;
      const uint32_t ftag_a = smem_u32(smem + L.off_fifo) + L.fifo_recs * kTcQRows * 16u + r * 4u;
      int fcnt = 0;
      TopkState st;
      st.min_key = kKeyEmpty; st.tau_local = -INFINITY; st.tau = -INFINITY;
      st.top[0] = st.top[1] = st.top[2] = st.top[3] = -INFINITY; st.minpos = 0; st.nfill = 0;
      if (qglob >= p.nq) st.tau = INFINITY;   // padding row of a partial query block: admits nothing, appends nothing
      float published = -INFINITY;
      bool booted = false;   // the bootstrap has already fed the first tile's chunk maxima into st.top
      int nslow = 0;
      long long t_wait = 0, t_slow = 0, t_ld = 0, t_top = 0, t_fast = 0, t_chunks = 0, t_pub = 0;
      const long long t_begin = TCLK();
      long long t_boot = 0, t_loop_end = 0;

      // inverse norms: tile(0) into buffer 0, tile(1) in flight in registers
      uint32_t mm0 = 0u, mm1 = 0u;                              // (masks of the tile in flight, beside its norms)
      auto load_norms = [&](int li2, float& a, float& b2) {
        a = __uint_as_float(kNaNBits); b2 = a;
        mm0 = 0u; mm1 = 0u;
        const int it2 = grp + li2 * kEpiGroups;
        if (it2 < my_tiles) {
          const int64_t rbase = static_cast<int64_t>(tset + it2 * n_tsets) * kTcTileN;
          if (rbase + lane < p.n_rows) a = __ldg(p.inv_norm + rbase + lane);
          if (rbase + lane + 32 < p.n_rows) b2 = __ldg(p.inv_norm + rbase + lane + 32);
          if constexpr (kMask) {
            if (rbase + lane < p.n_rows) mm0 = __ldg(p.row_mask + rbase + lane);
            if (rbase + lane + 32 < p.n_rows) mm1 = __ldg(p.row_mask + rbase + lane + 32);
          }
        }
      };
      float nn0, nn1;
      load_norms(0, nn0, nn1);
      mynorm[lane] = nn0; mynorm[lane + 32] = nn1;
      if constexpr (kMask) { mymask[lane] = mm0; mymask[lane + 32] = mm1; }
      load_norms(1, nn0, nn1);
      __syncwarp();

      // Bootstrap on the first tile: nobody has a threshold yet, and pushing 64 arbitrary rows
      // through the list would be all waste.  Read the tile once just for its best score(s),
      // publish them, wait until enough CTAs have done the same (~5 us, once), and let the
      // main loop examine the tile against the first certified threshold.  The accumulator
      // is not released here, so the MMA cannot overwrite it before the loop reads it again.
      if (xchg && grp < my_tiles) {
        mbar_wait(&tmem_full[grp & 1], 0);
        tc_fence_after();
#pragma unroll 1
        for (int c = 0; c < 4; ++c) {
          uint32_t a16[16];
          tmem_ld_x16(lane_addr + kTcAccCol0 + (grp & 1) * kTcTileN + c * 16, a16);
          tmem_wait_ld();
          float m = -INFINITY;
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            float v = __uint_as_float(a16[j]) * mynorm[c * 16 + j];
            if constexpr (kMask) { if (!(mymask[c * 16 + j] & mybit)) v = __uint_as_float(kNaNBits); }
            m = fmaxf(m, v);
          }
          top4_insert(st.top, m);       // (the main loop skips the tracker for this tile: it is counted here)
        }
        const float pv0 = top4_get(st.top, xm);
        if (pv0 > -INFINITY) {
          published = pv0;
          atomicMax(pubrow + tset, (static_cast<unsigned long long>(epoch) << 32) | f32_to_ord(pv0));
        }
        booted = true;
        const long long tb = clock64();
        float tboot = -INFINITY;
        do {
          __nanosleep(100);
          tboot = tau_s[r];
        } while (__any_sync(0xffffffffu, tboot == -INFINITY) && clock64() - tb < 100000);
        st.tau = fmaxf(st.tau, tboot);
        t_boot = TCLK() - t_begin;
      }

      int li = 0;  // this group's iteration count
      for (int it = grp; it < my_tiles; it += kEpiGroups, ++li) {
        const int tile = tset + it * n_tsets;
        const int row0 = tile * kTcTileN;
        const int b = it & 1;
        const float* nb = mynorm + (li & 1) * kTcTileN;
        const long long t_top0 = TCLK();
        const float thr_now = (xchg && !(p.dbg_flags & 16)) ? tau_s[r] : -INFINITY;   // shared-memory copy kept by the threshold warp

        // norms of the next tile (loaded one iteration ago) -> the other buffer; start the
        // loads for the tile after that.  A whole tile period hides the HBM latency.
        if (!(p.dbg_flags & 32)) {
          float* nnext = mynorm + ((li + 1) & 1) * kTcTileN;
          nnext[lane] = nn0; nnext[lane + 32] = nn1;
          if constexpr (kMask) { uint32_t* mnext = mymask + ((li + 1) & 1) * kTcTileN; mnext[lane] = mm0; mnext[lane + 32] = mm1; }
          load_norms(li + 2, nn0, nn1);
          __syncwarp();
        }

        {
          const long long t0 = TCLK();
          mbar_wait(&tmem_full[b], (static_cast<uint32_t>(it) >> 1) & 1u);
          t_wait += TCLK() - t0;
        }
        tc_fence_after();
        const long long t_ld0 = TCLK();
        uint32_t acc[4][16];
        if (!(p.dbg_flags & 1)) {
#pragma unroll
          for (int c = 0; c < 4; ++c) tmem_ld_x16(lane_addr + kTcAccCol0 + b * kTcTileN + c * 16, acc[c]);
          tmem_wait_ld();
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {  // accumulator drained into registers: hand the buffer back to the MMA warp
          if constexpr (kCtaGroup == 2) mbar_arrive_cluster(&tmem_empty[b], 0); else mbar_arrive(&tmem_empty[b]);
        }
        t_ld += TCLK() - t_ld0;
        if (p.dbg_flags & (1 | 4)) continue;
        const long long t_fast0 = TCLK();

        // Fast path: scale by 1/|c_j| in place (packed FMUL2) and keep one running max per 16
        // scores.  (NaN norm = tombstone / out of range: fmaxf drops it, `>=` rejects it.)
        if constexpr (kMask) {   // rows this query's tenant scope may not see: NaN, like tombstones
          const uint32_t* mb = mymask + (li & 1) * kTcTileN;
#pragma unroll
          for (int c = 0; c < 4; ++c)
#pragma unroll
            for (int j4 = 0; j4 < 4; ++j4) {
              const uint4 mk = *reinterpret_cast<const uint4*>(mb + c * 16 + j4 * 4);
              if (!(mk.x & mybit)) acc[c][j4 * 4 + 0] = kNaNBits;
              if (!(mk.y & mybit)) acc[c][j4 * 4 + 1] = kNaNBits;
              if (!(mk.z & mybit)) acc[c][j4 * 4 + 2] = kNaNBits;
              if (!(mk.w & mybit)) acc[c][j4 * 4 + 3] = kNaNBits;
            }
        }
        float cmax[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          float m = -INFINITY;
#pragma unroll
          for (int j4 = 0; j4 < 4; ++j4) {
            const ulonglong2 nv = *reinterpret_cast<const ulonglong2*>(nb + c * 16 + j4 * 4);
            const uint64_t p0 = mul_f32x2(pack_u64(acc[c][j4 * 4 + 0], acc[c][j4 * 4 + 1]), nv.x);
            const uint64_t p1 = mul_f32x2(pack_u64(acc[c][j4 * 4 + 2], acc[c][j4 * 4 + 3]), nv.y);
            unpack_u64(p0, acc[c][j4 * 4 + 0], acc[c][j4 * 4 + 1]);
            unpack_u64(p1, acc[c][j4 * 4 + 2], acc[c][j4 * 4 + 3]);
            m = fmaxf(fmaxf(m, fmaxf(__uint_as_float(acc[c][j4 * 4 + 0]), __uint_as_float(acc[c][j4 * 4 + 1]))),
                      fmaxf(__uint_as_float(acc[c][j4 * 4 + 2]), __uint_as_float(acc[c][j4 * 4 + 3])));
          }
          cmax[c] = m;
        }
        if (p.dbg_scores != nullptr && it == 0 && !(p.dbg_flags & 64)) {  // (group 0 owns tile 0)
#pragma unroll
          for (int c = 0; c < 4; ++c)
#pragma unroll
            for (int j = 0; j < 16; ++j)
              p.dbg_scores[(static_cast<size_t>(blockIdx.x) * kTcQRows + r) * kTcTileN + c * 16 + j] =
                  __uint_as_float(acc[c][j]);
        }

        t_fast += TCLK() - t_fast0;
        // what this CTA can vouch for (published below): chunk maxima are distinct rows
        if (!(booted && li == 0)) {
          if (xm == 1) st.top[0] = fmaxf(st.top[0], fmaxf(fmaxf(cmax[0], cmax[1]), fmaxf(cmax[2], cmax[3])));
          else { top4_insert(st.top, cmax[0]); top4_insert(st.top, cmax[1]); top4_insert(st.top, cmax[2]); top4_insert(st.top, cmax[3]); }
        }
        if (p.dbg_flags & 8) continue;
        const long long t_ch0 = TCLK();
        // A group of four scores whose max reaches this query's threshold goes out of line.
        st.tau = fmaxf(st.tau, thr_now);
        if (use_fifo) {
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            if (cmax[c] >= st.tau) {
              if (fcnt > static_cast<int>(L.fifo_recs) - 4) {   // no room for four more groups: make room (rare)
                const long long t0 = TCLK();
                st = drain_fifo(st, fifo_a, ftag_a, fcnt, list_a, ksel);
                fcnt = 0;
                ++nslow;
                t_slow += TCLK() - t0;
              }
#pragma unroll
              for (int g = 0; g < 4; ++g) {   // park the groups that reach the threshold: two stores each, no call
                const float m4 = fmaxf(fmaxf(__uint_as_float(acc[c][4 * g]), __uint_as_float(acc[c][4 * g + 1])),
                                       fmaxf(__uint_as_float(acc[c][4 * g + 2]), __uint_as_float(acc[c][4 * g + 3])));
                if (m4 >= st.tau) {
                  asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(fifo_a + static_cast<uint32_t>(fcnt) * (kTcQRows * 16u)),
                               "r"(acc[c][4 * g]), "r"(acc[c][4 * g + 1]), "r"(acc[c][4 * g + 2]), "r"(acc[c][4 * g + 3]) : "memory");
                  asm volatile("st.shared.b32 [%0], %1;" ::"r"(ftag_a + static_cast<uint32_t>(fcnt) * (kTcQRows * 4u)),
                               "r"(row0 + c * 16 + 4 * g) : "memory");
                  ++fcnt;
                }
              }
            }
          }
        } else {
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          if (cmax[c] >= st.tau) {
            const long long t0 = TCLK();
#pragma unroll
            for (int g = 0; g < 4; ++g) {
              const float s0 = __uint_as_float(acc[c][4 * g]), s1 = __uint_as_float(acc[c][4 * g + 1]);
              const float s2 = __uint_as_float(acc[c][4 * g + 2]), s3 = __uint_as_float(acc[c][4 * g + 3]);
              if (fmaxf(fmaxf(s0, s1), fmaxf(s2, s3)) >= st.tau)
                st = push_group4(st, s0, s1, s2, s3, row0 + c * 16 + 4 * g, list_a, ksel);
            }
            ++nslow;
            t_slow += TCLK() - t0;
          }
        }
        }

        if (li < 8) t_top += TCLK() - t_ch0; else t_chunks += TCLK() - t_ch0;   // t_top reused: early tiles
        const long long t_pub0 = TCLK();
        // publish this CTA's m-th best for the exchange (monotone, so stale reads stay valid)
        const float pv = top4_get(st.top, xm);
        if (xchg && pv > published) {
          published = pv;
          atomicMax(pubrow + tset, (static_cast<unsigned long long>(epoch) << 32) | f32_to_ord(pv));
        }
        t_pub += TCLK() - t_pub0;
      }
      t_loop_end = TCLK();
      __syncwarp();
      if (lane == 0) atomicAdd(const_cast<int*>(epi_done), 1);   // lets the threshold warps go
      // one read of the certified threshold (a global round trip) serves both the parked candidates and the final cut
      float tau_end = -INFINITY;
      if (xchg && my_tiles > 0) tau_end = read_threshold(thr_q, epoch);
      st.tau = fmaxf(st.tau, tau_end);
      if (use_fifo) {   // whatever is still parked meets the final threshold
        if (__any_sync(0xffffffffu, fcnt > 0)) st = drain_fifo(st, fifo_a, ftag_a, fcnt, list_a, ksel);
        fcnt = 0;
      }

      // ---- append the survivors (score >= the certified threshold, at most ksel of them)
      //      to this query's compact candidate row
      st = compact_list(st, list_a, ksel);
      {
        const size_t cap = static_cast<size_t>(n_tsets) * kEpiGroups * ksel;
        uint64_t* out = p.cand + static_cast<size_t>(qglob) * cap;
        if (st.nfill > 0 && qglob < p.nq) {
          const uint32_t slot0 = atomicAdd(p.cand_count + qglob, static_cast<uint32_t>(st.nfill));
          for (int t = 0; t < st.nfill; ++t) out[slot0 + t] = lds_u64(list_a + t * kSlot);
        }
      }
      if ((p.dbg_flags & 64) && p.dbg_scores != nullptr && grp == 0) {
        float* d = p.dbg_scores + (static_cast<size_t>(blockIdx.x) * kTcQRows + r) * kTcTileN;
        d[0] = static_cast<float>(st.nfill); d[1] = static_cast<float>(nslow);
        d[2] = tau_end; d[3] = st.tau_local;
        d[8] = static_cast<float>(t_wait); d[9] = static_cast<float>(t_slow);
        d[10] = 0.f; d[11] = static_cast<float>(t_ld);
        d[12] = static_cast<float>(TCLK() - t_begin);
        d[13] = static_cast<float>(t_top); d[14] = static_cast<float>(t_fast);
        d[15] = static_cast<float>(t_chunks); d[16] = static_cast<float>(t_pub);
        d[17] = static_cast<float>(t_boot); d[18] = static_cast<float>(t_begin - t_kernel0);
        d[19] = static_cast<float>(TCLK() - t_loop_end);
      }
    }
  }

  // ============================== teardown ==============================
  __syncwarp();
  tc_fence_before();
  if constexpr (kCtaGroup == 2) cluster_sync_all(); else __syncthreads();
  tc_fence_after();
  if (warp == kProducerWarp) tmem_dealloc<kCtaGroup>(tmem_base, 512);
}

template <int kCtaGroup, int kEpiGroups, bool kMask>
cudaError_t launch_variant(const cudaLaunchConfig_t& cfg, const CUtensorMap& tm, const TcParams& p, size_t smem) {
  auto kern = simtopk_tc_kernel<kCtaGroup, kEpiGroups, kMask>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
  if (e != cudaSuccess) return e;
  return cudaLaunchKernelEx(&cfg, kern, tm, p);
}

}  // namespace

size_t tc_smem_bytes(int cta_group, int epi_groups, int num_stages, int ksel, int dim) {
  return make_layout(cta_group, epi_groups, num_stages, ksel, dim).total + 1024;  // + alignment slack
}

int tc_pick_stages(int cta_group, int epi_groups, int ksel, int dim, size_t smem_limit) {
  for (int s = kTcMaxStages; s >= 2; --s)
    if (tc_smem_bytes(cta_group, epi_groups, s, ksel, dim) <= smem_limit) return s;
  return 0;
}

cudaError_t tc_launch(int cta_group, int epi_groups, int grid, const void* tmap, const TcParams& p, size_t smem,
                      cudaStream_t s) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(32 * (4 * epi_groups + 3));
  cfg.dynamicSmemBytes = smem;
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = cta_group;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  const CUtensorMap& tm = *reinterpret_cast<const CUtensorMap*>(tmap);
  const bool mask = p.row_mask != nullptr;
  if (cta_group == 2) {
    if (epi_groups == 2) return mask ? launch_variant<2, 2, true>(cfg, tm, p, smem) : launch_variant<2, 2, false>(cfg, tm, p, smem);
    return mask ? launch_variant<2, 1, true>(cfg, tm, p, smem) : launch_variant<2, 1, false>(cfg, tm, p, smem);
  }
  if (epi_groups == 2) return mask ? launch_variant<1, 2, true>(cfg, tm, p, smem) : launch_variant<1, 2, false>(cfg, tm, p, smem);
  return mask ? launch_variant<1, 1, true>(cfg, tm, p, smem) : launch_variant<1, 1, false>(cfg, tm, p, smem);
}

}  // namespace aur
