// Fused similarity + top-k for sm_100a: S = Q . C^T on tcgen05 tensor cores with the
// query block resident in TMEM, the corpus streamed once from HBM by TMA, and a per-query
// running top-k kept in shared memory by the epilogue warps.
//
// Replaces the dense leg of collection.query.hybrid(...) / near_text(...) that the
// reference sends to Weaviate (server/routes/knowledge_base/weaviate_client.py:252-259,
// server/routes/incident_feedback/weaviate_client.py:286-291).
//
// Shape of one CTA (192 threads, persistent, 1 CTA / SM):
//   warp 0   TMA producer : corpus tiles [64 rows x 256 k] -> smem ring (SWIZZLE_128B)
//   warp 1   MMA issuer   : tcgen05.mma kind::f16, A = queries from TMEM (128 lanes =
//                           128 queries, dim/2 columns), B = corpus tile from smem,
//                           D = [128 queries x 64 rows] fp32 in one of two TMEM buffers
//   warps 2-5 epilogue    : thread r owns query r: tcgen05.ld its 64 scores, scale by the
//                           row's inverse norm, compare against its threshold, push the rare
//                           survivors to a pending list, drain that into a k-slot list.
// cta_group::2: a CTA pair shares every corpus tile -- each CTA TMA-loads 32 of the 64
// rows, the leader issues M=256 MMAs, each CTA's TMEM holds its own 128 queries.
// cta_group::1: M=128; when nq > 128 two CTAs take the same tiles for the two query halves.
//
// Threshold exchange.  A CTA sees only 1/74 of the corpus, so its own k-th best is a loose
// filter (~270 insertions per query per CTA over 1M rows).  Every CTA therefore publishes,
// per query, its best (or 2nd best) score so far; a thread periodically reads the values
// of up to 74 CTAs and takes the R-th largest (R * m >= k + slack): at least k + slack
// rows with a score >= that value exist somewhere, so nothing below it can reach the final
// top-k.  This certified global threshold cuts insertions to a handful per query.
#include <cuda.h>
#include "internal.h"
#include "ptx.cuh"

namespace aur {
using namespace ptx;

namespace {

struct SmemLayout {
  uint32_t stage_bytes, box_bytes;
  uint32_t off_list, off_pend, off_norm, off_bar, total;
};
__host__ __device__ inline SmemLayout make_layout(int cta_group, int num_stages, int ksel) {
  SmemLayout L;
  L.box_bytes = (kTcTileN / cta_group) * 128u;
  L.stage_bytes = L.box_bytes * kTcKbPerStage;
  uint32_t o = L.stage_bytes * num_stages;
  L.off_list = o;  o += static_cast<uint32_t>(ksel) * kTcQRows * 8u;
  L.off_pend = o;  o += kTcPendCap * kTcQRows * 8u;
  L.off_norm = o;  o += 4u * 2u * kTcTileN * 4u;
  L.off_bar = o;   o += (2u * kTcMaxStages + 2u + 2u + 1u) * 8u + 16u;
  L.total = o;
  return L;
}

// Per-thread (= per-query) selection state of the epilogue.
struct TopkState {
  uint64_t tau_key;   // smallest key in the full list
  float tau_local;    // its score (own k-th best)
  float tau_glob;     // threshold certified by the cross-CTA exchange
  float tau;          // max of the two: the admission filter
  int minpos, cnt, nfill;
};

// Fold this lane's pending candidates into its k-slot list.  Out of line on purpose: it
// runs a handful of times per query and must stay out of the hot loop's instruction stream.
// list_a / pend_a: shared addresses of slot 0 for this lane (slot stride 1024 B).
__device__ __noinline__ TopkState drain_pending(TopkState st, uint32_t list_a, uint32_t pend_a, int ksel) {
  constexpr uint32_t kStride = kTcQRows * 8u;
  for (int c = 0; c < st.cnt; ++c) {
    const uint64_t key = lds_u64(pend_a + c * kStride);
    bool rescan = false;
    if (st.nfill < ksel) {             // still filling: append, no scan
      sts_u64(list_a + st.nfill * kStride, key);
      rescan = (++st.nfill == ksel);
    } else if (key > st.tau_key) {     // replace the current minimum
      sts_u64(list_a + st.minpos * kStride, key);
      rescan = true;
    }
    if (rescan) {
      uint64_t m = lds_u64(list_a); int mp = 0;
#pragma unroll 8
      for (int t = 1; t < ksel; ++t) {
        const uint64_t v = lds_u64(list_a + t * kStride);
        if (v < m) { m = v; mp = t; }
      }
      st.tau_key = m; st.minpos = mp; st.tau_local = key_score(m);
    }
  }
  st.cnt = 0;
  st.tau = fmaxf(st.tau_local, st.tau_glob);
  return st;
}

// R-th largest of the values published for this query by up to kTcPubMax CTAs (bisection
// on the value; entries of other launches or not yet written read as NaN and are skipped).
// Returns -inf when fewer than R CTAs have published.
__device__ __noinline__ float exchange_threshold(const unsigned long long* pubq, int nuse, int R, uint32_t epoch) {
  float v[kTcPubMax];
  float lo = INFINITY, hi = -INFINITY;
  int nvalid = 0;
#pragma unroll
  for (int i = 0; i < kTcPubMax; ++i) {
    unsigned long long e = 0ull;
    if (i < nuse) e = __ldcg(pubq + static_cast<size_t>(i) * kTcQRows);
    const bool ok = static_cast<uint32_t>(e >> 32) == epoch;
    v[i] = ok ? __uint_as_float(static_cast<uint32_t>(e)) : __int_as_float(0x7FC00000);
    nvalid += ok ? 1 : 0;
    lo = fminf(lo, v[i]); hi = fmaxf(hi, v[i]);   // fminf / fmaxf skip NaN
  }
  if (nvalid < R) return -INFINITY;
  // invariant: count(v >= lo) >= R
#pragma unroll 1
  for (int round = 0; round < 12; ++round) {
    const float mid = 0.5f * (lo + hi);
    int c = 0;
#pragma unroll
    for (int i = 0; i < kTcPubMax; ++i) c += (v[i] >= mid) ? 1 : 0;
    if (c >= R) lo = mid; else hi = mid;
  }
  return lo;
}

constexpr uint32_t kDescHi = 0x40004040u;  // SBO = 1024 B, descriptor version 1, SWIZZLE_128B

template <int kCtaGroup>
__global__ void __launch_bounds__(kTcThreads, 1)
simtopk_tc_kernel(const __grid_constant__ CUtensorMap tmap, const TcParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // SWIZZLE_128B tiles need 1024-byte alignment; the runtime only guarantees 16.  Offsetting
  // the declared array (rather than round-tripping through an integer) keeps the compiler's
  // shared-address-space inference, i.e. LDS/STS instead of generic loads.
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);

  const SmemLayout L = make_layout(kCtaGroup, p.num_stages, p.ksel);
  uint64_t* list = reinterpret_cast<uint64_t*>(smem + L.off_list);   // [ksel][128]
  uint64_t* pend = reinterpret_cast<uint64_t*>(smem + L.off_pend);   // [kTcPendCap][128]
  float* normbuf = reinterpret_cast<float*>(smem + L.off_norm);      // [4][2][64]
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L.off_bar);
  uint64_t* full_bar = bars;                              // [kTcMaxStages]
  uint64_t* empty_bar = bars + kTcMaxStages;              // [kTcMaxStages]
  uint64_t* tmem_full = bars + 2 * kTcMaxStages;          // [2]
  uint64_t* tmem_empty = bars + 2 * kTcMaxStages + 2;     // [2]
  uint64_t* q_ready = bars + 2 * kTcMaxStages + 4;        // [1]
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 2 * kTcMaxStages + 5);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = (kCtaGroup == 2) ? cluster_ctarank() : 0u;

  // Work split.  list = which set of corpus tiles; qblock = which 128 queries.
  int qblock, my_list, n_lists;
  if constexpr (kCtaGroup == 2) {
    qblock = static_cast<int>(rank);
    my_list = blockIdx.x >> 1;
    n_lists = gridDim.x >> 1;
  } else {
    qblock = blockIdx.x % p.n_qblocks;
    my_list = blockIdx.x / p.n_qblocks;
    n_lists = gridDim.x / p.n_qblocks;
  }
  const int my_tiles = (p.n_tiles > my_list) ? (p.n_tiles - my_list + n_lists - 1) / n_lists : 0;
  const int kbs = p.dim / kTcKBlock;  // 128-byte k-blocks per row

  if constexpr (kCtaGroup == 2) cluster_sync_all();  // both CTAs resident before the paired TMEM alloc

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmap);
    for (int i = 0; i < p.num_stages; ++i) {
      mbar_init(&full_bar[i], kCtaGroup);  // leader's expect_tx arrive (+ the peer producer's arrive)
      mbar_init(&empty_bar[i], 1);         // one tcgen05.commit
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full[i], 1);               // one tcgen05.commit
      mbar_init(&tmem_empty[i], 4 * kCtaGroup);  // one arrive per epilogue warp (of both CTAs)
    }
    mbar_init(q_ready, 4 * kCtaGroup);
    fence_mbar_init();
  } else if (warp == 2) {
    tmem_alloc<kCtaGroup>(tmem_ptr_smem, 512);
    tmem_relinquish<kCtaGroup>();
  }
  tc_fence_before();
  if constexpr (kCtaGroup == 2) cluster_sync_all(); else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  if (warp == 0) {
    // ============================== TMA producer ==============================
    if (lane == 0) {
      const uint64_t hint = (kCtaGroup == 2 || p.n_qblocks == 1) ? kEvictFirst : kEvictNormal;
      int stage = 0; uint32_t phase = 0;
      for (int it = 0; it < my_tiles; ++it) {
        const int tile = my_list + it * n_lists;
        const int row0 = tile * kTcTileN + static_cast<int>(rank) * (kTcTileN / kCtaGroup);
        for (int kb0 = 0; kb0 < kbs; kb0 += kTcKbPerStage) {
          const int nkb = min(kTcKbPerStage, kbs - kb0);
          mbar_wait(&empty_bar[stage], phase ^ 1u);
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
    }
  } else if (warp == 1) {
    // ============================== MMA issuer ==============================
    // The whole warp walks the pipeline (so every operand stays warp-uniform and lives in
    // uniform registers); one elected lane issues the MMAs and their commits.
    if (rank == 0) {
      mbar_wait(q_ready, 0);  // queries of (both) CTAs are in TMEM
      tc_fence_after();
      constexpr uint32_t idesc = idesc_bf16_f32(128 * kCtaGroup, kTcTileN);
      constexpr uint32_t kBox16 = ((kTcTileN / kCtaGroup) * 128u) >> 4;  // box stride in descriptor units
      int stage = 0; uint32_t phase = 0;
      for (int it = 0; it < my_tiles; ++it) {
        const int b = it & 1;
        mbar_wait(&tmem_empty[b], ((static_cast<uint32_t>(it) >> 1) & 1u) ^ 1u);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + kTcAccCol0 + b * kTcTileN;
        for (int kb0 = 0; kb0 < kbs; kb0 += kTcKbPerStage) {
          const int nkb = min(kTcKbPerStage, kbs - kb0);
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t base_lo = (smem_u32(smem + static_cast<uint32_t>(stage) * L.stage_bytes) & 0x3FFFFu) >> 4;
          const uint32_t a_col = tmem_base + static_cast<uint32_t>(kb0) * 32u;
          if (elect_one()) {
#pragma unroll
            for (int j = 0; j < kTcKbPerStage; ++j) {
              if (j < nkb) {
#pragma unroll
                for (int k = 0; k < 4; ++k) {  // 4 x K=16 per 128-byte k-block
                  const uint32_t acc = (j | k) != 0 ? 1u : (kb0 != 0 ? 1u : 0u);
                  mma_ts_bf16<kCtaGroup>(d_tmem, a_col + j * 32 + k * 8, pack_u64(base_lo + j * kBox16 + k * 2, kDescHi),
                                         idesc, acc);
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
    }
  } else {
    // ============================== epilogue: per-query top-k ==============================
    const int quarter = warp & 3;              // TMEM lane quarter this warp may touch
    const int r = quarter * 32 + lane;         // TMEM lane == query row inside the CTA
    const int qglob = qblock * kTcQRows + r;   // query index inside this launch
    const uint32_t lane_addr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16);
    float* mynorm = normbuf + (warp - 2) * 2 * kTcTileN;

    // ---- 1. park this thread's query row in TMEM (bf16 pairs, K ascending along columns)
    {
      const uint4* src = reinterpret_cast<const uint4*>(p.q + static_cast<size_t>(qglob) * p.dim);
      for (int kb = 0; kb < kbs; ++kb) {
        uint32_t v[2][16];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          uint4 x = make_uint4(0, 0, 0, 0);
          if (qglob < p.nq) x = ldg_nc_v4(src + kb * 8 + i);
          v[i >> 2][(i & 3) * 4 + 0] = x.x; v[i >> 2][(i & 3) * 4 + 1] = x.y;
          v[i >> 2][(i & 3) * 4 + 2] = x.z; v[i >> 2][(i & 3) * 4 + 3] = x.w;
        }
        tmem_st_x16(lane_addr + kb * 32, v[0]);
        tmem_st_x16(lane_addr + kb * 32 + 16, v[1]);
      }
      tmem_wait_st();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if constexpr (kCtaGroup == 2) mbar_arrive_cluster(q_ready, 0); else mbar_arrive(q_ready);
      }
    }

    // ---- 2. running top-k state: list[t][r], t < ksel; the minimum is tracked once full
    const int ksel = p.ksel;
    const uint32_t list_a = smem_u32(list) + r * 8u;   // + t * 1024
    const uint32_t pend_a = smem_u32(pend) + r * 8u;   // + c * 1024
    for (int t = 0; t < ksel; ++t) sts_u64(list_a + t * (kTcQRows * 8u), kKeyEmpty);
    TopkState st;
    st.tau_key = kKeyEmpty; st.tau_local = -INFINITY; st.tau_glob = -INFINITY; st.tau = -INFINITY;
    st.minpos = 0; st.cnt = 0; st.nfill = 0;
    float top1 = -INFINITY, top2 = -INFINITY, published = -INFINITY;

    // exchange geometry: R-th largest of the m-th best of `nuse` CTAs is a valid threshold
    const int nuse = min(n_lists, kTcPubMax);
    const int xm = (ksel <= nuse) ? 1 : 2;
    const int xR = (ksel + xm - 1) / xm;
    const bool xchg = (p.pub != nullptr) && (xR <= nuse);
    unsigned long long* pubq =
        reinterpret_cast<unsigned long long*>(p.pub) + (static_cast<size_t>(qblock) * n_lists) * kTcQRows + r;

    // inverse norms of the first tile
    if (my_tiles > 0) {
      const int64_t rbase = static_cast<int64_t>(my_list) * kTcTileN;
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int64_t row = rbase + lane + 32 * h;
        mynorm[lane + 32 * h] = (row < p.n_rows) ? __ldg(p.inv_norm + row) : __int_as_float(0x7FC00000);
      }
      __syncwarp();
    }

    for (int it = 0; it < my_tiles; ++it) {
      const int tile = my_list + it * n_lists;
      const int row0 = tile * kTcTileN;
      const int b = it & 1;
      const float* nb = mynorm + b * kTcTileN;

      // ---- threshold exchange (tiles 1..7, then every power of two)
      if (xchg && it >= 1 && (it < 8 || (it & (it - 1)) == 0)) {
        st.tau_glob = fmaxf(st.tau_glob, exchange_threshold(pubq, nuse, xR, p.epoch));
        st.tau = fmaxf(st.tau_local, st.tau_glob);
      }

      mbar_wait(&tmem_full[b], (static_cast<uint32_t>(it) >> 1) & 1u);
      tc_fence_after();
      uint32_t acc[4][16];
#pragma unroll
      for (int c = 0; c < 4; ++c) tmem_ld_x16(lane_addr + kTcAccCol0 + b * kTcTileN + c * 16, acc[c]);
      tmem_wait_ld();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {  // accumulator drained into registers: hand the buffer back to the MMA warp
        if constexpr (kCtaGroup == 2) mbar_arrive_cluster(&tmem_empty[b], 0); else mbar_arrive(&tmem_empty[b]);
      }

      // prefetch the next tile's inverse norms (latency hidden behind this tile's work)
      float nn0 = __int_as_float(0x7FC00000), nn1 = nn0;
      if (it + 1 < my_tiles) {
        const int64_t rbase = static_cast<int64_t>(tile + n_lists) * kTcTileN;
        if (rbase + lane < p.n_rows) nn0 = __ldg(p.inv_norm + rbase + lane);
        if (rbase + lane + 32 < p.n_rows) nn1 = __ldg(p.inv_norm + rbase + lane + 32);
      }

      // Fast path: scale by 1/|c_j| in place and keep one running max per 16 scores; a
      // chunk is looked at score by score only if some lane's max reaches its threshold.
      // (NaN norm = tombstone / out of range: fmaxf drops it and `>=` rejects it.)
      float cmax[4];
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        float m = -INFINITY;
#pragma unroll
        for (int j4 = 0; j4 < 4; ++j4) {
          const float4 nv = *reinterpret_cast<const float4*>(nb + c * 16 + j4 * 4);
          const float s0 = __uint_as_float(acc[c][j4 * 4 + 0]) * nv.x;
          const float s1 = __uint_as_float(acc[c][j4 * 4 + 1]) * nv.y;
          const float s2 = __uint_as_float(acc[c][j4 * 4 + 2]) * nv.z;
          const float s3 = __uint_as_float(acc[c][j4 * 4 + 3]) * nv.w;
          acc[c][j4 * 4 + 0] = __float_as_uint(s0); acc[c][j4 * 4 + 1] = __float_as_uint(s1);
          acc[c][j4 * 4 + 2] = __float_as_uint(s2); acc[c][j4 * 4 + 3] = __float_as_uint(s3);
          m = fmaxf(fmaxf(m, fmaxf(s0, s1)), fmaxf(s2, s3));
        }
        cmax[c] = m;
      }
      if (p.dbg_scores != nullptr && it == 0) {
#pragma unroll
        for (int c = 0; c < 4; ++c)
#pragma unroll
          for (int j = 0; j < 16; ++j)
            p.dbg_scores[(static_cast<size_t>(blockIdx.x) * kTcQRows + r) * kTcTileN + c * 16 + j] =
                __uint_as_float(acc[c][j]);
      }

#pragma unroll
      for (int c = 0; c < 4; ++c) {
        if (__any_sync(0xffffffffu, cmax[c] >= st.tau)) {
          if (__any_sync(0xffffffffu, st.cnt > kTcPendCap - 16)) st = drain_pending(st, list_a, pend_a, ksel);
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const float s = __uint_as_float(acc[c][j]);
            if (s >= st.tau) {
              sts_u64(pend_a + st.cnt * (kTcQRows * 8u), make_key(s, row0 + c * 16 + j));
              ++st.cnt;
              if (s > top1) { top2 = top1; top1 = s; } else if (s > top2) top2 = s;
            }
          }
        }
      }

      // publish this CTA's m-th best for the exchange (monotone, so stale reads stay valid)
      const float pv = (xm == 1) ? top1 : top2;
      if (xchg && pv > published) {
        published = pv;
        __stcg(pubq + static_cast<size_t>(my_list) * kTcQRows,
               (static_cast<unsigned long long>(p.epoch) << 32) | __float_as_uint(pv));
      }

      float* nnext = mynorm + (b ^ 1) * kTcTileN;
      nnext[lane] = nn0; nnext[lane + 32] = nn1;
      __syncwarp();
    }
    st = drain_pending(st, list_a, pend_a, ksel);

    // ---- 3. append the survivors (score >= the certified threshold) to this query's
    //         compact candidate row
    if (xchg && my_tiles > 0) st.tau_glob = fmaxf(st.tau_glob, exchange_threshold(pubq, nuse, xR, p.epoch));
    {
      const uint32_t thr = f32_to_ord(st.tau_glob);  // tau_glob only: ties at the threshold are kept
      const size_t cap = static_cast<size_t>(n_lists) * ksel;
      uint64_t* out = p.cand + static_cast<size_t>(qglob) * cap;
      for (int t = 0; t < ksel; ++t) {
        const uint64_t key = lds_u64(list_a + t * (kTcQRows * 8u));
        if (key != kKeyEmpty && static_cast<uint32_t>(key >> 32) >= thr) {
          const uint32_t slot = atomicAdd(p.cand_count + qglob, 1u);
          out[slot] = key;
        }
      }
    }
  }

  // ============================== teardown ==============================
  __syncwarp();
  tc_fence_before();
  if constexpr (kCtaGroup == 2) cluster_sync_all(); else __syncthreads();
  tc_fence_after();
  if (warp == 2) tmem_dealloc<kCtaGroup>(tmem_base, 512);
}

}  // namespace

size_t tc_smem_bytes(int cta_group, int num_stages, int ksel) {
  return make_layout(cta_group, num_stages, ksel).total + 1024;  // + alignment slack
}

int tc_pick_stages(int cta_group, int ksel, size_t smem_limit) {
  for (int s = kTcMaxStages; s >= 2; --s)
    if (tc_smem_bytes(cta_group, s, ksel) <= smem_limit) return s;
  return 0;
}

cudaError_t tc_launch(int cta_group, int grid, const void* tmap, const TcParams& p, size_t smem, cudaStream_t s) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(kTcThreads);
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
  cudaError_t e;
  if (cta_group == 2) {
    e = cudaFuncSetAttribute(simtopk_tc_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    if (e != cudaSuccess) return e;
    return cudaLaunchKernelEx(&cfg, simtopk_tc_kernel<2>, tm, p);
  }
  e = cudaFuncSetAttribute(simtopk_tc_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
  if (e != cudaSuccess) return e;
  return cudaLaunchKernelEx(&cfg, simtopk_tc_kernel<1>, tm, p);
}

}  // namespace aur
