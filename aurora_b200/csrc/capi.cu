// Host runtime + C ABI (include/aurora_b200.h).  Owns one corpus shard in HBM:
//   rows      [capacity, dim]  bf16 or f32, row-major (K-major for the MMA, 128-bit
//                              coalesced for everything else)
//   inv_norm  [capacity] f32   1/|row|; 0 for zero rows; NaN marks a tombstone
//   ids       [capacity] i64   caller ids;  user / org [capacity] i32 tenant codes
// plus grow-only scratch for candidate lists.  No CPU compute path exists here: without
// a CUDA device every entry point fails with AUR_ERR_NO_DEVICE.
//
// Concurrency (BASELINE config 5: streaming ingest while queries run; the reference's callers are 8 gunicorn
// threads + 4 Celery children, docker-compose.yaml:191,283-285).  The shard is append-only with a PUBLISHED row
// count: a writer copies rows / ids / tenant codes / inverse norms into [rows_pub, rows_pub + n) on the ingest
// stream, waits for them to land, and only then stores rows_pub + n (release).  A search loads rows_pub once
// (acquire) when it is enqueued and scans exactly that prefix -- rows behind it are masked by the kernels' own
// n_rows bound -- so every answer is the top-k of a consistent prefix, which it can report
// (aur_search_ex).  Searches never take the writers' lock; each in-flight search owns a SearchCtx (candidate
// lists, threshold-exchange table, staging buffers, events, stream), so several can be enqueued at once.
// Compaction and export are the only exclusive operations.
#include <cuda.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <string.h>

#include <atomic>
#include <memory>
#include <mutex>
#include <shared_mutex>
#include <string>
#include <unordered_map>
#include <algorithm>
#include <vector>

#include "../../include/aurora_b200.h"
#include "internal.h"

using namespace aur;

namespace {

thread_local std::string g_err;
int fail(int code, const char* fmt, ...) {
  char buf[512];
  va_list ap; va_start(ap, fmt); vsnprintf(buf, sizeof buf, fmt, ap); va_end(ap);
  g_err = buf;
  return code;
}
#define CU_TRY(expr)                                                                               \
  do {                                                                                             \
    cudaError_t e_ = (expr);                                                                       \
    if (e_ != cudaSuccess) return fail(AUR_ERR_CUDA, "%s: %s (%s:%d)", #expr, cudaGetErrorString(e_), \
                                       __FILE__, __LINE__);                                        \
  } while (0)

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

template <typename T>
struct DevBuf {  // grow-only device scratch
  T* p = nullptr; size_t n = 0;
  cudaError_t reserve(size_t want) {
    if (want <= n) return cudaSuccess;
    if (p) cudaFree(p);
    p = nullptr; n = 0;
    cudaError_t e = cudaMalloc(&p, want * sizeof(T));
    if (e == cudaSuccess) n = want;
    return e;
  }
  void release() { if (p) cudaFree(p); p = nullptr; n = 0; }
};

}  // namespace

namespace aur {
int report_error(int code, const char* fmt, ...) {
  char buf[512];
  va_list ap; va_start(ap, fmt); vsnprintf(buf, sizeof buf, fmt, ap); va_end(ap);
  g_err = buf;
  return code;
}
int encode_tmap_2d_bf16(void* tmap, const void* base, uint64_t cols, uint64_t rows, uint64_t row_stride_bytes,
                        uint32_t box_cols, uint32_t box_rows) {
  EncodeTiledFn enc = get_encode_fn();
  if (!enc) return -1;
  cuuint64_t gdim[2] = {cols, rows};
  cuuint64_t gstride[1] = {row_stride_bytes};
  cuuint32_t box[2] = {box_cols, box_rows};
  cuuint32_t estr[2] = {1, 1};
  return static_cast<int>(enc(static_cast<CUtensorMap*>(tmap), CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2,
                              const_cast<void*>(base), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                              CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE));
}
}  // namespace aur

// Scratch and bookkeeping of ONE in-flight search.  Device-pointer searches are bound to the caller's stream
// (same stream -> same context -> stream order protects the scratch); host-buffer searches take a context from a
// pool and run on its own stream, so N threads can search at once.
struct SearchCtx {
  std::mutex mu;                  // one enqueue at a time
  cudaStream_t own_stream = nullptr;
  cudaStream_t bound = nullptr;   // caller stream this context serves (nullptr = pool context)
  cudaEvent_t ev_begin = nullptr, ev_k0 = nullptr, ev_k1 = nullptr, ev_fin = nullptr, ev_end = nullptr;
  bool have_timing = false;
  int last_kernel = 0, last_launches = 0;
  int64_t snapshot_rows = 0;
  uint32_t epoch = 0;
  DevBuf<uint64_t> cand_a, cand_b;
  DevBuf<uint64_t> pub;          // tcgen05 kernel's cross-CTA threshold exchange
  DevBuf<uint32_t> cand_count;   // compacted candidates per query (self-resetting)
  DevBuf<uint32_t> d_epoch;      // the tcgen05 kernel's launch counter, device resident (CUDA-graph replays advance it)
  DevBuf<float> score_chunk;
  DevBuf<float> masked_inv;      // inverse norms with the invisible rows turned into NaN (tenant scope / id subset)
  DevBuf<int32_t> allow_rows;    // subset search: rows that stay visible
  DevBuf<uint32_t> row_mask;     // per-query tenant scopes on the tensor path: bit s = scope s of the batch sees the row
  DevBuf<int32_t> scope_tab, q_scope;   // the batch's distinct {user, org} scopes (<= 32) and every query's scope index
  DevBuf<uint8_t> stage_q;       // host-entry staging: queries
  DevBuf<int32_t> stage_quser, stage_qorg;
  DevBuf<float> stage_scores;
  DevBuf<int64_t> stage_ids;
  void release() {
    cand_a.release(); cand_b.release(); pub.release(); cand_count.release(); d_epoch.release(); score_chunk.release(); masked_inv.release();
    allow_rows.release(); row_mask.release(); scope_tab.release(); q_scope.release(); stage_q.release(); stage_quser.release(); stage_qorg.release(); stage_scores.release();
    stage_ids.release();
    if (ev_begin) cudaEventDestroy(ev_begin);
    if (ev_k0) cudaEventDestroy(ev_k0);
    if (ev_k1) cudaEventDestroy(ev_k1);
    if (ev_fin) cudaEventDestroy(ev_fin);
    if (ev_end) cudaEventDestroy(ev_end);
    if (own_stream) cudaStreamDestroy(own_stream);
  }
};

struct aur_index {
  std::shared_mutex rw;          // shared: searches and appends; exclusive: compaction, export, close
  std::mutex mu;                 // host metadata: id2row, live, options, the context pool
  std::mutex mu_write;           // one writer at a time (append / remove)
  int device = 0, dim = 0, dtype = 0;
  int64_t capacity = 0, live = 0;
  std::atomic<int64_t> rows_pub{0};   // rows visible to searches (published after the data landed)
  size_t elt = 2;
  void* d_rows = nullptr;
  float* d_inv_norm = nullptr;
  int64_t* d_ids = nullptr;
  int32_t* d_user = nullptr;
  int32_t* d_org = nullptr;
  std::unordered_map<int64_t, int64_t> id2row;
  cudaStream_t stream = nullptr;         // the index's own stream: device-pointer calls with stream == NULL
  cudaStream_t ingest_stream = nullptr;  // host appends / tombstones, lowest priority so queries overtake them
  CUtensorMap tmap[2];  // box rows 64 (cta_group::1) and 32 (cta_group::2)
  bool tmap_ok = false;
  int sm_count = 0;
  size_t smem_optin = 0;
  int opt_kernel = AUR_KERNEL_AUTO;
  int opt_dbg_flags = 0;
  int opt_epi_groups = 0;        // 0 = auto
  std::vector<std::unique_ptr<SearchCtx>> ctxs;
  std::vector<SearchCtx*> free_ctxs;   // idle pool contexts
  SearchCtx* last_ctx = nullptr;       // context of the most recently enqueued search (aur_get_stats)
};

namespace {

int build_tmaps(aur_index* ix) {
  ix->tmap_ok = false;
  if (ix->dtype != AUR_BF16 || ix->dim % kTcKBlock != 0 || ix->dim > kTcMaxDim) return AUR_OK;  // SIMT only
  EncodeTiledFn enc = get_encode_fn();
  if (!enc) return fail(AUR_ERR_CUDA, "cuTensorMapEncodeTiled entry point not found");
  for (int g = 1; g <= 2; ++g) {
    cuuint64_t gdim[2] = {static_cast<cuuint64_t>(ix->dim), static_cast<cuuint64_t>(ix->capacity)};
    cuuint64_t gstride[1] = {static_cast<cuuint64_t>(ix->dim) * 2};
    cuuint32_t box[2] = {static_cast<cuuint32_t>(kTcKBlock), static_cast<cuuint32_t>(kTcTileN / g)};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(&ix->tmap[g - 1], CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, ix->d_rows, gdim, gstride, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(AUR_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d)", static_cast<int>(r));
  }
  ix->tmap_ok = true;
  return AUR_OK;
}

bool tc_shape_ok(const aur_index* ix, int k, bool filtered) {
  if (!ix->tmap_ok || filtered || k > kMaxK) return false;
  // candidate lists (k + slack per query) and, past 768 dims, part of the queries share the SM's
  // shared memory with the TMA ring: large k at large dim leaves no room for a pipeline
  return tc_pick_stages(2, 1, k + kSlack, ix->dim, ix->smem_optin) >= 2;
}

int ctx_init(aur_index* ix, SearchCtx* c) {
  int lo = 0, hi = 0;
  CU_TRY(cudaDeviceGetStreamPriorityRange(&lo, &hi));   // hi = numerically lowest = highest priority
  CU_TRY(cudaStreamCreateWithPriority(&c->own_stream, cudaStreamNonBlocking, hi));
  CU_TRY(cudaEventCreate(&c->ev_begin)); CU_TRY(cudaEventCreate(&c->ev_k0));
  CU_TRY(cudaEventCreate(&c->ev_k1));    CU_TRY(cudaEventCreate(&c->ev_end));
  CU_TRY(cudaEventCreate(&c->ev_fin));
  (void)ix;
  return AUR_OK;
}

// Context for a search on the caller's stream `s` (bound) or, with s == nullptr, an idle pool context.
int acquire_ctx(aur_index* ix, cudaStream_t s, SearchCtx** out) {
  std::lock_guard<std::mutex> lk(ix->mu);
  if (s) {
    for (auto& c : ix->ctxs)
      if (c->bound == s) { *out = c.get(); return AUR_OK; }
  } else if (!ix->free_ctxs.empty()) {
    *out = ix->free_ctxs.back(); ix->free_ctxs.pop_back();
    return AUR_OK;
  }
  std::unique_ptr<SearchCtx> c(new SearchCtx());
  int rc = ctx_init(ix, c.get());
  if (rc != AUR_OK) { c->release(); return rc; }
  c->bound = s;
  *out = c.get();
  ix->ctxs.push_back(std::move(c));
  return AUR_OK;
}
void release_ctx(aur_index* ix, SearchCtx* c) {
  std::lock_guard<std::mutex> lk(ix->mu);
  if (!c->bound) ix->free_ctxs.push_back(c);
}

// Runs one block of <= 256 queries through the tcgen05 kernel over the first n_rows rows.  Leaves candidate keys
// in c->cand_a as [nqb_pad, n_lists, ksel]; returns n_lists.
int run_tc_block(aur_index* ix, SearchCtx* c, int cta_group, const void* q_dev, int nqb, int ksel, int64_t n_rows, float* dbg,
                 int* n_lists_out, cudaStream_t s, const float* inv_norm = nullptr, const uint32_t* row_mask = nullptr,
                 const int32_t* q_scope = nullptr) {
  int n_qblocks = (nqb > kTcQRows) ? 2 : 1;
  if (cta_group == 2 && n_qblocks != 2) {
    // a pair works on 256 query rows.  A short tail block normally runs as single CTAs; when their larger
    // TMA stages do not fit next to the lists (large k at dim > 768) it runs as a pair with a padding block
    if (tc_pick_stages(1, 1, ksel, ix->dim, ix->smem_optin) >= 2) cta_group = 1; else n_qblocks = 2;
  }
  const int pairs = ix->sm_count / 2;
  int n_super = 1;   // query super-blocks of 256 (CTA pairs that walk the same tiles side by side, sharing them through L2)
  if (cta_group == 2) {
    n_super = (nqb + 2 * kTcQRows - 1) / (2 * kTcQRows);
    // the threshold exchange needs ceil(ksel / tile sets) <= 4 rows vouched for per CTA
    while (n_super > 1 && (ksel + pairs / n_super - 1) / (pairs / n_super) > 4) --n_super;
    if (nqb > n_super * 2 * kTcQRows) return fail(AUR_ERR_INVALID, "internal: query block larger than the launch geometry");
    n_qblocks = 2 * n_super;
  }
  int grid = (cta_group == 2) ? (pairs / n_super) * n_super * 2 : (ix->sm_count & ~1);
  const int n_tsets = (cta_group == 2) ? pairs / n_super : grid / n_qblocks;
  // epilogue groups: 1 by default; 2 (alternating tiles) stays selectable for experiments
  int epi_groups = ix->opt_epi_groups;
  if (epi_groups == 0) epi_groups = 1;   // measured: one group + a deeper TMA ring (11 stages) beats two groups + 8
  const int stages = tc_pick_stages(cta_group, epi_groups, ksel, ix->dim, ix->smem_optin);
  if (stages < 2) return fail(AUR_ERR_UNSUPPORTED, "k too large for the tcgen05 path's shared memory");
  const size_t smem = tc_smem_bytes(cta_group, epi_groups, stages, ksel, ix->dim);
  const int n_lists = n_tsets * epi_groups;   // candidate lists per query
  const size_t ncand = static_cast<size_t>(n_qblocks) * kTcQRows * n_lists * ksel;
  CU_TRY(c->cand_a.reserve(ncand));
  const size_t npub = static_cast<size_t>(n_qblocks) * kTcQRows * (((n_tsets + 1) & ~1) + 1);
  if (npub > c->pub.n) {
    CU_TRY(c->pub.reserve(npub));
    CU_TRY(cudaMemsetAsync(c->pub.p, 0, npub * 8, s));  // epoch 0 is never used by a launch
  }
  if (c->cand_count.n < 8 * kTcQRows) {
    CU_TRY(c->cand_count.reserve(8 * kTcQRows));
    CU_TRY(cudaMemsetAsync(c->cand_count.p, 0, 8 * kTcQRows * 4, s));
  }
  if (c->d_epoch.n == 0) {
    static const uint32_t one = 1;
    CU_TRY(c->d_epoch.reserve(1));
    CU_TRY(cudaMemcpyAsync(c->d_epoch.p, &one, 4, cudaMemcpyHostToDevice, s));
  }
  ++c->epoch;
  TcParams p;
  p.q = static_cast<const __nv_bfloat16*>(q_dev);
  p.inv_norm = inv_norm ? inv_norm : ix->d_inv_norm;
  p.row_mask = row_mask; p.q_scope = q_scope;
  p.cand = c->cand_a.p;
  p.cand_count = c->cand_count.p;
  p.dbg_scores = dbg;
  p.pub = c->pub.p;
  // searches: device-resident counter, advanced by the finalize kernel that follows; the bring-up entry point (no
  // finalize behind it) tags its entries from a disjoint range on the host
  p.epoch = 0x80000000u | c->epoch;
  p.epoch_ptr = dbg ? nullptr : c->d_epoch.p;
  p.n_rows = n_rows;
  p.nq = nqb; p.dim = ix->dim; p.ksel = ksel; p.n_lists = n_lists; p.n_qblocks = n_qblocks;
  p.num_stages = stages;
  p.dbg_flags = ix->opt_dbg_flags;
  p.n_tiles = static_cast<int>((n_rows + kTcTileN - 1) / kTcTileN);
  CU_TRY(tc_launch(cta_group, epi_groups, grid, &ix->tmap[cta_group - 1], p, smem, s));
  *n_lists_out = n_lists;
  return AUR_OK;
}

// Visibility of a search besides tombstones: per-query tenant codes (generic kernel), one tenant scope for the
// whole batch, or an explicit list of visible rows -- the last two fold into the row scale (NaN = invisible), so
// the tensor-core kernel serves them.
struct Scope {
  const int32_t* q_user = nullptr;       // device, per query (nullptr = no tenant filter)
  const int32_t* q_org = nullptr;
  const int32_t* uniform = nullptr;      // host {user, org}: every query of the batch carries this scope
  const int32_t* allow_rows = nullptr;   // device: rows that stay visible (subset search)
  int64_t n_allow = -1;                  // -1 = no subset
  const int32_t* scope_tab = nullptr;    // device [n_scopes][2]: the batch's distinct tenant scopes, when there are <= 32
  const int32_t* q_scope = nullptr;      // device [nq]: scope index of every query
  int n_scopes = 0;
};

// Enqueues one search over the published prefix `n_rows` on stream s using context c.
int search_enqueue(aur_index* ix, SearchCtx* c, const void* q_dev, int nq, int k, const Scope& sc, int64_t n_rows,
                   float* scores, int64_t* ids, double* scores64, cudaStream_t s,
                   const FinalizeArgs::ExchangeOut* ex = nullptr) {
  if (nq <= 0 || k <= 0) return fail(AUR_ERR_INVALID, "nq and k must be positive");
  if (k > kMaxK) return fail(AUR_ERR_UNSUPPORTED, "k > %d", kMaxK);
  if (nq > 65535) return fail(AUR_ERR_UNSUPPORTED, "nq > 65535: split the batch");
  const bool subset = sc.n_allow >= 0;
  const bool scoped_tc = sc.n_scopes > 0 && !subset && sc.uniform == nullptr;        // per-query scopes through bit masks
  const bool filtered = sc.q_user != nullptr && sc.uniform == nullptr && !subset && !scoped_tc;   // else: generic kernel only
  const int ksel = k + kSlack;
  int kernel = ix->opt_kernel;
  if (kernel == AUR_KERNEL_AUTO) kernel = tc_shape_ok(ix, k, filtered) ? AUR_KERNEL_TC2 : AUR_KERNEL_SIMT;
  if (kernel != AUR_KERNEL_SIMT && !tc_shape_ok(ix, k, filtered))
    return fail(AUR_ERR_UNSUPPORTED, "tcgen05 path needs bf16, dim %% 64 == 0, dim <= %d, no per-query tenant filter, and k small "
                "enough for its shared-memory lists at this dim", kTcMaxDim);
  c->last_kernel = kernel;
  c->last_launches = 0;
  c->snapshot_rows = n_rows;
  CU_TRY(cudaEventRecord(c->ev_begin, s));
  bool k_timed = false;
  const float* inv = nullptr;   // masked inverse norms (nullptr = the shard's own)
  if (subset && n_rows > 0) {
    CU_TRY(c->masked_inv.reserve(static_cast<size_t>(ix->capacity) + 64));
    CU_TRY(launch_fill_f32(c->masked_inv.p, nanf(""), n_rows, s));
    CU_TRY(launch_scatter_inv_norm(ix->d_inv_norm, sc.allow_rows, sc.n_allow, n_rows, c->masked_inv.p, s));
    c->last_launches += 2;
    inv = c->masked_inv.p;
  } else if (kernel != AUR_KERNEL_SIMT && scoped_tc && n_rows > 0) {              // up to 32 scopes in the batch: row bit masks
    CU_TRY(c->row_mask.reserve(static_cast<size_t>(ix->capacity) + 64));
    CU_TRY(launch_row_scope_mask(ix->d_user, ix->d_org, sc.scope_tab, sc.n_scopes, n_rows, c->row_mask.p, s));
    ++c->last_launches;
  } else if (kernel != AUR_KERNEL_SIMT && sc.q_user != nullptr && sc.uniform != nullptr && n_rows > 0) {   // one scope for the whole batch
    CU_TRY(c->masked_inv.reserve(static_cast<size_t>(ix->capacity) + 64));
    CU_TRY(launch_mask_inv_norm(ix->d_inv_norm, ix->d_user, ix->d_org, sc.uniform[0], sc.uniform[1], n_rows, c->masked_inv.p, s));
    ++c->last_launches;
    inv = c->masked_inv.p;
  }

  // queries per launch: the generic kernel takes 1024; the tcgen05 kernel 256 per CTA pair and up to four pairs side by
  // side on the same corpus tiles (fewer when k + slack is too large for the threshold exchange of that geometry)
  int qstep = 1024;
  if (kernel == AUR_KERNEL_TC1) qstep = 2 * kTcQRows;
  else if (kernel == AUR_KERNEL_TC2) {
    const int pairs = ix->sm_count / 2;
    int n_super = 4;
    while (n_super > 1 && (ksel + pairs / n_super - 1) / (pairs / n_super) > 4) --n_super;
    qstep = n_super * 2 * kTcQRows;
  }
  for (int q0 = 0; q0 < nq; q0 += qstep) {
    const int nqb = (nq - q0 < qstep) ? nq - q0 : qstep;
    const uint8_t* qb = static_cast<const uint8_t*>(q_dev) + static_cast<size_t>(q0) * ix->dim * ix->elt;
    int n_lists = 0;
    uint64_t* cur = nullptr;
    if (kernel == AUR_KERNEL_SIMT) {
      if (n_rows == 0) {
        n_lists = 1;
        CU_TRY(c->cand_a.reserve(static_cast<size_t>(nqb) * ksel));
        CU_TRY(cudaMemsetAsync(c->cand_a.p, 0, static_cast<size_t>(nqb) * ksel * 8, s));
      } else {
        n_lists = static_cast<int>((n_rows + kSimtSeg - 1) / kSimtSeg);
        CU_TRY(c->cand_a.reserve(static_cast<size_t>(nqb) * n_lists * ksel));
        const int64_t chunk = 16 * kSimtSeg;  // 32768 rows of scores at a time
        CU_TRY(c->score_chunk.reserve(static_cast<size_t>(nqb) * chunk));
        FilterArgs f{ix->d_user, ix->d_org, nullptr, nullptr};
        if (!subset && sc.q_user && !inv) { f.q_user = sc.q_user + q0; f.q_org = sc.q_org ? sc.q_org + q0 : nullptr; }
        if (!k_timed) CU_TRY(cudaEventRecord(c->ev_k0, s));
        for (int64_t r0 = 0; r0 < n_rows; r0 += chunk) {
          const int64_t nr = (n_rows - r0 < chunk) ? n_rows - r0 : chunk;
          CU_TRY(launch_simt_scores(qb, ix->d_rows, ix->dtype, ix->dim, nqb, r0, nr, n_rows, inv ? inv : ix->d_inv_norm, f,
                                    c->score_chunk.p, s));
          CU_TRY(launch_simt_select(c->score_chunk.p, nqb, r0, nr, ksel, c->cand_a.p, n_lists,
                                    static_cast<int>(r0 / kSimtSeg), s));
          c->last_launches += 2;
        }
        if (!k_timed) { CU_TRY(cudaEventRecord(c->ev_k1, s)); k_timed = true; }
      }
      cur = c->cand_a.p;
    } else {
      if (!k_timed) CU_TRY(cudaEventRecord(c->ev_k0, s));
      const bool use_mask = scoped_tc && n_rows > 0;
      int rc = run_tc_block(ix, c, kernel == AUR_KERNEL_TC1 ? 1 : 2, qb, nqb, ksel, n_rows, nullptr, &n_lists, s, inv,
                            use_mask ? c->row_mask.p : nullptr, use_mask ? sc.q_scope + q0 : nullptr);
      if (rc != AUR_OK) return rc;
      if (!k_timed) { CU_TRY(cudaEventRecord(c->ev_k1, s)); k_timed = true; }
      c->last_launches += 1;
      cur = c->cand_a.p;
    }
    // dense candidate lists (SIMT path): fold until one sort of <= 4096 keys finishes the
    // job.  The tcgen05 kernel already compacted its survivors per query.
    const bool compact = kernel != AUR_KERNEL_SIMT;
    bool in_a = true;
    while (!compact && static_cast<int64_t>(n_lists) * ksel > 4096) {
      const int group = 4096 / ksel;
      const int n_groups = (n_lists + group - 1) / group;
      DevBuf<uint64_t>& dst = in_a ? c->cand_b : c->cand_a;
      // rows of cand are indexed by the query position inside the block (TC pads to 128/256)
      CU_TRY(dst.reserve(static_cast<size_t>(nqb) * n_groups * ksel));
      CU_TRY(launch_reduce_lists(cur, nqb, n_lists, ksel, group, dst.p, s));
      c->last_launches += 1;
      cur = dst.p; n_lists = n_groups; in_a = !in_a;
    }
    FinalizeArgs fa{};
    fa.cand = cur; fa.n_lists = n_lists; fa.ksel = ksel;
    fa.counts = compact ? c->cand_count.p : nullptr;
    fa.epoch_bump = compact ? c->d_epoch.p : nullptr;
    fa.q = qb; fa.rows = ix->d_rows; fa.dtype = ix->dtype; fa.dim = ix->dim; fa.nq = nqb; fa.k = k;
    fa.ids = ix->d_ids;
    fa.out_scores = scores ? scores + static_cast<size_t>(q0) * k : nullptr;
    fa.out_ids = ids ? ids + static_cast<size_t>(q0) * k : nullptr;
    fa.out_scores64 = scores64 ? scores64 + static_cast<size_t>(q0) * k : nullptr;
    if (ex) { fa.ex = *ex; fa.ex.q0 = q0; }
    CU_TRY(launch_finalize(fa, s));
    c->last_launches += 1;
  }
  if (!k_timed) { CU_TRY(cudaEventRecord(c->ev_k0, s)); CU_TRY(cudaEventRecord(c->ev_k1, s)); }
  CU_TRY(cudaEventRecord(c->ev_fin, s));
  CU_TRY(cudaEventRecord(c->ev_end, s));
  c->have_timing = true;
  {
    std::lock_guard<std::mutex> lk(ix->mu);
    ix->last_ctx = c;
  }
  return AUR_OK;
}

// Append n rows behind the published prefix and publish them once they have landed.  Caller holds mu_write.
int add_common(aur_index* ix, const void* rows, bool rows_on_device, const int64_t* ids, const int32_t* users,
               const int32_t* orgs, int64_t n, cudaStream_t s) {
  if (n < 0) return fail(AUR_ERR_INVALID, "n < 0");
  if (n == 0) return AUR_OK;
  if (!rows || !ids) return fail(AUR_ERR_INVALID, "rows and ids are required");
  const int64_t base = ix->rows_pub.load(std::memory_order_relaxed);   // writers are serialised: nobody else moves it
  if (base + n > ix->capacity)
    return fail(AUR_ERR_NOMEM, "shard full: %lld + %lld > capacity %lld (aur_compact reclaims tombstones)", (long long)base,
                (long long)n, (long long)ix->capacity);
  for (int64_t i = 0; i < n; ++i)
    if (ids[i] < 0) return fail(AUR_ERR_INVALID, "ids must be >= 0");
  uint8_t* dst = static_cast<uint8_t*>(ix->d_rows) + static_cast<size_t>(base) * ix->dim * ix->elt;
  CU_TRY(cudaMemcpyAsync(dst, rows, static_cast<size_t>(n) * ix->dim * ix->elt,
                         rows_on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, s));
  CU_TRY(cudaMemcpyAsync(ix->d_ids + base, ids, static_cast<size_t>(n) * 8, cudaMemcpyHostToDevice, s));
  std::vector<int32_t> fill;
  if (!users) { fill.assign(static_cast<size_t>(n), 0); users = fill.data(); }
  CU_TRY(cudaMemcpyAsync(ix->d_user + base, users, static_cast<size_t>(n) * 4, cudaMemcpyHostToDevice, s));
  std::vector<int32_t> fill2;
  if (!orgs) { fill2.assign(static_cast<size_t>(n), -1); orgs = fill2.data(); }
  CU_TRY(cudaMemcpyAsync(ix->d_org + base, orgs, static_cast<size_t>(n) * 4, cudaMemcpyHostToDevice, s));
  CU_TRY(launch_row_inv_norms(dst, ix->dtype, ix->dim, n, ix->d_inv_norm + base, s));
  // upsert: an id that already exists loses its old row (weaviate_client.py:172 uuid5 semantics).  The old rows
  // are tombstoned before the new ones are published: a racing search may briefly miss the object, never see it twice.
  std::vector<int64_t> dead;
  {
    std::lock_guard<std::mutex> lk(ix->mu);
    for (int64_t i = 0; i < n; ++i) {
      auto it = ix->id2row.find(ids[i]);
      if (it != ix->id2row.end()) { dead.push_back(it->second); it->second = base + i; }
      else { ix->id2row.emplace(ids[i], base + i); ++ix->live; }
    }
  }
  const float nanv = nanf("");
  for (int64_t row : dead) CU_TRY(cudaMemcpyAsync(ix->d_inv_norm + row, &nanv, 4, cudaMemcpyHostToDevice, s));
  CU_TRY(cudaStreamSynchronize(s));  // the data has landed (and the host staging vectors may go out of scope)
  ix->rows_pub.store(base + n, std::memory_order_release);
  return AUR_OK;
}

int check_search_args(aur_index* ix, const void* q, int32_t nq, int32_t k, const void* s_out, const void* i_out) {
  if (!ix || !q || !s_out || !i_out) return fail(AUR_ERR_INVALID, "null argument");
  if (nq <= 0 || k <= 0) return fail(AUR_ERR_INVALID, "nq and k must be positive");
  return AUR_OK;
}

// Host-buffer search: H2D of the queries, kernels, D2H of the results on a pool context's own stream.
int search_host(aur_index* ix, const void* queries_host, int32_t nq, int32_t k, const int32_t* q_user, const int32_t* q_org,
                const int64_t* allow_ids, int64_t n_allow, float* scores_out, int64_t* ids_out, int64_t* snapshot_out) {
  int rc = check_search_args(ix, queries_host, nq, k, scores_out, ids_out);
  if (rc != AUR_OK) return rc;
  std::shared_lock<std::shared_mutex> rl(ix->rw);
  CU_TRY(cudaSetDevice(ix->device));
  SearchCtx* c = nullptr;
  if ((rc = acquire_ctx(ix, nullptr, &c)) != AUR_OK) return rc;
  struct Guard { aur_index* ix; SearchCtx* c; ~Guard() { release_ctx(ix, c); } } guard{ix, c};
  std::lock_guard<std::mutex> cl(c->mu);
  cudaStream_t s = c->own_stream;
  const int64_t n_rows = ix->rows_pub.load(std::memory_order_acquire);
  const size_t qbytes = static_cast<size_t>(nq) * ix->dim * ix->elt;
  const size_t nout = static_cast<size_t>(nq) * k;
  CU_TRY(c->stage_q.reserve(qbytes));
  CU_TRY(c->stage_scores.reserve(nout));
  CU_TRY(c->stage_ids.reserve(nout));
  CU_TRY(cudaMemcpyAsync(c->stage_q.p, queries_host, qbytes, cudaMemcpyHostToDevice, s));
  Scope sc;
  int32_t scope[2] = {0, -1};
  std::vector<int32_t> rows_host;
  if (allow_ids || n_allow > 0) {
    // resolved metadata pre-filter: ids -> rows of the published prefix (unknown / newer ids drop out)
    if (n_allow < 0 || (n_allow > 0 && !allow_ids)) return fail(AUR_ERR_INVALID, "allow_ids / n_allow");
    rows_host.reserve(static_cast<size_t>(n_allow));
    {
      std::lock_guard<std::mutex> lk(ix->mu);
      for (int64_t i = 0; i < n_allow; ++i) {
        auto it = ix->id2row.find(allow_ids[i]);
        if (it != ix->id2row.end() && it->second < n_rows) rows_host.push_back(static_cast<int32_t>(it->second));
      }
    }
    CU_TRY(c->allow_rows.reserve(rows_host.size() + 1));
    if (!rows_host.empty())
      CU_TRY(cudaMemcpyAsync(c->allow_rows.p, rows_host.data(), rows_host.size() * 4, cudaMemcpyHostToDevice, s));
    sc.allow_rows = c->allow_rows.p;
    sc.n_allow = static_cast<int64_t>(rows_host.size());
  } else if (q_user) {
    CU_TRY(c->stage_quser.reserve(nq));
    CU_TRY(cudaMemcpyAsync(c->stage_quser.p, q_user, static_cast<size_t>(nq) * 4, cudaMemcpyHostToDevice, s));
    sc.q_user = c->stage_quser.p;
    if (q_org) {
      CU_TRY(c->stage_qorg.reserve(nq));
      CU_TRY(cudaMemcpyAsync(c->stage_qorg.p, q_org, static_cast<size_t>(nq) * 4, cudaMemcpyHostToDevice, s));
      sc.q_org = c->stage_qorg.p;
    }
    // the reference asks one tenant's question at a time (weaviate_client.py:244-249): when every query of the
    // batch carries the same (user, org) scope the filter folds into the row scale and the tcgen05 kernel serves it
    bool uniform = true;
    scope[0] = q_user[0]; scope[1] = q_org ? q_org[0] : -1;
    for (int i = 1; i < nq && uniform; ++i) uniform = q_user[i] == scope[0] && (q_org ? q_org[i] : -1) == scope[1];
    if (uniform) sc.uniform = scope;
    else {
      // a coalesced batch of several tenants' questions: with at most 32 distinct scopes every corpus row gets a bit
      // mask (one pre-pass) and the tensor-core kernel serves the batch; more scopes -> the generic kernel
      std::vector<int32_t> tab; std::vector<int32_t> qs(static_cast<size_t>(nq));
      bool fits = true;
      for (int i = 0; i < nq && fits; ++i) {
        const int32_t u = q_user[i], o = q_org ? q_org[i] : -1;
        int found = -1;
        for (size_t t = 0; t < tab.size() / 2; ++t) if (tab[2 * t] == u && tab[2 * t + 1] == o) { found = static_cast<int>(t); break; }
        if (found < 0) {
          if (tab.size() / 2 == 32) { fits = false; break; }
          found = static_cast<int>(tab.size() / 2); tab.push_back(u); tab.push_back(o);
        }
        qs[static_cast<size_t>(i)] = found;
      }
      if (fits) {
        CU_TRY(c->scope_tab.reserve(64)); CU_TRY(c->q_scope.reserve(static_cast<size_t>(nq)));
        CU_TRY(cudaMemcpyAsync(c->scope_tab.p, tab.data(), tab.size() * 4, cudaMemcpyHostToDevice, s));
        CU_TRY(cudaMemcpyAsync(c->q_scope.p, qs.data(), static_cast<size_t>(nq) * 4, cudaMemcpyHostToDevice, s));
        CU_TRY(cudaStreamSynchronize(s));          // tab / qs are host temporaries
        sc.scope_tab = c->scope_tab.p; sc.q_scope = c->q_scope.p; sc.n_scopes = static_cast<int>(tab.size() / 2);
      }
    }
  }
  // Pinned caller buffers are mapped into the device's address space (UVA): the re-rank kernel then writes its nq x k
  // results straight into them over PCIe and the two device-to-host copies (a launch + ~8 us of latency each) disappear.
  float* d_scores = c->stage_scores.p;
  int64_t* d_ids = c->stage_ids.p;
  bool direct = false;
  static const bool no_direct = getenv("AUR_NO_DIRECT_OUT") != nullptr;     // A/B switch
  if (!no_direct) {
    cudaPointerAttributes as{}, ai{};
    if (cudaPointerGetAttributes(&as, scores_out) == cudaSuccess && cudaPointerGetAttributes(&ai, ids_out) == cudaSuccess &&
        as.type == cudaMemoryTypeHost && ai.type == cudaMemoryTypeHost && as.devicePointer && ai.devicePointer) {
      d_scores = static_cast<float*>(as.devicePointer);
      d_ids = static_cast<int64_t*>(ai.devicePointer);
      direct = true;
    } else {
      cudaGetLastError();
    }
  }
  rc = search_enqueue(ix, c, c->stage_q.p, nq, k, sc, n_rows, d_scores, d_ids, nullptr, s);
  if (rc != AUR_OK) { cudaStreamSynchronize(s); return rc; }
  if (!direct) {
    // into the caller's pageable buffers; nothing is written unless every kernel above was enqueued successfully
    CU_TRY(cudaMemcpyAsync(scores_out, c->stage_scores.p, nout * 4, cudaMemcpyDeviceToHost, s));
    CU_TRY(cudaMemcpyAsync(ids_out, c->stage_ids.p, nout * 8, cudaMemcpyDeviceToHost, s));
  }
  CU_TRY(cudaStreamSynchronize(s));
  if (snapshot_out) *snapshot_out = n_rows;
  return AUR_OK;
}

}  // namespace

extern "C" {

int aur_abi_version(void) { return AUR_ABI_VERSION; }
const char* aur_last_error(void) { return g_err.c_str(); }

int aur_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
  return n;
}

int aur_open(const aur_config* cfg, aur_index** out) {
  if (!cfg || !out) return fail(AUR_ERR_INVALID, "null argument");
  *out = nullptr;
  if (cfg->dim <= 0 || cfg->capacity <= 0) return fail(AUR_ERR_INVALID, "dim and capacity must be positive");
  if (cfg->dtype != AUR_BF16 && cfg->dtype != AUR_F32) return fail(AUR_ERR_INVALID, "dtype must be AUR_BF16 or AUR_F32");
  if (cfg->dtype == AUR_BF16 && cfg->dim % 8 != 0) return fail(AUR_ERR_INVALID, "bf16 rows need dim %% 8 == 0");
  if (cfg->capacity > 0x7FFFFFC0ll) return fail(AUR_ERR_INVALID, "capacity exceeds int32 row indexing");
  int ndev = aur_device_count();
  if (ndev == 0) return fail(AUR_ERR_NO_DEVICE, "no CUDA device: aurora_b200 has no CPU fallback");
  if (cfg->device < 0 || cfg->device >= ndev) return fail(AUR_ERR_INVALID, "device %d out of range", cfg->device);
  CU_TRY(cudaSetDevice(cfg->device));
  cudaDeviceProp prop;
  CU_TRY(cudaGetDeviceProperties(&prop, cfg->device));
  if (prop.major < 10) return fail(AUR_ERR_UNSUPPORTED, "sm_%d%d device: this library is built for sm_100a only", prop.major, prop.minor);
  aur_index* ix = new aur_index();
  ix->device = cfg->device; ix->dim = cfg->dim; ix->dtype = cfg->dtype; ix->capacity = cfg->capacity;
  ix->elt = cfg->dtype == AUR_BF16 ? 2 : 4;
  ix->sm_count = prop.multiProcessorCount;
  ix->smem_optin = prop.sharedMemPerBlockOptin;
  auto bail = [&](int rc) { aur_close(ix); return rc; };
#define OPEN_TRY(expr)                                                                                  \
  do { cudaError_t e_ = (expr); if (e_ != cudaSuccess) return bail(fail(e_ == cudaErrorMemoryAllocation ? AUR_ERR_NOMEM : AUR_ERR_CUDA, \
                                                                 "%s: %s", #expr, cudaGetErrorString(e_))); } while (0)
  int prio_lo = 0, prio_hi = 0;
  OPEN_TRY(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
  OPEN_TRY(cudaStreamCreateWithPriority(&ix->stream, cudaStreamNonBlocking, prio_hi));
  OPEN_TRY(cudaStreamCreateWithPriority(&ix->ingest_stream, cudaStreamNonBlocking, prio_lo));
  // round the row store up to a whole tile so TMA boxes never straddle the allocation
  const int64_t cap_pad = (cfg->capacity + kTcTileN - 1) / kTcTileN * kTcTileN;
  OPEN_TRY(cudaMalloc(&ix->d_rows, static_cast<size_t>(cap_pad) * ix->dim * ix->elt));
  OPEN_TRY(cudaMalloc(&ix->d_inv_norm, static_cast<size_t>(cap_pad) * 4));
  OPEN_TRY(cudaMalloc(&ix->d_ids, static_cast<size_t>(cap_pad) * 8));
  OPEN_TRY(cudaMalloc(&ix->d_user, static_cast<size_t>(cap_pad) * 4));
  OPEN_TRY(cudaMalloc(&ix->d_org, static_cast<size_t>(cap_pad) * 4));
#undef OPEN_TRY
  int rc = build_tmaps(ix);
  if (rc != AUR_OK) return bail(rc);
  *out = ix;
  return AUR_OK;
}

int aur_close(aur_index* ix) {
  if (!ix) return AUR_OK;
  cudaSetDevice(ix->device);
  cudaDeviceSynchronize();   // searches bound to caller streams may still be in flight
  cudaFree(ix->d_rows); cudaFree(ix->d_inv_norm); cudaFree(ix->d_ids); cudaFree(ix->d_user); cudaFree(ix->d_org);
  for (auto& c : ix->ctxs) c->release();
  if (ix->stream) cudaStreamDestroy(ix->stream);
  if (ix->ingest_stream) cudaStreamDestroy(ix->ingest_stream);
  delete ix;
  return AUR_OK;
}

int aur_get_stats(aur_index* ix, aur_stats* out) {
  if (!ix || !out) return fail(AUR_ERR_INVALID, "null argument");
  memset(out, 0, sizeof *out);
  SearchCtx* c = nullptr;
  {
    std::lock_guard<std::mutex> lk(ix->mu);
    out->rows = ix->rows_pub.load(std::memory_order_acquire); out->live = ix->live; out->capacity = ix->capacity;
    out->dim = ix->dim; out->dtype = ix->dtype;
    c = ix->last_ctx;
  }
  if (c) {
    std::lock_guard<std::mutex> cl(c->mu);
    out->last_kernel = c->last_kernel; out->last_launches = c->last_launches;
    if (c->have_timing) {
      CU_TRY(cudaSetDevice(ix->device));
      CU_TRY(cudaEventSynchronize(c->ev_end));
      CU_TRY(cudaEventElapsedTime(&out->last_kernel_ms, c->ev_k0, c->ev_k1));
      CU_TRY(cudaEventElapsedTime(&out->last_total_ms, c->ev_begin, c->ev_end));
      CU_TRY(cudaEventElapsedTime(&out->last_finalize_ms, c->ev_k1, c->ev_fin));
      CU_TRY(cudaEventElapsedTime(&out->last_merge_ms, c->ev_fin, c->ev_end));
    }
  }
  return AUR_OK;
}

int aur_export(aur_index* ix, void* rows_out, int64_t* ids_out, int32_t* user_out, int32_t* org_out, uint8_t* live_out,
               int64_t n) {
  if (!ix || !rows_out || !ids_out || !live_out) return fail(AUR_ERR_INVALID, "null argument");
  std::unique_lock<std::shared_mutex> wl(ix->rw);
  const int64_t rows = ix->rows_pub.load(std::memory_order_acquire);
  if (n != rows) return fail(AUR_ERR_INVALID, "n must equal aur_stats.rows (%lld)", (long long)rows);
  if (n == 0) return AUR_OK;
  CU_TRY(cudaSetDevice(ix->device));
  CU_TRY(cudaDeviceSynchronize());
  std::vector<float> inv(static_cast<size_t>(n));
  CU_TRY(cudaMemcpy(inv.data(), ix->d_inv_norm, sizeof(float) * n, cudaMemcpyDeviceToHost));
  CU_TRY(cudaMemcpy(rows_out, ix->d_rows, static_cast<size_t>(n) * ix->dim * ix->elt, cudaMemcpyDeviceToHost));
  CU_TRY(cudaMemcpy(ids_out, ix->d_ids, sizeof(int64_t) * n, cudaMemcpyDeviceToHost));
  if (user_out) CU_TRY(cudaMemcpy(user_out, ix->d_user, sizeof(int32_t) * n, cudaMemcpyDeviceToHost));
  if (org_out) CU_TRY(cudaMemcpy(org_out, ix->d_org, sizeof(int32_t) * n, cudaMemcpyDeviceToHost));
  for (int64_t i = 0; i < n; ++i) live_out[i] = inv[static_cast<size_t>(i)] == inv[static_cast<size_t>(i)];   // NaN = tombstone
  return AUR_OK;
}

int aur_read_rows(aur_index* ix, int64_t row0, int64_t n, void* rows_out, int64_t* ids_out) {
  if (!ix || (n > 0 && (!rows_out || !ids_out))) return fail(AUR_ERR_INVALID, "null argument");
  std::shared_lock<std::shared_mutex> rl(ix->rw);
  const int64_t rows = ix->rows_pub.load(std::memory_order_acquire);
  if (row0 < 0 || n < 0 || row0 + n > rows) return fail(AUR_ERR_INVALID, "rows [%lld, %lld) are not inside the published prefix of %lld",
                                                        (long long)row0, (long long)(row0 + n), (long long)rows);
  if (n == 0) return AUR_OK;
  CU_TRY(cudaSetDevice(ix->device));
  const size_t rb = static_cast<size_t>(ix->dim) * ix->elt;
  CU_TRY(cudaMemcpy(rows_out, static_cast<const uint8_t*>(ix->d_rows) + static_cast<size_t>(row0) * rb, static_cast<size_t>(n) * rb,
                    cudaMemcpyDeviceToHost));
  CU_TRY(cudaMemcpy(ids_out, ix->d_ids + row0, static_cast<size_t>(n) * 8, cudaMemcpyDeviceToHost));
  return AUR_OK;
}

int aur_compact(aur_index* ix, int64_t* reclaimed) {
  if (!ix) return fail(AUR_ERR_INVALID, "null index");
  if (reclaimed) *reclaimed = 0;
  std::unique_lock<std::shared_mutex> wl(ix->rw);   // no search or append is being enqueued ...
  std::lock_guard<std::mutex> wk(ix->mu_write);
  CU_TRY(cudaSetDevice(ix->device));
  CU_TRY(cudaDeviceSynchronize());                  // ... and none is still running
  const int64_t rows = ix->rows_pub.load(std::memory_order_acquire);
  std::vector<std::pair<int64_t, int64_t>> order;   // (old row, id) of the live rows, in row order
  {
    std::lock_guard<std::mutex> lk(ix->mu);
    order.reserve(ix->id2row.size());
    for (auto& kv : ix->id2row) order.emplace_back(kv.second, kv.first);
  }
  std::sort(order.begin(), order.end());
  const int64_t nlive = static_cast<int64_t>(order.size());
  if (nlive == rows) return AUR_OK;                 // nothing to reclaim
  // Stable compaction in place, a bounce buffer at a time: live row j moves down to row j.  Chunk [i0, i1) only
  // overwrites rows < i1 <= old(i1), i.e. rows whose content has already been gathered or moved.
  const int64_t chunk = 65536;
  DevBuf<int32_t> d_map; DevBuf<uint8_t> bounce; DevBuf<float> b_inv; DevBuf<int64_t> b_ids; DevBuf<int32_t> b_user, b_org;
  cudaError_t e = d_map.reserve(static_cast<size_t>(chunk));
  if (e == cudaSuccess) e = bounce.reserve(static_cast<size_t>(chunk) * ix->dim * ix->elt);
  if (e == cudaSuccess) e = b_inv.reserve(chunk);
  if (e == cudaSuccess) e = b_ids.reserve(chunk);
  if (e == cudaSuccess) e = b_user.reserve(chunk);
  if (e == cudaSuccess) e = b_org.reserve(chunk);
  auto drop = [&]() { d_map.release(); bounce.release(); b_inv.release(); b_ids.release(); b_user.release(); b_org.release(); };
  if (e != cudaSuccess) { drop(); return fail(AUR_ERR_NOMEM, "aur_compact: %s", cudaGetErrorString(e)); }
  cudaStream_t s = ix->ingest_stream;
  std::vector<int32_t> map_host(static_cast<size_t>(chunk));
  int64_t first_moved = 0;
  while (first_moved < nlive && order[static_cast<size_t>(first_moved)].first == first_moved) ++first_moved;   // untouched prefix
  for (int64_t i0 = first_moved; i0 < nlive && e == cudaSuccess; i0 += chunk) {
    const int64_t m = (nlive - i0 < chunk) ? nlive - i0 : chunk;
    for (int64_t j = 0; j < m; ++j) map_host[static_cast<size_t>(j)] = static_cast<int32_t>(order[static_cast<size_t>(i0 + j)].first);
    e = cudaMemcpyAsync(d_map.p, map_host.data(), static_cast<size_t>(m) * 4, cudaMemcpyHostToDevice, s);
    if (e == cudaSuccess) e = launch_gather_rows(ix->d_rows, ix->d_inv_norm, ix->d_ids, ix->d_user, ix->d_org, d_map.p, m,
                                                 static_cast<int>(ix->dim * ix->elt), bounce.p, b_inv.p, b_ids.p, b_user.p, b_org.p, s);
    const size_t rb = static_cast<size_t>(ix->dim) * ix->elt;
    if (e == cudaSuccess) e = cudaMemcpyAsync(static_cast<uint8_t*>(ix->d_rows) + static_cast<size_t>(i0) * rb, bounce.p, static_cast<size_t>(m) * rb, cudaMemcpyDeviceToDevice, s);
    if (e == cudaSuccess) e = cudaMemcpyAsync(ix->d_inv_norm + i0, b_inv.p, static_cast<size_t>(m) * 4, cudaMemcpyDeviceToDevice, s);
    if (e == cudaSuccess) e = cudaMemcpyAsync(ix->d_ids + i0, b_ids.p, static_cast<size_t>(m) * 8, cudaMemcpyDeviceToDevice, s);
    if (e == cudaSuccess) e = cudaMemcpyAsync(ix->d_user + i0, b_user.p, static_cast<size_t>(m) * 4, cudaMemcpyDeviceToDevice, s);
    if (e == cudaSuccess) e = cudaMemcpyAsync(ix->d_org + i0, b_org.p, static_cast<size_t>(m) * 4, cudaMemcpyDeviceToDevice, s);
    if (e == cudaSuccess) e = cudaStreamSynchronize(s);   // map_host is reused by the next chunk
  }
  drop();
  if (e != cudaSuccess) return fail(AUR_ERR_CUDA, "aur_compact: %s (the shard may be inconsistent: restore the last snapshot)", cudaGetErrorString(e));
  {
    std::lock_guard<std::mutex> lk(ix->mu);
    for (int64_t j = 0; j < nlive; ++j) ix->id2row[order[static_cast<size_t>(j)].second] = j;
    ix->live = nlive;
  }
  ix->rows_pub.store(nlive, std::memory_order_release);
  if (reclaimed) *reclaimed = rows - nlive;
  return AUR_OK;
}

int aur_set_option(aur_index* ix, const char* key, int64_t value) {
  if (!ix || !key) return fail(AUR_ERR_INVALID, "null argument");
  std::lock_guard<std::mutex> lk(ix->mu);
  if (strcmp(key, "kernel") == 0) {
    if (value < AUR_KERNEL_AUTO || value > AUR_KERNEL_TC2) return fail(AUR_ERR_INVALID, "unknown kernel %lld", (long long)value);
    ix->opt_kernel = static_cast<int>(value);
    return AUR_OK;
  }
  if (strcmp(key, "epi_groups") == 0) {
    if (value < 0 || value > 2) return fail(AUR_ERR_INVALID, "epi_groups must be 0 (auto), 1 or 2");
    ix->opt_epi_groups = static_cast<int>(value);
    return AUR_OK;
  }
  if (strcmp(key, "dbg_flags") == 0) { ix->opt_dbg_flags = static_cast<int>(value); return AUR_OK; }
  return fail(AUR_ERR_INVALID, "unknown option '%s'", key);
}

int aur_sync(aur_index* ix) {
  if (!ix) return fail(AUR_ERR_INVALID, "null argument");
  CU_TRY(cudaSetDevice(ix->device));
  CU_TRY(cudaStreamSynchronize(ix->stream));
  return AUR_OK;
}

int aur_add(aur_index* ix, const void* rows_host, const int64_t* ids, const int32_t* user_codes,
            const int32_t* org_codes, int64_t n) {
  if (!ix) return fail(AUR_ERR_INVALID, "null index");
  std::shared_lock<std::shared_mutex> rl(ix->rw);
  std::lock_guard<std::mutex> wk(ix->mu_write);
  CU_TRY(cudaSetDevice(ix->device));
  return add_common(ix, rows_host, false, ids, user_codes, org_codes, n, ix->ingest_stream);
}

int aur_add_dev(aur_index* ix, const void* rows_dev, const int64_t* ids_host, const int32_t* user_codes_host,
                const int32_t* org_codes_host, int64_t n, void* stream) {
  if (!ix) return fail(AUR_ERR_INVALID, "null index");
  std::shared_lock<std::shared_mutex> rl(ix->rw);
  std::lock_guard<std::mutex> wk(ix->mu_write);
  CU_TRY(cudaSetDevice(ix->device));
  return add_common(ix, rows_dev, true, ids_host, user_codes_host, org_codes_host, n,
                    stream ? static_cast<cudaStream_t>(stream) : ix->ingest_stream);
}

int aur_remove(aur_index* ix, const int64_t* ids, int64_t n, int64_t* removed) {
  if (!ix || (n > 0 && !ids)) return fail(AUR_ERR_INVALID, "null argument");
  std::shared_lock<std::shared_mutex> rl(ix->rw);
  std::lock_guard<std::mutex> wk(ix->mu_write);
  CU_TRY(cudaSetDevice(ix->device));
  const float nanv = nanf("");
  std::vector<int64_t> dead;
  {
    std::lock_guard<std::mutex> lk(ix->mu);
    for (int64_t i = 0; i < n; ++i) {
      auto it = ix->id2row.find(ids[i]);
      if (it == ix->id2row.end()) continue;
      dead.push_back(it->second);
      ix->id2row.erase(it);
      --ix->live;
    }
  }
  for (int64_t row : dead) CU_TRY(cudaMemcpyAsync(ix->d_inv_norm + row, &nanv, 4, cudaMemcpyHostToDevice, ix->ingest_stream));
  CU_TRY(cudaStreamSynchronize(ix->ingest_stream));
  if (removed) *removed = static_cast<int64_t>(dead.size());
  return AUR_OK;
}

int aur_search_dev(aur_index* ix, const void* queries_dev, int32_t nq, int32_t k, const int32_t* q_user_dev,
                   const int32_t* q_org_dev, float* scores_dev, int64_t* ids_dev, double* scores64_dev, void* stream) {
  int rc = check_search_args(ix, queries_dev, nq, k, scores_dev, ids_dev);
  if (rc != AUR_OK) return rc;
  std::shared_lock<std::shared_mutex> rl(ix->rw);
  CU_TRY(cudaSetDevice(ix->device));
  cudaStream_t s = stream ? static_cast<cudaStream_t>(stream) : ix->stream;
  SearchCtx* c = nullptr;
  if ((rc = acquire_ctx(ix, s, &c)) != AUR_OK) return rc;
  std::lock_guard<std::mutex> cl(c->mu);
  Scope sc;
  sc.q_user = q_user_dev; sc.q_org = q_org_dev;
  return search_enqueue(ix, c, queries_dev, nq, k, sc, ix->rows_pub.load(std::memory_order_acquire), scores_dev, ids_dev,
                        scores64_dev, s);
}

int aur_search(aur_index* ix, const void* queries_host, int32_t nq, int32_t k, const int32_t* q_user,
               const int32_t* q_org, float* scores_out, int64_t* ids_out) {
  return search_host(ix, queries_host, nq, k, q_user, q_org, nullptr, -1, scores_out, ids_out, nullptr);
}

int aur_search_ex(aur_index* ix, const void* queries_host, int32_t nq, int32_t k, const int32_t* q_user,
                  const int32_t* q_org, float* scores_out, int64_t* ids_out, int64_t* snapshot_rows_out) {
  return search_host(ix, queries_host, nq, k, q_user, q_org, nullptr, -1, scores_out, ids_out, snapshot_rows_out);
}

int aur_search_subset(aur_index* ix, const void* queries_host, int32_t nq, int32_t k, const int64_t* allow_ids,
                      int64_t n_allow, float* scores_out, int64_t* ids_out) {
  if (n_allow < 0) return fail(AUR_ERR_INVALID, "n_allow < 0");
  static const int64_t none = 0;
  return search_host(ix, queries_host, nq, k, nullptr, nullptr, allow_ids ? allow_ids : &none, n_allow, scores_out, ids_out, nullptr);
}


// ---------------------------------------------------------------------------------------------------------------
// Fused cross-shard exchange (row-sharded corpus, one process per GPU).  Each rank owns one buffer in its HBM that
// every other rank maps through CUDA IPC; a shard's exact top-k rows are stored into all buffers by the finalize
// kernel itself and a flag per (parity, source rank) says when a slot is complete.  No NCCL call, no host
// round trip: local search -> peer stores -> merge is three kernels on one stream.
struct aur_exchange {
  int device = 0, rank = 0, world = 1, nq_max = 0, k_max = 0;
  size_t entries = 0, slot_stride = 0, parity_stride = 0, bytes = 0;   // slot_stride / parity_stride in 8-byte words
  uint64_t* local = nullptr;           // this rank's buffer (cudaMalloc, exported through cudaIpc)
  uint64_t* peer[8] = {};              // every rank's buffer as mapped here (peer[rank] == local)
  uint64_t* d_seq = nullptr;           // exchanges completed
  uint32_t* d_done = nullptr;          // merge kernel's block counter, then the status word
  bool connected = false;
};

int aur_merge_topk_dev(int32_t device, const double* in_scores64, const int64_t* in_ids, int32_t n_shards, int32_t nq,
                       int32_t k, float* out_scores, int64_t* out_ids, double* out_scores64, void* stream) {
  if (!in_scores64 || !in_ids || !out_scores || !out_ids) return fail(AUR_ERR_INVALID, "null argument");
  if (n_shards <= 0 || nq <= 0 || k <= 0 || k > kMaxK || n_shards * k > 2048) return fail(AUR_ERR_INVALID, "bad merge shape");
  if (aur_device_count() == 0) return fail(AUR_ERR_NO_DEVICE, "no CUDA device");
  CU_TRY(cudaSetDevice(device));
  CU_TRY(launch_merge_topk(in_scores64, in_ids, static_cast<size_t>(nq) * k, n_shards, nq, k, out_scores, out_ids,
                           out_scores64, static_cast<cudaStream_t>(stream)));
  return AUR_OK;
}

int aur_merge_topk_packed_dev(int32_t device, const void* packed, int32_t n_shards, int32_t nq, int32_t k,
                              float* out_scores, int64_t* out_ids, double* out_scores64, void* stream) {
  if (!packed || !out_scores || !out_ids) return fail(AUR_ERR_INVALID, "null argument");
  if (n_shards <= 0 || nq <= 0 || k <= 0 || k > kMaxK || n_shards * k > 2048) return fail(AUR_ERR_INVALID, "bad merge shape");
  if (aur_device_count() == 0) return fail(AUR_ERR_NO_DEVICE, "no CUDA device");
  CU_TRY(cudaSetDevice(device));
  const size_t plane = static_cast<size_t>(nq) * k;
  const double* s64 = static_cast<const double*>(packed);
  const int64_t* ids = static_cast<const int64_t*>(packed) + plane;
  CU_TRY(launch_merge_topk(s64, ids, 2 * plane, n_shards, nq, k, out_scores, out_ids, out_scores64,
                           static_cast<cudaStream_t>(stream)));
  return AUR_OK;
}

int aur_exchange_create(int32_t device, int32_t rank, int32_t world, int32_t nq_max, int32_t k_max, aur_exchange** out,
                        uint8_t* handle_out /* [64] */) {
  if (!out || !handle_out) return fail(AUR_ERR_INVALID, "null argument");
  *out = nullptr;
  if (world < 1 || world > 8 || rank < 0 || rank >= world) return fail(AUR_ERR_INVALID, "1 <= world <= 8, 0 <= rank < world");
  if (nq_max <= 0 || k_max <= 0 || k_max > kMaxK || world * k_max > 2048) return fail(AUR_ERR_INVALID, "bad nq_max / k_max");
  if (aur_device_count() == 0) return fail(AUR_ERR_NO_DEVICE, "no CUDA device");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "handle size");
  CU_TRY(cudaSetDevice(device));
  aur_exchange* ex = new aur_exchange();
  ex->device = device; ex->rank = rank; ex->world = world; ex->nq_max = nq_max; ex->k_max = k_max;
  ex->entries = static_cast<size_t>(nq_max) * k_max;
  ex->slot_stride = 4 * ex->entries;                       // an entry = 4 tagged words (score lo / hi, id lo / hi)
  ex->parity_stride = static_cast<size_t>(world) * ex->slot_stride;
  ex->bytes = 2 * ex->parity_stride * 8 + 256;
  cudaError_t e = cudaMalloc(&ex->local, ex->bytes);
  if (e == cudaSuccess) e = cudaMemset(ex->local, 0, ex->bytes);
  if (e == cudaSuccess) e = cudaMalloc(&ex->d_seq, 8);
  if (e == cudaSuccess) e = cudaMemset(ex->d_seq, 0, 8);
  if (e == cudaSuccess) e = cudaMalloc(&ex->d_done, 16);      // [0] merge block counter, [1] status, [2] finalize block counter
  if (e == cudaSuccess) e = cudaMemset(ex->d_done, 0, 16);
  if (e == cudaSuccess) e = cudaDeviceSynchronize();
  cudaIpcMemHandle_t h;
  if (e == cudaSuccess && world > 1) e = cudaIpcGetMemHandle(&h, ex->local);
  if (e != cudaSuccess) {
    cudaFree(ex->local); cudaFree(ex->d_seq); cudaFree(ex->d_done); delete ex;
    return fail(AUR_ERR_CUDA, "aur_exchange_create: %s", cudaGetErrorString(e));
  }
  if (world > 1) memcpy(handle_out, &h, 64); else memset(handle_out, 0, 64);
  ex->peer[rank] = ex->local;
  ex->connected = world == 1;
  *out = ex;
  return AUR_OK;
}

int aur_exchange_connect(aur_exchange* ex, const uint8_t* all_handles /* [world][64], rank order */) {
  if (!ex || !all_handles) return fail(AUR_ERR_INVALID, "null argument");
  CU_TRY(cudaSetDevice(ex->device));
  for (int r = 0; r < ex->world; ++r) {
    if (r == ex->rank || ex->peer[r]) continue;
    cudaIpcMemHandle_t h;
    memcpy(&h, all_handles + static_cast<size_t>(r) * 64, 64);
    void* p = nullptr;
    cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) return fail(AUR_ERR_CUDA, "cudaIpcOpenMemHandle(rank %d): %s", r, cudaGetErrorString(e));
    ex->peer[r] = static_cast<uint64_t*>(p);
  }
  ex->connected = true;
  return AUR_OK;
}

int aur_exchange_close(aur_exchange* ex) {
  if (!ex) return AUR_OK;
  cudaSetDevice(ex->device);
  cudaDeviceSynchronize();
  for (int r = 0; r < ex->world; ++r)
    if (r != ex->rank && ex->peer[r]) cudaIpcCloseMemHandle(ex->peer[r]);
  cudaFree(ex->local); cudaFree(ex->d_seq); cudaFree(ex->d_done);
  delete ex;
  return AUR_OK;
}

int aur_exchange_status(aur_exchange* ex, int64_t* exchanges_done, int32_t* status) {
  if (!ex) return fail(AUR_ERR_INVALID, "null argument");
  CU_TRY(cudaSetDevice(ex->device));
  uint64_t seq = 0; uint32_t st[2] = {0, 0};
  CU_TRY(cudaMemcpy(&seq, ex->d_seq, 8, cudaMemcpyDeviceToHost));
  CU_TRY(cudaMemcpy(st, ex->d_done, 8, cudaMemcpyDeviceToHost));
  if (exchanges_done) *exchanges_done = static_cast<int64_t>(seq);
  if (status) *status = static_cast<int32_t>(st[1]);
  return AUR_OK;
}

int aur_search_exchange_dev(aur_index* ix, aur_exchange* ex, const void* queries_dev, int32_t nq, int32_t k,
                            float* scores_dev, int64_t* ids_dev, void* stream) {
  int rc = check_search_args(ix, queries_dev, nq, k, scores_dev, ids_dev);
  if (rc != AUR_OK) return rc;
  if (!ex || !ex->connected) return fail(AUR_ERR_INVALID, "exchange not connected");
  if (ex->device != ix->device) return fail(AUR_ERR_INVALID, "exchange and index live on different devices");
  if (nq > ex->nq_max || k > ex->k_max || static_cast<size_t>(nq) * k > ex->entries)
    return fail(AUR_ERR_INVALID, "batch %d x top-%d exceeds the exchange's %d x %d", nq, k, ex->nq_max, ex->k_max);
  std::shared_lock<std::shared_mutex> rl(ix->rw);
  CU_TRY(cudaSetDevice(ix->device));
  cudaStream_t s = stream ? static_cast<cudaStream_t>(stream) : ix->stream;
  SearchCtx* c = nullptr;
  if ((rc = acquire_ctx(ix, s, &c)) != AUR_OK) return rc;
  std::lock_guard<std::mutex> cl(c->mu);
  FinalizeArgs::ExchangeOut eo{};
  eo.n_peers = ex->world;
  for (int r = 0; r < ex->world; ++r) eo.slot[r] = ex->peer[r] + static_cast<size_t>(ex->rank) * ex->slot_stride;
  eo.seq = ex->d_seq;
  eo.parity_stride = ex->parity_stride;
  Scope sc;
  rc = search_enqueue(ix, c, queries_dev, nq, k, sc, ix->rows_pub.load(std::memory_order_acquire), nullptr, nullptr, nullptr, s, &eo);
  if (rc != AUR_OK) return rc;
  ExchangeParams p{};
  p.slots = ex->local;
  p.seq = ex->d_seq; p.done = ex->d_done; p.status = ex->d_done + 1;
  p.world = ex->world; p.rank = ex->rank; p.nq = nq; p.k = k;
  p.parity_stride = ex->parity_stride; p.slot_stride = ex->slot_stride;
  p.out_scores = scores_dev; p.out_ids = ids_dev;
  CU_TRY(launch_exchange_merge(p, s));
  c->last_launches += 1;
  CU_TRY(cudaEventRecord(c->ev_end, s));
  return AUR_OK;
}

int aur_cosine_pairs(int32_t device, const float* a_host, const float* b_host, int64_t n, int32_t dim, int32_t clamp,
                     double* out_host) {
  if (n < 0 || dim < 0) return fail(AUR_ERR_INVALID, "negative size");
  if (n == 0) return AUR_OK;
  if (!a_host || !b_host || !out_host) return fail(AUR_ERR_INVALID, "null argument");
  if (aur_device_count() == 0) return fail(AUR_ERR_NO_DEVICE, "no CUDA device: aurora_b200 has no CPU fallback");
  CU_TRY(cudaSetDevice(device));
  if (dim == 0) { for (int64_t i = 0; i < n; ++i) out_host[i] = 0.0; return AUR_OK; }
  float *da = nullptr, *db = nullptr; double* dout = nullptr;
  const size_t bytes = static_cast<size_t>(n) * dim * 4;
  cudaError_t e = cudaMalloc(&da, bytes);
  if (e == cudaSuccess) e = cudaMalloc(&db, bytes);
  if (e == cudaSuccess) e = cudaMalloc(&dout, static_cast<size_t>(n) * 8);
  if (e == cudaSuccess) e = cudaMemcpy(da, a_host, bytes, cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMemcpy(db, b_host, bytes, cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = launch_cosine_pairs(da, db, n, dim, clamp, dout, nullptr);
  std::vector<double> tmp(static_cast<size_t>(n));
  if (e == cudaSuccess) e = cudaMemcpy(tmp.data(), dout, static_cast<size_t>(n) * 8, cudaMemcpyDeviceToHost);
  cudaFree(da); cudaFree(db); cudaFree(dout);
  if (e != cudaSuccess) return fail(AUR_ERR_CUDA, "cosine_pairs: %s", cudaGetErrorString(e));
  memcpy(out_host, tmp.data(), static_cast<size_t>(n) * 8);
  return AUR_OK;
}

int aur_dev_malloc(int32_t device, uint64_t bytes, void** out) {
  if (!out) return fail(AUR_ERR_INVALID, "null argument");
  *out = nullptr;
  if (aur_device_count() == 0) return fail(AUR_ERR_NO_DEVICE, "no CUDA device");
  CU_TRY(cudaSetDevice(device));
  void* p = nullptr;
  cudaError_t e = cudaMalloc(&p, bytes ? bytes : 1);
  if (e != cudaSuccess) return fail(AUR_ERR_NOMEM, "cudaMalloc(%llu): %s", (unsigned long long)bytes, cudaGetErrorString(e));
  *out = p;
  return AUR_OK;
}
int aur_dev_free(int32_t device, void* p) {
  if (!p) return AUR_OK;
  CU_TRY(cudaSetDevice(device));
  CU_TRY(cudaFree(p));
  return AUR_OK;
}
int aur_memcpy_h2d(int32_t device, void* dst_dev, const void* src_host, uint64_t bytes) {
  if (bytes == 0) return AUR_OK;
  if (!dst_dev || !src_host) return fail(AUR_ERR_INVALID, "null argument");
  CU_TRY(cudaSetDevice(device));
  CU_TRY(cudaMemcpy(dst_dev, src_host, bytes, cudaMemcpyHostToDevice));
  return AUR_OK;
}
int aur_memcpy_d2h(int32_t device, void* dst_host, const void* src_dev, uint64_t bytes) {
  if (bytes == 0) return AUR_OK;
  if (!dst_host || !src_dev) return fail(AUR_ERR_INVALID, "null argument");
  CU_TRY(cudaSetDevice(device));
  CU_TRY(cudaDeviceSynchronize());
  CU_TRY(cudaMemcpy(dst_host, src_dev, bytes, cudaMemcpyDeviceToHost));
  return AUR_OK;
}

int aur_debug_tc_scores(aur_index* ix, const void* queries_dev, int32_t nq, int32_t cta_group, float* out_dev,
                        int32_t* n_ctas_out, void* stream) {
  if (!ix || !queries_dev || !out_dev) return fail(AUR_ERR_INVALID, "null argument");
  std::shared_lock<std::shared_mutex> rl(ix->rw);
  CU_TRY(cudaSetDevice(ix->device));
  if (!ix->tmap_ok) return fail(AUR_ERR_UNSUPPORTED, "index shape has no tcgen05 path");
  if (nq <= 0 || nq > 2 * kTcQRows) return fail(AUR_ERR_INVALID, "1..256 queries");
  int n_lists = 0;
  cudaStream_t s = stream ? static_cast<cudaStream_t>(stream) : ix->stream;
  SearchCtx* c = nullptr;
  int rc = acquire_ctx(ix, s, &c);
  if (rc != AUR_OK) return rc;
  std::lock_guard<std::mutex> cl(c->mu);
  rc = run_tc_block(ix, c, cta_group, queries_dev, nq, 32 + kSlack, ix->rows_pub.load(std::memory_order_acquire), out_dev,
                    &n_lists, s);
  if (rc != AUR_OK) return rc;
  CU_TRY(cudaMemsetAsync(c->cand_count.p, 0, 8 * kTcQRows * 4, s));  // no finalize ran to reset them
  if (n_ctas_out) *n_ctas_out = ix->sm_count & ~1;
  return AUR_OK;
}

}  // extern "C"
