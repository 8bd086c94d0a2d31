// Host runtime + C ABI (include/aurora_b200.h).  Owns one corpus shard in HBM:
//   rows      [capacity, dim]  bf16 or f32, row-major (K-major for the MMA, 128-bit
//                              coalesced for everything else)
//   inv_norm  [capacity] f32   1/|row|; 0 for zero rows; NaN marks a tombstone
//   ids       [capacity] i64   caller ids;  user / org [capacity] i32 tenant codes
// plus grow-only scratch for candidate lists.  No CPU compute path exists here: without
// a CUDA device every entry point fails with AUR_ERR_NO_DEVICE.
#include <cuda.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <string.h>

#include <mutex>
#include <string>
#include <unordered_map>
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

struct aur_index {
  std::mutex mu;
  int device = 0, dim = 0, dtype = 0;
  int64_t capacity = 0, rows = 0, live = 0;
  size_t elt = 2;
  void* d_rows = nullptr;
  float* d_inv_norm = nullptr;
  int64_t* d_ids = nullptr;
  int32_t* d_user = nullptr;
  int32_t* d_org = nullptr;
  std::unordered_map<int64_t, int64_t> id2row;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev_begin = nullptr, ev_k0 = nullptr, ev_k1 = nullptr, ev_end = nullptr;
  bool have_timing = false;
  CUtensorMap tmap[2];  // box rows 64 (cta_group::1) and 32 (cta_group::2)
  bool tmap_ok = false;
  int sm_count = 0;
  size_t smem_optin = 0;
  int opt_kernel = AUR_KERNEL_AUTO;
  int opt_dbg_flags = 0;
  int opt_epi_groups = 0;        // 0 = auto
  int last_kernel = 0, last_launches = 0;
  DevBuf<uint64_t> cand_a, cand_b;
  DevBuf<uint64_t> pub;          // tcgen05 kernel's cross-CTA threshold exchange
  DevBuf<uint32_t> cand_count;   // compacted candidates per query (self-resetting)
  uint32_t epoch = 0;
  DevBuf<float> score_chunk;
  DevBuf<float> masked_inv;      // inverse norms with one tenant's invisible rows turned into NaN
  DevBuf<uint8_t> stage_q;       // host-entry staging: queries
  DevBuf<int32_t> stage_quser, stage_qorg;
  DevBuf<float> stage_scores;
  DevBuf<int64_t> stage_ids;
  DevBuf<float> dbg;
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

// Runs one block of <= 256 queries through the tcgen05 kernel.  Leaves candidate keys in
// ix->cand_a as [nqb_pad, n_lists, ksel]; returns n_lists.
int run_tc_block(aur_index* ix, int cta_group, const void* q_dev, int nqb, int ksel, float* dbg, int* n_lists_out,
                 cudaStream_t s, const float* inv_norm = nullptr) {
  int n_qblocks = (nqb > kTcQRows) ? 2 : 1;
  if (cta_group == 2 && n_qblocks != 2) {
    // a pair works on 256 query rows.  A short tail block normally runs as single CTAs; when their larger
    // TMA stages do not fit next to the lists (large k at dim > 768) it runs as a pair with a padding block
    if (tc_pick_stages(1, 1, ksel, ix->dim, ix->smem_optin) >= 2) cta_group = 1; else n_qblocks = 2;
  }
  int grid = ix->sm_count & ~1;
  const int n_tsets = (cta_group == 2) ? grid / 2 : grid / n_qblocks;
  // epilogue groups: 1 by default; 2 (alternating tiles) stays selectable for experiments
  int epi_groups = ix->opt_epi_groups;
  if (epi_groups == 0) epi_groups = 1;   // measured: one group + a deeper TMA ring (11 stages) beats two groups + 8
  const int stages = tc_pick_stages(cta_group, epi_groups, ksel, ix->dim, ix->smem_optin);
  if (stages < 2) return fail(AUR_ERR_UNSUPPORTED, "k too large for the tcgen05 path's shared memory");
  const size_t smem = tc_smem_bytes(cta_group, epi_groups, stages, ksel, ix->dim);
  const int n_lists = n_tsets * epi_groups;   // candidate lists per query
  const size_t ncand = static_cast<size_t>(n_qblocks) * kTcQRows * n_lists * ksel;
  CU_TRY(ix->cand_a.reserve(ncand));
  const size_t npub = static_cast<size_t>(n_qblocks) * kTcQRows * (((n_tsets + 1) & ~1) + 1);
  if (npub > ix->pub.n) {
    CU_TRY(ix->pub.reserve(npub));
    CU_TRY(cudaMemsetAsync(ix->pub.p, 0, npub * 8, s));  // epoch 0 is never used by a launch
  }
  if (ix->cand_count.n < 2 * kTcQRows) {
    CU_TRY(ix->cand_count.reserve(2 * kTcQRows));
    CU_TRY(cudaMemsetAsync(ix->cand_count.p, 0, 2 * kTcQRows * 4, s));
  }
  if (++ix->epoch == 0) ix->epoch = 1;
  TcParams p;
  p.q = static_cast<const __nv_bfloat16*>(q_dev);
  p.inv_norm = inv_norm ? inv_norm : ix->d_inv_norm;
  p.cand = ix->cand_a.p;
  p.cand_count = ix->cand_count.p;
  p.dbg_scores = dbg;
  p.pub = ix->pub.p;
  p.epoch = ix->epoch;
  p.n_rows = ix->rows;
  p.nq = nqb; p.dim = ix->dim; p.ksel = ksel; p.n_lists = n_lists; p.n_qblocks = n_qblocks;
  p.num_stages = stages;
  p.dbg_flags = ix->opt_dbg_flags;
  p.n_tiles = static_cast<int>((ix->rows + kTcTileN - 1) / kTcTileN);
  CU_TRY(tc_launch(cta_group, epi_groups, grid, &ix->tmap[cta_group - 1], p, smem, s));
  *n_lists_out = n_lists;
  return AUR_OK;
}

// uniform_scope (nullable): {user, org} when the host knows every query of the batch carries the same tenant scope;
// the filter then folds into the row scale and the tcgen05 kernel serves the batch.
int search_dev_locked(aur_index* ix, const void* q_dev, int nq, int k, const int32_t* q_user, const int32_t* q_org,
                      float* scores, int64_t* ids, double* scores64, cudaStream_t s, const int32_t* uniform_scope = nullptr) {
  if (nq <= 0 || k <= 0) return fail(AUR_ERR_INVALID, "nq and k must be positive");
  if (k > kMaxK) return fail(AUR_ERR_UNSUPPORTED, "k > %d", kMaxK);
  if (nq > 65535) return fail(AUR_ERR_UNSUPPORTED, "nq > 65535: split the batch");
  const bool filtered = q_user != nullptr && uniform_scope == nullptr;   // per-query scopes: generic kernel only
  const int ksel = k + kSlack;
  int kernel = ix->opt_kernel;
  if (kernel == AUR_KERNEL_AUTO) kernel = tc_shape_ok(ix, k, filtered) ? AUR_KERNEL_TC2 : AUR_KERNEL_SIMT;
  if (kernel != AUR_KERNEL_SIMT && !tc_shape_ok(ix, k, filtered))
    return fail(AUR_ERR_UNSUPPORTED, "tcgen05 path needs bf16, dim %% 64 == 0, dim <= %d, no tenant filter, and k small enough "
                "for its shared-memory lists at this dim", kTcMaxDim);
  ix->last_kernel = kernel;
  ix->last_launches = 0;
  CU_TRY(cudaEventRecord(ix->ev_begin, s));
  bool k_timed = false;
  const float* tc_inv = nullptr;
  if (kernel != AUR_KERNEL_SIMT && q_user != nullptr && ix->rows > 0) {   // one scope for the whole batch
    CU_TRY(ix->masked_inv.reserve(static_cast<size_t>(ix->capacity) + 64));
    CU_TRY(launch_mask_inv_norm(ix->d_inv_norm, ix->d_user, ix->d_org, uniform_scope[0], uniform_scope[1], ix->rows,
                                ix->masked_inv.p, s));
    ++ix->last_launches;
    tc_inv = ix->masked_inv.p;
  }

  const int qstep = (kernel == AUR_KERNEL_SIMT) ? 1024 : 2 * kTcQRows;
  for (int q0 = 0; q0 < nq; q0 += qstep) {
    const int nqb = (nq - q0 < qstep) ? nq - q0 : qstep;
    const uint8_t* qb = static_cast<const uint8_t*>(q_dev) + static_cast<size_t>(q0) * ix->dim * ix->elt;
    int n_lists = 0;
    uint64_t* cur = nullptr;
    if (kernel == AUR_KERNEL_SIMT) {
      if (ix->rows == 0) {
        n_lists = 1;
        CU_TRY(ix->cand_a.reserve(static_cast<size_t>(nqb) * ksel));
        CU_TRY(cudaMemsetAsync(ix->cand_a.p, 0, static_cast<size_t>(nqb) * ksel * 8, s));
      } else {
        n_lists = static_cast<int>((ix->rows + kSimtSeg - 1) / kSimtSeg);
        CU_TRY(ix->cand_a.reserve(static_cast<size_t>(nqb) * n_lists * ksel));
        const int64_t chunk = 16 * kSimtSeg;  // 32768 rows of scores at a time
        CU_TRY(ix->score_chunk.reserve(static_cast<size_t>(nqb) * chunk));
        FilterArgs f{ix->d_user, ix->d_org, q_user ? q_user + q0 : nullptr, q_org ? q_org + q0 : nullptr};
        if (!k_timed) CU_TRY(cudaEventRecord(ix->ev_k0, s));
        for (int64_t r0 = 0; r0 < ix->rows; r0 += chunk) {
          const int64_t nr = (ix->rows - r0 < chunk) ? ix->rows - r0 : chunk;
          CU_TRY(launch_simt_scores(qb, ix->d_rows, ix->dtype, ix->dim, nqb, r0, nr, ix->rows, ix->d_inv_norm, f,
                                    ix->score_chunk.p, s));
          CU_TRY(launch_simt_select(ix->score_chunk.p, nqb, r0, nr, ksel, ix->cand_a.p, n_lists,
                                    static_cast<int>(r0 / kSimtSeg), s));
          ix->last_launches += 2;
        }
        if (!k_timed) { CU_TRY(cudaEventRecord(ix->ev_k1, s)); k_timed = true; }
      }
      cur = ix->cand_a.p;
    } else {
      if (!k_timed) CU_TRY(cudaEventRecord(ix->ev_k0, s));
      int rc = run_tc_block(ix, kernel == AUR_KERNEL_TC1 ? 1 : 2, qb, nqb, ksel, nullptr, &n_lists, s, tc_inv);
      if (rc != AUR_OK) return rc;
      if (!k_timed) { CU_TRY(cudaEventRecord(ix->ev_k1, s)); k_timed = true; }
      ix->last_launches += 1;
      cur = ix->cand_a.p;
    }
    // dense candidate lists (SIMT path): fold until one sort of <= 4096 keys finishes the
    // job.  The tcgen05 kernel already compacted its survivors per query.
    const bool compact = kernel != AUR_KERNEL_SIMT;
    bool in_a = true;
    while (!compact && static_cast<int64_t>(n_lists) * ksel > 4096) {
      const int group = 4096 / ksel;
      const int n_groups = (n_lists + group - 1) / group;
      DevBuf<uint64_t>& dst = in_a ? ix->cand_b : ix->cand_a;
      // rows of cand are indexed by the query position inside the block (TC pads to 128/256)
      CU_TRY(dst.reserve(static_cast<size_t>(nqb) * n_groups * ksel));
      CU_TRY(launch_reduce_lists(cur, nqb, n_lists, ksel, group, dst.p, s));
      ix->last_launches += 1;
      cur = dst.p; n_lists = n_groups; in_a = !in_a;
    }
    FinalizeArgs fa;
    fa.cand = cur; fa.n_lists = n_lists; fa.ksel = ksel;
    fa.counts = compact ? ix->cand_count.p : nullptr;
    fa.q = qb; fa.rows = ix->d_rows; fa.dtype = ix->dtype; fa.dim = ix->dim; fa.nq = nqb; fa.k = k;
    fa.ids = ix->d_ids;
    fa.out_scores = scores + static_cast<size_t>(q0) * k;
    fa.out_ids = ids + static_cast<size_t>(q0) * k;
    fa.out_scores64 = scores64 ? scores64 + static_cast<size_t>(q0) * k : nullptr;
    CU_TRY(launch_finalize(fa, s));
    ix->last_launches += 1;
  }
  if (!k_timed) { CU_TRY(cudaEventRecord(ix->ev_k0, s)); CU_TRY(cudaEventRecord(ix->ev_k1, s)); }
  CU_TRY(cudaEventRecord(ix->ev_end, s));
  ix->have_timing = true;
  return AUR_OK;
}

int add_common(aur_index* ix, const void* rows, bool rows_on_device, const int64_t* ids, const int32_t* users,
               const int32_t* orgs, int64_t n, cudaStream_t s) {
  if (n < 0) return fail(AUR_ERR_INVALID, "n < 0");
  if (n == 0) return AUR_OK;
  if (!rows || !ids) return fail(AUR_ERR_INVALID, "rows and ids are required");
  if (ix->rows + n > ix->capacity)
    return fail(AUR_ERR_NOMEM, "shard full: %lld + %lld > capacity %lld", (long long)ix->rows, (long long)n,
                (long long)ix->capacity);
  for (int64_t i = 0; i < n; ++i)
    if (ids[i] < 0) return fail(AUR_ERR_INVALID, "ids must be >= 0");
  const int64_t base = ix->rows;
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
  // upsert: an id that already exists loses its old row (weaviate_client.py:172 uuid5 semantics)
  std::vector<int64_t> dead;
  for (int64_t i = 0; i < n; ++i) {
    auto it = ix->id2row.find(ids[i]);
    if (it != ix->id2row.end()) { dead.push_back(it->second); it->second = base + i; }
    else { ix->id2row.emplace(ids[i], base + i); ++ix->live; }
  }
  const float nanv = nanf("");
  for (int64_t row : dead) CU_TRY(cudaMemcpyAsync(ix->d_inv_norm + row, &nanv, 4, cudaMemcpyHostToDevice, s));
  ix->rows += n;
  CU_TRY(cudaStreamSynchronize(s));  // host staging vectors go out of scope
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
  OPEN_TRY(cudaStreamCreateWithFlags(&ix->stream, cudaStreamNonBlocking));
  OPEN_TRY(cudaEventCreate(&ix->ev_begin)); OPEN_TRY(cudaEventCreate(&ix->ev_k0));
  OPEN_TRY(cudaEventCreate(&ix->ev_k1));    OPEN_TRY(cudaEventCreate(&ix->ev_end));
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
  if (ix->stream) cudaStreamSynchronize(ix->stream);
  cudaFree(ix->d_rows); cudaFree(ix->d_inv_norm); cudaFree(ix->d_ids); cudaFree(ix->d_user); cudaFree(ix->d_org);
  ix->cand_a.release(); ix->cand_b.release(); ix->pub.release(); ix->cand_count.release(); ix->score_chunk.release(); ix->masked_inv.release(); ix->stage_q.release();
  ix->stage_quser.release(); ix->stage_qorg.release(); ix->stage_scores.release(); ix->stage_ids.release();
  ix->dbg.release();
  if (ix->ev_begin) cudaEventDestroy(ix->ev_begin);
  if (ix->ev_k0) cudaEventDestroy(ix->ev_k0);
  if (ix->ev_k1) cudaEventDestroy(ix->ev_k1);
  if (ix->ev_end) cudaEventDestroy(ix->ev_end);
  if (ix->stream) cudaStreamDestroy(ix->stream);
  delete ix;
  return AUR_OK;
}

int aur_get_stats(aur_index* ix, aur_stats* out) {
  if (!ix || !out) return fail(AUR_ERR_INVALID, "null argument");
  std::lock_guard<std::mutex> lk(ix->mu);
  memset(out, 0, sizeof *out);
  out->rows = ix->rows; out->live = ix->live; out->capacity = ix->capacity;
  out->dim = ix->dim; out->dtype = ix->dtype;
  out->last_kernel = ix->last_kernel; out->last_launches = ix->last_launches;
  if (ix->have_timing) {
    CU_TRY(cudaSetDevice(ix->device));
    CU_TRY(cudaEventSynchronize(ix->ev_end));
    CU_TRY(cudaEventElapsedTime(&out->last_kernel_ms, ix->ev_k0, ix->ev_k1));
    CU_TRY(cudaEventElapsedTime(&out->last_total_ms, ix->ev_begin, ix->ev_end));
  }
  return AUR_OK;
}

int aur_export(aur_index* ix, void* rows_out, int64_t* ids_out, int32_t* user_out, int32_t* org_out, uint8_t* live_out,
               int64_t n) {
  if (!ix || !rows_out || !ids_out || !live_out) return fail(AUR_ERR_INVALID, "null argument");
  std::lock_guard<std::mutex> lk(ix->mu);
  if (n != ix->rows) return fail(AUR_ERR_INVALID, "n must equal aur_stats.rows (%lld)", (long long)ix->rows);
  if (n == 0) return AUR_OK;
  CU_TRY(cudaSetDevice(ix->device));
  CU_TRY(cudaStreamSynchronize(ix->stream));
  std::vector<float> inv(static_cast<size_t>(n));
  CU_TRY(cudaMemcpy(inv.data(), ix->d_inv_norm, sizeof(float) * n, cudaMemcpyDeviceToHost));
  CU_TRY(cudaMemcpy(rows_out, ix->d_rows, static_cast<size_t>(n) * ix->dim * ix->elt, cudaMemcpyDeviceToHost));
  CU_TRY(cudaMemcpy(ids_out, ix->d_ids, sizeof(int64_t) * n, cudaMemcpyDeviceToHost));
  if (user_out) CU_TRY(cudaMemcpy(user_out, ix->d_user, sizeof(int32_t) * n, cudaMemcpyDeviceToHost));
  if (org_out) CU_TRY(cudaMemcpy(org_out, ix->d_org, sizeof(int32_t) * n, cudaMemcpyDeviceToHost));
  for (int64_t i = 0; i < n; ++i) live_out[i] = inv[static_cast<size_t>(i)] == inv[static_cast<size_t>(i)];   // NaN = tombstone
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
  std::lock_guard<std::mutex> lk(ix->mu);
  CU_TRY(cudaSetDevice(ix->device));
  return add_common(ix, rows_host, false, ids, user_codes, org_codes, n, ix->stream);
}

int aur_add_dev(aur_index* ix, const void* rows_dev, const int64_t* ids_host, const int32_t* user_codes_host,
                const int32_t* org_codes_host, int64_t n, void* stream) {
  if (!ix) return fail(AUR_ERR_INVALID, "null index");
  std::lock_guard<std::mutex> lk(ix->mu);
  CU_TRY(cudaSetDevice(ix->device));
  return add_common(ix, rows_dev, true, ids_host, user_codes_host, org_codes_host, n,
                    stream ? static_cast<cudaStream_t>(stream) : ix->stream);
}

int aur_remove(aur_index* ix, const int64_t* ids, int64_t n, int64_t* removed) {
  if (!ix || (n > 0 && !ids)) return fail(AUR_ERR_INVALID, "null argument");
  std::lock_guard<std::mutex> lk(ix->mu);
  CU_TRY(cudaSetDevice(ix->device));
  const float nanv = nanf("");
  int64_t cnt = 0;
  for (int64_t i = 0; i < n; ++i) {
    auto it = ix->id2row.find(ids[i]);
    if (it == ix->id2row.end()) continue;
    CU_TRY(cudaMemcpyAsync(ix->d_inv_norm + it->second, &nanv, 4, cudaMemcpyHostToDevice, ix->stream));
    ix->id2row.erase(it);
    --ix->live; ++cnt;
  }
  CU_TRY(cudaStreamSynchronize(ix->stream));
  if (removed) *removed = cnt;
  return AUR_OK;
}

int aur_search_dev(aur_index* ix, const void* queries_dev, int32_t nq, int32_t k, const int32_t* q_user_dev,
                   const int32_t* q_org_dev, float* scores_dev, int64_t* ids_dev, double* scores64_dev, void* stream) {
  if (!ix || !queries_dev || !scores_dev || !ids_dev) return fail(AUR_ERR_INVALID, "null argument");
  std::lock_guard<std::mutex> lk(ix->mu);
  CU_TRY(cudaSetDevice(ix->device));
  return search_dev_locked(ix, queries_dev, nq, k, q_user_dev, q_org_dev, scores_dev, ids_dev, scores64_dev,
                           stream ? static_cast<cudaStream_t>(stream) : ix->stream);
}

int aur_search(aur_index* ix, const void* queries_host, int32_t nq, int32_t k, const int32_t* q_user,
               const int32_t* q_org, float* scores_out, int64_t* ids_out) {
  if (!ix || !queries_host || !scores_out || !ids_out) return fail(AUR_ERR_INVALID, "null argument");
  if (nq <= 0 || k <= 0) return fail(AUR_ERR_INVALID, "nq and k must be positive");
  std::lock_guard<std::mutex> lk(ix->mu);
  CU_TRY(cudaSetDevice(ix->device));
  cudaStream_t s = ix->stream;
  const size_t qbytes = static_cast<size_t>(nq) * ix->dim * ix->elt;
  const size_t nout = static_cast<size_t>(nq) * k;
  CU_TRY(ix->stage_q.reserve(qbytes));
  CU_TRY(ix->stage_scores.reserve(nout));
  CU_TRY(ix->stage_ids.reserve(nout));
  CU_TRY(cudaMemcpyAsync(ix->stage_q.p, queries_host, qbytes, cudaMemcpyHostToDevice, s));
  const int32_t* du = nullptr; const int32_t* dorg = nullptr;
  if (q_user) {
    CU_TRY(ix->stage_quser.reserve(nq));
    CU_TRY(cudaMemcpyAsync(ix->stage_quser.p, q_user, static_cast<size_t>(nq) * 4, cudaMemcpyHostToDevice, s));
    du = ix->stage_quser.p;
    if (q_org) {
      CU_TRY(ix->stage_qorg.reserve(nq));
      CU_TRY(cudaMemcpyAsync(ix->stage_qorg.p, q_org, static_cast<size_t>(nq) * 4, cudaMemcpyHostToDevice, s));
      dorg = ix->stage_qorg.p;
    }
  }
  // the reference asks one tenant's question at a time (weaviate_client.py:244-249): when every query of the
  // batch carries the same (user, org) scope the filter folds into the row scale and the tcgen05 kernel serves it
  int32_t scope[2] = {0, -1};
  bool uniform = q_user != nullptr;
  if (uniform) {
    scope[0] = q_user[0]; scope[1] = q_org ? q_org[0] : -1;
    for (int i = 1; i < nq && uniform; ++i) uniform = q_user[i] == scope[0] && (q_org ? q_org[i] : -1) == scope[1];
  }
  int rc = search_dev_locked(ix, ix->stage_q.p, nq, k, du, dorg, ix->stage_scores.p, ix->stage_ids.p, nullptr, s,
                             uniform ? scope : nullptr);
  if (rc != AUR_OK) return rc;
  // straight into the caller's buffers (async when they are pinned); nothing is written
  // unless every kernel above was enqueued successfully
  CU_TRY(cudaMemcpyAsync(scores_out, ix->stage_scores.p, nout * 4, cudaMemcpyDeviceToHost, s));
  CU_TRY(cudaMemcpyAsync(ids_out, ix->stage_ids.p, nout * 8, cudaMemcpyDeviceToHost, s));
  CU_TRY(cudaStreamSynchronize(s));
  return AUR_OK;
}

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
  std::lock_guard<std::mutex> lk(ix->mu);
  CU_TRY(cudaSetDevice(ix->device));
  if (!ix->tmap_ok) return fail(AUR_ERR_UNSUPPORTED, "index shape has no tcgen05 path");
  if (nq <= 0 || nq > 2 * kTcQRows) return fail(AUR_ERR_INVALID, "1..256 queries");
  int n_lists = 0;
  cudaStream_t s = stream ? static_cast<cudaStream_t>(stream) : ix->stream;
  int rc = run_tc_block(ix, cta_group, queries_dev, nq, 32 + kSlack, out_dev, &n_lists, s);
  if (rc != AUR_OK) return rc;
  CU_TRY(cudaMemsetAsync(ix->cand_count.p, 0, 2 * kTcQRows * 4, s));  // no finalize ran to reset them
  if (n_ctas_out) *n_ctas_out = ix->sm_count & ~1;
  return AUR_OK;
}

}  // extern "C"
