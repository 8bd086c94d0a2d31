// Encoder self-attention on tcgen05, version 2: TWO work items in flight per CTA, 64 keys at a time with an online
// softmax and every hand-off double-buffered -- the tensor core works ahead of and behind the exponentials.
//
// Why: version 1 (attn_tc.cu) keeps the whole score row of ONE (sequence, 128-query block, head) item in the 512 TMEM
// columns and walks a strictly serial chain per item (load -> S MMAs -> row max -> exponentials -> PV MMAs -> store);
// ncu: tensor pipe 15 %, MUFU 37 %, half of the warp samples waiting on a barrier hand-off.  The work per 128 x 128
// block is ~1k cycles of MUFU (16 Ki exponentials at 16 / clk / SM) against ~0.5k cycles of MMA: overlapped they cost
// ~1k cycles, chained ~2k+.
//
// Layout per CTA (1 CTA / SM, persistent):
//   two SLOTS, each with its own stream of work items (slot s of CTA b takes units s*G + b, + 2G, ...):
//     TMEM   S0 S1 [64 cols each] fp32 scores of two consecutive 64-key sub-blocks | P0 P1 [32 cols each] bf16 pairs
//            (TS-MMA A operand) | O [64 cols] fp32 output accumulator      -> 256 columns per slot, 512 per CTA
//     smem   Q ring 2 x 16 KB + K ring 2 x 16 KB + V ring 2 x 16 KB (128-row tiles)  -> 96 KB per slot
//   warps 0-1   TMA producers, one per slot (Q once per item, K_j / V_j of 128 keys through the rings)
//   warps 2-3   MMA issuers, one per slot:  S_t = Q K_t^T (SS, M128 N64 K16 x4),  O (+)= P_t V_t (TS, N64 K16 x4, V read
//               MN-major) per 64-key sub-block t; S_{t+1} and S_{t+2} are issued while the softmax warps still work on
//               S_t, PV_t runs under the exponentials of t+1: nobody waits for the MMA it just asked for
//   warps 4-7   softmax of slot 0, warps 8-11 of slot 1 (thread = query row, the 64 columns of a sub-block in registers)
// Online softmax with a lazy reference maximum: exponentials are taken against m_ref; O and the running sum are
// rescaled only when a sub-block's maximum exceeds m_ref by more than 8 (log2 units, P <= 256 -- harmless in bf16 x
// fp32); the result is exact up to rounding either way.  Keys past the end of a sequence are masked to -inf (P = 0),
// warps whose 32 query rows are all padding only keep the barrier protocol going.
// Every barrier wait is bounded (~1 s): a protocol error aborts the kernel with a message instead of hanging the GPU.
// Restates eager_attention_forward + softmax of transformers' modeling_bert.py:115-140 (oracle/bert_encoder.py).
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "internal.h"
#include "ptx.cuh"

namespace aur {
namespace {

using namespace ptx;

constexpr int kQB = 128, kKB = 128, kSub = 64, kDh = 64;
constexpr int kTile = 128 * kDh * 2;          // 16 KB: 128 rows x 128 B
constexpr int kSlots = 2, kRing = 2;
constexpr int kThreads = (2 * kSlots + 4 * kSlots) * 32;   // per slot: producer, MMA issuer, four softmax warps
constexpr int kMaxUnits = 256;
constexpr int kStageRow = 144;                                     // 128 B of one output row + 16 B pad (bank spread)
constexpr int kStageWarp = 16 * kStageRow;                          // half a warp's rows at a time
constexpr int kSlotSmem = (3 * kRing) * kTile;                     // Q ring + K ring + V ring (Q of the next item is prefetched)
constexpr int kColS = 0, kColP = 128, kColO = 192, kSlotCols = 256;   // S0 S1 | P0 P1 | O
constexpr size_t kSmemBytes = 1024 + static_cast<size_t>(kSlots) * kSlotSmem + 512 + kMaxUnits * 16 + 4 * kSlots * kStageWarp;
constexpr float kRescaleGap = 8.0f;

// per slot: q_full[2], q_free[2], k_full[2], k_free[2], v_full[2], v_free[2], s_full[2], s_free[2], p_full[2], pv_done[2], o_free
enum { B_QFULL = 0, B_QFREE = B_QFULL + kRing, B_KFULL = B_QFREE + kRing, B_KFREE = B_KFULL + kRing, B_VFULL = B_KFREE + kRing, B_VFREE = B_VFULL + kRing,
       B_SFULL = B_VFREE + kRing, B_SFREE = B_SFULL + 2, B_PFULL = B_SFREE + 2, B_PVDONE = B_PFULL + 2, B_OFREE = B_PVDONE + 2, B_COUNT };

__device__ __forceinline__ void tmem_ld_x32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
      "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_st_x32(uint32_t taddr, const uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,"
      "%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
      "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]),
      "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]),
      "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31]) : "memory");
}
__device__ __forceinline__ float ex2_approx(float x) {
  float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y;
}
__device__ __forceinline__ float max3(float a, float b, float c) {
  float d; asm("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c)); return d;
}
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  uint32_t d;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi), "f"(lo));
  return d;
}
__host__ __device__ constexpr uint32_t idesc_bf16_f32_bmn(int m, int n) { return idesc_bf16_f32(m, n) | (1u << 16); }

// Bounded wait: false (and the CTA-wide abort flag set) after 1 s.
// The probe carries a suspend-time hint: a waiting warp is parked by the hardware (no issue slots taken from the
// softmax warps sharing its scheduler) and woken when the phase completes.
__device__ __forceinline__ bool mbar_test_park(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t"
      "}" : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity), "r"(20000u) : "memory");
  return ok != 0;
}
__device__ __forceinline__ bool wait_b(uint64_t* bar, uint32_t parity, volatile int* abort_flag) {
  uint64_t t0 = 0;
  for (uint32_t spin = 0;; ++spin) {
    if (mbar_test_park(bar, parity)) return true;
    if ((spin & 15u) == 15u) {
      if (*abort_flag) return false;
      uint64_t now;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
      if (t0 == 0) t0 = now;
      else if (now - t0 > 1000000000ull) { *abort_flag = 1; return false; }
    }
  }
}

// One work unit = (sequence, 128-query block, head).  CTA b owns units b, b + G, b + 2G, ...; local index li -> slot
// li & 1.  The first kMaxUnits of a CTA's list are decoded once into shared memory (the item table sits in global
// memory: a dependent ~700-cycle load per item transition and role otherwise); longer lists fall back to global loads.
struct Unit { int tok0, len, q0, head; };
__device__ __forceinline__ int unit_nkb(const Unit& u) { return (u.len + kKB - 1) / kKB; }     // 128-key tiles to load
__device__ __forceinline__ int unit_nsb(const Unit& u) { return (u.len + kSub - 1) / kSub; }   // 64-key sub-blocks to compute
__device__ __forceinline__ Unit fetch_unit_global(const AttnParams& p, int w) {
  const AttnItem it = p.items[w / p.heads];
  Unit u; u.tok0 = it.tok0; u.len = it.len; u.q0 = it.q0; u.head = w % p.heads;
  return u;
}
__device__ __forceinline__ Unit get_unit(const AttnParams& p, const int4* tab, int li, int G) {
  if (li < kMaxUnits) { const int4 v = tab[li]; Unit u; u.tok0 = v.x; u.len = v.y; u.q0 = v.z; u.head = v.w; return u; }
  return fetch_unit_global(p, li * G + static_cast<int>(blockIdx.x));
}

// Epilogue of one work item for one softmax warp: wait for the item's last PV, read O, hand the accumulator back, scale by
// 1 / row sum and store.  Kept out of line: it runs once per item from two call sites and would otherwise triple the
// softmax loop's footprint in the instruction cache.
struct Pending { bool have, dead; uint32_t last; float sum; int tok, n_rows, head; };
__device__ __noinline__ bool drain_item(const Pending& d, uint64_t* b, uint32_t trow, uint8_t* stage, int quarter, int lane,
                                        const AttnParams& p, volatile int* abort_flag) {
  if (!wait_b(&b[B_PVDONE + (d.last & 1)], (d.last >> 1) & 1, abort_flag)) return false;   // O complete: in-order pipe
  tc_fence_after();
  if (d.dead) {
    if (lane == 0) mbar_arrive(&b[B_OFREE]);
    return true;
  }
  const float inv = 1.0f / d.sum;
  uint32_t o[2][32];
  tmem_ld_x32(trow + kColO, o[0]);
  tmem_ld_x32(trow + kColO + 32, o[1]);
  tmem_wait_ld();
  tc_fence_before();
  __syncwarp();
  if (lane == 0) mbar_arrive(&b[B_OFREE]);                     // the next item's first PV may overwrite O
  // 32 rows x 128 B, staged 16 rows at a time so that every store instruction writes four whole 128-byte rows
  __nv_bfloat16* dst = p.ctx + static_cast<size_t>(d.tok + quarter * 32) * p.ld_ctx + d.head * kDh;
#pragma unroll
  for (int half = 0; half < 2; ++half) {
    if ((lane >> 4) == half) {
      uint4* srow = reinterpret_cast<uint4*>(stage + (lane & 15) * kStageRow);
#pragma unroll
      for (int h = 0; h < 2; ++h)
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          uint4 q4;
          q4.x = pack_bf16x2(__uint_as_float(o[h][8 * e + 0]) * inv, __uint_as_float(o[h][8 * e + 1]) * inv);
          q4.y = pack_bf16x2(__uint_as_float(o[h][8 * e + 2]) * inv, __uint_as_float(o[h][8 * e + 3]) * inv);
          q4.z = pack_bf16x2(__uint_as_float(o[h][8 * e + 4]) * inv, __uint_as_float(o[h][8 * e + 5]) * inv);
          q4.w = pack_bf16x2(__uint_as_float(o[h][8 * e + 6]) * inv, __uint_as_float(o[h][8 * e + 7]) * inv);
          srow[h * 4 + e] = q4;
        }
    }
    __syncwarp();
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int r = i * 4 + (lane >> 3), rg = half * 16 + r;      // row inside this warp's 32
      const uint4 q4 = *reinterpret_cast<const uint4*>(stage + r * kStageRow + (lane & 7) * 16);
      if (quarter * 32 + rg < d.n_rows)
        *reinterpret_cast<uint4*>(reinterpret_cast<uint8_t*>(dst + static_cast<size_t>(rg) * p.ld_ctx) + (lane & 7) * 16) = q4;
    }
    __syncwarp();
  }
  return true;
}

__global__ void __launch_bounds__(kThreads, 1)
attn_tc2_kernel(const __grid_constant__ CUtensorMap tmap_qkv, const AttnParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (base & 1023u)) & 1023u);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kSlots * kSlotSmem);   // [kSlots][B_COUNT]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + kSlots * B_COUNT);
  volatile int* abort_flag = reinterpret_cast<volatile int*>(tmem_slot + 1);
  int4* unit_tab = reinterpret_cast<int4*>(smem + kSlots * kSlotSmem + 512);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int total = p.n_items * p.heads, G = gridDim.x;
  const int n_local = (total - static_cast<int>(blockIdx.x) + G - 1) / G;      // units of this CTA (>= 1: grid <= total)

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmap_qkv);
    for (int s = 0; s < kSlots; ++s) {
      uint64_t* b = bars + s * B_COUNT;
      for (int i = 0; i < B_COUNT; ++i) mbar_init(&b[i], ((i >= B_SFREE && i < B_PVDONE) || i == B_OFREE) ? 4 : 1);
    }
    *abort_flag = 0;
    fence_mbar_init();
  }
  if (warp == 2) { tmem_alloc<1>(tmem_slot, 512); tmem_relinquish<1>(); }
  for (int li = threadIdx.x; li < n_local && li < kMaxUnits; li += kThreads) {   // the item table is an input, not produced upstream
    const Unit u = fetch_unit_global(p, li * G + static_cast<int>(blockIdx.x));
    unit_tab[li] = make_int4(u.tok0, u.len, u.q0, u.head);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  grid_dep_launch();
  grid_dep_wait();     // the qkv projections come from the previous kernel

  if (warp < kSlots) {
    // ================================================================ TMA producer of slot `warp`
    if (lane == 0) {
      const int s = warp;
      uint64_t* b = bars + s * B_COUNT;
      uint8_t* sm = smem + s * kSlotSmem;
      uint32_t items = 0, blk = 0;      // running item / 128-key tile counts (ring index and barrier parities)
      bool ok = true;
      for (int li = s; li < n_local && ok; li += kSlots, ++items) {
        const Unit u = get_unit(p, unit_tab, li, G);
        const uint32_t rq = items & 1;
        if (items >= kRing) ok = wait_b(&b[B_QFREE + rq], ((items >> 1) - 1) & 1, abort_flag);      // S MMAs of item - 2 retired
        if (!ok) break;
        mbar_arrive_expect_tx(&b[B_QFULL + rq], kTile);
        tma_load_2d(sm + rq * kTile, &tmap_qkv, &b[B_QFULL + rq], u.head * kDh, u.tok0 + u.q0, kEvictNormal);
        const int nkb = unit_nkb(u);
        for (int j = 0; j < nkb && ok; ++j, ++blk) {
          const uint32_t r = blk & 1;
          if (blk >= kRing) ok = wait_b(&b[B_KFREE + r], ((blk >> 1) - 1) & 1, abort_flag);
          if (!ok) break;
          mbar_arrive_expect_tx(&b[B_KFULL + r], kTile);
          tma_load_2d(sm + (kRing + r) * kTile, &tmap_qkv, &b[B_KFULL + r], p.hidden + u.head * kDh, u.tok0 + j * kKB, kEvictNormal);
          if (blk >= kRing) ok = wait_b(&b[B_VFREE + r], ((blk >> 1) - 1) & 1, abort_flag);
          if (!ok) break;
          mbar_arrive_expect_tx(&b[B_VFULL + r], kTile);
          tma_load_2d(sm + (2 * kRing + r) * kTile, &tmap_qkv, &b[B_VFULL + r], 2 * p.hidden + u.head * kDh, u.tok0 + j * kKB, kEvictNormal);
        }
      }
    }
  } else if (warp < 2 * kSlots) {
    // ================================================================ MMA issuer of slot `warp - 2`
    constexpr uint32_t idesc_s = idesc_bf16_f32(kQB, kSub);
    constexpr uint32_t idesc_o = idesc_bf16_f32_bmn(kQB, kDh);
    if (lane == 0) {
      const int s = warp - kSlots;
      uint64_t* b = bars + s * B_COUNT;
      const uint32_t sm = smem_u32(smem + s * kSlotSmem);
      const uint32_t tb = tmem_base + s * kSlotCols;
      // Two program counters over the slot's stream of 64-key sub-blocks: S (item li_s, sub-block ts of nsb_s, running
      // count sb) runs up to three ahead of PV (li_p, tp of nsb_p, pb).  The issue order is static -- S_0 S_1 S_2 PV_0
      // S_3 PV_1 ...: score buffer g & 1 is released early in the softmax of g, P_g arrives at its end -- so every wait
      // below is a plain blocking one.  kb_s / kb_p count the 128-key tiles taken from the K / V rings.
      int li_s = s, li_p = s, ts = 0, tp = 0;
      uint32_t sb = 0, pb = 0, kb_s = 0, kb_p = 0, s_items = 0, p_items = 0;
      bool have_s = li_s < n_local, have_p = have_s, ok = true;
      int nsb_s = have_s ? unit_nsb(get_unit(p, unit_tab, li_s, G)) : 0, nsb_p = nsb_s;
      while (have_p && ok) {
        if (have_s && sb < pb + 3) {
          // ---- S_t = Q K_t^T : needs the K tile (and Q for t == 0) in smem, and score buffer sb & 1 read by the softmax warps
          const uint32_t r = kb_s & 1, sbuf = sb & 1, half = ts & 1, rq = s_items & 1;
          if (half == 0) ok = wait_b(&b[B_KFULL + r], (kb_s >> 1) & 1, abort_flag);
          if (ok && ts == 0) ok = wait_b(&b[B_QFULL + rq], (s_items >> 1) & 1, abort_flag);
          if (ok && sb >= 2) ok = wait_b(&b[B_SFREE + sbuf], ((sb >> 1) - 1) & 1, abort_flag);
          if (!ok) break;
          tc_fence_after();
          const uint64_t q_desc = smem_desc_sw128(sm + rq * kTile);
          const uint64_t k_desc = smem_desc_sw128(sm + (kRing + r) * kTile + half * (kSub * 128));   // key rows 64 .. 127
#pragma unroll
          for (int ks = 0; ks < kDh / 16; ++ks) mma_ss_bf16<1>(tb + kColS + sbuf * kSub, q_desc + 2 * ks, k_desc + 2 * ks, idesc_s, ks != 0);
          mma_commit<1>(&b[B_SFULL + sbuf]);
          ++sb;
          const bool last = ++ts == nsb_s;
          if (half == 1 || last) { mma_commit<1>(&b[B_KFREE + r]); ++kb_s; }
          if (last) {
            mma_commit<1>(&b[B_QFREE + rq]);
            ++s_items; ts = 0; li_s += kSlots; have_s = li_s < n_local;
            if (have_s) nsb_s = unit_nsb(get_unit(p, unit_tab, li_s, G));
          }
          continue;
        }
        // ---- O (+)= P_t V_t : needs P_t written, the V tile in smem, and (t == 0) the previous item's O read out
        const uint32_t r = kb_p & 1, pbuf = pb & 1, half = tp & 1;
        ok = wait_b(&b[B_PFULL + pbuf], (pb >> 1) & 1, abort_flag);
        if (ok && half == 0) ok = wait_b(&b[B_VFULL + r], (kb_p >> 1) & 1, abort_flag);
        if (ok && tp == 0 && p_items > 0) ok = wait_b(&b[B_OFREE], (p_items - 1) & 1, abort_flag);
        if (!ok) break;
        tc_fence_after();
        const uint64_t v_desc = smem_desc_sw128(sm + (2 * kRing + r) * kTile + half * (kSub * 128));
#pragma unroll
        for (int ks = 0; ks < kSub / 16; ++ks)   // 16 keys = two 8-row swizzle groups = 2048 B of V
          mma_ts_bf16<1>(tb + kColO, tb + kColP + pbuf * (kSub / 2) + ks * 8, v_desc + (2048 >> 4) * ks, idesc_o, (tp | ks) != 0);
        mma_commit<1>(&b[B_PVDONE + pbuf]);
        ++pb;
        const bool last = ++tp == nsb_p;
        if (half == 1 || last) { mma_commit<1>(&b[B_VFREE + r]); ++kb_p; }
        if (last) {
          ++p_items; tp = 0; li_p += kSlots; have_p = li_p < n_local;
          if (have_p) nsb_p = unit_nsb(get_unit(p, unit_tab, li_p, G));
        }
      }
    }
  } else {
    // ================================================================ softmax warps (4 per slot)
    const int s = (warp - 2 * kSlots) >> 2, quarter = warp & 3;
    uint64_t* b = bars + s * B_COUNT;
    const uint32_t trow = tmem_base + s * kSlotCols + (static_cast<uint32_t>(quarter * 32) << 16);
    const float sc = p.scale_log2e;
    uint8_t* stage = smem + kSlots * kSlotSmem + 512 + kMaxUnits * 16 + (warp - 2 * kSlots) * kStageWarp;
    uint32_t sbk = 0;                 // running sub-block count of this slot (buffer index and barrier parities)
    // The epilogue of an item (wait for its last PV, read O, normalise, store) runs AFTER the first sub-block of the next
    // item: that sub-block's scores are in TMEM long before, so the warp computes instead of idling on the tensor pipe.
    Pending pend;
    pend.have = false; pend.dead = false; pend.last = 0; pend.sum = 1.f; pend.tok = 0; pend.n_rows = 0; pend.head = 0;
    bool ok = true;
    for (int li = s; li < n_local && ok && !*abort_flag; li += kSlots) {
      const Unit u = get_unit(p, unit_tab, li, G);
      const int nsb = unit_nsb(u);
      float m_ref = -INFINITY, sum = 0.f;
      const bool dead = u.q0 + quarter * 32 >= u.len;              // all 32 query rows of this warp are padding
      for (int t = 0; t < nsb && ok; ++t, ++sbk) {
        const uint32_t buf = sbk & 1, use = sbk >> 1;
        ok = wait_b(&b[B_SFULL + buf], use & 1, abort_flag);
        if (!ok) break;
        if (dead) {
          // keep the barrier protocol going, compute nothing: row r of P only feeds row r of O, which is never stored.
          // (the PV of two sub-blocks back must have retired before PFULL of this buffer may complete again)
          if (lane == 0) mbar_arrive(&b[B_SFREE + buf]);
          if (sbk >= 2) { ok = wait_b(&b[B_PVDONE + buf], (use - 1) & 1, abort_flag); if (!ok) break; }
          if (lane == 0) mbar_arrive(&b[B_PFULL + buf]);
          if (t == 0 && pend.have) { ok = drain_item(pend, b, trow, stage, quarter, lane, p, abort_flag); pend.have = false; }
          continue;
        }
        tc_fence_after();
        uint32_t v[2][32];
        tmem_ld_x32(trow + kColS + buf * kSub, v[0]);
        tmem_ld_x32(trow + kColS + buf * kSub + 32, v[1]);
        tmem_wait_ld();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&b[B_SFREE + buf]);            // S_{t+2} may overwrite this score buffer now
        const int nv = u.len - t * kSub;                           // valid keys in this sub-block (>= 1)
        if (nv < kSub) {                                           // a sequence's last sub-block: keys past its end
#pragma unroll
          for (int c = 0; c < 2; ++c)
#pragma unroll
            for (int e = 0; e < 32; ++e)
              if (c * 32 + e >= nv) v[c][e] = 0xFF800000u;         // -inf: P = 0
        }
        float mb = -INFINITY;
#pragma unroll
        for (int c = 0; c < 2; ++c)
#pragma unroll
          for (int e = 0; e < 32; e += 8) {                       // independent chains
            const float m0 = max3(__uint_as_float(v[c][e]), __uint_as_float(v[c][e + 1]), __uint_as_float(v[c][e + 2]));
            const float m1 = max3(__uint_as_float(v[c][e + 3]), __uint_as_float(v[c][e + 4]), __uint_as_float(v[c][e + 5]));
            mb = max3(mb, max3(m0, m1, __uint_as_float(v[c][e + 6])), __uint_as_float(v[c][e + 7]));
          }
        mb *= sc;                                                  // log2 domain (sc > 0)
        // lazy rescale: only when this sub-block's maximum leaves the reference far behind
        const bool grow = mb > m_ref + kRescaleGap;
        if (t > 0 && __any_sync(0xffffffffu, grow)) {
          // O is touched: every PV issued so far for this item must have retired (the one of t - 1 is the last)
          ok = wait_b(&b[B_PVDONE + (buf ^ 1)], ((sbk - 1) >> 1) & 1, abort_flag);
          if (!ok) break;
          tc_fence_after();
          const float alpha = grow ? ex2_approx(m_ref - mb) : 1.0f;     // m_ref = -inf cannot happen for t > 0
          uint32_t o[32];
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            tmem_ld_x32(trow + kColO + h * 32, o);
            tmem_wait_ld();
#pragma unroll
            for (int e = 0; e < 32; ++e) o[e] = __float_as_uint(__uint_as_float(o[e]) * alpha);
            tmem_st_x32(trow + kColO + h * 32, o);
          }
          tmem_wait_st();
          sum *= alpha;
        }
        if (grow) m_ref = mb;
        if (sbk >= 2) {
          // P buffer `buf` still belongs to the PV of two sub-blocks back (retired long ago in the steady state)
          ok = wait_b(&b[B_PVDONE + buf], (use - 1) & 1, abort_flag);
          if (!ok) break;
          tc_fence_after();
        }
        // exponentials against the reference maximum; P as bf16 pairs (two keys per 32-bit column)
        float s0 = 0.f, s1 = 0.f;
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          uint32_t o16[16];
          if (c * 32 >= nv) {                                      // 32 keys past the sequence end: P = 0, no exponentials
#pragma unroll
            for (int e = 0; e < 16; ++e) o16[e] = 0u;
            tmem_st_x16(trow + kColP + buf * (kSub / 2) + c * 16, o16);
            continue;
          }
#pragma unroll
          for (int e = 0; e < 32; e += 2) {
            // (moving a quarter of these to an FMA-pipe polynomial, as MUFU-bound kernels do, measured 4-5 % SLOWER here:
            //  the loop is issue/latency bound at two warps per scheduler, profiles/attn_ab_r2.txt)
            const float p0 = ex2_approx(fmaf(__uint_as_float(v[c][e]), sc, -m_ref));
            const float p1 = ex2_approx(fmaf(__uint_as_float(v[c][e + 1]), sc, -m_ref));
            s0 += p0; s1 += p1;
            o16[e >> 1] = pack_bf16x2(p0, p1);
          }
          tmem_st_x16(trow + kColP + buf * (kSub / 2) + c * 16, o16);
        }
        sum += s0 + s1;
        tmem_wait_st();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&b[B_PFULL + buf]);
        if (t == 0 && pend.have) { ok = drain_item(pend, b, trow, stage, quarter, lane, p, abort_flag); pend.have = false; }
      }
      if (!ok) break;
      pend.have = true; pend.dead = dead; pend.last = sbk - 1; pend.sum = sum;
      pend.tok = u.tok0 + u.q0; pend.n_rows = u.len - u.q0; pend.head = u.head;
    }
    if (ok && pend.have) drain_item(pend, b, trow, stage, quarter, lane, p, abort_flag);
  }

  tc_fence_before();
  __syncthreads();
  if (*abort_flag && threadIdx.x == 0) printf("attn_tc2_kernel: barrier time-out in CTA %d (protocol error) -- output invalid\n", blockIdx.x);
  if (warp == 2) { tc_fence_after(); tmem_dealloc<1>(tmem_base, 512); }
}

}  // namespace

cudaError_t attn_tc2_launch(int sm_count, const void* tmap_qkv, const AttnParams& p, cudaStream_t s) {
  const int total = p.n_items * p.heads;
  if (total <= 0) return cudaSuccess;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(attn_tc2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(kSmemBytes));
    if (e != cudaSuccess) return e;
    attr_set = true;
  }
  int grid = (total + kSlots - 1) / kSlots;
  if (grid > sm_count) grid = sm_count;
  return launch_pdl(attn_tc2_kernel, dim3(grid), dim3(kThreads), kSmemBytes, s, 1, *reinterpret_cast<const CUtensorMap*>(tmap_qkv), p);
}

}  // namespace aur
