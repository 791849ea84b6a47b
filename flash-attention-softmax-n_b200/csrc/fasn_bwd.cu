// Backward kernel: dQ, dK, dV of attention-with-softmax_n for sm_100a.
//
// Replaces the reference's `_bwd_kernel` (flash_attention_softmax_n/core/flash_attn_triton.py:146-235), which runs
// one program per head, serial over KV blocks, with a read-modify-write of dQ in HBM per inner iteration.  Here one
// CTA owns one 128-row K/V tile of one (batch, head) unit and streams the 128-row Q/dO tiles past it, everything
// transposed so that the probabilities land in TMEM as the A operand of the next MMA (no shared-memory round trip):
//
//   S^T  = K Q_i^T            SS-MMA, K-major operands              -> TMEM [0,128)
//   dP^T = V dO_i^T           SS-MMA                                -> TMEM [128,256)
//   P^T  = exp2(S^T c - LSE2_n[q])        compute warps, in registers; dropout mask regenerated from philox
//   dS^T = P^T o (Z/(1-p) dP^T - delta[q])
//   dV  += (P^T o Z/(1-p)) dO_i   TS-MMA, A = P^T in TMEM (16-bit, written over the first half of each thread's own
//                                 S^T columns: [0,32) and [64,96)), B = dO_i MN-major        -> TMEM [256,256+D)
//   dK  += dS^T Q_i               SS-MMA, A = dS^T in smem (K-major),    B = Q_i MN-major   -> TMEM [256+D,256+2D)
//   dQ_i = dS K                   SS-MMA, A = the same smem tile read MN-major, B = K MN-major -> TMEM [128,128+D)
//                                 (aliases dP^T), then added into the fp32 dq_accum by TMA reduce-add (cp.reduce.async.bulk.tensor)
//
// The only difference from softmax_0 attention is that P is recomputed from LSE_n = ln(n + sum exp s) (SURVEY.md
// section 9): the Jacobian keeps the softmax form dS = P o (dP - delta) with delta_i = sum_d O_id dO_id.
//
//   warps 0-7   compute: warp w owns TMEM lanes 32(w%4).. (kv rows) and q-columns 64(w/4)..64(w/4)+63.  Inside the warp
//               the fast kernels use the "quad" register layout of tcgen05.ld.16x256b: a thread holds 4 kv rows x 16
//               query columns (not 1 row x 64 columns), so it needs LSE2 / delta of 16 queries per tile instead of 64
//               and the four threads of a quad fetch them with one 32-byte shared-memory wavefront -- the per-query
//               constants cost 16 wavefronts per warp and tile instead of 128 (they were 950 of the ~5300 shared-memory
//               port cycles per tile pair)
//   warps 8-11  dQ reducers: TMEM -> registers -> fp32 staging in smem -> TMA reduce-add
//   warp 12     TMA producer (K,V once; Q_i + LSE2 + delta through a 2-deep ring, dO_i single-buffered)
//   warp 13     MMA issuer (one elected lane)  warp 14  TMEM allocator        warp 15  issues the dV / dK stores of the epilogue
//
// Persistent: the grid is one CTA per SM; every CTA takes (K/V tile, unit) work items from a global counter in the heavy-first
// order of decode_block until they run out.  Barrier set-up, tensor-memory allocation and the launch gap are paid once per SM,
// and the head of the next item overlaps the tail of the current one: Q / dO of its first tile are loaded as soon as the ring
// slot / the dO buffer is free, K as soon as the last dQ MMA has completed (dK is staged in the dS^T tile, not in sK), V as soon
// as the TMA store of dV has read its staging tile (sV; warp 15 issues the epilogue stores and waits for them, so no compute
// thread does), and S^T / dP^T of the first tile are issued while the compute warps are still draining dK.  Per-CTA overhead was 9760 cycles + a 770 ns launch gap per item against 4718 cycles per Q tile
// (12 % of the C3 backward, scripts/cta_profile.py).
#include "fasn_common.cuh"
#include "fasn_ptx.cuh"

namespace fasn {

namespace {

// Optional phase timeline (compile with -DFASN_TIMELINE): one CTA records clock64() at its pipeline events into
// BwdArgs.dbg, [role][slot] = (tag << 48) | clock.  Used by scripts/timeline.py; compiled out of the product build.
#ifdef FASN_TIMELINE
#define TL_DECL(role) unsigned long long* tl_p = (a.dbg && blockIdx.x == a.dbg_x) ? a.dbg + (role) * 2048 : nullptr; int tl_i = 0;   // one CTA, all of its items
#define TL_ONLY(cond) do { if (!(cond)) tl_p = nullptr; } while (0)
#define TL(tag) do { if (tl_p && tl_i < 2048) tl_p[tl_i++] = ((unsigned long long)(tag) << 48) | (clock64() & 0xFFFFFFFFFFFFull); } while (0)
// Per-CTA record (every CTA, thread 0): [4 b] = globaltimer at entry, [4 b + 1] = at exit, [4 b + 2] = clock64 cycles in
// between, [4 b + 3] = (smid << 32) | iterations, stored behind the five role timelines.  scripts/cta_profile.py fits
// lifetime = overhead + period * iterations and adds up the per-SM busy time.
#define TL_CTA_BEGIN() unsigned long long cta_g0 = 0, cta_c0 = 0; if (a.dbg && threadIdx.x == 0) { asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(cta_g0)); cta_c0 = clock64(); }
#define TL_CTA_END(iters) do { if (a.dbg && threadIdx.x == 0) { unsigned long long g1; unsigned int sm; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g1)); asm volatile("mov.u32 %0, %%smid;" : "=r"(sm)); \
    unsigned long long* rec = a.dbg + 5 * 2048 + 4ull * blockIdx.x; rec[0] = cta_g0; rec[1] = g1; rec[2] = clock64() - cta_c0; rec[3] = ((unsigned long long)sm << 32) | (unsigned int)(iters); } } while (0)
#else
#define TL_DECL(role)
#define TL_ONLY(cond)
#define TL(tag)
#define TL_CTA_BEGIN()
#define TL_CTA_END(iters)
#endif


constexpr int kBwdThreads = 512;
// K-steps (of 8) of dK_i issued BEFORE dP^T of the next Q tile; the rest follow it.  The MMAs execute in issue order, and
// the critical path of an iteration is dS_i -> dQ_i -> (drain of dQ_i from tensor memory) -> dP^T_{i+1} -> dS_{i+1}: with all of
// dK_i queued ahead of dP^T_{i+1} its ~900 cycles sit on that path; only as much of it as fits into the drain belongs there.
#ifndef FASN_BWD_DK_SPLIT
#define FASN_BWD_DK_SPLIT 4
#endif
// Register budgets after the role split (setmaxnreg): 8 compute warps, 4 dQ reducer warps (the whole fp32 dQ tile row of a
// thread, D values, lives in registers), 4 producer / MMA / idle warps.  256 x 152 + 128 x 152 + 128 x 56 = 65536 at D = 128,
// 256 x 176 + 128 x 104 + 128 x 56 = 65536 at D = 64.
template <int D> struct BwdRegs {
  static constexpr int kCompute = (D == 128) ? 152 : 176, kReduce = (D == 128) ? 152 : 104, kOther = 56;
};

template <int D> struct BwdCfg {
  static constexpr int DB = D / 64;
  static constexpr int TILE_BYTES = 128 * D * 2;
  static constexpr int BLK_BYTES = 128 * 128;
  static constexpr int DS_BYTES = 2 * BLK_BYTES;                  // dS^T: [2 q-blocks][128 kv rows][128 B]
  static constexpr int NUM_BARS = 27;
  static constexpr int DQ_STAGE_BYTES = 128 * 32 * 4;             // dQ staging chunk: 128 rows x 32 fp32 columns
  // K, V, Q ring (2), dO (1), dS^T, dQ staging (2 chunks), LSE2 ring + delta ring (2 x 2 x 512 B), barriers, tmem slot
  static constexpr int SMEM_BYTES = 5 * TILE_BYTES + DS_BYTES + 2 * DQ_STAGE_BYTES + 4 * 512 + NUM_BARS * 8 + 16;
  static_assert(SMEM_BYTES <= 232448, "shared memory per CTA");
};

// One work item = one 128-row K/V tile of one (batch, head) unit and the Q / dO tiles that can see it.
struct BwdItem { int kt, k0, bh, b, h, hk, i_start, n_iter; };

template <bool CAUSAL> FASN_DEVICE BwdItem bwd_decode(const BwdArgs& a, int lin, int nkv, int nq) {
  BwdItem w;
  const TileCoord tcd = decode_block((uint32_t)lin, nkv, a.B * a.H, a.sched_group);
  w.kt = tcd.tile;
  w.k0 = tcd.tile * 128;
  w.bh = tcd.bh;
  w.b = tcd.bh / a.H;
  w.h = tcd.bh - w.b * a.H;
  w.hk = (a.Hkv == 1) ? 0 : w.h;
  w.i_start = 0;
  if (CAUSAL) {
    const int first_q = w.k0 - a.causal_off;           // first query row that sees key k0
    w.i_start = first_q > 0 ? (first_q >> 7) : 0;
  }
  w.n_iter = nq - w.i_start;                           // <= 0: no query sees this K/V tile (dK = dV = 0)
  return w;
}

// TMA reduce-add shared -> global (fp32 tile added into the tensor at L2), bulk async-group completion
FASN_DEVICE void tma_reduce_add_4d(const CUtensorMap* m, const void* smem_src, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.reduce.async.bulk.tensor.4d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3, %4, %5}], [%1];\n" ::"l"(
          reinterpret_cast<uint64_t>(m)),
      "r"(smem_u32(smem_src)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
template <int N> FASN_DEVICE void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;\n" ::"n"(N) : "memory"); }

// cp.async.bulk 1-D global -> shared with mbarrier completion
FASN_DEVICE void bulk_load_1d(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(smem_u32(smem_dst)),
               "l"(reinterpret_cast<uint64_t>(gsrc)), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// AUX = true: dense attn_mask / attn_bias tensors (generic path); AUX = false keeps that code out of the fast kernels
// (a key-padding mask with row stride 0 is handled by both).
template <int D, bool BF16, bool CAUSAL, bool DROPOUT, bool AUX>
__global__ void __launch_bounds__(kBwdThreads, 1)
fasn_bwd_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_k,
                const __grid_constant__ CUtensorMap tm_v, const __grid_constant__ CUtensorMap tm_do,
                const __grid_constant__ CUtensorMap tm_dk, const __grid_constant__ CUtensorMap tm_dv,
                const __grid_constant__ CUtensorMap tm_dq, const BwdArgs a,
                const TensorView dk_view, const TensorView dv_view) {
  using Cfg = BwdCfg<D>;
  constexpr int DB = Cfg::DB, TILE_BYTES = Cfg::TILE_BYTES, BLK_BYTES = Cfg::BLK_BYTES;
  constexpr uint32_t TM_S = 0, TM_DP = 128, TM_DQ = 128, TM_DV = 256, TM_DK = 256 + D;

  TL_CTA_BEGIN();
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int nkv = (a.Skv + 127) >> 7;
  const int nq = (a.Sq + 127) >> 7;
  const int total = nkv * a.B * a.H;

  extern __shared__ __align__(1024) uint8_t smem[];
  if ((smem_u32(smem) & 1023u) != 0) __trap();      // 128B-swizzled tiles need 1024-byte alignment
  uint8_t* sK = smem;
  uint8_t* sV = sK + TILE_BYTES;
  uint8_t* sQ = sV + TILE_BYTES;              // [2]
  uint8_t* sDO = sQ + 2 * TILE_BYTES;         // [1]
  uint8_t* sDS = sDO + TILE_BYTES;
  uint8_t* sDQ = sDS + Cfg::DS_BYTES;         // [2] fp32 staging chunks for the TMA reduce-add of dQ
  float* sLse = reinterpret_cast<float*>(sDQ + 2 * Cfg::DQ_STAGE_BYTES);   // [2][128]
  float* sDelta = sLse + 256;                                              // [2][128]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sDelta + 256);
  // Parities: barriers that complete once per Q tile are indexed by the CTA's running tile count g (over all its items),
  // barriers that complete once per item by the count m of items with work.
  uint64_t* k_full = bars + 0;     // per item
  uint64_t* q_full = bars + 1;     // [2]
  uint64_t* q_empty = bars + 3;    // [2]
  uint64_t* do_full = bars + 5;
  uint64_t* v_free = bars + 6;     // per item: the TMA store of dV has read its staging tile (sV)
  uint64_t* do_empty = bars + 7;
  uint64_t* s_full = bars + 9;
  uint64_t* dp_full = bars + 10;
  uint64_t* p_full = bars + 11;    // 256 arrivals
  uint64_t* ds_full = bars + 12;   // 256 arrivals
  uint64_t* ds_empty = bars + 13;
  uint64_t* dq_full = bars + 14;
  uint64_t* dq_empty = bars + 15;  // 128 arrivals
  uint64_t* dkv_full = bars + 16;  // per item: every MMA of the item has completed
  uint64_t* dv_full = bars + 17;   // per item: the last dV MMA has completed: the dV epilogue overlaps the last dQ / dK MMAs
  uint64_t* item_full = bars + 18; // [2]  work-item ring: the producer has published an item index
  uint64_t* item_empty = bars + 20;// [2]  14 arrivals: the MMA warp, the store warp, the eight compute warps and the four reducer warps have read it
  uint64_t* v_full = bars + 22;    // per item
  uint64_t* dv_staged = bars + 23; // per item, 256 arrivals: dV is in its staging tile (sV)
  uint64_t* dk_staged = bars + 24; // per item, 256 arrivals: dK is in its staging tile (the dS^T tile)
  uint64_t* ds_free = bars + 25;   // per item: the TMA store of dK has read the dS^T tile
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + Cfg::NUM_BARS);
  volatile int* sItem = reinterpret_cast<volatile int*>(tmem_slot + 2);   // [2]

  const float* lse2_ws = a.delta + (long long)a.B * a.H * a.Sqp;     // second half of the workspace: LSE_n * log2e
  if (warp == 12 && lane == 0) {
    tma_prefetch_desc(&tm_q); tma_prefetch_desc(&tm_k); tma_prefetch_desc(&tm_v); tma_prefetch_desc(&tm_do);
    tma_prefetch_desc(&tm_dk); tma_prefetch_desc(&tm_dv); tma_prefetch_desc(&tm_dq);
    mbar_init(k_full, 1); mbar_init(v_full, 1);
    for (int i = 0; i < 2; ++i) { mbar_init(&q_full[i], 1); mbar_init(&q_empty[i], 1); mbar_init(&item_full[i], 1); mbar_init(&item_empty[i], 14); }
    mbar_init(do_full, 1); mbar_init(do_empty, 1); mbar_init(v_free, 1);
    mbar_init(dv_staged, 256); mbar_init(dk_staged, 256); mbar_init(ds_free, 1);
    mbar_init(s_full, 1); mbar_init(dp_full, 1); mbar_init(p_full, 256); mbar_init(ds_full, 256); mbar_init(ds_empty, 1);
    mbar_init(dq_full, 1); mbar_init(dq_empty, 128); mbar_init(dkv_full, 1); mbar_init(dv_full, 1);
    fence_mbar_init();
    fence_proxy_async_smem();
  }
  if (warp == 14) { tmem_alloc<512>(tmem_slot); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  uint32_t iters_done = 0;                    // timeline builds: Q tiles of all items of this CTA (thread 0)

  // Every consumer role reads the next item index from the two-deep ring the producer publishes.
  auto next_item = [&](uint32_t item_n) -> int {
    const int si = item_n & 1;
    mbar_wait(&item_full[si], (item_n >> 1) & 1);
    const int lin = sItem[si];
    __syncwarp();
    if (lane == 0) mbar_arrive(&item_empty[si]);
    return lin;
  };

  if (warp >= 12) {
    setmaxnreg_dec<BwdRegs<D>::kOther>();
    if (warp == 12 && lane == 0) {
      // ---------------------------------------------------------------- TMA producer
      int lin = blockIdx.x;
      uint32_t item_n = 0, m = 0, g = 0;     // items published, items with work, Q tiles loaded
      TL_DECL(4)
      while (true) {
        const int si = item_n & 1;
        if (item_n >= 2) mbar_wait(&item_empty[si], ((item_n >> 1) - 1) & 1);
        sItem[si] = (lin < total) ? lin : -1;
        mbar_arrive(&item_full[si]);             // release: the index is visible to whoever observes the phase
        if (lin >= total) break;
        const BwdItem w = bwd_decode<CAUSAL>(a, lin, nkv, nq);
        lin = (int)gridDim.x + atomicAdd(a.sched, 1);      // the next item (its latency hides behind this item's loads)
        ++item_n;
        if (w.n_iter <= 0) continue;
        for (int it = 0; it < w.n_iter; ++it, ++g) {
          const int s = g & 1;
          const uint32_t ph = (g >> 1) & 1;
          const int qi0 = (w.i_start + it) * 128;
          mbar_wait(&q_empty[s], ph ^ 1);
          mbar_arrive_expect_tx(&q_full[s], TILE_BYTES + 1024);
#pragma unroll
          for (int db = 0; db < DB; ++db) tma_load_4d(sQ + s * TILE_BYTES + db * BLK_BYTES, &tm_q, &q_full[s], db * 64, qi0, w.h, w.b);
          bulk_load_1d(sLse + s * 128, lse2_ws + (long long)w.bh * a.Sqp + qi0, 512, &q_full[s]);
          bulk_load_1d(sDelta + s * 128, a.delta + (long long)w.bh * a.Sqp + qi0, 512, &q_full[s]);
          mbar_wait(do_empty, (g & 1) ^ 1);     // single dO buffer: free once dV of the previous tile has completed
          mbar_arrive_expect_tx(do_full, TILE_BYTES);
#pragma unroll
          for (int db = 0; db < DB; ++db) tma_load_4d(sDO + db * BLK_BYTES, &tm_do, do_full, db * 64, qi0, w.h, w.b);
          if (it == 0) {
            // K / V of this item, after Q / dO of its first tile (their buffers are free earlier).  The last reader of sK is the dQ
            // MMA of the previous item's last tile (tile g - 1 of this CTA); sV is the staging tile of the previous item's dV store.
            TL(40);
            if (g > 0) mbar_wait(dq_full, (g - 1) & 1);
            TL(41);
            mbar_arrive_expect_tx(k_full, TILE_BYTES);
#pragma unroll
            for (int db = 0; db < DB; ++db) tma_load_4d(sK + db * BLK_BYTES, &tm_k, k_full, db * 64, w.k0, w.hk, w.b);
            if (m > 0) mbar_wait(v_free, (m - 1) & 1);
            TL(42);
            mbar_arrive_expect_tx(v_full, TILE_BYTES);
#pragma unroll
            for (int db = 0; db < DB; ++db) tma_load_4d(sV + db * BLK_BYTES, &tm_v, v_full, db * 64, w.k0, w.hk, w.b);
          }
        }
        ++m;
      }
    } else if (warp == 13) {
      // ---------------------------------------------------------------- MMA issuer
      // The whole warp walks the schedule (operands stay in uniform registers), one elected lane issues; descriptors
      // are base words computed once plus compile-time offsets, so a tcgen05.mma costs a few issue slots.
      constexpr uint32_t idesc_kk = umma_idesc(BF16, 128, 128, false, false);   // S^T, dP^T
      constexpr uint32_t idesc_dv = umma_idesc(BF16, 128, D, false, true);      // A in TMEM, B MN-major
      constexpr uint32_t idesc_dk = umma_idesc(BF16, 128, D, false, true);      // A K-major smem, B MN-major
      constexpr uint32_t idesc_dq = umma_idesc(BF16, 128, D, true, true);       // A, B MN-major
      constexpr uint32_t hi_desc = umma_desc_hi(1024);
      constexpr uint32_t TILE16 = TILE_BYTES >> 4;
      const uint32_t tm = __shfl_sync(0xffffffffu, tmem_base, 0);
      // Descriptor low words = (shared-memory address >> 4) | (LBO >> 4) << 16.  Every operand tile sits at a compile-time
      // offset from the CTA's shared-memory base, so one register plus immediates describes them all; the base is made
      // opaque once per Q tile (asm below) so that the compiler recomputes base + immediate next to each MMA instead of
      // hoisting ~60 loop-invariant descriptor words into registers it does not have (this warp runs on 56 registers).
      //   K-major views: LBO unused (1);  MN-major views: LBO = next 64-wide block
      constexpr uint32_t KM = 1u << 16, MN = static_cast<uint32_t>(BLK_BYTES >> 4) << 16;
      constexpr uint32_t oK = 0, oV = TILE16, oQ = 2 * TILE16, oDO = 4 * TILE16, oDS = 5 * TILE16;
      uint32_t sb = (smem_u32(smem) & 0x3FFFF) >> 4;
      auto issue_kmajor = [&](uint32_t tm_dst, uint32_t a_lo, uint32_t b_lo) {  // D[128x128] = A B^T over K = head dim
#pragma unroll
        for (int kb = 0; kb < D / 16; ++kb) {
          const uint32_t off = ((kb >> 2) * BLK_BYTES + (kb & 3) * 32) >> 4;
          umma_ss(tm + tm_dst, umma_desc_join(a_lo + off, hi_desc), umma_desc_join(b_lo + off, hi_desc), idesc_kk, kb > 0 ? 1u : 0u);
        }
      };
      uint32_t item_n = 0, m = 0, g0 = 0;      // items seen, items with work, Q tiles of the items before this one
      TL_DECL(0)
      TL_ONLY(lane == 0);
      while (true) {
        const int lin = next_item(item_n);
        if (lin < 0) break;
        ++item_n;
        const BwdItem w = bwd_decode<CAUSAL>(a, lin, nkv, nq);
        const int n_iter = w.n_iter;
        if (n_iter <= 0) continue;
        // First tile of the item.  S^T overwrites P^T of the previous item's last tile (its dV MMA was issued earlier on the same
        // pipe); dP^T overwrites the previous item's last dQ, which the reducers must have drained.
        mbar_wait(k_full, m & 1);
        mbar_wait(&q_full[g0 & 1], (g0 >> 1) & 1);
        tc_fence_after();
        TL(1);
        if (elect_one()) {
          issue_kmajor(TM_S, sb + oK + KM, sb + oQ + KM + (g0 & 1) * TILE16);
          tc_commit(s_full);
        }
        __syncwarp();
        mbar_wait(v_full, m & 1);
        mbar_wait(do_full, g0 & 1);
        if (g0 > 0) mbar_wait(dq_empty, (g0 - 1) & 1);
        tc_fence_after();
        if (elect_one()) { issue_kmajor(TM_DP, sb + oV + KM, sb + oDO + KM); tc_commit(dp_full); }
        __syncwarp();
        for (int it = 0; it < n_iter; ++it) {
          asm volatile("" : "+r"(sb));
          const uint32_t g = g0 + it;
          const int s = g & 1;
          const int s1 = s ^ 1;
          const uint32_t ph1 = ((g + 1) >> 1) & 1;
          const bool more = it + 1 < n_iter;
          // dV += P^T dO_i
          mbar_wait(p_full, g & 1);
          tc_fence_after();
          TL(2);
          if (elect_one()) {
#pragma unroll
            for (int kb = 0; kb < 8; ++kb)
              umma_ts(tm + TM_DV, tm + TM_S + (kb >> 2) * 64 + (kb & 3) * 8, umma_desc_join(sb + oDO + MN + kb * (2048 >> 4), hi_desc),
                      idesc_dv, (it > 0 || kb > 0) ? 1u : 0u);
            tc_commit(do_empty);         // dO_i is dead once dP^T_i (issued earlier) and dV_i have completed
            if (!more) tc_commit(dv_full);
          }
          __syncwarp();
          // S^T of the next Q tile (overwrites P^T: ordered behind the dV MMAs on the tensor pipe)
          if (more) {
            mbar_wait(&q_full[s1], ph1);
            tc_fence_after();
            TL(3);
            if (elect_one()) {
              issue_kmajor(TM_S, sb + oK + KM, sb + oQ + KM + s1 * TILE16);
              tc_commit(s_full);
            }
            __syncwarp();
          }
          // dQ_i = dS K first (its consumers, the reducer warps, then drain TMEM while dK executes) ;  dK += dS^T Q_i
          mbar_wait(ds_full, g & 1);
          tc_fence_after();
          TL(4);
          if (elect_one()) {
#pragma unroll
            for (int kb = 0; kb < 8; ++kb)
              umma_ss(tm + TM_DQ, umma_desc_join(sb + oDS + MN + kb * (2048 >> 4), hi_desc), umma_desc_join(sb + oK + MN + kb * (2048 >> 4), hi_desc),
                      idesc_dq, kb > 0 ? 1u : 0u);
            tc_commit(dq_full);
          }
          __syncwarp();
          auto issue_dk = [&](int kb0, int kb1) {
#pragma unroll
            for (int kb = kb0; kb < kb1; ++kb) {
              const uint32_t off = ((kb >> 2) * BLK_BYTES + (kb & 3) * 32) >> 4;
              umma_ss(tm + TM_DK, umma_desc_join(sb + oDS + KM + off, hi_desc), umma_desc_join(sb + oQ + MN + s * TILE16 + kb * (2048 >> 4), hi_desc),
                      idesc_dk, (it > 0 || kb > 0) ? 1u : 0u);
            }
          };
          constexpr int kSplit = FASN_BWD_DK_SPLIT;
          if (elect_one()) issue_dk(0, more ? kSplit : 8);
          __syncwarp();
          // dP^T of the next tile reuses the dQ columns: wait until the reducers have drained dQ_i
          if (more) {
            mbar_wait(do_full, (g + 1) & 1);
            TL(5);
            mbar_wait(dq_empty, g & 1);
            tc_fence_after();
            TL(6);
            if (elect_one()) { issue_kmajor(TM_DP, sb + oV + KM, sb + oDO + KM); tc_commit(dp_full); }
            __syncwarp();
          }
          if (elect_one()) {
            if (more) issue_dk(kSplit, 8);
            tc_commit(&q_empty[s]);      // Q_i, LSE2_i and delta_i stay valid until the compute warps are done with tile i
                                         // (ds_full above) and the MMAs that read Q_i have completed
            tc_commit(ds_empty);
            if (!more) tc_commit(dkv_full);
          }
          __syncwarp();
        }
        g0 += n_iter; ++m;
      }
    } else if (warp == 15) {
      // ---------------------------------------------------------------- epilogue stores
      // dV / dK of an item leave through TMA stores from their staging tiles.  This warp issues them and waits until they have
      // read shared memory, then hands the tiles on (sV to the producer, the dS^T tile to the compute warps of the next item):
      // a compute thread that waited here kept its whole warp, and with it the dK epilogue, back by ~1500 cycles per item.
      uint32_t item_n = 0, m = 0;
      while (true) {
        const int lin = next_item(item_n);
        if (lin < 0) break;
        ++item_n;
        const BwdItem w = bwd_decode<CAUSAL>(a, lin, nkv, nq);
        if (w.n_iter <= 0) continue;
        const bool head_sum = a.dk_accum != nullptr;     // shared K/V: the compute warps added dK / dV into the accumulators themselves
        mbar_wait(dv_staged, m & 1);
        if (lane == 0) {
          if (!head_sum) {
#pragma unroll
            for (int db = 0; db < DB; ++db) tma_store_4d(&tm_dv, sV + db * BLK_BYTES, db * 64, w.k0, w.h, w.b);
            tma_store_commit();
            tma_store_wait_read_all();
          }
          mbar_arrive(v_free);
        }
        __syncwarp();
        mbar_wait(dk_staged, m & 1);
        if (lane == 0) {
          if (!head_sum) {
#pragma unroll
            for (int db = 0; db < DB; ++db) tma_store_4d(&tm_dk, sDS + db * BLK_BYTES, db * 64, w.k0, w.h, w.b);
            tma_store_commit();
            tma_store_wait_read_all();
          }
          mbar_arrive(ds_free);
        }
        __syncwarp();
        ++m;
      }
    }
  } else if (warp >= 8) {
    // ------------------------------------------------------------------ dQ reducers
    // TMEM -> registers -> swizzled fp32 staging chunk in smem -> TMA reduce-add into dq_accum (the L2 does the adds
    // a full 128-byte line at a time; per-thread red.global instructions are an order of magnitude slower here).
    // The whole dQ tile is pulled into registers first and its tensor-memory columns are released at once: dP^T of the next
    // Q tile reuses them, and the chain dS -> dQ MMA -> drain -> dP^T MMA -> dS is the critical path of an iteration (the
    // staging stores below compete for the shared-memory port with the tensor core's operand reads and take ~1000 cycles).
    if constexpr (BwdRegs<D>::kReduce > 128) setmaxnreg_inc<BwdRegs<D>::kReduce>(); else setmaxnreg_dec<BwdRegs<D>::kReduce>();
    const int r = (warp & 3) * 32 + lane;                 // query row inside the tile
    const uint32_t lane_off = static_cast<uint32_t>((warp & 3) * 32) << 16;
    constexpr int NCH = D / 32;                           // 32-column chunks per dQ tile
    uint8_t* const stage_row = sDQ + r * 128;
    const int rx = (r & 7) << 4;
    uint32_t item_n = 0, g0 = 0;
    TL_DECL(3)
    TL_ONLY(threadIdx.x == 256);
    while (true) {
      const int lin = next_item(item_n);
      if (lin < 0) break;
      ++item_n;
      const BwdItem w = bwd_decode<CAUSAL>(a, lin, nkv, nq);
      const int n_iter = w.n_iter, bh = w.bh;
      if (n_iter <= 0) continue;
      for (int it = 0; it < n_iter; ++it) {
        const uint32_t g = g0 + it;
        const int qi0 = (w.i_start + it) * 128;
        mbar_wait(dq_full, g & 1);
        tc_fence_after();
        TL(30);
        uint32_t v[D];
#pragma unroll
        for (int cb = 0; cb < NCH; ++cb) tmem_ld_x32(tmem_base + lane_off + TM_DQ + cb * 32, v + cb * 32);
        tmem_wait_ld();
        tc_fence_before();
        mbar_arrive(dq_empty);                                // the dQ columns may be overwritten by dP^T now
        TL(31);
#pragma unroll
        for (int hb = 0; hb < NCH / 2; ++hb) {                // 64 columns (both staging chunks) per round
          if (threadIdx.x == 256) tma_store_wait_read<0>();   // the previous round's reduces have read both staging chunks
          named_bar_sync(2, 128);
#pragma unroll
          for (int ch = 0; ch < 2; ++ch)
#pragma unroll
            for (int gg = 0; gg < 8; ++gg)
              *reinterpret_cast<uint4*>(stage_row + ch * Cfg::DQ_STAGE_BYTES + ((gg << 4) ^ rx)) =
                  make_uint4(v[hb * 64 + ch * 32 + gg * 4], v[hb * 64 + ch * 32 + gg * 4 + 1], v[hb * 64 + ch * 32 + gg * 4 + 2], v[hb * 64 + ch * 32 + gg * 4 + 3]);
          fence_proxy_async_smem();
          named_bar_sync(3, 128);
          if (threadIdx.x == 256) {
            tma_reduce_add_4d(&tm_dq, sDQ, hb * 64, qi0, bh, 0);
            tma_reduce_add_4d(&tm_dq, sDQ + Cfg::DQ_STAGE_BYTES, hb * 64 + 32, qi0, bh, 0);
            tma_store_commit();
          }
        }
      }
      g0 += n_iter;
    }
    if (threadIdx.x == 256) tma_store_wait_read<0>();   // the staging chunks have been read; the L2 adds complete on their own before the grid ends
  } else {
    // -------------------------------------------------------------------- compute warps
    setmaxnreg_inc<BwdRegs<D>::kCompute>();
    const int quarter = warp & 3;
    const int half = warp >> 2;
    const int r = quarter * 32 + lane;                     // kv row inside the tile (row layout: AUX loop, epilogue)
    const uint32_t lane_off = static_cast<uint32_t>(quarter * 32) << 16;
    const float2 c2 = make_float2(a.scale_log2, a.scale_log2);
    uint32_t item_n = 0, m = 0, g0 = 0;                    // items seen, items with work, Q tiles of the items before this one
    TL_DECL(1 + half)
    TL_ONLY(threadIdx.x == 0 || threadIdx.x == 128);
    while (true) {
    const int lin = next_item(item_n);
    if (lin < 0) break;
    ++item_n;
    const BwdItem w = bwd_decode<CAUSAL>(a, lin, nkv, nq);
    const int k0 = w.k0, bh = w.bh, b = w.b, h = w.h, i_start = w.i_start, n_iter = w.n_iter;
    TL(20);                                                // item fetched
    if (n_iter <= 0) {
      // no query sees this K/V tile: dK = dV = 0 (nothing to add when the heads are summed into zero-filled accumulators)
      if (threadIdx.x < 128 && a.dk_accum == nullptr) {
        const int row = k0 + threadIdx.x;
        if (row < a.Skv) {
          uint4* pk = reinterpret_cast<uint4*>(reinterpret_cast<uint16_t*>(dk_view.ptr) + b * dk_view.sb + h * dk_view.sh + (long long)row * dk_view.ss);
          uint4* pv = reinterpret_cast<uint4*>(reinterpret_cast<uint16_t*>(dv_view.ptr) + b * dv_view.sb + h * dv_view.sh + (long long)row * dv_view.ss);
#pragma unroll
          for (int i = 0; i < D / 8; ++i) { pk[i] = make_uint4(0, 0, 0, 0); pv[i] = make_uint4(0, 0, 0, 0); }
        }
      }
      continue;
    }
    const uint8_t* mask_bh = a.mask.ptr ? reinterpret_cast<const uint8_t*>(a.mask.ptr) + b * a.mask.sb + h * a.mask.sh : nullptr;
    // A mask that is broadcast over the query axis (key padding, (B|1, H|1, 1, S)) is one byte per key: a key is either
    // visible to every query or to none, which the fast path handles like a row beyond Skv.
    const bool key_only_mask = (mask_bh != nullptr) && (a.mask.sq == 0);
    const uint32_t bh_global = a.bh_offset + bh;
    const uint32_t kvw = (uint32_t)((k0 + quarter * 32) >> 5);   // 32-key word index of this warp's kv rows
    const bool kv_tail = (k0 + 128 > a.Skv) || key_only_mask;   // this K/V tile (may) have rows that no query sees

    if constexpr (!AUX) {
      // ---------------------------------------------------------------- fast kernels: quad layout
      // tcgen05.ld.16x256b hands thread (g = lane / 4, t4 = lane % 4) the kv rows 8 j + g (j = 0..3) of this warp's 32 and
      // the query columns 8 c + 2 t4 + {0, 1} (c = 0..7) of its 64:  v[32 (j / 2) + 4 c + 2 (j % 2) + e].  The packed 16-bit
      // pairs go back to tensor memory with tcgen05.st.16x128b (column 4 c + t4 of the 32 packed columns), and into the
      // dS^T tile with 4-byte stores that are conflict-free under the 128-byte swizzle (8 rows x 4 lanes = 32 banks).
      const int g = lane >> 2, t4 = lane & 3;
      int rowpad[4];                                       // 0 for a real, visible key; 64 pushes the row's first visible column out of the tile
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int kvr = k0 + quarter * 32 + 8 * j + g;
        bool valid = kvr < a.Skv;
        if (key_only_mask && valid) valid = mask_bh[kvr] != 0;
        rowpad[j] = valid ? 0 : 64;
      }
      // Dropout keep words: lane L generates the 32-key words of query rows qc0 + L and qc0 + 32 + L one iteration ahead
      // (in the slot where this warp would otherwise wait for dP^T); each thread then fetches the words of its 16 query
      // columns with shuffles and shifts them so that the bit of kv row 8 j + g sits in the sign position of byte j.
      uint32_t keep_raw0 = 0xFFFFFFFFu, keep_raw1 = 0xFFFFFFFFu;
      auto make_keep = [&](int qc0n) {
        if constexpr (DROPOUT) {
          keep_raw0 = dropout_keep_word(a.key, bh_global, (uint32_t)(qc0n + lane), kvw, a.drop_thr);
          keep_raw1 = dropout_keep_word(a.key, bh_global, (uint32_t)(qc0n + 32 + lane), kvw, a.drop_thr);
        }
      };
      make_keep(i_start * 128 + half * 64);
      uint8_t* const ds_base = sDS + half * BLK_BYTES + (quarter * 32 + g) * 128 + 4 * t4;

      for (int it = 0; it < n_iter; ++it) {
        const uint32_t gt = g0 + it;                        // running Q-tile count of this CTA: ring slot and barrier parities
        const int s = gt & 1;
        const uint32_t ph = (gt >> 1) & 1;
        const int qi0 = (i_start + it) * 128;
        const int qc0 = qi0 + half * 64;                   // first query column of this warp
        TL(10);
        uint32_t kx[DROPOUT ? 16 : 1];                     // kx[2 c + e]: keep word of query column 8 c + 2 t4 + e, shifted left by 7 - g
        if constexpr (DROPOUT) {
#pragma unroll
          for (int c = 0; c < 8; ++c)
#pragma unroll
            for (int e = 0; e < 2; ++e)
              kx[DROPOUT ? 2 * c + e : 0] = __shfl_sync(0xffffffffu, (c < 4) ? keep_raw0 : keep_raw1, ((c & 3) << 3) + 2 * t4 + e) << (7 - g);
        }
        // ---- P^T = 2^(S^T c - LSE2)
        mbar_wait(&q_full[s], ph);                         // LSE2 and delta of this tile have landed (same barrier as Q_i)
        mbar_wait(s_full, gt & 1);
        tc_fence_after();
        TL(11);
        float p[64];
        {
          uint32_t* pr = reinterpret_cast<uint32_t*>(p);
          tmem_ld_16x256b_x8(tmem_base + lane_off + TM_S + half * 64, pr);
          tmem_ld_16x256b_x8(tmem_base + lane_off + (16u << 16) + TM_S + half * 64, pr + 32);
          tmem_wait_ld();
        }
        const float* lse_s = sLse + s * 128 + half * 64 + 2 * t4;
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          const float2 l2 = *reinterpret_cast<const float2*>(lse_s + 8 * c);
          const float2 nl = make_float2(-l2.x, -l2.y);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int idx = (j >> 1) * 32 + 4 * c + 2 * (j & 1);
            const float2 a01 = __ffma2_rn(make_float2(p[idx], p[idx + 1]), c2, nl);
            p[idx] = ex2(a01.x); p[idx + 1] = ex2(a01.y);
          }
        }
        const bool diag = CAUSAL && (qi0 + a.causal_off < k0 + 127);     // some (q, kv) of this tile pair lies above the diagonal
        if (diag || kv_tail) {
          // query column x of this warp's 64 is visible to kv row kvr iff kvr <= q + off  <=>  x >= kvr - off - qc0;
          // x = 8 c + 2 t4 + e, so the compile-time part 8 c + e is compared with a per-row threshold
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int kvr = k0 + quarter * 32 + 8 * j + g;
            const int fc = (diag ? max(kvr - a.causal_off - qc0, 0) : 0) + rowpad[j] - 2 * t4;
#pragma unroll
            for (int c = 0; c < 8; ++c)
#pragma unroll
              for (int e = 0; e < 2; ++e) {
                const int idx = (j >> 1) * 32 + 4 * c + 2 * (j & 1) + e;
                p[idx] = (8 * c + e >= fc) ? p[idx] : 0.f;
              }
          }
        }
#pragma unroll
        for (int h16 = 0; h16 < 2; ++h16) {
          uint32_t pk[16];
#pragma unroll
          for (int c = 0; c < 8; ++c)
#pragma unroll
            for (int u = 0; u < 2; ++u) {
              const int idx = h16 * 32 + 4 * c + 2 * u;
              uint32_t w01 = pack2<BF16>(p[idx], p[idx + 1]);
              // dropped entries are zeroed on the packed pair (1 PRMT + 1 AND per two elements); the 1/(1-p) factor of the
              // kept entries is applied to dV in the epilogue
              if constexpr (DROPOUT) w01 &= keep_byte_pair_mask(kx[DROPOUT ? 2 * c : 0], kx[DROPOUT ? 2 * c + 1 : 0], 2 * h16 + u);
              pk[2 * c + u] = w01;
            }
          tmem_st_16x128b_x8(tmem_base + lane_off + (static_cast<uint32_t>(h16 * 16) << 16) + TM_S + half * 64, pk);   // over S^T columns this warp has read
        }
        tmem_wait_st();
        tc_fence_before();
        mbar_arrive(p_full);
        TL(12);
        if (it + 1 < n_iter) make_keep(qc0 + 128);          // next tile's keep words, while dP^T is still in flight
        // ---- dS'^T = P^T o (Z dP^T - (1-p) delta)      (dS = dS' / (1-p); the factor is folded into the dK / dQ scales)
        const float* del_s = sDelta + s * 128 + half * 64 + 2 * t4;
        float2 nd[8];
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          const float2 d2 = *reinterpret_cast<const float2*>(del_s + 8 * c);
          nd[c] = make_float2(-d2.x, -d2.y);
        }
        mbar_wait(dp_full, gt & 1);
        tc_fence_after();
        TL(13);
        mbar_wait(ds_empty, (gt & 1) ^ 1);  // MMAs of the previous iteration no longer read the dS^T tile
        if (it == 0 && m > 0) mbar_wait(ds_free, (m - 1) & 1);   // ... and the TMA store of the previous item's dK, staged in this tile, has read it
        TL(14);
#pragma unroll
        for (int h16 = 0; h16 < 2; ++h16) {
          uint32_t dpr[32];
          tmem_ld_16x256b_x8(tmem_base + lane_off + (static_cast<uint32_t>(h16 * 16) << 16) + TM_DP + half * 64, dpr);
          tmem_wait_ld();
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            const uint32_t choff = static_cast<uint32_t>((c ^ g) << 4);      // 16-byte chunk c of the row under the 128B swizzle (row & 7 == g)
#pragma unroll
            for (int u = 0; u < 2; ++u) {
              const int j = 2 * h16 + u, idx = 4 * c + 2 * u;
              float dp0 = __uint_as_float(dpr[idx]), dp1 = __uint_as_float(dpr[idx + 1]);
              if constexpr (DROPOUT) {
                dp0 = (kx[DROPOUT ? 2 * c : 0] & (0x80u << (8 * j))) ? dp0 : 0.f;
                dp1 = (kx[DROPOUT ? 2 * c + 1 : 0] & (0x80u << (8 * j))) ? dp1 : 0.f;
              }
              const float2 e01 = __fadd2_rn(make_float2(dp0, dp1), nd[c]);
              const float2 s01 = __fmul2_rn(make_float2(p[h16 * 32 + idx], p[h16 * 32 + idx + 1]), e01);
              *reinterpret_cast<uint32_t*>(ds_base + j * (8 * 128) + choff) = pack2<BF16>(s01.x, s01.y);
            }
          }
        }
        tc_fence_before();
        fence_proxy_async_smem();
        mbar_arrive(ds_full);
        TL(15);
      }
    } else {
      // ---------------------------------------------------------------- dense mask / bias / ALiBi kernels: row layout
      // (thread = one kv row x 64 query columns: its bias / mask elements are one coalesced 64-byte segment per load)
      const int kv_row = k0 + r;
      const int kv_c = min(kv_row, a.Skv - 1);
      const uint8_t* mbase = (mask_bh != nullptr && !key_only_mask) ? mask_bh + kv_c : nullptr;
      const bool key_masked = key_only_mask && (mask_bh[kv_c] == 0);
      const bool kv_valid = (kv_row < a.Skv) && !key_masked;
      const bool has_aux = (mbase != nullptr) || (a.bias.ptr != nullptr) || (a.alibi != nullptr);
      const float alibi2 = (a.alibi != nullptr) ? a.alibi[h] * kLog2e : 0.f;
      const uint16_t* bbase = a.bias.ptr ? reinterpret_cast<const uint16_t*>(a.bias.ptr) + b * a.bias.sb + h * a.bias.sh + kv_c : nullptr;
      uint16_t* const dbias_row = a.dbias.ptr ? reinterpret_cast<uint16_t*>(const_cast<void*>(a.dbias.ptr)) + b * a.dbias.sb + h * a.dbias.sh + kv_c : nullptr;
      // Dropout keep bits: lane L generates the 32-key keep words of query rows qc0+L and qc0+32+L; a 32x32 bit
      // transpose across the warp then hands every lane (= kv row) its own bit of all 64 query columns.
      uint32_t keep_next0 = 0xFFFFFFFFu, keep_next1 = 0xFFFFFFFFu;
      auto make_keep = [&](int qc0n) {
        if constexpr (DROPOUT) {
          keep_next0 = warp_transpose_bits(dropout_keep_word(a.key, bh_global, (uint32_t)(qc0n + lane), kvw, a.drop_thr), lane);
          keep_next1 = warp_transpose_bits(dropout_keep_word(a.key, bh_global, (uint32_t)(qc0n + 32 + lane), kvw, a.drop_thr), lane);
        }
      };
      make_keep(i_start * 128 + half * 64);

      for (int it = 0; it < n_iter; ++it) {
        const uint32_t gt = g0 + it;                        // running Q-tile count of this CTA: ring slot and barrier parities
        const int s = gt & 1;
        const uint32_t ph = (gt >> 1) & 1;
        const int qi0 = (i_start + it) * 128;
        const int qc0 = qi0 + half * 64;                     // first query column of this thread
        // keep0 / keep1: bit c = keep decision for (query qc0 + c [+32], this thread's kv row); generated one
        // iteration ahead (below), in the slot where this warp would otherwise wait for dP^T
        const uint32_t keep0 = keep_next0, keep1 = keep_next1;
        // ---- P^T = 2^(S^T c + bias - LSE2)
        TL(10);
        // This thread needs one column of the (L, S) matrix -- its key, 64 queries.  The 32 lanes of a warp read 32
        // consecutive keys, so each load instruction is one coalesced 64-byte segment; all 64 (+64) loads are independent
        // of S^T and are issued here, before the wait for the tensor core, so their latency overlaps it.  Rows beyond Sq
        // are clamped (their P is 0 anyway: LSE2 = +inf on padding rows).
        uint32_t bpk[32];
        uint32_t mb0 = 0xFFFFFFFFu, mb1 = 0xFFFFFFFFu;
        if (has_aux) {
          if (bbase) {
#pragma unroll
            for (int c = 0; c < 64; c += 2) {
              const uint32_t lo = __ldg(bbase + (long long)min(qc0 + c, a.Sq - 1) * a.bias.sq);
              const uint32_t hi = __ldg(bbase + (long long)min(qc0 + c + 1, a.Sq - 1) * a.bias.sq);
              bpk[c >> 1] = lo | (hi << 16);
            }
          }
          if (mbase) {
            uint32_t m0 = 0u, m1 = 0u;
#pragma unroll
            for (int c = 0; c < 32; ++c) {
              m0 |= (__ldg(mbase + (long long)min(qc0 + c, a.Sq - 1) * a.mask.sq) != 0 ? 1u : 0u) << c;
              m1 |= (__ldg(mbase + (long long)min(qc0 + 32 + c, a.Sq - 1) * a.mask.sq) != 0 ? 1u : 0u) << c;
            }
            mb0 = m0; mb1 = m1;
          }
        }
        mbar_wait(&q_full[s], ph);                           // LSE2 and delta of this tile have landed (same barrier as Q_i)
        mbar_wait(s_full, gt & 1);
        tc_fence_after();
        TL(11);
        float p[64];
        {
          uint32_t* pr = reinterpret_cast<uint32_t*>(p);
          tmem_ld_x32(tmem_base + lane_off + TM_S + half * 64, pr);
          tmem_ld_x32(tmem_base + lane_off + TM_S + half * 64 + 32, pr + 32);
          tmem_wait_ld();
        }
        const float* lse_s = sLse + s * 128 + half * 64;
        if (has_aux) {
          // bias added / mask applied in the log2 domain before the exponent
#pragma unroll
          for (int c = 0; c < 64; c += 2) {
            const float2 l2 = *reinterpret_cast<const float2*>(lse_s + c);
            float b0 = alibi2 * (float)(kv_row - a.causal_off - qc0 - c) * kLn2, b1 = b0 - alibi2 * kLn2;   // ALiBi (natural-log units here)
            if (bbase) { b0 = cvt16_to_f32<BF16>((uint16_t)(bpk[c >> 1] & 0xFFFF)); b1 = cvt16_to_f32<BF16>((uint16_t)(bpk[c >> 1] >> 16)); }
            const float x0 = fmaf(p[c], a.scale_log2, b0 * kLog2e) - l2.x;
            const float x1 = fmaf(p[c + 1], a.scale_log2, b1 * kLog2e) - l2.y;
            const uint32_t mw = (c < 32) ? mb0 : mb1;
            p[c] = ((mw >> (c & 31)) & 1u) ? ex2(x0) : 0.f;
            p[c + 1] = ((mw >> ((c + 1) & 31)) & 1u) ? ex2(x1) : 0.f;
          }
        } else {
#pragma unroll
          for (int c = 0; c < 64; c += 4) {
            const float4 l4 = *reinterpret_cast<const float4*>(lse_s + c);
            const float2 a01 = __ffma2_rn(make_float2(p[c], p[c + 1]), c2, make_float2(-l4.x, -l4.y));
            const float2 a23 = __ffma2_rn(make_float2(p[c + 2], p[c + 3]), c2, make_float2(-l4.z, -l4.w));
            p[c] = ex2(a01.x); p[c + 1] = ex2(a01.y); p[c + 2] = ex2(a23.x); p[c + 3] = ex2(a23.y);
          }
        }
        const bool diag = CAUSAL && (qi0 + a.causal_off < k0 + 127);     // some (q, kv) of this tile pair lies above the diagonal
        if (diag || kv_tail) {
          // column c is visible to this kv row iff kv_row <= q + off  <=>  c >= kv_row - off - qc0
          const int first_c = (diag ? max(kv_row - a.causal_off - qc0, 0) : 0) + (kv_valid ? 0 : 64);
#pragma unroll
          for (int c = 0; c < 64; ++c) p[c] = (c >= first_c) ? p[c] : 0.f;
        }
        {
          uint32_t pk[32];
#pragma unroll
          for (int c = 0; c < 64; c += 2) {
            uint32_t w01 = pack2<BF16>(p[c], p[c + 1]);
            if constexpr (DROPOUT) w01 &= keep_pair_mask((c < 32) ? keep0 : keep1, c & 31);
            pk[c >> 1] = w01;
          }
          tmem_st_x32(tmem_base + lane_off + TM_S + half * 64, pk);   // over this thread's own S^T columns only
          tmem_wait_st();
        }
        tc_fence_before();
        mbar_arrive(p_full);
        TL(12);
        if (it + 1 < n_iter) make_keep(qc0 + 128);            // next tile's keep bits, while dP^T is still in flight
        // ---- dS'^T = P^T o (Z dP^T - (1-p) delta)      (dS = dS' / (1-p); the factor is folded into the dK / dQ scales)
        mbar_wait(dp_full, gt & 1);
        tc_fence_after();
        TL(13);
        mbar_wait(ds_empty, (gt & 1) ^ 1);  // MMAs of the previous iteration no longer read the dS^T tile
        if (it == 0 && m > 0) mbar_wait(ds_free, (m - 1) & 1);   // ... and the TMA store of the previous item's dK, staged in this tile, has read it
        TL(14);
        const float* del_s = sDelta + s * 128 + half * 64;
#pragma unroll
        for (int g = 0; g < 2; ++g) {
          uint32_t dpr[32];
          tmem_ld_x32(tmem_base + lane_off + TM_DP + half * 64 + g * 32, dpr);
          tmem_wait_ld();
          uint32_t out[16];
#pragma unroll
          for (int c = 0; c < 32; c += 4) {
            const float4 d4 = *reinterpret_cast<const float4*>(del_s + g * 32 + c);
            float dp0 = __uint_as_float(dpr[c]), dp1 = __uint_as_float(dpr[c + 1]), dp2 = __uint_as_float(dpr[c + 2]), dp3 = __uint_as_float(dpr[c + 3]);
            if constexpr (DROPOUT) {
              const uint32_t w = (g == 0) ? keep0 : keep1;
              dp0 = (w & (1u << (c + 0))) ? dp0 : 0.f;
              dp1 = (w & (1u << (c + 1))) ? dp1 : 0.f;
              dp2 = (w & (1u << (c + 2))) ? dp2 : 0.f;
              dp3 = (w & (1u << (c + 3))) ? dp3 : 0.f;
            }
            const float2 e01 = __fadd2_rn(make_float2(dp0, dp1), make_float2(-d4.x, -d4.y));
            const float2 e23 = __fadd2_rn(make_float2(dp2, dp3), make_float2(-d4.z, -d4.w));
            const float2 s01 = __fmul2_rn(make_float2(p[g * 32 + c], p[g * 32 + c + 1]), e01);
            const float2 s23 = __fmul2_rn(make_float2(p[g * 32 + c + 2], p[g * 32 + c + 3]), e23);
            out[(c >> 1)] = pack2<BF16>(s01.x, s01.y);
            out[(c >> 1) + 1] = pack2<BF16>(s23.x, s23.y);
            if (dbias_row != nullptr && kv_row < a.Skv) {
              // gradient of the logits (= of a dense attn_bias): dS = dS' / P(keep); one 2-byte store per element, the 32 lanes
              // of a warp cover 32 consecutive keys of one query row
              const int qq = qc0 + g * 32 + c;
              const float v4[4] = {s01.x, s01.y, s23.x, s23.y};
#pragma unroll
              for (int e = 0; e < 4; ++e)
                if (qq + e < a.Sq)
                  dbias_row[(long long)(qq + e) * a.dbias.sq] = (uint16_t)(pack2<BF16>(v4[e] * (DROPOUT ? a.inv_keep : 1.f), 0.f) & 0xFFFF);
            }
          }
#pragma unroll
          for (int q4 = 0; q4 < 4; ++q4) {
            const int cc = g * 4 + q4;
            *reinterpret_cast<uint4*>(sDS + half * BLK_BYTES + r * 128 + ((cc ^ (r & 7)) << 4)) =
                make_uint4(out[q4 * 4], out[q4 * 4 + 1], out[q4 * 4 + 2], out[q4 * 4 + 3]);
          }
        }
        tc_fence_before();
        fence_proxy_async_smem();
        mbar_arrive(ds_full);
        TL(15);
      }
    }

    // ------------------------------------------------------------------ dK, dV epilogue
    // V is dead after the last dP^T: stage dV into sV as soon as the last dV MMA is done (the last dQ / dK MMAs are still
    // running), then dK into the dS^T tile, which is dead with the last MMA (same swizzled [block][row][128 B] layout) -- sK
    // stays free for the K tile of the next item, which the producer loads while dK is drained.
#pragma unroll
    for (int which = 0; which < 2; ++which) {
      mbar_wait(which == 0 ? dv_full : dkv_full, m & 1);
      tc_fence_after();
      TL(21 + which);                                      // 21: last dV MMA done, 22: every MMA of the item done
      uint8_t* stage = which == 0 ? sV : sDS;
      const uint32_t tm_src = which == 0 ? TM_DV : TM_DK;
      const float mul = which == 0 ? (DROPOUT ? a.inv_keep : 1.f) : a.scale;   // a.scale already carries 1/(1-p)
      constexpr int COLS = D / 2;                          // columns per thread (this warp's half)
      // Shared K/V (3-D key / value: one K/V head serves every query head): dK / dV of all heads of a batch element belong to the
      // same rows, so they are added into float32 accumulators with 16-byte reductions at the L2 instead of being written per
      // head for the host to sum (H times the memory and a pass over it).  Off the headline path: taken only when the caller
      // passes the accumulators.
      float* const acc = which == 0 ? a.dv_accum : a.dk_accum;
      float* const acc_row = acc ? acc + ((long long)b * a.Skv + (k0 + r)) * D + half * COLS : nullptr;
#pragma unroll
      for (int cb = 0; cb < COLS / 32; ++cb) {
        uint32_t v[32];
        tmem_ld_x32(tmem_base + lane_off + tm_src + half * COLS + cb * 32, v);
        tmem_wait_ld();
        if (acc != nullptr) {
          if (k0 + r < a.Skv) {
#pragma unroll
            for (int g4 = 0; g4 < 8; ++g4)
              red_add_v4(acc_row + cb * 32 + g4 * 4, __uint_as_float(v[g4 * 4]) * mul, __uint_as_float(v[g4 * 4 + 1]) * mul,
                         __uint_as_float(v[g4 * 4 + 2]) * mul, __uint_as_float(v[g4 * 4 + 3]) * mul);
          }
          continue;
        }
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          uint4 w;
          w.x = pack2<BF16>(__uint_as_float(v[g * 8 + 0]) * mul, __uint_as_float(v[g * 8 + 1]) * mul);
          w.y = pack2<BF16>(__uint_as_float(v[g * 8 + 2]) * mul, __uint_as_float(v[g * 8 + 3]) * mul);
          w.z = pack2<BF16>(__uint_as_float(v[g * 8 + 4]) * mul, __uint_as_float(v[g * 8 + 5]) * mul);
          w.w = pack2<BF16>(__uint_as_float(v[g * 8 + 6]) * mul, __uint_as_float(v[g * 8 + 7]) * mul);
          const int col = half * COLS + cb * 32 + g * 8;
          const int db = col >> 6;
          const int cc = (col & 63) >> 3;
          *reinterpret_cast<uint4*>(stage + db * BLK_BYTES + r * 128 + ((cc ^ (r & 7)) << 4)) = w;
        }
      }
      fence_proxy_async_smem();
      mbar_arrive(which == 0 ? dv_staged : dk_staged);     // warp 15 issues the store
    }
    TL(23);                                                // dK staged, store issued
    g0 += n_iter; ++m;
    if (threadIdx.x == 0) iters_done += n_iter;
    }   // items
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 14) tmem_dealloc<512>(tmem_base);
  // the last CTA to finish hands the work counter back at zero for the next launch that uses this slot
  if (threadIdx.x == 0 && atomicAdd(a.sched + 1, 1) == (int)gridDim.x - 1) { a.sched[0] = 0; a.sched[1] = 0; __threadfence(); }
  TL_CTA_END(iters_done);
}

}  // namespace

template <int D, bool BF16, bool CAUSAL, bool DROPOUT, bool AUX>
static cudaError_t launch_bwd_t2(const CUtensorMap& tq, const CUtensorMap& tk, const CUtensorMap& tv, const CUtensorMap& tdo,
                                const CUtensorMap& tdk, const CUtensorMap& tdv, const CUtensorMap& tdq, const BwdArgs& a, const TensorView& dk,
                                const TensorView& dv, cudaStream_t stream) {
  auto kern = fasn_bwd_kernel<D, BF16, CAUSAL, DROPOUT, AUX>;
  constexpr int smem = BwdCfg<D>::SMEM_BYTES;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  if (e != cudaSuccess) return e;
  dim3 grid(a.grid_ctas, 1, 1);      // persistent: one CTA per SM (or per work item when there are fewer)
  kern<<<grid, kBwdThreads, smem, stream>>>(tq, tk, tv, tdo, tdk, tdv, tdq, a, dk, dv);
  return cudaGetLastError();
}

template <int D, bool BF16, bool CAUSAL, bool DROPOUT>
static cudaError_t launch_bwd_t(const CUtensorMap& tq, const CUtensorMap& tk, const CUtensorMap& tv, const CUtensorMap& tdo,
                                const CUtensorMap& tdk, const CUtensorMap& tdv, const CUtensorMap& tdq, const BwdArgs& a, const TensorView& dk,
                                const TensorView& dv, cudaStream_t stream) {
  const bool aux = a.bias.ptr != nullptr || a.alibi != nullptr || (a.mask.ptr != nullptr && a.mask.sq != 0);
  return aux ? launch_bwd_t2<D, BF16, CAUSAL, DROPOUT, true>(tq, tk, tv, tdo, tdk, tdv, tdq, a, dk, dv, stream)
             : launch_bwd_t2<D, BF16, CAUSAL, DROPOUT, false>(tq, tk, tv, tdo, tdk, tdv, tdq, a, dk, dv, stream);
}

cudaError_t launch_bwd(int head_dim, bool bf16, bool causal, bool dropout, const CUtensorMap& tq, const CUtensorMap& tk,
                       const CUtensorMap& tv, const CUtensorMap& tdo, const CUtensorMap& tdk, const CUtensorMap& tdv,
                       const CUtensorMap& tdq, const BwdArgs& a, const TensorView& dk, const TensorView& dv, cudaStream_t stream) {
#define FASN_BWD_CASE(D_, BF_, C_, DR_) \
  if (head_dim == D_ && bf16 == BF_ && causal == C_ && dropout == DR_) return launch_bwd_t<D_, BF_, C_, DR_>(tq, tk, tv, tdo, tdk, tdv, tdq, a, dk, dv, stream);
  FASN_BWD_CASE(64, false, false, false) FASN_BWD_CASE(64, false, false, true)
  FASN_BWD_CASE(64, false, true, false)  FASN_BWD_CASE(64, false, true, true)
  FASN_BWD_CASE(64, true, false, false)  FASN_BWD_CASE(64, true, false, true)
  FASN_BWD_CASE(64, true, true, false)   FASN_BWD_CASE(64, true, true, true)
  FASN_BWD_CASE(128, false, false, false) FASN_BWD_CASE(128, false, false, true)
  FASN_BWD_CASE(128, false, true, false)  FASN_BWD_CASE(128, false, true, true)
  FASN_BWD_CASE(128, true, false, false)  FASN_BWD_CASE(128, true, false, true)
  FASN_BWD_CASE(128, true, true, false)   FASN_BWD_CASE(128, true, true, true)
#undef FASN_BWD_CASE
  return cudaErrorInvalidValue;
}

}  // namespace fasn
