// Forward kernel: O = dropout(softmax_n(scale Q K^T + bias, masked)) V for sm_100a.
//
// Replaces the reference's `_fwd_kernel` (flash_attention_softmax_n/core/flash_attn_triton.py:30-126) and the
// SDPA call of `flash_attention_n` (flash_attention_softmax_n/core/flash_attn.py:115-124).
//
// One CTA owns 256 query rows (two 128-row tiles, "ping-pong") of one (batch, head) unit and streams the
// 128-row K/V tiles past them:
//
//   warp 8      TMA producer   Q tiles once, K/V tiles through an NS-deep shared-memory ring (128B swizzle)
//   warp 9      MMA issuer     S_t = Q_t K_j^T  (tcgen05.mma, operands in smem, accumulator in TMEM)
//                              O_t += P_t V_j   (A = P_t read straight from TMEM, B = V_j MN-major in smem)
//   warps 0-3   softmax, tile 0   one thread per query row: tcgen05.ld S, online softmax_n in the log2
//   warps 4-7   softmax, tile 1   domain, P (16-bit) written back over S with tcgen05.st; epilogue
//
// softmax_n costs nothing per tile: the "+n" of the denominator is a virtual key with logit 0, weight n and
// value 0, i.e. the running (max, sum) start at (0, n) instead of (-inf, 0).  This equals the reference's n
// zero-padded K/V rows (flash_attn.py:66-67) for integer n and its epilogue acc/(n exp(-m) + l)
// (flash_attn_triton.py:114) for real n, without the overflow of exp(-m).
//
// The running max is only moved when it grows by more than 2^8 (lazy rescale), so the O accumulator in TMEM is
// rescaled rarely; exponent arguments stay <= 8 and P fits fp16/bf16.
//
// TMEM columns: S0 [0,128) S1 [128,256) O0 [256,256+D) O1 [256+D,256+2D);  P_t aliases the first 64 columns of S_t at
// D=128 and has columns of its own ([384,512)) at D=64 (kSepP below).
//
// Launch: one-dimensional grid, work order decode_block (fasn_common.cuh): unit-major with a tile-major tail, heavy
// causal blocks first.  GENERIC instantiations carry the dense attn_mask / attn_bias / ALiBi code; a key-padding mask
// (row stride 0) is handled on the fast path.
#include "fasn_common.cuh"
#include "fasn_ptx.cuh"

namespace fasn {

namespace {

// Optional phase timeline (compile with -DFASN_TIMELINE; see fasn_bwd.cu / scripts/timeline.py)
#ifdef FASN_TIMELINE
#define TLF_DECL(role) unsigned long long* tl_p = (a.dbg && tc.tile == (int)a.dbg_x && bh == (int)a.dbg_y) ? a.dbg + (role) * 2048 : nullptr; int tl_i = 0;
#define TLF_ONLY(cond) do { if (!(cond)) tl_p = nullptr; } while (0)
#define TLF(tag) do { if (tl_p && tl_i < 2048) tl_p[tl_i++] = ((unsigned long long)(tag) << 48) | (clock64() & 0xFFFFFFFFFFFFull); } while (0)
#define TLF_CTA_BEGIN() unsigned long long cta_g0 = 0, cta_c0 = 0; if (a.dbg && threadIdx.x == 0) { asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(cta_g0)); cta_c0 = clock64(); }
#define TLF_CTA_END(iters) do { if (a.dbg && threadIdx.x == 0) { unsigned long long g1; unsigned int sm; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g1)); asm volatile("mov.u32 %0, %%smid;" : "=r"(sm)); \
    unsigned long long* rec = a.dbg + 5 * 2048 + 4ull * blockIdx.x; rec[0] = cta_g0; rec[1] = g1; rec[2] = clock64() - cta_c0; rec[3] = ((unsigned long long)sm << 32) | (unsigned int)(iters); } } while (0)
#else
#define TLF_DECL(role)
#define TLF_ONLY(cond)
#define TLF(tag)
#define TLF_CTA_BEGIN()
#define TLF_CTA_END(iters)
#endif

constexpr int kFwdThreads = 384;
// Debug build (-DFASN_DEBUG_FP32_P=1, libfasn_debug32.so): the probabilities reach the tensor core as
// P_hi + P_lo, two 16-bit terms that together carry ~22 mantissa bits (P.V is issued twice per K-step), every exponential runs
// on the MUFU, and the output can be taken in float32 (FasnParams.o_f32) -- the accuracy of the algorithm without the 16-bit
// rounding of P and O.  Checked to <= 1e-5 relative L2 against the float64 oracle (tests/test_gpu_debug32.py).  Forward only.
#ifndef FASN_DEBUG_FP32_P
#define FASN_DEBUG_FP32_P 0
#endif

constexpr float kRescaleThreshold = 8.0f;   // log2 units
// Share of the exponentials evaluated by exp2_poly_pair instead of MUFU.EX2 on unmasked tiles: in kPolyCount of
// every kPolyPeriod groups of four elements, one of the two pairs is a polynomial (1 of 2 -> 25 %).  0 disables.
// Measured on B200 (profiles/README.md): 25 % helps head dim 64 (+4..8 %) and the dropout kernels (+3 %), and costs
// 2 % on the D=128 no-dropout kernel, which therefore keeps every exponential on the MUFU.
#ifndef FASN_POLY_PERIOD_D64
#define FASN_POLY_PERIOD_D64 2
#endif
constexpr int kPolyPeriod = 2;


template <int D> struct FwdCfg {
  static constexpr int NS = (D == 128) ? 2 : 4;          // K/V ring depth
  static constexpr int DB = D / 64;                      // 128-byte blocks per row
  static constexpr int TILE_BYTES = 128 * D * 2;
  static constexpr int BLK_BYTES = 128 * 128;
  static constexpr int SO_TILES = (D == 128) ? 1 : 2;    // output staging tiles: one per Q tile at D=64, shared at D=128 (227 KB limit)
  // q_full[2] s_full[2] p_full[2] o_full[2] p_full2[2] | k_full k_empty v_full v_empty [NS] | q_empty[2] o_drained[2] so_done[2]
  // item_full[2] item_empty[2]
  static constexpr int NUM_BARS = 10 + 4 * NS + 10;
  static constexpr int SMEM_BYTES = 1024 + (2 + 2 * NS + SO_TILES) * TILE_BYTES + NUM_BARS * 8 + 32;
};

// One work item = 256 query rows (two 128-row tiles) of one (batch, head) unit.
struct FwdItem { int q0, bh, b, h, hk, tile, n_tiles, n_tiles0, n_tiles1; };

template <bool CAUSAL> FASN_DEVICE FwdItem fwd_decode(const FwdArgs& a, int lin, int nqb) {
  FwdItem w;
  const TileCoord tc = decode_block((uint32_t)lin, nqb, a.B * a.H, a.sched_group);
  const int qb = CAUSAL ? (nqb - 1 - tc.tile) : tc.tile;   // heavy causal blocks first
  w.tile = tc.tile;
  w.q0 = qb * 256;
  w.bh = tc.bh;
  w.b = tc.bh / a.H;
  w.h = tc.bh - w.b * a.H;
  w.hk = (a.Hkv == 1) ? 0 : w.h;
  // keys visible to this item: [0, kv_end)
  int kv_end = a.Skv;
  if (CAUSAL) kv_end = min(a.Skv, min(w.q0 + 256, a.Sq) + a.causal_off);
  kv_end = max(kv_end, 0);
  w.n_tiles = (kv_end + 127) >> 7;
  // per Q tile: tile 0 usually needs one K/V tile less than tile 1 under a causal mask; tile 1 is skipped
  // entirely when all of its rows lie beyond Sq
  w.n_tiles0 = w.n_tiles;
  if (CAUSAL) w.n_tiles0 = (max(min(a.Skv, min(w.q0 + 128, a.Sq) + a.causal_off), 0) + 127) >> 7;
  w.n_tiles1 = (w.q0 + 128 < a.Sq) ? w.n_tiles : 0;
  return w;
}

// GENERIC = false: the fast path (no dense mask / bias, positive scale; a key-padding mask with row stride 0 is fine).
// GENERIC = true:  dense attn_mask / attn_bias tensors and non-positive scales (scores are scaled before the maximum).
// Two instantiations so that the dense-tensor code (and its registers) stays out of the kernels the headline shapes run.
//
// Persistent: the grid is one CTA per SM; every CTA fetches work items from a global counter (the same heavy-first order
// the hardware block scheduler used to walk, decode_block) until the items run out.  The producer runs ahead of the other
// roles -- K/V tiles of the next item flow through the same ring, its Q tiles are loaded as soon as the last Q.K^T of the
// current item has read them, the output goes through a staging tile of its own -- so launch latency, barrier set-up,
// tensor-memory allocation, the first loads and the epilogue of an item (~10 k cycles per 256-row block, 12 % of the C3
// forward and 30 % of C2 when every block was a CTA of its own; profiles/README.md) overlap the neighbouring items.
template <int D, bool BF16, bool CAUSAL, bool DROPOUT, bool GENERIC>
__global__ void __launch_bounds__(kFwdThreads, 1)
fasn_fwd_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_k,
                const __grid_constant__ CUtensorMap tm_v, const __grid_constant__ CUtensorMap tm_o, const FwdArgs a) {
  using Cfg = FwdCfg<D>;
  constexpr int NS = Cfg::NS, DB = Cfg::DB, TILE_BYTES = Cfg::TILE_BYTES, BLK_BYTES = Cfg::BLK_BYTES;
  // D=128: P is handed to the tensor pipe in two halves (measured +6 % at S=8192); at D=64 the MMAs are too short
  // for the extra barrier round to pay (-3 %), so P is handed over whole.
  constexpr bool kSplitPV = (D == 128);
  // D=64: tensor memory has room for P_t in columns of its own ([384,512)), so S_t is free as soon as the softmax threads
  // have READ it and Q.K^T of the next K/V tile is issued during the softmax of the current one, instead of after its
  // P.V -- at D=64 the MMAs are short and the serial chain Q.K^T -> softmax -> P.V per tile was the limiter.
  // (At D=128 the 512 columns are full: S0 S1 O0 O1.)
  constexpr bool kSepP = (D == 64) && !FASN_DEBUG_FP32_P;   // (the debug build keeps P_lo beside P_hi in the S_t columns)

  TLF_CTA_BEGIN();
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int nqb = (a.Sq + 255) >> 8;
  const int total = nqb * a.B * a.H;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;
  uint8_t* sK = sQ + 2 * TILE_BYTES;
  uint8_t* sV = sK + NS * TILE_BYTES;
  uint8_t* sO = sV + NS * TILE_BYTES;    // [SO_TILES]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sO + Cfg::SO_TILES * TILE_BYTES);
  uint64_t* q_full = bars;               // [2]
  uint64_t* s_full = bars + 2;           // [2]
  uint64_t* p_full = bars + 4;           // [2]
  uint64_t* o_full = bars + 6;           // [2]
  uint64_t* k_full = bars + 8;           // [NS]
  uint64_t* k_empty = k_full + NS;       // [NS]
  uint64_t* v_full = k_empty + NS;       // [NS]
  uint64_t* v_empty = v_full + NS;       // [NS]
  uint64_t* p_full2 = v_empty + NS;      // [2]  second half (keys 64..127) of P_t written
  uint64_t* q_empty = p_full2 + 2;       // [2]  the last Q.K^T of the item has read Q_t
  uint64_t* o_drained = q_empty + 2;     // [2]  128 arrivals: the epilogue has read O_t out of tensor memory
  uint64_t* so_done = o_drained + 2;     // [2]  the TMA store of an epilogue has read its staging tile
  uint64_t* item_full = so_done + 2;     // [2]  work-item ring: the producer has published an item index
  uint64_t* item_empty = item_full + 2;  // [2]  9 arrivals: the MMA warp and the eight softmax warps have read it
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + Cfg::NUM_BARS);
  volatile int* sItem = reinterpret_cast<volatile int*>(tmem_slot + 2);   // [2]

  if (warp == 8 && lane == 0) {
    tma_prefetch_desc(&tm_q); tma_prefetch_desc(&tm_k); tma_prefetch_desc(&tm_v); tma_prefetch_desc(&tm_o);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&q_full[i], 1); mbar_init(&s_full[i], 1); mbar_init(&p_full[i], 128); mbar_init(&p_full2[i], 128); mbar_init(&o_full[i], 1);
      mbar_init(&q_empty[i], 1); mbar_init(&o_drained[i], 128); mbar_init(&so_done[i], 1); mbar_init(&item_full[i], 1); mbar_init(&item_empty[i], 9);
    }
    for (int i = 0; i < NS; ++i) { mbar_init(&k_full[i], 1); mbar_init(&k_empty[i], 1); mbar_init(&v_full[i], 1); mbar_init(&v_empty[i], 1); }
    fence_mbar_init();
    fence_proxy_async_smem();
  }
  if (warp == 10) { tmem_alloc<512>(tmem_slot); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  uint32_t iters_done = 0;               // timeline builds: K/V steps of both tiles over all items of this CTA
  if (warp >= 8) {
    setmaxnreg_dec<48>();   // 256 x 224 + 128 x 48 = 63488 <= 384 x 168 registers granted at launch
    if (warp == 8) {
      // ------------------------------------------------------------------ TMA producer
      if (lane == 0) {
        int lin = blockIdx.x;
        uint32_t item_n = 0, g = 0, quses = 0;     // items published, K/V tiles loaded, items that loaded Q
        while (true) {
          const int si = item_n & 1;
          if (item_n >= 2) mbar_wait(&item_empty[si], ((item_n >> 1) - 1) & 1);
          sItem[si] = (lin < total) ? lin : -1;
          mbar_arrive(&item_full[si]);             // release: the index is visible to whoever observes the phase
          if (lin >= total) break;
          const FwdItem w = fwd_decode<CAUSAL>(a, lin, nqb);
          if (w.n_tiles > 0) {
#pragma unroll
            for (int t = 0; t < 2; ++t) {
              if (quses > 0) mbar_wait(&q_empty[t], (quses - 1) & 1);
              mbar_arrive_expect_tx(&q_full[t], TILE_BYTES);
#pragma unroll
              for (int db = 0; db < DB; ++db)
                tma_load_4d(sQ + t * TILE_BYTES + db * BLK_BYTES, &tm_q, &q_full[t], db * 64, w.q0 + t * 128, w.h, w.b);
            }
            for (int j = 0; j < w.n_tiles; ++j, ++g) {
              const int s = g % NS;
              const uint32_t ph = (g / NS) & 1;
              mbar_wait(&k_empty[s], ph ^ 1);
              mbar_arrive_expect_tx(&k_full[s], TILE_BYTES);
#pragma unroll
              for (int db = 0; db < DB; ++db)
                tma_load_4d(sK + s * TILE_BYTES + db * BLK_BYTES, &tm_k, &k_full[s], db * 64, j * 128, w.hk, w.b);
              mbar_wait(&v_empty[s], ph ^ 1);
              mbar_arrive_expect_tx(&v_full[s], TILE_BYTES);
#pragma unroll
              for (int db = 0; db < DB; ++db)
                tma_load_4d(sV + s * TILE_BYTES + db * BLK_BYTES, &tm_v, &v_full[s], db * 64, j * 128, w.hk, w.b);
            }
            ++quses;
          }
          ++item_n;
          lin = (int)gridDim.x + atomicAdd(a.sched, 1);
        }
      }
    } else if (warp == 9) {
      // ------------------------------------------------------------------ MMA issuer
      // The whole warp walks the schedule (so every operand stays in uniform registers); one elected lane issues.
      // Descriptors are base words computed once plus compile-time offsets: the issue loop is a handful of
      // instructions per tcgen05.mma, well under the 64 cycles each 128x128x16 MMA occupies the tensor pipe.
      constexpr uint32_t idesc_qk = umma_idesc(BF16, 128, 128, false, false);
      constexpr uint32_t idesc_pv = umma_idesc(BF16, 128, D, false, true);
      constexpr uint32_t hi_desc = umma_desc_hi(1024);
      const uint32_t tm = __shfl_sync(0xffffffffu, tmem_base, 0);
      const uint32_t q_lo = umma_desc_lo(smem_u32(sQ), 16);            // K-major: LBO unused
      const uint32_t k_lo = umma_desc_lo(smem_u32(sK), 16);
      const uint32_t v_lo = umma_desc_lo(smem_u32(sV), BLK_BYTES);     // MN-major: LBO = next 64-wide block of D
      auto issue_qk = [&](int t, int s) {
        const uint32_t a0 = q_lo + t * (TILE_BYTES >> 4), b0 = k_lo + s * (TILE_BYTES >> 4);
#pragma unroll
        for (int kb = 0; kb < D / 16; ++kb) {
          const uint32_t off = ((kb >> 2) * BLK_BYTES + (kb & 3) * 32) >> 4;
          umma_ss(tm + t * 128, umma_desc_join(a0 + off, hi_desc), umma_desc_join(b0 + off, hi_desc), idesc_qk, kb > 0 ? 1u : 0u);
        }
      };
      // P.V is issued in two halves (keys 0..63, then 64..127 of the tile) so the first half runs on the tensor pipe
      // while the softmax warps are still producing the second half of P
      auto issue_pv_half = [&](int t, int s, int half, uint32_t acc) {
        const uint32_t b0 = v_lo + s * (TILE_BYTES >> 4);
#pragma unroll
        for (int kb = 0; kb < 4; ++kb) {
          const int k = half * 4 + kb;
          umma_ts(tm + 256 + t * D, tm + (kSepP ? 384 + t * 64 : t * 128) + k * 8, umma_desc_join(b0 + k * (2048 >> 4), hi_desc), idesc_pv,
                  (half > 0 || kb > 0) ? 1u : acc);
#if FASN_DEBUG_FP32_P
          umma_ts(tm + 256 + t * D, tm + t * 128 + 64 + k * 8, umma_desc_join(b0 + k * (2048 >> 4), hi_desc), idesc_pv, 1u);   // P_lo
#endif
        }
      };
      // counters that run across work items: items seen, K/V tiles consumed, items with Q loads, per Q tile: steps done and
      // items that accumulated into O_t (its epilogue arrives on o_drained[t] once per such item)
      uint32_t item_n = 0, g = 0, quses = 0, pc[2] = {0, 0}, ouses[2] = {0, 0};
      while (true) {
        const int si = item_n & 1;
        mbar_wait(&item_full[si], (item_n >> 1) & 1);
        const int lin = sItem[si];
        __syncwarp();
        if (lane == 0) mbar_arrive(&item_empty[si]);
        if (lin < 0) break;
        ++item_n;
        const FwdItem w = fwd_decode<CAUSAL>(a, lin, nqb);
        const int n_tiles = w.n_tiles, n_tiles0 = w.n_tiles0, n_tiles1 = w.n_tiles1;
        if (n_tiles == 0) continue;
        const int n_tl[2] = {n_tiles0, n_tiles1};
#ifdef FASN_TIMELINE
        const FwdItem& tc = w; const int bh = w.bh;
#endif
        TLF_DECL(0)
        TLF_ONLY(lane == 0);
        mbar_wait(&q_full[0], quses & 1);
        mbar_wait(&q_full[1], quses & 1);
        mbar_wait(&k_full[g % NS], (g / NS) & 1);
        tc_fence_after();
        TLF(1);
        if (elect_one()) {
          const int s0 = g % NS;
          // S_t of the previous item is dead: its last P.V was issued before these MMAs (same pipe, program order)
          if (n_tiles0 > 0) { issue_qk(0, s0); tc_commit(&s_full[0]); if (n_tiles0 == 1) tc_commit(&q_empty[0]); }
          if (n_tiles1 > 0) { issue_qk(1, s0); tc_commit(&s_full[1]); if (n_tiles1 == 1) tc_commit(&q_empty[1]); }
          tc_commit(&k_empty[s0]);
          // a Q tile without visible keys is never read: hand its buffer back at once
          if (n_tiles0 == 0) mbar_arrive(&q_empty[0]);
          if (n_tiles1 == 0) mbar_arrive(&q_empty[1]);
        }
        __syncwarp();
        for (int j = 0; j < n_tiles; ++j) {
          const uint32_t gj = g + j;
          const int s = gj % NS;
          const uint32_t ph = (gj / NS) & 1;
          const int s1 = (gj + 1) % NS;
          const uint32_t ph1 = ((gj + 1) / NS) & 1;
          const bool more = (j + 1 < n_tiles);
          mbar_wait(&v_full[s], ph);
          if (more) mbar_wait(&k_full[s1], ph1);
          if (kSepP && more) {
            // Q.K^T of the next K/V tile as soon as the softmax threads have read S_t (p_full2 = "S_t has been read")
#pragma unroll
            for (int t = 0; t < 2; ++t) {
              if (j + 1 < n_tl[t]) {
                mbar_wait(&p_full2[t], (pc[t] + j) & 1);
                tc_fence_after();
                if (elect_one()) { issue_qk(t, s1); tc_commit(&s_full[t]); if (j + 2 == n_tl[t]) tc_commit(&q_empty[t]); }
                __syncwarp();
              }
            }
            if (elect_one()) tc_commit(&k_empty[s1]);
            __syncwarp();
          }
#pragma unroll
          for (int t = 0; t < 2; ++t) {
            if (j < n_tl[t]) {
              mbar_wait(&p_full[t], (pc[t] + j) & 1);
              // the first P.V of an item overwrites O_t: the previous item's epilogue must have read it
              if (j == 0 && ouses[t] > 0) mbar_wait(&o_drained[t], (ouses[t] - 1) & 1);
              tc_fence_after();
              TLF(2 + t);
              if (kSplitPV) {
                if (elect_one()) issue_pv_half(t, s, 0, j > 0 ? 1u : 0u);
                __syncwarp();
                mbar_wait(&p_full2[t], (pc[t] + j) & 1);
                tc_fence_after();
              }
              if (elect_one()) {
                if (!kSplitPV) issue_pv_half(t, s, 0, j > 0 ? 1u : 0u);
                issue_pv_half(t, s, 1, 1u);
                tc_commit(&o_full[t]);
                if (!kSepP && j + 1 < n_tl[t]) { issue_qk(t, s1); tc_commit(&s_full[t]); if (j + 2 == n_tl[t]) tc_commit(&q_empty[t]); }
              }
              __syncwarp();
            }
          }
          if (elect_one()) {
            tc_commit(&v_empty[s]);
            if (!kSepP && more) tc_commit(&k_empty[s1]);
          }
          __syncwarp();
          TLF(4);
        }
        g += n_tiles; pc[0] += n_tiles0; pc[1] += n_tiles1; ++quses;
        if (n_tiles0 > 0) ++ouses[0];
        if (n_tiles1 > 0) ++ouses[1];
      }
    }
  } else {
    // -------------------------------------------------------------------- softmax warpgroups
    setmaxnreg_inc<224>();
    const int t = warp >> 2;
    const int r = threadIdx.x & 127;
    const uint32_t lane_off = static_cast<uint32_t>((warp & 3) * 32) << 16;
    const uint32_t tS = tmem_base + lane_off + t * 128;
    const uint32_t tO = tmem_base + lane_off + 256 + t * D;
    const bool has_aux = (a.mask.ptr != nullptr) || (a.bias.ptr != nullptr);
    // Fast path: the logit scale is folded into the exponent FFMA (p = 2^(s*c - m)) and the running max is taken on
    // the raw scores (valid for c > 0).  Generic path (mask / bias / non-positive scale): scores are scaled first.
    constexpr bool generic = GENERIC;
    const float cmul = generic ? 1.f : a.scale_log2;
    const float2 cmul2 = make_float2(cmul, cmul);
    // output staging tile and its hand-over barrier: per Q tile at D=64, one shared by both tiles at D=128
    uint8_t* const sOt = sO + (Cfg::SO_TILES == 2 ? t : 0) * TILE_BYTES;
    uint64_t* const so_bar = &so_done[Cfg::SO_TILES == 2 ? t : 0];
    uint32_t item_n = 0, sc = 0, kuses = 0;   // items seen, K/V steps of this Q tile so far, items that went through the epilogue

    while (true) {
      const int si = item_n & 1;
      mbar_wait(&item_full[si], (item_n >> 1) & 1);
      const int lin = sItem[si];
      __syncwarp();
      if (lane == 0) mbar_arrive(&item_empty[si]);
      if (lin < 0) break;
      ++item_n;
      const FwdItem w = fwd_decode<CAUSAL>(a, lin, nqb);
      const int q0 = w.q0, bh = w.bh, b = w.b, h = w.h;
      const int row = q0 + t * 128 + r;
      if (w.n_tiles == 0) {
        // No key is visible to any row of this block: softmax_n gives exactly 0 (n > 0), defined 0 for n == 0.
        if (row < a.Sq) {
          uint4* dst = reinterpret_cast<uint4*>(reinterpret_cast<uint16_t*>(a.o.ptr) + b * a.o.sb + h * a.o.sh + (long long)row * a.o.ss);
#pragma unroll
          for (int i = 0; i < D / 8; ++i) dst[i] = make_uint4(0, 0, 0, 0);
          a.lse[(long long)bh * a.Sq + row] = (a.softmax_n > 0.f) ? logf(a.softmax_n) : INFINITY;
        }
        continue;
      }
#ifdef FASN_TIMELINE
      const FwdItem& tc = w;
#endif
      const int row_lim = CAUSAL ? min(a.Skv, row + a.causal_off + 1) : a.Skv;      // visible keys: [0,row_lim)
      const int warp_row_lim = __shfl_sync(0xffffffffu, row_lim, 0);                // smallest in the warp
      const int row_c = min(row, a.Sq - 1);
      const uint8_t* mrow = a.mask.ptr ? reinterpret_cast<const uint8_t*>(a.mask.ptr) + b * a.mask.sb + h * a.mask.sh +
                                             (long long)row_c * a.mask.sq
                                       : nullptr;
      const uint16_t* brow = a.bias.ptr ? reinterpret_cast<const uint16_t*>(a.bias.ptr) + b * a.bias.sb + h * a.bias.sh +
                                              (long long)row_c * a.bias.sq
                                        : nullptr;
      const uint32_t bh_global = a.bh_offset + bh;
      const int n_t = (t == 0) ? w.n_tiles0 : w.n_tiles1;      // K/V tiles this Q tile needs
      // A mask broadcast over the query axis (key padding, row stride 0) without bias stays on the fast path: the 128 mask
      // bytes of a K/V tile are the same for every row, so each warp turns them into four 32-bit visibility words with
      // ballots and only tiles that contain a hidden key pay for the selects.
      const bool key_only_mask = !GENERIC && (mrow != nullptr);      // (the host picks GENERIC unless row stride 0, no bias, scale > 0)
      const float alibi2 = (GENERIC && a.alibi != nullptr) ? a.alibi[h] * kLog2e : 0.f;
      float m = (a.softmax_n > 0.f) ? 0.f : -INFINITY;   // running reference max (log2 domain)
      float l = a.softmax_n;                             // running sum, starts at n (the virtual zero-logit key)

      TLF_DECL(1 + t)
      TLF_ONLY(r == 0);
      for (int j = 0; j < n_t; ++j) {
        const int j0 = j * 128;
        TLF(10);
        uint32_t kw[4];
        if constexpr (DROPOUT) {   // independent of S: issue before waiting for the tensor core
#pragma unroll
          for (int i = 0; i < 4; ++i) kw[i] = dropout_keep_word(a.key, bh_global, (uint32_t)row, (uint32_t)(j0 >> 5) + i, a.drop_thr);
        }
        uint32_t vis[4] = {0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu};
        if (key_only_mask) {
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int col = j0 + 32 * i + lane;
            const uint8_t mb = (col < a.Skv) ? __ldg(mrow + col) : (uint8_t)1;      // keys beyond Skv are cut by row_lim below
            vis[i] = __ballot_sync(0xffffffffu, mb != 0);
          }
        }
        const bool hidden_keys = (vis[0] & vis[1] & vis[2] & vis[3]) != 0xFFFFFFFFu;   // warp-uniform
        // Dense bias (GENERIC kernels): this thread's 128 elements of the tile are sixteen 16-byte loads.  On interior tiles
        // with an aligned row they are issued here, unconditionally and before the wait for the tensor core, so their
        // latency overlaps it (conditional loads serialise: one DRAM round trip each).
        const bool bias_fast = GENERIC && brow != nullptr && (j0 + 128 <= a.Skv) && ((reinterpret_cast<uintptr_t>(brow) & 15) == 0);
        const bool mask_fast = GENERIC && mrow != nullptr && (j0 + 128 <= a.Skv) && ((reinterpret_cast<uintptr_t>(mrow) & 15) == 0);
        uint4 bq[GENERIC ? 16 : 1];
        if constexpr (GENERIC) {
          if (bias_fast) {
#pragma unroll
            for (int g = 0; g < 16; ++g) bq[g] = __ldg(reinterpret_cast<const uint4*>(brow + j0) + g);
          }
        }
        mbar_wait(&s_full[t], (sc + j) & 1);
        tc_fence_after();
        float s[128];
        {
          uint32_t* sr = reinterpret_cast<uint32_t*>(s);
          tmem_ld_x32(tS + 0, sr + 0);
          tmem_ld_x32(tS + 32, sr + 32);
          tmem_ld_x32(tS + 64, sr + 64);
          tmem_ld_x32(tS + 96, sr + 96);
          TLF(11);
          tmem_wait_ld();
          TLF(12);
          if (kSepP) { tc_fence_before(); mbar_arrive(&p_full2[t]); }      // S_t may be overwritten by the next Q.K^T
        }
        if constexpr (GENERIC) {
          if (a.alibi != nullptr) {       // ALiBi generated in place: + slope (j - i - (S - L)), log2 domain
            const float base = alibi2 * (float)(j0 - row - a.causal_off);
#pragma unroll
            for (int c = 0; c < 128; ++c) s[c] = fmaf(s[c], a.scale_log2, fmaf(alibi2, (float)c, base));
          } else {
#pragma unroll
            for (int c = 0; c < 128; ++c) s[c] *= a.scale_log2;
          }
          if (has_aux) {
            // dense bias / mask rows of this thread: 16-byte loads where the row segment is aligned and in range
            // (each thread streams its own 256 B / 128 B per tile; lines are shared by consecutive instructions via L1)
            if (bias_fast) {
#pragma unroll
              for (int g = 0; g < 16; ++g) {
                const uint32_t w[4] = {bq[g].x, bq[g].y, bq[g].z, bq[g].w};
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                  s[g * 8 + 2 * e] = fmaf(cvt16_to_f32<BF16>(w[e] & 0xFFFF), kLog2e, s[g * 8 + 2 * e]);
                  s[g * 8 + 2 * e + 1] = fmaf(cvt16_to_f32<BF16>(w[e] >> 16), kLog2e, s[g * 8 + 2 * e + 1]);
                }
              }
            } else if (brow) {
#pragma unroll
              for (int g = 0; g < 16; ++g) {                    // 8 bias elements per 16-byte load
                const int col = j0 + g * 8;
                const uint16_t* p = brow + col;
                if (col + 8 <= a.Skv && (reinterpret_cast<uintptr_t>(p) & 15) == 0) {
                  const uint4 v4 = __ldg(reinterpret_cast<const uint4*>(p));
                  const uint32_t w[4] = {v4.x, v4.y, v4.z, v4.w};
#pragma unroll
                  for (int e = 0; e < 4; ++e) {
                    s[g * 8 + 2 * e] = fmaf(cvt16_to_f32<BF16>(w[e] & 0xFFFF), kLog2e, s[g * 8 + 2 * e]);
                    s[g * 8 + 2 * e + 1] = fmaf(cvt16_to_f32<BF16>(w[e] >> 16), kLog2e, s[g * 8 + 2 * e + 1]);
                  }
                } else {
#pragma unroll
                  for (int e = 0; e < 8; ++e)
                    if (col + e < a.Skv) s[g * 8 + e] = fmaf(cvt16_to_f32<BF16>(p[e]), kLog2e, s[g * 8 + e]);
                }
              }
            }
            if (mask_fast) {
              uint4 mq[8];
#pragma unroll
              for (int g = 0; g < 8; ++g) mq[g] = __ldg(reinterpret_cast<const uint4*>(mrow + j0) + g);
#pragma unroll
              for (int g = 0; g < 8; ++g) {
                const uint32_t w[4] = {mq[g].x, mq[g].y, mq[g].z, mq[g].w};
#pragma unroll
                for (int e = 0; e < 16; ++e)
                  if (((w[e >> 2] >> (8 * (e & 3))) & 0xFF) == 0) s[g * 16 + e] = -INFINITY;
              }
            } else if (mrow) {
#pragma unroll
              for (int g = 0; g < 8; ++g) {                     // 16 mask bytes per 16-byte load
                const int col = j0 + g * 16;
                const uint8_t* p = mrow + col;
                if (col + 16 <= a.Skv && (reinterpret_cast<uintptr_t>(p) & 15) == 0) {
                  const uint4 v4 = __ldg(reinterpret_cast<const uint4*>(p));
                  const uint32_t w[4] = {v4.x, v4.y, v4.z, v4.w};
#pragma unroll
                  for (int e = 0; e < 16; ++e)
                    if (((w[e >> 2] >> (8 * (e & 3))) & 0xFF) == 0) s[g * 16 + e] = -INFINITY;
                } else {
#pragma unroll
                  for (int e = 0; e < 16; ++e)
                    if (col + e < a.Skv && p[e] == 0) s[g * 16 + e] = -INFINITY;
                }
              }
            }
          }
        }
        const bool masked_tile = (j0 + 128 > warp_row_lim) || hidden_keys;      // warp-uniform
        if (j0 + 128 > warp_row_lim) {
          const int lim = row_lim - j0;
#pragma unroll
          for (int c = 0; c < 128; ++c) s[c] = (c < lim) ? s[c] : -INFINITY;
        }
        if (hidden_keys) {
#pragma unroll
          for (int c = 0; c < 128; ++c) s[c] = ((vis[c >> 5] >> (c & 31)) & 1u) ? s[c] : -INFINITY;
        }
        float mx0 = s[0], mx1 = s[1], mx2 = s[2], mx3 = s[3];
#pragma unroll
        for (int c = 4; c < 124; c += 8) {
          mx0 = fmax3(mx0, s[c], s[c + 1]); mx1 = fmax3(mx1, s[c + 2], s[c + 3]);
          mx2 = fmax3(mx2, s[c + 4], s[c + 5]); mx3 = fmax3(mx3, s[c + 6], s[c + 7]);
        }
        mx0 = fmax3(mx0, s[124], s[125]); mx1 = fmax3(mx1, s[126], s[127]);
        const float tmax = fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3)) * cmul;
        const bool upd = tmax > m + kRescaleThreshold;
        float alpha = 1.f;
        if (upd) {
          alpha = ex2(m - tmax);
          m = tmax;
          l *= alpha;
        }
        if (j > 0 && __any_sync(0xffffffffu, upd)) {
          // rescale the O accumulator of this tile (rare): PV_{j-1} must have landed first
          mbar_wait(&o_full[t], (sc + j - 1) & 1);
          tc_fence_after();
#pragma unroll
          for (int cb = 0; cb < D / 32; ++cb) {
            uint32_t o[32];
            tmem_ld_x32(tO + cb * 32, o);
            tmem_wait_ld();
#pragma unroll
            for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
            tmem_st_x32(tO + cb * 32, o);
          }
        }
        TLF(13);
        const float m_use = (m == -INFINITY) ? 0.f : m;
        const float2 negm2 = make_float2(-m_use, -m_use);
        uint32_t pr[64];
#if FASN_DEBUG_FP32_P
        uint32_t pr_lo[64];
#endif
        float2 l01 = make_float2(0.f, 0.f), l23 = make_float2(0.f, 0.f);
        auto finish4 = [&](int c, float p0, float p1, float p2, float p3) {
          l01 = __fadd2_rn(l01, make_float2(p0, p1));
          l23 = __fadd2_rn(l23, make_float2(p2, p3));
          uint32_t w01 = pack2<BF16>(p0, p1), w23 = pack2<BF16>(p2, p3);
#if FASN_DEBUG_FP32_P
          uint32_t r01 = pack2<BF16>(p0 - cvt16_to_f32<BF16>((uint16_t)(w01 & 0xFFFF)), p1 - cvt16_to_f32<BF16>((uint16_t)(w01 >> 16)));
          uint32_t r23 = pack2<BF16>(p2 - cvt16_to_f32<BF16>((uint16_t)(w23 & 0xFFFF)), p3 - cvt16_to_f32<BF16>((uint16_t)(w23 >> 16)));
#endif
          if constexpr (DROPOUT) {     // zero the dropped entries on the packed pairs: 1 PRMT + 1 AND per two elements
            const uint32_t w = kw[c >> 5];
            w01 &= keep_pair_mask(w, c & 31);
            w23 &= keep_pair_mask(w, (c + 2) & 31);
#if FASN_DEBUG_FP32_P
            r01 &= keep_pair_mask(w, c & 31);
            r23 &= keep_pair_mask(w, (c + 2) & 31);
#endif
          }
          pr[c >> 1] = w01;
          pr[(c >> 1) + 1] = w23;
#if FASN_DEBUG_FP32_P
          pr_lo[c >> 1] = r01;
          pr_lo[(c >> 1) + 1] = r23;
#endif
        };
#ifndef FASN_POLY_DROPOUT
#define FASN_POLY_DROPOUT 1
#endif
        constexpr int kPolyCount = (!FASN_DEBUG_FP32_P && (D == 64 || (DROPOUT && FASN_POLY_DROPOUT))) ? 1 : 0;
        const bool use_poly = kPolyCount > 0 && !generic && !masked_tile;
#pragma unroll
        for (int hf = 0; hf < 2; ++hf) {
          if (use_poly) {
            // interior tile, all scores finite: a fixed share of the exponentials runs as a polynomial on the FMA pipes
#pragma unroll
            for (int c = hf * 64; c < hf * 64 + 64; c += 4) {
              const float2 a01 = __ffma2_rn(make_float2(s[c], s[c + 1]), cmul2, negm2);
              const float2 a23 = __ffma2_rn(make_float2(s[c + 2], s[c + 3]), cmul2, negm2);
              float p0 = ex2(a01.x), p1 = ex2(a01.y), p2, p3;
              if (((c >> 2) % (D == 64 ? FASN_POLY_PERIOD_D64 : kPolyPeriod)) < kPolyCount) { const float2 e = exp2_poly_pair(a23); p2 = e.x; p3 = e.y; }
              else { p2 = ex2(a23.x); p3 = ex2(a23.y); }
              finish4(c, p0, p1, p2, p3);
            }
          } else {
#pragma unroll
            for (int c = hf * 64; c < hf * 64 + 64; c += 4) {
              const float2 a01 = __ffma2_rn(make_float2(s[c], s[c + 1]), cmul2, negm2);
              const float2 a23 = __ffma2_rn(make_float2(s[c + 2], s[c + 3]), cmul2, negm2);
              finish4(c, ex2(a01.x), ex2(a01.y), ex2(a23.x), ex2(a23.y));
            }
          }
          // P columns [32 hf, 32 hf + 32) <- keys [64 hf, 64 hf + 64): over S columns this thread has already read (D=128),
          // or in P_t's own columns once the previous P.V has consumed them (D=64)
          if (kSepP && hf == 0 && j > 0) { mbar_wait(&o_full[t], (sc + j - 1) & 1); tc_fence_after(); }
          tmem_st_x32((kSepP ? tmem_base + lane_off + 384 + t * 64 : tS) + hf * 32, pr + hf * 32);
#if FASN_DEBUG_FP32_P
          tmem_st_x32(tS + 64 + hf * 32, pr_lo + hf * 32);      // S_t columns [64,128) were read into registers above
#endif
          if (kSplitPV || hf == 1) {
            tmem_wait_st();
            tc_fence_before();
            mbar_arrive((kSplitPV && hf == 1) ? &p_full2[t] : &p_full[t]);
          }
          if (hf == 0) TLF(14);
        }
        l += (l01.x + l01.y) + (l23.x + l23.y);
        TLF(15);
      }

      // ---------------------------------------------------------------- epilogue
      if (n_t > 0) {
        mbar_wait(&o_full[t], (sc + n_t - 1) & 1);
        tc_fence_after();
      }
      // the staging tile is free once the previous epilogue's TMA store has read it (uses alternate t = 0, 1, 0, ... when
      // the two Q tiles share one tile)
      const uint32_t use = (Cfg::SO_TILES == 2) ? kuses : 2 * kuses + t;
      if (use > 0) mbar_wait(so_bar, (use - 1) & 1);
      const float inv = (l > 0.f && n_t > 0) ? (DROPOUT ? a.inv_keep : 1.f) / l : 0.f;
#pragma unroll
      for (int cb = 0; cb < D / 32; ++cb) {
        uint32_t o[32];
        if (n_t > 0) {
          tmem_ld_x32(tO + cb * 32, o);
          tmem_wait_ld();
          if (cb == D / 32 - 1) { tc_fence_before(); mbar_arrive(&o_drained[t]); }   // O_t may be overwritten by the next item
        } else {                        // no visible key for this Q tile: the accumulator was never written
#pragma unroll
          for (int i = 0; i < 32; ++i) o[i] = 0u;
        }
#if FASN_DEBUG_FP32_P
        if (a.o_f32 != nullptr && row < a.Sq) {
          float4* dst = reinterpret_cast<float4*>(a.o_f32 + ((long long)bh * a.Sq + row) * D + cb * 32);
#pragma unroll
          for (int i = 0; i < 8; ++i)
            dst[i] = make_float4(__uint_as_float(o[4 * i]) * inv, __uint_as_float(o[4 * i + 1]) * inv, __uint_as_float(o[4 * i + 2]) * inv,
                                 __uint_as_float(o[4 * i + 3]) * inv);
        }
#endif
#pragma unroll
        for (int g4 = 0; g4 < 4; ++g4) {   // 4 x 16-byte chunks (8 elements each)
          uint4 v;
          v.x = pack2<BF16>(__uint_as_float(o[g4 * 8 + 0]) * inv, __uint_as_float(o[g4 * 8 + 1]) * inv);
          v.y = pack2<BF16>(__uint_as_float(o[g4 * 8 + 2]) * inv, __uint_as_float(o[g4 * 8 + 3]) * inv);
          v.z = pack2<BF16>(__uint_as_float(o[g4 * 8 + 4]) * inv, __uint_as_float(o[g4 * 8 + 5]) * inv);
          v.w = pack2<BF16>(__uint_as_float(o[g4 * 8 + 6]) * inv, __uint_as_float(o[g4 * 8 + 7]) * inv);
          const int col = cb * 32 + g4 * 8;
          const int db = col >> 6;
          const int cc = (col & 63) >> 3;
          *reinterpret_cast<uint4*>(sOt + db * BLK_BYTES + r * 128 + ((cc ^ (r & 7)) << 4)) = v;
        }
      }
      if (row < a.Sq) a.lse[(long long)bh * a.Sq + row] = (l > 0.f) ? (m + log2f(l)) * kLn2 : INFINITY;
      fence_proxy_async_smem();
      named_bar_sync(1 + t, 128);
      if (r == 0) {
#pragma unroll
        for (int db = 0; db < DB; ++db) tma_store_4d(&tm_o, sOt + db * BLK_BYTES, db * 64, q0 + t * 128, h, b);
        tma_store_commit();
        tma_store_wait_read_all();     // the staging tile has been read; the global writes complete on their own
        mbar_arrive(so_bar);
      }
      sc += n_t;
      ++kuses;
      iters_done += n_t;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 10) tmem_dealloc<512>(tmem_base);
  // the last CTA to finish hands the work counter back at zero for the next launch that uses this slot
  if (threadIdx.x == 0 && atomicAdd(a.sched + 1, 1) == (int)gridDim.x - 1) { a.sched[0] = 0; a.sched[1] = 0; __threadfence(); }
  TLF_CTA_END(iters_done);
}

}  // namespace

template <int D, bool BF16, bool CAUSAL, bool DROPOUT, bool GENERIC>
static cudaError_t launch_fwd_t2(const CUtensorMap& tq, const CUtensorMap& tk, const CUtensorMap& tv, const CUtensorMap& to,
                                const FwdArgs& a, cudaStream_t stream) {
  auto kern = fasn_fwd_kernel<D, BF16, CAUSAL, DROPOUT, GENERIC>;
  constexpr int smem = FwdCfg<D>::SMEM_BYTES;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  if (e != cudaSuccess) return e;
  dim3 grid(a.grid_ctas, 1, 1);      // persistent: one CTA per SM (or per work item when there are fewer)
  kern<<<grid, kFwdThreads, smem, stream>>>(tq, tk, tv, to, a);
  return cudaGetLastError();
}

template <int D, bool BF16, bool CAUSAL, bool DROPOUT>
static cudaError_t launch_fwd_t(const CUtensorMap& tq, const CUtensorMap& tk, const CUtensorMap& tv, const CUtensorMap& to,
                                const FwdArgs& a, cudaStream_t stream) {
  const bool generic = a.bias.ptr != nullptr || a.alibi != nullptr || (a.mask.ptr != nullptr && a.mask.sq != 0) || !(a.scale_log2 > 0.f);
  return generic ? launch_fwd_t2<D, BF16, CAUSAL, DROPOUT, true>(tq, tk, tv, to, a, stream)
                 : launch_fwd_t2<D, BF16, CAUSAL, DROPOUT, false>(tq, tk, tv, to, a, stream);
}

cudaError_t launch_fwd(int head_dim, bool bf16, bool causal, bool dropout, const CUtensorMap& tq, const CUtensorMap& tk,
                       const CUtensorMap& tv, const CUtensorMap& to, const FwdArgs& a, cudaStream_t stream) {
#define FASN_FWD_CASE(D_, BF_, C_, DR_) \
  if (head_dim == D_ && bf16 == BF_ && causal == C_ && dropout == DR_) return launch_fwd_t<D_, BF_, C_, DR_>(tq, tk, tv, to, a, stream);
  FASN_FWD_CASE(64, false, false, false) FASN_FWD_CASE(64, false, false, true)
  FASN_FWD_CASE(64, false, true, false)  FASN_FWD_CASE(64, false, true, true)
  FASN_FWD_CASE(64, true, false, false)  FASN_FWD_CASE(64, true, false, true)
  FASN_FWD_CASE(64, true, true, false)   FASN_FWD_CASE(64, true, true, true)
  FASN_FWD_CASE(128, false, false, false) FASN_FWD_CASE(128, false, false, true)
  FASN_FWD_CASE(128, false, true, false)  FASN_FWD_CASE(128, false, true, true)
  FASN_FWD_CASE(128, true, false, false)  FASN_FWD_CASE(128, true, false, true)
  FASN_FWD_CASE(128, true, true, false)   FASN_FWD_CASE(128, true, true, true)
#undef FASN_FWD_CASE
  return cudaErrorInvalidValue;
}

}  // namespace fasn
