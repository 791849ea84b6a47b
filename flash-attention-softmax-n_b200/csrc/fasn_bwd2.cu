// Paired backward kernel (head dim 128, no dense mask / bias): two CTAs on neighbouring SMs (a cluster of 2) own two
// adjacent 128-row K/V tiles of one (batch, head) unit and issue every tensor-core operation as ONE tcgen05
// `cta_group::2` MMA spanning both SMs.
//
// Why: the single-CTA kernel (fasn_bwd.cu) is bound by shared-memory bandwidth, not by the tensor pipe -- a 128x128x16
// MMA with both operands in shared memory reads 8 KB per 64 cycles, the whole 128 B/clk of an SM, and the Q / dO tile
// loads, the dS tile and the fp32 dQ staging compete for the same port (~512 KB of traffic per tile pair against 2560
// tensor-pipe cycles; profiles/README.md).  In a pair the B operand of every MMA is split between the two SMs, the
// probabilities AND dS^T are fed from tensor memory, and each SM stages only half of dQ:
//
//   S^T  = K Q_i^T     M=256 (128 kv rows per CTA) N=128   A = own K tile (smem), B = 64 query rows of Q_i per CTA
//   dP^T = V dO_i^T    same with V, dO_i
//   dV  += P^T dO_i    A = P^T in each CTA's TMEM, B = dO_i[:, 64 c .. 64 c + 63] (CTA c holds one half of the head dim)
//   dK  += dS^T Q_i    A = dS^T in TMEM (16-bit, beside dQ in the columns dP^T occupied), B = Q_i[:, 64 c ..]
//   dQ_i = dS K        M=128 (64 query rows per CTA), K = 256 kv rows of the pair, N=128: A = dS[q half c, all 256 kv]
//                      -- the half computed by the peer CTA arrives through distributed shared memory --,
//                      B = K[all 256 kv, 64 c ..];  dQ lands as 64 rows x 128 columns per CTA (columns 0-63 on TMEM
//                      lanes 0-63, 64-127 on lanes 64-127) and is reduced into dq_accum by TMA reduce-add.
//
// Per CTA and Q tile this moves ~336 KB through shared memory instead of ~512 KB, and halves the dQ reduce traffic.
//
// Only CTA 0's warp 13 issues MMAs.  Operand loads of both CTAs complete on CTA 0's mbarriers (TMA .cta_group::2),
// tcgen05.commit multicasts "done" to the same barrier in both CTAs, compute / reducer warps of CTA 1 arrive remotely on
// CTA 0's barriers.  Replaces the reference's `_bwd_kernel` (flash_attention_softmax_n/core/flash_attn_triton.py:146-235)
// for the headline shape; semantics identical to fasn_bwd.cu (P recomputed from LSE_n).
//
//   warps 0-7   compute (warp w: TMEM lanes 32(w%4).., query columns 64(w/4)..)     warps 8-11  dQ reducers
//   warp 12     TMA producer        warp 13  MMA issuer (CTA 0 only)        warp 14  TMEM allocator
#include "fasn_common.cuh"
#include "fasn_ptx.cuh"

namespace fasn {

namespace {

// Optional phase timeline (compile with -DFASN_TIMELINE; same tags as fasn_bwd.cu, scripts/timeline.py)
#ifdef FASN_TIMELINE
#define TL_DECL(role) unsigned long long* tl_p = (a.dbg && kt == (int)a.dbg_x && bh == (int)a.dbg_y) ? a.dbg + (role) * 2048 : nullptr; int tl_i = 0;
#define TL_ONLY(cond) do { if (!(cond)) tl_p = nullptr; } while (0)
#define TL(tag) do { if (tl_p && tl_i < 2048) tl_p[tl_i++] = ((unsigned long long)(tag) << 48) | (clock64() & 0xFFFFFFFFFFFFull); } while (0)
#else
#define TL_DECL(role)
#define TL_ONLY(cond)
#define TL(tag)
#endif

constexpr int kBwd2Threads = 512;

struct Bwd2Smem {   // byte offsets; every tile buffer is 1024-byte aligned (128-byte swizzle)
  static constexpr int K = 0;             // own K tile     [2 D-blocks][128 rows][128 B]         A of S^T
  static constexpr int V = 32768;         // own V tile                                           A of dP^T
  static constexpr int KQ = 65536;        // K[256 kv rows of the pair][D half c]  [256][128 B]   B of dQ
  static constexpr int QR = 98304;        // Q_i rows 64c..64c+63  [2 D-blocks][64 rows][128 B]   B of S^T
  static constexpr int QC = 114688;       // Q_i[:, D half c]      [128 rows][128 B]              B of dK
  static constexpr int DOR = 131072;      // dO_i rows 64c..       [2][64][128 B]                 B of dP^T
  static constexpr int DOC = 147456;      // dO_i[:, D half c]     [128][128 B]                   B of dV
  static constexpr int DSQ = 163840;      // dS[q half c][256 kv] as [256 kv rows][128 B]         A of dQ (MN-major)
  static constexpr int DQ = 196608;       // fp32 dQ staging: 2 chunks of [64 rows][128 B]
  static constexpr int DSX = 212992;      // staging of the dS half owed to the peer CTA: [128 kv rows][128 B], bulk-copied into its DSQ
  static constexpr int LSE = 229376;      // [2][128] fp32
  static constexpr int DELTA = 230400;    // [2][128] fp32
  static constexpr int BARS = 231424;
  static constexpr int NUM_BARS = 24;
  static constexpr int TMEM_SLOT = BARS + NUM_BARS * 8;
  static constexpr int BYTES = TMEM_SLOT + 16;
};

FASN_DEVICE void tma_reduce_add_4d_b2(const CUtensorMap* m, const void* smem_src, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.reduce.async.bulk.tensor.4d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3, %4, %5}], [%1];\n" ::"l"(
          reinterpret_cast<uint64_t>(m)),
      "r"(smem_u32(smem_src)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
FASN_DEVICE void bulk_load_1d_b2(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(smem_u32(smem_dst)),
               "l"(reinterpret_cast<uint64_t>(gsrc)), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
// Bulk copy from this CTA's shared memory into the peer's (DSMEM) by the copy engine; the bytes complete on an mbarrier
// of the peer CTA.  (Per-thread st.shared::cluster / st.async stores of the same 16 KB congest the load/store path of
// the whole SM: measured 2x longer dS phases.)
FASN_DEVICE void bulk_copy_to_peer(uint32_t dst_cluster_addr, const void* smem_src, uint32_t bytes, uint32_t bar_cluster_addr) {
  asm volatile("cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(dst_cluster_addr),
               "r"(smem_u32(smem_src)), "r"(bytes), "r"(bar_cluster_addr)
               : "memory");
}
FASN_DEVICE void mbar_arrive_expect_tx_remote(uint32_t bar_cluster_addr, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cluster.b64 _, [%0], %1;\n" ::"r"(bar_cluster_addr), "r"(bytes) : "memory");
}
FASN_DEVICE void named_bar_arrive(uint32_t id, uint32_t nthreads) {
  asm volatile("bar.arrive %0, %1;\n" ::"r"(id), "r"(nthreads) : "memory");
}

template <bool BF16, bool CAUSAL, bool DROPOUT>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kBwd2Threads, 1)
fasn_bwd2_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_q64,
                 const __grid_constant__ CUtensorMap tm_k, const __grid_constant__ CUtensorMap tm_v,
                 const __grid_constant__ CUtensorMap tm_do, const __grid_constant__ CUtensorMap tm_do64,
                 const __grid_constant__ CUtensorMap tm_dk, const __grid_constant__ CUtensorMap tm_dv,
                 const __grid_constant__ CUtensorMap tm_dq64, const BwdArgs a, const TensorView dk_view,
                 const TensorView dv_view) {
  constexpr int D = 128;
  constexpr int BLK_BYTES = 128 * 128;
  // TMEM columns: S^T [0,128) (P^T over [0,32) and [64,96));  dP^T [128,256), then dS^T (16-bit) [128,192) and dQ [192,256);
  // dV [256,384);  dK [384,512)
  constexpr uint32_t TM_S = 0, TM_DP = 128, TM_DST = 128, TM_DQ = 192, TM_DV = 256, TM_DK = 384;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t cr = cluster_ctarank();              // 0 = leader; equals blockIdx.x & 1
  const TileCoord tcd = decode_block(blockIdx.x >> 1, (((a.Skv + 127) >> 7) + 1) >> 1, a.B * a.H, a.sched_group);   // tile = K/V tile pair
  const int kt = 2 * tcd.tile + (int)cr;
  const int k0 = kt * 128;                            // first key of this CTA's tile
  const int k0p = (kt & ~1) * 128;                    // first key of the pair
  const int bh = tcd.bh;
  const int b = bh / a.H;
  const int h = bh - b * a.H;
  const int hk = (a.Hkv == 1) ? 0 : h;

  const int nq = (a.Sq + 127) >> 7;
  int i_start = 0;
  if (CAUSAL) {
    const int first_q = k0p - a.causal_off;           // first query row that sees the pair's first key
    i_start = first_q > 0 ? (first_q >> 7) : 0;
  }
  const int n_iter = nq - i_start;                    // identical in both CTAs of the pair

  if (n_iter <= 0) {
    if (threadIdx.x < 128) {
      const int row = k0 + threadIdx.x;
      if (row < a.Skv) {
        uint4* pk = reinterpret_cast<uint4*>(reinterpret_cast<uint16_t*>(dk_view.ptr) + b * dk_view.sb + h * dk_view.sh + (long long)row * dk_view.ss);
        uint4* pv = reinterpret_cast<uint4*>(reinterpret_cast<uint16_t*>(dv_view.ptr) + b * dv_view.sb + h * dv_view.sh + (long long)row * dv_view.ss);
#pragma unroll
        for (int i = 0; i < D / 8; ++i) { pk[i] = make_uint4(0, 0, 0, 0); pv[i] = make_uint4(0, 0, 0, 0); }
      }
    }
    return;
  }

  extern __shared__ __align__(1024) uint8_t smem[];
  if ((smem_u32(smem) & 1023u) != 0) __trap();
  uint8_t* sK = smem + Bwd2Smem::K;
  uint8_t* sV = smem + Bwd2Smem::V;
  uint8_t* sKq = smem + Bwd2Smem::KQ;
  uint8_t* sQr = smem + Bwd2Smem::QR;
  uint8_t* sQc = smem + Bwd2Smem::QC;
  uint8_t* sDOr = smem + Bwd2Smem::DOR;
  uint8_t* sDOc = smem + Bwd2Smem::DOC;
  uint8_t* sDSq = smem + Bwd2Smem::DSQ;
  uint8_t* sDQ = smem + Bwd2Smem::DQ;
  uint8_t* sDSX = smem + Bwd2Smem::DSX;
  float* sLse = reinterpret_cast<float*>(smem + Bwd2Smem::LSE);
  float* sDelta = reinterpret_cast<float*>(smem + Bwd2Smem::DELTA);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Bwd2Smem::BARS);
  // waited on by the MMA thread of CTA 0 ("leader" barriers: only CTA 0's copy is used)
  uint64_t* kv_full = bars + 0;
  uint64_t* qr_full = bars + 1;
  uint64_t* qc_full = bars + 2;
  uint64_t* dor_full = bars + 3;
  uint64_t* doc_full = bars + 4;
  uint64_t* p_full = bars + 5;      // 16 warp arrivals (8 compute warps x 2 CTAs)
  uint64_t* ds_full = bars + 6;     // dS rows in shared memory: the 8 locally-writing warps (4 per CTA) + CTA 1's relay of its dsx_full
  uint64_t* dq_empty = bars + 7;    // 8 (4 reducer warps x 2 CTAs)
  // signalled in both CTAs by tcgen05.commit multicast
  uint64_t* qr_empty = bars + 8;
  uint64_t* qc_empty = bars + 9;
  uint64_t* dor_empty = bars + 10;
  uint64_t* doc_empty = bars + 11;
  uint64_t* s_full = bars + 12;
  uint64_t* dp_full = bars + 13;
  uint64_t* dq_full = bars + 14;
  uint64_t* dkv_full = bars + 15;
  // CTA-local ring of LSE2 / delta rows
  uint64_t* ld_full = bars + 16;    // [2]
  uint64_t* ld_empty = bars + 18;   // [2], 8 warp arrivals
  // dS exchange: the four compute warps whose query half belongs to the peer stage their rows locally (dsx_ready, 4
  // arrivals), warp 15 bulk-copies the 16 KB into the peer's DSQ buffer and the bytes complete on the PEER's dsx_full
  uint64_t* dsx_ready = bars + 20;
  uint64_t* dsx_full = bars + 21;
  uint64_t* dst_full = bars + 22;   // leader: dS^T of both CTAs is in tensor memory (16 warp arrivals) -> dK may issue
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + Bwd2Smem::TMEM_SLOT);

  if (warp == 12 && lane == 0) {
    tma_prefetch_desc(&tm_q); tma_prefetch_desc(&tm_q64); tma_prefetch_desc(&tm_k); tma_prefetch_desc(&tm_v);
    tma_prefetch_desc(&tm_do); tma_prefetch_desc(&tm_do64); tma_prefetch_desc(&tm_dk); tma_prefetch_desc(&tm_dv);
    tma_prefetch_desc(&tm_dq64);
  }
  if (warp == 13 && lane == 0) {
    mbar_init(kv_full, 1); mbar_init(qr_full, 1); mbar_init(qc_full, 1); mbar_init(dor_full, 1); mbar_init(doc_full, 1);
    mbar_init(p_full, 16); mbar_init(ds_full, 9); mbar_init(dq_empty, 8);
    mbar_init(qr_empty, 1); mbar_init(qc_empty, 1); mbar_init(dor_empty, 1); mbar_init(doc_empty, 1);
    mbar_init(s_full, 1); mbar_init(dp_full, 1); mbar_init(dq_full, 1); mbar_init(dkv_full, 1);
    for (int i = 0; i < 2; ++i) { mbar_init(&ld_full[i], 1); mbar_init(&ld_empty[i], 8); }
    mbar_init(dsx_ready, 4); mbar_init(dsx_full, 1); mbar_init(dst_full, 16);
    fence_mbar_init();
    fence_proxy_async_smem();
  }
  if (warp == 14) { tmem_alloc_pair<512>(tmem_slot); tmem_relinquish_pair(); }
  tc_fence_before();
  cluster_sync_all();              // barriers of both CTAs are initialised before any remote arrive / TMA completion
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp >= 12) {
    setmaxnreg_dec<56>();
    if (warp == 12 && lane == 0) {
      // ---------------------------------------------------------------- TMA producer (both CTAs, own shared memory)
      const bool leader = (cr == 0);
      if (leader) mbar_arrive_expect_tx(kv_full, 2 * 3 * 32768);
#pragma unroll
      for (int db = 0; db < 2; ++db) {
        tma_load_4d_pair(sK + db * BLK_BYTES, &tm_k, kv_full, db * 64, k0, hk, b);
        tma_load_4d_pair(sV + db * BLK_BYTES, &tm_v, kv_full, db * 64, k0, hk, b);
        tma_load_4d_pair(sKq + db * BLK_BYTES, &tm_k, kv_full, 64 * (int)cr, k0p + db * 128, hk, b);
      }
      const float* lse2 = a.delta + (long long)a.B * a.H * a.Sqp;
      TL_DECL(4)
      // Every buffer feeds exactly one MMA group per iteration, so one buffer each gives a full iteration of prefetch.  The
      // loads are issued in the order in which the tensor pipe frees the buffers (dV_j, S^T_{j+1}, dK_j, dP^T_{j+1}), so
      // this thread never waits on a late buffer while an early one is already free.
      auto load_qr = [&](int j) {
        mbar_wait(qr_empty, (j & 1) ^ 1);
        if (leader) mbar_arrive_expect_tx(qr_full, 2 * 16384);
#pragma unroll
        for (int db = 0; db < 2; ++db) tma_load_4d_pair(sQr + db * 8192, &tm_q64, qr_full, db * 64, (i_start + j) * 128 + 64 * (int)cr, h, b);
      };
      auto load_dor = [&](int j) {
        const int qi0 = (i_start + j) * 128;
        mbar_wait(dor_empty, (j & 1) ^ 1);
        if (leader) mbar_arrive_expect_tx(dor_full, 2 * 16384);
#pragma unroll
        for (int db = 0; db < 2; ++db) tma_load_4d_pair(sDOr + db * 8192, &tm_do64, dor_full, db * 64, qi0 + 64 * (int)cr, h, b);
        const int rs = j & 1;
        mbar_wait(&ld_empty[rs], ((j >> 1) & 1) ^ 1);
        mbar_arrive_expect_tx(&ld_full[rs], 1024);
        bulk_load_1d_b2(sLse + rs * 128, lse2 + (long long)bh * a.Sqp + qi0, 512, &ld_full[rs]);
        bulk_load_1d_b2(sDelta + rs * 128, a.delta + (long long)bh * a.Sqp + qi0, 512, &ld_full[rs]);
      };
      load_qr(0);
      load_dor(0);
      for (int it = 0; it < n_iter; ++it) {
        const uint32_t ph = it & 1;
        const int qi0 = (i_start + it) * 128;
        TL(40);
        mbar_wait(doc_empty, ph ^ 1);
        TL(44);
        if (leader) mbar_arrive_expect_tx(doc_full, 2 * 16384);
        tma_load_4d_pair(sDOc, &tm_do, doc_full, 64 * (int)cr, qi0, h, b);
        if (it + 1 < n_iter) load_qr(it + 1);
        TL(41);
        mbar_wait(qc_empty, ph ^ 1);
        TL(45);
        if (leader) mbar_arrive_expect_tx(qc_full, 2 * 16384);
        tma_load_4d_pair(sQc, &tm_q, qc_full, 64 * (int)cr, qi0, h, b);
        if (it + 1 < n_iter) load_dor(it + 1);
        TL(42);
      }
    } else if (warp == 13 && cr == 0) {
      // ---------------------------------------------------------------- MMA issuer (leader CTA only)
      constexpr uint32_t idesc_kk = umma_idesc(BF16, 256, 128, false, false);   // S^T, dP^T
      constexpr uint32_t idesc_ts = umma_idesc(BF16, 256, 128, false, true);    // dV, dK: A in TMEM, B MN-major
      constexpr uint32_t idesc_dq = umma_idesc(BF16, 128, 128, true, true);     // dQ: A, B MN-major, 64 rows per CTA
      constexpr uint32_t hi_desc = umma_desc_hi(1024);
      const uint32_t tm = __shfl_sync(0xffffffffu, tmem_base, 0);
      const uint32_t k_km = umma_desc_lo(smem_u32(sK), 16), v_km = umma_desc_lo(smem_u32(sV), 16);
      const uint32_t qr_km = umma_desc_lo(smem_u32(sQr), 16), dor_km = umma_desc_lo(smem_u32(sDOr), 16);
      const uint32_t qc_mn = umma_desc_lo(smem_u32(sQc), BLK_BYTES), doc_mn = umma_desc_lo(smem_u32(sDOc), BLK_BYTES);
      const uint32_t dsq_mn = umma_desc_lo(smem_u32(sDSq), BLK_BYTES), kq_mn = umma_desc_lo(smem_u32(sKq), BLK_BYTES);
      auto issue_kmajor = [&](uint32_t tm_dst, uint32_t a_lo, uint32_t b_lo) {   // D[256 x 128] = A B^T over the head dim
#pragma unroll
        for (int kb = 0; kb < D / 16; ++kb) {
          const uint32_t offa = ((kb >> 2) * BLK_BYTES + (kb & 3) * 32) >> 4;    // A: [2][128 rows][128 B]
          const uint32_t offb = ((kb >> 2) * 8192 + (kb & 3) * 32) >> 4;         // B: [2][64 rows][128 B] per CTA
          umma2_ss(tm + tm_dst, umma_desc_join(a_lo + offa, hi_desc), umma_desc_join(b_lo + offb, hi_desc), idesc_kk, kb > 0 ? 1u : 0u);
        }
      };
      TL_DECL(0)
      TL_ONLY(lane == 0);
      mbar_wait(kv_full, 0);
      mbar_wait(qr_full, 0);
      tc_fence_after();
      TL(1);
      if (elect_one()) { issue_kmajor(TM_S, k_km, qr_km); tc_commit_pair(s_full); tc_commit_pair(qr_empty); }
      __syncwarp();
      mbar_wait(dor_full, 0);
      tc_fence_after();
      if (elect_one()) { issue_kmajor(TM_DP, v_km, dor_km); tc_commit_pair(dp_full); tc_commit_pair(dor_empty); }
      __syncwarp();
      for (int it = 0; it < n_iter; ++it) {
        const uint32_t ph = it & 1;
        const bool more = it + 1 < n_iter;
        // dV += P^T dO_i
        mbar_wait(doc_full, ph);
        TL(7);
        mbar_wait(p_full, ph);
        tc_fence_after();
        TL(2);
        if (elect_one()) {
#pragma unroll
          for (int kb = 0; kb < 8; ++kb)
            umma2_ts(tm + TM_DV, tm + TM_S + (kb >> 2) * 64 + (kb & 3) * 8, umma_desc_join(doc_mn + kb * (2048 >> 4), hi_desc),
                     idesc_ts, (it > 0 || kb > 0) ? 1u : 0u);
          tc_commit_pair(doc_empty);
        }
        __syncwarp();
        // S^T of the next Q tile (overwrites P^T: ordered behind the dV MMAs on the tensor pipe)
        if (more) {
          mbar_wait(qr_full, ph ^ 1);
          tc_fence_after();
          TL(3);
          if (elect_one()) { issue_kmajor(TM_S, k_km, qr_km); tc_commit_pair(s_full); tc_commit_pair(qr_empty); }
          __syncwarp();
        }
        // dK += dS^T Q_i first: its A operand is in tensor memory, so it runs while the peer's half of dS is still in flight
        mbar_wait(qc_full, ph);
        TL(8);
        mbar_wait(dst_full, ph);
        tc_fence_after();
        if (elect_one()) {
#pragma unroll
          for (int kb = 0; kb < 8; ++kb)
            umma2_ts(tm + TM_DK, tm + TM_DST + kb * 8, umma_desc_join(qc_mn + kb * (2048 >> 4), hi_desc), idesc_ts,
                     (it > 0 || kb > 0) ? 1u : 0u);
          tc_commit_pair(qc_empty);
        }
        __syncwarp();
        // dQ_i = dS K once both halves of dS are in place in both CTAs
        mbar_wait(ds_full, ph);
        mbar_wait(dsx_full, ph);       // the peer's half of dS has landed in this CTA's DSQ buffer
        tc_fence_after();
        TL(4);
        if (elect_one()) {
#pragma unroll
          for (int kb = 0; kb < 16; ++kb)
            umma2_ss(tm + TM_DQ, umma_desc_join(dsq_mn + kb * (2048 >> 4), hi_desc), umma_desc_join(kq_mn + kb * (2048 >> 4), hi_desc),
                     idesc_dq, kb > 0 ? 1u : 0u);
          tc_commit_pair(dq_full);
        }
        __syncwarp();
        // dP^T of the next tile overwrites dS^T (read by dK: earlier on the same pipe) and dQ (drained by the reducers)
        if (more) {
          mbar_wait(dor_full, ph ^ 1);
          TL(5);
          mbar_wait(dq_empty, ph);
          tc_fence_after();
          TL(6);
          if (elect_one()) { issue_kmajor(TM_DP, v_km, dor_km); tc_commit_pair(dp_full); tc_commit_pair(dor_empty); }
          __syncwarp();
        }
      }
      if (elect_one()) tc_commit_pair(dkv_full);
      __syncwarp();
    } else if (warp == 13 && lane == 0) {
      // ---------------------------------------------------------------- CTA 1: relay "the peer's dS half has landed here" to the leader
      const uint32_t ds_full_c = mapa_shared(smem_u32(ds_full), 0);
      for (int it = 0; it < n_iter; ++it) {
        mbar_wait(dsx_full, it & 1);
        mbar_arrive_remote(ds_full_c);
      }
    } else if (warp == 15 && lane == 0) {
      // ---------------------------------------------------------------- dS exchange: 16 KB bulk copy into the peer's DSQ rows 128 c ..
      const uint32_t peer = cr ^ 1u;
      const uint32_t dst = mapa_shared(smem_u32(sDSq) + 128u * cr * 128u, peer);
      const uint32_t bar = mapa_shared(smem_u32(dsx_full), peer);
      for (int it = 0; it < n_iter; ++it) {
        mbar_wait(dsx_ready, it & 1);
        mbar_arrive_expect_tx_remote(bar, 16384);
#pragma unroll
        for (int ch = 0; ch < 4; ++ch) bulk_copy_to_peer(dst + ch * 4096, sDSX + ch * 4096, 4096, bar);
      }
    }
  } else if (warp >= 8) {
    // ------------------------------------------------------------------ dQ reducers
    // This CTA holds dQ rows 64 c .. 64 c + 63 of the tile: columns 0-63 on TMEM lanes 0-63, columns 64-127 on lanes
    // 64-127 (the M = 128 cta_group::2 accumulator layout), 64 TMEM columns.
    setmaxnreg_dec<104>();
    const int w = warp & 3;
    const int row = 32 * (w & 1) + lane;                  // query row inside this CTA's 64-row half
    const int dhalf = w >> 1;                             // which 64 columns of the head dim
    const uint32_t lane_off = static_cast<uint32_t>(w * 32) << 16;
    const uint32_t dq_empty_c = mapa_shared(smem_u32(dq_empty), 0);
    TL_DECL(3)
    TL_ONLY(threadIdx.x == 256);
    for (int it = 0; it < n_iter; ++it) {
      const int qi0 = (i_start + it) * 128 + 64 * (int)cr;
      mbar_wait(dq_full, it & 1);
      tc_fence_after();
      TL(30);
      uint32_t v[64];
      tmem_ld_x32(tmem_base + lane_off + TM_DQ, v);
      tmem_ld_x32(tmem_base + lane_off + TM_DQ + 32, v + 32);
      tmem_wait_ld();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_remote(dq_empty_c);     // the dQ columns may be overwritten by dP^T now
      TL(31);
#pragma unroll
      for (int hb = 0; hb < 2; ++hb) {
        if (threadIdx.x == 256) tma_store_wait_read_all();   // the previous round's reduces have read both staging chunks
        named_bar_sync(2, 128);
#pragma unroll
        for (int g = 0; g < 8; ++g)
          *reinterpret_cast<uint4*>(sDQ + dhalf * 8192 + row * 128 + ((g ^ (row & 7)) << 4)) =
              make_uint4(v[hb * 32 + g * 4], v[hb * 32 + g * 4 + 1], v[hb * 32 + g * 4 + 2], v[hb * 32 + g * 4 + 3]);
        fence_proxy_async_smem();
        named_bar_sync(3, 128);
        if (threadIdx.x == 256) {
          tma_reduce_add_4d_b2(&tm_dq64, sDQ, hb * 32, qi0, bh, 0);
          tma_reduce_add_4d_b2(&tm_dq64, sDQ + 8192, 64 + hb * 32, qi0, bh, 0);
          tma_store_commit();
        }
      }
      TL(32);
    }
    if (threadIdx.x == 256) tma_store_wait_all();
  } else {
    // -------------------------------------------------------------------- compute warps
    setmaxnreg_inc<176>();
    const int quarter = warp & 3;
    const int half = warp >> 2;
    const int r = quarter * 32 + lane;                     // kv row inside this CTA's tile
    const int kv_row = k0 + r;
    const uint32_t lane_off = static_cast<uint32_t>(quarter * 32) << 16;
    const bool kv_valid = kv_row < a.Skv;
    const uint32_t bh_global = a.bh_offset + bh;
    const uint32_t kvw = (uint32_t)(kv_row >> 5);
    const bool kv_tail = (k0 + 128 > a.Skv);
    const float2 c2 = make_float2(a.scale_log2, a.scale_log2);
    const uint32_t p_full_c = mapa_shared(smem_u32(p_full), 0);
    const uint32_t ds_full_c = mapa_shared(smem_u32(ds_full), 0);
    // dS row of this thread: row 128 c + r of the dS[q half][256 kv] buffer of CTA `half` (local if half == c)
    // (the peer's half is staged in DSX row r and bulk-copied by warp 15)
    const bool ds_is_local = ((uint32_t)half == cr);
    uint8_t* ds_local = ds_is_local ? sDSq + (uint32_t)(128 * (int)cr + r) * 128 : sDSX + r * 128;
    const uint32_t dst_full_c = mapa_shared(smem_u32(dst_full), 0);

    uint32_t keep_next0 = 0xFFFFFFFFu, keep_next1 = 0xFFFFFFFFu;
    auto make_keep = [&](int qc0n) {
      if constexpr (DROPOUT) {
        keep_next0 = warp_transpose_bits(dropout_keep_word(a.key, bh_global, (uint32_t)(qc0n + lane), kvw, a.drop_thr), lane);
        keep_next1 = warp_transpose_bits(dropout_keep_word(a.key, bh_global, (uint32_t)(qc0n + 32 + lane), kvw, a.drop_thr), lane);
      }
    };
    make_keep(i_start * 128 + half * 64);
    TL_DECL(1 + half)
    TL_ONLY(threadIdx.x == 0 || threadIdx.x == 128);

    for (int it = 0; it < n_iter; ++it) {
      const int rs = it & 1;
      const uint32_t rph = (it >> 1) & 1;
      const uint32_t ph = it & 1;
      const int qi0 = (i_start + it) * 128;
      const int qc0 = qi0 + half * 64;
      const uint32_t keep0 = keep_next0, keep1 = keep_next1;
      // ---- P^T = 2^(S^T c - LSE2)
      TL(10);
      mbar_wait(&ld_full[rs], rph);
      mbar_wait(s_full, ph);
      tc_fence_after();
      TL(11);
      float p[64];
      {
        uint32_t* pr = reinterpret_cast<uint32_t*>(p);
        tmem_ld_x32(tmem_base + lane_off + TM_S + half * 64, pr);
        tmem_ld_x32(tmem_base + lane_off + TM_S + half * 64 + 32, pr + 32);
        tmem_wait_ld();
      }
      const float* lse_s = sLse + rs * 128 + half * 64;
#pragma unroll
      for (int c = 0; c < 64; c += 4) {
        const float4 l4 = *reinterpret_cast<const float4*>(lse_s + c);
        const float2 a01 = __ffma2_rn(make_float2(p[c], p[c + 1]), c2, make_float2(-l4.x, -l4.y));
        const float2 a23 = __ffma2_rn(make_float2(p[c + 2], p[c + 3]), c2, make_float2(-l4.z, -l4.w));
        p[c] = ex2(a01.x); p[c + 1] = ex2(a01.y); p[c + 2] = ex2(a23.x); p[c + 3] = ex2(a23.y);
      }
      const bool diag = CAUSAL && (qi0 + a.causal_off < k0 + 127);
      if (diag || kv_tail) {
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
      __syncwarp();
      if (lane == 0) mbar_arrive_remote(p_full_c);
      TL(12);
      if (it + 1 < n_iter) make_keep(qc0 + 128);
      // ---- dS'^T = P^T o (Z dP^T - (1-p) delta)
      mbar_wait(dp_full, ph);      // also: dQ / dK of the previous iteration (readers of dS) have completed in both CTAs
      tc_fence_after();
      TL(13);
      const float* del_s = sDelta + rs * 128 + half * 64;
      uint32_t outs[32];
#pragma unroll
      for (int g = 0; g < 2; ++g) {
        uint32_t dpr[32];
        tmem_ld_x32(tmem_base + lane_off + TM_DP + half * 64 + g * 32, dpr);
        tmem_wait_ld();
        // the other half's threads write their dS^T over columns [160,192) = this half-0 thread's second chunk
        if (half == 0 && g == 1) named_bar_arrive(4 + quarter, 64);
        uint32_t* out = outs + g * 16;
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
        }
        // dS for the dQ MMA: 64 bytes of the [kv row][64 queries of this half] row, in the CTA that owns this query half
#pragma unroll
        for (int q4 = 0; q4 < 4; ++q4) {
          const uint32_t off = (uint32_t)(((g * 4 + q4) ^ (r & 7)) << 4);
          *reinterpret_cast<uint4*>(ds_local + off) = make_uint4(out[q4 * 4], out[q4 * 4 + 1], out[q4 * 4 + 2], out[q4 * 4 + 3]);
        }
        // dS^T for the dK MMA (A operand in TMEM): query columns 64 half + 32 g .. -> TMEM columns TM_DST + 32 half + 16 g ..
        if (half == 0) tmem_st_x16(tmem_base + lane_off + TM_DST + g * 16, out);
      }
      fence_proxy_async_smem();                          // generic-proxy dS stores -> tensor-core reads / bulk copy
      __syncwarp();
      if (lane == 0) {                                   // shared-memory half first: the exchange starts as early as possible
        if (ds_is_local) mbar_arrive_remote(ds_full_c); else mbar_arrive(dsx_ready);
      }
      TL(16);
      if (half == 1) {
        named_bar_sync(4 + quarter, 64);                 // the half-0 thread of this lane has read dP^T columns [160,192)
        tmem_st_x32(tmem_base + lane_off + TM_DST + 32, outs);
      }
      tmem_wait_st();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) { mbar_arrive_remote(dst_full_c); mbar_arrive(&ld_empty[rs]); }
      TL(15);
    }

    // ------------------------------------------------------------------ dK, dV epilogue (own 128 kv rows)
    mbar_wait(dkv_full, 0);
    tc_fence_after();
#pragma unroll
    for (int which = 0; which < 2; ++which) {
      uint8_t* stage = which == 0 ? sV : sK;
      const uint32_t tm_src = which == 0 ? TM_DV : TM_DK;
      const float mul = which == 0 ? (DROPOUT ? a.inv_keep : 1.f) : a.scale;   // a.scale already carries 1/(1-p)
      constexpr int COLS = D / 2;
#pragma unroll
      for (int cb = 0; cb < COLS / 32; ++cb) {
        uint32_t v[32];
        tmem_ld_x32(tmem_base + lane_off + tm_src + half * COLS + cb * 32, v);
        tmem_wait_ld();
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
    }
    fence_proxy_async_smem();
    named_bar_sync(1, 256);
    if (threadIdx.x == 0) {
#pragma unroll
      for (int db = 0; db < 2; ++db) {
        tma_store_4d(&tm_dv, sV + db * BLK_BYTES, db * 64, k0, h, b);
        tma_store_4d(&tm_dk, sK + db * BLK_BYTES, db * 64, k0, h, b);
      }
      tma_store_commit();
      tma_store_wait_all();
    }
  }

  tc_fence_before();
  cluster_sync_all();              // no CTA leaves (or frees TMEM) while its peer may still read its shared memory / arrive on its barriers
  if (warp == 14) tmem_dealloc_pair<512>(tmem_base);
}

}  // namespace

template <bool BF16, bool CAUSAL, bool DROPOUT>
static cudaError_t launch_bwd2_t(const CUtensorMap& tq, const CUtensorMap& tq64, const CUtensorMap& tk, const CUtensorMap& tv,
                                 const CUtensorMap& tdo, const CUtensorMap& tdo64, const CUtensorMap& tdk, const CUtensorMap& tdv,
                                 const CUtensorMap& tdq64, const BwdArgs& a, const TensorView& dk, const TensorView& dv, cudaStream_t stream) {
  auto kern = fasn_bwd2_kernel<BF16, CAUSAL, DROPOUT>;
  constexpr int smem = Bwd2Smem::BYTES;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  if (e != cudaSuccess) return e;
  const int nkv = (a.Skv + 127) / 128;
  dim3 grid(2 * ((nkv + 1) / 2) * a.B * a.H, 1, 1);  // CTA pairs along x (static cluster dims 2 x 1 x 1)
  kern<<<grid, kBwd2Threads, smem, stream>>>(tq, tq64, tk, tv, tdo, tdo64, tdk, tdv, tdq64, a, dk, dv);
  return cudaGetLastError();
}

cudaError_t launch_bwd2(bool bf16, bool causal, bool dropout, const CUtensorMap& tq, const CUtensorMap& tq64, const CUtensorMap& tk,
                        const CUtensorMap& tv, const CUtensorMap& tdo, const CUtensorMap& tdo64, const CUtensorMap& tdk,
                        const CUtensorMap& tdv, const CUtensorMap& tdq64, const BwdArgs& a, const TensorView& dk, const TensorView& dv,
                        cudaStream_t stream) {
#define FASN_BWD2_CASE(BF_, C_, DR_) \
  if (bf16 == BF_ && causal == C_ && dropout == DR_) return launch_bwd2_t<BF_, C_, DR_>(tq, tq64, tk, tv, tdo, tdo64, tdk, tdv, tdq64, a, dk, dv, stream);
  FASN_BWD2_CASE(false, false, false) FASN_BWD2_CASE(false, false, true) FASN_BWD2_CASE(false, true, false) FASN_BWD2_CASE(false, true, true)
  FASN_BWD2_CASE(true, false, false)  FASN_BWD2_CASE(true, false, true)  FASN_BWD2_CASE(true, true, false)  FASN_BWD2_CASE(true, true, true)
#undef FASN_BWD2_CASE
  return cudaErrorInvalidValue;
}

}  // namespace fasn
