// Small auxiliary kernels: the dropout keep-mask dump (test hook) and the single-tile tcgen05 bring-up probe.
#include "fasn_common.cuh"
#include "fasn_ptx.cuh"

namespace fasn {

namespace {

__global__ void fasn_dropout_mask_kernel(uint8_t* __restrict__ out, int BH, int Sq, int Skv, uint32_t thr, PhiloxKey key,
                                         uint32_t bh_offset) {
  const int nw = (Skv + 31) >> 5;
  const long long total = (long long)BH * Sq * nw;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int w = (int)(i % nw);
    const long long t = i / nw;
    const int q = (int)(t % Sq);
    const int bh = (int)(t / Sq);
    const uint32_t word = dropout_keep_word(key, bh_offset + (uint32_t)bh, (uint32_t)q, (uint32_t)w, thr);
    uint8_t* dst = out + ((long long)bh * Sq + q) * Skv + (long long)w * 32;
    const int nvalid = min(32, Skv - w * 32);
    for (int bit = 0; bit < nvalid; ++bit) dst[bit] = (word >> bit) & 1u;
  }
}

// One 128x128x128 MMA through the same descriptor builders as the attention kernels (see fasn.h: fasn_probe).
template <bool BF16>
__global__ void __launch_bounds__(160, 1)
fasn_probe_kernel(const __grid_constant__ CUtensorMap tm_x, const __grid_constant__ CUtensorMap tm_y, int mode,
                  const uint16_t* __restrict__ x, float* __restrict__ c) {
  constexpr int BLK = 128 * 128;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sX = smem;
  uint8_t* sY = smem + 2 * BLK;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sY + 2 * BLK);   // [0] tma, [1] mma done, [2] A staged in TMEM
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 3);
  uint8_t* sAx = sY + 2 * BLK + 1024;          // modes 4-6: extension operands (no-swizzle K-major, one K-step of 16)
  uint8_t* sBx = sAx + 4096;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 4 && lane == 0) {
    mbar_init(&bars[0], 1); mbar_init(&bars[1], 1); mbar_init(&bars[2], 128);
    fence_mbar_init();
    fence_proxy_async_smem();
  }
  if (warp == 0) { tmem_alloc<512>(tmem_slot); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 4) {
    if (lane == 0) {
      mbar_arrive_expect_tx(&bars[0], 4 * BLK);
      for (int db = 0; db < 2; ++db) {
        tma_load_4d(sX + db * BLK, &tm_x, &bars[0], db * 64, 0, 0, 0);
        tma_load_4d(sY + db * BLK, &tm_y, &bars[0], db * 64, 0, 0, 0);
      }
      mbar_wait(&bars[0], 0);
      if (mode == 1 || mode >= 4) mbar_wait(&bars[2], 0);
      tc_fence_after();
      const uint32_t sX_u = smem_u32(sX), sY_u = smem_u32(sY);
      for (int kb = 0; kb < 8; ++kb) {
        const uint32_t koff = (kb >> 2) * BLK + (kb & 3) * 32;      // K-major advance
        const uint32_t moff = kb * 2048;                            // MN-major advance
        const uint32_t acc = kb > 0 ? 1u : 0u;
        if (mode == 0 || mode >= 4)
          umma_ss(tmem_base, umma_smem_desc(sX_u + koff, 16, 1024), umma_smem_desc(sY_u + koff, 16, 1024),
                  umma_idesc(BF16, 128, 128, false, false), acc);
        else if (mode == 1)
          umma_ts(tmem_base, tmem_base + 256 + kb * 8, umma_smem_desc(sY_u + moff, BLK, 1024),
                  umma_idesc(BF16, 128, 128, false, true), acc);
        else if (mode == 2)
          umma_ss(tmem_base, umma_smem_desc(sX_u + moff, BLK, 1024), umma_smem_desc(sY_u + moff, BLK, 1024),
                  umma_idesc(BF16, 128, 128, true, true), acc);
        else
          umma_ss(tmem_base, umma_smem_desc(sX_u + koff, 16, 1024), umma_smem_desc(sY_u + moff, BLK, 1024),
                  umma_idesc(BF16, 128, 128, false, true), acc);
      }
      if (mode >= 4) {
        // ninth K-step: C[i][j] += sum_k Ax[i][k] * Bx[j][k] with the extension operands written below.
        //   mode 4: both 16-byte K-chunks and all sixteen 8-row groups stored (LBO = 128, SBO = 256)
        //   mode 5: the second K-chunk aliases the first (LBO = 0, SBO = 128): the step counts every term twice
        //   mode 6: mode 5 + the A side is ONE core matrix shared by all row groups (SBO = 0)
        const uint32_t lbo = mode == 4 ? 128u : 0u, sbo = mode == 4 ? 256u : 128u;
        umma_ss(tmem_base, umma_smem_desc_noswz(smem_u32(sAx), lbo, mode == 6 ? 0u : sbo), umma_smem_desc_noswz(smem_u32(sBx), lbo, sbo),
                umma_idesc(BF16, 128, 128, false, false), 1u);
      }
      tc_commit(&bars[1]);
    }
  } else {
    const int r = threadIdx.x;   // 0..127
    const uint32_t lane_off = static_cast<uint32_t>(warp * 32) << 16;
    if (mode >= 4) {
      // A side: rows of (1, 1, 1, 0, ...); B side: row j = the three-term 16-bit split of e_j = 3.25 * x[j][0] (+ zeros).
      // Expected: C = X Y^T + e_j (mode 4), + 2 e_j (modes 5, 6).
      const uint16_t one = BF16 ? 0x3F80 : 0x3C00;
      const float e = 3.25f * cvt16_to_f32<BF16>(x[r * 128]);
      const uint16_t h0 = (uint16_t)(pack2<BF16>(e, 0.f) & 0xFFFF);
      const float r1 = e - cvt16_to_f32<BF16>(h0);
      const uint16_t h1 = (uint16_t)(pack2<BF16>(r1, 0.f) & 0xFFFF);
      const uint16_t h2 = (uint16_t)(pack2<BF16>(r1 - cvt16_to_f32<BF16>(h1), 0.f) & 0xFFFF);
      const uint4 arow = make_uint4(one | ((uint32_t)one << 16), one, 0u, 0u);
      const uint4 brow = make_uint4(h0 | ((uint32_t)h1 << 16), h2, 0u, 0u);
      const uint4 zero = make_uint4(0u, 0u, 0u, 0u);
      if (mode == 4) {            // [row group][K-chunk][row in group][16 B]
        *reinterpret_cast<uint4*>(sAx + (r >> 3) * 256 + (r & 7) * 16) = arow;
        *reinterpret_cast<uint4*>(sAx + (r >> 3) * 256 + 128 + (r & 7) * 16) = zero;
        *reinterpret_cast<uint4*>(sBx + (r >> 3) * 256 + (r & 7) * 16) = brow;
        *reinterpret_cast<uint4*>(sBx + (r >> 3) * 256 + 128 + (r & 7) * 16) = zero;
      } else {                    // [row group][row in group][16 B]; mode 6 reads only the first group of the A side
        *reinterpret_cast<uint4*>(sAx + r * 16) = arow;
        *reinterpret_cast<uint4*>(sBx + r * 16) = brow;
      }
      fence_proxy_async_smem();
      mbar_arrive(&bars[2]);
    }
    if (mode == 1) {
      uint32_t a[64];
      const uint32_t* src = reinterpret_cast<const uint32_t*>(x + r * 128);
#pragma unroll
      for (int i = 0; i < 64; ++i) a[i] = src[i];
      tmem_st_x32(tmem_base + lane_off + 256, a);
      tmem_st_x32(tmem_base + lane_off + 256 + 32, a + 32);
      tmem_wait_st();
      tc_fence_before();
      mbar_arrive(&bars[2]);
    }
    mbar_wait(&bars[1], 0);
    tc_fence_after();
#pragma unroll
    for (int cb = 0; cb < 4; ++cb) {
      uint32_t v[32];
      tmem_ld_x32(tmem_base + lane_off + cb * 32, v);
      tmem_wait_ld();
#pragma unroll
      for (int i = 0; i < 32; ++i) c[r * 128 + cb * 32 + i] = __uint_as_float(v[i]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<512>(tmem_base);
}

// CTA-pair probe (cluster of 2, tcgen05 cta_group::2): the operand forms of the paired backward kernel (fasn_bwd2.cu).
//   mode 10: C[256x128] = X[256x128] Y[128x128]^T   M=256 SS, both K-major; B split along N (64 rows of Y per CTA)
//   mode 11: C[128x128] = X[256x128]^T Y[256x128]   M=128 SS, A and B MN-major, K = 256; A is written with generic
//            (local and remote st.shared::cluster) stores by both CTAs, as the dS exchange of the backward does
//   mode 12: C[256x128] = X[256x128] Y[128x128]     M=256 TS, A in each CTA's TMEM, B MN-major split along N
template <bool BF16>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(160, 1)
fasn_probe_pair_kernel(const __grid_constant__ CUtensorMap tm_x, const __grid_constant__ CUtensorMap tm_y,
                       const __grid_constant__ CUtensorMap tm_y64, int mode, const uint16_t* __restrict__ x,
                       float* __restrict__ c) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;                 // 32 KB
  uint8_t* sB = smem + 32768;         // 32 KB
  uint64_t* bars = reinterpret_cast<uint64_t*>(sB + 32768);   // [0] tma (leader), [1] mma done (both), [2] A ready (leader)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 3);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t cr = cluster_ctarank();

  if (warp == 4 && lane == 0) {
    mbar_init(&bars[0], 1); mbar_init(&bars[1], 1); mbar_init(&bars[2], 8);
    fence_mbar_init();
    fence_proxy_async_smem();
  }
  if (warp == 0) { tmem_alloc_pair<512>(tmem_slot); tmem_relinquish_pair(); }
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 4) {
    if (lane == 0) {
      if (mode == 10) {
        if (cr == 0) mbar_arrive_expect_tx(&bars[0], 2 * (32768 + 16384));
        for (int db = 0; db < 2; ++db) {
          tma_load_4d_pair(sA + db * 16384, &tm_x, &bars[0], db * 64, 128 * cr, 0, 0);
          tma_load_4d_pair(sB + db * 8192, &tm_y64, &bars[0], db * 64, 64 * cr, 0, 0);
        }
      } else if (mode == 11) {
        if (cr == 0) mbar_arrive_expect_tx(&bars[0], 2 * 32768);
        tma_load_4d_pair(sB, &tm_y, &bars[0], 64 * cr, 0, 0, 0);
        tma_load_4d_pair(sB + 16384, &tm_y, &bars[0], 64 * cr, 128, 0, 0);
      } else {
        if (cr == 0) mbar_arrive_expect_tx(&bars[0], 2 * 16384);
        tma_load_4d_pair(sB, &tm_y, &bars[0], 64 * cr, 0, 0, 0);
      }
      if (cr == 0) {
        mbar_wait(&bars[0], 0);
        if (mode != 10) mbar_wait_cluster(&bars[2], 0);
        tc_fence_after();
        const uint32_t sA_u = smem_u32(sA), sB_u = smem_u32(sB);
        if (mode == 10) {
          for (int kb = 0; kb < 8; ++kb)
            umma2_ss(tmem_base, umma_smem_desc(sA_u + (kb >> 2) * 16384 + (kb & 3) * 32, 16, 1024),
                     umma_smem_desc(sB_u + (kb >> 2) * 8192 + (kb & 3) * 32, 16, 1024), umma_idesc(BF16, 256, 128, false, false), kb > 0);
        } else if (mode == 11) {
          for (int kb = 0; kb < 16; ++kb)
            umma2_ss(tmem_base, umma_smem_desc(sA_u + kb * 2048, 16384, 1024), umma_smem_desc(sB_u + kb * 2048, 16384, 1024),
                     umma_idesc(BF16, 128, 128, true, true), kb > 0);
        } else {
          for (int kb = 0; kb < 8; ++kb)
            umma2_ts(tmem_base, tmem_base + 256 + kb * 8, umma_smem_desc(sB_u + kb * 2048, 16384, 1024),
                     umma_idesc(BF16, 256, 128, false, true), kb > 0);
        }
        tc_commit_pair(&bars[1]);
      }
    }
  } else {
    const int r = threadIdx.x;   // 0..127
    const uint32_t lane_off = static_cast<uint32_t>(warp * 32) << 16;
    const uint32_t ready_bar = mapa_shared(smem_u32(&bars[2]), 0);
    if (mode == 11) {
      const uint4* src = reinterpret_cast<const uint4*>(x + (size_t)(128 * cr + r) * 128);
#pragma unroll
      for (int ch = 0; ch < 16; ++ch) {
        const uint4 v = src[ch];
        const uint32_t dst_cta = ch >> 3;                                   // q half -> owning CTA
        const uint32_t off = (128 * cr + r) * 128 + (((ch & 7) ^ (r & 7)) << 4);
        if (dst_cta == cr) *reinterpret_cast<uint4*>(sA + off) = v;
        else st_cluster_v4(mapa_shared(smem_u32(sA) + off, dst_cta), v.x, v.y, v.z, v.w);
      }
      fence_proxy_async_all();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(ready_bar);
    } else if (mode == 12) {
      uint32_t a[64];
      const uint32_t* src = reinterpret_cast<const uint32_t*>(x + (size_t)(128 * cr + r) * 128);
#pragma unroll
      for (int i = 0; i < 64; ++i) a[i] = src[i];
      tmem_st_x32(tmem_base + lane_off + 256, a);
      tmem_st_x32(tmem_base + lane_off + 256 + 32, a + 32);
      tmem_wait_st();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(ready_bar);
    }
    mbar_wait(&bars[1], 0);
    tc_fence_after();
    if (mode == 11) {
      uint32_t v[64];
      tmem_ld_x32(tmem_base + lane_off, v);
      tmem_ld_x32(tmem_base + lane_off + 32, v + 32);
      tmem_wait_ld();
      const int row = 64 * cr + 32 * (warp & 1) + lane, col0 = 64 * (warp >> 1);
#pragma unroll
      for (int i = 0; i < 64; ++i) c[row * 128 + col0 + i] = __uint_as_float(v[i]);
    } else {
#pragma unroll
      for (int cb = 0; cb < 4; ++cb) {
        uint32_t v[32];
        tmem_ld_x32(tmem_base + lane_off + cb * 32, v);
        tmem_wait_ld();
#pragma unroll
        for (int i = 0; i < 32; ++i) c[(128 * cr + r) * 128 + cb * 32 + i] = __uint_as_float(v[i]);
      }
    }
  }
  tc_fence_before();
  cluster_sync_all();
  if (warp == 0) tmem_dealloc_pair<512>(tmem_base);
}

}  // namespace

cudaError_t launch_probe_pair(int mode, bool bf16, const CUtensorMap& tx, const CUtensorMap& ty, const CUtensorMap& ty64,
                              const void* x, float* c, cudaStream_t stream) {
  constexpr int smem = 1024 + 65536 + 64;
  cudaError_t e;
  if (bf16) {
    e = cudaFuncSetAttribute(fasn_probe_pair_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return e;
    fasn_probe_pair_kernel<true><<<2, 160, smem, stream>>>(tx, ty, ty64, mode, reinterpret_cast<const uint16_t*>(x), c);
  } else {
    e = cudaFuncSetAttribute(fasn_probe_pair_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return e;
    fasn_probe_pair_kernel<false><<<2, 160, smem, stream>>>(tx, ty, ty64, mode, reinterpret_cast<const uint16_t*>(x), c);
  }
  return cudaGetLastError();
}

cudaError_t launch_dropout_mask(uint8_t* out, int B, int H, int Sq, int Skv, uint32_t thr, PhiloxKey key,
                                uint32_t bh_offset, cudaStream_t stream) {
  const long long total = (long long)B * H * Sq * ((Skv + 31) / 32);
  int blocks = (int)((total + 255) / 256);
  if (blocks > 148 * 16) blocks = 148 * 16;
  if (blocks < 1) blocks = 1;
  fasn_dropout_mask_kernel<<<blocks, 256, 0, stream>>>(out, B * H, Sq, Skv, thr, key, bh_offset);
  return cudaGetLastError();
}

cudaError_t launch_probe(int mode, bool bf16, const CUtensorMap& tx, const CUtensorMap& ty, const void* x, float* c,
                         cudaStream_t stream) {
  constexpr int smem = 1024 + 4 * 128 * 128 + 1024 + 2 * 4096;
  cudaError_t e;
  if (bf16) {
    e = cudaFuncSetAttribute(fasn_probe_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return e;
    fasn_probe_kernel<true><<<1, 160, smem, stream>>>(tx, ty, mode, reinterpret_cast<const uint16_t*>(x), c);
  } else {
    e = cudaFuncSetAttribute(fasn_probe_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return e;
    fasn_probe_kernel<false><<<1, 160, smem, stream>>>(tx, ty, mode, reinterpret_cast<const uint16_t*>(x), c);
  }
  return cudaGetLastError();
}

}  // namespace fasn
