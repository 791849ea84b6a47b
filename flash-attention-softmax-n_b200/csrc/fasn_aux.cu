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
      if (mode == 1 || mode == 4) mbar_wait(&bars[2], 0);
      tc_fence_after();
      const uint32_t sX_u = smem_u32(sX), sY_u = smem_u32(sY);
      for (int kb = 0; kb < 8; ++kb) {
        const uint32_t koff = (kb >> 2) * BLK + (kb & 3) * 32;      // K-major advance
        const uint32_t moff = kb * 2048;                            // MN-major advance
        const uint32_t acc = kb > 0 ? 1u : 0u;
        if (mode == 0)
          umma_ss(tmem_base, umma_smem_desc(sX_u + koff, 16, 1024), umma_smem_desc(sY_u + koff, 16, 1024),
                  umma_idesc(BF16, 128, 128, false, false), acc);
        else if (mode == 1 || mode == 4)
          umma_ts(tmem_base, tmem_base + 256 + kb * 8, umma_smem_desc(sY_u + moff, BLK, 1024),
                  umma_idesc(BF16, 128, 128, false, true), acc);
        else if (mode == 2)
          umma_ss(tmem_base, umma_smem_desc(sX_u + moff, BLK, 1024), umma_smem_desc(sY_u + moff, BLK, 1024),
                  umma_idesc(BF16, 128, 128, true, true), acc);
        else
          umma_ss(tmem_base, umma_smem_desc(sX_u + koff, 16, 1024), umma_smem_desc(sY_u + moff, BLK, 1024),
                  umma_idesc(BF16, 128, 128, false, true), acc);
      }
      tc_commit(&bars[1]);
    }
  } else {
    const int r = threadIdx.x;   // 0..127
    const uint32_t lane_off = static_cast<uint32_t>(warp * 32) << 16;
    if (mode == 4) {
      // The register layout of the backward kernel's compute warps: X goes into tensor memory as fp32 (thread = row), is read
      // back with tcgen05.ld.16x256b (thread = 4 rows x 16 columns per 64-column half), packed to 16-bit pairs and stored as
      // the A operand with tcgen05.st.16x128b.  Expected: the same C = X Y as mode 1.
      const uint32_t* src = reinterpret_cast<const uint32_t*>(x + r * 128);
#pragma unroll
      for (int cb = 0; cb < 4; ++cb) {
        uint32_t f[32];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const uint32_t w = src[cb * 16 + i];
          f[2 * i] = __float_as_uint(cvt16_to_f32<BF16>((uint16_t)(w & 0xFFFF)));
          f[2 * i + 1] = __float_as_uint(cvt16_to_f32<BF16>((uint16_t)(w >> 16)));
        }
        tmem_st_x32(tmem_base + lane_off + 128 + cb * 32, f);
      }
      tmem_wait_st();
      __syncwarp();
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        uint32_t v[64];
        tmem_ld_16x256b_x8(tmem_base + lane_off + 128 + half * 64, v);
        tmem_ld_16x256b_x8(tmem_base + lane_off + (16u << 16) + 128 + half * 64, v + 32);
        tmem_wait_ld();
#pragma unroll
        for (int h16 = 0; h16 < 2; ++h16) {
          uint32_t pk[16];
#pragma unroll
          for (int cc = 0; cc < 8; ++cc)
#pragma unroll
            for (int u = 0; u < 2; ++u)
              pk[2 * cc + u] = pack2<BF16>(__uint_as_float(v[h16 * 32 + 4 * cc + 2 * u]), __uint_as_float(v[h16 * 32 + 4 * cc + 2 * u + 1]));
          tmem_st_16x128b_x8(tmem_base + lane_off + (static_cast<uint32_t>(h16 * 16) << 16) + 256 + half * 32, pk);
        }
      }
      tmem_wait_st();
      tc_fence_before();
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

}  // namespace

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
  constexpr int smem = 1024 + 4 * 128 * 128 + 1024;
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
