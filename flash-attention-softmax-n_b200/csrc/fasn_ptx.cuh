// Thin inline-PTX layer for sm_100a: mbarrier, TMA (cp.async.bulk.tensor), tcgen05 (MMA / TMEM ld,st /
// alloc / commit / fences), UMMA shared-memory and instruction descriptors.
// Everything here is hand-written against the PTX ISA; no CUTLASS/CuTe types are used.
#pragma once
#include <cstdint>
#include <cuda.h>
#include <cuda_runtime.h>

namespace fasn {

#define FASN_DEVICE __device__ __forceinline__

// ------------------------------------------------------------------------------------------------
// misc
// ------------------------------------------------------------------------------------------------
FASN_DEVICE uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

FASN_DEVICE bool elect_one() {
  uint32_t pred = 0;
  asm volatile(
      "{\n\t.reg .pred P;\n\t"
      "elect.sync _|P, 0xffffffff;\n\t"
      "selp.b32 %0, 1, 0, P;\n\t}\n"
      : "=r"(pred));
  return pred != 0;
}

template <int N> FASN_DEVICE void setmaxnreg_inc() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;\n" ::"n"(N)); }
template <int N> FASN_DEVICE void setmaxnreg_dec() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;\n" ::"n"(N)); }

FASN_DEVICE void named_bar_sync(uint32_t id, uint32_t nthreads) {
  asm volatile("bar.sync %0, %1;\n" ::"r"(id), "r"(nthreads) : "memory");
}

FASN_DEVICE float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;\n" : "=f"(y) : "f"(x));
  return y;
}
FASN_DEVICE float lg2(float x) {
  float y;
  asm("lg2.approx.ftz.f32 %0, %1;\n" : "=f"(y) : "f"(x));
  return y;
}

// 2^x for a pair of finite x <= ~100 on the FMA/ALU pipes instead of the MUFU (16 ex2/clk/SM is the scarcest
// resource of the softmax): round-to-nearest range reduction with the 1.5*2^23 magic constant, degree-3 minimax
// polynomial for 2^f on [-0.5, 0.5] (max relative error 7.5e-5, below the 16-bit precision P is rounded to),
// exponent inserted with one integer multiply-add.  Inputs below -126 return ~1e-38 (never exactly 0): callers
// use it only on tiles without masked (-inf) scores.
FASN_DEVICE float2 exp2_poly_pair(float2 x) {
  const float2 magic = make_float2(12582912.f, 12582912.f);
  x.x = fmaxf(x.x, -126.f);
  x.y = fmaxf(x.y, -126.f);
  const float2 t = __fadd2_rn(x, magic);
  const float2 n = __fadd2_rn(t, make_float2(-12582912.f, -12582912.f));
  const float2 f = __fadd2_rn(x, make_float2(-n.x, -n.y));
  float2 p = __ffma2_rn(f, make_float2(0.055171654f, 0.055171654f), make_float2(0.24261113f, 0.24261113f));
  p = __ffma2_rn(p, f, make_float2(0.69326097f, 0.69326097f));
  p = __ffma2_rn(p, f, make_float2(0.99992806f, 0.99992806f));
  float2 r;
  r.x = __int_as_float(__float_as_int(t.x) * 0x800000 + __float_as_int(p.x));
  r.y = __int_as_float(__float_as_int(t.y) * 0x800000 + __float_as_int(p.y));
  return r;
}

// 3-input max (one FMNMX3 on sm_100)
FASN_DEVICE float fmax3(float a, float b, float c) {
  float y;
  asm("max.f32 %0, %1, %2, %3;\n" : "=f"(y) : "f"(a), "f"(b), "f"(c));
  return y;
}

// Transpose a 32x32 bit matrix held one row per lane: afterwards bit L of lane j's word = bit j of lane L's
// input word.  Five butterfly steps (shfl.bfly + select) instead of 32 broadcasts.
FASN_DEVICE uint32_t warp_transpose_bits(uint32_t x, int lane) {
#pragma unroll
  for (int i = 0; i < 5; ++i) {
    const int k = 16 >> i;
    const uint32_t m = (k == 16) ? 0x0000FFFFu : (k == 8) ? 0x00FF00FFu : (k == 4) ? 0x0F0F0F0Fu : (k == 2) ? 0x33333333u : 0x55555555u;
    const uint32_t y = __shfl_xor_sync(0xffffffffu, x, k);
    x = (lane & k) ? ((x & ~m) | ((y >> k) & m)) : ((x & m) | ((y & m) << k));
  }
  return x;
}

// Mask for a packed pair of 16-bit values from two adjacent keep bits: 0xFFFF in the low half iff bit `pos` of w is
// set, 0xFFFF in the high half iff bit `pos + 1` is set (pos even, compile-time after unrolling).  w << t puts bit
// 8 b + 7 - t in the sign position of byte b; PRMT with the sign-replicate flag (selector | 8) expands those sign
// bits to whole bytes, so the mask costs one PRMT (the shifted copies of w are shared by the whole 32-key group).
FASN_DEVICE uint32_t keep_pair_mask(uint32_t w, int pos) {
  const int b = pos >> 3;                        // both bits live in byte b
  const uint32_t x_lo = w << (7 - (pos & 7));    // bit pos     -> sign of byte b
  const uint32_t x_hi = w << (6 - (pos & 7));    // bit pos + 1 -> sign of byte b
  const uint32_t sel = (uint32_t)(b | 8) | ((uint32_t)(b | 8) << 4) | ((uint32_t)((4 + b) | 8) << 8) | ((uint32_t)((4 + b) | 8) << 12);
  uint32_t m;
  asm("prmt.b32 %0, %1, %2, %3;\n" : "=r"(m) : "r"(x_lo), "r"(x_hi), "r"(sel));
  return m;
}

// The same for keep words that were shifted so that the wanted bit is the sign bit of byte b of x_lo (low half of the
// pair) and of x_hi (high half): the quad layout of the backward kernel, where byte b <-> kv row 8 b + lane / 4.
FASN_DEVICE uint32_t keep_byte_pair_mask(uint32_t x_lo, uint32_t x_hi, int b) {
  const uint32_t sel = (uint32_t)(b | 8) | ((uint32_t)(b | 8) << 4) | ((uint32_t)((4 + b) | 8) << 8) | ((uint32_t)((4 + b) | 8) << 12);
  uint32_t m;
  asm("prmt.b32 %0, %1, %2, %3;\n" : "=r"(m) : "r"(x_lo), "r"(x_hi), "r"(sel));
  return m;
}

// pack two fp32 into one 32-bit word of two 16-bit floats; `lo` lands in bits [0,16)
template <bool BF16> FASN_DEVICE uint32_t pack2(float lo, float hi) {
  uint32_t r;
  if constexpr (BF16) asm("cvt.rn.bf16x2.f32 %0, %1, %2;\n" : "=r"(r) : "f"(hi), "f"(lo));
  else                asm("cvt.rn.f16x2.f32 %0, %1, %2;\n" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}

template <bool BF16> FASN_DEVICE float cvt16_to_f32(uint16_t h) {
  if constexpr (BF16) return __uint_as_float(static_cast<uint32_t>(h) << 16);
  else {
    float f;
    asm("cvt.f32.f16 %0, %1;\n" : "=f"(f) : "h"(h));
    return f;
  }
}

// ------------------------------------------------------------------------------------------------
// mbarrier
// ------------------------------------------------------------------------------------------------
FASN_DEVICE void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(count));
}
FASN_DEVICE void fence_mbar_init() {
  asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
}
FASN_DEVICE void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory"); }

FASN_DEVICE void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(smem_u32(bar)) : "memory");
}
FASN_DEVICE void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
FASN_DEVICE bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred P;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2, %3;\n\t"
      "selp.b32 %0, 1, 0, P;\n\t}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity), "r"(0x989680u)   // suspend-time hint: sleep in hardware instead of polling
      : "memory");
  return ok != 0;
}
// Wait until the phase with the given parity has completed.  A watchdog turns a protocol bug into a
// trap (launch failure) instead of a hung GPU: every legal wait in these kernels is bounded by a few tile
// times (microseconds), 2^22 polls of a HW-suspending try_wait is seconds.
#ifndef FASN_WATCHDOG_POLLS
#define FASN_WATCHDOG_POLLS (1u << 22)
#endif
FASN_DEVICE void mbar_wait(uint64_t* bar, uint32_t parity) {
#ifndef FASN_NO_WATCHDOG
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if (++spins > FASN_WATCHDOG_POLLS) { __trap(); }
  }
#else
  while (!mbar_try_wait(bar, parity)) {}
#endif
}

// ------------------------------------------------------------------------------------------------
// TMA
// ------------------------------------------------------------------------------------------------
FASN_DEVICE void tma_prefetch_desc(const CUtensorMap* m) {
  asm volatile("prefetch.tensormap [%0];\n" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
// 4-D tiled load global -> shared, completion on an mbarrier (complete_tx::bytes)
FASN_DEVICE void tma_load_4d(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];\n" ::
          "r"(smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
// 4-D tiled store shared -> global (bulk async-group completion)
FASN_DEVICE void tma_store_4d(const CUtensorMap* m, const void* smem_src, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];\n" ::"l"(
          reinterpret_cast<uint64_t>(m)),
      "r"(smem_u32(smem_src)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
FASN_DEVICE void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;\n" ::: "memory"); }
FASN_DEVICE void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;\n" ::: "memory"); }
FASN_DEVICE void tma_store_wait_read_all() { asm volatile("cp.async.bulk.wait_group.read 0;\n" ::: "memory"); }

// ------------------------------------------------------------------------------------------------
// tcgen05: TMEM allocation
// ------------------------------------------------------------------------------------------------
template <uint32_t COLS> FASN_DEVICE void tmem_alloc(uint32_t* smem_result) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(smem_result)), "n"(COLS)
               : "memory");
}
FASN_DEVICE void tmem_relinquish() { asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory"); }
template <uint32_t COLS> FASN_DEVICE void tmem_dealloc(uint32_t taddr) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(taddr), "n"(COLS) : "memory");
}

FASN_DEVICE void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory"); }
FASN_DEVICE void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory"); }
FASN_DEVICE void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory"); }
FASN_DEVICE void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;\n" ::: "memory"); }

// All previously issued tcgen05.mma of this thread arrive (once) on `bar` when they complete.
// Implies tcgen05.fence::before_thread_sync.
FASN_DEVICE void tc_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(smem_u32(bar)) : "memory");
}

// ------------------------------------------------------------------------------------------------
// tcgen05.ld / st, shape 32x32b: thread i of the warp <-> TMEM lane (base_lane + i), N consecutive columns
// ------------------------------------------------------------------------------------------------
FASN_DEVICE void tmem_ld_x32(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
// shape 16x256b.x8: 16 TMEM lanes x 64 columns.  Thread t holds, for column chunk c = 0..7 (8 columns each):
//   r[4c], r[4c+1] = lane (base + t/4),     columns 8c + 2(t%4), +1
//   r[4c+2], r[4c+3] = lane (base + t/4 + 8), the same columns
// i.e. the four threads of a quad hold 32 contiguous bytes of one row (one L2 sector).
FASN_DEVICE void tmem_ld_16x256b_x8(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.16x256b.x8.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
// shape 16x128b.x8 store: 16 TMEM lanes x 32 columns.  Thread t supplies, for n = 0..7,
//   r[2n] = lane (base + t/4), column 4n + t%4        r[2n+1] = lane (base + t/4 + 8), the same column
// which is where the packed 16-bit pairs of a 16x256b.x8 load (columns 8n + 2(t%4), +1) belong.
FASN_DEVICE void tmem_st_16x128b_x8(uint32_t taddr, const uint32_t* r) {
  asm volatile(
      "tcgen05.st.sync.aligned.16x128b.x8.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};\n" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
      "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
FASN_DEVICE void red_add_v2(float* gptr, float a, float b) {
  asm volatile("red.global.add.v2.f32 [%0], {%1, %2};\n" ::"l"(gptr), "f"(a), "f"(b) : "memory");
}
FASN_DEVICE void tmem_ld_x16(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
FASN_DEVICE void tmem_st_x32(uint32_t taddr, const uint32_t* r) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};\n" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
      "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]),
      "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]),
      "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
      : "memory");
}
FASN_DEVICE void tmem_st_x16(uint32_t taddr, const uint32_t* r) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};\n" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
      "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}

// ------------------------------------------------------------------------------------------------
// UMMA descriptors
// ------------------------------------------------------------------------------------------------
// Shared-memory matrix descriptor (64-bit), 128-byte swizzle, sm_100 version field = 1:
//   [0,14)  start address >> 4      [16,30) leading byte offset >> 4     [32,46) stride byte offset >> 4
//   [46,48) version = 1             [49,52) base offset = 0              [61,64) layout: 2 = SWIZZLE_128B
// Tiles are [rows][128 bytes] with the TMA 128B swizzle (16-byte chunk index XOR (row & 7)), 1024-byte aligned.
//   K-major operand  (rows = M/N index, 128 B = 64 K-elements):  SBO = 1024 (next 8 rows), LBO unused (1).
//       advancing K by 16 elements inside the 128-byte span = +32 bytes on the start address.
//   MN-major operand (rows = K index, 128 B = 64 M/N-elements):  SBO = 1024 (next 8 K rows),
//       LBO = byte distance to the next 64-wide M/N block; advancing K by 16 = +16 rows = +2048 bytes.
FASN_DEVICE uint64_t umma_smem_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFF) >> 4);
  d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}

// Split form for issue loops: the high word is a constant per operand kind, the low word is
// (start address >> 4) | (LBO >> 4) << 16, so stepping through a tile is one 32-bit add of (bytes >> 4).
__host__ __device__ constexpr uint32_t umma_desc_hi(uint32_t sbo_bytes) { return (sbo_bytes >> 4) | (1u << 14) | (2u << 29); }
FASN_DEVICE uint32_t umma_desc_lo(uint32_t smem_addr, uint32_t lbo_bytes) {
  return ((smem_addr & 0x3FFFF) >> 4) | ((lbo_bytes >> 4) << 16);
}
FASN_DEVICE uint64_t umma_desc_join(uint32_t lo, uint32_t hi) {
  uint64_t d;
  asm("mov.b64 %0, {%1, %2};\n" : "=l"(d) : "r"(lo), "r"(hi));
  return d;
}

// Instruction descriptor for kind::f16 (fp16/bf16 inputs, fp32 accumulate), dense, no negate:
//   [4,6) c_format = 1 (F32)   [7,10) a_format   [10,13) b_format  (0 = F16, 1 = BF16)
//   [15] a_major  [16] b_major (0 = K-major, 1 = MN-major)   [17,23) N >> 3   [24,29) M >> 4
__host__ __device__ constexpr uint32_t umma_idesc(bool bf16, uint32_t M, uint32_t N, bool a_mn_major, bool b_mn_major) {
  return (1u << 4) | ((bf16 ? 1u : 0u) << 7) | ((bf16 ? 1u : 0u) << 10) | ((a_mn_major ? 1u : 0u) << 15) |
         ((b_mn_major ? 1u : 0u) << 16) | ((N >> 3) << 17) | ((M >> 4) << 24);
}

// D[tmem] (+)= A[smem] * B[smem]
FASN_DEVICE void umma_ss(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// D[tmem] (+)= A[tmem] * B[smem]     (A: lane = row, two 16-bit K-elements per 32-bit column)
FASN_DEVICE void umma_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(tmem_d),
      "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}

// ------------------------------------------------------------------------------------------------
// global memory helpers
// ------------------------------------------------------------------------------------------------
FASN_DEVICE void red_add_v4(float* gptr, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};\n" ::"l"(gptr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

}  // namespace fasn
