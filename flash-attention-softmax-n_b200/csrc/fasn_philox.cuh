// Counter-based dropout for the attention kernels.
//
// A keep decision is a pure function of (seed, offset, global batch*head unit, query row, key column), so the
// forward kernel, the backward kernel and the CPU oracle (oracle/attention_oracle.py: dropout_keep_words)
// regenerate the same mask in whatever thread layout they like, and results do not depend on how the
// batch x head axis is sharded over GPUs.
//
// One 32-bit "keep word" covers keys 32*w .. 32*w+31 of one query row.  Two Philox-4x32-7 calls (FASN_PHILOX_ROUNDS: the
// variant FlashAttention-2 uses and the smallest Crush-resistant one of Salmon et al., SC'11) give eight
// 32-bit planes; plane p carries bit (7-p) of 32 independent 8-bit uniforms (bit-sliced), and u < T is
// evaluated with a bitwise comparator (MSB first) -- 16 keep decisions per Philox call, ~0.5 ALU op per
// decision for the compare.  P(keep) = T/256 with T = round(256 (1-p)).
#pragma once
#include <cstdint>

#ifndef FASN_PHILOX_ROUNDS
#define FASN_PHILOX_ROUNDS 7
#endif

namespace fasn {

struct PhiloxKey {
  uint32_t k0, k1;   // seed
  uint32_t offset;   // per-call stream offset
  // Expanded on the host (make_philox_key) so the kernels read them straight from the constant bank as instruction
  // operands instead of keeping 28 loop-invariant values in registers:
  uint32_t rk0[10], rk1[10];   // round keys k0 + r W0, k1 + r W1
  uint32_t tmask[8];           // tmask[p] = all-ones iff bit (7-p) of the keep threshold T is set
};

inline PhiloxKey make_philox_key(uint64_t seed, uint64_t offset, uint32_t thr) {
  PhiloxKey k{};
  k.k0 = (uint32_t)(seed & 0xFFFFFFFFull);
  k.k1 = (uint32_t)(seed >> 32);
  k.offset = (uint32_t)(offset & 0xFFFFFFFFull);
  for (int r = 0; r < 10; ++r) { k.rk0[r] = k.k0 + (uint32_t)r * 0x9E3779B9u; k.rk1[r] = k.k1 + (uint32_t)r * 0xBB67AE85u; }
  for (int p = 0; p < 8; ++p) k.tmask[p] = ((thr >> (7 - p)) & 1u) ? 0xFFFFFFFFu : 0u;
  return k;
}

__device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, const PhiloxKey& key,
                                              uint32_t (&out)[4]) {
  constexpr uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u;
#pragma unroll
  for (int r = 0; r < FASN_PHILOX_ROUNDS; ++r) {
    const uint32_t hi0 = __umulhi(M0, c0), lo0 = M0 * c0;
    const uint32_t hi1 = __umulhi(M1, c2), lo1 = M1 * c2;
    c0 = hi1 ^ c1 ^ key.rk0[r];
    c1 = lo1;
    c2 = hi0 ^ c3 ^ key.rk1[r];
    c3 = lo0;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

// 32 keep bits for (unit bh, query row q, key word w).  `thr` in [0,256].
__device__ __forceinline__ uint32_t dropout_keep_word(const PhiloxKey& key, uint32_t bh, uint32_t q, uint32_t w,
                                                      uint32_t thr) {
  if (thr >= 256u) return 0xFFFFFFFFu;
  uint32_t pl[8];
  {
    uint32_t o[4];
    philox4x32_10(q, (w << 1), bh, key.offset, key, o);
    pl[0] = o[0]; pl[1] = o[1]; pl[2] = o[2]; pl[3] = o[3];
    philox4x32_10(q, (w << 1) | 1u, bh, key.offset, key, o);
    pl[4] = o[0]; pl[5] = o[1]; pl[6] = o[2]; pl[7] = o[3];
  }
  uint32_t lt = 0u, eq = 0xFFFFFFFFu;
#pragma unroll
  for (int p = 0; p < 8; ++p) {
    const uint32_t tmask = key.tmask[p];   // warp-uniform, read from the constant bank
    lt |= eq & ~pl[p] & tmask;
    eq &= ~(pl[p] ^ tmask);
  }
  return lt;
}

}  // namespace fasn
