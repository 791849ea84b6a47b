// Internal argument blocks shared by the kernels and the C-ABI translation unit.
#pragma once
#include <cstdint>
#include <cuda.h>
#include <cuda_runtime.h>
#include "fasn_philox.cuh"

namespace fasn {

constexpr float kLog2e = 1.4426950408889634f;
constexpr float kLn2 = 0.6931471805599453f;

struct AuxView {          // dense mask / bias, element (b,h,i,j) at ptr[b*sb + h*sh + i*sq + j]
  const void* ptr;
  long long sb, sh, sq;
};

struct TensorView {       // (B,H,S,D) view, element strides, unit stride on D
  void* ptr;
  long long sb, sh, ss;
};

struct FwdArgs {
  int B, H, Hkv, Sq, Skv;
  int causal_off;         // Skv - Sq (bottom-right alignment)
  float scale_log2;       // scale * log2(e)
  float softmax_n;        // n
  float* lse;             // (B,H,Sq)
  TensorView o;           // only used by the "no visible keys" early-out
  AuxView mask, bias;
  const float* alibi;     // H slopes or null: logit(i, j) += alibi[h] * (j - i - causal_off)   (generic kernels)
  uint32_t drop_thr;      // keep iff u8 < drop_thr
  float inv_keep;         // 1/(1-p)
  PhiloxKey key;
  uint32_t bh_offset;
  int sched_group;        // units scheduled tile-major at the end of the launch (decode_block)
  float* o_f32;           // debug builds (FASN_DEBUG_FP32_P) only: (B,H,Sq,D) float32 copy of the output, or null
  int* sched;             // persistent kernels: [0] next work item (beyond the first wave), [1] CTAs finished; both zero before
                          // and after every launch (the last CTA to finish resets them)
  int grid_ctas;          // CTAs to launch: min(SMs of the device, work items)
  unsigned long long* dbg;   // FASN_TIMELINE builds only
  unsigned int dbg_x, dbg_y;
};

struct BwdArgs {
  int B, H, Hkv, Sq, Skv;
  int causal_off;
  float scale;            // logit scale / (1-p): multiplies dQ and dK (dS is computed without the dropout factor)
  float scale_log2;       // scale * log2(e)
  const float* lse;       // (B,H,Sq) natural log
  const float* delta;     // workspace: [0] delta (B,H,Sqp), [1] LSE_n * log2e (B,H,Sqp), written by the prep kernel
  float* dq_accum;        // (B,H,Sqp,D) fp32
  int Sqp;                // Sq rounded up to 128
  AuxView mask, bias;
  const float* alibi;     // H slopes or null (generic kernels)
  AuxView dbias;          // optional output of the dense-tensor kernels: dS in the I/O dtype (const-ness of AuxView::ptr is cast away)
  float* dk_accum;        // shared K/V (Hkv == 1), optional: (B,1,Skv,D) fp32 accumulators the kernel adds every head's dK / dV into
  float* dv_accum;        //   (zero-filled by the caller); null = per-head dK / dV tensors through TMA stores
  uint32_t drop_thr;
  float inv_keep;         // 1 / P(keep)
  float keep_prob;        // P(keep) = T / 256
  PhiloxKey key;
  uint32_t bh_offset;
  int sched_group;        // units scheduled tile-major at the end of the launch (decode_block)
  int* sched;             // work counters of the persistent kernel (see FwdArgs::sched)
  int grid_ctas;          // CTAs to launch: min(SMs of the device, work items)
  unsigned long long* dbg;   // FASN_TIMELINE builds only: phase timeline buffer (see fasn_bwd.cu)
  unsigned int dbg_x, dbg_y;
};

// Work order of the tensor-core kernels.  The grid is one-dimensional; CTA `lin` works on tile `tile` of unit `bh`.
// Units are taken one after the other with all their tiles adjacent (tiles of a unit start together and stream the same
// K/V or Q/dO tiles in lockstep, so one DRAM fetch serves them all through L2) -- except for the last `tail` units, which
// are taken tile-major: with a causal mask the work per tile falls with the tile index (heavy tiles are numbered first),
// so the launch ends with the lightest tiles of many units instead of the heaviest tiles of the last unit.
// List-scheduling simulation at S=4096 on 148 SMs: 3.5 % of the backward and 5 % of the forward; measured 7 % on the
// C3 forward.  `tail` is sized on the host so that the tail's operands stay within ~96 MB (1..32 units).
struct TileCoord { int tile, bh; };
__device__ __forceinline__ TileCoord decode_block(uint32_t lin, int ntiles, int BH, int tail) {
  tail = min(tail, BH);
  const uint32_t head = (uint32_t)(BH - tail) * (uint32_t)ntiles;
  TileCoord c;
  if (lin < head) {
    c.bh = (int)(lin / (uint32_t)ntiles);
    c.tile = (int)(lin - (uint32_t)c.bh * (uint32_t)ntiles);
  } else {
    const uint32_t rem = lin - head;
    c.tile = (int)(rem / (uint32_t)tail);
    c.bh = (BH - tail) + (int)(rem - (uint32_t)c.tile * (uint32_t)tail);
  }
  return c;
}

inline int sched_group_size(long long unit_bytes) {
  long long g = (96ll << 20) / (unit_bytes > 0 ? unit_bytes : 1);
  return (int)(g < 1 ? 1 : (g > 32 ? 32 : g));
}

// launchers implemented in the kernel translation units
cudaError_t launch_fwd(int head_dim, bool bf16, bool causal, bool dropout, const CUtensorMap& tq, const CUtensorMap& tk,
                       const CUtensorMap& tv, const CUtensorMap& to, const FwdArgs& a, cudaStream_t stream);

cudaError_t launch_bwd_prep(int head_dim, bool bf16, const TensorView& o, const TensorView& dout, const BwdArgs& a,
                            cudaStream_t stream);
cudaError_t launch_bwd(int head_dim, bool bf16, bool causal, bool dropout, const CUtensorMap& tq, const CUtensorMap& tk,
                       const CUtensorMap& tv, const CUtensorMap& tdo, const CUtensorMap& tdk, const CUtensorMap& tdv,
                       const CUtensorMap& tdq_accum, const BwdArgs& a, const TensorView& dk, const TensorView& dv,
                       cudaStream_t stream);
cudaError_t launch_bwd_finish(int head_dim, bool bf16, const TensorView& dq, const BwdArgs& a, cudaStream_t stream);

// standalone softmax_n over the last axis (fasn_softmax.cu); dtype codes 0 fp16, 1 bf16, 2 fp32
cudaError_t launch_softmax_n_fwd(const void* x, void* y, long long rows, int cols, long long sx, long long sy, int dt_in, int dt_out,
                                 float n, int vec, cudaStream_t st);
cudaError_t launch_softmax_n_bwd(const void* y, const void* dy, void* dx, long long rows, int cols, long long sy, long long sdy,
                                 long long sdx, int dt_in, int dt_out, int vec, cudaStream_t st);
cudaError_t launch_dropout_mask(uint8_t* out, int B, int H, int Sq, int Skv, uint32_t thr, PhiloxKey key,
                                uint32_t bh_offset, cudaStream_t stream);
cudaError_t launch_probe(int mode, bool bf16, const CUtensorMap& tx, const CUtensorMap& ty, const void* x, float* c,
                         cudaStream_t stream);

}  // namespace fasn
