/* fasn.h -- C ABI of libfasn.so: fused attention with softmax_n for NVIDIA B200 (sm_100a).
 *
 * This is the drop-in boundary for ONE hot path of softmax1/Flash-Attention-Softmax-N (reference paths are
 * relative to the reference repository root):
 *
 *   fasn_fwd  replaces  flash_attention_n(query, key, value, softmax_n_param, scale, dropout_p, attn_mask,
 *                       attn_bias, is_causal)           flash_attention_softmax_n/core/flash_attn.py:42-124
 *             and       _FlashAttentionN.forward        flash_attention_softmax_n/core/flash_attn_triton.py:243-299
 *                       (_fwd_kernel                    flash_attention_softmax_n/core/flash_attn_triton.py:30-126)
 *   fasn_bwd  replaces  _FlashAttentionN.backward       flash_attention_softmax_n/core/flash_attn_triton.py:301-336
 *                       (_bwd_preprocess :129-143, _bwd_kernel :146-235) and the aten autograd of the SDPA call
 *                       at flash_attn.py:115-124.
 *
 * The reference has no native code, so there is no existing FFI to mirror: these entry points are what a
 * ctypes/cffi binding inside flash_attention_softmax_n/core/flash_attn.py binds.
 *
 * Conventions
 *   - plain pointers and sizes only; all tensor pointers are DEVICE pointers unless the name says host;
 *   - tensors are (B, H, S, D) with arbitrary element strides for B, H, S and unit stride for D;
 *     strides are in ELEMENTS; base pointers and byte strides must be multiples of 16 bytes;
 *   - every function returns 0 on success, a negative FASN_E* code for a rejected argument, or a positive
 *     cudaError_t; fasn_last_error() returns a thread-local description of the last failure;
 *   - launches are asynchronous on `stream`; nothing here synchronises the device;
 *   - re-entrant.  Process-wide state, all mutex-guarded: the driver entry point looked up once (immutable), the
 *     fasn_profile event lists, the device arena of fasn_attention_host (one per device, grown on demand) and, per device,
 *     a pool of work counters for the persistent kernels (one pair per launch, handed out round-robin; every kernel
 *     leaves its pair at zero).
 */
#ifndef FASN_H_
#define FASN_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FASN_ABI_VERSION 3

enum { FASN_FP16 = 0, FASN_BF16 = 1 };

enum {
  FASN_EINVAL = -1,      /* null pointer, bad size, bad struct_size                       */
  FASN_EUNSUPPORTED = -2,/* head dim / dtype / alignment outside what the kernels handle   */
  FASN_EDRIVER = -3      /* cuTensorMapEncodeTiled unavailable or failed                   */
};

/* One (B, H, S, D) tensor view. */
typedef struct FasnTensor {
  void*   ptr;        /* device pointer to element (0,0,0,0)                                            */
  int64_t stride_b;   /* element stride between batches                                                 */
  int64_t stride_h;   /* element stride between heads (ignored for K/V when heads_kv == 1)              */
  int64_t stride_s;   /* element stride between sequence positions (>= D)                               */
} FasnTensor;

/* Optional dense mask / bias views, element (b,h,i,j) at ptr[b*stride_b + h*stride_h + i*stride_q + j];
 * a stride of 0 broadcasts that axis (flash_attn.py:88-89 expands the mask, :100-103 the bias). */
typedef struct FasnAux {
  const void* ptr;    /* NULL = absent                                                                  */
  int64_t stride_b, stride_h, stride_q;
} FasnAux;

typedef struct FasnParams {
  uint32_t struct_size;   /* sizeof(FasnParams), checked                                               */
  uint32_t dtype;         /* FASN_FP16 | FASN_BF16 : element type of q,k,v,o,dout,dq,dk,dv and bias     */
  int32_t  batch;         /* B                                                                          */
  int32_t  heads;         /* H  (query heads)                                                           */
  int32_t  heads_kv;      /* H, or 1 when K/V are shared by all heads (3-D key/value, flash_attn.py:75-79) */
  int32_t  seqlen_q;      /* L                                                                          */
  int32_t  seqlen_kv;     /* S                                                                          */
  int32_t  head_dim;      /* E == Ev, 64 or 128                                                         */

  FasnTensor q, k, v;     /* inputs                                                                     */
  FasnTensor o;           /* fwd: output (B,H,L,D);  bwd: the forward's output (input)                  */
  float*     lse;         /* (B,H,L) fp32 contiguous: ln(n + sum_j exp s_ij).  fwd writes, bwd reads    */

  /* backward only */
  FasnTensor dout;        /* dL/dO                                                                      */
  FasnTensor dq, dk, dv;  /* outputs; dk, dv are per QUERY head (B,H,S,D) even when heads_kv == 1 (but see dk_accum) */
  float*     delta;       /* workspace 2 x (B,H,Lp) fp32 (delta, then LSE*log2e), Lp = L rounded up to 128 */
  float*     dq_accum;    /* workspace (B,H,Lp,D) fp32; fasn_bwd zero-fills it itself                   */

  float    softmax_n;     /* n >= 0 (real valued: superset of the reference's integer zero-pad trick)   */
  float    scale;         /* logit scale; the caller resolves the 1/sqrt(E) default (flash_attn.py:59)  */
  int32_t  is_causal;     /* bottom-right aligned: row i sees keys j <= i + (S - L) (flash_attn.py:38-39) */
  float    dropout_p;     /* in [0,1); kept probabilities are scaled by 1/(1-p)                         */
  uint64_t philox_seed;   /* dropout stream; the keep mask is a pure function of                        */
  uint64_t philox_offset; /*   (seed, offset, bh_offset + b*H + h, i, j)                                */
  int64_t  bh_offset;     /* global index of this call's first (batch, head) unit (multi-GPU sharding)  */

  FasnAux  mask;          /* uint8/bool, nonzero = attend (flash_attn.py:61,70)                         */
  FasnAux  bias;          /* same dtype as q; added after scaling (flash_attn.py:81-83,100-113)         */

  void*    stream;        /* cudaStream_t                                                               */

  /* ALiBi generated in the kernels instead of a dense (H,L,S) attn_bias tensor:
   * NULL, or H fp32 slopes on the device; the logit of (i, j) gets  + alibi_slopes[h] * (j - i - (S - L)),
   * i.e. slope times the signed distance to the bottom-right aligned diagonal.  Exclusive with `bias`. */
  const float* alibi_slopes;

  /* backward only, optional: gradient with respect to the logits dS = P o (dP - delta) (dropout scaling included), i.e. the
   * gradient of a dense `bias`, written as (B,H,L,S) elements of the I/O dtype at dbias[b*stride_b + h*stride_h + i*stride_q + j].
   * Entries no CTA visits (above the causal diagonal) are not written: the caller zero-fills.  The caller reduces over the
   * axes its bias broadcasts (the SDPA route of the reference gives this gradient through aten autograd, flash_attn.py:100-124).
   * Requires `bias` or `alibi_slopes` or a dense mask (the dense-tensor kernels); NULL = not wanted. */
  void*   dbias;
  int64_t dbias_stride_b, dbias_stride_h, dbias_stride_q;

  /* forward, debug library only (libfasn_debug32.so, built with -DFASN_DEBUG_FP32_P=1: P as two 16-bit terms, float32 output):
   * contiguous (B,H,L,D) float32 copy of the output, or NULL.  The product library rejects a non-NULL value. */
  float*  o_f32;

  /* backward only, optional, heads_kv == 1 (K/V shared by all heads): float32 accumulators (B,1,S,D), contiguous, ZERO-FILLED BY THE
   * CALLER.  When both are given the kernel adds every query head's dK / dV (already scaled) into them with vector reductions at
   * the L2 instead of writing per-head (B,H,S,D) tensors for the caller to sum; `dk` / `dv` are then not written and may be null. */
  float*  dk_accum;
  float*  dv_accum;
} FasnParams;

/* ABI version of the loaded library (== FASN_ABI_VERSION it was built with). */
int fasn_version(void);

/* Thread-local description of the most recent failure on the calling thread ("" if none). */
const char* fasn_last_error(void);

/* O = dropout(softmax_n(scale Q K^T + bias, masked)) V ; also writes lse. */
int fasn_fwd(const FasnParams* p);

/* dQ, dK, dV from (Q, K, V, O, dO, lse); regenerates the forward's dropout mask from the philox fields. */
int fasn_bwd(const FasnParams* p);

/* Bytes the caller must provide for FasnParams.delta and FasnParams.dq_accum. */
int fasn_bwd_workspace(const FasnParams* p, uint64_t* delta_bytes, uint64_t* dq_accum_bytes);

/* Kernel timing for roofline reports: while enabled, fasn_fwd and fasn_bwd bracket their tensor-core kernel
 * (the forward kernel / the main backward kernel) with CUDA events on the caller's stream.
 * fasn_profile_read returns the summed durations (ms) and launch counts since the last read and clears them;
 * the caller must have synchronised the stream(s) first. */
int fasn_profile(int enable);

int fasn_profile_read(double* fwd_ms, int32_t* fwd_launches, double* bwd_ms, int32_t* bwd_launches);

/* Test hook: write the dropout keep mask the kernels use, as (B,H,L,S) uint8 (1 = keep), to `out`. */
int fasn_dropout_mask(uint8_t* out, int32_t batch, int32_t heads, int32_t seqlen_q, int32_t seqlen_kv,
                      float dropout_p, uint64_t philox_seed, uint64_t philox_offset, int64_t bh_offset,
                      void* stream);

/* Bring-up hook: one 128x128x128 tcgen05 MMA through the same descriptor builders the kernels use.
 *   mode 0: C = X * Y^T  (A, B K-major in shared memory)         -- the Q K^T form
 *   mode 1: C = X * Y    (A from tensor memory, B MN-major)      -- the P V form
 *   mode 2: C = X^T * Y  (A, B MN-major in shared memory)        -- the dQ = dS K form
 *   mode 3: C = X * Y    (A K-major, B MN-major in shared memory)-- the dK = dS^T Q form
 *   mode 4: mode 1 with A staged the way the backward kernel's compute warps do it: read from tensor memory as fp32 with
 *           tcgen05.ld.16x256b, packed to 16-bit pairs, stored with tcgen05.st.16x128b.
 * x, y: 128x128 row-major 16-bit device arrays; c: 128x128 row-major fp32. */
int fasn_probe(int mode, uint32_t dtype, const void* x, const void* y, float* c, void* stream);

/* Standalone fused softmax_n over the last (contiguous) axis: y_i = exp(x_i) / (n + sum_j exp(x_j)) for `rows` rows of
 * `cols` elements, row strides in elements.  Replaces the eager `softmax_n`
 * (flash_attention_softmax_n/core/functional.py:15-29: four elementwise passes) where a model cannot use the fused
 * attention (the reference's surgery: surgery_functions/_bert.py:101, _xlnet.py:62).  dtype codes: FASN_FP16,
 * FASN_BF16, FASN_FP32; (in, out) pairs: equal types, 16-bit -> fp32, fp32 -> 16-bit.  A row without any finite entry
 * gives 0 (the reference gives NaN for n = 0).  Backward: dx_i = y_i (dy_i - sum_j y_j dy_j); y and dy in the forward's
 * output dtype, dx in its input dtype. */
#define FASN_FP32 2u
int fasn_softmax_n_fwd(const void* x, void* y, int64_t rows, int32_t cols, int64_t x_row_stride, int64_t y_row_stride,
                       uint32_t dtype_in, uint32_t dtype_out, float n, void* stream);
int fasn_softmax_n_bwd(const void* y, const void* dy, void* dx, int64_t rows, int32_t cols, int64_t y_row_stride,
                       int64_t dy_row_stride, int64_t dx_row_stride, uint32_t dtype_in, uint32_t dtype_out, void* stream);

/* Asynchronous copy of `bytes` between any two device (or pinned host) buffers on `stream`: cudaMemcpyAsync with
 * cudaMemcpyDefault.  parallel.py moves (batch, head) slabs between GPUs with it, through peer memory mapped by CUDA IPC
 * handles: the copy engines of the issuing device do the transfer over NVLink (measured 790 GB/s per direction between two
 * B200s; a framework-level tensor copy into IPC-mapped memory took a 35 GB/s path). */
int fasn_copy_async(void* dst, const void* src, uint64_t bytes, void* stream);

/* Host-buffer convenience used for end-to-end measurement: copies q,k,v (and dout) from HOST memory,
 * runs fwd (+bwd when dout_host != NULL) and copies o (and dq,dk,dv) back.  Contiguous (B,H,S,D) layouts.
 * Uses an internal device arena sized on first use; synchronises `stream` before returning. */
int fasn_attention_host(uint32_t dtype, int32_t batch, int32_t heads, int32_t seqlen_q, int32_t seqlen_kv,
                        int32_t head_dim, const void* q_host, const void* k_host, const void* v_host,
                        void* o_host, const void* dout_host, void* dq_host, void* dk_host, void* dv_host,
                        float softmax_n, float scale, int32_t is_causal, float dropout_p,
                        uint64_t philox_seed, uint64_t philox_offset, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* FASN_H_ */
