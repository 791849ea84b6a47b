// Standalone fused softmax_n over the last (contiguous) axis, forward and backward -- SURVEY.md section 8(f) rank 2: the
// operator the reference's surgery injects into models that cannot use fused attention
// (flash_attention_softmax_n/core/functional.py:15-29; used by surgery/surgery_functions/_bert.py:101, _xlnet.py:62).
//
//   y_i = exp(x_i) / (n + sum_j exp(x_j))          dx_i = y_i (dy_i - sum_j y_j dy_j)
//
// The reference does four elementwise passes over the (.., L, S) scores in DRAM (max, exp, sum, divide); here a row is
// read once and written once.  The shift is max(row max, 0): the "+n" is a virtual entry with logit 0, so the shifted
// form exp(x - m) / (n exp(-m) + sum exp(x - m)) never overflows (the reference's exp(-max) does for very negative rows).
// HBM-bound: 2 x (bytes per element) per element; one warp per row up to 1024 columns, one 128-thread group per row up
// to 4096 columns (values stay in registers), beyond that an online (max, sum) pass and a second read that hits L2.
#include <cstdint>
#include <cuda_runtime.h>
#include "fasn_ptx.cuh"

namespace fasn {

namespace {

constexpr int kSmBlock = 128;
constexpr float kLog2eSm = 1.4426950408889634f;
// element type codes (include/fasn.h): 0 fp16, 1 bf16, 2 fp32

template <int DT> __device__ __forceinline__ float ld_elem(const void* p, long long i) {
  if constexpr (DT == 2) return reinterpret_cast<const float*>(p)[i];
  else return cvt16_to_f32<DT == 1>(reinterpret_cast<const uint16_t*>(p)[i]);
}
template <int DT> __device__ __forceinline__ void st_elem(void* p, long long i, float v) {
  if constexpr (DT == 2) reinterpret_cast<float*>(p)[i] = v;
  else reinterpret_cast<uint16_t*>(p)[i] = (uint16_t)(pack2<DT == 1>(v, 0.f) & 0xFFFFu);
}

// eight consecutive elements (the caller guarantees 16-byte alignment for 16-bit types, 2 x 16 bytes for fp32)
template <int DT> __device__ __forceinline__ void ld8(const void* p, long long i, float (&v)[8]) {
  if constexpr (DT == 2) {
    const float4 a = *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(p) + i);
    const float4 b = *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(p) + i + 4);
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
  } else {
    const uint4 u = *reinterpret_cast<const uint4*>(reinterpret_cast<const uint16_t*>(p) + i);
    const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
    for (int e = 0; e < 4; ++e) { v[2 * e] = cvt16_to_f32<DT == 1>((uint16_t)(w[e] & 0xFFFF)); v[2 * e + 1] = cvt16_to_f32<DT == 1>((uint16_t)(w[e] >> 16)); }
  }
}
template <int DT> __device__ __forceinline__ void st8(void* p, long long i, const float (&v)[8]) {
  if constexpr (DT == 2) {
    *reinterpret_cast<float4*>(reinterpret_cast<float*>(p) + i) = make_float4(v[0], v[1], v[2], v[3]);
    *reinterpret_cast<float4*>(reinterpret_cast<float*>(p) + i + 4) = make_float4(v[4], v[5], v[6], v[7]);
  } else {
    uint4 u;
    u.x = pack2<DT == 1>(v[0], v[1]); u.y = pack2<DT == 1>(v[2], v[3]); u.z = pack2<DT == 1>(v[4], v[5]); u.w = pack2<DT == 1>(v[6], v[7]);
    *reinterpret_cast<uint4*>(reinterpret_cast<uint16_t*>(p) + i) = u;
  }
}

// reductions over a group of G threads (32: one warp; 128: the whole block through shared memory)
template <int G> __device__ __forceinline__ float group_max(float v, float* scratch) {
#pragma unroll
  for (int s = 16; s > 0; s >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, s));
  if constexpr (G > 32) {
    const int w = threadIdx.x >> 5;
    __syncthreads();
    if ((threadIdx.x & 31) == 0) scratch[w] = v;
    __syncthreads();
    v = fmaxf(fmaxf(scratch[0], scratch[1]), fmaxf(scratch[2], scratch[3]));
  }
  return v;
}
template <int G> __device__ __forceinline__ float group_sum(float v, float* scratch) {
#pragma unroll
  for (int s = 16; s > 0; s >>= 1) v += __shfl_xor_sync(0xffffffffu, v, s);
  if constexpr (G > 32) {
    const int w = threadIdx.x >> 5;
    __syncthreads();
    if ((threadIdx.x & 31) == 0) scratch[w] = v;
    __syncthreads();
    v = (scratch[0] + scratch[1]) + (scratch[2] + scratch[3]);
  }
  return v;
}

struct SmArgs {
  const void* a;        // forward: x            backward: y
  const void* b;        // forward: unused       backward: dy
  void* out;            // forward: y            backward: dx
  long long rows;
  int cols;
  long long sa, sb, so; // row strides (elements)
  float n;
  int vec;              // 1: every row start is 16-byte aligned and cols % 8 == 0 (8-element vector accesses)
};

// ---- rows that fit in registers: G threads per row, CH chunks of 8 elements per thread (cols <= G * CH * 8)
template <int DTI, int DTO, int G, int CH, bool BWD>
__global__ void __launch_bounds__(kSmBlock)
fasn_softmax_n_reg_kernel(SmArgs p) {
  __shared__ float scratch[4];
  const int g = threadIdx.x / G, tg = threadIdx.x % G;
  const long long row = (long long)blockIdx.x * (kSmBlock / G) + g;
  const bool live = row < p.rows;                    // G == 128: uniform; G == 32: per warp
  if (G == 32 && !live) return;
  const long long ra = (live ? row : 0) * p.sa, rb = (live ? row : 0) * p.sb, ro = (live ? row : 0) * p.so;
  float v[CH][8], w[CH][8];
#pragma unroll
  for (int c = 0; c < CH; ++c) {
    const int col = (tg + c * G) * 8;
    if (p.vec && col + 8 <= p.cols) {
      ld8<BWD ? DTO : DTI>(p.a, ra + col, v[c]);
      if (BWD) ld8<DTO>(p.b, rb + col, w[c]);
    } else {
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const bool in = col + e < p.cols;
        v[c][e] = in ? ld_elem<BWD ? DTO : DTI>(p.a, ra + col + e) : (BWD ? 0.f : -INFINITY);
        if (BWD) w[c][e] = in ? ld_elem<DTO>(p.b, rb + col + e) : 0.f;
      }
    }
  }
  if constexpr (!BWD) {
    float m = 0.f;                                   // the virtual zero-logit entry of softmax_n (weight n)
    if (!(p.n > 0.f)) m = -INFINITY;
#pragma unroll
    for (int c = 0; c < CH; ++c)
#pragma unroll
      for (int e = 0; e < 8; ++e) m = fmaxf(m, v[c][e]);
    m = group_max<G>(m, scratch);
    const float ms = (m == -INFINITY) ? 0.f : m * kLog2eSm;
    float l = 0.f;
#pragma unroll
    for (int c = 0; c < CH; ++c)
#pragma unroll
      for (int e = 0; e < 8; ++e) { v[c][e] = ex2(fmaf(v[c][e], kLog2eSm, -ms)); l += v[c][e]; }
    l = group_sum<G>(l, scratch) + ((p.n > 0.f) ? p.n * ex2(-ms) : 0.f);
    const float inv = l > 0.f ? 1.f / l : 0.f;       // a row with no finite entry and n == 0 is defined as 0
#pragma unroll
    for (int c = 0; c < CH; ++c)
#pragma unroll
      for (int e = 0; e < 8; ++e) v[c][e] *= inv;
  } else {
    float d = 0.f;
#pragma unroll
    for (int c = 0; c < CH; ++c)
#pragma unroll
      for (int e = 0; e < 8; ++e) d = fmaf(v[c][e], w[c][e], d);
    d = group_sum<G>(d, scratch);
#pragma unroll
    for (int c = 0; c < CH; ++c)
#pragma unroll
      for (int e = 0; e < 8; ++e) v[c][e] *= (w[c][e] - d);
  }
  if (!live) return;
#pragma unroll
  for (int c = 0; c < CH; ++c) {
    const int col = (tg + c * G) * 8;
    if (p.vec && col + 8 <= p.cols) {
      st8<BWD ? DTI : DTO>(p.out, ro + col, v[c]);
    } else {
#pragma unroll
      for (int e = 0; e < 8; ++e)
        if (col + e < p.cols) st_elem<BWD ? DTI : DTO>(p.out, ro + col + e, v[c][e]);
    }
  }
}

// ---- long rows: one block per row, online (max, sum) pass, then a second read (L2) and the write; 8-element vector
// accesses when the rows allow it
template <int DTI, int DTO, bool BWD>
__global__ void __launch_bounds__(kSmBlock)
fasn_softmax_n_long_kernel(SmArgs p) {
  __shared__ float scratch[4];
  const long long row = blockIdx.x;
  const long long ra = row * p.sa, rb = row * p.sb, ro = row * p.so;
  const int nvec = p.vec ? p.cols / 8 : 0;            // chunks of 8 handled with vector accesses (vec => cols % 8 == 0)
  if constexpr (!BWD) {
    float m = (p.n > 0.f) ? 0.f : -INFINITY, l = 0.f;
    auto online = [&](float x) {
      x *= kLog2eSm;
      if (x > m) { l *= ex2(m - x); m = x; }          // m == -inf, x finite: l is 0, ex2(-inf) = 0
      if (x != -INFINITY) l += ex2(x - m);
    };
    for (int ch = threadIdx.x; ch < nvec; ch += kSmBlock) {
      float v[8];
      ld8<DTI>(p.a, ra + ch * 8, v);
      float cm = fmaxf(fmaxf(fmaxf(v[0], v[1]), fmaxf(v[2], v[3])), fmaxf(fmaxf(v[4], v[5]), fmaxf(v[6], v[7]))) * kLog2eSm;
      if (cm > m) { l *= ex2(m - cm); m = cm; }       // one rescale per chunk
      if (m != -INFINITY) {
#pragma unroll
        for (int e = 0; e < 8; ++e) l += ex2(fmaf(v[e], kLog2eSm, -m));
      }
    }
    for (int col = nvec * 8 + threadIdx.x; col < p.cols; col += kSmBlock) online(ld_elem<DTI>(p.a, ra + col));
    const float mm = group_max<kSmBlock>(m, scratch);
    const float ms = (mm == -INFINITY) ? 0.f : mm;
    l = (m == -INFINITY) ? 0.f : l * ex2(m - ms);
    l = group_sum<kSmBlock>(l, scratch) + ((p.n > 0.f) ? p.n * ex2(-ms) : 0.f);
    const float inv = l > 0.f ? 1.f / l : 0.f;
    for (int ch = threadIdx.x; ch < nvec; ch += kSmBlock) {
      float v[8];
      ld8<DTI>(p.a, ra + ch * 8, v);
#pragma unroll
      for (int e = 0; e < 8; ++e) v[e] = ex2(fmaf(v[e], kLog2eSm, -ms)) * inv;
      st8<DTO>(p.out, ro + ch * 8, v);
    }
    for (int col = nvec * 8 + threadIdx.x; col < p.cols; col += kSmBlock)
      st_elem<DTO>(p.out, ro + col, ex2(fmaf(ld_elem<DTI>(p.a, ra + col), kLog2eSm, -ms)) * inv);
  } else {
    float d = 0.f;
    for (int ch = threadIdx.x; ch < nvec; ch += kSmBlock) {
      float v[8], w[8];
      ld8<DTO>(p.a, ra + ch * 8, v);
      ld8<DTO>(p.b, rb + ch * 8, w);
#pragma unroll
      for (int e = 0; e < 8; ++e) d = fmaf(v[e], w[e], d);
    }
    for (int col = nvec * 8 + threadIdx.x; col < p.cols; col += kSmBlock) d = fmaf(ld_elem<DTO>(p.a, ra + col), ld_elem<DTO>(p.b, rb + col), d);
    d = group_sum<kSmBlock>(d, scratch);
    for (int ch = threadIdx.x; ch < nvec; ch += kSmBlock) {
      float v[8], w[8];
      ld8<DTO>(p.a, ra + ch * 8, v);
      ld8<DTO>(p.b, rb + ch * 8, w);
#pragma unroll
      for (int e = 0; e < 8; ++e) v[e] *= (w[e] - d);
      st8<DTI>(p.out, ro + ch * 8, v);
    }
    for (int col = nvec * 8 + threadIdx.x; col < p.cols; col += kSmBlock)
      st_elem<DTI>(p.out, ro + col, ld_elem<DTO>(p.a, ra + col) * (ld_elem<DTO>(p.b, rb + col) - d));
  }
}

template <int DTI, int DTO, bool BWD>
cudaError_t launch_sm_t(const SmArgs& p, cudaStream_t st) {
  if (p.rows <= 0 || p.cols <= 0) return cudaSuccess;
  const int chunks = (p.cols + 7) / 8;
  if (chunks <= 32 * 4) {                 // <= 1024 columns: one warp per row
    const unsigned blocks = (unsigned)((p.rows + 3) / 4);
    if (chunks <= 32) fasn_softmax_n_reg_kernel<DTI, DTO, 32, 1, BWD><<<blocks, kSmBlock, 0, st>>>(p);
    else if (chunks <= 64) fasn_softmax_n_reg_kernel<DTI, DTO, 32, 2, BWD><<<blocks, kSmBlock, 0, st>>>(p);
    else fasn_softmax_n_reg_kernel<DTI, DTO, 32, 4, BWD><<<blocks, kSmBlock, 0, st>>>(p);
  } else if (chunks <= 128 * 4) {         // <= 4096 columns: one 128-thread block per row, values in registers
    if (chunks <= 256) fasn_softmax_n_reg_kernel<DTI, DTO, 128, 2, BWD><<<(unsigned)p.rows, kSmBlock, 0, st>>>(p);
    else fasn_softmax_n_reg_kernel<DTI, DTO, 128, 4, BWD><<<(unsigned)p.rows, kSmBlock, 0, st>>>(p);
  } else {
    fasn_softmax_n_long_kernel<DTI, DTO, BWD><<<(unsigned)p.rows, kSmBlock, 0, st>>>(p);
  }
  return cudaGetLastError();
}

template <bool BWD>
cudaError_t launch_sm(int dti, int dto, const SmArgs& p, cudaStream_t st) {
#define FASN_SM_CASE(I_, O_) if (dti == I_ && dto == O_) return launch_sm_t<I_, O_, BWD>(p, st);
  FASN_SM_CASE(0, 0) FASN_SM_CASE(0, 2) FASN_SM_CASE(1, 1) FASN_SM_CASE(1, 2) FASN_SM_CASE(2, 2) FASN_SM_CASE(2, 0) FASN_SM_CASE(2, 1)
#undef FASN_SM_CASE
  return cudaErrorInvalidValue;
}

}  // namespace

// dtype codes: 0 fp16, 1 bf16, 2 fp32.  Supported (in, out) pairs: equal types, 16-bit -> fp32, fp32 -> 16-bit.
cudaError_t launch_softmax_n_fwd(const void* x, void* y, long long rows, int cols, long long sx, long long sy, int dt_in, int dt_out,
                                 float n, int vec, cudaStream_t st) {
  SmArgs p{x, nullptr, y, rows, cols, sx, 0, sy, n, vec};
  return launch_sm<false>(dt_in, dt_out, p, st);
}
// y, dy in the OUTPUT dtype of the forward, dx in its INPUT dtype
cudaError_t launch_softmax_n_bwd(const void* y, const void* dy, void* dx, long long rows, int cols, long long sy, long long sdy,
                                 long long sdx, int dt_in, int dt_out, int vec, cudaStream_t st) {
  SmArgs p{y, dy, dx, rows, cols, sy, sdy, sdx, 0.f, vec};
  return launch_sm<true>(dt_in, dt_out, p, st);
}

}  // namespace fasn
