// Backward helper kernels (HBM-bound, elementwise):
//   prep   : delta_i = sum_d O_id dO_id  (the reference's `_bwd_preprocess`, flash_attn_triton.py:129-143),
//            LSE2 = LSE_n * log2(e) padded to a multiple of 128 rows (so the main kernel can bulk-copy it),
//            and zero-fill of the fp32 dQ accumulator that the main kernel reduces into;
//   finish : dQ = scale * dq_accum, cast to the I/O dtype (the reference returns the fp32 buffer as is,
//            flash_attn_triton.py:312,336; we return dQ in the input dtype like the SDPA route does).
#include "fasn_common.cuh"
#include "fasn_ptx.cuh"

namespace fasn {

namespace {

template <int D, bool BF16>
__global__ void __launch_bounds__(256)
fasn_bwd_prep_kernel(TensorView o, TensorView dout, BwdArgs a) {
  // one warp per (bh, row); rows in [Sq, Sqp) are padding and get delta = 0
  const int lane = threadIdx.x & 31;
  const long long warp_global = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const long long total = (long long)a.B * a.H * a.Sqp;
  if (warp_global >= total) return;
  const int row = (int)(warp_global % a.Sqp);
  const int bh = (int)(warp_global / a.Sqp);
  const int b = bh / a.H, h = bh - b * a.H;
  float acc = 0.f;
  if (row < a.Sq) {
    const uint16_t* po = reinterpret_cast<const uint16_t*>(o.ptr) + b * o.sb + h * o.sh + (long long)row * o.ss;
    const uint16_t* pd = reinterpret_cast<const uint16_t*>(dout.ptr) + b * dout.sb + h * dout.sh + (long long)row * dout.ss;
    constexpr int PER = D / 32;   // elements per lane: 4 (D=128) or 2 (D=64)
    if constexpr (PER == 4) {
      const uint2 vo = *reinterpret_cast<const uint2*>(po + lane * 4);
      const uint2 vd = *reinterpret_cast<const uint2*>(pd + lane * 4);
      acc += cvt16_to_f32<BF16>(vo.x & 0xFFFF) * cvt16_to_f32<BF16>(vd.x & 0xFFFF);
      acc += cvt16_to_f32<BF16>(vo.x >> 16) * cvt16_to_f32<BF16>(vd.x >> 16);
      acc += cvt16_to_f32<BF16>(vo.y & 0xFFFF) * cvt16_to_f32<BF16>(vd.y & 0xFFFF);
      acc += cvt16_to_f32<BF16>(vo.y >> 16) * cvt16_to_f32<BF16>(vd.y >> 16);
    } else {
      const uint32_t vo = *reinterpret_cast<const uint32_t*>(po + lane * 2);
      const uint32_t vd = *reinterpret_cast<const uint32_t*>(pd + lane * 2);
      acc += cvt16_to_f32<BF16>(vo & 0xFFFF) * cvt16_to_f32<BF16>(vd & 0xFFFF);
      acc += cvt16_to_f32<BF16>(vo >> 16) * cvt16_to_f32<BF16>(vd >> 16);
    }
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, s);
  }
  if (lane == 0) {
    float* ws = const_cast<float*>(a.delta);
    ws[(long long)bh * a.Sqp + row] = acc * a.keep_prob;    // (1-p) delta: see the dS' formulation in fasn_bwd.cu
    // second half of the workspace: LSE_n in the log2 domain, +inf on padding rows (=> P = 0 there)
    ws[(long long)a.B * a.H * a.Sqp + (long long)bh * a.Sqp + row] =
        (row < a.Sq) ? a.lse[(long long)bh * a.Sq + row] * kLog2e : INFINITY;
  }
  float* acc_row = a.dq_accum + ((long long)bh * a.Sqp + row) * D;
  if constexpr (D == 128) reinterpret_cast<float4*>(acc_row)[lane] = make_float4(0.f, 0.f, 0.f, 0.f);
  else                    reinterpret_cast<float2*>(acc_row)[lane] = make_float2(0.f, 0.f);
}

template <int D, bool BF16>
__global__ void __launch_bounds__(256)
fasn_bwd_finish_kernel(TensorView dq, BwdArgs a) {
  // one thread per 8 output elements (16 bytes)
  constexpr int VPR = D / 8;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long total = (long long)a.B * a.H * a.Sq * VPR;
  if (idx >= total) return;
  const int v = (int)(idx % VPR);
  const long long t = idx / VPR;
  const int row = (int)(t % a.Sq);
  const int bh = (int)(t / a.Sq);
  const int b = bh / a.H, h = bh - b * a.H;
  const float4* src = reinterpret_cast<const float4*>(a.dq_accum + ((long long)bh * a.Sqp + row) * D + v * 8);
  const float4 x = src[0], y = src[1];
  uint4 out;
  out.x = pack2<BF16>(x.x * a.scale, x.y * a.scale);
  out.y = pack2<BF16>(x.z * a.scale, x.w * a.scale);
  out.z = pack2<BF16>(y.x * a.scale, y.y * a.scale);
  out.w = pack2<BF16>(y.z * a.scale, y.w * a.scale);
  uint16_t* dst = reinterpret_cast<uint16_t*>(dq.ptr) + b * dq.sb + h * dq.sh + (long long)row * dq.ss + v * 8;
  *reinterpret_cast<uint4*>(dst) = out;
}

}  // namespace

cudaError_t launch_bwd_prep(int head_dim, bool bf16, const TensorView& o, const TensorView& dout, const BwdArgs& a,
                            cudaStream_t stream) {
  const long long warps = (long long)a.B * a.H * a.Sqp;
  const int blocks = (int)((warps + 7) / 8);
  if (head_dim == 128 && bf16) fasn_bwd_prep_kernel<128, true><<<blocks, 256, 0, stream>>>(o, dout, a);
  else if (head_dim == 128) fasn_bwd_prep_kernel<128, false><<<blocks, 256, 0, stream>>>(o, dout, a);
  else if (head_dim == 64 && bf16) fasn_bwd_prep_kernel<64, true><<<blocks, 256, 0, stream>>>(o, dout, a);
  else if (head_dim == 64) fasn_bwd_prep_kernel<64, false><<<blocks, 256, 0, stream>>>(o, dout, a);
  else return cudaErrorInvalidValue;
  return cudaGetLastError();
}

cudaError_t launch_bwd_finish(int head_dim, bool bf16, const TensorView& dq, const BwdArgs& a, cudaStream_t stream) {
  const long long threads = (long long)a.B * a.H * a.Sq * (head_dim / 8);
  const int blocks = (int)((threads + 255) / 256);
  if (head_dim == 128 && bf16) fasn_bwd_finish_kernel<128, true><<<blocks, 256, 0, stream>>>(dq, a);
  else if (head_dim == 128) fasn_bwd_finish_kernel<128, false><<<blocks, 256, 0, stream>>>(dq, a);
  else if (head_dim == 64 && bf16) fasn_bwd_finish_kernel<64, true><<<blocks, 256, 0, stream>>>(dq, a);
  else if (head_dim == 64) fasn_bwd_finish_kernel<64, false><<<blocks, 256, 0, stream>>>(dq, a);
  else return cudaErrorInvalidValue;
  return cudaGetLastError();
}

}  // namespace fasn
