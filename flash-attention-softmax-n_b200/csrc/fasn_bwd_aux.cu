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
  // D / 8 lanes per (bh, row), 16-byte loads of O and dO (8 elements per lane), four rows per group of lanes in flight;
  // rows in [Sq, Sqp) are padding and get delta = 0.  HBM-bound: 2 x 2 D bytes read + 4 D bytes of zero-fill per row.
  constexpr int LPR = D / 8;                 // lanes per row: 16 (D=128) or 8 (D=64)
  constexpr int RPW = 32 / LPR;              // rows per warp pass
  constexpr int PASSES = 4;
  const int lane = threadIdx.x & 31;
  const int sub = lane / LPR, l = lane % LPR;
  const long long total = (long long)a.B * a.H * a.Sqp;
  const long long warp_global = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const long long row0 = warp_global * (RPW * PASSES) + sub;
  uint4 vo[PASSES], vd[PASSES];
  bool live[PASSES];
#pragma unroll
  for (int p = 0; p < PASSES; ++p) {
    const long long rg = row0 + p * RPW;
    live[p] = false;
    vo[p] = make_uint4(0, 0, 0, 0); vd[p] = make_uint4(0, 0, 0, 0);
    if (rg < total) {
      const int row = (int)(rg % a.Sqp);
      const int bh = (int)(rg / a.Sqp);
      const int b = bh / a.H, h = bh - b * a.H;
      if (row < a.Sq) {
        live[p] = true;
        vo[p] = __ldg(reinterpret_cast<const uint4*>(reinterpret_cast<const uint16_t*>(o.ptr) + b * o.sb + h * o.sh + (long long)row * o.ss) + l);
        vd[p] = __ldg(reinterpret_cast<const uint4*>(reinterpret_cast<const uint16_t*>(dout.ptr) + b * dout.sb + h * dout.sh + (long long)row * dout.ss) + l);
      }
    }
  }
#pragma unroll
  for (int p = 0; p < PASSES; ++p) {
    const long long rg = row0 + p * RPW;
    const uint32_t wo[4] = {vo[p].x, vo[p].y, vo[p].z, vo[p].w}, wd[4] = {vd[p].x, vd[p].y, vd[p].z, vd[p].w};
    float acc = 0.f;
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      acc = fmaf(cvt16_to_f32<BF16>(wo[e] & 0xFFFF), cvt16_to_f32<BF16>(wd[e] & 0xFFFF), acc);
      acc = fmaf(cvt16_to_f32<BF16>(wo[e] >> 16), cvt16_to_f32<BF16>(wd[e] >> 16), acc);
    }
#pragma unroll
    for (int sft = LPR / 2; sft > 0; sft >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, sft);
    if (rg < total) {
      const int row = (int)(rg % a.Sqp);
      const int bh = (int)(rg / a.Sqp);
      if (l == 0) {
        float* ws = const_cast<float*>(a.delta);
        ws[rg] = live[p] ? acc * a.keep_prob : 0.f;      // (1-p) delta: see the dS' formulation in fasn_bwd.cu
        // second half of the workspace: LSE_n in the log2 domain, +inf on padding rows (=> P = 0 there)
        ws[total + rg] = live[p] ? a.lse[(long long)bh * a.Sq + row] * kLog2e : INFINITY;
      }
      float4* acc_row = reinterpret_cast<float4*>(a.dq_accum + rg * D);   // D fp32 per row = 2 float4 per lane
      acc_row[l * 2] = make_float4(0.f, 0.f, 0.f, 0.f);
      acc_row[l * 2 + 1] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }
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
  const long long rows = (long long)a.B * a.H * a.Sqp;
  const int rows_per_warp = (32 / (head_dim / 8)) * 4;
  const long long warps = (rows + rows_per_warp - 1) / rows_per_warp;
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
