// C ABI of libfasn.so (declared in include/fasn.h): argument checking, TMA tensor-map construction, dispatch.
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <initializer_list>
#include <map>
#include <mutex>
#include <set>
#include <utility>
#include <string>
#include <vector>

#include "../../include/fasn.h"
#include "fasn_common.cuh"

namespace {

thread_local std::string g_last_error;

int fail(int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  g_last_error = buf;
  return code;
}
int fail_cuda(cudaError_t e, const char* what) {
  return fail(static_cast<int>(e), "%s: %s (%s)", what, cudaGetErrorName(e), cudaGetErrorString(e));
}

using EncodeTiledFn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                   const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                   CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_tiled() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

// (D, S, H, B) 16-bit tensor, box = 64 x 128 x 1 x 1, 128-byte swizzle, out-of-bounds rows read as zero.
int make_map(CUtensorMap* m, const void* ptr, long long sb, long long sh, long long ss, int B, int H, int S, int D,
             bool bf16, const char* name, int box_rows = 128) {
  EncodeTiledFn fn = encode_tiled();
  if (!fn) return fail(FASN_EDRIVER, "cuTensorMapEncodeTiled is not available from this driver");
  if (ptr == nullptr) return fail(FASN_EINVAL, "%s: null pointer", name);
  if ((reinterpret_cast<uintptr_t>(ptr) & 15) != 0) return fail(FASN_EUNSUPPORTED, "%s: pointer is not 16-byte aligned", name);
  if (ss < D || (ss % 8) != 0) return fail(FASN_EUNSUPPORTED, "%s: row stride %lld must be >= head_dim and a multiple of 8 elements", name, ss);
  // strides of size-1 axes are irrelevant to addressing: replace them by a packed value the encoder accepts
  if (H == 1 || sh <= 0) { if (H != 1) return fail(FASN_EUNSUPPORTED, "%s: head stride must be positive", name); sh = (long long)S * ss; }
  if (B == 1 || sb <= 0) { if (B != 1) return fail(FASN_EUNSUPPORTED, "%s: batch stride must be positive", name); sb = (long long)H * sh; }
  if ((sh % 8) != 0 || (sb % 8) != 0) return fail(FASN_EUNSUPPORTED, "%s: head/batch strides must be multiples of 8 elements", name);
  cuuint64_t dims[4] = {(cuuint64_t)D, (cuuint64_t)S, (cuuint64_t)H, (cuuint64_t)B};
  cuuint64_t strides[3] = {(cuuint64_t)ss * 2, (cuuint64_t)sh * 2, (cuuint64_t)sb * 2};
  cuuint32_t box[4] = {64, (cuuint32_t)box_rows, 1, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = fn(m, bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<void*>(ptr), dims,
                  strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(FASN_EDRIVER, "%s: cuTensorMapEncodeTiled failed with CUresult %d", name, (int)r);
  return 0;
}

// Launches (and cuTensorMapEncodeTiled, a driver call) need the context of the device that owns the tensors to be
// current on the calling thread.  Callers such as PyTorch's autograd worker threads may not have bound one yet
// (allocations served from a cache never touch the runtime), so bind it here and restore the previous device.
struct DeviceGuard {
  int prev = -1, dev = -1;
  cudaError_t err = cudaSuccess;
  explicit DeviceGuard(const void* device_ptr) {
    cudaPointerAttributes at{};
    err = cudaPointerGetAttributes(&at, device_ptr);
    if (err != cudaSuccess) return;
    if (at.type != cudaMemoryTypeDevice && at.type != cudaMemoryTypeManaged) { err = cudaErrorInvalidDevicePointer; return; }
    dev = at.device;
    if ((err = cudaGetDevice(&prev)) != cudaSuccess) return;
    err = cudaSetDevice(dev);
  }
  ~DeviceGuard() { if (prev >= 0 && prev != dev) cudaSetDevice(prev); }
};

// (D, Sqp, B*H, 1) fp32 accumulator, box = 32 x 128 x 1 x 1 (128-byte rows), 128-byte swizzle: target of the TMA reduce-add
int make_accum_map(CUtensorMap* m, float* ptr, long long BH, int Sqp, int D, int box_rows = 128) {
  EncodeTiledFn fn = encode_tiled();
  if (!fn) return fail(FASN_EDRIVER, "cuTensorMapEncodeTiled is not available from this driver");
  if ((reinterpret_cast<uintptr_t>(ptr) & 15) != 0) return fail(FASN_EUNSUPPORTED, "dq_accum: pointer is not 16-byte aligned");
  cuuint64_t dims[4] = {(cuuint64_t)D, (cuuint64_t)Sqp, (cuuint64_t)BH, 1};
  cuuint64_t strides[3] = {(cuuint64_t)D * 4, (cuuint64_t)Sqp * D * 4, (cuuint64_t)BH * Sqp * D * 4};
  cuuint32_t box[4] = {32, (cuuint32_t)box_rows, 1, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, ptr, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(FASN_EDRIVER, "dq_accum: cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
  return 0;
}

int check_common(const FasnParams* p) {
  if (p == nullptr) return fail(FASN_EINVAL, "params is null");
  if (p->struct_size != sizeof(FasnParams)) return fail(FASN_EINVAL, "struct_size %u != sizeof(FasnParams) %zu", p->struct_size, sizeof(FasnParams));
  if (p->dtype != FASN_FP16 && p->dtype != FASN_BF16) return fail(FASN_EUNSUPPORTED, "dtype %u: only fp16 (0) and bf16 (1)", p->dtype);
  if (p->head_dim != 64 && p->head_dim != 128) return fail(FASN_EUNSUPPORTED, "head_dim %d: only 64 and 128", p->head_dim);
  if (p->batch <= 0 || p->heads <= 0 || p->seqlen_q <= 0 || p->seqlen_kv <= 0) return fail(FASN_EINVAL, "batch/heads/seqlen must be positive");
  if (p->heads_kv != p->heads && p->heads_kv != 1) return fail(FASN_EINVAL, "heads_kv must equal heads or 1");
  {  // one-dimensional grids: (batch*heads) x tiles CTAs per launch
    const long long tiles = ((long long)(p->seqlen_q > p->seqlen_kv ? p->seqlen_q : p->seqlen_kv) + 127) / 128 + 1;
    if ((long long)p->batch * p->heads * tiles > 0x7FFFFFFFll) return fail(FASN_EUNSUPPORTED, "batch*heads*tiles exceeds 2^31-1 CTAs per launch: shard the batch x head axis");
  }
  if (!(p->softmax_n >= 0.f)) return fail(FASN_EINVAL, "softmax_n must be >= 0");
  if (!(p->dropout_p >= 0.f && p->dropout_p < 1.f)) return fail(FASN_EINVAL, "dropout_p must be in [0,1)");
  if (p->dropout_p > 0.f && lroundf((1.0f - p->dropout_p) * 256.0f) < 1)
    return fail(FASN_EINVAL, "dropout_p %.6f quantises to keep probability 0 (the generator resolves 1/256)", (double)p->dropout_p);
  if (!std::isfinite(p->scale)) return fail(FASN_EINVAL, "scale must be finite");
  if (p->lse == nullptr) return fail(FASN_EINVAL, "lse is null");
  if (p->alibi_slopes != nullptr && p->bias.ptr != nullptr) return fail(FASN_EINVAL, "alibi_slopes and bias are mutually exclusive");
  return 0;
}

// Optional per-kernel timing (fasn_profile): CUDA event pairs recorded around the two tensor-core kernels on the
// caller's stream, read back (and cleared) by fasn_profile_read after the caller has synchronised.
struct EventPair { cudaEvent_t a, b; };
struct Profile {
  std::mutex mu;
  bool enabled = false;
  std::vector<EventPair> fwd, bwd;
};
Profile g_prof;

struct ScopedEvents {
  std::vector<EventPair>* sink = nullptr;
  EventPair ev{};
  cudaStream_t st;
  ScopedEvents(std::vector<EventPair>& v, cudaStream_t s) : st(s) {
    std::lock_guard<std::mutex> l(g_prof.mu);
    if (!g_prof.enabled) return;
    if (cudaEventCreate(&ev.a) != cudaSuccess || cudaEventCreate(&ev.b) != cudaSuccess) return;
    sink = &v;
    cudaEventRecord(ev.a, st);
  }
  ~ScopedEvents() {
    if (!sink) return;
    cudaEventRecord(ev.b, st);
    std::lock_guard<std::mutex> l(g_prof.mu);
    sink->push_back(ev);
  }
};

// Work counters of the persistent kernels: a pool of (next item, CTAs finished) pairs per device, handed out round-robin,
// one pair per launch.  Every kernel leaves its pair at zero (the last CTA to finish resets it), so a pair is ready again
// long before its turn comes round: a collision would need kSchedSlots launches in flight at once.
constexpr int kSchedSlots = 4096;
struct DeviceState {
  int* sched = nullptr;
  int num_sms = 0;
  unsigned next = 0;
};
std::mutex g_dev_mu;
std::map<int, DeviceState> g_dev;

// (counter pair for one launch, SM count) of the device that is current (the DeviceGuard has bound the tensors' device)
int sched_slot(int** slot, int* num_sms) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return fail_cuda(e, "cudaGetDevice");
  std::lock_guard<std::mutex> l(g_dev_mu);
  DeviceState& d = g_dev[dev];
  if (d.sched == nullptr) {
    if ((e = cudaDeviceGetAttribute(&d.num_sms, cudaDevAttrMultiProcessorCount, dev)) != cudaSuccess) return fail_cuda(e, "cudaDeviceGetAttribute");
    if ((e = cudaMalloc(&d.sched, 2 * kSchedSlots * sizeof(int))) != cudaSuccess) return fail_cuda(e, "scheduler counters cudaMalloc");
    if ((e = cudaMemset(d.sched, 0, 2 * kSchedSlots * sizeof(int))) != cudaSuccess) return fail_cuda(e, "scheduler counters cudaMemset");
  }
  *slot = d.sched + 2 * (d.next++ % kSchedSlots);
  *num_sms = d.num_sms;
  return 0;
}

fasn::AuxView aux_view(const FasnAux& a) { return fasn::AuxView{a.ptr, a.stride_b, a.stride_h, a.stride_q}; }
fasn::TensorView tensor_view(const FasnTensor& t) { return fasn::TensorView{t.ptr, t.stride_b, t.stride_h, t.stride_s}; }

// Dropout keeps an element iff an 8-bit uniform u < T, T = round(256 (1 - p)): the keep probability actually realised is
// T / 256, and kept elements are scaled by its inverse 256 / T (not by 1 / (1 - p)), so that E[dropout(P)] = P exactly.
// p < 1/512 quantises to T = 256 (nothing dropped, scale 1); p > 1 - 1/512 would quantise to T = 0 and is rejected.
uint32_t keep_threshold(float dropout_p) {
  long t = lroundf((1.0f - dropout_p) * 256.0f);
  if (t < 0) t = 0;
  if (t > 256) t = 256;
  return (uint32_t)t;
}
float keep_probability(float dropout_p) { return dropout_p > 0.f ? (float)keep_threshold(dropout_p) / 256.0f : 1.0f; }


}  // namespace

#ifdef FASN_TIMELINE
unsigned long long* g_fasn_timeline = nullptr;
unsigned int g_fasn_timeline_xy[2] = {0, 0};
unsigned long long* g_fasn_timeline_fwd = nullptr;
unsigned int g_fasn_timeline_fwd_xy[2] = {0, 0};
extern "C" void fasn_set_timeline_fwd(unsigned long long* buf, unsigned int x, unsigned int y) {
  g_fasn_timeline_fwd = buf; g_fasn_timeline_fwd_xy[0] = x; g_fasn_timeline_fwd_xy[1] = y;
}
extern "C" void fasn_set_timeline(unsigned long long* buf, unsigned int x, unsigned int y) {
  g_fasn_timeline = buf; g_fasn_timeline_xy[0] = x; g_fasn_timeline_xy[1] = y;
}
#endif

extern "C" {

int fasn_version(void) { return FASN_ABI_VERSION; }

const char* fasn_last_error(void) { return g_last_error.c_str(); }

int fasn_fwd(const FasnParams* p) {
  if (int rc = check_common(p)) return rc;
  if (p->q.ptr == nullptr) return fail(FASN_EINVAL, "q: null pointer");
  DeviceGuard guard(p->q.ptr);
  if (guard.err != cudaSuccess) return fail_cuda(guard.err, "q is not a device pointer / cannot bind its device");
  const bool bf16 = p->dtype == FASN_BF16;
  const int B = p->batch, H = p->heads, Hkv = p->heads_kv, L = p->seqlen_q, S = p->seqlen_kv, D = p->head_dim;
  CUtensorMap tq, tk, tv, to;
  if (int rc = make_map(&tq, p->q.ptr, p->q.stride_b, p->q.stride_h, p->q.stride_s, B, H, L, D, bf16, "q")) return rc;
  if (int rc = make_map(&tk, p->k.ptr, p->k.stride_b, p->k.stride_h, p->k.stride_s, B, Hkv, S, D, bf16, "k")) return rc;
  if (int rc = make_map(&tv, p->v.ptr, p->v.stride_b, p->v.stride_h, p->v.stride_s, B, Hkv, S, D, bf16, "v")) return rc;
  if (int rc = make_map(&to, p->o.ptr, p->o.stride_b, p->o.stride_h, p->o.stride_s, B, H, L, D, bf16, "o")) return rc;
  fasn::FwdArgs a{};
  a.B = B; a.H = H; a.Hkv = Hkv; a.Sq = L; a.Skv = S;
  a.causal_off = S - L;
  a.scale_log2 = p->scale * fasn::kLog2e;
  a.softmax_n = p->softmax_n;
  a.lse = p->lse;
  a.o = tensor_view(p->o);
  a.mask = aux_view(p->mask);
  a.bias = aux_view(p->bias);
  a.alibi = p->alibi_slopes;
#if defined(FASN_DEBUG_FP32_P) && FASN_DEBUG_FP32_P
  a.o_f32 = p->o_f32;
#else
  if (p->o_f32 != nullptr) return fail(FASN_EUNSUPPORTED, "o_f32 is an output of the debug library (libfasn_debug32.so) only");
  a.o_f32 = nullptr;
#endif
  a.drop_thr = keep_threshold(p->dropout_p);
  a.inv_keep = 1.0f / keep_probability(p->dropout_p);
  a.key = fasn::make_philox_key(p->philox_seed, p->philox_offset, a.drop_thr);
  a.bh_offset = (uint32_t)p->bh_offset;
  a.sched_group = fasn::sched_group_size(2ll * D * (2ll * L + 2ll * S));            // Q, O, K, V of one unit
  {
    int num_sms = 0;
    if (int rc = sched_slot(&a.sched, &num_sms)) return rc;
    const long long items = (long long)B * H * ((L + 255) / 256);
    if (items > 0x7FFFFFFFll) return fail(FASN_EUNSUPPORTED, "more than 2^31-1 forward work items (batch x heads x 256-row blocks) per call");
    a.grid_ctas = (int)(items < num_sms ? items : num_sms);
  }
#ifdef FASN_TIMELINE
  {
    extern unsigned long long* g_fasn_timeline_fwd; extern unsigned int g_fasn_timeline_fwd_xy[2];
    a.dbg = g_fasn_timeline_fwd; a.dbg_x = g_fasn_timeline_fwd_xy[0]; a.dbg_y = g_fasn_timeline_fwd_xy[1];
  }
#endif
  cudaError_t e;
  {
    ScopedEvents prof(g_prof.fwd, (cudaStream_t)p->stream);
    e = fasn::launch_fwd(D, bf16, p->is_causal != 0, p->dropout_p > 0.f, tq, tk, tv, to, a, (cudaStream_t)p->stream);
  }
  if (e != cudaSuccess) return fail_cuda(e, "fasn_fwd launch");
  return 0;
}

int fasn_bwd_workspace(const FasnParams* p, uint64_t* delta_bytes, uint64_t* dq_accum_bytes) {
  if (p == nullptr || delta_bytes == nullptr || dq_accum_bytes == nullptr) return fail(FASN_EINVAL, "null argument");
  const uint64_t Lp = ((uint64_t)p->seqlen_q + 127) / 128 * 128;
  *delta_bytes = 2ull * (uint64_t)p->batch * p->heads * Lp * sizeof(float);   // delta + LSE_n*log2e
  *dq_accum_bytes = (uint64_t)p->batch * p->heads * Lp * p->head_dim * sizeof(float);
  return 0;
}

int fasn_bwd(const FasnParams* p) {
  if (int rc = check_common(p)) return rc;
  if (p->delta == nullptr || p->dq_accum == nullptr) return fail(FASN_EINVAL, "delta / dq_accum workspace is null");
  if (p->dq.ptr == nullptr || p->q.ptr == nullptr) return fail(FASN_EINVAL, "q / dq is null");
  DeviceGuard guard(p->q.ptr);
  if (guard.err != cudaSuccess) return fail_cuda(guard.err, "q is not a device pointer / cannot bind its device");
  const bool bf16 = p->dtype == FASN_BF16;
  const int B = p->batch, H = p->heads, Hkv = p->heads_kv, L = p->seqlen_q, S = p->seqlen_kv, D = p->head_dim;
  CUtensorMap tq, tk, tv, tdo, tdk, tdv;
  if (int rc = make_map(&tq, p->q.ptr, p->q.stride_b, p->q.stride_h, p->q.stride_s, B, H, L, D, bf16, "q")) return rc;
  if (int rc = make_map(&tk, p->k.ptr, p->k.stride_b, p->k.stride_h, p->k.stride_s, B, Hkv, S, D, bf16, "k")) return rc;
  if (int rc = make_map(&tv, p->v.ptr, p->v.stride_b, p->v.stride_h, p->v.stride_s, B, Hkv, S, D, bf16, "v")) return rc;
  if (int rc = make_map(&tdo, p->dout.ptr, p->dout.stride_b, p->dout.stride_h, p->dout.stride_s, B, H, L, D, bf16, "dout")) return rc;
  const bool head_sum = p->dk_accum != nullptr || p->dv_accum != nullptr;
  if (head_sum) {
    if (p->dk_accum == nullptr || p->dv_accum == nullptr) return fail(FASN_EINVAL, "dk_accum and dv_accum come as a pair");
    if (Hkv != 1) return fail(FASN_EINVAL, "dk_accum / dv_accum are for shared K/V (heads_kv == 1)");
    if ((reinterpret_cast<uintptr_t>(p->dk_accum) | reinterpret_cast<uintptr_t>(p->dv_accum)) & 15) return fail(FASN_EINVAL, "dk_accum / dv_accum must be 16-byte aligned");
    tdk = tq; tdv = tq;      // never used: the kernel adds into the accumulators instead of storing dk / dv
  } else {
    if (int rc = make_map(&tdk, p->dk.ptr, p->dk.stride_b, p->dk.stride_h, p->dk.stride_s, B, H, S, D, bf16, "dk")) return rc;
    if (int rc = make_map(&tdv, p->dv.ptr, p->dv.stride_b, p->dv.stride_h, p->dv.stride_s, B, H, S, D, bf16, "dv")) return rc;
  }
  if (p->o.ptr == nullptr) return fail(FASN_EINVAL, "o is null");
  // the delta pre-pass reads O and dO with 16-byte loads
  if ((reinterpret_cast<uintptr_t>(p->o.ptr) & 15) != 0 || (p->o.stride_s % 8) != 0 || (p->o.stride_h % 8) != 0 || (p->o.stride_b % 8) != 0)
    return fail(FASN_EUNSUPPORTED, "o: pointer must be 16-byte aligned and strides multiples of 8 elements");
  CUtensorMap tdq;
  if (int rc = make_accum_map(&tdq, p->dq_accum, (long long)B * H, (L + 127) / 128 * 128, D)) return rc;
  fasn::BwdArgs a{};
  a.B = B; a.H = H; a.Hkv = Hkv; a.Sq = L; a.Skv = S;
  a.causal_off = S - L;
  a.scale = p->scale / keep_probability(p->dropout_p);
  a.scale_log2 = p->scale * fasn::kLog2e;
  a.keep_prob = keep_probability(p->dropout_p);
  a.lse = p->lse;
  a.delta = p->delta;
  a.dq_accum = p->dq_accum;
  a.Sqp = (L + 127) / 128 * 128;
  a.mask = aux_view(p->mask);
  a.bias = aux_view(p->bias);
  a.alibi = p->alibi_slopes;
  a.dbias = fasn::AuxView{p->dbias, p->dbias_stride_b, p->dbias_stride_h, p->dbias_stride_q};
  a.dk_accum = p->dk_accum; a.dv_accum = p->dv_accum;
  if (p->dbias != nullptr && p->bias.ptr == nullptr && p->alibi_slopes == nullptr && !(p->mask.ptr != nullptr && p->mask.stride_q != 0))
    return fail(FASN_EINVAL, "dbias is produced by the dense-tensor kernels: pass the bias (or ALiBi slopes / a dense mask) it belongs to");
  a.drop_thr = keep_threshold(p->dropout_p);
  a.inv_keep = 1.0f / keep_probability(p->dropout_p);
  a.key = fasn::make_philox_key(p->philox_seed, p->philox_offset, a.drop_thr);
  a.bh_offset = (uint32_t)p->bh_offset;
  a.sched_group = fasn::sched_group_size(2ll * D * (2ll * L + 2ll * S) + 4ll * D * L);   // Q, dO, K, V + the fp32 dQ accumulator
  {
    int num_sms = 0;
    if (int rc = sched_slot(&a.sched, &num_sms)) return rc;
    const long long items = (long long)B * H * ((S + 127) / 128);
    if (items > 0x7FFFFFFFll) return fail(FASN_EUNSUPPORTED, "more than 2^31-1 backward work items (batch x heads x 128-row K/V tiles) per call");
    a.grid_ctas = (int)(items < num_sms ? items : num_sms);
  }
#ifdef FASN_TIMELINE
  {
    extern unsigned long long* g_fasn_timeline; extern unsigned int g_fasn_timeline_xy[2];
    a.dbg = g_fasn_timeline; a.dbg_x = g_fasn_timeline_xy[0]; a.dbg_y = g_fasn_timeline_xy[1];
  }
#endif
  cudaStream_t st = (cudaStream_t)p->stream;
  cudaError_t e = fasn::launch_bwd_prep(D, bf16, tensor_view(p->o), tensor_view(p->dout), a, st);
  if (e != cudaSuccess) return fail_cuda(e, "fasn_bwd prep launch");
  {
    ScopedEvents prof(g_prof.bwd, st);
    e = fasn::launch_bwd(D, bf16, p->is_causal != 0, p->dropout_p > 0.f, tq, tk, tv, tdo, tdk, tdv, tdq, a, tensor_view(p->dk), tensor_view(p->dv), st);
  }
  if (e != cudaSuccess) return fail_cuda(e, "fasn_bwd main launch");
  e = fasn::launch_bwd_finish(D, bf16, tensor_view(p->dq), a, st);
  if (e != cudaSuccess) return fail_cuda(e, "fasn_bwd finish launch");
  return 0;
}

int fasn_profile(int enable) {
  std::lock_guard<std::mutex> l(g_prof.mu);
  g_prof.enabled = enable != 0;
  return 0;
}

int fasn_profile_read(double* fwd_ms, int32_t* fwd_launches, double* bwd_ms, int32_t* bwd_launches) {
  if (!fwd_ms || !fwd_launches || !bwd_ms || !bwd_launches) return fail(FASN_EINVAL, "null argument");
  std::lock_guard<std::mutex> l(g_prof.mu);
  auto drain = [](std::vector<EventPair>& v, double* ms, int32_t* n) -> cudaError_t {
    *ms = 0.0; *n = 0;
    cudaError_t err = cudaSuccess;
    for (auto& e : v) {
      float t = 0.f;
      cudaError_t r = cudaEventElapsedTime(&t, e.a, e.b);
      if (r == cudaSuccess) { *ms += t; *n += 1; } else err = r;
      cudaEventDestroy(e.a); cudaEventDestroy(e.b);
    }
    v.clear();
    return err;
  };
  cudaError_t e1 = drain(g_prof.fwd, fwd_ms, fwd_launches);
  cudaError_t e2 = drain(g_prof.bwd, bwd_ms, bwd_launches);
  if (e1 != cudaSuccess) return fail_cuda(e1, "fasn_profile_read (forward events; synchronise the stream first)");
  if (e2 != cudaSuccess) return fail_cuda(e2, "fasn_profile_read (backward events; synchronise the stream first)");
  return 0;
}

int fasn_dropout_mask(uint8_t* out, int32_t batch, int32_t heads, int32_t seqlen_q, int32_t seqlen_kv, float dropout_p,
                      uint64_t philox_seed, uint64_t philox_offset, int64_t bh_offset, void* stream) {
  if (out == nullptr || batch <= 0 || heads <= 0 || seqlen_q <= 0 || seqlen_kv <= 0) return fail(FASN_EINVAL, "bad argument");
  if (!(dropout_p >= 0.f && dropout_p < 1.f)) return fail(FASN_EINVAL, "dropout_p must be in [0,1)");
  cudaError_t e = fasn::launch_dropout_mask(out, batch, heads, seqlen_q, seqlen_kv, keep_threshold(dropout_p),
                                            fasn::make_philox_key(philox_seed, philox_offset, keep_threshold(dropout_p)), (uint32_t)bh_offset, (cudaStream_t)stream);
  if (e != cudaSuccess) return fail_cuda(e, "fasn_dropout_mask launch");
  return 0;
}

namespace {
bool sm_pair_ok(uint32_t in, uint32_t out) { return in <= 2 && out <= 2 && (in == out || in == 2 || out == 2); }
int sm_vec(int cols, std::initializer_list<const void*> ptrs, std::initializer_list<long long> strides) {
  if (cols % 8) return 0;
  for (const void* q : ptrs) if (reinterpret_cast<uintptr_t>(q) & 15) return 0;
  for (long long st : strides) if (st % 8) return 0;
  return 1;
}
}  // namespace

int fasn_softmax_n_fwd(const void* x, void* y, int64_t rows, int32_t cols, int64_t x_row_stride, int64_t y_row_stride,
                       uint32_t dtype_in, uint32_t dtype_out, float n, void* stream) {
  if (x == nullptr || y == nullptr || rows < 0 || cols <= 0) return fail(FASN_EINVAL, "fasn_softmax_n_fwd: bad argument");
  if (!sm_pair_ok(dtype_in, dtype_out)) return fail(FASN_EUNSUPPORTED, "fasn_softmax_n_fwd: unsupported dtype pair (%u -> %u)", dtype_in, dtype_out);
  if (!(n >= 0.f)) return fail(FASN_EINVAL, "softmax_n must be >= 0");
  if (rows > 0x7FFFFFFFll) return fail(FASN_EUNSUPPORTED, "fasn_softmax_n_fwd: more than 2^31-1 rows per call");
  if (rows == 0) return 0;
  DeviceGuard guard(x);
  if (guard.err != cudaSuccess) return fail_cuda(guard.err, "x is not a device pointer / cannot bind its device");
  cudaError_t e = fasn::launch_softmax_n_fwd(x, y, rows, cols, x_row_stride, y_row_stride, (int)dtype_in, (int)dtype_out, n,
                                             sm_vec(cols, {x, y}, {x_row_stride, y_row_stride}), (cudaStream_t)stream);
  if (e != cudaSuccess) return fail_cuda(e, "fasn_softmax_n_fwd launch");
  return 0;
}

int fasn_softmax_n_bwd(const void* y, const void* dy, void* dx, int64_t rows, int32_t cols, int64_t y_row_stride,
                       int64_t dy_row_stride, int64_t dx_row_stride, uint32_t dtype_in, uint32_t dtype_out, void* stream) {
  if (y == nullptr || dy == nullptr || dx == nullptr || rows < 0 || cols <= 0) return fail(FASN_EINVAL, "fasn_softmax_n_bwd: bad argument");
  if (!sm_pair_ok(dtype_in, dtype_out)) return fail(FASN_EUNSUPPORTED, "fasn_softmax_n_bwd: unsupported dtype pair (%u -> %u)", dtype_in, dtype_out);
  if (rows > 0x7FFFFFFFll) return fail(FASN_EUNSUPPORTED, "fasn_softmax_n_bwd: more than 2^31-1 rows per call");
  if (rows == 0) return 0;
  DeviceGuard guard(y);
  if (guard.err != cudaSuccess) return fail_cuda(guard.err, "y is not a device pointer / cannot bind its device");
  cudaError_t e = fasn::launch_softmax_n_bwd(y, dy, dx, rows, cols, y_row_stride, dy_row_stride, dx_row_stride, (int)dtype_in, (int)dtype_out,
                                             sm_vec(cols, {y, dy, dx}, {y_row_stride, dy_row_stride, dx_row_stride}), (cudaStream_t)stream);
  if (e != cudaSuccess) return fail_cuda(e, "fasn_softmax_n_bwd launch");
  return 0;
}

int fasn_probe(int mode, uint32_t dtype, const void* x, const void* y, float* c, void* stream) {
  if (mode < 0 || mode > 4 || x == nullptr || y == nullptr || c == nullptr) return fail(FASN_EINVAL, "bad argument");
  if (dtype != FASN_FP16 && dtype != FASN_BF16) return fail(FASN_EUNSUPPORTED, "dtype");
  const bool bf16 = dtype == FASN_BF16;
  DeviceGuard guard(x);
  if (guard.err != cudaSuccess) return fail_cuda(guard.err, "x is not a device pointer / cannot bind its device");
  CUtensorMap tx, ty;
  if (int rc = make_map(&tx, x, 0, 0, 128, 1, 1, 128, 128, bf16, "x")) return rc;
  if (int rc = make_map(&ty, y, 0, 0, 128, 1, 1, 128, 128, bf16, "y")) return rc;
  cudaError_t e = fasn::launch_probe(mode, bf16, tx, ty, x, c, (cudaStream_t)stream);
  if (e != cudaSuccess) return fail_cuda(e, "fasn_probe launch");
  return 0;
}

namespace {
std::mutex g_peer_mu;
std::set<std::pair<int, int>> g_peer_enabled;     // (accessing device, accessed device) pairs already enabled in this process

// Device that owns `p`, or -1 for host memory / unknown pointers.
int device_of(const void* p) {
  cudaPointerAttributes at{};
  if (cudaPointerGetAttributes(&at, p) != cudaSuccess) { cudaGetLastError(); return -1; }
  return (at.type == cudaMemoryTypeDevice || at.type == cudaMemoryTypeManaged) ? at.device : -1;
}

// Without peer access enabled between the two devices IN THIS PROCESS the runtime stages a device-to-device copy through
// host memory (measured 30 GB/s instead of 790 GB/s over NVLink), also for memory mapped with cudaIpcOpenMemHandle.
cudaError_t ensure_peer_access(int from, int to) {
  if (from < 0 || to < 0 || from == to) return cudaSuccess;
  std::lock_guard<std::mutex> l(g_peer_mu);
  if (g_peer_enabled.count({from, to})) return cudaSuccess;
  int can = 0;
  cudaError_t e = cudaDeviceCanAccessPeer(&can, from, to);
  if (e != cudaSuccess) return e;
  if (can) {
    int prev = -1;
    if ((e = cudaGetDevice(&prev)) != cudaSuccess) return e;
    if ((e = cudaSetDevice(from)) != cudaSuccess) return e;
    e = cudaDeviceEnablePeerAccess(to, 0);
    if (e == cudaErrorPeerAccessAlreadyEnabled) { cudaGetLastError(); e = cudaSuccess; }
    cudaSetDevice(prev);
    if (e != cudaSuccess) return e;
  }
  g_peer_enabled.insert({from, to});
  return cudaSuccess;
}
}  // namespace

int fasn_copy_async(void* dst, const void* src, uint64_t bytes, void* stream) {
  if (dst == nullptr || src == nullptr) return fail(FASN_EINVAL, "fasn_copy_async: null pointer");
  if (bytes == 0) return 0;
  {
    const int ds = device_of(src), dd = device_of(dst);
    cudaError_t pe = ensure_peer_access(ds, dd);
    if (pe == cudaSuccess) pe = ensure_peer_access(dd, ds);
    if (pe != cudaSuccess) return fail_cuda(pe, "fasn_copy_async: enabling peer access");
  }
  cudaError_t e = cudaMemcpyAsync(dst, src, (size_t)bytes, cudaMemcpyDefault, (cudaStream_t)stream);
  if (e != cudaSuccess) return fail_cuda(e, "fasn_copy_async");
  return 0;
}

// -------------------------------------------------------------------------------------------------
// host-buffer entry point (end-to-end measurement: H2D + kernels + D2H inside one call)
// -------------------------------------------------------------------------------------------------
namespace {
// One device arena per device (keyed by the device that is current when fasn_attention_host is called), grown on demand.
struct Arena {
  void* base = nullptr;
  size_t cap = 0;
};
std::mutex g_arena_mu;
std::map<int, Arena> g_arenas;
}  // namespace

int fasn_attention_host(uint32_t dtype, int32_t batch, int32_t heads, int32_t seqlen_q, int32_t seqlen_kv, int32_t head_dim,
                        const void* q_host, const void* k_host, const void* v_host, void* o_host, const void* dout_host,
                        void* dq_host, void* dk_host, void* dv_host, float softmax_n, float scale, int32_t is_causal,
                        float dropout_p, uint64_t philox_seed, uint64_t philox_offset, void* stream) {
  if (!q_host || !k_host || !v_host || !o_host) return fail(FASN_EINVAL, "null host pointer");
  const bool bwd = dout_host != nullptr;
  if (bwd && (!dq_host || !dk_host || !dv_host)) return fail(FASN_EINVAL, "backward requested but a gradient host pointer is null");
  if (batch <= 0 || heads <= 0 || seqlen_q <= 0 || seqlen_kv <= 0 || (head_dim != 64 && head_dim != 128))
    return fail(FASN_EINVAL, "bad shape");
  cudaStream_t st = (cudaStream_t)stream;
  const size_t BH = (size_t)batch * heads;
  const size_t Lp = ((size_t)seqlen_q + 127) / 128 * 128;
  const size_t nq = BH * seqlen_q * head_dim * 2, nkv = BH * seqlen_kv * head_dim * 2;
  auto up = [](size_t x) { return (x + 255) / 256 * 256; };
  // layout: q k v o lse | dout dq dk dv delta dq_accum
  size_t off = 0;
  auto take = [&](size_t bytes) { size_t o = off; off += up(bytes); return o; };
  const size_t o_q = take(nq), o_k = take(nkv), o_v = take(nkv), o_o = take(nq), o_lse = take(BH * seqlen_q * 4);
  size_t o_do = 0, o_dq = 0, o_dk = 0, o_dv = 0, o_delta = 0, o_acc = 0;
  if (bwd) {
    o_do = take(nq); o_dq = take(nq); o_dk = take(nkv); o_dv = take(nkv); o_delta = take(2 * BH * Lp * 4);
    o_acc = take(BH * Lp * head_dim * 4);
  }
  int dev = 0;
  if (cudaError_t e0 = cudaGetDevice(&dev); e0 != cudaSuccess) return fail_cuda(e0, "cudaGetDevice");
  std::lock_guard<std::mutex> lock(g_arena_mu);
  Arena& g_arena = g_arenas[dev];
  if (g_arena.cap < off) {
    if (g_arena.base) cudaFree(g_arena.base);
    g_arena.base = nullptr; g_arena.cap = 0;
    cudaError_t e = cudaMalloc(&g_arena.base, off);
    if (e != cudaSuccess) return fail_cuda(e, "arena cudaMalloc");
    g_arena.cap = off;
  }
  char* base = static_cast<char*>(g_arena.base);
  cudaError_t e;
  if ((e = cudaMemcpyAsync(base + o_q, q_host, nq, cudaMemcpyHostToDevice, st)) != cudaSuccess) return fail_cuda(e, "H2D q");
  if ((e = cudaMemcpyAsync(base + o_k, k_host, nkv, cudaMemcpyHostToDevice, st)) != cudaSuccess) return fail_cuda(e, "H2D k");
  if ((e = cudaMemcpyAsync(base + o_v, v_host, nkv, cudaMemcpyHostToDevice, st)) != cudaSuccess) return fail_cuda(e, "H2D v");
  if (bwd && (e = cudaMemcpyAsync(base + o_do, dout_host, nq, cudaMemcpyHostToDevice, st)) != cudaSuccess) return fail_cuda(e, "H2D dout");

  FasnParams p;
  memset(&p, 0, sizeof(p));
  p.struct_size = sizeof(FasnParams);
  p.dtype = dtype;
  p.batch = batch; p.heads = heads; p.heads_kv = heads; p.seqlen_q = seqlen_q; p.seqlen_kv = seqlen_kv; p.head_dim = head_dim;
  auto tv4 = [&](size_t o, int S) {
    FasnTensor t; t.ptr = base + o; t.stride_s = head_dim; t.stride_h = (int64_t)S * head_dim; t.stride_b = (int64_t)heads * S * head_dim;
    return t;
  };
  p.q = tv4(o_q, seqlen_q); p.k = tv4(o_k, seqlen_kv); p.v = tv4(o_v, seqlen_kv); p.o = tv4(o_o, seqlen_q);
  p.lse = reinterpret_cast<float*>(base + o_lse);
  p.softmax_n = softmax_n; p.scale = scale; p.is_causal = is_causal; p.dropout_p = dropout_p;
  p.philox_seed = philox_seed; p.philox_offset = philox_offset; p.bh_offset = 0;
  p.stream = stream;
  if (int rc = fasn_fwd(&p)) return rc;
  if ((e = cudaMemcpyAsync(o_host, base + o_o, nq, cudaMemcpyDeviceToHost, st)) != cudaSuccess) return fail_cuda(e, "D2H o");
  if (bwd) {
    p.dout = tv4(o_do, seqlen_q); p.dq = tv4(o_dq, seqlen_q); p.dk = tv4(o_dk, seqlen_kv); p.dv = tv4(o_dv, seqlen_kv);
    p.delta = reinterpret_cast<float*>(base + o_delta);
    p.dq_accum = reinterpret_cast<float*>(base + o_acc);
    if (int rc = fasn_bwd(&p)) return rc;
    if ((e = cudaMemcpyAsync(dq_host, base + o_dq, nq, cudaMemcpyDeviceToHost, st)) != cudaSuccess) return fail_cuda(e, "D2H dq");
    if ((e = cudaMemcpyAsync(dk_host, base + o_dk, nkv, cudaMemcpyDeviceToHost, st)) != cudaSuccess) return fail_cuda(e, "D2H dk");
    if ((e = cudaMemcpyAsync(dv_host, base + o_dv, nkv, cudaMemcpyDeviceToHost, st)) != cudaSuccess) return fail_cuda(e, "D2H dv");
  }
  if ((e = cudaStreamSynchronize(st)) != cudaSuccess) return fail_cuda(e, "stream synchronize");
  return 0;
}

}  // extern "C"
