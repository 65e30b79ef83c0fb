// Fused tensor health check: ONE launch over up to 8 tensors, one status record per tensor.
//
// Replaces check_audio_tensor (ref:src/utils/debugging_utils.py:4-30), which the reference's step calls four times
// (ref:train_byol.py:52-59: clean / noisy waveforms, online prediction, target projection) and which costs >= 4 full
// passes and >= 4 host synchronisations per tensor (isnan().any(), isinf().any(), abs().sum() < t, abs().max().item(),
// plus mean / std / min / max .item() at DEBUG level).  Here every tensor is read ONCE (HBM-bound, 4 B per element) and
// the host reads one 64-byte record per tensor:
//   flags   bit 0 NaN present, bit 1 Inf present, bit 2 sum|x| < min_threshold, bit 3 max|x| > max_threshold
//           (the reference tests them in this order and reports the first that fires)
//   abs_max, max, min, abs_sum, sum, sum of squares, numel  (what the DEBUG statistics are derived from)
// Partial results meet in the record through atomics (the order of the fp64 additions is not fixed: the sums are
// diagnostics, compared against thresholds / printed with 4 decimals); the last CTA of a tensor finalises its record.
#include "common.cuh"

namespace nrse {
namespace {

constexpr int kCheckThreads = 256;
constexpr int kCheckWarps = kCheckThreads / 32;
constexpr int kCheckMaxTensors = NRSE_CHECK_MAX_TENSORS;

struct CheckArgs {
  const float* ptr[kCheckMaxTensors];
  long long numel[kCheckMaxTensors];
  nrse_tensor_check* out;  // [n]
  float max_threshold, min_threshold;
};

// order-preserving map float -> uint32 (for atomicMax / atomicMin on signed floats)
__device__ __forceinline__ unsigned ordered_key(float v) {
  const unsigned b = __float_as_uint(v);
  return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float ordered_value(unsigned k) {
  return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k);
}

struct Acc {
  unsigned flags = 0;
  float amax = 0.f, vmax = -INFINITY, vmin = INFINITY;
  float asum = 0.f, sum = 0.f, sq = 0.f;  // per-thread fp32 partials over <= a few thousand elements, then fp64
  __device__ __forceinline__ void add(float v) {
    const float a = fabsf(v);
    flags |= (v != v) ? 1u : 0u;
    flags |= (a == INFINITY) ? 2u : 0u;
    amax = fmaxf(amax, a);  // fmaxf drops NaNs: a NaN shows up in the flags, not in the peak
    vmax = fmaxf(vmax, v);
    vmin = fminf(vmin, v);
    asum += a;
    sum += v;
    sq = fmaf(v, v, sq);
  }
};

__global__ void __launch_bounds__(kCheckThreads) check_tensors_kernel(const CheckArgs a) {
  const int t = blockIdx.y;
  const float* __restrict__ p = a.ptr[t];
  const long long n = a.numel[t];
  // the record doubles as the working area until the last CTA finalises it: see nrse_tensor_check
  unsigned* rec_u = reinterpret_cast<unsigned*>(a.out + t);
  double* rec_d = reinterpret_cast<double*>(a.out + t);
  Acc acc;
  double asum = 0.0, sum = 0.0, sq = 0.0;
  const long long stride = static_cast<long long>(gridDim.x) * kCheckThreads;
  const long long tid = static_cast<long long>(blockIdx.x) * kCheckThreads + threadIdx.x;
  const bool vec = (reinterpret_cast<uintptr_t>(p) & 15u) == 0;
  const long long nvec = vec ? n / 4 : 0;
  int since_flush = 0;
  auto flush = [&] {
    asum += acc.asum; sum += acc.sum; sq += acc.sq;
    acc.asum = acc.sum = acc.sq = 0.f;
    since_flush = 0;
  };
  for (long long v = tid; v < nvec; v += stride) {
    const float4 q = ld_stream_f4(reinterpret_cast<const float4*>(p) + v);
    acc.add(q.x); acc.add(q.y); acc.add(q.z); acc.add(q.w);
    if (++since_flush == 64) flush();
  }
  for (long long i = nvec * 4 + tid; i < n; i += stride) {
    acc.add(__ldg(p + i));
    if (++since_flush == 64) flush();
  }
  flush();
  // CTA reduction
  __shared__ double s_d[kCheckWarps][3];
  __shared__ float s_f[kCheckWarps][3];
  __shared__ unsigned s_flags[kCheckWarps];
  __shared__ unsigned s_last;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  asum = warp_sum(asum); sum = warp_sum(sum); sq = warp_sum(sq);
  const float amax = warp_max(acc.amax), vmax = warp_max(acc.vmax), vmin = -warp_max(-acc.vmin);
  const unsigned flags = warp_or(acc.flags);
  if (lane == 0) {
    s_d[warp][0] = asum; s_d[warp][1] = sum; s_d[warp][2] = sq;
    s_f[warp][0] = amax; s_f[warp][1] = vmax; s_f[warp][2] = vmin;
    s_flags[warp] = flags;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double d0 = 0.0, d1 = 0.0, d2 = 0.0;
    float f0 = 0.f, f1 = -INFINITY, f2 = INFINITY;
    unsigned fl = 0;
    for (int w = 0; w < kCheckWarps; ++w) {
      d0 += s_d[w][0]; d1 += s_d[w][1]; d2 += s_d[w][2];
      f0 = fmaxf(f0, s_f[w][0]); f1 = fmaxf(f1, s_f[w][1]); f2 = fminf(f2, s_f[w][2]);
      fl |= s_flags[w];
    }
    // working layout (zeroed by the launcher): u[0] flags, u[1] abs-max bits, u[2] max key, u[3] ~min key,
    // d[2] abs_sum, d[3] sum, d[4] sumsq, u[12] ticket
    if (fl) atomicOr(rec_u + 0, fl);
    atomicMax(rec_u + 1, __float_as_uint(f0));
    atomicMax(rec_u + 2, ordered_key(f1));
    atomicMax(rec_u + 3, ~ordered_key(f2));
    atomicAdd(rec_d + 2, d0);
    atomicAdd(rec_d + 3, d1);
    atomicAdd(rec_d + 4, d2);
    __threadfence();
    s_last = atomicAdd(rec_u + 12, 1u) == gridDim.x - 1 ? 1u : 0u;
  }
  __syncthreads();
  if (s_last && threadIdx.x == 0) {  // every other CTA of this tensor has published: finalise the record
    __threadfence();
    volatile unsigned* vu = rec_u;
    volatile double* vd = rec_d;
    nrse_tensor_check r;
    unsigned fl = vu[0];
    r.abs_max = __uint_as_float(vu[1]);
    r.max = n > 0 ? ordered_value(vu[2]) : 0.f;
    r.min = n > 0 ? ordered_value(~vu[3]) : 0.f;
    r.abs_sum = vd[2];
    r.sum = vd[3];
    r.sumsq = vd[4];
    if (r.abs_sum < static_cast<double>(a.min_threshold)) fl |= 4u;  // NaN compares false, as in the reference
    if (r.abs_max > a.max_threshold) fl |= 8u;
    r.flags = static_cast<int32_t>(fl);
    r.numel = n;
    r.reserved = 0;
    r.reserved2 = 0;
    a.out[t] = r;
  }
}

}  // namespace
}  // namespace nrse

extern "C" {

int nrse_check_tensors_f32(const float* const* tensors_host, const int64_t* numel_host, int n_tensors,
                           float max_threshold, float min_threshold, nrse_tensor_check* out, nrse_stream_t stream) {
  using namespace nrse;
  static_assert(sizeof(nrse_tensor_check) == 64, "record layout");
  if (!tensors_host || !numel_host || !out || n_tensors < 1 || n_tensors > kCheckMaxTensors) return NRSE_ERR_INVALID_ARG;
  if (reinterpret_cast<uintptr_t>(out) & 7u) return NRSE_ERR_INVALID_ARG;
  CheckArgs a;
  long long longest = 0;
  for (int i = 0; i < kCheckMaxTensors; ++i) {
    a.ptr[i] = i < n_tensors ? tensors_host[i] : nullptr;
    a.numel[i] = i < n_tensors ? numel_host[i] : 0;
    if (i < n_tensors) {
      if (numel_host[i] < 0 || (numel_host[i] > 0 && !tensors_host[i])) return NRSE_ERR_INVALID_ARG;
      if (reinterpret_cast<uintptr_t>(tensors_host[i]) & 3u) return NRSE_ERR_INVALID_ARG;
      longest = numel_host[i] > longest ? numel_host[i] : longest;
    }
  }
  a.out = out;
  a.max_threshold = max_threshold;
  a.min_threshold = min_threshold;
  cudaStream_t s = as_stream(stream);
  NRSE_CUDA_TRY(cudaMemsetAsync(out, 0, sizeof(nrse_tensor_check) * n_tensors, s));
  // enough CTAs per tensor to fill the machine for the longest one, at >= 16 KB per CTA
  long long ctas = ceil_div(longest, static_cast<long long>(4096));
  const long long cap = ceil_div(static_cast<long long>(4 * kNumSMs), static_cast<long long>(n_tensors));
  ctas = ctas < 1 ? 1 : (ctas > cap ? cap : ctas);
  check_tensors_kernel<<<dim3(static_cast<unsigned>(ctas), n_tensors), kCheckThreads, 0, s>>>(a);
  NRSE_CHECK_LAUNCH();
  return NRSE_OK;
}

}  // extern "C"
