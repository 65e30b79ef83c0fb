// Multi-tensor EMA of the BYOL target network (fp32, in place, HBM-bound: 12 B per parameter).
//
// Replaces BYOLSpeechModel._update_target_network (ref:src/models/byol.py:62-73), which issues three
// elementwise kernels and three allocations per parameter tensor (1488 launches for WavLM-large).  Here
// the 496 tensors are cut once into a table of equally sized chunks; one launch walks the table with a
// grid sized to the machine, 128-bit loads/stores, and the same rounding as the reference expression
// `decay * t + (1 - decay) * o` (two products and one sum, each rounded to fp32 -- no FMA contraction),
// so the result is bit-identical to the reference.
#include "common.cuh"

namespace nrse {
namespace {

constexpr int kEmaThreads = 256;
constexpr int kEmaUnroll = 4;  // float4 per thread per array in flight

__device__ __forceinline__ float ema1(float t, float o, float decay, float omd) {
  return __fadd_rn(__fmul_rn(decay, t), __fmul_rn(omd, o));
}

__global__ void __launch_bounds__(kEmaThreads) ema_chunks_kernel(const uint64_t* __restrict__ chunk_target,
                                                                 const uint64_t* __restrict__ chunk_online,
                                                                 const int32_t* __restrict__ chunk_numel,
                                                                 long long n_chunks, float decay, float omd) {
  for (long long ch = blockIdx.x; ch < n_chunks; ch += gridDim.x) {
    float* __restrict__ t = reinterpret_cast<float*>(chunk_target[ch]);
    const float* __restrict__ o = reinterpret_cast<const float*>(chunk_online[ch]);
    const int n = chunk_numel[ch];
    const bool vec_ok = ((reinterpret_cast<uintptr_t>(t) | reinterpret_cast<uintptr_t>(o)) & 15u) == 0;
    const int nvec = vec_ok ? (n >> 2) : 0;
    float4* t4 = reinterpret_cast<float4*>(t);
    const float4* o4 = reinterpret_cast<const float4*>(o);
    for (int v = threadIdx.x; v < nvec; v += kEmaThreads * kEmaUnroll) {
      float4 tv[kEmaUnroll], ov[kEmaUnroll];
#pragma unroll
      for (int u = 0; u < kEmaUnroll; ++u) {
        const int i = v + u * kEmaThreads;
        if (i < nvec) {
          tv[u] = t4[i];               // target is rewritten below: plain (coherent) load
          ov[u] = ld_stream_f4(o4 + i);  // online weights are read-only here
        }
      }
#pragma unroll
      for (int u = 0; u < kEmaUnroll; ++u) {
        const int i = v + u * kEmaThreads;
        if (i < nvec) {
          float4 r;
          r.x = ema1(tv[u].x, ov[u].x, decay, omd);
          r.y = ema1(tv[u].y, ov[u].y, decay, omd);
          r.z = ema1(tv[u].z, ov[u].z, decay, omd);
          r.w = ema1(tv[u].w, ov[u].w, decay, omd);
          t4[i] = r;
        }
      }
    }
    for (int i = (nvec << 2) + threadIdx.x; i < n; i += kEmaThreads) t[i] = ema1(t[i], o[i], decay, omd);
  }
}

}  // namespace
}  // namespace nrse

extern "C" {

int64_t nrse_ema_plan_chunks_host(const uint64_t* target_ptrs_host, const uint64_t* online_ptrs_host,
                                  const int64_t* numel_host, int n_tensors, int64_t chunk_elems,
                                  uint64_t* chunk_target_host, uint64_t* chunk_online_host,
                                  int32_t* chunk_numel_host, int64_t max_chunks) {
  if (!target_ptrs_host || !online_ptrs_host || !numel_host || n_tensors < 0) return NRSE_ERR_INVALID_ARG;
  if (chunk_elems <= 0 || chunk_elems > (1 << 30) || (chunk_elems & 3) != 0) return NRSE_ERR_INVALID_ARG;
  const bool fill = chunk_target_host && chunk_online_host && chunk_numel_host;
  int64_t n = 0;
  for (int i = 0; i < n_tensors; ++i) {
    if (numel_host[i] < 0) return NRSE_ERR_INVALID_ARG;
    for (int64_t off = 0; off < numel_host[i]; off += chunk_elems) {
      if (fill) {
        if (n >= max_chunks) return NRSE_ERR_WORKSPACE;
        const int64_t len = numel_host[i] - off < chunk_elems ? numel_host[i] - off : chunk_elems;
        chunk_target_host[n] = target_ptrs_host[i] + static_cast<uint64_t>(off) * sizeof(float);
        chunk_online_host[n] = online_ptrs_host[i] + static_cast<uint64_t>(off) * sizeof(float);
        chunk_numel_host[n] = static_cast<int32_t>(len);
      }
      ++n;
    }
  }
  return n;
}

int nrse_ema_chunks_f32(const uint64_t* chunk_target, const uint64_t* chunk_online, const int32_t* chunk_numel,
                        int64_t n_chunks, float decay, float one_minus_decay, nrse_stream_t stream) {
  using namespace nrse;
  if (n_chunks < 0) return NRSE_ERR_INVALID_ARG;
  if (n_chunks == 0) return NRSE_OK;
  if (!chunk_target || !chunk_online || !chunk_numel) return NRSE_ERR_INVALID_ARG;
  // 8 resident CTAs of 256 threads per SM: 148 * 8 CTAs, each streaming whole chunks.
  const long long max_grid = static_cast<long long>(kNumSMs) * 8;
  const unsigned grid = static_cast<unsigned>(n_chunks < max_grid ? n_chunks : max_grid);
  ema_chunks_kernel<<<grid, kEmaThreads, 0, as_stream(stream)>>>(chunk_target, chunk_online, chunk_numel,
                                                                 static_cast<long long>(n_chunks), decay,
                                                                 one_minus_decay);
  NRSE_CHECK_LAUNCH();
  return NRSE_OK;
}

}  // extern "C"
