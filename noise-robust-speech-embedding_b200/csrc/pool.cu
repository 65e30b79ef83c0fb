// Batched Attentive Statistics Pooling (fp32, HBM-bound), forward and backward.
//
// Replaces the per-utterance Python loop of AttentiveStatisticsPooling.forward (ref:src/models/pool.py:37-58), which for
// each of the B utterances slices x[:feat_len], runs tanh(sap_linear(x)), a matmul with the attention vector, a softmax
// over time and two weighted reductions (~12 launches and one .item()-free but serial iteration per utterance: ~770
// launches at B = 64).  Here the linear layer stays one library GEMM over all B*T frames (the caller runs it), and the
// rest is two launches forward, two backward:
//   asp_logits_kernel   one warp per frame:  e[b,t] = sum_d tanh(hl[b,t,d]) * a[d]                (reads hl once)
//   asp_stats_kernel    one CTA per (utterance, 128-channel slice): softmax of e[b,:len_b] (recomputed per CTA, T values),
//                       mu[d] = sum_t w_t x[t,d],  rh[d] = sqrt(clamp(sum_t w_t x[t,d]^2 - mu[d]^2, 1e-5))  (reads x once)
//   asp_bwd_dw_kernel   one warp per frame:  dw[b,t] = sum_d x (dmu' + x dm2)   with dmu' = dmu - 2 mu dm2,
//                       dm2 = drh / (2 rh) where the clamp was inactive, else 0
//   asp_bwd_dx_kernel   one warp per frame:  softmax backward de = w (dw - sum_s w_s dw_s) (recomputed per CTA), then
//                       dx = w (dmu' + 2 x dm2),  dhl = de * a * (1 - tanh(hl)^2),  da += de * tanh(hl) (CTA-reduced atomics)
// Frames t >= len_b (padding) get zero weight / zero gradients, exactly as the reference's slice drops them.
#include <cfloat>

#include "common.cuh"

namespace nrse {
namespace {

constexpr int kPoolThreads = 256;             // 8 warps = 8 frames per CTA in the per-frame kernels
constexpr int kPoolWarps = kPoolThreads / 32;
constexpr int kStatsChannels = 128;           // channels per CTA in the statistics kernel (one per thread)
constexpr int kMaxT = 4096;                   // softmax weights staged in shared memory (16 KB)

__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }

// e[b,t] = sum_d tanh(hl[b,t,d]) * a[d]
__global__ void __launch_bounds__(kPoolThreads) asp_logits_kernel(const float* __restrict__ hl,
                                                                  const float* __restrict__ att,
                                                                  const int32_t* __restrict__ lens,
                                                                  float* __restrict__ logits, int B, int T, int D) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const long long row = static_cast<long long>(blockIdx.x) * kPoolWarps + warp;
  if (row >= static_cast<long long>(B) * T) return;
  const int b = static_cast<int>(row / T), t = static_cast<int>(row % T);
  if (t >= lens[b]) return;
  const float* h = hl + row * D;
  float acc = 0.f;
  for (int d = lane * 4; d < D; d += 128) {
    const float4 hv = ld_stream_f4(reinterpret_cast<const float4*>(h + d));
    const float4 av = ld4(att + d);
    acc += tanhf(hv.x) * av.x + tanhf(hv.y) * av.y + tanhf(hv.z) * av.z + tanhf(hv.w) * av.w;
  }
  acc = warp_sum(acc);
  if (lane == 0) logits[row] = acc;
}

// softmax over the valid frames of utterance b into shared memory; returns nothing, s_w[t] = w_t for t < len
__device__ __forceinline__ void softmax_to_smem(const float* __restrict__ logits_b, int len, float* s_w,
                                                float* s_red) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
  float m = -FLT_MAX;
  for (int t = tid; t < len; t += blockDim.x) {
    const float e = logits_b[t];
    s_w[t] = e;
    m = fmaxf(m, e);
  }
  m = warp_max(m);
  if (lane == 0) s_red[warp] = m;
  __syncthreads();
  m = s_red[0];
  for (int w = 1; w < nwarps; ++w) m = fmaxf(m, s_red[w]);
  __syncthreads();
  float s = 0.f;
  for (int t = tid; t < len; t += blockDim.x) {
    const float p = expf(s_w[t] - m);
    s_w[t] = p;
    s += p;
  }
  s = warp_sum(s);
  if (lane == 0) s_red[warp] = s;
  __syncthreads();
  s = 0.f;
  for (int w = 0; w < nwarps; ++w) s += s_red[w];  // fixed order: deterministic
  const float inv = 1.0f / s;
  __syncthreads();
  for (int t = tid; t < len; t += blockDim.x) s_w[t] *= inv;
  __syncthreads();
}

__global__ void __launch_bounds__(kStatsChannels) asp_stats_kernel(const float* __restrict__ x,
                                                                   const float* __restrict__ logits,
                                                                   const int32_t* __restrict__ lens,
                                                                   float* __restrict__ out, float* __restrict__ w_out,
                                                                   int B, int T, int D) {
  __shared__ float s_w[kMaxT];
  __shared__ float s_red[kStatsChannels / 32];
  const int b = blockIdx.y;
  const int d = blockIdx.x * kStatsChannels + threadIdx.x;
  const int len = lens[b];
  softmax_to_smem(logits + static_cast<long long>(b) * T, len, s_w, s_red);
  if (blockIdx.x == 0)  // the weights are saved once per utterance for the backward
    for (int t = threadIdx.x; t < T; t += kStatsChannels) w_out[static_cast<long long>(b) * T + t] = t < len ? s_w[t] : 0.f;
  if (d >= D) return;
  const float* xb = x + static_cast<long long>(b) * T * D + d;
  float mu = 0.f, m2 = 0.f;
  int t = 0;
  for (; t + 4 <= len; t += 4) {  // four independent loads in flight per thread, coalesced across the channel slice
    const float x0 = xb[static_cast<long long>(t) * D], x1 = xb[static_cast<long long>(t + 1) * D];
    const float x2 = xb[static_cast<long long>(t + 2) * D], x3 = xb[static_cast<long long>(t + 3) * D];
    const float w0 = s_w[t], w1 = s_w[t + 1], w2 = s_w[t + 2], w3 = s_w[t + 3];
    mu += x0 * w0; m2 += x0 * x0 * w0;
    mu += x1 * w1; m2 += x1 * x1 * w1;
    mu += x2 * w2; m2 += x2 * x2 * w2;
    mu += x3 * w3; m2 += x3 * x3 * w3;
  }
  for (; t < len; ++t) {
    const float xv = xb[static_cast<long long>(t) * D], w = s_w[t];
    mu += xv * w;
    m2 += xv * xv * w;
  }
  float* ob = out + static_cast<long long>(b) * 2 * D;
  ob[d] = mu;
  ob[D + d] = sqrtf(fmaxf(m2 - mu * mu, 1e-5f));  // pool.py:55: sqrt((sum(x^2 w) - mu^2).clamp(min=1e-5))
}

// per-channel gradient terms of the pooled statistics: dmu' = dmu - 2 mu dm2, dm2 = drh / (2 rh) if unclamped
__device__ __forceinline__ void stat_grads(const float* __restrict__ out_b, const float* __restrict__ dout_b, int D,
                                           int d, float& dmu_p, float& dm2) {
  const float mu = out_b[d], rh = out_b[D + d];
  // a clamped variance gives exactly sqrtf(1e-5f) in the forward (same instruction), so `rh > sqrtf(1e-5f)` recovers
  // "the clamp was inactive" without saving the variance
  dm2 = rh > sqrtf(1e-5f) ? dout_b[D + d] / (2.0f * rh) : 0.f;
  dmu_p = dout_b[d] - 2.0f * mu * dm2;
}

__global__ void __launch_bounds__(kPoolThreads) asp_bwd_dw_kernel(const float* __restrict__ x,
                                                                  const float* __restrict__ out,
                                                                  const float* __restrict__ dout,
                                                                  const int32_t* __restrict__ lens,
                                                                  float* __restrict__ dw, int B, int T, int D) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const long long row = static_cast<long long>(blockIdx.x) * kPoolWarps + warp;
  if (row >= static_cast<long long>(B) * T) return;
  const int b = static_cast<int>(row / T), t = static_cast<int>(row % T);
  if (t >= lens[b]) return;
  const float* xr = x + row * D;
  const float* ob = out + static_cast<long long>(b) * 2 * D;
  const float* gb = dout + static_cast<long long>(b) * 2 * D;
  float acc = 0.f;
  for (int d0 = lane * 4; d0 < D; d0 += 128) {
    const float4 xv = ld4(xr + d0);
    const float xs[4] = {xv.x, xv.y, xv.z, xv.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float dmu_p, dm2;
      stat_grads(ob, gb, D, d0 + j, dmu_p, dm2);
      acc += xs[j] * (dmu_p + xs[j] * dm2);
    }
  }
  acc = warp_sum(acc);
  if (lane == 0) dw[row] = acc;
}

__global__ void __launch_bounds__(kPoolThreads) asp_bwd_dx_kernel(const float* __restrict__ x,
                                                                  const float* __restrict__ hl,
                                                                  const float* __restrict__ att,
                                                                  const float* __restrict__ w,
                                                                  const float* __restrict__ dw,
                                                                  const float* __restrict__ out,
                                                                  const float* __restrict__ dout,
                                                                  const int32_t* __restrict__ lens,
                                                                  float* __restrict__ dx, float* __restrict__ dhl,
                                                                  float* __restrict__ datt, int B, int T, int D) {
  extern __shared__ float s_da[];  // [D] per-CTA partial of the attention-vector gradient
  __shared__ float s_red[kPoolWarps];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  // grid: (ceil(T / kPoolWarps), B) so that a CTA's frames share the utterance
  const int b = blockIdx.y;
  const int t = blockIdx.x * kPoolWarps + warp;
  const int len = lens[b];
  for (int d = threadIdx.x; d < D; d += kPoolThreads) s_da[d] = 0.f;
  // sum_s w_s dw_s over the utterance (every CTA of the utterance recomputes it: T values)
  const float* wb = w + static_cast<long long>(b) * T;
  const float* dwb = dw + static_cast<long long>(b) * T;
  float part = 0.f;
  for (int s = threadIdx.x; s < len; s += kPoolThreads) part += wb[s] * dwb[s];
  part = warp_sum(part);
  if (lane == 0) s_red[warp] = part;
  __syncthreads();
  float wdw = 0.f;
  for (int k = 0; k < kPoolWarps; ++k) wdw += s_red[k];

  if (t < T) {
    const long long row = static_cast<long long>(b) * T + t;
    float* dxr = dx + row * D;
    float* dhr = dhl + row * D;
    if (t >= len) {  // padded frame: the reference's slice never sees it
      for (int d0 = lane * 4; d0 < D; d0 += 128) {
        *reinterpret_cast<float4*>(dxr + d0) = make_float4(0.f, 0.f, 0.f, 0.f);
        *reinterpret_cast<float4*>(dhr + d0) = make_float4(0.f, 0.f, 0.f, 0.f);
      }
    } else {
      const float wt = wb[t];
      const float de = wt * (dwb[t] - wdw);  // softmax backward
      const float* xr = x + row * D;
      const float* hr = hl + row * D;
      const float* ob = out + static_cast<long long>(b) * 2 * D;
      const float* gb = dout + static_cast<long long>(b) * 2 * D;
      for (int d0 = lane * 4; d0 < D; d0 += 128) {
        const float4 xv = ld_stream_f4(reinterpret_cast<const float4*>(xr + d0));
        const float4 hv = ld_stream_f4(reinterpret_cast<const float4*>(hr + d0));
        const float4 av = ld4(att + d0);
        const float xs[4] = {xv.x, xv.y, xv.z, xv.w}, hs[4] = {hv.x, hv.y, hv.z, hv.w}, as[4] = {av.x, av.y, av.z, av.w};
        float ox[4], oh[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          float dmu_p, dm2;
          stat_grads(ob, gb, D, d0 + j, dmu_p, dm2);
          ox[j] = wt * (dmu_p + 2.0f * xs[j] * dm2);
          const float th = tanhf(hs[j]);
          oh[j] = de * as[j] * (1.0f - th * th);
          atomicAdd(&s_da[d0 + j], de * th);  // shared-memory atomics: 8 warps x distinct channels per lane
        }
        *reinterpret_cast<float4*>(dxr + d0) = make_float4(ox[0], ox[1], ox[2], ox[3]);
        *reinterpret_cast<float4*>(dhr + d0) = make_float4(oh[0], oh[1], oh[2], oh[3]);
      }
    }
  }
  __syncthreads();
  for (int d = threadIdx.x; d < D; d += kPoolThreads)
    if (s_da[d] != 0.f) atomicAdd(datt + d, s_da[d]);
}

}  // namespace
}  // namespace nrse

extern "C" {

int nrse_asp_pool_fwd(const float* x, const float* hl, const float* attention, const int32_t* lens, float* out,
                      float* weights, float* logits_ws, int B, int T, int D, nrse_stream_t stream) {
  using namespace nrse;
  if (B < 0 || T <= 0 || D <= 0 || (D & 3) != 0 || T > kMaxT) return NRSE_ERR_INVALID_ARG;
  if (B == 0) return NRSE_OK;
  if (!x || !hl || !attention || !lens || !out || !weights || !logits_ws) return NRSE_ERR_INVALID_ARG;
  const long long rows = static_cast<long long>(B) * T;
  asp_logits_kernel<<<static_cast<unsigned>(ceil_div<long long>(rows, kPoolWarps)), kPoolThreads, 0, as_stream(stream)>>>(
      hl, attention, lens, logits_ws, B, T, D);
  NRSE_CHECK_LAUNCH();
  dim3 grid(static_cast<unsigned>(ceil_div(D, kStatsChannels)), static_cast<unsigned>(B));
  asp_stats_kernel<<<grid, kStatsChannels, 0, as_stream(stream)>>>(x, logits_ws, lens, out, weights, B, T, D);
  NRSE_CHECK_LAUNCH();
  return NRSE_OK;
}

int nrse_asp_pool_bwd(const float* x, const float* hl, const float* attention, const int32_t* lens, const float* out,
                      const float* weights, const float* grad_out, float* grad_x, float* grad_hl,
                      float* grad_attention, float* dw_ws, int B, int T, int D, nrse_stream_t stream) {
  using namespace nrse;
  if (B < 0 || T <= 0 || D <= 0 || (D & 3) != 0 || T > kMaxT || D > 8192) return NRSE_ERR_INVALID_ARG;
  if (!grad_attention) return NRSE_ERR_INVALID_ARG;
  NRSE_CUDA_TRY(cudaMemsetAsync(grad_attention, 0, sizeof(float) * D, as_stream(stream)));
  if (B == 0) return NRSE_OK;
  if (!x || !hl || !attention || !lens || !out || !weights || !grad_out || !grad_x || !grad_hl || !dw_ws)
    return NRSE_ERR_INVALID_ARG;
  const long long rows = static_cast<long long>(B) * T;
  asp_bwd_dw_kernel<<<static_cast<unsigned>(ceil_div<long long>(rows, kPoolWarps)), kPoolThreads, 0, as_stream(stream)>>>(
      x, out, grad_out, lens, dw_ws, B, T, D);
  NRSE_CHECK_LAUNCH();
  dim3 grid(static_cast<unsigned>(ceil_div(T, kPoolWarps)), static_cast<unsigned>(B));
  asp_bwd_dx_kernel<<<grid, kPoolThreads, sizeof(float) * D, as_stream(stream)>>>(
      x, hl, attention, weights, dw_ws, out, grad_out, lens, grad_x, grad_hl, grad_attention, B, T, D);
  NRSE_CHECK_LAUNCH();
  return NRSE_OK;
}

}  // extern "C"
