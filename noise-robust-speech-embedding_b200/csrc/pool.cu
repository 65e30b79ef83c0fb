// Batched Attentive Statistics Pooling (fp32, HBM-bound), forward and backward.
//
// Replaces the per-utterance Python loop of AttentiveStatisticsPooling.forward (ref:src/models/pool.py:37-58), which for
// each of the B utterances slices x[:feat_len], runs tanh(sap_linear(x)), a matmul with the attention vector, a softmax
// over time and two weighted reductions (~12 launches per utterance: ~770 launches at B = 64).  Here the linear layer
// stays one library GEMM over all B*T frames (the caller runs it), and the rest is two launches forward, three backward:
//   asp_logits_kernel     warp per frame:  e[b,t] = sum_d tanh(hl[b,t,d]) * a[d]                         (reads hl once)
//   asp_stats_kernel      CTA per (utterance, 128-channel slice), 8 warps striding over time: softmax of e[b,:len_b]
//                         (recomputed per CTA, T values), mu[d] = sum_t w_t x[t,d],
//                         rh[d] = sqrt(clamp(sum_t w_t x[t,d]^2 - mu[d]^2, 1e-5))                        (reads x once)
//   asp_bwd_dw_kernel     CTA per (utterance, 32 frames): dw[b,t] = sum_d x (dmu' + x dm2), with the per-utterance vectors
//                         dmu' = dmu - 2 mu dm2 and dm2 = drh / (2 rh) (0 where the clamp was active) staged in shared memory
//   asp_bwd_dx_kernel     same grid: softmax backward de = w (dw - sum_s w_s dw_s), then dx = w (dmu' + 2 x dm2),
//                         dhl = de * a * (1 - tanh(hl)^2); the attention-vector gradient sum_t de * tanh(hl) is kept in
//                         registers across the CTA's frames and written as ONE partial row per CTA
//   asp_bwd_da_kernel     sums the partial rows in a fixed order: the gradient is deterministic (no atomics anywhere)
// Frames t >= len_b (padding) get zero weight / zero gradients, exactly as the reference's slice drops them.
// Every per-frame pass keeps 4-8 independent 128-bit loads in flight per lane (a 512-channel chunk of x and of hl).
#include <cfloat>

#include "common.cuh"

namespace nrse {
namespace {

constexpr int kPoolThreads = 256;             // 8 warps
constexpr int kPoolWarps = kPoolThreads / 32;
constexpr int kStatsChannels = 128;           // channels per CTA in the statistics kernel (4 per lane)
constexpr int kRowsPerCta = 32;               // frames per CTA in the backward kernels (4 per warp)
constexpr int kChunk = 512;                   // channels per register chunk: 4 float4 per lane
constexpr int kMaxT = 4096;                   // softmax weights staged in shared memory (16 KB)
constexpr int kMaxD = 8192;

__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ float4 zero4() { return make_float4(0.f, 0.f, 0.f, 0.f); }

// tanh with ~1e-7 ABSOLUTE error: 1 - 2 / (exp(2x) + 1) on ex2.approx / rcp.approx (exp overflow -> 1, underflow -> -1).
// The consumers need absolute accuracy (a 1024-term dot product of tanh values; 1 - tanh^2), not relative accuracy near 0.
__device__ __forceinline__ float tanh_fast(float x) {
  const float e = __expf(2.0f * x);
  return 1.0f - __fdividef(2.0f, e + 1.0f);
}

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
  for (int c0 = 0; c0 < D; c0 += kChunk) {
    float4 hv[4], av[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int d = c0 + j * 128 + lane * 4;
      hv[j] = d < D ? ld_stream_f4(reinterpret_cast<const float4*>(h + d)) : zero4();
      av[j] = d < D ? ld4(att + d) : zero4();
    }
#pragma unroll
    for (int j = 0; j < 4; ++j)
      acc += tanh_fast(hv[j].x) * av[j].x + tanh_fast(hv[j].y) * av[j].y + tanh_fast(hv[j].z) * av[j].z +
             tanh_fast(hv[j].w) * av[j].w;
  }
  acc = warp_sum(acc);
  if (lane == 0) logits[row] = acc;
}

// softmax over the valid frames of utterance b into shared memory: s_w[t] = w_t for t < len
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

__global__ void __launch_bounds__(kPoolThreads) asp_stats_kernel(const float* __restrict__ x,
                                                                 const float* __restrict__ logits,
                                                                 const int32_t* __restrict__ lens,
                                                                 float* __restrict__ out, float* __restrict__ w_out,
                                                                 int B, int T, int D) {
  __shared__ float s_w[kMaxT];
  __shared__ float s_red[kPoolWarps];
  __shared__ float4 s_part[2][kPoolWarps][32];  // per-warp partial (mu, m2) of the lane's 4 channels
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int b = blockIdx.y;
  const int d = blockIdx.x * kStatsChannels + lane * 4;
  const int len = lens[b];
  softmax_to_smem(logits + static_cast<long long>(b) * T, len, s_w, s_red);
  if (blockIdx.x == 0)  // the weights are saved once per utterance for the backward
    for (int t = threadIdx.x; t < T; t += kPoolThreads) w_out[static_cast<long long>(b) * T + t] = t < len ? s_w[t] : 0.f;
  float4 mu = zero4(), m2 = zero4();
  if (d < D) {
    const float* xb = x + static_cast<long long>(b) * T * D + d;
    // warp w takes frames w, w + 8, ...; four frames (independent 128-bit loads) in flight per lane
    for (int t0 = warp; t0 < len; t0 += 4 * kPoolWarps) {
      float4 xv[4];
      float wv[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int t = t0 + j * kPoolWarps;
        xv[j] = t < len ? ld_stream_f4(reinterpret_cast<const float4*>(xb + static_cast<long long>(t) * D)) : zero4();
        wv[j] = t < len ? s_w[t] : 0.f;
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        mu.x += xv[j].x * wv[j]; m2.x += xv[j].x * xv[j].x * wv[j];
        mu.y += xv[j].y * wv[j]; m2.y += xv[j].y * xv[j].y * wv[j];
        mu.z += xv[j].z * wv[j]; m2.z += xv[j].z * xv[j].z * wv[j];
        mu.w += xv[j].w * wv[j]; m2.w += xv[j].w * xv[j].w * wv[j];
      }
    }
  }
  s_part[0][warp][lane] = mu;
  s_part[1][warp][lane] = m2;
  __syncthreads();
  if (warp == 0 && d < D) {
    mu = zero4();
    m2 = zero4();
    for (int w = 0; w < kPoolWarps; ++w) {  // fixed order: deterministic
      const float4 a = s_part[0][w][lane], c = s_part[1][w][lane];
      mu.x += a.x; mu.y += a.y; mu.z += a.z; mu.w += a.w;
      m2.x += c.x; m2.y += c.y; m2.z += c.z; m2.w += c.w;
    }
    float* ob = out + static_cast<long long>(b) * 2 * D;
    *reinterpret_cast<float4*>(ob + d) = mu;
    // pool.py:55: sqrt((sum(x^2 w) - mu^2).clamp(min=1e-5))
    *reinterpret_cast<float4*>(ob + D + d) =
        make_float4(sqrtf(fmaxf(m2.x - mu.x * mu.x, 1e-5f)), sqrtf(fmaxf(m2.y - mu.y * mu.y, 1e-5f)),
                    sqrtf(fmaxf(m2.z - mu.z * mu.z, 1e-5f)), sqrtf(fmaxf(m2.w - mu.w * mu.w, 1e-5f)));
  }
}

// Per-utterance gradient vectors of the pooled statistics into shared memory:
//   s_g[d] = dmu' = dmu - 2 mu dm2,  s_g[D + d] = dm2 = drh / (2 rh) if the clamp was inactive, else 0.
// A clamped variance gives exactly sqrtf(1e-5f) in the forward (same instruction), so `rh > sqrtf(1e-5f)` recovers "the
// clamp was inactive" without saving the variance.
__device__ __forceinline__ void stage_stat_grads(const float* __restrict__ out_b, const float* __restrict__ dout_b,
                                                 int D, float* s_g) {
  for (int d = threadIdx.x; d < D; d += blockDim.x) {
    const float mu = out_b[d], rh = out_b[D + d];
    const float dm2 = rh > sqrtf(1e-5f) ? dout_b[D + d] / (2.0f * rh) : 0.f;
    s_g[d] = dout_b[d] - 2.0f * mu * dm2;
    s_g[D + d] = dm2;
  }
}

// grid (ceil(T / kRowsPerCta), B); dynamic shared memory: 2 * D floats
__global__ void __launch_bounds__(kPoolThreads) asp_bwd_dw_kernel(const float* __restrict__ x,
                                                                  const float* __restrict__ out,
                                                                  const float* __restrict__ dout,
                                                                  const int32_t* __restrict__ lens,
                                                                  float* __restrict__ dw, int B, int T, int D) {
  extern __shared__ float s_g[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int b = blockIdx.y;
  const int len = lens[b];
  if (blockIdx.x * kRowsPerCta >= len) return;  // whole CTA in the padding
  stage_stat_grads(out + static_cast<long long>(b) * 2 * D, dout + static_cast<long long>(b) * 2 * D, D, s_g);
  __syncthreads();
  for (int r = warp; r < kRowsPerCta; r += kPoolWarps) {
    const int t = blockIdx.x * kRowsPerCta + r;
    if (t >= len) break;
    const long long row = static_cast<long long>(b) * T + t;
    const float* xr = x + row * D;
    float acc = 0.f;
    for (int c0 = 0; c0 < D; c0 += kChunk) {
      float4 xv[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int d = c0 + j * 128 + lane * 4;
        xv[j] = d < D ? ld4(xr + d) : zero4();  // plain load: the dx pass reads the row again (L2)
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int d = c0 + j * 128 + lane * 4;
        if (d < D) {
          const float4 gm = ld4(s_g + d), g2 = ld4(s_g + D + d);
          acc += xv[j].x * (gm.x + xv[j].x * g2.x) + xv[j].y * (gm.y + xv[j].y * g2.y) +
                 xv[j].z * (gm.z + xv[j].z * g2.z) + xv[j].w * (gm.w + xv[j].w * g2.w);
        }
      }
    }
    acc = warp_sum(acc);
    if (lane == 0) dw[row] = acc;
  }
}

// grid (ceil(T / kRowsPerCta), B); dynamic shared memory: 2 * D + kPoolWarps * kChunk floats
__global__ void __launch_bounds__(kPoolThreads) asp_bwd_dx_kernel(const float* __restrict__ x,
                                                                  const float* __restrict__ hl,
                                                                  const float* __restrict__ att,
                                                                  const float* __restrict__ w,
                                                                  const float* __restrict__ dw,
                                                                  const float* __restrict__ out,
                                                                  const float* __restrict__ dout,
                                                                  const int32_t* __restrict__ lens,
                                                                  float* __restrict__ dx, float* __restrict__ dhl,
                                                                  float* __restrict__ da_part, int B, int T, int D) {
  extern __shared__ float s_g[];               // [2 * D] stat gradients, then [kPoolWarps][kChunk] da partials
  float* s_da = s_g + 2 * D;
  __shared__ float s_red[kPoolWarps];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int b = blockIdx.y;
  const int len = lens[b];
  const int cta = blockIdx.y * gridDim.x + blockIdx.x;
  float* da_row = da_part + static_cast<long long>(cta) * D;
  const int t_base = blockIdx.x * kRowsPerCta;
  if (t_base >= len) {  // whole CTA in the padding: zero gradients, zero partial
    for (int r = warp; r < kRowsPerCta; r += kPoolWarps) {
      const int t = t_base + r;
      if (t >= T) break;
      const long long row = static_cast<long long>(b) * T + t;
      for (int d = lane * 4; d < D; d += 128) {
        *reinterpret_cast<float4*>(dx + row * D + d) = zero4();
        *reinterpret_cast<float4*>(dhl + row * D + d) = zero4();
      }
    }
    for (int d = threadIdx.x; d < D; d += kPoolThreads) da_row[d] = 0.f;
    return;
  }
  stage_stat_grads(out + static_cast<long long>(b) * 2 * D, dout + static_cast<long long>(b) * 2 * D, D, s_g);
  // sum_s w_s dw_s over the utterance (every CTA of the utterance recomputes it: T values, L2-resident)
  const float* wb = w + static_cast<long long>(b) * T;
  const float* dwb = dw + static_cast<long long>(b) * T;
  float part = 0.f;
  for (int s = threadIdx.x; s < len; s += kPoolThreads) part += wb[s] * dwb[s];
  part = warp_sum(part);
  if (lane == 0) s_red[warp] = part;
  __syncthreads();
  float wdw = 0.f;
  for (int k = 0; k < kPoolWarps; ++k) wdw += s_red[k];  // fixed order

  for (int c0 = 0; c0 < D; c0 += kChunk) {
    float4 da[4] = {zero4(), zero4(), zero4(), zero4()};  // this lane's 16 channels of the chunk, over the warp's frames
    float4 av[4], gm[4], g2[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int d = c0 + j * 128 + lane * 4;
      av[j] = d < D ? ld4(att + d) : zero4();
      gm[j] = d < D ? ld4(s_g + d) : zero4();
      g2[j] = d < D ? ld4(s_g + D + d) : zero4();
    }
    for (int r = warp; r < kRowsPerCta; r += kPoolWarps) {
      const int t = t_base + r;
      if (t >= T) break;
      const long long row = static_cast<long long>(b) * T + t;
      float* dxr = dx + row * D;
      float* dhr = dhl + row * D;
      if (t >= len) {  // padded frame: the reference's slice never sees it
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int d = c0 + j * 128 + lane * 4;
          if (d < D) {
            *reinterpret_cast<float4*>(dxr + d) = zero4();
            *reinterpret_cast<float4*>(dhr + d) = zero4();
          }
        }
        continue;
      }
      const float wt = wb[t];
      const float de = wt * (dwb[t] - wdw);  // softmax backward
      float4 xv[4], hv[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {  // 8 independent 128-bit loads in flight
        const int d = c0 + j * 128 + lane * 4;
        xv[j] = d < D ? ld_stream_f4(reinterpret_cast<const float4*>(x + row * D + d)) : zero4();
        hv[j] = d < D ? ld_stream_f4(reinterpret_cast<const float4*>(hl + row * D + d)) : zero4();
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int d = c0 + j * 128 + lane * 4;
        if (d >= D) continue;
        float4 ox, oh;
        const float tx = tanh_fast(hv[j].x), ty = tanh_fast(hv[j].y), tz = tanh_fast(hv[j].z), tw = tanh_fast(hv[j].w);
        ox.x = wt * (gm[j].x + 2.0f * xv[j].x * g2[j].x);
        ox.y = wt * (gm[j].y + 2.0f * xv[j].y * g2[j].y);
        ox.z = wt * (gm[j].z + 2.0f * xv[j].z * g2[j].z);
        ox.w = wt * (gm[j].w + 2.0f * xv[j].w * g2[j].w);
        oh.x = de * av[j].x * (1.0f - tx * tx);
        oh.y = de * av[j].y * (1.0f - ty * ty);
        oh.z = de * av[j].z * (1.0f - tz * tz);
        oh.w = de * av[j].w * (1.0f - tw * tw);
        da[j].x += de * tx; da[j].y += de * ty; da[j].z += de * tz; da[j].w += de * tw;
        *reinterpret_cast<float4*>(dxr + d) = ox;
        *reinterpret_cast<float4*>(dhr + d) = oh;
      }
    }
    // combine the 8 warps' partials of this chunk in a fixed order, one partial row per CTA
#pragma unroll
    for (int j = 0; j < 4; ++j) *reinterpret_cast<float4*>(s_da + warp * kChunk + j * 128 + lane * 4) = da[j];
    __syncthreads();
    for (int k = threadIdx.x; k < kChunk; k += kPoolThreads) {
      if (c0 + k < D) {
        float s = 0.f;
        for (int wi = 0; wi < kPoolWarps; ++wi) s += s_da[wi * kChunk + k];
        da_row[c0 + k] = s;
      }
    }
    __syncthreads();
  }
}

// grad_attention[d] = sum over the per-CTA partial rows, fixed order.  One CTA per 32 channels; its 32 warps stride over
// the rows (4 independent loads in flight each), then one warp adds the 32 warp partials in order.
constexpr int kDaThreads = 1024;
__global__ void __launch_bounds__(kDaThreads) asp_bwd_da_kernel(const float* __restrict__ da_part, int n_rows, int D,
                                                                float* __restrict__ datt) {
  __shared__ float s_p[kDaThreads / 32][33];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int d = blockIdx.x * 32 + lane;
  float s = 0.f;
  if (d < D) {
    constexpr int kW = kDaThreads / 32;
    for (int r0 = warp; r0 < n_rows; r0 += 4 * kW) {
      float v[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int r = r0 + j * kW;
        v[j] = r < n_rows ? da_part[static_cast<long long>(r) * D + d] : 0.f;
      }
      s += (v[0] + v[1]) + (v[2] + v[3]);
    }
  }
  s_p[warp][lane] = s;
  __syncthreads();
  if (warp == 0 && d < D) {
    float t = 0.f;
    for (int w = 0; w < kDaThreads / 32; ++w) t += s_p[w][lane];
    datt[d] = t;
  }
}

}  // namespace
}  // namespace nrse

extern "C" {

size_t nrse_asp_pool_bwd_workspace_bytes(int B, int T, int D) {
  if (B <= 0 || T <= 0 || D <= 0) return 0;
  // dw [B,T] + one attention-gradient partial row per backward CTA
  const size_t ctas = static_cast<size_t>(B) * static_cast<size_t>(nrse::ceil_div(T, nrse::kRowsPerCta));
  return sizeof(float) * (static_cast<size_t>(B) * T + ctas * D);
}

int nrse_asp_pool_fwd(const float* x, const float* hl, const float* attention, const int32_t* lens, float* out,
                      float* weights, float* logits_ws, int B, int T, int D, nrse_stream_t stream) {
  using namespace nrse;
  if (B < 0 || T <= 0 || D <= 0 || (D & 3) != 0 || T > kMaxT || D > kMaxD) return NRSE_ERR_INVALID_ARG;
  if (B == 0) return NRSE_OK;
  if (!x || !hl || !attention || !lens || !out || !weights || !logits_ws) return NRSE_ERR_INVALID_ARG;
  const long long rows = static_cast<long long>(B) * T;
  asp_logits_kernel<<<static_cast<unsigned>(ceil_div<long long>(rows, kPoolWarps)), kPoolThreads, 0, as_stream(stream)>>>(
      hl, attention, lens, logits_ws, B, T, D);
  NRSE_CHECK_LAUNCH();
  dim3 grid(static_cast<unsigned>(ceil_div(D, kStatsChannels)), static_cast<unsigned>(B));
  asp_stats_kernel<<<grid, kPoolThreads, 0, as_stream(stream)>>>(x, logits_ws, lens, out, weights, B, T, D);
  NRSE_CHECK_LAUNCH();
  return NRSE_OK;
}

int nrse_asp_pool_bwd(const float* x, const float* hl, const float* attention, const int32_t* lens, const float* out,
                      const float* weights, const float* grad_out, float* grad_x, float* grad_hl,
                      float* grad_attention, void* workspace, size_t workspace_bytes, int B, int T, int D,
                      nrse_stream_t stream) {
  using namespace nrse;
  if (B < 0 || T <= 0 || D <= 0 || (D & 3) != 0 || T > kMaxT || D > kMaxD) return NRSE_ERR_INVALID_ARG;
  if (!grad_attention) return NRSE_ERR_INVALID_ARG;
  if (B == 0) {
    NRSE_CUDA_TRY(cudaMemsetAsync(grad_attention, 0, sizeof(float) * D, as_stream(stream)));
    return NRSE_OK;
  }
  if (!x || !hl || !attention || !lens || !out || !weights || !grad_out || !grad_x || !grad_hl || !workspace)
    return NRSE_ERR_INVALID_ARG;
  if (workspace_bytes < nrse_asp_pool_bwd_workspace_bytes(B, T, D)) return NRSE_ERR_WORKSPACE;
  float* dw_ws = static_cast<float*>(workspace);
  float* da_part = dw_ws + static_cast<size_t>(B) * T;
  dim3 grid(static_cast<unsigned>(ceil_div(T, kRowsPerCta)), static_cast<unsigned>(B));
  const size_t smem_dw = sizeof(float) * 2 * D;
  const size_t smem_dx = sizeof(float) * (2 * D + kPoolWarps * kChunk);
  static bool attr_set = false;  // benign race: idempotent attributes
  if (!attr_set) {
    NRSE_CUDA_TRY(cudaFuncSetAttribute(asp_bwd_dw_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       static_cast<int>(sizeof(float) * 2 * kMaxD)));
    NRSE_CUDA_TRY(cudaFuncSetAttribute(asp_bwd_dx_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       static_cast<int>(sizeof(float) * (2 * kMaxD + kPoolWarps * kChunk))));
    attr_set = true;
  }
  asp_bwd_dw_kernel<<<grid, kPoolThreads, smem_dw, as_stream(stream)>>>(x, out, grad_out, lens, dw_ws, B, T, D);
  NRSE_CHECK_LAUNCH();
  asp_bwd_dx_kernel<<<grid, kPoolThreads, smem_dx, as_stream(stream)>>>(x, hl, attention, weights, dw_ws, out, grad_out,
                                                                       lens, grad_x, grad_hl, da_part, B, T, D);
  NRSE_CHECK_LAUNCH();
  asp_bwd_da_kernel<<<static_cast<unsigned>(ceil_div(D, 32)), kDaThreads, 0, as_stream(stream)>>>(
      da_part, static_cast<int>(grid.x * grid.y), D, grad_attention);
  NRSE_CHECK_LAUNCH();
  return NRSE_OK;
}

}  // extern "C"
