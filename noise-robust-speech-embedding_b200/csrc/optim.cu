// Fused optimizer tail of the BYOL step (fp32, HBM-bound): gradient-norm clip + AdamW + EMA of the target network.
//
// Replaces three back-to-back stages of the reference's step body (ref:train_byol.py:67-71):
//   torch.nn.utils.clip_grad_norm_(model.parameters(), max_norm=1.0)      :67
//   optimizer.step()                       (torch.optim.AdamW, lr 1e-5, weight_decay 1e-5)      :70
//   model._update_target_network()         (ref:src/models/byol.py:62-73)      :71
// which in stock torch cost a multi-tensor norm + a scaling pass over every gradient, ~10 multi-tensor passes of the
// foreach AdamW and 3 kernels per target tensor (~110 B of HBM traffic per parameter, ~2000 launches).  Here:
//   launch 1  grad_sqnorm_chunks_kernel   reads every gradient once (4 B/param), one fp64 partial per CTA;
//   launch 2  adamw_ema_chunks_kernel     every CTA re-derives the clip coefficient from the partials (fixed order, so
//             it is deterministic and identical in all CTAs), then per element: g*coef, decoupled weight decay,
//             first/second moment, bias-corrected step, and -- for parameters that have a target twin -- the EMA of
//             the freshly updated value.  Reads p, g, m, v, t and writes p, m, v, t: 36 B/param (28 without a twin).
// The arithmetic follows torch's single-tensor AdamW (torch/optim/adam.py::_single_tensor_adam with
// decoupled_weight_decay), with the step-dependent scalars computed on the host in double exactly as that code does
// (see adamw1 for the one deliberate deviation: approximate sqrt / reciprocal in the step term); the EMA uses the
// reference's rounding (two products, one sum, no FMA), so given the updated parameter the target is bit-identical to
// ref:src/models/byol.py:67-68.
#include <cmath>

#include "common.cuh"

namespace nrse {
namespace {

constexpr int kOptThreads = 256;
constexpr int kOptCtasPerSm = 8;
constexpr int kNormUnroll = 4;
constexpr int kOptUnroll = 2;  // float4 groups (p, g, m, v, t) per thread in flight

struct OptTable {
  const uint64_t* p;
  const uint64_t* g;
  const uint64_t* m;
  const uint64_t* v;
  const uint64_t* t;
  const int32_t* numel;
  long long n_chunks;
};

struct AdamScalars {
  float decay_mul;    // 1 - lr * weight_decay
  float w1;           // 1 - beta1   (lerp weight)
  float beta2;
  float omb2;         // 1 - beta2
  float inv_bc2_sqrt; // 1 / sqrt(1 - beta2^step)
  float eps;
  float neg_step;     // -(lr / (1 - beta1^step))
  float ema_decay, ema_omd;
  float max_norm;     // <= 0: no clipping
};

__device__ __forceinline__ double block_sum_double(double v, double* scratch) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  v = warp_sum(v);
  if (lane == 0) scratch[warp] = v;
  __syncthreads();
  double s = 0.0;
  if (warp == 0) {
    s = lane < kOptThreads / 32 ? scratch[lane] : 0.0;
    s = warp_sum(s);
  }
  return s;  // valid in warp 0
}

__global__ void __launch_bounds__(kOptThreads) grad_sqnorm_chunks_kernel(const uint64_t* __restrict__ chunk_g,
                                                                         const int32_t* __restrict__ chunk_numel,
                                                                         long long n_chunks,
                                                                         double* __restrict__ partials) {
  __shared__ double scratch[kOptThreads / 32];
  double total = 0.0;
  for (long long ch = blockIdx.x; ch < n_chunks; ch += gridDim.x) {
    const float* __restrict__ g = reinterpret_cast<const float*>(chunk_g[ch]);
    if (!g) continue;
    const int n = chunk_numel[ch];
    const bool vec_ok = (reinterpret_cast<uintptr_t>(g) & 15u) == 0;
    const int nvec = vec_ok ? (n >> 2) : 0;
    const float4* g4 = reinterpret_cast<const float4*>(g);
    float acc = 0.f;  // <= 64 values per thread per chunk in fp32, fp64 across chunks
    for (int v = threadIdx.x; v < nvec; v += kOptThreads * kNormUnroll) {
      float4 gv[kNormUnroll];
#pragma unroll
      for (int u = 0; u < kNormUnroll; ++u) {
        const int i = v + u * kOptThreads;
        gv[u] = i < nvec ? g4[i] : make_float4(0.f, 0.f, 0.f, 0.f);  // plain load: kept in L2 for the update pass
      }
#pragma unroll
      for (int u = 0; u < kNormUnroll; ++u)
        acc += gv[u].x * gv[u].x + gv[u].y * gv[u].y + gv[u].z * gv[u].z + gv[u].w * gv[u].w;
    }
    for (int i = (nvec << 2) + threadIdx.x; i < n; i += kOptThreads) acc += g[i] * g[i];
    total += static_cast<double>(acc);
  }
  const double s = block_sum_double(total, scratch);
  if (threadIdx.x == 0) partials[blockIdx.x] = s;
}

// One element of torch/optim/adam.py::_single_tensor_adam (decoupled weight decay).  The moments and the weight decay are
// the same IEEE operations torch issues; the step itself, p += -step_size * m / (sqrt(v) / bc2_sqrt + eps), uses the
// hardware approximations (MUFU.SQRT / MUFU.RCP, <= 2 ulp) and a reciprocal of bc2_sqrt: correctly rounded division and
// square root cost ~25 instructions and two conditional slow-path calls per element, which made the kernel latency-
// instead of HBM-bound (52 % of peak).  The approximation perturbs the update term by < 3e-7 of ITS size, i.e. the
// parameter by < 3e-7 * lr relative -- three orders of magnitude inside the 1e-6 tolerance (torch's own CUDA kernels
// differ from its CPU path by the same kind of rounding).
__device__ __forceinline__ float sqrt_approx(float x) {
  float r;
  asm("sqrt.approx.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ float rcp_approx(float x) {
  float r;
  asm("rcp.approx.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ void adamw1(float& p, float g, float& m, float& v, const AdamScalars& a, float coef) {
  g = __fmul_rn(g, coef);                                     // clip_grad_norm_: grads.mul_(clip_coef_clamped)
  p = __fmul_rn(p, a.decay_mul);                              // param.mul_(1 - lr * weight_decay)
  m = __fmaf_rn(a.w1, __fsub_rn(g, m), m);                    // exp_avg.lerp_(grad, 1 - beta1)
  v = __fadd_rn(__fmul_rn(v, a.beta2), __fmul_rn(__fmul_rn(a.omb2, g), g));  // mul_(beta2).addcmul_(g, g, 1 - beta2)
  const float denom = __fmaf_rn(sqrt_approx(v), a.inv_bc2_sqrt, a.eps);      // sqrt(v) / bc2_sqrt + eps
  p = __fmaf_rn(a.neg_step, __fmul_rn(m, rcp_approx(denom)), p);             // param.addcdiv_(m, denom, -step_size)
}
__device__ __forceinline__ float ema1(float t, float o, float decay, float omd) {
  return __fadd_rn(__fmul_rn(decay, t), __fmul_rn(omd, o));
}

__global__ void __launch_bounds__(kOptThreads, 4) adamw_ema_chunks_kernel(const OptTable tab, const AdamScalars a,
                                                                       const double* __restrict__ partials,
                                                                       int n_partials, float* __restrict__ out_norm) {
  __shared__ double scratch[kOptThreads / 32];
  __shared__ float s_coef;
  // clip coefficient: every CTA sums the same partials in the same order
  if (a.max_norm > 0.f) {
    double s = 0.0;
    for (int i = threadIdx.x; i < n_partials; i += kOptThreads) s += partials[i];
    s = block_sum_double(s, scratch);
    if (threadIdx.x == 0) {
      const float total_norm = static_cast<float>(sqrt(s));
      const float c = __fdiv_rn(a.max_norm, __fadd_rn(total_norm, 1e-6f));  // max_norm / (total_norm + 1e-6)
      s_coef = fminf(c, 1.0f);                                              // clamp(max=1.0)
      if (blockIdx.x == 0 && out_norm) out_norm[0] = total_norm;
    }
  } else if (threadIdx.x == 0) {
    s_coef = 1.0f;
  }
  __syncthreads();
  const float coef = s_coef;

  for (long long ch = blockIdx.x; ch < tab.n_chunks; ch += gridDim.x) {
    float* __restrict__ p = reinterpret_cast<float*>(tab.p[ch]);
    const float* __restrict__ g = reinterpret_cast<const float*>(tab.g[ch]);
    float* __restrict__ m = reinterpret_cast<float*>(tab.m[ch]);
    float* __restrict__ v = reinterpret_cast<float*>(tab.v[ch]);
    float* __restrict__ t = reinterpret_cast<float*>(tab.t[ch]);
    const int n = tab.numel[ch];
    const uintptr_t all = reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g) |
                          reinterpret_cast<uintptr_t>(m) | reinterpret_cast<uintptr_t>(v) |
                          reinterpret_cast<uintptr_t>(t);
    const int nvec = (all & 15u) == 0 ? (n >> 2) : 0;
    if (g) {
      for (int i0 = threadIdx.x; i0 < nvec; i0 += kOptThreads * kOptUnroll) {
        float4 pv[kOptUnroll], gv[kOptUnroll], mv[kOptUnroll], vv[kOptUnroll], tv[kOptUnroll];
#pragma unroll
        for (int u = 0; u < kOptUnroll; ++u) {  // all loads of the unrolled group in flight before any arithmetic
          const int i = i0 + u * kOptThreads;
          if (i < nvec) {
            pv[u] = reinterpret_cast<float4*>(p)[i];
            gv[u] = ld_stream_f4(reinterpret_cast<const float4*>(g) + i);  // last use of the gradient
            mv[u] = reinterpret_cast<float4*>(m)[i];
            vv[u] = reinterpret_cast<float4*>(v)[i];
            if (t) tv[u] = reinterpret_cast<float4*>(t)[i];
          }
        }
#pragma unroll
        for (int u = 0; u < kOptUnroll; ++u) {
          const int i = i0 + u * kOptThreads;
          if (i >= nvec) break;
          adamw1(pv[u].x, gv[u].x, mv[u].x, vv[u].x, a, coef);
          adamw1(pv[u].y, gv[u].y, mv[u].y, vv[u].y, a, coef);
          adamw1(pv[u].z, gv[u].z, mv[u].z, vv[u].z, a, coef);
          adamw1(pv[u].w, gv[u].w, mv[u].w, vv[u].w, a, coef);
          reinterpret_cast<float4*>(p)[i] = pv[u];
          reinterpret_cast<float4*>(m)[i] = mv[u];
          reinterpret_cast<float4*>(v)[i] = vv[u];
          if (t) {
            tv[u].x = ema1(tv[u].x, pv[u].x, a.ema_decay, a.ema_omd);
            tv[u].y = ema1(tv[u].y, pv[u].y, a.ema_decay, a.ema_omd);
            tv[u].z = ema1(tv[u].z, pv[u].z, a.ema_decay, a.ema_omd);
            tv[u].w = ema1(tv[u].w, pv[u].w, a.ema_decay, a.ema_omd);
            reinterpret_cast<float4*>(t)[i] = tv[u];
          }
        }
      }
      for (int i = (nvec << 2) + threadIdx.x; i < n; i += kOptThreads) {
        float pv = p[i], mv = m[i], vv = v[i];
        adamw1(pv, g[i], mv, vv, a, coef);
        p[i] = pv; m[i] = mv; v[i] = vv;
        if (t) t[i] = ema1(t[i], pv, a.ema_decay, a.ema_omd);
      }
    } else if (t) {  // parameter without a gradient this step (frozen / unused): AdamW skips it, the EMA does not
      for (int i = threadIdx.x; i < nvec; i += kOptThreads) {
        const float4 pv = ld_stream_f4(reinterpret_cast<const float4*>(p) + i);
        float4 tv = reinterpret_cast<float4*>(t)[i];
        tv.x = ema1(tv.x, pv.x, a.ema_decay, a.ema_omd);
        tv.y = ema1(tv.y, pv.y, a.ema_decay, a.ema_omd);
        tv.z = ema1(tv.z, pv.z, a.ema_decay, a.ema_omd);
        tv.w = ema1(tv.w, pv.w, a.ema_decay, a.ema_omd);
        reinterpret_cast<float4*>(t)[i] = tv;
      }
      for (int i = (nvec << 2) + threadIdx.x; i < n; i += kOptThreads) t[i] = ema1(t[i], p[i], a.ema_decay, a.ema_omd);
    }
  }
}

}  // namespace
}  // namespace nrse

extern "C" {

int64_t nrse_optim_plan_chunks_host(const uint64_t* p_host, const uint64_t* g_host, const uint64_t* m_host,
                                    const uint64_t* v_host, const uint64_t* t_host, const int64_t* numel_host,
                                    int n_tensors, int64_t chunk_elems, uint64_t* chunk_ptrs_host,
                                    int32_t* chunk_numel_host, int64_t max_chunks) {
  if (!p_host || !g_host || !m_host || !v_host || !t_host || !numel_host || n_tensors < 0)
    return NRSE_ERR_INVALID_ARG;
  if (chunk_elems <= 0 || chunk_elems > (1 << 30) || (chunk_elems & 3) != 0) return NRSE_ERR_INVALID_ARG;
  int64_t total = 0;
  for (int i = 0; i < n_tensors; ++i) {
    if (numel_host[i] < 0 || !p_host[i]) return NRSE_ERR_INVALID_ARG;
    if (g_host[i] && (!m_host[i] || !v_host[i])) return NRSE_ERR_INVALID_ARG;
    if (!g_host[i] && !t_host[i]) continue;  // nothing to do for this tensor
    total += (numel_host[i] + chunk_elems - 1) / chunk_elems;
  }
  if (!chunk_ptrs_host || !chunk_numel_host) return total;
  if (total > max_chunks) return NRSE_ERR_WORKSPACE;
  // layout of chunk_ptrs_host: five consecutive arrays of `max_chunks` entries: p | g | m | v | t
  int64_t n = 0;
  for (int i = 0; i < n_tensors; ++i) {
    if (!g_host[i] && !t_host[i]) continue;
    for (int64_t off = 0; off < numel_host[i]; off += chunk_elems) {
      const int64_t len = numel_host[i] - off < chunk_elems ? numel_host[i] - off : chunk_elems;
      const uint64_t b = static_cast<uint64_t>(off) * sizeof(float);
      chunk_ptrs_host[0 * max_chunks + n] = p_host[i] + b;
      chunk_ptrs_host[1 * max_chunks + n] = g_host[i] ? g_host[i] + b : 0;
      chunk_ptrs_host[2 * max_chunks + n] = g_host[i] ? m_host[i] + b : 0;
      chunk_ptrs_host[3 * max_chunks + n] = g_host[i] ? v_host[i] + b : 0;
      chunk_ptrs_host[4 * max_chunks + n] = t_host[i] ? t_host[i] + b : 0;
      chunk_numel_host[n] = static_cast<int32_t>(len);
      ++n;
    }
  }
  return n;
}

int nrse_optim_partials_count(void) { return nrse::kNumSMs * nrse::kOptCtasPerSm; }

int nrse_grad_sqnorm_chunks_f32(const uint64_t* chunk_g, const int32_t* chunk_numel, int64_t n_chunks,
                                double* partials, nrse_stream_t stream) {
  using namespace nrse;
  if (n_chunks < 0 || !partials || (n_chunks > 0 && (!chunk_g || !chunk_numel))) return NRSE_ERR_INVALID_ARG;
  // always the full grid: CTAs without a chunk write a zero partial, so the consumer can sum a fixed count
  grad_sqnorm_chunks_kernel<<<kNumSMs * kOptCtasPerSm, kOptThreads, 0, as_stream(stream)>>>(
      chunk_g, chunk_numel, static_cast<long long>(n_chunks), partials);
  NRSE_CHECK_LAUNCH();
  return NRSE_OK;
}

int nrse_clip_adamw_ema_chunks_f32(const uint64_t* chunk_ptrs, int64_t chunk_pitch, const int32_t* chunk_numel,
                                   int64_t n_chunks, double lr, double beta1, double beta2, double eps,
                                   double weight_decay, int64_t step, double max_grad_norm, double ema_decay,
                                   const double* partials, int n_partials, float* grad_norm_out,
                                   nrse_stream_t stream) {
  using namespace nrse;
  if (n_chunks < 0 || step < 1) return NRSE_ERR_INVALID_ARG;
  if (n_chunks == 0) return NRSE_OK;
  if (!chunk_ptrs || !chunk_numel || chunk_pitch < n_chunks) return NRSE_ERR_INVALID_ARG;
  if (max_grad_norm > 0.0 && (!partials || n_partials <= 0)) return NRSE_ERR_INVALID_ARG;

  OptTable tab;
  tab.p = chunk_ptrs;
  tab.g = chunk_ptrs + chunk_pitch;
  tab.m = chunk_ptrs + 2 * chunk_pitch;
  tab.v = chunk_ptrs + 3 * chunk_pitch;
  tab.t = chunk_ptrs + 4 * chunk_pitch;
  tab.numel = chunk_numel;
  tab.n_chunks = n_chunks;

  // step-dependent scalars in double, as torch/optim/adam.py::_single_tensor_adam computes them in Python
  const double bc1 = 1.0 - std::pow(beta1, static_cast<double>(step));
  const double bc2 = 1.0 - std::pow(beta2, static_cast<double>(step));
  AdamScalars a;
  a.decay_mul = static_cast<float>(1.0 - lr * weight_decay);
  a.w1 = static_cast<float>(1.0 - beta1);
  a.beta2 = static_cast<float>(beta2);
  a.omb2 = static_cast<float>(1.0 - beta2);
  a.inv_bc2_sqrt = static_cast<float>(1.0 / std::sqrt(bc2));
  a.eps = static_cast<float>(eps);
  a.neg_step = static_cast<float>(-(lr / bc1));
  a.ema_decay = static_cast<float>(ema_decay);
  a.ema_omd = static_cast<float>(1.0 - ema_decay);
  a.max_norm = static_cast<float>(max_grad_norm);

  const long long max_grid = static_cast<long long>(kNumSMs) * kOptCtasPerSm;
  const unsigned grid = static_cast<unsigned>(n_chunks < max_grid ? n_chunks : max_grid);
  adamw_ema_chunks_kernel<<<grid, kOptThreads, 0, as_stream(stream)>>>(tab, a, partials, n_partials, grad_norm_out);
  NRSE_CHECK_LAUNCH();
  return NRSE_OK;
}

}  // extern "C"
