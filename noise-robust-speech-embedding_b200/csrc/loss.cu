// BYOL cosine loss, forward and backward, one launch each, no host sync.
//
// Replaces byol_loss (ref:src/models/byol.py:104-129): +1e-10, F.normalize(dim=1, eps=1e-10) on both
// inputs, row dot product, clamp to [-1,1], 2 - 2*mean -- about 15 ATen launches and 4 isnan().any()
// host syncs in the reference.  Forward: one warp per row (128-bit loads, shuffle reductions) keeps
// (||p'||, ||z'||, <p^,z^>) per row; the batch mean is reduced across the CTAs of ONE thread-block
// cluster through distributed shared memory in a fixed order (deterministic, no atomics, no workspace).
// Backward (w.r.t. online_pred only; the target branch is under no_grad, byol.py:94-96):
//   dL/dp = g * (-2/B) * 1[-1 <= s <= 1] * (z^ - s p^) / max(||p'||, eps)
#include <cooperative_groups.h>

#include "common.cuh"

namespace cg = cooperative_groups;

namespace nrse {
namespace {

constexpr int kLossThreads = 512;
constexpr int kLossWarps = kLossThreads / 32;
constexpr int kLossMaxCluster = 8;
constexpr float kEps = 1e-10f;

// 8 consecutive elements of a row as fp32
template <int kDtype>
__device__ __forceinline__ void load8(const void* row, int v, float (&x)[8]) {
  if constexpr (kDtype == NRSE_DTYPE_F32) {
    const float4* q = reinterpret_cast<const float4*>(row) + 2 * v;
    const float4 a = __ldg(q), b = __ldg(q + 1);
    x[0] = a.x; x[1] = a.y; x[2] = a.z; x[3] = a.w;
    x[4] = b.x; x[5] = b.y; x[6] = b.z; x[7] = b.w;
  } else {
    const uint4 a = __ldg(reinterpret_cast<const uint4*>(row) + v);
    const unsigned w[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      x[2 * j] = __uint_as_float(w[j] << 16);
      x[2 * j + 1] = __uint_as_float(w[j] & 0xffff0000u);
    }
  }
}

template <int kDtype>
__device__ __forceinline__ float load1(const void* row, int i) {
  if constexpr (kDtype == NRSE_DTYPE_F32) return __ldg(reinterpret_cast<const float*>(row) + i);
  else return __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(row)[i]);
}

template <int kDtype>
__global__ void __launch_bounds__(kLossThreads) byol_loss_fwd_kernel(const void* __restrict__ p,
                                                                     const void* __restrict__ z,
                                                                     float* __restrict__ loss,
                                                                     float* __restrict__ saved,
                                                                     float* __restrict__ row_sim,
                                                                     int32_t* __restrict__ flags_out, int B, int D) {
  cg::cluster_group cluster = cg::this_cluster();
  const int rank = static_cast<int>(cluster.block_rank());
  const int n_cta = static_cast<int>(cluster.num_blocks());
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  constexpr size_t kElem = kDtype == NRSE_DTYPE_F32 ? 4 : 2;

  __shared__ float warp_part[kLossWarps];
  __shared__ float cta_part[kLossMaxCluster];
  __shared__ unsigned warp_flags[kLossWarps];
  __shared__ unsigned cta_flags[kLossMaxCluster];
  unsigned flags = 0;  // bit 0 / 1: NaN in p / z; bit 2 / 3: Inf in p / z (NaN after normalisation), byol.py:109-122

  const int nvec = D / 8;
  float sim_sum = 0.f;  // this warp's clamped similarities (lane 0 holds the value)
  for (int row = rank * kLossWarps + warp; row < B; row += n_cta * kLossWarps) {
    const char* pr = reinterpret_cast<const char*>(p) + static_cast<size_t>(row) * D * kElem;
    const char* zr = reinterpret_cast<const char*>(z) + static_cast<size_t>(row) * D * kElem;
    float pp = 0.f, zz = 0.f, pz = 0.f;
    for (int v = lane; v < nvec; v += 32) {
      float a[8], b[8];
      load8<kDtype>(pr, v, a);
      load8<kDtype>(zr, v, b);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float x = a[j] + kEps, y = b[j] + kEps;  // byol.py:113-114
        flags |= (x != x ? 1u : 0u) | (y != y ? 2u : 0u) | (fabsf(x) == INFINITY ? 4u : 0u) | (fabsf(y) == INFINITY ? 8u : 0u);
        pp = fmaf(x, x, pp);
        zz = fmaf(y, y, zz);
        pz = fmaf(x, y, pz);
      }
    }
    for (int i = nvec * 8 + lane; i < D; i += 32) {
      const float x = load1<kDtype>(pr, i) + kEps, y = load1<kDtype>(zr, i) + kEps;
      flags |= (x != x ? 1u : 0u) | (y != y ? 2u : 0u) | (fabsf(x) == INFINITY ? 4u : 0u) | (fabsf(y) == INFINITY ? 8u : 0u);
      pp = fmaf(x, x, pp);
      zz = fmaf(y, y, zz);
      pz = fmaf(x, y, pz);
    }
    pp = warp_sum(pp);
    zz = warp_sum(zz);
    pz = warp_sum(pz);
    const float np = sqrtf(pp), nz = sqrtf(zz);
    const float s = pz / (fmaxf(np, kEps) * fmaxf(nz, kEps));  // F.normalize(eps=1e-10), then the row dot
    const float sc = fminf(fmaxf(s, -1.f), 1.f);               // byol.py:126
    if (lane == 0) {
      if (saved) reinterpret_cast<float4*>(saved)[row] = make_float4(np, nz, s, 0.f);
      if (row_sim) row_sim[row] = sc;
      sim_sum += sc;
    }
  }
  flags = warp_or(flags);
  if (lane == 0) {
    warp_part[warp] = sim_sum;
    warp_flags[warp] = flags;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    float s = 0.f;
    unsigned f = 0;
    for (int w = 0; w < kLossWarps; ++w) {
      s += warp_part[w];
      f |= warp_flags[w];
    }
    *cluster.map_shared_rank(&cta_part[rank], 0) = s;  // DSMEM write into the rank-0 CTA
    *cluster.map_shared_rank(&cta_flags[rank], 0) = f;
  }
  cluster.sync();
  if (rank == 0 && threadIdx.x == 0) {
    float s = 0.f;
    unsigned f = 0;
    for (int r = 0; r < n_cta; ++r) {
      s += cta_part[r];
      f |= cta_flags[r];
    }
    *loss = 2.f - 2.f * (s / static_cast<float>(B));  // byol.py:127
    // before normalisation: NaN present; after: a NaN stays NaN and an Inf becomes Inf / Inf = NaN
    if (flags_out != nullptr) *flags_out = static_cast<int32_t>((f & 3u) | (((f | (f >> 2)) & 3u) << 2));
  }
}

template <int kDtype>
__global__ void __launch_bounds__(256) byol_loss_bwd_kernel(const void* __restrict__ p, const void* __restrict__ z,
                                                            const float* __restrict__ saved,
                                                            const float* __restrict__ grad_loss,
                                                            void* __restrict__ grad_p, int B, int D) {
  constexpr size_t kElem = kDtype == NRSE_DTYPE_F32 ? 4 : 2;
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= B) return;
  const float4 sv = __ldg(reinterpret_cast<const float4*>(saved) + row);
  const float dp = fmaxf(sv.x, kEps), dz = fmaxf(sv.y, kEps), s = sv.z;
  const bool pass = (s >= -1.f) && (s <= 1.f);  // clamp backward; NaN similarity gives zero grad like torch
  // d s / d p' = (z^ - s p^)/||p'|| when ||p'|| > eps; when the norm is clamped p^ = p'/eps is linear in p'.
  const float k = pass ? (*grad_loss) * (-2.f / static_cast<float>(B)) / dp : 0.f;
  const float cz = k / dz;                              // coefficient of z'
  const float cp = (sv.x > kEps) ? -k * s / dp : 0.f;  // coefficient of p'
  const char* pr = reinterpret_cast<const char*>(p) + static_cast<size_t>(row) * D * kElem;
  const char* zr = reinterpret_cast<const char*>(z) + static_cast<size_t>(row) * D * kElem;
  char* gr = reinterpret_cast<char*>(grad_p) + static_cast<size_t>(row) * D * kElem;
  const int nvec = D / 8;
  for (int v = lane; v < nvec; v += 32) {
    float a[8], b[8], g[8];
    load8<kDtype>(pr, v, a);
    load8<kDtype>(zr, v, b);
#pragma unroll
    for (int j = 0; j < 8; ++j) g[j] = cz * (b[j] + kEps) + cp * (a[j] + kEps);
    if constexpr (kDtype == NRSE_DTYPE_F32) {
      float4* q = reinterpret_cast<float4*>(gr) + 2 * v;
      q[0] = make_float4(g[0], g[1], g[2], g[3]);
      q[1] = make_float4(g[4], g[5], g[6], g[7]);
    } else {
      uint4 o;
      __nv_bfloat162 h;
      h = __floats2bfloat162_rn(g[0], g[1]); o.x = *reinterpret_cast<unsigned*>(&h);
      h = __floats2bfloat162_rn(g[2], g[3]); o.y = *reinterpret_cast<unsigned*>(&h);
      h = __floats2bfloat162_rn(g[4], g[5]); o.z = *reinterpret_cast<unsigned*>(&h);
      h = __floats2bfloat162_rn(g[6], g[7]); o.w = *reinterpret_cast<unsigned*>(&h);
      reinterpret_cast<uint4*>(gr)[v] = o;
    }
  }
  for (int i = nvec * 8 + lane; i < D; i += 32) {
    const float g = cz * (load1<kDtype>(zr, i) + kEps) + cp * (load1<kDtype>(pr, i) + kEps);
    if constexpr (kDtype == NRSE_DTYPE_F32) reinterpret_cast<float*>(gr)[i] = g;
    else reinterpret_cast<__nv_bfloat16*>(gr)[i] = __float2bfloat16_rn(g);
  }
}

bool loss_args_ok(const void* p, const void* z, int B, int D, int dtype) {
  if (!p || !z || B <= 0 || D <= 0) return false;
  if (dtype != NRSE_DTYPE_F32 && dtype != NRSE_DTYPE_BF16) return false;
  // 128-bit row loads: rows must start 16-byte aligned
  const size_t row_bytes = static_cast<size_t>(D) * (dtype == NRSE_DTYPE_F32 ? 4 : 2);
  if (row_bytes % 16 != 0) return false;
  return ((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(z)) & 15u) == 0;
}

}  // namespace
}  // namespace nrse

extern "C" {

int nrse_byol_loss_fwd(const void* p, const void* z, float* loss, float* saved, float* row_sim, int32_t* flags, int B,
                       int D, int dtype, nrse_stream_t stream) {
  using namespace nrse;
  if (!loss_args_ok(p, z, B, D, dtype) || !loss) return NRSE_ERR_INVALID_ARG;
  if (saved && (reinterpret_cast<uintptr_t>(saved) & 15u)) return NRSE_ERR_INVALID_ARG;
  int n_cta = ceil_div(B, kLossWarps);
  n_cta = n_cta > kLossMaxCluster ? kLossMaxCluster : n_cta;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(n_cta);
  cfg.blockDim = dim3(kLossThreads);
  cfg.stream = as_stream(stream);
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = n_cta;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  if (dtype == NRSE_DTYPE_F32)
    NRSE_CUDA_TRY(cudaLaunchKernelEx(&cfg, byol_loss_fwd_kernel<NRSE_DTYPE_F32>, p, z, loss, saved, row_sim, flags, B, D));
  else
    NRSE_CUDA_TRY(cudaLaunchKernelEx(&cfg, byol_loss_fwd_kernel<NRSE_DTYPE_BF16>, p, z, loss, saved, row_sim, flags, B, D));
  return NRSE_OK;
}

int nrse_byol_loss_bwd(const void* p, const void* z, const float* saved, const float* grad_loss, void* grad_p,
                       int B, int D, int dtype, nrse_stream_t stream) {
  using namespace nrse;
  if (!loss_args_ok(p, z, B, D, dtype) || !saved || !grad_loss || !grad_p) return NRSE_ERR_INVALID_ARG;
  if ((reinterpret_cast<uintptr_t>(saved) | reinterpret_cast<uintptr_t>(grad_p)) & 15u) return NRSE_ERR_INVALID_ARG;
  const int rows_per_cta = 256 / 32;
  const unsigned grid = static_cast<unsigned>(ceil_div(B, rows_per_cta));
  if (dtype == NRSE_DTYPE_F32)
    byol_loss_bwd_kernel<NRSE_DTYPE_F32><<<grid, 256, 0, as_stream(stream)>>>(p, z, saved, grad_loss, grad_p, B, D);
  else
    byol_loss_bwd_kernel<NRSE_DTYPE_BF16><<<grid, 256, 0, as_stream(stream)>>>(p, z, saved, grad_loss, grad_p, B, D);
  NRSE_CHECK_LAUNCH();
  return NRSE_OK;
}

}  // extern "C"
