// WavLM convolutional feature encoder, forward (hf:models/wavlm/modeling_wavlm.py:682-789, reached from
// ref:src/models/encoder.py:25):  7 x { Conv1d(bias=False) ; LayerNorm over C | GroupNorm | none ; exact GELU }.
//
// Data layout.  Activations are channels-last bf16 [B * P_i, 512] with a per-utterance frame pitch P_i chosen
// so that P_{i-1} = 2 * P_i for the stride-2 layers 1..6.  Output frame m = b*P_i + t of layer i then reads
// the k_i consecutive input frames 2m .. 2m+k_i-1, and the whole batch is ONE implicit GEMM
//     Y[M = B*P_i, 512] = A[M, k_i*512] * W[512, k_i*512]^T
// with no im2col buffer: A is addressed by a 3-D TMA tensor map (channel, frame parity, frame pair) of the
// previous activation, weights are pre-packed [512, tap*512 + c_in] bf16.
//
// Kernels.
//   layer 0   (C_in = 1, k = 10, stride 5): SIMT, a warp per output frame, 16 channels per lane with the
//             10-tap filters held in registers; LayerNorm statistics by warp shuffles.  ALU/HBM-write bound.
//   layers 1-6: warp-specialised tcgen05 kernel.  warp 0 = TMA producer (128B-swizzled A/W tiles into a
//             multi-stage mbarrier ring), warp 1 = single-thread tcgen05.mma issuer (128 x 256 x 16 bf16 UMMA,
//             fp32 accumulators in TMEM), warps 2-5 = epilogue (tcgen05.ld -> LayerNorm over the 512 channels
//             -> exact-erf GELU -> bf16 -> global).  Two variants:
//               kClusterN = 1: one CTA owns all 512 channels of a 128-frame tile (accumulator = all of TMEM).
//               kClusterN = 2: a 2-CTA cluster splits the channels 256/256; each CTA double-buffers its
//                              accumulator in TMEM so the epilogue of tile j overlaps the MMAs of tile j+1;
//                              the per-frame LayerNorm partials (mean, M2) are exchanged through distributed
//                              shared memory with a remote mbarrier arrive.
#include <cuda.h>

#include <cstdlib>
#include <cstring>

#include "common.cuh"
#include "ptx.cuh"

namespace nrse {
namespace {

constexpr int kC = NRSE_FRONTEND_CHANNELS;  // 512
constexpr int kLayers = NRSE_FRONTEND_LAYERS;
constexpr int kKernel[kLayers] = {10, 3, 3, 3, 3, 2, 2};
constexpr int kStride[kLayers] = {5, 2, 2, 2, 2, 2, 2};
constexpr float kNormEps = 1e-5f;

// ---- exact GELU -----------------------------------------------------------------------------------------
// 0.5 x (1 + erf(x / sqrt 2)), erf by Abramowitz-Stegun 7.1.26 (|abs err| <= 1.5e-7, far below the bf16
// rounding of the stored activation): 2 MUFU (rcp, ex2) + 9 FMA-class instructions.
__device__ __forceinline__ float gelu_erf(float x) {
  const float ax = fabsf(x);
  float t;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(fmaf(0.3275911f * 0.70710678f, ax, 1.0f)));
  float e;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(x * x * (-0.5f * 1.44269504f)));
  float poly = fmaf(1.061405429f, t, -1.453152027f);
  poly = fmaf(poly, t, 1.421413741f);
  poly = fmaf(poly, t, -0.284496736f);
  poly = fmaf(poly, t, 0.254829592f);
  const float erf_abs = fmaf(-poly * t, e, 1.0f);  // erf(|x|/sqrt2)
  const float half_x = 0.5f * x;
  return fmaf(fabsf(half_x), erf_abs, half_x);  // 0.5x + 0.5|x| erf(|x|/sqrt2) = 0.5x(1 + erf(x/sqrt2))
}

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&h);
}

// =========================================================================================================
// Layer 0
// =========================================================================================================
constexpr int kL0Threads = 256;
constexpr int kL0Warps = kL0Threads / 32;
constexpr int kGnSlots = 32;  // partial-sum slots per utterance for the GroupNorm statistics

struct L0Args {
  const float* x;      // [B, L]
  const float* w;      // [512, 10]
  const float* gamma;  // layer mode: [512]; group mode: per-(b,c) scale  [B, 512]
  const float* beta;   //                              per-(b,c) shift  [B, 512]
  __nv_bfloat16* out;  // [B*P0, 512]
  int B, L, T0, P0;
};

// lane owns channels [8*lane, 8*lane+8) and [256 + 8*lane, 256 + 8*lane + 8)
__device__ __forceinline__ int l0_channel(int lane, int j) { return (j < 8 ? 0 : 256) + 8 * lane + (j & 7); }

__device__ __forceinline__ void l0_load_weights(const float* __restrict__ w, int lane, float (&wr)[16][10]) {
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    const float* wc = w + l0_channel(lane, j) * 10;
#pragma unroll
    for (int k = 0; k < 10; ++k) wr[j][k] = __ldg(wc + k);
  }
}

__device__ __forceinline__ void l0_conv(const float* __restrict__ xw, const float (&wr)[16][10], float (&y)[16]) {
  float xv[10];
#pragma unroll
  for (int k = 0; k < 10; ++k) xv[k] = __ldg(xw + k);  // warp-uniform address: one broadcast transaction
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    float a = 0.f;
#pragma unroll
    for (int k = 0; k < 10; ++k) a = fmaf(wr[j][k], xv[k], a);
    y[j] = a;
  }
}

// kGroup = false: LayerNorm over the 512 channels of each frame (wavlm-large).
// kGroup = true : per-(utterance, channel) affine prepared by the statistics kernels below (wavlm-base).
template <bool kGroup>
__global__ void __launch_bounds__(kL0Threads, 1) layer0_kernel(const L0Args a) {
  __shared__ __align__(16) float s_gamma[kC];
  __shared__ __align__(16) float s_beta[kC];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if constexpr (!kGroup) {
    for (int i = threadIdx.x; i < kC; i += kL0Threads) {
      s_gamma[i] = a.gamma[i];
      s_beta[i] = a.beta[i];
    }
    __syncthreads();
  }
  float wr[16][10];
  l0_load_weights(a.w, lane, wr);

  const long long rows = static_cast<long long>(a.B) * a.P0;
  const long long warps_total = static_cast<long long>(gridDim.x) * kL0Warps;
  for (long long m = static_cast<long long>(blockIdx.x) * kL0Warps + warp; m < rows; m += warps_total) {
    const int b = static_cast<int>(m / a.P0), t = static_cast<int>(m % a.P0);
    uint4* orow = reinterpret_cast<uint4*>(a.out + m * kC);
    if (t >= a.T0) {  // pitch padding: keep it finite, it is never read by a valid frame
      orow[lane] = make_uint4(0, 0, 0, 0);
      orow[32 + lane] = make_uint4(0, 0, 0, 0);
      continue;
    }
    float y[16];
    l0_conv(a.x + static_cast<size_t>(b) * a.L + 5 * t, wr, y);
    float v[16];
    if constexpr (!kGroup) {
      float s = 0.f;
#pragma unroll
      for (int j = 0; j < 16; ++j) s += y[j];
      const float mean = warp_sum(s) * (1.0f / kC);
      float q = 0.f;
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const float d = y[j] - mean;
        q = fmaf(d, d, q);
      }
      const float rstd = rsqrtf(warp_sum(q) * (1.0f / kC) + kNormEps);
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const float4* g4 = reinterpret_cast<const float4*>(s_gamma + h * 256 + 8 * lane);
        const float4* b4 = reinterpret_cast<const float4*>(s_beta + h * 256 + 8 * lane);
        const float4 g0 = g4[0], g1 = g4[1], b0 = b4[0], b1 = b4[1];
        const float g[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
        const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
        for (int j = 0; j < 8; ++j) v[h * 8 + j] = fmaf((y[h * 8 + j] - mean) * rstd, g[j], bb[j]);
      }
    } else {
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const float4* g4 = reinterpret_cast<const float4*>(a.gamma + static_cast<size_t>(b) * kC + h * 256 + 8 * lane);
        const float4* b4 = reinterpret_cast<const float4*>(a.beta + static_cast<size_t>(b) * kC + h * 256 + 8 * lane);
        const float4 g0 = __ldg(g4), g1 = __ldg(g4 + 1), b0 = __ldg(b4), b1 = __ldg(b4 + 1);
        const float g[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
        const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
        for (int j = 0; j < 8; ++j) v[h * 8 + j] = fmaf(y[h * 8 + j], g[j], bb[j]);
      }
    }
#pragma unroll
    for (int j = 0; j < 16; ++j) v[j] = gelu_erf(v[j]);
    orow[lane] = make_uint4(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]), pack_bf16x2(v[4], v[5]),
                            pack_bf16x2(v[6], v[7]));
    orow[32 + lane] = make_uint4(pack_bf16x2(v[8], v[9]), pack_bf16x2(v[10], v[11]), pack_bf16x2(v[12], v[13]),
                                 pack_bf16x2(v[14], v[15]));
  }
}

// GroupNorm(512 groups of 1 channel) statistics: per (utterance, channel) sum and sum of squares over time.
// grid = B * kGnSlots / kL0Warps CTAs; warp (b, slot) covers frames slot, slot + kGnSlots, ...
__global__ void __launch_bounds__(kL0Threads, 1) layer0_gn_partial_kernel(const L0Args a, float* __restrict__ part) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int gw = blockIdx.x * kL0Warps + warp;
  const int b = gw / kGnSlots, slot = gw % kGnSlots;
  if (b >= a.B) return;
  float wr[16][10];
  l0_load_weights(a.w, lane, wr);
  float s[16], q[16];
#pragma unroll
  for (int j = 0; j < 16; ++j) s[j] = q[j] = 0.f;
  for (int t = slot; t < a.T0; t += kGnSlots) {
    float y[16];
    l0_conv(a.x + static_cast<size_t>(b) * a.L + 5 * t, wr, y);
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      s[j] += y[j];
      q[j] = fmaf(y[j], y[j], q[j]);
    }
  }
  float* dst = part + (static_cast<size_t>(b) * kGnSlots + slot) * (2 * kC);
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    dst[l0_channel(lane, j)] = s[j];
    dst[kC + l0_channel(lane, j)] = q[j];
  }
}

// (sum, sumsq) partials -> per-(b,c) scale = gamma * rstd, shift = beta - mean * scale
__global__ void layer0_gn_finalize_kernel(const float* __restrict__ part, const float* __restrict__ gamma,
                                          const float* __restrict__ beta, float* __restrict__ scale,
                                          float* __restrict__ shift, int B, int T0) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * kC) return;
  const int b = i / kC, c = i % kC;
  double s = 0.0, q = 0.0;
  for (int slot = 0; slot < kGnSlots; ++slot) {
    const float* src = part + (static_cast<size_t>(b) * kGnSlots + slot) * (2 * kC);
    s += static_cast<double>(src[c]);
    q += static_cast<double>(src[kC + c]);
  }
  const double mean = s / T0;
  const double var = fmax(q / T0 - mean * mean, 0.0);
  const float rstd = static_cast<float>(1.0 / sqrt(var + static_cast<double>(kNormEps)));
  const float sc = gamma[c] * rstd;
  scale[i] = sc;
  shift[i] = beta[c] - static_cast<float>(mean) * sc;
}

// =========================================================================================================
// Weight packing: checkpoint layout [512, 512, k] fp32 -> bf16 [512, k*512], K index = tap*512 + c_in
// =========================================================================================================
__global__ void pack_weights_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ out, int k) {
  const int n = blockIdx.x;
  for (int i = threadIdx.x; i < k * kC; i += blockDim.x) {
    const int tap = i / kC, c = i % kC;
    out[static_cast<size_t>(n) * k * kC + i] = __float2bfloat16_rn(w[(static_cast<size_t>(n) * kC + c) * k + tap]);
  }
}

// =========================================================================================================
// Layers 1..6: tcgen05 implicit GEMM with fused LayerNorm + GELU epilogue
// =========================================================================================================
constexpr int kBlockM = 128;
constexpr int kBlockK = 64;  // 64 bf16 = one 128-byte swizzled row
constexpr int kUmmaK = 16;
constexpr int kUmmaN = 256;
constexpr int kGemmThreads = 192;  // warp 0 TMA, warp 1 MMA, warps 2..5 epilogue
constexpr int kEpiThreads = 128;

template <int kClusterN>
struct GemmCfg {
  static constexpr int kNPC = kC / kClusterN;            // channels per CTA
  static constexpr int kNumMma = kNPC / kUmmaN;          // UMMA instructions per K step
  static constexpr int kAccBufs = 512 / kNPC;            // TMEM accumulator buffers (512 columns total)
  static constexpr int kStages = kClusterN == 2 ? 4 : 2;
  static constexpr int kABytes = kBlockM * kBlockK * 2;  // 16 KB
  static constexpr int kBBytes = kNPC * kBlockK * 2;     // 32 / 64 KB
  static constexpr int kStageBytes = kABytes + kBBytes;
  // after the stage ring: gamma/beta (float2 per channel), LN partials (2 buffers x 128 rows x float2),
  // mbarriers, TMEM base address
  static constexpr int kGbOff = kStages * kStageBytes;
  static constexpr int kStatsOff = kGbOff + kNPC * 8;
  static constexpr int kBarOff = kStatsOff + 2 * kBlockM * 8;
  static constexpr int kNumBars = 2 * kStages + 2 * kAccBufs + 2;
  static constexpr int kTmemPtrOff = kBarOff + kNumBars * 8;
  static constexpr int kSmemBytes = kTmemPtrOff + 16 + 1024;  // + slack for the 1024-byte alignment
};

struct GemmArgs {
  const float* gamma;  // [512] or nullptr (no norm)
  const float* beta;
  void* out;           // [M_total, 512] bf16 or fp32
  int out_f32;
  int M_total;         // rows to produce (B * P_i)
  int num_tiles;       // ceil(M_total / 128)
  int k_stages;        // k_i * 512 / 64
  int stride;          // s_i (frame parity dimension of the A tensor map)
};

template <int kClusterN>
__global__ void __launch_bounds__(kGemmThreads, 1)
conv_gemm_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_w,
                 const GemmArgs g) {
  using Cfg = GemmCfg<kClusterN>;
  extern __shared__ uint8_t smem_raw[];
  // SWIZZLE_128B tiles need 1024-byte alignment (in the shared window, which is what TMA/UMMA see)
  const uint32_t smem_base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (smem_base - ptx::smem_u32(smem_raw));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t cta_rank = kClusterN == 2 ? ptx::cluster_ctarank() : 0u;
  const int n0 = static_cast<int>(cta_rank) * Cfg::kNPC;  // first channel owned by this CTA

  auto bar = [&](int i) { return smem_base + Cfg::kBarOff + 8u * static_cast<uint32_t>(i); };
  const int kFull = 0, kEmpty = Cfg::kStages, kTmemFull = 2 * Cfg::kStages,
            kTmemEmpty = 2 * Cfg::kStages + Cfg::kAccBufs, kStats = 2 * Cfg::kStages + 2 * Cfg::kAccBufs;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(smem + Cfg::kTmemPtrOff);
  float2* s_gb = reinterpret_cast<float2*>(smem + Cfg::kGbOff);
  float2* s_stats = reinterpret_cast<float2*>(smem + Cfg::kStatsOff);

  // ---- one-time setup -------------------------------------------------------------------------------
  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tmap_a);
    ptx::prefetch_tmap(&tmap_w);
    for (int s = 0; s < Cfg::kStages; ++s) {
      ptx::mbar_init(bar(kFull + s), 1);
      ptx::mbar_init(bar(kEmpty + s), 1);
    }
    for (int b = 0; b < Cfg::kAccBufs; ++b) {
      ptx::mbar_init(bar(kTmemFull + b), 1);
      ptx::mbar_init(bar(kTmemEmpty + b), kEpiThreads);
    }
    ptx::mbar_init(bar(kStats + 0), kEpiThreads);
    ptx::mbar_init(bar(kStats + 1), kEpiThreads);
    ptx::fence_mbar_init();
  }
  if (warp == 1) {
    ptx::tmem_alloc(ptx::smem_u32(tmem_ptr_smem), 512);
    ptx::tmem_relinquish();
  }
  const bool has_norm = g.gamma != nullptr;
  for (int i = threadIdx.x; i < Cfg::kNPC; i += kGemmThreads)
    s_gb[i] = has_norm ? make_float2(g.gamma[n0 + i], g.beta[n0 + i]) : make_float2(1.f, 0.f);
  ptx::tc_fence_before();
  if constexpr (kClusterN == 2) ptx::cluster_sync_all();  // peer barriers must be initialised before remote arrives
  else __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  const int first_tile = static_cast<int>(blockIdx.x) / kClusterN;
  const int tile_step = static_cast<int>(gridDim.x) / kClusterN;

  if (warp == 0) {
    // ===== TMA producer ================================================================================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = first_tile; tile < g.num_tiles; tile += tile_step) {
        const int m0 = tile * kBlockM;
        for (int kb = 0; kb < g.k_stages; ++kb) {
          ptx::mbar_wait(bar(kEmpty + stage), phase ^ 1u);
          const uint32_t a_dst = smem_base + stage * Cfg::kStageBytes;
          const uint32_t b_dst = a_dst + Cfg::kABytes;
          ptx::mbar_arrive_expect_tx(bar(kFull + stage), Cfg::kStageBytes);
          // input frame of tap j for output frame m is stride*m + j = stride*(m + j/stride) + j%stride
          const int tap = kb >> 3, c0 = (kb & 7) * kBlockK;
          ptx::tma_load_3d(a_dst, &tmap_a, bar(kFull + stage), c0, tap % g.stride, m0 + tap / g.stride);
#pragma unroll
          for (int h = 0; h < Cfg::kNumMma; ++h)
            ptx::tma_load_2d(b_dst + h * (kUmmaN * kBlockK * 2), &tmap_w, bar(kFull + stage), kb * kBlockK,
                             n0 + h * kUmmaN);
          if (++stage == Cfg::kStages) { stage = 0; phase ^= 1u; }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ===== MMA issuer (one thread) ======================================================================
    if (lane == 0) {
      constexpr uint32_t idesc = ptx::umma_idesc_bf16(kBlockM, kUmmaN);
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int tile = first_tile; tile < g.num_tiles; tile += tile_step, ++it) {
        const int buf = it % Cfg::kAccBufs;
        const uint32_t acc_phase = static_cast<uint32_t>(it / Cfg::kAccBufs) & 1u;
        ptx::mbar_wait(bar(kTmemEmpty + buf), acc_phase ^ 1u);  // epilogue has drained this accumulator
        ptx::tc_fence_after();
        const uint32_t tmem_acc = tmem_base + static_cast<uint32_t>(buf * Cfg::kNPC);
        for (int kb = 0; kb < g.k_stages; ++kb) {
          ptx::mbar_wait(bar(kFull + stage), phase);
          ptx::tc_fence_after();
          const uint32_t a_src = smem_base + stage * Cfg::kStageBytes;
          const uint32_t b_src = a_src + Cfg::kABytes;
#pragma unroll
          for (int k = 0; k < kBlockK / kUmmaK; ++k) {
            const uint64_t da = ptx::umma_desc_sw128(a_src + k * (kUmmaK * 2));
#pragma unroll
            for (int h = 0; h < Cfg::kNumMma; ++h) {
              const uint64_t db = ptx::umma_desc_sw128(b_src + h * (kUmmaN * kBlockK * 2) + k * (kUmmaK * 2));
              ptx::umma_bf16(tmem_acc + h * kUmmaN, da, db, idesc, (kb | k) != 0 ? 1u : 0u);
            }
          }
          ptx::umma_commit(bar(kEmpty + stage));  // frees the smem slot once these MMAs have read it
          if (++stage == Cfg::kStages) { stage = 0; phase ^= 1u; }
        }
        ptx::umma_commit(bar(kTmemFull + buf));  // accumulator complete -> epilogue
      }
    }
    __syncwarp();
  } else {
    // ===== epilogue: TMEM -> LayerNorm -> GELU -> global ================================================
    const int quad = warp & 3;               // TMEM lane quadrant this warp may access
    const int row = quad * 32 + lane;        // accumulator row == TMEM lane
    const uint32_t peer = cta_rank ^ 1u;
    int it = 0;
    for (int tile = first_tile; tile < g.num_tiles; tile += tile_step, ++it) {
      const int buf = it % Cfg::kAccBufs;
      const uint32_t acc_phase = static_cast<uint32_t>(it / Cfg::kAccBufs) & 1u;
      const long long m = static_cast<long long>(tile) * kBlockM + row;
      ptx::mbar_wait(bar(kTmemFull + buf), acc_phase);
      ptx::tc_fence_after();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + static_cast<uint32_t>(buf * Cfg::kNPC);

      float mean = 0.f, rstd = 1.f;
      if (has_norm) {
        // pass 1: shifted sums over this CTA's channels (shift = first element: no cancellation)
        float shift = 0.f, s1 = 0.f, s2 = 0.f;
#pragma unroll 1
        for (int c = 0; c < Cfg::kNPC; c += 32) {
          uint32_t r[32];
          ptx::tmem_ld32(taddr + c, r);
          ptx::tmem_ld_wait();
          if (c == 0) shift = __uint_as_float(r[0]);
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const float d = __uint_as_float(r[j]) - shift;
            s1 += d;
            s2 = fmaf(d, d, s2);
          }
        }
        constexpr float kInvN = 1.0f / Cfg::kNPC;
        float mean_c = shift + s1 * kInvN;
        float m2 = fmaxf(s2 - s1 * s1 * kInvN, 0.f);
        if constexpr (kClusterN == 2) {
          // exchange (mean, M2) of my 256 channels with the peer CTA that holds the other 256 (Chan et al.)
          const int sb = it & 1;
          const uint32_t slot = smem_base + Cfg::kStatsOff + static_cast<uint32_t>((sb * kBlockM + row) * 8);
          ptx::st_cluster_f2(ptx::mapa(slot, peer), mean_c, m2);
          ptx::mbar_arrive_remote_release(ptx::mapa(bar(kStats + sb), peer));
          ptx::mbar_wait_cluster(bar(kStats + sb), static_cast<uint32_t>(it >> 1) & 1u);
          const float2 o = s_stats[sb * kBlockM + row];
          const float delta = mean_c - o.x;
          m2 = m2 + o.y + delta * delta * (0.5f * Cfg::kNPC);
          mean_c = 0.5f * (mean_c + o.x);
        }
        mean = mean_c;
        rstd = rsqrtf(m2 * (1.0f / kC) + kNormEps);
      }

      // pass 2: normalise, GELU, store this row's channels [n0, n0 + kNPC)
      const bool in_range = m < g.M_total;
#pragma unroll 1
      for (int c = 0; c < Cfg::kNPC; c += 32) {
        uint32_t r[32];
        ptx::tmem_ld32(taddr + c, r);
        ptx::tmem_ld_wait();
        if (c + 32 >= Cfg::kNPC) {  // accumulator fully read: hand it back to the MMA warp
          ptx::tc_fence_before();
          ptx::mbar_arrive(bar(kTmemEmpty + buf));
        }
        float v[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const float2 gb = s_gb[c + j];  // warp-uniform address: broadcast
          const float x = has_norm ? fmaf((__uint_as_float(r[j]) - mean) * rstd, gb.x, gb.y) : __uint_as_float(r[j]);
          v[j] = gelu_erf(x);
        }
        if (in_range) {
          if (g.out_f32) {
            float4* dst = reinterpret_cast<float4*>(reinterpret_cast<float*>(g.out) + m * kC + n0 + c);
#pragma unroll
            for (int j = 0; j < 8; ++j) dst[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
          } else {
            uint4* dst = reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(g.out) + m * kC + n0 + c);
#pragma unroll
            for (int j = 0; j < 4; ++j)
              dst[j] = make_uint4(pack_bf16x2(v[8 * j], v[8 * j + 1]), pack_bf16x2(v[8 * j + 2], v[8 * j + 3]),
                                  pack_bf16x2(v[8 * j + 4], v[8 * j + 5]), pack_bf16x2(v[8 * j + 6], v[8 * j + 7]));
          }
        }
      }
    }
  }

  // ---- teardown -------------------------------------------------------------------------------------
  ptx::tc_fence_before();
  if constexpr (kClusterN == 2) ptx::cluster_sync_all();  // no CTA may exit while its peer can still write to it
  else __syncthreads();
  if (warp == 1) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, 512);
  }
}

// ---- host side --------------------------------------------------------------------------------------------
using EncodeTiledFn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                   const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                   CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
      p = nullptr;
    return reinterpret_cast<EncodeTiledFn>(p);
  }();
  return fn;
}

// A operand: previous activation [rows_prev, 512] bf16 seen as (channel, frame parity, frame pair).
int make_tmap_a(CUtensorMap* m, const void* act_prev, int64_t rows_prev, int stride) {
  EncodeTiledFn enc = get_encode_fn();
  if (!enc) return NRSE_ERR_CUDA;
  const cuuint64_t dims[3] = {static_cast<cuuint64_t>(kC), static_cast<cuuint64_t>(stride),
                              static_cast<cuuint64_t>(rows_prev / stride)};
  const cuuint64_t strides[2] = {static_cast<cuuint64_t>(kC) * 2, static_cast<cuuint64_t>(kC) * 2 * stride};
  const cuuint32_t box[3] = {kBlockK, 1, kBlockM};
  const cuuint32_t estr[3] = {1, 1, 1};
  const CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(act_prev), dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? NRSE_OK : NRSE_ERR_CUDA;
}

// B operand: packed weights [512, K] bf16, box = 256 output channels x 64 K
int make_tmap_w(CUtensorMap* m, const void* w_packed, int K) {
  EncodeTiledFn enc = get_encode_fn();
  if (!enc) return NRSE_ERR_CUDA;
  const cuuint64_t dims[2] = {static_cast<cuuint64_t>(K), static_cast<cuuint64_t>(kC)};
  const cuuint64_t strides[1] = {static_cast<cuuint64_t>(K) * 2};
  const cuuint32_t box[2] = {kBlockK, kUmmaN};
  const cuuint32_t estr[2] = {1, 1};
  const CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(w_packed), dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? NRSE_OK : NRSE_ERR_CUDA;
}

int g_variant = 2;  // 1: single CTA per tile, 2: 2-CTA cluster splitting the channels (default)

template <int kClusterN>
int launch_gemm(const CUtensorMap& ta, const CUtensorMap& tw, const GemmArgs& g, cudaStream_t stream) {
  using Cfg = GemmCfg<kClusterN>;
  static bool attr_set = false;  // benign race: the attribute is idempotent
  if (!attr_set) {
    NRSE_CUDA_TRY(cudaFuncSetAttribute(conv_gemm_kernel<kClusterN>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       Cfg::kSmemBytes));
    attr_set = true;
  }
  const int max_groups = kNumSMs / kClusterN;  // persistent: one CTA (or CTA pair) per SM (pair)
  const int groups = g.num_tiles < max_groups ? g.num_tiles : max_groups;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(static_cast<unsigned>(groups * kClusterN));
  cfg.blockDim = dim3(kGemmThreads);
  cfg.dynamicSmemBytes = Cfg::kSmemBytes;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = kClusterN;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  NRSE_CUDA_TRY(cudaLaunchKernelEx(&cfg, conv_gemm_kernel<kClusterN>, ta, tw, g));
  return NRSE_OK;
}

int geometry(int L, int32_t* T, int32_t* P) {
  long long t = L;
  for (int i = 0; i < kLayers; ++i) {
    if (t < kKernel[i]) return NRSE_ERR_INVALID_ARG;  // empty output (L < 400)
    t = (t - kKernel[i]) / kStride[i] + 1;
    T[i] = static_cast<int32_t>(t);
  }
  // P_6 = max_i ceil(T_i / 2^(6-i)); P_i = 2^(6-i) * P_6  =>  P_i >= T_i and P_{i-1} = 2 P_i
  long long p6 = 1;
  for (int i = 0; i < kLayers; ++i) {
    const long long f = 1ll << (kLayers - 1 - i);
    const long long need = (T[i] + f - 1) / f;
    p6 = need > p6 ? need : p6;
  }
  for (int i = 0; i < kLayers; ++i) P[i] = static_cast<int32_t>(p6 << (kLayers - 1 - i));
  return NRSE_OK;
}

size_t act_bytes(int B, int P) { return round_up(static_cast<size_t>(B) * P * kC * 2, static_cast<size_t>(1024)); }
size_t gn_part_bytes(int B) { return round_up(static_cast<size_t>(B) * kGnSlots * 2 * kC * 4, static_cast<size_t>(1024)); }
size_t gn_affine_bytes(int B) { return round_up(static_cast<size_t>(B) * kC * 4, static_cast<size_t>(1024)); }

}  // namespace
}  // namespace nrse

extern "C" {

int nrse_conv_frontend_geometry(int L, int32_t* T_out_host, int32_t* P_out_host) {
  if (!T_out_host || !P_out_host || L < 1) return NRSE_ERR_INVALID_ARG;
  return nrse::geometry(L, T_out_host, P_out_host);
}

size_t nrse_conv_frontend_workspace_bytes(int B, int L) {
  int32_t T[nrse::kLayers], P[nrse::kLayers];
  if (B < 1 || nrse::geometry(L, T, P) != NRSE_OK) return 0;
  size_t n = 0;
  for (int i = 0; i < nrse::kLayers - 1; ++i) n += nrse::act_bytes(B, P[i]);
  n += nrse::gn_part_bytes(B) + 2 * nrse::gn_affine_bytes(B);  // GroupNorm statistics (wavlm-base mode)
  return n;
}

int nrse_conv_frontend_set_variant(int variant) {
  if (variant != 1 && variant != 2) return NRSE_ERR_INVALID_ARG;
  nrse::g_variant = variant;
  return NRSE_OK;
}

int nrse_conv_frontend_pack_weights(const float* w, void* w_packed, int k, nrse_stream_t stream) {
  using namespace nrse;
  if (!w || !w_packed || (k != 2 && k != 3)) return NRSE_ERR_INVALID_ARG;
  pack_weights_kernel<<<kC, 256, 0, as_stream(stream)>>>(w, reinterpret_cast<__nv_bfloat16*>(w_packed), k);
  NRSE_CHECK_LAUNCH();
  return NRSE_OK;
}

int nrse_conv_layer0_fwd(const float* x, const float* w0, const float* gamma, const float* beta, int norm_mode,
                         void* out, void* gn_scratch, int B, int L, int T0, int P0, nrse_stream_t stream) {
  using namespace nrse;
  if (!x || !w0 || !gamma || !beta || !out || B < 1 || T0 < 1 || P0 < T0 || L < 5 * (T0 - 1) + 10)
    return NRSE_ERR_INVALID_ARG;
  if (reinterpret_cast<uintptr_t>(out) & 15u) return NRSE_ERR_INVALID_ARG;
  cudaStream_t s = as_stream(stream);
  L0Args a;
  a.x = x; a.w = w0; a.gamma = gamma; a.beta = beta;
  a.out = reinterpret_cast<__nv_bfloat16*>(out);
  a.B = B; a.L = L; a.T0 = T0; a.P0 = P0;
  const long long rows = static_cast<long long>(B) * P0;
  const long long want = ceil_div(rows, static_cast<long long>(kL0Warps));
  const unsigned grid = static_cast<unsigned>(want < kNumSMs ? want : kNumSMs);
  if (norm_mode == NRSE_NORM_LAYER) {
    layer0_kernel<false><<<grid, kL0Threads, 0, s>>>(a);
    NRSE_CHECK_LAUNCH();
    return NRSE_OK;
  }
  if (norm_mode != NRSE_NORM_GROUP || !gn_scratch) return NRSE_ERR_INVALID_ARG;
  float* part = reinterpret_cast<float*>(gn_scratch);
  float* scale = reinterpret_cast<float*>(reinterpret_cast<char*>(gn_scratch) + gn_part_bytes(B));
  float* shift = reinterpret_cast<float*>(reinterpret_cast<char*>(scale) + gn_affine_bytes(B));
  layer0_gn_partial_kernel<<<ceil_div(B * kGnSlots, kL0Warps), kL0Threads, 0, s>>>(a, part);
  NRSE_CHECK_LAUNCH();
  layer0_gn_finalize_kernel<<<ceil_div(B * kC, 256), 256, 0, s>>>(part, gamma, beta, scale, shift, B, T0);
  NRSE_CHECK_LAUNCH();
  a.gamma = scale;
  a.beta = shift;
  layer0_kernel<true><<<grid, kL0Threads, 0, s>>>(a);
  NRSE_CHECK_LAUNCH();
  return NRSE_OK;
}

int nrse_conv_layer_fwd(const void* act_prev, int64_t rows_prev, const void* w_packed, int k, int stride,
                        const float* gamma, const float* beta, void* out, int out_dtype, int64_t rows_out,
                        nrse_stream_t stream) {
  using namespace nrse;
  if (!act_prev || !w_packed || !out || (k != 2 && k != 3) || stride != 2) return NRSE_ERR_INVALID_ARG;
  if ((gamma == nullptr) != (beta == nullptr)) return NRSE_ERR_INVALID_ARG;
  if (rows_out < 1 || rows_prev != rows_out * stride || rows_out > (1ll << 30)) return NRSE_ERR_INVALID_ARG;
  if (out_dtype != NRSE_DTYPE_BF16 && out_dtype != NRSE_DTYPE_F32) return NRSE_ERR_INVALID_ARG;
  if ((reinterpret_cast<uintptr_t>(act_prev) | reinterpret_cast<uintptr_t>(w_packed) |
       reinterpret_cast<uintptr_t>(out)) & 15u)
    return NRSE_ERR_INVALID_ARG;
  CUtensorMap ta, tw;
  int rc = make_tmap_a(&ta, act_prev, rows_prev, stride);
  if (rc != NRSE_OK) return rc;
  rc = make_tmap_w(&tw, w_packed, k * kC);
  if (rc != NRSE_OK) return rc;
  GemmArgs g;
  g.gamma = gamma; g.beta = beta; g.out = out;
  g.out_f32 = out_dtype == NRSE_DTYPE_F32 ? 1 : 0;
  g.M_total = static_cast<int>(rows_out);
  g.num_tiles = static_cast<int>(ceil_div(rows_out, static_cast<int64_t>(kBlockM)));
  g.k_stages = k * kC / kBlockK;
  g.stride = stride;
  return g_variant == 2 ? launch_gemm<2>(ta, tw, g, as_stream(stream)) : launch_gemm<1>(ta, tw, g, as_stream(stream));
}

int nrse_conv_frontend_fwd(const float* x, const nrse_frontend_params* prm, int norm_mode, void* y, int y_dtype,
                           void* workspace, size_t workspace_bytes, uint64_t* acts_out_host, int B, int L,
                           nrse_stream_t stream) {
  using namespace nrse;
  if (!x || !prm || !y || !workspace || B < 1) return NRSE_ERR_INVALID_ARG;
  if (norm_mode != NRSE_NORM_LAYER && norm_mode != NRSE_NORM_GROUP) return NRSE_ERR_INVALID_ARG;
  if (reinterpret_cast<uintptr_t>(workspace) & 1023u) return NRSE_ERR_INVALID_ARG;
  int32_t T[kLayers], P[kLayers];
  int rc = geometry(L, T, P);
  if (rc != NRSE_OK) return rc;
  if (workspace_bytes < nrse_conv_frontend_workspace_bytes(B, L)) return NRSE_ERR_WORKSPACE;

  char* ws = reinterpret_cast<char*>(workspace);
  void* act[kLayers];
  for (int i = 0; i < kLayers - 1; ++i) {
    act[i] = ws;
    ws += act_bytes(B, P[i]);
    if (acts_out_host) acts_out_host[i] = reinterpret_cast<uint64_t>(act[i]);
  }
  act[kLayers - 1] = y;
  void* gn_scratch = ws;

  rc = nrse_conv_layer0_fwd(x, prm->w0, prm->gamma[0], prm->beta[0], norm_mode, act[0], gn_scratch, B, L, T[0], P[0],
                            stream);
  if (rc != NRSE_OK) return rc;
  for (int i = 1; i < kLayers; ++i) {
    const bool norm = norm_mode == NRSE_NORM_LAYER;
    if (norm && (!prm->gamma[i] || !prm->beta[i])) return NRSE_ERR_INVALID_ARG;
    rc = nrse_conv_layer_fwd(act[i - 1], static_cast<int64_t>(B) * P[i - 1], prm->w_packed[i - 1], kKernel[i],
                             kStride[i], norm ? prm->gamma[i] : nullptr, norm ? prm->beta[i] : nullptr, act[i],
                             i == kLayers - 1 ? y_dtype : NRSE_DTYPE_BF16, static_cast<int64_t>(B) * P[i], stream);
    if (rc != NRSE_OK) return rc;
  }
  return NRSE_OK;
}

}  // extern "C"
