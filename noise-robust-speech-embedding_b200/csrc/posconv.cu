// Positional convolution embedding of WavLM, forward and backward:
//   WavLMPositionalConvEmbedding.forward, hf:models/wavlm/modeling_wavlm.py:48-90 -- Conv1d(1024, 1024, k = 128, padding = 64,
//   groups = 16) with weight-norm over dim 2, drop the last frame (WavLMSamePadLayer, :37-45), exact GELU -- the consumer of
//   the feature projection inside the encoder (reached from ref:src/models/encoder.py:25; on the emotion fine-tune step
//   through ref:src/models/emotion.py:60-79).
//
// The grouped convolution is an implicit GEMM per group g (64 input / 64 output channels, K = 128 taps x 64 channels):
//     out[b, t, g*64 + n] = bias + sum_tap sum_k  X[b, t + tap - 64, g*64 + k] * W[g*64 + n, k, tap]
// on channels-last bf16 activations [B, T, 1024]: the A tile of (group, tap) is a TMA box {64 channels, 128 frames} of X at
// frame offset tap - 64 (a 3-D tensor map (channel, frame, utterance): frames outside [0, T) are zero-filled by the TMA
// engine, which IS the convolution's zero padding -- no padded copy), the B tile an 8 KB block of the pre-packed weights.
//   posconv_gemm_kernel   warp-specialised tcgen05: warp 0 TMA producer (8-stage ring of 24 KB), warp 1 MMA issuer
//                         (128 x 64 x 16 UMMAs, four groups = 256 output channels per tile, two 256-column TMEM accumulator
//                         buffers), two epilogue teams (bias + GELU + fp32 store; the training forward also keeps the
//                         pre-GELU activation in bf16).  The SAME kernel computes the data gradient: dX is the convolution
//                         of dZ with the tap-reversed, transposed weights at frame offset tap - 63.
//   posconv_wgrad_kernel  dW[tap][n][k] = sum_{b,t} dZ[b, t, n] X[b, t + tap - 64, k]: one CTA per (pair of groups, block of
//                         4 taps), both operands MN-major straight from their channels-last homes (as conv_wgrad_kernel),
//                         128 x 128 accumulator per tap (its two diagonal 64 x 64 blocks are the two groups; the
//                         off-diagonal cross-group products are discarded), no split-K: single owner, plain stores.
//   weight-norm           w = g[tap] v / ||v[:, :, tap]||: norms + packing kernels forward, its backward
//                         (dv, dg from dW) as a reduction + a transposing elementwise kernel.
// Geometry: hidden 1024, 16 groups, 128 taps (wavlm-large); other sizes keep the stock module.
#include <cuda.h>

#include "common.cuh"
#include "epilogue_math.cuh"
#include "ptx.cuh"

namespace nrse {
namespace {

constexpr int kH = 1024;    // hidden size
constexpr int kG = 16;      // groups
constexpr int kCG = 64;     // channels per group
constexpr int kTaps = 128;
constexpr int kPcBlockM = 128;
constexpr int kPcUmmaK = 16;

using EncodeTiledFn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                   const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                   CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn pc_encode_fn() {
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

// channels-last bf16 activation [B, T, 1024] as (channel, frame, utterance); box = 64 channels x box_frames frames
int pc_tmap_act(CUtensorMap* m, const void* ptr, int B, int T, int box_frames) {
  EncodeTiledFn enc = pc_encode_fn();
  if (!enc) return NRSE_ERR_CUDA;
  const cuuint64_t dims[3] = {kH, static_cast<cuuint64_t>(T), static_cast<cuuint64_t>(B)};
  const cuuint64_t strides[2] = {kH * 2, static_cast<cuuint64_t>(T) * kH * 2};
  const cuuint32_t box[3] = {kCG, static_cast<cuuint32_t>(box_frames), 1};
  const cuuint32_t estr[3] = {1, 1, 1};
  const CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(ptr), dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? NRSE_OK : NRSE_ERR_CUDA;
}

// packed weights: rows of 64 bf16 (one 128-byte swizzled row per output channel), [(group*128 + tap)*64 + n][k]
int pc_tmap_w(CUtensorMap* m, const void* ptr) {
  EncodeTiledFn enc = pc_encode_fn();
  if (!enc) return NRSE_ERR_CUDA;
  const cuuint64_t dims[2] = {kCG, static_cast<cuuint64_t>(kG) * kTaps * kCG};
  const cuuint64_t strides[1] = {kCG * 2};
  const cuuint32_t box[2] = {kCG, kCG};
  const cuuint32_t estr[2] = {1, 1};
  const CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? NRSE_OK : NRSE_ERR_CUDA;
}

// ---- fp32 -> bf16 cast (the GEMM operand copy of the hidden states) ---------------------------------------------------
__global__ void __launch_bounds__(256) pc_cast_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ y, long long n4) {
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n4; i += stride) {
    const float4 v = ld_stream_f4(reinterpret_cast<const float4*>(x) + i);
    reinterpret_cast<uint2*>(y)[i] = make_uint2(pack_bf16x2(v.x, v.y), pack_bf16x2(v.z, v.w));
  }
}

// ---- weight norm: ||v[:, :, tap]||^2 for the 128 taps; v [1024, 64, 128] (tap fastest) -----------------------------------
__global__ void __launch_bounds__(kTaps) pc_normsq_kernel(const float* __restrict__ v, float* __restrict__ normsq) {
  const int rows_per_block = (kH * kCG) / gridDim.x;
  const float* p = v + static_cast<size_t>(blockIdx.x) * rows_per_block * kTaps + threadIdx.x;
  float acc = 0.f;
  for (int r = 0; r < rows_per_block; ++r) {
    const float a = __ldg(p + static_cast<size_t>(r) * kTaps);
    acc = fmaf(a, a, acc);
  }
  atomicAdd(normsq + threadIdx.x, acc);
}

// w = g[tap] v / ||v||_tap  ->  forward pack  wf[((grp*128 + tap)*64 + n)*64 + k]  = w[grp*64 + n][k][tap]       (blockIdx.y = 0)
//                               backward pack wb[((grp*128 + tp)*64 + k)*64 + n]   = w[grp*64 + n][k][127 - tp]  (blockIdx.y = 1)
// One CTA per (group, fixed row index r): forward r = n (reads v[grp*64 + n][0..63][0..127], contiguous 32 KB), backward
// r = k (reads v[grp*64 + 0..63][k][0..127], 64 runs of 512 B); a [64][128] tile goes through shared memory so that both
// the reads (tap fastest) and the writes (the 64-element GEMM-K index fastest) are coalesced.
__global__ void __launch_bounds__(256) pc_pack_kernel(const float* __restrict__ v, const float* __restrict__ g,
                                                      const float* __restrict__ normsq, __nv_bfloat16* __restrict__ wf,
                                                      __nv_bfloat16* __restrict__ wb) {
  __shared__ float tile[kCG][kTaps + 1];
  __shared__ float scale[kTaps];
  const int grp = blockIdx.x / kCG, r = blockIdx.x % kCG;
  const bool bwd = blockIdx.y == 1;
  if (threadIdx.x < kTaps) scale[threadIdx.x] = g[threadIdx.x] * rsqrtf(normsq[threadIdx.x]);
  __syncthreads();
  for (int i = threadIdx.x; i < kCG * kTaps; i += blockDim.x) {
    const int j = i / kTaps, tap = i % kTaps;  // j = the OTHER channel index (k forward, n backward)
    const size_t src = bwd ? ((static_cast<size_t>(grp) * kCG + j) * kCG + r) * kTaps + tap
                           : ((static_cast<size_t>(grp) * kCG + r) * kCG + j) * kTaps + tap;
    tile[j][tap] = __ldg(v + src) * scale[tap];
  }
  __syncthreads();
  __nv_bfloat16* dst = bwd ? wb : wf;
  for (int i = threadIdx.x; i < kCG * kTaps; i += blockDim.x) {
    const int tp = i / kCG, j = i % kCG;
    const float val = bwd ? tile[j][kTaps - 1 - tp] : tile[j][tp];
    dst[((static_cast<size_t>(grp) * kTaps + tp) * kCG + r) * kCG + j] = __float2bfloat16_rn(val);
  }
}

// =========================================================================================================
// Main implicit GEMM (forward, and data gradient with the tap-reversed / transposed weight pack)
// =========================================================================================================
struct PcCfg {
  static constexpr int kThreads = 64 + 2 * 128;  // warp 0 TMA, warp 1 MMA, two epilogue teams
  static constexpr int kStages = 8;
  static constexpr int kABytes = kPcBlockM * kCG * 2;  // 16 KB: 128 frames x 64 channels
  static constexpr int kBBytes = kCG * kCG * 2;        //  8 KB: 64 output channels x 64 input channels of one tap
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kBiasOff = kStages * kStageBytes;
  static constexpr int kBarOff = kBiasOff + kH * 4;
  static constexpr int kNumBars = 2 * kStages + 4;
  static constexpr int kTmemPtrOff = kBarOff + kNumBars * 8;
  static constexpr int kSmemBytes = kTmemPtrOff + 16 + 1024;
};
static_assert(PcCfg::kSmemBytes <= 227 * 1024, "shared memory");

struct PcArgs {
  float* out;             // [B*T, 1024] fp32
  __nv_bfloat16* zsave;   // nullable: pre-GELU activation (training forward)
  const float* bias;      // nullable (data gradient)
  int B, T, tiles_per_utt, num_tiles;
  int pad;                // 64 forward, 63 data gradient
  int gelu;               // 1 forward, 0 data gradient
};

__global__ void __launch_bounds__(PcCfg::kThreads, 1)
posconv_gemm_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_w, const PcArgs a) {
  using Cfg = PcCfg;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (smem_base - ptx::smem_u32(smem_raw));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  auto bar = [&](int i) { return smem_base + Cfg::kBarOff + 8u * static_cast<uint32_t>(i); };
  const int kFull = 0, kEmpty = Cfg::kStages, kTmemFull = 2 * Cfg::kStages, kTmemEmpty = kTmemFull + 2;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(smem + Cfg::kTmemPtrOff);
  float* s_bias = reinterpret_cast<float*>(smem + Cfg::kBiasOff);

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tmap_x);
    ptx::prefetch_tmap(&tmap_w);
    for (int s = 0; s < Cfg::kStages; ++s) {
      ptx::mbar_init(bar(kFull + s), 1);
      ptx::mbar_init(bar(kEmpty + s), 1);
    }
    for (int b = 0; b < 2; ++b) {
      ptx::mbar_init(bar(kTmemFull + b), 1);
      ptx::mbar_init(bar(kTmemEmpty + b), 128);
    }
    ptx::fence_mbar_init();
  }
  if (warp == 1) {
    ptx::tmem_alloc(ptx::smem_u32(tmem_ptr_smem), 512);
    ptx::tmem_relinquish();
  }
  for (int i = threadIdx.x; i < kH; i += Cfg::kThreads) s_bias[i] = a.bias != nullptr ? a.bias[i] : 0.f;
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  // tile -> (utterance, 128-frame block, quarter of the output channels = four groups); the four quarters of a frame block
  // are neighbours in the tile order, so its input rows are read from HBM once and from L2 three times
  auto decode = [&](int tile, int& b, int& t0, int& q) {
    q = tile & 3;
    const int mt = tile >> 2;
    b = mt / a.tiles_per_utt;
    t0 = (mt % a.tiles_per_utt) * kPcBlockM;
  };

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < a.num_tiles; tile += gridDim.x) {
        int b, t0, q;
        decode(tile, b, t0, q);
        for (int gi = 0; gi < 4; ++gi) {
          const int grp = 4 * q + gi;
          for (int tap = 0; tap < kTaps; ++tap) {
            ptx::mbar_wait(bar(kEmpty + stage), phase ^ 1u);
            const uint32_t a_dst = smem_base + stage * Cfg::kStageBytes;
            ptx::mbar_arrive_expect_tx(bar(kFull + stage), Cfg::kStageBytes);
            // frames outside [0, T) are zero-filled by the TMA engine: the convolution's zero padding
            ptx::tma_load_3d(a_dst, &tmap_x, bar(kFull + stage), grp * kCG, t0 + tap - a.pad, b);
            ptx::tma_load_2d(a_dst + Cfg::kABytes, &tmap_w, bar(kFull + stage), 0, (grp * kTaps + tap) * kCG);
            if (++stage == Cfg::kStages) { stage = 0; phase ^= 1u; }
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = ptx::umma_idesc_bf16(kPcBlockM, kCG);
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int tile = blockIdx.x; tile < a.num_tiles; tile += gridDim.x, ++it) {
        const int buf = it & 1;
        ptx::mbar_wait(bar(kTmemEmpty + buf), (static_cast<uint32_t>(it >> 1) & 1u) ^ 1u);
        ptx::tc_fence_after();
        for (int gi = 0; gi < 4; ++gi) {
          const uint32_t acc = tmem_base + static_cast<uint32_t>(buf * 256 + gi * kCG);
          for (int tap = 0; tap < kTaps; ++tap) {
            ptx::mbar_wait(bar(kFull + stage), phase);
            ptx::tc_fence_after();
            const uint32_t a_src = smem_base + stage * Cfg::kStageBytes;
            const uint32_t b_src = a_src + Cfg::kABytes;
#pragma unroll
            for (int k = 0; k < kCG / kPcUmmaK; ++k)
              ptx::umma_bf16(acc, ptx::umma_desc_sw128(a_src + k * (kPcUmmaK * 2)), ptx::umma_desc_sw128(b_src + k * (kPcUmmaK * 2)),
                             idesc, (tap | k) != 0 ? 1u : 0u);
            ptx::umma_commit(bar(kEmpty + stage));
            if (++stage == Cfg::kStages) { stage = 0; phase ^= 1u; }
          }
        }
        ptx::umma_commit(bar(kTmemFull + buf));
      }
    }
    __syncwarp();
  } else {
    // epilogue: team = accumulator buffer; one thread = one frame of the tile
    const int team = (warp - 2) >> 2;
    const int quad = warp & 3;
    const int row = quad * 32 + lane;
    int it = team;
    for (int tile = blockIdx.x + team * gridDim.x; tile < a.num_tiles; tile += 2 * gridDim.x, it += 2) {
      int b, t0, q;
      decode(tile, b, t0, q);
      const int t = t0 + row;
      const bool store = t < a.T;
      ptx::mbar_wait(bar(kTmemFull + team), static_cast<uint32_t>(it >> 1) & 1u);
      ptx::tc_fence_after();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + static_cast<uint32_t>(team * 256);
      const size_t orow = (static_cast<size_t>(b) * a.T + (store ? t : 0)) * kH + q * 256;
      float* out_row = a.out + orow;
      __nv_bfloat16* z_row = a.zsave != nullptr ? a.zsave + orow : nullptr;
      const float* bias = s_bias + q * 256;
      uint32_t ra[32], rb[32];
      auto emit = [&](const uint32_t (&r)[32], int c) {
        if (!store) return;
        float z[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) z[j] = __uint_as_float(r[j]) + bias[c * 32 + j];
        if (z_row != nullptr) {
          char* zd = reinterpret_cast<char*>(z_row + c * 32);
#pragma unroll
          for (int j = 0; j < 2; ++j)
            st_global_256(zd + 32 * j, pack_bf16x2(z[16 * j], z[16 * j + 1]), pack_bf16x2(z[16 * j + 2], z[16 * j + 3]),
                          pack_bf16x2(z[16 * j + 4], z[16 * j + 5]), pack_bf16x2(z[16 * j + 6], z[16 * j + 7]),
                          pack_bf16x2(z[16 * j + 8], z[16 * j + 9]), pack_bf16x2(z[16 * j + 10], z[16 * j + 11]),
                          pack_bf16x2(z[16 * j + 12], z[16 * j + 13]), pack_bf16x2(z[16 * j + 14], z[16 * j + 15]));
        }
        if (a.gelu) {
#pragma unroll
          for (int j = 0; j < 16; ++j) f2_split(gelu2h(f2_make(0.5f * z[2 * j], 0.5f * z[2 * j + 1])), z[2 * j], z[2 * j + 1]);
        }
        char* od = reinterpret_cast<char*>(out_row + c * 32);
#pragma unroll
        for (int j = 0; j < 4; ++j)
          st_global_256(od + 32 * j, __float_as_uint(z[8 * j]), __float_as_uint(z[8 * j + 1]), __float_as_uint(z[8 * j + 2]),
                        __float_as_uint(z[8 * j + 3]), __float_as_uint(z[8 * j + 4]), __float_as_uint(z[8 * j + 5]),
                        __float_as_uint(z[8 * j + 6]), __float_as_uint(z[8 * j + 7]));
      };
      ptx::tmem_ld32(taddr, ra);
#pragma unroll 1
      for (int c = 0; c < 8; c += 2) {
        ptx::tmem_ld_wait();
        ptx::tmem_ld32(taddr + (c + 1) * 32, rb);
        emit(ra, c);
        ptx::tmem_ld_wait();
        if (c + 2 < 8) {
          ptx::tmem_ld32(taddr + (c + 2) * 32, ra);
        } else {
          ptx::tc_fence_before();
          ptx::mbar_arrive(bar(kTmemEmpty + team));
        }
        emit(rb, c + 1);
      }
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, 512);
  }
}

// ---- backward: dOut -> dZ = dOut gelu'(Z) (bf16 GEMM operand), dbias[c] += sum dZ -----------------------------------------
__global__ void __launch_bounds__(256) pc_bwd_prep_kernel(const float* __restrict__ dout, const __nv_bfloat16* __restrict__ z,
                                                          __nv_bfloat16* __restrict__ dz, float* __restrict__ dbias,
                                                          long long rows) {
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  for (long long m = blockIdx.x; m < rows; m += gridDim.x) {
    const float4 g = __ldg(reinterpret_cast<const float4*>(dout + m * kH) + threadIdx.x);
    const uint2 zz = __ldg(reinterpret_cast<const uint2*>(z + m * kH) + threadIdx.x);
    float d0, d1, d2, d3;
    f2_split(f2_mul(f2_make(g.x, g.y), gelu_grad2(f2_bits(zz.x << 16, zz.x & 0xffff0000u))), d0, d1);
    f2_split(f2_mul(f2_make(g.z, g.w), gelu_grad2(f2_bits(zz.y << 16, zz.y & 0xffff0000u))), d2, d3);
    acc[0] += d0; acc[1] += d1; acc[2] += d2; acc[3] += d3;
    reinterpret_cast<uint2*>(dz + m * kH)[threadIdx.x] = make_uint2(pack_bf16x2(d0, d1), pack_bf16x2(d2, d3));
  }
  if (dbias != nullptr) {
#pragma unroll
    for (int j = 0; j < 4; ++j) atomicAdd(dbias + 4 * threadIdx.x + j, acc[j]);
  }
}

// =========================================================================================================
// Weight gradient
// =========================================================================================================
constexpr int kPwThreads = 192;
constexpr int kPwStages = 2;
constexpr int kPwKm = 64;                       // frames per pipeline stage
constexpr int kPwTapsPerCta = 4;
constexpr int kPwABytes = 2 * kPwKm * 128;      // dZ: 128 channels (two groups) = 2 boxes of {64 ch, 64 frames}
constexpr int kPwBBytes = 2 * kPwKm * 128;      // X at one tap offset: the same two groups
constexpr int kPwStageBytes = kPwABytes + kPwTapsPerCta * kPwBBytes;  // 80 KB
constexpr int kPwBarOff = kPwStages * kPwStageBytes;
constexpr int kPwSmemBytes = kPwBarOff + (2 * kPwStages + 1) * 8 + 16 + 1024;
static_assert(kPwSmemBytes <= 227 * 1024, "shared memory");

__device__ __forceinline__ uint64_t pc_desc_sw128_mn(uint32_t smem_addr) {
  // MN-major, 128B swizzle: 64-element MN blocks are 8 KB apart (LBO), 8-row K groups 1 KB apart (SBO)
  return static_cast<uint64_t>((smem_addr & 0x3FFFFu) >> 4) | (static_cast<uint64_t>((kPwKm * 128) >> 4) << 16) |
         (static_cast<uint64_t>(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}

// dwt [128 taps][1024 n][64 k] fp32, every element written exactly once
__global__ void __launch_bounds__(kPwThreads, 1)
posconv_wgrad_kernel(const __grid_constant__ CUtensorMap tmap_dz, const __grid_constant__ CUtensorMap tmap_x,
                     float* __restrict__ dwt, int B, int T) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (smem_base - ptx::smem_u32(smem_raw));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  auto bar = [&](int i) { return smem_base + kPwBarOff + 8u * static_cast<uint32_t>(i); };
  const int kFull = 0, kEmpty = kPwStages, kDone = 2 * kPwStages;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(smem + kPwBarOff + (2 * kPwStages + 1) * 8);
  const int gp = blockIdx.x / (kTaps / kPwTapsPerCta);           // pair of groups: channels [128 gp, 128 gp + 128)
  const int tap0 = (blockIdx.x % (kTaps / kPwTapsPerCta)) * kPwTapsPerCta;
  const int kblocks = (T + kPwKm - 1) / kPwKm;
  const int n_stages = B * kblocks;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tmap_dz);
    ptx::prefetch_tmap(&tmap_x);
    for (int s = 0; s < kPwStages; ++s) {
      ptx::mbar_init(bar(kFull + s), 1);
      ptx::mbar_init(bar(kEmpty + s), 1);
    }
    ptx::mbar_init(bar(kDone), 1);
    ptx::fence_mbar_init();
  }
  if (warp == 1) {
    ptx::tmem_alloc(ptx::smem_u32(tmem_ptr_smem), 512);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int st = 0; st < n_stages; ++st) {
        const int b = st / kblocks, f0 = (st % kblocks) * kPwKm;
        ptx::mbar_wait(bar(kEmpty + stage), phase ^ 1u);
        const uint32_t a_dst = smem_base + stage * kPwStageBytes;
        ptx::mbar_arrive_expect_tx(bar(kFull + stage), kPwStageBytes);
#pragma unroll
        for (int j = 0; j < 2; ++j) ptx::tma_load_3d(a_dst + j * (kPwKm * 128), &tmap_dz, bar(kFull + stage), gp * 128 + 64 * j, f0, b);
#pragma unroll
        for (int i = 0; i < kPwTapsPerCta; ++i) {
          const uint32_t b_dst = a_dst + kPwABytes + i * kPwBBytes;
#pragma unroll
          for (int j = 0; j < 2; ++j)  // frames outside [0, T): zero fill = the convolution's padding
            ptx::tma_load_3d(b_dst + j * (kPwKm * 128), &tmap_x, bar(kFull + stage), gp * 128 + 64 * j, f0 + tap0 + i - 64, b);
        }
        if (++stage == kPwStages) { stage = 0; phase ^= 1u; }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = ptx::umma_idesc_bf16(128, 128) | (1u << 15) | (1u << 16);  // both operands MN-major
      int stage = 0;
      uint32_t phase = 0;
      for (int st = 0; st < n_stages; ++st) {
        ptx::mbar_wait(bar(kFull + stage), phase);
        ptx::tc_fence_after();
        const uint32_t a_src = smem_base + stage * kPwStageBytes;
#pragma unroll
        for (int i = 0; i < kPwTapsPerCta; ++i) {
          const uint32_t b_src = a_src + kPwABytes + i * kPwBBytes;
#pragma unroll
          for (int k = 0; k < kPwKm / kPcUmmaK; ++k)  // 16 frames (= 16 shared-memory rows = 2 KB) per instruction
            ptx::umma_bf16(tmem_base + static_cast<uint32_t>(i * 128), pc_desc_sw128_mn(a_src + k * (kPcUmmaK * 128)),
                           pc_desc_sw128_mn(b_src + k * (kPcUmmaK * 128)), idesc, (st > 0 || k > 0) ? 1u : 0u);
        }
        ptx::umma_commit(bar(kEmpty + stage));
        if (++stage == kPwStages) { stage = 0; phase ^= 1u; }
      }
      ptx::umma_commit(bar(kDone));
    }
    __syncwarp();
  } else {
    // epilogue: TMEM lane = output channel n within the pair; the 64 columns of ITS group are the wanted block
    const int quad = warp & 3;
    const int nl = quad * 32 + lane;            // 0..127
    const int n = gp * 128 + nl;
    const int col0 = (nl >> 6) * kCG;           // diagonal block: input channels of the same group
    ptx::mbar_wait(bar(kDone), 0);
    ptx::tc_fence_after();
    const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quad * 32) << 16);
#pragma unroll 1
    for (int i = 0; i < kPwTapsPerCta; ++i) {
      float* dst = dwt + (static_cast<size_t>(tap0 + i) * kH + n) * kCG;
#pragma unroll 1
      for (int c = 0; c < 2; ++c) {
        uint32_t r[32];
        ptx::tmem_ld32(taddr + static_cast<uint32_t>(i * 128 + col0 + c * 32), r);
        ptx::tmem_ld_wait();
        char* d = reinterpret_cast<char*>(dst + c * 32);
#pragma unroll
        for (int j = 0; j < 4; ++j)
          st_global_256(d + 32 * j, r[8 * j], r[8 * j + 1], r[8 * j + 2], r[8 * j + 3], r[8 * j + 4], r[8 * j + 5], r[8 * j + 6],
                        r[8 * j + 7]);
      }
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, 512);
  }
}

// ---- weight-norm backward -----------------------------------------------------------------------------------------------
// s[tap] = sum_{n,k} dW[n][k][tap] v[n][k][tap]     (dwt is [tap][n][k], v is [n][k][tap])
__global__ void __launch_bounds__(256) pc_wn_dot_kernel(const float* __restrict__ dwt, const float* __restrict__ v,
                                                        float* __restrict__ s) {
  const int tap = blockIdx.x, part = blockIdx.y, parts = gridDim.y;
  const int per = (kH * kCG) / parts;
  float acc = 0.f;
  for (int i = part * per + threadIdx.x; i < (part + 1) * per; i += blockDim.x)
    acc = fmaf(__ldg(dwt + static_cast<size_t>(tap) * kH * kCG + i), __ldg(v + static_cast<size_t>(i) * kTaps + tap), acc);
  acc = warp_sum(acc);
  __shared__ float red[8];
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int w = 0; w < 8; ++w) t += red[w];
    atomicAdd(s + tap, t);
  }
}

// dv[n][k][tap] += (g/||v||) (dW - v s / ||v||^2);  dg[tap] += s / ||v||  (block (0, 0) only).  One CTA per output channel n:
// the [128 taps][64 k] slab of dwt goes through shared memory so that reads (k fastest) and writes (tap fastest) coalesce.
__global__ void __launch_bounds__(256) pc_wn_bwd_kernel(const float* __restrict__ dwt, const float* __restrict__ v,
                                                        const float* __restrict__ g, const float* __restrict__ normsq,
                                                        const float* __restrict__ s, float* __restrict__ dv,
                                                        float* __restrict__ dg) {
  __shared__ float tile[kTaps][kCG + 1];
  __shared__ float c1[kTaps], c2[kTaps];
  const int n = blockIdx.x;
  if (threadIdx.x < kTaps) {
    const float inv = rsqrtf(normsq[threadIdx.x]);
    c1[threadIdx.x] = g[threadIdx.x] * inv;                    // g / ||v||
    c2[threadIdx.x] = s[threadIdx.x] * inv * inv;              // s / ||v||^2
    if (n == 0 && dg != nullptr) dg[threadIdx.x] += s[threadIdx.x] * inv;
  }
  for (int i = threadIdx.x; i < kTaps * kCG; i += blockDim.x) {
    const int tap = i / kCG, k = i % kCG;
    tile[tap][k] = __ldg(dwt + (static_cast<size_t>(tap) * kH + n) * kCG + k);
  }
  __syncthreads();
  if (dv == nullptr) return;
  for (int i = threadIdx.x; i < kTaps * kCG; i += blockDim.x) {
    const int k = i / kTaps, tap = i % kTaps;
    const size_t idx = (static_cast<size_t>(n) * kCG + k) * kTaps + tap;
    dv[idx] += c1[tap] * (tile[tap][k] - __ldg(v + idx) * c2[tap]);
  }
}

size_t pc_round(size_t n) { return round_up(n, static_cast<size_t>(1024)); }

int pc_launch_gemm(const void* act_bf16, const void* w_packed, const float* bias, float* out, void* zsave, int B, int T,
                   int pad, int gelu, cudaStream_t s) {
  CUtensorMap tx, tw;
  int rc = pc_tmap_act(&tx, act_bf16, B, T, kPcBlockM);
  if (rc != NRSE_OK) return rc;
  rc = pc_tmap_w(&tw, w_packed);
  if (rc != NRSE_OK) return rc;
  static bool attr_set = false;  // benign race: idempotent attribute
  if (!attr_set) {
    NRSE_CUDA_TRY(cudaFuncSetAttribute(posconv_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, PcCfg::kSmemBytes));
    attr_set = true;
  }
  PcArgs a;
  a.out = out;
  a.zsave = reinterpret_cast<__nv_bfloat16*>(zsave);
  a.bias = bias;
  a.B = B; a.T = T;
  a.tiles_per_utt = ceil_div(T, kPcBlockM);
  a.num_tiles = B * a.tiles_per_utt * 4;
  a.pad = pad;
  a.gelu = gelu;
  const int grid = a.num_tiles < kNumSMs ? a.num_tiles : kNumSMs;
  posconv_gemm_kernel<<<grid, PcCfg::kThreads, PcCfg::kSmemBytes, s>>>(tx, tw, a);
  NRSE_CHECK_LAUNCH();
  return NRSE_OK;
}

}  // namespace
}  // namespace nrse

extern "C" {

size_t nrse_pos_conv_pack_bytes(void) { return static_cast<size_t>(nrse::kG) * nrse::kTaps * nrse::kCG * nrse::kCG * 2; }

int nrse_pos_conv_pack(const float* v, const float* g, void* w_fwd, void* w_bwd, float* normsq, nrse_stream_t stream) {
  using namespace nrse;
  if (!v || !g || !w_fwd || !w_bwd || !normsq) return NRSE_ERR_INVALID_ARG;
  cudaStream_t s = as_stream(stream);
  NRSE_CUDA_TRY(cudaMemsetAsync(normsq, 0, kTaps * sizeof(float), s));
  pc_normsq_kernel<<<256, kTaps, 0, s>>>(v, normsq);
  NRSE_CHECK_LAUNCH();
  pc_pack_kernel<<<dim3(kG * kCG, 2), 256, 0, s>>>(v, g, normsq, reinterpret_cast<__nv_bfloat16*>(w_fwd),
                                                   reinterpret_cast<__nv_bfloat16*>(w_bwd));
  NRSE_CHECK_LAUNCH();
  return NRSE_OK;
}

int nrse_pos_conv_fwd(const float* x, const void* w_fwd, const float* bias, float* y, void* x_bf16, void* z_save, int B, int T,
                      nrse_stream_t stream) {
  using namespace nrse;
  if (!x || !w_fwd || !bias || !y || !x_bf16 || B < 1 || T < 1 || static_cast<long long>(B) * T > (1ll << 24))
    return NRSE_ERR_INVALID_ARG;
  if ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(x_bf16) | reinterpret_cast<uintptr_t>(w_fwd)) & 15u)
    return NRSE_ERR_INVALID_ARG;
  if ((reinterpret_cast<uintptr_t>(y) & 31u) || (z_save && (reinterpret_cast<uintptr_t>(z_save) & 31u))) return NRSE_ERR_INVALID_ARG;
  cudaStream_t s = as_stream(stream);
  const long long n4 = static_cast<long long>(B) * T * kH / 4;
  const long long want = ceil_div(n4, 256ll);
  pc_cast_kernel<<<static_cast<unsigned>(want < 8 * kNumSMs ? want : 8 * kNumSMs), 256, 0, s>>>(
      x, reinterpret_cast<__nv_bfloat16*>(x_bf16), n4);
  NRSE_CHECK_LAUNCH();
  return pc_launch_gemm(x_bf16, w_fwd, bias, y, z_save, B, T, /*pad=*/64, /*gelu=*/1, s);
}

/* workspace: dZ bf16 [B*T, 1024] | dW [128][1024][64] fp32 | s [128] fp32 */
size_t nrse_pos_conv_bwd_workspace_bytes(int B, int T) {
  using namespace nrse;
  if (B < 1 || T < 1) return 0;
  return pc_round(static_cast<size_t>(B) * T * kH * 2) + pc_round(static_cast<size_t>(kTaps) * kH * kCG * 4) + pc_round(kTaps * 4);
}

int nrse_pos_conv_bwd(const float* d_y, const void* x_bf16, const void* z_save, const void* w_bwd, const float* v,
                      const float* g, const float* normsq, float* d_x, float* d_v, float* d_g, float* d_bias,
                      void* workspace, int B, int T, nrse_stream_t stream) {
  using namespace nrse;
  if (!d_y || !x_bf16 || !z_save || !w_bwd || !v || !g || !normsq || !workspace || B < 1 || T < 1) return NRSE_ERR_INVALID_ARG;
  if ((d_v == nullptr) != (d_g == nullptr)) return NRSE_ERR_INVALID_ARG;
  if (reinterpret_cast<uintptr_t>(workspace) & 1023u) return NRSE_ERR_INVALID_ARG;
  if (d_x && (reinterpret_cast<uintptr_t>(d_x) & 31u)) return NRSE_ERR_INVALID_ARG;
  cudaStream_t s = as_stream(stream);
  const long long rows = static_cast<long long>(B) * T;
  char* ws = reinterpret_cast<char*>(workspace);
  __nv_bfloat16* dz = reinterpret_cast<__nv_bfloat16*>(ws);
  float* dwt = reinterpret_cast<float*>(ws + pc_round(static_cast<size_t>(rows) * kH * 2));
  float* sdot = reinterpret_cast<float*>(reinterpret_cast<char*>(dwt) + pc_round(static_cast<size_t>(kTaps) * kH * kCG * 4));
  pc_bwd_prep_kernel<<<static_cast<unsigned>(rows < 4 * kNumSMs ? rows : 4 * kNumSMs), 256, 0, s>>>(
      d_y, reinterpret_cast<const __nv_bfloat16*>(z_save), dz, d_bias, rows);
  NRSE_CHECK_LAUNCH();
  int rc;
  if (d_x != nullptr) {  // dX = conv(dZ, reversed / transposed weights), frame offset tap - 63
    rc = pc_launch_gemm(dz, w_bwd, nullptr, d_x, nullptr, B, T, /*pad=*/63, /*gelu=*/0, s);
    if (rc != NRSE_OK) return rc;
  }
  if (d_v == nullptr) return NRSE_OK;
  CUtensorMap tdz, tx;
  rc = pc_tmap_act(&tdz, dz, B, T, kPwKm);
  if (rc != NRSE_OK) return rc;
  rc = pc_tmap_act(&tx, x_bf16, B, T, kPwKm);
  if (rc != NRSE_OK) return rc;
  static bool attr_set = false;  // benign race: idempotent attribute
  if (!attr_set) {
    NRSE_CUDA_TRY(cudaFuncSetAttribute(posconv_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kPwSmemBytes));
    attr_set = true;
  }
  posconv_wgrad_kernel<<<(kG / 2) * (kTaps / kPwTapsPerCta), kPwThreads, kPwSmemBytes, s>>>(tdz, tx, dwt, B, T);
  NRSE_CHECK_LAUNCH();
  NRSE_CUDA_TRY(cudaMemsetAsync(sdot, 0, kTaps * sizeof(float), s));
  pc_wn_dot_kernel<<<dim3(kTaps, 4), 256, 0, s>>>(dwt, v, sdot);
  NRSE_CHECK_LAUNCH();
  pc_wn_bwd_kernel<<<kH, 256, 0, s>>>(dwt, v, g, normsq, sdot, d_v, d_g);
  NRSE_CHECK_LAUNCH();
  return NRSE_OK;
}

}  // extern "C"
