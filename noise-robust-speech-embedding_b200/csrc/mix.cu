// Fused SNR mix + peak normalisation + z-normalisation (fp32, HBM-bound).
//
// One thread-block CLUSTER per utterance row.  The reference does this per utterance on a DataLoader
// worker with ~20 tensor ops and three dependent full-row reductions:
//   add_noise_to_speech          ref:src/data/augment.py:4-66
//   peak normalisation           ref:src/data/noisy_speech_dataset.py:88-116
//   zero_mean_unit_var_norm      hf:models/wav2vec2/feature_extraction_wav2vec2.py:95
// Here the row is streamed from HBM once (pass 1: powers, sums, cross term, peaks), re-read from L2 for
// the peak of the mixed signal (pass 2, BYOL mode only) and for the output pass.  Every later statistic
// (mean / variance of both normalised views) is derived algebraically from the pass-1 sums, so the three
// dependent reductions of the reference cost one HBM read.  Partial sums are combined across the CTAs of
// the cluster through distributed shared memory; there is no workspace and no atomics, and the result is
// deterministic.  Algorithmic traffic: 16 B per sample (BYOL mode), 12 B (emotion mode).
#include <cooperative_groups.h>

#include <cmath>

#include "common.cuh"
#include "ptx.cuh"

namespace cg = cooperative_groups;

namespace nrse {
namespace {

constexpr int kMixThreads = 512;
constexpr int kMixWarps = kMixThreads / 32;
constexpr int kMixCluster = 4;  // CTAs per utterance row
constexpr int kMaxSnr = 32;

struct MixParams {
  const float* clean;
  const float* noise;
  const int32_t* snr_idx;
  float* clean_out;
  float* noisy_out;
  int32_t* status;
  int32_t* snr_used;  // nullable [B]: the table index a processed row was mixed at (a retry re-draws it, see noise_shift)
  int B, L, Ln, peak_norm, n_snr;
  int raw;  // 1: write the un-normalised mix c + s*n (add_noise_to_speech alone); peak_norm must be 0
  int retry;        // 1: only rows whose status is != 0 are processed (the others keep their outputs), ...
  int noise_shift;  // ... with the noise AND the SNR draw of row (row + noise_shift) % B: the device-side "try another
                    // noise file, draw another SNR" of ref:src/data/noisy_speech_dataset.py:69-81
  int n_attempts;   // retry launches only: attempts made inside ONE launch (shift noise_shift, noise_shift + 1, ...) while
                    // the row keeps failing -- the reference's `for attempt in range(max_attempts)` loop (:58) on the device
  float snr_lin[kMaxSnr];  // float(10 ** (snr_db / 10)), ref:src/data/augment.py:39
};

// 4 consecutive samples of the clean row and of the length-matched noise row (augment.py:16-21:
// truncate if longer, tile if shorter).  Samples at index >= L read as 0 and contribute nothing.
template <bool kVec>
__device__ __forceinline__ void load4(const float* __restrict__ c_row, const float* __restrict__ n_row,
                                      int L, int Ln, int v, float (&c)[4], float (&n)[4]) {
  if constexpr (kVec) {
    const float4 cv = ld_stream_f4(reinterpret_cast<const float4*>(c_row) + v);
    const float4 nv = ld_stream_f4(reinterpret_cast<const float4*>(n_row) + v);
    c[0] = cv.x; c[1] = cv.y; c[2] = cv.z; c[3] = cv.w;
    n[0] = nv.x; n[1] = nv.y; n[2] = nv.z; n[3] = nv.w;
  } else {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int i = 4 * v + j;
      const bool in = i < L;
      c[j] = in ? __ldg(c_row + i) : 0.f;
      n[j] = in ? __ldg(n_row + (Ln >= L ? i : i % Ln)) : 0.f;
    }
  }
}

template <bool kVec>
__device__ __forceinline__ void store4(float* __restrict__ row, int L, int v, const float (&o)[4]) {
  if constexpr (kVec) {
    reinterpret_cast<float4*>(row)[v] = make_float4(o[0], o[1], o[2], o[3]);
  } else {
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if (4 * v + j < L) row[4 * v + j] = o[j];
  }
}

// Block-wide sum of kN doubles; result valid in every thread.  `scratch` holds kMixWarps*kN doubles.
template <int kN>
__device__ __forceinline__ void block_sum(double (&v)[kN], double* scratch) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < kN; ++k) v[k] = warp_sum(v[k]);
  __syncthreads();  // scratch may still be read from a previous call
  if (lane == 0) {
#pragma unroll
    for (int k = 0; k < kN; ++k) scratch[warp * kN + k] = v[k];
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < kN; ++k) {
    double s = 0.0;
    for (int w = 0; w < kMixWarps; ++w) s += scratch[w * kN + k];  // fixed order: deterministic
    v[k] = s;
  }
}

__device__ __forceinline__ float block_max(float v, float* scratch) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  v = warp_max(v);
  __syncthreads();
  if (lane == 0) scratch[warp] = v;
  __syncthreads();
  float m = scratch[0];
  for (int w = 1; w < kMixWarps; ++w) m = fmaxf(m, scratch[w]);
  return m;
}

// fmaxf drops NaNs, which is what we want for peaks: rows containing NaN are rejected through the sums.
__device__ __forceinline__ float absmax4(float m, const float (&x)[4]) {
  return fmaxf(fmaxf(m, fmaxf(fabsf(x[0]), fabsf(x[1]))), fmaxf(fabsf(x[2]), fabsf(x[3])));
}

// kRetry: the launch redoes rejected rows only, up to p.n_attempts times with successive noise / SNR donors (every CTA
// of a row's cluster derives the same verdict, so all of them leave the loop together).
template <bool kVec, bool kRetry>
__global__ void __launch_bounds__(kMixThreads) mix_normalize_kernel(const MixParams p) {
  cg::cluster_group cluster = cg::this_cluster();
  const int rank = static_cast<int>(cluster.block_rank());
  const int row = blockIdx.x / kMixCluster;
  const int tid = threadIdx.x;

  __shared__ double red_d[kMixWarps * 5];
  __shared__ float red_f[kMixWarps];
  __shared__ double xch1_d[kMixCluster][5];
  __shared__ float xch1_f[kMixCluster][2];
  __shared__ float xch2_f[kMixCluster];
  __shared__ unsigned xch2_u[kMixCluster];

  if (kRetry && p.status[row] == 0) return;  // whole cluster (same row): nothing to redo
  const int L = p.L, Ln = p.Ln;
  const float* c_row = p.clean + static_cast<size_t>(row) * L;
  const int nvec = (L + 3) / 4;
  const int seg = (nvec + kMixCluster - 1) / kMixCluster;
  const int v_begin = rank * seg;
  const int v_end = min(nvec, v_begin + seg);
  const int n_att = kRetry ? p.n_attempts : 1;
  for (int att = 0; att < n_att; ++att) {
  const int donor = kRetry ? (row + p.noise_shift + att) % p.B : row;  // whose noise crop and SNR draw this attempt uses
  const float* n_row = p.noise + static_cast<size_t>(donor) * Ln;
  // exchange buffers / reduction scratch of the previous attempt have been read by every CTA of the cluster
  if (kRetry && att > 0) cluster.sync();

  // ---- pass 1 (HBM): sum c^2, n^2, c, n, c*n; max|c|, max|n| -------------------------------------
  double acc[5] = {0.0, 0.0, 0.0, 0.0, 0.0};
  float cmax = 0.f, nmax_in = 0.f;
  for (int v0 = v_begin + tid; v0 < v_end; v0 += 2 * kMixThreads) {
    float c[2][4], n[2][4];
    const int v1 = v0 + kMixThreads;
    load4<kVec>(c_row, n_row, L, Ln, v0, c[0], n[0]);
    if (v1 < v_end) {
      load4<kVec>(c_row, n_row, L, Ln, v1, c[1], n[1]);
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j) c[1][j] = n[1][j] = 0.f;
    }
    // 8 samples are summed in fp32, then folded into the fp64 running sums (keeps F2F/DADD traffic low
    // while the long accumulation stays in double: no cancellation trouble in var = E[x^2] - mean^2).
    float f[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int u = 0; u < 2; ++u) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        f[0] = fmaf(c[u][j], c[u][j], f[0]);
        f[1] = fmaf(n[u][j], n[u][j], f[1]);
        f[2] += c[u][j];
        f[3] += n[u][j];
        f[4] = fmaf(c[u][j], n[u][j], f[4]);
      }
      cmax = absmax4(cmax, c[u]);
      nmax_in = absmax4(nmax_in, n[u]);
    }
#pragma unroll
    for (int k = 0; k < 5; ++k) acc[k] += static_cast<double>(f[k]);
  }
  block_sum<5>(acc, red_d);
  cmax = block_max(cmax, red_f);
  nmax_in = block_max(nmax_in, red_f);

  // all-to-all inside the cluster through distributed shared memory
  if (tid < kMixCluster) {
    double* dst_d = cluster.map_shared_rank(&xch1_d[rank][0], tid);
    float* dst_f = cluster.map_shared_rank(&xch1_f[rank][0], tid);
#pragma unroll
    for (int k = 0; k < 5; ++k) dst_d[k] = acc[k];
    dst_f[0] = cmax;
    dst_f[1] = nmax_in;
  }
  cluster.sync();
  double s_cc = 0.0, s_nn = 0.0, s_c = 0.0, s_n = 0.0, s_cn = 0.0;
  cmax = 0.f;
  nmax_in = 0.f;
#pragma unroll
  for (int r = 0; r < kMixCluster; ++r) {
    s_cc += xch1_d[r][0];
    s_nn += xch1_d[r][1];
    s_c += xch1_d[r][2];
    s_n += xch1_d[r][3];
    s_cn += xch1_d[r][4];
    cmax = fmaxf(cmax, xch1_f[r][0]);
    nmax_in = fmaxf(nmax_in, xch1_f[r][1]);
  }

  // ---- add_noise_to_speech decisions (augment.py:7-51), evaluated identically by every thread ----
  const double Ld = static_cast<double>(L);
  const float Ps = static_cast<float>(s_cc / Ld);  // torch.mean(speech ** 2)
  const float Pn = static_cast<float>(s_nn / Ld);
  int idx = p.snr_idx[donor];
  idx = idx < 0 ? 0 : (idx >= p.n_snr ? p.n_snr - 1 : idx);
  if (rank == 0 && tid == 0 && p.snr_used != nullptr) p.snr_used[row] = idx;
  const float scale = __fsqrt_rn(__fdiv_rn(Ps, __fmul_rn(Pn, p.snr_lin[idx])));  // augment.py:40, fp32
  int st = 0;
  if (isnan(Ps)) st = 1;                         // isnan(speech).any()
  else if (isnan(Pn)) st = 2;                    // isnan(noise).any()
  else if (Ps < 1e-10f) st = 3;
  else if (Pn < 1e-10f) st = 4;
  else if (isinf(scale) || isnan(scale)) st = 5;
  else if (scale > 1e6f) st = 6;
  else if (isinf(nmax_in)) st = 7;               // inf * 0 -> NaN in noise * scale (augment.py:56)

  float nmax = 0.f;
  if (p.peak_norm && st == 0) {
    // ---- pass 2 (L2): peak of the mixed signal, NaN flags of augment.py:56-64 ---------------------
    unsigned flags = 0;
    for (int v0 = v_begin + tid; v0 < v_end; v0 += kMixThreads) {
      float c[4], n[4], y[4];
      load4<kVec>(c_row, n_row, L, Ln, v0, c, n);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float sn = __fmul_rn(n[j], scale);
        y[j] = __fadd_rn(c[j], sn);
        flags |= (isnan(sn) ? 1u : 0u) | (isnan(y[j]) ? 2u : 0u);
      }
      nmax = absmax4(nmax, y);
    }
    nmax = block_max(nmax, red_f);
    flags = __syncthreads_or(static_cast<int>(flags));
    if (tid < kMixCluster) {
      *cluster.map_shared_rank(&xch2_f[rank], tid) = nmax;
      *cluster.map_shared_rank(&xch2_u[rank], tid) = flags;
    }
    cluster.sync();
    nmax = 0.f;
    flags = 0;
#pragma unroll
    for (int r = 0; r < kMixCluster; ++r) {
      nmax = fmaxf(nmax, xch2_f[r]);
      flags |= xch2_u[r];
    }
    if (flags & 1u) st = 7;
    else if (flags & 2u) st = 8;
    else if (cmax < 1e-8f) st = 9;               // noisy_speech_dataset.py:95
    else if (nmax < 1e-8f) st = 10;              // :99
    else if (isinf(cmax)) st = 11;               // inf / inf -> NaN after the peak division (:107)
    else if (isinf(nmax)) st = 12;               // (:111)
  }

  // ---- statistics of the normalised views, from the pass-1 sums -----------------------------------
  // clean view : x = c / dc             noisy view : x = (c + s n) / dn
  const double s = static_cast<double>(scale);
  float a_c = 0.f, b_c = 0.f, a_n = 0.f, b_n = 0.f, inv_dc = 1.f, inv_dn = 1.f;
  bool mixed = (st == 0);
  if (p.peak_norm) {
    if (st == 0) {
      const double dc = static_cast<double>(__fadd_rn(cmax, 1e-8f));
      const double dn = static_cast<double>(__fadd_rn(nmax, 1e-8f));
      const double mc = s_c / Ld / dc;
      const double vc = s_cc / Ld / (dc * dc) - mc * mc;
      const double mn = (s_c + s * s_n) / Ld / dn;
      const double vn = (s_cc + 2.0 * s * s_cn + s * s * s_nn) / Ld / (dn * dn) - mn * mn;
      const float vcf = static_cast<float>(vc), vnf = static_cast<float>(vn);
      const float mcf = static_cast<float>(mc), mnf = static_cast<float>(mn);
      if (!isfinite(mcf) || !isfinite(vcf)) st = 13;
      else if (!isfinite(mnf) || !isfinite(vnf)) st = 14;
      inv_dc = static_cast<float>(1.0 / dc);
      inv_dn = static_cast<float>(1.0 / dn);
      a_c = mcf;
      b_c = 1.0f / __fsqrt_rn(__fadd_rn(vcf, 1e-7f));  // np.sqrt(x.var() + 1e-7) in float32
      a_n = mnf;
      b_n = 1.0f / __fsqrt_rn(__fadd_rn(vnf, 1e-7f));
    }
  } else {
    // emotion mode (emotion_dataset.py:177-203): no peak normalisation; a failed mix keeps the clean wave
    const double sw = mixed ? (s_c + s * s_n) : s_c;
    const double sww = mixed ? (s_cc + 2.0 * s * s_cn + s * s * s_nn) : s_cc;
    const double mn = sw / Ld;
    const double vn = sww / Ld - mn * mn;
    a_n = static_cast<float>(mn);
    b_n = 1.0f / __fsqrt_rn(__fadd_rn(static_cast<float>(vn), 1e-7f));
    if (p.raw) {  // (y - 0) * 1 is exact: the output is the mixed signal itself
      a_n = 0.f;
      b_n = 1.f;
    }
  }
  if (rank == 0 && tid == 0) p.status[row] = st;

  // ---- output pass (inputs from L2) ----------------------------------------------------------------
  float* co_row = p.clean_out ? p.clean_out + static_cast<size_t>(row) * L : nullptr;
  float* no_row = p.noisy_out + static_cast<size_t>(row) * L;
  if (p.peak_norm && st != 0) {  // the reference would re-draw this item: hand back zeros + status
    const float z[4] = {0.f, 0.f, 0.f, 0.f};
    for (int v0 = v_begin + tid; v0 < v_end; v0 += kMixThreads) {
      store4<kVec>(co_row, L, v0, z);
      store4<kVec>(no_row, L, v0, z);
    }
    continue;  // next attempt, if this launch makes one
  }
  for (int v0 = v_begin + tid; v0 < v_end; v0 += 2 * kMixThreads) {
    float c[2][4], n[2][4];
    const int v1 = v0 + kMixThreads;
    const bool has1 = v1 < v_end;
    load4<kVec>(c_row, n_row, L, Ln, v0, c[0], n[0]);
    if (has1) load4<kVec>(c_row, n_row, L, Ln, v1, c[1], n[1]);
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      if (u == 1 && !has1) break;
      float oc[4], on[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float y = mixed ? __fadd_rn(c[u][j], __fmul_rn(n[u][j], scale)) : c[u][j];  // augment.py:54,60
        if (p.peak_norm) {
          oc[j] = (__fmul_rn(c[u][j], inv_dc) - a_c) * b_c;
          on[j] = (__fmul_rn(y, inv_dn) - a_n) * b_n;
        } else {
          on[j] = (y - a_n) * b_n;
        }
      }
      if (p.peak_norm) store4<kVec>(co_row, L, u ? v1 : v0, oc);
      store4<kVec>(no_row, L, u ? v1 : v0, on);
    }
  }
  if (st == 0) break;  // (emotion mode: a rejected mix wrote the clean waveform and may be retried)
  }  // attempts
}

// ---------------------------------------------------------------------------------------------------------
// Helpers of the on-chip-resident kernel below (packed fp32x2 arithmetic, NaN-propagating peaks, streaming stores).
// ---------------------------------------------------------------------------------------------------------
constexpr int kSmemMaxCluster = 8;  // CTAs per row (portable cluster limit)

// packed fp32x2 helpers (FFMA2 / FMUL2 / FADD2): same IEEE rounding per lane as the scalar instructions
typedef unsigned long long f2;
__device__ __forceinline__ f2 f2_make(float lo, float hi) {
  f2 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void f2_split(f2 v, float& lo, float& hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ f2 f2_fma(f2 a, f2 b, f2 c) {
  f2 d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ f2 f2_mul(f2 a, f2 b) {
  f2 d;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ f2 f2_add(f2 a, f2 b) {
  f2 d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ float f2_hsum(f2 v) {
  float a, b;
  f2_split(v, a, b);
  return a + b;
}
// NaN-propagating |.| max: a NaN anywhere in the row surfaces in the reduced peak
__device__ __forceinline__ float absmax_nan(float m, float x) {
  float r;
  asm("max.NaN.f32 %0, %1, %2;" : "=f"(r) : "f"(m), "f"(fabsf(x)));
  return r;
}

// Scalars every thread needs after a phase; computed once by thread 0 (fp64 divisions are ~40 instructions each)
struct MixScalars {
  float scale, inv_dc, inv_dn, a_c, b_c, a_n, b_n;
  int st1;        // verdict of add_noise_to_speech (after pass 1)
  int st;         // final verdict
  int clean_ok;   // the clean view passes its own peak / z-norm checks (it may be written before the mixed peak is known)
  int clean_nan;  // mean / variance of the peak-normalised clean view is not finite (status 13 if nothing earlier applies)
};


__device__ __forceinline__ void st_stream_cs_f4(float4* p, const float4& v) {
  asm volatile("st.global.cs.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

// ---------------------------------------------------------------------------------------------------------
// On-chip-resident variant (default when the row fits): the whole row lives ON CHIP between the passes -- the first
// 4096 float4 of each CTA's segment in REGISTERS (4 + 4 float4 per thread, 128 KB per SM: the register file is the
// largest memory of the SM), the rest in shared memory (bulk async copies, <= 96 KB per array).  One 1024-thread CTA
// per SM holds up to 320 KB of input, so a 4 s row needs a cluster of only two CTAs (74 rows in flight instead of the
// 37 of the shared-memory-only design), every byte of the row is in flight the moment the CTA starts, and passes 2
// and 3 touch neither L2 nor HBM for reads: the memory system sees exactly the algorithmic 16 B per sample.  (The
// streaming variant pushes 4x that through the L2 slices, whose ~6300 B/clk cap is then as tight as HBM itself.)
// ---------------------------------------------------------------------------------------------------------
// Two shapes of the same kernel: 512 threads x 2 CTAs per SM (the default: the two co-resident CTAs belong to different
// rows, so one CTA's reduction / exchange / scalar bubbles -- about 5 us per row, measured -- are covered by the other's
// HBM phases) and 1024 threads x 1 CTA per SM (twice the capacity per CTA, for rows too long for eight small CTAs).
constexpr int kResRegVec = 4;  // float4 per thread per array held in registers
constexpr int kResChunks = 4;
template <int kThreads>
struct ResCfg {
  static constexpr int kWarps = kThreads / 32;
  static constexpr int kCtasPerSm = 1024 / kThreads;        // 1024 threads x 64 registers fill the register file
  static constexpr int kRegCap = kResRegVec * kThreads;     // float4 per array in registers (64 B per thread)
  static constexpr int kSmemCap = kThreads == 1024 ? 6144 : 3456;  // float4 per array in shared memory
  static constexpr int kCap = kRegCap + kSmemCap;
};

// kRetry: the launch redoes rejected rows only and makes up to p.n_attempts attempts per row inside the launch (see MixParams);
// every barrier then completes one phase per attempt, and a cluster barrier between attempts keeps a fast CTA's next
// messages out of buffers a slower CTA of the row is still reading.  The first-attempt instantiation carries none of this.
template <int kResThreads, bool kRetry>
__global__ void __launch_bounds__(kResThreads, 1024 / kResThreads)
mix_normalize_resident_kernel(const MixParams p, int seg_vec, int smem_pitch) {
  constexpr int kResWarps = ResCfg<kResThreads>::kWarps;
  constexpr int kResRegCap = ResCfg<kResThreads>::kRegCap;
  cg::cluster_group cluster = cg::this_cluster();
  const int rank = static_cast<int>(cluster.block_rank());
  const int cs = static_cast<int>(cluster.num_blocks());
  const int row = blockIdx.x / cs;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

  extern __shared__ __align__(128) unsigned char mix_smem[];
  __shared__ __align__(8) unsigned long long bars[kResChunks];
  __shared__ __align__(8) unsigned long long xbar[2];  // cluster exchange 1 (sums, input peaks) and 2 (mixed peak)
  __shared__ double red_d[kResWarps][5];
  __shared__ float red_f[kResWarps][2];
  __shared__ __align__(16) double xch1_d[kSmemMaxCluster][6];  // 5 sums, then {max|c|, max|n|} as two floats
  __shared__ float xch2_f[kSmemMaxCluster];
  __shared__ MixScalars sc;

  const int L = p.L;
  const int nvec = L >> 2;  // L % 4 == 0 on this path
  const int v_begin = min(nvec, rank * seg_vec);
  const int v_end = min(nvec, v_begin + seg_vec);
  const int n_my = v_end - v_begin;
  const int n_reg = min(n_my, kResRegCap);
  const int n_sm = n_my - n_reg;  // local vectors [n_reg, n_my) live in shared memory
  const int chunk_vec = (n_sm + kResChunks - 1) / kResChunks;
  float4* s_c = reinterpret_cast<float4*>(mix_smem);
  float4* s_n = s_c + smem_pitch;
  if (kRetry && p.status[row] == 0) return;  // whole cluster (same row): nothing to redo
  const float4* g_c = reinterpret_cast<const float4*>(p.clean + static_cast<size_t>(row) * L) + v_begin;

  // The cluster exchanges are st.async messages that complete_tx on the RECEIVER's mbarrier: no cluster barrier and no
  // gpu-scope fence on the critical path (cg::cluster.sync() costs MEMBAR.GPU + ERRBAR twice per row).  Each CTA
  // sends its partials to every CTA of the cluster (itself included), in both exchanges, before it waits for the
  // second one -- so once a CTA has received everything nothing can still be addressed to it and it may exit.
  const unsigned xbar1 = ptx::smem_u32(&xbar[0]), xbar2 = ptx::smem_u32(&xbar[1]);
  const int n_att = kRetry ? p.n_attempts : 1;
  unsigned x2_phase = 0;  // phases xbar2 has completed: the second exchange only happens for rows that pass add_noise_to_speech,
  bool x2_armed = false;  // so an attempt that skipped it leaves the barrier armed for the next one
  for (int att = 0; att < n_att; ++att) {
  const unsigned ph = static_cast<unsigned>(att) & 1u;  // phase parity of the per-attempt barriers
  const int donor = kRetry ? (row + p.noise_shift + att) % p.B : row;  // whose noise crop and SNR draw this attempt uses
  const float4* g_n = reinterpret_cast<const float4*>(p.noise + static_cast<size_t>(donor) * p.Ln) + v_begin;
  if (tid == 0) {
    if (att == 0) {
      for (int c = 0; c < kResChunks; ++c) ptx::mbar_init(ptx::smem_u32(&bars[c]), 1);
      ptx::mbar_init(xbar1, 1);
      ptx::mbar_init(xbar2, 1);
      ptx::fence_mbar_init();
    }
    ptx::mbar_arrive_expect_tx(xbar1, static_cast<unsigned>(cs) * 48u);
    if (!x2_armed) ptx::mbar_arrive_expect_tx(xbar2, static_cast<unsigned>(cs) * 4u);
    for (int c = 0; c < kResChunks; ++c) {
      const int c0 = c * chunk_vec;
      const int len = min(chunk_vec, n_sm - c0);
      if (len <= 0) break;
      const unsigned bar = static_cast<unsigned>(__cvta_generic_to_shared(&bars[c]));
      const unsigned bytes = static_cast<unsigned>(len) * 16u;
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(2u * bytes) : "memory");
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                       static_cast<unsigned>(__cvta_generic_to_shared(s_c + c0))),
                   "l"(g_c + n_reg + c0), "r"(bytes), "r"(bar)
                   : "memory");
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                       static_cast<unsigned>(__cvta_generic_to_shared(s_n + c0))),
                   "l"(g_n + n_reg + c0), "r"(bytes), "r"(bar)
                   : "memory");
    }
  }

  // ---- register part: 8 x 128-bit loads per thread, all in flight at once ---------------------------------------
  float4 rc[kResRegVec], rn[kResRegVec];
#pragma unroll
  for (int u = 0; u < kResRegVec; ++u) {
    const int v = u * kResThreads + tid;
    if (v < n_reg) {
      rc[u] = ld_stream_f4(g_c + v);
      rn[u] = ld_stream_f4(g_n + v);
    } else {
      rc[u] = rn[u] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }
  x2_armed = true;
  __syncthreads();  // mbarrier init visible to the waiting threads
  // peers: my barriers exist (and, from the second attempt on: I have read everything they sent me for the previous one)
  if (cs > 1) {
    if (kRetry) asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    else asm volatile("barrier.cluster.arrive.relaxed.aligned;" ::: "memory");
  }

  // ---- pass 1: packed fp32 partial sums per thread (<= 40 samples per lane-half), fp64 across ---------------------
  const f2 zero2 = f2_make(0.f, 0.f);
  f2 a_cc = zero2, a_nn = zero2, a_c1 = zero2, a_n1 = zero2, a_cn = zero2;
  float cmax = 0.f, nmax_in = 0.f;
  auto accum = [&](const float4& cv, const float4& nv) {
    const f2 c01 = f2_make(cv.x, cv.y), c23 = f2_make(cv.z, cv.w);
    const f2 n01 = f2_make(nv.x, nv.y), n23 = f2_make(nv.z, nv.w);
    a_cc = f2_fma(c01, c01, a_cc); a_cc = f2_fma(c23, c23, a_cc);
    a_nn = f2_fma(n01, n01, a_nn); a_nn = f2_fma(n23, n23, a_nn);
    a_cn = f2_fma(c01, n01, a_cn); a_cn = f2_fma(c23, n23, a_cn);
    a_c1 = f2_add(a_c1, c01); a_c1 = f2_add(a_c1, c23);
    a_n1 = f2_add(a_n1, n01); a_n1 = f2_add(a_n1, n23);
    cmax = fmaxf(fmaxf(cmax, fmaxf(fabsf(cv.x), fabsf(cv.y))), fmaxf(fabsf(cv.z), fabsf(cv.w)));
    nmax_in = fmaxf(fmaxf(nmax_in, fmaxf(fabsf(nv.x), fabsf(nv.y))), fmaxf(fabsf(nv.z), fabsf(nv.w)));
  };
#pragma unroll
  for (int u = 0; u < kResRegVec; ++u) accum(rc[u], rn[u]);
  if (n_sm > 0) {
    for (int c = 0; c < kResChunks; ++c) {
      const int c0 = c * chunk_vec;
      const int c_end = min(n_sm, c0 + chunk_vec);
      if (c0 >= c_end) break;
      const unsigned bar = static_cast<unsigned>(__cvta_generic_to_shared(&bars[c]));
      asm volatile(
          "{\n\t.reg .pred p;\n\tRES_WAIT_LOOP:\n\t"
          "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
          "@p bra RES_WAIT_DONE;\n\tbra RES_WAIT_LOOP;\n\tRES_WAIT_DONE:\n\t}\n" ::"r"(bar), "r"(ph)
          : "memory");
      for (int v = c0 + tid; v < c_end; v += kResThreads) accum(s_c[v], s_n[v]);
    }
  }

  // CTA-wide then cluster-wide combine.  Warp shuffles (xor butterfly: every lane ends with the same bits), one warp
  // over the 32 warp partials, then every CTA receives every CTA's partial in rank order: deterministic, and all
  // CTAs of the row derive bit-identical scalars.
  {
    double acc[5] = {static_cast<double>(f2_hsum(a_cc)), static_cast<double>(f2_hsum(a_nn)),
                     static_cast<double>(f2_hsum(a_c1)), static_cast<double>(f2_hsum(a_n1)),
                     static_cast<double>(f2_hsum(a_cn))};
#pragma unroll
    for (int k = 0; k < 5; ++k) acc[k] = warp_sum(acc[k]);
    cmax = warp_max(cmax);
    nmax_in = warp_max(nmax_in);
    if (lane == 0) {
#pragma unroll
      for (int k = 0; k < 5; ++k) red_d[warp][k] = acc[k];
      red_f[warp][0] = cmax;
      red_f[warp][1] = nmax_in;
    }
    __syncthreads();
    if (cs > 1) asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
    if (warp == 0) {
#pragma unroll
      for (int k = 0; k < 5; ++k) acc[k] = warp_sum(lane < kResWarps ? red_d[lane][k] : 0.0);
      const float m0 = warp_max(lane < kResWarps ? red_f[lane][0] : 0.f);
      const float m1 = warp_max(lane < kResWarps ? red_f[lane][1] : 0.f);
      if (lane < cs) {
        const unsigned dst = ptx::mapa(ptx::smem_u32(&xch1_d[rank][0]), lane);
        const unsigned dbar = ptx::mapa(xbar1, lane);
#pragma unroll
        for (int k = 0; k < 5; ++k)
          asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.b64 [%0], %1, [%2];" ::"r"(
                           dst + 8u * k),
                       "l"(__double_as_longlong(acc[k])), "r"(dbar)
                       : "memory");
        ptx::st_async_f2(dst + 40u, m0, m1, dbar);
      }
    }
  }
  // Two threads wait for the first exchange: thread 0 derives the mix scale (what pass 2 waits for), lane 0 of warp 1 the
  // scalars of the CLEAN view -- they depend on nothing but the clean sums and peak, so the clean view can leave for HBM while
  // pass 2, the second exchange and the noisy view's scalars (~2.5 us of latency chain per row) are still under way.
  constexpr int kCleanTid = 32;
  if (tid == 0 || (p.peak_norm && tid == kCleanTid)) ptx::mbar_wait(xbar1, ph);

  double s_cc = 0.0, s_nn = 0.0, s_c1 = 0.0, s_n1 = 0.0, s_cn = 0.0;
  const double inv_L = 1.0 / static_cast<double>(L);
  if (p.peak_norm && tid == kCleanTid) {
    double q_cc = 0.0, q_c1 = 0.0;
    float pk_c = 0.f;
    for (int r = 0; r < cs; ++r) {
      q_cc += xch1_d[r][0];
      q_c1 += xch1_d[r][2];
      pk_c = fmaxf(pk_c, reinterpret_cast<const float2*>(&xch1_d[r][5])->x);
    }
    const double rdc = 1.0 / static_cast<double>(__fadd_rn(pk_c, 1e-8f));
    const double mc = q_c1 * inv_L * rdc;
    const double vc = q_cc * inv_L * rdc * rdc - mc * mc;
    const float vcf = static_cast<float>(vc), mcf = static_cast<float>(mc);
    const bool bad = !isfinite(mcf) || !isfinite(vcf);
    sc.inv_dc = static_cast<float>(rdc);
    sc.a_c = mcf;
    sc.b_c = 1.0f / __fsqrt_rn(__fadd_rn(vcf, 1e-7f));  // np.sqrt(x.var() + 1e-7) in float32
    sc.clean_nan = bad ? 1 : 0;
    sc.clean_ok = (!bad && !(pk_c < 1e-8f) && !isinf(pk_c)) ? 1 : 0;
  }
  if (tid == 0) {
    cmax = 0.f;
    nmax_in = 0.f;
    for (int r = 0; r < cs; ++r) {
      s_cc += xch1_d[r][0];
      s_nn += xch1_d[r][1];
      s_c1 += xch1_d[r][2];
      s_n1 += xch1_d[r][3];
      s_cn += xch1_d[r][4];
      const float2 pk = *reinterpret_cast<const float2*>(&xch1_d[r][5]);
      cmax = fmaxf(cmax, pk.x);
      nmax_in = fmaxf(nmax_in, pk.y);
    }
    const float Ps = static_cast<float>(s_cc * inv_L);
    const float Pn = static_cast<float>(s_nn * inv_L);
    int idx = p.snr_idx[donor];
    idx = idx < 0 ? 0 : (idx >= p.n_snr ? p.n_snr - 1 : idx);
    if (rank == 0 && p.snr_used != nullptr) p.snr_used[row] = idx;
    const float scale = __fsqrt_rn(__fdiv_rn(Ps, __fmul_rn(Pn, p.snr_lin[idx])));
    int st = 0;
    if (isnan(Ps)) st = 1;
    else if (isnan(Pn)) st = 2;
    else if (Ps < 1e-10f) st = 3;
    else if (Pn < 1e-10f) st = 4;
    else if (isinf(scale) || isnan(scale)) st = 5;
    else if (scale > 1e6f) st = 6;
    else if (isinf(nmax_in)) st = 7;
    sc.scale = scale;
    sc.st1 = st;
  }
  __syncthreads();
  const float scale = sc.scale;
  const int st1 = sc.st1;
  const bool mixed = st1 == 0;
  const f2 s2 = f2_make(scale, scale);

  // ---- clean view out (BYOL mode), ahead of pass 2.  If a later check rejects the row after all, pass 3 overwrites it with
  // zeros: same thread, same addresses, program order.
  float4* co = p.clean_out ? reinterpret_cast<float4*>(p.clean_out + static_cast<size_t>(row) * L) + v_begin : nullptr;
  if (p.peak_norm && mixed && sc.clean_ok) {
    const f2 nac2 = f2_make(-sc.a_c, -sc.a_c), bc2 = f2_make(sc.b_c, sc.b_c), idc2 = f2_make(sc.inv_dc, sc.inv_dc);
    auto emit_clean = [&](int v, const float4& cv) {
      float4 oc;
      f2_split(f2_mul(f2_add(f2_mul(f2_make(cv.x, cv.y), idc2), nac2), bc2), oc.x, oc.y);  // (c/dc - mean) / std, 3 roundings
      f2_split(f2_mul(f2_add(f2_mul(f2_make(cv.z, cv.w), idc2), nac2), bc2), oc.z, oc.w);
      st_stream_cs_f4(co + v, oc);
    };
#pragma unroll
    for (int u = 0; u < kResRegVec; ++u) {
      const int v = u * kResThreads + tid;
      if (v < n_reg) emit_clean(v, rc[u]);
    }
    for (int v = tid; v < n_sm; v += kResThreads) emit_clean(n_reg + v, s_c[v]);
  }

  // ---- pass 2 (on chip): y = c + s*n replaces n; peak of the mixed signal (BYOL mode) ----------------------------
  float nmax = 0.f;
  if (p.peak_norm && mixed) {
    auto mix_in_place = [&](const float4& cv, float4& nv) {
      f2_split(f2_add(f2_make(cv.x, cv.y), f2_mul(f2_make(nv.x, nv.y), s2)), nv.x, nv.y);  // augment.py:54,60
      f2_split(f2_add(f2_make(cv.z, cv.w), f2_mul(f2_make(nv.z, nv.w), s2)), nv.z, nv.w);
      nmax = absmax_nan(absmax_nan(absmax_nan(absmax_nan(nmax, nv.x), nv.y), nv.z), nv.w);
    };
#pragma unroll
    for (int u = 0; u < kResRegVec; ++u) mix_in_place(rc[u], rn[u]);
    for (int v = tid; v < n_sm; v += kResThreads) {
      const float4 cv = s_c[v];
      float4 nv = s_n[v];
      mix_in_place(cv, nv);
      s_n[v] = nv;  // re-read in pass 3 by the same thread
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float other = __shfl_xor_sync(0xffffffffu, nmax, o);
      asm("max.NaN.f32 %0, %0, %1;" : "+f"(nmax) : "f"(other));
    }
    if (lane == 0) red_f[warp][0] = nmax;
    __syncthreads();
    if (warp == 0) {
      float m = lane < kResWarps ? red_f[lane][0] : 0.f;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const float other = __shfl_xor_sync(0xffffffffu, m, o);
        asm("max.NaN.f32 %0, %0, %1;" : "+f"(m) : "f"(other));
      }
      if (lane < cs)
        asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.b32 [%0], %1, [%2];" ::"r"(
                         ptx::mapa(ptx::smem_u32(&xch2_f[rank]), lane)),
                     "r"(__float_as_uint(m)), "r"(ptx::mapa(xbar2, lane))
                     : "memory");
    }
    if (tid == 0) ptx::mbar_wait(xbar2, x2_phase & 1u);
    ++x2_phase;
    x2_armed = false;
  }

  if (tid == 0) {
    int st = st1;
    const double s = static_cast<double>(scale);
    float a_n = 0.f, b_n = 0.f, inv_dn = 1.f;
    if (p.peak_norm) {
      if (st == 0) {
        nmax = xch2_f[0];
        for (int r = 1; r < cs; ++r) asm("max.NaN.f32 %0, %0, %1;" : "+f"(nmax) : "f"(xch2_f[r]));
        if (isnan(nmax)) st = 8;
        else if (cmax < 1e-8f) st = 9;
        else if (nmax < 1e-8f) st = 10;
        else if (isinf(cmax)) st = 11;
        else if (isinf(nmax)) st = 12;
      }
      if (st == 0) {  // (the clean view's scalars are in `sc` already: kCleanTid, after the first exchange)
        const double rdn = 1.0 / static_cast<double>(__fadd_rn(nmax, 1e-8f));
        const double mn = (s_c1 + s * s_n1) * inv_L * rdn;
        const double vn = (s_cc + 2.0 * s * s_cn + s * s * s_nn) * inv_L * rdn * rdn - mn * mn;
        const float vnf = static_cast<float>(vn), mnf = static_cast<float>(mn);
        if (sc.clean_nan) st = 13;
        else if (!isfinite(mnf) || !isfinite(vnf)) st = 14;
        inv_dn = static_cast<float>(rdn);
        a_n = mnf;
        b_n = 1.0f / __fsqrt_rn(__fadd_rn(vnf, 1e-7f));
      }
    } else {
      const double sw = mixed ? (s_c1 + s * s_n1) : s_c1;
      const double sww = mixed ? (s_cc + 2.0 * s * s_cn + s * s * s_nn) : s_cc;
      const double mn = sw * inv_L;
      const double vn = sww * inv_L - mn * mn;
      a_n = static_cast<float>(mn);
      b_n = 1.0f / __fsqrt_rn(__fadd_rn(static_cast<float>(vn), 1e-7f));
      if (p.raw) {
        a_n = 0.f;
        b_n = 1.f;
      }
    }
    sc.inv_dn = inv_dn;
    sc.a_n = a_n; sc.b_n = b_n;
    sc.st = st;
    if (rank == 0) p.status[row] = st;
  }
  __syncthreads();
  const int st = sc.st;

  // ---- pass 3 (on chip -> HBM): the noisy view (the clean one left before pass 2) -----------------------------------------
  float4* no = reinterpret_cast<float4*>(p.noisy_out + static_cast<size_t>(row) * L) + v_begin;
  if (p.peak_norm && st != 0) {
    const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int v = tid; v < n_my; v += kResThreads) {
      st_stream_cs_f4(co + v, z);
      st_stream_cs_f4(no + v, z);
    }
    if (kRetry) __syncthreads();
    continue;  // next attempt, if this launch makes one
  }
  const f2 nan2 = f2_make(-sc.a_n, -sc.a_n), bn2 = f2_make(sc.b_n, sc.b_n), idn2 = f2_make(sc.inv_dn, sc.inv_dn);
  auto emit = [&](int v, const float4& cv, const float4& yv) {
    f2 y01 = f2_make(yv.x, yv.y), y23 = f2_make(yv.z, yv.w);
    float4 on;
    if (p.peak_norm) {  // the noise slot already holds the mixed signal: (y/dn - mean) / std, 3 roundings
      f2_split(f2_mul(f2_add(f2_mul(y01, idn2), nan2), bn2), on.x, on.y);
      f2_split(f2_mul(f2_add(f2_mul(y23, idn2), nan2), bn2), on.z, on.w);
    } else {
      const f2 c01 = f2_make(cv.x, cv.y), c23 = f2_make(cv.z, cv.w);
      if (mixed) {
        y01 = f2_add(c01, f2_mul(y01, s2));
        y23 = f2_add(c23, f2_mul(y23, s2));
      } else {
        y01 = c01;
        y23 = c23;
      }
      f2_split(f2_mul(f2_add(y01, nan2), bn2), on.x, on.y);
      f2_split(f2_mul(f2_add(y23, nan2), bn2), on.z, on.w);
    }
    st_stream_cs_f4(no + v, on);
  };
#pragma unroll
  for (int u = 0; u < kResRegVec; ++u) {
    const int v = u * kResThreads + tid;
    if (v < n_reg) emit(v, rc[u], rn[u]);
  }
  for (int v = tid; v < n_sm; v += kResThreads) emit(n_reg + v, s_c[v], s_n[v]);
  if (st == 0) break;  // (emotion mode: a rejected mix wrote the clean waveform and may be retried)
  if (kRetry) __syncthreads();  // every thread is done with the on-chip row before the next attempt's copies overwrite it
  }  // attempts
}

// The tail of the batch: (1) rows that failed every attempt (status != 0 after the retries) take over the outputs of the
// nearest following good row of the batch -- the device-side form of the reference's "move on to the next item"
// (ref:src/data/noisy_speech_dataset.py:60-66) -- so that no all-zero waveform reaches BatchNorm statistics or the loss
// mean; (2) the batch's "snr" labels (the value of the table entry every row was FINALLY mixed at, :140-144) are gathered;
// (3) the number of rows that were rejected for good is written for the host's log.  One CTA per row; on a healthy batch
// the launch moves no waveform data.  `status` is not modified (it keeps reporting why the row failed); a batch without
// any good row is left as it is.
__global__ void __launch_bounds__(256) mix_finish_kernel(float* __restrict__ clean_out, float* __restrict__ noisy_out,
                                                         const int32_t* __restrict__ status,
                                                         int32_t* __restrict__ snr_used,
                                                         const int64_t* __restrict__ label_table,
                                                         int64_t* __restrict__ labels_out,
                                                         int32_t* __restrict__ n_rejected, int substitute, int B, int L) {
  const int row = blockIdx.x;
  __shared__ int s_src;
  __shared__ int s_cnt;
  const bool bad = status[row] != 0;
  if (row == 0 && n_rejected != nullptr) {  // status is read-only here: CTA 0 counts the whole batch
    if (threadIdx.x == 0) s_cnt = 0;
    __syncthreads();
    int c = 0;
    for (int r = threadIdx.x; r < B; r += blockDim.x) c += status[r] != 0 ? 1 : 0;
    c = __reduce_add_sync(0xffffffffu, c);
    if ((threadIdx.x & 31) == 0 && c) atomicAdd(&s_cnt, c);
    __syncthreads();
    if (threadIdx.x == 0) *n_rejected = s_cnt;
  }
  if (threadIdx.x == 0) {
    int src = -1;
    if (bad && substitute) {
      for (int d = 1; d < B; ++d) {
        const int r = (row + d) % B;
        if (status[r] == 0) { src = r; break; }
      }
    }
    s_src = src;
    // good rows' snr_used entries are never written here, so reading the donor's is race-free
    const int used = snr_used != nullptr ? snr_used[src >= 0 ? src : row] : 0;
    if (src >= 0 && snr_used != nullptr) snr_used[row] = used;
    if (labels_out != nullptr) labels_out[row] = label_table[used];
  }
  if (!bad || !substitute) return;
  __syncthreads();
  const int src = s_src;
  if (src < 0) return;
  for (int k = 0; k < 2; ++k) {
    float* base = k == 0 ? clean_out : noisy_out;
    if (base == nullptr) continue;
    const float* from = base + static_cast<size_t>(src) * L;
    float* to = base + static_cast<size_t>(row) * L;
    if (((reinterpret_cast<uintptr_t>(from) | reinterpret_cast<uintptr_t>(to)) & 15u) == 0 && L % 4 == 0) {
      for (int v = threadIdx.x; v < L / 4; v += blockDim.x)
        reinterpret_cast<float4*>(to)[v] = reinterpret_cast<const float4*>(from)[v];
    } else {
      for (int i = threadIdx.x; i < L; i += blockDim.x) to[i] = from[i];
    }
  }
}

int g_mix_carveout = -1;  // shared-memory carveout (percent) of the resident kernel; -1: just what the CTAs need
int g_mix_cluster = 0;    // 0: automatic; 1..8 force the CTAs per row of the resident kernel (tuning)
int g_mix_variant = 4;    // 4: on-chip resident (registers + shared memory) whenever the row fits 8 CTAs (default);
                          // 5: the same, forcing the 1024-thread / one-CTA-per-SM shape;
                          // 0: always the generic re-read-from-L2 kernel (the fallback for unaligned rows, tiled noise
                          //    and rows beyond 8 x 40960 samples)

const char* const kMixStatusNames[] = {
    "ok", "speech_nan", "noise_nan", "speech_power_too_small", "noise_power_too_small", "scale_invalid",
    "scale_too_large", "scaled_noise_nan", "noisy_nan", "clean_peak_too_small", "noisy_peak_too_small",
    "clean_norm_nan", "noisy_norm_nan", "clean_znorm_nan", "noisy_znorm_nan"};

}  // namespace
}  // namespace nrse

extern "C" {

int nrse_mix_set_cluster(int ctas_per_row) {
  if (ctas_per_row < 0 || ctas_per_row > nrse::kSmemMaxCluster) return NRSE_ERR_INVALID_ARG;
  nrse::g_mix_cluster = ctas_per_row;
  return NRSE_OK;
}

int nrse_mix_set_carveout(int percent) {
  if (percent < -1 || percent > 100) return NRSE_ERR_INVALID_ARG;
  nrse::g_mix_carveout = percent;
  return NRSE_OK;
}

int nrse_mix_set_variant(int variant) {
  if (variant != 0 && variant != 4 && variant != 5) return NRSE_ERR_INVALID_ARG;
  nrse::g_mix_variant = variant;
  return NRSE_OK;
}

const char* nrse_mix_status_name(int code) {
  return (code >= 0 && code <= 14) ? nrse::kMixStatusNames[code] : "unknown";
}

static int mix_normalize_impl(const float* clean, const float* noise, const int32_t* snr_idx,
                              const double* snr_db_table_host, int n_snr, float* clean_out, float* noisy_out,
                              int32_t* status, int32_t* snr_used, int B, int L, int L_noise, int peak_norm, int retry,
                              int noise_shift, int n_attempts, nrse_stream_t stream) {
  using namespace nrse;
  if (!clean || !noise || !snr_idx || !snr_db_table_host || !noisy_out || !status) return NRSE_ERR_INVALID_ARG;
  if (B < 0 || L <= 0 || L_noise <= 0 || n_snr <= 0 || n_snr > kMaxSnr) return NRSE_ERR_INVALID_ARG;
  if (peak_norm < 0 || peak_norm > 2) return NRSE_ERR_INVALID_ARG;
  if (peak_norm == 1 && !clean_out) return NRSE_ERR_INVALID_ARG;
  if (retry && (noise_shift < 0 || n_attempts < 1)) return NRSE_ERR_INVALID_ARG;
  if (B == 0) return NRSE_OK;

  MixParams p;
  p.clean = clean; p.noise = noise; p.snr_idx = snr_idx;
  p.clean_out = peak_norm == 1 ? clean_out : nullptr;
  p.noisy_out = noisy_out; p.status = status; p.snr_used = snr_used;
  p.B = B; p.L = L; p.Ln = L_noise; p.peak_norm = peak_norm == 1 ? 1 : 0; p.n_snr = n_snr;
  p.raw = peak_norm == 2 ? 1 : 0;
  p.retry = retry ? 1 : 0;
  p.noise_shift = retry ? noise_shift % B : 0;
  p.n_attempts = retry ? n_attempts : 1;
  for (int i = 0; i < kMaxSnr; ++i)
    p.snr_lin[i] = i < n_snr ? static_cast<float>(std::pow(10.0, snr_db_table_host[i] / 10.0)) : 1.0f;

  auto aligned16 = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15u) == 0; };
  const bool vec = (L % 4 == 0) && (L_noise % 4 == 0) && (L_noise >= L) && aligned16(clean) && aligned16(noise) &&
                   aligned16(noisy_out) && (peak_norm != 1 || aligned16(clean_out));

  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cudaLaunchConfig_t cfg = {};
  cfg.stream = as_stream(stream);
  cfg.attrs = attr;
  cfg.numAttrs = 1;

  if (vec && g_mix_variant >= 4) {
    // CTAs per row (<= 8, any size; sweep in scripts/bench_mix_sweep.py, B200): the 512-thread shape with segments of
    // <= 4096 float4 (half in registers, half in shared memory: 2 x 64 KB of shared memory per SM leaves ~100 KB of
    // L1 for the register-bound loads in flight) is best at every length it can serve; segments that fill the whole
    // shared-memory capacity (5334 float4: 61 % instead of 73 % at 4 s) or larger clusters than needed lose 5-15 %.
    // Longer rows: 512 threads up to the full capacity, then the 1024-thread shape; rows beyond 8 x 40960 samples
    // (20 s) go to the generic kernel.
    const int nvec = L / 4;
    auto min_cluster = [&](int cap) {
      const int cs = ceil_div(nvec, cap);
      return cs < 1 ? 1 : cs;
    };
    int threads = 512, cs = min_cluster(2 * ResCfg<512>::kRegCap);
    if (cs > kSmemMaxCluster) cs = min_cluster(ResCfg<512>::kCap);
    if (cs > kSmemMaxCluster || g_mix_variant == 5) {
      threads = 1024;
      cs = min_cluster(ResCfg<1024>::kCap);
    }
    if (cs <= kSmemMaxCluster) {
      const int cap = threads == 512 ? ResCfg<512>::kCap : ResCfg<1024>::kCap;
      const int reg_cap = threads == 512 ? ResCfg<512>::kRegCap : ResCfg<1024>::kRegCap;
      if (g_mix_cluster > 0 && ceil_div(nvec, g_mix_cluster) <= cap) cs = g_mix_cluster;
      const int seg_vec = ceil_div(nvec, cs);
      const int smem_pitch = seg_vec > reg_cap ? seg_vec - reg_cap : 0;
      const size_t dyn = static_cast<size_t>(smem_pitch) * 32;
      static bool res_attr_set = false;  // benign race: idempotent attributes
      if (!res_attr_set) {
        NRSE_CUDA_TRY(cudaFuncSetAttribute(mix_normalize_resident_kernel<512, false>,
                                           cudaFuncAttributeMaxDynamicSharedMemorySize, ResCfg<512>::kSmemCap * 32));
        NRSE_CUDA_TRY(cudaFuncSetAttribute(mix_normalize_resident_kernel<1024, false>,
                                           cudaFuncAttributeMaxDynamicSharedMemorySize, ResCfg<1024>::kSmemCap * 32));
        NRSE_CUDA_TRY(cudaFuncSetAttribute(mix_normalize_resident_kernel<512, true>,
                                           cudaFuncAttributeMaxDynamicSharedMemorySize, ResCfg<512>::kSmemCap * 32));
        NRSE_CUDA_TRY(cudaFuncSetAttribute(mix_normalize_resident_kernel<1024, true>,
                                           cudaFuncAttributeMaxDynamicSharedMemorySize, ResCfg<1024>::kSmemCap * 32));
        res_attr_set = true;
      }
      // Shared-memory carveout: just enough for the resident CTAs.  Whatever is left stays L1, and L1 is where the
      // 128 KB of register-bound loads per SM land while in flight: with the carveout at 100 % the same kernel is
      // 15-25 % slower (measured), the loads being throttled by the 28 KB of L1 that remain.
      {
        const int ctas_per_sm = threads == 512 ? ResCfg<512>::kCtasPerSm : ResCfg<1024>::kCtasPerSm;
        const size_t need = ctas_per_sm * (dyn + 2048 + 1024);
        int pct = g_mix_carveout >= 0 ? g_mix_carveout : static_cast<int>((need * 100 + 228 * 1024 - 1) / (228 * 1024));
        pct = pct > 100 ? 100 : pct;
        static int last_pct[2] = {-1, -1};  // benign race: idempotent attribute
        int& last = last_pct[threads == 512 ? 0 : 1];
        if (pct != last) {
          if (threads == 512) {
            NRSE_CUDA_TRY(cudaFuncSetAttribute(mix_normalize_resident_kernel<512, false>,
                                               cudaFuncAttributePreferredSharedMemoryCarveout, pct));
            NRSE_CUDA_TRY(cudaFuncSetAttribute(mix_normalize_resident_kernel<512, true>,
                                               cudaFuncAttributePreferredSharedMemoryCarveout, pct));
          } else {
            NRSE_CUDA_TRY(cudaFuncSetAttribute(mix_normalize_resident_kernel<1024, false>,
                                               cudaFuncAttributePreferredSharedMemoryCarveout, pct));
            NRSE_CUDA_TRY(cudaFuncSetAttribute(mix_normalize_resident_kernel<1024, true>,
                                               cudaFuncAttributePreferredSharedMemoryCarveout, pct));
          }
          last = pct;
        }
      }
      cfg.gridDim = dim3(static_cast<unsigned>(B) * cs);
      cfg.blockDim = dim3(threads);
      cfg.dynamicSmemBytes = dyn;
      attr[0].val.clusterDim.x = cs;
      if (threads == 512) {
        if (retry) NRSE_CUDA_TRY(cudaLaunchKernelEx(&cfg, mix_normalize_resident_kernel<512, true>, p, seg_vec, smem_pitch));
        else NRSE_CUDA_TRY(cudaLaunchKernelEx(&cfg, mix_normalize_resident_kernel<512, false>, p, seg_vec, smem_pitch));
      } else {
        if (retry) NRSE_CUDA_TRY(cudaLaunchKernelEx(&cfg, mix_normalize_resident_kernel<1024, true>, p, seg_vec, smem_pitch));
        else NRSE_CUDA_TRY(cudaLaunchKernelEx(&cfg, mix_normalize_resident_kernel<1024, false>, p, seg_vec, smem_pitch));
      }
      return NRSE_OK;
    }
  }
  // generic kernel: a fixed cluster of 4 CTAs per row, pass 1 from HBM, passes 2-3 re-read from L2; any alignment, any
  // length, tiled noise
  cfg.gridDim = dim3(static_cast<unsigned>(B) * kMixCluster);
  cfg.blockDim = dim3(kMixThreads);
  cfg.dynamicSmemBytes = 0;
  attr[0].val.clusterDim.x = kMixCluster;
  if (vec) {
    if (retry) NRSE_CUDA_TRY(cudaLaunchKernelEx(&cfg, mix_normalize_kernel<true, true>, p));
    else NRSE_CUDA_TRY(cudaLaunchKernelEx(&cfg, mix_normalize_kernel<true, false>, p));
  } else {
    if (retry) NRSE_CUDA_TRY(cudaLaunchKernelEx(&cfg, mix_normalize_kernel<false, true>, p));
    else NRSE_CUDA_TRY(cudaLaunchKernelEx(&cfg, mix_normalize_kernel<false, false>, p));
  }
  return NRSE_OK;
}

static int mix_finish_impl(float* clean_out, float* noisy_out, const int32_t* status, int32_t* snr_used,
                           const int64_t* label_table, int64_t* labels_out, int32_t* n_rejected, int substitute, int B, int L,
                           nrse_stream_t stream) {
  using namespace nrse;
  if (!noisy_out || !status || B < 0 || L <= 0) return NRSE_ERR_INVALID_ARG;
  if ((labels_out != nullptr) && (label_table == nullptr || snr_used == nullptr)) return NRSE_ERR_INVALID_ARG;
  if (B == 0) return NRSE_OK;
  if (B < 2) substitute = 0;  // nothing to substitute from
  if (!substitute && !labels_out && !n_rejected) return NRSE_OK;
  mix_finish_kernel<<<B, 256, 0, as_stream(stream)>>>(clean_out, noisy_out, status, snr_used, label_table, labels_out,
                                                     n_rejected, substitute, B, L);
  NRSE_CHECK_LAUNCH();
  return NRSE_OK;
}

int nrse_mix_substitute_rows_f32(float* clean_out, float* noisy_out, const int32_t* status, int32_t* snr_idx_used, int B,
                                 int L, nrse_stream_t stream) {
  return mix_finish_impl(clean_out, noisy_out, status, snr_idx_used, nullptr, nullptr, nullptr, 1, B, L, stream);
}

int nrse_mix_normalize_f32(const float* clean, const float* noise, const int32_t* snr_idx,
                           const double* snr_db_table_host, int n_snr, float* clean_out, float* noisy_out,
                           int32_t* status, int B, int L, int L_noise, int peak_norm, nrse_stream_t stream) {
  return mix_normalize_impl(clean, noise, snr_idx, snr_db_table_host, n_snr, clean_out, noisy_out, status, nullptr, B, L,
                            L_noise, peak_norm, 0, 0, 1, stream);
}

int nrse_mix_normalize_retry_f32(const float* clean, const float* noise, const int32_t* snr_idx,
                                 const double* snr_db_table_host, int n_snr, float* clean_out, float* noisy_out,
                                 int32_t* status, int32_t* snr_idx_used, int B, int L, int L_noise, int peak_norm,
                                 int noise_row_shift, nrse_stream_t stream) {
  return mix_normalize_impl(clean, noise, snr_idx, snr_db_table_host, n_snr, clean_out, noisy_out, status, snr_idx_used,
                            B, L, L_noise, peak_norm, 1, noise_row_shift, 1, stream);
}

int nrse_mix_batch_f32(const float* clean, const float* noise, const int32_t* snr_idx, const double* snr_db_table_host,
                       int n_snr, float* clean_out, float* noisy_out, int32_t* status, int32_t* snr_idx_used,
                       const int64_t* snr_label_table, int64_t* snr_labels_out, int32_t* n_rejected, int B, int L,
                       int L_noise, int peak_norm, int max_attempts, int substitute_bad_rows, nrse_stream_t stream) {
  if (!snr_idx_used || max_attempts < 1) return NRSE_ERR_INVALID_ARG;
  int rc = mix_normalize_impl(clean, noise, snr_idx, snr_db_table_host, n_snr, clean_out, noisy_out, status, snr_idx_used, B,
                              L, L_noise, peak_norm, 0, 0, 1, stream);
  if (rc != NRSE_OK) return rc;
  // attempts 2..max_attempts: ONE launch, rows still rejected loop inside it.  BYOL mode only: the emotion path never
  // retries, a failed mix keeps the clean waveform (ref:src/data/emotion_dataset.py:190-194)
  if (max_attempts > 1 && B > 1 && peak_norm == 1) {
    rc = mix_normalize_impl(clean, noise, snr_idx, snr_db_table_host, n_snr, clean_out, noisy_out, status, snr_idx_used, B, L,
                            L_noise, peak_norm, 1, 1, max_attempts - 1, stream);
    if (rc != NRSE_OK) return rc;
  }
  // (emotion mode: a row whose mix was rejected carries the normalised CLEAN waveform -- a valid item, never substituted)
  return mix_finish_impl(peak_norm == 1 ? clean_out : nullptr, noisy_out, status, snr_idx_used, snr_label_table,
                         snr_labels_out, n_rejected, peak_norm == 1 ? substitute_bad_rows : 0, B, L, stream);
}

}  // extern "C"
