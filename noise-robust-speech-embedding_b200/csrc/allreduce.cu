// Gradient all-reduce over NVSwitch multicast memory (NVLink SHARP): the data-parallel step's one collective
// (the mean of the ranks' gradients, what DistributedDataParallel's all-reduce computes for
// ref:train_byol.py:66-70 once the batch is sharded over the GPUs of a box) as ONE small kernel per rank.
//
// The gradient arena of every rank is a symmetric allocation mapped into one multicast object (host side:
// torch.distributed._symmetric_memory, plumbing only).  Rank r owns the r-th slice of the range: for every 16 bytes of it
//     v = multimem.ld_reduce.add [mc + i]     -- the SWITCH reads the 16 bytes from all W arenas and returns their sum
//     multimem.st [mc + i], v / W             -- the switch writes the mean into all W arenas
// so a GPU's NVLink ports carry each byte once in and once out (2 x bytes x (W-1)/W, both directions busy at once) and
// its SMs do no reduction arithmetic.  The kernel is register-light (128 threads, no shared memory).  Cross-rank ordering
// (everybody's gradients written before / everybody's slice stored after) is the caller's: a symmetric-memory barrier on
// the same stream before and after the launch.
// Measured (profiles/r2_allreduce_overlap.md): correct to 5e-8 against the exact mean; at 2 GPUs -- where the multicast
// scheme also loops a rank's own slice through the switch -- 3.2 ms for 1.3 GB against NCCL's 2.6 ms, and next to the
// conv-frontend backward it overlaps no better than NCCL does (static persistent tile schedules lose a whole wave to any
// SM they have to share), so GradArena keeps NCCL as its default and this kernel as the opt-in path.
#include "common.cuh"

namespace nrse {
namespace {

__device__ __forceinline__ float4 multimem_ld_reduce_f4(const float* mc) {
  float4 v;
  asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "l"(mc)
               : "memory");
  return v;
}
__device__ __forceinline__ void multimem_st_f4(float* mc, const float4& v) {
  asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(mc), "f"(v.x), "f"(v.y), "f"(v.z),
               "f"(v.w)
               : "memory");
}

constexpr int kArThreads = 128;  // x 8 loads of 16 bytes in flight per thread: 16 KB per CTA, ~6 K registers
constexpr int kArUnroll = 8;

__global__ void __launch_bounds__(kArThreads) multimem_allreduce_kernel(float* __restrict__ mc, long long begin4,
                                                                        long long end4, float scale) {
  const long long stride = static_cast<long long>(gridDim.x) * kArThreads;
  long long i = begin4 + static_cast<long long>(blockIdx.x) * kArThreads + threadIdx.x;
  for (; i + (kArUnroll - 1) * stride < end4; i += kArUnroll * stride) {
    float4 v[kArUnroll];
#pragma unroll
    for (int u = 0; u < kArUnroll; ++u) v[u] = multimem_ld_reduce_f4(mc + 4 * (i + u * stride));
#pragma unroll
    for (int u = 0; u < kArUnroll; ++u) {
      v[u].x *= scale; v[u].y *= scale; v[u].z *= scale; v[u].w *= scale;
      multimem_st_f4(mc + 4 * (i + u * stride), v[u]);
    }
  }
  for (; i < end4; i += stride) {
    float4 v = multimem_ld_reduce_f4(mc + 4 * i);
    v.x *= scale; v.y *= scale; v.z *= scale; v.w *= scale;
    multimem_st_f4(mc + 4 * i, v);
  }
}

}  // namespace
}  // namespace nrse

extern "C" {

int nrse_multimem_allreduce_mean_f32(void* multicast_ptr, int64_t elem_offset, int64_t numel, int rank, int world,
                                     int max_ctas, nrse_stream_t stream) {
  using namespace nrse;
  if (!multicast_ptr || numel < 0 || elem_offset < 0 || world < 1 || rank < 0 || rank >= world) return NRSE_ERR_INVALID_ARG;
  if ((elem_offset & 3) || (numel & 3) || (reinterpret_cast<uintptr_t>(multicast_ptr) & 15u)) return NRSE_ERR_INVALID_ARG;
  if (numel == 0) return NRSE_OK;
  const long long n4 = numel / 4;
  const long long per = ceil_div(n4, static_cast<long long>(world));
  const long long b4 = elem_offset / 4 + per * rank;
  long long e4 = b4 + per;
  const long long end_all = elem_offset / 4 + n4;
  if (e4 > end_all) e4 = end_all;
  if (b4 >= e4) return NRSE_OK;  // this rank's slice is empty (tiny ranges)
  long long ctas = ceil_div(e4 - b4, static_cast<long long>(kArThreads * kArUnroll));
  const long long cap = max_ctas > 0 ? max_ctas : kNumSMs;
  if (ctas > cap) ctas = cap;
  multimem_allreduce_kernel<<<static_cast<unsigned>(ctas), kArThreads, 0, as_stream(stream)>>>(
      reinterpret_cast<float*>(multicast_ptr), b4, e4, 1.0f / static_cast<float>(world));
  NRSE_CHECK_LAUNCH();
  return NRSE_OK;
}

}  // extern "C"
