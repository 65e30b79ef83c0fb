// Library-wide plumbing of libnrse_b200.so: version, error strings, device check.
#include "common.cuh"

namespace nrse {
static thread_local cudaError_t g_last_cuda_error = cudaSuccess;
void set_last_cuda_error(cudaError_t e) { g_last_cuda_error = e; }
}  // namespace nrse

extern "C" {

int nrse_version(void) { return 200; /* 0.2.0 */ }

int nrse_experiments_build(void) {
#ifdef NRSE_EXPERIMENTS
  return 1;
#else
  return 0;
#endif
}

const char* nrse_strerror(int status) {
  switch (status) {
    case NRSE_OK: return "ok";
    case NRSE_ERR_INVALID_ARG: return "invalid argument";
    case NRSE_ERR_UNSUPPORTED: return "unsupported shape or mode";
    case NRSE_ERR_CUDA: return "CUDA call failed (see nrse_last_cuda_error)";
    case NRSE_ERR_WORKSPACE: return "workspace too small";
    case NRSE_ERR_NO_DEVICE: return "current device is not an sm_100 (B200) GPU";
    default: return "unknown nrse status";
  }
}

int nrse_last_cuda_error(void) { return static_cast<int>(nrse::g_last_cuda_error); }

int nrse_check_device(void) {
  int dev = 0;
  NRSE_CUDA_TRY(cudaGetDevice(&dev));
  int major = 0;
  NRSE_CUDA_TRY(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
  return major == 10 ? NRSE_OK : NRSE_ERR_NO_DEVICE;
}

}  // extern "C"
