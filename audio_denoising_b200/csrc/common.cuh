// common.cuh -- shared host/device helpers for libb200denoise (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <atomic>
#include <string>

#include "../../include/b200denoise.h"

namespace b2d {

// ---- error plumbing -------------------------------------------------------------------------
std::string& last_error();                  // thread-local
int fail(int code, const char* fmt, ...);   // records message, returns code
extern std::atomic<unsigned long long> g_launches;

#define B2D_REQUIRE(cond, code, ...)                   \
  do {                                                 \
    if (!(cond)) return ::b2d::fail((code), __VA_ARGS__); \
  } while (0)

#define B2D_CUDA(call)                                                                          \
  do {                                                                                          \
    cudaError_t e__ = (call);                                                                   \
    if (e__ != cudaSuccess)                                                                     \
      return ::b2d::fail(B2D_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), \
                         __FILE__, __LINE__);                                                   \
  } while (0)

// after every kernel launch
#define B2D_LAUNCH_CHECK(name)                                                                    \
  do {                                                                                            \
    ::b2d::g_launches.fetch_add(1, std::memory_order_relaxed);                                    \
    cudaError_t e__ = cudaGetLastError();                                                         \
    if (e__ != cudaSuccess)                                                                       \
      return ::b2d::fail(B2D_ERR_CUDA, "launch of %s failed: %s", name, cudaGetErrorString(e__)); \
  } while (0)

// Opt a kernel in to `bytes` of dynamic shared memory.  The attribute belongs to the function (per device), not to the
// launch, so it is set when a launch site first needs more than it has asked for before -- not on every launch.
#define B2D_SMEM_OPT_IN(bytes, ...)                                                                              \
  do {                                                                                                           \
    static std::atomic<int> have__[16];                                                                          \
    int dev__ = 0;                                                                                               \
    cudaGetDevice(&dev__);                                                                                       \
    const int want__ = (int)(bytes);                                                                             \
    if (have__[dev__ & 15].load(std::memory_order_acquire) < want__) {                                           \
      B2D_CUDA(cudaFuncSetAttribute(__VA_ARGS__, cudaFuncAttributeMaxDynamicSharedMemorySize, want__));          \
      have__[dev__ & 15].store(want__, std::memory_order_release);                                               \
    }                                                                                                            \
  } while (0)

// ---- programmatic dependent launch (PDL) ------------------------------------------------------
// Back-to-back launches of the same persistent kernel (32 Griffin-Lim iterations) pay launch latency + the table / twiddle
// prologue + the drain of the slowest SM between every pair.  A kernel launched with launch_pdl() may start its CTAs as soon as
// the CTAs of the kernel before it in the stream have exited (or called pdl_trigger()); everything it does before pdl_wait()
// must touch only memory no earlier kernel of the chain writes (plan tables); pdl_wait() returns once the previous kernel has
// completed and its writes are visible.  Launched without the attribute both calls are no-ops.
#ifdef __CUDACC__
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
template <typename... KArgs, typename... Args>
static inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}
#endif

static inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }
static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// ---- plan / model (host structs; device tables owned by them) -------------------------------
constexpr int kMaxPass = 8;

struct FftDesc {  // passed by value to kernels
  int M;          // complex length = n_fft / 2
  int npass;
  int radix[kMaxPass];
};

}  // namespace b2d

struct b2d_plan {
  int n_fft, hop, n_mels, F, M, Fp;
  unsigned flags;     // B2D_PLAN_* (include/b200denoise.h), fixed at creation
  b2d::FftDesc fft;
  int device;
  int num_sms;
  float2* d_tw;       // [M]      W_M^k   = exp(-2 pi i k / M)
  float2* d_tw512;    // [512]    W_512^k (register FFT of the n_fft = 512 / 1024 fast paths)
  float2* d_rtw;      // [M]      W_N^k   = exp(-2 pi i k / N)   (real-FFT split twiddles)
  float* d_win;       // [N]      periodic Hann
  float* d_winn;      // [N]      Hann / N (synthesis window with the irfft 1/N folded in)
  float* d_inv_env;   // [hop]    1 / (w^2[i] + w^2[i+hop])      (valid when hop == N/2)
  int* d_mel_lo;      // [n_mels] first frequency bin with non-zero weight
  int* d_mel_cnt;     // [n_mels] number of consecutive bins
  int* d_mel_off;     // [n_mels] offset into d_mel_w
  float* d_mel_w;     // compact column weights
  int mel_nnz;
  // the same columns cut into segments of <= 8 bins for the warp-per-frame STFT kernel (balanced lanes):
  float* d_seg_w;     // [2][mel_seg_pad][4] taps 0-3 / 4-7 of each segment, zero padded
  int* d_seg_lo;      // [mel_seg_pad] first bin of the segment
  int* d_seg_first;   // [n_mels + 1] first segment of each mel column
  int mel_seg_pad;    // number of segments rounded up to a multiple of 32
  float* d_pinv;      // [Fp, n_mels] rows F..Fp-1 zero
  float2* d_tw8;      // TF32 big/small weight images of pinv for the tcgen05 inverse-mel GEMM (invmel_tc.cu), may be null
};

struct b2d_model {
  b2d_model_config cfg;
  int n_mels;
  int device;
  float* d_blob;      // all packed tensors, one allocation
  float* h_blob;      // host copy of the same (weights passed by value as __grid_constant__ kernel parameters)
  size_t blob_floats;
  // offsets (in floats) into d_blob, see model.cu
  int enc_w[6], enc_pb[6];
  int rec_w, rec_pb;
  int dec_w[6], dec_pb[6];
  float* d_mma;       // weight fragment images for the warp-level MMA decoder (unet_mma.cu)
  float* d_utc;       // weight images of the fused tcgen05 encoder (unet_tc.cu)
};
