// ingest.cu -- audio ingest adjacent to the hot path:
//   int16 PCM -> float (app3.py:168-172, utils.py:109-116), mono mix-down (app.py:184-186),
//   float -> int16 with clipping (app3.py:244-245), and the 44.1 k <-> 48 k polyphase sinc resampler
//   (utils.py:48-49: torchaudio.transforms.Resample; TA:functional/functional.py:1305-1432).
#include "kernels.cuh"

struct b2d_resampler {
  int orig, neu, K, width;  // rates reduced by their gcd, taps per phase, left padding
  float* d_kernel;          // [neu][K]
  int device;
};

namespace b2d {

__global__ void __launch_bounds__(256) pcm16_to_float_kernel(const short* __restrict__ pcm, size_t n, int ch, int channel,
                                                             float* __restrict__ out) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    float v;
    if (channel >= 0) {
      v = (float)pcm[i * ch + channel] / 32767.0f;
    } else {  // mean over channels of the scaled samples
      float acc = 0.f;
      for (int c = 0; c < ch; ++c) acc += (float)pcm[i * ch + c] / 32767.0f;
      v = acc / (float)ch;
    }
    out[i] = v;
  }
}

__global__ void __launch_bounds__(256) float_to_pcm16_kernel(const float* __restrict__ in, size_t n, short* __restrict__ out) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const float c = fminf(fmaxf(in[i], -1.0f), 1.0f) * 32767.0f;  // np.clip(x, -1, 1) * 32767
    out[i] = (short)(int)c;                                        // .astype(np.int16): truncation toward zero
  }
}

// out[b, m * neu + p] = sum_k kern[p, k] * xpad[b, m * orig + k],  xpad = zero-pad(x, (width, width + orig))
__global__ void __launch_bounds__(256) resample_kernel(const float* __restrict__ x, int L, const float* __restrict__ kern, int orig,
                                                       int neu, int K, int width, float* __restrict__ out, int Lout) {
  const int b = blockIdx.y;
  const float* xb = x + (size_t)b * L;
  for (int o = blockIdx.x * blockDim.x + threadIdx.x; o < Lout; o += gridDim.x * blockDim.x) {
    const int m = o / neu, p = o - m * neu;
    const float* kp = kern + (size_t)p * K;
    const int base = m * orig - width;
    float acc = 0.f;
    const int k0 = base < 0 ? -base : 0;
    const int k1 = (base + K > L) ? L - base : K;
    for (int k = k0; k < k1; ++k) acc = fmaf(kp[k], xb[base + k], acc);
    out[(size_t)b * Lout + o] = acc;
  }
}

}  // namespace b2d

using namespace b2d;

extern "C" {

int b2d_pcm16_to_float(const short* pcm, size_t n, int channels, int channel, float* out, void* stream) {
  B2D_REQUIRE(channels >= 1 && channel < channels, B2D_ERR_BAD_ARG, "bad channel selection %d of %d", channel, channels);
  if (n == 0) return B2D_OK;
  B2D_REQUIRE(pcm && out, B2D_ERR_BAD_ARG, "NULL pointer");
  const unsigned blocks = (unsigned)((n + 255) / 256 < 148 * 8 ? (n + 255) / 256 : 148 * 8);
  pcm16_to_float_kernel<<<blocks, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(pcm, n, channels, channel, out);
  B2D_LAUNCH_CHECK("pcm16_to_float_kernel");
  return B2D_OK;
}

int b2d_float_to_pcm16(const float* in, size_t n, short* out, void* stream) {
  if (n == 0) return B2D_OK;
  B2D_REQUIRE(in && out, B2D_ERR_BAD_ARG, "NULL pointer");
  const unsigned blocks = (unsigned)((n + 255) / 256 < 148 * 8 ? (n + 255) / 256 : 148 * 8);
  float_to_pcm16_kernel<<<blocks, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(in, n, out);
  B2D_LAUNCH_CHECK("float_to_pcm16_kernel");
  return B2D_OK;
}

int b2d_resampler_create(int orig_reduced, int new_reduced, int taps, int width, const float* h_kernel, b2d_resampler** out) {
  B2D_REQUIRE(out && h_kernel, B2D_ERR_BAD_ARG, "NULL pointer");
  *out = nullptr;
  B2D_REQUIRE(orig_reduced >= 1 && new_reduced >= 1 && taps >= 1 && width >= 0, B2D_ERR_BAD_ARG, "bad resampler geometry");
  b2d_resampler* r = new b2d_resampler();
  r->orig = orig_reduced; r->neu = new_reduced; r->K = taps; r->width = width; r->d_kernel = nullptr;
  B2D_CUDA(cudaGetDevice(&r->device));
  const size_t bytes = sizeof(float) * (size_t)new_reduced * taps;
  cudaError_t e = cudaMalloc(&r->d_kernel, bytes);
  if (e == cudaSuccess) e = cudaMemcpy(r->d_kernel, h_kernel, bytes, cudaMemcpyHostToDevice);
  if (e != cudaSuccess) {
    cudaFree(r->d_kernel);
    delete r;
    return fail(B2D_ERR_CUDA, "resampler table upload failed: %s", cudaGetErrorString(e));
  }
  *out = r;
  return B2D_OK;
}
void b2d_resampler_destroy(b2d_resampler* r) {
  if (!r) return;
  cudaFree(r->d_kernel);
  delete r;
}
int b2d_resample_length(const b2d_resampler* r, int L) {
  if (!r || L < 0) return B2D_ERR_BAD_ARG;
  return (int)(((long long)r->neu * L + r->orig - 1) / r->orig);  // ceil(new * L / orig)
}
int b2d_resample(const b2d_resampler* r, const float* in, int B, int L, float* out, void* stream) {
  B2D_REQUIRE(r && in && out, B2D_ERR_BAD_ARG, "NULL pointer");
  B2D_REQUIRE(B >= 1 && B <= 65535 && L >= 1, B2D_ERR_BAD_ARG, "bad shape [%d, %d]", B, L);
  const int Lout = b2d_resample_length(r, L);
  dim3 grid((Lout + 255) / 256 < 2048 ? (Lout + 255) / 256 : 2048, B);
  resample_kernel<<<grid, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(in, L, r->d_kernel, r->orig, r->neu, r->K, r->width, out, Lout);
  B2D_LAUNCH_CHECK("resample_kernel");
  return B2D_OK;
}

}  // extern "C"
