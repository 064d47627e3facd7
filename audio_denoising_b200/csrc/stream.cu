// stream.cu -- one hop of the streaming loop of app3.py:178-226 for S independent sessions.
//
// v1 composition: the per-hop chain is issued as the same kernels the batched path uses, on a
// 3-frame problem per session (quirks Q2-Q4 of SURVEY.md Appendix C kept), bracketed by two small
// kernels for the chunk conditioning (peak normalise + Hann pre-window, app3.py:179-188) and the
// output overlap-add ring (app3.py:219-224).
#include "kernels.cuh"

namespace b2d {

int launch_inverse_mel_tc(const b2d_plan* p, const float* mel_bt, size_t nframes, float* mag_tf, int terms, cudaStream_t st);  // invmel_tc.cu

// workspace: peak[S] | x[S,N] | logmel[S,3,M] | pred | mel | mag[S,3,Fp] | y[S,N] | model ws | GL ws
struct StreamWs {
  float *peak, *x, *logmel, *pred, *mel, *mag, *y;
  unsigned char* model_ws; size_t model_bytes;
  unsigned char* gl_ws; size_t gl_bytes;
  size_t total;
};
static StreamWs stream_layout(const b2d_plan* p, const b2d_model* m, int S, void* base) {
  StreamWs w;
  unsigned char* q = static_cast<unsigned char*>(base);
  size_t o = 0;
  auto take = [&](size_t bytes) { size_t at = o; o += align_up(bytes, 256); return q ? q + at : nullptr; };
  const int T = 1 + p->n_fft / p->hop;
  const size_t nm = (size_t)S * T * p->n_mels * sizeof(float);
  w.peak = reinterpret_cast<float*>(take((size_t)S * sizeof(float)));
  w.x = reinterpret_cast<float*>(take((size_t)S * p->n_fft * sizeof(float)));
  w.logmel = reinterpret_cast<float*>(take(nm));
  w.pred = reinterpret_cast<float*>(take(nm));
  w.mel = reinterpret_cast<float*>(take(nm));
  w.mag = reinterpret_cast<float*>(take((size_t)S * T * p->Fp * sizeof(float)));
  w.y = reinterpret_cast<float*>(take((size_t)S * p->hop * (T - 1) * sizeof(float)));
  w.model_bytes = model_workspace_bytes(m, S, T);
  w.model_ws = take(w.model_bytes);
  w.gl_bytes = gl_workspace_bytes(p, S, T, false);
  w.gl_ws = take(w.gl_bytes);
  w.total = o;
  return w;
}
size_t stream_step_ws(const b2d_plan* p, const b2d_model* m, int S) { return stream_layout(p, m, S, nullptr).total; }

// peak normalise + Hann pre-window (app3.py:179-188)
__global__ void __launch_bounds__(256) stream_pre_kernel(const float* __restrict__ chunk, const float* __restrict__ win, int N,
                                                         float* __restrict__ x, float* __restrict__ peak) {
  const int s = blockIdx.x;
  pdl_wait();
  pdl_trigger();
  const float* c = chunk + (size_t)s * N;  // may be pinned HOST memory (zero-copy streaming graph): read each sample once
  constexpr int KEEP = 8;                  // N <= 2048: the chunk stays in registers between the two passes
  float keep[KEEP];
  float m = 0.f;
#pragma unroll
  for (int q = 0; q < KEEP; ++q) {
    const int i = threadIdx.x + q * 256;
    keep[q] = i < N ? c[i] : 0.f;
    m = fmaxf(m, fabsf(keep[q]));
  }
  for (int i = threadIdx.x + KEEP * 256; i < N; i += blockDim.x) m = fmaxf(m, fabsf(c[i]));
  for (int o = 16; o; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  __shared__ float red[8];
  __shared__ float pk;
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x == 0) {
    float v = red[0];
    for (int i = 1; i < 8; ++i) v = fmaxf(v, red[i]);
    pk = (v > 1e-6f) ? v : 1.0f;
    peak[s] = pk;
  }
  __syncthreads();
  const float p = pk;
#pragma unroll
  for (int q = 0; q < KEEP; ++q) {
    const int i = threadIdx.x + q * 256;
    if (i < N) x[(size_t)s * N + i] = (keep[q] / p) * win[i];
  }
  for (int i = threadIdx.x + KEEP * 256; i < N; i += blockDim.x) x[(size_t)s * N + i] = (c[i] / p) * win[i];
}

// emit ola[:hop]; shift; ola += y (app3.py:219-224).  N == 2*hop.
__global__ void __launch_bounds__(256) stream_ola_kernel(const float* __restrict__ y, float* __restrict__ ola,
                                                         float* __restrict__ out, int hop) {
  const int s = blockIdx.x;
  const int N = 2 * hop;
  pdl_wait();  // y comes from the Griffin-Lim kernel before
  pdl_trigger();
  for (int i = threadIdx.x; i < hop; i += blockDim.x) {
    const float o0 = ola[(size_t)s * N + i], o1 = ola[(size_t)s * N + hop + i];
    out[(size_t)s * hop + i] = o0;
    ola[(size_t)s * N + i] = o1 + y[(size_t)s * N + i];
    ola[(size_t)s * N + hop + i] = y[(size_t)s * N + hop + i];
  }
}

int stream_step_impl(const b2d_plan* p, const b2d_model* m, const float* chunk, int S, float* hx, float* ola,
                     const float2* init_angles, unsigned long long seed, const unsigned long long* d_seed, int n_iter, float momentum,
                     int conv_mode, float* out, void* ws, size_t ws_bytes, cudaStream_t st) {
  B2D_REQUIRE(S >= 1 && S <= 65535, B2D_ERR_BAD_ARG, "sessions must be in [1, 65535] (got %d)", S);
  B2D_REQUIRE(p->hop * 2 == p->n_fft, B2D_ERR_UNSUPPORTED, "streaming requires hop == n_fft/2");
  B2D_REQUIRE(m->n_mels == p->n_mels, B2D_ERR_BAD_ARG, "plan n_mels (%d) != model n_mels (%d)", p->n_mels, m->n_mels);
  B2D_REQUIRE(aligned16(ws), B2D_ERR_ALIGN, "workspace must be 16-byte aligned");
  const StreamWs w = stream_layout(p, m, S, ws);
  B2D_REQUIRE(ws_bytes >= w.total, B2D_ERR_WORKSPACE, "stream workspace too small (%zu < %zu)", ws_bytes, w.total);
  const int N = p->n_fft, T = 3;
  int rc;
  if (stft_frames_per_block(p) >= T) {
    // peak normalise + Hann pre-window fused into the STFT kernel (one CTA per session reads the chunk once, possibly from pinned host memory)
    if ((rc = launch_stft(p, chunk, nullptr, S, N, w.logmel, nullptr, nullptr, st, w.peak))) return rc;
  } else {
    B2D_CUDA(launch_pdl(stream_pre_kernel, dim3(S), dim3(256), 0, st, chunk, p->d_win, N, w.x, w.peak));
    B2D_LAUNCH_CHECK("stream_pre_kernel");
    if ((rc = launch_stft(p, w.x, nullptr, S, N, w.logmel, nullptr, nullptr, st))) return rc;
  }
  if ((rc = model_forward(m, w.logmel, hx, w.pred, w.mel, 1, 0.f, S, T, conv_mode, w.model_ws, w.model_bytes, st))) return rc;
  if (p->d_tw8 != nullptr && !(p->flags & B2D_PLAN_FP32_INVMEL)) {
    if ((rc = launch_inverse_mel_tc(p, w.mel, (size_t)S * T, w.mag, 3, st))) return rc;
  } else if ((rc = launch_inverse_mel(p, w.mel, S, T, w.mag, false, st))) {
    return rc;
  }
  if (gl_fuses_ola(p, S, T))  // emit + shift + add of the overlap-add ring behind the last iteration of the single-launch hop kernel
    return gl_run(p, w.mag, init_angles, seed, S, T, n_iter, momentum, w.peak, w.y, w.gl_ws, w.gl_bytes, st, d_seed, ola, out);
  if ((rc = gl_run(p, w.mag, init_angles, seed, S, T, n_iter, momentum, w.peak, w.y, w.gl_ws, w.gl_bytes, st, d_seed))) return rc;
  B2D_CUDA(launch_pdl(stream_ola_kernel, dim3(S), dim3(256), 0, st, w.y, ola, out, p->hop));
  B2D_LAUNCH_CHECK("stream_ola_kernel");
  return B2D_OK;
}

}  // namespace b2d
