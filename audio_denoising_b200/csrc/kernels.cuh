// kernels.cuh -- internal launcher declarations shared by the translation units of libb200denoise.
#pragma once

#include "common.cuh"

namespace b2d {

// dsp.cu
int launch_peak(const float* wave, int B, int L, float* peak, float* partial, int chunks, cudaStream_t st);
// int16 PCM [B, L] -> float wave [B, L] (x / 32767) + peak[B] (same rule as launch_peak) in one pass
int launch_pcm16_ingest_peak(const short* pcm, int B, int L, float* wave, float* peak, float* partial, int chunks, cudaStream_t st);
// pre_peak != nullptr (streaming hop, L == n_fft): `wave` is the raw chunk; (chunk / peak) * window is formed in-kernel, peak -> pre_peak[B]
int launch_stft(const b2d_plan* p, const float* wave, const float* inv_scale, int B, int L, float* logmel_bt,
                float* logmel_bm, float2* spec, cudaStream_t st, float* pre_peak = nullptr);
int stft_frames_per_block(const b2d_plan* p);
int launch_mel_scale(const b2d_plan* p, const float* mag, int B, int T, float* mel, cudaStream_t st);
int launch_residual(const float* x, const float* pred, float* out, size_t n, int mode, float out_scale, cudaStream_t st);
int launch_inverse_mel(const b2d_plan* p, const float* mel, int B, int T, float* out, bool torch_layout, cudaStream_t st);
int launch_istft(const b2d_plan* p, const float2* spec, const float* mag, int B, int T, float* wave, cudaStream_t st);
int launch_to_frame_layout(const float* in, float* out, int B, int F, int T, int Fp, cudaStream_t st);

// griffinlim.cu
struct GlPartition {
  int n;  // frames per run
  int R;  // runs per clip
  int G;  // frames per round inside a run
  int fast;  // 0 = generic shared-memory kernel, 1 = n_fft 1024 register kernel, 2 = n_fft 512 (two frames per transform),
             // 3 = n_fft 2048 (warp pair), 4 = warp-synchronous Stockham (gl_warp.cu), 5 = generic-radix register FFT (gl_reg.cu),
             // 6 = the same in a single launch, one cluster of `csize` CTAs per clip (small problems)
             // 7 = a streaming hop (T <= 4) in a single launch, one CTA per session, iterates in shared memory
             // 8 = small / medium problems in one cooperative launch over the whole GPU (grid barrier between iterations)
  int csize;
};
GlPartition gl_partition(const b2d_plan* p, int B, int T);
size_t gl_workspace_bytes(const b2d_plan* p, int B, int T, bool need_mag_copy);
// ola / hop_out: streaming hop only, honoured when gl_fuses_ola(p, B, T) -- the overlap-add ring update of app3.py:219-224 runs
// behind the last iteration inside the single-launch hop kernel and `wave` is not written
int gl_run(const b2d_plan* p, const float* mag_tf, const float2* init_angles, unsigned long long seed, int B, int T, int n_iter,
           float momentum, const float* out_scale, float* wave, void* ws, size_t ws_bytes, cudaStream_t st,
           const unsigned long long* seed_ptr = nullptr, float* ola = nullptr, float* hop_out = nullptr);
bool gl_fuses_ola(const b2d_plan* p, int B, int T);

// model.cu
size_t model_workspace_bytes(const b2d_model* m, int B, int T);
// fused_mode: 0 = write pred only, 1 = also mel (app3 residual), 2 = also mel (server residual, out_scale)
int model_forward(const b2d_model* m, const float* x, float* hx, float* pred, float* mel_bt, int fused_mode,
                  float out_scale, int B, int T, int conv_mode, void* ws, size_t ws_bytes, cudaStream_t st);
int scale_inplace(float* p, size_t n, float s, cudaStream_t st);

}  // namespace b2d
