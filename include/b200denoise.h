/*
 * b200denoise.h -- C-ABI of libb200denoise.so: the B200 (sm_100a) implementation of the
 * belacks/audio-denoising inference hot path
 *
 *   waveform -> STFT -> Mel log-magnitude -> GRUUNet2 -> inverse Mel -> Griffin-Lim -> iSTFT/OLA
 *
 * The reference is pure Python: it has no FFI.  Each entry point below replaces one *library
 * call site* of the reference (torchaudio / torch ops, or the reference's own nn.Module), cited
 * as file:line into the reference tree (`TA:` = torchaudio site-packages).  INTEGRATION.md shows
 * the ctypes binding a maintainer of the reference would add.
 *
 * Conventions
 *   - every pointer argument is a raw pointer; `const float* wave` etc. are DEVICE pointers unless
 *     the name starts with `h_` (host);  sizes are plain ints;  `stream` is a cudaStream_t passed
 *     as void* (NULL = legacy default stream).
 *   - return value: 0 = B2D_OK, <0 = error; b2d_last_error_string() describes the last error of
 *     the calling thread.  Nothing throws, nothing synchronises the stream, nothing allocates
 *     device memory after plan/model creation: scratch space is a caller-owned workspace sized by
 *     the matching *_workspace_bytes() call.
 *   - plans and models are immutable after creation and may be shared by threads/streams;
 *     per-stream mutable state (hx, Griffin-Lim iterates, streaming rings) lives in caller buffers.
 *   - layouts: "torch layout" spectrogram = [B, F, T] (time contiguous, F = n_fft/2+1,
 *     T = 1 + L/hop); "frame layout" = [B, T, Fp] with Fp = n_fft/2 + 4 (frequency contiguous,
 *     16-byte aligned rows) is what the fused chain uses internally.
 *   - there is NO CPU fallback: every compute entry point launches sm_100a kernels.
 */
#ifndef B200DENOISE_H_
#define B200DENOISE_H_

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B2D_OK 0
#define B2D_ERR_BAD_ARG (-1)     /* NULL pointer, non-positive size, inconsistent shapes        */
#define B2D_ERR_UNSUPPORTED (-2) /* geometry / model config outside what the kernels implement  */
#define B2D_ERR_ALIGN (-3)       /* pointer not aligned as required (16 B for wave/spec rows)   */
#define B2D_ERR_CUDA (-4)        /* a CUDA runtime call or launch failed                        */
#define B2D_ERR_WORKSPACE (-5)   /* workspace NULL or smaller than *_workspace_bytes()          */

typedef struct b2d_plan b2d_plan;   /* DSP geometry + constant tables (twiddles, window, mel, pinv) */
typedef struct b2d_model b2d_model; /* packed GRUUNet2 weights                                      */

/* Complex numbers are interleaved (re, im) float pairs == torch.complex64 == float2. */
typedef struct { float re, im; } b2d_c64;

int b2d_version(void);
const char* b2d_last_error_string(void);

/* ---- plan -------------------------------------------------------------------------------------
 * Geometry of Spectrogram/MelScale/InverseMelScale/GriffinLim as constructed at app3.py:135-153 and
 * server.py:173-176: win_length == n_fft, periodic Hann, center=True, reflect padding, onesided.
 * h_mel_fb: host [F, n_mels] row-major triangular filterbank (TA:functional/functional.py:518-588),
 * h_pinv  : host [F, n_mels] row-major pinv(fb^T) (min-norm lstsq of TA:transforms/_transforms.py:508).
 * The STFT supports any 1 <= hop <= n_fft; iSTFT / Griffin-Lim require hop == n_fft/2 (the only
 * setting the reference uses: app3.py:29-33, server.py:166-170).
 * n_fft must be even with n_fft/2 = 2^a 3^b 5^c and 64 <= n_fft <= 4096. */
int b2d_plan_create(int n_fft, int hop, int n_mels, const float* h_mel_fb, const float* h_pinv, b2d_plan** out);
/* The same with option flags, fixed for the life of the plan (nothing on the compute path reads the environment):
 * B2D_PLAN_GENERIC_KERNELS  every transform runs on the generic shared-memory Stockham kernels -- the cross-check engine
 *                           the tests hold the register / TMA fast paths against;
 * the B2D_PLAN_EXACT_* / FP32_* flags swap one approximation each for the exact operation the reference performs, so that
 * tests/test_gpu_parity.py::test_exact_math_attribution can attribute the end-to-end deviation from the CPU oracle:
 *   EXACT_SQRT      |STFT| with an IEEE square root instead of sqrt.approx.ftz (app3.py:192)
 *   EXACT_UNIT      Griffin-Lim projection a / (|a| + 1e-16) with IEEE sqrt + division instead of rsqrt.approx (TA functional.py:343)
 *   EXACT_PEAK_DIV  x / peak as a division instead of x * (1 / peak) (app3.py:183)
 *   FP32_INVMEL     inverse mel as an fp32 FMA GEMM instead of the TF32x3 tcgen05 GEMM
 * (GENERIC_KERNELS also forms the Griffin-Lim momentum term in the frequency domain from a stored complex `tprev`, literally
 *  as TA functional.py:337-341 does; the n_fft = 1024 fast path forms it on the time-domain iterates, see csrc/gl_fast.cu.) */
#define B2D_PLAN_GENERIC_KERNELS 1u
#define B2D_PLAN_EXACT_SQRT 2u
#define B2D_PLAN_EXACT_UNIT 4u
#define B2D_PLAN_EXACT_PEAK_DIV 8u
#define B2D_PLAN_FP32_INVMEL 16u
/* CLUSTER_GL: small problems keep the round-2a single-launch Griffin-Lim (one thread-block cluster per clip, barrier.cluster
 * between iterations) instead of the cooperative whole-GPU kernel (grid barrier) -- same arithmetic, kept for the tests. */
#define B2D_PLAN_CLUSTER_GL 32u
int b2d_plan_create_ex(int n_fft, int hop, int n_mels, const float* h_mel_fb, const float* h_pinv, unsigned flags, b2d_plan** out);
void b2d_plan_destroy(b2d_plan* plan);
int b2d_plan_num_frames(const b2d_plan* plan, int L);     /* T = 1 + L / hop                 */
int b2d_plan_output_length(const b2d_plan* plan, int T);  /* hop * (T - 1)  (length=None)    */
int b2d_plan_frame_stride(const b2d_plan* plan);          /* Fp = n_fft/2 + 4                */

/* ---- model ------------------------------------------------------------------------------------
 * GRUUNet2(num_compressed_bins, in_size=1, hidden_sizes, kernel_sizes, strides, paddings,
 * num_gaussians) -- gruunet2.py:246-264.  h_params: n_params host pointers in state_dict
 * `parameters()` order: input_gate.downs.{i}.conv.{weight,bias} (levels),
 * reset_gate.downs.0.conv.{weight,bias}, output_gate.ups.{i}.conv.{weight,bias} (levels).
 * h_gs_offsets: 3 host pointers to the `gs.offset` buffers (input, reset, output gate), each
 * [num_gaussians].  Supported: uniform hidden size <= 32, kernel 3, stride 2, padding 1,
 * levels in [1, 6]; anything else -> B2D_ERR_UNSUPPORTED. */
typedef struct {
  int num_compressed_bins;
  int hidden;        /* hidden_sizes[i], all equal */
  int levels;        /* len(hidden_sizes)          */
  int kernel, stride, padding;
  int num_gaussians;
} b2d_model_config;

int b2d_model_create(const b2d_model_config* cfg, const float* const* h_params, int n_params,
                     const float* const* h_gs_offsets, b2d_model** out);
void b2d_model_destroy(b2d_model* model);
int b2d_model_n_mels(const b2d_model* model); /* num_compressed_bins << levels */

/* ---- the same cell for ANY configuration, and its sibling MOMO3 ---------------
 * b2d_model above is the tuned engine of the shipped GRUUNet2 configuration.  b2d_cell runs the reference's conv-GRU U-Net
 * cell generically (fp32 FMA kernels, same three phases): per-level hidden sizes, kernel sizes, strides, paddings, any
 * n_mels that compresses to num_compressed_bins (odd lengths, padding 0).
 *   B2D_ARCH_GRUUNET2  gruunet2.py:71-306 (and gruunet.py: the same maths)
 *   B2D_ARCH_MOMO3     momo3.py:191-324: input channels (x_t, x_t - x_{t-1}), Gaussian channels at the encoder input only
 * h_params: state_dict parameters() order as for b2d_model_create; h_gs_offsets: input, reset (and output for GRUUNET2) gate.
 * b2d_cell_forward: x [B, T, n_mels], prev [B, n_mels] or NULL (MOMO3: the frame before x[:, 0]; NULL = x[:, 0] itself,
 * momo3.py:277-278), hx [B, hidden[-1], bins] in/out, out [B, T, n_mels]. */
#define B2D_ARCH_GRUUNET2 0
#define B2D_ARCH_MOMO3 1
typedef struct b2d_cell b2d_cell;
typedef struct {
  int arch;
  int num_compressed_bins;
  int levels;
  int num_gaussians;
  int n_mels;
  int hidden[8], kernel[8], stride[8], padding[8];
} b2d_cell_config;
int b2d_cell_create(const b2d_cell_config* cfg, const float* const* h_params, int n_params,
                    const float* const* h_gs_offsets, b2d_cell** out);
void b2d_cell_destroy(b2d_cell* cell);
size_t b2d_cell_workspace_bytes(const b2d_cell* cell, int B, int T);
int b2d_cell_forward(const b2d_cell* cell, const float* x, const float* prev, float* hx, float* out, int B, int T,
                     void* workspace, size_t workspace_bytes, void* stream);

/* ---- fp32 backward of the cell (fine-tuning; server.py:86-142 wraps the module in AdamW) --------
 * b2d_cell_backward consumes what b2d_cell_forward left in ITS workspace (pass the same buffer, untouched, as forward_workspace)
 * and the inputs of that forward call (x, prev, the hx the call STARTED from), plus grad_out [B, T, n_mels] and grad_hx_out
 * [B, hidden[-1], bins] (nullable = zero).  It writes grad_x [B, T, n_mels], grad_hx_in [B, hidden[-1], bins] and grad_params:
 * b2d_cell_num_param_floats() floats, the gradients of every parameter tensor in state_dict parameters() order and torch
 * layout, back to back (weights of the Gaussian position channels and biases included).  Deterministic up to the order of the
 * float atomics that combine per-CTA partial sums. */
int b2d_cell_num_param_floats(const b2d_cell* cell);
size_t b2d_cell_backward_workspace_bytes(const b2d_cell* cell, int B, int T);
int b2d_cell_backward(const b2d_cell* cell, const float* x, const float* prev, const float* hx_in, const void* forward_workspace,
                      const float* grad_out, const float* grad_hx_out, float* grad_x, float* grad_hx_in, float* grad_params,
                      int B, int T, void* workspace, size_t workspace_bytes, void* stream);

/* ---- K0: per-clip peak (app3.py:181-186) ------------------------------------------------------
 * peak[b] = max|wave[b,:]| if > 1e-6 else 1. */
int b2d_peak(const float* wave, int B, int L, float* peak, void* stream);

/* ---- K1: Spectrogram(power=None) forward (app3.py:191, server.py:207; TA:functional/functional.py:54-145)
 * wave [B, L] -> spec [B, F, T] complex (torch layout).  L > n_fft/2 (reflect padding). */
int b2d_stft(const b2d_plan* plan, const float* wave, int B, int L, b2d_c64* spec, void* stream);

/* ---- K1+K2 fused: log1p(MelScale(|Spectrogram(x)|)) (app3.py:191-193, server.py:207-210) --------
 * Any of the outputs may be NULL.  logmel_bt [B, T, n_mels] is the model-input layout
 * (app3.py:195 transposes to it), logmel_bm [B, n_mels, T] is MelScale's own layout, spec as b2d_stft.
 * inv_scale (nullable, [B]): every sample of clip b is divided by inv_scale[b] first (a1 fused). */
int b2d_stft_mel_log1p(const b2d_plan* plan, const float* wave, const float* inv_scale, int B, int L,
                       float* logmel_bt, float* logmel_bm, b2d_c64* spec, void* stream);

/* ---- K2 alone: MelScale.forward (TA:transforms/_transforms.py:407-419): mag [B,F,T] -> mel [B,n_mels,T] */
int b2d_mel_scale(const b2d_plan* plan, const float* mag, int B, int T, float* mel, void* stream);

/* ---- K3: GRUUNet2.forward (gruunet2.py:290-306) -----------------------------------------------
 * x [B, T, n_mels], hx [B, hidden, bins] in/out (caller zero-fills it for hx=None), out [B, T, n_mels].
 * Runs encoder (time-parallel) -> persistent recurrence -> decoder (time-parallel).
 * conv_mode (low byte) selects the convolution engine:
 *   B2D_CONV_MMA   (default) encoder + decoder on warp-level m16n8k8 TF32 tensor-core MMAs, every operand split into big +
 *                  small TF32 parts (3 MMAs per tile: fp32-class accuracy); the fastest engine for this 17-channel model;
 *   B2D_CONV_UTC   the encoder as ONE persistent tcgen05 / TMEM kernel (implicit GEMMs chained through shared memory, weights by
 *                  TMA, same 3 x TF32 split) + the warp-MMA decoder;
 *   B2D_CONV_FP32  fp32 CUDA-core FMA kernels: the exact engine the attribution test compares the others with.
 * B2D_CONV_EXACT_GATES or-ed in: GRU gate sigmoid / tanh through expf and IEEE division instead of ex2.approx / rcp.approx
 * (the reference's torch.sigmoid / torch.tanh, gruunet2.py:231-240) -- the model's switch for test_exact_math_attribution. */
#define B2D_CONV_FP32 0
#define B2D_CONV_MMA 3
#define B2D_CONV_UTC 5
#define B2D_CONV_EXACT_GATES 0x100
size_t b2d_gruunet2_workspace_bytes(const b2d_model* model, int B, int T);
int b2d_gruunet2_forward(const b2d_model* model, const float* x, float* hx, float* out, int B, int T,
                         int conv_mode, void* workspace, size_t workspace_bytes, void* stream);

/* ---- K4: residual + nonlinearity (app3.py:203-208 / server.py:213-215) -------------------------
 * mode 0 (app):    mel = max(expm1(leaky_relu(x - pred, 0.2)), 0)
 * mode 1 (server): mel = exp(x - relu(pred) * out_scale) - 1            (out_scale = 3)
 * x, pred, mel_bt: [B, T, n_mels] (n = B*T*n_mels elements). */
int b2d_residual_mel(const float* x, const float* pred, float* mel_bt, size_t n, int mode, float out_scale, void* stream);

/* ---- K5: InverseMelScale.forward (TA:transforms/_transforms.py:491-512) = relu(pinv(fb^T) @ mel)
 * torch layout: mel [B, n_mels, T] -> lin [B, F, T]. */
int b2d_inverse_mel(const b2d_plan* plan, const float* mel, int B, int T, float* lin, void* stream);
/* frame layout: mel_bt [B, T, n_mels] -> mag_tf [B, T, Fp] (columns F..Fp-1 are written as 0). */
int b2d_inverse_mel_frames(const b2d_plan* plan, const float* mel_bt, int B, int T, float* mag_tf, void* stream);

/* ---- K6: GriffinLim(power=1) (app3.py:213; TA:functional/functional.py:255-353) -----------------
 * mag [B, F, T] torch layout (b2d_griffinlim) or [B, T, Fp] frame layout (b2d_griffinlim_frames);
 * init_angles [B, F, T] complex torch layout = the `angles` tensor of TA functional.py:309-312
 * (when NULL: seed != 0 -> rand_init=True with in-kernel counter-based U[0,1) draws for real and imaginary parts,
 * seed == 0 -> rand_init=False, all ones);  momentum is the user-facing value (0.99), rescaled inside as
 * at TA functional.py:300;  out_scale (nullable, [B]) multiplies clip b's waveform (app3.py:217).
 * wave [B, hop*(T-1)].  T >= 3. */
size_t b2d_griffinlim_workspace_bytes(const b2d_plan* plan, int B, int T);
int b2d_griffinlim(const b2d_plan* plan, const float* mag, const b2d_c64* init_angles, unsigned long long seed,
                   int B, int T, int n_iter, float momentum, const float* out_scale, float* wave,
                   void* workspace, size_t workspace_bytes, void* stream);
int b2d_griffinlim_frames(const b2d_plan* plan, const float* mag_tf, const b2d_c64* init_angles, unsigned long long seed,
                          int B, int T, int n_iter, float momentum, const float* out_scale, float* wave,
                          void* workspace, size_t workspace_bytes, void* stream);

/* ---- K7: InverseSpectrogram / torch.istft (server.py:216; TA:functional/functional.py:205) ------
 * spec [B, F, T] complex torch layout -> wave [B, hop*(T-1)].  If mag != NULL the spectrum used is
 * torch.polar(mag, angle(spec)) = mag * spec/|spec| (server.py:216 with phase from :208). */
int b2d_istft(const b2d_plan* plan, const b2d_c64* spec, const float* mag, int B, int T, float* wave, void* stream);

/* ---- whole chain, app3.py:181-217 applied to whole clips ---------------
 * noisy [B, L] -> wave [B, hop*(T-1)].  hx [B, hidden, bins] in/out.  normalise != 0 applies the
 * per-clip peak normalisation of app3.py:181-186 and the final `* peak` of :217.
 * Optional debug outputs (nullable): logmel_bt, pred_bt [B,T,n_mels], mag_tf [B,T,Fp]. */
size_t b2d_denoise_workspace_bytes(const b2d_plan* plan, const b2d_model* model, int B, int L);
int b2d_denoise_batch(const b2d_plan* plan, const b2d_model* model, const float* noisy, int B, int L,
                      float* hx, const b2d_c64* init_angles, unsigned long long seed, int n_iter, float momentum,
                      int normalise, int conv_mode, float* wave, float* logmel_bt, float* pred_bt, float* mag_tf,
                      void* workspace, size_t workspace_bytes, void* stream);

/* ---- the same chain on the int16 PCM link (what app3.py's recv exchanges: :168-172 in, :244-245 out) ----
 * pcm [B, L] int16 -> x = pcm / 32767 (ingest fused with the peak pass) -> b2d_denoise_batch -> pcm_out [B, hop*(T-1)] int16
 * = (clip(wave, -1, 1) * 32767) truncated toward zero.  The float staging lives in the workspace. */
size_t b2d_denoise_pcm16_workspace_bytes(const b2d_plan* plan, const b2d_model* model, int B, int L);
int b2d_denoise_batch_pcm16(const b2d_plan* plan, const b2d_model* model, const short* pcm, int B, int L,
                            float* hx, const b2d_c64* init_angles, unsigned long long seed, int n_iter, float momentum,
                            int normalise, int conv_mode, short* pcm_out,
                            void* workspace, size_t workspace_bytes, void* stream);

/* ---- server.py:207-216 chain (noisy-phase iSTFT, no Griffin-Lim) --------------------------------
 * x [B, L] -> wave [B, hop*(T-1)];  hx in/out and multiplied by hx_decay (0.9) afterwards. */
size_t b2d_denoise_noisy_phase_workspace_bytes(const b2d_plan* plan, const b2d_model* model, int B, int L);
int b2d_denoise_noisy_phase(const b2d_plan* plan, const b2d_model* model, const float* x, int B, int L,
                            float* hx, float out_scale, float hx_decay, int conv_mode, float* wave,
                            void* workspace, size_t workspace_bytes, void* stream);

/* ---- streaming hop, app3.py:178-226 (one `while` iteration for S independent sessions) ----------
 * chunk [S, n_fft] raw float samples (the current input window), hx [S, hidden, bins] in/out,
 * ola [S, n_fft] in/out output overlap-add ring, out [S, hop] the hop of audio emitted by this step.
 * init_angles [S, F, 3] or NULL (then seed as for b2d_griffinlim; if d_seed != NULL the seed is read from that device
 * word at run time instead, so a captured CUDA graph of this call can be replayed with a fresh seed per hop).
 * Quirks Q2-Q4 of SURVEY.md Appendix C are kept. */
size_t b2d_stream_step_workspace_bytes(const b2d_plan* plan, const b2d_model* model, int S);
int b2d_stream_step(const b2d_plan* plan, const b2d_model* model, const float* chunk, int S, float* hx,
                    float* ola, const b2d_c64* init_angles, unsigned long long seed, const unsigned long long* d_seed,
                    int n_iter, float momentum, int conv_mode, float* out, void* workspace, size_t workspace_bytes, void* stream);

/* ---- audio ingest adjacent to the path ------------------------------
 * b2d_pcm16_to_float: int16 PCM [n, channels] (sample-major, as `frame.to_ndarray(format="s16")` reshaped by app3.py:168-170)
 *   -> float32 [n] = pcm[:, channel] / 32767 (app3.py:172), or the mean over channels when channel < 0 (app.py:184-186).
 * b2d_float_to_pcm16: (clip(x, -1, 1) * 32767).astype(int16)  (app3.py:244-245).
 * b2d_resample*: torchaudio.transforms.Resample(orig, new) of utils.py:48-49 (TA:functional/functional.py:1305-1432): the
 *   caller passes the sinc-Hann polyphase table h_kernel [new_reduced, taps] (rates reduced by their gcd, taps = 2*width +
 *   orig_reduced); in [B, L] -> out [B, ceil(new * L / orig)]. */
typedef struct b2d_resampler b2d_resampler;
int b2d_pcm16_to_float(const short* pcm, size_t n, int channels, int channel, float* out, void* stream);
int b2d_float_to_pcm16(const float* in, size_t n, short* out, void* stream);
int b2d_resampler_create(int orig_reduced, int new_reduced, int taps, int width, const float* h_kernel, b2d_resampler** out);
void b2d_resampler_destroy(b2d_resampler* r);
int b2d_resample_length(const b2d_resampler* r, int L);
int b2d_resample(const b2d_resampler* r, const float* in, int B, int L, float* out, void* stream);

/* Number of kernels this library has launched from the calling process (all threads); bench.py
 * reports the difference across the timed region as "gpu_launches". */
unsigned long long b2d_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* B200DENOISE_H_ */
