// dsp.cu -- K0 peak, K1/K2 STFT (+Mel+log1p), K2 MelScale, K4 residual, K5 inverse Mel, K7 iSTFT.
// Generic over every supported n_fft (shared-memory Stockham FFT, fft.cuh).
#include <stdlib.h>

#include "fft.cuh"
#include "kernels.cuh"

namespace b2d {

// ------------------------------------------------------------------------------------------------
// K0: peak[b] = max |x| (app3.py:181-186)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) peak_partial_kernel(const float* __restrict__ wave, int L, int chunks,
                                                           float* __restrict__ partial) {
  const int b = blockIdx.y;
  const float* x = wave + (size_t)b * L;
  const int per = ((L + chunks - 1) / chunks + 3) & ~3;
  const int lo = min(L, (int)blockIdx.x * per), hi = min(L, lo + per);
  float m = 0.f;
  if ((((size_t)x) & 15) == 0) {  // 16-byte aligned clip row: float4 body, four loads in flight per thread
    const float4* x4 = reinterpret_cast<const float4*>(x + lo);
    const int n4 = (hi - lo) >> 2;
    int i = threadIdx.x;
    for (; i + 3 * (int)blockDim.x < n4; i += 4 * blockDim.x) {
      const float4 a = x4[i], b = x4[i + blockDim.x], c = x4[i + 2 * blockDim.x], d = x4[i + 3 * blockDim.x];
      m = fmaxf(m, fmaxf(fmaxf(fmaxf(fabsf(a.x), fabsf(a.y)), fmaxf(fabsf(a.z), fabsf(a.w))), fmaxf(fmaxf(fabsf(b.x), fabsf(b.y)), fmaxf(fabsf(b.z), fabsf(b.w)))));
      m = fmaxf(m, fmaxf(fmaxf(fmaxf(fabsf(c.x), fabsf(c.y)), fmaxf(fabsf(c.z), fabsf(c.w))), fmaxf(fmaxf(fabsf(d.x), fabsf(d.y)), fmaxf(fabsf(d.z), fabsf(d.w)))));
    }
    for (; i < n4; i += blockDim.x) {
      const float4 a = x4[i];
      m = fmaxf(m, fmaxf(fmaxf(fabsf(a.x), fabsf(a.y)), fmaxf(fabsf(a.z), fabsf(a.w))));
    }
    for (int j = lo + 4 * n4 + threadIdx.x; j < hi; j += blockDim.x) m = fmaxf(m, fabsf(x[j]));
  } else {
    for (int i = lo + threadIdx.x; i < hi; i += blockDim.x) m = fmaxf(m, fabsf(x[i]));
  }
  for (int o = 16; o; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  __shared__ float s[8];
  if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x < 8) {
    m = s[threadIdx.x];
    for (int o = 4; o; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffu, m, o));
    if (threadIdx.x == 0) partial[b * chunks + blockIdx.x] = m;
  }
}
// int16 PCM rows -> float (x / 32767, app3.py:172) with the peak partials of K0 taken in the same pass (the float copy is
// written once and never re-read for the peak): the ingest of the int16 host link (b2d_denoise_batch_pcm16)
__global__ void __launch_bounds__(256) pcm16_ingest_peak_kernel(const short* __restrict__ pcm, int L, int chunks,
                                                                float* __restrict__ wave, float* __restrict__ partial) {
  const int b = blockIdx.y;
  const short* x = pcm + (size_t)b * L;
  float* y = wave + (size_t)b * L;
  const int per = ((L + chunks - 1) / chunks + 7) & ~7;
  const int lo = min(L, (int)blockIdx.x * per), hi = min(L, lo + per);
  float m = 0.f;
  if (((((size_t)x) & 15) | (((size_t)y) & 15)) == 0) {  // 8 samples per thread: one 16-byte load, two 16-byte stores
    const int4* x8 = reinterpret_cast<const int4*>(x + lo);
    float4* y4 = reinterpret_cast<float4*>(y + lo);
    const int n8 = (hi - lo) >> 3;
    for (int i = threadIdx.x; i < n8; i += blockDim.x) {
      const int4 q = x8[i];
      const int w[4] = {q.x, q.y, q.z, q.w};
      float v[8];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        v[2 * e] = (float)(short)(w[e] & 0xffff) / 32767.0f;
        v[2 * e + 1] = (float)(short)(w[e] >> 16) / 32767.0f;
      }
      y4[2 * i] = make_float4(v[0], v[1], v[2], v[3]);
      y4[2 * i + 1] = make_float4(v[4], v[5], v[6], v[7]);
#pragma unroll
      for (int e = 0; e < 8; ++e) m = fmaxf(m, fabsf(v[e]));
    }
    for (int j = lo + 8 * n8 + threadIdx.x; j < hi; j += blockDim.x) {
      const float v = (float)x[j] / 32767.0f;
      y[j] = v;
      m = fmaxf(m, fabsf(v));
    }
  } else {
    for (int j = lo + threadIdx.x; j < hi; j += blockDim.x) {
      const float v = (float)x[j] / 32767.0f;
      y[j] = v;
      m = fmaxf(m, fabsf(v));
    }
  }
  for (int o = 16; o; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  __shared__ float s[8];
  if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x < 8) {
    m = s[threadIdx.x];
    for (int o = 4; o; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffu, m, o));
    if (threadIdx.x == 0) partial[b * chunks + blockIdx.x] = m;
  }
}
__global__ void peak_final_kernel(const float* __restrict__ partial, int B, int chunks, float* __restrict__ peak) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  float m = 0.f;
  for (int c = 0; c < chunks; ++c) m = fmaxf(m, partial[b * chunks + c]);
  peak[b] = (m > 1e-6f) ? m : 1.0f;
}

// ------------------------------------------------------------------------------------------------
// K1 (+K2): frames -> window -> rFFT (-> |.| -> mel -> log1p)
// grid (ceil(T/G), B), 256 threads, G frames per block.
// ------------------------------------------------------------------------------------------------
struct StftArgs {
  const float* wave;
  const float* inv_scale;
  int B, L, T, G;
  int n_fft, hop, M, F, n_mels;
  FftDesc fd;
  const float2* tw;
  const float2* rtw;
  const float* win;
  const int* mel_lo;
  const int* mel_cnt;
  const int* mel_off;
  const float* mel_w;
  float2* spec;      // [B,F,T] or null
  float* logmel_bt;  // [B,T,n_mels] or null
  float* logmel_bm;  // [B,n_mels,T] or null
  // streaming hop (stream.cu): the chunk conditioning of app3.py:179-188 fused in -- wave = (chunk / peak) * win with
  // peak[b] = max|chunk[b]| (1 when <= 1e-6), written to pre_peak; one CTA per session (T <= G), wave is not read
  const float* pre_chunk;  // [B, L] raw chunk (device or pinned host memory) or null
  float* pre_peak;         // [B]
};

// MT: complex transform length known at compile time (fft_rows_t: 320 / 512 / 768) or 0
template <int MT>
__global__ void __launch_bounds__(512) stft_kernel(const StftArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int M = MT ? MT : a.M, N = MT ? 2 * MT : a.n_fft, G = a.G, hop = a.hop;
  float2* tw_s = reinterpret_cast<float2*>(smem_raw);
  float2* bufA = tw_s + M;
  float2* bufB = bufA + G * M;
  float* xin = reinterpret_cast<float*>(bufB + G * M);
  const int xlen = (G - 1) * hop + N;
  float* mag_s = xin + xlen;  // [G][M+1]

  const int b = blockIdx.y;
  const int t0 = blockIdx.x * G;
  const int p = N / 2;
  const FastDiv d_M(M), d_half(M / 2 + 1), d_mel(a.n_mels > 0 ? a.n_mels : 1);  // index math without integer division (fft.cuh)
  const float* x = a.wave + (size_t)b * a.L;
  for (int i = threadIdx.x; i < M; i += blockDim.x) tw_s[i] = a.tw[i];
  pdl_wait();  // the twiddle table is a plan constant; the waveform and the scale come from the kernels before
  pdl_trigger();
  const float sc = a.inv_scale ? a.inv_scale[b] : 1.0f;
  float* raw = reinterpret_cast<float*>(bufB);  // pre mode: the raw chunk (L <= 2 G M floats), read once from (possibly host) memory
  float pk = 1.0f;
  if (a.pre_chunk) {
    __shared__ float red[32];
    __shared__ float pk_s;
    const float* c = a.pre_chunk + (size_t)b * a.L;
    float m = 0.f;
    for (int i = threadIdx.x; i < a.L; i += blockDim.x) {
      const float v = c[i];
      raw[i] = v;
      m = fmaxf(m, fabsf(v));
    }
    for (int o = 16; o; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = m;
    __syncthreads();
    if (threadIdx.x == 0) {
      float v = red[0];
      for (int i = 1; i < (int)(blockDim.x >> 5); ++i) v = fmaxf(v, red[i]);
      pk_s = (v > 1e-6f) ? v : 1.0f;
      a.pre_peak[b] = pk_s;
    }
    __syncthreads();
    pk = pk_s;
  }
  const int padded = a.L + 2 * p;
  for (int i = threadIdx.x; i < xlen; i += blockDim.x) {
    const int c = t0 * hop + i;
    float v = 0.f;
    if (c < padded) {
      int s = c - p;
      if (s < 0) s = -s;
      if (s >= a.L) s = 2 * (a.L - 1) - s;
      v = a.pre_chunk ? (raw[s] / pk) * a.win[s] : x[s] / sc;
    }
    xin[i] = v;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < G * M; i += blockDim.x) {
    const int g = d_M.div(i), m = i - g * M;
    float2 z = make_float2(0.f, 0.f);
    if (t0 + g < a.T) {
      const float* xf = xin + g * hop + 2 * m;  // scalar loads: hop may be odd
      z = make_float2(xf[0] * a.win[2 * m], xf[1] * a.win[2 * m + 1]);
    }
    bufA[i] = z;
  }
  __syncthreads();
  float2* res = fft_rows_t<false, MT>(bufA, bufB, G, M, a.fd, tw_s);
  // split into the one-sided spectrum
  const int half = M / 2 + 1;
  for (int i = threadIdx.x; i < G * half; i += blockDim.x) {
    const int g = d_half.div(i), k = i - g * half;
    const int t = t0 + g;
    if (t >= a.T) continue;
    const float2 zk = res[g * M + k];
    const float2 zmk = res[g * M + ((M - k) & -(k != 0))];  // k == 0 pairs with itself
    float2 xk, xmk;
    rfft_split(zk, zmk, a.rtw[k], xk, xmk);
    if (k == 0) { xk.y = 0.f; xmk.y = 0.f; }
    if (a.spec) {
      a.spec[((size_t)b * a.F + k) * a.T + t] = xk;
      a.spec[((size_t)b * a.F + (M - k)) * a.T + t] = xmk;
    }
    mag_s[g * (M + 1) + k] = sqrtf(xk.x * xk.x + xk.y * xk.y);
    mag_s[g * (M + 1) + (M - k)] = sqrtf(xmk.x * xmk.x + xmk.y * xmk.y);
  }
  if (a.logmel_bt == nullptr && a.logmel_bm == nullptr) return;
  __syncthreads();
  for (int i = threadIdx.x; i < G * a.n_mels; i += blockDim.x) {
    const int g = d_mel.div(i), m = i - g * a.n_mels;
    const int t = t0 + g;
    if (t >= a.T) continue;
    const int lo = a.mel_lo[m], cnt = a.mel_cnt[m];
    const float* w = a.mel_w + a.mel_off[m];
    const float* mg = mag_s + g * (M + 1) + lo;
    float acc = 0.f;
    for (int q = 0; q < cnt; ++q) acc = fmaf(mg[q], w[q], acc);
    const float v = log1pf(acc);
    if (a.logmel_bt) a.logmel_bt[((size_t)b * a.T + t) * a.n_mels + m] = v;
    if (a.logmel_bm) a.logmel_bm[((size_t)b * a.n_mels + m) * a.T + t] = v;
  }
}

// ------------------------------------------------------------------------------------------------
// K2 alone: MelScale on an arbitrary magnitude tensor, torch layout [B,F,T] -> [B,n_mels,T]
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) mel_scale_kernel(const float* __restrict__ mag, int F, int T, int n_mels,
                                                        const int* __restrict__ mel_lo, const int* __restrict__ mel_cnt,
                                                        const int* __restrict__ mel_off, const float* __restrict__ mel_w,
                                                        float* __restrict__ mel) {
  const int b = blockIdx.z, m = blockIdx.y;
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= T) return;
  const int lo = mel_lo[m], cnt = mel_cnt[m];
  const float* w = mel_w + mel_off[m];
  const float* src = mag + ((size_t)b * F + lo) * T + t;
  float acc = 0.f;
  for (int q = 0; q < cnt; ++q) acc = fmaf(src[(size_t)q * T], w[q], acc);
  mel[((size_t)b * n_mels + m) * T + t] = acc;
}

// ------------------------------------------------------------------------------------------------
// K4: residual + nonlinearity
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) residual_mel_kernel(const float* __restrict__ x, const float* __restrict__ pred,
                                                           float* __restrict__ out, size_t n, int mode, float out_scale) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) {
    float v;
    if (mode == 0) {
      float r = x[i] - pred[i];
      r = r > 0.f ? r : 0.2f * r;
      v = fmaxf(expm1f(r), 0.f);
    } else {
      v = expf(x[i] - fmaxf(pred[i], 0.f) * out_scale) - 1.0f;
    }
    out[i] = v;
  }
}

// ------------------------------------------------------------------------------------------------
// K5: lin[frame, f] = relu(sum_m mel[frame, m] * P[f, m]) -- 64x64 output tile, 4x4 per thread.
// kTorch = false: mel [NF, n_mels] row-major, out [NF, Fp] row-major (frame layout)
// kTorch = true : mel [B, n_mels, T], out [B, F, T] (torch layout); tiles never straddle clips.
// ------------------------------------------------------------------------------------------------
template <bool kTorch>
__global__ void __launch_bounds__(256) inverse_mel_kernel(const float* __restrict__ mel, const float* __restrict__ pinv,
                                                          float* __restrict__ out, int B, int T, int K, int F, int Fp) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  constexpr int LD = 68;                            // padded row stride (keeps float4 alignment)
  float* As = reinterpret_cast<float*>(smem_raw);  // [K][LD] (frames)
  float* Ps = As + K * LD;                          // [K][LD] (freqs)
  const int f0 = blockIdx.x * 64;
  int b = 0, t0 = 0;
  size_t frame0 = 0;
  int nframes;
  if (kTorch) {
    b = blockIdx.z;
    t0 = blockIdx.y * 64;
    nframes = min(64, T - t0);
  } else {
    frame0 = (size_t)blockIdx.y * 64;
    const size_t NF = (size_t)B * T;
    nframes = (int)min((size_t)64, NF - frame0);
  }
  const int nfreq = min(64, (kTorch ? F : Fp) - f0);
  for (int i = threadIdx.x; i < K * 64; i += blockDim.x) {
    float va, vp;
    if (kTorch) {
      const int k = i >> 6, r = i & 63;  // r = frame index in tile (t contiguous in memory)
      va = (r < nframes) ? mel[((size_t)b * K + k) * T + t0 + r] : 0.f;
      As[k * LD + r] = va;
    } else {
      const int r = i / K, k = i - r * K;  // k contiguous in memory
      va = (r < nframes) ? mel[(frame0 + r) * K + k] : 0.f;
      As[k * LD + r] = va;
    }
    {
      const int r = i / K, k = i - r * K;
      vp = (r < nfreq) ? pinv[(size_t)(f0 + r) * K + k] : 0.f;
      Ps[k * LD + r] = vp;
    }
  }
  __syncthreads();
  // frame layout: tx -> freq (contiguous in out), ty -> frames.  torch layout: tx -> frames, ty -> freq.
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int fr = (kTorch ? tx : ty) * 4, fq = (kTorch ? ty : tx) * 4;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  for (int k = 0; k < K; ++k) {
    const float4 av = *reinterpret_cast<const float4*>(As + k * LD + fr);
    const float4 pv = *reinterpret_cast<const float4*>(Ps + k * LD + fq);
    const float a4[4] = {av.x, av.y, av.z, av.w};
    const float p4[4] = {pv.x, pv.y, pv.z, pv.w};
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a4[i], p4[j], acc[i][j]);
  }
  if (kTorch) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      if (fq + j >= nfreq) continue;
#pragma unroll
      for (int i = 0; i < 4; ++i)
        if (fr + i < nframes) out[((size_t)b * F + f0 + fq + j) * T + t0 + fr + i] = fmaxf(acc[i][j], 0.f);
    }
  } else {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      if (fr + i >= nframes) continue;
      float* o = out + (frame0 + fr + i) * Fp + f0 + fq;
      if (fq + 3 < nfreq) {
        *reinterpret_cast<float4*>(o) = make_float4(fmaxf(acc[i][0], 0.f), fmaxf(acc[i][1], 0.f), fmaxf(acc[i][2], 0.f), fmaxf(acc[i][3], 0.f));
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (fq + j < nfreq) o[j] = fmaxf(acc[i][j], 0.f);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// K7: iSTFT, torch layout in, standard waveform out.  hop == N/2.
// Block handles frames [t0, t0+G) with t0 = blockIdx.x*(G-1) and emits hop-blocks t0+1 .. t0+G-1.
// ------------------------------------------------------------------------------------------------
struct IstftArgs {
  const float2* spec;  // [B,F,T]
  const float* mag;    // [B,F,T] or null: polar(mag, angle(spec))
  int B, T, G;
  int n_fft, hop, M, F;
  FftDesc fd;
  const float2* tw;
  const float2* rtw;
  const float* winn;
  const float* inv_env;
  float* wave;  // [B, hop*(T-1)]
};

__device__ __forceinline__ float2 polar_like(float2 s, float m) {
  const float n = sqrtf(s.x * s.x + s.y * s.y);
  if (n == 0.f) return make_float2(m, 0.f);
  const float r = m / n;
  return make_float2(s.x * r, s.y * r);
}

__global__ void __launch_bounds__(256) istft_kernel(const IstftArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int M = a.M, G = a.G, hop = a.hop;
  float2* tw_s = reinterpret_cast<float2*>(smem_raw);
  float2* bufA = tw_s + M;
  float2* bufB = bufA + G * M;
  const int b = blockIdx.y;
  const int t0 = blockIdx.x * (G - 1);
  for (int i = threadIdx.x; i < M; i += blockDim.x) tw_s[i] = a.tw[i];
  const int half = M / 2 + 1;
  // t fastest so that the strided [B,F,T] reads touch G consecutive elements
  for (int i = threadIdx.x; i < G * half; i += blockDim.x) {
    const int k = i / G, g = i - k * G;
    const int t = t0 + g;
    float2 zk = make_float2(0.f, 0.f), zmk = zk;
    if (t < a.T) {
      const size_t ik = ((size_t)b * a.F + k) * a.T + t, imk = ((size_t)b * a.F + (M - k)) * a.T + t;
      float2 yk = a.spec[ik], ymk = a.spec[imk];
      if (a.mag) { yk = polar_like(yk, a.mag[ik]); ymk = polar_like(ymk, a.mag[imk]); }
      if (k == 0) { yk.y = 0.f; ymk.y = 0.f; }  // C2R ignores Im of DC / Nyquist
      irfft_merge(yk, ymk, a.rtw[k], zk, zmk);
    }
    bufA[g * M + k] = zk;
    if (k != 0 && 2 * k != M) bufA[g * M + M - k] = zmk;
  }
  __syncthreads();
  float2* res = fft_rows<true>(bufA, bufB, G, M, a.fd, tw_s);
  const float* y = reinterpret_cast<const float*>(res);  // row g: 2*M = N floats
  const int N = a.n_fft;
  const int Lout = hop * (a.T - 1);
  for (int i = threadIdx.x; i < (G - 1) * hop; i += blockDim.x) {
    const int c = i / hop, s = i - c * hop;  // output block j = t0 + 1 + c from frames (c, c+1)
    const int j = t0 + 1 + c;
    if (j > a.T - 1) continue;
    const float v = y[c * N + hop + s] * a.winn[hop + s] + y[(c + 1) * N + s] * a.winn[s];
    a.wave[(size_t)b * Lout + (size_t)(j - 1) * hop + s] = v * a.inv_env[s];
  }
}

// ------------------------------------------------------------------------------------------------
// [B,F,T] (torch) -> [B,T,Fp] (frame layout), pad columns zero-filled.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) to_frame_layout_kernel(const float* __restrict__ in, float* __restrict__ out, int F,
                                                              int T, int Fp) {
  __shared__ float tile[32][33];
  const int b = blockIdx.z;
  const int f0 = blockIdx.x * 32, t0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 8 rows per pass
  for (int r = ty; r < 32; r += 8) {
    const int f = f0 + r, t = t0 + tx;
    tile[r][tx] = (f < F && t < T) ? in[((size_t)b * F + f) * T + t] : 0.f;
  }
  __syncthreads();
  for (int r = ty; r < 32; r += 8) {
    const int t = t0 + r, f = f0 + tx;
    if (t < T && f < Fp) out[((size_t)b * T + t) * Fp + f] = tile[tx][r];
  }
}

// ------------------------------------------------------------------------------------------------
// host launchers
// ------------------------------------------------------------------------------------------------
static int frames_per_block(const b2d_plan* p) { return p->M <= 1024 ? 4 : 2; }
int stft_frames_per_block(const b2d_plan* p) { return frames_per_block(p); }

int launch_peak(const float* wave, int B, int L, float* peak, float* partial, int chunks, cudaStream_t st) {
  peak_partial_kernel<<<dim3(chunks, B), 256, 0, st>>>(wave, L, chunks, partial);
  B2D_LAUNCH_CHECK("peak_partial_kernel");
  peak_final_kernel<<<(B + 127) / 128, 128, 0, st>>>(partial, B, chunks, peak);
  B2D_LAUNCH_CHECK("peak_final_kernel");
  return B2D_OK;
}

int launch_pcm16_ingest_peak(const short* pcm, int B, int L, float* wave, float* peak, float* partial, int chunks, cudaStream_t st) {
  pcm16_ingest_peak_kernel<<<dim3(chunks, B), 256, 0, st>>>(pcm, L, chunks, wave, partial);
  B2D_LAUNCH_CHECK("pcm16_ingest_peak_kernel");
  peak_final_kernel<<<(B + 127) / 128, 128, 0, st>>>(partial, B, chunks, peak);
  B2D_LAUNCH_CHECK("peak_final_kernel");
  return B2D_OK;
}

int launch_stft_fast512(const b2d_plan* p, const float* wave, const float* inv_scale, int B, int L, float* logmel_bt,
                        cudaStream_t st);  // gl_fast.cu

bool stft_reg_supported(const b2d_plan* p, int B, int L);  // gl_reg.cu
int launch_stft_reg(const b2d_plan* p, const float* wave, const float* inv_scale, int B, int L, float* logmel_bt, cudaStream_t st);

int launch_stft(const b2d_plan* p, const float* wave, const float* inv_scale, int B, int L, float* logmel_bt,
                float* logmel_bm, float2* spec, cudaStream_t st, float* pre_peak) {
  if (pre_peak) goto generic;  // streaming hop with the chunk conditioning fused in: block-cooperative kernel, one CTA per session
  if (p->n_fft == 1024 && p->hop == 512 && logmel_bt && !logmel_bm && !spec && L >= 1024 && p->n_mels <= 128 && p->mel_seg_pad <= 320 && (long long)B * (1 + L / p->hop) < (1ll << 30) &&
      !(p->flags & B2D_PLAN_GENERIC_KERNELS))
    return launch_stft_fast512(p, wave, inv_scale, B, L, logmel_bt, st);
  // n_fft 640 / 1536 batches: register-FFT kernel (a single streaming hop -- 3 frames per session -- stays on the block kernel)
  if (logmel_bt && !logmel_bm && !spec && !(p->flags & B2D_PLAN_GENERIC_KERNELS) && stft_reg_supported(p, B, L) &&
      (long long)B * (1 + L / p->hop) >= 64)
    return launch_stft_reg(p, wave, inv_scale, B, L, logmel_bt, st);
generic:
  StftArgs a;
  a.wave = wave; a.inv_scale = inv_scale; a.B = B; a.L = L; a.T = 1 + L / p->hop; a.G = frames_per_block(p);
  a.n_fft = p->n_fft; a.hop = p->hop; a.M = p->M; a.F = p->F; a.n_mels = p->n_mels; a.fd = p->fft;
  a.tw = p->d_tw; a.rtw = p->d_rtw; a.win = p->d_win;
  a.mel_lo = p->d_mel_lo; a.mel_cnt = p->d_mel_cnt; a.mel_off = p->d_mel_off; a.mel_w = p->d_mel_w;
  a.spec = spec; a.logmel_bt = logmel_bt; a.logmel_bm = logmel_bm;
  a.pre_chunk = nullptr; a.pre_peak = nullptr;
  if (pre_peak) {  // streaming hop: `wave` is the raw chunk, conditioning fused in (one CTA per session)
    B2D_REQUIRE(L == p->n_fft && a.T <= a.G && !inv_scale, B2D_ERR_UNSUPPORTED, "fused chunk conditioning needs one n_fft window per session");
    a.pre_chunk = wave; a.pre_peak = pre_peak;
  }
  const int xlen = (a.G - 1) * p->hop + p->n_fft;
  const size_t smem = sizeof(float2) * (size_t)(p->M + 2 * a.G * p->M) + sizeof(float) * (size_t)(xlen + a.G * (p->M + 1)) + 16;
  dim3 grid((a.T + a.G - 1) / a.G, B);
  // a streaming hop is a single CTA per session: give it more threads; batches keep 256 (several CTAs per SM)
  const int threads = ((long)grid.x * grid.y <= 2 * p->num_sms) ? 512 : 256;
  if (p->M == 320) {
    B2D_SMEM_OPT_IN(smem, stft_kernel<320>);
    B2D_CUDA(launch_pdl(stft_kernel<320>, grid, dim3(threads), smem, st, a));
  } else if (p->M == 512) {
    B2D_SMEM_OPT_IN(smem, stft_kernel<512>);
    B2D_CUDA(launch_pdl(stft_kernel<512>, grid, dim3(threads), smem, st, a));
  } else if (p->M == 768) {
    B2D_SMEM_OPT_IN(smem, stft_kernel<768>);
    B2D_CUDA(launch_pdl(stft_kernel<768>, grid, dim3(threads), smem, st, a));
  } else {
    B2D_SMEM_OPT_IN(smem, stft_kernel<0>);
    B2D_CUDA(launch_pdl(stft_kernel<0>, grid, dim3(threads), smem, st, a));
  }
  B2D_LAUNCH_CHECK("stft_kernel");
  return B2D_OK;
}

int launch_mel_scale(const b2d_plan* p, const float* mag, int B, int T, float* mel, cudaStream_t st) {
  dim3 grid((T + 127) / 128, p->n_mels, B);
  mel_scale_kernel<<<grid, 128, 0, st>>>(mag, p->F, T, p->n_mels, p->d_mel_lo, p->d_mel_cnt, p->d_mel_off, p->d_mel_w, mel);
  B2D_LAUNCH_CHECK("mel_scale_kernel");
  return B2D_OK;
}

int launch_residual(const float* x, const float* pred, float* out, size_t n, int mode, float out_scale, cudaStream_t st) {
  const int blocks = (int)((n + 255) / 256 < 148 * 8 ? (n + 255) / 256 : 148 * 8);
  residual_mel_kernel<<<blocks > 0 ? blocks : 1, 256, 0, st>>>(x, pred, out, n, mode, out_scale);
  B2D_LAUNCH_CHECK("residual_mel_kernel");
  return B2D_OK;
}

int launch_inverse_mel(const b2d_plan* p, const float* mel, int B, int T, float* out, bool torch_layout, cudaStream_t st) {
  const int K = p->n_mels;
  const size_t smem = sizeof(float) * 2 * K * 68;
  if (torch_layout) {
    B2D_SMEM_OPT_IN(smem, inverse_mel_kernel<true>);
    dim3 grid((p->F + 63) / 64, (T + 63) / 64, B);
    inverse_mel_kernel<true><<<grid, 256, smem, st>>>(mel, p->d_pinv, out, B, T, K, p->F, p->Fp);
  } else {
    B2D_SMEM_OPT_IN(smem, inverse_mel_kernel<false>);
    const size_t NF = (size_t)B * T;
    dim3 grid((p->Fp + 63) / 64, (unsigned)((NF + 63) / 64), 1);
    inverse_mel_kernel<false><<<grid, 256, smem, st>>>(mel, p->d_pinv, out, B, T, K, p->F, p->Fp);
  }
  B2D_LAUNCH_CHECK("inverse_mel_kernel");
  return B2D_OK;
}

int launch_istft(const b2d_plan* p, const float2* spec, const float* mag, int B, int T, float* wave, cudaStream_t st) {
  IstftArgs a;
  a.spec = spec; a.mag = mag; a.B = B; a.T = T; a.G = frames_per_block(p) < 3 ? 3 : frames_per_block(p);
  if (p->M > 1024) a.G = 3;
  a.n_fft = p->n_fft; a.hop = p->hop; a.M = p->M; a.F = p->F; a.fd = p->fft;
  a.tw = p->d_tw; a.rtw = p->d_rtw; a.winn = p->d_winn; a.inv_env = p->d_inv_env; a.wave = wave;
  const size_t smem = sizeof(float2) * (size_t)(p->M + 2 * a.G * p->M) + 16;
  B2D_SMEM_OPT_IN(smem, istft_kernel);
  dim3 grid((T - 1 + a.G - 2) / (a.G - 1), B);
  istft_kernel<<<grid, 256, smem, st>>>(a);
  B2D_LAUNCH_CHECK("istft_kernel");
  return B2D_OK;
}

int launch_to_frame_layout(const float* in, float* out, int B, int F, int T, int Fp, cudaStream_t st) {
  dim3 grid((Fp + 31) / 32, (T + 31) / 32, B);
  to_frame_layout_kernel<<<grid, 256, 0, st>>>(in, out, F, T, Fp);
  B2D_LAUNCH_CHECK("to_frame_layout_kernel");
  return B2D_OK;
}

}  // namespace b2d
