// griffinlim.cu -- K6: fused fast Griffin-Lim iteration (TA:functional/functional.py:255-353): the dispatcher (gl_run), the
// generic shared-memory kernel (any n_fft / 2 = 2^a 3^b 5^c, the cross-check engine and the single-launch streaming mode)
// and the stitch kernel.  The register / TMA fast paths live in gl_fast*.cu and gl_warp.cu.
//
// Generic kernel: one launch = one iteration.  The launch reads the current time-domain iterate x_k, and per frame
//   rebuilt = rfft(window * x_k)                      (torch.stft, center/reflect)
//   a       = rebuilt - m * tprev ; a /= |a| + 1e-16  (momentum + projection to unit modulus)
//   tprev   = rebuilt                                 (in place; each frame has one owner)
//   y       = irfft(mag * a) * window                 (torch.istft, first half)
// and overlap-adds y into x_{k+1}.  `angles` is never materialised.  Algorithmic HBM bytes per
// iteration per clip: read x (4 L), read tprev (8 F T), read mag (4 F T), write tprev (8 F T),
// write x (4 L) = 20 F T + 8 L  (SURVEY.md section 8d).
//
// The iterate x is kept in a "partial hop-block" format so that no CTA ever needs another CTA's
// frames and no atomics are used: a clip's T frames are cut into R runs of n frames; run r owns
// n+1 hop-blocks, slot c holding  [second half of frame r*n+c-1] + [first half of frame r*n+c]
// restricted to the frames of that run.  A hop-block on a run boundary is the sum of two slots
// (x_block_sample).  The window-envelope division and the reflect padding of torch.stft are applied
// when the next iteration stages its input.
#include <stdlib.h>

#include "fft.cuh"
#include "kernels.cuh"

namespace b2d {

struct GlArgs {
  const float* mag_tf;    // [B,T,Fp]
  const float2* angles0;  // [B,F,T] torch layout or null (init only)
  float2* tprev;          // [B,T,M]   slot 0 = (Re X[0], Re X[M])
  const float* xin;       // partial hop-block format
  float* xout;            // partial hop-block format
  int B, T, n, R, G;
  int n_fft, hop, M, F, Fp;
  FftDesc fd;
  const float2* tw;
  const float2* rtw;
  const float* win;
  const float* winn;
  const float* inv_env;
  float mom;
  int init, use_prev, store_prev;
  unsigned long long seed;  // init only: angles0 == null and seed != 0 -> in-kernel uniform draws
  const unsigned long long* seed_ptr;  // optional device-resident seed (CUDA-graph replays change it without re-capture)
  // fused mode (R == 1, short clips): the init and all n_iter iterations in ONE launch, ping-ponging xa / xb
  int fused_iters;  // -1: one step per launch (flags above); >= 0: run init + this many iterations here
  float* xa;
  float* xb;
};

// normalised interior hop-block sample, 1 <= j <= T-1
// (`part` is deliberately not __restrict__: in fused mode the same launch wrote it one step earlier, so the loads must
//  stay on the coherent path rather than ld.global.nc)
__device__ __forceinline__ float x_block_sample(const float* part, const float* __restrict__ inv_env,
                                                int b, int R, const FastDiv& dn, int hop, int j, int i) {
  const int n = dn.d;
  const int r1 = dn.div(j - 1), r2 = dn.div(j);
  float v = part[((size_t)(b * R + r1) * (n + 1) + (j - r1 * n)) * hop + i];
  if (r2 != r1) v += part[((size_t)(b * R + r2) * (n + 1)) * hop + i];
  return v * inv_env[i];
}
__device__ __forceinline__ float x_block_sample(const float* part, const float* __restrict__ inv_env,
                                                int b, int R, int n, int hop, int j, int i) {
  return x_block_sample(part, inv_env, b, R, FastDiv(n), hop, j, i);
}
// sample i of padded hop-block j (0 <= j <= T) of the reflect-padded iterate
__device__ __forceinline__ float x_padded(const float* part, const float* __restrict__ inv_env, int b,
                                          int R, const FastDiv& dn, int hop, int T, int j, int i) {
  if (j == 0) return (i == 0) ? x_block_sample(part, inv_env, b, R, dn, hop, 2, 0)
                              : x_block_sample(part, inv_env, b, R, dn, hop, 1, hop - i);
  if (j == T) return (i == hop - 1) ? x_block_sample(part, inv_env, b, R, dn, hop, T - 2, hop - 1)
                                    : x_block_sample(part, inv_env, b, R, dn, hop, T - 1, hop - 2 - i);
  return x_block_sample(part, inv_env, b, R, dn, hop, j, i);
}

// a / (|a| + 1e-16) through the reciprocal-square-root unit (one MUFU instead of an IEEE square root and division): for
// |a| >= 1e-15 the epsilon is below fp32 resolution, for |a| -> 0 both forms tend to a * 1e16 and give 0 for a == 0
__device__ __forceinline__ float inv_abs(float s) {
  float r;
  asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(fmaxf(s, 1e-32f)));
  return r;
}
__device__ __forceinline__ float2 unit_dir(float2 a) {
  const float inv = inv_abs(a.x * a.x + a.y * a.y);
  return make_float2(a.x * inv, a.y * inv);
}

// MT: complex transform length known at compile time (320 / 768: the streaming geometries n_fft 640 / 1536, hop = M) or 0
template <int MT>
__global__ void __launch_bounds__(1024) gl_generic_kernel(const GlArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int M = MT ? MT : a.M, N = MT ? 2 * MT : a.n_fft, G = a.G, hop = MT ? MT : a.hop, n = a.n, R = a.R, T = a.T;
  float2* tw_s = reinterpret_cast<float2*>(smem_raw);
  float2* bufA = tw_s + M;
  float2* bufB = bufA + G * M;
  float* xin_s = reinterpret_cast<float*>(bufB + G * M);  // (G+1)*hop
  float* carry_s = xin_s + (G + 1) * hop;                 // hop

  const int b = blockIdx.y, r = blockIdx.x;
  const int tbeg = r * n, tend = min(T, tbeg + n);
  if (tbeg >= tend) return;
  for (int i = threadIdx.x; i < M; i += blockDim.x) tw_s[i] = a.tw[i];
  for (int i = threadIdx.x; i < hop; i += blockDim.x) carry_s[i] = 0.f;
  __syncthreads();
  const int half = M / 2 + 1;
  const FastDiv d_hop(hop), d_M(M), d_half(half), d_G(G), d_n(n);  // index math without integer division (fft.cuh)
  const unsigned long long seed = a.seed_ptr ? *a.seed_ptr : a.seed;
  const int last_step = a.fused_iters >= 0 ? a.fused_iters : 0;
  const float* cur_in = a.xin;
  float* cur_out = a.xout;
#pragma unroll 1
  for (int step = 0; step <= last_step; ++step) {
  // per-step flags: one step per launch takes them from the arguments; fused mode derives them (step 0 = init)
  int f_init = a.init, f_use_prev = a.use_prev, f_store_prev = a.store_prev;
  if (a.fused_iters >= 0) {
    const int it = step - 1;
    f_init = (step == 0);
    f_use_prev = (it > 0 && a.mom != 0.f);
    f_store_prev = (it + 1 < a.fused_iters && a.mom != 0.f);
    cur_in = (step & 1) ? a.xa : a.xb;   // step 0 writes xa, step 1 reads xa writes xb, ...
    cur_out = (step & 1) ? a.xb : a.xa;
  }
  float* xo = cur_out + (size_t)(b * R + r) * (n + 1) * hop;

  for (int tb = tbeg; tb < tend; tb += G) {
    const int gv = min(G, tend - tb);
    float2 *src, *other;
    if (!f_init) {
      int first = 0;
      if (tb != tbeg) {
        for (int i = threadIdx.x; i < hop; i += blockDim.x) xin_s[i] = xin_s[G * hop + i];
        first = 1;
        __syncthreads();
      }
      for (int idx = threadIdx.x; idx < (gv + 1 - first) * hop; idx += blockDim.x) {
        const int cq = d_hop.div(idx);
        const int c = first + cq, i = idx - cq * hop;
        if (a.fused_iters >= 0) {
          // fused mode: the run is the whole clip (R == 1, n >= T), so no hop-block is split over two runs
          const int j = tb + c;
          int js = j, is = i;
          if (j == 0) { js = (i == 0) ? 2 : 1; is = (i == 0) ? 0 : hop - i; }
          else if (j == T) { js = (i == hop - 1) ? T - 2 : T - 1; is = (i == hop - 1) ? hop - 1 : hop - 2 - i; }
          xin_s[c * hop + i] = cur_in[((size_t)b * (n + 1) + js) * hop + is] * a.inv_env[is];
        } else {
          xin_s[c * hop + i] = x_padded(cur_in, a.inv_env, b, R, d_n, hop, T, tb + c, i);
        }
      }
      __syncthreads();
      for (int idx = threadIdx.x; idx < G * M; idx += blockDim.x) {
        const int g = d_M.div(idx), m = idx - g * M;
        float2 z = make_float2(0.f, 0.f);
        if (g < gv) {
          const float2 xv = *reinterpret_cast<const float2*>(xin_s + g * hop + 2 * m);
          const float2 wv = *reinterpret_cast<const float2*>(a.win + 2 * m);
          z = make_float2(xv.x * wv.x, xv.y * wv.y);
        }
        bufA[idx] = z;
      }
      __syncthreads();
      float2* res = fft_rows_t<false, MT>(bufA, bufB, G, M, a.fd, tw_s);
      for (int idx = threadIdx.x; idx < gv * half; idx += blockDim.x) {
        const int g = d_half.div(idx), k = idx - g * half;
        const int t = tb + g;
        const float2 zk = res[g * M + k];
        const float2 zmk = res[g * M + ((M - k) & -(k != 0))];
        float2 xk, xmk;
        rfft_split(zk, zmk, a.rtw[k], xk, xmk);
        float2* tp = a.tprev + ((size_t)b * T + t) * M;
        const float* mg = a.mag_tf + ((size_t)b * T + t) * a.Fp;
        const float mk = mg[k], mmk = mg[M - k];
        float2 yk, ymk;
        if (k == 0) {
          float a0 = xk.x, aM = xmk.x;
          if (f_use_prev) {
            const float2 pv = tp[0];
            a0 -= a.mom * pv.x;
            aM -= a.mom * pv.y;
          }
          yk = make_float2(mk * (a0 * inv_abs(a0 * a0)), 0.f);
          ymk = make_float2(mmk * (aM * inv_abs(aM * aM)), 0.f);
          if (f_store_prev) tp[0] = make_float2(xk.x, xmk.x);
        } else {
          float2 ak = xk, amk = xmk;
          if (f_use_prev) {
            const float2 pk = tp[k], pmk = tp[M - k];
            ak = make_float2(xk.x - a.mom * pk.x, xk.y - a.mom * pk.y);
            amk = make_float2(xmk.x - a.mom * pmk.x, xmk.y - a.mom * pmk.y);
          }
          const float2 uk = unit_dir(ak), umk = unit_dir(amk);
          yk = make_float2(mk * uk.x, mk * uk.y);
          ymk = make_float2(mmk * umk.x, mmk * umk.y);
          if (f_store_prev) {
            tp[k] = xk;
            if (2 * k != M) tp[M - k] = xmk;
          }
        }
        float2 z1, z2;
        irfft_merge(yk, ymk, a.rtw[k], z1, z2);
        res[g * M + k] = z1;
        if (k != 0 && 2 * k != M) res[g * M + M - k] = z2;
      }
      __syncthreads();
      src = res;
      other = (res == bufA) ? bufB : bufA;
    } else {
      for (int idx = threadIdx.x; idx < G * half; idx += blockDim.x) {
        const int k = d_G.div(idx), g = idx - k * G;
        float2 z1 = make_float2(0.f, 0.f), z2 = z1;
        if (g < gv) {
          const int t = tb + g;
          const float* mg = a.mag_tf + ((size_t)b * T + t) * a.Fp;
          float2 yk = make_float2(mg[k], 0.f), ymk = make_float2(mg[M - k], 0.f);
          if (a.angles0) {
            const float2 ak = a.angles0[((size_t)b * a.F + k) * T + t];
            const float2 amk = a.angles0[((size_t)b * a.F + (M - k)) * T + t];
            yk = make_float2(yk.x * ak.x, yk.x * ak.y);
            ymk = make_float2(ymk.x * amk.x, ymk.x * amk.y);
          } else if (seed) {
            const unsigned long long base = ((unsigned long long)b * T + t) * (M + 1);
            const float2 ak = rand_angle(seed, base + k), amk = rand_angle(seed, base + (M - k));
            yk = make_float2(yk.x * ak.x, yk.x * ak.y);
            ymk = make_float2(ymk.x * amk.x, ymk.x * amk.y);
          }
          if (k == 0) { yk.y = 0.f; ymk.y = 0.f; }
          irfft_merge(yk, ymk, a.rtw[k], z1, z2);
        }
        bufA[g * M + k] = z1;
        if (k != 0 && 2 * k != M) bufA[g * M + M - k] = z2;
      }
      __syncthreads();
      src = bufA;
      other = bufB;
    }
    float2* out = fft_rows_t<true, MT>(src, other, G, M, a.fd, tw_s);
    const float* y = reinterpret_cast<const float*>(out);
    for (int idx = threadIdx.x; idx < gv * hop; idx += blockDim.x) {
      const int c = d_hop.div(idx), i = idx - c * hop;
      const float prev = (c == 0) ? carry_s[i] : y[(c - 1) * N + hop + i] * a.winn[hop + i];
      xo[(size_t)(tb - tbeg + c) * hop + i] = prev + y[c * N + i] * a.winn[i];
      // the thread that consumed carry_s[i] (c == 0, idx == i) is the one that replaces it: no barrier in between
      if (c == 0) carry_s[i] = y[(gv - 1) * N + hop + i] * a.winn[hop + i];
    }
    __syncthreads();
  }
  for (int i = threadIdx.x; i < hop; i += blockDim.x) {
    xo[(size_t)(tend - tbeg) * hop + i] = carry_s[i];
    carry_s[i] = 0.f;  // ready for the next fused step
  }
  __syncthreads();  // fused mode: this CTA's global writes are visible to all its threads before the next step reads them
  }  // step
}

// partial hop-block format -> [B, hop*(T-1)] waveform (envelope-normalised, optional per-clip scale)
__global__ void __launch_bounds__(256) gl_stitch_kernel(const float* __restrict__ part, const float* __restrict__ inv_env,
                                                        const float* __restrict__ out_scale, float* __restrict__ wave,
                                                        int T, int n, int R, int hop, int jstep) {
  const int b = blockIdx.y;
  const int j = (blockIdx.x + 1) * jstep;  // hop-block 1..T-1 (jstep = 1) or the run boundaries n, 2n, ... (jstep = n)
  if (j > T - 1) return;
  const float sc = out_scale ? out_scale[b] : 1.0f;
  float* dst = wave + (size_t)b * hop * (T - 1) + (size_t)(j - 1) * hop;
  for (int i = threadIdx.x; i < hop; i += blockDim.x) dst[i] = x_block_sample(part, inv_env, b, R, n, hop, j, i) * sc;
}

// ------------------------------------------------------------------------------------------------
int launch_gl_fast512(const b2d_plan* p, const float* mag_tf, const float* xin, const float* xprev, float* xout, int B, int T,
                      int n, int R, float mom, int use_prev, float* wave, const float* out_scale, cudaStream_t st);  // gl_fast.cu
int gl_fast_warps_per_sm();
int launch_gl_fast_n512(const b2d_plan* p, const float* mag_tf, float2* tprev, const float* xin, float* xout, int B, int T, int n,
                        int R, float mom, int use_prev, int store_prev, cudaStream_t st);  // gl_fast_n512.cu
int launch_gl_fast_n2048(const b2d_plan* p, const float* mag_tf, float2* tprev, const float* xin, float* xout, int B, int T, int n,
                         int R, float mom, int use_prev, int store_prev, float* wave, const float* out_scale,
                         cudaStream_t st);  // gl_fast_n2048.cu
int gl_fast_n2048_warps();
int launch_gl_fast_n2048_init(const b2d_plan* p, const float* mag_tf, float* xout, int B, int T, int n, int R, unsigned long long seed,
                              const unsigned long long* seed_ptr, cudaStream_t st);
int launch_gl_warp(const b2d_plan* p, const float* mag_tf, float2* tprev, const float* xin, float* xout, int B, int T, int n, int R,
                   float mom, int use_prev, int store_prev, cudaStream_t st);  // gl_warp.cu
int gl_warp_warps(const b2d_plan* p);
bool gl_warp_supported(const b2d_plan* p);
int launch_gl_fast512_init(const b2d_plan* p, const float* mag_tf, float* xout, int B, int T, int n, int R, unsigned long long seed,
                           const unsigned long long* seed_ptr, cudaStream_t st);
int launch_gl_reg(const b2d_plan* p, const float* mag_tf, const float* xin, const float* xprev, float* xout, int B, int T, int n,
                  int R, float mom, int use_prev, float* wave, const float* out_scale, cudaStream_t st);  // gl_reg.cu
int launch_gl_reg_init(const b2d_plan* p, const float* mag_tf, float* xout, int B, int T, int n, int R, unsigned long long seed,
                       const unsigned long long* seed_ptr, cudaStream_t st);
bool gl_reg_fused_plan(const b2d_plan* p, int B, int T, int* n_out, int* R_out, int* csize_out);
int launch_gl_reg_fused(const b2d_plan* p, const float* mag_tf, const float2* angles0, unsigned long long seed,
                        const unsigned long long* seed_ptr, float* x0, float* x1, float* x2, int B, int T, int n, int R, int csize,
                        int n_iter, float mom, const float* out_scale, float* wave, cudaStream_t st);
int gl_reg_r3(const b2d_plan* p);
int gl_reg_warps(const b2d_plan* p);
bool gl_reg_hop_plan(const b2d_plan* p, int B, int T);
bool gl_reg_coop_plan(const b2d_plan* p, int B, int T);
int launch_gl_reg_coop(const b2d_plan* p, const float* mag_tf, const float2* angles0, unsigned long long seed,
                       const unsigned long long* seed_ptr, float* x0, float* x1, float* x2, int B, int T, int n_iter,
                       float mom, const float* out_scale, float* wave, cudaStream_t st);
int launch_gl_reg_hop(const b2d_plan* p, const float* mag_tf, const float2* angles0, unsigned long long seed,
                      const unsigned long long* seed_ptr, int B, int T, int n_iter, float mom, const float* out_scale, float* wave,
                      cudaStream_t st, float* ola, float* hop_out);

// uniform cut for the generic shared-memory kernel: one CTA per run, runs a multiple of G frames long
static GlPartition generic_partition(const b2d_plan* p, int B, int T) {
  GlPartition q;
  q.G = (p->M <= 1024) ? 4 : 2;
  q.fast = 0;
  const int target = 4 * p->num_sms;  // CTAs wanted in flight
  int R = (target + B - 1) / B;
  const int maxR = (T + q.G - 1) / q.G;
  if (R > maxR) R = maxR;
  if (R < 1) R = 1;
  int n = (T + R - 1) / R;
  n = (n + q.G - 1) / q.G * q.G;
  q.n = n;
  q.R = (T + n - 1) / n;
  return q;
}

GlPartition gl_partition(const b2d_plan* p, int B, int T) {
  GlPartition q;
  q.G = (p->M <= 1024) ? 4 : 2;
  q.fast = 0;
  q.csize = 1;
  // small problems (a streaming hop, a handful of clips): init + all iterations in ONE launch, a cluster per clip (gl_reg.cu)
  // one cooperative launch over the whole GPU, a frame per warp group (gl_reg.cu); B2D_PLAN_CLUSTER_GL keeps the cluster kernel
  if (!(p->flags & B2D_PLAN_CLUSTER_GL) && gl_reg_coop_plan(p, B, T)) { q.fast = 8; q.n = 1; q.R = T; return q; }
  if (gl_reg_fused_plan(p, B, T, &q.n, &q.R, &q.csize)) { q.fast = 6; return q; }
  // a streaming hop (T = 3 frames): the same in one CTA per session with the iterates in shared memory (gl_reg.cu)
  if (gl_reg_hop_plan(p, B, T)) { q.fast = 7; q.n = 1; q.R = T; return q; }
  // a clip of at most one group of frames (a streaming hop: T = 3) runs init + all iterations in ONE launch of the generic
  // kernel, one CTA per clip: faster than 33 launches of the register kernels (n_fft 1024: 0.41 -> 0.3 ms per hop)
  if (T <= q.G) return generic_partition(p, B, T);
  if (!(p->flags & B2D_PLAN_GENERIC_KERNELS)) {
    if (p->n_fft == 1024 && p->hop == 512) q.fast = 1;
    if (p->n_fft == 512 && p->hop == 256) q.fast = 2;
    if (p->n_fft == 2048 && p->hop == 1024) q.fast = 3;
    if (gl_reg_r3(p)) q.fast = 5;  // n_fft 640 / 1536: generic-radix register FFT (gl_reg.cu)
    // every other length with hop = n_fft / 2: warp-synchronous Stockham kernel (gl_warp.cu)
    if (!q.fast && gl_warp_supported(p)) q.fast = 4;
  }
  if (q.fast) {
    // One warp per run, persistent CTAs.  Measured: a warp needs about the same time per frame whether 1 or 12 warps share
    // the SM (the kernels are latency-bound per warp), so a launch lasts (rounds of runs per warp slot) x (frames per run).
    // Pick the number of runs per clip that minimises that; ties -> longer runs (less boundary traffic).
    // (Measured dead ends, round 1: dealing the runs longest-first in serpentine order, and a balanced run table of equal
    //  shares split at clip boundaries -- both correct, both 4-10 % slower: 1536 busy warps already saturate the SMs' issue /
    //  shared-memory throughput, more runs only add boundary traffic.)
    const int wps = q.fast == 3 ? gl_fast_n2048_warps() : q.fast == 4 ? gl_warp_warps(p) : q.fast == 5 ? gl_reg_warps(p) : gl_fast_warps_per_sm();
    const long slots = (long)wps * p->num_sms;
    const int min_run = 2;  // frames per run: short runs cut the latency of small batches
    const int maxR = (T + min_run - 1) / min_run;
    long best = -1; int bestR = 1;
    int bestN = T;
    for (int R = 1; R <= maxR; ++R) {
      int n = (T + R - 1) / R;
      if (q.fast == 2) n = (n + 1) & ~1;  // n_fft 512 walks a run two frames at a time
      const int Reff = (T + n - 1) / n;
      const long runs = (long)B * Reff;
      const long rounds = (runs + slots - 1) / slots;
      const long cost = rounds * ((q.fast == 2 ? n / 2 : n) + 1);  // + 1: a run's fixed cost (first rows, boundary blocks, carry flush)
      if (best < 0 || cost < best) { best = cost; bestR = Reff; bestN = n; }
    }
    q.n = bestN;
    q.R = bestR;
    return q;
  }
  return generic_partition(p, B, T);
}

static size_t part_floats(const b2d_plan* p, const GlPartition& q, int B) {
  return (size_t)B * q.R * (q.n + 1) * p->hop;
}

// workspace: the time-domain-momentum fast paths (n_fft 1024; 640 / 1536) keep three time-domain iterates (x_{k-1}, x_k, x_{k+1}) and no spectrogram state;
// the other kernels keep two iterates and the complex `tprev` [B, T, M]
size_t gl_workspace_bytes(const b2d_plan* p, int B, int T, bool need_mag_copy) {
  const GlPartition q = gl_partition(p, B, T);
  const size_t pbytes = align_up(part_floats(p, q, B) * sizeof(float), 256);
  size_t bytes = (q.fast == 7) ? 256 : (q.fast == 1 || q.fast == 5 || q.fast == 6 || q.fast == 8) ? 3 * pbytes : 2 * pbytes + align_up((size_t)B * T * p->M * sizeof(float2), 256);
  if (need_mag_copy) bytes += align_up((size_t)B * T * p->Fp * sizeof(float), 256);
  return bytes;
}

bool gl_fuses_ola(const b2d_plan* p, int B, int T) { return T == 3 && gl_partition(p, B, T).fast == 7; }

int gl_run(const b2d_plan* p, const float* mag_tf, const float2* init_angles, unsigned long long seed, int B, int T, int n_iter,
           float momentum, const float* out_scale, float* wave, void* ws, size_t ws_bytes, cudaStream_t st,
           const unsigned long long* seed_ptr, float* ola, float* hop_out) {
  B2D_REQUIRE(p->hop * 2 == p->n_fft, B2D_ERR_UNSUPPORTED, "Griffin-Lim requires hop == n_fft/2 (got n_fft=%d hop=%d)", p->n_fft, p->hop);
  B2D_REQUIRE(T >= 3, B2D_ERR_BAD_ARG, "Griffin-Lim needs at least 3 frames (got %d)", T);
  B2D_REQUIRE(momentum >= 0.f && momentum < 1.f, B2D_ERR_BAD_ARG, "momentum must be in range [0, 1). Found: %g", (double)momentum);
  B2D_REQUIRE(n_iter >= 0, B2D_ERR_BAD_ARG, "n_iter must be >= 0");
  B2D_REQUIRE(B >= 1 && B <= 65535, B2D_ERR_BAD_ARG, "batch must be in [1, 65535] per call (got %d)", B);
  B2D_REQUIRE(ws_bytes >= gl_workspace_bytes(p, B, T, false), B2D_ERR_WORKSPACE, "Griffin-Lim workspace too small");
  const GlPartition q = gl_partition(p, B, T);
  const size_t pbytes = align_up(part_floats(p, q, B) * sizeof(float), 256);
  unsigned char* base = static_cast<unsigned char*>(ws);
  float* xa = reinterpret_cast<float*>(base);
  float* xb = reinterpret_cast<float*>(base + pbytes);
  float* xc = reinterpret_cast<float*>(base + 2 * pbytes);       // fast == 1 / 5: third iterate
  float2* tprev = reinterpret_cast<float2*>(base + 2 * pbytes);  // otherwise: complex spectrogram state

  GlArgs a;
  a.mag_tf = mag_tf; a.angles0 = init_angles; a.tprev = tprev;
  a.B = B; a.T = T; a.n = q.n; a.R = q.R; a.G = q.G;
  a.n_fft = p->n_fft; a.hop = p->hop; a.M = p->M; a.F = p->F; a.Fp = p->Fp; a.fd = p->fft;
  a.tw = p->d_tw; a.rtw = p->d_rtw; a.win = p->d_win; a.winn = p->d_winn; a.inv_env = p->d_inv_env;
  a.mom = momentum / (1.0f + momentum);
  const size_t smem = sizeof(float2) * (size_t)(p->M + 2 * q.G * p->M) + sizeof(float) * (size_t)((q.G + 2) * p->hop) + 16;
  const int mt = (p->hop != p->M) ? 0 : (p->M == 320 || p->M == 512 || p->M == 768) ? p->M : 0;
  auto launch_generic = [&](dim3 g, const GlArgs& args) -> int {
    int threads = p->M >= 640 ? 512 : 256;  // batch mode, measured at B = 256: n_fft 1536 1259 -> 1128 us per iteration with 512
    // fused mode: one CTA walks all the steps of a short clip alone, give it the whole SM (measured per-hop latency, n_fft 640 /
    // 1536: 0.222 / 0.579 ms with 256 threads, 0.207 / 0.371 with 512, 0.211 / 0.357 with 1024)
    if (args.fused_iters >= 0) threads = p->M >= 640 ? 1024 : (p->M >= 256 ? 512 : 256);
    if (mt == 320) { B2D_SMEM_OPT_IN(smem, gl_generic_kernel<320>); gl_generic_kernel<320><<<g, threads, smem, st>>>(args); }
    else if (mt == 768) { B2D_SMEM_OPT_IN(smem, gl_generic_kernel<768>); gl_generic_kernel<768><<<g, threads, smem, st>>>(args); }
    else if (mt == 512) { B2D_SMEM_OPT_IN(smem, gl_generic_kernel<512>); gl_generic_kernel<512><<<g, threads, smem, st>>>(args); }
    else { B2D_SMEM_OPT_IN(smem, gl_generic_kernel<0>); gl_generic_kernel<0><<<g, threads, smem, st>>>(args); }
    return B2D_OK;
  };
  dim3 grid(q.R, B);
  a.fused_iters = -1; a.xa = xa; a.xb = xb;
  float* cur = xa;
  float* nxt = xb;
  float* prv = xc;
  bool direct_interior = false;
  int rc;
  if (q.fast == 7) {
    return launch_gl_reg_hop(p, mag_tf, init_angles, seed, seed_ptr, B, T, n_iter, a.mom, out_scale, wave, st, ola, hop_out);
  }
  if (q.fast == 8) {
    return launch_gl_reg_coop(p, mag_tf, init_angles, seed, seed_ptr, xa, xb, xc, B, T, n_iter, a.mom, out_scale, wave, st);
  }
  if (q.fast == 6) {
    return launch_gl_reg_fused(p, mag_tf, init_angles, seed, seed_ptr, xa, xb, xc, B, T, q.n, q.R, q.csize, n_iter, a.mom, out_scale, wave, st);
  }
  if (!q.fast && q.R == 1 && T <= 16) {
    // short clips (streaming hops: T = 3): every dependency stays inside one CTA -> init + all iterations in one launch
    a.init = 1; a.use_prev = 0; a.store_prev = 0; a.xin = nullptr; a.xout = xa; a.seed = seed; a.seed_ptr = seed_ptr;
    a.fused_iters = n_iter;
    if ((rc = launch_generic(grid, a))) return rc;
    B2D_LAUNCH_CHECK("gl_generic_kernel(fused)");
    cur = (n_iter & 1) ? xb : xa;  // step s writes xa when s is even; the last step is s = n_iter
  } else {
  // x_0 = istft(mag * angles_0)
  a.init = 1; a.use_prev = 0; a.store_prev = 0; a.xin = nullptr; a.xout = xa; a.seed = seed; a.seed_ptr = seed_ptr;
  if (q.fast == 1 && init_angles == nullptr) {
    if ((rc = launch_gl_fast512_init(p, mag_tf, xa, B, T, q.n, q.R, seed, seed_ptr, st))) return rc;
  } else if (q.fast == 3 && init_angles == nullptr) {
    if ((rc = launch_gl_fast_n2048_init(p, mag_tf, xa, B, T, q.n, q.R, seed, seed_ptr, st))) return rc;
  } else if (q.fast == 5 && init_angles == nullptr) {
    if ((rc = launch_gl_reg_init(p, mag_tf, xa, B, T, q.n, q.R, seed, seed_ptr, st))) return rc;
  } else {
    if ((rc = launch_generic(grid, a))) return rc;
    B2D_LAUNCH_CHECK("gl_generic_kernel(init)");
  }
  a.init = 0; a.angles0 = nullptr;
  for (int it = 0; it < n_iter; ++it) {
    a.use_prev = (it > 0 && a.mom != 0.f) ? 1 : 0;
    a.store_prev = (it + 1 < n_iter && a.mom != 0.f) ? 1 : 0;
    a.xin = cur; a.xout = nxt;
    const bool last = (it + 1 == n_iter);  // last iteration: run-interior hop-blocks go straight to `wave`
    if (q.fast == 1 || q.fast == 5) {
      // time-domain momentum: reads x_k (cur) and x_{k-1} (prv), writes x_{k+1} (nxt); three buffers rotate
      if (q.fast == 1) rc = launch_gl_fast512(p, mag_tf, cur, prv, nxt, B, T, q.n, q.R, a.mom, a.use_prev, last ? wave : nullptr, out_scale, st);
      else rc = launch_gl_reg(p, mag_tf, cur, prv, nxt, B, T, q.n, q.R, a.mom, a.use_prev, last ? wave : nullptr, out_scale, st);
      if (rc) return rc;
      direct_interior = last;
      float* t = prv; prv = cur; cur = nxt; nxt = t;
      continue;
    } else if (q.fast == 2) {
      if ((rc = launch_gl_fast_n512(p, mag_tf, tprev, cur, nxt, B, T, q.n, q.R, a.mom, a.use_prev, a.store_prev, st))) return rc;
    } else if (q.fast == 4) {
      if ((rc = launch_gl_warp(p, mag_tf, tprev, cur, nxt, B, T, q.n, q.R, a.mom, a.use_prev, a.store_prev, st))) return rc;
    } else if (q.fast == 3) {
      if ((rc = launch_gl_fast_n2048(p, mag_tf, tprev, cur, nxt, B, T, q.n, q.R, a.mom, a.use_prev, a.store_prev,
                                     last ? wave : nullptr, out_scale, st))) return rc;
      direct_interior = last;
    } else {
      if ((rc = launch_generic(grid, a))) return rc;
      B2D_LAUNCH_CHECK("gl_generic_kernel");
    }
    float* t = cur; cur = nxt; nxt = t;
  }
  }
  if (direct_interior) {
    // only the hop-blocks on run boundaries (j = r * n, r = 1 .. R-1) still need the sum of two partial slots
    if (q.R > 1) {
      gl_stitch_kernel<<<dim3(q.R - 1, B), 256, 0, st>>>(cur, p->d_inv_env, out_scale, wave, T, q.n, q.R, p->hop, q.n);
      B2D_LAUNCH_CHECK("gl_stitch_kernel(boundaries)");
    }
  } else {
    gl_stitch_kernel<<<dim3(T - 1, B), 256, 0, st>>>(cur, p->d_inv_env, out_scale, wave, T, q.n, q.R, p->hop, 1);
    B2D_LAUNCH_CHECK("gl_stitch_kernel");
  }
  return B2D_OK;
}

}  // namespace b2d
