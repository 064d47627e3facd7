// gl_reg.cu -- Griffin-Lim iteration and fused STFT + Mel for n_fft = 128 * R3 with a register-resident FFT (gl_reg.cuh):
// n_fft 640 (R3 = 5: the 20 ms hop of BASELINE config 3 in batch mode) and n_fft 1536 (R3 = 12: the reference app's own
// 48 kHz geometry, app3.py:29-33).  Same structure as the n_fft = 1024 kernel (gl_fast.cu): one warp owns a frame, walks a
// run of frames with the overlap-add carry in registers, the momentum term is formed on the time-domain iterates
// (d = x_k - m x_{k-1}, see gl_fast.cu), the mag row and the hop-blocks of both iterates arrive by TMA one frame ahead, one
// persistent CTA per SM.  Round 1 ran these lengths on shared-memory Stockham kernels at 19 % / 25 % of HBM peak.
#include <cooperative_groups.h>

#include "gl_reg.cuh"
#include "kernels.cuh"
#include "tma.cuh"

namespace b2d {

using namespace regfft;

struct GlRegArgs {
  const float* mag_tf;   // [B,T,Fp]
  const float* xin;      // x_k, partial hop-block format
  const float* xprev;    // x_{k-1} (read only when USE_PREV)
  float* xout;           // x_{k+1}
  int B, T, n, R, Fp;
  const float2* tw;      // W_M^k
  const float2* rtw;     // W_N^k, k < M
  const float* win;      // [N]
  const float* winn;     // [N] win / N
  const float* inv_env;  // [HOP]
  float mom;
  float* wave;             // last iteration: [B, HOP*(T-1)] or null
  const float* out_scale;  // [B] or null
  unsigned long long seed;             // MODE_INIT: 0 = all-ones angles, else in-kernel counter-based U[0,1)^2 draws
  const unsigned long long* seed_ptr;  // optional device-resident seed
};
// kernel modes: one iteration without / with the momentum term, or x_0 = istft(mag * angles_0) (the inverse half only)
constexpr int MODE_FIRST = 0, MODE_ITER = 1, MODE_INIT = 2;

template <int R3>
struct RegSmem {
  typedef Geo<R3> G;
  static constexpr int MAG_BYTES = (G::M + 4) * 4;                       // Fp floats, a multiple of 16
  static constexpr int OFF_MAG = G::XCH * 8;
  static constexpr int OFF_BAR = OFF_MAG + ((MAG_BYTES + 15) & ~15);
  static constexpr int OFF_RING = OFF_BAR + 16;
  static constexpr int OFF_XBAR = OFF_RING + 2 * 2 * G::HOP * 4;         // two slots x [x_k block | x_{k-1} block]
  static constexpr int WARP_BYTES = OFF_XBAR + 16;
  static constexpr int TABLE_BYTES = (G::M / 2 + G::M / 2 + G::M + G::M) * 8;  // WA | WB | WN | RT (float2)
};

// sample `is` of interior hop-block js of clip b in the partial format: the sum of two slots on a run boundary
// (the loads go to L2 with ld.global.cg: in the single-launch kernel `part` was written by other CTAs of the cluster a step ago)
template <int HOP>
__device__ __forceinline__ float reg_partial_sample(const float* part, int b, int R, int n, int js, int is) {
  const int r1 = (js - 1) / n, r2 = js / n;
  float v = __ldcg(part + ((size_t)(b * R + r1) * (n + 1) + (js - r1 * n)) * HOP + is);
  if (r2 != r1) v += __ldcg(part + ((size_t)(b * R + r2) * (n + 1)) * HOP + is);
  return v;
}
template <int R3, int WARPS, int MODE>
__global__ void __launch_bounds__(WARPS * 32, 1) gl_reg_kernel(const GlRegArgs a) {
  constexpr bool USE_PREV = (MODE == MODE_ITER);
  constexpr bool INIT = (MODE == MODE_INIT);
  typedef Geo<R3> G;
  typedef RegSmem<R3> SM;
  constexpr int M = G::M, HOP = G::HOP, NB = G::NB, NR = G::NR, H2 = M / 2;  // H2: float2 per hop-block
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float2* WA = reinterpret_cast<float2*>(smem_raw);  // inv_env * win, first half   [H2]
  float2* WB = WA + H2;                              // inv_env * win, second half  [H2]
  float2* WN = WB + H2;                              // win / N                     [M]
  float2* RT = WN + M;                               // W_N^k                       [M]
  unsigned char* warp_base = reinterpret_cast<unsigned char*>(RT + M);
  for (int i = threadIdx.x; i < H2; i += blockDim.x) {
    WA[i] = make_float2(a.inv_env[2 * i] * a.win[2 * i], a.inv_env[2 * i + 1] * a.win[2 * i + 1]);
    WB[i] = make_float2(a.inv_env[2 * i] * a.win[HOP + 2 * i], a.inv_env[2 * i + 1] * a.win[HOP + 2 * i + 1]);
  }
  for (int i = threadIdx.x; i < M; i += blockDim.x) {
    WN[i] = make_float2(a.winn[2 * i], a.winn[2 * i + 1]);
    RT[i] = a.rtw[i];
  }
  __syncthreads();

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n = a.n, R = a.R, T = a.T;
  const int nruns = a.B * a.R;
  const int gw0 = warp * (int)gridDim.x + (int)blockIdx.x;  // runs dealt round-robin over the CTAs
  const int gstep = WARPS * (int)gridDim.x;
  if (gw0 >= nruns) return;
  unsigned char* wsm = warp_base + (size_t)warp * SM::WARP_BYTES;
  float2* S = reinterpret_cast<float2*>(wsm);
  float* mg_s = reinterpret_cast<float*>(wsm + SM::OFF_MAG);
  uint64_t* bar = reinterpret_cast<uint64_t*>(wsm + SM::OFF_BAR);
  float* xs = reinterpret_cast<float*>(wsm + SM::OFF_RING);  // slot s: x_k block at xs + s * 2 * HOP, x_{k-1} block behind it
  uint64_t* xbar = reinterpret_cast<uint64_t*>(wsm + SM::OFF_XBAR);
  if (lane == 0) {
    tma::barrier_init(bar, 1);
    tma::barrier_init(xbar, 1);
    tma::barrier_init(xbar + 1, 1);
    tma::fence_barrier_init();
  }
  __syncwarp();
  LaneTwR<R3> tw;
  lane_twiddles_r<R3>(lane, a.tw, tw);
  const size_t run_stride = (size_t)(n + 1) * HOP;
  const float2 nmom = make_float2(-a.mom, -a.mom);
  constexpr uint32_t ring_bytes = (USE_PREV ? 2u : 1u) * HOP * 4u;
  uint32_t tma_uses = 0, xuse0 = 0, xuse1 = 0;
  pdl_wait();  // the prologue above read plan tables only (common.cuh: programmatic dependent launch)
  pdl_trigger();

#pragma unroll 1
  for (int gw = gw0; gw < nruns; gw += gstep) {
    const int b = gw / R, r = gw - b * R;
    const int tb = r * n, te = min(T, tb + n);
    const float* xrun = a.xin + (size_t)(b * R + r) * run_stride;
    const float* prun = a.xprev + (size_t)(b * R + r) * run_stride;
    float* xo = a.xout + (size_t)(b * R + r) * run_stride;
    float2 carry[NR * 4];
#pragma unroll
    for (int q = 0; q < NR * 4; ++q) carry[q] = make_float2(0.f, 0.f);
    if (lane == 0) {
      tma::expect_bytes(bar, SM::MAG_BYTES);
      tma::load(mg_s, a.mag_tf + ((size_t)b * T + tb) * a.Fp, SM::MAG_BYTES, bar);
    }
    const int nrun = te - tb;
    if (!INIT && lane == 0 && nrun > 1) {
      tma::expect_bytes(xbar + 1, ring_bytes);
      tma::load(xs + 2 * HOP, xrun + HOP, HOP * 4, xbar + 1);
      if (USE_PREV) tma::load(xs + 3 * HOP, prun + HOP, HOP * 4, xbar + 1);
    }
#pragma unroll 1
    for (int t = tb; t < te; ++t) {
      const int c = t - tb;
      float2 v[G::NV];
      float2 wA[R3], wB[R3];
      if (!INIT) {
      // ---- stage the frame: v[8 r + n1] = d[(lane + 32 r) + NB n1], n1 < 4 from hop-block t, n1 >= 4 from hop-block t + 1 ----
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int j = t + h, cs = c + h;
        const float2* wtab = h ? WB : WA;
        if (j == 0 || j == T) {  // reflect-padded edge of the clip
          __syncwarp();
          stage_reflect_wide<HOP>(a.xin, USE_PREV ? a.xprev : nullptr, a.mom, a.inv_env, a.win + h * HOP, b, R, n, T, j,
                                 reinterpret_cast<float*>(S), lane);
          __syncwarp();
#pragma unroll
          for (int rr = 0; rr < NR; ++rr)
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const int i = lane + 32 * rr;
              v[8 * rr + 4 * h + q] = (G::FULL || i < NB) ? S[i + NB * q] : make_float2(0.f, 0.f);
            }
        } else if (cs >= 1 && cs <= nrun - 1) {  // interior block of this run: in the shared-memory ring
          if (h == 1) {
            if (cs & 1) { tma::wait(xbar + 1, xuse1 & 1); ++xuse1; } else { tma::wait(xbar, xuse0 & 1); ++xuse0; }
          }
          const float2* src = reinterpret_cast<const float2*>(xs + (cs & 1) * 2 * HOP);
#pragma unroll
          for (int rr = 0; rr < NR; ++rr)
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const int i = lane + 32 * rr;
              float2 xv = make_float2(0.f, 0.f);
              if (G::FULL || i < NB) {
                xv = src[i + NB * q];
                if (USE_PREV) xv = cfma2(src[H2 + i + NB * q], nmom, xv);
                xv = cscale2(xv, wtab[i + NB * q]);
              }
              v[8 * rr + 4 * h + q] = xv;
            }
        } else {  // block on a run boundary: the sum of this run's slot and the neighbouring run's slot
          const size_t o1 = (size_t)cs * HOP;
          const ptrdiff_t o2 = (cs == 0) ? -(ptrdiff_t)run_stride + (ptrdiff_t)n * HOP : (ptrdiff_t)run_stride;
          const bool two = (cs == 0) || (j == te);
          const float2* p1 = reinterpret_cast<const float2*>(xrun + o1);
          const float2* q1 = reinterpret_cast<const float2*>(prun + o1);
#pragma unroll
          for (int rr = 0; rr < NR; ++rr)
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const int i = lane + 32 * rr;
              float2 xv = make_float2(0.f, 0.f);
              if (G::FULL || i < NB) {
                const int m = i + NB * q;
                xv = p1[m];
                if (two) xv = cadd(xv, reinterpret_cast<const float2*>(xrun + o2)[m]);
                if (USE_PREV) {
                  float2 pv = q1[m];
                  if (two) pv = cadd(pv, reinterpret_cast<const float2*>(prun + o2)[m]);
                  xv = cfma2(pv, nmom, xv);
                }
                xv = cscale2(xv, wtab[m]);
              }
              v[8 * rr + 4 * h + q] = xv;
            }
        }
      }
      {  // block c is consumed: its ring slot takes block c + 2 while this frame computes
        __syncwarp();
        if (lane == 0 && c + 2 <= nrun - 1) {
          float* slot = xs + (c & 1) * 2 * HOP;
          tma::expect_bytes(xbar + (c & 1), ring_bytes);
          tma::load(slot, xrun + (size_t)(c + 2) * HOP, HOP * 4, xbar + (c & 1));
          if (USE_PREV) tma::load(slot + HOP, prun + (size_t)(c + 2) * HOP, HOP * 4, xbar + (c & 1));
        }
      }
      // ---- forward FFT ----
      __syncwarp();
      fwd1_store_r<R3>(lane, v, tw, S);
      __syncwarp();
      fwd2_load_r<R3>(lane, v, S);
      __syncwarp();
      fwd2_store_r<R3>(lane, v, tw, S);
      __syncwarp();
      fwd3_load_r<R3>(lane, wA, wB, S);
      }
      // ---- projection to unit modulus, x mag (INIT: mag x initial angles) ----
      tma::wait(bar, tma_uses & 1);
      ++tma_uses;
      if (INIT) {
        const unsigned long long seed = a.seed_ptr ? *a.seed_ptr : a.seed;
        init_frame<R3>(lane, wA, wB, RT, mg_s, seed, ((unsigned long long)b * T + t) * (M + 1));
      } else {
        project_frame<R3>(lane, wA, wB, RT, mg_s);
      }
      __syncwarp();  // every lane is done reading the staged row
      if (lane == 0 && t + 1 < te) {
        tma::expect_bytes(bar, SM::MAG_BYTES);
        tma::load(mg_s, a.mag_tf + ((size_t)b * T + t + 1) * a.Fp, SM::MAG_BYTES, bar);
      }
      // ---- inverse FFT ----
      inv1_store_r<R3>(lane, wA, wB, S);
      __syncwarp();
      inv2_load_r<R3>(lane, v, tw, S);
      __syncwarp();
      inv2_store_r<R3>(lane, v, S);
      __syncwarp();
      inv3_load_r<R3>(lane, v, tw, S);
      // ---- synthesis window + overlap-add: block c = carry + first half ; carry = second half ----
      const bool direct = (a.wave != nullptr && c >= 1);  // last iteration, block interior to this run: final samples
      const float sc = (direct && a.out_scale) ? a.out_scale[b] : 1.0f;
      float2* dst = direct ? reinterpret_cast<float2*>(a.wave + (size_t)b * HOP * (T - 1) + (size_t)(t - 1) * HOP)
                           : reinterpret_cast<float2*>(xo + (size_t)c * HOP);
      const float2* ie = reinterpret_cast<const float2*>(a.inv_env);
#pragma unroll
      for (int rr = 0; rr < NR; ++rr)
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int i = lane + 32 * rr;
          if (G::FULL || i < NB) {
            const int m = i + NB * q;
            float2 o = cfma2(v[8 * rr + q], WN[m], carry[4 * rr + q]);
            if (direct) o = cscale2(o, cscale2(ie[m], make_float2(sc, sc)));
            dst[m] = o;
            carry[4 * rr + q] = cscale2(v[8 * rr + 4 + q], WN[H2 + m]);
          }
        }
    }
    float2* dl = reinterpret_cast<float2*>(xo + (size_t)(te - tb) * HOP);
#pragma unroll
    for (int rr = 0; rr < NR; ++rr)
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int i = lane + 32 * rr;
        if (G::FULL || i < NB) dl[i + NB * q] = carry[4 * rr + q];
      }
    __syncwarp();
  }
}

// warps per CTA: as many as registers (65536 / warps / 32 per thread) and shared memory allow
template <int R3> struct RegWarps;
template <> struct RegWarps<5> { static constexpr int W = 16; };    // 128 registers / thread
template <> struct RegWarps<12> { static constexpr int W = 9; };    // 22.3 KB of staging per warp

int gl_reg_r3(const b2d_plan* p) {
  if (p->hop * 2 != p->n_fft) return 0;
  if (p->n_fft == 640) return 5;
  if (p->n_fft == 1536) return 12;
  return 0;
}
int gl_reg_warps(const b2d_plan* p) {
  const int r3 = gl_reg_r3(p);
  return r3 == 5 ? RegWarps<5>::W : r3 == 12 ? RegWarps<12>::W : 0;
}

template <int R3, int MODE>
static int launch_reg(const GlRegArgs& a, int num_sms, cudaStream_t st) {
  constexpr int W = RegWarps<R3>::W;
  const size_t smem = (size_t)RegSmem<R3>::TABLE_BYTES + (size_t)W * RegSmem<R3>::WARP_BYTES;
  static_assert(RegSmem<R3>::TABLE_BYTES + W * RegSmem<R3>::WARP_BYTES <= 232448, "per-CTA shared memory exceeds 227 KB");
  B2D_SMEM_OPT_IN(smem, gl_reg_kernel<R3, W, MODE>);
  const int runs = a.B * a.R;
  B2D_CUDA(launch_pdl(gl_reg_kernel<R3, W, MODE>, dim3(runs < num_sms ? runs : num_sms), dim3(W * 32), smem, st, a));
  B2D_LAUNCH_CHECK("gl_reg_kernel");
  return B2D_OK;
}
template <int MODE>
static int launch_reg_r3(const b2d_plan* p, const GlRegArgs& a, cudaStream_t st) {
  switch (gl_reg_r3(p)) {
    case 5: return launch_reg<5, MODE>(a, p->num_sms, st);
    case 12: return launch_reg<12, MODE>(a, p->num_sms, st);
  }
  return fail(B2D_ERR_UNSUPPORTED, "no register-FFT Griffin-Lim kernel for n_fft = %d", p->n_fft);
}
static GlRegArgs reg_args(const b2d_plan* p, const float* mag_tf, int B, int T, int n, int R) {
  GlRegArgs a{};
  a.mag_tf = mag_tf; a.B = B; a.T = T; a.n = n; a.R = R; a.Fp = p->Fp;
  a.tw = p->d_tw; a.rtw = p->d_rtw; a.win = p->d_win; a.winn = p->d_winn; a.inv_env = p->d_inv_env;
  return a;
}

int launch_gl_reg(const b2d_plan* p, const float* mag_tf, const float* xin, const float* xprev, float* xout, int B, int T, int n,
                  int R, float mom, int use_prev, float* wave, const float* out_scale, cudaStream_t st) {
  GlRegArgs a = reg_args(p, mag_tf, B, T, n, R);
  a.xin = xin; a.xprev = use_prev ? xprev : xin; a.xout = xout;
  a.mom = mom; a.wave = wave; a.out_scale = out_scale;
  return use_prev ? launch_reg_r3<MODE_ITER>(p, a, st) : launch_reg_r3<MODE_FIRST>(p, a, st);
}
// x_0 = istft(mag * angles_0), angles_0 all ones (seed 0) or drawn in-kernel (same draws as the generic kernel for the same seed)
int launch_gl_reg_init(const b2d_plan* p, const float* mag_tf, float* xout, int B, int T, int n, int R, unsigned long long seed,
                       const unsigned long long* seed_ptr, cudaStream_t st) {
  GlRegArgs a = reg_args(p, mag_tf, B, T, n, R);
  a.xin = xout; a.xprev = xout; a.xout = xout; a.seed = seed; a.seed_ptr = seed_ptr;
  return launch_reg_r3<MODE_INIT>(p, a, st);
}


// ------------------------------------------------------------------------------------------------
// K1 + K2 for the same lengths: one warp per frame -- window, register FFT, |.|, mel, log1p (the n_fft = 1024 kernel
// stft_fast512_kernel of gl_fast.cu with the generic last radix).  Interior frames (contiguous, 16-byte aligned in the clip)
// arrive by TMA into a two-slot per-warp ring one frame ahead; the reflect-padded edge frames are staged by the warp.
// ------------------------------------------------------------------------------------------------
struct StftRegArgs {
  const float* wave;
  const float* inv_scale;
  int B, L, T, n_mels;
  const float2* tw;
  const float2* rtw;
  const float* win;
  const float* seg_w;     // [2][seg_pad][4] mel column segments of <= 8 bins (see b2d_plan)
  const int* seg_lo;      // [seg_pad]
  const int* seg_first;   // [n_mels + 1]
  int seg_pad;
  float* logmel_bt;
  int exact_sqrt, exact_div;
};
template <int R3> struct StftRegWarps;
template <> struct StftRegWarps<5> { static constexpr int W = 16; };
template <> struct StftRegWarps<12> { static constexpr int W = 10; };

template <int R3, int WARPS>
__global__ void __launch_bounds__(WARPS * 32, 1) stft_reg_kernel(const StftRegArgs a) {
  typedef Geo<R3> G;
  constexpr int M = G::M, N = G::N, HOP = G::HOP, NB = G::NB, NR = G::NR;
  constexpr int WSMEM = G::XCH * 8 + 2 * N * 4 + 16;  // exchange | two frame buffers | two mbarriers
  constexpr int PBASE = (M + 16 + 3) & ~3;            // segment partial sums live behind the magnitudes in the exchange buffer
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float2* WIN = reinterpret_cast<float2*>(smem_raw);  // [M] window pairs
  float2* RT = WIN + M;                               // [M]
  float4* SEGW = reinterpret_cast<float4*>(RT + M);   // [2][seg_pad]
  int* SEGLO = reinterpret_cast<int*>(SEGW + 2 * a.seg_pad);
  int* SEGF = SEGLO + a.seg_pad;
  unsigned char* warp_base = reinterpret_cast<unsigned char*>(SEGF + ((a.n_mels + 4) & ~3));
  for (int i = threadIdx.x; i < M; i += blockDim.x) {
    WIN[i] = make_float2(a.win[2 * i], a.win[2 * i + 1]);
    RT[i] = a.rtw[i];
  }
  for (int i = threadIdx.x; i < 2 * a.seg_pad; i += blockDim.x) SEGW[i] = reinterpret_cast<const float4*>(a.seg_w)[i];
  for (int i = threadIdx.x; i < a.seg_pad; i += blockDim.x) SEGLO[i] = a.seg_lo[i];
  for (int i = threadIdx.x; i <= a.n_mels; i += blockDim.x) SEGF[i] = a.seg_first[i];
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  unsigned char* wsm = warp_base + (size_t)warp * WSMEM;
  float2* S = reinterpret_cast<float2*>(wsm);
  float* Sf = reinterpret_cast<float*>(S);
  float* fbuf = reinterpret_cast<float*>(wsm + G::XCH * 8);  // [2][N]
  uint64_t* fbar = reinterpret_cast<uint64_t*>(wsm + G::XCH * 8 + 2 * N * 4);
  if (lane == 0) {
    tma::barrier_init(fbar, 1);
    tma::barrier_init(fbar + 1, 1);
    tma::fence_barrier_init();
  }
  __syncwarp();
  LaneTwR<R3> tw;
  lane_twiddles_r<R3>(lane, a.tw, tw);
  const unsigned nframes = (unsigned)a.B * (unsigned)a.T;
  const unsigned stride = gridDim.x * WARPS;
  const bool tma_ok = (a.L % 4) == 0;  // 16-byte aligned clip rows
  auto interior = [&](unsigned b, unsigned t, const float*& src) {
    const long s0 = (long)t * HOP - HOP;
    src = a.wave + (size_t)b * a.L + s0;
    return tma_ok && s0 >= 0 && s0 + N <= a.L;
  };
  uint32_t use0 = 0, use1 = 0;
  unsigned f = blockIdx.x * WARPS + warp;
  unsigned b = f / (unsigned)a.T, t = f - b * (unsigned)a.T;
  int slot = 0;
  if (f < nframes && lane == 0) {
    const float* src;
    if (interior(b, t, src)) { tma::expect_bytes(fbar, N * 4); tma::load(fbuf, src, N * 4, fbar); }
  }
  unsigned bn = 0, tn = 0;
#pragma unroll 1
  for (; f < nframes; f += stride, slot ^= 1, b = bn, t = tn) {
    const float* x = a.wave + (size_t)b * a.L;
    const float pk = a.inv_scale ? a.inv_scale[b] : 1.0f;
    const float rsc = 1.0f / pk;  // x / peak as x * (1 / peak): within one ulp of the division (exact_div: the division itself)
    float* cur = fbuf + slot * N;
    const float* src_cur;
    const bool cur_tma = interior(b, t, src_cur);
    bn = (f + stride) / (unsigned)a.T;
    tn = (f + stride) - bn * (unsigned)a.T;
    if (lane == 0 && f + stride < nframes) {  // next frame of this warp into the other slot
      const float* src;
      if (interior(bn, tn, src)) { tma::expect_bytes(fbar + (slot ^ 1), N * 4); tma::load(fbuf + (slot ^ 1) * N, src, N * 4, fbar + (slot ^ 1)); }
    }
    if (cur_tma) {
      if (slot) { tma::wait(fbar + 1, use1 & 1); ++use1; } else { tma::wait(fbar, use0 & 1); ++use0; }
    } else {  // reflect-padded edge frame (or unaligned clip length): stage it here
      const long s0 = (long)t * HOP - HOP;
      for (int i = lane; i < N; i += 32) {
        long sidx = s0 + i;
        if (sidx < 0) sidx = -sidx;
        if (sidx >= a.L) sidx = 2L * (a.L - 1) - sidx;
        cur[i] = (sidx >= 0 && sidx < a.L) ? x[sidx] : 0.f;
      }
      __syncwarp();
    }
    float2 v[G::NV];
    {
      const float2* c2 = reinterpret_cast<const float2*>(cur);
#pragma unroll
      for (int rr = 0; rr < NR; ++rr)
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const int i = lane + 32 * rr;
          float2 xv = make_float2(0.f, 0.f);
          if (G::FULL || i < NB) {
            const int m = i + NB * q;
            xv = c2[m];
            xv = a.exact_div ? make_float2(xv.x / pk, xv.y / pk) : cscale2(xv, make_float2(rsc, rsc));
            xv = cscale2(xv, WIN[m]);
          }
          v[8 * rr + q] = xv;
        }
    }
    __syncwarp();
    fwd1_store_r<R3>(lane, v, tw, S);
    __syncwarp();
    fwd2_load_r<R3>(lane, v, S);
    __syncwarp();
    fwd2_store_r<R3>(lane, v, tw, S);
    __syncwarp();
    float2 wA[R3], wB[R3], U[R3], V[R3];
    fwd3_load_r<R3>(lane, wA, wB, S);
    __syncwarp();  // the exchange buffer is re-used for the magnitudes
    gather_pairs<R3>(lane, wA, wB, U, V);
#pragma unroll
    for (int r = 0; r < R3; ++r) {
      const int k = slot_k_r<R3>(lane, r);
      if (r == 0 && lane == 0) {
        Sf[0] = fabsf(U[0].x + U[0].y);
        Sf[M] = fabsf(U[0].x - U[0].y);
        const float sh = V[0].x * V[0].x + V[0].y * V[0].y;
        Sf[M / 2] = sqrtf(sh);
      } else {
        float2 xk, xmk;
        rfft_split(U[r], V[r], RT[k], xk, xmk);
        const float s1 = xk.x * xk.x + xk.y * xk.y, s2 = xmk.x * xmk.x + xmk.y * xmk.y;
        Sf[k] = sqrtf(s1);
        Sf[M - k] = sqrtf(s2);
      }
    }
    if (lane < 8) Sf[M + 1 + lane] = 0.f;  // the zero-weight taps of a column's last segment read up to 7 bins past the Nyquist bin
    __syncwarp();
    float* P = Sf + PBASE;
    for (int s0 = 0; s0 < a.seg_pad; s0 += 32) {  // lane s accumulates segment s (<= 8 consecutive bins of one mel column) ...
      const int sg = s0 + lane;
      const float* mp = Sf + SEGLO[sg];
      const float4 w0 = SEGW[sg], w1 = SEGW[a.seg_pad + sg];
      float acc = mp[0] * w0.x;
      acc = fmaf(mp[1], w0.y, acc); acc = fmaf(mp[2], w0.z, acc); acc = fmaf(mp[3], w0.w, acc);
      acc = fmaf(mp[4], w1.x, acc); acc = fmaf(mp[5], w1.y, acc); acc = fmaf(mp[6], w1.z, acc); acc = fmaf(mp[7], w1.w, acc);
      P[sg] = acc;
    }
    __syncwarp();
    for (int m = lane; m < a.n_mels; m += 32) {  // ... then every mel column sums its segments in order
      float acc = 0.f;
      for (int sg = SEGF[m]; sg < SEGF[m + 1]; ++sg) acc += P[sg];
      a.logmel_bt[(size_t)f * a.n_mels + m] = log1pf(acc);
    }
    __syncwarp();
  }
}

template <int R3>
static int launch_stft_reg_t(const b2d_plan* p, const StftRegArgs& a, cudaStream_t st) {
  constexpr int W = StftRegWarps<R3>::W;
  typedef Geo<R3> G;
  const size_t smem = sizeof(float2) * 2 * G::M + (size_t)p->mel_seg_pad * 36 + sizeof(int) * ((p->n_mels + 4) & ~3) +
                      (size_t)W * (G::XCH * 8 + 2 * G::N * 4 + 16);
  B2D_REQUIRE(smem <= 232448, B2D_ERR_UNSUPPORTED, "mel filterbank too dense for the register STFT kernel");
  B2D_SMEM_OPT_IN(smem, stft_reg_kernel<R3, W>);
  const size_t nframes = (size_t)a.B * a.T;
  const size_t want = (nframes + W - 1) / W;
  const int grid = (int)(want < (size_t)p->num_sms ? want : (size_t)p->num_sms);
  stft_reg_kernel<R3, W><<<grid, W * 32, smem, st>>>(a);
  B2D_LAUNCH_CHECK("stft_reg_kernel");
  return B2D_OK;
}

bool stft_reg_supported(const b2d_plan* p, int B, int L) {
  return gl_reg_r3(p) != 0 && L >= p->n_fft && p->n_mels <= 128 && p->mel_seg_pad <= 320 && (long long)B * (1 + L / p->hop) < (1ll << 30);
}
int launch_stft_reg(const b2d_plan* p, const float* wave, const float* inv_scale, int B, int L, float* logmel_bt, cudaStream_t st) {
  StftRegArgs a;
  a.wave = wave; a.inv_scale = inv_scale; a.B = B; a.L = L; a.T = 1 + L / p->hop; a.n_mels = p->n_mels;
  a.tw = p->d_tw; a.rtw = p->d_rtw; a.win = p->d_win;
  a.seg_w = p->d_seg_w; a.seg_lo = p->d_seg_lo; a.seg_first = p->d_seg_first; a.seg_pad = p->mel_seg_pad;
  a.logmel_bt = logmel_bt;
  a.exact_sqrt = (p->flags & B2D_PLAN_EXACT_SQRT) ? 1 : 0;
  a.exact_div = (p->flags & B2D_PLAN_EXACT_PEAK_DIV) ? 1 : 0;
  return gl_reg_r3(p) == 5 ? launch_stft_reg_t<5>(p, a, st) : launch_stft_reg_t<12>(p, a, st);
}


// reflect-padded edge block with every load in flight at once (the single-launch kernel has no other warps to hide a
// serial chain of L2 round trips behind: the compact loop above costs ~6 us there)
template <int HOP>
__device__ __forceinline__ void reg_stage_reflect_wide(const float* part, const float* prev, float mom, const float* __restrict__ inv_env,
                                                       const float* __restrict__ win_half, int b, int R, int n, int T, int j, float* dst,
                                                       int lane) {
  constexpr int PER = HOP / 32;
  float v[PER], pv[PER];
#pragma unroll
  for (int q = 0; q < PER; ++q) {
    const int i = lane + 32 * q;
    int js, is;
    if (j == 0) { js = (i == 0) ? 2 : 1; is = (i == 0) ? 0 : HOP - i; }
    else        { js = (i == HOP - 1) ? T - 2 : T - 1; is = (i == HOP - 1) ? HOP - 1 : HOP - 2 - i; }
    v[q] = reg_partial_sample<HOP>(part, b, R, n, js, is);
    pv[q] = prev ? reg_partial_sample<HOP>(prev, b, R, n, js, is) : 0.f;
  }
#pragma unroll
  for (int q = 0; q < PER; ++q) {
    const int i = lane + 32 * q;
    const int is = (j == 0) ? ((i == 0) ? 0 : HOP - i) : ((i == HOP - 1) ? HOP - 1 : HOP - 2 - i);
    dst[i] = fmaf(-mom, pv[q], v[q]) * inv_env[is] * win_half[i];
  }
}

// ------------------------------------------------------------------------------------------------
// Small problems in ONE launch: a streaming hop (T = 3 frames per session, app3.py:213) or a handful of whole clips
// (BASELINE config 1: one 4 s clip; the B <= 16 points of the config-5 sweep).  Per-iteration launches are launch-bound
// there (12 us per iteration for one clip).  Here a clip belongs to one thread-block CLUSTER: its runs are dealt over the
// cluster's warps (one frame per warp when the clip has at most 8 x WARPS frames), the init and all n_iter iterations run
// inside the kernel, separated by barrier.cluster (release / acquire: the iterates live in global memory, i.e. in L2, and
// cross CTA boundaries at run edges).  Same arithmetic and same partial hop-block format as the batch kernel above.
// ------------------------------------------------------------------------------------------------
struct GlRegFusedArgs {
  const float* mag_tf;      // [B,T,Fp]
  const float2* angles0;    // [B,F,T] torch layout or null
  float* x[3];              // three iterates, partial hop-block format
  int B, T, n, R, Fp, F, n_iter, csize;
  const float2* tw;
  const float2* rtw;
  const float* win;
  const float* winn;
  const float* inv_env;
  float mom;
  float* wave;              // [B, HOP*(T-1)]
  const float* out_scale;   // [B] or null
  unsigned long long seed;
  const unsigned long long* seed_ptr;
  // streaming hop only (gl_reg_hop_kernel, T == 3): overlap-add ring update fused behind the last iteration (app3.py:219-224):
  // hop_out[s] = ola[s][:hop]; ola[s] = [ola[s][hop:] + y[:hop], y[hop:]] with y = the hop's 2 * hop output samples; wave unused
  float* ola;               // [B, 2 * HOP] or null
  float* hop_out;           // [B, HOP]
};

__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("fence.proxy.async;\n" ::: "memory");  // this thread's global writes -> visible to the TMA (async proxy) reads of the next step
  asm volatile("barrier.cluster.arrive.release.aligned;\n" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;\n" ::: "memory");
}

template <int R3>
struct FusedSmem {  // per warp: exchange | mag row | mbarrier  (no iterate ring: every hop-block is a run-boundary block when n == 1)
  typedef Geo<R3> G;
  static constexpr int OFF_MAG = G::XCH * 8;
  static constexpr int OFF_BAR = OFF_MAG + ((RegSmem<R3>::MAG_BYTES + 15) & ~15);
  static constexpr int WARP_BYTES = OFF_BAR + 16;
};

template <int R3, int WARPS>
__global__ void __launch_bounds__(WARPS * 32, 1) gl_reg_fused_kernel(const GlRegFusedArgs a) {
  typedef Geo<R3> G;
  typedef FusedSmem<R3> SM;
  constexpr int M = G::M, HOP = G::HOP, NB = G::NB, NR = G::NR, H2 = M / 2;
  constexpr int MAG_BYTES = RegSmem<R3>::MAG_BYTES;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float2* WA = reinterpret_cast<float2*>(smem_raw);
  float2* WB = WA + H2;
  float2* WN = WB + H2;
  float2* RT = WN + M;
  unsigned char* warp_base = reinterpret_cast<unsigned char*>(RT + M);
  for (int i = threadIdx.x; i < H2; i += blockDim.x) {
    WA[i] = make_float2(a.inv_env[2 * i] * a.win[2 * i], a.inv_env[2 * i + 1] * a.win[2 * i + 1]);
    WB[i] = make_float2(a.inv_env[2 * i] * a.win[HOP + 2 * i], a.inv_env[2 * i + 1] * a.win[HOP + 2 * i + 1]);
  }
  for (int i = threadIdx.x; i < M; i += blockDim.x) {
    WN[i] = make_float2(a.winn[2 * i], a.winn[2 * i + 1]);
    RT[i] = a.rtw[i];
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int T = a.T;  // one frame per run: R == T, n == 1; run r owns slots 2 r (first half of frame r) and 2 r + 1 (second half)
  const int b = (int)blockIdx.x / a.csize, crank = (int)blockIdx.x % a.csize;
  const int t = crank * WARPS + warp;  // this warp's frame
  const bool active = t < T;
  unsigned char* wsm = warp_base + (size_t)warp * SM::WARP_BYTES;
  float2* S = reinterpret_cast<float2*>(wsm);
  float* mg_s = reinterpret_cast<float*>(wsm + SM::OFF_MAG);
  uint64_t* bar = reinterpret_cast<uint64_t*>(wsm + SM::OFF_BAR);
  if (lane == 0) {
    tma::barrier_init(bar, 1);
    tma::fence_barrier_init();
    if (active) {  // the frame's magnitude row: loaded once, resident for the init and all iterations
      tma::expect_bytes(bar, MAG_BYTES);
      tma::load(mg_s, a.mag_tf + ((size_t)b * T + t) * a.Fp, MAG_BYTES, bar);
    }
  }
  __syncwarp();
  LaneTwR<R3> tw;
  lane_twiddles_r<R3>(lane, a.tw, tw);
  const float2 nmom = make_float2(-a.mom, -a.mom);
  const unsigned long long seed = a.seed_ptr ? *a.seed_ptr : a.seed;
  const size_t clip = (size_t)b * T * 2 * HOP;  // this clip's slots in an iterate buffer
  if (active) tma::wait(bar, 0);

#pragma unroll 1
  for (int step = 0; step <= a.n_iter; ++step) {
    const bool init = (step == 0);
    const bool use_prev = (step >= 2) && (a.mom != 0.f);
    const float* xin = a.x[(step + 2) % 3] + clip;    // x_k       (written by step - 1)
    const float* xprev = a.x[(step + 1) % 3] + clip;  // x_{k-1}   (written by step - 2)
    float* xout = a.x[step % 3] + clip;
    if (active) {
      float2 v[G::NV];
      float2 wA[R3], wB[R3];
      if (!init) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int j = t + h;  // padded hop-block index: block j = slot (2 (j - 1) + 1) + slot (2 j), 1 <= j <= T - 1
          const float2* wtab = h ? WB : WA;
          if (j == 0 || j == T) {
            __syncwarp();
            reg_stage_reflect_wide<HOP>(xin - clip, use_prev ? xprev - clip : nullptr, a.mom, a.inv_env, a.win + h * HOP, b, T, 1, T, j,
                                        reinterpret_cast<float*>(S), lane);
            __syncwarp();
#pragma unroll
            for (int rr = 0; rr < NR; ++rr)
#pragma unroll
              for (int q = 0; q < 4; ++q) {
                const int i = lane + 32 * rr;
                v[8 * rr + 4 * h + q] = (G::FULL || i < NB) ? S[i + NB * q] : make_float2(0.f, 0.f);
              }
          } else {
            const float2* p1 = reinterpret_cast<const float2*>(xin + (size_t)(2 * j - 1) * HOP);   // second half of frame j - 1
            const float2* p2 = reinterpret_cast<const float2*>(xin + (size_t)(2 * j) * HOP);       // first half of frame j
            const float2* q1 = reinterpret_cast<const float2*>(xprev + (size_t)(2 * j - 1) * HOP);
            const float2* q2 = reinterpret_cast<const float2*>(xprev + (size_t)(2 * j) * HOP);
#pragma unroll
            for (int rr = 0; rr < NR; ++rr)
#pragma unroll
              for (int q = 0; q < 4; ++q) {
                const int i = lane + 32 * rr;
                float2 xv = make_float2(0.f, 0.f);
                if (G::FULL || i < NB) {
                  const int m = i + NB * q;
                  xv = cadd(__ldcg(p1 + m), __ldcg(p2 + m));  // written by other warps / CTAs of the cluster one step ago: read from L2
                  if (use_prev) xv = cfma2(cadd(__ldcg(q1 + m), __ldcg(q2 + m)), nmom, xv);
                  xv = cscale2(xv, wtab[m]);
                }
                v[8 * rr + 4 * h + q] = xv;
              }
          }
        }
        __syncwarp();
        fwd1_store_r<R3>(lane, v, tw, S);
        __syncwarp();
        fwd2_load_r<R3>(lane, v, S);
        __syncwarp();
        fwd2_store_r<R3>(lane, v, tw, S);
        __syncwarp();
        fwd3_load_r<R3>(lane, wA, wB, S);
        project_frame<R3>(lane, wA, wB, RT, mg_s);
      } else if (a.angles0) {
        init_frame_angles<R3>(lane, wA, wB, RT, mg_s, a.angles0 + (size_t)b * a.F * T + t, T);
      } else {
        init_frame<R3>(lane, wA, wB, RT, mg_s, seed, ((unsigned long long)b * T + t) * (M + 1));
      }
      __syncwarp();
      inv1_store_r<R3>(lane, wA, wB, S);
      __syncwarp();
      inv2_load_r<R3>(lane, v, tw, S);
      __syncwarp();
      inv2_store_r<R3>(lane, v, S);
      __syncwarp();
      inv3_load_r<R3>(lane, v, tw, S);
      float2* d0 = reinterpret_cast<float2*>(xout + (size_t)(2 * t) * HOP);      // slot 0: first half x window
      float2* d1 = reinterpret_cast<float2*>(xout + (size_t)(2 * t + 1) * HOP);  // slot 1: second half x window
#pragma unroll
      for (int rr = 0; rr < NR; ++rr)
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int i = lane + 32 * rr;
          if (G::FULL || i < NB) {
            const int m = i + NB * q;
            d0[m] = cscale2(v[8 * rr + q], WN[m]);
            d1[m] = cscale2(v[8 * rr + 4 + q], WN[H2 + m]);
          }
        }
    }
    cluster_sync_all();  // x_{k+1} complete and visible to every CTA of the clip's cluster
  }
  // ---- stitch: hop-block j (1 .. T-1) = its two partial slots x 1 / envelope x clip scale ----
  const float* fin = a.x[a.n_iter % 3] + clip;
  const float sc = a.out_scale ? a.out_scale[b] : 1.0f;
  for (int j = 1 + t; j <= T - 1; j += a.csize * WARPS) {
    float* dst = a.wave + (size_t)b * HOP * (T - 1) + (size_t)(j - 1) * HOP;
    const float* s1 = fin + (size_t)(2 * j - 1) * HOP;
    const float* s2 = fin + (size_t)(2 * j) * HOP;
    for (int i = lane; i < HOP; i += 32) dst[i] = (__ldcg(s1 + i) + __ldcg(s2 + i)) * a.inv_env[i] * sc;
  }
}

// ------------------------------------------------------------------------------------------------
// A streaming hop (T <= 4 frames per session, app3.py:213: one n_fft window = 3 frames) in ONE launch with the iterates in
// SHARED memory: one CTA per session, __syncthreads between iterations.  Same arithmetic and the same two-slots-per-frame
// format as gl_reg_fused_kernel above, but x_{k-1}, x_k, x_{k+1} never leave the SM, and a frame belongs to a GROUP of
// NR = ceil(8 R3 / 32) warps: the radix-8 stages 1 and 2 of the forward and inverse transforms are cut by rounds of 32
// butterflies (gl_reg.cuh), and here every round has its own warp (n_fft 640: 2, n_fft 1536: 3), so a lone frame's critical
// path is one round per stage instead of NR; the radix-R3 stage and the projection stay on the group's first warp (bin k and
// its mirror sit in one lane there).  The group's warps meet at a named barrier between the stages.  Bit-identical to the
// one-warp-per-frame kernels (same butterflies, same order).  Measured per hop: DESIGN.md section 4.4.
// ------------------------------------------------------------------------------------------------
constexpr int HOP_FRAMES = 4;  // frames (warp groups) per CTA

template <int R3>
struct HopSmem {  // per frame: exchange buffer (shared by the group's warps) | mag row | mbarrier
  typedef Geo<R3> G;
  static constexpr int OFF_MAG = G::XCH * 8;
  static constexpr int OFF_BAR = OFF_MAG + ((RegSmem<R3>::MAG_BYTES + 15) & ~15);
  static constexpr int FRAME_BYTES = OFF_BAR + 16;
};

__device__ __forceinline__ void group_sync(int id, int nthreads) {  // named barriers 1 .. 15; a one-warp group needs none
  if (nthreads == 32) __syncwarp();
  else asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

template <int R3>
__global__ void __launch_bounds__(HOP_FRAMES * Geo<R3>::NR * 32, 1) gl_reg_hop_kernel(const GlRegFusedArgs a) {
  typedef Geo<R3> G;
  typedef HopSmem<R3> SM;
  constexpr int M = G::M, HOP = G::HOP, NB = G::NB, NR = G::NR, H2 = M / 2;
  constexpr int MAG_BYTES = RegSmem<R3>::MAG_BYTES;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float2* WA = reinterpret_cast<float2*>(smem_raw);
  float2* WB = WA + H2;
  float2* WN = WB + H2;
  float2* RT = WN + M;
  unsigned char* frame_base = reinterpret_cast<unsigned char*>(RT + M);
  float* XS = reinterpret_cast<float*>(frame_base + (size_t)HOP_FRAMES * SM::FRAME_BYTES);  // [3 iterates][T frames][2 slots][HOP]
  for (int i = threadIdx.x; i < H2; i += blockDim.x) {
    WA[i] = make_float2(a.inv_env[2 * i] * a.win[2 * i], a.inv_env[2 * i + 1] * a.win[2 * i + 1]);
    WB[i] = make_float2(a.inv_env[2 * i] * a.win[HOP + 2 * i], a.inv_env[2 * i + 1] * a.win[HOP + 2 * i + 1]);
  }
  for (int i = threadIdx.x; i < M; i += blockDim.x) {
    WN[i] = make_float2(a.winn[2 * i], a.winn[2 * i + 1]);
    RT[i] = a.rtw[i];
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int T = a.T, b = (int)blockIdx.x;
  const int t = warp / NR, r = warp - t * NR;  // this warp's frame and its round of the radix-8 stages
  const bool active = t < T;
  const int bar_id = 1 + t;
  unsigned char* fsm = frame_base + (size_t)t * SM::FRAME_BYTES;
  float2* S = reinterpret_cast<float2*>(fsm);
  float* mg_s = reinterpret_cast<float*>(fsm + SM::OFF_MAG);
  uint64_t* bar = reinterpret_cast<uint64_t*>(fsm + SM::OFF_BAR);
  if (lane == 0 && r == 0) {
    tma::barrier_init(bar, 1);
    tma::fence_barrier_init();
  }
  // this warp's round: butterflies i = lane + 32 r of stages 1 / 2 (the last round is partial when 8 R3 % 32 != 0)
  const int i = lane + 32 * r;
  const bool work = G::FULL || i < NB;
  const int n3 = (lane >> 3) + 4 * r, k1 = lane & 7;
  float2 t1[3], t2[3];
  {
    const int ii = work ? i : 0, nn = work ? n3 : 0;
    t1[0] = a.tw[ii]; t1[1] = a.tw[2 * ii]; t1[2] = a.tw[4 * ii];
    t2[0] = a.tw[8 * nn]; t2[1] = a.tw[16 * nn]; t2[2] = a.tw[32 * nn];
  }
  pdl_wait();  // tables and twiddles above are plan constants; the magnitudes come from the kernel before
  pdl_trigger();
  if (lane == 0 && r == 0 && active) {  // the frame's magnitude row: loaded once, resident for the init and all iterations
    tma::expect_bytes(bar, MAG_BYTES);
    tma::load(mg_s, a.mag_tf + ((size_t)b * T + t) * a.Fp, MAG_BYTES, bar);
  }
  const float2 nmom = make_float2(-a.mom, -a.mom);
  const unsigned long long seed = a.seed_ptr ? *a.seed_ptr : a.seed;
  const int xbuf = T * 2 * HOP;  // floats per iterate
  __syncthreads();               // tables, mbarrier init
  if (active && r == 0) tma::wait(bar, 0);
  // sample `is` of interior hop-block js (1 .. T-1): second half of frame js - 1 + first half of frame js
  auto blk = [&](const float* x, int js, int is) { return x[(2 * js - 1) * HOP + is] + x[(2 * js) * HOP + is]; };

#pragma unroll 1
  for (int step = 0; step <= a.n_iter; ++step) {
    const bool init = (step == 0);
    const bool use_prev = (step >= 2) && (a.mom != 0.f);
    const float* xin = XS + ((step + 2) % 3) * xbuf;    // x_k       (written by step - 1)
    const float* xprev = XS + ((step + 1) % 3) * xbuf;  // x_{k-1}   (written by step - 2)
    float* xout = XS + (step % 3) * xbuf;
    if (active) {
      float2 v[8];
      if (!init) {
        // ---- stage: v[4 h + q] = windowed (x_k - m x_{k-1})[pair i + NB q of half-block h] ----
        if (work) {
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const int j = t + h;  // padded hop-block index
            const float2* wtab = h ? WB : WA;
            if (j == 0 || j == T) {  // reflect-padded edge of the clip (torch.stft center=True)
              const float* wh = a.win + h * HOP;
#pragma unroll
              for (int q = 0; q < 4; ++q) {
                float e[2];
#pragma unroll
                for (int c = 0; c < 2; ++c) {
                  const int idx = 2 * (i + NB * q) + c;
                  int js, is;
                  if (j == 0) { js = (idx == 0) ? 2 : 1; is = (idx == 0) ? 0 : HOP - idx; }
                  else        { js = (idx == HOP - 1) ? T - 2 : T - 1; is = (idx == HOP - 1) ? HOP - 1 : HOP - 2 - idx; }
                  float xv = blk(xin, js, is);
                  if (use_prev) xv = fmaf(-a.mom, blk(xprev, js, is), xv);
                  e[c] = xv * a.inv_env[is] * wh[idx];
                }
                v[4 * h + q] = make_float2(e[0], e[1]);
              }
            } else {
              const float2* p1 = reinterpret_cast<const float2*>(xin + (2 * j - 1) * HOP);
              const float2* p2 = reinterpret_cast<const float2*>(xin + (2 * j) * HOP);
              const float2* q1 = reinterpret_cast<const float2*>(xprev + (2 * j - 1) * HOP);
              const float2* q2 = reinterpret_cast<const float2*>(xprev + (2 * j) * HOP);
#pragma unroll
              for (int q = 0; q < 4; ++q) {
                const int m = i + NB * q;
                float2 xv = cadd(p1[m], p2[m]);
                if (use_prev) xv = cfma2(cadd(q1[m], q2[m]), nmom, xv);
                v[4 * h + q] = cscale2(xv, wtab[m]);
              }
            }
          }
          // ---- forward stage 1: radix 8 over n1, twiddle W_M^{i k1} ----
          float2 p[8];
          tw_powers(t1, p);
          dft8<false>(v);
#pragma unroll
          for (int kk = 0; kk < 8; ++kk) S[kk * G::LD1 + i] = kk ? cmul(v[kk], p[kk]) : v[0];
        }
        group_sync(bar_id, NR * 32);
        // ---- forward stage 2: radix 8 over n2, twiddle W_M^{8 n3 k2} ----
        if (work) {
#pragma unroll
          for (int n2 = 0; n2 < 8; ++n2) v[n2] = S[k1 * G::LD1 + R3 * n2 + n3];
        }
        group_sync(bar_id, NR * 32);  // S is rewritten in the other layout
        if (work) {
          float2 p[8];
          tw_powers(t2, p);
          dft8<false>(v);
#pragma unroll
          for (int k2 = 0; k2 < 8; ++k2) S[n3 * G::LD2 + k1 + 8 * k2] = k2 ? cmul(v[k2], p[k2]) : v[0];
        }
        group_sync(bar_id, NR * 32);
      }
      // ---- stage 3 + projection (or the initial spectrum) + inverse stage 3: the group's first warp ----
      if (r == 0) {
        float2 wA[R3], wB[R3];
        if (!init) {
          fwd3_load_r<R3>(lane, wA, wB, S);
          project_frame<R3>(lane, wA, wB, RT, mg_s);
        } else if (a.angles0) {
          init_frame_angles<R3>(lane, wA, wB, RT, mg_s, a.angles0 + (size_t)b * a.F * T + t, T);
        } else {
          init_frame<R3>(lane, wA, wB, RT, mg_s, seed, ((unsigned long long)b * T + t) * (M + 1));
        }
        __syncwarp();
        inv1_store_r<R3>(lane, wA, wB, S);
      }
      group_sync(bar_id, NR * 32);
      // ---- inverse stage 2, stage 1, synthesis window ----
      if (work) {
        float2 p[8];
        tw_powers(t2, p);
#pragma unroll
        for (int k2 = 0; k2 < 8; ++k2) {
          const float2 c = S[n3 * G::LD2 + k1 + 8 * k2];
          v[k2] = k2 ? cmulc(c, p[k2]) : c;
        }
        dft8<true>(v);
      }
      group_sync(bar_id, NR * 32);  // S is rewritten in the other layout
      if (work) {
#pragma unroll
        for (int n2 = 0; n2 < 8; ++n2) S[k1 * G::LD1 + R3 * n2 + n3] = v[n2];
      }
      group_sync(bar_id, NR * 32);
      if (work) {
        float2 p[8];
        tw_powers(t1, p);
#pragma unroll
        for (int kk = 0; kk < 8; ++kk) {
          const float2 c = S[kk * G::LD1 + i];
          v[kk] = kk ? cmulc(c, p[kk]) : c;
        }
        dft8<true>(v);
        float2* d0 = reinterpret_cast<float2*>(xout + (2 * t) * HOP);      // slot 0: first half x window
        float2* d1 = reinterpret_cast<float2*>(xout + (2 * t + 1) * HOP);  // slot 1: second half x window
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int m = i + NB * q;
          d0[m] = cscale2(v[q], WN[m]);
          d1[m] = cscale2(v[4 + q], WN[H2 + m]);
        }
      }
    }
    __syncthreads();  // x_{k+1} complete (also fences the exchange buffers for the next step)
  }
  // ---- stitch: hop-block j (1 .. T-1) = its two partial slots x 1 / envelope x clip scale ----
  const float* fin = XS + (a.n_iter % 3) * xbuf;
  const float sc = a.out_scale ? a.out_scale[b] : 1.0f;
  const int nw = (int)blockDim.x >> 5;
  if (a.ola != nullptr) {  // T == 3: y = hop-blocks 1 and 2; emit the ring's first half, shift, add
    float* ring = a.ola + (size_t)b * 2 * HOP;
    float* out = a.hop_out + (size_t)b * HOP;
    for (int q = threadIdx.x; q < HOP; q += blockDim.x) {
      const float y0 = blk(fin, 1, q) * a.inv_env[q] * sc, y1 = blk(fin, 2, q) * a.inv_env[q] * sc;
      const float o0 = ring[q], o1 = ring[HOP + q];
      out[q] = o0;
      ring[q] = o1 + y0;
      ring[HOP + q] = y1;
    }
    return;
  }
  for (int j = 1 + warp; j <= T - 1; j += nw) {
    float* dst = a.wave + (size_t)b * HOP * (T - 1) + (size_t)(j - 1) * HOP;
    for (int q = lane; q < HOP; q += 32) dst[q] = blk(fin, j, q) * a.inv_env[q] * sc;
  }
}

// ------------------------------------------------------------------------------------------------
// Small and medium problems (one clip ... a few dozen clips) in ONE COOPERATIVE launch over the whole GPU: persistent CTAs,
// one per SM, each with COOP_GROUPS frame groups of NR warps (a warp per radix-8 round, as in gl_reg_hop_kernel); the frames
// of all clips are dealt round-robin over the groups of the grid; the iterates live in global memory (L2, ld.global.cg) in
// the two-slots-per-frame format, and a grid-wide barrier separates the iterations.  (Measured alternative: per-frame progress
// flags -- st.release after a frame's step, ld.acquire polling of its two or three neighbours before the next -- instead of the
// grid barrier: 6.8 instead of 5.6 us per iteration for one clip, 12.0 instead of 9.3 us for 16; the release -> poll -> load chain
// is three dependent L2 round trips.)  The cluster kernel above keeps a clip
// on 8 - 16 SMs (4 frames per SM sub-partition for a 4 s clip) and the device holds only 8 clusters; here a clip's 126 frames
// spread over 126 groups on as many SM sub-partitions as there are, and B = 16 clips still fit one wave.
// ------------------------------------------------------------------------------------------------
template <int R3> struct CoopGroups { static constexpr int G = 16 / Geo<R3>::NR; };  // 16 warps per CTA (15 at n_fft 1536)

template <int R3>
__global__ void __launch_bounds__(CoopGroups<R3>::G * Geo<R3>::NR * 32, 1) gl_reg_coop_kernel(const GlRegFusedArgs a) {
  typedef Geo<R3> G;
  typedef HopSmem<R3> SM;
  constexpr int M = G::M, HOP = G::HOP, NB = G::NB, NR = G::NR, H2 = M / 2;
  constexpr int MAG_BYTES = RegSmem<R3>::MAG_BYTES;
  cooperative_groups::grid_group grid = cooperative_groups::this_grid();
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float2* WA = reinterpret_cast<float2*>(smem_raw);
  float2* WB = WA + H2;
  float2* WN = WB + H2;
  float2* RT = WN + M;
  unsigned char* frame_base = reinterpret_cast<unsigned char*>(RT + M);
  for (int i = threadIdx.x; i < H2; i += blockDim.x) {
    WA[i] = make_float2(a.inv_env[2 * i] * a.win[2 * i], a.inv_env[2 * i + 1] * a.win[2 * i + 1]);
    WB[i] = make_float2(a.inv_env[2 * i] * a.win[HOP + 2 * i], a.inv_env[2 * i + 1] * a.win[HOP + 2 * i + 1]);
  }
  for (int i = threadIdx.x; i < M; i += blockDim.x) {
    WN[i] = make_float2(a.winn[2 * i], a.winn[2 * i + 1]);
    RT[i] = a.rtw[i];
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int T = a.T;
  const int g = warp / NR, r = warp - g * NR;  // frame group inside the CTA, round of the radix-8 stages
  const int bar_id = 1 + g;
  unsigned char* fsm = frame_base + (size_t)g * SM::FRAME_BYTES;
  float2* S = reinterpret_cast<float2*>(fsm);
  float* mg_s = reinterpret_cast<float*>(fsm + SM::OFF_MAG);
  uint64_t* bar = reinterpret_cast<uint64_t*>(fsm + SM::OFF_BAR);
  if (lane == 0 && r == 0) {
    tma::barrier_init(bar, 1);
    tma::fence_barrier_init();
  }
  const int i = lane + 32 * r;
  const bool work = G::FULL || i < NB;
  const int n3 = (lane >> 3) + 4 * r, k1 = lane & 7;
  float2 t1[3], t2[3];
  {
    const int ii = work ? i : 0, nn = work ? n3 : 0;
    t1[0] = a.tw[ii]; t1[1] = a.tw[2 * ii]; t1[2] = a.tw[4 * ii];
    t2[0] = a.tw[8 * nn]; t2[1] = a.tw[16 * nn]; t2[2] = a.tw[32 * nn];
  }
  const float2 nmom = make_float2(-a.mom, -a.mom);
  const unsigned long long seed = a.seed_ptr ? *a.seed_ptr : a.seed;
  const int nframes = a.B * T;
  // a.csize = groups in use per CTA (the launcher spreads a small problem over all SMs before it fills the groups of any):
  // groups of one rank across the CTAs first, so a clip's frames sit on as many SMs as possible
  const int gu = a.csize;
  const int f0 = g < gu ? ((int)blockIdx.x + g * (int)gridDim.x) : nframes;
  const int fstep = gu * (int)gridDim.x;
  const bool resident = f0 + fstep >= nframes;                      // at most one frame for this group: its mag row stays in smem
  const size_t xclip = (size_t)T * 2 * HOP;                         // floats per clip in an iterate buffer
  uint32_t uses = 0;
  __syncthreads();  // tables, mbarrier init
  if (resident && f0 < nframes && lane == 0 && r == 0) {
    tma::expect_bytes(bar, MAG_BYTES);
    tma::load(mg_s, a.mag_tf + (size_t)f0 * a.Fp, MAG_BYTES, bar);
  }
  bool mag_ready = false;
  auto blk = [&](const float* x, int js, int is) { return __ldcg(x + (size_t)(2 * js - 1) * HOP + is) + __ldcg(x + (size_t)(2 * js) * HOP + is); };

#pragma unroll 1
  for (int step = 0; step <= a.n_iter; ++step) {
    const bool init = (step == 0);
    const bool use_prev = (step >= 2) && (a.mom != 0.f);
#pragma unroll 1
    for (int f = f0; f < nframes; f += fstep) {
      const int b = f / T, t = f - b * T;
      const float* xin = a.x[(step + 2) % 3] + (size_t)b * xclip;    // x_k       (written by step - 1)
      const float* xprev = a.x[(step + 1) % 3] + (size_t)b * xclip;  // x_{k-1}   (written by step - 2)
      float* xout = a.x[step % 3] + (size_t)b * xclip;
      if (!resident && lane == 0 && r == 0) {  // this frame's magnitude row lands while the forward transform runs
        tma::expect_bytes(bar, MAG_BYTES);
        tma::load(mg_s, a.mag_tf + (size_t)f * a.Fp, MAG_BYTES, bar);
      }
      float2 v[8];
      if (!init) {
        if (work) {
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const int j = t + h;  // padded hop-block index
            const float2* wtab = h ? WB : WA;
            if (j == 0 || j == T) {  // reflect-padded edge of the clip (torch.stft center=True)
              const float* wh = a.win + h * HOP;
#pragma unroll
              for (int q = 0; q < 4; ++q) {
                float e[2];
#pragma unroll
                for (int c = 0; c < 2; ++c) {
                  const int idx = 2 * (i + NB * q) + c;
                  int js, is;
                  if (j == 0) { js = (idx == 0) ? 2 : 1; is = (idx == 0) ? 0 : HOP - idx; }
                  else        { js = (idx == HOP - 1) ? T - 2 : T - 1; is = (idx == HOP - 1) ? HOP - 1 : HOP - 2 - idx; }
                  float xv = blk(xin, js, is);
                  if (use_prev) xv = fmaf(-a.mom, blk(xprev, js, is), xv);
                  e[c] = xv * a.inv_env[is] * wh[idx];
                }
                v[4 * h + q] = make_float2(e[0], e[1]);
              }
            } else {
              const float2* p1 = reinterpret_cast<const float2*>(xin + (size_t)(2 * j - 1) * HOP);
              const float2* p2 = reinterpret_cast<const float2*>(xin + (size_t)(2 * j) * HOP);
              const float2* q1 = reinterpret_cast<const float2*>(xprev + (size_t)(2 * j - 1) * HOP);
              const float2* q2 = reinterpret_cast<const float2*>(xprev + (size_t)(2 * j) * HOP);
              float2 xa[4], xb[4], pa[4], pb[4];
#pragma unroll
              for (int q = 0; q < 4; ++q) {  // all loads in flight before the first use (L2 round trips)
                const int m = i + NB * q;
                xa[q] = __ldcg(p1 + m); xb[q] = __ldcg(p2 + m);
                if (use_prev) { pa[q] = __ldcg(q1 + m); pb[q] = __ldcg(q2 + m); }
              }
#pragma unroll
              for (int q = 0; q < 4; ++q) {
                const int m = i + NB * q;
                float2 xv = cadd(xa[q], xb[q]);
                if (use_prev) xv = cfma2(cadd(pa[q], pb[q]), nmom, xv);
                v[4 * h + q] = cscale2(xv, wtab[m]);
              }
            }
          }
          float2 p[8];
          tw_powers(t1, p);
          dft8<false>(v);
#pragma unroll
          for (int kk = 0; kk < 8; ++kk) S[kk * G::LD1 + i] = kk ? cmul(v[kk], p[kk]) : v[0];
        }
        group_sync(bar_id, NR * 32);
        if (work) {
#pragma unroll
          for (int n2 = 0; n2 < 8; ++n2) v[n2] = S[k1 * G::LD1 + R3 * n2 + n3];
        }
        group_sync(bar_id, NR * 32);
        if (work) {
          float2 p[8];
          tw_powers(t2, p);
          dft8<false>(v);
#pragma unroll
          for (int k2 = 0; k2 < 8; ++k2) S[n3 * G::LD2 + k1 + 8 * k2] = k2 ? cmul(v[k2], p[k2]) : v[0];
        }
        group_sync(bar_id, NR * 32);
      }
      if (r == 0) {
        float2 wA[R3], wB[R3];
        if (!resident || !mag_ready) { tma::wait(bar, uses & 1); ++uses; mag_ready = true; }
        if (!init) {
          fwd3_load_r<R3>(lane, wA, wB, S);
          project_frame<R3>(lane, wA, wB, RT, mg_s);
        } else if (a.angles0) {
          init_frame_angles<R3>(lane, wA, wB, RT, mg_s, a.angles0 + (size_t)b * a.F * T + t, T);
        } else {
          init_frame<R3>(lane, wA, wB, RT, mg_s, seed, ((unsigned long long)b * T + t) * (M + 1));
        }
        __syncwarp();
        inv1_store_r<R3>(lane, wA, wB, S);
      }
      group_sync(bar_id, NR * 32);
      if (work) {
        float2 p[8];
        tw_powers(t2, p);
#pragma unroll
        for (int k2 = 0; k2 < 8; ++k2) {
          const float2 c = S[n3 * G::LD2 + k1 + 8 * k2];
          v[k2] = k2 ? cmulc(c, p[k2]) : c;
        }
        dft8<true>(v);
      }
      group_sync(bar_id, NR * 32);
      if (work) {
#pragma unroll
        for (int n2 = 0; n2 < 8; ++n2) S[k1 * G::LD1 + R3 * n2 + n3] = v[n2];
      }
      group_sync(bar_id, NR * 32);
      if (work) {
        float2 p[8];
        tw_powers(t1, p);
#pragma unroll
        for (int kk = 0; kk < 8; ++kk) {
          const float2 c = S[kk * G::LD1 + i];
          v[kk] = kk ? cmulc(c, p[kk]) : c;
        }
        dft8<true>(v);
        float2* d0 = reinterpret_cast<float2*>(xout + (size_t)(2 * t) * HOP);      // slot 0: first half x window
        float2* d1 = reinterpret_cast<float2*>(xout + (size_t)(2 * t + 1) * HOP);  // slot 1: second half x window
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int m = i + NB * q;
          d0[m] = cscale2(v[q], WN[m]);
          d1[m] = cscale2(v[4 + q], WN[H2 + m]);
        }
      }
      group_sync(bar_id, NR * 32);  // the exchange buffer (and, when not resident, the mag row) is free for the next frame
    }
    grid.sync();  // x_{k+1} of every clip complete and visible
  }
  // ---- stitch: hop-block j (1 .. T-1) of clip b = its two partial slots x 1 / envelope x clip scale ----
  const float* fin = a.x[a.n_iter % 3];
  const int nw = (int)(blockDim.x >> 5) * (int)gridDim.x, gwarp = warp * (int)gridDim.x + (int)blockIdx.x;
  for (int e = gwarp; e < a.B * (T - 1); e += nw) {
    const int b = e / (T - 1), j = 1 + e - b * (T - 1);
    const float sc = a.out_scale ? a.out_scale[b] : 1.0f;
    float* dst = a.wave + (size_t)b * HOP * (T - 1) + (size_t)(j - 1) * HOP;
    const float* x = fin + (size_t)b * xclip;
    for (int q = lane; q < HOP; q += 32) dst[q] = blk(x, j, q) * a.inv_env[q] * sc;
  }
}

template <int R3> struct FusedWarps;
template <> struct FusedWarps<4> { static constexpr int W = 16; };
template <> struct FusedWarps<5> { static constexpr int W = 16; };
template <> struct FusedWarps<8> { static constexpr int W = 16; };   // 8 x 16 warps: one frame per warp for a 4 s / 16 kHz clip
template <> struct FusedWarps<12> { static constexpr int W = 12; };

static int fused_r3(const b2d_plan* p) {
  if (p->hop * 2 != p->n_fft) return 0;
  switch (p->n_fft) { case 512: return 4; case 640: return 5; case 1024: return 8; case 1536: return 12; }
  return 0;
}
static int fused_warps(int r3) { return r3 == 4 ? FusedWarps<4>::W : r3 == 5 ? FusedWarps<5>::W : r3 == 8 ? FusedWarps<8>::W : FusedWarps<12>::W; }

template <int R3>
static int launch_fused_t(const b2d_plan* p, const GlRegFusedArgs& a, cudaStream_t st) {
  constexpr int W = FusedWarps<R3>::W;
  const size_t smem = (size_t)RegSmem<R3>::TABLE_BYTES + (size_t)W * FusedSmem<R3>::WARP_BYTES;
  static_assert(RegSmem<R3>::TABLE_BYTES + W * FusedSmem<R3>::WARP_BYTES <= 232448, "per-CTA shared memory exceeds 227 KB");
  B2D_SMEM_OPT_IN(smem, gl_reg_fused_kernel<R3, W>);
  if (a.csize > 8) {
    static std::atomic<int> nonportable{0};
    if (!nonportable.load()) {
      B2D_CUDA(cudaFuncSetAttribute(gl_reg_fused_kernel<R3, W>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
      nonportable.store(1);
    }
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(a.B * a.csize));
  cfg.blockDim = dim3(W * 32);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = (unsigned)a.csize; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  B2D_CUDA(cudaLaunchKernelEx(&cfg, gl_reg_fused_kernel<R3, W>, a));
  B2D_LAUNCH_CHECK("gl_reg_fused_kernel");
  return B2D_OK;
}

// how many clusters of `csize` CTAs of the single-launch kernel the device keeps resident at once (measured on B200: 8 for
// csize 8 -- one per GPC -- although 16 x 8 CTAs are fewer than the 148 SMs); queried once per (length, cluster size)
template <int R3>
static int fused_max_clusters_t(int csize) {
  constexpr int W = FusedWarps<R3>::W;
  static std::atomic<int> cache[5];  // csize 1, 2, 4, 8, 16
  int slot = 0;
  while ((1 << slot) < csize) ++slot;
  int v = cache[slot].load();
  if (v) return v;
  const size_t smem = (size_t)RegSmem<R3>::TABLE_BYTES + (size_t)W * FusedSmem<R3>::WARP_BYTES;
  if (cudaFuncSetAttribute(gl_reg_fused_kernel<R3, W>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return 0;
  if (csize > 8 && cudaFuncSetAttribute(gl_reg_fused_kernel<R3, W>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) != cudaSuccess) return 0;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)csize);
  cfg.blockDim = dim3(W * 32);
  cfg.dynamicSmemBytes = smem;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = (unsigned)csize; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  int n = 0;
  if (cudaOccupancyMaxActiveClusters(&n, gl_reg_fused_kernel<R3, W>, &cfg) != cudaSuccess) { cudaGetLastError(); n = 0; }
  cache[slot].store(n > 0 ? n : -1);
  return n > 0 ? n : -1;
}
static int fused_max_clusters(int r3, int csize) {
  switch (r3) {
    case 4: return fused_max_clusters_t<4>(csize);
    case 5: return fused_max_clusters_t<5>(csize);
    case 8: return fused_max_clusters_t<8>(csize);
    case 12: return fused_max_clusters_t<12>(csize);
  }
  return 0;
}

// Is the single-launch kernel the right tool?  Yes when a clip's frames fit one warp each in a cluster (<= 16 CTAs) and every
// clip's cluster is resident at once.  (n == 1, R == T in the partial hop-block format.)
bool gl_reg_fused_plan(const b2d_plan* p, int B, int T, int* n_out, int* R_out, int* csize_out) {
  const int r3 = fused_r3(p);
  if (!r3 || (p->flags & B2D_PLAN_GENERIC_KERNELS)) return false;
  // a streaming hop (T = 3) stays on the block-cooperative single-launch kernel of griffinlim.cu: 512-1024 threads share
  // three frames there, which beats one warp per frame on latency (measured 0.25 vs 0.50 ms per 20 ms hop)
  if (T <= 4) return false;
  const int W = fused_warps(r3);
  int csize = 1;
  while (csize < 16 && csize * W < T) csize *= 2;
  if (csize * W < T) return false;
  if (B > fused_max_clusters(r3, csize)) return false;  // a second wave of clusters would cost what the per-iteration launches cost
  *n_out = 1; *R_out = T; *csize_out = csize;
  return true;
}

// ---- the streaming-hop kernel: plan + launch ----
template <int R3>
static size_t hop_smem_bytes(int T) {
  return (size_t)RegSmem<R3>::TABLE_BYTES + (size_t)HOP_FRAMES * HopSmem<R3>::FRAME_BYTES + (size_t)3 * T * 2 * Geo<R3>::HOP * sizeof(float);
}
bool gl_reg_hop_plan(const b2d_plan* p, int B, int T) {
  (void)B;
  return fused_r3(p) != 0 && !(p->flags & B2D_PLAN_GENERIC_KERNELS) && T >= 3 && T <= HOP_FRAMES;
}
template <int R3>
static int launch_hop_t(const GlRegFusedArgs& a, cudaStream_t st) {
  const size_t smem = hop_smem_bytes<R3>(a.T);
  static_assert(RegSmem<R3>::TABLE_BYTES + HOP_FRAMES * HopSmem<R3>::FRAME_BYTES + 3 * HOP_FRAMES * 2 * Geo<R3>::HOP * 4 <= 232448,
                "per-CTA shared memory exceeds 227 KB");
  B2D_SMEM_OPT_IN(hop_smem_bytes<R3>(HOP_FRAMES), gl_reg_hop_kernel<R3>);
  B2D_CUDA(launch_pdl(gl_reg_hop_kernel<R3>, dim3((unsigned)a.B), dim3(HOP_FRAMES * Geo<R3>::NR * 32), smem, st, a));
  B2D_LAUNCH_CHECK("gl_reg_hop_kernel");
  return B2D_OK;
}
int launch_gl_reg_hop(const b2d_plan* p, const float* mag_tf, const float2* angles0, unsigned long long seed,
                      const unsigned long long* seed_ptr, int B, int T, int n_iter, float mom, const float* out_scale, float* wave,
                      cudaStream_t st, float* ola, float* hop_out) {
  GlRegFusedArgs a{};
  B2D_REQUIRE(ola == nullptr || (T == 3 && hop_out != nullptr), B2D_ERR_BAD_ARG, "fused overlap-add needs a 3-frame hop and an output buffer");
  a.ola = ola; a.hop_out = hop_out;
  a.mag_tf = mag_tf; a.angles0 = angles0;
  a.B = B; a.T = T; a.n = 1; a.R = T; a.Fp = p->Fp; a.F = p->F; a.n_iter = n_iter; a.csize = 1;
  a.tw = p->d_tw; a.rtw = p->d_rtw; a.win = p->d_win; a.winn = p->d_winn; a.inv_env = p->d_inv_env;
  a.mom = mom; a.wave = wave; a.out_scale = out_scale; a.seed = seed; a.seed_ptr = seed_ptr;
  switch (fused_r3(p)) {
    case 4: return launch_hop_t<4>(a, st);
    case 5: return launch_hop_t<5>(a, st);
    case 8: return launch_hop_t<8>(a, st);
    case 12: return launch_hop_t<12>(a, st);
  }
  return fail(B2D_ERR_UNSUPPORTED, "no streaming-hop Griffin-Lim kernel for n_fft = %d", p->n_fft);
}

// ---- the cooperative kernel: plan + launch ----
template <int R3>
static size_t coop_smem_bytes() { return (size_t)RegSmem<R3>::TABLE_BYTES + (size_t)CoopGroups<R3>::G * HopSmem<R3>::FRAME_BYTES; }
template <int R3>
static int coop_max_ctas_t(int num_sms) {  // co-resident CTAs (one per SM when the shared memory fits), queried once
  static std::atomic<int> cache{0};
  int v = cache.load();
  if (v) return v;
  constexpr int threads = CoopGroups<R3>::G * Geo<R3>::NR * 32;
  const size_t smem = coop_smem_bytes<R3>();
  if (cudaFuncSetAttribute(gl_reg_coop_kernel<R3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return 0;
  int per_sm = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, gl_reg_coop_kernel<R3>, threads, smem) != cudaSuccess) return 0;
  v = per_sm > 0 ? num_sms : 0;  // one CTA per SM is all the kernel asks for
  cache.store(v > 0 ? v : -1);
  return v;
}
static int coop_max_ctas(int r3, int num_sms) {
  int v = 0;
  switch (r3) {
    case 4: v = coop_max_ctas_t<4>(num_sms); break;
    case 5: v = coop_max_ctas_t<5>(num_sms); break;
    case 8: v = coop_max_ctas_t<8>(num_sms); break;
    case 12: v = coop_max_ctas_t<12>(num_sms); break;
  }
  return v > 0 ? v : 0;
}
static int coop_groups(int r3) { return r3 == 4 ? CoopGroups<4>::G : r3 == 5 ? CoopGroups<5>::G : r3 == 8 ? CoopGroups<8>::G : CoopGroups<12>::G; }
// worth it while a frame group carries at most a few frames per iteration: beyond that the per-iteration batch kernels (TMA rings,
// register-resident overlap-add) win
bool gl_reg_coop_plan(const b2d_plan* p, int B, int T) {
  const int r3 = fused_r3(p);
  if (!r3 || (p->flags & B2D_PLAN_GENERIC_KERNELS) || T <= HOP_FRAMES) return false;
  const int ctas = coop_max_ctas(r3, p->num_sms);
  if (!ctas) return false;
  return (long)B * T <= 3L * ctas * coop_groups(r3);
}
template <int R3>
static int launch_coop_t(const b2d_plan* p, GlRegFusedArgs a, cudaStream_t st) {
  constexpr int G = CoopGroups<R3>::G, threads = G * Geo<R3>::NR * 32;
  const size_t smem = coop_smem_bytes<R3>();
  static_assert(RegSmem<R3>::TABLE_BYTES + G * HopSmem<R3>::FRAME_BYTES <= 232448, "per-CTA shared memory exceeds 227 KB");
  B2D_SMEM_OPT_IN(smem, gl_reg_coop_kernel<R3>);
  const int nframes = a.B * a.T;
  const int grid = nframes < p->num_sms ? nframes : p->num_sms;
  int gu = (nframes + grid - 1) / grid;  // frame groups in use per CTA
  if (gu > G) gu = G;
  a.csize = gu;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)grid);
  cfg.blockDim = dim3(threads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeCooperative;
  attr[0].val.cooperative = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  B2D_CUDA(cudaLaunchKernelEx(&cfg, gl_reg_coop_kernel<R3>, a));
  B2D_LAUNCH_CHECK("gl_reg_coop_kernel");
  return B2D_OK;
}
int launch_gl_reg_coop(const b2d_plan* p, const float* mag_tf, const float2* angles0, unsigned long long seed,
                       const unsigned long long* seed_ptr, float* x0, float* x1, float* x2, int B, int T, int n_iter,
                       float mom, const float* out_scale, float* wave, cudaStream_t st) {
  GlRegFusedArgs a{};
  a.mag_tf = mag_tf; a.angles0 = angles0; a.x[0] = x0; a.x[1] = x1; a.x[2] = x2;
  a.B = B; a.T = T; a.n = 1; a.R = T; a.Fp = p->Fp; a.F = p->F; a.n_iter = n_iter; a.csize = 1;
  a.tw = p->d_tw; a.rtw = p->d_rtw; a.win = p->d_win; a.winn = p->d_winn; a.inv_env = p->d_inv_env;
  a.mom = mom; a.wave = wave; a.out_scale = out_scale; a.seed = seed; a.seed_ptr = seed_ptr;
  switch (fused_r3(p)) {
    case 4: return launch_coop_t<4>(p, a, st);
    case 5: return launch_coop_t<5>(p, a, st);
    case 8: return launch_coop_t<8>(p, a, st);
    case 12: return launch_coop_t<12>(p, a, st);
  }
  return fail(B2D_ERR_UNSUPPORTED, "no cooperative Griffin-Lim kernel for n_fft = %d", p->n_fft);
}

int launch_gl_reg_fused(const b2d_plan* p, const float* mag_tf, const float2* angles0, unsigned long long seed,
                        const unsigned long long* seed_ptr, float* x0, float* x1, float* x2, int B, int T, int n, int R, int csize,
                        int n_iter, float mom, const float* out_scale, float* wave, cudaStream_t st) {
  GlRegFusedArgs a{};
  a.mag_tf = mag_tf; a.angles0 = angles0; a.x[0] = x0; a.x[1] = x1; a.x[2] = x2;
  a.B = B; a.T = T; a.n = n; a.R = R; a.Fp = p->Fp; a.F = p->F; a.n_iter = n_iter; a.csize = csize;
  a.tw = p->d_tw; a.rtw = p->d_rtw; a.win = p->d_win; a.winn = p->d_winn; a.inv_env = p->d_inv_env;
  a.mom = mom; a.wave = wave; a.out_scale = out_scale; a.seed = seed; a.seed_ptr = seed_ptr;
  switch (fused_r3(p)) {
    case 4: return launch_fused_t<4>(p, a, st);
    case 5: return launch_fused_t<5>(p, a, st);
    case 8: return launch_fused_t<8>(p, a, st);
    case 12: return launch_fused_t<12>(p, a, st);
  }
  return fail(B2D_ERR_UNSUPPORTED, "no single-launch Griffin-Lim kernel for n_fft = %d", p->n_fft);
}

}  // namespace b2d
