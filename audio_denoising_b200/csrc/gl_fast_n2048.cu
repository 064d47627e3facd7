// gl_fast_n2048.cu -- Griffin-Lim iteration for n_fft = 2048 (hop 1024) on the register FFT of gl_fast.cuh.
//
// The 1024-point complex transform of a frame is ONE radix-2 decimation-in-time step over TWO 512-point register
// transforms.  A PAIR OF WARPS owns a frame: the even warp transforms the even complex samples z[2m] = (x[4m], x[4m+1]),
// the odd warp z[2m+1] = (x[4m+2], x[4m+3]).  After the forward passes each lane holds bins k and 512-k of its transform;
// the warps swap half of their pair slots through shared memory so that every lane ends up with E[k], E[512-k], O[k],
// O[512-k] for four slots = bins k, 512-k, 512+k, 1024-k of the frame: the radix-2 combine, the real-FFT split, the phase
// update and their inverses are lane-local (quad_update / quad_special in gl_fast.cuh), then the halves swap back.
// Per frame the pair meets at two named barriers (bar.sync id, 64); there is no CTA-wide barrier after start-up.
// tprev (8 KB) and mag (4 KB) rows and the iterate's hop-blocks (4 KB) travel by TMA one frame ahead, as in the
// n_fft = 1024 kernel.  One persistent 12-warp CTA per SM = six frames in flight.
#include <stdlib.h>

#include "gl_fast.cuh"
#include "kernels.cuh"

namespace b2d {

using namespace fast512;

namespace n2048 {
constexpr int HOP2 = 1024;          // hop = complex points per frame
constexpr int FP2 = 1028;           // frame-layout row stride of the magnitudes (floats)
constexpr int TP_BYTES = HOP2 * 8;  // tprev row
constexpr int MG_BYTES = FP2 * 4;   // mag row (4112 B)
constexpr int PAIRS = 6;            // warp pairs per CTA
// per pair: exchange buffers of the even / odd warp | swap-back area | tprev row | mag row | two hop-blocks | mbarriers
constexpr int OFF_X = 2 * XCH * 8;
constexpr int OFF_TP = OFF_X + 4096;
constexpr int OFF_MG = OFF_TP + TP_BYTES;
constexpr int OFF_XR = OFF_MG + 4128;
constexpr int OFF_BAR = OFF_XR + 2 * HOP2 * 4;
constexpr int PSMEM = OFF_BAR + 32;                          // 33856 B
constexpr int TABLE_BYTES = (1024 + 1024 + 2048) * 4 + 512 * 8;  // WA, WB, WN + W2048^k (k < 512)
}  // namespace n2048

struct GlN2048Args {
  const float* mag_tf;   // [B,T,1028]
  float2* tprev;         // [B,T,1024]  (2 x rebuilt, bin 0 = (DC, Nyquist))
  const float* xin;      // partial hop-block format, hop 1024
  float* xout;
  int B, T, n, R;
  const float2* tw512;   // W512^k
  const float2* rtw;     // W2048^k, k < 1024
  const float* win;      // [2048]
  const float* winn;     // [2048] win / 2048
  const float* inv_env;  // [1024]
  float mom;
  int use_prev, store_prev;
  float* wave;             // last iteration: run-interior hop-blocks are final (see gl_fast.cu)
  const float* out_scale;
  unsigned long long seed;             // INIT mode: x_0 = istft(mag * angles_0), angles drawn in-kernel (0 = all ones)
  const unsigned long long* seed_ptr;
};

__device__ __forceinline__ uint32_t n20_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void n20_mbar_init(uint64_t* bar) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(n20_u32(bar)) : "memory");
}
__device__ __forceinline__ void n20_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(n20_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void n20_bulk(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(n20_u32(dst)), "l"(src),
               "r"(bytes), "r"(n20_u32(bar))
               : "memory");
}
__device__ __forceinline__ void n20_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "W_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra D_%=;\n"
      "bra W_%=;\n"
      "D_%=:\n"
      "}\n" ::"r"(n20_u32(bar)),
      "r"(parity)
      : "memory");
}

__device__ __forceinline__ void pair_sync(int id) { asm volatile("bar.sync %0, 64;" ::"r"(id) : "memory"); }

template <bool USE_PREV, bool INIT>
__global__ void __launch_bounds__(n2048::PAIRS * 64, 1) gl_fast_n2048_kernel(const GlN2048Args a) {
  using namespace n2048;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  // window tables, de-interleaved per warp role so that a lane's float2 reads are conflict-free:
  // entry [role * 256 + m] covers samples 4m + 2 role, 4m + 2 role + 1 of a hop-block
  float2* WA = reinterpret_cast<float2*>(smem_raw);  // inv_env * analysis window, first hop   [2][256]
  float2* WB = WA + 512;                             // inv_env * analysis window, second hop  [2][256]
  float2* WN0 = WB + 512;                            // synthesis window / 2048, first hop     [2][256]
  float2* WN1 = WN0 + 512;                           // synthesis window / 2048, second hop    [2][256]
  float2* RT = WN1 + 512;                            // W2048^k, k < 512
  unsigned char* pair_base = reinterpret_cast<unsigned char*>(RT + 512);
  for (int i = threadIdx.x; i < 512; i += blockDim.x) {
    const int s0 = 4 * (i & 255) + 2 * (i >> 8);
    const float e0 = a.inv_env[s0], e1 = a.inv_env[s0 + 1];
    WA[i] = make_float2(e0 * a.win[s0], e1 * a.win[s0 + 1]);
    WB[i] = make_float2(e0 * a.win[1024 + s0], e1 * a.win[1024 + s0 + 1]);
    WN0[i] = make_float2(a.winn[s0], a.winn[s0 + 1]);
    WN1[i] = make_float2(a.winn[1024 + s0], a.winn[1024 + s0 + 1]);
    RT[i] = a.rtw[i];
  }
  __syncthreads();

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int pair = warp >> 1, odd = warp & 1;
  const int bar_id = 1 + pair;
  const int n = a.n, R = a.R, T = a.T;
  const int nruns = a.B * R;
  const int gp0 = pair * (int)gridDim.x + (int)blockIdx.x;  // runs dealt round-robin over the CTAs (one CTA per SM)
  const int gstep = PAIRS * (int)gridDim.x;
  if (gp0 >= nruns) return;
  unsigned char* psm = pair_base + (size_t)pair * PSMEM;
  float2* S = reinterpret_cast<float2*>(psm) + odd * XCH;          // this warp's exchange buffer
  float2* So = reinterpret_cast<float2*>(psm) + (odd ^ 1) * XCH;   // the partner's
  float2* X = reinterpret_cast<float2*>(psm + OFF_X);              // swap-back area: [0,256) even -> odd, [256,512) odd -> even
  const float2* tp_s = reinterpret_cast<const float2*>(psm + OFF_TP);
  const float* mg_s = reinterpret_cast<const float*>(psm + OFF_MG);
  float* xs = reinterpret_cast<float*>(psm + OFF_XR);
  uint64_t* bar = reinterpret_cast<uint64_t*>(psm + OFF_BAR);
  uint64_t* xbar = bar + 1;
  const bool leader = (odd == 0 && lane == 0);
  if (leader) {
    n20_mbar_init(bar);
    n20_mbar_init(xbar);
    n20_mbar_init(xbar + 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  pair_sync(bar_id);
  LaneTw tw;
  lane_twiddles(lane, a.tw512, tw);
  const int kU0 = lane, kU4 = lane - (lane == 0 ? 224 : 0);
  const size_t run_stride = (size_t)(n + 1) * HOP2;
  const int sub = 2 * odd;  // float offset of this warp's complex sample inside a float4 of the frame
  uint32_t uses = 0, xuse0 = 0, xuse1 = 0;
  pdl_wait();  // the prologue above read plan tables only (common.cuh: programmatic dependent launch)
  pdl_trigger();

#pragma unroll 1
  for (int gp = gp0; gp < nruns; gp += gstep) {
    const int b = gp / R, r = gp - b * R;
    const int tb = r * n, te = min(T, tb + n);
    const int nrun = te - tb;
    const float* xrun = a.xin + (size_t)(b * R + r) * run_stride;
    float* xo = a.xout + (size_t)(b * R + r) * run_stride;
    float2 carry[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) carry[q] = make_float2(0.f, 0.f);

    const uint32_t row_bytes = MG_BYTES + (USE_PREV ? TP_BYTES : 0);
    if (leader) {
      n20_expect_tx(bar, row_bytes);
      n20_bulk(psm + OFF_MG, a.mag_tf + ((size_t)b * T + tb) * FP2, MG_BYTES, bar);
      if (USE_PREV) n20_bulk(psm + OFF_TP, a.tprev + ((size_t)b * T + tb) * HOP2, TP_BYTES, bar);
      if (!INIT && nrun > 1) {  // own-slot hop-blocks 1 .. nrun-1 travel by TMA, one frame ahead
        n20_expect_tx(xbar + 1, HOP2 * 4);
        n20_bulk(xs + HOP2, xrun + HOP2, HOP2 * 4, xbar + 1);
      }
    }
#pragma unroll 1
    for (int t = tb; t < te; ++t) {
      const int c = t - tb;
      float2 v[16];
      if (!INIT) {
      // ---- stage this warp's half of the frame: complex samples (x[4m + sub], x[4m + sub + 1]), m = lane + 32 q ---------
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int j = t + h, cs = c + h;
        const float2* wtab = (h ? WB : WA) + odd * 256;
        if (j == 0 || j == T) {  // reflect-padded edge of the clip: generic path through this warp's exchange buffer
          __syncwarp();
          stage_reflect_wide<HOP2>(a.xin, nullptr, 0.f, a.inv_env, a.win + h * HOP2, b, R, n, T, j, reinterpret_cast<float*>(S), lane);
          __syncwarp();
#pragma unroll
          for (int q = 0; q < 8; ++q) v[8 * h + q] = *reinterpret_cast<const float2*>(reinterpret_cast<const float*>(S) + 4 * (lane + 32 * q) + sub);
          __syncwarp();
        } else if (cs >= 1 && cs <= nrun - 1) {  // interior block of this run: in the shared-memory ring
          if (h == 1) {
            if (cs & 1) { n20_wait(xbar + 1, xuse1 & 1); ++xuse1; } else { n20_wait(xbar, xuse0 & 1); ++xuse0; }
          }
          const float* src = xs + (cs & 1) * HOP2 + sub;
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            const float2 xv = *reinterpret_cast<const float2*>(src + 4 * (lane + 32 * q));
            const float2 wv = wtab[lane + 32 * q];
            v[8 * h + q] = cscale2(xv, wv);
          }
        } else {
          const float* p1 = xrun + (size_t)cs * HOP2 + sub;
          const float* p2 = nullptr;  // second partial when the block sits on a run boundary
          if (cs == 0) p2 = xrun - run_stride + (size_t)n * HOP2 + sub;
          else if (j == te) p2 = xrun + run_stride + sub;
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            float2 xv = *reinterpret_cast<const float2*>(p1 + 4 * (lane + 32 * q));
            if (p2) { const float2 x2 = *reinterpret_cast<const float2*>(p2 + 4 * (lane + 32 * q)); xv.x += x2.x; xv.y += x2.y; }
            const float2 wv = wtab[lane + 32 * q];
            v[8 * h + q] = cscale2(xv, wv);
          }
        }
      }
      // ---- forward 512-point transform of this warp's half ------------------------------------------------------------
      __syncwarp();
      fwd1_store(lane, v, tw, S);
      __syncwarp();
      fwd2_load(lane, v, S);
      __syncwarp();
      fwd2_store(lane, v, tw, S);
      __syncwarp();
      fwd3_load(lane, v, S);
      if (lane == 0) lane0_permute(v);
      // ---- swap: the even warp updates pair slots 0..3, the odd warp slots 4..7; each sends the other four ------------
      __syncwarp();
#pragma unroll
      for (int i = 0; i < 4; ++i) {  // slots given away: i + 4 (even warp) / i (odd warp); static register indices + selects
        S[(2 * i) * 32 + lane] = odd ? v[2 * i] : v[2 * (i + 4)];
        S[(2 * i + 1) * 32 + lane] = odd ? v[2 * (7 - i) + 1] : v[2 * (3 - i) + 1];
      }
      pair_sync(bar_id);  // (1) both halves staged and transformed, swap data visible
      if (leader && c + 2 <= nrun - 1) {  // hop-block c of the ring is consumed by both warps: its buffer takes block c+2
        n20_expect_tx(xbar + (c & 1), HOP2 * 4);
        n20_bulk(xs + (c & 1) * HOP2, xrun + (size_t)(c + 2) * HOP2, HOP2 * 4, xbar + (c & 1));
      }
      }  // !INIT
      const unsigned long long seed = INIT ? (a.seed_ptr ? *a.seed_ptr : a.seed) : 0ull;
      const unsigned long long frame_base = ((unsigned long long)b * T + t) * 1025ull;
      float2* tp = a.store_prev ? a.tprev + ((size_t)b * T + t) * HOP2 : nullptr;
      n20_wait(bar, uses & 1);
      ++uses;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int rr = i + 4 * odd;  // slot this warp updates
        float2 Ek, Emk, Ok, Omk;
        const int k = (rr < 4 ? kU0 : kU4) + 64 * rr;
        if (INIT) {
          if (rr == 0 && lane == 0) {
            quad_special_init(Ek, Emk, Ok, Omk, RT[256], mg_s, seed, frame_base);
          } else {
            const float2 rk = RT[k];
            quad_init(Ek, Emk, Ok, Omk, k, cmul(rk, rk), rk, make_float2(-rk.y, -rk.x), mg_s, seed, frame_base);
          }
        } else {
          const float2 oU = So[(2 * i) * 32 + lane], oV = So[(2 * i + 1) * 32 + lane];
          const float2 mU = odd ? v[2 * (i + 4)] : v[2 * i];
          const float2 mV = odd ? v[2 * (3 - i) + 1] : v[2 * (7 - i) + 1];
          Ek = odd ? oU : mU; Emk = odd ? oV : mV; Ok = odd ? mU : oU; Omk = odd ? mV : oV;
          if (rr == 0 && lane == 0) {
            quad_special(Ek, Emk, Ok, Omk, RT[256], tp_s, mg_s, a.mom, USE_PREV, tp);
          } else {
            const float2 rk = RT[k];
            quad_update(Ek, Emk, Ok, Omk, k, cmul(rk, rk), rk, make_float2(-rk.y, -rk.x), tp_s, mg_s, a.mom, USE_PREV, tp);
          }
        }
        if (odd) { v[2 * (i + 4)] = Ok; v[2 * (3 - i) + 1] = Omk; } else { v[2 * i] = Ek; v[2 * (7 - i) + 1] = Emk; }
        X[odd * 256 + (2 * i) * 32 + lane] = odd ? Ek : Ok;      // the partner's transform of this slot goes back
        X[odd * 256 + (2 * i + 1) * 32 + lane] = odd ? Emk : Omk;
      }
      pair_sync(bar_id);  // (2) both updates done: staged rows free, swap-back data visible
      if (leader && t + 1 < te) {
        n20_expect_tx(bar, row_bytes);
        n20_bulk(psm + OFF_MG, a.mag_tf + ((size_t)b * T + t + 1) * FP2, MG_BYTES, bar);
        if (USE_PREV) n20_bulk(psm + OFF_TP, a.tprev + ((size_t)b * T + t + 1) * HOP2, TP_BYTES, bar);
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float2 bU = X[(odd ^ 1) * 256 + (2 * i) * 32 + lane], bV = X[(odd ^ 1) * 256 + (2 * i + 1) * 32 + lane];
        if (odd) { v[2 * i] = bU; v[2 * (7 - i) + 1] = bV; } else { v[2 * (i + 4)] = bU; v[2 * (3 - i) + 1] = bV; }
      }
      if (lane == 0) lane0_unpermute(v);
      // ---- inverse transform -------------------------------------------------------------------------------------------
      __syncwarp();
      inv1_store(lane, v, S);
      __syncwarp();
      inv2_load(lane, v, tw, S);
      __syncwarp();
      inv2_store(lane, v, S);
      __syncwarp();
      inv3_load(lane, v, tw, S);
      // ---- synthesis window + overlap-add of this warp's samples ----------------------------------------------------
      const float2* wn0 = WN0 + odd * 256;
      const float2* wn1 = WN1 + odd * 256;
      if (a.wave != nullptr && c >= 1) {
        const float sc = a.out_scale ? a.out_scale[b] : 1.0f;
        float* wdst = a.wave + (size_t)b * HOP2 * (T - 1) + (size_t)(t - 1) * HOP2 + sub;
        const float* ie = a.inv_env + sub;
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const int o = 4 * (lane + 32 * q);
          const float2 w0 = wn0[lane + 32 * q], w1 = wn1[lane + 32 * q];
          const float2 e = __ldg(reinterpret_cast<const float2*>(ie + o));
          *reinterpret_cast<float2*>(wdst + o) = cscale2(cfma2(v[q], w0, carry[q]), cscale2(e, make_float2(sc, sc)));
          carry[q] = cscale2(v[8 + q], w1);
        }
      } else {
        float* dst = xo + (size_t)c * HOP2 + sub;
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const int o = 4 * (lane + 32 * q);
          const float2 w0 = wn0[lane + 32 * q], w1 = wn1[lane + 32 * q];
          *reinterpret_cast<float2*>(dst + o) = cfma2(v[q], w0, carry[q]);
          carry[q] = cscale2(v[8 + q], w1);
        }
      }
    }
    float* dl = xo + (size_t)nrun * HOP2 + sub;
#pragma unroll
    for (int q = 0; q < 8; ++q) *reinterpret_cast<float2*>(dl + 4 * (lane + 32 * q)) = carry[q];
    __syncwarp();
  }
}

int gl_fast_n2048_warps() { return n2048::PAIRS; }  // run slots (warp pairs) per SM

template <bool USE_PREV, bool INIT>
static int launch_n2048(const GlN2048Args& a, int grid, cudaStream_t st) {
  using namespace n2048;
  const size_t smem = TABLE_BYTES + (size_t)PAIRS * PSMEM;
  B2D_SMEM_OPT_IN(smem, gl_fast_n2048_kernel<USE_PREV, INIT>);
  B2D_CUDA(launch_pdl(gl_fast_n2048_kernel<USE_PREV, INIT>, dim3(grid), dim3(PAIRS * 64), smem, st, a));
  B2D_LAUNCH_CHECK("gl_fast_n2048_kernel");
  return B2D_OK;
}

static GlN2048Args n2048_args(const b2d_plan* p, const float* mag_tf, float2* tprev, const float* xin, float* xout, int B, int T, int n, int R) {
  GlN2048Args a;
  a.mag_tf = mag_tf; a.tprev = tprev; a.xin = xin; a.xout = xout; a.B = B; a.T = T; a.n = n; a.R = R;
  a.tw512 = p->d_tw512; a.rtw = p->d_rtw; a.win = p->d_win; a.winn = p->d_winn; a.inv_env = p->d_inv_env;
  a.mom = 0.f; a.use_prev = 0; a.store_prev = 0; a.wave = nullptr; a.out_scale = nullptr; a.seed = 0; a.seed_ptr = nullptr;
  return a;
}

int launch_gl_fast_n2048(const b2d_plan* p, const float* mag_tf, float2* tprev, const float* xin, float* xout, int B, int T, int n,
                         int R, float mom, int use_prev, int store_prev, float* wave, const float* out_scale, cudaStream_t st) {
  GlN2048Args a = n2048_args(p, mag_tf, tprev, xin, xout, B, T, n, R);
  a.mom = mom; a.use_prev = use_prev; a.store_prev = store_prev; a.wave = wave; a.out_scale = out_scale;
  const int runs = B * R;
  const int grid = runs < p->num_sms ? runs : p->num_sms;
  return use_prev ? launch_n2048<true, false>(a, grid, st) : launch_n2048<false, false>(a, grid, st);
}

// x_0 = istft(mag * angles_0) with in-kernel angle draws (seed != 0) or all-ones angles
int launch_gl_fast_n2048_init(const b2d_plan* p, const float* mag_tf, float* xout, int B, int T, int n, int R, unsigned long long seed,
                              const unsigned long long* seed_ptr, cudaStream_t st) {
  GlN2048Args a = n2048_args(p, mag_tf, nullptr, nullptr, xout, B, T, n, R);
  a.seed = seed; a.seed_ptr = seed_ptr;
  const int runs = B * R;
  const int grid = runs < p->num_sms ? runs : p->num_sms;
  return launch_n2048<false, true>(a, grid, st);
}

}  // namespace b2d
