// gl_fast.cu -- the n_fft = 1024 Griffin-Lim iteration kernel (see gl_fast.cuh for the data flow).
#include <stdlib.h>

#include "gl_fast.cuh"
#include "kernels.cuh"

namespace b2d {

using namespace fast512;

struct GlFastArgs {
  const float* mag_tf;   // [B,T,Fp]
  const float* xin;      // x_k, partial hop-block format (run = one warp's frames)
  const float* xprev;    // x_{k-1}, same format (read only when USE_PREV)
  float* xout;           // x_{k+1}
  int B, T, n, R, Fp;
  const float2* tw512;   // W512^k
  const float2* rtw;     // W1024^k, k = 0..511
  const float* win;      // [1024]
  const float* winn;     // [1024] win / N
  const float* inv_env;  // [512]
  float mom;             // momentum / (1 + momentum), TA functional.py:300
  // last iteration only: hop-blocks interior to a run are final -> written straight to the waveform
  float* wave;             // [B, HOP*(T-1)] or null
  const float* out_scale;  // [B] or null
};

// ---- TMA bulk copy + mbarrier helpers (sm_90+/sm_100a PTX) ------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "W_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra D_%=;\n"
      "bra W_%=;\n"
      "D_%=:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

// per-warp shared-memory staging
constexpr int MAG_ROW_BYTES = (M + 4) * 4;                      // Fp floats = 2064 B
constexpr int WARP_SMEM = XCH * 8 + M * 8 + 2080 + 16;          // init / STFT kernels: exchange | (unused row) | mag row | mbarrier
// iteration kernel: exchange | mag row | mbarrier | ring of two slots, each [x_k hop-block | x_{k-1} hop-block] | two mbarriers
constexpr int IT_OFF_MAG = XCH * 8;
constexpr int IT_OFF_BAR = IT_OFF_MAG + 2080;
constexpr int IT_OFF_RING = IT_OFF_BAR + 16;
constexpr int IT_OFF_XBAR = IT_OFF_RING + 2 * 2 * HOP * 4;
constexpr int IT_WARP_SMEM = IT_OFF_XBAR + 16;                  // 14912 B
#ifndef B2D_IT_WARPS
#define B2D_IT_WARPS 12
#endif
constexpr int IT_WARPS = B2D_IT_WARPS;                                    // one persistent 12-warp CTA per SM (168 registers / thread)
// (Measured, round 2: with `tprev` gone the kernel no longer waits on HBM -- DRAM traffic fell from 436 MB to ~290 MB per
//  launch for 89.3 -> 86.9 us -- it is bound by the shared-memory pipe and issue slots.  Reading the 21 lane twiddles from
//  CTA-wide shared-memory tables instead of 42 registers fits 128 registers and 14 warps per SM, and is SLOWER: 94.8 us at
//  12 warps, 96.9 at 13, 91.9 at 14 (B = 256; 281 -> 306 us at B = 1024): +42 LDS.64 per frame cost more than two more
//  warps give.  tools/tune_gl.py history, gpurun r2.)
// (Measured, round 2, B = 256 x 126 frames = 1536 runs of 21 frames on 148 x 12 warp slots, 83.5 us: other cuts of the same work
//  through an experimental run table -- 3 x 22 + 3 x 20 frames per clip with the short runs on the 56 SMs that carry 11 warps
//  (busiest SM 220 instead of 231 frames): -0.8 %; 5 x 26: +13 %; 4 x 32: +20 %; two rounds of short runs, 9 x 14 ... 21 x 6:
//  +6 ... +40 %.  A warp needs ~4.0 us per frame with 10 - 12 warps on the SM (3.1 us with 7, 2.9 us with 5), so a launch lasts
//  as long as its longest run; SM throughput still rises with every added warp (2.61 frames/us/SM at 10.4 warps, 2.94 at 12:
//  B = 296 fills every slot and reaches 95 % of the HBM roofline figure) and registers cap it at 12.)

// normalised (not yet windowed) sample `is` of interior hop-block js of clip b: the sum of the two partial slots on a run boundary
__device__ __forceinline__ float partial_sample(const float* __restrict__ part, int b, int R, int n, int js, int is) {
  const int r1 = (js - 1) / n, r2 = js / n;
  float v = part[((size_t)(b * R + r1) * (n + 1) + (js - r1 * n)) * HOP + is];
  if (r2 != r1) v += part[((size_t)(b * R + r2) * (n + 1)) * HOP + is];
  return v;
}
// One fused Griffin-Lim iteration (TA:functional/functional.py:316-343) with the momentum term moved to the time domain.
// The reference forms  angles = rebuilt_k - m * rebuilt_{k-1}  from two complex spectrograms; the STFT is linear, so
//   rebuilt_k - m * rebuilt_{k-1} = STFT(x_k - m * x_{k-1}),
// and only the DIRECTION of that difference is used.  The kernel therefore reads the two previous iterates (4 L bytes each)
// instead of reading and writing a complex `tprev` (8 F T bytes each way): per clip and iteration
//   read x_k, x_{k-1} (8 L) + mag (4 F T) + write x_{k+1} (4 L)  =  12 L + 4 F T   instead of   8 L + 20 F T,
// 1.03 MB instead of 1.80 MB at config 2, with one forward and one inverse transform per frame as before.
template <bool USE_PREV, bool EXACT>
__global__ void __launch_bounds__(IT_WARPS * 32, 1) gl_fast512_kernel(const GlFastArgs a) {
  constexpr int WARPS = IT_WARPS;
  // tables (float2 views): WA[256] = inv_env*win (first half), WB[256] = inv_env*win (second half),
  // WN[512] = win/N, RT[512] = W1024^k ; per-warp buffers behind them
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float2* WA = reinterpret_cast<float2*>(smem_raw);
  float2* WB = WA + 256;
  float2* WN = WB + 256;
  float2* RT = WN + 512;
  unsigned char* warp_base = reinterpret_cast<unsigned char*>(RT + 512);

  for (int i = threadIdx.x; i < 256; i += blockDim.x) {
    WA[i] = make_float2(a.inv_env[2 * i] * a.win[2 * i], a.inv_env[2 * i + 1] * a.win[2 * i + 1]);
    WB[i] = make_float2(a.inv_env[2 * i] * a.win[HOP + 2 * i], a.inv_env[2 * i + 1] * a.win[HOP + 2 * i + 1]);
  }
  for (int i = threadIdx.x; i < 512; i += blockDim.x) {
    WN[i] = make_float2(a.winn[2 * i], a.winn[2 * i + 1]);
    RT[i] = a.rtw[i];
  }
  __syncthreads();

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n = a.n, R = a.R, T = a.T;
  const int nruns = a.B * a.R;
  // one CTA per SM; run ids are dealt round-robin over the CTAs first, so every SM carries the same number of busy warps (+-1)
  const int gw0 = warp * (int)gridDim.x + (int)blockIdx.x;
  const int gstep = WARPS * (int)gridDim.x;
  if (gw0 >= nruns) return;
  unsigned char* wsm = warp_base + (size_t)warp * IT_WARP_SMEM;
  float2* S = reinterpret_cast<float2*>(wsm);                       // exchange buffer
  float* mg_s = reinterpret_cast<float*>(wsm + IT_OFF_MAG);         // staged mag row of the current frame
  uint64_t* bar = reinterpret_cast<uint64_t*>(wsm + IT_OFF_BAR);
  float* xs = reinterpret_cast<float*>(wsm + IT_OFF_RING);          // ring slot s: x_k block at xs + s*2*HOP, x_{k-1} block behind it
  uint64_t* xbar = reinterpret_cast<uint64_t*>(wsm + IT_OFF_XBAR);
  if (lane == 0) {
    mbar_init(bar, 1);
    mbar_init(xbar, 1);
    mbar_init(xbar + 1, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncwarp();
  LaneTw tw;
  lane_twiddles(lane, a.tw512, tw);
  // lane 0 owns families 0 and 32: its slots r >= 4 sit 224 bins below the regular lane + 64 r pattern
  const int kU0 = lane, kU4 = lane - (lane == 0 ? 224 : 0);
  const size_t run_stride = (size_t)(n + 1) * HOP;
  const float2 nmom = make_float2(-a.mom, -a.mom);
  constexpr uint32_t ring_bytes = (USE_PREV ? 2u : 1u) * HOP * 4u;
  uint32_t tma_uses = 0, xuse0 = 0, xuse1 = 0;  // completed phases of the three mbarriers (parity tracking across runs)
  // everything above read plan tables only; the iterates (and, for the first iteration, the magnitudes) come from the kernel before
  pdl_wait();
  pdl_trigger();

#pragma unroll 1
  for (int gw = gw0; gw < nruns; gw += gstep) {
  const int b = gw / R, r = gw - b * R;
  const int tb = r * n, te = min(T, tb + n);
  const float* xrun = a.xin + (size_t)(b * R + r) * run_stride;
  const float* prun = a.xprev + (size_t)(b * R + r) * run_stride;
  float* xo = a.xout + (size_t)(b * R + r) * run_stride;
  float2 carry[8];
#pragma unroll
  for (int q = 0; q < 8; ++q) carry[q] = make_float2(0.f, 0.f);

  if (lane == 0) {  // TMA: mag row of the first frame
    mbar_expect_tx(bar, MAG_ROW_BYTES);
    bulk_g2s(mg_s, a.mag_tf + ((size_t)b * T + tb) * a.Fp, MAG_ROW_BYTES, bar);
  }
  const int nrun = te - tb;
  if (lane == 0 && nrun > 1) {  // own-slot hop-blocks 1 .. nrun-1 of both iterates travel by TMA, one frame ahead
    mbar_expect_tx(xbar + 1, ring_bytes);
    bulk_g2s(xs + 2 * HOP, xrun + HOP, HOP * 4, xbar + 1);
    if (USE_PREV) bulk_g2s(xs + 3 * HOP, prun + HOP, HOP * 4, xbar + 1);
  }
#pragma unroll 1
  for (int t = tb; t < te; ++t) {
    const int c = t - tb;
    float2 v[16];
    // ---- stage the frame: d = x_k - m x_{k-1} on hop-blocks t (first half) and t+1 (second half), envelope + window applied ---
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int j = t + h;            // padded hop-block index
      const int cs = c + h;           // slot inside this run
      const float2* wtab = h ? WB : WA;
      if (j == 0 || j == T) {
        // reflect-padded edge of the clip (2 blocks per clip): sample e of the padded block is sample `is` of interior block jsA
        // read in reverse (one sample comes from block jsB).  Every load is issued before the first use: the launch lasts as
        // long as its slowest warp, and the two runs per clip that own an edge used to spend ~6 us here in a serial chain of
        // L2 round trips (16 dependent loop trips with two integer divisions each).
        const int jsA = (j == 0) ? 1 : T - 1, jsB = (j == 0) ? 2 : T - 2;
        auto slots = [&](const float* part, int js, const float*& p1, const float*& p2) {  // block js = slot of run (js-1)/n (+ slot 0 of run js/n)
          const int r1 = (js - 1) / n, r2 = js / n;
          p1 = part + ((size_t)(b * R + r1) * (n + 1) + (js - r1 * n)) * HOP;
          p2 = (r2 != r1) ? part + ((size_t)(b * R + r2) * (n + 1)) * HOP : nullptr;
        };
        const float *xa1, *xa2, *xb1, *xb2, *pa1 = nullptr, *pa2 = nullptr, *pb1 = nullptr, *pb2 = nullptr;
        slots(a.xin, jsA, xa1, xa2);
        slots(a.xin, jsB, xb1, xb2);
        if (USE_PREV) { slots(a.xprev, jsA, pa1, pa2); slots(a.xprev, jsB, pb1, pb2); }
        const float* wh = a.win + h * HOP;
        float xr[16], pr[16], sc[16];
#pragma unroll
        for (int u = 0; u < 16; ++u) {
          const int e = 2 * (lane + 32 * (u >> 1)) + (u & 1);
          const bool odd_one = (j == 0) ? (e == 0) : (e == HOP - 1);  // the one sample that comes from block jsB
          const int is = (j == 0) ? (e == 0 ? 0 : HOP - e) : (e == HOP - 1 ? HOP - 1 : HOP - 2 - e);
          const float* s1 = odd_one ? xb1 : xa1;
          const float* s2 = odd_one ? xb2 : xa2;
          float x = s1[is];
          if (s2) x += s2[is];
          xr[u] = x;
          if (USE_PREV) {
            const float* t1 = odd_one ? pb1 : pa1;
            const float* t2 = odd_one ? pb2 : pa2;
            float pv = t1[is];
            if (t2) pv += t2[is];
            pr[u] = pv;
          }
          sc[u] = a.inv_env[is];
        }
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const int e = 2 * (lane + 32 * q);
          float x0 = xr[2 * q], x1 = xr[2 * q + 1];
          if (USE_PREV) { x0 = fmaf(-a.mom, pr[2 * q], x0); x1 = fmaf(-a.mom, pr[2 * q + 1], x1); }
          v[8 * h + q] = make_float2(x0 * sc[2 * q] * wh[e], x1 * sc[2 * q + 1] * wh[e + 1]);
        }
      } else if (cs >= 1 && cs <= nrun - 1) {  // interior block of this run: already in the shared-memory ring
        if (h == 1) {  // (h == 0: the same block was awaited one frame ago)
          if (cs & 1) { mbar_wait(xbar + 1, xuse1 & 1); ++xuse1; } else { mbar_wait(xbar, xuse0 & 1); ++xuse0; }
        }
        const float2* src = reinterpret_cast<const float2*>(xs + (cs & 1) * 2 * HOP);
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          float2 xv = src[lane + 32 * q];
          if (USE_PREV) xv = cfma2(src[256 + lane + 32 * q], nmom, xv);
          v[8 * h + q] = cscale2(xv, wtab[lane + 32 * q]);
        }
      } else {
        // block on a run boundary: the sum of this run's slot and the neighbouring run's slot
        const size_t o1 = (size_t)cs * HOP;
        const ptrdiff_t o2 = (cs == 0) ? -(ptrdiff_t)run_stride + (ptrdiff_t)n * HOP : (ptrdiff_t)run_stride;  // previous run's last slot / next run's slot 0
        const bool two = (cs == 0) || (j == te);
        const float2* p1 = reinterpret_cast<const float2*>(xrun + o1);
        const float2* q1 = reinterpret_cast<const float2*>(prun + o1);
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          float2 xv = p1[lane + 32 * q];
          if (two) xv = cadd(xv, reinterpret_cast<const float2*>(xrun + o2)[lane + 32 * q]);
          if (USE_PREV) {
            float2 pv = q1[lane + 32 * q];
            if (two) pv = cadd(pv, reinterpret_cast<const float2*>(prun + o2)[lane + 32 * q]);
            xv = cfma2(pv, nmom, xv);
          }
          v[8 * h + q] = cscale2(xv, wtab[lane + 32 * q]);
        }
      }
    }
    {  // block c is consumed: its ring slot takes block c+2 while this frame computes
      __syncwarp();
      if (lane == 0 && c + 2 <= nrun - 1) {
        float* slot = xs + (c & 1) * 2 * HOP;
        mbar_expect_tx(xbar + (c & 1), ring_bytes);
        bulk_g2s(slot, xrun + (size_t)(c + 2) * HOP, HOP * 4, xbar + (c & 1));
        if (USE_PREV) bulk_g2s(slot + HOP, prun + (size_t)(c + 2) * HOP, HOP * 4, xbar + (c & 1));
      }
    }
    // ---- forward FFT ------------------------------------------------------------------------------------
    __syncwarp();
    fwd1_store(lane, v, tw, S);
    __syncwarp();
    fwd2_load(lane, v, S);
    __syncwarp();
    fwd2_store(lane, v, tw, S);
    __syncwarp();
    fwd3_load(lane, v, S);
    // ---- projection to unit modulus, x mag (the mag row was bulk-copied into smem one frame ahead) ---------
    mbar_wait(bar, tma_uses & 1);
    ++tma_uses;
    if (lane == 0) lane0_permute(v);
#pragma unroll
    for (int rr = 0; rr < 8; ++rr) {
      const int k = (rr < 4 ? kU0 : kU4) + 64 * rr;
      float2& U = v[2 * rr];
      float2& V = v[2 * (7 - rr) + 1];
      if (rr == 0 && lane == 0) special_project<EXACT>(U, V, mg_s[0], mg_s[M], mg_s[256]);
      else pair_project<EXACT>(U, V, RT[k], mg_s[k], mg_s[M - k]);
    }
    __syncwarp();  // every lane is done reading the staged row
    if (lane == 0 && t + 1 < te) {  // TMA: mag row of the next frame lands while this frame's inverse FFT runs
      mbar_expect_tx(bar, MAG_ROW_BYTES);
      bulk_g2s(mg_s, a.mag_tf + ((size_t)b * T + t + 1) * a.Fp, MAG_ROW_BYTES, bar);
    }
    if (lane == 0) lane0_unpermute(v);
    // ---- inverse FFT ------------------------------------------------------------------------------------
    __syncwarp();
    inv1_store(lane, v, S);
    __syncwarp();
    inv2_load(lane, v, tw, S);
    __syncwarp();
    inv2_store(lane, v, S);
    __syncwarp();
    inv3_load(lane, v, tw, S);
    // ---- synthesis window + overlap-add: block c = carry + first half ; carry = second half --------------
    if (a.wave != nullptr && c >= 1) {
      // last iteration, block interior to this run: both contributions are here, so this is the final sample --
      // divide by the window envelope, apply the clip's scale and write the waveform directly (skips the stitch)
      const float sc = a.out_scale ? a.out_scale[b] : 1.0f;
      float2* wdst = reinterpret_cast<float2*>(a.wave + (size_t)b * HOP * (T - 1) + (size_t)(t - 1) * HOP);
      const float2* ie = reinterpret_cast<const float2*>(a.inv_env);
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const float2 w0 = WN[lane + 32 * q], w1 = WN[256 + lane + 32 * q];
        const float2 e = ie[lane + 32 * q];
        wdst[lane + 32 * q] = cscale2(cfma2(v[q], w0, carry[q]), cscale2(e, make_float2(sc, sc)));
        carry[q] = cscale2(v[8 + q], w1);
      }
    } else {
      float2* dst = reinterpret_cast<float2*>(xo + (size_t)c * HOP);
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const float2 w0 = WN[lane + 32 * q], w1 = WN[256 + lane + 32 * q];
        dst[lane + 32 * q] = cfma2(v[q], w0, carry[q]);
        carry[q] = cscale2(v[8 + q], w1);
      }
    }
  }
  float2* dst = reinterpret_cast<float2*>(xo + (size_t)(te - tb) * HOP);
#pragma unroll
  for (int q = 0; q < 8; ++q) dst[lane + 32 * q] = carry[q];
  __syncwarp();
  }  // next run of this warp
}

// ------------------------------------------------------------------------------------------------
// x_0 = istft(mag * angles_0) for rand_init (in-kernel counter-based draws) or all-ones angles:
// the inverse half of the iteration kernel (merge -> inverse FFT -> overlap-add), mag rows via TMA.
// ------------------------------------------------------------------------------------------------
struct GlInitArgs {
  const float* mag_tf;
  float* xout;
  int B, T, n, R, Fp;
  const float2* tw512;
  const float2* rtw;
  const float* winn;
  unsigned long long seed;
  const unsigned long long* seed_ptr;
};

template <int WARPS>
__global__ void __launch_bounds__(WARPS * 32, (WARPS > 8 ? 1 : 2)) gl_fast512_init_kernel(const GlInitArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float2* WN = reinterpret_cast<float2*>(smem_raw);
  float2* RT = WN + 512;
  unsigned char* warp_base = reinterpret_cast<unsigned char*>(RT + 512);
  for (int i = threadIdx.x; i < 512; i += blockDim.x) {
    WN[i] = make_float2(a.winn[2 * i], a.winn[2 * i + 1]);
    RT[i] = a.rtw[i];
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int gw = blockIdx.x * (int)(blockDim.x >> 5) + warp;  // CTAs of blockDim.x / 32 <= WARPS warps
  if (gw >= a.B * a.R) return;
  const int b = gw / a.R, r = gw - b * a.R;
  const int n = a.n, R = a.R, T = a.T;
  const int tb = r * n, te = min(T, tb + n);
  unsigned char* wsm = warp_base + (size_t)warp * WARP_SMEM;
  float2* S = reinterpret_cast<float2*>(wsm);
  float* mg_s = reinterpret_cast<float*>(wsm + XCH * 8 + M * 8);
  uint64_t* bar = reinterpret_cast<uint64_t*>(wsm + XCH * 8 + M * 8 + 2080);
  if (lane == 0) {
    mbar_init(bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncwarp();
  LaneTw tw;
  lane_twiddles(lane, a.tw512, tw);
  const int kU0 = lane, kU4 = lane - (lane == 0 ? 224 : 0);
  float* xo = a.xout + (size_t)(b * R + r) * (size_t)(n + 1) * HOP;
  float2 carry[8];
#pragma unroll
  for (int q = 0; q < 8; ++q) carry[q] = make_float2(0.f, 0.f);
  pdl_wait();  // magnitudes (and the seed) come from the kernels before
  pdl_trigger();
  if (lane == 0) {
    mbar_expect_tx(bar, MAG_ROW_BYTES);
    bulk_g2s(mg_s, a.mag_tf + ((size_t)b * T + tb) * a.Fp, MAG_ROW_BYTES, bar);
  }
#pragma unroll 1
  for (int t = tb; t < te; ++t) {
    const int c = t - tb;
    float2 v[16];
    const unsigned long long base = ((unsigned long long)b * T + t) * (M + 1);
    const unsigned long long seed = a.seed_ptr ? *a.seed_ptr : a.seed;
    mbar_wait(bar, c & 1);
#pragma unroll
    for (int rr = 0; rr < 8; ++rr) {
      const int k = (rr < 4 ? kU0 : kU4) + 64 * rr;
      float2& U = v[2 * rr];
      float2& V = v[2 * (7 - rr) + 1];
      const float2 one = make_float2(1.f, 0.f);
      if (rr == 0 && lane == 0) {
        const float2 a0 = seed ? rand_angle(seed, base) : one, aM = seed ? rand_angle(seed, base + M) : one;
        const float2 a256 = seed ? rand_angle(seed, base + 256) : one;
        const float y0 = mg_s[0] * a0.x, yM = mg_s[M] * aM.x, m256 = mg_s[256];
        U = make_float2(y0 + yM, y0 - yM);
        V = make_float2(2.0f * m256 * a256.x, -2.0f * m256 * a256.y);
      } else {
        const float2 ak = seed ? rand_angle(seed, base + k) : one, amk = seed ? rand_angle(seed, base + (M - k)) : one;
        const float mk = mg_s[k], mmk = mg_s[M - k];
        irfft_merge(make_float2(mk * ak.x, mk * ak.y), make_float2(mmk * amk.x, mmk * amk.y), RT[k], U, V);
      }
    }
    __syncwarp();
    if (lane == 0 && t + 1 < te) {
      mbar_expect_tx(bar, MAG_ROW_BYTES);
      bulk_g2s(mg_s, a.mag_tf + ((size_t)b * T + t + 1) * a.Fp, MAG_ROW_BYTES, bar);
    }
    if (lane == 0) lane0_unpermute(v);
    inv1_store(lane, v, S);
    __syncwarp();
    inv2_load(lane, v, tw, S);
    __syncwarp();
    inv2_store(lane, v, S);
    __syncwarp();
    inv3_load(lane, v, tw, S);
    float2* dst = reinterpret_cast<float2*>(xo + (size_t)c * HOP);
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const float2 w0 = WN[lane + 32 * q], w1 = WN[256 + lane + 32 * q];
      dst[lane + 32 * q] = cfma2(v[q], w0, carry[q]);
      carry[q] = cscale2(v[8 + q], w1);
    }
  }
  float2* dst = reinterpret_cast<float2*>(xo + (size_t)(te - tb) * HOP);
#pragma unroll
  for (int q = 0; q < 8; ++q) dst[lane + 32 * q] = carry[q];
}

template <int WMAX>  // kernel instantiated for CTAs of up to WMAX warps, launched with `warps` of them
static int launch_init_w(const GlInitArgs& a, int warps, cudaStream_t st) {
  const size_t smem = sizeof(float2) * 1024 + (size_t)warps * WARP_SMEM;
  B2D_SMEM_OPT_IN(sizeof(float2) * 1024 + (size_t)WMAX * WARP_SMEM, gl_fast512_init_kernel<WMAX>);
  B2D_CUDA(launch_pdl(gl_fast512_init_kernel<WMAX>, dim3((a.B * a.R + warps - 1) / warps), dim3(warps * 32), smem, st, a));
  B2D_LAUNCH_CHECK("gl_fast512_init_kernel");
  return B2D_OK;
}

int launch_gl_fast512_init(const b2d_plan* p, const float* mag_tf, float* xout, int B, int T, int n, int R, unsigned long long seed,
                           const unsigned long long* seed_ptr, cudaStream_t st) {
  GlInitArgs a;
  a.mag_tf = mag_tf; a.xout = xout; a.B = B; a.T = T; a.n = n; a.R = R; a.Fp = p->Fp;
  a.tw512 = p->d_tw512; a.rtw = p->d_rtw; a.winn = p->d_winn; a.seed = seed; a.seed_ptr = seed_ptr;
  // a launch lasts as long as its busiest SM: when the runs fit one CTA per SM, size the CTAs so that they do (config 2: 1536 runs
  // = 140 CTAs of 11 warps instead of 192 CTAs of 8, which put 16 warps on 44 of the 148 SMs and 8 on the rest)
  const int runs = B * R;
  const int per_sm = (runs + p->num_sms - 1) / p->num_sms;
  if (per_sm > 8 && per_sm <= 12) return launch_init_w<12>(a, per_sm, st);
  return launch_init_w<8>(a, 8, st);
}

// ------------------------------------------------------------------------------------------------
// K1+K2 for n_fft = 1024, hop = 512: one warp per frame -- window, register FFT, |.|, mel, log1p.
// ------------------------------------------------------------------------------------------------
struct StftFastArgs {
  const float* wave;
  const float* inv_scale;
  int B, L, T, n_mels;
  const float2* tw512;
  const float2* rtw;
  const float* win;
  const float* seg_w;     // [2][seg_pad][4] mel column segments of <= 8 bins (see b2d_plan)
  const int* seg_lo;      // [seg_pad]
  const int* seg_first;   // [n_mels + 1]
  int seg_pad;
  float* logmel_bt;
  int exact_sqrt;   // B2D_PLAN_EXACT_SQRT: IEEE square root for |X|
  int exact_div;    // B2D_PLAN_EXACT_PEAK_DIV: x / peak as a division
};

__device__ __forceinline__ float sqrt_fast(float x) {
  float r;
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}

// One warp per frame; the frame's 4 KB of samples arrive by TMA bulk copy into a two-slot per-warp ring one frame
// ahead (interior frames are contiguous in the clip); the two reflect-padded frames at the clip edges are staged
// by the warp itself.
constexpr int STFT_WARPS = 16;
constexpr int STFT_WSMEM = XCH * 8 + 2 * N * 4 + 16;  // exchange | two frame buffers | two mbarriers

__global__ void __launch_bounds__(STFT_WARPS * 32, 1) stft_fast512_kernel(const StftFastArgs a) {
  constexpr int WARPS = STFT_WARPS;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float2* WIN = reinterpret_cast<float2*>(smem_raw);  // [512] plain window pairs
  float2* RT = WIN + 512;
  float4* SEGW = reinterpret_cast<float4*>(RT + 512);  // [2][seg_pad]
  int* SEGLO = reinterpret_cast<int*>(SEGW + 2 * a.seg_pad);
  int* SEGF = SEGLO + a.seg_pad;                      // [n_mels + 1], padded to a multiple of 4 ints
  unsigned char* warp_base = reinterpret_cast<unsigned char*>(SEGF + ((a.n_mels + 4) & ~3));
  for (int i = threadIdx.x; i < 512; i += blockDim.x) {
    WIN[i] = make_float2(a.win[2 * i], a.win[2 * i + 1]);
    RT[i] = a.rtw[i];
  }
  for (int i = threadIdx.x; i < 2 * a.seg_pad; i += blockDim.x) SEGW[i] = reinterpret_cast<const float4*>(a.seg_w)[i];
  for (int i = threadIdx.x; i < a.seg_pad; i += blockDim.x) SEGLO[i] = a.seg_lo[i];
  for (int i = threadIdx.x; i <= a.n_mels; i += blockDim.x) SEGF[i] = a.seg_first[i];
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  unsigned char* wsm = warp_base + (size_t)warp * STFT_WSMEM;
  float2* S = reinterpret_cast<float2*>(wsm);
  float* Sf = reinterpret_cast<float*>(S);
  float* fbuf = reinterpret_cast<float*>(wsm + XCH * 8);  // [2][1024]
  uint64_t* fbar = reinterpret_cast<uint64_t*>(wsm + XCH * 8 + 2 * N * 4);
  if (lane == 0) {
    mbar_init(fbar, 1);
    mbar_init(fbar + 1, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncwarp();
  LaneTw tw;
  lane_twiddles(lane, a.tw512, tw);
  const int kU0 = lane, kU4 = lane - (lane == 0 ? 224 : 0);
  const unsigned nframes = (unsigned)a.B * (unsigned)a.T;  // < 2^31 (checked by the launcher): 32-bit index math throughout
  const unsigned stride = gridDim.x * WARPS;
  const bool tma_ok = (a.L % 4) == 0;  // 16-byte aligned clip rows
  pdl_wait();  // the waveform and the peaks come from the kernels before
  pdl_trigger();
  // frame (b, t) is "interior" when its 1024 samples lie inside the clip: then it is one contiguous, aligned 4 KB read
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
    if (interior(b, t, src)) { mbar_expect_tx(fbar, N * 4); bulk_g2s(fbuf, src, N * 4, fbar); }
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
      if (interior(bn, tn, src)) { mbar_expect_tx(fbar + (slot ^ 1), N * 4); bulk_g2s(fbuf + (slot ^ 1) * N, src, N * 4, fbar + (slot ^ 1)); }
    }
    if (cur_tma) {
      if (slot) { mbar_wait(fbar + 1, use1 & 1); ++use1; } else { mbar_wait(fbar, use0 & 1); ++use0; }
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
    float2 v[16];
    {
      const float2* c2 = reinterpret_cast<const float2*>(cur);
#pragma unroll
      for (int q = 0; q < 16; ++q) {
        const float2 xv = c2[lane + 32 * q];
        const float2 wv = WIN[lane + 32 * q];
        const float2 xn = a.exact_div ? make_float2(xv.x / pk, xv.y / pk) : cscale2(xv, make_float2(rsc, rsc));
        v[q] = cscale2(xn, wv);
      }
    }
    __syncwarp();
    fwd1_store(lane, v, tw, S);
    __syncwarp();
    fwd2_load(lane, v, S);
    __syncwarp();
    fwd2_store(lane, v, tw, S);
    __syncwarp();
    fwd3_load(lane, v, S);
    __syncwarp();  // exchange buffer is re-used for the magnitudes
    if (lane == 0) lane0_permute(v);
#pragma unroll
    for (int rr = 0; rr < 8; ++rr) {
      const int k = (rr < 4 ? kU0 : kU4) + 64 * rr;
      const float2 U = v[2 * rr], V = v[2 * (7 - rr) + 1];
      if (rr == 0 && lane == 0) {
        Sf[0] = fabsf(U.x + U.y);
        Sf[M] = fabsf(U.x - U.y);
        Sf[256] = a.exact_sqrt ? sqrtf(V.x * V.x + V.y * V.y) : sqrt_fast(V.x * V.x + V.y * V.y);
      } else {
        float2 xk, xmk;
        rfft_split(U, V, RT[k], xk, xmk);
        const float s1 = xk.x * xk.x + xk.y * xk.y, s2 = xmk.x * xmk.x + xmk.y * xmk.y;
        Sf[k] = a.exact_sqrt ? sqrtf(s1) : sqrt_fast(s1);
        Sf[M - k] = a.exact_sqrt ? sqrtf(s2) : sqrt_fast(s2);
      }
    }
    if (lane < 8) Sf[M + 1 + lane] = 0.f;  // the zero-weight taps of a column's last segment read up to 7 bins past the Nyquist bin
    __syncwarp();
    // mel projection: lane s accumulates segment s (<= 8 consecutive bins of one mel column) ...
    float* P = Sf + 576;
    for (int s0 = 0; s0 < a.seg_pad; s0 += 32) {
      const int sg = s0 + lane;
      const float* mp = Sf + SEGLO[sg];
      const float4 w0 = SEGW[sg], w1 = SEGW[a.seg_pad + sg];
      float acc = mp[0] * w0.x;
      acc = fmaf(mp[1], w0.y, acc); acc = fmaf(mp[2], w0.z, acc); acc = fmaf(mp[3], w0.w, acc);
      acc = fmaf(mp[4], w1.x, acc); acc = fmaf(mp[5], w1.y, acc); acc = fmaf(mp[6], w1.z, acc); acc = fmaf(mp[7], w1.w, acc);
      P[sg] = acc;
    }
    __syncwarp();
    // ... then every mel column sums its segments in order
    for (int m = lane; m < a.n_mels; m += 32) {
      float acc = 0.f;
      for (int sg = SEGF[m]; sg < SEGF[m + 1]; ++sg) acc += P[sg];
      a.logmel_bt[(size_t)f * a.n_mels + m] = log1pf(acc);
    }
    __syncwarp();
  }
}

int launch_stft_fast512(const b2d_plan* p, const float* wave, const float* inv_scale, int B, int L, float* logmel_bt,
                        cudaStream_t st) {
  StftFastArgs a;
  a.wave = wave; a.inv_scale = inv_scale; a.B = B; a.L = L; a.T = 1 + L / p->hop; a.n_mels = p->n_mels;
  a.tw512 = p->d_tw512; a.rtw = p->d_rtw; a.win = p->d_win;
  a.seg_w = p->d_seg_w; a.seg_lo = p->d_seg_lo; a.seg_first = p->d_seg_first; a.seg_pad = p->mel_seg_pad;
  a.logmel_bt = logmel_bt;
  a.exact_sqrt = (p->flags & B2D_PLAN_EXACT_SQRT) ? 1 : 0;
  a.exact_div = (p->flags & B2D_PLAN_EXACT_PEAK_DIV) ? 1 : 0;
  constexpr int W = STFT_WARPS;
  B2D_REQUIRE(p->mel_seg_pad <= 320, B2D_ERR_UNSUPPORTED, "mel filterbank too dense for the n_fft = 1024 fast kernel");
  const size_t smem = sizeof(float2) * 1024 + (size_t)p->mel_seg_pad * 36 + sizeof(int) * ((p->n_mels + 4) & ~3) + (size_t)W * STFT_WSMEM;
  B2D_SMEM_OPT_IN(smem, stft_fast512_kernel);
  const size_t nframes = (size_t)B * a.T;
  const size_t want = (nframes + W - 1) / W;
  const int grid = (int)(want < (size_t)p->num_sms ? want : (size_t)p->num_sms);
  B2D_CUDA(launch_pdl(stft_fast512_kernel, dim3(grid), dim3(W * 32), smem, st, a));
  B2D_LAUNCH_CHECK("stft_fast512_kernel");
  return B2D_OK;
}

int gl_fast_warps_per_sm() { return IT_WARPS; }

template <bool USE_PREV, bool EXACT>
static int launch_iteration(const GlFastArgs& a, int num_sms, cudaStream_t st) {
  const size_t smem = sizeof(float2) * (256 + 256 + 512 + 512) + (size_t)IT_WARPS * IT_WARP_SMEM;
  B2D_SMEM_OPT_IN(smem, gl_fast512_kernel<USE_PREV, EXACT>);
  const int runs = a.B * a.R;
  B2D_CUDA(launch_pdl(gl_fast512_kernel<USE_PREV, EXACT>, dim3(runs < num_sms ? runs : num_sms), dim3(IT_WARPS * 32), smem, st, a));
  B2D_LAUNCH_CHECK("gl_fast512_kernel");
  return B2D_OK;
}
template <bool EXACT>
static int launch_iteration_p(const GlFastArgs& a, int use_prev, int num_sms, cudaStream_t st) {
  return use_prev ? launch_iteration<true, EXACT>(a, num_sms, st) : launch_iteration<false, EXACT>(a, num_sms, st);
}

int launch_gl_fast512(const b2d_plan* p, const float* mag_tf, const float* xin, const float* xprev, float* xout, int B, int T,
                      int n, int R, float mom, int use_prev, float* wave, const float* out_scale, cudaStream_t st) {
  GlFastArgs a;
  a.mag_tf = mag_tf; a.xin = xin; a.xprev = use_prev ? xprev : xin; a.xout = xout;
  a.B = B; a.T = T; a.n = n; a.R = R; a.Fp = p->Fp;
  a.tw512 = p->d_tw512; a.rtw = p->d_rtw; a.win = p->d_win; a.winn = p->d_winn; a.inv_env = p->d_inv_env;
  a.mom = mom; a.wave = wave; a.out_scale = out_scale;
  const bool exact = (p->flags & B2D_PLAN_EXACT_UNIT) != 0;
  if (exact) return launch_iteration_p<true>(a, use_prev, p->num_sms, st);
  return launch_iteration_p<false>(a, use_prev, p->num_sms, st);
}

}  // namespace b2d
