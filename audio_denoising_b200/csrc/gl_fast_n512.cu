// gl_fast_n512.cu -- Griffin-Lim iteration for n_fft = 512 (hop 256) on the register FFT of gl_fast.cuh.
//
// Two consecutive real frames a = frame t, b = frame t+1 are packed into ONE 512-point complex transform
// z = a + i b (the classic two-for-one trick): the forward / inverse passes are exactly the 8 x 8 x 8 register
// stages of the n_fft = 1024 kernel, the pairing of bin k with 512 - k is lane-local as before, and no real-FFT
// twiddles are needed:  2 A[k] = Z[k] + conj Z[512-k],  2 B[k] = -i (Z[k] - conj Z[512-k])  (pair2_update).
// A warp therefore walks its run two frames at a time; the overlap-add of the pair and the carry into the next
// pair stay in registers.  tprev / mag rows of the two frames are adjacent in memory: one TMA bulk copy each.
#include <stdlib.h>

#include "gl_fast.cuh"
#include "kernels.cuh"

namespace b2d {

using namespace fast512;

namespace n512 {
constexpr int HOPN = 256;   // hop = bins per frame (one-sided 257 with DC/Nyquist packed)
constexpr int FPN = 260;    // frame-layout row stride of the magnitudes
constexpr int WARPS = 12;
constexpr int WSMEM = XCH * 8 + 2 * HOPN * 8 + 2 * FPN * 4 + 16;  // exchange | 2 tprev rows | 2 mag rows | mbarrier = 10800 B
}  // namespace n512

struct GlN512Args {
  const float* mag_tf;   // [B,T,260]
  float2* tprev;         // [B,T,256]  (2 x rebuilt, bin 0 = (DC, Nyquist))
  const float* xin;      // partial hop-block format, hop 256
  float* xout;
  int B, T, n, R;
  const float2* tw512;
  const float* win;      // [512]
  const float* winn;     // [512] win / 512
  const float* inv_env;  // [256]
  float mom;
  int use_prev, store_prev;
};

__device__ __forceinline__ uint32_t n5_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void n5_mbar_init(uint64_t* bar) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(n5_u32(bar)) : "memory");
}
__device__ __forceinline__ void n5_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(n5_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void n5_bulk(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(n5_u32(dst)), "l"(src),
               "r"(bytes), "r"(n5_u32(bar))
               : "memory");
}
__device__ __forceinline__ void n5_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "W_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra D_%=;\n"
      "bra W_%=;\n"
      "D_%=:\n"
      "}\n" ::"r"(n5_u32(bar)),
      "r"(parity)
      : "memory");
}

template <bool USE_PREV>
__global__ void __launch_bounds__(n512::WARPS * 32, 1) gl_fast_n512_kernel(const GlN512Args a) {
  using namespace n512;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float* W1 = reinterpret_cast<float*>(smem_raw);  // analysis window, first half  [256]
  float* W2 = W1 + HOPN;                           // analysis window, second half [256]
  float* WNn = W2 + HOPN;                          // synthesis window / 512       [512]
  float* IE = WNn + 2 * HOPN;                      // 1 / envelope                 [256]
  unsigned char* warp_base = reinterpret_cast<unsigned char*>(IE + HOPN);
  for (int i = threadIdx.x; i < HOPN; i += blockDim.x) {
    W1[i] = a.win[i];
    W2[i] = a.win[HOPN + i];
    IE[i] = a.inv_env[i];
  }
  for (int i = threadIdx.x; i < 2 * HOPN; i += blockDim.x) WNn[i] = a.winn[i];
  __syncthreads();

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n = a.n, R = a.R, T = a.T;
  const int nruns = a.B * R;
  const int gw0 = warp * (int)gridDim.x + (int)blockIdx.x;  // runs dealt round-robin over the CTAs (one CTA per SM)
  const int gstep = WARPS * (int)gridDim.x;
  if (gw0 >= nruns) return;
  unsigned char* wsm = warp_base + (size_t)warp * WSMEM;
  float2* S = reinterpret_cast<float2*>(wsm);
  float* Sf = reinterpret_cast<float*>(wsm);
  float2* tp_s = reinterpret_cast<float2*>(wsm + XCH * 8);            // rows of frames a, b
  float* mg_s = reinterpret_cast<float*>(wsm + XCH * 8 + 2 * HOPN * 8);
  uint64_t* bar = reinterpret_cast<uint64_t*>(wsm + XCH * 8 + 2 * HOPN * 8 + 2 * FPN * 4);
  if (lane == 0) {
    n5_mbar_init(bar);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncwarp();
  LaneTw tw;
  lane_twiddles(lane, a.tw512, tw);
  const int kU0 = lane, kU4 = lane - (lane == 0 ? 224 : 0);
  const size_t run_stride = (size_t)(n + 1) * HOPN;
  uint32_t uses = 0;
  pdl_wait();  // the prologue above read plan tables only (common.cuh: programmatic dependent launch)
  pdl_trigger();

#pragma unroll 1
  for (int gw = gw0; gw < nruns; gw += gstep) {
    const int b = gw / R, r = gw - b * R;
    const int tb = r * n, te = min(T, tb + n);
    const int nrun = te - tb;
    const float* xrun = a.xin + (size_t)(b * R + r) * run_stride;
    float* xo = a.xout + (size_t)(b * R + r) * run_stride;
    float carry[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) carry[q] = 0.f;

    auto issue_rows = [&](int t) {  // TMA: tprev / mag rows of frames t and (if present) t+1 are adjacent in memory
      const uint32_t nfr = (t + 1 < te) ? 2u : 1u;
      n5_expect_tx(bar, nfr * (FPN * 4 + (USE_PREV ? HOPN * 8 : 0)));
      n5_bulk(mg_s, a.mag_tf + ((size_t)b * T + t) * FPN, nfr * FPN * 4, bar);
      if (USE_PREV) n5_bulk(tp_s, a.tprev + ((size_t)b * T + t) * HOPN, nfr * HOPN * 8, bar);
    };
    auto load_block = [&](int j, int slot, float* x) {  // envelope-normalised samples lane + 32 q of padded hop-block j
      if (j == 0 || j == T) {
        __syncwarp();
        stage_reflect_wide<HOPN>(a.xin, nullptr, 0.f, a.inv_env, nullptr, b, R, n, T, j, Sf, lane);
        __syncwarp();
#pragma unroll
        for (int q = 0; q < 8; ++q) x[q] = Sf[lane + 32 * q];
        __syncwarp();
      } else {
        const float* p1 = xrun + (size_t)slot * HOPN;
        const float* p2 = nullptr;
        if (slot == 0) p2 = xrun - run_stride + (size_t)n * HOPN;  // previous run, last slot
        else if (slot == nrun) p2 = xrun + run_stride;             // next run, slot 0
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          float v = p1[lane + 32 * q];
          if (p2) v += p2[lane + 32 * q];
          x[q] = v * IE[lane + 32 * q];
        }
      }
    };

    if (lane == 0) issue_rows(tb);
#pragma unroll 1
    for (int t = tb; t < te; t += 2) {
      const int c = t - tb;
      const bool has_b = (t + 1 < te);
      float2 v[16];
      {
        float x0[8], x1[8], x2[8];
        load_block(t, c, x0);
        load_block(t + 1, c + 1, x1);
        if (has_b) {
          load_block(t + 2, c + 2, x2);
        } else {
#pragma unroll
          for (int q = 0; q < 8; ++q) x2[q] = 0.f;
        }
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const float w1 = W1[lane + 32 * q], w2 = W2[lane + 32 * q];
          v[q] = make_float2(x0[q] * w1, has_b ? x1[q] * w1 : 0.f);
          v[8 + q] = make_float2(x1[q] * w2, x2[q] * w2);
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
      // ---- spectral update of both frames ------------------------------------------------------------
      float2* tpa = a.tprev + ((size_t)b * T + t) * HOPN;
      float2* tpb = tpa + HOPN;
      n5_wait(bar, uses & 1);
      ++uses;
      const float2* pa_s = tp_s;
      const float2* pb_s = tp_s + HOPN;
      const float* ma_s = mg_s;
      const float* mb_s = mg_s + FPN;
      const float2 zero2 = make_float2(0.f, 0.f);
      if (lane == 0) lane0_permute(v);
#pragma unroll
      for (int rr = 0; rr < 8; ++rr) {
        const int kU = (rr < 4 ? kU0 : kU4) + 64 * rr;
        float2& U = v[2 * rr];
        float2& V = v[2 * (7 - rr) + 1];
        if (rr == 0 && lane == 0) {
          float2 x0a, x0b;
          special2_update(U, V, USE_PREV ? pa_s[0] : zero2, (USE_PREV && has_b) ? pb_s[0] : zero2, ma_s[0], ma_s[HOPN],
                          has_b ? mb_s[0] : 0.f, has_b ? mb_s[HOPN] : 0.f, a.mom, USE_PREV, x0a, x0b);
          if (a.store_prev) {
            __stcs(tpa, x0a);
            if (has_b) __stcs(tpb, x0b);
          }
        } else {
          const bool swap = (rr >= 4) && (lane != 0);  // U holds the bin above 256: its mirror carries the bin index
          const int kb = swap ? 512 - kU : kU;
          float2 lo = swap ? V : U, hi = swap ? U : V;
          float2 xa, xb;
          pair2_update(lo, hi, USE_PREV ? pa_s[kb] : zero2, (USE_PREV && has_b) ? pb_s[kb] : zero2, ma_s[kb], has_b ? mb_s[kb] : 0.f,
                       a.mom, USE_PREV, xa, xb);
          U = swap ? hi : lo;
          V = swap ? lo : hi;
          if (a.store_prev) {
            __stcs(tpa + kb, xa);
            if (has_b) __stcs(tpb + kb, xb);
          }
        }
      }
      __syncwarp();
      if (lane == 0 && t + 2 < te) issue_rows(t + 2);
      if (lane == 0) lane0_unpermute(v);
      // ---- inverse FFT ---------------------------------------------------------------------------------
      __syncwarp();
      inv1_store(lane, v, S);
      __syncwarp();
      inv2_load(lane, v, tw, S);
      __syncwarp();
      inv2_store(lane, v, S);
      __syncwarp();
      inv3_load(lane, v, tw, S);
      // ---- synthesis window + overlap-add of the pair ---------------------------------------------------
      float* d0 = xo + (size_t)c * HOPN;
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const int i = lane + 32 * q;
        const float wa = WNn[i], wb = WNn[HOPN + i];
        d0[i] = fmaf(v[q].x, wa, carry[q]);
        if (has_b) {
          d0[HOPN + i] = fmaf(v[8 + q].x, wb, v[q].y * wa);
          carry[q] = v[8 + q].y * wb;
        } else {
          carry[q] = v[8 + q].x * wb;
        }
      }
    }
    float* dl = xo + (size_t)nrun * HOPN;
#pragma unroll
    for (int q = 0; q < 8; ++q) dl[lane + 32 * q] = carry[q];
    __syncwarp();
  }
}

int launch_gl_fast_n512(const b2d_plan* p, const float* mag_tf, float2* tprev, const float* xin, float* xout, int B, int T, int n,
                        int R, float mom, int use_prev, int store_prev, cudaStream_t st) {
  using namespace n512;
  GlN512Args a;
  a.mag_tf = mag_tf; a.tprev = tprev; a.xin = xin; a.xout = xout; a.B = B; a.T = T; a.n = n; a.R = R;
  a.tw512 = p->d_tw512; a.win = p->d_win; a.winn = p->d_winn; a.inv_env = p->d_inv_env;
  a.mom = mom; a.use_prev = use_prev; a.store_prev = store_prev;
  const size_t smem = sizeof(float) * (5 * HOPN) + (size_t)WARPS * WSMEM;
  const int runs = B * R;
  const int grid = runs < p->num_sms ? runs : p->num_sms;
  if (use_prev) {
    B2D_SMEM_OPT_IN(smem, gl_fast_n512_kernel<true>);
    B2D_CUDA(launch_pdl(gl_fast_n512_kernel<true>, dim3(grid), dim3(WARPS * 32), smem, st, a));
  } else {
    B2D_SMEM_OPT_IN(smem, gl_fast_n512_kernel<false>);
    B2D_CUDA(launch_pdl(gl_fast_n512_kernel<false>, dim3(grid), dim3(WARPS * 32), smem, st, a));
  }
  B2D_LAUNCH_CHECK("gl_fast_n512_kernel");
  return B2D_OK;
}

}  // namespace b2d
