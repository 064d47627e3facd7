// gl_warp.cu -- Griffin-Lim iteration for the transform lengths without a register-FFT path (n_fft 640, 1536, ...:
// any n_fft / 2 = 2^a 3^b 5^c with hop = n_fft / 2), organised like the n_fft = 1024 kernel instead of like the generic
// block-cooperative one: a WARP owns a frame and walks a run of frames, the mixed-radix Stockham passes run warp-
// synchronously in the warp's own shared-memory buffers (no block barrier anywhere after start-up), the overlap-add carry
// stays in shared memory between frames, tprev / mag rows arrive by TMA one frame ahead.  The iterate uses the same
// partial hop-block format and the same (n, R) partition as the other kernels, so init and stitch are shared.
// Used where at least 12 warps fit an SM (n_fft <= ~1100, e.g. the 16 kHz / n_fft 640 geometry of BASELINE config 3 in batch mode).
#include <stdlib.h>

#include "fft.cuh"
#include "kernels.cuh"

namespace b2d {

struct GlWarpArgs {
  const float* mag_tf;   // [B,T,Fp]
  float2* tprev;         // [B,T,M]  bin 0 = (Re X[0], Re X[M])
  const float* xin;      // partial hop-block format
  float* xout;
  int B, T, n, R;
  int M, Fp;             // hop == M, n_fft == 2 M
  FftDesc fd;
  const float2* tw;      // W_M^k
  const float2* rtw;     // W_N^k
  const float* win;      // [N]
  const float* winn;     // [N] win / N
  const float* inv_env;  // [hop]
  float mom;
  int use_prev, store_prev;
  int warps;             // warps per CTA
};

__device__ __forceinline__ uint32_t gw_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void gw_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(gw_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void gw_bulk(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(gw_u32(dst)), "l"(src),
               "r"(bytes), "r"(gw_u32(bar))
               : "memory");
}
__device__ __forceinline__ void gw_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "W_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra D_%=;\n"
      "bra W_%=;\n"
      "D_%=:\n"
      "}\n" ::"r"(gw_u32(bar)),
      "r"(parity)
      : "memory");
}

// one Stockham pass over a single row, executed by one warp
template <bool INV, int R>
__device__ __forceinline__ void warp_pass(const float2* __restrict__ src, float2* __restrict__ dst, int M, int Ns,
                                          const float2* __restrict__ tw, int lane) {
  for (int w = lane; w < M / R; w += 32) stockham_item<INV, R>(src, dst, w, M, M, Ns, tw);
}
template <bool INV>
__device__ __forceinline__ float2* warp_fft(float2* a, float2* b, const FftDesc& fd, const float2* __restrict__ tw, int lane) {
  int Ns = 1;
  for (int p = 0; p < fd.npass; ++p) {
    const int R = fd.radix[p];
    switch (R) {
      case 8: warp_pass<INV, 8>(a, b, fd.M, Ns, tw, lane); break;
      case 4: warp_pass<INV, 4>(a, b, fd.M, Ns, tw, lane); break;
      case 2: warp_pass<INV, 2>(a, b, fd.M, Ns, tw, lane); break;
      case 3: warp_pass<INV, 3>(a, b, fd.M, Ns, tw, lane); break;
      default: warp_pass<INV, 5>(a, b, fd.M, Ns, tw, lane); break;
    }
    __syncwarp();
    float2* t = a;
    a = b;
    b = t;
    Ns *= R;
  }
  return a;
}

__device__ __forceinline__ float gw_inv_abs(float s) {
  float r;
  asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(fmaxf(s, 1e-32f)));
  return r;
}

// envelope-normalised sample `is` of hop-block js (1 <= js <= T - 1) of clip b
__device__ __forceinline__ float gw_block_sample(const float* part, const float* __restrict__ inv_env, int b, int R, const FastDiv& dn,
                                                 int hop, int js, int is) {
  const int n = dn.d;
  const int r1 = dn.div(js - 1), r2 = dn.div(js);
  float v = part[((size_t)(b * R + r1) * (n + 1) + (js - r1 * n)) * hop + is];
  if (r2 != r1) v += part[((size_t)(b * R + r2) * (n + 1)) * hop + is];
  return v * inv_env[is];
}

template <bool USE_PREV>
__global__ void __launch_bounds__(512, 1) gl_warp_kernel(const GlWarpArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int M = a.M, hop = a.M, T = a.T, n = a.n, R = a.R;
  float2* tw_s = reinterpret_cast<float2*>(smem_raw);  // [M] shared by the CTA
  const size_t wbytes = (size_t)28 * M + (((size_t)a.Fp * 4 + 15) & ~(size_t)15) + 16;  // bufA | bufB | tprev row | carry | mag row | mbarrier
  unsigned char* wbase = smem_raw + (size_t)8 * M;
  for (int i = threadIdx.x; i < M; i += blockDim.x) tw_s[i] = a.tw[i];
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nruns = a.B * R;
  const int gw0 = warp * (int)gridDim.x + (int)blockIdx.x;  // runs dealt round-robin over the CTAs
  const int gstep = a.warps * (int)gridDim.x;
  if (gw0 >= nruns) return;
  unsigned char* wsm = wbase + (size_t)warp * wbytes;
  float2* bufA = reinterpret_cast<float2*>(wsm);
  float2* bufB = bufA + M;
  float2* tp_s = bufB + M;
  float* carry = reinterpret_cast<float*>(tp_s + M);
  float* mg_s = carry + hop;
  uint64_t* bar = reinterpret_cast<uint64_t*>(reinterpret_cast<unsigned char*>(mg_s) + (((size_t)a.Fp * 4 + 15) & ~(size_t)15));
  if (lane == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(gw_u32(bar)) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncwarp();
  const FastDiv d_n(n);
  const int half = M / 2;  // pairs k = 1 .. half (k == half is the self-paired bin when M is even)
  const uint32_t mag_bytes = (uint32_t)(((size_t)a.Fp * 4) & ~(size_t)15);  // Fp is a multiple of 4 floats
  const uint32_t row_bytes = mag_bytes + (USE_PREV ? (uint32_t)M * 8u : 0u);
  const size_t run_stride = (size_t)(n + 1) * hop;
  uint32_t uses = 0;

#pragma unroll 1
  for (int gw = gw0; gw < nruns; gw += gstep) {
    const int b = gw / R, r = gw - b * R;
    const int tb = r * n, te = min(T, tb + n);
    const float* xrun = a.xin + (size_t)(b * R + r) * run_stride;
    float* xo = a.xout + (size_t)(b * R + r) * run_stride;
    for (int i = lane; i < hop; i += 32) carry[i] = 0.f;
    if (lane == 0) {
      gw_expect_tx(bar, row_bytes);
      gw_bulk(mg_s, a.mag_tf + ((size_t)b * T + tb) * a.Fp, mag_bytes, bar);
      if (USE_PREV) gw_bulk(tp_s, a.tprev + ((size_t)b * T + tb) * M, (uint32_t)M * 8u, bar);
    }
    __syncwarp();
#pragma unroll 1
    for (int t = tb; t < te; ++t) {
      const int c = t - tb;
      // ---- stage the windowed frame as M complex values: hop-blocks t (first half) and t + 1 (second half) --------------
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int j = t + h, cs = c + h;
        float2* dst = bufA + h * (hop / 2);
        const float* wn = a.win + h * hop;
        if (j == 0 || j == T) {  // reflect-padded edge of the clip
          for (int m = lane; m < hop / 2; m += 32) {
            float v[2];
#pragma unroll
            for (int e = 0; e < 2; ++e) {
              const int i = 2 * m + e;
              int js, is;
              if (j == 0) { js = (i == 0) ? 2 : 1; is = (i == 0) ? 0 : hop - i; }
              else        { js = (i == hop - 1) ? T - 2 : T - 1; is = (i == hop - 1) ? hop - 1 : hop - 2 - i; }
              v[e] = gw_block_sample(a.xin, a.inv_env, b, R, d_n, hop, js, is) * wn[i];
            }
            dst[m] = make_float2(v[0], v[1]);
          }
        } else {
          const float2* p1 = reinterpret_cast<const float2*>(xrun + (size_t)cs * hop);
          const float2* p2 = nullptr;  // second partial when the block sits on a run boundary
          if (cs == 0) p2 = reinterpret_cast<const float2*>(xrun - run_stride + (size_t)n * hop);
          else if (j == te) p2 = reinterpret_cast<const float2*>(xrun + run_stride);
          const float2* e2 = reinterpret_cast<const float2*>(a.inv_env);
          const float2* w2 = reinterpret_cast<const float2*>(wn);
          for (int m = lane; m < hop / 2; m += 32) {
            float2 xv = p1[m];
            if (p2) { const float2 x2 = p2[m]; xv.x += x2.x; xv.y += x2.y; }
            const float2 ev = __ldg(e2 + m), wv = __ldg(w2 + m);
            dst[m] = make_float2(xv.x * ev.x * wv.x, xv.y * ev.y * wv.y);
          }
        }
      }
      __syncwarp();
      float2* res = warp_fft<false>(bufA, bufB, a.fd, tw_s, lane);
      // ---- split -> momentum / projection -> (store rebuilt) -> x mag -> merge, in place -----------------------------------
      float2* tp = a.tprev + ((size_t)b * T + t) * M;
      gw_wait(bar, uses & 1);
      ++uses;
      for (int k = lane; k <= half; k += 32) {
        const int mk = (M - k) & -(k != 0);  // k == 0 pairs with itself
        const float2 zk = res[k], zmk = res[mk];
        float2 xk, xmk;
        rfft_split(zk, zmk, __ldg(a.rtw + k), xk, xmk);
        const float m1 = mg_s[k], m2 = mg_s[M - k];
        float2 yk, ymk;
        if (k == 0) {
          float a0 = xk.x, aM = xmk.x;
          if (USE_PREV) { const float2 pv = tp_s[0]; a0 -= a.mom * pv.x; aM -= a.mom * pv.y; }
          yk = make_float2(m1 * (a0 * gw_inv_abs(a0 * a0)), 0.f);
          ymk = make_float2(m2 * (aM * gw_inv_abs(aM * aM)), 0.f);
          if (a.store_prev) __stcs(tp, make_float2(xk.x, xmk.x));
        } else {
          float2 ak = xk, amk = xmk;
          if (USE_PREV) {
            const float2 pk = tp_s[k], pmk = tp_s[M - k];
            ak = make_float2(xk.x - a.mom * pk.x, xk.y - a.mom * pk.y);
            amk = make_float2(xmk.x - a.mom * pmk.x, xmk.y - a.mom * pmk.y);
          }
          const float i1 = gw_inv_abs(ak.x * ak.x + ak.y * ak.y), i2 = gw_inv_abs(amk.x * amk.x + amk.y * amk.y);
          yk = make_float2(m1 * (ak.x * i1), m1 * (ak.y * i1));  // same association as the block-cooperative kernel: bit-identical results
          ymk = make_float2(m2 * (amk.x * i2), m2 * (amk.y * i2));
          if (a.store_prev) {
            __stcs(tp + k, xk);
            if (2 * k != M) __stcs(tp + (M - k), xmk);
          }
        }
        float2 z1, z2;
        irfft_merge(yk, ymk, __ldg(a.rtw + k), z1, z2);
        res[k] = z1;
        if (k != 0 && 2 * k != M) res[M - k] = z2;
      }
      __syncwarp();  // every lane is done with the staged rows
      if (lane == 0 && t + 1 < te) {
        gw_expect_tx(bar, row_bytes);
        gw_bulk(mg_s, a.mag_tf + ((size_t)b * T + t + 1) * a.Fp, mag_bytes, bar);
        if (USE_PREV) gw_bulk(tp_s, a.tprev + ((size_t)b * T + t + 1) * M, (uint32_t)M * 8u, bar);
      }
      float2* other = (res == bufA) ? bufB : bufA;
      const float* y = reinterpret_cast<const float*>(warp_fft<true>(res, other, a.fd, tw_s, lane));
      // ---- synthesis window + overlap-add: block c = carry + first half ; carry = second half ------------------------------
      float* dsto = xo + (size_t)c * hop;
      for (int i = lane; i < hop; i += 32) {
        dsto[i] = fmaf(y[i], __ldg(a.winn + i), carry[i]);
        carry[i] = y[hop + i] * __ldg(a.winn + hop + i);
      }
      __syncwarp();
    }
    float* dl = xo + (size_t)(te - tb) * hop;
    for (int i = lane; i < hop; i += 32) dl[i] = carry[i];
    __syncwarp();
  }
}

static size_t gl_warp_bytes_per_warp(const b2d_plan* p) {
  return (size_t)28 * p->M + (((size_t)p->Fp * 4 + 15) & ~(size_t)15) + 16;
}
// warps per CTA (one CTA per SM) for this plan: as many as the shared memory holds, 16 at most (512 threads)
int gl_warp_warps(const b2d_plan* p) {
  const size_t avail = (size_t)227 * 1024 - (size_t)8 * p->M - 1024;
  size_t w = avail / gl_warp_bytes_per_warp(p);
  if (w > 16) w = 16;
  return (int)w;
}
bool gl_warp_supported(const b2d_plan* p) {
  // measured at B = 256: n_fft 640 (16 warps / SM) 392 -> 286 us per iteration; n_fft 1536 (9 warps / SM) 1128 -> 1282 us, so
  // the long transforms stay on the block-cooperative kernel
  return p->hop == p->M && (p->M % 4) == 0 && (p->Fp % 4) == 0 && gl_warp_warps(p) >= 12;
}

int launch_gl_warp(const b2d_plan* p, const float* mag_tf, float2* tprev, const float* xin, float* xout, int B, int T, int n, int R,
                   float mom, int use_prev, int store_prev, cudaStream_t st) {
  GlWarpArgs a;
  a.mag_tf = mag_tf; a.tprev = tprev; a.xin = xin; a.xout = xout; a.B = B; a.T = T; a.n = n; a.R = R;
  a.M = p->M; a.Fp = p->Fp; a.fd = p->fft; a.tw = p->d_tw; a.rtw = p->d_rtw; a.win = p->d_win; a.winn = p->d_winn;
  a.inv_env = p->d_inv_env; a.mom = mom; a.use_prev = use_prev; a.store_prev = store_prev;
  a.warps = gl_warp_warps(p);
  const size_t smem = (size_t)8 * p->M + (size_t)a.warps * gl_warp_bytes_per_warp(p);
  const int runs = B * R;
  const int grid = runs < p->num_sms ? runs : p->num_sms;
  if (use_prev) {
    B2D_SMEM_OPT_IN(smem, gl_warp_kernel<true>);
    gl_warp_kernel<true><<<grid, a.warps * 32, smem, st>>>(a);
  } else {
    B2D_SMEM_OPT_IN(smem, gl_warp_kernel<false>);
    gl_warp_kernel<false><<<grid, a.warps * 32, smem, st>>>(a);
  }
  B2D_LAUNCH_CHECK("gl_warp_kernel");
  return B2D_OK;
}

}  // namespace b2d
