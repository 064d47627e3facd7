// model.cu -- K3: GRUUNet2 forward (gruunet2.py:54-306), restructured for the GPU:
//
//   encoder  (input_gate DownBlocks, gruunet2.py:127-144)  : depends only on x_t  -> all B*T frames in parallel
//   recurrence (reset_gate conv + gates, gruunet2.py:146-155, 231-240) : the only sequential part; one CTA per
//              clip, the [51,17,3] recurrent weights live in registers for the whole sequence, gate_x hoisted
//   decoder  (output_gate UpBlocks, gruunet2.py:184-199)   : depends only on h_t and the skips -> all frames in parallel
//
// The GaussianSmearing position channels (gruunet2.py:54-68) are input independent, so their
// contribution is folded into a per-position bias when the model is packed.
// This file holds the fp32 CUDA-core convolutions (B2D_CONV_FP32, the exact engine) and the recurrence; unet_mma.cu holds the
// default warp-level tensor-core encoder / decoder, unet_tc.cu the fused tcgen05 encoder.
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <vector>

#include "kernels.cuh"
#include "tma.cuh"
#include "model_layout.cuh"

namespace b2d {

__device__ __forceinline__ void prefetch_l2_line(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
// warm L2 with [p, p + bytes): one 128-byte line per lane per round
__device__ __forceinline__ void prefetch_l2_range(const float* p, int bytes, int lane) {
  for (int o = lane * 128; o < bytes; o += 32 * 128) prefetch_l2_line(reinterpret_cast<const char*>(p) + o);
}

// ------------------------------------------------------------------------------------------------
// encoder: one warp per frame.  Work item = (output position j, group of 4 output channels).
// ------------------------------------------------------------------------------------------------
template <int CIN, int COP, int LIN>
__device__ __forceinline__ void enc_layer(const float* __restrict__ in, const float* __restrict__ W,
                                          const float* __restrict__ PB, float* __restrict__ out, int cout, int lane) {
  constexpr int LOUT = LIN / 2;
  constexpr int NG = COP / 4;
  for (int item = lane; item < LOUT * NG; item += 32) {
    const int j = item % LOUT, cg = item / LOUT;
    float4 acc = *reinterpret_cast<const float4*>(PB + j * COP + cg * 4);
#pragma unroll 6
    for (int ci = 0; ci < CIN; ++ci) {
      const float* row = in + ci * LIN;
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        const int p = 2 * j - 1 + k;
        const float v = (p >= 0 && p < LIN) ? row[p] : 0.f;
        const float4 w = *reinterpret_cast<const float4*>(W + (ci * 3 + k) * COP + cg * 4);
        acc.x = fmaf(v, w.x, acc.x);
        acc.y = fmaf(v, w.y, acc.y);
        acc.z = fmaf(v, w.z, acc.z);
        acc.w = fmaf(v, w.w, acc.w);
      }
    }
    const float r[4] = {fmaxf(acc.x, 0.f), fmaxf(acc.y, 0.f), fmaxf(acc.z, 0.f), fmaxf(acc.w, 0.f)};
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int co = cg * 4 + q;
      if (co < cout) out[co * LOUT + j] = r[q];
    }
  }
}

constexpr int ENC_WARPS = 8;
constexpr int ENC_ACT = NMEL + D0 + D1 + D2 + GX;  // 1220 floats per warp

__global__ void __launch_bounds__(ENC_WARPS * 32) encoder_kernel(const float* __restrict__ blob, const float* __restrict__ x,
                                                                 size_t nframes, float* __restrict__ d0, float* __restrict__ d1,
                                                                 float* __restrict__ d2, float* __restrict__ gx) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const Packed L = packed_layout();
  float* wts = reinterpret_cast<float*>(smem_raw);  // encoder part of the blob: [0, rec_w)
  const int nw = L.rec_w;
  float* act = wts + ((nw + 3) & ~3);
  for (int i = threadIdx.x; i < nw; i += blockDim.x) wts[i] = blob[i];
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* a_in = act + warp * ENC_ACT;
  float* a0 = a_in + NMEL;
  float* a1 = a0 + D0;
  float* a2 = a1 + D1;
  float* a3 = a2 + D2;
  for (size_t f = (size_t)blockIdx.x * ENC_WARPS + warp; f < nframes; f += (size_t)gridDim.x * ENC_WARPS) {
    const float* xf = x + f * NMEL;
    if (lane < 2 && f + (size_t)gridDim.x * ENC_WARPS < nframes) prefetch_l2_line(xf + (size_t)gridDim.x * ENC_WARPS * NMEL + lane * 32);
    a_in[lane] = xf[lane];
    a_in[lane + 32] = xf[lane + 32];
    __syncwarp();
    enc_layer<1, HP, 64>(a_in, wts + L.enc_w[0], wts + L.enc_pb[0], a0, H, lane);
    __syncwarp();
    enc_layer<H, HP, 32>(a0, wts + L.enc_w[1], wts + L.enc_pb[1], a1, H, lane);
    __syncwarp();
    enc_layer<H, HP, 16>(a1, wts + L.enc_w[2], wts + L.enc_pb[2], a2, H, lane);
    __syncwarp();
    enc_layer<H, H3P, 8>(a2, wts + L.enc_w[3], wts + L.enc_pb[3], a3, H3, lane);
    __syncwarp();
    for (int i = lane; i < D0; i += 32) d0[f * D0 + i] = a0[i];
    for (int i = lane; i < D1; i += 32) d1[f * D1 + i] = a1[i];
    for (int i = lane; i < D2; i += 32) d2[f * D2 + i] = a2[i];
    for (int i = lane; i < GX; i += 32) gx[f * GX + i] = a3[i];
    __syncwarp();
  }
}

// (A register-tiled variant of this encoder -- a lane owning one output position of one frame, same scheme as the
//  decoder below -- measured 25 us slower: both are bound by the shared-memory weight loads.  The default encoder is the
//  tensor-core one in unet_mma.cu; this kernel is conv_mode "fp32".)
// ------------------------------------------------------------------------------------------------
// recurrence: one CTA (96 threads, 68 active) per clip; thread (c, j) owns hidden channel c at
// compressed bin j and keeps its 3 x 17 x 3 recurrent weights in registers for all T steps.
// ------------------------------------------------------------------------------------------------
// Gate non-linearities on the exp2 unit (ex2.approx has 2 ulp of error, the reciprocal 1 ulp: ~3e-7 relative, far inside
// the 1e-5 model parity budget); tanh(v) = 1 - 2 / (1 + e^{2v}) saturates cleanly to +-1 when the exponential overflows / underflows.
// EXACT (B2D_CONV_EXACT_GATES): expf and IEEE division -- the attribution switch of test_exact_math_attribution.
template <bool EXACT>
__device__ __forceinline__ float sigmoidf_(float v) { return EXACT ? 1.0f / (1.0f + expf(-v)) : __fdividef(1.0f, 1.0f + __expf(-v)); }
template <bool EXACT>
__device__ __forceinline__ float tanhf_(float v) { return EXACT ? tanhf(v) : 1.0f - __fdividef(2.0f, 1.0f + __expf(2.0f * v)); }

// The hoisted input gates gx [B, T, 204] stream in by TMA, REC_CHUNK steps (6.5 KB, contiguous) at a time into a two-slot
// shared-memory ring, two chunks ahead of the step that consumes them.  (Round 1 prefetched one step ahead with plain loads:
// a step is shorter than an L2 round trip, so every step waited for its gates -- 0.58 us per step.)
constexpr int REC_CHUNK = 8;
// (Measured, round 2 -- ncu on this kernel at B = 256: 231 warp instructions per warp and step (204 of them the 51 LDS + 153 FMA
//  of the three gates), issue slots 44 % busy with two CTAs = 6 warps per SM: the step is bound by the instructions the two
//  resident clips issue, not by one warp's latency.  Two restructurings were built, bit-identical, and dropped: (a) 4 warps =
//  bins, lane = channel, the three taps as one float4 broadcast LDS.128 + packed FMAs (119 instead of 204 instructions per
//  thread-step, but 254 registers and 8-way bank conflicts on the tap stores): 66 us, no gain; (b) a warp per (gate, bin), 12
//  warps, two barriers per step: 139 us -- 60 % more instructions per clip-step at 17 of 32 lanes busy; (c) this layout with every
//  hidden value stored twice so that one LDS.64 feeds a packed FMA for the reset / update sums (159 instead of 204 instructions):
//  75 us and 255 registers with spills -- the packed FMA does not issue at the scalar rate here.)

template <bool EXACT>
__global__ void __launch_bounds__(96) recurrence_kernel(const float* __restrict__ blob, const float* __restrict__ gx,
                                                        float* __restrict__ hx, float* __restrict__ hseq, int T) {
  const Packed L = packed_layout();
  __shared__ float hp[2][H][BINS + 2];  // double-buffered hidden state, zero padded at both ends: one barrier per step
  __shared__ __align__(16) float gxs[2][REC_CHUNK * GX];
  __shared__ uint64_t gbar[2];
  const int b = blockIdx.x;
  const int tid = threadIdx.x;
  const bool active = tid < H * BINS;
  const int c = active ? tid / BINS : 0, j = active ? tid % BINS : 0;
  const float* gxb = gx + (size_t)b * T * GX;
  const int nchunks = (T + REC_CHUNK - 1) / REC_CHUNK;
  auto fetch = [&](int chunk) {  // thread 0 only
    const int steps = min(REC_CHUNK, T - chunk * REC_CHUNK);
    const uint32_t bytes = (uint32_t)steps * GX * sizeof(float);  // 816 B per step: a multiple of 16
    tma::expect_bytes(&gbar[chunk & 1], bytes);
    tma::load(gxs[chunk & 1], gxb + (size_t)chunk * REC_CHUNK * GX, bytes, &gbar[chunk & 1]);
  };
  if (tid == 0) {
    tma::barrier_init(&gbar[0], 1);
    tma::barrier_init(&gbar[1], 1);
    tma::fence_barrier_init();
  }
  float w[3][H][3];
  float pb[3];
  {
    const float* W = blob + L.rec_w;
#pragma unroll
    for (int g = 0; g < 3; ++g) {
#pragma unroll
      for (int ci = 0; ci < H; ++ci)
#pragma unroll
        for (int k = 0; k < 3; ++k) w[g][ci][k] = W[((ci * 3 + k) * 3 + g) * HP + c];
      pb[g] = blob[L.rec_pb + (g * H + c) * BINS + j];
    }
  }
  for (int i = tid; i < 2 * H * (BINS + 2); i += blockDim.x) (&hp[0][0][0])[i] = 0.f;
  pdl_wait();  // weights are in registers; gx / hx come from the kernels before (common.cuh: programmatic dependent launch)
  pdl_trigger();
  if (tid == 0) {
    if (nchunks > 0) fetch(0);
    if (nchunks > 1) fetch(1);
  }
  __syncthreads();  // also publishes the mbarrier initialisation to every thread
  float h = 0.f;
  if (active) {
    h = hx[(size_t)b * HS + c * BINS + j];
    hp[0][c][j + 1] = h;
  }
  __syncthreads();
  float* hsb = hseq + (size_t)b * T * HS;
  for (int t = 0; t < T; ++t) {
    const int chunk = t / REC_CHUNK, s = t - chunk * REC_CHUNK;
    if (s == 0) tma::wait(&gbar[chunk & 1], (chunk >> 1) & 1);
    const float* g1 = gxs[chunk & 1] + s * GX;
    const float xr = g1[c * BINS + j], xz = g1[(H + c) * BINS + j], xn = g1[(2 * H + c) * BINS + j];
    // three partial sums per gate (one per tap): dependency chains of 17 instead of 51 FMAs
    const float (*hc)[BINS + 2] = hp[t & 1];
    float ar[3] = {pb[0], 0.f, 0.f}, az[3] = {pb[1], 0.f, 0.f}, an[3] = {pb[2], 0.f, 0.f};
#pragma unroll
    for (int ci = 0; ci < H; ++ci) {
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        const float v = hc[ci][j + k];
        ar[k] = fmaf(w[0][ci][k], v, ar[k]);
        az[k] = fmaf(w[1][ci][k], v, az[k]);
        an[k] = fmaf(w[2][ci][k], v, an[k]);
      }
    }
    const float sr = fmaxf(ar[0] + ar[1] + ar[2], 0.f);  // both pre-activations went through ReLU (gruunet2.py:71-79, Q8)
    const float sz = fmaxf(az[0] + az[1] + az[2], 0.f);
    const float sn = fmaxf(an[0] + an[1] + an[2], 0.f);
    const float z = sigmoidf_<EXACT>(xz + sz);
    const float r = sigmoidf_<EXACT>(xr + sr);
    const float nw = tanhf_<EXACT>(xn + r * sn);
    const float hn = nw + z * (h - nw);
    if (active) {
      hp[(t + 1) & 1][c][j + 1] = hn;  // the other buffer: nobody reads it during this step
      hsb[(size_t)t * HS + c * BINS + j] = hn;
    }
    h = hn;
    __syncthreads();
    // every thread has read its gates of this chunk: the slot can take the chunk after next
    if (tid == 0 && s == REC_CHUNK - 1 && chunk + 2 < nchunks) fetch(chunk + 2);
  }
  if (active) hx[(size_t)b * HS + c * BINS + j] = h;
}

// ------------------------------------------------------------------------------------------------
// decoder (conv_mode "fp32"; the default is the tensor-core decoder in unet_mma.cu).  ConvTranspose1d(k3,s2,p1,op1):
//   out[co,2j]   = pb + sum_ci x[ci,j] W[ci,1,co]
//   out[co,2j+1] = pb + sum_ci (x[ci,j] W[ci,2,co] + x[ci,j+1] W[ci,0,co])
// (The first version -- one warp per frame, work item = (output position, 4 output channels) -- took 655 us.)
// ------------------------------------------------------------------------------------------------
// (Measured dead end, kept as a note: passing the 27.5 KB decoder weights as a __grid_constant__ kernel parameter so that
//  FMAs take warp-uniform constant operands was slower -- register-indexed LDC.64 when partially unrolled (+0.4 ms), and
//  110 KB of straight-line code when fully unrolled (+0.05 ms).  These kernels are bound by the shared-memory -> register
//  path: a broadcast float4 weight load still moves 512 B to the register file.)
// decoder, register-tiled: a warp works on FW = 2 frames at once; a lane owns one INPUT position j of one
// frame and computes all output channels of both outputs 2j and 2j+1 (2 x 17 accumulators in registers).
// Per input channel: 2 activation loads, 15 broadcast float4 weight loads, 60 FMAs.
// Per-frame staging (floats): region A [544] = s2 [34][16] (also s0 [17][4] and the final [64]),
//                             region B [1088] = s3 [34][32] (also s1 [34][8]); skips sit in the upper channel halves.
// ------------------------------------------------------------------------------------------------
constexpr int DEC2_WARPS = 6;
constexpr int DEC2_FW = 2;
constexpr int DEC2_FR = 544 + 1088 + 16;  // +16: the two frames of a warp land on different banks

// CSPLIT lanes share one (frame, j): lane `cs` computes the output-channel groups [cs*NGL, (cs+1)*NGL) -- keeps all 32
// lanes busy in the small layers (up0: 2 frames x 4 positions x 4 splits, up1: 2 x 8 x 2).
template <int CIN, int COUT, int COP, int LIN, int CSPLIT, bool RELU>
__device__ __forceinline__ void dec2_layer(const float* __restrict__ in, const float* __restrict__ W, const float* __restrict__ PB,
                                           float* __restrict__ out, int j, int cs) {
  constexpr int NG = COP / 4;
  constexpr int NGL = (NG + CSPLIT - 1) / CSPLIT;
  const int g0 = cs * NGL;
  float4 ae[NGL], ao[NGL];
#pragma unroll
  for (int gl = 0; gl < NGL; ++gl) {
    const bool on = (CSPLIT == 1) || (g0 + gl < NG);
    ae[gl] = on ? *reinterpret_cast<const float4*>(PB + (2 * j) * COP + 4 * (g0 + gl)) : make_float4(0.f, 0.f, 0.f, 0.f);
    ao[gl] = on ? *reinterpret_cast<const float4*>(PB + (2 * j + 1) * COP + 4 * (g0 + gl)) : make_float4(0.f, 0.f, 0.f, 0.f);
  }
  const bool has_next = (j + 1 < LIN);
#pragma unroll 2
  for (int ci = 0; ci < CIN; ++ci) {
    const float a0 = in[ci * LIN + j];
    const float a1 = has_next ? in[ci * LIN + j + 1] : 0.f;
    const float* w = W + ci * 3 * COP + 4 * g0;
#pragma unroll
    for (int gl = 0; gl < NGL; ++gl) {
      if ((CSPLIT == 1) || (g0 + gl < NG)) {
        const float4 w0 = *reinterpret_cast<const float4*>(w + 4 * gl);
        const float4 w1 = *reinterpret_cast<const float4*>(w + COP + 4 * gl);
        const float4 w2 = *reinterpret_cast<const float4*>(w + 2 * COP + 4 * gl);
        ae[gl].x = fmaf(a0, w1.x, ae[gl].x); ae[gl].y = fmaf(a0, w1.y, ae[gl].y);
        ae[gl].z = fmaf(a0, w1.z, ae[gl].z); ae[gl].w = fmaf(a0, w1.w, ae[gl].w);
        ao[gl].x = fmaf(a0, w2.x, fmaf(a1, w0.x, ao[gl].x)); ao[gl].y = fmaf(a0, w2.y, fmaf(a1, w0.y, ao[gl].y));
        ao[gl].z = fmaf(a0, w2.z, fmaf(a1, w0.z, ao[gl].z)); ao[gl].w = fmaf(a0, w2.w, fmaf(a1, w0.w, ao[gl].w));
      }
    }
  }
#pragma unroll
  for (int gl = 0; gl < NGL; ++gl) {
    const float e[4] = {ae[gl].x, ae[gl].y, ae[gl].z, ae[gl].w};
    const float o[4] = {ao[gl].x, ao[gl].y, ao[gl].z, ao[gl].w};
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int co = 4 * (g0 + gl) + q;
      if (co < COUT) {
        float2 r = make_float2(e[q], o[q]);
        if (RELU) r = make_float2(fmaxf(r.x, 0.f), fmaxf(r.y, 0.f));
        *reinterpret_cast<float2*>(out + co * (2 * LIN) + 2 * j) = r;
      }
    }
  }
}

__global__ void __launch_bounds__(DEC2_WARPS * 32, 2) decoder2_kernel(const float* __restrict__ blob, const float* __restrict__ hseq,
                                                                      const float* __restrict__ d0, const float* __restrict__ d1,
                                                                      const float* __restrict__ d2, const float* __restrict__ x,
                                                                      size_t nframes, float* __restrict__ pred, float* __restrict__ mel,
                                                                      int fused_mode, float out_scale) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const Packed L = packed_layout();
  float* wts = reinterpret_cast<float*>(smem_raw);
  const int w0 = L.dec_w[0];
  const int nw = L.total - w0;
  float* act = wts + ((nw + 3) & ~3);
  for (int i = threadIdx.x; i < nw; i += blockDim.x) wts[i] = blob[w0 + i];
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* base = act + warp * (DEC2_FW * DEC2_FR);
  const size_t stride = (size_t)gridDim.x * DEC2_WARPS * DEC2_FW;
  for (size_t f0 = ((size_t)blockIdx.x * DEC2_WARPS + warp) * DEC2_FW; f0 < nframes; f0 += stride) {
    const int nf = (int)min((size_t)DEC2_FW, nframes - f0);
    if (f0 + stride + DEC2_FW <= nframes) {  // next iteration's inputs -> L2 while this one computes
      const size_t fn = f0 + stride;
      prefetch_l2_range(d0 + fn * D0, DEC2_FW * D0 * 4, lane);
      prefetch_l2_range(d1 + fn * D1, DEC2_FW * D1 * 4, lane);
      prefetch_l2_range(d2 + fn * D2, DEC2_FW * D2 * 4, lane);
      prefetch_l2_range(hseq + fn * HS, DEC2_FW * HS * 4, lane);
      prefetch_l2_range(x + fn * NMEL, DEC2_FW * NMEL * 4, lane);
    }
    // stage hidden state and skips of both frames (frames are contiguous in global memory)
    for (int f = 0; f < nf; ++f) {
      float* A = base + f * DEC2_FR;  // region A
      float* Bq = A + 544;            // region B
      const size_t fr = f0 + f;
      for (int i = lane; i < HS; i += 32) A[i] = hseq[fr * HS + i];                 // s0 [17][4]
      for (int i = lane; i < D2; i += 32) Bq[H * 8 + i] = d2[fr * D2 + i];          // s1 upper half
      for (int i = lane; i < D1; i += 32) A[H * 16 + i] = d1[fr * D1 + i];          // s2 upper half
      for (int i = lane; i < D0; i += 32) Bq[H * 32 + i] = d0[fr * D0 + i];         // s3 upper half
    }
    __syncwarp();
    {  // up0: [17][4] -> [17][8]   lane = f*16 + j*4 + channel split (4 ways)
      const int f = lane >> 4, j = (lane >> 2) & 3, cs = lane & 3;
      if (f < nf) {
        float* A = base + f * DEC2_FR;
        dec2_layer<H, H, HP, 4, 4, true>(A, wts + (L.dec_w[0] - w0), wts + (L.dec_pb[0] - w0), A + 544, j, cs);
      }
    }
    __syncwarp();
    {  // up1: [34][8] -> [17][16]  lane = f*16 + j*2 + channel split (2 ways)
      const int f = lane >> 4, j = (lane >> 1) & 7, cs = lane & 1;
      if (f < nf) {
        float* A = base + f * DEC2_FR;
        dec2_layer<2 * H, H, HP, 8, 2, true>(A + 544, wts + (L.dec_w[1] - w0), wts + (L.dec_pb[1] - w0), A, j, cs);
      }
    }
    __syncwarp();
    {  // up2: [34][16] -> [17][32]
      const int f = lane >> 4, j = lane & 15;
      if (f < nf) {
        float* A = base + f * DEC2_FR;
        dec2_layer<2 * H, H, HP, 16, 1, true>(A, wts + (L.dec_w[2] - w0), wts + (L.dec_pb[2] - w0), A + 544, j, 0);
      }
    }
    __syncwarp();
    for (int f = 0; f < nf; ++f) {  // up3: [34][32] -> [1][64]
      float* A = base + f * DEC2_FR;
      dec2_layer<2 * H, 1, 4, 32, 1, false>(A + 544, wts + (L.dec_w[3] - w0), wts + (L.dec_pb[3] - w0), A, lane, 0);
    }
    __syncwarp();
    for (int f = 0; f < nf; ++f) {
      const float* A = base + f * DEC2_FR;
      const size_t fr = f0 + f;
      for (int i = lane; i < NMEL; i += 32) {
        const float p = A[i];
        pred[fr * NMEL + i] = p;
        if (fused_mode) {
          const float xv = x[fr * NMEL + i];
          float v;
          if (fused_mode == 1) {
            float r = xv - p;
            r = r > 0.f ? r : 0.2f * r;
            v = fmaxf(expm1f(r), 0.f);
          } else {
            v = expf(xv - fmaxf(p, 0.f) * out_scale) - 1.0f;
          }
          mel[fr * NMEL + i] = v;
        }
      }
    }
    __syncwarp();
  }
}

__global__ void scale_kernel(float* p, size_t n, float s) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] *= s;
}
int scale_inplace(float* p, size_t n, float s, cudaStream_t st) {
  if (n == 0) return B2D_OK;
  scale_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(p, n, s);
  B2D_LAUNCH_CHECK("scale_kernel");
  return B2D_OK;
}

// ------------------------------------------------------------------------------------------------
// host: packing
// ------------------------------------------------------------------------------------------------
static void smear_table(const float* off, int G, int nbins, std::vector<double>& s) {
  // gruunet2.py:54-68: exp(coeff * (p - o)^2), p = linspace(0,1,nbins), coeff = -0.5/(o[1]-o[0])^2 (f32 gap)
  const double gap = (double)(float)(off[1] - off[0]);
  const double coeff = -0.5 / (gap * gap);
  s.assign((size_t)G * nbins, 0.0);
  for (int p = 0; p < nbins; ++p) {
    const double pos = (nbins > 1) ? (double)(float)((double)p / (double)(nbins - 1)) : 0.0;
    for (int g = 0; g < G; ++g) {
      const double d = (double)(float)(pos - (double)off[g]);
      s[(size_t)g * nbins + p] = exp((double)(float)coeff * (double)(float)(d * d));
    }
  }
}

int model_pack_mma(b2d_model* m);                                    // unet_mma.cu
int model_pack_utc(b2d_model* m, const float* const* hp);             // unet_tc.cu
int model_encode_utc(const b2d_model* m, const float* x, size_t nframes, float* d0, float* d1, float* d2, float* gx, int num_sms,
                     cudaStream_t st);
int model_encode_mma(const b2d_model* m, const float* x, size_t nframes, float* d0, float* d1, float* d2, float* gx, int terms,
                     int num_sms, cudaStream_t st);
int model_decode_mma(const b2d_model* m, const float* hseq, const float* d0, const float* d1, const float* d2, const float* x,
                     size_t nframes, float* pred, float* mel, int fused_mode, float out_scale, int terms, int num_sms, cudaStream_t st);

int model_pack(b2d_model* m, const float* const* hp, const float* const* offs) {
  const Packed L = packed_layout();
  const int G = m->cfg.num_gaussians;
  std::vector<float> blob((size_t)L.total, 0.f);
  std::vector<double> S;
  // encoder: params 2l (weight [co][ci+G][3]) , 2l+1 (bias [co])
  const int cin[LEVELS] = {1, H, H, H};
  const int cout[LEVELS] = {H, H, H, H3};
  const int cop[LEVELS] = {HP, HP, HP, H3P};
  const int lin[LEVELS] = {64, 32, 16, 8};
  for (int l = 0; l < LEVELS; ++l) {
    const float* W = hp[2 * l];
    const float* bias = hp[2 * l + 1];
    const int CT = cin[l] + G, Lin = lin[l], Lout = Lin / 2;
    smear_table(offs[0], G, Lin, S);
    for (int co = 0; co < cout[l]; ++co) {
      for (int ci = 0; ci < cin[l]; ++ci)
        for (int k = 0; k < 3; ++k) blob[L.enc_w[l] + (ci * 3 + k) * cop[l] + co] = W[(co * CT + ci) * 3 + k];
      for (int j = 0; j < Lout; ++j) {
        double acc = bias[co];
        for (int g = 0; g < G; ++g)
          for (int k = 0; k < 3; ++k) {
            const int p = 2 * j - 1 + k;
            if (p >= 0 && p < Lin) acc += (double)W[(co * CT + cin[l] + g) * 3 + k] * S[(size_t)g * Lin + p];
          }
        blob[L.enc_pb[l] + j * cop[l] + co] = (float)acc;
      }
    }
  }
  // recurrent conv: params 2*LEVELS, 2*LEVELS+1 ; weight [3H][H+G][3], stride 1, pad 1, length BINS
  {
    const float* W = hp[2 * LEVELS];
    const float* bias = hp[2 * LEVELS + 1];
    const int CT = H + G;
    smear_table(offs[1], G, BINS, S);
    for (int g3 = 0; g3 < 3; ++g3)
      for (int c = 0; c < H; ++c) {
        const int co = g3 * H + c;
        for (int ci = 0; ci < H; ++ci)
          for (int k = 0; k < 3; ++k) blob[L.rec_w + ((ci * 3 + k) * 3 + g3) * HP + c] = W[(co * CT + ci) * 3 + k];
        for (int j = 0; j < BINS; ++j) {
          double acc = bias[co];
          for (int g = 0; g < G; ++g)
            for (int k = 0; k < 3; ++k) {
              const int p = j - 1 + k;
              if (p >= 0 && p < BINS) acc += (double)W[(co * CT + H + g) * 3 + k] * S[(size_t)g * BINS + p];
            }
          blob[L.rec_pb + co * BINS + j] = (float)acc;
        }
      }
  }
  // decoder: params 2*LEVELS+2+2i ; ConvTranspose weight [ci+G][co][3]
  const int dcin[LEVELS] = {H, 2 * H, 2 * H, 2 * H};
  const int dcout[LEVELS] = {H, H, H, 1};
  const int dcop[LEVELS] = {HP, HP, HP, 4};
  const int dlin[LEVELS] = {4, 8, 16, 32};
  for (int i = 0; i < LEVELS; ++i) {
    const float* W = hp[2 * LEVELS + 2 + 2 * i];
    const float* bias = hp[2 * LEVELS + 3 + 2 * i];
    const int Lin = dlin[i], CO = dcout[i];
    smear_table(offs[2], G, Lin, S);
    for (int co = 0; co < CO; ++co) {
      for (int ci = 0; ci < dcin[i]; ++ci)
        for (int k = 0; k < 3; ++k) blob[L.dec_w[i] + (ci * 3 + k) * dcop[i] + co] = W[(ci * CO + co) * 3 + k];
      for (int o = 0; o < 2 * Lin; ++o) {
        const int j = o >> 1;
        double acc = bias[co];
        for (int g = 0; g < G; ++g) {
          const float* Wg = W + ((dcin[i] + g) * CO + co) * 3;
          if (o & 1) {
            acc += (double)Wg[2] * S[(size_t)g * Lin + j];
            if (j + 1 < Lin) acc += (double)Wg[0] * S[(size_t)g * Lin + j + 1];
          } else {
            acc += (double)Wg[1] * S[(size_t)g * Lin + j];
          }
        }
        blob[L.dec_pb[i] + o * dcop[i] + co] = (float)acc;
      }
    }
  }
  m->blob_floats = blob.size();
  m->h_blob = static_cast<float*>(malloc(blob.size() * sizeof(float)));  // host copy: constant-bank kernels take weights by value
  if (m->h_blob) memcpy(m->h_blob, blob.data(), blob.size() * sizeof(float));
  B2D_CUDA(cudaMalloc(&m->d_blob, blob.size() * sizeof(float)));
  B2D_CUDA(cudaMemcpy(m->d_blob, blob.data(), blob.size() * sizeof(float), cudaMemcpyHostToDevice));
  int rc = model_pack_mma(m);  // weight fragments for the warp-level MMA decoder (unet_mma.cu)
  if (rc != B2D_OK) return rc;
  return model_pack_utc(m, hp);  // weight images of the fused tcgen05 encoder (unet_tc.cu)
}

bool model_config_supported(const b2d_model_config* c) {
  return c->hidden == H && c->levels == LEVELS && c->num_compressed_bins == BINS && c->kernel == 3 && c->stride == 2 &&
         c->padding == 1 && c->num_gaussians >= 2 && c->num_gaussians <= 16;
}

// workspace: d0 | d1 | d2 | gx | hseq
size_t model_workspace_bytes(const b2d_model* m, int B, int T) {
  (void)m;
  const size_t nf = (size_t)B * T;
  return align_up(nf * D0 * 4, 256) + align_up(nf * D1 * 4, 256) + align_up(nf * D2 * 4, 256) + align_up(nf * GX * 4, 256) +
         align_up(nf * HS * 4, 256);
}

// conv_mode (low byte): B2D_CONV_FP32 = fp32 FMA kernels of this file (the exact reference engine), B2D_CONV_MMA = warp-level
// tensor-core MMAs with the fp32-class TF32 split (unet_mma.cu; the default: fastest), B2D_CONV_UTC = the fused tcgen05 / TMEM
// encoder (unet_tc.cu) + the warp-MMA decoder.
int model_forward(const b2d_model* m, const float* x, float* hx, float* pred, float* mel_bt, int fused_mode,
                  float out_scale, int B, int T, int conv_mode, void* ws, size_t ws_bytes, cudaStream_t st) {
  B2D_REQUIRE(B >= 1 && T >= 1, B2D_ERR_BAD_ARG, "GRUUNet2 forward needs B >= 1 and T >= 1 (got %d, %d)", B, T);
  B2D_REQUIRE(ws != nullptr && ws_bytes >= model_workspace_bytes(m, B, T), B2D_ERR_WORKSPACE, "GRUUNet2 workspace too small");
  const bool exact_gates = (conv_mode & B2D_CONV_EXACT_GATES) != 0;
  conv_mode &= 0xff;
  B2D_REQUIRE(conv_mode == B2D_CONV_FP32 || conv_mode == B2D_CONV_MMA || conv_mode == B2D_CONV_UTC, B2D_ERR_BAD_ARG,
              "conv_mode must be B2D_CONV_FP32 (0), B2D_CONV_MMA (3) or B2D_CONV_UTC (5), got %d", conv_mode);
  const size_t nf = (size_t)B * T;
  unsigned char* base = static_cast<unsigned char*>(ws);
  float* d0 = reinterpret_cast<float*>(base); base += align_up(nf * D0 * 4, 256);
  float* d1 = reinterpret_cast<float*>(base); base += align_up(nf * D1 * 4, 256);
  float* d2 = reinterpret_cast<float*>(base); base += align_up(nf * D2 * 4, 256);
  float* gx = reinterpret_cast<float*>(base); base += align_up(nf * GX * 4, 256);
  float* hseq = reinterpret_cast<float*>(base);
  const Packed L = packed_layout();
  int dev_sms = 148;
  if (cudaDeviceGetAttribute(&dev_sms, cudaDevAttrMultiProcessorCount, m->device) != cudaSuccess || dev_sms < 1) dev_sms = 148;
  int rc;
  if (conv_mode == B2D_CONV_FP32) {
    const size_t smem = sizeof(float) * (size_t)(((L.rec_w + 3) & ~3) + ENC_WARPS * ENC_ACT);
    B2D_SMEM_OPT_IN(smem, encoder_kernel);
    const size_t want = (nf + ENC_WARPS - 1) / ENC_WARPS;
    const int grid = (int)(want < (size_t)dev_sms * 3 ? want : (size_t)dev_sms * 3);
    encoder_kernel<<<grid, ENC_WARPS * 32, smem, st>>>(m->d_blob, x, nf, d0, d1, d2, gx);
    B2D_LAUNCH_CHECK("encoder_kernel");
  } else if (conv_mode == B2D_CONV_UTC) {
    if ((rc = model_encode_utc(m, x, nf, d0, d1, d2, gx, dev_sms, st))) return rc;
  } else {
    if ((rc = model_encode_mma(m, x, nf, d0, d1, d2, gx, 3, dev_sms, st))) return rc;
  }
  if (exact_gates) B2D_CUDA(launch_pdl(recurrence_kernel<true>, dim3(B), dim3(96), 0, st, m->d_blob, gx, hx, hseq, T));
  else B2D_CUDA(launch_pdl(recurrence_kernel<false>, dim3(B), dim3(96), 0, st, m->d_blob, gx, hx, hseq, T));
  B2D_LAUNCH_CHECK("recurrence_kernel");
  if (conv_mode == B2D_CONV_FP32) {
    const int nw = L.total - L.dec_w[0];
    const size_t smem = sizeof(float) * (size_t)(((nw + 3) & ~3) + DEC2_WARPS * DEC2_FW * DEC2_FR);
    B2D_SMEM_OPT_IN(smem, decoder2_kernel);
    const size_t want = (nf + DEC2_WARPS * DEC2_FW - 1) / (DEC2_WARPS * DEC2_FW);
    const int grid = (int)(want < (size_t)dev_sms * 2 ? want : (size_t)dev_sms * 2);
    decoder2_kernel<<<grid, DEC2_WARPS * 32, smem, st>>>(m->d_blob, hseq, d0, d1, d2, x, nf, pred, mel_bt, fused_mode, out_scale);
    B2D_LAUNCH_CHECK("decoder2_kernel");
  } else {
    if ((rc = model_decode_mma(m, hseq, d0, d1, d2, x, nf, pred, mel_bt, fused_mode, out_scale, 3, dev_sms, st))) return rc;
  }
  return B2D_OK;
}

}  // namespace b2d
