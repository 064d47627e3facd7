// unet_mma.cu -- GRUUNet2 decoder (output_gate UpBlocks, gruunet2.py:184-199) on warp-level tensor-core MMAs.
//
// Why warp-level mma.sync is the default engine and not tcgen05: the U-Net has 17 channels.  One ConvTranspose1d(k3, s2, p1, op1) is
//   out[2i]   = pb + x[i] W1,      out[2i+1] = pb + x[i] W2 + x[i+1] W0        (x[i]: the Cin-vector at input position i)
// i.e. per frame a [Lin x Cin] x [Cin x 17] GEMM with Lin = 4..32 rows.  tcgen05 needs 128-row operand tiles staged in shared
// memory in core-matrix layout and a TMEM round trip per layer.  The fair trial of that design is unet_tc.cu: all encoder
// layers fused in one persistent kernel, activations written from the epilogue straight into the next layer's operand, weights
// by TMA, two tile slots per SM -- correct to 4e-6 and 120-130 us against this file's 98 us (tensor pipe 17 % busy, issue slots
// 17 %: the time goes into the per-layer hand-offs, fence -> barrier -> one thread issuing 24 MMAs -> commit -> mbarrier ->
// tcgen05.ld, of which only two run per SM at a time because a tile slot's A operand with its big + small planes takes 64 KB).
// A warp-level m16n8k8 MMA takes its operands straight from registers, so a warp keeps two frames' activations in its own
// shared-memory rows ([position][channel], stride 44 floats: conflict-free fragment loads) and walks the four layers without
// leaving the SM or waiting for any other warp.  Measured on B200: 919 m16n8k8 TF32 MMAs / us / SM (tools/micro/mma_rate.cu),
// 3.9x the FP32 FMA peak -- and, more to the point, one shared-memory operand fetch feeds 1024 MACs instead of 128.
//
// Precision: every operand is split into TF32 big + small parts (big = round-to-nearest TF32, small = exact remainder)
// and three MMAs are issued per tile (big*big + small*big + big*small), fp32 accumulate: fp32-class results (model parity
// < 1e-5, same as the FMA kernels).
//
// Weight fragments are laid out at model-pack time exactly as the B operand registers want them:
//   float4 (b0_big, b1_big, b0_small, b1_small) per lane per (layer, operand set, k-step, n-tile).
#include <stdlib.h>
#include <string.h>
#include <vector>

#include "kernels.cuh"
#include "model_layout.cuh"

namespace b2d {
namespace umma {

constexpr int WARPS = 8;
constexpr int G = 2;    // frames per warp iteration
// Activation row stride (floats): 36 mod 32 = 4 -> an A fragment's 8 rows x 4 columns hit 32 distinct banks.  A row holds
// channels 0..33; the k-steps read up to channel 39, i.e. into the first floats of the next row: finite values that meet
// all-zero weight rows.  Every layer input has its own buffer, so the zero rows that close each frame and the unused
// channels are written once (at kernel start) and never again.
constexpr int S = 36;
constexpr int ROWS0 = G * 5, ROWS1 = G * 9, ROWS2 = G * 17, ROWS3 = G * 33;
constexpr int WARP_FLOATS = (ROWS0 + ROWS1 + ROWS2 + ROWS3 + 1) * S + G * 64;  // four layer inputs (+1 row of slack) | output staging
// decoder fragment counts: layers 0..2 have three operand sets (even | odd from x[i] | odd from x[i+1]) x k-steps x 3 n-tiles
__host__ __device__ constexpr int dks(int l) { return l == 0 ? 3 : 5; }
__host__ __device__ constexpr int dfrag_off(int l) { return l == 0 ? 0 : l == 1 ? 27 : l == 2 ? 72 : l == 3 ? 117 : 127; }
__host__ __device__ constexpr int dpb_off(int l) { return l == 0 ? 0 : l == 1 ? 8 * HP : l == 2 ? 24 * HP : 56 * HP; }
constexpr int DPB_FLOATS = 56 * HP + 64 * 4;

// round-to-nearest (ties away) TF32 of a finite value: two integer instructions (cvt.rna.tf32.f32 expands to five with
// its NaN / infinity handling, and activations here are finite)
__device__ __forceinline__ uint32_t tf32_rna(float v) { return (__float_as_uint(v) + 0x1000u) & 0xFFFFE000u; }
__device__ __forceinline__ void mma_tf32(float* c, const uint32_t* a, uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
// D += A * B with both operands split: big*big + small*big + big*small (TERMS == 3) or big*big only
template <int TERMS>
__device__ __forceinline__ void mma_split(float* c, const uint32_t* abig, const uint32_t* asmall, const float4 bf) {
  const uint32_t b0 = __float_as_uint(bf.x), b1 = __float_as_uint(bf.y);
  mma_tf32(c, abig, b0, b1);
  if (TERMS == 3) {
    mma_tf32(c, asmall, b0, b1);
    mma_tf32(c, abig, __float_as_uint(bf.z), __float_as_uint(bf.w));
  }
}
// A fragment of one 16-row tile at k-step ks: rows r0 / r1 are float offsets of tile rows g / g + 8
template <int TERMS>
__device__ __forceinline__ void load_a(const float* buf, int r0, int r1, int ks, int t, uint32_t* big, uint32_t* small) {
  const float v[4] = {buf[r0 + 8 * ks + t], buf[r1 + 8 * ks + t], buf[r0 + 8 * ks + t + 4], buf[r1 + 8 * ks + t + 4]};
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    big[e] = tf32_rna(v[e]);
    if (TERMS == 3) small[e] = __float_as_uint(v[e] - __uint_as_float(big[e]));
  }
}

// One up-block for NT 16-row tiles at once.  in: rows frame * (LIN + 1) + i, out: rows frame * (2 LIN + 1) + o (channels 0..16).
template <int LIN, int KS, int NT, int TERMS>
__device__ __forceinline__ void up_layer(const float* __restrict__ in, float* __restrict__ out, const float4* __restrict__ FR,
                                         const float* __restrict__ PB, int tile0, int lane) {
  const int g = lane >> 2, t = lane & 3;
  int r0[NT], r1[NT];
  bool ok0[NT], ok1[NT];
#pragma unroll
  for (int m = 0; m < NT; ++m) {
    const int R0 = (tile0 + m) * 16 + g, R1 = R0 + 8;
    ok0[m] = R0 < G * LIN;
    ok1[m] = R1 < G * LIN;
    r0[m] = ok0[m] ? ((R0 / LIN) * (LIN + 1) + (R0 % LIN)) * S : 0;
    r1[m] = ok1[m] ? ((R1 / LIN) * (LIN + 1) + (R1 % LIN)) * S : 0;
  }
  float ce[NT][3][4], co[NT][3][4];
#pragma unroll
  for (int m = 0; m < NT; ++m)
#pragma unroll
    for (int n = 0; n < 3; ++n)
#pragma unroll
      for (int e = 0; e < 4; ++e) { ce[m][n][e] = 0.f; co[m][n][e] = 0.f; }
#pragma unroll
  for (int ks = 0; ks < KS; ++ks) {
    uint32_t a0b[NT][4], a0s[NT][4], a1b[NT][4], a1s[NT][4];
#pragma unroll
    for (int m = 0; m < NT; ++m) {
      load_a<TERMS>(in, r0[m], r1[m], ks, t, a0b[m], a0s[m]);          // x[i]
      load_a<TERMS>(in, r0[m] + S, r1[m] + S, ks, t, a1b[m], a1s[m]);  // x[i + 1] (a zero row closes every frame)
    }
#pragma unroll
    for (int n = 0; n < 3; ++n) {
      const float4 fe = FR[((0 * KS + ks) * 3 + n) * 32 + lane];
      const float4 f0 = FR[((1 * KS + ks) * 3 + n) * 32 + lane];
      const float4 f1 = FR[((2 * KS + ks) * 3 + n) * 32 + lane];
#pragma unroll
      for (int m = 0; m < NT; ++m) {
        mma_split<TERMS>(ce[m][n], a0b[m], a0s[m], fe);
        mma_split<TERMS>(co[m][n], a0b[m], a0s[m], f0);
        mma_split<TERMS>(co[m][n], a1b[m], a1s[m], f1);
      }
    }
  }
  // epilogue: relu(acc + position bias) -> next layer's rows; c0/c1: tile row g, c2/c3: row g + 8; columns 8 n + 2 t (+1)
#pragma unroll
  for (int m = 0; m < NT; ++m) {
#pragma unroll
    for (int hrow = 0; hrow < 2; ++hrow) {
      const int R = (tile0 + m) * 16 + g + 8 * hrow;
      if (!(hrow ? ok1[m] : ok0[m])) continue;
      const int fl = R / LIN, i = R % LIN;
#pragma unroll
      for (int par = 0; par < 2; ++par) {
        const int o = 2 * i + par;
        float* dst = out + (fl * (2 * LIN + 1) + o) * S;
        const float* pb = PB + o * HP;
#pragma unroll
        for (int n = 0; n < 3; ++n) {
          const int c = 8 * n + 2 * t;
          const float* acc = par ? co[m][n] : ce[m][n];
          const float v0 = fmaxf(acc[2 * hrow] + pb[c], 0.f);
          if (c + 1 < H) {
            const float v1 = fmaxf(acc[2 * hrow + 1] + pb[c + 1], 0.f);
            *reinterpret_cast<float2*>(dst + c) = make_float2(v0, v1);
          } else if (c < H) {
            dst[c] = v0;
          }
        }
      }
    }
  }
}

// last up-block (one output channel): columns 0 / 1 of the single n-tile are the even / odd output of input position i
template <int NT, int TERMS>
__device__ __forceinline__ void last_layer(const float* __restrict__ in, float* __restrict__ outbuf, const float4* __restrict__ FR,
                                           const float* __restrict__ PB, int lane) {
  constexpr int LIN = 32, KS = 5;
  const int g = lane >> 2, t = lane & 3;
  float c[NT][4];
  int r0[NT], r1[NT];
#pragma unroll
  for (int m = 0; m < NT; ++m) {
    const int R0 = m * 16 + g, R1 = R0 + 8;
    r0[m] = ((R0 / LIN) * (LIN + 1) + (R0 % LIN)) * S;
    r1[m] = ((R1 / LIN) * (LIN + 1) + (R1 % LIN)) * S;
#pragma unroll
    for (int e = 0; e < 4; ++e) c[m][e] = 0.f;
  }
#pragma unroll
  for (int ks = 0; ks < KS; ++ks) {
    const float4 f0 = FR[(0 * KS + ks) * 32 + lane], f1 = FR[(1 * KS + ks) * 32 + lane];
#pragma unroll
    for (int m = 0; m < NT; ++m) {
      uint32_t ab[4], as[4];
      load_a<TERMS>(in, r0[m], r1[m], ks, t, ab, as);
      mma_split<TERMS>(c[m], ab, as, f0);
      load_a<TERMS>(in, r0[m] + S, r1[m] + S, ks, t, ab, as);
      mma_split<TERMS>(c[m], ab, as, f1);
    }
  }
  if (t == 0) {
#pragma unroll
    for (int m = 0; m < NT; ++m)
#pragma unroll
      for (int hrow = 0; hrow < 2; ++hrow) {
        const int R = m * 16 + g + 8 * hrow;
        const int fl = R / LIN, i = R % LIN;
        const float v0 = c[m][2 * hrow] + PB[(2 * i) * 4], v1 = c[m][2 * hrow + 1] + PB[(2 * i + 1) * 4];
        *reinterpret_cast<float2*>(outbuf + fl * NMEL + 2 * i) = make_float2(v0, v1);
      }
  }
}

// TMA bulk copy of the weight fragment image into shared memory (one instruction instead of a 16-round load / store loop:
// it matters for the streaming hop, where a launch processes three frames and the prologue is most of the kernel)
__device__ __forceinline__ uint32_t um_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void um_bulk_load(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(um_u32(bar)) : "memory");
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(um_u32(bar)), "r"(bytes) : "memory");
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(um_u32(dst)), "l"(src),
               "r"(bytes), "r"(um_u32(bar))
               : "memory");
}
__device__ __forceinline__ void um_bulk_wait(uint64_t* bar) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "W_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n"
      "@p bra D_%=;\n"
      "bra W_%=;\n"
      "D_%=:\n"
      "}\n" ::"r"(um_u32(bar))
      : "memory");
}

__device__ __forceinline__ void prefetch_range(const float* p, int bytes, int lane) {  // one 128-byte line per lane per round
  for (int o = lane * 128; o < bytes; o += 32 * 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(reinterpret_cast<const char*>(p) + o));
}

// 4-byte asynchronous global -> shared copy (LDGSTS) and its group bookkeeping
__device__ __forceinline__ void cp_async4(float* smem_dst, const float* gsrc) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(static_cast<uint32_t>(__cvta_generic_to_shared(smem_dst))), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

template <int TERMS>
__global__ void __launch_bounds__(WARPS * 32, 1) decoder_mma_kernel(const float4* __restrict__ frags, const float* __restrict__ blob,
                                                                    const float* __restrict__ hseq, const float* __restrict__ d0,
                                                                    const float* __restrict__ d1, const float* __restrict__ d2,
                                                                    const float* __restrict__ x, size_t nframes, float* __restrict__ pred,
                                                                    float* __restrict__ mel, int fused_mode, float out_scale) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float4* FR = reinterpret_cast<float4*>(smem_raw);                    // [127][32] weight fragments
  float* PB = reinterpret_cast<float*>(FR + dfrag_off(4) * 32);        // position biases of the four layers
  float* act = PB + DPB_FLOATS;
  const Packed L = packed_layout();
  __shared__ uint64_t frag_bar;
  if (threadIdx.x == 0) um_bulk_load(FR, frags, dfrag_off(4) * 32 * 16, &frag_bar);
  for (int l = 0; l < 4; ++l) {
    const int n = (l < 3 ? (8 << l) * HP : 64 * 4);
    for (int i = threadIdx.x; i < n; i += blockDim.x) PB[dpb_off(l) + i] = blob[L.dec_pb[l] + i];
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (((size_t)blockIdx.x * WARPS + warp) * G < nframes)  // only warps with work: padding channels / rows stay zero for good
    for (int i = lane; i < WARP_FLOATS; i += 32) act[warp * WARP_FLOATS + i] = 0.f;
  __syncthreads();
  um_bulk_wait(&frag_bar);
  pdl_wait();  // weights and biases staged; activations come from the kernels before (common.cuh: programmatic dependent launch)
  pdl_trigger();
  float* B0 = act + warp * WARP_FLOATS;  // layer inputs: rows fl * (Lin + 1) + i
  float* B1 = B0 + ROWS0 * S;
  float* B2 = B1 + ROWS1 * S;
  float* B3 = B2 + ROWS2 * S;
  float* OUT = B3 + (ROWS3 + 1) * S;
  const size_t stride = (size_t)gridDim.x * WARPS * G;
  // skip tensors are [frame][channel][position] in HBM; a warp step moves 8 positions x 4 channels: whole 32-byte sectors in,
  // conflict-free rows out (bank = 4 i + channel)
  const int si = lane & 7, sc = lane >> 3;
  for (size_t f0 = ((size_t)blockIdx.x * WARPS + warp) * G; f0 < nframes; f0 += stride) {
    const int nf = (int)min((size_t)G, nframes - f0);
    if (f0 + stride + G <= nframes) {  // next iteration's inputs -> L2 while this one computes
      const size_t fn = f0 + stride;
      prefetch_range(d0 + fn * D0, G * D0 * 4, lane);
      prefetch_range(d1 + fn * D1, G * D1 * 4, lane);
      prefetch_range(d2 + fn * D2, G * D2 * 4, lane);
      prefetch_range(hseq + fn * HS, G * HS * 4, lane);
      prefetch_range(x + fn * NMEL, G * NMEL * 4, lane);
    }
    // ---- stage everything this frame pair needs (all loads are independent: one exposed latency per iteration) ----------
    // Every input element goes global -> shared memory with a 4-byte cp.async (the [channel][position] -> [position][channel]
    // transpose happens in the destination address), in four commit groups in the order the layers need them: the hidden
    // state before layer 0, skip d2 before layer 1, d1 before layer 2, d0 -- half of the bytes -- before the last layer.  Only
    // the first group's latency is exposed; round 1 staged everything with loads + stores up front (ncu: long_scoreboard 1.3 per issue).
    for (int e = lane; e < G * HS; e += 32) {
      const int fl = e / HS, r = e - fl * HS, ch = r >> 2, i = r & 3;
      float* dst = &B0[(fl * 5 + i) * S + ch];
      if (fl < nf) cp_async4(dst, hseq + (f0 + fl) * HS + r); else *dst = 0.f;
    }
    cp_async_commit();
#pragma unroll
    for (int fl = 0; fl < G; ++fl) {
      const bool live = fl < nf;
      const float* s2 = d2 + (f0 + fl) * D2;
#pragma unroll
      for (int cg = 0; cg < 5; ++cg) {
        const int ch = 4 * cg + sc;
        if (ch < H) {
          float* dst = &B1[(fl * 9 + si) * S + H + ch];
          if (live) cp_async4(dst, s2 + ch * 8 + si); else *dst = 0.f;
        }
      }
    }
    cp_async_commit();
#pragma unroll
    for (int fl = 0; fl < G; ++fl) {
      const bool live = fl < nf;
      const float* s1 = d1 + (f0 + fl) * D1;
#pragma unroll
      for (int cg = 0; cg < 5; ++cg) {
        const int ch = 4 * cg + sc;
        if (ch < H) {
#pragma unroll
          for (int pg = 0; pg < 2; ++pg) {
            float* dst = &B2[(fl * 17 + 8 * pg + si) * S + H + ch];
            if (live) cp_async4(dst, s1 + ch * 16 + 8 * pg + si); else *dst = 0.f;
          }
        }
      }
    }
    cp_async_commit();
#pragma unroll
    for (int fl = 0; fl < G; ++fl) {
      const bool live = fl < nf;
      const float* s0 = d0 + (f0 + fl) * D0;
#pragma unroll
      for (int cg = 0; cg < 5; ++cg) {
        const int ch = 4 * cg + sc;
        if (ch < H) {
#pragma unroll
          for (int pg = 0; pg < 4; ++pg) {
            float* dst = &B3[(fl * 33 + 8 * pg + si) * S + H + ch];
            if (live) cp_async4(dst, s0 + ch * 32 + 8 * pg + si); else *dst = 0.f;
          }
        }
      }
    }
    cp_async_commit();
    cp_async_wait<3>();
    __syncwarp();
    up_layer<4, 3, 1, TERMS>(B0, B1, FR + dfrag_off(0) * 32, PB + dpb_off(0), 0, lane);
    cp_async_wait<2>();
    __syncwarp();
    up_layer<8, 5, 1, TERMS>(B1, B2, FR + dfrag_off(1) * 32, PB + dpb_off(1), 0, lane);
    cp_async_wait<1>();
    __syncwarp();
    up_layer<16, 5, 2, TERMS>(B2, B3, FR + dfrag_off(2) * 32, PB + dpb_off(2), 0, lane);
    cp_async_wait<0>();
    __syncwarp();
    last_layer<4, TERMS>(B3, OUT, FR + dfrag_off(3) * 32, PB + dpb_off(3), lane);
    __syncwarp();
    for (int e = lane; e < nf * NMEL; e += 32) {
      const size_t idx = f0 * NMEL + e;
      const float p = OUT[e];
      pred[idx] = p;
      if (fused_mode) {
        const float xv = x[idx];
        float v;
        if (fused_mode == 1) {
          float r = xv - p;
          r = r > 0.f ? r : 0.2f * r;
          v = fmaxf(expm1f(r), 0.f);
        } else {
          v = expf(xv - fmaxf(p, 0.f) * out_scale) - 1.0f;
        }
        mel[idx] = v;
      }
    }
    __syncwarp();
  }
}

// ====================================================================================================================
// Encoder (input_gate DownBlocks, gruunet2.py:127-144): Conv1d(k3, s2, p1), out[j] = relu(pb[j] + sum_k x[2j-1+k] W_k).
// Rows = output positions, one K-block per tap.  A layer input lives in rows [position + 1][channel] (row 0 = the zero
// left of the frame) with stride ES = 18 floats: consecutive output positions are 2 rows = 36 floats apart, so the 8 rows
// x 4 columns of an A fragment again hit 32 distinct banks.  Layer 0 has one input channel: its K-block is the three taps.
// ====================================================================================================================
constexpr int EWARPS = 12;
constexpr int ES = 18;
constexpr int EX = 72;                                   // padded input frame: [0] = 0, [1..64] = x, rest 0
constexpr int ER0 = 33, ER1 = 17, ER2 = 9;               // rows per frame of the inputs of layers 1, 2, 3
constexpr int EWARP_FLOATS = G * EX + (G * (ER0 + ER1 + ER2) + 2) * ES;
__host__ __device__ constexpr int efrag_off(int l) { return l == 0 ? 0 : l == 1 ? 3 : l == 2 ? 30 : l == 3 ? 57 : 120; }
__host__ __device__ constexpr int epb_off(int l) { return l == 0 ? 0 : l == 1 ? 32 * HP : l == 2 ? 48 * HP : 56 * HP; }
constexpr int EPB_FLOATS = 56 * HP + 4 * H3P;

// LOUT output positions per frame, NN n-tiles, NT tiles at once; FIRST: layer 0 (A straight from the padded input frame).
// Results go to the next layer's rows (act_out, may be null) and to HBM as [frame][channel][position] (gout).
template <int LOUT, int NN, int NT, int COUT, int COP, bool FIRST, int TERMS>
__device__ __forceinline__ void down_layer(const float* __restrict__ in, float* __restrict__ act_out, float* __restrict__ gout,
                                           const float4* __restrict__ FR, const float* __restrict__ PB, int nf, int lane) {
  constexpr int LIN = 2 * LOUT;
  constexpr int KS = FIRST ? 1 : 3;  // k-steps per tap block
  const int g = lane >> 2, t = lane & 3;
  int r0[NT], r1[NT];
  bool ok0[NT], ok1[NT];
#pragma unroll
  for (int m = 0; m < NT; ++m) {
    const int R0 = m * 16 + g, R1 = R0 + 8;
    ok0[m] = R0 < G * LOUT;
    ok1[m] = R1 < G * LOUT;
    const int f0 = ok0[m] ? R0 / LOUT : 0, j0 = ok0[m] ? R0 % LOUT : 0, f1 = ok1[m] ? R1 / LOUT : 0, j1 = ok1[m] ? R1 % LOUT : 0;
    r0[m] = FIRST ? f0 * EX + 2 * j0 : (f0 * (LIN + 1) + 2 * j0) * ES;
    r1[m] = FIRST ? f1 * EX + 2 * j1 : (f1 * (LIN + 1) + 2 * j1) * ES;
  }
  float c[NT][NN][4];
#pragma unroll
  for (int m = 0; m < NT; ++m)
#pragma unroll
    for (int n = 0; n < NN; ++n)
#pragma unroll
      for (int e = 0; e < 4; ++e) c[m][n][e] = 0.f;
  if (FIRST) {
    uint32_t ab[NT][4], as[NT][4];
#pragma unroll
    for (int m = 0; m < NT; ++m) load_a<TERMS>(in, r0[m], r1[m], 0, t, ab[m], as[m]);  // columns = taps: x[2j - 1 + t]
#pragma unroll
    for (int n = 0; n < NN; ++n) {
      const float4 f = FR[n * 32 + lane];
#pragma unroll
      for (int m = 0; m < NT; ++m) mma_split<TERMS>(c[m][n], ab[m], as[m], f);
    }
  } else {
#pragma unroll
    for (int tap = 0; tap < 3; ++tap)
#pragma unroll
      for (int ks = 0; ks < KS; ++ks) {
        uint32_t ab[NT][4], as[NT][4];
#pragma unroll
        for (int m = 0; m < NT; ++m) load_a<TERMS>(in, r0[m] + tap * ES, r1[m] + tap * ES, ks, t, ab[m], as[m]);
#pragma unroll
        for (int n = 0; n < NN; ++n) {
          const float4 f = FR[((tap * KS + ks) * NN + n) * 32 + lane];
#pragma unroll
          for (int m = 0; m < NT; ++m) mma_split<TERMS>(c[m][n], ab[m], as[m], f);
        }
      }
  }
#pragma unroll
  for (int m = 0; m < NT; ++m)
#pragma unroll
    for (int hrow = 0; hrow < 2; ++hrow) {
      if (!(hrow ? ok1[m] : ok0[m])) continue;
      const int R = m * 16 + g + 8 * hrow;
      const int fl = R / LOUT, j = R % LOUT;
      const float* pb = PB + j * COP;
#pragma unroll
      for (int n = 0; n < NN; ++n) {
        const int co = 8 * n + 2 * t;
        const float v0 = fmaxf(c[m][n][2 * hrow] + pb[co < COP ? co : 0], 0.f);
        const float v1 = fmaxf(c[m][n][2 * hrow + 1] + pb[co + 1 < COP ? co + 1 : 0], 0.f);
        if (act_out != nullptr) {
          float* dst = act_out + (fl * (LOUT + 1) + j + 1) * ES;
          if (co + 1 < COUT) *reinterpret_cast<float2*>(dst + co) = make_float2(v0, v1);
          else if (co < COUT) dst[co] = v0;
        }
        if (fl < nf) {
          float* gd = gout + (size_t)fl * COUT * LOUT + j;
          if (co < COUT) gd[co * LOUT] = v0;
          if (co + 1 < COUT) gd[(co + 1) * LOUT] = v1;
        }
      }
    }
}

template <int TERMS>
__global__ void __launch_bounds__(EWARPS * 32, 1) encoder_mma_kernel(const float4* __restrict__ frags, const float* __restrict__ blob,
                                                                     const float* __restrict__ x, size_t nframes, float* __restrict__ d0,
                                                                     float* __restrict__ d1, float* __restrict__ d2, float* __restrict__ gx) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float4* FR = reinterpret_cast<float4*>(smem_raw);
  float* PB = reinterpret_cast<float*>(FR + efrag_off(4) * 32);
  float* act = PB + EPB_FLOATS;
  const Packed L = packed_layout();
  __shared__ uint64_t frag_bar;
  if (threadIdx.x == 0) um_bulk_load(FR, frags, efrag_off(4) * 32 * 16, &frag_bar);
  for (int l = 0; l < 4; ++l) {
    const int n = (l < 3 ? (32 >> l) * HP : 4 * H3P);
    for (int i = threadIdx.x; i < n; i += blockDim.x) PB[epb_off(l) + i] = blob[L.enc_pb[l] + i];
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (((size_t)blockIdx.x * EWARPS + warp) * G < nframes)  // only warps with work: zero rows / unused channels stay zero
    for (int i = lane; i < EWARP_FLOATS + 8; i += 32) act[warp * EWARP_FLOATS + i] = 0.f;
  __syncthreads();
  um_bulk_wait(&frag_bar);
  pdl_wait();
  pdl_trigger();
  float* X = act + warp * EWARP_FLOATS;
  float* A0 = X + G * EX;            // layer-1 input rows
  float* A1 = A0 + G * ER0 * ES;
  float* A2 = A1 + G * ER1 * ES;
  const size_t stride = (size_t)gridDim.x * EWARPS * G;
  // (Measured, round 2: the input frames of trip k + 1 by cp.async into a second X buffer while trip k computes: 99 us instead of
  //  98 us -- the L2 prefetch below already hides that latency; the same idea is worth 23 us in the decoder, whose inputs are 18 x larger.)
  for (size_t f0 = ((size_t)blockIdx.x * EWARPS + warp) * G; f0 < nframes; f0 += stride) {
    const int nf = (int)min((size_t)G, nframes - f0);
    if (f0 + stride + G <= nframes) prefetch_range(x + (f0 + stride) * NMEL, G * NMEL * 4, lane);
#pragma unroll
    for (int fl = 0; fl < G; ++fl) {
      X[fl * EX + 1 + lane] = fl < nf ? x[(f0 + fl) * NMEL + lane] : 0.f;
      X[fl * EX + 33 + lane] = fl < nf ? x[(f0 + fl) * NMEL + 32 + lane] : 0.f;
    }
    __syncwarp();
    down_layer<32, 3, 4, H, HP, true, TERMS>(X, A0, d0 + f0 * D0, FR + efrag_off(0) * 32, PB + epb_off(0), nf, lane);
    __syncwarp();
    down_layer<16, 3, 2, H, HP, false, TERMS>(A0, A1, d1 + f0 * D1, FR + efrag_off(1) * 32, PB + epb_off(1), nf, lane);
    __syncwarp();
    down_layer<8, 3, 1, H, HP, false, TERMS>(A1, A2, d2 + f0 * D2, FR + efrag_off(2) * 32, PB + epb_off(2), nf, lane);
    __syncwarp();
    down_layer<4, 7, 1, H3, H3P, false, TERMS>(A2, nullptr, gx + f0 * GX, FR + efrag_off(3) * 32, PB + epb_off(3), nf, lane);
    __syncwarp();
  }
}

// ---- host: weight fragment images ---------------------------------------------------------------------------------
static float tf32_rna_host(float v) {
  uint32_t u;
  memcpy(&u, &v, 4);
  u = (u + 0x1000u) & 0xFFFFE000u;  // round to nearest, ties away from zero (cvt.rna.tf32.f32)
  float r;
  memcpy(&r, &u, 4);
  return r;
}
static void put_frag(std::vector<float>& img, int frag, int lane, float b0, float b1) {
  float* p = img.data() + ((size_t)frag * 32 + lane) * 4;
  p[0] = tf32_rna_host(b0);
  p[1] = tf32_rna_host(b1);
  p[2] = b0 - p[0];
  p[3] = b1 - p[1];
}

}  // namespace umma

// Builds the decoder fragment image from the packed fp32 blob (dec W[ci][k][coP]) and uploads it.
int model_pack_mma(b2d_model* m) {
  using namespace umma;
  const Packed L = packed_layout();
  const float* blob = m->h_blob;
  B2D_REQUIRE(blob != nullptr, B2D_ERR_CUDA, "host copy of the packed weights is missing");
  std::vector<float> img((size_t)dfrag_off(4) * 32 * 4, 0.f);
  const int cin[4] = {H, 2 * H, 2 * H, 2 * H};
  const int cop[4] = {HP, HP, HP, 4};
  for (int l = 0; l < 4; ++l) {
    const float* W = blob + L.dec_w[l];
    auto w = [&](int ci, int tap, int co) { return (ci < cin[l]) ? W[(ci * 3 + tap) * cop[l] + co] : 0.f; };
    for (int ks = 0; ks < dks(l); ++ks)
      for (int lane = 0; lane < 32; ++lane) {
        const int g = lane >> 2, t = lane & 3;
        const int k0 = 8 * ks + t, k1 = k0 + 4;
        if (l < 3) {
          const int tap_of_set[3] = {1, 2, 0};  // even <- x[i] W1 ; odd <- x[i] W2 + x[i+1] W0
          for (int set = 0; set < 3; ++set)
            for (int n = 0; n < 3; ++n) {
              const int co = 8 * n + g;
              const float b0 = co < H ? w(k0, tap_of_set[set], co) : 0.f, b1 = co < H ? w(k1, tap_of_set[set], co) : 0.f;
              put_frag(img, dfrag_off(l) + (set * dks(l) + ks) * 3 + n, lane, b0, b1);
            }
        } else {
          // column g of the single n-tile: 0 = even output (x[i] W1), 1 = odd output (x[i] W2 | x[i+1] W0)
          float a0 = 0.f, a1 = 0.f, c0 = 0.f, c1 = 0.f;
          if (g == 0) { a0 = w(k0, 1, 0); a1 = w(k1, 1, 0); }
          if (g == 1) { a0 = w(k0, 2, 0); a1 = w(k1, 2, 0); c0 = w(k0, 0, 0); c1 = w(k1, 0, 0); }
          put_frag(img, dfrag_off(l) + 0 * dks(l) + ks, lane, a0, a1);
          put_frag(img, dfrag_off(l) + 1 * dks(l) + ks, lane, c0, c1);
        }
      }
  }
  // encoder: W[ci][k][coP]; layer 0 (one input channel): K index = tap; layers 1..3: one K-block (3 k-steps) per tap
  const size_t ebase = img.size();
  img.resize(ebase + (size_t)efrag_off(4) * 32 * 4, 0.f);
  std::vector<float> eimg((size_t)efrag_off(4) * 32 * 4, 0.f);
  const int ecop[4] = {HP, HP, HP, H3P};
  const int ecout[4] = {H, H, H, H3};
  const int enn[4] = {3, 3, 3, 7};
  for (int l = 0; l < 4; ++l) {
    const float* W = blob + L.enc_w[l];
    for (int lane = 0; lane < 32; ++lane) {
      const int g = lane >> 2, t = lane & 3;
      for (int n = 0; n < enn[l]; ++n) {
        const int co = 8 * n + g;
        if (l == 0) {
          const float b0 = (co < ecout[l] && t < 3) ? W[(0 * 3 + t) * ecop[l] + co] : 0.f;
          put_frag(eimg, efrag_off(0) + n, lane, b0, 0.f);
        } else {
          for (int tap = 0; tap < 3; ++tap)
            for (int ks = 0; ks < 3; ++ks) {
              const int k0 = 8 * ks + t, k1 = k0 + 4;
              const float b0 = (co < ecout[l] && k0 < H) ? W[(k0 * 3 + tap) * ecop[l] + co] : 0.f;
              const float b1 = (co < ecout[l] && k1 < H) ? W[(k1 * 3 + tap) * ecop[l] + co] : 0.f;
              put_frag(eimg, efrag_off(l) + (tap * 3 + ks) * enn[l] + n, lane, b0, b1);
            }
        }
      }
    }
  }
  memcpy(img.data() + ebase, eimg.data(), eimg.size() * sizeof(float));
  B2D_CUDA(cudaMalloc(&m->d_mma, img.size() * sizeof(float)));
  B2D_CUDA(cudaMemcpy(m->d_mma, img.data(), img.size() * sizeof(float), cudaMemcpyHostToDevice));
  return B2D_OK;
}

int model_encode_mma(const b2d_model* m, const float* x, size_t nframes, float* d0, float* d1, float* d2, float* gx, int terms,
                     int num_sms, cudaStream_t st) {
  using namespace umma;
  B2D_REQUIRE(m->d_mma != nullptr, B2D_ERR_CUDA, "tensor-core weight fragments are missing");
  const size_t smem = (size_t)efrag_off(4) * 32 * 16 + sizeof(float) * (EPB_FLOATS + (size_t)EWARPS * EWARP_FLOATS) + 64;  // + slack behind the last warp
  const size_t want = (nframes + EWARPS * G - 1) / (EWARPS * G);
  const int grid = (int)(want < (size_t)num_sms ? want : (size_t)num_sms);
  const float4* fr = reinterpret_cast<const float4*>(m->d_mma) + (size_t)dfrag_off(4) * 32;
  (void)terms;  // always the fp32-class 3 x TF32 split
  B2D_SMEM_OPT_IN(smem, encoder_mma_kernel<3>);
  B2D_CUDA(launch_pdl(encoder_mma_kernel<3>, dim3(grid), dim3(EWARPS * 32), smem, st, fr, m->d_blob, x, nframes, d0, d1, d2, gx));
  B2D_LAUNCH_CHECK("encoder_mma_kernel");
  return B2D_OK;
}

int model_decode_mma(const b2d_model* m, const float* hseq, const float* d0, const float* d1, const float* d2, const float* x,
                     size_t nframes, float* pred, float* mel, int fused_mode, float out_scale, int terms, int num_sms, cudaStream_t st) {
  using namespace umma;
  B2D_REQUIRE(m->d_mma != nullptr, B2D_ERR_CUDA, "tensor-core weight fragments are missing");
  const size_t smem = (size_t)dfrag_off(4) * 32 * 16 + sizeof(float) * (DPB_FLOATS + (size_t)WARPS * WARP_FLOATS);
  const size_t want = (nframes + WARPS * G - 1) / (WARPS * G);
  const int grid = (int)(want < (size_t)num_sms ? want : (size_t)num_sms);
  const float4* fr = reinterpret_cast<const float4*>(m->d_mma);
  (void)terms;
  B2D_SMEM_OPT_IN(smem, decoder_mma_kernel<3>);
  B2D_CUDA(launch_pdl(decoder_mma_kernel<3>, dim3(grid), dim3(WARPS * 32), smem, st, fr, m->d_blob, hseq, d0, d1, d2, x, nframes, pred, mel,
                      fused_mode, out_scale));
  B2D_LAUNCH_CHECK("decoder_mma_kernel");
  return B2D_OK;
}

}  // namespace b2d
