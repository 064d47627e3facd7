// unet_tc.cu -- the GRUUNet2 encoder (gruunet2.py:71-79, 127-157) as ONE persistent tcgen05 / TMEM kernel.
//
// The four strided convolutions are implicit GEMMs  D[rows, N] = A[rows, K] * B[N, K]^T  with rows = (frame, output
// position), K = (tap, input channel), N = output channels:
//     layer   rows / frame   K (padded)      N (padded)
//       0         32         3  (8)          17 (32)
//       1         16         3 x 20 (64)     17 (32)
//       2          8         3 x 20 (64)     17 (32)
//       3          4         3 x 20 (64)     51 (64)
// A tile is 8 frames: 256 / 128 / 64 / 32 rows, i.e. two, one, half and a quarter of a 128-row UMMA tile.
// All four layers of a tile run back to back inside the SM: the weight images (TF32 big | small parts, canonical K-major
// core-matrix layout, 66 KB) arrive once per CTA by TMA and stay; a layer's accumulator comes back from TMEM with
// tcgen05.ld, gets its position bias + ReLU, goes to HBM once (the skip tensor the decoder needs) and is written STRAIGHT
// INTO THE NEXT LAYER'S A OPERAND in shared memory, in the canonical layout, already split into big + small TF32 parts
// (position p feeds tap 1 of row p / 2 when even, tap 2 of row (p - 1) / 2 and tap 0 of row (p + 1) / 2 when odd): no
// im2col gather, no HBM round trip between layers.  fp32-class accuracy from TF32 tensor cores: three MMAs per k-step
// (big*big + small*big + big*small), fp32 accumulate in TMEM.
// Two warpgroups per CTA each own a tile slot (their own A buffers, TMEM columns and mbarrier) and run the same sequential
// chain on alternate tiles, so one group's epilogue overlaps the other's MMAs; one elected thread per group issues.
#include <vector>

#include "kernels.cuh"
#include "model_layout.cuh"
#include "tcgen05.cuh"

namespace b2d {
namespace utc {

using namespace tc5;

constexpr int FT = 8;                   // frames per tile
constexpr int K0 = 8, K1 = 64, CH = 20; // K of layer 0; K of layers 1-3 = 3 taps x CH channel slots (17 used) + 4 pad
constexpr int N0 = 32, N3 = 64;
// weight images (floats): per layer big | small
constexpr int W0_OFF = 0, W1_OFF = W0_OFF + 2 * N0 * K0, W2_OFF = W1_OFF + 2 * N0 * K1, W3_OFF = W2_OFF + 2 * N0 * K1;
constexpr int W_FLOATS = W3_OFF + 2 * N3 * K1;                       // 16896 floats = 67584 B
constexpr int AX_FLOATS = 2 * 128 * K1;                              // [128 rows][64] big | small, reused by layers 1-3
constexpr int WG_FLOATS = AX_FLOATS;                                 // 16384 floats = 65536 B per warpgroup
constexpr int SMEM_BYTES = (W_FLOATS + 2 * WG_FLOATS) * 4 + 64;      // + mbarriers / TMEM slot  = 198720 B
constexpr int TM_COLS_WG = 256;                                      // D1 64 | D2 96 | D3 128 (64 wide)

// an activation row split once into TF32 big / small parts, as the five 16-byte k-chunks of one tap (17 channels + 3 zero slots)
struct SplitRow { float4 big[5], small[5]; };
__device__ __forceinline__ void split_row(const float* v, SplitRow& r) {
#pragma unroll
  for (int q = 0; q < 5; ++q) {
    const float x0 = v[4 * q], x1 = (4 * q + 1 < H) ? v[4 * q + 1] : 0.f, x2 = (4 * q + 2 < H) ? v[4 * q + 2] : 0.f,
                x3 = (4 * q + 3 < H) ? v[4 * q + 3] : 0.f;
    r.big[q] = make_float4(tf32_big(x0), tf32_big(x1), tf32_big(x2), tf32_big(x3));
    r.small[q] = make_float4(tf32_big(x0 - r.big[q].x), tf32_big(x1 - r.big[q].y), tf32_big(x2 - r.big[q].z), tf32_big(x3 - r.big[q].w));
  }
}
// write one (row, tap) of a layer-1..3 A operand; consecutive k-chunks of a row are 128 B (32 floats) apart in the canonical layout
__device__ __forceinline__ void put_tap(float* AX, int row, int tap, const SplitRow& r) {
  float* big = AX + canon_off(row, tap * CH, K1);
  float* small = big + 128 * K1;
#pragma unroll
  for (int q = 0; q < 5; ++q) {
    *reinterpret_cast<float4*>(big + 32 * q) = r.big[q];
    *reinterpret_cast<float4*>(small + 32 * q) = r.small[q];
  }
}
__device__ __forceinline__ void zero_tap(float* AX, int row, int tap) {
  float* big = AX + canon_off(row, tap * CH, K1);
  float* small = big + 128 * K1;
  const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
  for (int q = 0; q < 5; ++q) {
    *reinterpret_cast<float4*>(big + 32 * q) = z;
    *reinterpret_cast<float4*>(small + 32 * q) = z;
  }
}
// output position p (of LOUT) of frame-local row base rb: tap 1 of row p / 2 when even; tap 2 of row (p - 1) / 2 and tap 0 of
// row (p + 1) / 2 when odd; input position -1 of the frame's first row is zero padding (the buffer is reused across layers)
template <int LOUT>
__device__ __forceinline__ void scatter_next(float* AX, int rb, int p, const float* v) {
  SplitRow r;
  split_row(v, r);
  if ((p & 1) == 0) {
    put_tap(AX, rb + (p >> 1), 1, r);
    if (p == 0) zero_tap(AX, rb, 0);
  } else {
    put_tap(AX, rb + (p >> 1), 2, r);
    if ((p >> 1) + 1 < LOUT / 2) put_tap(AX, rb + (p >> 1) + 1, 0, r);
  }
}
// three MMAs per k-step: big*big + small*big + big*small.  The descriptors of a k-step differ from the first one's by a constant
// in the address field (two 128-byte core matrices = 256 B = 16 units per K = 8 step): built once, bumped by immediates.
template <int K>
__device__ __forceinline__ void mma_layer(uint32_t d_tmem, uint32_t a_big, uint32_t a_small, uint32_t b_big, uint32_t b_small, uint32_t idesc) {
  const uint64_t dab = make_desc(a_big, K), das = make_desc(a_small, K), dbb = make_desc(b_big, K), dbs = make_desc(b_small, K);
#pragma unroll
  for (int ks = 0; ks < K / 8; ++ks) {
    const uint64_t ko = (uint64_t)(ks * 16);
    if (ks == 0) mma_tf32_imm<false>(d_tmem, dab, dbb, idesc);
    else mma_tf32_imm<true>(d_tmem, dab + ko, dbb + ko, idesc);
    mma_tf32_imm<true>(d_tmem, das + ko, dbb + ko, idesc);
    mma_tf32_imm<true>(d_tmem, dab + ko, dbs + ko, idesc);
  }
}

__global__ void __launch_bounds__(256, 1) encoder_tc_kernel(const float* __restrict__ wimg, const float* __restrict__ blob,
                                                            const float* __restrict__ x, size_t nframes, float* __restrict__ d0,
                                                            float* __restrict__ d1, float* __restrict__ d2, float* __restrict__ gx) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  float* Wimg = reinterpret_cast<float*>(smem_raw);
  const int tid = threadIdx.x, wg = tid >> 7, tw = tid & 127, warp4 = tw >> 5, lane = tid & 31;
  float* AX = Wimg + W_FLOATS + wg * WG_FLOATS;
  uint64_t* bars = reinterpret_cast<uint64_t*>(Wimg + W_FLOATS + 2 * WG_FLOATS);  // [0] weights, [1 + wg] MMA completion
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 4);
  const Packed L = packed_layout();

  if (tid == 0) {
    tma::barrier_init(&bars[0], 1);
    tma::barrier_init(&bars[1], 1);
    tma::barrier_init(&bars[2], 1);
    tma::fence_barrier_init();
  }
  if ((tid >> 5) == 0) tmem_alloc(tmem_slot, 512);
  // zero both A regions once: K padding, unused channel slots and the rows a short layer does not fill must hold finite values
  for (int i = tw; i < WG_FLOATS / 4; i += 128) reinterpret_cast<float4*>(AX)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem = *tmem_slot + (uint32_t)(wg * TM_COLS_WG);
  if (tid == 0) {  // TMA: all four weight images in one bulk copy
    tma::expect_bytes(&bars[0], W_FLOATS * 4);
    tma::load(Wimg, wimg, W_FLOATS * 4, &bars[0]);
  }
  const uint32_t sW = tma::smem_addr(Wimg), sAX = tma::smem_addr(AX);
  uint64_t* mbar = &bars[1 + wg];
  uint32_t phase = 0;
  bool weights_ready = false;
  const uint32_t lane_addr = tmem + ((uint32_t)(warp4 * 32) << 16);
  const size_t ntiles = (nframes + FT - 1) / FT;
  // position biases (bias + folded Gaussian channels): a thread always serves the same position of layers 0-2, so its 3 x 17
  // values live in registers for the whole kernel (global loads inside the epilogues were the top stall: long_scoreboard 5.1 / issue)
  // layer 0 has one input channel and three taps: 51 FMAs per output row on the CUDA cores, exact fp32, straight into layer 1's
  // A operand -- cheaper than a tensor-core stage with its TMEM round trip (its weights: 3 x 17 registers)
  float pb0[H], pb1[H], pb2[H], w0[3][H];
#pragma unroll
  for (int c = 0; c < H; ++c) {
#pragma unroll
    for (int k = 0; k < 3; ++k) w0[k][c] = blob[L.enc_w[0] + k * HP + c];
    pb0[c] = blob[L.enc_pb[0] + lane * HP + c];
    pb1[c] = blob[L.enc_pb[1] + (tw & 15) * HP + c];
    pb2[c] = blob[L.enc_pb[2] + (tw & 7) * HP + c];
  }
  // taps of this thread's two layer-0 rows, fetched one tile ahead
  float nx[6];
  auto fetch_x = [&](size_t tl) {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int r = h * 128 + tw, fl = r >> 5, j = r & 31;
      const size_t f = tl * FT + fl;
      nx[3 * h] = nx[3 * h + 1] = nx[3 * h + 2] = 0.f;
      if (tl < ntiles && f < nframes) {
        const float* xf = x + f * NMEL;
        nx[3 * h + 1] = xf[2 * j];
        nx[3 * h + 2] = xf[2 * j + 1];
        if (j > 0) nx[3 * h] = xf[2 * j - 1];
      }
    }
  };
  fetch_x((size_t)blockIdx.x * 2 + wg);

  for (size_t tile = (size_t)blockIdx.x * 2 + wg; tile < ntiles; tile += (size_t)gridDim.x * 2) {
    const size_t f0 = tile * FT;
    // ---- layer 0 on the CUDA cores: row (frame h * 4 + warp, position lane) -> d0, -> A1 rows (frame * 16 + j1) -------------
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const float t0 = nx[3 * h], t1 = nx[3 * h + 1], t2 = nx[3 * h + 2];
      const int fl = h * 4 + warp4, p = lane;
      const size_t f = f0 + fl;
      float v[H];
#pragma unroll
      for (int c = 0; c < H; ++c) v[c] = fmaxf(fmaf(w0[2][c], t2, fmaf(w0[1][c], t1, fmaf(w0[0][c], t0, pb0[c]))), 0.f);
      if (f < nframes) {
#pragma unroll
        for (int c = 0; c < H; ++c) d0[f * D0 + c * 32 + p] = v[c];
      }
      scatter_next<32>(AX, fl * 16, p, v);
    }
    fetch_x(tile + (size_t)gridDim.x * 2);  // next tile's taps: in flight during this tile's three tensor-core layers
    fence_async_smem();
    named_barrier(1 + wg, 128);
    // ---- layer 1 -----------------------------------------------------------------------------------------------------
    if (tw == 0) {
      if (!weights_ready) tma::wait(&bars[0], 0);
      fence_after_sync();
      mma_layer<K1>(tmem + 64, sAX, sAX + 128 * K1 * 4, sW + W1_OFF * 4, sW + (W1_OFF + N0 * K1) * 4, idesc_tf32(N0));
      commit(mbar);
    }
    weights_ready = true;
    tma::wait(mbar, phase);
    phase ^= 1;
    fence_after_sync();
    {  // epilogue 1: row = tw = (frame tw / 16, position tw % 16) -> d1, -> A2 rows (frame * 8 + j2)
      float acc[32];
      ld16(lane_addr + 64, acc);
      ld16(lane_addr + 64 + 16, acc + 16);
      ld_wait();
      const int fl = tw >> 4, p = tw & 15;
      const size_t f = f0 + fl;
      float v[H];
#pragma unroll
      for (int c = 0; c < H; ++c) v[c] = fmaxf(acc[c] + pb1[c], 0.f);
      if (f < nframes) {
#pragma unroll
        for (int c = 0; c < H; ++c) d1[f * D1 + c * 16 + p] = v[c];
      }
      scatter_next<16>(AX, fl * 8, p, v);  // (the A buffer is free: MMA 1, its only reader, has completed)
    }
    fence_async_smem();
    fence_before_sync();
    named_barrier(1 + wg, 128);
    // ---- layer 2 (64 live rows of a 128-row tile) ------------------------------------------------------------------------
    if (tw == 0) {
      fence_after_sync();
      mma_layer<K1>(tmem + 96, sAX, sAX + 128 * K1 * 4, sW + W2_OFF * 4, sW + (W2_OFF + N0 * K1) * 4, idesc_tf32(N0));
      commit(mbar);
    }
    tma::wait(mbar, phase);
    phase ^= 1;
    fence_after_sync();
    {
      float acc[32];
      ld16(lane_addr + 96, acc);  // all four warps execute the (warp-collective) TMEM load; rows >= 64 are discarded
      ld16(lane_addr + 96 + 16, acc + 16);
      ld_wait();
      if (tw < 64) {
        const int fl = tw >> 3, p = tw & 7;
        const size_t f = f0 + fl;
        float v[H];
#pragma unroll
        for (int c = 0; c < H; ++c) v[c] = fmaxf(acc[c] + pb2[c], 0.f);
        if (f < nframes) {
#pragma unroll
          for (int c = 0; c < H; ++c) d2[f * D2 + c * 8 + p] = v[c];
        }
        scatter_next<8>(AX, fl * 4, p, v);
      }
    }
    fence_async_smem();
    fence_before_sync();
    named_barrier(1 + wg, 128);
    // ---- layer 3 (32 live rows, N = 64: the 51 gate channels) ----------------------------------------------------------------
    if (tw == 0) {
      fence_after_sync();
      mma_layer<K1>(tmem + 128, sAX, sAX + 128 * K1 * 4, sW + W3_OFF * 4, sW + (W3_OFF + N3 * K1) * 4, idesc_tf32(N3));
      commit(mbar);
    }
    float pb3[H3];  // warp 0 only: the gate biases, loaded while the MMAs run
    if (tw < 32) {
#pragma unroll
      for (int c = 0; c < H3; ++c) pb3[c] = blob[L.enc_pb[3] + (tw & 3) * H3P + c];
    }
    tma::wait(mbar, phase);
    phase ^= 1;
    fence_after_sync();
    {
      float acc[64];
#pragma unroll
      for (int c0 = 0; c0 < 64; c0 += 16) ld16(lane_addr + 128 + c0, acc + c0);
      ld_wait();
      if (tw < 32) {
        const int fl = tw >> 2, p = tw & 3;
        const size_t f = f0 + fl;
        if (f < nframes) {
#pragma unroll
          for (int c = 0; c < H3; ++c) gx[f * GX + c * 4 + p] = fmaxf(acc[c] + pb3[c], 0.f);
        }
      }
    }
    fence_before_sync();
    named_barrier(1 + wg, 128);  // TMEM and the A buffers of this warpgroup are free for its next tile
    fence_after_sync();
  }
  if (!weights_ready && tw == 0 && wg == 0) tma::wait(&bars[0], 0);  // never leave with a bulk copy in flight
  fence_before_sync();
  __syncthreads();
  if ((tid >> 5) == 0) tmem_dealloc(*tmem_slot, 512);
}

}  // namespace utc

// ---- host side ------------------------------------------------------------------------------------------
static void put_w(std::vector<float>& img, int base, int NP, int KP, int n, int k, float w) {
  const float big = tc5::tf32_big(w);
  img[base + tc5::canon_off(n, k, KP)] = big;
  img[base + NP * KP + tc5::canon_off(n, k, KP)] = tc5::tf32_big(w - big);
}

// weight images of the four encoder layers (Conv1d weights [co][ci + G][3], data channels only): K index = tap * 20 + ci
int model_pack_utc(b2d_model* m, const float* const* hp) {
  using namespace utc;
  const int G = m->cfg.num_gaussians;
  std::vector<float> img(W_FLOATS, 0.f);
  const int base[4] = {W0_OFF, W1_OFF, W2_OFF, W3_OFF};
  const int NP[4] = {N0, N0, N0, N3}, KP[4] = {K0, K1, K1, K1};
  for (int l = 0; l < 4; ++l) {
    const int cin = l == 0 ? 1 : H, cout = l == 3 ? H3 : H, CT = cin + G;
    const float* W = hp[2 * l];
    for (int co = 0; co < cout; ++co)
      for (int ci = 0; ci < cin; ++ci)
        for (int t = 0; t < 3; ++t) put_w(img, base[l], NP[l], KP[l], co, l == 0 ? t : t * CH + ci, W[(co * CT + ci) * 3 + t]);
  }
  B2D_CUDA(cudaMalloc(&m->d_utc, img.size() * sizeof(float)));
  B2D_CUDA(cudaMemcpy(m->d_utc, img.data(), img.size() * sizeof(float), cudaMemcpyHostToDevice));
  return B2D_OK;
}

int model_encode_utc(const b2d_model* m, const float* x, size_t nframes, float* d0, float* d1, float* d2, float* gx, int num_sms,
                     cudaStream_t st) {
  using namespace utc;
  B2D_REQUIRE(m->d_utc != nullptr, B2D_ERR_CUDA, "tcgen05 encoder weight images are missing");
  const size_t smem = SMEM_BYTES;
  static_assert(SMEM_BYTES <= 232448, "encoder_tc_kernel shared memory exceeds 227 KB");
  B2D_SMEM_OPT_IN(smem, encoder_tc_kernel);
  const size_t ntiles = (nframes + FT - 1) / FT;
  const size_t want = (ntiles + 1) / 2;
  const int grid = (int)(want < (size_t)num_sms ? want : (size_t)num_sms);
  encoder_tc_kernel<<<grid, 256, smem, st>>>(m->d_utc, m->d_blob, x, nframes, d0, d1, d2, gx);
  B2D_LAUNCH_CHECK("encoder_tc_kernel");
  return B2D_OK;
}

}  // namespace b2d
