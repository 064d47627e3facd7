// invmel_tc.cu -- the inverse-mel projection relu(pinv(fb^T) @ mel) (TA:transforms/_transforms.py:508; app3.py:210-211) as a
// tcgen05 / TMEM GEMM:  D[128 frames, 176] += A[128 frames, 64] * B[176, 64]^T  per CTA column (3 columns cover Fp = 516).
// Operands are fp32 values consumed as TF32 (kind::tf32, K = 8 per instruction) from shared memory in the canonical K-major
// no-swizzle core-matrix layout; every operand is split into big + small TF32 parts and three MMAs are issued per k-step
// (big*big + small*big + big*small): fp32-class results (~1e-6) from 11-bit tensor-core multiplies.  Accumulators live in
// TMEM and come back with tcgen05.ld (one TMEM lane = one frame per thread); the weight images (split and laid out at plan
// creation) arrive with one TMA bulk copy per CTA.
// (Round 1 also ran the U-Net convolutions layer by layer on this kernel shape -- 460 us encoder, 5.7 ms decoder; the fused
//  tcgen05 encoder that replaced that trial is unet_tc.cu.)
#include <math.h>
#include <string.h>
#include <vector>

#include "kernels.cuh"

namespace b2d {

// ---- PTX wrappers ---------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t s_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void tc_mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void tc_mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tc_bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(s_u32(dst)), "l"(src),
               "r"(bytes), "r"(s_u32(bar))
               : "memory");
}
__device__ __forceinline__ void tc_mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "W_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra D_%=;\n"
      "bra W_%=;\n"
      "D_%=:\n"
      "}\n" ::"r"(s_u32(bar)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s_u32(dst_smem)), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma_tf32(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(s_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void proxy_fence_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
// 16 consecutive fp32 columns of this thread's TMEM lane
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float* v) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

// the same without the wait, and a wait that names the destination registers of two such loads as in / out operands so that the
// compiler cannot move a use of them above it (the loads of the next group are in flight while a group is being stored)
__device__ __forceinline__ void tmem_ld16_async(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_wait_ld32(uint32_t* a, uint32_t* b) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(a[0]), "+r"(a[1]), "+r"(a[2]), "+r"(a[3]), "+r"(a[4]), "+r"(a[5]), "+r"(a[6]), "+r"(a[7]), "+r"(a[8]), "+r"(a[9]),
                 "+r"(a[10]), "+r"(a[11]), "+r"(a[12]), "+r"(a[13]), "+r"(a[14]), "+r"(a[15]), "+r"(b[0]), "+r"(b[1]), "+r"(b[2]),
                 "+r"(b[3]), "+r"(b[4]), "+r"(b[5]), "+r"(b[6]), "+r"(b[7]), "+r"(b[8]), "+r"(b[9]), "+r"(b[10]), "+r"(b[11]),
                 "+r"(b[12]), "+r"(b[13]), "+r"(b[14]), "+r"(b[15])
               :
               : "memory");
}

// ---- canonical K-major no-swizzle layout (TF32: 4 elements per 16-byte core-matrix row) ---------------
// element (row, k) of a [rows, KP] operand: ((row/8) * (KP/4) + k/4) * 128 + (row%8) * 16 + (k%4) * 4   [bytes]
__host__ __device__ inline int canon_off_f(int row, int k, int KP) { return (((row >> 3) * (KP >> 2) + (k >> 2)) << 5) + ((row & 7) << 2) + (k & 3); }  // in floats
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, int KP) {
  const uint64_t lbo = 128 >> 4, sbo = (uint64_t)((KP >> 2) * 128) >> 4;
  return (uint64_t)((saddr >> 4) & 0x3FFF) | (lbo << 16) | (sbo << 32) | (1ull << 46);  // version 1, SWIZZLE_NONE
}
__host__ __device__ inline float tf32_big(float v) {
#ifdef __CUDA_ARCH__
  return __uint_as_float(__float_as_uint(v) & 0xFFFFE000u);
#else
  uint32_t u; memcpy(&u, &v, 4); u &= 0xFFFFE000u; float r; memcpy(&r, &u, 4); return r;
#endif
}
__host__ __device__ constexpr uint32_t idesc_tf32(int N) { return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((128u >> 4) << 24); }

// ---- inverse mel: out[frame, f] = relu(sum_m mel[frame, m] * P[f, m]) -------------------------------------
// rows = frames, K = n_mels (64), N = 176 per CTA column (3 columns cover Fp = 516 padded to 528).
struct TcInvMel {
  const float* mel;    // [NF][K]
  const float* wimg;   // 3 x (big | small) images of [176, 64]
  float* out;          // [NF][Fp]
  size_t nframes;
  int Fp, terms;
};
constexpr int IM_N = 176, IM_K = 64;

// Two-stage software pipeline (round 2): the A operand (frames x 64 mel values, big / small TF32 images) and
// the TMEM accumulator are double-buffered; while the 128 threads read tile i out of TMEM, clamp it and store it, the MMAs of
// tile i + 1 are already running on the tensor pipe and its mel rows were gathered before the wait.  (Round 1: gather ->
// barrier -> MMA -> wait -> epilogue strictly in sequence, one tile at a time.)
__global__ void __launch_bounds__(128) invmel_tc_kernel(const TcInvMel L) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  float* A_big0 = reinterpret_cast<float*>(smem_raw);       // [2 buffers][big | small][128 x 64]
  float* B_big = A_big0 + 4 * 128 * IM_K;
  float* B_small = B_big + IM_N * IM_K;
  uint64_t* bar_w = reinterpret_cast<uint64_t*>(B_small + IM_N * IM_K);
  uint64_t* bar_mma = bar_w + 1;                             // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_mma + 2);
  const int tid = threadIdx.x, warp = tid >> 5;
  constexpr uint32_t TM_COLS = 512, TM_BUF = 256;            // two accumulators of 176 columns
  if (tid == 0) {
    tc_mbar_init(bar_w, 1);
    tc_mbar_init(bar_mma, 1);
    tc_mbar_init(bar_mma + 1, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) tmem_alloc(tmem_slot, TM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const int ncol = blockIdx.y;  // which 176-wide slice of the output
  if (tid == 0) {
    constexpr uint32_t wbytes = 2u * IM_N * IM_K * 4u;
    tc_mbar_expect_tx(bar_w, wbytes);
    tc_bulk_g2s(B_big, L.wimg + (size_t)ncol * 2 * IM_N * IM_K, wbytes, bar_w);
  }
  const size_t tiles = (L.nframes + 127) / 128;
  pdl_wait();  // TMEM allocated, weight image in flight; mel comes from the decoder (common.cuh: programmatic dependent launch)
  pdl_trigger();

  auto gather = [&](size_t tile, int buf) {  // this thread's frame of the tile: 64 mel values -> big / small TF32 images, canonical layout
    // (Measured: lanes walking the warp's contiguous 8 KB of mel rows in float4 order -- 4 cache lines per load instead of 32 -- makes
    //  the scatter into the core-matrix layout 16-way bank-conflicted: 64 us instead of 44 us for the kernel.  A row per thread
    //  puts 8 rows x 16 B side by side.)
    float* A_big = A_big0 + buf * 2 * 128 * IM_K;
    float* A_small = A_big + 128 * IM_K;
    const size_t frame = tile * 128 + tid;
    const bool live = frame < L.nframes;
    const float4* src = reinterpret_cast<const float4*>(L.mel + frame * IM_K);
    float4 vv[IM_K / 4];  // the frame's 256 bytes: every load in flight before the first use (ncu: this latency was 30 % of the kernel)
#pragma unroll
    for (int kc = 0; kc < IM_K / 4; ++kc) vv[kc] = live ? src[kc] : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int kc = 0; kc < IM_K / 4; ++kc) {
      const float4 v = vv[kc];
      const float4 big = make_float4(tf32_big(v.x), tf32_big(v.y), tf32_big(v.z), tf32_big(v.w));
      const float4 small = make_float4(tf32_big(v.x - big.x), tf32_big(v.y - big.y), tf32_big(v.z - big.z), tf32_big(v.w - big.w));
      const int off = canon_off_f(tid, kc * 4, IM_K);
      *reinterpret_cast<float4*>(A_big + off) = big;
      *reinterpret_cast<float4*>(A_small + off) = small;
    }
    proxy_fence_async();  // these generic-proxy stores -> visible to the tensor core's async-proxy reads after the next barrier
  };
  bool weights_ready = false;
  auto issue = [&](int buf) {  // thread 0: the 8 x (1 or 3) MMAs of one tile into accumulator `buf`
    if (!weights_ready) { tc_mbar_wait(bar_w, 0); weights_ready = true; }
    tc_fence_after();
    const uint32_t ab = s_u32(A_big0 + buf * 2 * 128 * IM_K), as = ab + 128 * IM_K * 4, bb = s_u32(B_big), bs = s_u32(B_small);
    const uint32_t acc = tmem + buf * TM_BUF;
    constexpr uint32_t idesc = idesc_tf32(IM_N);
#pragma unroll 1
    for (int ks = 0; ks < IM_K / 8; ++ks) {
      const uint32_t ko = ks * 256;
      umma_tf32(acc, make_desc(ab + ko, IM_K), make_desc(bb + ko, IM_K), idesc, ks > 0);
      if (L.terms == 3) {
        umma_tf32(acc, make_desc(as + ko, IM_K), make_desc(bb + ko, IM_K), idesc, 1);
        umma_tf32(acc, make_desc(ab + ko, IM_K), make_desc(bs + ko, IM_K), idesc, 1);
      }
    }
    umma_commit(bar_mma + buf);
  };

  size_t tile = blockIdx.x;
  uint32_t phase[2] = {0, 0};
  int buf = 0;
  if (tile < tiles) {
    gather(tile, 0);
    __syncthreads();
    if (tid == 0) issue(0);
  }
  for (; tile < tiles; tile += gridDim.x, buf ^= 1) {
    const size_t next = tile + gridDim.x;
    const bool has_next = next < tiles;
    // A[buf ^ 1] and accumulator buf ^ 1 are free: the MMAs that read / wrote them were awaited, and their epilogue ended
    // before the barrier that closed the previous trip
    if (has_next) gather(next, buf ^ 1);
    __syncthreads();
    if (has_next && tid == 0) issue(buf ^ 1);  // runs on the tensor pipe during this tile's epilogue
    tc_mbar_wait(bar_mma + buf, phase[buf]);
    phase[buf] ^= 1;
    tc_fence_after();
    // epilogue: TMEM lane = frame, a thread holds one frame's columns and stores them as float4s (64 contiguous bytes per load).
    // (Measured: staging the tile through shared memory so that consecutive lanes walk a frame's row -- 2 rows x 256 B per store
    //  instruction instead of 32 rows x 16 B -- is slower, 63 us instead of 44 us for the kernel: the extra pass costs more than the
    //  scattered stores.)
    const size_t frame = tile * 128 + tid;
    const bool live = frame < L.nframes;
    const uint32_t lane_addr = tmem + buf * TM_BUF + ((uint32_t)(warp * 32) << 16);
    float* orow = L.out + frame * L.Fp + (size_t)ncol * IM_N;
    auto store16 = [&](const uint32_t* r, int c0) {
      if (!live) return;
#pragma unroll
      for (int q = 0; q < 16; q += 4) {
        if (ncol * IM_N + c0 + q < L.Fp)  // Fp is a multiple of 4: whole float4 groups are in or out
          *reinterpret_cast<float4*>(orow + c0 + q) =
              make_float4(fmaxf(__uint_as_float(r[q]), 0.f), fmaxf(__uint_as_float(r[q + 1]), 0.f), fmaxf(__uint_as_float(r[q + 2]), 0.f),
                          fmaxf(__uint_as_float(r[q + 3]), 0.f));
      }
    };
    // 176 columns in groups of 32 (two 16-column loads), two register sets: the loads of group g + 1 fly while group g is stored
    static_assert(IM_N == 176, "epilogue schedule: 5 groups of 32 columns + one of 16");
    uint32_t ra0[16], ra1[16], rb0[16], rb1[16];
    tmem_ld16_async(lane_addr + 0, ra0);
    tmem_ld16_async(lane_addr + 16, ra1);
    tmem_wait_ld32(ra0, ra1);
#pragma unroll
    for (int c0 = 0; c0 < IM_N; c0 += 64) {
      if (c0 + 32 < IM_N) tmem_ld16_async(lane_addr + c0 + 32, rb0);
      if (c0 + 48 < IM_N) tmem_ld16_async(lane_addr + c0 + 48, rb1);
      store16(ra0, c0);
      if (c0 + 16 < IM_N) store16(ra1, c0 + 16);
      tmem_wait_ld32(rb0, rb1);
      if (c0 + 64 < IM_N) tmem_ld16_async(lane_addr + c0 + 64, ra0);
      if (c0 + 80 < IM_N) tmem_ld16_async(lane_addr + c0 + 80, ra1);
      if (c0 + 32 < IM_N) store16(rb0, c0 + 32);
      if (c0 + 48 < IM_N) store16(rb1, c0 + 48);
      tmem_wait_ld32(ra0, ra1);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
  }
  if (!weights_ready && tid == 0) tc_mbar_wait(bar_w, 0);
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, TM_COLS);
}

// ---- host side ------------------------------------------------------------------------------------------
static void put_image(std::vector<float>& img, size_t base, int NP, int KP, int n, int k, float w) {
  const float big = tf32_big(w);
  img[base + canon_off_f(n, k, KP)] = big;
  img[base + (size_t)NP * KP + canon_off_f(n, k, KP)] = tf32_big(w - big);
}

// plan: inverse-mel weight images (ceil(Fp / 176) column slices of [176, 64], big | small each; 3 slices at n_fft 1024, 5 at 1536)
int plan_pack_invmel_tc(b2d_plan* p, const float* h_pinv) {
  constexpr int kMaxSlices = 12;  // n_fft <= 4096: Fp = 2052 <= 12 x 176
  const int nslices = (p->Fp + IM_N - 1) / IM_N;
  if (p->n_mels != IM_K || nslices > kMaxSlices) { p->d_tw8 = nullptr; return B2D_OK; }  // 64 mel bins only
  std::vector<float> img((size_t)nslices * 2 * IM_N * IM_K, 0.f);
  for (int f = 0; f < p->F; ++f) {
    const int col = f / IM_N, n = f - col * IM_N;
    for (int k = 0; k < IM_K; ++k) put_image(img, (size_t)col * 2 * IM_N * IM_K, IM_N, IM_K, n, k, h_pinv[(size_t)f * IM_K + k]);
  }
  float* d = nullptr;
  B2D_CUDA(cudaMalloc(&d, img.size() * sizeof(float)));
  B2D_CUDA(cudaMemcpy(d, img.data(), img.size() * sizeof(float), cudaMemcpyHostToDevice));
  p->d_tw8 = reinterpret_cast<float2*>(d);
  return B2D_OK;
}

int launch_inverse_mel_tc(const b2d_plan* p, const float* mel_bt, size_t nframes, float* mag_tf, int terms, cudaStream_t st) {
  B2D_REQUIRE(p->d_tw8 != nullptr, B2D_ERR_UNSUPPORTED, "tensor-core inverse mel needs n_mels == 64");
  TcInvMel L;
  L.mel = mel_bt; L.wimg = reinterpret_cast<const float*>(p->d_tw8); L.out = mag_tf; L.nframes = nframes; L.Fp = p->Fp; L.terms = terms;
  const size_t smem = sizeof(float) * (size_t)(4 * 128 * IM_K + 2 * IM_N * IM_K) + 64;  // two A buffers (big | small) + the weight images
  B2D_SMEM_OPT_IN(smem, invmel_tc_kernel);
  const size_t tiles = (nframes + 127) / 128;
  const int ncols = (p->Fp + IM_N - 1) / IM_N;
  const size_t cap = (size_t)(p->num_sms / ncols > 0 ? p->num_sms / ncols : 1);
  dim3 grid((unsigned)(tiles < cap ? tiles : cap), ncols);
  B2D_CUDA(launch_pdl(invmel_tc_kernel, grid, dim3(128), smem, st, L));
  B2D_LAUNCH_CHECK("invmel_tc_kernel");
  return B2D_OK;
}

}  // namespace b2d
