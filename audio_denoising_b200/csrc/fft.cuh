// fft.cuh -- block-cooperative mixed-radix Stockham FFT in shared memory (any M = 2^a 3^b 5^c).
//
// This is the generic engine used by the STFT / iSTFT kernels for every supported n_fft and by
// the generic Griffin-Lim iteration.  The power-of-two fast path (gl_fast.cuh) keeps data in
// registers between passes instead.  Real transforms of length N run as complex transforms of
// length M = N/2 plus a split/merge step (rfft_split / irfft_merge below).
#pragma once

#include "common.cuh"

#define B2D_HD __host__ __device__ __forceinline__

namespace b2d {

// floor(n / d) for 0 <= n < 2^22 through the float unit: (n + 0.5) / d is never closer than 0.5 / d to an integer, far more
// than the rounding error of the product, so the truncation is exact.  The kernels' index math (w / Q, j / Ns, idx / hop ...)
// uses it instead of runtime integer division, which costs ~25 dependent instructions per quotient.
struct FastDiv {
  int d;
  float inv;
  B2D_HD FastDiv(int d_) : d(d_), inv(1.0f / (float)d_) {}
  B2D_HD int div(int n) const { return (int)(((float)n + 0.5f) * inv); }
};

B2D_HD float2 cmul(float2 a, float2 b) {
  return make_float2(fmaf(a.x, b.x, -a.y * b.y), fmaf(a.x, b.y, a.y * b.x));
}
B2D_HD float2 cmulc(float2 a, float2 b) {  // a * conj(b)
  return make_float2(fmaf(a.x, b.x, a.y * b.y), fmaf(a.y, b.x, -a.x * b.y));
}
// Complex add / subtract / scale.  On the device these are Blackwell's packed fp32 instructions (add.f32x2 / fma.rn.f32x2:
// one issue slot for both components, SASS FADD2 / FFMA2); a - b is fma(b, -1, a), which rounds exactly like the
// subtraction, so host and device stay bit-identical.
#if defined(__CUDA_ARCH__) && __CUDA_ARCH__ >= 1000
B2D_HD float2 cadd(float2 a, float2 b) { return __fadd2_rn(a, b); }
B2D_HD float2 csub(float2 a, float2 b) { return __ffma2_rn(b, make_float2(-1.0f, -1.0f), a); }
B2D_HD float2 cscale2(float2 a, float2 s) { return __fmul2_rn(a, s); }
B2D_HD float2 cfma2(float2 a, float2 b, float2 c) { return __ffma2_rn(a, b, c); }
#else
B2D_HD float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
B2D_HD float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
B2D_HD float2 cscale2(float2 a, float2 s) { return make_float2(a.x * s.x, a.y * s.y); }
B2D_HD float2 cfma2(float2 a, float2 b, float2 c) { return make_float2(fmaf(a.x, b.x, c.x), fmaf(a.y, b.y, c.y)); }
#endif
// rot90(a - b) computed directly into the rotated pair (two scalar subtractions; a packed subtraction would need its
// components swapped afterwards)
template <bool INV>
B2D_HD float2 csub_rot90(float2 a, float2 b) {
  return INV ? make_float2(b.y - a.y, a.x - b.x) : make_float2(a.y - b.y, b.x - a.x);
}
// multiply by -i (forward) or +i (inverse)
template <bool INV>
B2D_HD float2 rot90(float2 a) {
  return INV ? make_float2(-a.y, a.x) : make_float2(a.y, -a.x);
}

template <bool INV>
B2D_HD void dft2(float2& a, float2& b) {
  float2 t = a;
  a = cadd(t, b);
  b = csub(t, b);
}

template <bool INV>
B2D_HD void dft4(float2* v) {
  float2 a0 = cadd(v[0], v[2]), a1 = csub(v[0], v[2]);
  float2 a2 = cadd(v[1], v[3]), a3 = csub_rot90<INV>(v[1], v[3]);
  v[0] = cadd(a0, a2);
  v[2] = csub(a0, a2);
  v[1] = cadd(a1, a3);
  v[3] = csub(a1, a3);
}

template <bool INV>
B2D_HD void dft8(float2* v) {
  const float h = 0.70710678118654752440f;
  // three radix-2 stages, decimation in frequency, output written back in natural order
  float2 a[8];
  a[0] = cadd(v[0], v[4]);
  a[4] = csub(v[0], v[4]);
  a[1] = cadd(v[1], v[5]);
  a[2] = cadd(v[2], v[6]);
  a[3] = cadd(v[3], v[7]);
  // twiddles W8^i on the lower half, folded into the subtractions
  {
    const float2 t = csub(v[1], v[5]);  // * W8^1: fwd ((x + y) h, (y - x) h), inv ((x - y) h, (x + y) h)
    a[5] = cscale2(INV ? make_float2(t.x - t.y, t.x + t.y) : make_float2(t.x + t.y, t.y - t.x), make_float2(h, h));
    a[6] = csub_rot90<INV>(v[2], v[6]);
    const float2 u = csub(v[3], v[7]);  // * W8^3: fwd ((y - x) h, -(x + y) h), inv (-(x + y) h, (x - y) h)
    a[7] = INV ? cscale2(make_float2(u.x + u.y, u.x - u.y), make_float2(-h, h)) : cscale2(make_float2(u.y - u.x, u.x + u.y), make_float2(h, -h));
  }
  float2 b[8];
#pragma unroll
  for (int g = 0; g < 2; ++g) {
    float2* p = a + 4 * g;
    float2* q = b + 4 * g;
    q[0] = cadd(p[0], p[2]);
    q[2] = csub(p[0], p[2]);
    q[1] = cadd(p[1], p[3]);
    q[3] = csub_rot90<INV>(p[1], p[3]);
  }
  // last stage; outputs X[k]: k = 4*k2 + 2*k1 + k0 from DIF order
  v[0] = cadd(b[0], b[1]);
  v[4] = csub(b[0], b[1]);
  v[2] = cadd(b[2], b[3]);
  v[6] = csub(b[2], b[3]);
  v[1] = cadd(b[4], b[5]);
  v[5] = csub(b[4], b[5]);
  v[3] = cadd(b[6], b[7]);
  v[7] = csub(b[6], b[7]);
}

template <bool INV>
B2D_HD void dft3(float2* v) {
  const float s = INV ? 0.86602540378443864676f : -0.86602540378443864676f;  // sin(2pi/3) with sign
  float2 t1 = cadd(v[1], v[2]);
  float2 t2 = make_float2(v[0].x - 0.5f * t1.x, v[0].y - 0.5f * t1.y);
  float2 d = csub(v[1], v[2]);
  float2 t3 = make_float2(-s * d.y, s * d.x);  // i*s*d
  v[0] = cadd(v[0], t1);
  v[1] = cadd(t2, t3);
  v[2] = csub(t2, t3);
}

template <bool INV>
B2D_HD void dft5(float2* v) {
  const float c1 = 0.30901699437494742410f;   // cos(2pi/5)
  const float c2 = -0.80901699437494742410f;  // cos(4pi/5)
  const float s1 = INV ? 0.95105651629515357212f : -0.95105651629515357212f;  // sin(2pi/5), signed
  const float s2 = INV ? 0.58778525229247312917f : -0.58778525229247312917f;  // sin(4pi/5), signed
  float2 a1 = cadd(v[1], v[4]), b1 = csub(v[1], v[4]);
  float2 a2 = cadd(v[2], v[3]), b2 = csub(v[2], v[3]);
  float2 x0 = v[0];
  v[0] = make_float2(x0.x + a1.x + a2.x, x0.y + a1.y + a2.y);
  float2 p1 = make_float2(x0.x + c1 * a1.x + c2 * a2.x, x0.y + c1 * a1.y + c2 * a2.y);
  float2 p2 = make_float2(x0.x + c2 * a1.x + c1 * a2.x, x0.y + c2 * a1.y + c1 * a2.y);
  // q = i * (s1*b1 + s2*b2), r = i * (s2*b1 - s1*b2)
  float2 q = make_float2(-(s1 * b1.y + s2 * b2.y), s1 * b1.x + s2 * b2.x);
  float2 r = make_float2(-(s2 * b1.y - s1 * b2.y), s2 * b1.x - s1 * b2.x);
  v[1] = cadd(p1, q);
  v[4] = csub(p1, q);
  v[2] = cadd(p2, r);
  v[3] = csub(p2, r);
}

// Counter-based uniform draw for the Griffin-Lim initial angles (rand_init=True, TA:functional/functional.py:310
// draws real and imaginary parts ~ U[0,1)): two 24-bit mantissas from (seed, element index) through 32-bit multiply-xorshift
// rounds (the `lowbias32` constants).  (Round 1 used splitmix64: its three 64-bit multiplies cost ~45 integer instructions per
// draw on a GPU without a 64-bit multiplier -- as much as the inverse FFT of the init kernels beside it.)
B2D_HD float2 rand_angle(unsigned long long seed, unsigned long long idx) {
  unsigned a = (unsigned)idx * 0x9E3779B9u + (unsigned)seed;
  unsigned b = ((unsigned)(idx >> 32) + 0x7F4A7C15u) * 0x85EBCA6Bu + (unsigned)(seed >> 32);
  a ^= b;
  a ^= a >> 16; a *= 0x21F0AAADu; a ^= a >> 15; a *= 0x735A2D97u; a ^= a >> 15;
  b += a * 0x9E3779B9u;
  b ^= b >> 16; b *= 0x21F0AAADu; b ^= b >> 15; b *= 0x735A2D97u; b ^= b >> 15;
  const float k = 1.0f / 16777216.0f;
  return make_float2((float)(a >> 8) * k, (float)(b >> 8) * k);
}

// One radix-R Stockham butterfly: work item w in [0, rows * M/R) of a pass over `rows` independent
// length-M rows (row stride ld, in float2).  tw[k] = exp(-2 pi i k / M) (forward table; conjugated
// on the fly for INV).  Host-callable so the index math is unit-tested on the CPU.
template <bool INV, int R>
B2D_HD void stockham_item(const float2* __restrict__ src, float2* __restrict__ dst, int w, int M, int ld, int Ns,
                          const float2* __restrict__ tw) {
  const int Q = M / R;
  const FastDiv dq(Q), dn(Ns);
  const int tstep = dn.div(Q);  // M / (Ns * R)
  const int row = dq.div(w);
  const int j = w - row * Q;
  const int jq = dn.div(j);
  const int k = j - jq * Ns;
  const float2* s = src + row * ld;
  float2* d = dst + row * ld;
  float2 v[R];
#pragma unroll
  for (int r = 0; r < R; ++r) v[r] = s[j + r * Q];
  if (Ns > 1) {
#pragma unroll
    for (int r = 1; r < R; ++r) {
      float2 t = tw[r * k * tstep];
      v[r] = INV ? cmulc(v[r], t) : cmul(v[r], t);
    }
  }
  if (R == 2) dft2<INV>(v[0], v[1]);
  if (R == 3) dft3<INV>(v);
  if (R == 4) dft4<INV>(v);
  if (R == 5) dft5<INV>(v);
  if (R == 8) dft8<INV>(v);
  const int j0 = jq * Ns * R + k;
#pragma unroll
  for (int q = 0; q < R; ++q) d[j0 + q * Ns] = v[q];
}

template <bool INV, int R>
__device__ __forceinline__ void stockham_pass(const float2* __restrict__ src, float2* __restrict__ dst, int rows,
                                              int M, int ld, int Ns, const float2* __restrict__ tw) {
  const int n = rows * (M / R);
  for (int w = threadIdx.x; w < n; w += blockDim.x) stockham_item<INV, R>(src, dst, w, M, ld, Ns, tw);
}

// Full transform of `rows` rows.  Data starts in `a`; returns the buffer holding the result
// (a or b).  Contains __syncthreads(): must be called by all threads of the block; the caller
// must have synchronised after filling `a`.  Unnormalised in both directions.
template <bool INV>
__device__ __forceinline__ float2* fft_rows(float2* a, float2* b, int rows, int ld, const FftDesc& fd,
                                            const float2* __restrict__ tw) {
  int Ns = 1;
  for (int p = 0; p < fd.npass; ++p) {
    const int R = fd.radix[p];
    switch (R) {
      case 8: stockham_pass<INV, 8>(a, b, rows, fd.M, ld, Ns, tw); break;
      case 4: stockham_pass<INV, 4>(a, b, rows, fd.M, ld, Ns, tw); break;
      case 2: stockham_pass<INV, 2>(a, b, rows, fd.M, ld, Ns, tw); break;
      case 3: stockham_pass<INV, 3>(a, b, rows, fd.M, ld, Ns, tw); break;
      default: stockham_pass<INV, 5>(a, b, rows, fd.M, ld, Ns, tw); break;
    }
    __syncthreads();
    float2* t = a;
    a = b;
    b = t;
    Ns *= R;
  }
  return a;
}

// The same with the length (and hence the radix sequence of api.cu's factor(): 8s, 4s, 2s, 3s, 5s) known at compile time,
// for the two streaming geometries (n_fft 640 -> M = 320 = 8 8 5, n_fft 1536 -> M = 768 = 8 8 4 3): every index
// computation folds to constants.  MT == 0 falls back to the run-time description.
template <bool INV, int MT>
__device__ __forceinline__ float2* fft_rows_t(float2* a, float2* b, int rows, int ld, const FftDesc& fd,
                                              const float2* __restrict__ tw) {
  if (MT == 320) {
    stockham_pass<INV, 8>(a, b, rows, 320, ld, 1, tw);
    __syncthreads();
    stockham_pass<INV, 8>(b, a, rows, 320, ld, 8, tw);
    __syncthreads();
    stockham_pass<INV, 5>(a, b, rows, 320, ld, 64, tw);
    __syncthreads();
    return b;
  }
  if (MT == 512) {
    stockham_pass<INV, 8>(a, b, rows, 512, ld, 1, tw);
    __syncthreads();
    stockham_pass<INV, 8>(b, a, rows, 512, ld, 8, tw);
    __syncthreads();
    stockham_pass<INV, 8>(a, b, rows, 512, ld, 64, tw);
    __syncthreads();
    return b;
  }
  if (MT == 768) {
    stockham_pass<INV, 8>(a, b, rows, 768, ld, 1, tw);
    __syncthreads();
    stockham_pass<INV, 8>(b, a, rows, 768, ld, 8, tw);
    __syncthreads();
    stockham_pass<INV, 4>(a, b, rows, 768, ld, 64, tw);
    __syncthreads();
    stockham_pass<INV, 3>(b, a, rows, 768, ld, 256, tw);
    __syncthreads();
    return a;
  }
  return fft_rows<INV>(a, b, rows, ld, fd, tw);
}

// Real-FFT split: Z = FFT_M(x[2m] + i x[2m+1]) -> X[k], X[M-k] for one pair index k in [0, M/2].
// rt = W_N^k.  (k == 0 yields X[0] and X[M], both real; k == M/2 yields the same bin twice.)
B2D_HD void rfft_split(float2 zk, float2 zmk, float2 rt, float2& xk, float2& xmk) {
  // E = (Zk + conj(Zmk))/2, D = (Zk - conj(Zmk))/2 ; X[k] = E - i W^k D ; X[M-k] = conj(E) - i conj(W^k) conj(D) ... derived below
  const float ex = 0.5f * (zk.x + zmk.x), ey = 0.5f * (zk.y - zmk.y);
  const float dx = 0.5f * (zk.x - zmk.x), dy = 0.5f * (zk.y + zmk.y);
  // t = W^k * D
  const float tx = fmaf(rt.x, dx, -rt.y * dy), ty = fmaf(rt.x, dy, rt.y * dx);
  // X[k] = E - i t = (ex + ty, ey - tx)
  xk = make_float2(ex + ty, ey - tx);
  // X[M-k] = conj(E) + i conj(t)... : (ex - ty, -ey - tx)
  xmk = make_float2(ex - ty, -ey - tx);
}

// Inverse of the above (unnormalised: IFFT_M of the result gives N * x packed as re/im pairs / ... see
// irfft scaling note): builds Z'[k], Z'[M-k] from Y[k], Y[M-k].  rt = W_N^k (forward twiddle).
// Z'[k] = (Yk + conj(Ymk)) + i conj(W^k) (Yk - conj(Ymk)).
B2D_HD void irfft_merge(float2 yk, float2 ymk, float2 rt, float2& zk, float2& zmk) {
  const float2 e = cfma2(ymk, make_float2(1.0f, -1.0f), yk), d = cfma2(ymk, make_float2(-1.0f, 1.0f), yk);
  const float ex = e.x, ey = e.y;   // Yk + conj(Ymk)
  const float dx = d.x, dy = d.y;   // Yk - conj(Ymk)
  // t = conj(W^k) * D
  const float tx = fmaf(rt.x, dx, rt.y * dy), ty = fmaf(rt.x, dy, -rt.y * dx);
  // Z'[k] = E + i t = (ex - ty, ey + tx)
  zk = make_float2(ex - ty, ey + tx);
  // Z'[M-k] = conj(E) + i conj(t)... : (ex + ty, -ey + tx)
  zmk = make_float2(ex + ty, tx - ey);
}

}  // namespace b2d
