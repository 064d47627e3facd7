// gl_fast.cuh -- register-resident Griffin-Lim iteration for n_fft = 1024 (M = 512 = 8 x 8 x 8).
//
// One WARP owns one frame at a time: each lane keeps 16 complex values in registers through
// forward FFT -> spectral update -> inverse FFT; the only shared-memory traffic is four
// conflict-free 4 KB exchanges per frame (one per inter-pass transpose) plus read-only tables.
// A warp walks a contiguous run of frames, so the overlap-add carry stays in registers and no
// block-level barrier is needed after start-up.  tprev / mag go HBM <-> registers directly with
// fully coalesced 256-byte warp accesses.
//
// Index maps (n = 64 n1 + 8 n2 + n3 time index of z, k = k1 + 8 k2 + 64 k3 frequency index):
//   X[k1,k2,k3] = sum_n3 W8^{n3 k3} W64^{n3 k2} sum_n2 W8^{n2 k2} W512^{(8 n2+n3) k1} sum_n1 W8^{n1 k1} z[n1,n2,n3]
// lane p, register slot q:
//   stage 1: v[q]      = z[p + 32 q]                     butterflies i = p (even q) and i = p + 32 (odd q) over n1
//   stage 2: u[2 n2+b] = A[k1 = (p>>3) + 4 b ; 8 n2 + (p&7)]   butterflies over n2
//   stage 3: w[2 n3+f] = B[j_f ; n3], families j_0 = p, j_1 = 64 - p (lane 0: 0 and 32); butterflies over n3
// so that after stage 3 lane p holds Z[j_f + 64 k3]: bin k and its mirror M - k sit in the SAME lane,
// which makes the real-FFT split, the phase update and the merge lane-local.
// The phase functions are __host__ __device__ so tests/host/gl_fast_host_test.cu can emulate a warp on the CPU.
#pragma once

#include "fft.cuh"

namespace b2d {
namespace fast512 {

constexpr int M = 512, N = 1024, HOP = 512;
constexpr int LD1 = 72;          // row stride of the [k1][i] exchange layout (float2 units)
constexpr int LD2 = 66;          // row stride of the [n3][j] exchange layout
constexpr int XCH = 8 * LD1;     // float2 per warp exchange buffer (both layouts alias it)

struct LaneTw {
  float2 a[8];   // W512^{p k1}
  float2 b[8];   // W512^{(p+32) k1}
  float2 c[8];   // W64^{(p&7) k2}
  B2D_HD float2 A(int k) const { return a[k]; }
  B2D_HD float2 Bq(int k) const { return b[k]; }
  B2D_HD float2 Cq(int k) const { return c[k]; }
};
B2D_HD void lane_twiddles(int lane, const float2* __restrict__ tw512, LaneTw& t) {
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    t.a[k] = tw512[lane * k];
    t.b[k] = tw512[(lane + 32) * k];
    t.c[k] = tw512[8 * (lane & 7) * k];
  }
}

// two interleaved radix-8 butterflies on r[2 s + b], s = 0..7
template <bool INV>
B2D_HD void bfly8x2(float2* r) {
#pragma unroll
  for (int b = 0; b < 2; ++b) {
    float2 t[8];
#pragma unroll
    for (int s = 0; s < 8; ++s) t[s] = r[2 * s + b];
    dft8<INV>(t);
#pragma unroll
    for (int s = 0; s < 8; ++s) r[2 * s + b] = t[s];
  }
}

B2D_HD int fam(int lane, int f) { return f == 0 ? lane : (lane == 0 ? 32 : 64 - lane); }

// ---- forward ------------------------------------------------------------------------------------
template <typename TW>
B2D_HD void fwd1_store(int lane, float2* v, const TW& t, float2* S) {
  bfly8x2<false>(v);
#pragma unroll
  for (int k1 = 0; k1 < 8; ++k1) {
    float2 a = v[2 * k1], b = v[2 * k1 + 1];
    if (k1) { a = cmul(a, t.A(k1)); b = cmul(b, t.Bq(k1)); }
    S[k1 * LD1 + lane] = a;
    S[k1 * LD1 + lane + 32] = b;
  }
}
B2D_HD void fwd2_load(int lane, float2* u, const float2* S) {
  const int kq = lane >> 3, n3 = lane & 7;
#pragma unroll
  for (int n2 = 0; n2 < 8; ++n2) {
    u[2 * n2] = S[kq * LD1 + 8 * n2 + n3];
    u[2 * n2 + 1] = S[(kq + 4) * LD1 + 8 * n2 + n3];
  }
}
template <typename TW>
B2D_HD void fwd2_store(int lane, float2* u, const TW& t, float2* S) {
  const int kq = lane >> 3, n3 = lane & 7;
  bfly8x2<false>(u);
#pragma unroll
  for (int k2 = 0; k2 < 8; ++k2) {
    float2 a = u[2 * k2], b = u[2 * k2 + 1];
    if (k2) { const float2 c = t.Cq(k2); a = cmul(a, c); b = cmul(b, c); }
    S[n3 * LD2 + kq + 8 * k2] = a;
    S[n3 * LD2 + kq + 4 + 8 * k2] = b;
  }
}
B2D_HD void fwd3_load(int lane, float2* w, const float2* S) {
  const int j0 = fam(lane, 0), j1 = fam(lane, 1);
#pragma unroll
  for (int n3 = 0; n3 < 8; ++n3) {
    w[2 * n3] = S[n3 * LD2 + j0];
    w[2 * n3 + 1] = S[n3 * LD2 + j1];
  }
  bfly8x2<false>(w);
}

// ---- inverse (mirror image; twiddles applied after the loads) --------------------------------------
B2D_HD void inv1_store(int lane, float2* w, float2* S) {
  const int j0 = fam(lane, 0), j1 = fam(lane, 1);
  bfly8x2<true>(w);
#pragma unroll
  for (int n3 = 0; n3 < 8; ++n3) {
    S[n3 * LD2 + j0] = w[2 * n3];
    S[n3 * LD2 + j1] = w[2 * n3 + 1];
  }
}
template <typename TW>
B2D_HD void inv2_load(int lane, float2* u, const TW& t, const float2* S) {
  const int kq = lane >> 3, n3 = lane & 7;
#pragma unroll
  for (int k2 = 0; k2 < 8; ++k2) {
    float2 a = S[n3 * LD2 + kq + 8 * k2], b = S[n3 * LD2 + kq + 4 + 8 * k2];
    if (k2) { const float2 c = t.Cq(k2); a = cmulc(a, c); b = cmulc(b, c); }
    u[2 * k2] = a;
    u[2 * k2 + 1] = b;
  }
  bfly8x2<true>(u);
}
B2D_HD void inv2_store(int lane, const float2* u, float2* S) {
  const int kq = lane >> 3, n3 = lane & 7;
#pragma unroll
  for (int n2 = 0; n2 < 8; ++n2) {
    S[kq * LD1 + 8 * n2 + n3] = u[2 * n2];
    S[(kq + 4) * LD1 + 8 * n2 + n3] = u[2 * n2 + 1];
  }
}
template <typename TW>
B2D_HD void inv3_load(int lane, float2* v, const TW& t, const float2* S) {
#pragma unroll
  for (int k1 = 0; k1 < 8; ++k1) {
    float2 a = S[k1 * LD1 + lane], b = S[k1 * LD1 + lane + 32];
    if (k1) { a = cmulc(a, t.A(k1)); b = cmulc(b, t.Bq(k1)); }
    v[2 * k1] = a;
    v[2 * k1 + 1] = b;
  }
  bfly8x2<true>(v);
}

// ---- spectral step ---------------------------------------------------------------------------------
// lane 0 owns the self-paired families 0 and 32: re-arrange its registers so that slot r pairs
// w[2r] with w[2(7-r)+1] for every lane (A' = [A0..A3,B0..B3], B' = [B4..B7,A5,A6,A7,A4]).
B2D_HD void lane0_permute(float2* w) {
  float2 A[8], Bf[8];
#pragma unroll
  for (int r = 0; r < 8; ++r) { A[r] = w[2 * r]; Bf[r] = w[2 * r + 1]; }
  const float2 nA[8] = {A[0], A[1], A[2], A[3], Bf[0], Bf[1], Bf[2], Bf[3]};
  const float2 nB[8] = {Bf[4], Bf[5], Bf[6], Bf[7], A[5], A[6], A[7], A[4]};
#pragma unroll
  for (int r = 0; r < 8; ++r) { w[2 * r] = nA[r]; w[2 * r + 1] = nB[r]; }
}
B2D_HD void lane0_unpermute(float2* w) {
  float2 A[8], Bf[8];
#pragma unroll
  for (int r = 0; r < 8; ++r) { A[r] = w[2 * r]; Bf[r] = w[2 * r + 1]; }
  const float2 oA[8] = {A[0], A[1], A[2], A[3], Bf[7], Bf[4], Bf[5], Bf[6]};
  const float2 oB[8] = {A[4], A[5], A[6], A[7], Bf[0], Bf[1], Bf[2], Bf[3]};
#pragma unroll
  for (int r = 0; r < 8; ++r) { w[2 * r] = oA[r]; w[2 * r + 1] = oB[r]; }
}
// frequency index handled by slot r of a lane (the partner is M - k)
B2D_HD int slot_k(int lane, int r) { return (lane == 0 && r >= 4) ? 32 + 64 * (r - 4) : lane + 64 * r; }

// a / (|a| + 1e-16) without a branch: for |a| >= 1e-15 the epsilon is below fp32 resolution, for
// |a| -> 0 both forms tend to a * 1e16 (and give exactly 0 for a == 0, e.g. digital silence).
B2D_HD float inv_norm(float s) {
#ifdef __CUDA_ARCH__
  float r;  // the clamp keeps the argument a normal number: the flush-to-zero approximate unit (one MUFU) is exact enough
  asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(fmaxf(s, 1e-32f)));
  return r;
#else
  return 1.0f / sqrtf(fmaxf(s, 1e-32f));
#endif
}
B2D_HD float2 unit_dir_fast(float2 a) {
  const float inv = inv_norm(a.x * a.x + a.y * a.y);
  return make_float2(a.x * inv, a.y * inv);
}

// rfft_split without the factor 1/2: returns 2 X[k], 2 X[M-k].  The fast path keeps 2 * rebuilt in tprev: the phase
// update only uses the direction of (rebuilt - m * tprev), which is invariant to the common factor.
B2D_HD void rfft_split2(float2 zk, float2 zmk, float2 rt, float2& xk, float2& xmk) {
  const float2 e = cfma2(zmk, make_float2(1.0f, -1.0f), zk), d = cfma2(zmk, make_float2(-1.0f, 1.0f), zk);
  const float ex = e.x, ey = e.y, dx = d.x, dy = d.y;
  const float tx = fmaf(rt.x, dx, -rt.y * dy), ty = fmaf(rt.x, dy, rt.y * dx);
  xk = make_float2(ex + ty, ey - tx);
  xmk = make_float2(ex - ty, -ey - tx);
}

// One pair slot: split -> momentum/projection -> (store rebuilt) -> magnitude -> merge, in place.
// pk/pmk: previous rebuilt values at k and M-k (ignored unless use_prev); returns rebuilt values in xk/xmk.
B2D_HD void pair_update(float2& U, float2& V, float2 rt, float2 pk, float2 pmk, float mk, float mmk, float mom,
                        bool use_prev, float2& xk, float2& xmk) {
  rfft_split2(U, V, rt, xk, xmk);
  float2 ak = xk, amk = xmk;
  if (use_prev) {
    ak = cfma2(pk, make_float2(-mom, -mom), xk);
    amk = cfma2(pmk, make_float2(-mom, -mom), xmk);
  }
  const float2 uk = unit_dir_fast(ak), umk = unit_dir_fast(amk);
  irfft_merge(make_float2(mk * uk.x, mk * uk.y), make_float2(mmk * umk.x, mmk * umk.y), rt, U, V);
}

// ---- time-domain-momentum form (gl_fast.cu): the transform input already is x_k - m x_{k-1}, so a pair slot only needs
// split -> projection to unit modulus -> x mag -> merge.  EXACT: a / (|a| + 1e-16) with an IEEE square root and division,
// exactly as TA:functional/functional.py:343 writes it (plan flag B2D_PLAN_EXACT_UNIT), instead of the reciprocal-square-
// root unit.
template <bool EXACT>
B2D_HD float2 unit_dir_t(float2 a) {
  if (EXACT) {
    const float d = sqrtf(a.x * a.x + a.y * a.y) + 1e-16f;
    return make_float2(a.x / d, a.y / d);
  }
  return unit_dir_fast(a);
}
template <bool EXACT>
B2D_HD float unit_real_t(float a) {
  if (EXACT) return a / (fabsf(a) + 1e-16f);
  return a * inv_norm(a * a);
}
template <bool EXACT>
B2D_HD void pair_project(float2& U, float2& V, float2 rt, float mk, float mmk) {
  float2 xk, xmk;
  rfft_split2(U, V, rt, xk, xmk);
  if (EXACT) { xk = make_float2(0.5f * xk.x, 0.5f * xk.y); xmk = make_float2(0.5f * xmk.x, 0.5f * xmk.y); }  // the true X[k] next to the 1e-16
  const float2 uk = unit_dir_t<EXACT>(xk), umk = unit_dir_t<EXACT>(xmk);
  irfft_merge(make_float2(mk * uk.x, mk * uk.y), make_float2(mmk * umk.x, mmk * umk.y), rt, U, V);
}
// lane 0, slot 0: U = Z[0] (DC + Nyquist packed), V = Z[256] (self-paired bin M/2)
template <bool EXACT>
B2D_HD void special_project(float2& U, float2& V, float m0, float mM, float m256) {
  const float X0 = U.x + U.y, XM = U.x - U.y;  // X[0], X[M] (real); X[256] = conj V
  const float2 x256 = make_float2(V.x, -V.y);
  const float y0 = m0 * unit_real_t<EXACT>(X0), yM = mM * unit_real_t<EXACT>(XM);
  const float2 u = unit_dir_t<EXACT>(x256);
  U = make_float2(y0 + yM, y0 - yM);
  V = make_float2(2.0f * m256 * u.x, -2.0f * m256 * u.y);
}

// lane 0, slot 0: U = Z[0] (DC + Nyquist packed), V = Z[256] (self-paired bin M/2).
// p0 = (Re P[0], Re P[M]) packed, p256 = P[256]; m0, mM, m256 magnitudes.  Returns packed rebuilt values.
B2D_HD void special_update(float2& U, float2& V, float2 p0, float2 p256, float m0, float mM, float m256, float mom,
                           bool use_prev, float2& x0M, float2& x256) {
  const float X0 = 2.0f * (U.x + U.y), XM = 2.0f * (U.x - U.y);  // 2 * rebuilt, like rfft_split2
  x0M = make_float2(X0, XM);
  x256 = make_float2(2.0f * V.x, -2.0f * V.y);
  float a0 = X0, aM = XM;
  float2 a256 = x256;
  if (use_prev) {
    a0 = fmaf(-mom, p0.x, X0);
    aM = fmaf(-mom, p0.y, XM);
    a256 = make_float2(fmaf(-mom, p256.x, x256.x), fmaf(-mom, p256.y, x256.y));
  }
  const float y0 = m0 * (a0 * inv_norm(a0 * a0)), yM = mM * (aM * inv_norm(aM * aM));
  const float2 u = unit_dir_fast(a256);
  U = make_float2(y0 + yM, y0 - yM);
  V = make_float2(2.0f * m256 * u.x, -2.0f * m256 * u.y);
}

// ---- n_fft = 512: two consecutive real frames a, b ride one 512-point complex transform z = a + i b -----------------
// For a pair slot holding Z[lo] (lo <= 256) and Z[hi = 512 - lo]:  2 A[lo] = Z[lo] + conj Z[hi],  2 B[lo] = -i (Z[lo] - conj Z[hi]).
// Update both frames' bins (tprev holds 2 x rebuilt, as above) and rebuild Z'[lo] = Ya + i Yb, Z'[hi] = conj Ya + i conj Yb.
B2D_HD void pair2_update(float2& Zlo, float2& Zhi, float2 pa, float2 pb, float ma, float mb, float mom, bool use_prev,
                         float2& xa, float2& xb) {
  xa = make_float2(Zlo.x + Zhi.x, Zlo.y - Zhi.y);
  xb = make_float2(Zlo.y + Zhi.y, Zhi.x - Zlo.x);
  float2 aa = xa, ab = xb;
  if (use_prev) {
    aa = make_float2(fmaf(-mom, pa.x, xa.x), fmaf(-mom, pa.y, xa.y));
    ab = make_float2(fmaf(-mom, pb.x, xb.x), fmaf(-mom, pb.y, xb.y));
  }
  const float2 ua = unit_dir_fast(aa), ub = unit_dir_fast(ab);
  const float2 ya = make_float2(ma * ua.x, ma * ua.y), yb = make_float2(mb * ub.x, mb * ub.y);
  Zlo = make_float2(ya.x - yb.y, ya.y + yb.x);
  Zhi = make_float2(ya.x + yb.y, yb.x - ya.y);
}
// lane 0, slot 0: U = Z[0] = (DC of a, DC of b), V = Z[256] = (Nyquist of a, Nyquist of b), all real.
// p0a / p0b: packed (DC, Nyquist) of the previous rebuilt of frames a / b; returns the new packed values.
B2D_HD void special2_update(float2& U, float2& V, float2 p0a, float2 p0b, float ma0, float maN, float mb0, float mbN, float mom,
                            bool use_prev, float2& x0a, float2& x0b) {
  x0a = make_float2(2.0f * U.x, 2.0f * V.x);
  x0b = make_float2(2.0f * U.y, 2.0f * V.y);
  float a0 = x0a.x, aN = x0a.y, b0 = x0b.x, bN = x0b.y;
  if (use_prev) {
    a0 = fmaf(-mom, p0a.x, a0); aN = fmaf(-mom, p0a.y, aN);
    b0 = fmaf(-mom, p0b.x, b0); bN = fmaf(-mom, p0b.y, bN);
  }
  U = make_float2(ma0 * (a0 * inv_norm(a0 * a0)), mb0 * (b0 * inv_norm(b0 * b0)));
  V = make_float2(maN * (aN * inv_norm(aN * aN)), mbN * (bN * inv_norm(bN * bN)));
}

// ---- n_fft = 2048: the 1024-point complex transform is one radix-2 step over two 512-point register transforms -------
// z[n] = x[2n] + i x[2n+1] (n < 1024);  E = FFT512(z[2m]), O = FFT512(z[2m+1]);  Z[k] = E[k] + W1024^k O[k],
// Z[k+512] = E[k] - W1024^k O[k].  A pair slot holds bins k and 512-k of E and of O in the same lane, i.e. the four
// bins k, 512-k, 512+k, 1024-k of Z = the two real-FFT pairs (k, 1024-k) and (512-k, 512+k): still lane-local.
// tp / mg: staged rows (tprev holds 2 x rebuilt, bin 0 packed) ; st: the frame's tprev row in HBM or nullptr.
B2D_HD void store_prev2(float2* st, int k, float2 v) {
#ifdef __CUDA_ARCH__
  __stcs(st + k, v);
#else
  st[k] = v;
#endif
}
// inverse radix-2 step: E'[k] = Z[k] + Z[k+512], O'[k] = (Z[k] - Z[k+512]) conj W1024^k, and the same for bin 512-k
B2D_HD void quad_uncombine(float2 Zk, float2 Zkp, float2 Zm, float2 Zmp, float2 c, float2& Ek, float2& Emk, float2& Ok, float2& Omk) {
  Ek = make_float2(Zk.x + Zkp.x, Zk.y + Zkp.y);
  Ok = cmulc(make_float2(Zk.x - Zkp.x, Zk.y - Zkp.y), c);
  Emk = make_float2(Zm.x + Zmp.x, Zm.y + Zmp.y);
  Omk = cmul(make_float2(Zmp.x - Zm.x, Zmp.y - Zm.y), c);  // (Zm - Zmp) * conj(-conj c)
}
B2D_HD void quad_update(float2& Ek, float2& Emk, float2& Ok, float2& Omk, int k, float2 c, float2 rtk, float2 rtm,
                        const float2* tp, const float* mg, float mom, bool use_prev, float2* st) {
  const float2 zero = make_float2(0.f, 0.f);
  const float2 t = cmul(c, Ok);
  const float2 tm = cmul(make_float2(-c.x, c.y), Omk);  // W1024^(512-k) = -conj W1024^k
  float2 Zk = make_float2(Ek.x + t.x, Ek.y + t.y), Zkp = make_float2(Ek.x - t.x, Ek.y - t.y);          // bins k, k+512
  float2 Zm = make_float2(Emk.x + tm.x, Emk.y + tm.y), Zmp = make_float2(Emk.x - tm.x, Emk.y - tm.y);  // bins 512-k, 1024-k
  float2 x1, x2, x3, x4;
  pair_update(Zk, Zmp, rtk, use_prev ? tp[k] : zero, use_prev ? tp[1024 - k] : zero, mg[k], mg[1024 - k], mom, use_prev, x1, x2);
  pair_update(Zm, Zkp, rtm, use_prev ? tp[512 - k] : zero, use_prev ? tp[512 + k] : zero, mg[512 - k], mg[512 + k], mom, use_prev, x3, x4);
  if (st) { store_prev2(st, k, x1); store_prev2(st, 1024 - k, x2); store_prev2(st, 512 - k, x3); store_prev2(st, 512 + k, x4); }
  quad_uncombine(Zk, Zkp, Zm, Zmp, c, Ek, Emk, Ok, Omk);
}
// x_0 = istft(mag * angles_0): the same slot filled straight from the magnitudes (angle draws as in the generic kernel:
// element index = frame_base + bin, frame_base = (b T + t) * 1025; seed 0 = all-ones angles)
B2D_HD float2 init_bin(const float* mg, int k, unsigned long long seed, unsigned long long frame_base) {
  const float m = mg[k];
  if (!seed) return make_float2(m, 0.f);
  const float2 a = rand_angle(seed, frame_base + k);
  return make_float2(m * a.x, m * a.y);
}
B2D_HD void quad_init(float2& Ek, float2& Emk, float2& Ok, float2& Omk, int k, float2 c, float2 rtk, float2 rtm, const float* mg,
                      unsigned long long seed, unsigned long long frame_base) {
  float2 Zk, Zkp, Zm, Zmp;
  irfft_merge(init_bin(mg, k, seed, frame_base), init_bin(mg, 1024 - k, seed, frame_base), rtk, Zk, Zmp);
  irfft_merge(init_bin(mg, 512 - k, seed, frame_base), init_bin(mg, 512 + k, seed, frame_base), rtm, Zm, Zkp);
  quad_uncombine(Zk, Zkp, Zm, Zmp, c, Ek, Emk, Ok, Omk);
}
B2D_HD void quad_special_init(float2& E0, float2& E256, float2& O0, float2& O256, float2 rt256, const float* mg, unsigned long long seed,
                              unsigned long long frame_base) {
  const float y0 = init_bin(mg, 0, seed, frame_base).x, yM = init_bin(mg, 1024, seed, frame_base).x;
  const float2 y512 = init_bin(mg, 512, seed, frame_base);
  const float2 Z0 = make_float2(y0 + yM, y0 - yM), Z512 = make_float2(2.0f * y512.x, -2.0f * y512.y);
  float2 Z256, Z768;
  irfft_merge(init_bin(mg, 256, seed, frame_base), init_bin(mg, 768, seed, frame_base), rt256, Z256, Z768);
  E0 = make_float2(Z0.x + Z512.x, Z0.y + Z512.y);
  O0 = make_float2(Z0.x - Z512.x, Z0.y - Z512.y);
  E256 = make_float2(Z256.x + Z768.x, Z256.y + Z768.y);
  const float2 d = make_float2(Z256.x - Z768.x, Z256.y - Z768.y);
  O256 = make_float2(-d.y, d.x);
}
// lane 0, slot 0: E0 = E[0], E256 = E[256], O0 = O[0], O256 = O[256] -> bins 0 / 1024 (packed), 512 (self-paired), 256 & 768.
B2D_HD void quad_special(float2& E0, float2& E256, float2& O0, float2& O256, float2 rt256, const float2* tp, const float* mg, float mom,
                         bool use_prev, float2* st) {
  const float2 zero = make_float2(0.f, 0.f);
  float2 Z0 = make_float2(E0.x + O0.x, E0.y + O0.y), Z512 = make_float2(E0.x - O0.x, E0.y - O0.y);
  float2 Z256 = make_float2(E256.x + O256.y, E256.y - O256.x), Z768 = make_float2(E256.x - O256.y, E256.y + O256.x);  // W1024^256 = -i
  float2 x0M, x512, x256, x768;
  special_update(Z0, Z512, use_prev ? tp[0] : zero, use_prev ? tp[512] : zero, mg[0], mg[1024], mg[512], mom, use_prev, x0M, x512);
  pair_update(Z256, Z768, rt256, use_prev ? tp[256] : zero, use_prev ? tp[768] : zero, mg[256], mg[768], mom, use_prev, x256, x768);
  if (st) { store_prev2(st, 0, x0M); store_prev2(st, 512, x512); store_prev2(st, 256, x256); store_prev2(st, 768, x768); }
  E0 = make_float2(Z0.x + Z512.x, Z0.y + Z512.y);
  O0 = make_float2(Z0.x - Z512.x, Z0.y - Z512.y);
  E256 = make_float2(Z256.x + Z768.x, Z256.y + Z768.y);
  const float2 d = make_float2(Z256.x - Z768.x, Z256.y - Z768.y);
  O256 = make_float2(-d.y, d.x);  // * conj(-i) = * i
}

}  // namespace fast512
#ifdef __CUDACC__
// reflect-padded edge block of a clip (j == 0 or j == T): dst[i] = (x_k - mom x_{k-1})[reflected index] * inv_env * win
// (prev == nullptr: no momentum term; win_half == nullptr: no window).
// Sample i of the padded block is sample `is` of interior block jsA read in reverse (one sample comes from block jsB); the slot
// pointers are resolved once and the loads of 8 samples per lane are all in flight before the first use: a launch lasts as long
// as its slowest warp, and the two runs per clip that own an edge used to spend a serial chain of HOP / 32 L2 round trips here.
template <int HOP>
__device__ __forceinline__ void stage_reflect_wide(const float* part, const float* prev, float mom, const float* __restrict__ inv_env,
                                                  const float* __restrict__ win_half, int b, int R, int n, int T, int j,
                                                  float* __restrict__ dst, int lane) {
  const int jsA = (j == 0) ? 1 : T - 1, jsB = (j == 0) ? 2 : T - 2;
  auto slots = [&](const float* x, int js, const float*& p1, const float*& p2) {  // block js = slot of run (js-1)/n (+ slot 0 of run js/n)
    const int r1 = (js - 1) / n, r2 = js / n;
    p1 = x + ((size_t)(b * R + r1) * (n + 1) + (js - r1 * n)) * HOP;
    p2 = (r2 != r1) ? x + ((size_t)(b * R + r2) * (n + 1)) * HOP : nullptr;
  };
  const float *xa1, *xa2, *xb1, *xb2, *pa1 = nullptr, *pa2 = nullptr, *pb1 = nullptr, *pb2 = nullptr;
  slots(part, jsA, xa1, xa2);
  slots(part, jsB, xb1, xb2);
  if (prev) { slots(prev, jsA, pa1, pa2); slots(prev, jsB, pb1, pb2); }
  constexpr int PER = HOP / 32, CH = 8;
  static_assert(HOP % 32 == 0, "hop must be a multiple of the warp size");
#pragma unroll 1
  for (int q0 = 0; q0 < PER; q0 += CH) {
    float xr[CH], pr[CH], sc[CH];
#pragma unroll
    for (int u = 0; u < CH; ++u) {
      const int i = lane + 32 * (q0 + u);
      xr[u] = 0.f; pr[u] = 0.f; sc[u] = 0.f;
      if (q0 + u < PER) {
        const bool odd_one = (j == 0) ? (i == 0) : (i == HOP - 1);
        const int is = (j == 0) ? (i == 0 ? 0 : HOP - i) : (i == HOP - 1 ? HOP - 1 : HOP - 2 - i);
        const float* s1 = odd_one ? xb1 : xa1;
        const float* s2 = odd_one ? xb2 : xa2;
        float x = __ldcg(s1 + is);
        if (s2) x += __ldcg(s2 + is);
        xr[u] = x;
        if (prev) {
          const float* t1 = odd_one ? pb1 : pa1;
          const float* t2 = odd_one ? pb2 : pa2;
          float pv = __ldcg(t1 + is);
          if (t2) pv += __ldcg(t2 + is);
          pr[u] = pv;
        }
        sc[u] = inv_env[is];
      }
    }
#pragma unroll
    for (int u = 0; u < CH; ++u) {
      const int i = lane + 32 * (q0 + u);
      if (q0 + u < PER) {
        float x = xr[u];
        if (prev) x = fmaf(-mom, pr[u], x);
        dst[i] = win_half ? x * sc[u] * win_half[i] : x * sc[u];
      }
    }
  }
}

#endif

}  // namespace b2d
