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
B2D_HD void fwd1_store(int lane, float2* v, const LaneTw& t, float2* S) {
  bfly8x2<false>(v);
#pragma unroll
  for (int k1 = 0; k1 < 8; ++k1) {
    float2 a = v[2 * k1], b = v[2 * k1 + 1];
    if (k1) { a = cmul(a, t.a[k1]); b = cmul(b, t.b[k1]); }
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
B2D_HD void fwd2_store(int lane, float2* u, const LaneTw& t, float2* S) {
  const int kq = lane >> 3, n3 = lane & 7;
  bfly8x2<false>(u);
#pragma unroll
  for (int k2 = 0; k2 < 8; ++k2) {
    float2 a = u[2 * k2], b = u[2 * k2 + 1];
    if (k2) { a = cmul(a, t.c[k2]); b = cmul(b, t.c[k2]); }
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
B2D_HD void inv2_load(int lane, float2* u, const LaneTw& t, const float2* S) {
  const int kq = lane >> 3, n3 = lane & 7;
#pragma unroll
  for (int k2 = 0; k2 < 8; ++k2) {
    float2 a = S[n3 * LD2 + kq + 8 * k2], b = S[n3 * LD2 + kq + 4 + 8 * k2];
    if (k2) { a = cmulc(a, t.c[k2]); b = cmulc(b, t.c[k2]); }
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
B2D_HD void inv3_load(int lane, float2* v, const LaneTw& t, const float2* S) {
#pragma unroll
  for (int k1 = 0; k1 < 8; ++k1) {
    float2 a = S[k1 * LD1 + lane], b = S[k1 * LD1 + lane + 32];
    if (k1) { a = cmulc(a, t.a[k1]); b = cmulc(b, t.b[k1]); }
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
  const float ex = zk.x + zmk.x, ey = zk.y - zmk.y;
  const float dx = zk.x - zmk.x, dy = zk.y + zmk.y;
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
    ak = make_float2(fmaf(-mom, pk.x, xk.x), fmaf(-mom, pk.y, xk.y));
    amk = make_float2(fmaf(-mom, pmk.x, xmk.x), fmaf(-mom, pmk.y, xmk.y));
  }
  const float2 uk = unit_dir_fast(ak), umk = unit_dir_fast(amk);
  irfft_merge(make_float2(mk * uk.x, mk * uk.y), make_float2(mmk * umk.x, mmk * umk.y), rt, U, V);
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

}  // namespace fast512
}  // namespace b2d
