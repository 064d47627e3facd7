// gl_reg.cuh -- register-resident FFT of length M = 64 * R3 for one warp (the n_fft = 1024 scheme of gl_fast.cuh made
// generic in its last radix): M = 8 x 8 x R3 with R3 in {4, 5, 8, 12}, i.e. n_fft = 512, 640, 1024, 1536 -- the 20 ms
// hop of BASELINE config 3 at 16 kHz (n_fft 640) and the reference app's own geometry (n_fft 1536 @ 48 kHz, app3.py:29-33).
//
// Index maps (time n = 8 R3 n1 + R3 n2 + n3, frequency k = k1 + 8 k2 + 64 k3):
//   X[k1,k2,k3] = sum_n3 W_R3^{n3 k3} W_M^{8 n3 k2} sum_n2 W_8^{n2 k2} W_M^{(R3 n2 + n3) k1} sum_n1 W_8^{n1 k1} z[n1,n2,n3]
//   stage 1: butterfly i = R3 n2 + n3 in [0, 8 R3): radix 8 over n1 (stride 8 R3), twiddle W_M^{i k1}
//   stage 2: butterfly (k1, n3), k1 < 8, n3 < R3: radix 8 over n2, twiddle W_M^{8 n3 k2}
//   stage 3: butterfly j = k1 + 8 k2 in [0, 64): radix R3 over n3; output bin k = j + 64 k3
// A lane runs butterflies lane + 32 r of stages 1 and 2 (r < NR = ceil(8 R3 / 32); the last round is partial when
// 8 R3 is not a multiple of 32) and the two stage-3 butterflies j = lane and j = 64 - lane (lane 0: j = 0 and j = 32),
// so that bin k and its mirror M - k always sit in the SAME lane: the real-FFT split, the projection and the merge are
// lane-local for every R3.  Between the stages the values cross lanes through the warp's shared-memory exchange buffer:
//   S1[k1][i]  row stride LD1 = 2 (mod 16)   (stage-1 stores: consecutive lanes -> consecutive i; stage-2 loads: k1 = u % 8,
//   S2[n3][j]  row stride LD2 = 8 (mod 16)    n3 = u / 8 -> banks 2 k1 + n3 / 8 n3 + k1: conflict-free 64-bit accesses)
// Twiddles: a lane keeps w, w^2, w^4 of its stage-1 / stage-2 twiddle bases per round in registers and forms the other
// powers with four complex multiplies per butterfly (shared-memory tables would load the LSU pipe that already bounds the
// iteration kernel; 7 powers per round would not fit the register file at R3 = 12).
// Everything is __host__ __device__ so tests/host/gl_reg_host_test.cu can emulate a warp on the CPU.
#pragma once

#include "gl_fast.cuh"

namespace b2d {
namespace regfft {

template <int R3>
struct Geo {
  static constexpr int M = 64 * R3, N = 2 * M, HOP = M;
  static constexpr int NB = 8 * R3;                              // butterflies of stages 1 and 2
  static constexpr int NR = (NB + 31) / 32;                      // rounds per lane
  static constexpr bool FULL = (NB % 32) == 0;
  static constexpr int LD1 = NB + ((2 - NB % 16) + 16) % 16;     // >= NB, = 2 (mod 16)
  static constexpr int LD2 = 72;                                 // >= 64, = 8 (mod 16)
  static constexpr int XCH = (8 * LD1 > R3 * LD2) ? 8 * LD1 : R3 * LD2;  // float2 per warp exchange buffer
  static constexpr int NV = NR * 8;                              // values per lane in stages 1 / 2
};

template <int R3>
struct LaneTwR {
  float2 t1[Geo<R3>::NR][3];  // W_M^{i}, W_M^{2i}, W_M^{4i}, i = lane + 32 r
  float2 t2[Geo<R3>::NR][3];  // W_M^{8 n3}, W_M^{16 n3}, W_M^{32 n3}, n3 = (lane >> 3) + 4 r
};

// tw[k] = W_M^k, k < M
template <int R3>
B2D_HD void lane_twiddles_r(int lane, const float2* __restrict__ tw, LaneTwR<R3>& t) {
  typedef Geo<R3> G;
#pragma unroll
  for (int r = 0; r < G::NR; ++r) {
    const int i = lane + 32 * r;
    const bool ok = G::FULL || i < G::NB;
    const int ii = ok ? i : 0, n3 = ok ? (lane >> 3) + 4 * r : 0;
    t.t1[r][0] = tw[ii]; t.t1[r][1] = tw[2 * ii]; t.t1[r][2] = tw[4 * ii];
    t.t2[r][0] = tw[8 * n3]; t.t2[r][1] = tw[16 * n3]; t.t2[r][2] = tw[32 * n3];
  }
}
// p[1..7] = w^1 .. w^7 from (w, w^2, w^4)
B2D_HD void tw_powers(const float2* b, float2* p) {
  p[1] = b[0]; p[2] = b[1]; p[4] = b[2];
  p[3] = cmul(b[0], b[1]); p[5] = cmul(b[0], b[2]); p[6] = cmul(b[1], b[2]); p[7] = cmul(p[3], b[2]);
}

// 12-point DFT as 3 x 4 by the prime-factor map (no twiddles): input n = (4 n1 + 3 n2) mod 12, output k = (4 k1 + 9 k2) mod 12
template <bool INV>
B2D_HD void dft12(float2* v) {
  float2 t[3][4];
#pragma unroll
  for (int n1 = 0; n1 < 3; ++n1) {
#pragma unroll
    for (int n2 = 0; n2 < 4; ++n2) t[n1][n2] = v[(4 * n1 + 3 * n2) % 12];
    dft4<INV>(t[n1]);
  }
#pragma unroll
  for (int k2 = 0; k2 < 4; ++k2) {
    float2 c[3] = {t[0][k2], t[1][k2], t[2][k2]};
    dft3<INV>(c);
#pragma unroll
    for (int k1 = 0; k1 < 3; ++k1) v[(4 * k1 + 9 * k2) % 12] = c[k1];
  }
}
template <int R3, bool INV>
B2D_HD void dft_r3(float2* v) {
  if (R3 == 4) dft4<INV>(v);
  if (R3 == 5) dft5<INV>(v);
  if (R3 == 8) dft8<INV>(v);
  if (R3 == 12) dft12<INV>(v);
}

// ---- forward -----------------------------------------------------------------------------------------
// v[8 r + n1] = z[(lane + 32 r) + 8 R3 n1]
template <int R3>
B2D_HD void fwd1_store_r(int lane, float2* v, const LaneTwR<R3>& t, float2* S) {
  typedef Geo<R3> G;
#pragma unroll
  for (int r = 0; r < G::NR; ++r) {
    const int i = lane + 32 * r;
    if (G::FULL || i < G::NB) {
      float2 p[8];
      tw_powers(t.t1[r], p);
      dft8<false>(v + 8 * r);
#pragma unroll
      for (int k1 = 0; k1 < 8; ++k1) S[k1 * G::LD1 + i] = k1 ? cmul(v[8 * r + k1], p[k1]) : v[8 * r];
    }
  }
}
template <int R3>
B2D_HD void fwd2_load_r(int lane, float2* u, const float2* S) {
  typedef Geo<R3> G;
  const int k1 = lane & 7;
#pragma unroll
  for (int r = 0; r < G::NR; ++r) {
    const int n3 = (lane >> 3) + 4 * r;
    if (G::FULL || lane + 32 * r < G::NB) {
#pragma unroll
      for (int n2 = 0; n2 < 8; ++n2) u[8 * r + n2] = S[k1 * G::LD1 + R3 * n2 + n3];
    }
  }
}
template <int R3>
B2D_HD void fwd2_store_r(int lane, float2* u, const LaneTwR<R3>& t, float2* S) {
  typedef Geo<R3> G;
  const int k1 = lane & 7;
#pragma unroll
  for (int r = 0; r < G::NR; ++r) {
    const int n3 = (lane >> 3) + 4 * r;
    if (G::FULL || lane + 32 * r < G::NB) {
      float2 p[8];
      tw_powers(t.t2[r], p);
      dft8<false>(u + 8 * r);
#pragma unroll
      for (int k2 = 0; k2 < 8; ++k2) S[n3 * G::LD2 + k1 + 8 * k2] = k2 ? cmul(u[8 * r + k2], p[k2]) : u[8 * r];
    }
  }
}
// wA[k3] = Z[jA + 64 k3], wB[k3] = Z[jB + 64 k3]
template <int R3>
B2D_HD void fwd3_load_r(int lane, float2* wA, float2* wB, const float2* S) {
  typedef Geo<R3> G;
  const int jA = fast512::fam(lane, 0), jB = fast512::fam(lane, 1);
#pragma unroll
  for (int n3 = 0; n3 < R3; ++n3) {
    wA[n3] = S[n3 * G::LD2 + jA];
    wB[n3] = S[n3 * G::LD2 + jB];
  }
  dft_r3<R3, false>(wA);
  dft_r3<R3, false>(wB);
}

// ---- inverse (mirror image; conjugate twiddles applied after the loads) ---------------------------------
template <int R3>
B2D_HD void inv1_store_r(int lane, float2* wA, float2* wB, float2* S) {
  typedef Geo<R3> G;
  const int jA = fast512::fam(lane, 0), jB = fast512::fam(lane, 1);
  dft_r3<R3, true>(wA);
  dft_r3<R3, true>(wB);
#pragma unroll
  for (int n3 = 0; n3 < R3; ++n3) {
    S[n3 * G::LD2 + jA] = wA[n3];
    S[n3 * G::LD2 + jB] = wB[n3];
  }
}
template <int R3>
B2D_HD void inv2_load_r(int lane, float2* u, const LaneTwR<R3>& t, const float2* S) {
  typedef Geo<R3> G;
  const int k1 = lane & 7;
#pragma unroll
  for (int r = 0; r < G::NR; ++r) {
    const int n3 = (lane >> 3) + 4 * r;
    if (G::FULL || lane + 32 * r < G::NB) {
      float2 p[8];
      tw_powers(t.t2[r], p);
#pragma unroll
      for (int k2 = 0; k2 < 8; ++k2) {
        const float2 a = S[n3 * G::LD2 + k1 + 8 * k2];
        u[8 * r + k2] = k2 ? cmulc(a, p[k2]) : a;
      }
      dft8<true>(u + 8 * r);
    }
  }
}
template <int R3>
B2D_HD void inv2_store_r(int lane, const float2* u, float2* S) {
  typedef Geo<R3> G;
  const int k1 = lane & 7;
#pragma unroll
  for (int r = 0; r < G::NR; ++r) {
    const int n3 = (lane >> 3) + 4 * r;
    if (G::FULL || lane + 32 * r < G::NB) {
#pragma unroll
      for (int n2 = 0; n2 < 8; ++n2) S[k1 * G::LD1 + R3 * n2 + n3] = u[8 * r + n2];
    }
  }
}
template <int R3>
B2D_HD void inv3_load_r(int lane, float2* v, const LaneTwR<R3>& t, const float2* S) {
  typedef Geo<R3> G;
#pragma unroll
  for (int r = 0; r < G::NR; ++r) {
    const int i = lane + 32 * r;
    if (G::FULL || i < G::NB) {
      float2 p[8];
      tw_powers(t.t1[r], p);
#pragma unroll
      for (int k1 = 0; k1 < 8; ++k1) {
        const float2 a = S[k1 * G::LD1 + i];
        v[8 * r + k1] = k1 ? cmulc(a, p[k1]) : a;
      }
      dft8<true>(v + 8 * r);
    }
  }
}

// ---- pairing --------------------------------------------------------------------------------------------------
// Lanes >= 1: slot r pairs U = wA[r] (bin lane + 64 r) with V = wB[R3 - 1 - r] (bin M - lane - 64 r).
// Lane 0 owns the self-mirrored families j = 0 (bins 64 a) and j = 32 (bins 32 + 64 b): slot 0 is the special pair
// (Z[0] = DC / Nyquist packed, Z[M/2] self-paired), slots 1 .. nA pair (64 a, M - 64 a), the rest pair (32 + 64 b, M - 32 - 64 b).
template <int R3>
struct Lane0Map {
  static constexpr int nA = (R3 % 2 == 0) ? R3 / 2 - 1 : (R3 - 1) / 2;
  // family (0 = A, 1 = B) and index of U / V of slot r
  __host__ __device__ static constexpr int u_fam(int r) { return (r <= nA) ? 0 : 1; }
  __host__ __device__ static constexpr int u_idx(int r) { return (r <= nA) ? r : r - nA - 1; }
  __host__ __device__ static constexpr int v_fam(int r) { return r == 0 ? ((R3 % 2 == 0) ? 0 : 1) : ((r <= nA) ? 0 : 1); }
  __host__ __device__ static constexpr int v_idx(int r) { return r == 0 ? ((R3 % 2 == 0) ? R3 / 2 : (R3 - 1) / 2) : ((r <= nA) ? R3 - r : R3 - 1 - (r - nA - 1)); }
  __host__ __device__ static constexpr int k(int r) { return (r <= nA) ? 64 * r : 32 + 64 * (r - nA - 1); }
};
template <int R3>
B2D_HD int slot_k_r(int lane, int r) { return lane ? lane + 64 * r : Lane0Map<R3>::k(r); }

template <int R3>
B2D_HD void gather_pairs(int lane, const float2* wA, const float2* wB, float2* U, float2* V) {
  typedef Lane0Map<R3> L0;
  if (lane == 0) {
#pragma unroll
    for (int r = 0; r < R3; ++r) {
      U[r] = L0::u_fam(r) ? wB[L0::u_idx(r)] : wA[L0::u_idx(r)];
      V[r] = L0::v_fam(r) ? wB[L0::v_idx(r)] : wA[L0::v_idx(r)];
    }
  } else {
#pragma unroll
    for (int r = 0; r < R3; ++r) { U[r] = wA[r]; V[r] = wB[R3 - 1 - r]; }
  }
}
template <int R3>
B2D_HD void scatter_pairs(int lane, float2* wA, float2* wB, const float2* U, const float2* V) {
  typedef Lane0Map<R3> L0;
  if (lane == 0) {
#pragma unroll
    for (int r = 0; r < R3; ++r) {
      if (L0::u_fam(r)) wB[L0::u_idx(r)] = U[r]; else wA[L0::u_idx(r)] = U[r];
      if (L0::v_fam(r)) wB[L0::v_idx(r)] = V[r]; else wA[L0::v_idx(r)] = V[r];
    }
  } else {
#pragma unroll
    for (int r = 0; r < R3; ++r) { wA[r] = U[r]; wB[R3 - 1 - r] = V[r]; }
  }
}

// projection of one frame's spectrum (TA functional.py:343 with the momentum already applied in the time domain):
// rt[k] = W_N^k (k < M), mg = the frame's magnitude row [M + 1]
template <int R3>
B2D_HD void project_frame(int lane, float2* wA, float2* wB, const float2* rt, const float* mg) {
  constexpr int M = Geo<R3>::M;
  float2 U[R3], V[R3];
  gather_pairs<R3>(lane, wA, wB, U, V);
#pragma unroll
  for (int r = 0; r < R3; ++r) {
    const int k = slot_k_r<R3>(lane, r);
    if (r == 0 && lane == 0) fast512::special_project<false>(U[0], V[0], mg[0], mg[M], mg[M / 2]);
    else fast512::pair_project<false>(U[r], V[r], rt[k], mg[k], mg[M - k]);
  }
  scatter_pairs<R3>(lane, wA, wB, U, V);
}

// x_0 = istft(mag * angles_0): the frame's packed spectrum straight from the magnitudes (element index of the draw = frame_base +
// bin with frame_base = (b T + t) * (M + 1), as in the generic kernel; seed 0 = all-ones angles)
template <int R3>
B2D_HD void init_frame(int lane, float2* wA, float2* wB, const float2* rt, const float* mg, unsigned long long seed,
                       unsigned long long frame_base) {
  constexpr int M = Geo<R3>::M;
  const float2 one = make_float2(1.f, 0.f);
  float2 U[R3], V[R3];
#pragma unroll
  for (int r = 0; r < R3; ++r) {
    const int k = slot_k_r<R3>(lane, r);
    if (r == 0 && lane == 0) {
      const float2 a0 = seed ? rand_angle(seed, frame_base) : one, aM = seed ? rand_angle(seed, frame_base + M) : one;
      const float2 ah = seed ? rand_angle(seed, frame_base + M / 2) : one;
      const float y0 = mg[0] * a0.x, yM = mg[M] * aM.x, mh = mg[M / 2];
      U[0] = make_float2(y0 + yM, y0 - yM);
      V[0] = make_float2(2.0f * mh * ah.x, -2.0f * mh * ah.y);
    } else {
      const float2 ak = seed ? rand_angle(seed, frame_base + k) : one, amk = seed ? rand_angle(seed, frame_base + (M - k)) : one;
      const float mk = mg[k], mmk = mg[M - k];
      irfft_merge(make_float2(mk * ak.x, mk * ak.y), make_float2(mmk * amk.x, mmk * amk.y), rt[k], U[r], V[r]);
    }
  }
  scatter_pairs<R3>(lane, wA, wB, U, V);
}

// the same from injected initial angles: ang = &angles0[b, 0, t] of the torch layout [B, F, T] (bin stride T)
template <int R3>
B2D_HD void init_frame_angles(int lane, float2* wA, float2* wB, const float2* rt, const float* mg, const float2* ang, int T) {
  constexpr int M = Geo<R3>::M;
  float2 U[R3], V[R3];
#pragma unroll
  for (int r = 0; r < R3; ++r) {
    const int k = slot_k_r<R3>(lane, r);
    if (r == 0 && lane == 0) {
      const float2 a0 = ang[0], aM = ang[(size_t)M * T], ah = ang[(size_t)(M / 2) * T];
      const float y0 = mg[0] * a0.x, yM = mg[M] * aM.x, mh = mg[M / 2];
      U[0] = make_float2(y0 + yM, y0 - yM);
      V[0] = make_float2(2.0f * mh * ah.x, -2.0f * mh * ah.y);
    } else {
      const float2 ak = ang[(size_t)k * T], amk = ang[(size_t)(M - k) * T];
      const float mk = mg[k], mmk = mg[M - k];
      irfft_merge(make_float2(mk * ak.x, mk * ak.y), make_float2(mmk * amk.x, mmk * amk.y), rt[k], U[r], V[r]);
    }
  }
  scatter_pairs<R3>(lane, wA, wB, U, V);
}

}  // namespace regfft
}  // namespace b2d
