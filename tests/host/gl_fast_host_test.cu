// CPU emulation of one warp of the n_fft=1024 fast Griffin-Lim path (csrc/gl_fast.cuh): checks the
// three-pass register/shared-memory data flow against a naive DFT, and the lane-local spectral update
// against a straightforward double-precision evaluation of one Griffin-Lim iteration on one frame.
#include <cmath>
#include <complex>
#include <cstdio>
#include <vector>
#include "../../audio_denoising_b200/csrc/gl_fast.cuh"
using namespace b2d;
using namespace b2d::fast512;
typedef std::complex<double> cd;

static float frand(unsigned& s) { s = s * 1664525u + 1013904223u; return (s >> 8) / 16777216.0f - 0.5f; }

int main() {
  int bad = 0;
  std::vector<float2> tw(M), rt(M);
  for (int k = 0; k < M; ++k) {
    tw[k] = make_float2((float)cos(-2 * M_PI * k / M), (float)sin(-2 * M_PI * k / M));
    rt[k] = make_float2((float)cos(-2 * M_PI * k / N), (float)sin(-2 * M_PI * k / N));
  }
  LaneTw lt[32];
  for (int l = 0; l < 32; ++l) lane_twiddles(l, tw.data(), lt[l]);
  unsigned seed = 7;
  std::vector<float> x(N);
  for (auto& v : x) v = frand(seed);
  std::vector<float2> z(M);
  for (int m = 0; m < M; ++m) z[m] = make_float2(x[2 * m], x[2 * m + 1]);
  std::vector<cd> Z(M);
  for (int k = 0; k < M; ++k) { cd a = 0; for (int m = 0; m < M; ++m) a += cd(z[m].x, z[m].y) * std::polar(1.0, -2 * M_PI * (double)((long)m * k % M) / M); Z[k] = a; }

  float2 v[32][16];
  std::vector<float2> S(XCH);
  auto forward = [&]() {
    for (int l = 0; l < 32; ++l) fwd1_store(l, v[l], lt[l], S.data());
    for (int l = 0; l < 32; ++l) fwd2_load(l, v[l], S.data());
    for (int l = 0; l < 32; ++l) fwd2_store(l, v[l], lt[l], S.data());
    for (int l = 0; l < 32; ++l) fwd3_load(l, v[l], S.data());
  };
  auto inverse = [&]() {
    for (int l = 0; l < 32; ++l) inv1_store(l, v[l], S.data());
    for (int l = 0; l < 32; ++l) inv2_load(l, v[l], lt[l], S.data());
    for (int l = 0; l < 32; ++l) inv2_store(l, v[l], S.data());
    for (int l = 0; l < 32; ++l) inv3_load(l, v[l], lt[l], S.data());
  };
  for (int l = 0; l < 32; ++l) for (int q = 0; q < 16; ++q) v[l][q] = z[l + 32 * q];
  forward();
  double e = 0, nrm = 0;
  std::vector<int> seen(M, 0);
  for (int l = 0; l < 32; ++l) for (int f = 0; f < 2; ++f) for (int k3 = 0; k3 < 8; ++k3) {
    int k = fam(l, f) + 64 * k3; seen[k]++;
    e += std::norm(cd(v[l][2 * k3 + f].x, v[l][2 * k3 + f].y) - Z[k]); nrm += std::norm(Z[k]);
  }
  for (int k = 0; k < M; ++k) if (seen[k] != 1) { printf("bin %d covered %d times\n", k, seen[k]); bad++; }
  printf("forward rel err %.3e\n", sqrt(e / nrm)); if (sqrt(e / nrm) > 2e-6) bad++;
  inverse();
  e = 0; nrm = 0;
  for (int l = 0; l < 32; ++l) for (int q = 0; q < 16; ++q) {
    cd want = 512.0 * cd(z[l + 32 * q].x, z[l + 32 * q].y);
    e += std::norm(cd(v[l][q].x, v[l][q].y) - want); nrm += std::norm(want);
  }
  printf("round trip rel err %.3e\n", sqrt(e / nrm)); if (sqrt(e / nrm) > 2e-6) bad++;

  // ---- one full frame iteration: window omitted (x is the already windowed frame) ----
  std::vector<float2> P(M);      // previous rebuilt, packed slot 0 = (Re P0, Re PM)
  std::vector<float> mag(M + 4);
  for (int k = 0; k < M; ++k) P[k] = make_float2(20 * frand(seed), 20 * frand(seed));
  for (int k = 0; k <= M; ++k) mag[k] = 3.0f * (frand(seed) + 0.5f);
  const float mom = 0.99f / 1.99f;
  // reference in double
  std::vector<cd> X(M + 1), Y(M + 1);
  for (int k = 0; k <= M; ++k) { cd a = 0; for (int n = 0; n < N; ++n) a += (double)x[n] * std::polar(1.0, -2 * M_PI * (double)((long)n * k % N) / N); X[k] = a; }
  for (int k = 0; k <= M; ++k) {
    cd p = (k == 0) ? cd(P[0].x, 0) : (k == M) ? cd(P[0].y, 0) : cd(P[k].x, P[k].y);  // P = 2 * previous rebuilt
    cd a = 2.0 * X[k] - (double)mom * p;
    if (k == 0 || k == M) a = cd(a.real(), 0);
    Y[k] = (double)mag[k] * a / (std::abs(a) + 1e-16);
  }
  std::vector<double> yref(N);
  for (int n = 0; n < N; ++n) {
    cd a = Y[0].real() + Y[M].real() * ((n & 1) ? -1.0 : 1.0);
    for (int k = 1; k < M; ++k) a += 2.0 * (Y[k] * std::polar(1.0, 2 * M_PI * (double)((long)n * k % N) / N)).real();
    yref[n] = a.real();  // = N * irfft(Y)
  }
  for (int l = 0; l < 32; ++l) for (int q = 0; q < 16; ++q) v[l][q] = z[l + 32 * q];
  forward();
  std::vector<float2> newP(M, make_float2(0, 0));
  std::vector<int> stored(M, 0);
  for (int l = 0; l < 32; ++l) {
    if (l == 0) lane0_permute(v[l]);
    for (int r = 0; r < 8; ++r) {
      const int k = slot_k(l, r);
      float2& U = v[l][2 * r];
      float2& V = v[l][2 * (7 - r) + 1];
      if (l == 0 && r == 0) {
        float2 x0M, x256;
        special_update(U, V, P[0], P[256], mag[0], mag[M], mag[256], mom, true, x0M, x256);
        newP[0] = x0M; newP[256] = x256; stored[0]++; stored[256]++;
      } else {
        float2 xk, xmk;
        pair_update(U, V, rt[k], P[k], P[M - k], mag[k], mag[M - k], mom, true, xk, xmk);
        newP[k] = xk; newP[M - k] = xmk; stored[k]++; stored[M - k]++;
      }
    }
    if (l == 0) lane0_unpermute(v[l]);
  }
  for (int k = 0; k < M; ++k) if (stored[k] != 1) { printf("tprev bin %d stored %d times\n", k, stored[k]); bad++; }
  e = 0; nrm = 0;
  for (int k = 1; k < M; ++k) { e += std::norm(cd(newP[k].x, newP[k].y) - 2.0 * X[k]); nrm += std::norm(2.0 * X[k]); }  // tprev holds 2 * rebuilt
  e += std::norm(cd(newP[0].x, 0) - 2.0 * X[0]) + std::norm(cd(newP[0].y, 0) - 2.0 * X[M]);
  printf("rebuilt (tprev) rel err %.3e\n", sqrt(e / nrm)); if (sqrt(e / nrm) > 2e-6) bad++;
  inverse();
  e = 0; nrm = 0;
  for (int l = 0; l < 32; ++l) for (int q = 0; q < 16; ++q) {
    int m = l + 32 * q;
    e += pow(v[l][q].x - yref[2 * m], 2) + pow(v[l][q].y - yref[2 * m + 1], 2);
    nrm += pow(yref[2 * m], 2) + pow(yref[2 * m + 1], 2);
  }
  printf("frame iteration rel err %.3e\n", sqrt(e / nrm)); if (sqrt(e / nrm) > 5e-6) bad++;
  // ---- n_fft = 512: two real frames in one complex transform -------------------------------------------------------
  {
    const int N2 = 512, H2 = 256;
    std::vector<float> fa(N2), fb(N2);
    for (auto& q : fa) q = frand(seed);
    for (auto& q : fb) q = frand(seed);
    std::vector<float2> Pa(H2), Pb(H2);  // 2 x previous rebuilt, slot 0 packed (DC, Nyquist)
    std::vector<float> ma(H2 + 4), mb(H2 + 4);
    for (int k = 0; k < H2; ++k) { Pa[k] = make_float2(10 * frand(seed), 10 * frand(seed)); Pb[k] = make_float2(10 * frand(seed), 10 * frand(seed)); }
    for (int k = 0; k <= H2; ++k) { ma[k] = 2.0f * (frand(seed) + 0.6f); mb[k] = 2.0f * (frand(seed) + 0.6f); }
    auto ref = [&](const std::vector<float>& f, const std::vector<float2>& P, const std::vector<float>& mg, std::vector<double>& y, std::vector<cd>& X2) {
      std::vector<cd> X(H2 + 1), Y(H2 + 1);
      for (int k = 0; k <= H2; ++k) { cd a = 0; for (int n = 0; n < N2; ++n) a += (double)f[n] * std::polar(1.0, -2 * M_PI * (double)((long)n * k % N2) / N2); X[k] = a; }
      X2.resize(H2 + 1);
      for (int k = 0; k <= H2; ++k) {
        cd p = (k == 0) ? cd(P[0].x, 0) : (k == H2) ? cd(P[0].y, 0) : cd(P[k].x, P[k].y);
        cd a = 2.0 * X[k] - (double)mom * p;
        if (k == 0 || k == H2) a = cd(a.real(), 0);
        Y[k] = (double)mg[k] * a / (std::abs(a) + 1e-16);
        X2[k] = 2.0 * X[k];
      }
      y.resize(N2);
      for (int n = 0; n < N2; ++n) {
        cd a = Y[0].real() + Y[H2].real() * ((n & 1) ? -1.0 : 1.0);
        for (int k = 1; k < H2; ++k) a += 2.0 * (Y[k] * std::polar(1.0, 2 * M_PI * (double)((long)n * k % N2) / N2)).real();
        y[n] = a.real();
      }
    };
    std::vector<double> ya, yb; std::vector<cd> Xa2, Xb2;
    ref(fa, Pa, ma, ya, Xa2); ref(fb, Pb, mb, yb, Xb2);
    for (int l = 0; l < 32; ++l) for (int q = 0; q < 16; ++q) v[l][q] = make_float2(fa[l + 32 * q], fb[l + 32 * q]);
    forward();
    std::vector<float2> nPa(H2, make_float2(0, 0)), nPb(H2, make_float2(0, 0));
    std::vector<int> st2(H2, 0);
    for (int l = 0; l < 32; ++l) {
      if (l == 0) lane0_permute(v[l]);
      for (int r = 0; r < 8; ++r) {
        const int kU = slot_k(l, r);
        float2& U = v[l][2 * r];
        float2& V = v[l][2 * (7 - r) + 1];
        if (l == 0 && r == 0) {
          float2 x0a, x0b;
          special2_update(U, V, Pa[0], Pb[0], ma[0], ma[H2], mb[0], mb[H2], mom, true, x0a, x0b);
          nPa[0] = x0a; nPb[0] = x0b; st2[0]++;
        } else {
          const bool swap = (kU > H2);
          const int kb = swap ? N2 - kU : kU;
          float2 xa, xb;
          if (swap) pair2_update(V, U, Pa[kb], Pb[kb], ma[kb], mb[kb], mom, true, xa, xb);
          else pair2_update(U, V, Pa[kb], Pb[kb], ma[kb], mb[kb], mom, true, xa, xb);
          nPa[kb] = xa; nPb[kb] = xb; st2[kb]++;
        }
      }
      if (l == 0) lane0_unpermute(v[l]);
    }
    for (int k = 0; k < H2; ++k) if (st2[k] != 1) { printf("n_fft512: bin %d handled %d times\n", k, st2[k]); bad++; }
    double e1 = 0, n1 = 0;
    for (int k = 1; k < H2; ++k) {
      e1 += std::norm(cd(nPa[k].x, nPa[k].y) - Xa2[k]) + std::norm(cd(nPb[k].x, nPb[k].y) - Xb2[k]);
      n1 += std::norm(Xa2[k]) + std::norm(Xb2[k]);
    }
    e1 += std::norm(cd(nPa[0].x, 0) - Xa2[0]) + std::norm(cd(nPa[0].y, 0) - Xa2[H2]) + std::norm(cd(nPb[0].x, 0) - Xb2[0]) + std::norm(cd(nPb[0].y, 0) - Xb2[H2]);
    printf("n_fft512 rebuilt rel err %.3e\n", sqrt(e1 / n1)); if (sqrt(e1 / n1) > 2e-6) bad++;
    inverse();
    double e2 = 0, n2 = 0;
    for (int l = 0; l < 32; ++l) for (int q = 0; q < 16; ++q) {
      const int m = l + 32 * q;
      e2 += pow(v[l][q].x - ya[m], 2) + pow(v[l][q].y - yb[m], 2);
      n2 += pow(ya[m], 2) + pow(yb[m], 2);
    }
    printf("n_fft512 frame-pair iteration rel err %.3e\n", sqrt(e2 / n2)); if (sqrt(e2 / n2) > 5e-6) bad++;
  }
  // ---- n_fft = 2048: radix-2 step over two 512-point register transforms ------------------------------------------------
  {
    const int N4 = 2048, M4 = 1024;
    std::vector<float> f(N4);
    for (auto& q : f) q = frand(seed);
    std::vector<float2> P4(M4), rt4(M4);
    std::vector<float> mg4(M4 + 4);
    for (int k = 0; k < M4; ++k) {
      P4[k] = make_float2(30 * frand(seed), 30 * frand(seed));
      rt4[k] = make_float2((float)cos(-2 * M_PI * k / N4), (float)sin(-2 * M_PI * k / N4));
    }
    for (int k = 0; k <= M4; ++k) mg4[k] = 3.0f * (frand(seed) + 0.6f);
    std::vector<cd> X(M4 + 1), Y(M4 + 1);
    for (int k = 0; k <= M4; ++k) { cd a = 0; for (int n = 0; n < N4; ++n) a += (double)f[n] * std::polar(1.0, -2 * M_PI * (double)((long)n * k % N4) / N4); X[k] = a; }
    for (int k = 0; k <= M4; ++k) {
      cd p = (k == 0) ? cd(P4[0].x, 0) : (k == M4) ? cd(P4[0].y, 0) : cd(P4[k].x, P4[k].y);
      cd a = 2.0 * X[k] - (double)mom * p;
      if (k == 0 || k == M4) a = cd(a.real(), 0);
      Y[k] = (double)mg4[k] * a / (std::abs(a) + 1e-16);
    }
    std::vector<double> yr(N4);
    for (int n = 0; n < N4; ++n) {
      cd a = Y[0].real() + Y[M4].real() * ((n & 1) ? -1.0 : 1.0);
      for (int k = 1; k < M4; ++k) a += 2.0 * (Y[k] * std::polar(1.0, 2 * M_PI * (double)((long)n * k % N4) / N4)).real();
      yr[n] = a.real();
    }
    static float2 ve[32][16], vo[32][16];
    std::vector<float2> S2(XCH);
    for (int l = 0; l < 32; ++l) for (int q = 0; q < 16; ++q) {
      const int m = l + 32 * q;
      ve[l][q] = make_float2(f[4 * m], f[4 * m + 1]);
      vo[l][q] = make_float2(f[4 * m + 2], f[4 * m + 3]);
    }
    auto fwd = [&](float2 (*w)[16]) {
      for (int l = 0; l < 32; ++l) fwd1_store(l, w[l], lt[l], S2.data());
      for (int l = 0; l < 32; ++l) fwd2_load(l, w[l], S2.data());
      for (int l = 0; l < 32; ++l) fwd2_store(l, w[l], lt[l], S2.data());
      for (int l = 0; l < 32; ++l) fwd3_load(l, w[l], S2.data());
    };
    auto inv = [&](float2 (*w)[16]) {
      for (int l = 0; l < 32; ++l) inv1_store(l, w[l], S2.data());
      for (int l = 0; l < 32; ++l) inv2_load(l, w[l], lt[l], S2.data());
      for (int l = 0; l < 32; ++l) inv2_store(l, w[l], S2.data());
      for (int l = 0; l < 32; ++l) inv3_load(l, w[l], lt[l], S2.data());
    };
    fwd(ve); fwd(vo);
    std::vector<float2> nP(M4, make_float2(1e30f, 1e30f));
    for (int l = 0; l < 32; ++l) {
      if (l == 0) { lane0_permute(ve[l]); lane0_permute(vo[l]); }
      for (int r = 0; r < 8; ++r) {
        const int k = slot_k(l, r);
        if (l == 0 && r == 0) quad_special(ve[l][0], ve[l][15], vo[l][0], vo[l][15], rt4[256], P4.data(), mg4.data(), mom, true, nP.data());
        else quad_update(ve[l][2 * r], ve[l][2 * (7 - r) + 1], vo[l][2 * r], vo[l][2 * (7 - r) + 1], k, rt4[2 * k], rt4[k], rt4[512 - k],
                         P4.data(), mg4.data(), mom, true, nP.data());
      }
      if (l == 0) { lane0_unpermute(ve[l]); lane0_unpermute(vo[l]); }
    }
    double e4 = 0, n4 = 0;
    for (int k = 1; k < M4; ++k) { e4 += std::norm(cd(nP[k].x, nP[k].y) - 2.0 * X[k]); n4 += std::norm(2.0 * X[k]); }
    e4 += std::norm(cd(nP[0].x, 0) - 2.0 * X[0]) + std::norm(cd(nP[0].y, 0) - 2.0 * X[M4]);
    printf("n_fft2048 rebuilt rel err %.3e\n", sqrt(e4 / n4)); if (!(sqrt(e4 / n4) < 2e-6)) bad++;
    inv(ve); inv(vo);
    e4 = 0; n4 = 0;
    for (int l = 0; l < 32; ++l) for (int q = 0; q < 16; ++q) {
      const int m = l + 32 * q;
      e4 += pow(ve[l][q].x - yr[4 * m], 2) + pow(ve[l][q].y - yr[4 * m + 1], 2) + pow(vo[l][q].x - yr[4 * m + 2], 2) + pow(vo[l][q].y - yr[4 * m + 3], 2);
      n4 += pow(yr[4 * m], 2) + pow(yr[4 * m + 1], 2) + pow(yr[4 * m + 2], 2) + pow(yr[4 * m + 3], 2);
    }
    printf("n_fft2048 frame iteration rel err %.3e\n", sqrt(e4 / n4)); if (!(sqrt(e4 / n4) < 5e-6)) bad++;
  }
  printf(bad ? "FAIL\n" : "OK\n");
  return bad;
}
