// CPU emulation of one warp of the generic register-FFT Griffin-Lim path (csrc/gl_reg.cuh), M = 64 * R3 for
// R3 = 4, 5, 8, 12 (n_fft 512, 640, 1024, 1536): the three-pass data flow against a naive DFT, the round trip, bin
// coverage of the lane-local pairing (every bin exactly once, each with its mirror), and one whole frame
// (forward -> projection x magnitude -> inverse) against a double-precision evaluation.
#include <cmath>
#include <complex>
#include <cstdio>
#include <vector>
#include "../../audio_denoising_b200/csrc/gl_reg.cuh"
using namespace b2d;
using namespace b2d::regfft;
typedef std::complex<double> cd;

static float frand(unsigned& s) { s = s * 1664525u + 1013904223u; return (s >> 8) / 16777216.0f - 0.5f; }

template <int R3>
int run() {
  typedef Geo<R3> G;
  constexpr int M = G::M, N = G::N, NV = G::NV;
  int bad = 0;
  std::vector<float2> tw(M), rt(M);
  for (int k = 0; k < M; ++k) {
    tw[k] = make_float2((float)cos(-2 * M_PI * k / M), (float)sin(-2 * M_PI * k / M));
    rt[k] = make_float2((float)cos(-2 * M_PI * k / N), (float)sin(-2 * M_PI * k / N));
  }
  std::vector<LaneTwR<R3>> lt(32);
  for (int l = 0; l < 32; ++l) lane_twiddles_r<R3>(l, tw.data(), lt[l]);
  unsigned seed = 7 + R3;
  std::vector<float> x(N);
  for (auto& v : x) v = frand(seed);
  std::vector<float2> z(M);
  for (int m = 0; m < M; ++m) z[m] = make_float2(x[2 * m], x[2 * m + 1]);
  std::vector<cd> Z(M);
  for (int k = 0; k < M; ++k) { cd a = 0; for (int m = 0; m < M; ++m) a += cd(z[m].x, z[m].y) * std::polar(1.0, -2 * M_PI * (double)((long)m * k % M) / M); Z[k] = a; }

  std::vector<std::vector<float2>> v(32, std::vector<float2>(NV)), wA(32, std::vector<float2>(R3)), wB(32, std::vector<float2>(R3));
  std::vector<float2> S(G::XCH, make_float2(1e30f, 1e30f));
  auto load = [&]() {
    for (int l = 0; l < 32; ++l) for (int r = 0; r < G::NR; ++r) for (int n1 = 0; n1 < 8; ++n1) {
      const int i = l + 32 * r;
      v[l][8 * r + n1] = (i < G::NB) ? z[i + G::NB * n1] : make_float2(0, 0);
    }
  };
  auto forward = [&]() {
    for (int l = 0; l < 32; ++l) fwd1_store_r<R3>(l, v[l].data(), lt[l], S.data());
    for (int l = 0; l < 32; ++l) fwd2_load_r<R3>(l, v[l].data(), S.data());
    for (int l = 0; l < 32; ++l) fwd2_store_r<R3>(l, v[l].data(), lt[l], S.data());
    for (int l = 0; l < 32; ++l) fwd3_load_r<R3>(l, wA[l].data(), wB[l].data(), S.data());
  };
  auto inverse = [&]() {
    for (int l = 0; l < 32; ++l) inv1_store_r<R3>(l, wA[l].data(), wB[l].data(), S.data());
    for (int l = 0; l < 32; ++l) inv2_load_r<R3>(l, v[l].data(), lt[l], S.data());
    for (int l = 0; l < 32; ++l) inv2_store_r<R3>(l, v[l].data(), S.data());
    for (int l = 0; l < 32; ++l) inv3_load_r<R3>(l, v[l].data(), lt[l], S.data());
  };
  load();
  forward();
  double e = 0, nrm = 0;
  std::vector<int> seen(M, 0);
  for (int l = 0; l < 32; ++l) for (int f = 0; f < 2; ++f) for (int k3 = 0; k3 < R3; ++k3) {
    const int k = fast512::fam(l, f) + 64 * k3;
    seen[k]++;
    const float2 g = f ? wB[l][k3] : wA[l][k3];
    e += std::norm(cd(g.x, g.y) - Z[k]); nrm += std::norm(Z[k]);
  }
  for (int k = 0; k < M; ++k) if (seen[k] != 1) { printf("R3=%d bin %d covered %d times\n", R3, k, seen[k]); bad++; }
  printf("R3=%2d forward rel err %.3e\n", R3, sqrt(e / nrm)); if (!(sqrt(e / nrm) < 2e-6)) bad++;
  inverse();
  e = 0; nrm = 0;
  for (int l = 0; l < 32; ++l) for (int r = 0; r < G::NR; ++r) for (int n1 = 0; n1 < 8; ++n1) {
    const int i = l + 32 * r;
    if (i >= G::NB) continue;
    const cd want = (double)M * cd(z[i + G::NB * n1].x, z[i + G::NB * n1].y);
    e += std::norm(cd(v[l][8 * r + n1].x, v[l][8 * r + n1].y) - want); nrm += std::norm(want);
  }
  printf("R3=%2d round trip rel err %.3e\n", R3, sqrt(e / nrm)); if (!(sqrt(e / nrm) < 2e-6)) bad++;

  // pairing: every bin 0..M exactly once, partner = mirror
  std::vector<int> hit(M + 1, 0);
  for (int l = 0; l < 32; ++l) {
    // tag each register with its bin through gather_pairs
    std::vector<float2> tA(R3), tB(R3), U(R3), V(R3);
    for (int k3 = 0; k3 < R3; ++k3) { tA[k3] = make_float2((float)(fast512::fam(l, 0) + 64 * k3), 0); tB[k3] = make_float2((float)(fast512::fam(l, 1) + 64 * k3), 0); }
    gather_pairs<R3>(l, tA.data(), tB.data(), U.data(), V.data());
    for (int r = 0; r < R3; ++r) {
      const int k = slot_k_r<R3>(l, r), ku = (int)U[r].x, kv = (int)V[r].x;
      if (l == 0 && r == 0) {
        if (ku != 0 || kv != M / 2) { printf("R3=%d special slot holds bins %d %d\n", R3, ku, kv); bad++; }
        hit[0]++; hit[M]++; hit[M / 2]++;
      } else {
        if (ku != k || kv != M - k) { printf("R3=%d lane %d slot %d: k=%d holds %d / %d\n", R3, l, r, k, ku, kv); bad++; }
        hit[k]++; hit[M - k]++;
      }
    }
    std::vector<float2> bA(R3), bB(R3);
    scatter_pairs<R3>(l, bA.data(), bB.data(), U.data(), V.data());
    for (int k3 = 0; k3 < R3; ++k3) if (bA[k3].x != tA[k3].x || bB[k3].x != tB[k3].x) { printf("R3=%d lane %d scatter is not the inverse of gather\n", R3, l); bad++; break; }
  }
  for (int k = 0; k <= M; ++k) if (hit[k] != 1) { printf("R3=%d bin %d paired %d times\n", R3, k, hit[k]); bad++; }

  // one whole frame: Y = mag * X / (|X| + 1e-16), y = N * irfft(Y)
  std::vector<float> mag(M + 4);
  for (int k = 0; k <= M; ++k) mag[k] = 3.0f * (frand(seed) + 0.5f);
  std::vector<cd> X(M + 1), Y(M + 1);
  for (int k = 0; k <= M; ++k) { cd a = 0; for (int n = 0; n < N; ++n) a += (double)x[n] * std::polar(1.0, -2 * M_PI * (double)((long)n * k % N) / N); X[k] = a; }
  for (int k = 0; k <= M; ++k) {
    cd a = X[k];
    if (k == 0 || k == M) a = cd(a.real(), 0);
    Y[k] = (double)mag[k] * a / (std::abs(a) + 1e-16);
  }
  std::vector<double> yref(N);
  for (int n = 0; n < N; ++n) {
    cd a = Y[0].real() + Y[M].real() * ((n & 1) ? -1.0 : 1.0);
    for (int k = 1; k < M; ++k) a += 2.0 * (Y[k] * std::polar(1.0, 2 * M_PI * (double)((long)n * k % N) / N)).real();
    yref[n] = a.real();
  }
  load();
  forward();
  for (int l = 0; l < 32; ++l) project_frame<R3>(l, wA[l].data(), wB[l].data(), rt.data(), mag.data());
  inverse();
  e = 0; nrm = 0;
  for (int l = 0; l < 32; ++l) for (int r = 0; r < G::NR; ++r) for (int n1 = 0; n1 < 8; ++n1) {
    const int i = l + 32 * r;
    if (i >= G::NB) continue;
    const int m = i + G::NB * n1;
    e += pow(v[l][8 * r + n1].x - yref[2 * m], 2) + pow(v[l][8 * r + n1].y - yref[2 * m + 1], 2);
    nrm += pow(yref[2 * m], 2) + pow(yref[2 * m + 1], 2);
  }
  printf("R3=%2d frame iteration rel err %.3e\n", R3, sqrt(e / nrm)); if (!(sqrt(e / nrm) < 5e-6)) bad++;
  return bad;
}

int main() {
  int bad = run<4>() + run<5>() + run<8>() + run<12>();
  printf(bad ? "FAILED (%d)\n" : "OK\n", bad);
  return bad ? 1 : 0;
}
