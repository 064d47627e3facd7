// CPU-side unit test of the index math in csrc/fft.cuh (compiled by nvcc, runs on the host only).
#include <cmath>
#include <complex>
#include <cstdio>
#include <vector>
#include "../../audio_denoising_b200/csrc/fft.cuh"
using namespace b2d;
typedef std::complex<double> cd;

template <bool INV>
static std::vector<float2> run_fft(std::vector<float2> a, int M, const std::vector<int>& radix, const std::vector<float2>& tw) {
  std::vector<float2> b(M);
  int Ns = 1;
  for (int R : radix) {
    for (int w = 0; w < M / R; ++w) {
      switch (R) {
        case 8: stockham_item<INV, 8>(a.data(), b.data(), w, M, M, Ns, tw.data()); break;
        case 4: stockham_item<INV, 4>(a.data(), b.data(), w, M, M, Ns, tw.data()); break;
        case 2: stockham_item<INV, 2>(a.data(), b.data(), w, M, M, Ns, tw.data()); break;
        case 3: stockham_item<INV, 3>(a.data(), b.data(), w, M, M, Ns, tw.data()); break;
        default: stockham_item<INV, 5>(a.data(), b.data(), w, M, M, Ns, tw.data()); break;
      }
    }
    a.swap(b);
    Ns *= R;
  }
  return a;
}

int main() {
  int bad = 0;
  struct Case { int M; std::vector<int> r; };
  std::vector<Case> cases = {{512, {8, 8, 8}}, {256, {8, 8, 4}}, {1024, {8, 8, 8, 2}}, {320, {8, 8, 5}}, {768, {8, 8, 4, 3}},
                             {32, {8, 4}}, {2048, {8, 8, 8, 4}}, {60, {4, 3, 5}}, {50, {2, 5, 5}}, {36, {4, 3, 3}}};
  for (auto& c : cases) {
    int M = c.M;
    std::vector<float2> tw(M), x(M);
    for (int k = 0; k < M; ++k) tw[k] = make_float2((float)cos(-2 * M_PI * k / M), (float)sin(-2 * M_PI * k / M));
    unsigned s = 12345u + M;
    for (int i = 0; i < M; ++i) {
      s = s * 1664525u + 1013904223u; float a = (s >> 8) / 16777216.0f - 0.5f;
      s = s * 1664525u + 1013904223u; float b = (s >> 8) / 16777216.0f - 0.5f;
      x[i] = make_float2(a, b);
    }
    for (int inv = 0; inv < 2; ++inv) {
      auto y = inv ? run_fft<true>(x, M, c.r, tw) : run_fft<false>(x, M, c.r, tw);
      double err = 0, nrm = 0;
      for (int k = 0; k < M; ++k) {
        cd acc = 0;
        for (int n = 0; n < M; ++n) acc += cd(x[n].x, x[n].y) * std::polar(1.0, (inv ? 2 : -2) * M_PI * (double)((long long)n * k % M) / M);
        err += std::norm(acc - cd(y[k].x, y[k].y)); nrm += std::norm(acc);
      }
      double rel = sqrt(err / nrm);
      printf("M=%d inv=%d rel=%.3e\n", M, inv, rel);
      if (rel > 2e-6) bad++;
    }
  }
  // real split / merge on N = 64
  {
    const int N = 64, M = N / 2;
    std::vector<double> x(N);
    for (int i = 0; i < N; ++i) x[i] = sin(0.3 * i) + 0.2 * cos(1.1 * i + 0.5) + 0.01 * i;
    std::vector<cd> Z(M), X(M + 1);
    for (int k = 0; k < M; ++k) { cd a = 0; for (int m = 0; m < M; ++m) a += cd(x[2 * m], x[2 * m + 1]) * std::polar(1.0, -2 * M_PI * m * k / M); Z[k] = a; }
    for (int k = 0; k <= M; ++k) { cd a = 0; for (int n = 0; n < N; ++n) a += x[n] * std::polar(1.0, -2 * M_PI * n * k / N); X[k] = a; }
    double e1 = 0, e2 = 0;
    for (int k = 0; k <= M / 2; ++k) {
      float2 rt = make_float2((float)cos(-2 * M_PI * k / N), (float)sin(-2 * M_PI * k / N));
      cd zk = Z[k], zm = Z[(M - k) % M];
      float2 xk, xmk;
      rfft_split(make_float2(zk.real(), zk.imag()), make_float2(zm.real(), zm.imag()), rt, xk, xmk);
      e1 = fmax(e1, std::abs(cd(xk.x, xk.y) - X[k]));
      e1 = fmax(e1, std::abs(cd(xmk.x, xmk.y) - X[M - k]));
      // merge: from X[k], X[M-k] back to Z' (= 2 * Z with this scaling)
      float2 zk2, zmk2;
      irfft_merge(make_float2(X[k].real(), X[k].imag()), make_float2(X[M - k].real(), X[M - k].imag()), rt, zk2, zmk2);
      e2 = fmax(e2, std::abs(cd(zk2.x, zk2.y) - 2.0 * zk));
      if (k != 0) e2 = fmax(e2, std::abs(cd(zmk2.x, zmk2.y) - 2.0 * zm));
    }
    printf("split err %.3e merge err %.3e\n", e1, e2);
    if (e1 > 1e-4 || e2 > 1e-4) bad++;
  }
  printf(bad ? "FAIL\n" : "OK\n");
  return bad;
}
