// Host-side feasibility probe for packed staging: T threads each walk their share of B dense QPs (Q 60x60, A 38x60, row-major
// doubles), (a) summing everything (pure read), (b) classifying + gathering the entries a reduced solve needs into a packed
// buffer.  Prints GB/s of dense input consumed and the time per 2^16 QPs.   g++ -O3 -march=x86-64-v3 -pthread
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>
int main(int argc, char** argv) {
  const int T = argc > 1 ? atoi(argv[1]) : (int)std::thread::hardware_concurrency();
  const long B = argc > 2 ? atol(argv[2]) : 32768;
  const int n = 60, m = 38;
  const size_t per = (size_t)n * n + (size_t)m * n;
  std::vector<double> in(per * B);
  for (long q = 0; q < B; ++q) {
    double* Q = &in[q * per]; double* A = Q + n * n;
    memset(Q, 0, per * sizeof(double));
    for (int i = 0; i < n; ++i) Q[i * n + i] = 1.0 + i;
    for (int i = 0; i < 22; ++i) for (int j = 0; j < 22; ++j) Q[i * n + j] += 0.01 * (i + j + 1);
    for (int k = 0; k < m; ++k) for (int j = 0; j < 39; ++j) A[k * n + j] = 0.1 * (k + j + 1);
    for (int j = 39; j < n; ++j) A[(j % m) * n + j] = 1.0;
  }
  std::vector<double> out((size_t)2200 * B);
  std::vector<double> sums(T);
  for (int mode = 0; mode < 2; ++mode) {
    for (int rep = 0; rep < 3; ++rep) {
      auto t0 = std::chrono::steady_clock::now();
      std::vector<std::thread> th;
      for (int t = 0; t < T; ++t) th.emplace_back([&, t] {
        double s = 0.0;
        for (long q = B * t / T; q < B * (t + 1) / T; ++q) {
          const double* Q = &in[q * per]; const double* A = Q + n * n;
          if (mode == 0) { for (size_t e = 0; e < per; ++e) s += Q[e]; continue; }
          // classification: off-diagonal flag per column (symmetric: per row), nnz per A column
          int type[64], nnz[64]; unsigned char off[64];
          for (int i = 0; i < n; ++i) { bool o = false; const double* r = Q + i * n; for (int j = 0; j < n; ++j) o |= (j != i) & (r[j] != 0.0); off[i] = o; }
          for (int j = 0; j < n; ++j) nnz[j] = 0;
          for (int k = 0; k < m; ++k) { const double* r = A + k * n; for (int j = 0; j < n; ++j) nnz[j] += r[j] != 0.0; }
          int cols[64], nc = 0, rl[64], nr = 0;
          for (int j = 0; j < n; ++j) { type[j] = off[j] ? 0 : (Q[j * n + j] == 0.0 ? 3 : (nnz[j] <= 1 ? 2 : 1)); if (type[j] == 0) rl[nr++] = j; if (type[j] != 2) cols[nc++] = j; }
          double* o = &out[(size_t)2200 * q];
          for (int a = 0; a < nr; ++a) for (int b = 0; b <= a; ++b) *o++ = Q[rl[a] * n + rl[b]];
          for (int k = 0; k < m; ++k) { const double* r = A + k * n; for (int c = 0; c < nc; ++c) *o++ = r[cols[c]]; }
          for (int j = 0; j < n; ++j) *o++ = Q[j * n + j];
          s += o[-1];
        }
        sums[t] = s;
      });
      for (auto& x : th) x.join();
      const double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
      printf("%s T=%d B=%ld rep %d: %.1f ms, %.1f GB/s of dense input, %.1f ms per 2^16 QPs\n", mode ? "classify+pack" : "read-only    ", T, B, rep,
             1e3 * dt, per * B * 8 / dt / 1e9, 1e3 * dt * 65536.0 / B);
    }
  }
  return sums[0] == 1.2345;
}
