// Host side of the pole expansion of x^(-1/2) (fcn_common.cuh): quadrature nodes and weights for the
// intervals [1, 2^q], q = 1..FCN_QMAX, with N = 32 nodes each.
//
//   x^(-1/2) = (2/pi) int_0^inf dt / (x + t^2),   t = sc(u | m1),  m1 = 1 - 2^-q,  u in (0, K(m1))
//            ~ sum_j c_j / (x + beta_j),   u_j = (j + 1/2) K/N,  beta_j = sc^2(u_j),  c_j = (2K/(pi N)) dn/cn^2
// (Hale, Higham, Trefethen 2008, method 3).  Jacobi elliptic functions by the descending Landen /
// AGM recurrence (Abramowitz & Stegun 16.4) in long double; nodes beyond K/2 use the quarter-period
// reflection so that cn is never formed by cancellation.  Every interval is checked against
// x^(-1/2) on a log-spaced grid when the table is built.
#include <cmath>
#include <mutex>
#include <vector>

#include "fcn_common.cuh"

namespace lk {

namespace {

typedef long double ld;

struct Landen {
  ld a[40], c[40];
  int n;
  ld K;
  explicit Landen(ld kc /* sqrt(1-m) */) {
    ld an = 1.0L, bn = kc, cn = sqrtl(1.0L - kc * kc);
    a[0] = an;
    c[0] = cn;
    n = 0;
    while (fabsl(cn) > 1e-19L * an && n < 38) {
      const ld a1 = 0.5L * (an + bn), c1 = 0.5L * (an - bn), b1 = sqrtl(an * bn);
      an = a1;
      bn = b1;
      cn = c1;
      ++n;
      a[n] = an;
      c[n] = cn;
    }
    K = 3.14159265358979323846264338327950288L / (2.0L * an);
  }
  void ellipj(ld u, ld &sn, ld &cn, ld &dn) const {
    ld phi = ldexpl(a[n] * u, n), prev = phi;
    for (int i = n; i >= 1; --i) {
      prev = phi;
      phi = 0.5L * (phi + asinl(c[i] * sinl(phi) / a[i]));
    }
    sn = sinl(phi);
    cn = cosl(phi);
    dn = n >= 1 ? cn / cosl(prev - phi) : 1.0L;
  }
};

std::vector<double> build_table(double *worst_err) {
  std::vector<double> tab((size_t)(FCN_QMAX + 1) * 2 * FCN_NP, 0.0);
  const int N = FCN_NP;
  double worst = 0.0;
  for (int q = 1; q <= FCN_QMAX; ++q) {
    const ld k2 = ldexpl(1.0L, -q);  // 1 - m1
    const ld kc = sqrtl(k2);
    Landen L(kc);
    double *c = &tab[(size_t)(q * 2 + 0) * N], *b = &tab[(size_t)(q * 2 + 1) * N];
    const ld h = 2.0L * L.K / (3.14159265358979323846264338327950288L * N);
    for (int j = 0; j < N; ++j) {
      ld sn, cn, dn, beta, w;
      if (2 * j + 1 <= N) {
        L.ellipj((j + 0.5L) * L.K / N, sn, cn, dn);
        beta = (sn * sn) / (cn * cn);
        w = dn / (cn * cn);
      } else {  // u = K - v: sn(u) = cn(v)/dn(v), cn(u) = kc sn(v)/dn(v), dn(u) = kc/dn(v)
        L.ellipj((N - j - 0.5L) * L.K / N, sn, cn, dn);
        beta = (cn * cn) / (k2 * sn * sn);
        w = dn / (kc * sn * sn);
      }
      c[j] = (double)(h * w);
      b[j] = (double)beta;
    }
    // check: relative error of the expansion on [1, 2^q]
    double err = 0.0;
    const int NS = 400;
    for (int s = 0; s <= NS; ++s) {
      const ld x = expl(logl(2.0L) * q * s / NS);
      ld r = 0.0L;
      for (int j = 0; j < N; ++j) r += (ld)c[j] / (x + (ld)b[j]);
      const double e = (double)fabsl(r * sqrtl(x) - 1.0L);
      if (e > err) err = e;
    }
    // q <= 27: < 5e-13; wider intervals degrade gracefully (q = 34: ~1e-10)
    if (q <= 27 && err > worst) worst = err;
  }
  if (worst_err) *worst_err = worst;
  return tab;
}

}  // namespace

const std::vector<double> &fcn_pole_table_host() {
  static std::vector<double> tab;
  static std::once_flag once;
  std::call_once(once, [] {
    double worst = 0.0;
    tab = build_table(&worst);
    if (!(worst < 2e-12)) throw Error("fcn pole table failed its accuracy check (" + std::to_string(worst) + ")");
  });
  return tab;
}

}  // namespace lk

// C entry point for the CPU test tier: copies the table (returns its length in doubles)
extern "C" int letkf_b200_selftest_pole_table(double *out, int cap) {
  try {
    const std::vector<double> &t = lk::fcn_pole_table_host();
    if (out)
      for (int i = 0; i < cap && i < (int)t.size(); ++i) out[i] = t[i];
    return (int)t.size();
  } catch (...) {
    return -1;
  }
}
