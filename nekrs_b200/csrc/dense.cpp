// dense.cpp -- the two small dense eigen-solves of the setup phase, hand-written because no LAPACK
// is linked:  (1) the symmetric-definite generalized problem A v = lambda B v of the FDM setup
// (reference: dsygv_, ellipticMultiGridSchwarz.cpp:455-511), (2) the spectral radius of the small
// Arnoldi Hessenberg matrix (reference: dgeev_, ellipticMultiGridLevelSetup.cpp:260-290,434-446).
#include <algorithm>
#include <cmath>
#include <complex>

#include "host.hpp"

namespace nrsb {

uint64_t splitmix64(uint64_t x)
{
  x += 0x9E3779B97F4A7C15ull;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  return x ^ (x >> 31);
}

// A, B: n x n symmetric (any storage order), B positive definite.  On return A holds the
// eigenvectors with dsygv's convention: column-major, A[i + n*j] = component i of eigenvector j,
// normalised so that V^T B V = I; lam ascending.  Cholesky reduction + cyclic Jacobi.
int sym_generalized_eig(int n, std::vector<double>& A, std::vector<double>& B, std::vector<double>& lam)
{
  // B = L L^T
  std::vector<double> L(n * n, 0.0);
  for (int j = 0; j < n; ++j) {
    double d = B[j + n * j];
    for (int k = 0; k < j; ++k) d -= L[j + n * k] * L[j + n * k];
    if (!(d > 0.0)) return j + 1;  // leading minor not positive definite (dsygv info = n + j)
    L[j + n * j] = std::sqrt(d);
    for (int i = j + 1; i < n; ++i) {
      double s = B[i + n * j];
      for (int k = 0; k < j; ++k) s -= L[i + n * k] * L[j + n * k];
      L[i + n * j] = s / L[j + n * j];
    }
  }
  // C = L^-1 A L^-T
  std::vector<double> C(n * n);
  for (int i = 0; i < n; ++i)
    for (int j = 0; j < n; ++j) C[i + n * j] = 0.5 * (A[i + n * j] + A[j + n * i]);
  // solve L X = C (column by column), then X L^T = Y
  for (int j = 0; j < n; ++j)
    for (int i = 0; i < n; ++i) {
      double s = C[i + n * j];
      for (int k = 0; k < i; ++k) s -= L[i + n * k] * C[k + n * j];
      C[i + n * j] = s / L[i + n * i];
    }
  for (int i = 0; i < n; ++i)
    for (int j = 0; j < n; ++j) {
      double s = C[i + n * j];
      for (int k = 0; k < j; ++k) s -= C[i + n * k] * L[j + n * k];
      C[i + n * j] = s / L[j + n * j];
    }
  for (int i = 0; i < n; ++i)
    for (int j = i + 1; j < n; ++j) C[i + n * j] = C[j + n * i] = 0.5 * (C[i + n * j] + C[j + n * i]);
  // cyclic Jacobi on C, accumulate rotations in Q
  std::vector<double> Q(n * n, 0.0);
  for (int i = 0; i < n; ++i) Q[i + n * i] = 1.0;
  for (int sweep = 0; sweep < 100; ++sweep) {
    double off = 0.0, diag = 0.0;
    for (int i = 0; i < n; ++i) {
      diag += C[i + n * i] * C[i + n * i];
      for (int j = i + 1; j < n; ++j) off += C[i + n * j] * C[i + n * j];
    }
    if (off <= 1e-32 * (diag + 1e-300)) break;
    for (int p = 0; p < n - 1; ++p)
      for (int q = p + 1; q < n; ++q) {
        const double apq = C[p + n * q];
        if (apq == 0.0) continue;
        const double app = C[p + n * p], aqq = C[q + n * q];
        const double theta = (aqq - app) / (2.0 * apq);
        const double t = (theta >= 0 ? 1.0 : -1.0) / (std::fabs(theta) + std::sqrt(theta * theta + 1.0));
        const double c = 1.0 / std::sqrt(t * t + 1.0), s = t * c;
        for (int k = 0; k < n; ++k) {
          const double akp = C[k + n * p], akq = C[k + n * q];
          C[k + n * p] = c * akp - s * akq;
          C[k + n * q] = s * akp + c * akq;
        }
        for (int k = 0; k < n; ++k) {
          const double apk = C[p + n * k], aqk = C[q + n * k];
          C[p + n * k] = c * apk - s * aqk;
          C[q + n * k] = s * apk + c * aqk;
        }
        for (int k = 0; k < n; ++k) {
          const double qkp = Q[k + n * p], qkq = Q[k + n * q];
          Q[k + n * p] = c * qkp - s * qkq;
          Q[k + n * q] = s * qkp + c * qkq;
        }
      }
  }
  // sort ascending
  std::vector<int> idx(n);
  for (int i = 0; i < n; ++i) idx[i] = i;
  std::sort(idx.begin(), idx.end(), [&](int a, int b) { return C[a + n * a] < C[b + n * b]; });
  lam.resize(n);
  // V = L^-T Q
  for (int jj = 0; jj < n; ++jj) {
    const int j = idx[jj];
    lam[jj] = C[j + n * j];
    for (int i = n - 1; i >= 0; --i) {
      double s = Q[i + n * j];
      for (int k = i + 1; k < n; ++k) s -= L[k + n * i] * A[k + n * jj];
      A[i + n * jj] = s / L[i + n * i];
    }
  }
  return 0;
}

// spectral radius of a small dense matrix (column-major H[i + n*j]); Hessenberg in practice.
// Unshifted-to-Wilkinson QR iterations on the complex Hessenberg form.
double hessenberg_spectral_radius(int n, std::vector<double> Hr)
{
  using cd = std::complex<double>;
  if (n == 0) return 0.0;
  std::vector<cd> H(n * n);
  for (int i = 0; i < n * n; ++i) H[i] = Hr[i];
  auto at = [&](int i, int j) -> cd& { return H[i + n * j]; };
  // reduce to Hessenberg by Householder-free Gaussian similarity is unnecessary: Arnoldi's H already is.
  // zero anything below the first subdiagonal (round-off only)
  for (int j = 0; j < n; ++j)
    for (int i = j + 2; i < n; ++i) at(i, j) = 0.0;
  double rho = 0.0;
  int hi = n - 1;
  int iter = 0;
  while (hi >= 0) {
    if (hi == 0) {
      rho = std::max(rho, std::abs(at(0, 0)));
      break;
    }
    // deflation check
    int l = hi;
    while (l > 0) {
      const double s = std::abs(at(l - 1, l - 1)) + std::abs(at(l, l));
      if (std::abs(at(l, l - 1)) <= 1e-15 * (s == 0.0 ? 1.0 : s)) {
        at(l, l - 1) = 0.0;
        break;
      }
      --l;
    }
    if (l == hi) {
      rho = std::max(rho, std::abs(at(hi, hi)));
      --hi;
      iter = 0;
      continue;
    }
    if (++iter > 500) {  // give up: fall back to a norm bound
      double nb = 0.0;
      for (int i = 0; i <= hi; ++i) {
        double s = 0.0;
        for (int j = 0; j <= hi; ++j) s += std::abs(at(i, j));
        nb = std::max(nb, s);
      }
      return std::max(rho, nb);
    }
    // Wilkinson shift from the trailing 2x2 of the active block [l..hi]
    const cd a = at(hi - 1, hi - 1), b = at(hi - 1, hi), c = at(hi, hi - 1), d = at(hi, hi);
    const cd tr = a + d, det = a * d - b * c;
    const cd disc = std::sqrt(tr * tr - 4.0 * det);
    cd mu1 = 0.5 * (tr + disc), mu2 = 0.5 * (tr - disc);
    cd mu = (std::abs(mu1 - d) < std::abs(mu2 - d)) ? mu1 : mu2;
    if (iter % 11 == 10) mu += cd(std::abs(at(hi, hi - 1)), 0.0);  // exceptional shift
    // QR step on block [l..hi] with Givens rotations
    const int m = hi - l + 1;
    std::vector<cd> cs(m), sn(m);
    for (int i = l; i <= hi; ++i) at(i, i) -= mu;
    for (int k = l; k < hi; ++k) {
      const cd x = at(k, k), y = at(k + 1, k);
      const double r = std::sqrt(std::norm(x) + std::norm(y));
      cd c_ = 1.0, s_ = 0.0;
      if (r > 0.0) {
        c_ = x / r;
        s_ = y / r;
      }
      cs[k - l] = c_;
      sn[k - l] = s_;
      for (int j = k; j <= hi; ++j) {
        const cd t1 = at(k, j), t2 = at(k + 1, j);
        at(k, j) = std::conj(c_) * t1 + std::conj(s_) * t2;
        at(k + 1, j) = -s_ * t1 + c_ * t2;
      }
    }
    for (int k = l; k < hi; ++k) {
      const cd c_ = cs[k - l], s_ = sn[k - l];
      for (int i = l; i <= std::min(k + 2, hi); ++i) {
        const cd t1 = at(i, k), t2 = at(i, k + 1);
        at(i, k) = t1 * c_ + t2 * s_;
        at(i, k + 1) = -t1 * std::conj(s_) + t2 * std::conj(c_);
      }
    }
    for (int i = l; i <= hi; ++i) at(i, i) += mu;
  }
  return rho;
}

}  // namespace nrsb
