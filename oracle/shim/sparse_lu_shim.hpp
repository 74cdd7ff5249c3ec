// SparseLU look-alike for the reference's call sites (Solver.h:35,
// Solver.cc:53-54 analyzePattern, :277 factorize, :278 solve).
//
// TEST INFRASTRUCTURE.  Direct LU without pivoting (valid because the PPFV
// matrix is strictly column diagonally dominant).  Two kernels:
//   * (default) left-looking sparse LU on a nested-dissection ordering of the
//     detected nx*ny grid graph, symbolic pattern computed once in analyzePattern
//     (mirrors Eigen's analyzePattern/factorize split);
//   * dense band (natural numbering K=j*nx+i, half bandwidth nx), selected with
//     SY2D_ORACLE_LU=band, as an independent cross-check at small N.
// Also keeps a pointer to the last factorised matrix / RHS so the oracle driver
// can dump (M, R) of a step without touching the reference's private members.
#pragma once
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <numeric>

namespace Eigen {
namespace shim {
struct LastSystem {
  const SparseMatrix<double, ColMajor>* M = nullptr;
  std::vector<double> R;
  double factor_seconds = 0.0, solve_seconds = 0.0;
  long nnz_LU = 0;
};
inline LastSystem& last_system() {
  static LastSystem s;
  return s;
}
inline double now() {
  return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}
}  // namespace shim

template <class Mat, class Ordering>
class SparseLU {
 public:
  void analyzePattern(const Mat& A) {
    n_ = A.rows();
    bw_ = 0;
    for (long c = 0; c < n_; ++c)
      for (long p = A.colptr[c]; p < A.colptr[c + 1]; ++p) bw_ = std::max<long>(bw_, std::labs((long)A.rowind[p] - c));
    // nested dissection is ~3x faster than the band already at 80x80; the band
    // kernel stays as an independent cross-check (SY2D_ORACLE_LU=band).
    const char* force = std::getenv("SY2D_ORACLE_LU");  // "band" | "nd"
    use_band_ = force ? (std::string(force) == "band") : false;
    if (!use_band_) nd_symbolic_(A);
  }

  void factorize(const Mat& A) {
    double t0 = shim::now();
    shim::last_system().M = &A;
    if (use_band_)
      band_factor_(A);
    else
      nd_numeric_(A);
    shim::last_system().factor_seconds += shim::now() - t0;
  }

  VectorXd solve(const VectorXd& b) const {
    double t0 = shim::now();
    shim::last_system().R = b.d;
    VectorXd x(b);
    if (use_band_)
      band_solve_(x.d);
    else
      nd_solve_(x.d);
    shim::last_system().solve_seconds += shim::now() - t0;
    return x;
  }

  int info() const { return 0; }

 private:
  long n_ = 0, bw_ = 0;
  bool use_band_ = true;

  // ------------------------------------------------------------- dense band
  std::vector<double> ab_;  // row r holds columns r-bw..r+bw at ab_[r*(2bw+1) + (c-r+bw)]
  double& B_(long r, long c) { return ab_[r * (2 * bw_ + 1) + (c - r + bw_)]; }
  double B_(long r, long c) const { return ab_[r * (2 * bw_ + 1) + (c - r + bw_)]; }

  void band_factor_(const Mat& A) {
    const long w = 2 * bw_ + 1;
    ab_.assign((std::size_t)n_ * w, 0.0);
    for (long c = 0; c < n_; ++c)
      for (long p = A.colptr[c]; p < A.colptr[c + 1]; ++p) B_(A.rowind[p], c) = A.val[p];
    for (long k = 0; k < n_; ++k) {
      const double piv = B_(k, k);
      const long last = std::min(n_ - 1, k + bw_);
      const double* rk = &ab_[k * w + bw_];  // rk[c-k]
      for (long i = k + 1; i <= last; ++i) {
        double l = B_(i, k);
        if (l == 0.0) continue;
        l /= piv;
        B_(i, k) = l;
        double* ri = &ab_[i * w + (k - i + bw_)];  // ri[c-k]
        const long len = last - k;
        for (long q = 1; q <= len; ++q) ri[q] -= l * rk[q];
      }
    }
    shim::last_system().nnz_LU = (long)ab_.size();
  }

  void band_solve_(std::vector<double>& x) const {
    for (long i = 0; i < n_; ++i) {
      double s = x[i];
      for (long c = std::max<long>(0, i - bw_); c < i; ++c) s -= B_(i, c) * x[c];
      x[i] = s;
    }
    for (long i = n_ - 1; i >= 0; --i) {
      double s = x[i];
      const long last = std::min(n_ - 1, i + bw_);
      for (long c = i + 1; c <= last; ++c) s -= B_(i, c) * x[c];
      x[i] = s / B_(i, i);
    }
  }

  // --------------------------------------------- nested-dissection sparse LU
  // perm_[new] = old.  L is unit lower, U upper, both CSC in the NEW numbering,
  // row indices ascending inside each column (ascending = a valid topological
  // order for the column-k triangular solve).
  std::vector<int> perm_, iperm_;
  std::vector<long> Lp_, Up_;
  std::vector<int> Li_, Ui_;
  std::vector<double> Lx_, Ux_;
  std::vector<long> Ap_;  // permuted A pattern (CSC, new numbering) -> index into A.val
  std::vector<int> Ai_;
  std::vector<long> Asrc_;

  static void nd_order_(int x0, int x1, int y0, int y1, int nx, std::vector<int>& out) {
    const int w = x1 - x0, h = y1 - y0;
    if (w <= 0 || h <= 0) return;
    if (w * h <= 24 || (w <= 2 && h <= 2)) {
      for (int j = y0; j < y1; ++j)
        for (int i = x0; i < x1; ++i) out.push_back(j * nx + i);
      return;
    }
    if (w >= h) {
      const int xm = x0 + w / 2;
      nd_order_(x0, xm, y0, y1, nx, out);
      nd_order_(xm + 1, x1, y0, y1, nx, out);
      for (int j = y0; j < y1; ++j) out.push_back(j * nx + xm);
    } else {
      const int ym = y0 + h / 2;
      nd_order_(x0, x1, y0, ym, nx, out);
      nd_order_(x0, x1, ym + 1, y1, nx, out);
      for (int i = x0; i < x1; ++i) out.push_back(ym * nx + i);
    }
  }

  void nd_symbolic_(const Mat& A) {
    const long n = n_;
    const int nx = (int)bw_;
    perm_.clear();
    if (nx > 0 && n % nx == 0) nd_order_(0, nx, 0, (int)(n / nx), nx, perm_);
    if ((long)perm_.size() != n) {  // not a grid graph: natural order
      perm_.resize(n);
      std::iota(perm_.begin(), perm_.end(), 0);
    }
    iperm_.assign(n, 0);
    for (long k = 0; k < n; ++k) iperm_[perm_[k]] = (int)k;
    // permuted pattern B = P A P^T, columns in new order, rows ascending
    Ap_.assign(n + 1, 0);
    Ai_.clear();
    Asrc_.clear();
    std::vector<std::pair<int, long>> tmp;
    for (long k = 0; k < n; ++k) {
      const long c = perm_[k];
      tmp.clear();
      for (long p = A.colptr[c]; p < A.colptr[c + 1]; ++p) tmp.emplace_back(iperm_[A.rowind[p]], p);
      std::sort(tmp.begin(), tmp.end());
      for (auto& t : tmp) {
        Ai_.push_back(t.first);
        Asrc_.push_back(t.second);
      }
      Ap_[k + 1] = (long)Ai_.size();
    }
    // Symbolic: the pattern is structurally symmetric and no pivoting happens, so
    // struct(L) = struct(U^T) = symbolic Cholesky of B.  Row-merge via the
    // elimination tree: struct(L(:,k)) = struct(B(k+1:,k)) U (struct(L(:,c)) \ {k}) for children c.
    std::vector<std::vector<int>> Ls(n);
    std::vector<int> parent(n, -1), mark(n, -1);
    std::vector<std::vector<int>> children(n);
    for (long k = 0; k < n; ++k) {
      std::vector<int>& col = Ls[k];
      mark[k] = (int)k;
      for (long p = Ap_[k]; p < Ap_[k + 1]; ++p) {
        int i = Ai_[p];
        if (i > k && mark[i] != k) {
          mark[i] = (int)k;
          col.push_back(i);
        }
      }
      for (int c : children[k])
        for (int i : Ls[c])
          if (i > k && mark[i] != k) {
            mark[i] = (int)k;
            col.push_back(i);
          }
      std::sort(col.begin(), col.end());
      if (!col.empty()) {
        parent[k] = col[0];
        children[col[0]].push_back((int)k);
      }
      for (int c : children[k]) {
        // children structures are no longer needed once merged into the parent
        (void)c;
      }
    }
    Lp_.assign(n + 1, 0);
    for (long k = 0; k < n; ++k) Lp_[k + 1] = Lp_[k] + (long)Ls[k].size();
    Li_.resize(Lp_[n]);
    for (long k = 0; k < n; ++k) std::copy(Ls[k].begin(), Ls[k].end(), Li_.begin() + Lp_[k]);
    // U(:,k) pattern = rows j<k with L(k,j) != 0  (transpose of L's pattern), plus the diagonal last
    std::vector<long> cnt(n + 1, 0);
    for (long j = 0; j < n; ++j)
      for (long p = Lp_[j]; p < Lp_[j + 1]; ++p) cnt[Li_[p] + 1]++;
    Up_.assign(n + 1, 0);
    for (long k = 0; k < n; ++k) Up_[k + 1] = Up_[k] + cnt[k + 1] + 1;
    Ui_.resize(Up_[n]);
    std::vector<long> pos(Up_.begin(), Up_.end() - 1);
    for (long j = 0; j < n; ++j)  // ascending j => ascending rows inside each U column
      for (long p = Lp_[j]; p < Lp_[j + 1]; ++p) Ui_[pos[Li_[p]]++] = (int)j;
    for (long k = 0; k < n; ++k) Ui_[pos[k]++] = (int)k;
    Lx_.assign(Lp_[n], 0.0);
    Ux_.assign(Up_[n], 0.0);
    shim::last_system().nnz_LU = Lp_[n] + Up_[n];
  }

  void nd_numeric_(const Mat& A) {
    const long n = n_;
    std::vector<double> x(n, 0.0);
    for (long k = 0; k < n; ++k) {
      for (long p = Ap_[k]; p < Ap_[k + 1]; ++p) x[Ai_[p]] = A.val[Asrc_[p]];
      // x = L \ B(:,k) restricted to the known pattern; rows ascending
      const long ub = Up_[k], ue = Up_[k + 1] - 1;  // last entry is the diagonal
      for (long q = ub; q < ue; ++q) {
        const int j = Ui_[q];
        const double xj = x[j];
        Ux_[q] = xj;
        x[j] = 0.0;
        if (xj != 0.0) {
          const long le = Lp_[j + 1];
          for (long p = Lp_[j]; p < le; ++p) x[Li_[p]] -= Lx_[p] * xj;
        }
      }
      const double piv = x[k];
      Ux_[ue] = piv;
      x[k] = 0.0;
      const double ipiv = 1.0 / piv;
      for (long p = Lp_[k]; p < Lp_[k + 1]; ++p) {
        const int i = Li_[p];
        Lx_[p] = x[i] * ipiv;
        x[i] = 0.0;
      }
    }
  }

  void nd_solve_(std::vector<double>& b) const {
    const long n = n_;
    std::vector<double> y(n);
    for (long k = 0; k < n; ++k) y[k] = b[perm_[k]];
    for (long j = 0; j < n; ++j) {
      const double yj = y[j];
      if (yj != 0.0)
        for (long p = Lp_[j]; p < Lp_[j + 1]; ++p) y[Li_[p]] -= Lx_[p] * yj;
    }
    for (long k = n - 1; k >= 0; --k) {
      const long ue = Up_[k + 1] - 1;
      const double xk = y[k] / Ux_[ue];
      y[k] = xk;
      if (xk != 0.0)
        for (long q = Up_[k]; q < ue; ++q) y[Ui_[q]] -= Ux_[q] * xk;
    }
    for (long k = 0; k < n; ++k) b[perm_[k]] = y[k];
  }
};

}  // namespace Eigen
