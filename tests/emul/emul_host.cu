// CPU emulation harness for tests/ ONLY: runs the SAME per-cell arithmetic the
// CUDA kernels run (the __host__ __device__ functions of
// sayram2d_b200/csrc/sy2d_kernels.cuh) in plain serial loops, so that the
// device code's formulas can be checked against the reference's (M,R) and f on a
// box without a GPU.  It is not part of the product and nothing under
// sayram2d_b200/ loads it; the product has no CPU path.
#include <cmath>
#include <cstring>
#include <vector>

#include "../../sayram2d_b200/csrc/sy2d_geometry.h"
#include "../../sayram2d_b200/csrc/sy2d_kernels.cuh"

using namespace sy2d;

namespace {
struct Emul {
  int nx, ny;
  double dt;
  HostGeometry hg;
  std::vector<double> tx, ty, cxy, U, Ud, bc[4];
  int bct[4];
  Geometry geo() const {
    Geometry g;
    g.wxL = hg.wxL.data(); g.wxR = hg.wxR.data(); g.wyB = hg.wyB.data(); g.wyT = hg.wyT.data();
    g.bc_xmin = bc[0].data(); g.bc_xmax = bc[1].data(); g.bc_ymin = bc[2].data(); g.bc_ymax = bc[3].data();
    for (int k = 0; k < 4; ++k) g.bc[k] = bct[k];
    g.nx = nx; g.ny = ny;
    return g;
  }
};
}  // namespace

extern "C" {

void* emul_create(int nx, int ny, const double* xe, const double* ye, double dt, const double* G, const double* Dxx,
                  const double* Dxy, const double* Dyy, const double* inv_tau, const int* bct, const double* xmin,
                  const double* xmax, const double* ymin, const double* ymax) {
  Emul* e = new Emul;
  e->nx = nx; e->ny = ny; e->dt = dt;
  e->hg = make_host_geometry(nx, ny, xe, ye);
  const size_t N = (size_t)nx * ny;
  e->tx.resize(N); e->ty.resize(N); e->cxy.resize(N); e->U.resize(N); e->Ud.resize(N);
  for (int i = 0; i < nx; ++i)
    for (int j = 0; j < ny; ++j) {
      const size_t n = (size_t)i * ny + j;
      const CellCoeffs c = prepare_cell(G[n], Dxx[n], Dxy[n], Dyy[n], inv_tau ? inv_tau[n] : 0.0, e->hg.dx[i], e->hg.dy[j], dt);
      e->tx[n] = c.tx; e->ty[n] = c.ty; e->cxy[n] = c.cxy; e->U[n] = c.U; e->Ud[n] = c.Ud;
    }
  const double* lines[4] = {xmin, xmax, ymin, ymax};
  for (int k = 0; k < 4; ++k) {
    e->bct[k] = bct[k];
    const size_t m = (k < 2 ? ny : nx) + 1;
    e->bc[k].assign(m, 0.0);
    if (lines[k]) std::memcpy(e->bc[k].data(), lines[k], m * sizeof(double));
  }
  return e;
}

void emul_destroy(void* h) { delete static_cast<Emul*>(h); }

// unscaled operator: diags [5][nx][ny] (diag, W, E, S, N), R [nx][ny], vf [(nx+1)][(ny+1)]
void emul_assemble(void* h, const double* f, double* diags, double* R, double* vf) {
  Emul* e = static_cast<Emul*>(h);
  const Geometry g = e->geo();
  const int nx = e->nx, ny = e->ny;
  const size_t N = (size_t)nx * ny;
  for (int i = 0; i < nx; ++i)
    for (int j = 0; j < ny; ++j) {
      Row r;
      assemble_row(f, e->tx.data(), e->ty.data(), e->cxy.data(), e->U.data(), e->Ud.data(), g, i, j, r);
      const size_t n = (size_t)i * ny + j;
      diags[n] = r.diag; diags[N + n] = r.oW; diags[2 * N + n] = r.oE; diags[3 * N + n] = r.oS; diags[4 * N + n] = r.oN;
      R[n] = r.R;
      if (vf) {
        vf[(size_t)i * (ny + 1) + j] = r.vSW;
        vf[(size_t)(i + 1) * (ny + 1) + j] = r.vSE;
        vf[(size_t)i * (ny + 1) + j + 1] = r.vNW;
        vf[(size_t)(i + 1) * (ny + 1) + j + 1] = r.vNE;
      }
    }
}

// One implicit step with the same formulation and iteration as the CUDA path
// (scaled system A d = rhs, BiCGSTAB, max-norm stop); returns iterations.
int emul_step(void* h, double* f, double* yprev, double tol, int maxit, int predictor, double* resid_out) {
  Emul* e = static_cast<Emul*>(h);
  const Geometry g = e->geo();
  const int nx = e->nx, ny = e->ny;
  const size_t N = (size_t)nx * ny;
  std::vector<double> wW(N), wE(N), wS(N), wN(N), rhs(N), cs(N), x(N, 0.0), r(N), p(N), v(N), s(N), t(N);
  double rho = 0.0, rmax = 0.0;
  for (int i = 0; i < nx; ++i)
    for (int j = 0; j < ny; ++j) {
      Row row;
      assemble_row(f, e->tx.data(), e->ty.data(), e->cxy.data(), e->U.data(), e->Ud.data(), g, i, j, row);
      const size_t n = (size_t)i * ny + j;
      const size_t nW = i > 0 ? n - ny : n, nE = i < nx - 1 ? n + ny : n, nS = j > 0 ? n - 1 : n, nN = j < ny - 1 ? n + 1 : n;
      Scaled sc;
      scale_row(row, yprev[n], yprev[nW], yprev[nE], yprev[nS], yprev[nN], sc);
      wW[n] = sc.wW; wE[n] = sc.wE; wS[n] = sc.wS; wN[n] = sc.wN; rhs[n] = sc.rhs; cs[n] = sc.cs;
      rho += sc.rhs * sc.rhs;
      rmax = std::fmax(rmax, std::fabs(sc.rhs));
    }
  int it = 0;
  double alpha = 1.0, omega = 1.0, beta = 0.0;
  bool first = true;
  while (!(rmax <= tol) && it < maxit) {
    for (size_t n = 0; n < N; ++n) p[n] = first ? rhs[n] : r[n] + beta * (p[n] - omega * v[n]);
    double rv = 0.0;
    for (size_t n = 0; n < N; ++n) {
      v[n] = stencil_apply(p.data(), n, N, ny, p[n], wW[n], wE[n], wS[n], wN[n]);
      rv += rhs[n] * v[n];
    }
    alpha = rv != 0.0 ? rho / rv : 0.0;
    for (size_t n = 0; n < N; ++n) s[n] = (first ? rhs[n] : r[n]) - alpha * v[n];
    double ts = 0.0, tt = 0.0;
    for (size_t n = 0; n < N; ++n) {
      t[n] = stencil_apply(s.data(), n, N, ny, s[n], wW[n], wE[n], wS[n], wN[n]);
      ts += t[n] * s[n];
      tt += t[n] * t[n];
    }
    omega = tt > 0.0 ? ts / tt : 0.0;
    double rho_new = 0.0;
    rmax = 0.0;
    for (size_t n = 0; n < N; ++n) {
      x[n] = (first ? 0.0 : x[n]) + (alpha * p[n] + omega * s[n]);
      r[n] = s[n] - omega * t[n];
      rho_new += rhs[n] * r[n];
      rmax = std::fmax(rmax, std::fabs(r[n]));
    }
    beta = (rho_new / rho) * (alpha / omega);
    rho = rho_new;
    first = false;
    ++it;
  }
  double res = 0.0;
  for (size_t n = 0; n < N; ++n) {
    const double ax = it > 0 ? stencil_apply(x.data(), n, N, ny, x[n], wW[n], wE[n], wS[n], wN[n]) : 0.0;
    res = std::fmax(res, std::fabs(rhs[n] - ax));
  }
  if (resid_out) *resid_out = res;
  for (size_t n = 0; n < N; ++n) {
    const double fold = f[n], fnew = cs[n] * (1.0 + (it > 0 ? x[n] : 0.0));
    f[n] = fnew;
    if (predictor) {
      double y = fnew / fold;
      y = std::fmin(std::fmax(y, kPredMin), kPredMax);
      yprev[n] = (y == y) ? y : 1.0;
    }
  }
  return rmax <= tol ? it : -it;
}

// Same step with the x-line (tridiagonal along i) right preconditioner, as engine 2's
// k_problem_xline runs it:  T = tridiag(wW, 1, wE) per column j,  phat = T^-1 p,
// v = A phat = p + wS phat_S + wN phat_N  (because T phat = p),  likewise for s.
int emul_step_xline(void* h, double* f, double* yprev, double tol, int maxit, int predictor, double* resid_out) {
  Emul* e = static_cast<Emul*>(h);
  const Geometry g = e->geo();
  const int nx = e->nx, ny = e->ny;
  const size_t N = (size_t)nx * ny;
  std::vector<double> wW(N), wE(N), wS(N), wN(N), rhs(N), cs(N), x(N, 0.0), r(N), p(N), v(N), t(N), hat(N), l(N), dinv(N), ee(N);
  double rho = 0.0, rmax = 0.0;
  for (int i = 0; i < nx; ++i)
    for (int j = 0; j < ny; ++j) {
      Row row;
      assemble_row(f, e->tx.data(), e->ty.data(), e->cxy.data(), e->U.data(), e->Ud.data(), g, i, j, row);
      const size_t n = (size_t)i * ny + j;
      const size_t nW = i > 0 ? n - ny : n, nE = i < nx - 1 ? n + ny : n, nS = j > 0 ? n - 1 : n, nN = j < ny - 1 ? n + 1 : n;
      Scaled sc;
      scale_row(row, yprev[n], yprev[nW], yprev[nE], yprev[nS], yprev[nN], sc);
      wW[n] = sc.wW; wE[n] = sc.wE; wS[n] = sc.wS; wN[n] = sc.wN; rhs[n] = sc.rhs; cs[n] = sc.cs;
      rho += sc.rhs * sc.rhs;
      rmax = std::fmax(rmax, std::fabs(sc.rhs));
    }
  for (int j = 0; j < ny; ++j) {  // LU of T per column: d_0 = 1, l_i = wW_i/d_{i-1}, d_i = 1 - l_i wE_{i-1}
    double dprev = 1.0;
    for (int i = 0; i < nx; ++i) {
      const size_t n = (size_t)i * ny + j;
      const XlineFactor fc = xline_factor(wW[n], i > 0 ? wE[n - ny] : 0.0, dprev, i == 0);
      l[n] = fc.l; dinv[n] = fc.dinv; ee[n] = wE[n] * fc.dinv;
      dprev = fc.d;
    }
  }
  auto tsolve = [&](const std::vector<double>& b, std::vector<double>& y) {
    for (int j = 0; j < ny; ++j) {
      double carry = 0.0;
      for (int i = 0; i < nx; ++i) { const size_t n = (size_t)i * ny + j; carry = b[n] - l[n] * carry; y[n] = carry; }
      carry = 0.0;
      for (int i = nx - 1; i >= 0; --i) { const size_t n = (size_t)i * ny + j; carry = y[n] * dinv[n] - ee[n] * carry; y[n] = carry; }
    }
  };
  auto sn_apply = [&](const std::vector<double>& c, const std::vector<double>& yh, size_t n) {
    const double yS = yh[n > 0 ? n - 1 : n], yN = yh[n + 1 < N ? n + 1 : n];
    return c[n] + (wS[n] * yS + wN[n] * yN);
  };
  int it = 0;
  double alpha = 1.0, omega = 1.0, beta = 0.0;
  bool first = true;
  r = rhs;
  while (!(rmax <= tol) && it < maxit) {
    for (size_t n = 0; n < N; ++n) p[n] = first ? r[n] : r[n] + beta * (p[n] - omega * v[n]);
    tsolve(p, hat);
    double rv = 0.0;
    for (size_t n = 0; n < N; ++n) { v[n] = sn_apply(p, hat, n); rv += rhs[n] * v[n]; }
    alpha = rv != 0.0 ? rho / rv : 0.0;
    for (size_t n = 0; n < N; ++n) { r[n] = r[n] - alpha * v[n]; x[n] += alpha * hat[n]; }  // r now holds s
    tsolve(r, hat);
    double ts = 0.0, tt = 0.0;
    for (size_t n = 0; n < N; ++n) { t[n] = sn_apply(r, hat, n); ts += t[n] * r[n]; tt += t[n] * t[n]; }
    omega = tt > 0.0 ? ts / tt : 0.0;
    double rho_new = 0.0;
    rmax = 0.0;
    for (size_t n = 0; n < N; ++n) {
      x[n] += omega * hat[n];
      r[n] = r[n] - omega * t[n];
      rho_new += rhs[n] * r[n];
      rmax = std::fmax(rmax, std::fabs(r[n]));
    }
    beta = (rho_new / rho) * (alpha / omega);
    rho = rho_new;
    first = false;
    ++it;
  }
  double res = 0.0;  // true residual with the FULL operator (all four neighbours)
  for (size_t n = 0; n < N; ++n) {
    const double ax = it > 0 ? stencil_apply(x.data(), n, N, ny, x[n], wW[n], wE[n], wS[n], wN[n]) : 0.0;
    res = std::fmax(res, std::fabs(rhs[n] - ax));
  }
  if (resid_out) *resid_out = res;
  for (size_t n = 0; n < N; ++n) {
    const double fold = f[n], fnew = cs[n] * (1.0 + (it > 0 ? x[n] : 0.0));
    f[n] = fnew;
    if (predictor) {
      double y = fnew / fold;
      y = std::fmin(std::fmax(y, kPredMin), kPredMax);
      yprev[n] = (y == y) ? y : 1.0;
    }
  }
  return rmax <= tol ? it : -it;
}

}  // extern "C"
