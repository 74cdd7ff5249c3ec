// Oracle driver around the UNMODIFIED reference classes (Parameters, Mesh,
// Albert_Young, Albert_Young_LC, Solver), which oracle/Makefile compiles from
// /root/reference/source/*.cc against the shim headers in oracle/shim/.
//
// TEST INFRASTRUCTURE: produces the golden f snapshots and (M, R) dumps that pin
// the NumPy restatement (oracle/ppfv_oracle.py) and the CUDA path, and is the
// "reference" CPU baseline that bench.py times.  It follows main.cc:41-85 (same
// construction order, same time loop) but writes .npy files and per-step timing.
//
//   ref_driver --case AY|LC|SYN|ENS|TD --ini <file> --out <dir> [--steps N]
//              [--every K] [--skip W] [--dump-op s1,s2,...] [--member a b] [--stretch s] [--cwd <dir>]
//
// SYN  = BASELINE config 3 (synthetic full tensor + loss strip on the AY domain);
// ENS  = BASELINE config 4 member: LC case with D scaled by a, inv_tau by b.
// TD   = time-dependent D, 1/tau and Dirichlet data (Equation::update(t) path).
#include <unistd.h>

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <ctime>
#include <set>
#include <sstream>
#include <string>

#include "Albert_Young.h"
#include "Albert_Young_LC.h"
#include "Mesh.h"
#include "Parameters.h"
#include "Solver.h"

namespace {

// A user-defined case plugged into the reference's Equation interface
// (README.md:148-168 describes this extension point).
class Synthetic_Tensor : public Albert_Young {
 public:
  Synthetic_Tensor(const Parameters& p, const Mesh& m) : Albert_Young(p, m) {
    const double x0 = m.x_edge(0), x1 = m.x_edge(m.nx());
    const double y0 = m.y_edge(0), y1 = m.y_edge(m.ny());
    for (std::size_t i = 0; i < m.nx(); ++i) {
      const double xi = (m.x(i) - x0) / (x1 - x0);
      for (std::size_t j = 0; j < m.ny(); ++j) {
        const double eta = (m.y(j) - y0) / (y1 - y0);
        const double s = std::sin(gPI * xi);
        Dxx_(i, j) = 10.0 * std::exp(-3.0 * eta) * (0.05 + s * s);
        Dyy_(i, j) = 2.0 * std::exp(-2.0 * eta) * (0.05 + 4.0 * xi * (1.0 - xi));
        const double rho = 0.8 * std::sin(2.0 * gPI * xi) * std::cos(gPI * eta);
        Dxy_(i, j) = rho * std::sqrt(Dxx_(i, j) * Dyy_(i, j));
        inv_tau_(i, j) = 5.0 * std::max(0.0, 1.0 - xi / 0.1);
      }
    }
  }
};

class Ensemble_Member : public Albert_Young_LC {
 public:
  Ensemble_Member(const Parameters& p, const Mesh& m, double a, double b) : Albert_Young_LC(p, m) {
    for (std::size_t i = 0; i < m.nx(); ++i)
      for (std::size_t j = 0; j < m.ny(); ++j) {
        Dxx_(i, j) *= a;
        Dxy_(i, j) *= a;
        Dyy_(i, j) *= a;
        inv_tau_(i, j) *= b;
      }
  }
};

// A time-dependent user case (README.md:177: "update(t) is where G, D and the BCs are updated"): the
// AY case with D(t) = D0 (1 + 0.5 sin(2 pi t / 0.1)), a loss term 1/tau(t) = 3 sin^2(pi t / 0.05) day^-1 on
// the first quarter of the alpha0 rows, and the low-energy Dirichlet line decaying like exp(-2 t).
// Exercises Solver.cc:286-289 (eq.update(t), update_Lambda, update_vertex_f after every step).
class Time_Dependent : public Albert_Young {
 public:
  Time_Dependent(const Parameters& p, const Mesh& m) : Albert_Young(p, m), m_(m), Dxx0_(Dxx_), Dxy0_(Dxy_), Dyy0_(Dyy_) { update(0.0); }
  void update(double t) override {
    const double a = 1.0 + 0.5 * std::sin(2.0 * gPI * t / 0.1);
    const double s = std::sin(gPI * t / 0.05);
    for (std::size_t i = 0; i < m_.nx(); ++i)
      for (std::size_t j = 0; j < m_.ny(); ++j) {
        Dxx_(i, j) = a * Dxx0_(i, j);
        Dxy_(i, j) = a * Dxy0_(i, j);
        Dyy_(i, j) = a * Dyy0_(i, j);
        inv_tau_(i, j) = i < m_.nx() / 4 ? 3.0 * s * s : 0.0;
      }
  }
  bool dirichlet_vertex_value(BoundaryID side, std::size_t i, std::size_t j, double t, double* out) const override {
    const bool ok = Albert_Young::dirichlet_vertex_value(side, i, j, t, out);
    if (ok && side == BoundaryID::YMIN) *out *= std::exp(-2.0 * t);
    return ok;
  }

 private:
  const Mesh& m_;
  Xtensor2d Dxx0_, Dxy0_, Dyy0_;
};

// main.cc:20-37; stretch != 0 warps the edges (xi -> xi + stretch*sin(2 pi xi)/(2 pi))
// to exercise the non-uniform weights of Solver.cc:326-381, which make_uniform never does.
Grid2D make_grid(const Parameters& p, double stretch) {
  std::vector<double> xe(p.nalpha0() + 1), ye(p.nE() + 1);
  const double dx = (p.alpha0_max() - p.alpha0_min()) / static_cast<double>(p.nalpha0());
  const double dy = (p.logEmax() - p.logEmin()) / static_cast<double>(p.nE());
  for (std::size_t i = 0; i <= p.nalpha0(); ++i) xe[i] = p.alpha0_min() + dx * static_cast<double>(i);
  for (std::size_t j = 0; j <= p.nE(); ++j) ye[j] = p.logEmin() + dy * static_cast<double>(j);
  if (stretch != 0.0) {
    auto warp = [&](std::vector<double>& e, double s) {
      const double a = e.front(), L = e.back() - e.front();
      const std::size_t n = e.size() - 1;
      for (std::size_t k = 1; k < n; ++k) {
        const double xi = static_cast<double>(k) / static_cast<double>(n);
        e[k] = a + L * (xi + s * std::sin(2.0 * gPI * xi) / (2.0 * gPI));
      }
    };
    warp(xe, stretch);
    warp(ye, -0.5 * stretch);
  }
  return Grid2D(std::move(xe), std::move(ye));
}

void dump_field(const std::string& path, const Equation& eq, const Mesh& m, int which) {
  std::vector<double> v(m.nx() * m.ny());
  for (std::size_t i = 0; i < m.nx(); ++i)
    for (std::size_t j = 0; j < m.ny(); ++j) {
      const Ind k{i, j};
      v[i * m.ny() + j] = which == 0 ? eq.G(k) : which == 1 ? eq.Dxx(k) : which == 2 ? eq.Dxy(k) : which == 3 ? eq.Dyy(k) : eq.inv_tau(k);
    }
  h5shim::write_npy(path, v.data(), {m.nx(), m.ny()});
}

// (M, R) of the step that was just solved, as five (nx,ny) diagonals in the
// reference's numbering K=j*nx+i (Mesh.h:66-68), read back from the shim's hook.
void dump_operator(const std::string& prefix, const Mesh& m) {
  const auto& ls = Eigen::shim::last_system();
  const auto& M = *ls.M;
  const long nx = (long)m.nx(), ny = (long)m.ny();
  std::vector<double> d[5];
  for (auto& a : d) a.assign(nx * ny, 0.0);
  for (long c = 0; c < M.cols(); ++c)
    for (long p = M.colptr[c]; p < M.colptr[c + 1]; ++p) {
      const long r = M.rowind[p], i = r % nx, j = r / nx, off = c - r;
      const int k = off == 0 ? 0 : off == -1 ? 1 : off == 1 ? 2 : off == -nx ? 3 : off == nx ? 4 : -1;
      if (k < 0) throw std::runtime_error("unexpected stencil offset");
      d[k][i * ny + j] = M.val[p];
    }
  const char* names[5] = {"diag", "W", "E", "S", "N"};
  for (int k = 0; k < 5; ++k) h5shim::write_npy(prefix + names[k] + ".npy", d[k].data(), {(std::size_t)nx, (std::size_t)ny});
  std::vector<double> R(nx * ny);
  for (long i = 0; i < nx; ++i)
    for (long j = 0; j < ny; ++j) R[i * ny + j] = ls.R[j * nx + i];
  h5shim::write_npy(prefix + "R.npy", R.data(), {(std::size_t)nx, (std::size_t)ny});
}

}  // namespace

int main(int argc, char** argv) {
  std::string kase = "AY", ini = "p.ini", out = "ref_out", cwd;
  long steps = -1, every = -1, skip = 0;
  double ma = 1.0, mb = 1.0, stretch = 0.0;
  std::set<long> dump_ops;
  for (int k = 1; k < argc; ++k) {
    const std::string a = argv[k];
    auto next = [&]() { if (k + 1 >= argc) { std::fprintf(stderr, "missing value for %s\n", a.c_str()); std::exit(2); } return std::string(argv[++k]); };
    if (a == "--case") kase = next();
    else if (a == "--ini") ini = next();
    else if (a == "--out") out = next();
    else if (a == "--cwd") cwd = next();
    else if (a == "--steps") steps = std::stol(next());
    else if (a == "--every") every = std::stol(next());
    else if (a == "--skip") skip = std::stol(next());  // warm-up steps excluded from the timings
    else if (a == "--stretch") stretch = std::stod(next());
    else if (a == "--member") { ma = std::stod(next()); mb = std::stod(next()); }
    else if (a == "--dump-op") { std::stringstream ss(next()); std::string t; while (std::getline(ss, t, ',')) dump_ops.insert(std::stol(t)); }
    else { std::fprintf(stderr, "unknown argument %s\n", a.c_str()); return 2; }
  }
  if (!cwd.empty() && chdir(cwd.c_str()) != 0) { std::perror("chdir"); return 2; }

  char* pargv[2] = {argv[0], const_cast<char*>(ini.c_str())};
  Parameters paras(2, pargv);
  Grid2D grid = make_grid(paras, stretch);
  Mesh m(grid, paras.dt());

  Equation* eq = nullptr;
  if (kase == "AY") eq = new Albert_Young(paras, m);
  else if (kase == "LC") eq = new Albert_Young_LC(paras, m);
  else if (kase == "SYN") eq = new Synthetic_Tensor(paras, m);
  else if (kase == "ENS") eq = new Ensemble_Member(paras, m, ma, mb);
  else if (kase == "TD") eq = new Time_Dependent(paras, m);
  else { std::fprintf(stderr, "unknown case %s\n", kase.c_str()); return 2; }

  std::filesystem::create_directories(out);
  const char* fields[5] = {"G", "Dxx", "Dxy", "Dyy", "inv_tau"};
  for (int w = 0; w < 5; ++w) dump_field(out + "/" + fields[w] + ".npy", *eq, m, w);
  h5shim::write_npy(out + "/x_edges.npy", grid.x_edges.data(), {grid.x_edges.size()});
  h5shim::write_npy(out + "/y_edges.npy", grid.y_edges.data(), {grid.y_edges.size()});

  const auto t_setup0 = std::chrono::steady_clock::now();
  Solver solver(m, eq);
  const double setup_s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t_setup0).count();
  h5shim::write_npy(out + "/f_0.npy", solver.f().data(), {m.nx(), m.ny()});

  const long nsteps = steps >= 0 ? steps : paras.nsteps();
  const long save_every = every > 0 ? every : (steps >= 0 ? 1 : paras.save_every_step());
  Eigen::shim::last_system().factor_seconds = Eigen::shim::last_system().solve_seconds = 0.0;
  double loop_s = 0.0;
  const clock_t c0 = clock();
  for (long tstep = 1; tstep <= nsteps; ++tstep) {
    const auto t0 = std::chrono::steady_clock::now();
    solver.update();  // main.cc:80
    if (tstep > skip) loop_s += std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    if (tstep == skip) Eigen::shim::last_system().factor_seconds = Eigen::shim::last_system().solve_seconds = 0.0;
    if (dump_ops.count(tstep)) dump_operator(out + "/op" + std::to_string(tstep) + "_", m);
    if (tstep % save_every == 0) h5shim::write_npy(out + "/f_" + std::to_string(tstep / save_every) + ".npy", solver.f().data(), {m.nx(), m.ny()});
  }
  const double cpu_s = double(clock() - c0) / CLOCKS_PER_SEC;  // main.cc:91-93 measures this (incl. dumps)
  long neg = 0;
  double fmin = 1e300, fmax = -1e300;
  for (std::size_t i = 0; i < m.nx(); ++i)
    for (std::size_t j = 0; j < m.ny(); ++j) {
      const double v = solver.f({i, j});
      neg += v < 0.0;
      fmin = std::min(fmin, v);
      fmax = std::max(fmax, v);
    }
  const auto& ls = Eigen::shim::last_system();
  std::printf(
      "{\"case\": \"%s\", \"nx\": %zu, \"ny\": %zu, \"dt\": %.17g, \"steps\": %ld, \"timed_steps\": %ld, \"save_every\": %ld, \"setup_s\": %.6f, "
      "\"loop_wall_s\": %.6f, \"loop_cpu_clock_s\": %.6f, \"lu_factor_s\": %.6f, \"lu_solve_s\": %.6f, \"nnz_LU\": %ld, "
      "\"t_end\": %.17g, \"fmin\": %.17g, \"fmax\": %.17g, \"negatives\": %ld}\n",
      kase.c_str(), m.nx(), m.ny(), m.dt(), nsteps, nsteps - std::min(skip, nsteps), save_every, setup_s, loop_s, cpu_s, ls.factor_seconds, ls.solve_seconds,
      ls.nnz_LU, solver.t(), fmin, fmax, neg);
  delete eq;
  return 0;
}
