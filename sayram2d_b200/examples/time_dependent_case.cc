// A user-defined TIME-DEPENDENT case on the reference's Equation interface (the extension point the
// reference documents in README.md:148-177: "update(t) is where G, D and the BCs are updated"), run
// with the GPU Solver.  While the GPU advances step n, the drop-in Solver calls eq.update(t^{n+1}), gathers the
// fields and the Dirichlet vertex lines and uploads them into the library's second buffer set, which becomes
// active at step n+1 (Solver.cc:286-289 of the reference: eq.update(t), update_Lambda, update_vertex_f);
// SY2D_SYNC_STAGING=1 selects the blocking route (stage after the step) instead.
//
//   sayram2d_td <ini> <outdir> <nsteps> <every>
//
// The case: Albert & Young with D(t) = D0 (1 + 0.5 sin(2 pi t / 0.1)), a loss term
// 1/tau(t) = 3 sin^2(pi t / 0.05) day^-1 on the first quarter of the alpha0 rows, and the low-energy
// Dirichlet line decaying like exp(-2 t).  The test suite runs the same class on the reference's own
// headers and CPU Solver; tests/golden/td64.npz is what that produced.
#include <cmath>
#include <cstdlib>
#include <iostream>
#include <string>

#include "Albert_Young.h"
#include "Mesh.h"
#include "Parameters.h"
#include "Solver.h"
#include "h5lite.h"

class Time_Dependent : public Albert_Young {
 public:
  Time_Dependent(const Parameters& p, const Mesh& m) : Albert_Young(p, m), m_(m), Dxx0_(Dxx_), Dxy0_(Dxy_), Dyy0_(Dyy_) { update(0.0); }
  bool is_static() const override { return false; }
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

int main(int argc, char** argv) {
  if (argc < 5) { std::cerr << "usage: sayram2d_td <ini> <outdir> <nsteps> <every>" << std::endl; return 2; }
  const std::string out = argv[2];
  const long nsteps = std::atol(argv[3]), every = std::atol(argv[4]);
  try {
    Parameters paras(argv[1]);
    std::vector<double> xe(paras.nalpha0() + 1), ye(paras.nE() + 1);
    const double dx = (paras.alpha0_max() - paras.alpha0_min()) / static_cast<double>(paras.nalpha0());
    const double dy = (paras.logEmax() - paras.logEmin()) / static_cast<double>(paras.nE());
    for (std::size_t i = 0; i <= paras.nalpha0(); ++i) xe[i] = paras.alpha0_min() + dx * static_cast<double>(i);
    for (std::size_t j = 0; j <= paras.nE(); ++j) ye[j] = paras.logEmin() + dy * static_cast<double>(j);
    Grid2D grid(std::move(xe), std::move(ye));
    Mesh m(grid, paras.dt());
    Time_Dependent eq(paras, m);
    Solver solver(m, &eq);
    if (const char* e = std::getenv("SY2D_SYNC_STAGING")) solver.set_async_staging(std::atoi(e) == 0);
    h5lite::write_npy(out + "/f_0.npy", solver.f().data(), {m.nx(), m.ny()});
    long iters = 0;
    for (long k = 1; k <= nsteps; ++k) {
      solver.update();
      iters += solver.iterations_last();
      if (k % every == 0) h5lite::write_npy(out + "/f_" + std::to_string(k / every) + ".npy", solver.f().data(), {m.nx(), m.ny()});
    }
    std::cout << "steps " << nsteps << " t " << solver.t() << " iterations " << iters << " negatives " << solver.negatives_last()
              << " device_seconds " << solver.seconds_device() << std::endl;
  } catch (const std::exception& e) {
    std::cerr << "error: " << e.what() << std::endl;
    return 1;
  }
  return 0;
}
