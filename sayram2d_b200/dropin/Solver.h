// GPU-backed Solver with the reference's public interface (source/Solver.h:18-25):
//
//     Solver(const Mesh& m, Equation* eq);   void update();   double t() const;
//     const Xtensor2d& f() const;            double f(const Ind&) const;
//
// It is written against the PUBLIC API of Mesh / Equation only (nx, ny, x_edge, y_edge,
// dt; G, Dxx, Dxy, Dyy, inv_tau, bc_type, dirichlet_vertex_value, init_f, update), so the
// same two files compile against this directory's host classes or, unchanged, against the
// reference's own headers: replace source/Solver.{h,cc} by these, link
// libsayram2d_b200.so, and main.cc builds as is.
//
// The PPFV assembly and the linear solve (Solver.cc:57-290) run on the GPU behind the C
// ABI of include/sayram2d.h; f stays on the device and is copied back lazily, only when
// f() is called after an update (main.cc does that on output steps only).
#ifndef SOLVER_H
#define SOLVER_H

#include <cstddef>
#include <vector>

#include "Equation.h"
#include "Mesh.h"
#include "common.h"
#include "sayram2d.h"

class Solver {
 public:
  Solver(const Mesh& m_in, Equation* eqp);
  ~Solver();
  Solver(const Solver&) = delete;
  Solver& operator=(const Solver&) = delete;

  void update();                                   // one implicit time step (Solver.cc:270-290)
  double t() const { return istep_ * m.dt(); }
  const Xtensor2d& f() const;                      // (nx, ny); downloads from the device if stale
  double f(const Ind& ind) const { return f()(ind.i, ind.j); }

  // ---- extensions (not in the reference) ----
  void update(int nsteps);                         // nsteps steps without leaving the device; needs a static Equation
  void set_static_equation(bool is_static) { static_eq_ = is_static; }  // skip the per-step re-staging check
  void set_async_staging(bool on) { async_staging_ = on; }  // time-dependent cases: stage t^{n+1} while step n runs (default on)
  void set_device(int device);                     // before the first update(); default 0 or $SY2D_DEVICE
  long long iterations_last() const { return iters_last_; }
  long long iterations_total() const { return iters_total_; }
  double residual_last() const { return resid_last_; }
  long long negatives_last() const { return negatives_last_; }
  double seconds_device() const { return seconds_device_; }

 private:
  const Mesh& m;
  Equation& eq;
  std::size_t istep_ = 0;
  sy2d_ctx* ctx_ = nullptr;
  bool static_eq_ = false;
  bool async_staging_ = true;
  long long fields_ver_ = -2, bc_ver_ = -2;            // Equation dirty counters at the last staging (-1: the Equation has none)

  mutable Xtensor2d f_;
  mutable bool f_stale_ = false;

  std::vector<double> G_, Dxx_, Dxy_, Dyy_, itau_;     // gathered fields (host staging)
  std::vector<double> bc_lines_[4];
  int bc_types_[4] = {1, 1, 1, 1};
  long long iters_last_ = 0, iters_total_ = 0, negatives_last_ = 0;
  double resid_last_ = 0.0, seconds_device_ = 0.0;

  void create_context(int device);
  bool gather_coefficients(bool force);   // false: the Equation's dirty counter says nothing changed
  bool gather_boundaries(double t, bool force);
  void stage(double t, bool force, bool async);
  void upload(bool coeffs, bool boundaries, bool async);
  void account(const sy2d_stats& st);
  void check(int rc) const;
};

#endif /* SOLVER_H */
