#include "Solver.h"

#include <cstdlib>
#include <cstring>
#include <exception>
#include <chrono>
#include <future>
#include <thread>
#include <stdexcept>
#include <string>

#include "BCTypes.h"
#include "sayram2d.h"

namespace {
// Equation::is_static() exists in this directory's Equation.h but not in the reference's:
// detect it so the same source compiles against both.
template <class E>
auto equation_is_static(const E& e, int) -> decltype(e.is_static()) { return e.is_static(); }
template <class E>
bool equation_is_static(const E&, long) { return false; }

// Equation::fields_version() / bc_version() (this directory's Equation.h, not the reference's): counters a time-dependent
// case bumps whenever update(t) really changed its fields / boundary data - the dirty flag that replaces gathering and
// comparing 5 N values per step.  Without them (an unmodified reference Equation) every step is treated as dirty.
template <class E>
auto equation_fields_version(const E& e, int) -> decltype(e.fields_version()) { return e.fields_version(); }
template <class E>
long long equation_fields_version(const E&, long) { return -1; }
template <class E>
auto equation_bc_version(const E& e, int) -> decltype(e.bc_version()) { return e.bc_version(); }
template <class E>
long long equation_bc_version(const E&, long) { return -1; }

int default_device() {
  const char* s = std::getenv("SY2D_DEVICE");
  return s ? std::atoi(s) : 0;
}
}  // namespace

Solver::Solver(const Mesh& m_in, Equation* eqp) : m(m_in), eq(*eqp) {
  static_eq_ = equation_is_static(eq, 0);
  f_.resize({m.nx(), m.ny()});
  for (std::size_t i = 0; i < m.nx(); ++i)            // Solver.cc:38-42
    for (std::size_t j = 0; j < m.ny(); ++j) f_(i, j) = eq.init_f({i, j});
  create_context(default_device());
}

Solver::~Solver() { sy2d_destroy(ctx_); }

void Solver::check(int rc) const {
  if (rc != SY2D_OK) throw std::runtime_error(sy2d_last_error(ctx_));
}

void Solver::create_context(int device) {
  std::vector<double> xe(m.nx() + 1), ye(m.ny() + 1);
  for (std::size_t i = 0; i <= m.nx(); ++i) xe[i] = m.x_edge(i);
  for (std::size_t j = 0; j <= m.ny(); ++j) ye[j] = m.y_edge(j);
  sy2d_ctx* c = nullptr;
  const int rc = sy2d_create(&c, device, static_cast<int>(m.nx()), static_cast<int>(m.ny()), 1, xe.data(), ye.data(), m.dt());
  if (rc != SY2D_OK) throw std::runtime_error(sy2d_last_error(nullptr));
  if (ctx_) sy2d_destroy(ctx_);
  ctx_ = c;
  stage(t(), true, false);                            // update_Lambda + update_vertex_f at t = 0 (Solver.cc:44-45)
  check(sy2d_set_f(ctx_, f_.data()));
  f_stale_ = false;
}

void Solver::set_device(int device) {
  if (istep_ != 0) throw std::runtime_error("Solver::set_device: call before the first update()");
  create_context(device);
}

// Equation fields through the per-cell accessors (Equation.h:44-50); Lambda = G*D is formed
// on the device (Solver.cc:57-65).  Returns false when the Equation's dirty counter says nothing changed.
bool Solver::gather_coefficients(bool force) {
  const long long ver = equation_fields_version(eq, 0);
  if (!force && ver >= 0 && ver == fields_ver_) return false;
  fields_ver_ = ver;
  const std::size_t nx = m.nx(), ny = m.ny(), N = nx * ny;
  G_.resize(N); Dxx_.resize(N); Dxy_.resize(N); Dyy_.resize(N); itau_.resize(N);
  for (std::size_t i = 0; i < nx; ++i)
    for (std::size_t j = 0; j < ny; ++j) {
      const Ind c{i, j};
      const std::size_t n = i * ny + j;
      G_[n] = eq.G(c); Dxx_[n] = eq.Dxx(c); Dxy_[n] = eq.Dxy(c); Dyy_[n] = eq.Dyy(c); itau_[n] = eq.inv_tau(c);
    }
  return true;
}

// Boundary types and Dirichlet vertex lines at time t (Solver.cc:385-422).
bool Solver::gather_boundaries(double tt, bool force) {
  const long long ver = equation_bc_version(eq, 0);
  if (!force && ver >= 0 && ver == bc_ver_) return false;
  bc_ver_ = ver;
  const std::size_t nx = m.nx(), ny = m.ny();
  const BoundaryID sides[4] = {BoundaryID::XMIN, BoundaryID::XMAX, BoundaryID::YMIN, BoundaryID::YMAX};
  for (int s = 0; s < 4; ++s) {
    const int type = eq.bc_type(sides[s]) == BCType::Dirichlet ? SY2D_DIRICHLET : SY2D_ZEROFLUX;
    std::vector<double>& line = bc_lines_[s];
    line.clear();
    if (type == SY2D_DIRICHLET) {
      const std::size_t n = (s < 2 ? ny : nx) + 1;
      line.resize(n);
      for (std::size_t k = 0; k < n; ++k) {
        const std::size_t vi = s == 0 ? 0 : s == 1 ? nx : k;
        const std::size_t vj = s == 2 ? 0 : s == 3 ? ny : k;
        double u = 0.0;
        if (!eq.dirichlet_vertex_value(sides[s], vi, vj, tt, &u)) throw std::runtime_error("Dirichlet BC: missing value.");
        line[k] = u;
      }
    }
    bc_types_[s] = type;
  }
  return true;
}

// update_Lambda + update_vertex_f of the reference (Solver.cc:57-65, 292-422) for time tt: gather what changed and hand
// it to the library - blocking (first staging) or asynchronously into the second buffer set (every later step).
void Solver::stage(double tt, bool force, bool async) {
  upload(gather_coefficients(force), gather_boundaries(tt, force), async);
}

void Solver::upload(bool coeffs, bool boundaries, bool async) {
  if (coeffs) {
    if (async) check(sy2d_set_coeffs_async(ctx_, G_.data(), Dxx_.data(), Dxy_.data(), Dyy_.data(), itau_.data()));
    else check(sy2d_set_coeffs(ctx_, G_.data(), Dxx_.data(), Dxy_.data(), Dyy_.data(), itau_.data()));
  }
  if (boundaries) {
    const double* lines[4];
    for (int s = 0; s < 4; ++s) lines[s] = bc_lines_[s].empty() ? nullptr : bc_lines_[s].data();
    if (async) check(sy2d_set_bc_async(ctx_, bc_types_, lines[0], lines[1], lines[2], lines[3]));
    else check(sy2d_set_bc(ctx_, bc_types_, lines[0], lines[1], lines[2], lines[3]));
  }
}

void Solver::account(const sy2d_stats& st) {
  iters_last_ = st.iters_last;
  iters_total_ += st.iters_total;
  resid_last_ = st.resid_last;
  negatives_last_ = st.negatives;
  seconds_device_ += st.seconds_device;
  f_stale_ = true;
}

void Solver::update() {
  sy2d_stats st;
  if (static_eq_ || !async_staging_) {
    check(sy2d_step(ctx_, 1, &st));                   // assemble + solve (Solver.cc:271-284)
    account(st);
    istep_ += 1;                                      // Solver.cc:286
    eq.update(t());                                   // Solver.cc:287
    if (!static_eq_) stage(t(), false, false);        // update_Lambda + update_vertex_f (Solver.cc:288-289)
    return;
  }
  // Time-dependent case: the fields and boundary data of t^{n+1} do not depend on the solve of step n, so
  // Equation::update(t^{n+1}), the gathering through the accessors and the upload (a second device buffer set filled
  // on a copy stream) run on this thread WHILE the GPU advances step n on a helper thread; the new set is swapped in at
  // the start of step n+1.  The reference's order of effects (Solver.cc:286-289) is kept for everything the Solver's
  // result depends on; the only visible difference: eq.update(t^{n+1}) has already run if step n throws.
  const long long begun = sy2d_steps_begun(ctx_);
  std::future<int> step = std::async(std::launch::async, [&] { return sy2d_step(ctx_, 1, &st); });
  std::exception_ptr err;
  try {
    const double tnext = static_cast<double>(istep_ + 1) * m.dt();
    eq.update(tnext);
    const bool coeffs = gather_coefficients(false), boundaries = gather_boundaries(tnext, false);
    // step n must have taken ITS fields (swap-in at the start of sy2d_step) before those of step n+1 are staged
    while (sy2d_steps_begun(ctx_) == begun && step.wait_for(std::chrono::seconds(0)) != std::future_status::ready) std::this_thread::yield();
    upload(coeffs, boundaries, true);
  } catch (...) {
    err = std::current_exception();
  }
  const int rc = step.get();
  if (err) std::rethrow_exception(err);
  check(rc);
  account(st);
  istep_ += 1;
}

void Solver::update(int nsteps) {
  if (nsteps <= 0) return;
  if (!static_eq_) {
    for (int k = 0; k < nsteps; ++k) update();
    return;
  }
  sy2d_stats st;
  check(sy2d_step(ctx_, nsteps, &st));
  account(st);
  istep_ += static_cast<std::size_t>(nsteps);
  eq.update(t());
}

const Xtensor2d& Solver::f() const {
  if (f_stale_) {
    if (sy2d_get_f(ctx_, f_.data()) != SY2D_OK) throw std::runtime_error(sy2d_last_error(ctx_));
    f_stale_ = false;
  }
  return f_;
}
