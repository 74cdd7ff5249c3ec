#include "Solver.h"

#include <cstdlib>
#include <cstring>
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
  G_.clear();
  for (auto& l : bc_lines_) l.clear();
  stage(true);                                        // update_Lambda + update_vertex_f at t = 0 (Solver.cc:44-45)
  check(sy2d_set_f(ctx_, f_.data()));
  f_stale_ = false;
}

void Solver::set_device(int device) {
  if (istep_ != 0) throw std::runtime_error("Solver::set_device: call before the first update()");
  create_context(device);
}

// Equation fields through the per-cell accessors (Equation.h:44-50); Lambda = G*D is formed
// on the device (Solver.cc:57-65).
bool Solver::gather_coefficients() {
  const std::size_t nx = m.nx(), ny = m.ny(), N = nx * ny;
  std::vector<double> g(N), dxx(N), dxy(N), dyy(N), it(N);
  for (std::size_t i = 0; i < nx; ++i)
    for (std::size_t j = 0; j < ny; ++j) {
      const Ind c{i, j};
      const std::size_t n = i * ny + j;
      g[n] = eq.G(c); dxx[n] = eq.Dxx(c); dxy[n] = eq.Dxy(c); dyy[n] = eq.Dyy(c); it[n] = eq.inv_tau(c);
    }
  const bool same = G_.size() == N && !std::memcmp(g.data(), G_.data(), N * 8) && !std::memcmp(dxx.data(), Dxx_.data(), N * 8) &&
                    !std::memcmp(dxy.data(), Dxy_.data(), N * 8) && !std::memcmp(dyy.data(), Dyy_.data(), N * 8) &&
                    !std::memcmp(it.data(), itau_.data(), N * 8);
  if (same) return false;
  G_.swap(g); Dxx_.swap(dxx); Dxy_.swap(dxy); Dyy_.swap(dyy); itau_.swap(it);
  return true;
}

// Boundary types and Dirichlet vertex lines at time t (Solver.cc:385-422).
bool Solver::gather_boundaries(double tt) {
  const std::size_t nx = m.nx(), ny = m.ny();
  const BoundaryID sides[4] = {BoundaryID::XMIN, BoundaryID::XMAX, BoundaryID::YMIN, BoundaryID::YMAX};
  bool changed = false;
  for (int s = 0; s < 4; ++s) {
    const int type = eq.bc_type(sides[s]) == BCType::Dirichlet ? SY2D_DIRICHLET : SY2D_ZEROFLUX;
    std::vector<double> line;
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
    if (type != bc_types_[s] || line != bc_lines_[s]) changed = true;
    bc_types_[s] = type;
    bc_lines_[s].swap(line);
  }
  return changed;
}

void Solver::stage(bool force) {
  if (gather_coefficients() || force)
    check(sy2d_set_coeffs(ctx_, G_.data(), Dxx_.data(), Dxy_.data(), Dyy_.data(), itau_.data()));
  if (gather_boundaries(t()) || force) {
    const double* lines[4];
    for (int s = 0; s < 4; ++s) lines[s] = bc_lines_[s].empty() ? nullptr : bc_lines_[s].data();
    check(sy2d_set_bc(ctx_, bc_types_, lines[0], lines[1], lines[2], lines[3]));
  }
}

void Solver::update() {
  sy2d_stats st;
  check(sy2d_step(ctx_, 1, &st));                     // assemble + solve (Solver.cc:271-284)
  iters_last_ = st.iters_last;
  iters_total_ += st.iters_total;
  resid_last_ = st.resid_last;
  negatives_last_ = st.negatives;
  seconds_device_ += st.seconds_device;
  f_stale_ = true;
  istep_ += 1;                                        // Solver.cc:286
  eq.update(t());                                     // Solver.cc:287
  if (!static_eq_) stage(false);                      // update_Lambda + update_vertex_f (Solver.cc:288-289)
}

void Solver::update(int nsteps) {
  if (nsteps <= 0) return;
  if (!static_eq_) {
    for (int k = 0; k < nsteps; ++k) update();
    return;
  }
  sy2d_stats st;
  check(sy2d_step(ctx_, nsteps, &st));
  iters_last_ = st.iters_last;
  iters_total_ += st.iters_total;
  resid_last_ = st.resid_last;
  negatives_last_ = st.negatives;
  seconds_device_ += st.seconds_device;
  f_stale_ = true;
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
