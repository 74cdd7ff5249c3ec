#include "Albert_Young.h"

Albert_Young_Base::Albert_Young_Base(const Parameters& paras_in, const Mesh& m_in, double sin_loss_cone)
    : Equation(m_in), paras(paras_in), m(m_in), io(paras_in), sin_lc_(sin_loss_cone) {
  constructG();
  constructD();
}

void Albert_Young_Base::constructG() {
  for (std::size_t i = 0; i < m.nx(); ++i)
    for (std::size_t j = 0; j < m.ny(); ++j) G_(i, j) = calculate_G(m.x(i), m.y(j));
}

// Fractional position in the table, which is uniform in alpha0 and in log E
// (Albert_Young.cc:94-111).
void Albert_Young_Base::locate(double alpha0, double logE, Loc* loc) const {
  const double pos_x = (alpha0 - io.xmin_D()) / (io.xmax_D() - io.xmin_D()) * (io.nx_D() - 1);
  const double pos_y = (logE - std::log(io.ymin_D())) / (std::log(io.ymax_D()) - std::log(io.ymin_D())) * (io.ny_D() - 1);
  std::size_t i0 = static_cast<std::size_t>(static_cast<long long>(std::floor(pos_x)));  // negative wraps, as in the reference
  std::size_t j0 = static_cast<std::size_t>(static_cast<long long>(std::floor(pos_y)));
  double wi, wj;
  calWeight(i0, wi, io.nx_D() - 1, pos_x);
  calWeight(j0, wj, io.ny_D() - 1, pos_y);
  loc->i0 = static_cast<int>(i0);
  loc->j0 = static_cast<int>(j0);
  loc->wi = wi;
  loc->wj = wj;
}

// Table values are normalised and per second; convert to (alpha0, log E) coordinates and
// per day (Albert_Young.cc:113-134).
void Albert_Young_Base::constructD() {
  const double denormalize_factor = gME * gME * gC * gC;
  const double second_to_day = 3600 * 24;
  Loc loc;
  for (std::size_t i = 0; i < m.nx(); ++i) {
    const double alpha0 = m.x(i);
    for (std::size_t j = 0; j < m.ny(); ++j) {
      const double logE = m.y(j);
      const double p = e2p(std::exp(logE), gE0);
      locate(alpha0, logE, &loc);
      Dxx_(i, j) = interp2D(io.Dxx_raw, loc) * denormalize_factor * second_to_day / (p * p);
      Dxy_(i, j) = interp2D(io.Dxy_raw, loc) * denormalize_factor * second_to_day * dlogE_dp(logE, gE0) / p;
      Dyy_(i, j) = interp2D(io.Dyy_raw, loc) * denormalize_factor * second_to_day * std::pow(dlogE_dp(logE, gE0), 2);
    }
  }
}

// ------------------------------------------------------------------ Albert_Young
Albert_Young::Albert_Young(const Parameters& paras_in, const Mesh& m_in)
    : Albert_Young_Base(paras_in, m_in, std::sin(5 * gPI / 180)) {}

BCType Albert_Young::bc_type(BoundaryID side) const {  // Albert_Young.cc:42-59
  switch (side) {
    case BoundaryID::XMIN: case BoundaryID::YMIN: case BoundaryID::YMAX: return BCType::Dirichlet;
    case BoundaryID::XMAX: return BCType::ZeroFlux;
  }
  throw std::runtime_error("Albert_Young::bc_type: unknown BoundaryID");
}

bool Albert_Young::dirichlet_vertex_value(BoundaryID side, std::size_t i, std::size_t, double, double* out) const {
  switch (side) {  // Albert_Young.cc:64-92, Albert_Young.h:52-62
    case BoundaryID::XMIN: *out = 0.0; return true;
    case BoundaryID::YMIN: *out = ymin(m.x_edge(i)); return true;
    case BoundaryID::YMAX: *out = 0.0; return true;
    case BoundaryID::XMAX: return false;
  }
  throw std::runtime_error("Albert_Young::dirichlet_value: unknown BoundaryID");
}

// --------------------------------------------------------------- Albert_Young_LC
Albert_Young_LC::Albert_Young_LC(const Parameters& paras_in, const Mesh& m_in) : Albert_Young_Base(paras_in, m_in, 0.0) {
  const double L = 4.5;                                                        // Albert_Young_LC.cc:42
  const double alpha0lc = std::asin(std::pow(std::pow(L, 5) * (4 * L - 3), -0.25));
  for (std::size_t i = 0; i < m.nx(); ++i) {
    const double alpha0 = m.x(i);
    if (alpha0 >= alpha0lc) continue;
    for (std::size_t j = 0; j < m.ny(); ++j) inv_tau_(i, j) = 4.0 / bounce_period(alpha0, e2p(std::exp(m.y(j)), gE0), L);
  }
}

BCType Albert_Young_LC::bc_type(BoundaryID side) const {  // Albert_Young_LC.cc:56-73
  switch (side) {
    case BoundaryID::XMIN: case BoundaryID::XMAX: return BCType::ZeroFlux;
    case BoundaryID::YMIN: case BoundaryID::YMAX: return BCType::Dirichlet;
  }
  throw std::runtime_error("Albert_Young_LC::bc_type: unknown BoundaryID");
}

bool Albert_Young_LC::dirichlet_vertex_value(BoundaryID side, std::size_t i, std::size_t, double, double* out) const {
  switch (side) {  // Albert_Young_LC.cc:78-106
    case BoundaryID::YMIN: *out = ymin(m.x_edge(i)); return true;
    case BoundaryID::YMAX: *out = 0.0; return true;
    case BoundaryID::XMIN: case BoundaryID::XMAX: return false;
  }
  throw std::runtime_error("Albert_Young_LC::dirichlet_value: unknown BoundaryID");
}

double Albert_Young_LC::bounce_period(double a0, double p, double L) const {  // Albert_Young_LC.cc:150-160
  const double T0 = 1.3802, T1 = 0.7405;
  const double y = std::sin(a0);
  const double Ty = T0 - 0.5 * (T0 - T1) * (y + std::sqrt(y));
  return 4 * L * gRE * ((gE0 + p2e(p, gE0)) / (gC * gC)) / p * Ty / (3e8 * 3600 * 24);
}
