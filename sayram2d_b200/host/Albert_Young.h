// Albert & Young (2005) chorus cases behind the reference's Equation interface:
//   Albert_Young    - source/Cases/Albert_Young.{h,cc}: Dirichlet at alpha0_min (f=0), E_min
//                     (initial profile) and E_max (f=0); zero flux at alpha0 = 90 deg.
//   Albert_Young_LC - source/Cases/Albert_Young_LC.{h,cc}: alpha0_min = 0, zero flux on both
//                     alpha0 sides, loss term 1/tau = 4/tau_bounce inside the loss cone (L = 4.5).
// Both share G, the table lookup of D and the unit conversion, factored into one base here
// (the reference duplicates them per case).
#ifndef SY2D_HOST_ALBERT_YOUNG_H_
#define SY2D_HOST_ALBERT_YOUNG_H_

#include "Albert_Young_IO.h"
#include "Equation.h"

class Albert_Young_Base : public Equation {
 public:
  void update(double) override {}           // static case: nothing depends on t
  bool is_static() const override { return true; }
  double init_f(const Ind& c) const override { return calculate_init_f(m.x(c.i), m.y(c.j)); }

 protected:
  Albert_Young_Base(const Parameters& paras_in, const Mesh& m_in, double sin_loss_cone);
  const Parameters& paras;
  const Mesh& m;
  Albert_Young_IO io;
  double sin_lc_;  // sin of the loss-cone angle subtracted in the initial profile (0 for the LC case)

  double calculate_init_f(double a, double logE) const {  // Albert_Young.h:37-40 / Albert_Young_LC.h:37-40
    const double p = e2p(std::exp(logE), gE0);
    return std::exp(-(std::exp(logE) - 0.2) / 0.1) * (std::sin(a) - sin_lc_) / (p * p) + gEPS;
  }
  static double calculate_G(double alpha, double logE) {  // Albert_Young.h:42-45
    const double t = 1.30 - 0.56 * std::sin(alpha);
    return std::pow(e2p(std::exp(logE), gE0), 2) * t * std::sin(alpha) * std::cos(alpha) / dlogE_dp(logE, gE0);
  }
  double ymin(double a0) const { return calculate_init_f(a0, paras.logEmin()); }
  void locate(double alpha0, double logE, Loc* loc) const;

 private:
  void constructG();
  void constructD();
};

class Albert_Young : public Albert_Young_Base {
 public:
  Albert_Young(const Parameters& paras_in, const Mesh& m_in);
  BCType bc_type(BoundaryID side) const override;
  bool dirichlet_vertex_value(BoundaryID side, std::size_t i, std::size_t j, double t, double* out) const override;
};

class Albert_Young_LC : public Albert_Young_Base {
 public:
  Albert_Young_LC(const Parameters& paras_in, const Mesh& m_in);
  BCType bc_type(BoundaryID side) const override;
  bool dirichlet_vertex_value(BoundaryID side, std::size_t i, std::size_t j, double t, double* out) const override;

 private:
  double bounce_period(double a0, double p, double L) const;
};

#endif
