// Problem definition ("plugin" point): the reference's abstract Equation interface
// (source/Equation.h:27-75) - per-cell fields G, Dxx, Dyy, Dxy, inv_tau on (nx, ny),
// boundary types, Dirichlet vertex values, initial f and the time-dependence hook.
#ifndef SY2D_HOST_EQUATION_H_
#define SY2D_HOST_EQUATION_H_

#include <cstddef>

#include "BCTypes.h"
#include "Mesh.h"
#include "Parameters.h"
#include "common.h"
#include "utils.h"

class Equation {
 public:
  explicit Equation(const Mesh& m)
      : G_(m.nx(), m.ny()), Dxx_(m.nx(), m.ny()), Dyy_(m.nx(), m.ny()), Dxy_(m.nx(), m.ny()), inv_tau_(m.nx(), m.ny(), 0.0) {}
  virtual ~Equation() = default;

  double G(const Ind& c) const { return G_(c.i, c.j); }
  double Dxx(const Ind& c) const { return Dxx_(c.i, c.j); }
  double Dyy(const Ind& c) const { return Dyy_(c.i, c.j); }
  double Dxy(const Ind& c) const { return Dxy_(c.i, c.j); }
  double inv_tau(const Ind& c) const { return inv_tau_(c.i, c.j); }  // loss term -f/tau; 0 = no loss

  virtual BCType bc_type(BoundaryID side) const = 0;
  virtual double init_f(const Ind& c) const = 0;
  virtual void update(double t) = 0;  // called once per step with the new time
  // Dirichlet value at boundary vertex (i, j) of `side`; false = no data (zero-flux side)
  virtual bool dirichlet_vertex_value(BoundaryID, std::size_t, std::size_t, double, double*) const { return false; }

  // Extension (not in the reference): a case whose fields and boundary values do not depend
  // on t says so, and the GPU Solver then never re-stages them.
  virtual bool is_static() const { return false; }

  // Extension: dirty counters for time-dependent cases.  A case that bumps fields_version() / bc_version() exactly when
  // update(t) changed its fields / Dirichlet data lets the GPU Solver skip the gathering of unchanged data; the default
  // (-1) means "unknown": the Solver then re-stages after every update(t), overlapped with the running time step.
  virtual long long fields_version() const { return -1; }
  virtual long long bc_version() const { return -1; }

  // whole fields, for staging to the device in one copy
  const Xtensor2d& G_field() const { return G_; }
  const Xtensor2d& Dxx_field() const { return Dxx_; }
  const Xtensor2d& Dyy_field() const { return Dyy_; }
  const Xtensor2d& Dxy_field() const { return Dxy_; }
  const Xtensor2d& inv_tau_field() const { return inv_tau_; }

 protected:
  Xtensor2d G_, Dxx_, Dyy_, Dxy_, inv_tau_;
};

#endif
