// Run parameters read from the ini file: the reference's Parameters interface
// (source/Parameters.h:24-49) with the same derived quantities
// (source/Parameters.cc:50-60): save_every_step = int(nsteps / nplots), nsteps is
// then rounded down to a multiple of it, dt = T / nsteps, angles converted to rad.
#ifndef SY2D_HOST_PARAMETERS_H_
#define SY2D_HOST_PARAMETERS_H_

#include <string>

#include "common.h"

class Parameters {
 public:
  Parameters(int argc, char** argv);
  // Same parsing without the side effects (no ./output directory, no ini copy): for
  // embedding and tests.
  explicit Parameters(const std::string& inp_file, bool make_output_dir = false);

  const std::string& inp_file() const { return inp_file_; }
  const std::string& run_id() const { return run_id_; }
  std::size_t nalpha0() const { return nalpha0_; }
  std::size_t nE() const { return nE_; }
  double alpha0_min() const { return alpha0_min_ * gPI / 180; }
  double alpha0_max() const { return alpha0_max_ * gPI / 180; }
  double Emin() const { return Emin_; }
  double Emax() const { return Emax_; }
  double logEmin() const { return logEmin_; }
  double logEmax() const { return logEmax_; }
  double T() const { return T_; }
  int nsteps() const { return nsteps_; }
  double dt() const { return T_ / nsteps_; }
  int nplots() const { return nplots_; }
  int save_every_step() const { return save_every_step_; }
  const std::string& output_path() const { return output_path_; }
  const std::string& dID() const { return dID_; }

 private:
  std::string inp_file_, run_id_, output_path_, dID_;
  std::size_t nalpha0_ = 0, nE_ = 0;
  double alpha0_min_ = 0, alpha0_max_ = 0, Emin_ = 0, Emax_ = 0, logEmin_ = 0, logEmax_ = 0, T_ = 0;
  double nsteps_ = 0;  // a double in the reference too (Parameters.h:66): dt() divides by it
  int nplots_ = 0, save_every_step_ = 0;

  void read_inp_file();
  void prepare_output_dir();
};

#endif
