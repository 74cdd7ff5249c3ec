#include "Parameters.h"

#include <cstdlib>
#include <filesystem>

#include "Ini_reader.h"

Parameters::Parameters(int argc, char** argv) {
  if (argc == 1) {
    inp_file_ = "p.ini";  // source/Parameters.cc:21-23
  } else if (argc == 2) {
    inp_file_ = argv[1];
  } else {
    std::cerr << "NParas() > 2! This program takes at most one argument: the parameter file name." << std::endl;
    std::exit(1);
  }
  read_inp_file();
  prepare_output_dir();
}

Parameters::Parameters(const std::string& inp_file, bool make_output_dir) : inp_file_(inp_file) {
  read_inp_file();
  if (make_output_dir) prepare_output_dir();
}

void Parameters::prepare_output_dir() {  // source/Parameters.cc:7-16
  output_path_ = "./output/" + run_id_ + "/";
  std::filesystem::create_directories(output_path_);
  std::error_code ec;
  std::filesystem::copy_file(inp_file_, output_path_ + run_id_ + ".ini", std::filesystem::copy_options::overwrite_existing, ec);
  if (ec) std::cerr << "Command failed with " << ec.value() << std::endl;
}

void Parameters::read_inp_file() {  // source/Parameters.cc:35-65
  Ini_reader ini(inp_file_);
  ini.set_section("basic");
  ini.read("run_id", &run_id_);
  ini.read("nalpha0", &nalpha0_);
  ini.read("nE", &nE_);
  ini.read("alpha0min", &alpha0_min_);
  ini.read("alpha0max", &alpha0_max_);
  ini.read("Emin", &Emin_);
  ini.read("Emax", &Emax_);
  logEmin_ = std::log(Emin_);
  logEmax_ = std::log(Emax_);
  ini.read("T", &T_);
  ini.read("nsteps", &nsteps_);
  ini.set_section("diagnostics");
  ini.read("nplots", &nplots_);
  save_every_step_ = static_cast<int>(nsteps_ / nplots_);
  nsteps_ = save_every_step_ * nplots_;
  ini.set_section("diffusion_coefficients");
  ini.read("dID", &dID_);
  output_path_ = "./output/" + run_id_ + "/";
}
