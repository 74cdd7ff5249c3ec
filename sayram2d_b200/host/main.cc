// Stand-alone driver with the flow of the reference's source/main.cc:39-96: parameters ->
// uniform (alpha0, log E) grid -> Mesh -> Equation case -> Solver -> time loop with
// nplots + 1 snapshots.  Without libhdf5 the snapshots are written as NumPy files into
// <output_path>/<run_id>_data/ (alpha0.npy [deg], logEN.npy, f_<k>.npy, t.npy - the datasets
// /alpha0, /logEN, /f/<k>, /t of the reference's HDF5 file).
//
//   sayram2d [p.ini] [--case AY|LC]     (the reference selects the case at compile time)
#include <chrono>
#include <cstring>
#include <ctime>
#include <filesystem>
#include <iostream>
#include <memory>

#include "Albert_Young.h"
#include "Mesh.h"
#include "Parameters.h"
#include "Solver.h"
#include "h5lite.h"

static Grid2D make_uniform(const Parameters& p) {  // main.cc:20-37
  std::vector<double> xe(p.nalpha0() + 1), ye(p.nE() + 1);
  const double dx = (p.alpha0_max() - p.alpha0_min()) / static_cast<double>(p.nalpha0());
  const double dy = (p.logEmax() - p.logEmin()) / static_cast<double>(p.nE());
  for (std::size_t i = 0; i <= p.nalpha0(); ++i) xe[i] = p.alpha0_min() + dx * static_cast<double>(i);
  for (std::size_t j = 0; j <= p.nE(); ++j) ye[j] = p.logEmin() + dy * static_cast<double>(j);
  return Grid2D(std::move(xe), std::move(ye));
}

int main(int argc, char** argv) {
  std::string kase = "AY";
  std::vector<char*> rest{argv[0]};
  for (int k = 1; k < argc; ++k) {
    if (!std::strcmp(argv[k], "--case") && k + 1 < argc) kase = argv[++k];
    else rest.push_back(argv[k]);
  }
  try {
    Parameters paras(static_cast<int>(rest.size()), rest.data());
    Grid2D grid = make_uniform(paras);
    Mesh m(grid, paras.dt());
    std::unique_ptr<Equation> eq;
    if (kase == "AY") eq.reset(new Albert_Young(paras, m));
    else if (kase == "LC") eq.reset(new Albert_Young_LC(paras, m));
    else { std::cerr << "unknown case " << kase << " (AY or LC)" << std::endl; return 2; }
    Solver solver(m, eq.get());

    const std::string dir = paras.output_path() + paras.run_id() + "_data";
    std::filesystem::create_directories(dir);
    std::vector<double> alpha0(m.nx()), logEN(m.ny());
    for (std::size_t i = 0; i < m.nx(); ++i) alpha0[i] = m.x(i) / gPI * 180.0;
    for (std::size_t j = 0; j < m.ny(); ++j) logEN[j] = m.y(j) - std::log(gE0);
    h5lite::write_npy(dir + "/alpha0.npy", alpha0.data(), {alpha0.size()});
    h5lite::write_npy(dir + "/logEN.npy", logEN.data(), {logEN.size()});

    const clock_t c0 = clock();
    const auto w0 = std::chrono::steady_clock::now();
    h5lite::write_npy(dir + "/f_0.npy", solver.f().data(), {m.nx(), m.ny()});
    for (int tstep = 1; tstep <= paras.nsteps(); ++tstep) {
      solver.update();
      if (tstep % paras.save_every_step() == 0)
        h5lite::write_npy(dir + "/f_" + std::to_string(tstep / paras.save_every_step()) + ".npy", solver.f().data(), {m.nx(), m.ny()});
    }
    std::vector<double> t(paras.nplots() + 1);
    for (int k = 0; k <= paras.nplots(); ++k) t[k] = paras.T() * k / paras.nplots();
    h5lite::write_npy(dir + "/t.npy", t.data(), {t.size()});
    const double cpu = double(clock() - c0) / CLOCKS_PER_SEC;
    const double wall = std::chrono::duration<double>(std::chrono::steady_clock::now() - w0).count();
    std::cout << "CPU time used " << cpu << " seconds" << std::endl;  // the reference's line (main.cc:93)
    std::cout << "wall " << wall << " s, device " << solver.seconds_device() << " s, " << solver.iterations_total()
              << " BiCGSTAB iterations, " << solver.negatives_last() << " negative cells, output in " << dir << std::endl;
  } catch (const std::exception& e) {
    std::cerr << "sayram2d: " << e.what() << std::endl;
    return 1;
  }
  return 0;
}
