// Stand-alone driver with the flow of the reference's source/main.cc:39-96: parameters ->
// uniform (alpha0, log E) grid -> Mesh -> Equation case -> Solver -> time loop with
// nplots + 1 snapshots.  The output is the reference's HDF5 file <output_path>/<run_id>_data.h5 with
// /alpha0 [deg], /logEN, /f/<k>, /t (main.cc:58-67,74,83,87-89; written by h5lite::Writer, no libhdf5
// needed, plot/cmp_ay.py reads it unchanged) plus the same arrays as NumPy files in <run_id>_data/.
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
    h5lite::Writer h5;
    h5.add("/alpha0", alpha0.data(), {alpha0.size()});     // main.cc:63-66
    h5.add("/logEN", logEN.data(), {logEN.size()});

    const clock_t c0 = clock();
    const auto w0 = std::chrono::steady_clock::now();
    h5lite::write_npy(dir + "/f_0.npy", solver.f().data(), {m.nx(), m.ny()});
    h5.add("/f/0", solver.f().data(), {m.nx(), m.ny()});  // main.cc:74
    for (int tstep = 1; tstep <= paras.nsteps(); ++tstep) {
      solver.update();
      if (tstep % paras.save_every_step() == 0) {
        const std::string k = std::to_string(tstep / paras.save_every_step());
        h5lite::write_npy(dir + "/f_" + k + ".npy", solver.f().data(), {m.nx(), m.ny()});
        h5.add("/f/" + k, solver.f().data(), {m.nx(), m.ny()});   // main.cc:83
      }
    }
    std::vector<double> t(paras.nplots() + 1);
    for (int k = 0; k <= paras.nplots(); ++k) t[k] = paras.T() * k / paras.nplots();
    h5lite::write_npy(dir + "/t.npy", t.data(), {t.size()});
    h5.add("/t", t.data(), {t.size()});                    // main.cc:87-89
    h5.save(paras.output_path() + paras.run_id() + "_data.h5");
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
