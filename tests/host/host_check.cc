// Test helper (CPU only): exercises the C++ host layer of sayram2d_b200/host without the
// GPU Solver and dumps what it computes, for comparison with the golden fixtures.
//   host_check dump <AY|LC> <ini> <outdir>     fields, init f, BC types and lines, mesh
//   host_check errors                          validation / error behaviour (prints PASS lines)
#include <cstring>
#include <iostream>
#include <memory>

#include "Albert_Young.h"
#include "Ini_reader.h"
#include "Mesh.h"
#include "Parameters.h"
#include <algorithm>
#include "h5lite.h"

static Grid2D make_uniform(const Parameters& p) {
  std::vector<double> xe(p.nalpha0() + 1), ye(p.nE() + 1);
  const double dx = (p.alpha0_max() - p.alpha0_min()) / static_cast<double>(p.nalpha0());
  const double dy = (p.logEmax() - p.logEmin()) / static_cast<double>(p.nE());
  for (std::size_t i = 0; i <= p.nalpha0(); ++i) xe[i] = p.alpha0_min() + dx * static_cast<double>(i);
  for (std::size_t j = 0; j <= p.nE(); ++j) ye[j] = p.logEmin() + dy * static_cast<double>(j);
  return Grid2D(std::move(xe), std::move(ye));
}

template <class F>
static bool throws_with(F&& fn, const char* needle) {
  try { fn(); } catch (const std::exception& e) { return std::strstr(e.what(), needle) != nullptr; }
  return false;
}

int main(int argc, char** argv) {
  if (argc >= 2 && !std::strcmp(argv[1], "errors")) {
    int bad = 0;
    auto expect = [&](bool ok, const char* what) { std::cout << (ok ? "PASS " : "FAIL ") << what << std::endl; bad += !ok; };
    expect(throws_with([] { Grid2D g({0.0}, {0.0, 1.0}); }, "x_edges must have size >= 2"), "Grid2D size");
    expect(throws_with([] { Grid2D g({0.0, 1.0, 1.0}, {0.0, 1.0}); }, "x_edges must be strictly increasing at i=1"), "Grid2D x order");
    expect(throws_with([] { Grid2D g({0.0, 1.0}, {0.0, 2.0, 1.0}); }, "y_edges must be strictly increasing at j=1"), "Grid2D y order");
    Grid2D g({0.0, 1.0, 3.0}, {0.0, 0.5, 1.5, 2.0});
    Mesh m(g, 0.1);
    Ind nb{};
    expect(m.nx() == 2 && m.ny() == 3 && m.x(1) == 2.0 && m.dx(1) == 2.0 && m.y(0) == 0.25, "Mesh geometry");
    expect(m.flatten_cell_index({1, 2}) == 5 && m.rinbr(0) == 2 && m.rinbr(1) == 3 && m.rinbr(2) == 0 && m.rinbr(3) == 1, "Mesh numbering");
    expect(!m.get_nbr_ind({0, 0}, m.inbr_im(), &nb) && !m.get_nbr_ind({0, 0}, m.inbr_jm(), &nb) && m.get_nbr_ind({0, 0}, m.inbr_ip(), &nb) && nb.i == 1 && nb.j == 0, "Mesh neighbours");
    Edge e;
    m.get_nbr_edge({1, 1}, m.inbr_ip(), &e);  // east face: A = SE, B = NE, normal +x, length dy
    expect(e.v[0][0] == 3.0 && e.v[0][1] == 0.5 && e.v[1][1] == 1.5 && e.length == 1.0 && e.n[0] == 1.0 && e.vind[0].i == 2 && e.vind[0].j == 1, "Mesh edge E");
    m.get_nbr_edge({1, 1}, m.inbr_im(), &e);  // west face: A = NW, B = SW
    expect(e.v[0][0] == 1.0 && e.v[0][1] == 1.5 && e.v[1][1] == 0.5 && e.n[0] == -1.0, "Mesh edge W");
    expect(std::abs(m.cell_area_dt({1, 1}) - 2.0 * 1.0 / 0.1) < 1e-12, "cell_area_dt");
    return bad;
  }
  if (argc >= 3 && !std::strcmp(argv[1], "ini")) {
    Ini_reader r(argv[2]);
    int bad = 0;
    auto expect = [&](bool ok, const char* what) { std::cout << (ok ? "PASS " : "FAIL ") << what << std::endl; bad += !ok; };
    std::size_t n = 0; double x = 0; std::string s; bool b = true;
    r.read("basic", "NALPHA0", &n); expect(n == 10, "case-insensitive key");
    r.set_section("Basic"); r.read("emin", &x); expect(x == 0.2, "case-insensitive section + set_section");
    r.read("run_id", &s); expect(s == "q", "string value");
    r.read("flag", &b); expect(!b, "bool value");
    expect(!r.has("basic", "commented"), "';' comment skipped");
    bool t1 = false, t2 = false;
    try { r.read("nosuch", "k", &x); } catch (const Ini_reader::section_not_found& e) { t1 = e.section == "nosuch"; }
    try { r.read("basic", "nokey", &x); } catch (const Ini_reader::key_not_found& e) { t2 = e.key == "nokey"; }
    expect(t1, "section_not_found"); expect(t2, "key_not_found");
    Parameters p(argv[2]);
    expect(p.save_every_step() == 50 && p.nsteps() == 500 && p.dt() == 1.0 / 500 && p.nplots() == 10, "nsteps rounding (Parameters.cc:59-60)");
    expect(p.alpha0_min() == 5 * gPI / 180 && p.logEmax() == std::log(5.0) && p.dID() == "X" && p.output_path() == "./output/q/", "derived values");
    return bad;
  }
  if (argc >= 3 && !std::strcmp(argv[1], "h5write")) {
    // h5write <file> [mirror]: "mirror" writes datasets with the names and shapes of D/AlbertYoung_chorus.h5
    // (so the structures can be compared byte for byte with a file written by libhdf5); otherwise the
    // layout of the reference's output file with 13 snapshots under /f (more than one default leaf node).
    h5lite::Writer w;
    auto ramp = [](std::size_t n, double a) { std::vector<double> v(n); for (std::size_t k = 0; k < n; ++k) v[k] = a + 0.5 * static_cast<double>(k); return v; };
    if (argc >= 4 && !std::strcmp(argv[3], "mirror")) {
      w.add("/alpha0", ramp(91, 0.0).data(), {91});
      w.add("/E", ramp(49, 1.0).data(), {49});
      w.add("/Daa", ramp(91 * 49, 2.0).data(), {91, 49});
      w.add("/Dap", ramp(91 * 49, 3.0).data(), {91, 49});
      w.add("/Dpp", ramp(91 * 49, 4.0).data(), {91, 49});
    } else {
      w.add("/alpha0", ramp(6, 5.0).data(), {6});
      w.add("/logEN", ramp(4, -1.0).data(), {4});
      for (int k = 0; k <= 12; ++k) w.add("/f/" + std::to_string(k), ramp(24, 100.0 * k).data(), {6, 4});
      w.add("/t", ramp(13, 0.0).data(), {13});
    }
    w.save(argv[2]);
    h5lite::File back(argv[2]);                      // this repo's C++ reader walks what the writer wrote
    std::cout << "datasets " << back.datasets().size() << " f/12[23] " << (back.datasets().count("/f/12") ? back.read("/f/12")[23] : -1.0) << std::endl;
    return 0;
  }
  if (argc >= 4 && !std::strcmp(argv[1], "h5read")) {
    // h5read <file> <outdir>: every dataset of the file as <outdir>/<path with '/' -> '_'>.npy; errors as "error: <what>", exit 3
    try {
      h5lite::File h(argv[2]);
      for (const auto& kv : h.datasets()) {
        std::vector<std::size_t> shape;
        const std::vector<double> v = h.read(kv.first, &shape);
        std::string leaf = kv.first;
        std::replace(leaf.begin(), leaf.end(), '/', '_');
        h5lite::write_npy(std::string(argv[3]) + "/" + leaf + ".npy", v.data(), shape);
        std::cout << kv.first << std::endl;
      }
    } catch (const std::exception& e) {
      std::cout << "error: " << e.what() << std::endl;
      return 3;
    }
    return 0;
  }
  if (argc < 5 || std::strcmp(argv[1], "dump")) { std::cerr << "usage: host_check dump <AY|LC> <ini> <outdir> | errors | ini <file>" << std::endl; return 2; }
  const std::string kase = argv[2], out = argv[4];
  Parameters paras(argv[3]);
  Grid2D grid = make_uniform(paras);
  Mesh m(grid, paras.dt());
  std::unique_ptr<Equation> eq;
  if (kase == "AY") eq.reset(new Albert_Young(paras, m)); else eq.reset(new Albert_Young_LC(paras, m));
  const std::size_t nx = m.nx(), ny = m.ny();
  h5lite::write_npy(out + "/x_edges.npy", m.x_edges().data(), {nx + 1});
  h5lite::write_npy(out + "/y_edges.npy", m.y_edges().data(), {ny + 1});
  h5lite::write_npy(out + "/G.npy", eq->G_field().data(), {nx, ny});
  h5lite::write_npy(out + "/Dxx.npy", eq->Dxx_field().data(), {nx, ny});
  h5lite::write_npy(out + "/Dxy.npy", eq->Dxy_field().data(), {nx, ny});
  h5lite::write_npy(out + "/Dyy.npy", eq->Dyy_field().data(), {nx, ny});
  h5lite::write_npy(out + "/inv_tau.npy", eq->inv_tau_field().data(), {nx, ny});
  std::vector<double> f0(nx * ny);
  for (std::size_t i = 0; i < nx; ++i) for (std::size_t j = 0; j < ny; ++j) f0[i * ny + j] = eq->init_f({i, j});
  h5lite::write_npy(out + "/f_0.npy", f0.data(), {nx, ny});
  const BoundaryID sides[4] = {BoundaryID::XMIN, BoundaryID::XMAX, BoundaryID::YMIN, BoundaryID::YMAX};
  std::vector<double> types(4);
  for (int s = 0; s < 4; ++s) {
    types[s] = eq->bc_type(sides[s]) == BCType::Dirichlet ? 0 : 1;
    const std::size_t n = (s < 2 ? ny : nx) + 1;
    std::vector<double> line(n, -1.0), has(1, 1.0);
    for (std::size_t k = 0; k < n; ++k) {
      double u = 0;
      const bool ok = eq->dirichlet_vertex_value(sides[s], s == 0 ? 0 : s == 1 ? nx : k, s == 2 ? 0 : s == 3 ? ny : k, 0.0, &u);
      if (!ok) has[0] = 0.0; else line[k] = u;
    }
    h5lite::write_npy(out + "/bc_line" + std::to_string(s) + ".npy", line.data(), {n});
    h5lite::write_npy(out + "/bc_has" + std::to_string(s) + ".npy", has.data(), {1});
  }
  h5lite::write_npy(out + "/bc_types.npy", types.data(), {4});
  std::cout << "dt " << m.dt() << " static " << eq->is_static() << std::endl;
  return 0;
}
