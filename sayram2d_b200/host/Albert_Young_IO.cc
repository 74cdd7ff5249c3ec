#include "Albert_Young_IO.h"

#include "h5lite.h"

Albert_Young_IO::Albert_Young_IO(const Parameters& paras) { read_D("D/" + paras.dID() + ".h5"); }
Albert_Young_IO::Albert_Young_IO(const std::string& h5_file) { read_D(h5_file); }

void Albert_Young_IO::read_D(const std::string& file) {
  const h5lite::File h5(file);
  x_D = Xarray1d(h5.read("/alpha0")) * gPI / 180;  // degrees -> rad (Albert_Young_IO.cc:22)
  y_D = Xarray1d(h5.read("/E"));
  nx_D_ = x_D.size();
  ny_D_ = y_D.size();
  xmin_D_ = x_D[0];
  xmax_D_ = x_D[nx_D_ - 1];
  ymin_D_ = y_D[0];
  ymax_D_ = y_D[ny_D_ - 1];
  auto table = [&](const char* name, Xtensor2d& out) {
    std::vector<std::size_t> shape;
    const std::vector<double> v = h5.read(name, &shape);
    if (shape.size() != 2 || shape[0] != nx_D_ || shape[1] != ny_D_) throw std::runtime_error(std::string("D table: bad shape of ") + name);
    out.resize({shape[0], shape[1]});
    std::copy(v.begin(), v.end(), out.data());
  };
  table("/Daa", Dxx_raw);
  table("/Dap", Dxy_raw);
  table("/Dpp", Dyy_raw);
}
