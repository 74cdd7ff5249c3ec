// Diffusion-coefficient table of the Albert & Young cases: the reference's
// Albert_Young_IO interface (source/Cases/Albert_Young_IO.h:19-51) reading
// D/<dID>.h5 (/alpha0 [deg], /E [MeV], /Daa, /Dap, /Dpp) through h5lite.
#ifndef SY2D_HOST_ALBERT_YOUNG_IO_H_
#define SY2D_HOST_ALBERT_YOUNG_IO_H_

#include <string>

#include "Parameters.h"
#include "common.h"

class Albert_Young_IO {
 public:
  explicit Albert_Young_IO(const Parameters& paras);
  explicit Albert_Young_IO(const std::string& h5_file);

  Xarray1d x_D;  // alpha0 in rad
  Xarray1d y_D;  // E in MeV
  Xtensor2d Dxx_raw, Dxy_raw, Dyy_raw;

  std::size_t nx_D() const { return nx_D_; }
  std::size_t ny_D() const { return ny_D_; }
  double xmin_D() const { return xmin_D_; }
  double xmax_D() const { return xmax_D_; }
  double ymin_D() const { return ymin_D_; }
  double ymax_D() const { return ymax_D_; }
  void update(double) {}

 private:
  std::size_t nx_D_ = 0, ny_D_ = 0;
  double xmin_D_ = 0, xmax_D_ = 0, ymin_D_ = 0, ymax_D_ = 0;
  void read_D(const std::string& file);
};

#endif
