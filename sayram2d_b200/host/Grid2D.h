// Rectilinear grid given by its edge coordinates (the reference's Grid2D interface,
// source/Grid2D.h:16-68, including the validation messages of :44-67).
#ifndef SY2D_HOST_GRID2D_H_
#define SY2D_HOST_GRID2D_H_

#include <stdexcept>
#include <string>
#include <vector>

class Grid2D {
 public:
  std::vector<double> x_edges;  // nx + 1
  std::vector<double> y_edges;  // ny + 1

  Grid2D() = default;
  Grid2D(std::vector<double> xe, std::vector<double> ye) : x_edges(std::move(xe)), y_edges(std::move(ye)) { validate(); }

  std::size_t nx() const { return x_edges.size() - 1; }
  std::size_t ny() const { return y_edges.size() - 1; }
  double x_min() const { return x_edges.front(); }
  double x_max() const { return x_edges.back(); }
  double y_min() const { return y_edges.front(); }
  double y_max() const { return y_edges.back(); }

  void validate() const {
    check(x_edges, "x_edges", "i");
    check(y_edges, "y_edges", "j");
  }

 private:
  static void check(const std::vector<double>& e, const char* name, const char* idx) {
    if (e.size() < 2) throw std::runtime_error(std::string("Grid2D: ") + name + " must have size >= 2.");
    for (std::size_t k = 0; k + 1 < e.size(); ++k)
      if (!(e[k + 1] > e[k]))
        throw std::runtime_error(std::string("Grid2D: ") + name + " must be strictly increasing at " + idx + "=" + std::to_string(k));
  }
};

#endif
