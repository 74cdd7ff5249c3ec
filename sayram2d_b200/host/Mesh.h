// Cell-centred rectilinear mesh: the reference's Mesh interface (source/Mesh.h:33-96).
// The reference stores 4*nx*ny Edge objects and neighbour records; on this mesh they
// are closed forms of (i, j), so they are computed on demand here and the GPU path
// keeps only the 1-D edge arrays.
#ifndef SY2D_HOST_MESH_H_
#define SY2D_HOST_MESH_H_

#include <array>
#include <cstddef>
#include <vector>

#include "Grid2D.h"

struct Ind {
  std::size_t i, j;
};
struct VtxInd {
  std::size_t i, j;
};
enum class Direction { XPOS, XNEG, YPOS, YNEG };

struct Edge {  // face of a cell: two vertices (A, B), their vertex indices, length, outward unit normal
  std::array<std::array<double, 2>, 2> v;
  std::array<VtxInd, 2> vind;
  double length;
  std::array<double, 2> n;
  Direction dir;
};

class Mesh {
 public:
  Mesh(const Grid2D& grid, double dt);

  const std::vector<double>& x() const { return x_; }
  const std::vector<double>& y() const { return y_; }
  double x(std::size_t i) const { return x_[i]; }
  double y(std::size_t j) const { return y_[j]; }
  std::size_t nx() const { return nx_; }
  std::size_t ny() const { return ny_; }
  double x_edge(std::size_t i) const { return xe_[i]; }
  double y_edge(std::size_t j) const { return ye_[j]; }
  const std::vector<double>& x_edges() const { return xe_; }
  const std::vector<double>& y_edges() const { return ye_; }
  double dx(std::size_t i) const { return dx_[i]; }
  double dy(std::size_t j) const { return dy_[j]; }
  double dt() const { return dt_; }
  double cell_area_dt(const Ind& c) const { return dx(c.i) * dy(c.j) / dt(); }
  std::size_t flatten_cell_index(const Ind& c) const { return c.j * nx_ + c.i; }  // the reference's matrix numbering

  // neighbour numbering: 0 = im (west), 1 = jp (north), 2 = ip (east), 3 = jm (south)
  std::size_t nnbrs() const { return 4; }
  int inbr_im() const { return 0; }
  int inbr_jp() const { return 1; }
  int inbr_ip() const { return 2; }
  int inbr_jm() const { return 3; }
  int rinbr(int inbr) const { return (inbr + 2) % 4; }
  bool get_nbr_ind(const Ind& c, int inbr, Ind* out) const;
  void get_nbr_edge(const Ind& c, int inbr, Edge* out) const;

 private:
  std::size_t nx_, ny_;
  double dt_;
  std::vector<double> xe_, ye_, dx_, dy_, x_, y_;
};

#endif
