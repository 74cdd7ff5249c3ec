#include "Mesh.h"

#include <cmath>
#include <stdexcept>
#include <string>

Mesh::Mesh(const Grid2D& grid, double dt)
    : nx_(grid.nx()), ny_(grid.ny()), dt_(dt), xe_(grid.x_edges), ye_(grid.y_edges), dx_(nx_), dy_(ny_), x_(nx_), y_(ny_) {
  auto build = [](const std::vector<double>& e, std::vector<double>& w, std::vector<double>& c, const char* what) {
    for (std::size_t k = 0; k < w.size(); ++k) {
      w[k] = e[k + 1] - e[k];
      if (!(w[k] > 0.0)) throw std::runtime_error(std::string("Mesh: non-positive ") + what + std::to_string(k));
      c[k] = 0.5 * (e[k] + e[k + 1]);
    }
  };
  if (xe_.size() != nx_ + 1) throw std::runtime_error("Mesh: x_edges size mismatch.");
  if (ye_.size() != ny_ + 1) throw std::runtime_error("Mesh: y_edges size mismatch.");
  build(xe_, dx_, x_, "dx at i=");
  build(ye_, dy_, y_, "dy at j=");
}

bool Mesh::get_nbr_ind(const Ind& c, int inbr, Ind* out) const {
  switch (inbr) {
    case 0: if (c.i == 0) return false; *out = {c.i - 1, c.j}; return true;
    case 1: if (c.j + 1 >= ny_) return false; *out = {c.i, c.j + 1}; return true;
    case 2: if (c.i + 1 >= nx_) return false; *out = {c.i + 1, c.j}; return true;
    case 3: if (c.j == 0) return false; *out = {c.i, c.j - 1}; return true;
    default: return false;
  }
}

// Vertex order (A, B) per face as in the reference (source/Mesh.cc:94-149):
// W: (NW, SW)   N: (NE, NW)   E: (SE, NE)   S: (SW, SE)
void Mesh::get_nbr_edge(const Ind& c, int inbr, Edge* e) const {
  const std::size_t i = c.i, j = c.j;
  VtxInd a{}, b{};
  switch (inbr) {
    case 0: a = {i, j + 1}; b = {i, j}; e->dir = Direction::XNEG; e->n = {-1.0, 0.0}; break;
    case 1: a = {i + 1, j + 1}; b = {i, j + 1}; e->dir = Direction::YPOS; e->n = {0.0, 1.0}; break;
    case 2: a = {i + 1, j}; b = {i + 1, j + 1}; e->dir = Direction::XPOS; e->n = {1.0, 0.0}; break;
    default: a = {i, j}; b = {i + 1, j}; e->dir = Direction::YNEG; e->n = {0.0, -1.0}; break;
  }
  e->vind = {a, b};
  e->v[0] = {xe_[a.i], ye_[a.j]};
  e->v[1] = {xe_[b.i], ye_[b.j]};
  e->length = std::hypot(e->v[1][0] - e->v[0][0], e->v[1][1] - e->v[0][1]);
}
