// Host-side 1-D geometry of the rectilinear mesh, computed once per context:
// cell widths/centres (Mesh.cc:36-65) and the vertex interpolation weights of
// Solver.cc:326-381, padded so that one bilinear formula covers interior
// vertices, boundary-line vertices and corners (see vertex_value()).
#pragma once
#include <vector>

namespace sy2d {

struct HostGeometry {
  std::vector<double> dx, dy, wxL, wxR, wyB, wyT;
};

inline HostGeometry make_host_geometry(int nx, int ny, const double* xe, const double* ye) {
  HostGeometry h;
  std::vector<double> x(nx), y(ny);
  h.dx.resize(nx); h.dy.resize(ny);
  h.wxL.resize(nx + 1); h.wxR.resize(nx + 1); h.wyB.resize(ny + 1); h.wyT.resize(ny + 1);
  for (int i = 0; i < nx; ++i) { h.dx[i] = xe[i + 1] - xe[i]; x[i] = 0.5 * (xe[i] + xe[i + 1]); }   // Mesh.cc:44-52
  for (int j = 0; j < ny; ++j) { h.dy[j] = ye[j + 1] - ye[j]; y[j] = 0.5 * (ye[j] + ye[j + 1]); }   // Mesh.cc:55-64
  h.wxL[0] = 0.0; h.wxR[0] = 1.0; h.wxL[nx] = 1.0; h.wxR[nx] = 0.0;
  for (int i = 1; i < nx; ++i) {                       // Solver.cc:337-341
    const double d = x[i] - x[i - 1];
    h.wxL[i] = (x[i] - xe[i]) / d;
    h.wxR[i] = (xe[i] - x[i - 1]) / d;
  }
  h.wyB[0] = 0.0; h.wyT[0] = 1.0; h.wyB[ny] = 1.0; h.wyT[ny] = 0.0;
  for (int j = 1; j < ny; ++j) {                       // Solver.cc:338-343
    const double d = y[j] - y[j - 1];
    h.wyB[j] = (y[j] - ye[j]) / d;
    h.wyT[j] = (ye[j] - y[j - 1]) / d;
  }
  return h;
}

}  // namespace sy2d
