// Boundary identifiers and condition types (same enumerators and order as the
// reference's source/BCTypes.h:13,16; the integer values are the SY2D_* constants
// of include/sayram2d.h).
#ifndef SY2D_HOST_BCTYPES_H_
#define SY2D_HOST_BCTYPES_H_
enum class BoundaryID { XMIN = 0, XMAX = 1, YMIN = 2, YMAX = 3 };
enum class BCType { Dirichlet = 0, ZeroFlux = 1 };
#endif
