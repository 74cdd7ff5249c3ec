// Translation unit of the drop-in demonstration (oracle/Makefile, target "dropin"):
// the GPU-backed Solver.h is included first, then the reference's main.cc verbatim from
// where it lies.  Both Solver.h files use the include guard SOLVER_H, so main.cc's own
// `#include "Solver.h"` (which would find source/Solver.h next to it) becomes a no-op.
#include "Solver.h"  // sayram2d_b200/dropin/Solver.h (first on the include path)
#include REF_MAIN
