// The loss-cone case lives next to Albert_Young (shared base); this header keeps the
// reference's include name (source/Cases/Albert_Young_LC.h).
#ifndef SY2D_HOST_ALBERT_YOUNG_LC_H_
#define SY2D_HOST_ALBERT_YOUNG_LC_H_
#include "Albert_Young.h"
#endif
