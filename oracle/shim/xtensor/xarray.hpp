// Shim: forwards to the xtensor look-alike used to compile the reference unmodified.
#pragma once
#include "../xt_shim.hpp"
