// Shim: HighFive/xtensor-io look-alike (HDF5 read of contiguous f64 datasets, .npy dump).
#pragma once
#include "../xt_shim.hpp"
#include "../xt_io_shim.hpp"
