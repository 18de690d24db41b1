// stand-in for <cuda_fp16.h>: the reference's structs only need a 16-bit `half` that converts to float;
// the vendored half_float library (external/half2.1.0, part of the reference tree) provides it.
#pragma once
#include <half.hpp>
using half_float::half;
