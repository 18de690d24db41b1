// host stand-in for <cuda_runtime.h>: qualifiers vanish, bit casts and the two CUDA vector helpers the reference
// headers use are provided.  OUR code.
#pragma once
#include <cmath>
#include <cstring>
#define __device__
#define __host__
#define __constant__
#define __global__
#ifndef __CUDACC_SHIM_TYPES__
#define __CUDACC_SHIM_TYPES__
struct float2
{
	float x, y;
};
static inline float2 make_float2(float x, float y) { return float2{x, y}; }
static inline unsigned int __float_as_uint(float f)
{
	unsigned int u;
	std::memcpy(&u, &f, 4);
	return u;
}
static inline int __float_as_int(float f)
{
	int u;
	std::memcpy(&u, &f, 4);
	return u;
}
static inline float __uint_as_float(unsigned int u)
{
	float f;
	std::memcpy(&f, &u, 4);
	return f;
}
static inline float __int_as_float(int u)
{
	float f;
	std::memcpy(&f, &u, 4);
	return f;
}
#endif
