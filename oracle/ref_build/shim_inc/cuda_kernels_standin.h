// Host stand-ins for the CUDA pieces RFW/backends/CUDART/src/Kernels.cu uses beyond shims/cuda_runtime.h, so the
// reference's own kernels can be compiled and executed thread by thread on the CPU (TEST INFRASTRUCTURE, OUR code):
// launch geometry as thread-local variables, symbol copies as memcpy, the output surface as a host array, atomics
// as plain sequential operations (the kernels use no shared memory, no barriers and no warp intrinsics).
#pragma once
#include <cuda_runtime.h> // shims/
#include <algorithm>
#include <cmath>
#include <cstddef>
#include <cstring>

#define __launch_bounds__(x, y)
struct uint3
{
	unsigned x, y, z;
};
struct dim3
{
	unsigned x, y, z;
	dim3(unsigned a = 1, unsigned b = 1, unsigned c = 1) : x(a), y(b), z(c) {}
};
struct float4
{
	float x, y, z, w;
};
inline thread_local uint3 threadIdx, blockIdx;
inline thread_local dim3 blockDim, gridDim;
typedef int cudaError;
typedef int cudaError_t;
constexpr int cudaSuccess = 0;
inline cudaError cudaGetLastError() { return cudaSuccess; }
template <typename T> inline cudaError cudaMemcpyToSymbol(T &symbol, const void *src, size_t n)
{
	std::memcpy(&symbol, src, std::min(n, sizeof(T))); // (setStride passes sizeof(void*) for a uint)
	return cudaSuccess;
}
constexpr int cudaSurfaceType2D = 2;
enum cudaSurfaceBoundaryMode
{
	cudaBoundaryModeZero,
	cudaBoundaryModeClamp,
	cudaBoundaryModeTrap
};
struct surfaceReference
{
};
template <typename T, int D> struct surface : surfaceReference
{
	float *pixels = nullptr; // RGBA32F rows of `width` pixels
	unsigned width = 0;
};
template <typename T, typename S> inline void surf2Dwrite(T value, S &surf, size_t x_bytes, size_t y, cudaSurfaceBoundaryMode)
{
	if (surf.pixels)
		std::memcpy(reinterpret_cast<char *>(surf.pixels) + (y * surf.width * 16 + x_bytes), &value, sizeof(T));
}
template <typename S> inline cudaError cudaGetSurfaceReference(const surfaceReference **ref, const S *s)
{
	*ref = s;
	return cudaSuccess;
}
template <typename T, typename B> inline T atomicAdd(T *p, B v)
{
	const T old = *p;
	*p = T(old + v);
	return old;
}
inline void __sincosf(float a, float *s, float *c) { *s = std::sin(a), *c = std::cos(a); }

// kernel<<<grid, block>>>(args) of the extracted source becomes RFW_LAUNCH(kernel, grid, block, args): every thread in turn
#define RFW_LAUNCH(KERNEL, GRID, BLOCK, ...)                                                                             \
	do                                                                                                                   \
	{                                                                                                                    \
		const dim3 g_ = dim3(GRID), b_ = dim3(BLOCK);                                                                    \
		::gridDim = g_, ::blockDim = b_;                                                                                 \
		for (unsigned bz_ = 0; bz_ < g_.z; bz_++)                                                                        \
			for (unsigned by_ = 0; by_ < g_.y; by_++)                                                                    \
				for (unsigned bx_ = 0; bx_ < g_.x; bx_++)                                                                \
					for (unsigned tz_ = 0; tz_ < b_.z; tz_++)                                                            \
						for (unsigned ty_ = 0; ty_ < b_.y; ty_++)                                                        \
							for (unsigned tx_ = 0; tx_ < b_.x; tx_++)                                                    \
							{                                                                                            \
								::blockIdx = uint3{bx_, by_, bz_}, ::threadIdx = uint3{tx_, ty_, tz_};                       \
								KERNEL(__VA_ARGS__);                                                                     \
							}                                                                                            \
	} while (0)
