// common.cuh -- shared device helpers for the morsi sm_100a kernels.
#pragma once
#include <cuda_runtime.h>
#include <math_constants.h>
#include <stdint.h>
#include <string.h>

__host__ __device__ __forceinline__ float morsi_bits_to_float(uint32_t u)
{
#ifdef __CUDA_ARCH__
	return __uint_as_float(u);
#else
	float f; memcpy(&f, &u, 4); return f;
#endif
}

// A band of one plane: sample (i,j) of the plane lives at p[(j-row0)*w + i];
// rows outside [0,h) do not exist (src/morsi.c:30-35), rows inside the image
// but outside the band are the caller's responsibility (never dereferenced by
// a correct launch).  pstride = floats between consecutive planes.
struct Band {
	const float *p;
	int row0;
	long long pstride;
};

// Final arithmetic of each operation (src/morsi.c:141-275).  a = erosion-side
// value, b = dilation-side value, x = the input pixel.  Written with explicit
// round-to-nearest intrinsics so nvcc never contracts to FMA: every step is
// rounded to float exactly like the reference's plain -O3 x86-64 build.
enum Epi {
	EPI_A = 0,    // y = a                      erosion, closing
	EPI_B,        // y = b                      dilation, opening
	EPI_B_SUB_A,  // y = b - a                  gradient        :165
	EPI_X_SUB_A,  // y = x - a                  igradient       :174
	EPI_B_SUB_X,  // y = b - x                  egradient       :183
	EPI_LAP,      // y = (a + b - 2x)/2         laplacian       :194
	EPI_ENH,      // y = x - lap                enhance         :204
	EPI_BLUR,     // y = x + lap                blur            :213
	EPI_A_SUB_B,  // y = a - b                  oscillation     :225 (closing - opening)
	EPI_X_SUB_B,  // y = x - b                  tophat          :234
	EPI_A_SUB_X,  // y = a - x                  bothat          :243
	EPI_IBLUR,    // y = (x + a)/2              iblur           :252
	EPI_EBLUR,    // y = (x + b)/2              eblur           :261
	EPI_CBLUR,    // y = 0.5x + 0.25a + 0.25b   cblur, in double :272
	EPI_AB        // two outputs: a and b (first stage of oscillation)
};

__device__ __forceinline__ float lap_of(float a, float b, float x)
{
	float s = __fadd_rn(a, b);
	float d = __fmul_rn(2.0f, x);
	return __fmul_rn(__fsub_rn(s, d), 0.5f);   // "/2" is exact scaling
}

template <int EPI>
__device__ __forceinline__ float epilogue(float a, float b, float x)
{
	if (EPI == EPI_A) return a;
	if (EPI == EPI_B) return b;
	if (EPI == EPI_B_SUB_A) return __fsub_rn(b, a);
	if (EPI == EPI_X_SUB_A) return __fsub_rn(x, a);
	if (EPI == EPI_B_SUB_X) return __fsub_rn(b, x);
	if (EPI == EPI_LAP) return lap_of(a, b, x);
	if (EPI == EPI_ENH) return __fsub_rn(x, lap_of(a, b, x));
	if (EPI == EPI_BLUR) return __fadd_rn(x, lap_of(a, b, x));
	if (EPI == EPI_A_SUB_B) return __fsub_rn(a, b);
	if (EPI == EPI_X_SUB_B) return __fsub_rn(x, b);
	if (EPI == EPI_A_SUB_X) return __fsub_rn(a, x);
	if (EPI == EPI_IBLUR) return __fmul_rn(__fadd_rn(x, a), 0.5f);
	if (EPI == EPI_EBLUR) return __fmul_rn(__fadd_rn(x, b), 0.5f);
	if (EPI == EPI_CBLUR)
		return (float)__dadd_rn(__dadd_rn(__dmul_rn(0.5, (double)x),
				__dmul_rn(0.25, (double)a)), __dmul_rn(0.25, (double)b));
	return a;
}

template <int EPI> struct EpiNeeds {
	static constexpr bool a = EPI == EPI_A || EPI == EPI_B_SUB_A || EPI == EPI_X_SUB_A ||
		EPI == EPI_LAP || EPI == EPI_ENH || EPI == EPI_BLUR || EPI == EPI_A_SUB_B ||
		EPI == EPI_A_SUB_X || EPI == EPI_IBLUR || EPI == EPI_CBLUR || EPI == EPI_AB;
	static constexpr bool b = EPI == EPI_B || EPI == EPI_B_SUB_A || EPI == EPI_B_SUB_X ||
		EPI == EPI_LAP || EPI == EPI_ENH || EPI == EPI_BLUR || EPI == EPI_A_SUB_B ||
		EPI == EPI_X_SUB_B || EPI == EPI_EBLUR || EPI == EPI_CBLUR || EPI == EPI_AB;
	static constexpr bool x = EPI == EPI_X_SUB_A || EPI == EPI_B_SUB_X || EPI == EPI_LAP ||
		EPI == EPI_ENH || EPI == EPI_BLUR || EPI == EPI_X_SUB_B || EPI == EPI_A_SUB_X ||
		EPI == EPI_IBLUR || EPI == EPI_EBLUR || EPI == EPI_CBLUR;
};

// Counter-based synthetic pixel; identical on host and device.
__host__ __device__ __forceinline__ uint32_t morsi_mix32(uint32_t h)
{
	h ^= h >> 16; h *= 0x85EBCA6Bu; h ^= h >> 13; h *= 0xC2B2AE35u; h ^= h >> 16;
	return h;
}
__host__ __device__ __forceinline__ float morsi_synth_value(uint32_t seed, uint32_t plane,
		uint32_t row, uint32_t col, int dist)
{
	uint32_t h = morsi_mix32(seed * 0x9E3779B1u + plane);
	h = morsi_mix32(h ^ (row * 0x27D4EB2Fu));
	h = morsi_mix32(h ^ (col * 0x165667B1u));
	if (dist == 1) return (float)(h >> 24);
	float v = (float)(h >> 8) * (1.0f / 16777216.0f);
	if (dist == 2) {
		uint32_t r = morsi_mix32(h ^ 0x5BD1E995u) % 1000u;
		if (r < 10) return morsi_bits_to_float(0x7FC00000u);
		if (r < 13) return morsi_bits_to_float(0x7F800000u);
		if (r < 15) return morsi_bits_to_float(0xFF800000u);
		if (r < 25) return 0.0f;
		if (r < 35) return morsi_bits_to_float(0x80000000u);
	}
	return v;
}
