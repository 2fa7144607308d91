// k_exact.cuh -- order-preserving kernels: the semantics of src/morsi.c:56-139
// evaluated in element order, straight from global memory (L1/L2 cached).
//
// These are the catch-all path: any element list, any size, any values
// (signed zeros resolved exactly like glibc's fmin/fmax: the last element-order
// occurrence of the extremum wins).  The fast kernel families
// (k_small / k_tiled / k_rowrun / k_median) reorder the reduction and hand
// images that contain -0.0 back to these via the `gate` word.
#pragma once
#include "common.cuh"

struct ExactArgs {
	Band a_src;        // source of the erosion-side pass (may be unused)
	Band b_src;        // source of the dilation-side pass
	Band x_src;        // the original image, for the epilogue
	float *y;          // output rows [y_row0, y_row0+y_rows)
	float *y2;         // second output for EPI_AB (the max), same band
	long long y_pstride;
	int y_row0, y_rows;
	int w, h;          // full plane size
	const int2 *offs;  // effective offsets (dx-e[2], dy-e[3]) in element order
	int n;
	const int *gate;   // if non-NULL: run only when *gate != 0
};

__device__ __forceinline__ float band_pixel(const Band &s, int plane, int w, int h, int i, int j)
{
	if (i < 0 || i >= w || j < 0 || j >= h)
		return CUDART_NAN_F;
	return __ldg(s.p + plane * s.pstride + (long long)(j - s.row0) * w + i);
}

template <int EPI>
__global__ void __launch_bounds__(256) k_exact_minmax(ExactArgs p)
{
	if (p.gate && *p.gate == 0) return;
	const int i = blockIdx.x * 32 + threadIdx.x;
	const int plane = blockIdx.z;
	for (int jj = blockIdx.y * 8 + threadIdx.y; jj < p.y_rows; jj += gridDim.y * 8) {
		if (i >= p.w) return;
		const int j = p.y_row0 + jj;
		float a = CUDART_INF_F, b = -CUDART_INF_F;
		for (int k = 0; k < p.n; k++) {
			const int2 o = p.offs[k];
			if (EpiNeeds<EPI>::a) {
				float v = band_pixel(p.a_src, plane, p.w, p.h, i + o.x, j + o.y);
				a = (v <= a) ? v : a;      // fmin(a,v): ties and -0/+0 -> v, NaN v ignored
			}
			if (EpiNeeds<EPI>::b) {
				float v = band_pixel(p.b_src, plane, p.w, p.h, i + o.x, j + o.y);
				b = (v >= b) ? v : b;      // fmax(b,v)
			}
		}
		float x = 0.f;
		if (EpiNeeds<EPI>::x) x = band_pixel(p.x_src, plane, p.w, p.h, i, j);
		const long long o = plane * p.y_pstride + (long long)jj * p.w + i;
		if (EPI == EPI_AB) { p.y[o] = a; p.y2[o] = b; }
		else p.y[o] = epilogue<EPI>(a, b, x);
	}
}

// ---- median (src/morsi.c:91-120) ------------------------------------------
// Stable-sort semantics without sorting: an order-preserving integer key per
// finite neighbour (+0 and -0 share a key, as compare_floats says they are
// equal), a 32-step binary search for the key of sorted position k, and, when
// that key is zero, a scan in element order to pick the sign the stable sort
// would have left at that position.

__device__ __forceinline__ uint32_t median_key(float v)
{
	uint32_t u = __float_as_uint(v);
	if (u == 0x80000000u) u = 0;                    // -0 == +0
	return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float median_unkey(uint32_t k)
{
	return __uint_as_float((k & 0x80000000u) ? (k & 0x7FFFFFFFu) : ~k);
}

struct MedianArgs {
	Band x_src;
	float *y;
	long long y_pstride;
	int y_row0, y_rows;
	int w, h;
	const int2 *offs;
	int n;
	const int *gate;
};

// value at stable-sorted position k among the finite neighbours of (i,j)
__device__ float median_select(const MedianArgs &p, int plane, int i, int j, int k)
{
	uint32_t prefix = 0;
	// invariant: the wanted key has the high bits of `prefix`; below = number
	// of finite keys smaller than any key with that prefix
	int below = 0;
	for (int bit = 31; bit >= 0; bit--) {
		const uint32_t mask_hi = bit == 31 ? 0u : (0xFFFFFFFFu << (bit + 1));
		int zeros = 0;   // keys matching prefix with this bit clear
		for (int m = 0; m < p.n; m++) {
			const int2 o = p.offs[m];
			float v = band_pixel(p.x_src, plane, p.w, p.h, i + o.x, j + o.y);
			if (!isfinite(v)) continue;
			uint32_t key = median_key(v);
			zeros += ((key & mask_hi) == prefix) && !((key >> bit) & 1u);
		}
		if (k < below + zeros) {
			// stays in the 0 branch
		} else {
			below += zeros;
			prefix |= 1u << bit;
		}
	}
	if (prefix != 0x80000000u)      // not a zero: the value is its own identity
		return median_unkey(prefix);
	// zero: the (k-below)-th zero in element order carries the sign
	int want = k - below, seen = 0;
	for (int m = 0; m < p.n; m++) {
		const int2 o = p.offs[m];
		float v = band_pixel(p.x_src, plane, p.w, p.h, i + o.x, j + o.y);
		if (v == 0.0f) {
			if (seen == want) return v;
			seen++;
		}
	}
	return 0.0f;
}

__global__ void __launch_bounds__(256) k_exact_median(MedianArgs p)
{
	if (p.gate && *p.gate == 0) return;
	const int i = blockIdx.x * 32 + threadIdx.x;
	const int plane = blockIdx.z;
	for (int jj = blockIdx.y * 8 + threadIdx.y; jj < p.y_rows; jj += gridDim.y * 8) {
		if (i >= p.w) return;
		const int j = p.y_row0 + jj;
		int cnt = 0;
		float first = 0.f, second = 0.f;   // gather order, for n <= 2
		for (int m = 0; m < p.n; m++) {
			const int2 o = p.offs[m];
			float v = band_pixel(p.x_src, plane, p.w, p.h, i + o.x, j + o.y);
			if (isfinite(v)) {
				if (cnt == 0) first = v;
				if (cnt == 1) second = v;
				cnt++;
			}
		}
		float r;
		if (cnt < 1) r = CUDART_NAN_F;                                   // :93
		else if (cnt == 1) r = first;                                     // :94
		else if (cnt == 2) r = __fmul_rn(__fadd_rn(first, second), 0.5f); // :95
		else if (cnt & 1) r = median_select(p, plane, i, j, cnt / 2);     // :100
		else r = __fmul_rn(__fadd_rn(median_select(p, plane, i, j, cnt / 2),
					median_select(p, plane, i, j, cnt / 2 + 1)), 0.5f); // :98
		p.y[plane * p.y_pstride + (long long)jj * p.w + i] = r;
	}
}

// ---- rank (src/morsi.c:122-139) -------------------------------------------
__global__ void __launch_bounds__(256) k_exact_rank(MedianArgs p)
{
	const int i = blockIdx.x * 32 + threadIdx.x;
	const int plane = blockIdx.z;
	for (int jj = blockIdx.y * 8 + threadIdx.y; jj < p.y_rows; jj += gridDim.y * 8) {
		if (i >= p.w) return;
		const int j = p.y_row0 + jj;
		const float u = band_pixel(p.x_src, plane, p.w, p.h, i, j);
		int cnt = 0;
		for (int m = 0; m < p.n; m++) {
			const int2 o = p.offs[m];
			float v = band_pixel(p.x_src, plane, p.w, p.h, i + o.x, j + o.y);
			cnt += isfinite(v) && (v < u);
		}
		p.y[plane * p.y_pstride + (long long)jj * p.w + i] = (float)cnt;
	}
}
