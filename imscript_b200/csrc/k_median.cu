// k_median.cu -- median filter (src/morsi.c:91-120) for small row-run elements
// (cross, square, disk2.5 ... disk5): BASELINE config C3 (disk5 median, n = 69).
//
// A CTA stages its 32x8 output tile plus halo in shared memory (out-of-image
// samples as NaN) and classifies it while loading: does it hold a non-finite
// value, a -0.0?  Pixels whose whole window is inside the image, in a tile of
// finite values, take the fast path: the n neighbours go to registers and a
// "forgetful selection" network (keep n/2+2 candidates; repeatedly find their
// minimum and maximum with compare-exchanges, forget both, admit the next
// neighbour) leaves the exact median -- the same value the reference obtains
// from qsort + a[n/2] (n is odd for these elements), with about n^2/2.4
// FMNMX instead of a libc sort per pixel.  All other pixels (image border:
// fewer than n neighbours, or tiles with NaN/Inf: the reference drops them,
// src/morsi.c:115) take a counting selection over the tile that implements
// the variable-count rules of src/morsi.c:91-101 (n<1, n==1, n==2, odd, even).
// Signed zeros are order-dependent in the reference (stable sort): a -0.0 in
// the data raises *flag and the order-preserving kernel re-runs the job.
#include "dispatch.cuh"
#include "shapes.cuh"

struct MedianFastArgs {
	Band x;
	float *y;
	long long y_pstride;
	int y_row0, y_rows;
	int x_rows;         // rows held by x (rows of the image outside the band are never needed)
	int w, h;
	int *flag;
};

template <class S> struct Win {
	static constexpr int R = S::R;
	static constexpr int RX = S::hw(R);
	__host__ __device__ static constexpr int count() { int n = 0; for (int i = 0; i <= 2 * R; i++) n += 2 * S::hw(i) + 1; return n; }
	static constexpr int N = count();
	// neighbour #i in row-run order (rows top to bottom, left to right)
	__host__ __device__ static constexpr int dy_of(int i) { int dy = -R; while (i >= 2 * S::hw(dy + R) + 1) { i -= 2 * S::hw(dy + R) + 1; dy++; } return dy; }
	__host__ __device__ static constexpr int dx_of(int i) { int dy = -R; while (i >= 2 * S::hw(dy + R) + 1) { i -= 2 * S::hw(dy + R) + 1; dy++; } return i - S::hw(dy + R); }
};

__device__ __forceinline__ void cex(float &a, float &b)    // a <- min, b <- max
{
	const float lo = fminf(a, b), hi = fmaxf(a, b);
	a = lo; b = hi;
}

// minimum of v[lo..hi] to v[lo], maximum to v[hi] (all indices compile time)
template <int NV, int lo, int hi>
__device__ __forceinline__ void minmax_ends(float (&v)[NV])
{
	constexpr int s = hi - lo + 1;
	// pair the two halves: afterwards the minimum is in the left part, the maximum in the right
#pragma unroll
	for (int i = 0; i < s / 2; i++) cex(v[lo + i], v[hi - i]);
	constexpr int nl = (s + 1) / 2;      // left part incl. the unpaired middle
	// tournament towards v[lo]
#pragma unroll
	for (int stride = 1; stride < nl; stride *= 2)
#pragma unroll
		for (int i = 0; i + stride < nl; i += 2 * stride) cex(v[lo + i], v[lo + i + stride]);
	// tournament towards v[hi] (mirror); the middle element takes part in both
#pragma unroll
	for (int stride = 1; stride < nl; stride *= 2)
#pragma unroll
		for (int i = 0; i + stride < nl; i += 2 * stride) cex(v[hi - i - stride], v[hi - i]);
}

// counting selection on canonical keys, reading the window from the tile;
// implements the general rules for windows with missing / non-finite values
__device__ __forceinline__ uint32_t mkey(float v)
{
	uint32_t u = __float_as_uint(v);
	if (u == 0x80000000u) u = 0;
	return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float munkey(uint32_t k)
{
	return __uint_as_float((k & 0x80000000u) ? (k & 0x7FFFFFFFu) : ~k);
}

template <class S>
__device__ float tile_select(const float *centre, int pitch, int k)
{
	constexpr int R = S::R;
	uint32_t prefix = 0;
	int below = 0;
	for (int bit = 31; bit >= 0; bit--) {
		const uint32_t mask_hi = bit == 31 ? 0u : (0xFFFFFFFFu << (bit + 1));
		int zeros = 0;
		for (int dy = -R; dy <= R; dy++) {
			const int hw = S::hw(dy + R);
			for (int dx = -hw; dx <= hw; dx++) {
				const float v = centre[dy * pitch + dx];
				if (!isfinite(v)) continue;
				const uint32_t key = mkey(v);
				zeros += ((key & mask_hi) == prefix) && !((key >> bit) & 1u);
			}
		}
		if (k >= below + zeros) { below += zeros; prefix |= 1u << bit; }
	}
	return munkey(prefix);
}

template <class S>
__device__ __noinline__ float tile_median_general(const float *centre, int pitch)
{
	constexpr int R = S::R;
	int cnt = 0;
	float first = 0.f, second = 0.f;
	for (int dy = -R; dy <= R; dy++) {
		const int hw = S::hw(dy + R);
		for (int dx = -hw; dx <= hw; dx++) {
			const float v = centre[dy * pitch + dx];
			if (isfinite(v)) {
				if (cnt == 0) first = v;
				if (cnt == 1) second = v;
				cnt++;
			}
		}
	}
	if (cnt < 1) return CUDART_NAN_F;                                     // src/morsi.c:93
	if (cnt == 1) return first;                                            // :94
	if (cnt == 2) return __fmul_rn(__fadd_rn(first, second), 0.5f);        // :95 (addition commutes)
	if (cnt & 1) return tile_select<S>(centre, pitch, cnt / 2);            // :100
	return __fmul_rn(__fadd_rn(tile_select<S>(centre, pitch, cnt / 2),
				tile_select<S>(centre, pitch, cnt / 2 + 1)), 0.5f);        // :98
}

// Forgetful selection, neighbour by neighbour (template recursion so that every
// register index is a compile-time constant).  The first SZ = N/2+2 neighbours
// fill the candidate set; each later one replaces the maximum after the
// minimum and maximum of the live set v[I-SZ .. SZ-1] have been found and
// forgotten.  Three candidates remain at the end; the middle one is the median.
template <class S, int PW, int I, int SZ>
__device__ __forceinline__ void forgetful_feed(float (&v)[SZ], const float *centre)
{
	if constexpr (I < Win<S>::N) {
		constexpr int dy = Win<S>::dy_of(I), dx = Win<S>::dx_of(I);
		const float nv = centre[dy * PW + dx];
		if constexpr (I < SZ) {
			v[I] = nv;
		} else {
			minmax_ends<SZ, I - SZ, SZ - 1>(v);
			v[SZ - 1] = nv;
		}
		forgetful_feed<S, PW, I + 1, SZ>(v, centre);
	}
}

template <class S>
__global__ void __launch_bounds__(256) k_median_fast(MedianFastArgs p)
{
	constexpr int R = Win<S>::R, RX = Win<S>::RX, N = Win<S>::N;
	constexpr int TX = 32, TY = 8;
	constexpr int PW = TX + 2 * RX + 1;             // +1: odd pitch, rows land on different banks
	constexpr int PH = TY + 2 * R;
	__shared__ float tile[PH * PW];

	const int tx = threadIdx.x, ty = threadIdx.y, tid = ty * TX + tx;
	const int plane = blockIdx.z;
	const int bx = blockIdx.x * TX, by = p.y_row0 + blockIdx.y * TY;
	const float *xp = p.x.p + plane * p.x.pstride;

	// ---- stage the tile, classify it ------------------------------------------
	bool nonfinite = false, negzero = false;
	for (int t = tid; t < PH * (TX + 2 * RX); t += 256) {
		const int r = t / (TX + 2 * RX), c = t - r * (TX + 2 * RX);
		const int gy = by - R + r, gx = bx - RX + c;
		float v = CUDART_NAN_F;
		if (gx >= 0 && gx < p.w && gy >= 0 && gy < p.h && gy >= p.x.row0 && gy < p.x.row0 + p.x_rows) {
			v = __ldg(xp + (long long)(gy - p.x.row0) * p.w + gx);
			nonfinite |= !isfinite(v);
			negzero |= __float_as_uint(v) == 0x80000000u;
		}
		tile[r * PW + c] = v;
	}
	const int dirty = __syncthreads_or(nonfinite);
	if (negzero) atomicOr(p.flag, 1);

	const int gx = bx + tx, gy = by + ty;
	if (gx >= p.w || gy >= p.y_row0 + p.y_rows) return;
	const float *centre = tile + (ty + R) * PW + tx + RX;
	float result;
	const bool inside = gx - RX >= 0 && gx + RX < p.w && gy - R >= 0 && gy + R < p.h;
	if (!dirty && inside) {
		// ---- forgetful selection of the median of N values (N odd) -------------
		constexpr int SZ = N / 2 + 2;                // candidates kept
		float v[SZ];
		forgetful_feed<S, PW, 0, SZ>(v, centre);
		if constexpr (N >= 3) {
			minmax_ends<SZ, SZ - 3, SZ - 1>(v);      // N - SZ values were forgotten on each side
			result = v[SZ - 2];
		} else {
			result = v[0];
		}
	} else {
		result = tile_median_general<S>(centre, PW);
	}
	p.y[plane * p.y_pstride + (long long)(gy - p.y_row0) * p.w + gx] = result;
}

// ---- host side --------------------------------------------------------------------
template <int ID>
static bool med_shape_matches(const DevElement *de)
{
	const RowRunPlan &rr = de->rowrun;
	if (!rr.ok || rr.reach != Shape<ID>::R || de->info.has_duplicates) return false;
	for (int i = 0; i <= 2 * rr.reach; i++)
		if (rr.hw[i] != Shape<ID>::hw(i)) return false;
	return de->n == Win<Shape<ID>>::N;
}

template <int ID>
static int launch_median(const MedianFastArgs &a, int planes, cudaStream_t st)
{
	static_assert(Win<Shape<ID>>::N % 2 == 1, "fast median needs an odd element count");
	dim3 grid((a.w + 31) / 32, (a.y_rows + 7) / 8, planes);
	k_median_fast<Shape<ID>><<<grid, dim3(32, 8), 0, st>>>(a);
	morsi_count_launch(1);
	MORSI_CU(cudaGetLastError());
	return MORSI_OK;
}

int morsi_run_median(MorsiCtx *, const DevElement *de, const MorsiJob &job, int *flag, int *handled)
{
	*handled = 0;
	if (job.op != MORSI_MEDIAN) return MORSI_OK;
	if ((job.y_rows + 7) / 8 > 65535) return MORSI_OK;
	MedianFastArgs a;
	a.x = Band{job.x, job.x_row0, job.x_pstride};
	a.y = job.y; a.y_pstride = job.y_pstride; a.y_row0 = job.y_row0; a.y_rows = job.y_rows;
	a.w = job.w; a.h = job.h; a.flag = flag; a.x_rows = job.x_rows;
	int rc = MORSI_OK;
#define T(ID) if (!*handled && med_shape_matches<ID>(de)) { rc = launch_median<ID>(a, job.planes, job.stream); *handled = 1; }
	T(14) T(15) T(0) T(1) T(2) T(3) T(4) T(5)
#undef T
	return rc;
}
