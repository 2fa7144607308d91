// k_line.cuh -- erosion / dilation family by LINE elements (hrecR: one row,
// vrecR: one column; src/morsi.c:351-383) with the van Herk / Gil-Werman
// scheme: 3 compares per sample whatever the length.
//
// A CTA stages NL lines (rows for hrec, columns for vrec) of (nb+1)*L samples
// along the element's axis in shared memory (L = element length, samples
// outside the image are NaN = absent).  The samples of a line are cut into
// blocks of L; one thread per (line, block) writes the running extremum from
// the block's end backwards (S) and from its start forwards (P, in place).
// Every window of L consecutive samples straddles at most one block border, so
// output o of a line is ext(S[o], P[o+L-1]).  min.f32/max.f32 ignore NaN like
// fmin/fmax; a -0.0 in the tile raises *flag for the order-preserving re-run
// (SURVEY.md 9.1-Z).  Two-stage operations run as two launches with the
// temporary in the workspace, like k_tiled.
#pragma once
#include "k_exact.cuh"

struct LineGeom {
	int vertical;      // 0: the element lies along x, 1: along y
	int lo;            // window of output o along the axis: [o + lo, o + lo + L - 1]
	int perp;          // offset of the line across the axis (0 for the reference's hrec / vrec)
	int L;
	int nb;            // output blocks per CTA along the axis (nb * L outputs per line)
	int nl;            // lines per CTA: 16 rows (horizontal) / 32 columns (vertical)
	int sl, sp;        // shared-memory strides of line / position
	int two_sources;   // the two reduction sides read different images
};

#define LINE_MAXOUT 40     // outputs per thread (256 threads): nl * nb * L <= 256 * LINE_MAXOUT

__device__ __forceinline__ int line_idx(const LineGeom &g, int line, int pos) { return line * g.sl + pos * g.sp; }

// stage the tile of `src` whose first output is (ox0, oy0) (global), then build S (in S) and P (in T)
// gy_lo .. gy_hi: the rows the stored outputs of this launch can need (a band
// holds exactly those; the tile of the last CTA reaches further and must not read there)
template <bool ISMAX>
__device__ __forceinline__ void line_scan(float *T, float *S, const LineGeom &g, const Band &src, int plane,
		int w, int h, int ox0, int oy0, int tid, bool &negzero, int gy_lo, int gy_hi)
{
	gy_lo = max(gy_lo, 0); gy_hi = min(gy_hi, h - 1);
	const int len = (g.nb + 1) * g.L;
	const float *sp = src.p + plane * src.pstride;
	const int total = g.nl * len;
	// consecutive threads -> consecutive global x; (line, pos) advance without divisions
	int line = g.vertical ? tid % g.nl : tid / len;
	int pos = g.vertical ? tid / g.nl : tid - line * len;
	const int dpos = g.vertical ? 256 / g.nl : 256;        // nl divides 256
	for (int t = tid; t < total; t += 256) {
		const int gx = g.vertical ? ox0 + line + g.perp : ox0 + g.lo + pos;
		const int gy = g.vertical ? oy0 + g.lo + pos : oy0 + line + g.perp;
		float v = CUDART_NAN_F;
		if (gx >= 0 && gx < w && gy >= gy_lo && gy <= gy_hi) {
			v = __ldg(sp + (long long)(gy - src.row0) * w + gx);
			negzero |= __float_as_uint(v) == 0x80000000u;
		}
		T[line_idx(g, line, pos)] = v;
		pos += dpos;
		if (!g.vertical) while (pos >= len) { pos -= len; line++; }
	}
	__syncthreads();
	const float init = ISMAX ? -CUDART_INF_F : CUDART_INF_F;
	for (int t = tid; t < g.nl * (g.nb + 1); t += 256) {
		const int line = t % g.nl, blk = t / g.nl;
		float s = init;
#pragma unroll 4
		for (int i = g.L - 1; i >= 0; i--) {
			const int k = line_idx(g, line, blk * g.L + i);
			s = ISMAX ? fmaxf(s, T[k]) : fminf(s, T[k]);
			S[k] = s;
		}
		float p = init;
#pragma unroll 4
		for (int i = 0; i < g.L; i++) {
			const int k = line_idx(g, line, blk * g.L + i);
			p = ISMAX ? fmaxf(p, T[k]) : fminf(p, T[k]);
			T[k] = p;
		}
	}
	__syncthreads();
}

// Outputs of a thread: #q is element q * 256 + tid of the CTA's nl x nout outputs,
// consecutive threads -> consecutive global x.  OutIter walks them without divisions
// (horizontal: nout > 256 by construction, so one wrap per step at most).
struct LineOutIter {
	int line, o, dline, dout, nout, wrap;
	__device__ __forceinline__ LineOutIter(const LineGeom &g, int tid)
	{
		nout = g.nb * g.L;
		if (g.vertical) { line = tid % g.nl; o = tid / g.nl; dout = 256 / g.nl; wrap = 0; }
		else { line = tid / nout; o = tid - line * nout; dout = 256; wrap = 1; }
	}
	__device__ __forceinline__ void next()
	{
		o += dout;
		if (wrap && o >= nout) { o -= nout; line++; }
	}
};

template <int EPI>
__global__ void __launch_bounds__(256) k_line_minmax(ExactArgs p, LineGeom g, int *flag)
{
	extern __shared__ float line_smem[];
	constexpr bool NA = EpiNeeds<EPI>::a, NB = EpiNeeds<EPI>::b;
	const int tid = threadIdx.x;
	const int plane = blockIdx.z;
	const int len = (g.nb + 1) * g.L;
	const int tile_floats = g.vertical ? len * g.sp : g.nl * g.sl;
	float *T = line_smem, *S = line_smem + tile_floats;
	const int nout = g.nb * g.L;                         // outputs per line
	// first output of this CTA (global coordinates; rows relative to the image)
	const int ox0 = g.vertical ? blockIdx.x * g.nl : blockIdx.x * nout;
	const int oy0 = p.y_row0 + (g.vertical ? blockIdx.y * nout : blockIdx.y * g.nl);
	const int total_out = g.nl * nout;

	bool negzero = false;
	const int need_lo = p.y_row0 + (g.vertical ? g.lo : g.perp);
	const int need_hi = p.y_row0 + p.y_rows - 1 + (g.vertical ? g.lo + g.L - 1 : g.perp);
	float a[LINE_MAXOUT], b[LINE_MAXOUT];
#pragma unroll
	for (int q = 0; q < LINE_MAXOUT; q++) { a[q] = 0.f; b[q] = 0.f; }
	if (NA) {
		line_scan<false>(T, S, g, p.a_src, plane, p.w, p.h, ox0, oy0, tid, negzero, need_lo, need_hi);
		LineOutIter it(g, tid);
#pragma unroll
		for (int q = 0; q < LINE_MAXOUT; q++) {
			if (q * 256 + tid < total_out)
				a[q] = fminf(S[line_idx(g, it.line, it.o)], T[line_idx(g, it.line, it.o + g.L - 1)]);
			it.next();
		}
		__syncthreads();
	}
	if (NB) {
		line_scan<true>(T, S, g, p.b_src, plane, p.w, p.h, ox0, oy0, tid, negzero, need_lo, need_hi);
		LineOutIter it(g, tid);
#pragma unroll
		for (int q = 0; q < LINE_MAXOUT; q++) {
			if (q * 256 + tid < total_out)
				b[q] = fmaxf(S[line_idx(g, it.line, it.o)], T[line_idx(g, it.line, it.o + g.L - 1)]);
			it.next();
		}
	}
	if (__syncthreads_or(negzero) && tid == 0) atomicOr(flag, 1);
	LineOutIter it(g, tid);
#pragma unroll
	for (int q = 0; q < LINE_MAXOUT; q++) {
		const int line = it.line, o = it.o;
		it.next();
		if (q * 256 + tid >= total_out) continue;
		const int i = g.vertical ? ox0 + line : ox0 + o;
		const int j = g.vertical ? oy0 + o : oy0 + line;       // global row
		const int jj = j - p.y_row0;
		if (i >= p.w || jj >= p.y_rows) continue;
		float x = 0.f;
		if (EpiNeeds<EPI>::x) x = band_pixel(p.x_src, plane, p.w, p.h, i, j);
		const long long off = plane * p.y_pstride + (long long)jj * p.w + i;
		if (EPI == EPI_AB) { p.y[off] = a[q]; p.y2[off] = b[q]; }
		else p.y[off] = epilogue<EPI>(a[q], b[q], x);
	}
}
