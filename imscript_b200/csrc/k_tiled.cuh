// k_tiled.cuh -- erosion / dilation family for ARBITRARY offset lists (dysk,
// drec/Drec, long hrec/vrec, disks outside shapes.cuh, user masks): the fast
// counterpart of k_exact_minmax for elements no specialised kernel takes.
//
// A CTA stages its 128x16 output tile plus the element's bounding-box halo in
// shared memory (samples outside the image are NaN = absent, src/morsi.c:30-35)
// and evaluates the list straight from the tile: per element one broadcast
// load of its tile offset, per output one shared-memory load and one FMNMX.
// A thread owns columns tx, tx+32, tx+64, tx+96 of rows ty and ty+8, so the
// lanes of a warp read consecutive banks.  min.f32/max.f32 ignore NaN like
// fmin/fmax but not the element order for +0/-0: a -0.0 in a tile raises
// *flag and the dispatcher re-runs k_exact_minmax.
// n compares per sample like the reference (src/morsi.c:56-82), at ~2.2
// instructions per compare instead of a bounds-checked global gather.
#pragma once
#include "k_exact.cuh"

struct TiledGeom {
	int xmin, xmax, ymin, ymax;   // box of the effective offsets
	int pw, ph;                   // tile pitch / rows: 128 + xmax - xmin, 16 + ymax - ymin
	const int *tile_offs;         // per element: (dy - ymin) * pw + (dx - xmin), on the device
	int two_tiles;                // the two reduction sides read different images
};

#define TILED_TX 128
#define TILED_TY 16

// gy_lo .. gy_hi: the rows the outputs of this LAUNCH can need.  A band holds
// exactly those (its caller's contract); the last tile of a launch reaches
// further down, and those rows -- used only by outputs that are not stored --
// must not be read.
__device__ __forceinline__ void tiled_load(float *tile, const Band &src, int plane, int w, int h,
		int gx0, int gy0, int pw, int ph, int tid, bool &negzero, int gy_lo, int gy_hi)
{
	const float *sp = src.p + plane * src.pstride;
	gy_lo = max(gy_lo, 0); gy_hi = min(gy_hi, h - 1);
	for (int t = tid; t < pw * ph; t += 256) {
		const int r = t / pw, cc = t - r * pw;
		const int gx = gx0 + cc, gy = gy0 + r;
		float v = CUDART_NAN_F;
		if (gx >= 0 && gx < w && gy >= gy_lo && gy <= gy_hi) {
			v = __ldg(sp + (long long)(gy - src.row0) * w + gx);
			negzero |= __float_as_uint(v) == 0x80000000u;
		}
		tile[t] = v;
	}
}

// EXACT: the order-preserving form -- `v <= a ? v : a` in element order, exactly glibc's
// fmin / fmax on ties -- at two instructions per compare instead of one.
// It is the re-run the dispatcher gates on the "saw a -0.0" word (p.gate) of the fast
// families: the same results as k_exact_minmax, from shared-memory tiles instead of
// bounds-checked global gathers.  Tile rows are walked with the stride of gridDim.y, so a
// gated launch can be a small grid that costs a few microseconds when it is a no-op.
template <int EPI, bool EXACT>
__global__ void __launch_bounds__(256) k_tiled_minmax(ExactArgs p, TiledGeom g, int *flag)
{
	if (p.gate && *p.gate == 0) return;
	extern __shared__ float tiled_smem[];
	constexpr bool NA = EpiNeeds<EPI>::a, NB = EpiNeeds<EPI>::b;
	const int tx = threadIdx.x, ty = threadIdx.y, tid = ty * 32 + tx;
	const int plane = blockIdx.z;
	const int pw = g.pw, ph = g.ph;
	const int bx = blockIdx.x * TILED_TX;
	for (int by = blockIdx.y * TILED_TY; by < p.y_rows; by += gridDim.y * TILED_TY) {   // tile origin, rows relative to y_row0
	__syncthreads();                                                      // the previous tile has been consumed
	float *tile_a = tiled_smem;                                  // erosion-side source (or the shared one)
	float *tile_b = g.two_tiles ? tiled_smem + pw * ph : tiled_smem;
	int *offs = reinterpret_cast<int *>(tiled_smem + (g.two_tiles ? 2 : 1) * pw * ph);

	bool negzero = false;
	const int gx0 = bx + g.xmin, gy0 = p.y_row0 + by + g.ymin;
	const int need_lo = p.y_row0 + g.ymin, need_hi = p.y_row0 + p.y_rows - 1 + g.ymax;
	if (NA || !g.two_tiles) tiled_load(tile_a, NA ? p.a_src : p.b_src, plane, p.w, p.h, gx0, gy0, pw, ph, tid, negzero, need_lo, need_hi);
	if (NB && g.two_tiles) tiled_load(tile_b, p.b_src, plane, p.w, p.h, gx0, gy0, pw, ph, tid, negzero, need_lo, need_hi);
	for (int k = tid; k < p.n; k += 256) offs[k] = g.tile_offs[k];
	if (__syncthreads_or(negzero) && tid == 0 && !EXACT) atomicOr(flag, 1);

	float a[2][4], b[2][4];
#pragma unroll
	for (int r = 0; r < 2; r++)
#pragma unroll
		for (int c = 0; c < 4; c++) { a[r][c] = CUDART_INF_F; b[r][c] = -CUDART_INF_F; }
	const float *qa = tile_a + ty * pw + tx, *qb = tile_b + ty * pw + tx;
	const int row8 = 8 * pw;
#pragma unroll 2
	for (int k = 0; k < p.n; k++) {
		const int o = offs[k];
#pragma unroll
		for (int r = 0; r < 2; r++)
#pragma unroll
			for (int c = 0; c < 4; c++) {
				if (NA) {
					const float v = qa[o + r * row8 + 32 * c];
					a[r][c] = EXACT ? ((v <= a[r][c]) ? v : a[r][c]) : fminf(a[r][c], v);
				}
				if (NB) {
					const float v = (NA && !g.two_tiles) ? qa[o + r * row8 + 32 * c] : qb[o + r * row8 + 32 * c];
					b[r][c] = EXACT ? ((v >= b[r][c]) ? v : b[r][c]) : fmaxf(b[r][c], v);
				}
			}
	}
#pragma unroll
	for (int r = 0; r < 2; r++) {
		const int jj = by + ty + 8 * r;
		if (jj >= p.y_rows) continue;
		const int j = p.y_row0 + jj;
#pragma unroll
		for (int c = 0; c < 4; c++) {
			const int i = bx + tx + 32 * c;
			if (i >= p.w) continue;
			float x = 0.f;
			if (EpiNeeds<EPI>::x) x = band_pixel(p.x_src, plane, p.w, p.h, i, j);
			const long long o = plane * p.y_pstride + (long long)jj * p.w + i;
			if (EPI == EPI_AB) { p.y[o] = a[r][c]; p.y2[o] = b[r][c]; }
			else p.y[o] = epilogue<EPI>(a[r][c], b[r][c], x);
		}
	}
	}   // tile rows
}

// ---- rank (src/morsi.c:122-139) from the same tiles ------------------------------------
// u = the pixel itself; the count of finite neighbours below u does not depend
// on the element order and treats +0/-0 like the reference's `<`: exact as is.
__global__ void __launch_bounds__(256) k_tiled_rank(MedianArgs p, TiledGeom g)
{
	extern __shared__ float tiled_smem[];
	const int tx = threadIdx.x, ty = threadIdx.y, tid = ty * 32 + tx;
	const int plane = blockIdx.z;
	const int bx = blockIdx.x * TILED_TX, by = blockIdx.y * TILED_TY;
	const int pw = g.pw, ph = g.ph;
	float *tile = tiled_smem;
	int *offs = reinterpret_cast<int *>(tiled_smem + pw * ph);
	bool unused = false;
	tiled_load(tile, p.x_src, plane, p.w, p.h, bx + g.xmin, p.y_row0 + by + g.ymin, pw, ph, tid, unused,
		p.y_row0 + g.ymin, p.y_row0 + p.y_rows - 1 + g.ymax);
	for (int k = tid; k < p.n; k += 256) offs[k] = g.tile_offs[k];
	__syncthreads();

	float u[2][4];
	int cnt[2][4];
#pragma unroll
	for (int r = 0; r < 2; r++)
#pragma unroll
		for (int c = 0; c < 4; c++) {
			// outside the image u is NaN and every comparison fails; those outputs are not stored anyway
			u[r][c] = by + ty + 8 * r < p.y_rows ?
				band_pixel(p.x_src, plane, p.w, p.h, bx + tx + 32 * c, p.y_row0 + by + ty + 8 * r) : CUDART_NAN_F;
			cnt[r][c] = 0;
		}
	const float *q = tile + ty * pw + tx;
	const int row8 = 8 * pw;
	for (int k = 0; k < p.n; k++) {
		const int o = offs[k];
#pragma unroll
		for (int r = 0; r < 2; r++)
#pragma unroll
			for (int c = 0; c < 4; c++) {
				const float v = q[o + r * row8 + 32 * c];
				cnt[r][c] += (isfinite(v) && v < u[r][c]) ? 1 : 0;
			}
	}
#pragma unroll
	for (int r = 0; r < 2; r++) {
		const int jj = by + ty + 8 * r;
		if (jj >= p.y_rows) continue;
#pragma unroll
		for (int c = 0; c < 4; c++) {
			const int i = bx + tx + 32 * c;
			if (i < p.w) p.y[plane * p.y_pstride + (long long)jj * p.w + i] = (float)cnt[r][c];
		}
	}
}
