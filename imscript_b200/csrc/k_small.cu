// k_small.cu -- HBM-bound kernels for structuring elements inside the 3x3
// neighbourhood (cross, square, disk2, hrec2, vrec2, ...): BASELINE configs
// C1 (square erosion) and C5 (cross gradient).
//
// No shared memory: every thread owns 4 adjacent columns (one float4 load and
// one float4 store per row, fully coalesced), gets the neighbouring columns
// from the adjacent lanes with warp shuffles, and marches down its rows with a
// 3-row window in registers.  Two-stage operations (opening, closing, tophat,
// bothat, oscillation) keep the intermediate erosion/dilation rows in
// registers too, so every operation moves 4 B in + 4 B out per sample.
// Reductions use min.f32/max.f32 (FMNMX/FMNMX3), which ignore NaN operands
// exactly like fmin/fmax; they are not order-preserving for +0/-0, so any -0.0
// seen in the loaded data raises *flag and the dispatcher re-runs the
// order-preserving kernels.
#include <cstdlib>
#include <type_traits>
#include <cstring>

#include "dispatch.cuh"

struct SmallArgs {
	Band x;
	float *y;
	long long y_pstride;
	int y_row0, y_rows;
	int x_rows;         // rows the band x holds: [x.row0, x.row0 + x_rows); no other row may be read
	int w, h;
	unsigned mask;      // bit (dy+1)*3+(dx+1)
	int rows_per_warp;
	int wx_log2;        // k_small_1: 2^wx_log2 warps of a CTA sit side by side on the same rows
	int inline_exact;   // k_small_1, canonical cross / square lists: warps that meet a -0.0 switch to the
	                    // order-preserving reduction themselves; no flag, no gated re-run
	int epi;            // Epi
	int stage1_min, stage1_max;   // two-stage: which temporaries exist
	int need_a, need_b;           // final pass: erosion side / dilation side
	int a_from_tmax, b_from_tmin; // two-stage: A = min over tmax, B = max over tmin
	int *flag;
};

template <bool ISMAX> __device__ __forceinline__ float mm(float a, float b)
{
	return ISMAX ? fmaxf(a, b) : fminf(a, b);
}

// reduction over the masked 3x3 neighbourhood of column c (window columns
// c..c+2 of the three rows)
template <int MASK, bool ISMAX, int N>
__device__ __forceinline__ float red3x3(const float (&up)[N], const float (&mid)[N],
		const float (&dn)[N], int c, unsigned rmask)
{
	float r = ISMAX ? -CUDART_INF_F : CUDART_INF_F;
	const unsigned m = MASK >= 0 ? (unsigned)MASK : rmask;
#pragma unroll
	for (int dx = 0; dx < 3; dx++) {
		if (m & (1u << dx))       r = mm<ISMAX>(r, up[c + dx]);
		if (m & (1u << (3 + dx))) r = mm<ISMAX>(r, mid[c + dx]);
		if (m & (1u << (6 + dx))) r = mm<ISMAX>(r, dn[c + dx]);
	}
	return r;
}

// The same reduction in the REFERENCE's element order with its tie rule (glibc
// fmin/fmax return the later operand on ties, so the last occurrence of the
// extremum wins: SURVEY.md 9.1-Z) -- only +0/-0 can tell the difference.  Lists
// in canonical order only: cross = (-1,0),(0,0),(1,0),(0,-1),(0,1) (src/morsi.c:484),
// square / disk2 = dx outer, dy inner (src/morsi.c:485, 319-320).
template <int MASK, bool ISMAX, int N>
__device__ __forceinline__ float red3x3_exact(const float (&up)[N], const float (&mid)[N], const float (&dn)[N], int c)
{
	float r = ISMAX ? -CUDART_INF_F : CUDART_INF_F;
#define STEP(v) do { const float t_ = (v); r = ISMAX ? (t_ >= r ? t_ : r) : (t_ <= r ? t_ : r); } while (0)
	if (MASK == 0272) {
		STEP(mid[c]); STEP(mid[c + 1]); STEP(mid[c + 2]); STEP(up[c + 1]); STEP(dn[c + 1]);
	} else {
#pragma unroll
		for (int dx = 0; dx < 3; dx++) { STEP(up[c + dx]); STEP(mid[c + dx]); STEP(dn[c + dx]); }
	}
#undef STEP
	return r;
}

// fast reduction, or the order-preserving one once the warp has met a -0.0 (canonical lists only)
template <int MASK, bool ISMAX, int N>
__device__ __forceinline__ float red3x3_sel(bool slow, const float (&up)[N], const float (&mid)[N],
		const float (&dn)[N], int c, unsigned rmask)
{
	if (MASK == 0272 || MASK == 0777) {
		if (slow) return red3x3_exact<(MASK == 0272 ? 0272 : 0777), ISMAX>(up, mid, dn, c);
	}
	return red3x3<MASK, ISMAX>(up, mid, dn, c, rmask);
}

__device__ __forceinline__ float apply_epi(int epi, float a, float b, float x)
{
	switch (epi) {
	case EPI_A: return a;
	case EPI_B: return b;
	case EPI_B_SUB_A: return epilogue<EPI_B_SUB_A>(a, b, x);
	case EPI_X_SUB_A: return epilogue<EPI_X_SUB_A>(a, b, x);
	case EPI_B_SUB_X: return epilogue<EPI_B_SUB_X>(a, b, x);
	case EPI_LAP: return epilogue<EPI_LAP>(a, b, x);
	case EPI_ENH: return epilogue<EPI_ENH>(a, b, x);
	case EPI_BLUR: return epilogue<EPI_BLUR>(a, b, x);
	case EPI_A_SUB_B: return epilogue<EPI_A_SUB_B>(a, b, x);
	case EPI_X_SUB_B: return epilogue<EPI_X_SUB_B>(a, b, x);
	case EPI_A_SUB_X: return epilogue<EPI_A_SUB_X>(a, b, x);
	case EPI_IBLUR: return epilogue<EPI_IBLUR>(a, b, x);
	case EPI_EBLUR: return epilogue<EPI_EBLUR>(a, b, x);
	case EPI_CBLUR: return epilogue<EPI_CBLUR>(a, b, x);
	}
	return a;
}

// One input row as fetched from global memory: this thread's 4 columns plus,
// on the warp's edge lanes, the columns just outside the warp's span.  Fetching
// is split from assembling so that the loads of a row PF rows ahead are in
// flight while the current row is reduced (one float4 per thread per row is far
// too little memory-level parallelism to fill HBM otherwise).
template <int HALO>
struct RawRow {
	float c[4];
	float e[HALO];      // lane 0: columns x0-1 (, x0-2); lane 31: columns x0+4 (, x0+5)
	unsigned ok;        // VEC: bit 0: c[] exist, bit 1+k: e[k] exists or is unused (absent samples are NaN on use)
};

// VEC: unconditional loads (absent samples read the plane's first floats
// instead) so that every prefetch slot keeps its own registers in flight; see
// fetch1 below for what a predicated load into a NaN-initialised register costs.
template <int HALO, bool VEC>
__device__ __forceinline__ void fetch_row(const float *plane, int row0, int w, int h,
		int j, int x0, int lane, RawRow<HALO> &r, int row_lo = 0, int row_hi = 0x7fffffff)
{
	const bool rowok = j >= 0 && j < h && j >= row_lo && j <= row_hi;   // in the image AND held by the band
	const float *row = plane + (long long)(j - row0) * w;
	if (VEC) {
		const bool cok = rowok && x0 < w;
		const float4 q = __ldg(reinterpret_cast<const float4 *>(cok ? row + x0 : plane));
		r.c[0] = q.x; r.c[1] = q.y; r.c[2] = q.z; r.c[3] = q.w;
		unsigned ok = cok ? 1u : 0u;
#pragma unroll
		for (int k = 0; k < HALO; k++) {
			const int off = lane == 0 ? -1 - k : 4 + k;
			const bool eok = rowok && ((lane == 0 && x0 - 1 - k >= 0 && x0 - 1 - k < w) || (lane == 31 && x0 + 4 + k < w));
			r.e[k] = __ldg(eok ? row + x0 + off : plane);
			ok |= (eok || (lane != 0 && lane != 31)) ? (2u << k) : 0u;   // inner lanes never use e[]
		}
		r.ok = ok;
		return;
	}
	const float nan = CUDART_NAN_F;
	r.ok = ~0u;
#pragma unroll
	for (int k = 0; k < 4; k++) r.c[k] = nan;
#pragma unroll
	for (int k = 0; k < HALO; k++) r.e[k] = nan;
	if (!rowok) return;
#pragma unroll
	for (int k = 0; k < 4; k++)
		if (x0 + k < w) r.c[k] = __ldg(row + x0 + k);
	if (lane == 0) {
#pragma unroll
		for (int k = 0; k < HALO; k++)
			if (x0 - 1 - k >= 0 && x0 - 1 - k < w) r.e[k] = __ldg(row + x0 - 1 - k);
	}
	if (lane == 31) {
#pragma unroll
		for (int k = 0; k < HALO; k++)
			if (x0 + 4 + k < w) r.e[k] = __ldg(row + x0 + 4 + k);
	}
}

// columns x0-HALO .. x0+3+HALO of a fetched row into v[0 .. 4+2*HALO)
template <int HALO>
__device__ __forceinline__ void assemble_row(const RawRow<HALO> &r0, int lane,
		float (&v)[4 + 2 * HALO], unsigned &negzero)
{
	RawRow<HALO> r = r0;
	const float nan = CUDART_NAN_F;
	// absent samples (rows outside the image, columns beyond it): rare, so a branch, not selects
	constexpr unsigned full = (2u << HALO) - 1u;
	if (r.ok != full && r.ok != ~0u) {
		if (!(r.ok & 1u)) { r.c[0] = nan; r.c[1] = nan; r.c[2] = nan; r.c[3] = nan; }
#pragma unroll
		for (int k = 0; k < HALO; k++) if (!(r.ok & (2u << k))) r.e[k] = nan;
	}
	// -0.0 is INT_MIN as a signed word: a running 3-input integer minimum sees it
	int z = (int)negzero;
	z = min(min(z, __float_as_int(r.c[0])), __float_as_int(r.c[1]));
	z = min(min(z, __float_as_int(r.c[2])), __float_as_int(r.c[3]));
	if (HALO == 2) z = min(min(z, __float_as_int(r.e[0])), __float_as_int(r.e[HALO - 1]));
	else z = min(z, __float_as_int(r.e[0]));
	negzero = (unsigned)z;
#pragma unroll
	for (int k = 0; k < 4; k++) v[HALO + k] = r.c[k];
	float l1 = __shfl_up_sync(0xffffffffu, r.c[3], 1);
	float r1 = __shfl_down_sync(0xffffffffu, r.c[0], 1);
	float l2 = 0.f, r2 = 0.f;
	if (HALO == 2) {
		l2 = __shfl_up_sync(0xffffffffu, r.c[2], 1);
		r2 = __shfl_down_sync(0xffffffffu, r.c[1], 1);
	}
	if (lane == 0) { l1 = r.e[0]; if (HALO == 2) l2 = r.e[HALO - 1]; }
	if (lane == 31) { r1 = r.e[0]; if (HALO == 2) r2 = r.e[HALO - 1]; }
	if (HALO == 1) { v[0] = l1; v[5] = r1; }
	else { v[0] = l2; v[1] = l1; v[6] = r1; v[7] = r2; }
}

// Loop-invariant part of fetch_row's VEC path (k_small_2): column pointers with
// absent columns redirected to the plane's first floats, which samples exist
// column-wise, and the row clamp that keeps every load inside the band.
template <int HALO>
struct FetchCtx {
	const float *pc;          // this thread's 4 columns of row `row0`
	const float *pe[HALO];    // its edge columns (lanes 0 / 31), or a dummy
	unsigned colbits;         // RawRow::ok of a row that exists
	int row0, h, w;
	int row_lo, row_hi;       // rows of the image held by the band
};

template <int HALO>
__device__ __forceinline__ FetchCtx<HALO> fetch_ctx(const float *plane, int row0, int x_rows, int w, int h, int x0, int lane)
{
	FetchCtx<HALO> f;
	f.row0 = row0; f.h = h; f.w = w;
	f.row_lo = max(0, row0); f.row_hi = min(h, row0 + x_rows) - 1;
	const bool cok = x0 < w;
	f.pc = plane + (cok ? x0 : 0);
	f.colbits = cok ? 1u : 0u;
#pragma unroll
	for (int k = 0; k < HALO; k++) {
		const int off = lane == 0 ? -1 - k : 4 + k;
		const bool eok = (lane == 0 && x0 - 1 - k >= 0 && x0 - 1 - k < w) || (lane == 31 && x0 + 4 + k < w);
		f.pe[k] = plane + (eok ? x0 + off : 0);
		f.colbits |= (eok || (lane != 0 && lane != 31)) ? (2u << k) : 0u;   // inner lanes never use e[]
	}
	return f;
}

// row j of the image; rows outside the image -- or outside the band: by the caller's contract
// no element offset reaches those -- load the nearest held row and are marked absent
template <int HALO>
__device__ __forceinline__ void fetch_row_vec(const FetchCtx<HALO> &f, int j, RawRow<HALO> &r)
{
	const int jc = min(max(j, f.row_lo), f.row_hi);
	const long long off = (long long)(jc - f.row0) * f.w;
	const float4 q = __ldg(reinterpret_cast<const float4 *>(f.pc + off));
	r.c[0] = q.x; r.c[1] = q.y; r.c[2] = q.z; r.c[3] = q.w;
#pragma unroll
	for (int k = 0; k < HALO; k++) r.e[k] = __ldg(f.pe[k] + off);
	r.ok = jc == j ? f.colbits : 0u;
}

#define SMALL_PF 3     // rows fetched ahead of the row being reduced

template <bool VEC>
__device__ __forceinline__ void store_row(float *yrow, int x0, int w, const float (&o)[4])
{
	if (VEC) {
		if (x0 < w) *reinterpret_cast<float4 *>(yrow + x0) = make_float4(o[0], o[1], o[2], o[3]);
	} else {
#pragma unroll
		for (int c = 0; c < 4; c++)
			if (x0 + c < w) yrow[x0 + c] = o[c];
	}
}

// ---- single stage -----------------------------------------------------------
// Row fetch for the single-stage kernel.  The loads are UNCONDITIONAL (an
// absent row or column reads a valid dummy address instead) and the NaN of an
// absent sample is selected when the row is consumed: a predicated load that
// merges into a NaN-initialised register makes the compiler funnel every
// prefetch slot through one register and wait for the load right after it was
// issued, which defeats the prefetch (seen in the round-1 ncu source view).
struct Raw1 { float4 q; float e; };

template <bool VEC>
__device__ __forceinline__ void fetch1(const float *row, const float *dummy, bool row_ok, bool col_ok, bool edge_ok,
		int eoff, int w, int x0, Raw1 &r)
{
	if (VEC) {
		const float *q = (row_ok && col_ok) ? row : dummy;
		const float *e = (row_ok && edge_ok) ? row + eoff : dummy;
		r.q = __ldg(reinterpret_cast<const float4 *>(q));
		r.e = __ldg(e);
	} else {
		const float nan = CUDART_NAN_F;
		r.q = make_float4(nan, nan, nan, nan);
		r.e = nan;
		if (row_ok && x0 < w) r.q.x = __ldg(row);
		if (row_ok && x0 + 1 < w) r.q.y = __ldg(row + 1);
		if (row_ok && x0 + 2 < w) r.q.z = __ldg(row + 2);
		if (row_ok && x0 + 3 < w) r.q.w = __ldg(row + 3);
		if (row_ok && edge_ok) r.e = __ldg(row + eoff);
	}
}

// ok / eok: the row's samples / its edge sample exist (VEC only; the scalar
// path stored NaN at fetch time)
template <bool VEC>
__device__ __forceinline__ void assemble1(const Raw1 &r, bool ok, bool eok, int lane, float (&v)[6], unsigned &negzero)
{
	const float nan = CUDART_NAN_F;
	float4 q = r.q;
	float e = r.e;
	if (VEC) {
		if (!ok) q = make_float4(nan, nan, nan, nan);
		if (!eok) e = nan;
	}
	// -0.0 is INT_MIN as a signed word: a running 3-input integer minimum sees it
	int z = (int)negzero;
	z = min(min(z, __float_as_int(q.x)), __float_as_int(q.y));
	z = min(min(z, __float_as_int(q.z)), __float_as_int(q.w));
	z = min(z, __float_as_int(e));
	negzero = (unsigned)z;
	const float l1 = __shfl_up_sync(0xffffffffu, q.w, 1);
	const float r1 = __shfl_down_sync(0xffffffffu, q.x, 1);
	v[0] = lane == 0 ? e : l1;
	v[1] = q.x; v[2] = q.y; v[3] = q.z; v[4] = q.w;
	v[5] = lane == 31 ? e : r1;
}

// EPI >= 0: the epilogue (and with it which of min / max is needed) is a
// compile-time constant; EPI < 0: taken from p.epi at run time.
// PF: rows in flight per thread (a multiple of 3).  The 8 warps of a CTA are
// laid out 2^wx_log2 side by side (adjacent 128-column spans of the same rows:
// longer contiguous runs per DRAM page) by 8 >> wx_log2 row segments.
template <int MASK, bool VEC, int EPI, int PF>
__global__ void __launch_bounds__(256) k_small_1(SmallArgs p)
{
	const int lane = threadIdx.x;
	const int wxm = (1 << p.wx_log2) - 1;
	const int x0 = (((blockIdx.x << p.wx_log2) + (threadIdx.y & wxm)) * 32 + lane) * 4;
	const int plane = blockIdx.z;
	const int seg = blockIdx.y * (blockDim.y >> p.wx_log2) + (threadIdx.y >> p.wx_log2);
	const int jj0 = seg * p.rows_per_warp;
	if (jj0 >= p.y_rows || x0 - 4 * lane >= p.w) return;      // the whole warp: no barrier in this kernel
	const int jj1 = min(p.y_rows, jj0 + p.rows_per_warp);
	const int y0 = p.y_row0 + jj0, y1 = p.y_row0 + jj1;
	const int w = p.w, h = p.h;
	unsigned negzero = 0;
	constexpr bool CT = EPI >= 0;
	const bool need_a = CT ? EpiNeeds<CT ? EPI : 0>::a : (p.need_a != 0);
	const bool need_b = CT ? EpiNeeds<CT ? EPI : 0>::b : (p.need_b != 0);
	const bool inl = (MASK == 0272 || MASK == 0777) && p.inline_exact != 0;

	// running pointers: next row to fetch, next row to store
	const float *dummy = p.x.p + plane * p.x.pstride;          // always readable, 16-byte aligned
	const float *fp = dummy + (long long)(y0 - 1 - p.x.row0) * w + x0;
	int fj = y0 - 1;                                  // global row fp points at
	float *yq = p.y + plane * p.y_pstride + (long long)(y0 - p.y_row0) * w + x0;

	// loop-invariant predicates of the fetch
	const bool col_ok = x0 < w;
	const int eoff = lane == 0 ? -1 : 4;
	const bool edge_ok = (lane == 0 && x0 > 0 && x0 - 1 < w) || (lane == 31 && x0 + 4 < w);

	float in[3][6];
	Raw1 raw[PF];
	// input rows j = y0-1 .. y1; window slot of row j is (j-(y0-1)) % 3, its
	// prefetch slot (j-(y0-1)) % PF; rows are fetched PF ahead, never below
	// row y1 (the source band may end there)
	// rows of the image that the band holds and this warp needs (an element without a
	// row above / below makes the caller's band end at the output rows: SURVEY 8e)
	const int row_lo = max(0, p.x.row0), row_hi = min(min(h, p.x.row0 + p.x_rows) - 1, y1);
#define ROW_OK(j) ((j) >= row_lo && (j) <= row_hi)
#pragma unroll
	for (int k = 0; k < PF; k++) {
		fetch1<VEC>(fp, dummy, ROW_OK(fj), col_ok, edge_ok, eoff, w, x0, raw[k]);
		fp += w; fj++;
	}
	assemble1<VEC>(raw[0], ROW_OK(y0 - 1) && col_ok, ROW_OK(y0 - 1) && edge_ok, lane, in[0], negzero);
	fetch1<VEC>(fp, dummy, ROW_OK(fj), col_ok, edge_ok, eoff, w, x0, raw[0]);
	fp += w; fj++;
	assemble1<VEC>(raw[1], ROW_OK(y0) && col_ok, ROW_OK(y0) && edge_ok, lane, in[1], negzero);
	fetch1<VEC>(fp, dummy, ROW_OK(fj), col_ok, edge_ok, eoff, w, x0, raw[1]);
	fp += w; fj++;
	for (int jb = y0 + 1; jb <= y1; jb += PF) {
#pragma unroll
		for (int u = 0; u < PF; u++) {
			const int j = jb + u;               // row slot (u+2)%3, prefetch slot (u+2)%PF
			if (j <= y1) {
				const bool rok = ROW_OK(j);
				assemble1<VEC>(raw[(u + 2) % PF], rok && col_ok, rok && edge_ok, lane, in[(u + 2) % 3], negzero);
				fetch1<VEC>(fp, dummy, ROW_OK(fj), col_ok, edge_ok, eoff, w, x0, raw[(u + 2) % PF]);
				fp += w; fj++;
				const float (&up)[6] = in[u % 3];
				const float (&mid)[6] = in[(u + 1) % 3];
				const float (&dn)[6] = in[(u + 2) % 3];
				float o[4];
				// a warp that has met a -0.0 reduces in the reference's order from then on
				const bool slow = inl && __any_sync(0xffffffffu, negzero == 0x80000000u);
#pragma unroll
				for (int c = 0; c < 4; c++) {
					float a = 0.f, b = 0.f;
					if (slow) {
						if (need_a) a = red3x3_exact<(MASK == 0272 ? 0272 : 0777), false>(up, mid, dn, c);
						if (need_b) b = red3x3_exact<(MASK == 0272 ? 0272 : 0777), true>(up, mid, dn, c);
					} else {
						if (need_a) a = red3x3<MASK, false>(up, mid, dn, c, p.mask);
						if (need_b) b = red3x3<MASK, true>(up, mid, dn, c, p.mask);
					}
					o[c] = CT ? epilogue<CT ? EPI : 0>(a, b, mid[c + 1]) : apply_epi(p.epi, a, b, mid[c + 1]);
				}
				if (VEC) { if (col_ok) *reinterpret_cast<float4 *>(yq) = make_float4(o[0], o[1], o[2], o[3]); }
				else store_row<VEC>(yq - x0, x0, w, o);
				yq += w;
			}
		}
	}
#undef ROW_OK
	if (!inl && __any_sync(0xffffffffu, negzero == 0x80000000u) && lane == 0) atomicOr(p.flag, 1);
}

// ---- two stages fused ---------------------------------------------------------
// S1: the first stage is a minimum (0: opening, tophat), a maximum (1: closing,
// bothat), or taken from p.stage1_max at run time (-1); EPI as in k_small_1.
template <int MASK, bool VEC, bool OSC, int S1, int EPI>
__global__ void __launch_bounds__(256, OSC ? 2 : 3) k_small_2(SmallArgs p)
{
	const bool s1max = S1 < 0 ? p.stage1_max != 0 : S1 == 1;
	constexpr bool CT = EPI >= 0;
	const bool inl = (MASK == 0272 || MASK == 0777) && p.inline_exact != 0;
	const int lane = threadIdx.x;
	const int x0 = (blockIdx.x * 32 + lane) * 4;
	const int plane = blockIdx.z;
	const int seg = blockIdx.y * blockDim.y + threadIdx.y;
	const int jj0 = seg * p.rows_per_warp;
	if (jj0 >= p.y_rows) return;
	const int jj1 = min(p.y_rows, jj0 + p.rows_per_warp);
	const int y0 = p.y_row0 + jj0, y1 = p.y_row0 + jj1;
	const float *xp = p.x.p + plane * p.x.pstride;
	float *yp = p.y + plane * p.y_pstride;
	unsigned negzero = 0;
	const float nan = CUDART_NAN_F;

	float in[3][8];     // input rows, columns x0-2 .. x0+5
	float t1[3][6];     // first temporary rows, columns x0-1 .. x0+4
	float t2[3][6];     // second temporary (oscillation only)
	// column validity of the temporaries (outside the image they are absent)
	bool cok[6];
#pragma unroll
	for (int c = 0; c < 6; c++) cok[c] = (x0 - 1 + c >= 0) && (x0 - 1 + c < p.w);
	const bool edge_thread = !(cok[0] && cok[5]);

	// input rows j = y0-2 .. y1+1, slot (j-(y0-2)) % 3
	// temporary row t = j-1 becomes available after loading row j; slot (t-(y0-1)) % 3
	// output row t-1 = j-2 after temporary rows j-3, j-2, j-1 exist
	RawRow<2> raw[SMALL_PF];
	const FetchCtx<2> fc = fetch_ctx<2>(xp, p.x.row0, p.x_rows, p.w, p.h, x0, lane);
#define FETCH2(j, slot) do { if (VEC) fetch_row_vec<2>(fc, (j), slot); \
		else fetch_row<2, VEC>(xp, p.x.row0, p.w, p.h, (j), x0, lane, slot, fc.row_lo, fc.row_hi); } while (0)
#pragma unroll
	for (int k = 0; k < SMALL_PF; k++)
		FETCH2(y0 - 2 + k, raw[k]);
	assemble_row<2>(raw[0], lane, in[0], negzero);
	FETCH2(y0 - 2 + SMALL_PF, raw[0]);
	assemble_row<2>(raw[1], lane, in[1], negzero);
	FETCH2(y0 - 1 + SMALL_PF, raw[1]);
	for (int jb = y0; jb <= y1 + 1; jb += 3) {
#pragma unroll
		for (int u = 0; u < 3; u++) {
			const int j = jb + u;
			if (j <= y1 + 1) {
				assemble_row<2>(raw[(u + 2) % 3], lane, in[(u + 2) % 3], negzero);
				if (j + SMALL_PF <= y1 + 1)
					FETCH2(j + SMALL_PF, raw[(u + 2) % 3]);
				const float (&iu)[8] = in[u % 3];
				const float (&im)[8] = in[(u + 1) % 3];
				const float (&id)[8] = in[(u + 2) % 3];
				// temporary row t = j-1 -> slot u; output row j-2 from temporary rows j-3 (slot u+1),
				// j-2 (slot u+2), j-1 (slot u).  SLOW: the reference's element order (a warp that has
				// met a -0.0 reduces both stages that way from then on); one branch per row, two bodies.
				auto row_body = [&](auto slow_c) {
					constexpr bool SLOW = decltype(slow_c)::value;
					const int t = j - 1;
					const bool trow = t >= 0 && t < p.h;
#pragma unroll
					for (int c = 0; c < 6; c++) {
						float v1, v2 = nan;
						if (OSC) {
							v1 = red3x3_sel<MASK, false>(SLOW, iu, im, id, c, p.mask);
							v2 = red3x3_sel<MASK, true>(SLOW, iu, im, id, c, p.mask);
						} else {
							v1 = s1max ? red3x3_sel<MASK, true>(SLOW, iu, im, id, c, p.mask)
							           : red3x3_sel<MASK, false>(SLOW, iu, im, id, c, p.mask);
						}
						t1[u][c] = v1;
						if (OSC) t2[u][c] = v2;
					}
					// temporaries outside the image are absent: only the threads on the
					// image's left / right edge and the rows above / below it have any
					if (!trow || edge_thread) {
#pragma unroll
						for (int c = 0; c < 6; c++)
							if (!(trow && cok[c])) { t1[u][c] = nan; if (OSC) t2[u][c] = nan; }
					}
					if (j >= y0 + 2) {
						const float (&tu)[6] = t1[(u + 1) % 3];
						const float (&tm)[6] = t1[(u + 2) % 3];
						const float (&td)[6] = t1[u % 3];
						float o[4];
#pragma unroll
						for (int c = 0; c < 4; c++) {
							float a = 0.f, b = 0.f;
							if (OSC) {
								// closing - opening: A = min over dilation, B = max over erosion
								a = red3x3_sel<MASK, false>(SLOW, t2[(u + 1) % 3], t2[(u + 2) % 3], t2[u % 3], c, p.mask);
								b = red3x3_sel<MASK, true>(SLOW, tu, tm, td, c, p.mask);
							} else if (s1max) {
								a = red3x3_sel<MASK, false>(SLOW, tu, tm, td, c, p.mask);
							} else {
								b = red3x3_sel<MASK, true>(SLOW, tu, tm, td, c, p.mask);
							}
							// x at row j-2: input slot of row j-2 is u % 3, columns offset by 2
							o[c] = CT ? epilogue<CT ? EPI : 0>(a, b, iu[c + 2]) : apply_epi(p.epi, a, b, iu[c + 2]);
						}
						store_row<VEC>(yp + (long long)(j - 2 - p.y_row0) * p.w, x0, p.w, o);
					}
				};
				if (inl && __any_sync(0xffffffffu, negzero == 0x80000000u)) row_body(std::true_type());
				else row_body(std::false_type());
			}
		}
	}
#undef FETCH2
	if (!inl && __any_sync(0xffffffffu, negzero == 0x80000000u) && lane == 0) atomicOr(p.flag, 1);
}

// ---- host side ------------------------------------------------------------------
static int small_pf()
{
	static const int pf = getenv("MORSI_SMALL_PF") ? atoi(getenv("MORSI_SMALL_PF")) : 3;
	return pf;
}

template <int MASK, bool VEC>
static void launch_small_t(const SmallArgs &a, int stages, bool osc, dim3 grid, dim3 block, cudaStream_t s)
{
	if (stages == 1) {
		if (VEC) {
			const bool pf3 = small_pf() == 3;
			switch (a.epi) {
#define E(X) case X: if (pf3) k_small_1<MASK, VEC, X, 3><<<grid, block, 0, s>>>(a); else k_small_1<MASK, VEC, X, 6><<<grid, block, 0, s>>>(a); return;
			E(EPI_A) E(EPI_B) E(EPI_B_SUB_A) E(EPI_X_SUB_A) E(EPI_B_SUB_X) E(EPI_LAP) E(EPI_ENH) E(EPI_BLUR)
			E(EPI_IBLUR) E(EPI_EBLUR) E(EPI_CBLUR)
#undef E
			}
		}
		k_small_1<MASK, VEC, -1, 3><<<grid, block, 0, s>>>(a);
	}
	else if (osc) {
		if (VEC && MASK >= 0) k_small_2<MASK, VEC, true, -1, EPI_A_SUB_B><<<grid, block, 0, s>>>(a);
		else k_small_2<MASK, VEC, true, -1, -1><<<grid, block, 0, s>>>(a);
	} else {
		if (VEC && MASK >= 0) {
			// opening, closing, tophat, bothat with everything known at compile time
			if (!a.stage1_max && a.epi == EPI_B) { k_small_2<MASK, VEC, false, 0, EPI_B><<<grid, block, 0, s>>>(a); return; }
			if (a.stage1_max && a.epi == EPI_A) { k_small_2<MASK, VEC, false, 1, EPI_A><<<grid, block, 0, s>>>(a); return; }
			if (!a.stage1_max && a.epi == EPI_X_SUB_B) { k_small_2<MASK, VEC, false, 0, EPI_X_SUB_B><<<grid, block, 0, s>>>(a); return; }
			if (a.stage1_max && a.epi == EPI_A_SUB_X) { k_small_2<MASK, VEC, false, 1, EPI_A_SUB_X><<<grid, block, 0, s>>>(a); return; }
		}
		k_small_2<MASK, VEC, false, -1, -1><<<grid, block, 0, s>>>(a);
	}
}

// returns MORSI_OK and *handled = 1 when it launched the job
int morsi_run_small(MorsiCtx *c, const DevElement *de, const MorsiJob &job, int *flag, int *handled)
{
	*handled = 0;
	const OpPlan plan = morsi_op_plan(job.op);
	if (plan.special || de->info.kind != MORSI_EK_SMALL || de->n == 0) return MORSI_OK;
	SmallArgs a;
	a.x = Band{job.x, job.x_row0, job.x_pstride};
	a.y = job.y; a.y_pstride = job.y_pstride; a.y_row0 = job.y_row0; a.y_rows = job.y_rows;
	a.x_rows = job.x_rows;
	a.w = job.w; a.h = job.h; a.mask = de->info.mask3x3; a.epi = plan.epi; a.flag = flag;
	a.stage1_min = plan.t_min; a.stage1_max = plan.t_max;
	a.need_a = plan.a_from != 0; a.need_b = plan.b_from != 0;
	a.a_from_tmax = plan.a_from == 3; a.b_from_tmin = plan.b_from == 2;
	const bool osc = plan.t_min && plan.t_max;
	const bool vec = (job.w % 4 == 0) && (((uintptr_t)job.x) % 16 == 0) && (((uintptr_t)job.y) % 16 == 0)
		&& (job.x_pstride % 4 == 0) && (job.y_pstride % 4 == 0);
	// single stage: the warps of a CTA side by side (as many as the width feeds)
	int wxl = 0;
	if (plan.stages == 1) {
		static const int forced = getenv("MORSI_SMALL_WX") ? atoi(getenv("MORSI_SMALL_WX")) : -1;
		while (wxl < 3 && (128 << (wxl + 1)) <= job.w + 127) wxl++;
		if (forced >= 0 && forced <= 3) wxl = forced;
	}
	a.wx_log2 = wxl;
	const int segs = 8 >> wxl;                         // row segments per CTA
	// Rows per warp.  Single stage: SHORT marches measure best on B200 (C5: 8 rows
	// 76 % of the HBM copy rate, 64 rows 72 %, 270 rows 60 %): the CTAs in flight
	// then sweep the image like a copy does and the re-read halo rows hit in L2;
	// small images get even shorter marches so that every SM has warps to run.
	// Two stages: longer marches, the warm-up is 4 rows of two reductions (16 rows measured best: cross opening of the C5 batch 0.84 ms, 32 rows 0.91, 64 rows 0.95).
	const int gx = (job.w + (128 << wxl) - 1) / (128 << wxl);
	static const int forced_rpw = getenv("MORSI_SMALL_RPW") ? atoi(getenv("MORSI_SMALL_RPW")) : 0;
	int rpw = plan.stages == 1 ? 8 : 16;
	const int min_rpw = plan.stages == 1 ? 2 : 8;
	while (rpw > min_rpw && (long long)gx * ((job.y_rows + rpw * segs - 1) / (rpw * segs)) * job.planes < 4LL * c->sm_count)
		rpw /= 2;
	if (forced_rpw > 0) rpw = forced_rpw;
	// gridDim.y <= 65535: very tall bands march longer per warp
	while ((job.y_rows + rpw * segs - 1) / (rpw * segs) > 65535) rpw *= 2;
	a.rows_per_warp = rpw;
	dim3 block(32, 8);
	dim3 grid(gx, (job.y_rows + rpw * segs - 1) / (rpw * segs), job.planes);
	const unsigned CROSS = 0272u /* .#. ### .#. */, SQUARE = 0777u;
	// a canonical cross / square list: the kernels resolve signed zeros themselves
	static const bool no_inline = getenv("MORSI_SMALL_INLINE") && !strcmp(getenv("MORSI_SMALL_INLINE"), "0");
	a.inline_exact = vec && !no_inline &&
		((a.mask == CROSS && de->canonical3x3 == 1) || (a.mask == SQUARE && de->canonical3x3 == 2));
	if (!a.inline_exact) MORSI_CU(cudaMemsetAsync(flag, 0, sizeof(int), job.stream));
	if (vec) {
		if (a.mask == CROSS) launch_small_t<(int)0272, true>(a, plan.stages, osc, grid, block, job.stream);
		else if (a.mask == SQUARE) launch_small_t<(int)0777, true>(a, plan.stages, osc, grid, block, job.stream);
		else launch_small_t<-1, true>(a, plan.stages, osc, grid, block, job.stream);
	} else {
		launch_small_t<-1, false>(a, plan.stages, osc, grid, block, job.stream);
	}
	morsi_count_launch(1);
	MORSI_CU(cudaGetLastError());
	*handled = a.inline_exact ? 6 : 1;                 // 6: complete, no gated re-run
	return MORSI_OK;
}

// ---- 3x3 median (src/morsi.c:91-120 over cross / square = disk2) ----------------------------
// The same row march as k_small_1: 4 columns per thread, 3-row register window,
// unconditional prefetch.  A window whose samples are all finite (the usual
// case) goes through a sorting network of which only the middle output is
// live (25 comparators for 9 samples, 9 for 5; verified exhaustively with the
// 0-1 principle); windows with absent or non-finite samples (image frame,
// NaN / Inf data: the reference drops them, src/morsi.c:115) sort with +INF
// stand-ins and pick the reference's variable-count positions (:93-100).
// Ties of +0 and -0 depend on the reference's stable sort: -0.0 raises *flag
// and the order-preserving kernel re-runs the job.
__device__ __forceinline__ void cex3(float &a, float &b)
{
	const float lo = fminf(a, b), hi = fmaxf(a, b);
	a = lo; b = hi;
}
template <int N> __device__ __forceinline__ void sort_small(float (&v)[N])
{
	if (N == 9) {
		cex3(v[0], v[3]); cex3(v[1], v[7]); cex3(v[2], v[5]); cex3(v[4], v[8]);
		cex3(v[0], v[7]); cex3(v[2], v[4]); cex3(v[3], v[8]); cex3(v[5], v[6]);
		cex3(v[0], v[2]); cex3(v[1], v[3]); cex3(v[4], v[5]); cex3(v[7], v[8]);
		cex3(v[1], v[4]); cex3(v[3], v[6]); cex3(v[5], v[7]);
		cex3(v[0], v[1]); cex3(v[2], v[4]); cex3(v[3], v[5]); cex3(v[6], v[8]);
		cex3(v[2], v[3]); cex3(v[4], v[5]); cex3(v[6], v[7]);
		cex3(v[1], v[2]); cex3(v[3], v[4]); cex3(v[5], v[6]);
	} else {
		cex3(v[0], v[1]); cex3(v[3], v[4]); cex3(v[2], v[4]); cex3(v[2], v[3]); cex3(v[1], v[4]);
		cex3(v[0], v[3]); cex3(v[0], v[2]); cex3(v[1], v[3]); cex3(v[1], v[2]);
	}
}
template <int N> __device__ __noinline__ float median_small_general(float v0, float v1, float v2, float v3, float v4,
		float v5, float v6, float v7, float v8)
{
	float v[9] = {v0, v1, v2, v3, v4, v5, v6, v7, v8};
	float s[N];
	int cnt = 0;
#pragma unroll
	for (int i = 0; i < N; i++) { const bool f = isfinite(v[i]); cnt += f; s[i] = f ? v[i] : CUDART_INF_F; }
	sort_small<N>(s);
	if (cnt < 1) return CUDART_NAN_F;                                     // src/morsi.c:93
	const int lo = cnt == 2 ? 0 : cnt / 2;                                // :94 (n == 1), :95 (n == 2), :100 (odd), :98 (even)
	float a = s[0], b = s[1];
#pragma unroll
	for (int i = 1; i < N; i++) { a = i == lo ? s[i] : a; b = i == lo + 1 ? s[i] : b; }
	return (cnt & 1) ? a : __fmul_rn(__fadd_rn(a, b), 0.5f);
}

template <int MASK>
__global__ void __launch_bounds__(256) k_median3(SmallArgs p)
{
	constexpr int N = MASK == 0272 ? 5 : 9;
	constexpr int PF = 3;
	const int lane = threadIdx.x;
	const int wxm = (1 << p.wx_log2) - 1;
	const int x0 = (((blockIdx.x << p.wx_log2) + (threadIdx.y & wxm)) * 32 + lane) * 4;
	const int plane = blockIdx.z;
	const int seg = blockIdx.y * (blockDim.y >> p.wx_log2) + (threadIdx.y >> p.wx_log2);
	const int jj0 = seg * p.rows_per_warp;
	if (jj0 >= p.y_rows || x0 - 4 * lane >= p.w) return;      // the whole warp: no barrier in this kernel
	const int jj1 = min(p.y_rows, jj0 + p.rows_per_warp);
	const int y0 = p.y_row0 + jj0, y1 = p.y_row0 + jj1;
	const int w = p.w, h = p.h;
	unsigned negzero = 0;

	const float *dummy = p.x.p + plane * p.x.pstride;
	const float *fp = dummy + (long long)(y0 - 1 - p.x.row0) * w + x0;
	int fj = y0 - 1;
	float *yq = p.y + plane * p.y_pstride + (long long)(y0 - p.y_row0) * w + x0;
	const bool col_ok = x0 < w;
	const int eoff = lane == 0 ? -1 : 4;
	const bool edge_ok = (lane == 0 && x0 > 0 && x0 - 1 < w) || (lane == 31 && x0 + 4 < w);

	float in[3][6];
	bool fin[3];            // every sample of the row's 6 columns is finite
	Raw1 raw[PF];
	const int row_lo = max(0, p.x.row0), row_hi = min(min(h, p.x.row0 + p.x_rows) - 1, y1);
#define ROW_OK(j) ((j) >= row_lo && (j) <= row_hi)
#define ROW_FIN(r) (isfinite(r[0]) && isfinite(r[1]) && isfinite(r[2]) && isfinite(r[3]) && isfinite(r[4]) && isfinite(r[5]))
#pragma unroll
	for (int k = 0; k < PF; k++) {
		fetch1<true>(fp, dummy, ROW_OK(fj), col_ok, edge_ok, eoff, w, x0, raw[k]);
		fp += w; fj++;
	}
	assemble1<true>(raw[0], ROW_OK(y0 - 1) && col_ok, ROW_OK(y0 - 1) && edge_ok, lane, in[0], negzero);
	fin[0] = ROW_FIN(in[0]);
	fetch1<true>(fp, dummy, ROW_OK(fj), col_ok, edge_ok, eoff, w, x0, raw[0]);
	fp += w; fj++;
	assemble1<true>(raw[1], ROW_OK(y0) && col_ok, ROW_OK(y0) && edge_ok, lane, in[1], negzero);
	fin[1] = ROW_FIN(in[1]);
	fetch1<true>(fp, dummy, ROW_OK(fj), col_ok, edge_ok, eoff, w, x0, raw[1]);
	fp += w; fj++;
	for (int jb = y0 + 1; jb <= y1; jb += PF) {
#pragma unroll
		for (int u = 0; u < PF; u++) {
			const int j = jb + u;               // row slot (u+2)%3
			if (j <= y1) {
				const bool rok = ROW_OK(j);
				assemble1<true>(raw[(u + 2) % PF], rok && col_ok, rok && edge_ok, lane, in[(u + 2) % 3], negzero);
				fin[(u + 2) % 3] = ROW_FIN(in[(u + 2) % 3]);
				fetch1<true>(fp, dummy, ROW_OK(fj), col_ok, edge_ok, eoff, w, x0, raw[(u + 2) % PF]);
				fp += w; fj++;
				const float (&up)[6] = in[u % 3];
				const float (&mid)[6] = in[(u + 1) % 3];
				const float (&dn)[6] = in[(u + 2) % 3];
				// (a thread on the image frame sees NaN = not finite in its edge columns / rows)
				const bool clean = fin[0] && fin[1] && fin[2];
				float o[4];
#pragma unroll
				for (int c = 0; c < 4; c++) {
					float v[9];
					if (N == 9) {
						v[0] = up[c]; v[1] = up[c + 1]; v[2] = up[c + 2]; v[3] = mid[c]; v[4] = mid[c + 1];
						v[5] = mid[c + 2]; v[6] = dn[c]; v[7] = dn[c + 1]; v[8] = dn[c + 2];
					} else {
						v[0] = mid[c]; v[1] = mid[c + 1]; v[2] = mid[c + 2]; v[3] = up[c + 1]; v[4] = dn[c + 1];
						v[5] = 0.f; v[6] = 0.f; v[7] = 0.f; v[8] = 0.f;
					}
					if (clean) {
						float s[N];
#pragma unroll
						for (int i = 0; i < N; i++) s[i] = v[i];
						sort_small<N>(s);
						o[c] = s[N / 2];
					} else {
						o[c] = median_small_general<N>(v[0], v[1], v[2], v[3], v[4], v[5], v[6], v[7], v[8]);
					}
				}
				if (col_ok) *reinterpret_cast<float4 *>(yq) = make_float4(o[0], o[1], o[2], o[3]);
				yq += w;
			}
		}
	}
#undef ROW_OK
#undef ROW_FIN
	if (__any_sync(0xffffffffu, negzero == 0x80000000u) && lane == 0) atomicOr(p.flag, 1);
}

// median by the reference's cross / square (= disk2) over 16-byte aligned planes
int morsi_run_median3(MorsiCtx *c, const DevElement *de, const MorsiJob &job, int *flag, int *handled)
{
	*handled = 0;
	if (job.op != MORSI_MEDIAN || de->info.kind != MORSI_EK_SMALL || de->info.has_duplicates) return MORSI_OK;
	static const bool off = getenv("MORSI_MEDIAN3") && !strcmp(getenv("MORSI_MEDIAN3"), "0");
	if (off) return MORSI_OK;
	const unsigned CROSS = 0272u, SQUARE = 0777u;
	const unsigned mask = de->info.mask3x3;
	if (!((mask == CROSS && de->n == 5) || (mask == SQUARE && de->n == 9))) return MORSI_OK;
	const bool vec = (job.w % 4 == 0) && (((uintptr_t)job.x) % 16 == 0) && (((uintptr_t)job.y) % 16 == 0)
		&& (job.x_pstride % 4 == 0) && (job.y_pstride % 4 == 0);
	if (!vec) return MORSI_OK;
	SmallArgs a;
	a.x = Band{job.x, job.x_row0, job.x_pstride};
	a.y = job.y; a.y_pstride = job.y_pstride; a.y_row0 = job.y_row0; a.y_rows = job.y_rows;
	a.x_rows = job.x_rows;
	a.w = job.w; a.h = job.h; a.mask = mask; a.epi = EPI_A; a.flag = flag;
	a.stage1_min = a.stage1_max = a.need_a = a.need_b = a.a_from_tmax = a.b_from_tmin = 0;
	a.inline_exact = 0;
	int wxl = 0;
	while (wxl < 3 && (128 << (wxl + 1)) <= job.w + 127) wxl++;
	a.wx_log2 = wxl;
	const int segs = 8 >> wxl;
	const int gx = (job.w + (128 << wxl) - 1) / (128 << wxl);
	int rpw = 8;
	while (rpw > 2 && (long long)gx * ((job.y_rows + rpw * segs - 1) / (rpw * segs)) * job.planes < 4LL * c->sm_count)
		rpw /= 2;
	a.rows_per_warp = rpw;
	const long long gy = (job.y_rows + rpw * segs - 1) / (rpw * segs);
	if (gy > 65535 || job.planes > 65535) return MORSI_OK;
	dim3 block(32, 8), grid(gx, (unsigned)gy, job.planes);
	if (mask == CROSS) k_median3<0272><<<grid, block, 0, job.stream>>>(a);
	else k_median3<0777><<<grid, block, 0, job.stream>>>(a);
	morsi_count_launch(1);
	MORSI_CU(cudaGetLastError());
	*handled = 1;
	return MORSI_OK;
}
