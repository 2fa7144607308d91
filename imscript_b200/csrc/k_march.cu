// k_march.cu -- erosion / dilation by convex symmetric "row-run" elements
// (disks): BASELINE configs C2 (disk7 opening/closing) and C4 (disk15 tophat).
//
// A disk of reach R is one centred run per row, half-width hw(dy).  The kernel
// is a streaming register march:
//   * a CTA owns a column strip and a row band; the strip's rows (plus RA halo
//     columns each side) stream through a shared-memory ring filled with
//     cp.async (16-byte LDGSTS), several row pairs ahead of the compute;
//     rows / columns outside the image are written as NaN, which min.f32 /
//     max.f32 ignore -- exactly the reference's getpixel_nan border rule
//     (src/morsi.c:30-35);
//   * a thread owns C adjacent columns and walks down the band two rows at a
//     time.  For each new row it builds the nested horizontal running
//     extrema H_k(x) = ext(in[x-k..x+k]) with one 3-input FMNMX3 per k, and
//     folds H_{hw(dy)} into the 2R+2 register accumulators of the output rows
//     y = r-dy that the row touches (one FMNMX3 per accumulator per row PAIR);
//     an output row leaves the registers when its last contributing row has
//     passed.  Cost per sample and stage: RX + (2R+1)/2 FMNMX instead of the
//     n = e[0] fmin/fmax calls of src/morsi.c:65 (12.5 vs 145 for disk7,
//     28.5 vs 697 for disk15), with ~(C+2RA)/C LDS words.
// The accumulator ring is addressed at compile time by unrolling R+1 steps, so
// the element's shape is a template parameter; the shapes below are the disks
// the CLI can name most often, everything else goes to the tiled kernels.
//
// min.f32/max.f32 do not keep the reference's last-wins order for +0/-0:
// the kernel raises *flag when it sees a -0.0 and the dispatcher re-runs the
// order-preserving path.
#include <cstdlib>

#include "dispatch.cuh"

struct MarchArgs {
	Band src;          // image the reduction runs over
	Band xop;          // x operand of the epilogue (may be unused)
	Band other;        // second reduction operand of the epilogue (may be unused)
	float *y;
	long long y_pstride;
	int y_row0, y_rows;
	int src_rows;      // rows held by src (for the loader's bounds)
	int w, h;
	int band_rows;     // output rows per CTA
	int epi;
	int *flag;
};

#include "shapes.cuh"

template <bool ISMAX> __device__ __forceinline__ float ext2(float a, float b)
{
	return ISMAX ? fmaxf(a, b) : fminf(a, b);
}
template <bool ISMAX> __device__ __forceinline__ float ext3(float a, float b, float c)
{
	return ISMAX ? fmaxf(fmaxf(a, b), c) : fminf(fminf(a, b), c);   // one FMNMX3
}

__device__ __forceinline__ void cp_async16(void *smem, const void *gmem)
{
	unsigned s = (unsigned)__cvta_generic_to_shared(smem);
	asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" :: "r"(s), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" :: "n"(N)); }

__device__ __forceinline__ float march_epi(int epi, float a, float b, float x)
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

// Epilogue with extra operands, kept out of line: the unrolled march calls it
// from 2(R+1) places and must stay small enough for the instruction cache.
template <bool ISMAX>
__device__ __noinline__ void march_emit_general(float *q, const float *qx, const float *qo, int epi, int ncol,
		float m0, float m1, float m2, float m3)
{
	float m[4] = {m0, m1, m2, m3};
	float xv[4] = {0.f, 0.f, 0.f, 0.f}, ov[4] = {0.f, 0.f, 0.f, 0.f}, out[4];
	if (qx) {
		if (ncol == 4) { float4 t = __ldg((const float4 *)qx); xv[0] = t.x; xv[1] = t.y; xv[2] = t.z; xv[3] = t.w; }
		else { float2 t = __ldg((const float2 *)qx); xv[0] = t.x; xv[1] = t.y; }
	}
	if (qo) {
		if (ncol == 4) { float4 t = __ldg((const float4 *)qo); ov[0] = t.x; ov[1] = t.y; ov[2] = t.z; ov[3] = t.w; }
		else { float2 t = __ldg((const float2 *)qo); ov[0] = t.x; ov[1] = t.y; }
	}
#pragma unroll
	for (int c = 0; c < 4; c++)
		out[c] = ISMAX ? march_epi(epi, ov[c], m[c], xv[c]) : march_epi(epi, m[c], ov[c], xv[c]);
	if (ncol == 4) *(float4 *)q = make_float4(out[0], out[1], out[2], out[3]);
	else *(float2 *)q = make_float2(out[0], out[1]);
}

template <class S, int C>
struct MarchCfg {
	static constexpr int R = S::R;
	static constexpr int RX = S::hw(R);               // half-width of the centre row
	static constexpr int RA = (RX + 3) / 4 * 4;       // halo columns, multiple of 4 (16-byte cp.async)
	static constexpr int NTM = 128;                   // marching threads per stage
	static constexpr int NV = C + 2 * RA;             // floats a thread reads per row
	static constexpr int NACC = 2 * R + 2;            // accumulator slots per column
	static constexpr int NRING = 16;                  // input ring rows (8 row pairs)
	static constexpr int DEPTH = NRING / 2 - 1;       // row pairs in flight
	// single-stage kernel: strip of NTM*C output columns
	static constexpr int TW1 = NTM * C;
	static constexpr int PITCH1 = TW1 + 2 * RA;
	// fused two-stage kernel: NTM*C temporary columns, 2*RA fewer output columns
	static constexpr int TW2 = NTM * C - 2 * RA;
	static constexpr int PITCH2 = NTM * C + 2 * RA;   // input ring row: TW2 + 4*RA
	static constexpr int PITCHT = NTM * C + 2 * RA;   // temporary ring row (padded: idle lanes read past TW2+2RA)
	static constexpr int NTRING = 4;                  // temporary ring rows (2 pairs)
};

// The per-thread march: C columns, 2R+2 accumulator slots per column.
template <class S, int C, bool ISMAX>
struct Marcher {
	using K = MarchCfg<S, C>;
	static constexpr int R = K::R, RX = K::RX, RA = K::RA, NV = K::NV, NACC = K::NACC;
	// the accumulators live in the kernel (the fused kernel shares one set
	// between its two warp roles, which are never active in the same thread)
	__device__ static __forceinline__ void reset(float (&acc)[C][NACC])
	{
#pragma unroll
		for (int c = 0; c < C; c++)
#pragma unroll
			for (int k = 0; k < NACC; k++) acc[c][k] = ISMAX ? -CUDART_INF_F : CUDART_INF_F;
	}

	// nested horizontal extrema of one ring row; `base` points at column x-RA
	template <bool CHECK0>
	__device__ static __forceinline__ void chain(const float *base, float (&H)[RX + 1][C], bool &negzero)
	{
		float v[NV];
#pragma unroll
		for (int q = 0; q < NV / C; q++) {
			if (C == 4) { float4 t = *(const float4 *)(base + 4 * q); v[4*q] = t.x; v[4*q+1] = t.y; v[4*q+2] = t.z; v[4*q+3] = t.w; }
			else { float2 t = *(const float2 *)(base + 2 * q); v[2*q] = t.x; v[2*q+1] = t.y; }
		}
#pragma unroll
		for (int c = 0; c < C; c++) {
			if (CHECK0) negzero |= (__float_as_uint(v[RA + c]) == 0x80000000u);
			H[0][c] = v[RA + c];
#pragma unroll
			for (int k = 1; k <= RX; k++)
				H[k][c] = ext3<ISMAX>(H[k - 1][c], v[RA + c - k], v[RA + c + k]);
		}
	}

	// One step = march rows i0 = 2g and i1 = 2g+1 (s = g mod (R+1); the caller's
	// loop over s is fully unrolled, so every index below folds to a constant).
	// They touch outputs o = i0 + d, d in [-2R, 1]; the accumulator slot of o
	// is (2s + d) mod NACC.  Outputs i0-2R and i0-2R+1 are complete afterwards
	// and handed to emit(which, m).
	template <bool CHECK0, bool FOLD, class Emit>
	__device__ static __forceinline__ void step(float (&acc)[C][NACC], const int s, const float *rowA, const float *rowB, bool &negzero, Emit &&emit)
	{
		float H0[RX + 1][C], H1[RX + 1][C];
		chain<CHECK0>(rowA, H0, negzero);
		chain<CHECK0>(rowB, H1, negzero);
#pragma unroll
		for (int d = -2 * R; d <= 1; d++) {
			const int slot = ((2 * s + d) % NACC + NACC) % NACC;
			// row i0 sits at dy0 = -d-R relative to output o, row i1 at dy1 = 1-d-R
			const int k0 = S::hw(-d < 0 ? 0 : -d);               // hw(dy0 + R)
			const int k1 = S::hw(1 - d > 2 * R ? 2 * R : 1 - d); // hw(dy1 + R)
#pragma unroll
			for (int c = 0; c < C; c++) {
				if (d == 1) acc[c][slot] = H1[k1][c];                                // first row of output i0+1
				else if (d == 0) acc[c][slot] = ext2<ISMAX>(H0[k0][c], H1[k1][c]);   // first two rows of output i0
				else if (d == -2 * R) acc[c][slot] = ext2<ISMAX>(acc[c][slot], H0[k0][c]);
				else acc[c][slot] = ext3<ISMAX>(acc[c][slot], H0[k0][c], H1[k1][c]);
			}
		}
		// FOLD: folding in the start value turns an all-NaN window into +-INF, as
		// the reference's a = +-INFINITY start does (src/morsi.c:63,77).  The
		// second stage of a fused pair never sees such a window (the element
		// holds its centre and first-stage results inside the image are not NaN).
		const float init = ISMAX ? -CUDART_INF_F : CUDART_INF_F;
		float m0[C], m1[C];
#pragma unroll
		for (int c = 0; c < C; c++) {
			m0[c] = acc[c][((2 * s - 2 * R) % NACC + NACC) % NACC];
			m1[c] = acc[c][((2 * s - 2 * R + 1) % NACC + NACC) % NACC];
			if (FOLD) { m0[c] = ext2<ISMAX>(m0[c], init); m1[c] = ext2<ISMAX>(m1[c], init); }
		}
		emit(0, m0);
		emit(1, m1);
	}
};

// Ring loader: thread `lt` (of the `nload` loading threads) owns the float4
// columns lt and lt+nload of every ring row.  Out-of-image rows / columns are
// written as NaN, the rest arrives by cp.async.
template <int PITCH>
struct RingLoader {
	const float *g[2];      // global address of (march row 0, column slot k); advanced by two rows per pair
	unsigned sm[2];         // byte offset of the column slot inside a ring row, or ~0u when unused
	bool col_ok[2];
	long long pair_stride;  // 2*w floats
	int w;
	int r;                  // global row of the next pair's first row
	int r_lo, r_hi;         // loadable rows: inside the image and inside the source band

	__device__ __forceinline__ void init(const float *src_plane, int src_row0, int src_rows, int w_, int h,
			int r_first, int gcol0, int lt, int nload)
	{
		w = w_; r = r_first; pair_stride = 2LL * w_;
		r_lo = max(0, src_row0); r_hi = min(h, src_row0 + src_rows);
#pragma unroll
		for (int k = 0; k < 2; k++) {
			const int c4 = lt + k * nload;
			const int gc = gcol0 + 4 * c4;
			sm[k] = c4 < PITCH / 4 ? (unsigned)(16 * c4) : ~0u;
			col_ok[k] = gc >= 0 && gc < w_;
			g[k] = src_plane + (long long)(r_first - src_row0) * w_ + gc;
		}
	}
	// ring_pair_base: shared-memory address of the pair's first row
	__device__ __forceinline__ void load_pair(float *ring_pair_base)
	{
		const unsigned base = (unsigned)__cvta_generic_to_shared(ring_pair_base);
		const bool ra = r >= r_lo && r < r_hi, rb = r + 1 >= r_lo && r + 1 < r_hi;
		const float4 nan4 = make_float4(CUDART_NAN_F, CUDART_NAN_F, CUDART_NAN_F, CUDART_NAN_F);
#pragma unroll
		for (int k = 0; k < 2; k++) {
			if (sm[k] != ~0u) {
				if (ra && col_ok[k]) asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" :: "r"(base + sm[k]), "l"(g[k]));
				else *reinterpret_cast<float4 *>((char *)ring_pair_base + sm[k]) = nan4;
				if (rb && col_ok[k]) asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" :: "r"(base + sm[k] + PITCH * 4), "l"(g[k] + w));
				else *reinterpret_cast<float4 *>((char *)ring_pair_base + sm[k] + PITCH * 4) = nan4;
			}
			g[k] += pair_stride;
		}
		r += 2;
	}
};

template <int C>
__device__ __forceinline__ void store_cols(float *q, const float (&m)[C])
{
	if (C == 4) *(float4 *)q = make_float4(m[0], m[1], m[2], m[3]);
	else *(float2 *)q = make_float2(m[0], m[1]);
}

// ---- single reduction pass -----------------------------------------------------
template <class S, int C, bool ISMAX>
__global__ void __launch_bounds__(128) k_march(MarchArgs p)
{
	using K = MarchCfg<S, C>;
	constexpr int R = K::R, RA = K::RA, PITCH = K::PITCH1;
	extern __shared__ __align__(16) float ring[];     // NRING x PITCH

	const int tid = threadIdx.x;
	const int plane = blockIdx.z;
	const int cx0 = blockIdx.x * K::TW1;              // first output column of the strip
	const int o_base = blockIdx.y * p.band_rows;      // first output row of the band (relative)
	const int nout = min(p.band_rows, p.y_rows - o_base);
	const int Y0 = p.y_row0 + o_base;                 // global row of relative output 0
	const int G = (nout + 2 * R + 1) / 2;             // row pairs to march
	const int x = cx0 + C * tid;                      // first column of this thread
	const bool col_ok = x < p.w;                      // w % 4 == 0: all C columns or none

	RingLoader<PITCH> ld;
	ld.init(p.src.p + plane * p.src.pstride, p.src.row0, p.src_rows, p.w, p.h, Y0 - R, cx0 - RA, tid, K::NTM);

	using M = Marcher<S, C, ISMAX>;
	float acc[C][K::NACC];
	M::reset(acc);
	bool negzero = false;

	// output pointers for relative output row 0
	float *yq = p.y + plane * p.y_pstride + (long long)(Y0 - p.y_row0) * p.w + x;
	const float *xq = p.xop.p ? p.xop.p + plane * p.xop.pstride + (long long)(Y0 - p.xop.row0) * p.w + x : nullptr;
	const float *oq = p.other.p ? p.other.p + plane * p.other.pstride + (long long)(Y0 - p.other.row0) * p.w + x : nullptr;
	const bool plain = p.epi == (ISMAX ? EPI_B : EPI_A);
	const int epi = p.epi, w = p.w;

#pragma unroll 1
	for (int g = 0; g < K::DEPTH; g++) {
		if (g < G) ld.load_pair(ring + (2 * g) * PITCH);
		cp_async_commit();
	}
	const float *my = ring + C * tid;
	long long yoff = -2LL * R * w;                    // element offset of output row 2g-2R
#pragma unroll 1
	for (int g0 = 0; g0 < G; g0 += R + 1) {
#pragma unroll
		for (int s = 0; s <= R; s++) {
			const int g = g0 + s;
			if (g < G) {
				cp_async_wait<K::DEPTH - 1>();          // pair g has landed (this thread's part)
				__syncthreads();                         // ... everyone's part; pair g-1 fully consumed
				if (g + K::DEPTH < G) ld.load_pair(ring + ((2 * (g + K::DEPTH)) & (K::NRING - 1)) * PITCH);
				cp_async_commit();
				const float *rowA = my + ((2 * g) & (K::NRING - 1)) * PITCH;
				const int o0 = 2 * g - 2 * R;
				M::template step<true, true>(acc, s, rowA, rowA + PITCH, negzero, [&](int which, const float (&m)[C]) {
					const int o = o0 + which;
					if (o < 0 || o >= nout || !col_ok) return;
					const long long off = which ? yoff + w : yoff;
					if (plain) { store_cols<C>(yq + off, m); return; }
					march_emit_general<ISMAX>(yq + off, xq ? xq + off : nullptr, oq ? oq + off : nullptr,
							epi, C, m[0], m[1], C == 4 ? m[2] : 0.f, C == 4 ? m[3] : 0.f);
				});
				yoff += 2 * w;
			}
		}
	}
	cp_async_wait<0>();
	if (__syncthreads_or(negzero) && tid == 0) atomicOr(p.flag, 1);
}

// ---- two stages fused -------------------------------------------------------------
// Warps 0-3 run the first reduction (S1MAX ? dilation : erosion) over the input
// ring and write its rows into a small shared-memory ring; warps 4-7 run the
// opposite reduction over that ring, one row pair behind, and write the final
// rows (with the epilogue) to global memory.  The temporary image of
// src/morsi.c:143-146 never exists in HBM: 4 B read + 4 B written per sample.
// Temporary samples outside the image are NaN (absent), SURVEY.md 9.1-B.
template <class S, int C, bool S1MAX>
__global__ void __launch_bounds__(256) k_march2(MarchArgs p)
{
	using K = MarchCfg<S, C>;
	constexpr int R = K::R, RA = K::RA, PITCH = K::PITCH2, PITCHT = K::PITCHT;
	extern __shared__ __align__(16) float smem[];
	float *ring = smem;                               // NRING x PITCH   input rows
	float *tring = smem + K::NRING * PITCH;           // NTRING x PITCHT temporary rows

	const int tid = threadIdx.x;
	const int role = tid >> 7;                        // 0: first stage, 1: second stage (warp-uniform)
	const int mt = tid & 127;
	const int plane = blockIdx.z;
	const int cx0 = blockIdx.x * K::TW2;              // first output column of the strip
	const int o_base = blockIdx.y * p.band_rows;
	const int nout = min(p.band_rows, p.y_rows - o_base);
	const int Y0 = p.y_row0 + o_base;                 // global row of output 0
	const int G2 = (nout + 2 * R + 1) / 2;            // second-stage row pairs
	const int G1 = G2 + R;                            // first-stage row pairs: temporary rows Y0-R .. Y0-R+2*G2-1 (+)
	const int GT = G1 + 1;                            // steps: the second stage runs R+1 steps behind
	const int w = p.w, h = p.h;

	RingLoader<PITCH> ld;
	// first-stage march row 0 is global row Y0-2R; ring column 0 is global column cx0-2RA
	ld.init(p.src.p + plane * p.src.pstride, p.src.row0, p.src_rows, w, h, Y0 - 2 * R, cx0 - 2 * RA, tid, 256);

	using M1 = Marcher<S, C, S1MAX>;
	using M2 = Marcher<S, C, !S1MAX>;
	float acc[C][K::NACC];
	if (role == 0) M1::reset(acc); else M2::reset(acc);
	bool negzero = false;

	// first stage: this thread's temporary columns are cx0-RA+C*mt ..
	const int tx = cx0 - RA + C * mt;
	const bool tcol_ok = tx >= 0 && tx < w;
	// second stage: output columns cx0+C*mt ..
	const int x = cx0 + C * mt;
	const bool col_ok = x < w && C * mt < K::TW2;
	float *yq = p.y + plane * p.y_pstride + (long long)(Y0 - p.y_row0) * w + x;
	const float *xq = p.xop.p ? p.xop.p + plane * p.xop.pstride + (long long)(Y0 - p.xop.row0) * w + x : nullptr;
	const bool plain = p.epi == (S1MAX ? EPI_A : EPI_B);
	const int epi = p.epi;

#pragma unroll 1
	for (int g = 0; g < K::DEPTH; g++) {
		if (g < G1) ld.load_pair(ring + (2 * g) * PITCH);
		cp_async_commit();
	}
	const float *my_in = ring + C * mt;
	const float *my_t = tring + C * mt;
	long long yoff = -2LL * R * w;                    // element offset of output row 2*g2-2R (second stage)
#pragma unroll 1
	for (int g0 = 0; g0 < GT; g0 += R + 1) {
#pragma unroll
		for (int s = 0; s <= R; s++) {
			const int g = g0 + s;
			if (g < GT) {
				cp_async_wait<K::DEPTH - 1>();
				__syncthreads();    // input pair g landed; temporary pair g-R-1 written; older slots free
				if (g + K::DEPTH < G1) ld.load_pair(ring + ((2 * (g + K::DEPTH)) & (K::NRING - 1)) * PITCH);
				cp_async_commit();
				if (role == 0) {
					if (g < G1) {
						const float *rowA = my_in + ((2 * g) & (K::NRING - 1)) * PITCH;
						// temporary rows completed by this step: index 2(g-R), 2(g-R)+1 from global row Y0-R
						const int tp = g - R;
						M1::template step<true, true>(acc, s, rowA, rowA + PITCH, negzero, [&](int which, const float (&m)[C]) {
							if (tp < 0) return;
							const int tr = Y0 - R + 2 * tp + which;          // global row of the temporary
							float *q = tring + ((2 * tp + which) & (K::NTRING - 1)) * PITCHT + C * mt;
							if (tr >= 0 && tr < h && tcol_ok) store_cols<C>(q, m);
							else {
								float nanv[C];
#pragma unroll
								for (int c = 0; c < C; c++) nanv[c] = CUDART_NAN_F;
								store_cols<C>(q, nanv);
							}
						});
					}
				} else {
					const int g2 = g - R - 1;                                 // pair written in the previous step
					if (g2 >= 0) {
						const float *rowA = my_t + ((2 * g2) & (K::NTRING - 1)) * PITCHT;
						const int o0 = 2 * g2 - 2 * R;
						M2::template step<false, false>(acc, s, rowA, rowA + PITCHT, negzero, [&](int which, const float (&m)[C]) {
							const int o = o0 + which;
							if (o < 0 || o >= nout || !col_ok) return;
							const long long off = which ? yoff + w : yoff;
							if (plain) { store_cols<C>(yq + off, m); return; }
							march_emit_general<!S1MAX>(yq + off, xq ? xq + off : nullptr, nullptr,
									epi, C, m[0], m[1], C == 4 ? m[2] : 0.f, C == 4 ? m[3] : 0.f);
						});
						yoff += 2 * w;
					}
				}
			}
		}
	}
	cp_async_wait<0>();
	if (__syncthreads_or(negzero) && tid == 0) atomicOr(p.flag, 1);
}

// ---- host side --------------------------------------------------------------------
// Bands: every CTA marches `rows` output rows plus a warm-up of 2*reach rows
// per stage, and CTAs run in waves of `slots`; pick the band count that
// minimises waves x (rows + warm-up), i.e. no half-empty last wave.
static int pick_band_rows(const MorsiCtx *c, int y_rows, long long strips_x_planes, int ctas_per_sm, int reach, int stages)
{
	const long long slots = (long long)c->sm_count * ctas_per_sm;
	const int warm = 2 * reach * stages + 12;          // rows of warm-up + fixed per-CTA cost
	const int min_rows = 8 * reach * stages > 32 ? 8 * reach * stages : 32;
	int best_rows = y_rows;
	double best = 1e300;
	for (int bands = 1; bands <= 4096; bands++) {
		int rows = (y_rows + bands - 1) / bands;
		rows = (rows + 1) & ~1;
		if (rows < min_rows && bands > 1) break;
		const long long ctas = strips_x_planes * ((y_rows + rows - 1) / rows);
		const long long waves = (ctas + slots - 1) / slots;
		const double cost = (double)waves * (rows + warm);
		if (cost < best * 0.999) { best = cost; best_rows = rows; }
	}
	return best_rows;
}

template <class S, int C, bool ISMAX>
static int launch_march(MorsiCtx *c, const MarchArgs &a0, int planes, cudaStream_t st)
{
	using K = MarchCfg<S, C>;
	MarchArgs a = a0;
	const int strips = (a.w + K::TW1 - 1) / K::TW1;
	constexpr size_t smem = (size_t)K::NRING * K::PITCH1 * sizeof(float);
	static_assert(smem <= 48 * 1024, "ring must fit the default 48 KB");
	int occ = 4;
	cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_march<S, C, ISMAX>, K::NTM, smem);
	a.band_rows = pick_band_rows(c, a.y_rows, (long long)strips * planes, occ > 0 ? occ : 1, S::R, 1);
	dim3 grid(strips, (a.y_rows + a.band_rows - 1) / a.band_rows, planes);
	k_march<S, C, ISMAX><<<grid, K::NTM, smem, st>>>(a);
	morsi_count_launch(1);
	MORSI_CU(cudaGetLastError());
	return MORSI_OK;
}

template <class S, int C, bool S1MAX>
static int launch_march2(MorsiCtx *c, const MarchArgs &a0, int planes, cudaStream_t st)
{
	using K = MarchCfg<S, C>;
	MarchArgs a = a0;
	const int strips = (a.w + K::TW2 - 1) / K::TW2;
	constexpr size_t smem = ((size_t)K::NRING * K::PITCH2 + (size_t)K::NTRING * K::PITCHT) * sizeof(float);
	static_assert(smem <= 48 * 1024, "rings must fit the default 48 KB");
	int occ = 2;
	cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_march2<S, C, S1MAX>, 256, smem);
	a.band_rows = pick_band_rows(c, a.y_rows, (long long)strips * planes, occ > 0 ? occ : 1, S::R, 2);
	dim3 grid(strips, (a.y_rows + a.band_rows - 1) / a.band_rows, planes);
	k_march2<S, C, S1MAX><<<grid, 256, smem, st>>>(a);
	morsi_count_launch(1);
	MORSI_CU(cudaGetLastError());
	return MORSI_OK;
}

template <int ID>
static int launch_shape(MorsiCtx *c, const MarchArgs &a, int planes, bool ismax, cudaStream_t st)
{
	constexpr int C = Shape<ID>::R <= 8 ? 4 : 2;
	return ismax ? launch_march<Shape<ID>, C, true>(c, a, planes, st)
	             : launch_march<Shape<ID>, C, false>(c, a, planes, st);
}

template <int ID>
static int launch_shape2(MorsiCtx *c, const MarchArgs &a, int planes, bool s1max, cudaStream_t st)
{
	constexpr int C = Shape<ID>::R <= 8 ? 4 : 2;
	if (ID == 13 && getenv("MORSI_MARCH_C4"))
		return s1max ? launch_march2<Shape<13>, 4, true>(c, a, planes, st)
		             : launch_march2<Shape<13>, 4, false>(c, a, planes, st);
	return s1max ? launch_march2<Shape<ID>, C, true>(c, a, planes, st)
	             : launch_march2<Shape<ID>, C, false>(c, a, planes, st);
}

template <int ID>
static bool shape_matches(const RowRunPlan &rr)
{
	if (rr.reach != Shape<ID>::R) return false;
	for (int i = 0; i <= 2 * rr.reach; i++)
		if (rr.hw[i] != Shape<ID>::hw(i)) return false;
	return true;
}

static int find_shape(const RowRunPlan &rr)
{
	if (!rr.ok) return -1;
#define T(ID) if (shape_matches<ID>(rr)) return ID;
	T(0) T(1) T(2) T(3) T(4) T(5) T(6) T(7) T(8) T(9) T(10) T(11) T(12) T(13)
#undef T
	return -1;
}

static int launch_by_id(int id, MorsiCtx *c, const MarchArgs &a, int planes, bool ismax, cudaStream_t st)
{
	switch (id) {
#define T(ID) case ID: return launch_shape<ID>(c, a, planes, ismax, st);
	T(0) T(1) T(2) T(3) T(4) T(5) T(6) T(7) T(8) T(9) T(10) T(11) T(12) T(13)
#undef T
	}
	return morsi_set_error(MORSI_ERR_INVALID, "no such shape %d", id);
}

static int launch2_by_id(int id, MorsiCtx *c, const MarchArgs &a, int planes, bool s1max, cudaStream_t st)
{
	switch (id) {
#define T(ID) case ID: return launch_shape2<ID>(c, a, planes, s1max, st);
	T(0) T(1) T(2) T(3) T(4) T(5) T(6) T(7) T(8) T(9) T(10) T(11) T(12) T(13)
#undef T
	}
	return morsi_set_error(MORSI_ERR_INVALID, "no such shape %d", id);
}

// One reduction pass over `src` for output rows [row0,row0+rows) into `dst`.
static int march_pass(MorsiCtx *c, int id, bool ismax, int epi, const MorsiJob &job, Band src, int src_rows,
		Band xop, Band other, float *dst, long long dst_pstride, int row0, int rows, int *flag)
{
	MarchArgs a;
	a.src = src; a.src_rows = src_rows; a.xop = xop; a.other = other;
	a.y = dst; a.y_pstride = dst_pstride; a.y_row0 = row0; a.y_rows = rows;
	a.w = job.w; a.h = job.h; a.epi = epi; a.flag = flag; a.band_rows = rows;
	return launch_by_id(id, c, a, job.planes, ismax, job.stream);
}

int morsi_run_march(MorsiCtx *c, const DevElement *de, const MorsiJob &job, int *flag, int *handled)
{
	*handled = 0;
	const OpPlan plan = morsi_op_plan(job.op);
	if (plan.special) return MORSI_OK;
	const int id = find_shape(de->rowrun);
	if (id < 0) return MORSI_OK;
	const bool aligned = (job.w % 4 == 0) && (((uintptr_t)job.x) % 16 == 0) && (((uintptr_t)job.y) % 16 == 0)
		&& (job.x_pstride % 4 == 0) && (job.y_pstride % 4 == 0);
	if (!aligned) return MORSI_OK;
	const int R = de->rowrun.reach;
	const Band none{nullptr, 0, 0};
	const Band xb{job.x, job.x_row0, job.x_pstride};
	int rc;

	// how many temporaries does the plan need?
	//   1 stage, one side      : 0      (erosion, dilation, i/egradient, i/eblur)
	//   1 stage, both sides    : 1      (gradient, laplacian, enhance, blur, cblur: a = erosion)
	//   2 stages               : 1      (opening, closing, tophat, bothat)
	//   oscillation            : 3
	const bool both1 = plan.stages == 1 && plan.a_from && plan.b_from;
	const bool osc = plan.t_min && plan.t_max;
	if (plan.stages == 1 && !both1) {
		rc = march_pass(c, id, plan.b_from != 0, plan.epi, job, xb, job.x_rows, xb, none,
				job.y, job.y_pstride, job.y_row0, job.y_rows, flag);
		if (rc) return rc;
		*handled = 1;
		return MORSI_OK;
	}
	if (plan.stages == 2 && !osc) {
		// opening, closing, tophat, bothat: both stages in one kernel
		MarchArgs a;
		a.src = xb; a.src_rows = job.x_rows; a.xop = xb; a.other = none;
		a.y = job.y; a.y_pstride = job.y_pstride; a.y_row0 = job.y_row0; a.y_rows = job.y_rows;
		a.w = job.w; a.h = job.h; a.epi = plan.epi; a.flag = flag; a.band_rows = job.y_rows;
		rc = launch2_by_id(id, c, a, job.planes, plan.t_max != 0, job.stream);
		if (rc) return rc;
		*handled = 1;
		return MORSI_OK;
	}
	// temporaries cover the output band grown by one reach (clipped); chunk the
	// band so that a temporary stays below 512 MiB
	const long long budget = 512LL << 20;
	long long rows_fit = budget / ((long long)job.w * 4 * job.planes) - 2 * R;
	if (rows_fit < 8 * R + 64) rows_fit = 8 * R + 64;
	const int chunk = (int)(rows_fit < job.y_rows ? rows_fit : job.y_rows);
	for (int r0 = 0; r0 < job.y_rows; r0 += chunk) {
		const int o0 = job.y_row0 + r0;
		const int orows = job.y_rows - r0 < chunk ? job.y_rows - r0 : chunk;
		float *ydst = job.y + (long long)r0 * job.w;
		if (both1) {
			void *p0; if ((rc = morsi_ws_get(c, job.lane, 0, (size_t)job.w * orows * job.planes * 4, &p0))) return rc;
			const long long tps = (long long)job.w * orows;
			rc = march_pass(c, id, false, EPI_A, job, xb, job.x_rows, none, none, (float *)p0, tps, o0, orows, flag);
			if (rc) return rc;
			rc = march_pass(c, id, true, plan.epi, job, xb, job.x_rows, xb, Band{(float *)p0, o0, tps},
					ydst, job.y_pstride, o0, orows, flag);
			if (rc) return rc;
			continue;
		}
		int t0 = o0 - R; if (t0 < 0) t0 = 0;
		int t1 = o0 + orows + R; if (t1 > job.h) t1 = job.h;
		const int trows = t1 - t0;
		const long long tps = (long long)job.w * trows;
		const size_t tbytes = (size_t)tps * job.planes * 4;
		void *p0, *p1, *p2;
		if ((rc = morsi_ws_get(c, job.lane, 0, tbytes, &p0))) return rc;
		if (!osc) {
			const bool s1max = plan.t_max != 0;
			rc = march_pass(c, id, s1max, s1max ? EPI_B : EPI_A, job, xb, job.x_rows, none, none,
					(float *)p0, tps, t0, trows, flag);
			if (rc) return rc;
			rc = march_pass(c, id, !s1max, plan.epi, job, Band{(float *)p0, t0, tps}, trows, xb, none,
					ydst, job.y_pstride, o0, orows, flag);
			if (rc) return rc;
		} else {
			if ((rc = morsi_ws_get(c, job.lane, 1, tbytes, &p1))) return rc;
			if ((rc = morsi_ws_get(c, job.lane, 2, (size_t)job.w * orows * job.planes * 4, &p2))) return rc;
			const long long ops = (long long)job.w * orows;
			rc = march_pass(c, id, false, EPI_A, job, xb, job.x_rows, none, none, (float *)p0, tps, t0, trows, flag);
			if (rc) return rc;
			rc = march_pass(c, id, true, EPI_B, job, xb, job.x_rows, none, none, (float *)p1, tps, t0, trows, flag);
			if (rc) return rc;
			// opening = max over erosion -> p2 ; closing = min over dilation, minus opening
			rc = march_pass(c, id, true, EPI_B, job, Band{(float *)p0, t0, tps}, trows, none, none,
					(float *)p2, ops, o0, orows, flag);
			if (rc) return rc;
			rc = march_pass(c, id, false, EPI_A_SUB_B, job, Band{(float *)p1, t0, tps}, trows, none,
					Band{(float *)p2, o0, ops}, ydst, job.y_pstride, o0, orows, flag);
			if (rc) return rc;
		}
	}
	*handled = 1;
	return MORSI_OK;
}
