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

// ---- shapes -------------------------------------------------------------------
template <int ID> struct Shape;
#define MORSI_SHAPE(ID, RY, ...) \
	template <> struct Shape<ID> { \
		static constexpr int R = RY; \
		__host__ __device__ static constexpr int hw(int i) { constexpr int t[2 * RY + 1] = {__VA_ARGS__}; return t[i]; } \
	};
MORSI_SHAPE(0, 2, 1, 2, 2, 2, 1)                                    // disk2.5
MORSI_SHAPE(1, 2, 2, 2, 2, 2, 2)                                    // disk3 (5x5)
MORSI_SHAPE(2, 3, 1, 2, 3, 3, 3, 2, 1)                              // disk3.5
MORSI_SHAPE(3, 3, 2, 3, 3, 3, 3, 3, 2)                              // disk4
MORSI_SHAPE(4, 4, 1, 2, 3, 4, 4, 4, 3, 2, 1)                        // disk4.2
MORSI_SHAPE(5, 4, 2, 3, 4, 4, 4, 4, 4, 3, 2)                        // disk4.5, disk5
MORSI_SHAPE(6, 5, 1, 3, 4, 4, 5, 5, 5, 4, 4, 3, 1)                  // disk5.1
MORSI_SHAPE(7, 5, 3, 4, 5, 5, 5, 5, 5, 5, 5, 4, 3)                  // disk6
MORSI_SHAPE(8, 6, 3, 4, 5, 6, 6, 6, 6, 6, 6, 6, 5, 4, 3)            // disk7
MORSI_SHAPE(9, 7, 3, 5, 6, 6, 7, 7, 7, 7, 7, 7, 7, 6, 6, 5, 3)      // disk8
MORSI_SHAPE(10, 8, 4, 5, 6, 7, 8, 8, 8, 8, 8, 8, 8, 8, 8, 7, 6, 5, 4)   // disk9
MORSI_SHAPE(11, 9, 4, 5, 7, 7, 8, 9, 9, 9, 9, 9, 9, 9, 9, 9, 8, 7, 7, 5, 4)   // disk10
MORSI_SHAPE(12, 11, 4, 6, 7, 8, 9, 10, 10, 11, 11, 11, 11, 11, 11, 11, 11, 11, 10, 10, 9, 8, 7, 6, 4)   // disk12
MORSI_SHAPE(13, 14, 5, 7, 8, 10, 11, 11, 12, 13, 13, 14, 14, 14, 14, 14, 14, 14, 14, 14, 14, 14, 13, 13, 12, 11, 11, 10, 8, 7, 5)   // disk15
#define MORSI_NSHAPES 14

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
	static constexpr int NT = 128;                    // threads per CTA
	static constexpr int TW = NT * C;                 // strip width
	static constexpr int PITCH = TW + 2 * RA;         // floats per ring row
	static constexpr int NV = C + 2 * RA;             // floats a thread reads per row
	static constexpr int NACC = 2 * R + 2;            // accumulator slots per column
	static constexpr int DEPTH = 6;                   // row pairs in flight
	static constexpr int NRING = 2 * (DEPTH + 1);     // ring rows
};

template <class S, int C, bool ISMAX>
__global__ void __launch_bounds__(128) k_march(MarchArgs p)
{
	using K = MarchCfg<S, C>;
	constexpr int R = K::R, RX = K::RX, RA = K::RA, PITCH = K::PITCH, NV = K::NV, NACC = K::NACC;
	extern __shared__ __align__(16) float ring[];     // NRING x PITCH

	const int tid = threadIdx.x;
	const int plane = blockIdx.z;
	const int cx0 = blockIdx.x * K::TW;               // first output column of the strip
	const int o_base = blockIdx.y * p.band_rows;      // first output row of the band (relative)
	const int nout = min(p.band_rows, p.y_rows - o_base);
	const int Y0 = p.y_row0 + o_base;                 // global row of relative output 0
	const int r_first = Y0 - R;                       // global row of march row i = 0
	const int NI = nout + 2 * R;                      // rows that matter
	const int G = (NI + 1) / 2;                       // row pairs
	const float *src = p.src.p + plane * p.src.pstride;
	const float init = ISMAX ? -CUDART_INF_F : CUDART_INF_F;

	// ---- loader: row pair g -> ring rows (2g)%NRING, (2g+1)%NRING --------------
	auto load_pair = [&](int g) {
		constexpr int Q = PITCH / 4;                  // float4 per ring row
		for (int q = tid; q < 2 * Q; q += K::NT) {
			const int half = q >= Q;
			const int c4 = q - half * Q;
			const int i = 2 * g + half;
			const int r = r_first + i;                // global row
			const int gc = cx0 - RA + 4 * c4;         // global column of the float4
			float *dst = ring + ((2 * g + half) % K::NRING) * PITCH + 4 * c4;
			const bool ok = r >= 0 && r < p.h && r >= p.src.row0 && r < p.src.row0 + p.src_rows
				&& gc >= 0 && gc < p.w;
			if (ok) cp_async16(dst, src + (long long)(r - p.src.row0) * p.w + gc);
			else *reinterpret_cast<float4 *>(dst) =
				make_float4(CUDART_NAN_F, CUDART_NAN_F, CUDART_NAN_F, CUDART_NAN_F);
		}
	};

	float acc[C][NACC];
#pragma unroll
	for (int c = 0; c < C; c++)
#pragma unroll
		for (int k = 0; k < NACC; k++) acc[c][k] = init;
	unsigned negzero = 0;

	const int x = cx0 + C * tid;                      // first column of this thread
	const bool col_ok = x < p.w;                      // w % 4 == 0: all C columns or none
	const float *xop = p.xop.p ? p.xop.p + plane * p.xop.pstride : nullptr;
	const float *oth = p.other.p ? p.other.p + plane * p.other.pstride : nullptr;
	float *yp = p.y + plane * p.y_pstride;

	auto emit = [&](int o, const float (&m)[C]) {
		if (o < 0 || o >= nout || !col_ok) return;
		const int gy = Y0 + o;
		float *q = yp + (long long)(gy - p.y_row0) * p.w + x;
		if (p.epi == (ISMAX ? EPI_B : EPI_A)) {          // plain erosion / dilation pass
			if (C == 4) *(float4 *)q = make_float4(m[0], m[1], m[2], m[3]);
			else *(float2 *)q = make_float2(m[0], m[1]);
			return;
		}
		const float *qx = xop ? xop + (long long)(gy - p.xop.row0) * p.w + x : nullptr;
		const float *qo = oth ? oth + (long long)(gy - p.other.row0) * p.w + x : nullptr;
		march_emit_general<ISMAX>(q, qx, qo, p.epi, C, m[0], m[1], C == 4 ? m[2] : 0.f, C == 4 ? m[3] : 0.f);
	};

	// nested horizontal extrema of one ring row for this thread's C columns
	auto chain = [&](const float *row, float (&H)[RX + 1][C]) {
		float v[NV];
		const float *base = row + C * tid;
#pragma unroll
		for (int q = 0; q < NV / C; q++) {
			if (C == 4) { float4 t = *(const float4 *)(base + 4 * q); v[4*q] = t.x; v[4*q+1] = t.y; v[4*q+2] = t.z; v[4*q+3] = t.w; }
			else { float2 t = *(const float2 *)(base + 2 * q); v[2*q] = t.x; v[2*q+1] = t.y; }
		}
#pragma unroll
		for (int c = 0; c < C; c++) {
			negzero |= (__float_as_uint(v[RA + c]) == 0x80000000u);
			H[0][c] = v[RA + c];
#pragma unroll
			for (int k = 1; k <= RX; k++)
				H[k][c] = ext3<ISMAX>(H[k - 1][c], v[RA + c - k], v[RA + c + k]);
		}
	};

	// ---- prologue: DEPTH pairs in flight ---------------------------------------
#pragma unroll 1
	for (int g = 0; g < K::DEPTH; g++) {
		if (g < G) load_pair(g);
		cp_async_commit();
	}

#pragma unroll 1
	for (int g0 = 0; g0 < G; g0 += R + 1) {
#pragma unroll
		for (int s = 0; s <= R; s++) {
			const int g = g0 + s;
			if (g < G) {
				cp_async_wait<K::DEPTH - 1>();          // pair g has landed (this thread's part)
				__syncthreads();                         // ... everyone's part; pair g-1 fully consumed
				if (g + K::DEPTH < G) load_pair(g + K::DEPTH);
				cp_async_commit();

				float H0[RX + 1][C], H1[RX + 1][C];
				chain(ring + ((2 * g) % K::NRING) * PITCH, H0);
				chain(ring + ((2 * g + 1) % K::NRING) * PITCH, H1);
				// rows i0 = 2g, i1 = 2g+1 touch outputs o = i0 + d, d in [-2R, 1];
				// accumulator slot of o is (2s + d) mod NACC
#pragma unroll
				for (int d = -2 * R; d <= 1; d++) {
					const int slot = ((2 * s + d) % NACC + NACC) % NACC;
					// row i0 sits at dy0 = -d-R below output o, row i1 at dy1 = 1-d-R
					const int k0 = S::hw(-d < 0 ? 0 : -d);               // hw(dy0 + R)
					const int k1 = S::hw(1 - d > 2 * R ? 2 * R : 1 - d); // hw(dy1 + R)
#pragma unroll
					for (int c = 0; c < C; c++) {
						if (d == 1) acc[c][slot] = H1[k1][c];                       // first row of output i0+1
						else if (d == 0) acc[c][slot] = ext2<ISMAX>(H0[k0][c], H1[k1][c]);   // first two rows of output i0
						else if (d == -2 * R) acc[c][slot] = ext2<ISMAX>(acc[c][slot], H0[k0][c]);
						else acc[c][slot] = ext3<ISMAX>(acc[c][slot], H0[k0][c], H1[k1][c]);
					}
				}
				// outputs o = i0-2R and i0-2R+1 are complete
				{
					// folding in the start value turns an all-NaN window into +-INF,
					// as the reference's a = +-INFINITY start does (src/morsi.c:63,77)
					float m0[C], m1[C];
#pragma unroll
					for (int c = 0; c < C; c++) {
						m0[c] = ext2<ISMAX>(acc[c][((2 * s - 2 * R) % NACC + NACC) % NACC], init);
						m1[c] = ext2<ISMAX>(acc[c][((2 * s - 2 * R + 1) % NACC + NACC) % NACC], init);
					}
					emit(2 * g - 2 * R, m0);
					emit(2 * g - 2 * R + 1, m1);
				}
			}
		}
	}
	cp_async_wait<0>();
	if (__syncthreads_or(negzero != 0) && tid == 0) atomicOr(p.flag, 1);
}

// ---- host side --------------------------------------------------------------------
template <class S, int C, bool ISMAX>
static int launch_march(MorsiCtx *c, const MarchArgs &a0, int planes, cudaStream_t st)
{
	using K = MarchCfg<S, C>;
	MarchArgs a = a0;
	const int strips = (a.w + K::TW - 1) / K::TW;
	// enough CTAs to fill the machine ~6x over, bands no shorter than 8 reaches
	long long target = 6LL * c->sm_count;
	long long per_band_row = (long long)strips * planes;
	int bands = (int)((target + per_band_row - 1) / per_band_row);
	if (bands < 1) bands = 1;
	int rows = (a.y_rows + bands - 1) / bands;
	const int min_rows = 16 * S::R > 64 ? 16 * S::R : 64;
	if (rows < min_rows) rows = min_rows;
	if (rows > a.y_rows) rows = a.y_rows;
	rows = (rows + 1) & ~1;
	bands = (a.y_rows + rows - 1) / rows;
	a.band_rows = rows;
	const size_t smem = (size_t)K::NRING * K::PITCH * sizeof(float);
	static_assert((size_t)K::NRING * K::PITCH * sizeof(float) <= 48 * 1024, "ring must fit the default 48 KB");
	dim3 grid(strips, bands, planes);
	k_march<S, C, ISMAX><<<grid, K::NT, smem, st>>>(a);
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
