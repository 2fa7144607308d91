// k_runs.cu -- erosion / dilation family for ANY "row-run" element up to reach 32 whose
// shape is not one of the compile-time shapes of k_disk: disk6.5, disk16 ... disk32.9,
// N x M rectangles given as user lists, hrecN / vrecN of moderate length.
// build_disk() takes any float radius (src/morsi.c:313-330,499), so the shape is a RUNTIME
// table here: half-width hw(dy) of the centred run on each row dy = -R..R.
//
// The same streaming march as k_disk -- per input row the nested horizontal extrema
// H_k(x) = ext(in[x-k..x+k]) by an FMNMX3 chain in registers, then one fold of
// H_{hw(dy)} into the accumulator of every output row the input row touches -- but the
// 2R+1 accumulators per column live in a per-thread slice of SHARED memory instead of
// registers, and so do the H_k of the current row (a fold is 2 LDS.128 + 4 FMNMX +
// STS.128, the half-width of each window row a run-time table), which is what frees the
// shape from compile time: only the chain length is a template parameter, in steps of 4.
// Shared-memory bandwidth is the bound (3 (2R+1) + RX 128-bit accesses per 4 samples).
// Measured on B200, disk20 erosion of 4096x4096x3: k_tiled 7.75 ms; folds ordered by
// half-width with the H_k in registers 1.90 ms (one dependent load-fold-store chain); this
// form 1.72 ms; H_k in registers picked through a jump table 2.27 ms.
// The folds of one input row touch 2R+1 different accumulators; they are issued in
// batches of four (eight independent loads in flight), because the occupancy is capped
// by the shared memory the accumulators take and latency has to be hidden inside a warp.
// A thread owns 4 adjacent columns and reads its window of the input row straight
// from global memory (float4, L1/L2 absorb the overlap between neighbours): threads
// never exchange data, so there is no barrier anywhere in the march.
// About R + 2.5 (2R+1) instructions per sample instead of n = e[0] gathers
// (disk20: ~115 vs 1257).  Rows / columns outside the image are absent (NaN,
// src/morsi.c:30-35); a -0.0 among the samples raises *flag.
#include <climits>
#include <cstdlib>
#include <cstring>

#include "dispatch.cuh"

#define RUNS_MAXR 32
#define RUNS_NT 32                       // threads per CTA (one warp); each owns 4 columns

struct RunsShape {
	int R, RX;
	signed char krow[2 * RUNS_MAXR + 1];   // half-width of window row dy = R - i, i = 0 .. 2R
};

struct RunsArgs {
	Band src, xop, other;
	float *y;
	long long y_pstride;
	int y_row0, y_rows;
	int src_row1;       // one past the last row src holds
	int w, h;
	int band_rows;
	int epi;
	int *flag;
	RunsShape s;
};

template <bool ISMAX> __device__ __forceinline__ float rext2(float a, float b) { return ISMAX ? fmaxf(a, b) : fminf(a, b); }
template <bool ISMAX> __device__ __forceinline__ float rext3(float a, float b, float c)
{
	return ISMAX ? fmaxf(fmaxf(a, b), c) : fminf(fminf(a, b), c);
}

__device__ __noinline__ float4 runs_epi4(int epi, bool ismax, float4 m, float4 o, float4 x)
{
	const float mm[4] = {m.x, m.y, m.z, m.w}, ov[4] = {o.x, o.y, o.z, o.w}, xv[4] = {x.x, x.y, x.z, x.w};
	float out[4] = {m.x, m.y, m.z, m.w};
#define CASE(E) case E: _Pragma("unroll") for (int c = 0; c < 4; c++) \
		out[c] = ismax ? epilogue<E>(ov[c], mm[c], xv[c]) : epilogue<E>(mm[c], ov[c], xv[c]); break;
	switch (epi) {
	CASE(EPI_B_SUB_A) CASE(EPI_X_SUB_A) CASE(EPI_B_SUB_X) CASE(EPI_LAP) CASE(EPI_ENH) CASE(EPI_BLUR)
	CASE(EPI_A_SUB_B) CASE(EPI_X_SUB_B) CASE(EPI_A_SUB_X) CASE(EPI_IBLUR) CASE(EPI_EBLUR) CASE(EPI_CBLUR)
	default: break;
	}
#undef CASE
	return make_float4(out[0], out[1], out[2], out[3]);
}

// KQ: the chain runs to half-width KMAX = 4 KQ >= RX.  VEC: rows are 16-byte aligned (w % 4 == 0).
template <int KQ, bool ISMAX, bool VEC>
__global__ void __launch_bounds__(RUNS_NT) k_runs(RunsArgs p)
{
	constexpr int KMAX = 4 * KQ, LH = KMAX, NV = 4 + 2 * KMAX;
	extern __shared__ float4 runs_acc[];               // [NS][RUNS_NT] accumulators, then [KMAX+1][RUNS_NT] H_k of the current row
	const int tid = threadIdx.x;
	const int plane = blockIdx.z;
	const int R = p.s.R, NS = 2 * R + 1;
	const int x0 = (blockIdx.x * RUNS_NT + tid) * 4;     // my 4 columns
	const int o_base = blockIdx.y * p.band_rows;
	const int nout = min(p.band_rows, p.y_rows - o_base);
	const int Y0 = p.y_row0 + o_base;                  // global row of my first output
	const int w = p.w, h = p.h;
	const float init = ISMAX ? -CUDART_INF_F : CUDART_INF_F;
	const bool col_ok = x0 < w;
	const float *sp = p.src.p + plane * p.src.pstride;
	float4 *acc = runs_acc + tid;
	float4 *hs = runs_acc + NS * RUNS_NT + tid;
	bool negzero = false;
	// input rows Y0-R .. Y0+nout-1+R; `base` = ring slot of the output row that equals the input row
	int base = 0;
#pragma unroll 1
	for (int r = Y0 - R; r < Y0 + nout + R; r++) {
		const bool row_ok = r >= 0 && r < h && r >= p.src.row0 && r < p.src_row1;
		if (row_ok) {
			float v[NV];
			const float *rp = sp + (long long)(r - p.src.row0) * w;
			const int c0 = x0 - LH;                           // first column of my window (a multiple of 4)
#pragma unroll
			for (int q = 0; q < NV / 4; q++) {
				const int c = c0 + 4 * q;
				if (VEC && c >= 0 && c + 3 < w) {
					const float4 t = __ldg((const float4 *)(rp + c));
					v[4*q] = t.x; v[4*q+1] = t.y; v[4*q+2] = t.z; v[4*q+3] = t.w;
				} else {
#pragma unroll
					for (int i = 0; i < 4; i++)
						v[4*q+i] = (c + i >= 0 && c + i < w) ? __ldg(rp + c + i) : CUDART_NAN_F;
				}
			}
			float H[4];
#pragma unroll
			for (int c = 0; c < 4; c++) {
				negzero |= __float_as_uint(v[LH + c]) == 0x80000000u;
				H[c] = v[LH + c];
			}
			hs[0] = make_float4(H[0], H[1], H[2], H[3]);
#pragma unroll
			for (int k = 1; k <= KMAX; k++) {
#pragma unroll
				for (int c = 0; c < 4; c++) H[c] = rext3<ISMAX>(H[c], v[LH + c - k], v[LH + c + k]);
				hs[k * RUNS_NT] = make_float4(H[0], H[1], H[2], H[3]);
			}
		} else {
			// an absent row: nothing to contribute (NaN is ignored by min / max)
			const float nan = CUDART_NAN_F;
#pragma unroll
			for (int k = 0; k <= KMAX; k++) hs[k * RUNS_NT] = make_float4(nan, nan, nan, nan);
		}
		// fold H_{hw(dy)} into the output row o = r - dy for dy = R .. -R (i = 0 .. 2R), slots
		// consecutive from (base - R) mod NS.  The row dy = -R of a window is the first to
		// arrive: it starts the accumulator.  Batches of four: the loads of a batch are independent.
		int sl = base - R; if (sl < 0) sl += NS;
		float4 m = make_float4(init, init, init, init);                 // the completed row (i = 0)
#pragma unroll 1
		for (int i0 = 0; i0 < NS; i0 += 4) {
			float4 a[4], hk[4];
			int slot[4];
#pragma unroll
			for (int j = 0; j < 4; j++) {
				const int i = min(i0 + j, NS - 1);
				slot[j] = sl;
				if (i0 + j < NS - 1) { if (++sl == NS) sl = 0; }
				hk[j] = hs[(int)p.s.krow[i] * RUNS_NT];
				a[j] = acc[slot[j] * RUNS_NT];
			}
#pragma unroll
			for (int j = 0; j < 4; j++) {
				const int i = i0 + j;
				if (i < NS) {
					if (i == NS - 1) a[j] = make_float4(init, init, init, init);    // first row of that window
					a[j].x = rext2<ISMAX>(a[j].x, hk[j].x); a[j].y = rext2<ISMAX>(a[j].y, hk[j].y);
					a[j].z = rext2<ISMAX>(a[j].z, hk[j].z); a[j].w = rext2<ISMAX>(a[j].w, hk[j].w);
					acc[slot[j] * RUNS_NT] = a[j];
					if (i == 0) m = a[j];
				}
			}
		}
		// output row o = r - R is complete
		const int o = r - R;
		if (o >= Y0 && col_ok) {
			const long long off = (long long)(o - p.y_row0) * w + x0;
			if (p.epi != (ISMAX ? EPI_B : EPI_A)) {
				float4 ov = make_float4(0.f, 0.f, 0.f, 0.f), xv = ov;
				float *ovp = &ov.x, *xvp = &xv.x;
#pragma unroll
				for (int c = 0; c < 4; c++) {
					if (x0 + c < w) {
						if (p.other.p) ovp[c] = __ldg(p.other.p + plane * p.other.pstride + (long long)(o - p.other.row0) * w + x0 + c);
						if (p.xop.p) xvp[c] = __ldg(p.xop.p + plane * p.xop.pstride + (long long)(o - p.xop.row0) * w + x0 + c);
					}
				}
				m = runs_epi4(p.epi, ISMAX, m, ov, xv);
			}
			float *yp = p.y + plane * p.y_pstride + off;
			if (VEC && x0 + 3 < w) *(float4 *)yp = m;
			else {
				const float *mp = &m.x;
#pragma unroll
				for (int c = 0; c < 4; c++) if (x0 + c < w) yp[c] = mp[c];
			}
		}
		if (++base == NS) base = 0;
	}
	if (__syncthreads_or(negzero) && tid == 0) atomicOr(p.flag, 1);
}

// ---- host side ---------------------------------------------------------------------------
template <int KQ>
static int runs_launch_kq(const RunsArgs &a, bool ismax, bool vec, dim3 grid, size_t smem, int device, cudaStream_t st)
{
	int rc = 0;
#define L(M, V) do { if ((rc = morsi_optin_smem((const void *)k_runs<KQ, M, V>, device, 160 * 1024))) return rc; \
		k_runs<KQ, M, V><<<grid, RUNS_NT, smem, st>>>(a); } while (0)
	if (ismax) { if (vec) L(true, true); else L(true, false); }
	else { if (vec) L(false, true); else L(false, false); }
#undef L
	morsi_count_launch(1);
	MORSI_CU(cudaGetLastError());
	return MORSI_OK;
}

static int runs_pass(MorsiCtx *c, const RunsShape &sh, bool ismax, int epi, const MorsiJob &job, Band src, int src_rows,
		Band xop, Band other, float *dst, long long dst_pstride, int row0, int rows, int *flag)
{
	RunsArgs a;
	a.src = src; a.src_row1 = src.row0 + src_rows; a.xop = xop; a.other = other;
	a.y = dst; a.y_pstride = dst_pstride; a.y_row0 = row0; a.y_rows = rows;
	a.w = job.w; a.h = job.h; a.epi = epi; a.flag = flag; a.s = sh;
	const int strips = (job.w + 4 * RUNS_NT - 1) / (4 * RUNS_NT);
	// bands: enough CTAs for ~2 per SM slot, each marching at least 8 R rows beyond its 2 R warm-up
	const int kmax = 4 * ((sh.RX + 3) / 4 > 0 ? (sh.RX + 3) / 4 : 1);
	const size_t smem = (size_t)(2 * sh.R + 1 + kmax + 1) * RUNS_NT * sizeof(float4);
	int per_sm = (int)((200 * 1024) / (smem + 1024)); if (per_sm > 16) per_sm = 16; if (per_sm < 1) per_sm = 1;
	long long want = 2LL * c->sm_count * per_sm / ((long long)strips * job.planes);
	if (want < 1) want = 1;
	int band = (int)((rows + want - 1) / want);
	const int min_band = 8 * sh.R > 64 ? 8 * sh.R : 64;
	if (band < min_band) band = min_band;
	if (band > rows) band = rows;
	a.band_rows = band;
	const bool vec = job.w % 4 == 0 && ((uintptr_t)src.p % 16 == 0) && ((uintptr_t)dst % 16 == 0) &&
		src.pstride % 4 == 0 && dst_pstride % 4 == 0;
	for (int b0 = 0; b0 * (long long)band < rows; b0 += 65535) {          // gridDim.y limit
		RunsArgs s = a;
		const long long done = (long long)b0 * band;
		s.y_row0 = row0 + (int)done; s.y_rows = (int)((rows - done) < 65535LL * band ? rows - done : 65535LL * band);
		s.y = dst + done * job.w;
		dim3 grid(strips, (s.y_rows + band - 1) / band, job.planes);
		int rc;
		switch ((sh.RX + 3) / 4) {
		case 0: case 1: rc = runs_launch_kq<1>(s, ismax, vec, grid, smem, c->device, job.stream); break;
		case 2: rc = runs_launch_kq<2>(s, ismax, vec, grid, smem, c->device, job.stream); break;
		case 3: rc = runs_launch_kq<3>(s, ismax, vec, grid, smem, c->device, job.stream); break;
		case 4: rc = runs_launch_kq<4>(s, ismax, vec, grid, smem, c->device, job.stream); break;
		case 5: rc = runs_launch_kq<5>(s, ismax, vec, grid, smem, c->device, job.stream); break;
		case 6: rc = runs_launch_kq<6>(s, ismax, vec, grid, smem, c->device, job.stream); break;
		case 7: rc = runs_launch_kq<7>(s, ismax, vec, grid, smem, c->device, job.stream); break;
		default: rc = runs_launch_kq<8>(s, ismax, vec, grid, smem, c->device, job.stream); break;
		}
		if (rc) return rc;
	}
	return MORSI_OK;
}

// The passes of one operation (src/morsi.c:141-275) over row-run kernels; temporaries in the
// workspace, bounded by row chunking.
int morsi_run_runs(MorsiCtx *c, const DevElement *de, const MorsiJob &job, int *flag, int *handled)
{
	*handled = 0;
	const OpPlan plan = morsi_op_plan(job.op);
	static const bool off = getenv("MORSI_RUNS") && !strcmp(getenv("MORSI_RUNS"), "0");
	if (off || plan.special || !de->rowrun.ok || de->n < 12) return MORSI_OK;
	const int R = de->rowrun.reach;
	// one-row and tall thin elements (hrecN, vrecN): k_tiled is faster there (hrec31 dilation of a 4096x4096x3
	// image: 0.45 vs 0.96 ms measured), the march pays off for two-dimensional shapes
	if (R < 2 || de->n < 3 * (2 * R + 1)) return MORSI_OK;
	if (R > RUNS_MAXR) return MORSI_OK;
	RunsShape sh;
	sh.R = R; sh.RX = de->rowrun.hw[R];
	if (sh.RX > RUNS_MAXR) return MORSI_OK;
	for (int i = 0; i <= 2 * R; i++) {
		const int k = de->rowrun.hw[2 * R - i];                 // window row dy = R - i
		if (k < 0 || k > RUNS_MAXR) return MORSI_OK;
		sh.krow[i] = (signed char)k;
	}
	if (job.planes > 65535) return MORSI_OK;
	const int w = job.w;
	const Band none{nullptr, 0, 0};
	const Band xb{job.x, job.x_row0, job.x_pstride};
	const bool both1 = plan.stages == 1 && plan.a_from && plan.b_from;
	const bool osc = plan.t_min && plan.t_max;
	int rc;
	if (plan.stages == 1 && !both1) {
		rc = runs_pass(c, sh, plan.b_from != 0, plan.epi, job, xb, job.x_rows, xb, none, job.y, job.y_pstride, job.y_row0, job.y_rows, flag);
		if (rc) return rc;
		*handled = 1;
		return MORSI_OK;
	}
	// every other operation needs temporaries: chunk the band so that one stays below 256 MiB
	long long rows_fit = (256LL << 20) / ((long long)w * 4 * job.planes) - 2 * R;
	if (rows_fit < 8 * R + 64) rows_fit = 8 * R + 64;
	const int chunk = (int)(rows_fit < job.y_rows ? rows_fit : job.y_rows);
	for (int r0 = 0; r0 < job.y_rows; r0 += chunk) {
		const int o0 = job.y_row0 + r0;
		const int orows = job.y_rows - r0 < chunk ? job.y_rows - r0 : chunk;
		float *ydst = job.y + (long long)r0 * w;
		const long long ops = (long long)w * orows;
		if (both1) {
			void *p0; if ((rc = morsi_ws_get(c, job.lane, 0, (size_t)ops * job.planes * 4, &p0))) return rc;
			if ((rc = runs_pass(c, sh, false, EPI_A, job, xb, job.x_rows, none, none, (float *)p0, ops, o0, orows, flag))) return rc;
			if ((rc = runs_pass(c, sh, true, plan.epi, job, xb, job.x_rows, xb, Band{(float *)p0, o0, ops}, ydst, job.y_pstride, o0, orows, flag))) return rc;
			continue;
		}
		// two stages: the first over the output rows grown by one reach (clipped to the image)
		int t0 = o0 - R; if (t0 < 0) t0 = 0;
		int t1 = o0 + orows + R; if (t1 > job.h) t1 = job.h;
		const int trows = t1 - t0;
		const long long tps = (long long)w * trows;
		void *p0, *p1 = nullptr;
		if ((rc = morsi_ws_get(c, job.lane, 0, (size_t)tps * job.planes * 4, &p0))) return rc;
		if (!osc) {
			const bool first_max = plan.t_max != 0;                  // closing / bothat: dilation first
			if ((rc = runs_pass(c, sh, first_max, first_max ? EPI_B : EPI_A, job, xb, job.x_rows, none, none, (float *)p0, tps, t0, trows, flag))) return rc;
			const Band xop = (plan.epi == EPI_X_SUB_B || plan.epi == EPI_A_SUB_X) ? xb : none;
			if ((rc = runs_pass(c, sh, !first_max, plan.epi, job, Band{(float *)p0, t0, tps}, trows, xop, none, ydst, job.y_pstride, o0, orows, flag))) return rc;
			continue;
		}
		// oscillation = closing - opening
		if ((rc = morsi_ws_get(c, job.lane, 1, (size_t)ops * job.planes * 4, &p1))) return rc;
		if ((rc = runs_pass(c, sh, true, EPI_B, job, xb, job.x_rows, none, none, (float *)p0, tps, t0, trows, flag))) return rc;
		if ((rc = runs_pass(c, sh, false, EPI_A, job, Band{(float *)p0, t0, tps}, trows, none, none, (float *)p1, ops, o0, orows, flag))) return rc;   // closing
		if ((rc = runs_pass(c, sh, false, EPI_A, job, xb, job.x_rows, none, none, (float *)p0, tps, t0, trows, flag))) return rc;
		if ((rc = runs_pass(c, sh, true, EPI_A_SUB_B, job, Band{(float *)p0, t0, tps}, trows, none, Band{(float *)p1, o0, ops}, ydst, job.y_pstride, o0, orows, flag))) return rc;   // closing - opening
	}
	*handled = 1;
	return MORSI_OK;
}
