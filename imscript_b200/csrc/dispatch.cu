// dispatch.cu -- picks a kernel family for (operation, element, image) and
// runs the passes.  The GPU counterpart of the function-pointer dispatch and
// composite wrappers of src/morsi.c:141-275,509-543.
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include <mutex>

#include "dispatch.cuh"
#include "k_exact.cuh"
#include "k_tiled.cuh"
#include "k_line.cuh"

OpPlan morsi_op_plan(int op)
{
	//                         stages tmin tmax a  b  epi          special
	switch (op) {
	case MORSI_EROSION:     return {1, 0, 0, 1, 0, EPI_A, 0};
	case MORSI_DILATION:    return {1, 0, 0, 0, 1, EPI_B, 0};
	case MORSI_MEDIAN:      return {1, 0, 0, 0, 0, EPI_A, 1};
	case MORSI_RANK:        return {1, 0, 0, 0, 0, EPI_A, 2};
	case MORSI_OPENING:     return {2, 1, 0, 0, 2, EPI_B, 0};
	case MORSI_CLOSING:     return {2, 0, 1, 3, 0, EPI_A, 0};
	case MORSI_GRADIENT:    return {1, 0, 0, 1, 1, EPI_B_SUB_A, 0};
	case MORSI_IGRADIENT:   return {1, 0, 0, 1, 0, EPI_X_SUB_A, 0};
	case MORSI_EGRADIENT:   return {1, 0, 0, 0, 1, EPI_B_SUB_X, 0};
	case MORSI_LAPLACIAN:   return {1, 0, 0, 1, 1, EPI_LAP, 0};
	case MORSI_ENHANCE:     return {1, 0, 0, 1, 1, EPI_ENH, 0};
	case MORSI_BLUR:        return {1, 0, 0, 1, 1, EPI_BLUR, 0};
	case MORSI_OSCILLATION: return {2, 1, 1, 3, 2, EPI_A_SUB_B, 0};
	case MORSI_TOPHAT:      return {2, 1, 0, 0, 2, EPI_X_SUB_B, 0};
	case MORSI_BOTHAT:      return {2, 0, 1, 3, 0, EPI_A_SUB_X, 0};
	case MORSI_IBLUR:       return {1, 0, 0, 1, 0, EPI_IBLUR, 0};
	case MORSI_EBLUR:       return {1, 0, 0, 0, 1, EPI_EBLUR, 0};
	case MORSI_CBLUR:       return {1, 0, 0, 1, 1, EPI_CBLUR, 0};
	}
	return {0, 0, 0, 0, 0, 0, 0};
}

void morsi_element_compile(MorsiCtx *, DevElement *d)
{
	d->rowrun.ok = 0;
	if (d->info.kind == MORSI_EK_ROWRUN || (d->info.kind == MORSI_EK_SMALL && d->info.reach > 0)) {
		d->rowrun.ok = 1;
		d->rowrun.reach = d->info.reach;
		for (int k = 0; k <= 2 * d->info.reach; k++) d->rowrun.hw[k] = d->info.halfwidth[k];
	}
}

// Opt a kernel in to > 48 KB of dynamic shared memory.  The attribute is per
// DEVICE: remember (kernel, device) pairs, under a lock (morsi_cuda_apply runs
// one host thread per device with MORSI_CUDA_DEVICES > 1).
int morsi_optin_smem(const void *kernel, int device, int bytes)
{
	static std::mutex mu;
	static std::map<const void *, unsigned long long> done;
	std::lock_guard<std::mutex> lk(mu);
	unsigned long long &m = done[kernel];
	const unsigned long long bit = 1ull << (device & 63);
	if (m & bit) return MORSI_OK;
	MORSI_CU(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
	m |= bit;
	return MORSI_OK;
}

// Gated launches (re-runs that are no-ops unless a -0.0 was seen) use a small
// grid-stride grid so that the usual no-op costs a few microseconds.
static dim3 exact_grid(int w, int rows, int planes, bool gated)
{
	unsigned gx = (unsigned)((w + 31) / 32);
	unsigned gy = (unsigned)((rows + 7) / 8);
	if (gy > 16384) gy = 16384;
	if (gated) {
		unsigned long long cap = 148ull * 4;   // a no-op grid costs per CTA: keep it to a few per SM
		unsigned long long per_row = (unsigned long long)gx * (unsigned)planes;
		unsigned want = (unsigned)(cap / (per_row ? per_row : 1));
		if (want < 1) want = 1;
		if (gy > want) gy = want;
	}
	return dim3(gx, gy, (unsigned)planes);
}

template <int EPI>
static void launch_exact_t(const ExactArgs &a, int planes, cudaStream_t s)
{
	k_exact_minmax<EPI><<<exact_grid(a.w, a.y_rows, planes, a.gate != nullptr), dim3(32, 8), 0, s>>>(a);
	morsi_count_launch(1);
}

// the tiled kernels for arbitrary lists (k_tiled.cuh); tg == nullptr: the exact ones
struct TiledLaunch {
	TiledGeom g;
	int *flag;
	size_t smem;
	const LineGeom *line = nullptr;   // non-NULL: the van Herk line kernels (k_line.cuh) instead
	size_t line_smem = 0;
	int exact = 0;                    // 1: the order-preserving form of k_tiled_minmax (the gated re-run)
};

template <int EPI>
static int launch_tiled_t(const ExactArgs &a, const TiledLaunch &t, int planes, int device, cudaStream_t s)
{
	int rc;
	if (t.line) {
		if ((rc = morsi_optin_smem((const void *)k_line_minmax<EPI>, device, 200 * 1024))) return rc;
		const LineGeom &g = *t.line;
		const int per_y = g.vertical ? g.nb * g.L : g.nl;     // output rows per CTA
		const int per_x = g.vertical ? g.nl : g.nb * g.L;
		const int blocks_y = (a.y_rows + per_y - 1) / per_y;
		for (int r0 = 0; r0 < blocks_y; r0 += 65535) {        // gridDim.y limit
			ExactArgs sub = a;
			const int nb = blocks_y - r0 < 65535 ? blocks_y - r0 : 65535;
			sub.y_row0 = a.y_row0 + r0 * per_y;
			sub.y_rows = a.y_rows - r0 * per_y < (long long)nb * per_y ? a.y_rows - r0 * per_y : nb * per_y;
			sub.y = a.y + (long long)r0 * per_y * a.w;
			if (a.y2) sub.y2 = a.y2 + (long long)r0 * per_y * a.w;
			dim3 grid((a.w + per_x - 1) / per_x, nb, planes);
			k_line_minmax<EPI><<<grid, 256, t.line_smem, s>>>(sub, g, t.flag);
			morsi_count_launch(1);
		}
		MORSI_CU(cudaGetLastError());
		return MORSI_OK;
	}
	if (t.exact) {
		if ((rc = morsi_optin_smem((const void *)k_tiled_minmax<EPI, true>, device, 200 * 1024))) return rc;
		// one launch, tile rows walked inside the kernel; a gated launch (a no-op nearly always) stays small
		const int rows_y = (a.y_rows + TILED_TY - 1) / TILED_TY;
		const unsigned gx = (unsigned)((a.w + TILED_TX - 1) / TILED_TX);
		unsigned gy = (unsigned)(rows_y < 65535 ? rows_y : 65535);
		if (a.gate) {
			const unsigned long long per_row = (unsigned long long)gx * (unsigned)planes;
			unsigned want = (unsigned)(148ull * 4 / (per_row ? per_row : 1));
			if (want < 1) want = 1;
			if (gy > want) gy = want;
		}
		k_tiled_minmax<EPI, true><<<dim3(gx, gy, planes), dim3(32, 8), t.smem, s>>>(a, t.g, t.flag);
		morsi_count_launch(1);
		MORSI_CU(cudaGetLastError());
		return MORSI_OK;
	}
	if ((rc = morsi_optin_smem((const void *)k_tiled_minmax<EPI, false>, device, 200 * 1024))) return rc;
	const int rows_y = (a.y_rows + TILED_TY - 1) / TILED_TY;
	for (int r0 = 0; r0 < rows_y; r0 += 65535) {   // gridDim.y limit
		ExactArgs sub = a;
		const int nb = rows_y - r0 < 65535 ? rows_y - r0 : 65535;
		sub.y_row0 = a.y_row0 + r0 * TILED_TY;
		sub.y_rows = a.y_rows - r0 * TILED_TY < nb * TILED_TY ? a.y_rows - r0 * TILED_TY : nb * TILED_TY;
		sub.y = a.y + (long long)r0 * TILED_TY * a.w;
		if (a.y2) sub.y2 = a.y2 + (long long)r0 * TILED_TY * a.w;
		dim3 grid((a.w + TILED_TX - 1) / TILED_TX, nb, planes);
		k_tiled_minmax<EPI, false><<<grid, dim3(32, 8), t.smem, s>>>(sub, t.g, t.flag);
		morsi_count_launch(1);
	}
	MORSI_CU(cudaGetLastError());
	return MORSI_OK;
}

static int launch_tiled(int epi, const ExactArgs &a, const TiledLaunch &t, int planes, int device, cudaStream_t s)
{
	switch (epi) {
#define C(E) case E: return launch_tiled_t<E>(a, t, planes, device, s);
	C(EPI_A) C(EPI_B) C(EPI_B_SUB_A) C(EPI_X_SUB_A) C(EPI_B_SUB_X) C(EPI_LAP) C(EPI_ENH)
	C(EPI_BLUR) C(EPI_A_SUB_B) C(EPI_X_SUB_B) C(EPI_A_SUB_X) C(EPI_IBLUR) C(EPI_EBLUR)
	C(EPI_CBLUR) C(EPI_AB)
#undef C
	}
	return morsi_set_error(MORSI_ERR_INVALID, "bad epilogue %d", epi);
}

static int launch_exact(int epi, const ExactArgs &a, int planes, cudaStream_t s)
{
	switch (epi) {
#define C(E) case E: launch_exact_t<E>(a, planes, s); break;
	C(EPI_A) C(EPI_B) C(EPI_B_SUB_A) C(EPI_X_SUB_A) C(EPI_B_SUB_X) C(EPI_LAP) C(EPI_ENH)
	C(EPI_BLUR) C(EPI_A_SUB_B) C(EPI_X_SUB_B) C(EPI_A_SUB_X) C(EPI_IBLUR) C(EPI_EBLUR)
	C(EPI_CBLUR) C(EPI_AB)
#undef C
	default: return morsi_set_error(MORSI_ERR_INVALID, "bad epilogue %d", epi);
	}
	MORSI_CU(cudaGetLastError());
	return MORSI_OK;
}

// The order-preserving path for every operation: one or two passes of
// k_exact_minmax, temporaries in the context workspace.  `gate` (device word)
// makes every kernel a no-op unless it is non-zero.
static int run_minmax_passes(MorsiCtx *c, const DevElement *de, const MorsiJob &job, const int *gate, const TiledLaunch *tl);

int morsi_run_exact(MorsiCtx *c, const DevElement *de, const MorsiJob &job, const int *gate)
{
	return run_minmax_passes(c, de, job, gate, nullptr);
}

// tl == nullptr: k_exact_* (order preserving); else k_tiled_minmax (min/max operations only)
static int run_minmax_passes(MorsiCtx *c, const DevElement *de, const MorsiJob &job, const int *gate, const TiledLaunch *tl)
{
	const OpPlan plan = morsi_op_plan(job.op);
	const int w = job.w, h = job.h;
	Band xb{job.x, job.x_row0, job.x_pstride};

	if (plan.special) {
		MedianArgs m;
		m.x_src = xb; m.y = job.y; m.y_pstride = job.y_pstride;
		m.y_row0 = job.y_row0; m.y_rows = job.y_rows; m.w = w; m.h = h;
		m.offs = de->d_offs; m.n = de->n; m.gate = gate;
		dim3 g = exact_grid(w, job.y_rows, job.planes, gate != nullptr);
		if (plan.special == 1) k_exact_median<<<g, dim3(32, 8), 0, job.stream>>>(m);
		else k_exact_rank<<<g, dim3(32, 8), 0, job.stream>>>(m);
		morsi_count_launch(1);
		MORSI_CU(cudaGetLastError());
		return MORSI_OK;
	}

	ExactArgs a;
	a.x_src = xb; a.w = w; a.h = h; a.offs = de->d_offs; a.n = de->n; a.gate = gate;
	a.y2 = nullptr;
	Band tmin{nullptr, 0, 0}, tmax{nullptr, 0, 0};
	if (plan.stages == 2) {
		// stage 1 over the output band grown by one reach, clipped to the image
		const int up = de->n ? (de->info.ymin < 0 ? -de->info.ymin : 0) : 0;
		const int dn = de->n ? (de->info.ymax > 0 ? de->info.ymax : 0) : 0;
		int t0 = job.y_row0 - up; if (t0 < 0) t0 = 0;
		long long t1 = (long long)job.y_row0 + job.y_rows + dn; if (t1 > h) t1 = h;
		const int t_rows = (int)(t1 - t0);
		const long long tps = (long long)w * t_rows;
		const size_t bytes = (size_t)tps * job.planes * sizeof(float);
		void *p0 = nullptr, *p1 = nullptr;
		int rc = morsi_ws_get(c, job.lane, 0, bytes, &p0);
		if (rc) return rc;
		if (plan.t_min && plan.t_max) { rc = morsi_ws_get(c, job.lane, 1, bytes, &p1); if (rc) return rc; }
		ExactArgs s1 = a;
		s1.a_src = xb; s1.b_src = xb;
		s1.y_row0 = t0; s1.y_rows = t_rows; s1.y_pstride = tps;
		int epi1;
		if (plan.t_min && plan.t_max) { s1.y = (float *)p0; s1.y2 = (float *)p1; epi1 = EPI_AB;
			tmin = Band{(float *)p0, t0, tps}; tmax = Band{(float *)p1, t0, tps}; }
		else if (plan.t_min) { s1.y = (float *)p0; epi1 = EPI_A; tmin = Band{(float *)p0, t0, tps}; }
		else { s1.y = (float *)p0; epi1 = EPI_B; tmax = Band{(float *)p0, t0, tps}; }
		if (tl) { TiledLaunch t1 = *tl; t1.g.two_tiles = 0; rc = launch_tiled(epi1, s1, t1, job.planes, c->device, job.stream); }
		else rc = launch_exact(epi1, s1, job.planes, job.stream);
		if (rc) return rc;
	}
	const Band *srcs[4] = {&xb, &xb, &tmin, &tmax};
	a.a_src = *srcs[plan.a_from]; a.b_src = *srcs[plan.b_from];
	a.y = job.y; a.y_pstride = job.y_pstride; a.y_row0 = job.y_row0; a.y_rows = job.y_rows;
	if (tl) {
		TiledLaunch t2 = *tl;
		t2.g.two_tiles = !tl->line && plan.a_from && plan.b_from && a.a_src.p != a.b_src.p;
		if (t2.g.two_tiles) t2.smem += (size_t)t2.g.pw * t2.g.ph * sizeof(float);
		if (t2.smem > 200 * 1024) return morsi_set_error(MORSI_ERR_INVALID, "tiled: element too large for two tiles");
		return launch_tiled(plan.epi, a, t2, job.planes, c->device, job.stream);
	}
	return launch_exact(plan.epi, a, job.planes, job.stream);
}

// Exact path with bounded temporaries: the job is cut into plane groups and
// row chunks so that a two-stage temporary never exceeds ~256 MiB per slot.
static int run_exact_chunked(MorsiCtx *c, const DevElement *de, const MorsiJob &job, const int *gate, const TiledLaunch *tl = nullptr)
{
	const OpPlan plan = morsi_op_plan(job.op);
	// the gated re-run is a no-op nearly always: fewer, larger chunks keep its launch count down
	// (C4: 6 launches per step instead of 50)
	const long long budget = (gate ? 2048LL : 256LL) << 20;
	const int halo = plan.stages == 2 ? (de->info.ymax - de->info.ymin + 1) : 0;
	long long row_bytes = (long long)job.w * 4;
	long long rows_fit = budget / row_bytes - halo;
	if (rows_fit < 16) rows_fit = 16;
	int planes_per = 1, rows_per = job.y_rows;
	if (plan.stages != 2) {
		planes_per = job.planes;                       // no temporaries at all
	} else if (rows_fit >= job.y_rows) {
		long long pp = budget / (row_bytes * (job.y_rows + halo));
		planes_per = (int)(pp < 1 ? 1 : pp > job.planes ? job.planes : pp);
	} else {
		rows_per = (int)rows_fit;
	}
	for (int p0 = 0; p0 < job.planes; p0 += planes_per)
		for (int r0 = 0; r0 < job.y_rows; r0 += rows_per) {
			MorsiJob sub = job;
			sub.planes = job.planes - p0 < planes_per ? job.planes - p0 : planes_per;
			sub.x = job.x + p0 * job.x_pstride;
			sub.y = job.y + p0 * job.y_pstride + (long long)r0 * job.w;
			sub.y_row0 = job.y_row0 + r0;
			sub.y_rows = job.y_rows - r0 < rows_per ? job.y_rows - r0 : rows_per;
			int rc = run_minmax_passes(c, de, sub, gate, tl);
			if (rc) return rc;
		}
	return MORSI_OK;
}

// Long line elements (hrecR / vrecR and any other list that is one contiguous run on a
// single row or column), length 96 ... 320: van Herk / Gil-Werman, 3 compares per sample
// whatever the length.  Measured on B200 (4096x4096x3 dilation): 0.6-0.9 ms for every
// length, against 0.0065 ms x length for k_tiled (hrec150, 299 samples: 0.86 vs 1.92 ms;
// hrec40, 79 samples: 0.78 vs 0.55 ms), hence the threshold.
int morsi_run_line(MorsiCtx *c, const DevElement *de, const MorsiJob &job, int *flag, int *handled)
{
	*handled = 0;
	const OpPlan plan = morsi_op_plan(job.op);
	const morsi_element_info &in = de->info;
	if (plan.special || de->n < 96 || de->n > 320 || in.has_duplicates) return MORSI_OK;
	static const bool off = getenv("MORSI_LINE") && !strcmp(getenv("MORSI_LINE"), "0");
	if (off) return MORSI_OK;
	const bool horizontal = in.ymin == in.ymax && in.xmax - in.xmin + 1 == de->n;
	const bool vertical = in.xmin == in.xmax && in.ymax - in.ymin + 1 == de->n;
	if (!horizontal && !vertical) return MORSI_OK;
	LineGeom g;
	g.vertical = vertical ? 1 : 0;
	g.L = de->n;
	g.lo = vertical ? in.ymin : in.xmin;
	g.perp = vertical ? in.xmin : in.ymin;
	g.nl = vertical ? 32 : 16;
	const int cap = vertical ? 256 : 640;                  // outputs per line and CTA (registers, shared memory)
	g.nb = cap / g.L < 1 ? 1 : cap / g.L;
	const int len = (g.nb + 1) * g.L;
	if (vertical) { g.sl = 1; g.sp = 32; }
	else { g.sp = 1; g.sl = len + ((33 - len % 32) % 32); }   // pitch = 1 (mod 32): the scans of 16 lines hit 16 banks
	g.two_sources = 0;
	if ((long long)g.nl * g.nb * g.L > 256LL * LINE_MAXOUT) return MORSI_OK;
	const size_t tile = (size_t)(vertical ? len * g.sp : g.nl * g.sl) * sizeof(float);
	if (2 * tile > 200 * 1024) return MORSI_OK;
	TiledLaunch t;
	t.g = TiledGeom{};
	t.flag = flag;
	t.smem = 0;
	t.line = &g;
	t.line_smem = 2 * tile;
	int rc = run_exact_chunked(c, de, job, nullptr, &t);
	if (rc) return rc;
	*handled = 1;
	return MORSI_OK;
}

// the tile geometry of an element for k_tiled_*; false when a tile does not fit shared memory
static bool tiled_setup(const DevElement *de, const OpPlan &plan, int *flag, TiledLaunch *t)
{
	if (de->n < 1 || de->n > 8192 || !de->d_tile_offs) return false;
	t->g.xmin = de->info.xmin; t->g.xmax = de->info.xmax; t->g.ymin = de->info.ymin; t->g.ymax = de->info.ymax;
	t->g.pw = TILED_TX + de->info.xmax - de->info.xmin;
	t->g.ph = TILED_TY + de->info.ymax - de->info.ymin;
	t->g.tile_offs = de->d_tile_offs;
	t->g.two_tiles = 0;
	t->flag = flag;
	t->smem = (size_t)t->g.pw * t->g.ph * sizeof(float) + (size_t)de->n * sizeof(int);
	// oscillation's last pass needs two tiles
	const size_t worst = t->smem + ((plan.t_min && plan.t_max) ? (size_t)t->g.pw * t->g.ph * sizeof(float) : 0);
	return worst <= 200 * 1024;
}

// Arbitrary lists the specialised families left: evaluate from shared-memory tiles.
int morsi_run_tiled(MorsiCtx *c, const DevElement *de, const MorsiJob &job, int *flag, int *handled)
{
	*handled = 0;
	const OpPlan plan = morsi_op_plan(job.op);
	if (plan.special == 1) return MORSI_OK;
	static const bool off = getenv("MORSI_TILED") && !strcmp(getenv("MORSI_TILED"), "0");
	if (off) return MORSI_OK;
	TiledLaunch t;
	if (!tiled_setup(de, plan, flag, &t)) return MORSI_OK;
	if (plan.special == 2) {
		// rank: one pass, no temporaries, exact in any order (no gated re-run needed)
		int rc = morsi_optin_smem((const void *)k_tiled_rank, c->device, 200 * 1024);
		if (rc) return rc;
		const int rows_y = (job.y_rows + TILED_TY - 1) / TILED_TY;
		for (int r0 = 0; r0 < rows_y; r0 += 65535) {
			const int nb = rows_y - r0 < 65535 ? rows_y - r0 : 65535;
			MedianArgs m;
			m.x_src = Band{job.x, job.x_row0, job.x_pstride};
			m.y = job.y + (long long)r0 * TILED_TY * job.w; m.y_pstride = job.y_pstride;
			m.y_row0 = job.y_row0 + r0 * TILED_TY;
			m.y_rows = job.y_rows - r0 * TILED_TY < nb * TILED_TY ? job.y_rows - r0 * TILED_TY : nb * TILED_TY;
			m.w = job.w; m.h = job.h; m.offs = de->d_offs; m.n = de->n; m.gate = nullptr;
			dim3 grid((job.w + TILED_TX - 1) / TILED_TX, nb, job.planes);
			k_tiled_rank<<<grid, dim3(32, 8), t.smem, job.stream>>>(m, t.g);
			morsi_count_launch(1);
		}
		MORSI_CU(cudaGetLastError());
		*handled = 6;                                  // complete: the caller skips the gated re-run
		return MORSI_OK;
	}
	int rc = run_exact_chunked(c, de, job, nullptr, &t);
	if (rc) return rc;
	*handled = 1;
	return MORSI_OK;
}

int morsi_dispatch(MorsiCtx *c, const int *e, const MorsiJob &job)
{
	const DevElement *de = nullptr;
	int rc = morsi_element_get(c, e, &de);
	if (rc) return rc;
	const int path = morsi_path();
	if (path != 1) {
		// fast families reorder the reduction; they raise *flag when the data
		// holds a -0.0, and the order-preserving kernels (no-ops while the
		// flag is clear) then redo the job
		int *flag = c->d_flag + job.lane;
		int handled = 0;
		rc = morsi_run_small(c, de, job, flag, &handled);   // clears the flag itself when its kernel uses it
		if (rc) return rc;
		if (!handled) MORSI_CU(cudaMemsetAsync(flag, 0, sizeof(int), job.stream));
		if (!handled) { rc = morsi_run_disk(c, de, job, flag, &handled); if (rc) return rc; if (handled) handled = 2; }
		if (!handled) { rc = morsi_run_runs(c, de, job, flag, &handled); if (rc) return rc; if (handled) handled = 8; }
		if (!handled) { rc = morsi_run_median3(c, de, job, flag, &handled); if (rc) return rc; if (handled) handled = 4; }
		if (!handled) { rc = morsi_run_median(c, de, job, flag, &handled); if (rc) return rc; if (handled) handled = 4; }
		if (!handled) { rc = morsi_run_line(c, de, job, flag, &handled); if (rc) return rc; if (handled) handled = 7; }
		if (!handled) { rc = morsi_run_tiled(c, de, job, flag, &handled); if (rc) return rc; if (handled == 1) handled = 5; }
		if (getenv("MORSI_CUDA_TRACE"))
			fprintf(stderr, "morsi_cuda: op %d n=%d %dx%dx%d rows [%d,+%d): %s\n", job.op, de->n, job.w, job.h, job.planes,
				job.y_row0, job.y_rows, handled == 1 ? "small" : handled == 2 ? "disk" :
				handled == 4 ? "median" : handled == 5 ? "tiled" : handled == 6 ? "complete in one pass (3x3 with in-kernel signed zeros / tiled rank)" : handled == 7 ? "line (van Herk)" : handled == 8 ? "runs (runtime row-run shape)" : "exact only");
		if (handled == 6) return MORSI_OK;             // exact as it stands: tiled rank, 3x3 with in-kernel signed zeros
		if (handled) {
			if (path == 2) return MORSI_OK;
			// the order-preserving re-run, gated on the flag: from shared-memory tiles where the
			// element fits one (every min / max operation), else straight from global memory
			const OpPlan plan = morsi_op_plan(job.op);
			TiledLaunch te;
			static const bool no_te = getenv("MORSI_TILED_EXACT") && !strcmp(getenv("MORSI_TILED_EXACT"), "0");
			if (!plan.special && !no_te && tiled_setup(de, plan, flag, &te)) {
				te.exact = 1;
				return run_exact_chunked(c, de, job, flag, &te);
			}
			return run_exact_chunked(c, de, job, flag);
		}
	}
	return run_exact_chunked(c, de, job, nullptr);
}
