// pipeline.cu -- the callers and data formats either side of the hot path
// (SURVEY.md 8f): morsi_all with shared erosion / dilation passes
// (src/morsi.c:278-310), pixel-interleaved ("vec") images converted and split
// on the device (src/iio.c:1416-1428 break_pixels_float / recover_broken_pixels_float
// and the sample conversions of src/iio.c:1139-1158), and a streaming entry
// point that moves an image through the device in row bands read and written
// by callbacks, so that neither the host nor the device ever holds the whole
// image (the role src/fancy_image.h:40-70 plays for the reference's tools; no
// int-sized byte counts as in src/iio.c:3759,4073).
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <vector>

#include "dispatch.cuh"

// ---------------------------------------------------------------------------
// morsi_all on the device
// ---------------------------------------------------------------------------
struct AllArgs {
	const float *x, *mn, *mx, *ope, *clo;
	float *grad, *igrad, *egrad, *lap, *enh, *str, *top, *bot;
	long long n;
};

// every pointwise output of src/morsi.c:296-303 in one pass over min / max / x / opening / closing
__global__ void __launch_bounds__(256) k_all_pointwise(AllArgs a)
{
	for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < a.n; i += (long long)gridDim.x * blockDim.x) {
		const float x = a.x[i];
		const float mn = a.mn ? a.mn[i] : 0.f, mx = a.mx ? a.mx[i] : 0.f;
		if (a.grad) a.grad[i] = __fsub_rn(mx, mn);                       // :296
		if (a.igrad) a.igrad[i] = __fsub_rn(x, mn);                      // :297
		if (a.egrad) a.egrad[i] = __fsub_rn(mx, x);                      // :298
		if (a.top) a.top[i] = __fsub_rn(x, a.ope[i]);                    // :299
		if (a.bot) a.bot[i] = __fsub_rn(a.clo[i], x);                    // :300
		if (a.str) a.str[i] = __fsub_rn(a.clo[i], a.ope[i]);             // :301
		if (a.lap || a.enh) {
			// (max + min - 2x)/2, left to right (:302); then x - lap (:303)
			const float l = __fmul_rn(__fsub_rn(__fadd_rn(mx, mn), __fmul_rn(2.0f, x)), 0.5f);
			if (a.lap) a.lap[i] = l;
			if (a.enh) a.enh[i] = __fsub_rn(x, l);
		}
	}
}

static int one_pass(MorsiCtx *c, const int *e, int op, const float *src, float *dst, int w, int h, int planes, int lane, cudaStream_t st)
{
	MorsiJob job;
	job.op = op; job.w = w; job.h = h; job.lane = lane;
	job.x_row0 = 0; job.x_rows = h; job.x_pstride = (long long)w * h;
	job.y_row0 = 0; job.y_rows = h; job.y_pstride = (long long)w * h;
	job.stream = st;
	for (int p0 = 0; p0 < planes; p0 += 32768) {
		job.planes = planes - p0 < 32768 ? planes - p0 : 32768;
		job.x = src + (long long)p0 * w * h;
		job.y = dst + (long long)p0 * w * h;
		int rc = morsi_dispatch(c, e, job);
		if (rc) return rc;
	}
	return MORSI_OK;
}

// out[k] in the order of morsi_cuda_apply_all; tmp[0..3]: device temporaries of the image's size for
// min / max / opening / closing when the caller did not ask for those outputs (may be NULL when unused)
static int all_on_device(MorsiCtx *c, const int *e, const float *d_x, float *const out[12], float *const tmp[4],
		int w, int h, int planes, int lane, cudaStream_t st)
{
	float *o_ero = out[0], *o_dil = out[1], *o_ope = out[2], *o_clo = out[3], *o_grad = out[4], *o_igrad = out[5],
		*o_egrad = out[6], *o_lap = out[7], *o_enh = out[8], *o_str = out[9], *o_top = out[10], *o_bot = out[11];
	const bool want_ope = o_ope || o_top || o_str, want_clo = o_clo || o_bot || o_str;
	const bool want_min = o_ero || want_ope || o_grad || o_igrad || o_lap || o_enh;
	const bool want_max = o_dil || want_clo || o_grad || o_egrad || o_lap || o_enh;
	float *mn = o_ero ? o_ero : tmp[0], *mx = o_dil ? o_dil : tmp[1];
	float *ope = o_ope ? o_ope : tmp[2], *clo = o_clo ? o_clo : tmp[3];
	int rc;
	// the reference's order (src/morsi.c:289-295): erosion and dilation of x once, then their
	// opposite reductions; every pass IS the reference call, so every value equals the reference's
	if (want_min && (rc = one_pass(c, e, MORSI_EROSION, d_x, mn, w, h, planes, lane, st))) return rc;
	if (want_max && (rc = one_pass(c, e, MORSI_DILATION, d_x, mx, w, h, planes, lane, st))) return rc;
	if (want_ope && (rc = one_pass(c, e, MORSI_DILATION, mn, ope, w, h, planes, lane, st))) return rc;
	if (want_clo && (rc = one_pass(c, e, MORSI_EROSION, mx, clo, w, h, planes, lane, st))) return rc;
	if (o_grad || o_igrad || o_egrad || o_lap || o_enh || o_str || o_top || o_bot) {
		AllArgs a;
		a.x = d_x; a.mn = want_min ? mn : nullptr; a.mx = want_max ? mx : nullptr;
		a.ope = want_ope ? ope : nullptr; a.clo = want_clo ? clo : nullptr;
		a.grad = o_grad; a.igrad = o_igrad; a.egrad = o_egrad; a.lap = o_lap; a.enh = o_enh;
		a.str = o_str; a.top = o_top; a.bot = o_bot;
		a.n = (long long)w * h * planes;
		k_all_pointwise<<<c->sm_count * 8, 256, 0, st>>>(a);
		morsi_count_launch(1);
		MORSI_CU(cudaGetLastError());
	}
	return MORSI_OK;
}

static int check_all_args(const int *e, const void *x, const void *out, int w, int h, int planes)
{
	if (!e || e[0] < 0) return morsi_set_error(MORSI_ERR_INVALID, "bad structuring element");
	if (!x || !out) return morsi_set_error(MORSI_ERR_INVALID, "NULL image pointer");
	if (w <= 0 || h <= 0 || planes <= 0)
		return morsi_set_error(MORSI_ERR_INVALID, "non-positive image size %dx%dx%d", w, h, planes);
	return MORSI_OK;
}

// which temporaries does a request need?
static void all_tmp_needs(float *const out[12], bool need[4])
{
	const bool want_ope = out[2] || out[10] || out[9], want_clo = out[3] || out[11] || out[9];
	const bool want_min = out[0] || want_ope || out[4] || out[5] || out[7] || out[8];
	const bool want_max = out[1] || want_clo || out[4] || out[6] || out[7] || out[8];
	need[0] = want_min && !out[0]; need[1] = want_max && !out[1];
	need[2] = want_ope && !out[2]; need[3] = want_clo && !out[3];
}

extern "C" int morsi_cuda_apply_all_device(const int *e, const float *d_x, float *const d_out[12], int w, int h, int planes, void *stream)
{
	int rc = check_all_args(e, d_x, d_out, w, h, planes);
	if (rc) return rc;
	MorsiCtx *c; rc = morsi_ctx_current(&c); if (rc) return rc;
	bool need[4];
	all_tmp_needs(d_out, need);
	float *tmp[4] = {nullptr, nullptr, nullptr, nullptr};
	const size_t bytes = (size_t)w * h * planes * sizeof(float);
	static const int slots[4] = {4, 5, 8, 9};          // lane 0: free of kernel temporaries (0-3) and pitched copies (6-7)
	for (int i = 0; i < 4; i++)
		if (need[i]) { void *p; if ((rc = morsi_ws_get(c, 0, slots[i], bytes, &p))) return rc; tmp[i] = (float *)p; }
	return all_on_device(c, e, d_x, d_out, tmp, w, h, planes, 0, stream ? (cudaStream_t)stream : c->stream);
}

// morsi_all (src/morsi.c:278-310) for host pointers: the input crosses PCIe once, erosion and
// dilation are computed once and shared by every output as the reference does, the
// pointwise outputs come from ONE pass over them, and the results stream back while
// the device is already free.  Planes are processed in groups that fit the device memory.
extern "C" int morsi_cuda_apply_all(const int *e, const float *x, float *const out[12], int w, int h, int planes)
{
	int rc = check_all_args(e, x, out, w, h, planes);
	if (rc) return rc;
	MorsiCtx *c;
	rc = morsi_ctx_current(&c);
	if (rc) return rc;
	std::lock_guard<std::mutex> host_lk(c->host_mu);
	int nout = 0;
	for (int k = 0; k < 12; k++) nout += out[k] != nullptr;
	if (!nout) return MORSI_OK;
	bool need[4];
	all_tmp_needs(out, need);
	const int nbuf = 1 + nout + need[0] + need[1] + need[2] + need[3];
	// plane groups: everything of a group is resident at once
	size_t free_b = 0, total_b = 0;
	MORSI_CU(cudaMemGetInfo(&free_b, &total_b));
	const size_t plane_bytes = (size_t)w * h * sizeof(float);
	long long pg = (long long)((free_b * 7 / 10) / ((size_t)nbuf * plane_bytes));
	if (pg < 1) return morsi_set_error(MORSI_ERR_TOO_LARGE, "morsi_all: %d buffers of one %dx%d plane do not fit the device (use morsi_cuda_apply per output)", nbuf, w, h);
	if (pg > planes) pg = planes;
	cudaStream_t s_run = c->lane_stream[1], s_copy = c->lane_stream[2];
	void *slab = nullptr;
	MORSI_CU(cudaMalloc(&slab, (size_t)nbuf * pg * plane_bytes));
	cudaEvent_t done, copied;
	cudaEventCreateWithFlags(&done, cudaEventDisableTiming);
	cudaEventCreateWithFlags(&copied, cudaEventDisableTiming);
	for (int p0 = 0; p0 < planes && !rc; p0 += (int)pg) {
		const int np = planes - p0 < pg ? planes - p0 : (int)pg;
		const size_t gbytes = (size_t)np * plane_bytes;
		float *base = (float *)slab;
		int b = 0;
		float *d_x = base + (size_t)(b++) * pg * w * h;
		float *d_out[12], *tmp[4];
		for (int k = 0; k < 12; k++) d_out[k] = out[k] ? base + (size_t)(b++) * pg * w * h : nullptr;
		for (int i = 0; i < 4; i++) tmp[i] = need[i] ? base + (size_t)(b++) * pg * w * h : nullptr;
		if (p0 > 0) MORSI_CU(cudaStreamWaitEvent(s_run, copied, 0));        // the previous group has left
		MORSI_CU(cudaMemcpyAsync(d_x, x + (size_t)p0 * w * h, gbytes, cudaMemcpyHostToDevice, s_run));
		rc = all_on_device(c, e, d_x, d_out, tmp, w, h, np, 1, s_run);
		if (rc) break;
		MORSI_CU(cudaEventRecord(done, s_run));
		MORSI_CU(cudaStreamWaitEvent(s_copy, done, 0));
		for (int k = 0; k < 12; k++)
			if (out[k]) MORSI_CU(cudaMemcpyAsync(out[k] + (size_t)p0 * w * h, d_out[k], gbytes, cudaMemcpyDeviceToHost, s_copy));
		MORSI_CU(cudaEventRecord(copied, s_copy));
	}
	cudaStreamSynchronize(s_run);
	cudaStreamSynchronize(s_copy);
	cudaEventDestroy(done); cudaEventDestroy(copied);
	cudaFree(slab);
	return rc;
}

// ---------------------------------------------------------------------------
// pixel-interleaved images: convert + split on the way in, join on the way out
// ---------------------------------------------------------------------------
// in:  rows x w pixels of pd interleaved samples of `type` -> pd planar float bands (plane stride ps)
template <typename T>
__global__ void __launch_bounds__(256) k_split(const T *in, float *out, long long npix, int pd, long long ps)
{
	for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < npix * pd; i += (long long)gridDim.x * blockDim.x) {
		const long long pix = i / pd;
		const int ch = (int)(i - pix * pd);
		out[ch * ps + pix] = (float)in[i];          // coalesced read; the pd-strided writes merge in L2
	}
}
__global__ void __launch_bounds__(256) k_join(const float *in, float *out, long long npix, int pd, long long ps)
{
	for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < npix * pd; i += (long long)gridDim.x * blockDim.x) {
		const long long pix = i / pd;
		const int ch = (int)(i - pix * pd);
		out[i] = in[ch * ps + pix];
	}
}

static size_t sample_bytes(int type) { return type == MORSI_SAMPLE_U8 ? 1 : type == MORSI_SAMPLE_U16 ? 2 : 4; }

// x: h rows of w pixels of pd interleaved samples (u8 / u16 / f32); y: the same layout, float32.
// Row bands go host -> device as they are (a u8 image moves a quarter of the bytes), are
// converted and split into planes on the device, processed with all planes in one launch,
// joined again and copied back: the CPU passes of src/iio.c:1416-1428 disappear.
extern "C" int morsi_cuda_apply_interleaved(int op, const int *e, const void *x, float *y, int w, int h, int pd, int sample_type)
{
	if (op < 0 || op >= MORSI_OP_COUNT) return morsi_set_error(MORSI_ERR_INVALID, "unknown operation %d", op);
	if (!e || e[0] < 0) return morsi_set_error(MORSI_ERR_INVALID, "bad structuring element");
	if (!x || !y) return morsi_set_error(MORSI_ERR_INVALID, "NULL image pointer");
	if (w <= 0 || h <= 0 || pd <= 0) return morsi_set_error(MORSI_ERR_INVALID, "non-positive image size %dx%dx%d", w, h, pd);
	if (sample_type < MORSI_SAMPLE_U8 || sample_type > MORSI_SAMPLE_F32)
		return morsi_set_error(MORSI_ERR_INVALID, "unknown sample type %d", sample_type);
	MorsiCtx *c;
	int rc = morsi_ctx_current(&c);
	if (rc) return rc;
	std::lock_guard<std::mutex> host_lk(c->host_mu);
	int up = 0, down = 0;
	if ((rc = morsi_cuda_halo_rows(op, e, &up, &down))) return rc;
	const size_t sb = sample_bytes(sample_type);
	long long target = (32LL << 20) / ((long long)w * pd * 4);
	int band = (int)std::max<long long>(std::max(64, 8 * (up + down)), target);
	if (const char *s = getenv("MORSI_CUDA_CHUNK_ROWS")) band = std::max(1, atoi(s));
	const int nchunks = (h + band - 1) / band;
	const int lanes = std::min(3, nchunks);
	const size_t max_in_rows = (size_t)std::min(h, band + up + down), max_out_rows = (size_t)std::min(h, band);
	void *d_raw[3], *d_pl[3], *d_res[3], *d_vec[3];
	for (int l = 0; l < lanes; l++) {
		if ((rc = morsi_ws_get(c, 1 + l, 4, max_in_rows * w * pd * sb, &d_raw[l]))) return rc;
		if ((rc = morsi_ws_get(c, 1 + l, 5, max_in_rows * w * pd * 4, &d_pl[l]))) return rc;
		if ((rc = morsi_ws_get(c, 1 + l, 8, max_out_rows * w * pd * 4, &d_res[l]))) return rc;
		if ((rc = morsi_ws_get(c, 1 + l, 9, max_out_rows * w * pd * 4, &d_vec[l]))) return rc;
	}
	for (int t = 0; t < nchunks; t++) {
		const int l = t % lanes;
		cudaStream_t s = c->lane_stream[1 + l];
		const int r0 = t * band, r1 = std::min(h, r0 + band);
		const int i0 = std::max(0, r0 - up), i1 = std::min(h, r1 + down);
		const long long npix_in = (long long)(i1 - i0) * w, npix_out = (long long)(r1 - r0) * w;
		MORSI_CU(cudaMemcpyAsync(d_raw[l], (const char *)x + (size_t)i0 * w * pd * sb, (size_t)npix_in * pd * sb, cudaMemcpyHostToDevice, s));
		const int grid = c->sm_count * 8;
		if (sample_type == MORSI_SAMPLE_U8) k_split<unsigned char><<<grid, 256, 0, s>>>((const unsigned char *)d_raw[l], (float *)d_pl[l], npix_in, pd, npix_in);
		else if (sample_type == MORSI_SAMPLE_U16) k_split<unsigned short><<<grid, 256, 0, s>>>((const unsigned short *)d_raw[l], (float *)d_pl[l], npix_in, pd, npix_in);
		else k_split<float><<<grid, 256, 0, s>>>((const float *)d_raw[l], (float *)d_pl[l], npix_in, pd, npix_in);
		morsi_count_launch(1);
		MorsiJob job;
		job.op = op; job.w = w; job.h = h; job.planes = pd; job.lane = 1 + l;
		job.x = (const float *)d_pl[l]; job.x_row0 = i0; job.x_rows = i1 - i0; job.x_pstride = npix_in;
		job.y = (float *)d_res[l]; job.y_row0 = r0; job.y_rows = r1 - r0; job.y_pstride = npix_out;
		job.stream = s;
		if ((rc = morsi_dispatch(c, e, job))) return rc;
		k_join<<<grid, 256, 0, s>>>((const float *)d_res[l], (float *)d_vec[l], npix_out, pd, npix_out);
		morsi_count_launch(1);
		MORSI_CU(cudaGetLastError());
		MORSI_CU(cudaMemcpyAsync(y + (size_t)r0 * w * pd, d_vec[l], (size_t)npix_out * pd * 4, cudaMemcpyDeviceToHost, s));
	}
	for (int l = 0; l < lanes; l++) MORSI_CU(cudaStreamSynchronize(c->lane_stream[1 + l]));
	return MORSI_OK;
}

// ---------------------------------------------------------------------------
// streaming: the image never exists as a whole, neither on the host nor on the device
// ---------------------------------------------------------------------------
struct StreamLane {
	float *h_in = nullptr, *h_out = nullptr;       // pinned staging
	cudaEvent_t done;
	int plane = -1, r0 = 0, r1 = 0;                // the chunk whose result sits in h_out (plane < 0: none)
};

extern "C" int morsi_cuda_apply_stream(int op, const int *e, int w, int h, int planes,
		morsi_read_rows_fn rd, morsi_write_rows_fn wr, void *user)
{
	if (op < 0 || op >= MORSI_OP_COUNT) return morsi_set_error(MORSI_ERR_INVALID, "unknown operation %d", op);
	if (!e || e[0] < 0) return morsi_set_error(MORSI_ERR_INVALID, "bad structuring element");
	if (!rd || !wr) return morsi_set_error(MORSI_ERR_INVALID, "NULL callback");
	if (w <= 0 || h <= 0 || planes <= 0) return morsi_set_error(MORSI_ERR_INVALID, "non-positive image size %dx%dx%d", w, h, planes);
	MorsiCtx *c;
	int rc = morsi_ctx_current(&c);
	if (rc) return rc;
	std::lock_guard<std::mutex> host_lk(c->host_mu);
	int up = 0, down = 0;
	if ((rc = morsi_cuda_halo_rows(op, e, &up, &down))) return rc;
	long long target = (32LL << 20) / ((long long)w * 4);
	int band = (int)std::max<long long>(std::max(64, 8 * (up + down)), target);
	if (const char *s = getenv("MORSI_CUDA_CHUNK_ROWS")) band = std::max(1, atoi(s));
	band = std::min(band, h);
	const size_t in_rows = (size_t)std::min(h, band + up + down);
	const long long nchunks = (long long)planes * ((h + band - 1) / band);
	const int lanes = (int)std::min<long long>(3, nchunks);
	StreamLane L[3];
	void *d_in[3], *d_out[3];
	auto cleanup = [&]() {
		for (int l = 0; l < lanes; l++) {
			if (L[l].h_in) cudaFreeHost(L[l].h_in);
			if (L[l].h_out) cudaFreeHost(L[l].h_out);
			cudaEventDestroy(L[l].done);
		}
	};
	for (int l = 0; l < lanes; l++) {
		cudaEventCreateWithFlags(&L[l].done, cudaEventDisableTiming);
		if (cudaMallocHost(&L[l].h_in, in_rows * w * 4) != cudaSuccess || cudaMallocHost(&L[l].h_out, (size_t)band * w * 4) != cudaSuccess) {
			cleanup();
			return morsi_set_error(MORSI_ERR_OOM, "stream: pinned staging of %zu rows", in_rows);
		}
		if ((rc = morsi_ws_get(c, 1 + l, 4, in_rows * w * 4, &d_in[l])) || (rc = morsi_ws_get(c, 1 + l, 5, (size_t)band * w * 4, &d_out[l]))) { cleanup(); return rc; }
	}
	// hand a lane's finished chunk to the writer
	auto flush = [&](int l) -> int {
		if (L[l].plane < 0) return 0;
		if (cudaEventSynchronize(L[l].done) != cudaSuccess) return morsi_set_error(MORSI_ERR_CUDA, "stream: %s", cudaGetErrorString(cudaGetLastError()));
		const int r = wr(user, L[l].plane, L[l].r0, L[l].r1 - L[l].r0, L[l].h_out);
		L[l].plane = -1;
		return r ? morsi_set_error(MORSI_ERR_INVALID, "stream: the write callback failed (%d)", r) : 0;
	};
	long long t = 0;
	for (int p = 0; p < planes && !rc; p++)
		for (int r0 = 0; r0 < h && !rc; r0 += band, t++) {
			const int l = (int)(t % lanes);
			cudaStream_t s = c->lane_stream[1 + l];
			const int r1 = std::min(h, r0 + band);
			const int i0 = std::max(0, r0 - up), i1 = std::min(h, r1 + down);
			if ((rc = flush(l))) break;                                    // its staging buffers are free again
			const int r = rd(user, p, i0, i1 - i0, L[l].h_in);
			if (r) { rc = morsi_set_error(MORSI_ERR_INVALID, "stream: the read callback failed (%d)", r); break; }
			if (cudaMemcpyAsync(d_in[l], L[l].h_in, (size_t)(i1 - i0) * w * 4, cudaMemcpyHostToDevice, s) != cudaSuccess) { rc = morsi_set_error(MORSI_ERR_CUDA, "stream: upload"); break; }
			MorsiJob job;
			job.op = op; job.w = w; job.h = h; job.planes = 1; job.lane = 1 + l;
			job.x = (const float *)d_in[l]; job.x_row0 = i0; job.x_rows = i1 - i0; job.x_pstride = (long long)w * (i1 - i0);
			job.y = (float *)d_out[l]; job.y_row0 = r0; job.y_rows = r1 - r0; job.y_pstride = (long long)w * (r1 - r0);
			job.stream = s;
			if ((rc = morsi_dispatch(c, e, job))) break;
			if (cudaMemcpyAsync(L[l].h_out, d_out[l], (size_t)(r1 - r0) * w * 4, cudaMemcpyDeviceToHost, s) != cudaSuccess) { rc = morsi_set_error(MORSI_ERR_CUDA, "stream: download"); break; }
			cudaEventRecord(L[l].done, s);
			L[l].plane = p; L[l].r0 = r0; L[l].r1 = r1;
		}
	// drain in submission order
	for (int k = 0; k < lanes && !rc; k++) rc = flush((int)((t + k) % lanes));
	for (int l = 0; l < lanes; l++) cudaStreamSynchronize(c->lane_stream[1 + l]);
	cleanup();
	return rc;
}

// ---------------------------------------------------------------------------
// pipe chains: `morsi E1 OP1 in | morsi E2 OP2 | qeasy black white - out.png`
// ---------------------------------------------------------------------------
// doc/tutorial/i.html:221-225 pipes morsi into qeasy (src/qeasy.c:55-73), scripts pipe
// morsi into morsi: every `|` is a full serialisation of the image and, here, a PCIe round
// trip.  The chain runs on the device instead: one upload, the operations ping-pong between
// two resident buffers, the quantiser of qeasy is the last kernel, and an 8-bit result
// crosses PCIe as bytes.
//   x[i] = floor(255 * (x[i] - black) / (white - black))       src/qeasy.c:57-58 (float arithmetic, left to right)
//   u8:  < 0 -> 0, > 255 -> 255, else the integer value         :64-69 (NaN -> 0, what x86's cvttss2si leaves in the low byte)
__global__ void __launch_bounds__(256) k_qeasy(const float *in, float *out_f, unsigned char *out_u8, long long n, float black, float white)
{
	const float range = __fsub_rn(white, black);
	for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
		const float v = floorf(__fdiv_rn(__fmul_rn(255.0f, __fsub_rn(in[i], black)), range));
		if (out_f) out_f[i] = v;
		if (out_u8) out_u8[i] = v < 0.f ? 0 : v > 255.f ? 255 : (v == v ? (unsigned char)(int)v : 0);
	}
}

extern "C" int morsi_cuda_apply_chain(int nops, const int *ops, const int *const *elements, const float *x, void *y,
		int w, int h, int planes, const morsi_quantizer *q)
{
	if (nops < 1 || !ops || !elements) return morsi_set_error(MORSI_ERR_INVALID, "chain: no operations");
	for (int k = 0; k < nops; k++) {
		if (ops[k] < 0 || ops[k] >= MORSI_OP_COUNT) return morsi_set_error(MORSI_ERR_INVALID, "chain: unknown operation %d", ops[k]);
		if (!elements[k] || elements[k][0] < 0) return morsi_set_error(MORSI_ERR_INVALID, "chain: bad structuring element %d", k);
	}
	if (!x || !y) return morsi_set_error(MORSI_ERR_INVALID, "NULL image pointer");
	if (w <= 0 || h <= 0 || planes <= 0) return morsi_set_error(MORSI_ERR_INVALID, "non-positive image size %dx%dx%d", w, h, planes);
	MorsiCtx *c;
	int rc = morsi_ctx_current(&c);
	if (rc) return rc;
	std::lock_guard<std::mutex> host_lk(c->host_mu);
	const size_t plane_bytes = (size_t)w * h * sizeof(float);
	size_t free_b = 0, total_b = 0;
	MORSI_CU(cudaMemGetInfo(&free_b, &total_b));
	long long pg = (long long)((free_b * 7 / 10) / (3 * plane_bytes));       // two ping-pong buffers + the quantised result
	if (pg < 1) return morsi_set_error(MORSI_ERR_TOO_LARGE, "chain: three buffers of one %dx%d plane do not fit the device", w, h);
	if (pg > planes) pg = planes;
	const bool to_u8 = q && q->to_uint8;
	cudaStream_t s_run = c->lane_stream[1], s_copy = c->lane_stream[2];
	void *slab = nullptr;
	MORSI_CU(cudaMalloc(&slab, 3 * (size_t)pg * plane_bytes));
	float *buf[2] = {(float *)slab, (float *)slab + (size_t)pg * w * h};
	void *d_q = (float *)slab + 2 * (size_t)pg * w * h;
	cudaEvent_t done, copied;
	cudaEventCreateWithFlags(&done, cudaEventDisableTiming);
	cudaEventCreateWithFlags(&copied, cudaEventDisableTiming);
	for (int p0 = 0; p0 < planes && !rc; p0 += (int)pg) {
		const int np = planes - p0 < pg ? planes - p0 : (int)pg;
		const long long n = (long long)np * w * h;
		if (p0 > 0) cudaStreamWaitEvent(s_run, copied, 0);
		if (cudaMemcpyAsync(buf[0], x + (size_t)p0 * w * h, (size_t)n * 4, cudaMemcpyHostToDevice, s_run) != cudaSuccess) { rc = morsi_set_error(MORSI_ERR_CUDA, "chain: upload"); break; }
		int cur = 0;
		for (int k = 0; k < nops && !rc; k++, cur ^= 1)
			rc = one_pass(c, elements[k], ops[k], buf[cur], buf[cur ^ 1], w, h, np, 1, s_run);
		if (rc) break;
		const void *src = buf[cur];
		size_t out_bytes = (size_t)n * 4;
		if (q) {
			k_qeasy<<<c->sm_count * 8, 256, 0, s_run>>>(buf[cur], to_u8 ? nullptr : (float *)d_q, to_u8 ? (unsigned char *)d_q : nullptr, n, q->black, q->white);
			morsi_count_launch(1);
			src = d_q;
			if (to_u8) out_bytes = (size_t)n;
		}
		cudaEventRecord(done, s_run);
		cudaStreamWaitEvent(s_copy, done, 0);
		char *dst = (char *)y + (size_t)p0 * w * h * (to_u8 ? 1 : 4);
		if (cudaMemcpyAsync(dst, src, out_bytes, cudaMemcpyDeviceToHost, s_copy) != cudaSuccess) { rc = morsi_set_error(MORSI_ERR_CUDA, "chain: download"); break; }
		cudaEventRecord(copied, s_copy);
		// the next group's upload overwrites buf[0]: the download of this group must have read its source first
		if (src == buf[0]) cudaStreamWaitEvent(s_run, copied, 0);
	}
	cudaStreamSynchronize(s_run);
	cudaStreamSynchronize(s_copy);
	cudaEventDestroy(done); cudaEventDestroy(copied);
	cudaFree(slab);
	if (!rc && cudaGetLastError() != cudaSuccess) rc = morsi_set_error(MORSI_ERR_CUDA, "chain: a CUDA call failed");
	return rc;
}
