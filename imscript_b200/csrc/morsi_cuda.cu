// morsi_cuda.cu -- C ABI, device context and operation dispatcher of
// libmorsi_cuda (see include/morsi_cuda.h).  Replaces, for the GPU, the
// operation table and channel loop of src/morsi.c:509-543 and the composite
// wrappers of src/morsi.c:141-275.
#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/morsi_cuda.h"
#include "element.h"
#include "dispatch.cuh"

// ---------------------------------------------------------------------------
// errors
// ---------------------------------------------------------------------------
static thread_local char g_err[512] = "";

int morsi_set_error(int code, const char *fmt, ...)
{
	va_list ap;
	va_start(ap, fmt);
	vsnprintf(g_err, sizeof g_err, fmt, ap);
	va_end(ap);
	if (getenv("MORSI_CUDA_TRACE")) fprintf(stderr, "morsi_cuda: error %d: %s\n", code, g_err);
	return code;
}

#define CU(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) \
	return morsi_set_error(e_ == cudaErrorMemoryAllocation ? MORSI_ERR_OOM : \
		(e_ == cudaErrorNoDevice || e_ == cudaErrorInsufficientDriver) ? MORSI_ERR_NO_DEVICE : MORSI_ERR_CUDA, \
		"%s: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); } while (0)

extern "C" const char *morsi_cuda_strerror(int s)
{
	switch (s) {
	case MORSI_OK: return "ok";
	case MORSI_ERR_INVALID: return "invalid argument";
	case MORSI_ERR_NO_DEVICE: return "no CUDA device (libmorsi_cuda has no CPU fallback)";
	case MORSI_ERR_CUDA: return "CUDA error";
	case MORSI_ERR_OOM: return "out of memory";
	case MORSI_ERR_TOO_LARGE: return "image too large";
	case MORSI_ERR_COMM: return "multi-device exchange failed";
	}
	return "unknown error";
}
extern "C" const char *morsi_cuda_last_error(void) { return g_err; }

// ---------------------------------------------------------------------------
// context
// ---------------------------------------------------------------------------
static std::atomic<long> g_launches{0};
void morsi_count_launch(int n) { g_launches += n; }
extern "C" long morsi_cuda_launch_count(void) { return g_launches.load(); }
extern "C" void morsi_cuda_launch_count_reset(void) { g_launches = 0; }

static std::mutex g_mu;
static std::map<int, MorsiCtx *> g_ctx;     // one per device
static int g_current = -1;
static int g_path = -1;

extern "C" int morsi_cuda_set_path(int path)
{
	if (path < 0 || path > 2) return MORSI_ERR_INVALID;
	g_path = path;
	return MORSI_OK;
}
int morsi_path(void)
{
	if (g_path < 0) {
		const char *s = getenv("MORSI_CUDA_PATH");
		g_path = 0;
		if (s && !strcmp(s, "exact")) g_path = 1;
		if (s && !strcmp(s, "fast")) g_path = 2;
	}
	return g_path;
}

extern "C" int morsi_cuda_device_count(void)
{
	int n = 0;
	if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
	return n;
}

static int ctx_create(int device, MorsiCtx **out)
{
	int n = morsi_cuda_device_count();
	if (n <= 0) return morsi_set_error(MORSI_ERR_NO_DEVICE, "no CUDA device visible");
	if (device < 0 || device >= n)
		return morsi_set_error(MORSI_ERR_INVALID, "device %d out of range (%d visible)", device, n);
	CU(cudaSetDevice(device));
	MorsiCtx *c = new MorsiCtx();
	c->device = device;
	cudaDeviceProp prop;
	CU(cudaGetDeviceProperties(&prop, device));
	c->sm_count = prop.multiProcessorCount;
	c->smem_optin = (int)prop.sharedMemPerBlockOptin;
	if (prop.major < 10 && !getenv("MORSI_CUDA_ANY_ARCH")) {
		delete c;
		return morsi_set_error(MORSI_ERR_NO_DEVICE,
			"device %d is sm_%d%d; libmorsi_cuda is built for sm_100a only", device, prop.major, prop.minor);
	}
	CU(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
	for (int l = 0; l < MORSI_LANES; l++)
		CU(cudaStreamCreateWithFlags(&c->lane_stream[l], cudaStreamNonBlocking));
	{
		// keep stream-ordered workspace memory cached in the pool between calls
		cudaMemPool_t pool;
		if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
			unsigned long long keep = ~0ull;
			cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
		}
		cudaGetLastError();
	}
	CU(cudaMalloc(&c->d_flag, 64));
	CU(cudaMemset(c->d_flag, 0, 64));
	CU(cudaMallocHost(&c->h_flag, 64));
	*out = c;
	return MORSI_OK;
}

int morsi_ctx_get(int device, MorsiCtx **out)
{
	std::lock_guard<std::mutex> lk(g_mu);
	auto it = g_ctx.find(device);
	if (it != g_ctx.end()) {
		*out = it->second;
		cudaSetDevice(device);
		return MORSI_OK;
	}
	MorsiCtx *c = nullptr;
	int rc = ctx_create(device, &c);
	if (rc) return rc;
	g_ctx[device] = c;
	*out = c;
	return MORSI_OK;
}

extern "C" int morsi_cuda_init(int device)
{
	MorsiCtx *c;
	int rc = morsi_ctx_get(device, &c);
	if (rc) return rc;
	g_current = device;
	return MORSI_OK;
}

int morsi_ctx_current(MorsiCtx **out)
{
	if (g_current < 0) {
		int rc = morsi_cuda_init(0);
		if (rc) return rc;
	}
	return morsi_ctx_get(g_current, out);
}

extern "C" void morsi_cuda_shutdown(void)
{
	std::lock_guard<std::mutex> lk(g_mu);
	for (auto &kv : g_ctx) {
		MorsiCtx *c = kv.second;
		cudaSetDevice(c->device);
		cudaStreamSynchronize(c->stream);
		for (auto &e : c->elements) { cudaFree(e.second.d_offs); cudaFree(e.second.d_tile_offs); }
		for (int l = 0; l < MORSI_LANES; l++)
			for (int i = 0; i < MORSI_WS_SLOTS; i++)
				if (c->ws[l][i]) { if (l >= 1 && l <= 3) cudaFreeAsync(c->ws[l][i], c->lane_stream[l]); else cudaFree(c->ws[l][i]); }
		for (int l = 0; l < MORSI_LANES; l++) if (c->lane_stream[l]) cudaStreamSynchronize(c->lane_stream[l]);
		for (int l = 0; l < MORSI_LANES; l++) if (c->lane_stream[l]) cudaStreamDestroy(c->lane_stream[l]);
		cudaFree(c->d_flag);
		cudaFreeHost(c->h_flag);
		cudaStreamDestroy(c->stream);
		delete c;
	}
	g_ctx.clear();
	g_current = -1;
}

// Workspace slots grow on demand.  Lanes 1..3 belong to the library's own
// streams: their slots are stream-ordered allocations (cudaMallocAsync /
// cudaFreeAsync on the lane's stream), so growing one never stalls the other
// lanes.  Lane 0 serves caller-supplied streams: a growth there waits for the
// device (rare: sizes only ever grow).
int morsi_ws_get(MorsiCtx *c, int lane, int slot, size_t bytes, void **out)
{
	std::lock_guard<std::mutex> lk(c->mu);
	if (c->ws_bytes[lane][slot] < bytes) {
		const bool async = lane >= 1 && lane <= 3;
		if (c->ws[lane][slot]) {
			if (async) CU(cudaFreeAsync(c->ws[lane][slot], c->lane_stream[lane]));
			else { CU(cudaDeviceSynchronize()); CU(cudaFree(c->ws[lane][slot])); }
			c->ws[lane][slot] = nullptr; c->ws_bytes[lane][slot] = 0;
		}
		if (async) CU(cudaMallocAsync(&c->ws[lane][slot], bytes, c->lane_stream[lane]));
		else CU(cudaMalloc(&c->ws[lane][slot], bytes));
		c->ws_bytes[lane][slot] = bytes;
	}
	*out = c->ws[lane][slot];
	return MORSI_OK;
}

// Element lists are uploaded once and cached by content.
int morsi_element_get(MorsiCtx *c, const int *e, const DevElement **out)
{
	if (!e || e[0] < 0) return morsi_set_error(MORSI_ERR_INVALID, "bad structuring element");
	std::lock_guard<std::mutex> lk(c->mu);
	std::string key((const char *)e, (size_t)(4 + 2 * (size_t)e[0]) * sizeof(int));
	auto it = c->elements.find(key);
	if (it != c->elements.end()) { *out = &it->second; return MORSI_OK; }
	DevElement d;
	if (morsi_element_analyze(e, &d.info))
		return morsi_set_error(MORSI_ERR_INVALID, "bad structuring element");
	d.n = e[0];
	d.d_offs = nullptr;
	std::vector<int2> offs((size_t)d.n + 1);
	for (int k = 0; k < d.n; k++)
		offs[k] = make_int2(e[4 + 2*k] - e[2], e[5 + 2*k] - e[3]);
	CU(cudaMalloc(&d.d_offs, ((size_t)d.n + 1) * sizeof(int2)));
	CU(cudaMemcpy(d.d_offs, offs.data(), ((size_t)d.n + 1) * sizeof(int2), cudaMemcpyHostToDevice));
	// the reference's own 3x3 literals, in their order (src/morsi.c:484-485; disk2 builds the square's list)
	{
		static const int cross[5][2] = {{-1, 0}, {0, 0}, {1, 0}, {0, -1}, {0, 1}};
		d.canonical3x3 = 0;
		if (d.n == 5) {
			bool same = true;
			for (int k = 0; k < 5; k++) same &= offs[k].x == cross[k][0] && offs[k].y == cross[k][1];
			if (same) d.canonical3x3 = 1;
		} else if (d.n == 9) {
			bool same = true;
			for (int k = 0; k < 9; k++) same &= offs[k].x == k / 3 - 1 && offs[k].y == k % 3 - 1;
			if (same) d.canonical3x3 = 2;
		}
	}
	// k_tiled: offsets into a shared-memory tile of pitch 128 + (xmax - xmin)
	d.d_tile_offs = nullptr;
	if (d.n > 0 && d.n <= 8192) {
		const int pw = 128 + d.info.xmax - d.info.xmin;
		std::vector<int> toffs((size_t)d.n);
		for (int k = 0; k < d.n; k++)
			toffs[k] = (offs[k].y - d.info.ymin) * pw + (offs[k].x - d.info.xmin);
		CU(cudaMalloc(&d.d_tile_offs, (size_t)d.n * sizeof(int)));
		CU(cudaMemcpy(d.d_tile_offs, toffs.data(), (size_t)d.n * sizeof(int), cudaMemcpyHostToDevice));
	}
	morsi_element_compile(c, &d);
	auto ins = c->elements.emplace(key, d);
	*out = &ins.first->second;
	return MORSI_OK;
}

// ---------------------------------------------------------------------------
// plumbing entry points
// ---------------------------------------------------------------------------
static cudaStream_t pick_stream(MorsiCtx *c, void *stream) { return stream ? (cudaStream_t)stream : c->stream; }

extern "C" int morsi_cuda_malloc(void **p, size_t bytes)
{
	MorsiCtx *c; int rc = morsi_ctx_current(&c); if (rc) return rc;
	CU(cudaMalloc(p, bytes ? bytes : 1));
	return MORSI_OK;
}
extern "C" int morsi_cuda_free(void *p)
{
	MorsiCtx *c; int rc = morsi_ctx_current(&c); if (rc) return rc;
	CU(cudaFree(p));
	return MORSI_OK;
}
extern "C" int morsi_cuda_host_alloc(void **p, size_t bytes)
{
	MorsiCtx *c; int rc = morsi_ctx_current(&c); if (rc) return rc;
	CU(cudaMallocHost(p, bytes ? bytes : 1));
	return MORSI_OK;
}
extern "C" int morsi_cuda_host_free(void *p)
{
	CU(cudaFreeHost(p));
	return MORSI_OK;
}
extern "C" int morsi_cuda_memcpy_h2d(void *d, const void *h, size_t bytes, void *stream)
{
	MorsiCtx *c; int rc = morsi_ctx_current(&c); if (rc) return rc;
	CU(cudaMemcpyAsync(d, h, bytes, cudaMemcpyHostToDevice, pick_stream(c, stream)));
	return MORSI_OK;
}
extern "C" int morsi_cuda_memcpy_d2h(void *h, const void *d, size_t bytes, void *stream)
{
	MorsiCtx *c; int rc = morsi_ctx_current(&c); if (rc) return rc;
	CU(cudaMemcpyAsync(h, d, bytes, cudaMemcpyDeviceToHost, pick_stream(c, stream)));
	return MORSI_OK;
}
extern "C" int morsi_cuda_memcpy_d2d(void *dst, const void *src, size_t bytes, void *stream)
{
	MorsiCtx *c; int rc = morsi_ctx_current(&c); if (rc) return rc;
	CU(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, pick_stream(c, stream)));
	return MORSI_OK;
}
extern "C" int morsi_cuda_sync(void *stream)
{
	MorsiCtx *c; int rc = morsi_ctx_current(&c); if (rc) return rc;
	CU(cudaStreamSynchronize(pick_stream(c, stream)));
	return MORSI_OK;
}
extern "C" int morsi_cuda_stream_create(void **stream)
{
	MorsiCtx *c; int rc = morsi_ctx_current(&c); if (rc) return rc;
	cudaStream_t s;
	CU(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
	*stream = (void *)s;
	return MORSI_OK;
}
extern "C" int morsi_cuda_stream_destroy(void *stream)
{
	CU(cudaStreamDestroy((cudaStream_t)stream));
	return MORSI_OK;
}
extern "C" int morsi_cuda_event_create(void **ev)
{
	MorsiCtx *c; int rc = morsi_ctx_current(&c); if (rc) return rc;
	cudaEvent_t e;
	CU(cudaEventCreate(&e));
	*ev = (void *)e;
	return MORSI_OK;
}
extern "C" int morsi_cuda_event_record(void *ev, void *stream)
{
	MorsiCtx *c; int rc = morsi_ctx_current(&c); if (rc) return rc;
	CU(cudaEventRecord((cudaEvent_t)ev, pick_stream(c, stream)));
	return MORSI_OK;
}
extern "C" int morsi_cuda_event_elapsed_ms(void *a, void *b, float *ms)
{
	CU(cudaEventSynchronize((cudaEvent_t)b));
	CU(cudaEventElapsedTime(ms, (cudaEvent_t)a, (cudaEvent_t)b));
	return MORSI_OK;
}
extern "C" int morsi_cuda_event_destroy(void *ev)
{
	CU(cudaEventDestroy((cudaEvent_t)ev));
	return MORSI_OK;
}

// ---- synthetic images -------------------------------------------------------
__global__ void k_synth(float *x, int w, int rows, int row0, int plane, unsigned seed, int dist)
{
	long long total = (long long)w * rows;
	for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total;
			t += (long long)gridDim.x * blockDim.x) {
		int r = (int)(t / w), col = (int)(t - (long long)r * w);
		x[t] = morsi_synth_value(seed, plane, row0 + r, col, dist);
	}
}
extern "C" int morsi_cuda_synth(float *d_x, int w, int rows, int row0, int plane,
		unsigned seed, int dist, void *stream)
{
	MorsiCtx *c; int rc = morsi_ctx_current(&c); if (rc) return rc;
	if (!d_x || w <= 0 || rows <= 0) return morsi_set_error(MORSI_ERR_INVALID, "synth: bad size");
	k_synth<<<c->sm_count * 8, 256, 0, pick_stream(c, stream)>>>(d_x, w, rows, row0, plane, seed, dist);
	morsi_count_launch(1);
	CU(cudaGetLastError());
	return MORSI_OK;
}
extern "C" void morsi_synth_host(float *x, int w, int rows, int row0, int plane, unsigned seed, int dist)
{
	for (int r = 0; r < rows; r++)
		for (int col = 0; col < w; col++)
			x[(size_t)r * w + col] = morsi_synth_value(seed, plane, row0 + r, col, dist);
}

// ---------------------------------------------------------------------------
// the hot path
// ---------------------------------------------------------------------------
extern "C" int morsi_cuda_halo_rows(int op, const int *e, int *up, int *down)
{
	morsi_element_info info;
	if (op < 0 || op >= MORSI_OP_COUNT || morsi_element_analyze(e, &info))
		return morsi_set_error(MORSI_ERR_INVALID, "halo_rows: bad op or element");
	int stages = morsi_op_plan(op).stages;
	int u = info.n ? (info.ymin < 0 ? -info.ymin : 0) : 0;
	int d = info.n ? (info.ymax > 0 ? info.ymax : 0) : 0;
	if (up) *up = u * stages;
	if (down) *down = d * stages;
	return MORSI_OK;
}

static int check_common(int op, const int *e, const void *x, const void *y, int w, int h)
{
	if (op < 0 || op >= MORSI_OP_COUNT) return morsi_set_error(MORSI_ERR_INVALID, "unknown operation %d", op);
	if (!e || e[0] < 0) return morsi_set_error(MORSI_ERR_INVALID, "bad structuring element");
	if (!x || !y) return morsi_set_error(MORSI_ERR_INVALID, "NULL image pointer");
	if (w <= 0 || h <= 0) return morsi_set_error(MORSI_ERR_INVALID, "non-positive image size %dx%d", w, h);
	return MORSI_OK;
}

extern "C" int morsi_cuda_apply_band_device(int op, const int *e,
		const float *d_x, int x_row0, int x_rows,
		float *d_y, int y_row0, int y_rows, int w, int h, void *stream)
{
	int rc = check_common(op, e, d_x, d_y, w, h);
	if (rc) return rc;
	MorsiCtx *c; rc = morsi_ctx_current(&c); if (rc) return rc;
	if (y_rows <= 0 || y_row0 < 0 || y_row0 + (long long)y_rows > h || x_rows <= 0)
		return morsi_set_error(MORSI_ERR_INVALID, "bad band [%d,+%d) of %d rows", y_row0, y_rows, h);
	int up, down;
	morsi_cuda_halo_rows(op, e, &up, &down);
	int need0 = y_row0 - up < 0 ? 0 : y_row0 - up;
	long long need1 = (long long)y_row0 + y_rows + down; if (need1 > h) need1 = h;
	if (x_row0 > need0 || x_row0 + (long long)x_rows < need1)
		return morsi_set_error(MORSI_ERR_INVALID,
			"input band [%d,+%d) does not cover rows [%d,%lld) needed by output band [%d,+%d)",
			x_row0, x_rows, need0, need1, y_row0, y_rows);
	MorsiJob job;
	job.op = op; job.w = w; job.h = h; job.planes = 1; job.lane = 0;
	job.x = d_x; job.x_row0 = x_row0; job.x_rows = x_rows; job.x_pstride = (long long)w * x_rows;
	job.y = d_y; job.y_row0 = y_row0; job.y_rows = y_rows; job.y_pstride = (long long)w * y_rows;
	job.stream = pick_stream(c, stream);
	return morsi_dispatch(c, e, job);
}

extern "C" int morsi_cuda_apply_device(int op, const int *e, const float *d_x, float *d_y,
		int w, int h, int planes, void *stream)
{
	int rc = check_common(op, e, d_x, d_y, w, h);
	if (rc) return rc;
	if (planes <= 0) return morsi_set_error(MORSI_ERR_INVALID, "non-positive plane count %d", planes);
	MorsiCtx *c; rc = morsi_ctx_current(&c); if (rc) return rc;
	MorsiJob job;
	job.op = op; job.w = w; job.h = h; job.lane = 0;
	job.x_row0 = 0; job.x_rows = h; job.x_pstride = (long long)w * h;
	job.y_row0 = 0; job.y_rows = h; job.y_pstride = (long long)w * h;
	job.stream = pick_stream(c, stream);
	// gridDim.z carries the plane index: go in chunks of at most 32768 planes
	for (int p0 = 0; p0 < planes; p0 += 32768) {
		job.planes = planes - p0 < 32768 ? planes - p0 : 32768;
		job.x = d_x + (long long)p0 * w * h;
		job.y = d_y + (long long)p0 * w * h;
		rc = morsi_dispatch(c, e, job);
		if (rc) return rc;
	}
	return MORSI_OK;
}

// Host-pointer entry point: see host_apply.cu
