// shard.cu -- row-band sharding of one plane across the GPUs of a box, in C
// (no torch, no NCCL): SURVEY.md 8(e), BASELINE config C4.
//
// Rank g of N owns output rows [g*h/N, (g+1)*h/N) of a w x h plane and keeps
// them in device memory together with `halo` rows of its vertical neighbours.
// The ranks may be N processes with one device each (torchrun, MPI: the
// 128-byte handles are exchanged by the caller, like an ncclUniqueId) or N
// devices of one process.  Every rank allocates ONE slab (flag words + its
// band buffers) and maps the slabs of its two neighbours -- CUDA IPC across
// processes, peer access inside one -- so a neighbour's halo rows are plain
// global addresses that go over NVLink.
//
// One step of morsi_shard_apply():
//   k_shard_push   tells both neighbours "my previous step is finished, your
//                  pushes may overwrite my halo rows" (credit), waits for
//                  their credit, STORES this rank's boundary rows straight
//                  into the neighbours' halo rows (remote st.global.v4 over
//                  NVLink: stages x reach rows, 4.5 MB each way for disk15 at
//                  w = 40000), fences, and raises their "halo ready" words;
//   kernels        the interior rows -- which need no halo -- are launched
//                  right behind the push, so the transfer hides behind them;
//   k_shard_wait   spins (a single thread) on this rank's two ready words;
//   kernels        the two edge strips.
// Everything is stream-ordered on the rank's own streams (main: interior rows; communication:
// the push; side: the halo wait and the edge strips): no host
// synchronisation, no collective, no reduction -- the path only has this
// point-to-point exchange.  The reference has no counterpart (one thread, one
// process); the rows needed per stage follow src/morsi.c:65, the border rule
// (image-edge bands get no neighbour data, rows outside the image are absent)
// src/morsi.c:30-35.
//
// All ranks must call morsi_shard_apply() / _apply_host() the same number of
// times, in the same order (the credit protocol counts steps).
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <unistd.h>
#include <vector>

#include "dispatch.cuh"

#define SHARD_MAX_BUF 4
#define SHARD_FLAG_BYTES 4096
// flag words inside a slab (each on its own 128-byte line)
enum { F_READY_UP = 0, F_READY_DOWN = 32, F_DONE_UP = 64, F_DONE_DOWN = 96, F_ERR = 128, F_COUNTER = 160 };

struct ShardHandle {              // what travels between ranks (MORSI_SHARD_HANDLE_BYTES)
	unsigned magic;
	int rank, nranks, device;
	long long pid;
	unsigned long long slab_addr;  // valid inside the owning process only
	unsigned long long slab_bytes;
	cudaIpcMemHandle_t ipc;        // 64 bytes
	int w, h, halo, nbuf;
};
static_assert(sizeof(ShardHandle) <= MORSI_SHARD_HANDLE_BYTES, "handle too large");

struct morsi_shard {
	int device, rank, nranks, w, h, halo, nbuf;
	int b0, b1, i0, i1;            // owned rows, held rows
	MorsiCtx *ctx;
	char *slab;
	size_t slab_bytes, buf_bytes;
	char *peer[2];                 // neighbours' slabs in this process' address space (0: up, 1: down)
	bool peer_ipc[2];
	int peer_i0[2];                // first held row of the neighbour
	size_t peer_buf_bytes[2];      // size of one of ITS buffers (bands differ by a row when h % N != 0)
	unsigned step;
	cudaStream_t stream, s_in, s_out, s_comm, s_side;
	cudaEvent_t ev[4];
	std::vector<cudaEvent_t> ev_chunks;
	int overlap;                   // 1: interior rows first, edge strips after the halo (default)
	unsigned timeout_ms;
	long long halo_bytes_last;     // bytes pushed to the neighbours by the last apply
};

static inline int band_first(int h, int rank, int n) { return (int)((long long)h * rank / n); }

// ---- device side ---------------------------------------------------------------------
__device__ __forceinline__ unsigned ld_flag(const unsigned *p)
{
	unsigned v;
	asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
	return v;
}
__device__ __forceinline__ void st_flag(unsigned *p, unsigned v)
{
	asm volatile("st.release.sys.global.u32 [%0], %1;" :: "l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned long long now_ns()
{
	unsigned long long t;
	asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
	return t;
}
// spin until *p >= want (steps only grow); false after the time limit
__device__ __forceinline__ bool wait_flag(const unsigned *p, unsigned want, unsigned long long limit_ns)
{
	const unsigned long long t0 = now_ns();
	int spins = 0;
	while ((int)(ld_flag(p) - want) < 0) {
		if (++spins > 64) { __nanosleep(200); }
		if ((spins & 1023) == 0 && now_ns() - t0 > limit_ns) return false;
	}
	return true;
}

struct PushArgs {
	const float *src[2];           // this rank's boundary rows: [0] its first rows (for the upper neighbour), [1] its last rows
	float *dst[2];                 // the neighbours' halo rows (peer memory)
	long long n[2];                // floats (0 for an operation without vertical reach); the flag pointers are NULL where there is no neighbour
	unsigned *peer_ready[2];       // neighbour's word "halo from below / above is in place"
	unsigned *peer_done[2];        // neighbour's word "the rank below / above finished its previous step"
	unsigned *my_done[2];          // my words, written by the neighbours
	unsigned *counter, *err;
	unsigned step;
	unsigned long long limit_ns;
};

__global__ void __launch_bounds__(256) k_shard_push(PushArgs a)
{
	__shared__ int ok;
	if (threadIdx.x == 0) {
		// credit first (never wait before giving it: two neighbours would deadlock)
		if (blockIdx.x == 0)
			for (int k = 0; k < 2; k++) if (a.peer_done[k]) st_flag(a.peer_done[k], a.step - 1);
		int good = 1;
		for (int k = 0; k < 2; k++)
			if (a.my_done[k] && !wait_flag(a.my_done[k], a.step - 1, a.limit_ns)) good = 0;
		ok = good;
	}
	__syncthreads();
	if (!ok) { if (threadIdx.x == 0) atomicExch(a.err, 1u); return; }
	for (int k = 0; k < 2; k++) {
		const long long n = a.n[k];
		if (!n) continue;
		const float *s = a.src[k];
		float *d = a.dst[k];
		const long long t0 = blockIdx.x * (long long)blockDim.x + threadIdx.x, ts = (long long)gridDim.x * blockDim.x;
		if ((((uintptr_t)s | (uintptr_t)d) & 15) == 0) {
			const long long n4 = n >> 2;
			long long t = t0;
			for (; t + 3 * ts < n4; t += 4 * ts) {      // four loads in flight per thread: NVLink latency is ~2 us
				const float4 v0 = __ldg((const float4 *)s + t), v1 = __ldg((const float4 *)s + t + ts);
				const float4 v2 = __ldg((const float4 *)s + t + 2 * ts), v3 = __ldg((const float4 *)s + t + 3 * ts);
				((float4 *)d)[t] = v0; ((float4 *)d)[t + ts] = v1; ((float4 *)d)[t + 2 * ts] = v2; ((float4 *)d)[t + 3 * ts] = v3;
			}
			for (; t < n4; t += ts) ((float4 *)d)[t] = __ldg((const float4 *)s + t);
			for (long long u = (n4 << 2) + t0; u < n; u += ts) d[u] = s[u];
		} else {
			for (long long t = t0; t < n; t += ts) d[t] = s[t];
		}
	}
	__threadfence_system();
	__syncthreads();
	if (threadIdx.x == 0) {
		const unsigned done = atomicAdd(a.counter, 1u);
		if (done == gridDim.x - 1) {           // the last CTA: every store of the grid is fenced
			*a.counter = 0;
			__threadfence_system();
			for (int k = 0; k < 2; k++) if (a.peer_ready[k]) st_flag(a.peer_ready[k], a.step);
		}
	}
}

__global__ void k_shard_wait(unsigned *ready_up, unsigned *ready_down, unsigned step, unsigned *err, unsigned long long limit_ns)
{
	bool good = true;
	if (ready_up) good &= wait_flag(ready_up, step, limit_ns);
	if (ready_down) good &= wait_flag(ready_down, step, limit_ns);
	if (!good) atomicExch(err, 1u);
	__threadfence_system();
}

// ---- host side -------------------------------------------------------------------------
#define SH_CU(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) \
	return morsi_set_error(e_ == cudaErrorMemoryAllocation ? MORSI_ERR_OOM : MORSI_ERR_COMM, \
		"shard: %s: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); } while (0)

static inline unsigned *flag_of(char *slab, int word) { return (unsigned *)slab + word; }
static inline size_t band_buf_bytes(int held_rows, int w) { return ((size_t)held_rows * w * sizeof(float) + 255) & ~(size_t)255; }
static inline float *buf_of(const morsi_shard *s, char *slab, int b) { return (float *)(slab + SHARD_FLAG_BYTES + (size_t)b * s->buf_bytes); }
static inline float *peer_buf_of(const morsi_shard *s, int k, int b) { return (float *)(s->peer[k] + SHARD_FLAG_BYTES + (size_t)b * s->peer_buf_bytes[k]); }

extern "C" int morsi_shard_create(morsi_shard **out, int device, int rank, int nranks, int w, int h, int halo_rows, int nbuf)
{
	if (!out || rank < 0 || nranks < 1 || rank >= nranks || w <= 0 || h <= 0 || halo_rows < 0 || nbuf < 2 || nbuf > SHARD_MAX_BUF)
		return morsi_set_error(MORSI_ERR_INVALID, "shard_create: bad arguments");
	if (h / nranks < (halo_rows > 0 ? halo_rows : 1))
		return morsi_set_error(MORSI_ERR_INVALID, "shard_create: bands of %d rows are shorter than the %d-row halo", h / nranks, halo_rows);
	MorsiCtx *c;
	int rc = morsi_ctx_get(device, &c);
	if (rc) return rc;
	morsi_shard *s = new morsi_shard();
	s->device = device; s->rank = rank; s->nranks = nranks; s->w = w; s->h = h; s->halo = halo_rows; s->nbuf = nbuf;
	s->ctx = c;
	s->b0 = band_first(h, rank, nranks); s->b1 = band_first(h, rank + 1, nranks);
	s->i0 = std::max(0, s->b0 - halo_rows); s->i1 = std::min(h, s->b1 + halo_rows);
	s->buf_bytes = band_buf_bytes(s->i1 - s->i0, w);
	s->slab_bytes = SHARD_FLAG_BYTES + (size_t)nbuf * s->buf_bytes;
	s->peer[0] = s->peer[1] = nullptr; s->peer_ipc[0] = s->peer_ipc[1] = false;
	s->step = 0;
	s->halo_bytes_last = 0;
	const char *ov = getenv("MORSI_SHARD_OVERLAP");
	s->overlap = ov ? atoi(ov) : 1;
	const char *tm = getenv("MORSI_SHARD_TIMEOUT_MS");
	s->timeout_ms = tm ? (unsigned)atoi(tm) : 20000u;
	cudaError_t e = cudaMalloc(&s->slab, s->slab_bytes);
	if (e != cudaSuccess) {
		const size_t want = s->slab_bytes;
		delete s;
		return morsi_set_error(MORSI_ERR_OOM, "shard_create: %zu bytes on device %d: %s", want, device, cudaGetErrorString(e));
	}
	cudaError_t ce = cudaMemset(s->slab, 0, SHARD_FLAG_BYTES);
	cudaStream_t *streams[5] = {&s->stream, &s->s_in, &s->s_out, &s->s_comm, &s->s_side};
	for (int i = 0; i < 5; i++) { *streams[i] = nullptr; if (ce == cudaSuccess) ce = cudaStreamCreateWithFlags(streams[i], cudaStreamNonBlocking); }
	for (int i = 0; i < 4; i++) { s->ev[i] = nullptr; if (ce == cudaSuccess) ce = cudaEventCreateWithFlags(&s->ev[i], cudaEventDisableTiming); }
	if (ce != cudaSuccess) {
		for (int i = 0; i < 5; i++) if (*streams[i]) cudaStreamDestroy(*streams[i]);
		for (int i = 0; i < 4; i++) if (s->ev[i]) cudaEventDestroy(s->ev[i]);
		cudaFree(s->slab);
		delete s;
		return morsi_set_error(MORSI_ERR_CUDA, "shard_create: %s", cudaGetErrorString(ce));
	}
	*out = s;
	return MORSI_OK;
}

extern "C" int morsi_shard_handle(const morsi_shard *s, void *handle)
{
	if (!s || !handle) return morsi_set_error(MORSI_ERR_INVALID, "shard_handle: NULL");
	ShardHandle hd;
	memset(&hd, 0, sizeof hd);
	hd.magic = 0x4D525348u;
	hd.rank = s->rank; hd.nranks = s->nranks; hd.device = s->device;
	hd.pid = (long long)getpid();
	hd.slab_addr = (unsigned long long)(uintptr_t)s->slab;
	hd.slab_bytes = s->slab_bytes;
	hd.w = s->w; hd.h = s->h; hd.halo = s->halo; hd.nbuf = s->nbuf;
	SH_CU(cudaSetDevice(s->device));
	SH_CU(cudaIpcGetMemHandle(&hd.ipc, s->slab));
	memset(handle, 0, MORSI_SHARD_HANDLE_BYTES);
	memcpy(handle, &hd, sizeof hd);
	return MORSI_OK;
}

extern "C" int morsi_shard_connect(morsi_shard *s, const void *handles)
{
	if (!s || !handles) return morsi_set_error(MORSI_ERR_INVALID, "shard_connect: NULL");
	SH_CU(cudaSetDevice(s->device));
	for (int k = 0; k < 2; k++) {
		const int nb = k == 0 ? s->rank - 1 : s->rank + 1;
		if (nb < 0 || nb >= s->nranks) continue;
		ShardHandle hd;
		memcpy(&hd, (const char *)handles + (size_t)nb * MORSI_SHARD_HANDLE_BYTES, sizeof hd);
		if (hd.magic != 0x4D525348u || hd.rank != nb || hd.nranks != s->nranks || hd.w != s->w || hd.h != s->h ||
				hd.halo != s->halo || hd.nbuf != s->nbuf)
			return morsi_set_error(MORSI_ERR_COMM, "shard_connect: handle %d does not describe rank %d of this plane", nb, nb);
		if (hd.pid == (long long)getpid()) {
			// same process: the neighbour's slab is addressable once peer access is on
			if (hd.device != s->device) {
				int can = 0;
				SH_CU(cudaDeviceCanAccessPeer(&can, s->device, hd.device));
				if (!can) return morsi_set_error(MORSI_ERR_COMM, "shard_connect: device %d cannot access device %d", s->device, hd.device);
				cudaError_t e = cudaDeviceEnablePeerAccess(hd.device, 0);
				if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled)
					return morsi_set_error(MORSI_ERR_COMM, "shard_connect: cudaDeviceEnablePeerAccess(%d): %s", hd.device, cudaGetErrorString(e));
				cudaGetLastError();
			}
			s->peer[k] = (char *)(uintptr_t)hd.slab_addr;
			s->peer_ipc[k] = false;
		} else {
			void *p = nullptr;
			cudaError_t e = cudaIpcOpenMemHandle(&p, hd.ipc, cudaIpcMemLazyEnablePeerAccess);
			if (e != cudaSuccess)
				return morsi_set_error(MORSI_ERR_COMM, "shard_connect: cudaIpcOpenMemHandle(rank %d, device %d): %s", nb, hd.device, cudaGetErrorString(e));
			s->peer[k] = (char *)p;
			s->peer_ipc[k] = true;
		}
		s->peer_i0[k] = std::max(0, band_first(s->h, nb, s->nranks) - s->halo);
		s->peer_buf_bytes[k] = band_buf_bytes(std::min(s->h, band_first(s->h, nb + 1, s->nranks) + s->halo) - s->peer_i0[k], s->w);
		if (SHARD_FLAG_BYTES + (size_t)s->nbuf * s->peer_buf_bytes[k] != hd.slab_bytes)
			return morsi_set_error(MORSI_ERR_COMM, "shard_connect: rank %d's slab is %llu bytes, expected %zu", nb, hd.slab_bytes, SHARD_FLAG_BYTES + (size_t)s->nbuf * s->peer_buf_bytes[k]);
	}
	return MORSI_OK;
}

extern "C" int morsi_shard_rows(const morsi_shard *s, int *own_row0, int *own_rows, int *held_row0, int *held_rows)
{
	if (!s) return morsi_set_error(MORSI_ERR_INVALID, "shard_rows: NULL");
	if (own_row0) *own_row0 = s->b0;
	if (own_rows) *own_rows = s->b1 - s->b0;
	if (held_row0) *held_row0 = s->i0;
	if (held_rows) *held_rows = s->i1 - s->i0;
	return MORSI_OK;
}

extern "C" float *morsi_shard_buffer(morsi_shard *s, int buf)
{
	if (!s || buf < 0 || buf >= s->nbuf) return nullptr;
	return buf_of(s, s->slab, buf);
}
extern "C" void *morsi_shard_stream(morsi_shard *s) { return s ? (void *)s->stream : nullptr; }
extern "C" long long morsi_shard_halo_bytes(const morsi_shard *s) { return s ? s->halo_bytes_last : 0; }

static int shard_job(morsi_shard *s, int op, const int *e, int src, int dst, int r0, int r1, cudaStream_t st, int lane = 0)
{
	if (r1 <= r0) return MORSI_OK;
	MorsiJob job;
	job.op = op; job.w = s->w; job.h = s->h; job.planes = 1; job.lane = lane;
	job.x = buf_of(s, s->slab, src); job.x_row0 = s->i0; job.x_rows = s->i1 - s->i0; job.x_pstride = (long long)s->w * (s->i1 - s->i0);
	job.y = buf_of(s, s->slab, dst) + (size_t)(r0 - s->i0) * s->w; job.y_row0 = r0; job.y_rows = r1 - r0;
	job.y_pstride = (long long)s->w * (s->i1 - s->i0);
	job.stream = st;
	return morsi_dispatch(s->ctx, e, job);
}

// the exchange of one step: credit, push, ready.  It runs on the rank's
// communication stream, behind `after` (everything the step's source buffer
// and the previous step depend on), so that the interior rows on the main
// stream overlap it.
static int shard_push(morsi_shard *s, int src, int up, int down, cudaEvent_t after)
{
	s->halo_bytes_last = 0;
	if (s->nranks == 1) return MORSI_OK;
	SH_CU(cudaStreamWaitEvent(s->s_comm, after, 0));
	PushArgs a;
	memset(&a, 0, sizeof a);
	float *mine = buf_of(s, s->slab, src);
	const int own = s->b1 - s->b0;
	if (s->rank > 0) {
		if (!s->peer[0]) return morsi_set_error(MORSI_ERR_COMM, "shard_apply: rank %d is not connected (morsi_shard_connect)", s->rank);
		// my first `down` rows are the upper neighbour's bottom halo
		const int n = std::min(down, own);
		a.src[0] = mine + (size_t)(s->b0 - s->i0) * s->w;
		a.dst[0] = peer_buf_of(s, 0, src) + (size_t)(s->b0 - s->peer_i0[0]) * s->w;
		a.n[0] = (long long)n * s->w;
		a.peer_ready[0] = flag_of(s->peer[0], F_READY_DOWN);
		a.peer_done[0] = flag_of(s->peer[0], F_DONE_DOWN);
		a.my_done[0] = flag_of(s->slab, F_DONE_UP);
	}
	if (s->rank < s->nranks - 1) {
		if (!s->peer[1]) return morsi_set_error(MORSI_ERR_COMM, "shard_apply: rank %d is not connected (morsi_shard_connect)", s->rank);
		const int n = std::min(up, own);
		a.src[1] = mine + (size_t)(s->b1 - n - s->i0) * s->w;
		a.dst[1] = peer_buf_of(s, 1, src) + (size_t)(s->b1 - n - s->peer_i0[1]) * s->w;
		a.n[1] = (long long)n * s->w;
		a.peer_ready[1] = flag_of(s->peer[1], F_READY_UP);
		a.peer_done[1] = flag_of(s->peer[1], F_DONE_UP);
		a.my_done[1] = flag_of(s->slab, F_DONE_DOWN);
	}
	// (an operation without vertical reach moves no rows but still runs the credit / ready protocol)
	a.counter = flag_of(s->slab, F_COUNTER);
	a.err = flag_of(s->slab, F_ERR);
	a.step = s->step;
	a.limit_ns = (unsigned long long)s->timeout_ms * 1000000ull;
	s->halo_bytes_last = (a.n[0] + a.n[1]) * (long long)sizeof(float);
	// enough CTAs to fill the NVLink pipe, few enough to be co-resident at once (they spin on the credit)
	k_shard_push<<<64, 256, 0, s->s_comm>>>(a);
	morsi_count_launch(1);
	SH_CU(cudaGetLastError());
	return MORSI_OK;
}

static int shard_wait(morsi_shard *s, cudaStream_t st)
{
	if (s->nranks == 1) return MORSI_OK;
	k_shard_wait<<<1, 1, 0, st>>>(s->rank > 0 ? flag_of(s->slab, F_READY_UP) : nullptr,
			s->rank < s->nranks - 1 ? flag_of(s->slab, F_READY_DOWN) : nullptr, s->step, flag_of(s->slab, F_ERR),
			(unsigned long long)s->timeout_ms * 1000000ull);
	morsi_count_launch(1);
	SH_CU(cudaGetLastError());
	return MORSI_OK;
}

static int shard_check_op(morsi_shard *s, int op, const int *e, int src, int dst, int *up, int *down)
{
	if (!s) return morsi_set_error(MORSI_ERR_INVALID, "shard: NULL");
	if (src < 0 || src >= s->nbuf || dst < 0 || dst >= s->nbuf || src == dst)
		return morsi_set_error(MORSI_ERR_INVALID, "shard: bad buffer indices %d -> %d", src, dst);
	int rc = morsi_cuda_halo_rows(op, e, up, down);
	if (rc) return rc;
	if (*up > s->halo || *down > s->halo)
		return morsi_set_error(MORSI_ERR_INVALID, "shard: the operation needs %d/%d halo rows, the shard holds %d", *up, *down, s->halo);
	SH_CU(cudaSetDevice(s->device));
	return MORSI_OK;
}

extern "C" int morsi_shard_apply(morsi_shard *s, int op, const int *e, int src, int dst)
{
	int up, down;
	int rc = shard_check_op(s, op, e, src, dst, &up, &down);
	if (rc) return rc;
	s->step++;
	SH_CU(cudaEventRecord(s->ev[2], s->stream));       // the previous step (and whatever filled the source buffer)
	if ((rc = shard_push(s, src, up, down, s->ev[2]))) return rc;
	const int top = s->rank > 0 ? up : 0, bot = s->rank < s->nranks - 1 ? down : 0;
	const int own = s->b1 - s->b0;
	if (s->overlap && s->nranks > 1 && own > 4 * (top + bot) + 64) {
		// interior rows on the main stream; the halo wait and the two edge strips on a side
		// stream (its own flag / workspace lane), so that the strips start the moment the halo
		// has landed and fill the SMs the interior's last wave leaves idle
		SH_CU(cudaStreamWaitEvent(s->s_side, s->ev[2], 0));
		if ((rc = shard_job(s, op, e, src, dst, s->b0 + top, s->b1 - bot, s->stream))) return rc;
		if ((rc = shard_wait(s, s->s_side))) return rc;
		if ((rc = shard_job(s, op, e, src, dst, s->b0, s->b0 + top, s->s_side, MORSI_LANE_SHARD_SIDE))) return rc;
		if ((rc = shard_job(s, op, e, src, dst, s->b1 - bot, s->b1, s->s_side, MORSI_LANE_SHARD_SIDE))) return rc;
		SH_CU(cudaEventRecord(s->ev[3], s->s_side));
		SH_CU(cudaStreamWaitEvent(s->stream, s->ev[3], 0));
	} else {
		if ((rc = shard_wait(s, s->stream))) return rc;
		if ((rc = shard_job(s, op, e, src, dst, s->b0, s->b1, s->stream))) return rc;
	}
	return MORSI_OK;
}

// Refresh the halo rows of buffer `buf` from the neighbours (`up` rows above,
// `down` below) without computing anything: the exchange of one step on its own.
extern "C" int morsi_shard_exchange(morsi_shard *s, int buf, int up, int down)
{
	if (!s || buf < 0 || buf >= s->nbuf || up < 0 || down < 0 || up > s->halo || down > s->halo)
		return morsi_set_error(MORSI_ERR_INVALID, "shard_exchange: bad arguments");
	SH_CU(cudaSetDevice(s->device));
	s->step++;
	SH_CU(cudaEventRecord(s->ev[2], s->stream));
	int rc = shard_push(s, buf, up, down, s->ev[2]);
	if (rc) return rc;
	return shard_wait(s, s->stream);
}

extern "C" int morsi_shard_sync(morsi_shard *s)
{
	if (!s) return morsi_set_error(MORSI_ERR_INVALID, "shard_sync: NULL");
	SH_CU(cudaSetDevice(s->device));
	SH_CU(cudaStreamSynchronize(s->stream));
	SH_CU(cudaStreamSynchronize(s->s_in));
	SH_CU(cudaStreamSynchronize(s->s_out));
	SH_CU(cudaStreamSynchronize(s->s_comm));
	SH_CU(cudaStreamSynchronize(s->s_side));
	unsigned err = 0;
	SH_CU(cudaMemcpy(&err, flag_of(s->slab, F_ERR), sizeof err, cudaMemcpyDeviceToHost));
	if (err)
		return morsi_set_error(MORSI_ERR_COMM, "shard: rank %d timed out waiting for a neighbour (step %u; all ranks must apply the same steps)", s->rank, s->step);
	return MORSI_OK;
}

// Host band in, host band out (x, y: this rank's OWNED rows, row pitch w;
// pinned memory gives the overlap).  The boundary rows cross PCIe first and
// are pushed to the neighbours at once; the band then streams through in row
// chunks on three streams (upload / kernels / download), chunk k computing
// while chunk k+1 arrives and chunk k-1 leaves.  src = buffer 0, dst = buffer 1.
extern "C" int morsi_shard_apply_host(morsi_shard *s, int op, const int *e, const float *x, float *y)
{
	int up, down;
	int rc = shard_check_op(s, op, e, 0, 1, &up, &down);
	if (rc) return rc;
	if (!x || !y) return morsi_set_error(MORSI_ERR_INVALID, "shard_apply_host: NULL band");
	s->step++;
	const int w = s->w, own = s->b1 - s->b0;
	float *d_in = buf_of(s, s->slab, 0) + (size_t)(s->b0 - s->i0) * w;     // owned row 0
	float *d_out = buf_of(s, s->slab, 1) + (size_t)(s->b0 - s->i0) * w;
	const size_t row_bytes = (size_t)w * sizeof(float);
	// the previous step's kernels (and, through the credit, the neighbours') are
	// done with buffer 0 before the uploads overwrite it
	SH_CU(cudaEventRecord(s->ev[0], s->stream));
	SH_CU(cudaStreamWaitEvent(s->s_in, s->ev[0], 0));
	if (s->nranks > 1) {
		const int ntop = s->rank > 0 ? std::min(down, own) : 0, nbot = s->rank < s->nranks - 1 ? std::min(up, own) : 0;
		if (ntop) SH_CU(cudaMemcpyAsync(d_in, x, ntop * row_bytes, cudaMemcpyHostToDevice, s->s_in));
		if (nbot) SH_CU(cudaMemcpyAsync(d_in + (size_t)(own - nbot) * w, x + (size_t)(own - nbot) * w, nbot * row_bytes, cudaMemcpyHostToDevice, s->s_in));
		SH_CU(cudaEventRecord(s->ev[1], s->s_in));
		if ((rc = shard_push(s, 0, up, down, s->ev[1]))) return rc;
	}
	long long target = (64LL << 20) / (long long)row_bytes;
	int chunk = (int)std::max<long long>(std::max(64, 8 * (up + down)), target);
	if (const char *cr = getenv("MORSI_SHARD_CHUNK_ROWS")) chunk = std::max(1, atoi(cr));
	const int nchunks = (own + chunk - 1) / chunk;
	while ((int)s->ev_chunks.size() < 2 * nchunks + 2) {
		cudaEvent_t evn;
		SH_CU(cudaEventCreateWithFlags(&evn, cudaEventDisableTiming));
		s->ev_chunks.push_back(evn);
	}
	for (int k = 0; k < nchunks; k++) {
		const int r0 = k * chunk, r1 = std::min(own, r0 + chunk);
		SH_CU(cudaMemcpyAsync(d_in + (size_t)r0 * w, x + (size_t)r0 * w, (size_t)(r1 - r0) * row_bytes, cudaMemcpyHostToDevice, s->s_in));
		SH_CU(cudaEventRecord(s->ev_chunks[2 * k], s->s_in));
	}
	bool waited = false;
	for (int k = 0; k < nchunks; k++) {
		const int r0 = k * chunk, r1 = std::min(own, r0 + chunk);
		// rows [r0-up, r1+down) of the band: the next chunk must have landed too
		SH_CU(cudaStreamWaitEvent(s->stream, s->ev_chunks[2 * std::min(k + 1, nchunks - 1)], 0));
		if (!waited) { if ((rc = shard_wait(s, s->stream))) return rc; waited = true; }
		if ((rc = shard_job(s, op, e, 0, 1, s->b0 + r0, s->b0 + r1, s->stream))) return rc;
		SH_CU(cudaEventRecord(s->ev_chunks[2 * k + 1], s->stream));
		SH_CU(cudaStreamWaitEvent(s->s_out, s->ev_chunks[2 * k + 1], 0));
		SH_CU(cudaMemcpyAsync(y + (size_t)r0 * w, d_out + (size_t)r0 * w, (size_t)(r1 - r0) * row_bytes, cudaMemcpyDeviceToHost, s->s_out));
	}
	return morsi_shard_sync(s);
}

extern "C" int morsi_shard_destroy(morsi_shard *s)
{
	if (!s) return MORSI_OK;
	cudaSetDevice(s->device);
	cudaStreamSynchronize(s->stream);
	cudaStreamSynchronize(s->s_in);
	cudaStreamSynchronize(s->s_out);
	cudaStreamSynchronize(s->s_comm);
	cudaStreamSynchronize(s->s_side);
	for (int k = 0; k < 2; k++) if (s->peer[k] && s->peer_ipc[k]) cudaIpcCloseMemHandle(s->peer[k]);
	for (int i = 0; i < 4; i++) cudaEventDestroy(s->ev[i]);
	for (cudaEvent_t evn : s->ev_chunks) cudaEventDestroy(evn);
	cudaStreamDestroy(s->stream); cudaStreamDestroy(s->s_in); cudaStreamDestroy(s->s_out); cudaStreamDestroy(s->s_comm); cudaStreamDestroy(s->s_side);
	cudaFree(s->slab);
	delete s;
	return MORSI_OK;
}

// ---- one process, N devices -------------------------------------------------------------
// Host plane in, host plane out, `iterations` applications of the operation
// (y = op(op(...op(x)))): the plane is cut into N row bands, one per device;
// the bands stay resident between iterations and only the halo rows travel,
// device to device (peer stores over NVLink).  iterations == 1 is the plain
// operation on an image too large (or too slow) for one device.
extern "C" int morsi_cuda_apply_sharded(int op, const int *e, const float *x, float *y, int w, int h, int ndev, int iterations)
{
	if (op < 0 || op >= MORSI_OP_COUNT || !e || e[0] < 0 || !x || !y || w <= 0 || h <= 0 || ndev < 1 || iterations < 1)
		return morsi_set_error(MORSI_ERR_INVALID, "apply_sharded: bad arguments");
	const int navail = morsi_cuda_device_count();
	if (navail <= 0) return morsi_set_error(MORSI_ERR_NO_DEVICE, "no CUDA device visible");
	if (ndev > navail) return morsi_set_error(MORSI_ERR_INVALID, "apply_sharded: %d devices asked, %d visible", ndev, navail);
	int up, down;
	int rc = morsi_cuda_halo_rows(op, e, &up, &down);
	if (rc) return rc;
	const int halo = std::max(up, down);
	std::vector<morsi_shard *> sh(ndev, nullptr);
	std::vector<char> handles((size_t)ndev * MORSI_SHARD_HANDLE_BYTES);
	auto cleanup = [&]() { for (morsi_shard *s : sh) morsi_shard_destroy(s); };
	for (int d = 0; d < ndev && !rc; d++) {
		rc = morsi_shard_create(&sh[d], d, d, ndev, w, h, halo, 2);
		if (!rc) rc = morsi_shard_handle(sh[d], handles.data() + (size_t)d * MORSI_SHARD_HANDLE_BYTES);
	}
	for (int d = 0; d < ndev && !rc; d++) rc = morsi_shard_connect(sh[d], handles.data());
	if (rc) { cleanup(); return rc; }
	// upload the owned rows (pageable or pinned: the copies of different devices overlap either way)
	for (int d = 0; d < ndev && !rc; d++) {
		morsi_shard *s = sh[d];
		cudaSetDevice(s->device);
		if (cudaMemcpyAsync(buf_of(s, s->slab, 0) + (size_t)(s->b0 - s->i0) * w, x + (size_t)s->b0 * w,
				(size_t)(s->b1 - s->b0) * w * sizeof(float), cudaMemcpyHostToDevice, s->stream) != cudaSuccess)
			rc = morsi_set_error(MORSI_ERR_CUDA, "apply_sharded: upload to device %d failed: %s", d, cudaGetErrorString(cudaGetLastError()));
	}
	// every iteration: all ranks enqueue their step (asynchronous; the device-side flags order them)
	for (int it = 0; it < iterations && !rc; it++)
		for (int d = 0; d < ndev && !rc; d++) rc = morsi_shard_apply(sh[d], op, e, it & 1, (it & 1) ^ 1);
	for (int d = 0; d < ndev && !rc; d++) {
		morsi_shard *s = sh[d];
		cudaSetDevice(s->device);
		if (cudaMemcpyAsync(y + (size_t)s->b0 * w, buf_of(s, s->slab, iterations & 1) + (size_t)(s->b0 - s->i0) * w,
				(size_t)(s->b1 - s->b0) * w * sizeof(float), cudaMemcpyDeviceToHost, s->stream) != cudaSuccess)
			rc = morsi_set_error(MORSI_ERR_CUDA, "apply_sharded: download from device %d failed: %s", d, cudaGetErrorString(cudaGetLastError()));
	}
	for (int d = 0; d < ndev; d++) { int r2 = morsi_shard_sync(sh[d]); if (!rc) rc = r2; }
	cleanup();
	return rc;
}
