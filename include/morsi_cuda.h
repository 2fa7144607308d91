/*
 * morsi_cuda.h -- C ABI of libmorsi_cuda, the B200 (sm_100a) implementation of
 * imscript's `morsi` gray-scale morphology hot path.
 *
 * Every entry point names the reference interface it replaces (paths relative
 * to the reference root, mnhrdt/imscript).  Plain C: pointers, ints, sizes.
 * There is no CPU fallback: without a usable CUDA device every compute call
 * returns MORSI_ERR_NO_DEVICE / MORSI_ERR_CUDA.
 *
 * Data contract (src/morsi.c:34,66,543; src/iio.h:42-43): IEEE float32,
 * planar, plane k at x + k*w*h, sample (i,j) of a plane at i + j*w, x fastest.
 * A structuring element is the reference's int list (src/morsi.c:48-54):
 * e[0]=count, e[1]=flags(0), (e[2],e[3])=centre, then (dx,dy) pairs from e[4];
 * neighbour k of (i,j) is (i-e[2]+e[2k+4], j-e[3]+e[2k+5]).  Arbitrary lists
 * (non-zero centre, repeats, any order) are legal: that is the reference's
 * "user-defined mask" mechanism (src/ftr/webcam/corrview.c:37).
 */
#ifndef MORSI_CUDA_H
#define MORSI_CUDA_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Operations, in the dispatcher order of src/morsi.c:510-527. */
enum morsi_op {
	MORSI_EROSION = 0,  /* src/morsi.c:56-68   */
	MORSI_DILATION,     /* src/morsi.c:70-82   */
	MORSI_MEDIAN,       /* src/morsi.c:103-120 */
	MORSI_RANK,         /* src/morsi.c:122-139 */
	MORSI_OPENING,      /* src/morsi.c:141-147 */
	MORSI_CLOSING,      /* src/morsi.c:149-155 */
	MORSI_GRADIENT,     /* src/morsi.c:157-167 */
	MORSI_IGRADIENT,    /* src/morsi.c:169-176 */
	MORSI_EGRADIENT,    /* src/morsi.c:178-185 */
	MORSI_LAPLACIAN,    /* src/morsi.c:187-197 */
	MORSI_ENHANCE,      /* src/morsi.c:199-206 */
	MORSI_BLUR,         /* src/morsi.c:208-215 */
	MORSI_OSCILLATION,  /* src/morsi.c:217-227 */
	MORSI_TOPHAT,       /* src/morsi.c:229-236 */
	MORSI_BOTHAT,       /* src/morsi.c:238-245 */
	MORSI_IBLUR,        /* src/morsi.c:247-254 */
	MORSI_EBLUR,        /* src/morsi.c:256-263 */
	MORSI_CBLUR,        /* src/morsi.c:265-275 */
	MORSI_OP_COUNT
};

/* Return codes (the reference functions return void and exit(-1) through
 * fail(), src/xmalloc.c:18-27, src/fail.c:65-81; here errors come back). */
enum morsi_status {
	MORSI_OK = 0,
	MORSI_ERR_INVALID = 1,    /* bad op / NULL pointer / non-positive size / bad element */
	MORSI_ERR_NO_DEVICE = 2,  /* no CUDA device: there is no CPU fallback */
	MORSI_ERR_CUDA = 3,       /* a CUDA call failed; see morsi_cuda_last_error() */
	MORSI_ERR_OOM = 4,        /* device or host allocation failed */
	MORSI_ERR_TOO_LARGE = 5,  /* w*h*planes beyond what one call supports */
	MORSI_ERR_COMM = 6        /* halo exchange / peer access failure */
};

const char *morsi_cuda_strerror(int status);
/* Detail of the last failure on the calling thread ("" if none). */
const char *morsi_cuda_last_error(void);

/* ---- structuring elements and names (host side, no GPU needed) ---------- */

/* Element-name grammar of src/morsi.c:484-485,496-508 ("cross", "square",
 * "diskR", "dyskR", "hrecR", "vrecR", "drecR", "DrecR", with the reference's
 * strspn() semantics).  On success *e receives a malloc'd list to be released
 * with morsi_element_free(); returns MORSI_ERR_INVALID where the reference
 * prints "elements = cross, square ..." and exits 1. */
int morsi_element_parse(const char *name, int **e);
/* The builders of src/morsi.c:313-417; NULL when radius <= 1 (or NaN). */
int *morsi_build_disk(float radius);   /* src/morsi.c:313-330 */
int *morsi_build_dysk(float radius);   /* src/morsi.c:332-349 */
int *morsi_build_hrec(float radius);   /* src/morsi.c:351-366 */
int *morsi_build_vrec(float radius);   /* src/morsi.c:368-383 */
int *morsi_build_drec(float radius);   /* src/morsi.c:385-400 */
int *morsi_build_Drec(float radius);   /* src/morsi.c:402-417 */
void morsi_element_free(int *e);
/* Operation-name table of src/morsi.c:509-527: index, or -1 if unknown. */
int morsi_operation_parse(const char *name);
const char *morsi_operation_name(int op);

/* How the element compiler classified a list (for reports and tests):
 * writes a short tag such as "small3x3", "rowrun", "direct" into buf. */
int morsi_element_describe(const int *e, char *buf, size_t buflen);

/* ---- device management -------------------------------------------------- */

int morsi_cuda_device_count(void);
/* Select the device used by the calling process for the *_device entry
 * points and create its context (stream, workspace).  Optional: the first
 * compute call does morsi_cuda_init(0) implicitly. */
int morsi_cuda_init(int device);
void morsi_cuda_shutdown(void);

/* ---- the hot path -------------------------------------------------------- */

/* Replaces the channel loop of src/morsi.c:539-543 around
 * operation(y+k*w*h, x+k*w*h, w, h, e): HOST pointers in, HOST pointers out,
 * synchronous.  Copies to the device(s), runs the sm_100a kernels, copies back.
 * When MORSI_CUDA_DEVICES=N (N>1) is set, planes (or, for a single plane, row
 * bands) are spread over N devices; on this entry point every device takes the
 * halo rows of its bands from the caller's host buffer together with the band,
 * so no device-to-device traffic is needed.  Device-resident sharding with a
 * halo exchange over NVLink is morsi_shard_* / morsi_cuda_apply_sharded below.
 * Thread safety: host-pointer calls on one device are serialised internally. */
int morsi_cuda_apply(int op, const int *e, const float *x, float *y,
		int w, int h, int planes);

/* morsi_all (src/morsi.c:278-310) for host pointers: out[] = {erosion, dilation,
 * opening, closing, gradient, igradient, egradient, laplacian, enhance,
 * oscillation ("o_str"), tophat, bothat}; NULL entries are skipped.  The input
 * is copied to the device once; erosion and dilation are computed once and
 * shared by every output exactly as the reference does (:289-295), the
 * pointwise outputs (:296-303) come from one pass over them.  Every result
 * equals the single operation's. */
int morsi_cuda_apply_all(const int *e, const float *x, float *const out[12],
		int w, int h, int planes);
/* The same on device-resident planar data, asynchronous on `stream`. */
int morsi_cuda_apply_all_device(const int *e, const float *d_x, float *const d_out[12],
		int w, int h, int planes, void *stream);

/* Pixel-interleaved images (iio's "vec" layout, sample c of pixel i at
 * x[i*pd + c]): the conversion to float and the split into planes that
 * iio_read_image_float_split does on the CPU (src/iio.c:1416-1428, 5763-5771;
 * sample conversions :1139-1158) happen on the device, and so does the join
 * on the way out (src/iio.c:1423, 6525-6531).  An 8-bit image crosses PCIe as
 * bytes.  y receives float32 samples in the same interleaved layout. */
enum morsi_sample_type { MORSI_SAMPLE_U8 = 0, MORSI_SAMPLE_U16 = 1, MORSI_SAMPLE_F32 = 2 };
int morsi_cuda_apply_interleaved(int op, const int *e, const void *x, float *y,
		int w, int h, int pd, int sample_type);

/* Streaming: the image is pulled and pushed in row bands through callbacks, so
 * neither the host nor the device ever holds it as a whole (the job
 * src/fancy_image.h:40-70 does for the reference's tools; all sizes are
 * size_t / long long here, unlike src/iio.c:3759,4073).  rd must fill dst with
 * rows [row0,row0+nrows) of `plane` (row pitch w); wr receives finished rows,
 * in order, plane by plane.  A non-zero return aborts with MORSI_ERR_INVALID. */
typedef int (*morsi_read_rows_fn)(void *user, int plane, int row0, int nrows, float *dst);
typedef int (*morsi_write_rows_fn)(void *user, int plane, int row0, int nrows, const float *src);
int morsi_cuda_apply_stream(int op, const int *e, int w, int h, int planes,
		morsi_read_rows_fn rd, morsi_write_rows_fn wr, void *user);

/* Same computation on DEVICE-resident planar data (row pitch = w), enqueued on
 * `stream` (a cudaStream_t; NULL = the context's stream) of the current
 * device, asynchronous. d_x and d_y must not overlap. */
int morsi_cuda_apply_device(int op, const int *e, const float *d_x, float *d_y,
		int w, int h, int planes, void *stream);

/* Row-band form, for images sharded across devices or streamed in tiles.
 * The plane has `h` rows in total; d_x holds its rows [x_row0, x_row0+x_rows),
 * d_y receives rows [y_row0, y_row0+y_rows) (row pitch w in both).  Rows outside
 * [0,h) are absent (the NaN rule of src/morsi.c:30-35); every in-image row
 * within morsi_cuda_halo_rows(op,e) of the output band must be present in d_x,
 * otherwise MORSI_ERR_INVALID. */
int morsi_cuda_apply_band_device(int op, const int *e,
		const float *d_x, int x_row0, int x_rows,
		float *d_y, int y_row0, int y_rows,
		int w, int h, void *stream);
/* Input rows needed above (*up) and below (*down) an output band. */
int morsi_cuda_halo_rows(int op, const int *e, int *up, int *down);

/* ---- row-band sharding of one plane across devices ---------
 * The reference is one thread in one process; this is how an image too large
 * (or too slow) for one GPU, or an iterated operation on it, runs on the GPUs
 * of a box.  Rank g of nranks owns rows [g*h/nranks, (g+1)*h/nranks) and holds
 * them plus `halo_rows` rows of each vertical neighbour in `nbuf` (2..4)
 * device buffers of equal shape.  A step pushes the rank's boundary rows
 * straight into the neighbours' halo rows (peer stores over NVLink, CUDA IPC
 * between processes / peer access inside one; device-side flag words order the
 * ranks, no host synchronisation, no collective) and runs the kernels of
 * morsi_cuda_apply_band_device: the interior rows overlap the transfer, the
 * edge strips follow it.  Ranks are processes with one device each or devices
 * of one process; all ranks must issue the same sequence of apply calls.
 * Image-edge bands get no neighbour data (src/morsi.c:30-35); the halo an
 * operation needs is stages x reach rows (src/morsi.c:65, morsi_cuda_halo_rows).
 * Failures of the exchange (unreachable peer, a neighbour that never arrives
 * within MORSI_SHARD_TIMEOUT_MS) return MORSI_ERR_COMM. */
typedef struct morsi_shard morsi_shard;
#define MORSI_SHARD_HANDLE_BYTES 128
int morsi_shard_create(morsi_shard **s, int device, int rank, int nranks,
		int w, int h, int halo_rows, int nbuf);
/* A rank's handle (MORSI_SHARD_HANDLE_BYTES bytes, position independent): the
 * caller gathers the handles of all ranks -- any transport: MPI, a file,
 * torch.distributed -- and passes the nranks x 128-byte table to connect(). */
int morsi_shard_handle(const morsi_shard *s, void *handle);
int morsi_shard_connect(morsi_shard *s, const void *handles);
int morsi_shard_rows(const morsi_shard *s, int *own_row0, int *own_rows,
		int *held_row0, int *held_rows);
/* Device pointer of the first HELD row of buffer `buf` (row pitch w). */
float *morsi_shard_buffer(morsi_shard *s, int buf);
/* The rank's stream (a cudaStream_t): fill buffers and record events on it. */
void *morsi_shard_stream(morsi_shard *s);
/* dst's owned rows = op(src) : halo exchange + kernels, asynchronous. */
int morsi_shard_apply(morsi_shard *s, int op, const int *e, int src_buf, int dst_buf);
/* Host rows in, host rows out (this rank's OWNED rows, pitch w; pinned memory
 * overlaps): boundary rows first and pushed at once, then the band streams
 * through in chunks on upload / kernel / download streams.  Synchronous. */
int morsi_shard_apply_host(morsi_shard *s, int op, const int *e, const float *x_own, float *y_own);
/* The exchange of a step on its own: refresh the halo rows of `buf` (up rows
 * above, down rows below) from the neighbours; asynchronous. */
int morsi_shard_exchange(morsi_shard *s, int buf, int up, int down);
/* Wait for the rank's streams; MORSI_ERR_COMM if an exchange timed out. */
int morsi_shard_sync(morsi_shard *s);
long long morsi_shard_halo_bytes(const morsi_shard *s);   /* pushed by the last apply */
int morsi_shard_destroy(morsi_shard *s);
/* One process, ndev devices: host plane in, host plane out, the operation
 * applied `iterations` times with the bands resident between iterations and
 * only halo rows travelling device to device. */
int morsi_cuda_apply_sharded(int op, const int *e, const float *x, float *y,
		int w, int h, int ndev, int iterations);

/* Pipe chains: `morsi E1 OP1 in | morsi E2 OP2 | qeasy black white - out`
 * (doc/tutorial/i.html:221-225, src/qeasy.c:55-73) as ONE call: one upload, the nops
 * operations back to back on the device, optionally qeasy's quantiser
 * floor(255 (v - black) / (white - black)) as the last kernel.  y receives planar
 * float32 (q == NULL or !q->to_uint8: qeasy -f) or planar uint8 saturated to
 * [0,255] (to_uint8: the 8-bit image crosses PCIe as bytes). */
typedef struct morsi_quantizer { float black, white; int to_uint8; } morsi_quantizer;
int morsi_cuda_apply_chain(int nops, const int *ops, const int *const *elements,
		const float *x, void *y, int w, int h, int planes, const morsi_quantizer *q);

/* Force a kernel family: 0 = automatic, 1 = order-preserving exact kernels
 * only (the signed-zero-safe path), 2 = fast kernels without the signed-zero
 * re-run (benchmark use).  Also settable with MORSI_CUDA_PATH=auto|exact|fast. */
int morsi_cuda_set_path(int path);

/* Number of kernels launched by this process since the last reset, and a
 * reset; bench.py reports it as "gpu_launches". */
long morsi_cuda_launch_count(void);
void morsi_cuda_launch_count_reset(void);

/* ---- plumbing for callers without a CUDA runtime of their own ----------- */

int morsi_cuda_malloc(void **d_ptr, size_t bytes);
int morsi_cuda_free(void *d_ptr);
int morsi_cuda_host_alloc(void **h_ptr, size_t bytes);   /* pinned */
int morsi_cuda_host_free(void *h_ptr);
int morsi_cuda_memcpy_h2d(void *d_dst, const void *h_src, size_t bytes, void *stream);
int morsi_cuda_memcpy_d2h(void *h_dst, const void *d_src, size_t bytes, void *stream);
int morsi_cuda_memcpy_d2d(void *d_dst, const void *d_src, size_t bytes, void *stream);
int morsi_cuda_sync(void *stream);
int morsi_cuda_stream_create(void **stream);
int morsi_cuda_stream_destroy(void *stream);
/* Device-side synthetic image of SURVEY.md 8(d): distribution 0 = uniform
 * [0,1) on a 2^-24 grid, 1 = integers 0..255, 2 = distribution 0 with NaN /
 * +-Inf / +-0 sprinkled in; value = f(seed, plane, row, column), so bands and
 * crops of the same image can be generated independently. */
int morsi_cuda_synth(float *d_x, int w, int rows, int row0, int plane,
		unsigned seed, int distribution, void *stream);
void morsi_synth_host(float *x, int w, int rows, int row0, int plane,
		unsigned seed, int distribution);
/* Events, for timing on the launching stream. */
int morsi_cuda_event_create(void **ev);
int morsi_cuda_event_record(void *ev, void *stream);
int morsi_cuda_event_elapsed_ms(void *ev_start, void *ev_stop, float *ms);
int morsi_cuda_event_destroy(void *ev);

/* ---- the reference's own function signatures ----------------------------
 * Drop-in for library-style callers of src/morsi.c (corrview.c:37-38,71-72):
 * same names, same arguments, host pointers; on failure they print one line
 * to stderr and exit(-1), the reference's fail() convention.  They live in
 * libmorsi_compat so the names do not clash inside the oracle tests. */
void morsi_erosion(float *y, float *x, int w, int h, int *e);
void morsi_dilation(float *y, float *x, int w, int h, int *e);
void morsi_median(float *y, float *x, int w, int h, int *e);
void morsi_rank(float *y, float *x, int w, int h, int *e);
void morsi_opening(float *y, float *x, int w, int h, int *e);
void morsi_closing(float *y, float *x, int w, int h, int *e);
void morsi_gradient(float *y, float *x, int w, int h, int *e);
void morsi_igradient(float *y, float *x, int w, int h, int *e);
void morsi_egradient(float *y, float *x, int w, int h, int *e);
void morsi_laplacian(float *y, float *x, int w, int h, int *e);
void morsi_enhance(float *y, float *x, int w, int h, int *e);
void morsi_blur(float *y, float *x, int w, int h, int *e);
void morsi_oscillation(float *y, float *x, int w, int h, int *e);
void morsi_tophat(float *y, float *x, int w, int h, int *e);
void morsi_bothat(float *y, float *x, int w, int h, int *e);
void morsi_iblur(float *y, float *x, int w, int h, int *e);
void morsi_eblur(float *y, float *x, int w, int h, int *e);
void morsi_cblur(float *y, float *x, int w, int h, int *e);
/* src/morsi.c:278-310: NULL outputs are skipped. */
void morsi_all(float *o_ero, float *o_dil, float *o_ope, float *o_clo,
		float *o_grad, float *o_igrad, float *o_egrad,
		float *o_lap, float *o_enh, float *o_str,
		float *o_top, float *o_bot, float *x, int w, int h, int *e);
int *build_disk(float radius);         /* src/morsi.c:313 (non-static there) */
/* src/morsi.c:478: the CLI entry kept callable for the `im` multi-call
 * binary (src/im.c:4-6); defined by the morsi host program, not the library. */
int main_morsi(int c, char **v);

#ifdef __cplusplus
}
#endif
#endif /* MORSI_CUDA_H */
