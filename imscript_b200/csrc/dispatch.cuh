// dispatch.cuh -- internal types shared by the translation units of
// libmorsi_cuda (not part of the C ABI).
#pragma once
#include <cuda_runtime.h>
#include <map>
#include <mutex>
#include <string>

#include "../../include/morsi_cuda.h"
#include "common.cuh"
#include "element.h"

#define MORSI_WS_SLOTS 10  // 0-3: kernel temporaries, 4-5: host-pipeline staging, 6-7: pitched copies (k_disk), 8-9: morsi_all results
#define MORSI_LANES 5   // lane 0: the *_device entry points; 1..3: host-pointer pipeline; 4: edge strips of a sharded band
#define MORSI_LANE_SHARD_SIDE 4

// compiled form of a row-run element (k_rowrun.cu)
struct RowRunPlan {
	int ok;
	int reach;
	int hw[2 * MORSI_MAX_REACH_ROWRUN + 1];
};

struct DevElement {
	int n;
	int2 *d_offs;                 // effective offsets, element order, on the device
	int canonical3x3;             // the list IS the reference's cross (1) / square = disk2 (2) literal, in its order
	int *d_tile_offs;             // k_tiled: (dy-ymin)*pw + (dx-xmin) per element, pw = 128 + xmax - xmin
	morsi_element_info info;
	RowRunPlan rowrun;
};

struct MorsiCtx {
	int device = 0;
	int sm_count = 148;
	int smem_optin = 0;
	cudaStream_t stream = nullptr;
	int *d_flag = nullptr;        // [0]: "saw -0.0" word written by the fast kernels
	int *h_flag = nullptr;        // pinned mirror
	cudaStream_t lane_stream[MORSI_LANES] = {};
	void *ws[MORSI_LANES][MORSI_WS_SLOTS] = {};
	size_t ws_bytes[MORSI_LANES][MORSI_WS_SLOTS] = {};
	std::map<std::string, DevElement> elements;
	// mu guards `elements` and the workspace table; host_mu serialises the
	// host-pointer entry points on this device (they own lanes 1..3), so two
	// host threads may call into one device.  The *_device entry points share
	// lane 0's temporaries: one stream at a time per device (documented).
	std::mutex mu, host_mu;
};

// one dispatch: `planes` planes, each a band of a w x h plane
struct MorsiJob {
	int op;
	int w, h, planes;
	const float *x; int x_row0, x_rows; long long x_pstride;
	float *y;       int y_row0, y_rows; long long y_pstride;
	cudaStream_t stream;
	int lane;          // workspace / flag lane
	int pitch = 0;     // floats between rows of x and y; 0: w (only k_disk's pitched copies use another)
};

// how an operation decomposes into min/max passes (src/morsi.c:141-275)
struct OpPlan {
	int stages;        // 1, or 2 when a temporary erosion/dilation feeds the final pass
	int t_min, t_max;  // stage 1 produces erosion (t_min) and/or dilation (t_max) of x
	int a_from;        // final erosion-side pass reads: 0 nothing, 1 x, 2 t_min, 3 t_max
	int b_from;        // final dilation-side pass reads: same codes
	int epi;           // Epi
	int special;       // 0 none, 1 median, 2 rank
};
OpPlan morsi_op_plan(int op);

int morsi_set_error(int code, const char *fmt, ...);
void morsi_count_launch(int n);
int morsi_path(void);
int morsi_ctx_get(int device, MorsiCtx **out);
int morsi_ctx_current(MorsiCtx **out);
int morsi_ws_get(MorsiCtx *c, int lane, int slot, size_t bytes, void **out);
int morsi_element_get(MorsiCtx *c, const int *e, const DevElement **out);
void morsi_element_compile(MorsiCtx *c, DevElement *d);
int morsi_dispatch(MorsiCtx *c, const int *e, const MorsiJob &job);
int morsi_optin_smem(const void *kernel, int device, int bytes);
// fast kernel families: set *handled = 1 when they took the job
int morsi_run_small(MorsiCtx *c, const DevElement *de, const MorsiJob &job, int *flag, int *handled);
int morsi_run_disk(MorsiCtx *c, const DevElement *de, const MorsiJob &job, int *flag, int *handled);
int morsi_run_runs(MorsiCtx *c, const DevElement *de, const MorsiJob &job, int *flag, int *handled);
int morsi_run_median(MorsiCtx *c, const DevElement *de, const MorsiJob &job, int *flag, int *handled);
int morsi_run_median3(MorsiCtx *c, const DevElement *de, const MorsiJob &job, int *flag, int *handled);
int morsi_run_tiled(MorsiCtx *c, const DevElement *de, const MorsiJob &job, int *flag, int *handled);
int morsi_run_line(MorsiCtx *c, const DevElement *de, const MorsiJob &job, int *flag, int *handled);

#define MORSI_CU(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) \
	return morsi_set_error(e_ == cudaErrorMemoryAllocation ? MORSI_ERR_OOM : MORSI_ERR_CUDA, \
		"%s: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); } while (0)
