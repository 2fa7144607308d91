// host_apply.cu -- morsi_cuda_apply(): the host-pointer entry point that
// replaces the channel loop of src/morsi.c:539-543.
//
// The planes are cut into row-band chunks; each chunk goes host->device,
// through the kernels (band form, halo rows re-sent with the chunk) and back
// device->host on one of three pipeline lanes, so the two copy engines and the
// SMs overlap.  With MORSI_CUDA_DEVICES=N the chunk list is dealt to N devices
// (whole planes when there are enough of them, otherwise row bands); the halo
// rows come straight from the caller's host buffer, so no peer traffic is
// needed on this entry point.
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <thread>
#include <vector>

#include "dispatch.cuh"

struct Chunk { int plane, r0, r1; };

// A caller's malloc'd (pageable) image: cudaMemcpyAsync would stage it through the
// driver synchronously and the three-lane overlap would be lost.  Page-lock it
// for the duration of the call instead (the reference CLI's buffers come from
// iio's malloc, src/morsi.c:531-532).  Already pinned / registered memory and
// small images are left alone; a failed registration only costs the overlap.
struct HostPin {
	void *p = nullptr;
	HostPin(const void *ptr, size_t bytes)
	{
		static const bool off = getenv("MORSI_CUDA_HOST_REGISTER") && !strcmp(getenv("MORSI_CUDA_HOST_REGISTER"), "0");
		if (off || bytes < (8u << 20)) return;
		cudaPointerAttributes at;
		if (cudaPointerGetAttributes(&at, ptr) != cudaSuccess) { cudaGetLastError(); return; }
		if (at.type != cudaMemoryTypeUnregistered) return;
		if (cudaHostRegister((void *)ptr, bytes, cudaHostRegisterPortable) == cudaSuccess) p = (void *)ptr;
		else cudaGetLastError();
	}
	~HostPin() { if (p) cudaHostUnregister(p); }
};

static int run_chunks(int device, int op, const int *e, const float *x, float *y,
		int w, int h, const std::vector<Chunk> &chunks, int up, int down)
{
	MorsiCtx *c;
	int rc = morsi_ctx_get(device, &c);
	if (rc) return rc;
	std::lock_guard<std::mutex> host_lk(c->host_mu);
	size_t max_in = 0, max_out = 0;
	for (const Chunk &k : chunks) {
		int i0 = std::max(0, k.r0 - up), i1 = std::min(h, k.r1 + down);
		max_in = std::max(max_in, (size_t)(i1 - i0) * w * sizeof(float));
		max_out = std::max(max_out, (size_t)(k.r1 - k.r0) * w * sizeof(float));
	}
	const int lanes = (int)std::min<size_t>(3, chunks.size());
	float *d_in[3], *d_out[3];
	for (int l = 0; l < lanes; l++) {
		void *p;
		if ((rc = morsi_ws_get(c, 1 + l, 4, max_in, &p))) return rc;
		d_in[l] = (float *)p;
		if ((rc = morsi_ws_get(c, 1 + l, 5, max_out, &p))) return rc;
		d_out[l] = (float *)p;
	}
	for (size_t t = 0; t < chunks.size(); t++) {
		const Chunk &k = chunks[t];
		const int l = (int)(t % lanes);
		cudaStream_t s = c->lane_stream[1 + l];
		const int i0 = std::max(0, k.r0 - up), i1 = std::min(h, k.r1 + down);
		const float *src = x + (size_t)k.plane * w * h + (size_t)i0 * w;
		float *dst = y + (size_t)k.plane * w * h + (size_t)k.r0 * w;
		MORSI_CU(cudaMemcpyAsync(d_in[l], src, (size_t)(i1 - i0) * w * sizeof(float),
				cudaMemcpyHostToDevice, s));
		MorsiJob job;
		job.op = op; job.w = w; job.h = h; job.planes = 1; job.lane = 1 + l;
		job.x = d_in[l]; job.x_row0 = i0; job.x_rows = i1 - i0; job.x_pstride = (long long)w * (i1 - i0);
		job.y = d_out[l]; job.y_row0 = k.r0; job.y_rows = k.r1 - k.r0; job.y_pstride = (long long)w * (k.r1 - k.r0);
		job.stream = s;
		if ((rc = morsi_dispatch(c, e, job))) return rc;
		MORSI_CU(cudaMemcpyAsync(dst, d_out[l], (size_t)(k.r1 - k.r0) * w * sizeof(float),
				cudaMemcpyDeviceToHost, s));
	}
	for (int l = 0; l < lanes; l++) MORSI_CU(cudaStreamSynchronize(c->lane_stream[1 + l]));
	return MORSI_OK;
}

extern "C" int morsi_cuda_apply(int op, const int *e, const float *x, float *y,
		int w, int h, int planes)
{
	if (op < 0 || op >= MORSI_OP_COUNT) return morsi_set_error(MORSI_ERR_INVALID, "unknown operation %d", op);
	if (!e || e[0] < 0) return morsi_set_error(MORSI_ERR_INVALID, "bad structuring element");
	if (!x || !y) return morsi_set_error(MORSI_ERR_INVALID, "NULL image pointer");
	if (w <= 0 || h <= 0 || planes <= 0)
		return morsi_set_error(MORSI_ERR_INVALID, "non-positive image size %dx%dx%d", w, h, planes);
	int navail = morsi_cuda_device_count();
	if (navail <= 0) return morsi_set_error(MORSI_ERR_NO_DEVICE, "no CUDA device visible");
	int ndev = 1;
	if (const char *s = getenv("MORSI_CUDA_DEVICES")) ndev = atoi(s);
	if (ndev < 1) ndev = 1;
	if (ndev > navail) ndev = navail;

	int up = 0, down = 0;
	int rc = morsi_cuda_halo_rows(op, e, &up, &down);
	if (rc) return rc;
	HostPin pin_x(x, (size_t)w * h * planes * sizeof(float)), pin_y(y, (size_t)w * h * planes * sizeof(float));

	// chunk height: ~32 MiB of input per chunk, at least 8x the halo
	long long chunk_mb = 32;                                 // MiB of input per chunk (MORSI_CUDA_CHUNK_MB; measured: profiles/r2_chunk_sweep.txt)
	if (const char *s = getenv("MORSI_CUDA_CHUNK_MB")) chunk_mb = std::max(1, atoi(s));
	long long target = (chunk_mb << 20) / ((long long)w * 4);
	int band = (int)std::max<long long>(std::max(64, 8 * (up + down)), target);
	if (const char *s = getenv("MORSI_CUDA_CHUNK_ROWS")) band = std::max(1, atoi(s));

	// deal the work: whole planes per device when possible, else row bands
	std::vector<std::vector<Chunk>> per_dev(ndev);
	if (planes >= ndev) {
		for (int p = 0; p < planes; p++) {
			int d = (int)((long long)p * ndev / planes);
			for (int r = 0; r < h; r += band)
				per_dev[d].push_back({p, r, std::min(h, r + band)});
		}
	} else {
		for (int p = 0; p < planes; p++)
			for (int d = 0; d < ndev; d++) {
				int b0 = (int)((long long)h * d / ndev), b1 = (int)((long long)h * (d + 1) / ndev);
				for (int r = b0; r < b1; r += band)
					per_dev[d].push_back({p, r, std::min(b1, r + band)});
			}
	}
	// Pipeline ramp: the first upload and the last download of a call overlap nothing, so the
	// first and the last chunk of every device are cut short (a quarter, then the rest).
	for (auto &list : per_dev) {
		const int min_rows = std::max(32, 4 * (up + down));
		if (list.size() >= 2) {
			Chunk f = list.front();
			const int q = (f.r1 - f.r0) / 4;
			if (q >= min_rows) {
				list.front().r0 = f.r0 + q;
				list.insert(list.begin(), Chunk{f.plane, f.r0, f.r0 + q});
			}
			Chunk l = list.back();
			const int ql = (l.r1 - l.r0) / 4;
			if (ql >= min_rows) {
				list.back().r1 = l.r1 - ql;
				list.push_back(Chunk{l.plane, l.r1 - ql, l.r1});
			}
		}
	}
	if (ndev == 1) {
		MorsiCtx *c; rc = morsi_ctx_current(&c); if (rc) return rc;
		return run_chunks(c->device, op, e, x, y, w, h, per_dev[0], up, down);
	}
	std::vector<int> rcs(ndev, MORSI_OK);
	std::vector<std::thread> th;
	for (int d = 0; d < ndev; d++)
		th.emplace_back([&, d] {
			if (!per_dev[d].empty())
				rcs[d] = run_chunks(d, op, e, x, y, w, h, per_dev[d], up, down);
		});
	for (auto &t : th) t.join();
	for (int d = 0; d < ndev; d++)
		if (rcs[d]) return morsi_set_error(rcs[d], "device %d failed", d);
	return MORSI_OK;
}

