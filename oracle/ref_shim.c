/*
 * ref_shim.c -- builds the UNMODIFIED reference compute functions into a
 * shared object / timer, the way the reference itself reuses them
 * (src/ftr/webcam/corrview.c:17-18: OMIT_MORSI_MAIN + #include "morsi.c").
 *
 * TEST INFRASTRUCTURE ONLY (see oracle/morsi_oracle.c header).  No reference
 * source is copied: the compiler reads it from $(REF)/src at build time and
 * the outputs go to oracle/_ref/ (git-ignored, shipped to the GPU box).
 */
#define OMIT_MORSI_MAIN
#include REF_MORSI_C
#include <string.h>
#include <time.h>

typedef void (*ref_op_t)(float*, float*, int, int, int*);

static ref_op_t ref_ops[18] = {
	morsi_erosion, morsi_dilation, morsi_median, morsi_rank, morsi_opening,
	morsi_closing, morsi_gradient, morsi_igradient, morsi_egradient,
	morsi_laplacian, morsi_enhance, morsi_blur, morsi_oscillation,
	morsi_tophat, morsi_bothat, morsi_iblur, morsi_eblur, morsi_cblur
};

/* op index = dispatcher order of src/morsi.c:510-527 */
int morsi_ref_apply(int op, int *e, float *x, float *y, int w, int h)
{
	if (op < 0 || op >= 18) return 1;
	ref_ops[op](y, x, w, h, e);
	return 0;
}

/* kind: 0 disk 1 dysk 2 hrec 3 vrec 4 drec 5 Drec; returns ints written,
 * 0 for NULL, -1 if cap too small */
int morsi_ref_build(int kind, float radius, int *out, int cap)
{
	int *e = NULL;
	switch (kind) {
	case 0: e = build_disk(radius); break;
	case 1: e = build_dysk(radius); break;
	case 2: e = build_hrec(radius); break;
	case 3: e = build_vrec(radius); break;
	case 4: e = build_drec(radius); break;
	case 5: e = build_Drec(radius); break;
	}
	if (!e) return 0;
	int n = 2*e[0] + 4;
	if (n > cap) { free(e); return -1; }
	memcpy(out, e, n * sizeof*e);
	free(e);
	return n;
}

/* seconds spent inside the reference function only (no I/O) */
double morsi_ref_time(int op, int *e, float *x, float *y, int w, int h)
{
	struct timespec t0, t1;
	clock_gettime(CLOCK_MONOTONIC, &t0);
	ref_ops[op](y, x, w, h, e);
	clock_gettime(CLOCK_MONOTONIC, &t1);
	return (t1.tv_sec - t0.tv_sec) + 1e-9 * (t1.tv_nsec - t0.tv_nsec);
}
