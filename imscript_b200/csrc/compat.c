/*
 * compat.c -- libmorsi_compat: the reference's own function signatures
 * (src/morsi.c:56-310,313) on top of the libmorsi_cuda C ABI, so that
 * library-style callers such as src/ftr/webcam/corrview.c:37-38,71-72
 * relink unchanged.  Host pointers, synchronous.  Failures follow the
 * reference's fail() convention (src/fail.c:65-81): one line on stderr and
 * exit(-1); there is no CPU fallback.
 */
#include <stdio.h>
#include <stdlib.h>

#include "../../include/morsi_cuda.h"

static void run(int op, float *y, float *x, int w, int h, int *e)
{
	int rc = morsi_cuda_apply(op, e, x, y, w, h, 1);
	if (rc != MORSI_OK) {
		fprintf(stderr, "FAIL(\"morsi_%s\"): libmorsi_cuda: %s: %s\n",
				morsi_operation_name(op), morsi_cuda_strerror(rc),
				morsi_cuda_last_error());
		exit(-1);
	}
}

#define WRAP(name, OP) \
	void morsi_##name(float *y, float *x, int w, int h, int *e) { run(OP, y, x, w, h, e); }
WRAP(erosion, MORSI_EROSION)
WRAP(dilation, MORSI_DILATION)
WRAP(median, MORSI_MEDIAN)
WRAP(rank, MORSI_RANK)
WRAP(opening, MORSI_OPENING)
WRAP(closing, MORSI_CLOSING)
WRAP(gradient, MORSI_GRADIENT)
WRAP(igradient, MORSI_IGRADIENT)
WRAP(egradient, MORSI_EGRADIENT)
WRAP(laplacian, MORSI_LAPLACIAN)
WRAP(enhance, MORSI_ENHANCE)
WRAP(blur, MORSI_BLUR)
WRAP(oscillation, MORSI_OSCILLATION)
WRAP(tophat, MORSI_TOPHAT)
WRAP(bothat, MORSI_BOTHAT)
WRAP(iblur, MORSI_IBLUR)
WRAP(eblur, MORSI_EBLUR)
WRAP(cblur, MORSI_CBLUR)

/* src/morsi.c:278-310: every non-NULL output is filled; `o_str` is the
 * oscillation.  Each output equals the single-operation result (the
 * reference's shared temporaries do not change any value); the input is
 * uploaded once (morsi_cuda_apply_all). */
void morsi_all(float *o_ero, float *o_dil, float *o_ope, float *o_clo,
		float *o_grad, float *o_igrad, float *o_egrad,
		float *o_lap, float *o_enh, float *o_str,
		float *o_top, float *o_bot, float *x, int w, int h, int *e)
{
	float *const out[12] = {o_ero, o_dil, o_ope, o_clo, o_grad, o_igrad, o_egrad, o_lap, o_enh, o_str, o_top, o_bot};
	int rc = morsi_cuda_apply_all(e, x, out, w, h, 1);
	if (rc != MORSI_OK) {
		fprintf(stderr, "FAIL(\"morsi_all\"): libmorsi_cuda: %s: %s\n", morsi_cuda_strerror(rc), morsi_cuda_last_error());
		exit(-1);
	}
}

int *build_disk(float radius) { return morsi_build_disk(radius); }
