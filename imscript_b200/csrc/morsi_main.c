/*
 * morsi_main.c -- the `morsi` command line on top of libmorsi_cuda.
 *
 * Drop-in for the reference front-end (src/morsi.c:419-558): same argv
 * grammar `morsi ELEMENT OPERATION [in [out]]`, same messages and exit codes,
 * same help family (src/help_stuff.c:23-34), image I/O through the
 * reference's own iio (linked from the reference tree, not rewritten).
 * The compute step is the only thing that changed: one call into the CUDA
 * library instead of the per-channel function-pointer loop (:539-543).
 * No GPU => error exit; there is no CPU fallback.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../../include/morsi_cuda.h"

/* the two iio entry points morsi uses (src/iio.h:42,182) */
float *iio_read_image_float_split(const char *fname, int *w, int *h, int *pd);
void iio_write_image_float_split(char *fname, float *x, int w, int h, int pd);

static const char usage_line[] =
"usage:\n\tmorsi {cross|square|disk5...} {erosion|dilation...} [in [out]]";
static const char oneliner[] = "simple gray-scale morphology";
static const char version_text[] = "morsi 1.0\n\nWritten by mnhrdt";

/* The help text is part of the CLI contract (doc/man/man1/morsi.1 is
 * generated from it); it is emitted from a table so that element and
 * operation names stay next to their descriptions. */
static const char *const element_help[][2] = {
	{"cross", "a cross of 5 pixels"},
	{"square", "a square of 3x3 pixels"},
	{"diskR", "a discrete disk of radius R (e.g., disk3.1)"},
	{"dyskR", "the boundary of a discrete disk of radius R"},
	{"hrecW", "a rectangle of size Wx1"},
	{"vrecH", "a rectangle of size 1xH"},
	{"drecX", "a 45-degree diagonal of X-pixels"},
	{"DrecX", "a -45-degree diagonal of X-pixels"},
};
static const char *const operation_help[][2] = {
	{"erosion", "min of neighboring pixels (smooth and darken)"},
	{"dilation", "max of neighboring pixels (smooth and brighten)"},
	{"median", "median of neighboring pixels (smooth)"},
	{"opening", "erosion and dilation (remove small dark spots)"},
	{"closing", "dilation and erosion (remove small bright spots)"},
	{"gradient", "dilation minus erosion (find centered boundaries)"},
	{"igradient", "image minus erosion (inner boundaries)"},
	{"egradient", "dilation minus image (outer boundaries)"},
	{"laplacian", "difference between outer and inner grads (signed boundaries)"},
	{"enhance", "image minus laplacian (sharpen boundaries and details)"},
	{"blur", "image plus laplacian (centered smoothing, removes deatils)"},
	{"tophat", "image minus opening (select bright blobs)"},
	{"bothat", "closing minus image (select dark blobs)"},
	{"oscillation", "closing minus opening (measure local oscillation)"},
	{"iblur", "average of image and its erosion (blur towards dark)"},
	{"eblur", "average of image and its dilation (blur towards light)"},
	{"cblur", "average of iblur and eblur (detail-preserving smoothing)"},
};

static void print_long_help(void)
{
	puts("Morsi applies a gray-scale morphological operator to an image.\n"
	     "\n"
	     "You have to specify a structuring element and a morphological operation.\n"
	     "There are no defaults.\n"
	     "\n"
	     "Usage: morsi ELEMENT OPERATION in out\n"
	     "   or: morsi ELEMENT OPERATION in > out\n"
	     "   or: cat in | morsi ELEMENT OPERATION > out\n"
	     "\n"
	     "Elements:");
	for (size_t i = 0; i < sizeof element_help / sizeof *element_help; i++)
		printf(" %-12s %s\n", element_help[i][0], element_help[i][1]);
	puts("\nOperations:");
	for (size_t i = 0; i < sizeof operation_help / sizeof *operation_help; i++)
		printf(" %-12s %s\n", operation_help[i][0], operation_help[i][1]);
	puts("\n"
	     "Examples:\n"
	     " morsi cross erosion i.png o.png    Erode by a \"cross\" structuring element\n"
	     " morsi disk4.2 median i.png o.png   Median filtering of radius 4.2\n"
	     "\n"
	     "Report bugs to <enric.meinhardt@ens-paris-saclay.fr>.");
}

/* src/help_stuff.c:5-21: the --man family shells out to help2man */
static int run_help2man(int raw)
{
	char cmd[0x200];
	snprintf(cmd, sizeof cmd, "help2man -N -S imscript -n \"%s\" %smorsi %s",
			oneliner, raw > 0 ? "" : "./", abs(raw) > 1 ? "" : "|man -l -");
	return system(cmd);
}

/* src/help_stuff.c:23-34; only consulted when argc == 2 (src/morsi.c:481) */
static void handle_help_argument(const char *s)
{
	if (!s || !strcmp(s, "-h"))   { puts(usage_line); exit(0); }
	if (!strcmp(s, "-?"))         { puts(oneliner); exit(0); }
	if (!strcmp(s, "--help"))     { print_long_help(); exit(0); }
	if (!strcmp(s, "--version"))  { puts(version_text); exit(0); }
	if (!strcmp(s, "--man"))      exit(run_help2man(1));
	if (!strcmp(s, "--manraw"))   exit(run_help2man(2));
	if (!strcmp(s, "--man-x"))    exit(run_help2man(-1));
	if (!strcmp(s, "--manraw-x")) exit(run_help2man(-2));
	if (!strcmp(s, "--help-oneliner")) puts(oneliner);   /* falls through to the usage error */
}

int main_morsi(int c, char **v)
{
	if (c == 2) handle_help_argument(v[1]);

	if (c != 3 && c != 4 && c != 5) {
		fprintf(stderr, "usage:\n\t%s element operation [in [out]]\n", *v);
		return 1;
	}
	char *filename_in  = c > 3 ? v[3] : "-";
	char *filename_out = c > 4 ? v[4] : "-";

	int *element = NULL;
	if (morsi_element_parse(v[1], &element) != MORSI_OK) {
		fprintf(stderr, "elements = cross, square ...\n");
		return 1;
	}
	int op = morsi_operation_parse(v[2]);
	if (op < 0) {
		fprintf(stderr, "operations = erosion, dilation, opening...\n");
		return 1;
	}

	int w, h, pd;
	float *x = iio_read_image_float_split(filename_in, &w, &h, &pd);
	float *y = malloc((size_t)w * h * pd * sizeof *y);
	if (!y) {
		fprintf(stderr, "FAIL(\"%s\"): out of memory\n", *v);
		exit(-1);
	}

	int rc = morsi_cuda_apply(op, element, x, y, w, h, pd);
	if (rc != MORSI_OK) {
		fprintf(stderr, "FAIL(\"%s\"): libmorsi_cuda: %s: %s\n", *v,
				morsi_cuda_strerror(rc), morsi_cuda_last_error());
		exit(-1);
	}

	iio_write_image_float_split(filename_out, y, w, h, pd);

	free(x);
	free(y);
	return 0;
}

#ifndef HIDE_ALL_MAINS
int main(int c, char **v) { return main_morsi(c, v); }
#endif
