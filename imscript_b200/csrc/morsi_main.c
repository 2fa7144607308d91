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
#define _FILE_OFFSET_BITS 64
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/types.h>

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

static void die(const char *prog, int rc)
{
	fprintf(stderr, "FAIL(\"%s\"): libmorsi_cuda: %s: %s\n", prog,
			morsi_cuda_strerror(rc), morsi_cuda_last_error());
	exit(-1);
}

/* ---- extension 1: `morsi ELEMENT all in out-%s.ext` ------------------------
 * The twelve results of morsi_all (src/morsi.c:278-310) from one read of the
 * input and one upload, written to the output name with %s replaced by the
 * operation's name (the tutorial, doc/tutorial/i.html:213-225, runs morsi 13
 * times on the same image for this).  "all" is not a name of the reference's
 * operation table, so no reference invocation changes meaning. */
static int run_all(const char *prog, int *element, const char *in, const char *out)
{
	static const char *const names[12] = {"erosion", "dilation", "opening", "closing", "gradient", "igradient",
		"egradient", "laplacian", "enhance", "oscillation", "tophat", "bothat"};
	const char *pct = strstr(out, "%s");
	if (!pct || strstr(pct + 2, "%")) {
		fprintf(stderr, "usage:\n\t%s element all in out-%%s.ext\n", prog);
		return 1;
	}
	int w, h, pd;
	float *x = iio_read_image_float_split(in, &w, &h, &pd);
	size_t n = (size_t)w * h * pd;
	float *y[12];
	for (int k = 0; k < 12; k++)
		if (morsi_cuda_host_alloc((void **)&y[k], n * sizeof(float)) != MORSI_OK) {
			fprintf(stderr, "FAIL(\"%s\"): out of memory\n", prog);
			exit(-1);
		}
	int rc = morsi_cuda_apply_all(element, x, y, w, h, pd);
	if (rc != MORSI_OK) die(prog, rc);
	for (int k = 0; k < 12; k++) {
		char name[0x400];
		snprintf(name, sizeof name, "%.*s%s%s", (int)(pct - out), out, names[k], pct + 2);
		iio_write_image_float_split(name, y[k], w, h, pd);
		morsi_cuda_host_free(y[k]);
	}
	free(x);
	return 0;
}

/* ---- extension 2: MORSI_CUDA_STREAM=1, float32 2-D .npy in and out ----------
 * The image streams from the input file through the device to the output file
 * in row bands (morsi_cuda_apply_stream): the host never holds it, so planes
 * larger than memory -- and than the int-sized buffers of src/iio.c:3759,4073 --
 * go through.  Anything else falls back to the ordinary path. */
struct npy_io { FILE *in, *out; long in_off, out_off; int w; };
static int npy_read_rows(void *u, int plane, int row0, int nrows, float *dst)
{
	struct npy_io *io = u;
	(void)plane;
	if (fseeko(io->in, io->in_off + (off_t)row0 * io->w * 4, SEEK_SET)) return 1;
	return fread(dst, 4, (size_t)nrows * io->w, io->in) != (size_t)nrows * io->w;
}
static int npy_write_rows(void *u, int plane, int row0, int nrows, const float *src)
{
	struct npy_io *io = u;
	(void)plane;
	if (fseeko(io->out, io->out_off + (off_t)row0 * io->w * 4, SEEK_SET)) return 1;
	return fwrite(src, 4, (size_t)nrows * io->w, io->out) != (size_t)nrows * io->w;
}
/* 0: streamed; -1: not applicable (caller takes the ordinary path) */
static int try_stream_npy(const char *prog, int op, int *element, const char *in, const char *out)
{
	const char *s = getenv("MORSI_CUDA_STREAM");
	size_t li = strlen(in), lo = strlen(out);
	if (!s || strcmp(s, "1") || li < 5 || lo < 5 || strcmp(in + li - 4, ".npy") || strcmp(out + lo - 4, ".npy"))
		return -1;
	FILE *f = fopen(in, "rb");
	if (!f) return -1;
	unsigned char hd[10];
	char dict[4096];
	if (fread(hd, 1, 10, f) != 10 || memcmp(hd, "\x93NUMPY", 6) || hd[6] != 1) { fclose(f); return -1; }
	unsigned hlen = hd[8] | (hd[9] << 8);
	if (hlen >= sizeof dict || fread(dict, 1, hlen, f) != hlen) { fclose(f); return -1; }
	dict[hlen] = 0;
	long hh = 0, ww = 0, dd = 1;
	const char *sh = strstr(dict, "'shape'");
	if (!strstr(dict, "'<f4'") || !strstr(dict, "False") || !sh) { fclose(f); return -1; }
	{
		/* (h, w) or (h, w, 1): one gray plane, rows contiguous */
		const char *p0 = strchr(sh, '('), *p1 = p0 ? strchr(p0, ')') : NULL;
		int nd = p0 && p1 ? sscanf(p0, "(%ld, %ld, %ld", &hh, &ww, &dd) : 0;
		if (nd < 2 || (nd == 3 && dd != 1) || hh <= 0 || ww <= 0 || hh > 0x7fffffff || ww > 0x7fffffff) { fclose(f); return -1; }
		int commas = 0;
		for (const char *q = p0; q < p1; q++) commas += *q == ',';
		if (commas > 3) { fclose(f); return -1; }
	}
	FILE *g = fopen(out, "wb");
	if (!g) { fclose(f); return -1; }
	/* same type, same shape: the output header is the input's */
	const int total = 10 + (int)hlen;
	fwrite(hd, 1, 10, g);
	fwrite(dict, 1, hlen, g);
	struct npy_io io = {f, g, 10 + (long)hlen, total, (int)ww};
	int rc = morsi_cuda_apply_stream(op, element, (int)ww, (int)hh, 1, npy_read_rows, npy_write_rows, &io);
	fclose(f);
	if (fclose(g)) rc = rc ? rc : MORSI_ERR_INVALID;
	if (rc != MORSI_OK) die(prog, rc);
	return 0;
}

/* the iio "vec" entry points (src/iio.h:33,176): pixel-interleaved samples */
float *iio_read_image_float_vec(const char *fname, int *w, int *h, int *pd);
void iio_write_image_float_vec(char *fname, float *x, int w, int h, int pd);
void iio_write_image_uint8_split(char *fname, unsigned char *x, int w, int h, int pd);

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
	if (!strcmp(v[2], "all") && c == 5)
		return run_all(*v, element, filename_in, filename_out);
	int op = morsi_operation_parse(v[2]);
	if (op < 0) {
		fprintf(stderr, "operations = erosion, dilation, opening...\n");
		return 1;
	}
	if (try_stream_npy(*v, op, element, filename_in, filename_out) == 0)
		return 0;

	/* extension 3: MORSI_CUDA_QEASY="black white" = `morsi E OP in | qeasy black white - out`
	 * (doc/tutorial/i.html:221-225, src/qeasy.c): the quantiser runs on the device as the last
	 * kernel of the chain and the 8-bit result comes back as bytes. */
	const char *qe = getenv("MORSI_CUDA_QEASY");
	float black, white;
	if (qe && sscanf(qe, "%f %f", &black, &white) == 2) {
		int w, h, pd;
		float *x = iio_read_image_float_split(filename_in, &w, &h, &pd);
		unsigned char *y8 = malloc((size_t)w * h * pd);
		if (!y8) { fprintf(stderr, "FAIL(\"%s\"): out of memory\n", *v); exit(-1); }
		morsi_quantizer q = {black, white, 1};
		const int *els[1] = {element};
		int rc = morsi_cuda_apply_chain(1, &op, els, x, y8, w, h, pd, &q);
		if (rc != MORSI_OK) die(*v, rc);
		iio_write_image_uint8_split(filename_out, y8, w, h, pd);
		free(x);
		free(y8);
		return 0;
	}

	/* The reference reads with iio_read_image_float_split (src/morsi.c:531): a
	 * "vec" read followed by a CPU pass that splits the pixels into planes
	 * (src/iio.c:5763-5771), and joins them again on the way out (:6525-6531).
	 * Here the vec image goes to the device as it is and is split / joined
	 * there (MORSI_CUDA_SPLIT=host restores the CPU passes). */
	int w, h, pd;
	const char *sp = getenv("MORSI_CUDA_SPLIT");
	const int split_on_host = sp && !strcmp(sp, "host");
	float *x = split_on_host ? iio_read_image_float_split(filename_in, &w, &h, &pd)
		: iio_read_image_float_vec(filename_in, &w, &h, &pd);
	float *y = NULL;
	if (morsi_cuda_host_alloc((void **)&y, (size_t)w * h * pd * sizeof *y) != MORSI_OK) {   /* page-locked: the download overlaps */
		fprintf(stderr, "FAIL(\"%s\"): out of memory\n", *v);
		exit(-1);
	}

	int rc = (split_on_host || pd == 1) ? morsi_cuda_apply(op, element, x, y, w, h, pd)
		: morsi_cuda_apply_interleaved(op, element, x, y, w, h, pd, MORSI_SAMPLE_F32);
	if (rc != MORSI_OK) die(*v, rc);

	if (split_on_host) iio_write_image_float_split(filename_out, y, w, h, pd);
	else iio_write_image_float_vec(filename_out, y, w, h, pd);

	free(x);
	morsi_cuda_host_free(y);
	return 0;
}

#ifndef HIDE_ALL_MAINS
int main(int c, char **v) { return main_morsi(c, v); }
#endif
