/* A library-style caller of the reference's morsi API, written the way
 * src/ftr/webcam/corrview.c:17-18,37-38,71-72 uses it -- build_disk(5.1),
 * morsi_bothat(y, x, w, h, E), morsi_enhance(...), morsi_all(...) -- but
 * LINKED against libmorsi_compat / libmorsi_cuda instead of #including
 * morsi.c.  Reads a raw float32 image, writes the raw results; the pytest
 * compares them with the compiled reference (oracle/_ref).
 *   compat_caller W H in.raw out_prefix
 */
#include <stdio.h>
#include <stdlib.h>

/* the reference's own prototypes (src/morsi.c): nothing of morsi_cuda.h is needed */
int *build_disk(float radius);
void morsi_bothat(float *y, float *x, int w, int h, int *e);
void morsi_enhance(float *y, float *x, int w, int h, int *e);
void morsi_median(float *y, float *x, int w, int h, int *e);
void morsi_all(float *o_ero, float *o_dil, float *o_ope, float *o_clo,
		float *o_grad, float *o_igrad, float *o_egrad,
		float *o_lap, float *o_enh, float *o_str,
		float *o_top, float *o_bot, float *x, int w, int h, int *e);

static void dump(const char *prefix, const char *name, const float *y, int n)
{
	char fn[1024];
	snprintf(fn, sizeof fn, "%s%s.raw", prefix, name);
	FILE *f = fopen(fn, "wb");
	if (!f || fwrite(y, sizeof *y, n, f) != (size_t)n) exit(3);
	fclose(f);
}

int main(int c, char **v)
{
	if (c != 5) return 2;
	int w = atoi(v[1]), h = atoi(v[2]), n = w * h;
	float *x = malloc(n * sizeof *x), *y = malloc(n * sizeof *y);
	FILE *f = fopen(v[3], "rb");
	if (!f || fread(x, sizeof *x, n, f) != (size_t)n) return 3;
	fclose(f);
	int *E = build_disk(5.1);                          /* corrview.c:37 */
	morsi_bothat(y, x, w, h, E);                       /* corrview.c:38 */
	dump(v[4], "bothat", y, n);
	morsi_enhance(y, x, w, h, E);                      /* corrview.c:72 (commented variants) */
	dump(v[4], "enhance", y, n);
	morsi_median(y, x, w, h, E);
	dump(v[4], "median", y, n);
	float *o[12];
	for (int k = 0; k < 12; k++) o[k] = (k == 1 || k == 6) ? NULL : malloc(n * sizeof(float));
	morsi_all(o[0], o[1], o[2], o[3], o[4], o[5], o[6], o[7], o[8], o[9], o[10], o[11], x, w, h, E);
	static const char *const names[12] = {"all_erosion", "all_dilation", "all_opening", "all_closing", "all_gradient",
		"all_igradient", "all_egradient", "all_laplacian", "all_enhance", "all_oscillation", "all_tophat", "all_bothat"};
	for (int k = 0; k < 12; k++) if (o[k]) dump(v[4], names[k], o[k], n);
	free(E);
	return 0;
}
