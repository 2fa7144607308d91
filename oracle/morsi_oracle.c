/*
 * morsi_oracle.c -- CPU restatement of the reference `morsi` algorithm.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product path (libmorsi_cuda, the
 * `morsi` CLI, imscript_b200/) may link, import or execute this file.  Only
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
 * reference legs use it, as the checker.
 *
 * Parity status: PINNED.  The reference has no tests or goldens for morsi
 * (SURVEY.md section 8c), but the reference itself compiles in the build
 * container (oracle/Makefile -> oracle/_ref/), and this restatement is checked
 * bit-for-bit against it in tests/test_oracle.py and against the committed
 * vectors under tests/golden/ that the reference binary produced.
 *
 * This is a restatement, not a copy: the reference evaluates libm fmin/fmax
 * and libc qsort; here their observable behaviour on x86-64 glibc 2.39 is
 * written out explicitly (last-wins ties, NaN operands ignored, stable sort).
 * Every function cites the reference lines it follows (paths relative to the
 * reference root).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

enum {
	OP_EROSION, OP_DILATION, OP_MEDIAN, OP_RANK, OP_OPENING, OP_CLOSING,
	OP_GRADIENT, OP_IGRADIENT, OP_EGRADIENT, OP_LAPLACIAN, OP_ENHANCE,
	OP_BLUR, OP_OSCILLATION, OP_TOPHAT, OP_BOTHAT, OP_IBLUR, OP_EBLUR,
	OP_CBLUR, OP_COUNT
}; /* dispatcher order, src/morsi.c:510-527 */

/* src/morsi.c:30-35 (getpixel_nan): out-of-image neighbours read as NaN.
 * 64-bit index so that crops of >2^31-sample images can be checked. */
static inline float pix_or_nan(const float *x, int w, int h, long i, long j)
{
	if (i < 0 || i >= w || j < 0 || j >= h)
		return NAN;
	return x[i + j * (long)w];
}

/* src/morsi.c:48-54,65: neighbour k of (i,j) is (i-e[2]+e[2k+4], j-e[3]+e[2k+5]) */
#define NEIGHBOUR(k) pix_or_nan(x, w, h, (long)i - e[2] + e[2*(k)+4], \
                                          (long)j - e[3] + e[2*(k)+5])

/* src/morsi.c:56-68.  fmin(a,v) on x86-64 glibc: NaN operand -> the other
 * one; otherwise minsd with v as the surviving operand on ties, i.e. the LAST
 * element-order occurrence of the minimum wins (observable for +0/-0 only).
 * `a` starts at +INF and is never NaN, so the rule reduces to one compare. */
static void o_erosion(float *y, const float *x, int w, int h, const int *e)
{
	for (int j = 0; j < h; j++)
	for (int i = 0; i < w; i++) {
		float a = INFINITY;
		for (int k = 0; k < e[0]; k++) {
			float v = NEIGHBOUR(k);
			if (v <= a) a = v;    /* false for NaN v */
		}
		y[(long)j * w + i] = a;
	}
}

/* src/morsi.c:70-82, same with fmax / -INF */
static void o_dilation(float *y, const float *x, int w, int h, const int *e)
{
	for (int j = 0; j < h; j++)
	for (int i = 0; i < w; i++) {
		float a = -INFINITY;
		for (int k = 0; k < e[0]; k++) {
			float v = NEIGHBOUR(k);
			if (v >= a) a = v;
		}
		y[(long)j * w + i] = a;
	}
}

/* Stable ascending sort under the order of src/morsi.c:84-89 (compare_floats:
 * (a>b)-(a<b), so +0 and -0 compare equal and keep their gather order, which
 * is what glibc 2.39's merge-sort qsort does).  Binary insertion is enough
 * for a checker. */
static void stable_sort(float *a, int n)
{
	for (int i = 1; i < n; i++) {
		float v = a[i];
		int p = i;
		while (p > 0 && a[p-1] > v) { a[p] = a[p-1]; p--; }
		a[p] = v;
	}
}

/* src/morsi.c:91-101.  Note the even case uses a[n/2] and a[n/2+1]. */
static float o_median_of(float *a, int n)
{
	if (n < 1) return NAN;
	if (n == 1) return a[0];
	if (n == 2) return (a[0] + a[1]) / 2;
	stable_sort(a, n);
	if (n % 2 == 0)
		return (a[n/2] + a[1 + n/2]) / 2;
	return a[n/2];
}

/* src/morsi.c:103-120: gather finite neighbours in element order */
static void o_median(float *y, const float *x, int w, int h, const int *e)
{
	float *a = malloc((e[0] > 0 ? e[0] : 1) * sizeof *a);
	for (int j = 0; j < h; j++)
	for (int i = 0; i < w; i++) {
		int cx = 0;
		for (int k = 0; k < e[0]; k++) {
			float v = NEIGHBOUR(k);
			if (isfinite(v)) a[cx++] = v;
		}
		y[(long)j * w + i] = o_median_of(a, cx);
	}
	free(a);
}

/* src/morsi.c:122-139: u is the pixel itself, strict <, finite neighbours */
static void o_rank(float *y, const float *x, int w, int h, const int *e)
{
	for (int j = 0; j < h; j++)
	for (int i = 0; i < w; i++) {
		int cx = 0;
		float u = pix_or_nan(x, w, h, i, j);
		for (int k = 0; k < e[0]; k++) {
			float v = NEIGHBOUR(k);
			if (isfinite(v)) cx += v < u;
		}
		y[(long)j * w + i] = cx;
	}
}

static float *tmp_plane(int w, int h)
{
	return malloc((size_t)w * h * sizeof(float));
}

/* composites: src/morsi.c:141-275, arithmetic exactly as written there
 * (float, left to right; cblur in double because its literals are double) */
int morsi_oracle_apply(int op, const int *e, const float *x, float *y,
		int w, int h)
{
	long n = (long)w * h;
	float *a = NULL, *b = NULL, *t = NULL;
	switch (op) {
	case OP_EROSION:  o_erosion(y, x, w, h, e); break;
	case OP_DILATION: o_dilation(y, x, w, h, e); break;
	case OP_MEDIAN:   o_median(y, x, w, h, e); break;
	case OP_RANK:     o_rank(y, x, w, h, e); break;
	case OP_OPENING:  /* :141-147 */
		t = tmp_plane(w, h);
		o_erosion(t, x, w, h, e); o_dilation(y, t, w, h, e); break;
	case OP_CLOSING:  /* :149-155 */
		t = tmp_plane(w, h);
		o_dilation(t, x, w, h, e); o_erosion(y, t, w, h, e); break;
	case OP_GRADIENT: /* :157-167 */
		a = tmp_plane(w, h); b = tmp_plane(w, h);
		o_erosion(a, x, w, h, e); o_dilation(b, x, w, h, e);
		for (long i = 0; i < n; i++) y[i] = b[i] - a[i];
		break;
	case OP_IGRADIENT: /* :169-176 */
		t = tmp_plane(w, h); o_erosion(t, x, w, h, e);
		for (long i = 0; i < n; i++) y[i] = x[i] - t[i];
		break;
	case OP_EGRADIENT: /* :178-185 */
		t = tmp_plane(w, h); o_dilation(t, x, w, h, e);
		for (long i = 0; i < n; i++) y[i] = t[i] - x[i];
		break;
	case OP_LAPLACIAN: case OP_ENHANCE: case OP_BLUR: /* :187-215 */
		a = tmp_plane(w, h); b = tmp_plane(w, h);
		o_erosion(a, x, w, h, e); o_dilation(b, x, w, h, e);
		for (long i = 0; i < n; i++) {
			volatile float s = a[i] + b[i];   /* round each step to float */
			volatile float d = 2 * x[i];
			volatile float m = s - d;
			volatile float l = m / 2;
			if (op == OP_LAPLACIAN) y[i] = l;
			if (op == OP_ENHANCE)   y[i] = x[i] - l;
			if (op == OP_BLUR)      y[i] = x[i] + l;
		}
		break;
	case OP_OSCILLATION: /* :217-227 */
		a = tmp_plane(w, h); b = tmp_plane(w, h); t = tmp_plane(w, h);
		o_erosion(t, x, w, h, e); o_dilation(a, t, w, h, e); /* opening */
		o_dilation(t, x, w, h, e); o_erosion(b, t, w, h, e); /* closing */
		for (long i = 0; i < n; i++) y[i] = b[i] - a[i];
		break;
	case OP_TOPHAT: /* :229-236 */
		a = tmp_plane(w, h); t = tmp_plane(w, h);
		o_erosion(t, x, w, h, e); o_dilation(a, t, w, h, e);
		for (long i = 0; i < n; i++) y[i] = x[i] - a[i];
		break;
	case OP_BOTHAT: /* :238-245 */
		a = tmp_plane(w, h); t = tmp_plane(w, h);
		o_dilation(t, x, w, h, e); o_erosion(a, t, w, h, e);
		for (long i = 0; i < n; i++) y[i] = a[i] - x[i];
		break;
	case OP_IBLUR: /* :247-254 */
		t = tmp_plane(w, h); o_erosion(t, x, w, h, e);
		for (long i = 0; i < n; i++) {
			volatile float s = x[i] + t[i];
			y[i] = s / 2;
		}
		break;
	case OP_EBLUR: /* :256-263 */
		t = tmp_plane(w, h); o_dilation(t, x, w, h, e);
		for (long i = 0; i < n; i++) {
			volatile float s = x[i] + t[i];
			y[i] = s / 2;
		}
		break;
	case OP_CBLUR: /* :265-275 */
		a = tmp_plane(w, h); b = tmp_plane(w, h);
		o_erosion(a, x, w, h, e); o_dilation(b, x, w, h, e);
		for (long i = 0; i < n; i++)
			y[i] = (float)(0.5 * (double)x[i] + 0.25 * (double)a[i]
					+ 0.25 * (double)b[i]);
		break;
	default:
		return 1;
	}
	free(a); free(b); free(t);
	return 0;
}

/* ---- structuring elements ------------------------------------------------ */

/* Builders follow src/morsi.c:313-417: candidates i (outer) and j (inner) run
 * over [-radius-1, radius+1] with the bounds truncated to int; membership
 * tests are done in double against the float radius.  kind: 0 disk, 1 dysk,
 * 2 hrec, 3 vrec, 4 drec, 5 Drec.  Returns the number of ints written to
 * `out` (capacity `cap` ints), 0 when the reference would return NULL
 * (radius <= 1 or NaN), -1 when `cap` is too small. */
int morsi_oracle_build(int kind, float radius, int *out, int cap)
{
	if (!(radius > 1)) return 0;
	int lo = -radius - 1, hi = radius + 1, cx = 0;
	for (int i = lo; i <= hi; i++) {
		int j0 = kind <= 1 ? lo : 0, j1 = kind <= 1 ? hi : 0;
		for (int j = j0; j <= j1; j++) {
			int keep, dx, dy;
			if (kind == 0) {          /* :321 */
				keep = hypot(i, j) < radius; dx = i; dy = j;
			} else if (kind == 1) {   /* :340 */
				keep = hypot(i, j) < radius && hypot(i, j) >= radius - 1;
				dx = i; dy = j;
			} else {                  /* :358,375,392,409 */
				keep = abs(i) < radius;
				dx = kind == 2 ? i : kind == 3 ? 0 : kind == 4 ? i : -i;
				dy = kind == 2 ? 0 : i;
			}
			if (!keep) continue;
			if (2*cx + 5 >= cap) return -1;
			out[2*cx+4] = dx; out[2*cx+5] = dy; cx++;
		}
	}
	if (cap < 4) return -1;
	out[0] = cx; out[1] = out[2] = out[3] = 0;
	return 2*cx + 4;
}

/* Element-name grammar of src/morsi.c:484-485,496-508, including its quirks:
 * strspn() is a character-SET match, every matching test runs and the last
 * one wins (a NULL from a later builder also overrides an earlier success). */
int morsi_oracle_parse_element(const char *name, int *out, int cap)
{
	static const int cross[]  = {5,0, 0,0, -1,0, 0,0, 1,0, 0,-1, 0,1};
	static const int square[] = {9,0, 0,0, -1,-1,-1,0,-1,1, 0,-1,0,0,0,1,
	                             1,-1,1,0,1,1};
	static const char *kinds[] = {"disk","dysk","hrec","vrec","drec","Drec"};
	int n = 0;
	if (0 == strcmp(name, "cross")) {
		if (cap < 14) return -1;
		memcpy(out, cross, sizeof cross); n = 14;
	}
	if (0 == strcmp(name, "square")) {
		if (cap < 22) return -1;
		memcpy(out, square, sizeof square); n = 22;
	}
	for (int k = 0; k < 6; k++)
		if (4 == strspn(name, kinds[k])) {
			n = morsi_oracle_build(k, atof(name + 4), out, cap);
			if (n < 0) return -1;
		}
	return n;
}

static const char *op_names[OP_COUNT] = {
	"erosion", "dilation", "median", "rank", "opening", "closing",
	"gradient", "igradient", "egradient", "laplacian", "enhance", "blur",
	"oscillation", "tophat", "bothat", "iblur", "eblur", "cblur"
};

int morsi_oracle_parse_operation(const char *name)
{
	for (int i = 0; i < OP_COUNT; i++)
		if (0 == strcmp(name, op_names[i])) return i;
	return -1;
}
