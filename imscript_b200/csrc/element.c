/*
 * element.c -- structuring elements for libmorsi_cuda (host side, plain C).
 *
 * Builders and the element-name grammar re-derive the SAME offset lists, in
 * the SAME order, as the reference (src/morsi.c:313-417,484-485,496-508):
 * element order decides signed-zero results.  The analysis
 * half classifies a list for the kernel dispatcher.
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../../include/morsi_cuda.h"
#include "element.h"

enum shape { DISK, DYSK, HREC, VREC, DREC, DREC_NEG };

/* One generator for the six parametric families.  The reference scans
 * candidates i (outer) and j (inner) over [-radius-1, radius+1], bounds
 * truncated to int, and keeps those passing the family's test (done in double
 * against the float radius).  Line elements only have the outer scan. */
static int *build_family(enum shape s, float radius)
{
	if (!(radius > 1)) return NULL;           /* also rejects NaN */
	int lo = (int)(-radius - 1), hi = (int)(radius + 1);
	long span = (long)hi - lo + 1;
	long cap = (s == DISK || s == DYSK) ? span * span : span;
	int *e = malloc((size_t)(2 * cap + 4) * sizeof *e);
	if (!e) return NULL;
	int count = 0;
	for (int i = lo; i <= hi; i++) {
		if (s == DISK || s == DYSK) {
			for (int j = lo; j <= hi; j++) {
				double d = hypot(i, j);
				int keep = d < radius && (s == DISK || d >= radius - 1);
				if (!keep) continue;
				e[4 + 2*count] = i;
				e[5 + 2*count] = j;
				count++;
			}
		} else if (abs(i) < radius) {
			e[4 + 2*count] = s == VREC ? 0 : s == DREC_NEG ? -i : i;
			e[5 + 2*count] = s == HREC ? 0 : i;
			count++;
		}
	}
	e[0] = count;
	e[1] = e[2] = e[3] = 0;
	return e;
}

int *morsi_build_disk(float r) { return build_family(DISK, r); }
int *morsi_build_dysk(float r) { return build_family(DYSK, r); }
int *morsi_build_hrec(float r) { return build_family(HREC, r); }
int *morsi_build_vrec(float r) { return build_family(VREC, r); }
int *morsi_build_drec(float r) { return build_family(DREC, r); }
int *morsi_build_Drec(float r) { return build_family(DREC_NEG, r); }

void morsi_element_free(int *e) { free(e); }

static int *dup_list(const int *src, int nints)
{
	int *e = malloc(nints * sizeof *e);
	if (e) memcpy(e, src, nints * sizeof *e);
	return e;
}

int morsi_element_parse(const char *name, int **out)
{
	/* src/morsi.c:484-485 */
	static const int cross[]  = {5,0, 0,0, -1,0, 0,0, 1,0, 0,-1, 0,1};
	static const int square[] = {9,0, 0,0, -1,-1, -1,0, -1,1, 0,-1, 0,0, 0,1,
	                             1,-1, 1,0, 1,1};
	static const struct { const char *set; enum shape s; } fam[] = {
		{"disk", DISK}, {"dysk", DYSK}, {"hrec", HREC},
		{"vrec", VREC}, {"drec", DREC}, {"Drec", DREC_NEG}
	};
	if (!name || !out) return MORSI_ERR_INVALID;
	int *e = NULL;
	if (!strcmp(name, "cross"))  e = dup_list(cross, 14);
	if (!strcmp(name, "square")) e = dup_list(square, 22);
	/* The reference runs all six strspn() tests in turn; each match
	 * overwrites the previous result, even with NULL (src/morsi.c:499-504). */
	for (int k = 0; k < 6; k++)
		if (strspn(name, fam[k].set) == 4) {
			free(e);
			e = build_family(fam[k].s, atof(name + 4));
		}
	*out = e;
	return e ? MORSI_OK : MORSI_ERR_INVALID;
}

static const char *const op_table[MORSI_OP_COUNT] = {
	"erosion", "dilation", "median", "rank", "opening", "closing", "gradient",
	"igradient", "egradient", "laplacian", "enhance", "blur", "oscillation",
	"tophat", "bothat", "iblur", "eblur", "cblur"
};

int morsi_operation_parse(const char *name)
{
	if (!name) return -1;
	for (int i = 0; i < MORSI_OP_COUNT; i++)
		if (!strcmp(name, op_table[i])) return i;
	return -1;
}

const char *morsi_operation_name(int op)
{
	return op >= 0 && op < MORSI_OP_COUNT ? op_table[op] : NULL;
}

/* ---- analysis ------------------------------------------------------------ */

int morsi_element_analyze(const int *e, morsi_element_info *info)
{
	memset(info, 0, sizeof *info);
	if (!e || e[0] < 0) return 1;
	int n = info->n = e[0];
	info->kind = MORSI_EK_DIRECT;
	if (n == 0) return 0;
	int xmin = e[4] - e[2], xmax = xmin, ymin = e[5] - e[3], ymax = ymin;
	for (int k = 0; k < n; k++) {
		int dx = e[4 + 2*k] - e[2], dy = e[5 + 2*k] - e[3];
		if (dx < xmin) xmin = dx;
		if (dx > xmax) xmax = dx;
		if (dy < ymin) ymin = dy;
		if (dy > ymax) ymax = dy;
	}
	info->xmin = xmin; info->xmax = xmax; info->ymin = ymin; info->ymax = ymax;

	/* occupancy grid over the box, only when it is small enough to matter */
	long bw = (long)xmax - xmin + 1, bh = (long)ymax - ymin + 1;
	if (bw > 2 * MORSI_MAX_REACH_ROWRUN + 1 || bh > 2 * MORSI_MAX_REACH_ROWRUN + 1) {
		/* long one-row / one-column lists (hrecR, vrecR, user lines): the line
		 * kernels need to know whether the list is a run without repeats */
		if ((bw == 1 || bh == 1) && bw * bh <= (1L << 24)) {
			unsigned char *seen = calloc((size_t)(bw * bh), 1);
			if (!seen) { info->has_duplicates = 1; return 0; }
			for (int k = 0; k < n; k++) {
				long t = bw == 1 ? (e[5 + 2*k] - e[3]) - ymin : (e[4 + 2*k] - e[2]) - xmin;
				if (seen[t]) info->has_duplicates = 1;
				seen[t] = 1;
			}
			free(seen);
		}
		return 0;
	}
	unsigned char *grid = calloc((size_t)(bw * bh), 1);
	if (!grid) return 0;
	for (int k = 0; k < n; k++) {
		int dx = e[4 + 2*k] - e[2], dy = e[5 + 2*k] - e[3];
		unsigned char *c = &grid[(dy - ymin) * bw + (dx - xmin)];
		if (*c) info->has_duplicates = 1;
		*c = 1;
	}
	if (xmin >= -1 && xmax <= 1 && ymin >= -1 && ymax <= 1) {
		info->kind = MORSI_EK_SMALL;
		for (int dy = ymin; dy <= ymax; dy++)
			for (int dx = xmin; dx <= xmax; dx++)
				if (grid[(dy - ymin) * bw + (dx - xmin)])
					info->mask3x3 |= 1u << ((dy + 1) * 3 + (dx + 1));
	}
	/* row-run test: box symmetric about 0, every row a centred run, widths
	 * non-increasing away from row 0 */
	if (xmin == -xmax && ymin == -ymax && ymax <= MORSI_MAX_REACH_ROWRUN
			&& xmax <= MORSI_MAX_REACH_ROWRUN) {
		int ok = 1, reach = ymax;
		for (int dy = -reach; dy <= reach && ok; dy++) {
			const unsigned char *row = &grid[(dy - ymin) * bw];
			int hw = -1;
			for (int dx = 0; dx <= xmax; dx++)
				if (row[dx - xmin]) hw = dx; else break;
			if (hw < 0) { ok = 0; break; }
			for (int dx = -xmax; dx <= xmax; dx++) {
				int want = abs(dx) <= hw;
				if (row[dx - xmin] != want) { ok = 0; break; }
			}
			info->halfwidth[dy + reach] = hw;
		}
		for (int d = 1; d <= reach && ok; d++) {
			if (info->halfwidth[reach + d] > info->halfwidth[reach + d - 1]) ok = 0;
			if (info->halfwidth[reach - d] != info->halfwidth[reach + d]) ok = 0;
		}
		if (ok && info->kind != MORSI_EK_SMALL) {
			info->kind = MORSI_EK_ROWRUN;
			info->reach = reach;
		} else if (ok) {
			info->reach = reach;
		}
	}
	free(grid);
	return 0;
}

int morsi_element_describe(const int *e, char *buf, size_t buflen)
{
	morsi_element_info info;
	if (!buf || !buflen || morsi_element_analyze(e, &info)) return MORSI_ERR_INVALID;
	const char *k = info.kind == MORSI_EK_SMALL ? "small3x3"
	              : info.kind == MORSI_EK_ROWRUN ? "rowrun" : "direct";
	snprintf(buf, buflen, "%s n=%d box=[%d,%d]x[%d,%d]%s", k, info.n,
			info.xmin, info.xmax, info.ymin, info.ymax,
			info.has_duplicates ? " dup" : "");
	return MORSI_OK;
}
