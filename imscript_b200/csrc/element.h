/* element.h -- host-side structuring-element compiler (internal). */
#ifndef MORSI_ELEMENT_H
#define MORSI_ELEMENT_H
#ifdef __cplusplus
extern "C" {
#endif

enum morsi_element_kind {
	MORSI_EK_DIRECT = 0,  /* arbitrary offset list: tiled direct kernels */
	MORSI_EK_SMALL,       /* all offsets within the 3x3 neighbourhood */
	MORSI_EK_ROWRUN       /* convex & symmetric: one centred run per row, widths
	                         non-increasing away from the centre row (disks,
	                         squares, hrec, vrec) */
};

#define MORSI_MAX_REACH_ROWRUN 32

typedef struct morsi_element_info {
	int n;                      /* e[0] */
	int xmin, xmax, ymin, ymax; /* box of effective offsets (dx-e[2], dy-e[3]) */
	int kind;
	int has_duplicates;
	/* MORSI_EK_SMALL: bit (dy+1)*3+(dx+1) set when the offset is present */
	unsigned mask3x3;
	/* MORSI_EK_ROWRUN: half-width of the run on row dy = -reach..reach */
	int reach;
	int halfwidth[2 * MORSI_MAX_REACH_ROWRUN + 1];
} morsi_element_info;

/* 0 on success, non-zero if the list is malformed (NULL, negative count). */
int morsi_element_analyze(const int *e, morsi_element_info *info);

#ifdef __cplusplus
}
#endif
#endif
