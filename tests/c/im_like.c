/* A multi-call binary in the style of src/im.c:4-6: the tool's main is
 * compiled with -DHIDE_ALL_MAINS and reached through main_morsi().
 *   im_like morsi ELEMENT OPERATION [in [out]]
 */
#include <stdio.h>
#include <string.h>
int main_morsi(int c, char **v);
int main(int c, char **v)
{
	if (c < 2 || strcmp(v[1], "morsi")) { fprintf(stderr, "usage:\n\tim_like morsi ...\n"); return 1; }
	return main_morsi(c - 1, v + 1);
}
