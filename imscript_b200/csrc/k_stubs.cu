// placeholders for kernel families not built yet
#include "dispatch.cuh"
int morsi_run_tiled(MorsiCtx *, const DevElement *, const MorsiJob &, int *, int *handled) { *handled = 0; return MORSI_OK; }
