// (no kernel family is a placeholder any more; kept so that the build list stays stable)
#include "dispatch.cuh"
