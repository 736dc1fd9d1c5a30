// fp64 (dfloat) instances of the axhelm kernels
#define NRSB_AX_TYPE double
#include "axhelm.inc"
