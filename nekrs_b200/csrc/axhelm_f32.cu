// fp32 (pfloat) instances of the axhelm kernels
#define NRSB_AX_TYPE float
#include "axhelm.inc"
