// centre_p512.cu -- the centre kernels for up to 512 neighbours per centre (centre_impl.cuh)
#include "centre_impl.cuh"
GAPCU_CENTRE_INSTANCE(512)
