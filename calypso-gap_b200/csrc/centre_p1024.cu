// centre_p1024.cu -- the centre kernels for up to 1024 neighbours per centre (centre_impl.cuh)
#include "centre_impl.cuh"
GAPCU_CENTRE_INSTANCE(1024)
