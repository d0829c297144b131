// centre_p128.cu -- the centre kernels for up to 128 neighbours per centre (centre_impl.cuh)
#include "centre_impl.cuh"
GAPCU_CENTRE_INSTANCE_SE(128)
