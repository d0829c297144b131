// centre_p256.cu -- the centre kernels for up to 256 neighbours per centre (centre_impl.cuh)
#include "centre_impl.cuh"
GAPCU_CENTRE_INSTANCE_SE(256)
