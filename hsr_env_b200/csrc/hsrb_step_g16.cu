// Action kernel instantiated for 16 lanes per environment (see hsrb_kernels.cuh).
#include "hsrb_kernels.cuh"
HSRB_DEFINE_G(16)
