// Kernels and drivers of the BabyBearRing instantiated in their own translation unit (see ring_ops.cuh).
#include "ring_ops.cuh"
namespace lf { RingOps* ring_ops_babybear() { static RingOpsImpl<BabyBearRing> ops; return &ops; } }
