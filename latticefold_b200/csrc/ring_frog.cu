// Kernels and drivers of the FrogRing instantiated in their own translation unit (see ring_ops.cuh).
#include "ring_ops.cuh"
namespace lf { RingOps* ring_ops_frog() { static RingOpsImpl<FrogRing> ops; return &ops; } }
