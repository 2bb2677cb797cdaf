// Kernels and drivers of the GoldilocksRing instantiated in their own translation unit (see ring_ops.cuh).
#include "ring_ops.cuh"
namespace lf { RingOps* ring_ops_goldilocks() { static RingOpsImpl<GoldilocksRing> ops; return &ops; } }
