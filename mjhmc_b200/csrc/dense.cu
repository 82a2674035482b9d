// K4 placeholder translation unit: filled in by the dense-contraction kernels.
#include "dense.h"
namespace mjhmc {
bool dense_supported(int, int, int, int) { return false; }
cudaError_t launch_dense(int, int, const LaunchParams&, cudaStream_t) { return cudaErrorNotSupported; }
}
