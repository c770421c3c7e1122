// placeholder until the fast scan lands: everything routes to the reference-order kernel
#include "common.cuh"
namespace gf {
bool scan_fast_supports(int, int) { return false; }
cudaError_t launch_scan_fast(int, const ScanArgs &, int, int, cudaStream_t, int *) { return cudaErrorNotSupported; }
}
