// placeholder
#include "../../include/sodso_pr.h"
#include "common.cuh"
namespace sodso {
size_t m2dp_generate_workspace_bytes(int, bool) { return 256; }
cudaError_t launch_m2dp_generate(const double *, const float *, const int64_t *, int, double, bool, double *, void *, size_t, int, cudaStream_t, int64_t *) { return cudaErrorNotSupported; }
}
