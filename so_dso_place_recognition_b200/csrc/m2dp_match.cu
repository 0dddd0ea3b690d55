// placeholder
#include "../../include/sodso_pr.h"
#include "common.cuh"
namespace sodso {
size_t m2dp_match_workspace_bytes(int, int) { return 256; }
cudaError_t launch_m2dp_match(const double *, int, const double *, int, float *, float *, int, void *, cudaStream_t, int64_t *) { return cudaErrorNotSupported; }
}
