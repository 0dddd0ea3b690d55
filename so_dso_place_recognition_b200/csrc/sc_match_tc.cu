// placeholder (replaced by the tcgen05 kernel)
#include "../../include/sodso_pr.h"
#include "common.cuh"
namespace sodso {
size_t sc_tc_db_bytes(int) { return 256; }
size_t sc_tc_query_bytes(int) { return 256; }
cudaError_t launch_sc_tc_prep_db(const double *, int, void *, cudaStream_t, int64_t *) { return cudaErrorNotSupported; }
cudaError_t launch_sc_tc_prep_query(const double *, int, void *, cudaStream_t, int64_t *) { return cudaErrorNotSupported; }
cudaError_t launch_sc_match_tc(const void *, int, const void *, int, float *, float *, int, int, cudaStream_t, int64_t *) { return cudaErrorNotSupported; }
}
