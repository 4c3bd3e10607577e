// placeholder until the tcgen05 kernel lands (next commit)
#include "dtb_common.cuh"
#include "cin_impl.h"
namespace dtb {
bool cin_tc_supported(const CinShape&) { return false; }
size_t cin_tc_saved_bytes(const CinShape&, int) { return 0; }
size_t cin_tc_workspace_bytes(const CinShape&, int, int) { return 0; }
int cin_tc_fwd(const CinShape&, const int32_t*, const float*, const int64_t*, const float*, const float*, float*,
               void*, void*, size_t, int, int, int, int*, cudaStream_t) { return DTB_ERR_UNSUPPORTED; }
int cin_tc_bwd(const CinShape&, const int32_t*, const float*, const int64_t*, const float*, const float*,
               const void*, float*, float*, float*, void*, size_t, int, int, int, cudaStream_t) { return DTB_ERR_UNSUPPORTED; }
}
