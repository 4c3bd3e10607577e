// InnerProduct / OuterProduct (layers.py:473-487, 541-581) and the AutoInt interacting layer
// (layers.py:115-150).  Kernels land in a later commit of this round; until then the entry points
// report DTB_ERR_UNSUPPORTED (the host raises -- there is no CPU fallback).
#include "dtb_common.cuh"
using namespace dtb;
extern "C" {
int dtb_pnn_fwd(const int32_t*, const float*, const int64_t*, const float*, float*, float*, int, int, int, int,
                int*, void*) { set_error("dtb_pnn_fwd: not implemented yet"); return DTB_ERR_UNSUPPORTED; }
int dtb_pnn_bwd(const int32_t*, const float*, const int64_t*, const float*, const float*, const float*, float*,
                float*, int, int, int, int, void*) { set_error("dtb_pnn_bwd: not implemented yet"); return DTB_ERR_UNSUPPORTED; }
int dtb_attention_fwd(const float*, const float*, const float*, float*, int, int, int, int, int, void*) {
  set_error("dtb_attention_fwd: not implemented yet"); return DTB_ERR_UNSUPPORTED; }
int dtb_attention_bwd(const float*, const float*, const float*, const float*, float*, float*, float*, int, int,
                      int, int, int, void*) { set_error("dtb_attention_bwd: not implemented yet"); return DTB_ERR_UNSUPPORTED; }
}
