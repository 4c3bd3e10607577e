// C-ABI entry points for CIN: shape validation and dispatch between the tensor-core kernel
// (cin_tc.cu, product path) and the exact-fp32 any-shape formulation (cin_fp32.cu).
#include "dtb_common.cuh"
#include "cin_shapes.h"
#include "cin_impl.h"

using namespace dtb;

static bool use_tc(const CinShape& s, int precision) {
  if (precision == DTB_CIN_FP32) return false;
  return cin_tc_supported(s);
}

// DTB_CIN_AUTO: single-pass fp16 tensor-core kernels where all three (forward, data gradient, weight gradient) take the
// shape, else the bf16x3 tensor-core kernels, else the any-shape formulation.  Forward and backward of one step resolve
// identically (same shape, same process-wide switches).
static int resolve_precision(const CinShape& s, int precision) {
  if (precision == DTB_CIN_AUTO && cin_tc_f16_auto(s)) return DTB_CIN_TC_F16X1;
  return precision;
}

extern "C" {

int dtb_cin_tc_supported(int F, int D, const int* layer_sizes_host, int n_layers, int direct) {
  CinShape s;
  if (!s.init(F, D, layer_sizes_host, n_layers, direct)) return 0;
  return cin_tc_supported(s) ? 1 : 0;
}

int dtb_cin_resolved_precision(int F, int D, const int* layer_sizes_host, int n_layers, int direct, int precision) {
  CinShape s;
  if (!s.init(F, D, layer_sizes_host, n_layers, direct)) return DTB_ERR_INVALID_ARG;
  precision = resolve_precision(s, precision);
  if (precision == DTB_CIN_AUTO) return cin_tc_supported(s) ? DTB_CIN_TC_BF16X3 : DTB_CIN_FP32;
  return precision;
}

size_t dtb_cin_saved_bytes(int B, int F, int D, const int* layer_sizes_host, int n_layers, int direct) {
  CinShape s;
  if (!s.init(F, D, layer_sizes_host, n_layers, direct) || B <= 0) return 0;
  size_t a = cin_fp32_saved_bytes(s, B);
  size_t b = cin_tc_supported(s) ? cin_tc_saved_bytes(s, B) : 0;
  return a > b ? a : b;
}

size_t dtb_cin_workspace_bytes(int B, int F, int D, const int* layer_sizes_host, int n_layers, int direct,
                               int training) {
  CinShape s;
  if (!s.init(F, D, layer_sizes_host, n_layers, direct) || B <= 0) return 0;
  size_t a = cin_fp32_workspace_bytes(s, B, training);
  size_t b = cin_tc_supported(s) ? cin_tc_workspace_bytes(s, B, training) : 0;
  return a > b ? a : b;
}

int dtb_cin_fwd(const int32_t* idx, const float* table, const int64_t* row_offsets, const float* weights,
                const float* bias, float* pooled, void* saved, void* workspace, size_t workspace_bytes, int B,
                int F, int D, const int* layer_sizes_host, int n_layers, int direct, int act, int precision,
                int* status, void* stream) {
  DTB_CHECK_ARG(idx && table && row_offsets && weights && pooled && workspace, "NULL argument");
  DTB_CHECK_ARG(act == DTB_ACT_NONE || act == DTB_ACT_RELU, "unsupported activation");
  DTB_CHECK_ARG(precision >= 0 && precision <= DTB_CIN_TC_F16X1, "bad precision code");
  CinShape s;
  if (!s.init(F, D, layer_sizes_host, n_layers, direct)) {
    set_error("dtb_cin_fwd: invalid CIN configuration (cross_layer_size must be even except for the last "
              "layer when direct=False; 1..%d layers)", kCinMaxLayers);
    return DTB_ERR_INVALID_ARG;
  }
  if (B <= 0) return DTB_OK;
  cudaStream_t st = (cudaStream_t)stream;
  precision = resolve_precision(s, precision);
  if (use_tc(s, precision)) {
    const int single = precision == DTB_CIN_TC_BF16X1 || precision == DTB_CIN_TC_F16X1;
    return cin_tc_fwd(s, idx, table, row_offsets, weights, bias, pooled, saved, workspace, workspace_bytes, B,
                      act, single ? 1 : 3, precision == DTB_CIN_TC_F16X1 ? 1 : 0, status, st);
  }
  if (precision == DTB_CIN_TC_BF16X3 || precision == DTB_CIN_TC_BF16X1 || precision == DTB_CIN_TC_F16X1) {
    set_error("dtb_cin_fwd: tensor-core path requested but shape unsupported (F=%d D=%d)", F, D);
    return DTB_ERR_UNSUPPORTED;
  }
  return cin_fp32_fwd(s, idx, table, row_offsets, weights, bias, pooled, saved, workspace, workspace_bytes, B,
                      act, status, st);
}

static int cin_bwd_impl(const int32_t* idx, const float* table, const int64_t* row_offsets, const float* weights,
                        const float* d_pooled, const void* saved, float* grad_table, float* d_weights, float* d_bias,
                        void* workspace, size_t workspace_bytes, int B, int F, int D, const int* layer_sizes_host,
                        int n_layers, int direct, int act, int precision, int phase, void* stream) {
  DTB_CHECK_ARG(idx && table && row_offsets && weights && d_pooled && saved && grad_table && d_weights &&
                    workspace,
                "NULL argument");
  DTB_CHECK_ARG(act == DTB_ACT_NONE || act == DTB_ACT_RELU, "unsupported activation");
  DTB_CHECK_ARG(precision >= 0 && precision <= DTB_CIN_TC_F16X1, "bad precision code");
  CinShape s;
  if (!s.init(F, D, layer_sizes_host, n_layers, direct)) {
    set_error("dtb_cin_bwd: invalid CIN configuration");
    return DTB_ERR_INVALID_ARG;
  }
  if (B <= 0) return DTB_OK;
  cudaStream_t st = (cudaStream_t)stream;
  precision = resolve_precision(s, precision);
  if (use_tc(s, precision))
    return cin_tc_bwd(s, idx, table, row_offsets, weights, d_pooled, saved, grad_table, d_weights, d_bias,
                      workspace, workspace_bytes, B, act, precision == DTB_CIN_TC_BF16X1 ? 1 : 3,
                      precision == DTB_CIN_TC_F16X1 ? 1 : 0, phase, st);
  if (precision == DTB_CIN_TC_BF16X3 || precision == DTB_CIN_TC_BF16X1 || precision == DTB_CIN_TC_F16X1) {
    set_error("dtb_cin_bwd: tensor-core path requested but shape unsupported (F=%d D=%d)", F, D);
    return DTB_ERR_UNSUPPORTED;
  }
  if (phase == 2) return DTB_OK;
  return cin_fp32_bwd(s, idx, table, row_offsets, weights, d_pooled, saved, grad_table, d_weights, d_bias,
                      workspace, workspace_bytes, B, act, st);
}

int dtb_cin_bwd(const int32_t* idx, const float* table, const int64_t* row_offsets, const float* weights,
                const float* d_pooled, const void* saved, float* grad_table, float* d_weights, float* d_bias,
                void* workspace, size_t workspace_bytes, int B, int F, int D, const int* layer_sizes_host,
                int n_layers, int direct, int act, int precision, void* stream) {
  return cin_bwd_impl(idx, table, row_offsets, weights, d_pooled, saved, grad_table, d_weights, d_bias, workspace,
                      workspace_bytes, B, F, D, layer_sizes_host, n_layers, direct, act, precision, 0, stream);
}

int dtb_cin_bwd_phase(const int32_t* idx, const float* table, const int64_t* row_offsets, const float* weights,
                      const float* d_pooled, const void* saved, float* grad_table, float* d_weights, float* d_bias,
                      void* workspace, size_t workspace_bytes, int B, int F, int D, const int* layer_sizes_host,
                      int n_layers, int direct, int act, int precision, int phase, void* stream) {
  DTB_CHECK_ARG(phase == 1 || phase == 2, "phase must be 1 (embedding gradient) or 2 (weight gradient)");
  return cin_bwd_impl(idx, table, row_offsets, weights, d_pooled, saved, grad_table, d_weights, d_bias, workspace,
                      workspace_bytes, B, F, D, layer_sizes_host, n_layers, direct, act, precision, phase, stream);
}

}  // extern "C"
