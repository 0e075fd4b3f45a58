// Internal interface between the conv dispatch (conv_api.cu) and the two kernel families.
#pragma once
#include <cuda_runtime.h>

#include "mmtpsm_b200.h"

namespace mmt {

// fp32 FFMA kernels (conv_simt.cu)
int conv_fprop_simt(const float* x, const float* w, const float* scale, const float* bias,
                    const float* residual, float* y, const mmt_conv_shape& s, int flags,
                    cudaStream_t stream);
int conv_dgrad_simt(const float* dy, const float* w, float* dx, const mmt_conv_shape& s,
                    cudaStream_t stream);
int conv_wgrad_simt(const float* x, const float* dy, float* dw, float* dbias, const mmt_conv_shape& s,
                    cudaStream_t stream);
// dbias[c] = sum over pixels of dy[p][c] (dbias is overwritten)
int conv_bias_grad(const float* dy, long long pixels, int out_c, float* dbias, cudaStream_t stream);

// TMA + tcgen05 (TF32) kernels (conv_tc.cu)
bool conv_tc_supported_fprop(const mmt_conv_shape& s);
bool conv_tc_supported_dgrad(const mmt_conv_shape& s);
bool conv_tc_supported_wgrad(const mmt_conv_shape& s);
int conv_fprop_tc(const float* x, const float* w, const float* scale, const float* bias,
                  const float* residual, float* y, const mmt_conv_shape& s, int flags,
                  cudaStream_t stream);
// 7x7 / stride-2 / pad-3 stem over 3 input channels as a 4x4 implicit GEMM over the space-to-depth image
bool conv_stem_tc_supported(const mmt_conv_shape& s);
size_t conv_stem_tc_workspace(const mmt_conv_shape& s);
int conv_stem_tc(const float* x, const float* w, const float* scale, const float* bias, float* y,
                 const mmt_conv_shape& s, int flags, void* workspace, size_t workspace_bytes, cudaStream_t stream);
// w_t must hold the dgrad-transformed filter [in_c][k_h][k_w][out_c] (rotated 180 degrees)
// `aux` (dense, shaped like dx) with aux_mode MMT_DGRAD_AUX_ADD: dx = conv_T(dy) + aux;  MMT_DGRAD_AUX_MASK: dx = aux > 0 ? conv_T(dy) : 0
int conv_dgrad_tc(const float* dy, const float* w_t, float* dx, const mmt_conv_shape& s,
                  cudaStream_t stream, const float* aux = nullptr, int aux_mode = 0);
// row_scale (nullable): dw[o] (+)= row_scale[o] * wgrad(x, dy)[o]
int conv_wgrad_tc(const float* x, const float* dy, float* dw, float* dbias, const mmt_conv_shape& s,
                  void* workspace, size_t workspace_bytes, cudaStream_t stream, bool accumulate = false,
                  const float* row_scale = nullptr);
// [out_c][k_h][k_w][in_c] -> [in_c][k_h][k_w][out_c] with the taps rotated by 180 degrees; k_scale (nullable) multiplies
// every element by its OUT channel's factor
int conv_filter_for_dgrad(const float* w, float* w_t, const mmt_conv_shape& s, cudaStream_t stream,
                          const float* k_scale = nullptr);

}  // namespace mmt
