// extern "C" entry points of the convolution engine: argument validation + kernel selection.
#include "common.cuh"
#include "conv_common.cuh"

namespace {

int check_shape(const mmt_conv_shape* s, const char* who) {
  MMT_REQUIRE(s != nullptr, "%s: null shape", who);
  MMT_REQUIRE(s->batch >= 0 && s->in_h > 0 && s->in_w > 0 && s->in_c > 0 && s->out_c > 0 && s->k_h > 0 &&
                  s->k_w > 0 && s->stride > 0 && s->pad >= 0,
              "%s: bad shape", who);
  const int oh = (s->in_h + 2 * s->pad - s->k_h) / s->stride + 1;
  const int ow = (s->in_w + 2 * s->pad - s->k_w) / s->stride + 1;
  MMT_REQUIRE(oh == s->out_h && ow == s->out_w, "%s: out_h/out_w (%d,%d) inconsistent with the input (%d,%d expected)",
              who, s->out_h, s->out_w, oh, ow);
  return MMT_OK;
}

size_t filter_bytes(const mmt_conv_shape& s) {
  return sizeof(float) * (size_t)s.out_c * s.k_h * s.k_w * s.in_c;
}

}  // namespace

extern "C" size_t mmt_conv_workspace_bytes(const mmt_conv_shape* s) {
  if (!s) return 0;
  return mmt::align_up(filter_bytes(*s), 256) + 256;
}

extern "C" int mmt_conv_fprop_f32(const float* x, const float* w, const float* scale, const float* bias,
                                  const float* residual, float* y, const mmt_conv_shape* s, int flags,
                                  int impl, void* stream_) {
  if (int rc = check_shape(s, "conv_fprop")) return rc;
  if (s->batch == 0) return MMT_OK;
  MMT_REQUIRE(x && w && y, "conv_fprop: null pointer");
  cudaStream_t stream = (cudaStream_t)stream_;
  const bool tc_ok = mmt::conv_tc_supported_fprop(*s);
  if (impl == MMT_CONV_TCGEN05 && !tc_ok)
    return mmt::fail(MMT_ERR_UNSUPPORTED, "conv_fprop: shape not supported by the tcgen05 kernel");
  if (impl == MMT_CONV_SIMT || !tc_ok) return mmt::conv_fprop_simt(x, w, scale, bias, residual, y, *s, flags, stream);
  return mmt::conv_fprop_tc(x, w, scale, bias, residual, y, *s, flags, stream);
}

extern "C" size_t mmt_conv_stem_workspace_bytes(const mmt_conv_shape* s) {
  if (!s || !mmt::conv_stem_tc_supported(*s)) return 0;
  return mmt::conv_stem_tc_workspace(*s);
}

extern "C" int mmt_conv_stem_fprop_f32(const float* x, const float* w, const float* scale, const float* bias, float* y,
                                       const mmt_conv_shape* s, int flags, void* workspace, size_t workspace_bytes,
                                       void* stream_) {
  if (int rc = check_shape(s, "conv_stem_fprop")) return rc;
  if (s->batch == 0) return MMT_OK;
  MMT_REQUIRE(x && w && y, "conv_stem_fprop: null pointer");
  if (!mmt::conv_stem_tc_supported(*s))
    return mmt::fail(MMT_ERR_UNSUPPORTED, "conv_stem_fprop: only 7x7 / stride 2 / pad 3 over 3 input channels");
  MMT_REQUIRE(((uintptr_t)workspace & 255) == 0, "conv_stem_fprop: workspace must be 256-byte aligned");
  return mmt::conv_stem_tc(x, w, scale, bias, y, *s, flags, workspace, workspace_bytes, (cudaStream_t)stream_);
}

extern "C" int mmt_conv_dgrad_f32(const float* dy, const float* w, float* dx, const mmt_conv_shape* s, int impl,
                                  void* workspace, size_t workspace_bytes, void* stream_) {
  if (int rc = check_shape(s, "conv_dgrad")) return rc;
  if (s->batch == 0) return MMT_OK;
  MMT_REQUIRE(dy && w && dx, "conv_dgrad: null pointer");
  cudaStream_t stream = (cudaStream_t)stream_;
  const bool tc_ok = mmt::conv_tc_supported_dgrad(*s);
  if (impl == MMT_CONV_TCGEN05 && !tc_ok)
    return mmt::fail(MMT_ERR_UNSUPPORTED, "conv_dgrad: shape not supported by the tcgen05 kernel");
  if (impl == MMT_CONV_SIMT || !tc_ok) return mmt::conv_dgrad_simt(dy, w, dx, *s, stream);
  const size_t need = mmt::align_up(filter_bytes(*s), 256);
  if (workspace == nullptr || workspace_bytes < need)
    return mmt::fail(MMT_ERR_WORKSPACE_TOO_SMALL, "conv_dgrad: workspace %zu < required %zu", workspace_bytes, need);
  MMT_REQUIRE(((uintptr_t)workspace & 255) == 0, "conv_dgrad: workspace must be 256-byte aligned");
  float* w_t = (float*)workspace;
  if (int rc = mmt::conv_filter_for_dgrad(w, w_t, *s, stream)) return rc;
  return mmt::conv_dgrad_tc(dy, w_t, dx, *s, stream);
}

// Fused backward of Conv -> FrozenBN(scale) [-> ReLU] chains (reference layers/batch_norm.py:19-24 + resnet.py:254-275 under
// autograd): `dy` is the gradient of the BN OUTPUT; the BN scale is folded into the transformed filter, and the epilogue
// either adds `aux` (the skip connection's gradient) or masks with aux > 0 (ReLU backward of the layer that produced this
// conv's input) — the elementwise passes torch's autograd would launch around the convolution.  tcgen05 path only.
extern "C" int mmt_conv_dgrad_fused_f32(const float* dy, const float* w, const float* k_scale, const float* aux,
                                        int aux_mode, float* dx, const mmt_conv_shape* s, void* workspace,
                                        size_t workspace_bytes, void* stream_) {
  if (int rc = check_shape(s, "conv_dgrad_fused")) return rc;
  if (s->batch == 0) return MMT_OK;
  MMT_REQUIRE(dy && w && dx, "conv_dgrad_fused: null pointer");
  MMT_REQUIRE(aux_mode == 0 || ((aux_mode == MMT_DGRAD_AUX_ADD || aux_mode == MMT_DGRAD_AUX_MASK) && aux != nullptr),
              "conv_dgrad_fused: aux_mode %d needs an aux tensor", aux_mode);
  if (!mmt::conv_tc_supported_dgrad(*s))
    return mmt::fail(MMT_ERR_UNSUPPORTED, "conv_dgrad_fused: shape not supported by the tcgen05 kernel");
  if (aux_mode && s->stride != 1)
    return mmt::fail(MMT_ERR_UNSUPPORTED, "conv_dgrad_fused: the epilogue tensor needs a stride-1 convolution");
  cudaStream_t stream = (cudaStream_t)stream_;
  const size_t need = mmt::align_up(filter_bytes(*s), 256);
  if (workspace == nullptr || workspace_bytes < need)
    return mmt::fail(MMT_ERR_WORKSPACE_TOO_SMALL, "conv_dgrad_fused: workspace %zu < required %zu", workspace_bytes, need);
  MMT_REQUIRE(((uintptr_t)workspace & 255) == 0, "conv_dgrad_fused: workspace must be 256-byte aligned");
  float* w_t = (float*)workspace;
  if (int rc = mmt::conv_filter_for_dgrad(w, w_t, *s, stream, k_scale)) return rc;
  return mmt::conv_dgrad_tc(dy, w_t, dx, *s, stream, aux, aux_mode);
}

// dw (+)= row_scale[o] * wgrad(x, dy): `dy` is the gradient of the BN output, row_scale the frozen-BN scale.
extern "C" int mmt_conv_wgrad_scaled_f32(const float* x, const float* dy, const float* row_scale, float* dw,
                                         const mmt_conv_shape* s, int accumulate, void* stream_) {
  if (int rc = check_shape(s, "conv_wgrad_scaled")) return rc;
  MMT_REQUIRE(dw != nullptr, "conv_wgrad_scaled: null dw");
  if (s->batch == 0) {
    if (!accumulate) MMT_CUDA(cudaMemsetAsync(dw, 0, filter_bytes(*s), (cudaStream_t)stream_));
    return MMT_OK;
  }
  MMT_REQUIRE(x && dy, "conv_wgrad_scaled: null pointer");
  if (!mmt::conv_tc_supported_wgrad(*s))
    return mmt::fail(MMT_ERR_UNSUPPORTED, "conv_wgrad_scaled: shape not supported by the tcgen05 kernel");
  return mmt::conv_wgrad_tc(x, dy, dw, nullptr, *s, nullptr, 0, (cudaStream_t)stream_, accumulate != 0, row_scale);
}

extern "C" int mmt_conv_wgrad_f32(const float* x, const float* dy, float* dw, float* dbias,
                                  const mmt_conv_shape* s, int impl, void* workspace, size_t workspace_bytes,
                                  void* stream_) {
  if (int rc = check_shape(s, "conv_wgrad")) return rc;
  MMT_REQUIRE(dw != nullptr, "conv_wgrad: null dw");
  cudaStream_t stream = (cudaStream_t)stream_;
  if (s->batch == 0) {
    MMT_CUDA(cudaMemsetAsync(dw, 0, filter_bytes(*s), stream));
    if (dbias) MMT_CUDA(cudaMemsetAsync(dbias, 0, sizeof(float) * (size_t)s->out_c, stream));
    return MMT_OK;
  }
  MMT_REQUIRE(x && dy, "conv_wgrad: null pointer");
  const bool tc_ok = mmt::conv_tc_supported_wgrad(*s);
  if (impl == MMT_CONV_TCGEN05 && !tc_ok)
    return mmt::fail(MMT_ERR_UNSUPPORTED, "conv_wgrad: shape not supported by the tcgen05 kernel");
  if (impl == MMT_CONV_SIMT || !tc_ok) return mmt::conv_wgrad_simt(x, dy, dw, dbias, *s, stream);
  if (int rc = mmt::conv_wgrad_tc(x, dy, dw, nullptr, *s, workspace, workspace_bytes, stream)) return rc;
  if (dbias) return mmt::conv_bias_grad(dy, (long long)s->batch * s->out_h * s->out_w, s->out_c, dbias, stream);
  return MMT_OK;
}

// dw += wgrad(x, dy): the tensor-core kernel's split-K partial tiles are reduced into dw by TMA anyway, so accumulating
// into a live gradient buffer (the optimizer's flat arena) costs nothing and saves the memset, the temporary and the
// `grad += dw` pass an autograd AccumulateGrad node would launch.  tcgen05 path only.
extern "C" int mmt_conv_wgrad_accumulate_f32(const float* x, const float* dy, float* dw, const mmt_conv_shape* s,
                                             void* stream_) {
  if (int rc = check_shape(s, "conv_wgrad_accumulate")) return rc;
  MMT_REQUIRE(dw != nullptr, "conv_wgrad_accumulate: null dw");
  if (s->batch == 0) return MMT_OK;
  MMT_REQUIRE(x && dy, "conv_wgrad_accumulate: null pointer");
  if (!mmt::conv_tc_supported_wgrad(*s))
    return mmt::fail(MMT_ERR_UNSUPPORTED, "conv_wgrad_accumulate: shape not supported by the tcgen05 kernel");
  return mmt::conv_wgrad_tc(x, dy, dw, nullptr, *s, nullptr, 0, (cudaStream_t)stream_, true);
}
