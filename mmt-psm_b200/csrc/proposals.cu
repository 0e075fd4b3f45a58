// Proposal decoding for the RPN / box post-processors: BoxCoder.decode + clip_to_image + remove_small_boxes as ONE
// elementwise pass over every candidate of every (image, level) of a pass.
//
// Reference chain (each step a handful of full-tensor torch ops, per level and per image):
//   maskrcnn_benchmark/modeling/box_coder.py:56-95           decode (legacy +1 extents, dw/dh clamped, x2/y2 carry -1)
//   maskrcnn_benchmark/structures/bounding_box.py:211-221    clip_to_image(remove_empty=False)
//   maskrcnn_benchmark/structures/boxlist_ops.py:37-54       remove_small_boxes (w, h >= min_size in xywh = +1 extents)
//   maskrcnn_benchmark/modeling/rpn/inference.py:107-124     call order
// The arithmetic keeps the reference's operation order with explicit round-to-nearest mul / add (no FMA
// contraction), so boxes agree with torch's eager result to the last bit of expf.
#include "common.cuh"

namespace {

__global__ void __launch_bounds__(256) decode_clip_kernel(
    const float4* __restrict__ deltas, const float4* __restrict__ anchors, const float* __restrict__ scores,
    const float* __restrict__ image_wh, int per_image, long long total, float wx, float wy, float ww, float wh,
    float xform_clip, float min_size, float4* __restrict__ boxes, float* __restrict__ scores_nms,
    unsigned char* __restrict__ valid) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const float4 d = __ldg(deltas + i);
    const float4 a = __ldg(anchors + i);
    const float w = __fadd_rn(__fsub_rn(a.z, a.x), 1.f);
    const float h = __fadd_rn(__fsub_rn(a.w, a.y), 1.f);
    const float cx = __fadd_rn(a.x, __fmul_rn(0.5f, w));
    const float cy = __fadd_rn(a.y, __fmul_rn(0.5f, h));
    const float dx = __fdiv_rn(d.x, wx), dy = __fdiv_rn(d.y, wy);
    const float dw = fminf(__fdiv_rn(d.z, ww), xform_clip), dh = fminf(__fdiv_rn(d.w, wh), xform_clip);
    const float pcx = __fadd_rn(__fmul_rn(dx, w), cx);
    const float pcy = __fadd_rn(__fmul_rn(dy, h), cy);
    const float pw = __fmul_rn(expf(dw), w);
    const float ph = __fmul_rn(expf(dh), h);
    float4 o;
    o.x = __fsub_rn(pcx, __fmul_rn(0.5f, pw));
    o.y = __fsub_rn(pcy, __fmul_rn(0.5f, ph));
    o.z = __fsub_rn(__fadd_rn(pcx, __fmul_rn(0.5f, pw)), 1.f);
    o.w = __fsub_rn(__fadd_rn(pcy, __fmul_rn(0.5f, ph)), 1.f);
    bool ok = true;
    if (image_wh != nullptr) {
      const int img = (int)(i / per_image);
      const float W1 = __fsub_rn(image_wh[2 * img], 1.f), H1 = __fsub_rn(image_wh[2 * img + 1], 1.f);
      o.x = fminf(fmaxf(o.x, 0.f), W1);
      o.y = fminf(fmaxf(o.y, 0.f), H1);
      o.z = fminf(fmaxf(o.z, 0.f), W1);
      o.w = fminf(fmaxf(o.w, 0.f), H1);
    }
    if (min_size > 0.f)
      ok = (__fadd_rn(__fsub_rn(o.z, o.x), 1.f) >= min_size) && (__fadd_rn(__fsub_rn(o.w, o.y), 1.f) >= min_size);
    boxes[i] = o;
    if (scores_nms) scores_nms[i] = ok ? __ldg(scores + i) : -1.f;
    if (valid) valid[i] = ok ? 1 : 0;
  }
}

}  // namespace

extern "C" int mmt_decode_clip_boxes_f32(const float* deltas, const float* anchors, const float* scores,
                                         const float* image_wh, int n_images, int per_image, float wx, float wy,
                                         float ww, float wh, float xform_clip, float min_size, float* boxes,
                                         float* scores_nms, unsigned char* valid, void* stream_) {
  MMT_REQUIRE(n_images >= 0 && per_image >= 0, "decode_clip_boxes: negative sizes");
  const long long total = (long long)n_images * per_image;
  if (total == 0) return MMT_OK;
  MMT_REQUIRE(deltas && anchors && boxes, "decode_clip_boxes: null pointer");
  MMT_REQUIRE(scores_nms == nullptr || scores != nullptr, "decode_clip_boxes: scores_nms requested without scores");
  MMT_REQUIRE(wx != 0.f && wy != 0.f && ww != 0.f && wh != 0.f, "decode_clip_boxes: zero coder weight");
  MMT_REQUIRE((((uintptr_t)deltas | (uintptr_t)anchors | (uintptr_t)boxes) & 15) == 0,
              "decode_clip_boxes: box arrays must be 16-byte aligned");
  long long blocks = (total + 255) / 256;
  if (blocks > mmt::kNumSMs * 8) blocks = mmt::kNumSMs * 8;
  decode_clip_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream_>>>(
      reinterpret_cast<const float4*>(deltas), reinterpret_cast<const float4*>(anchors), scores, image_wh, per_image,
      total, wx, wy, ww, wh, xform_clip, min_size, reinterpret_cast<float4*>(boxes), scores_nms, valid);
  return mmt::check_launch("decode_clip_boxes");
}
