// Mask-target rasterisation on the device.
//
// Replaces the per-positive-ROI HOST loop of maskrcnn_benchmark/modeling/roi_heads/mask_head/loss.py:37-75
// (`project_masks_on_boxes`): Polygons.crop (structures/segmentation_mask.py:94-108, fp32 subtract),
// Polygons.resize (:110-125, fp32 multiply by float(M)/float(w)), Polygons.convert("mask") (:127-137)
// = pycocotools frPyObjects -> rleFrPoly (pycoco/maskApi.c:166-206), rleMerge (:53-75, union) and
// rleDecode (:47-51).  The integer/double arithmetic of rleFrPoly is reproduced statement by statement, so
// the M x M target is bit-identical to the reference's.
//
// The RLE built by rleFrPoly is "toggle the value at every sorted boundary index a = x*h + y (column major),
// equal indices cancel", i.e. pixel i is set iff an odd number of boundary indices are <= i.  One CTA per ROI:
// boundary indices are counted into a shared histogram with atomics, a prefix parity gives the polygon's mask,
// the instance's polygons are OR-ed.
#include "common.cuh"

namespace {

constexpr int kMaxM = 56;

struct EdgeWalk {
  int xs, ys, xe, ye, dx, dy, flip;
  double s;
  int steps;  // number of points on this edge: max(dx,dy)+1
};

__device__ __forceinline__ EdgeWalk make_edge(int x0, int y0, int x1, int y1) {
  EdgeWalk e;
  e.xs = x0; e.xe = x1; e.ys = y0; e.ye = y1;
  e.dx = abs(e.xe - e.xs);
  e.dy = abs(e.ys - e.ye);
  e.flip = (e.dx >= e.dy && e.xs > e.xe) || (e.dx < e.dy && e.ys > e.ye);
  if (e.flip) { int t = e.xs; e.xs = e.xe; e.xe = t; t = e.ys; e.ys = e.ye; e.ye = t; }
  e.s = e.dx >= e.dy ? (double)(e.ye - e.ys) / e.dx : (double)(e.xe - e.xs) / e.dy;
  e.steps = max(e.dx, e.dy) + 1;
  return e;
}

__device__ __forceinline__ void edge_point(const EdgeWalk& e, int d, int& u, int& v) {
  if (e.dx >= e.dy) {
    const int t = e.flip ? e.dx - d : d;
    u = t + e.xs;
    v = (int)(e.ys + e.s * t + .5);
  } else {
    const int t = e.flip ? e.dy - d : d;
    v = t + e.ys;
    u = (int)(e.xs + e.s * t + .5);
  }
}

__global__ void __launch_bounds__(256) rasterize_kernel(
    const float* __restrict__ xy, const int* __restrict__ poly_off, const int* __restrict__ inst_off,
    const float* __restrict__ boxes, const int* __restrict__ roi_inst, int M, float* __restrict__ out) {
  __shared__ int cnt[kMaxM * kMaxM + 1];
  __shared__ unsigned char acc[kMaxM * kMaxM];
  const int roi = blockIdx.x;
  const int tid = threadIdx.x;
  const int hw = M * M;
  const float bx0 = boxes[roi * 4 + 0], by0 = boxes[roi * 4 + 1];
  const float bw = fmaxf(__fsub_rn(boxes[roi * 4 + 2], bx0), 1.f);
  const float bh = fmaxf(__fsub_rn(boxes[roi * 4 + 3], by0), 1.f);
  // ratios are python doubles float(M)/float(w); the tensor multiply happens in fp32
  const float rw = (float)((double)M / (double)bw), rh = (float)((double)M / (double)bh);
  const double scale = 5.0;
  const int inst = roi_inst[roi];
  for (int i = tid; i < hw; i += blockDim.x) acc[i] = 0;
  const int p_begin = inst_off[inst], p_end = inst_off[inst + 1];
  for (int pi = p_begin; pi < p_end; ++pi) {
    const int v0 = poly_off[pi], k = poly_off[pi + 1] - v0;
    for (int i = tid; i <= hw; i += blockDim.x) cnt[i] = 0;
    __syncthreads();
    auto vertex = [&](int j, int& X, int& Y) {
      j = j % k;
      const float px = __fmul_rn(__fsub_rn(xy[2 * (v0 + j)], bx0), rw);
      const float py = __fmul_rn(__fsub_rn(xy[2 * (v0 + j) + 1], by0), rh);
      X = (int)(scale * (double)px + .5);
      Y = (int)(scale * (double)py + .5);
    };
    for (int e = 0; e < k; ++e) {
      int x0, y0, x1, y1;
      vertex(e, x0, y0);
      vertex(e + 1, x1, y1);
      const EdgeWalk ew = make_edge(x0, y0, x1, y1);
      for (int d = tid; d < ew.steps; d += blockDim.x) {
        int u, v, up, vp;
        edge_point(ew, d, u, v);
        if (d > 0) {
          edge_point(ew, d - 1, up, vp);
        } else {
          if (e == 0) continue;  // the very first point of the polygon has no predecessor
          int px0, py0;
          vertex(e - 1, px0, py0);
          const EdgeWalk pe = make_edge(px0, py0, x0, y0);
          edge_point(pe, pe.steps - 1, up, vp);
        }
        if (u == up) continue;
        double xd = (double)(u < up ? u : u - 1);
        xd = (xd + .5) / scale - .5;
        if (floor(xd) != xd || xd < 0 || xd > M - 1) continue;
        double yd = (double)(v < vp ? v : vp);
        yd = (yd + .5) / scale - .5;
        if (yd < 0) yd = 0; else if (yd > M) yd = M;
        yd = ceil(yd);
        atomicAdd(&cnt[(int)xd * M + (int)yd], 1);
      }
    }
    __syncthreads();
    if (tid < 32) {  // prefix parity over hw entries by one warp
      int carry = 0;
      for (int base = 0; base < hw; base += 32) {
        const int i = base + tid;
        int v = (i < hw) ? (cnt[i] & 1) : 0;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const int n = __shfl_up_sync(0xffffffffu, v, o);
          if (tid >= o) v ^= n;
        }
        v ^= carry;
        if (i < hw && v) acc[i] = 1;
        carry = __shfl_sync(0xffffffffu, v, 31);
      }
    }
    __syncthreads();
  }
  // column-major RLE index i = x*h + y  ->  out[roi][y][x]
  for (int i = tid; i < hw; i += blockDim.x) {
    const int y = i / M, x = i % M;
    out[(long long)roi * hw + i] = acc[x * M + y] ? 1.f : 0.f;
  }
}

}  // namespace

extern "C" int mmt_rasterize_polygons_f32(const float* poly_xy, const int* poly_offsets,
                                          const int* inst_poly_offsets, const float* roi_boxes,
                                          const int* roi_instance, int n_rois, int mask_size,
                                          float* masks_out, void* stream_) {
  MMT_REQUIRE(n_rois >= 0 && mask_size > 0 && mask_size <= kMaxM, "rasterize: mask_size must be in [1,%d]", kMaxM);
  if (n_rois == 0) return MMT_OK;
  MMT_REQUIRE(poly_xy && poly_offsets && inst_poly_offsets && roi_boxes && roi_instance && masks_out,
              "rasterize: null pointer");
  rasterize_kernel<<<n_rois, 256, 0, (cudaStream_t)stream_>>>(poly_xy, poly_offsets, inst_poly_offsets, roi_boxes,
                                                             roi_instance, mask_size, masks_out);
  return mmt::check_launch("rasterize_polygons");
}

// ------------------------------------------------------------------------------------------------
// Teacher mask paste-and-sum.
//
// Replaces the CPU loop of maskrcnn_benchmark/modeling/roi_heads/mask_head/inference.py:169-231
// (`Masker.forward_single_image` -> `paste_mask_in_image` per detection: pad the MxM probability map by 1,
// expand the box by (M+2)/M, truncate to int32, bilinear-resize (align_corners=False) to the box, threshold,
// paste into an im_h x im_w uint8 canvas) followed by `.sum(0)` over detections
// (modeling/detector/generalized_rcnn.py:129-132).  One thread per pasted pixel; the only consumer is the
// integral foreground map, so the canvases are never materialised.
namespace {

__global__ void __launch_bounds__(256) paste_sum_kernel(const float* __restrict__ probs,
                                                        const float* __restrict__ boxes,
                                                        const int* __restrict__ img_idx, int M, int im_h,
                                                        int im_w, float thresh, float* __restrict__ out) {
  const int d = blockIdx.y;
  const float* b = boxes + d * 4;
  // expand_boxes (inference.py:94-110), fp32 like torch; scale = float(M + 2) / M is a python double
  const float scale = (float)((double)(M + 2) / (double)M);
  float w_half = __fmul_rn(__fsub_rn(b[2], b[0]), .5f), h_half = __fmul_rn(__fsub_rn(b[3], b[1]), .5f);
  const float x_c = __fmul_rn(__fadd_rn(b[2], b[0]), .5f), y_c = __fmul_rn(__fadd_rn(b[3], b[1]), .5f);
  w_half = __fmul_rn(w_half, scale);
  h_half = __fmul_rn(h_half, scale);
  const int bx0 = (int)__fsub_rn(x_c, w_half), bx1 = (int)__fadd_rn(x_c, w_half);
  const int by0 = (int)__fsub_rn(y_c, h_half), by1 = (int)__fadd_rn(y_c, h_half);
  const int w = max(bx1 - bx0 + 1, 1), h = max(by1 - by0 + 1, 1);
  const int x_0 = max(bx0, 0), x_1 = min(bx1 + 1, im_w), y_0 = max(by0, 0), y_1 = min(by1 + 1, im_h);
  const int rw = x_1 - x_0, rh = y_1 - y_0;
  if (rw <= 0 || rh <= 0) return;
  const int P = M + 2;
  const float sx = (float)P / (float)w, sy = (float)P / (float)h;  // area_pixel_compute_scale
  const float* src = probs + (long long)d * M * M;
  float* dst = out + (long long)img_idx[d] * im_h * im_w;
  auto padded = [&](int yy, int xx) -> float {
    return (yy >= 1 && yy <= M && xx >= 1 && xx <= M) ? __ldg(src + (yy - 1) * M + (xx - 1)) : 0.f;
  };
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < rw * rh; i += gridDim.x * blockDim.x) {
    const int px = x_0 + i % rw, py = y_0 + i / rw;
    const int ox = px - bx0, oy = py - by0;  // pixel of the resized mask
    float fx = sx * ((float)ox + .5f) - .5f, fy = sy * ((float)oy + .5f) - .5f;
    fx = fx < 0.f ? 0.f : fx;
    fy = fy < 0.f ? 0.f : fy;
    const int ix0 = (int)fx, iy0 = (int)fy;
    const int ix1 = ix0 + (ix0 < P - 1 ? 1 : 0), iy1 = iy0 + (iy0 < P - 1 ? 1 : 0);
    const float lx1 = fx - (float)ix0, ly1 = fy - (float)iy0, lx0 = 1.f - lx1, ly0 = 1.f - ly1;
    const float v = ly0 * (lx0 * padded(iy0, ix0) + lx1 * padded(iy0, ix1)) +
                    ly1 * (lx0 * padded(iy1, ix0) + lx1 * padded(iy1, ix1));
    if (v > thresh) atomicAdd(dst + (long long)py * im_w + px, 1.f);
  }
}

}  // namespace

extern "C" int mmt_paste_masks_sum_f32(const float* mask_probs, const float* boxes, const int* image_index,
                                       int n_dets, int mask_size, int n_images, int im_h, int im_w,
                                       float threshold, float* integral_out, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  MMT_REQUIRE(n_dets >= 0 && mask_size > 0 && n_images > 0 && im_h > 0 && im_w > 0, "paste_masks: bad sizes");
  MMT_REQUIRE(integral_out != nullptr, "paste_masks: null output");
  MMT_CUDA(cudaMemsetAsync(integral_out, 0, sizeof(float) * (size_t)n_images * im_h * im_w, stream));
  if (n_dets == 0) return MMT_OK;
  MMT_REQUIRE(mask_probs && boxes && image_index, "paste_masks: null pointer");
  MMT_REQUIRE(n_dets <= 65535, "paste_masks: too many detections");
  paste_sum_kernel<<<dim3(16, n_dets), 256, 0, stream>>>(mask_probs, boxes, image_index, mask_size, im_h, im_w,
                                                        threshold, integral_out);
  return mmt::check_launch("paste_masks_sum");
}
