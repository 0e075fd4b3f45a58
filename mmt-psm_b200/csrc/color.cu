// Colour views of unlabeled crops on the device, bit-exact with the Pillow / torchvision arithmetic the
// reference's host pipeline runs per view (maskrcnn_benchmark/data/transforms/build.py:27-35 `transpart2`,
// transforms.py:84-206; data/datasets/Pap.py:818-831):
//   AdjustBrightness  ImageEnhance.Brightness = Image.blend(black, img, f): float32 in1 + f * (in2 - in1), clip, truncate
//   AdjustContrast    blend(grey(int(mean(L) + 0.5)), img, f), L = (19595 R + 38470 G + 7471 B + 0x8000) >> 16
//   AdjustHue         uint8 wrap-around add on the H plane of Pillow's RGB <-> HSV (Convert.c: colorsys in float with
//                     double-promoted constants; h truncated, p / q / t rounded)
//   RandomErasing     rectangles overwritten with uint8 noise (the host draws rectangles and fill: same random stream)
//   ToTensor + Normalize(to_bgr255): float32 (u / 255) * 255 - PIXEL_MEAN, channels B, G, R
// Every float / double operation is written with round-to-nearest intrinsics so that no FMA contraction changes a bit.
#include <algorithm>

#include "common.cuh"

namespace {

__device__ __forceinline__ unsigned char blend_u8(float degenerate, unsigned char v, float factor) {
  const float t = __fadd_rn(degenerate, __fmul_rn(factor, __fsub_rn((float)v, degenerate)));
  if (t <= 0.f) return 0;
  if (t >= 255.f) return 255;
  return (unsigned char)t;  // truncation, t > 0
}
__device__ __forceinline__ int luma(int r, int g, int b) { return (r * 19595 + g * 38470 + b * 7471 + 0x8000) >> 16; }

__global__ void __launch_bounds__(256) luma_sums_kernel(const uchar3* __restrict__ img, const float* __restrict__ brightness,
                                                        int B, long long pixels, unsigned long long* __restrict__ sums) {
  __shared__ unsigned long long s_red[8];
  const int vb = blockIdx.y, b = vb % B;
  const float f = brightness[vb];
  const uchar3* src = img + (long long)b * pixels;
  unsigned long long acc = 0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < pixels; i += (long long)gridDim.x * blockDim.x) {
    const uchar3 p = src[i];
    acc += (unsigned long long)luma(blend_u8(0.f, p.x, f), blend_u8(0.f, p.y, f), blend_u8(0.f, p.z, f));
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned long long t = 0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += s_red[w];
    atomicAdd(sums + vb, t);
  }
}

// Pillow Convert.c rgb2hsv_row
__device__ __forceinline__ void rgb2hsv(int r, int g, int b, int& uh, int& us, int& uv) {
  const int maxc = max(r, max(g, b)), minc = min(r, min(g, b));
  uv = maxc;
  if (minc == maxc) { uh = 0; us = 0; return; }
  const float cr = (float)(maxc - minc);
  const float s = __fdiv_rn(cr, (float)maxc);
  const float rc = __fdiv_rn((float)(maxc - r), cr), gc = __fdiv_rn((float)(maxc - g), cr), bc = __fdiv_rn((float)(maxc - b), cr);
  float h;
  if (r == maxc) h = __fsub_rn(bc, gc);
  else if (g == maxc) h = __double2float_rn(__dsub_rn(__dadd_rn(2.0, (double)rc), (double)bc));
  else h = __double2float_rn(__dsub_rn(__dadd_rn(4.0, (double)gc), (double)rc));
  double hd = __dadd_rn(__ddiv_rn((double)h, 6.0), 1.0);   // in [5/6, 11/6]: fmod(., 1) is an exact subtraction
  if (hd >= 1.0) hd = __dsub_rn(hd, 1.0);
  h = __double2float_rn(hd);
  uh = min(max((int)__dmul_rn((double)h, 255.0), 0), 255);
  us = min(max((int)__dmul_rn((double)s, 255.0), 0), 255);
}
// Pillow Convert.c hsv2rgb
__device__ __forceinline__ void hsv2rgb(int h, int s, int v, int& r, int& g, int& b) {
  if (s == 0) { r = g = b = v; return; }
  const double hf = __ddiv_rn(__dmul_rn((double)(float)h, 6.0), 255.0);
  const int i = (int)floor(hf);
  const double f = (double)__double2float_rn(__dsub_rn(hf, (double)(float)i));
  const double fs = (double)__double2float_rn(__ddiv_rn((double)(float)s, 255.0));
  const double vf = (double)(float)v;
  const int p = min(max((int)floor(__dadd_rn(__dmul_rn(vf, __dsub_rn(1.0, fs)), 0.5)), 0), 255);
  const int q = min(max((int)floor(__dadd_rn(__dmul_rn(vf, __dsub_rn(1.0, __dmul_rn(fs, f))), 0.5)), 0), 255);
  const int t = min(max((int)floor(__dadd_rn(__dmul_rn(vf, __dsub_rn(1.0, __dmul_rn(fs, __dsub_rn(1.0, f)))), 0.5)), 0), 255);
  switch (i % 6) {
    case 0: r = v; g = t; b = p; break;
    case 1: r = q; g = v; b = p; break;
    case 2: r = p; g = v; b = t; break;
    case 3: r = p; g = q; b = v; break;
    case 4: r = t; g = p; b = v; break;
    default: r = v; g = p; b = q; break;
  }
}

struct EraseRect { int vb, top, left, h, w; long long fill; };  // fill: offset of the h*w*3 uint8 block in the pool

__global__ void __launch_bounds__(256) color_views_kernel(const uchar3* __restrict__ img, const float* __restrict__ brightness,
                                                          const float* __restrict__ contrast,
                                                          const int* __restrict__ hue_shift,
                                                          const unsigned long long* __restrict__ luma_sums, int B, int H,
                                                          int W, const EraseRect* __restrict__ rects, int n_rects,
                                                          const unsigned char* __restrict__ fill_pool, float mean_b,
                                                          float mean_g, float mean_r, float* __restrict__ out) {
  const long long pixels = (long long)H * W;
  const int vb = blockIdx.y, b = vb % B;
  const float fb = brightness[vb], fc = contrast[vb];
  const int shift = hue_shift[vb];
  const float pivot = (float)(long long)(__dadd_rn(__ddiv_rn((double)luma_sums[vb], (double)pixels), 0.5));
  const uchar3* src = img + (long long)b * pixels;
  float* dst = out + (long long)vb * 3 * pixels;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < pixels; i += (long long)gridDim.x * blockDim.x) {
    const uchar3 p = src[i];
    int r = blend_u8(pivot, blend_u8(0.f, p.x, fb), fc);
    int g = blend_u8(pivot, blend_u8(0.f, p.y, fb), fc);
    int bl = blend_u8(pivot, blend_u8(0.f, p.z, fb), fc);
    int uh, us, uv;
    rgb2hsv(r, g, bl, uh, us, uv);
    uh = (uh + shift) & 255;
    hsv2rgb(uh, us, uv, r, g, bl);
    const int y = (int)(i / W), x = (int)(i - (long long)y * W);
    for (int k = 0; k < n_rects; ++k) {   // applied in order: a later rectangle overwrites an earlier one
      const EraseRect e = rects[k];
      if (e.vb == vb && y >= e.top && y < e.top + e.h && x >= e.left && x < e.left + e.w) {
        const unsigned char* f = fill_pool + e.fill + ((long long)(y - e.top) * e.w + (x - e.left)) * 3;
        r = f[0]; g = f[1]; bl = f[2];
      }
    }
    // ToTensor (/255) -> [2,1,0] * 255 -> - mean
    dst[i] = __fsub_rn(__fmul_rn(__fdiv_rn((float)bl, 255.f), 255.f), mean_b);
    dst[pixels + i] = __fsub_rn(__fmul_rn(__fdiv_rn((float)g, 255.f), 255.f), mean_g);
    dst[2 * pixels + i] = __fsub_rn(__fmul_rn(__fdiv_rn((float)r, 255.f), 255.f), mean_r);
  }
}

}  // namespace

extern "C" int mmt_color_luma_sums_u8(const unsigned char* images_hwc, const float* brightness, int n_views, int batch,
                                      int height, int width, unsigned long long* luma_sums_out, void* stream_) {
  MMT_REQUIRE(n_views > 0 && batch > 0 && height > 0 && width > 0, "color_luma_sums: bad sizes");
  MMT_REQUIRE(images_hwc && brightness && luma_sums_out, "color_luma_sums: null pointer");
  cudaStream_t stream = (cudaStream_t)stream_;
  const long long pixels = (long long)height * width;
  MMT_CUDA(cudaMemsetAsync(luma_sums_out, 0, sizeof(unsigned long long) * (size_t)n_views * batch, stream));
  const int gx = (int)std::min<long long>((pixels + 255) / 256, 4 * mmt::kNumSMs);
  luma_sums_kernel<<<dim3(gx, n_views * batch), 256, 0, stream>>>(reinterpret_cast<const uchar3*>(images_hwc), brightness,
                                                                   batch, pixels, luma_sums_out);
  return mmt::check_launch("color_luma_sums");
}

extern "C" int mmt_color_views_u8(const unsigned char* images_hwc, const float* brightness, const float* contrast,
                                  const int* hue_shift, const unsigned long long* luma_sums, int n_views, int batch,
                                  int height, int width, const void* erase_rects, int n_rects,
                                  const unsigned char* fill_pool, float mean_b, float mean_g, float mean_r,
                                  float* views_out, void* stream_) {
  MMT_REQUIRE(n_views > 0 && batch > 0 && height > 0 && width > 0 && n_rects >= 0, "color_views: bad sizes");
  MMT_REQUIRE(images_hwc && brightness && contrast && hue_shift && luma_sums && views_out, "color_views: null pointer");
  MMT_REQUIRE(n_rects == 0 || (erase_rects && fill_pool), "color_views: rectangles without table / fill pool");
  const long long pixels = (long long)height * width;
  const int gx = (int)std::min<long long>((pixels + 255) / 256, 4 * mmt::kNumSMs);
  color_views_kernel<<<dim3(gx, n_views * batch), 256, 0, (cudaStream_t)stream_>>>(
      reinterpret_cast<const uchar3*>(images_hwc), brightness, contrast, hue_shift, luma_sums, batch, height, width,
      reinterpret_cast<const EraseRect*>(erase_rects), n_rects, fill_pool, mean_b, mean_g, mean_r, views_out);
  return mmt::check_launch("color_views");
}
