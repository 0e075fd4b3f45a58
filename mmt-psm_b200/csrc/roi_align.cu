// ROIAlign forward/backward (+ thin ROIPool) for sm_100a.
//
// Replaces maskrcnn_benchmark/csrc/cuda/ROIAlign_cuda.cu:65-122 (RoIAlignForward), :178-254
// (RoIAlignBackwardFeature) and csrc/cuda/ROIPool_cuda.cu:17-108.  Sampling geometry follows the
// reference statement by statement (ROIAlign_cuda.cu:16-63,81-118 == cpu/ROIAlign_cpu.cpp:17-120);
// all coordinate / weight arithmetic uses explicitly rounded fp32 ops so the forward result is
// bit-identical to the reference CPU kernel (no FMA contraction).
//
// B200 mapping: tensors are addressed through element strides.  When the channel stride is 1
// (NHWC / channels_last — what the conv engine of this library emits) one thread owns 4
// consecutive channels, a warp covers 128 consecutive channels of ONE sampling point: every tap is
// a 512-byte fully coalesced float4 load and the warp is branch-uniform.  The backward scatters
// with `red.global.add.v4.f32` (one 16-byte reduction instead of four atomics) into a buffer the
// caller zero-fills ONCE for all levels.  For NCHW inputs (the reference's layout) the kernel
// falls back to the reference's pw-fastest thread order with scalar accesses.
#include "common.cuh"

namespace {

struct RoiGeom {
  int batch_ind;
  float start_w, start_h, bin_w, bin_h;
  int grid_h, grid_w;
  float count;
};

__device__ __forceinline__ RoiGeom roi_geometry(const float* __restrict__ roi, float scale,
                                                int pooled_h, int pooled_w, int sampling_ratio) {
  RoiGeom g;
  g.batch_ind = (int)roi[0];
  g.start_w = __fmul_rn(roi[1], scale);
  g.start_h = __fmul_rn(roi[2], scale);
  const float end_w = __fmul_rn(roi[3], scale);
  const float end_h = __fmul_rn(roi[4], scale);
  const float roi_w = fmaxf(__fsub_rn(end_w, g.start_w), 1.f);
  const float roi_h = fmaxf(__fsub_rn(end_h, g.start_h), 1.f);
  g.bin_h = __fdiv_rn(roi_h, (float)pooled_h);
  g.bin_w = __fdiv_rn(roi_w, (float)pooled_w);
  g.grid_h = sampling_ratio > 0 ? sampling_ratio : (int)ceilf(__fdiv_rn(roi_h, (float)pooled_h));
  g.grid_w = sampling_ratio > 0 ? sampling_ratio : (int)ceilf(__fdiv_rn(roi_w, (float)pooled_w));
  g.count = (float)(g.grid_h * g.grid_w);
  return g;
}

__device__ __forceinline__ float sample_coord(float start, int p, float bin, int i, int grid) {
  // start + p*bin + (i+.5f)*bin/grid, left to right, each op rounded (ROIAlign_cuda.cu:108,111)
  const float a = __fadd_rn(start, __fmul_rn((float)p, bin));
  const float b = __fdiv_rn(__fmul_rn((float)i + .5f, bin), (float)grid);
  return __fadd_rn(a, b);
}

struct Taps {
  int y_low, y_high, x_low, x_high;
  float w1, w2, w3, w4;
  bool valid;
};

__device__ __forceinline__ Taps bilinear_taps(float y, float x, int height, int width) {
  Taps t;
  if (y < -1.0f || y > (float)height || x < -1.0f || x > (float)width) {
    t.valid = false;
    t.y_low = t.y_high = t.x_low = t.x_high = 0;
    t.w1 = t.w2 = t.w3 = t.w4 = 0.f;
    return t;
  }
  t.valid = true;
  if (y <= 0.f) y = 0.f;
  if (x <= 0.f) x = 0.f;
  t.y_low = (int)y;
  t.x_low = (int)x;
  if (t.y_low >= height - 1) {
    t.y_high = t.y_low = height - 1;
    y = (float)t.y_low;
  } else {
    t.y_high = t.y_low + 1;
  }
  if (t.x_low >= width - 1) {
    t.x_high = t.x_low = width - 1;
    x = (float)t.x_low;
  } else {
    t.x_high = t.x_low + 1;
  }
  const float ly = __fsub_rn(y, (float)t.y_low), lx = __fsub_rn(x, (float)t.x_low);
  const float hy = __fsub_rn(1.f, ly), hx = __fsub_rn(1.f, lx);
  t.w1 = __fmul_rn(hy, hx);
  t.w2 = __fmul_rn(hy, lx);
  t.w3 = __fmul_rn(ly, hx);
  t.w4 = __fmul_rn(ly, lx);
  return t;
}

// w1*v1 + w2*v2 + w3*v3 + w4*v4, left to right, unfused (ROIAlign_cuda.cu:59)
__device__ __forceinline__ float blend(const Taps& t, float v1, float v2, float v3, float v4) {
  float r = __fmul_rn(t.w1, v1);
  r = __fadd_rn(r, __fmul_rn(t.w2, v2));
  r = __fadd_rn(r, __fmul_rn(t.w3, v3));
  r = __fadd_rn(r, __fmul_rn(t.w4, v4));
  return r;
}

// Decompose a flat work index.  CH_FAST: channel-vector fastest (NHWC); else pw fastest (NCHW).
template <bool CH_FAST>
__device__ __forceinline__ void decompose(long long idx, int cvecs, int pooled_h, int pooled_w,
                                          int& n, int& cv, int& ph, int& pw) {
  if (CH_FAST) {
    cv = (int)(idx % cvecs);
    idx /= cvecs;
    pw = (int)(idx % pooled_w);
    idx /= pooled_w;
    ph = (int)(idx % pooled_h);
    n = (int)(idx / pooled_h);
  } else {
    pw = (int)(idx % pooled_w);
    idx /= pooled_w;
    ph = (int)(idx % pooled_h);
    idx /= pooled_h;
    cv = (int)(idx % cvecs);
    n = (int)(idx / cvecs);
  }
}

template <int VEC, bool CH_FAST>
__global__ void __launch_bounds__(256) roi_align_fwd_kernel(
    const float* __restrict__ in, mmt_strides4 is, int channels, int height, int width,
    const float* __restrict__ rois, float scale, int pooled_h, int pooled_w, int sampling_ratio,
    float* __restrict__ out, mmt_strides4 os, long long total) {
  const int cvecs = channels / VEC;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    int n, cv, ph, pw;
    decompose<CH_FAST>(idx, cvecs, pooled_h, pooled_w, n, cv, ph, pw);
    const int c = cv * VEC;
    const RoiGeom g = roi_geometry(rois + (long long)n * 5, scale, pooled_h, pooled_w, sampling_ratio);
    const float* base = in + g.batch_ind * is.n + c * is.c;
    float acc[VEC];
#pragma unroll
    for (int v = 0; v < VEC; ++v) acc[v] = 0.f;
    for (int iy = 0; iy < g.grid_h; ++iy) {
      const float y = sample_coord(g.start_h, ph, g.bin_h, iy, g.grid_h);
      for (int ix = 0; ix < g.grid_w; ++ix) {
        const float x = sample_coord(g.start_w, pw, g.bin_w, ix, g.grid_w);
        const Taps t = bilinear_taps(y, x, height, width);
        if (!t.valid) continue;  // contributes +0
        const float* p1 = base + t.y_low * is.h + t.x_low * is.w;
        const float* p2 = base + t.y_low * is.h + t.x_high * is.w;
        const float* p3 = base + t.y_high * is.h + t.x_low * is.w;
        const float* p4 = base + t.y_high * is.h + t.x_high * is.w;
        if (VEC == 4) {
          const float4 a = __ldg(reinterpret_cast<const float4*>(p1));
          const float4 b = __ldg(reinterpret_cast<const float4*>(p2));
          const float4 cc = __ldg(reinterpret_cast<const float4*>(p3));
          const float4 d = __ldg(reinterpret_cast<const float4*>(p4));
          acc[0] = __fadd_rn(acc[0], blend(t, a.x, b.x, cc.x, d.x));
          acc[1 % VEC] = __fadd_rn(acc[1 % VEC], blend(t, a.y, b.y, cc.y, d.y));
          acc[2 % VEC] = __fadd_rn(acc[2 % VEC], blend(t, a.z, b.z, cc.z, d.z));
          acc[3 % VEC] = __fadd_rn(acc[3 % VEC], blend(t, a.w, b.w, cc.w, d.w));
        } else {
          acc[0] = __fadd_rn(acc[0], blend(t, __ldg(p1), __ldg(p2), __ldg(p3), __ldg(p4)));
        }
      }
    }
    float* o = out + n * os.n + c * os.c + ph * os.h + pw * os.w;
    if (VEC == 4 && os.c == 1) {
      float4 r;
      r.x = __fdiv_rn(acc[0], g.count);
      r.y = __fdiv_rn(acc[1 % VEC], g.count);
      r.z = __fdiv_rn(acc[2 % VEC], g.count);
      r.w = __fdiv_rn(acc[3 % VEC], g.count);
      *reinterpret_cast<float4*>(o) = r;
    } else {
#pragma unroll
      for (int v = 0; v < VEC; ++v) o[v * os.c] = __fdiv_rn(acc[v], g.count);
    }
  }
}

template <int VEC, bool CH_FAST>
__global__ void __launch_bounds__(256) roi_align_bwd_kernel(
    const float* __restrict__ gout, mmt_strides4 gs, const float* __restrict__ rois, float scale,
    int pooled_h, int pooled_w, int sampling_ratio, float* __restrict__ gin, mmt_strides4 is,
    int channels, int height, int width, long long total) {
  const int cvecs = channels / VEC;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    int n, cv, ph, pw;
    decompose<CH_FAST>(idx, cvecs, pooled_h, pooled_w, n, cv, ph, pw);
    const int c = cv * VEC;
    const RoiGeom g = roi_geometry(rois + (long long)n * 5, scale, pooled_h, pooled_w, sampling_ratio);
    float go[VEC];
    const float* gp = gout + n * gs.n + c * gs.c + ph * gs.h + pw * gs.w;
    if (VEC == 4 && gs.c == 1) {
      const float4 v = __ldg(reinterpret_cast<const float4*>(gp));
      go[0] = v.x; go[1 % VEC] = v.y; go[2 % VEC] = v.z; go[3 % VEC] = v.w;
    } else {
#pragma unroll
      for (int v = 0; v < VEC; ++v) go[v] = __ldg(gp + v * gs.c);
    }
    float* base = gin + g.batch_ind * is.n + c * is.c;
    for (int iy = 0; iy < g.grid_h; ++iy) {
      const float y = sample_coord(g.start_h, ph, g.bin_h, iy, g.grid_h);
      for (int ix = 0; ix < g.grid_w; ++ix) {
        const float x = sample_coord(g.start_w, pw, g.bin_w, ix, g.grid_w);
        const Taps t = bilinear_taps(y, x, height, width);
        if (!t.valid) continue;
        float* p1 = base + t.y_low * is.h + t.x_low * is.w;
        float* p2 = base + t.y_low * is.h + t.x_high * is.w;
        float* p3 = base + t.y_high * is.h + t.x_low * is.w;
        float* p4 = base + t.y_high * is.h + t.x_high * is.w;
        // g_k = top_diff * w_k / count  (ROIAlign_cuda.cu:238-241)
        if (VEC == 4) {
          float4 a, b, cc, d;
          a.x = __fdiv_rn(__fmul_rn(go[0], t.w1), g.count); a.y = __fdiv_rn(__fmul_rn(go[1 % VEC], t.w1), g.count);
          a.z = __fdiv_rn(__fmul_rn(go[2 % VEC], t.w1), g.count); a.w = __fdiv_rn(__fmul_rn(go[3 % VEC], t.w1), g.count);
          b.x = __fdiv_rn(__fmul_rn(go[0], t.w2), g.count); b.y = __fdiv_rn(__fmul_rn(go[1 % VEC], t.w2), g.count);
          b.z = __fdiv_rn(__fmul_rn(go[2 % VEC], t.w2), g.count); b.w = __fdiv_rn(__fmul_rn(go[3 % VEC], t.w2), g.count);
          cc.x = __fdiv_rn(__fmul_rn(go[0], t.w3), g.count); cc.y = __fdiv_rn(__fmul_rn(go[1 % VEC], t.w3), g.count);
          cc.z = __fdiv_rn(__fmul_rn(go[2 % VEC], t.w3), g.count); cc.w = __fdiv_rn(__fmul_rn(go[3 % VEC], t.w3), g.count);
          d.x = __fdiv_rn(__fmul_rn(go[0], t.w4), g.count); d.y = __fdiv_rn(__fmul_rn(go[1 % VEC], t.w4), g.count);
          d.z = __fdiv_rn(__fmul_rn(go[2 % VEC], t.w4), g.count); d.w = __fdiv_rn(__fmul_rn(go[3 % VEC], t.w4), g.count);
          mmt::red_add_v4(p1, a);
          mmt::red_add_v4(p2, b);
          mmt::red_add_v4(p3, cc);
          mmt::red_add_v4(p4, d);
        } else {
          atomicAdd(p1, __fdiv_rn(__fmul_rn(go[0], t.w1), g.count));
          atomicAdd(p2, __fdiv_rn(__fmul_rn(go[0], t.w2), g.count));
          atomicAdd(p3, __fdiv_rn(__fmul_rn(go[0], t.w3), g.count));
          atomicAdd(p4, __fdiv_rn(__fmul_rn(go[0], t.w4), g.count));
        }
      }
    }
  }
}

// ---------------------------------------------------------------- fused FPN pooler (all levels)
// Replaces the per-level loop of maskrcnn_benchmark/modeling/poolers.py:91-121 (4 nonzero syncs, 4
// ROIAlign launches and 4 index_put scatters per call): each ROI carries its level id (computed by
// the caller with the reference's LevelMapper formula, poolers.py:31-42) and samples that level's
// NHWC map; the result lands directly at row n of the [n_rois, ph, pw, C] output.
constexpr int kMaxLevels = 8;
struct LevelTab {
  const float* feat[kMaxLevels];
  float* grad[kMaxLevels];
  int h[kMaxLevels], w[kMaxLevels];
  float scale[kMaxLevels];
};

__global__ void __launch_bounds__(256) roi_align_fpn_fwd_kernel(
    LevelTab tab, int channels, const float* __restrict__ rois, const int* __restrict__ levels,
    int pooled_h, int pooled_w, int sampling_ratio, float* __restrict__ out, long long total) {
  const int cvecs = channels >> 2;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    int n, cv, ph, pw;
    decompose<true>(idx, cvecs, pooled_h, pooled_w, n, cv, ph, pw);
    const int l = levels[n];
    const int height = tab.h[l], width = tab.w[l];
    const RoiGeom g = roi_geometry(rois + (long long)n * 5, tab.scale[l], pooled_h, pooled_w, sampling_ratio);
    const long long sw = channels, sh = (long long)width * channels, sn = (long long)height * sh;
    const float* base = tab.feat[l] + g.batch_ind * sn + cv * 4;
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
    for (int iy = 0; iy < g.grid_h; ++iy) {
      const float y = sample_coord(g.start_h, ph, g.bin_h, iy, g.grid_h);
      for (int ix = 0; ix < g.grid_w; ++ix) {
        const float x = sample_coord(g.start_w, pw, g.bin_w, ix, g.grid_w);
        const Taps t = bilinear_taps(y, x, height, width);
        if (!t.valid) continue;
        const float4 a = __ldg(reinterpret_cast<const float4*>(base + t.y_low * sh + t.x_low * sw));
        const float4 b = __ldg(reinterpret_cast<const float4*>(base + t.y_low * sh + t.x_high * sw));
        const float4 c = __ldg(reinterpret_cast<const float4*>(base + t.y_high * sh + t.x_low * sw));
        const float4 d = __ldg(reinterpret_cast<const float4*>(base + t.y_high * sh + t.x_high * sw));
        a0 = __fadd_rn(a0, blend(t, a.x, b.x, c.x, d.x));
        a1 = __fadd_rn(a1, blend(t, a.y, b.y, c.y, d.y));
        a2 = __fadd_rn(a2, blend(t, a.z, b.z, c.z, d.z));
        a3 = __fadd_rn(a3, blend(t, a.w, b.w, c.w, d.w));
      }
    }
    float4 r;
    r.x = __fdiv_rn(a0, g.count); r.y = __fdiv_rn(a1, g.count);
    r.z = __fdiv_rn(a2, g.count); r.w = __fdiv_rn(a3, g.count);
    reinterpret_cast<float4*>(out)[idx] = r;  // idx enumerates [n][ph][pw][cvec] == NHWC order
  }
}

__global__ void __launch_bounds__(256) roi_align_fpn_bwd_kernel(
    LevelTab tab, int channels, const float* __restrict__ gout, const float* __restrict__ rois,
    const int* __restrict__ levels, int pooled_h, int pooled_w, int sampling_ratio, long long total) {
  const int cvecs = channels >> 2;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    int n, cv, ph, pw;
    decompose<true>(idx, cvecs, pooled_h, pooled_w, n, cv, ph, pw);
    const int l = levels[n];
    const int height = tab.h[l], width = tab.w[l];
    const RoiGeom g = roi_geometry(rois + (long long)n * 5, tab.scale[l], pooled_h, pooled_w, sampling_ratio);
    const long long sw = channels, sh = (long long)width * channels, sn = (long long)height * sh;
    float* base = tab.grad[l] + g.batch_ind * sn + cv * 4;
    const float4 go = __ldg(reinterpret_cast<const float4*>(gout) + idx);
    for (int iy = 0; iy < g.grid_h; ++iy) {
      const float y = sample_coord(g.start_h, ph, g.bin_h, iy, g.grid_h);
      for (int ix = 0; ix < g.grid_w; ++ix) {
        const float x = sample_coord(g.start_w, pw, g.bin_w, ix, g.grid_w);
        const Taps t = bilinear_taps(y, x, height, width);
        if (!t.valid) continue;
        const float ws[4] = {t.w1, t.w2, t.w3, t.w4};
        float* ps[4] = {base + t.y_low * sh + t.x_low * sw, base + t.y_low * sh + t.x_high * sw,
                        base + t.y_high * sh + t.x_low * sw, base + t.y_high * sh + t.x_high * sw};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          float4 v;
          v.x = __fdiv_rn(__fmul_rn(go.x, ws[k]), g.count);
          v.y = __fdiv_rn(__fmul_rn(go.y, ws[k]), g.count);
          v.z = __fdiv_rn(__fmul_rn(go.z, ws[k]), g.count);
          v.w = __fdiv_rn(__fmul_rn(go.w, ws[k]), g.count);
          mmt::red_add_v4(ps[k], v);
        }
      }
    }
  }
}


// ---- v2: one CTA per ROI, bilinear tap table in shared memory ---------------------------------------------------------
// The kernels above recompute the ROI geometry and the 4 bilinear taps of every sample in EVERY thread (64 threads per
// bin at 256 channels: ~13 IEEE divisions each) and hand consecutive bins of a ROI to different CTAs, so neighbouring
// bins — whose taps overlap — never share an L1.  Here a CTA owns a whole ROI: its threads first compute each sample's
// taps ONCE (one sample per thread) into shared memory, then sweep the ROI's bins in raster order, a warp covering 32
// consecutive channel-quads of one bin per step (512-byte coalesced tap reads, table entries broadcast from shared
// memory), so a tap row fetched for one bin is still in this SM's L1 for the next bin and the next row.  Same
// arithmetic, same order of additions: the forward result stays bit-identical to the reference.
struct TapEntry {
  int o1, o2, o3, o4;      // pixel offsets (in pixels: y * width + x) of the four taps; o1 < 0 = sample outside the map
  float w1, w2, w3, w4;
};
constexpr int kMaxSamplesV2 = 1024;   // pooled_h * pooled_w * sampling_ratio^2 (14 x 14 x 4 = 784; 7 x 7 x 4 = 196)

__device__ __forceinline__ void build_tap_table(TapEntry* tab, const RoiGeom& g, int pooled_h, int pooled_w, int height,
                                                int width) {
  const int per_bin = g.grid_h * g.grid_w;
  const int n_samples = pooled_h * pooled_w * per_bin;
  for (int s = threadIdx.x; s < n_samples; s += blockDim.x) {
    const int bin = s / per_bin, k = s % per_bin;
    const int ph = bin / pooled_w, pw = bin % pooled_w;
    const int iy = k / g.grid_w, ix = k % g.grid_w;
    const float y = sample_coord(g.start_h, ph, g.bin_h, iy, g.grid_h);
    const float x = sample_coord(g.start_w, pw, g.bin_w, ix, g.grid_w);
    const Taps t = bilinear_taps(y, x, height, width);
    TapEntry e;
    e.o1 = t.valid ? t.y_low * width + t.x_low : -1;
    e.o2 = t.y_low * width + t.x_high;
    e.o3 = t.y_high * width + t.x_low;
    e.o4 = t.y_high * width + t.x_high;
    e.w1 = t.w1; e.w2 = t.w2; e.w3 = t.w3; e.w4 = t.w4;
    tab[s] = e;
  }
}

__global__ void __launch_bounds__(256) roi_align_fpn_fwd2_kernel(
    LevelTab tab, int channels, const float* __restrict__ rois, const int* __restrict__ levels, int pooled_h,
    int pooled_w, int sampling_ratio, float* __restrict__ out) {
  __shared__ TapEntry taps[kMaxSamplesV2];
  const int n = blockIdx.x;
  const int l = levels[n];
  const int height = tab.h[l], width = tab.w[l];
  const RoiGeom g = roi_geometry(rois + (long long)n * 5, tab.scale[l], pooled_h, pooled_w, sampling_ratio);
  build_tap_table(taps, g, pooled_h, pooled_w, height, width);
  __syncthreads();
  const int cvecs = channels >> 2;
  const int per_bin = g.grid_h * g.grid_w;
  const float4* base = reinterpret_cast<const float4*>(tab.feat[l] + (long long)g.batch_ind * height * width * channels);
  float4* o = reinterpret_cast<float4*>(out) + (long long)n * pooled_h * pooled_w * cvecs;
  const int total = pooled_h * pooled_w * cvecs;
  for (int i = threadIdx.x; i < total; i += blockDim.x) {
    const int cv = i % cvecs, bin = i / cvecs;
    const TapEntry* e = taps + bin * per_bin;
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
    for (int k = 0; k < per_bin; ++k) {   // iy outer, ix inner: the reference's accumulation order
      const TapEntry t = e[k];
      if (t.o1 < 0) continue;  // contributes +0
      const float4 a = __ldg(base + (long long)t.o1 * cvecs + cv);
      const float4 b = __ldg(base + (long long)t.o2 * cvecs + cv);
      const float4 c = __ldg(base + (long long)t.o3 * cvecs + cv);
      const float4 d = __ldg(base + (long long)t.o4 * cvecs + cv);
      Taps w;
      w.w1 = t.w1; w.w2 = t.w2; w.w3 = t.w3; w.w4 = t.w4;
      a0 = __fadd_rn(a0, blend(w, a.x, b.x, c.x, d.x));
      a1 = __fadd_rn(a1, blend(w, a.y, b.y, c.y, d.y));
      a2 = __fadd_rn(a2, blend(w, a.z, b.z, c.z, d.z));
      a3 = __fadd_rn(a3, blend(w, a.w, b.w, c.w, d.w));
    }
    float4 r;
    r.x = __fdiv_rn(a0, g.count); r.y = __fdiv_rn(a1, g.count);
    r.z = __fdiv_rn(a2, g.count); r.w = __fdiv_rn(a3, g.count);
    o[i] = r;   // i enumerates [ph][pw][cvec] == NHWC order inside the ROI
  }
}

__global__ void __launch_bounds__(256) roi_align_fpn_bwd2_kernel(
    LevelTab tab, int channels, const float* __restrict__ gout, const float* __restrict__ rois,
    const int* __restrict__ levels, int pooled_h, int pooled_w, int sampling_ratio) {
  __shared__ TapEntry taps[kMaxSamplesV2];
  const int n = blockIdx.x;
  const int l = levels[n];
  const int height = tab.h[l], width = tab.w[l];
  const RoiGeom g = roi_geometry(rois + (long long)n * 5, tab.scale[l], pooled_h, pooled_w, sampling_ratio);
  build_tap_table(taps, g, pooled_h, pooled_w, height, width);
  __syncthreads();
  const int cvecs = channels >> 2;
  const int per_bin = g.grid_h * g.grid_w;
  float* base = tab.grad[l] + (long long)g.batch_ind * height * width * channels;
  const float4* go4 = reinterpret_cast<const float4*>(gout) + (long long)n * pooled_h * pooled_w * cvecs;
  const int total = pooled_h * pooled_w * cvecs;
  for (int i = threadIdx.x; i < total; i += blockDim.x) {
    const int cv = i % cvecs, bin = i / cvecs;
    const TapEntry* e = taps + bin * per_bin;
    const float4 go = __ldg(go4 + i);
    for (int k = 0; k < per_bin; ++k) {
      const TapEntry t = e[k];
      if (t.o1 < 0) continue;
      const float ws[4] = {t.w1, t.w2, t.w3, t.w4};
      const int os[4] = {t.o1, t.o2, t.o3, t.o4};
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        float4 v;   // g_k = top_diff * w_k / count  (ROIAlign_cuda.cu:238-241)
        v.x = __fdiv_rn(__fmul_rn(go.x, ws[q]), g.count);
        v.y = __fdiv_rn(__fmul_rn(go.y, ws[q]), g.count);
        v.z = __fdiv_rn(__fmul_rn(go.z, ws[q]), g.count);
        v.w = __fdiv_rn(__fmul_rn(go.w, ws[q]), g.count);
        mmt::red_add_v4(base + ((long long)os[q] * cvecs + cv) * 4, v);
      }
    }
  }
}

inline int grid_for(long long total, int block) {
  long long blocks = (total + block - 1) / block;
  const long long cap = (long long)mmt::kNumSMs * 8 * 4;  // 8 CTAs of 256 thr/SM resident, 4 waves
  if (blocks > cap) blocks = cap - (cap % mmt::kNumSMs);
  return (int)(blocks < 1 ? 1 : blocks);
}

// ---------------------------------------------------------------- ROIPool (ROIPool_cuda.cu)
__global__ void roi_pool_fwd_kernel(const float* __restrict__ in, float scale, int channels,
                                    int height, int width, int pooled_h, int pooled_w,
                                    const float* __restrict__ rois, float* __restrict__ out,
                                    int* __restrict__ argmax, long long total) {
  for (long long index = (long long)blockIdx.x * blockDim.x + threadIdx.x; index < total;
       index += (long long)gridDim.x * blockDim.x) {
    const int pw = index % pooled_w, ph = (index / pooled_w) % pooled_h;
    const int c = (index / pooled_w / pooled_h) % channels;
    const int n = index / pooled_w / pooled_h / channels;
    const float* r = rois + (long long)n * 5;
    const int b = (int)r[0];
    const int sw = (int)roundf(r[1] * scale), sh = (int)roundf(r[2] * scale);
    const int ew = (int)roundf(r[3] * scale), eh = (int)roundf(r[4] * scale);
    const int rw = max(ew - sw + 1, 1), rh = max(eh - sh + 1, 1);
    const float bh = (float)rh / (float)pooled_h, bw = (float)rw / (float)pooled_w;
    int hs = (int)floorf((float)ph * bh), ws = (int)floorf((float)pw * bw);
    int he = (int)ceilf((float)(ph + 1) * bh), we = (int)ceilf((float)(pw + 1) * bw);
    hs = min(max(hs + sh, 0), height); he = min(max(he + sh, 0), height);
    ws = min(max(ws + sw, 0), width); we = min(max(we + sw, 0), width);
    const bool empty = (he <= hs) || (we <= ws);
    float maxval = empty ? 0.f : -3.402823466e+38f;
    int maxidx = -1;
    const float* src = in + ((long long)b * channels + c) * height * width;
    for (int h = hs; h < he; ++h)
      for (int w = ws; w < we; ++w) {
        const int bi = h * width + w;
        if (src[bi] > maxval) { maxval = src[bi]; maxidx = bi; }
      }
    out[index] = maxval;
    argmax[index] = maxidx;
  }
}

__global__ void roi_pool_bwd_kernel(const float* __restrict__ gout, const int* __restrict__ argmax,
                                    const float* __restrict__ rois, int channels, int height,
                                    int width, int pooled_h, int pooled_w, float* __restrict__ gin,
                                    long long total) {
  for (long long index = (long long)blockIdx.x * blockDim.x + threadIdx.x; index < total;
       index += (long long)gridDim.x * blockDim.x) {
    const int c = (index / pooled_w / pooled_h) % channels;
    const int n = index / pooled_w / pooled_h / channels;
    const int b = (int)rois[(long long)n * 5];
    const int am = argmax[index];
    if (am != -1) atomicAdd(gin + ((long long)b * channels + c) * height * width + am, gout[index]);
  }
}

}  // namespace

extern "C" int mmt_roi_align_forward_f32(const float* input, mmt_strides4 is, int batch, int channels,
                                         int height, int width, const float* rois, int n_rois,
                                         float spatial_scale, int pooled_h, int pooled_w,
                                         int sampling_ratio, float* output, mmt_strides4 os,
                                         void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  MMT_REQUIRE(n_rois >= 0 && channels > 0 && pooled_h > 0 && pooled_w > 0 && batch >= 0,
              "roi_align_forward: bad sizes");
  if (n_rois == 0) return MMT_OK;
  MMT_REQUIRE(input && rois && output, "roi_align_forward: null pointer");
  const bool ch_fast = (is.c == 1);
  const bool vec4 = ch_fast && channels % 4 == 0 && ((uintptr_t)input % 16 == 0) && is.n % 4 == 0 &&
                    is.h % 4 == 0 && is.w % 4 == 0 &&
                    (os.c != 1 || (((uintptr_t)output % 16 == 0) && os.n % 4 == 0 && os.h % 4 == 0 && os.w % 4 == 0));
  const long long total = (long long)n_rois * pooled_h * pooled_w * (channels / (vec4 ? 4 : 1));
  const int grid = grid_for(total, 256);
  if (vec4)
    roi_align_fwd_kernel<4, true><<<grid, 256, 0, stream>>>(input, is, channels, height, width, rois,
                                                            spatial_scale, pooled_h, pooled_w,
                                                            sampling_ratio, output, os, total);
  else if (ch_fast)
    roi_align_fwd_kernel<1, true><<<grid, 256, 0, stream>>>(input, is, channels, height, width, rois,
                                                            spatial_scale, pooled_h, pooled_w,
                                                            sampling_ratio, output, os, total);
  else
    roi_align_fwd_kernel<1, false><<<grid, 256, 0, stream>>>(input, is, channels, height, width, rois,
                                                             spatial_scale, pooled_h, pooled_w,
                                                             sampling_ratio, output, os, total);
  return mmt::check_launch("roi_align_forward");
}

extern "C" int mmt_roi_align_backward_f32(const float* grad_output, mmt_strides4 gs, const float* rois,
                                          int n_rois, float spatial_scale, int pooled_h, int pooled_w,
                                          int sampling_ratio, float* grad_input, mmt_strides4 is,
                                          int batch, int channels, int height, int width,
                                          void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  MMT_REQUIRE(n_rois >= 0 && channels > 0 && pooled_h > 0 && pooled_w > 0 && batch >= 0,
              "roi_align_backward: bad sizes");
  if (n_rois == 0) return MMT_OK;
  MMT_REQUIRE(grad_output && rois && grad_input, "roi_align_backward: null pointer");
  const bool ch_fast = (is.c == 1);
  const bool vec4 = ch_fast && channels % 4 == 0 && ((uintptr_t)grad_input % 16 == 0) &&
                    is.n % 4 == 0 && is.h % 4 == 0 && is.w % 4 == 0 &&
                    (gs.c != 1 || (((uintptr_t)grad_output % 16 == 0) && gs.n % 4 == 0 && gs.h % 4 == 0 && gs.w % 4 == 0));
  const long long total = (long long)n_rois * pooled_h * pooled_w * (channels / (vec4 ? 4 : 1));
  const int grid = grid_for(total, 256);
  if (vec4)
    roi_align_bwd_kernel<4, true><<<grid, 256, 0, stream>>>(grad_output, gs, rois, spatial_scale,
                                                            pooled_h, pooled_w, sampling_ratio,
                                                            grad_input, is, channels, height, width, total);
  else if (ch_fast)
    roi_align_bwd_kernel<1, true><<<grid, 256, 0, stream>>>(grad_output, gs, rois, spatial_scale,
                                                            pooled_h, pooled_w, sampling_ratio,
                                                            grad_input, is, channels, height, width, total);
  else
    roi_align_bwd_kernel<1, false><<<grid, 256, 0, stream>>>(grad_output, gs, rois, spatial_scale,
                                                             pooled_h, pooled_w, sampling_ratio,
                                                             grad_input, is, channels, height, width, total);
  return mmt::check_launch("roi_align_backward");
}

extern "C" int mmt_roi_pool_forward_f32(const float* input, int batch, int channels, int height,
                                        int width, const float* rois, int n_rois, float spatial_scale,
                                        int pooled_h, int pooled_w, float* output, int* argmax,
                                        void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  MMT_REQUIRE(n_rois >= 0 && channels > 0 && batch >= 0, "roi_pool_forward: bad sizes");
  if (n_rois == 0) return MMT_OK;
  MMT_REQUIRE(input && rois && output && argmax, "roi_pool_forward: null pointer");
  const long long total = (long long)n_rois * channels * pooled_h * pooled_w;
  roi_pool_fwd_kernel<<<grid_for(total, 256), 256, 0, stream>>>(input, spatial_scale, channels, height,
                                                                width, pooled_h, pooled_w, rois,
                                                                output, argmax, total);
  return mmt::check_launch("roi_pool_forward");
}

extern "C" int mmt_roi_pool_backward_f32(const float* grad_output, const int* argmax, const float* rois,
                                         int n_rois, int pooled_h, int pooled_w, float* grad_input,
                                         int batch, int channels, int height, int width, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  MMT_REQUIRE(n_rois >= 0 && channels > 0 && batch >= 0, "roi_pool_backward: bad sizes");
  if (n_rois == 0) return MMT_OK;
  MMT_REQUIRE(grad_output && argmax && rois && grad_input, "roi_pool_backward: null pointer");
  const long long total = (long long)n_rois * channels * pooled_h * pooled_w;
  roi_pool_bwd_kernel<<<grid_for(total, 256), 256, 0, stream>>>(grad_output, argmax, rois, channels,
                                                                height, width, pooled_h, pooled_w,
                                                                grad_input, total);
  return mmt::check_launch("roi_pool_backward");
}

static int fill_level_tab(LevelTab& tab, const float* const* feats, float* const* grads, const int* hs,
                          const int* ws, const float* scales, int n_levels, const char* who) {
  MMT_REQUIRE(n_levels >= 1 && n_levels <= kMaxLevels, "%s: n_levels must be in [1,%d]", who, kMaxLevels);
  MMT_REQUIRE(hs && ws && scales, "%s: null level table", who);
  for (int l = 0; l < kMaxLevels; ++l) {
    tab.feat[l] = nullptr; tab.grad[l] = nullptr; tab.h[l] = tab.w[l] = 1; tab.scale[l] = 1.f;
  }
  for (int l = 0; l < n_levels; ++l) {
    const void* p = feats ? (const void*)feats[l] : (const void*)grads[l];
    MMT_REQUIRE(p != nullptr && ((uintptr_t)p % 16) == 0, "%s: level %d pointer null or not 16-byte aligned", who, l);
    MMT_REQUIRE(hs[l] > 0 && ws[l] > 0, "%s: bad level size", who);
    if (feats) tab.feat[l] = feats[l];
    if (grads) tab.grad[l] = grads[l];
    tab.h[l] = hs[l]; tab.w[l] = ws[l]; tab.scale[l] = scales[l];
  }
  return MMT_OK;
}

extern "C" int mmt_roi_align_fpn_forward_f32(const float* const* level_feats_host, const int* level_h_host,
                                             const int* level_w_host, const float* level_scale_host,
                                             int n_levels, int batch, int channels, const float* rois,
                                             int n_rois, int pooled_h, int pooled_w, int sampling_ratio,
                                             const int* roi_levels, float* output, void* stream_) {
  MMT_REQUIRE(n_rois >= 0 && channels > 0 && channels % 4 == 0 && pooled_h > 0 && pooled_w > 0 && batch >= 0,
              "roi_align_fpn_forward: bad sizes (channels must be a multiple of 4)");
  if (n_rois == 0) return MMT_OK;
  MMT_REQUIRE(level_feats_host && rois && roi_levels && output, "roi_align_fpn_forward: null pointer");
  MMT_REQUIRE((uintptr_t)output % 16 == 0, "roi_align_fpn_forward: output must be 16-byte aligned");
  LevelTab tab;
  if (int rc = fill_level_tab(tab, level_feats_host, nullptr, level_h_host, level_w_host, level_scale_host,
                              n_levels, "roi_align_fpn_forward")) return rc;
  const long long total = (long long)n_rois * pooled_h * pooled_w * (channels / 4);
  // algorithmic bytes: the pooled output written once (the taps are gathers from L2-resident maps)
  mmt::ProfScope prof(MMT_PROF_ROI_ALIGN, 0.0, 16.0 * total, (cudaStream_t)stream_);
  if (sampling_ratio > 0 && pooled_h * pooled_w * sampling_ratio * sampling_ratio <= kMaxSamplesV2)
    roi_align_fpn_fwd2_kernel<<<n_rois, 256, 0, (cudaStream_t)stream_>>>(tab, channels, rois, roi_levels, pooled_h,
                                                                        pooled_w, sampling_ratio, output);
  else   // adaptive sampling grid (sampling_ratio <= 0): per-thread geometry
    roi_align_fpn_fwd_kernel<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream_>>>(
        tab, channels, rois, roi_levels, pooled_h, pooled_w, sampling_ratio, output, total);
  return mmt::check_launch("roi_align_fpn_forward");
}

extern "C" int mmt_roi_align_fpn_backward_f32(const float* grad_output, const float* rois,
                                              const int* roi_levels, int n_rois, int pooled_h, int pooled_w,
                                              int sampling_ratio, float* const* level_grads_host,
                                              const int* level_h_host, const int* level_w_host,
                                              const float* level_scale_host, int n_levels, int batch,
                                              int channels, void* stream_) {
  MMT_REQUIRE(n_rois >= 0 && channels > 0 && channels % 4 == 0 && pooled_h > 0 && pooled_w > 0 && batch >= 0,
              "roi_align_fpn_backward: bad sizes (channels must be a multiple of 4)");
  if (n_rois == 0) return MMT_OK;
  MMT_REQUIRE(grad_output && rois && roi_levels && level_grads_host, "roi_align_fpn_backward: null pointer");
  MMT_REQUIRE((uintptr_t)grad_output % 16 == 0, "roi_align_fpn_backward: grad_output must be 16-byte aligned");
  LevelTab tab;
  if (int rc = fill_level_tab(tab, nullptr, level_grads_host, level_h_host, level_w_host, level_scale_host,
                              n_levels, "roi_align_fpn_backward")) return rc;
  const long long total = (long long)n_rois * pooled_h * pooled_w * (channels / 4);
  mmt::ProfScope prof(MMT_PROF_ROI_ALIGN, 0.0, 16.0 * total, (cudaStream_t)stream_);
  if (sampling_ratio > 0 && pooled_h * pooled_w * sampling_ratio * sampling_ratio <= kMaxSamplesV2)
    roi_align_fpn_bwd2_kernel<<<n_rois, 256, 0, (cudaStream_t)stream_>>>(tab, channels, grad_output, rois, roi_levels,
                                                                        pooled_h, pooled_w, sampling_ratio);
  else
    roi_align_fpn_bwd_kernel<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream_>>>(
        tab, channels, grad_output, rois, roi_levels, pooled_h, pooled_w, sampling_ratio, total);
  return mmt::check_launch("roi_align_fpn_backward");
}
