// Elementwise / data-movement kernels around the conv engine (NHWC fp32, HBM bound).
//
// Reference chains these replace (each a separate full-tensor pass there):
//   FrozenBN + ReLU backward   maskrcnn_benchmark/layers/batch_norm.py:19-24, resnet.py:254-275
//   stem max-pool 3x3/s2/p1    maskrcnn_benchmark/modeling/backbone/resnet.py:292
//   P6 = stride-2 subsample    maskrcnn_benchmark/modeling/backbone/fpn.py:72-74
//   FPN top-down merge         maskrcnn_benchmark/modeling/backbone/fpn.py:57-62
#include "common.cuh"

namespace {

inline int stream_grid(long long items, int block = 256) {
  long long blocks = (items + block - 1) / block;
  const long long cap = (long long)mmt::kNumSMs * 8;
  if (blocks > cap) blocks = cap;
  return (int)(blocks < 1 ? 1 : blocks);
}

__global__ void __launch_bounds__(256) bn_relu_bwd_kernel(const float4* __restrict__ gy,
                                                          const float4* __restrict__ y,
                                                          const float* __restrict__ scale,
                                                          float4* __restrict__ gz, long long n4,
                                                          int c4, int relu) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4;
       i += (long long)gridDim.x * blockDim.x) {
    float4 g = __ldg(gy + i);
    if (relu) {
      const float4 v = __ldg(y + i);
      g.x = v.x > 0.f ? g.x : 0.f; g.y = v.y > 0.f ? g.y : 0.f;
      g.z = v.z > 0.f ? g.z : 0.f; g.w = v.w > 0.f ? g.w : 0.f;
    }
    if (scale) {
      const float4 s = __ldg(reinterpret_cast<const float4*>(scale) + (i % c4));
      g.x *= s.x; g.y *= s.y; g.z *= s.z; g.w *= s.w;
    }
    gz[i] = g;
  }
}

__global__ void __launch_bounds__(256) maxpool3x3s2_kernel(const float4* __restrict__ x,
                                                           float4* __restrict__ y, int batch,
                                                           int H, int W, int OH, int OW, int c4) {
  const long long total = (long long)batch * OH * OW * c4;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % c4);
    long long r = i / c4;
    const int ox = (int)(r % OW); r /= OW;
    const int oy = (int)(r % OH);
    const int b = (int)(r / OH);
    float4 m = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
#pragma unroll
    for (int dy = 0; dy < 3; ++dy) {
      const int iy = oy * 2 - 1 + dy;
      if (iy < 0 || iy >= H) continue;
#pragma unroll
      for (int dx = 0; dx < 3; ++dx) {
        const int ix = ox * 2 - 1 + dx;
        if (ix < 0 || ix >= W) continue;
        const float4 v = __ldg(x + (((long long)b * H + iy) * W + ix) * c4 + c);
        m.x = fmaxf(m.x, v.x); m.y = fmaxf(m.y, v.y); m.z = fmaxf(m.z, v.z); m.w = fmaxf(m.w, v.w);
      }
    }
    y[i] = m;
  }
}

// dir 0: y[oy][ox] = x[2oy][2ox]; dir 1: gx[2oy][2ox] += gy[oy][ox]
__global__ void __launch_bounds__(256) subsample2_kernel(const float4* __restrict__ src,
                                                         float4* __restrict__ dst, int batch, int H,
                                                         int W, int OH, int OW, int c4, int dir) {
  const long long total = (long long)batch * OH * OW * c4;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % c4);
    long long r = i / c4;
    const int ox = (int)(r % OW); r /= OW;
    const int oy = (int)(r % OH);
    const int b = (int)(r / OH);
    const long long big = (((long long)b * H + 2 * oy) * W + 2 * ox) * c4 + c;
    if (dir == 0) dst[i] = __ldg(src + big);
    else {
      float4 a = dst[big];
      const float4 g = __ldg(src + i);
      a.x += g.x; a.y += g.y; a.z += g.z; a.w += g.w;
      dst[big] = a;
    }
  }
}

__global__ void __launch_bounds__(256) upsample2_add_kernel(const float4* __restrict__ lat,
                                                            const float4* __restrict__ top,
                                                            float4* __restrict__ out, int batch,
                                                            int H, int W, int TH, int TW, int c4) {
  const long long total = (long long)batch * H * W * c4;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % c4);
    long long r = i / c4;
    const int x = (int)(r % W); r /= W;
    const int y = (int)(r % H);
    const int b = (int)(r / H);
    // F.interpolate(scale_factor=2, mode='nearest'): src = floor(dst / 2)
    const int ty = min(y >> 1, TH - 1), tx = min(x >> 1, TW - 1);
    const float4 a = __ldg(lat + i);
    const float4 t = __ldg(top + (((long long)b * TH + ty) * TW + tx) * c4 + c);
    out[i] = make_float4(a.x + t.x, a.y + t.y, a.z + t.z, a.w + t.w);
  }
}

// grad_top[ty][tx] += sum over the (up to) 2x2 children of grad_out
__global__ void __launch_bounds__(256) upsample2_bwd_kernel(const float4* __restrict__ gout,
                                                            float4* __restrict__ gtop, int batch,
                                                            int H, int W, int TH, int TW, int c4) {
  const long long total = (long long)batch * TH * TW * c4;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % c4);
    long long r = i / c4;
    const int tx = (int)(r % TW); r /= TW;
    const int ty = (int)(r % TH);
    const int b = (int)(r / TH);
    float4 acc = gtop[i];
#pragma unroll
    for (int dy = 0; dy < 2; ++dy) {
      const int y = 2 * ty + dy;
      if (y >= H) continue;
#pragma unroll
      for (int dx = 0; dx < 2; ++dx) {
        const int x = 2 * tx + dx;
        if (x >= W) continue;
        const float4 g = __ldg(gout + (((long long)b * H + y) * W + x) * c4 + c);
        acc.x += g.x; acc.y += g.y; acc.z += g.z; acc.w += g.w;
      }
    }
    gtop[i] = acc;
  }
}

// [n][rows][cols] -> [n][cols][rows] through a padded smem tile (both sides coalesced)
__global__ void __launch_bounds__(256) transpose_kernel(const float* __restrict__ src,
                                                        float* __restrict__ dst, int rows, int cols) {
  __shared__ float tile[32][33];
  const long long plane = (long long)rows * cols;
  const float* s = src + blockIdx.z * plane;
  float* d = dst + blockIdx.z * plane;
  const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
  for (int k = ty; k < 32; k += 8) {
    const int r = r0 + k, c = c0 + tx;
    if (r < rows && c < cols) tile[k][tx] = s[(long long)r * cols + c];
  }
  __syncthreads();
  for (int k = ty; k < 32; k += 8) {
    const int c = c0 + k, r = r0 + tx;
    if (r < rows && c < cols) d[(long long)c * rows + r] = tile[tx][k];
  }
}

}  // namespace

extern "C" int mmt_bn_relu_backward_f32(const float* grad_y, const float* y, const float* scale,
                                        float* grad_z, int64_t pixels, int channels,
                                        int apply_relu_mask, void* stream_) {
  MMT_REQUIRE(pixels >= 0 && channels > 0 && channels % 4 == 0, "bn_relu_backward: channels must be a multiple of 4");
  if (pixels == 0) return MMT_OK;
  MMT_REQUIRE(grad_y && grad_z && (!apply_relu_mask || y), "bn_relu_backward: null pointer");
  const long long n4 = pixels * channels / 4;
  bn_relu_bwd_kernel<<<stream_grid(n4), 256, 0, (cudaStream_t)stream_>>>(
      (const float4*)grad_y, (const float4*)y, scale, (float4*)grad_z, n4, channels / 4, apply_relu_mask);
  return mmt::check_launch("bn_relu_backward");
}

extern "C" int mmt_maxpool3x3s2_forward_f32(const float* x, float* y, int batch, int in_h, int in_w,
                                            int channels, void* stream_) {
  MMT_REQUIRE(batch > 0 && in_h > 0 && in_w > 0 && channels % 4 == 0, "maxpool: bad sizes");
  MMT_REQUIRE(x && y, "maxpool: null pointer");
  const int OH = (in_h + 2 - 3) / 2 + 1, OW = (in_w + 2 - 3) / 2 + 1;
  const long long total = (long long)batch * OH * OW * (channels / 4);
  maxpool3x3s2_kernel<<<stream_grid(total), 256, 0, (cudaStream_t)stream_>>>(
      (const float4*)x, (float4*)y, batch, in_h, in_w, OH, OW, channels / 4);
  return mmt::check_launch("maxpool3x3s2");
}

extern "C" int mmt_subsample2_forward_f32(const float* x, float* y, int batch, int in_h, int in_w,
                                          int channels, void* stream_) {
  MMT_REQUIRE(batch > 0 && in_h > 0 && in_w > 0 && channels % 4 == 0, "subsample2: bad sizes");
  MMT_REQUIRE(x && y, "subsample2: null pointer");
  const int OH = (in_h - 1) / 2 + 1, OW = (in_w - 1) / 2 + 1;
  const long long total = (long long)batch * OH * OW * (channels / 4);
  subsample2_kernel<<<stream_grid(total), 256, 0, (cudaStream_t)stream_>>>(
      (const float4*)x, (float4*)y, batch, in_h, in_w, OH, OW, channels / 4, 0);
  return mmt::check_launch("subsample2");
}

extern "C" int mmt_subsample2_backward_f32(const float* grad_y, float* grad_x, int batch, int in_h,
                                           int in_w, int channels, void* stream_) {
  MMT_REQUIRE(batch > 0 && in_h > 0 && in_w > 0 && channels % 4 == 0, "subsample2_backward: bad sizes");
  MMT_REQUIRE(grad_y && grad_x, "subsample2_backward: null pointer");
  const int OH = (in_h - 1) / 2 + 1, OW = (in_w - 1) / 2 + 1;
  const long long total = (long long)batch * OH * OW * (channels / 4);
  subsample2_kernel<<<stream_grid(total), 256, 0, (cudaStream_t)stream_>>>(
      (const float4*)grad_y, (float4*)grad_x, batch, in_h, in_w, OH, OW, channels / 4, 1);
  return mmt::check_launch("subsample2_backward");
}

extern "C" int mmt_upsample2_add_f32(const float* lateral, const float* top, float* out, int batch,
                                     int h, int w, int top_h, int top_w, int channels, void* stream_) {
  MMT_REQUIRE(batch > 0 && h > 0 && w > 0 && channels % 4 == 0, "upsample2_add: bad sizes");
  MMT_REQUIRE(h == 2 * top_h && w == 2 * top_w, "upsample2_add: lateral must be exactly 2x the top map (%dx%d vs %dx%d)", h, w, top_h, top_w);
  MMT_REQUIRE(lateral && top && out, "upsample2_add: null pointer");
  const long long total = (long long)batch * h * w * (channels / 4);
  upsample2_add_kernel<<<stream_grid(total), 256, 0, (cudaStream_t)stream_>>>(
      (const float4*)lateral, (const float4*)top, (float4*)out, batch, h, w, top_h, top_w, channels / 4);
  return mmt::check_launch("upsample2_add");
}

extern "C" int mmt_upsample2_backward_f32(const float* grad_out, float* grad_top, int batch, int h,
                                          int w, int top_h, int top_w, int channels, void* stream_) {
  MMT_REQUIRE(batch > 0 && h > 0 && w > 0 && channels % 4 == 0, "upsample2_backward: bad sizes");
  MMT_REQUIRE(grad_out && grad_top, "upsample2_backward: null pointer");
  const long long total = (long long)batch * top_h * top_w * (channels / 4);
  upsample2_bwd_kernel<<<stream_grid(total), 256, 0, (cudaStream_t)stream_>>>(
      (const float4*)grad_out, (float4*)grad_top, batch, h, w, top_h, top_w, channels / 4);
  return mmt::check_launch("upsample2_backward");
}

extern "C" int mmt_nchw_to_nhwc_f32(const float* src, float* dst, int n, int c, int h, int w, void* stream_) {
  MMT_REQUIRE(n > 0 && c > 0 && h > 0 && w > 0 && src && dst, "nchw_to_nhwc: bad arguments");
  const int rows = c, cols = h * w;  // [n][c][hw] -> [n][hw][c]
  MMT_REQUIRE(n <= 65535, "nchw_to_nhwc: batch too large");
  transpose_kernel<<<dim3((cols + 31) / 32, (rows + 31) / 32, n), 256, 0, (cudaStream_t)stream_>>>(src, dst, rows, cols);
  return mmt::check_launch("nchw_to_nhwc");
}

extern "C" int mmt_nhwc_to_nchw_f32(const float* src, float* dst, int n, int c, int h, int w, void* stream_) {
  MMT_REQUIRE(n > 0 && c > 0 && h > 0 && w > 0 && src && dst, "nhwc_to_nchw: bad arguments");
  const int rows = h * w, cols = c;  // [n][hw][c] -> [n][c][hw]
  MMT_REQUIRE(n <= 65535, "nhwc_to_nchw: batch too large");
  MMT_REQUIRE((rows + 31) / 32 <= 65535, "nhwc_to_nchw: plane too large");
  transpose_kernel<<<dim3((cols + 31) / 32, (rows + 31) / 32, n), 256, 0, (cudaStream_t)stream_>>>(src, dst, rows, cols);
  return mmt::check_launch("nhwc_to_nchw");
}
