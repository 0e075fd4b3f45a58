// Plain fp32 (FFMA) implicit-GEMM convolution kernels, NHWC / OHWI, any shape.
//
// Role: (1) the shapes the tensor-core engine does not take — the 7x7/s2 stem with in_c = 3
// (maskrcnn_benchmark/modeling/backbone/resnet.py:288-293) and the skinny RPN / mask-logit heads
// (out_c = 3, 12; modeling/rpn/rpn.py:27-33); (2) the exact-fp32 on-device cross-check of
// conv_tc.cu at full problem sizes.  64x64 output tile, BK = 16, 256 threads, 4x4 micro-tile.
#include "common.cuh"
#include "conv_common.cuh"
#include <algorithm>

namespace {

constexpr int BM = 64, BN = 64, BK = 16;

struct Epilogue {
  const float* scale;
  const float* bias;
  const float* residual;
  int relu;
};

// --------------------------------------------------------------------------- element gathers
// fprop: A[m][k] = x[b, oy*s - p + r, ox*s - p + q, c],  k = (r*KW + q)*Cin + c
struct FpropA {
  const float* x;
  mmt_conv_shape s;
  __device__ __forceinline__ float operator()(long long m, int k) const {
    const int c = k % s.in_c;
    const int t = k / s.in_c;
    const int q = t % s.k_w, r = t / s.k_w;
    const int ox = (int)(m % s.out_w);
    const long long u = m / s.out_w;
    const int oy = (int)(u % s.out_h);
    const int b = (int)(u / s.out_h);
    const int iy = oy * s.stride - s.pad + r, ix = ox * s.stride - s.pad + q;
    if (iy < 0 || iy >= s.in_h || ix < 0 || ix >= s.in_w) return 0.f;
    return __ldg(x + (((long long)b * s.in_h + iy) * s.in_w + ix) * s.in_c + c);
  }
};
// fprop: B[k][n] = w[n][k]
struct FpropB {
  const float* w;
  int K;
  __device__ __forceinline__ float operator()(int k, int n) const { return __ldg(w + (long long)n * K + k); }
};
// dgrad: A[m][k] = dy[b, (iy+p-r)/s, (ix+p-q)/s, co], k = (r*KW + q)*Cout + co ; m = input pixel
struct DgradA {
  const float* dy;
  mmt_conv_shape s;
  __device__ __forceinline__ float operator()(long long m, int k) const {
    const int co = k % s.out_c;
    const int t = k / s.out_c;
    const int q = t % s.k_w, r = t / s.k_w;
    const int ix = (int)(m % s.in_w);
    const long long u = m / s.in_w;
    const int iy = (int)(u % s.in_h);
    const int b = (int)(u / s.in_h);
    const int ny = iy + s.pad - r, nx = ix + s.pad - q;
    if (ny < 0 || nx < 0 || ny % s.stride || nx % s.stride) return 0.f;
    const int oy = ny / s.stride, ox = nx / s.stride;
    if (oy >= s.out_h || ox >= s.out_w) return 0.f;
    return __ldg(dy + (((long long)b * s.out_h + oy) * s.out_w + ox) * s.out_c + co);
  }
};
// dgrad: B[k][n] = w[co][r][q][n]
struct DgradB {
  const float* w;
  mmt_conv_shape s;
  __device__ __forceinline__ float operator()(int k, int n) const {
    const int co = k % s.out_c;
    const int t = k / s.out_c;  // r*KW + q
    return __ldg(w + ((long long)co * s.k_h * s.k_w + t) * s.in_c + n);
  }
};
// wgrad: A[m=co][k=p] = dy[p][co] ; B[k=p][n=(r,q,ci)] = x[b, oy*s-p+r, ox*s-p+q, ci]
struct WgradA {
  const float* dy;
  int out_c;
  __device__ __forceinline__ float operator()(long long m, long long k) const { return __ldg(dy + k * out_c + m); }
};
struct WgradB {
  const float* x;
  mmt_conv_shape s;
  __device__ __forceinline__ float operator()(long long p, int n) const {
    const int c = n % s.in_c;
    const int t = n / s.in_c;
    const int q = t % s.k_w, r = t / s.k_w;
    const int ox = (int)(p % s.out_w);
    const long long u = p / s.out_w;
    const int oy = (int)(u % s.out_h);
    const int b = (int)(u / s.out_h);
    const int iy = oy * s.stride - s.pad + r, ix = ox * s.stride - s.pad + q;
    if (iy < 0 || iy >= s.in_h || ix < 0 || ix >= s.in_w) return 0.f;
    return __ldg(x + (((long long)b * s.in_h + iy) * s.in_w + ix) * s.in_c + c);
  }
};

// Generic tile GEMM: C[M][N] (+)= sum_{k in [k_begin,k_end)} A(m,k) * B(k,n)
template <class LA, class LB, bool ATOMIC>
__global__ void __launch_bounds__(256) tile_gemm_kernel(LA la, LB lb, long long M, int N, long long K,
                                                        long long k_chunk, float* __restrict__ C,
                                                        Epilogue ep) {
  __shared__ float As[BK][BM + 4];
  __shared__ float Bs[BK][BN + 4];
  const long long m0 = (long long)blockIdx.x * BM;
  const int n0 = blockIdx.y * BN;
  const long long k_begin = (long long)blockIdx.z * k_chunk;
  const long long k_end = min(K, k_begin + k_chunk);
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  for (long long k0 = k_begin; k0 < k_end; k0 += BK) {
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int idx = tid + e * 256;  // 1024 elements per operand tile
      {
        const int kk = idx & 15, mm = idx >> 4;  // k fastest: contiguous channels in NHWC
        const long long m = m0 + mm, k = k0 + kk;
        As[kk][mm] = (m < M && k < k_end) ? la(m, k) : 0.f;
      }
      {
        const int kk = idx & 15, nn = idx >> 4;
        const int n = n0 + nn;
        const long long k = k0 + kk;
        Bs[kk][nn] = (n < N && k < k_end) ? lb(k, n) : 0.f;
      }
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      float a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = As[kk][ty * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = Bs[kk][tx * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const long long m = m0 + ty * 4 + i;
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n >= N) continue;
      float v = acc[i][j];
      if (ATOMIC) {
        atomicAdd(C + m * N + n, v);
      } else {
        if (ep.scale) v *= __ldg(ep.scale + n);
        if (ep.bias) v += __ldg(ep.bias + n);
        if (ep.residual) v += __ldg(ep.residual + m * N + n);
        if (ep.relu) v = fmaxf(v, 0.f);
        C[m * N + n] = v;
      }
    }
  }
}

__global__ void __launch_bounds__(256) bias_grad_kernel(const float* __restrict__ dy, long long pixels,
                                                        int out_c, float* __restrict__ dbias) {
  // grid.x = out_c / 32 (ceil); block = 32 x 8: lanes over channels, 8 pixel slices
  __shared__ float part[8][33];
  const int c = blockIdx.x * 32 + (threadIdx.x & 31);
  const int slice = threadIdx.x >> 5;
  float acc = 0.f;
  if (c < out_c)
    for (long long p = (long long)blockIdx.y * 8 + slice; p < pixels; p += (long long)gridDim.y * 8)
      acc += __ldg(dy + p * out_c + c);
  part[slice][threadIdx.x & 31] = acc;
  __syncthreads();
  if (slice == 0 && c < out_c) {
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += part[i][threadIdx.x & 31];
    atomicAdd(dbias + c, s);
  }
}

}  // namespace

namespace mmt {

int conv_fprop_simt(const float* x, const float* w, const float* scale, const float* bias,
                    const float* residual, float* y, const mmt_conv_shape& s, int flags,
                    cudaStream_t stream) {
  const long long M = (long long)s.batch * s.out_h * s.out_w;
  const int K = s.k_h * s.k_w * s.in_c;
  if (M == 0) return MMT_OK;
  FpropA la{x, s};
  FpropB lb{w, K};
  Epilogue ep{scale, bias, residual, (flags & MMT_EPI_RELU) ? 1 : 0};
  dim3 grid((unsigned)ceil_div<long long>(M, BM), (unsigned)ceil_div(s.out_c, BN), 1);
  ProfScope prof(MMT_PROF_CONV_SIMT, 2.0 * M * s.out_c * K, 4.0 * (M * (double)s.out_c + (double)s.batch * s.in_h * s.in_w * s.in_c), stream);
  tile_gemm_kernel<FpropA, FpropB, false><<<grid, 256, 0, stream>>>(la, lb, M, s.out_c, K, K, y, ep);
  return check_launch("conv_fprop_simt");
}

int conv_dgrad_simt(const float* dy, const float* w, float* dx, const mmt_conv_shape& s,
                    cudaStream_t stream) {
  const long long M = (long long)s.batch * s.in_h * s.in_w;
  const int K = s.k_h * s.k_w * s.out_c;
  if (M == 0) return MMT_OK;
  DgradA la{dy, s};
  DgradB lb{w, s};
  Epilogue ep{nullptr, nullptr, nullptr, 0};
  dim3 grid((unsigned)ceil_div<long long>(M, BM), (unsigned)ceil_div(s.in_c, BN), 1);
  ProfScope prof(MMT_PROF_CONV_SIMT, 2.0 * M * s.in_c * K, 4.0 * (M * (double)s.in_c + (double)s.batch * s.out_h * s.out_w * s.out_c), stream);
  tile_gemm_kernel<DgradA, DgradB, false><<<grid, 256, 0, stream>>>(la, lb, M, s.in_c, K, K, dx, ep);
  return check_launch("conv_dgrad_simt");
}

int conv_wgrad_simt(const float* x, const float* dy, float* dw, float* dbias, const mmt_conv_shape& s,
                    cudaStream_t stream) {
  const long long P = (long long)s.batch * s.out_h * s.out_w;
  const int N = s.k_h * s.k_w * s.in_c;
  cudaError_t e = cudaMemsetAsync(dw, 0, sizeof(float) * (size_t)s.out_c * N, stream);
  if (e != cudaSuccess) return fail(MMT_ERR_CUDA, "conv_wgrad memset: %s", cudaGetErrorString(e));
  if (P == 0) {
    if (dbias) cudaMemsetAsync(dbias, 0, sizeof(float) * (size_t)s.out_c, stream);
    return MMT_OK;
  }
  WgradA la{dy, s.out_c};
  WgradB lb{x, s};
  Epilogue ep{nullptr, nullptr, nullptr, 0};
  const int tiles = ceil_div(s.out_c, BM) * ceil_div(N, BN);
  int split = (int)std::min<long long>(ceil_div<long long>(P, 256), std::max(1, (kNumSMs * 4) / tiles));
  split = min(split, 65535);
  long long k_chunk = ceil_div<long long>(ceil_div<long long>(P, split), BK) * BK;
  split = (int)ceil_div<long long>(P, k_chunk);
  dim3 grid((unsigned)ceil_div(s.out_c, BM), (unsigned)ceil_div(N, BN), (unsigned)split);
  ProfScope prof(MMT_PROF_CONV_SIMT, 2.0 * P * s.out_c * N, 4.0 * (P * (double)s.out_c + (double)s.batch * s.in_h * s.in_w * s.in_c), stream);
  tile_gemm_kernel<WgradA, WgradB, true><<<grid, 256, 0, stream>>>(la, lb, s.out_c, N, P, k_chunk, dw, ep);
  if (int rc = check_launch("conv_wgrad_simt")) return rc;
  if (dbias) return conv_bias_grad(dy, P, s.out_c, dbias, stream);
  return MMT_OK;
}

int conv_bias_grad(const float* dy, long long pixels, int out_c, float* dbias, cudaStream_t stream) {
  cudaError_t e = cudaMemsetAsync(dbias, 0, sizeof(float) * (size_t)out_c, stream);
  if (e != cudaSuccess) return fail(MMT_ERR_CUDA, "conv_bias_grad memset: %s", cudaGetErrorString(e));
  if (pixels == 0) return MMT_OK;
  dim3 g2((unsigned)ceil_div(out_c, 32), (unsigned)std::min<long long>(256, ceil_div<long long>(pixels, 64)));
  bias_grad_kernel<<<g2, 256, 0, stream>>>(dy, pixels, out_c, dbias);
  return check_launch("conv_bias_grad");
}

}  // namespace mmt
