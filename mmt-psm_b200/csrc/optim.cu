// Teacher EMA and fused SGD(+momentum, weight decay)+EMA over flat fp32 parameter arenas.
//
// Replaces `MTtrainer.update_teacher` (maskrcnn_benchmark/engine/MTtrainer.py:277-281: a mul_ and
// an add_ per parameter tensor = 264 launches per iteration) and torch.optim.SGD as configured by
// maskrcnn_benchmark/solver/build.py:5-23 (one param group per tensor).  One launch, each
// element of {grad, momentum, student, teacher} read once and written once: 28 B/element
// algorithmic traffic for SGD+EMA, 12 B/element for EMA alone — pure HBM streaming.
#include "common.cuh"

namespace {

constexpr int kMaxSegSmem = 1024;

__global__ void __launch_bounds__(256) ema_kernel(float4* __restrict__ t, const float4* __restrict__ s,
                                                  long long n4, float alpha, float one_minus,
                                                  float* __restrict__ t_tail,
                                                  const float* __restrict__ s_tail, int tail) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4;
       i += (long long)gridDim.x * blockDim.x) {
    float4 a = t[i];
    const float4 b = __ldg(s + i);
    // reference: mul_(alpha) then add_(1-alpha, param): round the product, then a*x+y
    a.x = __fmul_rn(a.x, alpha) + one_minus * b.x;
    a.y = __fmul_rn(a.y, alpha) + one_minus * b.y;
    a.z = __fmul_rn(a.z, alpha) + one_minus * b.z;
    a.w = __fmul_rn(a.w, alpha) + one_minus * b.w;
    t[i] = a;
  }
  if (blockIdx.x == 0 && (int)threadIdx.x < tail)
    t_tail[threadIdx.x] = __fmul_rn(t_tail[threadIdx.x], alpha) + one_minus * s_tail[threadIdx.x];
}

__global__ void __launch_bounds__(256) sgd_ema_kernel(
    float* __restrict__ p, float* __restrict__ t, const float* __restrict__ g, float* __restrict__ buf,
    long long n, const mmt_param_segment* __restrict__ segs, int n_segs, float momentum, float alpha,
    float one_minus, int first_step, float grad_scale) {
  __shared__ long long s_begin[kMaxSegSmem];
  const int cached = min(n_segs, kMaxSegSmem);
  for (int i = threadIdx.x; i < cached; i += blockDim.x) s_begin[i] = segs[i].begin;
  __syncthreads();
  const long long n4 = (n + 3) / 4;
  for (long long i4 = (long long)blockIdx.x * blockDim.x + threadIdx.x; i4 < n4;
       i4 += (long long)gridDim.x * blockDim.x) {
    const long long e = i4 * 4;
    // The four streams are read FIRST, unconditionally (the arenas are padded to whole float4 groups, so the vector
    // loads are always in bounds): 64 bytes per thread are in flight while the segment lookup below runs, instead of
    // the loads waiting for a 9-step dependent search.  Frozen tensors (< 1 % of the arena) pay two unused loads.
    const float4 P = *reinterpret_cast<const float4*>(p + e);
    const float4 G = __ldg(reinterpret_cast<const float4*>(g + e));
    float4 B = make_float4(0.f, 0.f, 0.f, 0.f), T = B;
    if (!first_step) B = *reinterpret_cast<const float4*>(buf + e);
    if (t) T = *reinterpret_cast<const float4*>(t + e);
    // segment lookup: last segment whose begin <= e
    int lo = 0, hi = n_segs - 1;
    while (lo < hi) {
      const int mid = (lo + hi + 1) >> 1;
      const long long b = mid < cached ? s_begin[mid] : segs[mid].begin;
      if (b <= e) lo = mid; else hi = mid - 1;
    }
    const mmt_param_segment sg = segs[lo];
    if (e < sg.begin || e >= sg.end) continue;  // padding between tensors
    const int cnt = (int)min((long long)4, sg.end - e);
    float pv[4] = {P.x, P.y, P.z, P.w}, tv[4] = {T.x, T.y, T.z, T.w};
    float gv[4] = {G.x, G.y, G.z, G.w}, bv[4] = {B.x, B.y, B.z, B.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      if (k < cnt) {
        if (sg.trainable) {
          // torch.optim.SGD: d = g + wd*p ; buf = mom*buf + d ; p = p - lr*buf
          float d = gv[k] * grad_scale;
          if (sg.weight_decay != 0.f) d = d + sg.weight_decay * pv[k];
          const float nb = first_step ? d : (__fmul_rn(bv[k], momentum) + d);
          bv[k] = nb;
          pv[k] = pv[k] - sg.lr * nb;
        }
        if (t) tv[k] = __fmul_rn(tv[k], alpha) + one_minus * pv[k];
      }
    }
    if (cnt == 4) {
      if (sg.trainable) {
        *reinterpret_cast<float4*>(p + e) = make_float4(pv[0], pv[1], pv[2], pv[3]);
        *reinterpret_cast<float4*>(buf + e) = make_float4(bv[0], bv[1], bv[2], bv[3]);
      }
      if (t) *reinterpret_cast<float4*>(t + e) = make_float4(tv[0], tv[1], tv[2], tv[3]);
    } else {  // last group of a tensor: the elements behind it are padding and stay untouched
      for (int k = 0; k < cnt; ++k) {
        if (sg.trainable) { p[e + k] = pv[k]; buf[e + k] = bv[k]; }
        if (t) t[e + k] = tv[k];
      }
    }
  }
}

inline int stream_grid(long long items) {
  long long blocks = (items + 255) / 256;
  const long long cap = (long long)mmt::kNumSMs * 8;  // 8 x 256 threads resident per SM
  if (blocks > cap) blocks = cap;
  return (int)(blocks < 1 ? 1 : blocks);
}

}  // namespace

extern "C" int mmt_ema_f32(float* teacher, const float* student, int64_t n, float alpha, void* stream_) {
  MMT_REQUIRE(n >= 0, "ema: negative size");
  if (n == 0) return MMT_OK;
  MMT_REQUIRE(teacher && student, "ema: null pointer");
  MMT_REQUIRE(((uintptr_t)teacher % 16 == 0) && ((uintptr_t)student % 16 == 0), "ema: arenas must be 16-byte aligned");
  const long long n4 = n / 4;
  const int tail = (int)(n - n4 * 4);
  ema_kernel<<<stream_grid(n4 ? n4 : 1), 256, 0, (cudaStream_t)stream_>>>(
      (float4*)teacher, (const float4*)student, n4, alpha, 1.f - alpha, teacher + n4 * 4,
      student + n4 * 4, tail);
  return mmt::check_launch("ema");
}

extern "C" int mmt_sgd_ema_f32(float* student, float* teacher, const float* grad, float* momentum_buf,
                               int64_t n, const mmt_param_segment* segments_dev, int n_segments,
                               float momentum, float alpha, int first_step, float grad_scale,
                               void* stream_) {
  MMT_REQUIRE(n >= 0 && n_segments >= 0, "sgd_ema: bad sizes");
  if (n == 0 || n_segments == 0) return MMT_OK;
  MMT_REQUIRE(student && grad && momentum_buf && segments_dev, "sgd_ema: null pointer");
  MMT_REQUIRE(n % 4 == 0, "sgd_ema: the arenas must be padded to a multiple of 4 elements (got %lld)", (long long)n);
  MMT_REQUIRE(((uintptr_t)student % 16 == 0) && ((uintptr_t)grad % 16 == 0) &&
                  ((uintptr_t)momentum_buf % 16 == 0) && (!teacher || (uintptr_t)teacher % 16 == 0),
              "sgd_ema: arenas must be 16-byte aligned");
  // student / momentum read+write, grad read, teacher read+write
  mmt::ProfScope prof(MMT_PROF_OPTIM, 0.0, 4.0 * n * (teacher ? 7.0 : 5.0), (cudaStream_t)stream_);
  sgd_ema_kernel<<<stream_grid((n + 3) / 4), 256, 0, (cudaStream_t)stream_>>>(
      student, teacher, grad, momentum_buf, n, segments_dev, n_segments, momentum, alpha,
      1.f - alpha, first_step, grad_scale);
  return mmt::check_launch("sgd_ema");
}
