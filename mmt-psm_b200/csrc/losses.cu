// Fused loss kernels: SigmoidFocalLoss, PSM (perturbation-sensitive sample mining) and MGD
// (mask-guided feature distillation).  sm_100a, fp32, HBM/latency bound — no tensor cores.
//
// Reference semantics:
//   focal : upstream maskrcnn-benchmark csrc/cuda/SigmoidFocalLoss_cuda.cu (absent from this tree)
//   PSM   : maskrcnn_benchmark/modeling/roi_heads/box_head/loss.py:164-237,267-287,311-315
//   MGD   : maskrcnn_benchmark/modeling/detector/generalized_rcnn.py:243-282
#include <float.h>
#include <algorithm>

#include "common.cuh"

namespace {

// ------------------------------------------------------------------------------------ focal
__global__ void __launch_bounds__(256) focal_fwd_kernel(const float* __restrict__ logits,
                                                        const int* __restrict__ targets,
                                                        int num_classes, float gamma, float alpha,
                                                        float* __restrict__ losses, long long total) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int n = (int)(i / num_classes), d = (int)(i % num_classes);
    const int t = targets[n];
    const float c1 = (t == d + 1) ? 1.f : 0.f;
    const float c2 = (t >= 0 && t != d + 1) ? 1.f : 0.f;
    const float x = logits[i];
    const float p = 1.f / (1.f + expf(-x));
    const float term1 = powf(1.f - p, gamma) * logf(fmaxf(p, FLT_MIN));
    const float pos = (x >= 0.f) ? 1.f : 0.f;
    const float term2 = powf(p, gamma) * (-1.f * x * pos - logf(1.f + expf(x - 2.f * x * pos)));
    losses[i] = -c1 * term1 * alpha + -c2 * term2 * (1.f - alpha);
  }
}

__global__ void __launch_bounds__(256) focal_bwd_kernel(const float* __restrict__ logits,
                                                        const int* __restrict__ targets,
                                                        const float* __restrict__ dloss,
                                                        int num_classes, float gamma, float alpha,
                                                        float* __restrict__ dlogits, long long total) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int n = (int)(i / num_classes), d = (int)(i % num_classes);
    const int t = targets[n];
    const float c1 = (t == d + 1) ? 1.f : 0.f;
    const float c2 = (t >= 0 && t != d + 1) ? 1.f : 0.f;
    const float x = logits[i];
    const float p = 1.f / (1.f + expf(-x));
    const float term1 = powf(1.f - p, gamma) * (1.f - p - (p * gamma * logf(fmaxf(p, FLT_MIN))));
    const float pos = (x >= 0.f) ? 1.f : 0.f;
    const float term2 =
        powf(p, gamma) *
        ((-1.f * x * pos - logf(1.f + expf(x - 2.f * x * pos))) * (1.f - p) * gamma - p);
    dlogits[i] = (-c1 * term1 * alpha + -c2 * term2 * (1.f - alpha)) * dloss[i];
  }
}

// -------------------------------------------------------------------------------------- PSM
constexpr int kMaxClasses = 16;
constexpr int kMaxViews = 16;

// One CTA, any n.  Workspace: v[n] float | kind[n] int (0 unused, 1 pos, 2 neg) | ord[n] int.
__global__ void __launch_bounds__(1024) psm_kernel(
    const float* __restrict__ student, const float* __restrict__ teacher, int n_views,
    const long long* __restrict__ labels, int n, int C, float temperature, int sharpen,
    int hard_negative, float neg_weight, const int* __restrict__ neg_order, float* __restrict__ loss_out,
    float* __restrict__ grad_student, int* __restrict__ selected_out, float* __restrict__ v,
    int* __restrict__ kind, int* __restrict__ ord) {
  __shared__ float red[32];
  __shared__ int s_counts[2];
  const int tid = threadIdx.x;

  // pass 1: per-row uncertainty v and pos/neg classification
  int my_pos = 0, my_neg = 0;
  for (int i = tid; i < n; i += blockDim.x) {
    const long long lab = labels[i];
    const int k = lab > 0 ? 1 : (lab == 0 ? 2 : 0);
    kind[i] = k;
    my_pos += (k == 1);
    my_neg += (k == 2);
    double vsum = 0.0;
    if (k != 0) {
      double mean[kMaxClasses], m2[kMaxClasses];
      for (int c = 0; c < C; ++c) mean[c] = 0.0, m2[c] = 0.0;
      // two passes over the views (probabilities are recomputed; n_views*C is tiny)
      for (int pass = 0; pass < 2; ++pass) {
        for (int kv = 0; kv < n_views; ++kv) {
          const float* z = teacher + ((long long)kv * n + i) * C;
          float mx = -INFINITY;
          for (int c = 0; c < C; ++c) mx = fmaxf(mx, z[c]);
          float den = 0.f;
          for (int c = 0; c < C; ++c) den += expf(z[c] - mx);
          for (int c = 0; c < C; ++c) {
            const double p = (double)(expf(z[c] - mx) / den);
            if (pass == 0) mean[c] += p;
            else { const double dd = p - mean[c]; m2[c] += dd * dd; }
          }
        }
        if (pass == 0) for (int c = 0; c < C; ++c) mean[c] /= (double)n_views;
      }
      for (int c = 0; c < C; ++c) vsum += sqrt(m2[c] / (double)(n_views - 1));
    }
    v[i] = (float)vsum;
  }
  {
    const float tp = mmt::block_sum((float)my_pos, red);
    const float tn = mmt::block_sum((float)my_neg, red);
    if (tid == 0) { s_counts[0] = (int)(tp + 0.5f); s_counts[1] = (int)(tn + 0.5f); }
  }
  __syncthreads();  // also publishes v[]/kind[] (global writes by this CTA) to the CTA
  const int P = s_counts[0], Nn = s_counts[1];
  const int m = min(Nn, P / 2);
  const int R = P + m;

  // pass 2: rank negatives (descending v, ties -> lower index) or take the caller's order
  for (int i = tid; i < n; i += blockDim.x) {
    int sel = 0;
    if (kind[i] == 1) sel = 1;
    else if (kind[i] == 2) {
      if (hard_negative) {
        const float vi = v[i];
        int rank = 0;
        for (int j = 0; j < n; ++j)
          if (kind[j] == 2) { const float vj = v[j]; rank += (vj > vi) || (vj == vi && j < i); }
        sel = (rank < m) ? 2 : 0;
      } else {
        int ordinal = 0;
        for (int j = 0; j < i; ++j) ordinal += (kind[j] == 2);
        ord[i] = ordinal;
      }
    }
    if (hard_negative || kind[i] != 2) selected_out[i] = sel;
  }
  if (!hard_negative) {
    __syncthreads();
    for (int i = tid; i < n; i += blockDim.x) {
      if (kind[i] != 2) continue;
      int sel = 0;
      for (int t = 0; t < m; ++t) if (neg_order[t] == ord[i]) { sel = 2; break; }
      selected_out[i] = sel;
    }
  }
  __syncthreads();

  // pass 3: soft-target CE on the selected rows + analytic gradient
  const float inv_R3 = 1.f / ((float)R * 3.f);  // R == 0 -> inf; 0 * inf = NaN as torch's empty mean
  const float inv_T = 1.f / temperature;
  float my_loss = 0.f;
  for (int i = tid; i < n; i += blockDim.x) {
    const int sel = selected_out[i];
    float* g = grad_student + (long long)i * C;
    if (sel == 0) { for (int c = 0; c < C; ++c) g[c] = 0.f; continue; }
    const float w = (sel == 1) ? 1.f : neg_weight;
    float q[kMaxClasses], ls[kMaxClasses];
    float mx = -INFINITY;
    for (int c = 0; c < C; ++c) {
      float acc = 0.f;
      for (int kv = 0; kv < n_views; ++kv) acc += teacher[((long long)kv * n + i) * C + c];
      q[c] = acc / (float)n_views;
      mx = fmaxf(mx, q[c]);
    }
    float den = 0.f;
    for (int c = 0; c < C; ++c) { q[c] = expf(q[c] - mx); den += q[c]; }
    for (int c = 0; c < C; ++c) q[c] /= den;
    if (sharpen) {
      float s = 0.f;
      for (int c = 0; c < C; ++c) { q[c] = powf(q[c], inv_T); s += q[c]; }
      for (int c = 0; c < C; ++c) q[c] /= s;
    }
    const float* sl = student + (long long)i * C;
    float smx = -INFINITY;
    for (int c = 0; c < C; ++c) smx = fmaxf(smx, sl[c]);
    float sden = 0.f;
    for (int c = 0; c < C; ++c) sden += expf(sl[c] - smx);
    const float lse = logf(sden);
    float qsum = 0.f, row = 0.f;
    for (int c = 0; c < C; ++c) { ls[c] = sl[c] - smx - lse; row -= q[c] * ls[c]; qsum += q[c]; }
    my_loss += w * row;
    for (int c = 0; c < C; ++c) g[c] = w * inv_R3 * (expf(ls[c]) * qsum - q[c]);
  }
  const float total = mmt::block_sum(my_loss, red);
  if (tid == 0) loss_out[0] = (R == 0) ? __int_as_float(0x7fc00000) : total * inv_R3;
}

// -------------------------------------------------------------------------------------- MGD
struct LevelTable {
  int n_levels;
  int h[8], w[8];
  long long off[9];
};

// One WARP per output pixel: the lanes sweep the pooling window (up to 62 x 62 inputs at the coarsest level — a
// thread per output serialised 3844 strided loads there).  The map holds small integers (a sum of binary masks), so
// the fp32 window sum is exact in any order.
__global__ void __launch_bounds__(256) mgd_masks_kernel(const float* __restrict__ M, int batch,
                                                        int H, int W, LevelTable tab,
                                                        unsigned char* __restrict__ masks,
                                                        float* __restrict__ sums) {
  const long long total = tab.off[tab.n_levels];
  __shared__ float s_level[8];
  if (threadIdx.x < 8) s_level[threadIdx.x] = 0.f;
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const long long warp0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long n_warps = ((long long)gridDim.x * blockDim.x) >> 5;
  for (long long idx = warp0; idx < total; idx += n_warps) {
    int l = 0;
    while (idx >= tab.off[l + 1]) ++l;
    const int h = tab.h[l], w = tab.w[l];
    long long r = idx - tab.off[l];
    const int ox = (int)(r % w); r /= w;
    const int oy = (int)(r % h);
    const int b = (int)(r / h);
    // adaptive_avg_pool2d window: [floor(o*I/O), ceil((o+1)*I/O))
    const int y0 = (int)(((long long)oy * H) / h), y1 = (int)((((long long)oy + 1) * H + h - 1) / h);
    const int x0 = (int)(((long long)ox * W) / w), x1 = (int)((((long long)ox + 1) * W + w - 1) / w);
    const int ww = x1 - x0, area = (y1 - y0) * ww;
    const float* src = M + (long long)b * H * W;
    float s = 0.f;
    for (int a = lane; a < area; a += 32) s += src[(long long)(y0 + a / ww) * W + x0 + a % ww];
    s = mmt::warp_sum(s);
    if (lane == 0) {
      const float mean = s / (float)area;
      const unsigned char bit = mean > 0.5f ? 1 : 0;
      masks[idx] = bit;
      if (bit) atomicAdd(&s_level[l], 1.f);
    }
  }
  __syncthreads();
  if (threadIdx.x < tab.n_levels && s_level[threadIdx.x] != 0.f)
    atomicAdd(&sums[threadIdx.x], s_level[threadIdx.x]);
}

struct ViewPtrs {
  const float* p[kMaxViews];
};

// One warp per pixel; lanes sweep the channel vector (C = 128 -> one float4 per lane).
template <int VEC>
__global__ void __launch_bounds__(256) mgd_level_kernel(
    const float* __restrict__ S, mmt_strides4 ss, ViewPtrs T, mmt_strides4 ts, int n_views,
    unsigned flip_mask, const unsigned char* __restrict__ mask, const float* __restrict__ mask_sum,
    int batch, int C, int H, int W, float term_weight, float* __restrict__ partial_out,
    float* __restrict__ grad, mmt_strides4 gs) {
  const int lane = threadIdx.x & 31;
  const long long warp_global = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long n_warps = ((long long)gridDim.x * blockDim.x) >> 5;
  const long long pixels = (long long)batch * H * W;
  const float coeff = 2.f * term_weight / (mask_sum[0] * (float)C + 1e-7f);
  float num[kMaxViews];
#pragma unroll
  for (int k = 0; k < kMaxViews; ++k) num[k] = 0.f;

  for (long long pix = warp_global; pix < pixels; pix += n_warps) {
    const int x = (int)(pix % W);
    const int y = (int)((pix / W) % H);
    const int b = (int)(pix / ((long long)W * H));
    const bool on = mask[pix] != 0;
    for (int c = lane * VEC; c < C; c += 32 * VEC) {
      float gacc[VEC];
#pragma unroll
      for (int v = 0; v < VEC; ++v) gacc[v] = 0.f;
      if (on) {
        float s[VEC];
        const float* sp = S + b * ss.n + y * ss.h + x * ss.w + c * ss.c;
        if (VEC == 4) {
          const float4 t4 = __ldg(reinterpret_cast<const float4*>(sp));
          s[0] = t4.x; s[1 % VEC] = t4.y; s[2 % VEC] = t4.z; s[3 % VEC] = t4.w;
        } else s[0] = __ldg(sp);
#pragma unroll
        for (int k = 0; k < kMaxViews; ++k) {
          if (k < n_views) {
            const int tx = ((flip_mask >> k) & 1u) ? (W - 1 - x) : x;
            const float* tp = T.p[k] + b * ts.n + y * ts.h + tx * ts.w + c * ts.c;
            float t[VEC];
            if (VEC == 4) {
              const float4 t4 = __ldg(reinterpret_cast<const float4*>(tp));
              t[0] = t4.x; t[1 % VEC] = t4.y; t[2 % VEC] = t4.z; t[3 % VEC] = t4.w;
            } else t[0] = __ldg(tp);
#pragma unroll
            for (int v = 0; v < VEC; ++v) {
              const float d = s[v] - t[v];
              num[k] += d * d;
              gacc[v] += d;
            }
          }
        }
      }
      if (grad != nullptr) {
        float* gp = grad + b * gs.n + y * gs.h + x * gs.w + c * gs.c;
        if (VEC == 4) {
          float4 o;
          o.x = coeff * gacc[0]; o.y = coeff * gacc[1 % VEC];
          o.z = coeff * gacc[2 % VEC]; o.w = coeff * gacc[3 % VEC];
          *reinterpret_cast<float4*>(gp) = o;
        } else gp[0] = coeff * gacc[0];
      }
    }
  }
  __shared__ float s_part[kMaxViews];
  if (threadIdx.x < kMaxViews) s_part[threadIdx.x] = 0.f;
  __syncthreads();
#pragma unroll
  for (int k = 0; k < kMaxViews; ++k) {
    if (k < n_views) {
      const float r = mmt::warp_sum(num[k]);
      if (lane == 0 && r != 0.f) atomicAdd(&s_part[k], r);
    }
  }
  __syncthreads();
  if ((int)threadIdx.x < n_views && s_part[threadIdx.x] != 0.f)
    atomicAdd(&partial_out[threadIdx.x], s_part[threadIdx.x]);
}

__global__ void mgd_finalize_kernel(const float* __restrict__ partials, const float* __restrict__ sums,
                                    int n_levels, int n_views, int C, float* __restrict__ loss) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    float acc = 0.f;
    // same term order as the reference loop (views outer, levels inner), generalized_rcnn.py:266-273
    for (int k = 0; k < n_views; ++k)
      for (int l = 0; l < n_levels; ++l)
        acc += partials[l * n_views + k] / (sums[l] * (float)C + 1e-7f);
    loss[0] = acc / (float)(n_levels * n_views);
  }
}

inline int grid_for(long long total, int block, int per_sm) {
  long long blocks = (total + block - 1) / block;
  const long long cap = (long long)mmt::kNumSMs * per_sm;
  if (blocks > cap) blocks = cap;
  return (int)(blocks < 1 ? 1 : blocks);
}

}  // namespace

extern "C" int mmt_sigmoid_focal_loss_forward_f32(const float* logits, const int* targets, int n,
                                                  int num_classes, float gamma, float alpha,
                                                  float* losses, void* stream_) {
  MMT_REQUIRE(n >= 0 && num_classes > 0, "focal_forward: bad sizes");
  if (n == 0) return MMT_OK;
  MMT_REQUIRE(logits && targets && losses, "focal_forward: null pointer");
  const long long total = (long long)n * num_classes;
  focal_fwd_kernel<<<grid_for(total, 256, 16), 256, 0, (cudaStream_t)stream_>>>(
      logits, targets, num_classes, gamma, alpha, losses, total);
  return mmt::check_launch("focal_forward");
}

extern "C" int mmt_sigmoid_focal_loss_backward_f32(const float* logits, const int* targets,
                                                   const float* grad_losses, int n, int num_classes,
                                                   float gamma, float alpha, float* grad_logits,
                                                   void* stream_) {
  MMT_REQUIRE(n >= 0 && num_classes > 0, "focal_backward: bad sizes");
  if (n == 0) return MMT_OK;
  MMT_REQUIRE(logits && targets && grad_losses && grad_logits, "focal_backward: null pointer");
  const long long total = (long long)n * num_classes;
  focal_bwd_kernel<<<grid_for(total, 256, 16), 256, 0, (cudaStream_t)stream_>>>(
      logits, targets, grad_losses, num_classes, gamma, alpha, grad_logits, total);
  return mmt::check_launch("focal_backward");
}

extern "C" size_t mmt_psm_workspace_bytes(int n) {
  return mmt::align_up((size_t)(n > 0 ? n : 1) * 4, 256) * 3;
}

extern "C" int mmt_psm_loss_f32(const float* student, const float* teacher, int n_views,
                                const int64_t* labels, int n, int C, float temperature, int sharpen,
                                int hard_negative, float neg_weight, const int* neg_order,
                                float* loss_out, float* grad_student, int* selected_out,
                                void* workspace, size_t workspace_bytes, void* stream_) {
  MMT_REQUIRE(n >= 0, "psm: negative row count");
  MMT_REQUIRE(C >= 1 && C <= kMaxClasses, "psm: num_classes must be in [1,%d]", kMaxClasses);
  MMT_REQUIRE(n_views >= 2 && n_views <= kMaxViews, "psm: n_views must be in [2,%d]", kMaxViews);
  MMT_REQUIRE(temperature > 0.f, "psm: temperature must be positive");
  MMT_REQUIRE(loss_out && (n == 0 || (student && teacher && labels && grad_student && selected_out)),
              "psm: null pointer");
  MMT_REQUIRE(hard_negative || neg_order || n == 0, "psm: neg_order required when hard_negative == 0");
  const size_t need = mmt_psm_workspace_bytes(n);
  if (workspace_bytes < need || workspace == nullptr)
    return mmt::fail(MMT_ERR_WORKSPACE_TOO_SMALL, "psm: workspace %zu < required %zu",
                     workspace_bytes, need);
  const size_t stride = need / 3;
  char* p = (char*)workspace;
  psm_kernel<<<1, 1024, 0, (cudaStream_t)stream_>>>(
      student, teacher, n_views, (const long long*)labels, n, C, temperature, sharpen, hard_negative,
      neg_weight, neg_order, loss_out, grad_student, selected_out, (float*)p, (int*)(p + stride),
      (int*)(p + 2 * stride));
  return mmt::check_launch("psm_loss");
}

extern "C" int mmt_mgd_masks_f32(const float* integral_mask, int batch, int height, int width,
                                 const int* level_h, const int* level_w, int n_levels,
                                 uint8_t* masks_out, float* mask_sums_out, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  MMT_REQUIRE(n_levels >= 1 && n_levels <= 8, "mgd_masks: n_levels must be in [1,8]");
  MMT_REQUIRE(batch > 0 && height > 0 && width > 0, "mgd_masks: bad sizes");
  MMT_REQUIRE(integral_mask && level_h && level_w && masks_out && mask_sums_out, "mgd_masks: null pointer");
  LevelTable tab;
  tab.n_levels = n_levels;
  tab.off[0] = 0;
  for (int l = 0; l < n_levels; ++l) {
    MMT_REQUIRE(level_h[l] > 0 && level_w[l] > 0, "mgd_masks: bad level size");
    tab.h[l] = level_h[l];
    tab.w[l] = level_w[l];
    tab.off[l + 1] = tab.off[l] + (long long)batch * level_h[l] * level_w[l];
  }
  MMT_CUDA(cudaMemsetAsync(mask_sums_out, 0, sizeof(float) * n_levels, stream));
  mgd_masks_kernel<<<grid_for(tab.off[n_levels] * 32, 256, 8), 256, 0, stream>>>(
      integral_mask, batch, height, width, tab, masks_out, mask_sums_out);
  return mmt::check_launch("mgd_masks");
}

extern "C" int mmt_mgd_level_f32(const float* student, mmt_strides4 ss, const float* const* teacher_views,
                                 mmt_strides4 ts, int n_views, unsigned flip_mask,
                                 const uint8_t* level_mask, const float* level_mask_sum, int batch,
                                 int channels, int height, int width, float term_weight,
                                 float* partial_out, float* grad_student, mmt_strides4 gs,
                                 void* stream_) {
  MMT_REQUIRE(n_views >= 1 && n_views <= kMaxViews, "mgd_level: n_views must be in [1,%d]", kMaxViews);
  MMT_REQUIRE(batch > 0 && channels > 0 && height > 0 && width > 0, "mgd_level: bad sizes");
  MMT_REQUIRE(student && teacher_views && level_mask && level_mask_sum && partial_out, "mgd_level: null pointer");
  ViewPtrs T;
  for (int k = 0; k < kMaxViews; ++k) T.p[k] = k < n_views ? teacher_views[k] : nullptr;
  for (int k = 0; k < n_views; ++k) MMT_REQUIRE(T.p[k] != nullptr, "mgd_level: null teacher view %d", k);
  auto aligned4 = [](const void* p, const mmt_strides4& s) {
    return ((uintptr_t)p % 16 == 0) && s.c == 1 && s.n % 4 == 0 && s.h % 4 == 0 && s.w % 4 == 0;
  };
  bool vec4 = channels % 4 == 0 && aligned4(student, ss) && (!grad_student || aligned4(grad_student, gs));
  for (int k = 0; k < n_views; ++k) vec4 = vec4 && aligned4(T.p[k], ts);
  const long long pixels = (long long)batch * height * width;
  const int grid = grid_for(pixels * 32, 256, 8);
  if (vec4)
    mgd_level_kernel<4><<<grid, 256, 0, (cudaStream_t)stream_>>>(
        student, ss, T, ts, n_views, flip_mask, level_mask, level_mask_sum, batch, channels, height,
        width, term_weight, partial_out, grad_student, gs);
  else
    mgd_level_kernel<1><<<grid, 256, 0, (cudaStream_t)stream_>>>(
        student, ss, T, ts, n_views, flip_mask, level_mask, level_mask_sum, batch, channels, height,
        width, term_weight, partial_out, grad_student, gs);
  return mmt::check_launch("mgd_level");
}

extern "C" int mmt_mgd_finalize_f32(const float* partials, const float* mask_sums, int n_levels,
                                    int n_views, int channels, float* loss_out, void* stream_) {
  MMT_REQUIRE(partials && mask_sums && loss_out && n_levels > 0 && n_views > 0, "mgd_finalize: bad arguments");
  mgd_finalize_kernel<<<1, 32, 0, (cudaStream_t)stream_>>>(partials, mask_sums, n_levels, n_views,
                                                           channels, loss_out);
  return mmt::check_launch("mgd_finalize");
}

// ------------------------------------------------------------------------------------ RPN loss
// Reference modeling/rpn/loss.py:138-196: over the sampled anchors of the batch,
//   objectness_loss = mean BCE-with-logits(objectness[sampled], labels[sampled])
//   box_loss        = sum smooth_l1(deltas[pos], targets[pos], beta = 1/9) / |sampled|
// The reference flattens the five per-level head maps (permute + reshape + cat), gathers the sampled rows with two
// `nonzero`s and runs ~35 elementwise launches, then as many again in autograd's backward.  Here one kernel reads the
// fused head maps [N, h*w, Cp] of all levels in place (channels [0,A) objectness, [A,5A) deltas) and reduces both sums
// under the sampling masks; a second one writes the analytic gradient straight into maps of the same layout.
namespace {

constexpr int kRpnMaxLevels = 8;
struct RpnLevels {
  int n_levels;
  const float* fused[kRpnMaxLevels];   // [N, pixels, Cp]
  float* dfused[kRpnMaxLevels];        // same layout (backward only)
  long long pixels[kRpnMaxLevels];
  long long anchor_start[kRpnMaxLevels + 1];   // first anchor of the level inside one image's anchor list
};

__device__ __forceinline__ int rpn_level_of(const RpnLevels& L, long long a) {
  int l = 0;
  while (l + 1 < L.n_levels && a >= L.anchor_start[l + 1]) ++l;
  return l;
}

__device__ __forceinline__ float bce_logits(float x, float y) {   // torch: max(x,0) - x*y + log1p(exp(-|x|))
  return fmaxf(x, 0.f) - x * y + log1pf(expf(-fabsf(x)));
}

__global__ void __launch_bounds__(256) rpn_loss_fwd_kernel(RpnLevels L, int A, int Cp, int n_images,
                                                           const float* __restrict__ labels,
                                                           const float4* __restrict__ targets,
                                                           const unsigned char* __restrict__ pos,
                                                           const unsigned char* __restrict__ neg, float beta,
                                                           float* __restrict__ sums /*[2]: bce, smooth-l1*/) {
  __shared__ float s_red[32];
  const long long per_image = L.anchor_start[L.n_levels];
  const long long total = per_image * n_images;
  float bce = 0.f, sl1 = 0.f;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total;
       t += (long long)gridDim.x * blockDim.x) {
    const bool p = pos[t] != 0, q = neg[t] != 0;
    if (!(p || q)) continue;
    const long long img = t / per_image, a_img = t - img * per_image;
    const int l = rpn_level_of(L, a_img);
    const long long a_lvl = a_img - L.anchor_start[l];
    const long long pix = a_lvl / A;
    const int a = (int)(a_lvl - pix * A);
    const float* row = L.fused[l] + (img * L.pixels[l] + pix) * Cp;
    bce += bce_logits(row[a], fmaxf(labels[t], 0.f));
    if (p) {
      const float4 tg = targets[t];
      const float d[4] = {row[A + 4 * a] - tg.x, row[A + 4 * a + 1] - tg.y, row[A + 4 * a + 2] - tg.z,
                          row[A + 4 * a + 3] - tg.w};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float n = fabsf(d[e]);
        sl1 += (n < beta) ? 0.5f * n * n / beta : n - 0.5f * beta;
      }
    }
  }
  bce = mmt::block_sum(bce, s_red);
  sl1 = mmt::block_sum(sl1, s_red);
  if (threadIdx.x == 0) {
    atomicAdd(sums, bce);
    atomicAdd(sums + 1, sl1);
  }
}

// d(loss)/d(fused): every element of every map is written (zeros off the sampled anchors and in the padding channels)
__global__ void __launch_bounds__(256) rpn_loss_bwd_kernel(RpnLevels L, int A, int Cp, int n_images,
                                                           const float* __restrict__ labels,
                                                           const float4* __restrict__ targets,
                                                           const unsigned char* __restrict__ pos,
                                                           const unsigned char* __restrict__ neg, float beta,
                                                           const float* __restrict__ n_sampled,
                                                           const float* __restrict__ g_obj,
                                                           const float* __restrict__ g_box) {
  const long long per_image = L.anchor_start[L.n_levels];
  const long long total = per_image * n_images;
  const float inv_n = 1.f / *n_sampled;
  const float go = *g_obj * inv_n, gb = *g_box * inv_n;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total;
       t += (long long)gridDim.x * blockDim.x) {
    const long long img = t / per_image, a_img = t - img * per_image;
    const int l = rpn_level_of(L, a_img);
    const long long a_lvl = a_img - L.anchor_start[l];
    const long long pix = a_lvl / A;
    const int a = (int)(a_lvl - pix * A);
    const long long off = (img * L.pixels[l] + pix) * Cp;
    const float* row = L.fused[l] + off;
    float* drow = L.dfused[l] + off;
    const bool p = pos[t] != 0, q = neg[t] != 0;
    float dobj = 0.f;
    float dd[4] = {0.f, 0.f, 0.f, 0.f};
    if (p || q) {
      const float x = row[a];
      dobj = (1.f / (1.f + expf(-x)) - fmaxf(labels[t], 0.f)) * go;
      if (p) {
        const float4 tg = targets[t];
        const float d[4] = {row[A + 4 * a] - tg.x, row[A + 4 * a + 1] - tg.y, row[A + 4 * a + 2] - tg.z,
                            row[A + 4 * a + 3] - tg.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float n = fabsf(d[e]);
          dd[e] = ((n < beta) ? d[e] / beta : (d[e] > 0.f ? 1.f : (d[e] < 0.f ? -1.f : 0.f))) * gb;
        }
      }
    }
    drow[a] = dobj;
#pragma unroll
    for (int e = 0; e < 4; ++e) drow[A + 4 * a + e] = dd[e];
    if (a == 0)
      for (int c = 5 * A; c < Cp; ++c) drow[c] = 0.f;
  }
}

int fill_levels(RpnLevels& L, const float* const* fused, float* const* dfused, const int64_t* pixels, int n_levels, int A) {
  if (n_levels < 1 || n_levels > kRpnMaxLevels) return mmt::fail(MMT_ERR_INVALID_ARGUMENT, "rpn_loss: 1..%d levels", kRpnMaxLevels);
  L.n_levels = n_levels;
  long long acc = 0;
  for (int l = 0; l < n_levels; ++l) {
    if (!fused[l] || pixels[l] <= 0) return mmt::fail(MMT_ERR_INVALID_ARGUMENT, "rpn_loss: bad level %d", l);
    L.fused[l] = fused[l];
    L.dfused[l] = dfused ? dfused[l] : nullptr;
    L.pixels[l] = pixels[l];
    L.anchor_start[l] = acc;
    acc += pixels[l] * A;
  }
  L.anchor_start[n_levels] = acc;
  return MMT_OK;
}

}  // namespace

extern "C" int mmt_rpn_loss_forward_f32(const float* const* fused_levels, const int64_t* level_pixels, int n_levels,
                                        int num_anchors, int padded_channels, int n_images, const float* labels,
                                        const float* regression_targets, const unsigned char* pos_mask,
                                        const unsigned char* neg_mask, float beta, float* sums_out, void* stream_) {
  MMT_REQUIRE(labels && regression_targets && pos_mask && neg_mask && sums_out && fused_levels && level_pixels,
              "rpn_loss_forward: null pointer");
  MMT_REQUIRE(padded_channels >= 5 * num_anchors && num_anchors > 0 && n_images > 0 && beta > 0.f, "rpn_loss_forward: bad sizes");
  RpnLevels L;
  if (int rc = fill_levels(L, fused_levels, nullptr, level_pixels, n_levels, num_anchors)) return rc;
  cudaStream_t stream = (cudaStream_t)stream_;
  MMT_CUDA(cudaMemsetAsync(sums_out, 0, 2 * sizeof(float), stream));
  const long long total = L.anchor_start[n_levels] * n_images;
  const int blocks = (int)std::min<long long>((total + 255) / 256, (long long)mmt::kNumSMs * 8);
  rpn_loss_fwd_kernel<<<blocks, 256, 0, stream>>>(L, num_anchors, padded_channels, n_images, labels,
                                                  reinterpret_cast<const float4*>(regression_targets), pos_mask,
                                                  neg_mask, beta, sums_out);
  return mmt::check_launch("rpn_loss_fwd");
}

extern "C" int mmt_rpn_loss_backward_f32(const float* const* fused_levels, float* const* dfused_levels,
                                         const int64_t* level_pixels, int n_levels, int num_anchors,
                                         int padded_channels, int n_images, const float* labels,
                                         const float* regression_targets, const unsigned char* pos_mask,
                                         const unsigned char* neg_mask, float beta, const float* n_sampled,
                                         const float* grad_objectness_loss, const float* grad_box_loss, void* stream_) {
  MMT_REQUIRE(labels && regression_targets && pos_mask && neg_mask && fused_levels && dfused_levels && level_pixels &&
                  n_sampled && grad_objectness_loss && grad_box_loss,
              "rpn_loss_backward: null pointer");
  MMT_REQUIRE(padded_channels >= 5 * num_anchors && num_anchors > 0 && n_images > 0 && beta > 0.f, "rpn_loss_backward: bad sizes");
  RpnLevels L;
  if (int rc = fill_levels(L, fused_levels, dfused_levels, level_pixels, n_levels, num_anchors)) return rc;
  for (int l = 0; l < n_levels; ++l) MMT_REQUIRE(dfused_levels[l] != nullptr, "rpn_loss_backward: null gradient map");
  const long long total = L.anchor_start[n_levels] * n_images;
  const int blocks = (int)std::min<long long>((total + 255) / 256, (long long)mmt::kNumSMs * 8);
  rpn_loss_bwd_kernel<<<blocks, 256, 0, (cudaStream_t)stream_>>>(
      L, num_anchors, padded_channels, n_images, labels, reinterpret_cast<const float4*>(regression_targets), pos_mask,
      neg_mask, beta, n_sampled, grad_objectness_loss, grad_box_loss);
  return mmt::check_launch("rpn_loss_bwd");
}
