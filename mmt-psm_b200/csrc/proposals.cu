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

// ------------------------------------------------------------------------------------------------------------
// Fused IoU + Matcher: for every prediction (anchor / proposal) the best-overlapping ground-truth box and the
// Matcher's threshold codes, without materialising the [G, A] IoU matrix.
//
// Reference chain:  structures/boxlist_ops.py:57-92 (boxlist_iou: +1 extents, [G,A] matrix through ~10 elementwise
// passes) -> modeling/matcher.py:44-138 (max over G, BELOW_LOW_THRESHOLD = -1, BETWEEN_THRESHOLDS = -2, and with
// allow_low_quality_matches every prediction attaining some ground truth's best IoU gets its arg-max back).
// At RPN scale that matrix is 40 x 159 882 floats per image; here each thread owns one prediction and walks the G
// boxes held in shared memory.  Same fp32 expression per pair as the reference, so `==` against the per-GT maxima
// selects exactly the pairs the reference's `quality == highest_quality_foreach_gt[:, None]` selects.
namespace {

constexpr int kMaxGtSmem = 1024;

__device__ __forceinline__ float pair_iou(const float4 g, const float ga, const float4 b, const float ba) {
  const float lx = fmaxf(g.x, b.x), ly = fmaxf(g.y, b.y);
  const float rx = fminf(g.z, b.z), ry = fminf(g.w, b.w);
  const float w = fmaxf(__fadd_rn(__fsub_rn(rx, lx), 1.f), 0.f), h = fmaxf(__fadd_rn(__fsub_rn(ry, ly), 1.f), 0.f);
  const float inter = __fmul_rn(w, h);
  return __fdiv_rn(inter, __fsub_rn(__fadd_rn(ga, ba), inter));
}
__device__ __forceinline__ float box_area(const float4 b) {
  return __fmul_rn(__fadd_rn(__fsub_rn(b.z, b.x), 1.f), __fadd_rn(__fsub_rn(b.w, b.y), 1.f));
}

// pass 1: per prediction max / arg-max over the ground truth (first maximum wins, like torch.max) + per-GT maxima
__global__ void __launch_bounds__(256) match_pass1_kernel(const float4* __restrict__ gt, int G,
                                                          const float4* __restrict__ boxes, int A,
                                                          float* __restrict__ best_val, int* __restrict__ best_idx,
                                                          unsigned int* __restrict__ best_per_gt) {
  extern __shared__ float4 s_gt[];  // G boxes then G areas (as float)
  float* s_area = reinterpret_cast<float*>(s_gt + G);
  for (int g = threadIdx.x; g < G; g += blockDim.x) {
    const float4 b = __ldg(gt + g);
    s_gt[g] = b;
    s_area[g] = box_area(b);
  }
  __syncthreads();
  const int a = blockIdx.x * blockDim.x + threadIdx.x;
  const bool live = a < A;
  float4 b = make_float4(0.f, 0.f, 0.f, 0.f);
  if (live) b = __ldg(boxes + a);
  const float ba = box_area(b);
  float best = -1.f;
  int arg = 0;
  const int lane = threadIdx.x & 31;
  for (int g = 0; g < G; ++g) {
    const float v = live ? pair_iou(s_gt[g], s_area[g], b, ba) : 0.f;
    if (v > best) { best = v; arg = g; }
    // IoU >= 0: the IEEE bit pattern orders like the value, so an unsigned max is a float max
    unsigned int m = __float_as_uint(v);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, o));
    if (lane == 0 && m != 0u) atomicMax(best_per_gt + g, m);
  }
  if (live) { best_val[a] = best; best_idx[a] = arg; }
}

// pass 2: threshold codes, and the low-quality restore
__global__ void __launch_bounds__(256) match_pass2_kernel(const float4* __restrict__ gt, int G,
                                                          const float4* __restrict__ boxes, int A,
                                                          const float* __restrict__ best_val,
                                                          const int* __restrict__ best_idx,
                                                          const unsigned int* __restrict__ best_per_gt, float high_thr,
                                                          float low_thr, int allow_low_quality,
                                                          long long* __restrict__ matches) {
  extern __shared__ float4 s_gt[];
  float* s_area = reinterpret_cast<float*>(s_gt + G);
  float* s_best = s_area + G;
  if (allow_low_quality) {
    for (int g = threadIdx.x; g < G; g += blockDim.x) {
      const float4 b = __ldg(gt + g);
      s_gt[g] = b;
      s_area[g] = box_area(b);
      s_best[g] = __uint_as_float(best_per_gt[g]);
    }
    __syncthreads();
  }
  const int a = blockIdx.x * blockDim.x + threadIdx.x;
  if (a >= A) return;
  const float v = best_val[a];
  long long m = best_idx[a];
  const long long all_m = m;
  if (v < low_thr) m = -1;
  else if (v < high_thr) m = -2;
  if (allow_low_quality) {
    const float4 b = __ldg(boxes + a);
    const float ba = box_area(b);
    bool restore = false;
    for (int g = 0; g < G; ++g) restore |= (pair_iou(s_gt[g], s_area[g], b, ba) == s_best[g]);
    if (restore) m = all_m;
  }
  matches[a] = m;
}

}  // namespace

extern "C" size_t mmt_match_workspace_bytes(int n_gt, int n_boxes) {
  return mmt::align_up(sizeof(float) * (size_t)n_boxes, 256) + mmt::align_up(sizeof(int) * (size_t)n_boxes, 256) +
         mmt::align_up(sizeof(unsigned int) * (size_t)(n_gt > 0 ? n_gt : 1), 256);
}

extern "C" int mmt_match_boxes_f32(const float* gt_xyxy, int n_gt, const float* boxes_xyxy, int n_boxes,
                                   float high_threshold, float low_threshold, int allow_low_quality_matches,
                                   int64_t* matches_out, void* workspace, size_t workspace_bytes, void* stream_) {
  MMT_REQUIRE(n_gt > 0 && n_boxes > 0, "match_boxes: empty ground truth or prediction set (the reference raises too)");
  MMT_REQUIRE(n_gt <= kMaxGtSmem, "match_boxes: more than %d ground-truth boxes per image", kMaxGtSmem);
  MMT_REQUIRE(gt_xyxy && boxes_xyxy && matches_out, "match_boxes: null pointer");
  MMT_REQUIRE((((uintptr_t)gt_xyxy | (uintptr_t)boxes_xyxy) & 15) == 0, "match_boxes: boxes must be 16-byte aligned");
  if (workspace == nullptr || workspace_bytes < mmt_match_workspace_bytes(n_gt, n_boxes))
    return mmt::fail(MMT_ERR_WORKSPACE_TOO_SMALL, "match_boxes: workspace too small");
  cudaStream_t stream = (cudaStream_t)stream_;
  char* w = (char*)workspace;
  float* best_val = (float*)w;
  w += mmt::align_up(sizeof(float) * (size_t)n_boxes, 256);
  int* best_idx = (int*)w;
  w += mmt::align_up(sizeof(int) * (size_t)n_boxes, 256);
  unsigned int* best_per_gt = (unsigned int*)w;
  MMT_CUDA(cudaMemsetAsync(best_per_gt, 0, sizeof(unsigned int) * (size_t)n_gt, stream));
  const int blocks = (n_boxes + 255) / 256;
  const size_t smem = (size_t)n_gt * (sizeof(float4) + 2 * sizeof(float));
  match_pass1_kernel<<<blocks, 256, smem, stream>>>(reinterpret_cast<const float4*>(gt_xyxy), n_gt,
                                                    reinterpret_cast<const float4*>(boxes_xyxy), n_boxes, best_val,
                                                    best_idx, best_per_gt);
  if (int rc = mmt::check_launch("match_pass1")) return rc;
  match_pass2_kernel<<<blocks, 256, smem, stream>>>(reinterpret_cast<const float4*>(gt_xyxy), n_gt,
                                                    reinterpret_cast<const float4*>(boxes_xyxy), n_boxes, best_val,
                                                    best_idx, best_per_gt, high_threshold, low_threshold,
                                                    allow_low_quality_matches, (long long*)matches_out);
  return mmt::check_launch("match_pass2");
}

// ------------------------------------------------------------------------------------------------------------
// BoxCoder.encode fused with the gather of the matched ground-truth box
// (reference modeling/box_coder.py:22-54: ~25 elementwise launches per call; called per image from
//  rpn/loss.py:79-82 and roi_heads/box_head/loss.py:75-78).  targets[t] = encode(gt[max(match[t], 0)], boxes[t]).
namespace {
__global__ void __launch_bounds__(256) encode_kernel(const float4* __restrict__ gt, const long long* __restrict__ match,
                                                     const float4* __restrict__ boxes, long long total, float wx,
                                                     float wy, float ww, float wh, float4* __restrict__ out) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    long long m = match ? match[i] : i;
    if (m < 0) m = 0;
    const float4 g = __ldg(gt + m), p = __ldg(boxes + i);
    const float ex_w = __fadd_rn(__fsub_rn(p.z, p.x), 1.f), ex_h = __fadd_rn(__fsub_rn(p.w, p.y), 1.f);
    const float ex_cx = __fadd_rn(p.x, __fmul_rn(0.5f, ex_w)), ex_cy = __fadd_rn(p.y, __fmul_rn(0.5f, ex_h));
    const float gt_w = __fadd_rn(__fsub_rn(g.z, g.x), 1.f), gt_h = __fadd_rn(__fsub_rn(g.w, g.y), 1.f);
    const float gt_cx = __fadd_rn(g.x, __fmul_rn(0.5f, gt_w)), gt_cy = __fadd_rn(g.y, __fmul_rn(0.5f, gt_h));
    float4 o;
    o.x = __fdiv_rn(__fmul_rn(wx, __fsub_rn(gt_cx, ex_cx)), ex_w);
    o.y = __fdiv_rn(__fmul_rn(wy, __fsub_rn(gt_cy, ex_cy)), ex_h);
    o.z = __fmul_rn(ww, logf(__fdiv_rn(gt_w, ex_w)));
    o.w = __fmul_rn(wh, logf(__fdiv_rn(gt_h, ex_h)));
    out[i] = o;
  }
}

// ROI table of the FPN pooler: rois[k] = (image index as float, x1, y1, x2, y2) and the pyramid level of
// modeling/poolers.py:31-42 (LevelMapper: floor(lvl0 + log2(sqrt(area) / s0 + eps)) clamped to [k_min, k_max], minus
// k_min; area with +1 extents) — reference poolers.py:78-89 + :31-42: ~15 launches per pooler call.
struct ImageOffsets { int n; int start[65]; };
__global__ void __launch_bounds__(256) fpn_rois_kernel(const float4* __restrict__ boxes, int total, ImageOffsets off,
                                                       float k_min, float k_max, float s0, float lvl0, float eps,
                                                       float* __restrict__ rois, int* __restrict__ levels) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  int img = 0;
  while (img + 1 < off.n && i >= off.start[img + 1]) ++img;
  const float4 b = __ldg(boxes + i);
  float* r = rois + 5ll * i;
  r[0] = (float)img; r[1] = b.x; r[2] = b.y; r[3] = b.z; r[4] = b.w;
  const float area = __fmul_rn(__fadd_rn(__fsub_rn(b.z, b.x), 1.f), __fadd_rn(__fsub_rn(b.w, b.y), 1.f));
  float lv = floorf(__fadd_rn(lvl0, log2f(__fadd_rn(__fdiv_rn(sqrtf(area), s0), eps))));
  lv = fminf(fmaxf(lv, k_min), k_max);
  levels[i] = (int)((long long)lv - (long long)k_min);
}
}  // namespace

extern "C" int mmt_encode_boxes_f32(const float* gt_xyxy, const int64_t* matched_idx, const float* boxes_xyxy,
                                    int64_t n_boxes, float wx, float wy, float ww, float wh, float* targets_out,
                                    void* stream_) {
  MMT_REQUIRE(n_boxes >= 0, "encode_boxes: negative size");
  if (n_boxes == 0) return MMT_OK;
  MMT_REQUIRE(gt_xyxy && boxes_xyxy && targets_out, "encode_boxes: null pointer");
  MMT_REQUIRE((((uintptr_t)gt_xyxy | (uintptr_t)boxes_xyxy | (uintptr_t)targets_out) & 15) == 0,
              "encode_boxes: box arrays must be 16-byte aligned");
  long long blocks = (n_boxes + 255) / 256;
  if (blocks > mmt::kNumSMs * 8) blocks = mmt::kNumSMs * 8;
  encode_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream_>>>(
      reinterpret_cast<const float4*>(gt_xyxy), (const long long*)matched_idx,
      reinterpret_cast<const float4*>(boxes_xyxy), n_boxes, wx, wy, ww, wh, reinterpret_cast<float4*>(targets_out));
  return mmt::check_launch("encode_boxes");
}

extern "C" int mmt_fpn_rois_f32(const float* boxes_xyxy, const int* image_counts_host, int n_images, float k_min,
                                float k_max, float canonical_scale, float canonical_level, float eps, float* rois_out,
                                int* levels_out, void* stream_) {
  MMT_REQUIRE(n_images >= 0 && n_images <= 64, "fpn_rois: at most 64 images per call (got %d)", n_images);
  ImageOffsets off;
  off.n = n_images;
  int total = 0;
  for (int i = 0; i < n_images; ++i) {
    MMT_REQUIRE(image_counts_host[i] >= 0, "fpn_rois: negative count");
    off.start[i] = total;
    total += image_counts_host[i];
  }
  off.start[n_images] = total;
  if (total == 0) return MMT_OK;
  MMT_REQUIRE(boxes_xyxy && rois_out && levels_out, "fpn_rois: null pointer");
  MMT_REQUIRE(((uintptr_t)boxes_xyxy & 15) == 0, "fpn_rois: boxes must be 16-byte aligned");
  fpn_rois_kernel<<<(total + 255) / 256, 256, 0, (cudaStream_t)stream_>>>(
      reinterpret_cast<const float4*>(boxes_xyxy), total, off, k_min, k_max, canonical_scale, canonical_level, eps,
      rois_out, levels_out);
  return mmt::check_launch("fpn_rois");
}

// ------------------------------------------------------------------------------------------------------------
// Balanced positive / negative sampling (reference modeling/balanced_positive_negative_sampler.py:16-71): per image
// a uniformly random subset of min(#pos, quota_pos) positives (label >= 1) and min(#neg, batch - #pos_taken)
// negatives (label == 0).  The reference draws torch.randperm(count)[:n] — which needs the counts on the host (a sync
// per image) and ~30 launches.  Here: every candidate gets a counter-based random 32-bit key (hash of seed, image,
// index); the n smallest keys of a class are its sample, found by a radix select on the key value (4 passes of 8 bits
// over a prefiltered key list in shared memory, one CTA per image), ties resolved by ticket.  Same distribution
// (every subset of the required size is equally likely), no host round trip, one launch for the whole batch.
namespace {

// 32-bit counter-based key: two rounds of the murmur3 finaliser over (seed, image, index) — ~12 integer instructions
// (the sampler evaluates it for every anchor in each of its ~5 passes, so a 64-bit mixer would dominate the kernel)
__device__ __forceinline__ unsigned int fmix32(unsigned int h) {
  h ^= h >> 16; h *= 0x85EBCA6Bu;
  h ^= h >> 13; h *= 0xC2B2AE35u;
  h ^= h >> 16;
  return h;
}
__device__ __forceinline__ unsigned int hash_key(unsigned long long seed, unsigned int img, unsigned int i) {
  const unsigned int s0 = (unsigned int)seed, s1 = (unsigned int)(seed >> 32);
  return fmix32(fmix32(i * 0x9E3779B1u + s0) ^ (img * 0x7FEB352Du + s1));
}

__device__ int block_count(int v, int* s_part) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  v = mmt::warp_sum_int(v);
  __syncthreads();
  if (lane == 0) s_part[wid] = v;
  __syncthreads();
  int r = (lane < (int)(blockDim.x >> 5)) ? s_part[lane] : 0;
  r = mmt::warp_sum_int(r);
  return r;  // valid in every thread of warp 0..; broadcast below
}

constexpr int kSampleCap = 8192;   // candidate keys held in shared memory during the threshold search

// smallest threshold T such that #{i in class : key_i <= T} >= want  (1 <= want < class size).
// The class can be ~150 K anchors while want <= 512: one pass first collects the keys below a prefilter threshold
// chosen to catch ~4 x want of them (uniform keys) into shared memory, and the 32-step bitwise search runs over that
// list; only if the prefilter caught too few / too many (probability ~0) does the search stream the whole image.
__device__ unsigned int select_threshold(const float* __restrict__ labels, int n, int cls_pos, unsigned long long seed,
                                         unsigned int img, int want, int cls_size, int* s_part, int* s_bcast,
                                         unsigned int* s_keys, int* s_hist, int* need_at_threshold) {
  // a class that fits in the list is collected whole (T0 = max key); a larger one is prefiltered to ~4 x want keys
  const unsigned long long target = (unsigned long long)want * 4ull + 64ull;
  const bool whole = cls_size <= kSampleCap;
  const bool prefilter = whole || target < (unsigned long long)cls_size;
  unsigned int T0 = 0xffffffffu;
  int m = 0;
  if (prefilter) {
    if (!whole) T0 = (unsigned int)((target << 32) / (unsigned long long)cls_size);
    if (threadIdx.x == 0) *s_bcast = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31;
    for (int base0 = 0; base0 < n; base0 += blockDim.x) {      // uniform trip count: the ballots below need whole warps
      const int i = base0 + threadIdx.x;
      bool hit = false;
      unsigned int k = 0u;
      if (i < n) {
        const float l = labels[i];
        if (cls_pos ? (l >= 1.f) : (l == 0.f)) {
          k = hash_key(seed, img, (unsigned int)i);
          hit = k <= T0;
        }
      }
      const unsigned int mask = __ballot_sync(0xffffffffu, hit);   // one slot reservation per warp, not per candidate
      if (mask) {
        const int leader = __ffs(mask) - 1;
        int slot0 = 0;
        if (lane == leader) slot0 = atomicAdd(s_bcast, __popc(mask));
        slot0 = __shfl_sync(0xffffffffu, slot0, leader);
        if (hit) {
          const int slot = slot0 + __popc(mask & ((1u << lane) - 1u));
          if (slot < kSampleCap) s_keys[slot] = k;
        }
      }
    }
    __syncthreads();
    m = *s_bcast;
    __syncthreads();
  }
  const bool use_list = prefilter && m >= want && m <= kSampleCap;
  // T = the want-th smallest key: MSB-first radix select, 4 passes of 8 bits with a 256-bin shared-memory histogram
  // (uniform keys: no bin is hot) — 4 block-wide steps instead of the 32 of a bit-by-bit search
  unsigned int prefix = 0u;
  int need = want;
  for (int pass = 0; pass < 4; ++pass) {
    const int shift = 24 - 8 * pass;
    const unsigned int hi_mask = pass == 0 ? 0u : (0xffffffffu << (shift + 8));
    for (int i = threadIdx.x; i < 256; i += blockDim.x) s_hist[i] = 0;
    __syncthreads();
    if (use_list) {
      for (int i = threadIdx.x; i < m; i += blockDim.x) {
        const unsigned int k = s_keys[i];
        if ((k & hi_mask) == prefix) atomicAdd(&s_hist[(k >> shift) & 255u], 1);
      }
    } else {
      for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const float l = labels[i];
        const bool in_cls = cls_pos ? (l >= 1.f) : (l == 0.f);
        if (in_cls) {
          const unsigned int k = hash_key(seed, img, (unsigned int)i);
          if ((k & hi_mask) == prefix) atomicAdd(&s_hist[(k >> shift) & 255u], 1);
        }
      }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      int d = 0, below = 0;
      while (d < 255 && below + s_hist[d] < need) { below += s_hist[d]; ++d; }
      *s_bcast = d;
      s_part[0] = need - below;
    }
    __syncthreads();
    prefix |= (unsigned int)(*s_bcast) << shift;
    need = s_part[0];
    __syncthreads();
  }
  *need_at_threshold = need;   // how many of the keys EQUAL to the threshold are still wanted (the rest, want - need, lie below it)
  return prefix;
}

__global__ void __launch_bounds__(1024) balanced_sample_kernel(const float* __restrict__ labels,
                                                               const int* __restrict__ offsets, int batch_per_image,
                                                               int quota_pos, unsigned long long seed,
                                                               unsigned char* __restrict__ pos_mask,
                                                               unsigned char* __restrict__ neg_mask) {
  __shared__ int s_part[32];
  __shared__ int s_bcast;
  __shared__ unsigned int s_keys[kSampleCap];
  __shared__ int s_hist[256];
  const unsigned int img = blockIdx.x;
  const int base = offsets[img], n = offsets[img + 1] - base;
  const float* lab = labels + base;
  // class sizes
  int cp = 0, cn = 0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const float l = lab[i];
    cp += (l >= 1.f);
    cn += (l == 0.f);
  }
  cp = block_count(cp, s_part);
  if (threadIdx.x == 0) s_bcast = cp;
  __syncthreads();
  cp = s_bcast;
  __syncthreads();
  cn = block_count(cn, s_part);
  if (threadIdx.x == 0) s_bcast = cn;
  __syncthreads();
  cn = s_bcast;
  __syncthreads();
  const int num_pos = min(cp, quota_pos);
  const int num_neg = min(cn, batch_per_image - num_pos);
  const bool all_pos = num_pos == cp, all_neg = num_neg == cn;
  unsigned int tp = 0xffffffffu, tn = 0xffffffffu;
  int need_p = 0, need_n = 0;
  if (!all_pos && num_pos > 0) tp = select_threshold(lab, n, 1, seed, img, num_pos, cp, s_part, &s_bcast, s_keys, s_hist, &need_p);
  if (!all_neg && num_neg > 0) tn = select_threshold(lab, n, 0, seed, img, num_neg, cn, s_part, &s_bcast, s_keys, s_hist, &need_n);
  // keys strictly below the threshold are in; of the keys equal to it (the threshold element itself, plus ties — a 2^-32
  // event per pair) the select wants need_p / need_n more: admitted by ticket
  __shared__ int s_tie_p, s_tie_n;
  if (threadIdx.x == 0) { s_tie_p = 0; s_tie_n = 0; }
  __syncthreads();
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const float l = lab[i];
    const unsigned int k = hash_key(seed, img, (unsigned int)i);
    bool p = (l >= 1.f) && num_pos > 0 && (all_pos || k < tp);
    bool q = (l == 0.f) && num_neg > 0 && (all_neg || k < tn);
    if (!p && (l >= 1.f) && !all_pos && num_pos > 0 && k == tp) p = atomicAdd(&s_tie_p, 1) < need_p;
    if (!q && (l == 0.f) && !all_neg && num_neg > 0 && k == tn) q = atomicAdd(&s_tie_n, 1) < need_n;
    pos_mask[base + i] = p;
    neg_mask[base + i] = q;
  }
}

}  // namespace

extern "C" int mmt_balanced_sample_f32(const float* labels, const int* image_offsets_dev, int n_images,
                                       int batch_size_per_image, int quota_pos, unsigned long long seed,
                                       unsigned char* pos_mask_out, unsigned char* neg_mask_out, void* stream_) {
  MMT_REQUIRE(n_images >= 0 && batch_size_per_image >= 0 && quota_pos >= 0, "balanced_sample: negative argument");
  if (n_images == 0) return MMT_OK;
  MMT_REQUIRE(labels && image_offsets_dev && pos_mask_out && neg_mask_out, "balanced_sample: null pointer");
  balanced_sample_kernel<<<n_images, 1024, 0, (cudaStream_t)stream_>>>(labels, image_offsets_dev, batch_size_per_image,
                                                                       quota_pos, seed, pos_mask_out, neg_mask_out);
  return mmt::check_launch("balanced_sample");
}

// ------------------------------------------------------------------------------------------------------------
// RPN candidate gather: for the top-k anchor indices of one pyramid level, fetch the box deltas straight from the
// fused head map [N, HW, Cp] (channels [A, 5A) hold the 4A deltas), the anchor boxes and the (sorted) scores, and
// write them into the pass-wide candidate tables at this level's column offset.  Replaces, per level, the reference's
// permute/reshape copy of the delta map, two advanced-index gathers and three concatenations
// (modeling/rpn/inference.py:86-105).
namespace {
__global__ void __launch_bounds__(256) rpn_gather_kernel(const float* __restrict__ fused, int Cp, int A, long long HW,
                                                         const long long* __restrict__ idx,
                                                         const float* __restrict__ vals,
                                                         const float4* __restrict__ anchors, int N, int k,
                                                         int K_total, int col_off, float4* __restrict__ deltas,
                                                         float4* __restrict__ anch, float* __restrict__ scores) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= N * k) return;
  const int n = t / k, j = t - n * k;
  const long long i = idx[t];
  const long long pix = i / A;
  const int a = (int)(i - pix * A);
  const float* src = fused + ((long long)n * HW + pix) * Cp + A + 4 * a;
  const long long o = (long long)n * K_total + col_off + j;
  deltas[o] = make_float4(__ldg(src), __ldg(src + 1), __ldg(src + 2), __ldg(src + 3));
  anch[o] = __ldg(anchors + i);
  scores[o] = vals[t];
}
}  // namespace

extern "C" int mmt_rpn_gather_f32(const float* fused_nhwc, int padded_channels, int num_anchors, int64_t pixels,
                                  const int64_t* topk_idx, const float* topk_vals, const float* level_anchors,
                                  int n_images, int k, int k_total, int col_offset, float* deltas_out,
                                  float* anchors_out, float* scores_out, void* stream_) {
  MMT_REQUIRE(n_images >= 0 && k >= 0 && col_offset >= 0 && col_offset + k <= k_total, "rpn_gather: bad sizes");
  if (n_images == 0 || k == 0) return MMT_OK;
  MMT_REQUIRE(fused_nhwc && topk_idx && topk_vals && level_anchors && deltas_out && anchors_out && scores_out,
              "rpn_gather: null pointer");
  MMT_REQUIRE(padded_channels >= 5 * num_anchors && num_anchors > 0, "rpn_gather: head map narrower than 5 x anchors");
  MMT_REQUIRE((((uintptr_t)level_anchors | (uintptr_t)deltas_out | (uintptr_t)anchors_out) & 15) == 0,
              "rpn_gather: box arrays must be 16-byte aligned");
  const int total = n_images * k;
  rpn_gather_kernel<<<(total + 255) / 256, 256, 0, (cudaStream_t)stream_>>>(
      fused_nhwc, padded_channels, num_anchors, pixels, (const long long*)topk_idx, topk_vals,
      reinterpret_cast<const float4*>(level_anchors), n_images, k, k_total, col_offset,
      reinterpret_cast<float4*>(deltas_out), reinterpret_cast<float4*>(anchors_out), scores_out);
  return mmt::check_launch("rpn_gather");
}


// ------------------------------------------------------------------------------------------------
// RPN candidate selection: per (image, level) the pre_nms_top_n highest objectness scores in descending order
// (reference modeling/rpn/inference.py:92-94: `objectness.topk(pre_nms_top_n, dim=1, sorted=True)`) FUSED with the gather
// above — ONE launch for every (image, level) problem of a pass instead of, per level, torch's multi-block radix top-k
// (5-6 kernels, ~55 us) + a sort + the gather.  One 1024-thread CTA per problem:
//   * rows of <= 2048 anchors (the two coarsest levels) are sorted whole;
//   * otherwise an MSB-first radix select (4 passes of 8 bits over the row, per-warp shared-memory histograms) finds the
//     bit pattern T of the k-th largest score exactly, a compaction pass collects the scores > T plus the lowest-index
//     scores == T, and a bitonic sort of (score bits, ~index) orders the <= 2048 survivors: descending score, ascending
//     anchor index among equal scores (torch.topk leaves that order unspecified).
namespace {
constexpr int kTopkThreads = 1024;
constexpr int kTopkCap = 2048;
constexpr int kTopkUnroll = 8;
constexpr int kTopkLevels = 8;

struct TopkTab {
  const float* scores[kTopkLevels];   // [N, n]
  const float* fused[kTopkLevels];    // [N, n / A, Cp]
  const float4* anchors[kTopkLevels]; // [n]
  int n[kTopkLevels], k[kTopkLevels], col[kTopkLevels];
};

__device__ __forceinline__ uint32_t ordered_key(float x) {   // monotone: a < b  <=>  key(a) < key(b)
  const uint32_t u = __float_as_uint(x);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float key_to_float(uint32_t key) {
  return __uint_as_float((key & 0x80000000u) ? (key & 0x7fffffffu) : ~key);
}

__global__ void __launch_bounds__(kTopkThreads) rpn_topk_gather_kernel(TopkTab tab, int n_levels, int N, int A, int Cp,
                                                                      int K_total, float4* __restrict__ deltas,
                                                                      float4* __restrict__ anch, float* __restrict__ scores_out) {
  __shared__ int hist[16][256];   // one histogram per pair of warps
  __shared__ unsigned long long buf[kTopkCap];
  __shared__ int s_total[256];
  __shared__ uint32_t s_prefix;
  __shared__ int s_want, s_cnt, s_tie_base, s_ties;
  const int l = blockIdx.x % n_levels, img = blockIdx.x / n_levels;
  const int n = tab.n[l], k = tab.k[l];
  const float* __restrict__ s = tab.scores[l] + (long long)img * n;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  for (int i = tid; i < kTopkCap; i += kTopkThreads) buf[i] = 0ull;   // 0 sorts behind every real entry (keys are >= 1)
  if (tid == 0) { s_prefix = 0u; s_want = k; s_cnt = 0; s_tie_base = 0; s_ties = 0; }
  __syncthreads();

  if (n <= kTopkCap) {
    for (int i = tid; i < n; i += kTopkThreads)
      buf[i] = ((unsigned long long)ordered_key(__ldg(s + i)) << 32) | (uint32_t)(~(uint32_t)i);
  } else {
    // ---- radix select: after pass p the top 8 * (p + 1) bits of the k-th largest key are known
    for (int pass = 0; pass < 4; ++pass) {
      const int shift = 24 - 8 * pass;
      for (int i = tid; i < 16 * 256; i += kTopkThreads) (&hist[0][0])[i] = 0;
      __syncthreads();
      const uint32_t prefix = s_prefix;
      const uint32_t hi_mask = pass == 0 ? 0u : (0xffffffffu << (shift + 8));
      // Pass 0 looks at the sign and the top 7 exponent bits: sigmoid scores share a handful of exponents, so almost every
      // element of a warp lands in the same 3-4 bins — the lanes with equal digits elect one of them to add their count
      // (__match_any_sync costs one step per DISTINCT value in the warp: cheap here, expensive in the later passes, whose
      // digits are spread over all 256 bins and whose plain shared-memory atomics do not collide).
      // (kTopkUnroll independent loads in flight per thread: with one, every iteration of this single-CTA loop pays a
      // full L2 round trip — 36 us per pass over 120 000 scores)
      for (int base = 0; base < n; base += kTopkThreads * kTopkUnroll) {
        uint32_t keys[kTopkUnroll];
#pragma unroll
        for (int u = 0; u < kTopkUnroll; ++u) {
          const int i = base + u * kTopkThreads + tid;
          keys[u] = i < n ? ordered_key(__ldg(s + i)) : 0u;      // key 0 is below every real score's key
        }
#pragma unroll
        for (int u = 0; u < kTopkUnroll; ++u) {
          const int i = base + u * kTopkThreads + tid;
          uint32_t digit = 0x100u;                         // "not counted"
          if (i < n && (keys[u] & hi_mask) == prefix) digit = (keys[u] >> shift) & 255u;
          if (pass == 0) {
            const unsigned peers = __match_any_sync(0xffffffffu, digit);
            if (digit < 256u && lane == __ffs(peers) - 1) atomicAdd(&hist[warp >> 1][digit], __popc(peers));
          } else if (digit < 256u) {
            atomicAdd(&hist[warp >> 1][digit], 1);
          }
        }
      }
      __syncthreads();
      if (tid < 256) {
        int t = 0;
#pragma unroll 8
        for (int w = 0; w < 16; ++w) t += hist[w][tid];
        s_total[tid] = t;
      }
      __syncthreads();
      if (tid == 0) {
        int want = s_want, d = 255, above = 0;
        while (d > 0 && above + s_total[d] < want) { above += s_total[d]; --d; }
        s_want = want - above;                       // how many of the elements with this digit are still needed
        s_prefix = prefix | ((uint32_t)d << shift);
        if (pass == 3) s_ties = s_total[d];          // elements whose key equals the threshold exactly
      }
      __syncthreads();
    }
    const uint32_t T = s_prefix;
    const int want = s_want, ties = s_ties;          // want >= 1 of the `ties` elements equal to T
    // ---- compaction: everything above T (count = k - want), then `want` of the ties — the lowest indices
    if (ties == want) {                              // common case: take them all, order restored by the sort
      for (int base = 0; base < n; base += kTopkThreads * kTopkUnroll) {
        uint32_t keys[kTopkUnroll];
#pragma unroll
        for (int u = 0; u < kTopkUnroll; ++u) {
          const int i = base + u * kTopkThreads + tid;
          keys[u] = i < n ? ordered_key(__ldg(s + i)) : 0u;
        }
#pragma unroll
        for (int u = 0; u < kTopkUnroll; ++u) {
          const int i = base + u * kTopkThreads + tid;
          if (i < n && keys[u] >= T) buf[atomicAdd(&s_cnt, 1)] = ((unsigned long long)keys[u] << 32) | (uint32_t)(~(uint32_t)i);
        }
      }
    } else {
      __shared__ int warp_ties[32];
      for (int base = 0; base < n; base += kTopkThreads) {
        const int i = base + tid;
        const uint32_t key = i < n ? ordered_key(__ldg(s + i)) : 0u;
        const bool gt = i < n && key > T, eq = i < n && key == T;
        if (gt) buf[atomicAdd(&s_cnt, 1)] = ((unsigned long long)key << 32) | (uint32_t)(~(uint32_t)i);
        const unsigned m = __ballot_sync(0xffffffffu, eq);
        if (lane == 0) warp_ties[warp] = __popc(m);
        __syncthreads();
        int before = s_tie_base;
        for (int w = 0; w < warp; ++w) before += warp_ties[w];
        const int rank = before + __popc(m & ((1u << lane) - 1u));
        if (eq && rank < want) buf[(k - want) + rank] = ((unsigned long long)key << 32) | (uint32_t)(~(uint32_t)i);
        __syncthreads();
        if (tid == 0) {
          int t = s_tie_base;
          for (int w = 0; w < 32; ++w) t += warp_ties[w];
          s_tie_base = t;
        }
        __syncthreads();
      }
    }
  }
  __syncthreads();
  // ---- bitonic sort, descending, of the kTopkCap 64-bit entries (empty slots are 0 and end up last)
  for (int size = 2; size <= kTopkCap; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      const int pos = 2 * tid - (tid & (stride - 1));
      const bool desc = (pos & size) == 0;
      const unsigned long long a = buf[pos], b = buf[pos + stride];
      if ((a < b) == desc) { buf[pos] = b; buf[pos + stride] = a; }
      __syncthreads();
    }
  }
  // ---- gather: deltas from the fused head map, anchors, scores -> the pass-wide tables at this level's columns
  const int HW = n / A;
  const float* __restrict__ f = tab.fused[l] + (long long)img * HW * Cp;
  for (int j = tid; j < k; j += kTopkThreads) {
    const unsigned long long e = buf[j];
    const uint32_t i = ~(uint32_t)(e & 0xffffffffull);
    const int pix = (int)(i / (uint32_t)A), a = (int)(i - (uint32_t)pix * A);
    const float* src = f + (long long)pix * Cp + A + 4 * a;
    const long long o = (long long)img * K_total + tab.col[l] + j;
    deltas[o] = make_float4(__ldg(src), __ldg(src + 1), __ldg(src + 2), __ldg(src + 3));
    anch[o] = __ldg(tab.anchors[l] + i);
    scores_out[o] = key_to_float((uint32_t)(e >> 32));
  }
}
}  // namespace

extern "C" int mmt_rpn_topk_gather_f32(const float* const* level_scores_host, const float* const* level_fused_host,
                                       const float* const* level_anchors_host, const int* level_n_host,
                                       const int* level_k_host, int n_levels, int n_images, int num_anchors,
                                       int padded_channels, int k_total, float* deltas_out, float* anchors_out,
                                       float* scores_out, void* stream_) {
  MMT_REQUIRE(n_levels >= 1 && n_levels <= kTopkLevels && n_images >= 0 && num_anchors > 0, "rpn_topk_gather: bad sizes");
  if (n_images == 0) return MMT_OK;
  MMT_REQUIRE(level_scores_host && level_fused_host && level_anchors_host && level_n_host && level_k_host && deltas_out &&
                  anchors_out && scores_out, "rpn_topk_gather: null pointer");
  MMT_REQUIRE(padded_channels >= 5 * num_anchors, "rpn_topk_gather: head map narrower than 5 x anchors");
  MMT_REQUIRE((((uintptr_t)deltas_out | (uintptr_t)anchors_out) & 15) == 0, "rpn_topk_gather: box arrays must be 16-byte aligned");
  TopkTab tab;
  int col = 0;
  for (int l = 0; l < n_levels; ++l) {
    MMT_REQUIRE(level_scores_host[l] && level_fused_host[l] && level_anchors_host[l], "rpn_topk_gather: null level pointer");
    MMT_REQUIRE(((uintptr_t)level_anchors_host[l] & 15) == 0, "rpn_topk_gather: anchors must be 16-byte aligned");
    MMT_REQUIRE(level_n_host[l] > 0 && level_n_host[l] % num_anchors == 0 && level_k_host[l] > 0 &&
                    level_k_host[l] <= level_n_host[l], "rpn_topk_gather: bad level size");
    if (level_k_host[l] > kTopkCap)
      return mmt::fail(MMT_ERR_UNSUPPORTED, "rpn_topk_gather: k = %d exceeds the in-kernel sort capacity %d", level_k_host[l], kTopkCap);
    tab.scores[l] = level_scores_host[l]; tab.fused[l] = level_fused_host[l];
    tab.anchors[l] = reinterpret_cast<const float4*>(level_anchors_host[l]);
    tab.n[l] = level_n_host[l]; tab.k[l] = level_k_host[l]; tab.col[l] = col;
    col += level_k_host[l];
  }
  MMT_REQUIRE(col == k_total, "rpn_topk_gather: k_total %d != sum of the level k (%d)", k_total, col);
  rpn_topk_gather_kernel<<<n_images * n_levels, kTopkThreads, 0, (cudaStream_t)stream_>>>(
      tab, n_levels, n_images, num_anchors, padded_channels, k_total, reinterpret_cast<float4*>(deltas_out),
      reinterpret_cast<float4*>(anchors_out), scores_out);
  return mmt::check_launch("rpn_topk_gather");
}
