// Segmented, fully on-device greedy NMS for sm_100a.
//
// Replaces maskrcnn_benchmark/csrc/cuda/nms.cu:23-131 (mask kernel + D2H copy + host scan) and the
// tensor plumbing of csrc/nms.h:10-28.  Results are bit-identical to the reference's CUDA flavour:
// "+1" areas and IoU in fp32 with the reference's operation order (nms.cu:13-21), suppression on
// IoU > thr (nms.cu:60), visiting order = descending score (nms.cu:74-75; ties -> lower original
// index, the reference leaves ties unspecified), output = kept original indices ascending
// (nms.cu:127-130).
//
// Pipeline (3 launches for ANY number of segments, no host synchronisation, no allocation):
//   1. rank_sort_kernel   : rank of every box inside its segment by counting (O(n^2) compares,
//                           the same order of work as the IoU matrix itself; no smem size limit)
//   2. iou_mask_kernel    : 64x64 tiles of the upper-triangular suppression bit-matrix
//   3. greedy_scan_kernel : one CTA per segment walks the 64-box blocks in order; the 64-step
//                           serial dependency runs out of shared memory, the row OR-reduction is
//                           spread over 32 warps; kept flags are compacted in index order.
#include "common.cuh"

namespace {

constexpr int kMaxSegPerLaunch = 48;
constexpr int kTile = 64;

struct SegTable {
  int n_seg;
  int box_off[kMaxSegPerLaunch + 1];       // box offsets (relative to the arrays passed in)
  long long mask_off[kMaxSegPerLaunch + 1];  // word offsets into the mask buffer
};

__device__ __forceinline__ float iou_plus_one(const float4 a, const float4 b) {
  // same expression order as reference nms.cu:13-21 (no FMA-contractable pattern)
  float left = fmaxf(a.x, b.x), right = fminf(a.z, b.z);
  float top = fmaxf(a.y, b.y), bottom = fminf(a.w, b.w);
  float width = fmaxf(right - left + 1.f, 0.f), height = fmaxf(bottom - top + 1.f, 0.f);
  float interS = __fmul_rn(width, height);
  float Sa = __fmul_rn(a.z - a.x + 1.f, a.w - a.y + 1.f);
  float Sb = __fmul_rn(b.z - b.x + 1.f, b.w - b.y + 1.f);
  return __fdiv_rn(interS, Sa + Sb - interS);
}

__global__ void __launch_bounds__(256) rank_sort_kernel(const float* __restrict__ boxes,
                                                        const float* __restrict__ scores,
                                                        SegTable tab, int* __restrict__ order,
                                                        float4* __restrict__ sorted_boxes) {
  const int s = blockIdx.y;
  const int base = tab.box_off[s];
  const int n = tab.box_off[s + 1] - base;
  if ((int)(blockIdx.x * blockDim.x) >= n) return;
  __shared__ float tile[256];
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  // NaN scores (a diverged network) are ranked last instead of corrupting the permutation
  float si = (i < n) ? scores[base + i] : 0.f;
  if (si != si) si = -INFINITY;
  int rank = 0;
  for (int j0 = 0; j0 < n; j0 += 256) {
    const int j = j0 + threadIdx.x;
    float sj_in = (j < n) ? scores[base + j] : 0.f;
    if (sj_in != sj_in) sj_in = -INFINITY;
    tile[threadIdx.x] = sj_in;
    __syncthreads();
    const int lim = min(256, n - j0);
    if (i < n) {
#pragma unroll 8
      for (int t = 0; t < lim; ++t) {
        const float sj = tile[t];
        const int jj = j0 + t;
        rank += (sj > si) || (sj == si && jj < i);
      }
    }
    __syncthreads();
  }
  if (i < n) {
    order[base + rank] = i;
    const float4 b = reinterpret_cast<const float4*>(boxes)[base + i];
    sorted_boxes[base + rank] = b;
  }
}

__global__ void __launch_bounds__(kTile) iou_mask_kernel(const float4* __restrict__ sorted_boxes,
                                                         SegTable tab, float thr,
                                                         unsigned long long* __restrict__ mask) {
  const int s = blockIdx.z;
  const int base = tab.box_off[s];
  const int n = tab.box_off[s + 1] - base;
  const int row_blk = blockIdx.y, col_blk = blockIdx.x;
  if (col_blk < row_blk) return;  // only the upper triangle is ever read by the scan
  if (row_blk * kTile >= n || col_blk * kTile >= n) return;
  const int col_blocks = (n + kTile - 1) / kTile;
  const int row_size = min(n - row_blk * kTile, kTile);
  const int col_size = min(n - col_blk * kTile, kTile);
  __shared__ float4 cols[kTile];
  if ((int)threadIdx.x < col_size) cols[threadIdx.x] = sorted_boxes[base + col_blk * kTile + threadIdx.x];
  __syncthreads();
  if ((int)threadIdx.x < row_size) {
    const int r = row_blk * kTile + threadIdx.x;
    const float4 me = sorted_boxes[base + r];
    unsigned long long bits = 0ull;
    const int start = (row_blk == col_blk) ? threadIdx.x + 1 : 0;
    for (int c = start; c < col_size; ++c)
      if (iou_plus_one(me, cols[c]) > thr) bits |= 1ull << c;
    mask[tab.mask_off[s] + (long long)r * col_blocks + col_blk] = bits;
  }
}

// One CTA (1024 threads) per segment.
__global__ void __launch_bounds__(1024) greedy_scan_kernel(const unsigned long long* __restrict__ mask,
                                                           const int* __restrict__ order,
                                                           SegTable tab,
                                                           unsigned char* __restrict__ keep_flag,
                                                           long long* __restrict__ keep_out,
                                                           int* __restrict__ n_keep_out, int seg_base_index) {
  extern __shared__ unsigned long long smem_u64[];
  const int s = blockIdx.x;
  const int base = tab.box_off[s];
  const int n = tab.box_off[s + 1] - base;
  const int cb = (n + kTile - 1) / kTile;
  unsigned long long* remv = smem_u64;       // [cb]
  unsigned long long* diag = smem_u64 + cb;  // [64]
  __shared__ unsigned long long s_kept;
  __shared__ int s_scan[32];
  __shared__ int s_running;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const unsigned long long* m = mask + tab.mask_off[s];

  for (int j = tid; j < cb; j += blockDim.x) remv[j] = 0ull;
  for (int i = tid; i < n; i += blockDim.x) keep_flag[base + i] = 0;
  __syncthreads();

  // Per 64-box block: (1) the block's diagonal 64x64 bit tile, (2) a 64-step serial pass over it by one thread, (3) the rows
  // of the boxes that pass kept are OR-ed into the suppression words of the blocks to the right.  (1) and (3) read the mask
  // from global memory; neither address depends on the outcome of (2), so both are issued BEFORE it — the next block's
  // diagonal one iteration ahead, this block's row words (warp w owns rows w and w + 32, one column word per lane) at the
  // top of the iteration — and only their use waits for `kept`.  The serial step then hides the memory latency instead of
  // queueing behind it (r01: ~2.6 us per block, 85 us per 2000-box segment).
  unsigned long long diag_next = 0ull;
  if (tid < kTile) diag_next = (tid < min(kTile, n)) ? m[(long long)tid * cb + 0] : 0ull;
  for (int b = 0; b < cb; ++b) {
    const int rows = min(kTile, n - b * kTile);
    if (tid < kTile) diag[tid] = diag_next;
    // row words of this block for the first 32 columns to the right (all of them when n <= 2112), loaded speculatively
    const int j0 = b + 1 + lane;
    unsigned long long w0 = 0ull, w1 = 0ull;
    if (j0 < cb) {
      if (wid < rows) w0 = m[(long long)(b * kTile + wid) * cb + j0];
      if (wid + 32 < rows) w1 = m[(long long)(b * kTile + wid + 32) * cb + j0];
    }
    if (tid < kTile && b + 1 < cb) {
      const int rows_n = min(kTile, n - (b + 1) * kTile);
      diag_next = (tid < rows_n) ? m[(long long)((b + 1) * kTile + tid) * cb + (b + 1)] : 0ull;
    }
    __syncthreads();
    if (tid == 0) {
      unsigned long long cur = remv[b], kept = 0ull;
      for (int t = 0; t < rows; ++t) {
        if (!((cur >> t) & 1ull)) {
          kept |= 1ull << t;
          cur |= diag[t];
        }
      }
      s_kept = kept;
    }
    __syncthreads();
    const unsigned long long kept = s_kept;
    if (tid < rows && ((kept >> tid) & 1ull)) keep_flag[base + order[base + b * kTile + tid]] = 1;
    {
      unsigned long long acc = 0ull;
      if ((kept >> wid) & 1ull) acc |= w0;
      if ((kept >> (wid + 32)) & 1ull) acc |= w1;
      if (acc) atomicOr(&remv[j0], acc);
    }
    for (int j = j0 + 32; j < cb; j += 32) {   // segments longer than 2112 boxes: the remaining columns, loaded on demand
      unsigned long long acc = 0ull;
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int t = wid + 32 * h;
        if (t < rows && ((kept >> t) & 1ull)) acc |= m[(long long)(b * kTile + t) * cb + j];
      }
      if (acc) atomicOr(&remv[j], acc);
    }
    __syncthreads();
  }

  // compact keep flags in index order -> ascending original indices
  if (tid == 0) s_running = 0;
  __syncthreads();
  for (int i0 = 0; i0 < n; i0 += blockDim.x) {
    const int i = i0 + tid;
    const int f = (i < n) ? keep_flag[base + i] : 0;
    const unsigned ballot = __ballot_sync(0xffffffffu, f);
    const int before = __popc(ballot & ((1u << lane) - 1u));
    if (lane == 0) s_scan[wid] = __popc(ballot);
    __syncthreads();
    if (wid == 0) {
      int v = s_scan[lane];
      int incl = v;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        int u = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += u;
      }
      s_scan[lane] = incl - v;  // exclusive
      if (lane == 31) diag[0] = (unsigned long long)incl;
    }
    __syncthreads();
    const int run = s_running;
    if (f) keep_out[base + run + s_scan[wid] + before] = i;
    __syncthreads();
    if (tid == 0) s_running = run + (int)diag[0];
    __syncthreads();
  }
  if (tid == 0) n_keep_out[seg_base_index + s] = s_running;
}

struct Workspace {
  int* order;
  float4* sorted;
  unsigned char* flags;
  unsigned long long* mask;
};

size_t mask_words(const int* off, int n_seg) {
  size_t w = 0;
  for (int s = 0; s < n_seg; ++s) {
    size_t n = (size_t)(off[s + 1] - off[s]);
    w += n * ((n + kTile - 1) / kTile);
  }
  return w;
}

}  // namespace

extern "C" size_t mmt_nms_workspace_bytes(int total_boxes, int n_segments, int max_segment_boxes) {
  if (total_boxes <= 0) return 256;
  // upper bound of the bit-matrix: every segment at most max_segment_boxes wide
  size_t cb = ((size_t)max_segment_boxes + kTile - 1) / kTile;
  size_t words = (size_t)total_boxes * cb;
  (void)n_segments;
  return mmt::align_up((size_t)total_boxes * sizeof(int), 256) +
         mmt::align_up((size_t)total_boxes * sizeof(float4), 256) +
         mmt::align_up((size_t)total_boxes, 256) + mmt::align_up(words * 8, 256) + 256;
}

static int nms_segmented_impl(const float* boxes, const float* scores, const int* seg_off, int n_segments, float thr,
                              int64_t* keep_out, int* n_keep_out, unsigned char* flags_ext, void* workspace,
                              size_t workspace_bytes, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  MMT_REQUIRE(n_segments >= 0 && seg_off != nullptr, "nms: bad segment table");
  if (n_segments == 0) return MMT_OK;
  const int total = seg_off[n_segments] - seg_off[0];
  MMT_REQUIRE(seg_off[0] == 0, "nms: seg_offsets[0] must be 0");
  int max_n = 0;
  for (int s = 0; s < n_segments; ++s) {
    MMT_REQUIRE(seg_off[s + 1] >= seg_off[s], "nms: seg_offsets must be non-decreasing");
    max_n = max(max_n, seg_off[s + 1] - seg_off[s]);
  }
  MMT_REQUIRE(n_keep_out != nullptr, "nms: n_keep_out is null");
  if (total == 0) {
    MMT_CUDA(cudaMemsetAsync(n_keep_out, 0, sizeof(int) * n_segments, stream));
    return MMT_OK;
  }
  MMT_REQUIRE(boxes && scores && keep_out && workspace, "nms: null pointer");
  MMT_REQUIRE((reinterpret_cast<uintptr_t>(boxes) & 15) == 0, "nms: boxes must be 16-byte aligned");
  const size_t words = mask_words(seg_off, n_segments);
  size_t need = mmt::align_up((size_t)total * sizeof(int), 256) +
                mmt::align_up((size_t)total * sizeof(float4), 256) +
                mmt::align_up((size_t)total, 256) + mmt::align_up(words * 8, 256);
  if (workspace_bytes < need)
    return mmt::fail(MMT_ERR_WORKSPACE_TOO_SMALL, "nms: workspace %zu < required %zu",
                     workspace_bytes, need);
  const int max_cb = (max_n + kTile - 1) / kTile;
  const size_t scan_smem = (size_t)(max_cb + kTile) * 8;
  if (scan_smem > 200 * 1024)
    return mmt::fail(MMT_ERR_UNSUPPORTED, "nms: segment of %d boxes exceeds the on-chip scan limit", max_n);

  char* p = (char*)workspace;
  Workspace ws;
  ws.order = (int*)p;
  p += mmt::align_up((size_t)total * sizeof(int), 256);
  ws.sorted = (float4*)p;
  p += mmt::align_up((size_t)total * sizeof(float4), 256);
  ws.flags = flags_ext ? flags_ext : (unsigned char*)p;   // per-row keep flags, original row order
  p += mmt::align_up((size_t)total, 256);
  ws.mask = (unsigned long long*)p;

  if (scan_smem > 48 * 1024)
    MMT_CUDA(cudaFuncSetAttribute(greedy_scan_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)scan_smem));

  long long mask_cursor = 0;
  for (int s0 = 0; s0 < n_segments; s0 += kMaxSegPerLaunch) {
    const int ns = min(kMaxSegPerLaunch, n_segments - s0);
    SegTable tab;
    tab.n_seg = ns;
    int chunk_max = 0;
    for (int s = 0; s <= ns; ++s) tab.box_off[s] = seg_off[s0 + s];
    for (int s = 0; s < ns; ++s) {
      const long long n = seg_off[s0 + s + 1] - seg_off[s0 + s];
      tab.mask_off[s] = mask_cursor;
      mask_cursor += n * ((n + kTile - 1) / kTile);
      chunk_max = max(chunk_max, (int)n);
    }
    tab.mask_off[ns] = mask_cursor;
    if (chunk_max == 0) {
      MMT_CUDA(cudaMemsetAsync(n_keep_out + s0, 0, sizeof(int) * ns, stream));
      continue;
    }
    const int cbm = (chunk_max + kTile - 1) / kTile;
    rank_sort_kernel<<<dim3((chunk_max + 255) / 256, ns), 256, 0, stream>>>(boxes, scores, tab,
                                                                             ws.order, ws.sorted);
    if (int rc = mmt::check_launch("nms rank_sort")) return rc;
    iou_mask_kernel<<<dim3(cbm, cbm, ns), kTile, 0, stream>>>(ws.sorted, tab, thr, ws.mask);
    if (int rc = mmt::check_launch("nms iou_mask")) return rc;
    greedy_scan_kernel<<<ns, 1024, (size_t)(cbm + kTile) * 8, stream>>>(
        ws.mask, ws.order, tab, ws.flags, (long long*)keep_out, n_keep_out, s0);
    if (int rc = mmt::check_launch("nms greedy_scan")) return rc;
  }
  return MMT_OK;
}

extern "C" int mmt_nms_segmented_f32(const float* boxes, const float* scores, const int* seg_off,
                                     int n_segments, float thr, int64_t* keep_out, int* n_keep_out,
                                     void* workspace, size_t workspace_bytes, void* stream_) {
  return nms_segmented_impl(boxes, scores, seg_off, n_segments, thr, keep_out, n_keep_out, nullptr, workspace,
                            workspace_bytes, stream_);
}

extern "C" int mmt_nms_segmented_flags_f32(const float* boxes, const float* scores, const int* seg_off,
                                           int n_segments, float thr, int64_t* keep_out, int* n_keep_out,
                                           unsigned char* keep_flags_out, void* workspace, size_t workspace_bytes,
                                           void* stream_) {
  MMT_REQUIRE(keep_flags_out != nullptr, "nms: keep_flags_out is null");
  return nms_segmented_impl(boxes, scores, seg_off, n_segments, thr, keep_out, n_keep_out, keep_flags_out, workspace,
                            workspace_bytes, stream_);
}

extern "C" int mmt_nms_f32(const float* boxes, const float* scores, int n, float thr,
                           int64_t* keep_out, int* n_keep_out, void* workspace,
                           size_t workspace_bytes, void* stream) {
  MMT_REQUIRE(n >= 0, "nms: negative box count");
  int off[2] = {0, n};
  return mmt_nms_segmented_f32(boxes, scores, off, 1, thr, keep_out, n_keep_out, workspace,
                               workspace_bytes, stream);
}
