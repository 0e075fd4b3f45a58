// TMA + tcgen05 (TF32, fp32 accumulate in TMEM) implicit-GEMM convolution engine for sm_100a.
//
// Replaces cuDNN behind torch.nn.Conv2d (+ FrozenBatchNorm2d + ReLU + residual add) on the path
//   maskrcnn_benchmark/layers/misc.py:30-43, layers/batch_norm.py:19-24,
//   modeling/backbone/resnet.py:254-275, modeling/backbone/fpn.py:43-69, modeling/rpn/rpn.py:39-46,
//   modeling/roi_heads/mask_head/roi_mask_feature_extractors.py:131-146 and the box-head FCs.
//
// Data layout: activations NHWC fp32, filters [out_c][k_h][k_w][in_c] fp32 (K contiguous).
//
// forward / data-gradient kernel (conv_tc_kernel)
//   GEMM view  D[128 pixels][BN out channels] += A[128 pixels][32 ch of one tap] * B[BN][32]
//   * the 128 pixels of a tile are a TW x TH spatial patch of ONE image; the A tile of filter tap
//     (r,s) is the same patch shifted by (r-pad, s-pad), fetched by ONE 4-D TMA box load whose
//     out-of-bounds elements are zero-filled by the hardware — the implicit-GEMM "im2col" costs
//     no instructions and no extra traffic.  1x1 convs flatten (n,h,w) into one axis (no tile
//     waste); 1x1/stride-2 convs read through a strided tensor map.
//   * warp roles: warp 0 = TMA producer, warp 1 = tcgen05.mma issuer (one elected lane) + TMEM
//     allocator, warps 2-5 = epilogue (TMEM -> registers -> scale/bias/residual/ReLU -> swizzled
//     smem -> TMA store).  smem ring of STAGES (A 16 KB + B BN*128 B) slots with full/empty
//     mbarriers; the accumulator is double-buffered in TMEM (2*BN columns) so the epilogue of
//     tile i overlaps the main loop of tile i+1; persistent CTAs, one per SM.
// weight-gradient kernel (wgrad_tc_kernel)
//   D[128 out_c][BN in_c] (one filter tap) += dy^T * x over 32-pixel patches; both operands are
//   MN-major (pixels are the GEMM K axis), split-K over pixel ranges, partial tiles are summed
//   into dw by TMA reduce-add (cp.reduce.async.bulk.tensor ... .add) — no atomics in SM code.
#include <cuda.h>
#include <atomic>
#include <cmath>
#include <mutex>

#include "common.cuh"
#include "conv_common.cuh"
#include "tc_ptx.cuh"

namespace {

using namespace tc;

constexpr int kThreads = 192;       // wgrad: TMA warp + MMA warp + 4 epilogue warps
constexpr int kFpropThreads = 320;  // fprop / dgrad: TMA warp + MMA warp + 8 epilogue warps (two per TMEM lane quadrant)
constexpr int kTileM = 128;
constexpr int kABytes = kTileM * 128;  // 128 rows x 32 fp32
constexpr int kStagingBytes = kTileM * 128;
constexpr int kSmemLimit = 232448;  // 227 KB

// CG = CTAs cooperating on one MMA (tcgen05 cta_group): 1, or 2 = a CTA pair computing a 256 x BN tile, each CTA staging
// its own 128 rows of A and BN / 2 rows of B.
template <int BN, int CG = 1>
struct Cfg {
  static constexpr int kBRows = BN / CG;
  static constexpr int kBBytes = kBRows * 128;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kAffineBytes = 2 * BN * 4;  // per-tile copy of the epilogue scale / shift vectors
  static constexpr int kFixedBytes = 2 * kStagingBytes + kAffineBytes + 256 /*barriers*/;
  static constexpr int kStagesRaw = (kSmemLimit - kFixedBytes) / kStageBytes;
  static constexpr int kStages = kStagesRaw > 8 ? 8 : kStagesRaw;
  static constexpr int kTmemCols = 2 * BN;  // 128 / 256 / 512: powers of two >= 32
  // the dynamic shared window is declared __align__(1024) (SWIZZLE_128B atoms); the kernels trap if it is not
  static constexpr int kSmemBytes = kStages * kStageBytes + kFixedBytes;
};

struct FpropParams {
  int tiles_w, tiles_h, tiles_b;  // spatial tiles per image and images
  int n_tiles;                    // ceil(out channels / BN)
  int n_out;                      // out channels (epilogue guard; the TMA store clips)
  int TW, TH;
  int kh, kw, cblocks;  // cblocks = channels of A / 32
  int pad_h, pad_w;
  int y_w, y_h;  // output extent (residual bounds)
  long long r_sb, r_sh, r_sw;
  const float* scale;
  const float* bias;
  const float* residual;
  int res_mode;  // 0: y += residual;  1: y = residual > 0 ? y : 0  (ReLU backward of the layer that produced this conv's input)
  int relu;
  int total_tiles;  // work units: (pixel tile, column tile) pairs; with CTA pairs a unit covers TWO consecutive pixel tiles
  int m_tiles;      // pixel tiles
  int stages;       // ring slots actually used (<= the kernel's compile-time STAGES)
  int prefetch;     // K blocks the producer runs its L2 prefetch ahead of its loads (0 = off)
  int sub_w;  // per-warp store box: 32 pixels = sub_w x (32 / sub_w) of the TW x TH patch
};

__device__ __forceinline__ void named_bar_sync(int id, int threads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory");
}

// Shared-memory map of one CTA.  Every address is affine in the stage index, so the per-stage lookups are computed
// (base + stage * stride) instead of being read from an array indexed by a runtime value: such an array lives in
// LOCAL memory (r01: a 160-byte stack frame, 20 STL in every prologue and an LDL per lookup inside the producer / MMA loops).
struct StageAddr {
  uint32_t base, stride;
  __device__ __forceinline__ uint32_t operator[](int s) const { return base + (uint32_t)s * stride; }
};
struct SmemLayout {
  StageAddr a, b;
  uint32_t staging;  // epilogue warps x 4 KB (32 rows x 128 B, SWIZZLE_128B)
  uint32_t affine;   // scale[BN] then shift[BN]
  StageAddr full, empty, tfull, tempty;
  uint32_t tmem_slot;
};

template <int STAGES, int A_BYTES, int B_BYTES, int AFFINE_BYTES>
__device__ __forceinline__ SmemLayout carve(uint8_t* raw) {
  SmemLayout L;
  const uint32_t base = smem_u32(raw);
  if (base & 1023u) __trap();  // extern __shared__ __align__(1024) was not honoured
  L.a = StageAddr{base, (uint32_t)(A_BYTES + B_BYTES)};
  L.b = StageAddr{base + (uint32_t)A_BYTES, (uint32_t)(A_BYTES + B_BYTES)};
  uint32_t p = base + STAGES * (A_BYTES + B_BYTES);
  L.staging = p;
  p += 2 * kStagingBytes;
  L.affine = p;
  p += AFFINE_BYTES;
  L.full = StageAddr{p, 8u};
  L.empty = StageAddr{p + 64u, 8u};
  L.tfull = StageAddr{p + 128u, 8u};
  L.tempty = StageAddr{p + 144u, 8u};
  L.tmem_slot = p + 160;
  return L;
}

__device__ __forceinline__ uint32_t ld_shared_u32(uint32_t addr) {
  uint32_t v;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ void st_shared_v4(uint32_t addr, float x, float y, float z, float w) {
  asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(x), "f"(y), "f"(z), "f"(w) : "memory");
}
__device__ __forceinline__ float4 ld_shared_v4(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ void prefetch_l2(const void* p) {
  asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
}
__device__ __forceinline__ void st_shared_f32(uint32_t addr, float x) {
  asm volatile("st.shared.f32 [%0], %1;" ::"r"(addr), "f"(x) : "memory");
}

template <int BN, int CG>
__global__ void __launch_bounds__(kFpropThreads, 1)
conv_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
               const __grid_constant__ CUtensorMap tmY, const FpropParams p) {
  using C = Cfg<BN, CG>;
  constexpr int STAGES = C::kStages;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const SmemLayout L = carve<STAGES, kABytes, C::kBBytes, C::kAffineBytes>(smem_raw);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // CTA pair: rank 0 (leader) issues the MMAs; both CTAs load operands and run the epilogue of their own 128 rows
  const uint32_t rank = CG == 2 ? cluster_ctarank() : 0u;
  const bool leader = rank == 0;

  if (warp == 0 && lane == 0) {
    prefetch_tensormap(&tmA);
    prefetch_tensormap(&tmB);
    prefetch_tensormap(&tmY);
  }
  if (warp == 1) {
    if (lane == 0) {
      for (int s = 0; s < STAGES; ++s) {
        mbar_init(L.full[s], 1);
        mbar_init(L.empty[s], 1);
      }
      mbar_init(L.tfull[0], 1);
      mbar_init(L.tfull[1], 1);
      mbar_init(L.tempty[0], 8 * CG);  // one arrival per epilogue warp (of both CTAs of a pair, on the leader's barrier)
      mbar_init(L.tempty[1], 8 * CG);
      fence_barrier_init();
    }
    __syncwarp();
    if constexpr (CG == 2) {
      tmem_alloc_cg2(L.tmem_slot, C::kTmemCols);
      tmem_relinquish_cg2();
    } else {
      tmem_alloc(L.tmem_slot, C::kTmemCols);
      tmem_relinquish();
    }
  }
  tc_fence_before();
  __syncthreads();
  if constexpr (CG == 2) cluster_sync_all();  // the peer's barriers are initialised and its TMEM allocated before anything remote
  tc_fence_after();
  const uint32_t tmem_base = ld_shared_u32(L.tmem_slot);
  // PDL: the prologue above ran while the previous kernel of the stream was finishing; from here on this kernel reads
  // (activations, a just-transposed filter) and writes global memory, so the predecessor must have completed
  pdl_launch_dependents();
  pdl_wait();

  const int kblocks = p.kh * p.kw * p.cblocks;
  const int tiles_per_img = p.tiles_w * p.tiles_h;
  const int unit0 = CG == 2 ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
  const int n_units = CG == 2 ? (int)(gridDim.x >> 1) : (int)gridDim.x;
  // work unit -> this CTA's (column tile, pixel tile origin).  In a pair CTA `rank` takes pixel tile 2 * i + rank; the odd
  // tile out at the end is computed twice (dup: same loads and MMA rows, no stores) so that every TMA box stays in bounds
  auto locate = [&](int t, int& nt, int& b, int& h0, int& w0, bool& dup) {
    nt = t % p.n_tiles;
    int mt = t / p.n_tiles;
    dup = false;
    if constexpr (CG == 2) {
      mt = mt * 2 + (int)rank;
      if (mt >= p.m_tiles) { dup = true; mt = p.m_tiles - 1; }
    }
    b = mt / tiles_per_img;
    const int rem = mt % tiles_per_img;
    h0 = (rem / p.tiles_w) * p.TH;
    w0 = (rem % p.tiles_w) * p.TW;
  };

  if (warp == 0) {
    if (lane == 0) {
      // pair: the transaction bytes of BOTH CTAs' loads are counted on the leader's full barrier
      const uint32_t full_base = CG == 2 ? mapa_shared(L.full[0], 0) : L.full[0];
      int stage = 0;
      uint32_t phase = 0;
      const int n_stage = p.stages;
      for (int t = unit0; t < p.total_tiles; t += n_units) {
        int nt, b, h0, w0;
        bool dup;
        locate(t, nt, b, h0, w0, dup);
        // L2 prefetch runs `p.prefetch` K blocks ahead of the loads (the A operand streams from HBM once per tile and tap;
        // the load that follows finds it in L2): prologue here, steady state inside the loop
        const int taps_w = p.kw, cbs = p.cblocks;
        auto prefetch_kb = [&](int q) {
          const int cb = q % cbs, tap = q / cbs;
          const int r = tap / taps_w, s2 = tap % taps_w;
          tma_prefetch_4d(&tmA, cb * 32, w0 + s2 - p.pad_w, h0 + r - p.pad_h, b);
        };
        if (p.prefetch > 0)
          for (int q = 0; q < min(p.prefetch, kblocks); ++q) prefetch_kb(q);
        int kb = 0;
        for (int r = 0; r < p.kh; ++r)
          for (int s = 0; s < p.kw; ++s)
            for (int cb = 0; cb < p.cblocks; ++cb, ++kb) {
              if (p.prefetch > 0 && kb + p.prefetch < kblocks) prefetch_kb(kb + p.prefetch);
              mbar_wait(L.empty[stage], phase ^ 1u);
              if constexpr (CG == 2) {
                const uint32_t full = full_base + 8u * (uint32_t)stage;
                if (leader) mbar_arrive_expect_tx(L.full[stage], 2u * (uint32_t)C::kStageBytes);
                tma_load_4d_cg2(L.a[stage], &tmA, full, cb * 32, w0 + s - p.pad_w, h0 + r - p.pad_h, b);
                tma_load_2d_cg2(L.b[stage], &tmB, full, kb * 32, nt * BN + (int)rank * C::kBRows);
              } else {
                mbar_arrive_expect_tx(L.full[stage], (uint32_t)C::kStageBytes);
                tma_load_4d(L.a[stage], &tmA, L.full[stage], cb * 32, w0 + s - p.pad_w, h0 + r - p.pad_h, b);
                tma_load_2d(L.b[stage], &tmB, L.full[stage], kb * 32, nt * BN);
              }
              if (++stage == n_stage) { stage = 0; phase ^= 1u; }
            }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (lane == 0 && leader) {
      constexpr uint32_t idesc = umma_idesc_tf32(kTileM * CG, BN, 0, 0);
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int t = unit0; t < p.total_tiles; t += n_units, ++it) {
        const int acc = it & 1;
        const uint32_t acc_phase = (uint32_t)(it >> 1) & 1u;
        mbar_wait(L.tempty[acc], acc_phase ^ 1u);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * BN);
        for (int kb = 0; kb < kblocks; ++kb) {
          mbar_wait(L.full[stage], phase);
          tc_fence_after();
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const uint64_t da = umma_desc_kmajor_sw128(L.a[stage] + k * 32);
            const uint64_t db = umma_desc_kmajor_sw128(L.b[stage] + k * 32);
            if constexpr (CG == 2) mma_tf32_ss_cg2(d_tmem, da, db, idesc, (kb | k) ? 1u : 0u);
            else mma_tf32_ss(d_tmem, da, db, idesc, (kb | k) ? 1u : 0u);
          }
          if constexpr (CG == 2) mma_commit_cg2(L.empty[stage], 3);  // frees the stage in BOTH CTAs
          else mma_commit(L.empty[stage]);
          if (++stage == p.stages) { stage = 0; phase ^= 1u; }
        }
        if constexpr (CG == 2) mma_commit_cg2(L.tfull[acc], 3);
        else mma_commit(L.tfull[acc]);
      }
    }
    __syncwarp();
  } else {
    // ------------------------------------------------------------------ epilogue (warps 2..9)
    // Two warps per TMEM lane quadrant (= 32 pixels of the tile), taking alternate 32-column chunks, so every SM
    // sub-partition has two epilogue warps to hide each other's TMEM / shared-memory / residual-load latencies.
    // Each warp runs on its own: one private 4 KB staging buffer, private TMA stores (bulk groups are per thread:
    // lane 0), no CTA-wide barrier per chunk.
    const int q = warp & 3;            // TMEM lane quadrant this warp may access
    const int half = (warp - 2) >> 2;  // which of the quadrant's two warps: chunks half, half + 2, ...
    const int epi_tid = threadIdx.x - 64;
    const uint32_t buf = L.staging + (uint32_t)(warp - 2) * 4096u;
    const int sub_tw0 = (q * 32) % p.TW, sub_th0 = (q * 32) / p.TW;  // origin of this warp's 32-pixel store box
    const bool affine = (p.scale != nullptr) || (p.bias != nullptr);
    const bool has_res = p.residual != nullptr;
    const uint32_t tempty_base = CG == 2 ? mapa_shared(L.tempty[0], 0) : L.tempty[0];
    // coalesced residual fetch: instruction i reads rows 4i + lane/8 (four whole 128-byte lines per warp instruction)
    const int col4 = lane & 7;
    int rtw[8], rth[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int r = q * 32 + 4 * i + (lane >> 3);
      rtw[i] = r % p.TW;
      rth[i] = r / p.TW;
    }
    int it = 0;
    int affine_nt = -1;
    for (int t = unit0; t < p.total_tiles; t += n_units, ++it) {
      int nt, b, h0, w0;
      bool dup;
      locate(t, nt, b, h0, w0, dup);
      const int acc = it & 1;
      const uint32_t acc_phase = (uint32_t)(it >> 1) & 1u;
      const int n0 = nt * BN;
      const int chunks = dup ? 0 : min(BN / 32, (p.n_out - n0 + 31) / 32);

      if (affine && nt != affine_nt) {  // (re)load this column block's scale / shift: uniform over the 8 warps
        named_bar_sync(1, 256);         // every warp is done reading the previous block's values
        for (int c = epi_tid; c < BN; c += 256) {
          const int col = n0 + c;
          const bool ok = col < p.n_out;
          st_shared_f32(L.affine + 4 * c, (ok && p.scale) ? __ldg(p.scale + col) : 1.f);
          st_shared_f32(L.affine + 4 * (BN + c), (ok && p.bias) ? __ldg(p.bias + col) : 0.f);
        }
        named_bar_sync(1, 256);
        affine_nt = nt;
      }

      // residual rows of this tile (element offsets) and the first chunk's prefetch, before waiting on the MMAs
      long long roff[8];
      bool rok[8];
      float4 rr[8];
      if (has_res && !dup) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          rok[i] = (w0 + rtw[i] < p.y_w) && (h0 + rth[i] < p.y_h);
          roff[i] = b * p.r_sb + (long long)(h0 + rth[i]) * p.r_sh + (long long)(w0 + rtw[i]) * p.r_sw + col4 * 4;
        }
        // pull this warp's share of the residual slab (32 rows x its chunks x 128 B) towards L2 now: the per-chunk
        // loads below then see L2 latency instead of DRAM latency
        if (col4 < chunks && (col4 & 1) == half) {
#pragma unroll
          for (int i = 0; i < 8; ++i)
            if (rok[i]) prefetch_l2(p.residual + roff[i] - col4 * 4 + n0 + col4 * 32);
        }
        const int cfirst = n0 + half * 32;
#pragma unroll
        for (int i = 0; i < 8; ++i)
          rr[i] = (rok[i] && half < chunks && cfirst + col4 * 4 < p.n_out)
                      ? __ldg(reinterpret_cast<const float4*>(p.residual + roff[i] + cfirst))
                      : make_float4(0.f, 0.f, 0.f, 0.f);
      }

      mbar_wait(L.tfull[acc], acc_phase);
      tc_fence_after();
#pragma unroll 1
      for (int j = half; j < chunks; j += 2) {
        const int c0 = n0 + j * 32;
        if (lane == 0) tma_wait_group_read<0>();  // this warp's previous store has read the staging buffer
        __syncwarp();
        uint32_t v[32];
        tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * BN + j * 32), v);
        if (has_res) {
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int r = 4 * i + (lane >> 3);
            st_shared_v4(buf + r * 128 + ((col4 ^ (r & 7)) << 4), rr[i].x, rr[i].y, rr[i].z, rr[i].w);
          }
          if (j + 2 < chunks) {  // prefetch this warp's next chunk of the residual while this one is processed
#pragma unroll
            for (int i = 0; i < 8; ++i)
              rr[i] = (rok[i] && c0 + 64 + col4 * 4 < p.n_out)
                          ? __ldg(reinterpret_cast<const float4*>(p.residual + roff[i] + c0 + 64))
                          : make_float4(0.f, 0.f, 0.f, 0.f);
          }
          __syncwarp();
        }
        tmem_ld_wait();
#pragma unroll
        for (int g = 0; g < 8; ++g) {
          float o0 = __uint_as_float(v[g * 4 + 0]), o1 = __uint_as_float(v[g * 4 + 1]);
          float o2 = __uint_as_float(v[g * 4 + 2]), o3 = __uint_as_float(v[g * 4 + 3]);
          if (affine) {
            const float4 sc = ld_shared_v4(L.affine + 4 * (j * 32 + g * 4));
            const float4 sh = ld_shared_v4(L.affine + 4 * (BN + j * 32 + g * 4));
            o0 = fmaf(o0, sc.x, sh.x); o1 = fmaf(o1, sc.y, sh.y);
            o2 = fmaf(o2, sc.z, sh.z); o3 = fmaf(o3, sc.w, sh.w);
          }
          const uint32_t addr = buf + lane * 128 + ((g ^ (lane & 7)) << 4);
          if (has_res) {
            const float4 r4 = ld_shared_v4(addr);
            if (p.res_mode) {
              o0 = r4.x > 0.f ? o0 : 0.f; o1 = r4.y > 0.f ? o1 : 0.f;
              o2 = r4.z > 0.f ? o2 : 0.f; o3 = r4.w > 0.f ? o3 : 0.f;
            } else {
              o0 += r4.x; o1 += r4.y; o2 += r4.z; o3 += r4.w;
            }
          }
          if (p.relu) {
            o0 = fmaxf(o0, 0.f); o1 = fmaxf(o1, 0.f); o2 = fmaxf(o2, 0.f); o3 = fmaxf(o3, 0.f);
          }
          st_shared_v4(addr, o0, o1, o2, o3);
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) {
          tma_store_4d(&tmY, buf, c0, w0 + sub_tw0, h0 + sub_th0, b);
          tma_commit_group();
        }
      }
      // this warp has finished its tcgen05.ld of the accumulator
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if constexpr (CG == 2) mbar_arrive_cluster(tempty_base + 8u * (uint32_t)acc);
        else mbar_arrive(L.tempty[acc]);
      }
    }
    if (lane == 0) tma_wait_group<0>();
  }

  tc_fence_before();
  __syncthreads();
  if constexpr (CG == 2) cluster_sync_all();  // nobody leaves while the leader may still read this CTA's smem / write its TMEM
  if (warp == 1) {
    tc_fence_after();
    if constexpr (CG == 2) tmem_dealloc_cg2(tmem_base, C::kTmemCols);
    else tmem_dealloc(tmem_base, C::kTmemCols);
  }
}

// ------------------------------------------------------------------------------------ wgrad
struct WgradParams {
  int tiles_w, tiles_h, tiles_b;  // 32-pixel patches per image, images
  int TW, TH;
  int kh, kw, pad_h, pad_w;
  int co_tiles, ci_tiles, splits;
  int pixel_tiles, tiles_per_split;
  int cin;
  int cout;
  const float* row_scale;  // nullable: per-out-channel multiplier of the result (frozen-BN scale: dw = scale[o] * wgrad(x, g_bn_out))
  int total_items;
};

template <int BN>
__global__ void __launch_bounds__(kThreads, 1)
wgrad_tc_kernel(const __grid_constant__ CUtensorMap tmDY, const __grid_constant__ CUtensorMap tmX,
                const __grid_constant__ CUtensorMap tmDW, const WgradParams p) {
  using C = Cfg<BN>;
  constexpr int STAGES = C::kStages;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const SmemLayout L = carve<STAGES, kABytes, C::kBBytes, C::kAffineBytes>(smem_raw);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    prefetch_tensormap(&tmDY);
    prefetch_tensormap(&tmX);
    prefetch_tensormap(&tmDW);
  }
  if (warp == 1) {
    if (lane == 0) {
      for (int s = 0; s < STAGES; ++s) {
        mbar_init(L.full[s], 1);
        mbar_init(L.empty[s], 1);
      }
      mbar_init(L.tfull[0], 1);
      mbar_init(L.tfull[1], 1);
      mbar_init(L.tempty[0], 1);
      mbar_init(L.tempty[1], 1);
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc(L.tmem_slot, C::kTmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = ld_shared_u32(L.tmem_slot);
  pdl_launch_dependents();
  pdl_wait();
  const int tiles_per_img = p.tiles_w * p.tiles_h;
  const int taps = p.kh * p.kw;

  // work item -> (co tile, tap, ci tile, split)
  auto decode = [&](int item, int& co_t, int& tap, int& ci_t, int& pt_begin, int& pt_end) {
    const int split = item % p.splits;
    int r = item / p.splits;
    ci_t = r % p.ci_tiles; r /= p.ci_tiles;
    tap = r % taps;
    co_t = r / taps;
    pt_begin = split * p.tiles_per_split;
    pt_end = min(p.pixel_tiles, pt_begin + p.tiles_per_split);
  };

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int item = blockIdx.x; item < p.total_items; item += gridDim.x) {
        int co_t, tap, ci_t, pb, pe;
        decode(item, co_t, tap, ci_t, pb, pe);
        const int r = tap / p.kw, s = tap % p.kw;
        for (int pt = pb; pt < pe; ++pt) {
          const int b = pt / tiles_per_img, rem = pt % tiles_per_img;
          const int h0 = (rem / p.tiles_w) * p.TH, w0 = (rem % p.tiles_w) * p.TW;
          mbar_wait(L.empty[stage], phase ^ 1u);
          mbar_arrive_expect_tx(L.full[stage], (uint32_t)C::kStageBytes);
#pragma unroll
          for (int g = 0; g < 4; ++g)
            tma_load_4d(L.a[stage] + g * 4096, &tmDY, L.full[stage], co_t * 128 + g * 32, w0, h0, b);
#pragma unroll
          for (int g = 0; g < BN / 32; ++g)
            tma_load_4d(L.b[stage] + g * 4096, &tmX, L.full[stage], ci_t * BN + g * 32,
                        w0 + s - p.pad_w, h0 + r - p.pad_h, b);
          if (++stage == STAGES) { stage = 0; phase ^= 1u; }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_tf32(kTileM, BN, 1, 1);
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int item = blockIdx.x; item < p.total_items; item += gridDim.x, ++it) {
        int co_t, tap, ci_t, pb, pe;
        decode(item, co_t, tap, ci_t, pb, pe);
        const int acc = it & 1;
        const uint32_t acc_phase = (uint32_t)(it >> 1) & 1u;
        mbar_wait(L.tempty[acc], acc_phase ^ 1u);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * BN);
        for (int pt = pb; pt < pe; ++pt) {
          mbar_wait(L.full[stage], phase);
          tc_fence_after();
#pragma unroll
          for (int k = 0; k < 4; ++k) {  // 4 x 8 pixels
            const uint64_t da = umma_desc_mnmajor_sw128_32b(L.a[stage] + k * 1024, 4096);
            const uint64_t db = umma_desc_mnmajor_sw128_32b(L.b[stage] + k * 1024, 4096);
            mma_tf32_ss(d_tmem, da, db, idesc, (pt > pb || k) ? 1u : 0u);
          }
          mma_commit(L.empty[stage]);
          if (++stage == STAGES) { stage = 0; phase ^= 1u; }
        }
        mma_commit(L.tfull[acc]);
      }
    }
    __syncwarp();
  } else {
    const int q = warp & 3;
    const int row = q * 32 + lane;  // out channel inside the 128-row tile
    const int epi_tid = threadIdx.x - 64;
    int it = 0;
    uint32_t chunk_counter = 0;
    for (int item = blockIdx.x; item < p.total_items; item += gridDim.x, ++it) {
      int co_t, tap, ci_t, pb, pe;
      decode(item, co_t, tap, ci_t, pb, pe);
      const int acc = it & 1;
      const uint32_t acc_phase = (uint32_t)(it >> 1) & 1u;
      const bool empty_range = pe <= pb;  // nothing accumulated: MMA warp still commits, D undefined
      const int orow = co_t * 128 + row;
      const float rs = empty_range ? 0.f : ((p.row_scale && orow < p.cout) ? __ldg(p.row_scale + orow) : 1.f);
      mbar_wait(L.tfull[acc], acc_phase);
      tc_fence_after();
#pragma unroll 1
      for (int j = 0; j < BN / 32; ++j, ++chunk_counter) {
        const uint32_t buf = L.staging + (chunk_counter & 1u) * (uint32_t)kStagingBytes;
        if (epi_tid == 0) tma_wait_group_read<1>();
        named_bar_sync(1, 128);
        uint32_t v[32];
        tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * BN + j * 32), v);
        tmem_ld_wait();
        tc_fence_before();
#pragma unroll
        for (int g = 0; g < 8; ++g)
          st_shared_v4(buf + row * 128 + ((g ^ (row & 7)) << 4),
                       empty_range ? 0.f : rs * __uint_as_float(v[g * 4 + 0]),
                       empty_range ? 0.f : rs * __uint_as_float(v[g * 4 + 1]),
                       empty_range ? 0.f : rs * __uint_as_float(v[g * 4 + 2]),
                       empty_range ? 0.f : rs * __uint_as_float(v[g * 4 + 3]));
        fence_proxy_async_smem();
        named_bar_sync(1, 128);
        if (epi_tid == 0) {
          tma_reduce_add_2d(&tmDW, buf, tap * p.cin + ci_t * BN + j * 32, co_t * 128);
          tma_commit_group();
        }
      }
      if (epi_tid == 0) mbar_arrive(L.tempty[acc]);
    }
    if (epi_tid == 0) tma_wait_group<0>();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, C::kTmemCols);
  }
}

// [out_c][kh][kw][in_c] -> [in_c][kh][kw][out_c], taps rotated 180 degrees
// `k_scale` (nullable): per-OUT-channel multiplier folded into the transformed filter — the frozen-BN scale of the
// layer, so that dgrad consumes the gradient of the BN OUTPUT directly (dx = conv_T(g * scale, w) = conv_T(g, w * scale))
__global__ void __launch_bounds__(256) filter_dgrad_kernel(const float* __restrict__ w, float* __restrict__ wt,
                                                           int out_c, int taps, int in_c,
                                                           const float* __restrict__ k_scale) {
  __shared__ float tile[32][33];
  // per tap: transpose the [out_c][in_c] matrix (row stride taps*in_c) into [in_c][out_c]
  const int tap = blockIdx.z;
  const int rtap = taps - 1 - tap;
  const int i0 = blockIdx.x * 32, o0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int k = ty; k < 32; k += 8) {
    const int o = o0 + k, i = i0 + tx;
    if (o < out_c && i < in_c) tile[k][tx] = w[((long long)o * taps + tap) * in_c + i] * (k_scale ? __ldg(k_scale + o) : 1.f);
  }
  __syncthreads();
  for (int k = ty; k < 32; k += 8) {
    const int i = i0 + k, o = o0 + tx;
    if (o < out_c && i < in_c) wt[((long long)i * taps + rtap) * out_c + o] = tile[tx][k];
  }
}

// ---------------------------------------------------------------------- stem (7x7 / stride 2, 3 channels)
// A 7x7 stride-2 convolution over C_in = 3 is a 4x4 stride-1 convolution over the 2x2 space-to-depth image
// (12 channels, zero-padded to one 32-channel K block): in[2y + r - 3] with r - 3 = 2a + dy, a in {-2..1}, dy in {0,1}
// -> s2d[y + a][dy] — which is a shape the implicit-GEMM kernel above runs as is (taps -2..+1 = "pad 2", TMA
// zero-fills the border).  x NHWC [B,H,W,3] -> s [B,ceil(H/2),ceil(W/2),32], channel = (dy*2 + dx)*3 + c.
__global__ void __launch_bounds__(256) stem_s2d_kernel(const float* __restrict__ x, float4* __restrict__ s, int B,
                                                       int H, int W, int H2, int W2, int x_is_nchw) {
  const long long total = (long long)B * H2 * W2 * 8;  // float4 slots: 8 per pixel
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int q = (int)(i & 7);
    const long long pix = i >> 3;
    const int x2 = (int)(pix % W2), y2 = (int)((pix / W2) % H2), b = (int)(pix / ((long long)W2 * H2));
    float v[4] = {0.f, 0.f, 0.f, 0.f};
    if (q < 3) {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int ch = q * 4 + e;  // < 12
        const int c = ch % 3, dx = (ch / 3) & 1, dy = ch / 6;
        const int yy = 2 * y2 + dy, xx = 2 * x2 + dx;
        if (yy < H && xx < W)
          v[e] = x_is_nchw ? __ldg(x + (((long long)b * 3 + c) * H + yy) * W + xx)
                           : __ldg(x + (((long long)b * H + yy) * W + xx) * 3 + c);
      }
    }
    s[i] = make_float4(v[0], v[1], v[2], v[3]);
  }
}
// w [out_c][7][7][3] -> w4 [out_c][4][4][32]
__global__ void __launch_bounds__(256) stem_filter_kernel(const float* __restrict__ w, float* __restrict__ w4, int out_c) {
  const int total = out_c * 16 * 32;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int ch = i & 31, tap = (i >> 5) & 15, o = i >> 9;
    float v = 0.f;
    if (ch < 12) {
      const int c = ch % 3, dx = (ch / 3) & 1, dy = ch / 6;
      const int r = 2 * ((tap >> 2) - 2) + dy + 3, t = 2 * ((tap & 3) - 2) + dx + 3;
      if (r >= 0 && r < 7 && t >= 0 && t < 7) v = __ldg(w + ((o * 7 + r) * 7 + t) * 3 + c);
    }
    w4[i] = v;
  }
}

// ------------------------------------------------------------------------------- host side
// Persistent-grid size.  One CTA per SM by default; mmt_conv_set_max_ctas(n) leaves 148 - n SMs to whatever runs on
// other streams (the trainer replays the conv trunks on a side stream while the latency-bound proposal / sampling
// glue runs on the main one: without free SMs every tiny glue kernel waits for a whole persistent conv to drain).
std::atomic<int> g_max_ctas{mmt::kNumSMs};
inline int persistent_grid(long long work_items) {
  return (int)std::min<long long>(work_items, g_max_ctas.load(std::memory_order_relaxed));
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                  CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                  CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)p;
  });
  return fn;
}

struct View4 {  // channels-innermost view: dims (c, w, h, n), element strides of w, h, n (c stride 1)
  const void* base;
  long long c, w, h, n;
  long long sw, sh, sn;
};

int make_map(CUtensorMap* m, const View4& v, int box_c, int box_w, int box_h, const char* what,
             CUtensorMapSwizzle swizzle = CU_TENSOR_MAP_SWIZZLE_128B) {
  EncodeTiledFn enc = get_encode();
  if (!enc) return mmt::fail(MMT_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
  cuuint64_t dims[4] = {(cuuint64_t)v.c, (cuuint64_t)v.w, (cuuint64_t)v.h, (cuuint64_t)v.n};
  cuuint64_t strides[3] = {(cuuint64_t)v.sw * 4, (cuuint64_t)v.sh * 4, (cuuint64_t)v.sn * 4};
  cuuint32_t box[4] = {(cuuint32_t)box_c, (cuuint32_t)box_w, (cuuint32_t)box_h, 1};
  cuuint32_t es[4] = {1, 1, 1, 1};
  if (((uintptr_t)v.base & 15) || (strides[0] & 15) || (strides[1] & 15) || (strides[2] & 15))
    return mmt::fail(MMT_ERR_INVALID_ARGUMENT, "%s: tensor must be 16-byte aligned with 16-byte strides", what);
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<void*>(v.base), dims, strides, box, es,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return mmt::fail(MMT_ERR_CUDA, "%s: cuTensorMapEncodeTiled failed (%d)", what, (int)r);
  return MMT_OK;
}

int make_map_2d(CUtensorMap* m, const void* base, long long inner, long long outer, long long row_stride,
                int box_inner, int box_outer, const char* what) {
  EncodeTiledFn enc = get_encode();
  if (!enc) return mmt::fail(MMT_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
  cuuint64_t dims[2] = {(cuuint64_t)inner, (cuuint64_t)outer};
  cuuint64_t strides[1] = {(cuuint64_t)row_stride * 4};
  cuuint32_t box[2] = {(cuuint32_t)box_inner, (cuuint32_t)box_outer};
  cuuint32_t es[2] = {1, 1};
  if (((uintptr_t)base & 15) || (strides[0] & 15))
    return mmt::fail(MMT_ERR_INVALID_ARGUMENT, "%s: matrix must be 16-byte aligned with 16-byte row stride", what);
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(base), dims, strides, box, es,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return mmt::fail(MMT_ERR_CUDA, "%s: cuTensorMapEncodeTiled failed (%d)", what, (int)r);
  return MMT_OK;
}

// choose the TW x TH (= pixels) patch that wastes the least area on a w x h map
void pick_patch(long long w, long long h, int pixels, int& TW, int& TH) {
  long long best = -1;
  for (int tw = pixels; tw >= 8; tw >>= 1) {
    const int th = pixels / tw;
    const long long area = mmt::ceil_div<long long>(w, tw) * tw * mmt::ceil_div<long long>(h, th) * th;
    if (best < 0 || area < best) { best = area; TW = tw; TH = th; }
  }
}

struct GemmProblem {
  View4 a;          // activations feeding the GEMM A operand (a.c = channels per tap)
  const float* b;   // [n_out][kh*kw*a.c]
  int n_out;
  int kh, kw, pad_h, pad_w;
  View4 y;          // output view, y.c == n_out
  const float* scale;
  const float* bias;
  const float* residual;  // addressed like y with strides r_*
  long long r_sw, r_sh, r_sn;
  int res_mode = 0;       // 0 add, 1 mask (see FpropParams)
  int relu;
  int prof_family;  // MMT_PROF_CONV_FPROP / MMT_PROF_CONV_DGRAD
  double flops_override = 0.0;  // algorithmic flops when the GEMM is padded (stem)
};

// algorithmic cost of one implicit GEMM: 2*M*N*K flops; bytes = A read once + filter + output (+ residual)
inline double gemm_flops(const GemmProblem& g) {
  if (g.flops_override > 0.0) return g.flops_override;
  return 2.0 * (double)g.y.n * g.y.h * g.y.w * g.n_out * g.kh * g.kw * g.a.c;
}
inline double gemm_bytes(const GemmProblem& g) {
  const double pix = (double)g.y.n * g.y.h * g.y.w;
  return 4.0 * (pix * g.a.c + (double)g.n_out * g.kh * g.kw * g.a.c + pix * g.n_out * (g.residual ? 2.0 : 1.0));
}

// Launch attributes of the tcgen05 kernels: the CTA pair (cluster of 2 = the two SMs of a TPC) and programmatic dependent
// launch (the kernel's prologue overlaps the tail of the previous kernel of the stream; see tc_ptx.cuh pdl_wait).
std::atomic<int> g_pdl{1};
std::atomic<int> g_dbg_stages{0};   // experiments: cap the ring depth (0 = all the shared memory allows)
std::atomic<int> g_prefetch{0};     // K blocks of L2 prefetch ahead of the loads
void fill_launch(cudaLaunchConfig_t& cfg, cudaLaunchAttribute* attr, int grid, int threads, size_t smem, cudaStream_t stream,
                 int cluster) {
  cfg.gridDim = dim3((unsigned)grid);
  cfg.blockDim = dim3((unsigned)threads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  int n = 0;
  if (cluster > 1) {
    attr[n].id = cudaLaunchAttributeClusterDimension;
    attr[n].val.clusterDim.x = (unsigned)cluster;
    attr[n].val.clusterDim.y = 1;
    attr[n].val.clusterDim.z = 1;
    ++n;
  }
  if (g_pdl.load(std::memory_order_relaxed)) {
    attr[n].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[n].val.programmaticStreamSerializationAllowed = 1;
    ++n;
  }
  cfg.attrs = attr;
  cfg.numAttrs = (unsigned)n;
}

template <int BN, int CG>
int launch_gemm(const GemmProblem& g, cudaStream_t stream) {
  using C = Cfg<BN, CG>;
  int TW, TH;
  pick_patch(g.y.w, g.y.h, kTileM, TW, TH);
  CUtensorMap tmA, tmB, tmY;
  if (int rc = make_map(&tmA, g.a, 32, TW, TH, "conv A operand")) return rc;
  const long long Ktot = (long long)g.kh * g.kw * g.a.c;
  if (int rc = make_map_2d(&tmB, g.b, Ktot, g.n_out, Ktot, 32, C::kBRows, "conv filter")) return rc;
  const int sub_w = TW < 32 ? TW : 32;  // each epilogue warp stores its own 32-pixel box
  if (int rc = make_map(&tmY, g.y, 32, sub_w, 32 / sub_w, "conv output")) return rc;
  FpropParams p;
  p.tiles_w = (int)mmt::ceil_div<long long>(g.y.w, TW);
  p.tiles_h = (int)mmt::ceil_div<long long>(g.y.h, TH);
  p.tiles_b = (int)g.y.n;
  p.n_tiles = (g.n_out + BN - 1) / BN;
  p.n_out = g.n_out;
  p.TW = TW; p.TH = TH;
  p.kh = g.kh; p.kw = g.kw; p.cblocks = (int)(g.a.c / 32);
  p.pad_h = g.pad_h; p.pad_w = g.pad_w;
  p.y_w = (int)g.y.w; p.y_h = (int)g.y.h;
  p.r_sb = g.r_sn; p.r_sh = g.r_sh; p.r_sw = g.r_sw;
  p.scale = g.scale; p.bias = g.bias; p.residual = g.residual; p.res_mode = g.res_mode; p.relu = g.relu;
  const long long m_tiles = (long long)p.tiles_w * p.tiles_h * p.tiles_b;
  const long long total = mmt::ceil_div<long long>(m_tiles, CG) * p.n_tiles;
  if (m_tiles * p.n_tiles > 0x7fffffffll) return mmt::fail(MMT_ERR_UNSUPPORTED, "conv: too many tiles");
  p.m_tiles = (int)m_tiles;
  p.total_tiles = (int)total;
  {
    const int dbg_stages = g_dbg_stages.load(std::memory_order_relaxed);
    p.stages = dbg_stages > 0 ? std::min(dbg_stages, C::kStages) : C::kStages;
    p.prefetch = g_prefetch.load(std::memory_order_relaxed);
  }
  p.sub_w = sub_w;
  static bool attr_set = false;
  if (!attr_set) {
    MMT_CUDA(cudaFuncSetAttribute(conv_tc_kernel<BN, CG>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::kSmemBytes));
    attr_set = true;
  }
  const int max_units = std::max(1, g_max_ctas.load(std::memory_order_relaxed) / CG);
  const int grid = CG * (int)std::min<long long>(total, max_units);
  mmt::ProfScope prof(g.prof_family, gemm_flops(g), gemm_bytes(g), stream);
  cudaLaunchConfig_t cfg = {};
  cudaLaunchAttribute attr[2];
  fill_launch(cfg, attr, grid, kFpropThreads, C::kSmemBytes, stream, CG);
  MMT_CUDA(cudaLaunchKernelEx(&cfg, conv_tc_kernel<BN, CG>, tmA, tmB, tmY, p));
  return mmt::check_launch("conv_tc_kernel");
}

// Tile configuration: column-tile width BN and CTAs per MMA (1, or 2 = cta_group::2 pair on a 256 x BN tile).  Every BN
// handles every n_out (weights beyond n_out are zero-filled by TMA, stores are clipped), so the choice is a cost question.
//
// One CTA per tile: per 32-channel K block a CTA streams (128 + BN) rows of 128 B from L2 into shared memory (~56 B/clk
// per SM sustained with every SM pulling) and the tensor core needs 2 * BN cycles for the four K = 8 MMAs; a launch takes
// ceil(tiles / CTAs) rounds of kblocks * max(those two).  Narrow tiles lose on operand re-reads, wide tiles on wave
// quantisation when there are few pixel tiles.
//
// CTA pairs: measured per layer (profiles/r02_conv_tile_cfg.log, every trunk conv at the batches the trainer runs): the
// pair wins 3-10 % where the main loop is long and the machine is full — K = k*k*C_in >= 512, >= 256 output channels,
// >= 48 pair tiles: the 3x3 256 -> 256 convs of FPN / RPN / layer3 / mask head, the 512 -> 1024 and 1024 -> 2048
// projections, layer4 at the teacher's batch — and LOSES 5-20 % on the short-K 1x1 convs (K <= 256), whose time is the
// epilogue's HBM traffic: a pair's two epilogues start and end together, a lone CTA's overlap other CTAs' main loops.
struct TileCfg { int bn, cg; };
std::atomic<int> g_force_cg{0};   // 0 = rule above, 1 / 2 = force (tests, per-layer benchmarks)

TileCfg pick_cfg(const GemmProblem& g) {
  int TW, TH;
  pick_patch(g.y.w, g.y.h, kTileM, TW, TH);
  const long long m_tiles = mmt::ceil_div<long long>(g.y.w, TW) * mmt::ceil_div<long long>(g.y.h, TH) * g.y.n;
  const int ctas = g_max_ctas.load(std::memory_order_relaxed);
  const int force = g_force_cg.load(std::memory_order_relaxed);
  const long long k_total = (long long)g.kh * g.kw * g.a.c;
  if (force == 2) return TileCfg{g.n_out > 128 ? 256 : 128, 2};
  if (force == 0 && ctas >= 2 && k_total >= 512 && g.n_out >= 256 &&
      mmt::ceil_div<long long>(m_tiles, 2) * ((g.n_out + 255) / 256) >= 48)
    return TileCfg{256, 2};
  TileCfg best{64, 1};
  double best_cost = -1.0;
  for (int bn = 256; bn >= 32; bn >>= 1) {
    if (bn > 32 && bn / 2 >= g.n_out) continue;          // no tile that is mostly padding
    if (bn == 32 && g.n_out > 32) continue;              // N = 32 MMAs only for the narrow heads
    const long long units = m_tiles * ((g.n_out + bn - 1) / bn);
    const double rounds = (double)mmt::ceil_div<long long>(units, std::max(1, ctas));
    const double load_clk = (128.0 + bn) * 128.0 / 56.0, mma_clk = 2.0 * bn;
    // + a per-round fixed cost (pipeline fill, epilogue of the last tile): ~4 K-blocks' worth
    const double cost = rounds * std::max(load_clk, mma_clk) * (1.0 + 4.0 / (double)(k_total / 32));
    if (best_cost < 0 || cost < best_cost * 0.999) { best_cost = cost; best = TileCfg{bn, 1}; }
  }
  return best;
}

int dispatch_gemm(const GemmProblem& g, cudaStream_t stream) {
  const TileCfg c = pick_cfg(g);
  if (c.cg == 2) return c.bn == 256 ? launch_gemm<256, 2>(g, stream) : launch_gemm<128, 2>(g, stream);
  switch (c.bn) {
    case 256: return launch_gemm<256, 1>(g, stream);
    case 128: return launch_gemm<128, 1>(g, stream);
    case 64: return launch_gemm<64, 1>(g, stream);
    default: return launch_gemm<32, 1>(g, stream);
  }
}

View4 dense_nhwc(const float* base, long long n, long long h, long long w, long long c) {
  return View4{base, c, w, h, n, c, w * c, h * w * c};
}
View4 flat_pixels(const float* base, long long pixels, long long c) {
  return View4{base, c, pixels, 1, 1, c, pixels * c, pixels * c};
}

template <int BN>
int launch_wgrad(const CUtensorMap& tmDY, const CUtensorMap& tmX, const CUtensorMap& tmDW,
                 const WgradParams& p, cudaStream_t stream) {
  using C = Cfg<BN>;
  static bool attr_set = false;
  if (!attr_set) {
    MMT_CUDA(cudaFuncSetAttribute(wgrad_tc_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::kSmemBytes));
    attr_set = true;
  }
  const int grid = persistent_grid(p.total_items);
  cudaLaunchConfig_t cfg = {};
  cudaLaunchAttribute attr[2];
  fill_launch(cfg, attr, grid, kThreads, C::kSmemBytes, stream, 1);
  MMT_CUDA(cudaLaunchKernelEx(&cfg, wgrad_tc_kernel<BN>, tmDY, tmX, tmDW, p));
  return mmt::check_launch("wgrad_tc_kernel");
}

}  // namespace

extern "C" int mmt_conv_set_pipeline(int stages, int prefetch_kblocks) {
  if (stages < 0 || stages > 8 || prefetch_kblocks < 0 || prefetch_kblocks > 64)
    return mmt::fail(MMT_ERR_INVALID_ARGUMENT, "conv_set_pipeline: stages in [0, 8], prefetch in [0, 64]");
  g_dbg_stages.store(stages);
  g_prefetch.store(prefetch_kblocks);
  return MMT_OK;
}

extern "C" int mmt_conv_set_pdl(int on) {
  g_pdl.store(on ? 1 : 0);
  return MMT_OK;
}

extern "C" int mmt_conv_force_cta_group(int cg) {
  if (cg < 0 || cg > 2) return mmt::fail(MMT_ERR_INVALID_ARGUMENT, "conv_force_cta_group: %d not in {0, 1, 2}", cg);
  g_force_cg.store(cg);
  return MMT_OK;
}

extern "C" int mmt_conv_set_max_ctas(int n) {
  if (n < 1 || n > mmt::kNumSMs) return mmt::fail(MMT_ERR_INVALID_ARGUMENT, "conv_set_max_ctas: %d not in [1, %d]", n, mmt::kNumSMs);
  g_max_ctas.store(n);
  return MMT_OK;
}

namespace mmt {

bool conv_tc_supported_fprop(const mmt_conv_shape& s) {
  if (s.in_c % 32 || s.out_c % 4) return false;
  if (s.stride == 1) return s.out_h == s.in_h + 2 * s.pad - s.k_h + 1 && s.out_w == s.in_w + 2 * s.pad - s.k_w + 1;
  return s.stride == 2 && s.k_h == 1 && s.k_w == 1 && s.pad == 0;
}
bool conv_tc_supported_dgrad(const mmt_conv_shape& s) {
  if (s.out_c % 32 || s.in_c % 4) return false;
  if (s.stride == 1) return s.k_h - 1 - s.pad >= 0 && s.k_w - 1 - s.pad >= 0;
  return s.stride == 2 && s.k_h == 1 && s.k_w == 1 && s.pad == 0;
}
bool conv_tc_supported_wgrad(const mmt_conv_shape& s) {
  if (s.out_c % 4 || s.in_c % 4) return false;  // partial tiles: TMA zero-fills loads, clips the reduce-add
  if (s.stride == 1) return true;
  return s.stride == 2 && s.k_h == 1 && s.k_w == 1 && s.pad == 0;
}

int conv_fprop_tc(const float* x, const float* w, const float* scale, const float* bias,
                  const float* residual, float* y, const mmt_conv_shape& s, int flags, cudaStream_t stream) {
  GemmProblem g;
  const bool pointwise = s.k_h == 1 && s.k_w == 1 && s.pad == 0;
  if (pointwise && s.stride == 1) {
    const long long P = (long long)s.batch * s.in_h * s.in_w;
    g.a = flat_pixels(x, P, s.in_c);
    g.y = flat_pixels(y, P, s.out_c);
    g.r_sw = s.out_c; g.r_sh = P * s.out_c; g.r_sn = P * s.out_c;
  } else if (pointwise && s.stride == 2) {
    g.a = View4{x, s.in_c, s.out_w, s.out_h, s.batch, 2ll * s.in_c, 2ll * s.in_w * s.in_c,
                (long long)s.in_h * s.in_w * s.in_c};
    g.y = dense_nhwc(y, s.batch, s.out_h, s.out_w, s.out_c);
    g.r_sw = s.out_c; g.r_sh = (long long)s.out_w * s.out_c; g.r_sn = (long long)s.out_h * s.out_w * s.out_c;
  } else {
    g.a = dense_nhwc(x, s.batch, s.in_h, s.in_w, s.in_c);
    g.y = dense_nhwc(y, s.batch, s.out_h, s.out_w, s.out_c);
    g.r_sw = s.out_c; g.r_sh = (long long)s.out_w * s.out_c; g.r_sn = (long long)s.out_h * s.out_w * s.out_c;
  }
  g.b = w; g.n_out = s.out_c;
  g.kh = s.k_h; g.kw = s.k_w; g.pad_h = s.pad; g.pad_w = s.pad;
  g.scale = scale; g.bias = bias; g.residual = residual;
  g.relu = (flags & MMT_EPI_RELU) ? 1 : 0;
  g.prof_family = MMT_PROF_CONV_FPROP;
  return dispatch_gemm(g, stream);
}

size_t conv_stem_tc_workspace(const mmt_conv_shape& s) {
  const size_t h2 = (s.in_h + 1) / 2, w2 = (s.in_w + 1) / 2;
  return align_up(sizeof(float) * (size_t)s.batch * h2 * w2 * 32, 256) + align_up(sizeof(float) * (size_t)s.out_c * 512, 256);
}

bool conv_stem_tc_supported(const mmt_conv_shape& s) {
  return s.in_c == 3 && s.k_h == 7 && s.k_w == 7 && s.stride == 2 && s.pad == 3 && s.out_c % 4 == 0 &&
         s.out_h == (s.in_h + 1) / 2 && s.out_w == (s.in_w + 1) / 2;
}

int conv_stem_tc(const float* x, const float* w, const float* scale, const float* bias, float* y,
                 const mmt_conv_shape& s, int flags, void* workspace, size_t workspace_bytes, cudaStream_t stream) {
  if (workspace == nullptr || workspace_bytes < conv_stem_tc_workspace(s))
    return fail(MMT_ERR_WORKSPACE_TOO_SMALL, "conv_stem: workspace %zu < required %zu", workspace_bytes,
                conv_stem_tc_workspace(s));
  const int h2 = (s.in_h + 1) / 2, w2 = (s.in_w + 1) / 2;
  float* s2d = (float*)workspace;
  float* w4 = (float*)((char*)workspace + align_up(sizeof(float) * (size_t)s.batch * h2 * w2 * 32, 256));
  {
    const long long slots = (long long)s.batch * h2 * w2 * 8;
    const int grid = (int)std::min<long long>((slots + 255) / 256, (long long)kNumSMs * 16);
    stem_s2d_kernel<<<grid, 256, 0, stream>>>(x, reinterpret_cast<float4*>(s2d), s.batch, s.in_h, s.in_w, h2, w2,
                                              (flags & MMT_STEM_INPUT_NCHW) ? 1 : 0);
    if (int rc = check_launch("stem_s2d")) return rc;
    stem_filter_kernel<<<(s.out_c * 512 + 255) / 256, 256, 0, stream>>>(w, w4, s.out_c);
    if (int rc = check_launch("stem_filter")) return rc;
  }
  GemmProblem g;
  g.a = dense_nhwc(s2d, s.batch, h2, w2, 32);
  g.y = dense_nhwc(y, s.batch, s.out_h, s.out_w, s.out_c);
  g.r_sw = g.r_sh = g.r_sn = 0;
  g.b = w4; g.n_out = s.out_c;
  g.kh = 4; g.kw = 4; g.pad_h = 2; g.pad_w = 2;
  g.scale = scale; g.bias = bias; g.residual = nullptr;
  g.relu = (flags & MMT_EPI_RELU) ? 1 : 0;
  g.prof_family = MMT_PROF_CONV_FPROP;
  g.flops_override = 2.0 * s.batch * s.out_h * s.out_w * (double)s.out_c * 147.0;  // algorithmic: the 7x7x3 taps
  return dispatch_gemm(g, stream);
}

int conv_dgrad_tc(const float* dy, const float* w_t, float* dx, const mmt_conv_shape& s, cudaStream_t stream,
                  const float* aux, int aux_mode) {
  GemmProblem g;
  const bool pointwise = s.k_h == 1 && s.k_w == 1 && s.pad == 0;
  if (pointwise && s.stride == 1) {
    const long long P = (long long)s.batch * s.in_h * s.in_w;
    g.a = flat_pixels(dy, P, s.out_c);
    g.y = flat_pixels(dx, P, s.in_c);
  } else if (pointwise && s.stride == 2) {
    // dx is zero except at the even positions, which receive the GEMM result through a strided map
    MMT_CUDA(cudaMemsetAsync(dx, 0, sizeof(float) * (size_t)s.batch * s.in_h * s.in_w * s.in_c, stream));
    g.a = dense_nhwc(dy, s.batch, s.out_h, s.out_w, s.out_c);
    g.y = View4{dx, s.in_c, s.out_w, s.out_h, s.batch, 2ll * s.in_c, 2ll * s.in_w * s.in_c,
                (long long)s.in_h * s.in_w * s.in_c};
  } else {
    g.a = dense_nhwc(dy, s.batch, s.out_h, s.out_w, s.out_c);
    g.y = dense_nhwc(dx, s.batch, s.in_h, s.in_w, s.in_c);
  }
  // the epilogue's auxiliary tensor (dense, shaped like dx) is addressed through the same strides as the output view
  g.r_sw = g.y.sw; g.r_sh = g.y.sh; g.r_sn = g.y.sn;
  g.b = w_t; g.n_out = s.in_c;
  g.kh = s.k_h; g.kw = s.k_w; g.pad_h = s.k_h - 1 - s.pad; g.pad_w = s.k_w - 1 - s.pad;
  g.scale = nullptr; g.bias = nullptr; g.relu = 0;
  g.residual = aux_mode ? aux : nullptr;
  g.res_mode = aux_mode == MMT_DGRAD_AUX_MASK ? 1 : 0;
  g.prof_family = MMT_PROF_CONV_DGRAD;
  return dispatch_gemm(g, stream);
}

int conv_filter_for_dgrad(const float* w, float* w_t, const mmt_conv_shape& s, cudaStream_t stream, const float* k_scale) {
  const int taps = s.k_h * s.k_w;
  dim3 grid((unsigned)ceil_div(s.in_c, 32), (unsigned)ceil_div(s.out_c, 32), (unsigned)taps);
  filter_dgrad_kernel<<<grid, 256, 0, stream>>>(w, w_t, s.out_c, taps, s.in_c, k_scale);
  return check_launch("filter_dgrad");
}

int conv_wgrad_tc(const float* x, const float* dy, float* dw, float* dbias, const mmt_conv_shape& s,
                  void* workspace, size_t workspace_bytes, cudaStream_t stream, bool accumulate, const float* row_scale) {
  (void)workspace; (void)workspace_bytes; (void)dbias;
  const bool pointwise = s.k_h == 1 && s.k_w == 1 && s.pad == 0;
  View4 vdy, vx;
  if (pointwise && s.stride == 1) {
    const long long P = (long long)s.batch * s.in_h * s.in_w;
    vdy = flat_pixels(dy, P, s.out_c);
    vx = flat_pixels(x, P, s.in_c);
  } else if (pointwise && s.stride == 2) {
    vdy = dense_nhwc(dy, s.batch, s.out_h, s.out_w, s.out_c);
    vx = View4{x, s.in_c, s.out_w, s.out_h, s.batch, 2ll * s.in_c, 2ll * s.in_w * s.in_c,
               (long long)s.in_h * s.in_w * s.in_c};
  } else {
    vdy = dense_nhwc(dy, s.batch, s.out_h, s.out_w, s.out_c);
    vx = dense_nhwc(x, s.batch, s.in_h, s.in_w, s.in_c);
  }
  const int BN = s.in_c % 256 == 0 ? 256 : (s.in_c % 128 == 0 ? 128 : 64);  // ci tiles may be partial
  int TW, TH;
  pick_patch(vdy.w, vdy.h, 32, TW, TH);
  CUtensorMap tmDY, tmX, tmDW;
  // MN-major TF32 operands only exist in the "128B swizzle, 32-byte atom" layout (4-row period)
  if (int rc = make_map(&tmDY, vdy, 32, TW, TH, "wgrad dy", CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B)) return rc;
  if (int rc = make_map(&tmX, vx, 32, TW, TH, "wgrad x", CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B)) return rc;
  const long long Ktot = (long long)s.k_h * s.k_w * s.in_c;
  if (int rc = make_map_2d(&tmDW, dw, Ktot, s.out_c, Ktot, 32, 128, "wgrad dw")) return rc;
  // every partial tile is ADDED into dw by TMA reduce: zero it first, unless the caller accumulates (dw += ...)
  if (!accumulate) MMT_CUDA(cudaMemsetAsync(dw, 0, sizeof(float) * (size_t)s.out_c * Ktot, stream));
  WgradParams p;
  p.tiles_w = (int)ceil_div<long long>(vdy.w, TW);
  p.tiles_h = (int)ceil_div<long long>(vdy.h, TH);
  p.tiles_b = (int)vdy.n;
  p.TW = TW; p.TH = TH;
  p.kh = s.k_h; p.kw = s.k_w; p.pad_h = s.pad; p.pad_w = s.pad;
  p.co_tiles = (s.out_c + 127) / 128;
  p.ci_tiles = (s.in_c + BN - 1) / BN;
  p.cin = s.in_c;
  p.cout = s.out_c;
  p.row_scale = row_scale;
  const long long ptiles = (long long)p.tiles_w * p.tiles_h * p.tiles_b;
  if (ptiles > 0x7fffffffll) return fail(MMT_ERR_UNSUPPORTED, "wgrad: too many pixel tiles");
  p.pixel_tiles = (int)ptiles;
  const int base_items = p.co_tiles * s.k_h * s.k_w * p.ci_tiles;
  int splits = std::max(1, (2 * kNumSMs) / base_items);
  splits = (int)std::min<long long>(splits, std::max<long long>(1, ptiles / 16));  // >= 16 K-steps per item
  p.tiles_per_split = (int)ceil_div<long long>(ptiles, splits);
  p.splits = (int)ceil_div<long long>(ptiles, p.tiles_per_split);
  p.total_items = base_items * p.splits;
  const double pix = (double)vdy.n * vdy.h * vdy.w;
  mmt::ProfScope prof(MMT_PROF_CONV_WGRAD, 2.0 * pix * s.out_c * (double)Ktot,
                      4.0 * (pix * (s.out_c + s.in_c) + (double)s.out_c * Ktot), stream);
  if (BN == 256) return launch_wgrad<256>(tmDY, tmX, tmDW, p, stream);
  if (BN == 128) return launch_wgrad<128>(tmDY, tmX, tmDW, p, stream);
  return launch_wgrad<64>(tmDY, tmX, tmDW, p, stream);
}

}  // namespace mmt
