// K6, tensor-core mode: TMA-fed tf32 GEMM / implicit-GEMM convolution on tcgen05 (fp32 operands straight from HBM,
// no register staging, no conversion pass, no transposed copies).
//
//   producer (1 thread)  cp.async.bulk.tensor (tiled 2-D, or 4-D im2col for convolutions), 128-byte swizzle
//                        -> 3-stage shared-memory ring, completion by mbarrier transaction bytes
//   MMA      (1 thread)  tcgen05.mma.cta_group::1.kind::tf32  (M = 128, N = BN, K = 8 per instruction),
//                        accumulator in TMEM; tcgen05.commit frees the stage / signals the epilogue
//   epilogue (4 warps)   tcgen05.ld -> per-warp slab in the drained stages -> row-wise coalesced stores with bias /
//                        activation / activation-derivative mask (or split-K partials, or a transposed store)
// An operand whose memory-contiguous direction is the reduction index is loaded K-major (box = 32 k x rows);
// otherwise MN-major (boxes of 32 rows/columns x 32 k; same 128-byte swizzled rows in shared memory, but the 32-byte
// atom flavour of the swizzle, the instruction descriptor's major bit and the LBO/SBO strides differ), so the
// data-gradient and weight-gradient products need no transposed copies.
// Entry points (each returns 1 when the shapes break TMA's rules; layers_api.cu then falls back to gemm_tc.cu):
//   tma_linear_fwd / dgrad / wgrad          dense layers
//   tma_conv_fwd / wgrad / dgrad            NHWC fp32 activations with C % 32 == 0 (dgrad: one sub-problem per stride phase)
//   tma_conv_fwd_u8 / wgrad_u8              first layer on uint8 frames (padded fp32 row image + overlapping wide pixels)
#include <algorithm>

#include <cuda.h>

#include "common.cuh"
#include "gemm_common.cuh"
#include "tc_common.cuh"
#include "tma_common.cuh"

namespace b200rl {

__device__ unsigned long long* g_tma_timeline = nullptr;   // tools/tc_timeline.py
__device__ __forceinline__ void tm_mark(int slot) {
  if (g_tma_timeline && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    g_tma_timeline[slot] = t;
  }
}

constexpr int GBM = 128, GBK = 32 /* fp32 elements = 128 bytes */, G_STAGES = 3, G_THREADS = 192;
int g_tma_dgrad_bn64 = 1;   // tools/layer_bench.py toggles it (B200RL_DGRAD_BN64)
int g_tma_stage_mode = 0;   // 0 = heuristic, 1 = always shallow, 2 = always deep (tools/layer_bench.py)

// NHWC fp32 activation tensor viewed through TMA's im2col mode: one instruction loads `pixels` consecutive
// output pixels (b, oy, ox order) x `channels` input channels of ONE filter tap; padding reads as zero.
// Corners as in CUTLASS (conv/collective/detail.hpp): lower = -pad_before, upper = pad_after - (k - 1).
static bool make_im2col_map(CUtensorMap* m, const float* x, const b200rl_conv_geom& g, int channels, int pixels,
                            bool mn_major) {
  EncodeIm2colFn enc = get_encode_im2col();
  if (!enc) return false;
  const int pad_bottom = (g.OH - 1) * g.stride + g.kh - g.H - g.pad_top;
  const int pad_right = (g.OW - 1) * g.stride + g.kw - g.W - g.pad_left;
  cuuint64_t dims[4] = {(cuuint64_t)g.C, (cuuint64_t)g.W, (cuuint64_t)g.H, (cuuint64_t)g.B};
  cuuint64_t strides[3] = {(cuuint64_t)g.C * 4, (cuuint64_t)g.W * g.C * 4, (cuuint64_t)g.H * g.W * g.C * 4};
  int lower[2] = {-g.pad_left, -g.pad_top};
  int upper[2] = {pad_right - (g.kw - 1), pad_bottom - (g.kh - 1)};
  cuuint32_t es[4] = {1, (cuuint32_t)g.stride, (cuuint32_t)g.stride, 1};
  return enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, (void*)x, dims, strides, lower, upper, (cuuint32_t)channels,
             (cuuint32_t)pixels, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
             mn_major ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B,
             CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// fp32 matrix X[lines][pos] (pos contiguous, row stride ld floats); box = box_pos x box_lines
static bool make_map(CUtensorMap* m, const float* p, int64_t lines, int64_t pos, int64_t ld, int box_pos, int box_lines,
                     bool mn_major) {
  EncodeTiledFn enc = get_encode();
  if (!enc) return false;
  cuuint64_t dims[2] = {(cuuint64_t)pos, (cuuint64_t)lines};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 4};
  cuuint32_t box[2] = {(cuuint32_t)box_pos, (cuuint32_t)box_lines};
  cuuint32_t es[2] = {1, 1};
  return enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)p, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
             // MN-major tf32 operands only exist in the 32-byte-atom flavour of the 128-byte swizzle
             mn_major ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}
static bool tma_ok(const float* p, int64_t ld) { return (((uintptr_t)p) & 15) == 0 && (ld * 4) % 16 == 0; }

__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// 128-byte-swizzled tile: start address, LBO, SBO (bytes), version 1, layout type 2 (SWIZZLE_128B, 16-byte
// atoms: K-major operands) or 1 (SWIZZLE_128B_BASE32B, 32-byte atoms: the only layout for MN-major tf32)
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint64_t layout) {
  return (uint64_t)((saddr >> 4) & 0x3FFFu) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16) |
         ((uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32) | (1ull << 46) | (layout << 61);
}
// kind::tf32: D = f32 (bits 4-5 = 1), A = B = tf32 (format 2 at bits 7-9 / 10-12), major bits 15 / 16
__host__ __device__ constexpr uint32_t umma_idesc_tf32(int M, int N, bool a_mn, bool b_mn) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((a_mn ? 1u : 0u) << 15) | ((b_mn ? 1u : 0u) << 16) |
         ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// C[row, col] = sum_k A(row, k) B(k, col).  Tile coordinates for the maps:
//   K-major operand : one box {32 k, R rows}       at (k0, row0)
//   MN-major operand: R/32 boxes {32 rows, 32 k}   at (row0 + 32 j, k0), stacked 4 KB apart
template <bool AMN, bool BMN, int BN, int STAGES>
__global__ void __launch_bounds__(G_THREADS)
tma_gemm_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b, Epilogue epi,
                int M, int N, int K, int kblocks_per_split, ConvA conv) {
  constexpr int A_BYTES = GBM * 128, B_BYTES = BN * 128, STAGE = A_BYTES + B_BYTES;
  constexpr uint32_t TMEM_COLS = BN < 32 ? 32 : BN;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;   // 128-byte swizzle atoms are 1 KB
  __shared__ __align__(8) uint64_t bar_full[STAGES], bar_empty[STAGES], bar_done;
  __shared__ uint32_t tmem_base_smem;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (!epi.pdl_late) pdl_launch_dependents();
  const int row0 = blockIdx.y * GBM, col0 = blockIdx.x * BN;
  const int total_kblocks = (K + GBK - 1) / GBK;
  const int kb_begin = blockIdx.z * kblocks_per_split;
  const int nkb = min(total_kblocks, kb_begin + kblocks_per_split) - kb_begin;

  if (tid == 0) {
    prefetch_tensormap(&map_a);   // the descriptor fetch otherwise sits in front of the first load
    prefetch_tensormap(&map_b);
#pragma unroll
    for (int s = 0; s < STAGES; ++s) { mbar_init(&bar_full[s], 1); mbar_init(&bar_empty[s], 1); }
    mbar_init(&bar_done, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) tmem_alloc(&tmem_base_smem, TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  pdl_wait();                  // the set-up above overlaps the previous kernel's tail (launch_pdl); no-op otherwise
  const uint32_t tmem_d = tmem_base_smem;
  if (tid == 0) tm_mark(0);

  if (warp == 0 && lane == 0) {
    // ---------------- TMA producer
    int ax = 0, ay = 0, an = 0;   // im2col anchor of the tile's first pixel (input coordinates)
    if (conv.enabled) {
      const int ox = row0 % conv.OW, t = row0 / conv.OW;
      ax = ox * conv.stride_w - conv.pad_left;
      ay = (t % conv.OH) * conv.stride_h - conv.pad_top;
      an = t / conv.OH;
    }
    // all per-k-block coordinates advance incrementally: the producer is a single thread and run-time integer
    // divisions (dozens of instructions each) would bound the rate at which it can issue loads
    int f_c0 = 0, f_tx = 0, f_ty = 0;          // conv forward: channel block / filter tap of the current k-block
    int w_ox = 0, w_oy = 0, w_n = 0;           // conv weight gradient: first pixel of the current k-block
    int w_c[GBM / 32], w_tx[GBM / 32], w_ty[GBM / 32], w_nblk = 0;   // ... and its (loop-invariant) row blocks
    if (conv.enabled && !AMN) {
      const int cb = conv.C / GBK, tap = kb_begin / cb;
      f_c0 = (kb_begin % cb) * GBK; f_tx = tap % conv.kw; f_ty = tap / conv.kw;
    }
    if (conv.enabled && AMN) {
      const int p0 = kb_begin * GBK, t = p0 / conv.OW;
      w_ox = p0 % conv.OW; w_oy = t % conv.OH; w_n = t / conv.OH;
#pragma unroll
      for (int j = 0; j < GBM / 32; ++j) {
        const int r = row0 + 32 * j, tap = r / conv.C;
        w_c[j] = r % conv.C; w_tx[j] = tap % conv.kw; w_ty[j] = tap / conv.kw;
        if (tap < conv.taps) w_nblk = j + 1;   // taps are ascending in j: blocks past the last tap are skipped
      }
    }
    for (int i = 0; i < nkb; ++i) {
      const int s = i % STAGES;
      if (i >= STAGES) mbar_wait(&bar_empty[s], ((i / STAGES) - 1) & 1);
      const uint32_t sa = base + s * STAGE, sb = sa + A_BYTES;
      const int k0 = (kb_begin + i) * GBK;
      if (AMN && conv.enabled) {
        // conv weight gradient: rows = patch indices (tap, channel), reduction = pixels.  Row block j of the tile
        // is 32 channels of one tap for the 32 pixels of this k-block; blocks past the last tap are skipped
        // (their accumulator rows are never stored).
        const int px = w_ox * conv.stride_w - conv.pad_left, py = w_oy * conv.stride_h - conv.pad_top;
        mbar_expect_tx(&bar_full[s], w_nblk * 4096 + B_BYTES);
#pragma unroll
        for (int j = 0; j < GBM / 32; ++j)
          if (j < w_nblk)
            tma_load_im2col(sa + j * 4096, &map_a, w_c[j], px, py, w_n, (uint16_t)w_tx[j], (uint16_t)w_ty[j], &bar_full[s]);
        w_ox += GBK;
        while (w_ox >= conv.OW) {
          w_ox -= conv.OW;
          if (++w_oy == conv.OH) { w_oy = 0; ++w_n; }
        }
      } else {
        mbar_expect_tx(&bar_full[s], STAGE);
      }
      if (AMN && conv.enabled) {
      } else if (!AMN && conv.enabled) {   // conv forward: 128 pixels x 32 channels of one tap
        tma_load_im2col(sa, &map_a, f_c0, ax, ay, an, (uint16_t)f_tx, (uint16_t)f_ty, &bar_full[s]);
        f_c0 += GBK;
        if (f_c0 == conv.C) {
          f_c0 = 0;
          if (++f_tx == conv.kw) { f_tx = 0; ++f_ty; }
        }
      } else if (AMN) {
#pragma unroll
        for (int j = 0; j < GBM / 32; ++j) tma_load_2d(sa + j * 4096, &map_a, row0 + 32 * j, k0, &bar_full[s]);
      } else {
        tma_load_2d(sa, &map_a, k0, row0, &bar_full[s]);
      }
      if (BMN) {
#pragma unroll
        for (int j = 0; j < BN / 32; ++j) tma_load_2d(sb + j * 4096, &map_b, col0 + 32 * j, k0, &bar_full[s]);
      } else {
        tma_load_2d(sb, &map_b, k0, col0, &bar_full[s]);
      }
    }
  } else if (warp == 1 && lane == 0) {
    // ---------------- MMA issuer
    constexpr uint32_t idesc = umma_idesc_tf32(GBM, BN, AMN, BMN);
    for (int i = 0; i < nkb; ++i) {
      const int s = i % STAGES;
      mbar_wait(&bar_full[s], (i / STAGES) & 1);
      if (i < 8) tm_mark(1 + i);
      tc_fence_after();
      const uint32_t sa = base + s * STAGE, sb = sa + A_BYTES;
#pragma unroll
      for (int kk = 0; kk < GBK / 8; ++kk) {
        // K-major : 8 k = 32 bytes inside the swizzled 128-byte row; 8-row groups 1 KB apart (SBO)
        // MN-major: 8 k = 8 rows of 128 bytes = two 4-row swizzle atoms 512 B apart (SBO); 32-wide row/column
        //           blocks 4 KB apart (LBO)
        const uint64_t da = AMN ? umma_desc_sw128(sa + kk * 1024, 4096, 512, 1) : umma_desc_sw128(sa + kk * 32, 16, 1024, 2);
        const uint64_t db = BMN ? umma_desc_sw128(sb + kk * 1024, 4096, 512, 1) : umma_desc_sw128(sb + kk * 32, 16, 1024, 2);
        umma_tf32(tmem_d, da, db, idesc, (i > 0 || kk > 0) ? 1u : 0u);
      }
      umma_commit(&bar_empty[s]);
      if (i == nkb - 1) umma_commit(&bar_done);
    }
  } else if (warp >= 2) {
    // ---------------- epilogue: warp w may read TMEM lanes 32 * (w % 4) ..
    if (nkb > 0) {
      mbar_wait(&bar_done, 0);
      tc_fence_after();
    }
    if (epi.pdl_late) pdl_launch_dependents();
    if (tid == 64) tm_mark(10);
    const int lane_base = (warp & 3) * 32;
    if (epi.transpose_out && !epi.partial) {
      // C^T is wanted: lanes (= rows) are already the contiguous direction of the destination
      const int row = row0 + lane_base + lane;
#pragma unroll 1
      for (int cc = 0; cc < BN; cc += 16) {
        float v[16];
        if (nkb > 0) tmem_ld16(tmem_d + ((uint32_t)lane_base << 16) + (uint32_t)cc, v);
        else {
#pragma unroll
          for (int j = 0; j < 16; ++j) v[j] = 0.f;
        }
        if (row < M) finish16(epi, row, col0 + cc, N, M, v);
      }
    } else {
      // transpose through shared memory (the drained pipeline stages) so that every store instruction of a warp
      // covers whole rows: 32 / (BN / 4) rows x BN floats
      float* slab = reinterpret_cast<float*>(smem_raw + (base - smem_u32(smem_raw))) + (warp & 3) * 32 * BN;
      stage_accumulator<BN>(tmem_d + ((uint32_t)lane_base << 16), slab, lane, nkb > 0);
      const int first = row0 + lane_base;
      store_staged_rows<BN>(epi, slab, lane, col0, N, M, [&](int r) -> long long { return first + r < M ? first + r : -1; });
    }
  }
  if (tid == 64) tm_mark(11);
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_d, TMEM_COLS);
  if (tid == 32) tm_mark(12);
}

template <bool AMN, bool BMN, int BN>
static int launch_tma_bn(const CUtensorMap& ma, const CUtensorMap& mb, Epilogue epi, int M, int N, int K, void* ws,
                         int64_t ws_bytes, cudaStream_t stream, const ConvA& conv) {
  const int tiles = ceil_div(M, GBM) * ceil_div(N, BN);
  const int kblocks = ceil_div(K, GBK);
  int splits = 1;
  if (2 * tiles <= kNumSMs && kblocks >= 16) {   // a partial pass + finish launch costs more than a part-filled wave
    splits = std::min(ceil_div(kNumSMs, tiles), kblocks / 8);
    splits = std::min(splits, 64);
    const int64_t cap = ws ? ws_bytes / ((int64_t)M * N * 4) : 0;
    splits = (int)std::max<int64_t>(1, std::min<int64_t>(splits, cap));
  }
  const int kps = ceil_div(kblocks, splits);
  splits = ceil_div(kblocks, kps);
  epi.partial = splits > 1 ? (float*)ws : nullptr;
  dim3 grid(ceil_div(N, BN), ceil_div(M, GBM), splits);
  // one CTA per SM anyway -> spend the shared memory on a deeper TMA ring (the loads are latency-bound);
  // more CTAs than SMs -> keep stages shallow so 2-3 CTAs co-reside and one's epilogue overlaps another's loads
  const bool deep = g_tma_stage_mode == 2 || (false);
  if (deep) {
    constexpr int DS = BN == 128 ? 6 : 8;
    constexpr int smem = DS * (GBM * 128 + BN * 128) + 1024;
    static bool attr = false;
    if (!attr) {
      B200RL_CUDA_OK(cudaFuncSetAttribute(tma_gemm_kernel<AMN, BMN, BN, DS>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
      attr = true;
    }
    epi.pdl_late = pdl_late_mode();
    B200RL_CUDA_OK(launch_pdl(tma_gemm_kernel<AMN, BMN, BN, DS>, dim3(grid), dim3(G_THREADS), smem, stream, ma, mb, epi, M, N, K, kps, conv));
  } else {
    constexpr int smem = G_STAGES * (GBM * 128 + BN * 128) + 1024;
    static bool attr = false;
    if (!attr) {
      B200RL_CUDA_OK(cudaFuncSetAttribute(tma_gemm_kernel<AMN, BMN, BN, G_STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
      attr = true;
    }
    epi.pdl_late = pdl_late_mode();
    B200RL_CUDA_OK(launch_pdl(tma_gemm_kernel<AMN, BMN, BN, G_STAGES>, dim3(grid), dim3(G_THREADS), smem, stream, ma, mb, epi, M, N, K, kps, conv));
  }
  B200RL_LAUNCH_OK();
  if (splits > 1) return launch_splitk_finish(epi, M, N, splits, stream);
  return B200RL_OK;
}

template <bool AMN, bool BMN>
static int launch_tma(const CUtensorMap& ma, const CUtensorMap& mb, const Epilogue& epi, int M, int N, int K, int BN,
                      void* ws, int64_t wsb, cudaStream_t s, const ConvA& conv = ConvA{0, 0, 0, 0, 0, 0, 0, 0, 0, 0}) {
  if (BN == 32) return launch_tma_bn<AMN, BMN, 32>(ma, mb, epi, M, N, K, ws, wsb, s, conv);
  if (BN == 64) return launch_tma_bn<AMN, BMN, 64>(ma, mb, epi, M, N, K, ws, wsb, s, conv);
  return launch_tma_bn<AMN, BMN, 128>(ma, mb, epi, M, N, K, ws, wsb, s, conv);
}
static int pick_bn(int N) { return N <= 32 ? 32 : (N <= 64 ? 64 : 128); }

// ---- entry points; return 1 when the shapes do not satisfy TMA's rules (caller falls back to gemm_tc.cu)
int tma_linear_fwd(int M, int N, int K, const float* x, int ldx, const float* w, const float* bias, float* y,
                   int ldy, int act, void* ws, int64_t wsb, cudaStream_t s) {
  if (!tma_ok(x, ldx) || !tma_ok(w, K) || K < GBK) return 1;
  const int BN = pick_bn(N);
  CUtensorMap ma, mb;
  if (!make_map(&ma, x, M, K, ldx, GBK, GBM, false) || !make_map(&mb, w, N, K, K, GBK, BN, false)) return 1;
  Epilogue e{y, ldy, bias, act, nullptr, 0, 0, nullptr, 0};
  return launch_tma<false, false>(ma, mb, e, M, N, K, BN, ws, wsb, s);
}
int tma_linear_dgrad(int M, int N, int K, const float* dy, int lddy, const float* w, float* dx, int lddx,
                     const float* mask, int ldmask, int mask_act, void* ws, int64_t wsb, cudaStream_t s) {
  // dx[m, k] = sum_n dy[m, n] w[n, k]: A = dy K-major (reduction n contiguous), B = w MN-major (line = n, pos = k)
  if (!tma_ok(dy, lddy) || !tma_ok(w, K) || N < GBK) return 1;
  int BN = pick_bn(K);
  // no split-K here (its finish pass would re-read the whole output): when 128-wide tiles leave SMs idle, halve them
  // so that two CTAs per SM overlap each other's load latency and epilogue
  if (BN == 128 && ceil_div(M, GBM) * ceil_div(K, 128) < kNumSMs && g_tma_dgrad_bn64) BN = 64;
  CUtensorMap ma, mb;
  if (!make_map(&ma, dy, M, N, lddy, GBK, GBM, false) || !make_map(&mb, w, N, K, K, 32, GBK, true)) return 1;
  Epilogue e{dx, lddx, nullptr, 0, mask, ldmask, mask_act, nullptr, 0};
  return launch_tma<false, true>(ma, mb, e, M, K, N, BN, ws, wsb, s);
}
int tma_linear_wgrad(int M, int N, int K, const float* dy, int lddy, const float* x, int ldx, float* dw, float* db,
                     void* ws, int64_t wsb, cudaStream_t s) {
  // dw[n, k] = sum_m dy[m, n] x[m, k]: both operands MN-major (line = reduction m)
  if (!tma_ok(dy, lddy) || !tma_ok(x, ldx) || M < GBK) return 1;
  const int BN = pick_bn(K);
  CUtensorMap ma, mb;
  if (!make_map(&ma, dy, M, N, lddy, 32, GBK, true) || !make_map(&mb, x, M, K, ldx, 32, GBK, true)) return 1;
  Epilogue e{dw, K, nullptr, 0, nullptr, 0, 0, nullptr, 0};
  int rc = launch_tma<true, true>(ma, mb, e, N, K, M, BN, ws, wsb, s);
  if (rc) return rc;
  if (db) return launch_colsum(M, N, dy, lddy, db, ws, wsb, s);
  return B200RL_OK;
}

// conv forward with an fp32 NHWC input whose channel count is a multiple of 32: A = TMA im2col, B = weights
int tma_conv_fwd(const float* x, const float* w, const float* bias, float* y, const b200rl_conv_geom& g, int act,
                 void* ws, int64_t wsb, cudaStream_t s) {
  const int M = g.B * g.OH * g.OW, K = g.kh * g.kw * g.C;
  if (g.C % GBK != 0 || !tma_ok(x, g.C) || !tma_ok(w, K) || (int64_t)g.B * g.H * g.W * g.C * 4 < 131072) return 1;
  const int BN = pick_bn(g.Cout);
  CUtensorMap ma, mb;
  if (!make_im2col_map(&ma, x, g, GBK, GBM, false) || !make_map(&mb, w, g.Cout, K, K, GBK, BN, false)) return 1;
  Epilogue e{y, g.Cout, bias, act, nullptr, 0, 0, nullptr, 0};
  ConvA conv{1, g.C, g.kw, g.OW, g.OH, g.stride, g.stride, g.pad_left, g.pad_top, g.kh * g.kw};
  return launch_tma<false, false>(ma, mb, e, M, g.Cout, K, BN, ws, wsb, s, conv);
}

// conv weight gradient, fp32 NHWC input with C % 32 == 0: dW^T[k, co] = sum_pixel col[pixel, k] dy[pixel, co]
int tma_conv_wgrad(const float* x, const float* dy, float* dw, float* db, const b200rl_conv_geom& g, void* ws, int64_t wsb,
                   cudaStream_t s) {
  const int M = g.B * g.OH * g.OW, K = g.kh * g.kw * g.C;
  if (g.C % GBK != 0 || !tma_ok(x, g.C) || !tma_ok(dy, g.Cout) || (int64_t)g.B * g.H * g.W * g.C * 4 < 131072 || M < GBK) return 1;
  const int BN = pick_bn(g.Cout);
  CUtensorMap ma, mb;
  if (!make_im2col_map(&ma, x, g, 32, GBK, true) || !make_map(&mb, dy, M, g.Cout, g.Cout, 32, GBK, true)) return 1;
  Epilogue e{dw, K, nullptr, 0, nullptr, 0, 0, nullptr, 1};   // stored transposed: dw[co][k]
  ConvA conv{1, g.C, g.kw, g.OW, g.OH, g.stride, g.stride, g.pad_left, g.pad_top, g.kh * g.kw};
  int rc = launch_tma<true, true>(ma, mb, e, K, g.Cout, M, BN, ws, wsb, s, conv);
  if (rc) return rc;
  if (db) return launch_colsum(M, g.Cout, dy, g.Cout, db, ws, wsb, s);
  return B200RL_OK;
}

// ---- first-layer convolutions on uint8 frames (DQN conv1: 84x84x4, 8x8 stride 4).  Four channels are far too
// few for TMA's im2col mode, but one filter ROW (kw * C = 32 values) is contiguous in an NHWC image.  The frames are
// therefore rewritten once per call as zero-padded fp32 rows, float(x) / 255 (`atari_wrapper.py:303-304`), and the
// convolution becomes a kh x 1 filter over "wide pixels" of 32 channels: wide pixel ox of an image row starts at
// float offset ox * stride * C, i.e. neighbouring wide pixels OVERLAP in memory (TMA strides need not be >= the
// extent of the inner dimension).  If the driver refuses the overlapping view the rows are materialised without
// overlap instead ([.., OW, 32], twice the bytes).
constexpr int kRowsPerCta = 8;   // one CTA per padded row made the CTA launch rate (152 tiny CTAs per SM) the bottleneck
__global__ void __launch_bounds__(256)
u8_rows_to_f32_kernel(const uint8_t* __restrict__ x, float* __restrict__ out, int H, int W, int Hp, int per_row /* float4s */,
                      int wq /* float4s per wide pixel */, int src_step /* source pixels between wide pixels */, int pad_left,
                      int pad_top, int total_rows) {
  // a CTA converts kRowsPerCta consecutive padded rows (b, hp); one float4 = the 4 channels of one pixel
  const int row0 = blockIdx.x * kRowsPerCta;
  const int nrows = min(kRowsPerCta, total_rows - row0);
  for (int e = threadIdx.x; e < nrows * per_row; e += blockDim.x) {
    const int r = e / per_row, q = e - r * per_row;
    const int row = row0 + r, b = row / Hp, hp = row - b * Hp, ys = hp - pad_top;
    const int xs = wq == 1 ? q - pad_left : (q / wq) * src_step + (q % wq) - pad_left;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (ys >= 0 && ys < H && xs >= 0 && xs < W) {
      const uchar4 p = __ldg(reinterpret_cast<const uchar4*>(x) + ((size_t)b * H + ys) * W + xs);
      v = make_float4(__fdiv_rn((float)p.x, 255.f), __fdiv_rn((float)p.y, 255.f), __fdiv_rn((float)p.z, 255.f), __fdiv_rn((float)p.w, 255.f));
    }
    reinterpret_cast<float4*>(out)[(size_t)row * per_row + q] = v;
  }
}

struct WideView {
  int Hp, row_floats, wide_stride;   // padded rows, floats per row, floats between wide pixels
  int64_t bytes;
};
static int g_wide_overlap = 1;   // 1 = try the overlapping view first, 0 = driver refused it once
static bool wide_eligible(const b200rl_conv_geom& g) {
  return g.C == 4 && g.kw * g.C == GBK && g.Cout % 8 == 0 && g.Cout <= 128 && (g.stride * g.C * 4) % 16 == 0;
}
static WideView wide_view(const b200rl_conv_geom& g, bool overlap) {
  WideView v;
  const int pad_bottom = (g.OH - 1) * g.stride + g.kh - g.H - g.pad_top;
  const int pad_right = (g.OW - 1) * g.stride + g.kw - g.W - g.pad_left;
  v.Hp = g.H + g.pad_top + std::max(pad_bottom, 0);
  if (overlap) { v.row_floats = (g.W + g.pad_left + std::max(pad_right, 0)) * g.C; v.wide_stride = g.stride * g.C; }
  else { v.row_floats = g.OW * GBK; v.wide_stride = GBK; }
  v.bytes = (int64_t)g.B * v.Hp * v.row_floats * 4;
  return v;
}
static int wide_convert(const uint8_t* x, float* out, const b200rl_conv_geom& g, const WideView& v, bool overlap, cudaStream_t s) {
  // overlapping view: the row is the padded image itself, "wide pixel" granularity = one source pixel
  const int per_row = v.row_floats / 4, wq = overlap ? 1 : v.wide_stride / 4, step = overlap ? 1 : g.stride;
  const int rows = g.B * v.Hp;
  u8_rows_to_f32_kernel<<<ceil_div(rows, kRowsPerCta), 256, 0, s>>>(x, out, g.H, g.W, v.Hp, per_row, wq, step, g.pad_left,
                                                                    g.pad_top, rows);
  B200RL_LAUNCH_OK();
  return B200RL_OK;
}
static bool make_wide_map(CUtensorMap* m, const float* xw, const b200rl_conv_geom& g, const WideView& v, int pixels, bool mn_major) {
  EncodeIm2colFn enc = get_encode_im2col();
  if (!enc) return false;
  cuuint64_t dims[4] = {(cuuint64_t)GBK, (cuuint64_t)g.OW, (cuuint64_t)v.Hp, (cuuint64_t)g.B};
  cuuint64_t strides[3] = {(cuuint64_t)v.wide_stride * 4, (cuuint64_t)v.row_floats * 4, (cuuint64_t)v.Hp * v.row_floats * 4};
  int lower[2] = {0, 0};
  int upper[2] = {0, -(g.kh - 1)};
  cuuint32_t es[4] = {1, 1, (cuuint32_t)g.stride, 1};
  return enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, (void*)xw, dims, strides, lower, upper, GBK, (cuuint32_t)pixels, es,
             CU_TENSOR_MAP_INTERLEAVE_NONE, mn_major ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B,
             CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}
// Size of the row image of `g` (0: geometry not eligible).  The larger (non-overlapping) layout is reported so that a
// buffer of this size works whichever view the driver accepts.
int64_t tma_rows_bytes(const b200rl_conv_geom& g) {
  if (!wide_eligible(g)) return 0;
  return std::max(wide_view(g, true).bytes, wide_view(g, false).bytes);
}
// Picks the view (overlapping first; remembered process-wide) and converts the frames into `rows`.
int tma_rows_from_u8(const uint8_t* x, const b200rl_conv_geom& g, float* rows, int64_t bytes, cudaStream_t s) {
  if (!wide_eligible(g) || !rows || (((uintptr_t)rows) & 127) != 0) return 1;
  for (int attempt = 0; attempt < 2; ++attempt) {
    const bool overlap = g_wide_overlap != 0;
    const WideView v = wide_view(g, overlap);
    if (v.bytes > bytes) return 1;
    CUtensorMap probe;
    if (make_wide_map(&probe, rows, g, v, GBM, false)) return wide_convert(x, rows, g, v, overlap, s);
    if (!overlap) return 1;
    g_wide_overlap = 0;
  }
  return 1;
}
static int64_t rows_used(const b200rl_conv_geom& g) {
  return (wide_view(g, g_wide_overlap != 0).bytes + 1023) & ~(int64_t)1023;
}

// first-layer convolution / weight gradient on a row image built by tma_rows_from_u8 (same geometry)
int tma_conv_fwd_rows(const float* rows, const float* w, const float* bias, float* y, const b200rl_conv_geom& g, int act,
                      void* ws, int64_t wsb, cudaStream_t s) {
  const int M = g.B * g.OH * g.OW, K = g.kh * g.kw * g.C;
  if (!wide_eligible(g) || !tma_ok(w, K)) return 1;
  CUtensorMap ma, mb;
  const int BN = pick_bn(g.Cout);
  if (!make_map(&mb, w, g.Cout, K, K, GBK, BN, false)) return 1;
  if (!make_wide_map(&ma, rows, g, wide_view(g, g_wide_overlap != 0), GBM, false)) return 1;
  Epilogue e{y, g.Cout, bias, act, nullptr, 0, 0, nullptr, 0};
  ConvA conv{1, GBK, 1, g.OW, g.OH, 1, g.stride, 0, 0, g.kh};
  return launch_tma<false, false>(ma, mb, e, M, g.Cout, K, BN, ws, wsb, s, conv);
}
int tma_conv_wgrad_rows(const float* rows, const float* dy, float* dw, float* db, const b200rl_conv_geom& g, void* ws,
                        int64_t wsb, cudaStream_t s) {
  const int M = g.B * g.OH * g.OW, K = g.kh * g.kw * g.C;
  if (!wide_eligible(g) || !tma_ok(dy, g.Cout) || g.Cout % 32 != 0 || M < GBK) return 1;
  CUtensorMap ma, mb;
  const int BN = pick_bn(g.Cout);
  if (!make_map(&mb, dy, M, g.Cout, g.Cout, 32, GBK, true)) return 1;
  if (!make_wide_map(&ma, rows, g, wide_view(g, g_wide_overlap != 0), GBK, true)) return 1;
  Epilogue e{dw, K, nullptr, 0, nullptr, 0, 0, nullptr, 1};
  ConvA conv{1, GBK, 1, g.OW, g.OH, 1, g.stride, 0, 0, g.kh};
  int rc = launch_tma<true, true>(ma, mb, e, K, g.Cout, M, BN, ws, wsb, s, conv);
  if (rc) return rc;
  if (db) return launch_colsum(M, g.Cout, dy, g.Cout, db, ws, wsb, s);
  return B200RL_OK;
}

// uint8 frames: the row image goes to the head of the workspace, then as above
int tma_conv_fwd_u8(const uint8_t* x, const float* w, const float* bias, float* y, const b200rl_conv_geom& g, int act,
                    void* ws, int64_t wsb, cudaStream_t s) {
  const int M = g.B * g.OH * g.OW, K = g.kh * g.kw * g.C;
  if (!tma_ok(w, K) || (int64_t)M * g.Cout * 4 < 131072 || !ws) return 1;
  if (int rc = tma_rows_from_u8(x, g, (float*)ws, wsb, s)) return rc;
  const int64_t used = rows_used(g);
  return tma_conv_fwd_rows((const float*)ws, w, bias, y, g, act, (char*)ws + used, wsb - used, s);
}
int tma_conv_wgrad_u8(const uint8_t* x, const float* dy, float* dw, float* db, const b200rl_conv_geom& g, void* ws, int64_t wsb,
                      cudaStream_t s) {
  const int M = g.B * g.OH * g.OW;
  if (!tma_ok(dy, g.Cout) || g.Cout % 32 != 0 || (int64_t)M * g.Cout * 4 < 131072 || M < GBK || !ws) return 1;
  if (int rc = tma_rows_from_u8(x, g, (float*)ws, wsb, s)) return rc;
  const int64_t used = rows_used(g);
  return tma_conv_wgrad_rows((const float*)ws, dy, dw, db, g, (char*)ws + used, wsb - used, s);
}

// ---- conv data gradient as an implicit GEMM over dy, one sub-problem per stride phase.
// With n = i + pad and n = s q + p (p = phase, 0 <= p < s) an input pixel i only meets the taps k = p + s t:
//     dx[i, ci] = sum_{t, co} dy[q - t, co] w[co, p + s t, ci]        (separately in y and x)
// i.e. every phase is a stride-1 correlation of dy with a T_y x T_x sub-filter, T = ceil((k - p) / s).  A = TMA im2col
// over dy (one map per phase: the bounding box selects that phase's q range, out-of-range dy reads as zero),
// B = 32 x C boxes of the untransposed weight matrix [Cout][kh kw C] (MN-major), accumulator rows are scattered
// to the phase's pixels of dx with the producer layer's activation derivative applied.
struct DgradPhase {
  int tile_begin;            // first CTA of this phase
  int cnt_x, cnt_y;          // q range sizes
  int lower_x, lower_y;      // im2col base coordinate of q = qmin (qmin - (T - 1))
  int Tx, Ty, px, py;
  int ix0, iy0;              // input coordinate of q = qmin
};
struct DgradParams {
  DgradPhase ph[4];
  int nphase, stride, kw, C, Cout, W, H, B;
};
struct DgradMaps { CUtensorMap m[4]; };

template <int BN>
__global__ void __launch_bounds__(G_THREADS)
tma_conv_dgrad_kernel(const __grid_constant__ DgradMaps maps, const __grid_constant__ CUtensorMap map_w, DgradParams P,
                      float* __restrict__ dx, const float* __restrict__ mask, int mask_act) {
  constexpr int A_BYTES = GBM * 128, B_BYTES = BN * 128, STAGE = A_BYTES + B_BYTES;
  constexpr uint32_t TMEM_COLS = BN < 32 ? 32 : BN;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  __shared__ __align__(8) uint64_t bar_full[G_STAGES], bar_empty[G_STAGES], bar_done;
  __shared__ uint32_t tmem_base_smem;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  int phase = 0;
#pragma unroll
  for (int i = 1; i < 4; ++i) if (i < P.nphase && (int)blockIdx.x >= P.ph[i].tile_begin) phase = i;
  const DgradPhase ph = P.ph[phase];
  const int row0 = ((int)blockIdx.x - ph.tile_begin) * GBM;
  const int Mp = P.B * ph.cnt_y * ph.cnt_x;
  const int cb = P.Cout / GBK;
  const int nkb = ph.Ty * ph.Tx * cb;

  if (tid == 0) {
    prefetch_tensormap(&maps.m[phase]);
    prefetch_tensormap(&map_w);
#pragma unroll
    for (int s = 0; s < G_STAGES; ++s) { mbar_init(&bar_full[s], 1); mbar_init(&bar_empty[s], 1); }
    mbar_init(&bar_done, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) tmem_alloc(&tmem_base_smem, TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  pdl_wait();
  const uint32_t tmem_d = tmem_base_smem;

  if (warp == 0 && lane == 0) {
    const int jx = row0 % ph.cnt_x, t = row0 / ph.cnt_x;
    const int ax = ph.lower_x + jx, ay = ph.lower_y + t % ph.cnt_y, an = t / ph.cnt_y;
    const int K = P.kw * P.C;   // weight row: tap (ky, kx) starts at (ky kw + kx) C; rows are kh*K long
    int co0 = 0, off_x = 0, off_y = 0;   // advanced incrementally (no divisions in the issue loop)
    for (int i = 0; i < nkb; ++i) {
      const int s = i % G_STAGES;
      if (i >= G_STAGES) mbar_wait(&bar_empty[s], ((i / G_STAGES) - 1) & 1);
      const uint32_t sa = base + s * STAGE, sb = sa + A_BYTES;
      const int ky = ph.py + P.stride * (ph.Ty - 1 - off_y), kx = ph.px + P.stride * (ph.Tx - 1 - off_x);
      mbar_expect_tx(&bar_full[s], STAGE);
      tma_load_im2col(sa, &maps.m[phase], co0, ax, ay, an, (uint16_t)off_x, (uint16_t)off_y, &bar_full[s]);
#pragma unroll
      for (int j = 0; j < BN / 32; ++j) tma_load_2d(sb + j * 4096, &map_w, ky * K + kx * P.C + 32 * j, co0, &bar_full[s]);
      co0 += GBK;
      if (co0 == P.Cout) {
        co0 = 0;
        if (++off_x == ph.Tx) { off_x = 0; ++off_y; }
      }
    }
  } else if (warp == 1 && lane == 0) {
    constexpr uint32_t idesc = umma_idesc_tf32(GBM, BN, false, true);
    for (int i = 0; i < nkb; ++i) {
      const int s = i % G_STAGES;
      mbar_wait(&bar_full[s], (i / G_STAGES) & 1);
      tc_fence_after();
      const uint32_t sa = base + s * STAGE, sb = sa + A_BYTES;
#pragma unroll
      for (int kk = 0; kk < GBK / 8; ++kk)
        umma_tf32(tmem_d, umma_desc_sw128(sa + kk * 32, 16, 1024, 2), umma_desc_sw128(sb + kk * 1024, 4096, 512, 1), idesc,
                  (i > 0 || kk > 0) ? 1u : 0u);
      umma_commit(&bar_empty[s]);
      if (i == nkb - 1) umma_commit(&bar_done);
    }
  } else if (warp >= 2) {
    mbar_wait(&bar_done, 0);
    tc_fence_after();
    pdl_launch_dependents();   // dependents are scheduled while this tile's epilogue runs
    const int lane_base = (warp & 3) * 32;
    const int row = row0 + lane_base + lane;
    long long orow = -1;   // destination pixel of this lane's accumulator row
    if (row < Mp) {
      const int jx = row % ph.cnt_x, t = row / ph.cnt_x;
      const int jy = t % ph.cnt_y, b = t / ph.cnt_y;
      orow = ((long long)b * P.H + (ph.iy0 + P.stride * jy)) * P.W + (ph.ix0 + P.stride * jx);
    }
    Epilogue e{dx, P.C, nullptr, 0, mask, P.C, mask_act, nullptr, 0};
    float* slab = reinterpret_cast<float*>(smem_raw + (base - smem_u32(smem_raw))) + (warp & 3) * 32 * BN;
    stage_accumulator<BN>(tmem_d + ((uint32_t)lane_base << 16), slab, lane, true);
    store_staged_rows<BN>(e, slab, lane, 0, P.C, 0, [&](int r) -> long long { return __shfl_sync(0xffffffffu, orow, r); });
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_d, TMEM_COLS);
}

template <int BN>
static int launch_dgrad(const DgradMaps& maps, const CUtensorMap& mw, const DgradParams& P, int tiles, float* dx,
                        const float* mask, int mask_act, cudaStream_t s) {
  constexpr int smem = G_STAGES * (GBM * 128 + BN * 128) + 1024;
  static bool attr = false;
  if (!attr) {
    B200RL_CUDA_OK(cudaFuncSetAttribute(tma_conv_dgrad_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    attr = true;
  }
  B200RL_CUDA_OK(launch_pdl(tma_conv_dgrad_kernel<BN>, dim3(tiles), dim3(G_THREADS), smem, s, maps, mw, P, dx, mask, mask_act));
  B200RL_LAUNCH_OK();
  return B200RL_OK;
}

int tma_conv_dgrad(const float* dy, const float* w, float* dx, const b200rl_conv_geom& g, const float* mask, int mask_act,
                   void* ws, int64_t wsb, cudaStream_t s) {
  (void)ws; (void)wsb;
  const int K = g.kh * g.kw * g.C, st = g.stride;
  if (st > 2 || g.kh < st || g.kw < st || g.Cout % GBK != 0 || (g.C != 32 && g.C != 64 && g.C != 128) || !tma_ok(dy, g.Cout) ||
      !tma_ok(w, K) || !tma_ok(dx, g.C) || (mask && !tma_ok(mask, g.C)) || (int64_t)g.B * g.OH * g.OW * g.Cout * 4 < 131072 ||
      (int64_t)g.B * g.H * g.W >= (1ll << 31) / g.C)
    return 1;
  EncodeIm2colFn enc = get_encode_im2col();
  if (!enc) return 1;
  DgradMaps maps;
  DgradParams P;
  P.nphase = st * st; P.stride = st; P.kw = g.kw; P.C = g.C; P.Cout = g.Cout; P.W = g.W; P.H = g.H; P.B = g.B;
  auto fdiv = [](int a, int b) { return a >= 0 ? a / b : -((-a + b - 1) / b); };
  int tiles = 0;
  for (int py = 0; py < st; ++py)
    for (int px = 0; px < st; ++px) {
      DgradPhase& ph = P.ph[py * st + px];
      const int Ty = (g.kh - py + st - 1) / st, Tx = (g.kw - px + st - 1) / st;
      const int qy0 = -fdiv(-(g.pad_top - py), st) , qx0 = -fdiv(-(g.pad_left - px), st);          // ceil
      const int qy1 = fdiv(g.H - 1 + g.pad_top - py, st), qx1 = fdiv(g.W - 1 + g.pad_left - px, st);  // floor
      ph.tile_begin = tiles;
      ph.cnt_x = qx1 - qx0 + 1; ph.cnt_y = qy1 - qy0 + 1;
      ph.Tx = Tx; ph.Ty = Ty; ph.px = px; ph.py = py;
      ph.lower_x = qx0 - (Tx - 1); ph.lower_y = qy0 - (Ty - 1);
      ph.ix0 = st * qx0 + px - g.pad_left; ph.iy0 = st * qy0 + py - g.pad_top;
      if (ph.cnt_x <= 0 || ph.cnt_y <= 0) return 1;
      tiles += ceil_div(g.B * ph.cnt_x * ph.cnt_y, GBM);
      cuuint64_t dims[4] = {(cuuint64_t)g.Cout, (cuuint64_t)g.OW, (cuuint64_t)g.OH, (cuuint64_t)g.B};
      cuuint64_t strides[3] = {(cuuint64_t)g.Cout * 4, (cuuint64_t)g.OW * g.Cout * 4, (cuuint64_t)g.OH * g.OW * g.Cout * 4};
      int lower[2] = {ph.lower_x, ph.lower_y};
      int upper[2] = {(qx1 - (Tx - 1)) - (g.OW - 1), (qy1 - (Ty - 1)) - (g.OH - 1)};
      cuuint32_t es[4] = {1, 1, 1, 1};
      if (enc(&maps.m[py * st + px], CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, (void*)dy, dims, strides, lower, upper, GBK, GBM, es,
              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
        return 1;
    }
  for (int i = P.nphase; i < 4; ++i) { P.ph[i] = P.ph[0]; maps.m[i] = maps.m[0]; }
  CUtensorMap mw;
  if (!make_map(&mw, w, g.Cout, K, K, 32, GBK, true)) return 1;
  if (g.C == 32) return launch_dgrad<32>(maps, mw, P, tiles, dx, mask, mask_act, s);
  if (g.C == 64) return launch_dgrad<64>(maps, mw, P, tiles, dx, mask, mask_act, s);
  return launch_dgrad<128>(maps, mw, P, tiles, dx, mask, mask_act, s);
}

}  // namespace b200rl

extern "C" int b200rl_debug_tma_stage_mode(int mode) { b200rl::g_tma_stage_mode = mode; return 0; }
extern "C" int b200rl_debug_tma_dgrad_bn64(int on) { b200rl::g_tma_dgrad_bn64 = on; return 0; }
extern "C" int b200rl_debug_tma_timeline(unsigned long long* buf_dev) {
  cudaError_t e = cudaMemcpyToSymbol(b200rl::g_tma_timeline, &buf_dev, sizeof(buf_dev));
  return e == cudaSuccess ? 0 : -2;
}
