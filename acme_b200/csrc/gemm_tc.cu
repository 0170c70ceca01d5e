// K6, speed mode: bf16 x bf16 -> fp32 GEMM on the 5th-generation tensor cores (tcgen05.mma,
// accumulator in TMEM), used for every dense contraction of the Q-network / critic / policy
// (acme/tf/networks/atari.py:36-69, duelling.py:37-59, continuous.py:37-68).
//
// Shape of the kernel (one CTA = one 128 x BN output tile, optionally one split of K):
//   * all 8 warps are loaders: they read the operands straight from their fp32 / uint8 home
//     (implicit im2col for the convolutions, transposed views for dgrad / wgrad), round to bf16 in
//     registers and store 16-byte chunks into shared memory in the canonical no-swizzle UMMA
//     layouts.  Every operand is read along its memory-contiguous direction: if that is the
//     reduction index the tile is K-major ([k-chunk][row] x 16 B), otherwise it is MN-major
//     ([k-group of 8][mn-chunk][k % 8] x 16 B) and the instruction descriptor says so -- dgrad and
//     wgrad need no transposed copies and no strided loads.  Both layouts use SBO = 128 B and
//     LBO = extent * 16 B.  No col buffer, no staging copy in HBM.
//   * one elected thread issues tcgen05.mma (M = 128, N = BN, K = 16 per instruction) on the two
//     shared-memory descriptors; tcgen05.commit arrives on an mbarrier when the stage is consumed,
//     so the loads of k-block i+1 overlap the MMAs of k-block i (2 stages).
//   * epilogue: tcgen05.ld (32 lanes x 16 columns per instruction) -> bias / activation /
//     activation-derivative mask -> global (or split-K partials).
// fp32 master weights, fp32 accumulation; only the multiplicands are rounded to bf16.
#include <algorithm>
#include <type_traits>

#include <cuda_bf16.h>

#include "common.cuh"
#include "gemm_common.cuh"
#include "tc_common.cuh"

namespace b200rl {

__device__ unsigned long long* g_tc_timeline = nullptr;   // tools/tc_timeline.py: phase timestamps of CTA 0
__device__ __forceinline__ void tl_mark(int slot) {
  if (g_tc_timeline && threadIdx.x == 0 && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    g_tc_timeline[slot] = t;
  }
}

constexpr int TBM = 128, TBK = 64, TC_THREADS = 256, TC_STAGES = 2;

// K-major, no swizzle: start address, LBO (k-chunk stride), SBO (8-row group stride), version 1
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  return (uint64_t)((saddr >> 4) & 0x3FFFu) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16) |
         ((uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32) | (1ull << 46);
}
// kind::f16: D = f32 (bits 4-5 = 1), A = B = bf16 (bits 7-9, 10-12 = 1), major bits 15 / 16
// (0 = K-major, 1 = MN-major), N >> 3 at bit 17, M >> 4 at bit 24
__host__ __device__ constexpr uint32_t umma_idesc(int M, int N, bool a_mn, bool b_mn) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((a_mn ? 1u : 0u) << 15) | ((b_mn ? 1u : 0u) << 16) |
         ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// ------------------------------------------------------------------------------ operand accessors
// An operand tile is filled in 16-byte units of 8 memory-contiguous elements.  An operand whose
// contiguous direction is the reduction index is used K-major (a unit = fixed row/column of the
// product, 8 consecutive k); otherwise MN-major (a unit = fixed k, 8 consecutive rows/columns).
// Per thread the non-k coordinates of its units never change, so every accessor splits its work:
//   tables(tab)          once per CTA: small lookup tables in shared memory (filter-tap offsets)
//   init(fix0, fix1)     once per unit slot: decode the fixed coordinates (pixel -> b, y, x ...)
//   fetch(state, k0)     per k-block: a couple of adds / compares and the predicated loads;
//                        the loaded registers are NOT touched, so loads stay in flight
//   pack_raw(raw)        when the tile is written to shared memory: round to bf16
// This keeps the integer work per unit to a handful of instructions (in this network the tensor
// cores wait on the loaders, not the other way round).
struct FastDiv {   // n / d for 0 <= n < 2^31 via one multiply-high (d >= 1)
  uint32_t d, magic, shift;
  __host__ static FastDiv make(uint32_t d) {
    FastDiv f;
    f.d = d;
    uint32_t s = 0;
    while ((1ull << s) < d) ++s;
    f.shift = s;
    f.magic = (uint32_t)(((1ull << 32) * ((1ull << s) - d)) / d + 1);
    return f;
  }
  __device__ __forceinline__ uint32_t div(uint32_t n) const { return (uint32_t)(((uint64_t)__umulhi(n, magic) + n) >> shift); }
  __device__ __forceinline__ void divmod(uint32_t n, uint32_t& q, uint32_t& r) const { q = div(n); r = n - q * d; }
};

struct Raw8f { float4 a, b; };
struct Raw8b { uint32_t a, b; };
__device__ __forceinline__ float4 zero4() { return make_float4(0.f, 0.f, 0.f, 0.f); }
__device__ __forceinline__ uint32_t bf16x2(float lo, float hi) {
  __nv_bfloat162 t = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&t);
}
__device__ __forceinline__ uint4 pack_raw(const Raw8f& r, const uint16_t*) {
  return make_uint4(bf16x2(r.a.x, r.a.y), bf16x2(r.a.z, r.a.w), bf16x2(r.b.x, r.b.y), bf16x2(r.b.z, r.b.w));
}
// uint8 frames: lut[u] = bf16(float32(u) / 255) (atari_wrapper.py:303-304 followed by the bf16 rounding)
__device__ __forceinline__ uint32_t lut2(const uint16_t* lut, uint32_t w, int s) {
  return (uint32_t)lut[(w >> s) & 0xffu] | ((uint32_t)lut[(w >> (s + 8)) & 0xffu] << 16);
}
__device__ __forceinline__ uint4 pack_raw(const Raw8b& r, const uint16_t* lut) {
  return make_uint4(lut2(lut, r.a, 0), lut2(lut, r.a, 16), lut2(lut, r.b, 0), lut2(lut, r.b, 16));
}
__device__ __forceinline__ Raw8f load8f(const float* q, bool ok) {
  Raw8f r{zero4(), zero4()};
  if (ok) {
    r.a = __ldg(reinterpret_cast<const float4*>(q));
    r.b = __ldg(reinterpret_cast<const float4*>(q) + 1);
  }
  return r;
}
__device__ __forceinline__ Raw8f load8f_ragged(const float* q, int pos, int n_pos) {
  float t[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) t[j] = (pos + j < n_pos) ? __ldg(q + j) : 0.f;
  return Raw8f{make_float4(t[0], t[1], t[2], t[3]), make_float4(t[4], t[5], t[6], t[7])};
}

constexpr int TAB_INTS = 1024;

// ---- dense fp32 matrix, element (line, pos) at p[line * ld + pos]
struct RowMatK {   // K-major use: fixed line, pos = k
  static constexpr bool MN = false;
  using raw_t = Raw8f;
  struct State { const float* q; int kofs; bool ok; };
  const float* p; int ld, n_lines, n_pos;
  __device__ __forceinline__ void tables(int*) const {}
  __device__ __forceinline__ State init(int line, int kofs, const int*) const {
    return State{p + (size_t)(line < n_lines ? line : 0) * ld + kofs, kofs, line < n_lines};
  }
  __device__ __forceinline__ raw_t fetch(const State& s, int k0, const int*) const {
    const int pos = k0 + s.kofs;
    const float* q = s.q + k0;
    if (s.ok && pos + 7 < n_pos && ((((uintptr_t)q) & 15) == 0)) return load8f(q, true);
    if (s.ok && pos < n_pos) return load8f_ragged(q, pos, n_pos);   // ragged edge / unaligned rows
    return raw_t{zero4(), zero4()};
  }
};
struct RowMatMN {  // MN-major use: fixed pos, line = k
  static constexpr bool MN = true;
  using raw_t = Raw8f;
  struct State { const float* q; int lofs; int pos; };
  const float* p; int ld, n_lines, n_pos;
  __device__ __forceinline__ void tables(int*) const {}
  __device__ __forceinline__ State init(int lofs, int pos, const int*) const {
    return State{p + (pos < n_pos ? pos : 0), lofs, pos};
  }
  __device__ __forceinline__ raw_t fetch(const State& s, int k0, const int*) const {
    const int line = k0 + s.lofs;
    const float* q = s.q + (size_t)line * ld;
    const bool in = line < n_lines && s.pos < n_pos;
    if (in && s.pos + 7 < n_pos && ((((uintptr_t)q) & 15) == 0)) return load8f(q, true);
    if (in) return load8f_ragged(q, s.pos, n_pos);
    return raw_t{zero4(), zero4()};
  }
};

// ---- implicit im2col of an NHWC tensor: pixel (b, oy, ox) x patch index (ky, kx, ci); C % 4 == 0.
// 8 consecutive patch indices = two groups of 4 channels (one tap when C % 8 == 0, two adjacent taps
// when C == 4).  tab[q] for channel group q = k / 4:  ((ky * W + kx) * C + ci) | ky << 20 | kx << 26.
struct Im2colGeom {
  const void* x; b200rl_conv_geom g; int Mtot, Ktot;
  FastDiv dOW, dOH, dC, dKW;
  __host__ static Im2colGeom make(const void* x, const b200rl_conv_geom& g) {
    Im2colGeom o;
    o.x = x; o.g = g; o.Mtot = g.B * g.OH * g.OW; o.Ktot = g.kh * g.kw * g.C;
    o.dOW = FastDiv::make(g.OW); o.dOH = FastDiv::make(g.OH); o.dC = FastDiv::make(g.C); o.dKW = FastDiv::make(g.kw);
    return o;
  }
  __device__ __forceinline__ void tables(int* tab) const {
    for (int q = threadIdx.x; q < Ktot / 4; q += blockDim.x) {
      uint32_t t2, ci, ky, kx;
      dC.divmod(q * 4, t2, ci);
      dKW.divmod(t2, ky, kx);
      tab[q] = (int)(((ky * g.W + kx) * g.C + ci) | (ky << 20) | (kx << 26));
    }
  }
  // origin of the receptive field of pixel m
  __device__ __forceinline__ bool pixel(int m, int& iy0, int& ix0, long long& base) const {
    uint32_t t, ox, b, oy;
    const bool ok = m < Mtot;
    dOW.divmod(ok ? m : 0, t, ox);
    dOH.divmod(t, b, oy);
    iy0 = (int)oy * g.stride - g.pad_top;
    ix0 = (int)ox * g.stride - g.pad_left;
    base = (((long long)b * g.H + iy0) * g.W + ix0) * g.C;
    return ok;
  }
  __device__ __forceinline__ long long tap(int e, int iy0, int ix0, long long base) const {
    const int ky = (e >> 20) & 63, kx = (e >> 26) & 31;
    const bool ok = (unsigned)(iy0 + ky) < (unsigned)g.H && (unsigned)(ix0 + kx) < (unsigned)g.W;
    return ok ? base + (e & 0xfffff) : -1;
  }
};
template <bool U8> struct Im2colRaw { using type = Raw8f; };
template <> struct Im2colRaw<true> { using type = Raw8b; };
template <bool U8>
__device__ __forceinline__ typename Im2colRaw<U8>::type im2col_load(const void* x, long long o0, long long o1) {
  typename Im2colRaw<U8>::type r;
  if constexpr (U8) {
    const uint8_t* base = (const uint8_t*)x;
    r.a = o0 >= 0 ? __ldg(reinterpret_cast<const uint32_t*>(base + o0)) : 0u;
    r.b = o1 >= 0 ? __ldg(reinterpret_cast<const uint32_t*>(base + o1)) : 0u;
  } else {
    const float* base = (const float*)x;
    r.a = o0 >= 0 ? __ldg(reinterpret_cast<const float4*>(base + o0)) : zero4();
    r.b = o1 >= 0 ? __ldg(reinterpret_cast<const float4*>(base + o1)) : zero4();
  }
  return r;
}
template <bool U8>
struct Im2colRowsK : Im2colGeom {   // conv fwd A: fixed pixel, k runs over the patch
  static constexpr bool MN = false;
  using raw_t = typename Im2colRaw<U8>::type;
  struct State { long long base; int iy0, ix0, kofs; bool ok; };
  __device__ __forceinline__ State init(int m, int kofs, const int*) const {
    State s;
    s.kofs = kofs;
    s.ok = pixel(m, s.iy0, s.ix0, s.base);
    return s;
  }
  __device__ __forceinline__ raw_t fetch(const State& s, int k0, const int* tab) const {
    const int q = (k0 + s.kofs) >> 2;
    long long o0 = -1, o1 = -1;
    if (s.ok && q * 4 < Ktot) o0 = tap(tab[q], s.iy0, s.ix0, s.base);
    if (s.ok && q * 4 + 4 < Ktot) o1 = tap(tab[q + 1], s.iy0, s.ix0, s.base);
    return im2col_load<U8>(x, o0, o1);
  }
};
template <bool U8>
struct Im2colPixelsMN : Im2colGeom {   // conv wgrad A: fixed patch index, k runs over the pixels
  static constexpr bool MN = true;
  using raw_t = typename Im2colRaw<U8>::type;
  struct State { int e0, e1, lofs; };
  __device__ __forceinline__ State init(int lofs, int pos, const int* tab) const {
    State s;
    s.lofs = lofs;
    s.e0 = pos < Ktot ? tab[pos >> 2] : -1;
    s.e1 = pos + 4 < Ktot ? tab[(pos >> 2) + 1] : -1;
    return s;
  }
  __device__ __forceinline__ raw_t fetch(const State& s, int k0, const int*) const {
    int iy0, ix0;
    long long base;
    const bool ok = pixel(k0 + s.lofs, iy0, ix0, base);
    const long long o0 = (ok && s.e0 >= 0) ? tap(s.e0, iy0, ix0, base) : -1;
    const long long o1 = (ok && s.e1 >= 0) ? tap(s.e1, iy0, ix0, base) : -1;
    return im2col_load<U8>(x, o0, o1);
  }
};

// ---- conv dgrad in gather form: dx[pixel, ci] = sum_r dy(pixel, r) * w(r, ci), r = (ky, kx, co); Cout % 8 == 0
struct DgradRowsK {   // A: fixed input pixel, 8 consecutive co of one tap.  tab[r / 8] = co | kx << 16 | ky << 24
  static constexpr bool MN = false;
  using raw_t = Raw8f;
  struct State { int b, iyp, ixp, kofs; bool ok; };
  const float* dy; b200rl_conv_geom g; int Mtot, Rtot;
  FastDiv dW, dH, dCout, dKW;
  __host__ static DgradRowsK make(const float* dy, const b200rl_conv_geom& g) {
    DgradRowsK o;
    o.dy = dy; o.g = g; o.Mtot = g.B * g.H * g.W; o.Rtot = g.kh * g.kw * g.Cout;
    o.dW = FastDiv::make(g.W); o.dH = FastDiv::make(g.H); o.dCout = FastDiv::make(g.Cout); o.dKW = FastDiv::make(g.kw);
    return o;
  }
  __device__ __forceinline__ void tables(int* tab) const {
    for (int q = threadIdx.x; q < Rtot / 8; q += blockDim.x) {
      uint32_t t2, co, ky, kx;
      dCout.divmod(q * 8, t2, co);
      dKW.divmod(t2, ky, kx);
      tab[q] = (int)(co | (kx << 16) | (ky << 24));
    }
  }
  __device__ __forceinline__ State init(int m, int kofs, const int*) const {
    State s;
    s.ok = m < Mtot;
    uint32_t t, ix, b, iy;
    dW.divmod(s.ok ? m : 0, t, ix);
    dH.divmod(t, b, iy);
    s.b = b; s.iyp = iy + g.pad_top; s.ixp = ix + g.pad_left; s.kofs = kofs;
    return s;
  }
  __device__ __forceinline__ raw_t fetch(const State& s, int k0, const int* tab) const {
    const int r = k0 + s.kofs;
    bool ok = s.ok && r < Rtot;
    const int e = tab[ok ? (r >> 3) : 0];
    const int co = e & 0xffff, kx = (e >> 16) & 0xff, ky = (e >> 24) & 0xff;
    const int ny = s.iyp - ky, nx = s.ixp - kx;
    int oy = ny, ox = nx;
    ok = ok && ny >= 0 && nx >= 0;
    if (g.stride == 2) { ok = ok && !((ny | nx) & 1); oy = ny >> 1; ox = nx >> 1; }
    else if (g.stride != 1) { ok = ok && (ny % g.stride == 0) && (nx % g.stride == 0); oy = ny / g.stride; ox = nx / g.stride; }
    ok = ok && oy < g.OH && ox < g.OW;
    return load8f(dy + (((size_t)s.b * g.OH + (ok ? oy : 0)) * g.OW + (ok ? ox : 0)) * g.Cout + co, ok);
  }
};
struct DgradWMN {   // B: fixed ci, k runs over r = (ky, kx, co): w[co][ky][kx][ci].  tab[r] = offset of (r, ci = 0)
  static constexpr bool MN = true;
  using raw_t = Raw8f;
  struct State { const float* q; int lofs; bool ok; };
  const float* w; b200rl_conv_geom g; int Rtot;
  FastDiv dCout, dKW;
  __host__ static DgradWMN make(const float* w, const b200rl_conv_geom& g) {
    DgradWMN o;
    o.w = w; o.g = g; o.Rtot = g.kh * g.kw * g.Cout;
    o.dCout = FastDiv::make(g.Cout); o.dKW = FastDiv::make(g.kw);
    return o;
  }
  __device__ __forceinline__ void tables(int* tab) const {
    for (int r = threadIdx.x; r < Rtot; r += blockDim.x) {
      uint32_t t2, co, ky, kx;
      dCout.divmod(r, t2, co);
      dKW.divmod(t2, ky, kx);
      tab[r] = (int)(((co * g.kh + ky) * g.kw + kx) * g.C);
    }
  }
  __device__ __forceinline__ State init(int lofs, int ci, const int*) const {
    return State{w + (ci < g.C ? ci : 0), lofs, ci < g.C};
  }
  __device__ __forceinline__ raw_t fetch(const State& s, int k0, const int* tab) const {
    const int r = k0 + s.lofs;
    const bool ok = s.ok && r < Rtot;
    return load8f(s.q + tab[ok ? r : 0], ok);
  }
};

// One operand tile = R (rows or columns of the product) x 64 (k) elements = R * 8 units of 16 bytes.
// A unit is (line, chunk): K-major: line = r in [0,R), chunk = k / 8 in [0,8);
//                          MN-major: line = k in [0,64), chunk = r / 8 in [0,R/8).
// A warp instruction covers 8 lines x 4 chunks: each quarter-warp writes 8 consecutive lines of one
// chunk (128 contiguous bytes of shared memory in both layouts -> conflict-free) and the 4
// quarter-warps read 4 adjacent chunks (128 contiguous bytes of global memory per line).
template <bool MN, int R>
struct TileMap {
  static constexpr int NCH = MN ? R / 8 : 8;
  static constexpr int NL = MN ? 64 : R;
  static constexpr int CG = NCH / 4;                 // chunk groups
  static constexpr int PASSES = (NL / 8) * CG / 8;   // 8 warps per pass
  static_assert(NCH % 4 == 0 && PASSES >= 1, "tile too small");
  __device__ static __forceinline__ void coords(int pass, int tid, int& line, int& chunk) {
    const int warp = tid >> 5, lane = tid & 31;
    const int bi = pass * 8 + warp;
    line = (bi / CG) * 8 + (lane & 7);
    chunk = (bi % CG) * 4 + (lane >> 3);
  }
  __device__ static __forceinline__ int smem_off(int line, int chunk) {
    return MN ? ((line >> 3) * R * 16 + chunk * 128 + (line & 7) * 16) : ((chunk * R + line) * 16);
  }
};


// ------------------------------------------------------------------------------ the kernel
template <class AL, class BL, int BN>
__global__ void __launch_bounds__(TC_THREADS)
tc_gemm_kernel(AL a, BL b, Epilogue epi, int M, int N, int K, int kblocks_per_split) {
  pdl_launch_dependents();
  pdl_wait();
  constexpr bool AMN = AL::MN, BMN = BL::MN;
  using MapA = TileMap<AMN, TBM>;
  using MapB = TileMap<BMN, BN>;
  constexpr int A_PASSES = MapA::PASSES, B_PASSES = MapB::PASSES;
  constexpr int A_BYTES = TBM * TBK * 2, B_BYTES = BN * TBK * 2;
  constexpr uint32_t TMEM_COLS = BN < 32 ? 32 : BN;
  extern __shared__ __align__(128) uint8_t smem[];
  uint8_t* sA[TC_STAGES];
  uint8_t* sB[TC_STAGES];
#pragma unroll
  for (int s = 0; s < TC_STAGES; ++s) { sA[s] = smem + s * (A_BYTES + B_BYTES); sB[s] = sA[s] + A_BYTES; }
  __shared__ __align__(8) uint64_t bar_free[TC_STAGES];
  __shared__ __align__(8) uint64_t bar_done;
  __shared__ uint32_t tmem_base_smem;
  __shared__ int tab_a[TAB_INTS], tab_b[TAB_INTS];
  __shared__ uint16_t lut[256];

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  tl_mark(0);
  const int row0 = blockIdx.y * TBM, col0 = blockIdx.x * BN;
  const int total_kblocks = (K + TBK - 1) / TBK;
  const int kb_begin = blockIdx.z * kblocks_per_split;
  const int kb_end = min(total_kblocks, kb_begin + kblocks_per_split);
  const int nkb = kb_end - kb_begin;

  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < TC_STAGES; ++s) mbar_init(&bar_free[s], 1);
    mbar_init(&bar_done, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) tmem_alloc(&tmem_base_smem, TMEM_COLS);
  a.tables(tab_a);
  b.tables(tab_b);
  {
    __nv_bfloat16 t = __float2bfloat16_rn(__fdiv_rn((float)tid, 255.f));
    lut[tid] = *reinterpret_cast<uint16_t*>(&t);   // TC_THREADS == 256
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_d = tmem_base_smem;
  tl_mark(1);

  // fixed coordinates of this thread's unit slots
  typename AL::State sta[A_PASSES];
  typename BL::State stb[B_PASSES];
#pragma unroll
  for (int it = 0; it < A_PASSES; ++it) {
    int l, c;
    MapA::coords(it, tid, l, c);
    sta[it] = AMN ? a.init(l, row0 + c * 8, tab_a) : a.init(row0 + l, c * 8, tab_a);
  }
#pragma unroll
  for (int it = 0; it < B_PASSES; ++it) {
    int l, c;
    MapB::coords(it, tid, l, c);
    stb[it] = BMN ? b.init(l, col0 + c * 8, tab_b) : b.init(col0 + l, c * 8, tab_b);
  }
  typename AL::raw_t ra[A_PASSES];
  typename BL::raw_t rb[B_PASSES];
  auto load_block = [&](int kb) {   // issue only: nothing below touches the loaded registers
    const int k0 = kb * TBK;
#pragma unroll
    for (int it = 0; it < A_PASSES; ++it) ra[it] = a.fetch(sta[it], k0, tab_a);
#pragma unroll
    for (int it = 0; it < B_PASSES; ++it) rb[it] = b.fetch(stb[it], k0, tab_b);
  };
  auto store_block = [&](int stage) {
#pragma unroll
    for (int it = 0; it < A_PASSES; ++it) {
      int l, c;
      MapA::coords(it, tid, l, c);
      *reinterpret_cast<uint4*>(sA[stage] + MapA::smem_off(l, c)) = pack_raw(ra[it], lut);
    }
#pragma unroll
    for (int it = 0; it < B_PASSES; ++it) {
      int l, c;
      MapB::coords(it, tid, l, c);
      *reinterpret_cast<uint4*>(sB[stage] + MapB::smem_off(l, c)) = pack_raw(rb[it], lut);
    }
  };

  constexpr uint32_t idesc = umma_idesc(TBM, BN, AMN, BMN);
  if (nkb > 0) load_block(kb_begin);
  for (int i = 0; i < nkb; ++i) {
    const int stage = i % TC_STAGES;
    if (i >= TC_STAGES) mbar_wait(&bar_free[stage], ((i / TC_STAGES) - 1) & 1);   // MMAs of block i-2 done
    if (i < 4) tl_mark(2 + 4 * i);
    store_block(stage);
    if (i < 4) tl_mark(3 + 4 * i);
    fence_async_smem();          // generic-proxy stores -> visible to the tensor core (async proxy)
    __syncthreads();
    if (tid == 0) {
      tc_fence_after();
      const uint32_t a0 = smem_u32(sA[stage]), b0 = smem_u32(sB[stage]);
#pragma unroll
      for (int kk = 0; kk < TBK / 16; ++kk) {
        // one MMA consumes 16 k = two 8-wide groups; in both layouts the group stride (LBO) is
        // extent * 16 B and the stride between 8-row / 8-column core matrices (SBO) is 128 B
        const uint64_t da = umma_desc(a0 + kk * 2 * TBM * 16, TBM * 16, 128);
        const uint64_t db = umma_desc(b0 + kk * 2 * BN * 16, BN * 16, 128);
        umma_bf16(tmem_d, da, db, idesc, (i > 0 || kk > 0) ? 1u : 0u);
      }
      umma_commit(&bar_free[stage]);
      if (i == nkb - 1) umma_commit(&bar_done);
    }
    if (i < 4) tl_mark(4 + 4 * i);
    if (i + 1 < nkb) load_block(kb_begin + i + 1);
    if (i < 4) tl_mark(5 + 4 * i);   // global loads in flight while the MMAs run
  }
  // ---- epilogue: warp w reads TMEM lanes 32*(w%4).., warps 0-3 take the low half of the columns
  tl_mark(18);
  if (nkb > 0) {
    mbar_wait(&bar_done, 0);
    tc_fence_after();
  }
  tl_mark(19);
  {
    const int lane_base = (warp & 3) * 32;
    const int row = row0 + lane_base + lane;
    constexpr int HALF = BN / 2 < 16 ? 16 : BN / 2;
    const int cbeg = (warp >> 2) * HALF;
#pragma unroll 1
    for (int c = 0; c < HALF; c += 16) {
      const int cc = cbeg + c;
      if (cc >= BN) break;
      float v[16];
      if (nkb > 0) tmem_ld16(tmem_d + ((uint32_t)lane_base << 16) + (uint32_t)cc, v);
      else {
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] = 0.f;
      }
      if (row < M) finish16(epi, row, col0 + cc, N, M, v);
    }
  }
  tl_mark(20);
  tc_fence_before();
  __syncthreads();
  tl_mark(21);
  if (warp == 0) tmem_dealloc(tmem_d, TMEM_COLS);
  tl_mark(22);
}

template <class AL, class BL, int BN>
static int launch_tc_bn(const AL& a, const BL& b, Epilogue epi, int M, int N, int K, void* ws, int64_t ws_bytes,
                        cudaStream_t stream) {
  constexpr int smem = TC_STAGES * (TBM * TBK * 2 + BN * TBK * 2);
  static bool attr = false;
  if (!attr) {
    B200RL_CUDA_OK(cudaFuncSetAttribute(tc_gemm_kernel<AL, BL, BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    attr = true;
  }
  const int tiles = ceil_div(M, TBM) * ceil_div(N, BN);
  const int kblocks = ceil_div(K, TBK);
  int splits = 1;
  if (tiles < kNumSMs && kblocks >= 8) {
    splits = std::min(ceil_div(2 * kNumSMs, tiles), kblocks / 4);
    splits = std::min(splits, 64);
    const int64_t cap = ws ? ws_bytes / ((int64_t)M * N * 4) : 0;
    splits = (int)std::max<int64_t>(1, std::min<int64_t>(splits, cap));
  }
  const int kps = ceil_div(kblocks, splits);
  splits = ceil_div(kblocks, kps);
  epi.partial = splits > 1 ? (float*)ws : nullptr;
  dim3 grid(ceil_div(N, BN), ceil_div(M, TBM), splits);
  B200RL_CUDA_OK(launch_pdl(tc_gemm_kernel<AL, BL, BN>, dim3(grid), dim3(TC_THREADS), smem, stream, a, b, epi, M, N, K, kps));
  B200RL_LAUNCH_OK();
  if (splits > 1) return launch_splitk_finish(epi, M, N, splits, stream);
  return B200RL_OK;
}

template <class AL, class BL>
static int launch_tc(const AL& a, const BL& b, const Epilogue& epi, int M, int N, int K, void* ws, int64_t wsb,
                     cudaStream_t s) {
  if (N <= 32) return launch_tc_bn<AL, BL, 32>(a, b, epi, M, N, K, ws, wsb, s);
  if (N <= 64) return launch_tc_bn<AL, BL, 64>(a, b, epi, M, N, K, ws, wsb, s);
  return launch_tc_bn<AL, BL, 128>(a, b, epi, M, N, K, ws, wsb, s);
}

// ---------------------------------------------------------------- bf16 entry points (precision 1)
// product dims (rows, cols, reduction)
int tc_linear_fwd(int M, int N, int K, const float* x, int ldx, const float* w, const float* bias, float* y,
                  int ldy, int act, void* ws, int64_t wsb, cudaStream_t s) {
  RowMatK a{x, ldx, M, K};            // x[m][k]
  RowMatK b{w, K, N, K};              // w[n][k]
  Epilogue e{y, ldy, bias, act, nullptr, 0, 0, nullptr, 0};
  return launch_tc(a, b, e, M, N, K, ws, wsb, s);
}
int tc_linear_dgrad(int M, int N, int K, const float* dy, int lddy, const float* w, float* dx, int lddx,
                    const float* mask, int ldmask, int mask_act, void* ws, int64_t wsb, cudaStream_t s) {
  RowMatK a{dy, lddy, M, N};          // dy[m][n]   (rows m, reduction n)
  RowMatMN b{w, K, N, K};             // w[n][k]    (line = reduction n, pos = column k)
  Epilogue e{dx, lddx, nullptr, 0, mask, ldmask, mask_act, nullptr, 0};
  return launch_tc(a, b, e, M, K, N, ws, wsb, s);
}
int tc_linear_wgrad(int M, int N, int K, const float* dy, int lddy, const float* x, int ldx, float* dw, float* db,
                    void* ws, int64_t wsb, cudaStream_t s) {
  RowMatMN a{dy, lddy, M, N};         // dy[m][n]   (line = reduction m, pos = row n)
  RowMatMN b{x, ldx, M, K};           // x[m][k]    (line = reduction m, pos = column k)
  Epilogue e{dw, K, nullptr, 0, nullptr, 0, 0, nullptr, 0};
  int rc = launch_tc(a, b, e, N, K, M, ws, wsb, s);
  if (rc) return rc;
  if (db) return launch_colsum(M, N, dy, lddy, db, ws, wsb, s);
  return B200RL_OK;
}
template <bool U8>
static int conv_fwd_t(const void* x, const float* w, const float* bias, float* y, const b200rl_conv_geom& g, int act,
                      void* ws, int64_t wsb, cudaStream_t s) {
  Im2colRowsK<U8> a;
  static_cast<Im2colGeom&>(a) = Im2colGeom::make(x, g);
  RowMatK b{w, a.Ktot, g.Cout, a.Ktot};   // w[co][k]
  Epilogue e{y, g.Cout, bias, act, nullptr, 0, 0, nullptr, 0};
  return launch_tc(a, b, e, a.Mtot, g.Cout, a.Ktot, ws, wsb, s);
}
int tc_conv_fwd(const void* x, int x_u8, const float* w, const float* bias, float* y, const b200rl_conv_geom& g,
                int act, void* ws, int64_t wsb, cudaStream_t s) {
  return x_u8 ? conv_fwd_t<true>(x, w, bias, y, g, act, ws, wsb, s) : conv_fwd_t<false>(x, w, bias, y, g, act, ws, wsb, s);
}
template <bool U8>
static int conv_wgrad_t(const void* x, const float* dy, float* dw, const b200rl_conv_geom& g, void* ws, int64_t wsb,
                        cudaStream_t s) {
  // dW^T[k, co] = sum_pixel col[pixel, k] * dy[pixel, co]: rows = patch index (256..576), cols = Cout
  Im2colPixelsMN<U8> a;
  static_cast<Im2colGeom&>(a) = Im2colGeom::make(x, g);
  RowMatMN b{dy, g.Cout, a.Mtot, g.Cout};   // dy[pixel][co]
  Epilogue e{dw, a.Ktot, nullptr, 0, nullptr, 0, 0, nullptr, 1};   // stored transposed: dw[co][k]
  return launch_tc(a, b, e, a.Ktot, g.Cout, a.Mtot, ws, wsb, s);
}
int tc_conv_wgrad(const void* x, int x_u8, const float* dy, float* dw, float* db, const b200rl_conv_geom& g,
                  void* ws, int64_t wsb, cudaStream_t s) {
  int rc = x_u8 ? conv_wgrad_t<true>(x, dy, dw, g, ws, wsb, s) : conv_wgrad_t<false>(x, dy, dw, g, ws, wsb, s);
  if (rc) return rc;
  if (db) return launch_colsum(g.B * g.OH * g.OW, g.Cout, dy, g.Cout, db, ws, wsb, s);
  return B200RL_OK;
}
int tc_conv_dgrad(const float* dy, const float* w, float* dx, const b200rl_conv_geom& g, const float* mask,
                  int mask_act, void* ws, int64_t wsb, cudaStream_t s) {
  DgradRowsK a = DgradRowsK::make(dy, g);
  DgradWMN b = DgradWMN::make(w, g);
  if (a.Rtot > TAB_INTS) {
    set_error("conv dgrad: kh*kw*Cout = %d exceeds the tap table (%d)", a.Rtot, TAB_INTS);
    return B200RL_EINVAL;
  }
  Epilogue e{dx, g.C, nullptr, 0, mask, g.C, mask_act, nullptr, 0};
  return launch_tc(a, b, e, a.Mtot, g.C, a.Rtot, ws, wsb, s);
}

}  // namespace b200rl

extern "C" int b200rl_debug_tc_timeline(unsigned long long* buf_dev) {
  cudaError_t e = cudaMemcpyToSymbol(b200rl::g_tc_timeline, &buf_dev, sizeof(buf_dev));
  return e == cudaSuccess ? 0 : -2;
}
