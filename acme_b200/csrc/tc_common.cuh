// tcgen05 / TMEM / mbarrier PTX wrappers and the register-level epilogue shared by the tensor-core
// GEMM kernels (gemm_tc.cu: loader-fed bf16, gemm_tma.cu: TMA-fed tf32).
#pragma once
#include <cuda_bf16.h>

#include "common.cuh"
#include "gemm_common.cuh"

namespace b200rl {

// ------------------------------------------------------------------------------ PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  const uint32_t addr = smem_u32(bar);
  do {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(addr), "r"(parity)
        : "memory");
  } while (!ok);
}
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

// Epilogue of 16 consecutive columns of one row held by one thread: vectorised (float4) whenever the
// destination allows it, so every 32-byte sector is written whole.  The transcendental activations are
// kept out of line: inlining them 16x per column group bloats the kernel past the instruction caches
// (a 128-column epilogue then runs at instruction-fetch speed).
static __device__ __noinline__ float act_eval_slow(float v, int act) { return apply_act(v, act); }
static __device__ __noinline__ float act_grad_slow(float y, int act) { return act_grad_out(y, act); }

__device__ __forceinline__ void finish16(const Epilogue& e, int row, int c0, int N, int M, float (&v)[16]) {
  const bool full = c0 + 15 < N;
  if (e.partial) {
    float* dst = e.partial + ((size_t)blockIdx.z * M + row) * N + c0;
    if (full && ((((uintptr_t)dst) & 15) == 0)) {
#pragma unroll
      for (int q = 0; q < 4; ++q) reinterpret_cast<float4*>(dst)[q] = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
    } else {
#pragma unroll
      for (int j = 0; j < 16; ++j) if (c0 + j < N) dst[j] = v[j];
    }
    return;
  }
  // bias, activation, activation-derivative mask -- each decided once per 16 columns
  if (e.bias) {
#pragma unroll
    for (int j = 0; j < 16; ++j) if (c0 + j < N) v[j] += __ldg(e.bias + c0 + j);
  }
  if (e.act == B200RL_ACT_RELU) {
#pragma unroll
    for (int j = 0; j < 16; ++j) v[j] = fmaxf(v[j], 0.f);
  } else if (e.act != B200RL_ACT_NONE) {
#pragma unroll 1
    for (int j = 0; j < 16; ++j) v[j] = act_eval_slow(v[j], e.act);
  }
  if (e.mask) {
    const float* msk = e.mask + (size_t)row * e.ldmask + c0;
    if (e.mask_act == B200RL_ACT_RELU) {
#pragma unroll
      for (int j = 0; j < 16; ++j) if (c0 + j < N) v[j] = __ldg(msk + j) > 0.f ? v[j] : 0.f;
    } else {
#pragma unroll 1
      for (int j = 0; j < 16; ++j) if (c0 + j < N) v[j] *= act_grad_slow(__ldg(msk + j), e.mask_act);
    }
  }
  if (e.transpose_out) {   // lanes hold consecutive rows -> consecutive addresses for a fixed column
#pragma unroll
    for (int j = 0; j < 16; ++j) if (c0 + j < N) e.out[(size_t)(c0 + j) * e.ldo + row] = v[j];
    return;
  }
  float* dst = e.out + (size_t)row * e.ldo + c0;
  if (full && ((((uintptr_t)dst) & 15) == 0)) {
#pragma unroll
    for (int q = 0; q < 4; ++q) reinterpret_cast<float4*>(dst)[q] = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
  } else {
#pragma unroll
    for (int j = 0; j < 16; ++j) if (c0 + j < N) dst[j] = v[j];
  }
}

}  // namespace b200rl
