// Pieces shared by the fp32 SIMT GEMM (gemm_simt.cu) and the bf16 tcgen05 GEMM (gemm_tc.cu).
#pragma once
#include <cuda_bf16.h>

#include "common.cuh"

namespace b200rl {

// ---- epilogue
struct Epilogue {
  float* out; int ldo;
  const float* bias;   // per column, nullable
  int act;             // applied after bias
  const float* mask;   // nullable: multiply by act'(mask[row, col])
  int ldmask; int mask_act;
  float* partial;      // non-null: split-K partial sums [split][M][N], epilogue deferred
  int transpose_out;   // store C^T: out[col * ldo + row] (conv wgrad computes dW^T)
  int out_bf16;        // `out` holds __nv_bfloat16 (bf16 activation / gradient dataflow), else fp32
  int mask_bf16;       // `mask` holds __nv_bfloat16
  float scale;         // accumulator multiplier applied before the bias; 0 means 1 (conv1 on integer-valued frames: 1/255)
  const int* row_map;  // nullable: GEMM row r is output / mask row row_map[r] (phase-ordered conv data gradient, fp32 path)
  int pdl_late;        // programmatic dependent launch: release the dependents when the accumulator is ready (1) instead of at
                       // CTA start (0): early-scheduled dependents hold SM slots that other streams' kernels could use
};
__device__ __forceinline__ float epi_scale(const Epilogue& e) { return e.scale == 0.f ? 1.f : e.scale; }

__device__ __forceinline__ float apply_act(float v, int act) {
  switch (act) {
    case B200RL_ACT_RELU: return fmaxf(v, 0.f);
    case B200RL_ACT_ELU: return v > 0.f ? v : expm1f(v);
    case B200RL_ACT_TANH: return tanhf(v);
    default: return v;
  }
}
__device__ __forceinline__ float act_grad_out(float y, int act) {
  switch (act) {
    case B200RL_ACT_RELU: return y > 0.f ? 1.f : 0.f;
    case B200RL_ACT_ELU: return y > 0.f ? 1.f : y + 1.f;
    case B200RL_ACT_TANH: return 1.f - y * y;
    default: return 1.f;
  }
}
__device__ __forceinline__ void finish(const Epilogue& e, int row, int col, float acc) {
  if (e.row_map) row = e.row_map[row];
  acc *= epi_scale(e);
  if (e.bias) acc += e.bias[col];
  acc = apply_act(acc, e.act);
  if (e.mask) {
    const size_t mi = (size_t)row * e.ldmask + col;
    const float y = e.mask_bf16 ? __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(e.mask)[mi]) : e.mask[mi];
    acc *= act_grad_out(y, e.mask_act);
  }
  const size_t oi = e.transpose_out ? (size_t)col * e.ldo + row : (size_t)row * e.ldo + col;
  if (e.out_bf16) reinterpret_cast<__nv_bfloat16*>(e.out)[oi] = __float2bfloat16_rn(acc);
  else e.out[oi] = acc;
}


int launch_splitk_finish(const Epilogue& epi, int M, int N, int splits, cudaStream_t stream);
int launch_colsum(int M, int N, const float* x, int ld, float* out, void* ws, int64_t ws_bytes, cudaStream_t stream);

}  // namespace b200rl
