// Pieces shared by the fp32 SIMT GEMM (gemm_simt.cu) and the bf16 tcgen05 GEMM (gemm_tc.cu).
#pragma once
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
};

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
  if (e.bias) acc += e.bias[col];
  acc = apply_act(acc, e.act);
  if (e.mask) acc *= act_grad_out(e.mask[(size_t)row * e.ldmask + col], e.mask_act);
  if (e.transpose_out) e.out[(size_t)col * e.ldo + row] = acc;
  else e.out[(size_t)row * e.ldo + col] = acc;
}


int launch_splitk_finish(const Epilogue& epi, int M, int N, int splits, cudaStream_t stream);
int launch_colsum(int M, int N, const float* x, int ld, float* out, void* ws, int64_t ws_bytes, cudaStream_t stream);

}  // namespace b200rl
