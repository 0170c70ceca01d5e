// DQfD demonstration mixing (`acme/agents/tf/dqfd/agent.py:111-122,160-219`).
//
// The reference draws every batch element from the replay dataset with probability 1 - ratio and from a stream of
// demonstration transitions with probability ratio (`tf.data.experimental.sample_from_datasets`); a demonstration
// transition is an n-step window starting at a uniformly drawn step of a demonstration episode, built by
// `_n_step_transition_from_episode` with ITS arithmetic (not the adder's):
//   first ~ U{0 .. max_index - 2},  last = min(first + n, max_index),  max_index = L - 1
//   discounts = [1, cumprod(d[first : last-1])] * g^[0 .. last-first-1]      (fp32; g^j by pow)
//   r = sum(rewards[first+1 : last+1] * discounts),  d = discounts[-1]
//   key = 0 (names no item; here ~0, keys count from 0), probability = 1.0
// Here the demonstration episodes live in HBM as flat step arrays; after K1 / K3 have filled the batch from replay, this
// kernel overwrites the rows whose first uniform is below `ratio` -- graph-capturable, no host decision.  One CTA per
// batch row; rows that stay replay rows cost nothing but the early exit.
#include "common.cuh"

namespace b200rl {

__global__ void __launch_bounds__(128)
demo_mix_kernel(int B, const uint8_t* __restrict__ obs, int obs_bytes, const uint8_t* __restrict__ act, int act_bytes,
                const float* __restrict__ rew, const float* __restrict__ disc, const long long* __restrict__ ep_off, int E,
                int n_step, float gamma, const float* __restrict__ u3, float ratio, uint8_t* __restrict__ o_tm1,
                uint8_t* __restrict__ a_tm1, float* __restrict__ R, float* __restrict__ D, uint8_t* __restrict__ o_t,
                unsigned long long* __restrict__ keys, float* __restrict__ prob, int* __restrict__ is_demo) {
  const int b = blockIdx.x;
  const bool demo = u3[3 * b] < ratio;
  if (threadIdx.x == 0 && is_demo) is_demo[b] = demo ? 1 : 0;
  if (!demo) return;
  int e = (int)(u3[3 * b + 1] * (float)E);
  e = min(max(e, 0), E - 1);
  const long long base = ep_off[e];
  const int L = (int)(ep_off[e + 1] - base);
  const int max_index = L - 1;
  int first = (int)(u3[3 * b + 2] * (float)(max_index - 1));     // tf.random.uniform(minval=0, maxval=max_index-1)
  first = min(max(first, 0), max_index - 2);
  const int last = min(first + n_step, max_index);
  const uint8_t* s0 = obs + (size_t)(base + first) * obs_bytes;
  const uint8_t* s1 = obs + (size_t)(base + last) * obs_bytes;
  uint8_t* d0 = o_tm1 + (size_t)b * obs_bytes;
  uint8_t* d1 = o_t + (size_t)b * obs_bytes;
  for (int i = threadIdx.x; i < obs_bytes; i += blockDim.x) { d0[i] = s0[i]; d1[i] = s1[i]; }
  const uint8_t* as = act + (size_t)(base + first) * act_bytes;
  for (int i = threadIdx.x; i < act_bytes; i += blockDim.x) a_tm1[(size_t)b * act_bytes + i] = as[i];
  if (threadIdx.x == 0) {
    const int m = last - first;
    float c = 1.f, r = 0.f, dj = 1.f;
    for (int j = 0; j < m; ++j) {
      if (j > 0) c = __fmul_rn(c, disc[base + first + j - 1]);          // cumprod(discounts[first : last-1])
      dj = __fmul_rn(c, (float)pow((double)gamma, (double)j));            // * g^j (correctly rounded fp32 power)
      r = __fadd_rn(r, __fmul_rn(rew[base + first + 1 + j], dj));         // reduce_sum in index order
    }
    R[b] = r;
    D[b] = dj;
    keys[b] = ~0ull;   // the reference's key 0 names no Reverb item; here keys count from 0, so "no item" is ~0
    prob[b] = 1.f;
  }
}

}  // namespace b200rl

using namespace b200rl;

extern "C" int b200rl_demo_mix(int32_t B, const void* obs, int32_t obs_bytes, const void* act, int32_t act_bytes,
                               const float* rew, const float* disc, const int64_t* episode_offsets, int32_t num_episodes,
                               int32_t n_step, float gamma, const float* uniforms3, float ratio, void* o_tm1, void* a_tm1,
                               float* R, float* D, void* o_t, uint64_t* keys, float* prob, int32_t* is_demo, void* stream) {
  B200RL_REQUIRE(obs && act && rew && disc && episode_offsets && uniforms3 && o_tm1 && a_tm1 && R && D && o_t && keys && prob,
                 "null argument");
  B200RL_REQUIRE(B >= 1 && num_episodes >= 1 && n_step >= 1 && obs_bytes >= 1 && act_bytes >= 1, "bad shape");
  demo_mix_kernel<<<B, 128, 0, as_stream(stream)>>>(B, (const uint8_t*)obs, obs_bytes, (const uint8_t*)act, act_bytes, rew, disc,
                                                   (const long long*)episode_offsets, num_episodes, n_step, gamma, uniforms3,
                                                   ratio, (uint8_t*)o_tm1, (uint8_t*)a_tm1, R, D, (uint8_t*)o_t,
                                                   (unsigned long long*)keys, prob, is_demo);
  B200RL_LAUNCH_OK();
  return B200RL_OK;
}
