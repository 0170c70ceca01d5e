// Shared helpers for the b200rl CUDA library (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <string>

#include "b200rl.h"

namespace b200rl {

void set_error(const char* fmt, ...);
void count_launch();

#define B200RL_CUDA_OK(expr)                                                              \
  do {                                                                                    \
    cudaError_t _e = (expr);                                                              \
    if (_e != cudaSuccess) {                                                              \
      ::b200rl::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
      return B200RL_ECUDA;                                                                \
    }                                                                                     \
  } while (0)

#define B200RL_REQUIRE(cond, ...)          \
  do {                                     \
    if (!(cond)) {                         \
      ::b200rl::set_error(__VA_ARGS__);    \
      return B200RL_EINVAL;                \
    }                                      \
  } while (0)

#define B200RL_LAUNCH_OK()                                                                 \
  do {                                                                                     \
 ::b200rl::count_launch();                                                                \
    cudaError_t _e = cudaGetLastError();                                                   \
    if (_e != cudaSuccess) {                                                               \
      ::b200rl::set_error("kernel launch failed: %s (%s:%d)", cudaGetErrorString(_e), __FILE__, __LINE__); \
      return B200RL_ECUDA;                                                                 \
    }                                                                                      \
  } while (0)

constexpr int kFanout = 32;
constexpr int kMaxLevels = 8;
constexpr int kNumSMs = 148;

// Device view of the fan-out-32 sum tree: lvl[0] = root (1 float) .. lvl[L] = leaf weights (raw child
// values); pre[l] = sequential fp32 inclusive prefixes of every node's 32 children at level l.
struct TreeView {
  float* lvl[kMaxLevels];
  float* pre[kMaxLevels];
  int64_t width[kMaxLevels];
  int32_t L;
};

// Live key range of the item ring, kept in device memory so graph replays see fresh values.
struct ReplayState {
  unsigned long long item_head;  // next key to be issued
  unsigned long long item_tail;  // oldest live key
};

static inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }

template <typename T>
static inline T ceil_div(T a, T b) {
  return (a + b - 1) / b;
}

}  // namespace b200rl
