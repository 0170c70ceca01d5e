// K1 (stratified prefix-sum sampling) and K2 (batched scatter priority update) over a fan-out-32
// fp32 sum tree that lives in HBM.
//
// Stands in for Reverb's Prioritized selector as configured at acme/agents/tf/dqn/agent.py:95-101
// (sampling) and for TFClient.update_priorities at acme/agents/tf/dqn/learning.py:151-154.
//
// Layout.  Level l (1..L, L = leaves) keeps two dense float arrays of the same width:
//   raw[l][32 g + j]  value of child j of node g (leaf level: the stored weight priority^alpha)
//   pre[l][32 g + j]  SEQUENTIAL fp32 inclusive prefix  p_j = fl(p_{j-1} + raw_j)  over the 32 children
// and raw[l-1][g] = pre[l][32 g + 31] (the root mass for l = 1).  A node's 32 prefixes are one
// 128-byte line.  Sequential prefixes are monotone, so "first child with target < prefix" is just the
// COUNT of prefixes <= target, and a child that is picked always has raw > 0:
//   sampling needs no scan at all -- a group of 4 lanes reads the line (32 bytes each), counts, takes
//   the largest prefix <= target as the exclusive prefix (two shuffles each) and descends; 8 samples per
//   warp, the top levels staged in shared memory once per CTA;
//   updates write leaves (last occurrence wins) and then recompute, level by level, the prefix line
//   of every touched node with ONE THREAD PER NODE (32 dependent adds, no atomics).
// The CPU oracle (oracle/sumtree.py) states the same arithmetic with np.cumsum(float32), so nodes,
// sampled indices and probabilities are bit-exact.
#include "common.cuh"

namespace b200rl {

// ------------------------------------------------------------------------------------------ K1
constexpr int kSampleThreads = 1024;

__device__ __forceinline__ void group_descend(const float* __restrict__ line, int gl, float& t, long long& node) {
  // line: the 32 prefixes of the current node; this lane owns entries 8*gl .. 8*gl+7
  const float4 a = *reinterpret_cast<const float4*>(line + gl * 8);
  const float4 b = *reinterpret_cast<const float4*>(line + gl * 8 + 4);
  const float p[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
  int cnt = 0;
  float mx = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const bool le = !(t < p[i]);       // prefix <= target
    cnt += le ? 1 : 0;
    mx = le ? fmaxf(mx, p[i]) : mx;    // prefixes are monotone: the largest one <= target is p[j-1]
  }
  cnt += __shfl_xor_sync(0xffffffffu, cnt, 1);
  cnt += __shfl_xor_sync(0xffffffffu, cnt, 2);
  mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
  mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
  int j = cnt;
  if (__any_sync(0xffffffffu, j == 32)) {
    // target >= node mass (fp rounding on the way down, or u*mass == mass): clamp to the last child
    // whose prefix increased, i.e. the last child with a non-empty subtree
    float prev = __shfl_up_sync(0xffffffffu, p[7], 1);
    if (gl == 0) prev = 0.f;
    int last = -1;
    float before = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (p[i] > prev) { last = gl * 8 + i; before = prev; }
      prev = p[i];
    }
#pragma unroll
    for (int d = 1; d <= 2; d <<= 1) {
      const int ol = __shfl_xor_sync(0xffffffffu, last, d);
      const float ob = __shfl_xor_sync(0xffffffffu, before, d);
      if (ol > last) { last = ol; before = ob; }
    }
    if (j == 32) { j = last < 0 ? 0 : last; mx = last < 0 ? 0.f : before; }
  }
  t = __fsub_rn(t, mx);
  node = node * kFanout + j;
}

template <int SPG>   // samples interleaved per 4-lane group (memory-level parallelism for large batches)
__global__ void __launch_bounds__(kSampleThreads)
sample_kernel(TreeView t, int S, const ReplayState* __restrict__ st, long long M, int B,
              const float* __restrict__ u, const long long* __restrict__ philox_counter, unsigned long long philox_seed,
              float* __restrict__ u_out, int stratified, float shards_f, const float* __restrict__ denom_dev,
              long long* __restrict__ idx, unsigned long long* __restrict__ keys,
              float* __restrict__ prob) {
  extern __shared__ __align__(16) float staged[];
  __shared__ int off[kMaxLevels];
  pdl_launch_dependents();     // K3 may be scheduled behind this grid (its own pdl_wait orders the data)
  if (threadIdx.x == 0) {
    int o = 0;
    for (int l = 1; l <= S; ++l) {
      off[l] = o;
      o += (int)t.width[l];
    }
  }
  __syncthreads();
  for (int l = 1; l <= S; ++l) {
    const float4* src = reinterpret_cast<const float4*>(t.pre[l]);
    float4* dst = reinterpret_cast<float4*>(staged + off[l]);
    for (int i = threadIdx.x; i < (int)(t.width[l] / 4); i += blockDim.x) dst[i] = src[i];
  }
  __syncthreads();

  const int gl = threadIdx.x & 3;
  const long long group_global = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 2;
  const long long total_groups = ((long long)gridDim.x * blockDim.x) >> 2;
  const float mass = t.lvl[0][0];
  // reported probability = weight / (R * M_r): the item's true inclusion probability when each of R shards draws the
  // same number of items; with a global mass installed (b200rl_replay_set_global_mass) = weight / sum_r M_r instead
  const float denom = denom_dev ? *denom_dev : __fmul_rn(shards_f, mass);
  const unsigned long long tail = st->item_tail;
  const unsigned long long key_base = tail - tail % (unsigned long long)M;
  const long long rounds = ((long long)B + total_groups * SPG - 1) / (total_groups * SPG);

  for (long long r = 0; r < rounds; ++r) {     // every lane runs every round: the shuffles are warp-wide
    float tg[SPG];
    long long node[SPG], b[SPG];
#pragma unroll
    for (int s = 0; s < SPG; ++s) {
      b[s] = (r * SPG + s) * total_groups + group_global;
      float ub = 0.f;
      if (b[s] < B) {
        if (philox_counter) {   // built-in draws: the same Philox stream as b200rl_uniform(seed, *counter)
          ub = philox_uniform((unsigned int)b[s], philox_seed, (unsigned long long)*philox_counter);
          if (u_out && gl == 0) u_out[b[s]] = ub;
        } else {
          ub = __ldg(u + b[s]);
        }
      }
      tg[s] = stratified ? __fmul_rn(__fdiv_rn(__fadd_rn((float)b[s], ub), (float)B), mass) : __fmul_rn(ub, mass);
      node[s] = 0;
    }
#pragma unroll 1
    for (int l = 1; l <= t.L; ++l) {
      const float* base = (l <= S) ? staged + off[l] : t.pre[l];
#pragma unroll
      for (int s = 0; s < SPG; ++s) group_descend(base + node[s] * kFanout, gl, tg[s], node[s]);
    }
    if (gl == 0) {
#pragma unroll
      for (int s = 0; s < SPG; ++s) {
        if (b[s] < B) {
          idx[b[s]] = node[s];
          if (keys) {
            unsigned long long k = key_base + (unsigned long long)node[s];
            if (k < tail) k += (unsigned long long)M;
            keys[b[s]] = k;
          }
          prob[b[s]] = __fdiv_rn(__ldg(t.lvl[t.L] + node[s]), denom);
        }
      }
    }
  }
}

int tree_staged_levels(const TreeView& t, int64_t budget_bytes) {
  int S = 0;
  int64_t used = 0;
  for (int l = 1; l <= t.L; ++l) {
    used += t.width[l] * 4;
    if (used > budget_bytes) break;
    S = l;
  }
  return S;
}

static int64_t staged_bytes(const TreeView& t, int S) {
  int64_t used = 0;
  for (int l = 1; l <= S; ++l) used += t.width[l] * 4;
  return used;
}

int tree_sample(const TreeView& t, const ReplayState* st_dev, int64_t M, int B, const float* u, const int64_t* philox_counter,
                uint64_t philox_seed, float* u_out, int stratified, int shard_count, const float* denom_dev, int64_t* idx,
                uint64_t* keys, float* prob, cudaStream_t stream, int* staged_out) {
  static bool attr_set = false;
  if (!attr_set) {
    B200RL_CUDA_OK(cudaFuncSetAttribute(sample_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    B200RL_CUDA_OK(cudaFuncSetAttribute(sample_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    attr_set = true;
  }
  const bool large = B >= 16384;
  // small batches: stage only what is nearly free (<= 16 KB); large batches amortise 160 KB per CTA
  int S = tree_staged_levels(t, large ? 160 * 1024 : 16 * 1024);
  size_t smem = (size_t)staged_bytes(t, S);
  if (staged_out) *staged_out = S;
  if (!large) {
    const int threads = 128;                                    // 32 samples per CTA
    int blocks = (int)ceil_div<int64_t>((int64_t)B * 4, threads);
    if (blocks > kNumSMs * 8) blocks = kNumSMs * 8;
    sample_kernel<1><<<blocks, threads, smem, stream>>>(t, S, st_dev, M, B, u, (const long long*)philox_counter, (unsigned long long)philox_seed, u_out, stratified,
                                                        (float)shard_count, denom_dev,
                                                        (long long*)idx, (unsigned long long*)keys, prob);
  } else {
    const bool fat = smem > 100 * 1024;   // one 1024-thread CTA per SM when the staged levels are big
    int threads = fat ? 1024 : 512;
    int blocks = kNumSMs * (fat ? 1 : 4);
    sample_kernel<2><<<blocks, threads, smem, stream>>>(t, S, st_dev, M, B, u, (const long long*)philox_counter, (unsigned long long)philox_seed, u_out, stratified,
                                                        (float)shard_count, denom_dev,
                                                        (long long*)idx, (unsigned long long*)keys, prob);
  }
  B200RL_LAUNCH_OK();
  return B200RL_OK;
}

// ------------------------------------------------------------------------------------------ K2
struct ScatterSrc {
  const unsigned long long* keys;  // non-null: entries come as (key, raw priority)
  const float* prio;
  const long long* pos;            // else: entries come as (position, weight); pos < 0 = skip
  const float* w;
  double alpha;
  long long M;
};

__device__ __forceinline__ float sanitize_weight(float w) { return (w > 0.f && w < 3.0e38f) ? w : 0.f; }

__device__ __forceinline__ void load_entry(const ScatterSrc& s, const ReplayState* st, int i,
                                           long long& pos, float& w) {
  if (s.keys) {
    unsigned long long k = s.keys[i];
    bool live = (k >= st->item_tail) && (k < st->item_head);
    pos = live ? (long long)(k % (unsigned long long)s.M) : -1;
    w = sanitize_weight((float)pow((double)s.prio[i], s.alpha));
  } else {
    pos = s.pos[i];
    w = sanitize_weight(s.w[i]);
  }
}

// One thread recomputes the prefix line of node g at level l from its 32 children and the node's own
// value one level up: p_j = fl(p_{j-1} + raw_j), sequential in fp32 (np.cumsum(float32) on the CPU).
__device__ __forceinline__ void recompute_node(const TreeView& t, int l, long long g) {
  const float4* src = reinterpret_cast<const float4*>(t.lvl[l] + g * kFanout);
  float4* dst = reinterpret_cast<float4*>(t.pre[l] + g * kFanout);
  float4 c[8];
#pragma unroll
  for (int q = 0; q < 8; ++q) c[q] = src[q];
  float acc = 0.f;
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    float4 o;
    acc = __fadd_rn(acc, c[q].x); o.x = acc;
    acc = __fadd_rn(acc, c[q].y); o.y = acc;
    acc = __fadd_rn(acc, c[q].z); o.z = acc;
    acc = __fadd_rn(acc, c[q].w); o.w = acc;
    dst[q] = o;
  }
  t.lvl[l - 1][l == 1 ? 0 : g] = acc;
}

// One CTA does the whole update for n <= 1024: leaf scatter with "last occurrence wins", then one
// pass per level (thread e recomputes the ancestor of entry e);
// only __syncthreads between levels, no atomics anywhere.
__global__ void __launch_bounds__(1024)
scatter_small_kernel(TreeView t, ScatterSrc src, const ReplayState* __restrict__ st, int n) {
  __shared__ int spos[1024];
  const int i = threadIdx.x;
  long long pos = -1;
  float w = 0.f;
  if (i < n) load_entry(src, st, i, pos, w);
  spos[i] = (int)pos;
  __syncthreads();
  if (pos >= 0) {
    bool winner = true;
    for (int j = i + 1; j < n; ++j) winner = winner && (spos[j] != (int)pos);
    if (winner) t.lvl[t.L][pos] = w;
  }
  __syncthreads();
  for (int l = t.L; l >= 1; --l) {
    const int shift = 5 * (t.L - l + 1);
    // entries that share a node recompute it redundantly and write identical values (cheaper than
    // electing one of them: the whole level is one load latency + 32 adds)
    if (pos >= 0) recompute_node(t, l, pos >> shift);
    __syncthreads();
  }
}

// Large batches.  Stage A: elect the last occurrence per position (one atomicMax per entry on a
// stamp array; this is the scatter's dedupe, the reduction below is atomic-free).
__global__ void scatter_stamp_kernel(ScatterSrc src, const ReplayState* __restrict__ st, int n,
                                     unsigned long long* __restrict__ stamp,
                                     const unsigned long long* __restrict__ epoch) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  long long pos;
  float w;
  load_entry(src, st, i, pos, w);
  if (pos >= 0) atomicMax(stamp + pos, (*epoch << 32) | (unsigned long long)(unsigned)i);
}

__global__ void scatter_leaf_kernel(TreeView t, ScatterSrc src, const ReplayState* __restrict__ st,
                                    int n, const unsigned long long* __restrict__ stamp,
                                    const unsigned long long* __restrict__ epoch) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  long long pos;
  float w;
  load_entry(src, st, i, pos, w);
  if (pos >= 0 && stamp[pos] == ((*epoch << 32) | (unsigned long long)(unsigned)i)) t.lvl[t.L][pos] = w;
}

// sparse level pass: thread e recomputes the ancestor (at level l-1) of entry e
__global__ void scatter_level_sparse_kernel(TreeView t, ScatterSrc src,
                                            const ReplayState* __restrict__ st, int n, int l) {
  int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n) return;
  long long pos;
  float w;
  load_entry(src, st, e, pos, w);
  if (pos < 0) return;
  recompute_node(t, l, pos >> (5 * (t.L - l + 1)));
}

// dense level pass: recompute every node of level l-1 from level l
__global__ void level_dense_kernel(TreeView t, int l, long long groups) {
  long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (; g < groups; g += stride) recompute_node(t, l, g);
}

__global__ void epoch_bump_kernel(unsigned long long* epoch) { *epoch += 1; }

static int launch_dense_level(const TreeView& t, int l, cudaStream_t stream) {
  long long groups = t.width[l] / kFanout;
  long long blocks = ceil_div<long long>(groups, 128);
  if (blocks > kNumSMs * 16) blocks = kNumSMs * 16;
  level_dense_kernel<<<(int)blocks, 128, 0, stream>>>(t, l, groups);
  B200RL_LAUNCH_OK();
  return B200RL_OK;
}

int tree_rebuild(const TreeView& t, cudaStream_t stream) {
  for (int l = t.L; l >= 1; --l) {
    int rc = launch_dense_level(t, l, stream);
    if (rc) return rc;
  }
  return B200RL_OK;
}

static int tree_scatter_impl(const TreeView& t, const ScatterSrc& src, const ReplayState* st_dev,
                             int n, unsigned long long* stamp, unsigned long long* epoch_dev,
                             cudaStream_t stream) {
  if (n <= 0) return B200RL_OK;
  if (n <= 1024) {
    int threads = ((n + 31) / 32) * 32;
    scatter_small_kernel<<<1, threads, 0, stream>>>(t, src, st_dev, n);
    B200RL_LAUNCH_OK();
    return B200RL_OK;
  }
  int blocks = ceil_div(n, 256);
  scatter_stamp_kernel<<<blocks, 256, 0, stream>>>(src, st_dev, n, stamp, epoch_dev);
  B200RL_LAUNCH_OK();
  scatter_leaf_kernel<<<blocks, 256, 0, stream>>>(t, src, st_dev, n, stamp, epoch_dev);
  B200RL_LAUNCH_OK();
  epoch_bump_kernel<<<1, 1, 0, stream>>>(epoch_dev);
  B200RL_LAUNCH_OK();
  for (int l = t.L; l >= 1; --l) {
    long long groups = t.width[l] / kFanout;
    if (groups <= (long long)n) {
      int rc = launch_dense_level(t, l, stream);
      if (rc) return rc;
    } else {
      scatter_level_sparse_kernel<<<ceil_div(n, 128), 128, 0, stream>>>(t, src, st_dev, n, l);
      B200RL_LAUNCH_OK();
    }
  }
  return B200RL_OK;
}

int tree_scatter_positions(const TreeView& t, int64_t M, int n, const int64_t* pos_dev,
                           const float* w_dev, const ReplayState* st_dev, unsigned long long* stamp,
                           unsigned long long* epoch_dev, cudaStream_t stream) {
  ScatterSrc src{nullptr, nullptr, (const long long*)pos_dev, w_dev, 1.0, (long long)M};
  return tree_scatter_impl(t, src, st_dev, n, stamp, epoch_dev, stream);
}

int tree_scatter_keys(const TreeView& t, int64_t M, int n, const uint64_t* keys_dev,
                      const float* prio_dev, double alpha, const ReplayState* st_dev,
                      unsigned long long* stamp, unsigned long long* epoch_dev, cudaStream_t stream) {
  ScatterSrc src{(const unsigned long long*)keys_dev, prio_dev, nullptr, nullptr, alpha, (long long)M};
  return tree_scatter_impl(t, src, st_dev, n, stamp, epoch_dev, stream);
}

}  // namespace b200rl
