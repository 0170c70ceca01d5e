// K1 (stratified prefix-sum sampling) and K2 (batched scatter priority update) over a fan-out-32
// fp32 sum tree that lives in HBM.
//
// Stands in for Reverb's Prioritized selector as configured at acme/agents/tf/dqn/agent.py:95-101
// (sampling) and for TFClient.update_priorities at acme/agents/tf/dqn/learning.py:151-154.
//
// Layout: level l (1..L) is a dense float array; the 32 children of node g are lvl[l][32g..32g+31]
// = one 128-byte line, so a warp fetches a node with ONE coalesced load, scans it with shuffles
// and descends.  The top levels are staged in shared memory once per CTA.  Every internal node is
// lane 31 of the Kogge-Stone scan of its children, so sampling and update share one summation
// order and the CPU oracle (oracle/sumtree.py) reproduces both bit for bit.
#include "common.cuh"

namespace b200rl {

// ------------------------------------------------------------------------------------------ K1
constexpr int kSampleThreads = 1024;

template <int SPW>
__global__ void __launch_bounds__(kSampleThreads)
sample_kernel(TreeView t, int S, const ReplayState* __restrict__ st, long long M, int B,
              const float* __restrict__ u, int stratified, float shards_f,
              long long* __restrict__ idx, unsigned long long* __restrict__ keys,
              float* __restrict__ prob) {
  extern __shared__ float staged[];
  __shared__ int off[kMaxLevels];
  if (threadIdx.x == 0) {
    int o = 0;
    for (int l = 1; l <= S; ++l) {
      off[l] = o;
      o += (int)t.width[l];
    }
  }
  __syncthreads();
  for (int l = 1; l <= S; ++l) {
    const float* src = t.lvl[l];
    float* dst = staged + off[l];
    for (int i = threadIdx.x; i < (int)t.width[l]; i += blockDim.x) dst[i] = src[i];
  }
  __syncthreads();

  const int lane = threadIdx.x & 31;
  const int warps_per_block = blockDim.x >> 5;
  const long long warp_global = (long long)blockIdx.x * warps_per_block + (threadIdx.x >> 5);
  const long long total_warps = (long long)gridDim.x * warps_per_block;
  const float mass = t.lvl[0][0];
  const float denom = __fmul_rn(shards_f, mass);
  const unsigned long long tail = st->item_tail;
  const unsigned long long key_base = tail - tail % (unsigned long long)M;

  for (long long base = warp_global * SPW; base < B; base += total_warps * SPW) {
    float tg[SPW];
    long long node[SPW];
    float leaf[SPW];
#pragma unroll
    for (int s = 0; s < SPW; ++s) {
      long long b = base + s;
      float ub = (b < B) ? __ldg(u + b) : 0.f;
      tg[s] = stratified ? __fmul_rn(__fdiv_rn(__fadd_rn((float)b, ub), (float)B), mass)
                         : __fmul_rn(ub, mass);
      node[s] = 0;
      leaf[s] = 0.f;
    }
#pragma unroll 1
    for (int l = 1; l <= t.L; ++l) {
      float c[SPW];
      if (l <= S) {
        const float* src = staged + off[l];
#pragma unroll
        for (int s = 0; s < SPW; ++s) c[s] = src[node[s] * kFanout + lane];
      } else {
        const float* src = t.lvl[l];
#pragma unroll
        for (int s = 0; s < SPW; ++s) c[s] = __ldg(src + node[s] * kFanout + lane);
      }
#pragma unroll
      for (int s = 0; s < SPW; ++s) {
        float p = warp_ks_scan(c[s], lane);
        unsigned ok = __ballot_sync(0xffffffffu, (tg[s] < p) && (c[s] > 0.f));
        int j;
        if (ok) {
          j = __ffs(ok) - 1;
        } else {
          unsigned nz = __ballot_sync(0xffffffffu, c[s] > 0.f);
          j = nz ? 31 - __clz(nz) : 0;
        }
        float pm1 = __shfl_sync(0xffffffffu, p, j > 0 ? j - 1 : 0);
        tg[s] = __fsub_rn(tg[s], j > 0 ? pm1 : 0.f);
        node[s] = node[s] * kFanout + j;
        leaf[s] = __shfl_sync(0xffffffffu, c[s], j);
      }
    }
    if (lane == 0) {
#pragma unroll
      for (int s = 0; s < SPW; ++s) {
        long long b = base + s;
        if (b < B) {
          idx[b] = node[s];
          if (keys) {
            unsigned long long k = key_base + (unsigned long long)node[s];
            if (k < tail) k += (unsigned long long)M;
            keys[b] = k;
          }
          prob[b] = __fdiv_rn(leaf[s], denom);
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

int tree_sample(const TreeView& t, const ReplayState* st_dev, int64_t M, int B, const float* u,
                int stratified, int shard_count, int64_t* idx, uint64_t* keys, float* prob,
                cudaStream_t stream, int* staged_out) {
  static bool attr_set = false;
  if (!attr_set) {
    B200RL_CUDA_OK(cudaFuncSetAttribute(sample_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    B200RL_CUDA_OK(cudaFuncSetAttribute(sample_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    attr_set = true;
  }
  const bool large = B >= 16384;
  // small batches: stage only what is nearly free (<= 16 KB); large batches amortise 160 KB per CTA
  int S = tree_staged_levels(t, large ? 160 * 1024 : 16 * 1024);
  size_t smem = (size_t)staged_bytes(t, S);
  if (staged_out) *staged_out = S;
  if (!large) {
    int warps = B;
    int blocks = (int)ceil_div<int64_t>(warps, 4);
    sample_kernel<1><<<blocks, 128, smem, stream>>>(t, S, st_dev, M, B, u, stratified,
                                                    (float)shard_count, (long long*)idx,
                                                    (unsigned long long*)keys, prob);
  } else {
    const bool fat = smem > 100 * 1024;  // one 1024-thread CTA per SM, else two 512-thread CTAs
    int threads = fat ? 1024 : 512;
    int blocks = kNumSMs * (fat ? 1 : 2);
    sample_kernel<4><<<blocks, threads, smem, stream>>>(
        t, S, st_dev, M, B, u, stratified, (float)shard_count, (long long*)idx,
        (unsigned long long*)keys, prob);
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

// recompute the parent of group g at level l: parent = lane 31 of the scan of its 32 children
__device__ __forceinline__ void recompute_parent(const TreeView& t, int l, long long g, int lane) {
  float c = t.lvl[l][g * kFanout + lane];
  float p = warp_ks_scan(c, lane);
  if (lane == 31) t.lvl[l - 1][l == 1 ? 0 : g] = p;
}

// One CTA does the whole update for n <= 1024: leaf scatter with "last occurrence wins", then one
// pass per level; only __syncthreads between levels, no atomics anywhere.
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
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  for (int l = t.L; l >= 1; --l) {
    const int shift = 5 * (t.L - l + 1);
    // each warp owns entries warp, warp + nwarps, ...; the (independent) node loads of up to 8 entries
    // are issued before the first scan so that their latencies overlap
    for (int e0 = warp; e0 < n; e0 += nwarps * 8) {
      float c[8];
      long long g[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int e = e0 + j * nwarps;
        const int p = e < n ? spos[e] : -1;
        g[j] = p < 0 ? -1 : ((long long)p >> shift);
        c[j] = g[j] < 0 ? 0.f : t.lvl[l][g[j] * kFanout + lane];
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        if (g[j] < 0) continue;
        float p = warp_ks_scan(c[j], lane);
        if (lane == 31) t.lvl[l - 1][l == 1 ? 0 : g[j]] = p;
      }
    }
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

// sparse level pass: one warp per entry recomputes that entry's ancestor at level l-1
__global__ void scatter_level_sparse_kernel(TreeView t, ScatterSrc src,
                                            const ReplayState* __restrict__ st, int n, int l) {
  const int lane = threadIdx.x & 31;
  long long e = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (e >= n) return;
  long long pos;
  float w;
  load_entry(src, st, (int)e, pos, w);
  if (pos < 0) return;
  recompute_parent(t, l, pos >> (5 * (t.L - l + 1)), lane);
}

// dense level pass: recompute every node of level l-1 from level l
__global__ void level_dense_kernel(TreeView t, int l, long long groups) {
  const int lane = threadIdx.x & 31;
  long long g = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long stride = ((long long)gridDim.x * blockDim.x) >> 5;
  for (; g < groups; g += stride) recompute_parent(t, l, g, lane);
}

__global__ void epoch_bump_kernel(unsigned long long* epoch) { *epoch += 1; }

static int launch_dense_level(const TreeView& t, int l, cudaStream_t stream) {
  long long groups = t.width[l] / kFanout;
  long long blocks = ceil_div<long long>(groups * 32, 256);
  if (blocks > kNumSMs * 8) blocks = kNumSMs * 8;
  level_dense_kernel<<<(int)blocks, 256, 0, stream>>>(t, l, groups);
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
    scatter_small_kernel<<<1, 1024, 0, stream>>>(t, src, st_dev, n);   // 32 warps share the per-level node recomputes
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
      long long b = ceil_div<long long>((long long)n * 32, 256);
      scatter_level_sparse_kernel<<<(int)b, 256, 0, stream>>>(t, src, st_dev, n, l);
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
