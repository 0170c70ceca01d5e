// Data-parallel learner exchange over NVLink / NVSwitch peer memory (include/b200rl.h "data parallel").
//
// The reference's multi-replica learner does all_reduce('mean', grads) -> optimizer.apply on every replica
// (`acme/agents/tf/crr/recurrent_learning.py:346-359`).  Here the collective and the optimizer are ONE kernel per
// gradient bucket:
//
//   barrier      every rank tells every peer (a flag in the peer's memory) that its gradients of this step are final
//   reduce       rank r owns shard r of the bucket: it reads that shard of ALL ranks' gradient buffers straight
//                through NVLink (peer loads), adds them in rank order and scales by 1/R            [reduce-scatter]
//   Adam         ... updates its shard of the parameters and moments (only the owner keeps moments for a shard)
//   broadcast    ... and stores the new parameters into EVERY rank's parameter buffer (peer stores)   [all-gather]
//   barrier      the kernel ends when every rank has finished its stores, so the next forward sees them
//
// so a step moves 2 x (R-1)/R x bytes per rank like a ring all-reduce, the optimizer work drops to 1/R per rank,
// no rank ever waits for a host, and the replicas are bit-identical by construction (one writer per parameter).
// The scalar all-reduce(MAX) of the importance-weight normaliser uses the same flag mechanism.
//
// Memory: one cudaMalloc region per rank [params | grads | mailbox], exported with cudaIpcGetMemHandle and opened by
// the peers (one process per GPU).  Flags are volatile (system-coherent) words, bulk data uses L1-bypassing accesses
// ordered by __threadfence_system(); flags carry the learner's step number, so the kernels are CUDA-graph capturable.
// Spin loops give up after a few seconds and raise the region's error word instead of hanging the GPU.
#include <algorithm>
#include <cstdlib>
#include <cstring>

#include <cuda_bf16.h>

#include "common.cuh"

namespace b200rl {

constexpr int kMaxRanks = 8;    // one NVSwitch box
constexpr int kBuckets = 4;

struct Mailbox {   // lives at the end of every rank's region; written by peers
  volatile long long grads_ready[kBuckets][kMaxRanks];   // [bucket][src rank] = epoch
  volatile long long params_done[kBuckets][kMaxRanks];
  volatile long long max_epoch[2][kMaxRanks];            // parity double-buffered scalar exchange
  volatile double max_value[2][kMaxRanks];
  unsigned int cta_done[kBuckets];                       // local: CTAs that finished their shard part
  volatile int error;                                    // 1 = a spin loop timed out
  long long dbg[8];                                      // tools/dp_bench.py: global-timer stamps of the last exchange
};

struct DpPeers {
  float* params[kMaxRanks];
  float* grads[kMaxRanks];
  Mailbox* mail[kMaxRanks];
};
constexpr int kCeStreams = 8;

struct DpState {
  int world, rank, device;
  int64_t n;
  void* region;              // local allocation
  size_t region_bytes;
  void* opened[kMaxRanks];   // peer regions from cudaIpcOpenMemHandle
  DpPeers peers;
  // copy-engine exchange: landing[j] = rank j's landing buffer ([world][chunk] floats inside its region): every rank
  // PUSHES the shard rank j owns of its own gradients there (DMA writes over NVLink run faster than DMA reads)
  float* landing[kMaxRanks];
  size_t landing_floats;
  bool external;             // region and peer mappings are owned by the caller (b200rl_dp_create_external)
  float* mc_params;          // multicast (NVSwitch) mappings of every rank's params / grads, or NULL
  float* mc_grads;
  cudaStream_t ce_stream[kCeStreams];   // the DMA copies of one phase are spread over several copy engines
  cudaEvent_t ce_fork[2], ce_join[2][kCeStreams];
};

static size_t params_bytes(int64_t n) { return ((size_t)n * 4 + 255) & ~(size_t)255; }
// landing buffer: a bucket [off, off + n) lands at off + bucket * 4 * kMaxRanks as R shards of ceil(n / R) floats rounded
// up to 4, so disjoint buckets never share landing space (two may be in flight at once)
static size_t landing_bytes(int64_t n) { return params_bytes(n + 4 * kMaxRanks * (kBuckets + 1)); }
static size_t landing_base(int64_t off, int bucket) { return (size_t)off + (size_t)bucket * 4 * kMaxRanks; }

// Bulk data moves with L1-bypassing (.cg) accesses: peer memory is only ever cached at its owner's L2, so these are
// coherent once the flag protocol (volatile flags + __threadfence_system on both sides) has ordered them, and -- unlike
// volatile accesses, which a thread issues one at a time -- they pipeline.
__device__ __forceinline__ long long gtime() {
  long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
__device__ __forceinline__ float4 ld_sys4(const float* p) {
  float4 v;
  asm volatile("ld.global.cg.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_sys4(float* p, float4 v) {
  asm volatile("st.global.cg.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
// waits until flag[j] >= epoch for every rank j; false on timeout
__device__ __forceinline__ bool wait_all(const volatile long long* flag, int world, long long epoch, volatile int* error) {
  for (int j = 0; j < world; ++j) {
    long long spins = 0;
    while (flag[j] < epoch) {
      __nanosleep(64);
      if (++spins > (1ll << 26)) { *error = 1; return false; }   // >= 4 s
    }
  }
  return true;
}

struct AdamC { float bc1, bc2, b1, b2, omb1, omb2, k1, lr, eps, gs; int eps_mode; };
__device__ __forceinline__ void adam1(float& p, float g, float& m, float& v, const AdamC& c) {
  const float gj = __fmul_rn(g, c.gs);
  m = __fadd_rn(__fmul_rn(c.b1, m), __fmul_rn(c.omb1, gj));
  v = __fadd_rn(__fmul_rn(c.b2, v), __fmul_rn(c.omb2, __fmul_rn(gj, gj)));
  m = fabsf(m) < 1.17549435e-38f ? 0.f : m;   // subnormal moments: see adam_one in learner_math.cu
  v = v < 1.17549435e-38f ? 0.f : v;
  // exact zeros: harmless substitute operands + selects instead of the IEEE slow paths (see adam_one)
  const bool vz = v == 0.f, mz = m == 0.f;
  const float v_in = vz ? 1.f : v, m_in = mz ? 1.f : m;
  float upd;
  if (c.eps_mode == 0) {
    const float root = vz ? 0.f : __fsqrt_rn(__fdiv_rn(v_in, c.bc2));
    upd = __fdiv_rn(__fdiv_rn(m_in, c.bc1), __fadd_rn(root, c.eps));
  } else {
    const float root = vz ? 0.f : __fsqrt_rn(v_in);
    upd = __fdiv_rn(__fmul_rn(c.k1, m_in), __fadd_rn(root, c.eps));
  }
  upd = mz ? 0.f : upd;
  p = __fsub_rn(p, __fmul_rn(c.lr, upd));
}

// One launch per gradient bucket [off, off + n).  The shard of rank r is [off + r * chunk, off + (r + 1) * chunk)
// with chunk a multiple of 4 floats.  (Same arithmetic as adam_kernel in learner_math.cu with gscale = 1/R.)
// W = compile-time bound on the world size, U = float4 groups per thread per iteration: all W * U peer loads of an
// iteration are issued before the first is consumed (NVLink round trips are microseconds; bytes in flight are
// what buys bandwidth).
template <int W, int U>
__global__ void __launch_bounds__(256)
dp_adam_kernel(DpPeers peers, int world, int rank, long long off, long long n, long long chunk, float* __restrict__ m,
               float* __restrict__ v, const long long* __restrict__ step_dev, float lr, double b1, double b2, float eps,
               int eps_mode, int bucket, int final_barrier) {
  Mailbox* mine = peers.mail[rank];
  const long long epoch = *step_dev + 1;
  __shared__ AdamC c;
  __shared__ bool ok;
  if (threadIdx.x == 0) {
    if (blockIdx.x == 0) {
      mine->dbg[0] = gtime();
      // this rank's gradients are final (previous kernels of the stream): tell every rank, including ourselves
      __threadfence_system();
      for (int j = 0; j < world; ++j) peers.mail[j]->grads_ready[bucket][rank] = epoch;
    }
    const double t = (double)epoch;
    c.bc1 = (float)(1.0 - pow(b1, t)); c.bc2 = (float)(1.0 - pow(b2, t));
    c.b1 = (float)b1; c.b2 = (float)b2; c.omb1 = (float)(1.0 - b1); c.omb2 = (float)(1.0 - b2);
    c.k1 = sqrtf(c.bc2) / c.bc1; c.lr = lr; c.eps = eps; c.gs = 1.f / (float)world; c.eps_mode = eps_mode;
    ok = wait_all(mine->grads_ready[bucket], world, epoch, &mine->error);
    __threadfence_system();
    if (blockIdx.x == 0) mine->dbg[1] = gtime();
  }
  __syncthreads();
  if (ok) {
    const long long s0 = off + (long long)rank * chunk;
    const long long s1 = min(off + n, s0 + chunk);
    const long long s1v = s0 + ((s1 - s0) & ~3ll);                     // whole float4 groups
    const long long tile = (long long)blockDim.x * 4 * U;              // floats per CTA per iteration
    const long long step_f = (long long)gridDim.x * tile;
    // software pipeline: the peer loads of iteration k+1 are in flight while iteration k does its Adam update and its
    // peer stores, so NVLink carries reads and writes at the same time
    float4 raw[W][U];
    auto issue = [&](long long base) {
#pragma unroll
      for (int j = 0; j < W; ++j)
        if (j < world) {
#pragma unroll
          for (int u = 0; u < U; ++u) {
            const long long i = base + ((long long)u * blockDim.x + threadIdx.x) * 4;
            if (i < s1v) raw[j][u] = ld_sys4(peers.grads[j] + i);
          }
        }
    };
    long long base = s0 + (long long)blockIdx.x * tile;
    if (base < s1v) issue(base);
    for (; base < s1v; base += step_f) {
      float4 g[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        g[u] = raw[0][u];
#pragma unroll
        for (int j = 1; j < W; ++j)            // fixed rank order: deterministic
          if (j < world) {
            g[u].x = __fadd_rn(g[u].x, raw[j][u].x); g[u].y = __fadd_rn(g[u].y, raw[j][u].y);
            g[u].z = __fadd_rn(g[u].z, raw[j][u].z); g[u].w = __fadd_rn(g[u].w, raw[j][u].w);
          }
      }
      if (base + step_f < s1v) issue(base + step_f);
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const long long i = base + ((long long)u * blockDim.x + threadIdx.x) * 4;
        if (i >= s1v) continue;
        float4 p = *reinterpret_cast<const float4*>(peers.params[rank] + i);
        float4 mm = *reinterpret_cast<const float4*>(m + i), vv = *reinterpret_cast<const float4*>(v + i);
        adam1(p.x, g[u].x, mm.x, vv.x, c); adam1(p.y, g[u].y, mm.y, vv.y, c);
        adam1(p.z, g[u].z, mm.z, vv.z, c); adam1(p.w, g[u].w, mm.w, vv.w, c);
        *reinterpret_cast<float4*>(m + i) = mm;
        *reinterpret_cast<float4*>(v + i) = vv;
#pragma unroll
        for (int j = 0; j < W; ++j)
          if (j < world) st_sys4(peers.params[j] + i, p);
      }
    }
    if (blockIdx.x == 0 && threadIdx.x < (int)(s1 - s1v)) {   // ragged end of the bucket (< 4 floats)
      const long long k = s1v + threadIdx.x;
      float g = *(volatile float*)(peers.grads[0] + k);
      for (int j = 1; j < world; ++j) g = __fadd_rn(g, *(volatile float*)(peers.grads[j] + k));
      float p = peers.params[rank][k], mm = m[k], vv = v[k];
      adam1(p, g, mm, vv, c);
      m[k] = mm; v[k] = vv;
      for (int j = 0; j < world; ++j) *(volatile float*)(peers.params[j] + k) = p;
    }
  }
  // the last CTA of this rank to finish announces it to every rank, then (unless a later exchange of the same step
  // does it) waits for everybody's announcement: when the kernel ends, all parameters are in place everywhere
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned int done = atomicAdd(&mine->cta_done[bucket], 1u);
    if (done == gridDim.x - 1) {
      mine->cta_done[bucket] = 0;
      mine->dbg[2] = gtime();
      __threadfence_system();
      for (int j = 0; j < world; ++j) peers.mail[j]->params_done[bucket][rank] = epoch;
      if (final_barrier) wait_all(mine->params_done[bucket], world, epoch, &mine->error);
      __threadfence_system();
      mine->dbg[3] = gtime();
    }
  }
}

// ---- Copy-engine form of the exchange.  The SM-issued kernel above needs ~450 resident CTAs to keep NVLink busy, and
// while they are resident the latency-bound GEMM / convolution kernels of the step cannot get their SM slots (measured
// on 2 GPUs: the target forward stalls from 15 to 94 us while the exchange runs).  Here the bulk bytes move by DMA
// (cudaMemcpyAsync between the peer-mapped regions: graph memcpy nodes, no SM), and the SMs only run
//   dp_flag_kernel       one warp: announce (grads_ready / params_done) to every peer, optionally wait for all peers
//   dp_adam_local_kernel the owner's shard: sum of R gradient shards in rank order (own one in place, the others from the
//                        landing buffer), x 1/R, Adam, bf16 shadow -- the same arithmetic, in the same order, as
//                        dp_adam_kernel, so both forms produce bit-identical parameters.
__global__ void dp_flag_kernel(DpPeers peers, int world, int rank, const long long* __restrict__ step_dev, int bucket,
                               int which, int wait) {
  if (threadIdx.x != 0) return;
  Mailbox* mine = peers.mail[rank];
  const long long epoch = *step_dev + 1;
  __threadfence_system();
  for (int j = 0; j < world; ++j) {
    if (which == 0) peers.mail[j]->grads_ready[bucket][rank] = epoch;
    else peers.mail[j]->params_done[bucket][rank] = epoch;
  }
  if (wait) wait_all(which == 0 ? mine->grads_ready[bucket] : mine->params_done[bucket], world, epoch, &mine->error);
  __threadfence_system();
}

template <int W>
__global__ void __launch_bounds__(256)
dp_adam_local_kernel(float* __restrict__ params, const float* __restrict__ grads, const float* __restrict__ landing,
                     int world, int rank, long long s0, long long len, long long chunk, float* __restrict__ m,
                     float* __restrict__ v, __nv_bfloat16* __restrict__ shadow, const long long* __restrict__ step_dev,
                     float lr, double b1, double b2, float eps, int eps_mode) {
  __shared__ AdamC c;
  if (threadIdx.x == 0) {
    const double t = (double)(*step_dev + 1);
    c.bc1 = (float)(1.0 - pow(b1, t)); c.bc2 = (float)(1.0 - pow(b2, t));
    c.b1 = (float)b1; c.b2 = (float)b2; c.omb1 = (float)(1.0 - b1); c.omb2 = (float)(1.0 - b2);
    c.k1 = sqrtf(c.bc2) / c.bc1; c.lr = lr; c.eps = eps; c.gs = 1.f / (float)world; c.eps_mode = eps_mode;
  }
  __syncthreads();
  const long long lenv = len & ~3ll;
  for (long long k = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 4; k < lenv; k += (long long)gridDim.x * blockDim.x * 4) {
    const long long i = s0 + k;
    float4 g;
#pragma unroll
    for (int j = 0; j < W; ++j)
      if (j < world) {
        const float4 x = j == rank ? *reinterpret_cast<const float4*>(grads + i)
                                   : *reinterpret_cast<const float4*>(landing + (long long)j * chunk + k);
        if (j == 0) g = x;
        else { g.x = __fadd_rn(g.x, x.x); g.y = __fadd_rn(g.y, x.y); g.z = __fadd_rn(g.z, x.z); g.w = __fadd_rn(g.w, x.w); }
      }
    float4 p = *reinterpret_cast<const float4*>(params + i);
    float4 mm = *reinterpret_cast<const float4*>(m + i), vv = *reinterpret_cast<const float4*>(v + i);
    adam1(p.x, g.x, mm.x, vv.x, c); adam1(p.y, g.y, mm.y, vv.y, c);
    adam1(p.z, g.z, mm.z, vv.z, c); adam1(p.w, g.w, mm.w, vv.w, c);
    *reinterpret_cast<float4*>(m + i) = mm;
    *reinterpret_cast<float4*>(v + i) = vv;
    *reinterpret_cast<float4*>(params + i) = p;
    if (shadow) {
      const __nv_bfloat162 lo = __floats2bfloat162_rn(p.x, p.y), hi = __floats2bfloat162_rn(p.z, p.w);
      uint2 o;
      o.x = *reinterpret_cast<const uint32_t*>(&lo); o.y = *reinterpret_cast<const uint32_t*>(&hi);
      *reinterpret_cast<uint2*>(shadow + i) = o;
    }
  }
  if (blockIdx.x == 0 && threadIdx.x < (int)(len - lenv)) {   // ragged end of the bucket (< 4 floats)
    const long long k = lenv + threadIdx.x, i = s0 + k;
    float g = 0.f;
    for (int j = 0; j < world; ++j) {
      const float x = j == rank ? grads[i] : landing[(long long)j * chunk + k];
      g = j == 0 ? x : __fadd_rn(g, x);
    }
    float p = params[i], mm = m[i], vv = v[i];
    adam1(p, g, mm, vv, c);
    m[i] = mm; v[i] = vv; params[i] = p;
    if (shadow) shadow[i] = __float2bfloat16_rn(p);
  }
}

// ---- NVSwitch multicast form (multimem.*): the reduce-scatter is ONE load per 16 bytes of the owned shard -- the
// switch fetches the operand from every rank's gradient buffer and adds them on the way -- and the all-gather ONE
// store that the switch replicates into every rank's parameter buffer.  Per rank and bucket of n floats the SMs move
// 2 x n/R x 4 bytes instead of 2 x (R-1)/R x n x 4, and the links carry n x 4 x (R+1)/R... see DESIGN.md §6: 32 MB per
// direction instead of 56 MB at R = 8.  The kernel needs few resident warps, so it runs underneath the convolution
// backward without taking the slots those kernels need.  The in-switch sum is not the rank-order sum of dp_adam_kernel
// (an ulp-level difference in the mean gradient); replicas stay bit-identical: one writer per parameter.
__device__ __forceinline__ float4 mc_ld_reduce4(const float* mc) {
  float4 v;
  asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(mc) : "memory");
  return v;
}
__device__ __forceinline__ void mc_st4(float* mc, float4 v) {
  asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};"
               :: "l"(mc), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

template <int U>
__global__ void __launch_bounds__(256)
dp_mc_adam_kernel(DpPeers peers, const float* __restrict__ mc_grads, float* __restrict__ mc_params, int world, int rank,
                  long long off, long long n, long long chunk, float* __restrict__ m, float* __restrict__ v,
                  const long long* __restrict__ step_dev, float lr, double b1, double b2, float eps, int eps_mode,
                  int bucket, int final_barrier) {
  Mailbox* mine = peers.mail[rank];
  const long long epoch = *step_dev + 1;
  __shared__ AdamC c;
  __shared__ bool ok;
  if (threadIdx.x == 0) {
    if (blockIdx.x == 0) {
      mine->dbg[0] = gtime();
      __threadfence_system();
      for (int j = 0; j < world; ++j) peers.mail[j]->grads_ready[bucket][rank] = epoch;
    }
    const double t = (double)epoch;
    c.bc1 = (float)(1.0 - pow(b1, t)); c.bc2 = (float)(1.0 - pow(b2, t));
    c.b1 = (float)b1; c.b2 = (float)b2; c.omb1 = (float)(1.0 - b1); c.omb2 = (float)(1.0 - b2);
    c.k1 = sqrtf(c.bc2) / c.bc1; c.lr = lr; c.eps = eps; c.gs = 1.f / (float)world; c.eps_mode = eps_mode;
    ok = wait_all(mine->grads_ready[bucket], world, epoch, &mine->error);
    __threadfence_system();
    if (blockIdx.x == 0) mine->dbg[1] = gtime();
  }
  __syncthreads();
  if (ok) {
    const long long s0 = off + (long long)rank * chunk;
    const long long s1 = min(off + n, s0 + chunk);
    const long long s1v = s0 + ((s1 - s0) & ~3ll);
    const long long tile = (long long)blockDim.x * 4 * U;
    const long long step_f = (long long)gridDim.x * tile;
    float* params = peers.params[rank];
    for (long long base = s0 + (long long)blockIdx.x * tile; base < s1v; base += step_f) {
      float4 g[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {   // all U in-switch reductions of the iteration in flight before the first is used
        const long long i = base + ((long long)u * blockDim.x + threadIdx.x) * 4;
        if (i < s1v) g[u] = mc_ld_reduce4(mc_grads + i);
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const long long i = base + ((long long)u * blockDim.x + threadIdx.x) * 4;
        if (i >= s1v) continue;
        float4 p = *reinterpret_cast<const float4*>(params + i);
        float4 mm = *reinterpret_cast<const float4*>(m + i), vv = *reinterpret_cast<const float4*>(v + i);
        adam1(p.x, g[u].x, mm.x, vv.x, c); adam1(p.y, g[u].y, mm.y, vv.y, c);
        adam1(p.z, g[u].z, mm.z, vv.z, c); adam1(p.w, g[u].w, mm.w, vv.w, c);
        *reinterpret_cast<float4*>(m + i) = mm;
        *reinterpret_cast<float4*>(v + i) = vv;
        mc_st4(mc_params + i, p);     // lands in every rank's parameter buffer, this one's included
      }
    }
    if (blockIdx.x == 0 && threadIdx.x < (int)(s1 - s1v)) {   // ragged end of the bucket (< 4 floats): plain peer accesses
      const long long k = s1v + threadIdx.x;
      float g = *(volatile float*)(peers.grads[0] + k);
      for (int j = 1; j < world; ++j) g = __fadd_rn(g, *(volatile float*)(peers.grads[j] + k));
      float p = params[k], mm = m[k], vv = v[k];
      adam1(p, g, mm, vv, c);
      m[k] = mm; v[k] = vv;
      for (int j = 0; j < world; ++j) *(volatile float*)(peers.params[j] + k) = p;
    }
  }
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned int done = atomicAdd(&mine->cta_done[bucket], 1u);
    if (done == gridDim.x - 1) {
      mine->cta_done[bucket] = 0;
      mine->dbg[2] = gtime();
      __threadfence_system();
      for (int j = 0; j < world; ++j) peers.mail[j]->params_done[bucket][rank] = epoch;
      if (final_barrier) wait_all(mine->params_done[bucket], world, epoch, &mine->error);
      __threadfence_system();
      mine->dbg[3] = gtime();
    }
  }
}

// The two halves on their own (hybrids with the copy-engine form): reduce + Adam with the in-switch reduction, new
// parameters written to this rank's buffer only (+ bf16 shadow) ...
template <int U>
__global__ void __launch_bounds__(256)
dp_mc_reduce_adam_kernel(DpPeers peers, const float* __restrict__ mc_grads, int world, int rank, long long off, long long n,
                         long long chunk, float* __restrict__ m, float* __restrict__ v, __nv_bfloat16* __restrict__ shadow,
                         const long long* __restrict__ step_dev, float lr, double b1, double b2, float eps, int eps_mode,
                         int bucket) {
  Mailbox* mine = peers.mail[rank];
  const long long epoch = *step_dev + 1;
  __shared__ AdamC c;
  __shared__ bool ok;
  if (threadIdx.x == 0) {
    if (blockIdx.x == 0) {
      mine->dbg[0] = gtime();
      __threadfence_system();
      for (int j = 0; j < world; ++j) peers.mail[j]->grads_ready[bucket][rank] = epoch;
    }
    const double t = (double)epoch;
    c.bc1 = (float)(1.0 - pow(b1, t)); c.bc2 = (float)(1.0 - pow(b2, t));
    c.b1 = (float)b1; c.b2 = (float)b2; c.omb1 = (float)(1.0 - b1); c.omb2 = (float)(1.0 - b2);
    c.k1 = sqrtf(c.bc2) / c.bc1; c.lr = lr; c.eps = eps; c.gs = 1.f / (float)world; c.eps_mode = eps_mode;
    ok = wait_all(mine->grads_ready[bucket], world, epoch, &mine->error);
    __threadfence_system();
    if (blockIdx.x == 0) mine->dbg[1] = gtime();
  }
  __syncthreads();
  if (!ok) return;
  const long long s0 = off + (long long)rank * chunk;
  const long long s1 = min(off + n, s0 + chunk);
  const long long s1v = s0 + ((s1 - s0) & ~3ll);
  const long long tile = (long long)blockDim.x * 4 * U;
  float* params = peers.params[rank];
  for (long long base = s0 + (long long)blockIdx.x * tile; base < s1v; base += (long long)gridDim.x * tile) {
    float4 g[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long i = base + ((long long)u * blockDim.x + threadIdx.x) * 4;
      if (i < s1v) g[u] = mc_ld_reduce4(mc_grads + i);
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long i = base + ((long long)u * blockDim.x + threadIdx.x) * 4;
      if (i >= s1v) continue;
      float4 p = *reinterpret_cast<const float4*>(params + i);
      float4 mm = *reinterpret_cast<const float4*>(m + i), vv = *reinterpret_cast<const float4*>(v + i);
      adam1(p.x, g[u].x, mm.x, vv.x, c); adam1(p.y, g[u].y, mm.y, vv.y, c);
      adam1(p.z, g[u].z, mm.z, vv.z, c); adam1(p.w, g[u].w, mm.w, vv.w, c);
      *reinterpret_cast<float4*>(m + i) = mm;
      *reinterpret_cast<float4*>(v + i) = vv;
      *reinterpret_cast<float4*>(params + i) = p;
      if (shadow) {
        const __nv_bfloat162 lo = __floats2bfloat162_rn(p.x, p.y), hi = __floats2bfloat162_rn(p.z, p.w);
        uint2 o;
        o.x = *reinterpret_cast<const uint32_t*>(&lo); o.y = *reinterpret_cast<const uint32_t*>(&hi);
        *reinterpret_cast<uint2*>(shadow + i) = o;
      }
    }
  }
  if (blockIdx.x == 0 && threadIdx.x < (int)(s1 - s1v)) {
    const long long k = s1v + threadIdx.x;
    float g = *(volatile float*)(peers.grads[0] + k);
    for (int j = 1; j < world; ++j) g = __fadd_rn(g, *(volatile float*)(peers.grads[j] + k));
    float p = params[k], mm = m[k], vv = v[k];
    adam1(p, g, mm, vv, c);
    m[k] = mm; v[k] = vv; params[k] = p;
    if (shadow) shadow[k] = __float2bfloat16_rn(p);
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) mine->dbg[2] = gtime();
}

// ... and the broadcast of the owned shard by multicast stores (then the params_done flags / barrier)
template <int U>
__global__ void __launch_bounds__(256)
dp_mc_broadcast_kernel(DpPeers peers, float* __restrict__ mc_params, int world, int rank, long long off, long long n,
                       long long chunk, const long long* __restrict__ step_dev, int bucket, int final_barrier) {
  Mailbox* mine = peers.mail[rank];
  const long long epoch = *step_dev + 1;
  const long long s0 = off + (long long)rank * chunk;
  const long long s1 = min(off + n, s0 + chunk);
  const long long s1v = s0 + ((s1 - s0) & ~3ll);
  const long long tile = (long long)blockDim.x * 4 * U;
  const float* params = peers.params[rank];
  for (long long base = s0 + (long long)blockIdx.x * tile; base < s1v; base += (long long)gridDim.x * tile) {
    float4 p[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long i = base + ((long long)u * blockDim.x + threadIdx.x) * 4;
      if (i < s1v) p[u] = *reinterpret_cast<const float4*>(params + i);
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long i = base + ((long long)u * blockDim.x + threadIdx.x) * 4;
      if (i < s1v) mc_st4(mc_params + i, p[u]);
    }
  }
  if (blockIdx.x == 0 && threadIdx.x < (int)(s1 - s1v)) {
    const long long k = s1v + threadIdx.x;
    const float p = params[k];
    for (int j = 0; j < world; ++j) if (j != rank) *(volatile float*)(peers.params[j] + k) = p;
  }
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned int done = atomicAdd(&mine->cta_done[bucket], 1u);
    if (done == gridDim.x - 1) {
      mine->cta_done[bucket] = 0;
      __threadfence_system();
      for (int j = 0; j < world; ++j) peers.mail[j]->params_done[bucket][rank] = epoch;
      if (final_barrier) wait_all(mine->params_done[bucket], world, epoch, &mine->error);
      __threadfence_system();
      mine->dbg[3] = gtime();
    }
  }
}

// all-reduce(MAX) of one double per rank: each rank drops (value, epoch) into every peer's mailbox
__global__ void dp_max_kernel(DpPeers peers, int world, int rank, double* __restrict__ value, const long long* __restrict__ step_dev) {
  if (threadIdx.x != 0) return;
  Mailbox* mine = peers.mail[rank];
  const long long epoch = *step_dev + 1;
  const int par = (int)(epoch & 1);
  const double x = *value;
  for (int j = 0; j < world; ++j) peers.mail[j]->max_value[par][rank] = x;
  __threadfence_system();
  for (int j = 0; j < world; ++j) peers.mail[j]->max_epoch[par][rank] = epoch;
  if (!wait_all(mine->max_epoch[par], world, epoch, &mine->error)) return;
  __threadfence_system();
  double best = mine->max_value[par][0];
  for (int j = 1; j < world; ++j) best = fmax(best, mine->max_value[par][j]);
  *value = best;
}

}  // namespace b200rl

using namespace b200rl;

static int dp_ready(DpState* s);

// Region layout shared by both creation paths: [params | grads | landing | mailbox]
extern "C" int64_t b200rl_dp_region_bytes(int64_t n_params) {
  return (int64_t)(2 * params_bytes(n_params) + landing_bytes(n_params) + ((sizeof(Mailbox) + 255) & ~(size_t)255));
}

static int dp_init_common(DpState* s) {
  for (int k = 0; k < kCeStreams; ++k) B200RL_CUDA_OK(cudaStreamCreateWithFlags(&s->ce_stream[k], cudaStreamNonBlocking));
  for (int c = 0; c < 2; ++c) {
    B200RL_CUDA_OK(cudaEventCreateWithFlags(&s->ce_fork[c], cudaEventDisableTiming));
    for (int k = 0; k < kCeStreams; ++k) B200RL_CUDA_OK(cudaEventCreateWithFlags(&s->ce_join[c][k], cudaEventDisableTiming));
  }
  return B200RL_OK;
}

static void dp_point(DpState* s, int r, void* base) {
  const size_t pb = params_bytes(s->n), lb = landing_bytes(s->n);
  s->peers.params[r] = (float*)base;
  s->peers.grads[r] = (float*)((char*)base + pb);
  s->landing[r] = (float*)((char*)base + 2 * pb);
  s->peers.mail[r] = (Mailbox*)((char*)base + 2 * pb + lb);
}

// The caller owns a symmetric allocation of b200rl_dp_region_bytes(n) on every rank (e.g. torch symmetric memory):
// bases[r] = this process's mapping of rank r's region, multicast_base = the NVSwitch multicast mapping of the same
// regions (NULL: none, b200rl_dp_adam_mc is then refused).  Zeroes this rank's region; the caller synchronises the
// ranks before the first exchange.
extern "C" int b200rl_dp_create_external(b200rl_dp_t* out, const b200rl_dp_cfg* cfg, void* const* bases, void* multicast_base) {
  B200RL_REQUIRE(out && cfg && bases, "null argument");
  B200RL_REQUIRE(cfg->world >= 1 && cfg->world <= kMaxRanks && cfg->rank >= 0 && cfg->rank < cfg->world, "bad world/rank");
  B200RL_REQUIRE(cfg->n_params >= 1, "bad parameter count");
  B200RL_CUDA_OK(cudaSetDevice(cfg->device));
  DpState* s = new DpState();
  memset(s, 0, sizeof(*s));
  s->world = cfg->world; s->rank = cfg->rank; s->device = cfg->device; s->n = cfg->n_params;
  s->external = true;
  s->region = bases[s->rank];
  s->region_bytes = (size_t)b200rl_dp_region_bytes(s->n);
  s->landing_floats = landing_bytes(s->n) / 4;
  for (int r = 0; r < s->world; ++r) {
    B200RL_REQUIRE(bases[r] && ((uintptr_t)bases[r] & 255) == 0, "region of rank %d is missing or misaligned", r);
    dp_point(s, r, bases[r]);
  }
  if (multicast_base) {
    s->mc_params = (float*)multicast_base;
    s->mc_grads = (float*)((char*)multicast_base + params_bytes(s->n));
  }
  B200RL_CUDA_OK(cudaMemset(s->region, 0, s->region_bytes));
  B200RL_CUDA_OK(cudaDeviceSynchronize());
  if (int rc = dp_init_common(s)) return rc;
  *out = (b200rl_dp_t)s;
  return B200RL_OK;
}

extern "C" int b200rl_dp_has_multicast(b200rl_dp_t h) {
  DpState* s = (DpState*)h;
  return s && s->mc_params ? 1 : 0;
}

extern "C" int b200rl_dp_adam_mc(b200rl_dp_t h, int64_t off, int64_t n, float* m, float* v, const int64_t* step_dev, float lr,
                                 double b1, double b2, float eps, int eps_mode, int32_t bucket, int32_t final_barrier,
                                 int32_t max_ctas, void* stream) {
  DpState* s = (DpState*)h;
  B200RL_REQUIRE(s && m && v && step_dev, "null argument");
  B200RL_REQUIRE(s->mc_params, "this exchange was created without a multicast mapping");
  B200RL_REQUIRE(off >= 0 && n >= 1 && off + n <= s->n && off % 4 == 0, "bad bucket range");
  B200RL_REQUIRE(bucket >= 0 && bucket < kBuckets && (eps_mode == 0 || eps_mode == 1), "bad argument");
  B200RL_REQUIRE((((uintptr_t)m | (uintptr_t)v) & 15) == 0, "moment buffers must be 16-byte aligned");
  if (int rc = dp_ready(s)) return rc;
  long long chunk = (n + s->world - 1) / s->world;
  chunk = (chunk + 3) & ~3ll;
  const long long vec = (chunk + 3) / 4;
  constexpr int U = 4;
  int blocks = (int)std::max<long long>(1, std::min<long long>((vec + 256 * U - 1) / (256 * U), 2ll * kNumSMs));
  if (max_ctas > 0) blocks = std::min(blocks, (int)max_ctas);
  dp_mc_adam_kernel<U><<<blocks, 256, 0, as_stream(stream)>>>(s->peers, s->mc_grads, s->mc_params, s->world, s->rank, off, n,
                                                             chunk, m, v, (const long long*)step_dev, lr, b1, b2, eps,
                                                             eps_mode, bucket, final_barrier);
  B200RL_LAUNCH_OK();
  return B200RL_OK;
}

extern "C" int b200rl_dp_reduce_adam_mc(b200rl_dp_t h, int64_t off, int64_t n, float* m, float* v, const int64_t* step_dev,
                                        float lr, double b1, double b2, float eps, int eps_mode, int32_t bucket,
                                        void* shadow_bf16, int32_t max_ctas, void* stream) {
  DpState* s = (DpState*)h;
  B200RL_REQUIRE(s && m && v && step_dev, "null argument");
  B200RL_REQUIRE(s->mc_params, "this exchange was created without a multicast mapping");
  B200RL_REQUIRE(off >= 0 && n >= 1 && off + n <= s->n && off % 4 == 0, "bad bucket range");
  B200RL_REQUIRE(bucket >= 0 && bucket < kBuckets && (eps_mode == 0 || eps_mode == 1), "bad argument");
  B200RL_REQUIRE((((uintptr_t)m | (uintptr_t)v) & 15) == 0 && ((uintptr_t)shadow_bf16 & 7) == 0, "misaligned buffers");
  if (int rc = dp_ready(s)) return rc;
  long long chunk = (n + s->world - 1) / s->world;
  chunk = (chunk + 3) & ~3ll;
  const long long vec = (chunk + 3) / 4;
  constexpr int U = 4;
  int blocks = (int)std::max<long long>(1, std::min<long long>((vec + 256 * U - 1) / (256 * U), 2ll * kNumSMs));
  if (max_ctas > 0) blocks = std::min(blocks, (int)max_ctas);
  dp_mc_reduce_adam_kernel<U><<<blocks, 256, 0, as_stream(stream)>>>(s->peers, s->mc_grads, s->world, s->rank, off, n, chunk, m, v,
                                                                    (__nv_bfloat16*)shadow_bf16, (const long long*)step_dev,
                                                                    lr, b1, b2, eps, eps_mode, bucket);
  B200RL_LAUNCH_OK();
  return B200RL_OK;
}

extern "C" int b200rl_dp_broadcast_mc(b200rl_dp_t h, int64_t off, int64_t n, const int64_t* step_dev, int32_t bucket,
                                      int32_t final_barrier, int32_t max_ctas, void* stream) {
  DpState* s = (DpState*)h;
  B200RL_REQUIRE(s && step_dev, "null argument");
  B200RL_REQUIRE(s->mc_params, "this exchange was created without a multicast mapping");
  B200RL_REQUIRE(off >= 0 && n >= 1 && off + n <= s->n && off % 4 == 0, "bad bucket range");
  B200RL_REQUIRE(bucket >= 0 && bucket < kBuckets, "bad argument");
  if (int rc = dp_ready(s)) return rc;
  long long chunk = (n + s->world - 1) / s->world;
  chunk = (chunk + 3) & ~3ll;
  const long long vec = (chunk + 3) / 4;
  constexpr int U = 4;
  int blocks = (int)std::max<long long>(1, std::min<long long>((vec + 256 * U - 1) / (256 * U), 2ll * kNumSMs));
  if (max_ctas > 0) blocks = std::min(blocks, (int)max_ctas);
  dp_mc_broadcast_kernel<U><<<blocks, 256, 0, as_stream(stream)>>>(s->peers, s->mc_params, s->world, s->rank, off, n, chunk,
                                                                  (const long long*)step_dev, bucket, final_barrier);
  B200RL_LAUNCH_OK();
  return B200RL_OK;
}

extern "C" int b200rl_dp_create(b200rl_dp_t* out, const b200rl_dp_cfg* cfg) {
  B200RL_REQUIRE(out && cfg, "null argument");
  B200RL_REQUIRE(cfg->world >= 1 && cfg->world <= kMaxRanks && cfg->rank >= 0 && cfg->rank < cfg->world, "bad world/rank");
  B200RL_REQUIRE(cfg->n_params >= 1, "bad parameter count");
  B200RL_CUDA_OK(cudaSetDevice(cfg->device));
  DpState* s = new DpState();
  memset(s, 0, sizeof(*s));
  s->world = cfg->world; s->rank = cfg->rank; s->device = cfg->device; s->n = cfg->n_params;
  const size_t pb = params_bytes(s->n);
  const size_t lb = landing_bytes(s->n);
  s->region_bytes = 2 * pb + lb + ((sizeof(Mailbox) + 255) & ~(size_t)255);
  cudaError_t e = cudaMalloc(&s->region, s->region_bytes);
  if (e != cudaSuccess) { delete s; set_error("cudaMalloc(%zu): %s", s->region_bytes, cudaGetErrorString(e)); return B200RL_ECUDA; }
  B200RL_CUDA_OK(cudaMemset(s->region, 0, s->region_bytes));
  s->landing_floats = lb / 4;
  if (int rc = dp_init_common(s)) return rc;
  B200RL_CUDA_OK(cudaDeviceSynchronize());
  s->peers.params[s->rank] = (float*)s->region;
  s->peers.grads[s->rank] = (float*)((char*)s->region + pb);
  s->landing[s->rank] = (float*)((char*)s->region + 2 * pb);
  s->peers.mail[s->rank] = (Mailbox*)((char*)s->region + 2 * pb + lb);
  *out = (b200rl_dp_t)s;
  return B200RL_OK;
}

extern "C" int b200rl_dp_destroy(b200rl_dp_t h) {
  DpState* s = (DpState*)h;
  if (!s) return B200RL_OK;
  cudaDeviceSynchronize();
  for (int j = 0; j < s->world; ++j)
    if (s->opened[j]) cudaIpcCloseMemHandle(s->opened[j]);
  if (s->region && !s->external) cudaFree(s->region);
  for (int k = 0; k < kCeStreams; ++k) if (s->ce_stream[k]) cudaStreamDestroy(s->ce_stream[k]);
  for (int c = 0; c < 2; ++c) {
    if (s->ce_fork[c]) cudaEventDestroy(s->ce_fork[c]);
    for (int k = 0; k < kCeStreams; ++k) if (s->ce_join[c][k]) cudaEventDestroy(s->ce_join[c][k]);
  }
  delete s;
  return B200RL_OK;
}

extern "C" int b200rl_dp_buffers(b200rl_dp_t h, float** params, float** grads) {
  DpState* s = (DpState*)h;
  B200RL_REQUIRE(s && params && grads, "null argument");
  *params = s->peers.params[s->rank];
  *grads = s->peers.grads[s->rank];
  return B200RL_OK;
}

extern "C" int b200rl_dp_export(b200rl_dp_t h, void* handle64) {
  DpState* s = (DpState*)h;
  B200RL_REQUIRE(s && handle64, "null argument");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "handle size");
  cudaIpcMemHandle_t hd;
  B200RL_CUDA_OK(cudaIpcGetMemHandle(&hd, s->region));
  memcpy(handle64, &hd, 64);
  return B200RL_OK;
}

extern "C" int b200rl_dp_import(b200rl_dp_t h, int32_t peer_rank, const void* handle64) {
  DpState* s = (DpState*)h;
  B200RL_REQUIRE(s && handle64, "null argument");
  B200RL_REQUIRE(peer_rank >= 0 && peer_rank < s->world && peer_rank != s->rank, "bad peer rank");
  B200RL_REQUIRE(!s->opened[peer_rank], "peer already imported");
  cudaIpcMemHandle_t hd;
  memcpy(&hd, handle64, 64);
  void* base = nullptr;
  B200RL_CUDA_OK(cudaIpcOpenMemHandle(&base, hd, cudaIpcMemLazyEnablePeerAccess));
  const size_t pb = params_bytes(s->n);
  s->opened[peer_rank] = base;
  s->peers.params[peer_rank] = (float*)base;
  s->peers.grads[peer_rank] = (float*)((char*)base + pb);
  s->landing[peer_rank] = (float*)((char*)base + 2 * pb);
  s->peers.mail[peer_rank] = (Mailbox*)((char*)base + 2 * pb + landing_bytes(s->n));
  return B200RL_OK;
}

static int dp_ready(DpState* s) {
  for (int j = 0; j < s->world; ++j)
    if (!s->peers.mail[j]) { set_error("data-parallel region of rank %d has not been imported", j); return B200RL_EINVAL; }
  return B200RL_OK;
}

extern "C" int b200rl_dp_max_f64(b200rl_dp_t h, double* value_dev, const int64_t* step_dev, void* stream) {
  DpState* s = (DpState*)h;
  B200RL_REQUIRE(s && value_dev && step_dev, "null argument");
  if (int rc = dp_ready(s)) return rc;
  dp_max_kernel<<<1, 32, 0, as_stream(stream)>>>(s->peers, s->world, s->rank, value_dev, (const long long*)step_dev);
  B200RL_LAUNCH_OK();
  return B200RL_OK;
}

extern "C" int b200rl_dp_adam(b200rl_dp_t h, int64_t off, int64_t n, float* m, float* v, const int64_t* step_dev, float lr,
                              double b1, double b2, float eps, int eps_mode, int32_t bucket, int32_t final_barrier,
                              void* stream) {
  DpState* s = (DpState*)h;
  B200RL_REQUIRE(s && m && v && step_dev, "null argument");
  B200RL_REQUIRE(off >= 0 && n >= 1 && off + n <= s->n && off % 4 == 0, "bad bucket range");
  B200RL_REQUIRE(bucket >= 0 && bucket < kBuckets && (eps_mode == 0 || eps_mode == 1), "bad argument");
  B200RL_REQUIRE((((uintptr_t)m | (uintptr_t)v) & 15) == 0, "moment buffers must be 16-byte aligned");
  if (int rc = dp_ready(s)) return rc;
  long long chunk = (n + s->world - 1) / s->world;
  chunk = (chunk + 3) & ~3ll;
  const long long vec = (chunk + 3) / 4;
  cudaStream_t st = as_stream(stream);
  // NVLink saturates with one CTA per SM; more only helps the local (HBM) part of the update, which dominates at R = 2
  static const int per_sm_env = getenv("B200RL_DP_CTAS_PER_SM") ? atoi(getenv("B200RL_DP_CTAS_PER_SM")) : 0;
  const int per_sm = per_sm_env > 0 ? per_sm_env : (s->world <= 2 ? 3 : 1);
#define DP_LAUNCH(W_, U_)                                                                                              \
  {                                                                                                                    \
    const int blocks = (int)std::max<long long>(1, std::min<long long>((vec + 256 * U_ - 1) / (256 * U_), per_sm * kNumSMs)); \
    dp_adam_kernel<W_, U_><<<blocks, 256, 0, st>>>(s->peers, s->world, s->rank, off, n, chunk, m, v,                   \
                                                  (const long long*)step_dev, lr, b1, b2, eps, eps_mode, bucket,        \
                                                  final_barrier);                                                      \
  }
  if (s->world <= 2) DP_LAUNCH(2, 2) else if (s->world <= 4) DP_LAUNCH(4, 2) else DP_LAUNCH(8, 2)
#undef DP_LAUNCH
  B200RL_LAUNCH_OK();
  return B200RL_OK;
}

// ---- copy-engine exchange (see dp_flag_kernel)
extern "C" int b200rl_debug_stamp(unsigned long long* buf, int slot, void* stream);
static unsigned long long* g_ce_marks = nullptr;   // tools/step_phases.py: 6 global-timer stamps of the last exchange
extern "C" int b200rl_debug_dp_ce_marks(unsigned long long* buf8) { g_ce_marks = buf8; return 0; }
#define CE_MARK(k) do { if (g_ce_marks) b200rl_debug_stamp(g_ce_marks, (k), (void*)st); } while (0)
static void dp_shard(const DpState* s, int64_t off, int64_t n, int r, long long* chunk_out, long long* s0, long long* len) {
  long long chunk = (n + s->world - 1) / s->world;
  chunk = (chunk + 3) & ~3ll;
  *chunk_out = chunk;
  *s0 = off + (long long)r * chunk;
  *len = std::max<long long>(0, std::min<long long>(off + n, *s0 + chunk) - *s0);
}

// One phase of DMA copies: (dst, src, floats) per peer, each cut into pieces and spread round-robin over kCeStreams
// internal streams that fork from / join back into `st` by events (capturable: they become parallel memcpy nodes).
struct CeCopy { float* dst; const float* src; long long n; };
static int ce_copies(DpState* s, cudaStream_t st, int phase, const CeCopy* list, int count) {
  if (count == 0) return B200RL_OK;
  static const int pieces_env = getenv("B200RL_DP_CE_PIECES") ? atoi(getenv("B200RL_DP_CE_PIECES")) : 0;
  static const int lanes_env = getenv("B200RL_DP_CE_LANES") ? atoi(getenv("B200RL_DP_CE_LANES")) : 0;
  // Measured on 2 GPUs: one 16 MB copy per direction runs at ~310 GB/s whether it is issued whole or in 2-8 pieces on
  // parallel streams (pieces only add overhead: 53 / 54 / 67 / 78 us) -- copies between one pair of GPUs share one
  // engine.  Copies to DIFFERENT peers (R > 2) can overlap, so each peer's copy stays whole and the peers are spread
  // over the lanes.
  const int pieces = pieces_env > 0 ? pieces_env : 1;
  B200RL_CUDA_OK(cudaEventRecord(s->ce_fork[phase], st));
  const int lanes = std::max(1, std::min(std::min(kCeStreams, lanes_env > 0 ? lanes_env : kCeStreams), count * pieces));
  for (int k = 0; k < lanes; ++k) B200RL_CUDA_OK(cudaStreamWaitEvent(s->ce_stream[k], s->ce_fork[phase], 0));
  int q = 0;
  for (int i = 0; i < count; ++i) {
    const long long per = ((list[i].n + pieces - 1) / pieces + 3) & ~3ll;
    for (long long o = 0; o < list[i].n; o += per, ++q)
      B200RL_CUDA_OK(cudaMemcpyAsync(list[i].dst + o, list[i].src + o, (size_t)std::min(per, list[i].n - o) * 4,
                                     cudaMemcpyDeviceToDevice, s->ce_stream[q % lanes]));
  }
  for (int k = 0; k < lanes; ++k) {
    B200RL_CUDA_OK(cudaEventRecord(s->ce_join[phase][k], s->ce_stream[k]));
    B200RL_CUDA_OK(cudaStreamWaitEvent(st, s->ce_join[phase][k], 0));
  }
  return B200RL_OK;
}

extern "C" int b200rl_dp_reduce_adam_ce(b200rl_dp_t h, int64_t off, int64_t n, float* m, float* v, const int64_t* step_dev,
                                        float lr, double b1, double b2, float eps, int eps_mode, int32_t bucket,
                                        void* shadow_bf16, int32_t max_ctas, void* stream) {
  DpState* s = (DpState*)h;
  B200RL_REQUIRE(s && m && v && step_dev, "null argument");
  B200RL_REQUIRE(off >= 0 && n >= 1 && off + n <= s->n && off % 4 == 0, "bad bucket range");
  B200RL_REQUIRE(bucket >= 0 && bucket < kBuckets && (eps_mode == 0 || eps_mode == 1), "bad argument");
  B200RL_REQUIRE((((uintptr_t)m | (uintptr_t)v) & 15) == 0 && ((uintptr_t)shadow_bf16 & 7) == 0, "misaligned buffers");
  if (int rc = dp_ready(s)) return rc;
  long long chunk, s0, len;
  dp_shard(s, off, n, s->rank, &chunk, &s0, &len);
  const size_t lbase = landing_base(off, bucket);
  B200RL_REQUIRE(s->landing_floats >= lbase + (size_t)chunk * s->world, "landing buffer too small");   // sized in dp_create
  cudaStream_t st = as_stream(stream);
  CE_MARK(0);
  // push the shard every peer owns of MY gradients into that peer's landing buffer (slot = my rank)
  CeCopy list[kMaxRanks];
  int count = 0;
  for (int d = 1; d < s->world; ++d) {   // start with the next rank: the links are hit evenly
    const int j = (s->rank + d) % s->world;
    long long cj, sj, lj;
    dp_shard(s, off, n, j, &cj, &sj, &lj);
    if (lj > 0) list[count++] = CeCopy{s->landing[j] + lbase + (size_t)s->rank * chunk, s->peers.grads[s->rank] + sj, lj};
  }
  if (int rc = ce_copies(s, st, 0, list, count)) return rc;
  CE_MARK(1);
  // "my pushes have landed" to every rank; wait until every rank's have landed here
  dp_flag_kernel<<<1, 32, 0, st>>>(s->peers, s->world, s->rank, (const long long*)step_dev, bucket, 0, 1);
  B200RL_LAUNCH_OK();
  CE_MARK(2);
  if (len > 0) {
    const long long vec = (len + 3) / 4;
    int blocks = (int)std::max<long long>(1, std::min<long long>((vec + 255) / 256, 4ll * kNumSMs));
    if (max_ctas > 0) blocks = std::min(blocks, (int)max_ctas);
#define DP_LOCAL(W_)                                                                                                    \
    dp_adam_local_kernel<W_><<<blocks, 256, 0, st>>>(s->peers.params[s->rank], s->peers.grads[s->rank], s->landing[s->rank] + lbase, \
                                                    s->world, s->rank, s0, len, chunk, m, v, (__nv_bfloat16*)shadow_bf16,  \
                                                    (const long long*)step_dev, lr, b1, b2, eps, eps_mode)
    if (s->world <= 2) DP_LOCAL(2); else if (s->world <= 4) DP_LOCAL(4); else DP_LOCAL(8);
#undef DP_LOCAL
    B200RL_LAUNCH_OK();
  }
  CE_MARK(3);
  return B200RL_OK;
}

extern "C" int b200rl_dp_broadcast_ce(b200rl_dp_t h, int64_t off, int64_t n, const int64_t* step_dev, int32_t bucket,
                                      int32_t final_barrier, void* stream) {
  DpState* s = (DpState*)h;
  B200RL_REQUIRE(s && step_dev, "null argument");
  B200RL_REQUIRE(off >= 0 && n >= 1 && off + n <= s->n && off % 4 == 0, "bad bucket range");
  B200RL_REQUIRE(bucket >= 0 && bucket < kBuckets, "bad argument");
  if (int rc = dp_ready(s)) return rc;
  long long chunk, s0, len;
  dp_shard(s, off, n, s->rank, &chunk, &s0, &len);
  cudaStream_t st = as_stream(stream);
  CE_MARK(4);
  CeCopy list[kMaxRanks];
  int count = 0;
  if (len > 0)
    for (int d = 1; d < s->world; ++d) {
      const int j = (s->rank + d) % s->world;
      list[count++] = CeCopy{s->peers.params[j] + s0, s->peers.params[s->rank] + s0, len};
    }
  if (int rc = ce_copies(s, st, 1, list, count)) return rc;
  CE_MARK(5);
  dp_flag_kernel<<<1, 32, 0, st>>>(s->peers, s->world, s->rank, (const long long*)step_dev, bucket, 1, final_barrier ? 1 : 0);
  B200RL_LAUNCH_OK();
  CE_MARK(6);
  return B200RL_OK;
}

// tools/dp_bench.py: raw peer-memory bandwidth as seen by SM loads / stores (mode 0 = read the peer's gradient
// region, 1 = write the peer's gradient region, 2 = read the local one)
__global__ void __launch_bounds__(256) dp_probe_kernel(b200rl::DpPeers peers, int peer, int rank, long long n4, int mode, float* sink) {
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  const long long stride = (long long)gridDim.x * blockDim.x;
  float* base = mode == 2 ? peers.grads[rank] : peers.grads[peer];
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += 4 * stride) {
    if (mode == 1) {
#pragma unroll
      for (int u = 0; u < 4; ++u)
        if (i + u * stride < n4) b200rl::st_sys4(base + 4 * (i + u * stride), make_float4(1.f, 2.f, 3.f, 4.f));
    } else {
      float4 t[4];
#pragma unroll
      for (int u = 0; u < 4; ++u)
        if (i + u * stride < n4) t[u] = b200rl::ld_sys4(base + 4 * (i + u * stride));
#pragma unroll
      for (int u = 0; u < 4; ++u)
        if (i + u * stride < n4) { acc.x += t[u].x; acc.y += t[u].y; acc.z += t[u].z; acc.w += t[u].w; }
    }
  }
  if (mode != 1 && acc.x + acc.y + acc.z + acc.w == 123.456f) sink[0] = acc.x;
}
extern "C" int b200rl_debug_dp_probe(b200rl_dp_t h, int peer, long long n_floats, int mode, int blocks, float* sink, void* stream) {
  DpState* s = (DpState*)h;
  dp_probe_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(s->peers, peer, s->rank, n_floats / 4, mode, sink);
  return cudaGetLastError() == cudaSuccess ? 0 : -2;
}

extern "C" int b200rl_debug_dp_stamps(b200rl_dp_t h, long long* out8) {
  DpState* s = (DpState*)h;
  const Mailbox* mb = s->peers.mail[s->rank];
  return cudaMemcpy(out8, (const void*)mb->dbg, 64, cudaMemcpyDeviceToHost) == cudaSuccess ? 0 : -2;
}

extern "C" int b200rl_dp_status(b200rl_dp_t h) {
  DpState* s = (DpState*)h;
  B200RL_REQUIRE(s, "null argument");
  int err = 0;
  const Mailbox* mb = s->peers.mail[s->rank];
  B200RL_CUDA_OK(cudaMemcpy(&err, (const void*)&mb->error, sizeof(int), cudaMemcpyDeviceToHost));
  if (err) { set_error("data-parallel exchange timed out waiting for a peer"); return B200RL_ECUDA; }
  return B200RL_OK;
}
