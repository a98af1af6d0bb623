// BatchNormalization (batch-statistics mode) fused with its neighbouring activation / MaxPool3D(2) /
// UpSampling3D(2), forward and backward.  HBM-bound passes: 16-byte vector loads/stores, fp32 math,
// fp64 statistics, deterministic two-stage reductions (partials per block, fixed-order final sum).
//
// Reference layers: keras BatchNormalization() after every Conv3D (lattice_vae.py:174,214,225;
// unet.py:278-338), LeakyReLU/ReLU (lattice_vae.py:175,215,226), MaxPool3D (lattice_vae.py:176;
// unet.py:282,291,300), UpSampling3D (lattice_vae.py:217; unet.py:309,319,329).
// Semantics register: SURVEY.md §8c R3 (eps 1e-3, momentum 0.99, biased variance, moving-variance
// factor n/(n-(1+eps))), R4 (LeakyReLU alpha, gradient at 0), R6 (max-pool gradient -> first maximum).
//
// Layouts: x, dy are [B,D,H,W,ld] channels-last with C used channels; T = bf16 (C % 8 == 0) or
// fp32 (C % 4 == 0).  "post" describes what follows BN+activation in the forward graph:
//   NONE : y has the shape of x;  POOL2 : y is [B,D/2,H/2,W/2,C] (+ uint8 argmax);  UP2 : y is [B,2D,2H,2W,C].
#include <cooperative_groups.h>
#include <cuda_fp16.h>

#include <type_traits>

#include "common.cuh"

namespace icsg3d {

template <typename T>
struct VecIO;
template <>
struct VecIO<__nv_bfloat16> {
  static constexpr int N = 8;
  typedef uint4 raw_t;
  __device__ static raw_t load_raw(const __nv_bfloat16* p) { return *reinterpret_cast<const uint4*>(p); }
  __device__ static void cvt(const raw_t& q, float (&v)[8]) {
    float2 a = unpack_bf16x2(q.x), b = unpack_bf16x2(q.y), c = unpack_bf16x2(q.z), d = unpack_bf16x2(q.w);
    v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y; v[4] = c.x; v[5] = c.y; v[6] = d.x; v[7] = d.y;
  }
  __device__ static void load(const __nv_bfloat16* p, float (&v)[8]) {
    const uint4 q = *reinterpret_cast<const uint4*>(p);
    float2 a = unpack_bf16x2(q.x), b = unpack_bf16x2(q.y), c = unpack_bf16x2(q.z), d = unpack_bf16x2(q.w);
    v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y; v[4] = c.x; v[5] = c.y; v[6] = d.x; v[7] = d.y;
  }
  __device__ static void store(__nv_bfloat16* p, const float (&v)[8]) {
    uint4 q;
    q.x = pack_bf16x2(v[0], v[1]); q.y = pack_bf16x2(v[2], v[3]);
    q.z = pack_bf16x2(v[4], v[5]); q.w = pack_bf16x2(v[6], v[7]);
    *reinterpret_cast<uint4*>(p) = q;
  }
};
template <>
struct VecIO<float> {
  static constexpr int N = 4;
  typedef float4 raw_t;
  __device__ static raw_t load_raw(const float* p) { return *reinterpret_cast<const float4*>(p); }
  __device__ static void cvt(const raw_t& q, float (&v)[4]) { v[0] = q.x; v[1] = q.y; v[2] = q.z; v[3] = q.w; }
  __device__ static void load(const float* p, float (&v)[4]) {
    const float4 q = *reinterpret_cast<const float4*>(p);
    v[0] = q.x; v[1] = q.y; v[2] = q.z; v[3] = q.w;
  }
  __device__ static void store(float* p, const float (&v)[4]) {
    *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
  }
};

template <int N>
__device__ __forceinline__ void store_bf16_vec(__nv_bfloat16* p, const float (&v)[N]) {
  if constexpr (N == 8) {
    VecIO<__nv_bfloat16>::store(p, v);
  } else {
    uint2 q;
    q.x = pack_bf16x2(v[0], v[1]);
    q.y = pack_bf16x2(v[2], v[3]);
    *reinterpret_cast<uint2*>(p) = q;
  }
}

__device__ __forceinline__ float act_fwd(float z, int act, float alpha) {
  if (act == ICSG3D_ACT_RELU) return z > 0.f ? z : 0.f;
  if (act == ICSG3D_ACT_LEAKY) return z > 0.f ? z : alpha * z;
  return z;
}
__device__ __forceinline__ float act_grad(float z, int act, float alpha) {
  if (act == ICSG3D_ACT_RELU) return z > 0.f ? 1.f : 0.f;
  if (act == ICSG3D_ACT_LEAKY) return z > 0.f ? 1.f : alpha;
  return 1.f;
}

static constexpr int kBnThreads = 256;

// ------------------------------------------------------------------------------------------------
// statistics: partials[blk][0][c] = sum x, partials[blk][1][c] = sum x^2 (fp64)
// ------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(kBnThreads, 4) bn_stats_kernel(const T* __restrict__ x, int ldx, long long M, int C,
                                                             double* __restrict__ partials) {
  pdl_prologue();
  constexpr int V = VecIO<T>::N;
  extern __shared__ double sred[];  // [2][rows_per_iter][C] would be too big: reduce per channel group instead
  const int cg = C / V;
  const int rpi = kBnThreads / cg;  // rows per iteration handled by this block
  const int g = threadIdx.x % cg;
  const int rl = threadIdx.x / cg;
  // per-thread partials in fp32: the grid is sized so that a thread sees at most a few dozen rows; everything
  // downstream (block reduction, partials, final sums) is fp64
  float s[V], q[V];
#pragma unroll
  for (int i = 0; i < V; ++i) s[i] = q[i] = 0.f;
  if (rl < rpi) {
    constexpr int U = 4;  // independent 16-byte loads in flight per thread
    const long long stride = static_cast<long long>(gridDim.x) * rpi;
    for (long long r0 = static_cast<long long>(blockIdx.x) * rpi + rl; r0 < M; r0 += U * stride) {
      typename VecIO<T>::raw_t raw[U];
#pragma unroll
      for (int j = 0; j < U; ++j)
        if (r0 + j * stride < M) raw[j] = VecIO<T>::load_raw(x + (r0 + j * stride) * ldx + g * V);
#pragma unroll
      for (int j = 0; j < U; ++j) {
        if (r0 + j * stride < M) {
          float v[V];
          VecIO<T>::cvt(raw[j], v);
#pragma unroll
          for (int i = 0; i < V; ++i) {
            s[i] += v[i];
            q[i] = fmaf(v[i], v[i], q[i]);
          }
        }
      }
    }
  }
  // block reduction over the row lanes: smem [rpi][cg*V] doubles, two passes (sum, sumsq)
  double* out = partials + static_cast<size_t>(blockIdx.x) * 2 * C;
  for (int pass = 0; pass < 2; ++pass) {
    __syncthreads();
    if (rl < rpi) {
#pragma unroll
      for (int i = 0; i < V; ++i) sred[rl * C + g * V + i] = static_cast<double>(pass == 0 ? s[i] : q[i]);
    }
    __syncthreads();
    for (int c = threadIdx.x; c < C; c += kBnThreads) {
      double a = 0.0;
      for (int r = 0; r < rpi; ++r) a += sred[r * C + c];
      out[pass * C + c] = a;
    }
  }
}

// sums[c] = sum_p partials[p][c] in a FIXED order (deterministic): 8 warps stride over the partials with
// coalesced 32-column reads, then the 8 warp partials are added in warp order.  Block = 32 columns.
__global__ void __launch_bounds__(256) bn_reduce_partials_kernel(const double* __restrict__ partials, int nparts, int C2,
                                                                 double* __restrict__ sums) {
  pdl_prologue();
  __shared__ double red[8][33];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + lane;
  double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
  if (c < C2) {
    int p = w;
    for (; p + 24 < nparts; p += 32) {
      a0 += partials[static_cast<size_t>(p) * C2 + c];
      a1 += partials[static_cast<size_t>(p + 8) * C2 + c];
      a2 += partials[static_cast<size_t>(p + 16) * C2 + c];
      a3 += partials[static_cast<size_t>(p + 24) * C2 + c];
    }
    for (; p < nparts; p += 8) a0 += partials[static_cast<size_t>(p) * C2 + c];
  }
  red[w][lane] = (a0 + a1) + (a2 + a3);
  __syncthreads();
  if (w == 0 && c < C2) {
    double t = 0.0;
#pragma unroll
    for (int i = 0; i < 8; ++i) t += red[i][lane];
    sums[c] = t;
  }
}

__global__ void bn_finalize_kernel(const double* __restrict__ sums, double count, const float* __restrict__ gamma,
                                   const float* __restrict__ beta, float eps, float* __restrict__ mean_out,
                                   float* __restrict__ rstd_out, float* __restrict__ scale, float* __restrict__ shift,
                                   float* __restrict__ moving_mean, float* __restrict__ moving_var, float momentum,
                                   int C) {
  pdl_prologue();
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const double mean = sums[c] / count;
  double var = sums[C + c] / count - mean * mean;
  if (var < 0.0) var = 0.0;
  const double rstd = 1.0 / sqrt(var + static_cast<double>(eps));
  const double g = gamma ? static_cast<double>(gamma[c]) : 1.0;
  const double b = beta ? static_cast<double>(beta[c]) : 0.0;
  mean_out[c] = static_cast<float>(mean);
  rstd_out[c] = static_cast<float>(rstd);
  scale[c] = static_cast<float>(g * rstd);
  shift[c] = static_cast<float>(b - mean * g * rstd);
  if (moving_mean) {
    // keras 2.3.1 / TF backend: mov -= (mov - value) * (1 - momentum); variance * n / (n - (1 + eps))
    const double var_unbiased = var * (count / (count - (1.0 + static_cast<double>(eps))));
    moving_mean[c] = static_cast<float>(moving_mean[c] - (moving_mean[c] - mean) * (1.0 - momentum));
    moving_var[c] = static_cast<float>(moving_var[c] - (moving_var[c] - var_unbiased) * (1.0 - momentum));
  }
}

// Fused single-GPU path: fixed-order reduction of the partials (as bn_reduce_partials_kernel) followed by the
// finalisation of the same 32 channels.  mode 0: forward statistics -> mean/rstd/scale/shift (+moving averages);
// mode 1: backward sums -> sums[2][C] (+ dgamma = sum g*xhat, dbeta = sum g).
static constexpr int kRedCh = 8;        // channels per block of bn_reduce_fused_kernel
static constexpr int kRedSlots = 64;    // partial-row slots per block (1024 threads = 64 slots x 16 columns)
__global__ void __launch_bounds__(1024) bn_reduce_fused_kernel(const double* __restrict__ partials, int nparts, int C,
                                                               int mode, double count, const float* __restrict__ gamma,
                                                               const float* __restrict__ beta, float eps,
                                                               double* __restrict__ sums, float* __restrict__ mean_out,
                                                               float* __restrict__ rstd_out, float* __restrict__ scale,
                                                               float* __restrict__ shift, float* __restrict__ moving_mean,
                                                               float* __restrict__ moving_var, float momentum,
                                                               float* __restrict__ dgamma, float* __restrict__ dbeta) {
  pdl_prologue();
  // The reduction is latency bound (a few hundred KB of fp64 partials): 1024 threads keep every load of a column
  // independent and in flight at once; the order of the additions is a function of (nparts) only => deterministic.
  __shared__ double red[kRedSlots][2 * kRedCh + 1];
  __shared__ double tot[2 * kRedCh];
  const int col = threadIdx.x & (2 * kRedCh - 1);   // 0..7 = sum x / sum g, 8..15 = sum x^2 / sum g*xhat
  const int slot = threadIdx.x >> 4;
  const int c = blockIdx.x * kRedCh + (col & (kRedCh - 1));
  const int gcol = (col >> 3) * C + c;
  const size_t C2 = 2 * static_cast<size_t>(C);
  double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
  if (c < C) {
    int p = slot;
    for (; p + 3 * kRedSlots < nparts; p += 4 * kRedSlots) {
      a0 += partials[static_cast<size_t>(p) * C2 + gcol];
      a1 += partials[static_cast<size_t>(p + kRedSlots) * C2 + gcol];
      a2 += partials[static_cast<size_t>(p + 2 * kRedSlots) * C2 + gcol];
      a3 += partials[static_cast<size_t>(p + 3 * kRedSlots) * C2 + gcol];
    }
    for (; p < nparts; p += kRedSlots) a0 += partials[static_cast<size_t>(p) * C2 + gcol];
  }
  red[slot][col] = (a0 + a1) + (a2 + a3);
  __syncthreads();
  if (threadIdx.x < 2 * kRedCh) {
    double t0 = 0.0, t1 = 0.0, t2 = 0.0, t3 = 0.0;
#pragma unroll
    for (int i = 0; i < kRedSlots; i += 4) {
      t0 += red[i][threadIdx.x];
      t1 += red[i + 1][threadIdx.x];
      t2 += red[i + 2][threadIdx.x];
      t3 += red[i + 3][threadIdx.x];
    }
    tot[threadIdx.x] = (t0 + t1) + (t2 + t3);
  }
  __syncthreads();
  if (threadIdx.x >= kRedCh || c >= C) return;
  const double s0 = tot[threadIdx.x], s1 = tot[kRedCh + threadIdx.x];
  if (sums) {
    sums[c] = s0;
    sums[C + c] = s1;
  }
  if (mode == 0) {
    const double mean = s0 / count;
    double var = s1 / count - mean * mean;
    if (var < 0.0) var = 0.0;
    const double rstd = 1.0 / sqrt(var + static_cast<double>(eps));
    const double g = gamma ? static_cast<double>(gamma[c]) : 1.0;
    const double b = beta ? static_cast<double>(beta[c]) : 0.0;
    mean_out[c] = static_cast<float>(mean);
    rstd_out[c] = static_cast<float>(rstd);
    scale[c] = static_cast<float>(g * rstd);
    shift[c] = static_cast<float>(b - mean * g * rstd);
    if (moving_mean) {
      const double var_unbiased = var * (count / (count - (1.0 + static_cast<double>(eps))));
      moving_mean[c] = static_cast<float>(moving_mean[c] - (moving_mean[c] - mean) * (1.0 - momentum));
      moving_var[c] = static_cast<float>(moving_var[c] - (moving_var[c] - var_unbiased) * (1.0 - momentum));
    }
  } else {
    if (dbeta) dbeta[c] = static_cast<float>(s0);
    if (dgamma) dgamma[c] = static_cast<float>(s1);
  }
}

// ------------------------------------------------------------------------------------------------
// Data-parallel fusion: partial reduction + ALL-REDUCE over NVLink peer memory + finalisation in ONE kernel.
// Replaces bn_reduce_partials -> NCCL all-reduce (2C doubles, pure latency) -> bn_finalize / the backward-sum
// exchange.  Every rank owns one symmetric buffer (torch symmetric memory; all peers' buffers are mapped into this
// process):   flags  uint64 [nslots][2][world][64]          (byte offset 0)
//             data   double [nslots][2][world][2*cmax]      (byte offset nslots*2*world*64*8)
// Protocol per call site (slot) and step (epoch, read from device memory so that a CUDA graph can be replayed):
//   1. block b reduces the local partials of its 8 channels (fixed order);
//   2. it PUSHES the 16 sums into data[slot][epoch&1][my_rank] of EVERY rank (remote stores), fences system-wide,
//      then writes flags[slot][epoch&1][my_rank][b] = epoch on every rank;
//   3. it polls its LOCAL flags of all source ranks, then adds the LOCAL copies in rank order: every rank computes
//      bit-identical global sums (deterministic, no atomics); parity double-buffering + the step's other exchanges keep
//      a fast rank from overwriting a slot a slow rank still reads.
// Spins are bounded (trap after ~4 s) so a mismatched launch sequence surfaces as an error, never as a hung GPU.
// ------------------------------------------------------------------------------------------------
struct BnPeerParams {
  const unsigned long long* peers;  // device array [world]: base address of every rank's symmetric buffer
  int world, rank, slot, nslots, cmax;
  const long long* epoch;           // device scalar, >= 1, incremented once per step
  long long timeout;                // spin budget in clock64 ticks (peer_timeout_cycles())
  int ll;                           // 1: flag-in-word exchange (every 8-byte word carries its epoch tag: no fence, no flag round trip)
};

// Bytes of the classic regions (flags + data) of one rank's symmetric buffer; the flag-in-word region follows them.
__host__ __device__ __forceinline__ size_t bn_peer_classic_bytes(int world, int nslots, int cmax) {
  return static_cast<size_t>(nslots) * 2 * world * (64 * 8 + 2 * static_cast<size_t>(cmax) * 8);
}
// Flag-in-word ("LL") slot of value v (0 .. 2*cmax) that rank `src` contributes to call site slot_idx, in the buffer at
// `base`: two 8-byte words, each = 32 data bits | (epoch tag << 32).  An 8-byte store is atomic, so a reader that sees the
// tag of this epoch in a word also sees that word's data: no __threadfence_system(), no separate flag write to wait for.
__device__ __forceinline__ unsigned long long* bn_ll_word(unsigned long long base, const BnPeerParams& pp, size_t slot_idx,
                                                          int src, int v) {
  return reinterpret_cast<unsigned long long*>(base + bn_peer_classic_bytes(pp.world, pp.nslots, pp.cmax)) +
         ((slot_idx * pp.world + src) * (2 * static_cast<size_t>(pp.cmax)) + v) * 2;
}
__device__ __forceinline__ void bn_ll_push(const BnPeerParams& pp, size_t slot_idx, int v, double val, unsigned long long epoch) {
  const unsigned long long bits = static_cast<unsigned long long>(__double_as_longlong(val));
  const unsigned long long tag = (epoch & 0xffffffffull) << 32;
  const unsigned long long w0 = (bits & 0xffffffffull) | tag, w1 = (bits >> 32) | tag;
  for (int r = 0; r < pp.world; ++r) {
    unsigned long long* q = bn_ll_word(pp.peers[r], pp, slot_idx, pp.rank, v);
    asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(q), "l"(w0) : "memory");
    asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(q + 1), "l"(w1) : "memory");
  }
}
// Sum over the ranks (rank order: identical on every rank) of value v, spinning until every rank's words carry this epoch.
__device__ __forceinline__ double bn_ll_sum(const BnPeerParams& pp, size_t slot_idx, int v, unsigned long long epoch) {
  const unsigned long long tag = epoch & 0xffffffffull;
  double s = 0.0;
  const long long t_start = clock64();
  for (int r = 0; r < pp.world; ++r) {
    const unsigned long long* q = bn_ll_word(pp.peers[pp.rank], pp, slot_idx, r, v);
    unsigned long long w0, w1;
    for (;;) {
      asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(w0) : "l"(q) : "memory");
      asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(w1) : "l"(q + 1) : "memory");
      if ((w0 >> 32) == tag && (w1 >> 32) == tag) break;
      if (clock64() - t_start > pp.timeout) {
        printf("icsg3d: bn all-reduce (flag-in-word) timeout rank %d slot %d block %d waiting for rank %d (epoch %llu)\n", pp.rank,
               pp.slot, blockIdx.x, r, epoch);
        __trap();
      }
    }
    s += __longlong_as_double(static_cast<long long>((w0 & 0xffffffffull) | (w1 << 32)));
  }
  return s;
}

__device__ __forceinline__ void st_release_sys_u64(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys_u64(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}

__global__ void __launch_bounds__(1024) bn_reduce_allreduce_kernel(const double* __restrict__ partials, int nparts, int C,
                                                                   int mode, double count, const float* __restrict__ gamma,
                                                                   const float* __restrict__ beta, float eps,
                                                                   double* __restrict__ sums, float* __restrict__ mean_out,
                                                                   float* __restrict__ rstd_out, float* __restrict__ scale,
                                                                   float* __restrict__ shift, float* __restrict__ moving_mean,
                                                                   float* __restrict__ moving_var, float momentum,
                                                                   float* __restrict__ dgamma, float* __restrict__ dbeta,
                                                                   const BnPeerParams pp) {
  pdl_prologue();
  __shared__ double red[kRedSlots][2 * kRedCh + 1];
  __shared__ double tot[2 * kRedCh];
  const int col = threadIdx.x & (2 * kRedCh - 1);
  const int slot_r = threadIdx.x >> 4;
  const int c = blockIdx.x * kRedCh + (col & (kRedCh - 1));
  const int gcol = (col >> 3) * C + c;
  const size_t C2 = 2 * static_cast<size_t>(C);
  double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
  if (c < C) {
    int p = slot_r;
    for (; p + 3 * kRedSlots < nparts; p += 4 * kRedSlots) {
      a0 += partials[static_cast<size_t>(p) * C2 + gcol];
      a1 += partials[static_cast<size_t>(p + kRedSlots) * C2 + gcol];
      a2 += partials[static_cast<size_t>(p + 2 * kRedSlots) * C2 + gcol];
      a3 += partials[static_cast<size_t>(p + 3 * kRedSlots) * C2 + gcol];
    }
    for (; p < nparts; p += kRedSlots) a0 += partials[static_cast<size_t>(p) * C2 + gcol];
  }
  red[slot_r][col] = (a0 + a1) + (a2 + a3);
  __syncthreads();
  const unsigned long long epoch = static_cast<unsigned long long>(*pp.epoch);
  const int par = static_cast<int>(epoch & 1ull);
  const size_t flags_bytes = static_cast<size_t>(pp.nslots) * 2 * pp.world * 64 * sizeof(unsigned long long);
  const size_t slot_idx = static_cast<size_t>(pp.slot) * 2 + par;
  if (threadIdx.x < 2 * kRedCh) {
    double t0 = 0.0, t1 = 0.0, t2 = 0.0, t3 = 0.0;
#pragma unroll
    for (int i = 0; i < kRedSlots; i += 4) {
      t0 += red[i][threadIdx.x];
      t1 += red[i + 1][threadIdx.x];
      t2 += red[i + 2][threadIdx.x];
      t3 += red[i + 3][threadIdx.x];
    }
    const double local = (t0 + t1) + (t2 + t3);
    tot[threadIdx.x] = local;  // local sums (dgamma / dbeta in backward mode: the gradient all-reduce adds the ranks)
    if (pp.ll) {
      if (c < C) bn_ll_push(pp, slot_idx, (threadIdx.x >> 3) * pp.cmax + c, local, epoch);
    } else {
      if (c < C) {
        for (int r = 0; r < pp.world; ++r) {  // push into every rank's copy (own included)
          double* data = reinterpret_cast<double*>(pp.peers[r] + flags_bytes) +
                         (slot_idx * pp.world + pp.rank) * (2 * static_cast<size_t>(pp.cmax));
          data[(threadIdx.x >> 3) * pp.cmax + c] = local;
        }
      }
      __threadfence_system();
    }
  }
  double s0 = 0.0, s1 = 0.0;
  if (pp.ll) {
    __shared__ double gl[2 * kRedCh];
    if (threadIdx.x < 2 * kRedCh && c < C) gl[threadIdx.x] = bn_ll_sum(pp, slot_idx, (threadIdx.x >> 3) * pp.cmax + c, epoch);
    __syncthreads();
    if (threadIdx.x >= kRedCh || c >= C) return;
    s0 = gl[threadIdx.x];
    s1 = gl[kRedCh + threadIdx.x];
  } else {
    __syncthreads();
    if (threadIdx.x < pp.world) {
      const int r = threadIdx.x;
      // publish: my contribution for block b is complete on rank r
      unsigned long long* rflag = reinterpret_cast<unsigned long long*>(pp.peers[r]) +
                                  (slot_idx * pp.world + pp.rank) * 64 + blockIdx.x;
      st_release_sys_u64(rflag, epoch);
      // wait: rank r's contribution has arrived in MY buffer
      const unsigned long long* lflag = reinterpret_cast<const unsigned long long*>(pp.peers[pp.rank]) +
                                        (slot_idx * pp.world + r) * 64 + blockIdx.x;
      const long long t_start = clock64();
      while (ld_acquire_sys_u64(lflag) != epoch) {
        if (clock64() - t_start > pp.timeout) {
          printf("icsg3d: bn all-reduce timeout rank %d slot %d block %d waiting for rank %d (epoch %llu)\n", pp.rank, pp.slot,
                 blockIdx.x, r, epoch);
          __trap();
        }
      }
    }
    __syncthreads();
    if (threadIdx.x >= kRedCh || c >= C) return;
    const double* mine = reinterpret_cast<const double*>(pp.peers[pp.rank] + flags_bytes) +
                         slot_idx * pp.world * (2 * static_cast<size_t>(pp.cmax));
    for (int r = 0; r < pp.world; ++r) {  // rank order: identical on every rank
      const volatile double* d = mine + static_cast<size_t>(r) * 2 * pp.cmax;
      s0 += d[c];
      s1 += d[pp.cmax + c];
    }
  }
  if (sums) {
    sums[c] = s0;
    sums[C + c] = s1;
  }
  if (mode == 0) {
    const double mean = s0 / count;
    double var = s1 / count - mean * mean;
    if (var < 0.0) var = 0.0;
    const double rstd = 1.0 / sqrt(var + static_cast<double>(eps));
    const double g = gamma ? static_cast<double>(gamma[c]) : 1.0;
    const double b = beta ? static_cast<double>(beta[c]) : 0.0;
    mean_out[c] = static_cast<float>(mean);
    rstd_out[c] = static_cast<float>(rstd);
    scale[c] = static_cast<float>(g * rstd);
    shift[c] = static_cast<float>(b - mean * g * rstd);
    if (moving_mean) {
      const double var_unbiased = var * (count / (count - (1.0 + static_cast<double>(eps))));
      moving_mean[c] = static_cast<float>(moving_mean[c] - (moving_mean[c] - mean) * (1.0 - momentum));
      moving_var[c] = static_cast<float>(moving_var[c] - (moving_var[c] - var_unbiased) * (1.0 - momentum));
    }
  } else {
    if (dbeta) dbeta[c] = static_cast<float>(tot[threadIdx.x]);
    if (dgamma) dgamma[c] = static_cast<float>(tot[kRedCh + threadIdx.x]);
  }
}

__global__ void bn_inference_coeffs_kernel(const float* __restrict__ gamma, const float* __restrict__ beta,
                                           const float* __restrict__ mm, const float* __restrict__ mv, float eps,
                                           float* __restrict__ scale, float* __restrict__ shift, int C) {
  pdl_prologue();
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const double rstd = 1.0 / sqrt(static_cast<double>(mv[c]) + static_cast<double>(eps));
  const double g = gamma ? gamma[c] : 1.0;
  scale[c] = static_cast<float>(g * rstd);
  shift[c] = static_cast<float>((beta ? beta[c] : 0.0) - mm[c] * g * rstd);
}

// ------------------------------------------------------------------------------------------------
// forward apply
// ------------------------------------------------------------------------------------------------
struct BnFwdParams {
  const void* x;
  int ldx;
  const float* scale;
  const float* shift;
  int act;
  float alpha;
  int post;
  int B, D, H, W, C;  // extents of x
  __nv_bfloat16* y;
  int ldy;
  float* y32;
  int ldy32;
  uint8_t* pool_idx;  // [B*D/2*H/2*W/2][C]
  int split_ctot;     // > 0: y is the fp32-class split tensor [rows][3*split_ctot] = [hi | lo | hi] (split3.cu) and this
  int split_coff;     //      layer's channels start at split_coff inside every part (U-Net skip concatenations)
  int split_fmt;      // 0: bf16 pairs, 1: IEEE fp16 pairs (raw 2-byte values in the same buffer)
};

// BN output vector -> y: plain bf16, or the [hi | lo | hi] split of the fp32 value
template <int V>
__device__ __forceinline__ void bn_store_out(const BnFwdParams& p, long long row, int c, const float (&v)[V]) {
  if (p.split_ctot == 0) {
    store_bf16_vec<V>(p.y + row * p.ldy + c, v);
    return;
  }
  __nv_bfloat16* d = p.y + row * p.ldy + p.split_coff + c;
  if (p.split_fmt == 0) {
    float hi[V], lo[V];
#pragma unroll
    for (int i = 0; i < V; ++i) {
      hi[i] = __bfloat162float(__float2bfloat16_rn(v[i]));
      lo[i] = v[i] - hi[i];
    }
    store_bf16_vec<V>(d, hi);
    store_bf16_vec<V>(d + p.split_ctot, lo);
    store_bf16_vec<V>(d + 2 * p.split_ctot, hi);
  } else {
    unsigned short* u = reinterpret_cast<unsigned short*>(d);
#pragma unroll
    for (int i = 0; i < V; ++i) {
      const __half h = __float2half_rn(v[i]);
      const __half l = __float2half_rn(v[i] - __half2float(h));
      u[i] = __half_as_ushort(h);
      u[p.split_ctot + i] = __half_as_ushort(l);
      u[2 * p.split_ctot + i] = __half_as_ushort(h);
    }
  }
}

template <typename T>
__global__ void __launch_bounds__(kBnThreads, 4) bn_apply_fwd_kernel(const BnFwdParams p) {
  pdl_prologue();
  constexpr int V = VecIO<T>::N;
  const T* x = static_cast<const T*>(p.x);
  const int cg = p.C / V;
  if (p.post == ICSG3D_POST_POOL2) {
    const int Do = p.D / 2, Ho = p.H / 2, Wo = p.W / 2;
    const long long total = static_cast<long long>(p.B) * Do * Ho * Wo * cg;
    for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total;
         idx += static_cast<long long>(gridDim.x) * blockDim.x) {
      const int g = static_cast<int>(idx % cg);
      long long o = idx / cg;
      const int wo = static_cast<int>(o % Wo);
      long long t = o / Wo;
      const int ho = static_cast<int>(t % Ho);
      t /= Ho;
      const int dz = static_cast<int>(t % Do);
      const int n = static_cast<int>(t / Do);
      float sc[V], sh[V], best[V];
      int bi[V];
#pragma unroll
      for (int i = 0; i < V; ++i) {
        sc[i] = p.scale[g * V + i];
        sh[i] = p.shift[g * V + i];
      }
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const int dd = 2 * dz + (k >> 2), hh = 2 * ho + ((k >> 1) & 1), ww = 2 * wo + (k & 1);
        const long long r = ((static_cast<long long>(n) * p.D + dd) * p.H + hh) * p.W + ww;
        float v[V];
        VecIO<T>::load(x + r * p.ldx + g * V, v);
#pragma unroll
        for (int i = 0; i < V; ++i) {
          const float z = act_fwd(fmaf(sc[i], v[i], sh[i]), p.act, p.alpha);
          if (k == 0 || z > best[i]) {  // strict '>' keeps the FIRST maximum in (d,h,w) scan order
            best[i] = z;
            bi[i] = k;
          }
        }
      }
      bn_store_out<V>(p, o, g * V, best);
      if (p.pool_idx) {
        uint8_t* ip = p.pool_idx + o * p.C + g * V;
        if constexpr (V == 8) {
          uint2 q;
          q.x = bi[0] | (bi[1] << 8) | (bi[2] << 16) | (bi[3] << 24);
          q.y = bi[4] | (bi[5] << 8) | (bi[6] << 16) | (bi[7] << 24);
          *reinterpret_cast<uint2*>(ip) = q;
        } else {
          *reinterpret_cast<uint32_t*>(ip) = bi[0] | (bi[1] << 8) | (bi[2] << 16) | (bi[3] << 24);
        }
      }
    }
    return;
  }
  const long long M = static_cast<long long>(p.B) * p.D * p.H * p.W;
  const long long total = M * cg;
  if (p.post == ICSG3D_POST_NONE) {
    // plain BN+activation: 4 independent vectors in flight per thread
    constexpr int U = 4;
    const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
    for (long long i0 = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i0 < total; i0 += U * stride) {
      typename VecIO<T>::raw_t raw[U];
#pragma unroll
      for (int j = 0; j < U; ++j) {
        const long long idx = i0 + j * stride;
        if (idx < total) raw[j] = VecIO<T>::load_raw(x + (idx / cg) * p.ldx + static_cast<int>(idx % cg) * V);
      }
#pragma unroll
      for (int j = 0; j < U; ++j) {
        const long long idx = i0 + j * stride;
        if (idx < total) {
          const int g = static_cast<int>(idx % cg);
          const long long r = idx / cg;
          float v[V];
          VecIO<T>::cvt(raw[j], v);
#pragma unroll
          for (int i = 0; i < V; ++i) v[i] = act_fwd(fmaf(p.scale[g * V + i], v[i], p.shift[g * V + i]), p.act, p.alpha);
          if (p.y) bn_store_out<V>(p, r, g * V, v);
          if (p.y32) {
            float* d32 = p.y32 + r * p.ldy32 + g * V;
#pragma unroll
            for (int i = 0; i < V; ++i) d32[i] = v[i];
          }
        }
      }
    }
    return;
  }
  for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int g = static_cast<int>(idx % cg);
    const long long r = idx / cg;
    float v[V];
    VecIO<T>::load(x + r * p.ldx + g * V, v);
#pragma unroll
    for (int i = 0; i < V; ++i) v[i] = act_fwd(fmaf(p.scale[g * V + i], v[i], p.shift[g * V + i]), p.act, p.alpha);
    if (p.post == ICSG3D_POST_UP2) {
      const int w = static_cast<int>(r % p.W);
      long long t = r / p.W;
      const int h = static_cast<int>(t % p.H);
      t /= p.H;
      const int d = static_cast<int>(t % p.D);
      const int n = static_cast<int>(t / p.D);
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const long long ro = ((static_cast<long long>(n) * 2 * p.D + 2 * d + (k >> 2)) * 2 * p.H + 2 * h + ((k >> 1) & 1)) *
                                 2 * p.W + 2 * w + (k & 1);
        bn_store_out<V>(p, ro, g * V, v);
      }
    } else {
      if (p.y) bn_store_out<V>(p, r, g * V, v);
      if (p.y32) {
        float* d32 = p.y32 + r * p.ldy32 + g * V;
#pragma unroll
        for (int i = 0; i < V; ++i) d32[i] = v[i];
      }
    }
  }
}

// Lean bf16 BN + activation + MaxPool3D(2) forward (plain bf16 output): thread = (window, 8-channel group) with the
// group fixed per thread (scale / shift live in registers), window coordinates in 32-bit arithmetic, the 8 rows of the
// window fetched in one batch from one base pointer, activation fixed at compile time.
template <int kAct>
__global__ void __launch_bounds__(kBnThreads, 3) bn_apply_fwd_pool_lean_kernel(const BnFwdParams p) {
  pdl_prologue();
  constexpr int V = 8;
  typedef __nv_bfloat16 T;
  const T* __restrict__ x = static_cast<const T*>(p.x);
  const int cg = p.C / V;
  const int rpi = kBnThreads / cg;
  const int g = threadIdx.x % cg;
  const int rl = threadIdx.x / cg;
  if (rl >= rpi) return;
  float sc[V], sh[V];
#pragma unroll
  for (int i = 0; i < V; ++i) {
    sc[i] = p.scale[g * V + i];
    sh[i] = p.shift[g * V + i];
  }
  const float alpha = p.alpha;
  const uint32_t Do = p.D / 2, Ho = p.H / 2, Wo = p.W / 2;
  const uint32_t Mo = static_cast<uint32_t>(p.B) * Do * Ho * Wo;
  const uint32_t stride = gridDim.x * rpi;
  const long long oW = p.ldx, oH = static_cast<long long>(p.W) * p.ldx, oD = static_cast<long long>(p.H) * p.W * p.ldx;
  for (uint32_t o = blockIdx.x * rpi + rl; o < Mo; o += stride) {
    const uint32_t t0 = o / Wo, wo = o - t0 * Wo;
    const uint32_t t1 = t0 / Ho, ho = t0 - t1 * Ho;
    const uint32_t n = t1 / Do, dz = t1 - n * Do;
    const long long rbase = ((static_cast<long long>(n) * p.D + 2 * dz) * p.H + 2 * ho) * p.W + 2 * wo;
    const T* xb = x + rbase * p.ldx + g * V;
    uint4 xr[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) xr[k] = *reinterpret_cast<const uint4*>(xb + (k >> 2) * oD + ((k >> 1) & 1) * oH + (k & 1) * oW);
    float best[V];
    uint32_t bi[V];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      float v[V];
      VecIO<T>::cvt(xr[k], v);
#pragma unroll
      for (int i = 0; i < V; ++i) {
        float z = fmaf(sc[i], v[i], sh[i]);
        if (kAct == ICSG3D_ACT_RELU) z = z > 0.f ? z : 0.f;
        if (kAct == ICSG3D_ACT_LEAKY) z = z > 0.f ? z : alpha * z;
        if (k == 0) {
          best[i] = z;
          bi[i] = 0;
        } else if (z > best[i]) {  // strict '>' keeps the FIRST maximum in (d,h,w) scan order
          best[i] = z;
          bi[i] = k;
        }
      }
    }
    VecIO<T>::store(p.y + static_cast<long long>(o) * p.ldy + g * V, best);
    if (p.pool_idx) {
      uint2 q;
      q.x = bi[0] | (bi[1] << 8) | (bi[2] << 16) | (bi[3] << 24);
      q.y = bi[4] | (bi[5] << 8) | (bi[6] << 16) | (bi[7] << 24);
      *reinterpret_cast<uint2*>(p.pool_idx + static_cast<size_t>(o) * p.C + g * V) = q;
    }
  }
}

// Lean bf16 BN + activation (+ UpSampling3D(2)) forward with a plain bf16 output: the channel group is fixed per thread
// (scale / shift in registers instead of 16 table loads per vector), rows walk by pointer increments, 4 rows in flight.
template <int kAct, bool kUp>
__global__ void __launch_bounds__(kBnThreads, 4) bn_apply_fwd_lean_kernel(const BnFwdParams p) {
  pdl_prologue();
  constexpr int V = 8, U = 4;
  typedef __nv_bfloat16 T;
  const T* __restrict__ x = static_cast<const T*>(p.x);
  const int cg = p.C / V;
  const int rpi = kBnThreads / cg;
  const int g = threadIdx.x % cg;
  const int rl = threadIdx.x / cg;
  if (rl >= rpi) return;
  float sc[V], sh[V];
#pragma unroll
  for (int i = 0; i < V; ++i) {
    sc[i] = p.scale[g * V + i];
    sh[i] = p.shift[g * V + i];
  }
  const float alpha = p.alpha;
  const uint32_t M = static_cast<uint32_t>(p.B) * p.D * p.H * p.W;
  const uint32_t stride = gridDim.x * rpi;
  auto finish = [&](const uint4& q, uint32_t r) {
    float v[V];
    VecIO<T>::cvt(q, v);
#pragma unroll
    for (int i = 0; i < V; ++i) {
      float z = fmaf(sc[i], v[i], sh[i]);
      if (kAct == ICSG3D_ACT_RELU) z = z > 0.f ? z : 0.f;
      if (kAct == ICSG3D_ACT_LEAKY) z = z > 0.f ? z : alpha * z;
      v[i] = z;
    }
    uint4 o;
    o.x = pack_bf16x2(v[0], v[1]);
    o.y = pack_bf16x2(v[2], v[3]);
    o.z = pack_bf16x2(v[4], v[5]);
    o.w = pack_bf16x2(v[6], v[7]);
    if constexpr (!kUp) {
      *reinterpret_cast<uint4*>(p.y + static_cast<long long>(r) * p.ldy + g * V) = o;
    } else {
      const uint32_t t0 = r / p.W, w = r - t0 * p.W;
      const uint32_t t1 = t0 / p.H, h = t0 - t1 * p.H;
      const uint32_t n = t1 / p.D, d = t1 - n * p.D;
      const long long W2 = 2ll * p.W, H2 = 2ll * p.H;
      T* ob = p.y + (((static_cast<long long>(n) * 2 * p.D + 2 * d) * H2 + 2 * h) * W2 + 2 * w) * p.ldy + g * V;
      const long long oW = p.ldy, oH = W2 * p.ldy, oD = H2 * W2 * p.ldy;
#pragma unroll
      for (int k = 0; k < 8; ++k) *reinterpret_cast<uint4*>(ob + (k >> 2) * oD + ((k >> 1) & 1) * oH + (k & 1) * oW) = o;
    }
  };
  uint32_t r = blockIdx.x * rpi + rl;
  const T* px = x + static_cast<long long>(r) * p.ldx + g * V;
  const long long sx = static_cast<long long>(stride) * p.ldx;
  for (; static_cast<unsigned long long>(r) + (U - 1ull) * stride < M; r += U * stride) {
    uint4 q[U];
#pragma unroll
    for (int j = 0; j < U; ++j) q[j] = *reinterpret_cast<const uint4*>(px + j * sx);
#pragma unroll
    for (int j = 0; j < U; ++j) finish(q[j], r + j * stride);
    px += U * sx;
  }
  for (; r < M; r += stride) {
    finish(*reinterpret_cast<const uint4*>(px), r);
    px += sx;
  }
}

// ------------------------------------------------------------------------------------------------
// backward
// ------------------------------------------------------------------------------------------------
struct BnBwdParams {
  const void* dy;   // gradient w.r.t. the forward output (post domain)
  int lddy;
  const void* dy2;  // optional second gradient w.r.t. the un-pooled BN output (U-Net skip connection), x's shape
  int lddy2;
  const void* x;    // BN input (pre-normalisation)
  int ldx;
  const float* mean;
  const float* rstd;
  const float* scale;  // gamma * rstd
  const float* shift;
  int act;
  float alpha;
  int post;
  const uint8_t* pool_idx;
  int B, D, H, W, C;   // extents of x
  double* partials;    // reduce: [nblk][2][C]
  const double* sums;  // apply:  [2][C] (sum g, sum g*xhat)
  double count;
  int pre_relu;        // U-Net ordering: x = ReLU(conv); multiply dx by (x > 0)
  const __nv_bfloat16* tap_other;  // DFC tap: dx += tap_coef * (x - tap_other) before the ReLU mask
  int ld_other;
  float tap_coef;
  __nv_bfloat16* dx;
  int lddx;
  int dx_f32;      // 1: dx is an fp32 tensor (fp32-class backward), else bf16
  double* tap_sq;  // optional (apply pass with a tap): per-block sum (x - tap_other)^2, the DFC feature loss of the layer
};

// per-block partial of the fused DFC feature loss: partial[blockIdx.x] = sum over the block's threads
__device__ __forceinline__ void bn_tap_sq_store(double* partials, double v) {
  __shared__ double tap_red[kBnThreads / 32];
  v = warp_sum(v);
  if ((threadIdx.x & 31) == 0) tap_red[threadIdx.x >> 5] = v;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
#pragma unroll
    for (int w = 0; w < kBnThreads / 32; ++w) t += tap_red[w];
    partials[blockIdx.x] = t;
  }
}

// g for one x row given the matching dy vector: g_i = dyv_i * act'(scale*x+shift)
template <int V>
__device__ __forceinline__ void g_from(const float (&xv)[V], const float (&dyv)[V], const float (&sc)[V],
                                       const float (&sh)[V], int act, float alpha, float (&g)[V]) {
#pragma unroll
  for (int i = 0; i < V; ++i) g[i] = dyv[i] * act_grad(fmaf(sc[i], xv[i], sh[i]), act, alpha);
}

template <typename T, bool kApply, int kPost = -1>
__device__ __forceinline__ void bn_bwd_body(const BnBwdParams& p, double* tap_sq_out = nullptr) {
  double tap_sq = 0.0;
  const bool want_sq = kApply && p.tap_sq != nullptr;
  // kPost >= 0 fixes the post-op at compile time (registers are then sized for that path alone), -1 reads p.post
  const int post = kPost >= 0 ? kPost : p.post;
  // HBM-bound: every thread keeps 4 independent rows (x, dy, optional skip gradient / tap) in flight before it
  // touches any of them (load phase, then compute phase), ~8-12 x 16 B per thread.
  constexpr int V = VecIO<T>::N;
  constexpr int U = 4;
  typedef typename VecIO<T>::raw_t raw_t;
  const T* x = static_cast<const T*>(p.x);
  const T* dy = static_cast<const T*>(p.dy);
  const T* dy2 = static_cast<const T*>(p.dy2);
  const int cg = p.C / V;
  // Thread -> channel group is fixed for the whole kernel (the grid strides in units of whole rows).
  const int rpi = kBnThreads / cg;
  const int g = threadIdx.x % cg;
  const int rl = threadIdx.x / cg;
  const bool has_tap = kApply && p.tap_other != nullptr;
  const bool has_dy2 = dy2 != nullptr;
  // reduce: s1 = sum g, s2 = sum g*(x-mean) (rstd applied once at the end)
  // apply : dx = scale*g + ca*x + cb  with  ca = -scale*rstd*k2,  cb = -scale*k1 - ca*mean
  float sc[V], sh[V], mu[V], s1[V], s2[V], ca[V], cb[V];
#pragma unroll
  for (int i = 0; i < V; ++i) {
    sc[i] = p.scale[g * V + i];
    sh[i] = p.shift[g * V + i];
    mu[i] = p.mean[g * V + i];
    s1[i] = s2[i] = 0.f;
    if (kApply) {
      const float k1 = static_cast<float>(p.sums[g * V + i] / p.count);
      const float k2 = static_cast<float>(p.sums[p.C + g * V + i] / p.count);
      ca[i] = -sc[i] * p.rstd[g * V + i] * k2;
      cb[i] = -sc[i] * k1 - ca[i] * mu[i];
    }
  }

  auto emit = [&](long long r, const float (&xv)[V], const float (&gv)[V], const raw_t& tapraw) {
    if (kApply) {
      float o[V];
#pragma unroll
      for (int i = 0; i < V; ++i) o[i] = fmaf(sc[i], gv[i], fmaf(ca[i], xv[i], cb[i]));
      if (has_tap) {
        float ov[V];
        VecIO<T>::cvt(tapraw, ov);
        float sq = 0.f;
#pragma unroll
        for (int i = 0; i < V; ++i) {
          const float d = xv[i] - ov[i];
          o[i] += p.tap_coef * d;
          sq = fmaf(d, d, sq);
        }
        if (want_sq) tap_sq += static_cast<double>(sq);
      }
      if (p.pre_relu) {
#pragma unroll
        for (int i = 0; i < V; ++i) o[i] = xv[i] > 0.f ? o[i] : 0.f;
      }
      if (p.dx_f32) {
        float* d32 = reinterpret_cast<float*>(p.dx) + r * p.lddx + g * V;
#pragma unroll
        for (int i = 0; i < V; i += 4) *reinterpret_cast<float4*>(d32 + i) = make_float4(o[i], o[i + 1], o[i + 2], o[i + 3]);
      } else {
        store_bf16_vec<V>(p.dx + r * p.lddx + g * V, o);
      }
    } else {
#pragma unroll
      for (int i = 0; i < V; ++i) {
        s1[i] += gv[i];
        s2[i] = fmaf(gv[i], xv[i] - mu[i], s2[i]);
      }
    }
  };
  const T* tap = reinterpret_cast<const T*>(p.tap_other);

  if (rl < rpi) {
    if (post == ICSG3D_POST_POOL2) {
      const int Do = p.D / 2, Ho = p.H / 2, Wo = p.W / 2;
      const long long Mo = static_cast<long long>(p.B) * Do * Ho * Wo;
      for (long long o = static_cast<long long>(blockIdx.x) * rpi + rl; o < Mo; o += static_cast<long long>(gridDim.x) * rpi) {
        const int wo = static_cast<int>(o % Wo);
        long long t = o / Wo;
        const int ho = static_cast<int>(t % Ho);
        t /= Ho;
        const int dz = static_cast<int>(t % Do);
        const int n = static_cast<int>(t / Do);
        const raw_t dyr = VecIO<T>::load_raw(dy + o * p.lddy + g * V);
        int bi[V];
        {
          const uint8_t* ip = p.pool_idx + o * p.C + g * V;
          if constexpr (V == 8) {
            const uint2 q = *reinterpret_cast<const uint2*>(ip);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              bi[i] = (q.x >> (8 * i)) & 0xff;
              bi[4 + i] = (q.y >> (8 * i)) & 0xff;
            }
          } else {
            const uint32_t q = *reinterpret_cast<const uint32_t*>(ip);
#pragma unroll
            for (int i = 0; i < 4; ++i) bi[i] = (q >> (8 * i)) & 0xff;
          }
        }
        float dyv[V];
        VecIO<T>::cvt(dyr, dyv);
#pragma unroll
        for (int k0 = 0; k0 < 8; k0 += U) {
          raw_t xr[U], er[U], tr[U];
          long long rr[U];
#pragma unroll
          for (int j = 0; j < U; ++j) {
            const int k = k0 + j;
            const int dd = 2 * dz + (k >> 2), hh = 2 * ho + ((k >> 1) & 1), ww = 2 * wo + (k & 1);
            rr[j] = ((static_cast<long long>(n) * p.D + dd) * p.H + hh) * p.W + ww;
            xr[j] = VecIO<T>::load_raw(x + rr[j] * p.ldx + g * V);
            if (has_dy2) er[j] = VecIO<T>::load_raw(dy2 + rr[j] * p.lddy2 + g * V);
            if (has_tap) tr[j] = VecIO<T>::load_raw(tap + rr[j] * p.ld_other + g * V);
          }
#pragma unroll
          for (int j = 0; j < U; ++j) {
            const int k = k0 + j;
            float xv[V], sel[V], gv[V];
            VecIO<T>::cvt(xr[j], xv);
#pragma unroll
            for (int i = 0; i < V; ++i) sel[i] = bi[i] == k ? dyv[i] : 0.f;
            if (has_dy2) {
              float e[V];
              VecIO<T>::cvt(er[j], e);
#pragma unroll
              for (int i = 0; i < V; ++i) sel[i] += e[i];
            }
            g_from<V>(xv, sel, sc, sh, p.act, p.alpha, gv);
            emit(rr[j], xv, gv, tr[j]);
          }
        }
      }
    } else if (post == ICSG3D_POST_UP2) {
      const long long M = static_cast<long long>(p.B) * p.D * p.H * p.W;
      for (long long r = static_cast<long long>(blockIdx.x) * rpi + rl; r < M; r += static_cast<long long>(gridDim.x) * rpi) {
        const int w = static_cast<int>(r % p.W);
        long long t = r / p.W;
        const int h = static_cast<int>(t % p.H);
        t /= p.H;
        const int d = static_cast<int>(t % p.D);
        const int n = static_cast<int>(t / p.D);
        raw_t cr[8], er, tr;
        const raw_t xr = VecIO<T>::load_raw(x + r * p.ldx + g * V);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const long long ro = ((static_cast<long long>(n) * 2 * p.D + 2 * d + (k >> 2)) * 2 * p.H + 2 * h + ((k >> 1) & 1)) *
                                   2 * p.W + 2 * w + (k & 1);
          cr[k] = VecIO<T>::load_raw(dy + ro * p.lddy + g * V);
        }
        if (has_dy2) er = VecIO<T>::load_raw(dy2 + r * p.lddy2 + g * V);
        if (has_tap) tr = VecIO<T>::load_raw(tap + r * p.ld_other + g * V);
        float xv[V], dyv[V], gv[V];
        VecIO<T>::cvt(xr, xv);
#pragma unroll
        for (int i = 0; i < V; ++i) dyv[i] = 0.f;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          float c[V];
          VecIO<T>::cvt(cr[k], c);
#pragma unroll
          for (int i = 0; i < V; ++i) dyv[i] += c[i];
        }
        if (has_dy2) {
          float e[V];
          VecIO<T>::cvt(er, e);
#pragma unroll
          for (int i = 0; i < V; ++i) dyv[i] += e[i];
        }
        g_from<V>(xv, dyv, sc, sh, p.act, p.alpha, gv);
        emit(r, xv, gv, tr);
      }
    } else {
      const long long M = static_cast<long long>(p.B) * p.D * p.H * p.W;
      const long long stride = static_cast<long long>(gridDim.x) * rpi;
      for (long long r0 = static_cast<long long>(blockIdx.x) * rpi + rl; r0 < M; r0 += U * stride) {
        raw_t xr[U], dr[U], er[U], tr[U];
#pragma unroll
        for (int j = 0; j < U; ++j) {
          const long long r = r0 + j * stride;
          if (r < M) {
            xr[j] = VecIO<T>::load_raw(x + r * p.ldx + g * V);
            dr[j] = VecIO<T>::load_raw(dy + r * p.lddy + g * V);
            if (has_dy2) er[j] = VecIO<T>::load_raw(dy2 + r * p.lddy2 + g * V);
            if (has_tap) tr[j] = VecIO<T>::load_raw(tap + r * p.ld_other + g * V);
          }
        }
#pragma unroll
        for (int j = 0; j < U; ++j) {
          const long long r = r0 + j * stride;
          if (r < M) {
            float xv[V], dyv[V], gv[V];
            VecIO<T>::cvt(xr[j], xv);
            VecIO<T>::cvt(dr[j], dyv);
            if (has_dy2) {
              float e[V];
              VecIO<T>::cvt(er[j], e);
#pragma unroll
              for (int i = 0; i < V; ++i) dyv[i] += e[i];
            }
            g_from<V>(xv, dyv, sc, sh, p.act, p.alpha, gv);
            emit(r, xv, gv, tr[j]);
          }
        }
      }
    }
  }
  if (tap_sq_out) *tap_sq_out = tap_sq;
  if (!kApply) {
    extern __shared__ double sred[];
    double* out = p.partials + static_cast<size_t>(blockIdx.x) * 2 * p.C;
    for (int pass = 0; pass < 2; ++pass) {
      __syncthreads();
      if (rl < rpi) {
#pragma unroll
        for (int i = 0; i < V; ++i)
          sred[rl * p.C + g * V + i] = pass == 0 ? static_cast<double>(s1[i])
                                                 : static_cast<double>(s2[i]) * static_cast<double>(p.rstd[g * V + i]);
      }
      __syncthreads();
      for (int c = threadIdx.x; c < p.C; c += kBnThreads) {
        double a = 0.0;
        for (int r = 0; r < rpi; ++r) a += sred[r * p.C + c];
        out[pass * p.C + c] = a;
      }
    }
  }
}

template <bool kApply, int kPost>
constexpr int kBnBwdMinBlocks() { return 2; }

template <typename T, bool kApply, int kPost, int kMinBlocks>
__global__ void __launch_bounds__(kBnThreads, kMinBlocks) bn_bwd_kernel(const BnBwdParams p) {
  pdl_prologue();
  double tap_sq = 0.0;
  bn_bwd_body<T, kApply, kPost>(p, &tap_sq);
  if (kApply && p.tap_sq) bn_tap_sq_store(p.tap_sq, tap_sq);
}

// ------------------------------------------------------------------------------------------------
// Lean bf16 backward passes for the layers that carry the step's BatchNorm traffic (no skip gradient; post = none or
// MaxPool3D(2)).  Same arithmetic as bn_bwd_body, restructured for the memory system:
//   * one instantiation per (pass, post, activation) so that each keeps only the per-channel state it needs
//     (24-40 registers instead of 56) and every load of an iteration is issued before the first use;
//   * pool: the 8 rows of a window are fetched in one batch from one base pointer; only the arg-max row of a channel
//     carries gradient, so the reduce pass picks that element with a 3-level select on the argmax bits (13
//     instructions per channel and WINDOW instead of 5 per channel and ROW) and the apply pass adds scale*g to the
//     one matching row;
//   * window coordinates in 32-bit arithmetic, row pointers advanced by a constant stride.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint4 ldg16(const __nv_bfloat16* p) { return *reinterpret_cast<const uint4*>(p); }
__device__ __forceinline__ uint32_t word_of(const uint4& q, int j) { return j == 0 ? q.x : j == 1 ? q.y : j == 2 ? q.z : q.w; }
__device__ __forceinline__ void cvt8(const uint4& q, float (&v)[8]) { VecIO<__nv_bfloat16>::cvt(q, v); }

template <bool kApply, int kPost, bool kHasAct, int kU = 4>
__global__ void __launch_bounds__(kBnThreads, (kApply || kU == 8) ? 2 : 3)
bn_bwd_lean_kernel(const BnBwdParams p) {
  pdl_prologue();
  constexpr int V = 8;
  typedef __nv_bfloat16 T;
  const T* __restrict__ x = static_cast<const T*>(p.x);
  const T* __restrict__ dy = static_cast<const T*>(p.dy);
  const T* __restrict__ tap = p.tap_other;
  const int cg = p.C / V;
  const int rpi = kBnThreads / cg;
  const int g = threadIdx.x % cg;
  const int rl = threadIdx.x / cg;
  const bool has_tap = kApply && tap != nullptr;
  const bool pre_relu = p.pre_relu != 0;
  const float tap_coef = p.tap_coef;
  const float slope = p.act == ICSG3D_ACT_RELU ? 0.f : p.alpha;  // act'(z) = z > 0 ? 1 : slope
  const bool want_sq = kApply && p.tap_sq != nullptr;
  double tap_sq = 0.0;
  float tap_sqf = 0.f;
  auto tap_flush = [&]() {
    if (want_sq) tap_sq += static_cast<double>(tap_sqf);
    tap_sqf = 0.f;
  };
  float sc[V], sh[V], mu[V], s1[V], s2[V], ca[V], cb[V];
#pragma unroll
  for (int i = 0; i < V; ++i) {
    sc[i] = p.scale[g * V + i];
    sh[i] = p.shift[g * V + i];
    mu[i] = p.mean[g * V + i];
    s1[i] = s2[i] = 0.f;
    if (kApply) {
      const float k1 = static_cast<float>(p.sums[g * V + i] / p.count);
      const float k2 = static_cast<float>(p.sums[p.C + g * V + i] / p.count);
      ca[i] = -sc[i] * p.rstd[g * V + i] * k2;
      cb[i] = -sc[i] * k1 - ca[i] * mu[i];
    }
  }
  // dx row from x row, g row (apply) -- or the running sums (reduce)
  auto finish = [&](const float (&xv)[V], float (&o)[V], const uint4& tq, T* out) {
    if (has_tap) {
      float ov[V];
      cvt8(tq, ov);
      float sq = 0.f;
#pragma unroll
      for (int i = 0; i < V; ++i) {
        const float d = xv[i] - ov[i];
        o[i] += tap_coef * d;
        sq = fmaf(d, d, sq);
      }
      tap_sqf += sq;  // fused DFC feature loss of the layer: fp32 within one loop iteration (<= 64 terms), then fp64
    }
    if (pre_relu) {
#pragma unroll
      for (int i = 0; i < V; ++i) o[i] = xv[i] > 0.f ? o[i] : 0.f;
    }
    VecIO<T>::store(out, o);
  };

  if (rl < rpi) {
    if constexpr (kPost == ICSG3D_POST_NONE) {
      constexpr int U = kU;
      const long long M = static_cast<long long>(p.B) * p.D * p.H * p.W;
      const long long stride = static_cast<long long>(gridDim.x) * rpi;
      long long r = static_cast<long long>(blockIdx.x) * rpi + rl;
      const T* px = x + r * p.ldx + g * V;
      const T* pd = dy + r * p.lddy + g * V;
      const T* pt = has_tap ? tap + r * p.ld_other + g * V : nullptr;
      T* po = kApply ? p.dx + r * p.lddx + g * V : nullptr;
      const long long sx = stride * p.ldx, sd = stride * p.lddy, st = stride * p.ld_other, so = stride * p.lddx;
      auto row = [&](const uint4& xq, const uint4& dq, const uint4& tq, T* out) {
        float xv[V], gv[V];
        cvt8(xq, xv);
        cvt8(dq, gv);
        if constexpr (kHasAct) {
#pragma unroll
          for (int i = 0; i < V; ++i) gv[i] *= fmaf(sc[i], xv[i], sh[i]) > 0.f ? 1.f : slope;
        }
        if constexpr (kApply) {
          float o[V];
#pragma unroll
          for (int i = 0; i < V; ++i) o[i] = fmaf(sc[i], gv[i], fmaf(ca[i], xv[i], cb[i]));
          finish(xv, o, tq, out);
        } else {
#pragma unroll
          for (int i = 0; i < V; ++i) {
            s1[i] += gv[i];
            s2[i] = fmaf(gv[i], xv[i] - mu[i], s2[i]);
          }
        }
      };
      for (; r + (U - 1) * stride < M; r += U * stride) {
        uint4 xr[U], dr[U], tr[U];
#pragma unroll
        for (int j = 0; j < U; ++j) {
          xr[j] = ldg16(px + j * sx);
          dr[j] = ldg16(pd + j * sd);
          if (has_tap) tr[j] = ldg16(pt + j * st);
        }
#pragma unroll
        for (int j = 0; j < U; ++j) row(xr[j], dr[j], tr[j], kApply ? po + j * so : nullptr);
        if (kApply) tap_flush();
        px += U * sx;
        pd += U * sd;
        if (has_tap) pt += U * st;
        if (kApply) po += U * so;
      }
      for (; r < M; r += stride) {
        uint4 tq = make_uint4(0, 0, 0, 0);
        if (has_tap) tq = ldg16(pt);
        row(ldg16(px), ldg16(pd), tq, po);
        if (kApply) tap_flush();
        px += sx;
        pd += sd;
        if (has_tap) pt += st;
        if (kApply) po += so;
      }
    } else {  // MaxPool3D(2)
      const uint32_t Do = p.D / 2, Ho = p.H / 2, Wo = p.W / 2;
      const uint32_t Mo = static_cast<uint32_t>(p.B) * Do * Ho * Wo;
      const uint32_t stride = gridDim.x * rpi;
      const long long oW = p.ldx, oH = static_cast<long long>(p.W) * p.ldx, oD = static_cast<long long>(p.H) * p.W * p.ldx;
      for (uint32_t o = blockIdx.x * rpi + rl; o < Mo; o += stride) {
        const uint32_t t0 = o / Wo, wo = o - t0 * Wo;
        const uint32_t t1 = t0 / Ho, ho = t0 - t1 * Ho;
        const uint32_t n = t1 / Do, dz = t1 - n * Do;
        const long long rbase = ((static_cast<long long>(n) * p.D + 2 * dz) * p.H + 2 * ho) * p.W + 2 * wo;
        const T* xb = x + rbase * p.ldx + g * V;
        uint4 xr[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) xr[k] = ldg16(xb + (k >> 2) * oD + ((k >> 1) & 1) * oH + (k & 1) * oW);
        const uint4 dyr = ldg16(dy + static_cast<long long>(o) * p.lddy + g * V);
        const uint2 q = *reinterpret_cast<const uint2*>(p.pool_idx + static_cast<size_t>(o) * p.C + g * V);
        float gw[V];
        cvt8(dyr, gw);
        if constexpr (kHasAct || !kApply) {
          // x of the arg-max row per channel: 3-level select on the bits of the window index (k = kd*4 + kh*2 + kw)
#pragma unroll
          for (int i = 0; i < V; ++i) {
            const uint32_t qq = i < 4 ? q.x : q.y;
            const int b = 8 * (i & 3);
            const bool b0 = (qq & (1u << b)) != 0, b1 = (qq & (2u << b)) != 0, b2 = (qq & (4u << b)) != 0;
            const int j = i >> 1;
            const uint32_t a0 = b0 ? word_of(xr[1], j) : word_of(xr[0], j), a1 = b0 ? word_of(xr[3], j) : word_of(xr[2], j);
            const uint32_t a2 = b0 ? word_of(xr[5], j) : word_of(xr[4], j), a3 = b0 ? word_of(xr[7], j) : word_of(xr[6], j);
            const uint32_t c0 = b1 ? a1 : a0, c1 = b1 ? a3 : a2;
            const uint32_t e = b2 ? c1 : c0;
            const float xs = __uint_as_float((i & 1) ? (e & 0xffff0000u) : (e << 16));
            if constexpr (kHasAct) gw[i] *= fmaf(sc[i], xs, sh[i]) > 0.f ? 1.f : slope;
            if constexpr (!kApply) {
              s1[i] += gw[i];
              s2[i] = fmaf(gw[i], xs - mu[i], s2[i]);
            }
          }
        }
        if constexpr (kApply) {
          const T* tb = has_tap ? tap + rbase * p.ld_other + g * V : nullptr;
          T* ob = p.dx + rbase * p.lddx + g * V;
          const long long tW = p.ld_other, tH = static_cast<long long>(p.W) * p.ld_other,
                          tD = static_cast<long long>(p.H) * p.W * p.ld_other;
          const long long dW = p.lddx, dH = static_cast<long long>(p.W) * p.lddx, dD = static_cast<long long>(p.H) * p.W * p.lddx;
          float dsc[V];
#pragma unroll
          for (int i = 0; i < V; ++i) dsc[i] = sc[i] * gw[i];
#pragma unroll
          for (int k0 = 0; k0 < 8; k0 += 4) {
            uint4 tr[4];
            if (has_tap) {
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const int k = k0 + j;
                tr[j] = ldg16(tb + (k >> 2) * tD + ((k >> 1) & 1) * tH + (k & 1) * tW);
              }
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const int k = k0 + j;
              float xv[V], ov[V];
              cvt8(xr[k], xv);
#pragma unroll
              for (int i = 0; i < V; ++i) {
                const uint32_t qq = i < 4 ? q.x : q.y;
                const bool hit = ((qq >> (8 * (i & 3))) & 0xffu) == static_cast<uint32_t>(k);
                ov[i] = fmaf(ca[i], xv[i], cb[i]) + (hit ? dsc[i] : 0.f);
              }
              finish(xv, ov, tr[j], ob + (k >> 2) * dD + ((k >> 1) & 1) * dH + (k & 1) * dW);
            }
          }
          tap_flush();
        }
      }
    }
  }
  if (kApply && p.tap_sq) bn_tap_sq_store(p.tap_sq, tap_sq);
  if (!kApply) {
    extern __shared__ double sred[];
    double* out = p.partials + static_cast<size_t>(blockIdx.x) * 2 * p.C;
    for (int pass = 0; pass < 2; ++pass) {
      __syncthreads();
      if (rl < rpi) {
#pragma unroll
        for (int i = 0; i < V; ++i)
          sred[rl * p.C + g * V + i] = pass == 0 ? static_cast<double>(s1[i])
                                                 : static_cast<double>(s2[i]) * static_cast<double>(p.rstd[g * V + i]);
      }
      __syncthreads();
      for (int c = threadIdx.x; c < p.C; c += kBnThreads) {
        double a = 0.0;
        for (int r = 0; r < rpi; ++r) a += sred[r * p.C + c];
        out[pass * p.C + c] = a;
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Whole BatchNorm backward of one layer in ONE cooperative launch (grid = co-resident blocks):
//   phase A  sum g, sum g*xhat partials per block        (reads dy, x)
//   grid sync
//   phase B  blocks 0..C/8-1 reduce the partials of 8 channels each in a fixed order, [data parallel: exchange the 16
//            sums with the other ranks over NVLink peer memory, as bn_reduce_allreduce_kernel], write the sums and the
//            LOCAL dgamma / dbeta
//   grid sync
//   phase C  dx                                          (re-reads dy, x: from L2 when the layer fits)
// Replaces bn_bwd_reduce + bn_reduce_grads (or bn_reduce_allreduce_grads) + bn_bwd_apply: two launches fewer per layer and
// the second read of the <= 100 MB layers comes out of the 126 MB L2 instead of HBM.
// ------------------------------------------------------------------------------------------------
struct BnFusedExtra {
  double* sums;   // [2][C] global sums (written in phase B, read in phase C)
  float* dgamma;
  float* dbeta;
  BnPeerParams pp;  // world == 1: no exchange
};

template <typename T>
__global__ void __launch_bounds__(kBnThreads, 2) bn_bwd_fused_kernel(BnBwdParams p, const BnFusedExtra e) {
  pdl_prologue();
  cooperative_groups::grid_group grid = cooperative_groups::this_grid();
  bn_bwd_body<T, false>(p);
  __threadfence();
  grid.sync();
  {
    __shared__ double red[16][17];
    __shared__ double tot[16];
    const int nblk = (p.C + kRedCh - 1) / kRedCh;
    if (static_cast<int>(blockIdx.x) < nblk) {
      const int col = threadIdx.x & 15;  // 0..7 sum g, 8..15 sum g*xhat
      const int slot = threadIdx.x >> 4;  // 16 slots
      const int c = blockIdx.x * kRedCh + (col & 7);
      const int gcol = (col >> 3) * p.C + c;
      const size_t C2 = 2 * static_cast<size_t>(p.C);
      const int nparts = gridDim.x;
      double a0 = 0.0, a1 = 0.0;
      if (c < p.C) {
        int q = slot;
        for (; q + 16 < nparts; q += 32) {
          a0 += p.partials[static_cast<size_t>(q) * C2 + gcol];
          a1 += p.partials[static_cast<size_t>(q + 16) * C2 + gcol];
        }
        if (q < nparts) a0 += p.partials[static_cast<size_t>(q) * C2 + gcol];
      }
      red[slot][col] = a0 + a1;
      __syncthreads();
      if (threadIdx.x < 16) {
        double t0 = 0.0, t1 = 0.0;
#pragma unroll
        for (int i = 0; i < 16; i += 2) {
          t0 += red[i][threadIdx.x];
          t1 += red[i + 1][threadIdx.x];
        }
        tot[threadIdx.x] = t0 + t1;
      }
      __syncthreads();
      const BnPeerParams& pp = e.pp;
      if (pp.world > 1) {
        const unsigned long long epoch = static_cast<unsigned long long>(*pp.epoch);
        const int par = static_cast<int>(epoch & 1ull);
        const size_t flags_bytes = static_cast<size_t>(pp.nslots) * 2 * pp.world * 64 * sizeof(unsigned long long);
        const size_t slot_idx = static_cast<size_t>(pp.slot) * 2 + par;
        if (threadIdx.x < 16) {
          const int cc = blockIdx.x * kRedCh + (threadIdx.x & 7);
          if (cc < p.C) {
            for (int r = 0; r < pp.world; ++r) {
              double* data = reinterpret_cast<double*>(pp.peers[r] + flags_bytes) +
                             (slot_idx * pp.world + pp.rank) * (2 * static_cast<size_t>(pp.cmax));
              data[(threadIdx.x >> 3) * pp.cmax + cc] = tot[threadIdx.x];
            }
          }
          __threadfence_system();
        }
        __syncthreads();
        if (threadIdx.x < pp.world) {
          const int r = threadIdx.x;
          unsigned long long* rflag = reinterpret_cast<unsigned long long*>(pp.peers[r]) +
                                      (slot_idx * pp.world + pp.rank) * 64 + blockIdx.x;
          st_release_sys_u64(rflag, epoch);
          const unsigned long long* lflag = reinterpret_cast<const unsigned long long*>(pp.peers[pp.rank]) +
                                            (slot_idx * pp.world + r) * 64 + blockIdx.x;
          const long long t_start = clock64();
          while (ld_acquire_sys_u64(lflag) != epoch) {
            if (clock64() - t_start > pp.timeout) {
              printf("icsg3d: bn backward all-reduce timeout rank %d slot %d block %d waiting for rank %d\n", pp.rank, pp.slot,
                     blockIdx.x, r);
              __trap();
            }
          }
        }
        __syncthreads();
      }
      if (threadIdx.x < kRedCh) {
        const int cc = blockIdx.x * kRedCh + threadIdx.x;
        if (cc < p.C) {
          double s0 = tot[threadIdx.x], s1 = tot[kRedCh + threadIdx.x];
          if (e.dbeta) e.dbeta[cc] = static_cast<float>(s0);
          if (e.dgamma) e.dgamma[cc] = static_cast<float>(s1);
          if (pp.world > 1) {
            const size_t flags_bytes = static_cast<size_t>(pp.nslots) * 2 * pp.world * 64 * sizeof(unsigned long long);
            const unsigned long long epoch = static_cast<unsigned long long>(*pp.epoch);
            const size_t slot_idx = static_cast<size_t>(pp.slot) * 2 + (epoch & 1ull);
            const double* mine = reinterpret_cast<const double*>(pp.peers[pp.rank] + flags_bytes) +
                                 slot_idx * pp.world * (2 * static_cast<size_t>(pp.cmax));
            s0 = 0.0;
            s1 = 0.0;
            for (int r = 0; r < pp.world; ++r) {
              const volatile double* d = mine + static_cast<size_t>(r) * 2 * pp.cmax;
              s0 += d[cc];
              s1 += d[pp.cmax + cc];
            }
          }
          e.sums[cc] = s0;
          e.sums[p.C + cc] = s1;
        }
      }
    }
  }
  __threadfence();
  grid.sync();
  p.sums = e.sums;
  bn_bwd_body<T, true>(p);
}

// dgamma = sum g*xhat, dbeta = sum g  (fp64 sums -> fp32 gradient slots)
__global__ void bn_param_grads_kernel(const double* __restrict__ sums, float* __restrict__ dgamma,
                                      float* __restrict__ dbeta, int C) {
  pdl_prologue();
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  if (dbeta) dbeta[c] = static_cast<float>(sums[c]);
  if (dgamma) dgamma[c] = static_cast<float>(sums[C + c]);
}

static int bn_grid(long long rows, int C, int V, int per_sm = 4) {
  const int cg = C / V;
  const int rpi = kBnThreads / cg;
  long long blocks = (rows + rpi - 1) / rpi;
  int sms = sm_count();
  if (sms <= 0) sms = 148;
  const long long cap = static_cast<long long>(sms) * per_sm;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return static_cast<int>(blocks);
}

static size_t bn_red_smem(int C, int V) {
  const int cg = C / V;
  const int rpi = kBnThreads / cg;
  return static_cast<size_t>(rpi) * C * sizeof(double);
}

static bool bn_shape_ok(int C, int dtype) {
  const int V = dtype == ICSG3D_DT_BF16 ? 8 : 4;
  return C > 0 && C % V == 0 && C / V <= kBnThreads;
}

}  // namespace icsg3d

using namespace icsg3d;

extern "C" int icsg3d_bn_nparts(int64_t rows, int C, int dtype) {
  if (!bn_shape_ok(C, dtype)) return -1;
  return bn_grid(rows, C, dtype == ICSG3D_DT_BF16 ? 8 : 4);
}

extern "C" int icsg3d_bn_stats(const void* x, int ldx, int dtype, int64_t rows, int C, double* partials, int nparts,
                               void* stream) {
  ICSG_REQUIRE(x && partials, "bn_stats: null pointer");
  ICSG_REQUIRE(bn_shape_ok(C, dtype), "bn_stats: unsupported C=%d for dtype %d", C, dtype);
  const int V = dtype == ICSG3D_DT_BF16 ? 8 : 4;
  ICSG_REQUIRE(ldx % V == 0, "bn_stats: ldx must be a multiple of %d", V);
  ICSG_REQUIRE(nparts == bn_grid(rows, C, V), "bn_stats: nparts %d != icsg3d_bn_nparts() %d", nparts, bn_grid(rows, C, V));
  const size_t smem = bn_red_smem(C, V);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (dtype == ICSG3D_DT_BF16) {
    if (smem > 48 * 1024) ICSG_CUDA(cudaFuncSetAttribute(bn_stats_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
    launch_k(bn_stats_kernel<__nv_bfloat16>, nparts, kBnThreads, smem, st, static_cast<const __nv_bfloat16*>(x), ldx, rows, C, partials);
  } else {
    launch_k(bn_stats_kernel<float>, nparts, kBnThreads, smem, st, static_cast<const float*>(x), ldx, rows, C, partials);
  }
  ICSG_CHECK_LAUNCH();
  return ICSG3D_OK;
}

extern "C" int icsg3d_bn_reduce_partials(const double* partials, int nparts, int C, double* sums, void* stream) {
  ICSG_REQUIRE(partials && sums && nparts > 0, "bn_reduce_partials: bad arguments");
  launch_k(bn_reduce_partials_kernel, ceil_div(2 * C, 32), 256, 0, static_cast<cudaStream_t>(stream), partials, nparts, 2 * C, sums);
  ICSG_CHECK_LAUNCH();
  return ICSG3D_OK;
}

extern "C" int icsg3d_bn_finalize(const double* sums, double count, const float* gamma, const float* beta, float eps,
                                  float* mean, float* rstd, float* scale, float* shift, float* moving_mean,
                                  float* moving_var, float momentum, int C, void* stream) {
  ICSG_REQUIRE(sums && mean && rstd && scale && shift && count > 0, "bn_finalize: bad arguments");
  ICSG_REQUIRE((moving_mean == nullptr) == (moving_var == nullptr), "bn_finalize: moving_mean/var must both be given");
  launch_k(bn_finalize_kernel, ceil_div(C, 128), 128, 0, static_cast<cudaStream_t>(stream), 
      sums, count, gamma, beta, eps, mean, rstd, scale, shift, moving_mean, moving_var, momentum, C);
  ICSG_CHECK_LAUNCH();
  return ICSG3D_OK;
}

extern "C" int icsg3d_bn_inference_coeffs(const float* gamma, const float* beta, const float* moving_mean,
                                          const float* moving_var, float eps, float* scale, float* shift, int C,
                                          void* stream) {
  ICSG_REQUIRE(moving_mean && moving_var && scale && shift, "bn_inference_coeffs: null pointer");
  launch_k(bn_inference_coeffs_kernel, ceil_div(C, 128), 128, 0, static_cast<cudaStream_t>(stream), gamma, beta, moving_mean, moving_var,
                                                                                        eps, scale, shift, C);
  ICSG_CHECK_LAUNCH();
  return ICSG3D_OK;
}

static int bn_apply_fwd_impl(const void* x, int ldx, int x_dtype, const float* scale, const float* shift, int act,
                             float alpha, int post, int B, int D, int H, int W, int C, void* y, int ldy,
                             float* y32, int ldy32, uint8_t* pool_idx, int split_ctot, int split_coff, int split_fmt,
                             void* stream) {
  ICSG_REQUIRE(x && scale && shift && (y || y32), "bn_apply_fwd: null pointer");
  ICSG_REQUIRE(bn_shape_ok(C, x_dtype), "bn_apply_fwd: unsupported C=%d for dtype %d", C, x_dtype);
  const int V = x_dtype == ICSG3D_DT_BF16 ? 8 : 4;
  ICSG_REQUIRE(ldx % V == 0 && (!y || ldy % V == 0), "bn_apply_fwd: ld must be a multiple of the vector width %d", V);
  ICSG_REQUIRE(post == ICSG3D_POST_NONE || y, "bn_apply_fwd: pool/upsample need the bf16 output");
  ICSG_REQUIRE(post != ICSG3D_POST_POOL2 || (D % 2 == 0 && H % 2 == 0 && W % 2 == 0), "bn_apply_fwd: odd extent for pool");
  ICSG_REQUIRE(split_ctot == 0 || (y && split_coff >= 0 && split_coff + C <= split_ctot && ldy >= 3 * split_ctot &&
                                   split_ctot % V == 0 && split_coff % V == 0),
               "bn_apply_fwd: bad split3 layout (ctot %d, coff %d, C %d, ldy %d)", split_ctot, split_coff, C, ldy);
  BnFwdParams p{x, ldx, scale, shift, act, alpha, post, B, D, H, W, C, static_cast<__nv_bfloat16*>(y), ldy, y32, ldy32, pool_idx,
                split_ctot, split_coff, split_fmt};
  long long items = static_cast<long long>(B) * D * H * W * (C / V);
  if (post == ICSG3D_POST_POOL2) items /= 8;
  long long blocks = (items + kBnThreads - 1) / kBnThreads;
  int sms = sm_count();
  if (sms <= 0) sms = 148;
  if (blocks > static_cast<long long>(sms) * 8) blocks = static_cast<long long>(sms) * 8;
  if (blocks < 1) blocks = 1;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  static const bool lean = [] { const char* e = getenv("ICSG3D_BN_FWD_LEAN"); return e ? atoi(e) != 0 : true; }();
  if (lean && x_dtype == ICSG3D_DT_BF16 && post == ICSG3D_POST_POOL2 && split_ctot == 0 && y &&
      static_cast<long long>(B) * D * H * W / 8 < (1ll << 31)) {
    const int rpi = kBnThreads / (C / V);
    const long long windows = static_cast<long long>(B) * D * H * W / 8;
    long long lb = (windows + rpi - 1) / rpi;
    if (lb > static_cast<long long>(sms) * 3) lb = static_cast<long long>(sms) * 3;
    if (lb < 1) lb = 1;
    const int g = static_cast<int>(lb);
    if (act == ICSG3D_ACT_NONE) launch_k(bn_apply_fwd_pool_lean_kernel<ICSG3D_ACT_NONE>, g, kBnThreads, 0, st, p);
    else if (act == ICSG3D_ACT_RELU) launch_k(bn_apply_fwd_pool_lean_kernel<ICSG3D_ACT_RELU>, g, kBnThreads, 0, st, p);
    else launch_k(bn_apply_fwd_pool_lean_kernel<ICSG3D_ACT_LEAKY>, g, kBnThreads, 0, st, p);
    ICSG_CHECK_LAUNCH();
    return ICSG3D_OK;
  }
  if (lean && x_dtype == ICSG3D_DT_BF16 && (post == ICSG3D_POST_NONE || post == ICSG3D_POST_UP2) && split_ctot == 0 && y && !y32 &&
      kBnThreads % (C / V) == 0 && (ldy & 7) == 0 && static_cast<long long>(B) * D * H * W < (1ll << 31) - (1ll << 24)) {
    const int rpi = kBnThreads / (C / V);
    const long long rows = static_cast<long long>(B) * D * H * W;
    long long lb = (rows + rpi - 1) / rpi;
    if (lb > static_cast<long long>(sms) * 4) lb = static_cast<long long>(sms) * 4;
    if (lb < 1) lb = 1;
    const int g = static_cast<int>(lb);
    const bool up = post == ICSG3D_POST_UP2;
#define ICSG_BN_LEAN(A)                                                                            \
  if (up) launch_k(bn_apply_fwd_lean_kernel<A, true>, g, kBnThreads, 0, st, p);                     \
  else launch_k(bn_apply_fwd_lean_kernel<A, false>, g, kBnThreads, 0, st, p)
    if (act == ICSG3D_ACT_NONE) { ICSG_BN_LEAN(ICSG3D_ACT_NONE); }
    else if (act == ICSG3D_ACT_RELU) { ICSG_BN_LEAN(ICSG3D_ACT_RELU); }
    else { ICSG_BN_LEAN(ICSG3D_ACT_LEAKY); }
#undef ICSG_BN_LEAN
    ICSG_CHECK_LAUNCH();
    return ICSG3D_OK;
  }
  if (x_dtype == ICSG3D_DT_BF16) launch_k(bn_apply_fwd_kernel<__nv_bfloat16>, static_cast<int>(blocks), kBnThreads, 0, st, p);
  else launch_k(bn_apply_fwd_kernel<float>, static_cast<int>(blocks), kBnThreads, 0, st, p);
  ICSG_CHECK_LAUNCH();
  return ICSG3D_OK;
}

extern "C" int icsg3d_bn_apply_fwd(const void* x, int ldx, int x_dtype, const float* scale, const float* shift, int act,
                                   float alpha, int post, int B, int D, int H, int W, int C, void* y, int ldy,
                                   float* y32, int ldy32, uint8_t* pool_idx, void* stream) {
  return bn_apply_fwd_impl(x, ldx, x_dtype, scale, shift, act, alpha, post, B, D, H, W, C, y, ldy, y32, ldy32, pool_idx, 0, 0, 0,
                           stream);
}

extern "C" int icsg3d_bn_apply_fwd_split3(const void* x, int ldx, int x_dtype, const float* scale, const float* shift, int act,
                                          float alpha, int post, int B, int D, int H, int W, int C, void* y, int ldy,
                                          uint8_t* pool_idx, int ctot, int coff, int fmt, void* stream) {
  ICSG_REQUIRE(ctot > 0 && (fmt == 0 || fmt == 1), "bn_apply_fwd_split3: bad ctot / fmt");
  return bn_apply_fwd_impl(x, ldx, x_dtype, scale, shift, act, alpha, post, B, D, H, W, C, y, ldy, nullptr, 0, pool_idx, ctot,
                           coff, fmt, stream);
}

// Backward grids are ONE wave of co-resident blocks (grid-stride loops): the lean bf16 kernels keep 3 blocks per SM in
// the reduce pass and 2 in the apply pass; the generic kernel (fp32, skip gradient, upsample) keeps the 4-per-SM cap.
static int g_bn_plain_u8 = -1;  // ICSG3D_BN_BWD_PLAIN_U8=1: plain reduce with 8 rows in flight, 2 blocks per SM
static bool bn_plain_u8() {
  if (g_bn_plain_u8 < 0) {
    const char* e = getenv("ICSG3D_BN_BWD_PLAIN_U8");
    g_bn_plain_u8 = e ? atoi(e) : 0;
  }
  return g_bn_plain_u8 != 0;
}
static int bn_bwd_grid(long long rows, int C, int dtype, int post, bool apply) {
  const int V = dtype == ICSG3D_DT_BF16 ? 8 : 4;
  if (dtype != ICSG3D_DT_BF16 || post == ICSG3D_POST_UP2) {
    const int g = bn_grid(rows, C, V);
    return apply && g * 2 < 148 * 8 ? g * 2 : g;
  }
  if (apply) return bn_grid(rows, C, V, 2);
  return bn_grid(rows, C, V, post == ICSG3D_POST_NONE && bn_plain_u8() ? 2 : 3);
}

// Launch the instantiation compiled for this post-op (each path gets its own register budget / occupancy).
static int g_bn_bwd_minb = -1;  // tuning knob: ICSG3D_BN_BWD_MINB = 2 | 3 | 4 resident blocks per SM
template <typename T, bool kApply, int kPost, int kMinB>
static void bn_bwd_launch_one(const BnBwdParams& p, int grid, size_t smem, cudaStream_t st) {
  auto* k = bn_bwd_kernel<T, kApply, kPost, kMinB>;
  if (smem > 48 * 1024) cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  launch_k(k, grid, kBnThreads, smem, st, p);
}
template <typename T, bool kApply, int kPost>
static void bn_bwd_dispatch_minb(const BnBwdParams& p, int grid, size_t smem, cudaStream_t st) {
  if (g_bn_bwd_minb < 0) {
    const char* e = getenv("ICSG3D_BN_BWD_MINB");
    g_bn_bwd_minb = e ? atoi(e) : 0;
  }
  const int def = kBnBwdMinBlocks<kApply, kPost>();
  const int mb = g_bn_bwd_minb >= 2 && g_bn_bwd_minb <= 4 ? g_bn_bwd_minb : def;
  if (mb == 2) bn_bwd_launch_one<T, kApply, kPost, 2>(p, grid, smem, st);
  else if (mb == 3) bn_bwd_launch_one<T, kApply, kPost, 3>(p, grid, smem, st);
  else bn_bwd_launch_one<T, kApply, kPost, 4>(p, grid, smem, st);
}
static bool bn_plain_u8();
static int g_bn_bwd_lean = -1;  // ICSG3D_BN_BWD_LEAN=0 keeps the generic kernel (A/B measurements)
template <bool kApply, int kPost, bool kHasAct, int kU = 4>
static void bn_bwd_lean_launch(const BnBwdParams& p, int grid, size_t smem, cudaStream_t st) {
  auto* k = bn_bwd_lean_kernel<kApply, kPost, kHasAct, kU>;
  if (smem > 48 * 1024) cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  launch_k(k, grid, kBnThreads, smem, st, p);
}
template <bool kApply>
static bool bn_bwd_try_lean(const BnBwdParams& p, int grid, size_t smem, cudaStream_t st) {
  if (g_bn_bwd_lean < 0) {
    const char* e = getenv("ICSG3D_BN_BWD_LEAN");
    g_bn_bwd_lean = e ? atoi(e) : 1;
  }
  if (!g_bn_bwd_lean || p.dy2 || p.post == ICSG3D_POST_UP2) return false;
  const bool act = p.act != ICSG3D_ACT_NONE;
  if (p.post == ICSG3D_POST_POOL2) {
    if (static_cast<long long>(p.B) * p.D * p.H * p.W / 8 >= (1ll << 31)) return false;  // 32-bit window arithmetic
    if (act) bn_bwd_lean_launch<kApply, ICSG3D_POST_POOL2, true>(p, grid, smem, st);
    else bn_bwd_lean_launch<kApply, ICSG3D_POST_POOL2, false>(p, grid, smem, st);
  } else if (!kApply && bn_plain_u8()) {
    if constexpr (!kApply) {
      if (act) bn_bwd_lean_launch<false, ICSG3D_POST_NONE, true, 8>(p, grid, smem, st);
      else bn_bwd_lean_launch<false, ICSG3D_POST_NONE, false, 8>(p, grid, smem, st);
    }
  } else {
    if (act) bn_bwd_lean_launch<kApply, ICSG3D_POST_NONE, true>(p, grid, smem, st);
    else bn_bwd_lean_launch<kApply, ICSG3D_POST_NONE, false>(p, grid, smem, st);
  }
  return true;
}

template <typename T, bool kApply>
static void bn_bwd_dispatch(const BnBwdParams& p, int grid, size_t smem, cudaStream_t st) {
  if constexpr (std::is_same<T, __nv_bfloat16>::value) {
    if (bn_bwd_try_lean<kApply>(p, grid, smem, st)) return;
  }
  if (p.post == ICSG3D_POST_POOL2) bn_bwd_dispatch_minb<T, kApply, ICSG3D_POST_POOL2>(p, grid, smem, st);
  else if (p.post == ICSG3D_POST_UP2) bn_bwd_dispatch_minb<T, kApply, ICSG3D_POST_UP2>(p, grid, smem, st);
  else bn_bwd_dispatch_minb<T, kApply, ICSG3D_POST_NONE>(p, grid, smem, st);
}

static int bn_bwd_launch(bool apply, const void* dy, int lddy, const void* dy2, int lddy2, const void* x, int ldx, int dtype, const float* mean,
                         const float* rstd, const float* scale, const float* shift, int act, float alpha, int post,
                         const uint8_t* pool_idx, int B, int D, int H, int W, int C, double* partials, int nparts,
                         const double* sums, double count, int pre_relu, const void* tap_other, int ld_other,
                         float tap_coef, void* dx, int lddx, void* stream, double* tap_sq = nullptr, int tap_sq_nparts = 0,
                         int dx_f32 = 0) {
  ICSG_REQUIRE(dy && x && mean && rstd && scale && shift, "bn_bwd: null pointer");
  ICSG_REQUIRE(bn_shape_ok(C, dtype), "bn_bwd: unsupported C=%d for dtype %d", C, dtype);
  const int V = dtype == ICSG3D_DT_BF16 ? 8 : 4;
  ICSG_REQUIRE(ldx % V == 0 && lddy % V == 0 && (!dy2 || lddy2 % V == 0), "bn_bwd: ld must be a multiple of %d", V);
  ICSG_REQUIRE(post != ICSG3D_POST_POOL2 || pool_idx, "bn_bwd: pool needs pool_idx");
  ICSG_REQUIRE(!dx_f32 || dtype == ICSG3D_DT_F32, "bn_bwd: an fp32 dx needs fp32 activations / gradients");
  long long rows = static_cast<long long>(B) * D * H * W;
  if (post == ICSG3D_POST_POOL2) rows /= 8;
  const int grid = bn_bwd_grid(rows, C, dtype, post, apply);
  BnBwdParams p{};
  p.dy = dy; p.lddy = lddy; p.dy2 = dy2; p.lddy2 = lddy2; p.x = x; p.ldx = ldx; p.mean = mean; p.rstd = rstd; p.scale = scale; p.shift = shift;
  p.act = act; p.alpha = alpha; p.post = post; p.pool_idx = pool_idx; p.B = B; p.D = D; p.H = H; p.W = W; p.C = C;
  p.partials = partials; p.sums = sums; p.count = count; p.pre_relu = pre_relu;
  p.tap_other = static_cast<const __nv_bfloat16*>(tap_other); p.ld_other = ld_other; p.tap_coef = tap_coef;
  p.dx = static_cast<__nv_bfloat16*>(dx); p.lddx = lddx; p.dx_f32 = dx_f32;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (!apply) {
    ICSG_REQUIRE(partials && nparts == grid, "bn_bwd_reduce: nparts %d != expected %d", nparts, grid);
    const size_t smem = bn_red_smem(C, V);
    if (dtype == ICSG3D_DT_BF16) bn_bwd_dispatch<__nv_bfloat16, false>(p, grid, smem, st);
    else bn_bwd_dispatch<float, false>(p, grid, smem, st);
  } else {
    ICSG_REQUIRE(sums && dx && count > 0, "bn_bwd_apply: bad arguments");
    ICSG_REQUIRE(lddx % 4 == 0, "bn_bwd_apply: lddx must be a multiple of 4");
    ICSG_REQUIRE(!tap_sq || (tap_other && tap_sq_nparts == grid), "bn_bwd_apply: tap_sq needs a tap and %d partials (got %d)",
                 grid, tap_sq_nparts);
    p.tap_sq = tap_sq;
    if (dtype == ICSG3D_DT_BF16) bn_bwd_dispatch<__nv_bfloat16, true>(p, grid, 0, st);
    else bn_bwd_dispatch<float, true>(p, grid, 0, st);
  }
  ICSG_CHECK_LAUNCH();
  return ICSG3D_OK;
}

extern "C" int icsg3d_bn_bwd_nparts(int B, int D, int H, int W, int C, int dtype, int post) {
  if (!bn_shape_ok(C, dtype)) return -1;
  long long rows = static_cast<long long>(B) * D * H * W;
  if (post == ICSG3D_POST_POOL2) rows /= 8;
  return bn_bwd_grid(rows, C, dtype, post, false);
}

extern "C" int icsg3d_bn_bwd_reduce(const void* dy, int lddy, const void* dy2, int lddy2, const void* x, int ldx, int dtype, const float* mean,
                                    const float* rstd, const float* scale, const float* shift, int act, float alpha,
                                    int post, const uint8_t* pool_idx, int B, int D, int H, int W, int C,
                                    double* partials, int nparts, void* stream) {
  return bn_bwd_launch(false, dy, lddy, dy2, lddy2, x, ldx, dtype, mean, rstd, scale, shift, act, alpha, post, pool_idx, B, D, H, W, C,
                       partials, nparts, nullptr, 0.0, 0, nullptr, 0, 0.f, nullptr, 0, stream);
}

extern "C" int icsg3d_bn_bwd_apply(const void* dy, int lddy, const void* dy2, int lddy2, const void* x, int ldx, int dtype, const float* mean,
                                   const float* rstd, const float* scale, const float* shift, int act, float alpha,
                                   int post, const uint8_t* pool_idx, int B, int D, int H, int W, int C,
                                   const double* sums, double count, int pre_relu, const void* tap_other, int ld_other,
                                   float tap_coef, void* dx, int lddx, void* stream) {
  return bn_bwd_launch(true, dy, lddy, dy2, lddy2, x, ldx, dtype, mean, rstd, scale, shift, act, alpha, post, pool_idx, B, D, H, W, C,
                       nullptr, 0, sums, count, pre_relu, tap_other, ld_other, tap_coef, dx, lddx, stream);
}

extern "C" int icsg3d_bn_bwd_apply_f32(const void* dy, int lddy, const void* dy2, int lddy2, const void* x, int ldx, const float* mean,
                                       const float* rstd, const float* scale, const float* shift, int act, float alpha,
                                       int post, const uint8_t* pool_idx, int B, int D, int H, int W, int C,
                                       const double* sums, double count, int pre_relu, const void* tap_other, int ld_other,
                                       float tap_coef, float* dx, int lddx, void* stream) {
  return bn_bwd_launch(true, dy, lddy, dy2, lddy2, x, ldx, ICSG3D_DT_F32, mean, rstd, scale, shift, act, alpha, post, pool_idx, B, D,
                       H, W, C, nullptr, 0, sums, count, pre_relu, tap_other, ld_other, tap_coef, dx, lddx, stream, nullptr, 0, 1);
}

extern "C" int icsg3d_bn_bwd_apply_nblocks(int B, int D, int H, int W, int C, int dtype, int post) {
  if (!bn_shape_ok(C, dtype)) return -1;
  long long rows = static_cast<long long>(B) * D * H * W;
  if (post == ICSG3D_POST_POOL2) rows /= 8;
  return bn_bwd_grid(rows, C, dtype, post, true);
}

extern "C" int icsg3d_bn_bwd_apply_tapsq(const void* dy, int lddy, const void* x, int ldx, int dtype, const float* mean,
                                         const float* rstd, const float* scale, const float* shift, int act, float alpha,
                                         int post, const uint8_t* pool_idx, int B, int D, int H, int W, int C,
                                         const double* sums, double count, int pre_relu, const void* tap_other,
                                         int ld_other, float tap_coef, void* dx, int lddx, double* tap_sq, int tap_sq_nparts,
                                         void* stream) {
  ICSG_REQUIRE(tap_sq && tap_other, "bn_bwd_apply_tapsq: tap_sq and tap_other are required");
  return bn_bwd_launch(true, dy, lddy, nullptr, 0, x, ldx, dtype, mean, rstd, scale, shift, act, alpha, post, pool_idx, B, D, H, W,
                       C, nullptr, 0, sums, count, pre_relu, tap_other, ld_other, tap_coef, dx, lddx, stream, tap_sq,
                       tap_sq_nparts);
}

extern "C" int icsg3d_bn_param_grads(const double* sums, float* dgamma, float* dbeta, int C, void* stream) {
  ICSG_REQUIRE(sums, "bn_param_grads: null pointer");
  launch_k(bn_param_grads_kernel, ceil_div(C, 128), 128, 0, static_cast<cudaStream_t>(stream), sums, dgamma, dbeta, C);
  ICSG_CHECK_LAUNCH();
  return ICSG3D_OK;
}

extern "C" int icsg3d_bn_reduce_finalize(const double* partials, int nparts, double count, const float* gamma,
                                         const float* beta, float eps, double* sums, float* mean, float* rstd, float* scale,
                                         float* shift, float* moving_mean, float* moving_var, float momentum, int C,
                                         void* stream) {
  ICSG_REQUIRE(partials && nparts > 0 && mean && rstd && scale && shift && count > 0, "bn_reduce_finalize: bad arguments");
  ICSG_REQUIRE((moving_mean == nullptr) == (moving_var == nullptr), "bn_reduce_finalize: moving_mean/var must both be given");
  launch_k(bn_reduce_fused_kernel, ceil_div(C, kRedCh), 1024, 0, static_cast<cudaStream_t>(stream), 
      partials, nparts, C, 0, count, gamma, beta, eps, sums, mean, rstd, scale, shift, moving_mean, moving_var, momentum, nullptr,
      nullptr);
  ICSG_CHECK_LAUNCH();
  return ICSG3D_OK;
}

extern "C" int icsg3d_bn_reduce_grads(const double* partials, int nparts, int C, double* sums, float* dgamma, float* dbeta,
                                      void* stream) {
  ICSG_REQUIRE(partials && nparts > 0 && sums, "bn_reduce_grads: bad arguments");
  launch_k(bn_reduce_fused_kernel, ceil_div(C, kRedCh), 1024, 0, static_cast<cudaStream_t>(stream), 
      partials, nparts, C, 1, 1.0, nullptr, nullptr, 0.f, sums, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, 0.f, dgamma,
      dbeta);
  ICSG_CHECK_LAUNCH();
  return ICSG3D_OK;
}


static int bn_peer_check(const uint64_t* peers, int world, int rank, int slot, int nslots, int cmax, const int64_t* epoch, int C) {
  ICSG_REQUIRE(peers && epoch, "bn all-reduce: null peer table / epoch");
  ICSG_REQUIRE(world >= 1 && world <= 16 && rank >= 0 && rank < world, "bn all-reduce: bad world/rank %d/%d", world, rank);
  ICSG_REQUIRE(slot >= 0 && slot < nslots && cmax >= C && ceil_div(C, kRedCh) <= 64, "bn all-reduce: bad slot/cmax (C=%d)", C);
  return ICSG3D_OK;
}

extern "C" int64_t icsg3d_bn_allreduce_buffer_bytes(int world, int nslots, int cmax) {
  if (world < 1 || nslots < 1 || cmax < 1) return -1;
  // classic regions (flags + data) followed by the flag-in-word region (two tagged 8-byte words per value)
  return static_cast<int64_t>(bn_peer_classic_bytes(world, nslots, cmax)) +
         static_cast<int64_t>(nslots) * 2 * world * (2 * static_cast<int64_t>(cmax)) * 16;
}

extern "C" int icsg3d_bn_reduce_allreduce_finalize(const double* partials, int nparts, double count_global, const float* gamma,
                                                   const float* beta, float eps, double* sums, float* mean, float* rstd,
                                                   float* scale, float* shift, float* moving_mean, float* moving_var,
                                                   float momentum, int C, const uint64_t* peers, int world, int rank, int slot,
                                                   int nslots, int cmax, const int64_t* epoch, void* stream) {
  ICSG_REQUIRE(partials && nparts > 0 && mean && rstd && scale && shift && count_global > 0, "bn_reduce_allreduce_finalize: bad arguments");
  int rc = bn_peer_check(peers, world, rank, slot, nslots, cmax, epoch, C);
  if (rc) return rc;
  BnPeerParams pp{reinterpret_cast<const unsigned long long*>(peers), world, rank, slot, nslots, cmax,
                  reinterpret_cast<const long long*>(epoch), peer_timeout_cycles(), peer_ll_enabled() ? 1 : 0};
  launch_k(bn_reduce_allreduce_kernel, ceil_div(C, kRedCh), 1024, 0, static_cast<cudaStream_t>(stream), 
      partials, nparts, C, 0, count_global, gamma, beta, eps, sums, mean, rstd, scale, shift, moving_mean, moving_var, momentum,
      nullptr, nullptr, pp);
  ICSG_CHECK_LAUNCH();
  return ICSG3D_OK;
}

extern "C" int icsg3d_bn_reduce_allreduce_grads(const double* partials, int nparts, int C, double* sums_global, float* dgamma,
                                                float* dbeta, const uint64_t* peers, int world, int rank, int slot, int nslots,
                                                int cmax, const int64_t* epoch, void* stream) {
  ICSG_REQUIRE(partials && nparts > 0 && sums_global, "bn_reduce_allreduce_grads: bad arguments");
  int rc = bn_peer_check(peers, world, rank, slot, nslots, cmax, epoch, C);
  if (rc) return rc;
  BnPeerParams pp{reinterpret_cast<const unsigned long long*>(peers), world, rank, slot, nslots, cmax,
                  reinterpret_cast<const long long*>(epoch), peer_timeout_cycles(), peer_ll_enabled() ? 1 : 0};
  launch_k(bn_reduce_allreduce_kernel, ceil_div(C, kRedCh), 1024, 0, static_cast<cudaStream_t>(stream), 
      partials, nparts, C, 1, 1.0, nullptr, nullptr, 0.f, sums_global, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, 0.f,
      dgamma, dbeta, pp);
  ICSG_CHECK_LAUNCH();
  return ICSG3D_OK;
}


// Whole BatchNorm backward in one cooperative launch (see bn_bwd_fused_kernel).  partials: scratch for at least
// icsg3d_bn_bwd_fused_nparts() x 2C doubles.  peers == NULL (world 1): single device.
extern "C" int icsg3d_bn_bwd_fused_nparts(int C, int dtype) {
  if (!bn_shape_ok(C, dtype)) return -1;
  int sms = sm_count();
  if (sms <= 0) sms = 148;
  return 2 * sms;  // upper bound of the co-resident grid (2 blocks of 256 threads per SM)
}

extern "C" int icsg3d_bn_bwd_fused(const void* dy, int lddy, const void* dy2, int lddy2, const void* x, int ldx, int dtype,
                                   const float* mean, const float* rstd, const float* scale, const float* shift, int act,
                                   float alpha, int post, const uint8_t* pool_idx, int B, int D, int H, int W, int C,
                                   double* partials, double* sums, double count_global, float* dgamma, float* dbeta, int pre_relu,
                                   const void* tap_other, int ld_other, float tap_coef, void* dx, int lddx,
                                   const uint64_t* peers, int world, int rank, int slot, int nslots, int cmax,
                                   const int64_t* epoch, void* stream) {
  ICSG_REQUIRE(dy && x && mean && rstd && scale && shift && partials && sums && dx && count_global > 0, "bn_bwd_fused: null pointer");
  ICSG_REQUIRE(bn_shape_ok(C, dtype), "bn_bwd_fused: unsupported C=%d for dtype %d", C, dtype);
  const int V = dtype == ICSG3D_DT_BF16 ? 8 : 4;
  ICSG_REQUIRE(ldx % V == 0 && lddy % V == 0 && (!dy2 || lddy2 % V == 0) && lddx % 4 == 0, "bn_bwd_fused: bad leading dimension");
  ICSG_REQUIRE(post != ICSG3D_POST_POOL2 || pool_idx, "bn_bwd_fused: pool needs pool_idx");
  ICSG_REQUIRE(!tap_other || dtype == ICSG3D_DT_BF16, "bn_bwd_fused: tap gradient needs bf16 activations");
  BnBwdParams p{};
  p.dy = dy; p.lddy = lddy; p.dy2 = dy2; p.lddy2 = lddy2; p.x = x; p.ldx = ldx; p.mean = mean; p.rstd = rstd; p.scale = scale; p.shift = shift;
  p.act = act; p.alpha = alpha; p.post = post; p.pool_idx = pool_idx; p.B = B; p.D = D; p.H = H; p.W = W; p.C = C;
  p.partials = partials; p.sums = sums; p.count = count_global; p.pre_relu = pre_relu;
  p.tap_other = static_cast<const __nv_bfloat16*>(tap_other); p.ld_other = ld_other; p.tap_coef = tap_coef;
  p.dx = static_cast<__nv_bfloat16*>(dx); p.lddx = lddx;
  BnFusedExtra e{};
  e.sums = sums; e.dgamma = dgamma; e.dbeta = dbeta;
  e.pp = BnPeerParams{reinterpret_cast<const unsigned long long*>(peers), peers ? world : 1, rank, slot, nslots, cmax,
                      reinterpret_cast<const long long*>(epoch), peer_timeout_cycles(), peer_ll_enabled() ? 1 : 0};
  if (peers) {
    int rc = bn_peer_check(peers, world, rank, slot, nslots, cmax, epoch, C);
    if (rc) return rc;
  }
  const size_t smem = bn_red_smem(C, V);
  const void* fn = dtype == ICSG3D_DT_BF16 ? reinterpret_cast<const void*>(bn_bwd_fused_kernel<__nv_bfloat16>)
                                           : reinterpret_cast<const void*>(bn_bwd_fused_kernel<float>);
  if (smem > 48 * 1024) ICSG_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
  int occ = 0;
  ICSG_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, fn, kBnThreads, smem));
  ICSG_REQUIRE(occ >= 1, "bn_bwd_fused: kernel does not fit an SM");
  if (occ > 2) occ = 2;
  int sms = sm_count();
  if (sms <= 0) sms = 148;
  long long rows = static_cast<long long>(B) * D * H * W;
  if (post == ICSG3D_POST_POOL2) rows /= 8;
  int grid = bn_grid(rows, C, V);
  if (grid > occ * sms) grid = occ * sms;
  const int nblk = ceil_div(C, kRedCh);
  if (grid < nblk) grid = nblk;  // phase B needs one block per 8 channels
  ICSG_REQUIRE(grid <= occ * sms, "bn_bwd_fused: C=%d needs more reduction blocks than can be co-resident", C);
  void* args[] = {&p, &e};
  ICSG_CUDA(cudaLaunchCooperativeKernel(fn, dim3(grid), dim3(kBnThreads), args, smem, static_cast<cudaStream_t>(stream)));
  ++g_launches;
  return ICSG3D_OK;
}
