// Device-side post-/pre-processing around the two networks (SURVEY §8f.1, §8f.3 and the metric closures of §8b):
//
//   * to_lattice_params / to_voxel_params (utils.py:160-190): per-sample min/max over the three coordinate channels
//     of the decoder output, fused with the fp32 -> bf16 packing of the U-Net input (the decoder output is read ONCE on its
//     way into the segmentation network; generate.py:208-220), then the reference's float32 arithmetic incl. its
//     a*(1-1/d) quirk, op by op with round-to-nearest intrinsics (no FMA contraction, no fast-math division);
//   * argmax species label + sigmoid >= threshold atom mask (generate.py:221-225) straight from the fp32 head logits;
//   * random_rotation_3d (utils.py:193-222): the reference rotates by exactly 90 degrees with a cubic spline, which
//     reproduces a signed axis permutation up to spline round-off (5e-16); the kernel applies the composed signed
//     permutation exactly, per sample, to any voxel payload (density fp32/fp64, species u8/fp64, 3 coordinate channels);
//   * r_m / p_m / f1_m / wr_m (unet.py:159-193): the K.round(K.clip(.,0,1)) counts over (y_true, y_pred) tensors.
//
// All of it is HBM-bound streaming work: 16-byte accesses, grids sized from the SM count, fp64/integer-exact sums.
#include "common.cuh"

namespace icsg3d {

// ------------------------------------------------------------------------------------------------
// min / max of the coordinate channels (+ optional bf16 pack of the 4-channel network input)
// ------------------------------------------------------------------------------------------------
static constexpr int kMmThreads = 256;

// p: T [B][vox][ld], channels c0..c0+2 are (x, y, z).  partials: T [B][nsplit][6] = (min x,y,z, max x,y,z).
// x16 (optional, T = float, ld == 4, c0 == 1 only): bf16 [B][vox][16] = (channels 0..3, zeros) — the U-Net input.
template <typename T>
__global__ void __launch_bounds__(kMmThreads) coord_minmax_kernel(const T* __restrict__ p, int ld, int c0, long long vox,
                                                                  int nsplit, T* __restrict__ partials,
                                                                  __nv_bfloat16* __restrict__ x16) {
  pdl_prologue();
  const int b = blockIdx.y, s = blockIdx.x;
  const long long per = (vox + nsplit - 1) / nsplit;
  const long long v0 = s * per, v1 = (v0 + per < vox) ? v0 + per : vox;
  const T* base = p + static_cast<long long>(b) * vox * ld;
  T mn[3], mx[3];
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    mn[c] = static_cast<T>(INFINITY);
    mx[c] = static_cast<T>(-INFINITY);
  }
  for (long long v = v0 + threadIdx.x; v < v1; v += kMmThreads) {
    T c[3];
    if constexpr (sizeof(T) == 4) {
      if (ld == 4 && c0 == 1) {
        const float4 q = *reinterpret_cast<const float4*>(base + v * 4);
        c[0] = q.y; c[1] = q.z; c[2] = q.w;
        if (x16) {
          uint4 lo = make_uint4(pack_bf16x2(q.x, q.y), pack_bf16x2(q.z, q.w), 0u, 0u);
          uint4* dst = reinterpret_cast<uint4*>(x16 + (static_cast<long long>(b) * vox + v) * 16);
          dst[0] = lo;
          dst[1] = make_uint4(0u, 0u, 0u, 0u);
        }
      } else {
#pragma unroll
        for (int k = 0; k < 3; ++k) c[k] = base[v * ld + c0 + k];
      }
    } else {
#pragma unroll
      for (int k = 0; k < 3; ++k) c[k] = base[v * ld + c0 + k];
    }
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      mn[k] = c[k] < mn[k] ? c[k] : mn[k];
      mx[k] = c[k] > mx[k] ? c[k] : mx[k];
    }
  }
  __shared__ T red[kMmThreads / 32][6];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < 3; ++k) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const T a = __shfl_xor_sync(0xffffffffu, mn[k], o), bb = __shfl_xor_sync(0xffffffffu, mx[k], o);
      mn[k] = a < mn[k] ? a : mn[k];
      mx[k] = bb > mx[k] ? bb : mx[k];
    }
    if (lane == 0) {
      red[w][k] = mn[k];
      red[w][3 + k] = mx[k];
    }
  }
  __syncthreads();
  if (threadIdx.x < 6) {
    T r = red[0][threadIdx.x];
    for (int i = 1; i < kMmThreads / 32; ++i) {
      const T o = red[i][threadIdx.x];
      r = threadIdx.x < 3 ? (o < r ? o : r) : (o > r ? o : r);
    }
    partials[(static_cast<long long>(b) * nsplit + s) * 6 + threadIdx.x] = r;
  }
}

__device__ __forceinline__ float div_rn(float a, float b) { return __fdiv_rn(a, b); }
__device__ __forceinline__ double div_rn(double a, double b) { return __ddiv_rn(a, b); }
__device__ __forceinline__ float sub_rn(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ double sub_rn(double a, double b) { return __dsub_rn(a, b); }
__device__ __forceinline__ float add_rn(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ double add_rn(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ float mul_rn(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ double mul_rn(double a, double b) { return __dmul_rn(a, b); }

// utils.py:160-190 in the array dtype T (numpy keeps float32 arrays float32 against python scalars):
//   ap = (max - min) / (1 + 2 eps) / (1 - 1/d);  ap -= ap / d;      dv = (lp + 2 lp eps) / d
template <typename T>
__global__ void lattice_finalize_kernel(const T* __restrict__ partials, int B, int nsplit, double eps_frac, int d,
                                        T* __restrict__ lp, T* __restrict__ dv) {
  pdl_prologue();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * 3) return;
  const int b = i / 3, k = i % 3;
  const T* pb = partials + static_cast<long long>(b) * nsplit * 6;
  T mn = pb[k], mx = pb[3 + k];
  for (int s = 1; s < nsplit; ++s) {
    const T a = pb[s * 6 + k], c = pb[s * 6 + 3 + k];
    mn = a < mn ? a : mn;
    mx = c > mx ? c : mx;
  }
  const T c1 = static_cast<T>(1.0 + 2.0 * eps_frac), c2 = static_cast<T>(1.0 - 1.0 / static_cast<double>(d));
  const T dd = static_cast<T>(d), e = static_cast<T>(eps_frac);
  T ap = div_rn(div_rn(sub_rn(mx, mn), c1), c2);
  ap = sub_rn(ap, div_rn(ap, dd));
  lp[i] = ap;
  if (dv) dv[i] = div_rn(add_rn(ap, mul_rn(mul_rn(static_cast<T>(2.0), ap), e)), dd);
}

// ------------------------------------------------------------------------------------------------
// generate.py:221-225: species = argmax_c of the float32 softmax output (first index on ties), mask = sigmoid >= thr
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) heads_predict_kernel(const float* __restrict__ logits, int ld, int c1, long long M,
                                                            float threshold, uint8_t* __restrict__ argmax_out,
                                                            uint8_t* __restrict__ mask_out, float* __restrict__ sig_prob) {
  pdl_prologue();
  const int lane = threadIdx.x & 31;
  const long long warp = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = (static_cast<long long>(gridDim.x) * blockDim.x) >> 5;
  for (long long v = warp; v < M; v += nwarps) {
    const float* row = logits + v * ld;
    // The reference takes np.argmax over the float32 SOFTMAX OUTPUT (generate.py:221): two classes whose probabilities
    // round to the same float tie and the first index wins.  Reproduce exactly that: the same p_j = exp(x_j - max) / sum
    // that heads_loss_kernel returns as `probs`, then arg-max over p with the first-index rule.
    float x[3], mx = -INFINITY;
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      const int col = lane + 32 * j;
      x[j] = col < c1 ? row[col] : -INFINITY;
      mx = fmaxf(mx, x[j]);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    float e[3], se = 0.f;
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      e[j] = lane + 32 * j < c1 ? expf(x[j] - mx) : 0.f;
      se += e[j];
    }
    se = warp_sum(se);
    const float inv = 1.f / se;
    float pmax = -1.f;
    int amax = 0;
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      const int col = lane + 32 * j;
      const float pj = col < c1 ? e[j] * inv : -1.f;
      if (pj > pmax) {
        pmax = pj;
        amax = col;
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float om = __shfl_xor_sync(0xffffffffu, pmax, o);
      const int oa = __shfl_xor_sync(0xffffffffu, amax, o);
      if (om > pmax || (om == pmax && oa < amax)) {
        pmax = om;
        amax = oa;
      }
    }
    if (lane == 0) {
      const float sp = 1.f / (1.f + expf(-row[c1]));
      if (argmax_out) argmax_out[v] = static_cast<uint8_t>(amax);
      if (mask_out) mask_out[v] = sp >= threshold ? 1 : 0;
      if (sig_prob) sig_prob[v] = sp;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// signed axis permutation of a batch of d^3 grids (random_rotation_3d with rot_angle = 90)
//   out[b][o0][o1][o2][:] = in[b][s0][s1][s2][:],  s_x = flip_x ? d-1-o[perm_x] : o[perm_x]
//   xf: int32 [B][6] = (perm0, perm1, perm2, flip0, flip1, flip2)
// One thread per output voxel and 4-/8-/16-byte word: writes are fully coalesced; reads are gathers whose
// footprint (one sample, <= 512 KB at 32^3 x 16 B) stays in L2.
// ------------------------------------------------------------------------------------------------
template <typename W>
__global__ void __launch_bounds__(256) rotate90_kernel(const W* __restrict__ in, W* __restrict__ out, int d, int wpv,
                                                       const int* __restrict__ xf, long long total) {
  pdl_prologue();
  const long long d3 = static_cast<long long>(d) * d * d;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long vox = i / wpv;
    const int wi = static_cast<int>(i - vox * wpv);
    const int b = static_cast<int>(vox / d3);
    const int r = static_cast<int>(vox - b * d3);
    int o[3];
    o[0] = r / (d * d);
    o[1] = (r / d) % d;
    o[2] = r % d;
    const int* t = xf + b * 6;
    int s[3];
#pragma unroll
    for (int x = 0; x < 3; ++x) {
      const int ox = t[x] == 0 ? o[0] : (t[x] == 1 ? o[1] : o[2]);
      s[x] = t[3 + x] ? d - 1 - ox : ox;
    }
    const long long src = (static_cast<long long>(b) * d3 + (static_cast<long long>(s[0]) * d + s[1]) * d + s[2]) * wpv + wi;
    out[i] = in[src];
  }
}

// ------------------------------------------------------------------------------------------------
// unet.py:159-193 metric counts over y_true / y_pred fp32 [rows][C]:
//   counts = [ sum round(clip(yt*yp)), sum round(clip(yt)), sum round(clip(yp)),
//              sum_{c>0} round(clip(yt*yp)), sum_{c>0} round(clip(yt)) ]          (K.round = half-to-even)
// Integer-valued, accumulated per thread as integers and added with fp64 atomics (exact => order independent).
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ int rclip(float v) { return static_cast<int>(rintf(fminf(fmaxf(v, 0.f), 1.f))); }

__global__ void __launch_bounds__(256) metric_counts_kernel(const float* __restrict__ yt, const float* __restrict__ yp,
                                                            long long n, int C, double* __restrict__ counts) {
  pdl_prologue();
  long long acc[5] = {0, 0, 0, 0, 0};
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const float t = yt[i], p = yp[i];
    const int tp = rclip(t * p), pos = rclip(t), pred = rclip(p);
    const bool nz = (i % C) != 0;
    acc[0] += tp;
    acc[1] += pos;
    acc[2] += pred;
    acc[3] += nz ? tp : 0;
    acc[4] += nz ? pos : 0;
  }
  __shared__ long long red[8][5];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < 5; ++k) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc[k] += __shfl_xor_sync(0xffffffffu, acc[k], o);
    if (lane == 0) red[w][k] = acc[k];
  }
  __syncthreads();
  if (threadIdx.x < 5) {
    long long tsum = 0;
    for (int i = 0; i < 8; ++i) tsum += red[i][threadIdx.x];
    if (tsum) atomicAdd(counts + threadIdx.x, static_cast<double>(tsum));
  }
}

static int grid_for(long long items, int threads, int per_sm) {
  int sms = sm_count();
  if (sms <= 0) sms = 148;
  long long b = (items + threads - 1) / threads;
  const long long cap = static_cast<long long>(sms) * per_sm;
  if (b > cap) b = cap;
  return static_cast<int>(b < 1 ? 1 : b);
}

}  // namespace icsg3d

using namespace icsg3d;
#define ST static_cast<cudaStream_t>(stream)

extern "C" int icsg3d_lattice_nsplit(int B, int64_t vox) {
  int sms = sm_count();
  if (sms <= 0) sms = 148;
  int n = (2 * sms + B - 1) / (B > 0 ? B : 1);
  const long long maxn = (vox + 1023) / 1024;  // at least 1024 voxels per block
  if (n > maxn) n = static_cast<int>(maxn);
  return n < 1 ? 1 : n;
}

extern "C" int icsg3d_coord_minmax(const void* p, int dtype, int ld, int c0, int B, int64_t vox, int nsplit, void* partials,
                                   void* x16, void* stream) {
  ICSG_REQUIRE(p && partials && B > 0 && vox > 0 && nsplit > 0, "coord_minmax: bad arguments");
  ICSG_REQUIRE(ld >= c0 + 3 && c0 >= 0, "coord_minmax: needs three coordinate channels at [c0, c0+3) of ld");
  ICSG_REQUIRE(dtype == ICSG3D_DT_F32 || dtype == ICSG3D_DT_F64, "coord_minmax: dtype must be fp32 or fp64");
  ICSG_REQUIRE(!x16 || (dtype == ICSG3D_DT_F32 && ld == 4 && c0 == 1), "coord_minmax: the fused pack needs fp32 [.,4] input");
  dim3 grid(nsplit, B);
  if (dtype == ICSG3D_DT_F32)
    launch_k(coord_minmax_kernel<float>, grid, kMmThreads, 0, ST, static_cast<const float*>(p), ld, c0, vox, nsplit,
                                                            static_cast<float*>(partials), static_cast<__nv_bfloat16*>(x16));
  else
    launch_k(coord_minmax_kernel<double>, grid, kMmThreads, 0, ST, static_cast<const double*>(p), ld, c0, vox, nsplit,
                                                             static_cast<double*>(partials), nullptr);
  ICSG_CHECK_LAUNCH();
  return ICSG3D_OK;
}

extern "C" int icsg3d_lattice_finalize(const void* partials, int dtype, int B, int nsplit, double eps_frac, int d, void* lp,
                                       void* dv, void* stream) {
  ICSG_REQUIRE(partials && lp && B > 0 && nsplit > 0 && d > 1, "lattice_finalize: bad arguments");
  ICSG_REQUIRE(dtype == ICSG3D_DT_F32 || dtype == ICSG3D_DT_F64, "lattice_finalize: dtype must be fp32 or fp64");
  const int n = B * 3;
  if (dtype == ICSG3D_DT_F32)
    launch_k(lattice_finalize_kernel<float>, ceil_div(n, 128), 128, 0, ST, static_cast<const float*>(partials), B, nsplit, eps_frac,
                                                                     d, static_cast<float*>(lp), static_cast<float*>(dv));
  else
    launch_k(lattice_finalize_kernel<double>, ceil_div(n, 128), 128, 0, ST, static_cast<const double*>(partials), B, nsplit,
                                                                      eps_frac, d, static_cast<double*>(lp),
                                                                      static_cast<double*>(dv));
  ICSG_CHECK_LAUNCH();
  return ICSG3D_OK;
}

extern "C" int icsg3d_heads_predict(const float* logits, int ld, int c1, int64_t M, float threshold, uint8_t* argmax_out,
                                    uint8_t* mask_out, float* sig_prob, void* stream) {
  ICSG_REQUIRE(logits && M > 0 && (argmax_out || mask_out || sig_prob), "heads_predict: bad arguments");
  ICSG_REQUIRE(c1 >= 1 && c1 <= 95 && ld > c1, "heads_predict: c1 must be in [1,95] and ld > c1");
  launch_k(heads_predict_kernel, grid_for(M * 32, 256, 16), 256, 0, ST, logits, ld, c1, M, threshold, argmax_out, mask_out, sig_prob);
  ICSG_CHECK_LAUNCH();
  return ICSG3D_OK;
}

extern "C" int icsg3d_rotate90_batch(const void* in, void* out, int B, int d, int voxel_bytes, const int* xforms,
                                     void* stream) {
  ICSG_REQUIRE(in && out && xforms && in != out && B > 0 && d > 0 && voxel_bytes > 0, "rotate90_batch: bad arguments");
  const long long vox = static_cast<long long>(B) * d * d * d;
  const uintptr_t al = reinterpret_cast<uintptr_t>(in) | reinterpret_cast<uintptr_t>(out);
  if (voxel_bytes % 16 == 0 && al % 16 == 0) {
    const int wpv = voxel_bytes / 16;
    launch_k(rotate90_kernel<uint4>, grid_for(vox * wpv, 256, 16), 256, 0, ST, static_cast<const uint4*>(in),
                                                                         static_cast<uint4*>(out), d, wpv, xforms, vox * wpv);
  } else if (voxel_bytes % 8 == 0 && al % 8 == 0) {
    const int wpv = voxel_bytes / 8;
    launch_k(rotate90_kernel<uint2>, grid_for(vox * wpv, 256, 16), 256, 0, ST, static_cast<const uint2*>(in),
                                                                         static_cast<uint2*>(out), d, wpv, xforms, vox * wpv);
  } else if (voxel_bytes % 4 == 0 && al % 4 == 0) {
    const int wpv = voxel_bytes / 4;
    launch_k(rotate90_kernel<uint32_t>, grid_for(vox * wpv, 256, 16), 256, 0, ST, 
        static_cast<const uint32_t*>(in), static_cast<uint32_t*>(out), d, wpv, xforms, vox * wpv);
  } else {
    launch_k(rotate90_kernel<uint8_t>, grid_for(vox * voxel_bytes, 256, 16), 256, 0, ST, 
        static_cast<const uint8_t*>(in), static_cast<uint8_t*>(out), d, voxel_bytes, xforms, vox * voxel_bytes);
  }
  ICSG_CHECK_LAUNCH();
  return ICSG3D_OK;
}

extern "C" int icsg3d_metric_counts(const float* y_true, const float* y_pred, int64_t n, int C, double* counts, void* stream) {
  ICSG_REQUIRE(y_true && y_pred && counts && n > 0 && C > 0, "metric_counts: bad arguments");
  ICSG_CUDA(cudaMemsetAsync(counts, 0, 5 * sizeof(double), ST));
  launch_k(metric_counts_kernel, grid_for(n, 256, 8), 256, 0, ST, y_true, y_pred, n, C, counts);
  ICSG_CHECK_LAUNCH();
  return ICSG3D_OK;
}
