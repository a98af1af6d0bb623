// Small fused kernels around the conv stacks: input packing, the VAE bottleneck (Dense / reparameterisation /
// KL), loss reductions (MSE, DFC feature MSE), loss-gradient seeds, and the Keras-form Adam update.
//
// Reference: vae/lattice_vae.py:53-66 (sampling), :182-185, :207-209 (Dense layers), :232-270 (losses),
// keras.optimizers.Adam (lattice_vae.py:98).
#include "common.cuh"

namespace icsg3d {

static int grid1d(long long n, int threads = 256, int cap_mult = 8) {
  long long b = (n + threads - 1) / threads;
  int sms = sm_count();
  if (sms <= 0) sms = 148;
  if (b > static_cast<long long>(sms) * cap_mult) b = static_cast<long long>(sms) * cap_mult;
  if (b < 1) b = 1;
  return static_cast<int>(b);
}

// ---- input packing ------------------------------------------------------------------------------
// m: fp32 [B*vox][4]; cond: fp32 [B][ncond] -> xe: bf16 [B*vox][16] = (M, cond one-hot, 0..) for the encoder
// (the 4x tiled condition is folded into the weights, see pack.cu), xp: bf16 [B*vox][16] = (M, 0..) for the
// perceptual U-Net.  Either output may be NULL.
__global__ void pack_vae_input_kernel(const float* __restrict__ m, const float* __restrict__ cond, int ncond,
                                      long long vox, long long total, __nv_bfloat16* __restrict__ xe,
                                      __nv_bfloat16* __restrict__ xp) {
  pdl_prologue();
  for (long long r = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; r < total;
       r += static_cast<long long>(gridDim.x) * blockDim.x) {
    const float4 v = *reinterpret_cast<const float4*>(m + r * 4);
    uint4 lo, hi;
    lo.x = pack_bf16x2(v.x, v.y);
    lo.y = pack_bf16x2(v.z, v.w);
    if (xp) {
      lo.z = lo.w = 0u;
      hi = make_uint4(0u, 0u, 0u, 0u);
      reinterpret_cast<uint4*>(xp + r * 16)[0] = lo;
      reinterpret_cast<uint4*>(xp + r * 16)[1] = hi;
    }
    if (xe) {
      const float* c = cond + (r / vox) * ncond;
      float cv[12];
#pragma unroll
      for (int i = 0; i < 12; ++i) cv[i] = i < ncond ? c[i] : 0.f;
      lo.z = pack_bf16x2(cv[0], cv[1]);
      lo.w = pack_bf16x2(cv[2], cv[3]);
      hi.x = pack_bf16x2(cv[4], cv[5]);
      hi.y = pack_bf16x2(cv[6], cv[7]);
      hi.z = pack_bf16x2(cv[8], cv[9]);
      hi.w = pack_bf16x2(cv[10], cv[11]);
      reinterpret_cast<uint4*>(xe + r * 16)[0] = lo;
      reinterpret_cast<uint4*>(xe + r * 16)[1] = hi;
    }
  }
}

// fp32 rows of `c` values -> first c channels of bf16 rows with stride ld (rest untouched), and back.
__global__ void f32_to_bf16_rows_kernel(const float* __restrict__ src, int c, long long rows,
                                        __nv_bfloat16* __restrict__ dst, int ld) {
  pdl_prologue();
  const long long total = rows * c;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x)
    dst[(i / c) * ld + (i % c)] = f2bf(src[i]);
}
__global__ void bf16_rows_to_f32_kernel(const __nv_bfloat16* __restrict__ src, int ld, int c, long long rows,
                                        float* __restrict__ dst) {
  pdl_prologue();
  const long long total = rows * c;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x)
    dst[i] = bf2f(src[(i / c) * ld + (i % c)]);
}

// ---- Dense ----------------------------------------------------------------------------------------
// y[b,n] = act( sum_k x1[b,k] W[k,n] + sum_k x2[b,k] W[K1+k,n] + bias[n] )   (Concatenate + Dense)
// Block = 32 outputs (n) x 8 k-slices; the 8 slice partials are added in slice order (deterministic).
__global__ void __launch_bounds__(256) dense_fwd_kernel(const float* __restrict__ x1, int k1, const float* __restrict__ x2,
                                                        int k2, const float* __restrict__ w, const float* __restrict__ bias,
                                                        int act, int B, int N, float* __restrict__ y) {
  pdl_prologue();
  __shared__ float red[8][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int n = blockIdx.x * 32 + tx, b = blockIdx.y;
  float acc = 0.f;
  if (n < N) {
    for (int k = ty; k < k1; k += 8) acc = fmaf(x1[b * k1 + k], w[static_cast<size_t>(k) * N + n], acc);
    for (int k = ty; k < k2; k += 8) acc = fmaf(x2[b * k2 + k], w[static_cast<size_t>(k1 + k) * N + n], acc);
  }
  red[ty][tx] = acc;
  __syncthreads();
  if (ty == 0 && n < N) {
    float t = bias ? bias[n] : 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) t += red[i][tx];
    if (act == ICSG3D_ACT_RELU) t = fmaxf(t, 0.f);
    y[b * N + n] = t;
  }
}
// dy <- dy * relu'(y) in place when act == RELU (y = saved forward output); then
// dx[b,k] = sum_n dy[b,n] W[k,n] for k < kx (the first kx rows of W), dW[k,n] = sum_b x[b,k] dy[b,n], db[n] = sum_b dy[b,n]
__global__ void dense_bwd_mask_kernel(float* __restrict__ dy, const float* __restrict__ y, int n) {
  pdl_prologue();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n && !(y[i] > 0.f)) dy[i] = 0.f;
}
// one warp per (b,k): lanes stride over n (coalesced W row), fixed-order shuffle tree
__global__ void __launch_bounds__(256) dense_bwd_input_kernel(const float* __restrict__ dy, const float* __restrict__ w, int B,
                                                              int N, int kx, float* __restrict__ dx, int accumulate) {
  pdl_prologue();
  const int lane = threadIdx.x & 31;
  const int idx = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (idx >= B * kx) return;
  const int b = idx / kx, k = idx % kx;
  float acc = 0.f;
  for (int n = lane; n < N; n += 32) acc = fmaf(dy[b * N + n], w[static_cast<size_t>(k) * N + n], acc);
  acc = warp_sum(acc);
  if (lane == 0) dx[idx] = accumulate ? dx[idx] + acc : acc;
}
__global__ void dense_bwd_weight_kernel(const float* __restrict__ x1, int k1, const float* __restrict__ x2, int k2,
                                        const float* __restrict__ dy, int B, int N, float* __restrict__ dw,
                                        float* __restrict__ db) {
  pdl_prologue();
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  const int K = k1 + k2;
  if (idx >= (K + 1) * N) return;
  const int k = idx / N, n = idx % N;
  float acc = 0.f;
  if (k == K) {
    for (int b = 0; b < B; ++b) acc += dy[b * N + n];
    db[n] = acc;
  } else {
    for (int b = 0; b < B; ++b) {
      const float xv = k < k1 ? x1[b * k1 + k] : x2[b * k2 + (k - k1)];
      acc = fmaf(xv, dy[b * N + n], acc);
    }
    dw[static_cast<size_t>(k) * N + n] = acc;
  }
}

// ---- reparameterisation + KL ----------------------------------------------------------------------
// z = mu + exp(0.5*lv)*eps ; kl[b] = -0.5 * sum_j (1 + lv - mu^2 - exp(lv))   (lattice_vae.py:53-66, 235-239)
__global__ void reparam_fwd_kernel(const float* __restrict__ mu, const float* __restrict__ lv,
                                   const float* __restrict__ eps, int L, float* __restrict__ z, float* __restrict__ kl) {
  pdl_prologue();
  const int b = blockIdx.x;
  float acc = 0.f;
  for (int j = threadIdx.x; j < L; j += blockDim.x) {
    const float m = mu[b * L + j], v = lv[b * L + j];
    z[b * L + j] = m + expf(0.5f * v) * eps[b * L + j];
    acc += 1.f + v - m * m - expf(v);
  }
  __shared__ float red[32];
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x < 32) {
    float t = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.f;
    t = warp_sum(t);
    if (threadIdx.x == 0) kl[b] = -0.5f * t;
  }
}
// dmu = dz + kl_coef*mu ; dlv = dz*0.5*exp(0.5 lv)*eps + 0.5*kl_coef*(exp(lv) - 1),  kl_coef = beta / B_global
__global__ void reparam_bwd_kernel(const float* __restrict__ dz, const float* __restrict__ mu,
                                   const float* __restrict__ lv, const float* __restrict__ eps, float kl_coef, int n,
                                   float* __restrict__ dmu, float* __restrict__ dlv) {
  pdl_prologue();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float v = lv[i];
  dmu[i] = dz[i] + kl_coef * mu[i];
  dlv[i] = dz[i] * 0.5f * expf(0.5f * v) * eps[i] + 0.5f * kl_coef * (expf(v) - 1.f);
}

// d(pre-activation) of a LeakyReLU output y given d(y), written as bf16 rows (stride ld, first c channels).
__global__ void leaky_bwd_rows_kernel(const float* __restrict__ dyv, const float* __restrict__ y, float alpha, int c,
                                      long long rows, __nv_bfloat16* __restrict__ dst, int ld) {
  pdl_prologue();
  const long long total = rows * c;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x)
    dst[(i / c) * ld + (i % c)] = f2bf(dyv[i] * (y[i] > 0.f ? 1.f : alpha));
}

// ---- loss reductions ------------------------------------------------------------------------------
// partials[blk] = sum (a-b)^2 over this block's slice (fp64); T = bf16 (n % 8 == 0) or fp32 (n % 4 == 0)
template <typename T>
__global__ void __launch_bounds__(256) sqdiff_partials_kernel(const T* __restrict__ a, const T* __restrict__ b,
                                                              long long n, double* __restrict__ partials) {
  pdl_prologue();
  constexpr int V = sizeof(T) == 2 ? 8 : 4;
  const long long nv = n / V;
  double dacc = 0.0;
  float acc = 0.f;
  int cnt = 0;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < nv;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    if constexpr (V == 8) {
      const uint4 qa = reinterpret_cast<const uint4*>(a)[i];
      const uint4 qb = reinterpret_cast<const uint4*>(b)[i];
      const uint32_t ua[4] = {qa.x, qa.y, qa.z, qa.w};
      const uint32_t ub[4] = {qb.x, qb.y, qb.z, qb.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 fa = unpack_bf16x2(ua[j]), fb = unpack_bf16x2(ub[j]);
        const float d0 = fa.x - fb.x, d1 = fa.y - fb.y;
        acc += d0 * d0 + d1 * d1;
      }
    } else {
      const float4 fa = reinterpret_cast<const float4*>(a)[i];
      const float4 fb = reinterpret_cast<const float4*>(b)[i];
      const float d0 = fa.x - fb.x, d1 = fa.y - fb.y, d2 = fa.z - fb.z, d3 = fa.w - fb.w;
      acc += d0 * d0 + d1 * d1 + d2 * d2 + d3 * d3;
    }
    if (++cnt == 16) {
      dacc += acc;
      acc = 0.f;
      cnt = 0;
    }
  }
  dacc += acc;
  dacc = warp_sum(dacc);
  __shared__ double red[8];
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = dacc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < 8; ++w) t += red[w];
    partials[blockIdx.x] = t;
  }
}

// VAE+DFC loss assembly (lattice_vae.py:241-255): one block.
//  sums layout: nterms segments of `stride` doubles each (only the first nparts[i] are valid);
//  term 0 = MSE sum, terms 1.. = DFC tap sums.  scales[i] multiplies term i's sum (1/numel etc.).
//  out[0..3] = [loss, pm, mse, kld] (batch means);  raw[0..5] (double) = [mse_sum, tap sums..., kl_sum] for DP.
__global__ void vae_loss_assemble_kernel(const double* __restrict__ partials, const int* __restrict__ nparts, int stride,
                                         int nterms, const double* __restrict__ scales, const float* __restrict__ kl,
                                         int B, double kl_scale, float alpha, float beta, float* __restrict__ out,
                                         double* __restrict__ raw) {
  pdl_prologue();
  __shared__ double term[8];
  __shared__ double klsum;
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (w < nterms) {
    double a = 0.0;
    for (int i = lane; i < nparts[w]; i += 32) a += partials[static_cast<size_t>(w) * stride + i];
    // fixed-order tree inside the warp: deterministic
    a = warp_sum(a);
    if (lane == 0) term[w] = a;
  }
  if (w == 7) {
    double a = 0.0;
    for (int i = lane; i < B; i += 32) a += kl[i];
    a = warp_sum(a);
    if (lane == 0) klsum = a;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    const double mse = term[0] * scales[0];
    double pm = 0.0;
    for (int i = 1; i < nterms; ++i) pm += term[i] * scales[i];
    const double kld = klsum * kl_scale;
    out[0] = static_cast<float>(mse + alpha * pm + beta * kld);
    out[1] = static_cast<float>(pm);
    out[2] = static_cast<float>(mse);
    out[3] = static_cast<float>(kld);
    if (raw) {
      for (int i = 0; i < nterms; ++i) raw[i] = term[i];
      raw[nterms] = klsum;
    }
  }
}

// ---- loss-gradient seeds ----------------------------------------------------------------------------
// d(loss)/d(x_hat) = mse_coef * (x_hat - x) + dgrad_c1[:, :4]      (fp32 [rows][4]; dgrad bf16 [rows][ld])
__global__ void xhat_grad_kernel(const float* __restrict__ x, const float* __restrict__ xhat, float mse_coef,
                                 const __nv_bfloat16* __restrict__ dpm, int ld, long long rows, float* __restrict__ dy) {
  pdl_prologue();
  for (long long r = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; r < rows;
       r += static_cast<long long>(gridDim.x) * blockDim.x) {
    const float4 a = reinterpret_cast<const float4*>(x)[r];
    const float4 b = reinterpret_cast<const float4*>(xhat)[r];
    float4 o = make_float4(mse_coef * (b.x - a.x), mse_coef * (b.y - a.y), mse_coef * (b.z - a.z), mse_coef * (b.w - a.w));
    if (dpm) {
      const uint2 q = *reinterpret_cast<const uint2*>(dpm + r * ld);
      const float2 p0 = unpack_bf16x2(q.x), p1 = unpack_bf16x2(q.y);
      o.x += p0.x; o.y += p0.y; o.z += p1.x; o.w += p1.y;
    }
    reinterpret_cast<float4*>(dy)[r] = o;
  }
}
// DFC tap without a BatchNorm behind it (c10): dc = coef * (a - a_other) * (a > 0)      bf16, n % 8 == 0
__global__ void tap_grad_relu_kernel(const __nv_bfloat16* __restrict__ a, const __nv_bfloat16* __restrict__ other,
                                     float coef, long long n, __nv_bfloat16* __restrict__ dc) {
  pdl_prologue();
  const long long nv = n / 8;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < nv;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const uint4 qa = reinterpret_cast<const uint4*>(a)[i];
    const uint4 qb = reinterpret_cast<const uint4*>(other)[i];
    const uint32_t ua[4] = {qa.x, qa.y, qa.z, qa.w};
    const uint32_t ub[4] = {qb.x, qb.y, qb.z, qb.w};
    uint32_t o[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 fa = unpack_bf16x2(ua[j]), fb = unpack_bf16x2(ub[j]);
      o[j] = pack_bf16x2(fa.x > 0.f ? coef * (fa.x - fb.x) : 0.f, fa.y > 0.f ? coef * (fa.y - fb.y) : 0.f);
    }
    reinterpret_cast<uint4*>(dc)[i] = make_uint4(o[0], o[1], o[2], o[3]);
  }
}

// ---- bias gradient: db[c] = sum_rows dy[r][c]  (bf16 rows; one block per 8-channel group; deterministic) ----
__global__ void __launch_bounds__(256) bias_grad_kernel(const __nv_bfloat16* __restrict__ dy, int ld, long long rows,
                                                        int C, float* __restrict__ db) {
  pdl_prologue();
  const int c = blockIdx.x;
  if (c >= C) return;
  double acc = 0.0;
  float f = 0.f;
  int cnt = 0;
  for (long long r = threadIdx.x; r < rows; r += blockDim.x) {
    f += bf2f(dy[r * ld + c]);
    if (++cnt == 64) {
      acc += f;
      f = 0.f;
      cnt = 0;
    }
  }
  acc += f;
  acc = warp_sum(acc);
  __shared__ double red[8];
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < 8; ++w) t += red[w];
    db[c] = static_cast<float>(t);
  }
}

// ---- Adam (Keras form) ------------------------------------------------------------------------------
// state[0] = t (as double), state[1] = lr_t.  adam_tick advances t and recomputes lr_t on the device so the
// whole train step (incl. the optimiser) can be replayed from a CUDA graph.
__global__ void adam_tick_kernel(double* __restrict__ state, double lr, double b1, double b2) {
  pdl_prologue();
  const double t = state[0] + 1.0;
  state[0] = t;
  state[1] = lr * sqrt(1.0 - pow(b2, t)) / (1.0 - pow(b1, t));
}
__device__ __forceinline__ void adam_one(float& p, float g, float& m, float& v, float b1, float b2, float eps, float lr_t,
                                         float grad_scale) {
  const float gi = g * grad_scale;
  const float mi = b1 * m + (1.f - b1) * gi;
  const float vi = b2 * v + (1.f - b2) * gi * gi;
  m = mi;
  v = vi;
  p -= lr_t * mi / (sqrtf(vi) + eps);
}
__global__ void adam_update_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                   float* __restrict__ v, const double* __restrict__ state, float b1, float b2, float eps,
                                   float grad_scale, long long n, int vec4) {
  pdl_prologue();
  const float lr_t = static_cast<float>(state[1]);
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  const long long t0 = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  long long done = 0;
  if (vec4) {  // all four arrays 16-byte aligned: 16-byte accesses for the bulk (same per-element arithmetic)
    const long long n4 = n >> 2;
    for (long long i = t0; i < n4; i += stride) {
      float4 pp = reinterpret_cast<float4*>(p)[i], mm = reinterpret_cast<float4*>(m)[i], vv = reinterpret_cast<float4*>(v)[i];
      const float4 gg = reinterpret_cast<const float4*>(g)[i];
      adam_one(pp.x, gg.x, mm.x, vv.x, b1, b2, eps, lr_t, grad_scale);
      adam_one(pp.y, gg.y, mm.y, vv.y, b1, b2, eps, lr_t, grad_scale);
      adam_one(pp.z, gg.z, mm.z, vv.z, b1, b2, eps, lr_t, grad_scale);
      adam_one(pp.w, gg.w, mm.w, vv.w, b1, b2, eps, lr_t, grad_scale);
      reinterpret_cast<float4*>(m)[i] = mm;
      reinterpret_cast<float4*>(v)[i] = vv;
      reinterpret_cast<float4*>(p)[i] = pp;
    }
    done = n4 << 2;
  }
  for (long long i = done + t0; i < n; i += stride) adam_one(p[i], g[i], m[i], v[i], b1, b2, eps, lr_t, grad_scale);
}

// ------------------------------------------------------------------------------------------------
// Data parallel: gradient ALL-REDUCE over NVLink peer memory fused with the Keras-Adam update (replaces
// NCCL all-reduce of the 3.36 MB flat gradient + adam_update).  Symmetric buffer per rank:
//   flags uint64 [2][world][nchunks]   data float [2][world][nchunks*kAdamChunk]   (parity = epoch & 1)
// Block b pushes its chunk of the LOCAL gradient into every rank's data[par][my_rank], publishes flag b on every rank,
// waits for flag b of every rank in its own buffer, then adds the copies in rank order (bit-identical on every rank),
// writes the global gradient back to g and applies Adam to the chunk.  Chunks pipeline independently.
// ------------------------------------------------------------------------------------------------
static constexpr int kAdamChunk = 4096;

__global__ void __launch_bounds__(256) adam_allreduce_kernel(float* __restrict__ p, float* __restrict__ g, float* __restrict__ m,
                                                             float* __restrict__ v, const double* __restrict__ state, float b1,
                                                             float b2, float eps, float grad_scale, long long n,
                                                             const unsigned long long* __restrict__ peers, int world, int rank,
                                                             int nchunks, size_t data_off, const long long* __restrict__ epoch_p,
                                                             long long timeout) {
  pdl_prologue();
  const unsigned long long epoch = static_cast<unsigned long long>(*epoch_p);
  const int par = static_cast<int>(epoch & 1ull);
  const long long i0 = static_cast<long long>(blockIdx.x) * kAdamChunk;
  const size_t slice = static_cast<size_t>(nchunks) * kAdamChunk;  // floats per (parity, rank)
  // 1. push (float4; the flat buffers are padded to a multiple of the chunk by the caller's layout, guard the tail)
  for (int r = 0; r < world; ++r) {
    float* dst = reinterpret_cast<float*>(peers[r] + data_off) + (static_cast<size_t>(par) * world + rank) * slice + i0;
    for (int j = threadIdx.x * 4; j < kAdamChunk; j += 256 * 4) {
      const long long i = i0 + j;
      if (i + 3 < n) {
        *reinterpret_cast<float4*>(dst + j) = *reinterpret_cast<const float4*>(g + i);
      } else {
        for (int k = 0; k < 4; ++k)
          if (i + k < n) dst[j + k] = g[i + k];
      }
    }
  }
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x < world) {
    const int r = threadIdx.x;
    unsigned long long* rflag = reinterpret_cast<unsigned long long*>(peers[r]) +
                                (static_cast<size_t>(par) * world + rank) * nchunks + blockIdx.x;
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(rflag), "l"(epoch) : "memory");
    const unsigned long long* lflag = reinterpret_cast<const unsigned long long*>(peers[rank]) +
                                      (static_cast<size_t>(par) * world + r) * nchunks + blockIdx.x;
    const long long t_start = clock64();
    unsigned long long seen;
    do {
      asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(seen) : "l"(lflag) : "memory");
      if (seen != epoch && clock64() - t_start > timeout) {
        printf("icsg3d: gradient all-reduce timeout rank %d chunk %d waiting for rank %d (epoch %llu)\n", rank, blockIdx.x, r,
               epoch);
        __trap();
      }
    } while (seen != epoch);
  }
  __syncthreads();
  // 2. rank-ordered sum + Adam
  const float lr_t = static_cast<float>(state[1]);
  const float* mine = reinterpret_cast<const float*>(peers[rank] + data_off) + static_cast<size_t>(par) * world * slice + i0;
  for (int j = threadIdx.x; j < kAdamChunk; j += 256) {
    const long long i = i0 + j;
    if (i >= n) break;
    float gs = 0.f;
    for (int r = 0; r < world; ++r) gs += __ldcv(mine + static_cast<size_t>(r) * slice + j);
    g[i] = gs;
    const float gi = gs * grad_scale;
    const float mi = b1 * m[i] + (1.f - b1) * gi;
    const float vi = b2 * v[i] + (1.f - b2) * gi * gi;
    m[i] = mi;
    v[i] = vi;
    p[i] -= lr_t * mi / (sqrtf(vi) + eps);
  }
}

}  // namespace icsg3d

using namespace icsg3d;
#define ST static_cast<cudaStream_t>(stream)

extern "C" int icsg3d_pack_vae_input(const float* m, const float* cond, int ncond, int B, int64_t vox, void* xe,
                                     void* xp, void* stream) {
  ICSG_REQUIRE(m && (xe || xp), "pack_vae_input: null pointer");
  ICSG_REQUIRE(!xe || (cond && ncond >= 0 && ncond <= 12), "pack_vae_input: ncond must be <= 12");
  const long long total = static_cast<long long>(B) * vox;
  launch_k(pack_vae_input_kernel, grid1d(total), 256, 0, ST, m, cond, ncond, vox, total, static_cast<__nv_bfloat16*>(xe),
                                                      static_cast<__nv_bfloat16*>(xp));
  ICSG_CHECK_LAUNCH();
  return ICSG3D_OK;
}

extern "C" int icsg3d_f32_to_bf16_rows(const float* src, int c, int64_t rows, void* dst, int ld, void* stream) {
  ICSG_REQUIRE(src && dst && c <= ld, "f32_to_bf16_rows: bad arguments");
  launch_k(f32_to_bf16_rows_kernel, grid1d(rows * c), 256, 0, ST, src, c, rows, static_cast<__nv_bfloat16*>(dst), ld);
  ICSG_CHECK_LAUNCH();
  return ICSG3D_OK;
}

extern "C" int icsg3d_bf16_rows_to_f32(const void* src, int ld, int c, int64_t rows, float* dst, void* stream) {
  ICSG_REQUIRE(src && dst && c <= ld, "bf16_rows_to_f32: bad arguments");
  launch_k(bf16_rows_to_f32_kernel, grid1d(rows * c), 256, 0, ST, static_cast<const __nv_bfloat16*>(src), ld, c, rows, dst);
  ICSG_CHECK_LAUNCH();
  return ICSG3D_OK;
}

extern "C" int icsg3d_dense_fwd(const float* x1, int k1, const float* x2, int k2, const float* w, const float* bias,
                                int act, int B, int N, float* y, void* stream) {
  ICSG_REQUIRE(x1 && w && y && (k2 == 0 || x2), "dense_fwd: null pointer");
  launch_k(dense_fwd_kernel, dim3(ceil_div(N, 32), B), 256, 0, ST, x1, k1, x2, k2, w, bias, act, B, N, y);
  ICSG_CHECK_LAUNCH();
  return ICSG3D_OK;
}

extern "C" int icsg3d_dense_bwd(const float* x1, int k1, const float* x2, int k2, const float* w, const float* y,
                                int act, float* dy, int B, int N, float* dx1, int accumulate_dx, float* dw, float* db,
                                void* stream) {
  ICSG_REQUIRE(x1 && w && dy && dw && db, "dense_bwd: null pointer");
  if (act == ICSG3D_ACT_RELU) {
    ICSG_REQUIRE(y, "dense_bwd: relu needs the forward output");
    launch_k(dense_bwd_mask_kernel, ceil_div(B * N, 128), 128, 0, ST, dy, y, B * N);
    ICSG_CHECK_LAUNCH();
  }
  if (dx1) {
    launch_k(dense_bwd_input_kernel, ceil_div(B * k1, 8), 256, 0, ST, dy, w, B, N, k1, dx1, accumulate_dx);
    ICSG_CHECK_LAUNCH();
  }
  launch_k(dense_bwd_weight_kernel, ceil_div((k1 + k2 + 1) * N, 128), 128, 0, ST, x1, k1, x2, k2, dy, B, N, dw, db);
  ICSG_CHECK_LAUNCH();
  return ICSG3D_OK;
}

extern "C" int icsg3d_reparam_fwd(const float* mu, const float* lv, const float* eps, int B, int L, float* z, float* kl,
                                  void* stream) {
  ICSG_REQUIRE(mu && lv && eps && z && kl, "reparam_fwd: null pointer");
  launch_k(reparam_fwd_kernel, B, 128, 0, ST, mu, lv, eps, L, z, kl);
  ICSG_CHECK_LAUNCH();
  return ICSG3D_OK;
}

extern "C" int icsg3d_reparam_bwd(const float* dz, const float* mu, const float* lv, const float* eps, float kl_coef,
                                  int B, int L, float* dmu, float* dlv, void* stream) {
  ICSG_REQUIRE(dz && mu && lv && eps && dmu && dlv, "reparam_bwd: null pointer");
  launch_k(reparam_bwd_kernel, ceil_div(B * L, 128), 128, 0, ST, dz, mu, lv, eps, kl_coef, B * L, dmu, dlv);
  ICSG_CHECK_LAUNCH();
  return ICSG3D_OK;
}

extern "C" int icsg3d_leaky_bwd_rows(const float* dy, const float* y, float alpha, int c, int64_t rows, void* dst, int ld,
                                     void* stream) {
  ICSG_REQUIRE(dy && y && dst && c <= ld, "leaky_bwd_rows: bad arguments");
  launch_k(leaky_bwd_rows_kernel, grid1d(rows * c), 256, 0, ST, dy, y, alpha, c, rows, static_cast<__nv_bfloat16*>(dst), ld);
  ICSG_CHECK_LAUNCH();
  return ICSG3D_OK;
}

extern "C" int icsg3d_sqdiff_nparts(int64_t n) { return grid1d(n / 8, 256, 4); }

extern "C" int icsg3d_sqdiff_partials(const void* a, const void* b, int dtype, int64_t n, double* partials, int nparts,
                                      void* stream) {
  ICSG_REQUIRE(a && b && partials, "sqdiff_partials: null pointer");
  ICSG_REQUIRE(nparts == icsg3d_sqdiff_nparts(n), "sqdiff_partials: nparts mismatch");
  if (dtype == ICSG3D_DT_BF16) {
    ICSG_REQUIRE(n % 8 == 0, "sqdiff_partials: n must be a multiple of 8 for bf16");
    launch_k(sqdiff_partials_kernel<__nv_bfloat16>, nparts, 256, 0, ST, static_cast<const __nv_bfloat16*>(a),
                                                                 static_cast<const __nv_bfloat16*>(b), n, partials);
  } else {
    ICSG_REQUIRE(n % 4 == 0, "sqdiff_partials: n must be a multiple of 4 for fp32");
    launch_k(sqdiff_partials_kernel<float>, nparts, 256, 0, ST, static_cast<const float*>(a), static_cast<const float*>(b), n, partials);
  }
  ICSG_CHECK_LAUNCH();
  return ICSG3D_OK;
}

extern "C" int icsg3d_vae_loss_assemble(const double* partials, const int* nparts, int stride, int nterms,
                                        const double* scales, const float* kl, int B, double kl_scale, float alpha,
                                        float beta, float* out, double* raw, void* stream) {
  ICSG_REQUIRE(partials && nparts && scales && kl && out, "vae_loss_assemble: null pointer");
  ICSG_REQUIRE(nterms >= 1 && nterms <= 7, "vae_loss_assemble: nterms must be in [1,7]");
  launch_k(vae_loss_assemble_kernel, 1, 256, 0, ST, partials, nparts, stride, nterms, scales, kl, B, kl_scale, alpha, beta, out, raw);
  ICSG_CHECK_LAUNCH();
  return ICSG3D_OK;
}

extern "C" int icsg3d_xhat_grad(const float* x, const float* xhat, float mse_coef, const void* dpm, int ld, int64_t rows,
                                float* dy, void* stream) {
  ICSG_REQUIRE(x && xhat && dy, "xhat_grad: null pointer");
  launch_k(xhat_grad_kernel, grid1d(rows), 256, 0, ST, x, xhat, mse_coef, static_cast<const __nv_bfloat16*>(dpm), ld, rows, dy);
  ICSG_CHECK_LAUNCH();
  return ICSG3D_OK;
}

extern "C" int icsg3d_tap_grad_relu(const void* a, const void* other, float coef, int64_t n, void* dc, void* stream) {
  ICSG_REQUIRE(a && other && dc && n % 8 == 0, "tap_grad_relu: bad arguments");
  launch_k(tap_grad_relu_kernel, grid1d(n / 8), 256, 0, ST, static_cast<const __nv_bfloat16*>(a),
                                                     static_cast<const __nv_bfloat16*>(other), coef, n,
                                                     static_cast<__nv_bfloat16*>(dc));
  ICSG_CHECK_LAUNCH();
  return ICSG3D_OK;
}

extern "C" int icsg3d_bias_grad(const void* dy, int ld, int64_t rows, int C, float* db, void* stream) {
  ICSG_REQUIRE(dy && db && C <= ld, "bias_grad: bad arguments");
  launch_k(bias_grad_kernel, C, 256, 0, ST, static_cast<const __nv_bfloat16*>(dy), ld, rows, C, db);
  ICSG_CHECK_LAUNCH();
  return ICSG3D_OK;
}

extern "C" int icsg3d_adam_keras_step(float* p, const float* g, float* m, float* v, double* state, double lr, double beta1,
                                      double beta2, double eps, float grad_scale, int64_t n, void* stream) {
  ICSG_REQUIRE(p && g && m && v && state, "adam_keras_step: null pointer");
  launch_k(adam_tick_kernel, 1, 1, 0, ST, state, lr, beta1, beta2);
  ICSG_CHECK_LAUNCH();
  const int vec4 = ((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(m) |
                     reinterpret_cast<uintptr_t>(v)) & 15) == 0;
  launch_k(adam_update_kernel, grid1d(vec4 ? (n + 3) / 4 : n), 256, 0, ST, p, g, m, v, state, static_cast<float>(beta1),
           static_cast<float>(beta2), static_cast<float>(eps), grad_scale, n, vec4);
  ICSG_CHECK_LAUNCH();
  return ICSG3D_OK;
}


extern "C" int64_t icsg3d_adam_allreduce_buffer_bytes(int world, int64_t n) {
  if (world < 1 || n < 1) return -1;
  const int64_t nchunks = (n + kAdamChunk - 1) / kAdamChunk;
  const int64_t flags = ((2 * world * nchunks * 8 + 255) / 256) * 256;
  return flags + 2 * world * nchunks * kAdamChunk * 4;
}

extern "C" int icsg3d_adam_keras_allreduce_step(float* p, float* g, float* m, float* v, double* state, double lr, double beta1,
                                                double beta2, double eps, float grad_scale, int64_t n, const uint64_t* peers,
                                                int world, int rank, const int64_t* epoch, void* stream) {
  ICSG_REQUIRE(p && g && m && v && state && peers && epoch, "adam_keras_allreduce_step: null pointer");
  ICSG_REQUIRE(world >= 1 && world <= 16 && rank >= 0 && rank < world && n > 0, "adam_keras_allreduce_step: bad world/rank/n");
  ICSG_REQUIRE((reinterpret_cast<uintptr_t>(g) & 15) == 0, "adam_keras_allreduce_step: g must be 16-byte aligned");
  const int64_t nchunks = (n + kAdamChunk - 1) / kAdamChunk;
  const size_t data_off = static_cast<size_t>(((2 * world * nchunks * 8 + 255) / 256) * 256);
  launch_k(adam_tick_kernel, 1, 1, 0, ST, state, lr, beta1, beta2);
  ICSG_CHECK_LAUNCH();
  launch_k(adam_allreduce_kernel, static_cast<int>(nchunks), 256, 0, ST, 
      p, g, m, v, state, static_cast<float>(beta1), static_cast<float>(beta2), static_cast<float>(eps), grad_scale, n,
      reinterpret_cast<const unsigned long long*>(peers), world, rank, static_cast<int>(nchunks), data_off,
      reinterpret_cast<const long long*>(epoch), peer_timeout_cycles());
  ICSG_CHECK_LAUNCH();
  return ICSG3D_OK;
}
