// U-Net output heads (unet/unet.py:339-352): `soft` = Conv3D(95, 1x1x1, softmax), `sig` = Conv3D(1, 1x1x1, sigmoid),
// both on c18.  The two 1x1x1 convolutions run as ONE tcgen05 GEMM [voxels x 128] x [128 x 96] (95 soft logits +
// 1 sigmoid logit = 96 columns, conv3d_k1_igemm); this file holds the weight packing for that GEMM and the fused
// per-voxel pass over the fp32 logits:
//   softmax -> class-weighted categorical cross-entropy (unet.py:196-221: renormalise, clip to [1e-7, 1-1e-7],
//   -sum_c y_c log p_c w_c), sigmoid binary cross-entropy in its logits form, argmax species label,
//   sigmoid probability, the f1_m / wr_m metric counts (unet.py:159-193) and d(loss)/d(logits) for the backward pass.
// One warp per voxel, three columns per lane, warp-shuffle reductions, fp64 per-block loss partials.
#include "common.cuh"

namespace icsg3d {

// wcat[k][n]: n < c1 -> w_soft[k][n], n == c1 -> w_sig[k], else 0.   fprop operand [nout][cin] (K-major),
// dgrad operand [cin][nout] (N = cin rows, K = nout contiguous).
__global__ void pack_heads_kernel(const float* __restrict__ w_soft, const float* __restrict__ w_sig,
                                  const float* __restrict__ b_soft, const float* __restrict__ b_sig, int cin, int c1,
                                  int nout, __nv_bfloat16* __restrict__ wf, __nv_bfloat16* __restrict__ wd,
                                  float* __restrict__ bias) {
  pdl_prologue();
  const int total = cin * nout;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int k = i / nout, n = i % nout;
    float v = 0.f;
    if (n < c1) v = w_soft[k * c1 + n];
    else if (n == c1) v = w_sig[k];
    wf[n * cin + k] = f2bf(v);
    wd[k * nout + n] = f2bf(v);
  }
  for (int n = blockIdx.x * blockDim.x + threadIdx.x; n < nout; n += gridDim.x * blockDim.x)
    bias[n] = n < c1 ? b_soft[n] : (n == c1 ? b_sig[0] : 0.f);
}

__global__ void unpack_heads_grad_kernel(const float* __restrict__ dwcat, const double* __restrict__ colsum, int cin,
                                         int c1, int nout, float* __restrict__ dw_soft, float* __restrict__ dw_sig,
                                         float* __restrict__ db_soft, float* __restrict__ db_sig) {
  pdl_prologue();
  const int total = cin * (c1 + 1);
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int k = i / (c1 + 1), n = i % (c1 + 1);
    const float v = dwcat[k * nout + n];
    if (n < c1) dw_soft[k * c1 + n] = v;
    else dw_sig[k] = v;
  }
  for (int n = blockIdx.x * blockDim.x + threadIdx.x; n <= c1; n += gridDim.x * blockDim.x) {
    if (n < c1) db_soft[n] = static_cast<float>(colsum[n]);
    else db_sig[0] = static_cast<float>(colsum[n]);
  }
}

static constexpr int kHeadsThreads = 256;
static constexpr int kHeadsTerms = 6;  // soft loss, sig loss, tp, predicted, tp_w, possible_w

// logits fp32 [M][ld] (c1 soft columns + 1 sigmoid column); species uint8 [M]; class_w fp32 [c1].
__global__ void __launch_bounds__(kHeadsThreads) heads_loss_kernel(const float* __restrict__ logits, int ld, int c1,
                                                                   const uint8_t* __restrict__ species,
                                                                   const float* __restrict__ class_w, long long M,
                                                                   float inv_count, uint8_t* __restrict__ argmax_out,
                                                                   float* __restrict__ sig_prob,
                                                                   float* __restrict__ probs,
                                                                   __nv_bfloat16* __restrict__ dlogits, int ldd,
                                                                   double* __restrict__ partials, int dl_f32) {
  pdl_prologue();
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int warps = kHeadsThreads / 32;
  double acc[kHeadsTerms] = {0, 0, 0, 0, 0, 0};
  for (long long v = static_cast<long long>(blockIdx.x) * warps + w; v < M; v += static_cast<long long>(gridDim.x) * warps) {
    const float* row = logits + v * ld;
    float x[3];
    int col[3];
    float mx = -INFINITY;
    int amax = 0;
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      col[j] = lane + 32 * j;
      x[j] = col[j] < c1 ? row[col[j]] : -INFINITY;
      if (x[j] > mx) {
        mx = x[j];
        amax = col[j];
      }
    }
    // warp arg-max with first-index tie rule (np.argmax, generate.py:221)
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float om = __shfl_xor_sync(0xffffffffu, mx, o);
      const int oa = __shfl_xor_sync(0xffffffffu, amax, o);
      if (om > mx || (om == mx && oa < amax)) {
        mx = om;
        amax = oa;
      }
    }
    float e[3], se = 0.f;
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      e[j] = col[j] < c1 ? expf(x[j] - mx) : 0.f;
      se += e[j];
    }
    se = warp_sum(se);
    const float inv = 1.f / se;
    const int t = species ? static_cast<int>(species[v]) : 0;
    const float xs = row[c1];  // sigmoid logit
    float pt = 0.f;
    int n_pred = 0;
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      const float pj = e[j] * inv;
      if (col[j] == t) pt = pj;
      if (col[j] < c1 && pj > 0.5f) n_pred = 1;  // K.round(p) == 1 (round-half-even: 0.5 -> 0)
    }
    pt = __shfl_sync(0xffffffffu, pt, t & 31);  // lane (t % 32) owns column t
    n_pred = __any_sync(0xffffffffu, n_pred) ? 1 : 0;
    const bool in_range = pt >= 1e-7f && pt <= 1.f - 1e-7f;
    const float ptc = fminf(fmaxf(pt, 1e-7f), 1.f - 1e-7f);
    const float wt = class_w ? class_w[t < c1 ? t : 0] : 1.f;
    const float tb = t != 0 ? 1.f : 0.f;
    const float sp = 1.f / (1.f + expf(-xs));
    if (lane == 0) {
      acc[0] += static_cast<double>(-wt * logf(ptc));
      acc[1] += static_cast<double>(fmaxf(xs, 0.f) - xs * tb + log1pf(expf(-fabsf(xs))));
      acc[2] += pt > 0.5f ? 1.0 : 0.0;
      acc[3] += n_pred;
      acc[4] += (t != 0 && pt > 0.5f) ? 1.0 : 0.0;
      acc[5] += t != 0 ? 1.0 : 0.0;
      if (argmax_out) argmax_out[v] = static_cast<uint8_t>(amax);
      if (sig_prob) sig_prob[v] = sp;
    }
    if (probs) {
#pragma unroll
      for (int j = 0; j < 3; ++j)
        if (col[j] < c1) probs[v * c1 + col[j]] = e[j] * inv;
    }
    if (dlogits) {
      const float gs = in_range ? wt * inv_count : 0.f;
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        float gv = 0.f;
        if (col[j] < c1) gv = gs * (e[j] * inv - (col[j] == t ? 1.f : 0.f));
        else if (col[j] == c1) gv = (sp - tb) * inv_count;
        if (col[j] < ldd) {
          if (dl_f32) reinterpret_cast<float*>(dlogits)[v * ldd + col[j]] = gv;  // fp32-class backward
          else dlogits[v * ldd + col[j]] = f2bf(gv);
        }
      }
    }
  }
  __shared__ double red[kHeadsThreads / 32][kHeadsTerms];
  if (lane == 0)
    for (int i = 0; i < kHeadsTerms; ++i) red[w][i] = acc[i];
  __syncthreads();
  if (threadIdx.x < kHeadsTerms) {
    double tsum = 0.0;
    for (int i = 0; i < warps; ++i) tsum += red[i][threadIdx.x];
    partials[static_cast<size_t>(blockIdx.x) * kHeadsTerms + threadIdx.x] = tsum;
  }
}

// out = [loss, soft_loss, sig_loss, f1_m, wr_m] (the order Keras reports; unet.py:249-259), raw[6] = the term sums.
__global__ void __launch_bounds__(256) heads_loss_finalize_kernel(const double* __restrict__ partials, int nparts, double count,
                                                                  float* __restrict__ out, double* __restrict__ raw) {
  pdl_prologue();
  // fixed-order two-level sum of the per-block partials (deterministic): 32 strided lanes per term, then lane order
  __shared__ double lanes[32][8];
  __shared__ double term[kHeadsTerms];
  const int tm = threadIdx.x & 7, ln = threadIdx.x >> 3;
  if (tm < kHeadsTerms) {
    double t = 0.0;
    for (int p = ln; p < nparts; p += 32) t += partials[static_cast<size_t>(p) * kHeadsTerms + tm];
    lanes[ln][tm] = t;
  }
  __syncthreads();
  if (threadIdx.x < kHeadsTerms) {
    double t = 0.0;
    for (int l = 0; l < 32; ++l) t += lanes[l][threadIdx.x];
    term[threadIdx.x] = t;
    if (raw) raw[threadIdx.x] = t;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    const double eps = 1e-7;  // K.epsilon()
    const double soft = term[0] / count, sig = term[1] / count;
    const double recall = term[2] / (count + eps);
    const double precision = term[2] / (term[3] + eps);
    const double f1 = 2.0 * ((precision * recall) / (precision + recall + eps));
    const double wr = term[4] / (term[5] + eps);
    out[0] = static_cast<float>(soft + sig);
    out[1] = static_cast<float>(soft);
    out[2] = static_cast<float>(sig);
    out[3] = static_cast<float>(f1);
    out[4] = static_cast<float>(wr);
  }
}

}  // namespace icsg3d

using namespace icsg3d;

extern "C" int icsg3d_pack_heads_w(const float* w_soft, const float* w_sig, const float* b_soft, const float* b_sig, int cin,
                                   int c1, int nout, void* wf, void* wd, float* bias, void* stream) {
  ICSG_REQUIRE(w_soft && w_sig && b_soft && b_sig && wf && wd && bias, "pack_heads_w: null pointer");
  ICSG_REQUIRE(nout % 16 == 0 && nout >= c1 + 1 && cin % 16 == 0, "pack_heads_w: bad sizes");
  launch_k(pack_heads_kernel, ceil_div(cin * nout, 256), 256, 0, static_cast<cudaStream_t>(stream), 
      w_soft, w_sig, b_soft, b_sig, cin, c1, nout, static_cast<__nv_bfloat16*>(wf), static_cast<__nv_bfloat16*>(wd), bias);
  ICSG_CHECK_LAUNCH();
  return ICSG3D_OK;
}

extern "C" int icsg3d_unpack_heads_grad(const float* dwcat, const double* colsum, int cin, int c1, int nout, float* dw_soft,
                                        float* dw_sig, float* db_soft, float* db_sig, void* stream) {
  ICSG_REQUIRE(dwcat && colsum && dw_soft && dw_sig && db_soft && db_sig, "unpack_heads_grad: null pointer");
  launch_k(unpack_heads_grad_kernel, ceil_div(cin * (c1 + 1), 256), 256, 0, static_cast<cudaStream_t>(stream), 
      dwcat, colsum, cin, c1, nout, dw_soft, dw_sig, db_soft, db_sig);
  ICSG_CHECK_LAUNCH();
  return ICSG3D_OK;
}

extern "C" int icsg3d_heads_loss_nparts(int64_t M) {
  long long b = (M + 7) / 8;
  int sms = sm_count();
  if (sms <= 0) sms = 148;
  if (b > static_cast<long long>(sms) * 8) b = static_cast<long long>(sms) * 8;
  return static_cast<int>(b < 1 ? 1 : b);
}

static int heads_loss_impl(const float* logits, int ld, int c1, const uint8_t* species, const float* class_w, int64_t M,
                           float inv_count, uint8_t* argmax_out, float* sig_prob, float* probs, void* dlogits, int ldd,
                           double* partials, int nparts, int dl_f32, void* stream) {
  ICSG_REQUIRE(logits && partials, "heads_loss: null pointer");
  ICSG_REQUIRE(c1 >= 1 && c1 <= 95 && ld > c1, "heads_loss: c1 must be in [1,95] and ld > c1 (three columns per lane)");
  ICSG_REQUIRE(!dlogits || (species && ldd > c1 && ldd <= 96), "heads_loss: gradient needs labels and c1 < ldd <= 96");
  ICSG_REQUIRE(nparts == icsg3d_heads_loss_nparts(M), "heads_loss: nparts mismatch");
  launch_k(heads_loss_kernel, nparts, kHeadsThreads, 0, static_cast<cudaStream_t>(stream), 
      logits, ld, c1, species, class_w, M, inv_count, argmax_out, sig_prob, probs, static_cast<__nv_bfloat16*>(dlogits), ldd,
      partials, dl_f32);
  ICSG_CHECK_LAUNCH();
  return ICSG3D_OK;
}

extern "C" int icsg3d_heads_loss(const float* logits, int ld, int c1, const uint8_t* species, const float* class_w,
                                 int64_t M, float inv_count, uint8_t* argmax_out, float* sig_prob, float* probs, void* dlogits,
                                 int ldd, double* partials, int nparts, void* stream) {
  return heads_loss_impl(logits, ld, c1, species, class_w, M, inv_count, argmax_out, sig_prob, probs, dlogits, ldd, partials,
                         nparts, 0, stream);
}

extern "C" int icsg3d_heads_loss_f32grad(const float* logits, int ld, int c1, const uint8_t* species, const float* class_w,
                                         int64_t M, float inv_count, uint8_t* argmax_out, float* sig_prob, float* probs,
                                         float* dlogits, int ldd, double* partials, int nparts, void* stream) {
  return heads_loss_impl(logits, ld, c1, species, class_w, M, inv_count, argmax_out, sig_prob, probs, dlogits, ldd, partials,
                         nparts, 1, stream);
}

extern "C" int icsg3d_heads_loss_finalize(const double* partials, int nparts, double count, float* out, double* raw,
                                          void* stream) {
  ICSG_REQUIRE(partials && out && count > 0, "heads_loss_finalize: bad arguments");
  launch_k(heads_loss_finalize_kernel, 1, 256, 0, static_cast<cudaStream_t>(stream), partials, nparts, count, out, raw);
  ICSG_CHECK_LAUNCH();
  return ICSG3D_OK;
}
