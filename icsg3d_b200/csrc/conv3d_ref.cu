// CUDA-core direct 3x3x3 convolutions with fp32 accumulation.
// TEST-ONLY on-device cross-checks for the tcgen05 kernels: same operand layouts, no tensor cores,
// no TMA, so a descriptor bug in the fast path cannot hide in a shared harness bug.
#include "common.cuh"

namespace icsg3d {

// One thread per (voxel, output channel).  wpack is [27][nout][cin] bf16 (the igemm B operand).
__global__ void ref_conv3d_k3_kernel(const __nv_bfloat16* __restrict__ x, int ldx,
                                     const __nv_bfloat16* __restrict__ wpack, const float* __restrict__ bias,
                                     float* __restrict__ y, int ldy, int B, int D, int H, int W, int cin, int nout,
                                     int act, float alpha) {
  const long long total = static_cast<long long>(B) * D * H * W * nout;
  for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int co = static_cast<int>(idx % nout);
    long long pix = idx / nout;
    const int w = static_cast<int>(pix % W);
    long long t = pix / W;
    const int h = static_cast<int>(t % H);
    t /= H;
    const int d = static_cast<int>(t % D);
    const int n = static_cast<int>(t / D);
    float acc = bias ? bias[co] : 0.f;
    for (int kd = 0; kd < 3; ++kd) {
      const int dd = d + kd - 1;
      if (dd < 0 || dd >= D) continue;
      for (int kh = 0; kh < 3; ++kh) {
        const int hh = h + kh - 1;
        if (hh < 0 || hh >= H) continue;
        for (int kw = 0; kw < 3; ++kw) {
          const int ww = w + kw - 1;
          if (ww < 0 || ww >= W) continue;
          const int tap = (kd * 3 + kh) * 3 + kw;
          const __nv_bfloat16* xp = x + ((((static_cast<long long>(n) * D + dd) * H + hh) * W + ww)) * ldx;
          const __nv_bfloat16* wp = wpack + (static_cast<long long>(tap) * nout + co) * cin;
          for (int ci = 0; ci < cin; ++ci) acc += bf2f(xp[ci]) * bf2f(wp[ci]);
        }
      }
    }
    if (act == ICSG3D_ACT_RELU) acc = fmaxf(acc, 0.f);
    else if (act == ICSG3D_ACT_LEAKY) acc = acc > 0.f ? acc : alpha * acc;
    y[pix * ldy + co] = acc;
  }
}

// One block per (tap, ci); threads over co accumulate over all voxels (slow, simple, deterministic).
__global__ void ref_conv3d_k3_wgrad_kernel(const __nv_bfloat16* __restrict__ x, int ldx,
                                           const __nv_bfloat16* __restrict__ dy, int ldy, float* __restrict__ dw,
                                           int B, int D, int H, int W, int cin, int cout) {
  const int tap = blockIdx.x / cin;
  const int ci = blockIdx.x % cin;
  const int kd = tap / 9, kh = (tap / 3) % 3, kw = tap % 3;
  __shared__ float red[256];
  for (int co = 0; co < cout; ++co) {
    float acc = 0.f;
    const long long npix = static_cast<long long>(B) * D * H * W;
    for (long long pix = threadIdx.x; pix < npix; pix += blockDim.x) {
      const int w = static_cast<int>(pix % W);
      long long t = pix / W;
      const int h = static_cast<int>(t % H);
      t /= H;
      const int d = static_cast<int>(t % D);
      const int n = static_cast<int>(t / D);
      const int dd = d + kd - 1, hh = h + kh - 1, ww = w + kw - 1;
      if (dd < 0 || dd >= D || hh < 0 || hh >= H || ww < 0 || ww >= W) continue;
      const float xv = bf2f(x[((((static_cast<long long>(n) * D + dd) * H + hh) * W + ww)) * ldx + ci]);
      acc += xv * bf2f(dy[pix * ldy + co]);
    }
    red[threadIdx.x] = acc;
    __syncthreads();
    for (int s = blockDim.x / 2; s > 0; s >>= 1) {
      if (threadIdx.x < s) red[threadIdx.x] += red[threadIdx.x + s];
      __syncthreads();
    }
    if (threadIdx.x == 0) dw[(static_cast<long long>(tap) * cin + ci) * cout + co] = red[0];
    __syncthreads();
  }
}

}  // namespace icsg3d

using namespace icsg3d;

extern "C" int icsg3d_ref_conv3d_k3(const void* x, int ldx, const void* wpack, const float* bias, float* y, int ldy,
                                    int B, int D, int H, int W, int cin, int nout, int act, float leaky_alpha,
                                    void* stream) {
  ICSG_REQUIRE(x && wpack && y, "ref_conv3d_k3: null pointer");
  const long long total = static_cast<long long>(B) * D * H * W * nout;
  const int threads = 256;
  long long blocks = (total + threads - 1) / threads;
  if (blocks > 148 * 32) blocks = 148 * 32;
  ref_conv3d_k3_kernel<<<static_cast<int>(blocks), threads, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __nv_bfloat16*>(x), ldx, static_cast<const __nv_bfloat16*>(wpack), bias, y, ldy, B, D, H, W,
      cin, nout, act, leaky_alpha);
  ICSG_CHECK_LAUNCH();
  return ICSG3D_OK;
}

extern "C" int icsg3d_ref_conv3d_k3_wgrad(const void* x, int ldx, const void* dy, int ldy, float* dw, int B, int D,
                                          int H, int W, int cin, int cout, void* stream) {
  ICSG_REQUIRE(x && dy && dw, "ref_conv3d_k3_wgrad: null pointer");
  ref_conv3d_k3_wgrad_kernel<<<27 * cin, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __nv_bfloat16*>(x), ldx, static_cast<const __nv_bfloat16*>(dy), ldy, dw, B, D, H, W, cin,
      cout);
  ICSG_CHECK_LAUNCH();
  return ICSG3D_OK;
}
