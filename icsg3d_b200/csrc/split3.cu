// "fp32-class" operand mode (north_star: 1e-4 parity tier, bit-exact argmax): every conv operand value v is carried as
// two bf16 numbers hi = bf16(v), lo = bf16(v - hi) and the tensor-core contraction computes
//     x_hi*w_hi + x_lo*w_hi + x_hi*w_lo        (fp32 accumulate; the dropped lo*lo term is ~2^-18 relative)
// WITHOUT any new conv kernel: activations are stored with 3x the channels  [hi | lo | hi]  and weights as
// [w_hi | w_hi | w_lo]  along Cin, so the ordinary bf16 implicit-GEMM kernels (fp32 output) do the rest.
// This file holds the producers of the split layout; the BatchNorm apply pass writes it directly (bn.cu, split3 mode).
#include <cuda_fp16.h>

#include "common.cuh"

namespace icsg3d {

// fmt 0: bf16 pair (16 significant bits); fmt 1: IEEE fp16 pair (22 bits — fp32 class; operands of magnitude < 65504).
// The 2-byte values are stored through __nv_bfloat16* as raw bits; the conv is told the format (icsg3d_conv3d_*_f16).
__device__ __forceinline__ void split_bf16(float v, __nv_bfloat16& hi, __nv_bfloat16& lo, int fmt) {
  if (fmt == 0) {
    hi = __float2bfloat16_rn(v);
    lo = __float2bfloat16_rn(v - __bfloat162float(hi));
  } else {
    const __half h = __float2half_rn(v);
    const __half l = __float2half_rn(v - __half2float(h));
    hi = __ushort_as_bfloat16(__half_as_ushort(h));
    lo = __ushort_as_bfloat16(__half_as_ushort(l));
  }
}

// src fp32 [rows][ld_src] (first c channels used) -> dst bf16 [rows][3*ctot]: part*ctot + coff + i  (part = hi, lo, hi)
__global__ void f32_to_split3_kernel(const float* __restrict__ src, int ld_src, int c, long long rows,
                                     __nv_bfloat16* __restrict__ dst, int ctot, int coff, int fmt) {
  pdl_prologue();
  const long long total = rows * c;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long r = i / c;
    const int ch = static_cast<int>(i - r * c);
    __nv_bfloat16 hi, lo;
    split_bf16(src[r * ld_src + ch], hi, lo, fmt);
    __nv_bfloat16* d = dst + r * 3 * ctot + coff + ch;
    d[0] = hi;
    d[ctot] = lo;
    d[2 * ctot] = hi;
  }
}

// VAE encoder / perceptual inputs in split form: xe3 [rows][48] from (M 4ch, cond one-hot ncond, 0..), xp3 [rows][48] from M;
// xp16 (optional): the plain bf16 [rows][16] perceptual-model operand (M, 0..) in the same pass.
// One row per thread, assembled in registers and written as whole 16-byte vectors (96 contiguous bytes per row).
__device__ __forceinline__ uint32_t pack_raw2(__nv_bfloat16 a, __nv_bfloat16 b) {
  return static_cast<uint32_t>(__bfloat16_as_ushort(a)) | (static_cast<uint32_t>(__bfloat16_as_ushort(b)) << 16);
}
__device__ __forceinline__ void store_split_row(__nv_bfloat16* dst, const float (&v)[16], int fmt) {
  uint32_t hi[8], lo[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    __nv_bfloat16 h0, l0, h1, l1;
    split_bf16(v[2 * i], h0, l0, fmt);
    split_bf16(v[2 * i + 1], h1, l1, fmt);
    hi[i] = pack_raw2(h0, h1);
    lo[i] = pack_raw2(l0, l1);
  }
  uint4* d = reinterpret_cast<uint4*>(dst);
  d[0] = make_uint4(hi[0], hi[1], hi[2], hi[3]);
  d[1] = make_uint4(hi[4], hi[5], hi[6], hi[7]);
  d[2] = make_uint4(lo[0], lo[1], lo[2], lo[3]);
  d[3] = make_uint4(lo[4], lo[5], lo[6], lo[7]);
  d[4] = d[0];
  d[5] = d[1];
}
// Lean split layout of the ENCODER input (bf16 pairs, 32 channels instead of 48): the one-hot condition is exact in bf16
// and needs no lo part, so  [M_hi(4) | cond(10) | M_lo(0:2)] [M_lo(2:4) | M_hi(4) | cond(10)]  against the weight pack
// [w_hi(M) | w_hi(cond) | w_hi(M 0:2)] [w_hi(M 2:4) | w_lo(M) | w_lo(cond)] (pack.cu mode 4) gives the same three products.
// The first 16 channels start with the plain bf16 operand (M_hi, cond): the bf16 filter gradient reads them in place.
__device__ __forceinline__ void store_lean_enc_row(__nv_bfloat16* dst, const float4 q, const float* c, int ncond) {
  const float mv[4] = {q.x, q.y, q.z, q.w};
  __nv_bfloat16 hi[4], lo[4], cv[10];
#pragma unroll
  for (int i = 0; i < 4; ++i) split_bf16(mv[i], hi[i], lo[i], 0);
#pragma unroll
  for (int i = 0; i < 10; ++i) cv[i] = __float2bfloat16_rn(i < ncond ? c[i] : 0.f);
  uint4* d = reinterpret_cast<uint4*>(dst);
  d[0] = make_uint4(pack_raw2(hi[0], hi[1]), pack_raw2(hi[2], hi[3]), pack_raw2(cv[0], cv[1]), pack_raw2(cv[2], cv[3]));
  d[1] = make_uint4(pack_raw2(cv[4], cv[5]), pack_raw2(cv[6], cv[7]), pack_raw2(cv[8], cv[9]), pack_raw2(lo[0], lo[1]));
  d[2] = make_uint4(pack_raw2(lo[2], lo[3]), pack_raw2(hi[0], hi[1]), pack_raw2(hi[2], hi[3]), pack_raw2(cv[0], cv[1]));
  d[3] = make_uint4(pack_raw2(cv[2], cv[3]), pack_raw2(cv[4], cv[5]), pack_raw2(cv[6], cv[7]), pack_raw2(cv[8], cv[9]));
}

__global__ void pack_vae_input_split3_kernel(const float* __restrict__ m, const float* __restrict__ cond, int ncond,
                                             long long vox, long long total, __nv_bfloat16* __restrict__ xe,
                                             __nv_bfloat16* __restrict__ xp, __nv_bfloat16* __restrict__ xp16, int fmt,
                                             int lean) {
  pdl_prologue();
  for (long long r = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; r < total;
       r += static_cast<long long>(gridDim.x) * blockDim.x) {
    float v[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = 0.f;
    const float4 q = *reinterpret_cast<const float4*>(m + r * 4);
    v[0] = q.x; v[1] = q.y; v[2] = q.z; v[3] = q.w;
    if (xp) store_split_row(xp + r * 48, v, fmt);
    if (xp16) {
      uint4* d = reinterpret_cast<uint4*>(xp16 + r * 16);
      d[0] = make_uint4(pack_bf16x2(q.x, q.y), pack_bf16x2(q.z, q.w), 0u, 0u);
      d[1] = make_uint4(0u, 0u, 0u, 0u);
    }
    if (xe) {
      const float* c = cond + (r / vox) * ncond;
      if (lean) {
        store_lean_enc_row(xe + r * 32, q, c, ncond);
      } else {
#pragma unroll
        for (int i = 0; i < 12; ++i)
          if (i < ncond) v[4 + i] = c[i];
        store_split_row(xe + r * 48, v, fmt);
      }
    }
  }
}

// w fp32 (ntaps, Cin, Cout) Keras layout -> bf16 [ntaps][cout_pad][3*cin_pad] = [w_hi | w_hi | w_lo] along Cin, with the
// same channel padding / condition fold as pack_w_fprop_kernel.
__global__ void pack_w_fprop_x3_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ wp, int ntaps, int cin, int cout,
                                       int cin_pad, int cout_pad, int cin_lead, int fold, int fold_c, int fmt, float wscale) {
  pdl_prologue();
  const long long total = static_cast<long long>(ntaps) * cout_pad * cin_pad;
  for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int ci = static_cast<int>(idx % cin_pad);
    const int co = static_cast<int>((idx / cin_pad) % cout_pad);
    const int tap = static_cast<int>(idx / (static_cast<long long>(cin_pad) * cout_pad));
    float v = 0.f;
    if (co < cout) {
      const float* wt = w + static_cast<long long>(tap) * cin * cout;
      if (fold <= 1) {
        if (ci < cin) v = wt[static_cast<long long>(ci) * cout + co];
      } else if (ci < cin_lead) {
        v = wt[static_cast<long long>(ci) * cout + co];
      } else if (ci < cin_lead + fold_c) {
        for (int r = 0; r < fold; ++r) v += wt[static_cast<long long>(cin_lead + r * fold_c + (ci - cin_lead)) * cout + co];
      }
    }
    __nv_bfloat16 hi, lo;
    split_bf16(v * wscale, hi, lo, fmt);  // power-of-two pre-scale keeps the fp16 lo parts out of the subnormal range
    __nv_bfloat16* d = wp + (static_cast<long long>(tap) * cout_pad + co) * 3 * cin_pad + ci;
    d[0] = hi;
    d[cin_pad] = hi;
    d[2 * cin_pad] = lo;
  }
}

// dgrad operand in split form: w fp32 (27|1, Cin, Cout) -> bf16 [ntaps][cin_pad][3*cout_pad] = [w_hi | w_hi | w_lo] along
// Cout (the K dimension of the input-gradient GEMM), taps mirrored like pack_w_dgrad_kernel.
__global__ void pack_w_dgrad_x3_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ wp, int ntaps, int cin, int cout,
                                       int cin_pad, int cout_pad, int fmt, float wscale) {
  pdl_prologue();
  const long long total = static_cast<long long>(ntaps) * cin_pad * cout_pad;
  for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int co = static_cast<int>(idx % cout_pad);
    const int ci = static_cast<int>((idx / cout_pad) % cin_pad);
    const int tap = static_cast<int>(idx / (static_cast<long long>(cin_pad) * cout_pad));
    float v = 0.f;
    if (ci < cin && co < cout) v = w[(static_cast<long long>(ntaps - 1 - tap) * cin + ci) * cout + co];
    __nv_bfloat16 hi, lo;
    split_bf16(v * wscale, hi, lo, fmt);
    __nv_bfloat16* d = wp + (static_cast<long long>(tap) * cin_pad + ci) * 3 * cout_pad + co;
    d[0] = hi;
    d[cout_pad] = hi;
    d[2 * cout_pad] = lo;
  }
}

// Filter gradient from split operands: the ordinary wgrad kernel run on x = [x_hi | x_lo] (2*cin_pad channels) and
// dy = [dy_hi | dy_lo] (2*cout_pad) yields P [ntaps][2 cin_pad][2 cout_pad]; dW = hi*hi + lo*hi + hi*lo (lo*lo dropped).
__global__ void wgrad_combine_x3_kernel(const float* __restrict__ P, float* __restrict__ dw, int ntaps, int cin_pad,
                                        int cout_pad) {
  pdl_prologue();
  const long long total = static_cast<long long>(ntaps) * cin_pad * cout_pad;
  for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int co = static_cast<int>(idx % cout_pad);
    const int ci = static_cast<int>((idx / cout_pad) % cin_pad);
    const int tap = static_cast<int>(idx / (static_cast<long long>(cin_pad) * cout_pad));
    const float* Pt = P + static_cast<long long>(tap) * 4 * cin_pad * cout_pad;
    const long long ld = 2ll * cout_pad;
    dw[idx] = Pt[ci * ld + co] + (Pt[(cin_pad + ci) * ld + co] + Pt[ci * ld + cout_pad + co]);
  }
}

// ---- fp32 loss-gradient seeds of the fp32-class backward (same maths as misc.cu's bf16 forms) ----
// dy = mse_coef*(xhat - x) + dpm   (rows of 4 channels; dpm fp32 with row stride ld, optional)
__global__ void xhat_grad_f32_kernel(const float* __restrict__ x, const float* __restrict__ xhat, float mse_coef,
                                     const float* __restrict__ dpm, int ld, long long rows, float* __restrict__ dy) {
  pdl_prologue();
  for (long long r = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; r < rows;
       r += static_cast<long long>(gridDim.x) * blockDim.x) {
    const float4 a = reinterpret_cast<const float4*>(x)[r];
    const float4 b = reinterpret_cast<const float4*>(xhat)[r];
    float4 o = make_float4(mse_coef * (b.x - a.x), mse_coef * (b.y - a.y), mse_coef * (b.z - a.z), mse_coef * (b.w - a.w));
    if (dpm) {
      const float4 q = *reinterpret_cast<const float4*>(dpm + r * ld);
      o.x += q.x; o.y += q.y; o.z += q.z; o.w += q.w;
    }
    reinterpret_cast<float4*>(dy)[r] = o;
  }
}
// dc = coef * (a - other) * (a > 0)   (DFC tap without a BatchNorm behind it: c10)
__global__ void tap_grad_relu_f32_kernel(const float* __restrict__ a, const float* __restrict__ other, float coef, long long n,
                                         float* __restrict__ dc) {
  pdl_prologue();
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const float v = a[i];
    dc[i] = v > 0.f ? coef * (v - other[i]) : 0.f;
  }
}
// dx = dy * act'(y) given the activation OUTPUT y (LeakyReLU: y > 0 ? 1 : alpha; ReLU: y > 0 ? 1 : 0)
__global__ void act_bwd_f32_kernel(const float* __restrict__ dy, const float* __restrict__ y, int act, float alpha, long long n,
                                   float* __restrict__ dx) {
  pdl_prologue();
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const float g = act == ICSG3D_ACT_NONE ? 1.f : (y[i] > 0.f ? 1.f : (act == ICSG3D_ACT_LEAKY ? alpha : 0.f));
    dx[i] = dy[i] * g;
  }
}

static int grid1(long long total) {
  long long b = (total + 255) / 256;
  if (b > 148 * 8) b = 148 * 8;
  if (b < 1) b = 1;
  return static_cast<int>(b);
}

}  // namespace icsg3d

using namespace icsg3d;

extern "C" int icsg3d_f32_to_split3(const float* src, int ld_src, int c, int64_t rows, void* dst, int ctot, int coff,
                                    int fmt, void* stream) {
  ICSG_REQUIRE(src && dst && c > 0 && c <= ld_src && coff >= 0 && coff + c <= ctot, "f32_to_split3: bad arguments");
  launch_k(f32_to_split3_kernel, grid1(rows * c), 256, 0, static_cast<cudaStream_t>(stream), src, ld_src, c, rows,
                                                                                       static_cast<__nv_bfloat16*>(dst), ctot, coff, fmt);
  ICSG_CHECK_LAUNCH();
  return ICSG3D_OK;
}

static int pack_vae_input_split3_impl(const float* m, const float* cond, int ncond, int B, int64_t vox, void* xe, void* xp,
                                      void* xp16, int fmt, void* stream, int lean = 0) {
  ICSG_REQUIRE(!lean || (fmt == 0 && ncond <= 10), "pack_vae_input: the lean encoder layout needs bf16 pairs and ncond <= 10");
  ICSG_REQUIRE(m && (xe || xp || xp16), "pack_vae_input_split3: null pointer");
  ICSG_REQUIRE(!xe || (cond && ncond >= 0 && ncond <= 12), "pack_vae_input_split3: ncond must be <= 12");
  const long long total = static_cast<long long>(B) * vox;
  launch_k(pack_vae_input_split3_kernel, grid1(total), 256, 0, static_cast<cudaStream_t>(stream), 
      m, cond, ncond, vox, total, static_cast<__nv_bfloat16*>(xe), static_cast<__nv_bfloat16*>(xp),
      static_cast<__nv_bfloat16*>(xp16), fmt, lean);
  ICSG_CHECK_LAUNCH();
  return ICSG3D_OK;
}

extern "C" int icsg3d_pack_vae_input_lean(const float* m, const float* cond, int ncond, int B, int64_t vox, void* xe32,
                                          void* xp16, void* stream) {
  return pack_vae_input_split3_impl(m, cond, ncond, B, vox, xe32, nullptr, xp16, 0, stream, 1);
}

extern "C" int icsg3d_pack_vae_input_split3(const float* m, const float* cond, int ncond, int B, int64_t vox, void* xe,
                                            void* xp, int fmt, void* stream) {
  return pack_vae_input_split3_impl(m, cond, ncond, B, vox, xe, xp, nullptr, fmt, stream);
}

extern "C" int icsg3d_pack_vae_input_mixed(const float* m, const float* cond, int ncond, int B, int64_t vox, void* xe3,
                                           void* xp16, int fmt, void* stream) {
  return pack_vae_input_split3_impl(m, cond, ncond, B, vox, xe3, nullptr, xp16, fmt, stream);
}

extern "C" int icsg3d_pack_conv_w_fprop_x3(const float* w, void* wpack, int ntaps, int cin, int cout, int cin_pad, int cout_pad,
                                           int cin_lead, int fold, int fold_c, int fmt, float wscale, void* stream) {
  ICSG_REQUIRE(w && wpack && (ntaps == 27 || ntaps == 1) && (fmt == 0 || fmt == 1), "pack_conv_w_fprop_x3: bad arguments");
  ICSG_REQUIRE(cout_pad >= cout && cin_pad > 0, "pack_conv_w_fprop_x3: bad padding");
  if (fold > 1) {
    ICSG_REQUIRE(cin == cin_lead + fold * fold_c && cin_pad >= cin_lead + fold_c, "pack_conv_w_fprop_x3: bad fold");
  } else {
    ICSG_REQUIRE(cin_pad >= cin, "pack_conv_w_fprop_x3: cin_pad < cin");
  }
  launch_k(pack_w_fprop_x3_kernel, grid1(static_cast<long long>(ntaps) * cout_pad * cin_pad), 256, 0, static_cast<cudaStream_t>(stream), 
      w, static_cast<__nv_bfloat16*>(wpack), ntaps, cin, cout, cin_pad, cout_pad, cin_lead, fold, fold_c, fmt, wscale);
  ICSG_CHECK_LAUNCH();
  return ICSG3D_OK;
}

extern "C" int icsg3d_pack_conv_w_dgrad_x3(const float* w, void* wpack, int ntaps, int cin, int cout, int cin_pad, int cout_pad,
                                           int fmt, float wscale, void* stream) {
  ICSG_REQUIRE(w && wpack && (ntaps == 27 || ntaps == 1) && (fmt == 0 || fmt == 1), "pack_conv_w_dgrad_x3: bad arguments");
  ICSG_REQUIRE(cin_pad >= cin && cout_pad >= cout, "pack_conv_w_dgrad_x3: bad padding");
  launch_k(pack_w_dgrad_x3_kernel, grid1(static_cast<long long>(ntaps) * cin_pad * cout_pad), 256, 0, static_cast<cudaStream_t>(stream), 
      w, static_cast<__nv_bfloat16*>(wpack), ntaps, cin, cout, cin_pad, cout_pad, fmt, wscale);
  ICSG_CHECK_LAUNCH();
  return ICSG3D_OK;
}

extern "C" int icsg3d_wgrad_combine_x3(const float* P, float* dw, int ntaps, int cin_pad, int cout_pad, void* stream) {
  ICSG_REQUIRE(P && dw && ntaps > 0 && cin_pad > 0 && cout_pad > 0, "wgrad_combine_x3: bad arguments");
  launch_k(wgrad_combine_x3_kernel, grid1(static_cast<long long>(ntaps) * cin_pad * cout_pad), 256, 0, static_cast<cudaStream_t>(stream), 
      P, dw, ntaps, cin_pad, cout_pad);
  ICSG_CHECK_LAUNCH();
  return ICSG3D_OK;
}

extern "C" int icsg3d_xhat_grad_f32(const float* x, const float* xhat, float mse_coef, const float* dpm, int ld, int64_t rows,
                                    float* dy, void* stream) {
  ICSG_REQUIRE(x && xhat && dy && (!dpm || ld % 4 == 0), "xhat_grad_f32: bad arguments");
  launch_k(xhat_grad_f32_kernel, grid1(rows), 256, 0, static_cast<cudaStream_t>(stream), x, xhat, mse_coef, dpm, ld, rows, dy);
  ICSG_CHECK_LAUNCH();
  return ICSG3D_OK;
}

extern "C" int icsg3d_tap_grad_relu_f32(const float* a, const float* other, float coef, int64_t n, float* dc, void* stream) {
  ICSG_REQUIRE(a && other && dc, "tap_grad_relu_f32: null pointer");
  launch_k(tap_grad_relu_f32_kernel, grid1(n), 256, 0, static_cast<cudaStream_t>(stream), a, other, coef, n, dc);
  ICSG_CHECK_LAUNCH();
  return ICSG3D_OK;
}

extern "C" int icsg3d_act_bwd_f32(const float* dy, const float* y, int act, float alpha, int64_t n, float* dx, void* stream) {
  ICSG_REQUIRE(dy && y && dx, "act_bwd_f32: null pointer");
  launch_k(act_bwd_f32_kernel, grid1(n), 256, 0, static_cast<cudaStream_t>(stream), dy, y, act, alpha, n, dx);
  ICSG_CHECK_LAUNCH();
  return ICSG3D_OK;
}
