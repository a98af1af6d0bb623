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

// VAE encoder / perceptual inputs in split form: xe3 [rows][48] from (M 4ch, cond one-hot ncond, 0..), xp3 [rows][48] from M
__global__ void pack_vae_input_split3_kernel(const float* __restrict__ m, const float* __restrict__ cond, int ncond,
                                             long long vox, long long total, __nv_bfloat16* __restrict__ xe,
                                             __nv_bfloat16* __restrict__ xp, int fmt) {
  for (long long r = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; r < total;
       r += static_cast<long long>(gridDim.x) * blockDim.x) {
    float v[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = 0.f;
    const float4 q = *reinterpret_cast<const float4*>(m + r * 4);
    v[0] = q.x; v[1] = q.y; v[2] = q.z; v[3] = q.w;
    if (xp) {
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        __nv_bfloat16 hi, lo;
        split_bf16(v[i], hi, lo, fmt);
        xp[r * 48 + i] = hi;
        xp[r * 48 + 16 + i] = lo;
        xp[r * 48 + 32 + i] = hi;
      }
    }
    if (xe) {
      const float* c = cond + (r / vox) * ncond;
      for (int i = 0; i < ncond && i < 12; ++i) v[4 + i] = c[i];
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        __nv_bfloat16 hi, lo;
        split_bf16(v[i], hi, lo, fmt);
        xe[r * 48 + i] = hi;
        xe[r * 48 + 16 + i] = lo;
        xe[r * 48 + 32 + i] = hi;
      }
    }
  }
}

// w fp32 (ntaps, Cin, Cout) Keras layout -> bf16 [ntaps][cout_pad][3*cin_pad] = [w_hi | w_hi | w_lo] along Cin, with the
// same channel padding / condition fold as pack_w_fprop_kernel.
__global__ void pack_w_fprop_x3_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ wp, int ntaps, int cin, int cout,
                                       int cin_pad, int cout_pad, int cin_lead, int fold, int fold_c, int fmt, float wscale) {
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
  f32_to_split3_kernel<<<grid1(rows * c), 256, 0, static_cast<cudaStream_t>(stream)>>>(src, ld_src, c, rows,
                                                                                       static_cast<__nv_bfloat16*>(dst), ctot, coff, fmt);
  ICSG_CHECK_LAUNCH();
  return ICSG3D_OK;
}

extern "C" int icsg3d_pack_vae_input_split3(const float* m, const float* cond, int ncond, int B, int64_t vox, void* xe,
                                            void* xp, int fmt, void* stream) {
  ICSG_REQUIRE(m && (xe || xp), "pack_vae_input_split3: null pointer");
  ICSG_REQUIRE(!xe || (cond && ncond >= 0 && ncond <= 12), "pack_vae_input_split3: ncond must be <= 12");
  const long long total = static_cast<long long>(B) * vox;
  pack_vae_input_split3_kernel<<<grid1(total), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      m, cond, ncond, vox, total, static_cast<__nv_bfloat16*>(xe), static_cast<__nv_bfloat16*>(xp), fmt);
  ICSG_CHECK_LAUNCH();
  return ICSG3D_OK;
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
  pack_w_fprop_x3_kernel<<<grid1(static_cast<long long>(ntaps) * cout_pad * cin_pad), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      w, static_cast<__nv_bfloat16*>(wpack), ntaps, cin, cout, cin_pad, cout_pad, cin_lead, fold, fold_c, fmt, wscale);
  ICSG_CHECK_LAUNCH();
  return ICSG3D_OK;
}
