// Weight packing between the Keras master layout (kd,kh,kw,Cin,Cout) fp32 and the bf16 GEMM
// operand layouts of the tcgen05 conv kernels, incl. channel padding to multiples of 16 and the
// fold of the encoder's tiled condition channels (vae/lattice_vae.py:167-169).
#include "common.cuh"

namespace icsg3d {

__global__ void pack_w_fprop_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ wp, int cin, int cout,
                                    int cin_pad, int cout_pad, int cin_lead, int fold, int fold_c) {
  pdl_prologue();
  const long long total = 27ll * cout_pad * cin_pad;
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
    wp[idx] = f2bf(v);
  }
}

__global__ void pack_w_dgrad_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ wp, int cin, int cout,
                                    int cin_pad, int cout_pad) {
  pdl_prologue();
  const long long total = 27ll * cin_pad * cout_pad;
  for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int co = static_cast<int>(idx % cout_pad);
    const int ci = static_cast<int>((idx / cout_pad) % cin_pad);
    const int tap = static_cast<int>(idx / (static_cast<long long>(cin_pad) * cout_pad));
    float v = 0.f;
    if (ci < cin && co < cout) v = w[(static_cast<long long>(26 - tap) * cin + ci) * cout + co];
    wp[idx] = f2bf(v);
  }
}

// Fast forms for the wide layers (no padding, no fold): the U-Net repacks 30 M weights per train step.
// fprop is a per-tap transpose [ci][co] -> [co][ci]: 64 x 32 tile through shared memory, 128-byte rows on both sides.
__global__ void __launch_bounds__(256) pack_w_fprop_tiled_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ wp,
                                                                 int cin, int cout) {
  pdl_prologue();
  __shared__ float tile[64][33];
  const int ci0 = blockIdx.x * 64, co0 = blockIdx.y * 32, tap = blockIdx.z;
  const int lane = threadIdx.x & 31, r0 = threadIdx.x >> 5;
  const float* src = w + (static_cast<long long>(tap) * cin + ci0) * cout + co0 + lane;
#pragma unroll
  for (int k = 0; k < 8; ++k) tile[r0 + 8 * k][lane] = src[static_cast<long long>(r0 + 8 * k) * cout];
  __syncthreads();
  uint32_t* dst = reinterpret_cast<uint32_t*>(wp + (static_cast<long long>(tap) * cout + co0) * cin + ci0) + lane;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int c = r0 + 8 * k;
    dst[static_cast<long long>(c) * (cin >> 1)] = pack_bf16x2(tile[2 * lane][c], tile[2 * lane + 1][c]);
  }
}

// dgrad keeps the orientation (taps mirrored): 8 output channels per thread, 16-byte stores; the input channels may be a
// slice [ci0, ci0 + cin) of a kernel with cin_total input channels (the skip half of a folded Upsample + Conv layer)
__global__ void __launch_bounds__(256) pack_w_dgrad_vec_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ wp,
                                                               int cin, int cout, int cin_total, int ci0) {
  pdl_prologue();
  const int c8 = cout >> 3;
  const long long per_tap = static_cast<long long>(cin) * c8, total = 27 * per_tap;
  for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int tap = static_cast<int>(idx / per_tap);
    const long long r = idx - tap * per_tap;  // (ci, co8) of this tap
    const float4* src = reinterpret_cast<const float4*>(
        w + ((static_cast<long long>(26 - tap) * cin_total + ci0) * c8 + r) * 8);
    const float4 a = src[0], b = src[1];
    uint4 q;
    q.x = pack_bf16x2(a.x, a.y);
    q.y = pack_bf16x2(a.z, a.w);
    q.z = pack_bf16x2(b.x, b.y);
    q.w = pack_bf16x2(b.z, b.w);
    reinterpret_cast<uint4*>(wp)[idx] = q;
  }
}

__global__ void unpack_dw_kernel(const float* __restrict__ dwp, float* __restrict__ dw, int cin, int cout, int cin_pad,
                                 int cout_pad, int cin_lead, int fold, int fold_c) {
  pdl_prologue();
  const long long total = 27ll * cin * cout;
  for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int co = static_cast<int>(idx % cout);
    const int ci = static_cast<int>((idx / cout) % cin);
    const int tap = static_cast<int>(idx / (static_cast<long long>(cin) * cout));
    int src = ci;
    if (fold > 1 && ci >= cin_lead) src = cin_lead + (ci - cin_lead) % fold_c;
    dw[idx] = dwp[(static_cast<long long>(tap) * cin_pad + src) * cout_pad + co];
  }
}

// All weight packs of a model in ONE launch: jobs[j] = {w, wp, cin, cout, cin_pad, cout_pad, cin_lead, fold, fold_c, mode}
// (int64 each; mode 0 = fprop layout, 1 = dgrad layout, 2 / 3 = the same as bf16-pair split operands); blockIdx.y = job.
__global__ void pack_w_batch_kernel(const long long* __restrict__ jobs) {
  pdl_prologue();
  const long long* jb = jobs + static_cast<size_t>(blockIdx.y) * 10;
  const float* w = reinterpret_cast<const float*>(jb[0]);
  __nv_bfloat16* wp = reinterpret_cast<__nv_bfloat16*>(jb[1]);
  const int cin = static_cast<int>(jb[2]), cout = static_cast<int>(jb[3]);
  const int cin_pad = static_cast<int>(jb[4]), cout_pad = static_cast<int>(jb[5]);
  const int cin_lead = static_cast<int>(jb[6]), fold = static_cast<int>(jb[7]), fold_c = static_cast<int>(jb[8]);
  const int mode = static_cast<int>(jb[9]);
  const long long total = 27ll * cout_pad * (mode == 4 ? 32 : cin_pad);
  for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    float v = 0.f;
    if (mode == 4) {
      // lean split pack of the encoder's first conv (split3.cu::store_lean_enc_row): [27][cout_pad][32]; cin = cin_lead +
      // fold * fold_c real channels (4 + 4 x 10), K = [w_hi(M) | w_hi(cond) | w_hi(M 0:2)] [w_hi(M 2:4) | w_lo(M) | w_lo(cond)]
      const int k = static_cast<int>(idx % 32);
      const int co = static_cast<int>((idx / 32) % cout_pad);
      const int tap = static_cast<int>(idx / (32ll * cout_pad));
      int src = -1, want_lo = 0, folded = 0;  // src: M channel or folded cond index
      if (k < 4) src = k;
      else if (k < 14) { src = k - 4; folded = 1; }
      else if (k < 16) src = k - 14;
      else if (k < 18) src = k - 16 + 2;
      else if (k < 22) { src = k - 18; want_lo = 1; }
      else { src = k - 22; folded = 1; want_lo = 1; }
      if (co < cout) {
        const float* wt = w + static_cast<long long>(tap) * cin * cout;
        if (!folded) v = wt[static_cast<long long>(src) * cout + co];
        else if (src < fold_c)
          for (int r = 0; r < fold; ++r) v += wt[static_cast<long long>(cin_lead + r * fold_c + src) * cout + co];
      }
      const __nv_bfloat16 hi = f2bf(v);
      wp[idx] = want_lo ? f2bf(v - __bfloat162float(hi)) : hi;
      continue;
    }
    if (mode == 0 || mode == 2) {
      const int ci = static_cast<int>(idx % cin_pad);
      const int co = static_cast<int>((idx / cin_pad) % cout_pad);
      const int tap = static_cast<int>(idx / (static_cast<long long>(cin_pad) * cout_pad));
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
    } else {
      const int co = static_cast<int>(idx % cout_pad);
      const int ci = static_cast<int>((idx / cout_pad) % cin_pad);
      const int tap = static_cast<int>(idx / (static_cast<long long>(cin_pad) * cout_pad));
      if (ci < cin && co < cout) v = w[(static_cast<long long>(26 - tap) * cin + ci) * cout + co];
    }
    if (mode < 2) {
      wp[idx] = f2bf(v);
    } else {
      // fp32-class split operands as bf16 pairs (csrc/split3.cu): K parts [w_hi | w_hi | w_lo]; mode 2 = fprop layout
      // [27][cout_pad][3*cin_pad], mode 3 = dgrad layout [27][cin_pad][3*cout_pad]
      const __nv_bfloat16 hi = f2bf(v), lo = f2bf(v - __bfloat162float(hi));
      const int kpad = mode == 2 ? cin_pad : cout_pad;
      __nv_bfloat16* d = wp + (idx / kpad) * 3 * kpad + (idx % kpad);
      d[0] = hi;
      d[kpad] = hi;
      d[2 * kpad] = lo;
    }
  }
}

static int grid_for(long long total) {
  long long b = (total + 255) / 256;
  if (b > 148 * 8) b = 148 * 8;
  if (b < 1) b = 1;
  return static_cast<int>(b);
}

}  // namespace icsg3d

using namespace icsg3d;

extern "C" int icsg3d_pack_conv_w_fprop(const float* w, void* wpack, int cin, int cout, int cin_pad, int cout_pad,
                                        int cin_lead, int fold, int fold_c, void* stream) {
  ICSG_REQUIRE(w && wpack, "pack_conv_w_fprop: null pointer");
  ICSG_REQUIRE(cout_pad >= cout && cin_pad > 0, "pack_conv_w_fprop: bad padding");
  if (fold > 1) {
    ICSG_REQUIRE(cin == cin_lead + fold * fold_c && cin_pad >= cin_lead + fold_c,
                 "pack_conv_w_fprop: fold %d x %d + lead %d does not match cin %d / cin_pad %d", fold, fold_c, cin_lead,
                 cin, cin_pad);
  } else {
    ICSG_REQUIRE(cin_pad >= cin, "pack_conv_w_fprop: cin_pad < cin");
  }
  if (fold <= 1 && cin_pad == cin && cout_pad == cout && cin % 64 == 0 && cout % 32 == 0 && cout / 32 <= 65535 &&
      (reinterpret_cast<uintptr_t>(wpack) & 3) == 0) {
    launch_k(pack_w_fprop_tiled_kernel, dim3(cin / 64, cout / 32, 27), 256, 0, static_cast<cudaStream_t>(stream), w,
             static_cast<__nv_bfloat16*>(wpack), cin, cout);
    ICSG_CHECK_LAUNCH();
    return ICSG3D_OK;
  }
  launch_k(pack_w_fprop_kernel, grid_for(27ll * cout_pad * cin_pad), 256, 0, static_cast<cudaStream_t>(stream), 
      w, static_cast<__nv_bfloat16*>(wpack), cin, cout, cin_pad, cout_pad, cin_lead, fold, fold_c);
  ICSG_CHECK_LAUNCH();
  return ICSG3D_OK;
}

extern "C" int icsg3d_pack_conv_w_dgrad(const float* w, void* wpack, int cin, int cout, int cin_pad, int cout_pad,
                                        void* stream) {
  ICSG_REQUIRE(w && wpack, "pack_conv_w_dgrad: null pointer");
  ICSG_REQUIRE(cin_pad >= cin && cout_pad >= cout, "pack_conv_w_dgrad: bad padding");
  if (cin_pad == cin && cout_pad == cout && cout % 8 == 0 &&
      ((reinterpret_cast<uintptr_t>(wpack) | reinterpret_cast<uintptr_t>(w)) & 15) == 0) {
    launch_k(pack_w_dgrad_vec_kernel, grid_for(27ll * cin * (cout / 8)), 256, 0, static_cast<cudaStream_t>(stream), w,
             static_cast<__nv_bfloat16*>(wpack), cin, cout, cin, 0);
    ICSG_CHECK_LAUNCH();
    return ICSG3D_OK;
  }
  launch_k(pack_w_dgrad_kernel, grid_for(27ll * cout_pad * cin_pad), 256, 0, static_cast<cudaStream_t>(stream), 
      w, static_cast<__nv_bfloat16*>(wpack), cin, cout, cin_pad, cout_pad);
  ICSG_CHECK_LAUNCH();
  return ICSG3D_OK;
}

// dgrad operand of the input-channel slice [ci0, ci0 + cin) of a kernel fp32 [27][cin_total][cout]: bf16 [27][cin][cout]
extern "C" int icsg3d_pack_conv_w_dgrad_slice(const float* w, void* wpack, int cin_total, int ci0, int cin, int cout,
                                              void* stream) {
  ICSG_REQUIRE(w && wpack && ci0 >= 0 && cin > 0 && ci0 + cin <= cin_total && cout % 8 == 0 && cin % 16 == 0 && cout % 16 == 0,
               "pack_conv_w_dgrad_slice: bad channel slice");
  ICSG_REQUIRE(((reinterpret_cast<uintptr_t>(wpack) | reinterpret_cast<uintptr_t>(w)) & 15) == 0,
               "pack_conv_w_dgrad_slice: operands must be 16-byte aligned");
  launch_k(pack_w_dgrad_vec_kernel, grid_for(27ll * cin * (cout / 8)), 256, 0, static_cast<cudaStream_t>(stream), w,
           static_cast<__nv_bfloat16*>(wpack), cin, cout, cin_total, ci0);
  ICSG_CHECK_LAUNCH();
  return ICSG3D_OK;
}

extern "C" int icsg3d_unpack_conv_dw(const float* dw_pad, float* dw, int cin, int cout, int cin_pad, int cout_pad,
                                     int cin_lead, int fold, int fold_c, void* stream) {
  ICSG_REQUIRE(dw_pad && dw, "unpack_conv_dw: null pointer");
  launch_k(unpack_dw_kernel, grid_for(27ll * cin * cout), 256, 0, static_cast<cudaStream_t>(stream), 
      dw_pad, dw, cin, cout, cin_pad, cout_pad, cin_lead, fold, fold_c);
  ICSG_CHECK_LAUNCH();
  return ICSG3D_OK;
}


extern "C" int icsg3d_pack_conv_w_batch(const int64_t* jobs, int njobs, int max_blocks, void* stream) {
  ICSG_REQUIRE(jobs && njobs > 0 && njobs <= 65535 && max_blocks > 0, "pack_conv_w_batch: bad arguments");
  launch_k(pack_w_batch_kernel, dim3(max_blocks, njobs), 256, 0, static_cast<cudaStream_t>(stream), 
      reinterpret_cast<const long long*>(jobs));
  ICSG_CHECK_LAUNCH();
  return ICSG3D_OK;
}
