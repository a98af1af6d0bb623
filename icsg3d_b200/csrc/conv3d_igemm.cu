// Conv3D 3x3x3 "same" as an implicit GEMM on tcgen05 / TMEM, operands fetched by TMA (sm_100a).
//
// Replaces the TF ops behind Keras Conv3D(kernel_size=3, padding="same") in the reference graphs
// (vae/lattice_vae.py:173,178,213,219-224; unet/unet.py:276-336): Conv3D + BiasAdd (+ the ReLU that
// follows it in the U-Net ordering), and Conv3DBackpropInputV2 when called with mirrored weights.
//
// GEMM view:  Y[M = B*D*H*W, N = Cout] = sum over 27 taps of  X_tap[M, Cin] * Wp[tap][N, Cin]^T
//   * one CTA tile = 128 consecutive output voxels (a box bw x bh x bd x bn of the NDHWC tensor) x NT
//     output channels; accumulator 128 lanes x NT fp32 columns in TMEM, double buffered so the epilogue
//     of tile i overlaps the MMAs of tile i+1;
//   * the A operand of tap (kd,kh,kw) is the same box shifted by (kd-1,kh-1,kw-1): one 5-D tiled TMA
//     load whose out-of-bounds zero fill IS the "same" padding (also across sample boundaries);
//   * both operands are K-major with the hardware swizzle matching the channel-chunk width
//     (KC = 16/32/64 channels -> 32/64/128-byte swizzle), so TMA output == UMMA canonical layout;
//   * warp roles: warp 0 = TMA producer, warp 1 = MMA issuer (one elected lane) + TMEM allocator,
//     warps 2..5 = epilogue (TMEM -> registers -> bias/activation -> global);
//   * persistent CTAs, static round-robin tile scheduler.
#include "common.cuh"

namespace icsg3d {

struct ConvIgemmParams {
  int m_total;          // B*D*H*W
  int D, H, W;          // spatial extents
  int bw, bh, bd, bn;   // TMA box (product == 128)
  int chunks;           // cin / KC
  int kc;               // channels per k-unit: 16, 32 or 64
  int nt;               // N tile
  int tiles_n, tiles_m;
  int mt;               // M tiles (128 voxels each) that share one weight tile per stage: 1, or 2 for the wide layers
                        // (operand traffic per FLOP 3/4: the per-tap kernel is bound by L2 -> shared-memory traffic)
  int ups;              // k-units per pipeline stage
  int ntaps;            // 27 (3x3x3) or 1 (1x1x1)
  int iters;            // ntaps*chunks/ups
  int ksplit, ips;      // split-K: number of K slices and pipeline iterations per slice (ksplit == 1: no split)
  float* ws;            // split-K partials fp32 [ksplit][m_total][tiles_n*nt]
  int stages;
  uint32_t a_unit_bytes;  // 128*kc*2
  uint32_t b_unit_bytes;  // nt*kc*2
  uint32_t stage_bytes;   // multiple of 1024
  uint32_t sbo;           // 8 rows * kc*2 bytes
  uint32_t layout;        // UMMA swizzle code
  uint32_t idesc;
  uint32_t tmem_cols;     // power of two >= 32, >= 2*nt
  // epilogue
  void* y;
  int ldy;
  int y_dtype;
  int n_store;
  const float* bias;
  int act;
  float alpha;
  float oscale;  // output = accumulator * oscale + bias (1 unless the weights were pre-scaled: fp16 split mode)
  const float* post_scale;  // optional per-channel affine applied AFTER the activation (inference BatchNorm that follows
  const float* post_shift;  // Conv3D+ReLU in the U-Net, unet.py:277-279): y = post_scale * act(.) + post_shift
};

static constexpr int kConvThreads = 192;
static constexpr int kMaxStages = 8;

template <int KSTEPS, int MT>
__global__ void __launch_bounds__(kConvThreads, 1)
conv3d_k3_igemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                       const ConvIgemmParams p) {
  pdl_prologue();
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t full_bar[kMaxStages];
  __shared__ __align__(8) uint64_t empty_bar[kMaxStages];
  __shared__ __align__(8) uint64_t tmem_full_bar[2];
  __shared__ __align__(8) uint64_t tmem_empty_bar[2];
  __shared__ uint32_t tmem_base_slot;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  // 1024-byte aligned operand ring (128B swizzle atoms repeat every 1024 bytes).
  const uint32_t ring_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* ring = smem_raw + (ring_base - smem_u32(smem_raw));

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tmem_full_bar[a], 1);
      mbar_init(&tmem_empty_bar[a], 4);
    }
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc(&tmem_base_slot, p.tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_slot;

  const int tiles_mg = (p.tiles_m + MT - 1) / MT;        // groups of mt M tiles
  const int total_tiles = tiles_mg * p.tiles_n * p.ksplit;   // tile index = (tm * tiles_n + tn) * ksplit + ks

  if (warp == 0) {
    // ===================== TMA producer (whole warp runs the loop; one elected lane issues) =====================
    const bool leader = elect_one();
    int stage = 0;
    uint32_t phase = 0;
    const uint32_t tx_bytes = static_cast<uint32_t>(p.ups) * (static_cast<uint32_t>(MT) * p.a_unit_bytes + p.b_unit_bytes);
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      const int ks = tile % p.ksplit;
      const int tmn = tile / p.ksplit;
      const int tm = tmn / p.tiles_n;
      const int tn = tmn - tm * p.tiles_n;
      int w0[MT], h0[MT], d0[MT], n0[MT];
#pragma unroll
      for (int t = 0; t < MT; ++t) {  // a tile past the end (odd tile count) lies beyond the last sample: zero filled
        int pix = (tm * MT + t) * 128;
        w0[t] = pix % p.W;
        pix /= p.W;
        h0[t] = pix % p.H;
        pix /= p.H;
        d0[t] = pix % p.D;
        n0[t] = pix / p.D;
      }
      const int it0 = ks * p.ips, it1 = min(p.iters, it0 + p.ips);
      const int unit0 = it0 * p.ups;
      int tap = unit0 / p.chunks, ch = unit0 - tap * p.chunks;
      for (int it = it0; it < it1; ++it) {
        mbar_wait(&empty_bar[stage], phase ^ 1u);
        uint8_t* sa = ring + static_cast<size_t>(stage) * p.stage_bytes;
        uint8_t* sb = sa + static_cast<size_t>(p.ups * MT) * p.a_unit_bytes;
        if (leader) mbar_expect_tx(&full_bar[stage], tx_bytes);
        for (int j = 0; j < p.ups; ++j) {
          int kd = tap / 9;
          int kh = (tap - kd * 9) / 3;
          int kw = tap - kd * 9 - kh * 3;
          if (p.ntaps == 1) kd = kh = kw = 1;  // 1x1x1: no spatial shift
          if (leader) {
#pragma unroll
            for (int t = 0; t < MT; ++t)
              tma_load_5d(sa + static_cast<size_t>(j * MT + t) * p.a_unit_bytes, &tmA, &full_bar[stage], ch * p.kc,
                          w0[t] + kw - 1, h0[t] + kh - 1, d0[t] + kd - 1, n0[t]);
            tma_load_3d(sb + static_cast<size_t>(j) * p.b_unit_bytes, &tmB, &full_bar[stage], ch * p.kc, tn * p.nt, tap);
          }
          if (++ch == p.chunks) {
            ch = 0;
            ++tap;
          }
        }
        if (++stage == p.stages) {
          stage = 0;
          phase ^= 1u;
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (whole warp runs the loop; one elected lane issues) =====================
    const bool leader = elect_one();
    int stage = 0;
    uint32_t phase = 0;
    int local = 0;
    const uint32_t desc_hi = umma_desc_hi(p.sbo, p.layout);
    const uint32_t ring_lo = umma_desc_lo(ring_base, 16u);
    const uint32_t stage_lo = p.stage_bytes >> 4, a_unit_lo = p.a_unit_bytes >> 4, b_unit_lo = p.b_unit_bytes >> 4;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++local) {
      const int acc = local & 1;
      const uint32_t acc_phase = static_cast<uint32_t>(local >> 1) & 1u;
      mbar_wait(&tmem_empty_bar[acc], acc_phase ^ 1u);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(acc * MT * p.nt);
      const int ks = tile % p.ksplit;
      const int it0 = ks * p.ips, it1 = min(p.iters, it0 + p.ips);
      for (int it = it0; it < it1; ++it) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        uint32_t a_lo = ring_lo + static_cast<uint32_t>(stage) * stage_lo;
        uint32_t b_lo = a_lo + static_cast<uint32_t>(p.ups * MT) * a_unit_lo;
        for (int j = 0; j < p.ups; ++j) {
          if (leader) {
#pragma unroll
            for (int t = 0; t < MT; ++t) {
#pragma unroll
              for (int k = 0; k < KSTEPS; ++k)
                umma_bf16_lohi(d_tmem + static_cast<uint32_t>(t * p.nt), a_lo + static_cast<uint32_t>(t) * a_unit_lo + 2u * k, desc_hi,
                               b_lo + 2u * k, desc_hi, p.idesc, ((it - it0) | j | k) != 0 ? 1u : 0u);
            }
          }
          a_lo += static_cast<uint32_t>(MT) * a_unit_lo;
          b_lo += b_unit_lo;
        }
        if (leader) umma_commit(&empty_bar[stage]);  // frees the smem slot once these MMAs retire
        if (++stage == p.stages) {
          stage = 0;
          phase ^= 1u;
        }
      }
      if (leader) umma_commit(&tmem_full_bar[acc]);  // accumulator complete -> epilogue
    }
  } else {
    // ===================== epilogue warps (2..5) =====================
    const int quarter = warp & 3;  // TMEM lane quarter this warp may access
    const int row = quarter * 32 + lane;
    int local = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++local) {
      const int acc = local & 1;
      const uint32_t acc_phase = static_cast<uint32_t>(local >> 1) & 1u;
      const int ks = tile % p.ksplit;
      const int tmn = tile / p.ksplit;
      const int tm = tmn / p.tiles_n;
      const int tn = tmn - tm * p.tiles_n;
      mbar_wait(&tmem_full_bar[acc], acc_phase);
      tc_fence_after();
      for (int t = 0; t < MT; ++t) {
      const long long pixel = (static_cast<long long>(tm) * MT + t) * 128 + row;
      const bool row_ok = pixel < p.m_total;
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + static_cast<uint32_t>((acc * MT + t) * p.nt);
      for (int c0 = 0; c0 < p.nt; c0 += 16) {
        uint32_t v[16];
        tmem_ld16(taddr + static_cast<uint32_t>(c0), v);
        tmem_ld_wait();
        const int col0 = tn * p.nt + c0;
        if (p.ksplit > 1) {  // raw fp32 partial of this K slice; bias/activation/conversion happen in the reduce kernel
          if (row_ok) {
            float4* dst = reinterpret_cast<float4*>(p.ws + (static_cast<size_t>(ks) * p.m_total + pixel) * (p.tiles_n * p.nt) + col0);
#pragma unroll
            for (int i = 0; i < 4; ++i)
              dst[i] = make_float4(__uint_as_float(v[4 * i]), __uint_as_float(v[4 * i + 1]), __uint_as_float(v[4 * i + 2]),
                                   __uint_as_float(v[4 * i + 3]));
          }
          continue;
        }
        if (!row_ok || col0 >= p.n_store) continue;
        float f[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          float x = __uint_as_float(v[i]) * p.oscale;
          if (p.bias != nullptr) x += __ldg(p.bias + col0 + i);
          if (p.act == ICSG3D_ACT_RELU) {
            x = fmaxf(x, 0.f);
          } else if (p.act == ICSG3D_ACT_LEAKY) {
            x = x > 0.f ? x : p.alpha * x;
          }
          if (p.post_scale != nullptr) x = fmaf(x, __ldg(p.post_scale + col0 + i), __ldg(p.post_shift + col0 + i));
          f[i] = x;
        }
        const int nvalid = min(16, p.n_store - col0);
        if (p.y_dtype == ICSG3D_DT_BF16) {
          __nv_bfloat16* dst = reinterpret_cast<__nv_bfloat16*>(p.y) + pixel * p.ldy + col0;
          if (nvalid == 16 && (p.ldy & 7) == 0) {
            uint4 q0, q1;
            q0.x = pack_bf16x2(f[0], f[1]);
            q0.y = pack_bf16x2(f[2], f[3]);
            q0.z = pack_bf16x2(f[4], f[5]);
            q0.w = pack_bf16x2(f[6], f[7]);
            q1.x = pack_bf16x2(f[8], f[9]);
            q1.y = pack_bf16x2(f[10], f[11]);
            q1.z = pack_bf16x2(f[12], f[13]);
            q1.w = pack_bf16x2(f[14], f[15]);
            reinterpret_cast<uint4*>(dst)[0] = q0;
            reinterpret_cast<uint4*>(dst)[1] = q1;
          } else {
            for (int i = 0; i < nvalid; ++i) dst[i] = f2bf(f[i]);
          }
        } else {
          float* dst = reinterpret_cast<float*>(p.y) + pixel * p.ldy + col0;
          if (nvalid == 16 && (p.ldy & 3) == 0) {
#pragma unroll
            for (int i = 0; i < 4; ++i)
              reinterpret_cast<float4*>(dst)[i] = make_float4(f[4 * i], f[4 * i + 1], f[4 * i + 2], f[4 * i + 3]);
          } else if (nvalid == 4 && (p.ldy & 3) == 0) {
            reinterpret_cast<float4*>(dst)[0] = make_float4(f[0], f[1], f[2], f[3]);
          } else {
            for (int i = 0; i < nvalid; ++i) dst[i] = f[i];
          }
        }
      }
      }  // t
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty_bar[acc]);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, p.tmem_cols);
  }
}

// y[pixel][c] = act(sum over K slices of ws[ks][pixel][c] + bias[c]), fixed slice order (deterministic)
__global__ void __launch_bounds__(256) conv_splitk_reduce_kernel(const float* __restrict__ ws, int ksplit, long long m_total,
                                                                 int ncols, const float* __restrict__ bias, int act, float alpha,
                                                                 void* __restrict__ y, int ldy, int y_dtype, int n_store,
                                                                 float oscale, const float* __restrict__ post_scale,
                                                                 const float* __restrict__ post_shift) {
  pdl_prologue();
  const int c4 = ncols >> 2;
  const long long total = m_total * c4;
  const size_t slice = static_cast<size_t>(m_total) * ncols;
  for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long pixel = idx / c4;
    const int c = static_cast<int>(idx - pixel * c4) << 2;
    if (c >= n_store) continue;
    const float* src = ws + static_cast<size_t>(pixel) * ncols + c;
    float4 a = *reinterpret_cast<const float4*>(src);
    for (int k = 1; k < ksplit; ++k) {
      const float4 b = *reinterpret_cast<const float4*>(src + k * slice);
      a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
    }
    float o[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float x = o[i] * oscale + (bias ? bias[c + i] : 0.f);
      if (act == ICSG3D_ACT_RELU) x = fmaxf(x, 0.f);
      else if (act == ICSG3D_ACT_LEAKY) x = x > 0.f ? x : alpha * x;
      if (post_scale != nullptr) x = fmaf(x, post_scale[c + i], post_shift[c + i]);
      o[i] = x;
    }
    const int nv = min(4, n_store - c);
    if (y_dtype == ICSG3D_DT_BF16) {
      __nv_bfloat16* dst = reinterpret_cast<__nv_bfloat16*>(y) + pixel * ldy + c;
      if (nv == 4 && (ldy & 3) == 0) {
        uint2 q;
        q.x = pack_bf16x2(o[0], o[1]);
        q.y = pack_bf16x2(o[2], o[3]);
        *reinterpret_cast<uint2*>(dst) = q;
      } else {
        for (int i = 0; i < nv; ++i) dst[i] = f2bf(o[i]);
      }
    } else {
      float* dst = reinterpret_cast<float*>(y) + pixel * ldy + c;
      if (nv == 4 && (ldy & 3) == 0) *reinterpret_cast<float4*>(dst) = make_float4(o[0], o[1], o[2], o[3]);
      else
        for (int i = 0; i < nv; ++i) dst[i] = o[i];
    }
  }
}

static bool is_pow2(int v) { return v > 0 && (v & (v - 1)) == 0; }

// Decompose a 128-voxel tile into a TMA box over (W, H, D, B).
static void tile_box(int B, int D, int H, int W, int* bw, int* bh, int* bd, int* bn) {
  int rem = 128;
  *bw = W < rem ? W : rem;
  rem /= *bw;
  *bh = H < rem ? H : rem;
  rem /= *bh;
  *bd = D < rem ? D : rem;
  rem /= *bd;
  *bn = rem;  // may exceed B for tiny problems: out-of-bounds samples are zero filled and never stored
  (void)B;
}

int encode_act_map(CUtensorMap* map, const void* x, int ldx, int B, int D, int H, int W, int c_extent, int kc) {
  int bw, bh, bd, bn;
  tile_box(B, D, H, W, &bw, &bh, &bd, &bn);
  uint64_t dims[5] = {static_cast<uint64_t>(c_extent), static_cast<uint64_t>(W), static_cast<uint64_t>(H),
                      static_cast<uint64_t>(D), static_cast<uint64_t>(B)};
  uint64_t strides[4] = {static_cast<uint64_t>(ldx) * 2, static_cast<uint64_t>(W) * ldx * 2,
                         static_cast<uint64_t>(H) * W * ldx * 2, static_cast<uint64_t>(D) * H * W * ldx * 2};
  uint32_t box[5] = {static_cast<uint32_t>(kc), static_cast<uint32_t>(bw), static_cast<uint32_t>(bh),
                     static_cast<uint32_t>(bd), static_cast<uint32_t>(bn)};
  return encode_tiled_bf16(map, x, 5, dims, strides, box, kc * 2);
}

}  // namespace icsg3d

#include <stdlib.h>
#include <string.h>

#include "conv3d_halo.cuh"
#include "conv3d_stream.cuh"
namespace icsg3d {
// ICSG3D_CONV_IMPL=v1 forces the per-tap TMA kernel, =halo the halo kernel (A/B comparisons); the default picks the
// plane-streaming kd-folded kernel when it applies, then the halo kernel, then the per-tap kernel.
extern int g_wgrad_impl;  // conv3d_wgrad.cu
static int g_conv_impl = -1;
static int conv_impl_choice() {
  int& v = g_conv_impl;
  if (v < 0) {
    const char* e = getenv("ICSG3D_CONV_IMPL");
    v = (e && strcmp(e, "v1") == 0) ? 1 : (e && strcmp(e, "halo") == 0) ? 2 : 0;
  }
  return v;
}
// Halo kernel or per-tap kernel?  conv_halo_plan() only says whether a halo configuration FITS.  Measured on B200
// (profiles/r02_conv_halo_vs_pertap.txt): whenever the halo plan has to split N (nt < nout: two passes over the same
// halo block at the 55-cycle MMA floor, 3-4 weight stages) the per-tap kernel is 1.9-2.5x faster (c18 128->128 @32^3:
// 453 -> 240 us at batch 8; c16 256->128 @16^3: 204 -> 80 us), and a narrow layer that leaves SMs idle (fewer items than
// SMs at nt <= 64: dec_conv2 128->64 @8^3) is 1.5x faster per tap with split K.  ICSG3D_CONV_IMPL=halo forces halo.
static bool conv_use_halo(int B, int D, int H, int W, int cin, int nout, int sms, ConvHaloParams* hp) {
  if (conv_impl_choice() == 1 || !conv_halo_plan(B, D, H, W, cin, nout, sms, hp)) return false;
  if (conv_impl_choice() == 2) return true;
  if (hp->tiles_n > 1) return false;
  if (hp->total_items < sms && hp->nt <= 64) return false;
  // two or more channel chunks (cin >= 128) shrink the halo block to one 50 %-useful tile per item; with a wide N
  // (>= 128: the per-tap MMAs are off the 55-cycle floor) and enough voxels to amortise the per-tap pipeline
  // (>= 256 M tiles) the per-tap kernel wins by 1.2-1.5x (128->256 @16^3 B=8: 90 -> 75 us; 256->128 @8^3 B=100: 190 -> 124;
  // 128->128 @16^3 B=16: 128 -> 90), while it loses below that size (128->128 @8^3 B=32: 30 vs 45) and for narrow N
  // (128->64 @16^3 B=100: 237 vs 355)
  if (cin >= 128 && nout >= 128 && static_cast<long long>(B) * D * H * W >= 32768) return false;
  return true;
}
}  // namespace icsg3d

using namespace icsg3d;

// N tile and K split of the per-tap kernel.  Without a workspace: the widest N tile that still yields about one tile per
// SM (never below 64 unless the layer is narrower).  With a workspace and too few tiles to fill the machine (the 4^3 / 2^3
// layers: M = B*64 voxels but K up to 13,824): the WIDEST N tile (least operand re-fetch through L2) and the K range
// (taps x channel chunks) split over ~sms/tiles CTAs, fp32 partials reduced in a fixed order.
static void igemm_tiling(int tiles_m, int nout, int iters, int sms, bool allow_split, int* nt_out, int* ksplit_out) {
  int wide = nout < 256 ? nout : 256;
  while (wide > 16 && nout % wide != 0) wide -= 16;
  *ksplit_out = 1;
  if (allow_split) {
    const long long tiles = static_cast<long long>(tiles_m) * (nout / wide);
    long long ks = sms / tiles;
    if (ks > iters / 4) ks = iters / 4;
    if (ks >= 2) {
      *nt_out = wide;
      *ksplit_out = static_cast<int>(ks);
      return;
    }
  }
  int nt = wide;
  while (nt > 64 && nout % nt != 0) nt -= 16;
  while (nt > 64 && (nt % 2 == 0) && nout % (nt / 2) == 0 && static_cast<long long>(tiles_m) * (nout / nt) < sms) nt /= 2;
  *nt_out = nt;
}

static int conv3d_igemm_impl(int ntaps, const void* x, int ldx, const void* wpack, const float* bias, void* y, int ldy,
                             int y_dtype, int n_store, int B, int D, int H, int W, int cin, int nout, int act,
                             float leaky_alpha, void* stream, double* stats = nullptr, int stats_parts = 0,
                             void* ws = nullptr, int64_t ws_bytes = 0, bool op_f16 = false, float oscale = 1.0f,
                             const float* post_scale = nullptr, const float* post_shift = nullptr) {
  // op_f16: the 2-byte operands are IEEE fp16 instead of bf16 (fp32-class split mode): same kernels and layouts, only
  // the a/b format fields (bits 7..9 / 10..12) of the tcgen05 instruction descriptor change from BF16 (1) to F16 (0)
  const uint32_t fmt_mask = op_f16 ? ~((7u << 7) | (7u << 10)) : ~0u;
  ICSG_REQUIRE(x && wpack && y, "conv3d_k3_igemm: null pointer");
  ICSG_REQUIRE(B > 0 && is_pow2(D) && is_pow2(H) && is_pow2(W) && D >= 2 && H >= 2 && W >= 2 && W <= 128,
               "conv3d_k3_igemm: D,H,W must be powers of two in [2,128] (got %d %d %d)", D, H, W);
  ICSG_REQUIRE(cin >= 16 && cin % 16 == 0, "conv3d_k3_igemm: cin must be a multiple of 16 (got %d)", cin);
  ICSG_REQUIRE(nout >= 16 && nout % 16 == 0, "conv3d_k3_igemm: nout must be a multiple of 16 (got %d)", nout);
  ICSG_REQUIRE(ldx % 8 == 0 && ldx >= cin, "conv3d_k3_igemm: ldx must be >= cin and a multiple of 8 (got %d)", ldx);
  ICSG_REQUIRE(n_store > 0 && n_store <= nout && ldy >= n_store, "conv3d_k3_igemm: bad n_store/ldy (%d/%d)", n_store, ldy);
  ICSG_REQUIRE((reinterpret_cast<uintptr_t>(x) & 15) == 0 && (reinterpret_cast<uintptr_t>(wpack) & 15) == 0,
               "conv3d_k3_igemm: x and wpack must be 16-byte aligned");
  ICSG_REQUIRE(y_dtype == ICSG3D_DT_BF16 || y_dtype == ICSG3D_DT_F32, "conv3d_k3_igemm: bad y_dtype");
  ICSG_REQUIRE((post_scale == nullptr) == (post_shift == nullptr) && !(post_scale && stats),
               "conv3d_k3_igemm: post_scale / post_shift come together and exclude fused statistics");
  const long long m_total = static_cast<long long>(B) * D * H * W;
  ICSG_REQUIRE(m_total < (1ll << 31), "conv3d_k3_igemm: too many voxels");

  {
    const int sms0 = sm_count();
    ConvStreamParams sp;
    if (ntaps == 27 && sms0 > 0 && conv_impl_choice() == 0 && conv_stream_plan(B, D, H, W, cin, nout, sms0, &sp)) {
      ICSG_REQUIRE(!stats || (sp.tiles_n == 1 && stats_parts == conv_stream_grid(sp)),
                   "conv3d_k3_igemm_stats: stats_parts %d does not match this layer's plan", stats_parts);
      for (int i = 0; i < 3; ++i) sp.idesc[i] &= fmt_mask;
      return launch_conv_stream(x, ldx, wpack, bias, y, ldy, y_dtype, n_store, cin, nout, act, leaky_alpha, stats, sp,
                                static_cast<cudaStream_t>(stream), oscale, post_scale, post_shift);
    }
    ConvHaloParams hp;
    if (ntaps == 27 && sms0 > 0 && conv_use_halo(B, D, H, W, cin, nout, sms0, &hp)) {
      ICSG_REQUIRE(!stats || (nout <= 512 && n_store == nout && stats_parts == conv_halo_grid(hp, sms0)),
                   "conv3d_k3_igemm_stats: stats_parts %d does not match this layer's plan", stats_parts);
      hp.idesc &= fmt_mask;
      return launch_conv_halo(x, ldx, wpack, bias, y, ldy, y_dtype, n_store, cin, nout, act, leaky_alpha, hp, sms0,
                              static_cast<cudaStream_t>(stream), oscale, stats, post_scale, post_shift);
    }
    ICSG_REQUIRE(!stats, "conv3d_k3_igemm_stats: this layer shape has no fused-statistics path (stats_parts() == 0)");
  }
  ConvIgemmParams p{};
  p.m_total = static_cast<int>(m_total);
  p.D = D;
  p.H = H;
  p.W = W;
  tile_box(B, D, H, W, &p.bw, &p.bh, &p.bd, &p.bn);
  p.kc = (cin % 64 == 0) ? 64 : (cin % 32 == 0 ? 32 : 16);
  p.chunks = cin / p.kc;
  p.tiles_m = static_cast<int>((m_total + 127) / 128);
  const int sms = sm_count();
  if (sms <= 0) return cuda_fail(cudaGetLastError(), "sm_count", __FILE__, __LINE__);
  p.ntaps = ntaps;
  p.ups = (p.kc == 64) ? 1 : 3;
  if ((ntaps * p.chunks) % p.ups != 0) p.ups = 1;
  p.iters = ntaps * p.chunks / p.ups;
  int nt = 0, ksplit = 1;
  igemm_tiling(p.tiles_m, nout, p.iters, sms, ws != nullptr, &nt, &ksplit);
  if (ksplit > 1 && ws_bytes < static_cast<int64_t>(ksplit) * m_total * nout * 4) {
    igemm_tiling(p.tiles_m, nout, p.iters, sms, false, &nt, &ksplit);  // workspace too small: no split
  }
  ICSG_REQUIRE(nt > 0 && nout % nt == 0, "conv3d_k3_igemm: unsupported nout %d", nout);
  p.nt = nt;
  p.tiles_n = nout / nt;
  p.ips = (p.iters + ksplit - 1) / ksplit;
  p.ksplit = (p.iters + p.ips - 1) / p.ips;
  // Two M tiles per weight tile (both accumulator pairs fit the 512 TMEM columns up to N = 128): the kernel is bound by
  // L2 -> shared-memory operand traffic (ncu: tensor pipe 51 % at N = 128 with 32 KB per 256 MMA cycles), and sharing the
  // weight tile makes it 48 KB per 512.  Needs enough tiles to keep every SM busy with pairs.
  static const int mt_env = [] { const char* e = getenv("ICSG3D_IGEMM_MT"); return e ? atoi(e) : 0; }();
  p.mt = 1;
  if (mt_env != 1 && p.kc == 64 && p.ups == 1 && p.ksplit == 1 && nt <= 128 &&
      static_cast<long long>(p.tiles_m / 2) * p.tiles_n >= 2ll * sms)
    p.mt = 2;
  p.ws = static_cast<float*>(ws);
  p.a_unit_bytes = 128u * p.kc * 2u;
  p.b_unit_bytes = static_cast<uint32_t>(nt) * p.kc * 2u;
  p.stage_bytes = (static_cast<uint32_t>(p.ups) * (static_cast<uint32_t>(p.mt) * p.a_unit_bytes + p.b_unit_bytes) + 1023u) & ~1023u;
  const uint32_t smem_budget = 200u * 1024u;
  int stages = static_cast<int>(smem_budget / p.stage_bytes);
  if (stages > kMaxStages) stages = kMaxStages;
  ICSG_REQUIRE(stages >= 2, "conv3d_k3_igemm: stage too large (%u bytes)", p.stage_bytes);
  p.stages = stages;
  p.sbo = 8u * p.kc * 2u;
  p.layout = umma_layout_for_swizzle(p.kc * 2);
  p.idesc = umma_idesc_bf16(nt, 0, 0) & fmt_mask;
  uint32_t cols = 32;
  while (cols < 2u * p.mt * nt) cols <<= 1;
  p.tmem_cols = cols;
  p.y = y;
  p.ldy = ldy;
  p.y_dtype = y_dtype;
  p.n_store = n_store;
  p.bias = bias;
  p.act = act;
  p.alpha = leaky_alpha;
  p.oscale = oscale;
  p.post_scale = post_scale;
  p.post_shift = post_shift;

  CUtensorMap tmA, tmB;
  int rc = encode_act_map(&tmA, x, ldx, B, D, H, W, cin, p.kc);
  if (rc) return rc;
  {
    uint64_t dims[3] = {static_cast<uint64_t>(cin), static_cast<uint64_t>(nout), static_cast<uint64_t>(ntaps)};
    uint64_t strides[2] = {static_cast<uint64_t>(cin) * 2, static_cast<uint64_t>(nout) * cin * 2};
    uint32_t box[3] = {static_cast<uint32_t>(p.kc), static_cast<uint32_t>(nt), 1};
    rc = encode_tiled_bf16(&tmB, wpack, 3, dims, strides, box, p.kc * 2);
    if (rc) return rc;
  }

  const size_t smem = static_cast<size_t>(p.stages) * p.stage_bytes + 1024;
  static bool configured = false;
  if (!configured) {
    ICSG_CUDA(cudaFuncSetAttribute(conv3d_k3_igemm_kernel<1, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 210 * 1024));
    ICSG_CUDA(cudaFuncSetAttribute(conv3d_k3_igemm_kernel<2, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 210 * 1024));
    ICSG_CUDA(cudaFuncSetAttribute(conv3d_k3_igemm_kernel<4, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 210 * 1024));
    ICSG_CUDA(cudaFuncSetAttribute(conv3d_k3_igemm_kernel<4, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 210 * 1024));
    configured = true;
  }
  const int total_tiles = ((p.tiles_m + p.mt - 1) / p.mt) * p.tiles_n * p.ksplit;
  const int grid = total_tiles < sms ? total_tiles : sms;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (p.kc == 16) launch_k(conv3d_k3_igemm_kernel<1, 1>, grid, kConvThreads, smem, st, tmA, tmB, p);
  else if (p.kc == 32) launch_k(conv3d_k3_igemm_kernel<2, 1>, grid, kConvThreads, smem, st, tmA, tmB, p);
  else if (p.mt == 2) launch_k(conv3d_k3_igemm_kernel<4, 2>, grid, kConvThreads, smem, st, tmA, tmB, p);
  else launch_k(conv3d_k3_igemm_kernel<4, 1>, grid, kConvThreads, smem, st, tmA, tmB, p);
  ICSG_CHECK_LAUNCH();
  if (p.ksplit > 1) {
    const long long items = m_total * (nout / 4);
    long long blocks = (items + 255) / 256;
    if (blocks > sms * 8) blocks = sms * 8;
    launch_k(conv_splitk_reduce_kernel, static_cast<int>(blocks), 256, 0, st, p.ws, p.ksplit, m_total, nout, bias, act, leaky_alpha, y,
                                                                      ldy, y_dtype, n_store, oscale, post_scale, post_shift);
    ICSG_CHECK_LAUNCH();
  }
  return ICSG3D_OK;
}

extern "C" int icsg3d_conv3d_k3_igemm(const void* x, int ldx, const void* wpack, const float* bias, void* y, int ldy,
                                      int y_dtype, int n_store, int B, int D, int H, int W, int cin, int nout, int act,
                                      float leaky_alpha, void* stream) {
  return conv3d_igemm_impl(27, x, ldx, wpack, bias, y, ldy, y_dtype, n_store, B, D, H, W, cin, nout, act, leaky_alpha, stream);
}

// Workspace (bytes) with which icsg3d_conv3d_k3_igemm_ws splits the K range of this layer; 0 = the layer is not split.
extern "C" int64_t icsg3d_conv3d_k3_workspace_bytes(int B, int D, int H, int W, int cin, int nout) {
  const int sms = sm_count();
  if (sms <= 0 || cin < 16 || cin % 16 || nout < 16 || nout % 16) return 0;
  ConvStreamParams sp;
  ConvHaloParams hp;
  if (conv_impl_choice() == 0 && conv_stream_plan(B, D, H, W, cin, nout, sms, &sp)) return 0;
  if (conv_use_halo(B, D, H, W, cin, nout, sms, &hp)) return 0;
  const long long m_total = static_cast<long long>(B) * D * H * W;
  const int kc = (cin % 64 == 0) ? 64 : (cin % 32 == 0 ? 32 : 16);
  const int chunks = cin / kc;
  int ups = (kc == 64) ? 1 : 3;
  if ((27 * chunks) % ups != 0) ups = 1;
  int nt = 0, ksplit = 1;
  igemm_tiling(static_cast<int>((m_total + 127) / 128), nout, 27 * chunks / ups, sms, true, &nt, &ksplit);
  return ksplit > 1 ? static_cast<int64_t>(ksplit) * m_total * nout * 4 : 0;
}

extern "C" int icsg3d_conv3d_k3_igemm_ws(const void* x, int ldx, const void* wpack, const float* bias, void* y, int ldy,
                                         int y_dtype, int n_store, int B, int D, int H, int W, int cin, int nout, int act,
                                         float leaky_alpha, void* ws, int64_t ws_bytes, void* stream) {
  return conv3d_igemm_impl(27, x, ldx, wpack, bias, y, ldy, y_dtype, n_store, B, D, H, W, cin, nout, act, leaky_alpha, stream,
                           nullptr, 0, ws, ws_bytes);
}

// Inference form of Conv3D + activation + BatchNorm (unet.py:277-279 in learning phase 0): the per-channel affine of the
// moving statistics is applied in the conv epilogue, y = post_scale * act(conv + bias) + post_shift
extern "C" int icsg3d_conv3d_k3_igemm_post(const void* x, int ldx, const void* wpack, const float* bias,
                                           const float* post_scale, const float* post_shift, void* y, int ldy, int y_dtype,
                                           int n_store, int B, int D, int H, int W, int cin, int nout, int act,
                                           float leaky_alpha, void* ws, int64_t ws_bytes, void* stream) {
  ICSG_REQUIRE(post_scale && post_shift, "conv3d_k3_igemm_post: post_scale and post_shift required");
  return conv3d_igemm_impl(27, x, ldx, wpack, bias, y, ldy, y_dtype, n_store, B, D, H, W, cin, nout, act, leaky_alpha, stream,
                           nullptr, 0, ws, ws_bytes, false, 1.0f, post_scale, post_shift);
}

// fp16 operands (the fp32-class split mode of csrc/split3.cu); same dispatcher, kernels and layouts
extern "C" int icsg3d_conv3d_k3_igemm_f16(const void* x, int ldx, const void* wpack, const float* bias, void* y, int ldy,
                                          int y_dtype, int n_store, int B, int D, int H, int W, int cin, int nout, int act,
                                          float leaky_alpha, float out_scale, void* stream) {
  return conv3d_igemm_impl(27, x, ldx, wpack, bias, y, ldy, y_dtype, n_store, B, D, H, W, cin, nout, act, leaky_alpha, stream,
                           nullptr, 0, nullptr, 0, true, out_scale);
}

extern "C" int icsg3d_conv3d_k1_igemm_f16(const void* x, int ldx, const void* wpack, const float* bias, void* y, int ldy,
                                          int y_dtype, int n_store, int B, int D, int H, int W, int cin, int nout, int act,
                                          float leaky_alpha, float out_scale, void* stream) {
  return conv3d_igemm_impl(1, x, ldx, wpack, bias, y, ldy, y_dtype, n_store, B, D, H, W, cin, nout, act, leaky_alpha, stream,
                           nullptr, 0, nullptr, 0, true, out_scale);
}

// Fused statistics in the HALO kernel accumulate per-CTA column sums with fp32 shared-memory atomics: the result depends
// (in the last bits) on the arrival order of the epilogue warps, which breaks the bit-for-bit reproducibility of a train
// step (graph replay == eager, N ranks == 1 rank) for a measured gain of 0.8 % of the step.  Off by default;
// ICSG3D_HALO_STATS=1 or icsg3d_conv3d_set_halo_stats(1) turns it on.
static int g_halo_stats = -1;
static bool halo_stats_enabled() {
  if (g_halo_stats < 0) {
    const char* e = getenv("ICSG3D_HALO_STATS");
    g_halo_stats = (e && e[0] == '1') ? 1 : 0;
  }
  return g_halo_stats == 1;
}
extern "C" int icsg3d_conv3d_stream_force(int n_hblk) {
  conv_stream_force_nb(n_hblk);
  return ICSG3D_OK;
}

extern "C" int icsg3d_conv3d_halo_force(int td, int th, int nt) {
  conv_halo_force(td, th, nt);
  return ICSG3D_OK;
}

extern "C" int icsg3d_conv3d_set_halo_stats(int on) {
  g_halo_stats = on ? 1 : 0;
  return ICSG3D_OK;
}

extern "C" int icsg3d_conv3d_k3_stats_parts(int B, int D, int H, int W, int cin, int nout) {
  const int sms = sm_count();
  if (sms <= 0) return 0;
  ConvStreamParams sp;
  if (conv_impl_choice() == 0 && conv_stream_plan(B, D, H, W, cin, nout, sms, &sp))
    return sp.tiles_n == 1 ? conv_stream_grid(sp) : 0;  // streaming kernel: fused statistics only for un-split layers
  ConvHaloParams hp;
  if (halo_stats_enabled() && nout <= 512 && conv_use_halo(B, D, H, W, cin, nout, sms, &hp))
    return conv_halo_grid(hp, sms);
  return 0;
}

extern "C" int icsg3d_conv3d_k3_igemm_stats(const void* x, int ldx, const void* wpack, const float* bias, void* y, int ldy,
                                            int y_dtype, int n_store, int B, int D, int H, int W, int cin, int nout, int act,
                                            float leaky_alpha, double* stats, int stats_parts, void* stream) {
  ICSG_REQUIRE(stats && stats_parts > 0, "conv3d_k3_igemm_stats: stats buffer required");
  return conv3d_igemm_impl(27, x, ldx, wpack, bias, y, ldy, y_dtype, n_store, B, D, H, W, cin, nout, act, leaky_alpha, stream,
                           stats, stats_parts);
}

extern "C" int icsg3d_conv3d_k1_igemm(const void* x, int ldx, const void* wpack, const float* bias, void* y, int ldy,
                                      int y_dtype, int n_store, int B, int D, int H, int W, int cin, int nout, int act,
                                      float leaky_alpha, void* stream) {
  return conv3d_igemm_impl(1, x, ldx, wpack, bias, y, ldy, y_dtype, n_store, B, D, H, W, cin, nout, act, leaky_alpha, stream);
}

extern "C" int icsg3d_conv3d_set_impl(int impl) {
  ICSG_REQUIRE(impl >= 0 && impl <= 2, "conv3d_set_impl: impl must be 0, 1 or 2");
  g_conv_impl = impl;
  g_wgrad_impl = impl == 0 ? 0 : 1;
  return ICSG3D_OK;
}

extern "C" int icsg3d_conv3d_stream_debug(void* buf, int steps) {
  ICSG_REQUIRE(steps >= 0 && (buf || steps == 0), "conv3d_stream_debug: bad arguments");
  conv_stream_set_debug(static_cast<long long*>(buf), steps);
  return ICSG3D_OK;
}

// Diagnostic: which kernel and tiling the dispatcher picks for a layer shape (host only; no device needed).
// out[0..9] = {impl (0 = per-tap TMA kernel, 1 = halo kernel), TD, TH, G, NT, a_bufs, b_stages, items, kc, smem_bytes}
extern "C" int icsg3d_conv3d_k3_plan(int B, int D, int H, int W, int cin, int nout, int sms, int* out) {
  ICSG_REQUIRE(out && sms > 0, "conv3d_k3_plan: bad arguments");
  for (int i = 0; i < 10; ++i) out[i] = 0;
  ConvStreamParams sp;
  if (conv_impl_choice() == 0 && conv_stream_plan(B, D, H, W, cin, nout, sms, &sp)) {
    out[0] = 2; out[1] = sp.R; out[2] = sp.TH; out[3] = sp.T; out[4] = sp.C; out[5] = sp.stages; out[6] = sp.issuers;
    out[7] = conv_stream_grid(sp); out[8] = sp.kc;
    out[9] = static_cast<int>(((sp.w_bytes + 1023u) & ~1023u) + sp.stages * sp.a_stage_bytes);
    return ICSG3D_OK;
  }
  ConvHaloParams hp;
  if (conv_use_halo(B, D, H, W, cin, nout, sms, &hp)) {
    out[0] = 1; out[1] = hp.TD; out[2] = hp.TH; out[3] = hp.G; out[4] = hp.nt; out[5] = hp.a_bufs; out[6] = hp.b_stages;
    out[7] = hp.total_items; out[8] = hp.kc;
    out[9] = static_cast<int>(hp.a_bufs * hp.a_buf_bytes + hp.b_stages * hp.b_unit_bytes);
  }
  return ICSG3D_OK;
}
