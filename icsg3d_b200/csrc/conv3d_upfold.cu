// Conv3D 3x3x3 "same" over concatenate([skip, UpSampling3D(2)(low)]) without the upsampled tensor
// (unet.py:309-332: c13 / c15 / c17 read UpSampling3D(c10 / c14 / c16) next to the encoder skip).
//
// Nearest-neighbour upsampling by 2 followed by a 3-tap filter is, per axis and per output parity r (o = 2i + r), a
// 2-tap filter on the LOW-resolution tensor with summed weights:
//     r = 0:  y[2i]   = w[-1] * low[i-1] + (w[0] + w[+1]) * low[i]
//     r = 1:  y[2i+1] = (w[-1] + w[0]) * low[i] + w[+1] * low[i+1]
// so in 3-D each of the 8 output phases (rd, rh, rw) is a 2x2x2 convolution of the low-resolution tensor (8 taps instead
// of 27: 3.4x fewer FLOPs on those channels, and the 8x larger upsampled tensor is never written or read).  The skip
// channels stay a 27-tap convolution at full resolution; both accumulate into the same TMEM tile.
//
// GEMM view (tcgen05, fp32 accumulate), one CTA tile = one phase x 128 consecutive LOW-resolution voxels x NT channels:
//   Y[(2i + r), :] = sum_{27 taps} Xskip[2i + r + k, :] Ws[k]^T  +  sum_{8 taps} Xlow[i + o(r, t), :] Wf[r][t]^T
//   * the voxels 2i + r + k of one parity class form an ordinary strided 5-D tensor: one TMA map per parity class (8),
//     out-of-bounds zero fill is the "same" padding exactly as in conv3d_igemm.cu;
//   * the packed weights (icsg3d_pack_conv_w_upfold) are a list of K units [unit][nout][64]: 27 * cs skip units shared
//     by all phases, then for each phase 8 * cu folded units;
//   * warp roles, pipeline and epilogue (bias, activation, optional inference-BatchNorm affine) as in conv3d_igemm.cu;
//     the epilogue scatters row i of the tile to output voxel 2i + r.
#include "common.cuh"

namespace icsg3d {

struct TmSet8 {
  CUtensorMap m[8];
};

struct ConvUpfoldParams {
  int mode;             // 0: fprop (tile = phase x low voxels, skip + folded units, output scattered to 2i + r)
                        // 1: dgrad w.r.t. the LOW tensor (tile = low voxels, 64 (phase, tap) x cout/64 units read from the
                        //    parity classes of dY, plain low-resolution output)
  int m_low;            // B * Dl * Hl * Wl
  int Dl, Hl, Wl;       // low-resolution extents (output: 2x)
  int cs, cu;           // 64-channel chunks of the skip / upsampled part
  int nt, tiles_n, tiles_m;
  int stages;
  uint32_t b_unit_bytes, stage_bytes;
  uint32_t idesc, tmem_cols;
  __nv_bfloat16* y;
  int ldy, n_store;
  const float* bias;
  int act;
  float alpha;
  const float* post_scale;
  const float* post_shift;
};

static constexpr int kUfThreads = 192;
static constexpr int kUfMaxStages = 8;
static constexpr uint32_t kUfAUnit = 128u * 64u * 2u;

template <int MT>  // low-resolution tiles (128 voxels) that share every weight tile: 1, or 2 when there are enough tiles
__global__ void __launch_bounds__(kUfThreads, 1)
conv3d_k3_upfold_kernel(const __grid_constant__ TmSet8 tmSkip, const __grid_constant__ CUtensorMap tmLow,
                        const __grid_constant__ CUtensorMap tmB, const ConvUpfoldParams p) {
  pdl_prologue();
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t full_bar[kUfMaxStages];
  __shared__ __align__(8) uint64_t empty_bar[kUfMaxStages];
  __shared__ __align__(8) uint64_t tmem_full_bar[2];
  __shared__ __align__(8) uint64_t tmem_empty_bar[2];
  __shared__ uint32_t tmem_base_slot;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t ring_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* ring = smem_raw + (ring_base - smem_u32(smem_raw));

  if (warp == 0 && lane == 0) {
    for (int i = 0; i < 8; ++i) tma_prefetch_desc(&tmSkip.m[i]);
    tma_prefetch_desc(&tmLow);
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

  // tile index = (tm * 8 + phase) * tiles_n + tn: the 8 phases of one low-resolution box run side by side (L2 reuse)
  const int tiles_mg = (p.tiles_m + MT - 1) / MT;
  const int total_tiles = p.mode == 0 ? tiles_mg * 8 * p.tiles_n : tiles_mg * p.tiles_n;
  const int skip_units = p.mode == 0 ? 27 * p.cs : 0;
  const int units = p.mode == 0 ? skip_units + 8 * p.cu : 64 * p.cs;  // mode 1: cs = cout / 64 chunks of dY

  if (warp == 0) {
    // ===================== TMA producer =====================
    const bool leader = elect_one();
    int stage = 0;
    uint32_t phase_bit = 0;
    const uint32_t tx_bytes = MT * kUfAUnit + p.b_unit_bytes;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      const int tn = tile % p.tiles_n;
      const int tmp = tile / p.tiles_n;
      const int ph = p.mode == 0 ? (tmp & 7) : 0, tm = p.mode == 0 ? (tmp >> 3) : tmp;
      const int rd = ph >> 2, rh = (ph >> 1) & 1, rw = ph & 1;
      int w0[MT], h0[MT], d0[MT], n0[MT];
#pragma unroll
      for (int t = 0; t < MT; ++t) {  // a tile past the end lies beyond the last sample: zero filled, never stored
        int pix = (tm * MT + t) * 128;
        w0[t] = pix % p.Wl;
        pix /= p.Wl;
        h0[t] = pix % p.Hl;
        pix /= p.Hl;
        d0[t] = pix % p.Dl;
        n0[t] = pix / p.Dl;
      }
      int tap = 0, ch = 0;
      for (int u = 0; u < units; ++u) {
        mbar_wait(&empty_bar[stage], phase_bit ^ 1u);
        uint8_t* sa = ring + static_cast<size_t>(stage) * p.stage_bytes;
        uint8_t* sb = sa + MT * kUfAUnit;
        if (leader) {
          mbar_expect_tx(&full_bar[stage], tx_bytes);
          if (p.mode == 1) {
            // dLow[i] += Wf[r][t]^T dY[2 (i - o(r,t)) + r]: `tap` counts the 64 (phase r, tap t) pairs
            const int r = tap >> 3, t = tap & 7;
            const int od = ((t >> 2) & 1) - (((r >> 2) & 1) ^ 1), oh = ((t >> 1) & 1) - (((r >> 1) & 1) ^ 1), ow = (t & 1) - ((r & 1) ^ 1);
#pragma unroll
            for (int t2 = 0; t2 < MT; ++t2)
              tma_load_5d(sa + t2 * kUfAUnit, &tmSkip.m[r], &full_bar[stage], ch * 64, w0[t2] - ow, h0[t2] - oh, d0[t2] - od, n0[t2]);
            tma_load_3d(sb, &tmB, &full_bar[stage], 0, tn * p.nt, u);
          } else if (u < skip_units) {
            // full-resolution voxel 2i + r + (k - 1) = 2 (i + o) + r' of parity class r'
            const int kd = tap / 9, kh = (tap - kd * 9) / 3, kw = tap - kd * 9 - kh * 3;
            const int qd = rd + kd - 1, qh = rh + kh - 1, qw = rw + kw - 1;  // in [-1, 2]
            const int pd = qd & 1, phh = qh & 1, pw = qw & 1;
            const int od = (qd - pd) >> 1, oh = (qh - phh) >> 1, ow = (qw - pw) >> 1;
#pragma unroll
            for (int t2 = 0; t2 < MT; ++t2)
              tma_load_5d(sa + t2 * kUfAUnit, &tmSkip.m[pd * 4 + phh * 2 + pw], &full_bar[stage], ch * 64, w0[t2] + ow, h0[t2] + oh,
                          d0[t2] + od, n0[t2]);
            tma_load_3d(sb, &tmB, &full_bar[stage], 0, tn * p.nt, u);
          } else {
            const int td = tap >> 2, th = (tap >> 1) & 1, tw = tap & 1;
#pragma unroll
            for (int t2 = 0; t2 < MT; ++t2)
              tma_load_5d(sa + t2 * kUfAUnit, &tmLow, &full_bar[stage], ch * 64, w0[t2] + tw - (rw ^ 1), h0[t2] + th - (rh ^ 1),
                          d0[t2] + td - (rd ^ 1), n0[t2]);
            tma_load_3d(sb, &tmB, &full_bar[stage], 0, tn * p.nt, skip_units + (ph * 8 + tap) * p.cu + ch);
          }
        }
        if (++ch == ((p.mode == 1 || u < skip_units) ? p.cs : p.cu)) {
          ch = 0;
          ++tap;
          if (u + 1 == skip_units) tap = 0;
        }
        if (++stage == p.stages) {
          stage = 0;
          phase_bit ^= 1u;
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    const bool leader = elect_one();
    int stage = 0;
    uint32_t phase_bit = 0;
    int local = 0;
    const uint32_t desc_hi = umma_desc_hi(1024u, umma_layout_for_swizzle(128));
    const uint32_t ring_lo = umma_desc_lo(ring_base, 16u);
    const uint32_t stage_lo = p.stage_bytes >> 4, a_unit_lo = kUfAUnit >> 4;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++local) {
      const int acc = local & 1;
      mbar_wait(&tmem_empty_bar[acc], (static_cast<uint32_t>(local >> 1) & 1u) ^ 1u);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(acc * MT * p.nt);
      for (int u = 0; u < units; ++u) {
        mbar_wait(&full_bar[stage], phase_bit);
        tc_fence_after();
        if (leader) {
          const uint32_t a_lo = ring_lo + static_cast<uint32_t>(stage) * stage_lo;
          const uint32_t b_lo = a_lo + MT * a_unit_lo;
#pragma unroll
          for (int t = 0; t < MT; ++t) {
#pragma unroll
            for (int k = 0; k < 4; ++k)
              umma_bf16_lohi(d_tmem + static_cast<uint32_t>(t * p.nt), a_lo + t * a_unit_lo + 2u * k, desc_hi, b_lo + 2u * k, desc_hi,
                             p.idesc, (u | k) != 0 ? 1u : 0u);
          }
          umma_commit(&empty_bar[stage]);
        }
        if (++stage == p.stages) {
          stage = 0;
          phase_bit ^= 1u;
        }
      }
      if (leader) umma_commit(&tmem_full_bar[acc]);
    }
  } else {
    // ===================== epilogue warps (2..5) =====================
    const int quarter = warp & 3;
    const int row = quarter * 32 + lane;
    int local = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++local) {
      const int acc = local & 1;
      const int tn = tile % p.tiles_n;
      const int tmp = tile / p.tiles_n;
      const int ph = p.mode == 0 ? (tmp & 7) : 0, tm = p.mode == 0 ? (tmp >> 3) : tmp;
      mbar_wait(&tmem_full_bar[acc], static_cast<uint32_t>(local >> 1) & 1u);
      tc_fence_after();
#pragma unroll 1
      for (int t2 = 0; t2 < MT; ++t2) {
      const int pl = (tm * MT + t2) * 128 + row;  // low-resolution voxel of this row
      const bool row_ok = pl < p.m_low;
      long long pixel = pl;
      if (p.mode == 0) {
        int t = pl;
        const int wl = t % p.Wl;
        t /= p.Wl;
        const int hl = t % p.Hl;
        t /= p.Hl;
        const int dl = t % p.Dl;
        const int n = t / p.Dl;
        pixel = ((static_cast<long long>(n) * (2 * p.Dl) + 2 * dl + (ph >> 2)) * (2 * p.Hl) + 2 * hl + ((ph >> 1) & 1)) * (2 * p.Wl) +
                2 * wl + (ph & 1);
      }
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + static_cast<uint32_t>((acc * MT + t2) * p.nt);
      for (int c0 = 0; c0 < p.nt; c0 += 16) {
        uint32_t v[16];
        tmem_ld16(taddr + static_cast<uint32_t>(c0), v);
        tmem_ld_wait();
        const int col0 = tn * p.nt + c0;
        if (!row_ok || col0 >= p.n_store) continue;
        float f[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          float x = __uint_as_float(v[i]);
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
        __nv_bfloat16* dst = p.y + pixel * p.ldy + col0;
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
      }
      }  // t2
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

// Folded weight pack.  w: fp32 [27][cin][cout] (Keras kernel (3,3,3,cin,cout)), input channels [c_skip0, c_skip0 + c_skip)
// are the skip tensor, [c_up0, c_up0 + c_up) the upsampled one.  out: bf16 [27*cs + 64*cu][nout_pad][64], K-major units:
//   unit tap*cs + ch                         : w[tap][c_skip0 + 64 ch + k][co]
//   unit 27 cs + (phase*8 + t)*cu + ch       : sum of w[kd][kh][kw][c_up0 + 64 ch + k][co] over the taps that phase
//                                              (rd,rh,rw) folds onto low-resolution offset t = (td,th,tw)
// per axis: r = 0: t = 0 <- {k = 0}, t = 1 <- {k = 1, 2};  r = 1: t = 0 <- {k = 0, 1}, t = 1 <- {k = 2}   (k in 0..2)
__global__ void pack_w_upfold_kernel(const float* __restrict__ w, int cin, int cout, int c_skip0, int c_skip, int c_up0,
                                     int c_up, int nout_pad, __nv_bfloat16* __restrict__ out) {
  pdl_prologue();
  const int cs = (c_skip + 63) / 64, cu = (c_up + 63) / 64;
  const long long total = static_cast<long long>(27 * cs + 64 * cu) * nout_pad * 64;
  for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int k = static_cast<int>(idx & 63);
    const int co = static_cast<int>((idx >> 6) % nout_pad);
    const int unit = static_cast<int>((idx >> 6) / nout_pad);
    float v = 0.f;
    if (co < cout) {
      if (unit < 27 * cs) {
        const int tap = unit / cs, c = (unit - tap * cs) * 64 + k;
        if (c < c_skip) v = w[(static_cast<long long>(tap) * cin + c_skip0 + c) * cout + co];
      } else {
        const int r = unit - 27 * cs;
        const int c = (r % cu) * 64 + k;
        const int pt = r / cu, ph = pt >> 3, t = pt & 7;
        if (c < c_up) {
          const int lo[3] = {((ph >> 2) & 1) == 0 ? ((t >> 2) & 1 ? 1 : 0) : ((t >> 2) & 1 ? 2 : 0),
                             ((ph >> 1) & 1) == 0 ? ((t >> 1) & 1 ? 1 : 0) : ((t >> 1) & 1 ? 2 : 0),
                             (ph & 1) == 0 ? (t & 1 ? 1 : 0) : (t & 1 ? 2 : 0)};
          const int hi[3] = {((ph >> 2) & 1) == 0 ? ((t >> 2) & 1 ? 2 : 0) : ((t >> 2) & 1 ? 2 : 1),
                             ((ph >> 1) & 1) == 0 ? ((t >> 1) & 1 ? 2 : 0) : ((t >> 1) & 1 ? 2 : 1),
                             (ph & 1) == 0 ? (t & 1 ? 2 : 0) : (t & 1 ? 2 : 1)};
          for (int kd = lo[0]; kd <= hi[0]; ++kd)
            for (int kh = lo[1]; kh <= hi[1]; ++kh)
              for (int kw = lo[2]; kw <= hi[2]; ++kw)
                v += w[(static_cast<long long>((kd * 3 + kh) * 3 + kw) * cin + c_up0 + c) * cout + co];
        }
      }
    }
    out[idx] = f2bf(v);
  }
}

// Folded weights for the gradient w.r.t. the LOW tensor (mode 1): out bf16 [64 * cc][cup_pad][64], cc = ceil(cout / 64):
//   unit (r*8 + t)*cc + ch, row n, column k  =  Wf[r][t][c_up0 + n][64 ch + k]   (the same tap sums, N = low channel, K = cout)
__global__ void pack_w_upfold_dgrad_kernel(const float* __restrict__ w, int cin, int cout, int c_up0, int c_up, int cup_pad,
                                           __nv_bfloat16* __restrict__ out) {
  pdl_prologue();
  const int cc = (cout + 63) / 64;
  const long long total = 64ll * cc * cup_pad * 64;
  for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int k = static_cast<int>(idx & 63);
    const int n = static_cast<int>((idx >> 6) % cup_pad);
    const int unit = static_cast<int>((idx >> 6) / cup_pad);
    const int co = (unit % cc) * 64 + k;
    const int pt = unit / cc, ph = pt >> 3, t = pt & 7;
    float v = 0.f;
    if (n < c_up && co < cout) {
      const int lo[3] = {((ph >> 2) & 1) == 0 ? ((t >> 2) & 1 ? 1 : 0) : ((t >> 2) & 1 ? 2 : 0),
                         ((ph >> 1) & 1) == 0 ? ((t >> 1) & 1 ? 1 : 0) : ((t >> 1) & 1 ? 2 : 0),
                         (ph & 1) == 0 ? (t & 1 ? 1 : 0) : (t & 1 ? 2 : 0)};
      const int hi[3] = {((ph >> 2) & 1) == 0 ? ((t >> 2) & 1 ? 2 : 0) : ((t >> 2) & 1 ? 2 : 1),
                         ((ph >> 1) & 1) == 0 ? ((t >> 1) & 1 ? 2 : 0) : ((t >> 1) & 1 ? 2 : 1),
                         (ph & 1) == 0 ? (t & 1 ? 2 : 0) : (t & 1 ? 2 : 1)};
      for (int kd = lo[0]; kd <= hi[0]; ++kd)
        for (int kh = lo[1]; kh <= hi[1]; ++kh)
          for (int kw = lo[2]; kw <= hi[2]; ++kw)
            v += w[(static_cast<long long>((kd * 3 + kh) * 3 + kw) * cin + c_up0 + n) * cout + co];
    }
    out[idx] = f2bf(v);
  }
}

static bool uf_pow2(int v) { return v > 0 && (v & (v - 1)) == 0; }

}  // namespace icsg3d

using namespace icsg3d;

extern "C" int64_t icsg3d_conv3d_upfold_wpack_elems(int c_skip, int c_up, int nout) {
  if (c_skip <= 0 || c_up <= 0 || nout <= 0) return -1;
  const int cs = (c_skip + 63) / 64, cu = (c_up + 63) / 64, np = (nout + 15) / 16 * 16;
  return static_cast<int64_t>(27 * cs + 64 * cu) * np * 64;
}

extern "C" int icsg3d_pack_conv_w_upfold(const float* w, int cin, int cout, int c_skip0, int c_skip, int c_up0, int c_up,
                                         void* wfold, void* stream) {
  ICSG_REQUIRE(w && wfold && cin > 0 && cout > 0 && c_skip > 0 && c_up > 0 && c_skip0 >= 0 && c_up0 >= 0 &&
                   c_skip0 + c_skip <= cin && c_up0 + c_up <= cin,
               "pack_conv_w_upfold: bad channel ranges");
  const int np = (cout + 15) / 16 * 16;
  const long long total = icsg3d_conv3d_upfold_wpack_elems(c_skip, c_up, cout);
  long long blocks = (total + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  launch_k(pack_w_upfold_kernel, static_cast<int>(blocks), 256, 0, static_cast<cudaStream_t>(stream), w, cin, cout, c_skip0,
           c_skip, c_up0, c_up, np, static_cast<__nv_bfloat16*>(wfold));
  ICSG_CHECK_LAUNCH();
  return ICSG3D_OK;
}

// Shared launcher.  x_par: the full-resolution tensor read through its 8 parity classes (mode 0: the skip input with
// c_par channels; mode 1: dY with c_par = cout channels); x_low (mode 0 only): the low-resolution input.
static int upfold_launch(int mode, const void* x_par, int ld_par, int c_par, const void* x_low, int ld_low, int c_low,
                         const void* wpack, int n_units, const float* bias, const float* post_scale, const float* post_shift,
                         void* y, int ldy, int n_store, int B, int D, int H, int W, int nout, int act, float leaky_alpha,
                         void* stream) {
  const int Dl = D / 2, Hl = H / 2, Wl = W / 2;
  const long long m_low = static_cast<long long>(B) * Dl * Hl * Wl;
  ICSG_REQUIRE(m_low * 8 < (1ll << 31), "conv3d_k3_upfold: too many voxels");
  const int sms = sm_count();
  if (sms <= 0) return cuda_fail(cudaGetLastError(), "sm_count", __FILE__, __LINE__);

  ConvUpfoldParams p{};
  p.mode = mode;
  p.m_low = static_cast<int>(m_low);
  p.Dl = Dl;
  p.Hl = Hl;
  p.Wl = Wl;
  p.cs = c_par / 64;
  p.cu = c_low / 64;
  p.tiles_m = static_cast<int>((m_low + 127) / 128);
  const int phases = mode == 0 ? 8 : 1;
  int nt = nout < 256 ? nout : 256;
  while (nt > 16 && nout % nt != 0) nt -= 16;
  while (nt > 64 && nt % 32 == 0 && static_cast<long long>(p.tiles_m) * phases * (nout / nt) < sms) nt /= 2;  // small problems: more tiles
  ICSG_REQUIRE(nout % nt == 0, "conv3d_k3_upfold: unsupported N %d", nout);
  p.nt = nt;
  p.tiles_n = nout / nt;
  p.b_unit_bytes = static_cast<uint32_t>(nt) * 128u;
  // two low-resolution tiles per weight tile where both accumulator pairs fit TMEM and there are enough tiles (as in the
  // per-tap kernel: the operand traffic L2 -> shared memory bounds the kernel)
  static const int mt_env = [] { const char* e = getenv("ICSG3D_IGEMM_MT"); return e ? atoi(e) : 0; }();
  const int mt = (mt_env != 1 && nt <= 128 && static_cast<long long>(p.tiles_m / 2) * phases * p.tiles_n >= 2ll * sms) ? 2 : 1;
  p.stage_bytes = (mt * kUfAUnit + p.b_unit_bytes + 1023u) & ~1023u;
  int stages = static_cast<int>(200u * 1024u / p.stage_bytes);
  if (stages > kUfMaxStages) stages = kUfMaxStages;
  p.stages = stages;
  p.idesc = umma_idesc_bf16(nt, 0, 0);
  uint32_t cols = 32;
  while (cols < 2u * mt * nt) cols <<= 1;
  p.tmem_cols = cols;
  p.y = static_cast<__nv_bfloat16*>(y);
  p.ldy = ldy;
  p.n_store = n_store;
  p.bias = bias;
  p.act = act;
  p.alpha = leaky_alpha;
  p.post_scale = post_scale;
  p.post_shift = post_shift;

  // 128 consecutive low-resolution voxels as a TMA box over (Wl, Hl, Dl, B)
  int rem = 128;
  const int bw = Wl < rem ? Wl : rem;
  rem /= bw;
  const int bh = Hl < rem ? Hl : rem;
  rem /= bh;
  const int bd = Dl < rem ? Dl : rem;
  rem /= bd;
  const int bn = rem;
  const uint32_t box[5] = {64u, static_cast<uint32_t>(bw), static_cast<uint32_t>(bh), static_cast<uint32_t>(bd),
                           static_cast<uint32_t>(bn)};
  TmSet8 tmPar;
  CUtensorMap tmLow, tmB;
  {
    // parity class (pd, ph, pw) of the full-resolution tensor: voxel (2i + p) -> index i, strides doubled
    const uint64_t dims[5] = {static_cast<uint64_t>(c_par), static_cast<uint64_t>(Wl), static_cast<uint64_t>(Hl),
                              static_cast<uint64_t>(Dl), static_cast<uint64_t>(B)};
    const uint64_t sW = static_cast<uint64_t>(ld_par) * 2, sH = sW * W, sD = sH * H, sN = sD * D;
    const uint64_t strides[4] = {2 * sW, 2 * sH, 2 * sD, sN};
    for (int c = 0; c < 8; ++c) {
      const uint8_t* base = static_cast<const uint8_t*>(x_par) + ((c >> 2) & 1) * sD + ((c >> 1) & 1) * sH + (c & 1) * sW;
      int rc = encode_tiled_bf16(&tmPar.m[c], base, 5, dims, strides, box, 128);
      if (rc) return rc;
    }
  }
  if (mode == 0) {
    const uint64_t dims[5] = {static_cast<uint64_t>(c_low), static_cast<uint64_t>(Wl), static_cast<uint64_t>(Hl),
                              static_cast<uint64_t>(Dl), static_cast<uint64_t>(B)};
    const uint64_t sW = static_cast<uint64_t>(ld_low) * 2;
    const uint64_t strides[4] = {sW, sW * Wl, sW * Wl * Hl, sW * Wl * Hl * Dl};
    int rc = encode_tiled_bf16(&tmLow, x_low, 5, dims, strides, box, 128);
    if (rc) return rc;
  } else {
    tmLow = tmPar.m[0];  // unused in mode 1
  }
  {
    const int np = (nout + 15) / 16 * 16;
    const uint64_t dims[3] = {64, static_cast<uint64_t>(np), static_cast<uint64_t>(n_units)};
    const uint64_t strides[2] = {128, static_cast<uint64_t>(np) * 128};
    const uint32_t bbox[3] = {64, static_cast<uint32_t>(nt), 1};
    int rc = encode_tiled_bf16(&tmB, wpack, 3, dims, strides, bbox, 128);
    if (rc) return rc;
  }
  const size_t smem = static_cast<size_t>(p.stages) * p.stage_bytes + 1024;
  static bool configured = false;
  if (!configured) {
    ICSG_CUDA(cudaFuncSetAttribute(conv3d_k3_upfold_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 210 * 1024));
    ICSG_CUDA(cudaFuncSetAttribute(conv3d_k3_upfold_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 210 * 1024));
    configured = true;
  }
  const int total_tiles = ((p.tiles_m + mt - 1) / mt) * phases * p.tiles_n;
  const int grid = total_tiles < sms ? total_tiles : sms;
  if (mt == 2) launch_k(conv3d_k3_upfold_kernel<2>, grid, kUfThreads, smem, static_cast<cudaStream_t>(stream), tmPar, tmLow, tmB, p);
  else launch_k(conv3d_k3_upfold_kernel<1>, grid, kUfThreads, smem, static_cast<cudaStream_t>(stream), tmPar, tmLow, tmB, p);
  ICSG_CHECK_LAUNCH();
  return ICSG3D_OK;
}

extern "C" int icsg3d_conv3d_k3_upfold(const void* x_skip, int ld_skip, int c_skip, const void* x_low, int ld_low, int c_low,
                                       const void* wfold, const float* bias, const float* post_scale, const float* post_shift,
                                       void* y, int ldy, int n_store, int B, int D, int H, int W, int nout, int act,
                                       float leaky_alpha, void* stream) {
  ICSG_REQUIRE(x_skip && x_low && wfold && y, "conv3d_k3_upfold: null pointer");
  ICSG_REQUIRE(B > 0 && uf_pow2(D) && uf_pow2(H) && uf_pow2(W) && D >= 4 && H >= 4 && W >= 4 && W <= 256,
               "conv3d_k3_upfold: D,H,W (output extents) must be powers of two in [4,256] (got %d %d %d)", D, H, W);
  ICSG_REQUIRE(c_skip >= 64 && c_skip % 64 == 0 && c_low >= 64 && c_low % 64 == 0,
               "conv3d_k3_upfold: channel counts must be multiples of 64 (got %d skip, %d low)", c_skip, c_low);
  ICSG_REQUIRE(nout >= 16 && nout % 16 == 0, "conv3d_k3_upfold: nout must be a multiple of 16 (got %d)", nout);
  ICSG_REQUIRE(ld_skip % 8 == 0 && ld_skip >= c_skip && ld_low % 8 == 0 && ld_low >= c_low && ldy >= n_store && n_store > 0 &&
                   n_store <= nout,
               "conv3d_k3_upfold: bad leading dimensions");
  ICSG_REQUIRE(((reinterpret_cast<uintptr_t>(x_skip) | reinterpret_cast<uintptr_t>(x_low) | reinterpret_cast<uintptr_t>(wfold)) & 15) == 0,
               "conv3d_k3_upfold: operands must be 16-byte aligned");
  ICSG_REQUIRE((post_scale == nullptr) == (post_shift == nullptr), "conv3d_k3_upfold: post_scale / post_shift come together");
  return upfold_launch(0, x_skip, ld_skip, c_skip, x_low, ld_low, c_low, wfold, 27 * (c_skip / 64) + 64 * (c_low / 64), bias,
                       post_scale, post_shift, y, ldy, n_store, B, D, H, W, nout, act, leaky_alpha, stream);
}

extern "C" int64_t icsg3d_conv3d_upfold_dgrad_wpack_elems(int cout, int c_up) {
  if (cout <= 0 || c_up <= 0) return -1;
  return 64ll * ((cout + 63) / 64) * ((c_up + 15) / 16 * 16) * 64;
}

extern "C" int icsg3d_pack_conv_w_upfold_dgrad(const float* w, int cin, int cout, int c_up0, int c_up, void* wpack, void* stream) {
  ICSG_REQUIRE(w && wpack && cin > 0 && cout > 0 && c_up > 0 && c_up0 >= 0 && c_up0 + c_up <= cin,
               "pack_conv_w_upfold_dgrad: bad channel range");
  const long long total = icsg3d_conv3d_upfold_dgrad_wpack_elems(cout, c_up);
  long long blocks = (total + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  launch_k(pack_w_upfold_dgrad_kernel, static_cast<int>(blocks), 256, 0, static_cast<cudaStream_t>(stream), w, cin, cout, c_up0,
           c_up, (c_up + 15) / 16 * 16, static_cast<__nv_bfloat16*>(wpack));
  ICSG_CHECK_LAUNCH();
  return ICSG3D_OK;
}

// Gradient of the folded convolution w.r.t. the LOW-resolution input: dlow [B,D/2,H/2,W/2,ld_low] (c_up channels) from
// dy [B,D,H,W,ld_dy] (cout channels) — what UpSampling3D's backward (the sum over the 8 children) of the full-resolution
// data gradient gives, at 8/27 of its multiply-adds and without the full-resolution gradient of those channels.
extern "C" int icsg3d_conv3d_k3_upfold_dgrad_low(const void* dy, int ld_dy, int cout, const void* wpack, void* dlow, int ld_low,
                                                 int c_up, int B, int D, int H, int W, void* stream) {
  ICSG_REQUIRE(dy && wpack && dlow, "conv3d_k3_upfold_dgrad_low: null pointer");
  ICSG_REQUIRE(B > 0 && uf_pow2(D) && uf_pow2(H) && uf_pow2(W) && D >= 4 && H >= 4 && W >= 4 && W <= 256,
               "conv3d_k3_upfold_dgrad_low: D,H,W (extents of dy) must be powers of two in [4,256] (got %d %d %d)", D, H, W);
  ICSG_REQUIRE(cout >= 64 && cout % 64 == 0 && c_up >= 16 && c_up % 16 == 0,
               "conv3d_k3_upfold_dgrad_low: cout must be a multiple of 64 and c_up of 16 (got %d, %d)", cout, c_up);
  ICSG_REQUIRE(ld_dy % 8 == 0 && ld_dy >= cout && ld_low >= c_up, "conv3d_k3_upfold_dgrad_low: bad leading dimensions");
  ICSG_REQUIRE(((reinterpret_cast<uintptr_t>(dy) | reinterpret_cast<uintptr_t>(wpack)) & 15) == 0,
               "conv3d_k3_upfold_dgrad_low: operands must be 16-byte aligned");
  return upfold_launch(1, dy, ld_dy, cout, nullptr, 0, 0, wpack, 64 * (cout / 64), nullptr, nullptr, nullptr, dlow, ld_low, c_up,
                       B, D, H, W, c_up, ICSG3D_ACT_NONE, 0.f, stream);
}
