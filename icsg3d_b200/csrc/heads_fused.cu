// generate.py:220-225 in one kernel: the two 1x1x1 head convolutions of the segmentation U-Net (unet.py:338-341: softmax
// head with c1 classes, sigmoid atom-mask head), np.argmax over the float32 softmax output and the sigmoid >= threshold
// mask.  The head logits (96 fp32 per voxel = 384 B against the 256 B of bf16 features they are computed from) never
// leave the SM: the GEMM accumulates them in TMEM, the epilogue threads (one per voxel) read them from there and write
// 1 + 1 (+ 4) bytes per voxel.  HBM traffic is the feature tensor, once.
//
//   logits[M = B*D*H*W, N = nout] = X[M, K = cin] * Wp[N, K]^T       (tcgen05, fp32 accumulate)
//   * X is streamed through a ring of 128-row x 64-channel TMA boxes (16 KB, 128-byte swizzle), the packed weights
//     (icsg3d_pack_heads_w: [nout][cin], K-major) are loaded once per CTA and stay in shared memory;
//   * warp 0 = TMA producer, warp 1 = MMA issuer + TMEM allocator, warps 2..5 / 6..9 = two epilogue groups that own
//     accumulator slot 0 / 1, so the softmax/argmax of tile i overlaps the loads and MMAs of tiles i+1, i+2;
//   * the softmax sum is added in exactly the order of heads_predict_kernel (post.cu) — lane partials over columns
//     l, l+32, l+64, then the xor butterfly — so both kernels return identical labels for identical logits.
#include "common.cuh"

namespace icsg3d {

struct HeadsFusedParams {
  long long m_total;
  int tiles_m;
  int chunks;  // cin / 64
  int nout;    // GEMM N (multiple of 16, <= 96)
  int c1;      // softmax classes; the sigmoid logit is column c1
  int stages;
  uint32_t idesc;
  float oscale;
  float threshold;
  int no_ties;  // experiment only: skip the tie evaluation (labels = arg-max of the logits)
  const float* bias;
  uint8_t* argmax;
  uint8_t* mask;
  float* sigp;
  // training form (heads_loss_fused): losses + metrics partials + gradient w.r.t. the logits
  const uint8_t* species;   // [M] class index (0 = empty); the binary target is species != 0
  const float* class_w;     // [c1] or null
  float inv_count;          // 1 / (voxels of the global batch)
  __nv_bfloat16* dlogits;   // [M][ldd] bf16 or null
  int ldd;
  double* partials;         // [grid][6]: soft loss, sig loss, tp, predicted, tp_w, possible_w (heads.cu::kHeadsTerms)
};

static constexpr int kHfThreads = 320;
static constexpr int kHfMaxStages = 10;
static constexpr uint32_t kHfChunkBytes = 128u * 64u * 2u;
static constexpr uint32_t kHfSlotCols = 128;  // TMEM columns per accumulator slot (>= 96)

// ---- per-voxel math, inference: label = np.argmax of the float32 soft-max output (generate.py:221) ----
__device__ __forceinline__ int hf_label(float (&x)[96], int c1, int no_ties) {
  float mxk[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) mxk[k] = -INFINITY;
#pragma unroll
  for (int c = 0; c < 96; ++c) mxk[c & 7] = fmaxf(mxk[c & 7], x[c]);
  const float mx = fmaxf(fmaxf(fmaxf(mxk[0], mxk[1]), fmaxf(mxk[2], mxk[3])), fmaxf(fmaxf(mxk[4], mxk[5]), fmaxf(mxk[6], mxk[7])));
  // np.argmax over the float32 soft-max output: p_c = e_c * inv with e_c = exp(x_c - max) <= 1, so the largest p is
  // inv itself (e == 1 at the largest logit) and the label is the FIRST class whose p rounds to inv.  Classes equal to
  // the maximum tie exactly; a class slightly below it can still tie (e_c rounds to 1, or e_c == 1 - 2^-24 and inv
  // is a power of two; e_c <= 1 - 2^-23 never does), and only matters when it comes BEFORE the first exact maximum.
  // e_c >= 1 - 2^-24 with ex2.approx good to 2^-22 needs x_c >= max - 3e-7: x_c >= max - 1e-6 bounds those classes.
  const float thr = mx - 1e-6f;
  uint32_t meq[3], mnear[3];
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    uint32_t eq[4] = {0u, 0u, 0u, 0u}, nr[4] = {0u, 0u, 0u, 0u};
#pragma unroll
    for (int l = 0; l < 32; ++l) {
      eq[l & 3] |= x[32 * j + l] == mx ? 1u << l : 0u;
      nr[l & 3] |= x[32 * j + l] >= thr ? 1u << l : 0u;
    }
    meq[j] = (eq[0] | eq[1]) | (eq[2] | eq[3]);
    mnear[j] = (nr[0] | nr[1]) | (nr[2] | nr[3]);
  }
  const int cmax = meq[0] ? __ffs(meq[0]) - 1 : meq[1] ? 31 + __ffs(meq[1]) : 63 + __ffs(meq[2]);
  int amax = cmax;
  uint32_t before[3];
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    const int nb = cmax - 32 * j;  // bits of word j below the first exact maximum
    before[j] = mnear[j] & (nb >= 32 ? 0xffffffffu : nb <= 0 ? 0u : (1u << nb) - 1u);
  }
  if ((before[0] | before[1] | before[2]) != 0u && !no_ties) {  // rare: evaluate those classes like heads_predict_kernel
    float inv = 0.f;
    bool done = false;
#pragma unroll 1
    for (int j = 0; j < 3 && !done; ++j) {
      uint32_t m = before[j];
      while (m != 0u && !done) {
        const int idx = 32 * j + __ffs(m) - 1;
        m &= m - 1u;
        float xv = 0.f;
#pragma unroll
        for (int c = 0; c < 96; ++c) xv = c == idx ? x[c] : xv;
        const float e = expf(xv - mx);
        if (e == 1.f) {
          amax = idx;
          done = true;
        } else if (__float_as_uint(e) == 0x3F7FFFFFu) {
          if (inv == 0.f) {  // soft-max denominator in heads_predict_kernel's order: lane partials, xor butterfly
            float s[32];
#pragma unroll
            for (int l = 0; l < 32; ++l) s[l] = 0.f;
#pragma unroll
            for (int jj = 0; jj < 3; ++jj) {
#pragma unroll
              for (int l = 0; l < 32; ++l) s[l] += 32 * jj + l < c1 ? expf(x[32 * jj + l] - mx) : 0.f;
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
              for (int l = 0; l < o; ++l) s[l] += s[l + o];
            }
            inv = 1.f / s[0];
          }
          if (e * inv == inv) {
            amax = idx;
            done = true;
          }
        }
      }
    }
  }
  return amax;
}

template <bool kTrain>
__global__ void __launch_bounds__(kHfThreads, 1)
heads_predict_fused_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                           const HeadsFusedParams p) {
  pdl_prologue();
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t full_bar[kHfMaxStages];
  __shared__ __align__(8) uint64_t empty_bar[kHfMaxStages];
  __shared__ __align__(8) uint64_t w_bar;
  __shared__ __align__(8) uint64_t tmem_full_bar[2];
  __shared__ __align__(8) uint64_t tmem_empty_bar[2];
  __shared__ uint32_t tmem_base_slot;
  __shared__ float s_bias[96];  // soft-max head biases; -inf from column c1 on (those columns never win the maximum)
  __shared__ float s_bias_sig;  // sigmoid head bias (column c1)

  __shared__ double s_red[8][6];
  double acc[2] = {0.0, 0.0};  // training: this thread's share of the two loss sums ...
  int cnt[4] = {0, 0, 0, 0};   // ... and of the four metric counts

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t w_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* w_smem = smem_raw + (w_base - smem_u32(smem_raw));
  const uint32_t w_chunk_bytes = static_cast<uint32_t>(p.nout) * 128u;
  const uint32_t ring_base = w_base + static_cast<uint32_t>(p.chunks) * w_chunk_bytes;
  uint8_t* ring = w_smem + static_cast<size_t>(p.chunks) * w_chunk_bytes;

  if (threadIdx.x < 96) {
    const int c = static_cast<int>(threadIdx.x);
    s_bias[c] = c < p.c1 ? (p.bias != nullptr ? p.bias[c] : 0.f) : -INFINITY;
    if (c == p.c1) s_bias_sig = p.bias != nullptr ? p.bias[c] : 0.f;
  }
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(&w_bar, 1);
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tmem_full_bar[a], 1);
      mbar_init(&tmem_empty_bar[a], 4);
    }
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc(&tmem_base_slot, 2 * kHfSlotCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_slot;

  if (warp == 0) {
    // ===================== TMA producer =====================
    const bool leader = elect_one();
    if (leader) {
      mbar_expect_tx(&w_bar, static_cast<uint32_t>(p.chunks) * w_chunk_bytes);
      for (int c = 0; c < p.chunks; ++c) tma_load_2d(w_smem + static_cast<size_t>(c) * w_chunk_bytes, &tmB, &w_bar, c * 64, 0);
    }
    int stage = 0;
    uint32_t phase = 0;
    for (int tile = blockIdx.x; tile < p.tiles_m; tile += gridDim.x) {
      for (int c = 0; c < p.chunks; ++c) {
        mbar_wait(&empty_bar[stage], phase ^ 1u);
        if (leader) {
          mbar_expect_tx(&full_bar[stage], kHfChunkBytes);
          tma_load_2d(ring + static_cast<size_t>(stage) * kHfChunkBytes, &tmA, &full_bar[stage], c * 64, tile * 128);
        }
        if (++stage == p.stages) {
          stage = 0;
          phase ^= 1u;
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    const bool leader = elect_one();
    int stage = 0;
    uint32_t phase = 0;
    int local = 0;
    const uint32_t desc_hi = umma_desc_hi(1024u, umma_layout_for_swizzle(128));
    const uint32_t ring_lo = umma_desc_lo(ring_base, 16u);
    const uint32_t w_lo = umma_desc_lo(w_base, 16u);
    mbar_wait(&w_bar, 0u);
    tc_fence_after();
    for (int tile = blockIdx.x; tile < p.tiles_m; tile += gridDim.x, ++local) {
      const int slot = local & 1;
      mbar_wait(&tmem_empty_bar[slot], (static_cast<uint32_t>(local >> 1) & 1u) ^ 1u);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(slot) * kHfSlotCols;
      for (int c = 0; c < p.chunks; ++c) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        if (leader) {
          const uint32_t a_lo = ring_lo + static_cast<uint32_t>(stage) * (kHfChunkBytes >> 4);
          const uint32_t b_lo = w_lo + static_cast<uint32_t>(c) * (w_chunk_bytes >> 4);
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_bf16_lohi(d_tmem, a_lo + 2u * k, desc_hi, b_lo + 2u * k, desc_hi, p.idesc, (c | k) != 0 ? 1u : 0u);
          umma_commit(&empty_bar[stage]);
        }
        if (++stage == p.stages) {
          stage = 0;
          phase ^= 1u;
        }
      }
      if (leader) umma_commit(&tmem_full_bar[slot]);
    }
  } else {
    // ===================== epilogue: group 0 = warps 2..5 (slot 0), group 1 = warps 6..9 (slot 1) =====================
    const int group = (warp - 2) >> 2;
    const int quarter = warp & 3;  // TMEM lane quarter this warp may access
    const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + static_cast<uint32_t>(group) * kHfSlotCols;
    int local = 0;
    for (int tile = blockIdx.x; tile < p.tiles_m; tile += gridDim.x, ++local) {
      if ((local & 1) != group) continue;
      mbar_wait(&tmem_full_bar[group], static_cast<uint32_t>(local >> 1) & 1u);
      tc_fence_after();
      const long long pixel = static_cast<long long>(tile) * 128 + quarter * 32 + lane;
      if constexpr (!kTrain) {
        // logits of this thread's voxel; columns >= c1 come out as -inf or NaN (never-written TMEM columns when
        // nout < 96): fmaxf and the comparisons below ignore both
        float x[96], sl;
        {
          uint32_t vs;
          asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(vs) : "r"(taddr + static_cast<uint32_t>(p.c1)));
#pragma unroll
          for (int j = 0; j < 3; ++j) {
            uint32_t v[32];
            tmem_ld32(taddr + 32u * j, v);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) x[32 * j + i] = fmaf(__uint_as_float(v[i]), p.oscale, s_bias[32 * j + i]);
          }
          sl = fmaf(__uint_as_float(vs), p.oscale, s_bias_sig);
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tmem_empty_bar[group]);  // the slot is free as soon as the logits are in registers
        const int amax = hf_label(x, p.c1, p.no_ties);
        const float sp = 1.f / (1.f + expf(-sl));
        if (pixel < p.m_total) {
          if (p.argmax) p.argmax[pixel] = static_cast<uint8_t>(amax);
          if (p.mask) p.mask[pixel] = sp >= p.threshold ? 1 : 0;
          if (p.sigp) p.sigp[pixel] = sp;
        }
      } else {
        // Training (unet.py:196-221 + the f1 / weighted-recall counts of unet.py:159-193 + d loss / d logits): three
        // passes over the accumulator in TMEM (max; soft-max sum; gradient) instead of 96 live logits per thread
        const bool ok = pixel < p.m_total;
        const int t = (ok && p.species) ? static_cast<int>(p.species[pixel]) : 0;
        float sl;
        {
          uint32_t vs;
          asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(vs) : "r"(taddr + static_cast<uint32_t>(p.c1)));
          tmem_ld_wait();
          sl = fmaf(__uint_as_float(vs), p.oscale, s_bias_sig);
        }
        float mxk[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
        for (int j = 0; j < 3; ++j) {
          uint32_t v[32];
          tmem_ld32(taddr + 32u * j, v);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i) mxk[i & 3] = fmaxf(mxk[i & 3], fmaf(__uint_as_float(v[i]), p.oscale, s_bias[32 * j + i]));
        }
        const float mx = fmaxf(fmaxf(mxk[0], mxk[1]), fmaxf(mxk[2], mxk[3]));
        float s[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f}, et = 0.f;
        int amax = 96;
#pragma unroll
        for (int j = 2; j >= 0; --j) {  // descending: the last assignment leaves the FIRST maximum of the logits
          uint32_t v[32];
          tmem_ld32(taddr + 32u * j, v);
          tmem_ld_wait();
#pragma unroll
          for (int i = 31; i >= 0; --i) {
            const int c = 32 * j + i;
            const float xv = fmaf(__uint_as_float(v[i]), p.oscale, s_bias[c]);
            const float e = c < p.c1 ? expf(xv - mx) : 0.f;
            s[i & 7] += e;
            et = c == t ? e : et;
            amax = xv == mx ? c : amax;
          }
        }
        const float inv = 1.f / (((s[0] + s[1]) + (s[2] + s[3])) + ((s[4] + s[5]) + (s[6] + s[7])));
        const float pt = et * inv;
        const bool in_range = pt >= 1e-7f && pt <= 1.f - 1e-7f;
        const float ptc = fminf(fmaxf(pt, 1e-7f), 1.f - 1e-7f);
        const float wt = p.class_w ? p.class_w[t < p.c1 ? t : 0] : 1.f;
        const float tb = t != 0 ? 1.f : 0.f;
        const float sp = 1.f / (1.f + expf(-sl));
        const float gs = in_range ? wt * p.inv_count : 0.f;
        if (ok) {
          acc[0] += static_cast<double>(-wt * logf(ptc));
          acc[1] += static_cast<double>(fmaxf(sl, 0.f) - sl * tb + log1pf(expf(-fabsf(sl))));
          cnt[0] += pt > 0.5f ? 1 : 0;
          cnt[1] += inv > 0.5f ? 1 : 0;  // K.round(p) == 1 for some class <=> the largest probability (e == 1: p = inv) > 0.5
          cnt[2] += (t != 0 && pt > 0.5f) ? 1 : 0;
          cnt[3] += t != 0 ? 1 : 0;
          if (p.argmax) p.argmax[pixel] = static_cast<uint8_t>(amax);
          if (p.sigp) p.sigp[pixel] = sp;
        }
        if (p.dlogits) {
          uint4* dst = reinterpret_cast<uint4*>(p.dlogits + pixel * p.ldd);
#pragma unroll
          for (int j = 0; j < 3; ++j) {
            uint32_t v[32];
            tmem_ld32(taddr + 32u * j, v);
            tmem_ld_wait();
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              float g[8];
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                const int c = 32 * j + 8 * q + i;
                const float xv = fmaf(__uint_as_float(v[8 * q + i]), p.oscale, s_bias[c]);
                const float pj = expf(xv - mx) * inv;
                g[i] = c < p.c1 ? gs * (pj - (c == t ? 1.f : 0.f)) : (c == p.c1 ? (sp - tb) * p.inv_count : 0.f);
              }
              uint4 qv;
              qv.x = pack_bf16x2(g[0], g[1]);
              qv.y = pack_bf16x2(g[2], g[3]);
              qv.z = pack_bf16x2(g[4], g[5]);
              qv.w = pack_bf16x2(g[6], g[7]);
              if (ok && 32 * j + 8 * q < p.ldd) dst[4 * j + q] = qv;
            }
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tmem_empty_bar[group]);
      }
    }
  }

  if constexpr (kTrain) {
    if (warp >= 2) {
#pragma unroll
      for (int i = 0; i < 6; ++i) {
        const double v = warp_sum(i < 2 ? acc[i] : static_cast<double>(cnt[i - 2]));
        if (lane == 0) s_red[warp - 2][i] = v;
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if constexpr (kTrain) {
    if (threadIdx.x < 6) {  // fixed order over the 8 epilogue warps: deterministic for a given grid
      double t = 0.0;
      for (int w = 0; w < 8; ++w) t += s_red[w][threadIdx.x];
      p.partials[static_cast<size_t>(blockIdx.x) * 6 + threadIdx.x] = t;
    }
  }
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 2 * kHfSlotCols);
  }
}

}  // namespace icsg3d

using namespace icsg3d;

static int heads_fused_launch(bool train, const void* x, int ldx, const void* wpack, const float* bias, int64_t M, int cin, int nout,
                              int c1, int op_f16, float out_scale, HeadsFusedParams p, void* stream) {
  ICSG_REQUIRE(x && wpack && M > 0, "heads_fused: bad arguments");
  ICSG_REQUIRE(cin >= 64 && cin % 64 == 0 && cin <= 384, "heads_fused: cin must be a multiple of 64 in [64,384] (got %d)", cin);
  ICSG_REQUIRE(nout >= 16 && nout % 16 == 0 && nout <= 96, "heads_fused: nout must be a multiple of 16 <= 96 (got %d)", nout);
  ICSG_REQUIRE(c1 >= 1 && c1 < nout, "heads_fused: c1 must be in [1, nout) (got %d)", c1);
  ICSG_REQUIRE(ldx % 8 == 0 && ldx >= cin, "heads_fused: ldx must be >= cin and a multiple of 8 (got %d)", ldx);
  ICSG_REQUIRE((reinterpret_cast<uintptr_t>(x) & 15) == 0 && (reinterpret_cast<uintptr_t>(wpack) & 15) == 0,
               "heads_fused: x and wpack must be 16-byte aligned");
  ICSG_REQUIRE((M + 127) / 128 < (1ll << 31) / 128, "heads_fused: too many voxels");
  const int sms = sm_count();
  if (sms <= 0) return cuda_fail(cudaGetLastError(), "sm_count", __FILE__, __LINE__);
  p.m_total = M;
  p.tiles_m = static_cast<int>((M + 127) / 128);
  p.chunks = cin / 64;
  p.nout = nout;
  p.c1 = c1;
  const uint32_t fmt_mask = op_f16 ? ~((7u << 7) | (7u << 10)) : ~0u;  // a/b format fields: BF16 (1) -> F16 (0)
  p.idesc = umma_idesc_bf16(nout, 0, 0) & fmt_mask;
  p.oscale = out_scale;
  static const int no_ties = [] { const char* e = getenv("ICSG3D_HEADS_NO_TIES"); return e ? atoi(e) : 0; }();
  p.no_ties = no_ties;
  p.bias = bias;
  const uint32_t w_bytes = static_cast<uint32_t>(p.chunks) * nout * 128u;
  int stages = static_cast<int>((200u * 1024u - w_bytes) / kHfChunkBytes);
  if (stages > kHfMaxStages) stages = kHfMaxStages;
  p.stages = stages;

  CUtensorMap tmA, tmB;
  {
    uint64_t dims[2] = {static_cast<uint64_t>(cin), static_cast<uint64_t>(M)};
    uint64_t strides[1] = {static_cast<uint64_t>(ldx) * 2};
    uint32_t box[2] = {64, 128};
    int rc = encode_tiled_bf16(&tmA, x, 2, dims, strides, box, 128);
    if (rc) return rc;
  }
  {
    uint64_t dims[2] = {static_cast<uint64_t>(cin), static_cast<uint64_t>(nout)};
    uint64_t strides[1] = {static_cast<uint64_t>(cin) * 2};
    uint32_t box[2] = {64, static_cast<uint32_t>(nout)};
    int rc = encode_tiled_bf16(&tmB, wpack, 2, dims, strides, box, 128);
    if (rc) return rc;
  }
  const size_t smem = 1024 + w_bytes + static_cast<size_t>(stages) * kHfChunkBytes;
  static bool configured = false;
  if (!configured) {
    ICSG_CUDA(cudaFuncSetAttribute(heads_predict_fused_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 210 * 1024));
    ICSG_CUDA(cudaFuncSetAttribute(heads_predict_fused_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 210 * 1024));
    configured = true;
  }
  const int grid = p.tiles_m < sms ? p.tiles_m : sms;
  if (train) launch_k(heads_predict_fused_kernel<true>, grid, kHfThreads, smem, static_cast<cudaStream_t>(stream), tmA, tmB, p);
  else launch_k(heads_predict_fused_kernel<false>, grid, kHfThreads, smem, static_cast<cudaStream_t>(stream), tmA, tmB, p);
  ICSG_CHECK_LAUNCH();
  return ICSG3D_OK;
}

extern "C" int icsg3d_heads_predict_fused(const void* x, int ldx, const void* wpack, const float* bias, int64_t M, int cin,
                                          int nout, int c1, int op_f16, float out_scale, float threshold,
                                          uint8_t* argmax_out, uint8_t* mask_out, float* sig_prob, void* stream) {
  ICSG_REQUIRE(argmax_out || mask_out || sig_prob, "heads_predict_fused: no output requested");
  HeadsFusedParams p{};
  p.threshold = threshold;
  p.argmax = argmax_out;
  p.mask = mask_out;
  p.sigp = sig_prob;
  return heads_fused_launch(false, x, ldx, wpack, bias, M, cin, nout, c1, op_f16, out_scale, p, stream);
}

// rows of the partials buffer icsg3d_heads_loss_fused writes (= its grid): feed them to icsg3d_heads_loss_finalize
extern "C" int icsg3d_heads_loss_fused_nparts(int64_t M) {
  if (M <= 0) return -1;
  const int sms = sm_count();
  const int64_t tiles = (M + 127) / 128;
  return static_cast<int>(tiles < sms ? tiles : (sms > 0 ? sms : 148));
}

extern "C" int icsg3d_heads_loss_fused(const void* x, int ldx, const void* wpack, const float* bias, int64_t M, int cin, int nout,
                                       int c1, const uint8_t* species, const float* class_w, float inv_count,
                                       double* partials, uint8_t* argmax_out, float* sig_prob, void* dlogits, int ldd,
                                       void* stream) {
  ICSG_REQUIRE(partials, "heads_loss_fused: partials required");
  ICSG_REQUIRE(!dlogits || (ldd % 8 == 0 && ldd > c1 && ldd <= 96 && (reinterpret_cast<uintptr_t>(dlogits) & 15) == 0),
               "heads_loss_fused: dlogits needs 8 | ldd in (c1, 96] and 16-byte alignment");
  HeadsFusedParams p{};
  p.argmax = argmax_out;
  p.sigp = sig_prob;
  p.species = species;
  p.class_w = class_w;
  p.inv_count = inv_count;
  p.dlogits = static_cast<__nv_bfloat16*>(dlogits);
  p.ldd = ldd;
  p.partials = partials;
  return heads_fused_launch(true, x, ldx, wpack, bias, M, cin, nout, c1, 0, 1.0f, p, stream);
}
