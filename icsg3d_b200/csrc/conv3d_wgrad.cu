// Conv3D 3x3x3 filter gradient (Conv3DBackpropFilterV2) on tcgen05 / TMEM (sm_100a).
//
//   dW[tap][ci][co] = sum over voxels v of  x[v + off(tap), ci] * dy[v, co]
//
// GEMM view (per 128-voxel tile, accumulated over all tiles a CTA owns):
//   D[(tap,ci) = 128 rows, co = NT cols] += A^T[128 x 16 voxels] * B[16 voxels x NT]     (8 k-steps / tile)
// Both operands are the NDHWC tiles exactly as TMA delivers them ([voxel][channel], channel contiguous),
// i.e. MN-major UMMA operands: no transposed copies of x or dy are ever materialised.
// The 128 MMA rows of one "group" are 128/KCA row-blocks, each row-block = one (tap, channel-chunk) and one
// TMA box (the tap shift and the "same" zero padding come from the box coordinates / OOB fill).
// Every group owns NT TMEM columns; a CTA keeps up to 512/NT groups resident for its whole voxel range,
// then dumps fp32 partials to the workspace; a second kernel reduces the voxel splits in a fixed order
// (deterministic: needed for the 1-rank vs N-rank data-parallel equivalence test).
#include <stdlib.h>
#include <string.h>

#include "common.cuh"

namespace icsg3d {

struct WgradParams {
  int m_total;
  int D, H, W;
  int tiles_m;
  int cin, cout;
  int ntaps;                   // 27 or 1
  int kca, chunks_a, bpg;      // channels per row-block, chunks per tap, row-blocks per group
  int total_blocks, total_groups;
  int groups_per_cta;
  int kcb, ntw, nb_boxes;      // B operand: channels per box, N tile, boxes per tile
  int splits;
  int stages;
  uint32_t a_box_bytes, b_box_bytes, a_bytes, b_bytes, stage_bytes;  // stage = the A row-blocks of one group; the dy tile has its own 2-slot ring
  uint32_t sbo_a, lbo_a, sbo_b, lbo_b, layout_a, layout_b, idesc, tmem_cols;
  float* ws;                   // [splits][27*cin*cout]
};

// warp 0: TMA producer | warps 1..kWgradIssuers: MMA issuers — accumulator group gl is owned by issuer gl % kWgradIssuers,
// so every accumulator sees its MMAs from one thread in a fixed order (deterministic) while the issue latencies of
// the issuers overlap | last 4 warps: epilogue
static constexpr int kWgradIssuers = 3;
static constexpr int kWgradThreads = (1 + kWgradIssuers + 4) * 32;
static constexpr int kWgradMaxStages = 6;

__global__ void __launch_bounds__(kWgradThreads, 1)
conv3d_k3_wgrad_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmDY,
                       const WgradParams p) {
  pdl_prologue();
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t full_bar[kWgradMaxStages];
  __shared__ __align__(8) uint64_t empty_bar[kWgradMaxStages];
  __shared__ __align__(8) uint64_t done_bar;
  __shared__ __align__(8) uint64_t b_full[2], b_empty[2];
  __shared__ uint32_t tmem_base_slot;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  // [dy tile slot 0 | dy tile slot 1 | A stage ring]: the dy tile of a voxel tile is loaded ONCE and shared by all the
  // accumulator groups of the CTA (it used to travel with every group's stage: half of the L2 -> smem traffic)
  const uint32_t bring_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* bring = smem_raw + (bring_base - smem_u32(smem_raw));
  const uint32_t ring_base = bring_base + 2u * p.b_bytes;
  uint8_t* ring = bring + 2u * p.b_bytes;

  const int split = blockIdx.x;
  const int g_begin = blockIdx.y * p.groups_per_cta;
  const int g_end = min(g_begin + p.groups_per_cta, p.total_groups);
  const int nz = blockIdx.z;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmX);
    tma_prefetch_desc(&tmDY);
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], kWgradIssuers);
    }
    mbar_init(&done_bar, kWgradIssuers);
    for (int s = 0; s < 2; ++s) {
      mbar_init(&b_full[s], 1);
      mbar_init(&b_empty[s], static_cast<uint32_t>(g_end - g_begin));  // one tcgen05.commit per group that read the slot
    }
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc(&tmem_base_slot, p.tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_slot;

  if (warp == 0) {
    // TMA producer: whole warp runs the loop, one elected lane issues
    const bool leader = elect_one();
    int stage = 0;
    uint32_t phase = 0;
    int ti = 0;
    for (int tile = split; tile < p.tiles_m; tile += p.splits, ++ti) {
      int pix = tile * 128;
      const int w0 = pix % p.W;
      pix /= p.W;
      const int h0 = pix % p.H;
      pix /= p.H;
      const int d0 = pix % p.D;
      const int n0 = pix / p.D;
      {
        const int bs = ti & 1;
        mbar_wait(&b_empty[bs], (static_cast<uint32_t>(ti >> 1) & 1u) ^ 1u);
        if (leader) {
          mbar_expect_tx(&b_full[bs], p.b_bytes);
          for (int j = 0; j < p.nb_boxes; ++j)
            tma_load_5d(bring + static_cast<size_t>(bs) * p.b_bytes + static_cast<size_t>(j) * p.b_box_bytes, &tmDY, &b_full[bs],
                        nz * p.ntw + j * p.kcb, w0, h0, d0, n0);
        }
      }
      for (int g = g_begin; g < g_end; ++g) {
        mbar_wait(&empty_bar[stage], phase ^ 1u);
        const int blk0 = g * p.bpg;
        int nblk = p.total_blocks - blk0;
        if (nblk > p.bpg) nblk = p.bpg;
        uint8_t* sa = ring + static_cast<size_t>(stage) * p.stage_bytes;
        if (leader) {
          mbar_expect_tx(&full_bar[stage], static_cast<uint32_t>(nblk) * p.a_box_bytes);
          for (int b = 0; b < nblk; ++b) {
            const int blk = blk0 + b;
            const int tap = blk / p.chunks_a;
            const int ch = blk - tap * p.chunks_a;
            int kd = tap / 9;
            int kh = (tap - kd * 9) / 3;
            int kw = tap - kd * 9 - kh * 3;
            if (p.ntaps == 1) kd = kh = kw = 1;
            tma_load_5d(sa + static_cast<size_t>(b) * p.a_box_bytes, &tmX, &full_bar[stage], ch * p.kca, w0 + kw - 1,
                        h0 + kh - 1, d0 + kd - 1, n0);
          }
        }
        if (++stage == p.stages) {
          stage = 0;
          phase ^= 1u;
        }
      }
    }
  } else if (warp <= kWgradIssuers) {
    // MMA issuers: whole warp runs the loop, one elected lane issues; descriptor high words are loop invariant
    const int issuer = warp - 1;
    const bool leader = elect_one();
    const int g_cnt = g_end - g_begin;
    const uint32_t a_hi = umma_desc_hi(p.sbo_a, p.layout_a), b_hi = umma_desc_hi(p.sbo_b, p.layout_b);
    const uint32_t ring_lo_a = umma_desc_lo(ring_base, p.lbo_a);
    const uint32_t ring_lo_b = umma_desc_lo(bring_base, p.lbo_b);
    const uint32_t bslot_lo = p.b_bytes >> 4;
    const uint32_t stage_lo = p.stage_bytes >> 4;
    const uint32_t ka = (2u * p.sbo_a) >> 4, kb = (2u * p.sbo_b) >> 4;
    // Every issuer walks the WHOLE stage sequence and waits on every full barrier in order (mbarrier parity waits are
    // only sound when a waiter observes each phase); it issues MMAs for the groups it owns and simply arrives on the
    // empty barrier for the others, so a stage is recycled after its owner's MMAs retired and all issuers moved on.
    int ti = 0, stage = 0;
    uint32_t phase = 0;
    for (int tile = split; tile < p.tiles_m; tile += p.splits, ++ti) {
      const int bs = ti & 1;
      mbar_wait(&b_full[bs], static_cast<uint32_t>(ti >> 1) & 1u);
      for (int gl = 0; gl < g_cnt; ++gl) {
        mbar_wait(&full_bar[stage], phase);
        if (gl % kWgradIssuers == issuer) {
          tc_fence_after();
          const uint32_t a_lo = ring_lo_a + static_cast<uint32_t>(stage) * stage_lo;
          const uint32_t b_lo = ring_lo_b + static_cast<uint32_t>(bs) * bslot_lo;
          const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(gl * p.ntw);
          if (leader) {
#pragma unroll
            for (int k = 0; k < 8; ++k)  // 8 x 16 voxels
              umma_bf16_lohi(d_tmem, a_lo + k * ka, a_hi, b_lo + k * kb, b_hi, p.idesc, (ti | k) != 0 ? 1u : 0u);
            umma_commit(&empty_bar[stage]);
            umma_commit(&b_empty[bs]);
          }
        } else if (leader) {
          mbar_arrive(&empty_bar[stage]);
        }
        if (++stage == p.stages) {
          stage = 0;
          phase ^= 1u;
        }
      }
    }
    if (leader) umma_commit(&done_bar);
  } else {
    const int quarter = warp & 3;
    const int row = quarter * 32 + lane;
    mbar_wait(&done_bar, 0);
    tc_fence_after();
    float* ws = p.ws + static_cast<size_t>(split) * p.ntaps * p.cin * p.cout;
    for (int g = g_begin; g < g_end; ++g) {
      const int blk = g * p.bpg + row / p.kca;
      const int tap = blk / p.chunks_a;
      const int ci = (blk - tap * p.chunks_a) * p.kca + row % p.kca;
      const bool ok = blk < p.total_blocks;
      const uint32_t taddr =
          tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + static_cast<uint32_t>((g - g_begin) * p.ntw);
      for (int c0 = 0; c0 < p.ntw; c0 += 16) {
        uint32_t v[16];
        tmem_ld16(taddr + static_cast<uint32_t>(c0), v);
        tmem_ld_wait();
        if (!ok) continue;
        float4* dst = reinterpret_cast<float4*>(ws + (static_cast<size_t>(tap) * p.cin + ci) * p.cout + nz * p.ntw + c0);
#pragma unroll
        for (int i = 0; i < 4; ++i)
          dst[i] = make_float4(__uint_as_float(v[4 * i]), __uint_as_float(v[4 * i + 1]), __uint_as_float(v[4 * i + 2]),
                               __uint_as_float(v[4 * i + 3]));
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, p.tmem_cols);
  }
}

// dw[i] = sum over splits of ws[s][i], in a FIXED order (deterministic): LANES interleaved split lanes per element (each
// with two independent accumulators), then the lane sums are added in a fixed tree.  Two shapes of the same kernel: many
// splits of a few-KB filter gradient (the VAE step: ~148 splits, latency bound) take 16 lanes x 16 elements so that the
// short strided columns stay in flight; few splits of a multi-MB gradient (the U-Net's 256..512-channel layers: 1..13
// splits of up to 28 MB, bandwidth bound) take 4 lanes x 64 elements so that the lanes are not idle.
template <int LANES, int ELEMS>
__global__ void __launch_bounds__(256) wgrad_reduce_kernel(const float* __restrict__ ws, float* __restrict__ dw, long long n,
                                                           int splits) {
  static_assert(LANES * ELEMS == 256, "one thread per (lane, element)");
  pdl_prologue();
  __shared__ float red[LANES][ELEMS + 1];
  const int e = threadIdx.x & (ELEMS - 1), sl = threadIdx.x / ELEMS;
  const long long i = static_cast<long long>(blockIdx.x) * ELEMS + e;
  float a0 = 0.f, a1 = 0.f;
  if (i < n) {
    int s = sl;
    for (; s + LANES < splits; s += 2 * LANES) {
      a0 += ws[static_cast<long long>(s) * n + i];
      a1 += ws[static_cast<long long>(s + LANES) * n + i];
    }
    if (s < splits) a0 += ws[static_cast<long long>(s) * n + i];
  }
  red[sl][e] = a0 + a1;
  __syncthreads();
  if (sl == 0 && i < n) {
    float t[LANES];
#pragma unroll
    for (int k = 0; k < LANES; ++k) t[k] = red[k][e];
#pragma unroll
    for (int w = LANES / 2; w > 0; w >>= 1)
#pragma unroll
      for (int k = 0; k < w; ++k) t[k] += t[k + w];
    dw[i] = t[0];
  }
}

// Few splits of a large gradient (the U-Net's 256..768-channel layers: up to 10.6 M elements, 2..13 splits): a plain
// streaming sum, one float4 per thread and grid-stride, splits added in index order.
__global__ void __launch_bounds__(256) wgrad_reduce_stream_kernel(const float4* __restrict__ ws, float4* __restrict__ dw,
                                                                  long long n4, int splits) {
  pdl_prologue();
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n4;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    float4 a = ws[i];
    for (int s = 1; s < splits; ++s) {
      const float4 b = ws[static_cast<long long>(s) * n4 + i];
      a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
    }
    dw[i] = a;
  }
}

// The shape is a function of the split count (and the alignment of the destination) only, and the split count of the layer
// shape only: a layer is always reduced in the same order (1 rank == N ranks, graph replay == eager).
void launch_wgrad_reduce(const float* ws, float* dw, long long n, int splits, cudaStream_t st) {
  if (splits >= 32) {
    launch_k(wgrad_reduce_kernel<16, 16>, static_cast<int>((n + 15) / 16), 256, 0, st, ws, dw, n, splits);
  } else if (n % 4 == 0 && ((reinterpret_cast<uintptr_t>(ws) | reinterpret_cast<uintptr_t>(dw)) & 15) == 0) {
    long long blocks = (n / 4 + 255) / 256;
    if (blocks > 148 * 8) blocks = 148 * 8;
    launch_k(wgrad_reduce_stream_kernel, static_cast<int>(blocks), 256, 0, st, reinterpret_cast<const float4*>(ws),
             reinterpret_cast<float4*>(dw), n / 4, splits);
  } else {
    launch_k(wgrad_reduce_kernel<4, 64>, static_cast<int>((n + 63) / 64), 256, 0, st, ws, dw, n, splits);
  }
}

int encode_act_map(CUtensorMap* map, const void* x, int ldx, int B, int D, int H, int W, int c_extent, int kc);
// conv3d_wgrad_stream.cu: plane-streaming variant for the narrow W >= 16 layers (returns < 0 when the shape does not fit)
int64_t wgrad_stream_workspace_bytes(int B, int D, int H, int W, int cin, int cout, int sms);
int wgrad_stream_run(const void* x, int ldx, const void* dy, int ldy, int B, int D, int H, int W, int cin, int cout, int sms,
                     float* ws, int64_t ws_bytes, int* splits_out, cudaStream_t st);

static bool wg_pow2(int v) { return v > 0 && (v & (v - 1)) == 0; }

static int wgrad_plan(int ntaps, int B, int D, int H, int W, int cin, int cout, WgradParams* p) {
  p->ntaps = ntaps;
  const long long m_total = static_cast<long long>(B) * D * H * W;
  p->m_total = static_cast<int>(m_total);
  p->D = D;
  p->H = H;
  p->W = W;
  p->tiles_m = static_cast<int>((m_total + 127) / 128);
  p->cin = cin;
  p->cout = cout;
  p->kca = (cin % 64 == 0) ? 64 : (cin % 32 == 0 ? 32 : 16);
  p->chunks_a = cin / p->kca;
  p->bpg = 128 / p->kca;
  p->total_blocks = ntaps * p->chunks_a;
  p->total_groups = (p->total_blocks + p->bpg - 1) / p->bpg;
  p->kcb = (cout % 64 == 0) ? 64 : (cout % 32 == 0 ? 32 : 16);
  p->ntw = cout < 128 ? cout : 128;
  if (cout % p->ntw != 0) p->ntw = p->kcb;
  p->nb_boxes = p->ntw / p->kcb;
  int gpc = 512 / p->ntw;
  if (gpc > p->total_groups) gpc = p->total_groups;
  p->groups_per_cta = gpc;
  p->a_box_bytes = 128u * p->kca * 2u;
  p->b_box_bytes = 128u * p->kcb * 2u;
  p->a_bytes = static_cast<uint32_t>(p->bpg) * p->a_box_bytes;  // 32 KB
  p->b_bytes = static_cast<uint32_t>(p->nb_boxes) * p->b_box_bytes;
  p->stage_bytes = p->a_bytes;
  int stages = static_cast<int>((200u * 1024u - 2u * p->b_bytes) / p->stage_bytes);
  if (stages > kWgradMaxStages) stages = kWgradMaxStages;
  p->stages = stages;
  p->sbo_a = 8u * p->kca * 2u;
  p->lbo_a = p->a_box_bytes;
  p->sbo_b = 8u * p->kcb * 2u;
  p->lbo_b = p->b_box_bytes;
  p->layout_a = umma_layout_for_swizzle(p->kca * 2);
  p->layout_b = umma_layout_for_swizzle(p->kcb * 2);
  p->idesc = umma_idesc_bf16(p->ntw, 1, 1);
  uint32_t cols = 32;
  while (cols < static_cast<uint32_t>(gpc * p->ntw)) cols <<= 1;
  p->tmem_cols = cols;
  const int gy = (p->total_groups + gpc - 1) / gpc;
  const int gz = cout / p->ntw;
  int sms = sm_count();
  if (sms <= 0) sms = 148;
  int splits = sms / (gy * gz);
  if (splits < 1) splits = 1;
  if (splits > p->tiles_m) splits = p->tiles_m;
  p->splits = splits;
  return ICSG3D_OK;
}

}  // namespace icsg3d

using namespace icsg3d;

static int64_t wgrad_workspace_impl(int ntaps, int B, int D, int H, int W, int cin, int cout) {
  if (B <= 0 || cin <= 0 || cout <= 0 || cin % 16 || cout % 16) return -1;
  WgradParams p{};
  wgrad_plan(ntaps, B, D, H, W, cin, cout, &p);
  int64_t need = static_cast<int64_t>(p.splits) * ntaps * cin * cout * 4;
  if (ntaps == 27) {
    int sms = sm_count();
    if (sms <= 0) sms = 148;
    const int64_t alt = wgrad_stream_workspace_bytes(B, D, H, W, cin, cout, sms);
    if (alt > need) need = alt;
  }
  return need;
}
extern "C" int64_t icsg3d_conv3d_k3_wgrad_workspace(int B, int D, int H, int W, int cin, int cout) {
  return wgrad_workspace_impl(27, B, D, H, W, cin, cout);
}
extern "C" int64_t icsg3d_conv3d_k1_wgrad_workspace(int B, int D, int H, int W, int cin, int cout) {
  return wgrad_workspace_impl(1, B, D, H, W, cin, cout);
}

// ICSG3D_WGRAD_IMPL=v1 forces the per-tap kernel (A/B comparisons); icsg3d_conv3d_set_impl(1|2) does the same at run time.
namespace icsg3d {
int g_wgrad_impl = -1;
}
static int wgrad_impl_choice() {
  if (g_wgrad_impl < 0) {
    const char* e = getenv("ICSG3D_WGRAD_IMPL");
    g_wgrad_impl = (e && strcmp(e, "v1") == 0) ? 1 : 0;
  }
  return g_wgrad_impl;
}

static int wgrad_impl(int ntaps, const void* x, int ldx, const void* dy, int ldy, float* dw, int B, int D, int H, int W,
                      int cin, int cout, void* workspace, int64_t workspace_bytes, void* stream) {
  ICSG_REQUIRE(x && dy && dw && workspace, "conv3d_k3_wgrad: null pointer");
  ICSG_REQUIRE(B > 0 && wg_pow2(D) && wg_pow2(H) && wg_pow2(W) && D >= 2 && H >= 2 && W >= 2 && W <= 128,
               "conv3d_k3_wgrad: D,H,W must be powers of two in [2,128]");
  ICSG_REQUIRE(cin % 16 == 0 && cout % 16 == 0 && cin >= 16 && cout >= 16,
               "conv3d_k3_wgrad: cin/cout must be multiples of 16 (got %d/%d)", cin, cout);
  ICSG_REQUIRE(ldx % 8 == 0 && ldy % 8 == 0 && ldx >= cin && ldy >= cout, "conv3d_k3_wgrad: bad ldx/ldy");
  const long long n_dw = static_cast<long long>(ntaps) * cin * cout;
  if (ntaps == 27 && wgrad_impl_choice() == 0) {
    int sms = sm_count();
    if (sms <= 0) sms = 148;
    int splits = 0;
    const int rc = wgrad_stream_run(x, ldx, dy, ldy, B, D, H, W, cin, cout, sms, static_cast<float*>(workspace), workspace_bytes,
                                    &splits, static_cast<cudaStream_t>(stream));
    if (rc == ICSG3D_OK) {
      launch_wgrad_reduce(static_cast<const float*>(workspace), dw, n_dw, splits, static_cast<cudaStream_t>(stream));
      ICSG_CHECK_LAUNCH();
      return ICSG3D_OK;
    }
    if (rc != 1) return rc;  // 1 = shape not eligible: fall through to the per-tap kernel
  }
  WgradParams p{};
  wgrad_plan(ntaps, B, D, H, W, cin, cout, &p);
  ICSG_REQUIRE(p.stages >= 2, "conv3d_k3_wgrad: stage too large");
  const int64_t need = static_cast<int64_t>(p.splits) * ntaps * cin * cout * 4;
  ICSG_REQUIRE(workspace_bytes >= need, "conv3d_k3_wgrad: workspace too small (%lld < %lld)",
               static_cast<long long>(workspace_bytes), static_cast<long long>(need));
  // a single split IS the gradient: written in place (16-byte stores), no reduction pass
  const bool direct = p.splits == 1 && (reinterpret_cast<uintptr_t>(dw) & 15) == 0;
  p.ws = direct ? dw : static_cast<float*>(workspace);

  CUtensorMap tmX, tmDY;
  int rc = encode_act_map(&tmX, x, ldx, B, D, H, W, cin, p.kca);
  if (rc) return rc;
  rc = encode_act_map(&tmDY, dy, ldy, B, D, H, W, cout, p.kcb);
  if (rc) return rc;

  static bool configured = false;
  if (!configured) {
    ICSG_CUDA(cudaFuncSetAttribute(conv3d_k3_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 210 * 1024));
    configured = true;
  }
  const size_t smem = 2u * p.b_bytes + static_cast<size_t>(p.stages) * p.stage_bytes + 1024;
  dim3 grid(p.splits, (p.total_groups + p.groups_per_cta - 1) / p.groups_per_cta, cout / p.ntw);
  launch_k(conv3d_k3_wgrad_kernel, grid, kWgradThreads, smem, static_cast<cudaStream_t>(stream), tmX, tmDY, p);
  ICSG_CHECK_LAUNCH();
  if (!direct) launch_wgrad_reduce(p.ws, dw, n_dw, p.splits, static_cast<cudaStream_t>(stream));
  ICSG_CHECK_LAUNCH();
  return ICSG3D_OK;
}

extern "C" int icsg3d_conv3d_k3_wgrad(const void* x, int ldx, const void* dy, int ldy, float* dw, int B, int D, int H,
                                      int W, int cin, int cout, void* workspace, int64_t workspace_bytes,
                                      void* stream) {
  return wgrad_impl(27, x, ldx, dy, ldy, dw, B, D, H, W, cin, cout, workspace, workspace_bytes, stream);
}

extern "C" int icsg3d_conv3d_k1_wgrad(const void* x, int ldx, const void* dy, int ldy, float* dw, int B, int D, int H,
                                      int W, int cin, int cout, void* workspace, int64_t workspace_bytes,
                                      void* stream) {
  return wgrad_impl(1, x, ldx, dy, ldy, dw, B, D, H, W, cin, cout, workspace, workspace_bytes, stream);
}
