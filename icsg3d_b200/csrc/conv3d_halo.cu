// Conv3D 3x3x3 "same" — halo-reuse implicit GEMM on tcgen05 / TMEM (sm_100a).  Second-generation kernel
// for layers with W >= 8: the per-tap re-fetch of the A operand (27x L2->SM traffic in conv3d_igemm.cu)
// is replaced by ONE TMA load of a halo'd activation block; every tap is the same block read through a
// UMMA descriptor whose start address is shifted by a whole number of rows.
//
//   * work item = output region TD x TH x W (full width) of one sample (x one N tile);
//   * A block in smem = (TD+2) x (TH+2) x (W+1) voxels x KC channels per channel chunk, row pitch KC*2 bytes
//     (hardware swizzle = row pitch), delivered by one 5-D tiled TMA box per chunk starting at
//     (w,h,d) = (-1, h0-1, d0-1): out-of-bounds zero fill = "same" padding.  The row pitch in W is W+1:
//     column 0 of every row is the zero pad w=-1 and doubles as the pad w=W of the previous row;
//   * flat row index f = (dl*HP + hl)*WP + wl addresses the block; output voxel (dl,hl,wl) has its
//     (kd,kh,kw) input at f + kd*HP*WP + kh*WP + kw, so an M tile of 128 consecutive f and a tap shift are
//     both plain row offsets of the descriptor start address (measured on B200: K-major swizzled
//     descriptors accept arbitrary row-shifted starts, profiles/r01_conv_probe_bringup.jsonl "shift");
//     rows with hl >= TH or wl == W are junk and are skipped by the epilogue;
//   * loop order is tap-outer: each weight tile B[tap][N x KC] streams through a small TMA ring once per
//     work item and is applied to all G = ceil(TD*HP*WP/128) M tiles, whose accumulators live side by side
//     in TMEM (G*NT <= 512 columns);
//   * two accumulator sets (2*G*NT <= 512 TMEM columns) alternate between consecutive items, so the epilogue
//     of item i (TMEM -> bias/activation -> global) overlaps all MMAs of item i+1; the A block is double
//     buffered when shared memory allows, and when all 27*chunks weight tiles fit beside it they are loaded
//     once and stay resident (every 32^3 layer of the VAE+DFC step).
#include "common.cuh"
#include "conv3d_halo.cuh"

namespace icsg3d {



// warp 0: TMA producer | warps 1..kHaloIssuers: MMA issuers (accumulator g is owned by issuer g % kHaloIssuers;
// measured: one issuing thread sustains only ~100-150 cycles per tcgen05.mma because of its own uniform-register
// dependency chains, vs a 55-cycle tensor floor at N <= 64 — profiles/r01_halo_pattern_probe.json) | last 4 warps: epilogue
static constexpr int kHaloIssuers = 3;
static constexpr int kHaloEpiWarps = 8;  // 2 per TMEM lane quarter: the epilogue is latency bound (TMEM round trips)
static constexpr int kHaloThreads = (1 + kHaloIssuers + kHaloEpiWarps) * 32;
static constexpr int kHaloMaxG = 32;
static constexpr int kHaloMaxBStages = 8;

template <int KSTEPS>
__global__ void __launch_bounds__(kHaloThreads, 1)
conv3d_k3_halo_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                      const ConvHaloParams p) {
  pdl_prologue();
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t a_full[2], a_empty[2];
  __shared__ __align__(8) uint64_t b_full[kHaloMaxBStages], b_empty[kHaloMaxBStages];
  __shared__ __align__(8) uint64_t acc_full[2], acc_empty[2];
  __shared__ __align__(8) uint64_t b_all_full;
  __shared__ float s_bias[512];
  __shared__ float s_stats[2][512];  // per-CTA BatchNorm partials of the stored output (fused statistics)
  __shared__ uint32_t tmem_base_slot;

  // warp index through a shuffle: warp-uniform for the compiler, keeps the MMA descriptors in uniform registers
  const int warp = __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* sm = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t b_ring_off = static_cast<uint32_t>(p.a_bufs) * p.a_buf_bytes;

  // Zero the guard row that follows every A chunk block: it is the w=W pad of the block's last row.
  {
    const int words = p.row_bytes / 4;
    for (int i = threadIdx.x; i < p.a_bufs * p.chunks * words; i += blockDim.x) {
      const int buf = i / (p.chunks * words);
      const int ch = (i / words) % p.chunks;
      const int w = i % words;
      uint32_t* dst = reinterpret_cast<uint32_t*>(sm + static_cast<size_t>(buf) * p.a_buf_bytes +
                                                  static_cast<size_t>(ch) * p.a_chunk_bytes +
                                                  static_cast<size_t>(p.block_rows) * p.row_bytes);
      dst[w] = 0u;
    }
    fence_proxy_async();
  }
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&a_full[i], 1);
      mbar_init(&a_empty[i], kHaloIssuers);
    }
    for (int s = 0; s < (p.b_resident ? 0 : p.b_stages); ++s) {
      mbar_init(&b_full[s], 1);
      mbar_init(&b_empty[s], kHaloIssuers);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&acc_full[i], kHaloIssuers);
      mbar_init(&acc_empty[i], kHaloEpiWarps);
    }
    mbar_init(&b_all_full, 1);
    fence_mbar_init();
  }
  for (int i = threadIdx.x; i < p.nt * p.tiles_n && i < 512; i += blockDim.x) s_bias[i] = p.bias ? p.bias[i] : 0.f;
  for (int i = threadIdx.x; i < 1024; i += blockDim.x) (&s_stats[0][0])[i] = 0.f;
  if (p.post_scale != nullptr) {  // inference: the statistics table carries the post-activation affine instead
    for (int i = threadIdx.x; i < p.nt * p.tiles_n && i < 512; i += blockDim.x) {
      s_stats[0][i] = p.post_scale[i];
      s_stats[1][i] = p.post_shift[i];
    }
  }
  if (warp == 1) tmem_alloc(&tmem_base_slot, p.tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_slot;

  if (warp == 0) {
    // ===================== TMA producer (whole warp runs the loop; one elected lane issues) =====================
    const bool leader = elect_one();
    int bstage = 0;
    uint32_t bphase = 0;
    int it = 0;
    for (int item = blockIdx.x; item < p.total_items; item += gridDim.x, ++it) {
      int t = item;
      const int tn = t % p.tiles_n;
      t /= p.tiles_n;
      const int hb = t % p.n_hblk;
      t /= p.n_hblk;
      const int db = t % p.n_dblk;
      const int n = t / p.n_dblk;
      const int buf = it % p.a_bufs;
      const uint32_t aph = static_cast<uint32_t>(it / p.a_bufs) & 1u;
      mbar_wait(&a_empty[buf], aph ^ 1u);
      if (leader) {
        mbar_expect_tx(&a_full[buf], p.a_tx_bytes);
        for (int ch = 0; ch < p.chunks; ++ch)
          tma_load_5d(sm + static_cast<size_t>(buf) * p.a_buf_bytes + static_cast<size_t>(ch) * p.a_chunk_bytes, &tmA,
                      &a_full[buf], ch * p.kc, -1, hb * p.TH - 1, db * p.TD - 1, n);
      }
      if (p.b_resident) {
        if (it == 0 && leader) {  // tiles_n == 1 in resident mode: one load of all weight tiles for the whole kernel
          mbar_expect_tx(&b_all_full, static_cast<uint32_t>(27 * p.chunks) * p.b_unit_bytes);
          for (int tap = 0; tap < 27; ++tap)
            for (int ch = 0; ch < p.chunks; ++ch)
              tma_load_3d(sm + b_ring_off + static_cast<size_t>(tap * p.chunks + ch) * p.b_unit_bytes, &tmB, &b_all_full,
                          ch * p.kc, 0, tap);
        }
        continue;
      }
      for (int tap = 0; tap < 27; ++tap) {
        for (int ch = 0; ch < p.chunks; ++ch) {
          mbar_wait(&b_empty[bstage], bphase ^ 1u);
          if (leader) {
            mbar_expect_tx(&b_full[bstage], p.b_unit_bytes);
            tma_load_3d(sm + b_ring_off + static_cast<size_t>(bstage) * p.b_unit_bytes, &tmB, &b_full[bstage], ch * p.kc,
                        tn * p.nt, tap);
          }
          if (++bstage == p.b_stages) {
            bstage = 0;
            bphase ^= 1u;
          }
        }
      }
    }
  } else if (warp <= kHaloIssuers) {
    // ===================== MMA issuers (whole warp runs the loop; one elected lane issues) =====================
    const int issuer = warp - 1;
    const bool leader = elect_one();
    int bstage = 0;
    uint32_t bphase = 0;
    int it = 0;
    const uint32_t desc_hi = umma_desc_hi(p.sbo, p.layout);
    const uint32_t tile_lo = (128u * static_cast<uint32_t>(p.row_bytes)) >> 4;  // descriptor-lo step between M tiles
    const uint32_t b_ring_lo = umma_desc_lo(base + b_ring_off, 16u);
    const uint32_t b_unit_lo = p.b_unit_bytes >> 4;
    for (int item = blockIdx.x; item < p.total_items; item += gridDim.x, ++it) {
      const int buf = it % p.a_bufs;
      const uint32_t aph = static_cast<uint32_t>(it / p.a_bufs) & 1u;
      const int set = it & 1;
      const uint32_t sph = static_cast<uint32_t>(it >> 1) & 1u;
      mbar_wait(&a_full[buf], aph);
      if (p.b_resident && it == 0) mbar_wait(&b_all_full, 0);
      mbar_wait(&acc_empty[set], sph ^ 1u);
      tc_fence_after();
      const uint32_t a_buf_lo = umma_desc_lo(base + static_cast<uint32_t>(buf) * p.a_buf_bytes, 16u);
      const uint32_t tmem_set = tmem_base + static_cast<uint32_t>(set * p.G * p.nt);
      bool first = true;
      int unit = 0;
      for (int kd = 0; kd < 3; ++kd) {
        for (int kh = 0; kh < 3; ++kh) {
          for (int kw = 0; kw < 3; ++kw) {
            const uint32_t shift_lo = (static_cast<uint32_t>(kd * p.plane_rows + kh * p.WP + kw) * static_cast<uint32_t>(p.row_bytes)) >> 4;
            for (int ch = 0; ch < p.chunks; ++ch, ++unit) {
              uint32_t b_lo;
              if (p.b_resident) {
                b_lo = b_ring_lo + static_cast<uint32_t>(unit) * b_unit_lo;
              } else {
                mbar_wait(&b_full[bstage], bphase);
                tc_fence_after();
                b_lo = b_ring_lo + static_cast<uint32_t>(bstage) * b_unit_lo;
              }
              uint32_t a_lo = a_buf_lo + static_cast<uint32_t>(ch) * (p.a_chunk_bytes >> 4) + shift_lo +
                              static_cast<uint32_t>(issuer) * tile_lo;
              uint32_t d_tmem = tmem_set + static_cast<uint32_t>(issuer * p.nt);
              if (leader) {
                for (int g = issuer; g < p.G; g += kHaloIssuers) {
#pragma unroll
                  for (int k = 0; k < KSTEPS; ++k)
                    umma_bf16_lohi(d_tmem, a_lo + 2u * k, desc_hi, b_lo + 2u * k, desc_hi, p.idesc, (!first || k != 0) ? 1u : 0u);
                  a_lo += kHaloIssuers * tile_lo;
                  d_tmem += static_cast<uint32_t>(kHaloIssuers * p.nt);
                }
              }
              first = false;
              if (!p.b_resident) {
                if (leader) umma_commit(&b_empty[bstage]);
                if (++bstage == p.b_stages) {
                  bstage = 0;
                  bphase ^= 1u;
                }
              }
            }
          }
        }
      }
      if (leader) {
        umma_commit(&a_empty[buf]);
        umma_commit(&acc_full[set]);
      }
    }
  } else {
    // ===================== epilogue warps: 2 per TMEM lane quarter, the (tile g, 32-column group) items alternate =====
    const int quarter = warp & 3;
    const int half = (warp - 1 - kHaloIssuers) >> 2;
    const int ngrp = (p.nt + 31) >> 5;  // 32-column groups per tile (the last one may be 16 wide)
    const float slope = p.act == ICSG3D_ACT_RELU ? 0.f : (p.act == ICSG3D_ACT_LEAKY ? p.alpha : 1.f);
    int it = 0;
    for (int item = blockIdx.x; item < p.total_items; item += gridDim.x, ++it) {
      int t = item;
      const int tn = t % p.tiles_n;
      t /= p.tiles_n;
      const int hb = t % p.n_hblk;
      t /= p.n_hblk;
      const int db = t % p.n_dblk;
      const int n = t / p.n_dblk;
      const int set = it & 1;
      mbar_wait(&acc_full[set], static_cast<uint32_t>(it >> 1) & 1u);
      tc_fence_after();
      const int nwork = p.G * ngrp;
      for (int wk = half; wk < nwork; wk += 2) {
        const int g = wk / ngrp;
        const int c0 = (wk - g * ngrp) << 5;
        const int f = g * 128 + quarter * 32 + lane;
        const int dl = f / p.plane_rows;
        const int rem = f - dl * p.plane_rows;
        const int hl = rem / p.WP;
        const int wl = rem - hl * p.WP;
        const int d = db * p.TD + dl;
        const bool ok = dl < p.TD && hl < p.TH && wl < p.W && d < p.D;
        const long long pixel = ((static_cast<long long>(n) * p.D + d) * p.H + hb * p.TH + hl) * p.W + wl;
        const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) +
                               static_cast<uint32_t>((set * p.G + g) * p.nt + c0);
        const int ncol = min(32, p.nt - c0);
        uint32_t v[32];
        if (ncol == 32) {
          tmem_ld32(taddr, v);
        } else {
          uint32_t v16[16];
          tmem_ld16(taddr, v16);
#pragma unroll
          for (int i = 0; i < 16; ++i) v[i] = v16[i];
#pragma unroll
          for (int i = 16; i < 32; ++i) v[i] = 0u;
        }
        tmem_ld_wait();
        const int col0 = tn * p.nt + c0;
        if (col0 >= p.n_store) continue;        // warp-uniform
        if (!ok && !p.stats) continue;
        float fv[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          const float x = fmaf(__uint_as_float(v[i]), p.oscale, s_bias[(col0 + i) & 511]);
          fv[i] = fmaxf(x, slope * x);
        }
        if (p.post_scale != nullptr) {
#pragma unroll
          for (int i = 0; i < 32; ++i) fv[i] = fmaf(fv[i], s_stats[0][(col0 + i) & 511], s_stats[1][(col0 + i) & 511]);
        }
        const int nvalid = min(ncol, p.n_store - col0);
        if (p.y_dtype == ICSG3D_DT_BF16) {
          __nv_bfloat16* dst = reinterpret_cast<__nv_bfloat16*>(p.y) + pixel * p.ldy + col0;
          if ((p.ldy & 7) == 0 && (nvalid & 7) == 0) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              if (8 * j < nvalid) {
                uint4 qv;
                qv.x = pack_bf16x2(fv[8 * j], fv[8 * j + 1]);
                qv.y = pack_bf16x2(fv[8 * j + 2], fv[8 * j + 3]);
                qv.z = pack_bf16x2(fv[8 * j + 4], fv[8 * j + 5]);
                qv.w = pack_bf16x2(fv[8 * j + 6], fv[8 * j + 7]);
                if (ok) reinterpret_cast<uint4*>(dst)[j] = qv;
                if (p.stats) {  // statistics of the values as stored (bf16-rounded)
                  float2 u;
                  u = unpack_bf16x2(qv.x); fv[8 * j] = u.x; fv[8 * j + 1] = u.y;
                  u = unpack_bf16x2(qv.y); fv[8 * j + 2] = u.x; fv[8 * j + 3] = u.y;
                  u = unpack_bf16x2(qv.z); fv[8 * j + 4] = u.x; fv[8 * j + 5] = u.y;
                  u = unpack_bf16x2(qv.w); fv[8 * j + 6] = u.x; fv[8 * j + 7] = u.y;
                }
              }
            }
          } else {
#pragma unroll
            for (int i = 0; i < 32; ++i) {
              if (i < nvalid) {
                const __nv_bfloat16 q = f2bf(fv[i]);
                if (ok) dst[i] = q;
                fv[i] = bf2f(q);
              }
            }
          }
        } else if (ok) {
          float* dst = reinterpret_cast<float*>(p.y) + pixel * p.ldy + col0;
          if ((p.ldy & 3) == 0 && (nvalid & 3) == 0) {
#pragma unroll
            for (int j = 0; j < 8; ++j)
              if (4 * j < nvalid)
                reinterpret_cast<float4*>(dst)[j] = make_float4(fv[4 * j], fv[4 * j + 1], fv[4 * j + 2], fv[4 * j + 3]);
          } else {
#pragma unroll
            for (int i = 0; i < 32; ++i)
              if (i < nvalid) dst[i] = fv[i];
          }
        }
        if (p.stats) {
          // column sums over the warp's 32 rows by a transpose-reduce butterfly (junk rows contribute zero), then one
          // shared-memory atomic per column and warp: the CTA's partial, written out after the last item
#pragma unroll
          for (int j = 0; j < 32; j += 16) {
            if (j < nvalid) {  // warp-uniform
              float a[16], b[16];
#pragma unroll
              for (int i = 0; i < 16; ++i) {
                const float t = ok ? fv[j + i] : 0.f;
                a[i] = t;
                b[i] = t * t;
              }
              const float s1 = warp_colsum16(a, lane), s2 = warp_colsum16(b, lane);
              if ((lane & 1) == 0) {
                const int c = (col0 + j + colsum16_owner(lane)) & 511;
                atomicAdd(&s_stats[0][c], s1);
                atomicAdd(&s_stats[1][c], s2);
              }
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc_empty[set]);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (p.stats) {
    const int nout = p.nt * p.tiles_n;
    for (int i = threadIdx.x; i < 2 * nout; i += blockDim.x) {
      const int half = i / nout, c = i - half * nout;
      p.stats[static_cast<size_t>(blockIdx.x) * 2 * nout + i] = static_cast<double>(s_stats[half][c]);
    }
  }
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, p.tmem_cols);
  }
}

static bool halo_pow2(int v) { return v > 0 && (v & (v - 1)) == 0; }

// Autotuning hook (tools/halo_autotune.py): when set, only the forced (TD, TH, NT) is considered.
static int g_halo_force[3] = {0, 0, 0};
void conv_halo_force(int td, int th, int nt) {
  g_halo_force[0] = td;
  g_halo_force[1] = th;
  g_halo_force[2] = nt;
}

// Pick (TD, TH, NT) minimising a simple time-per-output model; returns false when no configuration fits.
bool conv_halo_plan(int B, int D, int H, int W, int cin, int nout, int sms, ConvHaloParams* out) {
  if (W < 8 || W > 64 || !halo_pow2(W) || !halo_pow2(H) || !halo_pow2(D)) return false;
  const int kc = (cin % 64 == 0) ? 64 : (cin % 32 == 0 ? 32 : 16);
  const int chunks = cin / kc;
  const int row_bytes = kc * 2;
  const int WP = W + 1;
  const uint32_t smem_budget = 216u * 1024u;
  double best = 1e30;
  bool found = false;
  ConvHaloParams bp{};
  for (int nt = (nout < 256 ? nout : 256); nt >= 16; nt -= 16) {
    if (nout % nt != 0) continue;
    if (nt < 64 && nt != nout) continue;  // never split a narrow layer
    for (int TH = H; TH >= 2; TH /= 2) {
      const int HP = TH + 2;
      const int plane_rows = HP * WP;
      for (int TD = 1; TD <= 8 && TD <= D; ++TD) {
        if (g_halo_force[0] > 0 && (TD != g_halo_force[0] || TH != g_halo_force[1] || nt != g_halo_force[2])) continue;
        const int out_rows = TD * plane_rows;
        const int G = (out_rows + 127) / 128;
        if (G > kHaloMaxG || 2 * G * nt > 512) continue;  // two accumulator sets
        const int block_rows = (TD + 2) * plane_rows;
        const int rows_alloc = G * 128 + 2 * plane_rows + 2 * WP + 3;
        const uint32_t a_chunk = (static_cast<uint32_t>((rows_alloc > block_rows + 1 ? rows_alloc : block_rows + 1)) * row_bytes + 1023u) & ~1023u;
        const uint32_t a_buf = a_chunk * chunks;
        const uint32_t b_unit = static_cast<uint32_t>(nt) * row_bytes;
        if (a_buf + 3 * b_unit > smem_budget) continue;
        const uint32_t b_all = 27u * chunks * b_unit;
        int a_bufs, b_stages, b_resident = 0;
        if (nt == nout && nout <= 512 && 2 * a_buf + b_all <= smem_budget) {
          a_bufs = 2; b_resident = 1; b_stages = 27 * chunks;
        } else if (nt == nout && nout <= 512 && a_buf + b_all <= smem_budget) {
          a_bufs = 1; b_resident = 1; b_stages = 27 * chunks;
        } else {
          a_bufs = (2 * a_buf + 4 * b_unit <= smem_budget) ? 2 : 1;
          b_stages = static_cast<int>((smem_budget - a_bufs * a_buf) / b_unit);
          if (b_stages > kHaloMaxBStages) b_stages = kHaloMaxBStages;
          if (b_stages < 3) continue;
        }
        const int n_dblk = (D + TD - 1) / TD;
        const long long items = static_cast<long long>(B) * n_dblk * (H / TH) * (nout / nt);
        // cycles per item: MMA pipe (A smem read floor ~32 clk / MMA, N/2 clk tensor floor) vs L2->SM fill
        // measured on B200 (profiles/r01_mma_rate_probe.json): SS-mode tcgen05.mma K=16 costs max(54.7, N/2) cycles
        const double mma_clk = static_cast<double>(G) * 27 * chunks * (kc / 16) * (nt / 2 > 55 ? nt / 2 : 55);
        const double a_rows = static_cast<double>(block_rows) * chunks;
        const double fill_bytes = a_rows * row_bytes + (b_resident ? 0.0 : 27.0 * chunks * nt * row_bytes);
        double fill_clk = fill_bytes / 28.0;
        if (fill_clk < a_rows * 2.5) fill_clk = a_rows * 2.5;
        double t = (a_bufs == 2) ? (mma_clk > fill_clk ? mma_clk : fill_clk) : (mma_clk + a_rows * row_bytes / 28.0);
        t += 2000.0;  // per-item fixed overhead (barrier round trips, accumulator drain)
        const double waves = static_cast<double>((items + sms - 1) / sms);
        const double score = waves * t;  // modelled cycles for the whole layer
        if (score < best) {
          best = score;
          found = true;
          bp = ConvHaloParams{};
          bp.B = B; bp.D = D; bp.H = H; bp.W = W;
          bp.TD = TD; bp.TH = TH; bp.HP = HP; bp.WP = WP;
          bp.plane_rows = plane_rows; bp.block_rows = block_rows; bp.out_rows = out_rows; bp.G = G;
          bp.kc = kc; bp.chunks = chunks; bp.row_bytes = row_bytes;
          bp.nt = nt; bp.tiles_n = nout / nt;
          bp.n_dblk = n_dblk; bp.n_hblk = H / TH;
          bp.total_items = static_cast<int>(items);
          bp.a_bufs = a_bufs; bp.b_stages = b_stages; bp.b_resident = b_resident;
          bp.a_chunk_bytes = a_chunk; bp.a_buf_bytes = a_buf;
          bp.a_tx_bytes = static_cast<uint32_t>(block_rows) * row_bytes * chunks;
          bp.b_unit_bytes = b_unit;
        }
      }
    }
  }
  if (!found) return false;
  bp.sbo = 8u * bp.row_bytes;
  bp.layout = umma_layout_for_swizzle(bp.row_bytes);
  bp.idesc = umma_idesc_bf16(bp.nt, 0, 0);
  uint32_t cols = 32;
  while (cols < static_cast<uint32_t>(2 * bp.G * bp.nt)) cols <<= 1;
  bp.tmem_cols = cols;
  *out = bp;
  return true;
}

int launch_conv_halo(const void* x, int ldx, const void* wpack, const float* bias, void* y, int ldy, int y_dtype,
                     int n_store, int cin, int nout, int act, float alpha, ConvHaloParams p, int sms, cudaStream_t st,
                     float oscale, double* stats, const float* post_scale, const float* post_shift) {
  CUtensorMap tmA, tmB;
  {
    uint64_t dims[5] = {static_cast<uint64_t>(cin), static_cast<uint64_t>(p.W), static_cast<uint64_t>(p.H),
                        static_cast<uint64_t>(p.D), static_cast<uint64_t>(p.B)};
    uint64_t strides[4] = {static_cast<uint64_t>(ldx) * 2, static_cast<uint64_t>(p.W) * ldx * 2,
                           static_cast<uint64_t>(p.H) * p.W * ldx * 2, static_cast<uint64_t>(p.D) * p.H * p.W * ldx * 2};
    uint32_t box[5] = {static_cast<uint32_t>(p.kc), static_cast<uint32_t>(p.WP), static_cast<uint32_t>(p.HP),
                       static_cast<uint32_t>(p.TD + 2), 1};
    int rc = encode_tiled_bf16(&tmA, x, 5, dims, strides, box, p.row_bytes);
    if (rc) return rc;
  }
  {
    uint64_t dims[3] = {static_cast<uint64_t>(cin), static_cast<uint64_t>(nout), 27};
    uint64_t strides[2] = {static_cast<uint64_t>(cin) * 2, static_cast<uint64_t>(nout) * cin * 2};
    uint32_t box[3] = {static_cast<uint32_t>(p.kc), static_cast<uint32_t>(p.nt), 1};
    int rc = encode_tiled_bf16(&tmB, wpack, 3, dims, strides, box, p.row_bytes);
    if (rc) return rc;
  }
  p.y = y; p.ldy = ldy; p.y_dtype = y_dtype; p.n_store = n_store; p.bias = bias; p.act = act; p.alpha = alpha;
  p.oscale = oscale;
  p.stats = stats;
  p.post_scale = post_scale;
  p.post_shift = post_shift;
  static bool configured = false;
  if (!configured) {
    ICSG_CUDA(cudaFuncSetAttribute(conv3d_k3_halo_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
    ICSG_CUDA(cudaFuncSetAttribute(conv3d_k3_halo_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
    ICSG_CUDA(cudaFuncSetAttribute(conv3d_k3_halo_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
    configured = true;
  }
  const size_t smem = static_cast<size_t>(p.a_bufs) * p.a_buf_bytes + static_cast<size_t>(p.b_stages) * p.b_unit_bytes + 1024;
  const int grid = conv_halo_grid(p, sms);
  if (p.kc == 16) launch_k(conv3d_k3_halo_kernel<1>, grid, kHaloThreads, smem, st, tmA, tmB, p);
  else if (p.kc == 32) launch_k(conv3d_k3_halo_kernel<2>, grid, kHaloThreads, smem, st, tmA, tmB, p);
  else launch_k(conv3d_k3_halo_kernel<4>, grid, kHaloThreads, smem, st, tmA, tmB, p);
  ICSG_CHECK_LAUNCH();
  return ICSG3D_OK;
}

}  // namespace icsg3d
