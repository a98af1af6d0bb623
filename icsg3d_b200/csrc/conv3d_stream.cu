// Conv3D 3x3x3 "same" — plane-streaming implicit GEMM with the kd taps folded into the MMA N dimension
// (tcgen05 / TMEM / TMA, sm_100a).  Third-generation kernel for the narrow layers (Cout <= 64) at W >= 16,
// which carry most of the VAE+DFC step (c1, c2, enc_conv1/2, dec_conv3/4, decoder_output and their
// dgrads).  Replaces the same reference ops as conv3d_igemm.cu (Keras Conv3D + BiasAdd [+ReLU/LeakyReLU],
// Conv3DBackpropInput): vae/lattice_vae.py:173,178,213,219-224; unet/unet.py:276-336.
//
// Why: measured on B200 (profiles/r01_mma_rate_probe.json) a tcgen05.mma (SS, bf16, K=16, M=128) costs
// max(54.7, N/2) cycles — the 128x32 B A-operand read from shared memory is the floor — so an N = Cout <= 64
// layer runs the tensor pipe at <= 58 %, and the tap-outer halo kernel also re-loads every input plane for each
// of the 3 output planes it feeds.  Here:
//   * a CTA walks a column (sample n, h-block hb) of the volume along d; every INPUT plane slab
//     (TH+2) x (W+1) voxels x Cin is TMA-loaded ONCE into a shared-memory ring (out-of-bounds zero fill = "same"
//     padding in h and w; the d padding planes are simply never issued);
//   * input plane i contributes to output planes i-1, i, i+1 through kd = 2, 1, 0.  The accumulators of
//     consecutive output planes sit side by side in TMEM columns (a ring of R slots), so ONE MMA with
//     N = 3*Cout and the weight tile [kd=2 | kd=1 | kd=0] x Cout rows updates all three: 3x fewer MMA
//     instructions and A-operand reads, N = 48..192 instead of 16..64;
//   * the (kh, kw) taps are row-shifted UMMA descriptors into the same slab (as in conv3d_halo.cu);
//   * all 27 weight tiles stay resident in shared memory (loaded once per CTA);
//   * the first MMA that touches a fresh slot is split off with accumulate = 0; a ring wrap splits the MMA in two;
//   * the epilogue drains output plane d (TMEM -> bias/activation -> global) while the MMAs of the following
//     planes run; optional per-channel sum / sum-of-squares of the stored values (BatchNorm statistics).
//   * work = B * n_hblk * D plane steps, split evenly over the CTAs (a CTA may start mid-column: it re-reads
//     one neighbour plane on each side of the cut).
#include "conv3d_stream.cuh"

namespace icsg3d {

static constexpr int kStreamMaxIssuers = 3;  // 1 + 3 + 8 = 12 warps = 3 per SM sub-partition: up to 168 registers per thread
static constexpr int kStreamDefaultIssuers = 3;
static constexpr int kStreamEpiWarps = 8;
static constexpr int kStreamThreads = (1 + kStreamMaxIssuers + kStreamEpiWarps) * 32;
static constexpr int kStreamMaxStages = 6;
static constexpr int kStreamMaxR = 8;

// One epilogue item: NC accumulator columns of 32 rows (TMEM -> bias/activation -> global, + BatchNorm running sums).
// NC = 32 gives the warp two independent 16-column chains per TMEM round trip (the epilogue is latency bound).
// tcgen05.wait::ld that also names the loaded registers as in/out operands: no use of them can be scheduled above it
template <int NC>
__device__ __forceinline__ void tmem_ld_wait_regs(uint32_t (&v)[NC]) {
  if constexpr (NC == 32) {
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(v[0]), "+r"(v[1]), "+r"(v[2]), "+r"(v[3]), "+r"(v[4]), "+r"(v[5]), "+r"(v[6]), "+r"(v[7]),
                   "+r"(v[8]), "+r"(v[9]), "+r"(v[10]), "+r"(v[11]), "+r"(v[12]), "+r"(v[13]), "+r"(v[14]), "+r"(v[15]),
                   "+r"(v[16]), "+r"(v[17]), "+r"(v[18]), "+r"(v[19]), "+r"(v[20]), "+r"(v[21]), "+r"(v[22]), "+r"(v[23]),
                   "+r"(v[24]), "+r"(v[25]), "+r"(v[26]), "+r"(v[27]), "+r"(v[28]), "+r"(v[29]), "+r"(v[30]), "+r"(v[31])
                 :
                 : "memory");
  } else {
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(v[0]), "+r"(v[1]), "+r"(v[2]), "+r"(v[3]), "+r"(v[4]), "+r"(v[5]), "+r"(v[6]), "+r"(v[7]),
                   "+r"(v[8]), "+r"(v[9]), "+r"(v[10]), "+r"(v[11]), "+r"(v[12]), "+r"(v[13]), "+r"(v[14]), "+r"(v[15])
                 :
                 : "memory");
  }
}
template <int NC>
__device__ __forceinline__ void tmem_ld_cols(uint32_t taddr, uint32_t (&v)[NC]) {
  if constexpr (NC == 32) tmem_ld32(taddr, v);
  else tmem_ld16(taddr, v);
}

// One drained accumulator item (32 lanes x NC columns, already in registers): scale + bias + activation, store,
// BatchNorm running sums.
template <int NC, bool kPost>
__device__ __forceinline__ void stream_epi_math(const ConvStreamParams& p, const uint32_t (&v)[NC], int c0, bool ok, long long pixel,
                                                const float* s_bias, float slope, float (&sacc)[32], float (&qacc)[32],
                                                const float* s_post) {
  const int nvalid = min(NC, p.n_store - c0);
  const bool bf16 = p.y_dtype == ICSG3D_DT_BF16;
  const bool vec = nvalid == NC && (p.ldy & (bf16 ? 7 : 3)) == 0;
  const bool stats = p.stats != nullptr && ok;
  __nv_bfloat16* dst16 = reinterpret_cast<__nv_bfloat16*>(p.y) + pixel * p.ldy + c0;
  float* dst32 = reinterpret_cast<float*>(p.y) + pixel * p.ldy + c0;
  // 8 columns at a time (few live temporaries: the next item's TMEM load occupies 32 registers meanwhile)
#pragma unroll
  for (int j = 0; j < NC / 8; ++j) {
    float fv[8];
#pragma unroll
    for (int i = 0; i < 8; i += 4) {
      const float4 b4 = *reinterpret_cast<const float4*>(&s_bias[c0 + 8 * j + i]);
      const float x0 = fmaf(__uint_as_float(v[8 * j + i]), p.oscale, b4.x), x1 = fmaf(__uint_as_float(v[8 * j + i + 1]), p.oscale, b4.y);
      const float x2 = fmaf(__uint_as_float(v[8 * j + i + 2]), p.oscale, b4.z), x3 = fmaf(__uint_as_float(v[8 * j + i + 3]), p.oscale, b4.w);
      fv[i] = fmaxf(x0, slope * x0);
      fv[i + 1] = fmaxf(x1, slope * x1);
      fv[i + 2] = fmaxf(x2, slope * x2);
      fv[i + 3] = fmaxf(x3, slope * x3);
      if constexpr (kPost) {  // inference BatchNorm affine after the activation
        const float4 s4 = *reinterpret_cast<const float4*>(&s_post[c0 + 8 * j + i]);
        const float4 t4 = *reinterpret_cast<const float4*>(&s_post[64 + c0 + 8 * j + i]);
        fv[i] = fmaf(fv[i], s4.x, t4.x);
        fv[i + 1] = fmaf(fv[i + 1], s4.y, t4.y);
        fv[i + 2] = fmaf(fv[i + 2], s4.z, t4.z);
        fv[i + 3] = fmaf(fv[i + 3], s4.w, t4.w);
      }
    }
    if (bf16) {
      uint4 qv;
      qv.x = pack_bf16x2(fv[0], fv[1]);
      qv.y = pack_bf16x2(fv[2], fv[3]);
      qv.z = pack_bf16x2(fv[4], fv[5]);
      qv.w = pack_bf16x2(fv[6], fv[7]);
      if (ok) {
        if (vec) {
          reinterpret_cast<uint4*>(dst16)[j] = qv;
        } else {
#pragma unroll
          for (int i = 0; i < 8; ++i)
            if (8 * j + i < nvalid) dst16[8 * j + i] = f2bf(fv[i]);
        }
      }
      if (stats) {  // statistics of the values as stored (bf16-rounded)
        float2 u;
        u = unpack_bf16x2(qv.x); fv[0] = u.x; fv[1] = u.y;
        u = unpack_bf16x2(qv.y); fv[2] = u.x; fv[3] = u.y;
        u = unpack_bf16x2(qv.z); fv[4] = u.x; fv[5] = u.y;
        u = unpack_bf16x2(qv.w); fv[6] = u.x; fv[7] = u.y;
      }
    } else if (ok) {
      if (vec) {
        reinterpret_cast<float4*>(dst32)[2 * j] = make_float4(fv[0], fv[1], fv[2], fv[3]);
        reinterpret_cast<float4*>(dst32)[2 * j + 1] = make_float4(fv[4], fv[5], fv[6], fv[7]);
      } else if (nvalid == 4 && (p.ldy & 3) == 0) {
        if (j == 0) reinterpret_cast<float4*>(dst32)[0] = make_float4(fv[0], fv[1], fv[2], fv[3]);
      } else {
#pragma unroll
        for (int i = 0; i < 8; ++i)
          if (8 * j + i < nvalid) dst32[8 * j + i] = fv[i];
      }
    }
    if (stats) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        sacc[8 * j + i] += fv[i];
        qacc[8 * j + i] = fmaf(fv[i], fv[i], qacc[8 * j + i]);
      }
    }
  }
}

template <int KSTEPS, bool kPost>
__global__ void __launch_bounds__(kStreamThreads, 1)
conv3d_k3_stream_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                        const ConvStreamParams p) {
  pdl_prologue();
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t a_full[kStreamMaxStages], a_empty[kStreamMaxStages];
  __shared__ __align__(8) uint64_t slot_full[kStreamMaxR], slot_empty[kStreamMaxR];
  __shared__ __align__(8) uint64_t w_full;
  __shared__ __align__(16) float s_bias[64];
  __shared__ __align__(16) float s_post[128];  // [0,64) post-activation scale, [64,128) shift (inference BatchNorm)
  __shared__ double s_stats[2][64];
  __shared__ uint32_t tmem_base_slot;

  // warp index through a shuffle: tells the compiler it is warp-uniform, so that everything derived from it (MMA
  // descriptors, TMEM addresses) stays in uniform registers — the MMA issue loop is instruction bound otherwise
  const int warp = __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* sm = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t a_off = (p.w_bytes + 1023u) & ~1023u;  // weights first, then the plane ring

  // zero the guard row behind every plane slab chunk: it is the w = W pad of the slab's last row
  {
    const int words = p.row_bytes / 4;
    const int slab_rows = p.HP * p.WP;
    for (int i = threadIdx.x; i < p.stages * p.chunks * words; i += blockDim.x) {
      const int st = i / (p.chunks * words);
      const int ch = (i / words) % p.chunks;
      const int w = i % words;
      uint32_t* dst = reinterpret_cast<uint32_t*>(sm + a_off + static_cast<size_t>(st) * p.a_stage_bytes +
                                                  static_cast<size_t>(ch) * p.a_chunk_bytes +
                                                  static_cast<size_t>(slab_rows) * p.row_bytes);
      dst[w] = 0u;
    }
    fence_proxy_async();
  }
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int i = 0; i < p.stages; ++i) {
      mbar_init(&a_full[i], 1);
      mbar_init(&a_empty[i], p.issuers);
    }
    for (int i = 0; i < p.R; ++i) {
      mbar_init(&slot_full[i], p.issuers);
      mbar_init(&slot_empty[i], kStreamEpiWarps);
    }
    mbar_init(&w_full, 1);
    fence_mbar_init();
  }
  const int n_off = static_cast<int>(blockIdx.y) * p.C;  // N split: this CTA's slice of the output channels
  if (threadIdx.x < 64) s_bias[threadIdx.x] = (p.bias && threadIdx.x < p.C) ? p.bias[n_off + threadIdx.x] : 0.f;
  if (kPost && threadIdx.x < 64) {
    s_post[threadIdx.x] = threadIdx.x < p.C ? p.post_scale[n_off + threadIdx.x] : 1.f;
    s_post[64 + threadIdx.x] = threadIdx.x < p.C ? p.post_shift[n_off + threadIdx.x] : 0.f;
  }
  if (threadIdx.x < 128) s_stats[threadIdx.x >> 6][threadIdx.x & 63] = 0.0;
  if (warp == 1) tmem_alloc(&tmem_base_slot, p.tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_slot;

  const bool dbg_on = p.dbg != nullptr && blockIdx.x == 1 && blockIdx.y == 0;  // CTA 1: a full-length interior segment
  const int s_begin = blockIdx.x * p.steps_per_cta;
  const int s_end = min(p.total_steps, s_begin + p.steps_per_cta);

  if (warp == 0) {
    // ===================== TMA producer =====================
    const bool leader = elect_one();
    if (leader) {  // all 27 * chunks weight tiles, laid out [kh*3+kw][chunk][kd = 2, 1, 0][C rows]
      mbar_expect_tx(&w_full, p.w_bytes);
      for (int khw = 0; khw < 9; ++khw)
        for (int ch = 0; ch < p.chunks; ++ch)
          for (int blk = 0; blk < 3; ++blk)
            tma_load_3d(sm + static_cast<size_t>((khw * p.chunks + ch) * 3 + blk) * p.blk_bytes, &tmB, &w_full, ch * p.kc, n_off,
                        (2 - blk) * 9 + khw);
    }
    int astep = 0;
    for (int s = s_begin; s < s_end;) {
      const int col = s / p.D, db = s - col * p.D;
      const int de = min(p.D, db + (s_end - s));
      const int n = col / p.n_hblk, hb = col - n * p.n_hblk;
      const int i_lo = max(0, db - 1), i_hi = min(p.D - 1, de);
      for (int i = i_lo; i <= i_hi; ++i, ++astep) {
        const int stage = astep & (p.stages - 1);
        mbar_wait(&a_empty[stage], (static_cast<uint32_t>(astep >> p.st_log2) & 1u) ^ 1u);
        if (leader) {
          if (dbg_on && astep < p.dbg_steps) p.dbg[(0 * p.dbg_steps + astep) * 4] = clock64();
          mbar_expect_tx(&a_full[stage], p.a_tx_bytes);
          for (int ch = 0; ch < p.chunks; ++ch)
            tma_load_5d(sm + a_off + static_cast<size_t>(stage) * p.a_stage_bytes + static_cast<size_t>(ch) * p.a_chunk_bytes,
                        &tmA, &a_full[stage], ch * p.kc, -1, hb * p.TH - 1, i, n);
        }
      }
      s += de - db;
    }
  } else if (warp <= kStreamMaxIssuers) {
    // ===================== MMA issuers (tile t is owned by issuer t % issuers) =====================
    const int issuer = warp - 1;
    if (issuer < p.issuers) {
      const bool leader = elect_one();
      const uint32_t desc_hi = umma_desc_hi(p.sbo, p.layout);
      const uint32_t w_lo = umma_desc_lo(base, 16u);
      const uint32_t blk_lo = p.blk_bytes >> 4;
      const uint32_t unit_lo = 3u * blk_lo;
      const uint32_t chunk_lo = p.a_chunk_bytes >> 4;
      const uint32_t row_lo = static_cast<uint32_t>(p.row_bytes) >> 4;
      const uint32_t a_ring_lo = umma_desc_lo(base + a_off, 16u);
      const uint32_t tile_lo = 128u * row_lo;
      mbar_wait(&w_full, 0);
      int astep = 0, qbase = 0;
      for (int s = s_begin; s < s_end;) {
        const int col = s / p.D, db = s - col * p.D;
        const int de = min(p.D, db + (s_end - s));
        const int hb = col % p.n_hblk;
        const int th_valid = min(p.TH, p.H - hb * p.TH);
        const int t_valid = (th_valid * p.WP + 127) >> 7;  // tiles that hold at least one real output row
        const int i_lo = max(0, db - 1), i_hi = min(p.D - 1, de);
        for (int i = i_lo; i <= i_hi; ++i, ++astep) {
          const int o_lo = max(db, i - 1), o_hi = min(de - 1, i + 1);
          const int o_f = (i == 0) ? o_lo : min(i + 1, o_hi + 1);  // outputs [o_f, o_hi] are touched for the first time
          const int rmask = p.R - 1;
          const bool dbg_i = dbg_on && issuer == 0 && leader && astep < p.dbg_steps;
          long long* dbg_row = dbg_i ? p.dbg + (1 * p.dbg_steps + astep) * 4 : nullptr;
          if (dbg_i) dbg_row[0] = clock64();
          for (int o = o_f; o <= o_hi; ++o) {
            const int q = qbase + o - db;
            mbar_wait(&slot_empty[q & rmask], static_cast<uint32_t>(q >> p.r_log2) & 1u);  // zeroed by the epilogue
          }
          if (dbg_i) dbg_row[1] = clock64();
          const int stage = astep & (p.stages - 1);
          mbar_wait(&a_full[stage], static_cast<uint32_t>(astep >> p.st_log2) & 1u);
          tc_fence_after();
          if (dbg_i) dbg_row[2] = clock64();
          // outputs o_lo..o_hi = consecutive ring positions; split once at the ring wrap: op0 (n0 blocks), op1 (n1 blocks)
          const int q_lo = qbase + o_lo - db;
          const int cnt = o_hi - o_lo + 1;
          const int s0 = q_lo & rmask;
          const int n0 = cnt < p.R - s0 ? cnt : p.R - s0;
          const int n1 = cnt - n0;
          const int blk0 = o_lo - (i - 1);
          const uint32_t d0 = static_cast<uint32_t>(s0 * p.C);
          const uint32_t b0 = static_cast<uint32_t>(blk0) * blk_lo;
          const uint32_t b1 = static_cast<uint32_t>(blk0 + n0) * blk_lo;
          const uint32_t idesc0 = p.idesc[n0 - 1];
          const uint32_t idesc1 = p.idesc[n1 > 0 ? n1 - 1 : 0];
          const uint32_t a_stage_lo = a_ring_lo + static_cast<uint32_t>(stage) * (p.a_stage_bytes >> 4);
          const uint32_t wp_lo = static_cast<uint32_t>(p.WP) * row_lo;
          if (leader) {
            for (int t = issuer; t < t_valid; t += p.issuers) {
              const uint32_t a_t = a_stage_lo + static_cast<uint32_t>(t) * tile_lo;
              const uint32_t d_t = tmem_base + static_cast<uint32_t>(t * p.R * p.C);
              const uint32_t da = d_t + d0;
              const uint32_t wa = w_lo + b0, wb = w_lo + b1;
              // every slot was zeroed by the epilogue when it was drained: all MMAs accumulate.  Unit (0, 0, 0):
#pragma unroll
              for (int k = 0; k < KSTEPS; ++k) {
                umma_bf16_lohi(da, a_t + 2u * k, desc_hi, wa + 2u * k, desc_hi, idesc0, 1u);
                if (n1 > 0) umma_bf16_lohi(d_t, a_t + 2u * k, desc_hi, wb + 2u * k, desc_hi, idesc1, 1u);
              }
              if (p.chunks == 1) {
                // single channel chunk (Cin <= 64): the other 8 taps as one flat unrolled sequence
#pragma unroll
                for (int tap = 1; tap < 9; ++tap) {
                  const uint32_t a_u = a_t + static_cast<uint32_t>(tap / 3) * wp_lo + static_cast<uint32_t>(tap % 3) * row_lo;
                  const uint32_t bo = static_cast<uint32_t>(tap) * unit_lo;
#pragma unroll
                  for (int k = 0; k < KSTEPS; ++k) umma_bf16_lohi(da, a_u + 2u * k, desc_hi, wa + bo + 2u * k, desc_hi, idesc0, 1u);
                  if (n1 > 0) {
#pragma unroll
                    for (int k = 0; k < KSTEPS; ++k)
                      umma_bf16_lohi(d_t, a_u + 2u * k, desc_hi, wb + bo + 2u * k, desc_hi, idesc1, 1u);
                  }
                }
              } else {
                uint32_t bo = 0;
                for (int kh = 0; kh < 3; ++kh) {
#pragma unroll
                  for (int kw = 0; kw < 3; ++kw) {
                    uint32_t a_u = a_t + static_cast<uint32_t>(kh) * wp_lo + static_cast<uint32_t>(kw) * row_lo;
                    for (int ch = 0; ch < p.chunks; ++ch, a_u += chunk_lo, bo += unit_lo) {
                      if ((kh | kw | ch) == 0) continue;  // done above
#pragma unroll
                      for (int k = 0; k < KSTEPS; ++k) umma_bf16_lohi(da, a_u + 2u * k, desc_hi, wa + bo + 2u * k, desc_hi, idesc0, 1u);
                      if (n1 > 0) {
#pragma unroll
                        for (int k = 0; k < KSTEPS; ++k)
                          umma_bf16_lohi(d_t, a_u + 2u * k, desc_hi, wb + bo + 2u * k, desc_hi, idesc1, 1u);
                      }
                    }
                  }
                }
              }
            }
            umma_commit(&a_empty[stage]);
            if (dbg_i) dbg_row[3] = clock64();
            if (i - 1 >= db) umma_commit(&slot_full[(qbase + i - 1 - db) & rmask]);            // output plane i-1 is complete
            if (i == p.D - 1 && de == p.D) umma_commit(&slot_full[(qbase + i - db) & rmask]);  // and the last plane of the column
          }
          __syncwarp();
        }
        qbase += de - db;
        s += de - db;
      }
    }
  } else {
    // ===================== epilogue warps: 2 per TMEM lane quarter =====================
    // C = 64: warp half h owns columns [32h, 32h+32) of every tile; C <= 32: the halves take alternate tiles.
    const int quarter = warp & 3;
    const int half = (warp - 1 - kStreamMaxIssuers) >> 2;
    const bool split_cols = p.C == 64;
    const int c0 = split_cols ? 32 * half : 0;
    const int t_step = split_cols ? 1 : 2;
    const uint32_t wp_magic = ((1u << 20) + static_cast<uint32_t>(p.WP) - 1u) / static_cast<uint32_t>(p.WP);
    ConvStreamParams q_ = p;  // this CTA's view of the output: channels [n_off, n_off + C)
    q_.n_store = p.n_store - n_off;
    q_.y = static_cast<char*>(p.y) + static_cast<size_t>(n_off) * (p.y_dtype == ICSG3D_DT_BF16 ? 2 : 4);
    const ConvStreamParams& pe = q_;
    const bool active = c0 < pe.n_store;
    // LeakyReLU / ReLU / identity as max(x, slope * x)
    const float slope = p.act == ICSG3D_ACT_RELU ? 0.f : (p.act == ICSG3D_ACT_LEAKY ? p.alpha : 1.f);
    // BatchNorm statistics of the stored values: per-thread fp32 running sums for this warp's (at most 32) columns,
    // reduced across lanes once at the end of the kernel
    float sacc[32], qacc[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) sacc[i] = qacc[i] = 0.f;
    // The accumulator slots are kept ZERO between uses (so every MMA accumulates and the first K step of a plane needs no
    // overwrite split): zero the whole allocation once, then every drained item is zeroed right after it was read.
    for (uint32_t cz = static_cast<uint32_t>(half) * 32u; cz < p.tmem_cols; cz += 64u)
      tmem_st_zero<32>(tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + cz);
    tmem_st_wait();
    tc_fence_before();
    __syncwarp();
    if (lane == 0)
      for (int i = 0; i < p.R; ++i) mbar_arrive(&slot_empty[i]);
    int q = 0;
    for (int s = s_begin; s < s_end;) {
      const int col = s / p.D, db = s - col * p.D;
      const int de = min(p.D, db + (s_end - s));
      const int n = col / p.n_hblk, hb = col - n * p.n_hblk;
      const int th_valid = min(p.TH, p.H - hb * p.TH);
      const int t_valid = (th_valid * p.WP + 127) >> 7;
      for (int o = db; o < de; ++o, ++q) {
        const int slot = q & (p.R - 1);
        const long long plane0 = ((static_cast<long long>(n) * p.D + o) * p.H + hb * p.TH) * p.W;
        const bool dbg_e = dbg_on && quarter == 0 && lane == 0 && q < p.dbg_steps;
        long long* dbg_row = dbg_e ? p.dbg + ((2 + half) * p.dbg_steps + q) * 4 : nullptr;
        if (dbg_e) dbg_row[0] = clock64();
        mbar_wait(&slot_full[slot], static_cast<uint32_t>(q >> p.r_log2) & 1u);
        tc_fence_after();
        if (dbg_e) dbg_row[1] = clock64();
        // This warp's tiles of the plane, software-pipelined: the TMEM load of the next tile is in flight while the
        // current one is converted and stored; a drained item is zeroed as soon as its load has completed.
        // C <= 32: the two warps of a quarter take alternate tiles, alternating with the plane as well (odd T balances).
        const uint32_t tq = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + static_cast<uint32_t>(slot * p.C + c0);
        const uint32_t t_stride = static_cast<uint32_t>(p.R * p.C);
        int t = split_cols ? 0 : ((half + q) & 1);
        auto coords = [&](int tt, bool& ok, long long& pixel) {
          const uint32_t f = static_cast<uint32_t>(tt * 128 + quarter * 32 + lane);
          const uint32_t hl = (f * wp_magic) >> 20;  // f / WP (exact for f < 4096, WP < 130)
          const uint32_t wl = f - hl * static_cast<uint32_t>(p.WP);
          ok = static_cast<int>(hl) < th_valid && static_cast<int>(wl) < p.W;
          pixel = plane0 + static_cast<long long>(hl * static_cast<uint32_t>(p.W) + wl);
        };
        if (!active) {
          for (; t < t_valid; t += t_step) {
            if (p.C >= 32) tmem_st_zero<32>(tq + t * t_stride);
            else tmem_st_zero<16>(tq + t * t_stride);
          }
        } else if (p.C >= 32) {
          uint32_t v[32];
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = 0u;
          if (t < t_valid && !(p.exp_flags & 2)) tmem_ld_cols<32>(tq + t * t_stride, v);
          while (t < t_valid) {
            const int tn = t + t_step;
            uint32_t cur[32];
            tmem_ld_wait_regs<32>(v);
#pragma unroll
            for (int i = 0; i < 32; ++i) cur[i] = v[i];
            if (!(p.exp_flags & 1)) tmem_st_zero<32>(tq + t * t_stride);
            if (tn < t_valid && !(p.exp_flags & 2)) tmem_ld_cols<32>(tq + tn * t_stride, v);
            bool ok;
            long long pixel;
            coords(t, ok, pixel);
            if (p.exp_flags & 4) ok = false;
            stream_epi_math<32, kPost>(pe, cur, c0, ok, pixel, s_bias, slope, sacc, qacc, s_post);
            t = tn;
          }
        } else {
          uint32_t v[16];
          if (t < t_valid) tmem_ld_cols<16>(tq + t * t_stride, v);
          while (t < t_valid) {
            const int tn = t + t_step;
            uint32_t cur[16];
            tmem_ld_wait_regs<16>(v);
#pragma unroll
            for (int i = 0; i < 16; ++i) cur[i] = v[i];
            tmem_st_zero<16>(tq + t * t_stride);
            if (tn < t_valid) tmem_ld_cols<16>(tq + tn * t_stride, v);
            bool ok;
            long long pixel;
            coords(t, ok, pixel);
            stream_epi_math<16, kPost>(pe, cur, c0, ok, pixel, s_bias, slope, sacc, qacc, s_post);
            t = tn;
          }
        }
        if (dbg_e) dbg_row[2] = clock64();
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&slot_empty[slot]);
        if (dbg_e) dbg_row[3] = clock64();
      }
      s += de - db;
    }
    if (p.stats && active) {
      const int ncol = p.C >= 32 ? 32 : 16;
      for (int j = 0; j < ncol; j += 16) {
        float a[16], b[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          a[i] = j == 0 ? sacc[i] : sacc[16 + i];
          b[i] = j == 0 ? qacc[i] : qacc[16 + i];
        }
        const float s1 = warp_colsum16(a, lane), s2 = warp_colsum16(b, lane);
        if ((lane & 1) == 0) {
          const int c = c0 + j + colsum16_owner(lane);
          atomicAdd(&s_stats[0][c], static_cast<double>(s1));
          atomicAdd(&s_stats[1][c], static_cast<double>(s2));
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (p.stats && threadIdx.x < 2 * p.C) {
    const int half = threadIdx.x / p.C, c = threadIdx.x - half * p.C;
    p.stats[static_cast<size_t>(blockIdx.x) * 2 * p.C + half * p.C + c] = s_stats[half][c];
  }
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, p.tmem_cols);
  }
}

// MMA issuer warps used (tiles are dealt round-robin); ICSG3D_STREAM_ISSUERS = 1..3 overrides for A/B timing
static int stream_max_issuers() {
  static const int v = [] {
    const char* e = getenv("ICSG3D_STREAM_ISSUERS");
    const int n = e ? atoi(e) : kStreamDefaultIssuers;
    return n < 1 ? 1 : (n > kStreamMaxIssuers ? kStreamMaxIssuers : n);
  }();
  return v;
}

static bool stream_pow2(int v) { return v > 0 && (v & (v - 1)) == 0; }

static int g_stream_force_nb = 0;
void conv_stream_force_nb(int nb) { g_stream_force_nb = nb; }

bool conv_stream_plan(int B, int D, int H, int W, int cin, int nout, int sms, ConvStreamParams* out) {
  if (W < 16 || W > 64 || !stream_pow2(W) || !stream_pow2(H) || !stream_pow2(D) || D < 4) return false;
  if (cin % 16 != 0 || cin > 256) return false;
  // channels per CTA (= per kd block): the whole layer when its 27 weight tiles fit in shared memory, else 32-channel
  // slices (N split over blockIdx.y: the input is streamed once per slice, N = 96 per MMA)
  int C = 0;
  if ((nout == 16 || nout == 32 || nout == 64) && 27u * cin * nout * 2u + 64u * 1024u <= 222u * 1024u) C = nout;
  else if (nout == 64 && cin <= 64 && 27u * cin * 32u * 2u + 64u * 1024u <= 222u * 1024u) C = 32;  // c3; measured: 4 slices
  // of a 128-wide layer (c4 fprop: 91 us) lose to the halo kernel (81 us), so wider layers stay there
  if (C == 0) return false;
  const int kc = (cin % 64 == 0) ? 64 : (cin % 32 == 0 ? 32 : 16);
  const int chunks = cin / kc;
  const int row_bytes = kc * 2;
  const int WP = W + 1;
  const uint32_t budget = 222u * 1024u;
  const uint32_t w_bytes = 27u * static_cast<uint32_t>(cin) * C * 2u;
  const uint32_t w_alloc = (w_bytes + 1023u) & ~1023u;
  int Tmax = 512 / (4 * C);
  if (Tmax > 8) Tmax = 8;
  double best = -1.0;
  ConvStreamParams bp{};
  for (int nb = 1; nb <= H; ++nb) {  // balanced h-blocks only: every plane step of the layer costs about the same
    if (g_stream_force_nb > 0 && nb != g_stream_force_nb) continue;  // autotuning hook
    const int TH = (H + nb - 1) / nb;
    if ((H + TH - 1) / TH != nb) continue;
    const int T = (TH * WP + 127) / 128;
    if (T > Tmax) continue;
    const int HP = TH + 2;
    const int n_hblk = nb;
    const int rows_alloc = (HP * WP + 1 > T * 128 + 2 * WP + 3) ? HP * WP + 1 : T * 128 + 2 * WP + 3;
    const uint32_t a_chunk = (static_cast<uint32_t>(rows_alloc) * row_bytes + 1023u) & ~1023u;
    const uint32_t a_stage = a_chunk * chunks;
    int stages = static_cast<int>((budget - w_alloc) / a_stage);
    if (stages < 2) continue;
    stages = stages >= 4 ? 4 : 2;  // power of two (ring arithmetic by mask)
    // tiles actually issued per column plane (the partial last block skips empty tiles)
    int tiles = 0;
    for (int hb = 0; hb < n_hblk; ++hb) {
      const int thv = (H - hb * TH) < TH ? (H - hb * TH) : TH;
      tiles += (thv * WP + 127) / 128;
    }
    double eff = static_cast<double>(H) * W / (tiles * 128.0);
    if (stages < 4) eff *= 0.9;
    eff -= 0.002 * HP / TH;  // tie-break: less h-halo re-read
    if (eff > best) {
      best = eff;
      bp = ConvStreamParams{};
      bp.B = B; bp.D = D; bp.H = H; bp.W = W;
      bp.TH = TH; bp.HP = HP; bp.WP = WP; bp.n_hblk = n_hblk;
      bp.T = T; bp.C = C;
      const int R = 512 / (T * C) >= 8 ? 8 : 4;  // power of two, >= 4 by the choice of Tmax
      bp.R = R;
      bp.r_log2 = R == 8 ? 3 : 2;
      bp.st_log2 = stages == 4 ? 2 : 1;
      bp.kc = kc; bp.chunks = chunks; bp.row_bytes = row_bytes;
      bp.stages = stages;
      bp.issuers = T < stream_max_issuers() ? T : stream_max_issuers();
      bp.a_chunk_bytes = a_chunk; bp.a_stage_bytes = a_stage;
      bp.a_tx_bytes = static_cast<uint32_t>(HP * WP) * row_bytes * chunks;
      bp.w_bytes = w_bytes; bp.blk_bytes = static_cast<uint32_t>(C) * row_bytes;
    }
  }
  if (best < 0.0) return false;
  bp.tiles_n = nout / C;
  bp.total_steps = B * bp.n_hblk * D;
  int grid = sms / bp.tiles_n;
  if (grid > bp.total_steps / 4) grid = bp.total_steps / 4;  // at least ~4 planes per CTA (each cut costs 2 extra plane passes)
  if (grid < 1) grid = 1;
  bp.steps_per_cta = (bp.total_steps + grid - 1) / grid;
  bp.sbo = 8u * bp.row_bytes;
  bp.layout = umma_layout_for_swizzle(bp.row_bytes);
  for (int n = 1; n <= 3; ++n) bp.idesc[n - 1] = umma_idesc_bf16(n * C, 0, 0);
  uint32_t cols = 32;
  while (cols < static_cast<uint32_t>(bp.R * bp.T * C)) cols <<= 1;
  bp.tmem_cols = cols;
  *out = bp;
  return true;
}

int conv_stream_grid(const ConvStreamParams& p) { return (p.total_steps + p.steps_per_cta - 1) / p.steps_per_cta; }

static long long* g_stream_dbg = nullptr;
static int g_stream_dbg_steps = 0;
void conv_stream_set_debug(long long* buf, int steps) {
  g_stream_dbg = buf;
  g_stream_dbg_steps = steps;
}

int launch_conv_stream(const void* x, int ldx, const void* wpack, const float* bias, void* y, int ldy, int y_dtype,
                       int n_store, int cin, int nout, int act, float alpha, double* stats, ConvStreamParams p,
                       cudaStream_t st, float oscale, const float* post_scale, const float* post_shift) {
  CUtensorMap tmA, tmB;
  {
    uint64_t dims[5] = {static_cast<uint64_t>(cin), static_cast<uint64_t>(p.W), static_cast<uint64_t>(p.H),
                        static_cast<uint64_t>(p.D), static_cast<uint64_t>(p.B)};
    uint64_t strides[4] = {static_cast<uint64_t>(ldx) * 2, static_cast<uint64_t>(p.W) * ldx * 2,
                           static_cast<uint64_t>(p.H) * p.W * ldx * 2, static_cast<uint64_t>(p.D) * p.H * p.W * ldx * 2};
    uint32_t box[5] = {static_cast<uint32_t>(p.kc), static_cast<uint32_t>(p.WP), static_cast<uint32_t>(p.HP), 1, 1};
    int rc = encode_tiled_bf16(&tmA, x, 5, dims, strides, box, p.row_bytes);
    if (rc) return rc;
  }
  {
    uint64_t dims[3] = {static_cast<uint64_t>(cin), static_cast<uint64_t>(nout), 27};
    uint64_t strides[2] = {static_cast<uint64_t>(cin) * 2, static_cast<uint64_t>(nout) * cin * 2};
    uint32_t box[3] = {static_cast<uint32_t>(p.kc), static_cast<uint32_t>(p.C), 1};
    int rc = encode_tiled_bf16(&tmB, wpack, 3, dims, strides, box, p.row_bytes);
    if (rc) return rc;
  }
  p.y = y; p.ldy = ldy; p.y_dtype = y_dtype; p.n_store = n_store; p.bias = bias; p.act = act; p.alpha = alpha;
  p.stats = stats;
  p.dbg = g_stream_dbg;
  p.dbg_steps = g_stream_dbg_steps;
  static const int exp_flags = [] { const char* e = getenv("ICSG3D_STREAM_EXP"); return e ? atoi(e) : 0; }();
  p.exp_flags = exp_flags;
  p.oscale = oscale;
  p.post_scale = post_scale;
  p.post_shift = post_shift;
  static bool configured = false;
  if (!configured) {
    ICSG_CUDA(cudaFuncSetAttribute(conv3d_k3_stream_kernel<1, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 224 * 1024));
    ICSG_CUDA(cudaFuncSetAttribute(conv3d_k3_stream_kernel<2, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 224 * 1024));
    ICSG_CUDA(cudaFuncSetAttribute(conv3d_k3_stream_kernel<4, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 224 * 1024));
    ICSG_CUDA(cudaFuncSetAttribute(conv3d_k3_stream_kernel<1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 224 * 1024));
    ICSG_CUDA(cudaFuncSetAttribute(conv3d_k3_stream_kernel<2, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 224 * 1024));
    ICSG_CUDA(cudaFuncSetAttribute(conv3d_k3_stream_kernel<4, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 224 * 1024));
    configured = true;
  }
  const size_t smem = ((p.w_bytes + 1023u) & ~1023u) + static_cast<size_t>(p.stages) * p.a_stage_bytes + 1024;
  const dim3 grid(conv_stream_grid(p), p.tiles_n);
  // the inference form (BatchNorm affine after the activation) is its own instantiation: the training kernel's epilogue
  // is latency bound and carries nothing it does not need
  if (post_scale != nullptr) {
    if (p.kc == 16) launch_k(conv3d_k3_stream_kernel<1, true>, grid, kStreamThreads, smem, st, tmA, tmB, p);
    else if (p.kc == 32) launch_k(conv3d_k3_stream_kernel<2, true>, grid, kStreamThreads, smem, st, tmA, tmB, p);
    else launch_k(conv3d_k3_stream_kernel<4, true>, grid, kStreamThreads, smem, st, tmA, tmB, p);
  } else {
    if (p.kc == 16) launch_k(conv3d_k3_stream_kernel<1, false>, grid, kStreamThreads, smem, st, tmA, tmB, p);
    else if (p.kc == 32) launch_k(conv3d_k3_stream_kernel<2, false>, grid, kStreamThreads, smem, st, tmA, tmB, p);
    else launch_k(conv3d_k3_stream_kernel<4, false>, grid, kStreamThreads, smem, st, tmA, tmB, p);
  }
  ICSG_CHECK_LAUNCH();
  return ICSG3D_OK;
}

}  // namespace icsg3d
