// Conv3D 3x3x3 filter gradient (Conv3DBackpropFilterV2) — plane-streaming kernel with the kh taps folded into the
// MMA M dimension and the kw taps folded into N (tcgen05 / TMEM / TMA, sm_100a).  For the narrow 16^3 .. 64^3 layers
// (Cout <= 32), where the per-tap im2col re-fetch of conv3d_wgrad.cu (27x the activation through L2 -> SM) is the
// bound.  Reference op: the filter gradient TF computes for every Keras Conv3D in vae/lattice_vae.py:173-224.
//
//   dW[kd][kh][kw][ci][co] = sum_{n,d,h,w} x[n, d+kd-1, h+kh-1, w+kw-1, ci] * dy[n, d, h, w, co]
//
//   * a CTA walks a column (sample n, h-block hb) along d; per plane ONE TMA load of the halo'd x slab
//     ((TH+2) x (W+1) voxels x ca channels, zero fill = "same" padding) and ONE of the dy slab (TH x (W+1), the extra
//     column is out of bounds = zero, so the flat row pitch WP = W+1 is common to both);
//   * GEMM per (x plane i, dy plane o), |i-o| <= 1, kd = i-o+1:  D_kd[(kh, ci), (kw, co)] += X^T * DY  with K = the
//     flat voxel rows.  Both operands are MN-major straight from TMA.  The 128 MMA rows are channel blocks whose
//     leading-dimension stride is WP ROWS (block j reads x shifted by kh = j rows of the slab), the N = 3*Cout
//     columns are channel blocks whose stride is ONE row (block l reads dy shifted by l, i.e. kw = 2-l): overlapping
//     MN blocks are legal for swizzled MN-major descriptors (measured: profiles/r01_mnfold_probe.json), so all 9
//     (kh, kw) taps of a plane pair come out of ONE MMA per 16 voxels instead of 7-27;
//   * three accumulators (kd = 0, 1, 2) live in TMEM for the whole CTA, one issuer warp each;
//   * fp32 partials per CTA -> workspace, fixed-order reduction by wgrad_reduce_kernel (deterministic).
#include "common.cuh"

namespace icsg3d {

struct WgradStreamParams {
  int B, D, H, W;
  int TH, HP, WP, n_hblk;
  int cin, cout, ca, cb, chunks;
  int ksteps;              // MMAs (16 voxel rows each) per plane pair
  int stages;
  int total_steps, steps_per_cta, splits;
  uint32_t x_bytes, stage_bytes, x_tx, dy_tx;
  uint32_t lbo_a, sbo_a, lbo_b, sbo_b, layout_a, layout_b, idesc, tmem_cols;
  float* ws;               // [splits][27*cin*cout]
};

static constexpr int kWgsThreads = (1 + 3 + 4) * 32;
static constexpr int kWgsMaxStages = 6;
static constexpr uint32_t kWgsDyPad = 1024;  // zeroed bytes in front of every dy slab (the kw fold reads 2 rows before it)

__global__ void __launch_bounds__(kWgsThreads, 1)
conv3d_k3_wgrad_stream_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmDY,
                              const WgradStreamParams p) {
  pdl_prologue();
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t full_bar[kWgsMaxStages], empty_bar[kWgsMaxStages];
  __shared__ __align__(8) uint64_t done_bar;
  __shared__ uint32_t tmem_base_slot;

  const int warp = __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* sm = smem_raw + (base - smem_u32(smem_raw));
  const int chunk = blockIdx.y;

  // Every row the MMAs read outside the TMA boxes (guard rows, the junk kh blocks, the K tail) must be finite AND, on
  // the dy side, zero: clear the whole ring once; TMA only ever rewrites the slabs.
  {
    uint4* z = reinterpret_cast<uint4*>(sm);
    const int n16 = static_cast<int>((static_cast<size_t>(p.stages) * p.stage_bytes) >> 4);
    for (int i = threadIdx.x; i < n16; i += blockDim.x) z[i] = make_uint4(0u, 0u, 0u, 0u);
    fence_proxy_async();
  }
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmX);
    tma_prefetch_desc(&tmDY);
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 3);
    }
    mbar_init(&done_bar, 3);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc(&tmem_base_slot, p.tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_slot;
  const int ncols = 3 * p.cb;

  const int s_begin = blockIdx.x * p.steps_per_cta;
  const int s_end = min(p.total_steps, s_begin + p.steps_per_cta);

  if (warp == 0) {
    // ===================== TMA producer =====================
    const bool leader = elect_one();
    int astep = 0;
    for (int s = s_begin; s < s_end;) {
      const int col = s / p.D, db = s - col * p.D;
      const int de = min(p.D, db + (s_end - s));
      const int n = col / p.n_hblk, hb = col - n * p.n_hblk;
      for (int pl = max(0, db - 1); pl < de; ++pl, ++astep) {
        const int stage = astep % p.stages;
        mbar_wait(&empty_bar[stage], (static_cast<uint32_t>(astep / p.stages) & 1u) ^ 1u);
        if (leader) {
          uint8_t* st = sm + static_cast<size_t>(stage) * p.stage_bytes;
          mbar_expect_tx(&full_bar[stage], p.x_tx + p.dy_tx);
          tma_load_5d(st, &tmX, &full_bar[stage], chunk * p.ca, -1, hb * p.TH - 1, pl, n);
          tma_load_5d(st + p.x_bytes + kWgsDyPad, &tmDY, &full_bar[stage], 0, 0, hb * p.TH, pl, n);
        }
      }
      s += de - db;
    }
  } else if (warp <= 3) {
    // ===================== MMA issuers: warp 1: (x_i, dy_i) kd=1 | warp 2: (x_i, dy_i-1) kd=2 | warp 3: (x_i-1, dy_i) kd=0
    const int role = warp - 1;
    const int kd = role == 0 ? 1 : (role == 1 ? 2 : 0);
    const bool leader = elect_one();
    const uint32_t a_hi = umma_desc_hi(p.sbo_a, p.layout_a), b_hi = umma_desc_hi(p.sbo_b, p.layout_b);
    const uint32_t stage_lo = p.stage_bytes >> 4;
    const uint32_t a_ring_lo = umma_desc_lo(base, p.lbo_a);
    const uint32_t b_ring_lo = umma_desc_lo(base + p.x_bytes + kWgsDyPad - 2u * p.cb * 2u, p.lbo_b);  // 2 rows before the slab
    const uint32_t ka = (16u * p.ca * 2u) >> 4, kb = (16u * p.cb * 2u) >> 4;
    const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(kd * ncols);
    uint32_t started = 0;
    int astep = 0;
    for (int s = s_begin; s < s_end;) {
      const int col = s / p.D, db = s - col * p.D;
      const int de = min(p.D, db + (s_end - s));
      const int p_lo = max(0, db - 1);
      for (int pl = p_lo; pl < de; ++pl, ++astep) {
        const int stage = astep % p.stages;
        const int prev = (astep + p.stages - 1) % p.stages;
        mbar_wait(&full_bar[stage], static_cast<uint32_t>(astep / p.stages) & 1u);
        tc_fence_after();
        if (pl >= db && (role == 0 || pl > p_lo)) {
          const int xs = role == 2 ? prev : stage;
          const int ds = role == 1 ? prev : stage;
          uint32_t a_lo = a_ring_lo + static_cast<uint32_t>(xs) * stage_lo;
          uint32_t b_lo = b_ring_lo + static_cast<uint32_t>(ds) * stage_lo;
          if (leader) {
            umma_bf16_lohi(d_tmem, a_lo, a_hi, b_lo, b_hi, p.idesc, started);
            for (int ks = 1; ks < p.ksteps; ++ks) {
              a_lo += ka;
              b_lo += kb;
              umma_bf16_lohi(d_tmem, a_lo, a_hi, b_lo, b_hi, p.idesc, 1u);
            }
          }
          started = 1u;
        }
        if (leader) {
          if (pl > p_lo) umma_commit(&empty_bar[prev]);      // plane pl-1 is not needed any more
          if (pl == de - 1) umma_commit(&empty_bar[stage]);  // end of the segment: nor is this one
        }
        __syncwarp();
      }
      s += de - db;
    }
    if (leader) umma_commit(&done_bar);
  } else {
    // ===================== epilogue: fp32 partials of this CTA -> workspace =====================
    const int quarter = warp & 3;
    const int row = quarter * 32 + lane;
    const int j = row / p.ca;             // kh block
    const int ci = chunk * p.ca + (row - j * p.ca);
    // kd = 0 and 2 pair a plane with its predecessor: a CTA whose range never has one leaves those accumulators unwritten
    bool has_prev = false;
    for (int s = s_begin; s < s_end;) {
      const int col = s / p.D, db = s - col * p.D;
      const int de = min(p.D, db + (s_end - s));
      if (de - 1 >= 1) has_prev = true;
      s += de - db;
    }
    mbar_wait(&done_bar, 0);
    tc_fence_after();
    float* ws = p.ws + static_cast<size_t>(blockIdx.x) * 27 * p.cin * p.cout;
    for (int kd = 0; kd < 3; ++kd) {
      const bool written = kd == 1 || has_prev;
      for (int c0 = 0; c0 < ncols; c0 += 16) {
        uint32_t v[16];
        if (written) {
          tmem_ld16(tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + static_cast<uint32_t>(kd * ncols + c0), v);
          tmem_ld_wait();
        } else {
#pragma unroll
          for (int i = 0; i < 16; ++i) v[i] = 0u;
        }
        if (j >= 3) continue;
        const int l = c0 / p.cb;
        const int co = c0 - l * p.cb;
        const int tap = (kd * 3 + j) * 3 + (2 - l);
        float4* dst = reinterpret_cast<float4*>(ws + (static_cast<size_t>(tap) * p.cin + ci) * p.cout + co);
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

static bool wgs_pow2(int v) { return v > 0 && (v & (v - 1)) == 0; }

bool wgrad_stream_plan(int B, int D, int H, int W, int cin, int cout, int sms, WgradStreamParams* out) {
  if (cout != 16 && cout != 32) return false;
  if (W < 16 || W > 64 || !wgs_pow2(W) || !wgs_pow2(H) || !wgs_pow2(D) || D < 2) return false;
  if (cin % 16 != 0 || cin > 256) return false;
  WgradStreamParams p{};
  p.B = B; p.D = D; p.H = H; p.W = W; p.cin = cin; p.cout = cout;
  p.ca = (cin % 32 == 0) ? 32 : 16;
  p.cb = cout;
  p.chunks = cin / p.ca;
  p.WP = W + 1;
  const int nblk_a = 128 / p.ca;
  const uint32_t rba = p.ca * 2u, rbb = p.cb * 2u;
  const uint32_t budget = 220u * 1024u;
  bool found = false;
  for (int nb = 1; nb <= H && !found; ++nb) {  // largest balanced h-block that leaves >= 3 ring stages
    const int TH = (H + nb - 1) / nb;
    if ((H + TH - 1) / TH != nb) continue;
    const int HP = TH + 2;
    const int kpad = ((TH * p.WP + 2 + 15) / 16) * 16;
    const int x_rows = (HP * p.WP > kpad + (nblk_a - 1) * p.WP ? HP * p.WP : kpad + (nblk_a - 1) * p.WP) + 1;
    const uint32_t x_bytes = (static_cast<uint32_t>(x_rows) * rba + 1023u) & ~1023u;
    const uint32_t dy_bytes = kWgsDyPad + ((static_cast<uint32_t>(kpad + 2) * rbb + 1023u) & ~1023u);
    const uint32_t stage = x_bytes + dy_bytes;
    int stages = static_cast<int>(budget / stage);
    if (stages > kWgsMaxStages) stages = kWgsMaxStages;
    if (stages < 3) continue;
    found = true;
    p.TH = TH; p.HP = HP; p.n_hblk = nb;
    p.ksteps = kpad / 16;
    p.stages = stages;
    p.x_bytes = x_bytes; p.stage_bytes = stage;
    p.x_tx = static_cast<uint32_t>(HP * p.WP) * rba;
    p.dy_tx = static_cast<uint32_t>(TH * p.WP) * rbb;
  }
  if (!found) return false;
  p.total_steps = B * p.n_hblk * D;
  int splits = sms / p.chunks;
  if (splits > p.total_steps / 2) splits = p.total_steps / 2;
  if (splits < 1) splits = 1;
  p.steps_per_cta = (p.total_steps + splits - 1) / splits;
  p.splits = (p.total_steps + p.steps_per_cta - 1) / p.steps_per_cta;
  p.lbo_a = static_cast<uint32_t>(p.WP) * rba;
  p.sbo_a = 8u * rba;
  p.lbo_b = rbb;
  p.sbo_b = 8u * rbb;
  p.layout_a = umma_layout_for_swizzle(static_cast<int>(rba));
  p.layout_b = umma_layout_for_swizzle(static_cast<int>(rbb));
  p.idesc = umma_idesc_bf16(3 * p.cb, 1, 1);
  uint32_t cols = 32;
  while (cols < static_cast<uint32_t>(9 * p.cb)) cols <<= 1;
  p.tmem_cols = cols;
  *out = p;
  return true;
}

int64_t wgrad_stream_workspace(const WgradStreamParams& p) {
  return static_cast<int64_t>(p.splits) * 27 * p.cin * p.cout * 4;
}

int launch_wgrad_stream(const void* x, int ldx, const void* dy, int ldy, WgradStreamParams p, float* ws, cudaStream_t st) {
  CUtensorMap tmX, tmDY;
  {
    uint64_t dims[5] = {static_cast<uint64_t>(p.cin), static_cast<uint64_t>(p.W), static_cast<uint64_t>(p.H),
                        static_cast<uint64_t>(p.D), static_cast<uint64_t>(p.B)};
    uint64_t strides[4] = {static_cast<uint64_t>(ldx) * 2, static_cast<uint64_t>(p.W) * ldx * 2,
                           static_cast<uint64_t>(p.H) * p.W * ldx * 2, static_cast<uint64_t>(p.D) * p.H * p.W * ldx * 2};
    uint32_t box[5] = {static_cast<uint32_t>(p.ca), static_cast<uint32_t>(p.WP), static_cast<uint32_t>(p.HP), 1, 1};
    int rc = encode_tiled_bf16(&tmX, x, 5, dims, strides, box, p.ca * 2);
    if (rc) return rc;
  }
  {
    uint64_t dims[5] = {static_cast<uint64_t>(p.cout), static_cast<uint64_t>(p.W), static_cast<uint64_t>(p.H),
                        static_cast<uint64_t>(p.D), static_cast<uint64_t>(p.B)};
    uint64_t strides[4] = {static_cast<uint64_t>(ldy) * 2, static_cast<uint64_t>(p.W) * ldy * 2,
                           static_cast<uint64_t>(p.H) * p.W * ldy * 2, static_cast<uint64_t>(p.D) * p.H * p.W * ldy * 2};
    uint32_t box[5] = {static_cast<uint32_t>(p.cb), static_cast<uint32_t>(p.WP), static_cast<uint32_t>(p.TH), 1, 1};
    int rc = encode_tiled_bf16(&tmDY, dy, 5, dims, strides, box, p.cb * 2);
    if (rc) return rc;
  }
  p.ws = ws;
  static bool configured = false;
  if (!configured) {
    ICSG_CUDA(cudaFuncSetAttribute(conv3d_k3_wgrad_stream_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 224 * 1024));
    configured = true;
  }
  const size_t smem = static_cast<size_t>(p.stages) * p.stage_bytes + 1024;
  dim3 grid(p.splits, p.chunks, 1);
  launch_k(conv3d_k3_wgrad_stream_kernel, grid, kWgsThreads, smem, st, tmX, tmDY, p);
  ICSG_CHECK_LAUNCH();
  return ICSG3D_OK;
}

int64_t wgrad_stream_workspace_bytes(int B, int D, int H, int W, int cin, int cout, int sms) {
  WgradStreamParams p;
  if (!wgrad_stream_plan(B, D, H, W, cin, cout, sms, &p)) return -1;
  return wgrad_stream_workspace(p);
}

// returns ICSG3D_OK, 1 when the shape is not eligible (caller falls back), or an error code
int wgrad_stream_run(const void* x, int ldx, const void* dy, int ldy, int B, int D, int H, int W, int cin, int cout, int sms,
                     float* ws, int64_t ws_bytes, int* splits_out, cudaStream_t st) {
  WgradStreamParams p;
  if (!wgrad_stream_plan(B, D, H, W, cin, cout, sms, &p)) return 1;
  if (ws_bytes < wgrad_stream_workspace(p)) return 1;
  *splits_out = p.splits;
  return launch_wgrad_stream(x, ldx, dy, ldy, p, ws, st);
}

}  // namespace icsg3d
