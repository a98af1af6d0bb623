// Hardware probe (tests/tools only): does a K-major swizzled UMMA smem descriptor accept a start address
// that is shifted by an arbitrary number of ROWS (not a multiple of the 8-row swizzle atom)?
// This decides whether a halo'd activation block can be loaded into shared memory ONCE and re-used for
// all 27 taps of a 3x3x3 convolution by moving the descriptor start address (DESIGN.md "halo reuse").
//
// A: [rows >= 128+shift_max, kc] bf16 row-major, B: [n, kc] bf16 row-major.
// out[mode][s][128][n] = A[s : s+128, :] * B^T   for s in [0, nshift), mode 0: base_offset=0,
// mode 1: base_offset = (start_address >> 7) & 7.
#include "common.cuh"

namespace icsg3d {

__global__ void __launch_bounds__(128, 1)
probe_shift_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, float* out, int rows,
                   int kc, int n, int nshift) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t load_bar;
  __shared__ __align__(8) uint64_t mma_bar;
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* sA = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t a_bytes = static_cast<uint32_t>(rows) * kc * 2u;
  const uint32_t a_bytes_al = (a_bytes + 1023u) & ~1023u;
  uint8_t* sB = sA + a_bytes_al;
  if (threadIdx.x == 0) {
    mbar_init(&load_bar, 1);
    mbar_init(&mma_bar, 1);
    fence_mbar_init();
  }
  if (warp == 0) tmem_alloc(&tmem_slot, 256);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  if (threadIdx.x == 0) {
    mbar_expect_tx(&load_bar, a_bytes + static_cast<uint32_t>(n) * kc * 2u);
    tma_load_2d(sA, &tmA, &load_bar, 0, 0);
    tma_load_2d(sB, &tmB, &load_bar, 0, 0);
  }
  mbar_wait(&load_bar, 0);
  tc_fence_after();
  const uint32_t layout = umma_layout_for_swizzle(kc * 2);
  const uint32_t sbo = 8u * kc * 2u;
  const uint32_t idesc = umma_idesc_bf16(n, 0, 0);
  uint32_t phase = 0;
  for (int mode = 0; mode < 2; ++mode) {
    for (int s = 0; s < nshift; ++s) {
      if (threadIdx.x == 0) {
        for (int k = 0; k < kc / 16; ++k) {
          const uint32_t a_addr = base + static_cast<uint32_t>(s) * kc * 2u + k * 32u;
          uint64_t adesc = umma_smem_desc(a_addr, 16u, sbo, layout);
          if (mode == 1) adesc |= static_cast<uint64_t>((a_addr >> 7) & 7u) << 49;
          const uint64_t bdesc = umma_smem_desc(base + a_bytes_al + k * 32u, 16u, sbo, layout);
          umma_bf16(tmem, adesc, bdesc, idesc, k != 0);
        }
        umma_commit(&mma_bar);
      }
      mbar_wait(&mma_bar, phase);
      phase ^= 1u;
      tc_fence_after();
      const int row = warp * 32 + lane;
      float* dst = out + ((static_cast<size_t>(mode) * nshift + s) * 128 + row) * n;
      for (int c0 = 0; c0 < n; c0 += 16) {
        uint32_t v[16];
        tmem_ld16(tmem + (static_cast<uint32_t>(warp * 32) << 16) + c0, v);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 16; ++i) dst[c0 + i] = __uint_as_float(v[i]);
      }
      tc_fence_before();
      __syncthreads();
      tc_fence_after();
    }
  }
  if (warp == 0) tmem_dealloc(tmem, 256);
}

}  // namespace icsg3d

using namespace icsg3d;

extern "C" int icsg3d_probe_shifted_desc(const void* a, const void* b, float* out, int rows, int kc, int n, int nshift,
                                         void* stream) {
  ICSG_REQUIRE(a && b && out, "probe: null pointer");
  ICSG_REQUIRE((kc == 16 || kc == 32 || kc == 64) && n % 16 == 0 && n >= 16 && n <= 256 && rows >= 128 + nshift &&
                   rows <= 256,
               "probe: bad shape");
  CUtensorMap tmA, tmB;
  {
    uint64_t dims[2] = {static_cast<uint64_t>(kc), static_cast<uint64_t>(rows)};
    uint64_t strides[1] = {static_cast<uint64_t>(kc) * 2};
    uint32_t box[2] = {static_cast<uint32_t>(kc), static_cast<uint32_t>(rows)};
    int rc = encode_tiled_bf16(&tmA, a, 2, dims, strides, box, kc * 2);
    if (rc) return rc;
  }
  {
    uint64_t dims[2] = {static_cast<uint64_t>(kc), static_cast<uint64_t>(n)};
    uint64_t strides[1] = {static_cast<uint64_t>(kc) * 2};
    uint32_t box[2] = {static_cast<uint32_t>(kc), static_cast<uint32_t>(n)};
    int rc = encode_tiled_bf16(&tmB, b, 2, dims, strides, box, kc * 2);
    if (rc) return rc;
  }
  const size_t smem = static_cast<size_t>(rows) * kc * 2 + static_cast<size_t>(n) * kc * 2 + 3072;
  ICSG_CUDA(cudaFuncSetAttribute(probe_shift_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
  probe_shift_kernel<<<1, 128, smem, static_cast<cudaStream_t>(stream)>>>(tmA, tmB, out, rows, kc, n, nshift);
  ICSG_CHECK_LAUNCH();
  return ICSG3D_OK;
}

// ------------------------------------------------------------------------------------------------------
// Probe 2: tensor-pipe occupancy of one tcgen05.mma (SS mode, bf16, K=16) as a function of (M, N):
// one CTA issues `reps` MMAs back to back (alternating `nacc` accumulators), commits, waits, and reports
// elapsed SM cycles.  Operand contents are irrelevant (zeros).
// ------------------------------------------------------------------------------------------------------
namespace icsg3d {
__global__ void __launch_bounds__(128, 1) probe_mma_rate_kernel(long long* out, int m, int n, int reps, int nacc, int swz,
                                                                int a_step, int b_step) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5;
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  if (threadIdx.x == 0) {
    mbar_init(&bar, 1);
    fence_mbar_init();
  }
  for (int i = threadIdx.x; i < 96 * 1024 / 4; i += blockDim.x)
    reinterpret_cast<uint32_t*>(smem_raw + (base - smem_u32(smem_raw)))[i] = 0u;
  fence_proxy_async();
  if (warp == 0) tmem_alloc(&tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  if (threadIdx.x == 0) {
    const uint32_t row_bytes = static_cast<uint32_t>(swz);
    const uint32_t layout = umma_layout_for_swizzle(swz);
    const uint32_t sbo = 8u * row_bytes;
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (static_cast<uint32_t>(n >> 3) << 17) |
                           (static_cast<uint32_t>(m >> 4) << 24);
    const uint64_t adesc = umma_smem_desc(base, 16u, sbo, layout);
    const uint64_t bdesc = umma_smem_desc(base + 48 * 1024, 16u, sbo, layout);
    const long long t0 = clock64();
    const uint32_t mask = static_cast<uint32_t>(nacc - 1);  // nacc is a power of two
    // a_step / b_step (bytes, multiples of 16): operand start address advances by that much per MMA, cycling
    // through 8 positions, so consecutive MMAs read DIFFERENT shared-memory operands (no operand reuse).
    const uint64_t astep = static_cast<uint64_t>(a_step >> 4), bstep = static_cast<uint64_t>(b_step >> 4);
#pragma unroll 8
    for (int r = 0; r < reps; ++r) {
      const uint64_t j = static_cast<uint64_t>(r & 7);
      umma_bf16(tmem + (static_cast<uint32_t>(r) & mask) * static_cast<uint32_t>(n), adesc + j * astep, bdesc + j * bstep, idesc, 1u);
    }
    umma_commit(&bar);
    const long long t1 = clock64();
    mbar_wait(&bar, 0);
    const long long t2 = clock64();
    out[0] = t1 - t0;
    out[1] = t2 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem, 512);
  }
}
}  // namespace icsg3d

extern "C" int icsg3d_probe_mma_rate(int64_t* out, int m, int n, int reps, int nacc, int swizzle_bytes, int a_step,
                                     int b_step, void* stream) {
  ICSG_REQUIRE(out && (m == 64 || m == 128) && n % 16 == 0 && n >= 16 && n <= 256 && nacc >= 1 && nacc * n <= 512,
               "probe_mma_rate: bad arguments");
  ICSG_CUDA(cudaFuncSetAttribute(icsg3d::probe_mma_rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
  icsg3d::probe_mma_rate_kernel<<<1, 128, 98 * 1024, static_cast<cudaStream_t>(stream)>>>(reinterpret_cast<long long*>(out), m, n,
                                                                                         reps, nacc, swizzle_bytes, a_step, b_step);
  ICSG_CHECK_LAUNCH();
  return ICSG3D_OK;
}

// ------------------------------------------------------------------------------------------------------
// Probe 2b: the same for MN-MAJOR operands as the filter-gradient kernels use them (A = [voxel][64 channels] boxes, two
// boxes side by side along M; B = [voxel][64 channels] boxes along N; K = voxels, 16 per MMA = two 8-row groups):
// cycles per MMA when consecutive MMAs walk the 8 k-steps of a 128-voxel tile.  mn = 0 runs the K-major form of probe 2
// with the same shared-memory footprint for comparison.
// ------------------------------------------------------------------------------------------------------
namespace icsg3d {
__global__ void __launch_bounds__(128, 1) probe_mma_rate_mn_kernel(long long* out, int n, int reps, int nacc, int mn) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5;
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  if (threadIdx.x == 0) {
    mbar_init(&bar, 1);
    fence_mbar_init();
  }
  for (int i = threadIdx.x; i < 96 * 1024 / 4; i += blockDim.x)
    reinterpret_cast<uint32_t*>(smem_raw + (base - smem_u32(smem_raw)))[i] = 0u;
  fence_proxy_async();
  if (warp == 0) tmem_alloc(&tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  if (threadIdx.x == 0) {
    const uint32_t box = 128u * 128u;  // [128 voxels][64 channels] bf16
    const uint32_t idesc = umma_idesc_bf16(n, mn, mn);
    const uint32_t hi = umma_desc_hi(1024u, umma_layout_for_swizzle(128));
    const uint32_t a_lo = umma_desc_lo(base, mn ? box : 16u), b_lo = umma_desc_lo(base + 32u * 1024u, mn ? box : 16u);
    const uint32_t kstep = mn ? (2u * 1024u) >> 4 : 2u;  // 16 voxels = two 8-row groups | 32 bytes along the 128-byte row
    const uint32_t mask = static_cast<uint32_t>(nacc - 1);
    const long long t0 = clock64();
#pragma unroll 8
    for (int r = 0; r < reps; ++r) {
      const uint32_t k = static_cast<uint32_t>(r) & (mn ? 7u : 3u);
      umma_bf16_lohi(tmem + (static_cast<uint32_t>(r >> 3) & mask) * static_cast<uint32_t>(n), a_lo + k * kstep, hi, b_lo + k * kstep, hi,
                     idesc, 1u);
    }
    umma_commit(&bar);
    mbar_wait(&bar, 0);
    const long long t2 = clock64();
    out[0] = reps;
    out[1] = t2 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem, 512);
  }
}
}  // namespace icsg3d

extern "C" int icsg3d_probe_mma_rate_mn(int64_t* out, int n, int reps, int nacc, int mn, void* stream) {
  ICSG_REQUIRE(out && n % 16 == 0 && n >= 16 && n <= 256 && nacc >= 1 && nacc * n <= 512 && (mn == 0 || mn == 1),
               "probe_mma_rate_mn: bad arguments");
  ICSG_CUDA(cudaFuncSetAttribute(icsg3d::probe_mma_rate_mn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
  icsg3d::probe_mma_rate_mn_kernel<<<1, 128, 98 * 1024, static_cast<cudaStream_t>(stream)>>>(reinterpret_cast<long long*>(out), n, reps,
                                                                                            nacc, mn);
  ICSG_CHECK_LAUNCH();
  return ICSG3D_OK;
}

// ------------------------------------------------------------------------------------------------------
// Probe 3: the halo kernel's exact MMA issue pattern (tap-outer, G accumulators, KSTEPS k-steps, per-tap row
// shifts into a resident A block, resident B tiles) with NO TMA, NO barriers and NO epilogue: isolates the
// tensor-pipe cost of the pattern itself.  out[0] = cycles for `items` items, out[1] = MMAs issued.
// ------------------------------------------------------------------------------------------------------
namespace icsg3d {
__global__ void __launch_bounds__(192, 1) probe_halo_pattern_kernel(long long* out, int G, int nt, int plane_rows, int WP,
                                                                    int row_bytes, int ksteps, int items, int mode) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5;
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  if (threadIdx.x == 0) {
    mbar_init(&bar, 1);
    fence_mbar_init();
  }
  for (int i = threadIdx.x; i < 200 * 1024 / 4; i += blockDim.x)
    reinterpret_cast<uint32_t*>(smem_raw + (base - smem_u32(smem_raw)))[i] = 0u;
  fence_proxy_async();
  if (warp == 1) tmem_alloc(&tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  // variant A (mode < 4): whole warp runs the loop, elected lane issues (the production kernels' structure)
  // variant B (mode >= 4): ONE thread (threadIdx.x == 32) runs the whole loop
  const bool single = mode >= 4;
  mode &= 3;
  if (warp == 1 && (!single || threadIdx.x == 32)) {
    const bool leader = single ? true : elect_one();
    const uint32_t desc_hi = umma_desc_hi(8u * row_bytes, umma_layout_for_swizzle(row_bytes));
    const uint32_t idesc = umma_idesc_bf16(nt, 0, 0);
    const uint32_t a_base_lo = umma_desc_lo(base, 16u);
    const uint32_t b_base_lo = umma_desc_lo(base + 96 * 1024, 16u);
    const uint32_t b_unit_lo = (static_cast<uint32_t>(nt) * row_bytes) >> 4;
    const uint32_t tile_lo = (128u * row_bytes) >> 4;
    long long n_mma = 0;
    const long long t0 = clock64();
    for (int it = 0; it < items; ++it) {
      const uint32_t tmem_set = tmem + static_cast<uint32_t>((it & 1) * G * nt);
      bool first = true;
      int unit = 0;
      for (int kd = 0; kd < 3; ++kd)
        for (int kh = 0; kh < 3; ++kh)
          for (int kw = 0; kw < 3; ++kw, ++unit) {
            uint32_t shift_rows = static_cast<uint32_t>(kd * plane_rows + kh * WP + kw);
            if (mode == 1) shift_rows = 0;
            if (mode == 2) shift_rows &= ~7u;
            const uint32_t b_lo = b_base_lo + static_cast<uint32_t>(mode == 3 ? 0 : unit) * b_unit_lo;
            uint32_t a_lo = a_base_lo + ((shift_rows * static_cast<uint32_t>(row_bytes)) >> 4);
            uint32_t d_tmem = tmem_set;
            if (leader) {
              for (int g = 0; g < G; ++g) {
                for (int k = 0; k < ksteps; ++k)
                  umma_bf16_lohi(d_tmem, a_lo + 2u * k, desc_hi, b_lo + 2u * k, desc_hi, idesc, (!first || k != 0) ? 1u : 0u);
                a_lo += tile_lo;
                d_tmem += static_cast<uint32_t>(nt);
              }
            }
            n_mma += G * ksteps;
            first = false;
          }
    }
    if (leader) {
      umma_commit(&bar);
      mbar_wait(&bar, 0);
      out[0] = clock64() - t0;
      out[1] = n_mma;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem, 512);
  }
}
}  // namespace icsg3d

extern "C" int icsg3d_probe_halo_pattern(int64_t* out, int G, int nt, int plane_rows, int WP, int row_bytes, int ksteps,
                                         int items, int mode, void* stream) {
  ICSG_REQUIRE(out && 2 * G * nt <= 512 && 27 * nt * row_bytes <= 112 * 1024, "probe_halo_pattern: bad arguments");
  ICSG_CUDA(cudaFuncSetAttribute(icsg3d::probe_halo_pattern_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
  icsg3d::probe_halo_pattern_kernel<<<1, 192, 210 * 1024, static_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<long long*>(out), G, nt, plane_rows, WP, row_bytes, ksteps, items, mode);
  ICSG_CHECK_LAUNCH();
  return ICSG3D_OK;
}


// ------------------------------------------------------------------------------------------------
// Probe 4: MN-major swizzled operands whose MN blocks OVERLAP (leading byte offset = a few rows instead of a whole
// atom-aligned slab).  Decides whether the filter-gradient GEMM can fold taps into M (x shifted by kw rows per
// channel block) and into N (dy shifted by kh*WP rows per channel block) straight from one resident halo slab:
//   out[j*ca + ci][l*cb + co] = sum_{r < 16*ksteps} X[r + j*a_shift][ci] * Y[r + l*b_shift][co]
// X: [rows][ca], Y: [rows][cb] bf16 row-major, TMA-loaded with swizzle = row bytes (the production layout).
// ------------------------------------------------------------------------------------------------
namespace icsg3d {
__global__ void __launch_bounds__(128, 1)
probe_mn_fold_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmY, float* out, int rows,
                     int ca, int cb, int a_shift, int nblk_b, int b_shift, int ksteps) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t load_bar;
  __shared__ __align__(8) uint64_t mma_bar;
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* sX = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t x_bytes = static_cast<uint32_t>(rows) * ca * 2u;
  const uint32_t x_bytes_al = (x_bytes + 1023u) & ~1023u;
  const uint32_t y_bytes = static_cast<uint32_t>(rows) * cb * 2u;
  uint8_t* sY = sX + x_bytes_al;
  if (threadIdx.x == 0) {
    mbar_init(&load_bar, 1);
    mbar_init(&mma_bar, 1);
    fence_mbar_init();
  }
  if (warp == 0) tmem_alloc(&tmem_slot, 256);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  if (threadIdx.x == 0) {
    mbar_expect_tx(&load_bar, x_bytes + y_bytes);
    tma_load_2d(sX, &tmX, &load_bar, 0, 0);
    tma_load_2d(sY, &tmY, &load_bar, 0, 0);
  }
  mbar_wait(&load_bar, 0);
  tc_fence_after();
  const int n = nblk_b * cb;
  if (threadIdx.x == 0) {
    const uint32_t rba = ca * 2u, rbb = cb * 2u;
    const uint32_t idesc = umma_idesc_bf16(n, 1, 1);
    for (int s = 0; s < ksteps; ++s) {
      const uint64_t adesc = umma_smem_desc(base + static_cast<uint32_t>(s) * 16u * rba, static_cast<uint32_t>(a_shift) * rba,
                                            8u * rba, umma_layout_for_swizzle(rba));
      const uint64_t bdesc = umma_smem_desc(base + x_bytes_al + static_cast<uint32_t>(s) * 16u * rbb,
                                            static_cast<uint32_t>(b_shift) * rbb, 8u * rbb, umma_layout_for_swizzle(rbb));
      umma_bf16(tmem, adesc, bdesc, idesc, s != 0);
    }
    umma_commit(&mma_bar);
  }
  mbar_wait(&mma_bar, 0);
  tc_fence_after();
  const int row = warp * 32 + lane;
  for (int c0 = 0; c0 < n; c0 += 16) {
    uint32_t v[16];
    tmem_ld16(tmem + (static_cast<uint32_t>(warp * 32) << 16) + c0, v);
    tmem_ld_wait();
#pragma unroll
    for (int i = 0; i < 16; ++i) out[static_cast<size_t>(row) * n + c0 + i] = __uint_as_float(v[i]);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == 0) tmem_dealloc(tmem, 256);
}
}  // namespace icsg3d

extern "C" int icsg3d_probe_mn_fold(const void* x, const void* y, float* out, int rows, int ca, int cb, int a_shift,
                                    int nblk_b, int b_shift, int ksteps, void* stream) {
  ICSG_REQUIRE(x && y && out, "probe_mn_fold: null pointer");
  ICSG_REQUIRE((ca == 16 || ca == 32 || ca == 64) && (cb == 16 || cb == 32 || cb == 64) && nblk_b >= 1 && nblk_b * cb <= 256 &&
                   rows <= 256 && ksteps >= 1,
               "probe_mn_fold: bad shape");
  ICSG_REQUIRE(16 * ksteps + (128 / ca - 1) * a_shift <= rows && 16 * ksteps + (nblk_b - 1) * b_shift <= rows,
               "probe_mn_fold: not enough rows");
  CUtensorMap tmX, tmY;
  {
    uint64_t dims[2] = {static_cast<uint64_t>(ca), static_cast<uint64_t>(rows)};
    uint64_t strides[1] = {static_cast<uint64_t>(ca) * 2};
    uint32_t box[2] = {static_cast<uint32_t>(ca), static_cast<uint32_t>(rows)};
    int rc = encode_tiled_bf16(&tmX, x, 2, dims, strides, box, ca * 2);
    if (rc) return rc;
  }
  {
    uint64_t dims[2] = {static_cast<uint64_t>(cb), static_cast<uint64_t>(rows)};
    uint64_t strides[1] = {static_cast<uint64_t>(cb) * 2};
    uint32_t box[2] = {static_cast<uint32_t>(cb), static_cast<uint32_t>(rows)};
    int rc = encode_tiled_bf16(&tmY, y, 2, dims, strides, box, cb * 2);
    if (rc) return rc;
  }
  const size_t smem = static_cast<size_t>(rows) * (ca + cb) * 2 + 3072;
  ICSG_CUDA(cudaFuncSetAttribute(probe_mn_fold_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
  probe_mn_fold_kernel<<<1, 128, smem, static_cast<cudaStream_t>(stream)>>>(tmX, tmY, out, rows, ca, cb, a_shift, nblk_b,
                                                                           b_shift, ksteps);
  ICSG_CHECK_LAUNCH();
  return ICSG3D_OK;
}
