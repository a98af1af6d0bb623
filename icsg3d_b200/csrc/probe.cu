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
