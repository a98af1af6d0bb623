// Host runtime bits of libicsg3d: error reporting, device queries, TMA tensor-map encoding.
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "common.cuh"

namespace icsg3d {

static thread_local char g_err[512] = "";
unsigned long long g_launches = 0;

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int cuda_fail(cudaError_t e, const char* what, const char* file, int line) {
  set_error("CUDA error %d (%s) at %s:%d in `%s`", static_cast<int>(e), cudaGetErrorString(e), file, line, what);
  return ICSG3D_ERR_CUDA;
}

void set_peer_timeout(double seconds);
void set_pdl(int on);
void set_peer_ll(int on);

int sm_count() {
  static int cached[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return -1;
  if (cached[dev] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return -1;
    cached[dev] = n;
  }
  return cached[dev];
}

static long long g_peer_timeout = -1;
long long peer_timeout_cycles() {
  if (g_peer_timeout < 0) {
    double sec = 600.0;
    if (const char* e = getenv("ICSG3D_PEER_TIMEOUT_S")) {
      const double v = atof(e);
      if (v > 0.0) sec = v;
    }
    g_peer_timeout = static_cast<long long>(sec * 2.0e9);  // clock64 ticks at <= 2 GHz
  }
  return g_peer_timeout;
}
void set_peer_timeout(double seconds) { g_peer_timeout = static_cast<long long>(seconds * 2.0e9); }

static int g_peer_ll = -1;
bool peer_ll_enabled() {
  if (g_peer_ll < 0) {
    const char* e = getenv("ICSG3D_PEER_LL");
    g_peer_ll = (e && e[0] == '0') ? 0 : 1;
  }
  return g_peer_ll == 1;
}
void set_peer_ll(int on) { g_peer_ll = on ? 1 : 0; }

static int g_pdl = -1;
bool pdl_enabled() {
  if (g_pdl < 0) {
    const char* e = getenv("ICSG3D_PDL");
    g_pdl = (e && e[0] == '1') ? 1 : 0;  // off by default: measured 3.173 vs 3.179 ms per captured step (no gain)
  }
  return g_pdl == 1;
}
void set_pdl(int on) { g_pdl = on ? 1 : 0; }

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_tiled() {
  static EncodeTiledFn fn = nullptr;
  if (fn) return fn;
  void* p = nullptr;
  cudaDriverEntryPointQueryResult qres;
  cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres);
  if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || p == nullptr) {
    set_error("cuTensorMapEncodeTiled not available from the driver (cudaGetDriverEntryPoint: %s)",
              cudaGetErrorString(e));
    return nullptr;
  }
  fn = reinterpret_cast<EncodeTiledFn>(p);
  return fn;
}

int encode_tiled_bf16(CUtensorMap* map, const void* base, int rank, const uint64_t* dims,
                      const uint64_t* strides_bytes, const uint32_t* box, int swizzle_bytes) {
  EncodeTiledFn fn = get_encode_tiled();
  if (!fn) return ICSG3D_ERR_CUDA;
  cuuint64_t gdim[5];
  cuuint64_t gstr[4];
  cuuint32_t bdim[5];
  cuuint32_t estr[5];
  for (int i = 0; i < rank; ++i) {
    gdim[i] = dims[i];
    bdim[i] = box[i];
    estr[i] = 1;
    if (i > 0) gstr[i - 1] = strides_bytes[i - 1];
  }
  CUtensorMapSwizzle sw = swizzle_bytes == 128  ? CU_TENSOR_MAP_SWIZZLE_128B
                          : swizzle_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B
                          : swizzle_bytes == 32 ? CU_TENSOR_MAP_SWIZZLE_32B
                                                : CU_TENSOR_MAP_SWIZZLE_NONE;
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, static_cast<cuuint32_t>(rank), const_cast<void*>(base),
                  gdim, gstr, bdim, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error(
        "cuTensorMapEncodeTiled failed (CUresult %d): rank %d dims [%llu %llu %llu %llu %llu] box [%u %u %u %u %u] "
        "swizzle %d base %p",
        static_cast<int>(r), rank, (unsigned long long)dims[0], (unsigned long long)(rank > 1 ? dims[1] : 0),
        (unsigned long long)(rank > 2 ? dims[2] : 0), (unsigned long long)(rank > 3 ? dims[3] : 0),
        (unsigned long long)(rank > 4 ? dims[4] : 0), box[0], rank > 1 ? box[1] : 0, rank > 2 ? box[2] : 0,
        rank > 3 ? box[3] : 0, rank > 4 ? box[4] : 0, swizzle_bytes, base);
    return ICSG3D_ERR_CUDA;
  }
  return ICSG3D_OK;
}

}  // namespace icsg3d

extern "C" {

const char* icsg3d_last_error(void) { return icsg3d::g_err; }

int icsg3d_version(void) { return 100; }

int icsg3d_sm_count(void) { return icsg3d::sm_count(); }

int64_t icsg3d_launch_count(void) { return static_cast<int64_t>(icsg3d::g_launches); }

int icsg3d_set_peer_ll(int on) {
  icsg3d::set_peer_ll(on);
  return ICSG3D_OK;
}

int icsg3d_set_pdl(int on) {
  icsg3d::set_pdl(on);
  return ICSG3D_OK;
}

int icsg3d_set_peer_timeout(double seconds) {
  if (!(seconds > 0.0)) {
    icsg3d::set_error("set_peer_timeout: seconds must be > 0");
    return ICSG3D_ERR_INVALID;
  }
  icsg3d::set_peer_timeout(seconds);
  return ICSG3D_OK;
}

}  // extern "C"
