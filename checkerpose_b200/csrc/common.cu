#include <stdarg.h>
#include <string.h>

#include <cuda.h>

#include "common.cuh"

namespace cp {
static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int num_sms() {
  static int cache[64] = {0};     // benign race: every writer stores the same value
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  if (cache[dev] == 0) {
    int n = 0;
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    cache[dev] = n > 0 ? n : 148;
  }
  return cache[dev];
}

// cuTensorMapEncodeTiled through the runtime's driver entry point (the library does not link libcuda)
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn encode_tiled_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess && qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(ptr);
  }
  return fn;
}

int make_out_tensor_map(void* map128, void* out, int ncols, int ld_out, int N, int B, const char* who) {
  static_assert(sizeof(CUtensorMap) == 128, "CUtensorMap is 128 bytes");
  EncodeTiledFn enc = encode_tiled_fn();
  CP_REQUIRE(enc, CP_E_CUDA, "%s: cuTensorMapEncodeTiled is not available from this driver", who);
  const cuuint64_t dims[3] = {(cuuint64_t)ncols, (cuuint64_t)N, (cuuint64_t)B};
  const cuuint64_t strides[2] = {(cuuint64_t)ld_out * 2, (cuuint64_t)N * ld_out * 2};
  const cuuint32_t box[3] = {32, 32, 1}, estr[3] = {1, 1, 1};
  const CUresult r = enc(reinterpret_cast<CUtensorMap*>(map128), CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, out, dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  CP_REQUIRE(r == CUDA_SUCCESS, CP_E_CUDA, "%s: cuTensorMapEncodeTiled failed (%d)", who, (int)r);
  return CP_OK;
}
int make_bf16_operand_map(void* map128, const void* src, int64_t ncols, int64_t nrows, int64_t ld, const char* who, int box_rows) {
  EncodeTiledFn enc = encode_tiled_fn();
  CP_REQUIRE(enc, CP_E_CUDA, "%s: cuTensorMapEncodeTiled is not available from this driver", who);
  const cuuint64_t dims[2] = {(cuuint64_t)ncols, (cuuint64_t)nrows};
  const cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
  const cuuint32_t box[2] = {64, (cuuint32_t)box_rows}, estr[2] = {1, 1};
  const CUresult r = enc(reinterpret_cast<CUtensorMap*>(map128), CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(src), dims, strides, box,
                         estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  CP_REQUIRE(r == CUDA_SUCCESS, CP_E_CUDA, "%s: cuTensorMapEncodeTiled failed (%d)", who, (int)r);
  return CP_OK;
}

int make_f32_tensor_map_2d(void* map128, void* out, int64_t ncols, int64_t nrows, int64_t ld, const char* who) {
  EncodeTiledFn enc = encode_tiled_fn();
  CP_REQUIRE(enc, CP_E_CUDA, "%s: cuTensorMapEncodeTiled is not available from this driver", who);
  const cuuint64_t dims[2] = {(cuuint64_t)ncols, (cuuint64_t)nrows};
  const cuuint64_t strides[1] = {(cuuint64_t)ld * 4};
  const cuuint32_t box[2] = {32, 32}, estr[2] = {1, 1};
  const CUresult r = enc(reinterpret_cast<CUtensorMap*>(map128), CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, out, dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  CP_REQUIRE(r == CUDA_SUCCESS, CP_E_CUDA, "%s: cuTensorMapEncodeTiled failed (%d)", who, (int)r);
  return CP_OK;
}
}  // namespace cp

extern "C" const char* cp_last_error_string(void) { return cp::g_err; }
extern "C" int cp_version(void) { return 100; }
extern "C" int cp_device_arch(void) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) {
    cp::set_error("cp_device_arch: no CUDA device");
    (void)cudaGetLastError();
    return CP_E_CUDA;
  }
  int major = 0, minor = 0;
  cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
  cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev);
  return major * 10 + minor;
}
