// tma.cu — cached TMA descriptor construction through the driver entry point (no -lcuda needed).
#include <map>
#include <mutex>
#include <string.h>
#include <vector>
#include "common.cuh"
#include "tma.cuh"

namespace d3fk {

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn g_encode = nullptr;
static std::mutex g_mu;
static std::map<std::vector<uint64_t>, CUtensorMap> g_cache;

static int load_entry_point() {
  if (g_encode) return D3FK_OK;
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
  if (e != cudaSuccess || q != cudaDriverEntryPointSuccess || !fn)
    return set_error(D3FK_ERR_CUDA, "cuTensorMapEncodeTiled entry point unavailable (%s)", cudaGetErrorString(e));
  g_encode = (EncodeTiledFn)fn;
  return D3FK_OK;
}

int get_tensor_map(CUtensorMap* out, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                   const uint32_t* box, int swizzle_bytes) {
  std::lock_guard<std::mutex> lock(g_mu);
  int rc = load_entry_point();
  if (rc) return rc;
  std::vector<uint64_t> key;
  key.push_back((uint64_t)(uintptr_t)base);
  key.push_back((uint64_t)rank * 1000 + swizzle_bytes);
  for (int i = 0; i < rank; ++i) key.push_back(dims[i]);
  for (int i = 0; i + 1 < rank; ++i) key.push_back(strides_bytes[i]);
  for (int i = 0; i < rank; ++i) key.push_back(box[i]);
  auto it = g_cache.find(key);
  if (it != g_cache.end()) {
    memcpy(out, &it->second, sizeof(CUtensorMap));
    return D3FK_OK;
  }
  alignas(64) CUtensorMap m;
  cuuint64_t gd[5], gs[5];
  cuuint32_t bx[5], es[5];
  for (int i = 0; i < rank; ++i) { gd[i] = dims[i]; bx[i] = box[i]; es[i] = 1; }
  for (int i = 0; i + 1 < rank; ++i) gs[i] = strides_bytes[i];
  CUtensorMapSwizzle sw = swizzle_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B
                          : swizzle_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B
                          : swizzle_bytes == 32 ? CU_TENSOR_MAP_SWIZZLE_32B : CU_TENSOR_MAP_SWIZZLE_NONE;
  CUresult r = g_encode(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void*>(base), gd, gs, bx, es,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return set_error(D3FK_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d) rank=%d", (int)r, rank);
  if (g_cache.size() > 65536) g_cache.clear();
  g_cache[key] = m;
  memcpy(out, &m, sizeof(CUtensorMap));
  return D3FK_OK;
}

}  // namespace d3fk
