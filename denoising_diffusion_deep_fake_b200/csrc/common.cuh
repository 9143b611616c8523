// common.cuh — shared device/host helpers for libd3fk (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include "../../include/d3fk.h"

namespace d3fk {

// ---- host-side bookkeeping -------------------------------------------------------------------
extern int64_t g_launch_count;
extern char g_last_error[512];
extern int* g_dev_error_flag;  // device int, set by kernel watchdogs

int set_error(int code, const char* fmt, ...);
int check_launch(const char* what);
inline void count_launch(int n = 1) { g_launch_count += n; }

#define D3FK_CHECK_ARG(cond, msg)                                        \
  do {                                                                   \
    if (!(cond)) return d3fk::set_error(D3FK_ERR_ARG, "%s: %s", __func__, msg); \
  } while (0)

static inline int cdiv(long long a, long long b) { return (int)((a + b - 1) / b); }

// ---- kernel launch: every libd3fk kernel goes through launch_k --------------------------------
// Programmatic dependent launch (PDL): the kernel may begin launching while its predecessor in the stream is still
// draining; every kernel therefore starts with pdl_enter() (griddepcontrol.wait = the predecessor has completed and its
// memory is visible; then griddepcontrol.launch_dependents = the successor may start its own launch / prologue).
// D3FK_PDL=0 turns the attribute off (the device-side instructions are then no-ops).
extern int g_use_pdl;
template <typename... KArgs, typename... Args>
static inline cudaError_t launch_k(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, dim3 cluster,
                                   Args&&... args) {
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = s;
  cudaLaunchAttribute attr[2];
  int na = 0;
  if (cluster.x * cluster.y * cluster.z > 1) {
    attr[na].id = cudaLaunchAttributeClusterDimension;
    attr[na].val.clusterDim.x = cluster.x;
    attr[na].val.clusterDim.y = cluster.y;
    attr[na].val.clusterDim.z = cluster.z;
    ++na;
  }
  if (g_use_pdl) {
    attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[na].val.programmaticStreamSerializationAllowed = 1;
    ++na;
  }
  cfg.attrs = attr;
  cfg.numAttrs = na;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}
#ifdef __CUDACC__
// grid-wide barrier for a co-resident grid (called by ONE thread per CTA, after its CTA's global writes / atomics)
__device__ __forceinline__ void grid_barrier_arrive_wait(unsigned* counter, unsigned expected, int* errflag) {
  // arrive: a RELEASE reduction at gpu scope (cumulative over what the CTA's other threads wrote before the CTA barrier that
  // precedes this call) — no returned value to wait for; wait: acquire loads, which order everything after them (the CTA
  // barrier that follows hands that order to the other threads), so no trailing fence.
  asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(counter) : "memory");
  const long long t0 = clock64();
  while (true) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(counter) : "memory");
    if (v >= expected) break;
    if (clock64() - t0 > 4000000000ll) {   // ~2 s: the grid was not co-resident — fatal (sticky CUDA error), never continue
      atomicExch(errflag, 2);
      printf("[d3fk] grid barrier timeout: block %d of %d saw %u arrivals (expected %u), counter %p\n", (int)blockIdx.x, (int)gridDim.x, v,
             expected, (void*)counter);
      __trap();
    }
  }
}

__device__ __forceinline__ void pdl_enter() {
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}
#endif

// ---- dtype traits ----------------------------------------------------------------------------
template <typename T> struct Vec;  // 16-byte vector of T
template <> struct Vec<float> {
  static constexpr int N = 4;
  float v[4];
};
template <> struct Vec<__nv_bfloat16> {
  static constexpr int N = 8;
  __nv_bfloat16 v[8];
};

template <typename T> __device__ __forceinline__ float to_f(T x);
template <> __device__ __forceinline__ float to_f<float>(float x) { return x; }
template <> __device__ __forceinline__ float to_f<__nv_bfloat16>(__nv_bfloat16 x) { return __bfloat162float(x); }
template <typename T> __device__ __forceinline__ T from_f(float x);
template <> __device__ __forceinline__ float from_f<float>(float x) { return x; }
template <> __device__ __forceinline__ __nv_bfloat16 from_f<__nv_bfloat16>(float x) { return __float2bfloat16_rn(x); }

// 16-byte vector load/store (addresses must be 16-byte aligned)
template <typename T> __device__ __forceinline__ void load_vec(const T* p, float* out) {
  constexpr int N = Vec<T>::N;
  uint4 raw = *reinterpret_cast<const uint4*>(p);
  const T* e = reinterpret_cast<const T*>(&raw);
#pragma unroll
  for (int i = 0; i < N; ++i) out[i] = to_f<T>(e[i]);
}
// the same in two steps, for loops that keep several vectors in flight: hold the packed 16 bytes (4 registers), widen at use
template <typename T> __device__ __forceinline__ uint4 load_raw(const T* p) { return *reinterpret_cast<const uint4*>(p); }
template <typename T> __device__ __forceinline__ void unpack_vec(const uint4& raw, float* out) {
  constexpr int N = Vec<T>::N;
  const T* e = reinterpret_cast<const T*>(&raw);
#pragma unroll
  for (int i = 0; i < N; ++i) out[i] = to_f<T>(e[i]);
}
template <typename T> __device__ __forceinline__ void store_vec(T* p, const float* in) {
  constexpr int N = Vec<T>::N;
  uint4 raw;
  T* e = reinterpret_cast<T*>(&raw);
#pragma unroll
  for (int i = 0; i < N; ++i) e[i] = from_f<T>(in[i]);
  *reinterpret_cast<uint4*>(p) = raw;
}

// ---- Philox4x32-10 ---------------------------------------------------------------------------
__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint2 k) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    uint32_t hi0 = __umulhi(M0, c.x), lo0 = M0 * c.x;
    uint32_t hi1 = __umulhi(M1, c.z), lo1 = M1 * c.z;
    c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
    k.x += W0;
    k.y += W1;
  }
  return c;
}
__device__ __forceinline__ float u32_to_uniform(uint32_t u) { return ((float)u + 0.5f) * 2.3283064365386963e-10f; }
// four standard normals from one Philox block (Box-Muller)
__device__ __forceinline__ float4 philox_normal4(uint64_t seed, uint64_t idx, uint64_t offset) {
  uint4 c = make_uint4((uint32_t)idx, (uint32_t)(idx >> 32), (uint32_t)offset, (uint32_t)(offset >> 32));
  uint4 r = philox4x32_10(c, make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
  float u0 = u32_to_uniform(r.x), u1 = u32_to_uniform(r.y), u2 = u32_to_uniform(r.z), u3 = u32_to_uniform(r.w);
  float ra = sqrtf(-2.f * __logf(u0)), rb = sqrtf(-2.f * __logf(u2));
  float s0, c0, s1, c1;
  __sincosf(6.283185307179586f * u1, &s0, &c0);
  __sincosf(6.283185307179586f * u3, &s1, &c1);
  return make_float4(ra * c0, ra * s0, rb * c1, rb * s1);
}

// ---- the A-operand gather shared by the FFMA and tcgen05 convolution kernels ------------------
struct Gather {
  const void* src0; const void* src1;
  int c0, c1, ld0, ld1, up0;
  int B, Hi, Wi, Ho, Wo, kh, kw, stride, pad, mode;
  int ctot;     // c0 + c1
  int K;        // kh*kw*ctot
  int M;        // B*Ho*Wo
  int sshift;   // log2(stride)
};

// Returns element offset (in elements of the selected source) or -1 when the tap is padding.
// h0/w0 are the per-row bases: mode 0: ho*stride - pad ; mode 1: ho + pad.
__device__ __forceinline__ long long gather_offset(const Gather& g, int n, int h0, int w0, int khi, int kwi, int c,
                                                   int& which) {
  int hi, wi;
  if (g.mode == 0) {
    hi = h0 + khi;
    wi = w0 + kwi;
  } else {
    int th = h0 - khi, tw = w0 - kwi;
    if ((th | tw) < 0) return -1;
    int mask = g.stride - 1;
    if ((th & mask) | (tw & mask)) return -1;
    hi = th >> g.sshift;
    wi = tw >> g.sshift;
  }
  if ((unsigned)hi >= (unsigned)g.Hi || (unsigned)wi >= (unsigned)g.Wi) return -1;
  if (c < g.c0) {
    which = 0;
    int hs = g.Hi >> g.up0, ws = g.Wi >> g.up0;
    return ((long long)(n * hs + (hi >> g.up0)) * ws + (wi >> g.up0)) * g.ld0 + c;
  }
  which = 1;
  return ((long long)(n * g.Hi + hi) * g.Wi + wi) * g.ld1 + (c - g.c0);
}

int make_gather(Gather& g, const void* src0, const void* src1, int c0, int c1, int ld0, int ld1, int up0, int B, int Hi,
                int Wi, int Ho, int Wo, int kh, int kw, int stride, int pad, int mode);

}  // namespace d3fk
