// What does a warp shuffle cost on this part?  Dependent chain, independent batch, the 31-shuffle transpose-reduce, the
// tcgen05.ld shapes with their results consumed.  One warp per SM sub-partition (128 threads), one block.
#include <cstdio>
#include <cuda_runtime.h>
#include "tc_common.cuh"
using namespace d3fk;
__device__ __forceinline__ void tmem_ld_16x256b_x4(uint32_t taddr, uint32_t* v) {
  asm volatile(
      "tcgen05.ld.sync.aligned.16x256b.x4.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr)
      : "memory");
}
__global__ void __launch_bounds__(256) probe(long long* out, float* sink, int nwarps) {
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0) tmem_alloc(smem_u32(&slot), 64);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_d = slot;
  if (warp >= nwarps) goto done;
  {
  float x = lane * 0.5f + 1.f;
  long long t[8];
  for (int rep = 0; rep < 3; ++rep) {
    t[0] = clock64();
#pragma unroll
    for (int i = 0; i < 32; ++i) x += __shfl_xor_sync(0xffffffffu, x, 1 + (i & 15));      // dependent chain of 32
    t[1] = clock64();
    float y[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) y[i] = __shfl_xor_sync(0xffffffffu, x + i, 16);           // 32 independent
    float acc = 0.f;
#pragma unroll
    for (int i = 0; i < 32; ++i) acc += y[i];
    t[2] = clock64();
#pragma unroll
    for (int i = 0; i < 32; ++i) y[i] = acc + i * x;
    float cs = warp_colsum32(y, lane);                                                     // 31-shuffle transpose-reduce
    t[3] = clock64();
    uint32_t raw[32];
    tmem_ld32(tmem_d + ((uint32_t)((warp & 3) * 32) << 16), raw);
    tmem_ld_wait();
    float s2 = 0.f;
#pragma unroll
    for (int i = 0; i < 32; ++i) s2 += __uint_as_float(raw[i]);
    t[4] = clock64();
    uint32_t a[16], b[16];
    tmem_ld_16x256b_x4(tmem_d + ((uint32_t)((warp & 3) * 32) << 16), a);
    tmem_ld_16x256b_x4(tmem_d + ((uint32_t)((warp & 3) * 32 + 16) << 16), b);
    tmem_ld_wait();
    float s3 = 0.f;
#pragma unroll
    for (int i = 0; i < 16; ++i) s3 += __uint_as_float(a[i]) + __uint_as_float(b[i]);
    t[5] = clock64();
    x = cs + s2 * 1e-30f + s3 * 1e-30f + acc * 1e-30f;
    if (x == 12345.f) x = 1.f;
  }
  if (lane == 0) for (int k = 0; k < 5; ++k) out[warp * 5 + k] = t[k + 1] - t[k];
  sink[threadIdx.x] = x;
  }
done:
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_d, 64);
}
int main() {
  long long* out; float* sink;
  cudaMalloc(&out, 8 * 5 * sizeof(long long));
  cudaMalloc(&sink, 256 * sizeof(float));
  for (int nw : {1, 4, 8}) {
    cudaMemset(out, 0, 8 * 5 * sizeof(long long));
    probe<<<1, 256>>>(out, sink, nw);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("error: %s\n", cudaGetErrorString(e)); return 1; }
    long long h[40];
    cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
    printf("%d active warps; warp 0 cycles: 32 dependent shfl %lld | 32 independent shfl + adds %lld | colsum32 (31 shfl) %lld | tcgen05.ld 32x32b.x32 + 32 adds %lld | 2 x 16x256b.x4 + adds %lld\n",
           nw, h[0], h[1], h[2], h[3], h[4]);
  }
  return 0;
}
