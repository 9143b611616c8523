// Epilogue micro-probe: cycles of the pieces of a tcgen05 epilogue, per warp, cold (first execution) and warm.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I denoising_diffusion_deep_fake_b200/csrc -o tools/probes/epi_probe tools/probes/epi_probe.cu
#include <cstdio>
#include <cuda_runtime.h>
#include "tc_common.cuh"
using namespace d3fk;

__global__ void __launch_bounds__(256) probe(__nv_bfloat16* out, int ld, long long* times, int reps) {
  __shared__ uint32_t slot;
  __shared__ float s_stat[64];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0) tmem_alloc(smem_u32(&slot), 128);
  if (threadIdx.x < 64) s_stat[threadIdx.x] = 0.f;
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_d = slot;
  const int q = warp & 3, half = warp >> 2;
  const long long m = (long long)blockIdx.x * 128 + q * 32 + lane;
  for (int r = 0; r < reps; ++r) {
    long long t0 = clock64();
    uint32_t raw[32];
    tmem_ld32(tmem_d + ((uint32_t)(q * 32) << 16) + half * 32, raw);
    tmem_ld_wait();
    long long t1 = clock64();
    float f[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) f[i] = __uint_as_float(raw[i]) * 0.5f;
    uint4* op = reinterpret_cast<uint4*>(out + m * ld + half * 32);
#pragma unroll
    for (int qq = 0; qq < 4; ++qq) {
      uint4 o;
      __nv_bfloat162* o2 = reinterpret_cast<__nv_bfloat162*>(&o);
#pragma unroll
      for (int i = 0; i < 4; ++i) o2[i] = __floats2bfloat162_rn(f[qq * 8 + 2 * i], f[qq * 8 + 2 * i + 1]);
      op[qq] = o;
    }
    long long t2 = clock64();
    float sq[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) sq[i] = f[i] * f[i];
    float cs = warp_colsum32(f, lane), cq = warp_colsum32(sq, lane);
    s_stat[lane] += cs; s_stat[32 + lane] += cq;
    long long t3 = clock64();
    __syncwarp();
    if (lane == 0) {
      long long* t = times + (((long long)blockIdx.x * 8 + warp) * reps + r) * 4;
      t[0] = t1 - t0; t[1] = t2 - t1; t[2] = t3 - t2; t[3] = t3 - t0;
    }
  }
  if (threadIdx.x == 0 && s_stat[0] == 123.f) out[0] = __float2bfloat16(s_stat[1]);
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_d, 128);
}

int main() {
  const int reps = 4, blocks = 148, ld = 128;
  __nv_bfloat16* out; long long* times;
  cudaMalloc(&out, (size_t)blocks * 128 * ld * 2);
  cudaMalloc(&times, sizeof(long long) * blocks * 8 * reps * 4);
  long long* h = new long long[blocks * 8 * reps * 4];
  for (int launch = 0; launch < 3; ++launch) {
    probe<<<blocks, 256>>>(out, ld, times, reps);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("error: %s\n", cudaGetErrorString(e)); return 1; }
    cudaMemcpy(h, times, sizeof(long long) * blocks * 8 * reps * 4, cudaMemcpyDeviceToHost);
    printf("launch %d (cycles: tmem_ld+wait | cvt+4xSTG.128 issue | stats 2x colsum32 | total), block 0 warp 0, then mean over all warps\n", launch);
    for (int r = 0; r < reps; ++r) {
      double mean[4] = {0, 0, 0, 0};
      for (int b = 0; b < blocks * 8; ++b) for (int k = 0; k < 4; ++k) mean[k] += (double)h[((long long)b * reps + r) * 4 + k] / (blocks * 8);
      printf("  rep %d: b0w0 %5lld %5lld %5lld %5lld   mean %7.0f %7.0f %7.0f %7.0f\n", r, h[r * 4], h[r * 4 + 1], h[r * 4 + 2], h[r * 4 + 3],
             mean[0], mean[1], mean[2], mean[3]);
    }
  }
  return 0;
}
