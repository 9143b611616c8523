// Which accumulator element does register r of thread t receive from tcgen05.ld.16x256b.x4?  Writes a known pattern with
// tcgen05.st.32x32b (thread = row) and reads it back in the 16x256b shape; then checks and times the column statistics
// (sum, sum of squares over the 32 rows of a warp) built on that shape against the 32x32b + 31-shuffle transpose-reduce.
#include <cstdio>
#include <cuda_runtime.h>
#include "tc_common.cuh"
using namespace d3fk;

__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t* v) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
      ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]),
        "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]), "r"(v[16]), "r"(v[17]), "r"(v[18]), "r"(v[19]),
        "r"(v[20]), "r"(v[21]), "r"(v[22]), "r"(v[23]), "r"(v[24]), "r"(v[25]), "r"(v[26]), "r"(v[27]), "r"(v[28]), "r"(v[29]),
        "r"(v[30]), "r"(v[31])
      : "memory");
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_ld_16x256b_x4(uint32_t taddr, uint32_t* v) {
  asm volatile(
      "tcgen05.ld.sync.aligned.16x256b.x4.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr)
      : "memory");
}

__global__ void __launch_bounds__(128) probe(int* map, float* stats_out, long long* cyc) {
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0) tmem_alloc(smem_u32(&slot), 64);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_d = slot;
  const int row = warp * 32 + lane;
  uint32_t v[32];
#pragma unroll
  for (int c = 0; c < 32; ++c) v[c] = (uint32_t)(row * 64 + c);
  tmem_st32(tmem_d + ((uint32_t)(warp * 32) << 16), v);
  __syncwarp();
  uint32_t a[16], b[16];
  tmem_ld_16x256b_x4(tmem_d + ((uint32_t)(warp * 32) << 16), a);
  tmem_ld_16x256b_x4(tmem_d + ((uint32_t)(warp * 32 + 16) << 16), b);
  tmem_ld_wait();
  for (int r = 0; r < 16; ++r) {
    map[(threadIdx.x * 32 + r) * 2 + 0] = (int)a[r] / 64;   // row
    map[(threadIdx.x * 32 + r) * 2 + 1] = (int)a[r] % 64;   // col
    map[(threadIdx.x * 32 + 16 + r) * 2 + 0] = (int)b[r] / 64;
    map[(threadIdx.x * 32 + 16 + r) * 2 + 1] = (int)b[r] % 64;
  }
  // ---- float pattern for the statistics check: x(row, col) = 0.01 * row + col
  float f[32];
#pragma unroll
  for (int c = 0; c < 32; ++c) { f[c] = 0.01f * row + c; v[c] = __float_as_uint(f[c]); }
  tmem_st32(tmem_d + ((uint32_t)(warp * 32) << 16), v);
  __syncwarp();
  for (int rep = 0; rep < 3; ++rep) {
    long long t0 = clock64();
    tmem_ld_16x256b_x4(tmem_d + ((uint32_t)(warp * 32) << 16), a);
    tmem_ld_16x256b_x4(tmem_d + ((uint32_t)(warp * 32 + 16) << 16), b);
    tmem_ld_wait();
    // assumed layout (mma C fragment): regs 4j+{0,1}: row t/4, cols 8j + 2(t%4) + {0,1}; regs 4j+{2,3}: row t/4 + 8
    float s[16];   // [0..7] sums, [8..15] sums of squares for cols 8j + 2(t%4) + e, index j*2+e
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const float x0 = __uint_as_float(a[4 * j + e]), x1 = __uint_as_float(a[4 * j + 2 + e]);
        const float x2 = __uint_as_float(b[4 * j + e]), x3 = __uint_as_float(b[4 * j + 2 + e]);
        s[j * 2 + e] = (x0 + x1) + (x2 + x3);
        s[8 + j * 2 + e] = fmaf(x0, x0, x1 * x1) + fmaf(x2, x2, x3 * x3);
      }
    // transpose-reduce over the 8 lanes that share t % 4 (lane bits 2..4): 8 + 4 + 2 shuffles
#pragma unroll
    for (int ofs = 16, n = 16; ofs >= 4; ofs >>= 1, n >>= 1) {
      const bool up = (lane & ofs) != 0;
#pragma unroll
      for (int i = 0; i < n / 2; ++i) {
        const float send = up ? s[i] : s[i + n / 2];
        const float keep = up ? s[i + n / 2] : s[i];
        s[i] = keep + __shfl_xor_sync(0xffffffffu, send, ofs);
      }
    }
    // now s[0], s[1] hold two fully reduced values: which ones?  index bits: (lane&16 ? hi8 : lo8), (lane&8 ? ..), (lane&4 ? ..)
    long long t1 = clock64();
    if (rep == 2) {
      // original index of s[k] (k = 0, 1): idx = ((lane >> 4) & 1) * 8 + ((lane >> 3) & 1) * 4 + ((lane >> 2) & 1) * 2 + k
      for (int k = 0; k < 2; ++k) {
        const int idx = ((lane >> 4) & 1) * 8 + ((lane >> 3) & 1) * 4 + ((lane >> 2) & 1) * 2 + k;
        const int which = idx >> 3, jj = (idx & 7) >> 1, e = idx & 1;
        const int col = 8 * jj + 2 * (lane & 3) + e;
        stats_out[(warp * 2 + which) * 32 + col] = s[k];
      }
      if (lane == 0) cyc[warp] = t1 - t0;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_d, 64);
}

int main() {
  int* map; float* st; long long* cyc;
  cudaMalloc(&map, 128 * 32 * 2 * sizeof(int));
  cudaMalloc(&st, 4 * 2 * 32 * sizeof(float));
  cudaMalloc(&cyc, 4 * sizeof(long long));
  cudaMemset(st, 0, 4 * 2 * 32 * sizeof(float));
  probe<<<1, 128>>>(map, st, cyc);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("error: %s\n", cudaGetErrorString(e)); return 1; }
  static int h[128 * 32 * 2]; static float hs[4 * 2 * 32]; long long hc[4];
  cudaMemcpy(h, map, sizeof(h), cudaMemcpyDeviceToHost);
  cudaMemcpy(hs, st, sizeof(hs), cudaMemcpyDeviceToHost);
  cudaMemcpy(hc, cyc, sizeof(hc), cudaMemcpyDeviceToHost);
  for (int t : {0, 1, 2, 3, 4, 5, 31, 32, 37}) {
    printf("thread %3d lanes+0 :", t);
    for (int r = 0; r < 16; ++r) printf(" (%d,%d)", h[(t * 32 + r) * 2], h[(t * 32 + r) * 2 + 1]);
    printf("\n           lanes+16:");
    for (int r = 0; r < 16; ++r) printf(" (%d,%d)", h[(t * 32 + 16 + r) * 2], h[(t * 32 + 16 + r) * 2 + 1]);
    printf("\n");
  }
  int bad = 0;
  for (int t = 0; t < 128; ++t)
    for (int r = 0; r < 32; ++r) {
      const int w = t / 32, l = t % 32, rr = r % 16, hi = r / 16;
      const int j = rr / 4, e2 = rr % 2, up = (rr % 4) / 2;
      const int erow = w * 32 + hi * 16 + l / 4 + 8 * up, ecol = 8 * j + 2 * (l % 4) + e2;
      if (h[(t * 32 + r) * 2] != erow || h[(t * 32 + r) * 2 + 1] != ecol) ++bad;
    }
  printf("layout assumption (mma C fragment): %s (%d mismatches)\n", bad ? "WRONG" : "confirmed", bad);
  double worst = 0;
  for (int w = 0; w < 4; ++w)
    for (int c = 0; c < 32; ++c) {
      double s1 = 0, s2 = 0;
      for (int r = 0; r < 32; ++r) { const double x = (double)(0.01f * (w * 32 + r) + c); s1 += x; s2 += x * x; }
      worst = fmax(worst, fabs(hs[(w * 2) * 32 + c] - s1) / s1);
      worst = fmax(worst, fabs(hs[(w * 2 + 1) * 32 + c] - s2) / s2);
    }
  printf("column statistics via 16x256b + 14 shuffles: worst rel err %.3e; cycles per warp (2 loads + reduce): %lld %lld %lld %lld\n", worst, hc[0], hc[1], hc[2], hc[3]);
  return 0;
}
