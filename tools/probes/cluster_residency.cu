// Probe: how many CTAs of a cluster launch are co-resident on a B200?  Each CTA records smid + start/end globaltimer
// and spins ~30 us; the host reports the peak number of simultaneously running CTAs and the SMs they used.
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <algorithm>
#include <cuda_runtime.h>
struct Rec { unsigned long long t0, t1; unsigned smid, rank; };
__global__ void probe(Rec* out, int spin_us, int use_tmem) {
  extern __shared__ unsigned char smem[];
  __shared__ unsigned tmem_slot;
  unsigned long long t0, t1;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
  if (use_tmem && threadIdx.x < 32) {
    unsigned addr = (unsigned)__cvta_generic_to_shared(&tmem_slot);
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(addr), "r"(128) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  __syncthreads();
  smem[threadIdx.x] = 1;
  do { asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1)); } while (t1 - t0 < (unsigned long long)spin_us * 1000ull);
  __syncthreads();
  if (use_tmem && threadIdx.x < 32) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_slot), "r"(128) : "memory");
  }
  if (threadIdx.x == 0) {
    unsigned smid, rank;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
    out[blockIdx.x] = Rec{t0, t1, smid, rank};
  }
}
int main() {
  const int grid = 296;
  Rec* d; cudaMalloc(&d, grid * sizeof(Rec));
  for (int use_tmem = 0; use_tmem < 2; ++use_tmem)
  for (int smem_kb : {40, 100, 200})
  for (int cl : {0, 1, 2, 4, 8}) {
    size_t smem = (size_t)smem_kb * 1024;
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaFuncSetAttribute(probe, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    cudaMemset(d, 0, grid * sizeof(Rec));
    cudaError_t e;
    if (cl == 0) { probe<<<grid, 160, smem>>>(d, 30, use_tmem); e = cudaGetLastError(); }
    else {
      cudaLaunchConfig_t cfg = {}; cfg.gridDim = dim3(grid); cfg.blockDim = dim3(160); cfg.dynamicSmemBytes = smem;
      cudaLaunchAttribute a[1]; a[0].id = cudaLaunchAttributeClusterDimension; a[0].val.clusterDim.x = cl; a[0].val.clusterDim.y = 1; a[0].val.clusterDim.z = 1;
      cfg.attrs = a; cfg.numAttrs = 1;
      e = cudaLaunchKernelEx(&cfg, probe, d, 30, use_tmem);
    }
    cudaError_t e2 = cudaDeviceSynchronize();
    std::vector<Rec> h(grid); cudaMemcpy(h.data(), d, grid * sizeof(Rec), cudaMemcpyDeviceToHost);
    std::vector<std::pair<unsigned long long,int>> ev; std::vector<int> per_sm(256, 0);
    unsigned long long tmin = ~0ull, tmax = 0;
    for (auto& r : h) { ev.push_back({r.t0, 1}); ev.push_back({r.t1, -1}); tmin = std::min(tmin, r.t0); tmax = std::max(tmax, r.t1); }
    std::sort(ev.begin(), ev.end());
    int cur = 0, peak = 0; for (auto& x : ev) { cur += x.second; peak = std::max(peak, cur); }
    // CTAs started within the first 5 us = first wave
    int first = 0; std::vector<int> sm_first(256, 0);
    for (auto& r : h) if (r.t0 - tmin < 5000) { ++first; sm_first[r.smid]++; }
    int sms = 0, max_per_sm = 0; for (int i = 0; i < 256; ++i) if (sm_first[i]) { ++sms; max_per_sm = std::max(max_per_sm, sm_first[i]); }
    printf("tmem=%d smem=%3dK cluster=%d: launch=%s sync=%s  total %.1f us  peak concurrent %d  first wave %d CTAs on %d SMs (max %d per SM)\n",
           use_tmem, smem_kb, cl, cudaGetErrorString(e), cudaGetErrorString(e2), (tmax - tmin) / 1000.0, peak, first, sms, max_per_sm);
  }
  return 0;
}
