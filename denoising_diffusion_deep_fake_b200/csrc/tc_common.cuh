// tc_common.cuh — what the tcgen05 translation units share (conv_tc.cu, conv_slab.cu, wgrad_tc.cu): PTX wrappers
// (mbarrier, cp.async, tcgen05 alloc / mma / commit / ld, cluster + DSMEM), UMMA descriptors, the fused epilogue and the
// launch-geometry globals.  Split out of conv_tc.cu so that the three kernel families compile in parallel.
#pragma once
#include <stdlib.h>
#include <string.h>
#include "common.cuh"
#include "tma.cuh"

namespace d3fk {

typedef __nv_bfloat16 bf16;

// -DD3FK_TIMELINE (tools/build_timeline.sh; never shipped): CTA 0 of every conv_tc launch stamps %globaltimer at its
// phase boundaries into the debug buffer behind the error flag — where do the ~10 us of a tiny tensor-core kernel go?
#ifdef D3FK_TIMELINE
#define TL_SLOTS 16
#define TL_MAX 512
__device__ __forceinline__ unsigned long long tl_now() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
#define TL_DECL unsigned long long* tl_buf = reinterpret_cast<unsigned long long*>(errflag + 16); __shared__ unsigned tl_idx;
#define TL_BEGIN if (blockIdx.x == 0 && threadIdx.x == 0) { tl_idx = atomicAdd(reinterpret_cast<unsigned*>(errflag + 4), 1u) % TL_MAX; tl_buf[tl_idx * TL_SLOTS + 0] = tl_now(); }
#define TL_STAMP(slot) if (blockIdx.x == 0) { tl_buf[tl_idx * TL_SLOTS + (slot)] = tl_now(); }
#else
#define TL_DECL
#define TL_BEGIN
#define TL_STAMP(slot)
#endif

// ------------------------------------------------------------------------------------------
// PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// One lane of a fully converged warp (elect.sync).  Every single-thread issue region (tcgen05.mma / commit, TMA loads)
// is entered through this and NOT through `lane == 0`: inside an elect region ptxas knows that one thread is active and
// moves descriptor / coordinate words into uniform registers with a plain R2UR (or computes them on the uniform datapath
// outright); behind a lane compare it wraps EVERY UTCHMMA / UTMALDG in an ELECT + R2UR.BROADCAST + BRA.U.ANY waterfall
// loop — ~31 SASS instructions per MMA instead of ~2 (ncu source page of conv_slab_kernel, profiles/r02_issue_loop.txt).
__device__ __forceinline__ bool elect_one_sync() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}
// warp index as a value the compiler can prove warp-uniform
__device__ __forceinline__ int warp_index_uniform() { return __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
// arrive-on triggered when all prior cp.async of this thread have landed (pending count +1 now, -1 then)
__device__ __forceinline__ void cp_async_mbar_arrive(uint32_t bar) {
  asm volatile("cp.async.mbarrier.arrive.shared::cta.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ uint32_t mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok;
}
// NOTE on proxies: the cp.async (LDGSTS) gather writes shared memory and tcgen05.mma reads it through the async proxy.
// There is deliberately no fence.proxy.async between the full-barrier wait and the MMA: completion of the copies is
// tracked by the mbarrier itself (cp.async.mbarrier.arrive -> ARRIVES.LDGSTSBAR), exactly the protocol of CUTLASS's
// sm100 cp.async main loop (cutlass/gemm/collective/sm100_mma_cpasync_warpspecialized.hpp:499-566: producer_commit with
// cpasync_barrier_arrive, consumer_wait, then cute::gemm -> tcgen05.mma, no proxy fence in between).  The fence lowers to
// MEMBAR.ALL.CTA, which drains every in-flight LDGSTS of the CTA and serialises the whole pipeline.  Every byte of a
// stage is written by cp.async (padding uses the zero-fill form), never by st.shared; tests/test_gpu_ops.py::
// test_cp_async_path_is_race_free re-runs ragged shapes 1000 times and demands bit-identical results.
//
// Bounded wait: a barrier that does not complete within ~2 s is a protocol bug or a lost TMA transaction.  The kernel
// sets the device error flag and TRAPS: the launch fails with a sticky CUDA error that the next libd3fk call (and the
// next CUDA call of the host framework) reports — garbage is never produced at full speed.
static __device__ __noinline__ void mbar_timeout(int* errflag, uint32_t bar, uint32_t parity) {
  atomicExch(errflag, 1);
  printf("[d3fk] mbarrier watchdog: block %d thread %d waited > 2 s on barrier 0x%x parity %u\n", (int)blockIdx.x,
         (int)threadIdx.x, bar, parity);
  __trap();
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity, int* errflag) {
  if (mbar_try_wait(bar, parity)) return;
  unsigned long long t0;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
  for (uint32_t it = 1;; ++it) {
    if (mbar_try_wait(bar, parity)) return;
    if ((it & 1023u) == 0) {
      unsigned long long t;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
      if (t - t0 > 2000000000ull) mbar_timeout(errflag, bar, parity);
    }
  }
}
__device__ __forceinline__ void cp_async_16(uint32_t dst, const void* src, uint32_t bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(bytes) : "memory");
}
// L1-allocating variant: the 3x3 taps of a tile re-read the same input lines, so the activation gather can hit L1
__device__ __forceinline__ void cp_async_16_ca(uint32_t dst, const void* src, uint32_t bytes) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tmem_alloc(uint32_t smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_dst), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// D[tmem] (+)= A[smem desc] * B[smem desc]
__device__ __forceinline__ void umma_f16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// same, descriptors given as (lo, hi) words: the issue loop only adds to the 14-bit address field of the low word
__device__ __forceinline__ void umma_f16_lohi(uint32_t d_tmem, uint32_t alo, uint32_t ahi, uint32_t blo, uint32_t bhi, uint32_t idesc,
                                              uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
      "mov.b64 da, {%1, %2};\n\t"
      "mov.b64 db, {%3, %4};\n\t"
      "setp.ne.b32 p, %6, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n\t}"
      ::"r"(d_tmem), "r"(alo), "r"(ahi), "r"(blo), "r"(bhi), "r"(idesc), "r"(accumulate)
      : "memory");
}
// mbarrier arrive once all previously issued tcgen05.mma of this thread have completed
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t* v) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t* v) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// ---- thread-block cluster / distributed shared memory helpers (split-K reduction across the CTAs of a cluster)
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
// all threads of all CTAs of the cluster; release/acquire orders the DSMEM traffic around it
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// The same barrier without memory ordering on the arrive side (no MEMBAR.ALL.GPU + ERRBAR in front of UCGABAR_ARV): for
// phases that only order control flow ("every CTA of the cluster is past its main loop").
__device__ __forceinline__ void cluster_sync_relaxed() {
  asm volatile("barrier.cluster.arrive.relaxed.aligned;\n\tbarrier.cluster.wait.aligned;" ::: "memory");
}
// Asynchronous 16-byte store into the shared memory of a CTA of the cluster; its completion is counted (complete_tx, 16
// bytes) on an mbarrier of the DESTINATION CTA — the receiver waits on its own barrier, no cluster-wide barrier or fence.
__device__ __forceinline__ void st_async_f4(uint32_t remote_addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d, uint32_t remote_mbar) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.b32 [%0], {%1, %2, %3, %4}, [%5];"
               ::"r"(remote_addr), "r"(a), "r"(b), "r"(c), "r"(d), "r"(remote_mbar) : "memory");
}
__device__ __forceinline__ uint32_t mapa_shared(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void st_cluster_f4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared::cluster.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ float4 ld_shared_f4(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr) : "memory");
  return v;
}

// Shared-memory matrix descriptor (cute::UMMA::SmemDescriptor layout), SWIZZLE_128B, version 1.
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;
  d |= (uint64_t)1 << 46;  // descriptor version (Blackwell)
  d |= (uint64_t)2 << 61;  // SWIZZLE_128B
  return d;
}
// Instruction descriptor (cute::UMMA::InstrDescriptor): bf16 x bf16 -> f32, M x N, majors selectable.
__host__ __device__ constexpr uint32_t make_idesc(int M, int N, int a_mn_major, int b_mn_major) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)a_mn_major << 15) | ((uint32_t)b_mn_major << 16) |
         ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// transpose-reduce: on return lane l holds the sum over the 32 lanes of v[l]
__device__ __forceinline__ float warp_colsum32(float* v, int lane) {
#pragma unroll
  for (int ofs = 16, n = 32; ofs >= 1; ofs >>= 1, n >>= 1) {
    const bool up = (lane & ofs) != 0;
#pragma unroll
    for (int i = 0; i < n / 2; ++i) {
      float send = up ? v[i] : v[i + n / 2];
      float keep = up ? v[i + n / 2] : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, ofs);
    }
  }
  return v[0];
}
// 16 columns: lanes l and l+16 both end with the sum of column (l & 15)
__device__ __forceinline__ float warp_colsum16(float* v, int lane) {
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] += __shfl_xor_sync(0xffffffffu, v[i], 16);
#pragma unroll
  for (int ofs = 8, n = 16; ofs >= 1; ofs >>= 1, n >>= 1) {
    const bool up = (lane & ofs) != 0;
#pragma unroll
    for (int i = 0; i < n / 2; ++i) {
      float send = up ? v[i] : v[i + n / 2];
      float keep = up ? v[i + n / 2] : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, ofs);
    }
  }
  return v[0];
}

// ---- per-channel batch statistics straight from tensor memory
// tcgen05.ld.16x256b hands a warp its 32 accumulator rows in the mma C-fragment arrangement (measured with
// tools/probes/tmem_layout_probe.cu): of one load at lane base L, thread t holds rows L + t/4 and L + t/4 + 8 and, per
// 8-column block j, the columns 8j + 2(t%4) + {0, 1} (registers 4j + {0,1} | {2,3}).  A thread therefore sums FOUR rows of
// its column pair in registers and the remaining reduction runs over the 8 lanes that share t % 4: 8 + 4 + 2 shuffles for
// the sums AND the sums of squares of 32 columns, against 2 x 31 for the transpose-reduce of the row-per-thread (32x32b)
// arrangement — which, queued behind the epilogue's scattered 16-byte stores in the same memory-instruction pipe, took
// 0.7 us per 32-column chunk and was the largest single piece of every training-forward epilogue.
template <int NREP> __device__ __forceinline__ void tmem_ld_16x256b(uint32_t taddr, uint32_t* v);
template <> __device__ __forceinline__ void tmem_ld_16x256b<4>(uint32_t taddr, uint32_t* v) {
  asm volatile(
      "tcgen05.ld.sync.aligned.16x256b.x4.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr)
      : "memory");
}
template <> __device__ __forceinline__ void tmem_ld_16x256b<2>(uint32_t taddr, uint32_t* v) {
  asm volatile("tcgen05.ld.sync.aligned.16x256b.x2.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
               : "r"(taddr)
               : "memory");
}
// 16 fp32 columns of this thread's accumulator row back into tensor memory (the split-K owner parks its reduced slice so
// that the statistics can be taken in the 16x256b arrangement)
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const float* f) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
      ::"r"(taddr), "f"(f[0]), "f"(f[1]), "f"(f[2]), "f"(f[3]), "f"(f[4]), "f"(f[5]), "f"(f[6]), "f"(f[7]), "f"(f[8]), "f"(f[9]),
        "f"(f[10]), "f"(f[11]), "f"(f[12]), "f"(f[13]), "f"(f[14]), "f"(f[15])
      : "memory");
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}
// Sums and sums of squares of CW (32 or 16) accumulator columns starting at taddr's column, over the 32 rows of this warp's
// lane quarter (taddr carries the quarter's lane base), ACCUMULATED into sstat_sum[0..CW) / sstat_sq[0..CW) (one owner
// lane per slot).  row_limit: rows (relative to the quarter's first row) >= row_limit are excluded.
template <int CW>
__device__ __forceinline__ void tmem_col_stats(uint32_t taddr, int row_limit, float* sstat_sum, float* sstat_sq, int lane) {
  constexpr int J = CW / 8;                 // 8-column blocks
  uint32_t a[4 * J], b[4 * J];
  tmem_ld_16x256b<J>(taddr, a);
  tmem_ld_16x256b<J>(taddr + (16u << 16), b);
  tmem_ld_wait();
  const int r0 = lane >> 2;
  const bool v0 = r0 < row_limit, v1 = r0 + 8 < row_limit, v2 = r0 + 16 < row_limit, v3 = r0 + 24 < row_limit;
  float s[4 * J];                           // [0, 2J): sums, [2J, 4J): sums of squares; index j * 2 + e
#pragma unroll
  for (int j = 0; j < J; ++j)
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const float x0 = v0 ? __uint_as_float(a[4 * j + e]) : 0.f, x1 = v1 ? __uint_as_float(a[4 * j + 2 + e]) : 0.f;
      const float x2 = v2 ? __uint_as_float(b[4 * j + e]) : 0.f, x3 = v3 ? __uint_as_float(b[4 * j + 2 + e]) : 0.f;
      s[j * 2 + e] = (x0 + x1) + (x2 + x3);
      s[2 * J + j * 2 + e] = fmaf(x0, x0, x1 * x1) + fmaf(x2, x2, x3 * x3);
    }
  // transpose-reduce over the lanes that share lane % 4 (lane bits 4, 3, 2): the surviving index bits are those lane bits
#define D3FK_TR_LEVEL(ofs, n)                                                 \
  {                                                                           \
    const bool up = (lane & (ofs)) != 0;                                      \
    _Pragma("unroll") for (int i = 0; i < (n) / 2; ++i) {                     \
      const float send = up ? s[i] : s[i + (n) / 2];                          \
      const float keep = up ? s[i + (n) / 2] : s[i];                          \
      s[i] = keep + __shfl_xor_sync(0xffffffffu, send, (ofs));                \
    }                                                                         \
  }
  D3FK_TR_LEVEL(16, 4 * J)
  D3FK_TR_LEVEL(8, 2 * J)
  D3FK_TR_LEVEL(4, J)
#undef D3FK_TR_LEVEL
  const int which = (lane >> 4) & 1;
  float* dst = which ? sstat_sq : sstat_sum;
  if (CW == 32) {                           // idx = which:1 | j:2 | e:1 ; two survivors (e = 0, 1)
    const int j = ((lane >> 3) & 1) * 2 + ((lane >> 2) & 1);
    const int col = 8 * j + 2 * (lane & 3);
    dst[col] += s[0];
    dst[col + 1] += s[1];
  } else {                                  // idx = which:1 | j:1 | e:1 ; one survivor
    const int col = 8 * ((lane >> 3) & 1) + 2 * (lane & 3) + ((lane >> 2) & 1);
    dst[col] += s[0];
  }
}

struct EpiTC {
  bf16* out; float* out_nchw; const float* scale; const float* shift; const bf16* res; double* stats;
  int ldo, ldr, relu, Cout, Ho, Wo;
  // BN-backward reduction fused into a dgrad epilogue (bw_x set): the stored output v is the gradient wrt the activation of
  // the previous layer; the statistics become sum(g') and sum(g' * xhat) with g' = v masked by that layer's ReLU
  // (bw_act > 0) and xhat = (bw_x - mean) * invstd — exactly what bn_bwd_reduce would compute in a second pass.
  const bf16* bw_x; const bf16* bw_act; const float* bw_mean; const float* bw_invstd;
  int bw_ldx, bw_ldact, bw_relu;
};

// in place: f -> g' ; sq -> g' * x (rows outside the tensor contribute zero).  The per-channel mean / invstd are applied
// once per CTA when the partial sums are flushed: sum(g' * xhat) = invstd * (sum(g' * x) - mean * sum(g')).
template <int CW>
__device__ __forceinline__ void bw_stat_terms(float (&f)[CW], float (&sq)[CW], const EpiTC& e, long long m, bool row_ok, int cbase) {
  if (!row_ok) {
#pragma unroll
    for (int i = 0; i < CW; ++i) { f[i] = 0.f; sq[i] = 0.f; }
    return;
  }
#pragma unroll
  for (int q8 = 0; q8 < CW / 8; ++q8) {
    const uint4 xr = __ldg(reinterpret_cast<const uint4*>(e.bw_x + m * e.bw_ldx + cbase + q8 * 8));
    const bf16* xb = reinterpret_cast<const bf16*>(&xr);
    uint4 ar = make_uint4(0u, 0u, 0u, 0u);
    if (e.bw_relu) ar = __ldg(reinterpret_cast<const uint4*>(e.bw_act + m * e.bw_ldact + cbase + q8 * 8));
    const bf16* ab = reinterpret_cast<const bf16*>(&ar);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int c = q8 * 8 + i;
      float gsel = f[c];
      if (e.bw_relu && !(__bfloat162float(ab[i]) > 0.f)) gsel = 0.f;
      f[c] = gsel;
      sq[c] = gsel * __bfloat162float(xb[i]);
    }
  }
}
// per-CTA partial sums (a = sum g', b = sum g' * x) of channel ch -> the second BN-backward sum
__device__ __forceinline__ double bw_second_sum(double a, double b, const EpiTC& e, int ch) {
  return (double)__ldg(e.bw_invstd + ch) * (b - (double)__ldg(e.bw_mean + ch) * a);
}

__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

struct FastDiv {
  uint32_t mul, shr;
};
static FastDiv make_fastdiv(uint32_t d) {
  FastDiv f;
  uint32_t l = 0;
  while ((1ull << l) < d) ++l;
  f.shr = l;
  f.mul = (uint32_t)(((1ull << 32) * ((1ull << l) - d)) / d + 1);
  return f;
}
__device__ __forceinline__ uint32_t fdiv(uint32_t n, FastDiv f) { return (__umulhi(f.mul, n) + n) >> f.shr; }

constexpr int TC_THREADS = 320;   // warps 0-3 gather/epilogue, warp 4 MMA issuer, warp 5 TMA producer, warps 6-9 epilogue helpers
constexpr int EPI_THREADS = 256;  // the 8 epilogue warps
constexpr int WG_THREADS = 160;   // weight-gradient kernel: warps 0-3 gather/epilogue, warp 4 MMA issuer
constexpr int TC_BM = 128;
constexpr int TC_BK = 64;                 // bf16 elements = 128 bytes = one swizzle row
constexpr int A_STAGE_BYTES = TC_BM * 128;

// Epilogue of CW accumulator columns [cbase, cbase+CW) of output row m held in f[]: folded-BN affine / bias, residual,
// ReLU, store (bf16 NHWC or fp32 NCHW) and per-channel batch statistics.  The statistics are folded over the 32 rows of
// the warp with a transpose-reduce and ACCUMULATED into this warp's shared-memory slots (sstat_warp[col], [BN + col]);
// the CTA flushes the slots to global memory with one double atomic per channel when its n tile changes / at the end.
// s_aff (nullable): the chunk's affine coefficients staged in shared memory by the CTA — s_aff[0..CW) = scale (1 for a plain
// bias), s_aff[aff_stride + 0..CW) = shift (0 beyond Cout) — read with broadcast 16-byte loads instead of 2 * CW global loads
// per thread and tile (the eval forward, where every convolution carries a folded BatchNorm, is epilogue-bound on them).
template <int CW, bool BW = true>
__device__ __forceinline__ void epilogue_chunk(float (&f)[CW], const EpiTC& e, long long m, bool row_ok, int cbase, int on,
                                               int oh, int ow, bool do_stats, float* sstat_sum, float* sstat_sq, int lane,
                                               const float* s_aff = nullptr, int aff_stride = 0) {
  if (s_aff) {
#pragma unroll
    for (int q = 0; q < CW / 4; ++q) {
      const float4 a = *reinterpret_cast<const float4*>(s_aff + 4 * q);
      const float4 b = *reinterpret_cast<const float4*>(s_aff + aff_stride + 4 * q);
      f[4 * q] = fmaf(f[4 * q], a.x, b.x); f[4 * q + 1] = fmaf(f[4 * q + 1], a.y, b.y);
      f[4 * q + 2] = fmaf(f[4 * q + 2], a.z, b.z); f[4 * q + 3] = fmaf(f[4 * q + 3], a.w, b.w);
    }
  } else if (e.scale) {
#pragma unroll
    for (int i = 0; i < CW; ++i)
      if (cbase + i < e.Cout) f[i] = fmaf(f[i], __ldg(e.scale + cbase + i), __ldg(e.shift + cbase + i));
  } else if (e.shift) {
#pragma unroll
    for (int i = 0; i < CW; ++i)
      if (cbase + i < e.Cout) f[i] += __ldg(e.shift + cbase + i);
  }
  if (e.res && row_ok) {
    const uint4* rp4 = reinterpret_cast<const uint4*>(e.res + m * e.ldr + cbase);
#pragma unroll
    for (int q = 0; q < CW / 8; ++q) {
      uint4 rr = __ldg(rp4 + q);
      const bf16* rb16 = reinterpret_cast<const bf16*>(&rr);
#pragma unroll
      for (int i = 0; i < 8; ++i) f[q * 8 + i] += __bfloat162float(rb16[i]);
    }
  }
  if (e.relu) {
#pragma unroll
    for (int i = 0; i < CW; ++i) f[i] = fmaxf(f[i], 0.f);
  }
  if (row_ok) {
    if (e.out_nchw) {
#pragma unroll
      for (int i = 0; i < CW; ++i)
        if (cbase + i < e.Cout) e.out_nchw[(((long long)on * e.Cout + cbase + i) * e.Ho + oh) * e.Wo + ow] = f[i];
    } else {
      uint4* op = reinterpret_cast<uint4*>(e.out + m * e.ldo + cbase);
#pragma unroll
      for (int q = 0; q < CW / 8; ++q) {
        uint4 o;
        __nv_bfloat162* o2 = reinterpret_cast<__nv_bfloat162*>(&o);
#pragma unroll
        for (int i = 0; i < 4; ++i) o2[i] = __floats2bfloat162_rn(f[q * 8 + 2 * i], f[q * 8 + 2 * i + 1]);
        op[q] = o;
      }
    }
  }
  if (do_stats) {
    float sq[CW];
    if (BW && e.bw_x) {
      bw_stat_terms<CW>(f, sq, e, m, row_ok, cbase);
    } else {
#pragma unroll
      for (int i = 0; i < CW; ++i) {
        if (!row_ok) f[i] = 0.f;
        sq[i] = f[i] * f[i];
      }
    }
    float cs, cq;
    if (CW == 32) { cs = warp_colsum32(f, lane); cq = warp_colsum32(sq, lane); }
    else { cs = warp_colsum16(f, lane); cq = warp_colsum16(sq, lane); }
    if (lane < CW) {           // one writer per (warp, column) slot: no shared-memory atomics, deterministic
      sstat_sum[lane] += cs;
      sstat_sq[lane] += cq;
    }
  }
}

// ---- launch-geometry state shared by the tensor-core translation units (defined in conv_tc.cu)
extern int g_num_sms;
extern int g_verbose;       // D3FK_VERBOSE=1: print launch geometry
extern int g_max_cluster;   // cap of the split-K / split-pixel cluster size (debug builds: D3FK_CLUSTER)
constexpr int SLAB_MAX_SMEM = 225 * 1024;

// Co-resident CTA capacity of a cluster launch, per CTAs-per-SM.  cudaOccupancyMaxActiveClusters reports one CTA per SM
// for kernels that allocate tensor memory; tools/probes/cluster_residency.cu measured, on B200 at 2 CTAs/SM, 296 CTAs for
// cluster sizes 1-2, 284 for 4 and 264 for 8 (GPC boundaries strand a few SMs) — the table below keeps a safety margin.
// GUARANTEED co-resident CTAs of a cluster launch (what cudaOccupancyMaxActiveClusters reports for these kernels: one CTA
// per SM, minus the SMs GPC boundaries strand) — the bound for anything that spins on a grid-wide barrier.  A 16 x 8-CTA
// cluster grid of the 320-thread kernel was observed NOT to be co-resident (15 clusters were), although the probe kernel
// reached 33: the optimistic table below is for wave sizing only.
static inline int cluster_capacity_safe(int cl) { return cl >= 8 ? 120 : cl >= 4 ? 132 : g_num_sms; }
static inline int cluster_capacity(int cl, int ctas_per_sm) {
  const int per_sm = cl >= 8 ? 120 : cl >= 4 ? 138 : g_num_sms;   // usable SMs (of 148) for this cluster size
  return per_sm * ctas_per_sm;
}

#define D3FK_SET_SMEM(k, bytes)                                                                     \
  if (e == cudaSuccess) e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);               \
  if (e == cudaSuccess) e = cudaFuncSetAttribute(k, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);

}  // namespace d3fk
