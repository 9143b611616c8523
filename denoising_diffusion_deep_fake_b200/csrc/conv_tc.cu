// conv_tc.cu — bf16 implicit-GEMM convolution, dgrad and wgrad on the Blackwell 5th-gen tensor
// cores: tcgen05.mma issued by one elected thread, accumulators in TMEM, operands staged in
// 128B-swizzled shared memory by an asynchronous gather (cp.async completing on mbarriers),
// tcgen05.ld epilogue fused with BN-statistics / folded-BN affine / residual / ReLU.
//
// Forward / dgrad:  D[M = B*Ho*Wo pixels, N = Cout] = A[M, K = taps*Cin] x W[N, K]^T     (both K-major)
// Weight gradient:  D[M = 128 k-rows, N = Cout]     = A[pixels, k]^T x dY[pixels, Cout]  (both MN-major)
//
// Warp roles (160 threads): warps 0-3 = gather producers, then epilogue (TMEM lane quarter = warp);
// warp 4 = TMEM allocator + single-thread MMA issuer.
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
// NOTE: no fence.proxy.async between the full-barrier wait and tcgen05.mma: cp.async completion is tracked by
// the mbarrier itself (ARRIVES.LDGSTSBAR), as in CUTLASS's sm100 cp.async mainloop; the fence lowers to
// MEMBAR.ALL.CTA, which drains every in-flight LDGSTS of the CTA and serialises the whole pipeline.
// bounded wait: a barrier that never completes sets the device error flag instead of hanging the GPU
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity, int* errflag) {
  for (uint32_t it = 0;; ++it) {
    if (mbar_try_wait(bar, parity)) return;
    if (it > (1u << 20)) {
      atomicExch(errflag, 1);
      return;
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
// Warp-uniform issue: the whole MMA warp runs the loop and only the leader lane's instruction takes effect (inside an
// `if (lane == 0)` region ptxas wraps every tcgen05.mma in an ELECT / R2UR serialisation loop).
__device__ __forceinline__ void umma_f16_lohi_p(uint32_t d_tmem, uint32_t alo, uint32_t ahi, uint32_t blo, uint32_t bhi, uint32_t idesc,
                                                uint32_t accumulate, uint32_t leader) {
  asm volatile(
      "{\n\t.reg .pred p, q;\n\t.reg .b64 da, db;\n\t"
      "mov.b64 da, {%1, %2};\n\t"
      "mov.b64 db, {%3, %4};\n\t"
      "setp.ne.b32 p, %6, 0;\n\t"
      "setp.ne.b32 q, %7, 0;\n\t"
      "@q tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n\t}"
      ::"r"(d_tmem), "r"(alo), "r"(ahi), "r"(blo), "r"(bhi), "r"(idesc), "r"(accumulate), "r"(leader)
      : "memory");
}
__device__ __forceinline__ void umma_commit_p(uint32_t bar, uint32_t leader) {
  asm volatile(
      "{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %1, 0;\n\t"
      "@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}" ::"r"(bar), "r"(leader)
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

template <int BN> struct ConvCfg {
  static constexpr int STAGES = BN >= 128 ? 3 : 4;
  static constexpr int B_STAGE_BYTES = BN * 128;
  static constexpr int SMEM = 1024 + STAGES * (A_STAGE_BYTES + B_STAGE_BYTES) + 256 + 8 * BN * 4;
  static constexpr int ACC_COLS = BN < 32 ? 32 : BN;   // one accumulator buffer
  static constexpr int TMEM_COLS = 2 * ACC_COLS;        // double buffered: epilogue of tile i overlaps MMAs of tile i+1
};

// Tile schedule.  KS == 1: tile t -> (m tile, n tile), consecutive CTAs walk consecutive m tiles, persistent loop.
// KS > 1 (split K over a thread-block cluster of KS CTAs): one tile per CTA, t -> (k split = cluster rank, m tile, n tile);
// the KS partial accumulators are reduce-scattered through distributed shared memory (no workspace, no second kernel).
struct TileSched {
  int MT, NT, KS, kb_per_split, nkb, total, a_ca;
  int bw, bh, bn, wt, ht;   // PATH 2: the 128-pixel M tile as a (w, h, n) box and the tile grid along w / h
};

__device__ __forceinline__ void decode_tile(const TileSched& ts, int t, int& mt, int& nt, int& ks) {
  if (ts.KS > 1) {
    ks = t % ts.KS;
    const int r = t / ts.KS;
    mt = r % ts.MT;
    nt = r / ts.MT;
  } else {
    ks = 0;
    mt = t % ts.MT;
    nt = t / ts.MT;
  }
}

// Epilogue of CW accumulator columns [cbase, cbase+CW) of output row m held in f[]: folded-BN affine / bias, residual,
// ReLU, store (bf16 NHWC or fp32 NCHW) and per-channel batch statistics.  The statistics are folded over the 32 rows of
// the warp with a transpose-reduce and ACCUMULATED into this warp's shared-memory slots (sstat_warp[col], [BN + col]);
// the CTA flushes the slots to global memory with one double atomic per channel when its n tile changes / at the end.
template <int CW, bool BW = true>
__device__ __forceinline__ void epilogue_chunk(float (&f)[CW], const EpiTC& e, long long m, bool row_ok, int cbase, int on,
                                               int oh, int ow, bool do_stats, float* sstat_sum, float* sstat_sq, int lane) {
  if (e.scale) {
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

// Fused train-mode BatchNorm (forward): when every output tile of the layer is resident at once (one tile per CTA, the
// accumulator parked in TMEM — or, for split-K clusters, the reduced slice in shared memory), the conv kernel itself
// finalises the batch statistics behind a grid-wide barrier and applies normalise + residual + ReLU from the fp32
// accumulator: raw conv output (kept for the backward pass) and activation are both written by this one kernel and the
// separate BN kernel (launch, prologue, a re-read of the raw tensor) disappears.
struct FuseBN {
  const float* gamma; const float* beta;
  float* mean; float* invstd; float* running_mean; float* running_var; long long* nbt;
  bf16* act; const bf16* res;
  unsigned* barrier;         // zeroed by the forward's statistics memset
  long long count;
  int ldact, ldr, relu;
  float eps, momentum;
};

// PATH 0 (LINEAR): one source, no upsample, forward gather or stride-1 transposed gather — the tap
//   offset is the same for every row, so a row costs two compares, one 64-bit add and the cp.async.
// PATH 1 (GENERIC): nearest-2x upsample + channel concat (decoder conv1) and stride-2 transposed gather.
// PATH 2 (TMA): stride-1 "same" convolutions with Cin % 64 == 0 and power-of-two extents: the A tile of tap (kh,kw)
//   is the NHWC box {64 channels, bw, bh, bn} shifted by the tap offset, loaded by ONE cp.async.bulk.tensor.4d with
//   hardware zero fill for the padding halo — no per-thread address arithmetic at all.
// The weight tile is always a TMA 2-D box {64, BN} of the packed [Cout][K] matrix (OOB rows / K tail zero filled).
template <int BN, int PATH, int FUSE>
__global__ void __launch_bounds__(TC_THREADS, 2) conv_tc_kernel(Gather g, FastDiv dWo, FastDiv dHo, const __grid_constant__ CUtensorMap tmA,
                                                             const __grid_constant__ CUtensorMap tmB, EpiTC e, TileSched ts,
                                                             FuseBN fb, int* errflag) {
  using Cfg = ConvCfg<BN>;
  constexpr int STAGES = Cfg::STAGES;
  constexpr int CW = BN >= 32 ? 32 : 16;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t a_base = base;
  const uint32_t b_base = base + STAGES * A_STAGE_BYTES;
  const uint32_t bar_base = b_base + STAGES * Cfg::B_STAGE_BYTES;  // full[S], empty[S], acc_full[2], acc_empty[2], tmem ptr
  uint8_t* gen_bar = smem_raw + (base - smem_u32(smem_raw)) + STAGES * (A_STAGE_BYTES + Cfg::B_STAGE_BYTES);
  volatile uint32_t* tmem_ptr_slot = reinterpret_cast<volatile uint32_t*>(gen_bar + 8 * (2 * STAGES + 4));
  float* s_stat = reinterpret_cast<float*>(gen_bar + 256);  // [4 warps][2][BN]
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (STAGES + s); };
  auto acc_full_bar = [&](int b) { return bar_base + 8u * (2 * STAGES + b); };
  auto acc_empty_bar = [&](int b) { return bar_base + 8u * (2 * STAGES + 2 + b); };

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  // Epilogue warps: 0-3 (also the gather producers of PATH 0/1) and the helpers 6-9.  A warp may touch TMEM lanes
  // 32*(warp%4)..+32 only, so helper warp w shares the lane quarter of primary warp w%4 and takes every other column
  // chunk: the epilogue is a dependent instruction chain per warp (~1 us per 32 columns with one warp per scheduler).
  const bool is_epi = warp < 4 || warp >= 6;
  const int q = warp & 3;                         // TMEM lane quarter / row group of this epilogue warp
  const int half = warp >= 6 ? 1 : 0;             // helpers take the odd column chunks
  const int etid = warp < 4 ? tid : 128 + (tid - 192);   // 0..255 over the 8 epilogue warps
  const bool clus = ts.KS > 1;
  const bool do_stats = e.stats != nullptr;
  TL_DECL
  TL_BEGIN

  if (tid == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(full_bar(s), PATH == 2 ? 1 : 129);   // 128 gather threads + the TMA thread's expect_tx arrive
      mbar_init(empty_bar(s), 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(acc_full_bar(b), 1);
      mbar_init(acc_empty_bar(b), EPI_THREADS);
    }
    fence_barrier_init();
  }
  if (is_epi) {
    for (int i = etid; i < 8 * BN; i += EPI_THREADS) s_stat[i] = 0.f;
  }
  if (warp == 5 && lane == 0) {
    tma_prefetch_desc(&tmB);
    if (PATH == 2) tma_prefetch_desc(&tmA);
    // The packed weights are not written by the preceding kernel (pack_all ran at the start of the step), so the first
    // pipeline stages' weight tiles are pulled into L2 now, while the previous kernel is still draining (PDL prologue).
    int mt0, nt0, ks0;
    decode_tile(ts, blockIdx.x, mt0, nt0, ks0);
    const int kbp = ks0 * ts.kb_per_split;
    for (int i = 0; i < STAGES && kbp + i < ts.nkb; ++i) tma_prefetch_l2_2d(&tmB, (kbp + i) * TC_BK, nt0 * BN);
  }
  if (FUSE == 2 && is_epi && e.bw_x) {
    // Fused BN-backward reduction: the epilogue reads one row segment of the forward's raw output / activation per lane.
    // They were written a whole forward ago (HBM, not L2) and not by the preceding kernel: pull the first tile's segments
    // into L2 now, so the loads on the epilogue's dependent chain are L2 hits.
    int mt0, nt0, ks0;
    decode_tile(ts, blockIdx.x, mt0, nt0, ks0);
    const long long mp = (long long)mt0 * TC_BM + q * 32 + lane;
    if (mp < g.M) {
      for (int b = half * 128; b < BN * 2; b += 256) {
        prefetch_l2(reinterpret_cast<const char*>(e.bw_x + mp * e.bw_ldx + nt0 * BN) + b);
        if (e.bw_relu) prefetch_l2(reinterpret_cast<const char*>(e.bw_act + mp * e.bw_ldact + nt0 * BN) + b);
      }
    }
  }
  if (warp == 4) tmem_alloc(smem_u32((const void*)tmem_ptr_slot), Cfg::TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_d = *tmem_ptr_slot;
  if (tid == 0) { TL_STAMP(1) }
  pdl_enter();   // prologue above overlaps the previous kernel; from here on its results are visible
  if (tid == 0) { TL_STAMP(2) }

  // flush this CTA's accumulated statistics of n tile `nt` (epilogue warps only)
  auto flush_stats = [&](int nt) {
    asm volatile("bar.sync 1, 256;" ::: "memory");
    const int n0 = nt * BN;
    if (tid < BN && n0 + tid < e.Cout) {
      const float a = (s_stat[tid] + s_stat[2 * BN + tid]) + (s_stat[4 * BN + tid] + s_stat[6 * BN + tid]);
      const float b = (s_stat[BN + tid] + s_stat[3 * BN + tid]) + (s_stat[5 * BN + tid] + s_stat[7 * BN + tid]);
      if (a != 0.f || b != 0.f) {
        atomicAdd(&e.stats[n0 + tid], (double)a);
        atomicAdd(&e.stats[e.Cout + n0 + tid], (FUSE == 2 && e.bw_x) ? bw_second_sum((double)a, (double)b, e, n0 + tid) : (double)b);
      }
    }
    asm volatile("bar.sync 1, 256;" ::: "memory");
    for (int i = etid; i < 8 * BN; i += EPI_THREADS) s_stat[i] = 0.f;
    asm volatile("bar.sync 1, 256;" ::: "memory");
  };

  // FUSE: after this CTA's statistics are flushed — grid barrier, then scale / shift of the tile's BN channels into s_stat
  // (free again after the flush): s_stat[c] = scale, s_stat[BN + c] = shift.  The CTA with m tile 0 (and k slice 0)
  // publishes mean / invstd / running statistics of its n tile.
  auto fuse_finalize = [&](int nt, bool publisher) {
    asm volatile("bar.sync 1, 256;" ::: "memory");
    if (tid == 0) grid_barrier_arrive_wait(fb.barrier, gridDim.x, errflag);
    asm volatile("bar.sync 1, 256;" ::: "memory");
    const int ch = nt * BN + tid;
    if (tid < BN && ch < e.Cout) {
      const double n = (double)fb.count;
      const double mean = __ldcg(e.stats + ch) / n;
      double var = __ldcg(e.stats + e.Cout + ch) / n - mean * mean;
      if (var < 0) var = 0;
      const double invstd = rsqrt(var + (double)fb.eps);
      const double gm = (double)__ldg(fb.gamma + ch), bt = (double)__ldg(fb.beta + ch);
      s_stat[tid] = (float)(gm * invstd);
      s_stat[BN + tid] = (float)(bt - mean * gm * invstd);
      if (publisher) {
        fb.mean[ch] = (float)mean;
        fb.invstd[ch] = (float)invstd;
        if (fb.running_mean) {
          const double unbiased = n > 1 ? var * n / (n - 1) : var;
          fb.running_mean[ch] = (float)((1.0 - fb.momentum) * fb.running_mean[ch] + fb.momentum * mean);
          fb.running_var[ch] = (float)((1.0 - fb.momentum) * fb.running_var[ch] + fb.momentum * unbiased);
        }
        if (ch == 0 && fb.nbt) *fb.nbt += 1;
      }
    }
    asm volatile("bar.sync 1, 256;" ::: "memory");
  };
  // FUSE: activation of CWF accumulator columns [ct, ct + CWF) (tile-relative) of output row m
  auto fuse_apply = [&](float* f, int ncol, long long m, bool row_ok, int nt, int ct) {
    if (!row_ok) return;
    const int cbase = nt * BN + ct;
    for (int q = 0; q < ncol / 8; ++q) {
      float y[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) y[i] = fmaf(f[q * 8 + i], s_stat[ct + q * 8 + i], s_stat[BN + ct + q * 8 + i]);
      if (fb.res) {
        const uint4 rr = __ldg(reinterpret_cast<const uint4*>(fb.res + m * fb.ldr + cbase + q * 8));
        const bf16* rb16 = reinterpret_cast<const bf16*>(&rr);
#pragma unroll
        for (int i = 0; i < 8; ++i) y[i] += __bfloat162float(rb16[i]);
      }
      if (fb.relu) {
#pragma unroll
        for (int i = 0; i < 8; ++i) y[i] = fmaxf(y[i], 0.f);
      }
      uint4 o;
      __nv_bfloat162* o2 = reinterpret_cast<__nv_bfloat162*>(&o);
#pragma unroll
      for (int i = 0; i < 4; ++i) o2[i] = __floats2bfloat162_rn(y[2 * i], y[2 * i + 1]);
      *reinterpret_cast<uint4*>(fb.act + m * fb.ldact + cbase + q * 8) = o;
    }
  };

  int my_mt = 0, my_nt = 0, my_ks = 0;   // cluster mode: the single tile of this CTA
  if (clus) decode_tile(ts, blockIdx.x, my_mt, my_nt, my_ks);

  if (warp == 5) {
    // ===================== TMA producer (one thread) =====================
    if (lane == 0) {
      uint32_t kbg = 0;
      const int sgn = g.mode ? -1 : 1;
      const int off = g.mode ? g.pad : -g.pad;
      for (int t = blockIdx.x; t < ts.total; t += gridDim.x) {
        int mt, nt, ks;
        decode_tile(ts, t, mt, nt, ks);
        const int kb0 = ks * ts.kb_per_split;
        const int kb1 = min(ts.nkb, kb0 + ts.kb_per_split);
        int w0 = 0, h0 = 0, i0 = 0, tap = 0, c = 0, khi = 0, kwi = 0;
        if (PATH == 2) {
          const int tw = mt % ts.wt;
          const int r2 = mt / ts.wt;
          w0 = tw * ts.bw + off;
          h0 = (r2 % ts.ht) * ts.bh + off;
          i0 = (r2 / ts.ht) * ts.bn;
          const int k = kb0 * TC_BK;
          tap = k / g.ctot;
          c = k - tap * g.ctot;
          khi = tap / g.kw;
          kwi = tap - khi * g.kw;
        }
        for (int kb = kb0; kb < kb1; ++kb, ++kbg) {
          const int s = kbg % STAGES;
          if (kbg >= STAGES) mbar_wait(empty_bar(s), ((kbg / STAGES) - 1) & 1, errflag);
          mbar_arrive_expect_tx(full_bar(s), Cfg::B_STAGE_BYTES + (PATH == 2 ? A_STAGE_BYTES : 0));
          tma_load_2d(b_base + s * Cfg::B_STAGE_BYTES, &tmB, kb * TC_BK, nt * BN, full_bar(s));
          if (PATH == 2) {
            tma_load_4d(a_base + s * A_STAGE_BYTES, &tmA, c, w0 + sgn * kwi, h0 + sgn * khi, i0, full_bar(s));
            c += TC_BK;
            if (c >= g.ctot) {
              c = 0;
              if (++kwi == g.kw) { kwi = 0; ++khi; }
            }
          }
        }
      }
    }
    __syncwarp();
  } else if (is_epi) {
    const int j = tid & 7;    // 16-byte chunk (8 channels) within the 128-byte k-row
    const int rb = tid >> 3;  // rows rb + 16*i
    const uint32_t sw = (uint32_t)((j ^ (rb & 7)) << 4);
    const int sgn = g.mode ? -1 : 1;
    const uint32_t smask = g.mode ? (uint32_t)(g.stride - 1) : 0u;   // transposed gather: coordinate must be a multiple
    const int sshift = g.mode ? g.sshift : 0;                        // of the stride (forward folds it into h0/w0)
    uint32_t kbg = 0;  // k-blocks issued by this CTA so far (pipeline stage / phase bookkeeping)
    uint32_t tile_iter = 0;
    int cur_nt = -1;
    for (int t = blockIdx.x; t < ts.total; t += gridDim.x, ++tile_iter) {
      int mt, nt, ks;
      decode_tile(ts, t, mt, nt, ks);
      const int m0 = mt * TC_BM, n0 = nt * BN;
      const int kb0 = ks * ts.kb_per_split;
      const int kb1 = min(ts.nkb, kb0 + ts.kb_per_split);

      if (PATH != 2 && warp < 4) {
        // ---- per-row state
        int rh[8], rw[8];
        int rn[8];                    // GENERIC: image index
        const bf16* rp[8];            // LINEAR: pointer of (n, h0, w0, channel 0) in src0
  #pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int m = m0 + rb + 16 * i;
          int n = 0, h0 = -(1 << 28), w0 = 0;
          if (m < g.M) {
            const uint32_t q = fdiv((uint32_t)m, dWo);
            const int wo = m - (int)q * g.Wo;
            n = (int)fdiv(q, dHo);
            const int ho = (int)q - n * g.Ho;
            if (g.mode == 0) { h0 = ho * g.stride - g.pad; w0 = wo * g.stride - g.pad; }
            else { h0 = ho + g.pad; w0 = wo + g.pad; }
          }
          rh[i] = h0; rw[i] = w0;
          if (PATH == 0) rp[i] = (const bf16*)g.src0 + ((long long)(n * g.Hi + h0) * g.Wi + w0) * g.ld0;
          else rn[i] = n;
        }
        // ---- k state of this thread's chunk at the first k-block of the split
        int k = kb0 * TC_BK + j * 8;
        int tap = k / g.ctot;
        int c = k - tap * g.ctot;
        int khi = tap / g.kw, kwi = tap - khi * g.kw;

        for (int kb = kb0; kb < kb1; ++kb, ++kbg) {
          const int s = kbg % STAGES;
          if (kbg >= STAGES) mbar_wait(empty_bar(s), ((kbg / STAGES) - 1) & 1, errflag);
          const bool k_ok = k < g.K;
          const uint32_t a_dst = a_base + s * A_STAGE_BYTES + rb * 128 + sw;
          const int dkh = sgn * khi, dkw = sgn * kwi;
          if (PATH == 0) {
            const long long koff = (long long)(dkh * g.Wi + dkw) * g.ld0 + c;
  #pragma unroll
            for (int i = 0; i < 8; ++i) {
              const bool ok = k_ok && (unsigned)(rh[i] + dkh) < (unsigned)g.Hi && (unsigned)(rw[i] + dkw) < (unsigned)g.Wi;
              const void* src = ok ? (const void*)(rp[i] + koff) : g.src0;
              if (ts.a_ca) cp_async_16_ca(a_dst + i * (16 * 128), src, ok ? 16u : 0u);
              else cp_async_16(a_dst + i * (16 * 128), src, ok ? 16u : 0u);
            }
          } else {
            const bool second = c >= g.c0;
            const bf16* sb = second ? (const bf16*)g.src1 + (c - g.c0) : (const bf16*)g.src0 + c;
            const int ld = second ? g.ld1 : g.ld0;
            const int up = second ? 0 : g.up0;
            const int hs = g.Hi >> up, wsz = g.Wi >> up;
  #pragma unroll
            for (int i = 0; i < 8; ++i) {
              const int th = rh[i] + dkh, tw = rw[i] + dkw;
              const int hi = th >> sshift, wi = tw >> sshift;
              const bool ok = k_ok && (((uint32_t)(th | tw)) & (0x80000000u | smask)) == 0 && hi < g.Hi && wi < g.Wi;
              const long long pix = (long long)((rn[i] * hs + (hi >> up)) * wsz + (wi >> up));
              const void* src = ok ? (const void*)(sb + pix * ld) : g.src0;
              if (ts.a_ca) cp_async_16_ca(a_dst + i * (16 * 128), src, ok ? 16u : 0u);
              else cp_async_16(a_dst + i * (16 * 128), src, ok ? 16u : 0u);
            }
          }
          cp_async_mbar_arrive(full_bar(s));
          mbar_arrive(full_bar(s));
          k += TC_BK;
          c += TC_BK;
          while (c >= g.ctot) {
            c -= g.ctot;
            if (++kwi == g.kw) { kwi = 0; ++khi; }
          }
        }
      }

      // ===================== epilogue: TMEM -> registers -> global =====================
      const uint32_t abuf = tile_iter & 1;
      mbar_wait(acc_full_bar(abuf), (tile_iter >> 1) & 1, errflag);
      if (tid == 0) { TL_STAMP(5) }
      tc_fence_after();
      if (clus) break;   // split K: the accumulator is reduced across the cluster below
      if (do_stats && cur_nt != nt) {
        if (cur_nt >= 0) flush_stats(cur_nt);
        cur_nt = nt;
      }
      const int row = q * 32 + lane;
      const int m = m0 + row;
      const bool row_ok = m < g.M;
      int on = 0, oh = 0, ow = 0;
      if (e.out_nchw && row_ok) {
        const uint32_t q = fdiv((uint32_t)m, dWo);
        ow = m - (int)q * e.Wo;
        on = (int)fdiv(q, dHo);
        oh = (int)q - on * e.Ho;
      }
#pragma unroll 1
      for (int cc = half * CW; cc < BN; cc += 2 * CW) {
        uint32_t raw[CW];
        const uint32_t taddr = tmem_d + ((uint32_t)(q * 32) << 16) + abuf * Cfg::ACC_COLS + (uint32_t)cc;
        if (CW == 32) tmem_ld32(taddr, raw); else tmem_ld16(taddr, raw);
        tmem_ld_wait();
        float f[CW];
#pragma unroll
        for (int i = 0; i < CW; ++i) f[i] = __uint_as_float(raw[i]);
        epilogue_chunk<CW, FUSE == 2>(f, e, (long long)m, row_ok, n0 + cc, on, oh, ow, do_stats, s_stat + q * 2 * BN + cc,
                           s_stat + q * 2 * BN + BN + cc, lane);
        if (tid == 0 && cc == 0) { TL_STAMP(8) }
      }
      if (tid == 0) { TL_STAMP(9) }
      if (FUSE == 1) {
        // one tile per CTA (the launcher guarantees it): statistics -> grid barrier -> activation from the parked accumulator
        flush_stats(nt);
        cur_nt = -1;
        fuse_finalize(nt, mt == 0);
#pragma unroll 1
        for (int cc = half * CW; cc < BN; cc += 2 * CW) {
          uint32_t raw[CW];
          const uint32_t taddr = tmem_d + ((uint32_t)(q * 32) << 16) + abuf * Cfg::ACC_COLS + (uint32_t)cc;
          if (CW == 32) tmem_ld32(taddr, raw); else tmem_ld16(taddr, raw);
          tmem_ld_wait();
          float f[CW];
#pragma unroll
          for (int i = 0; i < CW; ++i) f[i] = __uint_as_float(raw[i]);
          fuse_apply(f, CW, (long long)m, row_ok, nt, cc);
        }
      }
      tc_fence_before();   // this tile's TMEM reads are done: hand the accumulator buffer back to the MMA issuer
      mbar_arrive(acc_empty_bar(abuf));
    }
    if (tid == 0) { TL_STAMP(10) }
    if (!clus && do_stats && cur_nt >= 0) flush_stats(cur_nt);
  } else if (warp == 4) {
    // ===================== MMA issuer (one thread) =====================
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc(TC_BM, BN, 0, 0);
      uint32_t kbg = 0, tile_iter = 0;
      for (int t = blockIdx.x; t < ts.total; t += gridDim.x, ++tile_iter) {
        int mt, nt, ks;
        decode_tile(ts, t, mt, nt, ks);
        const int kb0 = ks * ts.kb_per_split;
        const int kb1 = min(ts.nkb, kb0 + ts.kb_per_split);
        const uint32_t abuf = tile_iter & 1;
        if (tile_iter >= 2) {   // the epilogue must have drained this accumulator buffer (two tiles ago)
          mbar_wait(acc_empty_bar(abuf), ((tile_iter >> 1) - 1) & 1, errflag);
          tc_fence_after();
        }
        const uint32_t d_addr = tmem_d + abuf * Cfg::ACC_COLS;
        for (int kb = kb0; kb < kb1; ++kb, ++kbg) {
          const int s = kbg % STAGES;
          mbar_wait(full_bar(s), (kbg / STAGES) & 1, errflag);
          if (kbg == 0) { TL_STAMP(3) }
          tc_fence_after();
          const uint32_t a_addr = a_base + s * A_STAGE_BYTES;
          const uint32_t b_addr = b_base + s * Cfg::B_STAGE_BYTES;
#pragma unroll
          for (int kk = 0; kk < TC_BK / 16; ++kk) {
            uint64_t ad = make_smem_desc(a_addr + kk * 32, 16, 1024);
            uint64_t bd = make_smem_desc(b_addr + kk * 32, 16, 1024);
            umma_f16(d_addr, ad, bd, idesc, (kb > kb0 || kk > 0) ? 1u : 0u);
          }
          umma_commit(empty_bar(s));
        }
        umma_commit(acc_full_bar(abuf));
        TL_STAMP(4)
      }
    }
    __syncwarp();
  }

  if (clus) {
    // ===================== split-K reduction across the cluster (KS CTAs, one k slice each) =====================
    // Every CTA holds a 128 x BN fp32 partial tile in TMEM.  Column slice j (SL = BN/KS columns) is owned by rank j:
    // each CTA writes its partial of slice j into slot [own rank] of rank j's receive buffer (the pipeline stages are
    // dead once every CTA has finished its main loop), then each owner sums KS slots and runs the epilogue on its slice.
    const int KS = ts.KS, SL = BN / KS, sl4 = SL >> 2;
    const uint32_t rank = cluster_ctarank();
    const uint32_t recv = a_base;   // [KS][SL/4][128 rows] float4
    const int row = q * 32 + lane;
    tc_fence_before();
    cluster_sync_all();             // #1: all main loops done (every epilogue warp saw acc_full)
    tc_fence_after();
    if (is_epi) {
#pragma unroll 1
      for (int cc = half * CW; cc < BN; cc += 2 * CW) {
        uint32_t raw[CW];
        const uint32_t taddr = tmem_d + ((uint32_t)(q * 32) << 16) + (uint32_t)cc;
        if (CW == 32) tmem_ld32(taddr, raw); else tmem_ld16(taddr, raw);
        tmem_ld_wait();
#pragma unroll
        for (int q = 0; q < CW / 4; ++q) {
          const int col = cc + 4 * q;
          const int owner = col / SL, within = col - owner * SL;
          const uint32_t la = recv + (uint32_t)((((int)rank * sl4 + (within >> 2)) * 128 + row) * 16);
          st_cluster_f4(mapa_shared(la, (uint32_t)owner), raw[4 * q], raw[4 * q + 1], raw[4 * q + 2], raw[4 * q + 3]);
        }
      }
    }
    cluster_sync_all();             // #2: all partials have landed
    if (is_epi) {
      const int m = my_mt * TC_BM + row;
      const bool row_ok = m < g.M;
      const int cslice = (int)rank * SL;     // first column of my slice within the tile
#pragma unroll 1
      for (int ch = half * 16; ch < SL; ch += 32) {
        float f[16];
#pragma unroll
        for (int c4 = 0; c4 < 4; ++c4) {
          float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
          for (int r = 0; r < KS; ++r) {
            const float4 v = ld_shared_f4(recv + (uint32_t)(((r * sl4 + (ch >> 2) + c4) * 128 + row) * 16));
            a.x += v.x; a.y += v.y; a.z += v.z; a.w += v.w;
          }
          f[4 * c4] = a.x; f[4 * c4 + 1] = a.y; f[4 * c4 + 2] = a.z; f[4 * c4 + 3] = a.w;
        }
        const int ct = cslice + ch;          // column within the tile
        epilogue_chunk<16, FUSE == 2>(f, e, (long long)m, row_ok, my_nt * BN + ct, 0, 0, 0, do_stats, s_stat + q * 2 * BN + ct,
                           s_stat + q * 2 * BN + BN + ct, lane);
      }
      if (do_stats) flush_stats(my_nt);
      if (FUSE == 1) {
        fuse_finalize(my_nt, my_mt == 0 && rank == 0);
#pragma unroll 1
        for (int ch = half * 16; ch < SL; ch += 32) {
          float f[16];
#pragma unroll
          for (int c4 = 0; c4 < 4; ++c4) {
            float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
            for (int r = 0; r < KS; ++r) {
              const float4 v = ld_shared_f4(recv + (uint32_t)(((r * sl4 + (ch >> 2) + c4) * 128 + row) * 16));
              a.x += v.x; a.y += v.y; a.z += v.z; a.w += v.w;
            }
            f[4 * c4] = a.x; f[4 * c4 + 1] = a.y; f[4 * c4 + 2] = a.z; f[4 * c4 + 3] = a.w;
          }
          fuse_apply(f, 16, (long long)m, row_ok, my_nt, cslice + ch);
        }
      }
    }
  }

  if (tid == 0) { TL_STAMP(6) }
  tc_fence_before();
  __syncthreads();
  if (warp == 4) tmem_dealloc(tmem_d, Cfg::TMEM_COLS);
  if (tid == 128) { TL_STAMP(7) }
}

static int g_num_sms = 148;
static int g_tile_loop = 1;   // D3FK_TILE_LOOP=0: one CTA per tile (debug aid)
static int g_a_ca = 0;        // D3FK_A_CA=1: L1-allocating activation gather
static int g_occ_cap = 0;     // D3FK_OCC=n: cap CTAs per SM
static int g_use_tma_a = 1;   // D3FK_TMA_A=0: force the gather producers (debug aid)
static int g_split_tiles = 74;  // split K only when the output tiles fill at most this many SMs
static int g_verbose = 0;      // D3FK_VERBOSE=1: print launch geometry
static int g_max_cluster = 8;   // D3FK_CLUSTER=n: cap the split-K cluster size (1 disables split K)

// Co-resident CTA capacity of a cluster launch, per CTAs-per-SM.  cudaOccupancyMaxActiveClusters reports one CTA per SM
// for kernels that allocate tensor memory; tools/probes/cluster_residency.cu measured, on B200 at 2 CTAs/SM, 296 CTAs for
// cluster sizes 1-2, 284 for 4 and 264 for 8 (GPC boundaries strand a few SMs) — the table below keeps a safety margin.
// GUARANTEED co-resident CTAs of a cluster launch (what cudaOccupancyMaxActiveClusters reports for these kernels: one CTA
// per SM, minus the SMs GPC boundaries strand) — the bound for anything that spins on a grid-wide barrier.  A 16 x 8-CTA
// cluster grid of the 320-thread kernel was observed NOT to be co-resident (15 clusters were), although the probe kernel
// reached 33: the optimistic table below is for wave sizing only.
static int cluster_capacity_safe(int cl) { return cl >= 8 ? 120 : cl >= 4 ? 132 : g_num_sms; }

static int cluster_capacity(int cl, int ctas_per_sm) {
  const int per_sm = cl >= 8 ? 120 : cl >= 4 ? 138 : g_num_sms;   // usable SMs (of 148) for this cluster size
  return per_sm * ctas_per_sm;
}

// Fusion request of launch_conv_bn (below): set around a launch_conv_tc call; the launcher takes it when the layer
// qualifies (BN == 128 tiles, one tile per co-resident CTA) and reports back through `taken`.
struct FuseReq { const d3fk_bn_params* bn; unsigned* barrier; bool taken; };
static thread_local FuseReq* t_fuse = nullptr;
static int g_fuse_bn = 1;   // D3FK_FUSE_BN=0: always run BatchNorm as its own kernel

template <int BN, int PATH>
static int launch_conv_tc_bn(const Gather& g, const d3fk_conv_params* p, cudaStream_t s, const TileSched& box) {
  EpiTC e{(bf16*)p->out, p->out_nchw, p->scale, p->shift, (const bf16*)p->res, p->stats, p->ldo, p->ldr, p->relu, p->Cout, p->Ho, p->Wo,
           (const bf16*)p->bw_x, (const bf16*)p->bw_act, p->bw_mean, p->bw_invstd, p->bw_ldx, p->bw_ldact, p->bw_relu};
  TileSched ts = box;
  ts.MT = cdiv(g.M, TC_BM);
  ts.NT = cdiv(p->Cout, BN);
  ts.nkb = cdiv(g.K, TC_BK);
  ts.KS = 1;
  const int tiles = ts.MT * ts.NT;
  ts.a_ca = g_a_ca && p->kh > 1;
  // split K over a cluster when the output tiles cannot fill the chip and the reduction is long
  if (BN == 128 && !p->out_nchw && p->Cout % BN == 0 && tiles <= g_split_tiles && ts.nkb >= 8 && g_max_cluster > 1) {
    int ks = 1;
    while (ks * 2 <= g_max_cluster && ks * 2 <= 8 && tiles * ks * 2 <= 2 * g_num_sms && ts.nkb / (ks * 2) >= 4) ks *= 2;
    while (ks > 1 && tiles * ks > cluster_capacity(ks, 2)) ks >>= 1;
    ts.KS = ks;
  }
  ts.kb_per_split = cdiv(ts.nkb, ts.KS);
  if (ts.KS > 1 && (ts.KS - 1) * ts.kb_per_split >= ts.nkb) {   // every rank needs at least one k-block
    ts.KS = 1;
    ts.kb_per_split = ts.nkb;
  }
  ts.total = tiles * ts.KS;
  int occ = (227 * 1024) / (ConvCfg<BN>::SMEM + 1024);
  if (occ * ConvCfg<BN>::TMEM_COLS > 512) occ = 512 / ConvCfg<BN>::TMEM_COLS;
  if (g_occ_cap > 0 && occ > g_occ_cap) occ = g_occ_cap;
  int grid = ts.total < g_num_sms * occ ? ts.total : g_num_sms * occ;
  if (!g_tile_loop || ts.KS > 1) grid = ts.total;

  // TMA descriptors: weights [Cout][K] as a {64, BN} box; PATH 2 also the activation tensor as a 4-D NHWC box
  alignas(64) CUtensorMap tmA, tmB;
  memset(&tmA, 0, sizeof(tmA));
  {
    uint64_t dims[2] = {(uint64_t)g.K, (uint64_t)p->Cout};
    uint64_t strides[1] = {(uint64_t)g.K * 2};
    uint32_t bx[2] = {TC_BK, (uint32_t)BN};
    int rc = get_tensor_map(&tmB, p->w, 2, dims, strides, bx, 128);
    if (rc) return rc;
  }
  if (PATH == 2) {
    uint64_t dims[4] = {(uint64_t)g.c0, (uint64_t)g.Wi, (uint64_t)g.Hi, (uint64_t)g.B};
    uint64_t strides[3] = {(uint64_t)g.ld0 * 2, (uint64_t)g.Wi * g.ld0 * 2, (uint64_t)g.Hi * g.Wi * g.ld0 * 2};
    uint32_t bx[4] = {TC_BK, (uint32_t)ts.bw, (uint32_t)ts.bh, (uint32_t)ts.bn};
    int rc = get_tensor_map(&tmA, p->src0, 4, dims, strides, bx, 128);
    if (rc) return rc;
  }
  // fused BatchNorm: every tile resident at once (grid == tiles * KS <= co-resident capacity), 128-wide tiles only
  FuseBN fb;
  memset(&fb, 0, sizeof(fb));
  bool fuse = false;
  if (BN == 128 && t_fuse && g_fuse_bn && p->stats && !p->scale && !p->shift && !p->res && !p->relu && !p->out_nchw && p->mode == 0 &&
      p->Cout % BN == 0 && ts.total <= cluster_capacity_safe(ts.KS)) {
    const d3fk_bn_params* b = t_fuse->bn;
    fb.gamma = b->gamma; fb.beta = b->beta; fb.mean = b->mean; fb.invstd = b->invstd;
    fb.running_mean = b->running_mean; fb.running_var = b->running_var; fb.nbt = (long long*)b->num_batches_tracked;
    fb.act = (bf16*)b->y; fb.res = (const bf16*)b->res; fb.barrier = t_fuse->barrier; fb.count = b->count;
    fb.ldact = b->ldy; fb.ldr = b->ldr; fb.relu = b->relu; fb.eps = b->eps; fb.momentum = b->momentum;
    fuse = true;
    grid = ts.total;
    t_fuse->taken = true;
  }
  if (g_verbose) fprintf(stderr, "[d3fk] conv<%d,%d> mode=%d M=%d K=%d Cout=%d tiles=%d KS=%d kbps=%d grid=%d fuse=%d\n", BN, PATH, g.mode, g.M, g.K, p->Cout, tiles, ts.KS, ts.kb_per_split, grid, (int)fuse);
  cudaError_t le;
  if (p->bw_x)
    le = launch_k(conv_tc_kernel<BN, PATH, 2>, dim3(grid), dim3(TC_THREADS), ConvCfg<BN>::SMEM, s, dim3(ts.KS, 1, 1), g,
                  make_fastdiv((uint32_t)g.Wo), make_fastdiv((uint32_t)g.Ho), tmA, tmB, e, ts, fb, g_dev_error_flag);
  else if (BN == 128 && fuse)
    le = launch_k(conv_tc_kernel<BN, PATH, (BN == 128 ? 1 : 0)>, dim3(grid), dim3(TC_THREADS), ConvCfg<BN>::SMEM, s, dim3(ts.KS, 1, 1), g,
                  make_fastdiv((uint32_t)g.Wo), make_fastdiv((uint32_t)g.Ho), tmA, tmB, e, ts, fb, g_dev_error_flag);
  else
    le = launch_k(conv_tc_kernel<BN, PATH, 0>, dim3(grid), dim3(TC_THREADS), ConvCfg<BN>::SMEM, s, dim3(ts.KS, 1, 1), g,
                  make_fastdiv((uint32_t)g.Wo), make_fastdiv((uint32_t)g.Ho), tmA, tmB, e, ts, fb, g_dev_error_flag);
  if (le != cudaSuccess) return set_error(D3FK_ERR_CUDA, "conv_tc launch: %s", cudaGetErrorString(le));
  count_launch();
  return check_launch("conv_tc");
}

template <int PATH>
static int launch_conv_tc_path(const Gather& g, const d3fk_conv_params* p, cudaStream_t s, const TileSched& box) {
  const int C = p->Cout;
  if (C <= 16) return launch_conv_tc_bn<16, PATH>(g, p, s, box);
  D3FK_CHECK_ARG(p->out_nchw == nullptr, "out_nchw only for Cout <= 16");
  if (C % 128 == 0) return launch_conv_tc_bn<128, PATH>(g, p, s, box);
  if (C % 64 == 0) return launch_conv_tc_bn<64, PATH>(g, p, s, box);
  if (C % 32 == 0) return launch_conv_tc_bn<32, PATH>(g, p, s, box);
  return set_error(D3FK_ERR_UNSUPPORTED, "conv_tc: Cout=%d (need <=16 or a multiple of 32)", C);
}

// PATH 2 eligibility: the 128-pixel M tile must be an axis-aligned (w, h, n) box whose pixel order equals the linear
// output-pixel order: bw = min(W,128) columns, then bh rows, then bn images, each level fully covered before the next.
static bool tma_box(const Gather& g, const d3fk_conv_params* p, TileSched& ts) {
  if (!g_use_tma_a) return false;
  if (p->c1 != 0 || p->up0 != 0 || p->stride != 1 || g.ctot % TC_BK != 0) return false;
  if (p->Ho != p->Hi || p->Wo != p->Wi || 2 * p->pad != p->kh - 1 || p->kh != p->kw) return false;
  if (((uintptr_t)p->src0 & 15) || (g.ld0 % 8)) return false;
  const int W = p->Wi, H = p->Hi;
  int bw = W < TC_BM ? W : TC_BM;
  if (TC_BM % bw || W % bw) return false;
  int bh = TC_BM / bw;
  if (bh > H) bh = H;
  if ((TC_BM / bw) % bh || H % bh) return false;
  int bn = TC_BM / (bw * bh);
  if (bw < W && bh != 1) return false;
  if (bh < H && bn != 1) return false;
  if (bn > 256) return false;
  ts.bw = bw; ts.bh = bh; ts.bn = bn; ts.wt = W / bw; ts.ht = H / bh;
  return true;
}

// ------------------------------------------------------------------------------------------
// Slab convolution: 3x3 / stride 1 / pad 1, one source, Cin in {16, 32, 64}, W in {16, 32, 64} (forward and dgrad).
// The implicit-GEMM kernels above re-read every activation pixel once per tap (9x) from L2.  Here a persistent CTA owns
// super-tiles of S*R full image rows (S sub-tiles of R*W = 128 pixels).  For each super-tile ONE TMA box per horizontal
// tap offset (3 boxes: columns shifted by -1/0/+1, S*R+2 rows, hardware zero fill for the halo) lands in shared memory;
// all 9 taps of all S sub-tiles are then UMMA operands that differ only by a row offset into those slabs, and the 9 weight
// tiles stay resident in shared memory for the whole kernel.  L2 -> SM traffic per pixel drops from 9x to 3*(S*R+2)/(S*R)
// and no thread computes an address: warp 5 issues 3 TMA loads per super-tile, warp 4 issues the MMAs, warps 0-3 only run
// the epilogue (double-buffered TMEM accumulators), so load, MMA and epilogue of consecutive super-tiles overlap.
struct SlabSched {
  int W, H, R, S;         // image extent; rows per 128-pixel sub-tile; sub-tiles per super-tile
  int Wt, wtiles;         // tile width min(W, 128) and tiles across the image width
  int row_bytes;          // Cin * 2 = bytes of one pixel row of the K-major operand = TMA / UMMA swizzle span (32/64/128)
  int slab_bytes;         // (S*R + 2) * W * row_bytes rounded up to 1 KB
  int slab_tx;            // bytes one slab load delivers
  int stages;             // slab pipeline depth
  int w_tile_bytes;       // BN * row_bytes: one tap's resident weight tile
  int total, tiles_per_img;
  int sgn;                // +1 forward taps, -1 transposed (dgrad)
  int ksteps;             // channels per chunk / 16
  int chunks, ctot;       // 64-channel chunks per tap when Cin > 64 (each chunk is one pipeline stage); total Cin
  uint32_t layout;        // UMMA smem-descriptor layout type: 2 = SWIZZLE_128B, 4 = 64B, 6 = 32B
  int M;                  // B*H*W
};

__device__ __forceinline__ uint64_t make_smem_desc_sw(uint32_t saddr, uint32_t sbo_bytes, uint32_t layout) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
  d |= (uint64_t)1 << 16;                              // leading byte offset (unused for swizzled K-major): 16 B
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;   // stride between 8-row groups
  d |= (uint64_t)1 << 46;                              // descriptor version (Blackwell)
  d |= (uint64_t)layout << 61;
  return d;
}

// Warps: 0-3 epilogue, 4 TMA producer, 5..5+NW-1 MMA issuers.  A single thread sustains only one of these small MMAs
// per ~90 cycles (descriptor moves into uniform registers + the issue itself are a dependent chain), so the 9*KSTEPS*S
// MMAs of a super-tile are spread over NW warps: warp w owns sub-tile w % S and every P-th tap (P = NW / S) in a private
// accumulator; the epilogue adds the P partial accumulators of a sub-tile.
template <int BN> struct SlabCfg {
  static constexpr int NW = BN >= 128 ? 2 : 4;
  static constexpr int THREADS = (5 + NW) * 32;
  static constexpr int ACC = BN < 32 ? 32 : BN;
  static constexpr int TMEM_COLS = 2 * NW * ACC;       // double buffered
};
template <int BN, int KSTEPS>
__global__ void __launch_bounds__(SlabCfg<BN>::THREADS) conv_slab_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                                                               EpiTC e, SlabSched ss, FastDiv dWo, FastDiv dHo, int* errflag) {
  constexpr int CW = BN >= 32 ? 32 : 16;
  constexpr int ACC = SlabCfg<BN>::ACC;
  constexpr int NW = SlabCfg<BN>::NW;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t w_base = base;                                         // [9 taps][chunks][BN rows][row_bytes]
  const uint32_t w_bytes = (uint32_t)((9 * ss.chunks * ss.w_tile_bytes + 1023) & ~1023);
  const uint32_t slab_base = base + w_bytes;                            // [stages][3][slab_bytes]
  const uint32_t stage_bytes = 3u * ss.slab_bytes;
  const uint32_t bar_base = slab_base + ss.stages * stage_bytes;        // full[4], empty[4], acc_full[2], acc_empty[2], wbar
  uint8_t* gen_bar = smem_raw + (bar_base - smem_u32(smem_raw));
  volatile uint32_t* tmem_ptr_slot = reinterpret_cast<volatile uint32_t*>(gen_bar + 8 * 13);
  float* s_stat = reinterpret_cast<float*>(gen_bar + 128);              // [4 warps][2][BN]
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (4 + s); };
  auto acc_full_bar = [&](int b) { return bar_base + 8u * (8 + b); };
  auto acc_empty_bar = [&](int b) { return bar_base + 8u * (10 + b); };
  const uint32_t wbar = bar_base + 8u * 12;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const bool do_stats = e.stats != nullptr;
  const int S = ss.S;
  const int P = NW / S;                                 // partial accumulators (tap subsets) per sub-tile
  constexpr uint32_t tmem_cols = (uint32_t)SlabCfg<BN>::TMEM_COLS;

  if (tid == 0) {
    for (int s = 0; s < 4; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), NW);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(acc_full_bar(b), NW);
      mbar_init(acc_empty_bar(b), 128);
    }
    mbar_init(wbar, 1);
    fence_barrier_init();
  }
  if (tid < 128) {
    for (int i = tid; i < 8 * BN; i += 128) s_stat[i] = 0.f;
  }
  if (warp == 4 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    const int cchp = ss.row_bytes >> 1;   // weights are not produced by the previous kernel: warm L2 during the PDL prologue
    for (int t = 0; t < 9; ++t)
      for (int c = 0; c < ss.chunks; ++c) tma_prefetch_l2_2d(&tmB, t * ss.ctot + c * cchp, 0);
  }
  if (warp == 5) tmem_alloc(smem_u32((const void*)tmem_ptr_slot), tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_d = *tmem_ptr_slot;
  pdl_enter();

  if (warp == 4) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      // resident weights: 9 * chunks boxes {chunk channels, BN} of the packed [Cout][9*Cin] matrix (rows >= Cout zero filled)
      mbar_arrive_expect_tx(wbar, 9u * ss.chunks * ss.w_tile_bytes);
      const int cch = ss.row_bytes >> 1;
      for (int t = 0; t < 9; ++t)
        for (int c = 0; c < ss.chunks; ++c)
          tma_load_2d(w_base + (t * ss.chunks + c) * ss.w_tile_bytes, &tmB, t * ss.ctot + c * cch, 0, wbar);
      const int rows = S * ss.R;
      // pipeline position over (super-tile, chunk) pairs as running counters: stage index, parity of the round, and whether
      // the ring has wrapped (a run-time `it % stages` / `it / stages` is a ~20-instruction sequence in every role's loop)
      int st = 0;
      uint32_t round_par = 0;
      bool wrapped = false;
      for (int t = blockIdx.x; t < ss.total; t += gridDim.x) {
        const int n = t / ss.tiles_per_img;
        const int rem = t - n * ss.tiles_per_img;
        const int hb = rem / ss.wtiles;
        const int h0 = hb * rows, w0 = (rem - hb * ss.wtiles) * ss.Wt;
        for (int c = 0; c < ss.chunks; ++c) {
          if (wrapped) mbar_wait(empty_bar(st), round_par ^ 1u, errflag);
          mbar_arrive_expect_tx(full_bar(st), 3u * ss.slab_tx);
          for (int sx = 0; sx < 3; ++sx)
            tma_load_4d(slab_base + st * stage_bytes + sx * ss.slab_bytes, &tmA, c * cch, w0 + sx - 1, h0 - 1, n, full_bar(st));
          if (++st == ss.stages) { st = 0; round_par ^= 1u; wrapped = true; }
        }
      }
    }
    __syncwarp();
  } else if (warp >= 5) {
    // ===================== MMA issuers =====================
    // The loop is kept to two adds per MMA: every descriptor is (constant high word, low word = constant | address >> 4)
    // and this warp's tap offsets live in registers.  Warp-uniform loop, leader-predicated issue.
    {
      const int w = warp - 5;
      const int my_s = w % S, my_p = w / S;            // sub-tile and tap subset of this warp
      const uint32_t leader = lane == 0 ? 1u : 0u;
      constexpr uint32_t idesc = make_idesc(TC_BM, BN, 0, 0);
      const uint32_t sbo = 8u * ss.row_bytes;
      const uint64_t dtempl = make_smem_desc_sw(0, sbo, ss.layout);
      const uint32_t dhi = (uint32_t)(dtempl >> 32), dlo = (uint32_t)dtempl;
      const uint32_t img_row16 = (uint32_t)(ss.Wt * ss.row_bytes) >> 4;  // one image row of the slab, in 16-byte units
      uint32_t a_off[9], b_lo[9];
#pragma unroll
      for (int tap = 0; tap < 9; ++tap) {
        const int kh = tap / 3, kw = tap - kh * 3;
        const int sy = ss.sgn > 0 ? kh : 2 - kh;
        const int sx = ss.sgn > 0 ? kw : 2 - kw;
        a_off[tap] = ((uint32_t)(sx * ss.slab_bytes) >> 4) + (uint32_t)(sy + my_s * ss.R) * img_row16;
        b_lo[tap] = dlo | ((w_base + (uint32_t)(tap * ss.chunks * ss.w_tile_bytes)) >> 4);
      }
      const uint32_t wchunk16 = (uint32_t)ss.w_tile_bytes >> 4;
      const bool active = my_p < P;                     // S * P == NW: always true; kept for clarity
      mbar_wait(wbar, 0, errflag);
      uint32_t tile_it = 0;
      int st = 0;
      uint32_t round_par = 0;
      const int pmask = P - 1;                          // P is 1, 2 or 4: tap % P == tap & (P - 1)
      for (int t = blockIdx.x; t < ss.total; t += gridDim.x, ++tile_it) {
        const uint32_t abuf = tile_it & 1;
        if (tile_it >= 2) mbar_wait(acc_empty_bar(abuf), ((tile_it >> 1) - 1) & 1, errflag);
        const uint32_t d_addr = tmem_d + abuf * (uint32_t)(NW * ACC) + (uint32_t)((my_s * P + my_p) * ACC);
        uint32_t first = 0u;
        for (int c = 0; c < ss.chunks; ++c) {
          mbar_wait(full_bar(st), round_par, errflag);
          tc_fence_after();
          const uint32_t sub_lo = dlo | ((slab_base + st * stage_bytes) >> 4);
          const uint32_t bc = (uint32_t)c * wchunk16;
          if (active) {
#pragma unroll
            for (int tap = 0; tap < 9; ++tap) {
              if ((tap & pmask) == my_p) {              // P in {1, 2, 4}
#pragma unroll
                for (int kk = 0; kk < KSTEPS; ++kk) {
                  umma_f16_lohi_p(d_addr, sub_lo + a_off[tap] + 2u * kk, dhi, b_lo[tap] + bc + 2u * kk, dhi, idesc, first, leader);
                  first = 1u;
                }
              }
            }
          }
          umma_commit_p(empty_bar(st), leader);
          if (++st == ss.stages) { st = 0; round_par ^= 1u; }
        }
        umma_commit_p(acc_full_bar(abuf), leader);
      }
    }
    __syncwarp();
  } else {
    // ===================== epilogue warps =====================
    // Narrow layers (BN <= 32) keep their batch statistics in registers: a thread owns row (warp*32+lane) of every
    // sub-tile, so it accumulates its own per-column sums over the whole kernel and the 128 rows are folded ONCE at
    // the end (instead of a shuffle transpose-reduce per tile).
    constexpr bool REG_STATS = BN <= 32;
    float rs[REG_STATS ? BN : 1], rq[REG_STATS ? BN : 1];
#pragma unroll
    for (int i = 0; i < (REG_STATS ? BN : 1); ++i) { rs[i] = 0.f; rq[i] = 0.f; }
    uint32_t it = 0;
    const int row = warp * 32 + lane;
    for (int t = blockIdx.x; t < ss.total; t += gridDim.x, ++it) {
      const uint32_t abuf = it & 1;
      const int n = t / ss.tiles_per_img;
      const int rem = t - n * ss.tiles_per_img;
      const int hb = rem / ss.wtiles;
      const int h0 = hb * S * ss.R, w0 = (rem - hb * ss.wtiles) * ss.Wt;
      mbar_wait(acc_full_bar(abuf), (it >> 1) & 1, errflag);
      tc_fence_after();
      for (int s = 0; s < S; ++s) {
        const long long m = ((long long)n * ss.H + h0 + s * ss.R) * ss.W + w0 + row;   // 128 consecutive pixels
        const bool row_ok = m < ss.M;
        int on = 0, oh = 0, ow = 0;
        if (e.out_nchw && row_ok) {
          const uint32_t q = fdiv((uint32_t)m, dWo);
          ow = (int)m - (int)q * e.Wo;
          on = (int)fdiv(q, dHo);
          oh = (int)q - on * e.Ho;
        }
#pragma unroll
        for (int cc = 0; cc < BN; cc += CW) {
          uint32_t raw[CW];
          const uint32_t taddr = tmem_d + ((uint32_t)(warp * 32) << 16) + abuf * (uint32_t)(NW * ACC) + (uint32_t)(s * P * ACC + cc);
          if (CW == 32) tmem_ld32(taddr, raw); else tmem_ld16(taddr, raw);
          tmem_ld_wait();
          float f[CW];
#pragma unroll
          for (int i = 0; i < CW; ++i) f[i] = __uint_as_float(raw[i]);
          for (int pp = 1; pp < P; ++pp) {              // add the other tap subsets' partial accumulators
            if (CW == 32) tmem_ld32(taddr + (uint32_t)(pp * ACC), raw); else tmem_ld16(taddr + (uint32_t)(pp * ACC), raw);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < CW; ++i) f[i] += __uint_as_float(raw[i]);
          }
          epilogue_chunk<CW>(f, e, m, row_ok, cc, on, oh, ow, do_stats && !REG_STATS, s_stat + warp * 2 * BN + cc,
                             s_stat + warp * 2 * BN + BN + cc, lane);
          if (REG_STATS && do_stats) {
            if (e.bw_x) {
              float sq[CW];
              bw_stat_terms<CW>(f, sq, e, m, row_ok, cc);
#pragma unroll
              for (int i = 0; i < CW; ++i) { rs[cc + i] += f[i]; rq[cc + i] += sq[i]; }
            } else {
#pragma unroll
              for (int i = 0; i < CW; ++i) { rs[cc + i] += f[i]; rq[cc + i] = fmaf(f[i], f[i], rq[cc + i]); }
            }
          }
        }
      }
      tc_fence_before();
      mbar_arrive(acc_empty_bar(abuf));
    }
    if (do_stats) {
      if (REG_STATS) {
#pragma unroll
        for (int cc = 0; cc < BN; cc += CW) {
          float a[CW], b[CW];
#pragma unroll
          for (int i = 0; i < CW; ++i) { a[i] = rs[cc + i]; b[i] = rq[cc + i]; }
          float cs, cq;
          if (CW == 32) { cs = warp_colsum32(a, lane); cq = warp_colsum32(b, lane); }
          else { cs = warp_colsum16(a, lane); cq = warp_colsum16(b, lane); }
          if (lane < CW) {
            s_stat[warp * 2 * BN + cc + lane] = cs;
            s_stat[warp * 2 * BN + BN + cc + lane] = cq;
          }
        }
      }
      asm volatile("bar.sync 1, 128;" ::: "memory");
      if (tid < BN && tid < e.Cout) {
        const float a = (s_stat[tid] + s_stat[2 * BN + tid]) + (s_stat[4 * BN + tid] + s_stat[6 * BN + tid]);
        const float b = (s_stat[BN + tid] + s_stat[3 * BN + tid]) + (s_stat[5 * BN + tid] + s_stat[7 * BN + tid]);
        atomicAdd(&e.stats[tid], (double)a);
        atomicAdd(&e.stats[e.Cout + tid], e.bw_x ? bw_second_sum((double)a, (double)b, e, tid) : (double)b);
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 5) tmem_dealloc(tmem_d, tmem_cols);
}

constexpr int SLAB_MAX_SMEM = 225 * 1024;
static int g_use_slab = 1;   // D3FK_SLAB=0: never take the slab path

// Slab-path eligibility and geometry.  Returns false when the generic kernels must be used.
static bool slab_plan(const Gather& g, const d3fk_conv_params* p, int BN, SlabSched& ss, int& smem) {
  if (!g_use_slab) return false;
  if (p->kh != 3 || p->kw != 3 || p->stride != 1 || p->pad != 1 || p->c1 != 0 || p->up0 != 0) return false;
  if (p->Ho != p->Hi || p->Wo != p->Wi) return false;
  const int C = g.ctot, W = p->Wi, H = p->Hi;
  if (C != 16 && C != 32 && C != 64 && C != 128) return false;   // 128 = two 64-channel chunks (one pipeline stage each)
  if (W != 16 && W != 32 && W != 64 && (W % 128)) return false;
  if (p->Cout > BN || (!p->out_nchw && p->Cout != BN)) return false;
  if (((uintptr_t)p->src0 & 15) || (g.ld0 % 8)) return false;
  const int Wt = W < TC_BM ? W : TC_BM;
  const int R = TC_BM / Wt;
  if (H % R) return false;
  ss.W = W; ss.H = H; ss.R = R; ss.Wt = Wt; ss.wtiles = W / Wt;
  const int cch = C > 64 ? 64 : C;
  ss.chunks = C / cch;
  ss.ctot = C;
  ss.row_bytes = cch * 2;
  ss.layout = cch == 64 ? 2u : cch == 32 ? 4u : 6u;
  ss.ksteps = cch / 16;
  ss.w_tile_bytes = BN * ss.row_bytes;
  ss.sgn = p->mode ? -1 : 1;
  ss.M = g.M;
  const int w_bytes = (9 * ss.chunks * ss.w_tile_bytes + 1023) & ~1023;
  // largest super-tile whose double-buffered accumulators fit TMEM and whose 2-stage slabs fit shared memory
  const int NW = BN >= 128 ? 2 : 4;                    // MMA warps (SlabCfg<BN>::NW): S must divide it
  for (int S = NW; S >= 1; S >>= 1) {
    if (H % (S * R)) continue;
    const int slab = ((S * R + 2) * Wt * ss.row_bytes + 1023) & ~1023;
    for (int stages = 3; stages >= 2; --stages) {
      const int need = 1024 + w_bytes + stages * 3 * slab + 128 + 8 * BN * 4;
      if (need > SLAB_MAX_SMEM) continue;
      ss.S = S;
      ss.slab_bytes = slab;
      ss.slab_tx = (S * R + 2) * Wt * ss.row_bytes;
      ss.stages = stages;
      ss.tiles_per_img = (H / (S * R)) * ss.wtiles;
      ss.total = p->B * ss.tiles_per_img;
      smem = need;
      return true;
    }
  }
  return false;
}

template <int BN, int KSTEPS>
static int launch_conv_slab_bn(const Gather& g, const d3fk_conv_params* p, cudaStream_t s, const SlabSched& ss, int smem) {
  EpiTC e{(bf16*)p->out, p->out_nchw, p->scale, p->shift, (const bf16*)p->res, p->stats, p->ldo, p->ldr, p->relu, p->Cout, p->Ho, p->Wo,
           (const bf16*)p->bw_x, (const bf16*)p->bw_act, p->bw_mean, p->bw_invstd, p->bw_ldx, p->bw_ldact, p->bw_relu};
  alignas(64) CUtensorMap tmA, tmB;
  const int C = g.ctot;
  {
    uint64_t dims[2] = {(uint64_t)g.K, (uint64_t)p->Cout};
    uint64_t strides[1] = {(uint64_t)g.K * 2};
    uint32_t bx[2] = {(uint32_t)(ss.row_bytes >> 1), (uint32_t)BN};
    int rc = get_tensor_map(&tmB, p->w, 2, dims, strides, bx, ss.row_bytes);
    if (rc) return rc;
  }
  {
    uint64_t dims[4] = {(uint64_t)C, (uint64_t)g.Wi, (uint64_t)g.Hi, (uint64_t)g.B};
    uint64_t strides[3] = {(uint64_t)g.ld0 * 2, (uint64_t)g.Wi * g.ld0 * 2, (uint64_t)g.Hi * g.Wi * g.ld0 * 2};
    uint32_t bx[4] = {(uint32_t)(ss.row_bytes >> 1), (uint32_t)ss.Wt, (uint32_t)(ss.S * ss.R + 2), 1u};
    int rc = get_tensor_map(&tmA, p->src0, 4, dims, strides, bx, ss.row_bytes);
    if (rc) return rc;
  }
  int occ = (227 * 1024) / (smem + 1024);
  const int tmem_cols = SlabCfg<BN>::TMEM_COLS;
  if (occ * tmem_cols > 512) occ = 512 / tmem_cols;
  if (occ < 1) occ = 1;
  int grid = ss.total < g_num_sms * occ ? ss.total : g_num_sms * occ;
  if (g_verbose) fprintf(stderr, "[d3fk] slab<%d> mode=%d M=%d C=%d Cout=%d W=%d S=%d stages=%d smem=%d grid=%d total=%d\n", BN, g.mode, g.M, C, p->Cout, ss.W, ss.S, ss.stages, smem, grid, ss.total);
  launch_k(conv_slab_kernel<BN, KSTEPS>, dim3(grid), dim3(SlabCfg<BN>::THREADS), (size_t)smem, s, dim3(1, 1, 1), tmA, tmB, e, ss,
           make_fastdiv((uint32_t)g.Wo), make_fastdiv((uint32_t)g.Ho), g_dev_error_flag);
  count_launch();
  return check_launch("conv_slab");
}

// returns 1 when the slab path took the op, 0 when it is not eligible, < 0 on error
static int try_launch_conv_slab(const Gather& g, const d3fk_conv_params* p, cudaStream_t s) {
  const int C = p->Cout;
  const int BN = C <= 16 ? 16 : C <= 32 ? 32 : C <= 64 ? 64 : C <= 128 ? 128 : 0;
  if (!BN) return 0;
  if (p->out_nchw && BN != 16) return 0;
  SlabSched ss;
  memset(&ss, 0, sizeof(ss));
  int smem = 0;
  if (!slab_plan(g, p, BN, ss, smem)) return 0;
  int rc;
#define SLAB_CASE(bn, ks) if (BN == bn && ss.ksteps == ks) rc = launch_conv_slab_bn<bn, ks>(g, p, s, ss, smem); else
  SLAB_CASE(16, 1) SLAB_CASE(16, 2) SLAB_CASE(16, 4) SLAB_CASE(32, 1) SLAB_CASE(32, 2) SLAB_CASE(32, 4)
  SLAB_CASE(64, 1) SLAB_CASE(64, 2) SLAB_CASE(64, 4) SLAB_CASE(128, 1) SLAB_CASE(128, 2) SLAB_CASE(128, 4)
  rc = set_error(D3FK_ERR_UNSUPPORTED, "slab: BN=%d ksteps=%d", BN, ss.ksteps);
#undef SLAB_CASE
  return rc ? rc : 1;
}

int try_launch_head_conv(const d3fk_conv_params* p, cudaStream_t s);
int launch_conv_tc(const d3fk_conv_params* p, cudaStream_t s) {
  Gather g;
  int rc = make_gather(g, p->src0, p->src1, p->c0, p->c1, p->ld0, p->ld1, p->up0, p->B, p->Hi, p->Wi, p->Ho, p->Wo, p->kh,
                       p->kw, p->stride, p->pad, p->mode);
  if (rc) return rc;
  D3FK_CHECK_ARG(p->out || p->out_nchw, "no output");
  D3FK_CHECK_ARG(p->out_nchw || (p->ldo % 8 == 0), "ldo must be a multiple of 8");
  D3FK_CHECK_ARG(!p->scale || p->shift, "scale requires shift");
  D3FK_CHECK_ARG(((uintptr_t)p->w & 15) == 0, "weights must be 16-byte aligned");
  const int head = try_launch_head_conv(p, s);          // 16 -> 3 channels, fp32 NCHW out: CUDA cores (head_conv.cu)
  if (head) return head < 0 ? head : D3FK_OK;
  const int slab = try_launch_conv_slab(g, p, s);
  if (slab) return slab < 0 ? slab : D3FK_OK;
  TileSched box;
  memset(&box, 0, sizeof(box));
  if (tma_box(g, p, box)) return launch_conv_tc_path<2>(g, p, s, box);
  const bool linear = p->c1 == 0 && p->up0 == 0 && (p->mode == 0 || p->stride == 1);
  return linear ? launch_conv_tc_path<0>(g, p, s, box) : launch_conv_tc_path<1>(g, p, s, box);
}

// conv + train-mode BatchNorm (+residual) + ReLU as one op: fused into the conv kernel when the layer qualifies, otherwise
// the two kernels back to back (same results up to the rounding of the raw tensor the unfused BN reads back).
int launch_bn_apply(const d3fk_bn_params* p, cudaStream_t s);
int launch_conv_bn_tc(const d3fk_convbn_params* p, cudaStream_t s) {
  FuseReq req{&p->bn, (unsigned*)p->barrier, false};
  t_fuse = p->barrier ? &req : nullptr;
  int rc = launch_conv_tc(&p->conv, s);
  t_fuse = nullptr;
  if (rc || req.taken) return rc;
  return launch_bn_apply(&p->bn, s);
}

// ------------------------------------------------------------------------------------------
// weight gradient.  Stage = 64 pixels (the MMA K dimension, 4 x K16).
//   A stage: 2 column blocks (64 k-columns each) x [64 pixels x 128 B]   (MN-major, M = k index)
//   B stage: BN/64 column blocks (64 channels each) x [64 pixels x 128 B] (MN-major, N = co)
// Grid (k tiles, cout tiles, pixel splits); the splits of one output tile form thread-block clusters of CL CTAs whose
// partial tiles are reduce-scattered through distributed shared memory, so a tile costs (splits / CL) atomic passes
// (none when splits == CL) instead of `splits`.
constexpr int WG_PIX = 64;
constexpr int WG_ONE_PER_SM_SMEM = 120 * 1024;   // more than half of the 227 KB an SM offers: one CTA per SM
constexpr int WG_A_STAGE = 2 * WG_PIX * 128;
template <int BN> struct WgradCfg {
  static constexpr int STAGES = BN >= 128 ? 3 : 4;
  static constexpr int NCB = BN / 64;
  static constexpr int B_STAGE = NCB * WG_PIX * 128;
  static constexpr int SMEM = 1024 + STAGES * (WG_A_STAGE + B_STAGE) + 256;
  static constexpr int TMEM_COLS = BN < 32 ? 32 : BN;
};

// A group of identically shaped problems shares one launch: blockIdx.z = problem * splits + pixel split (splits is a
// multiple of the cluster size, so a cluster never straddles two problems).
struct WgGroup {
  int count, splits;
  const void* src0[D3FK_WGRAD_GROUP_MAX];
  const bf16* dy[D3FK_WGRAD_GROUP_MAX];
  float* dw[D3FK_WGRAD_GROUP_MAX];
};

template <int BN>
__global__ void __launch_bounds__(WG_THREADS) wgrad_tc_kernel(Gather g, FastDiv dWo, FastDiv dHo, const bf16* __restrict__ dy_one,
                                                              int ldy, int Cout, float* __restrict__ dw_one, int cin_real,
                                                              int cout_real, int blocks_per_split, int lbo_a, int lbo_b,
                                                              int CL, int* errflag, const __grid_constant__ WgGroup grp) {
  using Cfg = WgradCfg<BN>;
  constexpr int STAGES = Cfg::STAGES;
  constexpr int CW = 32;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t a_base = base;
  const uint32_t b_base = base + STAGES * WG_A_STAGE;
  const uint32_t bar_base = b_base + STAGES * Cfg::B_STAGE;
  uint8_t* gen_bar = smem_raw + (base - smem_u32(smem_raw)) + STAGES * (WG_A_STAGE + Cfg::B_STAGE);
  volatile uint32_t* tmem_ptr_slot = reinterpret_cast<volatile uint32_t*>(gen_bar + 8 * (2 * STAGES + 1));
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (STAGES + s); };
  const uint32_t accum_bar = bar_base + 8u * (2 * STAGES);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int k0 = blockIdx.x * 128, co0 = blockIdx.y * BN;
  const int nblk_total = (g.M + WG_PIX - 1) / WG_PIX;
  const int prob = grp.count > 0 ? (int)blockIdx.z / grp.splits : 0;
  const int zsplit = (int)blockIdx.z - prob * grp.splits;
  const bf16* __restrict__ dy = grp.count > 0 ? grp.dy[prob] : dy_one;
  float* __restrict__ dw = grp.count > 0 ? grp.dw[prob] : dw_one;
  const void* src0 = grp.count > 0 ? grp.src0[prob] : g.src0;
  const int blk_beg = zsplit * blocks_per_split;
  const int blk_end = min(nblk_total, blk_beg + blocks_per_split);
  const int nblk = max(0, blk_end - blk_beg);
  const bool use_atomic = grp.splits > CL;

  if (tid == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(full_bar(s), 128);
      mbar_init(empty_bar(s), 1);
    }
    mbar_init(accum_bar, 1);
    fence_barrier_init();
  }
  if (warp == 4) tmem_alloc(smem_u32((const void*)tmem_ptr_slot), Cfg::TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_d = *tmem_ptr_slot;
  pdl_enter();   // prologue above overlaps the previous kernel; from here on its results are visible

  // output row of this thread (epilogue): k index -> (tap, ci)
  const int krow = k0 + (warp & 3) * 32 + lane;
  int tap_o = 0, ci_o = 0;
  if (krow < g.K) { tap_o = krow / g.ctot; ci_o = krow - tap_o * g.ctot; }
  const bool row_ok = krow < g.K && ci_o < cin_real;
  const int taps = g.kh * g.kw;
  auto emit4 = [&](int co, float a, float b, float c, float d) {   // columns co..co+3 of this thread's row
    float v[4] = {a, b, c, d};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      if (co + i < cout_real) {
        float* dst = dw + ((long long)(co + i) * cin_real + ci_o) * taps + tap_o;
        if (use_atomic) atomicAdd(dst, v[i]);
        else *dst = v[i];
      }
    }
  };

  if (warp < 4) {
    const int j = tid & 7;
    const int rb = tid >> 3;  // pixel rows rb + 16*i, i < 4
    const uint32_t sw = (uint32_t)((j ^ (rb & 7)) << 4);
    // the two k chunks (column blocks 0/1) this thread gathers are fixed for the whole kernel
    const bf16* sb[2];
    int kkh[2], kkw[2], sld[2], sup[2], shs[2], sws[2];
    bool kok[2];
#pragma unroll
    for (int cb = 0; cb < 2; ++cb) {
      const int k = k0 + cb * 64 + j * 8;
      kok[cb] = k < g.K;
      const int tap = kok[cb] ? k / g.ctot : 0;
      const int kc = kok[cb] ? k - tap * g.ctot : 0;
      kkh[cb] = tap / g.kw;
      kkw[cb] = tap - kkh[cb] * g.kw;
      const bool second = kc >= g.c0;
      sb[cb] = second ? (const bf16*)g.src1 + (kc - g.c0) : (const bf16*)src0 + kc;
      sld[cb] = second ? g.ld1 : g.ld0;
      sup[cb] = second ? 0 : g.up0;
      shs[cb] = g.Hi >> sup[cb];
      sws[cb] = g.Wi >> sup[cb];
    }
    for (int it = 0; it < nblk; ++it) {
      const int s = it % STAGES;
      if (it >= STAGES) mbar_wait(empty_bar(s), ((it / STAGES) - 1) & 1, errflag);
      const int mbase = (blk_beg + it) * WG_PIX;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int prow = rb + 16 * i;
        const int m = mbase + prow;
        const bool m_ok = m < g.M;
        int n = 0, h0 = -(1 << 28), w0 = 0;
        if (m_ok) {
          const uint32_t t = fdiv((uint32_t)m, dWo);
          const int wo = m - (int)t * g.Wo;
          n = (int)fdiv(t, dHo);
          const int ho = (int)t - n * g.Ho;
          h0 = ho * g.stride - g.pad;
          w0 = wo * g.stride - g.pad;
        }
#pragma unroll
        for (int cb = 0; cb < 2; ++cb) {
          const int hi = h0 + kkh[cb], wi = w0 + kkw[cb];
          const bool ok = kok[cb] && (unsigned)hi < (unsigned)g.Hi && (unsigned)wi < (unsigned)g.Wi;
          const long long pix = (long long)((n * shs[cb] + (hi >> sup[cb])) * sws[cb] + (wi >> sup[cb]));
          const void* src = ok ? (const void*)(sb[cb] + pix * sld[cb]) : src0;
          cp_async_16(a_base + s * WG_A_STAGE + cb * (WG_PIX * 128) + prow * 128 + sw, src, ok ? 16u : 0u);
        }
#pragma unroll
        for (int cb = 0; cb < Cfg::NCB; ++cb) {
          const int co = co0 + cb * 64 + j * 8;
          const bool ok = m_ok && co < Cout;
          const void* src = ok ? (const void*)(dy + (long long)m * ldy + co) : (const void*)dy;
          cp_async_16(b_base + s * Cfg::B_STAGE + cb * (WG_PIX * 128) + prow * 128 + sw, src, ok ? 16u : 0u);
        }
      }
      cp_async_mbar_arrive(full_bar(s));
      mbar_arrive(full_bar(s));
    }
    if (nblk > 0) {
      mbar_wait(accum_bar, 0, errflag);
      tc_fence_after();
    }
    if (CL == 1 && nblk > 0) {
      // no cluster: D[k row][co col] straight from TMEM (atomic when the pixels are split over several CTAs)
#pragma unroll 1
      for (int cc = 0; cc < BN; cc += CW) {
        uint32_t raw[CW];
        tmem_ld32(tmem_d + ((uint32_t)(warp * 32) << 16) + (uint32_t)cc, raw);
        tmem_ld_wait();
        if (row_ok) {
#pragma unroll
          for (int q = 0; q < CW / 4; ++q)
            emit4(co0 + cc + 4 * q, __uint_as_float(raw[4 * q]), __uint_as_float(raw[4 * q + 1]), __uint_as_float(raw[4 * q + 2]),
                  __uint_as_float(raw[4 * q + 3]));
        }
      }
    }
  } else if (warp == 4) {
    if (lane == 0 && nblk > 0) {
      constexpr uint32_t idesc = make_idesc(128, BN, 1, 1);
      for (int it = 0; it < nblk; ++it) {
        const int s = it % STAGES;
        mbar_wait(full_bar(s), (it / STAGES) & 1, errflag);
        tc_fence_after();
        const uint32_t a_addr = a_base + s * WG_A_STAGE;
        const uint32_t b_addr = b_base + s * Cfg::B_STAGE;
#pragma unroll
        for (int kk = 0; kk < WG_PIX / 16; ++kk) {
          uint64_t ad = make_smem_desc(a_addr + kk * (16 * 128), (uint32_t)lbo_a, 1024);
          uint64_t bd = make_smem_desc(b_addr + kk * (16 * 128), (uint32_t)lbo_b, 1024);
          umma_f16(tmem_d, ad, bd, idesc, (it | kk) ? 1u : 0u);
        }
        umma_commit(empty_bar(s));
      }
      umma_commit(accum_bar);
    }
    __syncwarp();
  }

  if (CL > 1) {
    // reduce-scatter the CL partial tiles through distributed shared memory (see conv_tc_kernel)
    const int SL = BN / CL, sl4 = SL >> 2;
    const uint32_t rank = cluster_ctarank();
    const uint32_t recv = a_base;   // [CL][SL/4][128 rows] float4 over the dead pipeline stages
    const int row = (warp & 3) * 32 + lane;
    tc_fence_before();
    cluster_sync_all();
    tc_fence_after();
    if (warp < 4) {
#pragma unroll 1
      for (int cc = 0; cc < BN; cc += CW) {
        uint32_t raw[CW];
        if (nblk > 0) {
          tmem_ld32(tmem_d + ((uint32_t)(warp * 32) << 16) + (uint32_t)cc, raw);
          tmem_ld_wait();
        } else {
#pragma unroll
          for (int i = 0; i < CW; ++i) raw[i] = 0u;   // a split with no pixel blocks contributes zeros
        }
#pragma unroll
        for (int q = 0; q < CW / 4; ++q) {
          const int col = cc + 4 * q;
          const int owner = col / SL, within = col - owner * SL;
          const uint32_t la = recv + (uint32_t)((((int)rank * sl4 + (within >> 2)) * 128 + row) * 16);
          st_cluster_f4(mapa_shared(la, (uint32_t)owner), raw[4 * q], raw[4 * q + 1], raw[4 * q + 2], raw[4 * q + 3]);
        }
      }
    }
    cluster_sync_all();
    if (warp < 4 && row_ok) {
      for (int c4 = 0; c4 < sl4; ++c4) {
        float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int r = 0; r < CL; ++r) {
          const float4 v = ld_shared_f4(recv + (uint32_t)(((r * sl4 + c4) * 128 + row) * 16));
          a.x += v.x; a.y += v.y; a.z += v.z; a.w += v.w;
        }
        emit4(co0 + (int)rank * SL + 4 * c4, a.x, a.y, a.z, a.w);
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 4) tmem_dealloc(tmem_d, Cfg::TMEM_COLS);
}

static int g_wg_ctas_per_sm = 2;
static int g_wg_cap = 0;     // D3FK_WG_CAP=n: override the co-resident CTA capacity used to size the pixel splits
static int g_wg_plain = 1;   // D3FK_WG_PLAIN=0: launch cluster-size-1 grids through cudaLaunchKernelEx too
// Grouped launches run for 50-90 us on a side stream while the latency-bound main chain needs a free CTA slot every few
// microseconds: D3FK_WG_GROUP_OCC=1 pads their shared-memory request so that only ONE of them fits an SM and the other slot
// stays available to the main chain.
static int g_wg_group_occ = 2;

template <int BN>
static int launch_wgrad_tc_bn(const Gather& g, const d3fk_wgrad_params* p, cudaStream_t s, const d3fk_wgrad_group_params* group = nullptr) {
  const int gx = cdiv(g.K, 128), gy = cdiv(p->Cout, BN);
  const int nblk = cdiv(g.M, WG_PIX);
  const int G = group ? group->count : 1;
  const int tiles = gx * gy * G;
  // Pixel splits and cluster size: one wave of co-resident CTAs.  Cost model (us): pipeline stages per CTA, plus the atomic
  // passes over the K x Cout outputs when a tile's splits span several clusters, plus the cluster reduction itself.
  const double elems = (double)g.K * p->cout_real * G;
  const int max_cl = g_max_cluster < 8 ? (g_max_cluster < 1 ? 1 : g_max_cluster) : 8;
  int cl = 1, splits = 1;
  double best = 1e30;
  for (int c = 1; c <= max_cl; c *= 2) {
    int cap = g_wg_cap > 0 ? g_wg_cap : cluster_capacity(c, (group && g_wg_group_occ == 1) ? 1 : g_wg_ctas_per_sm);
    int smax = cap / tiles;
    if (smax > nblk) smax = nblk;
    smax = (smax / c) * c;
    if (smax < c) {
      if (c == 1) { best = (double)nblk * 0.4 * cdiv(tiles, cap); cl = 1; splits = 1; }   // more tiles than one wave: no split
      continue;
    }
    const int cand[2] = {smax, c};
    for (int i = 0; i < 2; ++i) {
      const int sp = cand[i];
      const double est = (double)cdiv(nblk, sp) * 0.4 + (sp > c ? (sp / c) * elems / 216e3 : 0.0) + (c > 1 ? 1.0 : 0.0);
      if (est < best) { best = est; cl = c; splits = sp; }
    }
  }
  const int bps = cdiv(nblk, splits);
  WgGroup grp;
  memset(&grp, 0, sizeof(grp));
  grp.splits = splits;
  if (group) {
    grp.count = G;
    for (int i = 0; i < G; ++i) { grp.src0[i] = group->src0[i]; grp.dy[i] = (const bf16*)group->dy[i]; grp.dw[i] = group->dw[i]; }
  }
  dim3 grid(gx, gy, splits * G);
  const size_t smem = (group && g_wg_group_occ == 1) ? (size_t)WG_ONE_PER_SM_SMEM : (size_t)WgradCfg<BN>::SMEM;
  if (g_verbose) fprintf(stderr, "[d3fk] wgrad<%d> M=%d K=%d Cout=%d group=%d tiles=%d cl=%d splits=%d bps=%d\n", BN, g.M, g.K, p->Cout, G, tiles, cl, splits, bps);
  if (cl == 1 && g_wg_plain) {
    launch_k(wgrad_tc_kernel<BN>, dim3(grid), dim3(WG_THREADS), smem, s, dim3(1, 1, 1), g, make_fastdiv((uint32_t)g.Wo), make_fastdiv((uint32_t)g.Ho),
                                                                  (const bf16*)p->dy, p->ldy, p->Cout, p->dw, p->cin_real, p->cout_real,
                                                                  bps, WG_PIX * 128, WG_PIX * 128, cl, g_dev_error_flag, grp);
  } else {
    cudaError_t le = launch_k(wgrad_tc_kernel<BN>, grid, dim3(WG_THREADS), smem, s, dim3(1, 1, cl), g,
                                    make_fastdiv((uint32_t)g.Wo), make_fastdiv((uint32_t)g.Ho), (const bf16*)p->dy, p->ldy, p->Cout,
                                    p->dw, p->cin_real, p->cout_real, bps, WG_PIX * 128, WG_PIX * 128, cl, g_dev_error_flag, grp);
    if (le != cudaSuccess) return set_error(D3FK_ERR_CUDA, "wgrad_tc launch: %s", cudaGetErrorString(le));
  }
  count_launch();
  return check_launch("wgrad_tc");
}

// ------------------------------------------------------------------------------------------
// Slab weight gradient: 3x3 / stride 1 / pad 1, one source, Cin in {16, 32, 64}, Cout in {16, 32}, large images.
// dW[kh][kw][ci][co] = sum_pix A[pix + (kh-1, kw-1)][ci] * dY[pix][co].  A persistent CTA walks super-tiles of S*128
// pixels; per super-tile it lands the same three column-shifted activation slabs as conv_slab_kernel plus the dY tile.
// Both operands are MN-major (pixel rows are the MMA K dimension).  The A operand of ONE tcgen05.mma is M = 128 =
// (128 / Cin) "atoms" of Cin channels whose leading-dimension stride is one image row of the slab — i.e. one MMA covers
// the taps kh = 0, 1, 2 (... surplus atoms read further rows and land in accumulator rows nobody reads) of one kw for 16
// pixels.  The 3 (x2 for Cin = 64) accumulators stay in TMEM for the whole kernel: no per-tile epilogue at all, one
// atomic pass per CTA at the end.  L2 -> SM traffic per pixel: 3*(S*R+2)/(S*R) activation reads + 1 dY read instead of 9 + 9.
struct WgSlabSched {
  int W, H, R, S, Wt, wtiles;
  int C;                  // input channels (16/32/64); a_row_bytes = 2*C
  int a_row_bytes, b_row_bytes;
  int slab_bytes, slab_tx, dy_bytes, stage_bytes, stages;
  int total, tiles_per_img;
  int MB;                 // accumulator row blocks per kw: 1 (Cin <= 32: kh 0..2 in one M = 128) or 2 (Cin = 64)
  uint32_t a_layout, b_layout;
};

__device__ __forceinline__ uint64_t make_smem_desc_mn(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;   // stride between M (N) atoms of one swizzle width
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;   // stride between 8-row (K) groups
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)layout << 61;
  return d;
}

constexpr int WGS_NW = 6;                      // MMA-issuing warps of wgrad_slab_kernel
constexpr int WGS_THREADS = (5 + WGS_NW) * 32;  // warps 0-3 epilogue, 4 TMA producer, 5-10 MMA issuers
template <int BN>
__global__ void __launch_bounds__(WGS_THREADS) wgrad_slab_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmD,
                                                                WgSlabSched ss, float* __restrict__ dw, int cin_real, int cout_real,
                                                                int* errflag) {
  constexpr int ACC = BN < 32 ? 32 : BN;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t stage_base = base;                                      // [stages][3 slabs | dY tile]
  const uint32_t bar_base = stage_base + ss.stages * ss.stage_bytes;     // full[4], empty[4], acc_full
  uint8_t* gen_bar = smem_raw + (bar_base - smem_u32(smem_raw));
  volatile uint32_t* tmem_ptr_slot = reinterpret_cast<volatile uint32_t*>(gen_bar + 8 * 9);
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (4 + s); };
  const uint32_t acc_full = bar_base + 8u * 8;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int nacc = 3 * ss.MB;                       // accumulators per K-step parity set
  uint32_t tmem_cols = 32;
  while ((int)tmem_cols < 2 * nacc * ACC) tmem_cols <<= 1;
  const bool has_work = (int)blockIdx.x < ss.total;

  if (tid == 0) {
    for (int s = 0; s < 4; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), WGS_NW);
    }
    mbar_init(acc_full, WGS_NW);
    fence_barrier_init();
  }
  if (warp == 4 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmD);
  }
  if (warp == 5) tmem_alloc(smem_u32((const void*)tmem_ptr_slot), tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_d = *tmem_ptr_slot;
  pdl_enter();

  if (warp == 4) {
    if (lane == 0) {
      const int rows = ss.S * ss.R;
      uint32_t it = 0;
      for (int t = blockIdx.x; t < ss.total; t += gridDim.x, ++it) {
        const int n = t / ss.tiles_per_img;
        const int rem = t - n * ss.tiles_per_img;
        const int hb = rem / ss.wtiles;
        const int h0 = hb * rows, w0 = (rem - hb * ss.wtiles) * ss.Wt;
        const int st = it % ss.stages;
        if (it >= (uint32_t)ss.stages) mbar_wait(empty_bar(st), ((it / ss.stages) - 1) & 1, errflag);
        mbar_arrive_expect_tx(full_bar(st), 3u * ss.slab_tx + (uint32_t)ss.dy_bytes);
        const uint32_t sb = stage_base + st * ss.stage_bytes;
        for (int sx = 0; sx < 3; ++sx) tma_load_4d(sb + sx * ss.slab_bytes, &tmA, 0, w0 + sx - 1, h0 - 1, n, full_bar(st));
        tma_load_4d(sb + 3 * ss.slab_bytes, &tmD, 0, w0, h0, n, full_bar(st));
      }
    }
    __syncwarp();
  } else if (warp >= 5) {
    // MMA issuers: warp w owns kw = w % 3 and the K steps (16-pixel groups) of parity w / 3, in its own accumulators —
    // six independent issue streams (one thread sustains only ~1 small MMA per 90 cycles).
    if (has_work) {
      const int w = warp - 5;
      const int sx = w % 3, par = w / 3;
      const uint32_t leader = lane == 0 ? 1u : 0u;
      constexpr uint32_t idesc = make_idesc(128, BN, 1, 1);
      const uint32_t img_row = (uint32_t)(ss.Wt * ss.a_row_bytes);       // M-atom stride of the A operand = one tap row (kh)
      const uint64_t a_t = make_smem_desc_mn(0, img_row, 8u * ss.a_row_bytes, ss.a_layout);
      const uint64_t b_t = make_smem_desc_mn(0, 8u * ss.b_row_bytes, 8u * ss.b_row_bytes, ss.b_layout);
      const uint32_t ahi = (uint32_t)(a_t >> 32), alo0 = (uint32_t)a_t, bhi = (uint32_t)(b_t >> 32), blo0 = (uint32_t)b_t;
      const uint32_t a_step2 = (32u * ss.a_row_bytes) >> 4, b_step2 = (32u * ss.b_row_bytes) >> 4;   // two K steps
      const uint32_t a_par = (uint32_t)par * ((16u * ss.a_row_bytes) >> 4), b_par = (uint32_t)par * ((16u * ss.b_row_bytes) >> 4);
      const uint32_t mb_off = ((uint32_t)(128 / ss.C) * img_row) >> 4;   // second row block (Cin = 64): taps kh = 2, (3)
      const int ksteps2 = ss.S * 4;                                       // K steps of this parity per super-tile
      const uint32_t d_base = tmem_d + (uint32_t)((par * nacc + sx * ss.MB) * ACC);
      uint32_t it = 0;
      for (int t = blockIdx.x; t < ss.total; t += gridDim.x, ++it) {
        const int st = it % ss.stages;
        mbar_wait(full_bar(st), (it / ss.stages) & 1, errflag);
        tc_fence_after();
        const uint32_t sb = stage_base + st * ss.stage_bytes;
        const uint32_t b_lo = (blo0 | ((sb + 3u * ss.slab_bytes) >> 4)) + b_par;
        const uint32_t a_lo = (alo0 | ((sb + (uint32_t)(sx * ss.slab_bytes)) >> 4)) + a_par;
        for (int mb = 0; mb < ss.MB; ++mb) {
          const uint32_t d_addr = d_base + (uint32_t)(mb * ACC);
          uint32_t a_cur = a_lo + (uint32_t)mb * mb_off, b_cur = b_lo;
          umma_f16_lohi_p(d_addr, a_cur, ahi, b_cur, bhi, idesc, it ? 1u : 0u, leader);
#pragma unroll 4
          for (int j = 1; j < ksteps2; ++j) {
            a_cur += a_step2;
            b_cur += b_step2;
            umma_f16_lohi_p(d_addr, a_cur, ahi, b_cur, bhi, idesc, 1u, leader);
          }
        }
        umma_commit_p(empty_bar(st), leader);
      }
      umma_commit_p(acc_full, leader);
    }
    __syncwarp();
  } else if (has_work) {
    // epilogue (once per CTA): accumulator row r of (kw, row block mb) = tap kh = mb*(128/C) + r / C, channel ci = r % C
    mbar_wait(acc_full, 0, errflag);
    tc_fence_after();
    const int r = warp * 32 + lane;
    const int apm = 128 / ss.C;
    for (int sx = 0; sx < 3; ++sx) {
      for (int mb = 0; mb < ss.MB; ++mb) {
        const int kh = mb * apm + r / ss.C, ci = r % ss.C;
        const bool ok = kh < 3 && ci < cin_real;
#pragma unroll 1
        for (int cc = 0; cc < BN; cc += 16) {
          uint32_t raw[16], raw2[16];
          const uint32_t ta = tmem_d + ((uint32_t)(warp * 32) << 16) + (uint32_t)((sx * ss.MB + mb) * ACC + cc);
          tmem_ld16(ta, raw);
          tmem_ld16(ta + (uint32_t)(nacc * ACC), raw2);     // the odd-K-step accumulator set
          tmem_ld_wait();
          if (ok) {
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              const int co = cc + i;
              if (co < cout_real)
                atomicAdd(dw + ((long long)co * cin_real + ci) * 9 + kh * 3 + sx, __uint_as_float(raw[i]) + __uint_as_float(raw2[i]));
            }
          }
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 5) tmem_dealloc(tmem_d, tmem_cols);
}

static int g_use_wg_slab = 1;   // D3FK_WG_SLAB=0: never take the slab weight-gradient path

template <int BN>
static int launch_wgrad_slab_bn(const Gather& g, const d3fk_wgrad_params* p, cudaStream_t s, WgSlabSched& ss) {
  const int C = g.ctot, W = p->Wi, H = p->Hi;
  const int Wt = W < 128 ? W : 128, R = 128 / Wt;
  ss.W = W; ss.H = H; ss.R = R; ss.Wt = Wt; ss.wtiles = W / Wt; ss.C = C;
  ss.a_row_bytes = 2 * C;
  ss.b_row_bytes = 2 * BN;
  ss.a_layout = C == 64 ? 2u : C == 32 ? 4u : 6u;
  ss.b_layout = BN == 64 ? 2u : BN == 32 ? 4u : 6u;
  ss.MB = C == 64 ? 2 : 1;
  const int ACC = BN < 32 ? 32 : BN;
  if (2 * 3 * ss.MB * ACC > 512) return 0;
  int smem = 0;
  bool found = false;
  for (int S = 4; S >= 1 && !found; S >>= 1) {
    if (H % (S * R)) continue;
    const int slab = ((S * R + 2) * Wt * ss.a_row_bytes + 1023) & ~1023;
    const int dyb = (S * 128 * ss.b_row_bytes + 1023) & ~1023;
    // surplus M atoms (Cin < 64: 128/C - 3 of them) read up to (128/C - 3) image rows past the last slab: they must stay
    // inside the stage (the dY tile that follows the slabs absorbs them)
    const int overrun = (128 / C > 3 ? 128 / C - 3 : (C == 64 ? 1 : 0)) * Wt * ss.a_row_bytes;
    if (overrun > dyb) continue;
    for (int stages = 3; stages >= 2; --stages) {
      const int need = 1024 + stages * (3 * slab + dyb) + 128;
      if (need > SLAB_MAX_SMEM) continue;
      ss.S = S; ss.slab_bytes = slab; ss.slab_tx = (S * R + 2) * Wt * ss.a_row_bytes; ss.dy_bytes = S * 128 * ss.b_row_bytes;
      ss.stage_bytes = 3 * slab + dyb; ss.stages = stages;
      ss.tiles_per_img = (H / (S * R)) * ss.wtiles;
      ss.total = p->B * ss.tiles_per_img;
      smem = need;
      found = true;
      break;
    }
  }
  if (!found) return 0;
  alignas(64) CUtensorMap tmA, tmD;
  {
    uint64_t dims[4] = {(uint64_t)C, (uint64_t)W, (uint64_t)H, (uint64_t)p->B};
    uint64_t strides[3] = {(uint64_t)p->ld0 * 2, (uint64_t)W * p->ld0 * 2, (uint64_t)H * W * p->ld0 * 2};
    uint32_t bx[4] = {(uint32_t)C, (uint32_t)Wt, (uint32_t)(ss.S * R + 2), 1u};
    int rc = get_tensor_map(&tmA, p->src0, 4, dims, strides, bx, ss.a_row_bytes);
    if (rc) return rc;
  }
  {
    uint64_t dims[4] = {(uint64_t)BN, (uint64_t)W, (uint64_t)H, (uint64_t)p->B};
    uint64_t strides[3] = {(uint64_t)p->ldy * 2, (uint64_t)W * p->ldy * 2, (uint64_t)H * W * p->ldy * 2};
    uint32_t bx[4] = {(uint32_t)BN, (uint32_t)Wt, (uint32_t)(ss.S * R), 1u};
    int rc = get_tensor_map(&tmD, p->dy, 4, dims, strides, bx, ss.b_row_bytes);
    if (rc) return rc;
  }
  int grid = ss.total < g_num_sms ? ss.total : g_num_sms;
  if (g_verbose) fprintf(stderr, "[d3fk] wgrad_slab<%d> M=%d C=%d W=%d S=%d stages=%d smem=%d grid=%d total=%d\n", BN, g.M, C, W, ss.S, ss.stages, smem, grid, ss.total);
  launch_k(wgrad_slab_kernel<BN>, dim3(grid), dim3(WGS_THREADS), (size_t)smem, s, dim3(1, 1, 1), tmA, tmD, ss, p->dw, p->cin_real,
           p->cout_real, g_dev_error_flag);
  count_launch();
  int rc = check_launch("wgrad_slab");
  return rc ? rc : 1;
}

// returns 1 when taken, 0 when not eligible, < 0 on error
static int try_launch_wgrad_slab(const Gather& g, const d3fk_wgrad_params* p, cudaStream_t s) {
  if (!g_use_wg_slab) return 0;
  if (p->kh != 3 || p->kw != 3 || p->stride != 1 || p->pad != 1 || p->c1 != 0 || p->up0 != 0) return 0;
  if (p->Ho != p->Hi || p->Wo != p->Wi) return 0;
  const int C = g.ctot, W = p->Wi;
  if (C != 16 && C != 32 && C != 64) return 0;
  if (p->Cout != 16 && p->Cout != 32) return 0;
  if (W != 16 && W != 32 && W != 64 && (W % 128)) return 0;
  if (((uintptr_t)p->src0 & 15) || ((uintptr_t)p->dy & 15) || (p->ld0 % 8) || (p->ldy % 8)) return 0;
  if ((long long)g.M < 128ll * 148 * 4) return 0;   // small problems: the per-CTA atomic pass would dominate
  WgSlabSched ss;
  memset(&ss, 0, sizeof(ss));
  if (p->Cout == 16) return launch_wgrad_slab_bn<16>(g, p, s, ss);
  return launch_wgrad_slab_bn<32>(g, p, s, ss);
}

int launch_wgrad_tc(const d3fk_wgrad_params* p, cudaStream_t s) {
  Gather g;
  int rc = make_gather(g, p->src0, p->src1, p->c0, p->c1, p->ld0, p->ld1, p->up0, p->B, p->Hi, p->Wi, p->Ho, p->Wo, p->kh,
                       p->kw, p->stride, p->pad, 0);
  if (rc) return rc;
  D3FK_CHECK_ARG(p->Cout % 8 == 0 && p->ldy % 8 == 0, "Cout and ldy must be multiples of 8");
  const int slab = try_launch_wgrad_slab(g, p, s);
  if (slab) return slab < 0 ? slab : D3FK_OK;
  if (p->Cout > 64) return launch_wgrad_tc_bn<128>(g, p, s);
  return launch_wgrad_tc_bn<64>(g, p, s);
}

int launch_wgrad_group_tc(const d3fk_wgrad_group_params* gp, cudaStream_t s) {
  const d3fk_wgrad_params* p = &gp->base;
  D3FK_CHECK_ARG(gp->count >= 1 && gp->count <= D3FK_WGRAD_GROUP_MAX, "wgrad group: count out of range");
  D3FK_CHECK_ARG(p->c1 == 0 && p->src1 == nullptr, "wgrad group: single-source layers only");
  for (int i = 0; i < gp->count; ++i) {
    D3FK_CHECK_ARG(gp->src0[i] && gp->dy[i] && gp->dw[i], "wgrad group: null pointer in a problem");
    D3FK_CHECK_ARG((((uintptr_t)gp->src0[i] | (uintptr_t)gp->dy[i]) & 15) == 0, "wgrad group: operands must be 16-byte aligned");
  }
  Gather g;
  int rc = make_gather(g, gp->src0[0], nullptr, p->c0, 0, p->ld0, p->ld1, p->up0, p->B, p->Hi, p->Wi, p->Ho, p->Wo, p->kh,
                       p->kw, p->stride, p->pad, 0);
  if (rc) return rc;
  D3FK_CHECK_ARG(p->Cout % 8 == 0 && p->ldy % 8 == 0, "Cout and ldy must be multiples of 8");
  if (p->Cout > 64) return launch_wgrad_tc_bn<128>(g, p, s, gp);
  return launch_wgrad_tc_bn<64>(g, p, s, gp);
}

int tc_init() {
  cudaError_t e = cudaSuccess;
  int dev = 0, sms = 0;
  if (cudaGetDevice(&dev) == cudaSuccess && cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && sms > 0)
    g_num_sms = sms;
  const char* tl = getenv("D3FK_TILE_LOOP");
  if (tl) g_tile_loop = atoi(tl);
  if (const char* v = getenv("D3FK_A_CA")) g_a_ca = atoi(v);
  if (const char* v = getenv("D3FK_OCC")) g_occ_cap = atoi(v);
  if (const char* v = getenv("D3FK_TMA_A")) g_use_tma_a = atoi(v);
  if (const char* v = getenv("D3FK_SPLIT_TILES")) g_split_tiles = atoi(v);
  if (const char* v = getenv("D3FK_CLUSTER")) g_max_cluster = atoi(v);
  if (const char* v = getenv("D3FK_VERBOSE")) g_verbose = atoi(v);
  if (const char* v = getenv("D3FK_SLAB")) g_use_slab = atoi(v);
  if (const char* v = getenv("D3FK_FUSE_BN")) g_fuse_bn = atoi(v);
  if (const char* v = getenv("D3FK_WG_SLAB")) g_use_wg_slab = atoi(v);
  if (const char* v = getenv("D3FK_WG_OCC")) g_wg_ctas_per_sm = atoi(v);
  if (const char* v = getenv("D3FK_WG_CAP")) g_wg_cap = atoi(v);
  if (const char* v = getenv("D3FK_WG_PLAIN")) g_wg_plain = atoi(v);
  if (const char* v = getenv("D3FK_WG_GROUP_OCC")) g_wg_group_occ = atoi(v);
#define SET_SMEM(k, bytes)                                                                          \
  if (e == cudaSuccess) e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);               \
  if (e == cudaSuccess) e = cudaFuncSetAttribute(k, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
  SET_SMEM((conv_tc_kernel<16, 0, 0>), ConvCfg<16>::SMEM)
  SET_SMEM((conv_tc_kernel<32, 0, 0>), ConvCfg<32>::SMEM)
  SET_SMEM((conv_tc_kernel<64, 0, 0>), ConvCfg<64>::SMEM)
  SET_SMEM((conv_tc_kernel<128, 0, 0>), ConvCfg<128>::SMEM)
  SET_SMEM((conv_tc_kernel<16, 1, 0>), ConvCfg<16>::SMEM)
  SET_SMEM((conv_tc_kernel<32, 1, 0>), ConvCfg<32>::SMEM)
  SET_SMEM((conv_tc_kernel<64, 1, 0>), ConvCfg<64>::SMEM)
  SET_SMEM((conv_tc_kernel<128, 1, 0>), ConvCfg<128>::SMEM)
  SET_SMEM((conv_tc_kernel<16, 2, 0>), ConvCfg<16>::SMEM)
  SET_SMEM((conv_tc_kernel<32, 2, 0>), ConvCfg<32>::SMEM)
  SET_SMEM((conv_tc_kernel<64, 2, 0>), ConvCfg<64>::SMEM)
  SET_SMEM((conv_tc_kernel<128, 2, 0>), ConvCfg<128>::SMEM)
  SET_SMEM((conv_tc_kernel<16, 0, 2>), ConvCfg<16>::SMEM)
  SET_SMEM((conv_tc_kernel<16, 1, 2>), ConvCfg<16>::SMEM)
  SET_SMEM((conv_tc_kernel<16, 2, 2>), ConvCfg<16>::SMEM)
  SET_SMEM((conv_tc_kernel<32, 0, 2>), ConvCfg<32>::SMEM)
  SET_SMEM((conv_tc_kernel<32, 1, 2>), ConvCfg<32>::SMEM)
  SET_SMEM((conv_tc_kernel<32, 2, 2>), ConvCfg<32>::SMEM)
  SET_SMEM((conv_tc_kernel<64, 0, 2>), ConvCfg<64>::SMEM)
  SET_SMEM((conv_tc_kernel<64, 1, 2>), ConvCfg<64>::SMEM)
  SET_SMEM((conv_tc_kernel<64, 2, 2>), ConvCfg<64>::SMEM)
  SET_SMEM((conv_tc_kernel<128, 0, 2>), ConvCfg<128>::SMEM)
  SET_SMEM((conv_tc_kernel<128, 1, 2>), ConvCfg<128>::SMEM)
  SET_SMEM((conv_tc_kernel<128, 2, 2>), ConvCfg<128>::SMEM)
  SET_SMEM((conv_tc_kernel<128, 0, 1>), ConvCfg<128>::SMEM)
  SET_SMEM((conv_tc_kernel<128, 1, 1>), ConvCfg<128>::SMEM)
  SET_SMEM((conv_tc_kernel<128, 2, 1>), ConvCfg<128>::SMEM)
  SET_SMEM((conv_slab_kernel<16, 1>), SLAB_MAX_SMEM)
  SET_SMEM((conv_slab_kernel<16, 2>), SLAB_MAX_SMEM)
  SET_SMEM((conv_slab_kernel<16, 4>), SLAB_MAX_SMEM)
  SET_SMEM((conv_slab_kernel<32, 1>), SLAB_MAX_SMEM)
  SET_SMEM((conv_slab_kernel<32, 2>), SLAB_MAX_SMEM)
  SET_SMEM((conv_slab_kernel<32, 4>), SLAB_MAX_SMEM)
  SET_SMEM((conv_slab_kernel<64, 1>), SLAB_MAX_SMEM)
  SET_SMEM((conv_slab_kernel<64, 2>), SLAB_MAX_SMEM)
  SET_SMEM((conv_slab_kernel<64, 4>), SLAB_MAX_SMEM)
  SET_SMEM((conv_slab_kernel<128, 1>), SLAB_MAX_SMEM)
  SET_SMEM((conv_slab_kernel<128, 2>), SLAB_MAX_SMEM)
  SET_SMEM((conv_slab_kernel<128, 4>), SLAB_MAX_SMEM)
  SET_SMEM(wgrad_slab_kernel<16>, SLAB_MAX_SMEM)
  SET_SMEM(wgrad_slab_kernel<32>, SLAB_MAX_SMEM)
  SET_SMEM(wgrad_tc_kernel<64>, WG_ONE_PER_SM_SMEM)
  SET_SMEM(wgrad_tc_kernel<128>, WG_ONE_PER_SM_SMEM)
#undef SET_SMEM
  if (e != cudaSuccess) return set_error(D3FK_ERR_CUDA, "cudaFuncSetAttribute: %s", cudaGetErrorString(e));
  return D3FK_OK;
}

}  // namespace d3fk
