// wgrad_tc.cu — weight gradients on the tcgen05 tensor cores: the generic MN-major kernel (cluster-reduced pixel
// splits, grouped launches) and the slab kernel for the narrow 3x3 decoder layers.
#include "tc_common.cuh"

namespace d3fk {

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

// TMA variant (3x3-style "same" convolutions at stride 1 with Cin % 64 == 0 and power-of-two extents — the 33 encoder /
// decoder-conv2 weight gradients of the step): a 64-pixel block is an axis-aligned (w, h, n) box, so one column block of the
// A stage (64 channels of one tap for 64 pixels) is ONE cp.async.bulk.tensor.4d — shifted by the tap, hardware zero fill for
// the halo — and the dY tile is a 2-D box per 64 output channels.  One elected thread of warp 0 is the producer; nobody
// computes an address (the gather variant issues 16 cp.async and ~150 instructions per thread and stage).
struct WgBox { int bw, bh, bn, wt, ht; int win; };   // win: windowed rows of the space-to-depth stem (d3fk_wgrad_params.mode 2)
template <int BN, bool TMA>
__global__ void __launch_bounds__(WG_THREADS) wgrad_tc_kernel(Gather g, FastDiv dWo, FastDiv dHo, const bf16* __restrict__ dy_one,
                                                              int ldy, int Cout, float* __restrict__ dw_one, int cin_real,
                                                              int cout_real, int blocks_per_split, int lbo_a, int lbo_b,
                                                              int CL, int* errflag, const __grid_constant__ WgGroup grp,
                                                              const __grid_constant__ CUtensorMap tmA,
                                                              const __grid_constant__ CUtensorMap tmD, WgBox box) {
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
  const uint32_t recv_bar = bar_base + 8u * (2 * STAGES + 2);   // cluster reduction: the partial tiles of my column slice have landed

  const int tid = threadIdx.x, warp = warp_index_uniform(), lane = tid & 31;
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
      mbar_init(full_bar(s), TMA ? 1 : 128);
      mbar_init(empty_bar(s), 1);
    }
    mbar_init(accum_bar, 1);
    mbar_init(recv_bar, 1);
    fence_barrier_init();
  }
  if (TMA && warp == 0 && elect_one_sync()) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmD);
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
  bool row_ok = krow < g.K && ci_o < cin_real;
  int taps = g.kh * g.kw;
  if (TMA && box.win) {
    // windowed rows: k = th*64 + tw*16 + (dy*2+dx)*4 + ci  ->  tap (2th+dy-1, 2tw+dx-1) of the 7x7 master gradient, channel ci
    const int kh7 = 2 * (krow >> 6) + ((krow >> 3) & 1) - 1, kw7 = 2 * ((krow >> 4) & 3) + ((krow >> 2) & 1) - 1;
    ci_o = krow & 3;
    tap_o = kh7 * 7 + kw7;
    taps = 49;
    row_ok = krow < g.K && ci_o < cin_real && kh7 >= 0 && kh7 < 7 && kw7 >= 0 && kw7 < 7;
  }
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
    if (TMA) {
      // ===================== TMA producer (one elected thread of warp 0) =====================
      if (warp == 0 && elect_one_sync()) {
        int kc[2], dh[2], dwv[2];
        bool kok[2];
#pragma unroll
        for (int cb = 0; cb < 2; ++cb) {
          const int k = k0 + cb * 64;
          kok[cb] = k < g.K;
          const int tap = kok[cb] ? k / g.ctot : 0;
          kc[cb] = kok[cb] ? k - tap * g.ctot : 0;
          const int khi = tap / g.kw;
          dh[cb] = khi - g.pad;
          dwv[cb] = box.win ? 0 : tap - khi * g.kw - g.pad;      // windowed rows: the kw taps are inside the 128-byte window
        }
        const uint32_t tx = (uint32_t)(((kok[0] ? 1 : 0) + (kok[1] ? 1 : 0) + Cfg::NCB) * (WG_PIX * 128));
        for (int it = 0; it < nblk; ++it) {
          const int s = it % STAGES;
          if (it >= STAGES) mbar_wait(empty_bar(s), ((it / STAGES) - 1) & 1, errflag);
          const int blk = blk_beg + it;
          const int tw = blk % box.wt, r2 = blk / box.wt;
          const int w0 = tw * box.bw, h0 = (r2 % box.ht) * box.bh, n0 = (r2 / box.ht) * box.bn;
          mbar_arrive_expect_tx(full_bar(s), tx);
#pragma unroll
          for (int cb = 0; cb < 2; ++cb)      // (a column block beyond K is not loaded: its accumulator rows are never stored)
            if (kok[cb]) tma_load_4d(a_base + s * WG_A_STAGE + cb * (WG_PIX * 128), &tmA, kc[cb], w0 + dwv[cb], h0 + dh[cb], n0, full_bar(s));
#pragma unroll
          for (int cb = 0; cb < Cfg::NCB; ++cb)
            tma_load_2d(b_base + s * Cfg::B_STAGE + cb * (WG_PIX * 128), &tmD, co0 + cb * 64, blk * WG_PIX, full_bar(s));
        }
      }
      __syncwarp();
    }
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
    for (int it = 0; it < (TMA ? 0 : nblk); ++it) {
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
    if (nblk > 0 && elect_one_sync()) {
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
    // reduce-scatter the CL partial tiles through distributed shared memory, as conv_tc_kernel does: one relaxed cluster
    // barrier ("every CTA is past its main loop": the stages are dead), then st.async stores that count their bytes on the
    // OWNER's mbarrier — no release / acquire cluster barrier (MEMBAR.ALL.GPU), an owner starts as soon as ITS slice is in.
    const int SL = BN / CL, sl4 = SL >> 2;
    const uint32_t rank = cluster_ctarank();
    const uint32_t recv = a_base;   // [CL][SL/4][128 rows] float4 over the dead pipeline stages
    const int row = (warp & 3) * 32 + lane;
    if (tid == 0) mbar_arrive_expect_tx(recv_bar, (uint32_t)(128 * BN * 4));
    tc_fence_before();
    cluster_sync_relaxed();
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
          st_async_f4(mapa_shared(la, (uint32_t)owner), raw[4 * q], raw[4 * q + 1], raw[4 * q + 2], raw[4 * q + 3],
                      mapa_shared(recv_bar, (uint32_t)owner));
        }
      }
      mbar_wait(recv_bar, 0, errflag);
      if (row_ok) {
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

static int g_wg_tma = 1;     // D3FK_WG_TMA=0 (debug builds): always gather

// TMA eligibility of the generic weight gradient: single source, no upsample, stride 1, "same" geometry, Cin % 64 == 0, and
// the 64-pixel block an axis-aligned (w, h, n) box in linear pixel order (as tma_box() of the convolution, with 64 pixels).
static bool wg_tma_box(const Gather& g, const d3fk_wgrad_params* p, WgBox& b) {
  if (!g_wg_tma) return false;
  if (p->c1 != 0 || p->src1 || p->up0 != 0 || p->stride != 1 || g.ctot % 64 != 0) return false;
  if (p->Ho != p->Hi || p->Wo != p->Wi || 2 * p->pad != p->kh - 1 || p->kh != p->kw) return false;
  if (((uintptr_t)p->src0 & 15) || ((uintptr_t)p->dy & 15) || (g.ld0 % 8) || (p->ldy % 8)) return false;
  const int W = p->Wi, H = p->Hi;
  int bw = W < WG_PIX ? W : WG_PIX;
  if (WG_PIX % bw || W % bw) return false;
  int bh = WG_PIX / bw;
  if (bh > H) bh = H;
  if ((WG_PIX / bw) % bh || H % bh) return false;
  const int bn = WG_PIX / (bw * bh);
  if (bw < W && bh != 1) return false;
  if (bh < H && bn != 1) return false;
  if (bn > 256) return false;
  b.bw = bw; b.bh = bh; b.bn = bn; b.wt = W / bw; b.ht = H / bh;
  return true;
}

template <int BN>
static int launch_wgrad_tc_bn(const Gather& g, const d3fk_wgrad_params* p, cudaStream_t s, const d3fk_wgrad_group_params* group = nullptr) {
  const int gx = cdiv(g.K, 128), gy = cdiv(p->Cout, BN);
  WgBox box;
  memset(&box, 0, sizeof(box));
  const bool win = p->mode == 2;
  bool tma;
  if (win) {
    D3FK_CHECK_ARG(!group && p->c0 == 64 && p->c1 == 0 && p->ld0 == 16 && p->kh == 4 && p->kw == 1 && p->stride == 1 && p->up0 == 0 &&
                   p->Ho == p->Hi && p->Wo == p->Wi && p->cin_real == 3, "wgrad mode 2: the space-to-depth stem geometry");
    d3fk_wgrad_params q = *p;
    q.pad = 0; q.kh = q.kw = 1;          // the window view is a "same" 1x1 geometry for the box test
    tma = wg_tma_box(g, &q, box);
    if (!tma) return set_error(D3FK_ERR_UNSUPPORTED, "wgrad mode 2: the 64-pixel block is not a (w, h, n) box for %d x %d", p->Ho, p->Wo);
    box.win = 1;
  } else {
    tma = !group && wg_tma_box(g, p, box);
  }
  alignas(64) CUtensorMap tmA, tmD;
  memset(&tmA, 0, sizeof(tmA));
  memset(&tmD, 0, sizeof(tmD));
  if (tma) {
    const uint64_t pitch = (uint64_t)(win ? g.Wi + 3 : g.Wi);      // pixels per image row in memory
    uint64_t dims[4] = {(uint64_t)g.c0, (uint64_t)g.Wi, (uint64_t)g.Hi, (uint64_t)g.B};
    uint64_t strides[3] = {(uint64_t)g.ld0 * 2, pitch * g.ld0 * 2, (uint64_t)g.Hi * pitch * g.ld0 * 2};
    uint32_t bx[4] = {64u, (uint32_t)box.bw, (uint32_t)box.bh, (uint32_t)box.bn};
    int rc = get_tensor_map(&tmA, p->src0, 4, dims, strides, bx, 128);
    if (rc) return rc;
    uint64_t ddims[2] = {(uint64_t)p->Cout, (uint64_t)g.M};
    uint64_t dstrides[1] = {(uint64_t)p->ldy * 2};
    uint32_t dbx[2] = {64u, (uint32_t)WG_PIX};
    rc = get_tensor_map(&tmD, p->dy, 2, ddims, dstrides, dbx, 128);
    if (rc) return rc;
  }
  const double stage_us = tma ? 0.2 : 0.4;      // measured main-loop time per 64-pixel stage
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
      if (c == 1) { best = (double)nblk * stage_us * cdiv(tiles, cap); cl = 1; splits = 1; }   // more tiles than one wave: no split
      continue;
    }
    const int cand[2] = {smax, c};
    for (int i = 0; i < 2; ++i) {
      const int sp = cand[i];
      // cluster reduction: (c - 1) / c of each CTA's 128 x BN fp32 tile crosses the SM-to-SM network (~19 B / cycle / SM, two CTAs per SM)
      const double dsmem_us = c > 1 ? 0.5 + (double)(c - 1) / c * (128.0 * BN * 4 * g_wg_ctas_per_sm) / (19.0 * 1965.0) : 0.0;
      // (an epilogue-store term — 128 x BN / c scattered 4-byte accesses per CTA — was tried: it only moves layer4 to a 2-CTA
      // cluster split, 20.3 -> 23.5 us; those kernels are bound by L2 -> SM operand traffic, 512 KB per CTA)
      const double est = (double)cdiv(nblk, sp) * stage_us + (sp > c ? (sp / c) * elems / 216e3 : 0.0) + dsmem_us;
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
  if (g_verbose) fprintf(stderr, "[d3fk] wgrad<%d> M=%d K=%d Cout=%d group=%d tiles=%d cl=%d splits=%d bps=%d tma=%d\n", BN, g.M, g.K, p->Cout, G, tiles, cl, splits, bps, (int)tma);
  const dim3 cluster = (cl == 1 && g_wg_plain) ? dim3(1, 1, 1) : dim3(1, 1, cl);
  cudaError_t le;
  if (tma)
    le = launch_k(wgrad_tc_kernel<BN, true>, grid, dim3(WG_THREADS), smem, s, cluster, g, make_fastdiv((uint32_t)g.Wo),
                  make_fastdiv((uint32_t)g.Ho), (const bf16*)p->dy, p->ldy, p->Cout, p->dw, p->cin_real, p->cout_real, bps,
                  WG_PIX * 128, WG_PIX * 128, cl, g_dev_error_flag, grp, tmA, tmD, box);
  else
    le = launch_k(wgrad_tc_kernel<BN, false>, grid, dim3(WG_THREADS), smem, s, cluster, g, make_fastdiv((uint32_t)g.Wo),
                  make_fastdiv((uint32_t)g.Ho), (const bf16*)p->dy, p->ldy, p->Cout, p->dw, p->cin_real, p->cout_real, bps,
                  WG_PIX * 128, WG_PIX * 128, cl, g_dev_error_flag, grp, tmA, tmD, box);
  if (le != cudaSuccess) return set_error(D3FK_ERR_CUDA, "wgrad_tc launch: %s", cudaGetErrorString(le));
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
  uint32_t coalesce;      // epilogue: 0 = one atomicAdd per lane and element, 1 = staged + coalesced, 2 = staged + 16-byte reductions
  uint32_t ablate;        // -DD3FK_DEBUG builds only (D3FK_WGS_ABLATE): 1 = no atomic epilogue, 2 = no MMAs, 4 = no slab loads, 8 = no dY load
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

  const int tid = threadIdx.x, warp = warp_index_uniform(), lane = tid & 31;
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
  if (warp == 4 && elect_one_sync()) {
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
    if (elect_one_sync()) {
      const int rows = ss.S * ss.R;
      uint32_t it = 0;
      for (int t = blockIdx.x; t < ss.total; t += gridDim.x, ++it) {
        const int n = t / ss.tiles_per_img;
        const int rem = t - n * ss.tiles_per_img;
        const int hb = rem / ss.wtiles;
        const int h0 = hb * rows, w0 = (rem - hb * ss.wtiles) * ss.Wt;
        const int st = it % ss.stages;
        if (it >= (uint32_t)ss.stages) mbar_wait(empty_bar(st), ((it / ss.stages) - 1) & 1, errflag);
        const uint32_t sb = stage_base + st * ss.stage_bytes;
#ifdef D3FK_DEBUG
        if (ss.ablate & 12u) {
          const uint32_t tx = ((ss.ablate & 4u) ? 0u : 3u * ss.slab_tx) + ((ss.ablate & 8u) ? 0u : (uint32_t)ss.dy_bytes);
          if (tx) mbar_arrive_expect_tx(full_bar(st), tx); else mbar_arrive(full_bar(st));
          if (!(ss.ablate & 4u)) for (int sx = 0; sx < 3; ++sx) tma_load_4d(sb + sx * ss.slab_bytes, &tmA, 0, w0 + sx - 1, h0 - 1, n, full_bar(st));
          if (!(ss.ablate & 8u)) tma_load_4d(sb + 3 * ss.slab_bytes, &tmD, 0, w0, h0, n, full_bar(st));
          continue;
        }
#endif
        mbar_arrive_expect_tx(full_bar(st), 3u * ss.slab_tx + (uint32_t)ss.dy_bytes);
        for (int sx = 0; sx < 3; ++sx) tma_load_4d(sb + sx * ss.slab_bytes, &tmA, 0, w0 + sx - 1, h0 - 1, n, full_bar(st));
        tma_load_4d(sb + 3 * ss.slab_bytes, &tmD, 0, w0, h0, n, full_bar(st));
      }
    }
    __syncwarp();
  } else if (warp >= 5) {
    // MMA issuers: warp w owns kw = w % 3 and the K steps (16-pixel groups) of parity w / 3, in its own accumulators —
    // six independent issue streams, each ONE elected thread (elect.sync: descriptors in uniform registers).
    if (has_work && elect_one_sync()) {
      const int w = warp - 5;
      const int sx = w % 3, par = w / 3;
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
#ifdef D3FK_DEBUG
        if (ss.ablate & 2u) { umma_commit(empty_bar(st)); continue; }
#endif
        for (int mb = 0; mb < ss.MB; ++mb) {
          const uint32_t d_addr = d_base + (uint32_t)(mb * ACC);
          uint32_t a_cur = a_lo + (uint32_t)mb * mb_off, b_cur = b_lo;
          umma_f16_lohi(d_addr, a_cur, ahi, b_cur, bhi, idesc, it ? 1u : 0u);
#pragma unroll 4
          for (int j = 1; j < ksteps2; ++j) {
            a_cur += a_step2;
            b_cur += b_step2;
            umma_f16_lohi(d_addr, a_cur, ahi, b_cur, bhi, idesc, 1u);
          }
        }
        umma_commit(empty_bar(st));
      }
      umma_commit(acc_full);
    }
    __syncwarp();
  } else if (has_work) {
    // epilogue (once per CTA): accumulator row r of (kw, row block mb) = tap kh = mb*(128/C) + r / C, channel ci = r % C.
    // The CTA's whole result is first laid out in shared memory (the pipeline stages are free by now) in dW's own order
    // [co][ci][kh][kw], then added to dW with COALESCED reductions (red.global.add.v4.f32 where the alignment allows): a
    // lane-per-row atomicAdd hits 32 different sectors per instruction, and 148 CTAs x 18 k of those cost more than the
    // main loop (ablation, tools/gpu_run57.sh: 288 -> 173 us over the six decoder-tail layers with the atomics removed).
    mbar_wait(acc_full, 0, errflag);
    tc_fence_after();
    const int r = warp * 32 + lane;
    const int apm = 128 / ss.C;
    const int ce = ss.C < cin_real ? ss.C : cin_real;      // channels of this launch that exist in dW
    const int run = ce * 9;                                // contiguous floats of dW per output channel
    float* stg = reinterpret_cast<float*>(smem_raw + (stage_base - smem_u32(smem_raw)));
    for (int sx = 0; sx < 3; ++sx) {
      for (int mb = 0; mb < ss.MB; ++mb) {
        const int kh = mb * apm + r / ss.C, ci = r % ss.C;
        const bool ok = kh < 3 && ci < ce;
#pragma unroll 1
        for (int cc = 0; cc < BN; cc += 16) {
          uint32_t raw[16], raw2[16];
          const uint32_t ta = tmem_d + ((uint32_t)(warp * 32) << 16) + (uint32_t)((sx * ss.MB + mb) * ACC + cc);
          tmem_ld16(ta, raw);
          tmem_ld16(ta + (uint32_t)(nacc * ACC), raw2);     // the odd-K-step accumulator set
          tmem_ld_wait();
#ifdef D3FK_DEBUG
          if (ss.ablate & 1u) continue;
#endif
          if (ok) {
            if (ss.coalesce) {
              float* q = stg + ci * 9 + kh * 3 + sx;
#pragma unroll
              for (int i = 0; i < 16; ++i)
                if (cc + i < cout_real) q[(cc + i) * run] = __uint_as_float(raw[i]) + __uint_as_float(raw2[i]);
            } else {
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
    if (ss.coalesce) {
      asm volatile("bar.sync 1, 128;" ::: "memory");        // the four epilogue warps only
      const int total = cout_real * run;
      const long long gstride = (long long)cin_real * 9;
#ifdef D3FK_DEBUG
      if (ss.ablate & 1u) { /* timing experiment: staging only */ } else
#endif
      if (ss.coalesce == 2) {                                // 16-byte aligned rows: vector reductions
        for (int j = tid * 4; j < total; j += 128 * 4) {
          const int co = j / run, o = j - co * run;
          const float4 v = *reinterpret_cast<const float4*>(stg + j);
          asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dw + co * gstride + o), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
        }
      } else {
        for (int j = tid; j < total; j += 128) {
          const int co = j / run, o = j - co * run;
          atomicAdd(dw + co * gstride + o, stg[j]);
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 5) tmem_dealloc(tmem_d, tmem_cols);
}

static int g_use_wg_slab = 1;   // D3FK_WG_SLAB=0: never take the slab weight-gradient path
static int g_wgs_smax = 4;      // D3FK_WGS_SMAX (debug builds): largest super-tile (sub-tiles of 128 pixels) of the slab weight gradient
static int g_wgs_stages = 3;    // D3FK_WGS_STAGES (debug builds): deepest pipeline tried (<= 4)
static int g_wgs_coalesce = 2;  // D3FK_WGS_COALESCE (debug builds): 0 = per-lane atomics, 1 = staged scalar, 2 = staged 16-byte reductions
static int g_wgs_grid = 0;      // D3FK_WGS_GRID (debug builds): cap of the persistent grid (SMs left to the main chain)
static int g_wgs_ablate = 0;    // D3FK_WGS_ABLATE (debug builds): see WgSlabSched::ablate

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
  ss.ablate = (uint32_t)g_wgs_ablate;
  const int ACC = BN < 32 ? 32 : BN;
  if (2 * 3 * ss.MB * ACC > 512) return 0;
  int smem = 0;
  bool found = false;
  for (int S = g_wgs_smax; S >= 1 && !found; S >>= 1) {
    if (H % (S * R)) continue;
    const int slab = ((S * R + 2) * Wt * ss.a_row_bytes + 1023) & ~1023;
    const int dyb = (S * 128 * ss.b_row_bytes + 1023) & ~1023;
    // surplus M atoms (Cin < 64: 128/C - 3 of them) read up to (128/C - 3) image rows past the last slab: they must stay
    // inside the stage (the dY tile that follows the slabs absorbs them)
    const int overrun = (128 / C > 3 ? 128 / C - 3 : (C == 64 ? 1 : 0)) * Wt * ss.a_row_bytes;
    if (overrun > dyb) continue;
    for (int stages = g_wgs_stages; stages >= 2; --stages) {
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
  {
    // staged epilogue: the CTA's [cout_real][min(C, cin_real) * 9] result must fit the (then idle) pipeline stages
    const int ce = C < p->cin_real ? C : p->cin_real;
    const size_t need = (size_t)p->cout_real * ce * 9 * sizeof(float);
    ss.coalesce = 0;
    if (g_wgs_coalesce && need <= (size_t)ss.stages * ss.stage_bytes)
      ss.coalesce = (g_wgs_coalesce > 1 && ce % 4 == 0 && p->cin_real % 4 == 0 && ((uintptr_t)p->dw & 15) == 0) ? 2u : 1u;
  }
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
  if (g_wgs_grid > 0 && grid > g_wgs_grid) grid = g_wgs_grid;
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
  if (p->mode != 2) {
    const int slab = try_launch_wgrad_slab(g, p, s);
    if (slab) return slab < 0 ? slab : D3FK_OK;
  }
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


int wgrad_init() {
  cudaError_t e = cudaSuccess;
#ifdef D3FK_DEBUG
  if (const char* v = getenv("D3FK_WG_SLAB")) g_use_wg_slab = atoi(v);
  if (const char* v = getenv("D3FK_WGS_SMAX")) g_wgs_smax = atoi(v);
  if (const char* v = getenv("D3FK_WGS_ABLATE")) g_wgs_ablate = atoi(v);
  if (const char* v = getenv("D3FK_WGS_GRID")) g_wgs_grid = atoi(v);
  if (const char* v = getenv("D3FK_WGS_COALESCE")) g_wgs_coalesce = atoi(v);
  if (const char* v = getenv("D3FK_WGS_STAGES")) { g_wgs_stages = atoi(v); if (g_wgs_stages > 4) g_wgs_stages = 4; if (g_wgs_stages < 2) g_wgs_stages = 2; }
  if (const char* v = getenv("D3FK_WG_OCC")) g_wg_ctas_per_sm = atoi(v);
  if (const char* v = getenv("D3FK_WG_TMA")) g_wg_tma = atoi(v);
  if (const char* v = getenv("D3FK_WG_CAP")) g_wg_cap = atoi(v);
  if (const char* v = getenv("D3FK_WG_PLAIN")) g_wg_plain = atoi(v);
  if (const char* v = getenv("D3FK_WG_GROUP_OCC")) g_wg_group_occ = atoi(v);
#endif
  D3FK_SET_SMEM(wgrad_slab_kernel<16>, SLAB_MAX_SMEM)
  D3FK_SET_SMEM(wgrad_slab_kernel<32>, SLAB_MAX_SMEM)
  D3FK_SET_SMEM((wgrad_tc_kernel<64, false>), WG_ONE_PER_SM_SMEM)
  D3FK_SET_SMEM((wgrad_tc_kernel<128, false>), WG_ONE_PER_SM_SMEM)
  D3FK_SET_SMEM((wgrad_tc_kernel<64, true>), WG_ONE_PER_SM_SMEM)
  D3FK_SET_SMEM((wgrad_tc_kernel<128, true>), WG_ONE_PER_SM_SMEM)
  if (e != cudaSuccess) return set_error(D3FK_ERR_CUDA, "cudaFuncSetAttribute (wgrad): %s", cudaGetErrorString(e));
  return D3FK_OK;
}

}  // namespace d3fk
