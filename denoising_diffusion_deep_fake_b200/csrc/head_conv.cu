// head_conv.cu — the segmentation head: 3x3 / stride 1 / pad 1 convolution from 16 channels to <= 4 (the network's 3 RGB
// outputs) with bias, NHWC bf16 in, fp32 NCHW out.  Replaces smp's SegmentationHead Conv2d(16, 3, 3, padding=1)
// (d3f/train_denoiser/lit_module.py:46-52 builds it through smp.Unet).
//
// Why not the tensor-core slab kernel: N = 3 pads to a 16-wide MMA (81 % of the tensor work is padding) and, worse, the
// kernel is bound by the TMA engine's box-row rate — three column-shifted slabs of 32-byte pixel rows, 58 us for 1 M pixels
// whatever N is.  The layer is 0.45 GFMA against 46 MB of compulsory HBM traffic: CUDA cores do it at
// the memory floor's order of magnitude.  A 256-thread block stages a (TH+2) x (TW+2) halo tile of the input with plain
// 16-byte loads (2 KB-contiguous image rows, no TMA), every thread produces 4 pixels of one row x 3 channels from
// packed shared-memory words (lanes = consecutive pixels, conflict-free), weights broadcast from shared memory as float4
// (co0, co1, co2, 0) per (tap, ci).
#include "common.cuh"

namespace d3fk {

constexpr int HC_C = 16;            // input channels
constexpr int HC_PX = 4;            // pixels per thread (along W)
constexpr int HC_THREADS = 256;

template <int TW>
__global__ void __launch_bounds__(HC_THREADS) head_conv_kernel(const __nv_bfloat16* __restrict__ x, int ldx, const __nv_bfloat16* __restrict__ w,
                                                              const float* __restrict__ bias, float* __restrict__ out, int B, int H,
                                                              int W, int Cout) {
  constexpr int TH = HC_THREADS * HC_PX / TW;          // 8 / 16 / 32 rows for TW = 128 / 64 / 32
  constexpr int SW = TW + 2, SH = TH + 2;
  extern __shared__ uint4 hc_smem[];
  uint4* tile = hc_smem;                                // [2 halves][SH][SW] : channels 0-7 / 8-15 of every staged pixel
  float4* wsm = reinterpret_cast<float4*>(tile + SH * SW * 2);   // [9][16] (co0, co1, co2, co3)
  const int tid = threadIdx.x;
  // weights: w[co][tap][ci] bf16 (rows = Cout) -> float4 per (tap, ci); independent of the previous kernel
  for (int i = tid; i < 9 * HC_C; i += HC_THREADS) {
    float v[4] = {0.f, 0.f, 0.f, 0.f};
    for (int co = 0; co < Cout; ++co) v[co] = __bfloat162float(w[co * 9 * HC_C + i]);
    wsm[i] = make_float4(v[0], v[1], v[2], v[3]);
  }
  pdl_enter();
  const int tiles_w = W / TW, tiles_h = H / TH;
  int b = blockIdx.x;
  const int tw = b % tiles_w; b /= tiles_w;
  const int th = b % tiles_h; b /= tiles_h;             // b = image
  const int h0 = th * TH - 1, w0 = tw * TW - 1;         // image coordinates of tile(0, 0)
  // stage: all of a thread's loads are independent (unrolled batches); the two 16-byte halves of a pixel go to separate
  // planes so that consecutive lanes (consecutive pixels) read consecutive 16-byte words — no bank conflicts
  constexpr int NV = SH * SW * 2;
  constexpr int PER = (NV + HC_THREADS - 1) / HC_THREADS;
#pragma unroll 4
  for (int k = 0; k < PER; ++k) {
    const int i = tid + k * HC_THREADS;
    if (i < NV) {
      const int half = i & 1, pix = i >> 1;
      const int r = pix / SW, c = pix - r * SW;
      const int gh = h0 + r, gw = w0 + c;
      uint4 v = make_uint4(0u, 0u, 0u, 0u);
      if ((unsigned)gh < (unsigned)H && (unsigned)gw < (unsigned)W)
        v = __ldg(reinterpret_cast<const uint4*>(x + ((long long)(b * H + gh) * W + gw) * ldx) + half);
      tile[half * SH * SW + pix] = v;
    }
  }
  __syncthreads();
  // a thread owns HC_PX pixels of one row, TW / HC_PX columns apart: the lanes of a warp are consecutive pixels
  constexpr int CG = TW / HC_PX;
  const int row = tid / CG, cg = tid - row * CG;
  float acc[HC_PX][3];
#pragma unroll
  for (int j = 0; j < HC_PX; ++j) { acc[j][0] = 0.f; acc[j][1] = 0.f; acc[j][2] = 0.f; }
#pragma unroll
  for (int ky = 0; ky < 3; ++ky) {
#pragma unroll
    for (int kx = 0; kx < 3; ++kx) {
      uint4 px[HC_PX][2];
#pragma unroll
      for (int j = 0; j < HC_PX; ++j) {
        const int o = (row + ky) * SW + cg + j * CG + kx;
        px[j][0] = tile[o];
        px[j][1] = tile[SH * SW + o];
      }
#pragma unroll
      for (int ci = 0; ci < HC_C; ++ci) {
        const float4 wv = wsm[(ky * 3 + kx) * HC_C + ci];
#pragma unroll
        for (int j = 0; j < HC_PX; ++j) {
          const uint4& q = px[j][ci >> 3];
          const uint32_t word = (ci & 7) < 2 ? q.x : (ci & 7) < 4 ? q.y : (ci & 7) < 6 ? q.z : q.w;
          const float v = __uint_as_float((ci & 1) ? (word & 0xFFFF0000u) : (word << 16));   // bf16 -> fp32
          acc[j][0] = fmaf(v, wv.x, acc[j][0]);
          acc[j][1] = fmaf(v, wv.y, acc[j][1]);
          acc[j][2] = fmaf(v, wv.z, acc[j][2]);
        }
      }
    }
  }
  const int oh = th * TH + row;
  for (int co = 0; co < Cout; ++co) {
    const float bv = bias ? __ldg(bias + co) : 0.f;
    float* orow = out + ((long long)(b * Cout + co) * H + oh) * W + tw * TW + cg;
#pragma unroll
    for (int j = 0; j < HC_PX; ++j) orow[j * CG] = acc[j][co < 3 ? co : 0] + bv;      // lanes = consecutive pixels: coalesced
  }
}

// The same layer on the warp-level tensor-core path: mma.sync.m16n8k16 (bf16 x bf16 -> fp32) fed by ldmatrix from the SAME
// two-plane halo tile.  A warp produces 16 consecutive pixels of a row x 8 output columns (3 real) per group: per tap ONE
// ldmatrix.x4 (the four 8x8 blocks = pixels 0-7 / 8-15 of channel plane 0 / 1: 128 contiguous bytes each, conflict-free) and ONE
// mma — 18 instructions per 16 pixels and tap set, against ~300 for the FFMA form above (which is issue-bound: IPC 2.5, 34 us
// for 1 M pixels against a 7 us HBM floor).  The B fragments of all 9 taps (W[n][tap][k]) live in 18 registers per thread.
__device__ __forceinline__ void hc_ldmatrix_x4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void hc_mma_bf16(float* c, uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

template <int TW>
__global__ void __launch_bounds__(HC_THREADS) head_conv_mma_kernel(const __nv_bfloat16* __restrict__ x, int ldx, const __nv_bfloat16* __restrict__ w,
                                                                  const float* __restrict__ bias, float* __restrict__ out, int B, int H,
                                                                  int W, int Cout) {
  constexpr int TH = HC_THREADS * HC_PX / TW;          // 8 / 16 / 32 rows for TW = 128 / 64 / 32 (1024 pixels per block)
  constexpr int SW = TW + 2, SH = TH + 2;
  extern __shared__ uint4 hc_smem[];
  uint4* tile = hc_smem;                                // [2 halves][SH][SW] : channels 0-7 / 8-15 of every staged pixel
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, t = lane & 3;
  // B fragments (col-major 16 x 8): b0 = W[n = g][tap][k = 2t, 2t+1], b1 = W[g][tap][2t+8, 2t+9]; columns >= Cout are zero.
  // Weights and bias do not depend on the previous kernel: loaded before griddepcontrol.wait.
  uint32_t bf[9][2];
#pragma unroll
  for (int tap = 0; tap < 9; ++tap) {
    bf[tap][0] = 0u;
    bf[tap][1] = 0u;
    if (g < Cout) {
      const uint32_t* wr = reinterpret_cast<const uint32_t*>(w + g * 9 * HC_C + tap * HC_C);
      bf[tap][0] = __ldg(wr + t);
      bf[tap][1] = __ldg(wr + t + 4);
    }
  }
  const float bv0 = (bias && 2 * t < Cout) ? __ldg(bias + 2 * t) : 0.f;
  const float bv1 = (bias && 2 * t + 1 < Cout) ? __ldg(bias + 2 * t + 1) : 0.f;
  pdl_enter();
  const int tiles_w = W / TW, tiles_h = H / TH;
  int b = blockIdx.x;
  const int tw = b % tiles_w; b /= tiles_w;
  const int th = b % tiles_h; b /= tiles_h;             // b = image
  const int h0 = th * TH - 1, w0 = tw * TW - 1;         // image coordinates of tile(0, 0)
  constexpr int NV = SH * SW * 2;
  constexpr int PER = (NV + HC_THREADS - 1) / HC_THREADS;
#pragma unroll 4
  for (int k = 0; k < PER; ++k) {
    const int i = tid + k * HC_THREADS;
    if (i < NV) {
      const int half = i & 1, pix = i >> 1;
      const int r = pix / SW, c = pix - r * SW;
      const int gh = h0 + r, gw = w0 + c;
      uint4 v = make_uint4(0u, 0u, 0u, 0u);
      if ((unsigned)gh < (unsigned)H && (unsigned)gw < (unsigned)W)
        v = __ldg(reinterpret_cast<const uint4*>(x + ((long long)(b * H + gh) * W + gw) * ldx) + half);
      tile[half * SH * SW + pix] = v;
    }
  }
  __syncthreads();
  constexpr int GW = TW / 16;                           // 16-pixel groups per tile row
  constexpr int NG = TH * GW;                           // 64 groups per block, 8 per warp
  const uint32_t tile_s = (uint32_t)__cvta_generic_to_shared(tile);
  // ldmatrix.x4: lanes 0-7 / 8-15 / 16-23 / 24-31 give the row addresses of (pixels 0-7, plane 0) / (8-15, plane 0) /
  // (0-7, plane 1) / (8-15, plane 1) = the a0 / a1 / a2 / a3 blocks of the 16 x 16 A fragment
  const int lrow = (lane >> 4) * SH * SW + ((lane >> 3) & 1) * 8 + (lane & 7);
  const long long plane = (long long)H * W;
  for (int grp = warp; grp < NG; grp += HC_THREADS / 32) {
    const int r = grp / GW, c0 = (grp - r * GW) * 16;
    float c[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int ky = 0; ky < 3; ++ky) {
#pragma unroll
      for (int kx = 0; kx < 3; ++kx) {
        uint32_t a0, a1, a2, a3;
        hc_ldmatrix_x4(tile_s + (uint32_t)(((r + ky) * SW + c0 + kx + lrow) * 16), a0, a1, a2, a3);
        hc_mma_bf16(c, a0, a1, a2, a3, bf[ky * 3 + kx][0], bf[ky * 3 + kx][1]);
      }
    }
    // C fragment: c0 / c1 = (pixel g, channels 2t / 2t+1), c2 / c3 = (pixel g + 8, same channels); fp32 NCHW out
    float* o = out + (long long)b * Cout * plane + (long long)(th * TH + r) * W + tw * TW + c0 + g;
    if (2 * t < Cout) {
      o[(2 * t) * plane] = c[0] + bv0;
      o[(2 * t) * plane + 8] = c[2] + bv0;
    }
    if (2 * t + 1 < Cout) {
      o[(2 * t + 1) * plane] = c[1] + bv1;
      o[(2 * t + 1) * plane + 8] = c[3] + bv1;
    }
  }
}

static int g_head_mma = 1;   // D3FK_HEAD_MMA=0 (debug builds): the FFMA form

template <int TW>
static int launch_head_tw(const d3fk_conv_params* p, cudaStream_t s) {
  constexpr int TH = HC_THREADS * HC_PX / TW;
  const size_t smem = (size_t)(TH + 2) * (TW + 2) * 32 + 9 * HC_C * sizeof(float4);
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(head_conv_kernel<TW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(head_conv_mma_kernel<TW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return set_error(D3FK_ERR_CUDA, "head conv smem attribute: %s", cudaGetErrorString(e));
    attr_set = true;
  }
  const long long blocks = (long long)p->B * (p->Hi / TH) * (p->Wi / TW);
  if (g_head_mma && ((uintptr_t)p->w & 3) == 0)
    launch_k(head_conv_mma_kernel<TW>, dim3((unsigned)blocks), dim3(HC_THREADS), smem - 9 * HC_C * sizeof(float4), s, dim3(1, 1, 1),
             (const __nv_bfloat16*)p->src0, p->ld0, (const __nv_bfloat16*)p->w, p->shift, p->out_nchw, p->B, p->Hi, p->Wi, p->Cout);
  else
    launch_k(head_conv_kernel<TW>, dim3((unsigned)blocks), dim3(HC_THREADS), smem, s, dim3(1, 1, 1), (const __nv_bfloat16*)p->src0, p->ld0,
             (const __nv_bfloat16*)p->w, p->shift, p->out_nchw, p->B, p->Hi, p->Wi, p->Cout);
  count_launch();
  return check_launch("head_conv");
}

// 1 = launched, 0 = not this kernel's shape (the caller falls through to the tensor-core paths), < 0 = error
int try_launch_head_conv(const d3fk_conv_params* p, cudaStream_t s) {
#ifdef D3FK_DEBUG
  static int enabled = -1;
  if (enabled < 0) { const char* v = getenv("D3FK_HEAD_CONV"); enabled = v ? atoi(v) : 1; if (const char* m = getenv("D3FK_HEAD_MMA")) g_head_mma = atoi(m); }
  if (!enabled) return 0;
#endif
  if (p->dtype != D3FK_BF16 || p->mode != 0 || p->kh != 3 || p->kw != 3 || p->stride != 1 || p->pad != 1) return 0;
  if (p->c0 != HC_C || p->c1 != 0 || p->up0 != 0 || p->src1 || p->Cout < 1 || p->Cout > 3) return 0;
  if (!p->out_nchw || p->out || p->res || p->stats || p->relu || p->scale || p->bw_x) return 0;
  if (p->Ho != p->Hi || p->Wo != p->Wi || (p->Hi % 32) || (p->Wi % 32) || (p->ld0 % 8)) return 0;
  if (((uintptr_t)p->src0 & 15) || ((uintptr_t)p->out_nchw & 15)) return 0;
  const long long blocks32 = (long long)p->B * (p->Hi / 32) * (p->Wi / 32);
  if (blocks32 >= (1ll << 31)) return 0;
  int rc;
  if (p->Wi % 128 == 0) rc = launch_head_tw<128>(p, s);
  else if (p->Wi % 64 == 0) rc = launch_head_tw<64>(p, s);
  else rc = launch_head_tw<32>(p, s);
  return rc ? rc : 1;
}

}  // namespace d3fk
