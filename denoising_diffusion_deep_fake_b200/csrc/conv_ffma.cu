// conv_ffma.cu — fp32 parity-mode convolution (CUDA-core FFMA implicit GEMM), its weight
// gradient, and the weight packing kernels shared with the bf16 tcgen05 engine.
// The fp32 engine exists for the 1e-5 parity mode named by BASELINE.json's north_star; the
// throughput path is conv_tc.cu.
#include "common.cuh"

namespace d3fk {

int make_gather(Gather& g, const void* src0, const void* src1, int c0, int c1, int ld0, int ld1, int up0, int B, int Hi,
                int Wi, int Ho, int Wo, int kh, int kw, int stride, int pad, int mode) {
  g.src0 = src0; g.src1 = src1; g.c0 = c0; g.c1 = c1; g.ld0 = ld0; g.ld1 = ld1; g.up0 = up0;
  g.B = B; g.Hi = Hi; g.Wi = Wi; g.Ho = Ho; g.Wo = Wo; g.kh = kh; g.kw = kw; g.stride = stride; g.pad = pad; g.mode = mode;
  g.ctot = c0 + c1;
  g.K = kh * kw * g.ctot;
  g.M = B * Ho * Wo;
  g.sshift = stride == 1 ? 0 : (stride == 2 ? 1 : -1);
  if (g.sshift < 0) return set_error(D3FK_ERR_UNSUPPORTED, "stride must be 1 or 2");
  if (c0 % 8 || c1 % 8 || c0 <= 0 || c1 < 0) return set_error(D3FK_ERR_ARG, "c0/c1 must be multiples of 8");
  if (ld0 % 8 || (c1 && ld1 % 8)) return set_error(D3FK_ERR_ARG, "pixel strides must be multiples of 8");
  if (c1 && !src1) return set_error(D3FK_ERR_ARG, "src1 missing");
  if (up0 && ((Hi | Wi) & 1)) return set_error(D3FK_ERR_ARG, "upsampled extent must be even");
  if ((long long)B * Ho * Wo >= (1ll << 31)) return set_error(D3FK_ERR_UNSUPPORTED, "M too large");
  return D3FK_OK;
}

// ---------------------------------------------------------------------------------------------
// shared epilogue for one output element
struct Epi {
  void* out; float* out_nchw; const float* scale; const float* shift; const void* res; double* stats;
  int ldo, ldr, relu, Cout, Ho, Wo;
};

// ---------------------------------------------------------------------------------------------
// forward / dgrad conv: 64x64 tile, BK=16, 256 threads, 4x4 micro-tile
constexpr int FBM = 64, FBN = 64, FBK = 16;

__global__ void __launch_bounds__(256) conv_ffma_kernel(Gather g, const float* __restrict__ w, Epi e) {
  pdl_enter();
  __shared__ float As[FBK][FBM + 4];
  __shared__ float Bs[FBK][FBN + 4];
  __shared__ double sstat[2][FBN];
  const int tid = threadIdx.x;
  const int m0 = blockIdx.x * FBM, n0 = blockIdx.y * FBN;
  const int tx = tid % 16, ty = tid / 16;
  if (tid < FBN) { sstat[0][tid] = 0.0; sstat[1][tid] = 0.0; }

  // loader mapping: one float4 of A and one of B per thread per k-block
  const int lrow = tid / 4, lk = (tid % 4) * 4;
  const int am = m0 + lrow;
  const bool arow_ok = am < g.M;
  int an = 0, ah0 = 0, aw0 = 0;
  if (arow_ok) {
    int wo = am % g.Wo;
    int t = am / g.Wo;
    int ho = t % g.Ho;
    an = t / g.Ho;
    if (g.mode == 0) { ah0 = ho * g.stride - g.pad; aw0 = wo * g.stride - g.pad; }
    else { ah0 = ho + g.pad; aw0 = wo + g.pad; }
  }
  const int bn = n0 + lrow;
  const bool brow_ok = bn < e.Cout;

  // parity mode: each 16-deep chunk is summed in fp32 and folded into a double accumulator, so the
  // result is the correctly rounded fp32 value for all practical purposes
  double acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.0;

  for (int kb = 0; kb < g.K; kb += FBK) {
    int k = kb + lk;
    float4 av = make_float4(0.f, 0.f, 0.f, 0.f), bv = make_float4(0.f, 0.f, 0.f, 0.f);
    if (k < g.K) {
      if (arow_ok) {
        int tap = k / g.ctot;
        int c = k - tap * g.ctot;
        int khi = tap / g.kw, kwi = tap - khi * g.kw;
        int which;
        long long off = gather_offset(g, an, ah0, aw0, khi, kwi, c, which);
        if (off >= 0) av = __ldg(reinterpret_cast<const float4*>((which ? (const float*)g.src1 : (const float*)g.src0) + off));
      }
      if (brow_ok) bv = __ldg(reinterpret_cast<const float4*>(w + (long long)bn * g.K + k));
    }
    As[lk + 0][lrow] = av.x; As[lk + 1][lrow] = av.y; As[lk + 2][lrow] = av.z; As[lk + 3][lrow] = av.w;
    Bs[lk + 0][lrow] = bv.x; Bs[lk + 1][lrow] = bv.y; Bs[lk + 2][lrow] = bv.z; Bs[lk + 3][lrow] = bv.w;
    __syncthreads();
    float part[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) part[i][j] = 0.f;
#pragma unroll
    for (int kk = 0; kk < FBK; ++kk) {
      float a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = As[kk][ty * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = Bs[kk][tx * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) part[i][j] = fmaf(a[i], b[j], part[i][j]);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j] += (double)part[i][j];
    __syncthreads();
  }

  // epilogue
  double cs[4] = {0, 0, 0, 0}, cq[4] = {0, 0, 0, 0};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int m = m0 + ty * 4 + i;
    if (m >= g.M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int c = n0 + tx * 4 + j;
      if (c >= e.Cout) continue;
      float v = (float)acc[i][j];
      if (e.scale) v = fmaf(v, e.scale[c], e.shift ? e.shift[c] : 0.f);
      else if (e.shift) v += e.shift[c];
      if (e.res) v += ((const float*)e.res)[(long long)m * e.ldr + c];
      if (e.relu) v = fmaxf(v, 0.f);
      cs[j] += (double)v;
      cq[j] += (double)v * (double)v;
      if (e.out_nchw) {
        int wo = m % e.Wo;
        int t = m / e.Wo;
        int ho = t % e.Ho;
        int n = t / e.Ho;
        e.out_nchw[(((long long)n * e.Cout + c) * e.Ho + ho) * e.Wo + wo] = v;
      } else {
        ((float*)e.out)[(long long)m * e.ldo + c] = v;
      }
    }
  }
  if (e.stats) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      atomicAdd(&sstat[0][tx * 4 + j], cs[j]);
      atomicAdd(&sstat[1][tx * 4 + j], cq[j]);
    }
    __syncthreads();
    if (tid < FBN && n0 + tid < e.Cout) {
      atomicAdd(&e.stats[n0 + tid], sstat[0][tid]);
      atomicAdd(&e.stats[e.Cout + n0 + tid], sstat[1][tid]);
    }
  }
}

int launch_conv_ffma(const d3fk_conv_params* p, cudaStream_t s) {
  Gather g;
  int rc = make_gather(g, p->src0, p->src1, p->c0, p->c1, p->ld0, p->ld1, p->up0, p->B, p->Hi, p->Wi, p->Ho, p->Wo, p->kh,
                       p->kw, p->stride, p->pad, p->mode);
  if (rc) return rc;
  D3FK_CHECK_ARG(p->out || p->out_nchw, "no output");
  Epi e{p->out, p->out_nchw, p->scale, p->shift, p->res, p->stats, p->ldo, p->ldr, p->relu, p->Cout, p->Ho, p->Wo};
  dim3 grid(cdiv(g.M, FBM), cdiv(p->Cout, FBN));
  launch_k(conv_ffma_kernel, dim3(grid), dim3(256), 0, s, dim3(1, 1, 1), g, (const float*)p->w, e);
  count_launch();
  return check_launch("conv_ffma");
}

// ---------------------------------------------------------------------------------------------
// wgrad: dw[co][ci][kh][kw] += sum_m dy[m][co] * A[m][k]; tile 64 (co) x 64 (k), split over m
__global__ void __launch_bounds__(256) wgrad_ffma_kernel(Gather g, const float* __restrict__ dy, int ldy, int Cout,
                                                         float* __restrict__ dw, int cin_real, int cout_real, int m_per_split) {
  pdl_enter();
  __shared__ float Ys[FBK][FBM + 4];  // [m][co]
  __shared__ float As[FBK][FBN + 4];  // [m][k]
  const int tid = threadIdx.x;
  const int co0 = blockIdx.x * 64, k0 = blockIdx.y * 64;
  const int tx = tid % 16, ty = tid / 16;
  const int mbeg = blockIdx.z * m_per_split;
  const int mend = min(g.M, mbeg + m_per_split);
  // loader: 16 rows x 64 cols = 256 float4 => one per thread for each operand
  const int lm = tid / 16, lc = (tid % 16) * 4;
  // this thread's k (fixed for the whole kernel)
  const int k = k0 + lc;
  int khi = 0, kwi = 0, kc = 0;
  const bool k_ok = k < g.K;
  if (k_ok) {
    int tap = k / g.ctot;
    kc = k - tap * g.ctot;
    khi = tap / g.kw;
    kwi = tap - khi * g.kw;
  }
  double acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.0;

  for (int mb = mbeg; mb < mend; mb += FBK) {
    int m = mb + lm;
    float4 yv = make_float4(0.f, 0.f, 0.f, 0.f), av = make_float4(0.f, 0.f, 0.f, 0.f);
    if (m < mend) {
      if (co0 + lc < Cout) yv = __ldg(reinterpret_cast<const float4*>(dy + (long long)m * ldy + co0 + lc));
      if (k_ok) {
        int wo = m % g.Wo;
        int t = m / g.Wo;
        int ho = t % g.Ho;
        int n = t / g.Ho;
        int which;
        long long off = gather_offset(g, n, ho * g.stride - g.pad, wo * g.stride - g.pad, khi, kwi, kc, which);
        if (off >= 0) av = __ldg(reinterpret_cast<const float4*>((which ? (const float*)g.src1 : (const float*)g.src0) + off));
      }
    }
    *reinterpret_cast<float4*>(&Ys[lm][lc]) = yv;
    *reinterpret_cast<float4*>(&As[lm][lc]) = av;
    __syncthreads();
    float part[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) part[i][j] = 0.f;
#pragma unroll
    for (int mm = 0; mm < FBK; ++mm) {
      float a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = Ys[mm][ty * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = As[mm][tx * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) part[i][j] = fmaf(a[i], b[j], part[i][j]);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j] += (double)part[i][j];
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int co = co0 + ty * 4 + i;
    if (co >= cout_real) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int kk = k0 + tx * 4 + j;
      if (kk >= g.K) continue;
      int tap = kk / g.ctot;
      int ci = kk - tap * g.ctot;
      if (ci >= cin_real) continue;
      atomicAdd(dw + ((long long)co * cin_real + ci) * (g.kh * g.kw) + tap, (float)acc[i][j]);
    }
  }
}

int launch_wgrad_ffma(const d3fk_wgrad_params* p, cudaStream_t s) {
  Gather g;
  int rc = make_gather(g, p->src0, p->src1, p->c0, p->c1, p->ld0, p->ld1, p->up0, p->B, p->Hi, p->Wi, p->Ho, p->Wo, p->kh,
                       p->kw, p->stride, p->pad, 0);
  if (rc) return rc;
  D3FK_CHECK_ARG(p->Cout % 4 == 0 && p->ldy % 4 == 0, "Cout and ldy must be multiples of 4");
  int tiles = cdiv(p->Cout, 64) * cdiv(g.K, 64);
  int splits = max(1, min(cdiv(g.M, 256), cdiv(148 * 4, tiles)));
  int m_per_split = cdiv(cdiv(g.M, splits), FBK) * FBK;
  splits = cdiv(g.M, m_per_split);
  dim3 grid(cdiv(p->Cout, 64), cdiv(g.K, 64), splits);
  launch_k(wgrad_ffma_kernel, dim3(grid), dim3(256), 0, s, dim3(1, 1, 1), g, (const float*)p->dy, p->ldy, p->Cout, p->dw, p->cin_real, p->cout_real, m_per_split);
  count_launch();
  return check_launch("wgrad_ffma");
}

// ---------------------------------------------------------------------------------------------
// weight packing: OIHW fp32 -> [Cout][kh][kw][cin_pad] and/or [Cin][kh][kw][cout_pad]
template <typename T>
__global__ void pack_weights_kernel(d3fk_pack_params p) {
  pdl_enter();
  const int taps = p.kh * p.kw;
  if (p.w_fwd) {
    long long total = (long long)p.Cout * taps * p.cin_pad;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
      int ci = (int)(i % p.cin_pad);
      long long t = i / p.cin_pad;
      int tap = (int)(t % taps);
      int co = (int)(t / taps);
      float v = ci < p.Cin ? p.w[((long long)co * p.Cin + ci) * taps + tap] : 0.f;
      ((T*)p.w_fwd)[i] = from_f<T>(v);
    }
  }
  if (p.w_dgrad) {
    long long total = (long long)p.Cin * taps * p.cout_pad;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
      int co = (int)(i % p.cout_pad);
      long long t = i / p.cout_pad;
      int tap = (int)(t % taps);
      int ci = (int)(t / taps);
      float v = co < p.Cout ? p.w[((long long)co * p.Cin + ci) * taps + tap] : 0.f;
      ((T*)p.w_dgrad)[i] = from_f<T>(v);
    }
  }
}

int launch_pack(const d3fk_pack_params* p, cudaStream_t s) {
  long long total = (long long)p->Cout * p->kh * p->kw * (p->cin_pad > p->Cin ? p->cin_pad : p->Cin);
  if (p->w_dgrad) {
    long long t2 = (long long)p->Cin * p->kh * p->kw * p->cout_pad;
    if (t2 > total) total = t2;
  }
  int grid = (int)((total + 255) / 256);
  if (grid > 148 * 8) grid = 148 * 8;
  if (p->dtype == D3FK_F32) launch_k(pack_weights_kernel<float>, dim3(grid), dim3(256), 0, s, dim3(1, 1, 1), *p);
  else if (p->dtype == D3FK_BF16) launch_k(pack_weights_kernel<__nv_bfloat16>, dim3(grid), dim3(256), 0, s, dim3(1, 1, 1), *p);
  else return set_error(D3FK_ERR_ARG, "bad dtype");
  count_launch();
  return check_launch("pack_weights");
}

}  // namespace d3fk
