// Generic CUDA-core implicit-GEMM convolution (fprop / dgrad / wgrad), fp32, any shape.
//
// This is the path for the shapes the tensor-core family (conv_tc.cu) does not take — the 7x7
// stem on 3/6-channel NCHW images (resnet_encoder.py:92-93), the C->1 disparity heads
// (depth_decoder.py:69-70), the 12-channel pose output (pose_decoder.py:33) — and the
// cross-check for the tcgen05 kernels in the GPU tests.  One "virtual input" fetch implements
// every input-side fusion of the decoder: channel concat of two sources, nearest x2 upsampling of
// source 1 (depth_decoder.py:93-100), reflection / zero padding (monodepth_layers.py:131-134),
// and the NCHW image + (x-0.45)/0.225 normalisation of the stem.
#include "common.cuh"

namespace segsde {

struct ConvP {
  View x1, x2, y;          // x1/x2: sources (strided NHWC, or NCHW planar when nchw); y: output / dy
  const float* w;          // [Cout][kh][kw][Ctot]
  const float* bias;
  int C1, C2, Ctot, Cout;
  int Hc, Wc;              // conv-input domain (after upsampling source 1)
  int Ho, Wo;
  int kh, kw, stride, pad, dil, pad_mode, up1, act, nchw;
  int Ktot;                // kh*kw*Ctot
  long long P;             // N*Ho*Wo
};

// value of the virtual conv input at (n, h, w, ci), h/w already resolved to inside the domain
__device__ __forceinline__ float fetch_in(const ConvP& p, int n, int h, int w, int ci) {
  if (p.nchw) {
    const float* s; int c, Cs;
    if (ci < p.C1) { s = p.x1.p; c = ci; Cs = p.C1; } else { s = p.x2.p; c = ci - p.C1; Cs = p.C2; }
    const float v = __ldg(s + (((long long)n * Cs + c) * p.Hc + h) * p.Wc + w);
    return (v - 0.45f) / 0.225f;
  }
  if (ci < p.C1) {
    if (p.up1) { h >>= 1; w >>= 1; }
    return __ldg(p.x1.p + p.x1.off(n, h, w) + ci);
  }
  return __ldg(p.x2.p + p.x2.off(n, h, w) + (ci - p.C1));
}

// resolves a padded coordinate; returns false when the tap reads a zero
__device__ __forceinline__ bool resolve(int& i, int n, int pad_mode) {
  if (i >= 0 && i < n) return true;
  if (pad_mode == SEGSDE_PAD_REFLECT) { i = reflect_idx(i, n); return true; }
  return false;
}

constexpr int BM = 64, BN = 64, BK = 16, CT = 256;

// ---------------------------------------------------------------------------------------------
// fprop: M = output pixels, N = Cout, K = (r,s,ci)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(CT) conv_fwd_kernel(ConvP p) {
  __shared__ float As[BK][BM + 4];
  __shared__ float Bs[BK][BN + 4];
  const int t = threadIdx.x;
  const long long m0 = (long long)blockIdx.x * BM;
  const int n0 = blockIdx.y * BN;
  const int lk = t % BK, lm = t / BK;     // loader coordinates
  int pn[4], ph[4], pw[4]; bool pv[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const long long m = m0 + lm + 16 * i;
    pv[i] = m < p.P;
    const long long mm = pv[i] ? m : 0;
    pw[i] = (int)(mm % p.Wo); const long long q = mm / p.Wo;
    ph[i] = (int)(q % p.Ho); pn[i] = (int)(q / p.Ho);
  }
  const int tx = t % 16, ty = t / 16;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  for (int k0 = 0; k0 < p.Ktot; k0 += BK) {
    const int k = k0 + lk;
    const bool kv = k < p.Ktot;
    int ci = 0, r = 0, s = 0;
    if (kv) { ci = k % p.Ctot; const int tap = k / p.Ctot; s = tap % p.kw; r = tap / p.kw; }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float v = 0.f;
      if (kv && pv[i]) {
        int h = ph[i] * p.stride - p.pad + r * p.dil;
        int w = pw[i] * p.stride - p.pad + s * p.dil;
        if (resolve(h, p.Hc, p.pad_mode) && resolve(w, p.Wc, p.pad_mode)) v = fetch_in(p, pn[i], h, w, ci);
      }
      As[lk][lm + 16 * i] = v;
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int co = n0 + lm + 16 * i;
      Bs[lk][lm + 16 * i] = (kv && co < p.Cout) ? __ldg(p.w + (long long)co * p.Ktot + k) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      const float4 a = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
      const float4 b = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
      const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const long long m = m0 + ty * 4 + i;
    if (m >= p.P) continue;
    const int ow = (int)(m % p.Wo); const long long q = m / p.Wo;
    const int oh = (int)(q % p.Ho), n = (int)(q / p.Ho);
    float* o = p.y.p + p.y.off(n, oh, ow);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int co = n0 + tx * 4 + j;
      if (co < p.Cout) o[co] = act_apply(acc[i][j] + (p.bias ? __ldg(p.bias + co) : 0.f), p.act);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// dgrad: M = padded-domain input pixels, N = ci, K = (r,s,co); p.y is dy.
// Results are folded (reflect / upsample / concat split) into dx1/dx2 = p.x1/p.x2.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(CT) conv_dgrad_kernel(ConvP p, int Hp, int Wp, int off, long long Pin,
                                                        int use_atomic) {
  __shared__ float As[BK][BM + 4];
  __shared__ float Bs[BK][BN + 4];
  const int t = threadIdx.x;
  const long long m0 = (long long)blockIdx.x * BM;
  const int n0 = blockIdx.y * BN;
  const int lk = t % BK, lm = t / BK;
  const int Kd = p.kh * p.kw * p.Cout;
  int pn[4], ph[4], pw[4]; bool pv[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const long long m = m0 + lm + 16 * i;
    pv[i] = m < Pin;
    const long long mm = pv[i] ? m : 0;
    pw[i] = (int)(mm % Wp) - off; const long long q = mm / Wp;
    ph[i] = (int)(q % Hp) - off; pn[i] = (int)(q / Hp);
  }
  const int tx = t % 16, ty = t / 16;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  for (int k0 = 0; k0 < Kd; k0 += BK) {
    const int k = k0 + lk;
    const bool kv = k < Kd;
    int co = 0, r = 0, s = 0;
    if (kv) { co = k % p.Cout; const int tap = k / p.Cout; s = tap % p.kw; r = tap / p.kw; }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float v = 0.f;
      if (kv && pv[i]) {
        const int hh = ph[i] + p.pad - r * p.dil, ww = pw[i] + p.pad - s * p.dil;
        if (hh >= 0 && ww >= 0 && hh % p.stride == 0 && ww % p.stride == 0) {
          const int oh = hh / p.stride, ow = ww / p.stride;
          if (oh < p.Ho && ow < p.Wo) v = __ldg(p.y.p + p.y.off(pn[i], oh, ow) + co);
        }
      }
      As[lk][lm + 16 * i] = v;
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int ci = n0 + lm + 16 * i;
      Bs[lk][lm + 16 * i] =
          (kv && ci < p.Ctot) ? __ldg(p.w + ((long long)co * p.kh * p.kw + r * p.kw + s) * p.Ctot + ci) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      const float4 a = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
      const float4 b = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
      const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const long long m = m0 + ty * 4 + i;
    if (m >= Pin) continue;
    int w = (int)(m % Wp) - off; const long long q = m / Wp;
    int h = (int)(q % Hp) - off; const int n = (int)(q / Hp);
    if (p.pad_mode == SEGSDE_PAD_REFLECT) { h = reflect_idx(h, p.Hc); w = reflect_idx(w, p.Wc); }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int ci = n0 + tx * 4 + j;
      if (ci >= p.Ctot) continue;
      float* dst;
      if (ci < p.C1) {
        if (!p.x1.p) continue;
        dst = p.x1.p + (p.up1 ? p.x1.off(n, h >> 1, w >> 1) : p.x1.off(n, h, w)) + ci;
      } else {
        if (!p.x2.p) continue;
        dst = p.x2.p + p.x2.off(n, h, w) + (ci - p.C1);
      }
      if (use_atomic) atomicAdd(dst, acc[i][j]); else *dst = acc[i][j];
    }
  }
}

// ---------------------------------------------------------------------------------------------
// wgrad: M = Cout, N = (r,s,ci), reduction over output pixels (split across blockIdx.z)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(CT) conv_wgrad_kernel(ConvP p, float* dw, float* dbias, long long chunk) {
  __shared__ float As[BK][BM + 4];   // [pixel][co]
  __shared__ float Bs[BK][BN + 4];   // [pixel][kcol]
  const int t = threadIdx.x;
  const int co0 = blockIdx.x * BM;
  const int kc0 = blockIdx.y * BN;
  const long long pbeg = (long long)blockIdx.z * chunk;
  const long long pend = min(p.P, pbeg + chunk);
  const int lc = t % 64, lp = t / 64;     // loader: column (co or kcol), pixel row 0..3
  const int kcol = kc0 + lc;
  const bool kv = kcol < p.Ktot;
  int ci = 0, r = 0, s = 0;
  if (kv) { ci = kcol % p.Ctot; const int tap = kcol / p.Ctot; s = tap % p.kw; r = tap / p.kw; }
  const int co_l = co0 + lc;
  const bool cov = co_l < p.Cout;
  const int tx = t % 16, ty = t / 16;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  float bsum = 0.f;

  for (long long p0 = pbeg; p0 < pend; p0 += BK) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const long long pp = p0 + lp + 4 * i;
      float a = 0.f, b = 0.f;
      if (pp < pend) {
        const int ow = (int)(pp % p.Wo); const long long q = pp / p.Wo;
        const int oh = (int)(q % p.Ho), n = (int)(q / p.Ho);
        if (cov) a = __ldg(p.y.p + p.y.off(n, oh, ow) + co_l);
        if (kv) {
          int h = oh * p.stride - p.pad + r * p.dil, w = ow * p.stride - p.pad + s * p.dil;
          if (resolve(h, p.Hc, p.pad_mode) && resolve(w, p.Wc, p.pad_mode)) b = fetch_in(p, n, h, w, ci);
        }
      }
      As[lp + 4 * i][lc] = a;
      Bs[lp + 4 * i][lc] = b;
      bsum += a;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      const float4 a = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
      const float4 b = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
      const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int co = co0 + ty * 4 + i;
    if (co >= p.Cout) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int kc = kc0 + tx * 4 + j;
      if (kc < p.Ktot) atomicAdd(dw + (long long)co * p.Ktot + kc, acc[i][j]);
    }
  }
  if (dbias && blockIdx.y == 0 && cov) atomicAdd(dbias + co_l, bsum);
}

__global__ void act_bwd_kernel(View y, View dy, View dz, int act) {
  const long long total = (long long)y.n * y.h * y.w * y.c;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int c = (int)(idx % y.c); long long q = idx / y.c;
  const int w = (int)(q % y.w); q /= y.w;
  const int h = (int)(q % y.h); const int n = (int)(q / y.h);
  const float yv = y.p[y.off(n, h, w) + c];
  dz.p[dz.off(n, h, w) + c] = dy.p[dy.off(n, h, w) + c] * act_grad_from_out(yv, act);
}

__global__ void act_fwd_kernel(View x, View y, int act) {
  const long long total = (long long)x.n * x.h * x.w * x.c;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int c = (int)(idx % x.c); long long q = idx / x.c;
  const int w = (int)(q % x.w); q /= x.w;
  const int h = (int)(q % x.h); const int n = (int)(q / x.h);
  y.p[y.off(n, h, w) + c] = act_apply(x.p[x.off(n, h, w) + c], act);
}

// 1x1 convolution over a handful of pixels (the ASPP image-pooling branch, model_parts.py:28-40: B x 2048 x 1 x 1):
// one warp per (pixel, output channel) dot product — the tiled kernel would run it on two CTAs.
__global__ void __launch_bounds__(256) conv1x1_fewpx_kernel(View x, View y, const float* __restrict__ w,
                                                            const float* __restrict__ bias, int C, int Cout, int act,
                                                            long long P) {
  const long long wid = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (wid >= P * Cout) return;
  const int co = (int)(wid % Cout); long long q = wid / Cout;
  const int wq = (int)(q % y.w); q /= y.w;
  const int hq = (int)(q % y.h); const int n = (int)(q / y.h);
  const float* xp = x.p + x.off(n, hq, wq);
  const float* wp = w + (long long)co * C;
  float s = 0.f;
  for (int c = lane; c < C; c += 32) s = fmaf(__ldg(xp + c), __ldg(wp + c), s);
  s = warp_sum(s);
  if (lane == 0) y.p[y.off(n, hq, wq) + co] = act_apply(s + (bias ? bias[co] : 0.f), act);
}

// conv_fewcout.cu
bool fewcout_ok(const View& x, const View& y, int kh, int kw, int stride, int pad, bool second_source, bool nchw, bool up1);
int fewcout_fwd(const View& x, const View& y, const float* w, const float* bias, int act, cudaStream_t st);
int fewcout_dgrad(const View& dy, const View& dx, const float* w, cudaStream_t st);
int fewcout_wgrad(const View& x, const View& dy, float* dw, cudaStream_t st);

static int fill(ConvP& p, const segsde_nhwc_t* x1, const segsde_nhwc_t* x2, const segsde_nhwc_t* y,
                const segsde_conv_desc_t* d, bool x_optional) {
  if (!y || !y->ptr || !d) return SEGSDE_E_ARG;
  if (!x_optional && (!x1 || !x1->ptr)) return SEGSDE_E_ARG;
  p.x1 = mk(x1); p.x2 = mk(x2); p.y = mk(y);
  p.C1 = x1 ? x1->c : 0; p.C2 = x2 ? x2->c : 0; p.Ctot = p.C1 + p.C2; p.Cout = y->c;
  if (p.Ctot < 1 || d->kh < 1 || d->kw < 1 || d->stride < 1 || d->dil < 1 || d->pad < 0) return SEGSDE_E_ARG;
  if (d->stride_w && d->stride_w != d->stride) return SEGSDE_E_UNSUPPORTED;   // anisotropic stride: tensor-core route only
  const segsde_nhwc_t* ref = x1 ? x1 : x2;
  const int up = (x1 && d->up1) ? 2 : 1;
  p.Hc = x1 ? x1->h * up : x2->h; p.Wc = x1 ? x1->w * up : x2->w;
  if (x1 && x2 && (x2->h != p.Hc || x2->w != p.Wc || x2->n != x1->n)) return SEGSDE_E_ARG;
  p.kh = d->kh; p.kw = d->kw; p.stride = d->stride; p.pad = d->pad; p.dil = d->dil;
  p.pad_mode = d->pad_mode; p.up1 = d->up1; p.act = d->act; p.nchw = d->nchw_norm_in;
  if (p.nchw && p.up1) return SEGSDE_E_ARG;
  if (p.pad_mode == SEGSDE_PAD_REFLECT && (p.pad >= p.Hc || p.pad >= p.Wc)) return SEGSDE_E_ARG;
  p.Ho = (p.Hc + 2 * p.pad - p.dil * (p.kh - 1) - 1) / p.stride + 1;
  p.Wo = (p.Wc + 2 * p.pad - p.dil * (p.kw - 1) - 1) / p.stride + 1;
  if (y->h != p.Ho || y->w != p.Wo || y->n != ref->n) return SEGSDE_E_ARG;
  p.Ktot = p.kh * p.kw * p.Ctot;
  p.P = (long long)y->n * p.Ho * p.Wo;
  return SEGSDE_OK;
}

}  // namespace segsde
using namespace segsde;

extern "C" int segsde_conv2d_fwd(const segsde_nhwc_t* x1, const segsde_nhwc_t* x2, const float* w,
                                 const float* bias, const segsde_nhwc_t* y, const segsde_conv_desc_t* d,
                                 void* stream) {
  ConvP p;
  int rc = fill(p, x1, x2, y, d, false);
  if (rc) return rc;
  if (!w) return SEGSDE_E_ARG;
  if (p.Cout == 1 && !p.x2.p) {
    rc = c1_fwd(p.x1, p.y, w, bias, d, as_stream(stream));
    if (rc != SEGSDE_E_UNSUPPORTED) return rc;
  }
  if (p.kh == 1 && p.kw == 1 && p.stride == 1 && p.pad == 0 && !p.x2.p && !p.nchw && !p.up1 && p.P <= 256) {
    conv1x1_fewpx_kernel<<<cdiv(p.P * p.Cout * 32, 256), 256, 0, as_stream(stream)>>>(p.x1, p.y, w, bias, p.C1, p.Cout,
                                                                                       p.act, p.P);
    return launched();
  }
  if (fewcout_ok(p.x1, p.y, p.kh, p.kw, p.stride, p.pad, p.x2.p != nullptr, p.nchw != 0, p.up1 != 0)) {
    rc = fewcout_fwd(p.x1, p.y, w, bias, p.act, as_stream(stream));
    if (rc != SEGSDE_E_UNSUPPORTED) return rc;
  }
  p.w = w; p.bias = bias;
  dim3 grid(cdiv(p.P, BM), cdiv(p.Cout, BN));
  conv_fwd_kernel<<<grid, CT, 0, as_stream(stream)>>>(p);
  return launched();
}

extern "C" int segsde_conv2d_dgrad(const segsde_nhwc_t* dy, const float* w, const segsde_nhwc_t* dx1,
                                   const segsde_nhwc_t* dx2, const segsde_conv_desc_t* d, void* stream) {
  ConvP p;
  if ((!dx1 || !dx1->ptr) && (!dx2 || !dx2->ptr)) return SEGSDE_OK;   // nothing asked for
  // shape bookkeeping needs both source shapes even when one gradient is skipped: callers pass the
  // view with ptr == NULL for a skipped source
  int rc = fill(p, dx1, dx2, dy, d, true);
  if (rc) return rc;
  if (!w || p.nchw) return SEGSDE_E_ARG;
  if (p.Cout == 1 && p.x1.p && !dx2) {
    rc = c1_dgrad(p.y, w, p.x1, d, as_stream(stream));
    if (rc != SEGSDE_E_UNSUPPORTED) return rc;
  }
  if (p.x1.p && (!dx2 || !dx2->ptr) && p.C2 == 0 &&
      fewcout_ok(p.x1, p.y, p.kh, p.kw, p.stride, p.pad, false, p.nchw != 0, p.up1 != 0)) {
    rc = fewcout_dgrad(p.y, p.x1, w, as_stream(stream));
    if (rc != SEGSDE_E_UNSUPPORTED) return rc;
  }
  p.w = w; p.bias = nullptr;
  const bool refl = p.pad_mode == SEGSDE_PAD_REFLECT;
  const int off = refl ? p.pad : 0;
  const int Hp = p.Hc + 2 * off, Wp = p.Wc + 2 * off;
  const long long Pin = (long long)dy->n * Hp * Wp;
  const int use_atomic = (refl || p.up1) ? 1 : 0;
  dim3 grid(cdiv(Pin, BM), cdiv(p.Ctot, BN));
  conv_dgrad_kernel<<<grid, CT, 0, as_stream(stream)>>>(p, Hp, Wp, off, Pin, use_atomic);
  return launched();
}

extern "C" int segsde_conv2d_wgrad(const segsde_nhwc_t* x1, const segsde_nhwc_t* x2,
                                   const segsde_nhwc_t* dy, float* dw, float* dbias,
                                   const segsde_conv_desc_t* d, void* stream) {
  ConvP p;
  int rc = fill(p, x1, x2, dy, d, false);
  if (rc) return rc;
  if (!dw) return SEGSDE_E_ARG;
  if (p.Cout == 1 && !p.x2.p && !dbias) {
    rc = c1_wgrad(p.x1, p.y, dw, d, as_stream(stream));
    if (rc != SEGSDE_E_UNSUPPORTED) return rc;
  }
  if (!dbias && fewcout_ok(p.x1, p.y, p.kh, p.kw, p.stride, p.pad, p.x2.p != nullptr, p.nchw != 0, p.up1 != 0))
    return fewcout_wgrad(p.x1, p.y, dw, as_stream(stream));
  p.w = nullptr; p.bias = nullptr;
  const int gx = cdiv(p.Cout, BM), gy = cdiv(p.Ktot, BN);
  long long want = (148LL * 6) / ((long long)gx * gy);
  if (want < 1) want = 1;
  long long chunks = cdiv(p.P, 128);
  if (chunks > want) chunks = want;
  if (chunks < 1) chunks = 1;
  if (chunks > 65535) chunks = 65535;
  long long chunk = (p.P + chunks - 1) / chunks;
  chunk = ((chunk + BK - 1) / BK) * BK;
  dim3 grid(gx, gy, cdiv(p.P, chunk));
  conv_wgrad_kernel<<<grid, CT, 0, as_stream(stream)>>>(p, dw, dbias, chunk);
  return launched();
}

extern "C" int segsde_act_bwd(const segsde_nhwc_t* y, const segsde_nhwc_t* dy, const segsde_nhwc_t* dz,
                              int act, void* stream) {
  if (!y || !dy || !dz || !y->ptr || !dy->ptr || !dz->ptr) return SEGSDE_E_ARG;
  View vy = mk(y), vdy = mk(dy), vdz = mk(dz);
  if (!same_shape(vy, vdy) || !same_shape(vy, vdz)) return SEGSDE_E_ARG;
  const long long total = (long long)vy.n * vy.h * vy.w * vy.c;
  if (total == 0) return SEGSDE_OK;
  act_bwd_kernel<<<cdiv(total, 256), 256, 0, as_stream(stream)>>>(vy, vdy, vdz, act);
  return launched();
}

extern "C" int segsde_act_fwd(const segsde_nhwc_t* x, const segsde_nhwc_t* y, int act, void* stream) {
  if (!x || !y || !x->ptr || !y->ptr) return SEGSDE_E_ARG;
  View vx = mk(x), vy = mk(y);
  if (!same_shape(vx, vy)) return SEGSDE_E_ARG;
  const long long total = (long long)vx.n * vx.h * vx.w * vx.c;
  if (total == 0) return SEGSDE_OK;
  act_fwd_kernel<<<cdiv(total, 256), 256, 0, as_stream(stream)>>>(vx, vy, act);
  return launched();
}
