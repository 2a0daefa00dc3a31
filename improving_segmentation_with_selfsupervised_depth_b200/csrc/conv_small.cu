// Single-output-channel convolutions (the four sigmoid disparity heads, depth_decoder.py:69-70 / :107-112):
// C -> 1, k x k, stride 1, reflection or zero padding.  These are HBM-bound stencils (9*C MACs per pixel), not
// GEMMs: a group of C/4 lanes owns one pixel (coalesced float4 channel reads), taps re-read through L1, the
// channel sum is a shuffle reduction.  Called from the generic entry points in conv_simt.cu.
#include "common.cuh"

namespace segsde {

struct C1P {
  View x, y;            // x: input [N,H,W,C] ; y: output / dy [N,Ho,Wo,1]
  const float* w;       // [1][kh][kw][C]
  const float* bias;
  int C, kh, kw, pad, dil, reflect, act, Ho, Wo;
};

__device__ __forceinline__ bool c1_resolve(int& i, int n, int reflect) {
  if (i >= 0 && i < n) return true;
  if (reflect) { i = reflect_idx(i, n); return true; }
  return false;
}

// One CTA per 32 x 8 output-pixel tile (taps of neighbouring pixels hit in L1); G lanes per pixel
// (G = C/4 capped at 32), 32/G pixels per warp iteration, one tile row per warp.
__global__ void __launch_bounds__(256) c1_fwd_kernel(C1P p, int G, int cpl, int tiles_x, int tiles_y) {
  extern __shared__ float sw[];       // weights [kh*kw][C]
  const int taps = p.kh * p.kw;
  for (int i = threadIdx.x; i < taps * p.C; i += blockDim.x) sw[i] = p.w[i];
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int ppw = 32 / G;
  const int sub = lane / G, gl = lane % G;
  int t = blockIdx.x;
  const int tx = t % tiles_x; t /= tiles_x;
  const int ty = t % tiles_y; const int n = t / tiles_y;
  const int oh = ty * 8 + warp;
  for (int it = 0; it < 32 / ppw; ++it) {
    const int ow = tx * 32 + it * ppw + sub;
    const bool pv = oh < p.Ho && ow < p.Wo;
    float acc = 0.f;
    if (pv) {
      for (int r = 0; r < p.kh; ++r) {
        int h = oh - p.pad + r * p.dil;
        if (!c1_resolve(h, p.x.h, p.reflect)) continue;
        for (int s = 0; s < p.kw; ++s) {
          int w = ow - p.pad + s * p.dil;
          if (!c1_resolve(w, p.x.w, p.reflect)) continue;
          const float* xp = p.x.p + p.x.off(n, h, w);
          const float* wp = sw + (r * p.kw + s) * p.C;
          for (int j = 0; j < cpl; ++j) {
            const int c = (gl + j * G) * 4;
            const float4 a = *reinterpret_cast<const float4*>(xp + c);
            const float4 b = *reinterpret_cast<const float4*>(wp + c);
            acc = fmaf(a.x, b.x, acc); acc = fmaf(a.y, b.y, acc); acc = fmaf(a.z, b.z, acc); acc = fmaf(a.w, b.w, acc);
          }
        }
      }
    }
    for (int o = G >> 1; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (pv && gl == 0) p.y.p[p.y.off(n, oh, ow)] = act_apply(acc + (p.bias ? p.bias[0] : 0.f), p.act);
  }
}

// dx[n,h,w,c] = sum over reflect preimages (hp,wp) of (h,w) of sum_taps dy[hp+pad-r*dil, wp+pad-s*dil] * w[r,s,c]
__global__ void __launch_bounds__(256) c1_dgrad_kernel(C1P p, View dx) {
  const int cq = p.C / 4;
  const long long total = (long long)dx.n * dx.h * dx.w * cq;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int c = (int)(idx % cq) * 4; long long q = idx / cq;
  const int w = (int)(q % dx.w); q /= dx.w;
  const int h = (int)(q % dx.h); const int n = (int)(q / dx.h);
  int rows[3], cols[3]; int nr = 0, nc = 0;
  rows[nr++] = h; cols[nc++] = w;
  if (p.reflect) {
    if (h >= 1 && h <= p.pad) rows[nr++] = -h;
    if (h >= dx.h - 1 - p.pad && h <= dx.h - 2) rows[nr++] = 2 * (dx.h - 1) - h;
    if (w >= 1 && w <= p.pad) cols[nc++] = -w;
    if (w >= dx.w - 1 - p.pad && w <= dx.w - 2) cols[nc++] = 2 * (dx.w - 1) - w;
  }
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int i = 0; i < nr; ++i)
    for (int j = 0; j < nc; ++j)
      for (int r = 0; r < p.kh; ++r) {
        const int oh = rows[i] + p.pad - r * p.dil;
        if (oh < 0 || oh >= p.Ho) continue;
        for (int s = 0; s < p.kw; ++s) {
          const int ow = cols[j] + p.pad - s * p.dil;
          if (ow < 0 || ow >= p.Wo) continue;
          const float g = __ldg(p.y.p + p.y.off(n, oh, ow));
          const float4 wv = *reinterpret_cast<const float4*>(p.w + (r * p.kw + s) * p.C + c);
          acc.x = fmaf(g, wv.x, acc.x); acc.y = fmaf(g, wv.y, acc.y); acc.z = fmaf(g, wv.z, acc.z); acc.w = fmaf(g, wv.w, acc.w);
        }
      }
  *reinterpret_cast<float4*>(dx.p + dx.off(n, h, w) + c) = acc;
}

// dw[tap][c] += sum_pixels dy[p] * xv[p + tap, c]; each thread owns 4 channels and (up to) 9 taps
template <int TAPS>
__global__ void __launch_bounds__(256) c1_wgrad_kernel(C1P p, float* __restrict__ dw, long long chunk) {
  const int cq = p.C / 4;                       // float4 lanes per pixel
  const int lanes = blockDim.x / cq * cq;       // threads used
  const int pl = threadIdx.x / cq, cl = threadIdx.x % cq;   // pixel lane, channel lane
  const int npl = blockDim.x / cq;
  // the CTA walks `chunk` consecutive 32 x 8 pixel tiles so that tap re-reads stay in L1
  const int tiles_x = (p.Wo + 31) / 32, tiles_y = (p.Ho + 7) / 8;
  const long long T = (long long)p.y.n * tiles_x * tiles_y;
  const long long tbeg = (long long)blockIdx.x * chunk, tend = min(T, tbeg + chunk);
  float4 acc[TAPS];
#pragma unroll
  for (int t = 0; t < TAPS; ++t) acc[t] = make_float4(0.f, 0.f, 0.f, 0.f);
  if ((int)threadIdx.x < lanes) {
    for (long long tile = tbeg; tile < tend; ++tile) {
      long long q = tile;
      const int tx = (int)(q % tiles_x); q /= tiles_x;
      const int ty = (int)(q % tiles_y); const int n = (int)(q / tiles_y);
      for (int lp = pl; lp < 256; lp += npl) {
        const int oh = ty * 8 + (lp >> 5), ow = tx * 32 + (lp & 31);
        if (oh >= p.Ho || ow >= p.Wo) continue;
        const float g = __ldg(p.y.p + p.y.off(n, oh, ow));
#pragma unroll
        for (int t = 0; t < TAPS; ++t) {
          int h = oh - p.pad + (t / p.kw) * p.dil, w = ow - p.pad + (t % p.kw) * p.dil;
          if (!c1_resolve(h, p.x.h, p.reflect) || !c1_resolve(w, p.x.w, p.reflect)) continue;
          const float4 a = *reinterpret_cast<const float4*>(p.x.p + p.x.off(n, h, w) + cl * 4);
          acc[t].x = fmaf(g, a.x, acc[t].x); acc[t].y = fmaf(g, a.y, acc[t].y);
          acc[t].z = fmaf(g, a.z, acc[t].z); acc[t].w = fmaf(g, a.w, acc[t].w);
        }
      }
    }
  }
  // reduce over the pixel lanes through shared memory, then one atomic per (tap, channel) and block
  extern __shared__ float red[];                 // [npl][TAPS][C]
  if ((int)threadIdx.x < lanes) {
#pragma unroll
    for (int t = 0; t < TAPS; ++t)
      *reinterpret_cast<float4*>(red + ((size_t)pl * TAPS + t) * p.C + cl * 4) = acc[t];
  }
  __syncthreads();
  for (int i = threadIdx.x; i < TAPS * p.C; i += blockDim.x) {
    float s = 0.f;
    for (int j = 0; j < npl; ++j) s += red[(size_t)j * TAPS * p.C + i];
    atomicAdd(dw + i, s);
  }
}

static bool c1_ok(const View& x, int cout, int kh, int kw) {
  return cout == 1 && x.c % 4 == 0 && x.c >= 16 && x.c <= 1024 && vec4_ok(x) && kh * kw <= 9 &&
         ((x.c / 4) & (x.c / 4 - 1)) == 0;      // C/4 a power of two (lane groups)
}

int c1_fwd(const View& x, const View& y, const float* w, const float* bias, const segsde_conv_desc_t* d, cudaStream_t st) {
  if (!c1_ok(x, y.c, d->kh, d->kw) || d->stride != 1 || d->up1 || d->nchw_norm_in) return SEGSDE_E_UNSUPPORTED;
  C1P p; p.x = x; p.y = y; p.w = w; p.bias = bias; p.C = x.c; p.kh = d->kh; p.kw = d->kw; p.pad = d->pad; p.dil = d->dil;
  p.reflect = d->pad_mode == SEGSDE_PAD_REFLECT; p.act = d->act; p.Ho = y.h; p.Wo = y.w;
  const int G = x.c / 4 < 32 ? x.c / 4 : 32, cpl = (x.c / 4 + G - 1) / G;
  const size_t smem = sizeof(float) * p.kh * p.kw * p.C;
  const int tiles_x = cdiv(y.w, 32), tiles_y = cdiv(y.h, 8);
  c1_fwd_kernel<<<tiles_x * tiles_y * y.n, 256, smem, st>>>(p, G, cpl, tiles_x, tiles_y);
  return launched();
}
int c1_dgrad(const View& dy, const float* w, const View& dx, const segsde_conv_desc_t* d, cudaStream_t st) {
  if (!c1_ok(dx, dy.c, d->kh, d->kw) || d->stride != 1 || d->up1 || d->nchw_norm_in) return SEGSDE_E_UNSUPPORTED;
  C1P p; p.x = dx; p.y = dy; p.w = w; p.bias = nullptr; p.C = dx.c; p.kh = d->kh; p.kw = d->kw; p.pad = d->pad; p.dil = d->dil;
  p.reflect = d->pad_mode == SEGSDE_PAD_REFLECT; p.act = 0; p.Ho = dy.h; p.Wo = dy.w;
  const long long total = (long long)dx.n * dx.h * dx.w * (dx.c / 4);
  c1_dgrad_kernel<<<cdiv(total, 256), 256, 0, st>>>(p, dx);
  return launched();
}
int c1_wgrad(const View& x, const View& dy, float* dw, const segsde_conv_desc_t* d, cudaStream_t st) {
  if (!c1_ok(x, dy.c, d->kh, d->kw) || d->stride != 1 || d->up1 || d->nchw_norm_in || x.c > 256) return SEGSDE_E_UNSUPPORTED;
  if (d->kh * d->kw != 9 && d->kh * d->kw != 1) return SEGSDE_E_UNSUPPORTED;
  C1P p; p.x = x; p.y = dy; p.w = nullptr; p.bias = nullptr; p.C = x.c; p.kh = d->kh; p.kw = d->kw; p.pad = d->pad; p.dil = d->dil;
  p.reflect = d->pad_mode == SEGSDE_PAD_REFLECT; p.act = 0; p.Ho = dy.h; p.Wo = dy.w;
  const long long P = (long long)dy.n * cdiv(dy.h, 8) * cdiv(dy.w, 32);     // number of 32x8 tiles
  int blocks = 148 * 4;
  if (blocks > P) blocks = (int)P;
  const long long chunk = (P + blocks - 1) / blocks;
  const int npl = 256 / (x.c / 4);
  const int taps = d->kh * d->kw;
  const size_t smem = sizeof(float) * (size_t)npl * taps * x.c;
  if (taps == 9) {
    static bool attr = false;
    if (!attr) { cudaFuncSetAttribute(c1_wgrad_kernel<9>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024); attr = true; }
    c1_wgrad_kernel<9><<<cdiv(P, chunk), 256, smem, st>>>(p, dw, chunk);
  } else {
    c1_wgrad_kernel<1><<<cdiv(P, chunk), 256, smem, st>>>(p, dw, chunk);
  }
  return launched();
}

}  // namespace segsde

// ---------------------------------------------------------------------------------------------------
// Disparity heads on the tensor cores: a C -> 1 3x3 conv is a 1x1 conv to 9 "tap planes" (z[p][t] = <x[p], w[t]>,
// a [P x C] x [C x 9] GEMM, N padded to 32) followed by a 9-tap scalar stencil; its backward needs the adjoint
// stencil of dy (gcol[q][t] = sum of dy over the outputs whose tap t reads q) and two more 1x1 GEMMs.
// ---------------------------------------------------------------------------------------------------
namespace segsde {

__global__ void head_stencil_fwd_kernel(View z, View y, const float* __restrict__ bias, int act, int reflect, int pad) {
  const long long total = (long long)y.n * y.h * y.w;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int w = (int)(idx % y.w); long long q = idx / y.w;
  const int h = (int)(q % y.h); const int n = (int)(q / y.h);
  float acc = bias ? bias[0] : 0.f;
#pragma unroll
  for (int r = 0; r < 3; ++r) {
    int hh = h - pad + r;
    if (!c1_resolve(hh, z.h, reflect)) continue;
#pragma unroll
    for (int s = 0; s < 3; ++s) {
      int ww = w - pad + s;
      if (!c1_resolve(ww, z.w, reflect)) continue;
      acc += __ldg(z.p + z.off(n, hh, ww) + r * 3 + s);
    }
  }
  y.p[y.off(n, h, w)] = act_apply(acc, act);
}

// gcol[n,h,w,t] = sum over padded-domain preimages (hp,wp) of (h,w) of dy[hp + pad - r, wp + pad - s], t = r*3+s
__global__ void head_gcol_kernel(View dy, View g, int reflect, int pad) {
  const long long total = (long long)g.n * g.h * g.w;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int w = (int)(idx % g.w); long long q = idx / g.w;
  const int h = (int)(q % g.h); const int n = (int)(q / g.h);
  int rows[3], cols[3]; int nr = 0, nc = 0;
  rows[nr++] = h; cols[nc++] = w;
  if (reflect) {
    if (h >= 1 && h <= pad) rows[nr++] = -h;
    if (h >= g.h - 1 - pad && h <= g.h - 2) rows[nr++] = 2 * (g.h - 1) - h;
    if (w >= 1 && w <= pad) cols[nc++] = -w;
    if (w >= g.w - 1 - pad && w <= g.w - 2) cols[nc++] = 2 * (g.w - 1) - w;
  }
  float v[9];
#pragma unroll
  for (int t = 0; t < 9; ++t) v[t] = 0.f;
  for (int i = 0; i < nr; ++i)
    for (int j = 0; j < nc; ++j)
#pragma unroll
      for (int r = 0; r < 3; ++r) {
        const int oh = rows[i] + pad - r;
        if (oh < 0 || oh >= dy.h) continue;
#pragma unroll
        for (int s = 0; s < 3; ++s) {
          const int ow = cols[j] + pad - s;
          if (ow < 0 || ow >= dy.w) continue;
          v[r * 3 + s] += __ldg(dy.p + dy.off(n, oh, ow));
        }
      }
  float* o = g.p + g.off(n, h, w);
  float4* o4 = reinterpret_cast<float4*>(o);
  o4[0] = make_float4(v[0], v[1], v[2], v[3]);
  o4[1] = make_float4(v[4], v[5], v[6], v[7]);
  o4[2] = make_float4(v[8], 0.f, 0.f, 0.f);
  const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int k = 3; k < (g.c >> 2); ++k) o4[k] = zero;       // 12 planes (CUDA-core heads) or 32 (tcgen05 heads)
}

}  // namespace segsde
using namespace segsde;

extern "C" int segsde_head_stencil_fwd(const segsde_nhwc_t* z, const segsde_nhwc_t* y, const float* bias, int act,
                                       int reflect, int pad, void* stream) {
  if (!z || !y || !z->ptr || !y->ptr || z->c < 9 || y->c != 1 || pad != 1) return SEGSDE_E_ARG;
  View vz = mk(z), vy = mk(y);
  if (vz.h != vy.h || vz.w != vy.w || vz.n != vy.n) return SEGSDE_E_ARG;
  const long long total = (long long)vy.n * vy.h * vy.w;
  head_stencil_fwd_kernel<<<cdiv(total, 256), 256, 0, as_stream(stream)>>>(vz, vy, bias, act, reflect, pad);
  return launched();
}
extern "C" int segsde_head_gcol(const segsde_nhwc_t* dy, const segsde_nhwc_t* gcol, int reflect, int pad, void* stream) {
  if (!dy || !gcol || !dy->ptr || !gcol->ptr || gcol->c < 12 || gcol->c % 4 || dy->c != 1 || pad != 1) return SEGSDE_E_ARG;
  View vd = mk(dy), vg = mk(gcol);
  if (vd.h != vg.h || vd.w != vg.w || vd.n != vg.n || !vec4_ok(vg)) return SEGSDE_E_ARG;
  const long long total = (long long)vg.n * vg.h * vg.w;
  head_gcol_kernel<<<cdiv(total, 256), 256, 0, as_stream(stream)>>>(vd, vg, reflect, pad);
  return launched();
}
