// Disparity heads: C -> 1, 3x3, pad 1 (reflection or zero), optional sigmoid — depth_decoder.py:69-70,107-112.
//
// A C -> 1 convolution has 2*9*C flops per C*4 input bytes: it is HBM-bound (at 512x1024, C = 64, B = 12 the input
// is 1.6 GB), so these kernels are organised around reading x (or writing dx) exactly once, with fully used 128-byte
// lines, and keeping everything else on chip:
//   fwd   : per 32x8 tile (+1 halo) the nine "tap planes" z[q][t] = <x[q,:], w[t,:]> are formed once per input pixel
//           (4 lanes per pixel, 16 channels each, two pixels in flight per lane so a weight vector fetched from shared
//           memory feeds 8 FMAs), parked in shared memory, then a 9-tap scalar stencil + bias + activation writes y.
//   dgrad : g[q][t] (adjoint stencil of dz incl. the reflection pre-images) per tile in shared memory, then
//           dx[q,:] = sum_t g[q][t] w[t,:], 64 contiguous bytes per lane.
//   wgrad : dw[t,:] += sum_q g[q][t] x[q,:] with 16 lanes per pixel (4 channels x 9 taps of accumulators per lane),
//           persistent CTAs, one atomic flush per CTA and 64-channel block.
// (Round-1 first form: a 1x1 tensor-core GEMM to 32 tap planes in HBM + stencil, 0.95 ms forward at scale 0; the
// tap planes alone cost 0.8 GB of writes and reads.)
#include "common.cuh"

namespace segsde {

constexpr int HT_X = 32, HT_Y = 8, HT_N = HT_X * HT_Y;          // output tile
constexpr int HZ_W = HT_X + 2, HZ_H = HT_Y + 2, HZ_N = HZ_W * HZ_H;   // tile + 1 halo (forward tap planes)

__device__ __forceinline__ bool head_resolve(int& i, int n, int reflect) {
  if (i >= 0 && i < n) return true;
  if (reflect) { i = reflect_idx(i, n); return true; }
  return false;
}
__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }

__global__ void __launch_bounds__(256) head_fwd_kernel(View x, View y, const float* __restrict__ w,
                                                       const float* __restrict__ bias, int act, int reflect,
                                                       int tiles_x, int tiles_y) {
  extern __shared__ __align__(16) float hsm[];
  const int C = x.c;
  float* ws = hsm;                 // [9][C]
  float* zs = hsm + 9 * C;         // [HZ_N][9]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  int b = blockIdx.x;
  const int tix = b % tiles_x; b /= tiles_x;
  const int tiy = b % tiles_y; const int n = b / tiles_y;
  const int w0 = tix * HT_X, h0 = tiy * HT_Y;
  for (int i = tid; i < 9 * C; i += 256) ws[i] = __ldg(w + i);
  __syncthreads();

  const int sub = lane & 3, pl = lane >> 2;          // channel quarter of a 64-channel block / pixel of the group
  for (int base = warp * 16; base < HZ_N; base += 128) {
    const int i0 = base + pl, i1 = base + 8 + pl;
    const int gh0 = h0 - 1 + i0 / HZ_W, gw0 = w0 - 1 + i0 % HZ_W;
    const int gh1 = h0 - 1 + i1 / HZ_W, gw1 = w0 - 1 + i1 % HZ_W;
    const bool v0 = i0 < HZ_N && gh0 >= 0 && gh0 < x.h && gw0 >= 0 && gw0 < x.w;
    const bool v1 = i1 < HZ_N && gh1 >= 0 && gh1 < x.h && gw1 >= 0 && gw1 < x.w;
    const float* p0 = x.p + (v0 ? x.off(n, gh0, gw0) : 0);
    const float* p1 = x.p + (v1 ? x.off(n, gh1, gw1) : 0);
    float a0[9], a1[9];
#pragma unroll
    for (int t = 0; t < 9; ++t) { a0[t] = 0.f; a1[t] = 0.f; }
    const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int cb = 0; cb < C; cb += 64) {
      const int c0 = cb + sub * 16;
      float4 xa[4], xb[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) { xa[j] = v0 ? ldg4(p0 + c0 + 4 * j) : zero4; xb[j] = v1 ? ldg4(p1 + c0 + 4 * j) : zero4; }
#pragma unroll
      for (int t = 0; t < 9; ++t) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float4 wv = *reinterpret_cast<const float4*>(ws + t * C + c0 + 4 * j);
          a0[t] = fmaf(xa[j].x, wv.x, fmaf(xa[j].y, wv.y, fmaf(xa[j].z, wv.z, fmaf(xa[j].w, wv.w, a0[t]))));
          a1[t] = fmaf(xb[j].x, wv.x, fmaf(xb[j].y, wv.y, fmaf(xb[j].z, wv.z, fmaf(xb[j].w, wv.w, a1[t]))));
        }
      }
    }
#pragma unroll
    for (int t = 0; t < 9; ++t) {
      a0[t] += __shfl_xor_sync(0xffffffffu, a0[t], 1); a0[t] += __shfl_xor_sync(0xffffffffu, a0[t], 2);
      a1[t] += __shfl_xor_sync(0xffffffffu, a1[t], 1); a1[t] += __shfl_xor_sync(0xffffffffu, a1[t], 2);
    }
    if (sub == 0) {
      if (i0 < HZ_N) {
#pragma unroll
        for (int t = 0; t < 9; ++t) zs[i0 * 9 + t] = a0[t];
      }
      if (i1 < HZ_N) {
#pragma unroll
        for (int t = 0; t < 9; ++t) zs[i1 * 9 + t] = a1[t];
      }
    }
  }
  __syncthreads();

  const int h = h0 + tid / HT_X, wq = w0 + tid % HT_X;
  if (h < y.h && wq < y.w) {
    float acc = bias ? __ldg(bias) : 0.f;
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      int hh = h - 1 + r;
      if (!head_resolve(hh, x.h, reflect)) continue;
#pragma unroll
      for (int s = 0; s < 3; ++s) {
        int ww = wq - 1 + s;
        if (!head_resolve(ww, x.w, reflect)) continue;
        acc += zs[((hh - (h0 - 1)) * HZ_W + (ww - (w0 - 1))) * 9 + r * 3 + s];
      }
    }
    y.p[y.off(n, h, wq)] = act_apply(acc, act);
  }
}

// g[q][t = r*3+s] = sum over the padded-domain pre-images (hp, wp) of q of dz[hp + 1 - r, wp + 1 - s]  (pad 1)
__device__ __forceinline__ void head_gcol_pixel(const View& dz, int n, int h, int w, int reflect, float (&v)[9]) {
#pragma unroll
  for (int t = 0; t < 9; ++t) v[t] = 0.f;
  if (h >= dz.h || w >= dz.w) return;
  int rows[3], cols[3]; int nr = 0, nc = 0;
  rows[nr++] = h; cols[nc++] = w;
  if (reflect) {
    if (h == 1) rows[nr++] = -1;
    if (h == dz.h - 2) rows[nr++] = dz.h;
    if (w == 1) cols[nc++] = -1;
    if (w == dz.w - 2) cols[nc++] = dz.w;
  }
  for (int i = 0; i < nr; ++i)
    for (int j = 0; j < nc; ++j)
#pragma unroll
      for (int r = 0; r < 3; ++r) {
        const int oh = rows[i] + 1 - r;
        if (oh < 0 || oh >= dz.h) continue;
#pragma unroll
        for (int s = 0; s < 3; ++s) {
          const int ow = cols[j] + 1 - s;
          if (ow < 0 || ow >= dz.w) continue;
          v[r * 3 + s] += __ldg(dz.p + dz.off(n, oh, ow));
        }
      }
}

__global__ void __launch_bounds__(256) head_dgrad_kernel(View dz, View dx, const float* __restrict__ w, int reflect,
                                                         int tiles_x, int tiles_y) {
  extern __shared__ __align__(16) float hsm[];
  const int C = dx.c;
  float* ws = hsm;                 // [9][C]
  float* gs = hsm + 9 * C;         // [HT_N][9]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  int b = blockIdx.x;
  const int tix = b % tiles_x; b /= tiles_x;
  const int tiy = b % tiles_y; const int n = b / tiles_y;
  const int w0 = tix * HT_X, h0 = tiy * HT_Y;
  for (int i = tid; i < 9 * C; i += 256) ws[i] = __ldg(w + i);
  {
    float v[9];
    head_gcol_pixel(dz, n, h0 + tid / HT_X, w0 + tid % HT_X, reflect, v);
#pragma unroll
    for (int t = 0; t < 9; ++t) gs[tid * 9 + t] = v[t];
  }
  __syncthreads();
  const int sub = lane & 3, pl = lane >> 2;
  for (int base = warp * 16; base < HT_N; base += 128) {
    const int i0 = base + pl, i1 = base + 8 + pl;
    const int gh0 = h0 + i0 / HT_X, gw0 = w0 + i0 % HT_X, gh1 = h0 + i1 / HT_X, gw1 = w0 + i1 % HT_X;
    const bool v0 = gh0 < dx.h && gw0 < dx.w, v1 = gh1 < dx.h && gw1 < dx.w;
    float* p0 = dx.p + (v0 ? dx.off(n, gh0, gw0) : 0);
    float* p1 = dx.p + (v1 ? dx.off(n, gh1, gw1) : 0);
    float g0[9], g1[9];
#pragma unroll
    for (int t = 0; t < 9; ++t) { g0[t] = gs[i0 * 9 + t]; g1[t] = gs[i1 * 9 + t]; }
    for (int cb = 0; cb < C; cb += 64) {
      const int c0 = cb + sub * 16;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float4 o0 = make_float4(0.f, 0.f, 0.f, 0.f), o1 = o0;
#pragma unroll
        for (int t = 0; t < 9; ++t) {
          const float4 wv = *reinterpret_cast<const float4*>(ws + t * C + c0 + 4 * j);
          o0.x = fmaf(g0[t], wv.x, o0.x); o0.y = fmaf(g0[t], wv.y, o0.y); o0.z = fmaf(g0[t], wv.z, o0.z); o0.w = fmaf(g0[t], wv.w, o0.w);
          o1.x = fmaf(g1[t], wv.x, o1.x); o1.y = fmaf(g1[t], wv.y, o1.y); o1.z = fmaf(g1[t], wv.z, o1.z); o1.w = fmaf(g1[t], wv.w, o1.w);
        }
        if (v0) *reinterpret_cast<float4*>(p0 + c0 + 4 * j) = o0;
        if (v1) *reinterpret_cast<float4*>(p1 + c0 + 4 * j) = o1;
      }
    }
  }
}

__global__ void __launch_bounds__(256) head_wgrad_kernel(View x, View dz, float* __restrict__ dw, int reflect, int tiles_x,
                                                         int tiles_y, long long total_tiles) {
  __shared__ float gs[HT_N * 9];
  __shared__ float red[8][9 * 64];
  const int C = x.c;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int l16 = lane & 15, half = lane >> 4;       // 16 lanes x 4 channels per pixel, two pixels per warp step
  for (int cb = 0; cb < C; cb += 64) {
    float acc[9][4];
#pragma unroll
    for (int t = 0; t < 9; ++t) { acc[t][0] = acc[t][1] = acc[t][2] = acc[t][3] = 0.f; }
    for (long long tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      long long b = tile;
      const int tix = (int)(b % tiles_x); b /= tiles_x;
      const int tiy = (int)(b % tiles_y); const int n = (int)(b / tiles_y);
      const int w0 = tix * HT_X, h0 = tiy * HT_Y;
      {
        float v[9];
        head_gcol_pixel(dz, n, h0 + tid / HT_X, w0 + tid % HT_X, reflect, v);
#pragma unroll
        for (int t = 0; t < 9; ++t) gs[tid * 9 + t] = v[t];
      }
      __syncthreads();
#pragma unroll 4
      for (int it = 0; it < 16; ++it) {
        const int i = warp * 32 + it * 2 + half;
        const int gh = h0 + i / HT_X, gw = w0 + i % HT_X;
        if (gh < x.h && gw < x.w) {
          const float4 xv = ldg4(x.p + x.off(n, gh, gw) + cb + l16 * 4);
#pragma unroll
          for (int t = 0; t < 9; ++t) {
            const float g = gs[i * 9 + t];
            acc[t][0] = fmaf(g, xv.x, acc[t][0]); acc[t][1] = fmaf(g, xv.y, acc[t][1]);
            acc[t][2] = fmaf(g, xv.z, acc[t][2]); acc[t][3] = fmaf(g, xv.w, acc[t][3]);
          }
        }
      }
      __syncthreads();
    }
    // CTA reduction: the two half-warps, then the eight warps, then one atomic per (tap, channel)
#pragma unroll
    for (int t = 0; t < 9; ++t)
#pragma unroll
      for (int k = 0; k < 4; ++k) acc[t][k] += __shfl_xor_sync(0xffffffffu, acc[t][k], 16);
    if (half == 0) {
#pragma unroll
      for (int t = 0; t < 9; ++t)
#pragma unroll
        for (int k = 0; k < 4; ++k) red[warp][t * 64 + l16 * 4 + k] = acc[t][k];
    }
    __syncthreads();
    for (int i = tid; i < 9 * 64; i += 256) {
      float s = 0.f;
#pragma unroll
      for (int wq = 0; wq < 8; ++wq) s += red[wq][i];
      atomicAdd(dw + (i / 64) * C + cb + (i % 64), s);
    }
    __syncthreads();
  }
}

static bool head_ok(const View& x, const View& y) {
  return x.p && y.p && y.c == 1 && x.c % 64 == 0 && x.c <= 512 && x.h == y.h && x.w == y.w && x.n == y.n && x.h >= 2 && x.w >= 2 &&
         vec4_ok(x);
}

}  // namespace segsde
using namespace segsde;

extern "C" int segsde_head_fwd_fused(const segsde_nhwc_t* x, const float* w, const float* bias, const segsde_nhwc_t* y,
                                     int act, int reflect, void* stream) {
  if (!x || !y || !w) return SEGSDE_E_ARG;
  View vx = mk(x), vy = mk(y);
  if (!head_ok(vx, vy)) return SEGSDE_E_UNSUPPORTED;
  const int tx = cdiv(vx.w, HT_X), ty = cdiv(vx.h, HT_Y);
  const size_t smem = sizeof(float) * (9 * (size_t)vx.c + HZ_N * 9);
  head_fwd_kernel<<<(unsigned)((long long)tx * ty * vx.n), 256, smem, as_stream(stream)>>>(vx, vy, w, bias, act, reflect, tx, ty);
  return launched();
}
extern "C" int segsde_head_dgrad_fused(const segsde_nhwc_t* dz, const float* w, const segsde_nhwc_t* dx, int reflect,
                                       void* stream) {
  if (!dz || !dx || !w) return SEGSDE_E_ARG;
  View vd = mk(dz), vx = mk(dx);
  if (!head_ok(vx, vd)) return SEGSDE_E_UNSUPPORTED;
  const int tx = cdiv(vx.w, HT_X), ty = cdiv(vx.h, HT_Y);
  const size_t smem = sizeof(float) * (9 * (size_t)vx.c + HT_N * 9);
  head_dgrad_kernel<<<(unsigned)((long long)tx * ty * vx.n), 256, smem, as_stream(stream)>>>(vd, vx, w, reflect, tx, ty);
  return launched();
}
extern "C" int segsde_head_wgrad_fused(const segsde_nhwc_t* x, const segsde_nhwc_t* dz, float* dw, int reflect,
                                       void* stream) {
  if (!x || !dz || !dw) return SEGSDE_E_ARG;
  View vx = mk(x), vd = mk(dz);
  if (!head_ok(vx, vd)) return SEGSDE_E_UNSUPPORTED;
  const int tx = cdiv(vx.w, HT_X), ty = cdiv(vx.h, HT_Y);
  const long long total = (long long)tx * ty * vx.n;
  const long long grid = total < 148 * 4 ? total : 148 * 4;
  head_wgrad_kernel<<<(unsigned)grid, 256, 0, as_stream(stream)>>>(vx, vd, dw, reflect, tx, ty, total);
  return launched();
}
