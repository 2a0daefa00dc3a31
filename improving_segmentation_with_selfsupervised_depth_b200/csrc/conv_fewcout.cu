// 1x1 convolutions with a FEW output channels on many pixels: the 19-class segmentation heads
// (joint_segmentation_depth_decoder.py:106-107 of the reference: 64 -> 19 at full resolution, 128 -> 19 at 1/4) and the
// like.  Cout is not a multiple of 32, so the tcgen05 family does not take them, and the tiled CUDA-core kernel
// (conv_simt.cu) ran them at ~3 TFLOP/s - 9 ms of the joint step.  These layers are HBM-bound (64 -> 19: 332 B per pixel
// for 2.4 kFLOP), so the kernels below are organised around the byte streams:
//   fprop : one thread per pixel, Cout accumulators in registers, the [Cin/4][CO][4] weight block broadcast from shared
//           memory, the CTA's 256 x Cout outputs staged in shared memory and written as one contiguous range;
//   dgrad : one thread per pixel and 32-channel chunk, the CTA's dy rows staged through shared memory;
//   wgrad : lane = input channel (mod 32), warp = pixel slab, Cout x Cin/32 accumulators per lane, dy rows broadcast from
//           shared memory; one atomicAdd per (co, ci) and CTA.
// All three are reached through segsde_conv2d_fwd / _dgrad / _wgrad (conv_simt.cu); fp32 math (no TF32 rounding).
#include "common.cuh"

namespace segsde {

constexpr int FC_T = 256;          // threads = pixels per CTA (fprop / dgrad)

template <int CO>                  // CO = Cout rounded up to a multiple of 4
__global__ void __launch_bounds__(FC_T) fewcout_fwd_kernel(View x, View y, const float* __restrict__ w,
                                                           const float* __restrict__ bias, int C, int Cout, int act,
                                                           long long P, int y_dense) {
  extern __shared__ __align__(16) float sm[];
  float* ws = sm;                              // [C/4][CO][4]
  float* os = sm + (size_t)C * CO;             // [FC_T][Cout] staging (dense outputs only)
  const int cq = C >> 2;
  for (int i = threadIdx.x; i < cq * CO * 4; i += FC_T) {
    const int k = i & 3, co = (i >> 2) % CO, c4 = (i >> 2) / CO;
    ws[i] = co < Cout ? __ldg(w + (long long)co * C + c4 * 4 + k) : 0.f;
  }
  __syncthreads();
  const long long p0 = (long long)blockIdx.x * FC_T;
  const long long p = p0 + threadIdx.x;
  const bool pv = p < P;
  float acc[CO];
#pragma unroll
  for (int j = 0; j < CO; ++j) acc[j] = 0.f;
  if (pv) {
    const int hw = y.h * y.w;
    const int n = (int)(p / hw), r = (int)(p - (long long)n * hw), hh = r / y.w, wq = r - hh * y.w;
    const float* xp = x.p + x.off(n, hh, wq);
#pragma unroll 2
    for (int c4 = 0; c4 < cq; ++c4) {
      const float4 v = __ldg(reinterpret_cast<const float4*>(xp) + c4);
      const float4* wr = reinterpret_cast<const float4*>(ws) + c4 * CO;
#pragma unroll
      for (int j = 0; j < CO; ++j) {
        const float4 q = wr[j];
        acc[j] = fmaf(v.x, q.x, fmaf(v.y, q.y, fmaf(v.z, q.z, fmaf(v.w, q.w, acc[j]))));
      }
    }
#pragma unroll
    for (int j = 0; j < CO; ++j)
      if (j < Cout) acc[j] = act_apply(acc[j] + (bias ? __ldg(bias + j) : 0.f), act);
    if (!y_dense) {
      float* yp = y.p + y.off(n, hh, wq);
#pragma unroll
      for (int j = 0; j < CO; ++j)
        if (j < Cout) yp[j] = acc[j];
    }
  }
  if (y_dense) {             // the CTA's outputs are one contiguous range of the tensor: coalesced write-out
#pragma unroll
    for (int j = 0; j < CO; ++j)
      if (j < Cout) os[threadIdx.x * Cout + j] = acc[j];
    __syncthreads();
    const long long rem = P - p0;
    const int cnt = (int)(rem < FC_T ? rem : FC_T) * Cout;
    float* dst = y.p + p0 * Cout;
    for (int i = threadIdx.x; i < cnt; i += FC_T) dst[i] = os[i];
  }
}

// dx[p, c] = sum_co dy[p, co] * w[co, c];  grid (pixel chunks, 32-channel chunks)
template <int CO>
__global__ void __launch_bounds__(FC_T) fewcout_dgrad_kernel(View dy, View dx, const float* __restrict__ w, int C, int Cout,
                                                             long long P, int dy_dense) {
  extern __shared__ __align__(16) float sm[];
  float* ws = sm;                              // [CO][32]
  float* ds = sm + CO * 32;                    // [FC_T][Cout]
  const int c0 = blockIdx.y * 32;
  for (int i = threadIdx.x; i < CO * 32; i += FC_T) {
    const int co = i >> 5, c = c0 + (i & 31);
    ws[i] = (co < Cout && c < C) ? __ldg(w + (long long)co * C + c) : 0.f;
  }
  const long long p0 = (long long)blockIdx.x * FC_T;
  const long long p = p0 + threadIdx.x;
  const bool pv = p < P;
  const int hw = dx.h * dx.w;
  int n = 0, hh = 0, wq = 0;
  if (pv) { n = (int)(p / hw); const int r = (int)(p - (long long)n * hw); hh = r / dx.w; wq = r - hh * dx.w; }
  if (dy_dense) {
    const long long rem = P - p0;
    const int cnt = (int)(rem < FC_T ? rem : FC_T) * Cout;
    const float* src = dy.p + p0 * Cout;
    for (int i = threadIdx.x; i < cnt; i += FC_T) ds[i] = __ldg(src + i);
  } else if (pv) {
    const float* gp = dy.p + dy.off(n, hh, wq);
    for (int j = 0; j < Cout; ++j) ds[threadIdx.x * Cout + j] = __ldg(gp + j);
  }
  __syncthreads();
  if (!pv) return;
  float acc[32];
#pragma unroll
  for (int k = 0; k < 32; ++k) acc[k] = 0.f;
  const float* g = ds + threadIdx.x * Cout;
  for (int co = 0; co < Cout; ++co) {
    const float d = g[co];
    const float4* wr = reinterpret_cast<const float4*>(ws + co * 32);
#pragma unroll
    for (int k4 = 0; k4 < 8; ++k4) {
      const float4 q = wr[k4];
      acc[k4 * 4 + 0] = fmaf(d, q.x, acc[k4 * 4 + 0]); acc[k4 * 4 + 1] = fmaf(d, q.y, acc[k4 * 4 + 1]);
      acc[k4 * 4 + 2] = fmaf(d, q.z, acc[k4 * 4 + 2]); acc[k4 * 4 + 3] = fmaf(d, q.w, acc[k4 * 4 + 3]);
    }
  }
  float* xp = dx.p + dx.off(n, hh, wq) + c0;
#pragma unroll
  for (int k4 = 0; k4 < 8; ++k4)
    if (c0 + k4 * 4 < C)
      *reinterpret_cast<float4*>(xp + k4 * 4) = make_float4(acc[k4 * 4], acc[k4 * 4 + 1], acc[k4 * 4 + 2], acc[k4 * 4 + 3]);
}

// dw[co, c] += sum_p dy[p, co] * x[p, c];  grid (pixel slabs, 64-channel chunks); lane owns channels c0 + lane, c0 + 32 + lane
constexpr int FW_PIX = 64;         // pixels staged per round and CTA
template <int CO>
__global__ void __launch_bounds__(FC_T) fewcout_wgrad_kernel(View x, View dy, float* __restrict__ dw, int C, int Cout,
                                                             long long P, long long slab, int dy_dense) {
  __shared__ __align__(16) float ds[FW_PIX * CO];          // [pixel][CO], zero-padded
  extern __shared__ float red_raw[];                       // [8][CO][65]
  float (*red)[CO][65] = reinterpret_cast<float (*)[CO][65]>(red_raw);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int c0 = blockIdx.y * 64;
  const bool v0 = c0 + lane < C, v1 = c0 + 32 + lane < C;
  const long long pbeg = (long long)blockIdx.x * slab, pend = min(P, pbeg + slab);
  const int hw = x.h * x.w;
  const bool x_flat = x.sh == (long long)x.w * x.sw && x.sn == (long long)x.h * x.sh;     // pixel p at p * sw
  float a0[CO], a1[CO];
#pragma unroll
  for (int j = 0; j < CO; ++j) { a0[j] = 0.f; a1[j] = 0.f; }
  for (long long q0 = pbeg; q0 < pend; q0 += FW_PIX) {
    const int np = (int)min((long long)FW_PIX, pend - q0);
    __syncthreads();
    for (int i = threadIdx.x; i < FW_PIX * CO; i += FC_T) {
      const int px = i / CO, co = i - px * CO;
      float v = 0.f;
      if (px < np && co < Cout) {
        if (dy_dense) v = __ldg(dy.p + (q0 + px) * Cout + co);
        else {
          const long long pp = q0 + px;
          const int n = (int)(pp / hw), r = (int)(pp - (long long)n * hw), hh = r / x.w;
          v = __ldg(dy.p + dy.off(n, hh, r - hh * x.w) + co);
        }
      }
      ds[i] = v;
    }
    __syncthreads();
    for (int px = warp; px < np; px += 8) {
      const long long pp = q0 + px;
      const float* xp;
      if (x_flat) xp = x.p + pp * x.sw + c0 + lane;
      else {
        const int n = (int)(pp / hw), r = (int)(pp - (long long)n * hw), hh = r / x.w;
        xp = x.p + x.off(n, hh, r - hh * x.w) + c0 + lane;
      }
      const float x0 = v0 ? __ldg(xp) : 0.f, x1 = v1 ? __ldg(xp + 32) : 0.f;
      const float4* g4 = reinterpret_cast<const float4*>(ds + px * CO);
#pragma unroll
      for (int j4 = 0; j4 < CO / 4; ++j4) {
        const float4 g = g4[j4];
        a0[j4 * 4 + 0] = fmaf(g.x, x0, a0[j4 * 4 + 0]); a1[j4 * 4 + 0] = fmaf(g.x, x1, a1[j4 * 4 + 0]);
        a0[j4 * 4 + 1] = fmaf(g.y, x0, a0[j4 * 4 + 1]); a1[j4 * 4 + 1] = fmaf(g.y, x1, a1[j4 * 4 + 1]);
        a0[j4 * 4 + 2] = fmaf(g.z, x0, a0[j4 * 4 + 2]); a1[j4 * 4 + 2] = fmaf(g.z, x1, a1[j4 * 4 + 2]);
        a0[j4 * 4 + 3] = fmaf(g.w, x0, a0[j4 * 4 + 3]); a1[j4 * 4 + 3] = fmaf(g.w, x1, a1[j4 * 4 + 3]);
      }
    }
  }
#pragma unroll
  for (int j = 0; j < CO; ++j) { red[warp][j][lane] = a0[j]; red[warp][j][32 + lane] = a1[j]; }
  __syncthreads();
  for (int i = threadIdx.x; i < CO * 64; i += FC_T) {
    const int co = i >> 6, c = i & 63;
    if (co >= Cout || c0 + c >= C) continue;
    float s = 0.f;
#pragma unroll
    for (int wv = 0; wv < 8; ++wv) s += red[wv][co][c];
    atomicAdd(dw + (long long)co * C + c0 + c, s);
  }
}

static bool dense_rows(const View& v) {        // pixel p of the flattened (n, h, w) index lives at p * c
  return v.sw == v.c && v.sh == (long long)v.w * v.sw && v.sn == (long long)v.h * v.sh;
}

#define FC_DISPATCH(CO_, ...)                  \
  switch (CO_) {                               \
    case 4: { constexpr int CO = 4; __VA_ARGS__; break; }    \
    case 8: { constexpr int CO = 8; __VA_ARGS__; break; }    \
    case 12: { constexpr int CO = 12; __VA_ARGS__; break; }  \
    case 16: { constexpr int CO = 16; __VA_ARGS__; break; }  \
    case 20: { constexpr int CO = 20; __VA_ARGS__; break; }  \
    case 24: { constexpr int CO = 24; __VA_ARGS__; break; }  \
    case 28: { constexpr int CO = 28; __VA_ARGS__; break; }  \
    default: { constexpr int CO = 32; __VA_ARGS__; break; }  \
  }

bool fewcout_ok(const View& x, const View& y, int kh, int kw, int stride, int pad, bool second_source, bool nchw, bool up1) {
  const long long P = (long long)y.n * y.h * y.w;
  return kh == 1 && kw == 1 && stride == 1 && pad == 0 && !second_source && !nchw && !up1 && y.c >= 2 && y.c <= 32 &&
         x.c % 4 == 0 && x.c >= 32 && x.c <= 512 && P >= 4096 && x.sw % 4 == 0 && x.sh % 4 == 0 && x.sn % 4 == 0 &&
         (reinterpret_cast<uintptr_t>(x.p) & 15) == 0 && (long long)y.h * y.w < (1LL << 31);
}

int fewcout_fwd(const View& x, const View& y, const float* w, const float* bias, int act, cudaStream_t st) {
  const long long P = (long long)y.n * y.h * y.w;
  const int C = x.c, Cout = y.c, co_pad = (Cout + 3) & ~3;
  const int dense = dense_rows(y) ? 1 : 0;
  const size_t smem = ((size_t)C * co_pad + (dense ? (size_t)FC_T * Cout : 0)) * sizeof(float);
  if (smem > 200 * 1024) return SEGSDE_E_UNSUPPORTED;
  FC_DISPATCH(co_pad, {
    if (smem > 48 * 1024) cudaFuncSetAttribute(fewcout_fwd_kernel<CO>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    fewcout_fwd_kernel<CO><<<(unsigned)cdiv(P, FC_T), FC_T, smem, st>>>(x, y, w, bias, C, Cout, act, P, dense);
  });
  return launched();
}

int fewcout_dgrad(const View& dy, const View& dx, const float* w, cudaStream_t st) {
  const long long P = (long long)dy.n * dy.h * dy.w;
  const int C = dx.c, Cout = dy.c, co_pad = (Cout + 3) & ~3;
  const int dense = dense_rows(dy) ? 1 : 0;
  const size_t smem = ((size_t)co_pad * 32 + (size_t)FC_T * Cout) * sizeof(float);
  dim3 grid((unsigned)cdiv(P, FC_T), (unsigned)cdiv(C, 32));
  FC_DISPATCH(co_pad, { fewcout_dgrad_kernel<CO><<<grid, FC_T, smem, st>>>(dy, dx, w, C, Cout, P, dense); });
  return launched();
}

int fewcout_wgrad(const View& x, const View& dy, float* dw, cudaStream_t st) {
  const long long P = (long long)dy.n * dy.h * dy.w;
  const int C = x.c, Cout = dy.c, co_pad = (Cout + 3) & ~3;
  const int cchunks = cdiv(C, 64);
  long long slabs = (148LL * 8) / cchunks; if (slabs < 1) slabs = 1;
  long long slab = cdiv(P, slabs); slab = cdiv(slab, (long long)FW_PIX) * FW_PIX;
  dim3 grid((unsigned)cdiv(P, slab), (unsigned)cchunks);
  const int dense = dense_rows(dy) ? 1 : 0;
  FC_DISPATCH(co_pad, {
    constexpr int smem = 8 * CO * 65 * (int)sizeof(float);
    if (smem > 40 * 1024) cudaFuncSetAttribute(fewcout_wgrad_kernel<CO>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    fewcout_wgrad_kernel<CO><<<grid, FC_T, smem, st>>>(x, dy, dw, C, Cout, P, slab, dense);
  });
  return launched();
}

}  // namespace segsde
