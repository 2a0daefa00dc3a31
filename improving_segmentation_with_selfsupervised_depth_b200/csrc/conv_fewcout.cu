// 1x1 convolutions with a FEW output channels on many pixels: the 19-class segmentation heads
// (joint_segmentation_depth_decoder.py:106-107 of the reference: 64 -> 19 at full resolution, 128 -> 19 at 1/4) and the
// like.  Cout is not a multiple of 32, so the tcgen05 family does not take them, and the tiled CUDA-core kernel
// (conv_simt.cu) ran them at ~3 TFLOP/s - 9 ms of the joint step.  These layers are HBM-bound (64 -> 19: 332 B per pixel
// for 2.4 kFLOP):
//   fprop : 2 pixels per thread, Cout accumulators each, the [Cin/4][CO][4] weight block broadcast from shared memory,
//           x rows through a two-stage cp.async ring, the CTA's outputs written as one contiguous range;
//   dgrad : one thread per pixel and 32-channel chunk, the CTA's dy rows staged through shared memory;
//   wgrad : lane = input channel (mod 32), warp = pixel slab, Cout x Cin/32 accumulators per lane, dy rows broadcast from
//           shared memory; one atomicAdd per (co, ci) and CTA.
// Measured (8 x 512 x 1024, 64 -> 19): fprop 0.49 ms, dgrad 0.57 ms, wgrad 0.57 ms (2.2 - 2.8 TB/s; the generic kernel:
// 3.5 / 1.9 / 3.0 ms).  All three are reached through segsde_conv2d_fwd / _dgrad / _wgrad (conv_simt.cu); fp32 math.
#include <cstdlib>
#include "common.cuh"

namespace segsde {

// Tile shape of fprop / dgrad: FP_T threads x FP_PPT pixels per thread, FP_CH channels per staged chunk.  The row pitch
// of a staged chunk is FP_CH + 4 floats (48 / 80 / 144 B): the LDS.128 of eight consecutive pixel-threads then fall into
// eight different 16-byte bank groups.

__device__ __forceinline__ const float* row_ptr(const View& v, long long p, bool flat, int hw) {
  if (flat) return v.p + p * v.sw;
  const int n = (int)(p / hw), r = (int)(p - (long long)n * hw), hh = r / v.w;
  return v.p + v.off(n, hh, r - hh * v.w);
}
__device__ __forceinline__ bool flat_rows(const View& v) {       // pixel p of the flattened (n, h, w) index at p * sw
  return v.sh == (long long)v.w * v.sw && v.sn == (long long)v.h * v.sh;
}
__device__ __forceinline__ void cp_async16(float* smem_dst, const float* gmem_src) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(d), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory"); }

// fprop.  A CTA owns 512 consecutive pixels; their x rows stream through a two-stage shared-memory ring 16 channels at a
// time (cp.async, 32 KB per stage, 64-byte row segments), each thread accumulates 4 pixels x Cout in registers.
template <int CO, int FP_PPT, int FP_CH, int FP_T>                  // CO = Cout rounded up to a multiple of 4
__global__ void __launch_bounds__(FP_T) fewcout_fwd_kernel(View x, View y, const float* __restrict__ w,
                                                           const float* __restrict__ bias, int C, int Cout, int act,
                                                           long long P, int y_dense) {
  constexpr int FP_TILE = FP_T * FP_PPT, FP_LD = FP_CH + 4, QPC = FP_CH / 4;
  extern __shared__ __align__(16) float sm[];
  float* ws = sm;                              // [C/4][CO][4]
  float* xs = sm + (size_t)C * CO;             // [2][FP_TILE][FP_LD]; reused as the [FP_TILE][Cout] output staging
  const int cq = C >> 2;
  for (int i = threadIdx.x; i < cq * CO * 4; i += FP_T) {
    const int k = i & 3, co = (i >> 2) % CO, c4 = (i >> 2) / CO;
    ws[i] = co < Cout ? __ldg(w + (long long)co * C + c4 * 4 + k) : 0.f;
  }
  const long long p0 = (long long)blockIdx.x * FP_TILE;
  const int hw = y.h * y.w;
  const bool xf = flat_rows(x);
  auto issue = [&](int c0, float* stage) {
#pragma unroll
    for (int k = 0; k < FP_TILE * QPC / FP_T; ++k) {
      const int idx = k * FP_T + threadIdx.x, px = idx / QPC, q = idx % QPC;
      float* dst = stage + px * FP_LD + q * 4;
      if (p0 + px < P && c0 + q * 4 < C) cp_async16(dst, row_ptr(x, p0 + px, xf, hw) + c0 + q * 4);
      else *reinterpret_cast<float4*>(dst) = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    cp_async_commit();
  };
  float acc[FP_PPT][CO];
#pragma unroll
  for (int i = 0; i < FP_PPT; ++i)
#pragma unroll
    for (int j = 0; j < CO; ++j) acc[i][j] = 0.f;
  const int nchunks = (C + FP_CH - 1) / FP_CH;
  issue(0, xs);
  for (int ck = 0; ck < nchunks; ++ck) {
    float* cur = xs + (ck & 1) * (FP_TILE * FP_LD);
    if (ck + 1 < nchunks) { issue((ck + 1) * FP_CH, xs + ((ck + 1) & 1) * (FP_TILE * FP_LD)); cp_async_wait<1>(); }
    else cp_async_wait<0>();
    __syncthreads();
    const int c0 = ck * FP_CH;
    const int nq = min(FP_CH / 4, (C - c0) >> 2);
    for (int q = 0; q < nq; ++q) {
      float4 xv[FP_PPT];
#pragma unroll
      for (int i = 0; i < FP_PPT; ++i)
        xv[i] = *reinterpret_cast<const float4*>(cur + (threadIdx.x + FP_T * i) * FP_LD + q * 4);
      const float4* wr = reinterpret_cast<const float4*>(ws) + ((c0 >> 2) + q) * CO;
#pragma unroll
      for (int j = 0; j < CO; ++j) {
        const float4 t = wr[j];
#pragma unroll
        for (int i = 0; i < FP_PPT; ++i)
          acc[i][j] = fmaf(xv[i].x, t.x, fmaf(xv[i].y, t.y, fmaf(xv[i].z, t.z, fmaf(xv[i].w, t.w, acc[i][j]))));
      }
    }
    __syncthreads();                  // the stage is free for chunk ck + 2
  }
#pragma unroll
  for (int i = 0; i < FP_PPT; ++i)
#pragma unroll
    for (int j = 0; j < CO; ++j)
      if (j < Cout) acc[i][j] = act_apply(acc[i][j] + (bias ? __ldg(bias + j) : 0.f), act);
  if (!y_dense) {
#pragma unroll
    for (int i = 0; i < FP_PPT; ++i) {
      const long long p = p0 + threadIdx.x + FP_T * i;
      if (p >= P) continue;
      float* yp = const_cast<float*>(row_ptr(y, p, false, hw));
#pragma unroll
      for (int j = 0; j < CO; ++j)
        if (j < Cout) yp[j] = acc[i][j];
    }
    return;
  }
  // the CTA's outputs are one contiguous range of the tensor: coalesced write-out through the (now idle) ring
#pragma unroll
  for (int i = 0; i < FP_PPT; ++i)
#pragma unroll
    for (int j = 0; j < CO; ++j)
      if (j < Cout) xs[(threadIdx.x + FP_T * i) * Cout + j] = acc[i][j];
  __syncthreads();
  const long long rem = P - p0;
  const int cnt = (int)(rem < FP_TILE ? rem : FP_TILE) * Cout;
  float* dst = y.p + p0 * Cout;
  for (int i = threadIdx.x; i < cnt; i += FP_T) dst[i] = xs[i];
}

constexpr int FC_T = 256;
constexpr int FW_PIX = 64;         // wgrad: pixels staged per round and CTA

// dgrad / wgrad: the many-small-CTA shapes below measured faster than register-tiled 4-pixels-per-thread forms (0.57 vs
// 0.87+ ms dgrad, 0.57 vs 0.66 ms wgrad): these passes are latency-bound, so resident warps count more than LDS traffic.
constexpr int FC_LD = 36;
// dgrad: one thread per pixel, 32 channels at a time; dy rows staged once per CTA, dx chunk leaves as 128-byte row segments
template <int CO>
__global__ void __launch_bounds__(FC_T) fewcout_dgrad_kernel(View dy, View dx, const float* __restrict__ w, int C, int Cout,
                                                              long long P, int dy_dense) {
  extern __shared__ __align__(16) float sm[];
  float* ws = sm;                              // [CO][C]
  float* ds = ws + (size_t)CO * C;             // [FC_T][Cout]
  float* xs = ds + FC_T * Cout + ((4 - ((FC_T * Cout) & 3)) & 3);      // [FC_T][FC_LD], 16-byte aligned
  for (int i = threadIdx.x; i < CO * C; i += FC_T) {
    const int co = i / C;
    ws[i] = co < Cout ? __ldg(w + i) : 0.f;
  }
  const long long p0 = (long long)blockIdx.x * FC_T;
  const long long p = p0 + threadIdx.x;
  const bool pv = p < P;
  const int hw = dx.h * dx.w;
  const bool xf = flat_rows(dx);
  if (dy_dense) {
    const long long rem = P - p0;
    const int cnt = (int)(rem < FC_T ? rem : FC_T) * Cout;
    const float* src = dy.p + p0 * Cout;
#pragma unroll 4
    for (int i = threadIdx.x; i < FC_T * Cout; i += FC_T) ds[i] = i < cnt ? __ldg(src + i) : 0.f;
  } else {
    const float* gp = pv ? row_ptr(dy, p, false, hw) : nullptr;
    for (int j = 0; j < Cout; ++j) ds[threadIdx.x * Cout + j] = pv ? __ldg(gp + j) : 0.f;
  }
  __syncthreads();
  const float* g = ds + threadIdx.x * Cout;
  for (int c0 = 0; c0 < C; c0 += 32) {
    float acc[32];
#pragma unroll
    for (int k = 0; k < 32; ++k) acc[k] = 0.f;
    const int nq = min(8, (C - c0) >> 2);
    for (int co = 0; co < Cout; ++co) {
      const float d = g[co];
      const float4* wr = reinterpret_cast<const float4*>(ws + (size_t)co * C + c0);
#pragma unroll
      for (int k4 = 0; k4 < 8; ++k4) {
        if (k4 < nq) {
          const float4 q = wr[k4];
          acc[k4 * 4 + 0] = fmaf(d, q.x, acc[k4 * 4 + 0]); acc[k4 * 4 + 1] = fmaf(d, q.y, acc[k4 * 4 + 1]);
          acc[k4 * 4 + 2] = fmaf(d, q.z, acc[k4 * 4 + 2]); acc[k4 * 4 + 3] = fmaf(d, q.w, acc[k4 * 4 + 3]);
        }
      }
    }
    __syncthreads();                  // the previous chunk has left the staging tile
#pragma unroll
    for (int k4 = 0; k4 < 8; ++k4)
      *reinterpret_cast<float4*>(xs + threadIdx.x * FC_LD + k4 * 4) =
          make_float4(acc[k4 * 4], acc[k4 * 4 + 1], acc[k4 * 4 + 2], acc[k4 * 4 + 3]);
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int idx = k * FC_T + threadIdx.x, px = idx >> 3, q = idx & 7;
      if (p0 + px < P && c0 + q * 4 < C)
        *reinterpret_cast<float4*>(const_cast<float*>(row_ptr(dx, p0 + px, xf, hw)) + c0 + q * 4) =
            *reinterpret_cast<const float4*>(xs + px * FC_LD + q * 4);
    }
  }
}

// wgrad: lane owns channels c0 + lane and c0 + 32 + lane (128-byte coalesced row segments), a warp takes 8 of the 64
// pixels of a round and issues all 16 of its x loads before the first FMA; the next round's dy rows are prefetched
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
  const bool xf = flat_rows(x);
  constexpr int NG = (FW_PIX * CO + FC_T - 1) / FC_T;
  float gpre[NG];
  auto fetch_dy = [&](long long q0) {
    const int np = (int)min((long long)FW_PIX, pend - q0);
#pragma unroll
    for (int k = 0; k < NG; ++k) {
      const int i = k * FC_T + threadIdx.x;
      const int px = i / CO, co = i - px * CO;
      float v = 0.f;
      if (i < FW_PIX * CO && px < np && co < Cout)
        v = dy_dense ? __ldg(dy.p + (q0 + px) * Cout + co) : __ldg(row_ptr(dy, q0 + px, false, hw) + co);
      gpre[k] = v;
    }
  };
  float a0[CO], a1[CO];
#pragma unroll
  for (int j = 0; j < CO; ++j) { a0[j] = 0.f; a1[j] = 0.f; }
  if (pbeg < pend) fetch_dy(pbeg);
  for (long long q0 = pbeg; q0 < pend; q0 += FW_PIX) {
    const int np = (int)min((long long)FW_PIX, pend - q0);
    __syncthreads();
#pragma unroll
    for (int k = 0; k < NG; ++k) {
      const int i = k * FC_T + threadIdx.x;
      if (i < FW_PIX * CO) ds[i] = gpre[k];
    }
    __syncthreads();
    float x0[8], x1[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int px = warp + 8 * i;
      x0[i] = 0.f; x1[i] = 0.f;
      if (px < np) {
        const float* xp = row_ptr(x, q0 + px, xf, hw) + c0 + lane;
        if (v0) x0[i] = __ldg(xp);
        if (v1) x1[i] = __ldg(xp + 32);
      }
    }
    if (q0 + FW_PIX < pend) fetch_dy(q0 + FW_PIX);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float4* g4 = reinterpret_cast<const float4*>(ds + (warp + 8 * i) * CO);      // zero rows beyond np
#pragma unroll
      for (int j4 = 0; j4 < CO / 4; ++j4) {
        const float4 g = g4[j4];
        a0[j4 * 4 + 0] = fmaf(g.x, x0[i], a0[j4 * 4 + 0]); a1[j4 * 4 + 0] = fmaf(g.x, x1[i], a1[j4 * 4 + 0]);
        a0[j4 * 4 + 1] = fmaf(g.y, x0[i], a0[j4 * 4 + 1]); a1[j4 * 4 + 1] = fmaf(g.y, x1[i], a1[j4 * 4 + 1]);
        a0[j4 * 4 + 2] = fmaf(g.z, x0[i], a0[j4 * 4 + 2]); a1[j4 * 4 + 2] = fmaf(g.z, x1[i], a1[j4 * 4 + 2]);
        a0[j4 * 4 + 3] = fmaf(g.w, x0[i], a0[j4 * 4 + 3]); a1[j4 * 4 + 3] = fmaf(g.w, x1[i], a1[j4 * 4 + 3]);
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

// fprop tile: 2 pixels per thread, 8-channel chunks, 256 threads (the best of five shapes measured on B200:
// {4px,16ch,128thr} 1.18 ms, {4,8,128} 0.93, {2,16,256} 0.53, {2,8,256} 0.49, {1,32,256} 0.70 for 64 -> 19 at 8x512x1024)
int fewcout_fwd(const View& x, const View& y, const float* w, const float* bias, int act, cudaStream_t st) {
  constexpr int PPT = 2, CH = 8, NT = 256, TILE = NT * PPT, LD = CH + 4;
  const long long P = (long long)y.n * y.h * y.w;
  const int C = x.c, Cout = y.c, co_pad = (Cout + 3) & ~3;
  const int dense = dense_rows(y) ? 1 : 0;
  const size_t ring = (size_t)2 * TILE * LD, stage_out = (size_t)TILE * Cout;
  const size_t smem = ((size_t)C * co_pad + (ring > stage_out ? ring : stage_out)) * sizeof(float);
  if (smem > 200 * 1024) return SEGSDE_E_UNSUPPORTED;
  FC_DISPATCH(co_pad, {
    auto k = fewcout_fwd_kernel<CO, PPT, CH, NT>;
    if (smem > 48 * 1024) cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    k<<<(unsigned)cdiv(P, TILE), NT, smem, st>>>(x, y, w, bias, C, Cout, act, P, dense);
  });
  return launched();
}

int fewcout_dgrad(const View& dy, const View& dx, const float* w, cudaStream_t st) {
  const long long P = (long long)dy.n * dy.h * dy.w;
  const int C = dx.c, Cout = dy.c, co_pad = (Cout + 3) & ~3;
  const int dense = dense_rows(dy) ? 1 : 0;
  const size_t smem = ((size_t)co_pad * C + (size_t)FC_T * Cout + 4 + (size_t)FC_T * FC_LD) * sizeof(float);
  if (smem > 200 * 1024) return SEGSDE_E_UNSUPPORTED;
  FC_DISPATCH(co_pad, {
    if (smem > 48 * 1024) cudaFuncSetAttribute(fewcout_dgrad_kernel<CO>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    fewcout_dgrad_kernel<CO><<<(unsigned)cdiv(P, FC_T), FC_T, smem, st>>>(dy, dx, w, C, Cout, P, dense);
  });
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
