// Helpers around the tensor-core convolution family: materialised reflection padding (+ nearest x2
// upsampling) of a conv input and its adjoint, the transposed / tap-flipped weight matrix that turns dgrad
// into an fprop, and the fused activation-backward + bias-gradient pass.
#include "common.cuh"

namespace segsde {

// y[n, hp, wp, c] = x[n, refl(hp - pad) >> up, refl(wp - pad) >> up, c]
// grid (chunks of one padded row, hp, n): 32-bit index math only - the 64-bit div/mod chain of a flat index made this
// copy issue-bound at 3.3 TB/s (profiles/r2_hot_kernels.md).
__global__ void __launch_bounds__(256) pad_prep_kernel(View x, View y, int up, int pad, int Hc, int Wc, int cq) {
  const int i = blockIdx.x * 256 + threadIdx.x;
  if (i >= y.w * cq) return;
  const int wp = i / cq, c = (i - wp * cq) * 4;
  const int hp = blockIdx.y, n = blockIdx.z;
  const int h = reflect_idx(hp - pad, Hc) >> up, w = reflect_idx(wp - pad, Wc) >> up;
  *reinterpret_cast<float4*>(y.p + y.off(n, hp, wp) + c) =
      *reinterpret_cast<const float4*>(x.p + x.off(n, h, w) + c);
}

// adjoint of pad_prep: dx[n,h,w,c] = sum over the (1<<up)^2 fine positions of the sum over their reflect preimages
// grid (chunks of one dx row, h, n)
__global__ void __launch_bounds__(256) pad_fold_kernel(View dyp, View dx, int up, int pad, int Hc, int Wc, int cq) {
  const int i = blockIdx.x * 256 + threadIdx.x;
  if (i >= dx.w * cq) return;
  const int w = i / cq, c = (i - w * cq) * 4;
  const int h = blockIdx.y, n = blockIdx.z;
  const int f = 1 << up;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int a = 0; a < f; ++a) {
    const int hh = h * f + a;
    int rows[3]; int nr = 0;
    rows[nr++] = hh;
    if (hh >= 1 && hh <= pad) rows[nr++] = -hh;
    if (hh >= Hc - 1 - pad && hh <= Hc - 2) rows[nr++] = 2 * (Hc - 1) - hh;
    for (int b = 0; b < f; ++b) {
      const int ww = w * f + b;
      int cols[3]; int nc = 0;
      cols[nc++] = ww;
      if (ww >= 1 && ww <= pad) cols[nc++] = -ww;
      if (ww >= Wc - 1 - pad && ww <= Wc - 2) cols[nc++] = 2 * (Wc - 1) - ww;
      for (int ii = 0; ii < nr; ++ii)
        for (int j = 0; j < nc; ++j) {
          const float4 v = *reinterpret_cast<const float4*>(dyp.p + dyp.off(n, rows[ii] + pad, cols[j] + pad) + c);
          acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
        }
    }
  }
  *reinterpret_cast<float4*>(dx.p + dx.off(n, h, w) + c) = acc;
}

// wt[ci - c0][kh-1-r][kw-1-s][co] = w[co][r][s][ci]
// Per tap a [Cout x cn] -> [cn x Cout] transpose through a 32 x 33 shared-memory tile: reads coalesced along ci, writes
// coalesced along co (the flat one-thread-per-element form gathered with a kh*kw*Ctot stride: ~5 ms/step of the
// DepthMix configuration, whose encoder trains).  grid (ci tiles, co tiles, taps), block (32, 8)
__global__ void __launch_bounds__(256) weight_tflip_kernel(const float* __restrict__ w, float* __restrict__ wt, int Cout, int kh,
                                                           int kw, int Ctot, int c0, int cn) {
  __shared__ float tile[32][33];
  const int taps = kh * kw, tap = blockIdx.z;
  const int r = tap / kw, s_ = tap - r * kw;
  const int tap_o = (kh - 1 - r) * kw + (kw - 1 - s_);
  const int ci0 = blockIdx.x * 32, co0 = blockIdx.y * 32;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int co = co0 + threadIdx.y + 8 * j, ci = ci0 + threadIdx.x;
    if (co < Cout && ci < cn) tile[threadIdx.y + 8 * j][threadIdx.x] = w[((long long)co * taps + tap) * Ctot + c0 + ci];
  }
  __syncthreads();
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int ci = ci0 + threadIdx.y + 8 * j, co = co0 + threadIdx.x;
    if (co < Cout && ci < cn) wt[((long long)ci * taps + tap_o) * Cout + co] = tile[threadIdx.x][threadIdx.y + 8 * j];
  }
}

// The same pass for FEW channels (the 1-channel disparity heads, the 12-channel pose output): one thread per pixel, the
// channels in a register loop, block-level reduction of the bias sums.  The 32-channels-per-warp kernel above leaves
// 31 of 32 lanes idle at C = 1 (ncu, round 2: 0.24 ms per head at 47 GB/s — 1.4 ms of a 57 ms step).
template <int MAXC>
__global__ void __launch_bounds__(256) act_bwd_bias_fewc_kernel(View y, View dy, View dz, int act, float* __restrict__ dbias) {
  const long long P = (long long)dy.n * dy.h * dy.w;
  const int C = dy.c;
  float s[MAXC];
#pragma unroll
  for (int c = 0; c < MAXC; ++c) s[c] = 0.f;
  for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < P; p += (long long)gridDim.x * blockDim.x) {
    const int w = (int)(p % dy.w); const long long q = p / dy.w;
    const int h = (int)(q % dy.h), n = (int)(q / dy.h);
    const long long od = dy.off(n, h, w), oy = y.p ? y.off(n, h, w) : 0, oz = dz.p ? dz.off(n, h, w) : 0;
#pragma unroll
    for (int c = 0; c < MAXC; ++c) {
      if (c >= C) break;
      float g = dy.p[od + c];
      if (act != SEGSDE_ACT_NONE) g *= act_grad_from_out(y.p[oy + c], act);
      if (dz.p) dz.p[oz + c] = g;
      s[c] += g;
    }
  }
  if (!dbias) return;
  __shared__ float red[8][MAXC];
#pragma unroll
  for (int c = 0; c < MAXC; ++c) {
    const float v = warp_sum(s[c]);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5][c] = v;
  }
  __syncthreads();
  if (threadIdx.x < C) {
    float t = 0.f;
    for (int i = 0; i < 8; ++i) t += red[i][threadIdx.x];
    atomicAdd(dbias + threadIdx.x, t);
  }
}

// Phase (a, b) of the dgrad of a 3x3 / stride-2 / pad-1 convolution: dx[2i+a, 2j+b] is a (1+a) x (1+b)-tap stride-1
// convolution of dy — row taps: a = 0 -> {w[1] at offset 0}; a = 1 -> {w[2] at offset 0, w[0] at offset +1}, columns alike —
// instead of a 9-tap convolution over the zero-stuffed dy (4x the MACs).  wt[ci - c0][th][tw][co] = w[co][r(a,th)][s(b,tw)][ci].
__global__ void weight_phase_s2_kernel(const float* __restrict__ w, float* __restrict__ wt, int Cout, int Ctot, int c0, int cn,
                                       int a, int b) {
  const int nth = 1 + a, ntw = 1 + b;
  const long long total = (long long)cn * nth * ntw * Cout;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int co = (int)(idx % Cout); long long q = idx / Cout;
  const int tw = (int)(q % ntw); q /= ntw;
  const int th = (int)(q % nth); const int ci = (int)(q / nth);
  const int r = a == 0 ? 1 : (th == 0 ? 2 : 0), s = b == 0 ? 1 : (tw == 0 ? 2 : 0);
  wt[idx] = w[(((long long)co * 3 + r) * 3 + s) * Ctot + c0 + ci];
}

// ---- nearest x2 upsampling + ReflectionPad2d(1) + 3x3 convolution as four 2x2 "phase" convolutions on the low-res input ----
// Output pixel (2i+a, 2j+b) sees the upsampled rows 2i+a-1 .. 2i+a+1, i.e. the low-res rows {i-1, i, i} (a = 0) or
// {i, i, i+1} (a = 1): a 2-tap convolution with pre-summed weights  a=0: {w0, w1+w2},  a=1: {w0+w1, w2}  (columns alike),
// and the reflection of the UPSAMPLED image at the border is a REPLICATION of the low-res one.  9/4 of the MACs, the
// upsampled + padded copy (4x the low-res bytes) is never written, its gradient never folded back.
//
// xp[n, u, v, c] = x[n, clamp(u - 1), clamp(v - 1), c]   (replicate padding by 1)
__global__ void __launch_bounds__(256) pad_replicate_kernel(View x, View y, int cq) {
  const int i = blockIdx.x * 256 + threadIdx.x;
  if (i >= y.w * cq) return;
  const int v = i / cq, c = (i - v * cq) * 4;
  const int u = blockIdx.y, n = blockIdx.z;
  const int h = min(max(u - 1, 0), x.h - 1), w = min(max(v - 1, 0), x.w - 1);
  *reinterpret_cast<float4*>(y.p + y.off(n, u, v) + c) = *reinterpret_cast<const float4*>(x.p + x.off(n, h, w) + c);
}
// wp[co][dr][ds][ci] = sum_{r in R(a,dr)} sum_{s in R(b,ds)} w[co][r][s][c0 + ci],  R(0,0)={0}, R(0,1)={1,2}, R(1,0)={0,1}, R(1,1)={2}
__device__ __forceinline__ void phase_taps(int a, int d, int& lo, int& hi) {
  if (a == 0) { lo = d == 0 ? 0 : 1; hi = d == 0 ? 0 : 2; } else { lo = d == 0 ? 0 : 2; hi = d == 0 ? 1 : 2; }
}
__global__ void weight_phase_up_kernel(const float* __restrict__ w, float* __restrict__ wp, int Cout, int Ctot, int c0, int cn,
                                       int a, int b) {
  const long long total = (long long)Cout * 4 * cn;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int ci = (int)(idx % cn); long long q = idx / cn;
  const int ds = (int)(q % 2); q /= 2;
  const int dr = (int)(q % 2); const int co = (int)(q / 2);
  int r0, r1, s0, s1;
  phase_taps(a, dr, r0, r1); phase_taps(b, ds, s0, s1);
  float acc = 0.f;
  for (int r = r0; r <= r1; ++r)
    for (int s2 = s0; s2 <= s1; ++s2) acc += w[(((long long)co * 3 + r) * 3 + s2) * Ctot + c0 + ci];
  wp[idx] = acc;
}
// adjoint of the pre-summation: dw[co][r][s][c0+ci] (+)= sum over the phases (a,b) and taps (dr,ds) with r in R(a,dr), s in R(b,ds)
// of dwp[a][b][co][dr][ds][ci]      (dwp: [2][2][Cout][2][2][cn])
__global__ void weight_phase_up_fold_kernel(const float* __restrict__ dwp, float* __restrict__ dw, int Cout, int Ctot, int c0,
                                            int cn) {
  const long long total = (long long)Cout * 9 * cn;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int ci = (int)(idx % cn); long long q = idx / cn;
  const int s = (int)(q % 3); q /= 3;
  const int r = (int)(q % 3); const int co = (int)(q / 3);
  float acc = 0.f;
  for (int a = 0; a < 2; ++a)
    for (int dr = 0; dr < 2; ++dr) {
      int r0, r1; phase_taps(a, dr, r0, r1);
      if (r < r0 || r > r1) continue;
      for (int b = 0; b < 2; ++b)
        for (int ds = 0; ds < 2; ++ds) {
          int s0, s1; phase_taps(b, ds, s0, s1);
          if (s < s0 || s > s1) continue;
          acc += dwp[(((((long long)(a * 2 + b) * Cout + co) * 2 + dr) * 2 + ds)) * cn + ci];
        }
    }
  dw[(((long long)co * 3 + r) * 3 + s) * Ctot + c0 + ci] += acc;
}
// dx[n,i,j,c] = sum over the replicate preimages (u,v) of (i,j) in xp and the phases (a,b) of g_ab[n, u-a, v-b, c]
// (g: [2][2] buffers of shape [n, h+1, w+1, c], the gradients w.r.t. the four phase views xp[:, a:a+h+1, b:b+w+1])
struct PhaseG { const float* p[4]; };
__global__ void phase_up_fold_kernel(PhaseG g, View dx) {
  const int cq = dx.c / 4;
  const long long total = (long long)dx.n * dx.h * dx.w * cq;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int c = (int)(idx % cq) * 4; long long q = idx / cq;
  const int j = (int)(q % dx.w); q /= dx.w;
  const int i = (int)(q % dx.h); const int n = (int)(q / dx.h);
  const int H1 = dx.h + 1, W1 = dx.w + 1;
  int us[3], vs[3]; int nu = 0, nv = 0;
  us[nu++] = i + 1; if (i == 0) us[nu++] = 0; if (i == dx.h - 1) us[nu++] = dx.h + 1;
  vs[nv++] = j + 1; if (j == 0) vs[nv++] = 0; if (j == dx.w - 1) vs[nv++] = dx.w + 1;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int iu = 0; iu < nu; ++iu)
    for (int iv = 0; iv < nv; ++iv)
#pragma unroll
      for (int a = 0; a < 2; ++a) {
        const int uu = us[iu] - a;
        if (uu < 0 || uu >= H1) continue;
#pragma unroll
        for (int b = 0; b < 2; ++b) {
          const int vv = vs[iv] - b;
          if (vv < 0 || vv >= W1) continue;
          const float4 t = *reinterpret_cast<const float4*>(g.p[a * 2 + b] + (((long long)n * H1 + uu) * W1 + vv) * dx.c + c);
          acc.x += t.x; acc.y += t.y; acc.z += t.z; acc.w += t.w;
        }
      }
  *reinterpret_cast<float4*>(dx.p + dx.off(n, i, j) + c) = acc;
}

// dz = dy * act'(y) ; dbias[c] += sum dz   (block = 32 channels x 8 pixel lanes, like the BN reductions)
__global__ void __launch_bounds__(256) act_bwd_bias_kernel(View y, View dy, View dz, int act, float* __restrict__ dbias,
                                                           long long slab) {
  const int c = blockIdx.x * 32 + threadIdx.x;
  const long long P = (long long)dy.n * dy.h * dy.w;
  const long long pbeg = (long long)blockIdx.y * slab, pend = min(P, pbeg + slab);
  float s = 0.f;
  if (c < dy.c) {
    for (long long p = pbeg + threadIdx.y; p < pend; p += 8) {
      const int w = (int)(p % dy.w); const long long q = p / dy.w;
      const int h = (int)(q % dy.h), n = (int)(q / dy.h);
      float g = dy.p[dy.off(n, h, w) + c];
      if (act != SEGSDE_ACT_NONE) g *= act_grad_from_out(y.p[y.off(n, h, w) + c], act);
      if (dz.p) dz.p[dz.off(n, h, w) + c] = g;
      s += g;
    }
  }
  __shared__ float sh[8][32];
  sh[threadIdx.y][threadIdx.x] = s;
  __syncthreads();
  if (dbias && threadIdx.y == 0 && c < dy.c) {
    float t = 0.f;
    for (int i = 0; i < 8; ++i) t += sh[i][threadIdx.x];
    atomicAdd(dbias + c, t);
  }
}

}  // namespace segsde
using namespace segsde;

extern "C" int segsde_pad_prep(const segsde_nhwc_t* x, const segsde_nhwc_t* y, int up, int pad, void* stream) {
  if (!x || !y || !x->ptr || !y->ptr || pad < 0 || (up != 0 && up != 1)) return SEGSDE_E_ARG;
  View vx = mk(x), vy = mk(y);
  const int Hc = vx.h << up, Wc = vx.w << up;
  if (vy.h != Hc + 2 * pad || vy.w != Wc + 2 * pad || vy.c != vx.c || vy.n != vx.n || pad >= Hc || pad >= Wc) return SEGSDE_E_ARG;
  if (!vec4_ok(vx) || !vec4_ok(vy)) return SEGSDE_E_ALIGN;
  if (vy.h > 65535 || vy.n > 65535) return SEGSDE_E_ARG;
  const int cq = vy.c / 4;
  pad_prep_kernel<<<dim3(cdiv(vy.w * cq, 256), vy.h, vy.n), 256, 0, as_stream(stream)>>>(vx, vy, up, pad, Hc, Wc, cq);
  return launched();
}
extern "C" int segsde_pad_fold(const segsde_nhwc_t* dyp, const segsde_nhwc_t* dx, int up, int pad, void* stream) {
  if (!dyp || !dx || !dyp->ptr || !dx->ptr || pad < 0 || (up != 0 && up != 1)) return SEGSDE_E_ARG;
  View vd = mk(dyp), vx = mk(dx);
  const int Hc = vx.h << up, Wc = vx.w << up;
  if (vd.h != Hc + 2 * pad || vd.w != Wc + 2 * pad || vd.c != vx.c || vd.n != vx.n) return SEGSDE_E_ARG;
  if (!vec4_ok(vx) || !vec4_ok(vd)) return SEGSDE_E_ALIGN;
  if (vx.h > 65535 || vx.n > 65535) return SEGSDE_E_ARG;
  const int cq = vx.c / 4;
  pad_fold_kernel<<<dim3(cdiv(vx.w * cq, 256), vx.h, vx.n), 256, 0, as_stream(stream)>>>(vd, vx, up, pad, Hc, Wc, cq);
  return launched();
}
extern "C" int segsde_weight_transpose_flip(const float* w, float* wt, int cout, int kh, int kw, int ctot,
                                            int c_begin, int c_count, void* stream) {
  if (!w || !wt || cout < 1 || kh < 1 || kw < 1 || c_begin < 0 || c_count < 1 || c_begin + c_count > ctot) return SEGSDE_E_ARG;
  if (kh * kw > 65535 || cdiv(cout, 32) > 65535) return SEGSDE_E_ARG;
  weight_tflip_kernel<<<dim3(cdiv(c_count, 32), cdiv(cout, 32), kh * kw), dim3(32, 8), 0, as_stream(stream)>>>(
      w, wt, cout, kh, kw, ctot, c_begin, c_count);
  return launched();
}
extern "C" int segsde_pad_replicate(const segsde_nhwc_t* x, const segsde_nhwc_t* y, void* stream) {
  if (!x || !y || !x->ptr || !y->ptr) return SEGSDE_E_ARG;
  View vx = mk(x), vy = mk(y);
  if (vy.h != vx.h + 2 || vy.w != vx.w + 2 || vy.c != vx.c || vy.n != vx.n) return SEGSDE_E_ARG;
  if (!vec4_ok(vx) || !vec4_ok(vy)) return SEGSDE_E_ALIGN;
  if (vy.h > 65535 || vy.n > 65535) return SEGSDE_E_ARG;
  const int cq = vy.c / 4;
  pad_replicate_kernel<<<dim3(cdiv(vy.w * cq, 256), vy.h, vy.n), 256, 0, as_stream(stream)>>>(vx, vy, cq);
  return launched();
}
extern "C" int segsde_weight_phase_up(const float* w, float* wp, int cout, int ctot, int c_begin, int c_count, int a, int b,
                                     void* stream) {
  if (!w || !wp || cout < 1 || c_begin < 0 || c_count < 1 || c_begin + c_count > ctot || a < 0 || a > 1 || b < 0 || b > 1)
    return SEGSDE_E_ARG;
  const long long total = (long long)cout * 4 * c_count;
  weight_phase_up_kernel<<<cdiv(total, 256), 256, 0, as_stream(stream)>>>(w, wp, cout, ctot, c_begin, c_count, a, b);
  return launched();
}
extern "C" int segsde_weight_phase_up_fold(const float* dwp, float* dw, int cout, int ctot, int c_begin, int c_count,
                                          void* stream) {
  if (!dwp || !dw || cout < 1 || c_begin < 0 || c_count < 1 || c_begin + c_count > ctot) return SEGSDE_E_ARG;
  const long long total = (long long)cout * 9 * c_count;
  weight_phase_up_fold_kernel<<<cdiv(total, 256), 256, 0, as_stream(stream)>>>(dwp, dw, cout, ctot, c_begin, c_count);
  return launched();
}
extern "C" int segsde_phase_up_fold(const float* g00, const float* g01, const float* g10, const float* g11,
                                   const segsde_nhwc_t* dx, void* stream) {
  if (!g00 || !g01 || !g10 || !g11 || !dx || !dx->ptr) return SEGSDE_E_ARG;
  View vx = mk(dx);
  if (!vec4_ok(vx)) return SEGSDE_E_ALIGN;
  PhaseG g; g.p[0] = g00; g.p[1] = g01; g.p[2] = g10; g.p[3] = g11;
  const long long total = (long long)vx.n * vx.h * vx.w * (vx.c / 4);
  phase_up_fold_kernel<<<cdiv(total, 256), 256, 0, as_stream(stream)>>>(g, vx);
  return launched();
}
extern "C" int segsde_weight_phase_s2(const float* w, float* wt, int cout, int ctot, int c_begin, int c_count, int a, int b,
                                     void* stream) {
  if (!w || !wt || cout < 1 || c_begin < 0 || c_count < 1 || c_begin + c_count > ctot || a < 0 || a > 1 || b < 0 || b > 1)
    return SEGSDE_E_ARG;
  const long long total = (long long)c_count * (1 + a) * (1 + b) * cout;
  weight_phase_s2_kernel<<<cdiv(total, 256), 256, 0, as_stream(stream)>>>(w, wt, cout, ctot, c_begin, c_count, a, b);
  return launched();
}
extern "C" int segsde_act_bwd_bias(const segsde_nhwc_t* y, const segsde_nhwc_t* dy, const segsde_nhwc_t* dz, int act,
                                   float* dbias, void* stream) {
  if (!dy || !dy->ptr) return SEGSDE_E_ARG;
  View vy = mk(y), vd = mk(dy), vz = mk(dz);
  if (act != SEGSDE_ACT_NONE && (!vy.p || !same_shape(vy, vd))) return SEGSDE_E_ARG;
  if (vz.p && !same_shape(vz, vd)) return SEGSDE_E_ARG;
  if (fast_reduce_ok(vd) && (act == SEGSDE_ACT_NONE || pix_contig(vy)) && (!vz.p || pix_contig(vz)))
    return act_bwd_bias_fast(vy, vd, vz, act, dbias, as_stream(stream));
  const long long P = (long long)vd.n * vd.h * vd.w;
  if (vd.c <= 16 || (vd.c <= 32 && vd.c % 4 != 0)) {      // incl. the 19-class segmentation heads' bias gradient
    long long blocks = cdiv(P, 256 * 4); if (blocks > 148 * 8) blocks = 148 * 8; if (blocks < 1) blocks = 1;
    if (vd.c <= 4) act_bwd_bias_fewc_kernel<4><<<(unsigned)blocks, 256, 0, as_stream(stream)>>>(vy, vd, vz, act, dbias);
    else if (vd.c <= 16) act_bwd_bias_fewc_kernel<16><<<(unsigned)blocks, 256, 0, as_stream(stream)>>>(vy, vd, vz, act, dbias);
    else act_bwd_bias_fewc_kernel<32><<<(unsigned)blocks, 256, 0, as_stream(stream)>>>(vy, vd, vz, act, dbias);
    return launched();
  }
  const int groups = cdiv(vd.c, 32);
  long long want = (148LL * 8) / groups; if (want < 1) want = 1;
  long long s = cdiv(P, 64); if (s > want) s = want; if (s < 1) s = 1;
  const long long slab = (P + s - 1) / s;
  dim3 grid(groups, cdiv(P, slab)), block(32, 8);
  act_bwd_bias_kernel<<<grid, block, 0, as_stream(stream)>>>(vy, vd, vz, act, dbias, slab);
  return launched();
}

// ---------------------------------------------------------------------------------------------------
// Stem (7x7 / stride 2 on 3- or 6-channel NCHW images, resnet_encoder.py:92-93) as a tensor-core GEMM:
// an im2col pass writes cols[n, oh, ow, (r*kw+s)*Cin + c] = (x[n,c,2oh-3+r,2ow-3+s] - 0.45)/0.225 (0 outside
// the image, K padded with zeros to a multiple of 32), after which conv1 is a 1x1 convolution over a
// Kpad-channel NHWC tensor and its wgrad is the 1x1 wgrad.
// ---------------------------------------------------------------------------------------------------
namespace segsde {
// One CTA per 32 consecutive output pixels of a row.  The input patch ([C][kh][32*stride + kw] floats, normalised,
// zero outside the image) is staged in shared memory with coalesced row reads, a k -> patch-offset table replaces
// the per-element div/mod, and the 32 x Kpad block leaves through a transposing tile with contiguous writes.
__global__ void __launch_bounds__(256) stem_im2col_kernel(const float* __restrict__ x1, const float* __restrict__ x2,
                                                          int C1, int C2, int N, int H, int W, int kh, int kw, int stride,
                                                          int pad, int Ho, int Wo, int Kpad, float* __restrict__ cols) {
  extern __shared__ float sm[];
  const int Ct = C1 + C2, K = kh * kw * Ct, ld = Kpad + 1;
  const int PW = 32 * stride + kw;                 // patch row length
  float* tile = sm;                                // [32][ld]
  float* patch = tile + 32 * ld;                   // [Ct][kh][PW]
  int* lut = reinterpret_cast<int*>(patch + Ct * kh * PW);   // [Kpad]
  const int segs = (Wo + 31) / 32;
  int b = blockIdx.x;
  const int seg = b % segs; b /= segs;
  const int oh = b % Ho; const int n = b / Ho;
  const int ow0 = seg * 32;
  for (int k = threadIdx.x; k < Kpad; k += blockDim.x) {
    int v = -1;
    if (k < K) { const int c = k % Ct, tap = k / Ct, s = tap % kw, r = tap / kw; v = (c * kh + r) * PW + s; }
    lut[k] = v;
  }
  const int w0 = ow0 * stride - pad, h0 = oh * stride - pad;
  for (int i = threadIdx.x; i < Ct * kh * PW; i += blockDim.x) {
    const int j = i % PW, cr = i / PW, r = cr % kh, c = cr / kh;
    const int h = h0 + r, w = w0 + j;
    float v = 0.f;
    if (h >= 0 && h < H && w >= 0 && w < W) {
      const float* src = c < C1 ? x1 + (((long long)n * C1 + c) * H + h) * W + w
                                : x2 + (((long long)n * C2 + (c - C1)) * H + h) * W + w;
      v = (__ldg(src) - 0.45f) / 0.225f;
    }
    patch[i] = v;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 32 * Kpad; i += blockDim.x) {
    const int px = i & 31, k = i >> 5;
    const int o = lut[k];
    tile[px * ld + k] = o >= 0 ? patch[o + px * stride] : 0.f;
  }
  __syncthreads();
  const int npx = min(32, Wo - ow0);
  float* dst = cols + (((long long)n * Ho + oh) * Wo + ow0) * Kpad;
  // one warp per pixel row: contiguous 128-byte segments, no integer division
  for (int px = threadIdx.x >> 5; px < npx; px += blockDim.x >> 5)
    for (int k = threadIdx.x & 31; k < Kpad; k += 32) dst[(long long)px * Kpad + k] = tile[px * ld + k];
}
// dst[r][0..cols_dst) = src[r][0..cols_src) zero-padded / truncated (row-major)
__global__ void copy_rows_kernel(const float* __restrict__ src, int ld_src, float* __restrict__ dst, int ld_dst,
                                 int rows, int ncopy) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * ld_dst) return;
  const int r = i / ld_dst, c = i % ld_dst;
  dst[i] = c < ncopy ? src[(long long)r * ld_src + c] : 0.f;
}
}  // namespace segsde

extern "C" int segsde_stem_im2col(const float* x1, const float* x2, int c1, int c2, int n, int h, int w, int kh, int kw,
                                  int stride, int pad, int kpad, float* cols, void* stream) {
  if (!x1 || !cols || c1 < 1 || (c2 > 0 && !x2) || kpad < kh * kw * (c1 + c2)) return SEGSDE_E_ARG;
  const int Ho = (h + 2 * pad - kh) / stride + 1, Wo = (w + 2 * pad - kw) / stride + 1;
  const long long blocks = (long long)n * Ho * cdiv(Wo, 32);
  const size_t smem = sizeof(float) * (32 * (kpad + 1) + (size_t)(c1 + c2) * kh * (32 * stride + kw) + kpad);
  static bool attr = false;
  if (!attr) { cudaFuncSetAttribute(stem_im2col_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024); attr = true; }
  if (smem > 96 * 1024) return SEGSDE_E_UNSUPPORTED;
  stem_im2col_kernel<<<(unsigned)blocks, 256, smem, as_stream(stream)>>>(x1, x2, c1, c2, n, h, w, kh, kw, stride, pad, Ho, Wo,
                                                                        kpad, cols);
  return launched();
}
extern "C" int segsde_copy_rows(const float* src, int ld_src, float* dst, int ld_dst, int rows, int ncopy, void* stream) {
  if (!src || !dst || rows < 1 || ld_dst < 1 || ncopy > ld_src || ncopy > ld_dst) return SEGSDE_E_ARG;
  copy_rows_kernel<<<cdiv((long long)rows * ld_dst, 256), 256, 0, as_stream(stream)>>>(src, ld_src, dst, ld_dst, rows, ncopy);
  return launched();
}

// ------------------------------------------------------------------------------------------------
// Stem without im2col: zero-haloed, channel-padded NHWC copy of the normalised frames (read by the tensor-core
// kernels through an overlapping row-band view) and the matching weight repack.
// ------------------------------------------------------------------------------------------------
namespace segsde {
// one thread per padded pixel: reads c1 + c2 planar values (coalesced along x), writes P contiguous floats
template <int P>
__global__ void __launch_bounds__(256) stem_pack_kernel(const float* __restrict__ x1, const float* __restrict__ x2, int c1,
                                                        int c2, int n, int h, int w, int halo, int wp,
                                                        float* __restrict__ xp) {
  const int hp = h + 2 * halo;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)n * hp * wp) return;
  const int px = (int)(i % wp); const long long q = i / wp;
  const int py = (int)(q % hp); const int b = (int)(q / hp);
  const int x = px - halo, y = py - halo;
  float v[P];
#pragma unroll
  for (int c = 0; c < P; ++c) v[c] = 0.f;
  if (x >= 0 && x < w && y >= 0 && y < h) {
    const long long plane = (long long)h * w, o = (long long)y * w + x;
#pragma unroll
    for (int c = 0; c < P; ++c) {
      if (c < c1) v[c] = (__ldg(x1 + ((long long)b * c1 + c) * plane + o) - 0.45f) / 0.225f;
      else if (c < c1 + c2) v[c] = (__ldg(x2 + ((long long)b * c2 + (c - c1)) * plane + o) - 0.45f) / 0.225f;
    }
  }
  float4* dst = reinterpret_cast<float4*>(xp + i * P);
#pragma unroll
  for (int j = 0; j < P / 4; ++j) dst[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
}
__global__ void stem_pack_w_kernel(float* __restrict__ w, float* __restrict__ wp, int cout, int kh, int kw, int cin, int P,
                                   int dir) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int row = 8 * P;
  if (i >= cout * kh * row) return;
  const int j = i % row, ky = (i / row) % kh, co = i / (row * kh);
  const int kx = j / P, c = j % P;
  const bool real = kx < kw && c < cin;
  const long long wi = (((long long)co * kh + ky) * kw + kx) * cin + c;
  if (dir == 0) wp[i] = real ? w[wi] : 0.f;
  else if (real) w[wi] = wp[i];
}
}  // namespace segsde

extern "C" int segsde_stem_pack(const float* x1, const float* x2, int c1, int c2, int n, int h, int w, int halo, int wp,
                                int P, float* xp, void* stream) {
  if (!x1 || !xp || c1 < 1 || c2 < 0 || (c2 > 0 && !x2) || n < 1 || h < 1 || w < 1 || halo < 0) return SEGSDE_E_ARG;
  if ((P != 4 && P != 8) || c1 + c2 > P || wp < w + 2 * halo) return SEGSDE_E_ARG;
  if (reinterpret_cast<uintptr_t>(xp) & 15) return SEGSDE_E_ARG;
  const long long total = (long long)n * (h + 2 * halo) * wp;
  if (P == 4) stem_pack_kernel<4><<<cdiv(total, 256), 256, 0, as_stream(stream)>>>(x1, x2, c1, c2, n, h, w, halo, wp, xp);
  else stem_pack_kernel<8><<<cdiv(total, 256), 256, 0, as_stream(stream)>>>(x1, x2, c1, c2, n, h, w, halo, wp, xp);
  return launched();
}
extern "C" int segsde_stem_pack_w(float* w, float* wp, int cout, int kh, int kw, int cin, int P, int dir, void* stream) {
  if (!w || !wp || cout < 1 || kh < 1 || kw < 1 || kw > 8 || cin < 1 || cin > P || (dir != 0 && dir != 1)) return SEGSDE_E_ARG;
  stem_pack_w_kernel<<<cdiv((long long)cout * kh * 8 * P, 256), 256, 0, as_stream(stream)>>>(w, wp, cout, kh, kw, cin, P, dir);
  return launched();
}
