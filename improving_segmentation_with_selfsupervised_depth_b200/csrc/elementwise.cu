// HBM-bound support kernels of the network path on strided NHWC views: BatchNorm (train/eval,
// forward/backward), max-pool, spatial reduce/broadcast, bilinear resize, dropout, layout changes,
// residual add and the self-attention gate.  All are single-pass, channel-coalesced, float4 when
// every view allows it.
#include "common.cuh"

namespace segsde {

// decompose a flat (pixel, channel-vector) index of view V
// grid = (chunks of one image, n): 32-bit index math only (a flat 64-bit index costs three 64-bit div/mod chains per
// element vector, which made the copy-class kernels issue-bound; profiles/r2_hot_kernels.md)
#define EW_DECOMP(V, VEC)                                                     \
  const unsigned cq_ = (unsigned)((V).c / (VEC));                             \
  const unsigned idx_ = blockIdx.x * 256u + threadIdx.x;                      \
  if (idx_ >= (unsigned)(V).h * (unsigned)(V).w * cq_) return;                \
  const unsigned q_ = idx_ / cq_;                                             \
  const int c_ = (int)(idx_ - q_ * cq_) * (VEC);                              \
  const int h_ = (int)(q_ / (unsigned)(V).w);                                 \
  const int w_ = (int)(q_ - (unsigned)h_ * (unsigned)(V).w);                  \
  const int n_ = (int)blockIdx.y;

template <int VEC> struct Vec;
template <> struct Vec<1> {
  float v[1];
  __device__ __forceinline__ void load(const float* p) { v[0] = *p; }
  __device__ __forceinline__ void store(float* p) const { *p = v[0]; }
};
template <> struct Vec<4> {
  float v[4];
  __device__ __forceinline__ void load(const float* p) {
    const float4 t = *reinterpret_cast<const float4*>(p); v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
  }
  __device__ __forceinline__ void store(float* p) const {
    *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
  }
};

static inline dim3 ew_blocks(const View& v, int vec) {
  return dim3((unsigned)cdiv((long long)v.h * v.w * (v.c / vec), 256), (unsigned)v.n);
}

// ---------------------------------------------------------------------------------------------
// per-channel reductions over N*H*W  (BN statistics, BN backward sums)
// block = (32 channels, 8 pixel lanes); grid = (channel groups, pixel slabs)
// MODE 0: sum x, sum x^2            MODE 1: sum dz, sum dz*xhat  (dz = dy * [y>0] when relu)
// ---------------------------------------------------------------------------------------------
template <int MODE>
__global__ void __launch_bounds__(256) chan_reduce_kernel(View x, View y, View dy, const float* __restrict__ mean,
                                                          const float* __restrict__ invstd, int relu,
                                                          double* __restrict__ out, long long slab,
                                                          const float* __restrict__ gamma = nullptr,
                                                          const float* __restrict__ beta = nullptr) {
  const int c = blockIdx.x * 32 + threadIdx.x;
  const long long P = (long long)x.n * x.h * x.w;
  const long long pbeg = (long long)blockIdx.y * slab, pend = min(P, pbeg + slab);
  float mu = 0.f, is = 0.f, sc = 0.f, bt = 0.f;
  if (MODE == 1 && c < x.c) { mu = mean[c]; is = invstd[c]; sc = is * (gamma ? gamma[c] : 1.f); bt = beta ? beta[c] : 0.f; }
  // fp64 accumulation: var = E[x^2] - mean^2 cancels catastrophically in fp32 when var << mean^2
  // (e.g. ASPPPooling's BatchNorm over B x 256 x 1 x 1 with similar samples); float*float is exact in double.
  double d0 = 0.0, d1 = 0.0;
  if (c < x.c) {
    for (long long p = pbeg + threadIdx.y; p < pend; p += 8) {
      const int w = (int)(p % x.w); const long long q = p / x.w;
      const int h = (int)(q % x.h), n = (int)(q / x.h);
      const float xv = x.p[x.off(n, h, w) + c];
      if (MODE == 0) { d0 += (double)xv; d1 += (double)xv * (double)xv; }
      else {
        float g = dy.p[dy.off(n, h, w) + c];
        if (relu) {      // mask from the saved output, or recomputed from x (no residual; the forward's exact expression)
          const float o = y.p ? y.p[y.off(n, h, w) + c] : (xv - mu) * sc + bt;
          if (!(o > 0.f)) g = 0.f;
        }
        d0 += (double)g; d1 += (double)g * (double)((xv - mu) * is);
      }
    }
  }
  __shared__ double sh[2][8][32];
  sh[0][threadIdx.y][threadIdx.x] = d0; sh[1][threadIdx.y][threadIdx.x] = d1;
  __syncthreads();
  if (threadIdx.y == 0 && c < x.c) {
    double a = 0.0, b = 0.0;
    for (int i = 0; i < 8; ++i) { a += sh[0][i][threadIdx.x]; b += sh[1][i][threadIdx.x]; }
    atomicAdd(out + c, a); atomicAdd(out + x.c + c, b);
  }
}

__global__ void bn_finalize_kernel(const double* __restrict__ sums, int C, double count, float eps, float momentum,
                                   float* __restrict__ mean, float* __restrict__ invstd,
                                   float* __restrict__ rmean, float* __restrict__ rvar) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  // sums hold sum(x - s), sum((x - s)^2) and the shift s (0 on the generic path)
  const double ms = sums[c] / count;
  double var = sums[C + c] / count - ms * ms;
  if (var < 0.0) var = 0.0;
  const double m = ms + sums[2 * C + c];
  mean[c] = (float)m;
  invstd[c] = (float)(1.0 / sqrt(var + (double)eps));
  if (rmean) {
    const double unbiased = count > 1.0 ? var * count / (count - 1.0) : var;
    rmean[c] = (1.f - momentum) * rmean[c] + momentum * (float)m;
    rvar[c] = (1.f - momentum) * rvar[c] + momentum * (float)unbiased;
  }
}

__global__ void bn_eval_prepare_kernel(const float* __restrict__ rmean, const float* __restrict__ rvar, int C,
                                       float eps, float* __restrict__ mean, float* __restrict__ invstd) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  mean[c] = rmean[c];
  invstd[c] = 1.f / sqrtf(rvar[c] + eps);
}

template <int VEC>
__global__ void bn_apply_kernel(View x, const float* __restrict__ mean, const float* __restrict__ invstd,
                                const float* __restrict__ gamma, const float* __restrict__ beta, View res, View y,
                                int act) {
  EW_DECOMP(x, VEC)
  Vec<VEC> xv, rv, o;
  xv.load(x.p + x.off(n_, h_, w_) + c_);
  if (res.p) rv.load(res.p + res.off(n_, h_, w_) + c_);
#pragma unroll
  for (int i = 0; i < VEC; ++i) {
    const int c = c_ + i;
    const float sc = invstd[c] * (gamma ? gamma[c] : 1.f);
    float v = (xv.v[i] - mean[c]) * sc + (beta ? beta[c] : 0.f);
    if (res.p) v += rv.v[i];
    o.v[i] = act == SEGSDE_ACT_RELU ? fmaxf(v, 0.f) : v;
  }
  o.store(y.p + y.off(n_, h_, w_) + c_);
}

template <int VEC>
__global__ void bn_bwd_apply_kernel(View x, View y, View dy, const float* __restrict__ mean,
                                    const float* __restrict__ invstd, const float* __restrict__ gamma, int relu,
                                    int training, const double* __restrict__ red, float inv_count, View dx,
                                    View dres, const float* __restrict__ beta) {
  EW_DECOMP(x, VEC)
  Vec<VEC> xv, yv, gv, o;
  xv.load(x.p + x.off(n_, h_, w_) + c_);
  gv.load(dy.p + dy.off(n_, h_, w_) + c_);
  if (relu && y.p) yv.load(y.p + y.off(n_, h_, w_) + c_);
#pragma unroll
  for (int i = 0; i < VEC; ++i) {
    const int c = c_ + i;
    float g = gv.v[i];
    const float sc = invstd[c] * (gamma ? gamma[c] : 1.f);
    if (relu) {
      const float fwd = y.p ? yv.v[i] : (xv.v[i] - mean[c]) * sc + (beta ? beta[c] : 0.f);
      if (!(fwd > 0.f)) g = 0.f;
    }
    gv.v[i] = g;
    if (training) {
      const float xh = (xv.v[i] - mean[c]) * invstd[c];
      const float m0 = (float)red[c] * inv_count, m1 = (float)red[x.c + c] * inv_count;
      o.v[i] = sc * (g - m0 - xh * m1);
    } else {
      o.v[i] = sc * g;
    }
  }
  if (dx.p) o.store(dx.p + dx.off(n_, h_, w_) + c_);
  if (dres.p) gv.store(dres.p + dres.off(n_, h_, w_) + c_);
}

__global__ void bn_param_grad_kernel(const double* __restrict__ red, int C, float* __restrict__ dgamma,
                                     float* __restrict__ dbeta) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  if (dbeta) dbeta[c] += (float)red[c];
  if (dgamma) dgamma[c] += (float)red[C + c];
}

// ---------------------------------------------------------------------------------------------
// MaxPool 3x3 / stride 2 / pad 1
// ---------------------------------------------------------------------------------------------
template <int VEC>
__global__ void maxpool_fwd_kernel(View x, View y, uint8_t* __restrict__ idx) {
  EW_DECOMP(y, VEC)
  float best[VEC]; int bi[VEC];
#pragma unroll
  for (int i = 0; i < VEC; ++i) { best[i] = -3.4e38f; bi[i] = 4; }
#pragma unroll
  for (int r = 0; r < 3; ++r) {
    const int h = h_ * 2 - 1 + r;
    if (h < 0 || h >= x.h) continue;
#pragma unroll
    for (int s = 0; s < 3; ++s) {
      const int w = w_ * 2 - 1 + s;
      if (w < 0 || w >= x.w) continue;
      Vec<VEC> v; v.load(x.p + x.off(n_, h, w) + c_);
#pragma unroll
      for (int i = 0; i < VEC; ++i)
        if (v.v[i] > best[i]) { best[i] = v.v[i]; bi[i] = r * 3 + s; }
    }
  }
  Vec<VEC> o;
#pragma unroll
  for (int i = 0; i < VEC; ++i) o.v[i] = best[i];
  o.store(y.p + y.off(n_, h_, w_) + c_);
  uint8_t* ip = idx + (((long long)n_ * y.h + h_) * y.w + w_) * y.c + c_;
  if (VEC == 4) {
    *reinterpret_cast<uchar4*>(ip) = make_uchar4((uint8_t)bi[0], (uint8_t)bi[VEC > 1 ? 1 : 0], (uint8_t)bi[VEC > 2 ? 2 : 0],
                                                 (uint8_t)bi[VEC > 3 ? 3 : 0]);
  } else {
#pragma unroll
    for (int i = 0; i < VEC; ++i) ip[i] = (uint8_t)bi[i];
  }
}

// gather form: every input element collects from the <=4 windows that contain it (no atomics)
template <int VEC>
__global__ void maxpool_bwd_kernel(View dy, const uint8_t* __restrict__ idx, View dx) {
  EW_DECOMP(dx, VEC)
  float acc[VEC];
#pragma unroll
  for (int i = 0; i < VEC; ++i) acc[i] = 0.f;
  const int oh0 = max(0, (h_ - 1 + 1) / 2), oh1 = min(dy.h - 1, (h_ + 1) / 2);
  const int ow0 = max(0, (w_ - 1 + 1) / 2), ow1 = min(dy.w - 1, (w_ + 1) / 2);
  for (int oh = oh0; oh <= oh1; ++oh)
    for (int ow = ow0; ow <= ow1; ++ow) {
      const int tap = (h_ - (2 * oh - 1)) * 3 + (w_ - (2 * ow - 1));
      const uint8_t* ip = idx + (((long long)n_ * dy.h + oh) * dy.w + ow) * dy.c + c_;
      Vec<VEC> g; g.load(dy.p + dy.off(n_, oh, ow) + c_);
      if (VEC == 4) {
        const uchar4 iv = *reinterpret_cast<const uchar4*>(ip);
        const int t4[4] = {iv.x, iv.y, iv.z, iv.w};
#pragma unroll
        for (int i = 0; i < VEC; ++i)
          if (t4[i] == tap) acc[i] += g.v[i];
      } else {
#pragma unroll
        for (int i = 0; i < VEC; ++i)
          if (ip[i] == tap) acc[i] += g.v[i];
      }
    }
  Vec<VEC> o;
#pragma unroll
  for (int i = 0; i < VEC; ++i) o.v[i] = acc[i];
  o.store(dx.p + dx.off(n_, h_, w_) + c_);
}

// ---------------------------------------------------------------------------------------------
// spatial reduce  y[n,c] = alpha * sum_hw x[n,h,w,c]   /   broadcast  y[n,h,w,c] = alpha * x[n,c]
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) spatial_reduce_kernel(View x, View y, float alpha) {
  const int c = blockIdx.x * 32 + threadIdx.x, n = blockIdx.y;
  const int HW = x.h * x.w;
  double a = 0.0;
  if (c < x.c)
    for (int p = threadIdx.y; p < HW; p += 8) a += (double)x.p[x.off(n, p / x.w, p % x.w) + c];
  __shared__ double sh[8][32];
  sh[threadIdx.y][threadIdx.x] = a;
  __syncthreads();
  if (threadIdx.y == 0 && c < x.c) {
    double t = 0.0;
    for (int i = 0; i < 8; ++i) t += sh[i][threadIdx.x];
    y.p[y.off(n, 0, 0) + c] = alpha * (float)t;
  }
}
template <int VEC>
__global__ void spatial_broadcast_kernel(View x, View y, float alpha) {
  EW_DECOMP(y, VEC)
  Vec<VEC> v; v.load(x.p + x.off(n_, 0, 0) + c_);
#pragma unroll
  for (int i = 0; i < VEC; ++i) v.v[i] *= alpha;
  v.store(y.p + y.off(n_, h_, w_) + c_);
}

// ---------------------------------------------------------------------------------------------
// copy / add / gate / layout
// ---------------------------------------------------------------------------------------------
template <int VEC>
__global__ void copy_kernel(View x, View y) {
  EW_DECOMP(x, VEC)
  Vec<VEC> v; v.load(x.p + x.off(n_, h_, w_) + c_);
  v.store(y.p + y.off(n_, h_, w_) + c_);
}
template <int VEC>
__global__ void add_kernel(View a, View b, View y) {
  EW_DECOMP(a, VEC)
  Vec<VEC> u, v; u.load(a.p + a.off(n_, h_, w_) + c_); v.load(b.p + b.off(n_, h_, w_) + c_);
#pragma unroll
  for (int i = 0; i < VEC; ++i) u.v[i] += v.v[i];
  u.store(y.p + y.off(n_, h_, w_) + c_);
}
template <int VEC>
__global__ void gate_fwd_kernel(View f, View a, View y) {
  EW_DECOMP(f, VEC)
  Vec<VEC> u, v; u.load(f.p + f.off(n_, h_, w_) + c_); v.load(a.p + a.off(n_, h_, w_) + c_);
#pragma unroll
  for (int i = 0; i < VEC; ++i) u.v[i] *= 1.f / (1.f + expf(-v.v[i]));
  u.store(y.p + y.off(n_, h_, w_) + c_);
}
template <int VEC>
__global__ void gate_bwd_kernel(View f, View a, View dy, View df, View da) {
  EW_DECOMP(f, VEC)
  Vec<VEC> u, v, g, o1, o2;
  u.load(f.p + f.off(n_, h_, w_) + c_); v.load(a.p + a.off(n_, h_, w_) + c_);
  g.load(dy.p + dy.off(n_, h_, w_) + c_);
#pragma unroll
  for (int i = 0; i < VEC; ++i) {
    const float s = 1.f / (1.f + expf(-v.v[i]));
    o1.v[i] = g.v[i] * s;
    o2.v[i] = g.v[i] * u.v[i] * s * (1.f - s);
  }
  if (df.p) o1.store(df.p + df.off(n_, h_, w_) + c_);
  if (da.p) o2.store(da.p + da.off(n_, h_, w_) + c_);
}
__global__ void nchw_to_nhwc_kernel(const float* __restrict__ x, long long sn, long long sc, long long sh,
                                    long long sw, View y) {
  EW_DECOMP(y, 1)
  y.p[y.off(n_, h_, w_) + c_] = x[n_ * sn + c_ * sc + h_ * sh + w_ * sw];
}
__global__ void nhwc_to_nchw_kernel(View x, float* __restrict__ y, long long sn, long long sc, long long sh,
                                    long long sw) {
  // iterate in NCHW order so the writes coalesce
  const long long total = (long long)x.n * x.c * x.h * x.w;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int w = (int)(idx % x.w); long long q = idx / x.w;
  const int h = (int)(q % x.h); q /= x.h;
  const int c = (int)(q % x.c); const int n = (int)(q / x.c);
  y[n * sn + c * sc + h * sh + w * sw] = x.p[x.off(n, h, w) + c];
}

// ---------------------------------------------------------------------------------------------
// bilinear resize (F.interpolate semantics)
// ---------------------------------------------------------------------------------------------
struct Lin { int i0, i1; float l0, l1; };
__device__ __forceinline__ Lin lin_src(int dst, int in, int out, int align) {
  Lin r;
  float src;
  if (align) {
    const float sc = out > 1 ? (float)(in - 1) / (float)(out - 1) : 0.f;
    src = sc * (float)dst;
  } else {
    const float sc = (float)in / (float)out;
    src = sc * ((float)dst + 0.5f) - 0.5f;
    if (src < 0.f) src = 0.f;
  }
  r.i0 = (int)src;
  if (r.i0 > in - 1) r.i0 = in - 1;
  r.i1 = r.i0 + (r.i0 < in - 1 ? 1 : 0);
  r.l1 = src - (float)r.i0; r.l0 = 1.f - r.l1;
  return r;
}
template <int VEC>
__global__ void bilinear_fwd_kernel(View x, View y, int align) {
  EW_DECOMP(y, VEC)
  const Lin ly = lin_src(h_, x.h, y.h, align), lx = lin_src(w_, x.w, y.w, align);
  Vec<VEC> a, b, c, d, o;
  a.load(x.p + x.off(n_, ly.i0, lx.i0) + c_); b.load(x.p + x.off(n_, ly.i0, lx.i1) + c_);
  c.load(x.p + x.off(n_, ly.i1, lx.i0) + c_); d.load(x.p + x.off(n_, ly.i1, lx.i1) + c_);
#pragma unroll
  for (int i = 0; i < VEC; ++i)
    o.v[i] = ly.l0 * (lx.l0 * a.v[i] + lx.l1 * b.v[i]) + ly.l1 * (lx.l0 * c.v[i] + lx.l1 * d.v[i]);
  o.store(y.p + y.off(n_, h_, w_) + c_);
}
template <int VEC>
__global__ void bilinear_bwd_kernel(View dy, View dx, int align) {
  EW_DECOMP(dy, VEC)
  const Lin ly = lin_src(h_, dx.h, dy.h, align), lx = lin_src(w_, dx.w, dy.w, align);
  Vec<VEC> g; g.load(dy.p + dy.off(n_, h_, w_) + c_);
  float* p00 = dx.p + dx.off(n_, ly.i0, lx.i0) + c_; float* p01 = dx.p + dx.off(n_, ly.i0, lx.i1) + c_;
  float* p10 = dx.p + dx.off(n_, ly.i1, lx.i0) + c_; float* p11 = dx.p + dx.off(n_, ly.i1, lx.i1) + c_;
#pragma unroll
  for (int i = 0; i < VEC; ++i) {
    atomicAdd(p00 + i, ly.l0 * lx.l0 * g.v[i]); atomicAdd(p01 + i, ly.l0 * lx.l1 * g.v[i]);
    atomicAdd(p10 + i, ly.l1 * lx.l0 * g.v[i]); atomicAdd(p11 + i, ly.l1 * lx.l1 * g.v[i]);
  }
}

// ---------------------------------------------------------------------------------------------
// dropout
// ---------------------------------------------------------------------------------------------
__global__ void dropout_fwd_kernel(View x, View y, float p, unsigned long long seed, unsigned long long offset,
                                   const float* __restrict__ mask_in, uint8_t* __restrict__ mask_out,
                                   int channelwise) {
  EW_DECOMP(x, 1)
  const long long lin = channelwise ? ((long long)n_ * x.c + c_)
                                    : ((((long long)n_ * x.h + h_) * x.w + w_) * x.c + c_);
  float keep;
  if (mask_in) {
    // replay masks come in the reference's NCHW element order
    const long long li = channelwise ? lin : ((((long long)n_ * x.c + c_) * x.h + h_) * x.w + w_);
    keep = mask_in[li];
  } else {
    Philox ph(seed, (unsigned long long)(lin >> 2), offset);
    ph.run();
    keep = u01(ph.c[lin & 3]) > p ? 1.f : 0.f;
  }
  if (mask_out) mask_out[channelwise ? lin : ((((long long)n_ * x.h + h_) * x.w + w_) * x.c + c_)] = keep != 0.f;
  y.p[y.off(n_, h_, w_) + c_] = keep != 0.f ? x.p[x.off(n_, h_, w_) + c_] / (1.f - p) : 0.f;
}
__global__ void dropout_bwd_kernel(View dy, const uint8_t* __restrict__ mask, float p, int channelwise, View dx) {
  EW_DECOMP(dy, 1)
  const long long lin = channelwise ? ((long long)n_ * dy.c + c_)
                                    : ((((long long)n_ * dy.h + h_) * dy.w + w_) * dy.c + c_);
  dx.p[dx.off(n_, h_, w_) + c_] = mask[lin] ? dy.p[dy.off(n_, h_, w_) + c_] / (1.f - p) : 0.f;
}

}  // namespace segsde
using namespace segsde;

#define DISPATCH_VEC(cond, KERNEL, blocks_view, ...)                                              \
  do {                                                                                            \
    if (cond) KERNEL<4><<<ew_blocks(blocks_view, 4), 256, 0, as_stream(stream)>>>(__VA_ARGS__);   \
    else KERNEL<1><<<ew_blocks(blocks_view, 1), 256, 0, as_stream(stream)>>>(__VA_ARGS__);        \
  } while (0)

static long long slab_for(const View& x, int& nslabs) {
  const long long P = (long long)x.n * x.h * x.w;
  const int groups = cdiv(x.c, 32);
  long long want = (148LL * 8) / groups;
  if (want < 1) want = 1;
  long long s = cdiv(P, 64);
  if (s > want) s = want;
  if (s < 1) s = 1;
  const long long slab = (P + s - 1) / s;
  nslabs = cdiv(P, slab);
  return slab;
}

extern "C" int segsde_bn_stats(const segsde_nhwc_t* x, double* sums, void* stream) {
  if (!x || !x->ptr || !sums) return SEGSDE_E_ARG;
  View v = mk(x), none = mk(nullptr);
  if (fast_reduce_ok(v)) return bn_stats_fast(v, sums, as_stream(stream));
  int ns; const long long slab = slab_for(v, ns);
  dim3 grid(cdiv(v.c, 32), ns), block(32, 8);
  chan_reduce_kernel<0><<<grid, block, 0, as_stream(stream)>>>(v, none, none, nullptr, nullptr, 0, sums, slab);
  return launched();
}
extern "C" int segsde_bn_finalize(const double* sums, int c, int64_t count, float eps, float momentum,
                                  float* mean, float* invstd, float* running_mean, float* running_var,
                                  void* stream) {
  if (!sums || !mean || !invstd || c < 1 || count < 1) return SEGSDE_E_ARG;
  if ((running_mean == nullptr) != (running_var == nullptr)) return SEGSDE_E_ARG;
  bn_finalize_kernel<<<cdiv(c, 128), 128, 0, as_stream(stream)>>>(sums, c, (double)count, eps, momentum, mean,
                                                                 invstd, running_mean, running_var);
  return launched();
}
extern "C" int segsde_bn_eval_prepare(const float* running_mean, const float* running_var, int c, float eps,
                                      float* mean, float* invstd, void* stream) {
  if (!running_mean || !running_var || !mean || !invstd || c < 1) return SEGSDE_E_ARG;
  bn_eval_prepare_kernel<<<cdiv(c, 128), 128, 0, as_stream(stream)>>>(running_mean, running_var, c, eps, mean, invstd);
  return launched();
}
extern "C" int segsde_bn_apply(const segsde_nhwc_t* x, const float* mean, const float* invstd,
                               const float* gamma, const float* beta, const segsde_nhwc_t* residual,
                               const segsde_nhwc_t* y, int act, void* stream) {
  if (!x || !x->ptr || !y || !y->ptr || !mean || !invstd) return SEGSDE_E_ARG;
  View vx = mk(x), vy = mk(y), vr = mk(residual);
  if (!same_shape(vx, vy) || (vr.p && !same_shape(vx, vr))) return SEGSDE_E_ARG;
  if (act != SEGSDE_ACT_NONE && act != SEGSDE_ACT_RELU) return SEGSDE_E_UNSUPPORTED;
  const bool v4 = vec4_ok(vx) && vec4_ok(vy) && (!vr.p || vec4_ok(vr));
  if (pix_contig(vx) && pix_contig(vy) && (!vr.p || pix_contig(vr)))
    return bn_apply_fast(vx, vr, vy, mean, invstd, gamma, beta, act, as_stream(stream));
  DISPATCH_VEC(v4, bn_apply_kernel, vx, vx, mean, invstd, gamma, beta, vr, vy, act);
  return launched();
}
extern "C" int segsde_bn_apply_train(const segsde_nhwc_t* x, const double* sums, int64_t count, float eps, float momentum,
                                     const float* gamma, const float* beta, const segsde_nhwc_t* residual,
                                     const segsde_nhwc_t* y, int act, float* mean, float* invstd, float* running_mean,
                                     float* running_var, void* stream) {
  if (!x || !x->ptr || !y || !y->ptr || !sums || !mean || !invstd || count < 1) return SEGSDE_E_ARG;
  if ((running_mean == nullptr) != (running_var == nullptr)) return SEGSDE_E_ARG;
  View vx = mk(x), vy = mk(y), vr = mk(residual);
  if (!same_shape(vx, vy) || (vr.p && !same_shape(vx, vr))) return SEGSDE_E_ARG;
  if (act != SEGSDE_ACT_NONE && act != SEGSDE_ACT_RELU) return SEGSDE_E_UNSUPPORTED;
  if (vx.c <= 2048 && pix_contig(vx) && pix_contig(vy) && (!vr.p || pix_contig(vr)))
    return bn_apply_train_fast(vx, vr, vy, sums, count, eps, momentum, gamma, beta, act, mean, invstd, running_mean,
                               running_var, as_stream(stream));
  int rc = segsde_bn_finalize(sums, vx.c, count, eps, momentum, mean, invstd, running_mean, running_var, stream);
  if (rc) return rc;
  return segsde_bn_apply(x, mean, invstd, gamma, beta, residual, y, act, stream);
}
extern "C" int segsde_bn_bwd_reduce(const segsde_nhwc_t* x, const segsde_nhwc_t* y, const segsde_nhwc_t* dy,
                                    const float* mean, const float* invstd, const float* gamma, const float* beta,
                                    int act, double* red, void* stream) {
  if (!x || !x->ptr || !dy || !dy->ptr || !mean || !invstd || !red) return SEGSDE_E_ARG;
  const int relu = act == SEGSDE_ACT_RELU;
  View vx = mk(x), vy = mk(y), vd = mk(dy);
  if (!same_shape(vx, vd)) return SEGSDE_E_ARG;
  if (fast_reduce_ok(vx) && pix_contig(vd) && (!relu || !vy.p || pix_contig(vy)))
    return bn_bwd_reduce_fast(vx, vy, vd, mean, invstd, act, red, as_stream(stream), gamma, beta);
  int ns; const long long slab = slab_for(vx, ns);
  dim3 grid(cdiv(vx.c, 32), ns), block(32, 8);
  chan_reduce_kernel<1><<<grid, block, 0, as_stream(stream)>>>(vx, vy, vd, mean, invstd, relu, red, slab, gamma, beta);
  return launched();
}
extern "C" int segsde_bn_bwd_apply(const segsde_nhwc_t* x, const segsde_nhwc_t* y, const segsde_nhwc_t* dy,
                                   const float* mean, const float* invstd, const float* gamma, const float* beta,
                                   int act, int training, const double* red, int64_t count, const segsde_nhwc_t* dx,
                                   const segsde_nhwc_t* dres, float* dgamma, float* dbeta, void* stream) {
  if (!x || !x->ptr || !dy || !dy->ptr || !mean || !invstd || !red || count < 1) return SEGSDE_E_ARG;
  const int relu = act == SEGSDE_ACT_RELU;
  View vx = mk(x), vy = mk(y), vd = mk(dy), vdx = mk(dx), vdr = mk(dres);
  if (relu && !vy.p && vdr.p) return SEGSDE_E_ARG;      // with a residual the mask needs the saved output
  int rc = SEGSDE_OK;
  const bool fast = vx.c <= 2048 && pix_contig(vx) && pix_contig(vd) && (!relu || !vy.p || pix_contig(vy)) && (!vdx.p || pix_contig(vdx)) &&
                    (!vdr.p || pix_contig(vdr));
  if ((vdx.p || vdr.p) && fast) {
    rc = bn_bwd_apply_fast(vx, vy, vd, vdx, vdr, mean, invstd, gamma, relu, training, red, count, as_stream(stream), beta);
    if (rc) return rc;
  } else if (vdx.p || vdr.p) {
    const bool v4 = vec4_ok(vx) && vec4_ok(vd) && (!relu || !vy.p || vec4_ok(vy)) && (!vdx.p || vec4_ok(vdx)) &&
                    (!vdr.p || vec4_ok(vdr));
    DISPATCH_VEC(v4, bn_bwd_apply_kernel, vx, vx, vy, vd, mean, invstd, gamma, relu, training, red,
                 (float)(1.0 / (double)count), vdx, vdr, beta);
    rc = launched();
    if (rc) return rc;
  }
  if (dgamma || dbeta) {
    bn_param_grad_kernel<<<cdiv(vx.c, 128), 128, 0, as_stream(stream)>>>(red, vx.c, dgamma, dbeta);
    rc = launched();
  }
  return rc;
}

extern "C" int segsde_maxpool3x3s2_fwd(const segsde_nhwc_t* x, const segsde_nhwc_t* y, uint8_t* idx, void* stream) {
  if (!x || !x->ptr || !y || !y->ptr || !idx) return SEGSDE_E_ARG;
  View vx = mk(x), vy = mk(y);
  if (vy.h != (vx.h + 2 - 3) / 2 + 1 || vy.w != (vx.w + 2 - 3) / 2 + 1 || vy.c != vx.c || vy.n != vx.n) return SEGSDE_E_ARG;
  DISPATCH_VEC(vec4_ok(vx) && vec4_ok(vy), maxpool_fwd_kernel, vy, vx, vy, idx);
  return launched();
}
extern "C" int segsde_maxpool3x3s2_bwd(const segsde_nhwc_t* dy, const uint8_t* idx, const segsde_nhwc_t* dx, void* stream) {
  if (!dy || !dy->ptr || !dx || !dx->ptr || !idx) return SEGSDE_E_ARG;
  View vd = mk(dy), vx = mk(dx);
  DISPATCH_VEC(vec4_ok(vd) && vec4_ok(vx), maxpool_bwd_kernel, vx, vd, idx, vx);
  return launched();
}

extern "C" int segsde_spatial_mean_fwd(const segsde_nhwc_t* x, const segsde_nhwc_t* y, float scale, void* stream) {
  if (!x || !x->ptr || !y || !y->ptr) return SEGSDE_E_ARG;
  View vx = mk(x), vy = mk(y);
  if (vy.c != vx.c || vy.n != vx.n) return SEGSDE_E_ARG;
  dim3 grid(cdiv(vx.c, 32), vx.n), block(32, 8);
  spatial_reduce_kernel<<<grid, block, 0, as_stream(stream)>>>(vx, vy, scale / (float)(vx.h * vx.w));
  return launched();
}
extern "C" int segsde_spatial_mean_bwd(const segsde_nhwc_t* dy, const segsde_nhwc_t* dx, float scale, void* stream) {
  if (!dy || !dy->ptr || !dx || !dx->ptr) return SEGSDE_E_ARG;
  View vd = mk(dy), vx = mk(dx);
  DISPATCH_VEC(vec4_ok(vd) && vec4_ok(vx), spatial_broadcast_kernel, vx, vd, vx, scale / (float)(vx.h * vx.w));
  return launched();
}
extern "C" int segsde_broadcast_hw_fwd(const segsde_nhwc_t* x, const segsde_nhwc_t* y, void* stream) {
  if (!x || !x->ptr || !y || !y->ptr) return SEGSDE_E_ARG;
  View vx = mk(x), vy = mk(y);
  DISPATCH_VEC(vec4_ok(vx) && vec4_ok(vy), spatial_broadcast_kernel, vy, vx, vy, 1.f);
  return launched();
}
extern "C" int segsde_broadcast_hw_bwd(const segsde_nhwc_t* dy, const segsde_nhwc_t* dx, void* stream) {
  if (!dy || !dy->ptr || !dx || !dx->ptr) return SEGSDE_E_ARG;
  View vd = mk(dy), vx = mk(dx);
  dim3 grid(cdiv(vd.c, 32), vd.n), block(32, 8);
  spatial_reduce_kernel<<<grid, block, 0, as_stream(stream)>>>(vd, vx, 1.f);
  return launched();
}
extern "C" int segsde_copy_nhwc(const segsde_nhwc_t* x, const segsde_nhwc_t* y, void* stream) {
  if (!x || !x->ptr || !y || !y->ptr) return SEGSDE_E_ARG;
  View vx = mk(x), vy = mk(y);
  if (!same_shape(vx, vy)) return SEGSDE_E_ARG;
  DISPATCH_VEC(vec4_ok(vx) && vec4_ok(vy), copy_kernel, vx, vx, vy);
  return launched();
}
extern "C" int segsde_add(const segsde_nhwc_t* a, const segsde_nhwc_t* b, const segsde_nhwc_t* y, void* stream) {
  if (!a || !b || !y || !a->ptr || !b->ptr || !y->ptr) return SEGSDE_E_ARG;
  View va = mk(a), vb = mk(b), vy = mk(y);
  if (!same_shape(va, vb) || !same_shape(va, vy)) return SEGSDE_E_ARG;
  DISPATCH_VEC(vec4_ok(va) && vec4_ok(vb) && vec4_ok(vy), add_kernel, va, va, vb, vy);
  return launched();
}
extern "C" int segsde_gate_fwd(const segsde_nhwc_t* f, const segsde_nhwc_t* a, const segsde_nhwc_t* y, void* stream) {
  if (!f || !a || !y || !f->ptr || !a->ptr || !y->ptr) return SEGSDE_E_ARG;
  View vf = mk(f), va = mk(a), vy = mk(y);
  if (!same_shape(vf, va) || !same_shape(vf, vy)) return SEGSDE_E_ARG;
  DISPATCH_VEC(vec4_ok(vf) && vec4_ok(va) && vec4_ok(vy), gate_fwd_kernel, vf, vf, va, vy);
  return launched();
}
extern "C" int segsde_gate_bwd(const segsde_nhwc_t* f, const segsde_nhwc_t* a, const segsde_nhwc_t* dy,
                               const segsde_nhwc_t* df, const segsde_nhwc_t* da, void* stream) {
  if (!f || !a || !dy || !f->ptr || !a->ptr || !dy->ptr) return SEGSDE_E_ARG;
  View vf = mk(f), va = mk(a), vd = mk(dy), vdf = mk(df), vda = mk(da);
  const bool v4 = vec4_ok(vf) && vec4_ok(va) && vec4_ok(vd) && (!vdf.p || vec4_ok(vdf)) && (!vda.p || vec4_ok(vda));
  DISPATCH_VEC(v4, gate_bwd_kernel, vf, vf, va, vd, vdf, vda);
  return launched();
}
extern "C" int segsde_nchw_to_nhwc(const float* x, int64_t sn, int64_t sc, int64_t sh, int64_t sw,
                                   const segsde_nhwc_t* y, void* stream) {
  if (!x || !y || !y->ptr) return SEGSDE_E_ARG;
  View vy = mk(y);
  nchw_to_nhwc_kernel<<<ew_blocks(vy, 1), 256, 0, as_stream(stream)>>>(x, sn, sc, sh, sw, vy);
  return launched();
}
extern "C" int segsde_nhwc_to_nchw(const segsde_nhwc_t* x, float* y, int64_t sn, int64_t sc, int64_t sh,
                                   int64_t sw, void* stream) {
  if (!x || !x->ptr || !y) return SEGSDE_E_ARG;
  View vx = mk(x);
  nhwc_to_nchw_kernel<<<ew_blocks(vx, 1), 256, 0, as_stream(stream)>>>(vx, y, sn, sc, sh, sw);
  return launched();
}
extern "C" int segsde_bilinear_fwd(const segsde_nhwc_t* x, const segsde_nhwc_t* y, int align_corners, void* stream) {
  if (!x || !x->ptr || !y || !y->ptr) return SEGSDE_E_ARG;
  View vx = mk(x), vy = mk(y);
  if (vx.c != vy.c || vx.n != vy.n) return SEGSDE_E_ARG;
  DISPATCH_VEC(vec4_ok(vx) && vec4_ok(vy), bilinear_fwd_kernel, vy, vx, vy, align_corners);
  return launched();
}
extern "C" int segsde_bilinear_bwd(const segsde_nhwc_t* dy, const segsde_nhwc_t* dx, int align_corners, void* stream) {
  if (!dy || !dy->ptr || !dx || !dx->ptr) return SEGSDE_E_ARG;
  View vd = mk(dy), vx = mk(dx);
  if (vx.c != vd.c || vx.n != vd.n) return SEGSDE_E_ARG;
  DISPATCH_VEC(vec4_ok(vd), bilinear_bwd_kernel, vd, vd, vx, align_corners);
  return launched();
}
extern "C" int segsde_dropout_fwd(const segsde_nhwc_t* x, const segsde_nhwc_t* y, float p, uint64_t seed,
                                  uint64_t offset, const float* mask_in, uint8_t* mask_out, int channelwise,
                                  void* stream) {
  if (!x || !x->ptr || !y || !y->ptr || p < 0.f || p >= 1.f) return SEGSDE_E_ARG;
  View vx = mk(x), vy = mk(y);
  if (!same_shape(vx, vy)) return SEGSDE_E_ARG;
  dropout_fwd_kernel<<<ew_blocks(vx, 1), 256, 0, as_stream(stream)>>>(vx, vy, p, seed, offset, mask_in, mask_out, channelwise);
  return launched();
}
extern "C" int segsde_dropout_bwd(const segsde_nhwc_t* dy, const uint8_t* mask, float p, int channelwise,
                                  const segsde_nhwc_t* dx, void* stream) {
  if (!dy || !dy->ptr || !dx || !dx->ptr || !mask) return SEGSDE_E_ARG;
  View vd = mk(dy), vx = mk(dx);
  if (!same_shape(vd, vx)) return SEGSDE_E_ARG;
  dropout_bwd_kernel<<<ew_blocks(vd, 1), 256, 0, as_stream(stream)>>>(vd, mask, p, channelwise, vx);
  return launched();
}
