// Bandwidth-tuned BatchNorm / activation-backward kernels for pixel-contiguous NHWC tensors
// (offset(pixel, c) = pixel*ld + c): no per-element div/mod, float4 channel vectors, fp32 partial sums around a
// per-channel shift (the value at pixel 0) so that E[(x-s)^2] - E[x-s]^2 does not cancel, fp64 across threads.
// The generic strided kernels in elementwise.cu remain the path for strided views and tiny tensors.
#include "common.cuh"

namespace segsde {

bool pix_contig(const View& v) {
  return v.p && v.sh == (long long)v.w * v.sw && v.sn == (long long)v.h * v.sh && vec4_ok(v);
}

struct Rows { float* p; long long ld; };
static inline Rows rows_of(const View& v) { Rows r; r.p = v.p; r.ld = v.sw; return r; }

__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ void st4(float* p, const float4& v) { *reinterpret_cast<float4*>(p) = v; }

// Column reductions.  block (32, 8); each warp row covers `ppw` pixels x `cq_w` float4 chunks.
// MODE 0: S1 = sum (x - s), S2 = sum (x - s)^2 with s = x[pixel 0]; out[2C..3C) = s
// MODE 1: S1 = sum dz, S2 = sum dz * xhat   (dz = dy masked by ReLU: from the saved output y, or - y.p == NULL, no residual -
//         RECOMPUTED from x as (x - mean) * (invstd * gamma) + beta > 0, the forward kernels' exact expression, which saves
//         one full tensor read per pass)
// MODE 2: dz = dy * act'(y) written out, S1 = sum dz (bias gradient, fp32 atomics into dbias)
template <int MODE>
__global__ void __launch_bounds__(256) colreduce_fast_kernel(Rows x, Rows y, Rows dy, Rows dz, long long P, int C,
                                                             const float* __restrict__ mean,
                                                             const float* __restrict__ invstd, int act,
                                                             double* __restrict__ out, float* __restrict__ dbias,
                                                             long long slab, const float* __restrict__ gamma = nullptr,
                                                             const float* __restrict__ beta = nullptr) {
  const int cq = C >> 2;
  const int cq_w = cq < 32 ? cq : 32;              // chunks per warp row
  const int ppw = 32 / cq_w;                        // pixels per warp row
  const int sub = threadIdx.x / cq_w;
  const int c4 = blockIdx.y * 32 + (threadIdx.x % cq_w);
  const bool cv = c4 < cq;
  const int c = c4 * 4;
  const long long pbeg = (long long)blockIdx.x * slab, pend = min(P, pbeg + slab);
  float4 sh = make_float4(0.f, 0.f, 0.f, 0.f), mu = sh, is = sh, sc = sh, bt = sh;
  if (cv) {
    if (MODE == 0) sh = ld4(x.p + c);
    if (MODE == 1) {
      mu = ld4(mean + c); is = ld4(invstd + c);
      const float4 g4 = gamma ? ld4(gamma + c) : make_float4(1.f, 1.f, 1.f, 1.f);
      sc = make_float4(is.x * g4.x, is.y * g4.y, is.z * g4.z, is.w * g4.w);
      if (beta) bt = ld4(beta + c);
    }
  }
  double d1[4] = {0, 0, 0, 0}, d2[4] = {0, 0, 0, 0};
  float4 a1 = make_float4(0.f, 0.f, 0.f, 0.f), a2 = a1;
  int cnt = 0;
  // U pixels per trip with every load issued before the first use: the reduction passes are latency-bound otherwise
  // (2 x 16 B in flight per thread at 37 % occupancy measured 2.9 TB/s; see profiles/r2_hot_kernels.md)
  constexpr int U = MODE == 0 ? 4 : 2;
  auto consume = [&](float4 g, const float4& v, const float4& o) {
    if (MODE == 0) {
      const float e0 = v.x - sh.x, e1 = v.y - sh.y, e2 = v.z - sh.z, e3 = v.w - sh.w;
      a1.x += e0; a1.y += e1; a1.z += e2; a1.w += e3;
      a2.x = fmaf(e0, e0, a2.x); a2.y = fmaf(e1, e1, a2.y); a2.z = fmaf(e2, e2, a2.z); a2.w = fmaf(e3, e3, a2.w);
    } else if (MODE == 1) {
      if (act == SEGSDE_ACT_RELU) {
        float4 t = o;
        if (!y.p) {
          t.x = (v.x - mu.x) * sc.x + bt.x; t.y = (v.y - mu.y) * sc.y + bt.y;
          t.z = (v.z - mu.z) * sc.z + bt.z; t.w = (v.w - mu.w) * sc.w + bt.w;
        }
        g.x = t.x > 0.f ? g.x : 0.f; g.y = t.y > 0.f ? g.y : 0.f; g.z = t.z > 0.f ? g.z : 0.f; g.w = t.w > 0.f ? g.w : 0.f;
      }
      a1.x += g.x; a1.y += g.y; a1.z += g.z; a1.w += g.w;
      a2.x = fmaf(g.x, (v.x - mu.x) * is.x, a2.x); a2.y = fmaf(g.y, (v.y - mu.y) * is.y, a2.y);
      a2.z = fmaf(g.z, (v.z - mu.z) * is.z, a2.z); a2.w = fmaf(g.w, (v.w - mu.w) * is.w, a2.w);
    }
    if (++cnt == 128) {       // bounded fp32 partials
      d1[0] += a1.x; d1[1] += a1.y; d1[2] += a1.z; d1[3] += a1.w;
      d2[0] += a2.x; d2[1] += a2.y; d2[2] += a2.z; d2[3] += a2.w;
      a1 = make_float4(0.f, 0.f, 0.f, 0.f); a2 = a1; cnt = 0;
    }
  };
  if (cv) {
    const long long step = 8 * ppw;
    long long p = pbeg + threadIdx.y * ppw + sub;
    if (MODE != 2) {
      const bool need_y = MODE == 1 && act == SEGSDE_ACT_RELU && y.p;
      for (; p + (U - 1) * step < pend; p += U * step) {
        float4 g[U], v[U], o[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const long long q = p + u * step;
          v[u] = ld4(x.p + q * x.ld + c);
          if (MODE == 1) g[u] = ld4(dy.p + q * dy.ld + c); else g[u] = v[u];
          if (need_y) o[u] = ld4(y.p + q * y.ld + c); else o[u] = v[u];
        }
#pragma unroll
        for (int u = 0; u < U; ++u) consume(g[u], v[u], o[u]);
      }
      for (; p < pend; p += step) {
        const float4 v = ld4(x.p + p * x.ld + c);
        const float4 g = MODE == 1 ? ld4(dy.p + p * dy.ld + c) : v;
        const float4 o = need_y ? ld4(y.p + p * y.ld + c) : v;
        consume(g, v, o);
      }
    } else {
      for (; p < pend; p += step) {
        float4 g = ld4(dy.p + p * dy.ld + c);
        if (act != SEGSDE_ACT_NONE) {
          const float4 o = ld4(y.p + p * y.ld + c);
          g.x *= act_grad_from_out(o.x, act); g.y *= act_grad_from_out(o.y, act);
          g.z *= act_grad_from_out(o.z, act); g.w *= act_grad_from_out(o.w, act);
        }
        if (dz.p) st4(dz.p + p * dz.ld + c, g);
        a1.x += g.x; a1.y += g.y; a1.z += g.z; a1.w += g.w;
        if (++cnt == 128) {
          d1[0] += a1.x; d1[1] += a1.y; d1[2] += a1.z; d1[3] += a1.w;
          a1 = make_float4(0.f, 0.f, 0.f, 0.f); cnt = 0;
        }
      }
    }
    d1[0] += a1.x; d1[1] += a1.y; d1[2] += a1.z; d1[3] += a1.w;
    d2[0] += a2.x; d2[1] += a2.y; d2[2] += a2.z; d2[3] += a2.w;
  }
  __shared__ double red[8][32][8];
#pragma unroll
  for (int i = 0; i < 4; ++i) { red[threadIdx.y][threadIdx.x][i] = d1[i]; red[threadIdx.y][threadIdx.x][4 + i] = d2[i]; }
  __syncthreads();
  if (threadIdx.y == 0 && sub == 0 && cv) {
    double t[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int yy = 0; yy < 8; ++yy)
      for (int s = 0; s < ppw; ++s)
#pragma unroll
        for (int i = 0; i < 8; ++i) t[i] += red[yy][s * cq_w + (threadIdx.x % cq_w)][i];
    if (MODE == 2) {
      if (dbias)
#pragma unroll
        for (int i = 0; i < 4; ++i) atomicAdd(dbias + c + i, (float)t[i]);
    } else {
#pragma unroll
      for (int i = 0; i < 4; ++i) { atomicAdd(out + c + i, t[i]); atomicAdd(out + C + c + i, t[4 + i]); }
      if (MODE == 0 && blockIdx.x == 0) {
        out[2 * C + c + 0] = sh.x; out[2 * C + c + 1] = sh.y; out[2 * C + c + 2] = sh.z; out[2 * C + c + 3] = sh.w;
      }
    }
  }
}

// y = act((x - mean) * invstd * gamma + beta [+ res])
__global__ void __launch_bounds__(256) bn_apply_fast_kernel(Rows x, Rows res, Rows y, long long total4, int cq, int cq_shift,
                                                            const float* __restrict__ mean, const float* __restrict__ invstd,
                                                            const float* __restrict__ gamma, const float* __restrict__ beta,
                                                            int act) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total4; i += (long long)gridDim.x * blockDim.x) {
    long long p; int c4;
    if (cq_shift >= 0) { p = i >> cq_shift; c4 = (int)(i & (cq - 1)); } else { p = i / cq; c4 = (int)(i - p * cq); }
    const int c = c4 * 4;
    const float4 v = ld4(x.p + p * x.ld + c);
    const float4 m = ld4(mean + c), is = ld4(invstd + c);
    const float4 g = gamma ? ld4(gamma + c) : make_float4(1.f, 1.f, 1.f, 1.f);
    const float4 b = beta ? ld4(beta + c) : make_float4(0.f, 0.f, 0.f, 0.f);
    float4 o;
    o.x = (v.x - m.x) * (is.x * g.x) + b.x; o.y = (v.y - m.y) * (is.y * g.y) + b.y;
    o.z = (v.z - m.z) * (is.z * g.z) + b.z; o.w = (v.w - m.w) * (is.w * g.w) + b.w;
    if (res.p) { const float4 r = ld4(res.p + p * res.ld + c); o.x += r.x; o.y += r.y; o.z += r.z; o.w += r.w; }
    if (act == SEGSDE_ACT_RELU) { o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f); }
    st4(y.p + p * y.ld + c, o);
  }
}

// Train-mode forward: bn_finalize folded into the apply pass.  Every CTA turns the fp64 sums into per-channel
// mean, scale = invstd * gamma and beta in shared memory (C <= 2048), CTA 0 also publishes mean / invstd for the
// backward pass and updates the running statistics.
__global__ void __launch_bounds__(256) bn_apply_train_fast_kernel(Rows x, Rows res, Rows y, long long total4, int cq,
                                                                  int cq_shift, const double* __restrict__ sums, int C,
                                                                  double count, float eps, float momentum,
                                                                  const float* __restrict__ gamma,
                                                                  const float* __restrict__ beta, int act,
                                                                  float* __restrict__ mean_o, float* __restrict__ invstd_o,
                                                                  float* __restrict__ rmean, float* __restrict__ rvar) {
  extern __shared__ __align__(16) float s_ss[];       // [C] mean, [C] scale, [C] beta
  float* s_mean = s_ss; float* s_scale = s_ss + C; float* s_beta = s_ss + 2 * C;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    // sums hold sum(x - s), sum((x - s)^2) and the shift s (0 when they come from a convolution epilogue)
    const double ms = sums[c] / count;
    double var = sums[C + c] / count - ms * ms;
    if (var < 0.0) var = 0.0;
    const double m = ms + sums[2 * C + c];
    const float mf = (float)m, is = (float)(1.0 / sqrt(var + (double)eps));
    s_mean[c] = mf;
    s_scale[c] = is * (gamma ? gamma[c] : 1.f);
    s_beta[c] = beta ? beta[c] : 0.f;
    if (blockIdx.x == 0) {
      mean_o[c] = mf; invstd_o[c] = is;
      if (rmean) {
        const double unbiased = count > 1.0 ? var * count / (count - 1.0) : var;
        rmean[c] = (1.f - momentum) * rmean[c] + momentum * mf;
        rvar[c] = (1.f - momentum) * rvar[c] + momentum * (float)unbiased;
      }
    }
  }
  __syncthreads();
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total4; i += (long long)gridDim.x * blockDim.x) {
    long long p; int c4;
    if (cq_shift >= 0) { p = i >> cq_shift; c4 = (int)(i & (cq - 1)); } else { p = i / cq; c4 = (int)(i - p * cq); }
    const int c = c4 * 4;
    const float4 v = ld4(x.p + p * x.ld + c);
    const float4 m = ld4(s_mean + c), sc = ld4(s_scale + c), b = ld4(s_beta + c);
    float4 o;
    o.x = (v.x - m.x) * sc.x + b.x; o.y = (v.y - m.y) * sc.y + b.y;
    o.z = (v.z - m.z) * sc.z + b.z; o.w = (v.w - m.w) * sc.w + b.w;
    if (res.p) { const float4 r = ld4(res.p + p * res.ld + c); o.x += r.x; o.y += r.y; o.z += r.z; o.w += r.w; }
    if (act == SEGSDE_ACT_RELU) { o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f); }
    st4(y.p + p * y.ld + c, o);
  }
}

// dx = scale * (dz - mean(dz) - xhat * mean(dz * xhat)); the per-channel terms are turned into five fp32 coefficients
// in shared memory once per CTA (the fp64 sums were converted per ELEMENT before: 8 LDG.64 + 8 F2F per float4 held the
// pass at 2.8 TB/s).
__global__ void __launch_bounds__(256) bn_bwd_apply_fast_kernel(Rows x, Rows y, Rows dy, Rows dx, Rows dres, long long total4,
                                                                int cq, int cq_shift, int C,
                                                                const float* __restrict__ mean,
                                                                const float* __restrict__ invstd,
                                                                const float* __restrict__ gamma, int relu, int training,
                                                                const double* __restrict__ red, float inv_count,
                                                                const float* __restrict__ beta) {
  extern __shared__ __align__(16) float s_co[];      // [C] mean, invstd, scale, m0, m1, beta
  float* s_mean = s_co; float* s_is = s_co + C; float* s_sc = s_co + 2 * C; float* s_m0 = s_co + 3 * C;
  float* s_m1 = s_co + 4 * C; float* s_bt = s_co + 5 * C;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const float is = invstd[c];
    s_mean[c] = mean[c]; s_is[c] = is; s_sc[c] = is * (gamma ? gamma[c] : 1.f);
    s_m0[c] = training ? (float)red[c] * inv_count : 0.f;
    s_m1[c] = training ? (float)red[C + c] * inv_count : 0.f;
    s_bt[c] = beta ? beta[c] : 0.f;
  }
  __syncthreads();
  const long long stride = (long long)gridDim.x * blockDim.x;
  auto one = [&](long long i, const float4& gin, const float4& xin, const float4& yin) {
    long long p; int c4;
    if (cq_shift >= 0) { p = i >> cq_shift; c4 = (int)(i & (cq - 1)); } else { p = i / cq; c4 = (int)(i - p * cq); }
    const int c = c4 * 4;
    const float4 m = ld4(s_mean + c), is4 = ld4(s_is + c), sc4 = ld4(s_sc + c);
    float g[4] = {gin.x, gin.y, gin.z, gin.w};
    const float xv[4] = {xin.x, xin.y, xin.z, xin.w};
    if (relu) {
      float4 t = yin;
      if (!y.p) {       // recompute the forward value (no residual): same expression as bn_apply*_fast_kernel
        const float4 b4 = ld4(s_bt + c);
        t.x = (xv[0] - m.x) * sc4.x + b4.x; t.y = (xv[1] - m.y) * sc4.y + b4.y;
        t.z = (xv[2] - m.z) * sc4.z + b4.z; t.w = (xv[3] - m.w) * sc4.w + b4.w;
      }
      if (!(t.x > 0.f)) g[0] = 0.f; if (!(t.y > 0.f)) g[1] = 0.f; if (!(t.z > 0.f)) g[2] = 0.f; if (!(t.w > 0.f)) g[3] = 0.f;
    }
    if (dx.p) {
      const float4 m0 = ld4(s_m0 + c), m1 = ld4(s_m1 + c);
      const float mm[4] = {m.x, m.y, m.z, m.w}, ii[4] = {is4.x, is4.y, is4.z, is4.w}, ss[4] = {sc4.x, sc4.y, sc4.z, sc4.w};
      const float a0[4] = {m0.x, m0.y, m0.z, m0.w}, a1[4] = {m1.x, m1.y, m1.z, m1.w};
      float o[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float xh = (xv[k] - mm[k]) * ii[k];
        o[k] = training ? ss[k] * (g[k] - a0[k] - xh * a1[k]) : ss[k] * g[k];
      }
      st4(dx.p + p * dx.ld + c, make_float4(o[0], o[1], o[2], o[3]));
    }
    if (dres.p) st4(dres.p + p * dres.ld + c, make_float4(g[0], g[1], g[2], g[3]));
  };
  auto addr = [&](const Rows& r, long long i) {
    long long p; int c4;
    if (cq_shift >= 0) { p = i >> cq_shift; c4 = (int)(i & (cq - 1)); } else { p = i / cq; c4 = (int)(i - p * cq); }
    return r.p + p * r.ld + c4 * 4;
  };
  const bool need_y = relu && y.p;
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  for (; i + stride < total4; i += 2 * stride) {      // two float4 per tensor in flight
    const float4 g0 = ld4(addr(dy, i)), g1 = ld4(addr(dy, i + stride));
    const float4 x0 = ld4(addr(x, i)), x1 = ld4(addr(x, i + stride));
    const float4 y0 = need_y ? ld4(addr(y, i)) : x0, y1 = need_y ? ld4(addr(y, i + stride)) : x1;
    one(i, g0, x0, y0);
    one(i + stride, g1, x1, y1);
  }
  if (i < total4) {
    const float4 g0 = ld4(addr(dy, i)), x0 = ld4(addr(x, i));
    const float4 y0 = need_y ? ld4(addr(y, i)) : x0;
    one(i, g0, x0, y0);
  }
}

static inline int shift_of(int v) { int s = 0; while ((1 << s) < v) ++s; return (1 << s) == v ? s : -1; }

static void reduce_geometry(long long P, int C, dim3& grid, long long& slab) {
  const int cq = C / 4;
  const int cblocks = cdiv(cq, 32);
  const int cq_w = cq < 32 ? cq : 32, ppw = 32 / cq_w;
  // 148 * 12 CTAs = whole waves at 3, 4 or 6 resident CTAs per SM (80 / 64 / 44 registers for MODE 1 / 0 / 2)
  long long want = (148LL * 12) / cblocks; if (want < 1) want = 1;
  long long s = cdiv(P, 8LL * ppw * 16); if (s > want) s = want; if (s < 1) s = 1;
  slab = (P + s - 1) / s;
  grid = dim3((unsigned)cdiv(P, slab), cblocks);
}

bool fast_reduce_ok(const View& x) {
  const int cq = x.c / 4;
  return pix_contig(x) && x.c % 4 == 0 && (cq >= 32 || (32 % cq) == 0) && (long long)x.n * x.h * x.w >= 512;
}

int bn_stats_fast(const View& x, double* sums, cudaStream_t st) {
  const long long P = (long long)x.n * x.h * x.w;
  dim3 grid; long long slab;
  reduce_geometry(P, x.c, grid, slab);
  Rows none; none.p = nullptr; none.ld = 0;
  colreduce_fast_kernel<0><<<grid, dim3(32, 8), 0, st>>>(rows_of(x), none, none, none, P, x.c, nullptr, nullptr, 0, sums,
                                                        nullptr, slab);
  return launched();
}
int bn_bwd_reduce_fast(const View& x, const View& y, const View& dy, const float* mean, const float* invstd, int act,
                       double* red, cudaStream_t st, const float* gamma, const float* beta) {
  const long long P = (long long)x.n * x.h * x.w;
  dim3 grid; long long slab;
  reduce_geometry(P, x.c, grid, slab);
  Rows none; none.p = nullptr; none.ld = 0;
  colreduce_fast_kernel<1><<<grid, dim3(32, 8), 0, st>>>(rows_of(x), y.p ? rows_of(y) : none, rows_of(dy), none, P, x.c, mean,
                                                        invstd, act, red, nullptr, slab, gamma, beta);
  return launched();
}
int act_bwd_bias_fast(const View& y, const View& dy, const View& dz, int act, float* dbias, cudaStream_t st) {
  const long long P = (long long)dy.n * dy.h * dy.w;
  dim3 grid; long long slab;
  reduce_geometry(P, dy.c, grid, slab);
  Rows none; none.p = nullptr; none.ld = 0;
  colreduce_fast_kernel<2><<<grid, dim3(32, 8), 0, st>>>(none, y.p ? rows_of(y) : none, rows_of(dy), dz.p ? rows_of(dz) : none,
                                                        P, dy.c, nullptr, nullptr, act, nullptr, dbias, slab);
  return launched();
}
int bn_apply_fast(const View& x, const View& res, const View& y, const float* mean, const float* invstd, const float* gamma,
                  const float* beta, int act, cudaStream_t st) {
  const long long P = (long long)x.n * x.h * x.w;
  const int cq = x.c / 4;
  const long long total4 = P * cq;
  long long blocks = cdiv(total4, 256 * 4); if (blocks > 148 * 16) blocks = 148 * 16; if (blocks < 1) blocks = 1;
  Rows none; none.p = nullptr; none.ld = 0;
  bn_apply_fast_kernel<<<(unsigned)blocks, 256, 0, st>>>(rows_of(x), res.p ? rows_of(res) : none, rows_of(y), total4, cq,
                                                        shift_of(cq), mean, invstd, gamma, beta, act);
  return launched();
}
int bn_apply_train_fast(const View& x, const View& res, const View& y, const double* sums, long long count, float eps,
                        float momentum, const float* gamma, const float* beta, int act, float* mean, float* invstd,
                        float* rmean, float* rvar, cudaStream_t st) {
  const long long P = (long long)x.n * x.h * x.w;
  const int cq = x.c / 4;
  const long long total4 = P * cq;
  long long blocks = cdiv(total4, 256 * 4); if (blocks > 148 * 16) blocks = 148 * 16; if (blocks < 1) blocks = 1;
  Rows none; none.p = nullptr; none.ld = 0;
  bn_apply_train_fast_kernel<<<(unsigned)blocks, 256, 3 * x.c * sizeof(float), st>>>(
      rows_of(x), res.p ? rows_of(res) : none, rows_of(y), total4, cq, shift_of(cq), sums, x.c, (double)count, eps, momentum,
      gamma, beta, act, mean, invstd, rmean, rvar);
  return launched();
}
int bn_bwd_apply_fast(const View& x, const View& y, const View& dy, const View& dx, const View& dres, const float* mean,
                      const float* invstd, const float* gamma, int relu, int training, const double* red, long long count,
                      cudaStream_t st, const float* beta) {
  const long long P = (long long)x.n * x.h * x.w;
  const int cq = x.c / 4;
  const long long total4 = P * cq;
  long long blocks = cdiv(total4, 256 * 4); if (blocks > 148 * 16) blocks = 148 * 16; if (blocks < 1) blocks = 1;
  Rows none; none.p = nullptr; none.ld = 0;
  bn_bwd_apply_fast_kernel<<<(unsigned)blocks, 256, 6 * x.c * sizeof(float), st>>>(rows_of(x), y.p ? rows_of(y) : none, rows_of(dy),
                                                            dx.p ? rows_of(dx) : none, dres.p ? rows_of(dres) : none, total4, cq,
                                                            shift_of(cq), x.c, mean, invstd, gamma, relu, training, red,
                                                            (float)(1.0 / (double)count), beta);
  return launched();
}

}  // namespace segsde
