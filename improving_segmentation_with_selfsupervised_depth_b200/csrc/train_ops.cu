// The step-level ops of SURVEY.md §8(a) rows T1-T4 that sit either side of the model / loss call in the reference's
// train.py (they are Trainer methods there, not part of models/ or loss/):
//   T1  feature-distance loss           torch.dist(enc_feat, imnet_feat, p=2)                 train.py:480-484
//   T2  DepthMix: per-sample disparity normalisation (train.py:688-692), depth-comparison mix mask
//       (generate_mix_mask "depthcomp", train.py:585-604) and the mix itself (loader/transformsgpu.py:33-47)
//   T3  pseudo-label selection          max / argmax of the teacher softmax, ignore where the max is 0, confidence
//       weight = share of pixels with max >= 0.968                                          train.py:644-651
//   T4  EMA teacher update              ema = a * ema + (1 - a) * p over all parameters       train.py:346-358
// All HBM-bound streaming kernels; the reference's host round trips (.item(), Python loops over the batch, ~600 tiny
// kernels for the EMA) are replaced by device-side scalars and one multi-tensor launch per 48 tensors.
#include "common.cuh"

namespace segsde {

// total order on floats for atomicMin / atomicMax
__device__ __forceinline__ unsigned int f2ord(float f) {
  const unsigned int b = __float_as_uint(f);
  return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float ord2f(unsigned int k) {
  return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k);
}

__device__ __forceinline__ double block_sum_d(double v) {
  __shared__ double red[32];
  v = warp_sum_d(v);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  double s = 0.0;
  if (threadIdx.x == 0)
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) s += red[i];
  return s;      // valid in thread 0
}

// ---- T1 -------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) sqdiff_sum_kernel(const float* __restrict__ a, const float* __restrict__ b,
                                                         long long n, double* __restrict__ sum) {
  double acc = 0.0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float d = a[i] - b[i];
    acc += (double)d * (double)d;
  }
  const double s = block_sum_d(acc);
  if (threadIdx.x == 0) atomicAdd(sum, s);
}
__global__ void dist_finalize_kernel(const double* __restrict__ sum, float* __restrict__ out) { out[0] = (float)sqrt(sum[0]); }
__global__ void __launch_bounds__(256) dist_bwd_kernel(const float* __restrict__ a, const float* __restrict__ b, long long n,
                                                       const float* __restrict__ dist, const float* __restrict__ g,
                                                       float* __restrict__ da, float* __restrict__ db) {
  const float d = dist[0];
  const float s = d > 0.f ? g[0] / d : 0.f;        // subgradient 0 at the origin, as torch's norm backward
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float v = s * (a[i] - b[i]);
    if (da) da[i] = v;
    if (db) db[i] = -v;
  }
}

// ---- T2 -------------------------------------------------------------------------------------------------------
// minmax: [B][2] ordered keys, initialised by the caller to {0xffffffff, 0}
__global__ void __launch_bounds__(256) sample_minmax_kernel(const float* __restrict__ d, long long hw,
                                                            unsigned int* __restrict__ minmax) {
  const int b = blockIdx.y;
  const float* p = d + (long long)b * hw;
  unsigned int lo = 0xffffffffu, hi = 0u;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < hw; i += (long long)gridDim.x * blockDim.x) {
    const unsigned int k = f2ord(p[i]);
    lo = min(lo, k); hi = max(hi, k);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    lo = min(lo, __shfl_xor_sync(0xffffffffu, lo, o));
    hi = max(hi, __shfl_xor_sync(0xffffffffu, hi, o));
  }
  if ((threadIdx.x & 31) == 0) { atomicMin(minmax + 2 * b, lo); atomicMax(minmax + 2 * b + 1, hi); }
}
__global__ void __launch_bounds__(256) sample_normalize_kernel(const float* __restrict__ d, float* __restrict__ out,
                                                               long long hw, const unsigned int* __restrict__ minmax) {
  const int b = blockIdx.y;
  const float lo = ord2f(minmax[2 * b]), hi = ord2f(minmax[2 * b + 1]);
  const float inv = 1.f / (hi - lo);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < hw; i += (long long)gridDim.x * blockDim.x)
    out[(long long)b * hw + i] = (fminf(fmaxf(d[(long long)b * hw + i], lo), hi) - lo) * inv;
}
// mask[i] = [d_i >= d_other - margin] * [d_i >= ft_i], other = (i + 1) % B; ft_i = ft_dev[i] when given (one threshold per
// sample, drawn on the device like the reference does), else the host scalar.  compare == 0 drops the first factor
// (mix-mask mode "depth": a plain per-sample disparity threshold).  Exactly one of mask_i / mask_f is written.
__global__ void __launch_bounds__(256) depthcomp_mask_kernel(const float* __restrict__ d, int B, long long hw, float margin,
                                                             float ft, const float* __restrict__ ft_dev, int compare,
                                                             long long* __restrict__ mask_i, float* __restrict__ mask_f) {
  const long long total = (long long)B * hw;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int b = (int)(i / hw); const long long p = i - (long long)b * hw;
    const float own = d[i];
    const float thr = ft_dev ? ft_dev[b] : ft;
    bool keep = own >= thr;
    if (compare) keep = keep && own >= d[(long long)((b + 1) % B) * hw + p] - margin;
    if (mask_i) mask_i[i] = keep ? 1 : 0;
    else mask_f[i] = keep ? 1.f : 0.f;
  }
}
// out[i,c,p] = m[i,p] * x[i,c,p] + (1 - m[i,p]) * x[(i+1)%B,c,p]; x / out addressed through (sn, sc, sp) element strides
__global__ void __launch_bounds__(256) mix_kernel(const float* __restrict__ x, float* __restrict__ out,
                                                  const long long* __restrict__ mask_i, const float* __restrict__ mask_f,
                                                  int B, int C, long long hw, long long xsn, long long xsc, long long xsp,
                                                  long long osn, long long osc, long long osp, int c_fastest) {
  const long long total = (long long)B * C * hw;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    int b, c; long long p;
    if (c_fastest) { c = (int)(i % C); const long long q = i / C; p = q % hw; b = (int)(q / hw); }
    else { p = i % hw; const long long q = i / hw; c = (int)(q % C); b = (int)(q / C); }
    const float m = mask_i ? (float)mask_i[(long long)b * hw + p] : mask_f[(long long)b * hw + p];
    const float own = x[b * xsn + c * xsc + p * xsp], oth = x[(long long)((b + 1) % B) * xsn + c * xsc + p * xsp];
    out[b * osn + c * osc + p * osp] = m * own + (1.f - m) * oth;
  }
}

// ---- T3 -------------------------------------------------------------------------------------------------------
// teacher softmax over the class axis (train.py:667): out[b,c,p] = exp(x - max_c) / sum_c, both tensors addressed by
// (sample, channel, pixel) element strides so NCHW-planar and channels-last logits need no copy; C <= 64
__global__ void __launch_bounds__(256) softmax_channels_kernel(const float* __restrict__ x, float* __restrict__ out, int B,
                                                               int C, long long hw, long long xsn, long long xsc,
                                                               long long xsp, long long osn, long long osc, long long osp) {
  const long long total = (long long)B * hw;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int b = (int)(i / hw); const long long p = i - (long long)b * hw;
    const float* q = x + b * xsn + p * xsp;
    float* o = out + b * osn + p * osp;
    float mx = q[0];
    for (int c = 1; c < C; ++c) mx = fmaxf(mx, q[c * xsc]);
    float se = 0.f;
    for (int c = 0; c < C; ++c) se += expf(q[c * xsc] - mx);
    const float inv = 1.f / se;
    for (int c = 0; c < C; ++c) o[c * osc] = expf(q[c * xsc] - mx) * inv;
  }
}
__global__ void __launch_bounds__(256) pseudo_label_kernel(const float* __restrict__ prob, int B, int C, long long hw,
                                                           long long sn, long long sc, long long sp, float thr,
                                                           long long ignore, long long* __restrict__ label,
                                                           float* __restrict__ max_prob,
                                                           unsigned long long* __restrict__ count) {
  const long long total = (long long)B * hw;
  unsigned int mine = 0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int b = (int)(i / hw); const long long p = i - (long long)b * hw;
    const float* q = prob + b * sn + p * sp;
    float best = q[0]; int bi = 0;
    for (int c = 1; c < C; ++c) { const float v = q[c * sc]; if (v > best) { best = v; bi = c; } }
    label[i] = best == 0.f ? ignore : (long long)bi;
    if (max_prob) max_prob[i] = best;
    mine += best >= thr ? 1u : 0u;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mine += __shfl_xor_sync(0xffffffffu, mine, o);
  if ((threadIdx.x & 31) == 0 && mine) atomicAdd(count, (unsigned long long)mine);
}
__global__ void __launch_bounds__(256) fill_ratio_kernel(const unsigned long long* __restrict__ count, double inv_total,
                                                         float scale, float* __restrict__ w, long long n) {
  const float v = scale * (float)((double)count[0] * inv_total);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) w[i] = v;
}

// ---- T4 -------------------------------------------------------------------------------------------------------
constexpr int MT_TENSORS = 48, MT_BLOCKS = 320, MT_CHUNK = 65536;
// The multi-tensor kernels walk a 64K-element chunk per CTA.  When every pointer of the tensor is 16-byte aligned (chunk
// bases are multiples of 4 elements) the chunk is walked as float4 with two vectors in flight per thread; the scalar loop
// is the path for odd offsets and the tail.  Same per-element expression either way.  (The scalar form ran the optimizer
// steps at 1.4 TB/s: 0.6 ms of Adam per dec5 step, 2.9 ms of SGD + clipping per joint step.)
__device__ __forceinline__ bool al16(const void* a, const void* b = nullptr, const void* c = nullptr, const void* d = nullptr) {
  return ((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(b) | reinterpret_cast<uintptr_t>(c) |
           reinterpret_cast<uintptr_t>(d)) & 15) == 0;
}
struct MultiAxpby {
  float* dst[MT_TENSORS];
  const float* src[MT_TENSORS];
  long long numel[MT_TENSORS];
  unsigned char blk_tensor[MT_BLOCKS];
  int blk_chunk[MT_BLOCKS];
};
// dst = alpha * dst + beta * src over a list of tensors; one CTA per (tensor, 64K-element chunk)
__global__ void __launch_bounds__(256) multi_axpby_kernel(const __grid_constant__ MultiAxpby t, float alpha, float beta) {
  const int ti = t.blk_tensor[blockIdx.x];
  const long long base = (long long)t.blk_chunk[blockIdx.x] * MT_CHUNK;
  const long long end = min(t.numel[ti], base + MT_CHUNK);
  float* __restrict__ d = t.dst[ti];
  const float* __restrict__ s = t.src[ti];
  long long i0 = base;
  if (al16(d, s)) {
    const long long n4 = (end - base) >> 2;
    float4* d4 = reinterpret_cast<float4*>(d + base);
    const float4* s4 = reinterpret_cast<const float4*>(s + base);
#pragma unroll 2
    for (long long j = threadIdx.x; j < n4; j += 256) {
      float4 a = d4[j]; const float4 b = s4[j];
      a.x = alpha * a.x + beta * b.x; a.y = alpha * a.y + beta * b.y; a.z = alpha * a.z + beta * b.z; a.w = alpha * a.w + beta * b.w;
      d4[j] = a;
    }
    i0 = base + n4 * 4;
  }
  for (long long i = i0 + threadIdx.x; i < end; i += blockDim.x) d[i] = alpha * d[i] + beta * s[i];
}

static unsigned grid_for(long long n) {
  long long b = cdiv(n, 256 * 8);
  if (b > 148 * 8) b = 148 * 8;
  if (b < 1) b = 1;
  return (unsigned)b;
}

}  // namespace segsde
using namespace segsde;

extern "C" int segsde_feature_distance_fwd(const float* a, const float* b, int64_t n, double* sum, float* dist, void* stream) {
  if (!a || !b || !sum || !dist || n < 1) return SEGSDE_E_ARG;
  cudaStream_t st = as_stream(stream);
  sqdiff_sum_kernel<<<grid_for(n), 256, 0, st>>>(a, b, n, sum);
  dist_finalize_kernel<<<1, 1, 0, st>>>(sum, dist);
  return launched();
}
extern "C" int segsde_feature_distance_bwd(const float* a, const float* b, int64_t n, const float* dist, const float* g,
                                           float* da, float* db, void* stream) {
  if (!a || !b || !dist || !g || (!da && !db) || n < 1) return SEGSDE_E_ARG;
  dist_bwd_kernel<<<grid_for(n), 256, 0, as_stream(stream)>>>(a, b, n, dist, g, da, db);
  return launched();
}
extern "C" int segsde_sample_minmax_normalize(const float* d, int b, int64_t hw, unsigned int* minmax, float* out, void* stream) {
  if (!d || !minmax || !out || b < 1 || b > 65535 || hw < 1) return SEGSDE_E_ARG;
  cudaStream_t st = as_stream(stream);
  long long gx = cdiv(hw, 256 * 8); if (gx > 148 * 4) gx = 148 * 4; if (gx < 1) gx = 1;
  dim3 grid((unsigned)gx, (unsigned)b);
  sample_minmax_kernel<<<grid, 256, 0, st>>>(d, hw, minmax);
  sample_normalize_kernel<<<grid, 256, 0, st>>>(d, out, hw, minmax);
  return launched();
}
extern "C" int segsde_depthcomp_mask(const float* d, int b, int64_t hw, float margin, float foreground_threshold,
                                     const float* threshold_dev, int compare, int64_t* mask_i64, float* mask_f32,
                                     void* stream) {
  if (!d || (!mask_i64 == !mask_f32) || b < 1 || hw < 1) return SEGSDE_E_ARG;
  depthcomp_mask_kernel<<<grid_for((long long)b * hw), 256, 0, as_stream(stream)>>>(
      d, b, hw, margin, foreground_threshold, threshold_dev, compare, (long long*)mask_i64, mask_f32);
  return launched();
}
extern "C" int segsde_mix(const float* x, float* out, const int64_t* mask_i64, const float* mask_f32, int b, int c,
                          int64_t hw, int64_t x_sn, int64_t x_sc, int64_t x_sp, int64_t o_sn, int64_t o_sc, int64_t o_sp,
                          void* stream) {
  if (!x || !out || (!mask_i64 && !mask_f32) || b < 1 || c < 1 || hw < 1) return SEGSDE_E_ARG;
  mix_kernel<<<grid_for((long long)b * c * hw), 256, 0, as_stream(stream)>>>(
      x, out, (const long long*)mask_i64, mask_f32, b, c, hw, x_sn, x_sc, x_sp, o_sn, o_sc, o_sp, (x_sc == 1 && o_sc == 1) ? 1 : 0);
  return launched();
}
extern "C" int segsde_pseudo_label(const float* prob, int b, int c, int64_t hw, int64_t sn, int64_t sc, int64_t sp,
                                   float threshold, int64_t ignore_index, int64_t* label, float* max_prob,
                                   unsigned long long* count, float weight_scale, float* pixel_weight, void* stream) {
  if (!prob || !label || !count || b < 1 || c < 1 || hw < 1) return SEGSDE_E_ARG;
  cudaStream_t st = as_stream(stream);
  const long long total = (long long)b * hw;
  pseudo_label_kernel<<<grid_for(total), 256, 0, st>>>(prob, b, c, hw, sn, sc, sp, threshold, ignore_index, (long long*)label,
                                                       max_prob, count);
  if (pixel_weight) fill_ratio_kernel<<<grid_for(total), 256, 0, st>>>(count, 1.0 / (double)total, weight_scale, pixel_weight, total);
  return launched();
}
extern "C" int segsde_softmax_channels(const float* x, float* out, int b, int c, int64_t hw, int64_t x_sn, int64_t x_sc,
                                       int64_t x_sp, int64_t o_sn, int64_t o_sc, int64_t o_sp, void* stream) {
  if (!x || !out || b < 1 || c < 1 || hw < 1) return SEGSDE_E_ARG;
  softmax_channels_kernel<<<grid_for((long long)b * hw), 256, 0, as_stream(stream)>>>(x, out, b, c, hw, x_sn, x_sc, x_sp,
                                                                                      o_sn, o_sc, o_sp);
  return launched();
}
extern "C" int segsde_multi_axpby(int ntensors, float* const* dst, const float* const* src, const int64_t* numel, float alpha,
                                  float beta, void* stream) {
  if (ntensors < 0 || (ntensors && (!dst || !src || !numel))) return SEGSDE_E_ARG;
  cudaStream_t st = as_stream(stream);
  MultiAxpby t;
  int nt = 0, nb = 0;
  for (int i = 0; i < ntensors; ++i) {
    if (numel[i] < 0 || (numel[i] && (!dst[i] || !src[i]))) return SEGSDE_E_ARG;
    const long long chunks = (numel[i] + MT_CHUNK - 1) / MT_CHUNK;
    long long done = 0;
    while (done < chunks) {
      if (nt == MT_TENSORS || nb == MT_BLOCKS) {       // table full: flush
        multi_axpby_kernel<<<nb, 256, 0, st>>>(t, alpha, beta);
        nt = nb = 0;
      }
      t.dst[nt] = dst[i]; t.src[nt] = src[i]; t.numel[nt] = numel[i];
      while (done < chunks && nb < MT_BLOCKS) { t.blk_tensor[nb] = (unsigned char)nt; t.blk_chunk[nb] = (int)done; ++nb; ++done; }
      ++nt;
    }
  }
  if (nb) multi_axpby_kernel<<<nb, 256, 0, st>>>(t, alpha, beta);
  return launched();
}

// ------------------------------------------------------------------------------------------------------------
// Optimizer step and gradient-norm clipping as multi-tensor kernels (the reference instantiates
// torch.optim.Adam / SGD through utils/optimizers.py:7-30 and clips with torch.nn.utils.clip_grad_norm_,
// train.py:516-524).  Same arithmetic as torch's single-tensor reference implementations.
// ------------------------------------------------------------------------------------------------------------
namespace segsde {

constexpr int MO_TENSORS = 36;
struct MultiOpt {
  float* p[MO_TENSORS];
  float* g[MO_TENSORS];        // gradients (read; scaled in place by multi_scale)
  float* s1[MO_TENSORS];       // Adam exp_avg / SGD momentum buffer (nullable)
  float* s2[MO_TENSORS];       // Adam exp_avg_sq (nullable)
  long long numel[MO_TENSORS];
  unsigned char blk_tensor[MT_BLOCKS];
  int blk_chunk[MT_BLOCKS];
};
struct AdamHyper { float beta1, beta2, eps, weight_decay, step_size, bias_c2_sqrt; };   // step_size = lr / (1 - beta1^t)
struct SgdHyper { float lr, momentum, dampening, weight_decay; int nesterov, first; };

__global__ void __launch_bounds__(256) multi_adam_kernel(const __grid_constant__ MultiOpt t, AdamHyper h) {
  const int ti = t.blk_tensor[blockIdx.x];
  const long long base = (long long)t.blk_chunk[blockIdx.x] * MT_CHUNK, end = min(t.numel[ti], base + MT_CHUNK);
  float* __restrict__ p = t.p[ti]; const float* __restrict__ g = t.g[ti];
  float* __restrict__ m = t.s1[ti]; float* __restrict__ v = t.s2[ti];
  const float step_size = h.step_size;
  auto one = [&](float& pi, float gi, float& mi, float& vi) {
    if (h.weight_decay != 0.f) gi = fmaf(h.weight_decay, pi, gi);
    mi = mi + (gi - mi) * (1.f - h.beta1);                                 // lerp, as torch's exp_avg.lerp_
    vi = h.beta2 * vi + (1.f - h.beta2) * (gi * gi);
    const float denom = sqrtf(vi) / h.bias_c2_sqrt + h.eps;
    pi = pi - step_size * (mi / denom);
  };
  long long i0 = base;
  if (al16(p, g, m, v)) {
    const long long n4 = (end - base) >> 2;
    float4* p4 = reinterpret_cast<float4*>(p + base); const float4* g4 = reinterpret_cast<const float4*>(g + base);
    float4* m4 = reinterpret_cast<float4*>(m + base); float4* v4 = reinterpret_cast<float4*>(v + base);
#pragma unroll 2
    for (long long j = threadIdx.x; j < n4; j += 256) {
      float4 pp = p4[j], mm = m4[j], vv = v4[j]; const float4 gg = g4[j];
      one(pp.x, gg.x, mm.x, vv.x); one(pp.y, gg.y, mm.y, vv.y); one(pp.z, gg.z, mm.z, vv.z); one(pp.w, gg.w, mm.w, vv.w);
      m4[j] = mm; v4[j] = vv; p4[j] = pp;
    }
    i0 = base + n4 * 4;
  }
  for (long long i = i0 + threadIdx.x; i < end; i += blockDim.x) {
    float pi = p[i], mi = m[i], vi = v[i];
    one(pi, g[i], mi, vi);
    m[i] = mi; v[i] = vi; p[i] = pi;
  }
}
__global__ void __launch_bounds__(256) multi_sgd_kernel(const __grid_constant__ MultiOpt t, SgdHyper h) {
  const int ti = t.blk_tensor[blockIdx.x];
  const long long base = (long long)t.blk_chunk[blockIdx.x] * MT_CHUNK, end = min(t.numel[ti], base + MT_CHUNK);
  float* __restrict__ p = t.p[ti]; const float* __restrict__ g = t.g[ti];
  float* __restrict__ buf = t.s1[ti];
  const bool mom = h.momentum != 0.f;
  auto one = [&](float& pi, float gi, float& bi) {
    if (h.weight_decay != 0.f) gi = fmaf(h.weight_decay, pi, gi);
    if (mom) {
      const float b = h.first ? gi : h.momentum * bi + (1.f - h.dampening) * gi;
      bi = b;
      gi = h.nesterov ? gi + h.momentum * b : b;
    }
    pi = pi - h.lr * gi;
  };
  long long i0 = base;
  if (al16(p, g, mom ? buf : nullptr)) {
    const long long n4 = (end - base) >> 2;
    float4* p4 = reinterpret_cast<float4*>(p + base); const float4* g4 = reinterpret_cast<const float4*>(g + base);
    float4* b4 = reinterpret_cast<float4*>(buf + base);
#pragma unroll 2
    for (long long j = threadIdx.x; j < n4; j += 256) {
      float4 pp = p4[j]; const float4 gg = g4[j];
      float4 bb = make_float4(0.f, 0.f, 0.f, 0.f);
      if (mom && !h.first) bb = b4[j];
      one(pp.x, gg.x, bb.x); one(pp.y, gg.y, bb.y); one(pp.z, gg.z, bb.z); one(pp.w, gg.w, bb.w);
      if (mom) b4[j] = bb;
      p4[j] = pp;
    }
    i0 = base + n4 * 4;
  }
  for (long long i = i0 + threadIdx.x; i < end; i += blockDim.x) {
    float pi = p[i], bi = (mom && !h.first) ? buf[i] : 0.f;
    one(pi, g[i], bi);
    if (mom) buf[i] = bi;
    p[i] = pi;
  }
}
__global__ void __launch_bounds__(256) multi_sqnorm_kernel(const __grid_constant__ MultiOpt t, double* __restrict__ sum) {
  const int ti = t.blk_tensor[blockIdx.x];
  const long long base = (long long)t.blk_chunk[blockIdx.x] * MT_CHUNK, end = min(t.numel[ti], base + MT_CHUNK);
  const float* __restrict__ g = t.g[ti];
  double acc = 0.0;
  long long i0 = base;
  if (al16(g)) {
    const long long n4 = (end - base) >> 2;
    const float4* g4 = reinterpret_cast<const float4*>(g + base);
#pragma unroll 4
    for (long long j = threadIdx.x; j < n4; j += 256) {
      const float4 q = g4[j];
      acc += (double)q.x * (double)q.x + (double)q.y * (double)q.y + (double)q.z * (double)q.z + (double)q.w * (double)q.w;
    }
    i0 = base + n4 * 4;
  }
  for (long long i = i0 + threadIdx.x; i < end; i += blockDim.x) acc += (double)g[i] * (double)g[i];
  const double s = block_sum_d(acc);
  if (threadIdx.x == 0) atomicAdd(sum, s);
}
// total_norm = sqrt(sum); coef = min(1, max_norm / (total_norm + 1e-6))   (torch.nn.utils.clip_grad_norm_)
__global__ void clip_coef_kernel(const double* __restrict__ sum, float max_norm, float* __restrict__ total_norm,
                                 float* __restrict__ coef) {
  const float n = (float)sqrt(sum[0]);
  total_norm[0] = n;
  coef[0] = fminf(1.f, max_norm / (n + 1e-6f));
}
__global__ void __launch_bounds__(256) multi_scale_kernel(const __grid_constant__ MultiOpt t, const float* __restrict__ coef) {
  const int ti = t.blk_tensor[blockIdx.x];
  const long long base = (long long)t.blk_chunk[blockIdx.x] * MT_CHUNK, end = min(t.numel[ti], base + MT_CHUNK);
  float* __restrict__ g = t.g[ti];
  const float c = coef[0];
  if (c == 1.f) return;           // gradient norm below the threshold: nothing to rewrite
  long long i0 = base;
  if (al16(g)) {
    const long long n4 = (end - base) >> 2;
    float4* g4 = reinterpret_cast<float4*>(g + base);
#pragma unroll 4
    for (long long j = threadIdx.x; j < n4; j += 256) {
      float4 q = g4[j];
      q.x *= c; q.y *= c; q.z *= c; q.w *= c;
      g4[j] = q;
    }
    i0 = base + n4 * 4;
  }
  for (long long i = i0 + threadIdx.x; i < end; i += blockDim.x) g[i] *= c;
}

// GradScaler.unscale_ (torch._amp_foreach_non_finite_check_and_unscale_): g *= 1 / scale[0]; found_inf[0] = 1 when any
// element is not finite (the check runs on the unscaled value, like torch's)
__global__ void __launch_bounds__(256) multi_unscale_kernel(const __grid_constant__ MultiOpt t, const float* __restrict__ scale,
                                                            float* __restrict__ found_inf) {
  const int ti = t.blk_tensor[blockIdx.x];
  const long long base = (long long)t.blk_chunk[blockIdx.x] * MT_CHUNK, end = min(t.numel[ti], base + MT_CHUNK);
  float* __restrict__ g = t.g[ti];
  const float c = 1.f / scale[0];
  bool bad = false;
  for (long long i = base + threadIdx.x; i < end; i += blockDim.x) {
    const float v = g[i];
    bad |= !isfinite(v);
    g[i] = (c == 1.f) ? v : v * c;
  }
  if (__any_sync(0xffffffffu, bad) && (threadIdx.x & 31) == 0) found_inf[0] = 1.f;
}
// GradScaler.update (torch._amp_update_scale_): back off on overflow, grow after `interval` clean steps
__global__ void amp_update_scale_kernel(float* __restrict__ scale, int* __restrict__ tracker, const float* __restrict__ found_inf,
                                        float growth, float backoff, int interval) {
  if (found_inf[0] != 0.f) {
    scale[0] *= backoff;
    tracker[0] = 0;
  } else {
    const int ok = tracker[0] + 1;
    if (ok == interval) {
      const float grown = scale[0] * growth;
      if (isfinite(grown)) scale[0] = grown;
      tracker[0] = 0;
    } else {
      tracker[0] = ok;
    }
  }
}

// walks the tensor list, filling tables of <= MO_TENSORS tensors / MT_BLOCKS chunks and launching `launch(table, nb)`
template <class F>
static int for_each_table(int n, float* const* p, float* const* g, float* const* s1, float* const* s2, const int64_t* numel,
                          F launch) {
  MultiOpt t;
  int nt = 0, nb = 0;
  for (int i = 0; i < n; ++i) {
    if (numel[i] < 0) return SEGSDE_E_ARG;
    const long long chunks = (numel[i] + MT_CHUNK - 1) / MT_CHUNK;
    long long done = 0;
    while (done < chunks) {
      if (nt == MO_TENSORS || nb == MT_BLOCKS) { launch(t, nb); nt = nb = 0; }
      t.p[nt] = p ? p[i] : nullptr; t.g[nt] = g ? g[i] : nullptr;
      t.s1[nt] = s1 ? s1[i] : nullptr; t.s2[nt] = s2 ? s2[i] : nullptr; t.numel[nt] = numel[i];
      while (done < chunks && nb < MT_BLOCKS) { t.blk_tensor[nb] = (unsigned char)nt; t.blk_chunk[nb] = (int)done; ++nb; ++done; }
      ++nt;
    }
  }
  if (nb) launch(t, nb);
  return launched();
}

}  // namespace segsde

extern "C" int segsde_multi_adam(int n, float* const* p, float* const* g, float* const* exp_avg, float* const* exp_avg_sq,
                                 const int64_t* numel, float lr, float beta1, float beta2, float eps, float weight_decay,
                                 int64_t step, void* stream) {
  if (n < 0 || (n && (!p || !g || !exp_avg || !exp_avg_sq || !numel)) || step < 1) return SEGSDE_E_ARG;
  cudaStream_t st = as_stream(stream);
  AdamHyper h;
  h.beta1 = beta1; h.beta2 = beta2; h.eps = eps; h.weight_decay = weight_decay;
  // bias corrections in double on the host, like torch's Python scalars
  h.step_size = (float)((double)lr / (1.0 - pow((double)beta1, (double)step)));
  h.bias_c2_sqrt = (float)sqrt(1.0 - pow((double)beta2, (double)step));
  return for_each_table(n, p, g, exp_avg, exp_avg_sq, numel,
                        [&](const MultiOpt& t, int nb) { multi_adam_kernel<<<nb, 256, 0, st>>>(t, h); });
}
extern "C" int segsde_multi_sgd(int n, float* const* p, float* const* g, float* const* momentum_buf, const int64_t* numel,
                                float lr, float momentum, float dampening, float weight_decay, int nesterov, int first_step,
                                void* stream) {
  if (n < 0 || (n && (!p || !g || !numel)) || (momentum != 0.f && n && !momentum_buf)) return SEGSDE_E_ARG;
  cudaStream_t st = as_stream(stream);
  SgdHyper h;
  h.lr = lr; h.momentum = momentum; h.dampening = dampening; h.weight_decay = weight_decay; h.nesterov = nesterov;
  h.first = first_step;
  return for_each_table(n, p, g, momentum_buf, nullptr, numel,
                        [&](const MultiOpt& t, int nb) { multi_sgd_kernel<<<nb, 256, 0, st>>>(t, h); });
}
extern "C" int segsde_multi_unscale(int n, float* const* g, const int64_t* numel, const float* scale, float* found_inf,
                                    void* stream) {
  if (n < 0 || (n && (!g || !numel)) || !scale || !found_inf) return SEGSDE_E_ARG;
  cudaStream_t st = as_stream(stream);
  return for_each_table(n, nullptr, g, nullptr, nullptr, numel,
                        [&](const MultiOpt& t, int nb) { multi_unscale_kernel<<<nb, 256, 0, st>>>(t, scale, found_inf); });
}
extern "C" int segsde_amp_update_scale(float* scale, int* growth_tracker, const float* found_inf, float growth_factor,
                                       float backoff_factor, int growth_interval, void* stream) {
  if (!scale || !growth_tracker || !found_inf || growth_interval < 1) return SEGSDE_E_ARG;
  amp_update_scale_kernel<<<1, 1, 0, as_stream(stream)>>>(scale, growth_tracker, found_inf, growth_factor, backoff_factor,
                                                         growth_interval);
  return launched();
}
extern "C" int segsde_multi_clip_grad_norm(int n, float* const* g, const int64_t* numel, float max_norm, double* sum,
                                           float* total_norm, float* coef, void* stream) {
  if (n < 0 || (n && (!g || !numel)) || !sum || !total_norm || !coef) return SEGSDE_E_ARG;
  cudaStream_t st = as_stream(stream);
  int rc = for_each_table(n, nullptr, g, nullptr, nullptr, numel,
                          [&](const MultiOpt& t, int nb) { multi_sqnorm_kernel<<<nb, 256, 0, st>>>(t, sum); });
  if (rc) return rc;
  clip_coef_kernel<<<1, 1, 0, st>>>(sum, max_norm, total_norm, coef);
  return for_each_table(n, nullptr, g, nullptr, nullptr, numel,
                        [&](const MultiOpt& t, int nb) { multi_scale_kernel<<<nb, 256, 0, st>>>(t, coef); });
}
