// Edge-aware smoothness (monodepth_layers.py:208-221 + mean normalisation monodepth_loss.py:182-184),
// segmentation cross-entropy (loss/loss.py:17-37) and the axpby used to scale stored unit gradients.
#include "common.cuh"

namespace segsde {

// ---- smoothness -------------------------------------------------------------------------------
__global__ void smooth_mean_kernel(const float* __restrict__ disp, int hw, float* __restrict__ mean_out) {
  // one block per sample; fixed-order reduction (deterministic)
  const int b = blockIdx.x;
  const float* d = disp + (size_t)b * hw;
  // per-thread partial sums in four independent fp32 chains (<= 128 terms each at 512x1024: disparities in (0,1), the
  // rounding of such a chain is ~1e-6 relative), fp64 only across threads — a dependent chain of fp64 adds per element
  // ran at 100 GB/s on the B200's reduced fp64 pipe
  float f0 = 0.f, f1 = 0.f, f2 = 0.f, f3 = 0.f;
  int i = threadIdx.x;
  for (; i + 3 * (int)blockDim.x < hw; i += 4 * blockDim.x) {
    f0 += d[i]; f1 += d[i + blockDim.x]; f2 += d[i + 2 * blockDim.x]; f3 += d[i + 3 * blockDim.x];
  }
  for (; i < hw; i += blockDim.x) f0 += d[i];
  double a = ((double)f0 + (double)f1) + ((double)f2 + (double)f3);
  __shared__ double sh[32];
  a = warp_sum_d(a);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = a;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) t += sh[i];
    mean_out[b] = (float)(t / (double)hw);
  }
}

// loss terms:  Lx = sum |dhat[x]-dhat[x+1]| * exp(-mean_c |I[x]-I[x+1]|),  Ly likewise;
// ghat = d(Lx/cnt_x + Ly/cnt_y)/d dhat   (cnt_x = B*h*(w-1), cnt_y = B*(h-1)*w)
__global__ void smooth_fused_kernel(const float* __restrict__ disp, const float* __restrict__ img,
                                    const float* __restrict__ mean, int B, int h, int w,
                                    float* __restrict__ acc, float* __restrict__ ghat) {
  const int b = blockIdx.z;
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  const int y = blockIdx.y * blockDim.y + threadIdx.y;
  const size_t hw = (size_t)h * w;
  const float inv_m = 1.f / (mean[b] + 1e-7f);
  const float* d = disp + b * hw;
  const float* I = img + (size_t)b * 3 * hw;
  float lx = 0.f, ly = 0.f, g = 0.f, dot = 0.f;
  if (x < w && y < h) {
    const size_t o = (size_t)y * w + x;
    const float dc = d[o] / (mean[b] + 1e-7f);
    const float wx_cnt = 1.f / ((float)B * (float)h * (float)(w - 1));
    const float wy_cnt = 1.f / ((float)B * (float)(h - 1) * (float)w);
    // edge (x, x+1): this pixel is the left end
    if (x + 1 < w) {
      const float dn = d[o + 1] / (mean[b] + 1e-7f);
      float gi = 0.f;
#pragma unroll
      for (int c = 0; c < 3; ++c) gi += fabsf(I[c * hw + o] - I[c * hw + o + 1]);
      const float e = expf(-(gi / 3.f));
      const float df = dc - dn;
      lx = fabsf(df) * e * wx_cnt;
      g += (df > 0.f ? 1.f : (df < 0.f ? -1.f : 0.f)) * e * wx_cnt;
    }
    if (x > 0) {  // edge (x-1, x): right end
      const float dp = d[o - 1] / (mean[b] + 1e-7f);
      float gi = 0.f;
#pragma unroll
      for (int c = 0; c < 3; ++c) gi += fabsf(I[c * hw + o - 1] - I[c * hw + o]);
      const float e = expf(-(gi / 3.f));
      const float df = dp - dc;
      g -= (df > 0.f ? 1.f : (df < 0.f ? -1.f : 0.f)) * e * wx_cnt;
    }
    if (y + 1 < h) {
      const float dn = d[o + w] / (mean[b] + 1e-7f);
      float gi = 0.f;
#pragma unroll
      for (int c = 0; c < 3; ++c) gi += fabsf(I[c * hw + o] - I[c * hw + o + w]);
      const float e = expf(-(gi / 3.f));
      const float df = dc - dn;
      ly = fabsf(df) * e * wy_cnt;
      g += (df > 0.f ? 1.f : (df < 0.f ? -1.f : 0.f)) * e * wy_cnt;
    }
    if (y > 0) {
      const float dp = d[o - w] / (mean[b] + 1e-7f);
      float gi = 0.f;
#pragma unroll
      for (int c = 0; c < 3; ++c) gi += fabsf(I[c * hw + o - w] - I[c * hw + o]);
      const float e = expf(-(gi / 3.f));
      const float df = dp - dc;
      g -= (df > 0.f ? 1.f : (df < 0.f ? -1.f : 0.f)) * e * wy_cnt;
    }
    if (ghat) ghat[b * hw + o] = g;
    dot = g * d[o];
    (void)inv_m;
  }
  __shared__ float sh[3][8];
  const int tid = threadIdx.y * blockDim.x + threadIdx.x;
  lx = warp_sum(lx); ly = warp_sum(ly); dot = warp_sum(dot);
  if ((tid & 31) == 0) { sh[0][tid >> 5] = lx; sh[1][tid >> 5] = ly; sh[2][tid >> 5] = dot; }
  __syncthreads();
  if (tid < 3) {
    float t = 0.f;
    for (int i = 0; i < 8; ++i) t += sh[tid][i];
    atomicAdd(tid < 2 ? acc + tid : acc + 2 + b, t);
  }
}

// d dhat_i / d d_j = delta_ij/(m+eps) - d_i/((m+eps)^2 hw)
__global__ void smooth_grad_finalize_kernel(const float* __restrict__ ghat, const float* __restrict__ mean,
                                            const float* __restrict__ acc, int hw, float wgt,
                                            float* __restrict__ gdisp) {
  const int b = blockIdx.y;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= hw) return;
  const float me = mean[b] + 1e-7f;
  const float dot = acc[2 + b];
  gdisp[(size_t)b * hw + i] += wgt * (ghat[(size_t)b * hw + i] / me - dot / (me * me * (float)hw));
}

__global__ void axpby_kernel(const float* __restrict__ x, float a, float* __restrict__ y, int accumulate,
                             long long n) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (; i < n; i += stride) y[i] = accumulate ? y[i] + a * x[i] : a * x[i];
}

// y (+)= (ca*a[0] + cb*b[0]) * x : scaling by upstream-gradient scalars that live on the device
__global__ void scale_by_dev_kernel(const float* __restrict__ x, const float* __restrict__ a, float ca,
                                    const float* __restrict__ b, float cb, float* __restrict__ y,
                                    int accumulate, long long n) {
  const float s = (a ? ca * a[0] : 0.f) + (b ? cb * b[0] : 0.f);
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (; i < n; i += stride) y[i] = accumulate ? y[i] + s * x[i] : s * x[i];
}

// out[s] = reproj[s] + wsm[s]*(sacc[s][0]+sacc[s][1]);  out[S] = mean_s out[s]
struct CombineW { float w[8]; };
__global__ void mono_combine_kernel(const float* __restrict__ reproj, const float* __restrict__ sacc,
                                    int S, int sacc_stride, CombineW wsm, float* __restrict__ out) {
  if (threadIdx.x == 0) {
    float tot = 0.f;
    for (int s = 0; s < S; ++s) {
      const float l = reproj[s] + wsm.w[s] * (sacc[s * sacc_stride] + sacc[s * sacc_stride + 1]);
      out[s] = l; tot += l;
    }
    out[S] = tot / (float)S;
  }
}

__global__ void ratio_kernel(const float* __restrict__ acc, float const_den, float* __restrict__ out,
                             float* __restrict__ inv) {
  const float den = const_den > 0.f ? const_den : acc[1];
  out[0] = acc[0] / den;     // 0/0 -> NaN exactly like F.cross_entropy on an all-ignored batch
  inv[0] = 1.f / den; inv[1] = 0.f;
}

// ---- cross entropy ----------------------------------------------------------------------------
// One thread per pixel; the C logits of a pixel are contiguous (NHWC view), C is small (19).
__global__ void nan_flag_kernel(const float* __restrict__ x, long long n, float* __restrict__ flag) {
  bool bad = false;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    bad |= isnan(x[i]);
  if (__any_sync(0xffffffffu, bad) && (threadIdx.x & 31) == 0) flag[0] = 1.f;
}
template <bool BWD>
__global__ void ce_kernel(View lg, const long long* __restrict__ target, const float* __restrict__ pw,
                          int ignore_index, float* __restrict__ acc, const float* __restrict__ gscale,
                          View dl, const float* __restrict__ flags) {
  const long long npix = (long long)lg.n * lg.h * lg.w;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  float nll = 0.f, valid = 0.f;
  if (i < npix) {
    const int x = (int)(i % lg.w);
    const int y = (int)((i / lg.w) % lg.h);
    const int n = (int)(i / ((long long)lg.w * lg.h));
    const float* p = lg.p + lg.off(n, y, x);
    const long long t = target[i];
    const int C = lg.c;
    // labels outside [0, C) other than ignore_index contribute nothing (torch raises a device assert there; here they
    // are counted in acc[3] so the host can turn them into an error without an out-of-bounds read)
    const bool in_range = t >= 0 && t < C;
    const bool ok = (t != ignore_index) && in_range;
    if (!BWD && t != ignore_index && !in_range) atomicAdd(acc + 3, 1.f);
    float mx = -3.4e38f;
    for (int c = 0; c < C; ++c) mx = fmaxf(mx, p[c]);
    float se = 0.f;
    for (int c = 0; c < C; ++c) se += expf(p[c] - mx);
    const float lse = mx + logf(se);
    // NaN anywhere in the pixel weights disables the weighting (reference loss/loss.py:31-32): flags[2] != 0
    const float wgt = (pw && flags[2] == 0.f) ? pw[i] : 1.f;
    if (!BWD) {
      if (ok) { nll = (lse - p[t]) * wgt; valid = 1.f; }
    } else {
      float* g = dl.p + dl.off(n, y, x);
      const float s = ok ? gscale[0] * wgt : 0.f;
      for (int c = 0; c < C; ++c) {
        const float sm = expf(p[c] - lse);
        g[c] = s * (sm - ((long long)c == t ? 1.f : 0.f));
      }
    }
  }
  if (!BWD) {
    __shared__ float sh[2][8];
    nll = warp_sum(nll); valid = warp_sum(valid);
    if ((threadIdx.x & 31) == 0) { sh[0][threadIdx.x >> 5] = nll; sh[1][threadIdx.x >> 5] = valid; }
    __syncthreads();
    if (threadIdx.x < 2) {
      float t = 0.f;
      for (int k = 0; k < (int)(blockDim.x >> 5); ++k) t += sh[threadIdx.x][k];
      atomicAdd(acc + threadIdx.x, t);
    }
  }
}

}  // namespace segsde
using namespace segsde;

extern "C" int segsde_smooth_mean(const float* disp, int B, int h, int w, float* mean_out, void* stream) {
  if (!disp || !mean_out || B < 1 || h < 1 || w < 1) return SEGSDE_E_ARG;
  smooth_mean_kernel<<<B, 1024, 0, as_stream(stream)>>>(disp, h * w, mean_out);
  return launched();
}
extern "C" int segsde_smooth_fused(const float* disp, const float* img, const float* mean, int B, int h,
                                   int w, float* acc, float* ghat, void* stream) {
  if (!disp || !img || !mean || !acc || B < 1 || h < 2 || w < 2) return SEGSDE_E_ARG;
  dim3 block(32, 8), grid(cdiv(w, 32), cdiv(h, 8), B);
  smooth_fused_kernel<<<grid, block, 0, as_stream(stream)>>>(disp, img, mean, B, h, w, acc, ghat);
  return launched();
}
extern "C" int segsde_smooth_grad_finalize(const float* ghat, const float* mean, const float* acc, int B,
                                           int h, int w, float wgt, float* gdisp, void* stream) {
  if (!ghat || !mean || !acc || !gdisp) return SEGSDE_E_ARG;
  dim3 grid(cdiv((int64_t)h * w, 256), B);
  smooth_grad_finalize_kernel<<<grid, 256, 0, as_stream(stream)>>>(ghat, mean, acc, h * w, wgt, gdisp);
  return launched();
}
extern "C" int segsde_axpby(const float* x, float a, float* y, int accumulate, int64_t n, void* stream) {
  if (!x || !y || n < 0) return SEGSDE_E_ARG;
  if (n == 0) return SEGSDE_OK;
  int blocks = (int)((n + 255) / 256);
  if (blocks > 148 * 16) blocks = 148 * 16;
  axpby_kernel<<<blocks, 256, 0, as_stream(stream)>>>(x, a, y, accumulate, (long long)n);
  return launched();
}

extern "C" int segsde_scale_by_dev(const float* x, const float* a, float ca, const float* b, float cb,
                                   float* y, int accumulate, int64_t n, void* stream) {
  if (!x || !y || n < 0) return SEGSDE_E_ARG;
  if (n == 0) return SEGSDE_OK;
  int blocks = (int)((n + 255) / 256);
  if (blocks > 148 * 16) blocks = 148 * 16;
  scale_by_dev_kernel<<<blocks, 256, 0, as_stream(stream)>>>(x, a, ca, b, cb, y, accumulate, (long long)n);
  return launched();
}
extern "C" int segsde_mono_combine(const float* reproj, const float* sacc, int S, int sacc_stride,
                                   const float* smooth_w_host, float* out, void* stream) {
  if (!reproj || !sacc || !smooth_w_host || !out || S < 1 || S > 8) return SEGSDE_E_ARG;
  CombineW w;
  for (int s = 0; s < 8; ++s) w.w[s] = s < S ? smooth_w_host[s] : 0.f;
  mono_combine_kernel<<<1, 32, 0, as_stream(stream)>>>(reproj, sacc, S, sacc_stride, w, out);
  return launched();
}

extern "C" int segsde_ratio(const float* acc, float const_den, float* out, float* inv, void* stream) {
  if (!acc || !out || !inv) return SEGSDE_E_ARG;
  ratio_kernel<<<1, 1, 0, as_stream(stream)>>>(acc, const_den, out, inv);
  return launched();
}

extern "C" int segsde_ce_fwd(const segsde_nhwc_t* logits, const int64_t* target, const float* pixel_w,
                             int ignore_index, float* acc, void* stream) {
  if (!logits || !logits->ptr || !target || !acc) return SEGSDE_E_ARG;
  View lg = mk(logits), none = mk(nullptr);
  const long long npix = (long long)lg.n * lg.h * lg.w;
  if (pixel_w) {
    int blocks = cdiv(npix, 1024);
    if (blocks > 148 * 8) blocks = 148 * 8;
    nan_flag_kernel<<<blocks, 256, 0, as_stream(stream)>>>(pixel_w, npix, acc + 2);
    int rc = launched();
    if (rc != SEGSDE_OK) return rc;
  }
  ce_kernel<false><<<cdiv(npix, 256), 256, 0, as_stream(stream)>>>(
      lg, (const long long*)target, pixel_w, ignore_index, acc, nullptr, none, acc);
  return launched();
}
extern "C" int segsde_ce_bwd(const segsde_nhwc_t* logits, const int64_t* target, const float* pixel_w,
                             int ignore_index, const float* gscale_dev, const float* acc,
                             const segsde_nhwc_t* dlogits, void* stream) {
  if (!logits || !logits->ptr || !target || !gscale_dev || !dlogits || !dlogits->ptr) return SEGSDE_E_ARG;
  if (pixel_w && !acc) return SEGSDE_E_ARG;
  View lg = mk(logits), dl = mk(dlogits);
  if (!same_shape(lg, dl)) return SEGSDE_E_ARG;
  const long long npix = (long long)lg.n * lg.h * lg.w;
  ce_kernel<true><<<cdiv(npix, 256), 256, 0, as_stream(stream)>>>(
      lg, (const long long*)target, pixel_w, ignore_index, nullptr, gscale_dev, dl, acc);
  return launched();
}

// ------------------------------------------------------------------------------------------------
// berHu pseudo-depth loss (loss/loss.py:5-15, called at train.py:494) and per-pixel normalised entropy
// (loss/loss.py:40-47, label_selection.py:449).
// ------------------------------------------------------------------------------------------------
namespace segsde {

__device__ __forceinline__ float berhu_absdiff(float x, float t, float m, int apply_log, float* sgn, float* dlog) {
  float dl = 1.f;
  if (apply_log) { dl = 1.f / (1.f + x); x = logf(1.f + x); t = logf(1.f + t); }
  const float d = t - x;
  if (sgn) *sgn = d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f);
  if (dlog) *dlog = dl;
  return fabsf(d) * m;
}

// pass 1: max |t - x| * m  (non-negative floats order like their bit patterns -> atomicMax on the bits)
__global__ void __launch_bounds__(256) berhu_max_kernel(const float* __restrict__ x, const float* __restrict__ t,
                                                        const float* __restrict__ m, long long n, int apply_log,
                                                        unsigned int* __restrict__ maxbits) {
  float best = 0.f;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    best = fmaxf(best, berhu_absdiff(x[i], t[i], m ? m[i] : 1.f, apply_log, nullptr, nullptr));
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) best = fmaxf(best, __shfl_xor_sync(0xffffffffu, best, o));
  if ((threadIdx.x & 31) == 0 && best > 0.f) atomicMax(maxbits, __float_as_uint(best));
}

// pass 2: sum of (a <= C ? a : (a^2 + C^2) / (2C)), C = thr * max — the threshold is read from device memory,
// the reference's .item() round trip (loss.py:11) is not needed
__global__ void __launch_bounds__(256) berhu_sum_kernel(const float* __restrict__ x, const float* __restrict__ t,
                                                        const float* __restrict__ m, long long n, int apply_log,
                                                        const unsigned int* __restrict__ maxbits, float thr,
                                                        double* __restrict__ sum) {
  const float C = thr * __uint_as_float(*maxbits);
  double acc = 0.0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float a = berhu_absdiff(x[i], t[i], m ? m[i] : 1.f, apply_log, nullptr, nullptr);
    acc += (double)(a <= C ? a : (a * a + C * C) / (2.f * C));
  }
  acc = warp_sum_d(acc);
  __shared__ double red[8];
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0.0;
    for (int i = 0; i < 8; ++i) s += red[i];
    atomicAdd(sum, s);
  }
}
__global__ void berhu_finalize_kernel(const double* __restrict__ sum, double inv_n, float* __restrict__ loss) {
  loss[0] = (float)(sum[0] * inv_n);
}
// d loss / d x = gloss / n * (a <= C ? 1 : a / C) * d a / d x,  d a / d x = -sign(t' - x') * m * (1 / (1 + x) with log)
__global__ void __launch_bounds__(256) berhu_bwd_kernel(const float* __restrict__ x, const float* __restrict__ t,
                                                        const float* __restrict__ m, long long n, int apply_log,
                                                        const unsigned int* __restrict__ maxbits, float thr,
                                                        const float* __restrict__ gloss, float inv_n,
                                                        float* __restrict__ dx) {
  const float C = thr * __uint_as_float(*maxbits);
  const float g = gloss[0] * inv_n;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    float sgn, dl;
    const float mi = m ? m[i] : 1.f;
    const float a = berhu_absdiff(x[i], t[i], mi, apply_log, &sgn, &dl);
    const float w = a <= C ? 1.f : a / C;
    dx[i] = -g * w * sgn * mi * dl;
  }
}

// entropy[n,h,w] = -sum_c p log2(p + 1e-30) / log2(C), p = softmax over the channel axis of NCHW logits
__global__ void __launch_bounds__(256) entropy_kernel(const float* __restrict__ logits, int N, int C, long long HW,
                                                      float* __restrict__ out, unsigned int* __restrict__ minmax) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  float e = 0.f;
  const bool valid = i < (long long)N * HW;
  if (valid) {
    const long long n = i / HW, p = i % HW;
    const float* l = logits + n * C * HW + p;
    float mx = -3.4e38f;
    for (int c = 0; c < C; ++c) mx = fmaxf(mx, l[c * HW]);
    float den = 0.f;
    for (int c = 0; c < C; ++c) den += expf(l[c * HW] - mx);
    const float inv = 1.f / den;
    float s = 0.f;
    for (int c = 0; c < C; ++c) { const float pr = expf(l[c * HW] - mx) * inv; s += pr * log2f(pr + 1e-30f); }
    e = -s / log2f((float)C);
    out[i] = e;
  }
  if (minmax) {       // entropy >= 0 up to rounding: clamp for the bit-pattern ordering of the atomics
    float lo = valid ? fmaxf(e, 0.f) : 3.4e38f, hi = valid ? fmaxf(e, 0.f) : 0.f;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      lo = fminf(lo, __shfl_xor_sync(0xffffffffu, lo, o));
      hi = fmaxf(hi, __shfl_xor_sync(0xffffffffu, hi, o));
    }
    if ((threadIdx.x & 31) == 0) { atomicMin(minmax, __float_as_uint(lo)); atomicMax(minmax + 1, __float_as_uint(hi)); }
  }
}
__global__ void entropy_normalize_kernel(float* __restrict__ e, long long n, const unsigned int* __restrict__ minmax) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float lo = __uint_as_float(minmax[0]), hi = __uint_as_float(minmax[1]);
  e[i] = (e[i] - lo) / (hi - lo);
}

}  // namespace segsde

extern "C" int segsde_berhu_fwd(const float* input, const float* target, const float* mask, int64_t n, int apply_log,
                                float threshold, unsigned int* maxbits, double* sum, float* loss, void* stream) {
  if (!input || !target || !maxbits || !sum || !loss || n < 1) return SEGSDE_E_ARG;
  cudaStream_t st = as_stream(stream);
  long long blocks = cdiv(n, 256 * 8); if (blocks > 148 * 8) blocks = 148 * 8; if (blocks < 1) blocks = 1;
  berhu_max_kernel<<<(unsigned)blocks, 256, 0, st>>>(input, target, mask, n, apply_log, maxbits);
  berhu_sum_kernel<<<(unsigned)blocks, 256, 0, st>>>(input, target, mask, n, apply_log, maxbits, threshold, sum);
  berhu_finalize_kernel<<<1, 1, 0, st>>>(sum, 1.0 / (double)n, loss);
  return launched();
}
extern "C" int segsde_berhu_bwd(const float* input, const float* target, const float* mask, int64_t n, int apply_log,
                                float threshold, const unsigned int* maxbits, const float* gloss, float* dinput,
                                void* stream) {
  if (!input || !target || !maxbits || !gloss || !dinput || n < 1) return SEGSDE_E_ARG;
  long long blocks = cdiv(n, 256 * 8); if (blocks > 148 * 8) blocks = 148 * 8; if (blocks < 1) blocks = 1;
  berhu_bwd_kernel<<<(unsigned)blocks, 256, 0, as_stream(stream)>>>(input, target, mask, n, apply_log, maxbits, threshold,
                                                                    gloss, (float)(1.0 / (double)n), dinput);
  return launched();
}
extern "C" int segsde_pixel_entropy(const float* logits, int n, int c, int h, int w, int normalize, float* entropy,
                                    unsigned int* minmax, void* stream) {
  if (!logits || !entropy || n < 1 || c < 2 || h < 1 || w < 1 || (normalize && !minmax)) return SEGSDE_E_ARG;
  const long long hw = (long long)h * w, total = (long long)n * hw;
  cudaStream_t st = as_stream(stream);
  entropy_kernel<<<cdiv(total, 256), 256, 0, st>>>(logits, n, c, hw, entropy, normalize ? minmax : nullptr);
  if (normalize) entropy_normalize_kernel<<<cdiv(total, 256), 256, 0, st>>>(entropy, total, minmax);
  return launched();
}
