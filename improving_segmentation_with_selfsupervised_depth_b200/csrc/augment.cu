// On-device input-pipeline pieces, sm_100a, NCHW planar fp32 like the loader's tensors:
//   * gaussian_blur  — loader/transformsgpu.py:21-30: kornia.filters.GaussianBlur2d((ky,kx), (sigma,sigma)), reflect border;
//     51 x 103 taps at 512x1024 -> separable: one horizontal and one vertical pass through shared-memory tiles
//     (2 x (read + write) of the image instead of 5253 taps per pixel);
//   * colour jitter  — loader/transformsgpu.py:10-18: the four kornia.augmentation.ColorJitter primitives (brightness,
//     contrast, saturation, hue) as one fused per-pixel kernel with the factors and the application order as arguments;
//   * area pyramid   — the loader's per-scale resize (sequence_segmentation_loader.py:96-99,308-309; SURVEY §8d uses
//     F.interpolate(mode="area")): exact 2^s x 2^s box averages for all scales in one launch.
// kornia is absent from this image: the blur / jitter arithmetic is restated from kornia 0.4's documented
// definitions and pinned against a PyTorch restatement in tests/test_augment.py ("parity unpinned" w.r.t. kornia itself).
#include "common.cuh"

namespace segsde {

constexpr int GB_MAXK = 255;

// horizontal pass: one block per (row = blockIdx.x, 256-pixel segment = blockIdx.y)
__global__ void __launch_bounds__(256) blur_h_kernel(const float* __restrict__ x, float* __restrict__ y, int W, int K,
                                                     const float* __restrict__ taps) {
  extern __shared__ float sm[];
  float* row = sm;            // [256 + K - 1]
  float* tw = sm + 256 + K - 1;
  const int r = K / 2;
  const long long base = (long long)blockIdx.x * W;
  const int x0 = blockIdx.y * 256;
  for (int i = threadIdx.x; i < 256 + K - 1; i += 256) {
    int gx = x0 - r + i;
    gx = reflect_idx(gx < -(W - 1) ? -(W - 1) : (gx > 2 * W - 2 ? 2 * W - 2 : gx), W);
    row[i] = x[base + gx];
  }
  for (int i = threadIdx.x; i < K; i += 256) tw[i] = taps[i];
  __syncthreads();
  const int ox = x0 + threadIdx.x;
  if (ox < W) {
    float acc = 0.f;
    for (int k = 0; k < K; ++k) acc = fmaf(tw[k], row[threadIdx.x + k], acc);
    y[base + ox] = acc;
  }
}
// vertical pass: one block per (plane, 32-column strip, 64-row segment)
__global__ void __launch_bounds__(256) blur_v_kernel(const float* __restrict__ x, float* __restrict__ y, int H, int W, int K,
                                                     const float* __restrict__ taps) {
  extern __shared__ float sm[];
  float* tile = sm;                        // [(64 + K - 1)][32]
  float* tw = sm + (64 + K - 1) * 32;
  const int r = K / 2;
  const long long base = (long long)blockIdx.z * H * W;
  const int x0 = blockIdx.x * 32, y0 = blockIdx.y * 64;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int i = ty; i < 64 + K - 1; i += 8) {
    int gy = y0 - r + i;
    gy = reflect_idx(gy < -(H - 1) ? -(H - 1) : (gy > 2 * H - 2 ? 2 * H - 2 : gy), H);
    tile[i * 32 + tx] = (x0 + tx < W) ? x[base + (long long)gy * W + x0 + tx] : 0.f;
  }
  for (int i = threadIdx.x; i < K; i += 256) tw[i] = taps[i];
  __syncthreads();
  for (int j = ty; j < 64; j += 8) {
    const int oy = y0 + j;
    if (oy < H && x0 + tx < W) {
      float acc = 0.f;
      for (int k = 0; k < K; ++k) acc = fmaf(tw[k], tile[(j + k) * 32 + tx], acc);
      y[base + (long long)oy * W + x0 + tx] = acc;
    }
  }
}

// ---- colour jitter ------------------------------------------------------------------------------------------------------
struct JitterP { float brightness, contrast, saturation, hue; int order[4]; };   // order: permutation of 0..3 = b, c, s, h
__device__ __forceinline__ void rgb2hsv(float r, float g, float b, float& h, float& s, float& v) {
  const float mx = fmaxf(r, fmaxf(g, b)), mn = fminf(r, fminf(g, b));
  v = mx;
  const float d = mx - mn;
  s = d / (mx + 1e-6f);
  const float dd = d == 0.f ? 1.f : d;
  float hh;
  if (mx == r) hh = (g - b) / dd;
  else if (mx == g) hh = 2.f + (b - r) / dd;
  else hh = 4.f + (r - g) / dd;
  hh = hh / 6.f;
  hh = hh - floorf(hh);
  h = hh * 6.283185307179586f;           // radians in [0, 2 pi)
}
__device__ __forceinline__ void hsv2rgb(float h, float s, float v, float& r, float& g, float& b) {
  const float hi = floorf(h / 6.283185307179586f * 6.f);
  const float f = h / 6.283185307179586f * 6.f - hi;
  const int i = ((int)hi) % 6;
  const float p = v * (1.f - s), q = v * (1.f - f * s), t = v * (1.f - (1.f - f) * s);
  switch (i) {
    case 0: r = v; g = t; b = p; break;
    case 1: r = q; g = v; b = p; break;
    case 2: r = p; g = v; b = t; break;
    case 3: r = p; g = q; b = v; break;
    case 4: r = t; g = p; b = v; break;
    default: r = v; g = p; b = q; break;
  }
}
__global__ void __launch_bounds__(256) jitter_kernel(const float* __restrict__ x, float* __restrict__ y, int B, long long hw,
                                                     JitterP jp) {
  const long long total = (long long)B * hw;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long b = i / hw, p = i - b * hw;
    const float* q = x + b * 3 * hw + p;
    float r = q[0], g = q[hw], bl = q[2 * hw];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int op = jp.order[k];
      if (op == 0) {            // brightness: additive shift, clamped to [0, 1]
        r = fminf(fmaxf(r + jp.brightness, 0.f), 1.f); g = fminf(fmaxf(g + jp.brightness, 0.f), 1.f);
        bl = fminf(fmaxf(bl + jp.brightness, 0.f), 1.f);
      } else if (op == 1) {     // contrast: multiplicative, clamped
        r = fminf(fmaxf(r * jp.contrast, 0.f), 1.f); g = fminf(fmaxf(g * jp.contrast, 0.f), 1.f);
        bl = fminf(fmaxf(bl * jp.contrast, 0.f), 1.f);
      } else if (op == 2) {     // saturation: scale S in HSV, clamped
        float h, s, v; rgb2hsv(r, g, bl, h, s, v);
        s = fminf(fmaxf(s * jp.saturation, 0.f), 1.f);
        hsv2rgb(h, s, v, r, g, bl);
      } else {                  // hue: shift H by hue (radians), wrapped
        float h, s, v; rgb2hsv(r, g, bl, h, s, v);
        h = fmodf(h + jp.hue, 6.283185307179586f);
        if (h < 0.f) h += 6.283185307179586f;
        hsv2rgb(h, s, v, r, g, bl);
      }
    }
    float* o = y + b * 3 * hw + p;
    o[0] = r; o[hw] = g; o[2 * hw] = bl;
  }
}

// ---- area pyramid: scale s output pixel = mean of the 2^s x 2^s block of the full-resolution image -----------------------------
__global__ void __launch_bounds__(256) pyramid_kernel(const float* __restrict__ x, int NC, int H, int W, float* __restrict__ y1,
                                                      float* __restrict__ y2, float* __restrict__ y3) {
  // one thread per scale-3 pixel (8x8 block): it also produces the 4 scale-2 and 16 scale-1 pixels inside it
  const int h3 = H >> 3, w3 = W >> 3;
  const long long total = (long long)NC * h3 * w3;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int ox = (int)(i % w3), oy = (int)((i / w3) % h3); const long long nc = i / ((long long)w3 * h3);
    const float* p = x + nc * (long long)H * W + (long long)(oy * 8) * W + ox * 8;
    float s3 = 0.f;
#pragma unroll
    for (int by = 0; by < 2; ++by)
#pragma unroll
      for (int bx = 0; bx < 2; ++bx) {
        float s2 = 0.f;
#pragma unroll
        for (int cy = 0; cy < 2; ++cy)
#pragma unroll
          for (int cx = 0; cx < 2; ++cx) {
            const float* q = p + (long long)(by * 4 + cy * 2) * W + bx * 4 + cx * 2;
            const float s1 = (q[0] + q[1]) + (q[W] + q[W + 1]);
            if (y1) y1[nc * (long long)(H >> 1) * (W >> 1) + (long long)(oy * 4 + by * 2 + cy) * (W >> 1) + ox * 4 + bx * 2 + cx] = s1 * 0.25f;
            s2 += s1;
          }
        if (y2) y2[nc * (long long)(H >> 2) * (W >> 2) + (long long)(oy * 2 + by) * (W >> 2) + ox * 2 + bx] = s2 * (1.f / 16.f);
        s3 += s2;
      }
    if (y3) y3[i] = s3 * (1.f / 64.f);
  }
}

}  // namespace segsde
using namespace segsde;

extern "C" int segsde_gaussian_blur(const float* x, float* tmp, float* y, int planes, int h, int w, int ky, int kx,
                                    const float* taps_y, const float* taps_x, void* stream) {
  if (!x || !tmp || !y || !taps_y || !taps_x || planes < 1 || h < 2 || w < 2) return SEGSDE_E_ARG;
  if (ky < 1 || kx < 1 || !(ky & 1) || !(kx & 1) || ky > GB_MAXK || kx > GB_MAXK || ky / 2 >= h || kx / 2 >= w) return SEGSDE_E_ARG;
  cudaStream_t st = as_stream(stream);
  dim3 gh(planes * h, cdiv(w, 256));
  blur_h_kernel<<<gh, 256, sizeof(float) * (256 + 2 * kx), st>>>(x, tmp, w, kx, taps_x);
  const size_t smv = sizeof(float) * ((size_t)(64 + ky - 1) * 32 + ky);
  static bool attr = false;
  if (!attr) { cudaFuncSetAttribute(blur_v_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(sizeof(float) * ((64 + GB_MAXK) * 32 + GB_MAXK))); attr = true; }
  dim3 gv(cdiv(w, 32), cdiv(h, 64), planes);
  blur_v_kernel<<<gv, 256, smv, st>>>(tmp, y, h, w, ky, taps_y);
  return launched();
}
extern "C" int segsde_color_jitter(const float* x, float* y, int b, int64_t hw, float brightness, float contrast, float saturation,
                                   float hue, const int* order4, void* stream) {
  if (!x || !y || !order4 || b < 1 || hw < 1) return SEGSDE_E_ARG;
  JitterP jp;
  jp.brightness = brightness; jp.contrast = contrast; jp.saturation = saturation; jp.hue = hue;
  int seen = 0;
  for (int k = 0; k < 4; ++k) { if (order4[k] < 0 || order4[k] > 3) return SEGSDE_E_ARG; jp.order[k] = order4[k]; seen |= 1 << order4[k]; }
  if (seen != 15) return SEGSDE_E_ARG;
  long long blocks = cdiv((long long)b * hw, 256 * 4);
  if (blocks > 148 * 8) blocks = 148 * 8;
  jitter_kernel<<<(unsigned)blocks, 256, 0, as_stream(stream)>>>(x, y, b, hw, jp);
  return launched();
}
extern "C" int segsde_area_pyramid(const float* x, int planes, int h, int w, float* y1, float* y2, float* y3, void* stream) {
  if (!x || planes < 1 || h < 8 || w < 8 || (h & 7) || (w & 7)) return SEGSDE_E_ARG;
  long long blocks = cdiv((long long)planes * (h >> 3) * (w >> 3), 256);
  if (blocks > 148 * 16) blocks = 148 * 16;
  pyramid_kernel<<<(unsigned)blocks, 256, 0, as_stream(stream)>>>(x, planes, h, w, y1, y2, y3);
  return launched();
}
