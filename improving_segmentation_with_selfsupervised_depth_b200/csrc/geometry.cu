// Standalone forms of the reference's geometry / image layers (models/monodepth_layers.py:145-254).
// The training path uses the fused kernel in reproj.cu; these exist so that the layer API
// (BackprojectDepth, Project3D, SSIM, upsample) stays callable with the reference's semantics.
#include "common.cuh"

namespace segsde {

__global__ void backproject_kernel(const float* __restrict__ depth, const float* __restrict__ invK, int H, int W,
                                   float* __restrict__ out) {
  const int b = blockIdx.y;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int HW = H * W;
  if (i >= HW) return;
  const float x = (float)(i % W), y = (float)(i / W);
  const float* ik = invK + b * 16;
  const float d = depth[(size_t)b * HW + i];
  float* o = out + (size_t)b * 4 * HW;
  o[i] = d * (ik[0] * x + ik[1] * y + ik[2]);
  o[HW + i] = d * (ik[4] * x + ik[5] * y + ik[6]);
  o[2 * HW + i] = d * (ik[8] * x + ik[9] * y + ik[10]);
  o[3 * HW + i] = 1.f;
}

__global__ void project3d_kernel(const float* __restrict__ pts, const float* __restrict__ K,
                                 const float* __restrict__ T, int H, int W, float eps, float* __restrict__ out) {
  const int b = blockIdx.y;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int HW = H * W;
  __shared__ float P[12];
  if (threadIdx.x < 12) {
    const int r = threadIdx.x / 4, c = threadIdx.x % 4;
    float a = 0.f;
    for (int q = 0; q < 4; ++q) a += K[b * 16 + r * 4 + q] * T[b * 16 + q * 4 + c];
    P[threadIdx.x] = a;
  }
  __syncthreads();
  if (i >= HW) return;
  const float* p = pts + (size_t)b * 4 * HW;
  const float x = p[i], y = p[HW + i], z = p[2 * HW + i], w = p[3 * HW + i];
  const float X = P[0] * x + P[1] * y + P[2] * z + P[3] * w;
  const float Y = P[4] * x + P[5] * y + P[6] * z + P[7] * w;
  const float Z = P[8] * x + P[9] * y + P[10] * z + P[11] * w;
  const float zz = Z + eps;
  out[((size_t)b * HW + i) * 2 + 0] = (X / zz / (float)(W - 1) - 0.5f) * 2.f;
  out[((size_t)b * HW + i) * 2 + 1] = (Y / zz / (float)(H - 1) - 0.5f) * 2.f;
}

__global__ void ssim_map_kernel(const float* __restrict__ x, const float* __restrict__ y, int H, int W,
                                float* __restrict__ out) {
  const int pl = blockIdx.z;
  const int px = blockIdx.x * blockDim.x + threadIdx.x, py = blockIdx.y * blockDim.y + threadIdx.y;
  if (px >= W || py >= H) return;
  const float* xp = x + (size_t)pl * H * W;
  const float* yp = y + (size_t)pl * H * W;
  float sx = 0.f, sy = 0.f, sxx = 0.f, syy = 0.f, sxy = 0.f;
  for (int dy = -1; dy <= 1; ++dy)
    for (int dx = -1; dx <= 1; ++dx) {
      const int o = reflect_idx(py + dy, H) * W + reflect_idx(px + dx, W);
      const float a = xp[o], b = yp[o];
      sx += a; sy += b; sxx += a * a; syy += b * b; sxy += a * b;
    }
  const float mu_x = sx / 9.f, mu_y = sy / 9.f;
  const float sig_x = sxx / 9.f - mu_x * mu_x, sig_y = syy / 9.f - mu_y * mu_y, sig_xy = sxy / 9.f - mu_x * mu_y;
  const float n = (2.f * mu_x * mu_y + 1e-4f) * (2.f * sig_xy + 9e-4f);
  const float d = (mu_x * mu_x + mu_y * mu_y + 1e-4f) * (sig_x + sig_y + 9e-4f);
  out[(size_t)pl * H * W + (size_t)py * W + px] = fminf(fmaxf((1.f - n / d) * 0.5f, 0.f), 1.f);
}

__global__ void upsample2x_kernel(View x, View y) {
  const long long total = (long long)y.n * y.h * y.w * y.c;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int c = (int)(idx % y.c); long long q = idx / y.c;
  const int w = (int)(q % y.w); q /= y.w;
  const int h = (int)(q % y.h); const int n = (int)(q / y.h);
  y.p[y.off(n, h, w) + c] = x.p[x.off(n, h >> 1, w >> 1) + c];
}

}  // namespace segsde
using namespace segsde;

extern "C" int segsde_backproject(const float* depth, const float* inv_K, int B, int H, int W, float* out,
                                  void* stream) {
  if (!depth || !inv_K || !out || B < 1) return SEGSDE_E_ARG;
  dim3 grid(cdiv((int64_t)H * W, 256), B);
  backproject_kernel<<<grid, 256, 0, as_stream(stream)>>>(depth, inv_K, H, W, out);
  return launched();
}
extern "C" int segsde_project3d(const float* points, const float* K, const float* T, int B, int H, int W,
                                float eps, float* out, void* stream) {
  if (!points || !K || !T || !out || B < 1) return SEGSDE_E_ARG;
  dim3 grid(cdiv((int64_t)H * W, 256), B);
  project3d_kernel<<<grid, 256, 0, as_stream(stream)>>>(points, K, T, H, W, eps, out);
  return launched();
}
extern "C" int segsde_ssim_map(const float* x, const float* y, int planes, int H, int W, float* out, void* stream) {
  if (!x || !y || !out || planes < 1 || H < 2 || W < 2) return SEGSDE_E_ARG;
  dim3 block(32, 8), grid(cdiv(W, 32), cdiv(H, 8), planes);
  ssim_map_kernel<<<grid, block, 0, as_stream(stream)>>>(x, y, H, W, out);
  return launched();
}
extern "C" int segsde_upsample2x_nearest(const segsde_nhwc_t* x, const segsde_nhwc_t* y, void* stream) {
  if (!x || !y || !x->ptr || !y->ptr) return SEGSDE_E_ARG;
  View vx = mk(x), vy = mk(y);
  if (vy.h != 2 * vx.h || vy.w != 2 * vx.w || vy.c != vx.c || vy.n != vx.n) return SEGSDE_E_ARG;
  const long long total = (long long)vy.n * vy.h * vy.w * vy.c;
  upsample2x_kernel<<<cdiv(total, 256), 256, 0, as_stream(stream)>>>(vx, vy);
  return launched();
}
