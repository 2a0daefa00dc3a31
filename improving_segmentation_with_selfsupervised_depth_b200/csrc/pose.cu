// Pose head geometry: (axis-angle, translation) -> 4x4 camera transform and its Jacobian.
// Restates transformation_from_parameters / rot_from_axisangle / get_translation_matrix
// (models/monodepth_layers.py:30-105).  The 12x6 Jacobian of M[:3,:4] is produced in the forward
// pass with forward-mode dual numbers, so the backward entry is a 12x6 mat-vec per sample.
#include "common.cuh"

namespace segsde {

struct D6 {
  float v; float d[6];
};
__device__ __forceinline__ D6 dconst(float c) { D6 r; r.v = c; for (int i = 0; i < 6; ++i) r.d[i] = 0.f; return r; }
__device__ __forceinline__ D6 dvar(float c, int k) { D6 r = dconst(c); r.d[k] = 1.f; return r; }
__device__ __forceinline__ D6 operator+(const D6& a, const D6& b) { D6 r; r.v = a.v + b.v; for (int i = 0; i < 6; ++i) r.d[i] = a.d[i] + b.d[i]; return r; }
__device__ __forceinline__ D6 operator-(const D6& a, const D6& b) { D6 r; r.v = a.v - b.v; for (int i = 0; i < 6; ++i) r.d[i] = a.d[i] - b.d[i]; return r; }
__device__ __forceinline__ D6 operator-(const D6& a) { D6 r; r.v = -a.v; for (int i = 0; i < 6; ++i) r.d[i] = -a.d[i]; return r; }
__device__ __forceinline__ D6 operator*(const D6& a, const D6& b) { D6 r; r.v = a.v * b.v; for (int i = 0; i < 6; ++i) r.d[i] = a.d[i] * b.v + a.v * b.d[i]; return r; }
__device__ __forceinline__ D6 operator/(const D6& a, const D6& b) {
  D6 r; r.v = a.v / b.v; const float ib = 1.f / b.v;
  for (int i = 0; i < 6; ++i) r.d[i] = (a.d[i] - r.v * b.d[i]) * ib;
  return r;
}
__device__ __forceinline__ D6 dsqrt(const D6& a) {
  D6 r; r.v = sqrtf(a.v); const float s = r.v > 0.f ? 0.5f / r.v : 0.f;   // torch.norm: zero subgradient at 0
  for (int i = 0; i < 6; ++i) r.d[i] = a.d[i] * s;
  return r;
}
__device__ __forceinline__ D6 dsin(const D6& a) { D6 r; r.v = sinf(a.v); const float c = cosf(a.v); for (int i = 0; i < 6; ++i) r.d[i] = a.d[i] * c; return r; }
__device__ __forceinline__ D6 dcos(const D6& a) { D6 r; r.v = cosf(a.v); const float s = -sinf(a.v); for (int i = 0; i < 6; ++i) r.d[i] = a.d[i] * s; return r; }

__global__ void pose_matrix_kernel(const float* __restrict__ vec, int B, int invert, float* __restrict__ M,
                                   float* __restrict__ jac) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  D6 a[3], t[3];
  for (int i = 0; i < 3; ++i) { a[i] = dvar(vec[b * 6 + i], i); t[i] = dvar(vec[b * 6 + 3 + i], 3 + i); }
  const D6 angle = dsqrt(a[0] * a[0] + a[1] * a[1] + a[2] * a[2]);
  const D6 den = angle + dconst(1e-7f);
  const D6 x = a[0] / den, y = a[1] / den, z = a[2] / den;
  const D6 ca = dcos(angle), sa = dsin(angle);
  const D6 C = dconst(1.f) - ca;
  const D6 xs = x * sa, ys = y * sa, zs = z * sa;
  const D6 xC = x * C, yC = y * C, zC = z * C;
  const D6 xyC = x * yC, yzC = y * zC, zxC = z * xC;
  D6 R[3][3];
  R[0][0] = x * xC + ca; R[0][1] = xyC - zs;   R[0][2] = zxC + ys;
  R[1][0] = xyC + zs;   R[1][1] = y * yC + ca; R[1][2] = yzC - xs;
  R[2][0] = zxC - ys;   R[2][1] = yzC + xs;   R[2][2] = z * zC + ca;
  D6 out[3][4];
  if (!invert) {                    // M = T * R  ->  [R | t]
    for (int i = 0; i < 3; ++i) { for (int j = 0; j < 3; ++j) out[i][j] = R[i][j]; out[i][3] = t[i]; }
  } else {                          // M = R^T * T(-t)  ->  [R^T | -R^T t]
    for (int i = 0; i < 3; ++i) {
      for (int j = 0; j < 3; ++j) out[i][j] = R[j][i];
      out[i][3] = -(R[0][i] * t[0] + R[1][i] * t[1] + R[2][i] * t[2]);
    }
  }
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 4; ++j) {
      M[b * 16 + i * 4 + j] = out[i][j].v;
      if (jac) for (int k = 0; k < 6; ++k) jac[(b * 12 + i * 4 + j) * 6 + k] = out[i][j].d[k];
    }
  M[b * 16 + 12] = 0.f; M[b * 16 + 13] = 0.f; M[b * 16 + 14] = 0.f; M[b * 16 + 15] = 1.f;
}

__global__ void pose_matrix_bwd_kernel(const float* __restrict__ jac, const float* __restrict__ dM, int B,
                                       float* __restrict__ dvec) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * 6) return;
  const int b = i / 6, k = i % 6;
  float a = 0.f;
  for (int e = 0; e < 12; ++e) a += jac[(b * 12 + e) * 6 + k] * dM[b * 16 + e];
  dvec[i] = a;
}

}  // namespace segsde
using namespace segsde;

extern "C" int segsde_pose_matrix_fwd(const float* vec, int b, int invert, float* M, float* jac, void* stream) {
  if (!vec || !M || b < 1) return SEGSDE_E_ARG;
  pose_matrix_kernel<<<cdiv(b, 64), 64, 0, as_stream(stream)>>>(vec, b, invert, M, jac);
  return launched();
}
extern "C" int segsde_pose_matrix_bwd(const float* jac, const float* dM, int b, float* dvec, void* stream) {
  if (!jac || !dM || !dvec || b < 1) return SEGSDE_E_ARG;
  pose_matrix_bwd_kernel<<<cdiv(b * 6, 64), 64, 0, as_stream(stream)>>>(jac, dM, b, dvec);
  return launched();
}
