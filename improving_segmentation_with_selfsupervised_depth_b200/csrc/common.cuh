// Shared device/host helpers for the segsde_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <atomic>
#include "../../include/segsde_b200.h"

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "segsde_b200 kernels are written for sm_100a (B200) only"
#endif

namespace segsde {

extern std::atomic<int64_t> g_launches;

inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }

// Call after every kernel launch: counts it and converts the launch status to the ABI code.
inline int launched() {
  g_launches.fetch_add(1, std::memory_order_relaxed);
  cudaError_t e = cudaPeekAtLastError();
  if (e != cudaSuccess) {
    cudaGetLastError();
    return (int)e;
  }
  return SEGSDE_OK;
}

inline int cdiv(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }

// Device-side copy of the ABI view (passed by value as a kernel parameter).
struct View {
  float* p;
  int n, h, w, c;
  long long sn, sh, sw;
  __host__ __device__ inline long long off(int in, int ih, int iw) const {
    return in * sn + ih * sh + iw * sw;
  }
};

inline View mk(const segsde_nhwc_t* t) {
  View v;
  if (!t) {
    v.p = nullptr; v.n = v.h = v.w = v.c = 0; v.sn = v.sh = v.sw = 0;
  } else {
    v.p = (float*)t->ptr; v.n = t->n; v.h = t->h; v.w = t->w; v.c = t->c;
    v.sn = t->sn; v.sh = t->sh; v.sw = t->sw;
  }
  return v;
}

inline bool same_shape(const View& a, const View& b) {
  return a.n == b.n && a.h == b.h && a.w == b.w && a.c == b.c;
}
// all rows 16-byte aligned so float4 access along C is legal
inline bool vec4_ok(const View& v) {
  return v.p && (v.c % 4 == 0) && (v.sn % 4 == 0) && (v.sh % 4 == 0) && (v.sw % 4 == 0) &&
         ((reinterpret_cast<uintptr_t>(v.p) & 15) == 0);
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__device__ __forceinline__ int reflect_idx(int i, int n) {
  // ReflectionPad semantics for |overhang| < n : -1 -> 1, n -> n-2
  if (i < 0) i = -i;
  if (i >= n) i = 2 * n - 2 - i;
  return i;
}

__device__ __forceinline__ float act_apply(float v, int act) {
  switch (act) {
    case SEGSDE_ACT_RELU: return v > 0.f ? v : 0.f;
    case SEGSDE_ACT_ELU: return v > 0.f ? v : expm1f(v);
    case SEGSDE_ACT_SIGMOID: return 1.f / (1.f + expf(-v));
    default: return v;
  }
}
// derivative expressed through the activation OUTPUT y
__device__ __forceinline__ float act_grad_from_out(float y, int act) {
  switch (act) {
    case SEGSDE_ACT_RELU: return y > 0.f ? 1.f : 0.f;
    case SEGSDE_ACT_ELU: return y > 0.f ? 1.f : (y + 1.f);
    case SEGSDE_ACT_SIGMOID: return y * (1.f - y);
    default: return 1.f;
  }
}

// Philox4x32-10 (counter-based; same stream regardless of launch geometry)
struct Philox {
  uint32_t c[4];
  uint32_t k[2];
  __device__ __forceinline__ Philox(uint64_t seed, uint64_t ctr_lo, uint64_t ctr_hi) {
    k[0] = (uint32_t)seed; k[1] = (uint32_t)(seed >> 32);
    c[0] = (uint32_t)ctr_lo; c[1] = (uint32_t)(ctr_lo >> 32);
    c[2] = (uint32_t)ctr_hi; c[3] = (uint32_t)(ctr_hi >> 32);
  }
  __device__ __forceinline__ void round() {
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u;
    uint32_t hi0 = __umulhi(M0, c[0]), lo0 = M0 * c[0];
    uint32_t hi1 = __umulhi(M1, c[2]), lo1 = M1 * c[2];
    uint32_t n0 = hi1 ^ c[1] ^ k[0], n1 = lo1, n2 = hi0 ^ c[3] ^ k[1], n3 = lo0;
    c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
    k[0] += 0x9E3779B9u; k[1] += 0xBB67AE85u;
  }
  __device__ __forceinline__ void run() {
#pragma unroll
    for (int i = 0; i < 10; ++i) round();
  }
};
__device__ __forceinline__ float u01(uint32_t x) {  // (0,1]
  return ((float)(x >> 8) + 1.0f) * (1.0f / 16777216.0f);
}

// single-output-channel fast paths (conv_small.cu); return SEGSDE_E_UNSUPPORTED when the shape is not theirs
int c1_fwd(const View& x, const View& y, const float* w, const float* bias, const segsde_conv_desc_t* d, cudaStream_t st);
int c1_dgrad(const View& dy, const float* w, const View& dx, const segsde_conv_desc_t* d, cudaStream_t st);
int c1_wgrad(const View& x, const View& dy, float* dw, const segsde_conv_desc_t* d, cudaStream_t st);

// bandwidth-tuned paths for pixel-contiguous tensors (bn_fast.cu)
bool pix_contig(const View& v);
bool fast_reduce_ok(const View& x);
int bn_stats_fast(const View& x, double* sums, cudaStream_t st);
int bn_bwd_reduce_fast(const View& x, const View& y, const View& dy, const float* mean, const float* invstd, int act,
                       double* red, cudaStream_t st, const float* gamma, const float* beta);
int act_bwd_bias_fast(const View& y, const View& dy, const View& dz, int act, float* dbias, cudaStream_t st);
int bn_apply_fast(const View& x, const View& res, const View& y, const float* mean, const float* invstd, const float* gamma,
                  const float* beta, int act, cudaStream_t st);
int bn_apply_train_fast(const View& x, const View& res, const View& y, const double* sums, long long count, float eps,
                        float momentum, const float* gamma, const float* beta, int act, float* mean, float* invstd,
                        float* rmean, float* rvar, cudaStream_t st);
int bn_bwd_apply_fast(const View& x, const View& y, const View& dy, const View& dx, const View& dres, const float* mean,
                      const float* invstd, const float* gamma, int relu, int training, const double* red, long long count,
                      cudaStream_t st, const float* beta);

}  // namespace segsde
