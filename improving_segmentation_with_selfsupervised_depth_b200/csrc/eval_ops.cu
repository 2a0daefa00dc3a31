// Validation / inference path and label-selection scoring ops (rank 4), sm_100a.
//   * confusion matrix on the device: per-pixel arg-max of the logits + histogram n*gt + pred, replacing
//     `semantics.data.max(1)[1].cpu().numpy()` + numpy bincount (train.py:846-850, evaluation/metrics.py:12-25);
//   * BatchNorm folding for eval-mode inference: w' = w * gamma / sqrt(var + eps), b' = beta - mean * gamma / sqrt(var + eps)
//     (+ conv bias), so conv -> BN -> ReLU runs as ONE convolution with a bias + ReLU epilogue;
//   * adaptive average / max pooling, pairwise p-norm distances (torch.cdist) and the iterative-farthest-point loop of
//     label_selection.py:347-650, the latter as ONE launch instead of 2 reductions + a host round trip per new sample.
// All HBM-bound or latency-bound streaming kernels.
#include "common.cuh"

namespace segsde {

// ---- confusion matrix ---------------------------------------------------------------------------------------------
// logits addressed by (sample, channel, pixel) element strides (NCHW planar or channels-last); hist: [n][n] int64,
// hist[gt][pred] += 1 for 0 <= gt < n.  pred = first maximum (np.argmax / torch.max semantics).  Block-private
// shared-memory histogram (n <= 32), one global atomic per non-empty bin per block.
__global__ void __launch_bounds__(256) confusion_kernel(const float* __restrict__ logits, const long long* __restrict__ pred_in,
                                                        const long long* __restrict__ gt, int B, int C, long long hw,
                                                        long long sn, long long sc, long long sp, int n,
                                                        unsigned long long* __restrict__ hist) {
  __shared__ unsigned int sh[32 * 32];
  for (int i = threadIdx.x; i < n * n; i += blockDim.x) sh[i] = 0;
  __syncthreads();
  const long long total = (long long)B * hw;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long t = gt[i];
    if (t < 0 || t >= n) continue;
    long long p;
    if (pred_in) p = pred_in[i];
    else {
      const int b = (int)(i / hw); const long long px = i - (long long)b * hw;
      const float* q = logits + b * sn + px * sp;
      float best = q[0]; int bi = 0;
      for (int c = 1; c < C; ++c) { const float v = q[c * sc]; if (v > best) { best = v; bi = c; } }
      p = bi;
    }
    if (p >= 0 && p < n) atomicAdd(&sh[(int)t * n + (int)p], 1u);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < n * n; i += blockDim.x)
    if (sh[i]) atomicAdd(hist + i, (unsigned long long)sh[i]);
}

// ---- BatchNorm folding ------------------------------------------------------------------------------------------------
// w: [O][K] (K = kh*kw*Cin, OHWI), one block per output channel
__global__ void __launch_bounds__(128) bn_fold_kernel(const float* __restrict__ w, const float* __restrict__ conv_bias,
                                                      const float* __restrict__ gamma, const float* __restrict__ beta,
                                                      const float* __restrict__ mean, const float* __restrict__ var, float eps,
                                                      int K, float* __restrict__ w_out, float* __restrict__ b_out) {
  const int o = blockIdx.x;
  const float s = (gamma ? gamma[o] : 1.f) / sqrtf(var[o] + eps);
  for (int k = threadIdx.x; k < K; k += blockDim.x) w_out[(long long)o * K + k] = w[(long long)o * K + k] * s;
  if (threadIdx.x == 0) b_out[o] = (beta ? beta[o] : 0.f) + ((conv_bias ? conv_bias[o] : 0.f) - mean[o]) * s;
}

// ---- adaptive pooling (NCHW planar) -------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) adaptive_pool_kernel(const float* __restrict__ x, int NC, int H, int W, int oh, int ow,
                                                            int is_max, float* __restrict__ y) {
  const long long total = (long long)NC * oh * ow;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int ox = (int)(i % ow), oy = (int)((i / ow) % oh); const long long nc = i / ((long long)ow * oh);
    // ATen: start = floor(o * in / out), end = ceil((o + 1) * in / out)
    const int y0 = (int)(((long long)oy * H) / oh), y1 = (int)((((long long)oy + 1) * H + oh - 1) / oh);
    const int x0 = (int)(((long long)ox * W) / ow), x1 = (int)((((long long)ox + 1) * W + ow - 1) / ow);
    const float* p = x + nc * (long long)H * W;
    float acc = is_max ? -3.4e38f : 0.f;
    for (int yy = y0; yy < y1; ++yy)
      for (int xx = x0; xx < x1; ++xx) {
        const float v = p[(long long)yy * W + xx];
        acc = is_max ? fmaxf(acc, v) : acc + v;
      }
    y[i] = is_max ? acc : acc / (float)((y1 - y0) * (x1 - x0));
  }
}

// ---- pairwise distances (torch.cdist of a set with itself) -----------------------------------------------------------------
// f: [N][D]; out[i][j] = (sum_d |f_i - f_j|^p)^(1/p).  16x16 pairs per block, D streamed through shared memory.
constexpr int PD_T = 16, PD_K = 32;
__global__ void __launch_bounds__(PD_T * PD_T) pairwise_kernel(const float* __restrict__ f, int N, long long D, float p,
                                                               float* __restrict__ out) {
  __shared__ float sa[PD_T][PD_K + 1], sb[PD_T][PD_K + 1];
  const int tx = threadIdx.x % PD_T, ty = threadIdx.x / PD_T;
  const int i = blockIdx.y * PD_T + ty, j = blockIdx.x * PD_T + tx;
  float acc = 0.f;
  for (long long d0 = 0; d0 < D; d0 += PD_K) {
    for (int e = threadIdx.x; e < PD_T * PD_K; e += PD_T * PD_T) {
      const int r = e / PD_K, c = e % PD_K;
      const int gi = blockIdx.y * PD_T + r, gj = blockIdx.x * PD_T + r;
      sa[r][c] = (gi < N && d0 + c < D) ? f[(long long)gi * D + d0 + c] : 0.f;
      sb[r][c] = (gj < N && d0 + c < D) ? f[(long long)gj * D + d0 + c] : 0.f;
    }
    __syncthreads();
#pragma unroll 8
    for (int c = 0; c < PD_K; ++c) {
      const float d = fabsf(sa[ty][c] - sb[tx][c]);
      acc += (p == 2.f) ? d * d : (p == 1.f ? d : powf(d, p));
    }
    __syncthreads();
  }
  if (i < N && j < N) out[(long long)i * N + j] = (p == 2.f) ? sqrtf(acc) : (p == 1.f ? acc : powf(acc, 1.f / p));
}

// ---- iterative farthest point (label_selection.py:617-640) ----------------------------------------------------------------
// dist: [N][N] (columns of ignored samples already zeroed by the caller); is_current: [N] 0/1 (updated in place);
// new_idx / new_dist: [n_new]; count: number of samples actually added (the loop stops when the farthest sample is
// already selected).  One CTA; every iteration = min-update of the running "distance to the selected set" + arg-max.
__global__ void __launch_bounds__(1024) farthest_point_kernel(const float* __restrict__ dist, int N, int* __restrict__ is_current,
                                                              int n_new, long long* __restrict__ new_idx,
                                                              float* __restrict__ new_dist, int* __restrict__ count,
                                                              float* __restrict__ mind) {
  __shared__ float sv[32]; __shared__ int si[32]; __shared__ int s_pick;
  const int tid = threadIdx.x;
  // distance of every sample to the current set
  for (int j = tid; j < N; j += blockDim.x) {
    float m = 3.4e38f;
    for (int i = 0; i < N; ++i) if (is_current[i]) m = fminf(m, dist[(long long)i * N + j]);
    mind[j] = m;
  }
  __syncthreads();
  int added = 0;
  for (int it = 0; it < n_new; ++it) {
    float bv = -3.4e38f; int bi = N;
    for (int j = tid; j < N; j += blockDim.x) { const float v = mind[j]; if (v > bv || (v == bv && j < bi)) { bv = v; bi = j; } }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, bv, o); const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
    }
    if ((tid & 31) == 0) { sv[tid >> 5] = bv; si[tid >> 5] = bi; }
    __syncthreads();
    if (tid == 0) {
      float v = sv[0]; int idx = si[0];
      for (int w = 1; w < (int)(blockDim.x >> 5); ++w) if (sv[w] > v || (sv[w] == v && si[w] < idx)) { v = sv[w]; idx = si[w]; }
      if (idx >= N || is_current[idx]) s_pick = -1;
      else { s_pick = idx; is_current[idx] = 1; new_idx[it] = idx; new_dist[it] = v; }
    }
    __syncthreads();
    const int pick = s_pick;
    if (pick < 0) break;
    ++added;
    for (int j = tid; j < N; j += blockDim.x) mind[j] = fminf(mind[j], dist[(long long)pick * N + j]);
    __syncthreads();
  }
  if (tid == 0) count[0] = added;
}

static unsigned grid1d(long long n) {
  long long b = cdiv(n, 256 * 4);
  if (b > 148 * 8) b = 148 * 8;
  if (b < 1) b = 1;
  return (unsigned)b;
}

}  // namespace segsde
using namespace segsde;

extern "C" int segsde_confusion_update(const float* logits, const int64_t* pred, const int64_t* gt, int b, int c, int64_t hw,
                                       int64_t sn, int64_t sc, int64_t sp, int n_classes, int64_t* hist, void* stream) {
  if ((!logits && !pred) || !gt || !hist || b < 1 || hw < 1 || n_classes < 1 || n_classes > 32) return SEGSDE_E_ARG;
  if (logits && c < 1) return SEGSDE_E_ARG;
  confusion_kernel<<<grid1d((long long)b * hw), 256, 0, as_stream(stream)>>>(
      logits, (const long long*)pred, (const long long*)gt, b, c, hw, sn, sc, sp, n_classes, (unsigned long long*)hist);
  return launched();
}
extern "C" int segsde_bn_fold(const float* w, const float* conv_bias, const float* gamma, const float* beta, const float* mean,
                              const float* var, float eps, int cout, int k, float* w_out, float* b_out, void* stream) {
  if (!w || !mean || !var || !w_out || !b_out || cout < 1 || k < 1) return SEGSDE_E_ARG;
  bn_fold_kernel<<<cout, 128, 0, as_stream(stream)>>>(w, conv_bias, gamma, beta, mean, var, eps, k, w_out, b_out);
  return launched();
}
extern "C" int segsde_adaptive_pool(const float* x, int nc, int h, int w, int oh, int ow, int is_max, float* y, void* stream) {
  if (!x || !y || nc < 1 || h < 1 || w < 1 || oh < 1 || ow < 1) return SEGSDE_E_ARG;
  adaptive_pool_kernel<<<grid1d((long long)nc * oh * ow), 256, 0, as_stream(stream)>>>(x, nc, h, w, oh, ow, is_max, y);
  return launched();
}
extern "C" int segsde_pairwise_distance(const float* f, int n, int64_t d, float p, float* out, void* stream) {
  if (!f || !out || n < 1 || d < 1 || !(p > 0.f)) return SEGSDE_E_ARG;
  dim3 grid(cdiv(n, PD_T), cdiv(n, PD_T));
  pairwise_kernel<<<grid, PD_T * PD_T, 0, as_stream(stream)>>>(f, n, d, p, out);
  return launched();
}
extern "C" int segsde_farthest_point(const float* dist, int n, int* is_current, int n_new, int64_t* new_idx, float* new_dist,
                                     int* count, float* scratch, void* stream) {
  if (!dist || !is_current || !new_idx || !new_dist || !count || !scratch || n < 1 || n_new < 0) return SEGSDE_E_ARG;
  farthest_point_kernel<<<1, 1024, 0, as_stream(stream)>>>(dist, n, is_current, n_new, (long long*)new_idx, new_dist, count,
                                                          scratch);
  return launched();
}
