// Fused monodepth photometric reprojection loss for one scale (sm_100a).
//
// Replaces, per scale, the op chain of the reference (all fp32):
//   F.interpolate(disp, bilinear, align_corners=False)           loss/monodepth_loss.py:71-73
//   disp_to_depth                                                models/monodepth_layers.py:18-27
//   BackprojectDepth / Project3D                                 models/monodepth_layers.py:169-199
//   F.grid_sample(border, align_corners=True)                    loss/monodepth_loss.py:94-98
//   SSIM (3x3 mean, reflection pad) + L1, 0.85/0.15 mix          monodepth_layers.py:240-254,
//                                                                monodepth_loss.py:104-116
//   identity auto-mask + tie-break noise + per-pixel min + mean  monodepth_loss.py:140-179
// and, when gdisp != NULL, produces in the same pass the gradient of that mean w.r.t. the
// scale's low-resolution disparity map and (as per-block partials) w.r.t. P = (K T)[:3,:].
//
// Work decomposition: one CTA per 32x8 pixel tile of one sample.  The warped images of both
// source frames are built in shared memory on the tile plus a halo (1 px for the loss, 2 px
// when gradients are requested, because d loss[p] / d pred[q] couples 3x3 neighbourhoods twice),
// so every colour value is fetched from global memory once per tile and the ~15 full-size
// temporaries of the reference never exist.
#include "common.cuh"

namespace segsde {

constexpr int TX = 32, TY = 8, NTHREADS = TX * TY;

struct ReprojK {
  const float* tgt; const float* src[2]; const float* disp;
  const float* K; const float* invK; const float* T[2];
  const float* noise; unsigned long long seed, offset;
  int B, H, W, hs, ws, F;
  float min_disp, max_disp;
  int flags;
  float* loss_partial; float* ident_sel; float* gdisp; float* gT_partial;
  float inv_count;
  int tiles_x, tiles_y;
  float* ident_cache; int ident_mode;
};

struct UpW {  // bilinear source taps of the low-res disparity for one full-res pixel
  int i00, i01, i10, i11;
  float w00, w01, w10, w11;
};

// F.interpolate(mode="bilinear", align_corners=False) with an explicit output size:
// scale = in/out, src = max(scale*(dst+0.5)-0.5, 0)   (ATen area_pixel_compute_source_index)
__device__ __forceinline__ float disp_up(const float* __restrict__ d, int hs, int ws, int H, int W,
                                         int x, int y, UpW* tap) {
  if (hs == H && ws == W) {
    int i = y * ws + x;
    if (tap) { tap->i00 = tap->i01 = tap->i10 = tap->i11 = i; tap->w00 = 1.f; tap->w01 = tap->w10 = tap->w11 = 0.f; }
    return __ldg(d + i);
  }
  const float sy = (float)hs / (float)H, sx = (float)ws / (float)W;
  float fy = sy * ((float)y + 0.5f) - 0.5f; fy = fy < 0.f ? 0.f : fy;
  float fx = sx * ((float)x + 0.5f) - 0.5f; fx = fx < 0.f ? 0.f : fx;
  int y0 = (int)fy, x0 = (int)fx;
  int y1 = y0 + (y0 < hs - 1 ? 1 : 0), x1 = x0 + (x0 < ws - 1 ? 1 : 0);
  float ly = fy - (float)y0, lx = fx - (float)x0;
  float hy = 1.f - ly, hx = 1.f - lx;
  float v00 = __ldg(d + y0 * ws + x0), v01 = __ldg(d + y0 * ws + x1);
  float v10 = __ldg(d + y1 * ws + x0), v11 = __ldg(d + y1 * ws + x1);
  if (tap) {
    tap->i00 = y0 * ws + x0; tap->i01 = y0 * ws + x1; tap->i10 = y1 * ws + x0; tap->i11 = y1 * ws + x1;
    tap->w00 = hy * hx; tap->w01 = hy * lx; tap->w10 = ly * hx; tap->w11 = ly * lx;
  }
  return hy * (hx * v00 + lx * v01) + ly * (hx * v10 + lx * v11);
}

struct Warp {   // everything the gradient chain needs about one (pixel, frame) sample
  float pred[3];
  float dpx[3], dpy[3];   // d pred_c / d ix, d pred_c / d iy, already multiplied by the clip mask
  float px, py, zinv;     // projected pixel coordinates and 1/(Z+eps)
};

// Projects the back-projected point of pixel (x,y) with depth `depth` into source frame `P`,
// and samples the three colour planes bilinearly (border padding, align_corners=True).
template <bool GRAD>
__device__ __forceinline__ void warp_pixel(const float* __restrict__ src, int H, int W,
                                           const float* P, float cx, float cy, float cz, Warp& o) {
  const float X = P[0] * cx + P[1] * cy + P[2] * cz + P[3];
  const float Y = P[4] * cx + P[5] * cy + P[6] * cz + P[7];
  const float Z = P[8] * cx + P[9] * cy + P[10] * cz + P[11];
  const float z = Z + 1e-7f;
  float px = X / z, py = Y / z;
  // Project3D's normalisation ((p/(W-1) - 0.5)*2) and grid_sample's un-normalisation (((g+1)/2)*(W-1)) are
  // exact inverses; composing them only adds a few ulp of rounding and four divisions per sample, so the
  // pixel coordinate is used directly.
  float ix = px, iy = py;
  float mx = 1.f, my = 1.f;
  const float maxx = (float)(W - 1), maxy = (float)(H - 1);
  if (!(ix > 0.f)) { ix = 0.f; mx = 0.f; } else if (ix >= maxx) { ix = maxx; mx = 0.f; }
  if (!(iy > 0.f)) { iy = 0.f; my = 0.f; } else if (iy >= maxy) { iy = maxy; my = 0.f; }
  const float fx0 = floorf(ix), fy0 = floorf(iy);
  const int x0 = (int)fx0, y0 = (int)fy0;
  const int x1 = min(x0 + 1, W - 1), y1 = min(y0 + 1, H - 1);
  const float ax = ix - fx0, ay = iy - fy0;          // ix - ix_nw
  const float bx = (fx0 + 1.f) - ix, by = (fy0 + 1.f) - iy;  // ix_se - ix
  const float nw = bx * by, ne = ax * by, sw = bx * ay, se = ax * ay;
  const size_t plane = (size_t)H * W;
  const float* r0 = src + (size_t)y0 * W;
  const float* r1 = src + (size_t)y1 * W;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const float vnw = __ldg(r0 + c * plane + x0), vne = __ldg(r0 + c * plane + x1);
    const float vsw = __ldg(r1 + c * plane + x0), vse = __ldg(r1 + c * plane + x1);
    o.pred[c] = vnw * nw + vne * ne + vsw * sw + vse * se;
    if (GRAD) {
      o.dpx[c] = mx * ((vne - vnw) * by + (vse - vsw) * ay);
      o.dpy[c] = my * ((vsw - vnw) * bx + (vse - vne) * ax);
    }
  }
  if (GRAD) { o.px = px; o.py = py; o.zinv = 1.f / z; }
}

struct Stats { float mu_x, mu_y, sxx, syy, sxy; };

// SSIM of monodepth_layers.py:240-254 for one channel at one centre; also returns the statistics.
__device__ __forceinline__ float ssim_from_sums(float sx, float sy, float sxx, float syy, float sxy,
                                                Stats* st) {
  constexpr float i9 = 1.f / 9.f;      // AvgPool2d(3,1): multiply instead of five fp32 divisions per window
  const float mu_x = sx * i9, mu_y = sy * i9;
  const float sig_x = sxx * i9 - mu_x * mu_x;
  const float sig_y = syy * i9 - mu_y * mu_y;
  const float sig_xy = sxy * i9 - mu_x * mu_y;
  const float n = (2.f * mu_x * mu_y + 1e-4f) * (2.f * sig_xy + 9e-4f);
  const float d = (mu_x * mu_x + mu_y * mu_y + 1e-4f) * (sig_x + sig_y + 9e-4f);
  if (st) { st->mu_x = mu_x; st->mu_y = mu_y; st->sxx = sig_x; st->syy = sig_y; st->sxy = sig_xy; }
  return (1.f - n / d) * 0.5f;
}

template <bool GRAD>
__global__ void __launch_bounds__(NTHREADS, GRAD ? 3 : 4) reproj_kernel(ReprojK k) {
  constexpr int R = GRAD ? 2 : 1;        // halo of the warped images
  constexpr int RC = R - 1;              // halo of the loss centres
  constexpr int RW = TX + 2 * R, RH = TY + 2 * R, NP = RW * RH;
  constexpr int CW = TX + 2 * RC, CH = TY + 2 * RC, NC = CW * CH;

  extern __shared__ float smem[];
  float* s_tgt = smem;                       // [3][NP]
  float* s_src = s_tgt + 3 * NP;             // [2][3][NP]
  float* s_pred = s_src + 6 * NP;            // [2][3][NP]
  float* s_wgt = s_pred + 6 * NP;            // [2][NC]   (GRAD) selection weight per frame
  float* s_coef = s_wgt + 2 * NC;            // [2][3][3][NC] (GRAD) affine SSIM-gradient coefficients per (frame, channel)
  float* s_red = s_coef + (GRAD ? 18 : 0) * NC;   // [8][25]
  __shared__ float sP[2][12];
  __shared__ float sIK[9];

  const int tid = threadIdx.x;
  const int tx = tid % TX, ty = tid / TX;
  const int b = blockIdx.z;
  const int x0 = blockIdx.x * TX, y0 = blockIdx.y * TY;
  const int H = k.H, W = k.W, F = k.F;
  const size_t plane = (size_t)H * W;
  const bool no_ssim = k.flags & SEGSDE_REPROJ_NO_SSIM;
  const bool avg = k.flags & SEGSDE_REPROJ_AVG;
  const bool automask = !(k.flags & SEGSDE_REPROJ_NO_AUTOMASK);

  if (tid < 12 * F) {
    const int f = tid / 12, e = tid % 12, i = e / 4, j = e % 4;
    const float* Km = k.K + b * 16;
    const float* Tm = k.T[f] + b * 16;
    float acc = 0.f;
#pragma unroll
    for (int q = 0; q < 4; ++q) acc += Km[i * 4 + q] * Tm[q * 4 + j];
    sP[f][e] = acc;
  }
  if (tid >= 32 && tid < 41) {
    const int e = tid - 32;
    sIK[e] = k.invK[b * 16 + (e / 3) * 4 + (e % 3)];
  }

  // ---- stage 1: target + raw source tiles (identity candidates) -> smem ------------------------
  const int rx0 = x0 - R, ry0 = y0 - R;
  for (int i = tid; i < NP; i += NTHREADS) {
    const int gx = rx0 + i % RW, gy = ry0 + i / RW;
    const bool in = gx >= 0 && gx < W && gy >= 0 && gy < H;
    const size_t o = (size_t)gy * W + gx;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      s_tgt[c * NP + i] = in ? __ldg(k.tgt + ((size_t)b * 3 + c) * plane + o) : 0.f;
      if (automask && k.ident_mode != 2)
        for (int f = 0; f < F; ++f)
          s_src[(f * 3 + c) * NP + i] = in ? __ldg(k.src[f] + ((size_t)b * 3 + c) * plane + o) : 0.f;
    }
  }
  __syncthreads();

  // ---- stage 2: warp both source frames on the tile + halo ---------------------------------------
  const float* dispb = k.disp + (size_t)b * k.hs * k.ws;
  const float dscale = k.max_disp - k.min_disp;
  const int qx = x0 + tx, qy = y0 + ty;
  const bool own_valid = qx < W && qy < H;
  Warp own[2];
  UpW tap;
  float ray[3] = {0.f, 0.f, 0.f}, depth_own = 0.f;
  if (own_valid) {
    const float d = disp_up(dispb, k.hs, k.ws, H, W, qx, qy, GRAD ? &tap : nullptr);
    const float depth = 1.f / (k.min_disp + dscale * d);
    const float fx = (float)qx, fy = (float)qy;
    ray[0] = sIK[0] * fx + sIK[1] * fy + sIK[2];
    ray[1] = sIK[3] * fx + sIK[4] * fy + sIK[5];
    ray[2] = sIK[6] * fx + sIK[7] * fy + sIK[8];
    depth_own = depth;
    const int si = (qy - ry0) * RW + (qx - rx0);
#pragma unroll
    for (int f = 0; f < 2; ++f) {
      if (f >= F) break;
      warp_pixel<GRAD>(k.src[f] + (size_t)b * 3 * plane, H, W, sP[f], depth * ray[0], depth * ray[1],
                       depth * ray[2], own[f]);
#pragma unroll
      for (int c = 0; c < 3; ++c) s_pred[(f * 3 + c) * NP + si] = own[f].pred[c];
    }
  }
  {
    // halo ring, compactly enumerated: R top rows, R bottom rows, 2R side columns of TY rows
    constexpr int NHALO = NP - TX * TY;
    for (int j = tid; j < NHALO; j += NTHREADS) {
      int lx, ly;
      if (j < R * RW) { ly = j / RW; lx = j % RW; }
      else if (j < 2 * R * RW) { int q = j - R * RW; ly = R + TY + q / RW; lx = q % RW; }
      else { int q = j - 2 * R * RW; ly = R + q / (2 * R); int cc = q % (2 * R); lx = cc < R ? cc : TX + cc; }
      const int gx = rx0 + lx, gy = ry0 + ly;
      if (gx < 0 || gx >= W || gy < 0 || gy >= H) continue;
      const float d = disp_up(dispb, k.hs, k.ws, H, W, gx, gy, nullptr);
      const float depth = 1.f / (k.min_disp + dscale * d);
      const float fx = (float)gx, fy = (float)gy;
      const float r0 = sIK[0] * fx + sIK[1] * fy + sIK[2];
      const float r1 = sIK[3] * fx + sIK[4] * fy + sIK[5];
      const float r2 = sIK[6] * fx + sIK[7] * fy + sIK[8];
      Warp h;
      for (int f = 0; f < F; ++f) {
        warp_pixel<false>(k.src[f] + (size_t)b * 3 * plane, H, W, sP[f], depth * r0, depth * r1,
                          depth * r2, h);
#pragma unroll
        for (int c = 0; c < 3; ++c) s_pred[(f * 3 + c) * NP + ly * RW + lx] = h.pred[c];
      }
    }
  }
  __syncthreads();

  // ---- stage 3: photometric candidates + min at every centre ------------------------------------
  // centre list: the thread's own pixel, then (GRAD only) the 1-px ring around the tile
  float my_loss = 0.f;
  constexpr int NCH = NC - TX * TY;   // ring centres (0 when !GRAD)
  const int n_iter = 1 + (GRAD ? (NCH + NTHREADS - 1) / NTHREADS : 0);
  for (int it = 0; it < n_iter; ++it) {
    int cx, cy;     // image coordinates of the centre
    bool is_own = (it == 0);
    if (is_own) { cx = qx; cy = qy; if (!own_valid) continue; }
    else {
      const int j = tid + (it - 1) * NTHREADS;
      if (j >= NCH) continue;
      int lx, ly;
      if (j < CW) { ly = 0; lx = j; }
      else if (j < 2 * CW) { ly = CH - 1; lx = j - CW; }
      else { int q = j - 2 * CW; ly = 1 + q / 2; lx = (q & 1) ? CW - 1 : 0; }
      cx = x0 - RC + lx; cy = y0 - RC + ly;
      if (cx < 0 || cx >= W || cy < 0 || cy >= H) continue;
    }
    // window offsets in smem (reflection at the IMAGE border, ReflectionPad2d(1))
    int wo[9];
#pragma unroll
    for (int dy = -1; dy <= 1; ++dy)
#pragma unroll
      for (int dx = -1; dx <= 1; ++dx)
        wo[(dy + 1) * 3 + dx + 1] = (reflect_idx(cy + dy, H) - ry0) * RW + (reflect_idx(cx + dx, W) - rx0);
    const int ctr = wo[4];
    const int ci = (cy - (y0 - RC)) * CW + (cx - (x0 - RC));     // this centre's slot in the [NC] arrays
    float cand[4] = {0.f, 0.f, 0.f, 0.f};   // [identity f0, identity f1, reproj f0, reproj f1]
    // images: m = 0,1 reprojected frames; m = 2,3 raw source frames (identity candidates).  The identity
    // candidates do not depend on the scale: scale 0 stores them (ident_mode 1), coarser scales read them back
    // (ident_mode 2) instead of recomputing two SSIM windows per pixel.
    const bool id_compute = automask && k.ident_mode != 2;
    float ssim_acc[4] = {0.f, 0.f, 0.f, 0.f}, l1_acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      float yw[9], sy = 0.f, syy = 0.f;
#pragma unroll
      for (int t = 0; t < 9; ++t) { yw[t] = s_tgt[c * NP + wo[t]]; sy += yw[t]; syy = fmaf(yw[t], yw[t], syy); }
#pragma unroll
      for (int m = 0; m < 4; ++m) {
        const int f = m & 1;
        if (f >= F || (m >= 2 && !id_compute)) continue;
        const float* X = (m < 2) ? (s_pred + (f * 3 + c) * NP) : (s_src + (f * 3 + c) * NP);
        l1_acc[m] += fabsf(yw[4] - X[ctr]);
        if (!no_ssim) {
          float sx = 0.f, sxx = 0.f, sxy = 0.f;
#pragma unroll
          for (int t = 0; t < 9; ++t) { const float xv = X[wo[t]]; sx += xv; sxx = fmaf(xv, xv, sxx); sxy = fmaf(xv, yw[t], sxy); }
          Stats st;
          const float v = ssim_from_sums(sx, sy, sxx, syy, sxy, (GRAD && m < 2) ? &st : nullptr);
          ssim_acc[m] += fminf(fmaxf(v, 0.f), 1.f);
          if (GRAD && m < 2) {
            // d(0.85/3 * clamp(ssim)) / d x_i = A + B x_i + C y_i for every x_i of this window (the window
            // statistics are already here; stage 4 only gathers).  Zero where the clamp is active.
            float A = 0.f, Bc = 0.f, Cc = 0.f;
            if (v >= 0.f && v <= 1.f) {
              const float n1 = 2.f * st.mu_x * st.mu_y + 1e-4f, n2 = 2.f * st.sxy + 9e-4f;
              const float d1 = st.mu_x * st.mu_x + st.mu_y * st.mu_y + 1e-4f, d2 = st.sxx + st.syy + 9e-4f;
              const float Nn = n1 * n2;
              const float iD = 1.f / (d1 * d2);
              const float sc = (0.85f / 3.f) * (1.f / 9.f);
              A = -sc * (st.mu_y * (n2 - n1) * iD - Nn * st.mu_x * (d2 - d1) * iD * iD);
              Bc = sc * Nn * d1 * iD * iD;
              Cc = -sc * n1 * iD;
            }
            float* cf = s_coef + (size_t)(f * 3 + c) * 3 * NC + ci;
            cf[0] = A; cf[NC] = Bc; cf[2 * NC] = Cc;
          }
        }
      }
    }
#pragma unroll
    for (int m = 0; m < 4; ++m) {
      const int f = m & 1;
      if (f >= F) continue;
      const float l1 = l1_acc[m] * (1.f / 3.f);
      const float v = no_ssim ? l1 : 0.85f * (ssim_acc[m] * (1.f / 3.f)) + 0.15f * l1;
      if (m < 2) cand[2 + f] = v;
      else if (id_compute) {
        cand[f] = v;
        if (k.ident_mode == 1 && is_own) k.ident_cache[(((size_t)b * F + f) * H + cy) * W + cx] = v;
      } else if (automask) {
        cand[f] = __ldg(k.ident_cache + (((size_t)b * F + f) * H + cy) * W + cx);
      }
    }
    // candidates in the reference's order: identity (+noise) first, then reprojection
    float best = 3.4e38f; int best_i = -1; int n_id = 0;
    float rp0 = cand[2], rp1 = cand[3], id0 = cand[0], id1 = cand[1];
    if (avg && F == 2) { rp0 = (rp0 + rp1) * 0.5f; id0 = (id0 + id1) * 0.5f; }
    const int ncand_f = (avg ? 1 : F);
    if (automask) {
      n_id = ncand_f;
      float nz[2];
      if (k.noise) {
        for (int f = 0; f < ncand_f; ++f)
          nz[f] = __ldg(k.noise + (((size_t)b * ncand_f + f) * H + cy) * W + cx);
      } else {
        Philox ph(k.seed, ((unsigned long long)b * H + cy) * W + cx, k.offset);
        ph.run();
        const float rad = sqrtf(-2.f * logf(u01(ph.c[0])));
        float sn, cs;
        sincospif(2.f * u01(ph.c[1]), &sn, &cs);
        nz[0] = rad * cs * 1e-5f; nz[1] = rad * sn * 1e-5f;
      }
      const float v0 = id0 + nz[0];
      best = v0; best_i = 0;
      if (ncand_f == 2) { const float v1 = id1 + nz[1]; if (v1 < best) { best = v1; best_i = 1; } }
    }
    if (rp0 < best) { best = rp0; best_i = n_id; }
    if (ncand_f == 2 && rp1 < best) { best = rp1; best_i = n_id + 1; }
    const bool reproj_won = best_i >= n_id;
    if (is_own) {
      my_loss = best;
      if (k.ident_sel) k.ident_sel[(size_t)b * plane + (size_t)cy * W + cx] = reproj_won ? 1.f : 0.f;
    }
    if (GRAD) {
      float w0 = 0.f, w1 = 0.f;
      if (reproj_won) {
        if (avg) { w0 = w1 = (F == 2 ? 0.5f : 1.f); }
        else if (best_i - n_id == 0) w0 = 1.f; else w1 = 1.f;
      }
      s_wgt[ci] = w0; s_wgt[NC + ci] = w1;
    }
  }

  // block-reduce the loss (deterministic: fixed tree, one partial per block)
  {
    float v = warp_sum(my_loss);
    if ((tid & 31) == 0) s_red[tid >> 5] = v;
    __syncthreads();
    if (tid == 0) {
      float t = 0.f;
      for (int i = 0; i < NTHREADS / 32; ++i) t += s_red[i];
      k.loss_partial[((size_t)b * k.tiles_y + blockIdx.y) * k.tiles_x + blockIdx.x] = t;
    }
  }
  if (!GRAD) return;

  // ---- stage 4 (GRAD): d mean-loss / d pred: gather the affine coefficients of the 3x3 centres around q -----
  // ring centres outside the image never wrote s_wgt / s_coef: they are never read either (their slot is replaced
  // by the thread's own with multiplicity 0)
  __syncthreads();
  float gpred[2][3] = {{0.f, 0.f, 0.f}, {0.f, 0.f, 0.f}};
  const float wl1 = no_ssim ? 1.f / 3.f : 0.15f / 3.f;
  if (own_valid) {
    const int si = (qy - ry0) * RW + (qx - rx0);
    const int cself = (qy - (y0 - RC)) * CW + (qx - (x0 - RC));
    int cidx[9]; float mult[9];
#pragma unroll
    for (int dy = -1; dy <= 1; ++dy) {
      const int py = qy + dy;
      // multiplicity of q in centre p's reflected window (adjoint of ReflectionPad2d(1))
      const float my = (py < 0 || py >= H) ? 0.f
                       : 1.f + ((py == 0 && qy == 1) ? 1.f : 0.f) + ((py == H - 1 && qy == H - 2) ? 1.f : 0.f);
#pragma unroll
      for (int dx = -1; dx <= 1; ++dx) {
        const int px = qx + dx;
        const float mx = (px < 0 || px >= W) ? 0.f
                         : 1.f + ((px == 0 && qx == 1) ? 1.f : 0.f) + ((px == W - 1 && qx == W - 2) ? 1.f : 0.f);
        const float m = mx * my;
        mult[(dy + 1) * 3 + dx + 1] = m;
        cidx[(dy + 1) * 3 + dx + 1] = m != 0.f ? cself + dy * CW + dx : cself;
      }
    }
#pragma unroll
    for (int f = 0; f < 2; ++f) {
      if (f >= F) break;
      float mw[9];
#pragma unroll
      for (int t = 0; t < 9; ++t) mw[t] = mult[t] * s_wgt[f * NC + cidx[t]];
      const float wself = s_wgt[f * NC + cself];
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const float xq = s_pred[(f * 3 + c) * NP + si], yq = s_tgt[c * NP + si];
        float g = 0.f;
        if (!no_ssim) {
          const float* cf = s_coef + (size_t)(f * 3 + c) * 3 * NC;
          float sa = 0.f, sb = 0.f, sc = 0.f;
#pragma unroll
          for (int t = 0; t < 9; ++t) {
            sa = fmaf(mw[t], cf[cidx[t]], sa); sb = fmaf(mw[t], cf[NC + cidx[t]], sb); sc = fmaf(mw[t], cf[2 * NC + cidx[t]], sc);
          }
          g = sa + sb * xq + sc * yq;
        }
        const float diff = xq - yq;
        const float sgn = diff > 0.f ? 1.f : (diff < 0.f ? -1.f : 0.f);
        g += wself * wl1 * sgn;
        gpred[f][c] = g * k.inv_count;
      }
    }
  }

  // ---- stage 5 (GRAD): chain to sample coordinates, depth, disparity, P -------------------------
  float gP[2][12];
  float gdepth = 0.f;
#pragma unroll
  for (int f = 0; f < 2; ++f)
#pragma unroll
    for (int e = 0; e < 12; ++e) gP[f][e] = 0.f;
  if (own_valid) {
    const float cam[4] = {depth_own * ray[0], depth_own * ray[1], depth_own * ray[2], 1.f};
#pragma unroll
    for (int f = 0; f < 2; ++f) {
      if (f >= F) break;
      float gix = 0.f, giy = 0.f;
#pragma unroll
      for (int c = 0; c < 3; ++c) { gix += gpred[f][c] * own[f].dpx[c]; giy += gpred[f][c] * own[f].dpy[c]; }
      // d ix / d px = 1 (the (W-1)/2 of grid_sample cancels Project3D's 2/(W-1))
      const float gX = gix * own[f].zinv, gY = giy * own[f].zinv;
      const float gZ = -(gix * own[f].px + giy * own[f].py) * own[f].zinv;
      const float gv[3] = {gX, gY, gZ};
#pragma unroll
      for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) gP[f][i * 4 + j] = gv[i] * cam[j];
      // d L / d cam = P[:, :3]^T gv ; d cam / d depth = ray
      const float* P = sP[f];
      const float gcx = P[0] * gX + P[4] * gY + P[8] * gZ;
      const float gcy = P[1] * gX + P[5] * gY + P[9] * gZ;
      const float gcz = P[2] * gX + P[6] * gY + P[10] * gZ;
      gdepth += gcx * ray[0] + gcy * ray[1] + gcz * ray[2];
    }
    // depth = 1/(min_disp + dscale*d)  ->  d depth / d d = -dscale * depth^2
    const float gd = -gdepth * dscale * depth_own * depth_own;
    float* g = k.gdisp + (size_t)b * k.hs * k.ws;
    if (k.hs == H && k.ws == W) {
      g[tap.i00] += gd;     // exactly one writer per element
    } else {
      atomicAdd(g + tap.i00, gd * tap.w00); atomicAdd(g + tap.i01, gd * tap.w01);
      atomicAdd(g + tap.i10, gd * tap.w10); atomicAdd(g + tap.i11, gd * tap.w11);
    }
  }
  // block-reduce the 12 d/dP entries per frame
#pragma unroll
  for (int f = 0; f < 2; ++f) {
    if (f >= F) break;
#pragma unroll
    for (int e = 0; e < 12; ++e) {
      const float v = warp_sum(gP[f][e]);
      if ((tid & 31) == 0) s_red[(tid >> 5) * 25 + e] = v;
    }
    __syncthreads();
    if (tid < 12) {
      float t = 0.f;
      for (int i = 0; i < NTHREADS / 32; ++i) t += s_red[i * 25 + tid];
      const size_t tile = (size_t)blockIdx.y * k.tiles_x + blockIdx.x;
      const size_t tiles = (size_t)k.tiles_x * k.tiles_y;
      k.gT_partial[(((size_t)f * k.B + b) * tiles + tile) * 12 + tid] = t;
    }
    __syncthreads();
  }
}

template <bool GRAD>
static size_t reproj_smem_bytes() {
  constexpr int R = GRAD ? 2 : 1, RC = R - 1;
  constexpr int NP = (TX + 2 * R) * (TY + 2 * R), NC = (TX + 2 * RC) * (TY + 2 * RC);
  return sizeof(float) * (size_t)(15 * NP + (GRAD ? 20 : 2) * NC + 8 * 25);
}

// loss = sum(partials) / count ; gT[f][b] = K[:3,:]^T (4x3) * gP (3x4)
__global__ void reproj_finalize_kernel(const float* __restrict__ part, int n, float inv_count,
                                       float* __restrict__ loss_out, const float* __restrict__ gTp,
                                       const float* __restrict__ K, int B, int tiles, int F,
                                       float* __restrict__ gT) {
  __shared__ double sh[256];
  __shared__ float sgp[12];
  const int tid = threadIdx.x;
  if (blockIdx.x == 0) {
    double a = 0.0;
    for (int i = tid; i < n; i += 256) a += (double)part[i];
    sh[tid] = a;
    __syncthreads();
    for (int s = 128; s > 0; s >>= 1) { if (tid < s) sh[tid] += sh[tid + s]; __syncthreads(); }
    if (tid == 0) loss_out[0] = (float)(sh[0] * (double)inv_count);
    return;
  }
  // blocks 1.. : one (f,b) pair each
  const int fb = blockIdx.x - 1;
  if (!gTp || fb >= F * B) return;
  const int b = fb % B;
  for (int e = 0; e < 12; ++e) {
    double a = 0.0;
    for (int i = tid; i < tiles; i += 256) a += (double)gTp[((size_t)fb * tiles + i) * 12 + e];
    sh[tid] = a;
    __syncthreads();
    for (int s = 128; s > 0; s >>= 1) { if (tid < s) sh[tid] += sh[tid + s]; __syncthreads(); }
    if (tid == 0) sgp[e] = (float)sh[0];
    __syncthreads();
  }
  if (tid < 16) {
    const int r = tid / 4, c = tid % 4;   // gT[r][c] = sum_i K[i][r] * gP[i][c], i<3
    const float* Km = K + b * 16;
    float a = 0.f;
    for (int i = 0; i < 3; ++i) a += Km[i * 4 + r] * sgp[i * 4 + c];
    gT[(size_t)fb * 16 + tid] = a;
  }
}

__global__ void reproj_materialize_kernel(const float* __restrict__ src, const float* __restrict__ disp,
                                          const float* __restrict__ K, const float* __restrict__ invK,
                                          const float* __restrict__ T, int B, int H, int W, int hs, int ws,
                                          float min_disp, float max_disp, float* __restrict__ depth_o,
                                          float* __restrict__ sample_o, float* __restrict__ color_o) {
  const int b = blockIdx.z;
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
  __shared__ float P[12];
  const int tid = threadIdx.y * blockDim.x + threadIdx.x;
  if (tid < 12 && T) {
    const int i = tid / 4, j = tid % 4;
    float a = 0.f;
    for (int q = 0; q < 4; ++q) a += K[b * 16 + i * 4 + q] * T[b * 16 + q * 4 + j];
    P[tid] = a;
  }
  __syncthreads();
  if (x >= W || y >= H) return;
  const size_t plane = (size_t)H * W, o = (size_t)y * W + x;
  const float d = disp_up(disp + (size_t)b * hs * ws, hs, ws, H, W, x, y, nullptr);
  const float depth = 1.f / (min_disp + (max_disp - min_disp) * d);
  if (depth_o) depth_o[b * plane + o] = depth;
  if (!T) return;
  const float* ik = invK + b * 16;
  const float fx = (float)x, fy = (float)y;
  const float r0 = ik[0] * fx + ik[1] * fy + ik[2], r1 = ik[4] * fx + ik[5] * fy + ik[6],
              r2 = ik[8] * fx + ik[9] * fy + ik[10];
  Warp w;
  if (src) warp_pixel<false>(src + (size_t)b * 3 * plane, H, W, P, depth * r0, depth * r1, depth * r2, w);
  if (sample_o) {
    const float cx = depth * r0, cy = depth * r1, cz = depth * r2;
    const float X = P[0] * cx + P[1] * cy + P[2] * cz + P[3];
    const float Y = P[4] * cx + P[5] * cy + P[6] * cz + P[7];
    const float Z = P[8] * cx + P[9] * cy + P[10] * cz + P[11];
    const float z = Z + 1e-7f;
    sample_o[(b * plane + o) * 2 + 0] = (X / z / (float)(W - 1) - 0.5f) * 2.f;
    sample_o[(b * plane + o) * 2 + 1] = (Y / z / (float)(H - 1) - 0.5f) * 2.f;
  }
  if (color_o && src)
    for (int c = 0; c < 3; ++c) color_o[((size_t)b * 3 + c) * plane + o] = w.pred[c];
}

}  // namespace segsde

using namespace segsde;

extern "C" int segsde_reproj_tiles(int H, int W) { return cdiv(W, TX) * cdiv(H, TY); }
extern "C" int segsde_reproj_num_partials(int B, int H, int W) { return B * segsde_reproj_tiles(H, W); }

extern "C" int segsde_reproj_fused(const segsde_reproj_args_t* a, void* stream) {
  if (!a || !a->tgt || !a->disp || !a->K || !a->inv_K || !a->loss_partial) return SEGSDE_E_ARG;
  if (a->F < 1 || a->F > 2 || a->B < 1 || a->H < 2 || a->W < 2 || a->hs < 1 || a->ws < 1) return SEGSDE_E_ARG;
  for (int f = 0; f < a->F; ++f)
    if (!a->src[f] || !a->T[f]) return SEGSDE_E_ARG;
  if (a->gdisp && !a->gT_partial) return SEGSDE_E_ARG;
  if (a->B > 65535) return SEGSDE_E_ARG;
  ReprojK k;
  k.tgt = a->tgt; k.src[0] = a->src[0]; k.src[1] = a->F > 1 ? a->src[1] : a->src[0];
  k.disp = a->disp; k.K = a->K; k.invK = a->inv_K;
  k.T[0] = a->T[0]; k.T[1] = a->F > 1 ? a->T[1] : a->T[0];
  k.noise = a->noise; k.seed = a->seed; k.offset = a->offset;
  k.B = a->B; k.H = a->H; k.W = a->W; k.hs = a->hs; k.ws = a->ws; k.F = a->F;
  k.min_disp = 1.f / a->max_depth; k.max_disp = 1.f / a->min_depth;
  k.flags = a->flags;
  k.loss_partial = a->loss_partial; k.ident_sel = a->ident_sel; k.gdisp = a->gdisp;
  k.gT_partial = a->gT_partial;
  k.inv_count = (float)(1.0 / ((double)a->B * a->H * a->W));
  k.tiles_x = cdiv(a->W, TX); k.tiles_y = cdiv(a->H, TY);
  k.ident_cache = a->ident_cache; k.ident_mode = a->ident_cache ? a->ident_mode : 0;
  if (k.ident_mode < 0 || k.ident_mode > 2) return SEGSDE_E_ARG;
  dim3 grid(k.tiles_x, k.tiles_y, a->B), block(NTHREADS);
  if (a->gdisp) {
    const size_t sm = reproj_smem_bytes<true>();
    static bool attr = false;
    if (!attr) { cudaFuncSetAttribute(reproj_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm); attr = true; }
    reproj_kernel<true><<<grid, block, sm, as_stream(stream)>>>(k);
  } else {
    const size_t sm = reproj_smem_bytes<false>();
    static bool attr = false;
    if (!attr) { cudaFuncSetAttribute(reproj_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm); attr = true; }
    reproj_kernel<false><<<grid, block, sm, as_stream(stream)>>>(k);
  }
  return launched();
}

extern "C" int segsde_reproj_finalize(const float* loss_partial, int n_partial, int64_t count,
                                      float* loss_out, const float* gT_partial, const float* K, int B,
                                      int tiles, int F, float* gT, void* stream) {
  if (!loss_partial || !loss_out || n_partial < 1 || count < 1) return SEGSDE_E_ARG;
  if (gT_partial && (!K || !gT)) return SEGSDE_E_ARG;
  const int blocks = 1 + (gT_partial ? F * B : 0);
  reproj_finalize_kernel<<<blocks, 256, 0, as_stream(stream)>>>(
      loss_partial, n_partial, (float)(1.0 / (double)count), loss_out, gT_partial, K, B, tiles, F, gT);
  return launched();
}

extern "C" int segsde_reproj_materialize(const float* src, const float* disp, const float* K,
                                         const float* inv_K, const float* T, int B, int H, int W, int hs,
                                         int ws, float min_depth, float max_depth, float* depth,
                                         float* sample, float* color, void* stream) {
  if (!disp || B < 1 || H < 2 || W < 2) return SEGSDE_E_ARG;
  if (T && (!K || !inv_K)) return SEGSDE_E_ARG;
  dim3 block(32, 8), grid(cdiv(W, 32), cdiv(H, 8), B);
  reproj_materialize_kernel<<<grid, block, 0, as_stream(stream)>>>(
      src, disp, K, inv_K, T, B, H, W, hs, ws, 1.f / max_depth, 1.f / min_depth, depth, sample, color);
  return launched();
}

extern "C" int segsde_disp_to_depth_up(const float* disp, int B, int hs, int ws, int H, int W,
                                       float min_depth, float max_depth, float* depth, void* stream) {
  return segsde_reproj_materialize(nullptr, disp, nullptr, nullptr, nullptr, B, H, W, hs, ws, min_depth,
                                   max_depth, depth, nullptr, nullptr, stream);
}
