// Fused monodepth photometric reprojection loss, all scales in ONE launch (sm_100a).
//
// Replaces, per scale, the op chain of the reference (all fp32):
//   F.interpolate(disp, bilinear, align_corners=False)           loss/monodepth_loss.py:71-73
//   disp_to_depth                                                models/monodepth_layers.py:18-27
//   BackprojectDepth / Project3D                                 models/monodepth_layers.py:169-199
//   F.grid_sample(border, align_corners=True)                    loss/monodepth_loss.py:94-98
//   SSIM (3x3 mean, reflection pad) + L1, 0.85/0.15 mix          monodepth_layers.py:240-254,
//                                                                monodepth_loss.py:104-116
//   identity auto-mask + tie-break noise + per-pixel min + mean  monodepth_loss.py:140-179
// and, when gdisp != NULL, produces in the same pass the gradient of those means w.r.t. every scale's
// low-resolution disparity map and (as per-warp partials) w.r.t. P = (K T)[:3,:] of each source frame.
//
// Work decomposition ("column march"): ONE WARP owns a strip of 28 output columns (32 lanes = 28 + a 2-column halo on
// each side) and a band of rows, and marches down the band one image row per step.  Everything a 3x3 window needs
// lives in registers:
//   * horizontal neighbours come from the adjacent lanes (__shfl_up / __shfl_down; reflection padding at the image
//     border = taking the other neighbour), vertical neighbours from a three-row ring of per-row horizontal sums
//     (sum x, sum x^2, sum x*y per frame and channel), i.e. the 3x3 sums are separable and never touch shared memory;
//   * d loss / d pred is affine in the window statistics: d/dx_q = sum_p w_p (A_p + B_p x_q + C_p y_q) over the 3x3
//     centres p around q, so the backward pass box-sums the selection-weighted coefficient fields the same way
//     (lanes + a three-row ring) two rows behind the forward front;
//   * the identity (auto-mask) candidates do not depend on the scale: a first sweep over the band computes them once
//     into a per-warp shared-memory stash, the four scale sweeps read them back (the round-1 kernel cached them in
//     HBM: 50 MB written and re-read three times per step);
//   * d/dP (12 values per frame) accumulates in registers over the whole band and is warp-reduced once per sweep.
// There is no __syncthreads anywhere, no halo-ring enumeration, no per-centre window re-summing; the halo recompute
// is (32/28) x (R+4)/R of the pixel count (1.29x at R = 32 rows) instead of 1.69x for the round-1 32x8 tile.
// Shared memory per warp: P / K^-1, the identity stash and a three-row ring of per-pixel "late" data (the bilinear
// slopes of the warped sample, needed again two rows later by the gradient chain).
#include "common.cuh"

namespace segsde {

constexpr int STRIP = 28;            // output columns per warp (lanes 2..29)
constexpr unsigned FULL = 0xffffffffu;

struct ReprojM {
  const float* tgt; const float* src[2];
  const float* K; const float* invK; const float* T[2];
  const float* disp[SEGSDE_REPROJ_MAX_SCALES]; int hs[SEGSDE_REPROJ_MAX_SCALES], ws[SEGSDE_REPROJ_MAX_SCALES];
  const float* noise[SEGSDE_REPROJ_MAX_SCALES];
  unsigned long long seed, offset;
  int B, H, W, S;
  float min_disp, max_disp;
  int flags;
  float* loss_partial;                              // [S][B][tiles]
  float* ident_sel[SEGSDE_REPROJ_MAX_SCALES];
  float* gdisp[SEGSDE_REPROJ_MAX_SCALES];
  float* gP_partial;                                // [S][F][B][tiles][12]
  float inv_count;
  int strips, bands, rows_per_band;
};

struct Warp {   // everything the gradient chain needs about one (pixel, frame) sample
  float pred[3];
  float dpx[3], dpy[3];   // d pred_c / d ix, d pred_c / d iy, already multiplied by the clip mask
  float px, py, zinv;     // projected pixel coordinates and 1/(Z+eps)
};

// Projects the back-projected point of a pixel with camera-space coordinates (cx,cy,cz) into source frame `P`,
// and samples the three colour planes bilinearly (border padding, align_corners=True).
// pl[c]: base pointer of colour plane c of the sample (uniform over the warp); the four taps are addressed by 32-bit
// offsets from it (one IMAD.WIDE per load instead of 64-bit pointer arithmetic per plane).
template <bool GRAD>
__device__ __forceinline__ void warp_pixel(const float* const (&pl)[3], int H, int W,
                                           const float* P, float cx, float cy, float cz, Warp& o) {
  const float X = P[0] * cx + P[1] * cy + P[2] * cz + P[3];
  const float Y = P[4] * cx + P[5] * cy + P[6] * cz + P[7];
  const float Z = P[8] * cx + P[9] * cy + P[10] * cz + P[11];
  const float z = Z + 1e-7f;
  const float rz = 1.f / z;              // one IEEE reciprocal instead of two divisions (<= 1 ulp more on px, py)
  float px = X * rz, py = Y * rz;
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
  const int o0 = y0 * W, o1 = y1 * W;
  const int onw = o0 + x0, one = o0 + x1, osw = o1 + x0, ose = o1 + x1;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const float vnw = __ldg(pl[c] + onw), vne = __ldg(pl[c] + one);
    const float vsw = __ldg(pl[c] + osw), vse = __ldg(pl[c] + ose);
    o.pred[c] = vnw * nw + vne * ne + vsw * sw + vse * se;
    if (GRAD) {
      o.dpx[c] = mx * ((vne - vnw) * by + (vse - vsw) * ay);
      o.dpy[c] = my * ((vsw - vnw) * bx + (vse - vne) * ax);
    }
  }
  if (GRAD) { o.px = px; o.py = py; o.zinv = rz; }
}

// SSIM of monodepth_layers.py:240-254 for one channel at one centre, evaluated on the 3x3 window SUMS (sx = sum x,
// sxx = sum x^2, sxy = sum x y, ...) instead of the means: with mu = s/9 and sigma = s2/9 - mu^2 both factors of the
// numerator and of the denominator pick up 81, which cancels in the quotient —
//   n1 = 2 sx sy + 81 C1            n2 = 2 (9 sxy - sx sy) + 81 C2
//   d1 = sx^2 + sy^2 + 81 C1        d2 = 9 sxx - sx^2 + 9 syy - sy^2 + 81 C2        ssim = n1 n2 / (d1 d2)
// — so the five divisions by 9 of AvgPool2d never happen.  The y-only parts (ay = sy^2 + 81 C1, by = 9 syy - sy^2 + 81 C2)
// are shared by both frames and come in precomputed.
struct SsimT { float n1, dn, d1, dd, r, iD; };      // dn = n2 - n1, dd = d2 - d1, r = N / D, iD = 1 / D
__device__ __forceinline__ float ssim_sums(float sx, float sxx, float sxy, float sy, float ay, float by, SsimT* t) {
  const float p = sx * sy;
  const float n1 = fmaf(2.f, p, 81.f * 1e-4f);
  const float n2 = fmaf(2.f, fmaf(9.f, sxy, -p), 81.f * 9e-4f);
  const float sx2 = sx * sx;
  const float d1 = sx2 + ay;
  const float d2 = fmaf(9.f, sxx, by) - sx2;
  const float iD = __fdividef(1.f, d1 * d2);      // SSIM in [0,1]: 2 ulp here is far below the 2e-5 loss tolerance
  const float r = (n1 * n2) * iD;
  if (t) { t->n1 = n1; t->dn = n2 - n1; t->d1 = d1; t->dd = d2 - d1; t->r = r; t->iD = iD; }
  return fmaf(r, -0.5f, 0.5f);
}

// left / right neighbour values of v along the strip.  ReflectionPad2d(1) at the image border (column -1 = column 1,
// column W = column W-2) costs nothing here: a halo lane whose column lies outside the image evaluates the REFLECTED
// column instead (xe below), so the lane next to it simply reads its neighbour.
__device__ __forceinline__ void neighbours(float v, float& l, float& r) {
  l = __shfl_up_sync(FULL, v, 1);
  r = __shfl_down_sync(FULL, v, 1);
}

template <int V> struct IC { static constexpr int value = V; };

template <bool GRAD, int F>
__global__ void __launch_bounds__(32, 8) reproj_march_kernel(const ReprojM k) {
  constexpr int LATE = F * 9 + 1 + F * 3 + 3;   // per pixel and row: (dpx3, dpy3, px, py, zinv) per frame, depth, pred, target
  extern __shared__ float smem[];
  float* sP = smem;                                   // [2][12]
  float* sIK = sP + 24;                               // [9] (+3 pad)
  float* stash = sIK + 12;                            // [rows_per_band + 2][2][32] identity candidates of the band's centres
  float* late = stash + (k.rows_per_band + 2) * 64;   // [3][LATE][32] (GRAD)

  const int lane = threadIdx.x;
  const int strip = blockIdx.x, band = blockIdx.y, b = blockIdx.z;
  const int H = k.H, W = k.W;
  const size_t plane = (size_t)H * W;
  const int x = strip * STRIP - 2 + lane;
  const bool xin = x >= 0 && x < W;
  // the column this lane EVALUATES: its own inside the image, the reflected one (-1 -> 1, W -> W-2) in the halo outside;
  // such a lane only serves as the window neighbour of the border column (it owns no centre: xin is false)
  const int xe = min(max(reflect_idx(min(max(x, -(W - 1)), 2 * W - 2), W), 0), W - 1);
  const bool own_col = lane >= 2 && lane < 2 + STRIP && x < W;
  // multiplicity of the left / right centre in this column's adjoint of ReflectionPad2d(1)
  const float ml = (x >= 1 && x < W) ? (x == 1 ? 2.f : 1.f) : 0.f;
  const float mr = (x >= 0 && x + 1 < W) ? (x == W - 2 ? 2.f : 1.f) : 0.f;
  const int ya = band * k.rows_per_band, yb = min(H, ya + k.rows_per_band);
  const bool no_ssim = k.flags & SEGSDE_REPROJ_NO_SSIM;
  const bool avg = (k.flags & SEGSDE_REPROJ_AVG) && F == 2;
  const bool automask = !(k.flags & SEGSDE_REPROJ_NO_AUTOMASK);
  const float w_ssim = no_ssim ? 0.f : 0.85f, w_l1 = no_ssim ? 1.f : 0.15f;
  const int ncand_f = avg ? 1 : F;
  const float dscale = k.max_disp - k.min_disp;
  const int tile = band * k.strips + strip, tiles = k.strips * k.bands;

  if (lane < 12 * F) {
    const int f = lane / 12, e = lane % 12, i = e / 4, j = e % 4;
    const float* Km = k.K + b * 16;
    const float* Tm = k.T[f] + b * 16;
    float acc = 0.f;
#pragma unroll
    for (int q = 0; q < 4; ++q) acc += Km[i * 4 + q] * Tm[q * 4 + j];
    sP[f * 12 + e] = acc;
  }
  if (lane < 9) sIK[lane] = k.invK[b * 16 + (lane / 3) * 4 + (lane % 3)];
  __syncwarp();

  const float* tgt_b = k.tgt + (size_t)b * 3 * plane;
  const float* src_b[2] = {k.src[0] + (size_t)b * 3 * plane, k.src[1] + (size_t)b * 3 * plane};
  const float* const spl[2][3] = {{src_b[0], src_b[0] + plane, src_b[0] + 2 * plane},
                                  {src_b[1], src_b[1] + plane, src_b[1] + 2 * plane}};

  // vertical reflection: window rows of centre c are (c-1, c, c+1) with -1 -> 1 and H -> H-2
  auto vweights = [&](int c, float& wt, float& wb) {
    wt = (c == 0) ? 0.f : (c == H - 1 ? 2.f : 1.f);
    wb = (c == H - 1) ? 0.f : (c == 0 ? 2.f : 1.f);
  };

  // ================= sweep 0: identity candidates of every centre of the band -> stash =================
  if (automask) {
    float xh[3][F][3][3], yh[3][3][2], sv[3][F][3], tv[3][3];
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        yh[i][c][0] = yh[i][c][1] = 0.f; tv[i][c] = 0.f;
#pragma unroll
        for (int f = 0; f < F; ++f) { xh[i][f][c][0] = xh[i][f][c][1] = xh[i][f][c][2] = 0.f; sv[i][f][c] = 0.f; }
      }
    auto step = [&](auto ph, int r) {
      constexpr int cur = decltype(ph)::value, p1 = (cur + 2) % 3, p2 = (cur + 1) % 3;
      if (r >= 0 && r < H) {
        const size_t o = (size_t)r * W + xe;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          const float y = __ldg(tgt_b + c * plane + o);
          float yl, yr;
          neighbours(y, yl, yr);
          tv[cur][c] = y;
          yh[cur][c][0] = yl + y + yr;
          yh[cur][c][1] = fmaf(yl, yl, fmaf(y, y, yr * yr));
#pragma unroll
          for (int f = 0; f < F; ++f) {
            const float v = __ldg(src_b[f] + c * plane + o);
            float vl, vr;
            neighbours(v, vl, vr);
            sv[cur][f][c] = v;
            xh[cur][f][c][0] = vl + v + vr;
            xh[cur][f][c][1] = fmaf(vl, vl, fmaf(v, v, vr * vr));
            xh[cur][f][c][2] = fmaf(vl, yl, fmaf(v, y, vr * yr));
          }
        }
      }
      const int c0 = r - 1;
      if (c0 >= ya - 1 && c0 >= 0 && c0 < H) {
        float wt, wb;
        vweights(c0, wt, wb);
        float ssim_acc[F], l1_acc[F];
#pragma unroll
        for (int f = 0; f < F; ++f) ssim_acc[f] = l1_acc[f] = 0.f;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          const float sy = yh[p1][c][0] + wt * yh[p2][c][0] + wb * yh[cur][c][0];
          const float syy = yh[p1][c][1] + wt * yh[p2][c][1] + wb * yh[cur][c][1];
          const float sy2 = sy * sy;
          const float ay = sy2 + 81.f * 1e-4f, by = fmaf(9.f, syy, 81.f * 9e-4f) - sy2;
#pragma unroll
          for (int f = 0; f < F; ++f) {
            const float sx = xh[p1][f][c][0] + wt * xh[p2][f][c][0] + wb * xh[cur][f][c][0];
            const float sxx = xh[p1][f][c][1] + wt * xh[p2][f][c][1] + wb * xh[cur][f][c][1];
            const float sxy = xh[p1][f][c][2] + wt * xh[p2][f][c][2] + wb * xh[cur][f][c][2];
            l1_acc[f] += fabsf(tv[p1][c] - sv[p1][f][c]);
            const float v = ssim_sums(sx, sxx, sxy, sy, ay, by, nullptr);
            ssim_acc[f] += fminf(fmaxf(v, 0.f), 1.f);
          }
        }
        float id[2] = {0.f, 0.f};
#pragma unroll
        for (int f = 0; f < F; ++f) id[f] = w_ssim * (ssim_acc[f] * (1.f / 3.f)) + w_l1 * (l1_acc[f] * (1.f / 3.f));
        if (avg) id[0] = (id[0] + id[1]) * 0.5f;
        float* st = stash + (c0 - (ya - 1)) * 64 + lane;
        st[0] = id[0]; st[32] = id[1];
      }
    };
    for (int r0 = ya - 2; r0 <= yb + 1; r0 += 3) {
      step(IC<0>{}, r0);
      if (r0 + 1 <= yb + 1) step(IC<1>{}, r0 + 1);
      if (r0 + 2 <= yb + 1) step(IC<2>{}, r0 + 2);
    }
    __syncwarp();
  }

  // ================= sweeps 1..S: one per scale =================
  const float rayx[3] = {sIK[0] * (float)xe + sIK[2], sIK[3] * (float)xe + sIK[5], sIK[6] * (float)xe + sIK[8]};
  for (int s = 0; s < k.S; ++s) {
    const int hs = k.hs[s], ws = k.ws[s];
    const bool dfull = (hs == H && ws == W);
    const float* dispb = k.disp[s] + (size_t)b * hs * ws;
    // horizontal taps of F.interpolate(bilinear, align_corners=False): scale = in/out, src = max(scale*(dst+0.5)-0.5, 0)
    int hx0 = xe, hx1 = xe; float hlx = 0.f, hhx = 1.f;
    const float sy_scale = (float)hs / (float)H;
    if (!dfull) {
      const float sx = (float)ws / (float)W;
      float fx = sx * ((float)xe + 0.5f) - 0.5f; fx = fx < 0.f ? 0.f : fx;
      hx0 = (int)fx; hx1 = hx0 + (hx0 < ws - 1 ? 1 : 0);
      hlx = fx - (float)hx0; hhx = 1.f - hlx;
    }
    auto vtaps = [&](int y, int& y0, int& y1, float& ly, float& hy) {
      if (dfull) { y0 = y1 = y; ly = 0.f; hy = 1.f; return; }
      float fy = sy_scale * ((float)y + 0.5f) - 0.5f; fy = fy < 0.f ? 0.f : fy;
      y0 = (int)fy; y1 = y0 + (y0 < hs - 1 ? 1 : 0);
      ly = fy - (float)y0; hy = 1.f - ly;
    };
    const float* noise_s = k.noise[s];
    float* sel_s = k.ident_sel[s];
    float* gdisp_s = GRAD ? k.gdisp[s] + (size_t)b * hs * ws : nullptr;

    float xh[3][F][3][3], yh[3][3][2], pr[3][F][3], tg[3][3], kh[3][F][3][3];
    float wprev[F], gP[F][12];
    float loss_acc = 0.f;
#pragma unroll
    for (int f = 0; f < F; ++f) {
      wprev[f] = 0.f;
#pragma unroll
      for (int e = 0; e < 12; ++e) gP[f][e] = 0.f;
    }
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        yh[i][c][0] = yh[i][c][1] = 0.f; tg[i][c] = 0.f;
#pragma unroll
        for (int f = 0; f < F; ++f) {
          pr[i][f][c] = 0.f;
#pragma unroll
          for (int q = 0; q < 3; ++q) { xh[i][f][c][q] = 0.f; kh[i][f][c][q] = 0.f; }
        }
      }

    auto step = [&](auto ph, int r) {
      constexpr int cur = decltype(ph)::value, p1 = (cur + 2) % 3, p2 = (cur + 1) % 3;
      // ---- (a) row r: warp both source frames, horizontal window sums ---------------------------------------
      if (r >= 0 && r < H) {
        float pred[F][3];
#pragma unroll
        for (int f = 0; f < F; ++f) pred[f][0] = pred[f][1] = pred[f][2] = 0.f;
        float y3[3];
        {
          const size_t o = (size_t)r * W + xe;
#pragma unroll
          for (int c = 0; c < 3; ++c) y3[c] = __ldg(tgt_b + c * plane + o);
          int y0, y1; float ly, hy;
          vtaps(r, y0, y1, ly, hy);
          float d;
          if (dfull) d = __ldg(dispb + (size_t)r * ws + xe);
          else {
            const float v00 = __ldg(dispb + y0 * ws + hx0), v01 = __ldg(dispb + y0 * ws + hx1);
            const float v10 = __ldg(dispb + y1 * ws + hx0), v11 = __ldg(dispb + y1 * ws + hx1);
            d = hy * (hhx * v00 + hlx * v01) + ly * (hhx * v10 + hlx * v11);
          }
          const float depth = 1.f / (k.min_disp + dscale * d);
          const float fy = (float)r;
          const float cx = depth * (rayx[0] + sIK[1] * fy), cy = depth * (rayx[1] + sIK[4] * fy),
                      cz = depth * (rayx[2] + sIK[7] * fy);
          float* lt = late + (cur * LATE) * 32 + lane;
#pragma unroll
          for (int f = 0; f < F; ++f) {
            Warp wp;
            warp_pixel<GRAD>(spl[f], H, W, sP + f * 12, cx, cy, cz, wp);
#pragma unroll
            for (int c = 0; c < 3; ++c) pred[f][c] = wp.pred[c];
            if (GRAD) {
#pragma unroll
              for (int c = 0; c < 3; ++c) { lt[(f * 9 + c) * 32] = wp.dpx[c]; lt[(f * 9 + 3 + c) * 32] = wp.dpy[c]; }
              lt[(f * 9 + 6) * 32] = wp.px; lt[(f * 9 + 7) * 32] = wp.py; lt[(f * 9 + 8) * 32] = wp.zinv;
#pragma unroll
              for (int c = 0; c < 3; ++c) lt[(F * 9 + 1 + f * 3 + c) * 32] = wp.pred[c];
            }
          }
          if (GRAD) {
            lt[(F * 9) * 32] = depth;
#pragma unroll
            for (int c = 0; c < 3; ++c) lt[(F * 9 + 1 + F * 3 + c) * 32] = y3[c];
          }
        }
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          const float y = y3[c];
          float yl, yr;
          neighbours(y, yl, yr);
          tg[cur][c] = y;
          yh[cur][c][0] = yl + y + yr;
          yh[cur][c][1] = fmaf(yl, yl, fmaf(y, y, yr * yr));
#pragma unroll
          for (int f = 0; f < F; ++f) {
            const float v = pred[f][c];
            float vl, vr;
            neighbours(v, vl, vr);
            pr[cur][f][c] = v;
            xh[cur][f][c][0] = vl + v + vr;
            xh[cur][f][c][1] = fmaf(vl, vl, fmaf(v, v, vr * vr));
            xh[cur][f][c][2] = fmaf(vl, yl, fmaf(v, y, vr * yr));
          }
        }
      }
      // ---- (b) centre row c0 = r - 1: candidates, selection, weighted gradient coefficients --------------------
      const int c0 = r - 1;
      float wnew[F];
#pragma unroll
      for (int f = 0; f < F; ++f) wnew[f] = 0.f;
      float coef[F][3][3];
#pragma unroll
      for (int f = 0; f < F; ++f)
#pragma unroll
        for (int c = 0; c < 3; ++c) coef[f][c][0] = coef[f][c][1] = coef[f][c][2] = 0.f;
      const bool centre_row = (c0 >= ya - 1 && c0 >= 0 && c0 < H);
      if (centre_row) {
        float wt, wb;
        vweights(c0, wt, wb);
        float ssim_acc[F], l1_acc[F];
#pragma unroll
        for (int f = 0; f < F; ++f) ssim_acc[f] = l1_acc[f] = 0.f;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          const float sy = yh[p1][c][0] + wt * yh[p2][c][0] + wb * yh[cur][c][0];
          const float syy = yh[p1][c][1] + wt * yh[p2][c][1] + wb * yh[cur][c][1];
          const float sy2 = sy * sy;
          const float ay = sy2 + 81.f * 1e-4f, by = fmaf(9.f, syy, 81.f * 9e-4f) - sy2;
#pragma unroll
          for (int f = 0; f < F; ++f) {
            const float sx = xh[p1][f][c][0] + wt * xh[p2][f][c][0] + wb * xh[cur][f][c][0];
            const float sxx = xh[p1][f][c][1] + wt * xh[p2][f][c][1] + wb * xh[cur][f][c][1];
            const float sxy = xh[p1][f][c][2] + wt * xh[p2][f][c][2] + wb * xh[cur][f][c][2];
            l1_acc[f] += fabsf(tg[p1][c] - pr[p1][f][c]);
            SsimT t;
            const float v = ssim_sums(sx, sxx, sxy, sy, ay, by, GRAD ? &t : nullptr);
            ssim_acc[f] += fminf(fmaxf(v, 0.f), 1.f);
            if (GRAD) {
              // d(w_ssim/3 * clamp(ssim)) / d x_i = A + B x_i + C y_i for every x_i of this window (zero where the clamp is
              // active); in the sum form, with k = (w_ssim / 3) / 9 and g = 81 k / D:
              //   C = -g n1      B = g r d1      A = -(g / 9) (sy (n2 - n1) - r sx (d2 - d1))
              const float g = (v >= 0.f && v <= 1.f) ? (81.f * (w_ssim / 27.f)) * t.iD : 0.f;
              const float gr = g * t.r;
              coef[f][c][0] = (g * (1.f / 9.f)) * fmaf(t.r, sx * t.dd, -(sy * t.dn));
              coef[f][c][1] = gr * t.d1;
              coef[f][c][2] = -(g * t.n1);
            }
          }
        }
        float rp[2] = {0.f, 0.f}, id[2] = {0.f, 0.f};
#pragma unroll
        for (int f = 0; f < F; ++f) rp[f] = w_ssim * (ssim_acc[f] * (1.f / 3.f)) + w_l1 * (l1_acc[f] * (1.f / 3.f));
        if (avg) rp[0] = (rp[0] + rp[1]) * 0.5f;
        // candidates in the reference's order: identity (+noise) first, then reprojection
        float best = 3.4e38f; int best_i = -1, n_id = 0;
        if (automask) {
          n_id = ncand_f;
          const float* st = stash + (c0 - (ya - 1)) * 64 + lane;
          id[0] = st[0]; id[1] = st[32];
          float nz[2] = {0.f, 0.f};
          if (xin) {
            if (noise_s) {
              nz[0] = __ldg(noise_s + (((size_t)b * ncand_f) * H + c0) * W + x);
              if (ncand_f == 2) nz[1] = __ldg(noise_s + (((size_t)b * ncand_f + 1) * H + c0) * W + x);
            } else {
              // tie-break noise ~ 1e-5 N(0,1): Philox4x32-7 (the shortest variant that passes BigCrush) + Box-Muller on
              // the fast intrinsics — it only has to decorrelate exact ties between the identity candidates
              Philox phx(k.seed, ((unsigned long long)b * H + c0) * W + x, k.offset + s);
#pragma unroll
              for (int rr = 0; rr < 7; ++rr) phx.round();
              const float rad = sqrtf(-2.f * __logf(u01(phx.c[0])));
              float sn, cs;
              __sincosf(6.283185307179586f * u01(phx.c[1]), &sn, &cs);
              nz[0] = rad * cs * 1e-5f; nz[1] = rad * sn * 1e-5f;
            }
          }
          best = id[0] + nz[0]; best_i = 0;
          if (ncand_f == 2) { const float v1 = id[1] + nz[1]; if (v1 < best) { best = v1; best_i = 1; } }
        }
        if (rp[0] < best) { best = rp[0]; best_i = n_id; }
        if (ncand_f == 2 && rp[1] < best) { best = rp[1]; best_i = n_id + 1; }
        const bool reproj_won = best_i >= n_id;
        if (own_col && c0 >= ya && c0 < yb) {
          loss_acc += best;
          if (sel_s) sel_s[(size_t)b * plane + (size_t)c0 * W + x] = reproj_won ? 1.f : 0.f;
        }
        if (GRAD && reproj_won && xin) {
          if (avg) { wnew[0] = 0.5f; wnew[F - 1] = 0.5f; }
          else if (best_i == n_id) wnew[0] = 1.f;
          else wnew[F - 1] = 1.f;
        }
      }
      if (GRAD) {
        // horizontal box sum of the weighted coefficient fields with the multiplicities of the reflection adjoint
#pragma unroll
        for (int f = 0; f < F; ++f)
#pragma unroll
          for (int c = 0; c < 3; ++c)
#pragma unroll
            for (int q = 0; q < 3; ++q) {
              const float v = wnew[f] * coef[f][c][q];
              const float l = __shfl_up_sync(FULL, v, 1), rr = __shfl_down_sync(FULL, v, 1);
              kh[cur][f][c][q] = fmaf(ml, l, fmaf(mr, rr, v));
            }
        // ---- (c) gradient row g = r - 2 -------------------------------------------------------------------------
        const int g = r - 2;
        if (g >= ya && g < yb && own_col) {
          const float mvt = (g >= 1) ? (g == 1 ? 2.f : 1.f) : 0.f;
          const float mvb = (g + 1 < H) ? (g == H - 2 ? 2.f : 1.f) : 0.f;
          const float* lt = late + (p2 * LATE) * 32 + lane;
          const float depth = lt[(F * 9) * 32];
          const float fy = (float)g;
          const float ray[3] = {rayx[0] + sIK[1] * fy, rayx[1] + sIK[4] * fy, rayx[2] + sIK[7] * fy};
          const float cam[4] = {depth * ray[0], depth * ray[1], depth * ray[2], 1.f};
          float gdepth = 0.f;
#pragma unroll
          for (int f = 0; f < F; ++f) {
            float gix = 0.f, giy = 0.f;
#pragma unroll
            for (int c = 0; c < 3; ++c) {
              const float xq = lt[(F * 9 + 1 + f * 3 + c) * 32], yq = lt[(F * 9 + 1 + F * 3 + c) * 32];
              const float sa = kh[p1][f][c][0] + mvt * kh[p2][f][c][0] + mvb * kh[cur][f][c][0];
              const float sb = kh[p1][f][c][1] + mvt * kh[p2][f][c][1] + mvb * kh[cur][f][c][1];
              const float sc = kh[p1][f][c][2] + mvt * kh[p2][f][c][2] + mvb * kh[cur][f][c][2];
              const float diff = xq - yq;
              const float sgn = diff > 0.f ? 1.f : (diff < 0.f ? -1.f : 0.f);
              const float gp = (sa + sb * xq + sc * yq + wprev[f] * (w_l1 / 3.f) * sgn) * k.inv_count;
              gix = fmaf(gp, lt[(f * 9 + c) * 32], gix);
              giy = fmaf(gp, lt[(f * 9 + 3 + c) * 32], giy);
            }
            const float px = lt[(f * 9 + 6) * 32], py = lt[(f * 9 + 7) * 32], zinv = lt[(f * 9 + 8) * 32];
            // d ix / d px = 1 (the (W-1)/2 of grid_sample cancels Project3D's 2/(W-1))
            const float gX = gix * zinv, gY = giy * zinv;
            const float gZ = -(gix * px + giy * py) * zinv;
            const float gv[3] = {gX, gY, gZ};
#pragma unroll
            for (int i = 0; i < 3; ++i)
#pragma unroll
              for (int j = 0; j < 4; ++j) gP[f][i * 4 + j] = fmaf(gv[i], cam[j], gP[f][i * 4 + j]);
            const float* P = sP + f * 12;
            const float gcx = P[0] * gX + P[4] * gY + P[8] * gZ;
            const float gcy = P[1] * gX + P[5] * gY + P[9] * gZ;
            const float gcz = P[2] * gX + P[6] * gY + P[10] * gZ;
            gdepth += gcx * ray[0] + gcy * ray[1] + gcz * ray[2];
          }
          // depth = 1/(min_disp + dscale*d)  ->  d depth / d d = -dscale * depth^2
          const float gd = -gdepth * dscale * depth * depth;
          if (dfull) {
            gdisp_s[(size_t)g * ws + x] += gd;          // exactly one writer per element
          } else {
            int y0, y1; float ly, hy;
            vtaps(g, y0, y1, ly, hy);
            atomicAdd(gdisp_s + y0 * ws + hx0, gd * (hy * hhx)); atomicAdd(gdisp_s + y0 * ws + hx1, gd * (hy * hlx));
            atomicAdd(gdisp_s + y1 * ws + hx0, gd * (ly * hhx)); atomicAdd(gdisp_s + y1 * ws + hx1, gd * (ly * hlx));
          }
        }
#pragma unroll
        for (int f = 0; f < F; ++f) wprev[f] = wnew[f];
      }
    };
    for (int r0 = ya - 2; r0 <= yb + 1; r0 += 3) {
      step(IC<0>{}, r0);
      if (r0 + 1 <= yb + 1) step(IC<1>{}, r0 + 1);
      if (r0 + 2 <= yb + 1) step(IC<2>{}, r0 + 2);
    }
    // per-warp partials (deterministic: fixed shuffle tree, one partial per warp)
    loss_acc = warp_sum(loss_acc);
    if (lane == 0) k.loss_partial[((size_t)s * k.B + b) * tiles + tile] = loss_acc;
    if (GRAD) {
#pragma unroll
      for (int f = 0; f < F; ++f) {
        float mine = 0.f;
#pragma unroll
        for (int e = 0; e < 12; ++e) {
          const float v = warp_sum(gP[f][e]);
          if (lane == e) mine = v;
        }
        if (lane < 12) k.gP_partial[((((size_t)s * F + f) * k.B + b) * tiles + tile) * 12 + lane] = mine;
      }
    }
    __syncwarp();
  }
}

// loss[s] = sum(partials[s]) / count ; gT[s][f][b] = K[:3,:]^T (4x3) * gP (3x4)
__global__ void reproj_finalize_kernel(const float* __restrict__ part, int n, float inv_count,
                                       float* __restrict__ loss_out, const float* __restrict__ gTp,
                                       const float* __restrict__ K, int B, int tiles, int F,
                                       float* __restrict__ gT) {
  __shared__ double sh[256];
  __shared__ float sgp[12];
  const int tid = threadIdx.x;
  const int s = blockIdx.y;
  if (blockIdx.x == 0) {
    double a = 0.0;
    for (int i = tid; i < n; i += 256) a += (double)part[(size_t)s * n + i];
    sh[tid] = a;
    __syncthreads();
    for (int st = 128; st > 0; st >>= 1) { if (tid < st) sh[tid] += sh[tid + st]; __syncthreads(); }
    if (tid == 0) loss_out[s] = (float)(sh[0] * (double)inv_count);
    return;
  }
  // blocks 1.. : one (f,b) pair each
  const int fb = blockIdx.x - 1;
  if (!gTp || fb >= F * B) return;
  const int b = fb % B;
  const float* src = gTp + ((size_t)s * F * B + fb) * tiles * 12;
  for (int e = 0; e < 12; ++e) {
    double a = 0.0;
    for (int i = tid; i < tiles; i += 256) a += (double)src[(size_t)i * 12 + e];
    sh[tid] = a;
    __syncthreads();
    for (int st = 128; st > 0; st >>= 1) { if (tid < st) sh[tid] += sh[tid + st]; __syncthreads(); }
    if (tid == 0) sgp[e] = (float)sh[0];
    __syncthreads();
  }
  if (tid < 16) {
    const int r = tid / 4, c = tid % 4;   // gT[r][c] = sum_i K[i][r] * gP[i][c], i<3
    const float* Km = K + b * 16;
    float a = 0.f;
    for (int i = 0; i < 3; ++i) a += Km[i * 4 + r] * sgp[i * 4 + c];
    gT[((size_t)s * F * B + fb) * 16 + tid] = a;
  }
}

// F.interpolate(mode="bilinear", align_corners=False) with an explicit output size:
// scale = in/out, src = max(scale*(dst+0.5)-0.5, 0)   (ATen area_pixel_compute_source_index)
__device__ __forceinline__ float disp_up(const float* __restrict__ d, int hs, int ws, int H, int W, int x, int y) {
  if (hs == H && ws == W) return __ldg(d + y * ws + x);
  const float sy = (float)hs / (float)H, sx = (float)ws / (float)W;
  float fy = sy * ((float)y + 0.5f) - 0.5f; fy = fy < 0.f ? 0.f : fy;
  float fx = sx * ((float)x + 0.5f) - 0.5f; fx = fx < 0.f ? 0.f : fx;
  int y0 = (int)fy, x0 = (int)fx;
  int y1 = y0 + (y0 < hs - 1 ? 1 : 0), x1 = x0 + (x0 < ws - 1 ? 1 : 0);
  float ly = fy - (float)y0, lx = fx - (float)x0;
  float hy = 1.f - ly, hx = 1.f - lx;
  float v00 = __ldg(d + y0 * ws + x0), v01 = __ldg(d + y0 * ws + x1);
  float v10 = __ldg(d + y1 * ws + x0), v11 = __ldg(d + y1 * ws + x1);
  return hy * (hx * v00 + lx * v01) + ly * (hx * v10 + lx * v11);
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
  const float d = disp_up(disp + (size_t)b * hs * ws, hs, ws, H, W, x, y);
  const float depth = 1.f / (min_disp + (max_disp - min_disp) * d);
  if (depth_o) depth_o[b * plane + o] = depth;
  if (!T) return;
  const float* ik = invK + b * 16;
  const float fx = (float)x, fy = (float)y;
  const float r0 = ik[0] * fx + ik[1] * fy + ik[2], r1 = ik[4] * fx + ik[5] * fy + ik[6],
              r2 = ik[8] * fx + ik[9] * fy + ik[10];
  Warp w;
  const float* const mpl[3] = {src + (size_t)b * 3 * plane, src + ((size_t)b * 3 + 1) * plane, src + ((size_t)b * 3 + 2) * plane};
  if (src) warp_pixel<false>(mpl, H, W, P, depth * r0, depth * r1, depth * r2, w);
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

static int rows_per_band(int H) {     // ~32 rows per warp sweep: halo overhead (R+4)/R = 1.125, >= 6 waves at 512x1024, B=12
  const int bands = cdiv(H, 32);
  return cdiv(H, bands);
}

template <bool GRAD, int F>
static int launch_march(const ReprojM& k, cudaStream_t st) {
  constexpr int LATE = F * 9 + 1 + F * 3 + 3;
  const size_t sm = sizeof(float) * (size_t)(36 + (k.rows_per_band + 2) * 64 + (GRAD ? 3 * LATE * 32 : 0));
  static int attr = 0;
  if ((int)sm > attr) {
    if (cudaFuncSetAttribute(reproj_march_kernel<GRAD, F>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm) != cudaSuccess) {
      cudaGetLastError();
      return SEGSDE_E_UNSUPPORTED;
    }
    attr = (int)sm;
  }
  dim3 grid(k.strips, k.bands, k.B);
  reproj_march_kernel<GRAD, F><<<grid, 32, sm, st>>>(k);
  return launched();
}

}  // namespace segsde

using namespace segsde;

extern "C" int segsde_reproj_tiles(int H, int W) { return cdiv(W, STRIP) * cdiv(H, rows_per_band(H)); }
extern "C" int segsde_reproj_num_partials(int B, int H, int W) { return B * segsde_reproj_tiles(H, W); }

extern "C" int segsde_reproj_fused(const segsde_reproj_args_t* a, void* stream) {
  if (!a || !a->tgt || !a->K || !a->inv_K || !a->loss_partial) return SEGSDE_E_ARG;
  if (a->F < 1 || a->F > 2 || a->B < 1 || a->H < 2 || a->W < 2 || a->S < 1 || a->S > SEGSDE_REPROJ_MAX_SCALES) return SEGSDE_E_ARG;
  for (int f = 0; f < a->F; ++f)
    if (!a->src[f] || !a->T[f]) return SEGSDE_E_ARG;
  const bool grad = a->gdisp[0] != nullptr;
  for (int s = 0; s < a->S; ++s) {
    if (!a->disp[s] || a->hs[s] < 1 || a->ws[s] < 1) return SEGSDE_E_ARG;
    if ((a->gdisp[s] != nullptr) != grad) return SEGSDE_E_ARG;
  }
  if (grad && !a->gT_partial) return SEGSDE_E_ARG;
  if (a->B > 65535) return SEGSDE_E_ARG;
  ReprojM k;
  k.tgt = a->tgt; k.src[0] = a->src[0]; k.src[1] = a->F > 1 ? a->src[1] : a->src[0];
  k.K = a->K; k.invK = a->inv_K;
  k.T[0] = a->T[0]; k.T[1] = a->F > 1 ? a->T[1] : a->T[0];
  for (int s = 0; s < SEGSDE_REPROJ_MAX_SCALES; ++s) {
    const bool on = s < a->S;
    k.disp[s] = on ? a->disp[s] : nullptr; k.hs[s] = on ? a->hs[s] : 0; k.ws[s] = on ? a->ws[s] : 0;
    k.noise[s] = on ? a->noise[s] : nullptr; k.ident_sel[s] = on ? a->ident_sel[s] : nullptr;
    k.gdisp[s] = on ? a->gdisp[s] : nullptr;
  }
  k.seed = a->seed; k.offset = a->offset;
  k.B = a->B; k.H = a->H; k.W = a->W; k.S = a->S;
  k.min_disp = 1.f / a->max_depth; k.max_disp = 1.f / a->min_depth;
  k.flags = a->flags;
  k.loss_partial = a->loss_partial; k.gP_partial = a->gT_partial;
  k.inv_count = (float)(1.0 / ((double)a->B * a->H * a->W));
  k.rows_per_band = rows_per_band(a->H);
  k.strips = cdiv(a->W, STRIP); k.bands = cdiv(a->H, k.rows_per_band);
  cudaStream_t st = as_stream(stream);
  if (grad) return a->F == 2 ? launch_march<true, 2>(k, st) : launch_march<true, 1>(k, st);
  return a->F == 2 ? launch_march<false, 2>(k, st) : launch_march<false, 1>(k, st);
}

extern "C" int segsde_reproj_finalize(const float* loss_partial, int n_partial, int64_t count, int S,
                                      float* loss_out, const float* gT_partial, const float* K, int B,
                                      int tiles, int F, float* gT, void* stream) {
  if (!loss_partial || !loss_out || n_partial < 1 || count < 1 || S < 1 || S > SEGSDE_REPROJ_MAX_SCALES) return SEGSDE_E_ARG;
  if (gT_partial && (!K || !gT)) return SEGSDE_E_ARG;
  dim3 grid(1 + (gT_partial ? F * B : 0), S);
  reproj_finalize_kernel<<<grid, 256, 0, as_stream(stream)>>>(
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
