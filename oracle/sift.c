/*
 * sift.c -- oracle for detectSIFTFeatures + extractFeatures("Method","SIFT") (VO.m:79-84).
 * TEST INFRASTRUCTURE ONLY.  OpenCV-convention SIFT (Lowe 2004):
 * 2x bilinear upsample, sigma0 = 1.6, 3 layers/octave, incremental separable Gaussian blurs,
 * DoG 3x3x3 extrema, quadratic sub-pixel refinement (<= 5 steps), contrast 0.04/3 and edge 10
 * tests, 36-bin orientation histogram (all peaks >= 0.8 max), 4x4x8 descriptor clipped at 0.2,
 * scaled by 512 and rounded to 0..255, keypoints sorted (x, y, -size, angle, -response, -octave)
 * with exact duplicates dropped.
 *
 * Arithmetic contract shared with the CUDA path (DESIGN.md "SIFT arithmetic"): FP32, every fused
 * multiply-add is an explicit fmaf(), blur = k0*x0 + sum_i k_i*(x[-i]+x[+i]) in ascending i,
 * exp/atan2 are the polynomial forms below, histogram contributions are rounded to 1/4096 and
 * summed as integers (order independent).  Compile with -ffp-contract=off.
 */
#include "vo_oracle.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <float.h>
#include <stddef.h>

#define SIFT_IMG_BORDER 5
#define SIFT_MAX_INTERP_STEPS 5
#define SIFT_ORI_HIST_BINS 36
#define SIFT_FIX 4096.0f
#define SIFT_INV_FIX (1.0f / 4096.0f)

/* ------------------------------------------------------------ primitives -- */
float vo_oracle_expf(float x) {
  float t = x * 1.44269504088896341f;
  float n = rintf(t);
  if (n < -126.f) return 0.f;
  if (n > 127.f) n = 127.f;
  float r = fmaf(n, -0.693145751953125f, x);          /* ln2 hi (exact in 12 bits) */
  r = fmaf(n, -1.42860682030941723e-6f, r);           /* ln2 lo */
  float p = 1.0f / 720.0f;
  p = fmaf(p, r, 1.0f / 120.0f);
  p = fmaf(p, r, 1.0f / 24.0f);
  p = fmaf(p, r, 1.0f / 6.0f);
  p = fmaf(p, r, 0.5f);
  p = fmaf(p, r, 1.0f);
  p = fmaf(p, r, 1.0f);
  union { float f; int32_t i; } u;
  u.f = p;
  u.i += ((int32_t)n) << 23;
  return u.f;
}

float vo_oracle_atan2deg(float y, float x) {
  const float p1 = 0.9997878412794807f * 57.29577951308232f;
  const float p3 = -0.3258083974640975f * 57.29577951308232f;
  const float p5 = 0.1555786518463281f * 57.29577951308232f;
  const float p7 = -0.04432655554792128f * 57.29577951308232f;
  float ax = fabsf(x), ay = fabsf(y), a, c, c2;
  if (ax >= ay) {
    c = ay / (ax + 2.220446049250313e-16f);
    c2 = c * c;
    a = fmaf(fmaf(fmaf(p7, c2, p5), c2, p3), c2, p1) * c;
  } else {
    c = ax / (ay + 2.220446049250313e-16f);
    c2 = c * c;
    a = 90.f - fmaf(fmaf(fmaf(p7, c2, p5), c2, p3), c2, p1) * c;
  }
  if (x < 0) a = 180.f - a;
  if (y < 0) a = 360.f - a;
  return a;
}

static inline int reflect101(int p, int len) {
  if (len == 1) return 0;
  while (p < 0 || p >= len) {
    if (p < 0) p = -p;
    else p = 2 * (len - 1) - p;
  }
  return p;
}

void vo_oracle_gauss_kernel(float sigma, int* radius, float* taps) {
  int ksize = ((int)lrint((double)sigma * 8.0 + 1.0)) | 1;
  int r = ksize / 2;
  double sum = 0, t[128];
  double s2 = -0.5 / ((double)sigma * (double)sigma);
  for (int i = 0; i < ksize; ++i) {
    double x = i - (ksize - 1) * 0.5;
    t[i] = exp(s2 * x * x);
    sum += t[i];
  }
  sum = 1.0 / sum;
  for (int i = 0; i <= r; ++i) taps[i] = (float)(t[r + i] * sum);
  *radius = r;
}

void vo_oracle_blur(const float* src, float* dst, int rows, int cols, float sigma) {
  int R; float k[64];
  vo_oracle_gauss_kernel(sigma, &R, k);
  float* tmp = (float*)malloc(sizeof(float) * (size_t)rows * cols);
  for (int r = 0; r < rows; ++r) {
    const float* s = src + (size_t)r * cols;
    float* d = tmp + (size_t)r * cols;
    for (int c = 0; c < cols; ++c) {
      float acc = k[0] * s[c];
      if (c >= R && c + R < cols) {
        for (int i = 1; i <= R; ++i) acc = fmaf(k[i], s[c - i] + s[c + i], acc);
      } else {
        for (int i = 1; i <= R; ++i)
          acc = fmaf(k[i], s[reflect101(c - i, cols)] + s[reflect101(c + i, cols)], acc);
      }
      d[c] = acc;
    }
  }
  for (int r = 0; r < rows; ++r) {
    float* d = dst + (size_t)r * cols;
    const float* s0 = tmp + (size_t)r * cols;
    for (int c = 0; c < cols; ++c) d[c] = k[0] * s0[c];
    for (int i = 1; i <= R; ++i) {
      const float* sm = tmp + (size_t)reflect101(r - i, rows) * cols;
      const float* sp = tmp + (size_t)reflect101(r + i, rows) * cols;
      float ki = k[i];
      for (int c = 0; c < cols; ++c) d[c] = fmaf(ki, sm[c] + sp[c], d[c]);
    }
  }
  free(tmp);
}

/* dbl = cv::resize(img, 2x, INTER_LINEAR): pixel-centre aligned, dbl(x) = src(x/2 - 0.25) with
 * clamped borders, i.e. weights (0.25, 0.75) / (0.75, 0.25); exact in FP32 for u8 input.  (This is
 * what OpenCV 4.13 does -- measured: keypoints carry the resulting +0.25 px offset.)  Then blur by
 * sig_diff = sqrt(sigma^2 - 4*0.5^2). */
void vo_oracle_base_image(const uint8_t* img, int rows, int cols, int ld, float sigma, float* base) {
  int R2 = rows * 2, C2 = cols * 2;
  float* dbl = (float*)malloc(sizeof(float) * (size_t)R2 * C2);
  for (int y = 0; y < R2; ++y) {
    int ya = (y & 1) ? (y >> 1) : (y >> 1) - 1, yb = ya + 1;
    float wyb = (y & 1) ? 0.25f : 0.75f, wya = 1.0f - wyb;
    if (ya < 0) ya = 0;
    if (yb > rows - 1) yb = rows - 1;
    for (int x = 0; x < C2; ++x) {
      int xa = (x & 1) ? (x >> 1) : (x >> 1) - 1, xb = xa + 1;
      float wxb = (x & 1) ? 0.25f : 0.75f, wxa = 1.0f - wxb;
      if (xa < 0) xa = 0;
      if (xb > cols - 1) xb = cols - 1;
      float a = img[(size_t)ya * ld + xa], b = img[(size_t)ya * ld + xb];
      float c = img[(size_t)yb * ld + xa], d = img[(size_t)yb * ld + xb];
      dbl[(size_t)y * C2 + x] = wya * (wxa * a + wxb * b) + wyb * (wxa * c + wxb * d);
    }
  }
  float sd2 = sigma * sigma - 0.5f * 0.5f * 4.0f;
  float sig_diff = sqrtf(sd2 > 0.01f ? sd2 : 0.01f);
  vo_oracle_blur(dbl, base, R2, C2, sig_diff);
  free(dbl);
}

/* ---------------------------------------------------------------- pyramid -- */
typedef struct {
  int n_oct, nl;            /* octaves, layers per octave (3)           */
  int rows[16], cols[16];
  float** g;                /* n_oct*(nl+3) gaussian images             */
  float** d;                /* n_oct*(nl+2) DoG images                  */
} pyr_t;

static void pyr_free(pyr_t* p) {
  for (int i = 0; i < p->n_oct * (p->nl + 3); ++i) free(p->g[i]);
  for (int i = 0; i < p->n_oct * (p->nl + 2); ++i) free(p->d[i]);
  free(p->g); free(p->d);
}

static void pyr_build(pyr_t* p, const uint8_t* img, int rows, int cols, int ld, int nl, float sigma) {
  int R2 = rows * 2, C2 = cols * 2;
  int mn = R2 < C2 ? R2 : C2;
  int n_oct = (int)lrint(log((double)mn) / log(2.0) - 2.0) + 1;
  if (n_oct < 1) n_oct = 1;
  if (n_oct > 16) n_oct = 16;
  p->n_oct = n_oct; p->nl = nl;
  p->g = (float**)calloc((size_t)n_oct * (nl + 3), sizeof(float*));
  p->d = (float**)calloc((size_t)n_oct * (nl + 2), sizeof(float*));
  float sig[16];
  double k = pow(2.0, 1.0 / nl);
  sig[0] = sigma;
  for (int i = 1; i < nl + 3; ++i) {
    double sp = pow(k, (double)(i - 1)) * sigma, st = sp * k;
    sig[i] = (float)sqrt(st * st - sp * sp);
  }
  for (int o = 0; o < n_oct; ++o) {
    int r = o == 0 ? R2 : p->rows[o - 1] / 2, c = o == 0 ? C2 : p->cols[o - 1] / 2;
    p->rows[o] = r; p->cols[o] = c;
    for (int i = 0; i < nl + 3; ++i) {
      float* dst = (float*)malloc(sizeof(float) * (size_t)(r > 0 ? r : 1) * (c > 0 ? c : 1));
      p->g[o * (nl + 3) + i] = dst;
      if (r == 0 || c == 0) continue;
      if (o == 0 && i == 0) {
        vo_oracle_base_image(img, rows, cols, ld, sigma, dst);
      } else if (i == 0) {
        const float* src = p->g[(o - 1) * (nl + 3) + nl];
        int sc = p->cols[o - 1];
        for (int y = 0; y < r; ++y)
          for (int x = 0; x < c; ++x) dst[(size_t)y * c + x] = src[(size_t)(2 * y) * sc + 2 * x];
      } else {
        vo_oracle_blur(p->g[o * (nl + 3) + i - 1], dst, r, c, sig[i]);
      }
    }
    for (int i = 0; i < nl + 2; ++i) {
      size_t n = (size_t)r * c;
      float* dst = (float*)malloc(sizeof(float) * (n ? n : 1));
      const float *a = p->g[o * (nl + 3) + i + 1], *b = p->g[o * (nl + 3) + i];
      for (size_t q = 0; q < n; ++q) dst[q] = a[q] - b[q];
      p->d[o * (nl + 2) + i] = dst;
    }
  }
}

/* ------------------------------------------------------------- keypoints -- */
typedef struct { float x, y, size, angle, response; int32_t octave; } kp_t;

static int adjust_extremum(const pyr_t* p, int o, int* layer, int* r, int* c, float contrast_thr,
                           float edge_thr, float sigma, kp_t* kp) {
  const int nl = p->nl, rows = p->rows[o], cols = p->cols[o];
  const float img_scale = 1.f / 255.f, deriv_scale = img_scale * 0.5f,
              second_deriv_scale = img_scale, cross_deriv_scale = img_scale * 0.25f;
  float xi = 0, xr = 0, xc = 0;
  int i = 0;
  for (; i < SIFT_MAX_INTERP_STEPS; ++i) {
    const float* img = p->d[o * (nl + 2) + *layer];
    const float* prv = p->d[o * (nl + 2) + *layer - 1];
    const float* nxt = p->d[o * (nl + 2) + *layer + 1];
    size_t q = (size_t)(*r) * cols + *c;
    float dD0 = (img[q + 1] - img[q - 1]) * deriv_scale;
    float dD1 = (img[q + cols] - img[q - cols]) * deriv_scale;
    float dD2 = (nxt[q] - prv[q]) * deriv_scale;
    float v2 = img[q] * 2.f;
    float dxx = (img[q + 1] + img[q - 1] - v2) * second_deriv_scale;
    float dyy = (img[q + cols] + img[q - cols] - v2) * second_deriv_scale;
    float dss = (nxt[q] + prv[q] - v2) * second_deriv_scale;
    float dxy = (img[q + cols + 1] - img[q + cols - 1] - img[q - cols + 1] + img[q - cols - 1]) * cross_deriv_scale;
    float dxs = (nxt[q + 1] - nxt[q - 1] - prv[q + 1] + prv[q - 1]) * cross_deriv_scale;
    float dys = (nxt[q + cols] - nxt[q - cols] - prv[q + cols] + prv[q - cols]) * cross_deriv_scale;
    /* Cramer's rule on H X = dD, H = [dxx dxy dxs; dxy dyy dys; dxs dys dss] */
    float a00 = dxx, a01 = dxy, a02 = dxs, a11 = dyy, a12 = dys, a22 = dss;
    float m0 = a11 * a22 - a12 * a12, m1 = a01 * a22 - a12 * a02, m2 = a01 * a12 - a11 * a02;
    float det = a00 * m0 - a01 * m1 + a02 * m2;
    float X0 = 0, X1 = 0, X2 = 0;
    if (det != 0.f) {
      float d = 1.f / det;
      X0 = d * (dD0 * m0 - a01 * (dD1 * a22 - a12 * dD2) + a02 * (dD1 * a12 - a11 * dD2));
      X1 = d * (a00 * (dD1 * a22 - a12 * dD2) - dD0 * m1 + a02 * (a01 * dD2 - dD1 * a02));
      X2 = d * (a00 * (a11 * dD2 - dD1 * a12) - a01 * (a01 * dD2 - dD1 * a02) + dD0 * m2);
    }
    xi = -X2; xr = -X1; xc = -X0;
    if (fabsf(xi) < 0.5f && fabsf(xr) < 0.5f && fabsf(xc) < 0.5f) break;
    if (fabsf(xi) > (float)(INT32_MAX / 3) || fabsf(xr) > (float)(INT32_MAX / 3) ||
        fabsf(xc) > (float)(INT32_MAX / 3))
      return 0;
    *c += (int)lrintf(xc); *r += (int)lrintf(xr); *layer += (int)lrintf(xi);
    if (*layer < 1 || *layer > nl || *c < SIFT_IMG_BORDER || *c >= cols - SIFT_IMG_BORDER ||
        *r < SIFT_IMG_BORDER || *r >= rows - SIFT_IMG_BORDER)
      return 0;
  }
  if (i >= SIFT_MAX_INTERP_STEPS) return 0;
  {
    const float* img = p->d[o * (nl + 2) + *layer];
    const float* prv = p->d[o * (nl + 2) + *layer - 1];
    const float* nxt = p->d[o * (nl + 2) + *layer + 1];
    size_t q = (size_t)(*r) * cols + *c;
    float dD0 = (img[q + 1] - img[q - 1]) * deriv_scale;
    float dD1 = (img[q + cols] - img[q - cols]) * deriv_scale;
    float dD2 = (nxt[q] - prv[q]) * deriv_scale;
    float t = dD0 * xc + dD1 * xr + dD2 * xi;
    float contr = img[q] * img_scale + t * 0.5f;
    if (fabsf(contr) * nl < contrast_thr) return 0;
    float v2 = img[q] * 2.f;
    float dxx = (img[q + 1] + img[q - 1] - v2) * second_deriv_scale;
    float dyy = (img[q + cols] + img[q - cols] - v2) * second_deriv_scale;
    float dxy = (img[q + cols + 1] - img[q + cols - 1] - img[q - cols + 1] + img[q - cols - 1]) * cross_deriv_scale;
    float tr = dxx + dyy, det = dxx * dyy - dxy * dxy;
    if (det <= 0 || tr * tr * edge_thr >= (edge_thr + 1) * (edge_thr + 1) * det) return 0;
    kp->x = (*c + xc) * (float)(1 << o);
    kp->y = (*r + xr) * (float)(1 << o);
    kp->octave = o + (*layer << 8) + ((int)lrintf((xi + 0.5f) * 255.f) << 16);
    kp->size = sigma * vo_oracle_expf(((*layer + xi) / nl) * 0.693147180559945f) * (float)(1 << o) * 2.f;
    kp->response = fabsf(contr);
  }
  return 1;
}

static float ori_hist(const float* img, int rows, int cols, int px, int py, int radius,
                      float sigma, float* hist) {
  const int n = SIFT_ORI_HIST_BINS;
  uint32_t acc[SIFT_ORI_HIST_BINS];
  memset(acc, 0, sizeof(acc));
  float expf_scale = -1.f / (2.f * sigma * sigma);
  for (int i = -radius; i <= radius; ++i) {
    int y = py + i;
    if (y <= 0 || y >= rows - 1) continue;
    for (int j = -radius; j <= radius; ++j) {
      int x = px + j;
      if (x <= 0 || x >= cols - 1) continue;
      float dx = img[(size_t)y * cols + x + 1] - img[(size_t)y * cols + x - 1];
      float dy = img[(size_t)(y - 1) * cols + x] - img[(size_t)(y + 1) * cols + x];
      float w = vo_oracle_expf((float)(i * i + j * j) * expf_scale);
      float ang = vo_oracle_atan2deg(dy, dx);
      float mag = sqrtf(fmaf(dx, dx, dy * dy));
      int bin = (int)lrintf((n / 360.f) * ang);
      if (bin >= n) bin -= n;
      if (bin < 0) bin += n;
      acc[bin] += (uint32_t)lrintf(w * mag * SIFT_FIX);
    }
  }
  float th[SIFT_ORI_HIST_BINS + 4];
  for (int i = 0; i < n; ++i) th[i + 2] = (float)acc[i] * SIFT_INV_FIX;
  th[1] = th[n + 1]; th[0] = th[n]; th[n + 2] = th[2]; th[n + 3] = th[3];
  float mx = 0;
  for (int i = 0; i < n; ++i) {
    hist[i] = (th[i] + th[i + 4]) * (1.f / 16.f) + (th[i + 1] + th[i + 3]) * (4.f / 16.f) +
              th[i + 2] * (6.f / 16.f);
    if (i == 0 || hist[i] > mx) mx = hist[i];
  }
  return mx;
}

static void descriptor(const float* img, int rows, int cols, float ptx, float pty, float ori,
                       float scl, float* dst) {
  const int d = 4, n = 8;
  int px = (int)lrintf(ptx), py = (int)lrintf(pty);
  double ang = (double)ori * (3.14159265358979323846 / 180.0);
  float cos_t = (float)cos(ang), sin_t = (float)sin(ang);
  float bins_per_rad = n / 360.f, exp_scale = -1.f / (d * d * 0.5f);
  float hist_width = 3.0f * scl;
  int radius = (int)lrintf(hist_width * 1.4142135623730951f * (d + 1) * 0.5f);
  int diag = (int)sqrt((double)cols * cols + (double)rows * rows);
  if (radius > diag) radius = diag;
  cos_t /= hist_width; sin_t /= hist_width;
  uint32_t hist[(4 + 2) * (4 + 2) * (8 + 2)];
  memset(hist, 0, sizeof(hist));
  for (int i = -radius; i <= radius; ++i)
    for (int j = -radius; j <= radius; ++j) {
      float c_rot = j * cos_t - i * sin_t;
      float r_rot = j * sin_t + i * cos_t;
      float rbin = r_rot + d / 2 - 0.5f, cbin = c_rot + d / 2 - 0.5f;
      int r = py + i, c = px + j;
      if (!(rbin > -1 && rbin < d && cbin > -1 && cbin < d && r > 0 && r < rows - 1 && c > 0 && c < cols - 1))
        continue;
      float dx = img[(size_t)r * cols + c + 1] - img[(size_t)r * cols + c - 1];
      float dy = img[(size_t)(r - 1) * cols + c] - img[(size_t)(r + 1) * cols + c];
      float w = vo_oracle_expf((c_rot * c_rot + r_rot * r_rot) * exp_scale);
      float a = vo_oracle_atan2deg(dy, dx);
      float mag = sqrtf(fmaf(dx, dx, dy * dy)) * w;
      float obin = (a - ori) * bins_per_rad;
      int r0 = (int)floorf(rbin), c0 = (int)floorf(cbin), o0 = (int)floorf(obin);
      rbin -= r0; cbin -= c0; obin -= o0;
      if (o0 < 0) o0 += n;
      if (o0 >= n) o0 -= n;
      float v_r1 = mag * rbin, v_r0 = mag - v_r1;
      float v_rc11 = v_r1 * cbin, v_rc10 = v_r1 - v_rc11;
      float v_rc01 = v_r0 * cbin, v_rc00 = v_r0 - v_rc01;
      float v111 = v_rc11 * obin, v110 = v_rc11 - v111;
      float v101 = v_rc10 * obin, v100 = v_rc10 - v101;
      float v011 = v_rc01 * obin, v010 = v_rc01 - v011;
      float v001 = v_rc00 * obin, v000 = v_rc00 - v001;
      int idx = ((r0 + 1) * (d + 2) + c0 + 1) * (n + 2) + o0;
      hist[idx] += (uint32_t)lrintf(v000 * SIFT_FIX);
      hist[idx + 1] += (uint32_t)lrintf(v001 * SIFT_FIX);
      hist[idx + (n + 2)] += (uint32_t)lrintf(v010 * SIFT_FIX);
      hist[idx + (n + 3)] += (uint32_t)lrintf(v011 * SIFT_FIX);
      hist[idx + (d + 2) * (n + 2)] += (uint32_t)lrintf(v100 * SIFT_FIX);
      hist[idx + (d + 2) * (n + 2) + 1] += (uint32_t)lrintf(v101 * SIFT_FIX);
      hist[idx + (d + 3) * (n + 2)] += (uint32_t)lrintf(v110 * SIFT_FIX);
      hist[idx + (d + 3) * (n + 2) + 1] += (uint32_t)lrintf(v111 * SIFT_FIX);
    }
  for (int i = 0; i < d; ++i)
    for (int j = 0; j < d; ++j) {
      int idx = ((i + 1) * (d + 2) + (j + 1)) * (n + 2);
      hist[idx] += hist[idx + n];
      hist[idx + 1] += hist[idx + n + 1];
      for (int k = 0; k < n; ++k) dst[(i * d + j) * n + k] = (float)hist[idx + k] * SIFT_INV_FIX;
    }
  float nrm2 = 0;
  for (int k = 0; k < 128; ++k) nrm2 = fmaf(dst[k], dst[k], nrm2);
  float thr = sqrtf(nrm2) * 0.2f;
  nrm2 = 0;
  for (int k = 0; k < 128; ++k) {
    float v = dst[k] < thr ? dst[k] : thr;
    dst[k] = v;
    nrm2 = fmaf(v, v, nrm2);
  }
  float s = sqrtf(nrm2);
  float scale = 512.f / (s > FLT_EPSILON ? s : FLT_EPSILON);
  for (int k = 0; k < 128; ++k) {
    float v = rintf(dst[k] * scale);
    dst[k] = v > 255.f ? 255.f : (v < 0.f ? 0.f : v);
  }
}

static int kp_less(const void* pa, const void* pb) {
  const kp_t *a = (const kp_t*)pa, *b = (const kp_t*)pb;
  if (a->x != b->x) return a->x < b->x ? -1 : 1;
  if (a->y != b->y) return a->y < b->y ? -1 : 1;
  if (a->size != b->size) return a->size > b->size ? -1 : 1;
  if (a->angle != b->angle) return a->angle < b->angle ? -1 : 1;
  if (a->response != b->response) return a->response > b->response ? -1 : 1;
  if (a->octave != b->octave) return a->octave > b->octave ? -1 : 1;
  return 0;
}

int vo_oracle_sift(const uint8_t* img, int rows, int cols, int ld,
                   const vo_oracle_sift_opts* opts, int capacity, vo_oracle_kp* out, float* desc) {
  vo_oracle_sift_opts o = {3, 0.04f, 10.f, 1.6f};
  if (opts) o = *opts;
  const int nl = o.n_octave_layers;
  pyr_t p;
  pyr_build(&p, img, rows, cols, ld, nl, o.sigma);
  int cap = 1 << 14, nk = 0;
  kp_t* kps = (kp_t*)malloc(sizeof(kp_t) * cap);
  const int threshold = (int)floor(0.5 * o.contrast_threshold / nl * 255.0);
  for (int oc = 0; oc < p.n_oct; ++oc) {
    const int R = p.rows[oc], C = p.cols[oc];
    for (int i = 1; i <= nl; ++i) {
      const float* cur = p.d[oc * (nl + 2) + i];
      const float* prv = p.d[oc * (nl + 2) + i - 1];
      const float* nxt = p.d[oc * (nl + 2) + i + 1];
      for (int r = SIFT_IMG_BORDER; r < R - SIFT_IMG_BORDER; ++r)
        for (int c = SIFT_IMG_BORDER; c < C - SIFT_IMG_BORDER; ++c) {
          size_t q = (size_t)r * C + c;
          float val = cur[q];
          if (!(fabsf(val) > (float)threshold)) continue;
          int ext = 1;
          if (val > 0) {
            for (int dr = -1; dr <= 1 && ext; ++dr)
              for (int dc = -1; dc <= 1; ++dc) {
                size_t qq = q + (ptrdiff_t)dr * C + dc;
                if (!(val >= cur[qq] && val >= prv[qq] && val >= nxt[qq])) { ext = 0; break; }
              }
          } else {
            for (int dr = -1; dr <= 1 && ext; ++dr)
              for (int dc = -1; dc <= 1; ++dc) {
                size_t qq = q + (ptrdiff_t)dr * C + dc;
                if (!(val <= cur[qq] && val <= prv[qq] && val <= nxt[qq])) { ext = 0; break; }
              }
          }
          if (!ext) continue;
          int layer = i, r1 = r, c1 = c;
          kp_t kp;
          if (!adjust_extremum(&p, oc, &layer, &r1, &c1, o.contrast_threshold, o.edge_threshold,
                               o.sigma, &kp))
            continue;
          float scl_octv = kp.size * 0.5f / (float)(1 << oc);
          float hist[SIFT_ORI_HIST_BINS];
          float omax = ori_hist(p.g[oc * (nl + 3) + layer], R, C, c1, r1,
                                (int)lrintf(4.5f * scl_octv), 1.5f * scl_octv, hist);
          float mag_thr = omax * 0.8f;
          const int n = SIFT_ORI_HIST_BINS;
          for (int j = 0; j < n; ++j) {
            int l = j > 0 ? j - 1 : n - 1, r2 = j < n - 1 ? j + 1 : 0;
            if (hist[j] > hist[l] && hist[j] > hist[r2] && hist[j] >= mag_thr) {
              float bin = j + 0.5f * (hist[l] - hist[r2]) / (hist[l] - 2 * hist[j] + hist[r2]);
              bin = bin < 0 ? n + bin : bin >= n ? bin - n : bin;
              kp.angle = 360.f - (360.f / n) * bin;
              if (fabsf(kp.angle - 360.f) < FLT_EPSILON) kp.angle = 0.f;
              if (nk == cap) { cap *= 2; kps = (kp_t*)realloc(kps, sizeof(kp_t) * cap); }
              kps[nk++] = kp;
            }
          }
        }
    }
  }
  qsort(kps, nk, sizeof(kp_t), kp_less);
  int m = 0;
  for (int j = 0; j < nk; ++j) {
    if (j == 0) { m = 1; continue; }
    const kp_t *a = &kps[m - 1], *b = &kps[j];
    if (a->x != b->x || a->y != b->y || a->size != b->size || a->angle != b->angle) kps[m++] = *b;
  }
  nk = nk ? m : 0;
  int ret = nk;
  if (nk > capacity) ret = -1;
  for (int q = 0; q < nk && ret >= 0; ++q) {
    /* first octave is -1: halve coordinates, octave word -1 */
    kp_t k = kps[q];
    k.octave = (k.octave & ~255) | ((k.octave - 1) & 255);
    k.x *= 0.5f; k.y *= 0.5f; k.size *= 0.5f;
    out[q].x = k.x; out[q].y = k.y; out[q].size = k.size; out[q].angle = k.angle;
    out[q].response = k.response; out[q].octave = k.octave;
    if (desc) {
      int oc = k.octave & 255, layer = (k.octave >> 8) & 255;
      oc = oc < 128 ? oc : (-128 | oc);
      float scale = oc >= 0 ? 1.f / (float)(1 << oc) : (float)(1 << -oc);
      float size = k.size * scale;
      float angle = 360.f - k.angle;
      if (fabsf(angle - 360.f) < FLT_EPSILON) angle = 0.f;
      int po = oc + 1;
      descriptor(p.g[po * (nl + 3) + layer], p.rows[po], p.cols[po], k.x * scale, k.y * scale,
                 angle, size * 0.5f, desc + (size_t)q * 128);
    }
  }
  free(kps);
  pyr_free(&p);
  return ret;
}
