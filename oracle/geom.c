/*
 * geom.c -- oracle for triangulate (VO.m:114-115, CreateLandmarksFromFeatures.m:7) and
 * estworldpose = P3P + MSAC (VO.m:123-127).  TEST INFRASTRUCTURE ONLY.
 *
 * triangulate: linear DLT, A = [x1*P1(3,:)-P1(1,:); y1*P1(3,:)-P1(2,:); x2*P2(3,:)-P2(1,:);
 * y2*P2(3,:)-P2(2,:)], right singular vector of the smallest singular value (one-sided Jacobi),
 * X = V(1:3,4)/V(4,4) (Hartley & Zisserman 12.2).
 *
 * estworldpose: MSAC over 4-point samples; Grunert/Gao law-of-cosines P3P on the first three
 * (quartic in v = s3/s1, Haralick et al. 1994 eq. for A4..A0), 4th point picks the root; cost
 * sum(min(d^2, tau)); adaptive trial bound; no refit.
 *
 * Everything here is + - * / sqrt in FP64 in a fixed order (no libm transcendentals), so a CUDA
 * kernel compiled with -fmad=false reproduces it bit for bit.  Compile with -ffp-contract=off.
 */
#include "vo_oracle.h"
#include <math.h>
#include <string.h>
#include <float.h>

/* ---------------------------------------------------------------- Philox -- */
static inline void mulhilo(uint32_t a, uint32_t b, uint32_t* hi, uint32_t* lo) {
  uint64_t p = (uint64_t)a * b;
  *hi = (uint32_t)(p >> 32); *lo = (uint32_t)p;
}

void vo_oracle_philox4x32(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                          uint32_t k0, uint32_t k1, uint32_t out[4]) {
  for (int r = 0; r < 10; ++r) {
    uint32_t hi0, lo0, hi1, lo1;
    mulhilo(0xD2511F53u, c0, &hi0, &lo0);
    mulhilo(0xCD9E8D57u, c2, &hi1, &lo1);
    uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

void vo_oracle_sample4(uint64_t seed, uint32_t trial, uint32_t n, uint32_t idx[4]) {
  for (uint32_t attempt = 0;; ++attempt) {
    uint32_t r[4];
    vo_oracle_philox4x32(trial, attempt, 0u, 0u, (uint32_t)seed, (uint32_t)(seed >> 32), r);
    for (int k = 0; k < 4; ++k) idx[k] = (uint32_t)(((uint64_t)r[k] * n) >> 32);
    if (idx[0] != idx[1] && idx[0] != idx[2] && idx[0] != idx[3] && idx[1] != idx[2] &&
        idx[1] != idx[3] && idx[2] != idx[3])
      return;
  }
}

/* ----------------------------------------------------------- triangulate -- */
static void dlt_point(const double* p1, const double* p2, const double* P1, const double* P2,
                      double X[4]) {
  double A[4][4], V[4][4];
  for (int k = 0; k < 4; ++k) {
    A[0][k] = p1[0] * P1[8 + k] - P1[k];
    A[1][k] = p1[1] * P1[8 + k] - P1[4 + k];
    A[2][k] = p2[0] * P2[8 + k] - P2[k];
    A[3][k] = p2[1] * P2[8 + k] - P2[4 + k];
  }
  for (int i = 0; i < 4; ++i)
    for (int j = 0; j < 4; ++j) V[i][j] = (i == j) ? 1.0 : 0.0;
  for (int sweep = 0; sweep < 30; ++sweep) {
    int rotated = 0;
    for (int p = 0; p < 3; ++p)
      for (int q = p + 1; q < 4; ++q) {
        double alpha = 0, beta = 0, gamma = 0;
        for (int i = 0; i < 4; ++i) {
          alpha += A[i][p] * A[i][p];
          beta += A[i][q] * A[i][q];
          gamma += A[i][p] * A[i][q];
        }
        if (fabs(gamma) <= 1e-15 * sqrt(alpha * beta) || gamma == 0.0) continue;
        rotated = 1;
        double zeta = (beta - alpha) / (2.0 * gamma);
        double t = 1.0 / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
        if (zeta < 0) t = -t;
        double c = 1.0 / sqrt(1.0 + t * t), s = c * t;
        for (int i = 0; i < 4; ++i) {
          double a = A[i][p], b = A[i][q];
          A[i][p] = c * a - s * b; A[i][q] = s * a + c * b;
          a = V[i][p]; b = V[i][q];
          V[i][p] = c * a - s * b; V[i][q] = s * a + c * b;
        }
      }
    if (!rotated) break;
  }
  int best = 0; double bn = DBL_MAX;
  for (int j = 0; j < 4; ++j) {
    double nn = 0;
    for (int i = 0; i < 4; ++i) nn += A[i][j] * A[i][j];
    if (nn < bn) { bn = nn; best = j; }
  }
  for (int i = 0; i < 4; ++i) X[i] = V[i][best];
}

void vo_oracle_triangulate(const double* pts1, const double* pts2, int n, const double* P1,
                           const double* P2, double* xyz, double* reproj_err, uint8_t* valid) {
  for (int i = 0; i < n; ++i) {
    double X[4];
    dlt_point(pts1 + 2 * i, pts2 + 2 * i, P1, P2, X);
    double x = X[0] / X[3], y = X[1] / X[3], z = X[2] / X[3];
    xyz[3 * i] = x; xyz[3 * i + 1] = y; xyz[3 * i + 2] = z;
    const double* Ps[2] = {P1, P2};
    const double* ps[2] = {pts1 + 2 * i, pts2 + 2 * i};
    double esum = 0; int ok = 1;
    for (int v = 0; v < 2; ++v) {
      const double* P = Ps[v];
      double u = P[0] * x + P[1] * y + P[2] * z + P[3];
      double w = P[4] * x + P[5] * y + P[6] * z + P[7];
      double d = P[8] * x + P[9] * y + P[10] * z + P[11];
      double du = u / d - ps[v][0], dv = w / d - ps[v][1];
      esum += sqrt(du * du + dv * dv);
      if (!(d > 0)) ok = 0;
    }
    if (reproj_err) reproj_err[i] = 0.5 * esum;
    if (valid) valid[i] = (uint8_t)ok;
  }
}

/* ------------------------------------------------------------------ P3P -- */
static double poly3(double A, double B, double C, double x) { return ((x + A) * x + B) * x + C; }

/* a positive real root of x^3 + A x^2 + B x + C with C <= 0 (safeguarded Newton/bisection,
 * fixed iteration cap, arithmetic only) */
static double cubic_pos_root(double A, double B, double C) {
  double m = fabs(A); if (fabs(B) > m) m = fabs(B); if (fabs(C) > m) m = fabs(C);
  double lo = 0.0, hi = 1.0 + m;
  double x = hi;
  for (int it = 0; it < 100; ++it) {
    double f = poly3(A, B, C, x);
    if (f == 0.0) return x;
    if (f > 0) hi = x; else lo = x;
    double df = (3.0 * x + 2.0 * A) * x + B;
    double xn = x - f / df;
    if (!(xn > lo && xn < hi)) xn = 0.5 * (lo + hi);
    if (xn == x) break;
    x = xn;
  }
  return x;
}

static int quad_real(double b, double c, double* r) { /* x^2 + b x + c */
  double disc = b * b - 4.0 * c;
  if (disc < 0) return 0;
  double sq = sqrt(disc);
  double q = (b >= 0) ? -0.5 * (b + sq) : -0.5 * (b - sq);
  r[0] = q;
  r[1] = (q != 0.0) ? c / q : 0.0;
  return 2;
}

/* real roots of a4 x^4 + ... + a0, Ferrari + 3 Newton polish steps; returns count (0..4) */
static int quartic_real(double a4, double a3, double a2, double a1, double a0, double* roots) {
  if (a4 == 0.0) return 0;
  double b = a3 / a4, c = a2 / a4, d = a1 / a4, e = a0 / a4;
  double b2 = b * b;
  double p = c - 0.375 * b2;
  double q = d - 0.5 * b * c + 0.125 * b2 * b;
  double r = e - 0.25 * b * d + 0.0625 * b2 * c - (3.0 / 256.0) * b2 * b2;
  double y[4]; int n = 0;
  double m = cubic_pos_root(p, 0.25 * p * p - r, -0.125 * q * q);
  if (m > 0) {
    double w = sqrt(2.0 * m);
    double h = q / (2.0 * w);
    n += quad_real(w, 0.5 * p + m - h, y + n);
    n += quad_real(-w, 0.5 * p + m + h, y + n);
  } else { /* biquadratic */
    double z[2];
    int nz = quad_real(p, r, z);
    for (int i = 0; i < nz; ++i)
      if (z[i] >= 0) { double s = sqrt(z[i]); y[n++] = s; y[n++] = -s; }
  }
  for (int i = 0; i < n; ++i) {
    double x = y[i] - 0.25 * b;
    for (int it = 0; it < 3; ++it) {
      double f = (((x + b) * x + c) * x + d) * x + e;
      double df = ((4.0 * x + 3.0 * b) * x + 2.0 * c) * x + d;
      if (df == 0.0) break;
      x -= f / df;
    }
    roots[i] = x;
  }
  return n;
}

static void cross3(const double* a, const double* b, double* c) {
  c[0] = a[1] * b[2] - a[2] * b[1]; c[1] = a[2] * b[0] - a[0] * b[2]; c[2] = a[0] * b[1] - a[1] * b[0];
}
static double dot3(const double* a, const double* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }

/* orthonormal frame (rows e1,e2,e3) of the triangle p0,p1,p2; returns 0 if degenerate */
static int tri_frame(const double* p0, const double* p1, const double* p2, double E[9]) {
  double d1[3], d2[3];
  for (int k = 0; k < 3; ++k) { d1[k] = p1[k] - p0[k]; d2[k] = p2[k] - p0[k]; }
  double n1 = sqrt(dot3(d1, d1));
  if (!(n1 > 0)) return 0;
  for (int k = 0; k < 3; ++k) E[k] = d1[k] / n1;
  cross3(E, d2, E + 6);
  double n3 = sqrt(dot3(E + 6, E + 6));
  if (!(n3 > 0)) return 0;
  for (int k = 0; k < 3; ++k) E[6 + k] /= n3;
  cross3(E + 6, E, E + 3);
  return 1;
}

int vo_oracle_p3p_solve(const double f[9], const double X[9], double R[4][9], double t[4][3]) {
  const double *P1 = X, *P2 = X + 3, *P3 = X + 6;
  double d23[3], d13[3], d12[3];
  for (int k = 0; k < 3; ++k) { d23[k] = P2[k] - P3[k]; d13[k] = P1[k] - P3[k]; d12[k] = P1[k] - P2[k]; }
  double a2 = dot3(d23, d23), b2 = dot3(d13, d13), c2 = dot3(d12, d12);
  if (!(a2 > 0 && b2 > 0 && c2 > 0)) return 0;
  double ca = dot3(f + 3, f + 6), cb = dot3(f, f + 6), cg = dot3(f, f + 3);
  double q = (a2 - c2) / b2, ac = (a2 + c2) / b2;
  double A4 = (q - 1.0) * (q - 1.0) - 4.0 * c2 / b2 * ca * ca;
  double A3 = 4.0 * (q * (1.0 - q) * cb - (1.0 - ac) * ca * cg + 2.0 * c2 / b2 * ca * ca * cb);
  double A2 = 2.0 * (q * q - 1.0 + 2.0 * q * q * cb * cb + 2.0 * (b2 - c2) / b2 * ca * ca -
                     4.0 * ac * ca * cb * cg + 2.0 * (b2 - a2) / b2 * cg * cg);
  double A1 = 4.0 * (-q * (1.0 + q) * cb + 2.0 * a2 / b2 * cg * cg * cb - (1.0 - ac) * ca * cg);
  double A0 = (1.0 + q) * (1.0 + q) - 4.0 * a2 / b2 * cg * cg;
  double vs[4];
  int nv = quartic_real(A4, A3, A2, A1, A0, vs);
  double Ew[9];
  if (!tri_frame(P1, P2, P3, Ew)) return 0;
  int ns = 0;
  for (int i = 0; i < nv; ++i) {
    double v = vs[i];
    if (!(v > 0)) continue;
    double den = 2.0 * (cg - v * ca);
    if (den == 0.0) continue;
    double u = ((q - 1.0) * v * v - 2.0 * q * cb * v + 1.0 + q) / den;
    if (!(u > 0)) continue;
    double dd = 1.0 + v * v - 2.0 * v * cb;
    if (!(dd > 0)) continue;
    double s1 = sqrt(b2 / dd), s2 = u * s1, s3 = v * s1;
    double Q[9];
    for (int k = 0; k < 3; ++k) { Q[k] = s1 * f[k]; Q[3 + k] = s2 * f[3 + k]; Q[6 + k] = s3 * f[6 + k]; }
    double Ec[9];
    if (!tri_frame(Q, Q + 3, Q + 6, Ec)) continue;
    /* R = Ec^T * Ew  (world -> camera) */
    double* Rm = R[ns];
    for (int r = 0; r < 3; ++r)
      for (int c = 0; c < 3; ++c)
        Rm[3 * r + c] = Ec[r] * Ew[c] + Ec[3 + r] * Ew[3 + c] + Ec[6 + r] * Ew[6 + c];
    for (int r = 0; r < 3; ++r) t[ns][r] = Q[r] - dot3(Rm + 3 * r, P1);
    int finite = 1;
    for (int k = 0; k < 9; ++k) if (!(fabs(Rm[k]) <= 2.0)) finite = 0;
    for (int k = 0; k < 3; ++k) if (!(fabs(t[ns][k]) < DBL_MAX)) finite = 0;
    if (finite) ++ns;
  }
  return ns;
}

static double reproj_d2(const double* R, const double* t, const double* Xw, const double* uv,
                        const double K[4]) {
  double x = dot3(R, Xw) + t[0], y = dot3(R + 3, Xw) + t[1], z = dot3(R + 6, Xw) + t[2];
  if (!(z > 0)) return DBL_MAX;
  double du = K[0] * x / z + K[2] - uv[0];
  double dv = K[1] * y / z + K[3] - uv[1];
  return du * du + dv * dv;
}

/* sum in the CUDA warp order: lane l accumulates points l, l+32, ...; then a shfl_down tree */
static double warp_order_sum(const double* v, int n) {
  double part[32];
  for (int l = 0; l < 32; ++l) part[l] = 0.0;
  for (int i = 0; i < n; ++i) part[i & 31] += v[i];
  for (int off = 16; off > 0; off >>= 1)
    for (int l = 0; l < off; ++l) part[l] += part[l + off];
  return part[0];
}

int vo_oracle_p3p(const double* img, const double* world, int n, const double K[4],
                  const vo_oracle_p3p_opts* opts, double A[16], uint8_t* inliers, int* n_inliers,
                  int* best_trial, int* trials_run) {
  vo_oracle_p3p_opts o = {1000, 99.0, 1.0, 0, 1};
  if (opts) o = *opts;
  memset(A, 0, sizeof(double) * 16);
  A[0] = A[5] = A[10] = A[15] = 1.0;
  if (n_inliers) *n_inliers = 0;
  if (best_trial) *best_trial = -1;
  if (trials_run) *trials_run = 0;
  if (inliers) memset(inliers, 0, (size_t)(n > 0 ? n : 0));
  if (n < 4) return 1;
  const double tau = o.max_reproj_error * o.max_reproj_error;
  double* cost = (double*)__builtin_malloc(sizeof(double) * n);
  double best_cost = DBL_MAX, bestR[9], bestT[3];
  int best = -1, T = o.max_num_trials, t_run = 0;
  for (int tr = 0; tr < o.max_num_trials; ++tr) {
    if (o.adaptive && tr >= T) break;
    ++t_run;
    uint32_t id[4];
    vo_oracle_sample4(o.seed, (uint32_t)tr, (uint32_t)n, id);
    double f[9], X[9];
    for (int k = 0; k < 3; ++k) {
      double bx = (img[2 * id[k]] - K[2]) / K[0], by = (img[2 * id[k] + 1] - K[3]) / K[1];
      double nn = sqrt(bx * bx + by * by + 1.0);
      f[3 * k] = bx / nn; f[3 * k + 1] = by / nn; f[3 * k + 2] = 1.0 / nn;
      for (int c = 0; c < 3; ++c) X[3 * k + c] = world[3 * id[k] + c];
    }
    double R[4][9], t[4][3];
    int ns = vo_oracle_p3p_solve(f, X, R, t);
    int pick = -1; double pd = DBL_MAX;
    for (int s = 0; s < ns; ++s) {
      double d2 = reproj_d2(R[s], t[s], world + 3 * id[3], img + 2 * id[3], K);
      if (d2 < pd) { pd = d2; pick = s; }
    }
    if (pick < 0) continue;
    int ninl = 0;
    for (int i = 0; i < n; ++i) {
      double d2 = reproj_d2(R[pick], t[pick], world + 3 * i, img + 2 * i, K);
      if (d2 < tau) { cost[i] = d2; ++ninl; } else cost[i] = tau;
    }
    double c = warp_order_sum(cost, n);
    if (c < best_cost) {
      best_cost = c; best = tr;
      memcpy(bestR, R[pick], sizeof(bestR)); memcpy(bestT, t[pick], sizeof(bestT));
      if (o.adaptive) {
        double w = (double)ninl / (double)n;
        double pg = w * w * w * w;            /* P(all four sampled points are inliers) */
        double miss = 1.0 - pg, target = 1.0 - 0.01 * o.confidence;
        int Tn = T;
        if (pg > 0) {                          /* smallest k with miss^k <= target */
          double prod = 1.0; Tn = 0;
          while (Tn < T) { prod *= miss; ++Tn; if (prod <= target) break; }
        }
        if (Tn < T) T = Tn;
      }
    }
  }
  if (trials_run) *trials_run = t_run;
  int status = 2, ninl = 0;
  if (best >= 0) {
    for (int i = 0; i < n; ++i) {
      double d2 = reproj_d2(bestR, bestT, world + 3 * i, img + 2 * i, K);
      int in = d2 < tau;
      if (inliers) inliers[i] = (uint8_t)in;
      ninl += in;
    }
    if (ninl >= 4) {
      status = 0;
      /* camera pose in world: R_wc = R^T, t_wc = -R^T t */
      for (int r = 0; r < 3; ++r) {
        for (int c = 0; c < 3; ++c) A[4 * r + c] = bestR[3 * c + r];
        A[4 * r + 3] = -(bestR[r] * bestT[0] + bestR[3 + r] * bestT[1] + bestR[6 + r] * bestT[2]);
      }
    }
  }
  if (n_inliers) *n_inliers = ninl;
  if (best_trial) *best_trial = best;
  __builtin_free(cost);
  return status;
}
