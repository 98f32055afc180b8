// vo_geom.cu -- triangulate (VO.m:114-115, CreateLandmarksFromFeatures.m:7) and estworldpose
// = P3P + MSAC (VO.m:123-127) on the GPU, FP64.
//
// triangulate: one thread per correspondence, 4x4 DLT system, one-sided Jacobi SVD in registers.
// P3P-MSAC:    one warp per hypothesis.  The warp draws its 4-point sample from a counter-based
//              Philox4x32-10 stream keyed by (seed, trial), solves the Grunert quartic for the first
//              three points (Ferrari + Newton polish, arithmetic only), lets the 4th point pick the
//              root, then scores all N points lane-strided with a shuffle-tree sum.  A second
//              kernel replays MSAC's sequential adaptive stopping rule over the per-trial costs, so
//              the result equals a sequential run with the same seed.
// Compiled with -fmad=false: the arithmetic contract (oracle/geom.c) is FP64
// + - * / sqrt in a fixed order, no contraction.
#include "vo_internal.h"
#include <cfloat>

namespace vo {

// --------------------------------------------------------------------------------- Philox
__device__ __forceinline__ void philox4x32(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                           uint32_t k0, uint32_t k1, uint32_t out[4]) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    const uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

__device__ __forceinline__ void sample4(uint64_t seed, uint32_t trial, uint32_t n, uint32_t idx[4]) {
  for (uint32_t attempt = 0;; ++attempt) {
    uint32_t r[4];
    philox4x32(trial, attempt, 0u, 0u, (uint32_t)seed, (uint32_t)(seed >> 32), r);
#pragma unroll
    for (int k = 0; k < 4; ++k) idx[k] = __umulhi(r[k], n);
    if (idx[0] != idx[1] && idx[0] != idx[2] && idx[0] != idx[3] && idx[1] != idx[2] &&
        idx[1] != idx[3] && idx[2] != idx[3])
      return;
  }
}

// ----------------------------------------------------------------------------- triangulate
__device__ void dlt_point(const double p1[2], const double p2[2], const double* __restrict__ P1,
                          const double* __restrict__ P2, double X[4]) {
  double A[4][4], V[4][4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    A[0][k] = p1[0] * P1[8 + k] - P1[k];
    A[1][k] = p1[1] * P1[8 + k] - P1[4 + k];
    A[2][k] = p2[0] * P2[8 + k] - P2[k];
    A[3][k] = p2[1] * P2[8 + k] - P2[4 + k];
  }
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) V[i][j] = (i == j) ? 1.0 : 0.0;
  for (int sweep = 0; sweep < 30; ++sweep) {
    int rotated = 0;
#pragma unroll
    for (int p = 0; p < 3; ++p)
#pragma unroll
      for (int q = p + 1; q < 4; ++q) {
        double alpha = 0, beta = 0, gamma = 0;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          alpha += A[i][p] * A[i][p];
          beta += A[i][q] * A[i][q];
          gamma += A[i][p] * A[i][q];
        }
        if (fabs(gamma) <= 1e-15 * sqrt(alpha * beta) || gamma == 0.0) continue;
        rotated = 1;
        const double zeta = (beta - alpha) / (2.0 * gamma);
        double t = 1.0 / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
        if (zeta < 0) t = -t;
        const double c = 1.0 / sqrt(1.0 + t * t), s = c * t;
#pragma unroll
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
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    double nn = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) nn += A[i][j] * A[i][j];
    if (nn < bn) { bn = nn; best = j; }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    double v = V[i][0];
    if (best == 1) v = V[i][1];
    if (best == 2) v = V[i][2];
    if (best == 3) v = V[i][3];
    X[i] = v;
  }
}

// P: [2][12] projection matrices in global memory.  pts as n x 2 (row-major: ld=2,stride 1;
// col-major: element (i,c) at [c*n + i]).  grid (ceil(cap/128), problems); problem p uses the
// slices [p*cap, (p+1)*cap) of every array and count np[p*np_stride] (np == null: cap).
template <typename T>
__global__ void __launch_bounds__(128)
triangulate_kernel(const T* __restrict__ pts1, const T* __restrict__ pts2, const int* __restrict__ np,
                   int np_stride, int n_cap, int col_major, const double* __restrict__ P, T* __restrict__ xyz,
                   T* __restrict__ err, uint8_t* __restrict__ valid) {
  const int prob = blockIdx.y;
  int n = np ? np[prob * np_stride] : n_cap;
  if (n > n_cap) n = n_cap;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  pts1 += (size_t)prob * n_cap * 2; pts2 += (size_t)prob * n_cap * 2; xyz += (size_t)prob * n_cap * 3;
  const int ld = col_major ? n : 1, st = col_major ? 1 : 2;
  const double p1[2] = {(double)pts1[(size_t)i * st], (double)pts1[(size_t)i * st + ld]};
  const double p2[2] = {(double)pts2[(size_t)i * st], (double)pts2[(size_t)i * st + ld]};
  const double *P1 = P, *P2 = P + 12;
  double X[4];
  dlt_point(p1, p2, P1, P2, X);
  const double x = X[0] / X[3], y = X[1] / X[3], z = X[2] / X[3];
  const int st3 = col_major ? 1 : 3;
  xyz[(size_t)i * st3] = (T)x;
  xyz[(size_t)i * st3 + ld] = (T)y;
  xyz[(size_t)i * st3 + 2 * ld] = (T)z;
  double esum = 0; int ok = 1;
#pragma unroll
  for (int v = 0; v < 2; ++v) {
    const double* Pv = v ? P2 : P1;
    const double* pv = v ? p2 : p1;
    const double u = Pv[0] * x + Pv[1] * y + Pv[2] * z + Pv[3];
    const double w = Pv[4] * x + Pv[5] * y + Pv[6] * z + Pv[7];
    const double d = Pv[8] * x + Pv[9] * y + Pv[10] * z + Pv[11];
    const double du = u / d - pv[0], dv = w / d - pv[1];
    esum += sqrt(du * du + dv * dv);
    if (!(d > 0)) ok = 0;
  }
  if (err) err[(size_t)prob * n_cap + i] = (T)(0.5 * esum);
  if (valid) valid[(size_t)prob * n_cap + i] = (uint8_t)ok;
}

// ------------------------------------------------------------------------------------- P3P
__device__ __forceinline__ double poly3(double A, double B, double C, double x) { return ((x + A) * x + B) * x + C; }

__device__ double cubic_pos_root(double A, double B, double C) {
  double m = fabs(A); if (fabs(B) > m) m = fabs(B); if (fabs(C) > m) m = fabs(C);
  double lo = 0.0, hi = 1.0 + m;
  double x = hi;
  for (int it = 0; it < 100; ++it) {
    const double f = poly3(A, B, C, x);
    if (f == 0.0) return x;
    if (f > 0) hi = x; else lo = x;
    const double df = (3.0 * x + 2.0 * A) * x + B;
    double xn = x - f / df;
    if (!(xn > lo && xn < hi)) xn = 0.5 * (lo + hi);
    if (xn == x) break;
    x = xn;
  }
  return x;
}

__device__ __forceinline__ int quad_real(double b, double c, double* r) {
  const double disc = b * b - 4.0 * c;
  if (disc < 0) return 0;
  const double sq = sqrt(disc);
  const double q = (b >= 0) ? -0.5 * (b + sq) : -0.5 * (b - sq);
  r[0] = q;
  r[1] = (q != 0.0) ? c / q : 0.0;
  return 2;
}

__device__ int quartic_real(double a4, double a3, double a2, double a1, double a0, double* roots) {
  if (a4 == 0.0) return 0;
  const double b = a3 / a4, c = a2 / a4, d = a1 / a4, e = a0 / a4;
  const double b2 = b * b;
  const double p = c - 0.375 * b2;
  const double q = d - 0.5 * b * c + 0.125 * b2 * b;
  const double r = e - 0.25 * b * d + 0.0625 * b2 * c - (3.0 / 256.0) * b2 * b2;
  double y[4]; int n = 0;
  const double m = cubic_pos_root(p, 0.25 * p * p - r, -0.125 * q * q);
  if (m > 0) {
    const double w = sqrt(2.0 * m);
    const double h = q / (2.0 * w);
    n += quad_real(w, 0.5 * p + m - h, y + n);
    n += quad_real(-w, 0.5 * p + m + h, y + n);
  } else {
    double z[2];
    const int nz = quad_real(p, r, z);
    for (int i = 0; i < nz; ++i)
      if (z[i] >= 0) { const double s = sqrt(z[i]); y[n++] = s; y[n++] = -s; }
  }
  for (int i = 0; i < n; ++i) {
    double x = y[i] - 0.25 * b;
    for (int it = 0; it < 3; ++it) {
      const double f = (((x + b) * x + c) * x + d) * x + e;
      const double df = ((4.0 * x + 3.0 * b) * x + 2.0 * c) * x + d;
      if (df == 0.0) break;
      x -= f / df;
    }
    roots[i] = x;
  }
  return n;
}

__device__ __forceinline__ void cross3(const double* a, const double* b, double* c) {
  c[0] = a[1] * b[2] - a[2] * b[1]; c[1] = a[2] * b[0] - a[0] * b[2]; c[2] = a[0] * b[1] - a[1] * b[0];
}
__device__ __forceinline__ double dot3(const double* a, const double* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }

__device__ int tri_frame(const double* p0, const double* p1, const double* p2, double E[9]) {
  double d1[3], d2[3];
  for (int k = 0; k < 3; ++k) { d1[k] = p1[k] - p0[k]; d2[k] = p2[k] - p0[k]; }
  const double n1 = sqrt(dot3(d1, d1));
  if (!(n1 > 0)) return 0;
  for (int k = 0; k < 3; ++k) E[k] = d1[k] / n1;
  cross3(E, d2, E + 6);
  const double n3 = sqrt(dot3(E + 6, E + 6));
  if (!(n3 > 0)) return 0;
  for (int k = 0; k < 3; ++k) E[6 + k] /= n3;
  cross3(E + 6, E, E + 3);
  return 1;
}

// up to 4 solutions; R row-major world->camera
__device__ int p3p_solve(const double f[9], const double X[9], double R[4][9], double t[4][3]) {
  const double *P1 = X, *P2 = X + 3, *P3 = X + 6;
  double d23[3], d13[3], d12[3];
  for (int k = 0; k < 3; ++k) { d23[k] = P2[k] - P3[k]; d13[k] = P1[k] - P3[k]; d12[k] = P1[k] - P2[k]; }
  const double a2 = dot3(d23, d23), b2 = dot3(d13, d13), c2 = dot3(d12, d12);
  if (!(a2 > 0 && b2 > 0 && c2 > 0)) return 0;
  const double ca = dot3(f + 3, f + 6), cb = dot3(f, f + 6), cg = dot3(f, f + 3);
  const double q = (a2 - c2) / b2, ac = (a2 + c2) / b2;
  const double A4 = (q - 1.0) * (q - 1.0) - 4.0 * c2 / b2 * ca * ca;
  const double A3 = 4.0 * (q * (1.0 - q) * cb - (1.0 - ac) * ca * cg + 2.0 * c2 / b2 * ca * ca * cb);
  const double A2 = 2.0 * (q * q - 1.0 + 2.0 * q * q * cb * cb + 2.0 * (b2 - c2) / b2 * ca * ca -
                           4.0 * ac * ca * cb * cg + 2.0 * (b2 - a2) / b2 * cg * cg);
  const double A1 = 4.0 * (-q * (1.0 + q) * cb + 2.0 * a2 / b2 * cg * cg * cb - (1.0 - ac) * ca * cg);
  const double A0 = (1.0 + q) * (1.0 + q) - 4.0 * a2 / b2 * cg * cg;
  double vs[4];
  const int nv = quartic_real(A4, A3, A2, A1, A0, vs);
  double Ew[9];
  if (!tri_frame(P1, P2, P3, Ew)) return 0;
  int ns = 0;
  for (int i = 0; i < nv; ++i) {
    const double v = vs[i];
    if (!(v > 0)) continue;
    const double den = 2.0 * (cg - v * ca);
    if (den == 0.0) continue;
    const double u = ((q - 1.0) * v * v - 2.0 * q * cb * v + 1.0 + q) / den;
    if (!(u > 0)) continue;
    const double dd = 1.0 + v * v - 2.0 * v * cb;
    if (!(dd > 0)) continue;
    const double s1 = sqrt(b2 / dd), s2 = u * s1, s3 = v * s1;
    double Q[9];
    for (int k = 0; k < 3; ++k) { Q[k] = s1 * f[k]; Q[3 + k] = s2 * f[3 + k]; Q[6 + k] = s3 * f[6 + k]; }
    double Ec[9];
    if (!tri_frame(Q, Q + 3, Q + 6, Ec)) continue;
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

__device__ __forceinline__ double reproj_d2(const double* R, const double* t, const double* Xw,
                                            const double* uv, const double* K) {
  const double x = dot3(R, Xw) + t[0], y = dot3(R + 3, Xw) + t[1], z = dot3(R + 6, Xw) + t[2];
  if (!(z > 0)) return DBL_MAX;
  const double du = K[0] * x / z + K[2] - uv[0];
  const double dv = K[1] * y / z + K[3] - uv[1];
  return du * du + dv * dv;
}

struct P3PProblem {
  const double* img;    // n x 2 row-major
  const double* world;  // n x 3 row-major
  const int* n_ptr;     // device count (may be null -> n)
  int n;
};

// grid (ceil(trials / 4), n_problems), block 128 = 4 warps = 4 hypotheses
__global__ void __launch_bounds__(128)
p3p_hypothesis_kernel(const double* __restrict__ img_base, const double* __restrict__ world_base,
                      const int* __restrict__ n_ptr, int cap, const double* __restrict__ K4,
                      uint64_t seed, int max_trials, double tau, double* __restrict__ hyp_cost,
                      int* __restrict__ hyp_ninl, double* __restrict__ hyp_rt, int trial0,
                      const int* __restrict__ trial_bound, const uint64_t* __restrict__ seed_add) {
  const int prob = blockIdx.y;
  const int trial = trial0 + blockIdx.x * 4 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (trial >= max_trials) return;
  if (trial_bound != nullptr && trial >= trial_bound[prob]) return;   // MSAC would never draw this trial
  const int n = n_ptr[prob] < cap ? n_ptr[prob] : cap;
  const size_t hidx = (size_t)prob * max_trials + trial;
  if (n < 4) { if (lane == 0) { hyp_cost[hidx] = DBL_MAX; hyp_ninl[hidx] = 0; } return; }
  const double* img = img_base + (size_t)prob * cap * 2;
  const double* world = world_base + (size_t)prob * cap * 3;
  const double K[4] = {K4[0], K4[1], K4[2], K4[3]};
  uint32_t id[4];
  // seed_add: a per-call part of the seed kept in device memory, so a captured launch sequence can be replayed with another one
  sample4(seed + (seed_add != nullptr ? *seed_add : 0ull) + (uint64_t)prob * 0x9E3779B97F4A7C15ull, (uint32_t)trial, (uint32_t)n, id);
  double f[9], X[9];
  for (int k = 0; k < 3; ++k) {
    const double bx = (img[2 * id[k]] - K[2]) / K[0], by = (img[2 * id[k] + 1] - K[3]) / K[1];
    const double nn = sqrt(bx * bx + by * by + 1.0);
    f[3 * k] = bx / nn; f[3 * k + 1] = by / nn; f[3 * k + 2] = 1.0 / nn;
    for (int c = 0; c < 3; ++c) X[3 * k + c] = world[3 * id[k] + c];
  }
  double R[4][9], t[4][3];
  const int ns = p3p_solve(f, X, R, t);
  int pick = -1; double pd = DBL_MAX;
  for (int s = 0; s < ns; ++s) {
    const double d2 = reproj_d2(R[s], t[s], world + 3 * id[3], img + 2 * id[3], K);
    if (d2 < pd) { pd = d2; pick = s; }
  }
  if (pick < 0) { if (lane == 0) { hyp_cost[hidx] = DBL_MAX; hyp_ninl[hidx] = 0; } return; }
  double Rp[9], tp[3];
  for (int k = 0; k < 9; ++k) Rp[k] = R[0][k];
  for (int k = 0; k < 3; ++k) tp[k] = t[0][k];
  for (int s = 1; s < 4; ++s)
    if (pick == s) {
      for (int k = 0; k < 9; ++k) Rp[k] = R[s][k];
      for (int k = 0; k < 3; ++k) tp[k] = t[s][k];
    }
  double part = 0.0; int ninl = 0;
  for (int i = lane; i < n; i += 32) {
    const double d2 = reproj_d2(Rp, tp, world + 3 * i, img + 2 * i, K);
    if (d2 < tau) { part += d2; ++ninl; } else part += tau;
  }
  for (int off = 16; off > 0; off >>= 1) {
    part += __shfl_down_sync(0xffffffffu, part, off);
    ninl += __shfl_down_sync(0xffffffffu, ninl, off);
  }
  if (lane == 0) {
    hyp_cost[hidx] = part; hyp_ninl[hidx] = ninl;
    double* o = hyp_rt + hidx * 12;
    for (int k = 0; k < 9; ++k) o[k] = Rp[k];
    for (int k = 0; k < 3; ++k) o[9 + k] = tp[k];
  }
}

// MSAC's sequential loop over precomputed per-trial costs: trial tr is visited iff tr < T, and T shrinks
// whenever the best cost improves (smallest k with (1 - w^4)^k <= 1 - confidence).
__device__ void msac_replay(const double* __restrict__ cost, const int* __restrict__ ninls, int n, int max_trials,
                            int upto, int adaptive, double confidence, int* best_out, int* t_run_out, int* T_out) {
  double best_cost = DBL_MAX; int T = max_trials, best = -1, t_run = 0;
  for (int tr = 0; tr < upto; ++tr) {
    if (adaptive && tr >= T) break;
    ++t_run;
    const double c = cost[tr];
    if (c < best_cost) {
      best_cost = c; best = tr;
      if (adaptive) {
        const double w = (double)ninls[tr] / (double)n;
        const double pg = w * w * w * w;
        const double miss = 1.0 - pg, target = 1.0 - 0.01 * confidence;
        int Tn = T;
        if (pg > 0) {
          double prod = 1.0; Tn = 0;
          while (Tn < T) { prod *= miss; ++Tn; if (prod <= target) break; }
        }
        if (Tn < T) T = Tn;
      }
    }
  }
  *best_out = best; *t_run_out = t_run; *T_out = T;
}

// after the first `head` trials: the largest trial index MSAC can still reach
__global__ void p3p_bound_kernel(const int* __restrict__ n_ptr, int n_prob, int cap, int max_trials, int head, int adaptive,
                                 double confidence, const double* __restrict__ hyp_cost,
                                 const int* __restrict__ hyp_ninl, int* __restrict__ bound) {
  const int prob = blockIdx.x * blockDim.x + threadIdx.x;
  if (prob >= n_prob) return;
  const int n = n_ptr[prob] < cap ? n_ptr[prob] : cap;
  int best, t_run, T = max_trials;
  if (n >= 4) msac_replay(hyp_cost + (size_t)prob * max_trials, hyp_ninl + (size_t)prob * max_trials, n, max_trials,
                          head < max_trials ? head : max_trials, adaptive, confidence, &best, &t_run, &T);
  bound[prob] = adaptive ? T : max_trials;
}

// one warp per problem: replay the sequential MSAC loop, then emit pose + inliers
__global__ void __launch_bounds__(32)
p3p_select_kernel(const double* __restrict__ img_base, const double* __restrict__ world_base,
                  const int* __restrict__ n_ptr, int cap, const double* __restrict__ K4, int max_trials,
                  double tau, double confidence, int adaptive, const double* __restrict__ hyp_cost,
                  const int* __restrict__ hyp_ninl, const double* __restrict__ hyp_rt,
                  double* __restrict__ A_out, uint8_t* __restrict__ inliers, int* __restrict__ status_out,
                  int* __restrict__ info_out) {
  const int prob = blockIdx.x, lane = threadIdx.x;
  const int n = n_ptr[prob] < cap ? n_ptr[prob] : cap;
  double* A = A_out + (size_t)prob * 16;
  if (lane < 16) A[lane] = (lane % 5 == 0) ? 1.0 : 0.0;
  uint8_t* inl = inliers ? inliers + (size_t)prob * cap : nullptr;
  if (inl) for (int i = lane; i < cap; i += 32) inl[i] = 0;
  int* info = info_out + (size_t)prob * 3;
  if (n < 4) {
    if (lane == 0) { status_out[prob] = 1; info[0] = 0; info[1] = -1; info[2] = 0; }
    return;
  }
  const double* cost = hyp_cost + (size_t)prob * max_trials;
  const int* ninls = hyp_ninl + (size_t)prob * max_trials;
  int best = -1, t_run = 0;
  if (lane == 0) { int T; msac_replay(cost, ninls, n, max_trials, max_trials, adaptive, confidence, &best, &t_run, &T); }
  best = __shfl_sync(0xffffffffu, best, 0);
  t_run = __shfl_sync(0xffffffffu, t_run, 0);
  int ninl = 0;
  if (best >= 0) {
    const double* rt = hyp_rt + ((size_t)prob * max_trials + best) * 12;
    double R[9], t[3];
    for (int k = 0; k < 9; ++k) R[k] = rt[k];
    for (int k = 0; k < 3; ++k) t[k] = rt[9 + k];
    const double K[4] = {K4[0], K4[1], K4[2], K4[3]};
    const double* img = img_base + (size_t)prob * cap * 2;
    const double* world = world_base + (size_t)prob * cap * 3;
    for (int i = lane; i < n; i += 32) {
      const int in = reproj_d2(R, t, world + 3 * i, img + 2 * i, K) < tau;
      if (inl) inl[i] = (uint8_t)in;
      ninl += in;
    }
    for (int off = 16; off > 0; off >>= 1) ninl += __shfl_xor_sync(0xffffffffu, ninl, off);
    if (ninl >= 4 && lane == 0) {
      for (int r = 0; r < 3; ++r) {
        for (int c = 0; c < 3; ++c) A[4 * r + c] = R[3 * c + r];
        A[4 * r + 3] = -(R[r] * t[0] + R[3 + r] * t[1] + R[6 + r] * t[2]);
      }
    }
  }
  if (lane == 0) {
    status_out[prob] = (best >= 0 && ninl >= 4) ? 0 : 2;
    info[0] = ninl; info[1] = best; info[2] = t_run;
  }
}

void fill_p3p_opts(const vo_p3p_opts* in, vo_p3p_opts* o) {
  o->max_num_trials = 1000; o->confidence = 99.0; o->max_reproj_error = 1.0; o->seed = 0; o->adaptive = 1;
  if (in) {
    if (in->max_num_trials > 0) o->max_num_trials = in->max_num_trials;
    if (in->confidence > 0) o->confidence = in->confidence;
    if (in->max_reproj_error > 0) o->max_reproj_error = in->max_reproj_error;
    o->seed = in->seed;
    o->adaptive = in->adaptive < 0 ? 1 : (in->adaptive != 0);
  }
}

// Batched device-resident P3P-MSAC.  img: [n_prob][cap][2], world: [n_prob][cap][3], n_dev[n_prob].
int p3p_batch_device(vo_ctx* ctx, const double* img, const double* world, const int* n_dev, int cap,
                     int n_prob, const double* K4_dev, const vo_p3p_opts& o, double* A_dev,
                     uint8_t* inliers_dev, int* status_dev, int* info_dev, cudaStream_t st, const uint64_t* seed_add_dev) {
  double *hc, *hrt; int* hn;
  VO_TRY(dev_buf(ctx, "p3p_cost", (size_t)n_prob * o.max_num_trials, &hc));
  VO_TRY(dev_buf(ctx, "p3p_ninl", (size_t)n_prob * o.max_num_trials, &hn));
  VO_TRY(dev_buf(ctx, "p3p_rt", (size_t)n_prob * o.max_num_trials * 12, &hrt));
  const double tau = o.max_reproj_error * o.max_reproj_error;
  // pass 1: the first P3P_HEAD trials; then the adaptive bound they imply; pass 2: only trials MSAC can
  // still reach (the rest exit immediately) -- identical result, a fraction of the work when inliers abound
  constexpr int P3P_HEAD = 32;
  int* bound; VO_TRY(dev_buf(ctx, "p3p_bound", (size_t)n_prob, &bound));
  const int head = o.max_num_trials < P3P_HEAD ? o.max_num_trials : P3P_HEAD;
  p3p_hypothesis_kernel<<<dim3(div_up(head, 4), n_prob), 128, 0, st>>>(img, world, n_dev, cap, K4_dev, o.seed, o.max_num_trials, tau,
                                                                      hc, hn, hrt, 0, nullptr, seed_add_dev);
  if (o.max_num_trials > head) {
    p3p_bound_kernel<<<div_up(n_prob, 64), 64, 0, st>>>(n_dev, n_prob, cap, o.max_num_trials, head, o.adaptive, o.confidence, hc, hn, bound);
    p3p_hypothesis_kernel<<<dim3(div_up(o.max_num_trials - head, 4), n_prob), 128, 0, st>>>(
        img, world, n_dev, cap, K4_dev, o.seed, o.max_num_trials, tau, hc, hn, hrt, head, bound, seed_add_dev);
  }
  p3p_select_kernel<<<n_prob, 32, 0, st>>>(img, world, n_dev, cap, K4_dev, o.max_num_trials, tau, o.confidence, o.adaptive,
                                           hc, hn, hrt, A_dev, inliers_dev, status_dev, info_dev);
  VO_CUDA(cudaGetLastError());
  return VO_OK;
}

// ------------------------------------------------------------------- landmark map
// VO.m:145-161 + CreateLandmarksFromFeatures.m:1-21 on the device, for frame pair p = (frame p, frame p+1):
//   matched positions of the CURRENT frame i = p + 1:  ml[k] = kps[2i][l0[i][k]], mr[k] = kps[2i+1][r0[i][k]], k < K0[i]
//   "old" = tracked positions of the PREVIOUS frame after find_remaining_points: old_l[p][m], old_r[p][m], m < K4[p]
// A feature is new iff NONE of the old points shares its x OR its y coordinate, left and right (the reference's
// `find(old.Location == loc(index,:), 1)` compares an N x 2 array with a 1 x 2 row: VO.m:147-154).
__global__ void __launch_bounds__(256)
landmark_select_kernel(const vo_keypoint* __restrict__ kps, int kc, const uint32_t* __restrict__ l0, const uint32_t* __restrict__ r0,
                       const int* __restrict__ K0, const double* __restrict__ old_l, const double* __restrict__ old_r,
                       const int* __restrict__ K4, uint32_t* __restrict__ newidx, int* __restrict__ n_new) {
  __shared__ float s_old[4][256];
  __shared__ int s_warp[8];
  __shared__ int s_carry;
  const int p = blockIdx.x, i = p + 1;
  const int n0 = min(K0[i], kc), n4 = min(K4[p], kc);
  const vo_keypoint* kl = kps + (size_t)(2 * i) * kc;
  const vo_keypoint* kr = kps + (size_t)(2 * i + 1) * kc;
  const double* ol = old_l + (size_t)p * kc * 2;
  const double* orr = old_r + (size_t)p * kc * 2;
  const int tid = threadIdx.x, lane = tid & 31, wrp = tid >> 5;
  if (tid == 0) s_carry = 0;
  __syncthreads();
  for (int base = 0; base < n0; base += 256) {
    const int k = base + tid;
    float lx = 0, ly = 0, rx = 0, ry = 0;
    if (k < n0) {
      const vo_keypoint a = kl[l0[(size_t)i * kc + k]], b = kr[r0[(size_t)i * kc + k]];
      lx = a.x; ly = a.y; rx = b.x; ry = b.y;
    }
    bool seen = false;
    for (int m0 = 0; m0 < n4; m0 += 256) {
      __syncthreads();
      if (m0 + tid < n4) {
        s_old[0][tid] = (float)ol[2 * (m0 + tid)]; s_old[1][tid] = (float)ol[2 * (m0 + tid) + 1];
        s_old[2][tid] = (float)orr[2 * (m0 + tid)]; s_old[3][tid] = (float)orr[2 * (m0 + tid) + 1];
      }
      __syncthreads();
      const int mm = min(256, n4 - m0);
      for (int m = 0; m < mm; ++m)
        seen = seen || lx == s_old[0][m] || ly == s_old[1][m] || rx == s_old[2][m] || ry == s_old[3][m];
    }
    const bool fresh = k < n0 && !seen;
    const unsigned bal = __ballot_sync(0xffffffffu, fresh);
    if (lane == 0) s_warp[wrp] = __popc(bal);
    __syncthreads();
    int before = s_carry;
    for (int w = 0; w < wrp; ++w) before += s_warp[w];
    if (fresh) newidx[(size_t)p * kc + before + __popc(bal & ((1u << lane) - 1u))] = (uint32_t)k;
    __syncthreads();
    if (tid == 0) { int t = 0; for (int w = 0; w < 8; ++w) t += s_warp[w]; s_carry += t; }
    __syncthreads();
  }
  if (tid == 0) n_new[p] = s_carry;
}

// Every second new feature (i = 1:2:n) is triangulated, kept when 0 <= z <= 80 and moved to world coordinates with
// the frame's pose (CreateLandmarksFromFeatures.m:4-18).  Row q of the frame's block is feature q of the new list;
// skipped and rejected rows stay zero like the reference's zero-initialised array, rows[p] = last written row + 1.
__global__ void __launch_bounds__(128)
landmark_build_kernel(const vo_keypoint* __restrict__ kps, int kc, const uint32_t* __restrict__ l0, const uint32_t* __restrict__ r0,
                      const uint32_t* __restrict__ newidx, const int* __restrict__ n_new, const int* __restrict__ status,
                      const double* __restrict__ P, const double* __restrict__ poses, double* __restrict__ out, int cap,
                      int* __restrict__ rows) {
  const int p = blockIdx.y, i = p + 1;
  if (status[p] != 0) return;                               // VO.m stops at a failed estworldpose: no landmarks
  const int n = min(n_new[p], cap);
  const int q = 2 * (blockIdx.x * 128 + threadIdx.x);
  if (q >= n) return;
  const uint32_t k = newidx[(size_t)p * kc + q];
  const vo_keypoint a = kps[(size_t)(2 * i) * kc + l0[(size_t)i * kc + k]];
  const vo_keypoint b = kps[(size_t)(2 * i + 1) * kc + r0[(size_t)i * kc + k]];
  const double p1[2] = {(double)a.x, (double)a.y}, p2[2] = {(double)b.x, (double)b.y};
  double X[4];
  dlt_point(p1, p2, P, P + 12, X);
  const double x = X[0] / X[3], y = X[1] / X[3], z = X[2] / X[3];
  if (z < 0 || z > 80) return;
  const double* A = poses + (size_t)i * 16;
  double* o = out + ((size_t)i * cap + q) * 3;
  o[0] = ((x * A[0] + y * A[1]) + z * A[2]) + A[3];
  o[1] = ((x * A[4] + y * A[5]) + z * A[6]) + A[7];
  o[2] = ((x * A[8] + y * A[9]) + z * A[10]) + A[11];
  atomicMax(&rows[i], q + 1);
}

int landmarks_device(const vo_keypoint* kps, int kc, const uint32_t* l0, const uint32_t* r0, const int* K0, const double* old_l,
                     const double* old_r, const int* K4, const int* status, const double* P_dev, const double* poses_dev,
                     int n_frames, uint32_t* newidx, int* n_new, double* out, int cap, int* rows, cudaStream_t st) {
  if (n_frames < 2) return VO_OK;
  landmark_select_kernel<<<n_frames - 1, 256, 0, st>>>(kps, kc, l0, r0, K0, old_l, old_r, K4, newidx, n_new);
  landmark_build_kernel<<<dim3(div_up(div_up(cap, 2), 128), n_frames - 1), 128, 0, st>>>(kps, kc, l0, r0, newidx, n_new, status, P_dev,
                                                                                        poses_dev, out, cap, rows);
  VO_CUDA(cudaGetLastError());
  return VO_OK;
}

int triangulate_batch_device(const double* pts1, const double* pts2, const int* n_dev, int n_stride, int cap, int n_prob,
                             const double* P_dev, double* xyz, cudaStream_t st) {
  if (cap <= 0 || n_prob <= 0) return VO_OK;
  triangulate_kernel<double><<<dim3(div_up(cap, 128), n_prob), 128, 0, st>>>(pts1, pts2, n_dev, n_stride, cap, 0, P_dev, xyz, nullptr, nullptr);
  VO_CUDA(cudaGetLastError());
  return VO_OK;
}

}  // namespace vo

using namespace vo;

extern "C" {

int vo_triangulate(vo_ctx* ctx, const void* pts1, const void* pts2, int n, int is_double, int col_major,
                   const double P1[12], const double P2[12], void* xyz, void* reproj_err, uint8_t* valid) {
  VO_CHECK_ARG(ctx && P1 && P2, "null argument");
  VO_CHECK_ARG(n >= 0, "negative n");
  if (n == 0) return VO_OK;
  VO_CHECK_ARG(pts1 && pts2 && xyz, "null point/output pointer");
  VO_CUDA(cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream;
  const size_t es = is_double ? sizeof(double) : sizeof(float);
  char *d1, *d2, *dx, *de; uint8_t* dv; double* dP;
  VO_TRY(dev_buf(ctx, "tri_p1", (size_t)n * 2 * es, &d1));
  VO_TRY(dev_buf(ctx, "tri_p2", (size_t)n * 2 * es, &d2));
  VO_TRY(dev_buf(ctx, "tri_xyz", (size_t)n * 3 * es, &dx));
  VO_TRY(dev_buf(ctx, "tri_err", (size_t)n * es, &de));
  VO_TRY(dev_buf(ctx, "tri_valid", (size_t)n, &dv));
  VO_TRY(dev_buf(ctx, "tri_P", 24, &dP));
  VO_CUDA(cudaMemcpyAsync(d1, pts1, (size_t)n * 2 * es, cudaMemcpyHostToDevice, st));
  VO_CUDA(cudaMemcpyAsync(d2, pts2, (size_t)n * 2 * es, cudaMemcpyHostToDevice, st));
  VO_CUDA(cudaMemcpyAsync(dP, P1, 12 * sizeof(double), cudaMemcpyHostToDevice, st));
  VO_CUDA(cudaMemcpyAsync(dP + 12, P2, 12 * sizeof(double), cudaMemcpyHostToDevice, st));
  if (is_double)
    triangulate_kernel<double><<<div_up(n, 128), 128, 0, st>>>((const double*)d1, (const double*)d2, nullptr, 0, n, col_major, dP,
                                                               (double*)dx, (double*)de, dv);
  else
    triangulate_kernel<float><<<div_up(n, 128), 128, 0, st>>>((const float*)d1, (const float*)d2, nullptr, 0, n, col_major, dP,
                                                              (float*)dx, (float*)de, dv);
  VO_CUDA(cudaGetLastError());
  VO_CUDA(cudaMemcpyAsync(xyz, dx, (size_t)n * 3 * es, cudaMemcpyDeviceToHost, st));
  if (reproj_err) VO_CUDA(cudaMemcpyAsync(reproj_err, de, (size_t)n * es, cudaMemcpyDeviceToHost, st));
  if (valid) VO_CUDA(cudaMemcpyAsync(valid, dv, (size_t)n, cudaMemcpyDeviceToHost, st));
  VO_CUDA(cudaStreamSynchronize(st));
  return VO_OK;
}

int vo_p3p(vo_ctx* ctx, const double* img, const double* world, int n, int col_major, const double K[4],
           const vo_p3p_opts* opts, double A[16], uint8_t* inliers, int* status, int info[3]) {
  VO_CHECK_ARG(ctx && K && A && status, "null argument");
  VO_CHECK_ARG(n >= 0, "negative n");
  VO_CHECK_ARG(n == 0 || (img && world), "null point pointer");
  VO_CUDA(cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream;
  vo_p3p_opts o; fill_p3p_opts(opts, &o);
  const int cap = n > 0 ? n : 1;
  // stage row-major copies on the host (tiny), then one upload
  double* h; VO_TRY(pin_buf(ctx, "p3p_host", (size_t)cap * 5 + 64, &h));
  for (int i = 0; i < n; ++i) {
    h[2 * i] = col_major ? img[i] : img[2 * i];
    h[2 * i + 1] = col_major ? img[(size_t)n + i] : img[2 * i + 1];
    for (int c = 0; c < 3; ++c) h[(size_t)2 * cap + 3 * i + c] = col_major ? world[(size_t)c * n + i] : world[3 * i + c];
  }
  double* hK = h + (size_t)cap * 5;
  for (int k = 0; k < 4; ++k) hK[k] = K[k];
  ((int*)(hK + 4))[0] = n;
  double* d; VO_TRY(dev_buf(ctx, "p3p_in", (size_t)cap * 5 + 64, &d));
  VO_CUDA(cudaMemcpyAsync(d, h, ((size_t)cap * 5 + 8) * sizeof(double), cudaMemcpyHostToDevice, st));
  double* dA; uint8_t* dinl; int* dst;
  VO_TRY(dev_buf(ctx, "p3p_A", 16, &dA));
  VO_TRY(dev_buf(ctx, "p3p_inl", (size_t)cap, &dinl));
  VO_TRY(dev_buf(ctx, "p3p_status", 8, &dst));
  const double* dK = d + (size_t)cap * 5;
  VO_TRY(p3p_batch_device(ctx, d, d + (size_t)2 * cap, (const int*)(dK + 4), cap, 1, dK, o, dA, dinl, dst, dst + 1, st, nullptr));
  double hA[16]; int hst[4];
  VO_CUDA(cudaMemcpyAsync(hA, dA, sizeof(hA), cudaMemcpyDeviceToHost, st));
  VO_CUDA(cudaMemcpyAsync(hst, dst, sizeof(hst), cudaMemcpyDeviceToHost, st));
  if (inliers && n > 0) VO_CUDA(cudaMemcpyAsync(inliers, dinl, (size_t)n, cudaMemcpyDeviceToHost, st));
  VO_CUDA(cudaStreamSynchronize(st));
  for (int r = 0; r < 4; ++r)
    for (int c = 0; c < 4; ++c) A[col_major ? 4 * c + r : 4 * r + c] = hA[4 * r + c];
  *status = hst[0];
  if (info) { info[0] = hst[1]; info[1] = hst[2]; info[2] = hst[3]; }
  return VO_OK;
}

}  // extern "C"
