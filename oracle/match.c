/*
 * match.c -- oracle for matchFeatures(f1, f2) with the reference's all-default call
 * (VO.m:87, 283, 293, 311, 323): Method Exhaustive, Metric SSD on unit-normalised rows,
 * MatchThreshold 1 %% (SSD <= 0.04), MaxRatio 0.6, Unique false.  TEST INFRASTRUCTURE ONLY.
 *
 * Bit-level definition (shared with the CUDA path, see DESIGN.md "match arithmetic"):
 *   n(x)    = fold_k fmaf(x_k, x_k, acc)          (k ascending, acc0 = 0)
 *   inv(x)  = 1.0f / sqrtf(n(x))                  (0 when n(x) == 0)
 *   dot     = fold_k fmaf(a_k, b_k, acc)
 *   key     = dot * inv(b)        c = key * inv(a)        s = max(fmaf(-2, c, 2), 0)
 * Nearest j1 = argmin_j s (lowest j on ties); s2 = min_{j != j1} s.
 * Keep row i iff s1 <= 0.04*MatchThreshold and (n2 < 2 or ratio <= MaxRatio) where
 * ratio = (s2 < 1e-6f) ? 1 : s1 / s2.
 * Compile with -ffp-contract=off: every fused operation is an explicit fmaf().
 */
#include "vo_oracle.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>
#if defined(__AVX2__) && defined(__FMA__)
#include <immintrin.h>
#endif

static float* to_row_major(const float* f, int n, int dim, int col_major) {
  float* r = (float*)malloc(sizeof(float) * (size_t)(n > 0 ? n : 1) * dim);
  if (!col_major) {
    memcpy(r, f, sizeof(float) * (size_t)n * dim);
  } else {
    for (int i = 0; i < n; ++i)
      for (int k = 0; k < dim; ++k) r[(size_t)i * dim + k] = f[(size_t)k * n + i];
  }
  return r;
}

static float inv_norm(const float* x, int dim) {
  float acc = 0.f;
  for (int k = 0; k < dim; ++k) acc = fmaf(x[k], x[k], acc);
  if (acc == 0.f) return 0.f;
  return 1.0f / sqrtf(acc);
}

static float ssd_unit(const float* a, float inva, const float* b, float invb, int dim) {
  float acc = 0.f;
  for (int k = 0; k < dim; ++k) acc = fmaf(a[k], b[k], acc);
  float key = acc * invb;
  float c = key * inva;
  float s = fmaf(-2.0f, c, 2.0f);
  return s < 0.f ? 0.f : s;
}

/* The scores of row a against all columns.  Each dot product is the same sequential fmaf chain as in ssd_unit(); with
 * AVX2 + FMA, 32 columns advance together (four vectors of eight lanes: one fused multiply-add per lane and k, in the
 * same order), which hides the latency of the chain without touching a single rounding.  Bt is B in blocks of eight
 * columns: Bt[(j / 8) * dim * 8 + k * 8 + j % 8]. */
static void scores_row(const float* a, float inva, const float* B, const float* Bt, const float* invb, int n2, int dim, float* s) {
  int j = 0;
#if defined(__AVX2__) && defined(__FMA__)
  const __m256 m2 = _mm256_set1_ps(-2.0f), two = _mm256_set1_ps(2.0f), zero = _mm256_setzero_ps(), via = _mm256_set1_ps(inva);
  for (; j + 32 <= n2; j += 32) {
    const float* p = Bt + (size_t)(j / 8) * dim * 8;
    __m256 c0 = zero, c1 = zero, c2 = zero, c3 = zero;
    for (int k = 0; k < dim; ++k) {
      const __m256 va = _mm256_set1_ps(a[k]);
      c0 = _mm256_fmadd_ps(va, _mm256_loadu_ps(p + (size_t)k * 8), c0);
      c1 = _mm256_fmadd_ps(va, _mm256_loadu_ps(p + (size_t)dim * 8 + (size_t)k * 8), c1);
      c2 = _mm256_fmadd_ps(va, _mm256_loadu_ps(p + (size_t)dim * 16 + (size_t)k * 8), c2);
      c3 = _mm256_fmadd_ps(va, _mm256_loadu_ps(p + (size_t)dim * 24 + (size_t)k * 8), c3);
    }
    __m256 acc[4] = {c0, c1, c2, c3};
    for (int q = 0; q < 4; ++q) {
      const __m256 key = _mm256_mul_ps(acc[q], _mm256_loadu_ps(invb + j + 8 * q));
      const __m256 c = _mm256_mul_ps(key, via);
      const __m256 sc = _mm256_fmadd_ps(m2, c, two);
      /* s < 0 ? 0 : s  (a NaN stays a NaN, as in the scalar form) */
      _mm256_storeu_ps(s + j + 8 * q, _mm256_blendv_ps(sc, zero, _mm256_cmp_ps(sc, zero, _CMP_LT_OQ)));
    }
  }
#endif
  (void)Bt;
  for (; j < n2; ++j) s[j] = ssd_unit(a, inva, B + (size_t)j * dim, invb[j], dim);
}

static void top2_rows(const float* A, int n1, const float* B, int n2, int dim,
                      uint32_t* j1, float* s1, float* s2) {
  float* invb = (float*)malloc(sizeof(float) * (size_t)(n2 > 0 ? n2 : 1));
  for (int j = 0; j < n2; ++j) invb[j] = inv_norm(B + (size_t)j * dim, dim);
  const int nb = n2 / 8;
  float* Bt = (float*)malloc(sizeof(float) * (size_t)(nb > 0 ? nb : 1) * dim * 8);
  for (int j = 0; j < nb * 8; ++j)
    for (int k = 0; k < dim; ++k) Bt[(size_t)(j / 8) * dim * 8 + (size_t)k * 8 + (j % 8)] = B[(size_t)j * dim + k];
  float* s = (float*)malloc(sizeof(float) * (size_t)(n2 > 0 ? n2 : 1));
  for (int i = 0; i < n1; ++i) {
    const float* a = A + (size_t)i * dim;
    float inva = inv_norm(a, dim);
    float b1 = INFINITY, b2 = INFINITY;
    uint32_t bj = UINT32_MAX;
    scores_row(a, inva, B, Bt, invb, n2, dim, s);
    for (int j = 0; j < n2; ++j) {
      const float sj = s[j];
      if (sj < b1) { b2 = b1; b1 = sj; bj = (uint32_t)j; }
      else if (sj < b2) { b2 = sj; }
    }
    j1[i] = bj; s1[i] = b1; s2[i] = b2;
  }
  free(s); free(Bt); free(invb);
}

void vo_oracle_match_top2(const float* f1, int n1, const float* f2, int n2, int dim,
                          int col_major, uint32_t* j1, float* s1, float* s2) {
  float* A = to_row_major(f1, n1, dim, col_major);
  float* B = to_row_major(f2, n2, dim, col_major);
  top2_rows(A, n1, B, n2, dim, j1, s1, s2);
  free(A); free(B);
}

int vo_oracle_match(const float* f1, int n1, const float* f2, int n2, int dim,
                    int col_major, const vo_oracle_match_opts* opts,
                    uint32_t* idx1, uint32_t* idx2, float* metric) {
  vo_oracle_match_opts o = {1.0f, 0.6f, 0, 0};
  if (opts) o = *opts;
  if (n1 <= 0 || n2 <= 0) return 0;
  float* A = to_row_major(f1, n1, dim, col_major);
  float* B = to_row_major(f2, n2, dim, col_major);
  uint32_t* j1 = (uint32_t*)malloc(sizeof(uint32_t) * n1);
  float* s1 = (float*)malloc(sizeof(float) * n1);
  float* s2 = (float*)malloc(sizeof(float) * n1);
  top2_rows(A, n1, B, n2, dim, j1, s1, s2);
  uint32_t* back = NULL;
  if (o.unique) {  /* forward-backward consistency: j1's own nearest row must be i */
    back = (uint32_t*)malloc(sizeof(uint32_t) * n2);
    float* t1 = (float*)malloc(sizeof(float) * n2);
    float* t2 = (float*)malloc(sizeof(float) * n2);
    top2_rows(B, n2, A, n1, dim, back, t1, t2);
    free(t1); free(t2);
  }
  const float thr = o.match_threshold * 0.04f;
  int p = 0;
  for (int i = 0; i < n1; ++i) {
    if (!(s1[i] <= thr)) continue;
    if (n2 >= 2) {
      float ratio = (s2[i] < 1e-6f) ? 1.0f : s1[i] / s2[i];
      if (!(ratio <= o.max_ratio)) continue;
    }
    if (o.unique && back[j1[i]] != (uint32_t)i) continue;
    idx1[p] = (uint32_t)i + (uint32_t)o.index_base;
    idx2[p] = j1[i] + (uint32_t)o.index_base;
    if (metric) metric[p] = s1[i];
    ++p;
  }
  free(A); free(B); free(j1); free(s1); free(s2); free(back);
  return p;
}
