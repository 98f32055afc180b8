/*
 * vo_oracle.h -- CPU oracle of the stereo-VO hot path (TEST INFRASTRUCTURE ONLY).
 *
 * This is a single-threaded plain-C restatement of the toolbox calls that the
 * reference (ivario123/r7020e-visual-odometry, a MATLAB script) makes on its
 * per-frame hot path:
 *
 *   detectSIFTFeatures + extractFeatures("Method","SIFT")   VO.m:79-84
 *   matchFeatures (all defaults)                            VO.m:87,283,293,311,323
 *   triangulate                                             VO.m:114-115, CreateLandmarksFromFeatures.m:7
 *   estworldpose (P3P + MSAC)                               VO.m:123-127
 *
 * PARITY STATUS: "parity unpinned" by the reference -- the arithmetic lives in the
 * closed-source MathWorks Computer Vision Toolbox (version unpinned, >= R2022b
 * inferred), MATLAB is not installed here, and the reference ships no tests,
 * golden vectors or saved outputs.  The oracle is therefore pinned against the
 * independent implementations that ARE available offline (OpenCV 4.13 cv2.SIFT /
 * cv2.triangulatePoints / cv2.solveP3P, NumPy float64 brute force); the vectors
 * and the script that generated them are under tests/golden/.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
 * reference legs may load this library.  The product (libvo_b200.so) never does.
 */
#ifndef VO_ORACLE_H
#define VO_ORACLE_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

/* ------------------------------------------------------------------ SIFT -- */
typedef struct {
  float x, y;        /* 0-based pixel position in the input image (OpenCV pt)   */
  float size;        /* OpenCV KeyPoint::size (diameter)                        */
  float angle;       /* degrees, OpenCV convention, [0,360)                     */
  float response;    /* |D(x^)| contrast                                        */
  int32_t octave;    /* packed OpenCV octave word: oct&255 | layer<<8 | xi<<16  */
} vo_oracle_kp;

typedef struct {
  int n_octave_layers;      /* 3      */
  float contrast_threshold; /* 0.04  (MATLAB ContrastThreshold 0.0133 * layers) */
  float edge_threshold;     /* 10     */
  float sigma;              /* 1.6    */
} vo_oracle_sift_opts;

/* img: rows x cols uint8, row-major with leading dimension ld.
 * Returns number of keypoints (<= capacity), or -1 if capacity was too small.
 * desc: capacity x 128 floats row-major (integer valued 0..255). */
int vo_oracle_sift(const uint8_t* img, int rows, int cols, int ld,
                   const vo_oracle_sift_opts* opts, int capacity,
                   vo_oracle_kp* kps, float* desc);

/* pieces exposed for unit tests */
void vo_oracle_gauss_kernel(float sigma, int* radius, float* taps /* >= 64 */);
void vo_oracle_blur(const float* src, float* dst, int rows, int cols, float sigma);
void vo_oracle_base_image(const uint8_t* img, int rows, int cols, int ld, float sigma,
                          float* base /* (2rows) x (2cols) */);
float vo_oracle_expf(float x);
float vo_oracle_atan2deg(float y, float x);

/* ----------------------------------------------------------------- match -- */
typedef struct {
  float match_threshold; /* percent, default 1.0  -> SSD <= 0.04              */
  float max_ratio;       /* default 0.6                                        */
  int unique;            /* default 0                                          */
  int index_base;        /* 0 (C) or 1 (MATLAB)                                */
} vo_oracle_match_opts;

/* f1: n1 x dim, f2: n2 x dim float32; col_major != 0 means MATLAB layout
 * (element (i,k) at f[k*n + i]).  Outputs idx1/idx2/metric need n1 entries. */
int vo_oracle_match(const float* f1, int n1, const float* f2, int n2, int dim,
                    int col_major, const vo_oracle_match_opts* opts,
                    uint32_t* idx1, uint32_t* idx2, float* metric);

/* exact per-row best/second-best (for kernel unit tests); j1 = UINT32_MAX if n2==0 */
void vo_oracle_match_top2(const float* f1, int n1, const float* f2, int n2, int dim,
                          int col_major, uint32_t* j1, float* s1, float* s2);

/* ----------------------------------------------------------- triangulate -- */
/* pts1/pts2: n x 2 doubles row-major (pixel coordinates as given -- the caller
 * decides 0/1-based, P must match).  P1,P2: 3x4 row-major.  xyz: n x 3. */
void vo_oracle_triangulate(const double* pts1, const double* pts2, int n,
                           const double* P1, const double* P2,
                           double* xyz, double* reproj_err, uint8_t* valid);

/* -------------------------------------------------------------- P3P MSAC -- */
typedef struct {
  int max_num_trials;          /* 1000 */
  double confidence;           /* 99 (percent) */
  double max_reproj_error;     /* 1 (pixel) */
  uint64_t seed;
  int adaptive;                /* 1: MSAC adaptive stopping; 0: run all trials */
} vo_oracle_p3p_opts;

/* img: n x 2 (row-major, pixels), world: n x 3, K = {fx, fy, cx, cy}.
 * A: 4x4 row-major camera->world pose (premultiply convention [R t; 0 1]).
 * status: 0 ok, 1 fewer than 4 points, 2 not enough inliers.  Returns status. */
int vo_oracle_p3p(const double* img, const double* world, int n, const double K[4],
                  const vo_oracle_p3p_opts* opts, double A[16], uint8_t* inliers,
                  int* n_inliers, int* best_trial, int* trials_run);

/* P3P minimal solver: three unit bearings f (3x3 row-major), three world points
 * X (3x3 row-major); up to 4 solutions R (row-major, world->camera), t.  Returns count. */
int vo_oracle_p3p_solve(const double f[9], const double X[9], double R[4][9], double t[4][3]);

void vo_oracle_philox4x32(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                          uint32_t k0, uint32_t k1, uint32_t out[4]);
void vo_oracle_sample4(uint64_t seed, uint32_t trial, uint32_t n, uint32_t idx[4]);

#ifdef __cplusplus
}
#endif
#endif
