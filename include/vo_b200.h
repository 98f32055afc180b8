/*
 * vo_b200.h -- C ABI of libvo_b200.so: the B200 (sm_100a) implementation of the per-frame hot
 * path of ivario123/r7020e-visual-odometry (a MATLAB stereo visual-odometry script).
 *
 * The reference has no FFI of its own: its hot path is six MathWorks toolbox calls.  Each entry
 * point below replaces one call site; the MEX gateways (csrc/mex/vo_*_mex.cpp) and the Python
 * mirror (api.py) only marshal arguments onto these functions.
 *
 *   vo_sift         <- detectSIFTFeatures + extractFeatures(...,"Method","SIFT")   VO.m:79-84
 *   vo_match        <- matchFeatures(f1, f2)                      VO.m:87, 283, 293, 311, 323
 *   vo_triangulate  <- triangulate(p_l, p_r, p1, p2)    VO.m:114-115, CreateLandmarksFromFeatures.m:7
 *   vo_p3p          <- estworldpose(imagePoints, worldPoints, intrinsics)         VO.m:123-127
 *   vo_frames       <- one pass of the `for i = 1:n_frames` body over a batch     VO.m:64-232
 *
 * Conventions: plain pointers and sizes only.  Host-pointer entry points copy in/out and
 * synchronise before returning (what a MEX call needs).  `_dev` entry points take device pointers
 * and a cudaStream_t (as void*) and do not synchronise.  Every function returns 0 on success or
 * a negative vo_status; vo_last_error() gives the message (thread-local).  There is NO CPU
 * fallback: without a CUDA device every compute entry point fails with VO_ERR_CUDA.
 */
#ifndef VO_B200_H
#define VO_B200_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct vo_ctx vo_ctx;

enum vo_status {
  VO_OK = 0,
  VO_ERR_ARG = -1,      /* bad argument (null pointer, bad size/class)            */
  VO_ERR_CUDA = -2,     /* CUDA runtime/driver failure, or no device              */
  VO_ERR_CAPACITY = -3, /* caller-provided capacity too small                     */
  VO_ERR_STATE = -4     /* call sequence error                                    */
};

int vo_version(void);
const char* vo_last_error(void);
/* A context owns one stream and all device buffers of the calls made through it; contexts share
 * nothing, so several may be used concurrently from different host threads (one thread per context:
 * the batches-in-flight pattern of INTEGRATION.md).  The
 * contexts of a process may live on different devices (kernel attributes are tracked per device). */
int vo_ctx_create(int device, vo_ctx** ctx);
void vo_ctx_destroy(vo_ctx* ctx);
int vo_ctx_sync(vo_ctx* ctx);
/* stream used by the host-pointer entry points (cudaStream_t) */
void* vo_ctx_stream(vo_ctx* ctx);

/* Stage profiler (used by bench.py for the roofline line): when enabled, every kernel stage is
 * bracketed by CUDA events on its launching stream.  vo_profile_get returns, per stage name, the
 * summed device time, the number of launches and the ALGORITHMIC bytes / flops those launches
 * were asked to process (DESIGN.md "roofline accounting"). */
int vo_profile_enable(vo_ctx* ctx, int on);
/* number of CUDA kernels this context has launched since creation */
long long vo_kernel_launches(vo_ctx* ctx);
int vo_profile_count(vo_ctx* ctx);
int vo_profile_get(vo_ctx* ctx, int i, char* name, int name_cap, double* ms, long long* launches,
                   double* bytes, double* flops);

/* ------------------------------------------------------------------------------------ SIFT */
typedef struct {
  float contrast_threshold; /* MATLAB ContrastThreshold (per layer); <= 0 -> 0.04/3          */
  float edge_threshold;     /* MATLAB EdgeThreshold; <= 0 -> 10                               */
  int num_layers_in_octave; /* MATLAB NumLayersInOctave; <= 0 -> 3                            */
  float sigma;              /* MATLAB Sigma; <= 0 -> 1.6                                      */
  int index_base;           /* 0: OpenCV pt; 1: MATLAB Location (adds 1 to x and y)           */
} vo_sift_opts;

/* One keypoint record, OpenCV KeyPoint fields (MATLAB SIFTPoints: Scale = size/2,
 * Orientation = angle*pi/180, Metric = response, Octave/Layer unpacked from `octave`). */
typedef struct {
  float x, y;
  float size;
  float angle;
  float response;
  int32_t octave; /* (octave & 255) | layer << 8 | round((xi+0.5)*255) << 16                  */
} vo_keypoint;

/* img: rows x cols uint8.  col_major = 0: C layout, element (r,c) at img[r*ld + c];
 * col_major = 1: MATLAB layout, element (r,c) at img[c*ld + r].
 * kps[capacity], desc[capacity*128] (row-major, one 128-float descriptor per keypoint, integer
 * valued 0..255 like OpenCV/MATLAB SIFT).  *n_out = number of keypoints found; if it exceeds
 * capacity the first `capacity` (in output order: ascending x) are returned with VO_ERR_CAPACITY. */
int vo_sift(vo_ctx* ctx, const uint8_t* img, int rows, int cols, int ld, int col_major,
            const vo_sift_opts* opts, int capacity, vo_keypoint* kps, float* desc, int* n_out);

/* n_img images of identical size, contiguous (image stride rows*cols, row-major).
 * kps[n_img*capacity], desc[n_img*capacity*128], n_out[n_img]. */
int vo_sift_batch(vo_ctx* ctx, const uint8_t* imgs, int n_img, int rows, int cols,
                  const vo_sift_opts* opts, int capacity, vo_keypoint* kps, float* desc,
                  int* n_out);

/* Same for a stack of n_img images of identical size held as ONE array: col_major = 1 is MATLAB's H x W x N
 * uint8 array (element (r,c,k) at imgs[k*rows*cols + c*rows + r]), col_major = 0 equals vo_sift_batch.
 * desc_col_major = 1: image b's descriptors are returned as MATLAB holds an M x 128 single matrix (element
 * (i,k) at desc[b*capacity*128 + k*M + i], M = min(n_out[b], capacity)); the transpose runs on the device. */
int vo_sift_stack(vo_ctx* ctx, const uint8_t* imgs, int n_img, int rows, int cols, int col_major,
                  const vo_sift_opts* opts, int capacity, vo_keypoint* kps, float* desc,
                  int desc_col_major, int* n_out);

/* ----------------------------------------------------------------------------------- match */
typedef struct {
  float match_threshold; /* MATLAB MatchThreshold in percent; <= 0 -> 1.0 (SSD <= 0.04)      */
  float max_ratio;       /* MATLAB MaxRatio; <= 0 -> 0.6                                      */
  int unique;            /* MATLAB Unique (forward-backward consistency); default 0           */
  int index_base;        /* 0 or 1 (MATLAB)                                                   */
} vo_match_opts;

/* f1: n1 x dim, f2: n2 x dim float32 (1 <= dim <= 128: one descriptor is one 128-byte
 * tensor-core K block; VO_ERR_ARG otherwise).  col_major = 1: MATLAB layout (element (i,k)
 * at f[k*n + i]).  idx1/idx2/metric: capacity n1 entries; rows ascending in idx1.
 * indexPairs(:,1) = idx1, indexPairs(:,2) = idx2, matchMetric = metric (may be NULL). */
int vo_match(vo_ctx* ctx, const float* f1, int n1, const float* f2, int n2, int dim,
             int col_major, const vo_match_opts* opts, uint32_t* idx1, uint32_t* idx2,
             float* metric, int* n_pairs);

/* Per-row nearest / second nearest before thresholding (diagnostic + relocalisation shards):
 * j1[n1] (0-based, UINT32_MAX when n2 == 0), s1[n1], s2[n1] (INF when n2 < 2). */
int vo_match_top2(vo_ctx* ctx, const float* f1, int n1, const float* f2, int n2, int dim,
                  int col_major, uint32_t* j1, float* s1, float* s2);

/* Device-pointer variant: f1_dev/f2_dev row-major float32 in device memory; outputs device
 * arrays of n1 entries; n_pairs_dev device int.  Runs on `stream`, no synchronisation. */
int vo_match_dev(vo_ctx* ctx, const float* f1_dev, int n1, const float* f2_dev, int n2, int dim,
                 const vo_match_opts* opts, uint32_t* idx1_dev, uint32_t* idx2_dev,
                 float* metric_dev, int* n_pairs_dev, void* stream);
int vo_match_top2_dev(vo_ctx* ctx, const float* f1_dev, int n1, const float* f2_dev, int n2,
                      int dim, uint32_t* j1_dev, float* s1_dev, float* s2_dev, void* stream);

/* Relocalisation shard (BASELINE config 5): one 16-byte record per query row,
 * records_dev[n1] = {uint32 j1, float s1, float s2, uint32 keep}.  keep = 1 iff the row passes the
 * MatchThreshold and MaxRatio tests of `opts` (then j1/s1 are exactly matchFeatures' pair and metric;
 * rows with keep = 0 carry j1 = UINT32_MAX, and s2 may be a lower bound of the second-nearest score).
 * Query rows are independent, so a row-sharded run followed by an all-gather of the records equals
 * the unsharded result. */
int vo_match_best2_dev(vo_ctx* ctx, const float* f1_dev, int n1, const float* f2_dev, int n2, int dim,
                       const vo_match_opts* opts, void* records_dev, void* stream);

/* The same shard with the all-gather fused into the match epilogue (one process per GPU, NVLink / NVSwitch peer
 * stores): peer_bufs[r] is rank r's gathered-record buffer (n_total x 16 bytes, this rank's own one included), each
 * made with vo_peer_alloc by its owner and mapped here with vo_peer_open (CUDA IPC; the 64-byte handles travel over any
 * host channel, e.g. torch.distributed.all_gather_object).  The kernel writes record i of this shard to row
 * first_row + i of EVERY buffer; afterwards one barrier across the ranks (no data) makes all slices visible. */
int vo_match_best2_gather_dev(vo_ctx* ctx, const float* f1_dev, int n1, const float* f2_dev, int n2, int dim,
                              const vo_match_opts* opts, void* const* peer_bufs, int n_peers, size_t first_row,
                              void* stream);
/* A landmark set that stays on the device (the relocalisation map): converted ONCE into the match operand form
 * (128-byte u8 rows, 1/||row||; 134 MB per million landmarks instead of 512 MB of float rows, which may be freed
 * afterwards).  vo_match_best2_dev / vo_match_best2_gather_dev then take f2_dev = NULL with n2 = the prepared count and
 * skip the conversion on every call (n2 = 0 with a NULL pointer is the empty set, not the prepared one).  Rows must be integers 0..255 (SIFT descriptors), dim 128; synchronises once. */
int vo_landmarks_prepare(vo_ctx* ctx, const float* f2_dev, int n2, int dim, void* stream);
int vo_peer_alloc(vo_ctx* ctx, size_t bytes, void** dev_ptr, uint8_t handle[64]);
int vo_peer_open(vo_ctx* ctx, const uint8_t handle[64], void** dev_ptr);
int vo_peer_close(vo_ctx* ctx, void* dev_ptr);
int vo_peer_free(vo_ctx* ctx, void* dev_ptr);

/* counters of the last vo_match / vo_match_top2 call on this ctx (after synchronisation):
 * stats[0] = 1 if the exact-integer u8 path ran (0: split-bf16 general path),
 * stats[1] = rows re-evaluated by the exact FP32 row scan, stats[2] = GEMM kernel launches,
 * stats[3] = K extent of the GEMM (general path: 128 per bf16 term; one term with a score bound, three without). */
int vo_match_stats(vo_ctx* ctx, int stats[4]);

/* Debug/unit-test hook: raw tensor-core dot products C = f1 * f2^T (n1 x n2 float32, row-major
 * host output), computed by the same tcgen05 pipeline as vo_match. */
int vo_match_debug_gemm(vo_ctx* ctx, const float* f1, int n1, const float* f2, int n2, int dim,
                        float* c_out);

/* ----------------------------------------------------------------------------- triangulate */
/* pts1/pts2: n x 2 (is_double ? double : float), row-major (col_major = 1: MATLAB n x 2 column
 * major).  P1, P2: 3x4 row-major double.  xyz: n x 3 in the class of the points, same majorness.
 * reproj_err (n, same class) and valid (n bytes) may be NULL. */
int vo_triangulate(vo_ctx* ctx, const void* pts1, const void* pts2, int n, int is_double,
                   int col_major, const double P1[12], const double P2[12], void* xyz,
                   void* reproj_err, uint8_t* valid);

/* ------------------------------------------------------------------------------- P3P MSAC */
typedef struct {
  int max_num_trials;      /* MATLAB MaxNumTrials; <= 0 -> 1000                              */
  double confidence;       /* MATLAB Confidence (percent); <= 0 -> 99                         */
  double max_reproj_error; /* MATLAB MaxReprojectionError (pixels); <= 0 -> 1                 */
  uint64_t seed;           /* counter-based RNG seed (the reference never seeds MATLAB's RNG) */
  int adaptive;            /* < 0 -> 1 (MSAC adaptive trial bound); 0: run every trial        */
} vo_p3p_opts;

/* img: n x 2 double pixels, world: n x 3 double (col_major = 1: MATLAB layout), K = {fx,fy,cx,cy}.
 * A: 4x4 camera-to-world pose, row-major (col_major = 1: column-major, ready for rigidtform3d).
 * inliers: n bytes (may be NULL).  *status: 0 ok, 1 fewer than 4 points, 2 not enough inliers.
 * info (may be NULL): {n_inliers, best_trial, trials_run}. */
int vo_p3p(vo_ctx* ctx, const double* img, const double* world, int n, int col_major,
           const double K[4], const vo_p3p_opts* opts, double A[16], uint8_t* inliers,
           int* status, int info[3]);

/* ---------------------------------------------------------------------------- input staging */
/* Replaces imageDatastore / readimage (VO.m:16-17, 71-72) for KITTI odometry frames: 8-bit grayscale,
 * non-interlaced PNG.  Host code (the library's own inflate and PNG row filters; zlib only adjudicates
 * streams that decoder rejects); other PNG flavours are rejected.
 * vo_png_read_batch decodes n files with n_threads workers (<= 0: one per hardware thread) into
 * out[n][rows][cols] -- typically the pinned batch buffer handed to vo_frames. */
int vo_png_info(const uint8_t* file, size_t n_bytes, int* rows, int* cols, int* bit_depth, int* color_type);
int vo_png_decode_gray8(const uint8_t* file, size_t n_bytes, uint8_t* out, int ld, int rows, int cols);
int vo_png_read_batch(const char* const* paths, int n, int rows, int cols, uint8_t* out, int n_threads);
/* The inflate stage of the reader on its own: the zlib stream in[0, n_in) must inflate to exactly n_out
 * bytes with a matching Adler-32 (this is the library's own decoder, without the zlib second opinion
 * the PNG reader takes on failure). */
int vo_inflate_zlib(const uint8_t* in, size_t n_in, uint8_t* out, size_t n_out);

/* Device-side decode of the same files (DEFLATE + PNG row filters as CUDA kernels, one warp per image; the host only
 * gathers the IDAT payloads): n encoded files in host memory (or n paths) -> out_dev[n][rows][cols] in DEVICE memory,
 * e.g. the batch buffer handed to vo_frames_dev.  For hosts with few cores per GPU: a core decodes 400-700 KITTI
 * frames per second, one GPU consumes 9 k.  Runs on the context's stream and returns when the decode has finished;
 * a malformed file fails with VO_ERR_ARG (vo_last_error names the image and the reason). */
int vo_png_decode_batch_dev(vo_ctx* ctx, const uint8_t* const* files, const size_t* sizes, int n, int rows, int cols,
                            uint8_t* out_dev);
int vo_png_read_batch_dev(vo_ctx* ctx, const char* const* paths, int n, int rows, int cols, uint8_t* out_dev,
                          int n_threads);

/* ------------------------------------------------------------------------- frame pipeline */
typedef struct {
  vo_sift_opts sift;
  vo_match_opts match;
  vo_p3p_opts p3p;
  int max_keypoints;   /* per-image capacity; <= 0 -> 8192 */
  int first_frame;     /* absolute index of frame 0 of this batch: keys the MSAC random stream so a
                          frame gets the same samples however the sequence is cut into batches */
  int col_major;       /* 1: every image is MATLAB-ordered (element (r,c) at img[c*rows + r], i.e. an
                          H x W x N uint8 array as MATLAB holds it); transposed on the device        */
} vo_frames_opts;

/* One pass of the VO.m loop body over n_frames consecutive stereo frames held in host memory
 * (left/right: n_frames x rows x cols uint8, row-major).  Frame 0 only seeds the tracker
 * (VO.m:207-210); for frame i >= 1: SIFT x2, stereo match, find_remaining_points against frame
 * i-1 (VO.m:280-334), batched triangulation of the previous pair, P3P-MSAC.
 * rel_pose[16*i] = rel_pose.A of frame i (row-major; identity for i = 0), status[i] = estworldpose
 * status (0 ok, 1 too few points, 2 too few inliers),
 * counts[8*i..] = {N_L, N_R, K0 (stereo), K1, K2, K3, K4 (tracked), inliers}  (may be NULL).
 * opts->match.unique = 1 applies Unique (forward-backward consistency) to all five matchFeatures calls (VO.m never
 * sets it; every match then also runs in the reverse direction).
 * The pose chain pose = pose * rel_pose (VO.m:130) is left to the caller: it is sequential.
 * To stream a long sequence call with overlapping batches [i0-1, i0+B). */
int vo_frames(vo_ctx* ctx, const uint8_t* left, const uint8_t* right, int n_frames, int rows,
              int cols, const double P1[12], const double P2[12], const vo_frames_opts* opts,
              double* rel_pose, int* status, int* counts);
/* Same, with the images already resident in device memory (left_dev/right_dev: device pointers). */
int vo_frames_dev(vo_ctx* ctx, const uint8_t* left_dev, const uint8_t* right_dev, int n_frames,
                  int rows, int cols, const double P1[12], const double P2[12],
                  const vo_frames_opts* opts, double* rel_pose, int* status, int* counts);

/* Landmark map (VO.m:145-161, CreateLandmarksFromFeatures.m:1-21) for the frames of the most recent
 * vo_frames / vo_frames_dev call on this context, entirely on the device: selection of the stereo-matched features
 * that "did not exist in the previous frame" (the reference's x-OR-y comparison), every second one triangulated,
 * depth filter 0 <= z <= 80, transformPointsForward with the frame's world pose.
 * poses: n_frames x 16 doubles, row-major 4x4 world pose of every frame of the batch AFTER the caller's pose chain
 * (pose = pose * rel_pose, VO.m:130; the chain is sequential and stays with the caller).
 * landmarks: [n_frames][cap][3] doubles, rows[n_frames]: frame i contributes rows[i] rows (what the reference
 * appends at VO.m:160: zero rows for skipped / rejected features included, 0 rows for frame 0 and for frames whose
 * estworldpose failed).  VO_ERR_STATE without a preceding vo_frames call, VO_ERR_CAPACITY if cap is too small. */
int vo_frames_landmarks(vo_ctx* ctx, const double* poses, int n_frames, int cap, double* landmarks, int* rows);

/* CUDA graph for the frame loop.  With enable != 0, vo_frames / vo_frames_dev capture their launch
 * sequence (SIFT ... P3P and the result copies; shapes depend on capacities only) the second time a (batch shape, option
 * set) is seen and replay it afterwards: a call then uploads its images and per-call parameters and launches ONE graph.
 * Results are bit-identical to the plain path.  The environment variable VO_FRAMES_GRAPH=1 enables it for every new
 * context.  vo_frames_graph_state: 0 = off, 1 = on (nothing captured yet), 2 = a captured graph is being replayed,
 * -1 = capture failed on this driver and the context fell back to plain launches. */
int vo_frames_use_graph(vo_ctx* ctx, int enable);
int vo_frames_graph_state(vo_ctx* ctx);

#ifdef __cplusplus
}
#endif
#endif
