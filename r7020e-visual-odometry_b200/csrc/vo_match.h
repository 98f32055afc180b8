// vo_match.h -- internal interface of the batched, device-count-driven match core (vo_match.cu),
// shared with the on-device frame loop (vo_frames.cu).
#pragma once
#include "vo_internal.h"

namespace vo {

// One side of a batch of match problems.  Problem p uses feature rows
//   row(i) = base + p*prob_stride + (gather ? gather[p*gather_stride + i] : i) * dim
// for i < min(count[p*count_stride], cap).
struct MatchOperand {
  const float* base = nullptr;
  size_t prob_stride = 0;
  const uint32_t* gather = nullptr;
  size_t gather_stride = 0;
  const int* count = nullptr;
  int count_stride = 0;
  int cap = 0;
  int col_major = 0;   // single problem only: element (i,k) at base[k*ld + i]
  int ld = 0;
  int integer_rows = 0;   // the caller vouches that every row holds integers 0..255 (descriptors written by sift_descriptor_kernel):
                          // with both sides vouched for, the general-float kernels are not even launched
  int prepared = 0;    // B side only: its u8 rows / 1/||row|| / max were written by an earlier call under the same tag
                       // (vo_landmarks_prepare: integer-valued rows) and are still there; base may be null
};

struct MatchTop2 {       // device results, [n_prob][cap_a_pad]
  uint32_t* j1; float* s1; float* s2; int row_stride;
  int* ctl;              // ctl[0] = non-integer flag, ctl[1] = rows sent to the exact scan
};

// Score bound for matchFeatures-style consumers: rows are only kept when s1 <= thr_score and
// s1/s2 <= max_ratio, so columns scoring above thr_score/max_ratio never matter.  key_floor is that
// bound as a cosine (0 = no bound: exact top-2 for every row).
struct MatchFilter { float key_floor; float thr_score; float max_ratio; };
MatchFilter make_match_filter(const vo_match_opts& o);

// prep + tcgen05 GEMM top-3 + finalize + exact row scan.  No host synchronisation.
// filter == nullptr: (j1, s1, s2) are the exact top-2 of every row.  With a filter, rows the filter
// rejects may carry s1 = inf / j1 = 0xFFFFFFFF and s2 may be a lower bound; match_batch_select with
// the same options returns exactly the pairs of the unfiltered run.
int match_batch_top2(vo_ctx* ctx, const MatchOperand& A, const MatchOperand& B, int n_prob, int dim,
                     const char* tag, float* dbg_c, cudaStream_t st, MatchTop2* out,
                     const MatchFilter* filter = nullptr);

// threshold / ratio tests + ordered compaction, one block per problem.
// idx1/idx2/metric: [n_prob][out_stride]; n_pairs[p*np_stride].  back (Unique): the top-2 of the reversed problems
// (match_batch_top2(B, A) under a different tag); a row is kept only if its landmark's nearest query row is that row.
int match_batch_select(vo_ctx* ctx, const MatchTop2& t, const MatchOperand& A, const MatchOperand& B, int n_prob,
                       const vo_match_opts& o, uint32_t* idx1, uint32_t* idx2, float* metric, int out_stride,
                       int* n_pairs, int np_stride, cudaStream_t st, const MatchTop2* back = nullptr);

void fill_match_opts(const vo_match_opts* in, vo_match_opts* o);
}  // namespace vo
