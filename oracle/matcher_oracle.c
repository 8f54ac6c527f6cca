/*
 * matcher_oracle.c -- CPU restatement of the secondary path:
 *   BruteForceFeatureMatcher::MatchImagePair  src/theia/matching/brute_force_feature_matcher.cc:49-117
 *   L2::operator()                            src/theia/matching/distance.h:48-57   (float squared distance)
 *   CompareFeaturesByDistance / top-2         src/theia/matching/indexed_feature_match.h:40-57
 *   IntersectMatches                          src/theia/matching/feature_matcher_utils.cc:48-71
 * TEST INFRASTRUCTURE (the checker for the round-2 CUDA matcher); nothing under theiasfm_b200/ uses it.
 * Pinned against the reference's own tests: brute_force_feature_matcher_test.cc:54-181,
 * feature_matcher_utils_test.cc:44-54, distance_test.cc:53-77 -- see tests/test_matcher_oracle.py.
 *
 * Where the reference leaves the result unspecified -- std::partial_sort among EQUAL distances, and the summation
 * order of Eigen's squaredNorm -- this restatement fixes: ties keep the lower feature index, distances are summed
 * left to right in float.  Parity tests must therefore treat exact ties / last-ulp distance differences as
 * "borderline" (SURVEY 8c), not as mismatches.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct { int32_t feature1_ind, feature2_ind; float distance; } indexed_feature_match;

typedef struct {
  int32_t keep_only_symmetric_matches; /* true  (feature_matcher_options.h:50) */
  int32_t use_lowes_ratio;             /* true  */
  float lowes_ratio;                   /* 0.8f  */
  int32_t min_num_feature_matches;     /* 30    */
} matcher_options;

void matcher_options_init(matcher_options* o) {
  o->keep_only_symmetric_matches = 1; o->use_lowes_ratio = 1; o->lowes_ratio = 0.8f; o->min_num_feature_matches = 30;
}

/* (a - b).squaredNorm() in float */
float matcher_l2(const float* a, const float* b, int dim) {
  float s = 0.0f;
  for (int k = 0; k < dim; ++k) { const float d = a[k] - b[k]; s += d * d; }
  return s;
}

/* One direction: for each descriptor of A the best match in B, kept if the (squared) ratio test passes.
 * Returns the number of matches written. */
static int match_one_way(const float* A, int nA, const float* B, int nB, int dim, const matcher_options* o, indexed_feature_match* out) {
  const float sq_lowes_ratio_f = o->lowes_ratio * o->lowes_ratio; /* :58-59: float * float, rounded to float, then widened */
  const double sq_lowes_ratio = (double)sq_lowes_ratio_f;
  int n = 0;
  for (int i = 0; i < nA; ++i) {
    float best = 0, second = 0; int best_j = -1, second_j = -1;
    for (int j = 0; j < nB; ++j) {
      const float d = matcher_l2(A + (size_t)i * dim, B + (size_t)j * dim, dim);
      if (best_j < 0 || d < best) { second = best; second_j = best_j; best = d; best_j = j; }
      else if (second_j < 0 || d < second) { second = d; second_j = j; }
    }
    if (best_j < 0) continue;
    /* :78-81; with a single candidate the reference reads past the valid range (undefined); we keep the match */
    if (!o->use_lowes_ratio || second_j < 0 || (double)best < sq_lowes_ratio * (double)second) {
      out[n].feature1_ind = i; out[n].feature2_ind = best_j; out[n].distance = best; ++n;
    }
  }
  return n;
}

/* IntersectMatches: keep forward (i -> j) only if the backward list holds (j -> i). Returns the new count. */
int matcher_intersect(const indexed_feature_match* backwards, int n_back, indexed_feature_match* forward, int n_fwd, int n2) {
  int32_t* map = (int32_t*)malloc(sizeof(int32_t) * (size_t)(n2 > 0 ? n2 : 1));
  for (int j = 0; j < n2; ++j) map[j] = -1;
  for (int k = 0; k < n_back; ++k) if (backwards[k].feature1_ind >= 0 && backwards[k].feature1_ind < n2) map[backwards[k].feature1_ind] = backwards[k].feature2_ind;
  int n = 0;
  for (int k = 0; k < n_fwd; ++k) {
    const int j = forward[k].feature2_ind;
    if (j >= 0 && j < n2 && map[j] == forward[k].feature1_ind) forward[n++] = forward[k];
  }
  free(map);
  return n;
}

/* MatchImagePair: returns 1 (and *n_matches) when the pair has enough matches, 0 otherwise (early exits included). */
int matcher_match_image_pair(const float* desc1, int n1, const float* desc2, int n2, int dim, const matcher_options* o,
                             indexed_feature_match* matches /* capacity n1 */, int* n_matches) {
  int n = match_one_way(desc1, n1, desc2, n2, dim, o, matches);
  *n_matches = n;
  if (n < o->min_num_feature_matches) return 0; /* :84-86 */
  if (o->keep_only_symmetric_matches) {        /* :89-113 */
    indexed_feature_match* rev = (indexed_feature_match*)malloc(sizeof(indexed_feature_match) * (size_t)(n2 > 0 ? n2 : 1));
    const int nr = match_one_way(desc2, n2, desc1, n1, dim, o, rev);
    n = matcher_intersect(rev, nr, matches, n, n2);
    free(rev);
    *n_matches = n;
  }
  return n >= o->min_num_feature_matches; /* :116 */
}
