// tbm_top2.h -- running top-2 (nearest / second-nearest squared distance) with the reference's tie semantics, shared by
// the CUDA kernel (tbm_matcher.cu) and the CPU unit test (tests/host_top2.cc).
//
// Reference semantics (brute_force_feature_matcher.cc:64-76: distances to all candidates, std::partial_sort of the
// two smallest, CompareFeaturesByDistance = strict "<"), made deterministic the way oracle/matcher_oracle.c does:
// scanning candidates in ascending index, best = first occurrence of the minimum value; second = the smallest value of
// the remaining multiset (equal to best when the minimum occurs twice).
#pragma once

#ifndef TBM_HD
#ifdef __CUDACC__
#define TBM_HD __host__ __device__ __forceinline__
#else
#define TBM_HD inline
#endif
#endif

namespace tbm {

struct Top2 {
  float bd;  // best (smallest) distance
  int bj;    // its index, -1 = empty
  float sd;  // second smallest distance
  int has2;  // at least two candidates seen
};

TBM_HD void top2_init(Top2& t) { t.bd = 0.0f; t.bj = -1; t.sd = 0.0f; t.has2 = 0; }

// Candidates must arrive in ascending index within one scanner: strict "<" keeps the lower index on ties.
TBM_HD void top2_push(Top2& t, float d, int j) {
  if (t.bj < 0 || d < t.bd) { t.sd = t.bd; t.has2 = t.bj >= 0; t.bd = d; t.bj = j; }
  else if (!t.has2 || d < t.sd) { t.sd = d; t.has2 = 1; }
}

// Merge the summary of another scanner (disjoint candidate set) into m: best = lexicographic minimum of (distance, index),
// second = smallest value among {the losing best, both seconds}.
TBM_HD void top2_merge(Top2& m, const Top2& o) {
  if (o.bj < 0) return;
  if (m.bj < 0) { m = o; return; }
  const bool o_wins = o.bd < m.bd || (o.bd == m.bd && o.bj < m.bj);
  float sd = o_wins ? m.bd : o.bd;
  if (m.has2 && m.sd < sd) sd = m.sd;
  if (o.has2 && o.sd < sd) sd = o.sd;
  if (o_wins) { m.bd = o.bd; m.bj = o.bj; }
  m.sd = sd;
  m.has2 = 1;
}

}  // namespace tbm
