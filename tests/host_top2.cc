// host_top2.cc -- CPU unit-test shim for theiasfm_b200/csrc/tbm_top2.h (the top-2 bookkeeping of the GPU matcher).
// emulate_kernel() reproduces k_nn2's scan order on the host: 8 scanners, scanner w takes candidates j with
// (j mod 32) mod 8 == w in ascending j, then the 8 summaries are merged in scanner order.
#include "../theiasfm_b200/csrc/tbm_top2.h"

extern "C" void emulate_kernel(const float* d, int n, int* best_j, float* best_d, float* second_d, int* has2) {
  tbm::Top2 t[8];
  for (int w = 0; w < 8; ++w) tbm::top2_init(t[w]);
  for (int j0 = 0; j0 < n; j0 += 32)
    for (int w = 0; w < 8; ++w)
      for (int jj = w; jj < 32; jj += 8) {
        const int j = j0 + jj;
        if (j >= n) break;
        tbm::top2_push(t[w], d[j], j);
      }
  tbm::Top2 m = t[0];
  for (int w = 1; w < 8; ++w) tbm::top2_merge(m, t[w]);
  *best_j = m.bj; *best_d = m.bd; *second_d = m.has2 ? m.sd : 0.0f; *has2 = m.has2;
}

// the reference semantics as oracle/matcher_oracle.c states them: one sequential scan
extern "C" void sequential(const float* d, int n, int* best_j, float* best_d, float* second_d, int* has2) {
  tbm::Top2 t;
  tbm::top2_init(t);
  for (int j = 0; j < n; ++j) tbm::top2_push(t, d[j], j);
  *best_j = t.bj; *best_d = t.bd; *second_d = t.has2 ? t.sd : 0.0f; *has2 = t.has2;
}
