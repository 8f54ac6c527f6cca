// Stress test of the host pack's worker pool (theiasfm_b200/csrc/tba_pack.h: PackPool, parallel_for): several caller threads at once
// (one gets the pool, the others fall back to their own threads -- the rank threads of tba_solve_multi do exactly that), loops of
// every size, nested use from inside a worker, results checked exactly.  Built and run by tests/test_pack_cpu.py (also under TSan
// by hand: g++ -fsanitize=thread).
#include <atomic>
#include <cstdio>
#include <numeric>
#include <thread>
#include <vector>

#include "../theiasfm_b200/csrc/tba_pack.h"

int main() {
  using tba::parallel_for;
  std::atomic<long long> bad(0);
  auto caller = [&](int id) {
    std::vector<long long> out(100000);
    for (int rep = 0; rep < 300; ++rep) {
      const int64_t n = 1 + (rep * 977 + id * 131) % 100000;
      const int T = 1 + (rep + id) % 24;
      std::fill(out.begin(), out.begin() + n, -1);
      parallel_for(n, T, [&](int64_t b, int64_t e, int) { for (int64_t i = b; i < e; ++i) out[i] = i * 3 + id; }, 64);
      for (int64_t i = 0; i < n; ++i) if (out[i] != i * 3 + id) { bad++; break; }
      if (rep % 50 == 0) {  // nested: a loop started from inside a parallel loop must not dead-lock
        std::atomic<long long> sum(0);
        parallel_for(8, 8, [&](int64_t b, int64_t e, int) {
          for (int64_t i = b; i < e; ++i) parallel_for(1000, 4, [&](int64_t bb, int64_t ee, int) { sum += ee - bb; }, 10);
        }, 1);
        if (sum.load() != 8000) bad++;
      }
    }
  };
  std::vector<std::thread> th;
  for (int id = 0; id < 6; ++id) th.emplace_back(caller, id);
  for (auto& t : th) t.join();
  std::printf(bad.load() == 0 ? "pack pool ok\n" : "pack pool FAILED\n");
  return bad.load() == 0 ? 0 : 1;
}
