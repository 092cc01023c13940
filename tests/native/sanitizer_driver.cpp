// Test driver (tests/test_sched_sanitizers.py): the C++ planners of csrc/sched.cpp under AddressSanitizer + UBSan over random
// geometries — capacity queries, exact and short buffers, 1F1B / inference / zero-bubble orders, the balanced partitioner.
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>
#include "dpipe.h"
int main() {
  unsigned s = 12345; auto rnd = [&](int lo, int hi) { s = s * 1664525u + 1013904223u; return lo + (int)((s >> 8) % (unsigned)(hi - lo + 1)); };
  long total = 0;
  for (int it = 0; it < 4000; ++it) {
    int S = rnd(1, 12), M = rnd(1, 40), st = rnd(0, S - 1);
    int need = dpipe_sched_train(M, S, st, nullptr, 0);
    if (need <= 0) { printf("train need %d\n", need); return 1; }
    std::vector<dpipe_instr> buf(need);
    if (dpipe_sched_train(M, S, st, buf.data(), need) != need) return 2;
    if (need > 1) {                       // a buffer that is too short must not be written past (ASAN watches the heap block)
      std::vector<dpipe_instr> small(need - 1);
      (void)dpipe_sched_train(M, S, st, small.data(), need - 1);
    }
    int ni = dpipe_sched_infer(M, S, st, nullptr, 0);
    std::vector<dpipe_instr> b2(ni > 0 ? ni : 1);
    if (ni > 0 && dpipe_sched_infer(M, S, st, b2.data(), ni) != ni) return 3;
    int infl = rnd(1, 2 * S + 2); if (infl > M) infl = M;
    std::vector<int> w(S); for (auto& x : w) x = rnd(1, 9);
    int tf = rnd(1, 30), tb = rnd(1, 40), tw = rnd(1, 30);
    int nz = dpipe_sched_zb_ex(M, S, st, tf, tb, tw, infl, (it & 1) ? w.data() : nullptr, nullptr, 0);
    if (nz <= 0) { printf("zb need %d (M=%d S=%d infl=%d)\n", nz, M, S, infl); return 4; }
    std::vector<dpipe_instr> b3(nz);
    if (dpipe_sched_zb_ex(M, S, st, tf, tb, tw, infl, (it & 1) ? w.data() : nullptr, b3.data(), nz) != nz) return 5;
    long long mk = dpipe_sched_zb_makespan_ex(M, S, tf, tb, tw, infl, (it & 1) ? w.data() : nullptr);
    if (mk <= 0) return 6;
    total += need + ni + nz;
    int n = rnd(1, 80), parts = rnd(1, n < 16 ? n : 16);
    std::vector<int64_t> wt(n); for (auto& x : wt) x = rnd(0, 1000000);
    std::vector<int> bounds(parts + 1);
    if (dpipe_partition_balanced(wt.data(), n, parts, bounds.data()) != 0) return 7;
    if (bounds[0] != 0 || bounds[parts] != n) return 8;
  }
  printf("ok %ld instructions\n", total);
  return 0;
}
