// Host-side check of the compile-time MMA schedule of K2c (corr_ozaki.cu::make_pass_plan).  The plan source is
// pasted in by tests/test_host_logic.py (PLAN_SNIPPET), so the test always checks the code the kernel is built from.
#include <cassert>
#include <cstdio>
#define __host__
#define __device__
#include PLAN_SNIPPET
int main() {
  const int cfg[7][2] = {{2, 2}, {4, 2}, {6, 2}, {8, 2}, {3, 1}, {5, 1}, {7, 1}};
  for (int mode = 0; mode < 2; ++mode)
    for (auto& c : cfg) {
      const int nload = c[0], ng = c[1], gl = nload - ng;
      const PassPlan pl = make_pass_plan(nload, ng, mode);
      unsigned seen = 0;
      for (int k = 0; k < nload; ++k) {  // load order is a permutation, pos its inverse
        const int x = pl.ord[k];
        assert(x >= 0 && x < nload && !((seen >> x) & 1u));
        seen |= 1u << x;
        assert(pl.pos[x] == k);
      }
      unsigned have = 0, freed = 0, started = 0;
      int pi = 0, cnt[2][8] = {};
      for (int st = 0; st < pl.nsteps; ++st) {
        assert(!(pl.wait_mask[st] & have));  // every unit is waited for exactly once
        have |= pl.wait_mask[st];
        const int n_have = __builtin_popcount(have);
        for (int k = 0; k < n_have; ++k) assert((have >> pl.ord[k]) & 1u);  // waits follow the load order
        for (; pi < pl.step_end[st]; ++pi) {
          const int a = pl.prod_a[pi], b = pl.prod_b[pi], g = pl.prod_g[pi];
          assert(g >= 0 && g < ng && a + b == gl + g);               // product belongs to its significance group
          assert(((have >> a) & 1u) && ((have >> b) & 1u));          // both units have landed
          assert(!((freed >> a) & 1u) && !((freed >> b) & 1u));      // and were not handed back yet
          assert(pl.prod_first[pi] == !((started >> g) & 1u));       // accumulate flag
          started |= 1u << g;
          cnt[g][a]++;
        }
        assert(!(pl.free_mask[st] & freed) && !(pl.free_mask[st] & ~have));
        freed |= pl.free_mask[st];
      }
      assert(have == (1u << nload) - 1u && freed == have);           // every unit waited for and freed once
      for (int g = 0; g < ng; ++g)
        for (int t = 0; t <= gl + g; ++t) assert(cnt[g][t] == 1);    // every product issued exactly once
      assert(pi == pl.nprod);
    }
  std::puts("pass plan ok");
  return 0;
}
