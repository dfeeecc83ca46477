// CPU test of acg::GroupSync (aphros_b200/csrc/cg_group.h): the thread barrier the in-process
// slab group uses instead of NCCL.  Built and run by tests/test_host.py.
#include <atomic>
#include <chrono>
#include <cstdio>
#include <thread>
#include <vector>

#include "cg_group.h"

namespace acg {
void AttachGroupSync(aphcg*, GroupSync*) {}
void SetLastError(const char*) {}
}  // namespace acg

int main() {
  using acg::GroupSync;
  // 1. rounds of the host-side all-reduce pattern of AllReduceInitial: write own slot, wait,
  //    sum in slab order, wait again before anybody overwrites
  {
    const int n = 6, rounds = 2000;
    GroupSync gs(n);
    std::vector<double> got(n, 0.0);
    std::atomic<int> bad{0};
    std::vector<std::thread> th;
    for (int q = 0; q < n; ++q)
      th.emplace_back([&, q] {
        for (int r = 0; r < rounds; ++r) {
          gs.red[q] = (double)(r * n + q);
          if (!gs.Wait()) bad++;
          double s = 0;
          for (int k = 0; k < n; ++k) s += gs.red[k];
          if (!gs.Wait()) bad++;
          const double want = (double)n * r * n + n * (n - 1) / 2.0;
          if (s != want) bad++;
          got[q] = s;
        }
      });
    for (auto& t : th) t.join();
    if (bad) {
      printf("FAIL all-reduce rounds: %d mismatches\n", bad.load());
      return 1;
    }
  }
  // 2. Abort releases the waiters, and the group stays aborted
  {
    const int n = 4;
    GroupSync gs(n);
    std::atomic<int> released{0};
    std::vector<std::thread> th;
    for (int q = 0; q < n - 1; ++q)
      th.emplace_back([&] {
        if (!gs.Wait()) released++;
      });
    std::this_thread::sleep_for(std::chrono::milliseconds(50));
    gs.Abort();  // the slab that failed never reaches the barrier
    for (auto& t : th) t.join();
    if (released != n - 1 || !gs.aborted() || gs.Wait()) {
      printf("FAIL abort: released=%d\n", released.load());
      return 1;
    }
  }
  printf("OK\n");
  return 0;
}
