// In-process slab group: what one process driving several z-slabs (aphcg_group_*,
// include/aphcg.h) uses instead of NCCL for the few host-ordered steps of a solve.
// Host-only; shared by aphcg.cu (the per-slab driver) and aphcg_group.cu (the workers).
#pragma once

#include <condition_variable>
#include <mutex>

#include "cg_types.h"

struct aphcg;

namespace acg {

// Reusable barrier over the slab threads of one group, with an abort path: a slab
// whose step failed calls Abort() so that its peers return instead of waiting forever.
class GroupSync {
 public:
  explicit GroupSync(int n) : n_(n) {}
  // false: the group was aborted (now or earlier)
  bool Wait() {
    std::unique_lock<std::mutex> lk(mu_);
    if (aborted_) return false;
    const unsigned long long gen = gen_;
    if (++count_ == n_) {
      count_ = 0;
      ++gen_;
      cv_.notify_all();
      return true;
    }
    cv_.wait(lk, [&] { return gen_ != gen || aborted_; });
    return !aborted_;
  }
  void Abort() {
    std::lock_guard<std::mutex> lk(mu_);
    aborted_ = true;
    cv_.notify_all();
  }
  bool aborted() {
    std::lock_guard<std::mutex> lk(mu_);
    return aborted_;
  }
  int size() const { return n_; }
  // scratch of the host-side scalar all-reduce (one value per slab, summed in slab order)
  double red[kMaxRanks] = {};
  double red2[kMaxRanks] = {};

 private:
  std::mutex mu_;
  std::condition_variable cv_;
  int n_, count_ = 0;
  unsigned long long gen_ = 0;
  bool aborted_ = false;
};

// Internal (not exported through the C ABI): puts a handle created with nranks > 1 under
// an in-process group.  From then on its cross-slab barriers and the once-per-solve
// all-reduce go through `gs` (stream synchronize + thread barrier), the per-iteration
// scalars always through the peer mailboxes, and NCCL is never loaded.
void AttachGroupSync(aphcg* h, GroupSync* gs);
// Sets the calling thread's aphcg_last_error() text.
void SetLastError(const char* msg);

}  // namespace acg
