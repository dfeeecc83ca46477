// Device-side building blocks shared by the CG kernels.
#pragma once

#include <cuda_runtime.h>

#include "cg_types.h"

namespace acg {

// ---- vector-of-VX-doubles helpers (VX = 2 -> 128-bit accesses) ---------------
template <int VX>
struct Vec;
template <>
struct Vec<1> {
  double v[1];
};
template <>
struct alignas(16) Vec<2> {
  double v[2];
};

template <int VX>
__device__ __forceinline__ Vec<VX> ldv(const double* p) {
  return *reinterpret_cast<const Vec<VX>*>(p);
}
// streaming (read-once) load: do not keep the line in L1
template <int VX>
__device__ __forceinline__ Vec<VX> ldv_stream(const double* p) {
  Vec<VX> r;
  if constexpr (VX == 2) {
    double2 t = __ldcs(reinterpret_cast<const double2*>(p));
    r.v[0] = t.x;
    r.v[1] = t.y;
  } else {
    r.v[0] = __ldcs(p);
  }
  return r;
}
template <int VX>
__device__ __forceinline__ void stv(double* p, const Vec<VX>& x) {
  *reinterpret_cast<Vec<VX>*>(p) = x;
}
template <int VX>
__device__ __forceinline__ void stv_stream(double* p, const Vec<VX>& x) {
  if constexpr (VX == 2) {
    __stcs(reinterpret_cast<double2*>(p), make_double2(x.v[0], x.v[1]));
  } else {
    __stcs(p, x.v[0]);
  }
}

// ---- deterministic reductions -------------------------------------------------
// Fixed-shape trees only (no floating-point atomics): butterfly inside a warp,
// then over the warps of a block, then one block sums the per-block slots in a
// fixed order.  The result depends on the launch geometry but never on timing.
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// All threads of the block must call. Result valid in every thread.
template <bool kMax>
__device__ __forceinline__ double block_reduce(double v, double* sm /*[32]*/) {
  const int tid = threadIdx.x + blockDim.x * (threadIdx.y + blockDim.y * threadIdx.z);
  const int nthr = blockDim.x * blockDim.y * blockDim.z;
  const int lane = tid & 31, warp = tid >> 5, nwarp = (nthr + 31) >> 5;
  v = kMax ? warp_max(v) : warp_sum(v);
  __syncthreads();  // sm may still be read by a previous call
  if (lane == 0) sm[warp] = v;
  __syncthreads();
  double t = (lane < nwarp) ? sm[lane] : (kMax ? 0.0 : 0.0);
  t = kMax ? warp_max(t) : warp_sum(t);
  return t;
}

// Ticket: returns true in every thread of the block that arrives last.
__device__ __forceinline__ bool last_block(unsigned* counter, unsigned nblocks, int* sm_flag) {
  const int tid = threadIdx.x + blockDim.x * (threadIdx.y + blockDim.y * threadIdx.z);
  __syncthreads();
  if (tid == 0) {
    __threadfence();  // this block's slot is visible before the ticket
    const unsigned t = atomicAdd(counter, 1u);
    *sm_flag = (t == nblocks - 1);
    if (t == nblocks - 1) {
      *counter = 0;  // re-arm for the next launch
      __threadfence();
    }
  }
  __syncthreads();
  return *sm_flag != 0;
}

// Sum (or max) of the per-block slots, fixed order; call from one whole block.
template <bool kMax>
__device__ __forceinline__ double reduce_slots(const double* slots, unsigned n, double* sm) {
  const int tid = threadIdx.x + blockDim.x * (threadIdx.y + blockDim.y * threadIdx.z);
  const int nthr = blockDim.x * blockDim.y * blockDim.z;
  double acc = 0.0;
  for (unsigned i = tid; i < n; i += nthr) {
    const double s = __ldcg(slots + i);
    acc = kMax ? fmax(acc, s) : acc + s;
  }
  return block_reduce<kMax>(acc, sm);
}

// ---- peer-memory all-reduce (see cg_types.h) -------------------------------------
// phase 0: p.Ap (after the direction kernel), phase 1: r.r / max|r| (after the update)
__device__ __forceinline__ unsigned mail_flag(unsigned long long run, int iter, int phase) {
  // never 0 (fresh memory); differs between the two uses of a slot that can be confused: the
  // same slot two iterations earlier, and the last iterations of the previous run
  return 0x80000000u | (((unsigned)run & 0x7ffu) << 20) |
         ((2u * (unsigned)iter + (unsigned)phase + 1u) & 0xfffffu);
}
__device__ __forceinline__ unsigned long long global_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
__device__ __forceinline__ void mail_store(MailWord* p, double v, unsigned flag) {
  const unsigned long long b = (unsigned long long)__double_as_longlong(v);
  asm volatile("st.volatile.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(p),
               "r"((unsigned)b), "r"(flag), "r"((unsigned)(b >> 32)), "r"(flag)
               : "memory");
}
__device__ __forceinline__ MailWord mail_load(const MailWord* p) {
  MailWord w;
  asm volatile("ld.volatile.global.v4.u32 {%0, %1, %2, %3}, [%4];"
               : "=r"(w.lo), "=r"(w.f0), "=r"(w.hi), "=r"(w.f1)
               : "l"(p)
               : "memory");
  return w;
}
__device__ __forceinline__ double mail_value(const MailWord& w) {
  return __longlong_as_double((long long)(((unsigned long long)w.hi << 32) | w.lo));
}
__device__ __forceinline__ MailSlot* mail_slot(MailSlot* box, int phase, int parity, int src) {
  return box + ((phase * 2 + parity) * kMaxRanks + src);
}
// Called by every thread of the CTA that finished the local reduction of iteration `iter`.
__device__ __forceinline__ void mail_push(const Comm& cm, unsigned long long run, int iter,
                                          int phase, double sum, double mx, double sum2 = 0.0) {
  const int tid = threadIdx.x + blockDim.x * (threadIdx.y + blockDim.y * threadIdx.z);
  if (tid < cm.nranks) {
    MailSlot* s = mail_slot(cm.box[tid], phase, iter & 1, cm.rank);
    const unsigned flag = mail_flag(run, iter, phase);
    __threadfence_system();  // this kernel's halo stores (ordered before the ticket) first
    mail_store(&s->w[0], sum, flag);
    mail_store(&s->w[1], mx, flag);
    mail_store(&s->w[2], sum2, flag);
  }
}
// What a lane of the waiting warp has read of "its" source rank's slots.  Peeking at both
// parities costs nothing and needs no loop state, so a CTA can issue these loads together
// with its loads of the state instead of after them (one memory round trip less at every
// CTA start; in the steady state the data is already there).
struct MailPeek {
  MailWord w[2][3];
};
__device__ __forceinline__ MailPeek mail_peek(const Comm& cm, int phase) {
  MailPeek pk;
  const int lane = threadIdx.x & 31;
  if (lane < cm.nranks) {
#pragma unroll
    for (int par = 0; par < 2; ++par) {
      const MailSlot* s = mail_slot(cm.box[cm.rank], phase, par, lane);
#pragma unroll
      for (int j = 0; j < 3; ++j) pk.w[par][j] = mail_load(&s->w[j]);
    }
  }
  return pk;
}
__device__ __forceinline__ bool mail_complete(const MailWord* w, unsigned flag) {
  bool ok = true;
#pragma unroll
  for (int j = 0; j < 3; ++j) ok = ok && w[j].f0 == flag && w[j].f1 == flag;
  return ok;
}
// Called by one full warp.  Waits for every rank's contribution of (iter, phase) in the LOCAL
// mailbox (first looking at what mail_peek already fetched) and adds them in rank order.
// Returns false on timeout (a peer died or never ran: the wall-clock limit cm.timeout_ns keeps
// the GPU from hanging).
__device__ __forceinline__ bool mail_wait(const Comm& cm, unsigned long long run, int iter,
                                          int phase, const MailPeek* peek, double* sum, double* mx,
                                          double* sum2) {
  const int lane = threadIdx.x & 31;
  const unsigned flag = mail_flag(run, iter, phase);
  double vs = 0.0, vm = 0.0, v2 = 0.0;
  bool ok = true;
  if (lane < cm.nranks) {
    MailWord w[3];
    bool have = false;
    if (peek) {
#pragma unroll
      for (int j = 0; j < 3; ++j) w[j] = (iter & 1) ? peek->w[1][j] : peek->w[0][j];
      have = mail_complete(w, flag);
    }
    const MailSlot* s = mail_slot(cm.box[cm.rank], phase, iter & 1, lane);
    unsigned long long t0 = 0;
    unsigned spins = 0;
    while (!have) {
#pragma unroll
      for (int j = 0; j < 3; ++j) w[j] = mail_load(&s->w[j]);
      have = mail_complete(w, flag);
      if (!have && (++spins & 1023u) == 0) {
        const unsigned long long now = global_ns();
        if (t0 == 0) t0 = now;
        if (now - t0 > cm.timeout_ns) {
          ok = false;
          break;
        }
      }
    }
    vs = mail_value(w[0]);
    vm = mail_value(w[1]);
    v2 = mail_value(w[2]);
  }
  ok = __all_sync(0xffffffffu, ok);
  double ts = 0.0, tm = 0.0, t2 = 0.0;
  for (int q = 0; q < cm.nranks; ++q) {  // rank order: the same sum on every rank
    ts += __shfl_sync(0xffffffffu, vs, q);
    tm = fmax(tm, __shfl_sync(0xffffffffu, vm, q));
    t2 += __shfl_sync(0xffffffffu, v2, q);
  }
  *sum = ts;
  *mx = tm;
  *sum2 = t2;
  return ok;
}

// ---- scalar recurrences (reference stages "iter2", "iter3", "check") -----------
__device__ __forceinline__ double cg_alpha(const CgState* st) {
  return st->rr / (st->pAp + 1e-100);  // linear.ipp:84
}
__device__ __forceinline__ double cg_beta(const CgState* st) {
  return st->iter == 0 ? 0.0 : st->rr / (st->rr_prev + 1e-100);  // linear.ipp:98
}

// after the direction/SpMV kernel: global p.Ap is known
__device__ __forceinline__ void cg_finish_dir(CgState* st, double pap) {
  st->pAp = pap;
  st->pend_dir = 0;
}

// What the end of the update stage does to the loop scalars (linear.ipp:83-114): computed as
// a value so that a kernel can look one stage ahead without writing the shared state.
// rr_new: numerator of the next alpha/beta (sum r^2, or sum r.z when preconditioned);
// rnorm2: sum r^2, the residual norm.
struct CgStep {
  double alpha_prev, alpha_prev2, rr, rr_prev, rnorm2, max_r, residual;
  int iter, done;
};
__device__ __forceinline__ CgStep cg_step(const CgState* st, double rr_new, double max_r,
                                          double rnorm2) {
  CgStep n;
  n.alpha_prev2 = st->alpha_prev;
  n.alpha_prev = cg_alpha(st);
  n.rr_prev = st->rr;
  n.rr = rr_new;
  n.rnorm2 = rnorm2;
  n.max_r = max_r;
  n.residual = st->maxnorm ? max_r / st->cell_volume : sqrt(rnorm2 / st->cell_volume);
  n.iter = st->iter + 1;
  n.done = (n.iter >= st->miniter && (n.iter > st->maxiter || n.residual < st->tol)) ? 1 : 0;
  return n;
}
__device__ __forceinline__ void cg_commit(CgState* st, double* history, const CgStep& n) {
  st->alpha_prev2 = n.alpha_prev2;
  st->alpha_prev = n.alpha_prev;
  st->rr_prev = n.rr_prev;
  st->rr = n.rr;
  st->rnorm2 = n.rnorm2;
  st->max_r = n.max_r;
  st->residual = n.residual;
  if (n.iter - 1 < st->hist_cap) history[n.iter - 1] = n.residual;
  st->iter = n.iter;
  st->pend_upd = 0;
  if (n.done) st->done = 1;
}
// after the update kernel: global sum r^2 and max|r| are known.
// Advances the iteration and evaluates the exit rule (linear.ipp:102-113).
__device__ __forceinline__ void cg_finish_upd(CgState* st, double* history, double rr_new,
                                              double max_r, double rnorm2) {
  cg_commit(st, history, cg_step(st, rr_new, max_r, rnorm2));
}

// ---- the loop scalars as a consumer kernel sees them (Comm::wait_in_kernel) ------------
// Direction kernel: beta, the alphas of the deferred x updates and the iteration number, with
// the update stage of the previous iteration folded in when its all-reduce is still pending.
struct DirView {
  CgStep next;      // valid when pend
  double beta, alpha_prev, alpha_prev2;
  int iter, done, pend, error;
};
// All threads of the CTA call; `sv` is shared memory.  When the previous update stage is
// pending, warp 0 waits for every rank's r.r / max|r| in the local mailbox; the fence chain
// (peer's halo stores -> its ticket -> its mailbox flag -> this acquire) also makes the
// neighbours' boundary planes of r visible before this CTA reads its ghost planes, through
// TMA included (fence.proxy.async).
__device__ __forceinline__ void dir_view(const DevPtrs& d, DirView* sv) {
  const int tid = threadIdx.x + blockDim.x * (threadIdx.y + blockDim.y * threadIdx.z);
  if (tid < 32) {
    const CgState* st = d.st;
    MailPeek peek;
    if (d.cm.wait_in_kernel) peek = mail_peek(d.cm, 1);  // in flight together with the state
    DirView v;
    v.pend = st->pend_upd;
    v.error = 0;
    if (!v.pend) {
      v.iter = st->iter;
      v.done = st->done;
      v.beta = cg_beta(st);
      v.alpha_prev = st->alpha_prev;
      v.alpha_prev2 = st->alpha_prev2;
    } else {
      double sum, mx, sum2;
      const bool ok = mail_wait(d.cm, st->seq_base, st->iter, 1, &peek, &sum, &mx, &sum2);
      // the neighbours' boundary planes of r were stored before their flags
      if (d.cm.reader_fence) __threadfence_system();
      v.next = cg_step(st, sum, mx, st->precond ? sum2 : sum);
      if (!ok) {
        v.next.done = 1;
        v.error = 1;
      }
      v.iter = v.next.iter;
      v.done = v.next.done;
      v.beta = v.next.rr / (v.next.rr_prev + 1e-100);  // iter >= 1 here (linear.ipp:98)
      v.alpha_prev = v.next.alpha_prev;
      v.alpha_prev2 = v.next.alpha_prev2;
      asm volatile("fence.proxy.async;" ::: "memory");
    }
    if (tid == 0) *sv = v;
  }
  __syncthreads();
}
// Last CTA of the direction kernel, thread 0: commit what dir_view folded.
__device__ __forceinline__ void dir_commit(const DevPtrs& d, const DirView& v) {
  if (!v.pend) return;
  cg_commit(d.st, d.history, v.next);
  if (v.error) {
    d.st->error = 1;
    d.st->done = 1;
  }
}

// Last CTA of a direction kernel, all threads: `tot` is this rank's sum p.Ap.
template <bool kSingle>
__device__ __forceinline__ void dir_epilogue(const DevPtrs& d, const DirView& v, double tot) {
  const int tid = threadIdx.x + blockDim.x * (threadIdx.y + blockDim.y * threadIdx.z);
  CgState* st = d.st;
  if (tid == 0) {
    dir_commit(d, v);  // every CTA has read the old state by now (ticket)
    if (!v.done) {
      st->loc_sum = tot;
      if (kSingle) cg_finish_dir(st, tot);
      if (!kSingle && d.cm.wait_in_kernel) st->pend_dir = 1;
    }
  }
  if (!kSingle && d.cm.use_mail && !v.done) mail_push(d.cm, st->seq_base, v.iter, 0, tot, 0.0);
}

// Update kernel: alpha, with the direction stage's p.Ap folded in when pending.
struct UpdView {
  double alpha, pAp;
  int iter, pend, error;
};
__device__ __forceinline__ void upd_view(const DevPtrs& d, UpdView* sv) {
  const int tid = threadIdx.x + blockDim.x * (threadIdx.y + blockDim.y * threadIdx.z);
  if (tid < 32) {
    const CgState* st = d.st;
    MailPeek peek;
    if (d.cm.wait_in_kernel) peek = mail_peek(d.cm, 0);
    UpdView v;
    v.pend = st->pend_dir;
    v.iter = st->iter;
    v.error = 0;
    if (!v.pend) {
      v.pAp = st->pAp;
    } else {
      double sum, mx, sum2;
      if (!mail_wait(d.cm, st->seq_base, st->iter, 0, &peek, &sum, &mx, &sum2)) v.error = 1;
      v.pAp = sum;
    }
    v.alpha = st->rr / (v.pAp + 1e-100);  // linear.ipp:84
    if (tid == 0) *sv = v;
  }
  __syncthreads();
}

}  // namespace acg
