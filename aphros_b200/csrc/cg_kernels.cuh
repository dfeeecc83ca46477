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
__device__ __forceinline__ unsigned long long mail_seq(const CgState* st, int phase) {
  return st->seq_base + 2ull * (unsigned long long)st->iter + (unsigned long long)phase + 1ull;
}
// Called by every thread of the CTA that finished the local reduction.
__device__ __forceinline__ void mail_push(const Comm& cm, const CgState* st, int phase, double sum,
                                          double mx, double sum2 = 0.0) {
  const int tid = threadIdx.x + blockDim.x * (threadIdx.y + blockDim.y * threadIdx.z);
  if (tid < cm.nranks) {
    MailSlot* s = cm.box[tid] + ((phase * 2 + (st->iter & 1)) * kMaxRanks + cm.rank);
    *reinterpret_cast<volatile double*>(&s->sum) = sum;
    *reinterpret_cast<volatile double*>(&s->mx) = mx;
    *reinterpret_cast<volatile double*>(&s->sum2) = sum2;
    __threadfence_system();  // values (and this kernel's halo stores) before the flag
    *reinterpret_cast<volatile unsigned long long*>(&s->seq) = mail_seq(st, phase);
  }
}
// Called by one warp.  Returns false (and flags the error) on timeout.
__device__ __forceinline__ bool mail_wait(const Comm& cm, CgState* st, int phase, double* sum,
                                          double* mx, double* sum2) {
  const int lane = threadIdx.x & 31;
  const unsigned long long want = mail_seq(st, phase);
  double vs = 0.0, vm = 0.0, v2 = 0.0;
  bool ok = true;
  if (lane < cm.nranks) {
    MailSlot* s = cm.box[cm.rank] + ((phase * 2 + (st->iter & 1)) * kMaxRanks + lane);
    const long long t0 = clock64();
    while (*reinterpret_cast<volatile unsigned long long*>(&s->seq) != want) {
      if (clock64() - t0 > 8000000000ll) {  // ~4 s: a peer died; do not hang the GPU
        ok = false;
        break;
      }
    }
    __threadfence_system();
    vs = *reinterpret_cast<volatile double*>(&s->sum);
    vm = *reinterpret_cast<volatile double*>(&s->mx);
    v2 = *reinterpret_cast<volatile double*>(&s->sum2);
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
  if (!ok && lane == 0) {
    st->error = 1;
    st->done = 1;
  }
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
__device__ __forceinline__ void cg_finish_dir(CgState* st, double pap) { st->pAp = pap; }

// after the update kernel: global sum r^2 and max|r| are known.
// Advances the iteration and evaluates the exit rule (linear.ipp:102-113).
// rr_new: numerator of the next alpha/beta (sum r^2, or sum r.z when preconditioned);
// rnorm2: sum r^2, the residual norm.
__device__ __forceinline__ void cg_finish_upd(CgState* st, double* history, double rr_new,
                                              double max_r, double rnorm2) {
  st->alpha_prev2 = st->alpha_prev;
  st->alpha_prev = cg_alpha(st);
  st->rr_prev = st->rr;
  st->rr = rr_new;
  st->rnorm2 = rnorm2;
  st->max_r = max_r;
  const double res = st->maxnorm ? max_r / st->cell_volume : sqrt(rnorm2 / st->cell_volume);
  st->residual = res;
  const int it = st->iter + 1;
  if (it - 1 < st->hist_cap) history[it - 1] = res;
  st->iter = it;
  if (it >= st->miniter && (it > st->maxiter || res < st->tol)) st->done = 1;
}

}  // namespace acg
