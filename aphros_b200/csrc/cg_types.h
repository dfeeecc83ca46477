// Shared host/device types of the CG solver (libaphcg.so).
#pragma once

#include <cstdint>

namespace acg {

// Streaming multiprocessors of the one target, B200 (sm_100a): grids of the grid-stride and
// persistent kernels are sized in multiples of it.
constexpr int kNumSMs = 148;

// Ghost offset of inner cell i=0 inside a padded row: keeps inner rows 128-byte
// aligned so that pair (128-bit) accesses and TMA boxes start on line boundaries.
constexpr int kGhostX = 16;

// Geometry of one rank's z-slab.
//
// "compact" arrays (coefficients, rhs, u, Ap) hold inner cells only:
//   idx = i + j*cy + k*cz.
// "padded" arrays (r, p0, p1) carry one ghost layer on each of the six faces
// (what m.Comm(&f, direct_one) maintains in the reference,
// src/linear/linear.ipp:57,100):
//   idx = poff + i + j*py + k*pz,  i in [-1,nx], j in [-1,ny], k in [-1,nzl].
struct Geom {
  int nx, ny, nzl;
  int zc;               // planes marched by one CTA of the tiled (non-TMA) kernels
  int utx;              // update kernel: threads along x (power of two, 32..256)
  int per_x, per_y;     // periodic in x / y (wrap inside the slab)
  int64_t cy, cz, ncell;
  int64_t py, pz, poff, ptotal;
};

// ---- scalar all-reduce through peer memory ("mailboxes") -----------------------
// Every rank owns a small array of slots in its own HBM, mapped into every other
// rank's address space (CUDA IPC).  The CTA that finishes a reduction stores this
// rank's partial result straight into slot [phase][iteration parity][my rank] of
// EVERY rank's array (NVLink stores), value first, then -- after a system-scope
// fence -- a sequence number.  The consumer waits until all nranks sequence numbers
// of the current step have arrived in its LOCAL array and adds the values in rank
// order, so every rank forms bitwise the same sum without a collective call.
constexpr int kMaxRanks = 16;
// One double travels as 16 bytes: its two halves, each followed by a copy of a 32-bit flag
// that identifies (run, iteration, phase) -- the "low latency" idea of NCCL's LL protocol.  The
// writer stores the 16 bytes with one vector store; the reader needs no fence between flag and
// data because each 8-byte half {data, flag} is written and read as a unit: a half whose flag
// is the expected one carries the data of that very store.
struct alignas(16) MailWord {
  unsigned lo, f0, hi, f1;
};
struct alignas(64) MailSlot {
  MailWord w[3];  // sum, max, second sum
  unsigned pad[4];
};
constexpr int kMailSlots = 2 * 2 * kMaxRanks;  // [phase][parity][source rank]
struct Comm {
  MailSlot* box[kMaxRanks];  // box[q]: rank q's slot array (own memory for q == rank)
  int rank, nranks;
  int use_mail;  // 1: mailboxes; 0: an NCCL all-reduce on loc_sum/loc_max follows the kernel
  // 1: the kernel that CONSUMES a reduced scalar waits for the mailboxes itself (every CTA of
  //    k_update for p.Ap, every CTA of the direction kernel for r.r), so an iteration is two
  //    launches on any number of GPUs; 0: one-warp k_finish_* kernels do the waiting (NCCL
  //    reduction, or slabs that share one GPU: a grid of spinning CTAs would keep the peer
  //    slab's producer kernel off the SMs)
  int wait_in_kernel;
  // 1 (always, except for measurements: APHCG_WAIT_FENCE=0): system-scope fence between the
  // arrival of the neighbours' flags and the first read of this slab's ghost planes of r
  int reader_fence;
  unsigned long long timeout_ns;  // a wait that sees nothing for this long flags APHCG_ERR_COMM
};

// Loop state; lives in device memory, updated by the kernels themselves so the
// host never has to read a scalar inside the loop (reference stages
// "iter2"/"iter3"/"check", src/linear/linear.ipp:83-114).
struct CgState {
  // all-reduced scalars
  // rr / rr_prev are the numerators of alpha and beta: sum r^2 in the reference
  // recurrence, sum r.z (z = r/diag) in the opt-in Jacobi-preconditioned mode
  double rr;        // current                            (dot_r / next dot_r_prev)
  double rr_prev;   // before the last update             (dot_r_prev)
  double rnorm2;    // sum r^2 (the residual norm; equals rr without preconditioner)
  double rnorm2_0;  // sum r^2 of the initial residual (reported as aphcg_info.residual0)
  double pAp;       // sum p*Ap                           (dot_p_lp)
  double max_r;     // max |r|
  double alpha_prev;  // alpha of the last completed iteration (x update is deferred
                      // into the next direction kernel)
  double alpha_prev2; // alpha of the iteration before that (batched x update, see xbatch)
  double residual;
  // this rank's partial results, all-reduced in place when nranks > 1
  double loc_sum;
  double loc_max;
  double loc_sum2;  // preconditioned mode: local sum r^2 (loc_sum holds r.z)
  // Conf (src/linear/linear.h:21-25) + Extra::residual_max
  double tol;
  double cell_volume;
  int miniter, maxiter, maxnorm;
  int precond;   // 1: Jacobi-preconditioned recurrence (opt-in; not the reference's)
  int iter;      // completed iterations
  int done;      // exit rule fired (linear.ipp:110-113); later kernels return at once
  int hist_cap;
  unsigned counter_a, counter_b;  // last-block-done tickets
  int error;     // 1: a peer's contribution did not arrive in time (multi-GPU)
  int xbatch;    // 1: the direction kernel applies the deferred x updates two at a time,
                 // on even iterations only:  x = fma(a_{k-1}, p_{k-1}, fma(a_{k-2}, p_{k-2}, x))
                 // -- the same FMAs in the same order as one per iteration, so x is bitwise
                 // unchanged, but x is read and written every other iteration (p_{k-2} is
                 // still in the buffer p_k is about to overwrite): 12 instead of 16 B/cell
  unsigned long long seq_base;  // run number: distinguishes the mailbox traffic of successive runs
  // Comm::wait_in_kernel: this rank's partial result of the direction / update stage has been
  // pushed, but the all-reduced value is not folded into the state above yet -- the consumer
  // kernel's CTAs fold it for themselves and its last CTA commits it (cg_kernels.cuh)
  int pend_dir, pend_upd;
};

struct DevPtrs {
  const double* a[7];   // SoA coefficients [c,x-,x+,y-,y+,z-,z+], compact
  const double* rhs;    // e7, compact
  double* u;            // iterate, compact
  double* ap;           // A*p, compact
  double* r;            // padded: residual r; in preconditioned mode z = r/diag instead
  double* rc;           // compact true residual (preconditioned mode only)
  double* p[2];         // search direction, padded, ping-pong by iteration parity
  // where this slab's bottom / top inner plane of a padded field must also be
  // stored: the neighbour's ghost plane (peer memory over NVLink when the
  // neighbour is another GPU, own ghost plane when the slab wraps onto itself).
  // Each points at element (i=0, j=0) of the destination plane.
  double* r_lo_dst;     // destination of inner plane k=0      (or nullptr)
  double* r_hi_dst;     // destination of inner plane k=nzl-1  (or nullptr)
  double* p_lo_dst[2];  // same for p[0], p[1] (initial guess copy, Jacobi iterate)
  double* p_hi_dst[2];
  CgState* st;
  double* history;
  double* partials;     // one slot per block for sums
  double* partials2;    // one slot per block for max
  double* partials3;    // one slot per block for the second sum (preconditioned mode)
  Comm cm;
};

}  // namespace acg
