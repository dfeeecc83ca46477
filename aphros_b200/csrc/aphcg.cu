// libaphcg.so: C ABI (include/aphcg.h) and host-side driver of the CG solver.
//
// Mirrors linear::Solver<M>::Solve of the reference (src/linear/linear.h:15-57):
// one call uploads the 8-coefficient rows and the initial guess, runs the
// device-resident loop (cg_kernels.cu) until the reference's exit rule fires
// (src/linear/linear.ipp:102-113) and downloads the solution.  There is no CPU
// fallback: without a usable CUDA device every compute entry point fails.
#include "../../include/aphcg.h"

#include <unistd.h>

#include <algorithm>
#include <cmath>
#include <cstddef>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "cg_group.h"
#include "cg_launch.h"
#include "cg_types.h"
#include "nccl_dl.h"

using namespace acg;

namespace {

thread_local std::string g_error;

int Fail(int code, const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  g_error = buf;
  return code;
}

#define CK(call)                                                                        \
  do {                                                                                  \
    cudaError_t e_ = (call);                                                            \
    if (e_ != cudaSuccess)                                                              \
      return Fail(APHCG_ERR_CUDA, "%s:%d: %s failed: %s", __FILE__, __LINE__, #call,    \
                  cudaGetErrorString(e_));                                              \
  } while (0)

#define NK(call)                                                                        \
  do {                                                                                  \
    ncclResult_t e_ = (call);                                                           \
    if (e_ != ncclSuccess)                                                              \
      return Fail(APHCG_ERR_COMM, "%s:%d: %s failed: %s", __FILE__, __LINE__, #call,    \
                  Nccl().GetErrorString(e_));                                           \
  } while (0)

struct IpcBlob {  // <= APHCG_IPC_BYTES
  cudaIpcMemHandle_t handle;  // 64 bytes
  int64_t ptotal, pz, poff, nzl;
  int32_t rank, pid;
  uint64_t base;  // device pointer (valid in the exporting process only)
  int32_t device;   // ordinal in the exporting process
  int32_t pci;      // PCI domain/bus/device of the GPU: the same for every process that uses it
};
static_assert(sizeof(IpcBlob) <= APHCG_IPC_BYTES, "blob too large");
static_assert(sizeof(ncclUniqueId) <= APHCG_UNIQUE_ID_BYTES, "id too large");

constexpr size_t kStageBytes = size_t(128) << 20;
// iterations enqueued (or replayed as one CUDA graph) between two looks at the exit flag
int ChunkIters() {
  static int n = [] {
    const char* e = getenv("APHCG_CHUNK");
    const int v = e ? atoi(e) : 16;
    return v >= 1 && v <= 1024 ? v : 16;
  }();
  return n;
}

}  // namespace

struct aphcg {
  aphcg_desc desc{};
  Geom g{};
  int vx = 1;
  bool single = true;
  cudaStream_t stream = nullptr, copy_stream = nullptr;
  double* coef[7] = {};
  double* rhs = nullptr;
  double* u = nullptr;
  double* ap = nullptr;
  double* slab = nullptr;  // r | p0 | p1, padded
  CgState* st = nullptr;
  CgState* h_st = nullptr;  // pinned
  double* history = nullptr;
  double* partials = nullptr;
  double* partials2 = nullptr;
  unsigned nslots = 0;
  int hist_cap = 0;
  double* stage[2] = {};
  size_t stage_bytes = 0;
  cudaEvent_t stage_free[2] = {}, stage_ready[2] = {};
  cudaEvent_t ev[4] = {};
  DevPtrs d{};
  bool have_system = false, have_guess = false;
  // loop graph
  cudaGraphExec_t gexec = nullptr;      // general-storage loop
  cudaGraphExec_t gexec_sym = nullptr;  // symmetric-storage loop
  cudaGraphExec_t gexec_jacobi = nullptr;
  bool use_graph = true;
  bool use_tma = false;
  bool xbatch = true;     // batched deferred x updates in the TMA kernel (CgState::xbatch)
  bool precond = false;   // opt-in Jacobi-preconditioned recurrence (APHCG_JACOBI_PRECOND)
  double* rc = nullptr;   // compact residual, preconditioned mode only
  double* partials3 = nullptr;
  bool sym = false;       // resident matrix verified symmetric -> 4-stream kernel
  bool allow_sym = true;
  int* d_flag = nullptr;
  TmaPlan* tma = nullptr;
  Tma2Plan* tma2 = nullptr;  // all-operands-by-TMA variant (symmetric storage only)
  double* scratch = nullptr;  // grow-only device scratch of the assemblers (inputs staged here)
  size_t scratch_bytes = 0;
  bool persist = false;   // the loop runs as one persistent cooperative kernel (small meshes)
  PersistPlan pplan{};
  // comm
  ncclComm_t comm = nullptr;
  GroupSync* gs = nullptr;             // in-process slab group (aphcg_group_*): replaces NCCL
  double* peer[kMaxRanks] = {};        // every rank's slab (own pointer for this rank)
  bool peer_opened[kMaxRanks] = {};
  bool use_mail = true;                // scalar all-reduce through peer mailboxes (else NCCL)
  bool wait_in_kernel = false;         // Comm::wait_in_kernel (decided in aphcg_ipc_connect)
  bool connected = false;
  unsigned long long runs = 0;
  int64_t launches = 0;
  int last_iter = 0;
};

namespace {

int SetDevice(aphcg_t* h) {
  CK(cudaSetDevice(h->desc.device));
  return 0;
}

void InvalidateGraphs(aphcg_t* h) {
  if (h->gexec) cudaGraphExecDestroy(h->gexec);
  if (h->gexec_sym) cudaGraphExecDestroy(h->gexec_sym);
  h->gexec_sym = nullptr;
  if (h->gexec_jacobi) cudaGraphExecDestroy(h->gexec_jacobi);
  h->gexec = nullptr;
  h->gexec_jacobi = nullptr;
}

void FillDevPtrs(aphcg_t* h) {
  DevPtrs& d = h->d;
  for (int q = 0; q < 7; ++q) d.a[q] = h->coef[q];
  d.rhs = h->rhs;
  d.u = h->u;
  d.ap = h->ap;
  d.r = h->slab;
  d.p[0] = h->slab + h->g.ptotal;
  d.p[1] = h->slab + 2 * h->g.ptotal;
  d.st = h->st;
  d.history = h->history;
  d.partials = h->partials;
  d.partials2 = h->partials2;
  d.partials3 = h->partials3;
  d.rc = h->rc;
  d.r_lo_dst = d.r_hi_dst = nullptr;
  d.p_lo_dst[0] = d.p_lo_dst[1] = d.p_hi_dst[0] = d.p_hi_dst[1] = nullptr;
  memset(&d.cm, 0, sizeof(d.cm));
  d.cm.nranks = 1;
}

// z images when the slab wraps onto itself (one rank, periodic z)
void SelfConnect(aphcg_t* h) {
  const Geom& g = h->g;
  DevPtrs& d = h->d;
  if (h->desc.nranks == 1 && h->desc.periodic[2]) {
    double* f[3] = {d.r, d.p[0], d.p[1]};
    double* lo[3];
    double* hi[3];
    for (int m = 0; m < 3; ++m) {
      lo[m] = f[m] + g.poff + (int64_t)g.nzl * g.pz;  // bottom plane -> top ghost
      hi[m] = f[m] + g.poff - g.pz;                    // top plane -> bottom ghost
    }
    d.r_lo_dst = lo[0];
    d.r_hi_dst = hi[0];
    d.p_lo_dst[0] = lo[1];
    d.p_hi_dst[0] = hi[1];
    d.p_lo_dst[1] = lo[2];
    d.p_hi_dst[1] = hi[2];
  }
}

int AllReduce(aphcg_t* h, double* p, ncclRedOp_t op) {
  NK(Nccl().AllReduce(p, p, 1, ncclDouble, op, h->comm, h->stream));
  return 0;
}

// One CG iteration enqueued on h->stream.
int EnqueueIteration(aphcg_t* h) {
  if (h->use_tma && h->tma2 && h->sym) {
    launch_dir_spmv_stream(h->tma2, h->g, h->d, h->single, h->stream);
  } else if (h->use_tma) {
    launch_dir_spmv_tma(h->tma, h->g, h->d, h->single, h->sym, h->stream);
  } else {
    launch_dir_spmv_plain(h->g, h->d, h->vx, h->single, h->stream);
  }
  if (!h->single && !h->wait_in_kernel) {
    if (!h->use_mail) {
      if (int rc = AllReduce(h, &h->st->loc_sum, ncclSum)) return rc;
    }
    launch_finish_dir(h->d, h->stream);  // with mailboxes: waits for all ranks' partials
  }
  launch_update(h->g, h->d, h->vx, h->single, h->precond, h->stream);
  if (!h->single && !h->wait_in_kernel) {
    if (!h->use_mail) {
      if (int rc = AllReduce(h, &h->st->loc_sum, ncclSum)) return rc;
      if (h->precond) {
        if (int rc = AllReduce(h, &h->st->loc_sum2, ncclSum)) return rc;
      }
      if (h->desc.flags & APHCG_MAXNORM) {
        if (int rc = AllReduce(h, &h->st->loc_max, ncclMax)) return rc;
      }
    }
    launch_finish_upd(h->d, h->stream);
  }
  return 0;
}

int EnqueueJacobiIteration(aphcg_t* h) {
  launch_jacobi(h->g, h->d, h->vx, h->single, h->stream);
  if (!h->single) {
    if (!h->use_mail) {
      if (int rc = AllReduce(h, &h->st->loc_max, ncclMax)) return rc;
    }
    launch_finish_jacobi(h->d, h->stream);  // with mailboxes: waits for all ranks' maxima
  }
  return 0;
}

int LaunchesPerIter(const aphcg_t* h) { return (h->single || h->wait_in_kernel) ? 2 : 4; }

// ChunkIters() iterations: what the host enqueues (or replays as one graph) between two looks
// at the exit flag.  With Comm::wait_in_kernel the update stage of the chunk's last iteration
// is still pending at its end; one k_finish_upd folds it so that the host reads a committed
// state (iteration count, residual, exit flag).
int EnqueueChunk(aphcg_t* h, bool jacobi) {
  for (int i = 0; i < ChunkIters(); ++i) {
    if (int rc = jacobi ? EnqueueJacobiIteration(h) : EnqueueIteration(h)) return rc;
  }
  if (!jacobi && h->wait_in_kernel) launch_finish_upd(h->d, h->stream);
  return 0;
}

int BuildGraph(aphcg_t* h, bool jacobi, cudaGraphExec_t* out) {
  cudaGraph_t graph = nullptr;
  CK(cudaStreamBeginCapture(h->stream, cudaStreamCaptureModeThreadLocal));
  const int rc = EnqueueChunk(h, jacobi);
  cudaError_t e = cudaStreamEndCapture(h->stream, &graph);
  if (rc) {
    if (graph) cudaGraphDestroy(graph);
    return rc;
  }
  if (e != cudaSuccess) return Fail(APHCG_ERR_CUDA, "graph capture failed: %s", cudaGetErrorString(e));
  e = cudaGraphInstantiate(out, graph, 0);
  cudaGraphDestroy(graph);
  if (e != cudaSuccess)
    return Fail(APHCG_ERR_CUDA, "graph instantiate failed: %s", cudaGetErrorString(e));
  return 0;
}

// cross-rank barrier on the stream (orders peer-memory writes before reads)
int StreamBarrier(aphcg_t* h) {
  if (h->single) return 0;
  if (h->gs) {  // slabs of one process: drain the stream, then meet the other slab threads
    CK(cudaStreamSynchronize(h->stream));
    if (!h->gs->Wait()) return Fail(APHCG_ERR_COMM, "another slab of the group failed");
    return 0;
  }
  return AllReduce(h, &h->st->loc_max, ncclMax);
}

// Sums over ranks of st->loc_sum and st->loc_sum2, once per solve (the numerator of the
// first alpha and sum r^2 of the initial residual), in rank order.
int AllReduceInitial(aphcg_t* h) {
  if (!h->gs) {
    if (int rc = AllReduce(h, &h->st->loc_sum, ncclSum)) return rc;
    return AllReduce(h, &h->st->loc_sum2, ncclSum);
  }
  static_assert(offsetof(CgState, loc_sum2) - offsetof(CgState, loc_sum) == 2 * sizeof(double),
                "loc_sum, loc_max, loc_sum2 are copied as one block");
  CK(cudaMemcpyAsync(&h->h_st->loc_sum, &h->st->loc_sum, 3 * sizeof(double),
                     cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  const int me = h->desc.rank, n = h->desc.nranks;
  h->gs->red[me] = h->h_st->loc_sum;
  h->gs->red2[me] = h->h_st->loc_sum2;
  if (!h->gs->Wait()) return Fail(APHCG_ERR_COMM, "another slab of the group failed");
  double sum = 0.0, sum2 = 0.0;
  for (int q = 0; q < n; ++q) {
    sum += h->gs->red[q];
    sum2 += h->gs->red2[q];
  }
  // nobody overwrites red[] before every slab has read it
  if (!h->gs->Wait()) return Fail(APHCG_ERR_COMM, "another slab of the group failed");
  h->h_st->loc_sum = sum;
  h->h_st->loc_sum2 = sum2;
  CK(cudaMemcpyAsync(&h->st->loc_sum, &h->h_st->loc_sum, 3 * sizeof(double),
                     cudaMemcpyHostToDevice, h->stream));
  // h_st is reused by the loop's polls: the copy must have read it before they land
  CK(cudaStreamSynchronize(h->stream));
  return 0;
}

int CheckLayout(const aphcg_t* h, const aphcg_layout* l, aphcg_layout* out) {
  if (l) {
    *out = *l;
    if (l->stride_y < h->g.nx || l->stride_z < l->stride_y * (int64_t)h->g.ny || l->offset < 0)
      return Fail(APHCG_ERR_ARG, "bad layout: offset=%lld stride_y=%lld stride_z=%lld",
                  (long long)l->offset, (long long)l->stride_y, (long long)l->stride_z);
  } else {
    out->offset = 0;
    out->stride_y = h->g.nx;
    out->stride_z = (int64_t)h->g.nx * h->g.ny;
  }
  return 0;
}

bool IsCompact(const aphcg_t* h, const aphcg_layout& l) {
  return l.stride_y == h->g.nx && l.stride_z == (int64_t)h->g.nx * h->g.ny;
}

// Copies planes [k0,k0+nk) of a laid-out host/device array of `elem`-byte cells
// into a compact device/host chunk (or back), on stream s.
int CopyPlanes(const aphcg_t* h, void* dst, const void* src, const aphcg_layout& l, bool src_laid,
               int k0, int nk, size_t elem, cudaMemcpyKind kind, cudaStream_t s) {
  const Geom& g = h->g;
  const size_t row = (size_t)g.nx * elem;
  if (IsCompact(h, l)) {
    const size_t off = ((size_t)l.offset + (size_t)k0 * l.stride_z) * elem;
    const size_t bytes = (size_t)nk * g.ny * row;
    if (src_laid) {
      CK(cudaMemcpyAsync(dst, (const char*)src + off, bytes, kind, s));
    } else {
      CK(cudaMemcpyAsync((char*)dst + off, src, bytes, kind, s));
    }
    return 0;
  }
  for (int k = 0; k < nk; ++k) {
    const size_t off = ((size_t)l.offset + (size_t)(k0 + k) * l.stride_z) * elem;
    const size_t coff = (size_t)k * g.ny * row;
    if (src_laid) {
      CK(cudaMemcpy2DAsync((char*)dst + coff, row, (const char*)src + off, l.stride_y * elem, row,
                           g.ny, kind, s));
    } else {
      CK(cudaMemcpy2DAsync((char*)dst + off, l.stride_y * elem, (const char*)src + coff, row, row,
                           g.ny, kind, s));
    }
  }
  return 0;
}

// The system just became resident: decide whether the 4-stream (symmetric
// storage) kernel may be used for it.  One pass over the off-diagonals.
int SystemResident(aphcg_t* h) {
  h->have_system = true;
  h->sym = false;
  if (!h->use_tma || !h->allow_sym) return 0;
  int flag = 0;
  CK(cudaMemsetAsync(h->d_flag, 0, sizeof(int), h->stream));
  launch_check_symmetry(h->g, h->d, h->d_flag, h->stream);
  h->launches++;
  CK(cudaMemcpyAsync(&flag, h->d_flag, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  CK(cudaGetLastError());
  h->sym = (flag == 0);
  return 0;
}

// Two staging buffers for the row upload: 128 MB each, or one xy-plane of rows if that is
// larger (the transpose works on whole planes).
int EnsureStage(aphcg_t* h) {
  h->stage_bytes = std::max(kStageBytes, (size_t)h->g.cz * 64);
  for (int b = 0; b < 2; ++b) {
    if (!h->stage[b]) {
      CK(cudaMalloc(&h->stage[b], h->stage_bytes));
      CK(cudaEventCreateWithFlags(&h->stage_free[b], cudaEventDisableTiming));
      CK(cudaEventCreateWithFlags(&h->stage_ready[b], cudaEventDisableTiming));
    }
  }
  return 0;
}

// Grow-only device scratch (the assemblers stage their inputs here): no cudaMalloc/cudaFree
// pair per call on the per-time-step path.
int EnsureScratch(aphcg_t* h, size_t bytes) {
  if (bytes <= h->scratch_bytes) return 0;
  CK(cudaStreamSynchronize(h->stream));
  if (h->scratch) CK(cudaFree(h->scratch));
  h->scratch = nullptr;
  h->scratch_bytes = 0;
  CK(cudaMalloc(&h->scratch, bytes));
  h->scratch_bytes = bytes;
  return 0;
}

int WriteState(aphcg_t* h, const aphcg_conf* conf) {
  CgState s{};
  s.tol = conf->tol;
  s.miniter = conf->miniter;
  s.maxiter = conf->maxiter;
  s.maxnorm = (h->desc.flags & APHCG_MAXNORM) ? 1 : 0;
  s.precond = h->precond ? 1 : 0;
  s.cell_volume = h->desc.cell_volume;
  s.hist_cap = h->hist_cap;
  s.seq_base = ++h->runs;
  s.xbatch = (h->use_tma && h->xbatch && !h->persist) ? 1 : 0;
  *h->h_st = s;
  CK(cudaMemcpyAsync(h->st, h->h_st, sizeof(CgState), cudaMemcpyHostToDevice, h->stream));
  return 0;
}

// The residual history keeps the first kHistoryMax iterations at most (writes are bounded by
// hist_cap on the device): a tolerance-driven Conf with a huge maxiter must not turn into a
// multi-GB allocation, and growing is rare because a reallocation invalidates the graphs.
constexpr int64_t kHistoryMax = 1 << 16;
int EnsureHistory(aphcg_t* h, int maxiter) {
  const int need = (int)std::min<int64_t>(std::max<int64_t>((int64_t)maxiter + 2, 16), kHistoryMax);
  if (need > h->hist_cap) {
    CK(cudaStreamSynchronize(h->stream));
    if (h->history) CK(cudaFree(h->history));
    CK(cudaMalloc(&h->history, sizeof(double) * (size_t)need));
    h->hist_cap = need;
    h->d.history = h->history;
    InvalidateGraphs(h);
  }
  return 0;
}

// Replays iterations until the device-side exit rule fires.
// The persistent kernel stops by itself; the host only bounds a launch (so that a
// tolerance-driven solve is looked at now and then) and repeats until the exit flag is set.
int RunLoop(aphcg_t* h, bool jacobi, const aphcg_conf* conf);

int RunLoopPersistent(aphcg_t* h, const aphcg_conf* conf) {
  const long limit = std::max<long>((long)conf->maxiter + 1, (long)conf->miniter);
  long enq = 0;
  for (;;) {
    const int n = (int)std::min<long>(limit - enq > 0 ? limit - enq : 1, 4096);
    const cudaError_t le = launch_cg_persistent(h->g, h->d, h->vx, h->pplan, n, h->stream);
    if (le != cudaSuccess && enq == 0) {
      // the cooperative launch was refused (co-residency not available in this context, e.g.
      // under MPS limits): nothing has run yet, so take the two-kernel iteration for this
      // handle from now on.  The state already says "no batched x update", which those
      // kernels honour.
      cudaGetLastError();
      h->persist = false;
      return RunLoop(h, false, conf);
    }
    CK(le);
    h->launches++;
    enq += n;
    CK(cudaMemcpyAsync(h->h_st, h->st, sizeof(CgState), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    if (h->h_st->done) break;
    if (enq > limit + 4096)
      return Fail(APHCG_ERR_STATE, "loop did not terminate (iter=%d)", h->h_st->iter);
  }
  return 0;
}

int RunLoop(aphcg_t* h, bool jacobi, const aphcg_conf* conf) {
  if (h->persist && !jacobi) return RunLoopPersistent(h, conf);
  cudaGraphExec_t* gx = jacobi ? &h->gexec_jacobi : (h->sym ? &h->gexec_sym : &h->gexec);
  if (h->use_graph && !*gx) {
    if (int rc = BuildGraph(h, jacobi, gx)) return rc;
  }
  // Upper bound on useful iterations: maxiter+1 (linear.ipp:110-113) or miniter.
  const long limit = std::max<long>((long)conf->maxiter + 1, (long)conf->miniter);
  long enq = 0;
  for (;;) {
    if (h->use_graph) {
      CK(cudaGraphLaunch(*gx, h->stream));
    } else {
      if (int rc = EnqueueChunk(h, jacobi)) return rc;
    }
    enq += ChunkIters();
    // poll the exit flag only when it can have fired: always with a tolerance,
    // otherwise once the iteration limit is covered
    if (conf->tol > 0 || enq >= limit) {
      CK(cudaMemcpyAsync(h->h_st, h->st, sizeof(CgState), cudaMemcpyDeviceToHost, h->stream));
      CK(cudaStreamSynchronize(h->stream));
      if (h->h_st->done) break;
      if (enq > limit + ChunkIters())
        return Fail(APHCG_ERR_STATE, "loop did not terminate (iter=%d)", h->h_st->iter);
    }
  }
  h->launches += (int64_t)h->h_st->iter * (jacobi ? (h->single ? 1 : 2) : LaunchesPerIter(h));
  if (!jacobi && h->wait_in_kernel) h->launches += (h->h_st->iter + ChunkIters() - 1) / ChunkIters();
  return 0;
}

}  // namespace

// ================================================================================
extern "C" {

const char* aphcg_last_error(void) { return g_error.c_str(); }
int aphcg_version(void) { return APHCG_VERSION; }

int aphcg_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return n;
}

int aphcg_host_alloc(void** out, uint64_t bytes) {
  if (!out) return Fail(APHCG_ERR_ARG, "null out");
  CK(cudaHostAlloc(out, bytes, cudaHostAllocDefault));
  return 0;
}
int aphcg_host_free(void* p) {
  CK(cudaFreeHost(p));
  return 0;
}

int aphcg_create(aphcg_t** out, const aphcg_desc* desc) {
  if (!out || !desc) return Fail(APHCG_ERR_ARG, "null argument");
  *out = nullptr;
  const aphcg_desc& ds = *desc;
  if (ds.nx < 1 || ds.ny < 1 || ds.nz < 1 || ds.nx > (1 << 20) || ds.ny > (1 << 20) ||
      ds.nz > (1 << 20))
    return Fail(APHCG_ERR_ARG, "bad mesh size %lld x %lld x %lld", (long long)ds.nx,
                (long long)ds.ny, (long long)ds.nz);
  if (ds.nranks > kMaxRanks) return Fail(APHCG_ERR_ARG, "at most %d ranks", kMaxRanks);
  if (ds.nranks < 1 || ds.rank < 0 || ds.rank >= ds.nranks)
    return Fail(APHCG_ERR_ARG, "bad rank %d of %d", ds.rank, ds.nranks);
  if (ds.nz_local < 1 || ds.z0 < 0 || ds.z0 + ds.nz_local > ds.nz)
    return Fail(APHCG_ERR_ARG, "bad slab z0=%lld nz_local=%lld of nz=%lld", (long long)ds.z0,
                (long long)ds.nz_local, (long long)ds.nz);
  if (ds.nranks == 1 && (ds.z0 != 0 || ds.nz_local != ds.nz))
    return Fail(APHCG_ERR_ARG, "one rank must own the whole domain");
  if (!(ds.cell_volume > 0)) return Fail(APHCG_ERR_ARG, "cell_volume must be positive");
  int ndev = aphcg_device_count();
  if (ndev <= 0)
    return Fail(APHCG_ERR_CUDA, "no CUDA device available (aphcg has no CPU fallback)");
  if (ds.device < 0 || ds.device >= ndev)
    return Fail(APHCG_ERR_ARG, "device %d out of range (%d devices)", ds.device, ndev);

  aphcg_t* h = new aphcg();
  h->desc = ds;
  h->single = (ds.nranks == 1);
  Geom& g = h->g;
  g.nx = (int)ds.nx;
  g.ny = (int)ds.ny;
  g.nzl = (int)ds.nz_local;
  g.per_x = ds.periodic[0] != 0;
  g.per_y = ds.periodic[1] != 0;
  g.cy = g.nx;
  g.cz = (int64_t)g.nx * g.ny;
  g.ncell = g.cz * g.nzl;
  g.py = ((int64_t)kGhostX + g.nx + 1 + 15) / 16 * 16;
  g.pz = g.py * (g.ny + 2);
  g.poff = kGhostX + g.py + g.pz;
  g.ptotal = g.pz * (g.nzl + 2);
  h->vx = (g.nx % 2 == 0) ? 2 : 1;
  // planes per CTA of the tiled kernels: 8 on large meshes; fewer on small ones so that
  // the grid still fills the 148 SMs several times (they are latency-bound there)
  g.zc = 8;
  while (g.zc > 1 && tile_blocks_for(g, h->vx) < (unsigned)kNumSMs * 8u) g.zc /= 2;
  if (const char* ez = getenv("APHCG_TILE_ZC")) g.zc = std::max(1, atoi(ez));
  // update kernel: as many threads along x as a row has 128-bit pairs (power of two)
  g.utx = 32;
  while (g.utx < 256 && g.utx * h->vx < g.nx) g.utx *= 2;
  h->precond = (ds.flags & APHCG_JACOBI_PRECOND) != 0;
  h->use_graph = !(ds.flags & APHCG_NO_GRAPH);
  if (const char* eg = getenv("APHCG_GRAPH")) h->use_graph = atoi(eg) != 0;

  auto cleanup = [&](int rc) {
    aphcg_destroy(h);
    return rc;
  };
#define CKC(call)                                                                      \
  do {                                                                                 \
    cudaError_t e_ = (call);                                                           \
    if (e_ != cudaSuccess)                                                             \
      return cleanup(Fail(APHCG_ERR_CUDA, "%s:%d: %s failed: %s", __FILE__, __LINE__,  \
                          #call, cudaGetErrorString(e_)));                             \
  } while (0)
  CKC(cudaSetDevice(ds.device));
  CKC(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
  CKC(cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking));
  for (auto& e : h->ev) CKC(cudaEventCreate(&e));
  const size_t nb = sizeof(double) * (size_t)g.ncell;
  for (int q = 0; q < 7; ++q) CKC(cudaMalloc(&h->coef[q], nb));
  CKC(cudaMalloc(&h->rhs, nb));
  CKC(cudaMalloc(&h->u, nb));
  CKC(cudaMalloc(&h->ap, nb));
  const size_t slab_bytes = sizeof(double) * 3 * (size_t)g.ptotal + sizeof(MailSlot) * kMailSlots;
  CKC(cudaMalloc(&h->slab, slab_bytes));
  CKC(cudaMemset(h->slab, 0, slab_bytes));
  CKC(cudaMalloc(&h->d_flag, sizeof(int)));
  CKC(cudaMalloc(&h->st, sizeof(CgState)));
  CKC(cudaMemset(h->st, 0, sizeof(CgState)));
  CKC(cudaHostAlloc(&h->h_st, sizeof(CgState), cudaHostAllocDefault));
  h->nslots = std::max(tile_blocks(g, h->vx), 1u);
  FillDevPtrs(h);
  SelfConnect(h);
  // TMA-staged kernel unless disabled or the geometry does not qualify
  h->allow_sym = !(ds.flags & APHCG_NO_SYM);
  h->use_mail = !(ds.flags & APHCG_NCCL_REDUCE);
  if (const char* em = getenv("APHCG_ALLREDUCE")) h->use_mail = strcmp(em, "nccl") != 0;
  if (const char* es = getenv("APHCG_SYM")) h->allow_sym = atoi(es) != 0;
  if (const char* ex = getenv("APHCG_XBATCH")) h->xbatch = atoi(ex) != 0;
  const char* env = getenv("APHCG_SPMV");
  bool want_tma = !(ds.flags & APHCG_NO_TMA);
  if (env && !strcmp(env, "plain")) want_tma = false;
  if (want_tma) {
    char err[256] = "";
    h->tma = tma_plan_create(g, h->d, err, sizeof(err));
    if (h->tma) {
      h->use_tma = true;
      h->nslots = std::max(h->nslots, tma_plan_blocks(h->tma));
      // symmetric storage: the variant with every operand staged by TMA (cg_spmv_tma2.cu),
      // 1.60 vs 1.84 ms at 512^3; APHCG_STREAM=0 / APHCG_NO_STREAM keep the LDG-fed kernel
      bool want2 = !(ds.flags & APHCG_NO_STREAM);
      if (const char* e2 = getenv("APHCG_STREAM")) want2 = atoi(e2) != 0;
      if (want2) {
        h->tma2 = tma2_plan_create(g, h->d, err, sizeof(err));
        if (h->tma2) h->nslots = std::max(h->nslots, tma2_plan_blocks(h->tma2));
      }
    } else if (env && !strcmp(env, "tma")) {
      return cleanup(Fail(APHCG_ERR_CUDA, "TMA kernel requested but unavailable: %s", err));
    }
  }
  // Small single-GPU meshes: the whole loop as one persistent cooperative kernel (DESIGN.md
  // section 3).  Limit: measured crossover with the two-kernel iteration (64^3: 10.3 vs 25.1 us
  // per iteration; 96^3: 32.8 vs 30.7).
  {
    int64_t max_cells = 700000;
    if (const char* ep = getenv("APHCG_PERSISTENT")) max_cells = atoll(ep) == 1 ? max_cells : atoll(ep);
    int coop = 0;
    cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, ds.device);
    h->persist = h->single && !h->precond && coop && !(ds.flags & APHCG_NO_PERSISTENT) &&
                 g.ncell <= max_cells && persistent_plan(g, h->vx, &h->pplan);
    if (h->persist) h->nslots = std::max(h->nslots, h->pplan.grid);
  }
  CKC(cudaMalloc(&h->partials, sizeof(double) * h->nslots));
  CKC(cudaMalloc(&h->partials2, sizeof(double) * h->nslots));
  CKC(cudaMalloc(&h->partials3, sizeof(double) * h->nslots));
  h->d.partials = h->partials;
  h->d.partials2 = h->partials2;
  h->d.partials3 = h->partials3;
  if (h->precond) {
    CKC(cudaMalloc(&h->rc, nb));
    h->d.rc = h->rc;
  }
  h->hist_cap = 1024;
  CKC(cudaMalloc(&h->history, sizeof(double) * h->hist_cap));
  h->d.history = h->history;
  CKC(cudaDeviceSynchronize());
#undef CKC
  *out = h;
  return 0;
}

int aphcg_destroy(aphcg_t* h) {
  if (!h) return 0;
  cudaSetDevice(h->desc.device);
  if (h->stream) cudaStreamSynchronize(h->stream);
  InvalidateGraphs(h);
  if (h->tma) tma_plan_destroy(h->tma);
  if (h->tma2) tma2_plan_destroy(h->tma2);
  if (h->comm) Nccl().CommDestroy(h->comm);
  for (int q = 0; q < kMaxRanks; ++q)
    if (h->peer_opened[q]) cudaIpcCloseMemHandle(h->peer[q]);
  for (int q = 0; q < 7; ++q) cudaFree(h->coef[q]);
  cudaFree(h->rhs);
  cudaFree(h->u);
  cudaFree(h->ap);
  cudaFree(h->slab);
  cudaFree(h->st);
  cudaFree(h->d_flag);
  cudaFree(h->history);
  cudaFree(h->partials);
  cudaFree(h->partials2);
  cudaFree(h->partials3);
  cudaFree(h->rc);
  cudaFree(h->scratch);
  if (h->h_st) cudaFreeHost(h->h_st);
  for (int b = 0; b < 2; ++b) {
    cudaFree(h->stage[b]);
    if (h->stage_free[b]) cudaEventDestroy(h->stage_free[b]);
    if (h->stage_ready[b]) cudaEventDestroy(h->stage_ready[b]);
  }
  for (auto& e : h->ev)
    if (e) cudaEventDestroy(e);
  if (h->copy_stream) cudaStreamDestroy(h->copy_stream);
  if (h->stream) cudaStreamDestroy(h->stream);
  cudaGetLastError();
  delete h;
  return 0;
}

// ---- uploads -------------------------------------------------------------------
int aphcg_upload_system(aphcg_t* h, const double* system, const aphcg_layout* layout) {
  if (!h || !system) return Fail(APHCG_ERR_ARG, "null argument");
  if (int rc = SetDevice(h)) return rc;
  aphcg_layout l;
  if (int rc = CheckLayout(h, layout, &l)) return rc;
  if (int rc = EnsureStage(h)) return rc;
  const Geom& g = h->g;
  const size_t plane = (size_t)g.cz * 64;
  const int per_chunk = (int)std::max<size_t>(1, h->stage_bytes / plane);
  int b = 0;
  for (int k0 = 0; k0 < g.nzl; k0 += per_chunk, b ^= 1) {
    const int nk = std::min(per_chunk, g.nzl - k0);
    // copy engine fills stage[b] once the transpose that last read it is done
    CK(cudaStreamWaitEvent(h->copy_stream, h->stage_free[b], 0));
    if (int rc = CopyPlanes(h, h->stage[b], system, l, true, k0, nk, 64, cudaMemcpyHostToDevice,
                            h->copy_stream))
      return rc;
    CK(cudaEventRecord(h->stage_ready[b], h->copy_stream));
    CK(cudaStreamWaitEvent(h->stream, h->stage_ready[b], 0));
    launch_rows_to_soa(g, h->stage[b], 0, g.cy, g.cz, k0, nk, h->coef, h->rhs, h->stream);
    h->launches++;
    CK(cudaEventRecord(h->stage_free[b], h->stream));
  }
  CK(cudaGetLastError());
  return SystemResident(h);
}

int aphcg_set_system_device(aphcg_t* h, const double* d_system, const aphcg_layout* layout) {
  if (!h || !d_system) return Fail(APHCG_ERR_ARG, "null argument");
  if (int rc = SetDevice(h)) return rc;
  aphcg_layout l;
  if (int rc = CheckLayout(h, layout, &l)) return rc;
  launch_rows_to_soa(h->g, d_system, l.offset, l.stride_y, l.stride_z, 0, h->g.nzl, h->coef,
                     h->rhs, h->stream);
  h->launches++;
  CK(cudaGetLastError());
  return SystemResident(h);
}

static int ScatterGuess(aphcg_t* h, const double* d_src, const aphcg_layout& l) {
  launch_scatter_field(h->g, d_src, l.offset, l.stride_y, l.stride_z, h->u, h->d.p[1],
                       h->d.p_lo_dst[1], h->d.p_hi_dst[1], h->vx, h->stream);
  h->launches++;
  CK(cudaGetLastError());
  h->have_guess = true;
  return 0;
}

int aphcg_upload_guess(aphcg_t* h, const double* x0, const aphcg_layout* layout) {
  if (!h) return Fail(APHCG_ERR_ARG, "null handle");
  if (int rc = SetDevice(h)) return rc;
  aphcg_layout l;
  if (int rc = CheckLayout(h, layout, &l)) return rc;
  if (!x0) {
    aphcg_layout c{0, h->g.cy, h->g.cz};
    return ScatterGuess(h, nullptr, c);
  }
  // stage through Ap (free between solves): compact copy, then scatter
  if (int rc = CopyPlanes(h, h->ap, x0, l, true, 0, h->g.nzl, 8, cudaMemcpyHostToDevice, h->stream))
    return rc;
  aphcg_layout c{0, h->g.cy, h->g.cz};
  return ScatterGuess(h, h->ap, c);
}

int aphcg_set_guess_device(aphcg_t* h, const double* d_x0, const aphcg_layout* layout) {
  if (!h) return Fail(APHCG_ERR_ARG, "null handle");
  if (int rc = SetDevice(h)) return rc;
  aphcg_layout l;
  if (int rc = CheckLayout(h, layout, &l)) return rc;
  return ScatterGuess(h, d_x0, l);
}

// ---- run -----------------------------------------------------------------------
static int CheckRun(aphcg_t* h, const aphcg_conf* conf) {
  if (!h || !conf) return Fail(APHCG_ERR_ARG, "null argument");
  if (!h->have_system) return Fail(APHCG_ERR_STATE, "no system uploaded");
  if (conf->maxiter < 0 || conf->miniter < 0) return Fail(APHCG_ERR_ARG, "negative iteration limit");
  if (!h->single && (!(h->comm || h->gs) || !h->connected))
    return Fail(APHCG_ERR_STATE, "nranks > 1 needs aphcg_comm_init and aphcg_ipc_connect first");
  return 0;
}

static int FinishRun(aphcg_t* h, aphcg_info* info) {
  CK(cudaEventRecord(h->ev[2], h->stream));
  CK(cudaMemcpyAsync(h->h_st, h->st, sizeof(CgState), cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  CK(cudaGetLastError());
  h->last_iter = h->h_st->iter;
  h->have_guess = false;  // the guess buffers were consumed
  if (h->h_st->error)
    return Fail(APHCG_ERR_COMM, "a peer rank's partial sum did not arrive (iteration %d)",
                h->h_st->iter);
  if (info) {
    info->residual = h->h_st->residual;
    info->iter = h->h_st->iter;
    info->reserved = 0;
    info->residual0 = sqrt(h->h_st->rnorm2_0 / h->desc.cell_volume);
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, h->ev[1], h->ev[2]));
    info->loop_ms = ms;
    CK(cudaEventElapsedTime(&ms, h->ev[0], h->ev[2]));
    info->total_ms = ms;
  }
  return 0;
}

int aphcg_run(aphcg_t* h, const aphcg_conf* conf, aphcg_info* info) {
  if (int rc = CheckRun(h, conf)) return rc;
  if (int rc = SetDevice(h)) return rc;
  if (!h->have_guess) {
    if (int rc = aphcg_upload_guess(h, nullptr, nullptr)) return rc;
  }
  if (int rc = EnsureHistory(h, conf->maxiter)) return rc;
  CK(cudaEventRecord(h->ev[0], h->stream));
  if (int rc = WriteState(h, conf)) return rc;
  // p_{-1} = 0: the first direction is p = r + 0*0 (linear.ipp:59-62)
  CK(cudaMemsetAsync(h->d.p[0], 0, sizeof(double) * (size_t)h->g.ptotal, h->stream));
  if (int rc = StreamBarrier(h)) return rc;  // neighbours' guess planes have landed
  launch_init_residual(h->g, h->d, h->vx, h->single, h->precond, h->stream);
  h->launches++;
  if (!h->single) {
    if (int rc = AllReduceInitial(h)) return rc;
    launch_finish_init(h->d, h->stream);
    h->launches++;
  }
  CK(cudaEventRecord(h->ev[1], h->stream));
  if (int rc = RunLoop(h, false, conf)) return rc;
  launch_final_update(h->g, h->d, h->vx, h->stream);
  h->launches++;
  return FinishRun(h, info);
}

int aphcg_run_jacobi(aphcg_t* h, const aphcg_conf* conf, aphcg_info* info) {
  if (int rc = CheckRun(h, conf)) return rc;
  if (int rc = SetDevice(h)) return rc;
  if (!h->have_guess) {
    if (int rc = aphcg_upload_guess(h, nullptr, nullptr)) return rc;
  }
  if (int rc = EnsureHistory(h, conf->maxiter)) return rc;
  CK(cudaEventRecord(h->ev[0], h->stream));
  if (int rc = WriteState(h, conf)) return rc;
  if (int rc = StreamBarrier(h)) return rc;  // neighbours' guess planes have landed
  // iterate starts in p[0] (parity of iter = 0): copy the padded guess over
  CK(cudaMemcpyAsync(h->d.p[0], h->d.p[1], sizeof(double) * (size_t)h->g.ptotal,
                     cudaMemcpyDeviceToDevice, h->stream));
  // Iteration 0 of a NEIGHBOUR stores its new boundary plane into this slab's p[1] ghost
  // plane, which the copy above still reads: nobody starts iterating before every slab's
  // copy is done.
  if (int rc = StreamBarrier(h)) return rc;
  CK(cudaEventRecord(h->ev[1], h->stream));
  if (int rc = RunLoop(h, true, conf)) return rc;
  // result: inner cells of the current iterate -> compact u
  {
    const Geom& g = h->g;
    const double* cur = h->d.p[h->h_st->iter & 1];
    CK(cudaMemcpy2DAsync(h->u, sizeof(double) * g.nx, cur + g.poff, sizeof(double) * g.py,
                         sizeof(double) * g.nx, (size_t)g.ny, cudaMemcpyDeviceToDevice, h->stream));
    // planes are not contiguous in the padded layout (ghost rows in between)
    for (int k = 1; k < g.nzl; ++k)
      CK(cudaMemcpy2DAsync(h->u + k * g.cz, sizeof(double) * g.nx, cur + g.poff + k * g.pz,
                           sizeof(double) * g.py, sizeof(double) * g.nx, (size_t)g.ny,
                           cudaMemcpyDeviceToDevice, h->stream));
  }
  return FinishRun(h, info);
}

int aphcg_download_solution(aphcg_t* h, double* x, const aphcg_layout* layout) {
  if (!h || !x) return Fail(APHCG_ERR_ARG, "null argument");
  if (int rc = SetDevice(h)) return rc;
  aphcg_layout l;
  if (int rc = CheckLayout(h, layout, &l)) return rc;
  if (int rc = CopyPlanes(h, x, h->u, l, false, 0, h->g.nzl, 8, cudaMemcpyDeviceToHost, h->stream))
    return rc;
  CK(cudaStreamSynchronize(h->stream));
  return 0;
}

int aphcg_get_solution_device(aphcg_t* h, double* d_x, const aphcg_layout* layout) {
  if (!h || !d_x) return Fail(APHCG_ERR_ARG, "null argument");
  if (int rc = SetDevice(h)) return rc;
  aphcg_layout l;
  if (int rc = CheckLayout(h, layout, &l)) return rc;
  launch_gather_field(h->g, h->u, d_x, l.offset, l.stride_y, l.stride_z, h->stream);
  h->launches++;
  CK(cudaStreamSynchronize(h->stream));
  return 0;
}

int aphcg_get_history(aphcg_t* h, double* out, int32_t n) {
  if (!h || !out || n < 0) return Fail(APHCG_ERR_ARG, "bad argument");
  if (int rc = SetDevice(h)) return rc;
  n = std::min<int32_t>(n, std::min(h->last_iter, h->hist_cap));
  if (n > 0) CK(cudaMemcpy(out, h->history, sizeof(double) * n, cudaMemcpyDeviceToHost));
  return n;
}

int aphcg_solve(aphcg_t* h, const double* system, const aphcg_layout* system_layout,
                const double* x0, const aphcg_layout* x0_layout, double* x,
                const aphcg_layout* x_layout, const aphcg_conf* conf, aphcg_info* info) {
  if (!h || !system || !x || !conf) return Fail(APHCG_ERR_ARG, "null argument");
  if (int rc = aphcg_upload_system(h, system, system_layout)) return rc;
  if (int rc = aphcg_upload_guess(h, x0, x0_layout)) return rc;
  if (int rc = aphcg_run(h, conf, info)) return rc;
  return aphcg_download_solution(h, x, x_layout);
}

int aphcg_apply(aphcg_t* h, const double* v, const aphcg_layout* v_layout, double* out,
                const aphcg_layout* out_layout) {
  if (!h || !v || !out) return Fail(APHCG_ERR_ARG, "null argument");
  if (!h->have_system) return Fail(APHCG_ERR_STATE, "no system uploaded");
  if (int rc = SetDevice(h)) return rc;
  aphcg_layout lv, lo;
  if (int rc = CheckLayout(h, v_layout, &lv)) return rc;
  if (int rc = CheckLayout(h, out_layout, &lo)) return rc;
  if (int rc = CopyPlanes(h, h->ap, v, lv, true, 0, h->g.nzl, 8, cudaMemcpyHostToDevice, h->stream))
    return rc;
  launch_scatter_field(h->g, h->ap, 0, h->g.cy, h->g.cz, nullptr, h->d.p[1], h->d.p_lo_dst[1],
                       h->d.p_hi_dst[1], h->vx, h->stream);
  if (int rc = StreamBarrier(h)) return rc;
  launch_apply(h->g, h->d, h->vx, h->stream);
  h->launches += 2;
  CK(cudaGetLastError());
  // the neighbours must not scatter their next field into this slab's ghost planes while
  // the operator above still reads them
  if (int rc = StreamBarrier(h)) return rc;
  if (int rc = CopyPlanes(h, out, h->ap, lo, false, 0, h->g.nzl, 8, cudaMemcpyDeviceToHost,
                          h->stream))
    return rc;
  CK(cudaStreamSynchronize(h->stream));
  h->have_guess = false;
  return 0;
}

int aphcg_true_residual(aphcg_t* h, double* sum_r2) {
  if (!h || !sum_r2) return Fail(APHCG_ERR_ARG, "null argument");
  if (!h->have_system) return Fail(APHCG_ERR_STATE, "no system uploaded");
  if (!h->single && (!(h->comm || h->gs) || !h->connected))
    return Fail(APHCG_ERR_STATE, "nranks > 1 needs aphcg_comm_init and aphcg_ipc_connect first");
  if (int rc = SetDevice(h)) return rc;
  // x -> padded field (+ the neighbours' ghost planes), then the stage-"init" kernel
  // (linear.ipp:48-56) on it; its local sum r^2 lands in loc_sum2
  launch_scatter_field(h->g, h->u, 0, h->g.cy, h->g.cz, nullptr, h->d.p[1], h->d.p_lo_dst[1],
                       h->d.p_hi_dst[1], h->vx, h->stream);
  if (int rc = StreamBarrier(h)) return rc;
  launch_init_residual(h->g, h->d, h->vx, h->single, h->precond, h->stream);
  h->launches += 2;
  CK(cudaGetLastError());
  CK(cudaMemcpyAsync(h->h_st, h->st, sizeof(CgState), cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  *sum_r2 = h->h_st->loc_sum2;
  // the neighbours must be done with this slab's ghost planes before anything else lands there
  if (int rc = StreamBarrier(h)) return rc;
  h->have_guess = false;
  return 0;
}

int aphcg_assemble_spheres(aphcg_t* h, const double* spheres, int32_t nspheres, double rho_in,
                           double rho_out, double dt) {
  if (!h || (nspheres > 0 && !spheres) || nspheres < 0) return Fail(APHCG_ERR_ARG, "bad argument");
  if (int rc = SetDevice(h)) return rc;
  double* d_sph = nullptr;
  if (nspheres > 0) {
    if (int rc = EnsureScratch(h, sizeof(double) * 4 * (size_t)nspheres)) return rc;
    d_sph = h->scratch;
    CK(cudaMemcpyAsync(d_sph, spheres, sizeof(double) * 4 * (size_t)nspheres,
                       cudaMemcpyHostToDevice, h->stream));
  }
  int per[3] = {h->desc.periodic[0], h->desc.periodic[1], h->desc.periodic[2]};
  launch_assemble_spheres(h->g, h->d, h->coef, h->rhs, d_sph, nspheres, rho_in, rho_out, dt,
                          h->desc.nz, h->desc.z0, (int)h->desc.nx, (int)h->desc.ny, per, h->stream);
  h->launches += 2;
  cudaError_t e = cudaStreamSynchronize(h->stream);
  if (e != cudaSuccess) return Fail(APHCG_ERR_CUDA, "assemble failed: %s", cudaGetErrorString(e));
  CK(cudaGetLastError());
  h->have_guess = false;
  return SystemResident(h);
}

int aphcg_assemble_projection(aphcg_t* h, const double* rho, const double* vx, const double* vy,
                              const double* vz, const double* source, double dt, double hcell) {
  if (!h || !rho || !vx || !vy || !vz) return Fail(APHCG_ERR_ARG, "null argument");
  if (!(dt > 0) || !(hcell > 0)) return Fail(APHCG_ERR_ARG, "dt and h must be positive");
  if (int rc = SetDevice(h)) return rc;
  const Geom& g = h->g;
  const size_t n_rho = (size_t)g.cz * (g.nzl + 2);
  const size_t n_vx = (size_t)(g.nx + 1) * g.ny * g.nzl;
  const size_t n_vy = (size_t)g.nx * (g.ny + 1) * g.nzl;
  const size_t n_vz = (size_t)g.cz * (g.nzl + 1);
  const size_t n_src = source ? (size_t)g.ncell : 0;
  if (int rc = EnsureScratch(h, sizeof(double) * (n_rho + n_vx + n_vy + n_vz + n_src))) return rc;
  double* buf = h->scratch;
  double* d_rho = buf;
  double* d_vx = d_rho + n_rho;
  double* d_vy = d_vx + n_vx;
  double* d_vz = d_vy + n_vy;
  double* d_src = source ? d_vz + n_vz : nullptr;
  auto up = [&](double* dst, const double* src, size_t n) {
    return cudaMemcpyAsync(dst, src, sizeof(double) * n, cudaMemcpyHostToDevice, h->stream);
  };
  cudaError_t e = up(d_rho, rho, n_rho);
  if (e == cudaSuccess) e = up(d_vx, vx, n_vx);
  if (e == cudaSuccess) e = up(d_vy, vy, n_vy);
  if (e == cudaSuccess) e = up(d_vz, vz, n_vz);
  if (e == cudaSuccess && source) e = up(d_src, source, n_src);
  if (e == cudaSuccess) {
    int per[3] = {h->desc.periodic[0], h->desc.periodic[1], h->desc.periodic[2]};
    launch_assemble_faces(g, h->d, h->coef, h->rhs, d_rho, d_vx, d_vy, d_vz, d_src, dt, hcell,
                          h->desc.cell_volume, h->desc.nz, h->desc.z0, (int)h->desc.nx,
                          (int)h->desc.ny, per, h->stream);
    h->launches += 2;
    e = cudaStreamSynchronize(h->stream);
  }
  if (e != cudaSuccess) return Fail(APHCG_ERR_CUDA, "assemble failed: %s", cudaGetErrorString(e));
  CK(cudaGetLastError());
  h->have_guess = false;
  return SystemResident(h);
}

int aphcg_download_system(aphcg_t* h, double* system, const aphcg_layout* layout) {
  if (!h || !system) return Fail(APHCG_ERR_ARG, "null argument");
  if (!h->have_system) return Fail(APHCG_ERR_STATE, "no system resident");
  if (int rc = SetDevice(h)) return rc;
  aphcg_layout l;
  if (int rc = CheckLayout(h, layout, &l)) return rc;
  double* d_rows = nullptr;
  CK(cudaMalloc(&d_rows, (size_t)h->g.ncell * 64));
  launch_soa_to_rows(h->g, h->d, d_rows, h->stream);
  h->launches++;
  int rc = CopyPlanes(h, system, d_rows, l, false, 0, h->g.nzl, 64, cudaMemcpyDeviceToHost,
                      h->stream);
  cudaError_t e = cudaStreamSynchronize(h->stream);
  cudaFree(d_rows);
  if (rc) return rc;
  if (e != cudaSuccess) return Fail(APHCG_ERR_CUDA, "download failed: %s", cudaGetErrorString(e));
  return 0;
}

// ---- multi-GPU -----------------------------------------------------------------
int aphcg_comm_unique_id(void* id_out) {
  if (!id_out) return Fail(APHCG_ERR_ARG, "null argument");
  if (const char* e = Nccl().Load()) return Fail(APHCG_ERR_COMM, "%s", e);
  ncclUniqueId id;
  NK(Nccl().GetUniqueId(&id));
  memset(id_out, 0, APHCG_UNIQUE_ID_BYTES);
  memcpy(id_out, &id, sizeof(id));
  return 0;
}

int aphcg_comm_init(aphcg_t* h, const void* id) {
  if (!h || !id) return Fail(APHCG_ERR_ARG, "null argument");
  if (h->single) return 0;
  if (const char* e = Nccl().Load()) return Fail(APHCG_ERR_COMM, "%s", e);
  if (int rc = SetDevice(h)) return rc;
  ncclUniqueId uid;
  memcpy(&uid, id, sizeof(uid));
  NK(Nccl().CommInitRank(&h->comm, h->desc.nranks, uid, h->desc.rank));
  return 0;
}

int aphcg_ipc_export(aphcg_t* h, void* blob_out) {
  if (!h || !blob_out) return Fail(APHCG_ERR_ARG, "null argument");
  if (int rc = SetDevice(h)) return rc;
  IpcBlob b{};
  if (!h->gs) CK(cudaIpcGetMemHandle(&b.handle, h->slab));  // in-process groups use plain pointers
  b.ptotal = h->g.ptotal;
  b.pz = h->g.pz;
  b.poff = h->g.poff;
  b.nzl = h->g.nzl;
  b.rank = h->desc.rank;
  b.pid = (int32_t)getpid();
  b.base = (uint64_t)(uintptr_t)h->slab;
  b.device = h->desc.device;
  {
    int dom = 0, bus = 0, dev = 0;
    CK(cudaDeviceGetAttribute(&dom, cudaDevAttrPciDomainId, h->desc.device));
    CK(cudaDeviceGetAttribute(&bus, cudaDevAttrPciBusId, h->desc.device));
    CK(cudaDeviceGetAttribute(&dev, cudaDevAttrPciDeviceId, h->desc.device));
    b.pci = (dom << 16) | (bus << 8) | dev;
  }
  memset(blob_out, 0, APHCG_IPC_BYTES);
  memcpy(blob_out, &b, sizeof(b));
  return 0;
}

int aphcg_ipc_connect(aphcg_t* h, const void* blobs, int32_t count) {
  if (!h) return Fail(APHCG_ERR_ARG, "null handle");
  if (h->single) return 0;
  if (!blobs || count != h->desc.nranks)
    return Fail(APHCG_ERR_ARG, "ipc_connect needs one blob per rank (%d), got %d", h->desc.nranks,
                (int)count);
  if (int rc = SetDevice(h)) return rc;
  const Geom& g = h->g;
  const int n = h->desc.nranks, me = h->desc.rank;
  std::vector<IpcBlob> b(n);
  for (int q = 0; q < n; ++q) {
    memcpy(&b[q], (const char*)blobs + (size_t)q * APHCG_IPC_BYTES, sizeof(IpcBlob));
    if (b[q].rank != q) return Fail(APHCG_ERR_COMM, "blob %d belongs to rank %d", q, b[q].rank);
    if (b[q].pz != g.pz || b[q].poff != g.poff)
      return Fail(APHCG_ERR_COMM, "rank %d has a different xy layout", q);
  }
  for (int q = 0; q < n; ++q) {
    if (q == me) {
      h->peer[q] = h->slab;
    } else if (b[q].pid == (int32_t)getpid()) {  // same process: plain peer pointer
      h->peer[q] = (double*)(uintptr_t)b[q].base;
      if (b[q].device != h->desc.device) {
        int can = 0;
        CK(cudaDeviceCanAccessPeer(&can, h->desc.device, b[q].device));
        if (!can)
          return Fail(APHCG_ERR_COMM, "device %d cannot access device %d's memory (no P2P path)",
                      h->desc.device, b[q].device);
        const cudaError_t e = cudaDeviceEnablePeerAccess(b[q].device, 0);
        if (e == cudaErrorPeerAccessAlreadyEnabled) {
          cudaGetLastError();
        } else if (e != cudaSuccess) {
          return Fail(APHCG_ERR_CUDA, "cudaDeviceEnablePeerAccess(%d) failed: %s", b[q].device,
                      cudaGetErrorString(e));
        }
      }
    } else if (!h->peer[q]) {
      void* ptr = nullptr;
      CK(cudaIpcOpenMemHandle(&ptr, b[q].handle, cudaIpcMemLazyEnablePeerAccess));
      h->peer[q] = (double*)ptr;
      h->peer_opened[q] = true;
    }
  }
  DevPtrs& d = h->d;
  const bool perz = h->desc.periodic[2] != 0;
  const int lo = me > 0 ? me - 1 : (perz ? n - 1 : -1);
  const int hi = me < n - 1 ? me + 1 : (perz ? 0 : -1);
  if (lo >= 0) {  // my bottom plane -> lower neighbour's top ghost plane
    const int64_t off = b[lo].poff + b[lo].nzl * b[lo].pz;
    d.r_lo_dst = h->peer[lo] + off;
    d.p_lo_dst[0] = h->peer[lo] + b[lo].ptotal + off;
    d.p_lo_dst[1] = h->peer[lo] + 2 * b[lo].ptotal + off;
  }
  if (hi >= 0) {  // my top plane -> upper neighbour's bottom ghost plane
    const int64_t off = b[hi].poff - b[hi].pz;
    d.r_hi_dst = h->peer[hi] + off;
    d.p_hi_dst[0] = h->peer[hi] + b[hi].ptotal + off;
    d.p_hi_dst[1] = h->peer[hi] + 2 * b[hi].ptotal + off;
  }
  // mailboxes live behind the three padded fields of every rank's slab
  for (int q = 0; q < n; ++q)
    d.cm.box[q] = reinterpret_cast<MailSlot*>(h->peer[q] + 3 * b[q].ptotal);
  d.cm.rank = me;
  d.cm.nranks = n;
  d.cm.use_mail = h->use_mail ? 1 : 0;
  // Who waits for the all-reduced scalars.  Default: one-warp k_finish_* kernels.  Opt-in
  // (APHCG_WAIT=kernel): the consumer kernels' own CTAs, i.e. two launches per iteration --
  // measured on 2 B200 (profiles/r02_wait_modes_2gpu.txt) it is no faster (64-plane slabs:
  // 0.3995 vs 0.3969 ms/iteration; 512^3 per GPU: 2.574 vs 2.550), because every CTA then pays
  // a system-scope fence before it may read the neighbours' planes; the hand-off cost is the
  // cross-GPU latency itself, not the extra launches.  Never when two slabs share a GPU: a
  // grid of spinning CTAs would keep the other slab's producer kernel off the SMs for good.
  bool shared_gpu = false;
  for (int q = 0; q < n; ++q)
    for (int w = q + 1; w < n; ++w) shared_gpu |= (b[q].pci == b[w].pci);
  h->wait_in_kernel = false;
  if (const char* ew = getenv("APHCG_WAIT"))
    h->wait_in_kernel = !strcmp(ew, "kernel") && h->use_mail && !shared_gpu;
  d.cm.wait_in_kernel = h->wait_in_kernel ? 1 : 0;
  d.cm.reader_fence = 1;
  if (const char* ef = getenv("APHCG_WAIT_FENCE")) d.cm.reader_fence = atoi(ef) != 0;  // measurements only
  {
    const char* et = getenv("APHCG_MAIL_TIMEOUT_MS");
    const double ms = et ? atof(et) : 20000.0;
    d.cm.timeout_ns = (unsigned long long)((ms > 1.0 ? ms : 1.0) * 1e6);
  }
  h->connected = true;
  InvalidateGraphs(h);
  return 0;
}

int aphcg_timer_start(aphcg_t* h) {
  if (!h) return Fail(APHCG_ERR_ARG, "null handle");
  if (int rc = SetDevice(h)) return rc;
  CK(cudaEventRecord(h->ev[3], h->stream));
  return 0;
}

int aphcg_timer_stop(aphcg_t* h, double* ms) {
  if (!h || !ms) return Fail(APHCG_ERR_ARG, "null argument");
  if (int rc = SetDevice(h)) return rc;
  cudaEvent_t e;
  CK(cudaEventCreate(&e));
  CK(cudaEventRecord(e, h->stream));
  CK(cudaEventSynchronize(e));
  float f = 0;
  cudaError_t err = cudaEventElapsedTime(&f, h->ev[3], e);
  cudaEventDestroy(e);
  if (err != cudaSuccess) return Fail(APHCG_ERR_CUDA, "timer: %s", cudaGetErrorString(err));
  *ms = f;
  return 0;
}

int aphcg_profile_kernels(aphcg_t* h, int32_t iters, double* ms_dir_spmv, double* ms_update) {
  if (!h || iters < 1 || !ms_dir_spmv || !ms_update) return Fail(APHCG_ERR_ARG, "bad argument");
  aphcg_conf conf{0.0, 0, iters + 1};
  if (int rc = CheckRun(h, &conf)) return rc;
  if (!h->single) return Fail(APHCG_ERR_STATE, "kernel profiling is single-rank only");
  if (int rc = SetDevice(h)) return rc;
  if (!h->have_guess) {
    if (int rc = aphcg_upload_guess(h, nullptr, nullptr)) return rc;
  }
  if (int rc = EnsureHistory(h, conf.maxiter)) return rc;
  if (int rc = WriteState(h, &conf)) return rc;
  CK(cudaMemsetAsync(h->d.p[0], 0, sizeof(double) * (size_t)h->g.ptotal, h->stream));
  launch_init_residual(h->g, h->d, h->vx, true, h->precond, h->stream);
  std::vector<cudaEvent_t> ev(3 * (size_t)iters);
  for (auto& e : ev) CK(cudaEventCreate(&e));
  for (int i = 0; i < iters; ++i) {
    CK(cudaEventRecord(ev[3 * i], h->stream));
    if (h->use_tma && h->tma2 && h->sym) {
      launch_dir_spmv_stream(h->tma2, h->g, h->d, true, h->stream);
    } else if (h->use_tma) {
      launch_dir_spmv_tma(h->tma, h->g, h->d, true, h->sym, h->stream);
    } else {
      launch_dir_spmv_plain(h->g, h->d, h->vx, true, h->stream);
    }
    CK(cudaEventRecord(ev[3 * i + 1], h->stream));
    launch_update(h->g, h->d, h->vx, true, h->precond, h->stream);
    CK(cudaEventRecord(ev[3 * i + 2], h->stream));
  }
  CK(cudaStreamSynchronize(h->stream));
  CK(cudaGetLastError());
  double sd = 0, su = 0, sd_par[2] = {0, 0};
  int n_par[2] = {0, 0};
  for (int i = 0; i < iters; ++i) {
    float a = 0, b = 0;
    CK(cudaEventElapsedTime(&a, ev[3 * i], ev[3 * i + 1]));
    CK(cudaEventElapsedTime(&b, ev[3 * i + 1], ev[3 * i + 2]));
    sd += a;
    su += b;
    sd_par[i & 1] += a;
    n_par[i & 1]++;
  }
  *ms_dir_spmv = sd / iters;
  *ms_update = su / iters;
  if (getenv("APHCG_VERBOSE")) {
    float tot = 0;
    cudaEventElapsedTime(&tot, ev[0], ev[3 * (size_t)iters - 1]);
    fprintf(stderr, "aphcg profile: %d iterations, dir %.4f ms (even %.4f, odd %.4f) + upd %.4f ms = %.4f; "
            "span/iter %.4f ms\n",
            iters, sd / iters, sd_par[0] / std::max(n_par[0], 1), sd_par[1] / std::max(n_par[1], 1),
            su / iters, (sd + su) / iters, tot / iters);
  }
  for (auto& e : ev) cudaEventDestroy(e);
  h->launches += 1 + 2 * (int64_t)iters;
  h->have_guess = false;
  return 0;
}

int aphcg_describe(aphcg_t* h, char* buf, int32_t buflen) {
  if (!h || !buf || buflen < 1) return Fail(APHCG_ERR_ARG, "bad argument");
  char t[200] = "";
  if (h->use_tma && h->tma2 && h->sym) {
    tma2_plan_describe(h->tma2, t, sizeof(t));
  } else if (h->use_tma) {
    tma_plan_describe(h->tma, t, sizeof(t));
  }
  if (h->persist) {
    snprintf(buf, buflen,
             "loop=persistent-cooperative ctas=%u planes_per_tile=%d update_rows=%d (two grid "
             "barriers per iteration, plain-load stencil) precond=none allreduce=none",
             h->pplan.grid, h->pplan.zc, h->pplan.ur);
    return 0;
  }
  snprintf(buf, buflen, "spmv=%s%s %s precond=%s graph=%d allreduce=%s", h->use_tma ? "tma" : "plain",
           h->use_tma ? (h->sym ? (h->tma2 ? "-stream-sym4" : "-sym4") : "-gen7") : "", t,
           h->precond ? "jacobi" : "none",
           h->use_graph ? 1 : 0,
           h->single ? "none"
                     : (h->use_mail ? (h->wait_in_kernel ? "peer-mailbox/in-kernel-wait"
                                                         : "peer-mailbox/finish-kernels")
                                    : "nccl"));
  return 0;
}

void* aphcg_stream(aphcg_t* h) { return h ? (void*)h->stream : nullptr; }
int64_t aphcg_launch_count(aphcg_t* h) { return h ? h->launches : 0; }
int aphcg_launches_per_iter(aphcg_t* h) { return h ? (h->persist ? 0 : LaunchesPerIter(h)) : 0; }

}  // extern "C"

namespace acg {
void AttachGroupSync(aphcg* h, GroupSync* gs) {
  h->gs = gs;
  h->use_mail = true;
}
void SetLastError(const char* msg) { g_error = msg ? msg : ""; }
}  // namespace acg
