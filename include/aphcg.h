/* aphcg -- B200-native conjugate-gradient solver for aphros' 7-point
 * FieldCell<Expr> pressure systems.  C ABI of libaphcg.so.
 *
 * This is the drop-in boundary: everything the reference's
 *   linear::Solver<M>::Solve / SetConf / GetConf      (src/linear/linear.h:15-57)
 *   linear::ModuleLinear<M>::Make                      (src/linear/linear.h:59-76)
 * needs from a device module, as plain pointers and sizes.  The aphros-side
 * adapter that binds it (ModuleLinear "conjugate_cuda") is
 * aphros_b200/plugin/linear_conjugate_cuda.cpp; INTEGRATION.md shows the wiring.
 *
 * The product path is CUDA (sm_100a) only: every compute entry point fails with
 * APHCG_ERR_CUDA when no device is usable.  There is no CPU fallback.
 *
 * Conventions
 *   - cells are indexed x fastest, then y, then z (src/geom/block.h:149-158);
 *   - a system row is 8 doubles [c, x-, x+, y-, y+, z-, z+, const] meaning
 *       e0*x[c] + sum_q e[1+q]*x[nb_q(c)] + e7 = 0
 *     (src/geom/mesh.h:484-485, src/linear/linear.h:34-44);
 *   - all functions return 0 on success, a negative APHCG_ERR_* otherwise, and
 *     leave a message for aphcg_last_error() (thread local).
 */
#ifndef APHCG_H_
#define APHCG_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define APHCG_VERSION 1

enum {
  APHCG_OK = 0,
  APHCG_ERR_ARG = -1,    /* bad argument */
  APHCG_ERR_CUDA = -2,   /* CUDA runtime/driver error, or no device */
  APHCG_ERR_COMM = -3,   /* NCCL / peer-memory error */
  APHCG_ERR_STATE = -4   /* call order (e.g. run before upload) */
};

/* aphcg_desc.flags */
enum {
  APHCG_MAXNORM = 1u << 0,   /* residual = max|r|/V  (Extra::residual_max,
                                src/linear/linear.ipp:103-107,245-249) */
  APHCG_NO_GRAPH = 1u << 1,  /* launch kernels one by one instead of replaying
                                a CUDA graph (debugging) */
  APHCG_NO_TMA = 1u << 2,    /* use the plain-load stencil kernel instead of the
                                TMA-staged one (debugging / comparison) */
  APHCG_NO_SYM = 1u << 3,    /* always stream all 7 coefficient arrays, even when
                                the resident matrix is verified symmetric */
  APHCG_NCCL_REDUCE = 1u << 4, /* multi-GPU: all-reduce the scalars with NCCL instead of
                                the peer-memory mailboxes (comparison baseline) */
  APHCG_JACOBI_PRECOND = 1u << 5, /* OPT-IN, not the reference's recurrence: conjugate
                                gradients preconditioned with diag(A) (z = r/e0,
                                alpha = r.z/p.Ap, beta = r.z_new/r.z, p = z + beta p);
                                residual and exit rule unchanged (norm of r).  The
                                reference's SolverConjugate is unpreconditioned, so
                                iteration counts differ unless diag(A) is constant. */
  APHCG_NO_STREAM = 1u << 7,  /* symmetric storage: keep the direction kernel whose coefficient
                                streams go HBM -> registers (k_dir_spmv_tma) instead of the
                                one that stages every operand in shared memory by TMA */
  APHCG_NO_PERSISTENT = 1u << 6 /* never run the loop as ONE persistent cooperative kernel.
                                By default a single-GPU solve of at most 700 000 cells (about
                                88^3) does: same arithmetic, two grid-wide
                                barriers per iteration instead of two launches. */
};

typedef struct aphcg aphcg_t;

/* Geometry of the solve.  The domain is cut into z-slabs, one per rank (one
 * rank = one GPU = one process); rank r owns the contiguous planes
 * [z0, z0+nz_local).  nranks == 1: z0 = 0, nz_local = nz. */
typedef struct {
  int64_t nx, ny, nz;    /* global inner cells */
  int32_t periodic[3];   /* m.flags.is_periodic (src/distr/distr.ipp:55-59); a
                            neighbour outside a non-periodic boundary contributes
                            the value 0 (every reference assembler zeroes that
                            coefficient) */
  double cell_volume;    /* m.GetCellSize().prod() (src/geom/mesh.h:171-173) */
  int32_t device;        /* CUDA device ordinal of this rank */
  int32_t rank, nranks;
  int64_t z0, nz_local;
  uint32_t flags;
} aphcg_desc;

/* Where inner cell (i,j,k) of THIS RANK'S slab lives in a caller array:
 * element index = offset + i + j*stride_y + k*stride_z (elements are doubles
 * for scalar fields, 8-double rows for the system).  A compact slab is
 * {0, nx, nx*ny}; a reference FieldCell with halos (src/geom/mesh.ipp:60-113)
 * is {hl*(1+sy+sz), nx+2*hl+1, sy*(ny+2*hl+1)}.  NULL layout = compact. */
typedef struct {
  int64_t offset, stride_y, stride_z;
} aphcg_layout;

/* linear::Solver<M>::Conf (src/linear/linear.h:21-25) */
typedef struct {
  double tol;
  int32_t miniter;
  int32_t maxiter;
} aphcg_conf;

/* linear::Solver<M>::Info (src/linear/linear.h:27-30) plus timings */
typedef struct {
  double residual;
  int32_t iter;
  int32_t reserved;
  double loop_ms;      /* device time of the CG loop alone (CUDA events) */
  double total_ms;     /* device time of everything the call enqueued */
  double residual0;    /* sqrt(sum r0^2 / V) of the initial residual r0 = -(A x0 + e7)
                          (linear.ipp:48-56): lets a caller express a relative
                          tolerance, which the reference's Conf cannot; 0 after
                          aphcg_run_jacobi, which forms no residual */
} aphcg_info;

const char* aphcg_last_error(void);
int aphcg_version(void);
/* number of usable CUDA devices (0 if none); never fails */
int aphcg_device_count(void);

int aphcg_create(aphcg_t** out, const aphcg_desc* desc);
int aphcg_destroy(aphcg_t* h);

/* Pinned host memory for callers that stage fields themselves (the adapter's
 * rank-wide shared fields): host<->device copies from it run at full link speed. */
int aphcg_host_alloc(void** out, uint64_t bytes);
int aphcg_host_free(void* p);

/* ---- one call = one linear::Solver::Solve (host buffers in, host buffer out) --
 * system: rows of this rank's slab; x0: initial guess or NULL (zero guess,
 * linear.ipp:43-47); x: solution out (may alias x0, linear.h:40). */
int aphcg_solve(
    aphcg_t* h, const double* system, const aphcg_layout* system_layout,
    const double* x0, const aphcg_layout* x0_layout, double* x,
    const aphcg_layout* x_layout, const aphcg_conf* conf, aphcg_info* info);

/* ---- the same, split so that fields can stay resident in HBM ---------------- */
int aphcg_upload_system(aphcg_t* h, const double* system, const aphcg_layout* layout);
int aphcg_upload_guess(aphcg_t* h, const double* x0, const aphcg_layout* layout);
/* device-resident inputs (device pointers; same row/field formats) */
int aphcg_set_system_device(aphcg_t* h, const double* d_system, const aphcg_layout* layout);
int aphcg_set_guess_device(aphcg_t* h, const double* d_x0, const aphcg_layout* layout);
/* initial residual + CG loop + final update of x; fields stay on the device */
int aphcg_run(aphcg_t* h, const aphcg_conf* conf, aphcg_info* info);
int aphcg_download_solution(aphcg_t* h, double* x, const aphcg_layout* layout);
int aphcg_get_solution_device(aphcg_t* h, double* d_x, const aphcg_layout* layout);
/* residual after each completed iteration of the last run (n <= iter) */
int aphcg_get_history(aphcg_t* h, double* out, int32_t n);

/* SolverJacobi twin (src/linear/linear.ipp:152-237) on the same resident system */
int aphcg_run_jacobi(aphcg_t* h, const aphcg_conf* conf, aphcg_info* info);

/* out = A*v for the resident system (host v/out, slab-local, compact or laid
 * out): the stage-"iter" operator alone, for operator-level parity tests. */
int aphcg_apply(aphcg_t* h, const double* v, const aphcg_layout* v_layout,
                double* out, const aphcg_layout* out_layout);

/* sum over this rank's cells of r^2, r = -(A x + e7) RECOMPUTED from the resident solution
 * of the last run (the loop itself only carries the recursively updated residual, like
 * linear.ipp:89,102-107).  With nranks > 1 it is a collective call (the neighbours' boundary
 * planes of x are exchanged) and the caller adds the ranks' values.  For checking a converged
 * solve at sizes where the solution cannot be compared on the host. */
int aphcg_true_residual(aphcg_t* h, double* sum_r2);

/* Synthetic variable-density projection system assembled on the device from a
 * sphere list (formulas of src/solver/proj.ipp:343-398):
 * spheres = nspheres x {cx,cy,cz,r}.  Replaces upload_system for benchmarks
 * whose system would not fit comfortably in host memory. */
int aphcg_assemble_spheres(
    aphcg_t* h, const double* spheres, int32_t nspheres, double rho_in,
    double rho_out, double dt);
/* Device-side assembly of the projection system from what the fluid solver holds
 * BEFORE it builds rows: what Proj::GetFlux + GetFluxSum compute
 * (src/solver/proj.ipp:343-383) on a uniform mesh without embedded boundaries, with
 * wall (zero pressure-gradient coefficient) or periodic domain faces:
 *   k_f = h*dt/rho_f,  rho_f = harmonic mean of the two cell densities,
 *   e0 = sum_f k_f,  e[1+q] = -k_f(q),  e7 = sum_q outward(q)*v_f(q) - source*V.
 * Host arrays of this rank's slab, compact, x fastest:
 *   rho (nz_local+2, ny, nx): planes -1..nz_local (neighbour slabs' boundary planes);
 *   vx (nz_local, ny, nx+1), vy (nz_local, ny+1, nx), vz (nz_local+1, ny, nx): volume
 *   fluxes through the faces;  source (nz_local, ny, nx) or NULL.
 * 32-40 bytes per cell cross the host link instead of the 64 of assembled rows. */
int aphcg_assemble_projection(
    aphcg_t* h, const double* rho, const double* vx, const double* vy, const double* vz,
    const double* source, double dt, double hcell);
/* copy the resident system back as rows (for checking the assembler) */
int aphcg_download_system(aphcg_t* h, double* system, const aphcg_layout* layout);

/* ---- multi-GPU (nranks > 1): one process per GPU -------------------------------
 * Residual halo planes are written straight into the neighbour's ghost planes
 * over NVLink, and the two scalars per iteration are all-reduced through small
 * "mailboxes" in every rank's memory, written by the kernel that finishes the
 * local reduction (peer memory again; NCCL is used only at the start of a
 * solve, outside the loop: for the initial residual and as a barrier).  Wiring, once after create:
 *   1. rank 0: aphcg_comm_unique_id(id); broadcast id to all ranks (the host
 *      harness does this with torch.distributed / MPI / a file);
 *   2. every rank: aphcg_comm_init(h, id)              (collective)
 *   3. every rank: aphcg_ipc_export(h, mine); all-gather the 128-byte blobs;
 *   4. every rank: aphcg_ipc_connect(h, blobs, nranks)  (blobs ordered by rank). */
#define APHCG_UNIQUE_ID_BYTES 128
#define APHCG_IPC_BYTES 128
int aphcg_comm_unique_id(void* id_out);
int aphcg_comm_init(aphcg_t* h, const void* id);
int aphcg_ipc_export(aphcg_t* h, void* blob_out);
int aphcg_ipc_connect(aphcg_t* h, const void* blobs, int32_t count);

/* ---- in-process slab group: ONE process drives several GPUs ----------------------
 * For callers that own the whole rank-wide arrays in one address space -- aphros
 * started without MPI on a multi-GPU node ("single process, G devices,
 * driven from the lead block of a single aphros rank").  The group cuts the domain of
 * `desc` (rank, nranks, z0, nz_local and device are ignored) into ndevices contiguous
 * z-slabs, one per entry of devices[] (sizes differ by at most one plane, larger
 * slabs first), creates an ordinary handle per slab and wires them with plain peer
 * pointers (cudaDeviceEnablePeerAccess).  Every group call runs one host thread per
 * slab; the loop is the multi-GPU loop above (halo planes and scalars through peer
 * memory, written by the kernels), and NCCL is not used at all: the two collective
 * steps of a solve are a stream synchronize plus a thread barrier.  A device ordinal
 * may appear more than once in devices[] (its slabs then share that GPU; meant for
 * testing the slab path on a single-GPU machine).
 * Arrays are RANK-WIDE: layouts address global inner cell (i,j,k) as
 * offset + i + j*stride_y + k*stride_z; NULL = compact (nz, ny, nx).
 * After any group call fails the group is unusable; destroy it. */
typedef struct aphcg_group aphcg_group_t;
int aphcg_group_create(aphcg_group_t** out, const aphcg_desc* desc, const int32_t* devices,
                       int32_t ndevices);
int aphcg_group_destroy(aphcg_group_t* g);
int aphcg_group_size(aphcg_group_t* g);
/* the handle of one slab (owned by the group), e.g. for aphcg_describe / timers */
aphcg_t* aphcg_group_member(aphcg_group_t* g, int32_t slab);
int aphcg_group_slab(aphcg_group_t* g, int32_t slab, int64_t* z0, int64_t* nz_local);
/* the group forms of aphcg_solve / upload / run / download (same meaning) */
int aphcg_group_solve(
    aphcg_group_t* g, const double* system, const aphcg_layout* system_layout,
    const double* x0, const aphcg_layout* x0_layout, double* x,
    const aphcg_layout* x_layout, const aphcg_conf* conf, aphcg_info* info);
int aphcg_group_upload_system(aphcg_group_t* g, const double* system, const aphcg_layout* layout);
int aphcg_group_upload_guess(aphcg_group_t* g, const double* x0, const aphcg_layout* layout);
int aphcg_group_run(aphcg_group_t* g, const aphcg_conf* conf, aphcg_info* info);
int aphcg_group_run_jacobi(aphcg_group_t* g, const aphcg_conf* conf, aphcg_info* info);
int aphcg_group_download_solution(aphcg_group_t* g, double* x, const aphcg_layout* layout);
int aphcg_group_assemble_spheres(
    aphcg_group_t* g, const double* spheres, int32_t nspheres, double rho_in, double rho_out,
    double dt);
/* sum over ALL slabs (added in slab order) */
int aphcg_group_true_residual(aphcg_group_t* g, double* sum_r2);
/* aphcg_assemble_projection for the group.  Arrays are RANK-WIDE and compact:
 * rho (nz+2, ny, nx) with one ghost plane below and above the domain (the periodic image, or
 * anything finite at a wall), vx (nz, ny, nx+1), vy (nz, ny+1, nx), vz (nz+1, ny, nx),
 * source (nz, ny, nx) or NULL. */
int aphcg_group_assemble_projection(
    aphcg_group_t* g, const double* rho, const double* vx, const double* vy, const double* vz,
    const double* source, double dt, double hcell);

/* Device-side timing on the handle's stream (CUDA events): start records an
 * event; stop records another, waits for it and returns the milliseconds between. */
int aphcg_timer_start(aphcg_t* h);
int aphcg_timer_stop(aphcg_t* h, double* ms);
/* Per-kernel timing of the loop on the resident system: restarts from the
 * resident guess, runs `iters` iterations launched one by one with an event
 * around every kernel, and returns the average duration of the direction+SpMV
 * kernel and of the update kernel (milliseconds).  Leaves no valid solution. */
int aphcg_profile_kernels(aphcg_t* h, int32_t iters, double* ms_dir_spmv, double* ms_update);

/* One-line description of the kernel configuration in use (for benchmark
 * records): stencil kernel variant, tile, prefetch distance, graph, all-reduce. */
int aphcg_describe(aphcg_t* h, char* buf, int32_t buflen);

/* cudaStream_t the handle enqueues on (as void*), for callers that time with
 * their own events */
void* aphcg_stream(aphcg_t* h);
/* number of kernels this handle has launched so far (graph replays counted per node) */
int64_t aphcg_launch_count(aphcg_t* h);
/* number of kernel launches per CG iteration of the current configuration */
int aphcg_launches_per_iter(aphcg_t* h);

#ifdef __cplusplus
}
#endif
#endif /* APHCG_H_ */
