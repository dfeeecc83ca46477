// Host-callable launch wrappers of the CG kernels (cg_kernels.cu, cg_spmv_tma.cu).
#pragma once

#include <cuda_runtime.h>

#include "cg_types.h"

namespace acg {

// vx: cells per thread along x (2 = 128-bit accesses, needs even nx; else 1).
// single: this rank is the only one -- the kernel that finishes a reduction also
// advances the loop state; otherwise an all-reduce and a k_finish_* follow.
unsigned tile_blocks(const Geom& g, int vx);
// CTAs of the tiled kernels for the current g.zc
unsigned tile_blocks_for(const Geom& g, int vx);

void launch_dir_spmv_plain(const Geom& g, const DevPtrs& d, int vx, bool single, cudaStream_t s);
void launch_update(const Geom& g, const DevPtrs& d, int vx, bool single, bool precond,
                   cudaStream_t s);
void launch_finish_dir(const DevPtrs& d, cudaStream_t s);
void launch_finish_upd(const DevPtrs& d, cudaStream_t s);
void launch_finish_init(const DevPtrs& d, cudaStream_t s);
void launch_finish_jacobi(const DevPtrs& d, cudaStream_t s);
void launch_init_residual(const Geom& g, const DevPtrs& d, int vx, bool single, bool precond,
                          cudaStream_t s);
void launch_apply(const Geom& g, const DevPtrs& d, int vx, cudaStream_t s);
void launch_scatter_field(const Geom& g, const double* src, int64_t off, int64_t sy, int64_t sz,
                          double* u, double* fpad, double* lo_dst, double* hi_dst, int vx,
                          cudaStream_t s);
void launch_gather_field(const Geom& g, const double* u, double* dst, int64_t off, int64_t sy,
                         int64_t sz, cudaStream_t s);
void launch_final_update(const Geom& g, const DevPtrs& d, int vx, cudaStream_t s);
void launch_rows_to_soa(const Geom& g, const double* rows, int64_t off, int64_t sy, int64_t sz,
                        int k0, int nk, double* const* a, double* rhs, cudaStream_t s);
void launch_soa_to_rows(const Geom& g, const DevPtrs& d, double* rows, cudaStream_t s);
void launch_jacobi(const Geom& g, const DevPtrs& d, int vx, bool single, cudaStream_t s);

// Persistent small-mesh loop (k_cg_persistent, cg_kernels.cu): one cooperative launch runs
// iterations until the exit rule fires or max_iters are done.
struct PersistPlan {
  dim3 tgrid;     // tiles of the direction stage
  int zc;         // planes per tile
  unsigned grid;  // CTAs (all co-resident)
  int ur;         // rows per thread of the update stage
};
bool persistent_plan(const Geom& g, int vx, PersistPlan* out);
cudaError_t launch_cg_persistent(const Geom& g, const DevPtrs& d, int vx, const PersistPlan& p,
                                 int max_iters, cudaStream_t s);

// TMA-staged direction+SpMV kernel (cg_spmv_tma.cu)
struct TmaPlan;  // opaque: tensor maps + launch geometry
TmaPlan* tma_plan_create(const Geom& g, const DevPtrs& d, char* err, int errlen);
void tma_plan_destroy(TmaPlan* p);
unsigned tma_plan_blocks(const TmaPlan* p);
void tma_plan_describe(const TmaPlan* p, char* buf, int buflen);
// sym: read 4 coefficient streams instead of 7 (matrix verified symmetric)
void launch_dir_spmv_tma(const TmaPlan* p, const Geom& g, const DevPtrs& d, bool single, bool sym,
                         cudaStream_t s);
// The same stage with EVERY operand staged in shared memory by TMA (cg_spmv_tma2.cu);
// symmetric storage only.
struct Tma2Plan;
Tma2Plan* tma2_plan_create(const Geom& g, const DevPtrs& d, char* err, int errlen);
void tma2_plan_destroy(Tma2Plan* p);
unsigned tma2_plan_blocks(const Tma2Plan* p);
void tma2_plan_describe(const Tma2Plan* p, char* buf, int buflen);
void launch_dir_spmv_stream(const Tma2Plan* p, const Geom& g, const DevPtrs& d, bool single,
                            cudaStream_t s);
// sets *flag (device int) to 1 if any off-diagonal pair differs: a2[c] != a1[c+1],
// a4[c] != a3[c+row], a6[c] != a5[c+plane] (inside this slab)
void launch_check_symmetry(const Geom& g, const DevPtrs& d, int* flag, cudaStream_t s);

// device-side synthetic assembly (cg_assemble.cu)
void launch_assemble_spheres(const Geom& g, const DevPtrs& d, double* const* a, double* rhs,
                             const double* spheres, int nspheres, double rho_in, double rho_out,
                             double dt, int64_t nz_global, int64_t z0, int nx_g, int ny_g,
                             const int* periodic, cudaStream_t s);

// general device-side assembly from a cell density and face fluxes (cg_assemble.cu)
void launch_assemble_faces(const Geom& g, const DevPtrs& d, double* const* a, double* rhs,
                           const double* rho_in, const double* vx, const double* vy,
                           const double* vz, const double* src, double dt, double h, double vol,
                           int64_t nz_global, int64_t z0, int nx_g, int ny_g, const int* periodic,
                           cudaStream_t s);

}  // namespace acg
