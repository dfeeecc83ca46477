// Optional capability of a linear solver module: solve the projection step's pressure system
// from what the fluid solver holds BEFORE it assembles rows -- the cell density, the face
// volume fluxes of the predicted velocity and the volume source -- instead of from a
// FieldCell<Expr>.  conjugate_cuda implements it by assembling the rows on
// the GPU (aphcg_assemble_projection: bit for bit what Proj::GetFlux + GetFluxSum produce,
// src/solver/proj.ipp:343-383), so that 32-40 bytes per cell cross the host link instead of 64
// and the host never writes the 64 B/cell system at all.
//
// A caller that has these fields asks its solver for the capability and falls back to the
// usual path otherwise -- e.g. in Proj<EB>::Imp::Project (src/solver/proj.ipp:386-398):
//
//   if (auto* ps = dynamic_cast<linear::ProjectionSolver<M>*>(linsolver_.get())) {
//     ps->SolveProjection(*owner_->fcr_, ffv.GetFieldFace(), owner_->fcsv_, dt, &fcp, fcp, m);
//   } else {
//     ctx->ffvc = GetFlux(ffv, dt); ctx->fcpcs = GetFluxSum(ctx->ffvc, *owner_->fcsv_); ...
//     linsolver_->Solve(ctx->fcpcs, &fcp, fcp, m);
//   }
//
// Preconditions (the caller's to check): uniform 3-D mesh without embedded boundaries; every
// non-periodic domain face a wall or another non-pressure boundary condition (zero
// pressure-gradient coefficient, proj.ipp:352-354); no cell conditions.
#pragma once

#include "linear/linear.h"

namespace linear {

template <class M>
class ProjectionSolver {
 public:
  using Scal = typename M::Scal;
  using Info = typename Solver<M>::Info;
  virtual ~ProjectionSolver() = default;
  // fc_dens: cell density with valid halos (as Comm leaves them); ff_flux: volume flux through
  // every face of the block; fc_source: volume source or nullptr; fc_init / fc_sol as in
  // Solver<M>::Solve.  A stage coroutine like Solve: call it from a nested stage.
  virtual Info SolveProjection(
      const FieldCell<Scal>& fc_dens, const FieldFace<Scal>& ff_flux,
      const FieldCell<Scal>* fc_source, Scal dt, const FieldCell<Scal>* fc_init,
      FieldCell<Scal>& fc_sol, M& m) = 0;
};

} // namespace linear
