/* TEST INFRASTRUCTURE -- the parity checker, never the product path.
 *
 * CPU restatement, on flat arrays, of the reference's unpreconditioned CG
 *   linear::SolverConjugate<M>::Imp::Solve   (src/linear/linear.ipp:24-129)
 * and its point-Jacobi sibling
 *   linear::SolverJacobi<M>::Imp::Solve      (src/linear/linear.ipp:152-237).
 *
 * Pinned (tests/test_oracle_vs_reference.py, tests/golden/): bit-for-bit equal
 * to the reference built from /root/reference/src (oracle/_ref/ref_cg) for
 * single- and multi-block meshes, iteration counts and residuals included.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this library.
 */
#ifndef CG_ORACLE_H_
#define CG_ORACLE_H_

#ifdef __cplusplus
extern "C" {
#endif

typedef struct {
  long nx, ny, nz;       /* inner cells, x fastest (src/geom/block.h:149-158) */
  int periodic[3];       /* wrap neighbours per direction; a non-periodic
                            out-of-domain neighbour contributes value 0
                            (its coefficient is 0 in every reference assembler,
                            SURVEY.md appendix B) */
  long bsx, bsy, bsz;    /* block shape used ONLY to reproduce the reference's
                            summation order: per-block partial sums in x-fastest
                            order, blocks appended x-fastest to a zero
                            (src/distr/distr.ipp:143-169, native.ipp:90-97);
                            0 => one block */
  double cell_volume;    /* m.GetCellSize().prod()  (src/geom/mesh.h:171-173) */
  double tol;            /* Conf::tol      (src/linear/linear.h:21-25) */
  int miniter;           /* Conf::miniter */
  int maxiter;           /* Conf::maxiter */
  int maxnorm;           /* Extra::residual_max (src/linear/linear.ipp:245-249) */
} cg_oracle_desc;

/* sys: nx*ny*nz rows of 8 doubles [c,x-,x+,y-,y+,z-,z+,const]
 *      (src/geom/mesh.h:484-485, src/linear/linear.h:34-44).
 * x0 : initial guess or NULL (zero guess, linear.ipp:43-47).
 * x  : out, solution. history: NULL or max(maxiter,miniter)+2 doubles, residual after each
 *      completed iteration.  Returns 0, or -1 on allocation failure. */
int cg_oracle_conjugate(
    const cg_oracle_desc* d, const double* sys, const double* x0, double* x,
    double* residual, int* iter, double* history);

/* OPT-IN mode of the CUDA module, NOT a reference algorithm ("parity unpinned":
 * the reference has no preconditioned CG to pin it against): conjugate gradients
 * preconditioned with diag(A), written in the reference's style (same stages,
 * same +1e-100 guards, same block-ordered sums, same residual norm and exit rule):
 *   z = r/e0;  alpha = (r.z)/(p.Ap + 1e-100);  beta = (r.z)_new/((r.z)_old + 1e-100);
 *   p = z + beta p. */
int cg_oracle_pconjugate(
    const cg_oracle_desc* d, const double* sys, const double* x0, double* x,
    double* residual, int* iter, double* history);

int cg_oracle_jacobi(
    const cg_oracle_desc* d, const double* sys, const double* x0, double* x,
    double* residual, int* iter, double* history);

/* out = A*v (no constant term): the "iter" stage operator, linear.ipp:65-72 */
int cg_oracle_apply(
    const cg_oracle_desc* d, const double* sys, const double* v, double* out);

#ifdef __cplusplus
}
#endif
#endif
