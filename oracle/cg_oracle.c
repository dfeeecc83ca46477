/* TEST INFRASTRUCTURE -- the parity checker, never the product path.
 * See cg_oracle.h.  Plain C, strict IEEE double arithmetic: compile with
 * -ffp-contract=off and without -ffast-math so that every sum is formed in
 * the order the reference forms it (oracle/Makefile does).
 */
#include "cg_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

/* Fields carry one ghost layer, like the reference's FieldCell after
 * Comm(direct_one); px = nx+2 etc.  Index of inner cell (i,j,k): */
typedef struct {
  long nx, ny, nz, px, py, pz, sy, sz, ntot;
  long bsx, bsy, bsz;
  int per[3];
} geom_t;

static geom_t make_geom(const cg_oracle_desc* d) {
  geom_t g;
  g.nx = d->nx; g.ny = d->ny; g.nz = d->nz;
  g.px = d->nx + 2; g.py = d->ny + 2; g.pz = d->nz + 2;
  g.sy = g.px; g.sz = g.px * g.py; g.ntot = g.px * g.py * g.pz;
  g.bsx = d->bsx > 0 ? d->bsx : d->nx;
  g.bsy = d->bsy > 0 ? d->bsy : d->ny;
  g.bsz = d->bsz > 0 ? d->bsz : d->nz;
  g.per[0] = d->periodic[0]; g.per[1] = d->periodic[1]; g.per[2] = d->periodic[2];
  return g;
}

static inline long gi(const geom_t* g, long i, long j, long k) {
  return (k + 1) * g->sz + (j + 1) * g->sy + (i + 1);
}

/* Halo fill = what m.Comm(&f, direct_one) leaves in the 6 face layers:
 * periodic images (src/distr/comm_manager_seq.ipp:21-23); zero where the
 * domain is not periodic (value is irrelevant there: coefficient 0). */
static void fill_ghosts(const geom_t* g, double* f) {
  long i, j, k;
  for (k = 0; k < g->nz; ++k)
    for (j = 0; j < g->ny; ++j) {
      f[gi(g, -1, j, k)] = g->per[0] ? f[gi(g, g->nx - 1, j, k)] : 0.;
      f[gi(g, g->nx, j, k)] = g->per[0] ? f[gi(g, 0, j, k)] : 0.;
    }
  for (k = 0; k < g->nz; ++k)
    for (i = 0; i < g->nx; ++i) {
      f[gi(g, i, -1, k)] = g->per[1] ? f[gi(g, i, g->ny - 1, k)] : 0.;
      f[gi(g, i, g->ny, k)] = g->per[1] ? f[gi(g, i, 0, k)] : 0.;
    }
  for (j = 0; j < g->ny; ++j)
    for (i = 0; i < g->nx; ++i) {
      f[gi(g, i, j, -1)] = g->per[2] ? f[gi(g, i, j, g->nz - 1)] : 0.;
      f[gi(g, i, j, g->nz)] = g->per[2] ? f[gi(g, i, j, 0)] : 0.;
    }
}

/* Row of the operator at inner cell (i,j,k), neighbour order q=0..5 =
 * x-,x+,y-,y+,z-,z+ (src/geom/mesh.h:253-255,294-296); the accumulation
 * order is the reference's: centre first, then q ascending
 * (linear.ipp:50-53 and :67-70). `start` is e7 for the residual, absent for A*p. */
static inline double row_apply(
    const geom_t* g, const double* e, const double* f, long c, int with_const) {
  double s = f[c] * e[0];
  if (with_const) s += e[7];
  s += f[c - 1] * e[1];
  s += f[c + 1] * e[2];
  s += f[c - g->sy] * e[3];
  s += f[c + g->sy] * e[4];
  s += f[c - g->sz] * e[5];
  s += f[c + g->sz] * e[6];
  return s;
}

/* Visits blocks x-fastest, and cells x-fastest inside a block
 * (src/distr/native.ipp:90-97, src/geom/rangein.h:25-36). */
#define FOR_BLOCKS(g)                                   \
  for (long b2 = 0; b2 < (g)->nz; b2 += (g)->bsz)       \
    for (long b1 = 0; b1 < (g)->ny; b1 += (g)->bsy)     \
      for (long b0 = 0; b0 < (g)->nx; b0 += (g)->bsx)
#define FOR_CELLS_IN_BLOCK(g)                                          \
  for (long k = b2; k < b2 + (g)->bsz && k < (g)->nz; ++k)             \
    for (long j = b1; j < b1 + (g)->bsy && j < (g)->ny; ++j)           \
      for (long i = b0; i < b0 + (g)->bsx && i < (g)->nx; ++i)

static double* padded_from_compact(const geom_t* g, const double* src) {
  double* f = (double*)calloc((size_t)g->ntot, sizeof(double));
  if (!f) return NULL;
  if (src) {
    for (long k = 0; k < g->nz; ++k)
      for (long j = 0; j < g->ny; ++j)
        memcpy(f + gi(g, 0, j, k), src + (k * g->ny + j) * g->nx,
               (size_t)g->nx * sizeof(double));
  }
  return f;
}

static void compact_from_padded(const geom_t* g, const double* f, double* dst) {
  for (long k = 0; k < g->nz; ++k)
    for (long j = 0; j < g->ny; ++j)
      memcpy(dst + (k * g->ny + j) * g->nx, f + gi(g, 0, j, k),
             (size_t)g->nx * sizeof(double));
}

static inline const double* row_of(const geom_t* g, const double* sys, long i,
                                   long j, long k) {
  return sys + 8 * ((k * g->ny + j) * g->nx + i);
}

int cg_oracle_apply(const cg_oracle_desc* d, const double* sys, const double* v,
                    double* out) {
  geom_t g = make_geom(d);
  double* f = padded_from_compact(&g, v);
  if (!f) return -1;
  fill_ghosts(&g, f);
  for (long k = 0; k < g.nz; ++k)
    for (long j = 0; j < g.ny; ++j)
      for (long i = 0; i < g.nx; ++i)
        out[(k * g.ny + j) * g.nx + i] =
            row_apply(&g, row_of(&g, sys, i, j, k), f, gi(&g, i, j, k), 0);
  free(f);
  return 0;
}

int cg_oracle_conjugate(const cg_oracle_desc* d, const double* sys,
                        const double* x0, double* x, double* residual,
                        int* iter_out, double* history) {
  geom_t g = make_geom(d);
  /* linear.ipp:28-40: u, r, p, lp */
  double* u = padded_from_compact(&g, x0); /* zero guess if x0 == NULL (:43-47) */
  double* r = (double*)calloc((size_t)g.ntot, sizeof(double));
  double* p = (double*)calloc((size_t)g.ntot, sizeof(double));
  double* lp = (double*)calloc((size_t)g.ntot, sizeof(double));
  if (!u || !r || !p || !lp) {
    free(u); free(r); free(p); free(lp);
    return -1;
  }
  /* the caller's guess arrives with valid halos (SURVEY.md 3.2 note) */
  fill_ghosts(&g, u);

  /* stage "init" (:42-58): r = -(A u + e7); Comm(r) */
  FOR_BLOCKS(&g) {
    FOR_CELLS_IN_BLOCK(&g) {
      const long c = gi(&g, i, j, k);
      r[c] = -row_apply(&g, row_of(&g, sys, i, j, k), u, c, 1);
    }
  }
  fill_ghosts(&g, r);
  /* stage "init" (:59-62): p = r */
  memcpy(p, r, (size_t)g.ntot * sizeof(double));

  int iter = 0;
  double res = 0;
  for (;;) {
    /* stage "iter" (:64-82): lp = A p; two block-ordered sums */
    double dot_r_prev = 0, dot_p_lp = 0;
    FOR_BLOCKS(&g) {
      double br = 0, bp = 0;
      FOR_CELLS_IN_BLOCK(&g) {
        const long c = gi(&g, i, j, k);
        lp[c] = row_apply(&g, row_of(&g, sys, i, j, k), p, c, 0);
      }
      FOR_CELLS_IN_BLOCK(&g) {
        const long c = gi(&g, i, j, k);
        br += r[c] * r[c];
        bp += p[c] * lp[c];
      }
      dot_r_prev += br;
      dot_p_lp += bp;
    }
    /* stage "iter2" (:83-95) */
    const double alpha = dot_r_prev / (dot_p_lp + 1e-100);
    double dot_r = 0;
    double max_r = -1.7976931348623157e308; /* OpMax::Neutral, reduce.h:86-88 */
    FOR_BLOCKS(&g) {
      double br = 0, bm = 0;
      FOR_CELLS_IN_BLOCK(&g) {
        const long c = gi(&g, i, j, k);
        u[c] += alpha * p[c];
        r[c] -= alpha * lp[c];
        br += r[c] * r[c];
        const double a = fabs(r[c]);
        bm = bm > a ? bm : a;
      }
      dot_r += br;
      max_r = max_r > bm ? max_r : bm;
    }
    /* stage "iter3" (:96-101): p = r + beta p; Comm(p) */
    const double beta = dot_r / (dot_r_prev + 1e-100);
    FOR_BLOCKS(&g) {
      FOR_CELLS_IN_BLOCK(&g) {
        const long c = gi(&g, i, j, k);
        p[c] = r[c] + beta * p[c];
      }
    }
    fill_ghosts(&g, p);
    /* stage "check" (:102-114) */
    if (d->maxnorm) {
      res = max_r / d->cell_volume;
    } else {
      res = sqrt(dot_r / d->cell_volume);
    }
    if (history) history[iter] = res;
    ++iter;
    if (iter >= d->miniter && (iter > d->maxiter || res < d->tol)) break;
  }
  /* stage "result" (:116-118) */
  compact_from_padded(&g, u, x);
  *residual = res;
  *iter_out = iter;
  free(u); free(r); free(p); free(lp);
  return 0;
}

int cg_oracle_pconjugate(const cg_oracle_desc* d, const double* sys,
                         const double* x0, double* x, double* residual,
                         int* iter_out, double* history) {
  geom_t g = make_geom(d);
  double* u = padded_from_compact(&g, x0);
  double* r = (double*)calloc((size_t)g.ntot, sizeof(double));
  double* z = (double*)calloc((size_t)g.ntot, sizeof(double));
  double* p = (double*)calloc((size_t)g.ntot, sizeof(double));
  double* lp = (double*)calloc((size_t)g.ntot, sizeof(double));
  if (!u || !r || !z || !p || !lp) {
    free(u); free(r); free(z); free(p); free(lp);
    return -1;
  }
  fill_ghosts(&g, u);
  double dot_rz = 0;
  FOR_BLOCKS(&g) {
    double b = 0;
    FOR_CELLS_IN_BLOCK(&g) {
      const long c = gi(&g, i, j, k);
      const double* e = row_of(&g, sys, i, j, k);
      r[c] = -row_apply(&g, e, u, c, 1);
      z[c] = r[c] / e[0];
      b += r[c] * z[c];
    }
    dot_rz += b;
  }
  fill_ghosts(&g, z);
  memcpy(p, z, (size_t)g.ntot * sizeof(double));
  int iter = 0;
  double res = 0;
  for (;;) {
    double dot_p_lp = 0;
    FOR_BLOCKS(&g) {
      double bp = 0;
      FOR_CELLS_IN_BLOCK(&g) {
        const long c = gi(&g, i, j, k);
        lp[c] = row_apply(&g, row_of(&g, sys, i, j, k), p, c, 0);
      }
      FOR_CELLS_IN_BLOCK(&g) {
        const long c = gi(&g, i, j, k);
        bp += p[c] * lp[c];
      }
      dot_p_lp += bp;
    }
    const double alpha = dot_rz / (dot_p_lp + 1e-100);
    double dot_rz_new = 0, dot_r = 0;
    double max_r = -1.7976931348623157e308;
    FOR_BLOCKS(&g) {
      double bz = 0, br = 0, bm = 0;
      FOR_CELLS_IN_BLOCK(&g) {
        const long c = gi(&g, i, j, k);
        const double* e = row_of(&g, sys, i, j, k);
        u[c] += alpha * p[c];
        r[c] -= alpha * lp[c];
        z[c] = r[c] / e[0];
        bz += r[c] * z[c];
        br += r[c] * r[c];
        const double a = fabs(r[c]);
        bm = bm > a ? bm : a;
      }
      dot_rz_new += bz;
      dot_r += br;
      max_r = max_r > bm ? max_r : bm;
    }
    const double beta = dot_rz_new / (dot_rz + 1e-100);
    dot_rz = dot_rz_new;
    FOR_BLOCKS(&g) {
      FOR_CELLS_IN_BLOCK(&g) {
        const long c = gi(&g, i, j, k);
        p[c] = z[c] + beta * p[c];
      }
    }
    fill_ghosts(&g, p);
    res = d->maxnorm ? max_r / d->cell_volume : sqrt(dot_r / d->cell_volume);
    if (history) history[iter] = res;
    ++iter;
    if (iter >= d->miniter && (iter > d->maxiter || res < d->tol)) break;
  }
  compact_from_padded(&g, u, x);
  *residual = res;
  *iter_out = iter;
  free(u); free(r); free(z); free(p); free(lp);
  return 0;
}

int cg_oracle_jacobi(const cg_oracle_desc* d, const double* sys,
                     const double* x0, double* x, double* residual,
                     int* iter_out, double* history) {
  geom_t g = make_geom(d);
  double* u = padded_from_compact(&g, x0);
  double* un = (double*)calloc((size_t)g.ntot, sizeof(double));
  double* const buf0 = u;  /* u and un swap every iteration; free the allocations by name */
  double* const buf1 = un;
  if (!u || !un) {
    free(u); free(un);
    return -1;
  }
  fill_ghosts(&g, u);
  int iter = 0;
  double res = 0;
  for (;;) {
    /* stage "iter" (linear.ipp:178-194) */
    double maxdiff = -1.7976931348623157e308;
    FOR_BLOCKS(&g) {
      double bm = 0;
      FOR_CELLS_IN_BLOCK(&g) {
        const long c = gi(&g, i, j, k);
        const double* e = row_of(&g, sys, i, j, k);
        double nd = e[7];
        nd += u[c - 1] * e[1];
        nd += u[c + 1] * e[2];
        nd += u[c - g.sy] * e[3];
        nd += u[c + g.sy] * e[4];
        nd += u[c - g.sz] * e[5];
        nd += u[c + g.sz] * e[6];
        un[c] = -nd / e[0];
        const double a = fabs(un[c] - u[c]);
        bm = bm > a ? bm : a;
      }
      maxdiff = maxdiff > bm ? maxdiff : bm;
    }
    { double* t = u; u = un; un = t; }
    fill_ghosts(&g, u);
    /* stage "check" (:195-204) */
    res = maxdiff;
    if (history) history[iter] = res;
    ++iter;
    if (iter >= d->miniter && (iter > d->maxiter || res < d->tol)) break;
  }
  compact_from_padded(&g, u, x);
  *residual = res;
  *iter_out = iter;
  free(buf0);
  free(buf1);
  return 0;
}
