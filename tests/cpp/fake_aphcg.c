/* TEST DOUBLE -- never part of the product, never shipped in aphros_b200/.
 *
 * tests/test_adapter_cpu.py preloads this object (LD_PRELOAD) in front of libaphcg.so so that
 * the aphros adapter (aphros_b200/plugin/linear_conjugate_cuda.cpp) can be driven by the
 * reference's own classes on a machine WITHOUT a GPU: what is under test is the adapter's
 * host logic -- the stage coroutine, block -> rank-wide gather/scatter, the geometry, flags,
 * Conf and device list it hands to the C ABI -- not any arithmetic.  The entry points the
 * adapter calls are answered by the CPU oracle (oracle/cg_oracle.c, test infrastructure) and
 * every call is logged to $FAKE_APHCG_LOG for the test to inspect.  With $FAKE_APHCG_TEE=<prefix>
 * and $FAKE_APHCG_TEE_INDEX=<n> the n-th solve's inputs are also written to <prefix>.sys /
 * <prefix>.x0 (raw float64): how a test captures a live system of the application (SURVEY.md 8d,
 * input S1) to replay it through the real library.
 *
 * The real library has no CPU path: without a CUDA device aphcg_group_create fails.
 */
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../../include/aphcg.h"
#include "../../oracle/cg_oracle.h"

struct aphcg_group {
  aphcg_desc desc;
  int ndev;
  double* rows; /* resident copies (upload/run split) */
  double* x0;
  double* x;
  int have_guess;
};

static void logf_(const char* fmt, ...) {
  const char* path = getenv("FAKE_APHCG_LOG");
  if (!path) return;
  FILE* f = fopen(path, "a");
  if (!f) return;
  va_list ap;
  va_start(ap, fmt);
  vfprintf(f, fmt, ap);
  va_end(ap);
  fputc('\n', f);
  fclose(f);
}

static int g_solves = 0;

static void tee(const aphcg_group_t* g, const double* system, const double* x0);

static size_t ncell(const aphcg_group_t* g) {
  return (size_t)g->desc.nx * g->desc.ny * g->desc.nz;
}

static cg_oracle_desc odesc(const aphcg_group_t* g, const aphcg_conf* c) {
  cg_oracle_desc d;
  memset(&d, 0, sizeof(d));
  d.nx = g->desc.nx;
  d.ny = g->desc.ny;
  d.nz = g->desc.nz;
  for (int i = 0; i < 3; ++i) d.periodic[i] = g->desc.periodic[i];
  d.cell_volume = g->desc.cell_volume;
  d.tol = c->tol;
  d.miniter = c->miniter;
  d.maxiter = c->maxiter;
  d.maxnorm = (g->desc.flags & APHCG_MAXNORM) ? 1 : 0;
  return d;
}

const char* aphcg_last_error(void) { return "fake_aphcg: error"; }

int aphcg_host_alloc(void** out, uint64_t bytes) {
  *out = malloc(bytes ? bytes : 8);
  return *out ? 0 : APHCG_ERR_CUDA;
}
int aphcg_host_free(void* p) {
  free(p);
  return 0;
}

int aphcg_group_create(aphcg_group_t** out, const aphcg_desc* desc, const int32_t* devices,
                       int32_t ndevices) {
  aphcg_group_t* g = (aphcg_group_t*)calloc(1, sizeof(*g));
  g->desc = *desc;
  g->ndev = ndevices;
  char dev[256] = "";
  for (int i = 0; i < ndevices && i < 32; ++i) sprintf(dev + strlen(dev), "%s%d", i ? "," : "", devices[i]);
  logf_("create nx=%lld ny=%lld nz=%lld periodic=%d%d%d volume=%.17g flags=%u devices=[%s]",
        (long long)desc->nx, (long long)desc->ny, (long long)desc->nz, desc->periodic[0],
        desc->periodic[1], desc->periodic[2], desc->cell_volume, desc->flags, dev);
  *out = g;
  return 0;
}

int aphcg_group_destroy(aphcg_group_t* g) {
  if (!g) return 0;
  logf_("destroy");
  free(g->rows);
  free(g->x0);
  free(g->x);
  free(g);
  return 0;
}

int aphcg_group_solve(aphcg_group_t* g, const double* system, const aphcg_layout* ls,
                      const double* x0, const aphcg_layout* l0, double* x, const aphcg_layout* lx,
                      const aphcg_conf* conf, aphcg_info* info) {
  if (ls || l0 || lx) return APHCG_ERR_ARG; /* the adapter passes compact rank-wide arrays */
  tee(g, system, x0);
  const cg_oracle_desc d = odesc(g, conf);
  /* x may alias x0 (linear.h:40): the oracle reads the guess before it writes x */
  double* guess = NULL;
  if (x0 && !getenv("FAKE_APHCG_NOSOLVE")) {
    guess = (double*)malloc(sizeof(double) * ncell(g));
    memcpy(guess, x0, sizeof(double) * ncell(g));
  }
  double res = 0;
  int it = 0;
  int rc = 0;
  if (getenv("FAKE_APHCG_NOSOLVE")) { /* timing of the adapter alone: leave x as it is */
    it = 1;
  } else {
    rc = cg_oracle_conjugate(&d, system, guess, x, &res, &it, NULL);
  }
  free(guess);
  logf_("solve guess=%d tol=%.17g miniter=%d maxiter=%d -> iter=%d residual=%.17g", x0 ? 1 : 0,
        conf->tol, conf->miniter, conf->maxiter, it, res);
  memset(info, 0, sizeof(*info));
  info->residual = res;
  info->iter = it;
  return rc;
}

int aphcg_group_upload_system(aphcg_group_t* g, const double* system, const aphcg_layout* l) {
  if (l) return APHCG_ERR_ARG;
  free(g->rows);
  g->rows = (double*)malloc(sizeof(double) * 8 * ncell(g));
  memcpy(g->rows, system, sizeof(double) * 8 * ncell(g));
  logf_("upload_system");
  return 0;
}

int aphcg_group_upload_guess(aphcg_group_t* g, const double* x0, const aphcg_layout* l) {
  if (l) return APHCG_ERR_ARG;
  free(g->x0);
  g->x0 = NULL;
  if (x0) {
    g->x0 = (double*)malloc(sizeof(double) * ncell(g));
    memcpy(g->x0, x0, sizeof(double) * ncell(g));
  }
  logf_("upload_guess guess=%d", x0 ? 1 : 0);
  return 0;
}

/* Rows from density + face fluxes in the reference's order of operations (what
 * aphcg_assemble_projection computes on the device; aphros_b200/systems.py:projection_rows). */
int aphcg_group_assemble_projection(aphcg_group_t* g, const double* rho, const double* vx,
                                    const double* vy, const double* vz, const double* source,
                                    double dt, double hcell) {
  const long nx = g->desc.nx, ny = g->desc.ny, nz = g->desc.nz;
  const double vol = g->desc.cell_volume, inv_h = 1.0 / hcell, area = vol / hcell;
  free(g->rows);
  g->rows = (double*)malloc(sizeof(double) * 8 * ncell(g));
  for (long k = 0; k < nz; ++k)
    for (long j = 0; j < ny; ++j)
      for (long i = 0; i < nx; ++i) {
        const long nn[3] = {nx, ny, nz}, w[3] = {i, j, k};
        double* e = g->rows + 8 * ((k * ny + j) * nx + i);
        const double inv_c = 1.0 / rho[((k + 1) * ny + j) * nx + i];
        double diag = 0.0;
        for (int q = 0; q < 6; ++q) {
          const int d = q / 2, up = q % 2;
          long v[3] = {i, j, k};
          v[d] += up ? 1 : -1;
          const int outside = v[d] < 0 || v[d] >= nn[d];
          double a = 0.0;
          if (!(outside && !g->desc.periodic[d])) {
            if (d < 2) v[d] = (v[d] + nn[d]) % nn[d]; /* z: the ghost planes hold the images */
            const double inv_n = 1.0 / rho[((v[2] + 1) * ny + v[1]) * nx + v[0]];
            const double rf = 1.0 / ((inv_c + inv_n) * 0.5);
            a = inv_h * ((area / rf) * dt);
          }
          (void)w;
          e[1 + q] = -a;
          diag = diag + a;
        }
        e[0] = diag;
        double e7 = -vx[(k * ny + j) * (nx + 1) + i] + vx[(k * ny + j) * (nx + 1) + i + 1];
        e7 = e7 - vy[(k * (ny + 1) + j) * nx + i];
        e7 = e7 + vy[(k * (ny + 1) + j + 1) * nx + i];
        e7 = e7 - vz[(k * ny + j) * nx + i];
        e7 = e7 + vz[((k + 1) * ny + j) * nx + i];
        e7 = e7 - (source ? source[(k * ny + j) * nx + i] : 0.0) * vol;
        e[7] = e7;
      }
  logf_("assemble_projection dt=%.17g h=%.17g source=%d", dt, hcell, source ? 1 : 0);
  return 0;
}

int aphcg_group_run(aphcg_group_t* g, const aphcg_conf* conf, aphcg_info* info) {
  if (!g->rows) return APHCG_ERR_STATE;
  const cg_oracle_desc d = odesc(g, conf);
  free(g->x);
  g->x = (double*)malloc(sizeof(double) * ncell(g));
  double res = 0;
  int it = 0;
  const int rc = cg_oracle_conjugate(&d, g->rows, g->x0, g->x, &res, &it, NULL);
  logf_("run tol=%.17g maxiter=%d -> iter=%d residual=%.17g", conf->tol, conf->maxiter, it, res);
  memset(info, 0, sizeof(*info));
  info->residual = res;
  info->iter = it;
  return rc;
}

int aphcg_group_run_jacobi(aphcg_group_t* g, const aphcg_conf* conf, aphcg_info* info) {
  if (!g->rows) return APHCG_ERR_STATE;
  const cg_oracle_desc d = odesc(g, conf);
  free(g->x);
  g->x = (double*)malloc(sizeof(double) * ncell(g));
  double res = 0;
  int it = 0;
  const int rc = cg_oracle_jacobi(&d, g->rows, g->x0, g->x, &res, &it, NULL);
  logf_("run_jacobi tol=%.17g maxiter=%d -> iter=%d residual=%.17g", conf->tol, conf->maxiter, it,
        res);
  memset(info, 0, sizeof(*info));
  info->residual = res;
  info->iter = it;
  return rc;
}

int aphcg_group_download_solution(aphcg_group_t* g, double* x, const aphcg_layout* l) {
  if (l || !g->x) return APHCG_ERR_ARG;
  memcpy(x, g->x, sizeof(double) * ncell(g));
  logf_("download_solution");
  return 0;
}

static void tee(const aphcg_group_t* g, const double* system, const double* x0) {
  const char* prefix = getenv("FAKE_APHCG_TEE");
  const char* index = getenv("FAKE_APHCG_TEE_INDEX");
  const int n = g_solves++;
  if (!prefix || !index || atoi(index) != n) return;
  char path[1024];
  snprintf(path, sizeof(path), "%s.sys", prefix);
  FILE* f = fopen(path, "wb");
  if (f) {
    fwrite(system, sizeof(double), 8 * ncell(g), f);
    fclose(f);
  }
  snprintf(path, sizeof(path), "%s.x0", prefix);
  f = fopen(path, "wb");
  if (f) {
    if (x0) fwrite(x0, sizeof(double), ncell(g), f);
    fclose(f);
  }
  logf_("tee solve %d -> %s.sys", n, prefix);
}
