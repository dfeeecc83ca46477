// Device-side assembly of the synthetic variable-density projection systems
// S2..S4 of SURVEY.md 8(d) (used by benchmarks whose rows would not fit in host
// memory, and as a first step towards assembling the pressure system where it is
// solved).  Restates the reference's face formulas (src/solver/proj.ipp:343-398):
//   a_f = h*dt/rho_f,  rho_f = 2/(1/rho_- + 1/rho_+),  zero through non-periodic
//   domain faces;  e0 = sum_f a_f,  e[1+q] = -a_f(q),
//   e7 = sum_q outward(q)*v_f,  v_f = (u.n_f)*h^2,
//   u = (sin(pi x) cos(2pi y), sin(pi y) cos(2pi z), sin(pi z) cos(2pi x)).
// Arithmetic is written with explicit round-to-nearest intrinsics so that the
// sphere classification matches aphros_b200/systems.py bit for bit.
#include "cg_kernels.cuh"
#include "cg_launch.h"

namespace acg {

namespace {

struct AsmPar {
  int nx_g, ny_g;
  int64_t nz_g, z0;
  int per[3];
  double h, rho_in, rho_out, dt;
  int nspheres;
};

// density at global cell (i,j,k), periodic wrap where the domain is periodic
__device__ double rho_at(const AsmPar& P, const double* __restrict__ sph, int i, int j,
                         int64_t k) {
  if (P.per[0]) i = (i + P.nx_g) % P.nx_g;
  if (P.per[1]) j = (j + P.ny_g) % P.ny_g;
  if (P.per[2]) k = (k + P.nz_g) % P.nz_g;
  const double x = __dmul_rn((double)i + 0.5, P.h);
  const double y = __dmul_rn((double)j + 0.5, P.h);
  const double z = __dmul_rn((double)k + 0.5, P.h);
  bool inside = false;
  for (int s = 0; s < P.nspheres; ++s) {
    const double dx = __dsub_rn(x, sph[4 * s + 0]);
    const double dy = __dsub_rn(y, sph[4 * s + 1]);
    const double dz = __dsub_rn(z, sph[4 * s + 2]);
    const double r = sph[4 * s + 3];
    const double d2 = __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
    inside = inside || (d2 < __dmul_rn(r, r));
  }
  return inside ? P.rho_in : P.rho_out;
}

// pass 1: density into a padded field (ghost layers included)
__global__ void k_density(const Geom g, const AsmPar P, const double* __restrict__ sph,
                          double* rho) {
  const int64_t nxy = (int64_t)(g.nx + 2) * (g.ny + 2);
  const int64_t n = nxy * (g.nzl + 2);
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < n;
       t += (int64_t)gridDim.x * blockDim.x) {
    const int i = (int)(t % (g.nx + 2)) - 1;
    const int j = (int)((t / (g.nx + 2)) % (g.ny + 2)) - 1;
    const int k = (int)(t / nxy) - 1;
    rho[g.poff + i + (int64_t)j * g.py + (int64_t)k * g.pz] = rho_at(P, sph, i, j, P.z0 + k);
  }
}

// pass 2: rows from the padded density
__global__ void k_rows_from_density(const Geom g, const AsmPar P, const double* __restrict__ rho,
                                    double* a0, double* a1, double* a2, double* a3, double* a4,
                                    double* a5, double* a6, double* rhs) {
  const double two_pi = 6.283185307179586476925286766559;
  const double hh = __dmul_rn(P.h, P.h);
  for (int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; c < g.ncell;
       c += (int64_t)gridDim.x * blockDim.x) {
    const int i = (int)(c % g.nx);
    const int j = (int)((c / g.nx) % g.ny);
    const int k = (int)(c / g.cz);
    const int64_t kg = P.z0 + k;
    const int64_t ip = g.poff + i + (int64_t)j * g.py + (int64_t)k * g.pz;
    const double rc = rho[ip];
    const int64_t off[6] = {-1, 1, -g.py, g.py, -g.pz, g.pz};
    const bool wall[6] = {!P.per[0] && i == 0,          !P.per[0] && i == P.nx_g - 1,
                          !P.per[1] && j == 0,          !P.per[1] && j == P.ny_g - 1,
                          !P.per[2] && kg == 0,         !P.per[2] && kg == P.nz_g - 1};
    double a[6];
    double diag = 0.0;
#pragma unroll
    for (int q = 0; q < 6; ++q) {
      const double rn = rho[ip + off[q]];
      // harmonic mean with the lower cell first, as systems.py forms it
      const double rlo = (q & 1) ? rc : rn, rhi = (q & 1) ? rn : rc;
      const double rf = __ddiv_rn(2.0, __dadd_rn(__ddiv_rn(1.0, rlo), __ddiv_rn(1.0, rhi)));
      a[q] = wall[q] ? 0.0 : __ddiv_rn(__dmul_rn(P.h, P.dt), rf);
      diag = __dadd_rn(diag, a[q]);
    }
    a0[c] = diag;
    a1[c] = -a[0];
    a2[c] = -a[1];
    a3[c] = -a[2];
    a4[c] = -a[3];
    a5[c] = -a[4];
    a6[c] = -a[5];
    // e7: divergence of the face-normal fluxes (zero flux through the walls)
    const double pi = 3.141592653589793238462643383279;
    const double xc = __dmul_rn((double)i + 0.5, P.h), yc = __dmul_rn((double)j + 0.5, P.h);
    const double zc = __dmul_rn((double)kg + 0.5, P.h);
    auto flux = [&](int64_t f, int64_t nf, int per, double ct) {
      // face f of nf+1 along one direction, tangential factor ct
      if (!per && (f == 0 || f == nf)) return 0.0;
      if (per && f == nf) f = 0;
      return __dmul_rn(__dmul_rn(sin(__dmul_rn(pi, __dmul_rn((double)f, P.h))), ct), hh);
    };
    const double cy = cos(two_pi * yc), cz = cos(two_pi * zc), cx = cos(two_pi * xc);
    const double dx = __dsub_rn(flux(i + 1, P.nx_g, P.per[0], cy), flux(i, P.nx_g, P.per[0], cy));
    const double dy = __dsub_rn(flux(j + 1, P.ny_g, P.per[1], cz), flux(j, P.ny_g, P.per[1], cz));
    const double dz = __dsub_rn(flux(kg + 1, P.nz_g, P.per[2], cx), flux(kg, P.nz_g, P.per[2], cx));
    rhs[c] = __dadd_rn(__dadd_rn(dx, dy), dz);
  }
}

// ---- general variant: density and face fluxes supplied by the caller ---------------
// rho_in: (nzl+2, ny, nx) cell densities of planes -1..nzl of this slab (the two extra
// planes are the z-neighbours' boundary planes; ignored at a non-periodic domain face)
__global__ void k_pad_density(const Geom g, const AsmPar P, const double* __restrict__ rho_in,
                              double* rho) {
  const int64_t nxy = (int64_t)(g.nx + 2) * (g.ny + 2);
  const int64_t n = nxy * (g.nzl + 2);
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < n;
       t += (int64_t)gridDim.x * blockDim.x) {
    int i = (int)(t % (g.nx + 2)) - 1;
    int j = (int)((t / (g.nx + 2)) % (g.ny + 2)) - 1;
    const int k = (int)(t / nxy) - 1;
    const int64_t ip = g.poff + i + (int64_t)j * g.py + (int64_t)k * g.pz;
    // wrap (periodic) or clamp (wall: the value is multiplied by a zero coefficient)
    i = P.per[0] ? (i + g.nx) % g.nx : min(max(i, 0), g.nx - 1);
    j = P.per[1] ? (j + g.ny) % g.ny : min(max(j, 0), g.ny - 1);
    rho[ip] = rho_in[i + (int64_t)j * g.nx + (int64_t)(k + 1) * g.cz];
  }
}

// Rows from the padded density and the caller's face fluxes:
//   vx (nzl, ny, nx+1), vy (nzl, ny+1, nx), vz (nzl+1, ny, nx); src (nzl, ny, nx) or null.
// The order of operations is the reference's own, so that the rows are bit for bit what
// Proj::GetFlux + GetFluxSum produce (src/solver/proj.ipp:343-383; pinned by
// oracle/_ref/ref_assemble, see aphros_b200/systems.py:projection_rows):
//   rho_f = 1 / ((1/rho_+ + 1/rho_-) * 0.5)        InterpolateHarmonic, approx_eb.h:351-363
//   k_f   = (1/h) * (((V/h) / rho_f) * dt)          GradientImplicit [-1/h, 1/h] scaled by
//                                                   -area/rho_f*dt, area = V/h (mesh.ipp:91)
//   e0 = sum_q k_f(q) in the order q = 0..5,  e[1+q] = -k_f(q)      AppendExpr, mesh.h:575-579
//   e7 = (((((-v_x- + v_x+) - v_y-) + v_y+) - v_z-) + v_z+) - src*V   proj.ipp:377-379
__global__ void k_rows_from_faces(const Geom g, const AsmPar P, const double* __restrict__ rho,
                                  const double* __restrict__ vx, const double* __restrict__ vy,
                                  const double* __restrict__ vz, const double* __restrict__ src,
                                  double vol, double* a0, double* a1, double* a2, double* a3,
                                  double* a4, double* a5, double* a6, double* rhs) {
  const double inv_h = __ddiv_rn(1.0, P.h);
  const double area = __ddiv_rn(vol, P.h);
  for (int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; c < g.ncell;
       c += (int64_t)gridDim.x * blockDim.x) {
    const int i = (int)(c % g.nx);
    const int j = (int)((c / g.nx) % g.ny);
    const int k = (int)(c / g.cz);
    const int64_t kg = P.z0 + k;
    const int64_t ip = g.poff + i + (int64_t)j * g.py + (int64_t)k * g.pz;
    const double inv_c = __ddiv_rn(1.0, rho[ip]);
    const int64_t off[6] = {-1, 1, -g.py, g.py, -g.pz, g.pz};
    const bool wall[6] = {!P.per[0] && i == 0,          !P.per[0] && i == P.nx_g - 1,
                          !P.per[1] && j == 0,          !P.per[1] && j == P.ny_g - 1,
                          !P.per[2] && kg == 0,         !P.per[2] && kg == P.nz_g - 1};
    double a[6];
    double diag = 0.0;
#pragma unroll
    for (int q = 0; q < 6; ++q) {
      const double inv_n = __ddiv_rn(1.0, rho[ip + off[q]]);
      const double rf = __ddiv_rn(1.0, __dmul_rn(__dadd_rn(inv_c, inv_n), 0.5));
      a[q] = wall[q] ? 0.0 : __dmul_rn(inv_h, __dmul_rn(__ddiv_rn(area, rf), P.dt));
      diag = __dadd_rn(diag, a[q]);
    }
    a0[c] = diag;
    a1[c] = -a[0];
    a2[c] = -a[1];
    a3[c] = -a[2];
    a4[c] = -a[3];
    a5[c] = -a[4];
    a6[c] = -a[5];
    const int64_t fx = i + (int64_t)j * (g.nx + 1) + (int64_t)k * (g.nx + 1) * g.ny;
    const int64_t fy = i + (int64_t)j * g.nx + (int64_t)k * g.nx * (g.ny + 1);
    double e7 = __dadd_rn(-vx[fx], vx[fx + 1]);
    e7 = __dsub_rn(e7, vy[fy]);
    e7 = __dadd_rn(e7, vy[fy + g.nx]);
    e7 = __dsub_rn(e7, vz[c]);
    e7 = __dadd_rn(e7, vz[c + g.cz]);
    e7 = __dsub_rn(e7, __dmul_rn(src ? src[c] : 0.0, vol));
    rhs[c] = e7;
  }
}

}  // namespace

void launch_assemble_faces(const Geom& g, const DevPtrs& d, double* const* a, double* rhs,
                           const double* rho_in, const double* vx, const double* vy,
                           const double* vz, const double* src, double dt, double h, double vol,
                           int64_t nz_global, int64_t z0, int nx_g, int ny_g, const int* periodic,
                           cudaStream_t s) {
  AsmPar P;
  P.nx_g = nx_g;
  P.ny_g = ny_g;
  P.nz_g = nz_global;
  P.z0 = z0;
  for (int i = 0; i < 3; ++i) P.per[i] = periodic[i];
  P.h = h;
  P.rho_in = P.rho_out = 0.0;
  P.dt = dt;
  P.nspheres = 0;
  // scratch: the padded residual field (see launch_assemble_spheres)
  k_pad_density<<<kNumSMs * 8, 256, 0, s>>>(g, P, rho_in, d.r);
  k_rows_from_faces<<<kNumSMs * 8, 256, 0, s>>>(g, P, d.r, vx, vy, vz, src, vol, a[0], a[1], a[2],
                                            a[3], a[4], a[5], a[6], rhs);
  cudaMemsetAsync(d.r, 0, sizeof(double) * (size_t)g.ptotal, s);
}

void launch_assemble_spheres(const Geom& g, const DevPtrs& d, double* const* a, double* rhs,
                             const double* spheres, int nspheres, double rho_in, double rho_out,
                             double dt, int64_t nz_global, int64_t z0, int nx_g, int ny_g,
                             const int* periodic, cudaStream_t s) {
  AsmPar P;
  P.nx_g = nx_g;
  P.ny_g = ny_g;
  P.nz_g = nz_global;
  P.z0 = z0;
  for (int i = 0; i < 3; ++i) P.per[i] = periodic[i];
  const int64_t nmax = std::max<int64_t>(std::max<int64_t>(nx_g, ny_g), nz_global);
  P.h = 1.0 / (double)nmax;
  P.rho_in = rho_in;
  P.rho_out = rho_out;
  P.dt = dt;
  P.nspheres = nspheres;
  // The density goes through the padded residual field r, free between solves; its ghost
  // layers are restored to zero afterwards because the solver relies on finite ghosts at
  // non-periodic boundaries.  NOT through p[1]: a neighbour rank's aphcg_upload_guess stores
  // its guess planes into this rank's p[1] ghost planes whenever it gets there -- it does not
  // wait for this rank -- and a zero landing between the two kernels below would be read as
  // a zero density (infinite coefficients).  Ghost planes of r are only written by peers
  // inside a run, which every rank enters through a barrier.
  k_density<<<kNumSMs * 8, 256, 0, s>>>(g, P, spheres, d.r);
  k_rows_from_density<<<kNumSMs * 8, 256, 0, s>>>(g, P, d.r, a[0], a[1], a[2], a[3], a[4], a[5],
                                              a[6], rhs);
  cudaMemsetAsync(d.r, 0, sizeof(double) * (size_t)g.ptotal, s);
}

}  // namespace acg
