// Hand-written FP64 CUDA kernels (sm_100a) of the CG hot path.
//
// Reference algorithm: linear::SolverConjugate<M>::Imp::Solve
// (src/linear/linear.ipp:42-125).  One iteration there is four stages and
// three sweeps over memory; here it is two kernels:
//
//   k_dir_spmv  ("iter3" of the previous iteration + "iter"):
//       x  += alpha_prev * p_old          (the x update of "iter2", deferred)
//       p   = r + beta * p_old            (on the tile and its 6 face neighbours)
//       Ap  = A p ,  partial sum p.Ap
//   k_update    ("iter2" + "check"):
//       r  -= alpha * Ap ,  partial sums r.r and max|r| ,  ghost copies of r
//
// The halo exchange of the reference (m.Comm(&p, direct_one), linear.ipp:100)
// disappears: ghost layers of p are recomputed locally from the ghost layers of
// r and p_old (bitwise the same numbers the owner computes), and ghost layers of
// r are written by k_update itself -- into this GPU's own ghost cells for
// periodic wrap, or straight into the neighbour GPU's ghost plane over NVLink.
#include <cooperative_groups.h>

#include <cstdlib>

#include "cg_kernels.cuh"
#include "cg_launch.h"

namespace cg = cooperative_groups;

namespace acg {

namespace {

constexpr int kBX = 32;  // threads along x (one warp = one contiguous row segment)
constexpr int kBY = 8;   // rows per block

struct Tile {
  int i, j, k0, k1;
  bool active;
};

// tile (bx, by, bz) of the tiled kernels' grid; tx, ty: this thread inside the kBX x kBY block
template <int VX>
__device__ __forceinline__ Tile tile_at(const Geom& g, int bx, int by, int bz, int tx, int ty,
                                        int zc) {
  Tile t;
  t.i = (bx * kBX + tx) * VX;
  t.j = by * kBY + ty;
  t.k0 = bz * zc;
  t.k1 = min(t.k0 + zc, g.nzl);
  t.active = (t.i < g.nx) && (t.j < g.ny);
  return t;
}
template <int VX>
__device__ __forceinline__ Tile my_tile(const Geom& g) {
  return tile_at<VX>(g, blockIdx.x, blockIdx.y, blockIdx.z, threadIdx.x, threadIdx.y, g.zc);
}

__device__ __forceinline__ unsigned num_blocks() { return gridDim.x * gridDim.y * gridDim.z; }
__device__ __forceinline__ unsigned block_id() {
  return blockIdx.x + gridDim.x * (blockIdx.y + gridDim.y * blockIdx.z);
}

// Stores the ghost copies of freshly computed inner values of a padded field:
// periodic images in x and y inside the slab, and the z images through
// lo_dst / hi_dst (own ghost plane or the neighbour GPU's, nullptr if none).
template <int VX>
__device__ __forceinline__ void store_images(const Geom& g, double* f, int64_t idp, int i, int j,
                                             int k, const Vec<VX>& val, double* lo_dst,
                                             double* hi_dst) {
  if (g.per_x) {
    if (i == 0) f[idp + g.nx] = val.v[0];
    if (i + VX == g.nx) f[idp - i - 1] = val.v[VX - 1];
  }
  if (g.per_y) {
    if (j == 0) stv<VX>(f + idp + (int64_t)g.ny * g.py, val);
    if (j == g.ny - 1) stv<VX>(f + idp - (int64_t)g.ny * g.py, val);
  }
  if (k == 0 && lo_dst) stv<VX>(lo_dst + i + (int64_t)j * g.py, val);
  if (k == g.nzl - 1 && hi_dst) stv<VX>(hi_dst + i + (int64_t)j * g.py, val);
}

}  // namespace

// ------------------------------------------------------------------------------
// k_dir_spmv (plain-load variant): neighbours of p come through L1/L2.
// ------------------------------------------------------------------------------
// One tile of the direction + SpMV stage with plain loads; returns this thread's part of
// sum p.Ap.  Shared by the stand-alone kernel and the persistent small-mesh kernel.
template <int VX>
__device__ __forceinline__ double dir_spmv_tile(const Geom& g, const DevPtrs& d, const Tile& t,
                                                const double beta, const double alpha_prev,
                                                const int par) {
  const double* __restrict__ po = d.p[par];
  double* __restrict__ pn = d.p[par ^ 1];
  const double* __restrict__ r = d.r;
  double acc = 0.0;
  if (t.active) {
    for (int k = t.k0; k < t.k1; ++k) {
      const int64_t idc = t.i + t.j * g.cy + k * g.cz;
      const int64_t idp = g.poff + t.i + t.j * g.py + k * g.pz;
      // p_new = r + beta*p_old (linear.ipp:97-99) at the cells and their neighbours
      auto pnew = [&](int64_t off) {
        const Vec<VX> rr = ldv<VX>(r + off), pp = ldv<VX>(po + off);
        Vec<VX> o;
#pragma unroll
        for (int v = 0; v < VX; ++v) o.v[v] = fma(beta, pp.v[v], rr.v[v]);
        return o;
      };
      auto pnew1 = [&](int64_t off) { return fma(beta, po[off], r[off]); };
      const Vec<VX> pold_c = ldv<VX>(po + idp);
      const Vec<VX> r_c = ldv<VX>(r + idp);
      Vec<VX> pc;
#pragma unroll
      for (int v = 0; v < VX; ++v) pc.v[v] = fma(beta, pold_c.v[v], r_c.v[v]);
      const double pxm = pnew1(idp - 1);
      const double pxp = pnew1(idp + VX);
      const Vec<VX> pym = pnew(idp - g.py), pyp = pnew(idp + g.py);
      const Vec<VX> pzm = pnew(idp - g.pz), pzp = pnew(idp + g.pz);

      Vec<VX> a[7];
#pragma unroll
      for (int q = 0; q < 7; ++q) a[q] = ldv_stream<VX>(d.a[q] + idc);
      Vec<VX> uu = ldv_stream<VX>(d.u + idc);
      Vec<VX> ap;
#pragma unroll
      for (int v = 0; v < VX; ++v) {
        const double xm = (v == 0) ? pxm : pc.v[v - 1];
        const double xp = (v == VX - 1) ? pxp : pc.v[v + 1];
        // accumulation order of the reference: centre, then q = 0..5 (linear.ipp:67-70)
        double s = pc.v[v] * a[0].v[v];
        s = fma(xm, a[1].v[v], s);
        s = fma(xp, a[2].v[v], s);
        s = fma(pym.v[v], a[3].v[v], s);
        s = fma(pyp.v[v], a[4].v[v], s);
        s = fma(pzm.v[v], a[5].v[v], s);
        s = fma(pzp.v[v], a[6].v[v], s);
        ap.v[v] = s;
        acc = fma(pc.v[v], s, acc);
        uu.v[v] = fma(alpha_prev, pold_c.v[v], uu.v[v]);  // linear.ipp:88, one iteration late
      }
      stv_stream<VX>(d.ap + idc, ap);
      stv_stream<VX>(d.u + idc, uu);
      stv<VX>(pn + idp, pc);
      // ghost cells of p_new owned by this thread
      if (t.i == 0) pn[idp - 1] = pxm;
      if (t.i + VX == g.nx) pn[idp + VX] = pxp;
      if (t.j == 0) stv<VX>(pn + idp - g.py, pym);
      if (t.j == g.ny - 1) stv<VX>(pn + idp + g.py, pyp);
      if (k == 0) stv<VX>(pn + idp - g.pz, pzm);
      if (k == g.nzl - 1) stv<VX>(pn + idp + g.pz, pzp);
    }
  }
  return acc;
}

template <int VX, bool kSingle>
__global__ void __launch_bounds__(kBX* kBY)
    k_dir_spmv_plain(const Geom g, const DevPtrs d) {
  __shared__ double sm[32];
  __shared__ int sm_flag;
  __shared__ DirView view;
  CgState* st = d.st;
  if (st->done) return;
  dir_view(d, &view);
  double acc = 0.0;
  if (!view.done)
    acc = dir_spmv_tile<VX>(g, d, my_tile<VX>(g), view.beta, view.alpha_prev, view.iter & 1);
  const double bsum = block_reduce<false>(acc, sm);
  const int tid = threadIdx.x + blockDim.x * threadIdx.y;
  if (tid == 0) d.partials[block_id()] = bsum;
  if (last_block(&st->counter_a, num_blocks(), &sm_flag)) {
    const double tot = reduce_slots<false>(d.partials, num_blocks(), sm);
    dir_epilogue<kSingle>(d, view, tot);
  }
}

// ------------------------------------------------------------------------------
// k_update: r -= alpha*Ap, sum r^2, max|r|, ghost copies of r ("iter2"+"check").
// Pure streaming (24 B/cell): a CTA owns kUR consecutive rows of one plane range,
// every thread keeps kUR independent 128-bit load pairs in flight.
// ------------------------------------------------------------------------------
constexpr int kUT = 256;  // threads

// Persistent: the grid is a fixed number of CTAs per SM and every CTA walks over
// work items (x-chunk, group of UR rows, plane) with a grid stride; each thread keeps
// UR independent 128-bit load pairs in flight.  Few CTAs -> few partial slots.
// kPre (opt-in Jacobi preconditioner, z = r/diag): the true residual lives in the
// compact array d.rc, the padded field d.r carries z (it is what the direction
// kernel combines with p_old), and the sums are r.z (-> alpha, beta) and r.r (-> norm).
// The update stage over the work items first, first + stride, ... (an item = UR rows of one
// x-chunk of one plane): r -= alpha*Ap with the partial sums of r.r (or r.z, kPre) and max|r|.
// Shared by the stand-alone persistent-grid kernel and the persistent small-mesh kernel.
template <int VX, int UR, bool kPre>
__device__ __forceinline__ void update_items(const Geom& g, const DevPtrs& d, const double alpha,
                                             const int64_t first, const int64_t stride,
                                             const int64_t nwork, double& acc, double& amax,
                                             double& acc2) {
  double* __restrict__ r = d.r;
  // g.utx threads span one row segment; on narrow meshes (nx/VX < kUT) the remaining
  // kUT/g.utx thread rows of the CTA take further row groups, so no thread idles
  const int utx = g.utx, uty = kUT / utx;
  const int tid = threadIdx.x + blockDim.x * threadIdx.y;
  const int tx = tid % utx, ty = tid / utx;
  const int xchunks = (g.nx + utx * VX - 1) / (utx * VX);
  const int jgroups = (g.ny + UR * uty - 1) / (UR * uty);
  for (int64_t w = first; w < nwork; w += stride) {
    const int xc = (int)(w % xchunks);
    const int64_t t = w / xchunks;
    const int j0 = ((int)(t % jgroups) * uty + ty) * UR;
    const int k = (int)(t / jgroups);
    const int i = (xc * utx + tx) * VX;
    if (i >= g.nx) continue;
    Vec<VX> ap[UR], rv[UR], dg[kPre ? UR : 1];
#pragma unroll
    for (int u = 0; u < UR; ++u) {
      const int j = j0 + u;
      if (j < g.ny) {
        const int64_t idc = i + j * g.cy + k * g.cz;
        ap[u] = ldv_stream<VX>(d.ap + idc);
        if constexpr (kPre) {
          rv[u] = ldv_stream<VX>(d.rc + idc);
          dg[u] = ldv_stream<VX>(d.a[0] + idc);
        } else {
          rv[u] = ldv_stream<VX>(r + g.poff + i + j * g.py + k * g.pz);
        }
      }
    }
#pragma unroll
    for (int u = 0; u < UR; ++u) {
      const int j = j0 + u;
      if (j < g.ny) {
        const int64_t idp = g.poff + i + j * g.py + k * g.pz;
        Vec<VX> zv;
#pragma unroll
        for (int v = 0; v < VX; ++v) {
          rv[u].v[v] = fma(-alpha, ap[u].v[v], rv[u].v[v]);  // linear.ipp:89
          amax = fmax(amax, fabs(rv[u].v[v]));               // :91
          if constexpr (kPre) {
            zv.v[v] = rv[u].v[v] / dg[u].v[v];
            acc = fma(rv[u].v[v], zv.v[v], acc);             // r.z
            acc2 = fma(rv[u].v[v], rv[u].v[v], acc2);        // r.r
          } else {
            acc = fma(rv[u].v[v], rv[u].v[v], acc);          // :90
          }
        }
        if (kPre) {
          stv_stream<VX>(d.rc + i + j * g.cy + k * g.cz, rv[u]);
          stv_stream<VX>(r + idp, zv);
          store_images<VX>(g, r, idp, i, j, k, zv, d.r_lo_dst, d.r_hi_dst);
        } else {
          stv_stream<VX>(r + idp, rv[u]);  // next read is a whole kernel away
          store_images<VX>(g, r, idp, i, j, k, rv[u], d.r_lo_dst, d.r_hi_dst);
        }
      }
    }
  }
}
__host__ __device__ inline int64_t update_nwork(const Geom& g, int vx, int ur) {
  const int uty = kUT / g.utx;
  const int xchunks = (g.nx + g.utx * vx - 1) / (g.utx * vx);
  return (int64_t)xchunks * ((g.ny + ur * uty - 1) / (ur * uty)) * g.nzl;
}

template <int VX, bool kSingle, int UR, bool kPre>
__global__ void __launch_bounds__(kUT) k_update(const Geom g, const DevPtrs d) {
  __shared__ double sm[32];
  __shared__ int sm_flag;
  __shared__ UpdView view;
  CgState* st = d.st;
  if (st->done) return;
  upd_view(d, &view);
  double acc = 0.0, amax = 0.0, acc2 = 0.0;
  update_items<VX, UR, kPre>(g, d, view.alpha, blockIdx.x, gridDim.x,
                             view.error ? 0 : update_nwork(g, VX, UR), acc, amax, acc2);
  if (d.r_lo_dst != nullptr || d.r_hi_dst != nullptr) __threadfence_system();
  const double bsum = block_reduce<false>(acc, sm);
  const double bmax = block_reduce<true>(amax, sm);
  const double bsum2 = kPre ? block_reduce<false>(acc2, sm) : 0.0;
  const int tid = threadIdx.x;
  const unsigned nblk = gridDim.x, bid = blockIdx.x;
  if (tid == 0) {
    d.partials[bid] = bsum;
    d.partials2[bid] = bmax;
    if (kPre) d.partials3[bid] = bsum2;
  }
  if (last_block(&st->counter_b, nblk, &sm_flag)) {
    const double tot = reduce_slots<false>(d.partials, nblk, sm);
    const double mx = reduce_slots<true>(d.partials2, nblk, sm);
    const double tot2 = kPre ? reduce_slots<false>(d.partials3, nblk, sm) : tot;
    if (tid == 0) {
      st->loc_sum = tot;
      st->loc_max = mx;
      st->loc_sum2 = tot2;
      if (view.pend) cg_finish_dir(st, view.pAp);  // every CTA has read the old state by now
      if (view.error) {
        st->error = 1;
        st->done = 1;
      }
      if (kSingle) cg_finish_upd(st, d.history, tot, mx, tot2);
      if (!kSingle && d.cm.wait_in_kernel && !view.error) st->pend_upd = 1;
    }
    if (!kSingle && d.cm.use_mail && !view.error)
      mail_push(d.cm, st->seq_base, view.iter, 1, tot, mx, tot2);
  }
}

// ------------------------------------------------------------------------------
// k_cg_persistent: the whole loop in ONE cooperative kernel, for meshes whose fields stay in
// the 126 MB L2 (config 1: 64^3).  There two launches per iteration plus their ramp-up cost
// more than the work (64^3: 29 us per iteration for ~4 us of memory traffic).  Same tiles,
// same arithmetic as k_dir_spmv_plain + k_update; the two scalar reductions per iteration
// become grid-wide barriers after which EVERY CTA adds the per-CTA slots in the same fixed
// order (bitwise the same sum everywhere, no broadcast needed), and every CTA carries the loop
// scalars in registers.  Runs until the exit rule fires or `max_iters` iterations are done.
// ------------------------------------------------------------------------------
template <int VX, int UR>
__global__ void __launch_bounds__(kBX* kBY, 2)
    k_cg_persistent(const Geom g, const DevPtrs d, const dim3 tgrid, const int zc,
                    const int max_iters) {
  cg::grid_group grid = cg::this_grid();
  __shared__ double sm[32];
  CgState* st = d.st;
  const int tid = threadIdx.x + blockDim.x * threadIdx.y;
  const unsigned G = gridDim.x;
  const unsigned ntiles = tgrid.x * tgrid.y * tgrid.z;
  const int64_t nwork = update_nwork(g, VX, UR);
  // loop scalars, identical in every thread of the grid
  double rr = st->rr, rr_prev = st->rr_prev, alpha_prev = st->alpha_prev, pAp = st->pAp;
  double alpha_prev2 = st->alpha_prev2, rnorm2 = st->rnorm2, max_r = st->max_r;
  double residual = st->residual;
  int iter = st->iter, done = st->done;
  const double tol = st->tol, vol = st->cell_volume;
  const int miniter = st->miniter, maxiter = st->maxiter, maxnorm = st->maxnorm;
  const int hist_cap = st->hist_cap;
  for (int it = 0; it < max_iters && !done; ++it) {
    // ---- stages "iter3" (of the previous iteration) + "iter" -----------------------------
    const double beta = iter == 0 ? 0.0 : rr / (rr_prev + 1e-100);  // linear.ipp:98
    double acc = 0.0;
    for (unsigned vb = blockIdx.x; vb < ntiles; vb += G) {
      const int bx = vb % tgrid.x, by = (vb / tgrid.x) % tgrid.y, bz = vb / (tgrid.x * tgrid.y);
      acc += dir_spmv_tile<VX>(g, d, tile_at<VX>(g, bx, by, bz, threadIdx.x, threadIdx.y, zc), beta,
                               alpha_prev, iter & 1);
    }
    const double bsum = block_reduce<false>(acc, sm);
    if (tid == 0) d.partials[blockIdx.x] = bsum;
    grid.sync();
    pAp = reduce_slots<false>(d.partials, G, sm);
    const double alpha = rr / (pAp + 1e-100);  // linear.ipp:84
    // ---- stages "iter2" + "check" ------------------------------------------------------------
    double a1 = 0.0, amax = 0.0, a2 = 0.0;
    update_items<VX, UR, false>(g, d, alpha, blockIdx.x, G, nwork, a1, amax, a2);
    const double bs = block_reduce<false>(a1, sm);
    const double bm = block_reduce<true>(amax, sm);
    if (tid == 0) {
      d.partials2[blockIdx.x] = bm;
      d.partials3[blockIdx.x] = bs;
    }
    grid.sync();
    const double tot = reduce_slots<false>(d.partials3, G, sm);
    const double mx = reduce_slots<true>(d.partials2, G, sm);
    alpha_prev2 = alpha_prev;
    alpha_prev = alpha;
    rr_prev = rr;
    rr = tot;
    rnorm2 = tot;
    max_r = mx;
    residual = maxnorm ? mx / vol : sqrt(tot / vol);  // linear.ipp:103-107
    ++iter;
    if (blockIdx.x == 0 && tid == 0 && iter - 1 < hist_cap) d.history[iter - 1] = residual;
    done = (iter >= miniter && (iter > maxiter || residual < tol)) ? 1 : 0;  // :110-113
  }
  if (blockIdx.x == 0 && tid == 0) {
    st->rr = rr;
    st->rr_prev = rr_prev;
    st->pAp = pAp;
    st->alpha_prev = alpha_prev;
    st->alpha_prev2 = alpha_prev2;
    st->rnorm2 = rnorm2;
    st->max_r = max_r;
    st->residual = residual;
    st->iter = iter;
    st->done = done;
  }
}

static int update_ur() {
  static int ur = [] {
    const char* e = getenv("APHCG_UPD_UR");
    const int v = e ? atoi(e) : 8;
    return (v == 1 || v == 2 || v == 4 || v == 8) ? v : 8;
  }();
  return ur;
}
static int update_ctas_per_sm() {
  static int c = [] {
    const char* e = getenv("APHCG_UPD_CTAS");
    const int v = e ? atoi(e) : 24;  // 8: 0.540, 12: 0.532, 16: 0.524, 24: 0.517 ms at 512^3
    return v >= 1 && v <= 32 ? v : 24;
  }();
  return c;
}

static dim3 update_grid(const Geom& g, int vx) {
  const int64_t nwork = update_nwork(g, vx, update_ur());
  const int64_t cap = kNumSMs * (int64_t)update_ctas_per_sm();
  return dim3((unsigned)(nwork < cap ? nwork : cap));
}

// symmetry check of the off-diagonals inside the slab (see cg_launch.h)
__global__ void k_check_symmetry(const Geom g, const DevPtrs d, int* flag) {
  bool bad = false;
  for (int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; c < g.ncell;
       c += (int64_t)gridDim.x * blockDim.x) {
    const int i = (int)(c % g.nx);
    const int j = (int)((c / g.nx) % g.ny);
    const int k = (int)(c / g.cz);
    if (i + 1 < g.nx) bad |= !(d.a[2][c] == d.a[1][c + 1]);
    if (j + 1 < g.ny) bad |= !(d.a[4][c] == d.a[3][c + g.cy]);
    if (k + 1 < g.nzl) bad |= !(d.a[6][c] == d.a[5][c + g.cz]);
  }
  if (__any_sync(0xffffffffu, bad) && (threadIdx.x & 31) == 0) atomicOr(flag, 1);
}

// Multi-GPU: runs after the all-reduce of loc_sum / loc_max.
// One warp.  With mailboxes it first waits for every rank's contribution of this
// step and sums them in rank order; otherwise NCCL has already reduced loc_sum/loc_max.
__global__ void k_finish_dir(const DevPtrs d) {
  CgState* st = d.st;
  if (st->done) return;
  double sum = st->loc_sum, mx = 0.0, sum2 = 0.0;
  if (d.cm.use_mail && !mail_wait(d.cm, st->seq_base, st->iter, 0, nullptr, &sum, &mx, &sum2)) {
    if (threadIdx.x == 0) st->error = st->done = 1;
    return;
  }
  if (threadIdx.x == 0) cg_finish_dir(st, sum);
}
__global__ void k_finish_upd(const DevPtrs d) {
  CgState* st = d.st;
  if (st->done) return;
  // Comm::wait_in_kernel: runs once per chunk of iterations, so that the host finds a fully
  // committed state when it looks; nothing to do unless an update stage is pending
  if (d.cm.wait_in_kernel && !st->pend_upd) return;
  double sum = st->loc_sum, mx = st->loc_max, sum2 = st->precond ? st->loc_sum2 : st->loc_sum;
  if (d.cm.use_mail && !mail_wait(d.cm, st->seq_base, st->iter, 1, nullptr, &sum, &mx, &sum2)) {
    if (threadIdx.x == 0) st->error = st->done = 1;
    return;
  }
  if (threadIdx.x == 0) cg_finish_upd(st, d.history, sum, mx, st->precond ? sum2 : sum);
}
__global__ void k_finish_init(CgState* st) {
  st->rr = st->loc_sum;
  st->rnorm2_0 = st->loc_sum2;
}

// ------------------------------------------------------------------------------
// k_residual: out = sign*(A f [+ rhs]) from a padded field f (stage "init",
// linear.ipp:48-56, with sign=-1 and rhs; the bare operator with sign=+1).
// kToPadded: result goes to the padded residual (with ghost copies) and sum r^2
// is formed; otherwise to a compact array.
// ------------------------------------------------------------------------------
template <int VX, bool kInit, bool kSingle, bool kPre = false>
__global__ void __launch_bounds__(kBX* kBY)
    k_residual(const Geom g, const DevPtrs d, const double* __restrict__ f, double* out) {
  __shared__ double sm[32];
  __shared__ int sm_flag;
  CgState* st = d.st;
  const Tile t = my_tile<VX>(g);
  double acc = 0.0, acc2 = 0.0;  // acc2: sum r^2 when acc is sum r.z (preconditioned start)
  if (t.active) {
    for (int k = t.k0; k < t.k1; ++k) {
      const int64_t idc = t.i + t.j * g.cy + k * g.cz;
      const int64_t idp = g.poff + t.i + t.j * g.py + k * g.pz;
      const Vec<VX> fc = ldv<VX>(f + idp);
      const double fxm = f[idp - 1], fxp = f[idp + VX];
      const Vec<VX> fym = ldv<VX>(f + idp - g.py), fyp = ldv<VX>(f + idp + g.py);
      const Vec<VX> fzm = ldv<VX>(f + idp - g.pz), fzp = ldv<VX>(f + idp + g.pz);
      Vec<VX> a[7];
#pragma unroll
      for (int q = 0; q < 7; ++q) a[q] = ldv_stream<VX>(d.a[q] + idc);
      Vec<VX> b;
      if (kInit) b = ldv_stream<VX>(d.rhs + idc);
      Vec<VX> res;
#pragma unroll
      for (int v = 0; v < VX; ++v) {
        const double xm = (v == 0) ? fxm : fc.v[v - 1];
        const double xp = (v == VX - 1) ? fxp : fc.v[v + 1];
        // linear.ipp:50-53: u*e0 + e7, then neighbours q = 0..5
        double s = kInit ? fma(fc.v[v], a[0].v[v], b.v[v]) : fc.v[v] * a[0].v[v];
        s = fma(xm, a[1].v[v], s);
        s = fma(xp, a[2].v[v], s);
        s = fma(fym.v[v], a[3].v[v], s);
        s = fma(fyp.v[v], a[4].v[v], s);
        s = fma(fzm.v[v], a[5].v[v], s);
        s = fma(fzp.v[v], a[6].v[v], s);
        res.v[v] = kInit ? -s : s;
        if (!kPre) acc = fma(res.v[v], res.v[v], acc);
      }
      if (kInit && kPre) {
        // preconditioned start: r compact, z = r/diag padded, sum r.z
        Vec<VX> zv;
#pragma unroll
        for (int v = 0; v < VX; ++v) {
          zv.v[v] = res.v[v] / a[0].v[v];
          acc = fma(res.v[v], zv.v[v], acc);
          acc2 = fma(res.v[v], res.v[v], acc2);
        }
        stv<VX>(d.rc + idc, res);
        stv<VX>(out + idp, zv);
        store_images<VX>(g, out, idp, t.i, t.j, k, zv, d.r_lo_dst, d.r_hi_dst);
      } else if (kInit) {
        stv<VX>(out + idp, res);
        store_images<VX>(g, out, idp, t.i, t.j, k, res, d.r_lo_dst, d.r_hi_dst);
      } else {
        stv<VX>(out + idc, res);
      }
    }
  }
  if (!kInit) return;
  if (d.r_lo_dst != nullptr || d.r_hi_dst != nullptr) __threadfence_system();
  const double bsum = block_reduce<false>(acc, sm);
  const double bsum2 = kPre ? block_reduce<false>(acc2, sm) : bsum;
  const int tid = threadIdx.x + blockDim.x * threadIdx.y;
  if (tid == 0) {
    d.partials[block_id()] = bsum;
    if (kPre) d.partials3[block_id()] = bsum2;
  }
  if (last_block(&st->counter_a, num_blocks(), &sm_flag)) {
    const double tot = reduce_slots<false>(d.partials, num_blocks(), sm);
    const double tot2 = kPre ? reduce_slots<false>(d.partials3, num_blocks(), sm) : tot;
    if (tid == 0) {
      st->loc_sum = tot;
      st->loc_sum2 = tot2;
      if (kSingle) {
        st->rr = tot;
        st->rnorm2_0 = tot2;
      }
    }
  }
}

// ------------------------------------------------------------------------------
// k_scatter_field: compact/laid-out source (or zero) -> compact u and padded copy
// with ghost images (what the caller's Comm'd initial guess provides,
// src/solver/proj.ipp:397).
// ------------------------------------------------------------------------------
template <int VX>
__global__ void __launch_bounds__(kBX* kBY)
    k_scatter_field(const Geom g, const double* __restrict__ src, int64_t s_off, int64_t s_sy,
                    int64_t s_sz, double* u, double* fpad, double* lo_dst, double* hi_dst) {
  const Tile t = my_tile<VX>(g);
  if (!t.active) return;
  for (int k = t.k0; k < t.k1; ++k) {
    const int64_t idc = t.i + t.j * g.cy + k * g.cz;
    const int64_t idp = g.poff + t.i + t.j * g.py + k * g.pz;
    Vec<VX> val;
#pragma unroll
    for (int v = 0; v < VX; ++v)
      val.v[v] = src ? src[s_off + t.i + v + t.j * s_sy + k * s_sz] : 0.0;
    if (u) stv<VX>(u + idc, val);
    stv<VX>(fpad + idp, val);
    store_images<VX>(g, fpad, idp, t.i, t.j, k, val, lo_dst, hi_dst);
  }
  if (lo_dst != nullptr || hi_dst != nullptr) __threadfence_system();
}

// compact u -> laid-out destination
__global__ void k_gather_field(const Geom g, const double* __restrict__ u, double* dst,
                               int64_t d_off, int64_t d_sy, int64_t d_sz) {
  const int64_t n = g.ncell;
  for (int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; c < n;
       c += (int64_t)gridDim.x * blockDim.x) {
    const int i = (int)(c % g.nx);
    const int64_t t = c / g.nx;
    const int j = (int)(t % g.ny);
    const int64_t k = t / g.ny;
    dst[d_off + i + j * d_sy + k * d_sz] = u[c];
  }
}

// x += alpha_prev * p : the deferred x update of the last iteration ("iter2", linear.ipp:88).
// With batched updates (CgState::xbatch) an even number of completed iterations leaves the
// last TWO updates pending; they are applied in iteration order.
template <int VX>
__global__ void __launch_bounds__(kBX* kBY) k_final_update(const Geom g, const DevPtrs d) {
  const CgState* st = d.st;
  const double alpha_prev = st->alpha_prev, alpha_prev2 = st->alpha_prev2;
  const double* __restrict__ p = d.p[st->iter & 1];
  const double* __restrict__ p2 = d.p[(st->iter & 1) ^ 1];
  const bool two = st->xbatch && st->iter >= 2 && (st->iter & 1) == 0;
  const Tile t = my_tile<VX>(g);
  if (!t.active) return;
  for (int k = t.k0; k < t.k1; ++k) {
    const int64_t idc = t.i + t.j * g.cy + k * g.cz;
    const int64_t idp = g.poff + t.i + t.j * g.py + k * g.pz;
    const Vec<VX> pv = ldv<VX>(p + idp);
    Vec<VX> uu = ldv<VX>(d.u + idc);
    if (two) {
      const Vec<VX> pv2 = ldv<VX>(p2 + idp);
#pragma unroll
      for (int v = 0; v < VX; ++v) uu.v[v] = fma(alpha_prev2, pv2.v[v], uu.v[v]);
    }
#pragma unroll
    for (int v = 0; v < VX; ++v) uu.v[v] = fma(alpha_prev, pv.v[v], uu.v[v]);
    stv<VX>(d.u + idc, uu);
  }
}

// ------------------------------------------------------------------------------
// AoS rows -> SoA coefficient arrays (one-time transpose at upload).
// rows: element (cell) index = r_off + i + j*r_sy + k*r_sz, 8 doubles per cell,
// for planes [k0, k0+nk) of the slab (rows is the base of the chunk: plane k0
// of the slab is plane 0 of the chunk).
// ------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
    k_rows_to_soa(const Geom g, const double* __restrict__ rows, int64_t r_off, int64_t r_sy,
                  int64_t r_sz, int k0, int nk, double* a0, double* a1, double* a2, double* a3,
                  double* a4, double* a5, double* a6, double* rhs) {
  // one block = 256 consecutive cells of one x-row segment; transpose through smem
  __shared__ double tile[8][256 + 1];
  double* outs[8] = {a0, a1, a2, a3, a4, a5, a6, rhs};
  const int segs = (g.nx + 255) / 256;
  const int64_t nrow = (int64_t)g.ny * nk;
  for (int64_t w = blockIdx.x; w < nrow * segs; w += gridDim.x) {
    const int seg = (int)(w % segs);
    const int64_t row = w / segs;
    const int j = (int)(row % g.ny);
    const int kk = (int)(row / g.ny);
    const int i0 = seg * 256;
    const int ncell = min(256, g.nx - i0);
    const double* src = rows + 8 * (r_off + i0 + j * r_sy + (int64_t)kk * r_sz);
    __syncthreads();
    for (int e = threadIdx.x; e < ncell * 8; e += 256) tile[e & 7][e >> 3] = __ldcs(src + e);
    __syncthreads();
    const int64_t dst = i0 + j * g.cy + (int64_t)(k0 + kk) * g.cz;
    for (int q = 0; q < 8; ++q)
      for (int c = threadIdx.x; c < ncell; c += 256) outs[q][dst + c] = tile[q][c];
  }
}

// SoA -> AoS rows (checking the device assembler)
__global__ void k_soa_to_rows(const Geom g, const DevPtrs d, double* rows) {
  const int64_t n = g.ncell;
  for (int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; c < n;
       c += (int64_t)gridDim.x * blockDim.x) {
#pragma unroll
    for (int q = 0; q < 7; ++q) rows[8 * c + q] = d.a[q][c];
    rows[8 * c + 7] = d.rhs[c];
  }
}

// ------------------------------------------------------------------------------
// Point Jacobi twin (linear::SolverJacobi, src/linear/linear.ipp:178-204):
// u_new = -(e7 + sum_q e[1+q] u[nb_q]) / e0 ; maxdiff = max|u_new - u|.
// Iterate lives in the padded ping-pong buffers p[0]/p[1].
// ------------------------------------------------------------------------------
// stage "check" of SolverJacobi (linear.ipp:195-204)
__device__ __forceinline__ void jacobi_finish(CgState* st, double* history, double maxdiff) {
  st->residual = maxdiff;
  const int it = st->iter + 1;
  if (it - 1 < st->hist_cap) history[it - 1] = maxdiff;
  st->iter = it;
  if (it >= st->miniter && (it > st->maxiter || maxdiff < st->tol)) st->done = 1;
}
// Multi-GPU: one warp.  With mailboxes it waits for every rank's max|u_new - u| of this
// iteration (phase 0); otherwise NCCL has already reduced loc_max.
__global__ void k_finish_jacobi(const DevPtrs d) {
  CgState* st = d.st;
  if (st->done) return;
  double sum = 0.0, mx = st->loc_max, sum2 = 0.0;
  if (d.cm.use_mail && !mail_wait(d.cm, st->seq_base, st->iter, 0, nullptr, &sum, &mx, &sum2)) {
    if (threadIdx.x == 0) st->error = st->done = 1;
    return;
  }
  if (threadIdx.x == 0) jacobi_finish(st, d.history, mx);
}

template <int VX, bool kSingle>
__global__ void __launch_bounds__(kBX* kBY) k_jacobi(const Geom g, const DevPtrs d) {
  __shared__ double sm[32];
  __shared__ int sm_flag;
  CgState* st = d.st;
  if (st->done) return;
  const int par = st->iter & 1;
  const double* __restrict__ f = d.p[par];
  double* __restrict__ fn = d.p[par ^ 1];
  const Tile t = my_tile<VX>(g);
  double amax = 0.0;
  if (t.active) {
    for (int k = t.k0; k < t.k1; ++k) {
      const int64_t idc = t.i + t.j * g.cy + k * g.cz;
      const int64_t idp = g.poff + t.i + t.j * g.py + k * g.pz;
      const Vec<VX> fc = ldv<VX>(f + idp);
      const double fxm = f[idp - 1], fxp = f[idp + VX];
      const Vec<VX> fym = ldv<VX>(f + idp - g.py), fyp = ldv<VX>(f + idp + g.py);
      const Vec<VX> fzm = ldv<VX>(f + idp - g.pz), fzp = ldv<VX>(f + idp + g.pz);
      Vec<VX> a[7];
#pragma unroll
      for (int q = 0; q < 7; ++q) a[q] = ldv_stream<VX>(d.a[q] + idc);
      const Vec<VX> b = ldv_stream<VX>(d.rhs + idc);
      Vec<VX> res;
#pragma unroll
      for (int v = 0; v < VX; ++v) {
        const double xm = (v == 0) ? fxm : fc.v[v - 1];
        const double xp = (v == VX - 1) ? fxp : fc.v[v + 1];
        double s = b.v[v];
        s = fma(xm, a[1].v[v], s);
        s = fma(xp, a[2].v[v], s);
        s = fma(fym.v[v], a[3].v[v], s);
        s = fma(fyp.v[v], a[4].v[v], s);
        s = fma(fzm.v[v], a[5].v[v], s);
        s = fma(fzp.v[v], a[6].v[v], s);
        res.v[v] = -s / a[0].v[v];
        amax = fmax(amax, fabs(res.v[v] - fc.v[v]));
      }
      stv<VX>(fn + idp, res);
      store_images<VX>(g, fn, idp, t.i, t.j, k, res, d.p_lo_dst[par ^ 1], d.p_hi_dst[par ^ 1]);
    }
  }
  if (d.p_lo_dst[0] != nullptr || d.p_hi_dst[0] != nullptr) __threadfence_system();
  const double bmax = block_reduce<true>(amax, sm);
  const int tid = threadIdx.x + blockDim.x * threadIdx.y;
  if (tid == 0) d.partials2[block_id()] = bmax;
  if (last_block(&st->counter_b, num_blocks(), &sm_flag)) {
    const double mx = reduce_slots<true>(d.partials2, num_blocks(), sm);
    if (tid == 0) {
      st->loc_max = mx;
      if (kSingle) jacobi_finish(st, d.history, mx);
    }
    if (!kSingle && d.cm.use_mail) mail_push(d.cm, st->seq_base, st->iter, 0, 0.0, mx);
  }
}

// ==============================================================================
// launch wrappers
// ==============================================================================
static dim3 tile_grid(const Geom& g, int vx) {
  return dim3((g.nx + kBX * vx - 1) / (kBX * vx), (g.ny + kBY - 1) / kBY, (g.nzl + g.zc - 1) / g.zc);
}

unsigned tile_blocks_for(const Geom& g, int vx) {
  const dim3 gr = tile_grid(g, vx);
  return gr.x * gr.y * gr.z;
}

// Upper bound on the CTAs of any kernel that writes per-CTA reduction slots: the tiled
// kernels (init residual, plain SpMV, Jacobi) and the persistent update kernel, whose
// grid is capped at 148 SMs x at most 32 CTAs whatever APHCG_UPD_CTAS says.
unsigned tile_blocks(const Geom& g, int vx) {
  const dim3 gr = tile_grid(g, vx);
  const unsigned tiled = gr.x * gr.y * gr.z, persistent = (unsigned)kNumSMs * 32u;
  return tiled > persistent ? tiled : persistent;
}

#define APHCG_DISPATCH_VX(vx, ...) \
  do {                             \
    if ((vx) == 2) {               \
      constexpr int VX = 2;        \
      __VA_ARGS__;                 \
    } else {                       \
      constexpr int VX = 1;        \
      __VA_ARGS__;                 \
    }                              \
  } while (0)

void launch_dir_spmv_plain(const Geom& g, const DevPtrs& d, int vx, bool single, cudaStream_t s) {
  const dim3 gr = tile_grid(g, vx), bl(kBX, kBY);
  APHCG_DISPATCH_VX(vx, {
    if (single)
      k_dir_spmv_plain<VX, true><<<gr, bl, 0, s>>>(g, d);
    else
      k_dir_spmv_plain<VX, false><<<gr, bl, 0, s>>>(g, d);
  });
}

void launch_check_symmetry(const Geom& g, const DevPtrs& d, int* flag, cudaStream_t s) {
  k_check_symmetry<<<kNumSMs * 8, 256, 0, s>>>(g, d, flag);
}

template <int VX, int UR, bool kPre>
static void launch_update_t(const Geom& g, const DevPtrs& d, bool single, dim3 gr, cudaStream_t s) {
  if (single)
    k_update<VX, true, UR, kPre><<<gr, kUT, 0, s>>>(g, d);
  else
    k_update<VX, false, UR, kPre><<<gr, kUT, 0, s>>>(g, d);
}

void launch_update(const Geom& g, const DevPtrs& d, int vx, bool single, bool precond,
                   cudaStream_t s) {
  const dim3 gr = update_grid(g, vx);
  const int ur = update_ur();
  APHCG_DISPATCH_VX(vx, {
    if (precond) {
      launch_update_t<VX, 4, true>(g, d, single, gr, s);
    } else {
      switch (ur) {
        case 1: launch_update_t<VX, 1, false>(g, d, single, gr, s); break;
        case 2: launch_update_t<VX, 2, false>(g, d, single, gr, s); break;
        case 4: launch_update_t<VX, 4, false>(g, d, single, gr, s); break;
        default: launch_update_t<VX, 8, false>(g, d, single, gr, s); break;
      }
    }
  });
}

// ---- persistent small-mesh loop ---------------------------------------------------------
template <int VX, int UR>
static int persistent_max_ctas() {
  int per_sm = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_cg_persistent<VX, UR>, kBX * kBY, 0) !=
      cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return per_sm * kNumSMs;
}

bool persistent_plan(const Geom& g, int vx, PersistPlan* out) {
  const int cap = vx == 2 ? persistent_max_ctas<2, 1>() : persistent_max_ctas<1, 1>();
  if (cap < kNumSMs) return false;
  // planes per tile: fewest sequential plane steps per CTA, one extra step per round of tiles
  int64_t best = -1;
  PersistPlan p{};
  for (int zc = 1; zc <= 16; zc *= 2) {
    if (zc > g.nzl && zc > 1) break;
    const dim3 tg((g.nx + kBX * vx - 1) / (kBX * vx), (g.ny + kBY - 1) / kBY, (g.nzl + zc - 1) / zc);
    const int64_t tiles = (int64_t)tg.x * tg.y * tg.z;
    const int64_t G = tiles < cap ? tiles : cap;
    const int64_t cost = ((tiles + G - 1) / G) * (zc + 1);
    if (best < 0 || cost < best) {
      best = cost;
      p.tgrid = tg;
      p.zc = zc;
      p.grid = (unsigned)G;
    }
  }
  // rows per thread of the update stage: as many loads in flight as still give every CTA work
  p.ur = 1;
  for (int ur = 4; ur > 1; ur /= 2) {
    if (update_nwork(g, vx, ur) >= (int64_t)p.grid) {
      p.ur = ur;
      break;
    }
  }
  *out = p;
  return true;
}

template <int VX, int UR>
static cudaError_t launch_persistent_t(const Geom& g, const DevPtrs& d, const PersistPlan& p,
                                       int max_iters, cudaStream_t s) {
  Geom gg = g;
  DevPtrs dd = d;
  dim3 tgrid = p.tgrid;
  int zc = p.zc, mi = max_iters;
  void* args[] = {&gg, &dd, &tgrid, &zc, &mi};
  return cudaLaunchCooperativeKernel((const void*)k_cg_persistent<VX, UR>, dim3(p.grid),
                                     dim3(kBX, kBY), args, 0, s);
}

cudaError_t launch_cg_persistent(const Geom& g, const DevPtrs& d, int vx, const PersistPlan& p,
                                 int max_iters, cudaStream_t s) {
  if (vx == 2) {
    switch (p.ur) {
      case 4: return launch_persistent_t<2, 4>(g, d, p, max_iters, s);
      case 2: return launch_persistent_t<2, 2>(g, d, p, max_iters, s);
      default: return launch_persistent_t<2, 1>(g, d, p, max_iters, s);
    }
  }
  switch (p.ur) {
    case 4: return launch_persistent_t<1, 4>(g, d, p, max_iters, s);
    case 2: return launch_persistent_t<1, 2>(g, d, p, max_iters, s);
    default: return launch_persistent_t<1, 1>(g, d, p, max_iters, s);
  }
}

void launch_finish_dir(const DevPtrs& d, cudaStream_t s) { k_finish_dir<<<1, 32, 0, s>>>(d); }
void launch_finish_upd(const DevPtrs& d, cudaStream_t s) { k_finish_upd<<<1, 32, 0, s>>>(d); }
void launch_finish_init(const DevPtrs& d, cudaStream_t s) { k_finish_init<<<1, 1, 0, s>>>(d.st); }
void launch_finish_jacobi(const DevPtrs& d, cudaStream_t s) { k_finish_jacobi<<<1, 32, 0, s>>>(d); }

void launch_init_residual(const Geom& g, const DevPtrs& d, int vx, bool single, bool precond,
                          cudaStream_t s) {
  const dim3 gr = tile_grid(g, vx), bl(kBX, kBY);
  APHCG_DISPATCH_VX(vx, {
    if (precond) {
      if (single)
        k_residual<VX, true, true, true><<<gr, bl, 0, s>>>(g, d, d.p[1], d.r);
      else
        k_residual<VX, true, false, true><<<gr, bl, 0, s>>>(g, d, d.p[1], d.r);
    } else {
      if (single)
        k_residual<VX, true, true><<<gr, bl, 0, s>>>(g, d, d.p[1], d.r);
      else
        k_residual<VX, true, false><<<gr, bl, 0, s>>>(g, d, d.p[1], d.r);
    }
  });
}

void launch_apply(const Geom& g, const DevPtrs& d, int vx, cudaStream_t s) {
  const dim3 gr = tile_grid(g, vx), bl(kBX, kBY);
  APHCG_DISPATCH_VX(vx, { (k_residual<VX, false, true><<<gr, bl, 0, s>>>(g, d, d.p[1], d.ap)); });
}

void launch_scatter_field(const Geom& g, const double* src, int64_t off, int64_t sy, int64_t sz,
                          double* u, double* fpad, double* lo_dst, double* hi_dst, int vx,
                          cudaStream_t s) {
  const dim3 gr = tile_grid(g, vx), bl(kBX, kBY);
  APHCG_DISPATCH_VX(
      vx, { (k_scatter_field<VX><<<gr, bl, 0, s>>>(g, src, off, sy, sz, u, fpad, lo_dst, hi_dst)); });
}

void launch_gather_field(const Geom& g, const double* u, double* dst, int64_t off, int64_t sy,
                         int64_t sz, cudaStream_t s) {
  k_gather_field<<<kNumSMs * 8, 256, 0, s>>>(g, u, dst, off, sy, sz);
}

void launch_final_update(const Geom& g, const DevPtrs& d, int vx, cudaStream_t s) {
  const dim3 gr = tile_grid(g, vx), bl(kBX, kBY);
  APHCG_DISPATCH_VX(vx, { (k_final_update<VX><<<gr, bl, 0, s>>>(g, d)); });
}

void launch_rows_to_soa(const Geom& g, const double* rows, int64_t off, int64_t sy, int64_t sz,
                        int k0, int nk, double* const* a, double* rhs, cudaStream_t s) {
  k_rows_to_soa<<<kNumSMs * 16, 256, 0, s>>>(g, rows, off, sy, sz, k0, nk, a[0], a[1], a[2], a[3],
                                         a[4], a[5], a[6], rhs);
}

void launch_soa_to_rows(const Geom& g, const DevPtrs& d, double* rows, cudaStream_t s) {
  k_soa_to_rows<<<kNumSMs * 8, 256, 0, s>>>(g, d, rows);
}

void launch_jacobi(const Geom& g, const DevPtrs& d, int vx, bool single, cudaStream_t s) {
  const dim3 gr = tile_grid(g, vx), bl(kBX, kBY);
  APHCG_DISPATCH_VX(vx, {
    if (single)
      k_jacobi<VX, true><<<gr, bl, 0, s>>>(g, d);
    else
      k_jacobi<VX, false><<<gr, bl, 0, s>>>(g, d);
  });
}

}  // namespace acg
