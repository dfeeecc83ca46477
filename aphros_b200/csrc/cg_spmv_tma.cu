// k_dir_spmv_tma: the direction + SpMV kernel of the CG loop with the p/r
// planes and their x/y halo staged in shared memory by TMA.
//
// Same arithmetic as k_dir_spmv_plain (cg_kernels.cu), i.e. reference stages
// "iter3" + "iter" (src/linear/linear.ipp:64-101):
//   x += alpha_prev*p_old ;  p = r + beta*p_old ;  Ap = A p ;  sum p.Ap
//
// One CTA owns a TX x TY column of cells and marches through ZC planes of it.
// For every plane, one elected thread issues two cp.async.bulk.tensor (TMA)
// loads -- the (TX+4) x (TY+2) boxes of r and p_old, halo included -- into a
// ring of S stages that complete on an mbarrier; all threads then form p_new
// for the whole box in shared memory (ring of 4 planes: z-1, z, z+1 and the one
// being written), store the cells this CTA owns, and evaluate the stencil for
// the previous plane with every neighbour of p read from shared memory.  The 7
// coefficient streams and x have no reuse: they go straight from HBM to
// registers with 128-bit loads issued a whole step ahead of their use.
#include <cuda.h>

#include <algorithm>
#include <cstdio>
#include <cstring>

#include "cg_kernels.cuh"
#include "cg_launch.h"
#include "cg_tma.cuh"

namespace acg {

namespace {

constexpr int TY = 8;     // tile rows
constexpr int S = 3;      // TMA stages in flight
constexpr int RING = 4;   // p_new planes kept

// Tile configuration: TX cells along x, TX/2 lanes x (TY / RPT) thread rows; a thread owns a
// pair of cells in RPT consecutive rows.
//   TX=128, RPT=2: 256 threads, 104 KB shared memory, 2 CTAs per SM, <= 128 registers
//   TX=128, RPT=1: 512 threads, same tile: twice the warps per SM (32) to hide the latency of
//                  the coefficient streams, at <= 64 registers per thread (APHCG_RPT=1)
//   TX=64 : 128 threads,  54 KB shared memory, 4 CTAs per SM (more independent
//           phases per SM to overlap one CTA's barrier/compute with another's loads)
template <int TX_, int RPT_ = 2>
struct Cfg {
  static constexpr int TX = TX_;
  static constexpr int RPT = RPT_;
  static constexpr int NT = TX * TY / (2 * RPT);
  static constexpr int LXN = TX / 2;  // threads along x
  static constexpr int MINB = TX == 128 ? 2 : 4;
  static constexpr int BW = TX + 4;   // box width: x0-2 .. x0+TX+1 (inner pairs 16-byte aligned)
  static constexpr int BH = TY + 2;   // box height: y0-1 .. y0+TY
  static constexpr int BOX = BW * BH;                        // doubles per box
  static constexpr int BOXB = (BOX * 8 + 127) / 128 * 128;   // bytes, 128-aligned for TMA
  static constexpr int kTxBytes = 2 * BOX * 8;
  static constexpr int kSmemBytes = S * 2 * BOXB + RING * BOXB + S * 8 + 128;
};

// kSym: the matrix is exactly symmetric (checked at upload), so the coefficient
// towards the upper neighbour is read as the lower coefficient OF that
// neighbour: 4 coefficient streams instead of 7 (x+ comes from the next lane by
// warp shuffle, y+ from the next row, z+ is carried in registers to become z- of
// the next plane).
// L2 prefetch of one row segment (bytes: multiple of 16), no register or
// shared-memory cost: lets the HBM->L2 stream of the read-once arrays run `pd`
// planes ahead of the CTA's compute phase.
// kDefer (symmetric storage only; OPT-IN via APHCG_DEFER=1, prepared at the end of round 1
// without GPU access and therefore not the default): everything that merely CONSUMES a freshly
// loaded coefficient -- the lane shuffle for x+, the y-/z- register aliases -- is moved from the
// load section to just in front of the stencil, behind the plane barrier, so that a warp issues
// all loads of a step back to back and their latency overlaps the p_new phase
// (profiles/r01d_dir_spmv_stalls_by_line.md: 22 % of the stall samples sit on those consumers).
// Same values into the same FMAs.
template <int TXT, bool kSingle, bool kSym, bool kDefer = false, int RPT = 2>
__global__ void __launch_bounds__(Cfg<TXT, RPT>::NT, Cfg<TXT, RPT>::MINB)
    k_dir_spmv_tma(const Geom g, const DevPtrs d, const int zc, const int pd, const int opts,
                   const __grid_constant__ CUtensorMap map_r,
                   const __grid_constant__ CUtensorMap map_p0,
                   const __grid_constant__ CUtensorMap map_p1) {
  using C = Cfg<TXT, RPT>;
  constexpr int TX = C::TX, NT = C::NT, BW = C::BW, BOX = C::BOX, BOXB = C::BOXB;
  // Dynamic shared memory, indexed only through pointers derived from the
  // __shared__ symbol itself so that every access compiles to LDS/STS (a detour
  // through an integer cast makes ptxas fall back to generic LD/ST).
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  __shared__ double sm_red[32];
  __shared__ int sm_flag;
  __shared__ DirView view;
  constexpr int BOXD = BOXB / 8;  // doubles per staged box (padded to 128 bytes)
  double* const smd = reinterpret_cast<double*>(smem_raw);
  // layout: [S x (r box | p_old box)] [RING x p_new box] [S mbarriers]
  uint64_t* const full = reinterpret_cast<uint64_t*>(smem_raw + (2 * S + RING) * BOXB);
  CgState* st = d.st;
  if (st->done) return;
  // loop scalars; with several GPUs this is where the CTA waits for the all-reduced r.r of the
  // previous update stage (and, by the same acquire, for the neighbours' boundary planes of r)
  dir_view(d, &view);
  const double beta = view.beta;
  const double alpha_prev = view.alpha_prev, alpha_prev2 = view.alpha_prev2;
  const int par = view.iter & 1;
  const CUtensorMap* map_po = par ? &map_p1 : &map_p0;
  // p_new goes where p_{k-2} lives; the batched x update reads it from there first
  double* pn_glob = d.p[par ^ 1];
  // deferred x updates applied by this launch (CgState::xbatch): 1 = the last one (every
  // iteration), 2 = the last two (even iterations), 0 = none (odd iterations; iteration 0)
  const int xmode = !st->xbatch ? 1 : ((par || view.iter == 0) ? 0 : 2);

  const int tid = threadIdx.x;
  const bool pstream = (opts & 1) != 0;  // streaming (evict-first) stores of p_new
  const int x0 = blockIdx.x * TX, y0 = blockIdx.y * TY;
  const int k0 = blockIdx.z * zc;
  const int k1 = min(k0 + zc, g.nzl);
  // planes k0-1 .. k1; none when the folded exit rule has just fired (the last CTA still
  // commits it below)
  const int np = view.done ? 0 : k1 - k0 + 2;

  // TMA coordinates of the box origin (element units of the padded tensor)
  const int c0 = kGhostX + x0 - 2, c1 = y0;  // row index 1+(y0-1)
  auto issue = [&](int n) {
    const int s = n % S;
    mbar_arrive_expect_tx(&full[s], C::kTxBytes);
    tma_load_3d(smd + (2 * s) * BOXD, &map_r, &full[s], c0, c1, k0 + n);  // plane 1+(k0-1+n)
    tma_load_3d(smd + (2 * s + 1) * BOXD, map_po, &full[s], c0, c1, k0 + n);
  };
  if (tid == 0) {
    for (int s = 0; s < S; ++s) mbar_init(&full[s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncthreads();
  if (tid == 0) {
    for (int n = 0; n < S && n < np; ++n) issue(n);
  }

  // cells of this thread: pair (2*lx, 2*lx+1) in rows RPT*ly .. RPT*ly+RPT-1 of the tile
  const int lx = tid % C::LXN, ly = tid / C::LXN;
  const int lane = tid & 31;
  const int ci = x0 + 2 * lx;
  const bool act_x = ci < g.nx;
  bool act[RPT];
  int cj[RPT];
#pragma unroll
  for (int h = 0; h < RPT; ++h) {
    cj[h] = y0 + RPT * ly + h;
    act[h] = act_x && cj[h] < g.ny;
  }
  // ownership of box columns/rows for the p_new stores (interior + face ghosts)
  const int own_xlo = x0 - (x0 == 0 ? 1 : 0);
  const int own_xhi = min(x0 + TX, g.nx) + (x0 + TX >= g.nx ? 1 : 0);  // exclusive
  const int own_ylo = y0 - (y0 == 0 ? 1 : 0);
  const int own_yhi = min(y0 + TY, g.ny) + (y0 + TY >= g.ny ? 1 : 0);

  double acc = 0.0;
  Vec<2> zcarry[RPT];
  Vec<2> p2n[RPT];
#pragma unroll
  for (int h = 0; h < RPT; ++h) {
    zcarry[h].v[0] = zcarry[h].v[1] = 0.0;
    p2n[h] = zcarry[h];
  }
  for (int n = 0; n < np; ++n) {
    const int z = k0 - 1 + n;  // plane whose p_new is formed in this step
    const int m = z - 1;       // plane whose stencil is evaluated in this step
    const bool do_stencil = (n >= 2);
    const bool z_inner = (z >= k0 && z < k1);

    // -- L2 prefetch of the coefficient / x rows `pd` planes ahead ---------------------
    if (pd > 0) {
      constexpr int NARR = kSym ? 5 : 8;  // coefficient arrays (+ x) to prefetch
      if (tid < NARR * TY) {
        const int arr = tid / TY, row = tid - arr * TY;
        const int jj = y0 + row;
        const int w = min(TX, g.nx - x0);
        // stencil plane of step n+pd is m+pd; x plane is z+pd
        const int pl = (arr == NARR - 1) ? z + pd : m + pd;
        const bool ok = jj < g.ny && ((arr == NARR - 1) ? (xmode != 0 && pl >= k0 && pl < k1)
                                                        : (pl >= k0 && pl < k1 + (kSym ? 1 : 0) && pl < g.nzl));
        if (ok) {
          const double* basep;
          if (arr == NARR - 1) {
            basep = d.u;
            if (xmode == 2 && pl + 1 < k1)  // p_{k-2} of the same cells (loaded one step early)
              prefetch_l2(pn_glob + g.poff + x0 + (int64_t)jj * g.py + (int64_t)(pl + 1) * g.pz,
                          (unsigned)w * 8u);
          } else if (kSym) {
            basep = d.a[arr == 0 ? 0 : 2 * arr - 1];  // a0, a1 (x-), a3 (y-), a5 (z-)
          } else {
            basep = d.a[arr];
          }
          prefetch_l2(basep + x0 + jj * g.cy + (int64_t)pl * g.cz, (unsigned)w * 8u);
        }
      }
    }

    // -- issue the read-once streams of this step before any waiting ---------------
    Vec<2> a[RPT][7];
    Vec<2> uu[RPT];
    if constexpr (!kSym) {
#pragma unroll
      for (int h = 0; h < RPT; ++h) {
        if (do_stencil && act[h]) {
          const int64_t idc = ci + cj[h] * g.cy + (int64_t)m * g.cz;
#pragma unroll
          for (int q = 0; q < 7; ++q) a[h][q] = ldv_stream<2>(d.a[q] + idc);
        }
      }
    } else {
      // a[h][1], a[h][3], a[h][5] <- lower coefficients of the own cells;
      // a[h][2], a[h][4], a[h][6] <- lower coefficients of the upper neighbours
      Vec<2> zero2;
      zero2.v[0] = zero2.v[1] = 0.0;
#pragma unroll
      for (int h = 0; h < RPT; ++h) a[h][1] = zero2;
      if (do_stencil) {
#pragma unroll
        for (int h = 0; h < RPT; ++h) {
          if (act[h]) {
            const int64_t idc = ci + cj[h] * g.cy + (int64_t)m * g.cz;
            a[h][0] = ldv_stream<2>(d.a[0] + idc);
            a[h][1] = ldv_stream<2>(d.a[1] + idc);
            if constexpr (!kDefer) {
              a[h][3] = (h == 1) ? a[0][4] : ldv_stream<2>(d.a[3] + idc);
            } else {
              if (h == 0) a[0][3] = ldv_stream<2>(d.a[3] + idc);  // a[1][3] = a[0][4]: later
            }
            // y+: next row's y-, or this row's own y+ on the last row of the domain
            a[h][4] = (cj[h] + 1 < g.ny) ? ldv<2>(d.a[3] + idc + g.cy) : ldv_stream<2>(d.a[4] + idc);
            // z-: carried from the previous plane (its z+), except on the first plane
            a[h][5] = (n == 2) ? ldv_stream<2>(d.a[5] + idc) : zcarry[h];
            a[h][6] = (m + 1 < g.nzl) ? ldv_stream<2>(d.a[5] + idc + g.cz) : ldv_stream<2>(d.a[6] + idc);
            if constexpr (!kDefer) {
              zcarry[h] = a[h][6];
            } else {
              // the two x+ values that do not come from the next lane are loads: issue them now
              if (ci + 2 >= g.nx) {
                a[h][2].v[1] = d.a[2][idc + 1];  // last cell of the row: its own x+
              } else if (lane == 31) {
                a[h][2].v[1] = d.a[1][idc + 2];  // next lane lives in another warp / CTA
              }
            }
          }
        }
        // x+ of the second cell of the pair = x- of the next lane's first cell
#pragma unroll
        for (int h = 0; h < RPT && !kDefer; ++h) {
          const double from_next = __shfl_down_sync(0xffffffffu, a[h][1].v[0], 1);
          if (act[h]) {
            const int64_t idc = ci + cj[h] * g.cy + (int64_t)m * g.cz;
            a[h][2].v[0] = a[h][1].v[1];
            if (ci + 2 >= g.nx) {
              a[h][2].v[1] = d.a[2][idc + 1];  // last cell of the row: its own x+
            } else if (lane == 31) {
              a[h][2].v[1] = d.a[1][idc + 2];  // next lane lives in another warp / CTA
            } else {
              a[h][2].v[1] = from_next;
            }
          }
        }
      }
    }
#pragma unroll
    for (int h = 0; h < RPT; ++h) {
      if (z_inner && xmode != 0 && act[h]) {
        uu[h] = ldv_stream<2>(d.u + ci + cj[h] * g.cy + (int64_t)z * g.cz);
      }
    }
    // p_{k-2} of plane z+1, consumed in the NEXT step: the __syncthreads that ends this step
    // orders the load before any thread of the CTA overwrites those cells with p_new(z+1)
    Vec<2> p2c[RPT];
#pragma unroll
    for (int h = 0; h < RPT; ++h) {
      p2c[h] = p2n[h];
      if (xmode == 2 && z + 1 >= k0 && z + 1 < k1 && act[h]) {
        p2n[h] = ldv_stream<2>(pn_glob + g.poff + ci + (int64_t)cj[h] * g.py + (int64_t)(z + 1) * g.pz);
      }
    }

    const int s = n % S;
    mbar_wait(&full[s], (n / S) & 1);
    const double* sr = smd + (2 * s) * BOXD;
    const double* sp = smd + (2 * s + 1) * BOXD;
    double* rg = smd + (2 * S + n % RING) * BOXD;

    // -- (a) p_new = r + beta*p_old on the whole box; store what this CTA owns --------
    const bool z_owned = z_inner || (z == -1 && k0 == 0) || (z == g.nzl && k1 == g.nzl);
    for (int e = tid; e < BOX / 2; e += NT) {
      const double2 rv = *reinterpret_cast<const double2*>(sr + 2 * e);
      const double2 pv = *reinterpret_cast<const double2*>(sp + 2 * e);
      double2 o;
      o.x = fma(beta, pv.x, rv.x);  // linear.ipp:97-99
      o.y = fma(beta, pv.y, rv.y);
      *reinterpret_cast<double2*>(rg + 2 * e) = o;
      if (z_owned) {
        const int row = e / (BW / 2), cp = e - row * (BW / 2);
        const int y = y0 - 1 + row, x = x0 - 2 + 2 * cp;
        if (y >= own_ylo && y < own_yhi) {
          const bool in0 = (x >= own_xlo && x < own_xhi);
          const bool in1 = (x + 1 >= own_xlo && x + 1 < own_xhi);
          double* dst = pn_glob + g.poff + x + (int64_t)y * g.py + (int64_t)z * g.pz;
          if (in0 && in1) {
            if (pstream) {
              __stcs(reinterpret_cast<double2*>(dst), o);  // next read is an iteration away
            } else {
              *reinterpret_cast<double2*>(dst) = o;
            }
          } else if (in0) {
            dst[0] = o.x;
          } else if (in1) {
            dst[1] = o.y;
          }
        }
      }
    }
    // -- (b) deferred x update (linear.ipp:88) with p_old of the own cells -------------
    if (z_inner && xmode != 0) {
#pragma unroll
      for (int h = 0; h < RPT; ++h) {
        if (act[h]) {
          const int o = (RPT * ly + h + 1) * BW + 2 + 2 * lx;
          const double2 pv = *reinterpret_cast<const double2*>(sp + o);
          if (xmode == 2) {  // the older update first: same FMAs, same order as one per iteration
            uu[h].v[0] = fma(alpha_prev2, p2c[h].v[0], uu[h].v[0]);
            uu[h].v[1] = fma(alpha_prev2, p2c[h].v[1], uu[h].v[1]);
          }
          uu[h].v[0] = fma(alpha_prev, pv.x, uu[h].v[0]);
          uu[h].v[1] = fma(alpha_prev, pv.y, uu[h].v[1]);
          stv_stream<2>(d.u + ci + cj[h] * g.cy + (int64_t)z * g.cz, uu[h]);
        }
      }
    }
    __syncthreads();  // p_new(z) visible; stage s fully consumed
    if (tid == 0 && n + S < np) issue(n + S);

    // -- (d) stencil of plane m = z-1 from the ring -------------------------------------
    if (do_stencil) {
      const double* rc = smd + (2 * S + (n + RING - 1) % RING) * BOXD;  // plane m
      const double* rm = smd + (2 * S + (n + RING - 2) % RING) * BOXD;  // plane m-1
      const double* rp = rg;                                            // plane m+1
      if constexpr (kSym && kDefer) {
        // the consumers of this step's coefficient loads, deferred to here (see kDefer)
#pragma unroll
        for (int h = 0; h < RPT; ++h) {
          const double from_next = __shfl_down_sync(0xffffffffu, a[h][1].v[0], 1);
          if (act[h]) {
            if (h == 1) a[h][3] = a[0][4];
            a[h][2].v[0] = a[h][1].v[1];
            if (ci + 2 < g.nx && lane != 31) a[h][2].v[1] = from_next;
          }
        }
      }
#pragma unroll
      for (int h = 0; h < RPT; ++h) {
        if (act[h]) {
          const int o = (RPT * ly + h + 1) * BW + 2 + 2 * lx;
          const double2 pc = *reinterpret_cast<const double2*>(rc + o);
          const double pxm = rc[o - 1], pxp = rc[o + 2];
          const double2 pym = *reinterpret_cast<const double2*>(rc + o - BW);
          const double2 pyp = *reinterpret_cast<const double2*>(rc + o + BW);
          const double2 pzm = *reinterpret_cast<const double2*>(rm + o);
          const double2 pzp = *reinterpret_cast<const double2*>(rp + o);
          Vec<2> ap;
          // accumulation order of the reference: centre, then q = 0..5 (linear.ipp:67-70)
          double t = pc.x * a[h][0].v[0];
          t = fma(pxm, a[h][1].v[0], t);
          t = fma(pc.y, a[h][2].v[0], t);
          t = fma(pym.x, a[h][3].v[0], t);
          t = fma(pyp.x, a[h][4].v[0], t);
          t = fma(pzm.x, a[h][5].v[0], t);
          t = fma(pzp.x, a[h][6].v[0], t);
          ap.v[0] = t;
          acc = fma(pc.x, t, acc);
          t = pc.y * a[h][0].v[1];
          t = fma(pc.x, a[h][1].v[1], t);
          t = fma(pxp, a[h][2].v[1], t);
          t = fma(pym.y, a[h][3].v[1], t);
          t = fma(pyp.y, a[h][4].v[1], t);
          t = fma(pzm.y, a[h][5].v[1], t);
          t = fma(pzp.y, a[h][6].v[1], t);
          ap.v[1] = t;
          acc = fma(pc.y, t, acc);
          stv_stream<2>(d.ap + ci + cj[h] * g.cy + (int64_t)m * g.cz, ap);
          if constexpr (kSym && kDefer) zcarry[h] = a[h][6];  // z+ becomes z- of the next plane
        }
      }
    }
  }

  const double bsum = block_reduce<false>(acc, sm_red);
  const unsigned nblk = gridDim.x * gridDim.y * gridDim.z;
  const unsigned bid = blockIdx.x + gridDim.x * (blockIdx.y + gridDim.y * blockIdx.z);
  if (tid == 0) d.partials[bid] = bsum;
  if (last_block(&st->counter_a, nblk, &sm_flag)) {
    const double tot = reduce_slots<false>(d.partials, nblk, sm_red);
    dir_epilogue<kSingle>(d, view, tot);
  }
}

using EncodeFn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                              const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                              const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                              CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

}  // namespace

struct TmaPlan {
  alignas(64) CUtensorMap map_r, map_p0, map_p1;
  dim3 grid;
  int zc;
  int pd;  // L2 prefetch distance in planes (0 = off)
  int opts;  // bit 0: streaming stores of p_new; bit 1: kDefer variant of the symmetric kernel
  int tx;  // tile width in use (64 or 128)
  int rpt;  // rows per thread: 2, or 1 (512-thread CTAs; 128-wide tiles, symmetric storage only)
};

template <int TXT>
static bool set_smem_limit() {
  using C = Cfg<TXT>;
  return cudaFuncSetAttribute(k_dir_spmv_tma<TXT, true, false>,
                              cudaFuncAttributeMaxDynamicSharedMemorySize, C::kSmemBytes) == cudaSuccess &&
         cudaFuncSetAttribute(k_dir_spmv_tma<TXT, false, false>,
                              cudaFuncAttributeMaxDynamicSharedMemorySize, C::kSmemBytes) == cudaSuccess &&
         cudaFuncSetAttribute(k_dir_spmv_tma<TXT, true, true>,
                              cudaFuncAttributeMaxDynamicSharedMemorySize, C::kSmemBytes) == cudaSuccess &&
         cudaFuncSetAttribute(k_dir_spmv_tma<TXT, false, true>,
                              cudaFuncAttributeMaxDynamicSharedMemorySize, C::kSmemBytes) == cudaSuccess &&
         cudaFuncSetAttribute(k_dir_spmv_tma<TXT, true, true, true>,
                              cudaFuncAttributeMaxDynamicSharedMemorySize, C::kSmemBytes) == cudaSuccess &&
         cudaFuncSetAttribute(k_dir_spmv_tma<TXT, false, true, true>,
                              cudaFuncAttributeMaxDynamicSharedMemorySize, C::kSmemBytes) == cudaSuccess &&
         cudaFuncSetAttribute(k_dir_spmv_tma<TXT, true, true, false, 1>,
                              cudaFuncAttributeMaxDynamicSharedMemorySize, C::kSmemBytes) == cudaSuccess &&
         cudaFuncSetAttribute(k_dir_spmv_tma<TXT, false, true, false, 1>,
                              cudaFuncAttributeMaxDynamicSharedMemorySize, C::kSmemBytes) == cudaSuccess;
}

TmaPlan* tma_plan_create(const Geom& g, const DevPtrs& d, char* err, int errlen) {
  auto fail = [&](const char* msg) -> TmaPlan* {
    if (err && errlen > 0) snprintf(err, errlen, "%s", msg);
    return nullptr;
  };
  if (g.nx % 2) return fail("nx is odd");
  EncodeFn encode = nullptr;
  cudaDriverEntryPointQueryResult qres;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void**)&encode, cudaEnableDefault,
                              &qres) != cudaSuccess ||
      qres != cudaDriverEntryPointSuccess || !encode) {
    cudaGetLastError();
    return fail("cuTensorMapEncodeTiled not available");
  }
  TmaPlan* p = new TmaPlan();
  p->tx = (g.nx <= 64) ? 64 : 128;
  if (const char* et = getenv("APHCG_TILE")) p->tx = (atoi(et) == 64) ? 64 : 128;
  const int TX = p->tx;
  const int BW = TX + 4, BH = TY + 2;
  const cuuint64_t gdim[3] = {(cuuint64_t)g.py, (cuuint64_t)(g.ny + 2), (cuuint64_t)(g.nzl + 2)};
  const cuuint64_t gstr[2] = {(cuuint64_t)g.py * 8, (cuuint64_t)g.pz * 8};
  const cuuint32_t box[3] = {(cuuint32_t)BW, (cuuint32_t)BH, 1};
  const cuuint32_t estr[3] = {1, 1, 1};
  double* bases[3] = {d.r, d.p[0], d.p[1]};
  CUtensorMap* maps[3] = {&p->map_r, &p->map_p0, &p->map_p1};
  for (int i = 0; i < 3; ++i) {
    const CUresult rc = encode(maps[i], CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, bases[i], gdim, gstr,
                               box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                               CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (rc != CUDA_SUCCESS) {
      delete p;
      char msg[96];
      snprintf(msg, sizeof(msg), "cuTensorMapEncodeTiled failed (CUresult %d)", (int)rc);
      return fail(msg);
    }
  }
  // Planes per CTA.  A column of z planes costs about z + 4 plane steps (2 halo
  // planes + pipeline fill) and the chip runs 296 columns at a time (2 CTAs/SM), so
  // the kernel takes about ceil(ctas/296) * (z + 4) steps: pick the z that minimises
  // it (ties: the longer column).  Agrees with the measurements: 128^3 8 planes x 256
  // CTAs beat 4 x 512 (46 vs 52 us), 256^3 16 x 1024 beat 8 x 2048 (0.291 vs 0.306 ms),
  // 512^3 32 x 4096 = 64 x 2048.
  const int tiles = ((g.nx + TX - 1) / TX) * ((g.ny + TY - 1) / TY);
  int zc = 32;
  if (const char* env = getenv("APHCG_ZC")) {
    zc = atoi(env);
  } else {
    // 64 planes per CTA only when that makes the whole grid ONE wave (measured, round 2:
    // 64x512x512 0.261 vs 0.271 ms, 256^3 0.270 vs 0.280; but 128x512x512 -- two waves
    // either way -- 0.526 vs 0.507).
    int64_t best = -1;
    const int64_t slots = 2 * kNumSMs;  // resident CTAs of this kernel (2 per SM)
    for (int z = 64; z >= 2; z /= 2) {
      const int64_t n = (int64_t)tiles * ((g.nzl + z - 1) / z);
      if (z == 64 && n > slots) continue;
      const int64_t cost = ((n + slots - 1) / slots) * (z + 4);
      if (best < 0 || cost < best) {
        best = cost;
        zc = z;
      }
    }
  }
  if (zc < 1) zc = 1;
  if (zc > g.nzl) zc = g.nzl;
  p->zc = zc;
  p->pd = 2;
  if (const char* ep = getenv("APHCG_PREFETCH")) p->pd = atoi(ep);
  p->opts = 1;  // streaming p_new stores: 1.87 vs 1.90 ms at 512^3
  if (const char* eo = getenv("APHCG_PSTREAM")) p->opts = atoi(eo) ? 1 : 0;
  if (const char* ed = getenv("APHCG_DEFER")) p->opts |= atoi(ed) ? 2 : 0;
  // one row pair per thread: 512-thread CTAs, twice the warps per SM at <= 64 registers.
  // Measured round 2 (profiles/r02_rows_per_thread_sweep.txt): 1.83-1.85 vs 1.89 ms at 512^3.
  p->rpt = 1;
  if (const char* er = getenv("APHCG_RPT")) p->rpt = atoi(er) == 2 ? 2 : 1;
  p->grid = dim3((g.nx + TX - 1) / TX, (g.ny + TY - 1) / TY, (g.nzl + zc - 1) / zc);
  if (!(TX == 64 ? set_smem_limit<64>() : set_smem_limit<128>())) {
    cudaGetLastError();
    delete p;
    return fail("cannot raise the dynamic shared memory limit");
  }
  return p;
}

void tma_plan_destroy(TmaPlan* p) { delete p; }
unsigned tma_plan_blocks(const TmaPlan* p) { return p->grid.x * p->grid.y * p->grid.z; }
void tma_plan_describe(const TmaPlan* p, char* buf, int buflen) {
  snprintf(buf, buflen, "tile=%dx%d planes_per_cta=%d stages=%d l2_prefetch=%d pstream=%d%s%s ctas=%u",
           p->tx, TY, p->zc, S, p->pd, p->opts & 1, (p->opts & 2) ? " defer=1" : "",
           p->rpt == 1 ? " rows_per_thread=1" : "", tma_plan_blocks(p));
}

template <int TXT>
static void launch_cfg(const TmaPlan* p, const Geom& g, const DevPtrs& d, bool single, bool sym,
                       cudaStream_t s) {
  using C = Cfg<TXT>;
#define APHCG_LAUNCH_TMA(SINGLE, SYM, DEFER)                                                    \
  k_dir_spmv_tma<TXT, SINGLE, SYM, DEFER><<<p->grid, C::NT, C::kSmemBytes, s>>>(                  \
      g, d, p->zc, p->pd, p->opts, p->map_r, p->map_p0, p->map_p1)
  const bool defer = sym && (p->opts & 2) != 0;
  if (sym && !defer && p->rpt == 1) {
    using C1 = Cfg<TXT, 1>;
    if (single) {
      k_dir_spmv_tma<TXT, true, true, false, 1><<<p->grid, C1::NT, C1::kSmemBytes, s>>>(
          g, d, p->zc, p->pd, p->opts, p->map_r, p->map_p0, p->map_p1);
    } else {
      k_dir_spmv_tma<TXT, false, true, false, 1><<<p->grid, C1::NT, C1::kSmemBytes, s>>>(
          g, d, p->zc, p->pd, p->opts, p->map_r, p->map_p0, p->map_p1);
    }
    return;
  }
  if (single) {
    if (defer) APHCG_LAUNCH_TMA(true, true, true);
    else if (sym) APHCG_LAUNCH_TMA(true, true, false);
    else APHCG_LAUNCH_TMA(true, false, false);
  } else {
    if (defer) APHCG_LAUNCH_TMA(false, true, true);
    else if (sym) APHCG_LAUNCH_TMA(false, true, false);
    else APHCG_LAUNCH_TMA(false, false, false);
  }
#undef APHCG_LAUNCH_TMA
}

void launch_dir_spmv_tma(const TmaPlan* p, const Geom& g, const DevPtrs& d, bool single, bool sym,
                         cudaStream_t s) {
  if (p->tx == 64) {
    launch_cfg<64>(p, g, d, single, sym, s);
  } else {
    launch_cfg<128>(p, g, d, single, sym, s);
  }
}

}  // namespace acg
