// k_dir_spmv_stream: the direction + SpMV kernel with EVERY operand staged in shared memory by
// TMA -- the r / p_old boxes as in k_dir_spmv_tma, and also the four coefficient streams of the
// symmetric storage, x and p_{k-2}.  Same arithmetic, same FMA order, same results
// (src/linear/linear.ipp:64-101:  x += alpha_prev*p_old ; p = r + beta*p_old ; Ap = A p ; sum p.Ap).
//
// Why: in k_dir_spmv_tma the read-once streams go HBM/L2 -> registers with LDG issued at the
// start of a plane step and consumed half a step later; ncu attributes half of all warp stalls
// to that wait (long scoreboard at the first consumers + the CTA barrier behind which the
// slowest warp's loads hide; profiles/r02_dir_spmv_stalls.md), while the same access pattern
// without any waiting reaches 6.5-6.6 TB/s (profiles/r02_access_pattern_microbench.txt).  Here
// no warp ever waits for a global load: one elected thread issues bulk tensor copies two plane
// steps ahead, completion is signalled on mbarriers, and all 8 warps of the CTA only read shared
// memory.  The price is shared memory (109 KB per CTA of 256 threads, 2 CTAs per SM), which is
// why the tile is 64 x 8 and the stages are few.
//
// Per plane step n (z = k0-1+n is the plane whose p_new is formed, m = z-1 the plane whose
// stencil is evaluated):
//   front stage n % SF : r(z), p_old(z) boxes with halo; x(z), p_{k-2}(z) tiles (when used)
//   back  stage j % SB : a0(m), a1(m) (66 wide: x+ of a cell is x- of the next), a3(m) (9 rows:
//                        y+ is y- of the row above), a5(m+1) (z+ is z- of the plane above;
//                        z- is carried in a register from the previous step), j = n - 2
// Front stages are consumed before the step's one __syncthreads, back stages after it; thread 0
// refills, right after that barrier, the front stage just consumed and the back stage the
// PREVIOUS step's stencil consumed.
#include <algorithm>
#include <cstdio>
#include <cstring>

#include "cg_kernels.cuh"
#include "cg_launch.h"
#include "cg_tma.cuh"

namespace acg {

namespace {

constexpr int TY = 8;
constexpr int SF = 2;    // front stages
constexpr int SB = 3;    // back stages
constexpr int RING = 4;  // p_new planes kept

constexpr int align128(int b) { return (b + 127) / 128 * 128; }

template <int TX_>
struct Cfg2 {
  static constexpr int TX = TX_;
  static constexpr int NT = TX * TY / 2;  // one pair of cells per thread
  static constexpr int LXN = TX / 2;
  static constexpr int BW = TX + 4, BH = TY + 2;
  static constexpr int BOX = BW * BH;
  static constexpr int BOXB = align128(BOX * 8);
  static constexpr int TILEB = TX * TY * 8;                  // a0, a5, x, p_{k-2}
  static constexpr int A1W = TX + 2;
  static constexpr int A1B = A1W * TY * 8, A1BP = align128(A1B);
  static constexpr int A3B = TX * (TY + 1) * 8, A3BP = align128(A3B);
  static constexpr int FRONTB = 2 * BOXB + 2 * TILEB;        // r | p_old | x | p_{k-2}
  static constexpr int BACKB = TILEB + A1BP + A3BP + TILEB;  // a0 | a1 | a3 | a5
  static constexpr int kSmemBytes = SF * FRONTB + SB * BACKB + RING * BOXB + (SF + SB) * 8 + 128;
  static constexpr int MINB = 2;
};

struct Maps2 {
  CUtensorMap r, p0, p1;  // padded fields, box (TX+4) x (TY+2)
  CUtensorMap q0, q1;     // padded p fields, box TX x TY (own cells: p_{k-2})
  CUtensorMap a0, a1, a3, a5, x;  // compact arrays: TX x TY, (TX+2) x TY, TX x (TY+1), TX x TY, TX x TY
};

template <int TXT, bool kSingle>
__global__ void __launch_bounds__(Cfg2<TXT>::NT, Cfg2<TXT>::MINB)
    k_dir_spmv_stream(const Geom g, const DevPtrs d, const int zc, const int pd, const int opts,
                      const __grid_constant__ Maps2 mp) {
  using C = Cfg2<TXT>;
  constexpr int TX = C::TX, NT = C::NT, BW = C::BW, BOX = C::BOX;
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  __shared__ double sm_red[32];
  __shared__ int sm_flag;
  __shared__ DirView view;
  // layout: [SF x front] [SB x back] [RING x p_new box] [SF + SB mbarriers]
  unsigned char* const front0 = smem_raw;
  unsigned char* const back0 = smem_raw + SF * C::FRONTB;
  unsigned char* const ring0 = back0 + SB * C::BACKB;
  uint64_t* const fbar = reinterpret_cast<uint64_t*>(ring0 + RING * C::BOXB);
  uint64_t* const bbar = fbar + SF;
  CgState* st = d.st;
  if (st->done) return;
  dir_view(d, &view);
  const double beta = view.beta;
  const double alpha_prev = view.alpha_prev, alpha_prev2 = view.alpha_prev2;
  const int par = view.iter & 1;
  const CUtensorMap* map_po = par ? &mp.p1 : &mp.p0;
  const CUtensorMap* map_q = par ? &mp.q0 : &mp.q1;  // p_{k-2} lives where p_new goes
  double* pn_glob = d.p[par ^ 1];
  const int xmode = !st->xbatch ? 1 : ((par || view.iter == 0) ? 0 : 2);

  const int tid = threadIdx.x;
  const bool pstream = (opts & 1) != 0;
  const int x0 = blockIdx.x * TX, y0 = blockIdx.y * TY;
  const int k0 = blockIdx.z * zc;
  const int k1 = min(k0 + zc, g.nzl);
  const int np = view.done ? 0 : k1 - k0 + 2;  // front loads / plane steps: planes k0-1 .. k1
  const int nj = view.done ? 0 : k1 - k0;      // back loads / stencil steps: planes k0 .. k1-1

  auto issue_front = [&](int n) {
    const int s = n % SF;
    const int z = k0 - 1 + n;
    const bool inner = (z >= k0 && z < k1);
    const bool want_x = inner && xmode != 0, want_q = inner && xmode == 2;
    unsigned char* base = front0 + s * C::FRONTB;
    mbar_arrive_expect_tx(&fbar[s], 2 * BOX * 8 + (want_x ? C::TILEB : 0) + (want_q ? C::TILEB : 0));
    tma_load_3d(base, &mp.r, &fbar[s], kGhostX + x0 - 2, y0, k0 + n);
    tma_load_3d(base + C::BOXB, map_po, &fbar[s], kGhostX + x0 - 2, y0, k0 + n);
    if (want_x) tma_load_3d(base + 2 * C::BOXB, &mp.x, &fbar[s], x0, y0, z);
    if (want_q) tma_load_3d(base + 2 * C::BOXB + C::TILEB, map_q, &fbar[s], kGhostX + x0, 1 + y0, 1 + z);
  };
  auto issue_back = [&](int j) {
    const int s = j % SB;
    const int m = k0 + j;
    const bool want_zp = m + 1 < g.nzl;
    unsigned char* base = back0 + s * C::BACKB;
    mbar_arrive_expect_tx(&bbar[s], C::TILEB + C::A1B + C::A3B + (want_zp ? C::TILEB : 0));
    tma_load_3d(base, &mp.a0, &bbar[s], x0, y0, m);
    tma_load_3d(base + C::TILEB, &mp.a1, &bbar[s], x0, y0, m);
    tma_load_3d(base + C::TILEB + C::A1BP, &mp.a3, &bbar[s], x0, y0, m);
    if (want_zp) tma_load_3d(base + C::TILEB + C::A1BP + C::A3BP, &mp.a5, &bbar[s], x0, y0, m + 1);
  };
  if (tid == 0) {
    for (int s = 0; s < SF; ++s) mbar_init(&fbar[s], 1);
    for (int s = 0; s < SB; ++s) mbar_init(&bbar[s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncthreads();
  if (tid == 0) {
    for (int n = 0; n < SF && n < np; ++n) issue_front(n);
    for (int j = 0; j < SB && j < nj; ++j) issue_back(j);
  }

  // cells of this thread: the pair (2*lx, 2*lx+1) of tile row ly
  const int lx = tid % C::LXN, ly = tid / C::LXN;
  const int ci = x0 + 2 * lx, cj = y0 + ly;
  const bool act = ci < g.nx && cj < g.ny;
  const int own_xlo = x0 - (x0 == 0 ? 1 : 0);
  const int own_xhi = min(x0 + TX, g.nx) + (x0 + TX >= g.nx ? 1 : 0);  // exclusive
  const int own_ylo = y0 - (y0 == 0 ? 1 : 0);
  const int own_yhi = min(y0 + TY, g.ny) + (y0 + TY >= g.ny ? 1 : 0);
  const int ob = (ly + 1) * BW + 2 + 2 * lx;  // own pair inside a box
  const int ot = ly * TX + 2 * lx;            // own pair inside a tile

  // The pairs of the (TX+4) x (TY+2) box this thread turns into p_new every plane, and what it
  // stores of them (interior + the face ghost cells of boundary tiles): the same for every
  // plane, so the index arithmetic is done once.  mode: -1 no pair, 0 compute only,
  // 1 first cell, 2 second cell, 3 both.
  constexpr int kPairs = (BOX / 2 + NT - 1) / NT;
  int pn_mode[kPairs];
  int64_t pn_off[kPairs];
#pragma unroll
  for (int k = 0; k < kPairs; ++k) {
    const int e = tid + k * NT;
    pn_mode[k] = -1;
    pn_off[k] = 0;
    if (e < BOX / 2) {
      const int row = e / (BW / 2), cp = e - row * (BW / 2);
      const int y = y0 - 1 + row, x = x0 - 2 + 2 * cp;
      const bool iny = (y >= own_ylo && y < own_yhi);
      const bool in0 = iny && (x >= own_xlo && x < own_xhi);
      const bool in1 = iny && (x + 1 >= own_xlo && x + 1 < own_xhi);
      pn_mode[k] = (in0 ? 1 : 0) | (in1 ? 2 : 0);
      pn_off[k] = g.poff + x + (int64_t)y * g.py;
    }
  }

  double acc = 0.0;
  Vec<2> zcarry;
  zcarry.v[0] = zcarry.v[1] = 0.0;
  for (int n = 0; n < np; ++n) {
    const int z = k0 - 1 + n;
    const int m = z - 1;
    const bool do_stencil = (n >= 2);
    const bool z_inner = (z >= k0 && z < k1);

    // -- L2 prefetch of the rows the TMA loads of step n + pd will want ------------------
    if (pd > 0 && tid < 5 * TY) {
      const int arr = tid / TY, row = tid - arr * TY;
      const int jj = y0 + row;
      const int w = min(TX, g.nx - x0);
      const int ahead = pd + (arr == 4 ? SF : SB);  // beyond what is already in flight
      const int pl = (arr == 4) ? z + ahead : m + ahead;
      const bool ok = jj < g.ny && ((arr == 4) ? (xmode != 0 && pl >= k0 && pl < k1)
                                               : (pl >= k0 && pl < k1 + 1 && pl < g.nzl));
      if (ok) {
        const double* basep = arr == 4 ? d.u : d.a[arr == 0 ? 0 : 2 * arr - 1];  // a0 a1 a3 a5 | x
        prefetch_l2(basep + x0 + jj * g.cy + (int64_t)pl * g.cz, (unsigned)w * 8u);
      }
    }

    // -- the few coefficients that do not come through the boxes (domain faces, first plane)
    Vec<2> edge_zm, edge_yp, edge_zp;
    double edge_xp = 0.0;
    if (do_stencil && act) {
      const int64_t idc = ci + cj * g.cy + (int64_t)m * g.cz;
      if (n == 2) edge_zm = ldv_stream<2>(d.a[5] + idc);               // z- of the first plane
      if (cj + 1 >= g.ny) edge_yp = ldv_stream<2>(d.a[4] + idc);       // last row: own y+
      if (m + 1 >= g.nzl) edge_zp = ldv_stream<2>(d.a[6] + idc);       // last plane: own z+
      if (ci + 2 >= g.nx) edge_xp = d.a[2][idc + 1];                   // last cell: own x+
    }

    const int sf = n % SF;
    mbar_wait(&fbar[sf], (n / SF) & 1);
    const unsigned char* fb = front0 + sf * C::FRONTB;
    const double* sr = reinterpret_cast<const double*>(fb);
    const double* sp = reinterpret_cast<const double*>(fb + C::BOXB);
    const double* sx = reinterpret_cast<const double*>(fb + 2 * C::BOXB);
    const double* sq = reinterpret_cast<const double*>(fb + 2 * C::BOXB + C::TILEB);
    double* rg = reinterpret_cast<double*>(ring0 + (n % RING) * C::BOXB);

    // -- (a) p_new = r + beta*p_old on the whole box; store what this CTA owns --------
    const bool z_owned = z_inner || (z == -1 && k0 == 0) || (z == g.nzl && k1 == g.nzl);
#pragma unroll
    for (int k = 0; k < kPairs; ++k) {
      if (pn_mode[k] < 0) continue;  // no such pair for this thread
      const int e2 = 2 * (tid + k * NT);
      const double2 rv = *reinterpret_cast<const double2*>(sr + e2);
      const double2 pv = *reinterpret_cast<const double2*>(sp + e2);
      double2 o;
      o.x = fma(beta, pv.x, rv.x);  // linear.ipp:97-99
      o.y = fma(beta, pv.y, rv.y);
      *reinterpret_cast<double2*>(rg + e2) = o;
      if (z_owned && pn_mode[k] > 0) {
        double* dst = pn_glob + pn_off[k] + (int64_t)z * g.pz;
        if (pn_mode[k] == 3) {
          if (pstream) {
            __stcs(reinterpret_cast<double2*>(dst), o);
          } else {
            *reinterpret_cast<double2*>(dst) = o;
          }
        } else if (pn_mode[k] == 1) {
          dst[0] = o.x;
        } else {
          dst[1] = o.y;
        }
      }
    }
    // -- (b) deferred x update (linear.ipp:88) with p_old of the own cells -------------
    if (z_inner && xmode != 0 && act) {
      const double2 xv = *reinterpret_cast<const double2*>(sx + ot);
      const double2 pv = *reinterpret_cast<const double2*>(sp + ob);
      Vec<2> uu;
      uu.v[0] = xv.x;
      uu.v[1] = xv.y;
      if (xmode == 2) {  // the older update first: same FMAs, same order as one per iteration
        const double2 qv = *reinterpret_cast<const double2*>(sq + ot);
        uu.v[0] = fma(alpha_prev2, qv.x, uu.v[0]);
        uu.v[1] = fma(alpha_prev2, qv.y, uu.v[1]);
      }
      uu.v[0] = fma(alpha_prev, pv.x, uu.v[0]);
      uu.v[1] = fma(alpha_prev, pv.y, uu.v[1]);
      stv_stream<2>(d.u + ci + cj * g.cy + (int64_t)z * g.cz, uu);
    }
    __syncthreads();  // p_new(z) visible; front stage sf consumed; last step's stencil done
    if (tid == 0) {
      if (n + SF < np) issue_front(n + SF);
      if (n >= 3 && (n - 3) + SB < nj) issue_back((n - 3) + SB);
    }

    // -- (c) stencil of plane m = z-1: p from the ring, coefficients from the back stage -
    if (do_stencil) {
      const int j = n - 2;
      const int sb = j % SB;
      mbar_wait(&bbar[sb], (j / SB) & 1);
      if (act) {
        const unsigned char* bb = back0 + sb * C::BACKB;
        const double* s0 = reinterpret_cast<const double*>(bb);
        const double* s1 = reinterpret_cast<const double*>(bb + C::TILEB);
        const double* s3 = reinterpret_cast<const double*>(bb + C::TILEB + C::A1BP);
        const double* s5 = reinterpret_cast<const double*>(bb + C::TILEB + C::A1BP + C::A3BP);
        const double* rc = reinterpret_cast<const double*>(ring0 + ((n + RING - 1) % RING) * C::BOXB);
        const double* rm = reinterpret_cast<const double*>(ring0 + ((n + RING - 2) % RING) * C::BOXB);
        const double* rp = rg;
        const double2 a0 = *reinterpret_cast<const double2*>(s0 + ot);
        const double2 a1 = *reinterpret_cast<const double2*>(s1 + ly * C::A1W + 2 * lx);
        const double a1n = (ci + 2 >= g.nx) ? edge_xp : s1[ly * C::A1W + 2 * lx + 2];
        const double2 a3 = *reinterpret_cast<const double2*>(s3 + ot);
        double2 a4, a5, a6;
        if (cj + 1 >= g.ny) {
          a4 = make_double2(edge_yp.v[0], edge_yp.v[1]);
        } else {
          a4 = *reinterpret_cast<const double2*>(s3 + ot + TX);
        }
        if (n == 2) {
          a5 = make_double2(edge_zm.v[0], edge_zm.v[1]);
        } else {
          a5 = make_double2(zcarry.v[0], zcarry.v[1]);
        }
        if (m + 1 >= g.nzl) {
          a6 = make_double2(edge_zp.v[0], edge_zp.v[1]);
        } else {
          a6 = *reinterpret_cast<const double2*>(s5 + ot);
        }
        zcarry.v[0] = a6.x;  // z+ becomes z- of the next plane
        zcarry.v[1] = a6.y;
        const double2 pc = *reinterpret_cast<const double2*>(rc + ob);
        const double pxm = rc[ob - 1], pxp = rc[ob + 2];
        const double2 pym = *reinterpret_cast<const double2*>(rc + ob - BW);
        const double2 pyp = *reinterpret_cast<const double2*>(rc + ob + BW);
        const double2 pzm = *reinterpret_cast<const double2*>(rm + ob);
        const double2 pzp = *reinterpret_cast<const double2*>(rp + ob);
        Vec<2> ap;
        // accumulation order of the reference: centre, then q = 0..5 (linear.ipp:67-70)
        double t = pc.x * a0.x;
        t = fma(pxm, a1.x, t);
        t = fma(pc.y, a1.y, t);  // x+ of the first cell = x- of the second
        t = fma(pym.x, a3.x, t);
        t = fma(pyp.x, a4.x, t);
        t = fma(pzm.x, a5.x, t);
        t = fma(pzp.x, a6.x, t);
        ap.v[0] = t;
        acc = fma(pc.x, t, acc);
        t = pc.y * a0.y;
        t = fma(pc.x, a1.y, t);
        t = fma(pxp, a1n, t);
        t = fma(pym.y, a3.y, t);
        t = fma(pyp.y, a4.y, t);
        t = fma(pzm.y, a5.y, t);
        t = fma(pzp.y, a6.y, t);
        ap.v[1] = t;
        acc = fma(pc.y, t, acc);
        stv_stream<2>(d.ap + ci + cj * g.cy + (int64_t)m * g.cz, ap);
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

}  // namespace

struct Tma2Plan {
  alignas(64) Maps2 maps;
  dim3 grid;
  int zc, pd, opts, tx;
};

Tma2Plan* tma2_plan_create(const Geom& g, const DevPtrs& d, char* err, int errlen) {
  auto fail = [&](const char* msg) -> Tma2Plan* {
    if (err && errlen > 0) snprintf(err, errlen, "%s", msg);
    return nullptr;
  };
  if (g.nx % 2) return fail("nx is odd");
  TensorMapEncodeFn encode = tensor_map_encoder();
  if (!encode) return fail("cuTensorMapEncodeTiled not available");
  Tma2Plan* p = new Tma2Plan();
  constexpr int TX = 64;
  p->tx = TX;
  auto make = [&](CUtensorMap* out, const double* base, bool padded, int bw, int bh) -> bool {
    cuuint64_t gdim[3], gstr[2];
    if (padded) {
      gdim[0] = (cuuint64_t)g.py;
      gdim[1] = (cuuint64_t)(g.ny + 2);
      gdim[2] = (cuuint64_t)(g.nzl + 2);
      gstr[0] = (cuuint64_t)g.py * 8;
      gstr[1] = (cuuint64_t)g.pz * 8;
    } else {
      gdim[0] = (cuuint64_t)g.nx;
      gdim[1] = (cuuint64_t)g.ny;
      gdim[2] = (cuuint64_t)g.nzl;
      gstr[0] = (cuuint64_t)g.cy * 8;
      gstr[1] = (cuuint64_t)g.cz * 8;
    }
    const cuuint32_t box[3] = {(cuuint32_t)bw, (cuuint32_t)bh, 1};
    const cuuint32_t estr[3] = {1, 1, 1};
    return encode(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, const_cast<double*>(base), gdim, gstr, box,
                  estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
  };
  Maps2& m = p->maps;
  const bool ok = make(&m.r, d.r, true, TX + 4, TY + 2) && make(&m.p0, d.p[0], true, TX + 4, TY + 2) &&
                  make(&m.p1, d.p[1], true, TX + 4, TY + 2) && make(&m.q0, d.p[0], true, TX, TY) &&
                  make(&m.q1, d.p[1], true, TX, TY) && make(&m.a0, d.a[0], false, TX, TY) &&
                  make(&m.a1, d.a[1], false, TX + 2, TY) && make(&m.a3, d.a[3], false, TX, TY + 1) &&
                  make(&m.a5, d.a[5], false, TX, TY) && make(&m.x, d.u, false, TX, TY);
  if (!ok) {
    delete p;
    return fail("cuTensorMapEncodeTiled failed");
  }
  const int tiles = ((g.nx + TX - 1) / TX) * ((g.ny + TY - 1) / TY);
  int zc = 32;
  if (const char* env = getenv("APHCG_ZC")) {
    zc = atoi(env);
  } else {
    int64_t best = -1;
    const int64_t slots = 2 * kNumSMs;
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
  zc = std::max(1, std::min(zc, g.nzl));
  p->zc = zc;
  // no L2 prefetch: the bulk copies are issued two plane steps ahead, which already covers the
  // DRAM latency; prefetching on top of them costs 25 % (measured: 2.00 vs 1.61 ms at 512^3)
  p->pd = 0;
  if (const char* ep = getenv("APHCG_PREFETCH2")) p->pd = atoi(ep);
  p->opts = 1;
  if (const char* eo = getenv("APHCG_PSTREAM")) p->opts = atoi(eo) ? 1 : 0;
  p->grid = dim3((g.nx + TX - 1) / TX, (g.ny + TY - 1) / TY, (g.nzl + zc - 1) / zc);
  using C = Cfg2<TX>;
  if (cudaFuncSetAttribute(k_dir_spmv_stream<TX, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                           C::kSmemBytes) != cudaSuccess ||
      cudaFuncSetAttribute(k_dir_spmv_stream<TX, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                           C::kSmemBytes) != cudaSuccess) {
    cudaGetLastError();
    delete p;
    return fail("cannot raise the dynamic shared memory limit");
  }
  return p;
}

void tma2_plan_destroy(Tma2Plan* p) { delete p; }
unsigned tma2_plan_blocks(const Tma2Plan* p) { return p->grid.x * p->grid.y * p->grid.z; }
void tma2_plan_describe(const Tma2Plan* p, char* buf, int buflen) {
  snprintf(buf, buflen,
           "tile=%dx%d planes_per_cta=%d stages=%d+%d all-operands-by-tma l2_prefetch=%d pstream=%d ctas=%u",
           p->tx, TY, p->zc, SF, SB, p->pd, p->opts & 1, tma2_plan_blocks(p));
}

void launch_dir_spmv_stream(const Tma2Plan* p, const Geom& g, const DevPtrs& d, bool single,
                            cudaStream_t s) {
  using C = Cfg2<64>;
  if (single) {
    k_dir_spmv_stream<64, true><<<p->grid, C::NT, C::kSmemBytes, s>>>(g, d, p->zc, p->pd, p->opts, p->maps);
  } else {
    k_dir_spmv_stream<64, false><<<p->grid, C::NT, C::kSmemBytes, s>>>(g, d, p->zc, p->pd, p->opts, p->maps);
  }
}

}  // namespace acg
