// TEST INFRASTRUCTURE -- not part of the product path.
//
// Driver around the reference's OWN classes (compiled from /root/reference/src,
// see Makefile): reads a 7-point system from raw files, solves it with whatever
// linear::ModuleLinear<M> the name selects (default "conjugate" =
// linear::SolverConjugate, src/linear/linear.ipp:18-150), writes the solution.
// Plays the role src/test/linear/main.cpp plays in the reference, but with
// file I/O that works without HDF5 (the reference's --system_in needs HDF5,
// src/test/linear/main.cpp:55-58).
//
// With `--plugin libX.so` the shared object is dlopen'ed first so that an
// out-of-tree module (our conjugate_cuda adapter) registers itself in the same
// ModuleLinear table (src/util/module.h:21-43) and can be selected by name:
// that is the drop-in path exercised end to end.
//
// Files (little-endian float64, global index, x fastest):
//   <sys>   : nx*ny*nz*8 doubles, AoS [c,x-,x+,y-,y+,z-,z+,const] per cell
//   <x0>    : nx*ny*nz doubles (optional; zero guess if absent)
//   <out>.x : nx*ny*nz doubles, <out>.info : text "iter residual time_s"

#include <dlfcn.h>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <mutex>
#include <sstream>
#include <string>
#include <vector>

#include "distr/distrbasic.h"
#include "linear/linear.h"
#include "util/distr.h"

// REFCG_DIM=2 (make dim2): the same driver on MeshCartesian<double,2>.  The file format stays
// the 3-D one (nz = 1, rows of 8 doubles); the z coefficients of the file are ignored.
#ifndef REFCG_DIM
#define REFCG_DIM 3
#endif
constexpr int kDim = REFCG_DIM;
using M = MeshCartesian<double, kDim>;
using Scal = typename M::Scal;
using MIdx = typename M::MIdx;
using Expr = typename M::Expr;

namespace {

struct Global {
  std::vector<double> sys; // N*8
  std::vector<double> x0; // N or empty
  std::vector<double> x; // N
  long nx = 0, ny = 0, nz = 0;
  int repeat = 1;
  double time_solve = 0;
  int iter = 0;
  double residual = 0;
  std::string out;
} g;

std::vector<double> ReadRaw(const std::string& path, size_t count) {
  std::vector<double> v(count);
  FILE* f = fopen(path.c_str(), "rb");
  if (!f) {
    std::cerr << "ref_cg: cannot open " << path << std::endl;
    std::exit(2);
  }
  const size_t got = fread(v.data(), sizeof(double), count, f);
  fclose(f);
  if (got != count) {
    std::cerr << "ref_cg: short read " << path << ": " << got << " of " << count
              << std::endl;
    std::exit(2);
  }
  return v;
}

void WriteRaw(const std::string& path, const std::vector<double>& v) {
  FILE* f = fopen(path.c_str(), "wb");
  fwrite(v.data(), sizeof(double), v.size(), f);
  fclose(f);
}

double Now() {
  using namespace std::chrono;
  return duration<double>(steady_clock::now().time_since_epoch()).count();
}

void Run(M& m, Vars& var) {
  auto sem = m.GetSem(__func__);
  struct {
    FieldCell<Expr> fc_system;
    FieldCell<Scal> fc_sol;
    std::unique_ptr<linear::Solver<M>> solver;
    typename linear::Solver<M>::Info info;
    double t0;
    int rep = 0;
  } * ctx(sem);
  auto& t = *ctx;
  auto gidx = [&](IdxCell c) -> size_t {
    const MIdx w = m.GetIndexCells().GetMIdx(c);
    const size_t k = kDim > 2 ? size_t(w[kDim - 1]) : 0;
    return (k * g.ny + w[1]) * g.nx + w[0];
  };
  if (sem("load")) {
    t.fc_system.Reinit(m, Expr(0));
    t.fc_system.SetName("pressure");
    for (auto c : m.Cells()) {
      const size_t i = gidx(c);
      Expr e;
      // [c, x-, x+, y-, y+, (z-, z+,) const]
      for (size_t k = 0; k < 2 * kDim + 1; ++k) {
        e[k] = g.sys[i * 8 + k];
      }
      e[2 * kDim + 1] = g.sys[i * 8 + 7];
      t.fc_system[c] = e;
    }
    t.fc_sol.Reinit(m, 0);
    if (!g.x0.empty()) {
      for (auto c : m.Cells()) {
        t.fc_sol[c] = g.x0[gidx(c)];
      }
      // callers hand over an initial guess with valid halos
      // (src/solver/proj.ipp:397 passes a field comm'd by the previous solve)
      m.Comm(&t.fc_sol, M::CommStencil::direct_one);
    }
    const auto name = var.String["linsolver_symm"];
    auto factory = linear::ModuleLinear<M>::GetInstance(name);
    fassert(factory, "Solver not found: " + name);
    t.solver = factory->Make(var, "symm", m);
    m.flags.linreport = var.Int["VERBOSE"];
  }
  sem.LoopBegin();
  if (sem("start")) {
    if (t.rep > 0 && !g.x0.empty()) {
      for (auto c : m.Cells()) {
        t.fc_sol[c] = g.x0[gidx(c)];
      }
      m.Comm(&t.fc_sol, M::CommStencil::direct_one);
    } else if (t.rep > 0) {
      t.fc_sol.Reinit(m, 0);
    }
  }
  if (sem("t0")) {
    t.t0 = Now();
  }
  if (sem.Nested("solve")) {
    t.info = t.solver->Solve(t.fc_system, &t.fc_sol, t.fc_sol, m);
  }
  if (sem("t1")) {
    if (m.IsRoot()) {
      const double dt = Now() - t.t0;
      if (t.rep == 0 || dt < g.time_solve) {
        g.time_solve = dt;
      }
      g.iter = t.info.iter;
      g.residual = t.info.residual;
    }
    ++t.rep;
    if (t.rep >= g.repeat) {
      sem.LoopBreak();
    }
  }
  sem.LoopEnd();
  if (sem("store")) {
    for (auto c : m.Cells()) {
      g.x[gidx(c)] = t.fc_sol[c];
    }
  }
  if (sem()) {
  }
}

const char* Arg(int argc, const char** argv, const char* key, const char* def) {
  for (int i = 1; i + 1 < argc; ++i) {
    if (!strcmp(argv[i], key)) return argv[i + 1];
  }
  return def;
}

} // namespace

int main(int argc, const char** argv) {
  if (argc < 2) {
    std::cerr
        << "usage: ref_cg --nx NX --ny NY --nz NZ --sys FILE [--x0 FILE] --out "
           "PREFIX\n"
           "  [--bsx B --bsy B --bsz B] [--tol T] [--maxiter K] [--miniter K] "
           "[--maxnorm 0|1]\n"
           "  [--px 0|1 --py 0|1 --pz 0|1] [--solver NAME] [--plugin LIB.so] "
           "[--extent L]\n"
           "  [--backend native|local] [--verbose 0|1] [--repeat R] [--extra "
           "'set ...']\n";
    return 1;
  }
  const char* plugin = Arg(argc, argv, "--plugin", "");
  if (plugin[0]) {
    if (!dlopen(plugin, RTLD_NOW | RTLD_GLOBAL)) {
      std::cerr << "ref_cg: dlopen failed: " << dlerror() << std::endl;
      return 2;
    }
  }
  g.nx = atol(Arg(argc, argv, "--nx", "32"));
  g.ny = atol(Arg(argc, argv, "--ny", Arg(argc, argv, "--nx", "32")));
  g.nz = atol(Arg(argc, argv, "--nz", Arg(argc, argv, "--nx", "32")));
  const long bsx = atol(Arg(argc, argv, "--bsx", "16"));
  const long bsy = atol(Arg(argc, argv, "--bsy", Arg(argc, argv, "--bsx", "16")));
  const long bsz = atol(Arg(argc, argv, "--bsz", Arg(argc, argv, "--bsx", "16")));
  if (kDim == 2 && g.nz != 1) {
    std::cerr << "ref_cg: the 2-D driver needs --nz 1" << std::endl;
    return 2;
  }
  if (g.nx % bsx || g.ny % bsy || (kDim > 2 && g.nz % bsz)) {
    std::cerr << "ref_cg: mesh not divisible by block" << std::endl;
    return 2;
  }
  const size_t n = size_t(g.nx) * g.ny * g.nz;
  g.sys = ReadRaw(Arg(argc, argv, "--sys", ""), n * 8);
  const char* x0 = Arg(argc, argv, "--x0", "");
  if (x0[0]) {
    g.x0 = ReadRaw(x0, n);
  }
  g.x.assign(n, 0.);
  g.out = Arg(argc, argv, "--out", "ref_cg_out");
  g.repeat = atoi(Arg(argc, argv, "--repeat", "1"));

  // the longest side has length `extent` (src/distr/distr.ipp:74-84)
  std::stringstream conf;
  conf << "set int bsx " << bsx << "\nset int bsy " << bsy << "\nset int bsz "
       << bsz << "\n";
  conf << "set int px 1\nset int py 1\nset int pz 1\n";
  conf << "set int bx " << g.nx / bsx << "\nset int by " << g.ny / bsy
       << "\nset int bz " << g.nz / bsz << "\n";
  conf << "set string linsolver_symm " << Arg(argc, argv, "--solver", "conjugate")
       << "\n";
  conf << "set double hypre_symm_tol " << Arg(argc, argv, "--tol", "1e-3") << "\n";
  conf << "set int hypre_symm_maxiter " << Arg(argc, argv, "--maxiter", "100")
       << "\n";
  conf << "set int hypre_symm_miniter " << Arg(argc, argv, "--miniter", "0")
       << "\n";
  conf << "set int linsolver_symm_maxnorm " << Arg(argc, argv, "--maxnorm", "0")
       << "\n";
  conf << "set int hypre_periodic_x " << Arg(argc, argv, "--px", "1") << "\n";
  conf << "set int hypre_periodic_y " << Arg(argc, argv, "--py", "1") << "\n";
  conf << "set int hypre_periodic_z " << Arg(argc, argv, "--pz", "1") << "\n";
  conf << "set string backend " << Arg(argc, argv, "--backend", "native") << "\n";
  conf << "set double extent " << Arg(argc, argv, "--extent", "1") << "\n";
  conf << "set int VERBOSE " << Arg(argc, argv, "--verbose", "0") << "\n";
  conf << Arg(argc, argv, "--extra", "") << "\n";

  MpiWrapper mpi(&argc, &argv);
  const int rc = RunMpiBasicString<M>(mpi, Run, conf.str());
  if (rc) return rc;

  WriteRaw(g.out + ".x", g.x);
  std::ofstream info(g.out + ".info");
  info.precision(17);
  info << g.iter << " " << g.residual << " " << g.time_solve << "\n";
  std::cout.precision(17);
  std::cout << "iter=" << g.iter << " residual=" << g.residual
            << " time=" << g.time_solve << std::endl;
  return 0;
}
