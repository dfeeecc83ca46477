// TEST INFRASTRUCTURE -- not part of the product path.
//
// Assembles the pressure system of the projection step with the reference's OWN
// functions (compiled from /root/reference/src, see Makefile) from a cell density and
// face volume fluxes given as raw files, and dumps the rows.  It is what pins
// aphcg_assemble_projection (device-side assembly) to the reference.
//
// The sequence is Proj<EB>::Imp::GetFlux followed by GetFluxSum
// (src/solver/proj.ipp:343-383) on a mesh without embedded boundaries, every
// non-periodic domain face being a wall (a non-pressure boundary condition:
// `ffe[cf] = ExprFace(0)`, proj.ipp:352-354).  Those two are private members of the
// fluid solver and cannot be called from outside, so the few statements that glue
// the reference's building blocks together are restated here; every building block
// itself is the reference's:
//   UEmbed<M>::InterpolateHarmonic  (src/solver/approx_eb.h:351-363)   face density
//   UEmbed<M>::GradientImplicit     (src/solver/approx_eb.ipp:1434-1468) [-1/h, 1/h, 0]
//   M::GetArea, GetOutwardFactor, GetVolume, LoopNci, GetFace
//   M::AppendExpr                   (src/geom/mesh.h:575-579)          face -> cell row
// (the same way src/test/linear/main.cpp:60-78 builds its test system).
//
// Files (little-endian float64, global index, x fastest):
//   --rho  nx*ny*nz          cell density
//   --vx   (nx+1)*ny*nz      volume flux through x faces (face i = lower face of cell i)
//   --vy   nx*(ny+1)*nz      y faces;   --vz  nx*ny*(nz+1)  z faces
//   --src  nx*ny*nz          volume source (optional)
//   --out  nx*ny*nz*8        rows [c, x-, x+, y-, y+, z-, z+, const]
//
// With `--plugin LIB.so --solver NAME --xout FILE` the module NAME (loaded from LIB.so like
// ref_cg does) is additionally asked for the linear::ProjectionSolver capability
// (aphros_b200/plugin/linear_projection.h) and solves the SAME projection from the same
// density / flux / source fields, zero guess; its solution goes to FILE (+ FILE.info with
// "iter residual").  That is the caller-side code a maintainer would put into
// Proj::Project, exercised through the reference's own mesh and coroutine machinery.

#include <dlfcn.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <memory>
#include <sstream>
#include <string>
#include <vector>

#include "distr/distrbasic.h"
#include "solver/approx_eb.h"
#include "solver/embed.h"
#include "util/distr.h"

#include "../../aphros_b200/plugin/linear_projection.h"

using M = MeshCartesian<double, 3>;
using Scal = typename M::Scal;
using MIdx = typename M::MIdx;
using Expr = typename M::Expr;
using ExprFace = typename M::ExprFace;
using UEB = UEmbed<M>;

namespace {

struct Global {
  std::vector<double> rho, v[3], src, rows, x;
  long n[3] = {0, 0, 0};
  double dt = 1;
  std::string solver;  // module asked for the ProjectionSolver capability ("" = none)
  int iter = 0;
  double residual = 0;
} g;

std::vector<double> ReadRaw(const std::string& path, size_t count) {
  std::vector<double> v(count);
  FILE* f = fopen(path.c_str(), "rb");
  if (!f || fread(v.data(), sizeof(double), count, f) != count) {
    std::cerr << "ref_assemble: cannot read " << count << " doubles from " << path
              << std::endl;
    std::exit(2);
  }
  fclose(f);
  return v;
}

void Run(M& m, Vars& var) {
  auto sem = m.GetSem(__func__);
  struct {
    FieldCell<Scal> fcr;
    FieldCell<Scal> fcsv;
    FieldFace<Scal> ffv;
    FieldCell<Scal> fcp;
    std::unique_ptr<linear::Solver<M>> solver;
    linear::ProjectionSolver<M>* proj = nullptr;
    typename linear::Solver<M>::Info info;
  } * ctx(sem);
  auto& t = *ctx;
  auto cidx = [&](MIdx w) -> size_t {
    return (size_t(w[2]) * g.n[1] + w[1]) * g.n[0] + w[0];
  };
  if (sem("load")) {
    t.fcr.Reinit(m, 1);
    t.fcsv.Reinit(m, 0);
    for (auto c : m.Cells()) {
      const MIdx w = m.GetIndexCells().GetMIdx(c);
      t.fcr[c] = g.rho[cidx(w)];
      if (!g.src.empty()) t.fcsv[c] = g.src[cidx(w)];
    }
    m.Comm(&t.fcr);  // halo cells: neighbour blocks / periodic images
  }
  if (sem("assemble")) {
    const MIdx gs = m.GetGlobalSize();
    // volume fluxes
    auto& ffv = t.ffv;
    ffv.Reinit(m, 0);
    for (auto f : m.Faces()) {
      const MIdx w = m.GetIndexFaces().GetMIdx(f);
      const size_t d = m.GetIndexFaces().GetDir(f).raw();
      long sz[3] = {g.n[0], g.n[1], g.n[2]};
      sz[d] += 1;
      ffv[f] = g.v[d][(size_t(w[2]) * sz[1] + w[1]) * sz[0] + w[0]];
    }
    const FieldFace<Scal> ffdens =
        UEB::InterpolateHarmonic(t.fcr, MapEmbed<BCond<Scal>>(), m);
    // GetFlux (proj.ipp:343-360)
    FieldFace<ExprFace> ffe = UEB::GradientImplicit(MapEmbed<BCond<Scal>>(), m);
    for (auto f : m.Faces()) {
      const MIdx w = m.GetIndexFaces().GetMIdx(f);
      const size_t d = m.GetIndexFaces().GetDir(f).raw();
      if (!m.flags.is_periodic[d] && (w[d] == 0 || w[d] == gs[d])) {
        ffe[f] = ExprFace(0);  // wall: no pressure-gradient term
      }
    }
    for (auto f : m.Faces()) {
      ffe[f] *= -m.GetArea(f) / ffdens[f] * g.dt;
      ffe[f][2] += ffv[f];
    }
    // GetFluxSum (proj.ipp:367-383)
    for (auto c : m.Cells()) {
      Expr sum(0);
      m.LoopNci(c, [&](auto q) {
        const auto cf = m.GetFace(c, q);
        const ExprFace v = ffe[cf] * m.GetOutwardFactor(c, q);
        m.AppendExpr(sum, v, q);
      });
      sum.back() -= t.fcsv[c] * m.GetVolume(c);
      const size_t i = cidx(m.GetIndexCells().GetMIdx(c));
      for (size_t k = 0; k < 8; ++k) g.rows[i * 8 + k] = sum[k];
    }
  }
  if (!g.solver.empty()) {
    if (sem("make")) {
      auto factory = linear::ModuleLinear<M>::GetInstance(g.solver);
      fassert(factory, "Solver not found: " + g.solver);
      t.solver = factory->Make(var, "symm", m);
      t.proj = dynamic_cast<linear::ProjectionSolver<M>*>(t.solver.get());
      fassert(t.proj, "module " + g.solver + " does not implement linear::ProjectionSolver");
      t.fcp.Reinit(m, 0);
    }
    if (sem.Nested("solve")) {
      t.info = t.proj->SolveProjection(
          t.fcr, t.ffv, g.src.empty() ? nullptr : &t.fcsv, g.dt, nullptr, t.fcp, m);
    }
    if (sem("store")) {
      for (auto c : m.Cells()) {
        g.x[cidx(m.GetIndexCells().GetMIdx(c))] = t.fcp[c];
      }
      if (m.IsRoot()) {
        g.iter = t.info.iter;
        g.residual = t.info.residual;
      }
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
    std::cerr << "usage: ref_assemble --nx NX --ny NY --nz NZ --rho F --vx F --vy F "
                 "--vz F [--src F] --dt DT --out F [--bsx B --bsy B --bsz B] "
                 "[--px 0|1 --py 0|1 --pz 0|1]\n";
    return 1;
  }
  g.n[0] = atol(Arg(argc, argv, "--nx", "32"));
  g.n[1] = atol(Arg(argc, argv, "--ny", "32"));
  g.n[2] = atol(Arg(argc, argv, "--nz", "32"));
  const long bs[3] = {
      atol(Arg(argc, argv, "--bsx", Arg(argc, argv, "--nx", "32"))),
      atol(Arg(argc, argv, "--bsy", Arg(argc, argv, "--ny", "32"))),
      atol(Arg(argc, argv, "--bsz", Arg(argc, argv, "--nz", "32")))};
  for (int d = 0; d < 3; ++d) {
    if (g.n[d] % bs[d]) {
      std::cerr << "ref_assemble: mesh not divisible by block" << std::endl;
      return 2;
    }
  }
  const size_t n = size_t(g.n[0]) * g.n[1] * g.n[2];
  g.rho = ReadRaw(Arg(argc, argv, "--rho", ""), n);
  g.v[0] = ReadRaw(Arg(argc, argv, "--vx", ""), n / g.n[0] * (g.n[0] + 1));
  g.v[1] = ReadRaw(Arg(argc, argv, "--vy", ""), n / g.n[1] * (g.n[1] + 1));
  g.v[2] = ReadRaw(Arg(argc, argv, "--vz", ""), n / g.n[2] * (g.n[2] + 1));
  const char* src = Arg(argc, argv, "--src", "");
  if (src[0]) g.src = ReadRaw(src, n);
  g.dt = atof(Arg(argc, argv, "--dt", "1"));
  g.rows.assign(n * 8, 0.);
  const char* plugin = Arg(argc, argv, "--plugin", "");
  if (plugin[0]) {
    if (!dlopen(plugin, RTLD_NOW | RTLD_GLOBAL)) {
      std::cerr << "ref_assemble: dlopen failed: " << dlerror() << std::endl;
      return 2;
    }
    g.solver = Arg(argc, argv, "--solver", "conjugate_cuda");
    g.x.assign(n, 0.);
  }

  std::stringstream conf;
  conf << "set int bsx " << bs[0] << "\nset int bsy " << bs[1] << "\nset int bsz "
       << bs[2] << "\n";
  conf << "set int px 1\nset int py 1\nset int pz 1\n";
  conf << "set int bx " << g.n[0] / bs[0] << "\nset int by " << g.n[1] / bs[1]
       << "\nset int bz " << g.n[2] / bs[2] << "\n";
  conf << "set int hypre_periodic_x " << Arg(argc, argv, "--px", "0") << "\n";
  conf << "set int hypre_periodic_y " << Arg(argc, argv, "--py", "0") << "\n";
  conf << "set int hypre_periodic_z " << Arg(argc, argv, "--pz", "0") << "\n";
  conf << "set string backend native\nset double extent 1\nset int VERBOSE 0\n";
  conf << "set double hypre_symm_tol " << Arg(argc, argv, "--tol", "1e-8") << "\n";
  conf << "set int hypre_symm_maxiter " << Arg(argc, argv, "--maxiter", "1000") << "\n";
  conf << Arg(argc, argv, "--extra", "") << "\n";

  MpiWrapper mpi(&argc, &argv);
  const int rc = RunMpiBasicString<M>(mpi, Run, conf.str());
  if (rc) return rc;
  FILE* f = fopen(Arg(argc, argv, "--out", "rows.f64"), "wb");
  fwrite(g.rows.data(), sizeof(double), g.rows.size(), f);
  fclose(f);
  if (!g.solver.empty()) {
    const std::string xout = Arg(argc, argv, "--xout", "x.f64");
    f = fopen(xout.c_str(), "wb");
    fwrite(g.x.data(), sizeof(double), g.x.size(), f);
    fclose(f);
    std::ofstream info(xout + ".info");
    info.precision(17);
    info << g.iter << " " << g.residual << "\n";
  }
  return 0;
}
