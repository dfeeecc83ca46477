// aphros-side adapter: registers libaphcg.so as the linear-solver modules
//   "conjugate_cuda"  (sibling of "conjugate",  src/linear/linear.ipp:239-253)
//   "jacobi_cuda"     (sibling of "jacobi",     src/linear/linear.ipp:255-265)
// behind the reference's own factory, so that `set string linsolver_symm
// conjugate_cuda` selects it with no change to aphros
// (ULinear<M>::MakeLinearSolver, src/util/linear.ipp:10-28).
//
// Compiled against the reference headers (-I/root/reference/src); loaded with
// LD_PRELOAD / dlopen / or linked in: the static registrar below adds the
// modules to ModuleLinear<M>'s table (src/util/module.h:21-43) at load time.
// All arithmetic happens in libaphcg.so through the C ABI (include/aphcg.h).
//
// Calling convention: Solve is a stage coroutine entered once
// per block per stage.  Blocks of one rank copy their inner cells into
// rank-wide pinned arrays owned by the lead block (the reference's precedent is
// LocalToShared/SharedToLocal, src/opencl/opencl.h:30-64); the lead block alone
// talks to the GPU, inside ONE stage; every block then copies its part of the
// solution back and requests the usual halo exchange of fc_sol
// (src/linear/linear.ipp:116-118).  Unlike conjugate_cl it needs no shared-mesh
// Comm, so it does not depend on the backend: tested under native and local; cubismnc
// (not buildable here) provides the same base-class services it uses
// (BcastFromLead, block-level Comm; src/distr/distr.ipp:171-282).
//
// Dimensions: the reference registers `conjugate` for every enabled
// MeshCartesian<double,dim> (src/linear/linear.cpp:16-23).  The C ABI is a 3-D 7-point
// solver; 1-D and 2-D meshes map onto it with ny and/or nz = 1, their rows
// [c, x-, x+, (y-, y+,) const] widened to the 8-double format with zero coefficients in the
// missing directions (a missing direction is neither periodic nor coupled).  4-D meshes are
// not a 7-point problem and are not registered.
#include <cstdio>
#include <cstdint>
#include <cstring>
#include <iostream>
#include <memory>
#include <string>
#include <vector>

#include "linear/linear.h"
#include "util/macros.h"

#include "aphcg.h"
#include "linear_projection.h"

#if defined(__SSE2__)
#include <emmintrin.h>
#endif

namespace linear {

// Copy into the rank-wide staging arrays with non-temporal stores: the destination is written
// once and next read by the GPU's copy engine, so it should not displace the block's working
// set from the caches (nor be read for ownership first).  dst must be 16-byte aligned.
inline void StreamCopy(double* dst, const double* src, size_t n) {
#if defined(__SSE2__)
  if ((reinterpret_cast<uintptr_t>(dst) & 15) == 0) {
    size_t i = 0;
    for (; i + 2 <= n; i += 2) _mm_stream_pd(dst + i, _mm_loadu_pd(src + i));
    for (; i < n; ++i) dst[i] = src[i];
    return;
  }
#endif
  std::memcpy(dst, src, n * sizeof(double));
}

template <class M>
class SolverCuda : public Solver<M>, public ProjectionSolver<M> {
 public:
  using Base = Solver<M>;
  using Conf = typename Base::Conf;
  using Info = typename Base::Info;
  using Scal = typename M::Scal;
  using Expr = typename M::Expr;
  using MIdx = typename M::MIdx;
  enum class Method { conjugate, jacobi };
  static constexpr int dim = int(M::dim);

  SolverCuda(const Conf& conf, Method method, bool maxnorm, int device, int ndevices,
             int slabs_per_device, unsigned flags, const M& m, std::string dump = "",
             int dump_index = 0)
      : Base(conf), method_(method), dump_(dump), dump_index_(dump_index) {
    static_assert(dim >= 1 && dim <= 3, "conjugate_cuda: 1-D, 2-D and 3-D meshes");
    static_assert(
        sizeof(Expr) == (2 * dim + 2) * sizeof(double), "row must be 2*dim+2 doubles");
    if (m.IsLead()) {
      // one device object per rank, owned by the lead block
      // (cf. src/linear/conjugate_cl.ipp:29-35)
      shared_obj_ = std::make_unique<Shared>();
      auto& s = *shared_obj_;
      const auto& ms = m.GetShared();
      const MIdx size = ms.GetInBlockCells().GetSize();
      s.origin = ms.GetInBlockCells().GetBegin();
      for (int i = 0; i < dim; ++i) s.size[i] = size[i];
      aphcg_desc d;
      std::memset(&d, 0, sizeof(d));
      d.nx = s.size[0];
      d.ny = s.size[1];
      d.nz = s.size[2];
      fassert(
          ms.GetGlobalSize() == size,
          "conjugate_cuda: one rank must own the whole domain (run aphros with "
          "px=py=pz=1 and set cuda_devices to use several GPUs from that rank)");
      for (int i = 0; i < dim; ++i) d.periodic[i] = m.flags.is_periodic[i] ? 1 : 0;
      d.cell_volume = m.GetCellSize().prod();
      d.rank = 0;
      d.nranks = 1;
      d.z0 = 0;
      d.nz_local = s.size[2];
      d.flags = flags | (maxnorm ? APHCG_MAXNORM : 0);
      // the rank-wide index space is cut into `ndevices` z-slabs, one per GPU
      // device..device+ndevices-1, all driven from this (lead) block
      // (cuda_slabs_per_device > 1 puts several consecutive slabs on each GPU)
      fassert(
          ndevices >= 1 && slabs_per_device >= 1 && ndevices * slabs_per_device <= 16,
          "conjugate_cuda: cuda_devices x cuda_slabs_per_device must be 1..16");
      std::vector<int32_t> devices;
      for (int i = 0; i < ndevices; ++i)
        for (int k = 0; k < slabs_per_device; ++k) devices.push_back(device + i);
      Check(aphcg_group_create(&s.group, &d, devices.data(), (int32_t)devices.size()));
      const uint64_t n = uint64_t(s.size[0]) * s.size[1] * s.size[2];
      Check(aphcg_host_alloc(reinterpret_cast<void**>(&s.rows), n * 8 * sizeof(double)));
      Check(aphcg_host_alloc(reinterpret_cast<void**>(&s.x), n * sizeof(double)));
      shared_ = &s;
    }
  }
  ~SolverCuda() override {
    if (shared_obj_) {
      auto& s = *shared_obj_;
      if (s.rows) aphcg_host_free(s.rows);
      if (s.x) aphcg_host_free(s.x);
      for (double* p : {s.rho, s.vx, s.vy, s.vz, s.src}) {
        if (p) aphcg_host_free(p);
      }
      if (s.group) aphcg_group_destroy(s.group);
    }
  }

  Info Solve(
      const FieldCell<Expr>& fc_system, const FieldCell<Scal>* fc_init,
      FieldCell<Scal>& fc_sol, M& m) override {
    auto sem = m.GetSem(__func__);
    struct {
      Info info;
    } * ctx(sem);
    auto& t = *ctx;
    if (sem("bcast")) {
      m.BcastFromLead(&shared_);
    }
    if (sem("gather")) {
      auto& s = *shared_;
      // rows and guess of this block -> rank-wide pinned arrays (disjoint ranges,
      // safe when blocks run concurrently under OpenMP, src/distr/distr.ipp:90-97).
      // Cells of one x-row are contiguous both in the block's fields (x fastest,
      // src/geom/block.h:149-158) and in the rank-wide arrays: whole rows are copied.
      ForEachRow(m, [&](IdxCell c0, size_t i0, size_t n) {
        const Expr* src = &fc_system[c0];
        if (dim == 3) {
          StreamCopy(s.rows + 8 * i0, &(*src)[0], n * 8);
        } else {  // [c, x-, x+, (y-, y+,) const] -> [c, x-, x+, y-, y+, z-, z+, const]
          for (size_t q = 0; q < n; ++q) {
            const Expr& e = src[q];
            double* row = s.rows + 8 * (i0 + q);
            for (int k = 0; k < 7; ++k) row[k] = k < 2 * dim + 1 ? e[k] : 0.;
            row[7] = e[2 * dim + 1];
          }
        }
        if (fc_init) {
          std::memcpy(s.x + i0, &(*fc_init)[c0], n * sizeof(double));
        } else {
          std::memset(s.x + i0, 0, n * sizeof(double));
        }
      });
#if defined(__SSE2__)
      _mm_sfence();  // the streamed rows are globally visible before the next stage reads them
#endif
    }
    if (sem("solve") && m.IsLead()) {
      auto& s = *shared_;
      aphcg_conf conf;
      conf.tol = this->conf.tol;
      conf.miniter = this->conf.miniter;
      conf.maxiter = this->conf.maxiter;
      aphcg_info info;
      if (!dump_.empty() && s.ncalls == dump_index_) {
        DumpSystem(s, fc_system.GetName(), fc_init != nullptr, m);
      }
      ++s.ncalls;
      if (method_ == Method::conjugate) {
        // x doubles as guess and solution (fc_init may alias fc_sol, linear.h:40)
        Check(aphcg_group_solve(
            s.group, s.rows, nullptr, fc_init ? s.x : nullptr, nullptr, s.x, nullptr, &conf,
            &info));
      } else {
        Check(aphcg_group_upload_system(s.group, s.rows, nullptr));
        Check(aphcg_group_upload_guess(s.group, fc_init ? s.x : nullptr, nullptr));
        Check(aphcg_group_run_jacobi(s.group, &conf, &info));
        Check(aphcg_group_download_solution(s.group, s.x, nullptr));
      }
      s.info.residual = info.residual;
      s.info.iter = info.iter;
    }
    if (sem("scatter")) {
      auto& s = *shared_;
      if (!fc_sol.size()) {
        fc_sol.Reinit(m);  // callers may pass an empty field (cf. opencl.h:36-40)
      }
      ForEachRow(m, [&](IdxCell c0, size_t i0, size_t n) {
        std::memcpy(&fc_sol[c0], s.x + i0, n * sizeof(double));
      });
      t.info = s.info;
      m.Comm(&fc_sol, M::CommStencil::direct_one);
      if (m.flags.linreport && m.IsRoot()) {
        std::cerr << std::scientific;
        std::cerr << std::string("linear(") +
                         (method_ == Method::conjugate ? "conjugate_cuda" : "jacobi_cuda") +
                         ") '" + fc_system.GetName() + "':"
                  << " res=" << t.info.residual << " iter=" << t.info.iter << std::endl;
      }
    }
    if (sem()) {
      // the Comm of fc_sol completes before the context is destroyed (linear.ipp:126-127)
    }
    return t.info;
  }

  // linear_projection.h: the pressure system of the projection step assembled on the GPU from
  // the density and the face fluxes (Proj::GetFlux + GetFluxSum, src/solver/proj.ipp:343-383)
  Info SolveProjection(
      const FieldCell<Scal>& fc_dens, const FieldFace<Scal>& ff_flux,
      const FieldCell<Scal>* fc_source, Scal dt, const FieldCell<Scal>* fc_init,
      FieldCell<Scal>& fc_sol, M& m) override {
    auto sem = m.GetSem(__func__);
    struct {
      Info info;
    } * ctx(sem);
    auto& t = *ctx;
    fassert(dim == 3, "conjugate_cuda: SolveProjection needs a 3-D mesh");
    fassert(method_ == Method::conjugate, "jacobi_cuda has no SolveProjection");
    if (sem("alloc") && m.IsLead()) {
      auto& s = *shared_obj_;
      if (!s.rho) {
        const uint64_t nx = s.size[0], ny = s.size[1], nz = s.size[2];
        auto alloc = [&](double** p, uint64_t n) {
          Check(aphcg_host_alloc(reinterpret_cast<void**>(p), n * sizeof(double)));
        };
        alloc(&s.rho, nx * ny * (nz + 2));
        alloc(&s.vx, (nx + 1) * ny * nz);
        alloc(&s.vy, nx * (ny + 1) * nz);
        alloc(&s.vz, nx * ny * (nz + 1));
        alloc(&s.src, nx * ny * nz);
      }
    }
    if (sem("bcast")) {
      m.BcastFromLead(&shared_);
    }
    if (sem("gather")) {
      GatherProjection(fc_dens, ff_flux, fc_source, fc_init, m);
    }
    if (sem("solve") && m.IsLead()) {
      auto& s = *shared_;
      aphcg_conf conf;
      conf.tol = this->conf.tol;
      conf.miniter = this->conf.miniter;
      conf.maxiter = this->conf.maxiter;
      aphcg_info info;
      ++s.ncalls;
      Check(aphcg_group_assemble_projection(
          s.group, s.rho, s.vx, s.vy, s.vz, fc_source ? s.src : nullptr, dt,
          m.GetCellSize()[0]));
      Check(aphcg_group_upload_guess(s.group, fc_init ? s.x : nullptr, nullptr));
      Check(aphcg_group_run(s.group, &conf, &info));
      Check(aphcg_group_download_solution(s.group, s.x, nullptr));
      s.info.residual = info.residual;
      s.info.iter = info.iter;
    }
    if (sem("scatter")) {
      auto& s = *shared_;
      if (!fc_sol.size()) {
        fc_sol.Reinit(m);
      }
      ForEachRow(m, [&](IdxCell c0, size_t i0, size_t n) {
        std::memcpy(&fc_sol[c0], s.x + i0, n * sizeof(double));
      });
      t.info = s.info;
      m.Comm(&fc_sol, M::CommStencil::direct_one);
      if (m.flags.linreport && m.IsRoot()) {
        std::cerr << std::scientific;
        std::cerr << "linear(conjugate_cuda) 'pressure' (device assembly):"
                  << " res=" << t.info.residual << " iter=" << t.info.iter << std::endl;
      }
    }
    if (sem()) {
    }
    return t.info;
  }

 private:
  struct Shared {
    aphcg_group_t* group = nullptr;  // one slab per GPU; a group of one is the plain solver
    double* rows = nullptr;  // pinned, rank-wide, [nz][ny][nx][8]
    double* x = nullptr;     // pinned, rank-wide, guess in / solution out
    // SolveProjection only (allocated on first use): rank-wide density with one ghost plane
    // below and above, face fluxes, source
    double *rho = nullptr, *vx = nullptr, *vy = nullptr, *vz = nullptr, *src = nullptr;
    MIdx origin;
    size_t size[3] = {1, 1, 1};  // rank-wide inner cells; 1 in the directions a mesh lacks
    Info info;
    int ncalls = 0;  // Solve calls so far (selects the call to dump)
    size_t Index(MIdx w) const {
      const MIdx l = w - origin;
      size_t i = 0;
      for (int d = dim - 1; d >= 0; --d) i = i * size[d] + size_t(l[d]);
      return i;
    }
  };
  // This block's part of the rank-wide inputs of aphcg_group_assemble_projection.  Every
  // location is written by exactly one block: a cell's density and LOWER faces by the block
  // that owns the cell, the domain's last upper faces by the block that touches that boundary.
  void GatherProjection(
      const FieldCell<Scal>& fc_dens, const FieldFace<Scal>& ff_flux,
      const FieldCell<Scal>* fc_source, const FieldCell<Scal>* fc_init, const M& m) const {
    auto& s = *shared_;
    const size_t nx = s.size[0], ny = s.size[1], nz = s.size[2];
    const bool perz = m.flags.is_periodic[dim - 1];
    const auto& ic = m.GetIndexCells();
    for (auto c : m.Cells()) {
      const MIdx w = ic.GetMIdx(c) - s.origin;
      const size_t i = w[0], j = w[1], k = dim > 2 ? w[dim - 1] : 0;
      const size_t cell = (k * ny + j) * nx + i;
      const Scal rho = fc_dens[c];
      s.rho[cell + nx * ny] = rho;  // plane k+1 of the ghost-extended array
      if (k == 0) s.rho[(perz ? nz + 1 : 0) * nx * ny + j * nx + i] = rho;
      if (k == nz - 1) s.rho[(perz ? 0 : nz + 1) * nx * ny + j * nx + i] = rho;
      s.src[cell] = fc_source ? (*fc_source)[c] : 0.;
      s.x[cell] = fc_init ? (*fc_init)[c] : 0.;
      s.vx[(k * ny + j) * (nx + 1) + i] = ff_flux[m.GetFace(c, IdxNci(0))];
      if (i == nx - 1) s.vx[(k * ny + j) * (nx + 1) + nx] = ff_flux[m.GetFace(c, IdxNci(1))];
      s.vy[(k * (ny + 1) + j) * nx + i] = ff_flux[m.GetFace(c, IdxNci(2))];
      if (j == ny - 1) s.vy[(k * (ny + 1) + ny) * nx + i] = ff_flux[m.GetFace(c, IdxNci(3))];
      s.vz[cell] = ff_flux[m.GetFace(c, IdxNci(4))];
      if (k == nz - 1) s.vz[cell + nx * ny] = ff_flux[m.GetFace(c, IdxNci(5))];
    }
  }
  // f(first cell of the row, its index in the rank-wide arrays, cells in the row) for every
  // x-row of the block's inner cells
  template <class F>
  void ForEachRow(const M& m, F f) const {
    const auto& bc = m.GetInBlockCells();
    const MIdx b0 = bc.GetBegin(), bs = bc.GetSize();
    const auto& ic = m.GetIndexCells();
    size_t nrows = 1;
    for (int d = 1; d < dim; ++d) nrows *= size_t(bs[d]);
    for (size_t r = 0; r < nrows; ++r) {
      MIdx w = b0;
      size_t t = r;
      for (int d = 1; d < dim; ++d) {
        w[d] = b0[d] + int(t % size_t(bs[d]));
        t /= size_t(bs[d]);
      }
      f(ic.GetIdx(w), shared_->Index(w), size_t(bs[0]));
    }
  }
  // Capture of a live system for replay outside aphros (the reference's
  // t.linear --system_out needs HDF5, src/test/linear/main.cpp:182-189): the rank-wide rows and
  // guess exactly as they go to the C ABI, raw little-endian float64, plus a text header.
  //   <prefix>.sys  nz*ny*nx*8   <prefix>.x0  nz*ny*nx (if a guess was given)   <prefix>.txt
  // Replay: python -m aphros_b200.tlinear --replay <prefix> [--solver ...]
  void DumpSystem(const Shared& s, const std::string& name, bool have_guess, const M& m) const {
    const size_t n = s.size[0] * s.size[1] * s.size[2];
    auto write = [&](const std::string& path, const double* p, size_t count) {
      FILE* f = std::fopen(path.c_str(), "wb");
      fassert(f, "conjugate_cuda: cannot write " + path);
      const size_t put = std::fwrite(p, sizeof(double), count, f);
      std::fclose(f);
      fassert(put == count, "conjugate_cuda: short write to " + path);
    };
    write(dump_ + ".sys", s.rows, n * 8);
    if (have_guess) write(dump_ + ".x0", s.x, n);
    FILE* f = std::fopen((dump_ + ".txt").c_str(), "w");
    fassert(f, "conjugate_cuda: cannot write " + dump_ + ".txt");
    std::fprintf(f, "nx %zu\nny %zu\nnz %zu\n", s.size[0], s.size[1], s.size[2]);
    std::fprintf(f, "periodic");
    for (int i = 0; i < 3; ++i) std::fprintf(f, " %d", i < dim && m.flags.is_periodic[i] ? 1 : 0);
    std::fprintf(f, "\ncell_volume %.17g\n", double(m.GetCellSize().prod()));
    std::fprintf(f, "tol %.17g\nminiter %d\nmaxiter %d\n", double(this->conf.tol),
                 this->conf.miniter, this->conf.maxiter);
    std::fprintf(f, "guess %d\ncall %d\nname %s\n", have_guess ? 1 : 0, s.ncalls, name.c_str());
    std::fclose(f);
  }
  // CUDA / NCCL failures surface like any other aphros error
  // (fassert -> aphros_SetError + throw, src/util/logger.h:44-60)
  static void Check(int rc) {
    fassert(rc == 0, std::string("conjugate_cuda: ") + aphcg_last_error());
  }

  Method method_;
  std::string dump_;  // linsolver_<prefix>_cuda_dump: file prefix, empty = off
  int dump_index_ = 0;
  std::unique_ptr<Shared> shared_obj_;
  Shared* shared_ = nullptr;
};

template <class M>
class ModuleLinearConjugateCuda : public ModuleLinear<M> {
 public:
  ModuleLinearConjugateCuda() : ModuleLinear<M>("conjugate_cuda") {}
  std::unique_ptr<Solver<M>> Make(const Vars& var, std::string prefix, const M& m) override {
    auto key = [prefix](std::string name) { return "linsolver_" + prefix + "_" + name; };
    // new keys are read with defaults: Vars::operator[] throws when a key is missing
    const bool maxnorm = var.Int(key("maxnorm"), 0);
    unsigned flags = 0;
    if (!var.Int(key("cuda_graph"), 1)) flags |= APHCG_NO_GRAPH;
    if (!var.Int(key("cuda_tma"), 1)) flags |= APHCG_NO_TMA;
    if (!var.Int(key("cuda_persistent"), 1)) flags |= APHCG_NO_PERSISTENT;
    if (!var.Int(key("cuda_stream"), 1)) flags |= APHCG_NO_STREAM;
    // opt-in diagonal preconditioner (NOT SolverConjugate's recurrence; default off)
    if (var.Int(key("jacobi"), 0)) flags |= APHCG_JACOBI_PRECOND;
    return std::make_unique<SolverCuda<M>>(
        this->GetConf(var, prefix), SolverCuda<M>::Method::conjugate, maxnorm,
        var.Int("cuda_device", 0), var.Int("cuda_devices", 1),
        var.Int("cuda_slabs_per_device", 1), flags, m, var.String(key("cuda_dump"), ""),
        var.Int(key("cuda_dump_index"), 0));
  }
};

template <class M>
class ModuleLinearJacobiCuda : public ModuleLinear<M> {
 public:
  ModuleLinearJacobiCuda() : ModuleLinear<M>("jacobi_cuda") {}
  std::unique_ptr<Solver<M>> Make(const Vars& var, std::string prefix, const M& m) override {
    return std::make_unique<SolverCuda<M>>(
        this->GetConf(var, prefix), SolverCuda<M>::Method::jacobi, false,
        var.Int("cuda_device", 0), var.Int("cuda_devices", 1),
        var.Int("cuda_slabs_per_device", 1), 0u, m);
  }
};

// one registration per enabled dimension, like src/linear/linear.cpp:16-23 (4-D excluded)
#define X(dim)                                                               \
  RegisterModule<ModuleLinearConjugateCuda<MeshCartesian<double, dim>>>(),   \
      RegisterModule<ModuleLinearJacobiCuda<MeshCartesian<double, dim>>>(),
bool kReg_conjugate_cuda[] = {MULTIDIMX1 MULTIDIMX2 MULTIDIMX3};
#undef X

} // namespace linear
