// aphcg_group_*: one process, several z-slabs (include/aphcg.h, "in-process slab group").
//
// The reference runs one solver object per rank and reaches other ranks through MPI
// (src/distr/distr.ipp:143-169, src/distr/native.ipp:147-365).  A process that owns
// several GPUs -- aphros started without MPI on a multi-GPU node, the case
// SURVEY.md 8(e) names first -- drives them through this file instead: the rank-wide
// arrays are cut into contiguous z-slabs, every slab is an ordinary aphcg handle on
// its own device, and one host thread per slab issues exactly the calls a rank of the
// process-per-GPU mode would.  Inside the loop nothing changes (halo planes and the
// scalars travel through peer memory, written by the kernels); outside it the two
// collective steps of a solve use a thread barrier (cg_group.h) instead of NCCL.
#include <cuda.h>
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <exception>
#include <functional>
#include <string>
#include <thread>
#include <vector>

#include "../../include/aphcg.h"
#include "cg_group.h"

using namespace acg;

struct aphcg_group {
  int n = 0;
  aphcg_desc desc{};  // global geometry
  std::vector<aphcg_t*> h;
  std::vector<int64_t> z0, nzl;
  GroupSync* gs = nullptr;
  bool broken = false;
};

namespace {

int GroupFail(int code, const std::string& msg) {
  SetLastError(msg.c_str());
  return code;
}

// fn(slab) on one host thread per slab; returns the first failure and leaves its message
// in the caller's aphcg_last_error().  A failing slab aborts the group's barrier so that
// its peers cannot wait for it forever; the group is unusable afterwards.
int RunAll(aphcg_group* g, const std::function<int(int)>& fn) {
  if (g->broken)
    return GroupFail(APHCG_ERR_STATE, "group is unusable after an earlier failure; destroy it");
  if (g->n == 1) return fn(0);
  std::vector<int> rc(g->n, 0);
  std::vector<std::string> err(g->n);
  auto body = [&](int q) {
    rc[q] = fn(q);
    if (rc[q] != 0) {
      err[q] = aphcg_last_error();
      g->gs->Abort();
    }
  };
  std::vector<std::thread> th;
  th.reserve(g->n - 1);
  try {
    for (int q = 1; q < g->n; ++q) th.emplace_back(body, q);
  } catch (const std::exception& e) {
    // a slab thread could not be started: release the ones that run (they would wait for it
    // at the first barrier) and give up
    g->gs->Abort();
    for (auto& t : th) t.join();
    g->broken = true;
    return GroupFail(APHCG_ERR_STATE, std::string("cannot start a slab thread: ") + e.what());
  }
  body(0);
  for (auto& t : th) t.join();
  // report the slab that failed on its own, not a peer that only saw the abort
  int first = -1;
  for (int q = 0; q < g->n; ++q) {
    if (rc[q] == 0) continue;
    if (first < 0 || (err[first].find("another slab") != std::string::npos &&
                      err[q].find("another slab") == std::string::npos))
      first = q;
  }
  if (first < 0) return 0;
  g->broken = true;
  char head[32];
  snprintf(head, sizeof(head), "slab %d: ", first);
  return GroupFail(rc[first], head + err[first]);
}

// layout of slab q inside a rank-wide array (NULL = compact rank-wide array)
aphcg_layout SlabLayout(const aphcg_group* g, const aphcg_layout* l, int q) {
  aphcg_layout out;
  if (l) {
    out = *l;
  } else {
    out.offset = 0;
    out.stride_y = g->desc.nx;
    out.stride_z = g->desc.nx * g->desc.ny;
  }
  out.offset += g->z0[q] * out.stride_z;
  return out;
}

// Slabs that share a GPU (a device ordinal repeats: the single-GPU test configuration,
// `cuda_slabs_per_device` in the adapter) hand their scalars to each other through one-warp
// kernels that spin on a mailbox, so the slabs' streams must sit on different hardware queues:
// the driver has 8 by default and reads CUDA_DEVICE_MAX_CONNECTIONS when it creates the
// device's context.  Raise it while that is still possible, refuse with a clear message when
// it is not -- a deadlock on the GPU is the alternative.
int EnsureConnections(const int32_t* devices, int n) {
  int worst = 1, worst_dev = -1;
  for (int q = 0; q < n; ++q) {
    int rep = 0;
    for (int w = 0; w < n; ++w) rep += (devices[w] == devices[q]);
    if (rep > worst) {
      worst = rep;
      worst_dev = devices[q];
    }
  }
  if (worst == 1) return 0;
  const int need = 32;  // the driver's maximum; streams are spread over the queues
  const char* e = getenv("CUDA_DEVICE_MAX_CONNECTIONS");
  if (e && atoi(e) >= need) return 0;
  // is the device's primary context already there?
  using GetStateFn = CUresult (*)(CUdevice, unsigned int*, int*);
  GetStateFn get_state = nullptr;
  cudaDriverEntryPointQueryResult qres;
  int active = 1;  // unknown counts as active
  if (cudaGetDriverEntryPoint("cuDevicePrimaryCtxGetState", (void**)&get_state, cudaEnableDefault,
                              &qres) == cudaSuccess &&
      qres == cudaDriverEntryPointSuccess && get_state) {
    unsigned flags = 0;
    active = 0;
    for (int q = 0; q < n; ++q) {
      int a = 1;
      if (get_state((CUdevice)devices[q], &flags, &a) != CUDA_SUCCESS) a = 1;
      active |= a;
    }
  } else {
    cudaGetLastError();
  }
  if (!active) {
    setenv("CUDA_DEVICE_MAX_CONNECTIONS", "32", 1);
    return 0;
  }
  return GroupFail(APHCG_ERR_STATE,
                   std::to_string(worst) + " slabs share device " + std::to_string(worst_dev) +
                       ": this needs CUDA_DEVICE_MAX_CONNECTIONS=32 in the environment before "
                       "the process first touches the GPU (now: " +
                       (e ? std::string(e) : std::string("unset")) +
                       ", and the CUDA context already exists)");
}

void MergeInfo(const std::vector<aphcg_info>& v, aphcg_info* info) {
  if (!info) return;
  *info = v[0];  // residual and iter are bitwise the same on every slab
  for (const auto& i : v) {
    if (i.loop_ms > info->loop_ms) info->loop_ms = i.loop_ms;
    if (i.total_ms > info->total_ms) info->total_ms = i.total_ms;
  }
}

}  // namespace

extern "C" {

int aphcg_group_create(aphcg_group_t** out, const aphcg_desc* desc, const int32_t* devices,
                       int32_t ndevices) {
  if (!out || !desc || !devices) return GroupFail(APHCG_ERR_ARG, "null argument");
  *out = nullptr;
  if (ndevices < 1 || ndevices > kMaxRanks)
    return GroupFail(APHCG_ERR_ARG, "a group has 1.." + std::to_string(kMaxRanks) + " slabs");
  if (desc->nz < ndevices)
    return GroupFail(APHCG_ERR_ARG, "cannot cut " + std::to_string(desc->nz) + " planes into " +
                                        std::to_string(ndevices) + " slabs");
  if (int rc = EnsureConnections(devices, ndevices)) return rc;
  aphcg_group* g = new aphcg_group();
  g->n = ndevices;
  g->desc = *desc;
  g->desc.rank = 0;
  g->desc.nranks = ndevices;
  g->desc.z0 = 0;
  g->desc.nz_local = desc->nz;
  g->h.assign(ndevices, nullptr);
  if (ndevices > 1) g->gs = new GroupSync(ndevices);
  // contiguous planes, sizes differ by at most one, larger slabs first
  const int64_t base = desc->nz / ndevices, extra = desc->nz % ndevices;
  int64_t z = 0;
  for (int q = 0; q < ndevices; ++q) {
    g->z0.push_back(z);
    g->nzl.push_back(base + (q < extra ? 1 : 0));
    z += g->nzl.back();
  }
  auto fail = [&](int rc) {
    const std::string msg = aphcg_last_error();
    aphcg_group_destroy(g);
    return GroupFail(rc, msg);
  };
  for (int q = 0; q < ndevices; ++q) {
    aphcg_desc d = g->desc;
    d.device = devices[q];
    d.rank = q;
    d.z0 = g->z0[q];
    d.nz_local = g->nzl[q];
    if (int rc = aphcg_create(&g->h[q], &d)) return fail(rc);
    if (g->gs) AttachGroupSync(g->h[q], g->gs);
  }
  if (ndevices > 1) {
    std::vector<char> blobs((size_t)ndevices * APHCG_IPC_BYTES);
    for (int q = 0; q < ndevices; ++q)
      if (int rc = aphcg_ipc_export(g->h[q], blobs.data() + (size_t)q * APHCG_IPC_BYTES))
        return fail(rc);
    for (int q = 0; q < ndevices; ++q)
      if (int rc = aphcg_ipc_connect(g->h[q], blobs.data(), ndevices)) return fail(rc);
  }
  *out = g;
  return 0;
}

int aphcg_group_destroy(aphcg_group_t* g) {
  if (!g) return 0;
  for (aphcg_t* h : g->h) aphcg_destroy(h);
  delete g->gs;
  delete g;
  return 0;
}

int aphcg_group_size(aphcg_group_t* g) { return g ? g->n : 0; }

aphcg_t* aphcg_group_member(aphcg_group_t* g, int32_t slab) {
  return (g && slab >= 0 && slab < g->n) ? g->h[slab] : nullptr;
}

int aphcg_group_slab(aphcg_group_t* g, int32_t slab, int64_t* z0, int64_t* nz_local) {
  if (!g || slab < 0 || slab >= g->n) return GroupFail(APHCG_ERR_ARG, "bad slab index");
  if (z0) *z0 = g->z0[slab];
  if (nz_local) *nz_local = g->nzl[slab];
  return 0;
}

int aphcg_group_upload_system(aphcg_group_t* g, const double* system, const aphcg_layout* layout) {
  if (!g || !system) return GroupFail(APHCG_ERR_ARG, "null argument");
  return RunAll(g, [&](int q) {
    const aphcg_layout l = SlabLayout(g, layout, q);
    return aphcg_upload_system(g->h[q], system, &l);
  });
}

int aphcg_group_upload_guess(aphcg_group_t* g, const double* x0, const aphcg_layout* layout) {
  if (!g) return GroupFail(APHCG_ERR_ARG, "null group");
  return RunAll(g, [&](int q) {
    const aphcg_layout l = SlabLayout(g, layout, q);
    return aphcg_upload_guess(g->h[q], x0, x0 ? &l : nullptr);
  });
}

int aphcg_group_run(aphcg_group_t* g, const aphcg_conf* conf, aphcg_info* info) {
  if (!g || !conf) return GroupFail(APHCG_ERR_ARG, "null argument");
  std::vector<aphcg_info> v(g->n);
  if (int rc = RunAll(g, [&](int q) { return aphcg_run(g->h[q], conf, &v[q]); })) return rc;
  MergeInfo(v, info);
  return 0;
}

int aphcg_group_run_jacobi(aphcg_group_t* g, const aphcg_conf* conf, aphcg_info* info) {
  if (!g || !conf) return GroupFail(APHCG_ERR_ARG, "null argument");
  std::vector<aphcg_info> v(g->n);
  if (int rc = RunAll(g, [&](int q) { return aphcg_run_jacobi(g->h[q], conf, &v[q]); })) return rc;
  MergeInfo(v, info);
  return 0;
}

int aphcg_group_download_solution(aphcg_group_t* g, double* x, const aphcg_layout* layout) {
  if (!g || !x) return GroupFail(APHCG_ERR_ARG, "null argument");
  return RunAll(g, [&](int q) {
    const aphcg_layout l = SlabLayout(g, layout, q);
    return aphcg_download_solution(g->h[q], x, &l);
  });
}

int aphcg_group_solve(aphcg_group_t* g, const double* system, const aphcg_layout* system_layout,
                      const double* x0, const aphcg_layout* x0_layout, double* x,
                      const aphcg_layout* x_layout, const aphcg_conf* conf, aphcg_info* info) {
  if (!g || !system || !x || !conf) return GroupFail(APHCG_ERR_ARG, "null argument");
  std::vector<aphcg_info> v(g->n);
  const int rc = RunAll(g, [&](int q) {
    const aphcg_layout ls = SlabLayout(g, system_layout, q);
    const aphcg_layout l0 = SlabLayout(g, x0_layout, q);
    const aphcg_layout lx = SlabLayout(g, x_layout, q);
    return aphcg_solve(g->h[q], system, &ls, x0, x0 ? &l0 : nullptr, x, &lx, conf, &v[q]);
  });
  if (rc) return rc;
  MergeInfo(v, info);
  return 0;
}

int aphcg_group_assemble_spheres(aphcg_group_t* g, const double* spheres, int32_t nspheres,
                                 double rho_in, double rho_out, double dt) {
  if (!g) return GroupFail(APHCG_ERR_ARG, "null group");
  return RunAll(g, [&](int q) {
    return aphcg_assemble_spheres(g->h[q], spheres, nspheres, rho_in, rho_out, dt);
  });
}

int aphcg_group_assemble_projection(aphcg_group_t* g, const double* rho, const double* vx,
                                    const double* vy, const double* vz, const double* source,
                                    double dt, double hcell) {
  if (!g || !rho || !vx || !vy || !vz) return GroupFail(APHCG_ERR_ARG, "null argument");
  const int64_t nx = g->desc.nx, ny = g->desc.ny;
  return RunAll(g, [&](int q) {
    const int64_t z0 = g->z0[q];
    // slab q: density planes z0-1 .. z0+nzl (rank-wide plane -1 is the array's plane 0)
    return aphcg_assemble_projection(g->h[q], rho + z0 * nx * ny, vx + z0 * (nx + 1) * ny,
                                     vy + z0 * nx * (ny + 1), vz + z0 * nx * ny,
                                     source ? source + z0 * nx * ny : nullptr, dt, hcell);
  });
}

int aphcg_group_true_residual(aphcg_group_t* g, double* sum_r2) {
  if (!g || !sum_r2) return GroupFail(APHCG_ERR_ARG, "null argument");
  std::vector<double> v(g->n, 0.0);
  if (int rc = RunAll(g, [&](int q) { return aphcg_true_residual(g->h[q], &v[q]); })) return rc;
  double s = 0.0;
  for (double x : v) s += x;
  *sum_r2 = s;
  return 0;
}

}  // extern "C"
