// Microbenchmark (measurement aid, not product code): what HBM bandwidth does the ACCESS
// PATTERN of the direction kernel reach when nothing else is in the way -- no shared-memory
// staging, no barriers, no arithmetic?  R read streams and W write streams over 512^3 doubles
// each, walked exactly like k_dir_spmv_tma walks them: a CTA owns a TX x TY column of cells and
// marches through ZC planes; per plane it touches TY row segments of TX*8 contiguous bytes in
// every array.  Compared with the same bytes walked as one flat stream per array.
//
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o scripts/stream_bench scripts/stream_bench.cu
//   scripts/stream_bench            (prints a table; GB/s = (R+W) * 8 B * cells / time)
#include <cstdio>
#include <cstdlib>
#include <vector>

#include <cuda_runtime.h>

#define CK(x)                                                                      \
  do {                                                                             \
    cudaError_t e = (x);                                                           \
    if (e != cudaSuccess) {                                                        \
      printf("%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e));             \
      exit(1);                                                                     \
    }                                                                              \
  } while (0)

constexpr int N = 512;
struct Ptrs {
  const double2* rd[8];
  double2* wr[4];
};

// tile walk: grid = (N/TX, N/TY, N/ZC), block = TX/2 threads x TY rows (flattened)
template <int R, int W>
__global__ void k_tiles(Ptrs p, int TX, int TY, int ZC) {
  const int lxn = TX / 2;
  const int lx = threadIdx.x % lxn, ly = threadIdx.x / lxn;
  const int rows_per_thread = TY / (blockDim.x / lxn);
  const int x = blockIdx.x * TX / 2 + lx;  // in double2 units
  const int k0 = blockIdx.z * ZC;
  double acc = 0;
  for (int k = k0; k < k0 + ZC; ++k) {
    for (int rr = 0; rr < rows_per_thread; ++rr) {
      const int y = blockIdx.y * TY + ly * rows_per_thread + rr;
      const size_t i = ((size_t)k * N + y) * (N / 2) + x;
      double2 v[R];
#pragma unroll
      for (int s = 0; s < R; ++s) v[s] = __ldcs(p.rd[s] + i);
      double t = 0;
#pragma unroll
      for (int s = 0; s < R; ++s) t += v[s].x + v[s].y;
      acc += t;
#pragma unroll
      for (int s = 0; s < W; ++s) __stcs(p.wr[s] + i, make_double2(t, acc));
    }
  }
  if (acc == 1.2345e-300) p.wr[0][0] = make_double2(acc, acc);
}

// flat walk: persistent grid-stride over double2 elements, 4 independent elements per thread
template <int R, int W>
__global__ void k_flat(Ptrs p, size_t n2) {
  double acc = 0;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n2; i += 4 * stride) {
    double2 v[4][R];
#pragma unroll
    for (int u = 0; u < 4; ++u)
#pragma unroll
      for (int s = 0; s < R; ++s)
        if (i + u * stride < n2) v[u][s] = __ldcs(p.rd[s] + i + u * stride);
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      if (i + u * stride >= n2) break;
      double t = 0;
#pragma unroll
      for (int s = 0; s < R; ++s) t += v[u][s].x + v[u][s].y;
      acc += t;
#pragma unroll
      for (int s = 0; s < W; ++s) __stcs(p.wr[s] + i + u * stride, make_double2(t, acc));
    }
  }
  if (acc == 1.2345e-300) p.wr[0][0] = make_double2(acc, acc);
}

template <int R, int W>
void run(Ptrs p, const char* name) {
  const size_t cells = (size_t)N * N * N;
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  auto time = [&](auto launch) {
    launch();
    launch();
    CK(cudaEventRecord(e0));
    for (int i = 0; i < 5; ++i) launch();
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    CK(cudaGetLastError());
    float ms;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    return ms / 5;
  };
  const double bytes = (double)(R + W) * 8.0 * cells;
  float ms = time([&] { k_flat<R, W><<<148 * 8, 256>>>(p, cells / 2); });
  printf("%-22s flat stream                         %7.3f ms  %7.1f GB/s\n", name, ms, bytes / ms / 1e6);
  const int shapes[][3] = {{128, 8, 32}, {128, 8, 64}, {256, 4, 32}, {256, 8, 32}, {512, 2, 32},
                           {512, 4, 32}, {512, 8, 32}, {64, 8, 32}, {128, 16, 32}};
  for (auto& s : shapes) {
    const int TX = s[0], TY = s[1], ZC = s[2];
    int threads = TX / 2 * TY;
    while (threads > 256) threads /= 2;  // several rows per thread
    dim3 grid(N / TX, N / TY, N / ZC);
    ms = time([&] { k_tiles<R, W><<<grid, threads>>>(p, TX, TY, ZC); });
    printf("%-22s tile %3dx%-2d planes/CTA %2d threads %3d  %7.3f ms  %7.1f GB/s\n", name, TX, TY, ZC,
           threads, ms, bytes / ms / 1e6);
  }
}

int main() {
  const size_t cells = (size_t)N * N * N;
  Ptrs p{};
  std::vector<double*> bufs;
  for (int i = 0; i < 12; ++i) {
    double* b;
    CK(cudaMalloc(&b, cells * 8));
    CK(cudaMemset(b, 0, cells * 8));
    bufs.push_back(b);
  }
  for (int i = 0; i < 8; ++i) p.rd[i] = (const double2*)bufs[i];
  for (int i = 0; i < 4; ++i) p.wr[i] = (double2*)bufs[8 + i];
  run<1, 1>(p, "copy 1R+1W");
  run<2, 1>(p, "update 2R+1W");
  run<6, 2>(p, "dir odd 6R+2W");
  run<8, 3>(p, "dir even 8R+3W");
  return 0;
}
