// runlen.cu -- measurement tool (not product code).  Two questions that decide the intermediate
// layouts and the transport of the stage kernels:
//  (1) how fast can B200 HBM stream a 16-byte-element array when one side of a tile
//      transposition is contiguous and the other side consists of runs of R elements at stride W;
//  (2) with two GPUs: what do SM stores / loads on a peer-mapped buffer and the copy engines reach
//      over NVLink, for contiguous and for run-structured access.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o runlen runlen.cu && ./runlen
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>

constexpr int TILE = 8192;     // elements per tile (8 lines of 1024 complex doubles = 128 KB)
constexpr int LT = 13;
constexpr int THREADS = 512;
constexpr int E = TILE / THREADS;

// runs of 2^lr elements, 2^lw elements between the runs of a tile; lr == LT: contiguous
struct Side { int lr, lw; };

__device__ __forceinline__ long long tile_base(const Side s, long long tile) {
  if (s.lr >= LT) return tile << LT;
  const int lk = LT - s.lr, lpg = s.lw - s.lr;      // runs per tile, tiles per group
  const long long g = tile >> lpg, w = tile & ((1ll << lpg) - 1);
  return (g << (lk + s.lw)) + (w << s.lr);
}
__device__ __forceinline__ unsigned in_tile(const Side s, unsigned j) {
  if (s.lr >= LT) return j;
  return ((j >> s.lr) << s.lw) + (j & ((1u << s.lr) - 1));
}

// half_remote: tiles with odd index write to out2 instead of out (a stage whose chunks go to two ranks)
__global__ void __launch_bounds__(THREADS, 2) copy_runs(const double2 *__restrict__ in, double2 *__restrict__ out,
                                                        double2 *__restrict__ out2, long long ntiles, Side si, Side so, int split = 0) {
  unsigned oi[E], oo[E];
#pragma unroll
  for (int e = 0; e < E; e++) {
    oi[e] = in_tile(si, e * THREADS + threadIdx.x);
    oo[e] = in_tile(so, e * THREADS + threadIdx.x);
  }
  for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const double2 *pi = in + tile_base(si, tile);
    double2 *po = ((tile & 1) && out2 && !split ? out2 : out) + tile_base(so, tile);
    double2 *po2 = split ? out2 + tile_base(so, tile) : po;   // split: the upper half of every tile goes to out2
    double2 x[E];
#pragma unroll
    for (int e = 0; e < E; e++) x[e] = pi[oi[e]];
#pragma unroll
    for (int e = 0; e < E; e++) (e < E / 2 ? po : po2)[oo[e]] = x[e];
  }
}

static cudaEvent_t e0, e1;
static double run(const double2 *in, double2 *out, double2 *out2, long long n, Side si, Side so, int grid) {
  float best = 1e30f;
  for (int rep = 0; rep < 4; rep++) {
    cudaEventRecord(e0);
    copy_runs<<<grid, THREADS>>>(in, out, out2, n >> LT, si, so);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    if (rep > 0 && ms < best) best = ms;
  }
  return 2.0 * n * 16 / best * 1e-6;   // read + write GB/s
}

int main() {
  const long long n = 1ll << 27;   // 2 GiB per array
  int ndev = 0;
  cudaGetDeviceCount(&ndev);
  cudaSetDevice(0);
  double2 *a, *b;
  cudaMalloc(&a, n * sizeof(double2));
  cudaMalloc(&b, n * sizeof(double2));
  cudaMemset(a, 1, n * sizeof(double2));
  cudaMemset(b, 0, n * sizeof(double2));
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  const Side C{LT, 0};
  printf("== local HBM, read+write GB/s (persistent grid 2/SM | one CTA per tile)\n");
  printf("%-28s %10.1f %10.1f\n", "contiguous -> contiguous", run(a, b, nullptr, n, C, C, sms * 2), run(a, b, nullptr, n, C, C, (int)(n >> LT)));
  // stride 128 KB (runs of neighbouring tiles are adjacent: blocked layouts), or one group spanning the
  // whole array (stride = n / runs per tile: the plain [k][everything else] layout)
  for (int mode = 0; mode < 2; mode++)
    for (int lr = 10; lr >= 2; lr--) {
      if (lr > 6 && lr != 10) continue;
      const int lw = mode == 0 ? 13 : 14 + lr;
      char nm[64];
      Side s{lr, lw};
      snprintf(nm, sizeof nm, "write runs %4d B stride 2^%d", 16 << lr, lw + 4);
      const double w = run(a, b, nullptr, n, C, s, sms * 2);
      const double r = run(a, b, nullptr, n, s, C, sms * 2);
      const double rw = run(a, b, nullptr, n, s, s, sms * 2);
      printf("%-28s %10.1f   (read runs instead: %8.1f, both: %8.1f)\n", nm, w, r, rw);
    }
  if (ndev >= 2) {
    int can = 0;
    cudaDeviceCanAccessPeer(&can, 0, 1);
    printf("== two GPUs, peer access 0->1: %d\n", can);
    cudaDeviceEnablePeerAccess(1, 0);
    cudaSetDevice(1);
    cudaDeviceEnablePeerAccess(0, 0);
    double2 *r;
    cudaMalloc(&r, n * sizeof(double2));
    cudaMemset(r, 0, n * sizeof(double2));
    cudaDeviceSynchronize();
    cudaSetDevice(0);
    // copy engine
    for (int bidir = 0; bidir < 2; bidir++) {
      cudaStream_t s0, s1;
      cudaStreamCreate(&s0);
      cudaSetDevice(1);
      cudaStreamCreate(&s1);
      cudaSetDevice(0);
      float best = 1e30f;
      for (int rep = 0; rep < 3; rep++) {
        cudaDeviceSynchronize();
        cudaEventRecord(e0, s0);
        cudaMemcpyPeerAsync(r, 1, a, 0, n * 16, s0);
        if (bidir) cudaMemcpyPeerAsync(b, 0, r + n / 2, 1, n * 8, s1);
        cudaEventRecord(e1, s0);
        cudaEventSynchronize(e1);
        cudaSetDevice(1);
        cudaDeviceSynchronize();
        cudaSetDevice(0);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        if (rep > 0 && ms < best) best = ms;
      }
      printf("copy engine 0->1 %s: %8.1f GB/s one direction\n", bidir ? "(with 1->0 traffic)" : "", n * 16 / best * 1e-6);
    }
    printf("SM kernel on GPU 0, GB/s of NVLink payload (one direction):\n");
    for (int grid : {sms / 2, sms, sms * 2}) {
      printf(" grid %3d: write remote contiguous %8.1f | read remote contiguous %8.1f\n", grid,
             run(a, r, nullptr, n, C, C, grid) / 2, run(r, b, nullptr, n, C, C, grid) / 2);
    }
    for (int lr = 6; lr >= 2; lr--) {
      Side s{lr, 13}, s2{lr, 14 + lr};
      printf(" write remote runs %4d B: stride 128 KB %8.1f | whole-array stride %8.1f | read remote runs (128 KB) %8.1f\n", 16 << lr,
             run(a, r, nullptr, n, C, s, sms * 2) / 2, run(a, r, nullptr, n, C, s2, sms * 2) / 2, run(r, b, nullptr, n, s, C, sms * 2) / 2);
    }
    printf(" half local / half remote, total read+write GB/s seen by GPU 0: contiguous %8.1f | runs 128 B %8.1f | runs 512 B %8.1f\n",
           run(a, b, r, n, C, C, sms * 2), run(a, b, r, n, C, Side{3, 17}, sms * 2), run(a, b, r, n, C, Side{5, 13}, sms * 2));
  }
  if (ndev >= 2) {
    // both GPUs at once (what a stage in front of an exchange does on every rank)
    cudaSetDevice(1);
    double2 *a1, *b1, *r0;
    cudaMalloc(&a1, n * sizeof(double2));
    cudaMalloc(&b1, n * sizeof(double2));
    cudaMemset(a1, 1, n * sizeof(double2));
    cudaEvent_t f0, f1;
    cudaEventCreate(&f0);
    cudaEventCreate(&f1);
    cudaSetDevice(0);
    cudaMalloc(&r0, n * sizeof(double2));
    double2 *r1;
    cudaSetDevice(1);
    cudaMalloc(&r1, n * sizeof(double2));
    cudaDeviceSynchronize();
    cudaSetDevice(0);
    cudaDeviceSynchronize();
    struct Mode { const char *name; bool remote0, remote1; Side so; };
    const Mode modes[] = {{"GPU0 half remote (128 B runs), GPU1 idle", true, false, Side{3, 17}},
                          {"GPU0 half remote, GPU1 local copy", true, true, Side{3, 17}},
                          {"both half remote, 128 B runs", true, true, Side{3, 17}},
                          {"both half remote, 512 B runs", true, true, Side{5, 13}},
                          {"both half remote, contiguous", true, true, C},
                          {"both, upper half of every tile remote, 128 B", true, true, Side{3, 17}}};
    for (int m = 0; m < 6; m++) {
      float best0 = 1e30f, best1 = 1e30f;
      for (int rep = 0; rep < 4; rep++) {
        cudaSetDevice(0);
        cudaDeviceSynchronize();
        cudaSetDevice(1);
        cudaDeviceSynchronize();
        cudaSetDevice(0);
        cudaEventRecord(e0);
        copy_runs<<<sms * 2, THREADS>>>(a, b, r1, n >> LT, C, modes[m].so, m == 5);
        cudaEventRecord(e1);
        if (m > 0) {
          cudaSetDevice(1);
          cudaEventRecord(f0);
          copy_runs<<<sms * 2, THREADS>>>(a1, b1, m == 1 ? nullptr : r0, n >> LT, C, modes[m].so, m == 5);
          cudaEventRecord(f1);
        }
        cudaSetDevice(0);
        cudaEventSynchronize(e1);
        float ms0 = 0, ms1 = 0;
        cudaEventElapsedTime(&ms0, e0, e1);
        if (m > 0) {
          cudaSetDevice(1);
          cudaEventSynchronize(f1);
          cudaEventElapsedTime(&ms1, f0, f1);
          cudaSetDevice(0);
        }
        if (rep > 0) { best0 = ms0 < best0 ? ms0 : best0; best1 = ms1 < best1 ? ms1 : best1; }
      }
      printf(" %-44s GPU0 %6.3f ms (%6.1f GB/s over NVLink)  GPU1 %6.3f ms\n", modes[m].name, best0, n * 8 / best0 * 1e-6, best1);
    }
  }
  cudaError_t err = cudaDeviceSynchronize();
  printf("status: %s\n", cudaGetErrorString(err));
  return 0;
}
