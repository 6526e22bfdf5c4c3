// Micro-benchmark of KMomentsSerial on synthetic (g, y): isolates the serial chain from its producers.
// dbg bit 0: no chain; bit 1: no addend production; bit 2: no staging copies.
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../../cauchyfriendly_b200/csrc/backend_cuda.cuh"
#include "../../cauchyfriendly_b200/csrc/mce_kern_prop.h"
using namespace mce;
int main(int argc, char** argv) {
  const long long n = argc > 1 ? atoll(argv[1]) : 1127690; const int d = 7, nq = 1 + d + d * d;
  std::vector<double> hg(2 * n), hy(2 * d * n);
  srand(1);
  for (auto& v : hg) v = rand() / (double)RAND_MAX - 0.5;
  for (auto& v : hy) v = rand() / (double)RAND_MAX - 0.5;
  double *g, *y, *out;
  cudaMalloc(&g, hg.size() * 8); cudaMalloc(&y, hy.size() * 8); cudaMalloc(&out, 2 * nq * 8 + 64);
  cudaMemcpy(g, hg.data(), hg.size() * 8, cudaMemcpyHostToDevice); cudaMemcpy(y, hy.data(), hy.size() * 8, cudaMemcpyHostToDevice);
  CudaBackend be; std::string why; if (!be.init(0, &why)) { printf("%s\n", why.c_str()); return 1; }
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int dbg : {0, 1, 2, 4, 3, 6, 7}) {
    KMomentsSerial k{(const cplx*)g, y, n, d, out, dbg};
    for (int rep = 0; rep < 3; rep++) {
      cudaEventRecord(e0, be.stream);
      be.launch(k, nq, 512, KMomentsSerial::smem_bytes(d));
      cudaEventRecord(e1, be.stream);
      cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      if (rep == 2) printf("dbg %d: %.3f ms  (%.2f ns/slot)\n", dbg, ms, ms * 1e6 / n);
    }
  }
  return 0;
}
