// setup_probe.cu -- where does engine set-up time go?  (diagnostic, run on the GPU box)
//   nvcc -O2 -o /tmp/setup_probe tools/setup_probe.cu -Iinclude -Llbzip2_b200 -lbz2b200 -Xlinker -rpath=$PWD/lbzip2_b200
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include "lbzip2_b200.h"
static double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
int main(int argc, char **argv) {
  const int chunks = argc > 1 ? atoi(argv[1]) : 32;
  double t = now();
  cudaFree(0);
  printf("context: %.3f s\n", now() - t);
  t = now(); void *p[8];
  for (int i = 0; i < 8; i++) cudaMalloc(&p[i], 512u << 20);
  printf("8 x cudaMalloc(512 MB): %.3f s\n", now() - t);
  t = now(); for (int i = 0; i < 8; i++) cudaFree(p[i]);
  printf("8 x cudaFree: %.3f s\n", now() - t);
  t = now(); void *h; cudaHostAlloc(&h, 64u << 20, cudaHostAllocDefault);
  printf("cudaHostAlloc(64 MB): %.3f s\n", now() - t);
  t = now(); cudaFreeHost(h);
  printf("cudaFreeHost: %.3f s\n", now() - t);
  for (int rep = 0; rep < 2; rep++) {
    t = now(); lbz_engine *e = lbz_engine_create(0, 9, chunks);
    printf("lbz_engine_create(%d chunks): %.3f s (%zu MB device)\n", chunks, now() - t, lbz_engine_device_bytes(e) >> 20);
    t = now(); lbz_engine_destroy(e);
    printf("lbz_engine_destroy: %.3f s\n", now() - t);
  }
  return 0;
}
