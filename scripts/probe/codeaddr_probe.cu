#include <cstdio>
#include <cuda_runtime.h>
__device__ __noinline__ void anchor() { asm volatile("" ::: "memory"); }
__global__ void k(unsigned long long* out, int mode) {
  anchor();
  void (*fp)() = anchor;
  unsigned long long a = (unsigned long long)fp;
  out[0] = a;
  if (mode & 1) { unsigned v; asm volatile("ld.global.u32 %0, [%1];" : "=r"(v) : "l"(a) : "memory"); out[1] = v; }
  if (mode & 2) { asm volatile("prefetch.global.L2 [%0];" ::"l"(a)); asm volatile("prefetch.global.L2 [%0];" ::"l"(a - (1ull << 20))); }
  if (mode & 4) { asm volatile("prefetch.global.L2 [%0];" ::"l"(0x10ull)); asm volatile("prefetch.global.L2 [%0];" ::"l"(0x7f0000000000ull)); }
}
int main() {
  unsigned long long* d; cudaMalloc(&d, 64); cudaMemset(d, 0, 64);
  for (int mode : {0, 2, 4, 1}) {
    k<<<1, 32>>>(d, mode);
    cudaError_t e = cudaDeviceSynchronize();
    unsigned long long h[2] = {0, 0}; cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
    printf("mode %d: %s  anchor=%llx word=%llx\n", mode, cudaGetErrorString(e), h[0], h[1]);
    if (e != cudaSuccess) break;
  }
  return 0;
}
