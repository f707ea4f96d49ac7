// Which scheduler (SMSP) does a warp land on? A dependent integer chain issues at most one instruction per ~4 cycles per
// warp; several such warps on ONE scheduler still fit (it issues one instruction per cycle), but a chain of INDEPENDENT
// instructions saturates a scheduler with one warp, so two warps on one scheduler take twice as long.
// build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o smsp_probe smsp_probe.cu
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(unsigned* out, unsigned mask, int iters) {
  const unsigned w = threadIdx.x >> 5;
  if (!((mask >> w) & 1u)) return;
  unsigned a = threadIdx.x, b = a * 3 + 1, c = a ^ 5, d = a + 7, e = a * 5, f = a ^ 9, g = a + 11, h = a * 7;
#pragma unroll 1
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int u = 0; u < 8; u++) {
      a = a * 3 + b; b = b * 5 + c; c = c * 7 + d; d = d * 9 + e; e = e * 11 + f; f = f * 13 + g; g = g * 15 + h; h = h * 17 + a;
    }
  }
  if ((a ^ b ^ c ^ d ^ e ^ f ^ g ^ h) == 0x12345u) out[0] = a;
}
static float run(int ctasPerSm, int threads, unsigned mask, int smem) {
  int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  unsigned* d; cudaMalloc(&d, 4);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<<<sms * ctasPerSm, threads, smem>>>(d, mask, 2000);
  cudaEventRecord(e0);
  k<<<sms * ctasPerSm, threads, smem>>>(d, mask, 20000);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1); cudaFree(d);
  return ms;
}
int main() {
  // single-warp CTAs, n per SM (80 KiB of shared memory each would allow 2; use small smem so that up to 8 fit)
  for (int n = 1; n <= 8; n++) printf("1-warp CTAs, %d per SM: %.3f ms\n", n, run(n, 32, 1u, 1024));
  // one CTA per SM, chosen warps busy
  const unsigned masks[] = {0x1, 0x3, 0x11, 0x5, 0xF, 0x1111, 0x33, 0xFF};
  for (unsigned m : masks) printf("1 CTA of 16 warps per SM, busy warps mask 0x%04x: %.3f ms\n", m, run(1, 512, m, 1024));
  // two single-warp CTAs per SM with 80 KiB of shared memory each (the sequence kernel's shape)
  printf("1-warp CTAs with 80 KiB smem, 2 per SM: %.3f ms\n", run(2, 32, 1u, 80 * 1024));
  printf("1-warp CTAs with 80 KiB smem, 1 per SM: %.3f ms\n", run(1, 32, 1u, 80 * 1024));
  return 0;
}
