// Microbenchmark: issue throughput of scalar FFMA vs packed FFMA2 (fma.rn.f32x2) on sm_100a.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ffma2_probe ffma2_probe.cu ; run on the GPU box.
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ unsigned long long pack2(float a, float b) {
  unsigned long long r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
  return r;
}
__device__ __forceinline__ void unpack2(unsigned long long v, float& a, float& b) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v));
}
__device__ __forceinline__ unsigned long long fma2(unsigned long long a, unsigned long long b, unsigned long long c) {
  unsigned long long r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}

template <int MODE> __global__ void probe(float* out, int iters, float s) {
  float acc[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) acc[i] = threadIdx.x * 0.001f + i;
  float x = s, y = s * 0.5f;
  if (MODE == 0) {
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int i = 0; i < 32; ++i) acc[i] = fmaf(acc[i], x, y);
    }
  } else if (MODE == 1) {  // packed, operands already paired
    unsigned long long a2[16], x2 = pack2(x, x), y2 = pack2(y, y);
#pragma unroll
    for (int i = 0; i < 16; ++i) a2[i] = pack2(acc[2 * i], acc[2 * i + 1]);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int i = 0; i < 16; ++i) a2[i] = fma2(a2[i], x2, y2);
    }
#pragma unroll
    for (int i = 0; i < 16; ++i) unpack2(a2[i], acc[2 * i], acc[2 * i + 1]);
  } else {  // packed accumulators, scalar multiplier re-paired each use (outer-product pattern)
    unsigned long long a2[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) a2[i] = pack2(acc[2 * i], acc[2 * i + 1]);
    for (int it = 0; it < iters; ++it) {
      float xs[4] = {x + it, x - it, x * it, x};
#pragma unroll
      for (int i = 0; i < 16; ++i) a2[i] = fma2(pack2(xs[i & 3], xs[i & 3]), pack2(y, x), a2[i]);
    }
#pragma unroll
    for (int i = 0; i < 16; ++i) unpack2(a2[i], acc[2 * i], acc[2 * i + 1]);
  }
  float r = 0;
#pragma unroll
  for (int i = 0; i < 32; ++i) r += acc[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}

template <int MODE> void run(const char* name, float* out, int sms) {
  const int iters = 20000;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  probe<MODE><<<sms * 2, 512>>>(out, 100, 1.0001f);
  cudaDeviceSynchronize();
  cudaEventRecord(e0);
  probe<MODE><<<sms * 2, 512>>>(out, iters, 1.0001f);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  const double fma = (double)sms * 2 * 512 * 32.0 * iters;
  printf("%-28s %8.3f ms  %7.2f TFMA/s  (%.1f fp32 TFLOP/s)\n", name, ms, fma / ms * 1e-9, 2 * fma / ms * 1e-9);
}

int main() {
  int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  float* out; cudaMalloc(&out, sms * 2 * 512 * sizeof(float));
  run<0>("FFMA scalar", out, sms);
  run<1>("FFMA2 packed", out, sms);
  run<2>("FFMA2 + repack multiplier", out, sms);
  return 0;
}
