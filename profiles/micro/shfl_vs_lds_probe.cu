// Microbenchmark: do warp shuffles share the shared-memory data pipe (the resource that bounds sample_elements_kernel)?
// A: broadcast LDS.128 only, B: SHFL.IDX only, C: both interleaved.  If time(C) ~ max(A, B) the pipes are separate and the
// Cholesky's column broadcast could move from shared memory to shuffles; if time(C) ~ A + B they are one pipe.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o shfl_vs_lds_probe shfl_vs_lds_probe.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE> __global__ void __launch_bounds__(448) probe(float* out, int iters) {
  __shared__ __align__(16) float buf[1024];
  for (int i = threadIdx.x; i < 1024; i += blockDim.x) buf[i] = i * 0.001f;
  __syncthreads();
  float acc0 = threadIdx.x, acc1 = 0.f, acc2 = 0.f, acc3 = 0.f;
  const int lane = threadIdx.x & 31;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < 16; ++u) {
      if (MODE == 0 || MODE == 2) {
        const float4 v = *reinterpret_cast<const float4*>(buf + ((it + u) & 63) * 4);  // broadcast: one address per warp
        acc0 += v.x; acc1 += v.y; acc2 += v.z; acc3 += v.w;
      }
      if (MODE == 1 || MODE == 2) {
        acc0 += __shfl_sync(0xffffffffu, acc1, (u + it) & 31);
        acc1 += __shfl_sync(0xffffffffu, acc2, (u + 1) & 31);
        acc2 += __shfl_sync(0xffffffffu, acc3, (u + 2) & 31);
        acc3 += __shfl_sync(0xffffffffu, acc0, (u + 3) & 31);
      }
    }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc0 + acc1 + acc2 + acc3 + lane;
}

template <int MODE> float run(float* out, int sms) {
  const int iters = 4000;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  probe<MODE><<<sms, 448>>>(out, 10);
  cudaDeviceSynchronize();
  cudaEventRecord(e0);
  probe<MODE><<<sms, 448>>>(out, iters);
  cudaEventRecord(e1);
  cudaDeviceSynchronize();
  float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
  return ms;
}

int main() {
  int sms = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  float* out; cudaMalloc(&out, (size_t)sms * 448 * 4);
  const float a = run<0>(out, sms), b = run<1>(out, sms), c = run<2>(out, sms);
  // per SM: 14 warps x 4000 x 16 = 896000 warp-level LDS.128 (A) / 4 x that many SHFL (B)
  const double n = 14.0 * 4000 * 16;
  printf("{\"lds128_broadcast_ms\": %.3f, \"shfl_ms\": %.3f, \"both_ms\": %.3f, \"lds128_per_clk_per_sm\": %.3f, "
         "\"shfl_per_clk_per_sm\": %.3f, \"clock_mhz_assumed\": 1965, \"cuda_error\": \"%s\"}\n",
         a, b, c, n / (a * 1e-3 * 1.965e9), 4 * n / (b * 1e-3 * 1.965e9), cudaGetErrorString(cudaGetLastError()));
  return 0;
}
