// Feasibility probe for a cross-layer persistent UNet kernel: what does one "op boundary" cost inside a cooperative
// persistent kernel (grid barrier between dependent ops) compared with a kernel boundary in a CUDA graph (9.6 us for a
// conv launch, ~5 us for an elementwise launch, profiles/README.md)?
// Each iteration: every CTA writes a few cache lines another CTA reads in the next iteration, then a grid barrier
// (monotonic arrival counter, ld.acquire spin).  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o gridbar_probe gridbar_probe.cu
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdint.h>

__device__ __forceinline__ void grid_barrier(int* counter, int target) {
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    atomicAdd(counter, 1);
    int v;
    uint32_t spins = 0;
    do {
      asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(counter) : "memory");
      if (++spins > (1u << 26)) __trap();
    } while (v < target);
  }
  __syncthreads();
}

__global__ void __launch_bounds__(352, 1) probe(int* counter, float* buf, int iters, int work) {
  const int G = gridDim.x;
  float acc = 0.f;
  for (int it = 0; it < iters; ++it) {
    // consume what the neighbour produced in the previous iteration, produce for the next
    const int src = (blockIdx.x + 1) % G;
    for (int w = 0; w < work; ++w) acc += __ldcg(buf + ((size_t)(it & 1) * G + src) * 1024 + (threadIdx.x + w * 352) % 1024);
    for (int w = 0; w < work; ++w) __stcg(buf + ((size_t)((it + 1) & 1) * G + blockIdx.x) * 1024 + (threadIdx.x + w * 352) % 1024, acc + it);
    grid_barrier(counter, (it + 1) * G);
  }
  if (acc == 123.456f) buf[0] = acc;
}

int main() {
  int dev = 0, sms = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  int* counter; float* buf;
  cudaMalloc(&counter, 4); cudaMalloc(&buf, (size_t)2 * sms * 1024 * 4);
  cudaMemset(buf, 0, (size_t)2 * sms * 1024 * 4);
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  for (int work = 1; work <= 4; work *= 4) {
    for (int iters : {100, 1000}) {
      cudaMemset(counter, 0, 4);
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3(sms); cfg.blockDim = dim3(352); cfg.dynamicSmemBytes = 200 * 1024; cfg.stream = 0;
      cudaLaunchAttribute attr[1];
      attr[0].id = cudaLaunchAttributeCooperative; attr[0].val.cooperative = 1;
      cfg.attrs = attr; cfg.numAttrs = 1;
      cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
      cudaEventRecord(e0);
      cudaError_t e = cudaLaunchKernelEx(&cfg, probe, counter, buf, iters, work);
      cudaEventRecord(e1);
      cudaError_t s = cudaDeviceSynchronize();
      float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
      printf("grid %d x 352 threads, work %d, %4d op boundaries: %.1f us total, %.2f us per boundary (launch %s, sync %s)\n", sms, work, iters,
             ms * 1e3, ms * 1e3 / iters, cudaGetErrorString(e), cudaGetErrorString(s));
    }
  }
  return 0;
}
